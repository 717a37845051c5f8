REGISTRY = {"photon_op": {}, "input": {}, "value": {}, "sensor": {}}


class PhotonOpBuilder:
    def buildPhotonOp(self, config, base, logger):
        raise NotImplementedError


class InputLoader:
    def __init__(self, init_func, takes_logger=False, **kw):
        self.init_func, self.takes_logger = init_func, takes_logger


def RegisterPhotonOpType(name, builder, input_type=None):
    REGISTRY["photon_op"][name] = (builder, input_type)


def RegisterInputType(name, loader):
    REGISTRY["input"][name] = loader


def RegisterValueType(name, func, types, input_type=None):
    REGISTRY["value"][name] = (func, types, input_type)


def GetAllParams(config, base, req=None, opt=None, single=None, ignore=()):
    req, opt = req or {}, opt or {}
    missing = [k for k in req if k not in config]
    if missing:
        raise KeyError("missing required parameters %s" % missing)
    unknown = [k for k in config if k not in req and k not in opt and k != "type" and k not in ignore]
    if unknown:
        raise KeyError("unexpected parameters %s" % unknown)
    return {k: v for k, v in config.items() if k != "type"}, True


def GetInputObj(name, config, base, where):
    return base["_input_objs"][name]


def GetRNG(config, base, logger=None, tag=None):
    from .. import BaseDeviate

    return base.get("rng", BaseDeviate(1234))


from . import sensor, util  # noqa: E402,F401


# ---- image / stamp types and the two helpers the pooled builder calls (round 2) ------------------------------------
REGISTRY.update({"image": {}, "stamp": {}})


def RegisterImageType(name, builder):
    REGISTRY["image"][name] = builder


def RegisterStampType(name, builder):
    REGISTRY["stamp"][name] = builder


def BuildPhotonOps(config, key, base, logger):
    """list of dicts -> photon ops: registered types through their builders, GalSim's own by class name"""
    import galsim

    ops = []
    for cfg in config.get(key, []):
        t = cfg["type"]
        if t in REGISTRY["photon_op"]:
            ops.append(REGISTRY["photon_op"][t][0].buildPhotonOp(cfg, base, logger))
        else:
            ops.append(getattr(galsim, t)(**{k: v for k, v in cfg.items() if k != "type"}))
    return ops


def FlattenNoiseVariance(base, full_image, stamps, current_vars, logger):
    return 0.0
