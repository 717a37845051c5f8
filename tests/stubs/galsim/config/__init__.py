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
