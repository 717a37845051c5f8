"""GalSim config registration of the B200 hot path.

List this module after ``imsim`` in a config's ``modules:`` and the same YAML
runs with the photon-shooting path swapped:

    modules: [imsim, imsim_b200.galsim_plugin]

It re-registers, under the reference's own names,

* photon ops ``RubinOptics``, ``RubinDiffractionOptics``, ``RubinDiffraction``
  (imsim/photon_ops.py:400-451),
* the sensor type ``Silicon`` (galsim.config sensor builder used by
  config/imsim-config.yaml:230-235),
* input ``tree_rings`` and value types ``TreeRingCenter`` / ``TreeRingFunc``
  (imsim/treerings.py:220-243).

``LSST_PhotonPoolingImage`` and ``LSST_Photons`` stay imSim's own builders: with the
ops and the sensor swapped, their hot loop (imsim/photon_pooling.py:141-160) already
runs on the device; ``B200SiliconSensor`` subclasses ``galsim.SiliconSensor`` so the
``isinstance`` test at photon_pooling.py:209 keeps passing ``recalc``.

GalSim, batoid and the LSST stack are not available in the build container: this
module is import-guarded and UNTESTED there.
"""
from __future__ import annotations

import galsim
from galsim.config import (GetAllParams, GetInputObj, InputLoader, PhotonOpBuilder, RegisterInputType,
                           RegisterPhotonOpType, RegisterValueType)
from galsim.config.util import get_cls_params

from . import photon_ops as _ops
from .diffraction import RUBIN_LATITUDE
from .sensor import SiliconSensor as _B200Sensor
from .treerings import TreeRings

_TYPES = {"CelestialCoord": galsim.CelestialCoord, "Angle": galsim.Angle, "PositionD": galsim.PositionD}


def _resolve(params):
    return {k: _TYPES.get(v, v) if isinstance(v, str) else v for k, v in params.items()}


for _cls in (_ops.RubinOptics, _ops.RubinDiffractionOptics, _ops.RubinDiffraction):
    _cls._req_params = _resolve(_cls._req_params)
    _cls._opt_params = _resolve(_cls._opt_params)


def photon_op_type(identifier: str, input_type=None):
    """Same decorator as imsim/photon_ops.py:361-386."""

    def decorator(deserializer):
        class Factory(PhotonOpBuilder):
            def buildPhotonOp(self, config, base, logger):
                return deserializer(config, base, logger)

        RegisterPhotonOpType(identifier, Factory(), input_type=input_type)
        return deserializer

    return decorator


def config_kwargs(config, base, cls, base_args=()):
    req, opt, single, _takes_rng = get_cls_params(cls)
    kwargs, _safe = GetAllParams(config, base, req, opt, single)
    kwargs.update({key: base[key] for key in base_args})
    return kwargs


def _camera(name):
    from imsim.camera import get_camera  # the reference's cached camera lookup

    return get_camera(name)


def _device(base):
    """One GPU per worker process: det_num % n_gpus (GalSim forks workers per output file)."""
    import torch

    n = max(torch.cuda.device_count(), 1)
    return int(base.get("det_num", base.get("file_num", 0))) % n


_rubin_optics_base_args = ("stamp_center",)


@photon_op_type("RubinOptics", input_type="telescope")
def deserialize_rubin_optics(config, base, _logger):
    kwargs = config_kwargs(config, base, _ops.RubinOptics, base_args=_rubin_optics_base_args)
    return _ops.RubinOptics(telescope=base["det_telescope"], icrf_to_field=base["_icrf_to_field"],
                            img_wcs=base["current_image"].wcs, camera=_camera(kwargs.pop("camera")),
                            device=_device(base), **kwargs)


@photon_op_type("RubinDiffractionOptics", input_type="telescope")
def deserialize_rubin_diffraction_optics(config, base, _logger):
    kwargs = config_kwargs(config, base, _ops.RubinDiffractionOptics, _rubin_optics_base_args)
    telescope = base["det_telescope"]
    rubin_diffraction = _ops.RubinDiffraction(
        telescope=telescope, latitude=kwargs.pop("latitude", RUBIN_LATITUDE), altitude=kwargs.pop("altitude"),
        azimuth=kwargs.pop("azimuth"), img_wcs=base["current_image"].wcs, icrf_to_field=base["_icrf_to_field"],
        disable_field_rotation=kwargs.pop("disable_field_rotation", False), device=_device(base))
    return _ops.RubinDiffractionOptics(telescope=telescope, camera=_camera(kwargs.pop("camera")),
                                       rubin_diffraction=rubin_diffraction, device=_device(base), **kwargs)


@photon_op_type("RubinDiffraction", input_type="telescope")
def deserialize_rubin_diffraction(config, base, _logger):
    kwargs = config_kwargs(config, base, _ops.RubinDiffraction)
    return _ops.RubinDiffraction(telescope=base["det_telescope"], icrf_to_field=base["_icrf_to_field"],
                                 img_wcs=base["current_image"].wcs, device=_device(base), **kwargs)


class B200SiliconSensor(_B200Sensor, galsim.SiliconSensor):
    """The device sensor, recognisable as a ``galsim.SiliconSensor`` (photon_pooling.py:209)."""

    def __init__(self, *args, **kwargs):
        _B200Sensor.__init__(self, *args, **kwargs)  # galsim's __init__ (C++ Silicon) is not run


class _SiliconSensorBuilder(galsim.config.sensor.SensorBuilder):
    def buildSensor(self, config, base, logger):
        opt = {"name": str, "strength": float, "diffusion_factor": float, "qdist": int, "nrecalc": float,
               "treering_func": None, "treering_center": galsim.PositionD, "transpose": bool}
        kwargs, _safe = GetAllParams(config, base, opt=opt)
        kwargs["rng"] = galsim.config.GetRNG(config, base, logger, "SiliconSensor")
        kwargs["device"] = _device(base)
        return B200SiliconSensor(**kwargs)


galsim.config.sensor.RegisterSensorType("Silicon", _SiliconSensorBuilder())


def TreeRingCenter(config, base, value_type):
    tree_rings = GetInputObj("tree_rings", config, base, "TreeRingCenter")
    kwargs, safe = GetAllParams(config, base, req={"det_name": str})
    return tree_rings.get_center(kwargs["det_name"]), safe


def TreeRingFunc(config, base, value_type):
    tree_rings = GetInputObj("tree_rings", config, base, "TreeRingCenter")
    kwargs, safe = GetAllParams(config, base, req={"det_name": str})
    return tree_rings.get_func(kwargs["det_name"]), safe


def _tree_rings_with_data_dir(file_name, only_dets=None, logger=None, defer_load=True):
    from imsim.meta_data import data_dir

    return TreeRings(file_name, only_dets=only_dets, logger=logger, defer_load=defer_load, data_dir=data_dir)


_tree_rings_with_data_dir._req_params = TreeRings._req_params
_tree_rings_with_data_dir._opt_params = TreeRings._opt_params
RegisterInputType("tree_rings", InputLoader(_tree_rings_with_data_dir, takes_logger=True))
RegisterValueType("TreeRingCenter", TreeRingCenter, [galsim.PositionD], input_type="tree_rings")
RegisterValueType("TreeRingFunc", TreeRingFunc, [object], input_type="tree_rings")
