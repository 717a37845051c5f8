"""Sky background and its shot noise on the device (mirror of ``LSST_ImageBuilderBase.addNoise``,
imsim/lsst_image.py:128-199).

The reference adds ``sky`` (a level in photons per pixel from the sky model, optionally multiplied by the sky
gradient, vignetting and fringing maps, lsst_image.py:158-196) to the finished image and hands it to GalSim's
config noise builder; with photon-shot objects only the sky's own Poisson noise is left to add, and when the
sky is put through the silicon model its level is weighted by the pixel areas (tree rings always,
brighter-fatter if ``use_flux_sky_areas``; config/imsim-config.yaml:222-228).  Here one kernel does
``image += Poisson(sky_level * area * modulation)`` with exact Poisson deviates from Philox.
"""
from __future__ import annotations

import ctypes as C

from . import _abi, _lib


def pixel_areas_device(sensor, use_flux: bool = True, orig_center=(0, 0)):
    """``sensor.calculate_pixel_areas`` of the sensor's bound image as a CUDA float64 tensor (the image and its
    charge stay on the device).  Like GalSim's, this rebuilds the pixel boundaries from the image in one step,
    so it belongs after the last ``accumulate`` on that image."""
    import torch

    ny, nx, _ = sensor._bound_shape
    areas = torch.empty((ny, nx), dtype=torch.float64, device="cuda:%d" % sensor.ctx.device)
    _lib.check(_lib.load().b2_sensor_pixel_areas(sensor._h, int(orig_center[0]), int(orig_center[1]), int(bool(use_flux)),
                                                 C.c_void_p(areas.data_ptr()), _abi.B2_DEVICE))
    sensor._last_image = None
    return areas


def add_sky(ctx, image, sky_level: float, seed: int, areas=None, modulation=None):
    """``image`` (CUDA float32 / float64 tensor) += Poisson(sky_level * areas * modulation), in place."""
    import torch

    assert image.is_cuda and image.is_contiguous() and image.dtype in (torch.float32, torch.float64)
    if areas is not None:
        assert areas.dtype == torch.float64 and areas.is_contiguous() and areas.numel() == image.numel()
    if modulation is not None:
        modulation = modulation.to(dtype=torch.float32, device=image.device).contiguous()
        assert modulation.numel() == image.numel()
    _lib.check(_lib.load().b2_add_sky(ctx.handle, C.c_void_p(image.data_ptr()), image.element_size(), image.numel(),
                                      float(sky_level), C.c_void_p(areas.data_ptr()) if areas is not None else None,
                                      C.c_void_p(modulation.data_ptr()) if modulation is not None else None,
                                      int(seed) & 0xFFFFFFFFFFFFFFFF))
    return image
