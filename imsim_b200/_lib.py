"""Loader and thin ctypes binding of ``libimsim_b200.so`` (include/imsim_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc, and if
no B200 is present every compute entry point raises ``B2Error`` (the library
itself refuses to create a context).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = [os.path.join(_HERE, "csrc", f) for f in ("abi.cu", "optics.cu", "sensor.cu", "pool.cu", "stamps.cu", "stage1.cu", "readout.cu",
                                                      "hostpipe.cu")]
_HDR = [os.path.join(_HERE, "csrc", f) for f in ("b2_common.cuh", "optics_device.cuh", "sensor_device.cuh")] + \
    [os.path.join(os.path.dirname(_HERE), "include", "imsim_b200.h")]
SO_PATH = os.path.join(_HERE, "_build", "libimsim_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--threads", "0",
]


class B2Error(RuntimeError):
    pass


def needs_build() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in _SRC + _HDR)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return SO_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise B2Error("nvcc not found and %s is missing or stale: cannot build the CUDA path" % SO_PATH)
    os.makedirs(os.path.dirname(SO_PATH), exist_ok=True)
    tmp = SO_PATH + ".tmp.%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + _SRC
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise B2Error("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, SO_PATH)
    if verbose:
        print(res.stderr)
    return SO_PATH


dp = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)
vp = C.c_void_p

_SIGNATURES = {
    # name: (restype, argtypes)
    "b2_last_error": (C.c_char_p, []),
    "b2_abi_version": (C.c_int, []),
    "b2_launch_count": (C.c_uint64, []),
    "b2_telescope_program": (C.c_int, [C.c_void_p]),
    "b2_sizeof": (C.c_int64, [C.c_int32]),
    "b2_ctx_create": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
    "b2_ctx_destroy": (C.c_int, [vp]),
    "b2_ctx_set_stream": (C.c_int, [vp, vp]),
    "b2_ctx_synchronize": (C.c_int, [vp]),
    "b2_fma_peak": (C.c_int, [vp, C.c_int32, dp]),
    "b2_copy_through_ring": (C.c_int, [vp, vp, vp, C.c_int64, C.c_int32]),
    "b2_segments_constant": (C.c_int, [C.c_int64, vp, vp, vp, vp]),
    "b2_fill_segments": (C.c_int, [vp, C.c_int64, vp, vp, vp]),
    "b2_test_round_f32": (C.c_int, [vp, C.c_int64, dp, dp]),
    "b2_timing_report": (C.c_int, [C.c_char_p, C.c_int64]),
    "b2_ctx_record_kernel_events": (C.c_int, [vp, C.c_int32]),
    "b2_ctx_kernel_ms": (C.c_int, [vp, dp, C.POINTER(C.c_int64)]),
    "b2_telescope_upload": (C.c_int, [vp, C.POINTER(_abi.B2Telescope)]),
    "b2_telescope_set_extra": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int64]),
    "b2_wcs_upload": (C.c_int, [vp, C.POINTER(_abi.B2TanSip), C.POINTER(_abi.B2TanSip)]),
    "b2_detector_upload": (C.c_int, [vp, C.POINTER(_abi.B2Detector)]),
    "b2_diffraction_config": (C.c_int, [vp, C.POINTER(_abi.B2Diffraction)]),
    "b2_xy_to_v": (C.c_int, [vp, C.c_int64, vp, vp, vp, vp, vp, C.c_int]),
    "b2_v_to_xy": (C.c_int, [vp, C.c_int64, vp, vp, vp, vp, vp, C.c_int]),
    "b2_trace_rays": (C.c_int, [vp, C.c_int64] + [vp] * 10 + [C.c_int]),
    "b2_rubin_optics": (C.c_int, [vp, C.c_int64] + [vp] * 11 + [C.POINTER(_abi.B2OpticsOptions), C.c_int,
                                                              C.POINTER(_abi.B2OpticsStats)]),
    "b2_rubin_diffraction": (C.c_int, [vp, C.c_int64] + [vp] * 7 + [C.POINTER(_abi.B2OpticsOptions), C.c_int]),
    "b2_sample_time_pupil": (C.c_int, [vp, C.c_int64, vp, vp, vp, C.c_double, C.c_double, C.c_double, C.c_double,
                                       C.c_uint64, C.c_uint64, C.c_int]),
    "b2_flat_photons": (C.c_int, [vp, C.c_int64, vp, vp, vp, vp, C.c_double, C.c_double, C.c_double, C.c_double, vp, vp,
                                  C.c_int32, C.c_uint64, C.c_uint64]),
    "b2_object_photons": (C.c_int, [vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int32, vp, vp, C.c_int32,
                                    C.c_uint64, C.c_uint64]),
    "b2_psf_upload": (C.c_int, [vp, C.POINTER(_abi.B2Psf), C.POINTER(vp), vp]),
    "b2_radial_luts_upload": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_double]),
    "b2_stage1_photons": (C.c_int, [vp, C.c_int64, vp, vp, vp, vp, vp, vp, C.c_int32, vp, vp, C.c_int32, C.c_int32, vp,
                                    C.c_uint64, C.c_uint64]),
    "b2_atomic_peak": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_double)]),
    "b2_scatter_add": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int64, vp, vp, vp]),
    "b2_add_sky": (C.c_int, [vp, vp, C.c_int32, C.c_int64, C.c_double, vp, vp, C.c_uint64]),
    "b2_bleed_trails": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int]),
    "b2_readout": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.POINTER(_abi.B2Amp), C.c_int32, vp, vp, vp, C.c_int32,
                             C.c_double, C.c_int32, C.c_double, C.c_uint64, vp, vp]),
    "b2_sensor_create": (C.c_int, [vp, C.POINTER(_abi.B2SensorConfig), vp, vp, vp, vp, vp, vp, C.POINTER(vp)]),
    "b2_sensor_destroy": (C.c_int, [vp]),
    "b2_sensor_set_treerings": (C.c_int, [vp, C.c_double, C.c_double, vp, vp, vp, C.c_int32]),
    "b2_sensor_bind_image": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int]),
    "b2_sensor_read_image": (C.c_int, [vp, vp, C.c_int]),
    "b2_sensor_accumulate": (C.c_int, [vp, C.c_int64] + [vp] * 7 + [C.c_uint64, C.c_uint64, C.c_int32, C.c_int32,
                                                                   C.c_int32, C.c_int32, C.c_int,
                                                                   C.POINTER(_abi.B2AccumStats)]),
    "b2_sensor_pixel_areas": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int]),
    "b2_plain_accumulate": (C.c_int, [vp, C.c_int64, vp, vp, vp, C.c_int, dp]),
    "b2_sensor_get_pixel": (C.c_int, [vp, C.c_int32, C.c_int32, vp, vp]),
    "b2_flat_step": (C.c_int, [vp, vp, vp, C.c_int64, C.c_int32, vp, vp, C.c_int32, C.c_uint64, C.c_uint64, C.c_uint64,
                               C.c_int32, C.c_int32, C.POINTER(_abi.B2AccumStats)]),
    "b2_sensor_accumulate_stamps": (C.c_int, [vp, C.c_int32, vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp, C.c_uint64,
                                              C.c_uint64, C.c_int32, C.c_int32, vp, C.c_int32, C.c_int32, C.c_int32,
                                              C.c_int32, C.c_int32, C.POINTER(_abi.B2AccumStats), vp]),
    "b2_photons_upload": (C.c_int, [vp, C.c_int32, C.c_int64, vp, vp, vp]),
    "b2_host_memcpy": (C.c_int, [vp, vp, C.c_int64]),
    "b2_xytov_compile": (C.c_int, [vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, dp]),
    "b2_pool_step": (C.c_int, [vp, vp, C.c_int64, vp, vp, vp, vp, vp, vp, C.POINTER(_abi.B2OpticsOptions), C.c_double,
                               C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                               C.c_int32, C.c_int32, C.c_int32, C.POINTER(_abi.B2OpticsStats),
                               C.POINTER(_abi.B2AccumStats)]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load (building if needed) the CUDA library and check the ABI."""
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.b2_abi_version() != _abi.B2_ABI_VERSION:
        raise B2Error("ABI version mismatch: library %d, binding %d" % (lib.b2_abi_version(), _abi.B2_ABI_VERSION))
    for which, cls in enumerate(_abi.SIZEOF_ORDER):
        if lib.b2_sizeof(which) != C.sizeof(cls):
            raise B2Error("struct %s: library sizeof %d != binding %d" % (cls.__name__, lib.b2_sizeof(which),
                                                                          C.sizeof(cls)))
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise B2Error(load().b2_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(load().b2_launch_count())


def timing_report() -> dict:
    """Per-kernel device times {name: [launches, total_ms]} since the last call (needs B2_TIMING=1)."""
    import json

    buf = C.create_string_buffer(1 << 16)
    check(load().b2_timing_report(buf, len(buf)))
    return json.loads(buf.value.decode())


# ---- pointer helpers -----------------------------------------------------
def _is_torch(a) -> bool:
    return type(a).__module__.startswith("torch")


def where_of(a) -> int:
    """B2_DEVICE for CUDA torch tensors, B2_HOST for numpy arrays."""
    if _is_torch(a):
        if not a.is_cuda:
            raise B2Error("torch tensors passed to the B200 path must live on the GPU")
        return _abi.B2_DEVICE
    return _abi.B2_HOST


def ptr(a, dtype=np.float64):
    """Raw pointer of a contiguous numpy array / CUDA tensor (None -> NULL)."""
    if a is None:
        return None
    if _is_torch(a):
        import torch

        want = {np.float64: torch.float64, np.uint8: torch.uint8, np.float32: torch.float32}[dtype]
        if a.dtype != want or not a.is_contiguous():
            raise B2Error("device arrays must be contiguous %s" % want)
        return C.c_void_p(a.data_ptr())
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags.c_contiguous:
        raise B2Error("host arrays must be C-contiguous numpy %s" % np.dtype(dtype).name)
    return C.c_void_p(a.ctypes.data)


def h2d_async(arr, device):
    """numpy array -> CUDA tensor without synchronising the stream: staged through torch's caching pinned-host
    allocator (which holds the staging block until the copy has run) and copied with ``non_blocking=True``.
    ``torch.as_tensor(arr, device=...)`` of pageable memory ends in a ``cudaStreamSynchronize``, which stops the
    host from queueing work ahead of the GPU (visit.DetectorRunner relies on running ahead)."""
    import torch

    return torch.from_numpy(np.ascontiguousarray(arr)).pin_memory().to(device, non_blocking=True)
