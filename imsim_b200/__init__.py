"""imsim_b200 -- B200-native photon-shooting hot path with imSim's plugin surface.

Host code is Python (this package); all per-photon arithmetic runs in the
hand-written sm_100a CUDA library ``_build/libimsim_b200.so`` behind the C ABI
of ``include/imsim_b200.h``.  There is no CPU fallback.
"""
from ._lib import B2Error, build, load, launch_count  # noqa: F401
from .photon_array import PhotonArray  # noqa: F401
from .telescope import Telescope, rubin_like  # noqa: F401
from .wcs import TanSipWCS, fit_tan_sip, field_wcs  # noqa: F401
from .detector import DetectorGeometry, lsstcam_like, lsstcam_science_detectors  # noqa: F401
from .diffraction import RUBIN_SPIDER_GEOMETRY, diffraction_config  # noqa: F401
from .context import OpticsContext  # noqa: F401
from .photon_ops import (  # noqa: F401
    BandpassRatio, RubinDiffraction, RubinDiffractionOptics, RubinOptics, XyToV, photon_velocity,
    ray_vector_to_photon_array, make_rubin_diffraction_optics,
)
from .treerings import TreeRings, TreeRingRadialFunction, RadialTable  # noqa: F401
from .sensor import SiliconSensor, Sensor, Image  # noqa: F401
from .photon_pooling import (  # noqa: F401
    LSST_PhotonPoolingImageBuilder, ObjectInfo, ProcessingMode, PhotonPool, DevicePhotons, PinnedPhotons,
)

__version__ = "0.1.0"
