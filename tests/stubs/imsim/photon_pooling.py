"""Stand-in for imsim.photon_pooling.LSST_PhotonPoolingImageBuilder -- TEST INFRASTRUCTURE.

The plugin's builder inherits everything except ``buildImage`` from imSim's class.  Here the inherited part is the
repo's mirror of the batching algebra (pinned to the reference's source, tests/golden/pooling.npz) plus minimal
instance methods: a catalogue of (x, y, flux) in ``base['_catalog']``, stamps whose photons come from
``base['_photon_source'](obj)`` -- GalSim's shooters stand behind that in a real run."""
import galsim
import numpy as np

from imsim_b200.photon_pooling import LSST_PhotonPoolingImageBuilder as _Mirror
from imsim_b200.photon_pooling import ObjectInfo, ProcessingMode


class _Stamp:
    def __init__(self, photons):
        self.photons = photons
        self.bounds = None


class LSST_PhotonPoolingImageBuilder(_Mirror):
    def setup(self, config, base, image_num, obj_num, ignore, logger):
        self.nbatch = int(config.get("nbatch", 10))
        self.nbatch_fft = int(config.get("nbatch_fft", 1))
        self.nsubbatch = int(config.get("nsubbatch", 50))
        self.det_name = config.get("det_name", base.get("det_name", "R22_S11"))
        self.checkpoint = None
        self.nobjects = len(base["_catalog"])
        self.xsize, self.ysize = int(config["xsize"]), int(config["ysize"])
        return self.xsize, self.ysize

    def _set_config_image_pos(self, config, base):
        pass

    def _create_full_image(self, config, base):
        img = galsim.ImageF(self.xsize, self.ysize, wcs=base.get("wcs"), dtype=np.dtype(config.get("dtype", "float32")))
        base["current_image"] = img
        return img

    def load_objects(self, obj_nums, config, base, logger):
        for k in obj_nums:
            yield ObjectInfo(k, int(base["_catalog"][k]["flux"]), ProcessingMode.PHOT)

    @staticmethod
    def build_stamps(base, logger, objects):
        if not objects:
            return [], []
        src = base["_photon_source"]
        images = [_Stamp(src(obj)) for obj in objects if obj.phot_flux > 0]
        return images, tuple(0.0 for _ in objects)
