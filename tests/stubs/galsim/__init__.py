"""Minimal stand-in for the parts of GalSim that imsim_b200.galsim_plugin touches -- TEST INFRASTRUCTURE.
It records registrations so that tests can check the plugin's wiring without GalSim installed."""


class Angle:
    def __init__(self, rad):
        self.rad = rad


class CelestialCoord:
    def __init__(self, ra=0.0, dec=0.0):
        self.ra, self.dec = ra, dec


class PositionD:
    def __init__(self, x=0.0, y=0.0):
        self.x, self.y = x, y


class BaseDeviate:
    def __init__(self, seed=0):
        self._seed = int(seed)

    def raw(self):
        return self._seed


class SiliconSensor:  # the plugin's sensor subclasses this for the isinstance test of photon_pooling.py:209
    def __init__(self, *a, **k):
        raise AssertionError("galsim.SiliconSensor.__init__ must not run for the B200 sensor")


from . import config  # noqa: E402,F401
