from . import REGISTRY


class SensorBuilder:
    def buildSensor(self, config, base, logger):
        raise NotImplementedError


def RegisterSensorType(name, builder):
    REGISTRY["sensor"][name] = builder
