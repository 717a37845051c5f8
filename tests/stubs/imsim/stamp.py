"""Stand-in for imsim.stamp -- TEST INFRASTRUCTURE (the plugin re-registers imSim's own LSST_Photons builder)."""


class LSST_PhotonsBuilder:
    pass
