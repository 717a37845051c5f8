"""Stand-in for imsim.flat -- TEST INFRASTRUCTURE (the plugin re-registers imSim's own LSST_Flat builder)."""


class LSST_FlatBuilder:
    pass
