class NxDPPModel:  # placeholder replaced below in this commit series
    pass
