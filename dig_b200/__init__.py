"""dig_b200 -- B200-native (sm_100a) implementation of DIG's 3D-graph message-passing hot path.

Drop-in surface (mirrors dig.threedgraph of the reference):

    from dig_b200.threedgraph.method import SchNet, SphereNet, DimeNetPP, ComENet, run
    from dig_b200.threedgraph.evaluation import ThreeDEvaluator
    from dig_b200.threedgraph.utils import xyz_to_dat
"""
__version__ = "0.1.0"
