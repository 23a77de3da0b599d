"""slam_decomposition_b200 -- B200-native engine for SLAM's template-evaluation hot path.

Drop-in Python surface of Pitt-JonesLab/slam_decomposition for that path (``CircuitTemplate``,
``CircuitTemplateV2``, cost functionals, samplers, ``TemplateOptimizer``) over hand-written
sm_100a CUDA kernels reached through the C ABI in ``include/slam_b200.h``.
"""
__version__ = "0.1.0"
