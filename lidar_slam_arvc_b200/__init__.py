"""B200-native ICP scan matching behind the KeyFrame / KeyFrameManager entry points of JudithV/LIDAR_SLAM_ARVC.

Only the hot path named by BASELINE.json's north_star lives here: hand-written sm_100a CUDA kernels
(``csrc/``) behind a C-ABI (``include/arvc_icp.h``) and the Python host mirror of the reference's
registration API (``keyframemanager``, ``config``).  There is no CPU fallback: importing the engine
without the built CUDA library raises.
"""
__version__ = "0.1.0"
