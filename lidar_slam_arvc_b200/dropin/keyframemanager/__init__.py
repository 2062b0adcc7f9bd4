"""Drop-in replacement of the reference's `keyframemanager` package (put lidar_slam_arvc_b200/dropin ahead of the
reference tree on sys.path).  Same import paths: `from keyframemanager.keyframemanager import KeyFrameManager`."""
import os
import sys

_REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
if _REPO not in sys.path:
    sys.path.append(_REPO)
