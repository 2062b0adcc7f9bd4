"""CPU oracle for the ICP scan-matching hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product (``lidar_slam_arvc_b200``) never does.

PARITY UNPINNED: the reference (JudithV/LIDAR_SLAM_ARVC) delegates the arithmetic of this path to
Open3D (unpinned pip dependency, ``requirements.txt:2``), which is absent from ``/root/reference`` and not
installable here, and the reference has no tests or golden vectors.  ``icp_oracle.cpp`` restates Open3D's
published legacy-CPU algorithms; ``numpy_ref.py`` is an independent second implementation used to
cross-check it.  The only part pinned by the reference's own code is ``filter_radius_height``
(keyframemanager/keyframe.py:74-94) and the ``artelib`` SE(3) helpers — see ``tests/golden``.
"""
