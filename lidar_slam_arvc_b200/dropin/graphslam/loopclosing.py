"""LoopClosing with the registrations of one invocation issued as ONE device batch.

Same class, constructor, methods and acceptance gates as the reference's graphslam/loopclosing.py:9-270; what
changes is the call pattern.  The reference registers the sampled candidates one after another
(loopclosing.py:54-56 and :80-99: up to 2 x number_of_triplets sequential ICPs, each re-running load + pre_process
of both keyframes).  The pose estimate the initial guesses come from is not modified inside that loop (add_edge only
appends a factor), so all the pairs are independent: here they are collected first, their keyframes are loaded and
pre-processed once, and `KeyFrameManager.compute_transformations` runs them as one batch on the GPU.  The random
draws (`np.random.choice`, same arguments, same order) and the order of `add_edge` calls are the reference's, so a
seeded run adds the same edges with the same transforms.

`graphslam` is duck-typed exactly as in the reference: `.current_estimate.atPose3(i).matrix()`,
`.current_estimate.exists(i)`, `.T0_gps`, `.add_edge(T, i, j, 'SM')`.
"""
import numpy as np

from lidar_slam_arvc_b200.homogeneousmatrix import result_type


def _H(array):
    return result_type()(array)


class LoopClosing():
    def __init__(self, graphslam, distance_backwards=7, radius_threshold=5.0):
        self.graphslam = graphslam
        self.distance_backwards = distance_backwards      # path length to skip before candidates are accepted
        self.radius_threshold = radius_threshold
        self.positions = None

    # ------------------------------------------------------------------ the two public procedures
    def loop_closing_simple(self, current_index, number_of_candidates_DA, keyframe_manager):
        """Registers the current keyframe against randomly chosen past keyframes within radius_threshold and adds every
        result as an edge (reference loopclosing.py:33-57)."""
        candidates = self.find_candidates()
        print(candidates)
        i = current_index
        n = np.min([len(candidates), number_of_candidates_DA])
        candidates = np.random.choice(candidates, size=n, replace=False)
        pairs = [(i, int(j)) for j in candidates]
        for (a, b), Tab in zip(pairs, self.compute_transformations_between_pairs(pairs, keyframe_manager)):
            self.add_loop_closing_observation(i=a, j=b, Tij=Tab)
        return

    def loop_closing_triangle(self, current_index, number_of_triplets_loop_closing, keyframe_manager):
        """Triplets (i, j1, j2): Tij1 * Tj1j2 * Tij2^-1 must be close to the identity for both observations to be kept
        (reference loopclosing.py:58-100)."""
        triplets = self.find_feasible_triplets(current_index=current_index)
        if len(triplets) == 0:
            return
        triplet_indexes = range(len(triplets))
        n = np.min([len(triplet_indexes), number_of_triplets_loop_closing])
        sampled = np.random.choice(triplet_indexes, size=n, replace=False)
        pairs = []
        for k in sampled:
            i, j1, j2 = (int(v) for v in triplets[k])
            pairs += [(i, j1), (i, j2)]
        observed = self.compute_transformations_between_pairs(pairs, keyframe_manager)
        added_loop_closures = []
        for m, k in enumerate(sampled):
            i, j1, j2 = (int(v) for v in triplets[k])
            print('Checking loop closing triplet: ', triplets[k])
            Tij1, Tij2 = observed[2 * m], observed[2 * m + 1]
            Tj1j2 = self.compute_consecutive_transformations(i=j1, j=j2)
            I = Tij1 * Tj1j2 * Tij2.inv()
            print('Found loop closing triplet I: ', I)
            if self.check_distances(I):
                print('Consistent triplet: adding both loop closing observations.')
                self.add_loop_closing_observation(i=i, j=j1, Tij=Tij1)
                self.add_loop_closing_observation(i=i, j=j2, Tij=Tij2)
                added_loop_closures.append([i, j1])
                added_loop_closures.append([i, j2])
        return added_loop_closures

    # ------------------------------------------------------------------ registration
    def compute_transformations_between_pairs(self, pairs, keyframe_manager):
        """Observed transforms for a list of (i, j), in the frame convention of
        compute_transformations_between_candidates (reference loopclosing.py:154-184), one device batch."""
        if len(pairs) == 0:
            return []
        T0_gps = self.graphslam.T0_gps
        guesses = [self.compute_consecutive_transformations(i=i, j=j) for i, j in pairs]
        if not hasattr(keyframe_manager, 'compute_transformations'):
            # a manager without the batched entry point (e.g. the reference's own class): pair by pair
            return [self._register_one(i, j, Tij, keyframe_manager) for (i, j), Tij in zip(pairs, guesses)]
        touched = sorted({k for p in pairs for k in p})
        for k in touched:
            keyframe_manager.load_pointcloud(k)
        keyframe_manager.pre_process_many(touched)
        observed, _ = keyframe_manager.compute_transformations(pairs, guesses)
        # ICP works LiDAR to LiDAR; the graph holds GPS-frame poses
        return [T0_gps.inv() * Tijsm * T0_gps for Tijsm in observed]

    def _register_one(self, i, j, Tij, keyframe_manager):
        T0_gps = self.graphslam.T0_gps
        keyframe_manager.load_pointcloud(i)
        keyframe_manager.pre_process(i)
        keyframe_manager.load_pointcloud(j)
        keyframe_manager.pre_process(j)
        return T0_gps.inv() * keyframe_manager.compute_transformation(i, j, Tij=Tij) * T0_gps

    def compute_transformations_between_candidates(self, i, j, keyframe_manager):
        """Single pair, kept for callers of the reference method (loopclosing.py:154)."""
        return self.compute_transformations_between_pairs([(i, j)], keyframe_manager)[0]

    def compute_consecutive_transformations(self, i, j):
        """iTj predicted by the current graph estimate, LiDAR to LiDAR (reference loopclosing.py:186-200)."""
        T0_gps = self.graphslam.T0_gps
        Ti = _H(self.graphslam.current_estimate.atPose3(i).matrix()) * T0_gps.inv()
        Tj = _H(self.graphslam.current_estimate.atPose3(j).matrix()) * T0_gps.inv()
        return Ti.inv() * Tj

    def add_loop_closing_observation(self, i, j, Tij):
        print('Adding loop_closing edge (i, j): ', i, j)
        self.graphslam.add_edge(Tij, i, j, 'SM')

    # ------------------------------------------------------------------ candidate selection (host, O(#poses))
    def find_feasible_triplets(self, current_index):
        candidates = self.find_candidates()
        if len(candidates) == 0:
            return []
        print('Found candidates within radius distance threshold:')
        print(candidates)
        candidates = np.sort(candidates)
        triplets = []
        for k in range(len(candidates)):
            j1 = candidates[k]
            j2 = self.look_for_valid_indexes(j1, candidates[k:])
            if j2 is not None:
                triplets.append([current_index, j1, j2])
        return triplets

    def check_distances(self, I):
        dp = np.linalg.norm(I.pos())
        eul = I.euler()
        da = min(np.linalg.norm(eul[0].abg), np.linalg.norm(eul[1].abg))
        print('Found triangle loop closing distances: ', dp, da)
        if dp < 0.1 and da < 0.05:
            return True
        print('Inconsistent loop closing triplet: discarded')
        return False

    def look_for_valid_indexes(self, i, rest_of_candidates):
        """First candidate u with 1 < |u - i| < 80 and 1.0 m < distance(i, u) < 2.0 m (reference loopclosing.py:131-145).
        The reference tests the candidates one by one (a Python loop calling distance(i, u) per candidate - quadratic in
        the number of candidates, seconds per invocation on a long trajectory); here the same predicate is evaluated for
        all of them at once on the positions store_positions() has just read from the same estimate."""
        rest = np.asarray(rest_of_candidates)
        if len(rest) == 0:
            return None
        if self.positions is None or len(self.positions) <= max(int(rest.max()), int(i)):
            for u in rest:                                    # no position table (called on its own): the reference's loop
                if (1 < abs(u - i) < 80) and (1.0 < self.distance(i, u) < 2.0):
                    return u
            return None
        gap = np.abs(rest - i)
        d = np.sqrt(((self.positions[rest] - self.positions[i]) ** 2).sum(axis=1))
        ok = np.nonzero((gap > 1) & (gap < 80) & (d > 1.0) & (d < 2.0))[0]
        return rest[ok[0]] if len(ok) else None

    def store_positions(self):
        # loopclosing.py:202-211 wraps every pose in a HomogeneousMatrix to read its position: the translation column of the
        # 4x4 is the same three numbers without the ~250 k temporary objects of a 5 000-keyframe run
        est = self.graphslam.current_estimate
        poses = []
        i = 0
        while est.exists(i):
            poses.append(np.asarray(est.atPose3(i).matrix(), dtype=np.float64)[0:3, 3])
            i += 1
        self.positions = np.array(poses).reshape(-1, 3)

    def find_candidates(self):
        self.store_positions()
        return self.find_candidates_within_radius(self.find_index_backwards())

    def find_index_backwards(self):
        """Newest index that lies more than distance_backwards of travelled path behind the current pose."""
        d = 0
        for i in reversed(range(len(self.positions) - 1)):
            d += np.linalg.norm(self.positions[i + 1] - self.positions[i])
            if d > self.distance_backwards:
                return i
        return None

    def find_candidates_within_radius(self, index):
        if index is None:
            return []
        d = np.linalg.norm(self.positions[0:index, :] - self.positions[-1, :], axis=1)
        return np.where(d < self.radius_threshold)[0]

    def distance(self, i, j):
        est = self.graphslam.current_estimate
        return np.linalg.norm(_H(est.atPose3(i).matrix()).pos() - _H(est.atPose3(j).matrix()).pos())
