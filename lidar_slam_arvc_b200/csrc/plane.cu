// Ground-plane model and split for the reference's 'icp2planes' method:
//   KeyFrame.calculate_plane (keyframe.py:417-436): plane through the points below a height, Open3D segment_plane
//     (RANSAC, 3-point hypotheses, 1000 iterations, unseeded) followed by its least-squares refit of the plane to
//     the inliers of the winning hypothesis.  Here: the same hypothesis test and the same refit, but the samples come
//     from a counter-based hash of (seed, iteration), so the result is reproducible and the CPU oracle repeats it bit
//     for bit.  One block per hypothesis counts its inliers; a second kernel keeps the best one; a third refits.
//     Stated differences: hypotheses with equal inlier counts are ranked by iteration (lowest wins) where Open3D takes
//     the lower inlier RMSE, and all `iterations` hypotheses are evaluated (Open3D may stop early at probability 1).
//   KeyFrame.segment_plane (keyframe.py:438-461): |a x + b y + c z + d| / sqrt(a^2+b^2+c^2) < threshold splits the
//     cloud, order preserved - a stable two-way compaction into the raw buffers of two new scans.
// Both work on the preprocessed cloud in the reference's point order, which is first restored from the Morton-sorted
// records.  Plain streaming kernels; every floating-point operation is rounded separately (no FMA), as numpy does.
#include "engine.cuh"

namespace arvc {

__global__ void __launch_bounds__(256) k_plane_unpermute(const ScanDev* __restrict__ sp, double* __restrict__ orig) {
    const ScanDev& s = *sp;
    const int n = s.counts[CNT_NPTS];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double x, y, z;
        int idx;
        if (s.wide) load_rec(reinterpret_cast<const RecD*>(s.recs) + i, x, y, z, idx);
        else load_rec(reinterpret_cast<const RecF*>(s.recs) + i, x, y, z, idx);
        orig[3 * (size_t)idx] = x; orig[3 * (size_t)idx + 1] = y; orig[3 * (size_t)idx + 2] = z;
    }
}

__device__ __forceinline__ double plane_value(const double* pl, double x, double y, double z) {
    return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(pl[0], x), __dmul_rn(pl[1], y)), __dmul_rn(pl[2], z)), pl[3]);
}

// ---- RANSAC -------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long plane_hash(unsigned long long seed, unsigned long long a, unsigned long long b) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (a + 1) + 0xBF58476D1CE4E5B9ull * (b + 1);     // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

constexpr int kSampleTries = 64;      // draws per hypothesis before it is declared void

// Hypothesis h: three distinct points with z < max_z (rejection sampling on the hashed index), unit normal, offset.
__device__ bool plane_hypothesis(const double* __restrict__ pts, int n, double max_z, unsigned long long seed, int h, double* pl) {
    int pick[3];
    int got = 0;
    for (int t = 0; t < kSampleTries && got < 3; ++t) {
        const int j = (int)(plane_hash(seed, (unsigned long long)h, (unsigned long long)t) % (unsigned long long)n);
        if (!(pts[3 * (size_t)j + 2] < max_z)) continue;
        bool dup = false;
        for (int k = 0; k < got; ++k) dup |= pick[k] == j;
        if (!dup) pick[got++] = j;
    }
    if (got < 3) return false;
    const double* p0 = pts + 3 * (size_t)pick[0];
    const double* p1 = pts + 3 * (size_t)pick[1];
    const double* p2 = pts + 3 * (size_t)pick[2];
    const double ux = __dsub_rn(p1[0], p0[0]), uy = __dsub_rn(p1[1], p0[1]), uz = __dsub_rn(p1[2], p0[2]);
    const double vx = __dsub_rn(p2[0], p0[0]), vy = __dsub_rn(p2[1], p0[1]), vz = __dsub_rn(p2[2], p0[2]);
    const double nx = __dsub_rn(__dmul_rn(uy, vz), __dmul_rn(uz, vy));
    const double ny = __dsub_rn(__dmul_rn(uz, vx), __dmul_rn(ux, vz));
    const double nz = __dsub_rn(__dmul_rn(ux, vy), __dmul_rn(uy, vx));
    const double len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny)), __dmul_rn(nz, nz)));
    if (!(len > 1e-12)) return false;
    pl[0] = __ddiv_rn(nx, len); pl[1] = __ddiv_rn(ny, len); pl[2] = __ddiv_rn(nz, len);
    pl[3] = -__dadd_rn(__dadd_rn(__dmul_rn(pl[0], p0[0]), __dmul_rn(pl[1], p0[1])), __dmul_rn(pl[2], p0[2]));
    return true;
}

__global__ void __launch_bounds__(256) k_plane_ransac(const ScanDev* __restrict__ sp, const double* __restrict__ pts, double max_z,
                                                      double thr, unsigned long long seed, int* __restrict__ score) {
    __shared__ double s_pl[4];
    __shared__ int s_ok, s_cnt[8];
    const int n = sp->counts[CNT_NPTS], h = blockIdx.x;
    if (threadIdx.x == 0) s_ok = (n >= 3 && plane_hypothesis(pts, n, max_z, seed, h, s_pl)) ? 1 : 0;
    __syncthreads();
    if (!s_ok) { if (threadIdx.x == 0) score[h] = 0; return; }
    const double pl[4] = {s_pl[0], s_pl[1], s_pl[2], s_pl[3]};
    int c = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
        if (z < max_z && fabs(plane_value(pl, x, y, z)) < thr) ++c;
    }
    c = warp_sum(c);
    if (lane_id() == 0) s_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_cnt[w];
        score[h] = t;
    }
}

// result[0..3] = plane of the best hypothesis (most inliers, lowest iteration on ties), result[4] = inliers (0: none)
__global__ void __launch_bounds__(1024) k_plane_pick(const ScanDev* __restrict__ sp, const double* __restrict__ pts, double max_z,
                                                     unsigned long long seed, const int* __restrict__ score, int iters, double* __restrict__ result) {
    __shared__ long long s_best[32];
    long long best = -1;                                  // (count << 32) | (0x7fffffff - h): max <=> most inliers, then lowest h
    for (int h = threadIdx.x; h < iters; h += blockDim.x)
        if (score[h] > 0) best = max(best, ((long long)score[h] << 32) | (long long)(0x7fffffff - h));
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(kFull, best, o));
    if (lane_id() == 0) s_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 0; w < 32; ++w) best = max(best, s_best[w]);
        double pl[4] = {0, 0, 0, 0};
        int cnt = 0;
        if (best >= 0) {
            cnt = (int)(best >> 32);
            plane_hypothesis(pts, sp->counts[CNT_NPTS], max_z, seed, 0x7fffffff - (int)(best & 0xffffffffll), pl);
        }
        result[0] = pl[0]; result[1] = pl[1]; result[2] = pl[2]; result[3] = pl[3]; result[4] = (double)cnt;
    }
}

// Open3D's segment_plane does not return the winning 3-point hypothesis: it collects the final inliers of that
// hypothesis (|plane . (x, y, z, 1)| < threshold) and refits the plane to ALL of them (GetPlaneFromPoints: centroid,
// centred second moments, normal = the largest of the three 2x2-determinant cross products, normalised; d = -n.c).
// Fixed summation order so that the oracle repeats it bit for bit: thread t of 1024 sums the inliers i = t, t + 1024, ...
// in ascending order, then xor-butterfly inside every warp (16, 8, 4, 2, 1), then the same butterfly over the 32 warp sums.
__device__ __forceinline__ double block1024_sum(double v, double* s_w /*[32]*/) {
    v = warp_sum(v);
    __syncthreads();
    if (lane_id() == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = s_w[lane_id()];
    t = warp_sum(t);
    return t;      // every thread returns the same value
}

__global__ void __launch_bounds__(1024) k_plane_refit(const ScanDev* __restrict__ sp, const double* __restrict__ pts, double max_z, double thr,
                                                      double* __restrict__ result) {
    __shared__ double s_w[32];
    const int n = sp->counts[CNT_NPTS];
    const double pl[4] = {result[0], result[1], result[2], result[3]};
    if (!(result[4] >= 3.0)) return;
    double sx = 0.0, sy = 0.0, sz = 0.0, cnt = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) {
        const double x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
        if (z < max_z && fabs(plane_value(pl, x, y, z)) < thr) { sx = __dadd_rn(sx, x); sy = __dadd_rn(sy, y); sz = __dadd_rn(sz, z); cnt += 1.0; }
    }
    sx = block1024_sum(sx, s_w); sy = block1024_sum(sy, s_w); sz = block1024_sum(sz, s_w); cnt = block1024_sum(cnt, s_w);
    const double cx = __ddiv_rn(sx, cnt), cy = __ddiv_rn(sy, cnt), cz = __ddiv_rn(sz, cnt);
    double xx = 0.0, xy = 0.0, xz = 0.0, yy = 0.0, yz = 0.0, zz = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) {
        const double x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
        if (z < max_z && fabs(plane_value(pl, x, y, z)) < thr) {
            const double r0 = __dsub_rn(x, cx), r1 = __dsub_rn(y, cy), r2 = __dsub_rn(z, cz);
            xx = __dadd_rn(xx, __dmul_rn(r0, r0)); xy = __dadd_rn(xy, __dmul_rn(r0, r1)); xz = __dadd_rn(xz, __dmul_rn(r0, r2));
            yy = __dadd_rn(yy, __dmul_rn(r1, r1)); yz = __dadd_rn(yz, __dmul_rn(r1, r2)); zz = __dadd_rn(zz, __dmul_rn(r2, r2));
        }
    }
    xx = block1024_sum(xx, s_w); xy = block1024_sum(xy, s_w); xz = block1024_sum(xz, s_w);
    yy = block1024_sum(yy, s_w); yz = block1024_sum(yz, s_w); zz = block1024_sum(zz, s_w);
    if (threadIdx.x != 0) return;
    const double det_x = __dsub_rn(__dmul_rn(yy, zz), __dmul_rn(yz, yz));
    const double det_y = __dsub_rn(__dmul_rn(xx, zz), __dmul_rn(xz, xz));
    const double det_z = __dsub_rn(__dmul_rn(xx, yy), __dmul_rn(xy, xy));
    double a, b, c;
    if (det_x > det_y && det_x > det_z) {
        a = det_x; b = __dsub_rn(__dmul_rn(xz, yz), __dmul_rn(xy, zz)); c = __dsub_rn(__dmul_rn(xy, yz), __dmul_rn(xz, yy));
    } else if (det_y > det_z) {
        a = __dsub_rn(__dmul_rn(xz, yz), __dmul_rn(xy, zz)); b = det_y; c = __dsub_rn(__dmul_rn(xy, xz), __dmul_rn(yz, xx));
    } else {
        a = __dsub_rn(__dmul_rn(xy, yz), __dmul_rn(xz, yy)); b = __dsub_rn(__dmul_rn(xy, xz), __dmul_rn(yz, xx)); c = det_z;
    }
    const double norm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c)));
    if (norm == 0.0) { result[0] = result[1] = result[2] = result[3] = 0.0; result[4] = 0.0; return; }      // the inliers do not span a plane
    a = __ddiv_rn(a, norm); b = __ddiv_rn(b, norm); c = __ddiv_rn(c, norm);
    result[0] = a; result[1] = b; result[2] = c;
    result[3] = -__dadd_rn(__dadd_rn(__dmul_rn(a, cx), __dmul_rn(b, cy)), __dmul_rn(c, cz));
}

void run_plane_fit(Launcher& L, const ScanDev* d_scan, int cap, double* d_orig, int* d_score, double* d_result, double max_z, double thr,
                   int iters, unsigned long long seed) {
    L.launch("plane_unpermute", k_plane_unpermute, dim3(max(1, min((cap + 255) / 256, 592))), dim3(256), d_scan, d_orig);
    L.launch("plane_ransac", k_plane_ransac, dim3(iters), dim3(256), d_scan, (const double*)d_orig, max_z, thr, seed, d_score);
    L.launch("plane_pick", k_plane_pick, dim3(1), dim3(1024), d_scan, (const double*)d_orig, max_z, seed, (const int*)d_score, iters, d_result);
    L.launch("plane_refit", k_plane_refit, dim3(1), dim3(1024), d_scan, (const double*)d_orig, max_z, thr, d_result);
}

// ---- split --------------------------------------------------------------------------------------------
struct PlaneSplit { double pl[4]; double norm; double thr; };
constexpr int kSplitBlock = 1024;

__device__ __forceinline__ bool near_plane(const PlaneSplit& ps, double x, double y, double z) {
    return __ddiv_rn(fabs(plane_value(ps.pl, x, y, z)), ps.norm) < ps.thr;      // keyframe.py:452-453
}

__global__ void __launch_bounds__(kSplitBlock) k_plane_count(const ScanDev* __restrict__ sp, const double* __restrict__ orig, PlaneSplit ps,
                                                            int* __restrict__ blk) {
    const int n = sp->counts[CNT_NPTS];
    if (blockIdx.x * kSplitBlock >= n && blockIdx.x > 0) return;
    const int i = blockIdx.x * kSplitBlock + threadIdx.x;
    const bool pred = i < n && near_plane(ps, orig[3 * (size_t)i], orig[3 * (size_t)i + 1], orig[3 * (size_t)i + 2]);
    const int c = __syncthreads_count(pred);
    if (threadIdx.x == 0) blk[blockIdx.x] = c;
}

__global__ void __launch_bounds__(kSplitBlock) k_plane_scatter(const ScanDev* __restrict__ sp, const double* __restrict__ orig, PlaneSplit ps,
                                                              const int* __restrict__ blk, double* __restrict__ near, double* __restrict__ far,
                                                              int* __restrict__ counts2) {
    __shared__ int s_warp[32], s_base;
    const int n = sp->counts[CNT_NPTS];
    if (blockIdx.x * kSplitBlock >= n && blockIdx.x > 0) return;
    const int i = blockIdx.x * kSplitBlock + threadIdx.x, lane = lane_id(), w = threadIdx.x >> 5;
    double x = 0, y = 0, z = 0;
    bool pred = false;
    if (i < n) {
        x = orig[3 * (size_t)i]; y = orig[3 * (size_t)i + 1]; z = orig[3 * (size_t)i + 2];
        pred = near_plane(ps, x, y, z);
    }
    if (w == 0) {                                         // near points in the blocks before this one
        int s = 0;
        for (int b = lane; b < (int)blockIdx.x; b += 32) s += blk[b];
        s = warp_sum(s);
        if (lane == 0) s_base = s;
    }
    const unsigned m = __ballot_sync(kFull, pred);
    if (lane == 0) s_warp[w] = __popc(m);
    __syncthreads();
    int before = s_base;
    for (int k = 0; k < w; ++k) before += s_warp[k];
    before += __popc(m & ((1u << lane) - 1u));
    if (i < n) {
        double* o = pred ? near + 3 * (size_t)before : far + 3 * (size_t)(i - before);
        o[0] = x; o[1] = y; o[2] = z;
    }
    if (i == n - 1 || (n == 0 && i == 0)) {
        const int n_near = n == 0 ? 0 : before + (pred ? 1 : 0);
        counts2[0] = n_near; counts2[1] = n - n_near;
    }
}

void run_plane_split(Launcher& L, const ScanDev* d_scan, int cap, double* d_orig, int* d_blk, double* d_near, double* d_far, int* d_counts2,
                     const double* plane, double norm, double thr) {
    PlaneSplit ps{{plane[0], plane[1], plane[2], plane[3]}, norm, thr};
    const int nb = max(1, (cap + kSplitBlock - 1) / kSplitBlock);
    L.launch("plane_unpermute", k_plane_unpermute, dim3(max(1, min((cap + 255) / 256, 592))), dim3(256), d_scan, d_orig);
    L.launch("plane_count", k_plane_count, dim3(nb), dim3(kSplitBlock), d_scan, (const double*)d_orig, ps, d_blk);
    L.launch("plane_scatter", k_plane_scatter, dim3(nb), dim3(kSplitBlock), d_scan, (const double*)d_orig, ps, (const int*)d_blk, d_near, d_far, d_counts2);
}

}  // namespace arvc
