// Map building on the device: the reference's KeyFrameManager.build_map
// (keyframemanager.py:154-184) filters and down-samples every keyframe, moves it to the global frame with
// KeyFrame.transform (keyframe.py:399-400 -> Open3D PointCloud::Transform) and concatenates the clouds on the host,
// one keyframe at a time.  Here the keyframes of a batch are preprocessed together (preprocess.cu); two small
// kernels then place every cloud at its offset of ONE output array, in the reference's point order.
// Both are plain HBM streaming kernels: 16 B (float32 records) or 32 B (float64) read and 24 B written per point.
#include "engine.cuh"

namespace arvc {

constexpr int kOffThreads = 1024;

// offsets[s] = number of points of the scans before s; offsets[n] = total; offsets[n+1] = OR of the scans' error
// flags (so the host needs one read-back for the whole batch).  One block.
__global__ void __launch_bounds__(kOffThreads) k_map_offsets(const ScanDev* const* __restrict__ scans, int n, long long* __restrict__ offsets) {
    __shared__ long long s_sum[kOffThreads];
    __shared__ int s_err;
    if (threadIdx.x == 0) s_err = 0;
    __syncthreads();
    const int t = threadIdx.x, per = (n + kOffThreads - 1) / kOffThreads;
    const int lo = min(n, t * per), hi = min(n, lo + per);
    long long local = 0;
    int err = 0;
    for (int k = lo; k < hi; ++k) { local += scans[k]->counts[CNT_NPTS]; err |= scans[k]->counts[CNT_ERR]; }
    if (err) atomicOr(&s_err, err);
    s_sum[t] = local;
    __syncthreads();
    for (int d = 1; d < kOffThreads; d <<= 1) {
        const long long v = t >= d ? s_sum[t - d] : 0;
        __syncthreads();
        s_sum[t] += v;
        __syncthreads();
    }
    long long run = s_sum[t] - local;
    for (int k = lo; k < hi; ++k) { offsets[k] = run; run += scans[k]->counts[CNT_NPTS]; }
    if (t == kOffThreads - 1) { offsets[n] = s_sum[t]; offsets[n + 1] = s_err; }
}

// out[offsets[s] + idx] = T_s * p: the oracle's operation order (products summed left to right, no FMA, then the
// division by w that Open3D applies), so the map is bit-identical to the CPU restatement.
__global__ void __launch_bounds__(256) k_map_transform(const ScanDev* const* __restrict__ scans, const double* __restrict__ T,
                                                       const long long* __restrict__ offsets, double* __restrict__ out, long long capacity) {
    const ScanDev& sc = *scans[blockIdx.y];
    const int n = sc.counts[CNT_NPTS];
    const long long off = offsets[blockIdx.y];
    if (off + n > capacity) return;                      // reported by the host from offsets[n]
    double M[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) M[k] = T[16 * (size_t)blockIdx.y + k];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double x, y, z;
        int idx;
        if (sc.wide) load_rec(reinterpret_cast<const RecD*>(sc.recs) + i, x, y, z, idx);
        else load_rec(reinterpret_cast<const RecF*>(sc.recs) + i, x, y, z, idx);
        double r[4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
            r[a] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M[4 * a], x), __dmul_rn(M[4 * a + 1], y)), __dmul_rn(M[4 * a + 2], z)), M[4 * a + 3]);
        double* o = out + 3 * (off + idx);
        o[0] = __ddiv_rn(r[0], r[3]);
        o[1] = __ddiv_rn(r[1], r[3]);
        o[2] = __ddiv_rn(r[2], r[3]);
    }
}

void run_map_build(Launcher& L, const ScanDev* const* d_scans, const double* d_T, int n_scans, int cap_max, long long* d_offsets,
                   double* d_out, long long capacity, bool offsets_only) {
    if (n_scans == 0) return;
    if (offsets_only) {
        L.launch("map_offsets", k_map_offsets, dim3(1), dim3(kOffThreads), d_scans, n_scans, d_offsets);
        return;
    }
    const int gx = max(1, min((cap_max + 255) / 256, 64));
    L.launch("map_transform", k_map_transform, dim3(gx, n_scans), dim3(256), d_scans, d_T, (const long long*)d_offsets, d_out, capacity);
}

}  // namespace arvc
