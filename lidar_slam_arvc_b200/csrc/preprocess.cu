// Preprocessing kernels: radius/height filter + stable compaction, voxel down-sampling (integer voxel
// keys, stable radix sort, segmented mean), Morton sort and multi-resolution hash-grid build.
// All kernels are batched over scans: blockIdx.y selects the scan, blockIdx.x the tile.
//
// Reference semantics restated (paths relative to the reference tree):
//   filter  : keyframemanager/keyframe.py:74-94 (the reference's own numpy)
//   voxel   : keyframe.py:111,151,159 -> Open3D PointCloud::VoxelDownSample
#include "engine.cuh"

namespace arvc {

// ---------------------------------------------------------------------------------------------------
// block-wide exclusive scan of a predicate (blockDim.x <= 1024)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_excl_scan_flag(bool pred, int& total, int* s_warp /*[33]*/) {
    const unsigned b = __ballot_sync(kFull, pred);
    const int lane = lane_id(), w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) s_warp[w] = __popc(b);
    __syncthreads();
    if (w == 0) {
        int v = lane < nw ? s_warp[lane] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += t;
        }
        s_warp[lane] = inc - v;
        if (lane == 31) s_warp[32] = inc;
    }
    __syncthreads();
    total = s_warp[32];
    const int r = s_warp[w] + __popc(b & ((1u << lane) - 1u));
    __syncthreads();
    return r;
}

// sum of blk[0..b) by one warp (b <= a few hundred)
__device__ __forceinline__ int warp_prefix_of_blocks(const int* blk, int b) {
    int s = 0;
    for (int i = lane_id(); i < b; i += 32) s += blk[i];
    return warp_sum(s);
}

__device__ __forceinline__ void load_raw(const ScanDev& s, int i, double& x, double& y, double& z) {
    if (s.raw_f64) {
        const double* p = reinterpret_cast<const double*>(s.raw) + 3 * (size_t)i;
        x = p[0]; y = p[1]; z = p[2];
    } else {
        const float* p = reinterpret_cast<const float*>(s.raw) + 3 * (size_t)i;
        x = (double)p[0]; y = (double)p[1]; z = (double)p[2];
    }
}

// keyframe.py:91-92: r2 = x**2 + y**2; (r2 < max_r**2) & (r2 > min_r**2) & (z > min_h) & (z < max_h)
__device__ __forceinline__ bool filter_pred(double x, double y, double z, const FilterParams& f) {
    const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
    return (r2 < f.max_r2) && (r2 > f.min_r2) && (z > f.min_h) && (z < f.max_h);
}

constexpr int kCompactBlock = 1024;

__global__ void __launch_bounds__(kCompactBlock) k_filter_count(const ScanDev* __restrict__ scans, FilterParams f) {
    const ScanDev& s = scans[blockIdx.y];
    if (blockIdx.x * kCompactBlock >= s.n_raw) return;
    const int i = blockIdx.x * kCompactBlock + threadIdx.x;
    bool pred = false;
    if (i < s.n_raw) {
        double x, y, z;
        load_raw(s, i, x, y, z);
        pred = filter_pred(x, y, z, f);
    }
    const int c = __syncthreads_count(pred);
    if (threadIdx.x == 0) s.blk[blockIdx.x] = c;
}

__global__ void __launch_bounds__(kCompactBlock) k_filter_scatter(const ScanDev* __restrict__ scans, FilterParams f, int voxel_on) {
    __shared__ int s_warp[33];
    __shared__ int s_base;
    const ScanDev& s = scans[blockIdx.y];
    if (blockIdx.x * kCompactBlock >= s.n_raw) return;
    const int nblk = (s.n_raw + kCompactBlock - 1) / kCompactBlock;
    if (threadIdx.x < 32) {
        const int base = warp_prefix_of_blocks(s.blk, blockIdx.x);
        if (threadIdx.x == 0) s_base = base;
        if (blockIdx.x == 0) {
            const int tot = warp_prefix_of_blocks(s.blk, nblk);
            if (threadIdx.x == 0) {
                s.counts[CNT_NFILT] = tot;
                if (!voxel_on) s.counts[CNT_NPTS] = tot;
            }
        }
    }
    const int i = blockIdx.x * kCompactBlock + threadIdx.x;
    bool pred = false;
    double x = 0, y = 0, z = 0;
    if (i < s.n_raw) {
        load_raw(s, i, x, y, z);
        pred = filter_pred(x, y, z, f);
    }
    int total;
    const int r = block_excl_scan_flag(pred, total, s_warp);
    if (pred) {
        const int pos = s_base + r;
        s.fx[pos] = x; s.fy[pos] = y; s.fz[pos] = z;
        s.raw_index[pos] = i;
    }
}

// ---------------------------------------------------------------------------------------------------
// stable LSD radix sort of (key, value) pairs, 8-bit digits, batched over scans
// ---------------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;

template <typename K>
__global__ void __launch_bounds__(kSortThreads) k_radix_hist(const ScanDev* __restrict__ scans, int cnt_index, int src, int shift) {
    __shared__ int h[256];
    const ScanDev& s = scans[blockIdx.y];
    const int n = s.counts[cnt_index];
    const int nblk = (n + kSortTile - 1) / kSortTile;
    if ((int)blockIdx.x >= nblk) return;
    h[threadIdx.x] = 0;
    __syncthreads();
    const K* keys = sizeof(K) == 8 ? reinterpret_cast<const K*>(s.key64[src]) : reinterpret_cast<const K*>(s.key32[src]);
    const int base = blockIdx.x * kSortTile;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int i = base + r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 255u], 1);
    }
    __syncthreads();
    s.hist[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of hist[0 .. 256*nblk) in place; one block per scan
__global__ void __launch_bounds__(1024) k_radix_scan(const ScanDev* __restrict__ scans, int cnt_index) {
    __shared__ int s_warp[33];
    __shared__ int s_carry;
    const ScanDev& s = scans[blockIdx.x];
    const int n = s.counts[cnt_index];
    const int nblk = (n + kSortTile - 1) / kSortTile;
    const int len = 256 * nblk;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = lane_id(), w = threadIdx.x >> 5;
    for (int base = 0; base < len; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < len ? s.hist[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[w] = inc;
        __syncthreads();
        if (w == 0) {
            const int wv = s_warp[lane];
            int winc = wv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, winc, o);
                if (lane >= o) winc += t;
            }
            s_warp[lane] = winc - wv;
            if (lane == 31) s_warp[32] = winc;
        }
        __syncthreads();
        const int carry = s_carry;
        if (i < len) s.hist[i] = carry + s_warp[w] + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + s_warp[32];
        __syncthreads();
    }
}

template <typename K>
__global__ void __launch_bounds__(kSortThreads) k_radix_scatter(const ScanDev* __restrict__ scans, int cnt_index, int src, int shift) {
    __shared__ int wh[kSortThreads / 32][256];
    __shared__ int gbase[256];
    const ScanDev& s = scans[blockIdx.y];
    const int n = s.counts[cnt_index];
    const int nblk = (n + kSortTile - 1) / kSortTile;
    if ((int)blockIdx.x >= nblk) return;
    const K* keys = sizeof(K) == 8 ? reinterpret_cast<const K*>(s.key64[src]) : reinterpret_cast<const K*>(s.key32[src]);
    K* okeys = sizeof(K) == 8 ? reinterpret_cast<K*>(s.key64[src ^ 1]) : reinterpret_cast<K*>(s.key32[src ^ 1]);
    const int* vals = s.val[src];
    int* ovals = s.val[src ^ 1];
    const int lane = lane_id(), w = threadIdx.x >> 5;
    for (int d = threadIdx.x; d < (kSortThreads / 32) * 256; d += kSortThreads) (&wh[0][0])[d] = 0;
    gbase[threadIdx.x] = s.hist[threadIdx.x * nblk + blockIdx.x];
    __syncthreads();
    // warp-major tile layout keeps the sort stable: warp w owns a contiguous chunk, processed in rounds of 32
    const int wbase = blockIdx.x * kSortTile + w * (32 * kSortItems);
    K key[kSortItems];
    int val[kSortItems], rank[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int i = wbase + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? keys[i] : (K)0;
        val[r] = valid ? vals[i] : 0;
        const unsigned digit = valid ? ((unsigned)(key[r] >> shift) & 255u) : (256u + lane);
        const unsigned peers = __match_any_sync(kFull, digit);
        const int leader = __ffs(peers) - 1;
        int old = 0;
        if (valid && lane == leader) {
            old = wh[w][digit];
            wh[w][digit] = old + __popc(peers);
        }
        old = __shfl_sync(kFull, old, leader);
        rank[r] = old + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    {   // exclusive prefix over warps for digit = threadIdx.x
        int acc = 0;
#pragma unroll
        for (int ww = 0; ww < kSortThreads / 32; ++ww) {
            const int t = wh[ww][threadIdx.x];
            wh[ww][threadIdx.x] = acc;
            acc += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int i = wbase + r * 32 + lane;
        if (i < n) {
            const unsigned digit = (unsigned)(key[r] >> shift) & 255u;
            const int dst = gbase[digit] + wh[w][digit] + rank[r];
            okeys[dst] = key[r];
            ovals[dst] = val[r];
        }
    }
}

// returns the index (0/1) of the buffer holding the sorted result
template <typename K>
static int radix_sort(Launcher& L, const ScanDev* d_scans, int n_scans, int cap_max, int cnt_index, int nbits) {
    const int nblk = (cap_max + kSortTile - 1) / kSortTile;
    int src = 0;
    for (int shift = 0; shift < nbits; shift += 8) {
        L.launch("radix_hist", k_radix_hist<K>, dim3(nblk, n_scans), dim3(kSortThreads), d_scans, cnt_index, src, shift);
        L.launch("radix_scan", k_radix_scan, dim3(n_scans), dim3(1024), d_scans, cnt_index);
        L.launch("radix_scatter", k_radix_scatter<K>, dim3(nblk, n_scans), dim3(kSortThreads), d_scans, cnt_index, src, shift);
        src ^= 1;
    }
    return src;
}

// ---------------------------------------------------------------------------------------------------
// voxel down-sampling
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long enc_min(double v) {   // order-preserving map double -> u64
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double dec_min(unsigned long long e) {
    const unsigned long long b = (e & 0x8000000000000000ULL) ? (e & 0x7fffffffffffffffULL) : ~e;
    return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(256) k_voxel_bbox(const ScanDev* __restrict__ scans) {
    const ScanDev& s = scans[blockIdx.y];
    const int n = s.counts[CNT_NFILT];
    double mx = INFINITY, my = INFINITY, mz = INFINITY;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        mx = fmin(mx, s.fx[i]); my = fmin(my, s.fy[i]); mz = fmin(mz, s.fz[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmin(mx, __shfl_xor_sync(kFull, mx, o));
        my = fmin(my, __shfl_xor_sync(kFull, my, o));
        mz = fmin(mz, __shfl_xor_sync(kFull, mz, o));
    }
    if (lane_id() == 0 && mx < INFINITY) {
        unsigned long long* b = reinterpret_cast<unsigned long long*>(s.bbox);
        atomicMin(b + 0, enc_min(mx)); atomicMin(b + 1, enc_min(my)); atomicMin(b + 2, enc_min(mz));
    }
}

// Open3D: voxel_min_bound = min_bound - voxel_size*0.5; ref = (p - voxel_min_bound)/voxel_size; index = int(floor(ref))
__global__ void __launch_bounds__(256) k_voxel_keys(const ScanDev* __restrict__ scans, VoxelParams vp) {
    const ScanDev& s = scans[blockIdx.y];
    const int n = s.counts[CNT_NFILT];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long* b = reinterpret_cast<const unsigned long long*>(s.bbox);
    const double half = __dmul_rn(vp.voxel, 0.5);
    const double ox = __dsub_rn(dec_min(b[0]), half), oy = __dsub_rn(dec_min(b[1]), half), oz = __dsub_rn(dec_min(b[2]), half);
    const double fxv = floor(__ddiv_rn(__dsub_rn(s.fx[i], ox), vp.voxel));
    const double fyv = floor(__ddiv_rn(__dsub_rn(s.fy[i], oy), vp.voxel));
    const double fzv = floor(__ddiv_rn(__dsub_rn(s.fz[i], oz), vp.voxel));
    long long ix = (long long)fxv, iy = (long long)fyv, iz = (long long)fzv;
    if (ix < 0 || iy < 0 || iz < 0 || ix >= (1ll << vp.bx) || iy >= (1ll << vp.by) || iz >= (1ll << vp.bz)) {
        atomicOr(&s.counts[CNT_ERR], ERR_VOXEL_RANGE);
        ix = min(max(ix, 0ll), (1ll << vp.bx) - 1); iy = min(max(iy, 0ll), (1ll << vp.by) - 1); iz = min(max(iz, 0ll), (1ll << vp.bz) - 1);
    }
    s.key64[0][i] = ((unsigned long long)ix << (vp.by + vp.bz)) | ((unsigned long long)iy << vp.bz) | (unsigned long long)iz;
    s.val[0][i] = i;
}

__global__ void __launch_bounds__(kCompactBlock) k_voxel_heads_count(const ScanDev* __restrict__ scans, int src) {
    const ScanDev& s = scans[blockIdx.y];
    const int n = s.counts[CNT_NFILT];
    if (blockIdx.x * kCompactBlock >= n) return;
    const int p = blockIdx.x * kCompactBlock + threadIdx.x;
    const unsigned long long* key = s.key64[src];
    const bool head = p < n && (p == 0 || key[p] != key[p - 1]);
    const int c = __syncthreads_count(head);
    if (threadIdx.x == 0) s.blk[blockIdx.x] = c;
}

// one thread per voxel head sums its run in ascending point index (stable sort) = Open3D's accumulation order
__global__ void __launch_bounds__(kCompactBlock) k_voxel_reduce(const ScanDev* __restrict__ scans, int src, VoxelParams vp) {
    __shared__ int s_warp[33];
    __shared__ int s_base;
    const ScanDev& s = scans[blockIdx.y];
    const int n = s.counts[CNT_NFILT];
    if (blockIdx.x * kCompactBlock >= n) return;
    const int nblk = (n + kCompactBlock - 1) / kCompactBlock;
    if (threadIdx.x < 32) {
        const int base = warp_prefix_of_blocks(s.blk, blockIdx.x);
        if (threadIdx.x == 0) s_base = base;
        if (blockIdx.x == 0) {
            const int tot = warp_prefix_of_blocks(s.blk, nblk);
            if (threadIdx.x == 0) s.counts[CNT_NPTS] = tot;
        }
    }
    const int p = blockIdx.x * kCompactBlock + threadIdx.x;
    const unsigned long long* key = s.key64[src];
    const int* val = s.val[src];
    const bool head = p < n && (p == 0 || key[p] != key[p - 1]);
    int total;
    const int r = block_excl_scan_flag(head, total, s_warp);
    if (head) {
        const int seg = s_base + r;
        const unsigned long long k = key[p];
        double sx = 0.0, sy = 0.0, sz = 0.0;
        int q = p;
        for (; q < n && key[q] == k; ++q) {
            const int i = val[q];
            sx = __dadd_rn(sx, s.fx[i]); sy = __dadd_rn(sy, s.fy[i]); sz = __dadd_rn(sz, s.fz[i]);
        }
        const double c = (double)(q - p);
        s.vx[seg] = __ddiv_rn(sx, c); s.vy[seg] = __ddiv_rn(sy, c); s.vz[seg] = __ddiv_rn(sz, c);
        s.vox_counts[seg] = q - p;
        s.vox_keys[3 * seg + 0] = (int)(k >> (vp.by + vp.bz));
        s.vox_keys[3 * seg + 1] = (int)((k >> vp.bz) & ((1ull << vp.by) - 1));
        s.vox_keys[3 * seg + 2] = (int)(k & ((1ull << vp.bz) - 1));
    }
}

// ---------------------------------------------------------------------------------------------------
// Morton sort + multi-resolution hash grid
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_morton_keys(const ScanDev* __restrict__ scans, int voxel_on) {
    const ScanDev& s = scans[blockIdx.y];
    const int n = s.counts[CNT_NPTS];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = voxel_on ? s.vx[i] : s.fx[i], y = voxel_on ? s.vy[i] : s.fy[i], z = voxel_on ? s.vz[i] : s.fz[i];
    const GridSpec& g = s.grid;
    s.key32[0][i] = morton3(cell_coord(x, g.ox, g.inv_c0), cell_coord(y, g.oy, g.inv_c0), cell_coord(z, g.oz, g.inv_c0));
    s.val[0][i] = i;
}

__global__ void __launch_bounds__(256) k_gather_build(const ScanDev* __restrict__ scans, int src, int voxel_on) {
    const ScanDev& s = scans[blockIdx.y];
    const int n = s.counts[CNT_NPTS];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const unsigned* key = s.key32[src];
    const int k = s.val[src][p];
    const double x = voxel_on ? s.vx[k] : s.fx[k], y = voxel_on ? s.vy[k] : s.fy[k], z = voxel_on ? s.vz[k] : s.fz[k];
    if (s.wide) {
        RecD r; r.x = x; r.y = y; r.z = z; r.idx = k; r.pad = 0;
        reinterpret_cast<RecD*>(s.recs)[p] = r;
    } else {
        RecF r; r.x = (float)x; r.y = (float)y; r.z = (float)z; r.idx = k;   // exact: float32 payload
        reinterpret_cast<RecF*>(s.recs)[p] = r;
    }
    // hash-grid insertion: this thread inserts every level whose cell starts at p
    const unsigned m = key[p];
    const unsigned prev = p > 0 ? key[p - 1] : 0u;
    const unsigned mask = s.table_mask;
    for (int l = 0; l <= s.grid.top_level; ++l) {
        const unsigned prefix = m >> (3 * l);                         // 3*l <= 30
        if (p > 0 && prefix == (prev >> (3 * l))) break;             // not a boundary here nor at any coarser level
        // end of the run: first position whose prefix exceeds ours
        int lo = p + 1, hi = n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((key[mid] >> (3 * l)) > prefix) hi = mid; else lo = mid + 1;
        }
        const unsigned long long ck = cell_key(l, prefix);
        unsigned slot = hash_key(ck) & mask;
        bool ok = false;
        for (unsigned probe = 0; probe <= mask; ++probe) {
            const unsigned long long old = atomicCAS(&s.table[slot].key, 0ull, ck);
            if (old == 0ull) { s.table[slot].start = (unsigned)p; s.table[slot].end = (unsigned)lo; ok = true; break; }
            slot = (slot + 1) & mask;
        }
        if (!ok) atomicOr(&s.counts[CNT_ERR], ERR_HASH_FULL);
        atomicAdd(&s.counts[CNT_NCELLS], 1);
    }
}

// ---------------------------------------------------------------------------------------------------
// host-side orchestration of one preprocessing batch (everything queued on L.stream, no sync)
// ---------------------------------------------------------------------------------------------------
void run_preprocess(Launcher& L, const ScanDev* d_scans, int n_scans, int cap_max, const FilterParams& fp, const VoxelParams& vp,
                    bool voxel_on) {
    if (n_scans == 0 || cap_max == 0) return;
    const int nb1024 = (cap_max + kCompactBlock - 1) / kCompactBlock;
    const int nb256 = (cap_max + 255) / 256;
    L.launch("filter_count", k_filter_count, dim3(nb1024, n_scans), dim3(kCompactBlock), d_scans, fp);
    L.launch("filter_scatter", k_filter_scatter, dim3(nb1024, n_scans), dim3(kCompactBlock), d_scans, fp, (int)voxel_on);
    if (voxel_on) {
        L.launch("voxel_bbox", k_voxel_bbox, dim3(min(nb256, 64), n_scans), dim3(256), d_scans);
        L.launch("voxel_keys", k_voxel_keys, dim3(nb256, n_scans), dim3(256), d_scans, vp);
        const int src = radix_sort<unsigned long long>(L, d_scans, n_scans, cap_max, CNT_NFILT, vp.bx + vp.by + vp.bz);
        L.launch("voxel_heads_count", k_voxel_heads_count, dim3(nb1024, n_scans), dim3(kCompactBlock), d_scans, src);
        L.launch("voxel_reduce", k_voxel_reduce, dim3(nb1024, n_scans), dim3(kCompactBlock), d_scans, src, vp);
    }
    L.launch("morton_keys", k_morton_keys, dim3(nb256, n_scans), dim3(256), d_scans, (int)voxel_on);
    const int src = radix_sort<unsigned>(L, d_scans, n_scans, cap_max, CNT_NPTS, 3 * kMortonBits);
    L.launch("gather_build", k_gather_build, dim3(nb256, n_scans), dim3(256), d_scans, src, (int)voxel_on);
}

}  // namespace arvc
