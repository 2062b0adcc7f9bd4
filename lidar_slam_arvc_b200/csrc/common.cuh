// Shared device-side types and helpers of the sm_100a ICP engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace arvc {

constexpr int kMortonBits = 10;           // bits per axis at the finest level -> 1024^3 cells
constexpr int kMaxLevels = kMortonBits + 1;
constexpr unsigned kFull = 0xffffffffu;

// ---- point records (sorted by Morton code of the finest grid cell) --------------------------------
// Narrow: the PCD payload is float32, exactly representable -> 16-byte record, one LDG.128.
// Wide:   voxel means / float64 uploads -> 32-byte record.
struct __align__(16) RecF { float x, y, z; int idx; };
struct __align__(16) RecD { double x, y, z; int idx; int pad; };

template <bool WIDE> struct RecT;
template <> struct RecT<false> { typedef RecF type; };
template <> struct RecT<true> { typedef RecD type; };

__device__ __forceinline__ void load_rec(const RecF* p, double& x, double& y, double& z, int& idx) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    x = (double)v.x; y = (double)v.y; z = (double)v.z; idx = __float_as_int(v.w);
}
__device__ __forceinline__ void load_rec(const RecD* p, double& x, double& y, double& z, int& idx) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    x = a.x; y = a.y; z = b.x; idx = __double2loint(b.y);
}

// Squared distance exactly as the oracle / nanoflann evaluate it: ((dx*dx) + dy*dy) + dz*dz, every
// operation rounded separately (no FMA contraction), so equal inputs give bit-equal distances.
__device__ __forceinline__ double sqdist(double ax, double ay, double az, double bx, double by, double bz) {
    const double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// t -> (t % nx, (t / nx) % ny, t / (nx * ny)) for the small cell blocks of the searches (t < 4096, nx, ny <= 64) without
// integer division: (t + 0.5) / n is at least 0.5 / n away from an integer, far beyond the error of the float reciprocal.
struct CellDecoder {
    int nx, nxy;
    float inx, inxy;
    __device__ __forceinline__ CellDecoder(int nx_, int ny_) : nx(nx_), nxy(nx_ * ny_) {
        inx = __frcp_rn((float)nx);
        inxy = __frcp_rn((float)nxy);
    }
    __device__ __forceinline__ void operator()(int t, int& ox, int& oy, int& oz) const {
        oz = (int)(((float)t + 0.5f) * inxy);
        const int rem = t - oz * nxy;
        oy = (int)(((float)rem + 0.5f) * inx);
        ox = rem - oy * nx;
    }
};

// ---- multi-resolution Morton hash grid ---------------------------------------------------------------
// Level l has cubic cells of edge c0 * 2^l; a level-l cell is the Morton prefix (code >> 3l).  Because the
// records are sorted by Morton code, every cell of every level is one contiguous run [start, end).
struct GridSpec {
    double ox, oy, oz;   // grid origin
    double inv_c0;       // 1 / finest cell edge
    double c0;
    int top_level;       // coarsest level a search may use
    int pad;
};

struct __align__(16) HashEntry { unsigned long long key; unsigned start; unsigned end; };

__host__ __device__ __forceinline__ unsigned long long cell_key(int level, unsigned prefix) {
    return ((unsigned long long)(level + 1) << 32) | (unsigned long long)prefix;
}
__host__ __device__ __forceinline__ unsigned hash_key(unsigned long long k) {
    // 32-bit mix of (Morton prefix, level): a handful of integer instructions instead of a 64-bit murmur finaliser
    unsigned h = (unsigned)k * 0x9E3779B1u ^ ((unsigned)(k >> 32) * 0x85EBCA6Bu);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}

__host__ __device__ __forceinline__ unsigned spread3(unsigned v) {   // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__host__ __device__ __forceinline__ unsigned morton3(unsigned ix, unsigned iy, unsigned iz) {
    return (spread3(ix) << 2) | (spread3(iy) << 1) | spread3(iz);
}

// finest-level integer cell coordinate of a coordinate value; monotone non-decreasing in v, clamped.
__device__ __forceinline__ int cell_coord(double v, double o, double inv_c0) {
    const double f = floor(__dmul_rn(__dsub_rn(v, o), inv_c0));
    const double hi = (double)((1 << kMortonBits) - 1);
    return (int)fmin(fmax(f, 0.0), hi);          // NaN -> 0 via fmax
}

__device__ __forceinline__ bool grid_lookup(const HashEntry* __restrict__ tab, unsigned mask, int level, unsigned prefix,
                                            unsigned& start, unsigned& end) {
    const unsigned long long key = cell_key(level, prefix);
    unsigned slot = hash_key(key) & mask;
    for (unsigned probe = 0; probe <= mask; ++probe) {
        const uint4 e = __ldg(reinterpret_cast<const uint4*>(tab + slot));
        const unsigned long long k = ((unsigned long long)e.y << 32) | e.x;
        if (k == key) { start = e.z; end = e.w; return true; }
        if (k == 0ull) return false;
        slot = (slot + 1) & mask;
    }
    return false;
}

// ---- per-scan device view ------------------------------------------------------------------------------
enum { CNT_NFILT = 0, CNT_NPTS = 1, CNT_ERR = 2, CNT_NCELLS = 3, CNT_NREDO = 4,
       CNT_NFB = 5,             // points on the normals fallback list (served by the per-point kernel)
       CNT_NFB_BLOCKS = 6,      // statistics: blocks of k_normals_blk that fell back as a whole
       CNT_NFB_POINTS = 7,      //             single points that fell back (trial radius too small, boundary overflow)
       CNT_NTRIAL_BLOCKS = 8,   //             blocks served at a trial radius
       CNT_WORDS = 16 };
enum { ERR_HASH_FULL = 1, ERR_VOXEL_RANGE = 2 };

struct ScanDev {
    // input
    const void* raw;        // float[3n] or double[3n]
    int n_raw;
    int raw_f64;
    int cap;                // capacity of every per-point buffer (= n_raw)
    int wide;               // records are RecD
    int* counts;            // device counters [CNT_WORDS]
    // persistent outputs
    void* recs;             // RecF[cap] or RecD[cap], Morton order
    double* normals;        // [cap][4], Morton order (xyz + pad)
    int* nn_count;          // [cap], Morton order
    HashEntry* table;
    unsigned table_mask;
    int* raw_index;         // [cap] raw index of filtered point k (filter order)
    int* vox_keys;          // [cap][3] voxel index of output point k (voxel mode)
    int* vox_counts;        // [cap]
    int* tap_idx;           // parity tap (ctx option "normals_tap"): [cap][max_nn] cloud indices of the neighbours each normal used
    int* tap_cnt;           //                                         [cap] how many of them (Morton order), else null
    int tap_stride;         // = max_nn of the preprocessing the tap was recorded with
    GridSpec grid;
    // scratch (valid during preprocessing only)
    double* fx; double* fy; double* fz;      // filtered cloud, filter order
    double* vx; double* vy; double* vz;      // voxel cloud, key order
    unsigned long long* key64[2];
    unsigned* key32[2];
    int* val[2];
    int* hist;              // [256][nblk]
    int* blk;               // [nblk1024 + 8] block counters
    double* bbox;           // [8] min bound / origin of the voxel grid
    double* moments;        // [cap][10] centred moment sums + neighbour count per point (normals)
    int* redo_list;         // [cap] points whose normal needs the canonical re-summation
    int* fb_list;           // [cap] points the block kernel hands to the per-point kernel
};

// ---- warp helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

}  // namespace arvc
