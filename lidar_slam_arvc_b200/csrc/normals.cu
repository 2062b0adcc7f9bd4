// Normal estimation: hybrid k-NN (k nearest, then d2 < r2) over the hash grid, per-point 3x3 covariance
// and analytic smallest-eigenvector solve in registers.  One warp per point.
//
// Reference semantics restated: keyframemanager/keyframe.py:160-162
//   pointcloud_filtered.estimate_normals(KDTreeSearchParamHybrid(radius=0.3, max_nn=300))
// -> Open3D EstimatePerPointCovariances (SearchHybrid, ComputeCovariance) + ComputeNormal (FastEigen3x3).
// Selection of the k nearest among more than k in-radius neighbours is exact: a monotone d2-histogram
// finds the bucket holding the k-th neighbour, the bucket is ranked exactly by (d2, index).
#include <cstdio>
#include <cstdlib>
#include <string>

#include "engine.cuh"

namespace arvc {

namespace {

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 cross3(const V3& a, const V3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double dot3(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

struct Sym3 { double a00, a01, a02, a11, a12, a22; };

__device__ V3 eigvec0(const Sym3& A, double ev) {
    const V3 r0{A.a00 - ev, A.a01, A.a02}, r1{A.a01, A.a11 - ev, A.a12}, r2{A.a02, A.a12, A.a22 - ev};
    const V3 c01 = cross3(r0, r1), c02 = cross3(r0, r2), c12 = cross3(r1, r2);
    const double d0 = dot3(c01, c01), d1 = dot3(c02, c02), d2 = dot3(c12, c12);
    double dmax = d0;
    int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) { imax = 2; }
    const V3 c = imax == 0 ? c01 : (imax == 1 ? c02 : c12);
    const double s = sqrt(imax == 0 ? d0 : (imax == 1 ? d1 : d2));
    return {c.x / s, c.y / s, c.z / s};
}

__device__ V3 eigvec1(const Sym3& A, const V3& e0, double ev1) {
    V3 U;
    if (fabs(e0.x) > fabs(e0.y)) {
        const double inv = 1.0 / sqrt(e0.x * e0.x + e0.z * e0.z);
        U = {-e0.z * inv, 0.0, e0.x * inv};
    } else {
        const double inv = 1.0 / sqrt(e0.y * e0.y + e0.z * e0.z);
        U = {0.0, e0.z * inv, -e0.y * inv};
    }
    const V3 V = cross3(e0, U);
    const V3 AU{A.a00 * U.x + A.a01 * U.y + A.a02 * U.z, A.a01 * U.x + A.a11 * U.y + A.a12 * U.z, A.a02 * U.x + A.a12 * U.y + A.a22 * U.z};
    const V3 AV{A.a00 * V.x + A.a01 * V.y + A.a02 * V.z, A.a01 * V.x + A.a11 * V.y + A.a12 * V.z, A.a02 * V.x + A.a12 * V.y + A.a22 * V.z};
    double m00 = dot3(U, AU) - ev1, m01 = dot3(U, AV), m11 = dot3(V, AV) - ev1;
    const double a00 = fabs(m00), a01 = fabs(m01), a11 = fabs(m11);
    if (a00 >= a11) {
        if (fmax(a00, a01) > 0) {
            if (a00 >= a01) { m01 /= m00; m00 = 1 / sqrt(1 + m01 * m01); m01 *= m00; }
            else            { m00 /= m01; m01 = 1 / sqrt(1 + m00 * m00); m00 *= m01; }
            return {m01 * U.x - m00 * V.x, m01 * U.y - m00 * V.y, m01 * U.z - m00 * V.z};
        }
        return U;
    }
    if (fmax(a11, a01) > 0) {
        if (a11 >= a01) { m01 /= m11; m11 = 1 / sqrt(1 + m01 * m01); m01 *= m11; }
        else            { m11 /= m01; m01 = 1 / sqrt(1 + m11 * m11); m11 *= m01; }
        return {m11 * U.x - m01 * V.x, m11 * U.y - m01 * V.y, m11 * U.z - m01 * V.z};
    }
    return U;
}

// covariance (symmetric, c00 c01 c02 c11 c12 c22) -> smallest-eigenvalue eigenvector, Open3D FastEigen3x3 scheme
__device__ V3 fast_eigen3x3(double c00, double c01, double c02, double c11, double c12, double c22, double* rel_gap) {
    const double mx = fmax(fmax(fmax(c00, c01), fmax(c02, c11)), fmax(c12, c22));   // maxCoeff (not max-abs)
    *rel_gap = 0.0;
    if (mx == 0) return {0, 0, 0};
    const Sym3 A{c00 / mx, c01 / mx, c02 / mx, c11 / mx, c12 / mx, c22 / mx};
    const double norm = A.a01 * A.a01 + A.a02 * A.a02 + A.a12 * A.a12;
    if (norm > 0) {
        const double q = (A.a00 + A.a11 + A.a22) / 3;
        const double b00 = A.a00 - q, b11 = A.a11 - q, b22 = A.a22 - q;
        const double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2) / 6);
        const double k00 = b11 * b22 - A.a12 * A.a12;
        const double k01 = A.a01 * b22 - A.a12 * A.a02;
        const double k02 = A.a01 * A.a12 - b11 * A.a02;
        const double det = (b00 * k00 - A.a01 * k01 + A.a02 * k02) / (p * p * p);
        double half_det = det * 0.5;
        half_det = fmin(fmax(half_det, -1.0), 1.0);
        const double angle = acos(half_det) / 3.0;
        const double two_thirds_pi = 2.09439510239319549;
        const double beta2 = cos(angle) * 2;
        const double beta0 = cos(angle + two_thirds_pi) * 2;
        const double beta1 = -(beta0 + beta2);
        const double e0 = q + p * beta0, e1 = q + p * beta1, e2 = q + p * beta2;
        *rel_gap = (e1 - e0) / fmax(fmax(fabs(e0), fabs(e2)), 1e-300);
        if (half_det >= 0) {
            const V3 v2 = eigvec0(A, e2);
            if (e2 < e0 && e2 < e1) return v2;
            const V3 v1 = eigvec1(A, v2, e1);
            if (e1 < e0 && e1 < e2) return v1;
            return cross3(v1, v2);
        }
        const V3 v0 = eigvec0(A, e0);
        if (e0 < e1 && e0 < e2) return v0;
        const V3 v1 = eigvec1(A, v0, e1);
        if (e1 < e0 && e1 < e2) return v1;
        return cross3(v0, v1);
    }
    {
        const double lo = fmin(c00, fmin(c11, c22)), hi = fmax(c00, fmax(c11, c22)), mid = c00 + c11 + c22 - lo - hi;
        *rel_gap = (mid - lo) / fmax(fmax(fabs(lo), fabs(hi)), 1e-300);
    }
    if (c00 < c11 && c00 < c22) return {1, 0, 0};
    if (c11 < c00 && c11 < c22) return {0, 1, 0};
    return {0, 0, 1};
}

constexpr int kNrmWarps = 8;
constexpr int kBins = 512;
constexpr int kCand = 128;
constexpr int kCanon = 320;          // most neighbours the canonical (sorted, sequential) re-summation handles
// per-warp shared memory, two layouts that are never live at the same time:
//   selection : int hist[kBins] | double cand_d2[kCand] | int cand_idx[kCand] | int cand_pos[kCand]
//   canonical : double key_d2[kCanon] | int key_idx[kCanon] | int key_pos[kCanon] | int order[kCanon]
constexpr int kScratch = kCanon * 20;  // bytes of the two overlapping layouts above
constexpr int kTryRuns = 160;          // cells of the trial ball (6^3 = 216 before box pruning)
constexpr int kScratchSel = (kBins * 4 + kCand * 16 + 16 + 255) / 256 * 256;   // mode 0 never uses the canonical layout
constexpr int kRunBytes = (64 + kTryRuns) * 8;              // cell runs, which outlive both layouts
constexpr int kWarpSmem = kScratch + kRunBytes;             // mode 1
constexpr int kWarpSmemSel = kScratchSel + kRunBytes;       // mode 0: 5.9 KB per warp instead of 8 KB leaves more L1
constexpr int kSpecFactor3 = 7;        // up to kSpecFactor3/3 * max_nn candidates: speculative single pass at the full radius
constexpr int kDenseFactor2 = 3;       // neighbourhoods with more than kDenseFactor2/2 * max_nn candidates try a smaller radius first
static_assert(kBins * 4 + kCand * 16 + 16 <= kScratch, "selection layout must fit");
constexpr double kIllGap = 2e-3;     // below this relative eigen-gap the normal is recomputed in canonical order

__device__ __forceinline__ bool key_less(double d2a, int ia, double d2b, int ib) { return d2a < d2b || (d2a == d2b && ia < ib); }

}  // namespace

// One candidate record, evaluated lazily: a float32 distance classifies most candidates (narrow records are exact
// float32 values, so d2f = d2 * (1 + theta), |theta| < 1e-6); the float64 distance - bit-equal to the oracle's -
// is only computed for candidates that are used or that sit within that error of a decision boundary.
template <bool WIDE> struct CandEval;

template <> struct CandEval<false> {
    float4 v;
    float d2f;
    __device__ __forceinline__ void load(const RecF* recs, unsigned j, float qxf, float qyf, float qzf, double, double, double) {
        v = __ldg(reinterpret_cast<const float4*>(recs + j));
        const float dx = qxf - v.x, dy = qyf - v.y, dz = qzf - v.z;
        d2f = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    }
    __device__ __forceinline__ double exact(double qx, double qy, double qz) const { return sqdist(qx, qy, qz, (double)v.x, (double)v.y, (double)v.z); }
    __device__ __forceinline__ double x() const { return (double)v.x; }
    __device__ __forceinline__ double y() const { return (double)v.y; }
    __device__ __forceinline__ double z() const { return (double)v.z; }
    __device__ __forceinline__ int idx() const { return __float_as_int(v.w); }
    // -1: certainly d2 >= bound, +1: certainly d2 < bound, 0: too close to call in float32
    __device__ __forceinline__ int below(float lo, float hi) const { return d2f > hi ? -1 : (d2f < lo ? 1 : 0); }
};

template <> struct CandEval<true> {
    double px, py, pz, d2;
    int id;
    __device__ __forceinline__ void load(const RecD* recs, unsigned j, float, float, float, double qx, double qy, double qz) {
        load_rec(recs + j, px, py, pz, id);
        d2 = sqdist(qx, qy, qz, px, py, pz);
    }
    __device__ __forceinline__ double exact(double, double, double) const { return d2; }
    __device__ __forceinline__ double x() const { return px; }
    __device__ __forceinline__ double y() const { return py; }
    __device__ __forceinline__ double z() const { return pz; }
    __device__ __forceinline__ int idx() const { return id; }
    __device__ __forceinline__ int below(float, float) const { return 0; }
};

// mode 0: neighbour search + selection for every point; the warp stores the centred moment sums and leaves the
//         eigen-solve to k_normals_eigen (one THREAD per point - the solve is scalar work);
// mode 1: only the points k_normals_eigen flagged as ill-conditioned: same search, then the canonical re-summation
//         and the solve inside the warp.
// TAP: record which neighbours were used (parity tap of arvc_scan_get_neighbors); a separate instantiation, so that the
// production kernels carry none of it.
template <bool WIDE, bool TAP>
__device__ __forceinline__ void normals_point(const ScanDev& s, const NormalParams& np, const int mode, const int p, unsigned char* s_raw) {
    typedef typename RecT<WIDE>::type Rec;
    const int w = threadIdx.x >> 5, lane = lane_id();
    const Rec* __restrict__ recs = reinterpret_cast<const Rec*>(s.recs);
    double qx, qy, qz;
    int qidx;
    load_rec(recs + p, qx, qy, qz, qidx);
    const float qxf = (float)qx, qyf = (float)qy, qzf = (float)qz;

    const GridSpec g = s.grid;
    const int L = np.level;
    const double r2 = np.r2;
    const float r2_lo = np.r2_lo, r2_hi = np.r2_hi;
    const double rinf = np.radius * (1.0 + 1e-9) + 1e-12;
    const double cl = g.c0 * (double)(1 << L);
    const int x0 = cell_coord(qx - rinf, g.ox, g.inv_c0) >> L, x1 = cell_coord(qx + rinf, g.ox, g.inv_c0) >> L;
    const int y0 = cell_coord(qy - rinf, g.oy, g.inv_c0) >> L, y1 = cell_coord(qy + rinf, g.oy, g.inv_c0) >> L;
    const int z0 = cell_coord(qz - rinf, g.oz, g.inv_c0) >> L, z1 = cell_coord(qz + rinf, g.oz, g.inv_c0) >> L;
    const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, nz = z1 - z0 + 1;
    const int ncell = nx * ny * nz;            // <= 27 (64 when the inflated ball grazes a fourth cell): cell edge >= radius

    unsigned char* wmem = s_raw + (size_t)w * (mode == 0 ? kWarpSmemSel : kWarpSmem);
    int* hist = reinterpret_cast<int*>(wmem);
    double* cand_d2 = reinterpret_cast<double*>(wmem + kBins * 4);
    int* cand_idx = reinterpret_cast<int*>(wmem + kBins * 4 + kCand * 8);
    int* cand_pos = reinterpret_cast<int*>(wmem + kBins * 4 + kCand * 12);
    uint2* runs_full = reinterpret_cast<uint2*>(wmem + (mode == 0 ? kScratchSel : kScratch));   // <= 64 runs: cells of the full radius
    uint2* runs_try = runs_full + 64;                                       // <= kTryRuns runs: finer cells of the trial radius

    // ---- cells of the full radius (<= 27 at the level whose edge >= radius): lane c owns cell c
    int nfull = 0, total = 0;
    const CellDecoder dec(nx, ny);
    for (int base = 0; base < ncell; base += 32) {      // one round unless the ball grazes a fourth cell along an axis
        const int t = base + lane;
        unsigned st = 0, en = 0;
        bool valid = false;
        if (t < ncell) {
            int cx, cy, cz;
            dec(t, cx, cy, cz);
            cx += x0; cy += y0; cz += z0;
            const double bx0 = g.ox + cx * cl, by0 = g.oy + cy * cl, bz0 = g.oz + cz * cl;
            const double ddx = fmax(0.0, fmax(bx0 - qx, qx - (bx0 + cl)));
            const double ddy = fmax(0.0, fmax(by0 - qy, qy - (by0 + cl)));
            const double ddz = fmax(0.0, fmax(bz0 - qz, qz - (bz0 + cl)));
            if (ddx * ddx + ddy * ddy + ddz * ddz <= r2 * (1.0 + 1e-9) + 1e-12)
                valid = grid_lookup(s.table, s.table_mask, L, morton3(cx, cy, cz), st, en);
        }
        const unsigned vm = __ballot_sync(kFull, valid);
        if (valid) runs_full[nfull + __popc(vm & ((1u << lane) - 1u))] = make_uint2(st, en);
        nfull += __popc(vm);
        total += warp_sum((int)(en - st));
    }
    // ---- dense neighbourhoods: the k nearest lie well inside the radius.  Estimate their radius from the local
    // density (k / total of the 3x3x3 block, surface model), cover that smaller ball with cells one level finer and
    // let the histogram pass verify the guess: if it holds fewer than k points, fall back to the full radius.
    int ntry = 0, total_try = 0;
    double rtry = 0.0;
    if (2 * total > kDenseFactor2 * np.max_nn && L > 0) {
        rtry = 1.1 * 3.0 * cl * sqrt((double)np.max_nn / (3.141592653589793 * (double)total));     // 10 % safety on the density model
        if (rtry < 0.95 * np.radius) {
            int Lf = 0;                                   // finest level whose cells are at least half the trial radius
            while (Lf < L - 1 && g.c0 * (double)(1 << Lf) < 0.5 * rtry) ++Lf;
            const double cf = g.c0 * (double)(1 << Lf), rt = rtry * (1.0 + 1e-9) + 1e-12, rt2 = rtry * rtry;
            const int fx0 = cell_coord(qx - rt, g.ox, g.inv_c0) >> Lf, fx1 = cell_coord(qx + rt, g.ox, g.inv_c0) >> Lf;
            const int fy0 = cell_coord(qy - rt, g.oy, g.inv_c0) >> Lf, fy1 = cell_coord(qy + rt, g.oy, g.inv_c0) >> Lf;
            const int fz0 = cell_coord(qz - rt, g.oz, g.inv_c0) >> Lf, fz1 = cell_coord(qz + rt, g.oz, g.inv_c0) >> Lf;
            const int fnx = fx1 - fx0 + 1, fny = fy1 - fy0 + 1, fncell = fnx * fny * (fz1 - fz0 + 1);
            const CellDecoder fdec(fnx, fny);
            for (int base = 0; base < fncell; base += 32) {
                const int t = base + lane;
                unsigned st = 0, en = 0;
                bool valid = false;
                if (t < fncell) {
                    int cx, cy, cz;
                    fdec(t, cx, cy, cz);
                    cx += fx0; cy += fy0; cz += fz0;
                    const double bx0 = g.ox + cx * cf, by0 = g.oy + cy * cf, bz0 = g.oz + cz * cf;
                    const double ddx = fmax(0.0, fmax(bx0 - qx, qx - (bx0 + cf)));
                    const double ddy = fmax(0.0, fmax(by0 - qy, qy - (by0 + cf)));
                    const double ddz = fmax(0.0, fmax(bz0 - qz, qz - (bz0 + cf)));
                    if (ddx * ddx + ddy * ddy + ddz * ddz <= rt2 * (1.0 + 1e-9) + 1e-12)
                        valid = grid_lookup(s.table, s.table_mask, Lf, morton3(cx, cy, cz), st, en);
                }
                const unsigned vm = __ballot_sync(kFull, valid);
                const int slot = ntry + __popc(vm & ((1u << lane) - 1u));
                if (valid && slot < kTryRuns) runs_try[slot] = make_uint2(st, en);
                ntry += __popc(vm);
                total_try += warp_sum((int)(en - st));
            }
            if (ntry > kTryRuns || total_try <= np.max_nn) ntry = 0;      // does not fit / cannot hold k points: full radius
        }
    }
    __syncwarp();

    double sx = 0, sy = 0, sz = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
    int cnt = 0, dbg_nin = -1;
    double tau_d2 = r2;      // selected <=> d2 < rq2 and (d2, idx) <= (tau_d2, tau_idx)
    int tau_idx = 0x7fffffff;
    double rq2 = r2;         // radius^2 the neighbourhood was finally searched with
    const uint2* runs = runs_full;
    int nruns = nfull;
    if (np.debug & 2) { nfull = 0; ntry = 0; total = 0; }      // ablation: fixed per-point overhead only
    const bool tap = TAP && mode != 1 && s.tap_idx != nullptr;      // the canonical re-summation (mode 1) uses the same set
    auto tap_reset = [&]() {
        if (tap) { __syncwarp(); if (lane == 0) s.tap_cnt[p] = 0; __syncwarp(); }
    };
    tap_reset();
    auto accumulate = [&](double x, double y, double z, int idx) {
        if (tap) {
            const int slot = atomicAdd(&s.tap_cnt[p], 1);
            if (slot < s.tap_stride) s.tap_idx[(size_t)p * s.tap_stride + slot] = idx;
        }
        const double ux = x - qx, uy = y - qy, uz = z - qz;      // centred: exact differences of float32 payloads
        sx += ux; sy += uy; sz += uz;
        sxx = fma(ux, ux, sxx); sxy = fma(ux, uy, sxy); sxz = fma(ux, uz, sxz);
        syy = fma(uy, uy, syy); syz = fma(uy, uz, syz); szz = fma(uz, uz, szz);
        ++cnt;
    };

    for (int attempt = (ntry > 0 ? 0 : 1); attempt < 2; ++attempt) {
        const bool trial = attempt == 0;
        runs = trial ? runs_try : runs_full;
        nruns = trial ? ntry : nfull;
        rq2 = trial ? rtry * rtry : r2;
        const int tot = trial ? total_try : total;
        const double bin_scale = trial ? (double)kBins / rq2 : np.bin_scale;
        const float bin_scale_f = (float)bin_scale;
        float rq2_lo = trial ? (float)(rq2 * (1.0 - 2e-6)) : np.r2_lo, rq2_hi = trial ? (float)(rq2 * (1.0 + 2e-6)) : np.r2_hi;
        float bsf = bin_scale_f;
        // pin the loop-invariant screening constants in registers (the compiler otherwise re-derives them from the
        // float64 radius inside the candidate loops: two DMUL + two F2F per iteration)
        asm volatile("" : "+f"(rq2_lo), "+f"(rq2_hi), "+f"(bsf));
        // d2 -> bucket, monotone in the exact d2; the float32 shortcut is taken only when it cannot cross a bucket edge
        auto bucket = [&](const CandEval<WIDE>& c, double& d2, bool& have) -> int {
            if constexpr (!WIDE) {
                const float u = c.d2f * bsf;
                const int b = (int)u;
                const float fr = u - (float)b;
                if (fr > 2e-3f && fr < 1.0f - 2e-3f && b < kBins - 1) return b;
            }
            if (!have) { d2 = c.exact(qx, qy, qz); have = true; }
            return min(kBins - 1, (int)(d2 * bin_scale));
        };
        auto in_radius = [&](const CandEval<WIDE>& c, double& d2, bool& have) -> bool {      // exact membership
            const int t = c.below(rq2_lo, rq2_hi);
            if (t != 0) return t > 0;
            if (!have) { d2 = c.exact(qx, qy, qz); have = true; }
            return d2 < rq2;
        };

        // ---- pass B: accumulate the buckets below bstar, collect bucket bstar for exact ranking
        auto pass_b = [&](const int bstar) -> int {
            int ncand = 0;
            // narrow records: the float32 histogram may have put a candidate one bucket off, so the k-th nearest lies in
            // exact bucket bstar-1 .. bstar+1: everything below is taken whole, those three buckets are ranked exactly
            const int blo = bstar == kBins ? kBins : (WIDE ? bstar : bstar - 1), bhi = WIDE ? bstar : bstar + 1;
            if constexpr (!WIDE) {
                // float32 bucket coordinate u = d2f * buckets / r^2 (absolute error < 1e-3): two compares settle almost every
                // candidate - certainly below bucket bstar (taken), certainly above it (dropped); the rest is decided exactly
                int* ncand_s = reinterpret_cast<int*>(wmem + kBins * 4 + kCand * 16);
                if (lane == 0) *ncand_s = 0;
                __syncwarp();
                const float u_lo = (float)blo - 2e-3f, u_hi = (float)(bstar == kBins ? kBins : bhi + 1) + 2e-3f;
                auto take_one = [&](const CandEval<WIDE>& c, unsigned j) {
                    const float u = c.d2f * bsf;
                    if (u < u_lo) {
                        accumulate(c.x(), c.y(), c.z(), c.idx());
                    } else if (u <= u_hi) {
                        const double d2 = c.exact(qx, qy, qz);
                        if (d2 < rq2) {
                            const int b = min(kBins - 1, (int)(d2 * bin_scale));
                            if (b < blo) {
                                accumulate(c.x(), c.y(), c.z(), c.idx());
                            } else if (b <= bhi) {
                                const int slot = atomicAdd(ncand_s, 1);
                                if (slot < kCand) { cand_d2[slot] = d2; cand_idx[slot] = c.idx(); cand_pos[slot] = (int)j; }
                            }
                        }
                    }
                };
                for (int rr = 0; rr < nruns; ++rr) {
                    const uint2 run = runs[rr];
                    for (unsigned j = run.x + lane; j < run.y; j += 64) {      // two records in flight per lane
                        CandEval<WIDE> c0, c1;
                        const bool two = j + 32 < run.y;
                        c0.load(recs, j, qxf, qyf, qzf, qx, qy, qz);
                        if (two) c1.load(recs, j + 32, qxf, qyf, qzf, qx, qy, qz);
                        take_one(c0, j);
                        if (two) take_one(c1, j + 32);
                    }
                }
                __syncwarp();
                ncand = *ncand_s;
            } else {
                for (int rr = 0; rr < nruns; ++rr) {
                    const uint2 run = runs[rr];
                    for (unsigned t = run.x; t < run.y; t += 32) {
                        const unsigned j = t + lane;
                        bool hit = false;
                        double d2 = 0;
                        CandEval<WIDE> c;
                        if (j < run.y) {
                            c.load(recs, j, qxf, qyf, qzf, qx, qy, qz);
                            bool have = false;
                            if (in_radius(c, d2, have)) {
                                const int b = bstar == kBins ? 0 : bucket(c, d2, have);
                                if (b < bstar) accumulate(c.x(), c.y(), c.z(), c.idx());
                                else if (b == bstar) { hit = true; if (!have) d2 = c.exact(qx, qy, qz); }
                            }
                        }
                        if (bstar != kBins) {
                            const unsigned m = __ballot_sync(kFull, hit);
                            if (hit) {
                                const int slot = ncand + __popc(m & ((1u << lane) - 1u));
                                if (slot < kCand) { cand_d2[slot] = d2; cand_idx[slot] = c.idx(); cand_pos[slot] = (int)j; }
                            }
                            ncand += __popc(m);
                        }
                    }
                }
            }
            return ncand;
        };

        int bstar = kBins, need = 0;       // buckets < bstar are taken whole; `need` more come from bucket bstar
        // Moderately populated neighbourhoods at the full radius mostly hold <= k points inside the radius: take all
        // of them in ONE pass and count; only when more than k turn up is the work discarded and the selection run.
        bool settled = false;
        if (!trial && tot > np.max_nn && 3 * tot <= kSpecFactor3 * np.max_nn) {
            pass_b(kBins);
            if (warp_sum(cnt) <= np.max_nn) settled = true;
            else { sx = sy = sz = sxx = sxy = sxz = syy = syz = szz = 0; cnt = 0; tap_reset(); }
        }
        if (settled) break;
        if (tot > np.max_nn) {
            // ---- pass A: bucket histogram of the in-radius candidates
            for (int b = lane; b < kBins; b += 32) hist[b] = 0;
            __syncwarp();
            auto count_one = [&](const CandEval<WIDE>& c) {
                if constexpr (!WIDE) {
                    // float32 bucket coordinate (error < 1e-3 buckets): membership in the radius is decided exactly, the
                    // bucket itself may be off by one - the selection allows for that (the boundary is three buckets wide)
                    const float u = c.d2f * bsf;
                    if (u < (float)kBins - 2e-3f) { atomicAdd(&hist[(int)u], 1); return; }      // certainly inside the radius
                    if (u > (float)kBins + 2e-3f) return;                                        // certainly outside
                    const double d2 = c.exact(qx, qy, qz);
                    if (d2 < rq2) atomicAdd(&hist[min(kBins - 1, (int)(d2 * bin_scale))], 1);
                } else {
                    double d2 = 0;
                    bool have = false;
                    if (in_radius(c, d2, have)) atomicAdd(&hist[bucket(c, d2, have)], 1);
                }
            };
            for (int rr = 0; rr < nruns; ++rr) {
                const uint2 run = runs[rr];
                for (unsigned j = run.x + lane; j < run.y; j += 64) {      // two records in flight per lane
                    CandEval<WIDE> c0, c1;
                    const bool two = j + 32 < run.y;
                    c0.load(recs, j, qxf, qyf, qzf, qx, qy, qz);
                    if (two) c1.load(recs, j + 32, qxf, qyf, qzf, qx, qy, qz);
                    count_one(c0);
                    if (two) count_one(c1);
                }
            }
            __syncwarp();
            constexpr int per = kBins / 32;      // lane owns `per` consecutive buckets
            int local = 0;
#pragma unroll
            for (int b = 0; b < per; ++b) local += hist[lane * per + b];
            int inc = local;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, inc, o);
                if (lane >= o) inc += t;
            }
            const int n_in = __shfl_sync(kFull, inc, 31);
            dbg_nin = n_in;
            if (trial && n_in < np.max_nn) continue;          // the guess was too small: search the full radius
            if (n_in > np.max_nn) {
                int before = inc - local, myb = -1, myneed = 0;
                if (before < np.max_nn && inc >= np.max_nn) {
#pragma unroll
                    for (int b = 0; b < per; ++b) {
                        const int h = hist[lane * per + b];
                        if (myb < 0 && before + h >= np.max_nn) { myb = lane * per + b; myneed = np.max_nn - before; }
                        before += h;
                    }
                }
                const unsigned who = __ballot_sync(kFull, myb >= 0);
                const int src_lane = __ffs(who) - 1;
                bstar = __shfl_sync(kFull, myb, src_lane);
                need = __shfl_sync(kFull, myneed, src_lane);
            }
            __syncwarp();
        }

        const int ncand = pass_b(bstar);
        if (bstar != kBins) {
            __syncwarp();
            const int blo = WIDE ? bstar : bstar - 1, bhi = WIDE ? bstar : bstar + 1;
            need = np.max_nn - warp_sum(cnt);          // still missing once the buckets below the boundary are taken whole
            if (ncand <= kCand) {
                // exact rank inside the boundary bucket; the `need` smallest (d2, index) keys join the neighbourhood
                double td2 = 0;
                int tidx = 0;
                bool have_tau = false;
                for (int a = lane; a < ncand; a += 32) {
                    const double d2a = cand_d2[a];
                    const int ia = cand_idx[a];
                    int rank = 0;
                    for (int b = 0; b < ncand; ++b) rank += key_less(cand_d2[b], cand_idx[b], d2a, ia) ? 1 : 0;
                    if (rank < need) {
                        double x, y, z;
                        int idx;
                        load_rec(recs + cand_pos[a], x, y, z, idx);
                        accumulate(x, y, z, idx);
                    }
                    if (rank == need - 1) { td2 = d2a; tidx = ia; have_tau = true; }
                }
                const unsigned who = __ballot_sync(kFull, have_tau);
                const int src_lane = __ffs(who) - 1;
                tau_d2 = __shfl_sync(kFull, td2, src_lane);
                tau_idx = __shfl_sync(kFull, tidx, src_lane);
            } else {
                // pathological bucket (many duplicates): extract the `need` smallest keys of the bucket one at a time ...
                double last_d2 = -1.0;
                int last_idx = -1;
                for (int t = 0; t < need; ++t) {
                    double md2 = INFINITY;
                    int midx = 0x7fffffff;
                    for (int rr = 0; rr < nruns; ++rr) {
                        const uint2 run = runs[rr];
                        for (unsigned u = run.x + lane; u < run.y; u += 32) {
                            double x, y, z;
                            int idx;
                            load_rec(recs + u, x, y, z, idx);
                            const double d2 = sqdist(qx, qy, qz, x, y, z);
                            const int bb = min(kBins - 1, (int)(d2 * bin_scale));
                            if (d2 < rq2 && bb >= blo && bb <= bhi && key_less(last_d2, last_idx, d2, idx) &&
                                key_less(d2, idx, md2, midx)) { md2 = d2; midx = idx; }
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double od2 = __shfl_xor_sync(kFull, md2, o);
                        const int oidx = __shfl_xor_sync(kFull, midx, o);
                        if (key_less(od2, oidx, md2, midx)) { md2 = od2; midx = oidx; }
                    }
                    last_d2 = md2; last_idx = midx;
                }
                tau_d2 = last_d2; tau_idx = last_idx;
                // ... and redo the accumulation with the exact threshold
                sx = sy = sz = sxx = sxy = sxz = syy = syz = szz = 0;
                cnt = 0;
                tap_reset();
                for (int rr = 0; rr < nruns; ++rr) {
                    const uint2 run = runs[rr];
                    for (unsigned u = run.x + lane; u < run.y; u += 32) {
                        double x, y, z;
                        int idx;
                        load_rec(recs + u, x, y, z, idx);
                        const double d2 = sqdist(qx, qy, qz, x, y, z);
                        if (d2 < rq2 && !key_less(tau_d2, tau_idx, d2, idx)) accumulate(x, y, z, idx);
                    }
                }
            }
        } else if (trial) {
            tau_d2 = rq2;      // exactly k points inside the trial radius: all of them, none on the boundary
        }
        break;
    }

    if ((np.debug & 1) && lane == 0 && (p % 97) == 0)      // ARVC_DEBUG_NORMALS=1: sampled per-point search statistics
        printf("NRM p=%d total=%d ntry=%d total_try=%d rtry=%.3f used_trial=%d nruns=%d n_in=%d\n", p, total, ntry, total_try, rtry, (int)(runs == runs_try), nruns, dbg_nin);
    cnt = warp_sum(cnt);
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    sxx = warp_sum(sxx); sxy = warp_sum(sxy); sxz = warp_sum(sxz);
    syy = warp_sum(syy); syz = warp_sum(syz); szz = warp_sum(szz);
    if (mode == 0) {
        if (lane == 0) {
            double* m = s.moments + 10 * (size_t)p;
            m[0] = sx; m[1] = sy; m[2] = sz; m[3] = sxx; m[4] = sxy; m[5] = sxz; m[6] = syy; m[7] = syz; m[8] = szz; m[9] = (double)cnt;
            s.nn_count[p] = cnt;
        }
        return;
    }
    V3 nv{0.0, 0.0, 1.0};
    int redo = 0;
    // mode 1 only gets here: k_normals_eigen has already found the eigen-gap of this point below kIllGap (with cnt <= kCanon)
    redo = (cnt >= 3 && cnt <= kCanon) ? 1 : 0;
    if (!redo && lane == 0) {
        double gap = 1.0;
        if (cnt >= 3) {
            // centred second moments: better conditioned than raw cumulants, equal to them up to rounding
            const double inv = 1.0 / (double)cnt;
            const double mx = sx * inv, my = sy * inv, mz = sz * inv;
            nv = fast_eigen3x3(sxx * inv - mx * mx, sxy * inv - mx * my, sxz * inv - mx * mz, syy * inv - my * my, syz * inv - my * mz,
                               szz * inv - mz * mz, &gap);
        } else {
            nv = fast_eigen3x3(1, 0, 0, 1, 0, 1, &gap);   // Open3D: covariance = Identity when fewer than 3 neighbours
        }
    }
    if (redo) {
        // Ill-conditioned neighbourhood (e.g. collinear points of one scan ring): the eigenvector amplifies the
        // rounding of the covariance by 1/gap, so reproduce Open3D's arithmetic exactly: raw-coordinate cumulants,
        // summed sequentially in ascending (d2, index) order (the k-NN result order), no FMA contraction.
        double* key_d2 = reinterpret_cast<double*>(wmem);
        int* key_idx = reinterpret_cast<int*>(wmem + kCanon * 8);
        int* key_pos = reinterpret_cast<int*>(wmem + kCanon * 12);
        int* order = reinterpret_cast<int*>(wmem + kCanon * 16);
        __syncwarp();
        int m = 0;
        for (int rr = 0; rr < nruns; ++rr) {
            const uint2 run = runs[rr];
            for (unsigned t = run.x; t < run.y; t += 32) {
                const unsigned j = t + lane;
                double x, y, z, d2 = INFINITY;
                int idx = 0;
                if (j < run.y) {
                    load_rec(recs + j, x, y, z, idx);
                    d2 = sqdist(qx, qy, qz, x, y, z);
                }
                const bool sel = d2 < rq2 && !key_less(tau_d2, tau_idx, d2, idx);
                const unsigned msk = __ballot_sync(kFull, sel);
                if (sel) {
                    const int slot = m + __popc(msk & ((1u << lane) - 1u));
                    key_d2[slot] = d2; key_idx[slot] = idx; key_pos[slot] = (int)j;
                }
                m += __popc(msk);
            }
        }
        __syncwarp();
        for (int a = lane; a < m; a += 32) {      // rank sort: keys are distinct (index breaks ties)
            const double d2a = key_d2[a];
            const int ia = key_idx[a];
            int rank = 0;
            for (int b = 0; b < m; ++b) rank += key_less(key_d2[b], key_idx[b], d2a, ia) ? 1 : 0;
            order[rank] = key_pos[a];
        }
        __syncwarp();
        // the nine cumulants are independent sequential sums: lane k carries cumulant k (x, y, z, xx, xy, xz, yy, yz, zz;
        // x * 1.0 is exact), every lane reads the same record (broadcast), so each sum keeps the oracle's order
        double acc = 0.0;
        {
            const int fa = lane < 3 ? lane : (lane < 6 ? 0 : (lane < 8 ? 1 : 2));          // first factor: 0 = x, 1 = y, 2 = z
            const int fb = lane < 3 ? 3 : (lane < 6 ? lane - 3 : (lane < 8 ? lane - 5 : 2));   // second factor, 3 = the constant 1
            for (int t = 0; t < m; ++t) {
                double x, y, z;
                int idx;
                load_rec(recs + order[t], x, y, z, idx);
                const double a = fa == 0 ? x : (fa == 1 ? y : z);
                const double b = fb == 0 ? x : (fb == 1 ? y : (fb == 2 ? z : 1.0));
                acc = __dadd_rn(acc, __dmul_rn(a, b));
            }
        }
        double c0 = __shfl_sync(kFull, acc, 0), c1 = __shfl_sync(kFull, acc, 1), c2 = __shfl_sync(kFull, acc, 2),
               c3 = __shfl_sync(kFull, acc, 3), c4 = __shfl_sync(kFull, acc, 4), c5 = __shfl_sync(kFull, acc, 5),
               c6 = __shfl_sync(kFull, acc, 6), c7 = __shfl_sync(kFull, acc, 7), c8 = __shfl_sync(kFull, acc, 8);
        if (lane == 0) {
            const double dn = (double)m;
            c0 = __ddiv_rn(c0, dn); c1 = __ddiv_rn(c1, dn); c2 = __ddiv_rn(c2, dn); c3 = __ddiv_rn(c3, dn); c4 = __ddiv_rn(c4, dn);
            c5 = __ddiv_rn(c5, dn); c6 = __ddiv_rn(c6, dn); c7 = __ddiv_rn(c7, dn); c8 = __ddiv_rn(c8, dn);
            double gap;
            nv = fast_eigen3x3(__dsub_rn(c3, __dmul_rn(c0, c0)), __dsub_rn(c4, __dmul_rn(c0, c1)), __dsub_rn(c5, __dmul_rn(c0, c2)),
                               __dsub_rn(c6, __dmul_rn(c1, c1)), __dsub_rn(c7, __dmul_rn(c1, c2)), __dsub_rn(c8, __dmul_rn(c2, c2)), &gap);
        }
    }
    if (lane == 0) {
        if (sqrt(nv.x * nv.x + nv.y * nv.y + nv.z * nv.z) == 0.0) nv = {0.0, 0.0, 1.0};
        reinterpret_cast<double4*>(s.normals)[p] = make_double4(nv.x, nv.y, nv.z, 0.0);
        s.nn_count[p] = cnt;
    }
}

template <bool WIDE, bool TAP>
__global__ void __launch_bounds__(kNrmWarps * 32, 3) k_normals(const ScanDev* __restrict__ scans, NormalParams np, int mode) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const ScanDev& s = scans[blockIdx.y];
    if ((s.wide != 0) != WIDE) return;
    const int n = s.counts[CNT_NPTS];
    const int w = threadIdx.x >> 5;
    if (mode == 0) {
        const int p = blockIdx.x * kNrmWarps + w;
        if (p < n) normals_point<WIDE, TAP>(s, np, 0, p, s_raw);
    } else if (mode == 2) {
        // the points the block kernel (normals_blk.cu) could not serve from its shared tile: same search, per point
        const int nfb = s.counts[CNT_NFB];
        for (int q = blockIdx.x * kNrmWarps + w; q < nfb; q += gridDim.x * kNrmWarps) {
            const int p = s.fb_list[q];
            if (p < n) normals_point<WIDE, TAP>(s, np, 0, p, s_raw);
            __syncwarp();
        }
    } else {
        // the redo list holds a few per cent of the points: a small grid strides over it (a full-size grid of blocks that
        // exit at once costs more than the work itself)
        const int nredo = s.counts[CNT_NREDO];
        for (int q = blockIdx.x * kNrmWarps + w; q < nredo; q += gridDim.x * kNrmWarps) {
            const int p = s.redo_list[q];
            if (p < n) normals_point<WIDE, TAP>(s, np, 1, p, s_raw);
            __syncwarp();
        }
    }
}

// One thread per point: covariance from the centred moment sums, analytic eigen-solve, normal.  Points whose two smallest
// eigenvalues nearly coincide are appended to the scan's redo list for the canonical (oracle-order) re-summation.
__global__ void __launch_bounds__(128) k_normals_eigen(const ScanDev* __restrict__ scans) {
    const ScanDev& s = scans[blockIdx.y];
    const int n = s.counts[CNT_NPTS];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const double* m = s.moments + 10 * (size_t)p;
    const int cnt = (int)m[9];
    V3 nv;
    double gap = 1.0;
    if (cnt >= 3) {
        // centred second moments: better conditioned than raw cumulants, equal to them up to rounding
        const double inv = 1.0 / (double)cnt;
        const double mx = m[0] * inv, my = m[1] * inv, mz = m[2] * inv;
        nv = fast_eigen3x3(m[3] * inv - mx * mx, m[4] * inv - mx * my, m[5] * inv - mx * mz, m[6] * inv - my * my, m[7] * inv - my * mz,
                           m[8] * inv - mz * mz, &gap);
        if (gap < kIllGap && cnt <= kCanon) s.redo_list[atomicAdd(&s.counts[CNT_NREDO], 1)] = p;
    } else {
        nv = fast_eigen3x3(1, 0, 0, 1, 0, 1, &gap);   // Open3D: covariance = Identity when fewer than 3 neighbours
    }
    if (sqrt(nv.x * nv.x + nv.y * nv.y + nv.z * nv.z) == 0.0) nv = {0.0, 0.0, 1.0};
    reinterpret_cast<double4*>(s.normals)[p] = make_double4(nv.x, nv.y, nv.z, 0.0);
}

void run_normals(Launcher& L, const ScanDev* d_scans, int n_scans, int cap_max, const NormalParams& np_in, bool any_wide, bool any_narrow, bool tap) {
    if (n_scans == 0 || cap_max == 0) return;
    NormalParams np = np_in;
    np.r2 = np.radius * np.radius;
    np.bin_scale = (double)kBins / np.r2;
    np.r2_lo = (float)(np.r2 * (1.0 - 2e-6));
    np.r2_hi = (float)(np.r2 * (1.0 + 2e-6));
    np.bin_scale_f = (float)np.bin_scale;
    np.debug = getenv("ARVC_DEBUG_NORMALS") ? atoi(getenv("ARVC_DEBUG_NORMALS")) : 0;
    np.crowded_ratio = getenv("ARVC_NB_CROWDED") ? (float)atof(getenv("ARVC_NB_CROWDED")) : 1.25f;      // tuning knob; results do not depend on it
    const dim3 grid((cap_max + kNrmWarps - 1) / kNrmWarps, n_scans), block(kNrmWarps * 32);
    const size_t smem = (size_t)kNrmWarps * kWarpSmem, smem_sel = (size_t)kNrmWarps * kWarpSmemSel;
    static unsigned long long attr_set = 0;      // per device: the attribute belongs to the device's copy of the kernel
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((attr_set >> (dev & 63)) & 1ull)) {
        cudaFuncSetAttribute(k_normals<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_normals<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_normals<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_normals<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set |= 1ull << (dev & 63);
    }
    auto point_kernel = [&](bool wide) { return wide ? (tap ? k_normals<true, true> : k_normals<true, false>) : (tap ? k_normals<false, true> : k_normals<false, false>); };
    // float32 records (the PCD payload, voxel off): block-cooperative kernel + per-point kernel for what it hands back;
    // float64 records (voxel means, float64 uploads): per-point kernel.  ARVC_NORMALS_IMPL=point forces the latter.
    static const bool per_point_only = getenv("ARVC_NORMALS_IMPL") && std::string(getenv("ARVC_NORMALS_IMPL")) == "point";
    if (any_narrow && per_point_only) L.launch_smem("normals", point_kernel(false), grid, block, smem_sel, d_scans, np, 0);
    if (any_narrow && !per_point_only) {
        launch_normals_blk(L, d_scans, n_scans, cap_max, np, tap);
        L.launch_smem("normals_fallback", point_kernel(false), dim3(min(grid.x, 192u), n_scans), block, smem_sel, d_scans, np, 2);
    }
    if (any_wide) L.launch_smem("normals", point_kernel(true), grid, block, smem_sel, d_scans, np, 0);
    L.launch("normals_eigen", k_normals_eigen, dim3((cap_max + 127) / 128, n_scans), dim3(128), d_scans);
    // ill-conditioned points (a few per cent)
    const dim3 grid_redo(min(grid.x, 96u), n_scans);
    if (any_narrow) L.launch_smem("normals_redo", point_kernel(false), grid_redo, block, smem, d_scans, np, 1);
    if (any_wide) L.launch_smem("normals_redo", point_kernel(true), grid_redo, block, smem, d_scans, np, 1);
}

}  // namespace arvc
