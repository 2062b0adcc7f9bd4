// ICP iteration, fully device-resident.  One pass over a batch of pairs = four kernels:
//   k_icp_select  (passes > 0) nearest-neighbour certificates: which source points need a new search;
//   k_icp_search  exact nearest neighbour on the target's multi-resolution hash grid for those points;
//   k_icp_accum   residual / Jacobian (point-to-plane) or Umeyama moments (point-to-point) of the matches, fixed-order
//                 warp reduction into one row of partial sums per warp;
//   k_icp_finish  per pair: reduction of the rows, 6x6 LDLT / 3x3 SVD solve, update of the cumulative transformation
//                 and Open3D's convergence test.
// The host enqueues ONE CUDA graph per batch: the first kIcpUnrolled passes as plain kernel nodes, then a conditional
// WHILE node whose body is one pass and whose condition - "some pair of the batch has not converged" - is set on the
// device (k_icp_cond), so the iteration ends with the last convergence and without any host round trip.
// (ARVC_ICP_LOOP=unrolled: max_iter + 1 passes enqueued unconditionally; finished pairs turn into no-ops.)
//
// Reference semantics restated: keyframemanager/keyframe.py:246-252 -> Open3D RegistrationICP,
// GetRegistrationResultAndCorrespondences, TransformationEstimationPointToPlane / PointToPoint.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "engine.cuh"

namespace arvc {

namespace {

struct Best { double d2; int idx; int pos; };

// ---------------------------------------------------------------------------------------------------
// Exact nearest neighbour on the multi-resolution hash grid, kG lanes cooperating on one query.
//
// Work items are grid cells (level, cx, cy, cz, [start, end)) on a small per-group stack in shared memory:
//   * the cells of the search ball at the start level are pushed home-cell-first;
//   * a popped cell is dropped when its box is farther than the best so far, scanned cooperatively
//     (coalesced 16-byte records, float32 screening, float64 evaluation only for candidates that can win or tie)
//     when it is small or at the finest level, and otherwise expanded into its eight children - lane r looks
//     up the r-th nearest octant - which go back on the stack nearest-on-top (branch and bound);
//   * when the level's ball is exhausted and the best is not yet certified, the search moves one level up.
// The result is the exact minimiser of (d2, cloud index): d2 evaluated exactly as the oracle evaluates it.
// ---------------------------------------------------------------------------------------------------
constexpr int kG = 8;            // lanes per query (16: unions outgrow the caps, 15x slower)
constexpr unsigned kGMask = kG == 32 ? 0xffffffffu : ((1u << kG) - 1u);
#ifndef ARVC_BIGRUN
#define ARVC_BIGRUN 160
#endif
constexpr int kBigRun = ARVC_BIGRUN;     // longer runs are expanded into their children instead of scanned
constexpr int kStack = 48;       // stack entries per group
constexpr int kGroupsPerBlock = kIcpBlock / kG;
constexpr int kUnionLevels = 5;  // finest levels the shared-candidate (union) phase may use (coarse ones only where sparse)
constexpr int kUnionCandCap = 1024;   // most records a union at level >= 2 may hold
// resident blocks per SM the search / far-query kernels are compiled for (register budgets: 8 -> 121, 10 -> 96, 16 -> 64
// registers per thread); the _DENSE builds serve batches of kDenseBatch pairs or more
#ifndef ARVC_SEARCH_OCC
#define ARVC_SEARCH_OCC 8
#endif
#ifndef ARVC_SEARCH_OCC_DENSE
#define ARVC_SEARCH_OCC_DENSE 10
#endif
#ifndef ARVC_DENSE_BATCH
#define ARVC_DENSE_BATCH 24
#endif
constexpr int kDenseBatch = ARVC_DENSE_BATCH;
#ifndef ARVC_FARSH0
#define ARVC_FARSH0 99     // far grid: no shrinking with the passes (measured: tails dominate)
#define ARVC_FARSH1 99
#endif
#ifndef ARVC_SCANCAP
#define ARVC_SCANCAP 1024
#endif
constexpr int kUnionScanCap = ARVC_SCANCAP;    // larger unions are split per query (lanes share the records of one query)
constexpr int kUnionMax = 64;    // most cells a group's union may span (<= 2 * kStack runs fit the stack memory)
#ifndef ARVC_STAGE
#define ARVC_STAGE 64
#endif
constexpr int kStage = ARVC_STAGE;        // records staged in shared memory per group and round
#ifndef ARVC_STAGE_BULK
#define ARVC_STAGE_BULK 0
#endif
#ifndef ARVC_FAR_OCC
#define ARVC_FAR_OCC 1
#endif
#ifndef ARVC_FAR_OCC_DENSE
#define ARVC_FAR_OCC_DENSE 16
#endif
#ifndef ARVC_FAR_G
#define ARVC_FAR_G 8      // lanes per query in the far kernel
#endif
#ifndef ARVC_FARBLOCKS
#define ARVC_FARBLOCKS 296
#endif
constexpr int kFarBlocks = ARVC_FARBLOCKS;   // blocks per pair of the far-query kernel (2 x 148 SMs)
constexpr int kSpreadBlocks = 3552;   // three waves of search blocks (148 SMs x 8 resident): below that, queries are spread thinner

__device__ __forceinline__ double box_d2(const GridSpec& g, int cx, int cy, int cz, double cl, double sx, double sy, double sz) {
    const double bx0 = g.ox + cx * cl, by0 = g.oy + cy * cl, bz0 = g.oz + cz * cl;
    const double ddx = fmax(0.0, fmax(bx0 - sx, sx - (bx0 + cl)));
    const double ddy = fmax(0.0, fmax(by0 - sy, sy - (by0 + cl)));
    const double ddz = fmax(0.0, fmax(bz0 - sz, sz - (bz0 + cl)));
    return ddx * ddx + ddy * ddy + ddz * ddz;
}

// float32 screening threshold for a best-so-far d2 (see DESIGN.md "float32 screening"): narrow target records are
// exact float32 values, so |d2f - d2| <= 2e-6 d2 + 3.5 sqrt(d2) e + 3 e^2 with e the rounding of the query coordinates.
// The threshold is re-evaluated under divergence every time a lane's best improves, so it is kept to one conversion and
// one FMA: sqrt(b) <= 5 b + 0.05 (AM-GM around 0.1 m, the typical match distance) turns the bound into k1 * b + k0.
// The slack this adds is ~1e-7 m^2 - it only lets a few more candidates through to the exact float64 comparison.
struct ScreenThr {
    float k0, k1;
    __device__ __forceinline__ explicit ScreenThr(float e)
        : k0((0.18f * e + 3.1f * e * e) * (1.0f + 1e-6f)), k1((1.0f + 4e-6f + 18.0f * e) * (1.0f + 1e-6f)) {}
    __device__ __forceinline__ float operator()(double best_d2) const { return fmaf(__double2float_ru(best_d2), k1, k0); }
};

template <bool TW, int G, int STK>
__device__ __forceinline__ void group_search(const ScanDev& tgt, double sx, double sy, double sz, double in_d2, int in_idx, int level,
                                             Best& lb, uint4* __restrict__ stk, float* __restrict__ stk_lb, int gl,
                                             unsigned gmask) {
    typedef typename RecT<TW>::type Rec;
    const Rec* __restrict__ recs = reinterpret_cast<const Rec*>(tgt.recs);
    const HashEntry* __restrict__ tab = tgt.table;
    const unsigned tmask = tgt.table_mask;
    const GridSpec g = tgt.grid;
    const int gshift = (lane_id() / G) * G;
    constexpr unsigned kMaskG = G >= 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u);
    const float sxf = (float)sx, syf = (float)sy, szf = (float)sz;
    const float e = (float)(fmax(fabs(sx), fmax(fabs(sy), fabs(sz))) * 6.0e-8 + 1e-30);
    double gbest = in_d2;                       // group-wide best d2 (every lane holds the same value)
    const ScreenThr screen_thr(e);
    float thr = screen_thr(gbest);
    lb.d2 = in_d2; lb.idx = in_idx; lb.pos = -1;

    for (int l = level; l <= g.top_level; ++l) {
        const double cl = g.c0 * (double)(1 << l);
        const double br = sqrt(gbest) * (1.0 + 1e-9) + 1e-12;
        const double r = fmin(br, cl);
        const int x0 = cell_coord(sx - r, g.ox, g.inv_c0) >> l, x1 = cell_coord(sx + r, g.ox, g.inv_c0) >> l;
        const int y0 = cell_coord(sy - r, g.oy, g.inv_c0) >> l, y1 = cell_coord(sy + r, g.oy, g.inv_c0) >> l;
        const int z0 = cell_coord(sz - r, g.oz, g.inv_c0) >> l, z1 = cell_coord(sz + r, g.oz, g.inv_c0) >> l;
        const int hx = min(max(cell_coord(sx, g.ox, g.inv_c0) >> l, x0), x1), hy = min(max(cell_coord(sy, g.oy, g.inv_c0) >> l, y0), y1),
                  hz = min(max(cell_coord(sz, g.oz, g.inv_c0) >> l, z0), z1);
        const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, ncell = nx * ny * (z1 - z0 + 1);     // <= 27
        const int home = (hx - x0) + nx * ((hy - y0) + ny * (hz - z0));
        const CellDecoder dec(nx, ny);
        int top = 0;
        // push the ball's cells, later rounds first, so that order position 0 (the home cell) ends on top
        for (int base = ((ncell - 1) / G) * G; base >= 0; base -= G) {
            const int t = base + gl;
            bool valid = false;
            unsigned st = 0, en = 0;
            int cx = 0, cy = 0, cz = 0;
            float lbf = 0.f;
            if (t < ncell) {
                const int c = t == 0 ? home : (t <= home ? t - 1 : t);
                dec(c, cx, cy, cz);
                cx += x0; cy += y0; cz += z0;
                const double bd = box_d2(g, cx, cy, cz, cl, sx, sy, sz);
                if (bd <= gbest * (1.0 + 1e-9) + 1e-12) {
                    valid = grid_lookup(tab, tmask, l, morton3(cx, cy, cz), st, en);
                    lbf = __double2float_rd(bd);
                }
            }
            const unsigned vm = (__ballot_sync(gmask, valid) >> gshift) & kMaskG;
            if (valid) {
                const int pos = top + __popc(vm & ~((2u << gl) - 1u));
                stk[pos] = make_uint4(st, en, (unsigned)cx | ((unsigned)cy << 10) | ((unsigned)cz << 20), (unsigned)l);
                stk_lb[pos] = lbf;
            }
            top += __popc(vm);
        }
        __syncwarp(gmask);
        while (top > 0) {
            --top;
            const uint4 en4 = stk[top];
            const float lbf = stk_lb[top];
            __syncwarp(gmask);
            if ((double)lbf > gbest * (1.0 + 1e-9) + 1e-12) continue;
            const int el = (int)en4.w;
            const unsigned cnt = en4.y - en4.x;
            if (el == 0 || cnt <= (unsigned)kBigRun || top + 8 > STK) {
                if constexpr (!TW) {
                    const float4* __restrict__ r4 = reinterpret_cast<const float4*>(recs);
                    auto eval = [&](const float4& v, unsigned p) {
                        const float dx = sxf - v.x, dy = syf - v.y, dz = szf - v.z;
                        const float d2f = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        if (d2f <= thr) {
                            const double d2 = sqdist(sx, sy, sz, (double)v.x, (double)v.y, (double)v.z);
                            const int idx = __float_as_int(v.w);
                            if (d2 < lb.d2 || (d2 == lb.d2 && idx < lb.idx)) { lb.d2 = d2; lb.idx = idx; lb.pos = (int)p; thr = screen_thr(d2); }
                        }
                    };
                    unsigned p0 = en4.x + gl;
                    for (; p0 + 3 * G < en4.y; p0 += 4 * G) {      // four loads in flight per lane, no bounds tests in the bulk
                        const float4 a = __ldg(r4 + p0), b = __ldg(r4 + p0 + G), c = __ldg(r4 + p0 + 2 * G), d = __ldg(r4 + p0 + 3 * G);
                        eval(a, p0); eval(b, p0 + G); eval(c, p0 + 2 * G); eval(d, p0 + 3 * G);
                    }
                    for (; p0 < en4.y; p0 += G) eval(__ldg(r4 + p0), p0);
                }
                for (unsigned p = en4.x + gl; TW && p < en4.y; p += G) {
                    if constexpr (!TW) {
                    } else {
                        double x, y, z;
                        int idx;
                        load_rec(recs + p, x, y, z, idx);
                        const double d2 = sqdist(sx, sy, sz, x, y, z);
                        if (d2 < lb.d2 || (d2 == lb.d2 && idx < lb.idx)) { lb.d2 = d2; lb.idx = idx; lb.pos = (int)p; }
                    }
                }
                // share the bound: group-wide minimum of the lanes' best distances
                double m = lb.d2;
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(gmask, m, o, G));
                if (m < gbest) { gbest = m; thr = fminf(thr, screen_thr(m)); }
            } else {
                // expand into the eight children, lane r <-> r-th nearest octant
                const int cx = (int)(en4.z & 1023u), cy = (int)((en4.z >> 10) & 1023u), cz = (int)(en4.z >> 20);
                const double pcl = g.c0 * (double)(1 << el), h = 0.5 * pcl;
                const unsigned near = ((sx >= g.ox + cx * pcl + h) ? 4u : 0u) | ((sy >= g.oy + cy * pcl + h) ? 2u : 0u) |
                                      ((sz >= g.oz + cz * pcl + h) ? 1u : 0u);
                // rounds of G octants, farthest first, so that the nearest octant ends on top of the stack
#pragma unroll
                for (int rnd = (8 + G - 1) / G - 1; rnd >= 0; --rnd) {
                    const int r = rnd * G + gl;
                    const unsigned child = near ^ ((0x76534210u >> (4 * (r & 7))) & 7u);      // 0,1,2,4,3,5,6,7 axis flips
                    const int ccx = 2 * cx + (int)((child >> 2) & 1u), ccy = 2 * cy + (int)((child >> 1) & 1u), ccz = 2 * cz + (int)(child & 1u);
                    const double bd = box_d2(g, ccx, ccy, ccz, h, sx, sy, sz);
                    bool valid = false;
                    unsigned st = 0, en = 0;
                    if (r < 8 && bd <= gbest * (1.0 + 1e-9) + 1e-12) valid = grid_lookup(tab, tmask, el - 1, morton3(ccx, ccy, ccz), st, en);
                    const unsigned vm = (__ballot_sync(gmask, valid) >> gshift) & kMaskG;
                    if (valid) {
                        const int pos = top + __popc(vm & ~((2u << gl) - 1u));
                        stk[pos] = make_uint4(st, en, (unsigned)ccx | ((unsigned)ccy << 10) | ((unsigned)ccz << 20), (unsigned)(el - 1));
                        stk_lb[pos] = __double2float_rd(bd);
                    }
                    top += __popc(vm);
                }
                __syncwarp(gmask);
            }
        }
        // every point within min(best radius, cl) has been seen: exact as soon as the best lies within cl
        if (sqrt(gbest) * (1.0 + 1e-9) + 1e-12 <= cl) break;
    }
    // lexicographic (d2, index) minimum over the group; lanes that found nothing carry the incoming bound with pos = -1
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        const double od2 = __shfl_xor_sync(gmask, lb.d2, o, G);
        const int oidx = __shfl_xor_sync(gmask, lb.idx, o, G);
        const int opos = __shfl_xor_sync(gmask, lb.pos, o, G);
        if (od2 < lb.d2 || (od2 == lb.d2 && (oidx < lb.idx || (oidx == lb.idx && opos > lb.pos)))) { lb.d2 = od2; lb.idx = oidx; lb.pos = opos; }
    }
}

// ---- small dense solvers (one thread) ---------------------------------------------------------------
__device__ void ldlt_solve6(double* A /*6x6 full, destroyed*/, const double* b, double* x) {
    const int n = 6;
    int piv[6];
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i) if (fabs(A[i * n + i]) > best) { best = fabs(A[i * n + i]); p = i; }
        piv[k] = p;
        if (p != k) {
            for (int j = 0; j < n; ++j) { const double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
            for (int i = 0; i < n; ++i) { const double t = A[i * n + k]; A[i * n + k] = A[i * n + p]; A[i * n + p] = t; }
        }
        const double d = A[k * n + k];
        if (d != 0.0) {
            for (int i = k + 1; i < n; ++i) A[i * n + k] /= d;
            for (int i = k + 1; i < n; ++i)
                for (int j = k + 1; j <= i; ++j) {
                    A[i * n + j] -= A[i * n + k] * d * A[j * n + k];
                    A[j * n + i] = A[i * n + j];
                }
        }
    }
    double y[6];
    for (int i = 0; i < n; ++i) y[i] = b[i];
    for (int k = 0; k < n; ++k) { const double t = y[k]; y[k] = y[piv[k]]; y[piv[k]] = t; }
    for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) y[i] -= A[i * n + j] * y[j];
    for (int i = 0; i < n; ++i) y[i] = fabs(A[i * n + i]) > 2.2250738585072014e-308 ? y[i] / A[i * n + i] : 0.0;
    for (int i = n - 1; i >= 0; --i) for (int j = i + 1; j < n; ++j) y[i] -= A[j * n + i] * y[j];
    for (int k = n - 1; k >= 0; --k) { const double t = y[k]; y[k] = y[piv[k]]; y[piv[k]] = t; }
    for (int i = 0; i < n; ++i) x[i] = y[i];
}

// R = Rz(x2) Ry(x1) Rx(x0), t = (x3,x4,x5)  (Open3D TransformVector6dToMatrix4d)
__device__ void vec6_to_mat4(const double* v, double* T) {
    double sa, ca, sb, cb, sg, cg;
    sincos(v[0], &sa, &ca); sincos(v[1], &sb, &cb); sincos(v[2], &sg, &cg);
    T[0] = cg * cb; T[1] = cg * sb * sa - sg * ca; T[2] = cg * sb * ca + sg * sa; T[3] = v[3];
    T[4] = sg * cb; T[5] = sg * sb * sa + cg * ca; T[6] = sg * sb * ca - cg * sa; T[7] = v[4];
    T[8] = -sb;     T[9] = cb * sa;                T[10] = cb * ca;               T[11] = v[5];
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

__device__ double det3(const double* M) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// one-sided Jacobi SVD of a 3x3: A = U diag(s) V^T, s descending
__device__ void svd3(const double* Ain, double* U, double* s, double* V) {
    double A[9];
    for (int i = 0; i < 9; ++i) { A[i] = Ain[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 3; ++i) { alpha += A[3 * i + p] * A[3 * i + p]; beta += A[3 * i + q] * A[3 * i + q]; gamma += A[3 * i + p] * A[3 * i + q]; }
                if (gamma == 0.0 || fabs(gamma) <= 1e-17 * sqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int i = 0; i < 3; ++i) {
                    const double ap = A[3 * i + p], aq = A[3 * i + q];
                    A[3 * i + p] = c * ap - sn * aq; A[3 * i + q] = sn * ap + c * aq;
                    const double vp = V[3 * i + p], vq = V[3 * i + q];
                    V[3 * i + p] = c * vp - sn * vq; V[3 * i + q] = sn * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    double nrm[3];
    for (int j = 0; j < 3; ++j) nrm[j] = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
    int ord[3] = {0, 1, 2};
    for (int a = 0; a < 2; ++a)
        for (int c = a + 1; c < 3; ++c)
            if (nrm[ord[c]] > nrm[ord[a]]) { const int t = ord[a]; ord[a] = ord[c]; ord[c] = t; }
    double Vs[9];
    for (int j = 0; j < 3; ++j) {
        const int o = ord[j];
        s[j] = nrm[o];
        for (int i = 0; i < 3; ++i) { Vs[3 * i + j] = V[3 * i + o]; U[3 * i + j] = nrm[o] > 0 ? A[3 * i + o] / nrm[o] : 0.0; }
    }
    for (int i = 0; i < 9; ++i) V[i] = Vs[i];
    const double tiny = s[0] * 1e-14;
    if (s[0] <= 0) { for (int i = 0; i < 9; ++i) U[i] = (i % 4 == 0) ? 1.0 : 0.0; return; }
    if (s[1] <= tiny) {
        const double u0x = U[0], u0y = U[3], u0z = U[6];
        const double ax = fabs(u0x) < 0.9 ? 1.0 : 0.0, ay = fabs(u0x) < 0.9 ? 0.0 : 1.0;
        double cx = u0y * 0.0 - u0z * ay, cy = u0z * ax - u0x * 0.0, cz = u0x * ay - u0y * ax;
        const double nn = sqrt(cx * cx + cy * cy + cz * cz);
        U[1] = cx / nn; U[4] = cy / nn; U[7] = cz / nn;
    }
    if (s[2] <= tiny) {
        U[2] = U[3] * U[7] - U[6] * U[4];
        U[5] = U[6] * U[1] - U[0] * U[7];
        U[8] = U[0] * U[4] - U[3] * U[1];
    }
}

__device__ __forceinline__ void mat4_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j];
            C[4 * i + j] = s;
        }
}

// sums layout: [0..20] JtJ upper triangle (row-major), [21..26] Jtr, [27] sum d2, [28] K          (point-to-plane)
//              [0..2] sum s, [3..5] sum t, [6..14] sum t s^T (row-major), [27] sum d2, [28] K    (point-to-point)
constexpr int kNS = 29;

__device__ void solve_update(int method, const double* S, double* upd) {
    for (int i = 0; i < 16; ++i) upd[i] = (i % 5 == 0) ? 1.0 : 0.0;
    const double K = S[28];
    if (K <= 0) return;
    if (method == 1) {
        double A[36], nb[6], x[6];
        int t = 0;
        for (int a = 0; a < 6; ++a)
            for (int c = a; c < 6; ++c) { A[6 * a + c] = S[t]; A[6 * c + a] = S[t]; ++t; }
        for (int a = 0; a < 6; ++a) nb[a] = -S[21 + a];
        ldlt_solve6(A, nb, x);
        vec6_to_mat4(x, upd);
    } else {
        const double inv = 1.0 / K;
        double ms[3], mt[3], Sg[9];
        for (int d = 0; d < 3; ++d) { ms[d] = S[d] * inv; mt[d] = S[3 + d] * inv; }
        for (int a = 0; a < 3; ++a)
            for (int c = 0; c < 3; ++c) Sg[3 * a + c] = S[6 + 3 * a + c] * inv - mt[a] * ms[c];
        double U[9], sv[3], V[9], R[9];
        svd3(Sg, U, sv, V);
        const double sgn = (det3(U) * det3(V) < 0) ? -1.0 : 1.0;
        for (int a = 0; a < 3; ++a)
            for (int c = 0; c < 3; ++c) R[3 * a + c] = U[3 * a + 0] * V[3 * c + 0] + U[3 * a + 1] * V[3 * c + 1] + sgn * U[3 * a + 2] * V[3 * c + 2];
        for (int a = 0; a < 3; ++a) {
            for (int c = 0; c < 3; ++c) upd[4 * a + c] = R[3 * a + c];
            upd[4 * a + 3] = mt[a] - (R[3 * a + 0] * ms[0] + R[3 * a + 1] * ms[1] + R[3 * a + 2] * ms[2]);
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// One ICP pass = three kernels + the per-pair finish kernel, all over the whole batch:
//   k_icp_select  (passes > 0) per source point: distance to the previous match and certificate test; points
//                 whose nearest neighbour is not certified are appended to the pair's work list;
//   k_icp_search  exact nearest-neighbour search for the listed points only (every point in pass 0), 8 consecutive
//                 list entries per group, so the cost follows the number of uncertified points;
//   k_icp_accum   per source point in Morton order: residual / Jacobian or Umeyama moments of its match, fixed-order
//                 warp reduction into one row of partial sums per warp (bit-reproducible run to run).
// ---------------------------------------------------------------------------------------------------
template <bool SW, bool TW>
__global__ void __launch_bounds__(256) k_icp_select(const BatchDesc* __restrict__ bd) {
    typedef typename RecT<SW>::type SRec;
    typedef typename RecT<TW>::type TRec;
    if ((int)blockIdx.y >= bd->n_pairs) return;
    const PairDev& pr = bd->pairs[blockIdx.y];
    const double ip_max_d2 = bd->ip.max_d2;
    const ScanDev& src = *pr.src;
    const ScanDev& tgt = *pr.tgt;
    if ((src.wide != 0) != SW || (tgt.wide != 0) != TW) return;
    PairState* st = pr.state;
    if (st->done) return;
    const int n = src.counts[CNT_NPTS];
    __shared__ int s_cnt[8], s_base;
#pragma unroll 1
    for (int chunk = blockIdx.x; chunk * 256 < n; chunk += gridDim.x) {
    const int i = chunk * 256 + threadIdx.x;
    bool need = false;
    if (i < n) {
        double px, py, pz;
        int sidx;
        load_rec(reinterpret_cast<const SRec*>(src.recs) + i, px, py, pz, sidx);
        const double* T = st->T;
        const double sx = T[0] * px + T[1] * py + T[2] * pz + T[3];
        const double sy = T[4] * px + T[5] * py + T[6] * pz + T[7];
        const double sz = T[8] * px + T[9] * py + T[10] * pz + T[11];
        need = (sx == sx && sy == sy && sz == sz) && ip_max_d2 > 0;
        const int pv = pr.prev[i];
        const int c = pr.cert_pass[i];
        if (need && pv >= 0 && c != 255) {
            // Certificate from pass c: every other target point was at least lb2 away from this source point's position
            // then.  It has moved by delta since, so every other point is still at least lb2 - delta away: if the previous
            // match is strictly closer (and inside the cut-off), it is the unique nearest neighbour - no search.
            double x, y, z;
            int idx;
            load_rec(reinterpret_cast<const TRec*>(tgt.recs) + pv, x, y, z, idx);
            const double d2 = sqdist(sx, sy, sz, x, y, z);
            const double* Tc = st->Thist + 12 * c;
            const double ex = sx - (Tc[0] * px + Tc[1] * py + Tc[2] * pz + Tc[3]);
            const double ey = sy - (Tc[4] * px + Tc[5] * py + Tc[6] * pz + Tc[7]);
            const double ez = sz - (Tc[8] * px + Tc[9] * py + Tc[10] * pz + Tc[11]);
            const double delta = sqrt(ex * ex + ey * ey + ez * ez);
            const double br = sqrt(d2) * (1.0 + 1e-9) + 1e-12;
            if (d2 < ip_max_d2 && (br + delta) * (1.0 + 1e-7) + 1e-9 < (double)pr.lb2[i]) need = false;
        }
    }
    // block-ordered append, padded to a multiple of kG with -1: the kG entries a search group serves then always come
    // from one block of 256 consecutive (Morton-sorted) source points, which keeps the groups spatially compact
    const int lane = lane_id(), w = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(kFull, need);
    if (lane == 0) s_cnt[w] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { before += k < w ? s_cnt[k] : 0; total += s_cnt[k]; }
    const int padded = (total + kG - 1) / kG * kG;
    if (threadIdx.x == 0) s_base = total ? atomicAdd(&st->nlist, padded) : 0;
    __syncthreads();
    if (need) pr.list[s_base + before + __popc(m & ((1u << lane) - 1u))] = i;
    if ((int)threadIdx.x < padded - total) pr.list[s_base + total + threadIdx.x] = -1;
    __syncthreads();
    }   // chunk loop
}

template <bool SW, bool TW, int OCC>
__global__ void __launch_bounds__(kIcpBlock, OCC) k_icp_search(const BatchDesc* __restrict__ bd) {
    typedef typename RecT<SW>::type SRec;
    typedef typename RecT<TW>::type TRec;
    __shared__ float4 s_stage[kGroupsPerBlock][kStage];
    __shared__ uint4 s_stk[kGroupsPerBlock][kStack];      // the union phase's run table
    __shared__ int s_roff[kGroupsPerBlock][kUnionMax + 1];   // prefix sums of the runs' lengths (flat candidate index -> run)
#if ARVC_STAGE_BULK
    __shared__ __align__(8) unsigned long long s_mbar[kGroupsPerBlock];
    if ((threadIdx.x & (kG - 1)) == 0) {
        const unsigned a = (unsigned)__cvta_generic_to_shared(&s_mbar[threadIdx.x / kG]);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned mbar_phase = 0;
#endif

    if ((int)blockIdx.y >= bd->n_pairs) return;
    const PairDev& pr = bd->pairs[blockIdx.y];
    const PairState* __restrict__ st = pr.state;
    if (st->done) return;
    const int pass = st->passes;         // completed passes = index of this one
    // a block beyond the work list has nothing to do whatever q turns out to be (a chunk holds at least kGroupsPerBlock
    // entries): gone before it touches the scans
    if (pass > 0 && (int)(blockIdx.x * kGroupsPerBlock) >= st->nlist) return;
    const ScanDev& src = *pr.src;
    const ScanDev& tgt = *pr.tgt;
    if ((src.wide != 0) != SW || (tgt.wide != 0) != TW) return;
    struct { double max_d2, cert_margin; int debug; } ip = {bd->ip.max_d2, bd->ip.cert_margin, bd->ip.debug};
    const int n = pass == 0 ? src.counts[CNT_NPTS] : st->nlist;       // pass 0 searches every point
    const int lane = lane_id(), gl = lane & (kG - 1), grp = threadIdx.x / kG;
    const unsigned gmask = (kG == 32) ? kFull : (((1u << kG) - 1u) << (lane & ~(kG - 1)));
    // Small batches leave most of the GPU idle while a warp works through its 32 queries (the groups of a warp run in
    // lockstep, dense unions serve their queries one after the other): when the whole batch fits a few waves of blocks,
    // every group takes only q < 8 queries - the lanes gl >= q still help with look-ups, staging and record scans - so
    // that the pass finishes after a fraction of the latency.  q is the same for all blocks of a pair (n is).
    int q = kG;
    for (int c = 1; c < kG; c <<= 1) {
        const int chunks = (n + kGroupsPerBlock * c - 1) / (kGroupsPerBlock * c);
        if (chunks <= (int)gridDim.x && (long long)chunks * bd->n_pairs <= kSpreadBlocks) { q = c; break; }
    }
    const int per_chunk = kGroupsPerBlock * q;
    // late passes run on a reduced grid (a finished pair then costs few block launches): each block strides over
    // the chunks of work-list entries; the warps of a block share nothing, so no block barrier is needed
#pragma unroll 1
    for (int chunk = blockIdx.x; chunk * per_chunk < n; chunk += gridDim.x) {
    const long long t_begin = clock64();
    const int t_idx = gl < q ? chunk * per_chunk + grp * q + gl : n;
    double T[12];                        // re-read per chunk: keeping it live across the loop costs 24 registers
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = st->T[k];
    const int i = t_idx < n ? (pass == 0 ? t_idx : pr.list[t_idx]) : -1;      // -1: padding entry
    // ---- per lane: transform the source point, bound the search with the previous pass' match
    double sx = 0, sy = 0, sz = 0;
    int level = 0;
    bool active = false, certified = true;
    float cert_lb = 0.f;                 // new certificate (0 = none)
    Best b;
    b.d2 = ip.max_d2; b.idx = -1; b.pos = -1;
    int stat_union = 0, stat_fb = 0;
    if (i >= 0) {
        double px, py, pz;
        int sidx;
        load_rec(reinterpret_cast<const SRec*>(src.recs) + i, px, py, pz, sidx);
        sx = T[0] * px + T[1] * py + T[2] * pz + T[3];
        sy = T[4] * px + T[5] * py + T[6] * pz + T[7];
        sz = T[8] * px + T[9] * py + T[10] * pz + T[11];
        const GridSpec& g = tgt.grid;
        active = (sx == sx && sy == sy && sz == sz) && ip.max_d2 > 0;
        certified = !active;
        if (pass > 0 && active) {
            const int pv = pr.prev[i];
            if (pv >= 0) {
                double x, y, z;
                int idx;
                load_rec(reinterpret_cast<const TRec*>(tgt.recs) + pv, x, y, z, idx);
                const double d2 = sqdist(sx, sy, sz, x, y, z);
                if (d2 < b.d2) {
                    // the previous match bounds the search ball; it stays unless something is strictly better
                    b.d2 = d2; b.idx = idx; b.pos = pv;
                    const double br = sqrt(d2) * (1.0 + 1e-9) + 1e-12;
                    while (level < g.top_level && g.c0 * (double)(1 << level) < br) ++level;
                }
            }
        }
    }
    // ---- union phase: the kG queries of a group are consecutive points of the Morton-sorted source, i.e. spatial
    // neighbours.  At the two finest levels their search balls share cells, so the group looks the union of the cells
    // up once, stages the records in shared memory and every lane screens every staged record against its own
    // query: no per-query set-up, no divergence, coalesced loads.  A lane is done when its best lies within the
    // level's cell edge.  The two smallest screened distances give the lane its certificate for later passes.
    {
        const GridSpec g = tgt.grid;
        const TRec* __restrict__ trecs = reinterpret_cast<const TRec*>(tgt.recs);
        uint2* runs = reinterpret_cast<uint2*>(s_stk[grp]);          // kStack * 16 bytes: kUnionMax runs (8 B) ...
        unsigned* rcell = reinterpret_cast<unsigned*>(runs + kUnionMax);   // ... + their cell offsets (4 B)
        const float sxf = (float)sx, syf = (float)sy, szf = (float)sz;
        const float e = (float)(fmax(fabs(sx), fmax(fabs(sy), fabs(sz))) * 6.0e-8 + 1e-30);
        const ScreenThr screen_thr(e);
        float thr = screen_thr(b.d2);
        const int gshift = lane & ~(kG - 1);
        // step -1 (queries without any bound, i.e. the cold pass): look only into the finest cell holding the query; whatever
        // is found there bounds the real search, whose ball then touches a few cells instead of the full 3x3x3 block
        for (int step = -1; step <= min(kUnionLevels - 1, g.top_level); ++step) {
            const int lu = max(step, 0);
            const bool probe = step < 0;
            const bool want = active && !certified && (probe ? (b.pos < 0 && level == 0) : level <= lu);
            if (!__any_sync(gmask, want) || (ip.debug & 4)) continue;
            const double cl = g.c0 * (double)(1 << lu);
            // cells are selected for a ball slightly larger than needed: the margin is what later certificates live on
            const double r = probe ? 1e-9 : fmin((double)(sqrtf(__double2float_ru(b.d2)) * (1.0f + 1e-6f)) + 1e-12 + ip.cert_margin, cl);
            int x0 = 1 << 30, x1 = -1, y0 = 1 << 30, y1 = -1, z0 = 1 << 30, z1 = -1;
            if (want) {
                x0 = cell_coord(sx - r, g.ox, g.inv_c0) >> lu; x1 = cell_coord(sx + r, g.ox, g.inv_c0) >> lu;
                y0 = cell_coord(sy - r, g.oy, g.inv_c0) >> lu; y1 = cell_coord(sy + r, g.oy, g.inv_c0) >> lu;
                z0 = cell_coord(sz - r, g.oz, g.inv_c0) >> lu; z1 = cell_coord(sz + r, g.oz, g.inv_c0) >> lu;
            }
#pragma unroll
            for (int o = kG / 2; o > 0; o >>= 1) {
                x0 = min(x0, __shfl_xor_sync(gmask, x0, o, kG)); x1 = max(x1, __shfl_xor_sync(gmask, x1, o, kG));
                y0 = min(y0, __shfl_xor_sync(gmask, y0, o, kG)); y1 = max(y1, __shfl_xor_sync(gmask, y1, o, kG));
                z0 = min(z0, __shfl_xor_sync(gmask, z0, o, kG)); z1 = max(z1, __shfl_xor_sync(gmask, z1, o, kG));
            }
            const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, ncell = nx * ny * (z1 - z0 + 1);
            // sparse regions / jumps of the space-filling curve: the 8 queries span too many cells here, try a coarser level
            if (ncell > kUnionMax) continue;
            int nr = 0, utotal = 0;
            const CellDecoder dec(nx, ny);
            for (int base = 0; base < ncell; base += kG) {
                const int t = base + gl;
                bool valid = false;
                unsigned rs = 0, re = 0;
                int ox = 0, oy = 0, oz = 0;
                if (t < ncell) {
                    dec(t, ox, oy, oz);
                    valid = grid_lookup(tgt.table, tgt.table_mask, lu, morton3(x0 + ox, y0 + oy, z0 + oz), rs, re);
                }
                const unsigned vm = (__ballot_sync(gmask, valid) >> gshift) & kGMask;
                if (valid) {
                    const int slot = nr + __popc(vm & ((1u << gl) - 1u));
                    runs[slot] = make_uint2(rs, re);
                    rcell[slot] = (unsigned)ox | ((unsigned)oy << 8) | ((unsigned)oz << 16);   // offsets from x0,y0,z0
                }
                nr += __popc(vm);
                utotal += (int)(re - rs);
            }
#pragma unroll
            for (int o = kG / 2; o > 0; o >>= 1) utotal += __shfl_xor_sync(gmask, utotal, o, kG);
            __syncwarp(gmask);
            // coarse levels pay off only where the scan is sparse; dense unions are left to the per-query search
            if (lu >= 2 && utotal > kUnionCandCap) break;
            float f1 = INFINITY, f2 = INFINITY;     // two smallest screened float32 distances (multiset)
            // The sweeps below are branch-free: they only track the two smallest float32 distances and WHERE the smallest
            // one was seen.  Afterwards the lane settles its query exactly: every staged record that could beat or tie
            // with the best so far has a float32 distance <= thr (see ScreenThr), so if the second smallest one is above
            // thr the only candidate is the record at `jmin` - one exact evaluation.  Otherwise (two records within the
            // float32 error of each other: duplicates, near ties) the lane re-reads the union and decides every record
            // below thr exactly, as a plain sequential scan.  Results are those of an all-float64 search.
            auto resolve = [&](int jmin, float m1, float m2) {
                if constexpr (!TW) {
                    if (!(m1 <= thr) || jmin < 0) return;
                    const float4* __restrict__ tr4 = reinterpret_cast<const float4*>(trecs);
                    {
                        const float4 v = __ldg(tr4 + jmin);
                        const double d2 = sqdist(sx, sy, sz, (double)v.x, (double)v.y, (double)v.z);
                        const int idx = __float_as_int(v.w);
                        if (d2 < b.d2 || (d2 == b.d2 && idx < b.idx)) { b.d2 = d2; b.idx = idx; b.pos = jmin; thr = screen_thr(d2); }
                    }
                    if (m2 <= thr) {
                        for (int rr = 0; rr < nr; ++rr) {
                            const uint2 run = runs[rr];
                            for (unsigned p = run.x; p < run.y; ++p) {
                                const float4 v = __ldg(tr4 + p);
                                const float dx = sxf - v.x, dy = syf - v.y, dz = szf - v.z;
                                if (fmaf(dz, dz, fmaf(dy, dy, dx * dx)) <= thr) {
                                    const double d2 = sqdist(sx, sy, sz, (double)v.x, (double)v.y, (double)v.z);
                                    const int idx = __float_as_int(v.w);
                                    if (d2 < b.d2 || (d2 == b.d2 && idx < b.idx)) { b.d2 = d2; b.idx = idx; b.pos = (int)p; thr = screen_thr(d2); }
                                }
                            }
                        }
                    }
                }
            };
            if (!TW && utotal > kUnionScanCap) {
                // Dense union: every lane screening every record would cost lanes x records.  Serve the wanting lanes one
                // at a time instead: the 8 lanes split the records of the cells that touch that lane's own ball (the
                // union's run table is reused, no further lookups), then agree on the (d2, index) minimum.
                const float4* __restrict__ tr4 = reinterpret_cast<const float4*>(trecs);
                for (int k = 0; k < kG; ++k) {
                    if (!__shfl_sync(gmask, (int)want, k, kG)) continue;
                    const double qx = __shfl_sync(gmask, sx, k, kG), qy = __shfl_sync(gmask, sy, k, kG), qz = __shfl_sync(gmask, sz, k, kG);
                    const double qr = __shfl_sync(gmask, r, k, kG);
                    const int kx0 = (cell_coord(qx - qr, g.ox, g.inv_c0) >> lu) - x0, kx1 = (cell_coord(qx + qr, g.ox, g.inv_c0) >> lu) - x0;
                    const int ky0 = (cell_coord(qy - qr, g.oy, g.inv_c0) >> lu) - y0, ky1 = (cell_coord(qy + qr, g.oy, g.inv_c0) >> lu) - y0;
                    const int kz0 = (cell_coord(qz - qr, g.oz, g.inv_c0) >> lu) - z0, kz1 = (cell_coord(qz + qr, g.oz, g.inv_c0) >> lu) - z0;
                    const float qxf = (float)qx, qyf = (float)qy, qzf = (float)qz;
                    float g1 = INFINITY, g2 = INFINITY;
                    int gj = -1;
                    for (int rr = 0; rr < nr; ++rr) {
                        const unsigned rc = rcell[rr];
                        const int ox = (int)(rc & 255u), oy = (int)((rc >> 8) & 255u), oz = (int)(rc >> 16);
                        if (ox < kx0 || ox > kx1 || oy < ky0 || oy > ky1 || oz < kz0 || oz > kz1) continue;
                        const uint2 run = runs[rr];
                        for (unsigned p0 = run.x + gl; p0 < run.y; p0 += 4 * kG) {      // four loads in flight per lane
                            float4 v4[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const unsigned p = p0 + u * kG;
                                v4[u] = p < run.y ? __ldg(tr4 + p) : make_float4(INFINITY, 0.f, 0.f, 0.f);
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float4 v = v4[u];
                                const float dx = qxf - v.x, dy = qyf - v.y, dz = qzf - v.z;
                                const float d2f = fmaf(dz, dz, fmaf(dy, dy, dx * dx));     // +inf for the padding lanes
                                gj = d2f < g1 ? (int)(p0 + u * kG) : gj;
                                g2 = fminf(g2, fmaxf(g1, d2f));
                                g1 = fminf(g1, d2f);
                            }
                        }
                    }
#pragma unroll
                    for (int o = kG / 2; o > 0; o >>= 1) {
                        const float h1 = __shfl_xor_sync(gmask, g1, o, kG), h2 = __shfl_xor_sync(gmask, g2, o, kG);
                        const int hj = __shfl_xor_sync(gmask, gj, o, kG);
                        g2 = fminf(fmaxf(g1, h1), fminf(g2, h2));      // two smallest of the merged multisets
                        gj = h1 < g1 ? hj : gj;
                        g1 = fminf(g1, h1);
                    }
                    if (gl == k) {      // the lane that owns the query settles it (ties between lanes show up as g2 == g1 <= thr)
                        f1 = g1; f2 = g2;
                        resolve(gj, g1, g2);
                    }
                }
            } else
            if constexpr (!TW) {
                float4* stage = s_stage[grp];
                int* roff = s_roff[grp];
                int jmin = -1;
                // The union's records form one flat sequence (the runs back to back): prefix sums of the run lengths map a
                // flat index to its run, so that staging needs no per-run loop - short runs (a handful of records per fine
                // cell) would leave most lanes idle.  Lane gl owns runs 8 gl .. 8 gl + 7 for the scan.
                {
                    int loc = 0, tmp[kUnionMax / kG];
#pragma unroll
                    for (int k = 0; k < kUnionMax / kG; ++k) {
                        const int r = gl * (kUnionMax / kG) + k;
                        tmp[k] = loc;
                        loc += r < nr ? (int)(runs[r].y - runs[r].x) : 0;
                    }
                    int inc = loc;
#pragma unroll
                    for (int o = 1; o < kG; o <<= 1) {
                        const int t = __shfl_up_sync(gmask, inc, o, kG);
                        if (gl >= o) inc += t;
                    }
#pragma unroll
                    for (int k = 0; k < kUnionMax / kG; ++k) {
                        const int r = gl * (kUnionMax / kG) + k;
                        if (r < nr) roff[r] = inc - loc + tmp[k];
                    }
                    if (gl == 0) roff[nr] = utotal;
                    __syncwarp(gmask);
                }
                // flat index of a staged record -> its position in the target's record array (6 halving steps over <= 64 runs)
                auto flat_to_pos = [&](int f) -> int {
                    int lo = 0, hi = nr - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (roff[mid] <= f) lo = mid; else hi = mid - 1;
                    }
                    return (int)runs[lo].x + (f - roff[lo]);
                };
#if ARVC_STAGE_BULK
                // Variant: one bulk asynchronous copy (cp.async.bulk, the non-tensor TMA path) per run piece, completion through
                // the group's mbarrier.  Measured slower than per-record LDGSTS for these short runs - kept
                // selectable for the comparison.
                unsigned long long* mbar = &s_mbar[grp];
                const unsigned mbar_a = (unsigned)__cvta_generic_to_shared(mbar);
                for (int base = 0; base < utotal; base += kStage) {
                    const int fill = min(kStage, utotal - base);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic reads of the last round before async writes
                    if (gl == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar_a), "r"(fill * 16) : "memory");
                    __syncwarp(gmask);
                    for (int c = gl; c < nr; c += kG) {
                        const int lo = max(roff[c], base), hi = min(roff[c + 1], base + fill);
                        if (lo < hi) {
                            const unsigned dst = (unsigned)__cvta_generic_to_shared(stage + (lo - base));
                            const TRec* src_p = trecs + runs[c].x + (unsigned)(lo - roff[c]);
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         :: "r"(dst), "l"(src_p), "r"((hi - lo) * 16), "r"(mbar_a) : "memory");
                        }
                    }
                    {
                        unsigned done = 0;
                        while (!done)
                            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                         : "=r"(done) : "r"(mbar_a), "r"(mbar_phase) : "memory");
                        mbar_phase ^= 1u;
                    }
#else
                int cur = 0;
                for (int base = 0; base < utotal; base += kStage) {
                    const int fill = min(kStage, utotal - base);
                    // asynchronous 16-byte copies global -> shared (LDGSTS): every lane has its kStage / kG records in
                    // flight at once, nothing passes through registers; the flat index advances monotonically per lane
#pragma unroll
                    for (int m = 0; m < kStage / kG; ++m) {
                        const int q = gl + m * kG;
                        if (q < fill) {
                            const int f = base + q;
                            while (f >= roff[cur + 1]) ++cur;
                            const unsigned pos = runs[cur].x + (unsigned)(f - roff[cur]);
                            const unsigned dst = (unsigned)__cvta_generic_to_shared(stage + q);
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(trecs + pos) : "memory");
                        }
                    }
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    __syncwarp(gmask);
#endif
                    int jloc = -1;      // where in THIS round the smallest float32 distance so far was seen
#pragma unroll 8
                    for (int j = 0; j < fill; ++j) {
                        const float4 v = stage[j];
                        const float dx = sxf - v.x, dy = syf - v.y, dz = szf - v.z;
                        const float d2f = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        jloc = d2f < f1 ? j : jloc;
                        f2 = fminf(f2, fmaxf(f1, d2f));
                        f1 = fminf(f1, d2f);
                    }
                    if (jloc >= 0) jmin = flat_to_pos(base + jloc);
                    __syncwarp(gmask);
                }
                if (want) resolve(jmin, f1, f2);
            } else {
                for (int rr = 0; rr < nr; ++rr) {
                    const uint2 run = runs[rr];
                    for (unsigned p = run.x; p < run.y; ++p) {
                        double x, y, z;
                        int idx;
                        load_rec(trecs + p, x, y, z, idx);
                        const double d2 = sqdist(sx, sy, sz, x, y, z);
                        if (want && (d2 < b.d2 || (d2 == b.d2 && idx < b.idx))) { b.d2 = d2; b.idx = idx; b.pos = (int)p; }
                    }
                }
            }
            __syncwarp(gmask);
            if (want && !probe) {
                // every target point within min(previous bound, cl) of this lane's query was screened
                if (sqrt(b.d2) * (1.0 + 1e-9) + 1e-12 <= cl) {
                    certified = true;
                    ++stat_union;
                    if (!TW && b.pos >= 0) {
                        // lower bound on the distance to every target point but the match: screened points via the second
                        // smallest float32 distance (minus its error bound), all others lie outside the selected cells
                        const float low2 = f2 * (1.0f - 4e-6f) - 3.6f * e * sqrtf(f2) - 3.1f * e * e;
                        const float l2 = low2 > 0.f ? sqrtf(low2) * (1.0f - 1e-6f) : 0.f;
                        cert_lb = fminf(l2, (float)r * (1.0f - 1e-6f));
                    }
                } else {
                    level = lu + 1;
                }
            }
        }
    }
    const long long t_union_end = clock64();
    // ---- remaining queries (nearest neighbour beyond the union levels): handed to k_icp_far through the pair's far
    // list, one entry per query, so that they spread over the whole GPU instead of queueing inside this warp.  The
    // bound travels implicitly: prev[i] (written below) holds the best match so far, k_icp_far re-derives its distance.
    {
        const bool pending = active && !certified && !(ip.debug & 2);
        const unsigned pend = __ballot_sync(kFull, pending);
        if (pend) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&pr.state->nfar, __popc(pend));
            base = __shfl_sync(kFull, base, 0);
            if (pending) { pr.far_list[base + __popc(pend & ((1u << lane) - 1u))] = i | (level << 24); ++stat_fb; }
        }
    }

    if (i >= 0) {
        pr.prev[i] = b.pos;
        const bool ok = cert_lb > 0.f && pass < kThist;
        pr.cert_pass[i] = ok ? (unsigned char)pass : (unsigned char)255;
        if (ok) pr.lb2[i] = cert_lb;
    }
    if ((ip.debug & 8) && lane == 0 && pass > 0) {
        const long long dt = clock64() - t_begin;
        int bkt = 0;
        while (bkt < 6 && dt > (20000ll << bkt)) ++bkt;      // 20k cycles ~ 10 us, doubling
        atomicAdd(&pr.state->dbg[1 + bkt], 1ull);
        atomicMax(&pr.state->dbg[0], (unsigned long long)dt);
        if (dt > 400000) printf("SLOW warp pass=%d pair=%d blk=%d total=%.0fus union=%.0fus fallback=%.0fus\n", pass, (int)blockIdx.y, chunk, dt / 1965.0, (t_union_end - t_begin) / 1965.0, (clock64() - t_union_end) / 1965.0);
    }
    if (ip.debug & 1) {
        const int u = warp_sum(stat_union), f = warp_sum(stat_fb), ac = warp_sum(i >= 0 ? 1 : 0);
        if (lane == 0) { atomicAdd(&pr.state->dbg[0], (unsigned long long)ac); atomicAdd(&pr.state->dbg[1], (unsigned long long)u);
                         atomicAdd(&pr.state->dbg[2], (unsigned long long)f); }
    }
    __syncwarp();
    }   // chunk loop
}

// Far queries of a pass (their nearest neighbour lies beyond the cells the shared-candidate phase covers): one 8-lane
// group per far-list entry, cooperative branch and bound over the coarser levels, groups striding over the list.
template <bool SW, bool TW, int OCC, int FG /* lanes per far query */>
__global__ void __launch_bounds__(kIcpBlock, OCC) k_icp_far(const BatchDesc* __restrict__ bd) {
    typedef typename RecT<SW>::type SRec;
    typedef typename RecT<TW>::type TRec;
    constexpr int kFarGroups = kIcpBlock / FG;
    constexpr int kFarStack = FG >= 8 ? kStack : 40;     // 27 cells of a ball + one expansion fit; deeper stacks cost occupancy
    __shared__ uint4 s_stk[kFarGroups][kFarStack];
    __shared__ float s_stk_lb[kFarGroups][kFarStack];
    if ((int)blockIdx.y >= bd->n_pairs) return;
    const PairDev& pr = bd->pairs[blockIdx.y];
    const PairState* __restrict__ st = pr.state;
    if (st->done) return;
    const int nfar = st->nfar;
    if ((int)(blockIdx.x * kFarGroups) >= nfar) return;      // most blocks of a late pass: gone before they touch the scans
    const ScanDev& src = *pr.src;
    const ScanDev& tgt = *pr.tgt;
    if ((src.wide != 0) != SW || (tgt.wide != 0) != TW) return;
    const int lane = lane_id(), gl = lane & (FG - 1), grp = threadIdx.x / FG;
    const unsigned gmask = FG >= 32 ? kFull : (((1u << (FG & 31)) - 1u) << (lane & ~(FG - 1)));
    const double max_d2 = bd->ip.max_d2;
    double T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = st->T[k];
#pragma unroll 1
    for (int f = blockIdx.x * kFarGroups + grp; f < nfar; f += gridDim.x * kFarGroups) {
        const int e = pr.far_list[f];
        const int i = e & 0xffffff, level = e >> 24;
        double px, py, pz;
        int sidx;
        load_rec(reinterpret_cast<const SRec*>(src.recs) + i, px, py, pz, sidx);
        const double sx = T[0] * px + T[1] * py + T[2] * pz + T[3];
        const double sy = T[4] * px + T[5] * py + T[6] * pz + T[7];
        const double sz = T[8] * px + T[9] * py + T[10] * pz + T[11];
        // the bound: best match so far (previous pass' match or what the shared-candidate phase found), else the cut-off
        double bd2 = max_d2;
        int bidx = -1;
        const int pv = pr.prev[i];
        if (pv >= 0) {
            double x, y, z;
            int idx;
            load_rec(reinterpret_cast<const TRec*>(tgt.recs) + pv, x, y, z, idx);
            const double d2 = sqdist(sx, sy, sz, x, y, z);
            if (d2 < bd2) { bd2 = d2; bidx = idx; }
        }
        Best lb;
        group_search<TW, FG, kFarStack>(tgt, sx, sy, sz, bd2, bidx, level, lb, s_stk[grp], s_stk_lb[grp], gl, gmask);
        if (gl == 0 && lb.pos >= 0) pr.prev[i] = lb.pos;      // else: nothing strictly better than the bound, prev[i] stands
        __syncwarp(gmask);
    }
}

constexpr int kAccPts = 4;                                  // source points per lane of the accumulation kernel
constexpr int kAccBlockPts = kIcpBlock * kAccPts;

template <int METHOD, bool SW, bool TW>
__global__ void __launch_bounds__(kIcpBlock) k_icp_accum(const BatchDesc* __restrict__ bd) {
    typedef typename RecT<SW>::type SRec;
    typedef typename RecT<TW>::type TRec;
    if ((int)blockIdx.y >= bd->n_pairs) return;
    const PairDev& pr = bd->pairs[blockIdx.y];
    const ScanDev& src = *pr.src;
    const ScanDev& tgt = *pr.tgt;
    if ((src.wide != 0) != SW || (tgt.wide != 0) != TW) return;
    const PairState* __restrict__ st = pr.state;
    if (st->done) return;
    const int pass = st->passes;
    const int n = src.counts[CNT_NPTS];
    const int nblk = max(1, (n + kAccBlockPts - 1) / kAccBlockPts);
    const int lane = lane_id();
#pragma unroll 1
    for (int chunk = blockIdx.x; chunk < nblk; chunk += gridDim.x) {
    const int base = chunk * kAccBlockPts + (threadIdx.x >> 5) * (32 * kAccPts) + lane;
    double acc[kNS];
#pragma unroll
    for (int k = 0; k < kNS; ++k) acc[k] = 0.0;
    const double* T = st->T;
    const double T0 = T[0], T1 = T[1], T2 = T[2], T3 = T[3], T4 = T[4], T5 = T[5], T6 = T[6], T7 = T[7], T8 = T[8], T9 = T[9],
                 T10 = T[10], T11 = T[11];
    int posv[kAccPts];                   // the matches of all the lane's points first: one exposed latency instead of four
#pragma unroll
    for (int j = 0; j < kAccPts; ++j) posv[j] = base + j * 32 < n ? pr.prev[base + j * 32] : -1;
    // all gathers of the lane's points are issued before the arithmetic (clamped indices, results of invalid points unused)
    double pxv[kAccPts], pyv[kAccPts], pzv[kAccPts], txv[kAccPts], tyv[kAccPts], tzv[kAccPts];
    int sidxv[kAccPts], tidxv[kAccPts];
    double4 nvv[kAccPts];
#pragma unroll
    for (int j = 0; j < kAccPts; ++j) {
        const int i = min(base + j * 32, max(n - 1, 0));
        const int pos = max(posv[j], 0);
        load_rec(reinterpret_cast<const SRec*>(src.recs) + i, pxv[j], pyv[j], pzv[j], sidxv[j]);
        load_rec(reinterpret_cast<const TRec*>(tgt.recs) + pos, txv[j], tyv[j], tzv[j], tidxv[j]);
        if (METHOD == 1) nvv[j] = reinterpret_cast<const double4*>(tgt.normals)[pos];
    }
#pragma unroll
    for (int j = 0; j < kAccPts; ++j) {
        const int i = base + j * 32;
        if (i >= n) break;
        const int pos = posv[j];
        const int sidx = sidxv[j];
        if (pos < 0) {
            if (pr.corr_trace) pr.corr_trace[(size_t)pass * src.cap + sidx] = -1;
            continue;
        }
        const double px = pxv[j], py = pyv[j], pz = pzv[j];
        const double sx = T0 * px + T1 * py + T2 * pz + T3;
        const double sy = T4 * px + T5 * py + T6 * pz + T7;
        const double sz = T8 * px + T9 * py + T10 * pz + T11;
        const double tx = txv[j], ty = tyv[j], tz = tzv[j];
        const int tidx = tidxv[j];
        acc[27] += sqdist(sx, sy, sz, tx, ty, tz);
        acc[28] += 1.0;
        if (pr.corr_trace) pr.corr_trace[(size_t)pass * src.cap + sidx] = tidx;
        if (METHOD == 1) {
            const double4 nv = nvv[j];
            const double r = (sx - tx) * nv.x + (sy - ty) * nv.y + (sz - tz) * nv.z;
            const double J[6] = {sy * nv.z - sz * nv.y, sz * nv.x - sx * nv.z, sx * nv.y - sy * nv.x, nv.x, nv.y, nv.z};
            int t = 0;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int c = a; c < 6; ++c) acc[t++] += J[a] * J[c];
#pragma unroll
            for (int a = 0; a < 6; ++a) acc[21 + a] += J[a] * r;
        } else {
            acc[0] += sx; acc[1] += sy; acc[2] += sz; acc[3] += tx; acc[4] += ty; acc[5] += tz;
            acc[6] += tx * sx; acc[7] += tx * sy; acc[8] += tx * sz;
            acc[9] += ty * sx; acc[10] += ty * sy; acc[11] += ty * sz;
            acc[12] += tz * sx; acc[13] += tz * sy; acc[14] += tz * sz;
        }
    }
    double* row = pr.partials + ((size_t)chunk * (kIcpBlock / 32) + (threadIdx.x >> 5)) * kSumStride;
#pragma unroll
    for (int k = 0; k < kNS; ++k) {
        if (METHOD == 0 && k >= 15 && k < 27) continue;
        const double v = warp_sum(acc[k]);
        if (lane == 0) row[k] = v;
    }
    }   // chunk loop
}

// Finish kernel, one block per pair: fixed-order reduction of the warps' partial sums, solve, cumulative
// transformation update and Open3D's convergence test.  Sets the `done` flag that turns later passes into no-ops.
template <int METHOD>
__global__ void __launch_bounds__(256) k_icp_finish(const BatchDesc* __restrict__ bd) {
    __shared__ double s_part[8][kSumStride];
    __shared__ double s_sum[kSumStride];
    if ((int)blockIdx.x >= bd->n_pairs) return;
    const PairDev& pr = bd->pairs[blockIdx.x];
    PairState* st = pr.state;
    if (st->done) return;
    const int pass = st->passes;
    const IcpParams ip = bd->ip;
    const ScanDev& src = *pr.src;
    const ScanDev& tgt = *pr.tgt;
    const int n = src.counts[CNT_NPTS];
    const int nrows = max(1, (n + kAccBlockPts - 1) / kAccBlockPts) * (kIcpBlock / 32);
    const int lane = lane_id(), w = threadIdx.x >> 5;
    double v = 0;
    for (int r = w; r < nrows; r += 8) v += pr.partials[(size_t)r * kSumStride + lane];
    s_part[w][lane] = v;
    __syncthreads();
    if (threadIdx.x < kSumStride) {
        double t = 0;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) t += s_part[ww][threadIdx.x];
        s_sum[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double K = s_sum[28];
        const double fitness = (K > 0 && n > 0) ? K / (double)n : 0.0;
        const double rmse = K > 0 ? sqrt(s_sum[27] / K) : 0.0;
        if (pr.state_trace) {
            double* tr = pr.state_trace + (size_t)pass * 18;
            for (int k = 0; k < 16; ++k) tr[k] = st->T[k];
            tr[16] = fitness; tr[17] = rmse;
        }
        bool stop = pass >= ip.max_iter;
        if (pass > 0 && fabs(st->fitness - fitness) < ip.rel_fitness && fabs(st->rmse - rmse) < ip.rel_rmse) stop = true;
        st->fitness = fitness;
        st->rmse = rmse;
        st->ncorr = (int)K;
        st->passes = pass + 1;
        for (int k = 0; k < kNS; ++k) st->sums[k] = s_sum[k];
        if (ip.debug & 1) st->dbg[3] += (unsigned long long)n;
        st->nlist = 0;
        st->nfar = 0;
        if (stop) {
            st->done = 1;
        } else {
            double upd[16], Tn[16];
            solve_update(METHOD, s_sum, upd);
            mat4_mul(upd, st->T, Tn);
            for (int k = 0; k < 16; ++k) st->T[k] = Tn[k];
            if (pass + 1 < kThist) for (int k = 0; k < 12; ++k) st->Thist[12 * (pass + 1) + k] = Tn[k];
            st->updates = st->updates + 1;
        }
        st->err = src.counts[CNT_ERR] | tgt.counts[CNT_ERR];
    }
}

// Batch set-up and delivery, one thread per pair: the pair states are initialised on the device from the compact array
// of initial guesses (128 B per pair instead of the whole state), and the results leave as 160-byte records - the unit
// the multi-GPU gather moves - plus one status word (error flags of the scans, 0x100 = a pair did not terminate).
__global__ void __launch_bounds__(128) k_icp_init(const PairDev* __restrict__ pairs, int n_pairs, const double* __restrict__ init_T, int* __restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *status = 0;
    if (i >= n_pairs) return;
    PairState* st = pairs[i].state;
    for (int k = 0; k < 16; ++k) st->T[k] = init_T[16 * (size_t)i + k];
    for (int k = 0; k < 12; ++k) st->Thist[k] = init_T[16 * (size_t)i + k];
    st->fitness = 0.0; st->rmse = 0.0;
    st->passes = 0; st->updates = 0; st->done = 0; st->ncorr = 0; st->err = 0; st->nlist = 0; st->nfar = 0;
    for (int k = 0; k < 8; ++k) st->dbg[k] = 0ull;
}

struct __align__(8) ResultRecordDev { int pair, updates, n_corr, passes; double T[16]; double fitness, rmse; };
static_assert(sizeof(ResultRecordDev) == 160, "result records are 160 bytes (include/arvc_icp.h)");

__global__ void __launch_bounds__(128) k_icp_pack(const PairDev* __restrict__ pairs, int n_pairs, ResultRecordDev* __restrict__ out, int* __restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const PairState* st = pairs[i].state;
    ResultRecordDev r;
    r.pair = i; r.updates = st->updates; r.n_corr = st->ncorr; r.passes = st->passes;
    for (int k = 0; k < 16; ++k) r.T[k] = st->T[k];
    r.fitness = st->fitness; r.rmse = st->rmse;
    out[i] = r;
    const int flags = st->err | (st->done ? 0 : 0x100);
    if (flags) atomicOr(status, flags);
}

void launch_icp_init(Launcher& L, const PairDev* d_pairs, int n_pairs, const double* d_init, int* d_status) {
    L.launch("icp_init", k_icp_init, dim3((n_pairs + 127) / 128), dim3(128), d_pairs, n_pairs, d_init, d_status);
}
void launch_icp_pack(Launcher& L, const PairDev* d_pairs, int n_pairs, void* d_records, int* d_status) {
    L.launch("icp_pack", k_icp_pack, dim3((n_pairs + 127) / 128), dim3(128), d_pairs, n_pairs, reinterpret_cast<ResultRecordDev*>(d_records), d_status);
}

// WHILE condition of the device-terminated loop: non-zero while some pair of the batch has not converged.
__global__ void __launch_bounds__(256) k_icp_cond(const BatchDesc* __restrict__ bd, cudaGraphConditionalHandle handle) {
    int open_pairs = 0;
    for (int i = threadIdx.x; i < bd->n_pairs; i += blockDim.x) open_pairs |= bd->pairs[i].state->done ? 0 : 1;
    const int any = __syncthreads_or(open_pairs);
    if (threadIdx.x == 0) cudaGraphSetConditional(handle, any ? 1u : 0u);
}

namespace {

const char* pass_name(int pass) {   // "icp_pass_00" ... so that the profile report separates the passes
    static char names[64][16];
    static bool init = false;
    if (!init) {
        for (int i = 0; i < 64; ++i) snprintf(names[i], sizeof(names[i]), "icp_pass_%02d", i);
        init = true;
    }
    return names[pass < 63 ? pass : 63];
}

// Receives the kernels of one pass in order: either launches them on the stream or chains them as graph nodes.
struct Emitter {
    Launcher* L = nullptr;                      // stream mode
    cudaGraph_t graph = nullptr;                // graph mode
    cudaGraphNode_t last = nullptr;
    cudaError_t err = cudaSuccess;
    int count = 0;
    void kernel(const char* name, const void* func, dim3 grid, dim3 block, void** args) {
        ++count;
        if (L) { L->launch_ptr(name, func, grid, block, args); return; }
        if (err != cudaSuccess) return;
        cudaKernelNodeParams kp{};
        kp.func = const_cast<void*>(func);
        kp.gridDim = grid; kp.blockDim = block; kp.sharedMemBytes = 0; kp.kernelParams = args; kp.extra = nullptr;
        cudaGraphNode_t node = nullptr;
        err = cudaGraphAddKernelNode(&node, graph, last ? &last : nullptr, last ? 1 : 0, &kp);
        last = node;
    }
};

// Pairs converge after ~5-12 passes: the later the pass, the smaller the grid (the kernels stride over their chunks),
// so that the blocks of finished pairs cost next to nothing.  `pass` >= kIcpUnrolled stands for the loop body.
template <int METHOD>
void emit_pass(Emitter& E, const BatchDesc* d_bd, int gx_search, int cap_max, int n_pairs_grid, int pass, int combos_mask) {
    const char* nm = pass_name(pass);
    const int sh_pts = pass < 4 ? 0 : (pass < 10 ? 2 : 3), sh_search = pass < 3 ? 0 : (pass < 5 ? 1 : 2);
    const int mult = n_pairs_grid <= 2 ? 4 : (n_pairs_grid <= 8 ? 2 : 1);      // small batches: room to spread the queries (blocks without work exit)
    const dim3 g256(max(1, ((cap_max + 255) / 256) >> sh_pts), n_pairs_grid);
    const dim3 gsearch(max(1, gx_search >> sh_search) * mult, n_pairs_grid);
    const dim3 gacc(max(1, ((cap_max + kAccBlockPts - 1) / kAccBlockPts) >> sh_pts), n_pairs_grid);
    void* args[] = {(void*)&d_bd};
    if (pass > 0) {
        if (combos_mask & 1) E.kernel("icp_select", (const void*)k_icp_select<false, false>, g256, dim3(256), args);
        if (combos_mask & 2) E.kernel("icp_select", (const void*)k_icp_select<false, true>, g256, dim3(256), args);
        if (combos_mask & 4) E.kernel("icp_select", (const void*)k_icp_select<true, false>, g256, dim3(256), args);
        if (combos_mask & 8) E.kernel("icp_select", (const void*)k_icp_select<true, true>, g256, dim3(256), args);
    }
    // Large batches are throughput bound - more resident warps pay for a few spilled registers; small ones (the 12 pairs of
    // a 128-beam step, one pair per call) are latency bound and want the spill-free build.  Float32 records only.
    static const int dense_min = getenv("ARVC_DENSE_BATCH") ? atoi(getenv("ARVC_DENSE_BATCH")) : kDenseBatch;   // profiling knob
    const bool dense_batch = n_pairs_grid >= dense_min;
    if (combos_mask & 1) E.kernel(nm, dense_batch ? (const void*)k_icp_search<false, false, ARVC_SEARCH_OCC_DENSE> : (const void*)k_icp_search<false, false, ARVC_SEARCH_OCC>, gsearch, dim3(kIcpBlock), args);
    if (combos_mask & 2) E.kernel(nm, (const void*)k_icp_search<false, true, ARVC_SEARCH_OCC>, gsearch, dim3(kIcpBlock), args);
    if (combos_mask & 4) E.kernel(nm, (const void*)k_icp_search<true, false, ARVC_SEARCH_OCC>, gsearch, dim3(kIcpBlock), args);
    if (combos_mask & 8) E.kernel(nm, (const void*)k_icp_search<true, true, ARVC_SEARCH_OCC>, gsearch, dim3(kIcpBlock), args);
    // far queries: a few per cent of the points; 296 blocks x 8 groups per pair stride over the pair's far list
    const int sh_far = pass < ARVC_FARSH0 ? 0 : (pass < ARVC_FARSH1 ? 1 : 2);      // far lists shrink with the passes
    const dim3 gfar((kFarBlocks >> sh_far) * mult, n_pairs_grid);
    static char far_names[64][16];
    snprintf(far_names[pass < 63 ? pass : 63], 16, "icp_far_%02d", pass < 63 ? pass : 63);
    const char* fn = far_names[pass < 63 ? pass : 63];
    // one or two pairs per call (the reference's own loop): a pass is as long as its slowest far query, so the whole warp
    // serves one query (scans 4x as wide; 0.89 -> 0.83 ms per sequential pair); larger batches are throughput bound, where
    // 8-lane groups do the least total work
    const bool tiny_batch = n_pairs_grid <= 2;
    if (combos_mask & 1) E.kernel(fn, tiny_batch ? (const void*)k_icp_far<false, false, ARVC_FAR_OCC, 32>
                                      : dense_batch ? (const void*)k_icp_far<false, false, ARVC_FAR_OCC_DENSE, ARVC_FAR_G>
                                                    : (const void*)k_icp_far<false, false, ARVC_FAR_OCC, ARVC_FAR_G>, gfar, dim3(kIcpBlock), args);
    if (combos_mask & 2) E.kernel(fn, (const void*)k_icp_far<false, true, ARVC_FAR_OCC, ARVC_FAR_G>, gfar, dim3(kIcpBlock), args);
    if (combos_mask & 4) E.kernel(fn, (const void*)k_icp_far<true, false, ARVC_FAR_OCC, ARVC_FAR_G>, gfar, dim3(kIcpBlock), args);
    if (combos_mask & 8) E.kernel(fn, (const void*)k_icp_far<true, true, ARVC_FAR_OCC, ARVC_FAR_G>, gfar, dim3(kIcpBlock), args);
    if (combos_mask & 1) E.kernel("icp_accum", (const void*)k_icp_accum<METHOD, false, false>, gacc, dim3(kIcpBlock), args);
    if (combos_mask & 2) E.kernel("icp_accum", (const void*)k_icp_accum<METHOD, false, true>, gacc, dim3(kIcpBlock), args);
    if (combos_mask & 4) E.kernel("icp_accum", (const void*)k_icp_accum<METHOD, true, false>, gacc, dim3(kIcpBlock), args);
    if (combos_mask & 8) E.kernel("icp_accum", (const void*)k_icp_accum<METHOD, true, true>, gacc, dim3(kIcpBlock), args);
    E.kernel("icp_finish", (const void*)k_icp_finish<METHOD>, dim3(n_pairs_grid), dim3(256), args);
}

void emit_pass_m(Emitter& E, int method, const BatchDesc* d_bd, int gx_search, int cap_max, int n_pairs_grid, int pass, int combos_mask) {
    if (method == 1) emit_pass<1>(E, d_bd, gx_search, cap_max, n_pairs_grid, pass, combos_mask);
    else emit_pass<0>(E, d_bd, gx_search, cap_max, n_pairs_grid, pass, combos_mask);
}

// batch shapes are bucketed so that a handful of instantiated graphs serves every batch size
int bucket_pairs(int n) {
    if (n <= 8) return n;
    if (n <= 64) return (n + 7) / 8 * 8;
    if (n <= 512) return (n + 31) / 32 * 32;
    return (n + 255) / 256 * 256;
}
int bucket_cap(int cap) { return (cap + 4095) / 4096 * 4096; }

const IcpGraph* get_graph(IcpGraphCache& cache, int method, int combos_mask, int cap_b, int pairs_b) {
    const unsigned long long key = ((unsigned long long)method << 60) | ((unsigned long long)combos_mask << 56) |
                                   ((unsigned long long)pairs_b << 32) | (unsigned long long)cap_b;
    for (const IcpGraph& g : cache.graphs) if (g.key == key) return &g;
    IcpGraph G;
    G.key = key;
    auto fail = [&](cudaError_t e, const char* what) -> const IcpGraph* {
        cache.error = std::string(what) + ": " + cudaGetErrorString(e);
        cache.disabled = true;
        if (G.exec) cudaGraphExecDestroy(G.exec);
        if (G.graph) cudaGraphDestroy(G.graph);
        cudaGetLastError();
        return nullptr;
    };
    cudaError_t e = cudaGraphCreate(&G.graph, 0);
    if (e != cudaSuccess) return fail(e, "cudaGraphCreate");
    const int gx_search = max(1, (cap_b + kIcpBlock - 1) / kIcpBlock);
    const BatchDesc* d_bd = cache.d_bd;
    Emitter E;
    E.graph = G.graph;
    for (int pass = 0; pass < kIcpUnrolled; ++pass) emit_pass_m(E, method, d_bd, gx_search, cap_b, pairs_b, pass, combos_mask);
    if (E.err != cudaSuccess) return fail(E.err, "cudaGraphAddKernelNode");
    cudaGraphConditionalHandle handle;
    e = cudaGraphConditionalHandleCreate(&handle, G.graph, 1, cudaGraphCondAssignDefault);
    if (e != cudaSuccess) return fail(e, "cudaGraphConditionalHandleCreate");
    void* cargs[] = {(void*)&d_bd, (void*)&handle};
    E.kernel("icp_cond", (const void*)k_icp_cond, dim3(1), dim3(256), cargs);
    if (E.err != cudaSuccess) return fail(E.err, "cudaGraphAddKernelNode(cond)");
    G.kernels_prefix = E.count;
    cudaGraphNodeParams cp{};
    cp.type = cudaGraphNodeTypeConditional;
    cp.conditional.handle = handle;
    cp.conditional.type = cudaGraphCondTypeWhile;
    cp.conditional.size = 1;
    cudaGraphNode_t wnode = nullptr;
    e = cudaGraphAddNode(&wnode, G.graph, &E.last, 1, &cp);
    if (e != cudaSuccess) return fail(e, "cudaGraphAddNode(conditional while)");
    Emitter B;
    B.graph = cp.conditional.phGraph_out[0];
    emit_pass_m(B, method, d_bd, gx_search, cap_b, pairs_b, kIcpUnrolled + 1, combos_mask);      // grid shape of passes >= 5
    B.kernel("icp_cond", (const void*)k_icp_cond, dim3(1), dim3(256), cargs);
    if (B.err != cudaSuccess) return fail(B.err, "cudaGraphAddKernelNode(body)");
    G.kernels_body = B.count;
    e = cudaGraphInstantiate(&G.exec, G.graph, 0);
    if (e != cudaSuccess) return fail(e, "cudaGraphInstantiate");
    cache.graphs.push_back(G);
    return &cache.graphs.back();
}

}  // namespace

void icp_graphs_destroy(IcpGraphCache& cache) {
    for (IcpGraph& g : cache.graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        if (g.graph) cudaGraphDestroy(g.graph);
    }
    cache.graphs.clear();
    if (cache.d_bd) cudaFree(cache.d_bd);
    cache.d_bd = nullptr;
}

// combos_mask bit (2*src_wide + tgt_wide) set when some pair of the batch has that record-type combination
const IcpGraph* run_icp(Launcher& L, IcpGraphCache& cache, const BatchDesc& h_bd, BatchDesc* d_bd_batch, int src_cap_max, int combos_mask,
                        bool use_graph) {
    if (h_bd.n_pairs == 0 || L.err != cudaSuccess) return nullptr;
    const int method = h_bd.ip.method;
    if (use_graph && !cache.disabled) {
        if (!cache.d_bd && cudaMalloc(&cache.d_bd, sizeof(BatchDesc)) != cudaSuccess) { cache.disabled = true; cache.error = "cudaMalloc(BatchDesc)"; cudaGetLastError(); }
        const IcpGraph* G = cache.disabled ? nullptr : get_graph(cache, method, combos_mask, bucket_cap(src_cap_max), bucket_pairs(h_bd.n_pairs));
        if (G) {
            // the graph's kernels read the batch from the context's fixed descriptor: stream-ordered update, then one launch
            cudaError_t e = cudaMemcpyAsync(cache.d_bd, &h_bd, sizeof(BatchDesc), cudaMemcpyHostToDevice, L.stream);
            Launcher::Rec r{"icp_graph", nullptr, nullptr};
            if (L.profile) { r.a = L.get_event(); r.b = L.get_event(); cudaEventRecord(r.a, L.stream); }
            if (e == cudaSuccess) e = cudaGraphLaunch(G->exec, L.stream);
            if (L.profile) { cudaEventRecord(r.b, L.stream); L.recs.push_back(r); }
            if (e != cudaSuccess) L.err = e;
            return G;
        }
        fprintf(stderr, "[arvc] device-terminated ICP loop unavailable (%s): enqueuing max_iter + 1 passes instead\n", cache.error.c_str());
    }
    if (cudaMemcpyAsync(d_bd_batch, &h_bd, sizeof(BatchDesc), cudaMemcpyHostToDevice, L.stream) != cudaSuccess) { L.err = cudaGetLastError(); return nullptr; }
    Emitter E;
    E.L = &L;
    const int gx_search = max(1, (src_cap_max + kIcpBlock - 1) / kIcpBlock);
    for (int pass = 0; pass <= h_bd.ip.max_iter; ++pass) emit_pass_m(E, method, d_bd_batch, gx_search, src_cap_max, h_bd.n_pairs, pass, combos_mask);
    return nullptr;
}

}  // namespace arvc
