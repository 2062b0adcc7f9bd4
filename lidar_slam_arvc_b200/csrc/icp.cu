// ICP iteration, fully device-resident: one fused pass kernel per iteration does
//   source transform -> exact nearest neighbour on the target's hash grid (warm-started from the previous
//   pass' match) -> residual/Jacobian (point-to-plane) or Umeyama moments (point-to-point) ->
//   warp + block reduction -> last-block grid reduction -> 6x6 LDLT / 3x3 SVD solve -> update of the
//   cumulative transformation and Open3D's convergence test.
// The host only enqueues max_iter+1 launches; converged pairs turn into no-ops via a device flag.
//
// Reference semantics restated: keyframemanager/keyframe.py:246-252 -> Open3D RegistrationICP,
// GetRegistrationResultAndCorrespondences, TransformationEstimationPointToPlane / PointToPoint.
#include "engine.cuh"

namespace arvc {

namespace {

struct Best { double d2; int idx; int pos; };

constexpr int kLeafMax = 48;   // cells with more points than this are descended octree-style instead of scanned

template <bool TW>
__device__ __forceinline__ void scan_run(const typename RecT<TW>::type* __restrict__ recs, unsigned st, unsigned en, double sx, double sy,
                                         double sz, Best& b) {
    for (unsigned p = st; p < en; ++p) {
        double x, y, z;
        int idx;
        load_rec(recs + p, x, y, z, idx);
        const double d2 = sqdist(sx, sy, sz, x, y, z);
        if (d2 < b.d2 || (d2 == b.d2 && idx < b.idx)) { b.d2 = d2; b.idx = idx; b.pos = (int)p; }
    }
}

__device__ __forceinline__ double box_d2(const GridSpec& g, int cx, int cy, int cz, double cl, double sx, double sy, double sz) {
    const double bx0 = g.ox + cx * cl, by0 = g.oy + cy * cl, bz0 = g.oz + cz * cl;
    const double ddx = fmax(0.0, fmax(bx0 - sx, sx - (bx0 + cl)));
    const double ddy = fmax(0.0, fmax(by0 - sy, sy - (by0 + cl)));
    const double ddz = fmax(0.0, fmax(bz0 - sz, sz - (bz0 + cl)));
    return ddx * ddx + ddy * ddy + ddz * ddz;
}

// Branch-and-bound descent of one big cell (level l > 0): its eight children are the level l-1 cells with
// prefix*8+k, visited nearest octant first and pruned by box distance against the best so far.  Keeps queries
// whose nearest neighbour is metres away (occlusion shadows) from scanning thousands of points linearly.
template <bool TW>
__device__ __noinline__ void descend_cell(const ScanDev& tgt, double sx, double sy, double sz, Best& b, int l, int cx, int cy, int cz) {
    typedef typename RecT<TW>::type Rec;
    const Rec* __restrict__ recs = reinterpret_cast<const Rec*>(tgt.recs);
    const GridSpec g = tgt.grid;
    int fcx[kMortonBits + 1], fcy[kMortonBits + 1], fcz[kMortonBits + 1];
    unsigned char fk[kMortonBits + 1], fnear[kMortonBits + 1];
    int depth = 0;
    auto near_octant = [&](int level, int x, int y, int z) -> unsigned char {
        const double cl = g.c0 * (double)(1 << level), h = 0.5 * cl;
        return (unsigned char)(((sx >= g.ox + x * cl + h) ? 4 : 0) | ((sy >= g.oy + y * cl + h) ? 2 : 0) | ((sz >= g.oz + z * cl + h) ? 1 : 0));
    };
    fcx[0] = cx; fcy[0] = cy; fcz[0] = cz; fk[0] = 0; fnear[0] = near_octant(l, cx, cy, cz);
    while (depth >= 0) {
        if (fk[depth] == 8) { --depth; continue; }
        const unsigned child = fnear[depth] ^ ((0x76534210u >> (4 * fk[depth])) & 7u);   // 0,1,2,4,3,5,6,7 axis flips
        ++fk[depth];
        const int cl_level = l - depth - 1;
        const int ccx = 2 * fcx[depth] + ((child >> 2) & 1), ccy = 2 * fcy[depth] + ((child >> 1) & 1), ccz = 2 * fcz[depth] + (child & 1);
        const double cl = g.c0 * (double)(1 << cl_level);
        if (box_d2(g, ccx, ccy, ccz, cl, sx, sy, sz) > b.d2 * (1.0 + 1e-9) + 1e-12) continue;
        unsigned st, en;
        if (!grid_lookup(tgt.table, tgt.table_mask, cl_level, morton3(ccx, ccy, ccz), st, en)) continue;
        if (cl_level == 0 || en - st <= (unsigned)kLeafMax) {
            scan_run<TW>(recs, st, en, sx, sy, sz, b);
        } else {
            ++depth;
            fcx[depth] = ccx; fcy[depth] = ccy; fcz[depth] = ccz; fk[depth] = 0; fnear[depth] = near_octant(cl_level, ccx, ccy, ccz);
        }
    }
}

// Exact nearest neighbour of (sx,sy,sz) among the target records, under the (d2, cloud index) order.
// On entry `b` holds an exclusive upper bound (candidates must be lexicographically smaller).
template <bool TW>
__device__ __forceinline__ void nn_search(const ScanDev& tgt, double sx, double sy, double sz, Best& b, int level) {
    typedef typename RecT<TW>::type Rec;
    const Rec* __restrict__ recs = reinterpret_cast<const Rec*>(tgt.recs);
    const HashEntry* __restrict__ tab = tgt.table;
    const unsigned mask = tgt.table_mask;
    const GridSpec g = tgt.grid;
    for (int l = level; l <= g.top_level; ++l) {
        const double cl = g.c0 * (double)(1 << l);
        const double br = sqrt(b.d2) * (1.0 + 1e-9) + 1e-12;
        const double r = fmin(br, cl);
        const int x0 = cell_coord(sx - r, g.ox, g.inv_c0) >> l, x1 = cell_coord(sx + r, g.ox, g.inv_c0) >> l;
        const int y0 = cell_coord(sy - r, g.oy, g.inv_c0) >> l, y1 = cell_coord(sy + r, g.oy, g.inv_c0) >> l;
        const int z0 = cell_coord(sz - r, g.oz, g.inv_c0) >> l, z1 = cell_coord(sz + r, g.oz, g.inv_c0) >> l;
        // the cell holding the query first: it usually yields the bound that prunes the others
        const int hx = min(max(cell_coord(sx, g.ox, g.inv_c0) >> l, x0), x1), hy = min(max(cell_coord(sy, g.oy, g.inv_c0) >> l, y0), y1),
                  hz = min(max(cell_coord(sz, g.oz, g.inv_c0) >> l, z0), z1);
        const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, ncell = nx * ny * (z1 - z0 + 1);
        const int home = (hx - x0) + nx * ((hy - y0) + ny * (hz - z0));
        for (int t = 0; t < ncell; ++t) {
            const int c = t == 0 ? home : (t <= home ? t - 1 : t);
            const int cx = x0 + c % nx, cy = y0 + (c / nx) % ny, cz = z0 + c / (nx * ny);
            if (box_d2(g, cx, cy, cz, cl, sx, sy, sz) > b.d2 * (1.0 + 1e-9) + 1e-12) continue;   // conservative slack
            unsigned st, en;
            if (!grid_lookup(tab, mask, l, morton3(cx, cy, cz), st, en)) continue;
            if (l > 0 && en - st > (unsigned)kLeafMax) descend_cell<TW>(tgt, sx, sy, sz, b, l, cx, cy, cz);
            else scan_run<TW>(recs, st, en, sx, sy, sz, b);
        }
        // every point within min(best radius, cl) has been seen: exact as soon as the best lies within cl
        if (sqrt(b.d2) * (1.0 + 1e-9) + 1e-12 <= cl) return;
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

template <int METHOD, bool SW, bool TW>
__global__ void __launch_bounds__(kIcpBlock) k_icp_pass(const PairDev* __restrict__ pairs, IcpParams ip, int pass) {
    typedef typename RecT<SW>::type SRec;
    typedef typename RecT<TW>::type TRec;
    __shared__ double s_T[16];
    __shared__ double s_red[kIcpBlock / 32][kNS];
    __shared__ int s_last;

    const PairDev& pr = pairs[blockIdx.y];
    const ScanDev& src = *pr.src;
    const ScanDev& tgt = *pr.tgt;
    if ((src.wide != 0) != SW || (tgt.wide != 0) != TW) return;
    PairState* st = pr.state;
    if (st->done) return;
    const int n = src.counts[CNT_NPTS];
    const int nblk = max(1, (n + kIcpBlock - 1) / kIcpBlock);
    if ((int)blockIdx.x >= nblk) return;
    if (threadIdx.x < 16) s_T[threadIdx.x] = st->T[threadIdx.x];
    __syncthreads();

    double acc[kNS];
#pragma unroll
    for (int k = 0; k < kNS; ++k) acc[k] = 0.0;

    const int i = blockIdx.x * kIcpBlock + threadIdx.x;
    if (i < n) {
        double px, py, pz;
        int sidx;
        load_rec(reinterpret_cast<const SRec*>(src.recs) + i, px, py, pz, sidx);
        const double sx = s_T[0] * px + s_T[1] * py + s_T[2] * pz + s_T[3];
        const double sy = s_T[4] * px + s_T[5] * py + s_T[6] * pz + s_T[7];
        const double sz = s_T[8] * px + s_T[9] * py + s_T[10] * pz + s_T[11];
        Best b;
        b.d2 = ip.max_d2; b.idx = -1; b.pos = -1;
        const GridSpec& g = tgt.grid;
        int level = min(g.cold_level, g.top_level);
        const bool finite = sx == sx && sy == sy && sz == sz;
        if (pass > 0) {
            const int pv = pr.prev[i];
            if (pv >= 0) {
                double x, y, z;
                int idx;
                load_rec(reinterpret_cast<const TRec*>(tgt.recs) + pv, x, y, z, idx);
                const double d2 = sqdist(sx, sy, sz, x, y, z);
                if (d2 < b.d2) {
                    // the previous match bounds the search ball; accept it unless something is strictly better
                    b.d2 = d2; b.idx = idx; b.pos = pv;
                    const double br = sqrt(d2) * (1.0 + 1e-9) + 1e-12;
                    level = 0;
                    while (level < g.top_level && g.c0 * (double)(1 << level) < br) ++level;
                }
            }
        }
        if (finite && ip.max_d2 > 0) nn_search<TW>(tgt, sx, sy, sz, b, level);
        pr.prev[i] = b.pos;
        if (pr.corr_trace) pr.corr_trace[(size_t)pass * src.cap + sidx] = b.pos >= 0 ? b.idx : -1;
        if (b.pos >= 0) {
            double tx, ty, tz;
            int tidx;
            load_rec(reinterpret_cast<const TRec*>(tgt.recs) + b.pos, tx, ty, tz, tidx);
            acc[27] = b.d2;
            acc[28] = 1.0;
            if (METHOD == 1) {
                const double4 nv = reinterpret_cast<const double4*>(tgt.normals)[b.pos];
                const double r = (sx - tx) * nv.x + (sy - ty) * nv.y + (sz - tz) * nv.z;
                const double J[6] = {sy * nv.z - sz * nv.y, sz * nv.x - sx * nv.z, sx * nv.y - sy * nv.x, nv.x, nv.y, nv.z};
                int t = 0;
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int c = a; c < 6; ++c) acc[t++] = J[a] * J[c];
#pragma unroll
                for (int a = 0; a < 6; ++a) acc[21 + a] = J[a] * r;
            } else {
                acc[0] = sx; acc[1] = sy; acc[2] = sz; acc[3] = tx; acc[4] = ty; acc[5] = tz;
                acc[6] = tx * sx; acc[7] = tx * sy; acc[8] = tx * sz;
                acc[9] = ty * sx; acc[10] = ty * sy; acc[11] = ty * sz;
                acc[12] = tz * sx; acc[13] = tz * sy; acc[14] = tz * sz;
            }
        }
    }
    // warp -> block reduction in a fixed order (bit-reproducible run to run)
    const int lane = lane_id(), w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kNS; ++k) {
        if (METHOD == 0 && k >= 15 && k < 27) continue;
        const double v = warp_sum(acc[k]);
        if (lane == 0) s_red[w][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < kNS) {
        double v = 0;
        if (!(METHOD == 0 && threadIdx.x >= 15 && threadIdx.x < 27)) {
#pragma unroll
            for (int ww = 0; ww < kIcpBlock / 32; ++ww) v += s_red[ww][threadIdx.x];
        }
        pr.partials[(size_t)blockIdx.x * kSumStride + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&st->ticket, 1u);
        s_last = (t == (unsigned)nblk - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;

    // ---- last block of this pair: grid reduction, solve, convergence test -------------------------------
    __shared__ double s_sum[kSumStride];
    if (threadIdx.x < kNS) {
        double v = 0;
        for (int bb = 0; bb < nblk; ++bb) v += __ldcg(pr.partials + (size_t)bb * kSumStride + threadIdx.x);
        s_sum[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double K = s_sum[28];
        const double fitness = (K > 0 && n > 0) ? K / (double)n : 0.0;
        const double rmse = K > 0 ? sqrt(s_sum[27] / K) : 0.0;
        if (pr.state_trace) {
            double* tr = pr.state_trace + (size_t)pass * 18;
            for (int k = 0; k < 16; ++k) tr[k] = s_T[k];
            tr[16] = fitness; tr[17] = rmse;
        }
        bool stop = pass >= ip.max_iter;
        if (pass > 0 && fabs(st->fitness - fitness) < ip.rel_fitness && fabs(st->rmse - rmse) < ip.rel_rmse) stop = true;
        st->fitness = fitness;
        st->rmse = rmse;
        st->ncorr = (int)K;
        st->passes = pass + 1;
        for (int k = 0; k < kNS; ++k) st->sums[k] = s_sum[k];
        if (stop) {
            st->done = 1;
        } else {
            double upd[16], Tn[16];
            solve_update(METHOD, s_sum, upd);
            mat4_mul(upd, s_T, Tn);
            for (int k = 0; k < 16; ++k) st->T[k] = Tn[k];
            st->updates = st->updates + 1;
        }
        st->err = src.counts[CNT_ERR] | tgt.counts[CNT_ERR];
        st->ticket = 0u;
        __threadfence();
    }
}

template <int METHOD>
static void launch_combos(Launcher& L, const PairDev* d_pairs, dim3 grid, const IcpParams& ip, int pass, int combos_mask) {
    if (combos_mask & 1) L.launch("icp_pass", k_icp_pass<METHOD, false, false>, grid, dim3(kIcpBlock), d_pairs, ip, pass);
    if (combos_mask & 2) L.launch("icp_pass", k_icp_pass<METHOD, false, true>, grid, dim3(kIcpBlock), d_pairs, ip, pass);
    if (combos_mask & 4) L.launch("icp_pass", k_icp_pass<METHOD, true, false>, grid, dim3(kIcpBlock), d_pairs, ip, pass);
    if (combos_mask & 8) L.launch("icp_pass", k_icp_pass<METHOD, true, true>, grid, dim3(kIcpBlock), d_pairs, ip, pass);
}

// combos_mask bit (2*src_wide + tgt_wide) set when some pair of the batch has that record-type combination
void run_icp(Launcher& L, const PairDev* d_pairs, int n_pairs, int src_cap_max, const IcpParams& ip, int combos_mask) {
    if (n_pairs == 0) return;
    const dim3 grid(max(1, (src_cap_max + kIcpBlock - 1) / kIcpBlock), n_pairs);
    for (int pass = 0; pass <= ip.max_iter; ++pass) {
        if (ip.method == 1) launch_combos<1>(L, d_pairs, grid, ip, pass, combos_mask);
        else launch_combos<0>(L, d_pairs, grid, ip, pass, combos_mask);
    }
}

}  // namespace arvc
