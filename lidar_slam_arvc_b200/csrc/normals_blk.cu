// Normal estimation, block-cooperative neighbour gathering (narrow = float32 records).
//
// Reference semantics restated: keyframemanager/keyframe.py:160-162
//   pointcloud_filtered.estimate_normals(KDTreeSearchParamHybrid(radius=0.3, max_nn=300))
// -> Open3D EstimatePerPointCovariances (SearchHybrid: k nearest, then d2 < r2) + ComputeCovariance.
//
// One block serves kNbG Morton-consecutive points, i.e. spatial neighbours whose search balls overlap almost entirely:
//   1. the block looks up ONCE the grid cells its points' balls touch (one hash probe per thread), estimates the local
//      density from their population and - where more than max_nn points crowd the radius - shrinks the search to a
//      trial radius that should still hold max_nn points (verified per point afterwards, never assumed);
//   2. the cells' records (contiguous runs of the Morton-sorted array) are streamed once, coalesced, through a
//      distance-to-bounding-box filter and compacted into a shared-memory tile: only records that can lie inside
//      some point's ball survive (a third to a quarter of the cells' content);
//   3. every warp then serves its points one after the other against the tile: lanes = candidates, float32
//      screening with exact float64 decisions at the boundaries (same rules as normals.cu), a 512-bucket d2 histogram
//      to select the max_nn nearest when more are inside the radius, centred float64 moment sums.
// Points the block cannot serve exactly from its tile (tile overflow, box too large after a jump of the space-filling
// curve, trial radius too small, pathological boundary bucket) go to the scan's fallback list, which the per-point
// kernel of normals.cu processes afterwards - results do not depend on which path a point takes.
#include "engine.cuh"

namespace arvc {

namespace {

#ifndef ARVC_NBQ
#define ARVC_NBQ 4
#endif
#ifndef ARVC_NBTILE
#define ARVC_NBTILE 2048
#endif
#ifndef ARVC_NBWARPS
#define ARVC_NBWARPS 8
#endif
constexpr int kNbWarps = ARVC_NBWARPS;
constexpr int kNbThreads = kNbWarps * 32;
constexpr int kNbQ = ARVC_NBQ;               // points per warp
constexpr int kNbG = kNbWarps * kNbQ;        // points per block (<= 32: one warp loads them)
constexpr int kNbTile = ARVC_NBTILE;         // records of the shared candidate tile
constexpr int kNbBins = 512;
constexpr int kNbCand = 64;                  // boundary candidates ranked exactly per point
constexpr int kNbCells = kNbThreads;         // one cell lookup per thread and round
static_assert(kNbG <= 32, "the points of a block are loaded by one warp");

struct NbWarp {
    unsigned hist[kNbBins / 2];              // 16-bit counters, two per word (a tile holds <= 2048 records)
    double cand_d2[kNbCand];
    int cand_idx[kNbCand];
    int cand_pos[kNbCand];
    int ncand;
    int pad[3];
};

struct NbShared {
    float4 tile[kNbTile + 64];
    float4 q[kNbG];
    uint2 runs[kNbCells];
    int roff[kNbCells + 1];
    int wcnt[2][kNbWarps];
    int wlen[kNbWarps];
    float bb[8];                             // min xyz, max xyz of the block's points
    int fb_base;
    NbWarp ws[kNbWarps];
};

__device__ __forceinline__ bool key_less_nb(double d2a, int ia, double d2b, int ib) { return d2a < d2b || (d2a == d2b && ia < ib); }

// distance^2 from a point to the axis-aligned box [lo, hi] in float32 (used as a conservative pre-filter only)
__device__ __forceinline__ float box_dist2f(float x, float y, float z, const float* bb) {
    const float dx = fmaxf(0.f, fmaxf(bb[0] - x, x - bb[3]));
    const float dy = fmaxf(0.f, fmaxf(bb[1] - y, y - bb[4]));
    const float dz = fmaxf(0.f, fmaxf(bb[2] - z, z - bb[5]));
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

// Block-wide compaction of the cells found by a lookup round: runs[] in thread order, roff[] = prefix sums of their
// lengths.  Returns the number of runs; roff[nruns] = total records.
__device__ __forceinline__ int compact_runs(NbShared& S, bool valid, unsigned st, unsigned en) {
    const int lane = lane_id(), w = threadIdx.x >> 5;
    const unsigned vm = __ballot_sync(kFull, valid);
    const int len = valid ? (int)(en - st) : 0;
    int inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) { S.wcnt[0][w] = __popc(vm); S.wlen[w] = inc; }
    __syncthreads();
    int slot = 0, off = 0, nruns = 0, total = 0;
#pragma unroll
    for (int k = 0; k < kNbWarps; ++k) {
        const int c = S.wcnt[0][k], l = S.wlen[k];
        if (k < w) { slot += c; off += l; }
        nruns += c; total += l;
    }
    if (valid) {
        const int sl = slot + __popc(vm & ((1u << lane) - 1u));
        S.runs[sl] = make_uint2(st, en);
        S.roff[sl] = off + inc - len;
    }
    if (threadIdx.x == 0) S.roff[nruns] = total;
    __syncthreads();
    return nruns;
}

}  // namespace

#ifndef ARVC_NBOCC
#define ARVC_NBOCC 4
#endif
template <bool TAP>      // TAP: record the neighbour indices every normal used (parity tap, separate instantiation)
__global__ void __launch_bounds__(kNbThreads, ARVC_NBOCC) k_normals_blk(const ScanDev* __restrict__ scans, NormalParams np) {
    extern __shared__ __align__(16) unsigned char nb_smem[];
    NbShared& S = *reinterpret_cast<NbShared*>(nb_smem);
    const ScanDev& s = scans[blockIdx.y];
    if (s.wide != 0) return;
    const int n = s.counts[CNT_NPTS];
    const int p0 = blockIdx.x * kNbG;
    if (p0 >= n) return;
    const int nq = min(kNbG, n - p0);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const float4* __restrict__ recs4 = reinterpret_cast<const float4*>(s.recs);
    const GridSpec g = s.grid;
    const int K = np.max_nn;

    // every point of the block joins the scan's fallback list (served by the per-point kernel)
    auto block_fallback = [&]() {
        if (tid == 0) {
            S.fb_base = atomicAdd(&s.counts[CNT_NFB], nq);
            atomicAdd(&s.counts[CNT_NFB_BLOCKS], 1);
        }
        __syncthreads();
        if (tid < nq) s.fb_list[S.fb_base + tid] = p0 + tid;
    };

    // ---- phase 0: the block's points and their bounding box
    if (w == 0) {
        const float4 v = __ldg(recs4 + p0 + (lane < nq ? lane : 0));
        if (lane < kNbG) S.q[lane] = v;
        float lx = v.x, ly = v.y, lz = v.z, hx = v.x, hy = v.y, hz = v.z;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lx = fminf(lx, __shfl_xor_sync(kFull, lx, o)); hx = fmaxf(hx, __shfl_xor_sync(kFull, hx, o));
            ly = fminf(ly, __shfl_xor_sync(kFull, ly, o)); hy = fmaxf(hy, __shfl_xor_sync(kFull, hy, o));
            lz = fminf(lz, __shfl_xor_sync(kFull, lz, o)); hz = fmaxf(hz, __shfl_xor_sync(kFull, hz, o));
        }
        if (lane == 0) { S.bb[0] = lx; S.bb[1] = ly; S.bb[2] = lz; S.bb[3] = hx; S.bb[4] = hy; S.bb[5] = hz; }
    }
    __syncthreads();
    const double blx = (double)S.bb[0], bly = (double)S.bb[1], blz = (double)S.bb[2];
    const double bhx = (double)S.bb[3], bhy = (double)S.bb[4], bhz = (double)S.bb[5];

    // ---- phase 1: cells of the full radius (level L: edge >= radius) around the bounding box
    const int L = np.level;
    const double cl = g.c0 * (double)(1 << L);
    int nruns = 0;
    int cnx, cny, cnz;
    {
        const double rinf = np.radius * (1.0 + 1e-9) + 1e-12;
        const int x0 = cell_coord(blx - rinf, g.ox, g.inv_c0) >> L, x1 = cell_coord(bhx + rinf, g.ox, g.inv_c0) >> L;
        const int y0 = cell_coord(bly - rinf, g.oy, g.inv_c0) >> L, y1 = cell_coord(bhy + rinf, g.oy, g.inv_c0) >> L;
        const int z0 = cell_coord(blz - rinf, g.oz, g.inv_c0) >> L, z1 = cell_coord(bhz + rinf, g.oz, g.inv_c0) >> L;
        cnx = x1 - x0 + 1; cny = y1 - y0 + 1; cnz = z1 - z0 + 1;
        const int ncell = cnx * cny * cnz;
        if (ncell > kNbCells) {                                 // a jump of the space-filling curve inside the block
            if ((np.debug & 4) && tid == 0) atomicAdd(&s.counts[15], 1);
            block_fallback();
            return;
        }
        bool valid = false;
        unsigned st = 0, en = 0;
        if (tid < ncell) {
            const CellDecoder dec(cnx, cny);
            int cx, cy, cz;
            dec(tid, cx, cy, cz);
            valid = grid_lookup(s.table, s.table_mask, L, morton3(x0 + cx, y0 + cy, z0 + cz), st, en);
        }
        nruns = compact_runs(S, valid, st, en);
    }
    int U = S.roff[nruns];

    // ---- density -> trial radius.  Surface model: the U points of the looked-up cells lie on a patch of area
    // A = cl^2 * (largest face of the cell box); the ball that holds K of them has radius sqrt(K A / (pi U)).
    // Three regimes: rt well below the radius -> search only the trial ball (`trial`); rt around the radius -> more than K
    // points are expected inside the radius, go straight to the selection (`crowded`); else one sweep usually settles it.
    double rq = np.radius;
    bool trial = false, crowded = false;
    if (U > K && L > 0) {
        const double A = cl * cl * (double)max(cnx * cny, max(cnx * cnz, cny * cnz));
        const double rt = 1.15 * sqrt((double)K * A / (3.141592653589793 * (double)U));      // 15 % safety on the model
        if (rt < 0.9 * np.radius) { trial = true; rq = rt; }
        else if (rt < (double)np.crowded_ratio * np.radius) crowded = true;
    }
    __syncthreads();      // every thread has read roff[nruns] before the table is rebuilt
    if (trial) {
        // ---- phase 2: finer cells around the box inflated by the trial radius (finest level with <= 256 cells)
        const double rinf = rq * (1.0 + 1e-9) + 1e-12;
        int Lf = 0, x0 = 0, y0 = 0, z0 = 0, fnx = 0, fny = 0, ncell = 0;
        for (; Lf < L; ++Lf) {
            x0 = cell_coord(blx - rinf, g.ox, g.inv_c0) >> Lf;
            y0 = cell_coord(bly - rinf, g.oy, g.inv_c0) >> Lf;
            z0 = cell_coord(blz - rinf, g.oz, g.inv_c0) >> Lf;
            fnx = (cell_coord(bhx + rinf, g.ox, g.inv_c0) >> Lf) - x0 + 1;
            fny = (cell_coord(bhy + rinf, g.oy, g.inv_c0) >> Lf) - y0 + 1;
            ncell = fnx * fny * ((cell_coord(bhz + rinf, g.oz, g.inv_c0) >> Lf) - z0 + 1);
            if (ncell <= kNbCells) break;
        }
        if (Lf < L) {      // otherwise the level-L runs of phase 1 stay (they cover the smaller ball as well)
            const double cf = g.c0 * (double)(1 << Lf);
            bool valid = false;
            unsigned st = 0, en = 0;
            if (tid < ncell) {
                const CellDecoder dec(fnx, fny);
                int cx, cy, cz;
                dec(tid, cx, cy, cz);
                cx += x0; cy += y0; cz += z0;
                // cell box against the points' bounding box: farther than the trial radius -> no point's ball touches it
                const double c0x = g.ox + cx * cf, c0y = g.oy + cy * cf, c0z = g.oz + cz * cf;
                const double ddx = fmax(0.0, fmax(c0x - bhx, blx - (c0x + cf)));
                const double ddy = fmax(0.0, fmax(c0y - bhy, bly - (c0y + cf)));
                const double ddz = fmax(0.0, fmax(c0z - bhz, blz - (c0z + cf)));
                if (ddx * ddx + ddy * ddy + ddz * ddz <= rq * rq * (1.0 + 1e-9) + 1e-12)
                    valid = grid_lookup(s.table, s.table_mask, Lf, morton3(cx, cy, cz), st, en);
            }
            nruns = compact_runs(S, valid, st, en);
            U = S.roff[nruns];
        }
    }

    // ---- phase 3: stream the runs once, keep what lies within rq of the points' bounding box
    const double rq2 = rq * rq;
    int ntile = 0;
    {
        const float keep2 = (float)(rq2 * (1.0 + 1e-5) + 1e-9);      // float32 box distance, generous slack: a pre-filter only
        int cur = 0, buf = 0;
        for (int base = 0; base < U; base += 2 * kNbThreads) {
            bool k0 = false, k1 = false;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            const int f0 = base + tid, f1 = base + kNbThreads + tid;
            if (f0 < U) {
                while (f0 >= S.roff[cur + 1]) ++cur;
                v0 = __ldg(recs4 + S.runs[cur].x + (unsigned)(f0 - S.roff[cur]));
            }
            if (f1 < U) {
                while (f1 >= S.roff[cur + 1]) ++cur;
                v1 = __ldg(recs4 + S.runs[cur].x + (unsigned)(f1 - S.roff[cur]));
            }
            if (f0 < U) k0 = box_dist2f(v0.x, v0.y, v0.z, S.bb) <= keep2;
            if (f1 < U) k1 = box_dist2f(v1.x, v1.y, v1.z, S.bb) <= keep2;
            const unsigned m0 = __ballot_sync(kFull, k0), m1 = __ballot_sync(kFull, k1);
            if (lane == 0) S.wcnt[buf][w] = __popc(m0) | (__popc(m1) << 16);
            __syncthreads();
            int before0 = 0, before1 = 0, tot0 = 0, tot1 = 0;
#pragma unroll
            for (int k = 0; k < kNbWarps; ++k) {
                const int c = S.wcnt[buf][k], c0 = c & 0xffff, c1 = c >> 16;
                if (k < w) { before0 += c0; before1 += c1; }
                tot0 += c0; tot1 += c1;
            }
            const unsigned below = (1u << lane) - 1u;
            if (k0) { const int pos = ntile + before0 + __popc(m0 & below); if (pos < kNbTile) S.tile[pos] = v0; }
            if (k1) { const int pos = ntile + tot0 + before1 + __popc(m1 & below); if (pos < kNbTile) S.tile[pos] = v1; }
            ntile += tot0 + tot1;
            buf ^= 1;      // the next round writes the other counter set: one barrier per round
        }
    }
    if (ntile > kNbTile) { __syncthreads(); block_fallback(); return; }
    // pad to a multiple of 64 with records no ball can contain: the sweeps then run two full chunks per round
    const int ntile_pad = (ntile + 63) & ~63;
    if (tid < ntile_pad - ntile) S.tile[ntile + tid] = make_float4(3.0e18f, 3.0e18f, 3.0e18f, 0.f);
    __syncthreads();
    if (tid == 0) {
        if (trial) atomicAdd(&s.counts[CNT_NTRIAL_BLOCKS], 1);
        if (np.debug & 4) { atomicAdd(&s.counts[9], ntile); atomicAdd(&s.counts[14], U); }
    }

    // ---- phase 4: every warp serves its points against the tile
    NbWarp& W = S.ws[w];
    const double bin_scale = (double)kNbBins / rq2;
    float rq2_lo = (float)(rq2 * (1.0 - 2e-6)), rq2_hi = (float)(rq2 * (1.0 + 2e-6)), bsf = (float)bin_scale;
    asm volatile("" : "+f"(rq2_lo), "+f"(rq2_hi), "+f"(bsf));      // keep the screening constants in registers
#pragma unroll 1
    for (int qi = 0; qi < kNbQ; ++qi) {
        const int ql = w * kNbQ + qi;
        if (ql >= nq) break;
        const float4 qv = S.q[ql];
        const float qxf = qv.x, qyf = qv.y, qzf = qv.z;
        const double qx = (double)qxf, qy = (double)qyf, qz = (double)qzf;
        double sx = 0, sy = 0, sz = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
        int cnt = 0;
        const int p = p0 + ql;
        const bool tap = TAP && s.tap_idx != nullptr;
        auto tap_reset = [&]() {
            if (tap) { __syncwarp(); if (lane == 0) s.tap_cnt[p] = 0; __syncwarp(); }
        };
        tap_reset();
        auto accumulate = [&](const float4& v) {
            if (tap) {
                const int slot = atomicAdd(&s.tap_cnt[p], 1);
                if (slot < s.tap_stride) s.tap_idx[(size_t)p * s.tap_stride + slot] = __float_as_int(v.w);
            }
            const double ux = (double)v.x - qx, uy = (double)v.y - qy, uz = (double)v.z - qz;      // exact differences
            sx += ux; sy += uy; sz += uz;
            sxx = fma(ux, ux, sxx); sxy = fma(ux, uy, sxy); sxz = fma(ux, uz, sxz);
            syy = fma(uy, uy, syy); syz = fma(uy, uz, syz); szz = fma(uz, uz, szz);
            ++cnt;
        };
        auto dist2f = [&](const float4& v) -> float {
            const float dx = qxf - v.x, dy = qyf - v.y, dz = qzf - v.z;
            return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        };
        auto exact = [&](const float4& v) -> double { return sqdist(qx, qy, qz, (double)v.x, (double)v.y, (double)v.z); };
        // everything inside the radius (decided exactly), no selection
        auto take_all = [&]() {
            for (int j = lane; j < ntile_pad; j += 64) {
                const float4 a = S.tile[j], b = S.tile[j + 32];
                const float da = dist2f(a), db = dist2f(b);
                if (da < rq2_lo || (da <= rq2_hi && exact(a) < rq2)) accumulate(a);
                if (db < rq2_lo || (db <= rq2_hi && exact(b) < rq2)) accumulate(b);
            }
        };
        bool settled = false, failed = false;
        if (!trial && !crowded) {
            // sparse / moderate neighbourhoods mostly hold <= K points inside the radius: take them all in one sweep and
            // count; only when more than K turn up is the work discarded and the selection run
            take_all();
            if (ntile <= K || warp_sum(cnt) <= K) settled = true;
            else { sx = sy = sz = sxx = sxy = sxz = syy = syz = szz = 0; cnt = 0; tap_reset(); }
        }
        if (!settled) {
            // ---- sweep A: d2 histogram of the in-radius candidates.  float32 bucket coordinate (error < 1e-3 buckets):
            // membership in the radius is decided exactly, the bucket itself may be off by one - the boundary that is
            // ranked exactly below is therefore three buckets wide
            for (int b = lane; b < kNbBins / 2; b += 32) W.hist[b] = 0u;
            __syncwarp();
            auto count_one = [&](const float4& v) {
                const float u = dist2f(v) * bsf;
                int b;
                if (u < (float)kNbBins - 2e-3f) b = (int)u;
                else if (u > (float)kNbBins + 2e-3f) return;
                else {
                    const double d2 = exact(v);
                    if (!(d2 < rq2)) return;
                    b = min(kNbBins - 1, (int)(d2 * bin_scale));
                }
                atomicAdd(&W.hist[b >> 1], (b & 1) ? 0x10000u : 1u);
            };
            for (int j = lane; j < ntile_pad; j += 64) {
                const float4 a = S.tile[j], b = S.tile[j + 32];
                count_one(a);
                count_one(b);
            }
            __syncwarp();
            constexpr int per = kNbBins / 32;      // lane owns `per` consecutive buckets
            int local = 0;
#pragma unroll
            for (int b = 0; b < per / 2; ++b) {
                const unsigned hw = W.hist[lane * (per / 2) + b];
                local += (int)(hw & 0xffffu) + (int)(hw >> 16);
            }
            int inc = local;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, inc, o);
                if (lane >= o) inc += t;
            }
            const int n_in = __shfl_sync(kFull, inc, 31);
            int bstar = kNbBins;      // buckets < bstar are taken whole; bstar == kNbBins: everything inside the radius
            if (trial && n_in < K) {
                failed = true;        // the trial radius holds fewer than K points: the per-point kernel searches the full radius
            } else if (n_in > K) {
                int before = inc - local, myb = -1;
                if (before < K && inc >= K) {      // the K-th nearest falls into one of this lane's buckets
                    for (int b = 0; b < per; ++b) {
                        const int h = (int)((W.hist[(lane * per + b) >> 1] >> (16 * (b & 1))) & 0xffffu);
                        if (myb < 0 && before + h >= K) myb = lane * per + b;
                        before += h;
                    }
                }
                const unsigned who = __ballot_sync(kFull, myb >= 0);
                bstar = __shfl_sync(kFull, myb, __ffs(who) - 1);
            }
            if (!failed) {
                // ---- sweep B: buckets certainly below the boundary are accumulated, the boundary is collected for exact ranking
                const int blo = bstar == kNbBins ? kNbBins : bstar - 1, bhi = bstar + 1;
                const float u_lo = (float)blo - 2e-3f, u_hi = (float)(bstar == kNbBins ? kNbBins : bhi + 1) + 2e-3f;
                if (lane == 0) W.ncand = 0;
                __syncwarp();
                auto take_one = [&](const float4& v, int j) {
                    const float u = dist2f(v) * bsf;
                    if (u < u_lo) {
                        accumulate(v);
                    } else if (u <= u_hi) {
                        const double d2 = exact(v);
                        if (d2 < rq2) {
                            const int b = min(kNbBins - 1, (int)(d2 * bin_scale));
                            if (b < blo) {
                                accumulate(v);
                            } else if (b <= bhi) {
                                const int slot = atomicAdd(&W.ncand, 1);
                                if (slot < kNbCand) { W.cand_d2[slot] = d2; W.cand_idx[slot] = __float_as_int(v.w); W.cand_pos[slot] = j; }
                            }
                        }
                    }
                };
                for (int j = lane; j < ntile_pad; j += 64) {
                    const float4 a = S.tile[j], b = S.tile[j + 32];
                    take_one(a, j);
                    take_one(b, j + 32);
                }
                __syncwarp();
                if (bstar != kNbBins) {
                    const int ncand = W.ncand;
                    const int need = K - warp_sum(cnt);      // still missing once the buckets below the boundary are taken whole
                    if (ncand > kNbCand) {
                        failed = true;                       // pathological boundary (many duplicates): per-point kernel
                    } else {
                        // exact rank inside the boundary; the `need` smallest (d2, index) keys join the neighbourhood
                        for (int a = lane; a < ncand; a += 32) {
                            const double d2a = W.cand_d2[a];
                            const int ia = W.cand_idx[a];
                            int rank = 0;
                            for (int b = 0; b < ncand; ++b) rank += key_less_nb(W.cand_d2[b], W.cand_idx[b], d2a, ia) ? 1 : 0;
                            if (rank < need) accumulate(S.tile[W.cand_pos[a]]);
                        }
                    }
                }
            }
        }
        if (failed) {
            tap_reset();
            if (lane == 0) {
                s.fb_list[atomicAdd(&s.counts[CNT_NFB], 1)] = p;
                atomicAdd(&s.counts[CNT_NFB_POINTS], 1);
            }
            continue;
        }
        cnt = warp_sum(cnt);
        if ((np.debug & 4) && lane == 0) {
            atomicAdd(&s.counts[10], cnt);
            atomicAdd(&s.counts[(trial || crowded) ? 13 : (settled ? 11 : 12)], 1);
        }
        // the nine sums reduced together: every butterfly step halves the values a lane still carries (16 -> 8 -> 4 -> 2 ->
        // 1, then one plain step), 16 double shuffles instead of 45; lane 4k ends up with sum k
        {
            double v[16] = {sx, sy, sz, sxx, sxy, sxz, syy, syz, szz, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int half = 8, o = 16; half >= 1; half >>= 1, o >>= 1) {
                const bool up = (lane & o) != 0;
#pragma unroll
                for (int k = 0; k < half; ++k) {
                    const double mine = up ? v[k + half] : v[k], theirs = up ? v[k] : v[k + half];
                    v[k] = mine + __shfl_xor_sync(kFull, theirs, o);
                }
            }
            const double tot = v[0] + __shfl_xor_sync(kFull, v[0], 1);      // lanes 2m and 2m + 1 hold the two halves of sum idx(m)
            // after the steps with o = 16, 8, 4, 2 lane bits 4, 3, 2, 1 select the sum: index = bit4 * 8 + bit3 * 4 + bit2 * 2 + bit1
            const int which = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            if ((lane & 1) == 0 && which < 9) s.moments[10 * (size_t)p + which] = tot;
        }
        if (lane == 0) {
            s.moments[10 * (size_t)p + 9] = (double)cnt;
            s.nn_count[p] = cnt;
        }
        __syncwarp();
    }
}

void launch_normals_blk(Launcher& L, const ScanDev* d_scans, int n_scans, int cap_max, const NormalParams& np, bool tap) {
    static unsigned long long attr_set = 0;      // per device: the attribute belongs to the device's copy of the kernel
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((attr_set >> (dev & 63)) & 1ull)) {
        cudaFuncSetAttribute(k_normals_blk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NbShared));
        cudaFuncSetAttribute(k_normals_blk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NbShared));
        attr_set |= 1ull << (dev & 63);
    }
    const dim3 grid((cap_max + kNbG - 1) / kNbG, n_scans);
    if (tap) L.launch_smem("normals", k_normals_blk<true>, grid, dim3(kNbThreads), sizeof(NbShared), d_scans, np);
    else L.launch_smem("normals", k_normals_blk<false>, grid, dim3(kNbThreads), sizeof(NbShared), d_scans, np);
}

}  // namespace arvc
