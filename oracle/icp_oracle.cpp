// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under lidar_slam_arvc_b200/ may link, import or call
// this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// CPU float64 restatement of the arithmetic behind the reference's ICP scan-matching hot path.
// The reference (/root/reference, pure Python) delegates that arithmetic to Open3D, an UNPINNED pip
// dependency (requirements.txt:2, bare `open3d`) whose sources are absent from /root/reference and
// which cannot be installed here.  PARITY IS THEREFORE UNPINNED against Open3D itself: the reference
// ships no tests, golden vectors or fixtures for this path.  This file restates
// the published algorithms of Open3D's legacy CPU pipeline (>= 0.13 semantics) at the reference's
// own call sites:
//   keyframemanager/keyframe.py:74-94    filter_radius_height     -> orc_filter_radius_height
//   keyframemanager/keyframe.py:111,151,159  voxel_down_sample    -> orc_voxel_down_sample
//   keyframemanager/keyframe.py:160-162  estimate_normals(Hybrid) -> orc_knn_hybrid, orc_estimate_normals
//   keyframemanager/keyframe.py:246-252  registration_icp         -> orc_icp, orc_correspondences
// and is cross-checked by an independent numpy + scipy.cKDTree implementation (oracle/numpy_ref.py)
// and known-answer tests (tests/test_oracle_*.py).
//
// Conventions fixed here because Open3D leaves them implementation-defined (documented in DESIGN.md):
//   * exact distance ties are broken towards the LOWEST point index (nanoflann: first visited);
//   * voxel_down_sample output is ordered by voxel key, lexicographic (ix, iy, iz)
//     (Open3D: std::unordered_map iteration order);
//   * all reductions run in ascending source-index order (Open3D: OpenMP-schedule dependent).
//
// Build: see oracle/Makefile  (g++ -O2 -fopenmp -ffp-contract=off, no -march flags: no FMA contraction,
// matching a generic x86-64 Open3D wheel).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------------
// KD-tree, leaf size 15 (nanoflann's leaf_max_size in Open3D's KDTreeFlann).  Results are the exact
// k nearest neighbours under the (d2, index) lexicographic order, so they do not depend on the tree
// shape; only the tie rule is a convention.
// ---------------------------------------------------------------------------------------------
inline double sqdist3(const double* a, const double* b) {
    // nanoflann L2_Adaptor::evalMetric for dim 3: ((dx*dx) + dy*dy) + dz*dz, each op rounded.
    const double dx = a[0] - b[0];
    const double dy = a[1] - b[1];
    const double dz = a[2] - b[2];
    double r = dx * dx;
    r = r + dy * dy;
    r = r + dz * dz;
    return r;
}

struct KDTree {
    struct Node {
        int lo, hi;         // range in perm
        int left, right;    // children (-1 for leaf)
        double bmin[3], bmax[3];
    };
    const double* pts = nullptr;
    int n = 0;
    std::vector<int> perm;
    std::vector<Node> nodes;
    static constexpr int kLeaf = 15;

    void build(const double* p, int count) {
        pts = p;
        n = count;
        perm.resize(n);
        for (int i = 0; i < n; ++i) perm[i] = i;
        nodes.clear();
        nodes.reserve(n / 4 + 16);
        if (n > 0) build_rec(0, n);
    }

    int build_rec(int lo, int hi) {
        Node nd;
        nd.lo = lo; nd.hi = hi; nd.left = nd.right = -1;
        for (int d = 0; d < 3; ++d) { nd.bmin[d] = std::numeric_limits<double>::infinity(); nd.bmax[d] = -nd.bmin[d]; }
        for (int i = lo; i < hi; ++i) {
            const double* q = pts + 3 * (size_t)perm[i];
            for (int d = 0; d < 3; ++d) { nd.bmin[d] = std::min(nd.bmin[d], q[d]); nd.bmax[d] = std::max(nd.bmax[d], q[d]); }
        }
        const int id = (int)nodes.size();
        nodes.push_back(nd);
        if (hi - lo > kLeaf) {
            int dim = 0;
            double ext = nd.bmax[0] - nd.bmin[0];
            for (int d = 1; d < 3; ++d) if (nd.bmax[d] - nd.bmin[d] > ext) { ext = nd.bmax[d] - nd.bmin[d]; dim = d; }
            if (ext > 0.0) {
                const int mid = lo + (hi - lo) / 2;
                const double* P = pts;
                std::nth_element(perm.begin() + lo, perm.begin() + mid, perm.begin() + hi, [P, dim](int a, int b) {
                    const double va = P[3 * (size_t)a + dim], vb = P[3 * (size_t)b + dim];
                    return va < vb || (va == vb && a < b);
                });
                const int l = build_rec(lo, mid);
                const int r = build_rec(mid, hi);
                nodes[id].left = l;
                nodes[id].right = r;
            }
        }
        return id;
    }

    inline double box_d2(const Node& nd, const double* q) const {
        double acc = 0.0;
        for (int d = 0; d < 3; ++d) {
            double diff = 0.0;
            if (q[d] < nd.bmin[d]) diff = q[d] - nd.bmin[d];
            else if (q[d] > nd.bmax[d]) diff = q[d] - nd.bmax[d];
            acc = acc + diff * diff;
        }
        return acc;
    }

    // ---- 1-NN
    void nn1_rec(int id, const double* q, double& bd, int& bi) const {
        const Node& nd = nodes[id];
        if (nd.left < 0) {
            for (int i = nd.lo; i < nd.hi; ++i) {
                const int j = perm[i];
                const double d = sqdist3(q, pts + 3 * (size_t)j);
                if (d < bd || (d == bd && j < bi)) { bd = d; bi = j; }
            }
            return;
        }
        const double dl = box_d2(nodes[nd.left], q), dr = box_d2(nodes[nd.right], q);
        const int first = dl <= dr ? nd.left : nd.right, second = dl <= dr ? nd.right : nd.left;
        const double dfirst = std::min(dl, dr), dsecond = std::max(dl, dr);
        if (dfirst <= bd) nn1_rec(first, q, bd, bi);
        if (dsecond <= bd) nn1_rec(second, q, bd, bi);
    }
    // returns index or -1 (NaN query / empty tree)
    int nn1(const double* q, double& d2) const {
        double bd = std::numeric_limits<double>::infinity();
        int bi = std::numeric_limits<int>::max();
        if (n > 0 && q[0] == q[0] && q[1] == q[1] && q[2] == q[2]) nn1_rec(0, q, bd, bi);
        if (bi == std::numeric_limits<int>::max()) { d2 = 0; return -1; }
        d2 = bd;
        return bi;
    }

    // ---- k-NN with a bounded max-heap on (d2, idx)
    typedef std::pair<double, int> DI;
    void knn_rec(int id, const double* q, int k, std::vector<DI>& heap) const {
        const Node& nd = nodes[id];
        if (nd.left < 0) {
            for (int i = nd.lo; i < nd.hi; ++i) {
                const int j = perm[i];
                const DI c(sqdist3(q, pts + 3 * (size_t)j), j);
                if ((int)heap.size() < k) { heap.push_back(c); std::push_heap(heap.begin(), heap.end()); }
                else if (c < heap.front()) { std::pop_heap(heap.begin(), heap.end()); heap.back() = c; std::push_heap(heap.begin(), heap.end()); }
            }
            return;
        }
        const double dl = box_d2(nodes[nd.left], q), dr = box_d2(nodes[nd.right], q);
        const int first = dl <= dr ? nd.left : nd.right, second = dl <= dr ? nd.right : nd.left;
        const double dfirst = std::min(dl, dr), dsecond = std::max(dl, dr);
        if ((int)heap.size() < k || dfirst <= heap.front().first) knn_rec(first, q, k, heap);
        if ((int)heap.size() < k || dsecond <= heap.front().first) knn_rec(second, q, k, heap);
    }
    // sorted ascending by (d2, idx)
    void knn(const double* q, int k, std::vector<DI>& out) const {
        out.clear();
        if (n == 0 || k <= 0) return;
        knn_rec(0, q, k, out);
        std::sort_heap(out.begin(), out.end());
    }
};

// ---------------------------------------------------------------------------------------------
// 3x3 symmetric eigen: smallest-eigenvalue eigenvector, analytic (Geometric Tools
// "RobustEigenSymmetric3x3" scheme as used by Open3D's FastEigen3x3 / ComputeEigenvector0/1).
// ---------------------------------------------------------------------------------------------
struct V3 { double x, y, z; };
inline V3 cross(const V3& a, const V3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 scale(const V3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }

// A is symmetric, stored a00 a01 a02 a11 a12 a22
struct Sym3 { double a00, a01, a02, a11, a12, a22; };

inline V3 divv(const V3& a, double s) { return {a.x / s, a.y / s, a.z / s}; }   // Eigen `v / s`: true division

V3 eigvec0(const Sym3& A, double ev) {
    const V3 r0{A.a00 - ev, A.a01, A.a02};
    const V3 r1{A.a01, A.a11 - ev, A.a12};
    const V3 r2{A.a02, A.a12, A.a22 - ev};
    const V3 c01 = cross(r0, r1), c02 = cross(r0, r2), c12 = cross(r1, r2);
    const double d0 = dot(c01, c01), d1 = dot(c02, c02), d2 = dot(c12, c12);
    double dmax = d0;
    int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) { imax = 2; }
    if (imax == 0) return divv(c01, std::sqrt(d0));
    if (imax == 1) return divv(c02, std::sqrt(d1));
    return divv(c12, std::sqrt(d2));
}

V3 eigvec1(const Sym3& A, const V3& e0, double ev1) {
    V3 U, V;
    if (std::fabs(e0.x) > std::fabs(e0.y)) {
        const double inv = 1.0 / std::sqrt(e0.x * e0.x + e0.z * e0.z);
        U = {-e0.z * inv, 0.0, e0.x * inv};
    } else {
        const double inv = 1.0 / std::sqrt(e0.y * e0.y + e0.z * e0.z);
        U = {0.0, e0.z * inv, -e0.y * inv};
    }
    V = cross(e0, U);
    const V3 AU{A.a00 * U.x + A.a01 * U.y + A.a02 * U.z, A.a01 * U.x + A.a11 * U.y + A.a12 * U.z,
                A.a02 * U.x + A.a12 * U.y + A.a22 * U.z};
    const V3 AV{A.a00 * V.x + A.a01 * V.y + A.a02 * V.z, A.a01 * V.x + A.a11 * V.y + A.a12 * V.z,
                A.a02 * V.x + A.a12 * V.y + A.a22 * V.z};
    double m00 = U.x * AU.x + U.y * AU.y + U.z * AU.z - ev1;
    double m01 = U.x * AV.x + U.y * AV.y + U.z * AV.z;
    double m11 = V.x * AV.x + V.y * AV.y + V.z * AV.z - ev1;
    const double a00 = std::fabs(m00), a01 = std::fabs(m01), a11 = std::fabs(m11);
    if (a00 >= a11) {
        if (std::max(a00, a01) > 0) {
            if (a00 >= a01) { m01 /= m00; m00 = 1 / std::sqrt(1 + m01 * m01); m01 *= m00; }
            else            { m00 /= m01; m01 = 1 / std::sqrt(1 + m00 * m00); m00 *= m01; }
            return {m01 * U.x - m00 * V.x, m01 * U.y - m00 * V.y, m01 * U.z - m00 * V.z};
        }
        return U;
    } else {
        if (std::max(a11, a01) > 0) {
            if (a11 >= a01) { m01 /= m11; m11 = 1 / std::sqrt(1 + m01 * m01); m01 *= m11; }
            else            { m11 /= m01; m01 = 1 / std::sqrt(1 + m11 * m11); m11 *= m01; }
            return {m11 * U.x - m01 * V.x, m11 * U.y - m01 * V.y, m11 * U.z - m01 * V.z};
        }
        return U;
    }
}

// covariance c[9] row-major (symmetric) -> normal
V3 fast_eigen3x3(const double* c) {
    double mx = c[0];
    for (int i = 1; i < 9; ++i) mx = std::max(mx, c[i]);   // Eigen maxCoeff(): max, not max-abs
    if (mx == 0) return {0, 0, 0};
    Sym3 A{c[0] / mx, c[1] / mx, c[2] / mx, c[4] / mx, c[5] / mx, c[8] / mx};
    const double norm = A.a01 * A.a01 + A.a02 * A.a02 + A.a12 * A.a12;
    if (norm > 0) {
        const double q = (A.a00 + A.a11 + A.a22) / 3;
        const double b00 = A.a00 - q, b11 = A.a11 - q, b22 = A.a22 - q;
        const double p = std::sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2) / 6);
        const double c00 = b11 * b22 - A.a12 * A.a12;
        const double c01 = A.a01 * b22 - A.a12 * A.a02;
        const double c02 = A.a01 * A.a12 - b11 * A.a02;
        const double det = (b00 * c00 - A.a01 * c01 + A.a02 * c02) / (p * p * p);
        double half_det = det * 0.5;
        half_det = std::min(std::max(half_det, -1.0), 1.0);
        const double angle = std::acos(half_det) / 3.0;
        const double two_thirds_pi = 2.09439510239319549;
        const double beta2 = std::cos(angle) * 2;
        const double beta0 = std::cos(angle + two_thirds_pi) * 2;
        const double beta1 = -(beta0 + beta2);
        const double e0 = q + p * beta0, e1 = q + p * beta1, e2 = q + p * beta2;
        if (half_det >= 0) {
            const V3 v2 = eigvec0(A, e2);
            if (e2 < e0 && e2 < e1) return v2;
            const V3 v1 = eigvec1(A, v2, e1);
            if (e1 < e0 && e1 < e2) return v1;
            return cross(v1, v2);
        } else {
            const V3 v0 = eigvec0(A, e0);
            if (e0 < e1 && e0 < e2) return v0;
            const V3 v1 = eigvec1(A, v0, e1);
            if (e1 < e0 && e1 < e2) return v1;
            return cross(v0, v1);
        }
    } else {
        // diagonal matrix (A *= max_coeff in Open3D restores the scale; the comparisons are scale-free)
        if (c[0] < c[4] && c[0] < c[8]) return {1, 0, 0};
        if (c[4] < c[0] && c[4] < c[8]) return {0, 1, 0};
        return {0, 0, 1};
    }
}

// Open3D utility::ComputeCovariance: single-pass cumulants over the listed indices, raw coordinates.
void covariance_from_indices(const double* pts, const int* idx, int m, double* cov) {
    double cu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < m; ++t) {
        const double* p = pts + 3 * (size_t)idx[t];
        cu[0] += p[0]; cu[1] += p[1]; cu[2] += p[2];
        cu[3] += p[0] * p[0]; cu[4] += p[0] * p[1]; cu[5] += p[0] * p[2];
        cu[6] += p[1] * p[1]; cu[7] += p[1] * p[2]; cu[8] += p[2] * p[2];
    }
    for (int i = 0; i < 9; ++i) cu[i] /= (double)m;
    cov[0] = cu[3] - cu[0] * cu[0];
    cov[4] = cu[6] - cu[1] * cu[1];
    cov[8] = cu[8] - cu[2] * cu[2];
    cov[1] = cov[3] = cu[4] - cu[0] * cu[1];
    cov[2] = cov[6] = cu[5] - cu[0] * cu[2];
    cov[5] = cov[7] = cu[7] - cu[1] * cu[2];
}

// ---------------------------------------------------------------------------------------------
// Small dense linear algebra
// ---------------------------------------------------------------------------------------------
inline void mat4_mul(const double* A, const double* B, double* C) {   // C = A*B, row-major, C may not alias
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j];
            C[4 * i + j] = s;
        }
}

// Open3D TransformPoints: homogeneous product then divide by w.
inline void transform_point(const double* T, double* p) {
    const double x = p[0], y = p[1], z = p[2];
    const double nx = T[0] * x + T[1] * y + T[2] * z + T[3];
    const double ny = T[4] * x + T[5] * y + T[6] * z + T[7];
    const double nz = T[8] * x + T[9] * y + T[10] * z + T[11];
    const double nw = T[12] * x + T[13] * y + T[14] * z + T[15];
    p[0] = nx / nw; p[1] = ny / nw; p[2] = nz / nw;
}

// Eigen-style LDLT with symmetric (largest |diagonal|) pivoting, n = 6.  Solves A x = b.
void ldlt_solve6(const double* Ain, const double* b, double* x) {
    const int n = 6;
    double A[36];
    std::memcpy(A, Ain, sizeof(A));
    int piv[6];
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = std::fabs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i) if (std::fabs(A[i * n + i]) > best) { best = std::fabs(A[i * n + i]); p = i; }
        piv[k] = p;
        if (p != k) {   // symmetric swap of rows/cols k and p (full storage)
            for (int j = 0; j < n; ++j) std::swap(A[k * n + j], A[p * n + j]);
            for (int i = 0; i < n; ++i) std::swap(A[i * n + k], A[i * n + p]);
        }
        const double d = A[k * n + k];
        if (d != 0.0) {
            for (int i = k + 1; i < n; ++i) A[i * n + k] /= d;             // L(i,k)
            for (int i = k + 1; i < n; ++i)
                for (int j = k + 1; j <= i; ++j) {
                    A[i * n + j] -= A[i * n + k] * d * A[j * n + k];
                    A[j * n + i] = A[i * n + j];
                }
        }
    }
    double y[6];
    for (int i = 0; i < n; ++i) y[i] = b[i];
    for (int k = 0; k < n; ++k) std::swap(y[k], y[piv[k]]);               // y = P b
    for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) y[i] -= A[i * n + j] * y[j];   // L y' = y
    const double tol = std::numeric_limits<double>::min();
    for (int i = 0; i < n; ++i) y[i] = std::fabs(A[i * n + i]) > tol ? y[i] / A[i * n + i] : 0.0;   // D
    for (int i = n - 1; i >= 0; --i) for (int j = i + 1; j < n; ++j) y[i] -= A[j * n + i] * y[j];   // L^T
    for (int k = n - 1; k >= 0; --k) std::swap(y[k], y[piv[k]]);          // x = P^T y
    for (int i = 0; i < n; ++i) x[i] = y[i];
}

// Open3D TransformVector6dToMatrix4d: R = Rz(x2) * Ry(x1) * Rx(x0), t = (x3,x4,x5)
void vec6_to_mat4(const double* v, double* T) {
    const double ca = std::cos(v[0]), sa = std::sin(v[0]);
    const double cb = std::cos(v[1]), sb = std::sin(v[1]);
    const double cg = std::cos(v[2]), sg = std::sin(v[2]);
    T[0] = cg * cb; T[1] = cg * sb * sa - sg * ca; T[2] = cg * sb * ca + sg * sa; T[3] = v[3];
    T[4] = sg * cb; T[5] = sg * sb * sa + cg * ca; T[6] = sg * sb * ca - cg * sa; T[7] = v[4];
    T[8] = -sb;     T[9] = cb * sa;                T[10] = cb * ca;               T[11] = v[5];
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

// 3x3 SVD by one-sided (Hestenes) Jacobi.  A = U diag(s) V^T, s descending, U,V orthogonal (full).
void svd3(const double* Ain, double* U, double* s, double* V) {
    double A[9];
    std::memcpy(A, Ain, sizeof(A));
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 3; ++i) { alpha += A[3 * i + p] * A[3 * i + p]; beta += A[3 * i + q] * A[3 * i + q]; gamma += A[3 * i + p] * A[3 * i + q]; }
                if (gamma == 0.0 || std::fabs(gamma) <= 1e-17 * std::sqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
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
    for (int j = 0; j < 3; ++j) nrm[j] = std::sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
    int ord[3] = {0, 1, 2};
    std::sort(ord, ord + 3, [&](int a, int b) { return nrm[a] > nrm[b]; });
    double Vs[9], Us[9];
    for (int j = 0; j < 3; ++j) {
        const int o = ord[j];
        s[j] = nrm[o];
        for (int i = 0; i < 3; ++i) { Vs[3 * i + j] = V[3 * i + o]; Us[3 * i + j] = nrm[o] > 0 ? A[3 * i + o] / nrm[o] : 0.0; }
    }
    // complete U for (numerically) zero singular values so that it stays orthogonal
    const double tiny = s[0] * 1e-14;
    auto col = [&](double* M, int j) { return V3{M[j], M[3 + j], M[6 + j]}; };
    auto setcol = [&](double* M, int j, const V3& v) { M[j] = v.x; M[3 + j] = v.y; M[6 + j] = v.z; };
    if (s[0] <= 0) { for (int i = 0; i < 9; ++i) Us[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    else {
        if (s[1] <= tiny) {   // rank 1: pick any unit vector orthogonal to u0
            const V3 u0 = col(Us, 0);
            V3 a = std::fabs(u0.x) < 0.9 ? V3{1, 0, 0} : V3{0, 1, 0};
            V3 u1 = cross(u0, a);
            u1 = scale(u1, 1.0 / std::sqrt(dot(u1, u1)));
            setcol(Us, 1, u1);
        }
        if (s[2] <= tiny) setcol(Us, 2, cross(col(Us, 0), col(Us, 1)));
    }
    std::memcpy(U, Us, sizeof(Us));
    std::memcpy(V, Vs, sizeof(Vs));
}

inline double det3(const double* M) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// One correspondence pass: Open3D GetRegistrationResultAndCorrespondences on already transformed
// source points.  corr[i] = target index or -1.  error2 summed in ascending i.
void correspondence_pass(const KDTree& tree, const double* src, int ns, double max_d, int* corr, double* d2out,
                         double& fitness, double& rmse, int& K) {
    const double r2 = max_d * max_d;
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < ns; ++i) {
        double d2;
        const int j = tree.nn1(src + 3 * (size_t)i, d2);
        if (j >= 0 && d2 < r2) { corr[i] = j; d2out[i] = d2; }
        else { corr[i] = -1; d2out[i] = 0; }
    }
    double e2 = 0;
    K = 0;
    for (int i = 0; i < ns; ++i) if (corr[i] >= 0) { e2 += d2out[i]; ++K; }
    if (K == 0) { fitness = 0; rmse = 0; }
    else { fitness = (double)K / (double)ns; rmse = std::sqrt(e2 / (double)K); }
}

// TransformationEstimationPointToPlane::ComputeTransformation
void p2plane_update(const double* src, const double* tgt, const double* nrm, const int* corr, int ns, int K, double* upd,
                    double* jtj_out, double* jtr_out) {
    double JTJ[36], JTr[6];
    std::memset(JTJ, 0, sizeof(JTJ));
    std::memset(JTr, 0, sizeof(JTr));
    for (int i = 0; i < ns; ++i) {
        if (corr[i] < 0) continue;
        const double* s = src + 3 * (size_t)i;
        const double* t = tgt + 3 * (size_t)corr[i];
        const double* n = nrm + 3 * (size_t)corr[i];
        const double r = (s[0] - t[0]) * n[0] + (s[1] - t[1]) * n[1] + (s[2] - t[2]) * n[2];
        const double J[6] = {s[1] * n[2] - s[2] * n[1], s[2] * n[0] - s[0] * n[2], s[0] * n[1] - s[1] * n[0], n[0], n[1], n[2]};
        for (int a = 0; a < 6; ++a) {
            for (int b = 0; b < 6; ++b) JTJ[6 * a + b] += J[a] * J[b];
            JTr[a] += J[a] * r;
        }
    }
    if (jtj_out) std::memcpy(jtj_out, JTJ, sizeof(JTJ));
    if (jtr_out) std::memcpy(jtr_out, JTr, sizeof(JTr));
    for (int i = 0; i < 16; ++i) upd[i] = (i % 5 == 0) ? 1.0 : 0.0;
    if (K == 0) return;
    double nb[6], x[6];
    for (int a = 0; a < 6; ++a) nb[a] = -JTr[a];
    ldlt_solve6(JTJ, nb, x);
    vec6_to_mat4(x, upd);
}

// TransformationEstimationPointToPoint (with_scaling = false): Eigen::umeyama, two-pass demeaned.
void p2p_update(const double* src, const double* tgt, const int* corr, int ns, int K, double* upd) {
    for (int i = 0; i < 16; ++i) upd[i] = (i % 5 == 0) ? 1.0 : 0.0;
    if (K == 0) return;
    double ms[3] = {0, 0, 0}, mt[3] = {0, 0, 0};
    for (int i = 0; i < ns; ++i) {
        if (corr[i] < 0) continue;
        const double* s = src + 3 * (size_t)i;
        const double* t = tgt + 3 * (size_t)corr[i];
        for (int d = 0; d < 3; ++d) { ms[d] += s[d]; mt[d] += t[d]; }
    }
    const double inv = 1.0 / (double)K;
    for (int d = 0; d < 3; ++d) { ms[d] *= inv; mt[d] *= inv; }
    double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // sigma = 1/K * sum (t - mt)(s - ms)^T
    for (int i = 0; i < ns; ++i) {
        if (corr[i] < 0) continue;
        const double* s = src + 3 * (size_t)i;
        const double* t = tgt + 3 * (size_t)corr[i];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) S[3 * a + b] += (t[a] - mt[a]) * (s[b] - ms[b]);
    }
    for (int i = 0; i < 9; ++i) S[i] *= inv;
    double U[9], sv[3], V[9];
    svd3(S, U, sv, V);
    const double sgn = (det3(U) * det3(V) < 0) ? -1.0 : 1.0;
    double R[9];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) R[3 * a + b] = U[3 * a + 0] * V[3 * b + 0] + U[3 * a + 1] * V[3 * b + 1] + sgn * U[3 * a + 2] * V[3 * b + 2];
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) upd[4 * a + b] = R[3 * a + b];
        upd[4 * a + 3] = mt[a] - (R[3 * a + 0] * ms[0] + R[3 * a + 1] * ms[1] + R[3 * a + 2] * ms[2]);
    }
}

bool is_identity4(const double* T) {   // Eigen isIdentity(prec = 1e-12)
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const double v = T[4 * i + j];
            if (i == j) { if (std::fabs(v - 1.0) > 1e-12) return false; }       // isApprox(1, prec)
            else { if (std::fabs(v) > 1e-12) return false; }                   // isMuchSmallerThan(1, prec)
        }
    return true;
}

}  // namespace

extern "C" {

int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// keyframe.py:88-93: r2 = x**2 + y**2; keep (r2 < max_r2) & (r2 > min_r2) & (z > min_h) & (z < max_h),
// strict, order preserved.  min_r2/max_r2 are the already squared radii (Python computes `radius ** 2`).
int orc_filter_radius_height(const double* pts, int n, double min_r2, double max_r2, double min_h, double max_h, int* keep) {
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const double x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
        const double r2 = x * x + y * y;
        if (r2 < max_r2 && r2 > min_r2 && z > min_h && z < max_h) keep[m++] = i;
    }
    return m;
}

// Open3D PointCloud::VoxelDownSample: origin = min_bound - 0.5*v; key = floor((p - origin)/v) per axis;
// output = mean of the points of each key (sum in ascending point index, divided by count).
// Output order: ascending (ix, iy, iz).  Returns M, or -1 on bad arguments.
int orc_voxel_down_sample(const double* pts, int n, double v, double* out_pts, int* out_keys, int* out_counts) {
    if (v <= 0.0) return -1;
    if (n == 0) return 0;
    double mn[3] = {pts[0], pts[1], pts[2]};
    for (int i = 1; i < n; ++i) for (int d = 0; d < 3; ++d) mn[d] = std::min(mn[d], pts[3 * (size_t)i + d]);
    double org[3];
    for (int d = 0; d < 3; ++d) org[d] = mn[d] - v * 0.5;
    struct Rec { int k[3]; int i; };
    std::vector<Rec> recs(n);
    for (int i = 0; i < n; ++i) {
        recs[i].i = i;
        for (int d = 0; d < 3; ++d) recs[i].k[d] = (int)std::floor((pts[3 * (size_t)i + d] - org[d]) / v);
    }
    std::stable_sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) {
        if (a.k[0] != b.k[0]) return a.k[0] < b.k[0];
        if (a.k[1] != b.k[1]) return a.k[1] < b.k[1];
        return a.k[2] < b.k[2];
    });
    int m = 0;
    for (int a = 0; a < n;) {
        int b = a;
        double s[3] = {0, 0, 0};
        while (b < n && recs[b].k[0] == recs[a].k[0] && recs[b].k[1] == recs[a].k[1] && recs[b].k[2] == recs[a].k[2]) {
            for (int d = 0; d < 3; ++d) s[d] += pts[3 * (size_t)recs[b].i + d];
            ++b;
        }
        const double cnt = (double)(b - a);
        for (int d = 0; d < 3; ++d) { out_pts[3 * (size_t)m + d] = s[d] / cnt; if (out_keys) out_keys[3 * (size_t)m + d] = recs[a].k[d]; }
        if (out_counts) out_counts[m] = b - a;
        ++m;
        a = b;
    }
    return m;
}

// KDTreeFlann::SearchHybrid(query, radius, max_nn): knnSearch(max_nn) sorted ascending, then the
// prefix with d2 < radius^2 (lower_bound => strict).  out_idx/out_d2 are [nq, max_nn]; unused = -1 / 0.
int orc_knn_hybrid(const double* pts, int n, const double* queries, int nq, double radius, int max_nn, int* out_idx, double* out_d2,
                   int* out_cnt) {
    KDTree tree;
    tree.build(pts, n);
    const double r2 = radius * radius;
#pragma omp parallel
    {
        std::vector<KDTree::DI> res;
        res.reserve(max_nn + 1);
#pragma omp for schedule(dynamic, 64)
        for (int q = 0; q < nq; ++q) {
            tree.knn(queries + 3 * (size_t)q, max_nn, res);
            int k = 0;
            while (k < (int)res.size() && res[k].first < r2) ++k;
            out_cnt[q] = k;
            for (int t = 0; t < max_nn; ++t) {
                if (out_idx) out_idx[(size_t)q * max_nn + t] = t < k ? res[t].second : -1;
                if (out_d2) out_d2[(size_t)q * max_nn + t] = t < k ? res[t].first : 0.0;
            }
        }
    }
    return 0;
}

// PointCloud::EstimateNormals(KDTreeSearchParamHybrid(radius, max_nn), fast_normal_computation = true)
// on a cloud without prior normals.  cov (optional) is [n,9]; nn_count (optional) [n].
int orc_estimate_normals(const double* pts, int n, double radius, int max_nn, double* normals, double* cov_out, int* nn_count) {
    KDTree tree;
    tree.build(pts, n);
    const double r2 = radius * radius;
#pragma omp parallel
    {
        std::vector<KDTree::DI> res;
        std::vector<int> idx;
        res.reserve(max_nn + 1);
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < n; ++i) {
            tree.knn(pts + 3 * (size_t)i, max_nn, res);
            int k = 0;
            while (k < (int)res.size() && res[k].first < r2) ++k;
            double cov[9];
            if (k >= 3) {
                idx.resize(k);
                for (int t = 0; t < k; ++t) idx[t] = res[t].second;
                covariance_from_indices(pts, idx.data(), k, cov);
            } else {
                for (int t = 0; t < 9; ++t) cov[t] = (t % 4 == 0) ? 1.0 : 0.0;
            }
            V3 nv = fast_eigen3x3(cov);
            if (std::sqrt(nv.x * nv.x + nv.y * nv.y + nv.z * nv.z) == 0.0) nv = {0.0, 0.0, 1.0};
            normals[3 * (size_t)i] = nv.x; normals[3 * (size_t)i + 1] = nv.y; normals[3 * (size_t)i + 2] = nv.z;
            if (cov_out) std::memcpy(cov_out + 9 * (size_t)i, cov, sizeof(cov));
            if (nn_count) nn_count[i] = k;
        }
    }
    return 0;
}

// normal from a given covariance (unit tests of the eigen-solver)
void orc_normal_from_covariance(const double* cov9, double* n3) {
    V3 nv = fast_eigen3x3(cov9);
    if (std::sqrt(nv.x * nv.x + nv.y * nv.y + nv.z * nv.z) == 0.0) nv = {0.0, 0.0, 1.0};
    n3[0] = nv.x; n3[1] = nv.y; n3[2] = nv.z;
}

// One correspondence pass of source points transformed by T (row-major 4x4) against target.
int orc_correspondences(const double* src, int ns, const double* tgt, int nt, const double* T16, double max_d, int* corr, double* d2,
                        double* fitness, double* rmse) {
    KDTree tree;
    tree.build(tgt, nt);
    std::vector<double> s(src, src + 3 * (size_t)ns);
    for (int i = 0; i < ns; ++i) transform_point(T16, s.data() + 3 * (size_t)i);
    int K;
    correspondence_pass(tree, s.data(), ns, max_d, corr, d2, *fitness, *rmse, K);
    return K;
}

// keyframe.py:399-400 KeyFrame.transform -> Open3D PointCloud::Transform (map building, keyframemanager.py:154-184)
void orc_transform_points(const double* pts, int n, const double* T16, double* out) {
    for (int i = 0; i < n; ++i) {
        double p[3] = {pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2]};
        transform_point(T16, p);
        out[3 * (size_t)i] = p[0]; out[3 * (size_t)i + 1] = p[1]; out[3 * (size_t)i + 2] = p[2];
    }
}

// keyframe.py:417-436 calculate_plane -> Open3D segment_plane(distance_threshold, ransac_n=3, num_iterations): RANSAC over
// the points with z < max_z.  Open3D draws its samples from an unseeded generator (irreproducible by design); the
// convention fixed here and in csrc/plane.cu: sample t of hypothesis h is splitmix64(seed, h, t) % n (rejecting points
// above max_z and duplicates, 64 tries), the hypothesis with most inliers wins, lowest h on ties, unit normal.
static unsigned long long plane_hash(unsigned long long seed, unsigned long long a, unsigned long long b) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (a + 1) + 0xBF58476D1CE4E5B9ull * (b + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static bool plane_hypothesis(const double* pts, int n, double max_z, unsigned long long seed, int h, double* pl) {
    int pick[3], got = 0;
    for (int t = 0; t < 64 && got < 3; ++t) {
        const int j = (int)(plane_hash(seed, (unsigned long long)h, (unsigned long long)t) % (unsigned long long)n);
        if (!(pts[3 * (size_t)j + 2] < max_z)) continue;
        bool dup = false;
        for (int k = 0; k < got; ++k) dup |= pick[k] == j;
        if (!dup) pick[got++] = j;
    }
    if (got < 3) return false;
    const double *p0 = pts + 3 * (size_t)pick[0], *p1 = pts + 3 * (size_t)pick[1], *p2 = pts + 3 * (size_t)pick[2];
    const double ux = p1[0] - p0[0], uy = p1[1] - p0[1], uz = p1[2] - p0[2];
    const double vx = p2[0] - p0[0], vy = p2[1] - p0[1], vz = p2[2] - p0[2];
    const double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
    const double len = std::sqrt(nx * nx + ny * ny + nz * nz);
    if (!(len > 1e-12)) return false;
    pl[0] = nx / len; pl[1] = ny / len; pl[2] = nz / len;
    pl[3] = -(pl[0] * p0[0] + pl[1] * p0[1] + pl[2] * p0[2]);
    return true;
}
// Open3D PointCloud::SegmentPlane, last step: the final inliers of the best hypothesis (|plane . (x, y, z, 1)| < thr) and
// GetPlaneFromPoints on them (centroid, centred second moments, the largest 2x2-determinant cross product, normalised).
// The summation order is the convention shared with the CUDA path (plane.cu k_plane_refit): 1024 strided partial sums,
// xor-butterfly over the 32 lanes of every warp, then over the 32 warp sums.
static double butterfly1024(std::vector<double> v) {
    for (int w = 0; w < 32; ++w)
        for (int o = 16; o > 0; o >>= 1) {
            double t[32];
            for (int l = 0; l < 32; ++l) t[l] = v[32 * w + l] + v[32 * w + (l ^ o)];
            for (int l = 0; l < 32; ++l) v[32 * w + l] = t[l];
        }
    double u[32];
    for (int w = 0; w < 32; ++w) u[w] = v[32 * w];
    for (int o = 16; o > 0; o >>= 1) {
        double t[32];
        for (int l = 0; l < 32; ++l) t[l] = u[l] + u[l ^ o];
        for (int l = 0; l < 32; ++l) u[l] = t[l];
    }
    return u[0];
}
static void plane_refit(const double* pts, int n, double max_z, double thr, double* pl) {
    std::vector<double> sx(1024, 0.0), sy(1024, 0.0), sz(1024, 0.0), sc(1024, 0.0);
    auto inlier = [&](int i) {
        const double x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
        return z < max_z && std::fabs(pl[0] * x + pl[1] * y + pl[2] * z + pl[3]) < thr;
    };
    for (int i = 0; i < n; ++i)
        if (inlier(i)) { const int t = i & 1023; sx[t] += pts[3 * (size_t)i]; sy[t] += pts[3 * (size_t)i + 1]; sz[t] += pts[3 * (size_t)i + 2]; sc[t] += 1.0; }
    const double cnt = butterfly1024(sc);
    const double cx = butterfly1024(sx) / cnt, cy = butterfly1024(sy) / cnt, cz = butterfly1024(sz) / cnt;
    std::vector<double> xx(1024, 0.0), xy(1024, 0.0), xz(1024, 0.0), yy(1024, 0.0), yz(1024, 0.0), zz(1024, 0.0);
    for (int i = 0; i < n; ++i)
        if (inlier(i)) {
            const int t = i & 1023;
            const double r0 = pts[3 * (size_t)i] - cx, r1 = pts[3 * (size_t)i + 1] - cy, r2 = pts[3 * (size_t)i + 2] - cz;
            xx[t] += r0 * r0; xy[t] += r0 * r1; xz[t] += r0 * r2; yy[t] += r1 * r1; yz[t] += r1 * r2; zz[t] += r2 * r2;
        }
    const double XX = butterfly1024(xx), XY = butterfly1024(xy), XZ = butterfly1024(xz), YY = butterfly1024(yy), YZ = butterfly1024(yz), ZZ = butterfly1024(zz);
    const double det_x = YY * ZZ - YZ * YZ, det_y = XX * ZZ - XZ * XZ, det_z = XX * YY - XY * XY;
    double a, b, c;
    if (det_x > det_y && det_x > det_z) { a = det_x; b = XZ * YZ - XY * ZZ; c = XY * YZ - XZ * YY; }
    else if (det_y > det_z) { a = XZ * YZ - XY * ZZ; b = det_y; c = XY * XZ - YZ * XX; }
    else { a = XY * YZ - XZ * YY; b = XY * XZ - YZ * XX; c = det_z; }
    const double norm = std::sqrt(a * a + b * b + c * c);
    if (norm == 0.0) { pl[0] = pl[1] = pl[2] = pl[3] = 0.0; return; }
    a /= norm; b /= norm; c /= norm;
    pl[0] = a; pl[1] = b; pl[2] = c;
    pl[3] = -(a * cx + b * cy + c * cz);
}
int orc_fit_plane(const double* pts, int n, double max_z, double thr, int iterations, unsigned long long seed, double* plane4) {
    int best_cnt = 0, best_h = -1;
    std::vector<int> score(iterations, 0);
    #pragma omp parallel for schedule(dynamic, 8)
    for (int h = 0; h < iterations; ++h) {
        double pl[4];
        if (n < 3 || !plane_hypothesis(pts, n, max_z, seed, h, pl)) continue;
        int c = 0;
        for (int i = 0; i < n; ++i) {
            const double x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
            if (z < max_z && std::fabs(pl[0] * x + pl[1] * y + pl[2] * z + pl[3]) < thr) ++c;
        }
        score[h] = c;
    }
    for (int h = 0; h < iterations; ++h) if (score[h] > best_cnt) { best_cnt = score[h]; best_h = h; }
    plane4[0] = plane4[1] = plane4[2] = plane4[3] = 0;
    if (best_h >= 0) plane_hypothesis(pts, n, max_z, seed, best_h, plane4);
    if (best_cnt >= 3) plane_refit(pts, n, max_z, thr, plane4);
    return best_cnt;
}

void orc_ldlt_solve6(const double* A, const double* b, double* x) { ldlt_solve6(A, b, x); }
void orc_vec6_to_mat4(const double* v, double* T) { vec6_to_mat4(v, T); }
void orc_svd3(const double* A, double* U, double* s, double* V) { svd3(A, U, s, V); }

// registration_icp(source, target, max_corr, init, estimation, ICPConvergenceCriteria(rel_fit, rel_rmse, max_iter))
// method: 0 = point-to-point (Umeyama), 1 = point-to-plane (needs tgt_normals).
// trace_T (optional) [(max_iter+1),16]: cumulative transformation used by pass p; trace_fit/trace_rmse [(max_iter+1)].
// corr_out (optional) [ns]: correspondence set of the final pass.  Returns number of passes executed (>=1) or <0.
int orc_icp(const double* src, int ns, const double* tgt, int nt, const double* tgt_normals, const double* init16, int method,
            double max_corr, double rel_fit, double rel_rmse, int max_iter, double* out_T16, double* out_fitness, double* out_rmse,
            int* out_updates, int* out_ncorr, int* corr_out, double* trace_T, double* trace_fit, double* trace_rmse) {
    if (method == 1 && !tgt_normals) return -2;
    double T[16];
    std::memcpy(T, init16, sizeof(T));
    KDTree tree;
    tree.build(tgt, nt);
    std::vector<double> pcd(src, src + 3 * (size_t)ns);
    if (!is_identity4(T)) for (int i = 0; i < ns; ++i) transform_point(T, pcd.data() + 3 * (size_t)i);
    std::vector<int> corr(ns > 0 ? ns : 1);
    std::vector<double> d2(ns > 0 ? ns : 1);
    double fit = 0, rmse = 0;
    int K = 0, pass = 0, updates = 0;
    if (max_corr > 0.0) correspondence_pass(tree, pcd.data(), ns, max_corr, corr.data(), d2.data(), fit, rmse, K);
    else { for (int i = 0; i < ns; ++i) corr[i] = -1; }
    if (trace_T) std::memcpy(trace_T, T, sizeof(T));
    if (trace_fit) trace_fit[0] = fit;
    if (trace_rmse) trace_rmse[0] = rmse;
    pass = 1;
    for (int it = 0; it < max_iter; ++it) {
        double upd[16], Tn[16];
        if (method == 1) p2plane_update(pcd.data(), tgt, tgt_normals, corr.data(), ns, K, upd, nullptr, nullptr);
        else p2p_update(pcd.data(), tgt, corr.data(), ns, K, upd);
        mat4_mul(upd, T, Tn);
        std::memcpy(T, Tn, sizeof(T));
        for (int i = 0; i < ns; ++i) transform_point(upd, pcd.data() + 3 * (size_t)i);
        ++updates;
        const double bfit = fit, brmse = rmse;
        if (max_corr > 0.0) correspondence_pass(tree, pcd.data(), ns, max_corr, corr.data(), d2.data(), fit, rmse, K);
        if (trace_T) std::memcpy(trace_T + 16 * (size_t)pass, T, sizeof(T));
        if (trace_fit) trace_fit[pass] = fit;
        if (trace_rmse) trace_rmse[pass] = rmse;
        ++pass;
        if (std::fabs(bfit - fit) < rel_fit && std::fabs(brmse - rmse) < rel_rmse) break;
    }
    std::memcpy(out_T16, T, sizeof(T));
    *out_fitness = fit;
    *out_rmse = rmse;
    if (out_updates) *out_updates = updates;
    if (out_ncorr) *out_ncorr = K;
    if (corr_out) std::memcpy(corr_out, corr.data(), sizeof(int) * (size_t)ns);
    return pass;
}

// p2plane normal equations for a given correspondence set (tests of the GPU reduction)
void orc_p2plane_system(const double* src_transformed, const double* tgt, const double* nrm, const int* corr, int ns, double* JTJ36,
                        double* JTr6, double* upd16) {
    int K = 0;
    for (int i = 0; i < ns; ++i) K += corr[i] >= 0;
    p2plane_update(src_transformed, tgt, nrm, corr, ns, K, upd16, JTJ36, JTr6);
}
void orc_p2p_update(const double* src_transformed, const double* tgt, const int* corr, int ns, double* upd16) {
    int K = 0;
    for (int i = 0; i < ns; ++i) K += corr[i] >= 0;
    p2p_update(src_transformed, tgt, corr, ns, K, upd16);
}

}  // extern "C"
