// Host-side plumbing shared by the .cu files: launch helper, parameter blocks, stage entry points.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "common.cuh"

namespace arvc {

struct FilterParams { double min_r2, max_r2, min_h, max_h; };
struct VoxelParams { double voxel; int bx, by, bz, pad; };
struct NormalParams {
    double radius, r2, bin_scale;       // r2 = radius^2, bin_scale = buckets / r2
    float r2_lo, r2_hi, bin_scale_f;    // float32 screening bounds r2 * (1 -/+ 2e-6)
    int max_nn;
    int level;
    int debug;
    float crowded_ratio;                // block kernel: trial radius / radius below which a neighbourhood counts as crowded
};

struct Launcher {
    cudaStream_t stream = nullptr;
    long long launches = 0;
    cudaError_t err = cudaSuccess;
    // optional per-kernel timing with CUDA events on the launching stream (arvc_profile_*)
    bool profile = false;
    struct Rec { const char* name; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get_event() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    // the same, for a kernel given as a plain function pointer with an argument array (shared with the graph builder)
    void launch_ptr(const char* name, const void* func, dim3 grid, dim3 block, void** args) {
        if (err != cudaSuccess) return;
        if (grid.x == 0 || grid.y == 0) return;
        Rec r{name, nullptr, nullptr};
        if (profile) { r.a = get_event(); r.b = get_event(); cudaEventRecord(r.a, stream); }
        const cudaError_t e = cudaLaunchKernel(func, grid, block, args, 0, stream);
        if (profile) { cudaEventRecord(r.b, stream); recs.push_back(r); }
        if (e != cudaSuccess) err = e;
        ++launches;
    }
    template <typename... KArgs, typename... Args>
    void launch(const char* name, void (*kernel)(KArgs...), dim3 grid, dim3 block, Args... args) {
        launch_smem(name, kernel, grid, block, 0, args...);
    }
    template <typename... KArgs, typename... Args>
    void launch_smem(const char* name, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
        if (err != cudaSuccess) return;
        if (grid.x == 0 || grid.y == 0) return;
        Rec r{name, nullptr, nullptr};
        if (profile) { r.a = get_event(); r.b = get_event(); cudaEventRecord(r.a, stream); }
        kernel<<<grid, block, smem, stream>>>(args...);
        if (profile) { cudaEventRecord(r.b, stream); recs.push_back(r); }
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) err = e;
        ++launches;
    }
};

// pair descriptor + state for the device-resident ICP iteration
constexpr int kThist = 32;     // passes whose transformation is kept for the nearest-neighbour certificates

struct __align__(16) PairState {
    double T[16];           // cumulative transformation used by the NEXT pass
    double fitness, rmse;   // of the last completed pass
    double sums[32];        // reduced normal-equation sums of the last pass (debug)
    int passes;             // completed correspondence passes
    int updates;            // executed updates
    int done;
    int ncorr;
    int err;                // OR of the two scans' device error flags
    int nlist;              // entries of the work list of the current pass
    int nfar;               // entries of the far list of the current pass
    int pad[1];
    unsigned long long dbg[8];   // search statistics (ARVC_DEBUG_STATS): skipped / union / fallback queries, queries, fallback by level
    // ---- device-only tail (not copied back): top three rows of the transformation each pass was evaluated at
    double Thist[12 * kThist];
};
constexpr size_t kPairStateHead = 16 * 8 + 2 * 8 + 32 * 8 + 8 * 4 + 8 * 8;   // bytes read back per pair

struct PairDev {
    const ScanDev* src;
    const ScanDev* tgt;
    PairState* state;
    double* partials;       // [warps][kSumStride]
    int* prev;              // [src cap] Morton position of last pass' match in the target, or -1
    float* lb2;             // [src cap] certificate: lower bound on the distance to every other target point ...
    unsigned char* cert_pass;   // [src cap] ... at the pass it was established (255 = none)
    int* list;              // [src cap] source points whose nearest neighbour has to be searched in this pass
    int* far_list;          // [src cap] of those, the ones the shared-candidate phase could not settle: point | start level << 24
    int* corr_trace;        // optional [(max_iter+1)][src cap] (cloud order), may be null
    double* state_trace;    // optional [(max_iter+1)][18]: T16, fitness, rmse
};

struct IcpParams {
    double max_d2;          // max_corr_dist^2
    double max_d;
    double rel_fitness, rel_rmse;
    double cert_margin;     // extra radius (m) the union phase covers so that later passes can skip the search
    int max_iter;
    int method;
    int debug;              // collect search statistics into PairState (ARVC_DEBUG_STATS)
    int pad;
};

constexpr int kSumStride = 32;
constexpr int kIcpBlock = 64;
constexpr int kIcpUnrolled = 4;     // passes enqueued as plain nodes before the device-terminated loop takes over

// Everything the ICP kernels need to know about the batch, in device memory: the kernels' only argument, so that one
// instantiated CUDA graph (fixed arguments) serves every batch of the same shape.
struct BatchDesc {
    const PairDev* pairs;
    int n_pairs;
    int pad;
    IcpParams ip;
};

// Instantiated ICP graphs of a context, keyed by batch shape: unrolled passes 0..kIcpUnrolled-1, then a conditional
// WHILE node whose body is one pass and whose condition ("some pair has not converged") is set on the device.
struct IcpGraph {
    unsigned long long key = 0;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int kernels_prefix = 0, kernels_body = 0;
};
struct IcpGraphCache {
    std::vector<IcpGraph> graphs;
    BatchDesc* d_bd = nullptr;          // fixed address the graphs' kernels read the current batch from
    bool disabled = false;              // graph construction failed once: stay on the unrolled path
    std::string error;
};

void run_preprocess(Launcher& L, const ScanDev* d_scans, int n_scans, int cap_max, const FilterParams& fp, const VoxelParams& vp,
                    bool voxel_on);
void run_normals(Launcher& L, const ScanDev* d_scans, int n_scans, int cap_max, const NormalParams& np, bool any_wide, bool any_narrow, bool tap);
void launch_normals_blk(Launcher& L, const ScanDev* d_scans, int n_scans, int cap_max, const NormalParams& np, bool tap);
// Enqueues the whole iteration of a batch.  `use_graph`: device-terminated loop (CUDA graph with a WHILE node); else
// max_iter + 1 passes are enqueued unconditionally (finished pairs turn into no-ops).  Returns the graph used (or null).
const IcpGraph* run_icp(Launcher& L, IcpGraphCache& cache, const BatchDesc& h_bd, BatchDesc* d_bd_batch, int src_cap_max, int combos_mask,
                        bool use_graph);
void icp_graphs_destroy(IcpGraphCache& cache);
void launch_icp_init(Launcher& L, const PairDev* d_pairs, int n_pairs, const double* d_init, int* d_status);
void launch_icp_pack(Launcher& L, const PairDev* d_pairs, int n_pairs, void* d_records, int* d_status);
void run_plane_fit(Launcher& L, const ScanDev* d_scan, int cap, double* d_orig, int* d_score, double* d_result, double max_z, double thr,
                   int iters, unsigned long long seed);
void run_plane_split(Launcher& L, const ScanDev* d_scan, int cap, double* d_orig, int* d_blk, double* d_near, double* d_far, int* d_counts2,
                     const double* plane, double norm, double thr);
void run_map_build(Launcher& L, const ScanDev* const* d_scans, const double* d_T, int n_scans, int cap_max, long long* d_offsets,
                   double* d_out, long long capacity, bool offsets_only);

}  // namespace arvc
