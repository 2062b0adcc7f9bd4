// Host-side plumbing shared by the .cu files: launch helper, parameter blocks, stage entry points.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "common.cuh"

namespace arvc {

struct FilterParams { double min_r2, max_r2, min_h, max_h; };
struct VoxelParams { double voxel; int bx, by, bz, pad; };
struct NormalParams { double radius; int max_nn; int level; };

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
struct __align__(16) PairState {
    double T[16];           // cumulative transformation used by the NEXT pass
    double fitness, rmse;   // of the last completed pass
    double sums[32];        // reduced normal-equation sums of the last pass (debug)
    int passes;             // completed correspondence passes
    int updates;            // executed updates
    int done;
    int ncorr;
    unsigned ticket;
    int err;                // OR of the two scans' device error flags
    int pad[2];
    unsigned long long dbg[8];   // search statistics, accumulated over all passes
    unsigned long long tl[8];    // per-phase clock cycles summed over blocks (debug timeline)
};

struct PairDev {
    const ScanDev* src;
    const ScanDev* tgt;
    PairState* state;
    double* partials;       // [nblk][kSumStride]
    int* prev;              // [src cap] Morton position of last pass' match in the target, or -1
    int* corr_trace;        // optional [(max_iter+1)][src cap] (cloud order), may be null
    double* state_trace;    // optional [(max_iter+1)][18]: T16, fitness, rmse
};

struct IcpParams {
    double max_d2;          // max_corr_dist^2
    double max_d;
    double rel_fitness, rel_rmse;
    int max_iter;
    int method;
    int defer_level;        // coarsest grid level the 8-lane search handles itself (farther queries: block-wide phase)
    int debug;              // collect search statistics / phase timeline into PairState (ARVC_DEBUG_STATS)
};

constexpr int kSumStride = 32;
constexpr int kIcpBlock = 128;

void run_preprocess(Launcher& L, const ScanDev* d_scans, int n_scans, int cap_max, const FilterParams& fp, const VoxelParams& vp,
                    bool voxel_on);
void run_normals(Launcher& L, const ScanDev* d_scans, int n_scans, int cap_max, const NormalParams& np, bool any_wide, bool any_narrow);
void run_icp(Launcher& L, const PairDev* d_pairs, int n_pairs, int src_cap_max, const IcpParams& ip, int combos_mask);

}  // namespace arvc
