/*
 * arvc_icp.h — C-ABI of the B200-native ICP scan-matching engine (libarvc_icp.so).
 *
 * Drop-in boundary for the hot path of JudithV/LIDAR_SLAM_ARVC.  The reference has no FFI layer of its
 * own: its registration API is the Python classes keyframemanager.KeyFrame / KeyFrameManager, which call
 * Open3D (CPU).  Each entry point below names the reference call it replaces (paths relative to the
 * reference tree).  Plain pointers and sizes only; all arrays are host memory unless stated; row-major.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (ARVC_E_*); arvc_last_error() gives the text;
 *   - a context owns one CUDA device + one stream; calls on one context must be serialised by the caller;
 *   - scan ids are caller-chosen 64-bit keys (the Python host uses the index in KeyFrameManager.keyframes);
 *   - "cloud order" of a preprocessed scan = order of `pointcloud_filtered.points` in the reference:
 *       filter-stable order when voxel_size is off, ascending voxel key (ix,iy,iz) when it is on;
 *     all indices that cross this boundary (correspondences, normals, points) are in cloud order;
 *   - there is NO CPU fallback: without a CUDA device arvc_ctx_create fails.
 */
#ifndef ARVC_ICP_H
#define ARVC_ICP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct arvc_ctx arvc_ctx;

enum {
    ARVC_OK = 0,
    ARVC_E_CUDA = -1,      /* a CUDA runtime call failed */
    ARVC_E_ARG = -2,       /* bad argument (null pointer, unknown scan id, bad method, ...) */
    ARVC_E_STATE = -3,     /* scan not uploaded / not preprocessed / normals missing for point-to-plane */
    ARVC_E_CAPACITY = -4,  /* device-side capacity exceeded (hash grid overflow, voxel index range) */
    ARVC_E_NOMEM = -5
};

enum { ARVC_P2P = 0, ARVC_P2PLANE = 1 };

/* --- context --------------------------------------------------------------------------------------- */
int arvc_ctx_create(int device, arvc_ctx** out);
void arvc_ctx_destroy(arvc_ctx* ctx);
const char* arvc_last_error(const arvc_ctx* ctx); /* ctx may be NULL: error of the last failed create */
int arvc_sync(arvc_ctx* ctx);                     /* wait for all queued work of the context */
void* arvc_stream(arvc_ctx* ctx);                 /* the context's cudaStream_t (for event timing) */
int arvc_version(void);
/* Engine switches (0 / 1), for A/B measurements and parity tests; results never depend on them:
 *   "icp_loop_graph"  1 (default): registration loops are ONE CUDA graph whose WHILE node ends the iteration on the
 *                     device with the last convergence; 0: max_iter + 1 passes are enqueued unconditionally.
 *   "normals_tap"     0 (default); 1: scans preprocessed from now on also record the neighbour set of every normal
 *                     (arvc_scan_get_neighbors) - separate kernel instantiations, max_nn x 4 bytes per point. */
int arvc_ctx_set_option(arvc_ctx* ctx, const char* name, int value);
/* number of kernels launched by this context so far (bench.py's gpu_launches) */
int64_t arvc_kernel_launches(const arvc_ctx* ctx);
/* Per-kernel device timing with CUDA events on the context stream (bench.py's roofline line).  enable(1) starts
 * recording every launch, report() synchronises and writes "name,launches,total_ms\n" lines into buf. */
int arvc_profile_enable(arvc_ctx* ctx, int on);
int arvc_profile_report(arvc_ctx* ctx, char* buf, size_t cap);

/* --- scans -----------------------------------------------------------------------------------------
 * Replaces KeyFrame.load_pointcloud's result handed to Open3D (keyframemanager/keyframe.py:41-45): the
 * PCD payload (float32 xyz, n points, NaNs allowed) is copied to the device.  Re-uploading an id replaces it.
 * The _f64 variant is for PCD files with double fields.  Asynchronous w.r.t. the host when `xyz` is pinned (keep the
 * buffer alive until arvc_sync): the copy runs on a dedicated copy stream and overlaps the kernels of scans uploaded
 * earlier; arvc_scan_preprocess waits for it on the device. */
int arvc_scan_upload_f32(arvc_ctx* ctx, int64_t scan_id, const float* xyz, int n);
int arvc_scan_upload_f64(arvc_ctx* ctx, int64_t scan_id, const double* xyz, int n);
/* Host wait for the upload of ONE scan (its copy-stream event): afterwards the host buffer handed to arvc_scan_upload_*
 * may be reused.  Returns at once when the copy has already finished.  (KeyFrame.load_pointcloud recycles its pinned
 * staging buffers with this instead of a device-wide synchronisation.) */
int arvc_scan_wait_upload(arvc_ctx* ctx, int64_t scan_id);
/* Forget the cached preprocessing of a scan (host-side flag only).  The reference never sets
 * KeyFrame.pre_processed (keyframe.py:39,114), i.e. every pre_process() call redoes the work; this engine caches by
 * scan id and parameters, and this call restores the reference's "redo" behaviour (used by bench.py). */
int arvc_scan_invalidate(arvc_ctx* ctx, int64_t scan_id);
/* KeyFrame.unload_pointcloud (keyframe.py:61-72): drop every device buffer of the scan. */
int arvc_scan_free(arvc_ctx* ctx, int64_t scan_id);

typedef struct arvc_preprocess_params {
    /* KeyFrame.filter_radius_height (keyframe.py:74-94): keep r2=x*x+y*y with
     *   r2 < max_radius2 && r2 > min_radius2 && z > min_height && z < max_height   (strict, float64).
     * The radii arrive already squared: Python computes `radius ** 2` exactly as keyframe.py:92 does. */
    double min_radius2, max_radius2, min_height, max_height;
    /* voxel_down_sample (keyframe.py:111,151,159): <= 0 or NaN means "voxel_size is None". */
    double voxel_size;
    /* estimate_normals(KDTreeSearchParamHybrid(radius, max_nn)) (keyframe.py:160-162); want_normals = 0
     * for 'icppointpoint' preprocessing (keyframe.py:148-151). */
    double normal_radius;
    int32_t max_nn;
    int32_t want_normals;
    /* hash-grid hints (0 = defaults): finest cell edge in metres; largest search distance the grid must
     * be able to answer exactly (the ICP max_correspondence_distance, icp_parameters.yaml:22). */
    double grid_cell;
    double grid_max_dist;
} arvc_preprocess_params;

/* KeyFrame.pre_process (keyframe.py:113-162) for a batch of uploaded scans: filter -> [voxel] -> spatial
 * sort + hash grid -> [normals].  Queued on the context stream; no host synchronisation. */
int arvc_scan_preprocess(arvc_ctx* ctx, int n_scans, const int64_t* scan_ids, const arvc_preprocess_params* p);
/* The same work for scans the caller will need NEXT, queued on a second compute stream so that it overlaps what is already
 * queued on the context stream - in the reference's loop (run_scanmatcher.py:196-213: load i+1, pre_process i+1,
 * compute_transformation(i, i+1), one pair per call) the scan i+2 is preprocessed while the pair (i, i+1) is registered.
 * Call it between arvc_icp_batch_async and arvc_icp_batch_finish.  Every later call on the context is ordered after it;
 * a later arvc_scan_preprocess with the same parameters finds the scan done.  Scans that already hold preprocessed
 * state are handled as by arvc_scan_preprocess (on the context stream). */
int arvc_scan_preprocess_ahead(arvc_ctx* ctx, int n_scans, const int64_t* scan_ids, const arvc_preprocess_params* p);

/* Sizes after preprocessing (synchronises): raw, after filter, final cloud size; has_normals. */
int arvc_scan_info(arvc_ctx* ctx, int64_t scan_id, int* n_raw, int* n_filtered, int* n_points, int* has_normals);
/* np.asarray(pointcloud_filtered.points / .normals): xyz[n_points*3], normals[n_points*3] (NULL to skip), cloud order. */
int arvc_scan_get_points(arvc_ctx* ctx, int64_t scan_id, double* xyz, double* normals);
/* KeyFrameManager.build_map (keyframemanager.py:154-184) without the GUI, for a batch of uploaded keyframes: per
 * keyframe filter_radius_height(radii, heights) (keyframe.py:74-94) -> down_sample (:108-111) ->
 * KeyFrame.transform(T_i) (:399-400, Open3D PointCloud::Transform incl. the division by w) -> concatenation in
 * keyframe order.  `p` carries the map's radii / heights / voxel size (want_normals = 0).  T: n_scans row-major 4x4.
 * xyz_out[capacity_points*3] receives the map, offsets_out[n_scans+1] the start of every keyframe's points
 * (offsets_out[n_scans] = total).  Returns ARVC_E_CAPACITY with valid offsets when capacity_points is too small
 * (sum of the raw sizes always suffices).  Synchronises. */
int arvc_map_build(arvc_ctx* ctx, int n_scans, const int64_t* scan_ids, const double* T, const arvc_preprocess_params* p,
                   double* xyz_out, int64_t capacity_points, int64_t* offsets_out);
/* 'icp2planes' preprocessing (keyframe.py:164-189), on the preprocessed cloud of a scan:
 * KeyFrame.calculate_plane (keyframe.py:417-436): plane a x + b y + c z + d = 0 (unit normal) through the points with
 * z < max_z (-0.5), RANSAC with 3-point hypotheses (`iterations` 1000, inlier distance `dist_threshold` 0.01).  Open3D's
 * RANSAC is unseeded; here the samples are a hash of (seed, iteration), so equal inputs give equal planes.  Synchronises.
 * KeyFrame.segment_plane (keyframe.py:438-461): points with |a x + b y + c z + d| / sqrt(a^2+b^2+c^2) < threshold (0.4)
 * become the raw cloud of scan `near_id`, the others of `far_id` (float64, order preserved: select_by_index /
 * invert=True); both are then preprocessed like any uploaded scan.  Synchronises. */
int arvc_scan_fit_plane(arvc_ctx* ctx, int64_t scan_id, double max_z, double dist_threshold, int iterations, uint64_t seed,
                        double* plane_out /* [4] */, int32_t* n_inliers);
int arvc_scan_split_plane(arvc_ctx* ctx, int64_t src_id, const double* plane /* [4] */, double threshold, int64_t near_id, int64_t far_id,
                          int32_t* n_near, int32_t* n_far);
/* Parity taps.  raw_index[n_filtered]: raw indices kept by the filter, ascending.
 * voxel keys[n_points*3] (Open3D voxel index per output point) and counts[n_points]; voxel mode only.
 * nn_count[n_points]: number of neighbours used by the normal of each point (after the k / radius cut). */
int arvc_scan_get_filter_indices(arvc_ctx* ctx, int64_t scan_id, int32_t* raw_index);
int arvc_scan_get_voxels(arvc_ctx* ctx, int64_t scan_id, int32_t* keys, int32_t* counts);
int arvc_scan_get_nn_counts(arvc_ctx* ctx, int64_t scan_id, int32_t* nn_count);
/* Neighbour SETS behind the normals (KDTreeSearchParamHybrid: the max_nn nearest, then d2 < radius^2; keyframe.py:160-162):
 * for every queried point (cloud order) the cloud indices of the neighbours whose covariance gave its normal, in no
 * particular order, padded with -1 to max_nn per row; out_cnt[q] = their number.  Only for scans preprocessed while the
 * context option "normals_tap" was set (the production kernels do not record anything). */
int arvc_scan_get_neighbors(arvc_ctx* ctx, int64_t scan_id, int n_query, const int32_t* point_ids, int32_t* out_idx /* [n_query*max_nn] */,
                            int32_t* out_cnt /* [n_query] */);
/* Device counters of a preprocessed scan, counters[16]: 0 points after the filter, 1 final points, 2 error flags,
 * 3 occupied grid cells (all levels), 4 normals recomputed in canonical order, 5 points served by the per-point
 * normals kernel, 6 blocks / 7 single points the block kernel handed back, 8 blocks served at a trial radius. */
int arvc_scan_get_counters(arvc_ctx* ctx, int64_t scan_id, int32_t* counters);

/* --- registration ---------------------------------------------------------------------------------- */
typedef struct arvc_icp_params {
    double max_corr_dist; /* ICP_PARAMETERS.distance_threshold, config/icp_parameters.yaml:22 (10.0) */
    double rel_fitness;   /* Open3D ICPConvergenceCriteria defaults, because keyframe.py:246-252 passes none: 1e-6 */
    double rel_rmse;      /* 1e-6 */
    int32_t max_iter;     /* 30 */
    int32_t method;       /* ARVC_P2P: TransformationEstimationPointToPoint (keyframe.py:248);
                             ARVC_P2PLANE: TransformationEstimationPointToPlane (keyframe.py:252) */
} arvc_icp_params;

/* KeyFrameManager.compute_transformation(i, j, Tij) (keyframemanager/keyframemanager.py:52-75) ->
 * KeyFrame.local_registration_simple (keyframe.py:231-260) -> o3d registration_icp(source = scan j,
 * target = scan i, threshold, init, estimation), for n_pairs independent pairs in one call.
 *   tgt_ids[p] = scan i (target), src_ids[p] = scan j (source), init_T[16*p..] = Tij.array (row-major).
 * Outputs (host): out_T[16*p..] = reg.transformation, fitness[p], rmse[p] = reg.fitness / reg.inlier_rmse,
 *   updates[p] = ICP iterations executed, n_corr[p] = len(reg.correspondence_set).  Any may be NULL.
 * The whole iteration (correspondence search, residual/Jacobian reduction, 6x6 solve or Umeyama,
 * convergence test) runs on the device; the call synchronises once at the end to deliver the results. */
int arvc_icp_batch(arvc_ctx* ctx, int n_pairs, const int64_t* tgt_ids, const int64_t* src_ids, const double* init_T,
                   const arvc_icp_params* p, double* out_T, double* fitness, double* rmse, int32_t* updates,
                   int32_t* n_corr);

/* Same, asynchronous: results stay on the device until arvc_icp_batch_finish; lets the host overlap the
 * upload of the next batch.  Records are 160 bytes: { int32 pair, updates, n_corr, passes; double T[16]; fitness; rmse }. */
typedef struct arvc_result_record {
    int32_t pair, updates, n_corr, passes;
    double T[16];
    double fitness, rmse;
} arvc_result_record;
int arvc_icp_batch_async(arvc_ctx* ctx, int n_pairs, const int64_t* tgt_ids, const int64_t* src_ids, const double* init_T,
                         const arvc_icp_params* p, uint64_t* ticket);
int arvc_icp_batch_finish(arvc_ctx* ctx, uint64_t ticket, arvc_result_record* records /* [n_pairs] host */);
/* Multi-GPU gather: device address of the pending batch's records, arvc_result_record[n_pairs], written
 * on the context's stream when the iteration has ended and valid until arvc_icp_batch_finish(ticket).  A collective
 * enqueued on (or ordered after) arvc_stream() can read them in place: no device -> host -> device bounce. */
int arvc_icp_batch_device_records(arvc_ctx* ctx, uint64_t ticket, const void** d_records, int* n_pairs);

/* Parity tap: one pair, recording every pass.  corr[pass*n_src + i] = target index (cloud order) matched to
 * source point i (cloud order) in that pass or -1; trace_T[pass*16..] = transformation the pass was evaluated
 * at; trace_fitness/rmse[pass].  Arrays sized for (max_iter+1) passes; n_passes receives the count. */
int arvc_icp_trace(arvc_ctx* ctx, int64_t tgt_id, int64_t src_id, const double* init_T, const arvc_icp_params* p,
                   int32_t* corr, double* trace_T, double* trace_fitness, double* trace_rmse, int32_t* n_passes,
                   arvc_result_record* result);

/* Load path helper (keyframe.py:41-45, o3d.io.read_point_cloud): LZF decompression of the payload of a
 * `DATA binary_compressed` PCD file.  Host only.  Returns bytes written or -1 on a malformed stream. */
long long arvc_lzf_decompress(const unsigned char* in, size_t n_in, unsigned char* out, size_t n_out);

/* pinned host staging (optional; plain malloc'ed pointers work too, just slower for H2D) */
void* arvc_host_alloc(size_t bytes);
void* arvc_ctx_host_alloc(arvc_ctx* ctx, size_t bytes);   /* same, after selecting the context's device (for helper threads) */
/* Optional: grow the device memory pool by `bytes` now.  Scans and batches allocate from that pool; when it has to grow
 * while kernels run, the allocating call waits for them.  A caller that knows its footprint (loop closing keeps every
 * visited keyframe resident - loopclosing.py:163-178 never unloads - at ~12 MB per 64-beam scan) reserves it up front:
 * once, right after arvc_ctx_create.  (On a pool that already holds memory it does not help: the first batches after it
 * were measured slower.) */
int arvc_ctx_reserve(arvc_ctx* ctx, size_t bytes);
void arvc_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* ARVC_ICP_H */
