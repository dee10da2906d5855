/*
 * libsleapb200 -- C-ABI of the B200-native SLEAP inference path.
 *
 * The reference (talmolab/sleap v1.4.1) has no FFI: its seam is the Python class surface
 * sleap.nn.inference.{Predictor, InferenceModel, InferenceLayer} plus the function-level
 * modules sleap.nn.peak_finding and sleap.nn.paf_grouping, all of which dispatch TensorFlow
 * ops.  Every entry point below names the reference interface (file:line under
 * /root/reference) whose device work it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain C: pointers + sizes only; all functions return 0 (SB_OK) or a negative SB_ERR_*;
 *     nothing throws.  sb_last_error() returns the message of the last failure.
 *   - one handle per GPU; calls on one handle are serialised by the caller; different handles
 *     are independent.  Device buffers are owned by the handle, host buffers by the caller.
 *   - pointers named *_host are host memory, *_dev are device memory on the handle's GPU.
 *   - images / maps are NHWC, row-major; points are (x, y) float32; missing values are NaN.
 */
#ifndef SLEAP_B200_H_
#define SLEAP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK 0
#define SB_ERR_INVALID (-1)
#define SB_ERR_CUDA (-2)
#define SB_ERR_UNSUPPORTED (-3)
#define SB_ERR_NO_DEVICE (-4)

/* refinement modes (sleap/nn/peak_finding.py:337-420, 451-532: refinement=None|"integral"|"local") */
#define SB_REFINE_NONE 0
#define SB_REFINE_INTEGRAL 1
#define SB_REFINE_LOCAL 2

/* overflow flags reported per sample (capacity-bounded device buffers; the reference is unbounded) */
#define SB_FLAG_PEAKS_TRUNCATED 1
#define SB_FLAG_NODE_PEAKS_TRUNCATED 2
#define SB_FLAG_INSTANCES_TRUNCATED 4

typedef struct sb_handle_s* sb_handle_t;

/* ---- lifetime --------------------------------------------------------------------------- */
int sb_version(void);
/* replaces the device selection of sleap/nn/system.py:29-110 (one GPU per process) */
int sb_create(int device_id, sb_handle_t* out_handle);
int sb_destroy(sb_handle_t h);
const char* sb_last_error(sb_handle_t h); /* h may be NULL: last error of this thread */
int sb_synchronize(sb_handle_t h);
/* number of kernels this handle has launched so far (bench.py "gpu_launches") */
int sb_gpu_launches(sb_handle_t h);
/* use the caller's CUDA stream (e.g. torch's current stream) for all subsequent work; NULL = own */
int sb_set_stream(sb_handle_t h, void* cuda_stream);

/* ---- stage level: peak finding (host buffers) --------------------------------------------
 * sleap/nn/peak_finding.py:451-532 find_local_peaks (+ :249-308 rough, :646-707 with offsets).
 * cms (B,H,W,C) f32.  offsets_host: NULL, or learned offset maps (B,H,W,2C) -> "with_offsets".
 * Outputs are ordered like tf.where: (sample, y, x, channel) row-major.  At most
 * max_peaks_per_sample peaks are kept per sample (first in order); out arrays must hold
 * B*max_peaks_per_sample entries.  out_flags (B) receives SB_FLAG_* (may be NULL). */
int sb_find_local_peaks(sb_handle_t h, const float* cms_host, int B, int H, int W, int C,
                        float threshold, int refinement, int integral_patch_size,
                        const float* offsets_host, int max_peaks_per_sample,
                        float* out_points, float* out_vals, int32_t* out_sample_inds,
                        int32_t* out_channel_inds, int32_t* out_n_peaks, int32_t* out_flags);

/* sleap/nn/peak_finding.py:337-420 find_global_peaks (+ :193-246 rough, :566-643 with offsets).
 * out_points (B,C,2), out_vals (B,C). */
int sb_find_global_peaks(sb_handle_t h, const float* cms_host, int B, int H, int W, int C,
                         float threshold, int refinement, int integral_patch_size,
                         const float* offsets_host, float* out_points, float* out_vals);

/* sleap/nn/peak_finding.py:135-190 crop_bboxes on make_centered_bboxes
 * (sleap/nn/data/instance_cropping.py:124-166).  images (B,H,W,C) u8 or f32; centroids (n,2) xy;
 * out (n,crop_h,crop_w,C) same dtype as images (u8: float->u8 truncation as the reference). */
int sb_crop_centered(sb_handle_t h, const void* images_host, int images_are_u8, int B, int H,
                     int W, int C, const float* centroids, const int32_t* sample_inds, int n,
                     int crop_h, int crop_w, void* out_crops);

/* ---- stage level: PAF grouping (host buffers) --------------------------------------------
 * sleap/nn/paf_grouping.py:406-550 score_paf_lines_batch (:82-142 candidates, :145-275 line
 * sampling, :278-403 scoring).  pafs (B,Hp,Wp,2E) f32; peaks (N,2) image px, peak_channel_inds
 * (N), grouped by sample through peak_offsets (B+1).  Outputs: flat candidate lists in
 * (sample, edge, src-major) order; cand_offsets (B+1).  cap = capacity of the out arrays. */
int sb_score_paf_lines_batch(sb_handle_t h, const float* pafs_host, int B, int Hp, int Wp, int C2,
                             const float* peaks, const int32_t* peak_channel_inds,
                             const int32_t* peak_offsets, const int32_t* skeleton_edges,
                             int n_edges, int n_nodes, int n_line_points, int pafs_stride,
                             float max_edge_length_ratio, float dist_penalty_weight,
                             int cap, int32_t* out_edge_inds, int32_t* out_edge_peak_inds,
                             float* out_line_scores, int32_t* out_cand_offsets);

/* sleap/nn/paf_grouping.py:145-222 make_line_subs, :225-275 get_paf_lines, :325-403
 * score_paf_lines for an explicit candidate list of ONE sample.  pafs_sample (Hp,Wp,C2) may be
 * NULL when lines_in (n,P,2) is given (score_paf_lines on pre-gathered lines).  Any of the
 * outputs may be NULL: out_subs (n,P,2) int32 [row,col]; out_lines (n,P,2); out_scores (n). */
int sb_paf_lines(sb_handle_t h, const float* pafs_sample, int Hp, int Wp, int C2,
                 const float* lines_in, const float* peaks, int n_peaks,
                 const int32_t* edge_peak_inds, const int32_t* edge_inds, int n, int n_line_points,
                 int pafs_stride, float max_edge_length, float dist_penalty_weight,
                 int32_t* out_subs, float* out_lines, float* out_scores);

/* sleap/nn/peak_finding.py:311-334 integral_regression: cms (N,h,w,C), xv (w), yv (h) ->
 * x_hat, y_hat (N,C). */
int sb_integral_regression(sb_handle_t h, const float* cms, int N, int Hh, int Ww, int C,
                           const float* xv, const float* yv, float* x_hat, float* y_hat);
/* sleap/nn/peak_finding.py:78-132 find_offsets_local_direction: patches (N,3,3,1) -> (N,2). */
int sb_find_offsets_local_direction(sb_handle_t h, const float* patches, int N, float delta,
                                    float* out_offsets);

/* sleap/nn/utils.py:79-98 tf_linear_sum_assignment, batched, as called from
 * sleap/nn/paf_grouping.py:621-650: problem p is the (n_src[p], n_dst[p]) row-major SCORE
 * matrix at scores + offsets[p]; cost = -score, NaN -> +inf; minimum-cost assignment with
 * SciPy's rectangular shortest-augmenting-path semantics.  Outputs per problem at
 * p*max_k: rows (ascending), cols, matched scores; out_counts[p] (0 when infeasible). */
int sb_linear_sum_assignment_batch(sb_handle_t h, const float* scores, const int32_t* n_src,
                                   const int32_t* n_dst, const int32_t* offsets, int n_problems,
                                   int max_k, int32_t* out_rows, int32_t* out_cols,
                                   float* out_scores, int32_t* out_counts);

/* sleap/nn/paf_grouping.py:1115-1290 group_instances_batch (:984-1112 per sample, :799-914
 * greedy assignment, :917-981 instance assembly).  Matches are edge-LOCAL indices grouped by
 * sample through match_offsets (B+1), in the order produced by match_candidates_batch.
 * Outputs: out_instances (B,max_instances,n_nodes,2), out_peak_scores (B,max_instances,n_nodes),
 * out_instance_scores (B,max_instances), out_n_instances (B). */
int sb_group_instances_batch(sb_handle_t h, int B, int n_nodes, const float* peaks,
                             const float* peak_vals, const int32_t* peak_channel_inds,
                             const int32_t* peak_offsets, const int32_t* match_edge_inds,
                             const int32_t* match_src_peak_inds, const int32_t* match_dst_peak_inds,
                             const float* match_line_scores, const int32_t* match_offsets,
                             const int32_t* edge_types, int n_edges,
                             const int32_t* sorted_edge_inds, int n_sorted, int min_instance_peaks,
                             float min_line_scores, int max_instances, float* out_instances,
                             float* out_peak_scores, float* out_instance_scores,
                             int32_t* out_n_instances);

/* ---- models --------------------------------------------------------------------------------
 * A model is the flat op-list compiled (host side, sleap_b200/nn/architectures.py) from the
 * reference's backbone + heads graph: sleap/nn/model.py:312-364, architectures/unet.py,
 * encoder_decoder.py, hourglass.py, heads.py:42-63.  ops: n_ops records of SB_OP_WORDS int32
 * (layout in sleap_b200/nn/oplist.py); weights: float32 blob the ops index into.
 * Replaces tf.keras.models.load_model + keras_model(imgs) (sleap/nn/inference.py:3203-3213,
 * :2875).  precision: 0 = fp16 tensor-core path (fp32 accumulate), 1 = fp32 CUDA-core path,
 * 2 = fp32-grade results on the tensor-core path: activations / weights as hi + lo fp16 pairs; the op-list must
 * have been compiled for it (compile_model(split=True): 3C physical channels [lo | hi | hi] per fp16 tensor, weight
 * rows [Wh | Wl | Wh], fp32 frame buffer; conv out_C stays the logical C_out) -- DESIGN.md 5.7. */
#define SB_OP_WORDS 24
int sb_load_model(sb_handle_t h, const int32_t* ops, int n_ops, const float* weights,
                  int64_t n_weights, int precision, int* out_model_id);

/* Plan device buffers for (max_batch, H, W, C_in) uint8/float input frames. */
int sb_model_configure(sb_handle_t h, int model_id, int max_batch, int H, int W, int C_in);

/* Forward only (conv parity tests): images_host (B,H,W,C_in) u8 (images_are_u8) or f32 in [0,1].
 * Copies each requested output tensor (op-list buffer id) to host as f32 NHWC. */
int sb_model_forward(sb_handle_t h, int model_id, const void* images_host, int images_are_u8,
                     int B, int n_outputs, const int32_t* output_buffer_ids,
                     float** out_host_ptrs);

/* Device-side timing of one forward pass, op by op (CUDA events on the launching stream).
 * out_kind: 0 other, 1 tensor-core conv, 2 CUDA-core conv; out_flops: 2*MACs for the batch. */
int sb_model_profile_ops(sb_handle_t h, int model_id, const uint8_t* frames_dev, int B, int cap,
                         float* out_ms, int32_t* out_kind, double* out_flops, int32_t* out_n_ops);

/* Live timing of the network part of every step: while enabled, each forward pass (whatever entry point runs it:
 * sb_model_forward, sb_infer_*, sb_bottomup_submit, sb_infer_topdown) is bracketed by a pair of CUDA events on the
 * launching stream.  A call synchronises that stream, returns the milliseconds of the passes recorded since the last
 * call (at most cap, at most 1024 are kept), clears them, and switches the recording on / off.  bench.py derives
 * `roofline.achieved` from the passes of its timed region.  Replaces nothing in the reference (it has no device
 * timers; tf.profiler is its tool). */
int sb_model_forward_times(sb_handle_t h, int model_id, int enable, int cap, float* out_ms, int32_t* out_n);

/* ---- fused predictors ------------------------------------------------------------------------
 * sleap/nn/inference.py:2737-3003 BottomUpInferenceLayer.call: preprocess -> net -> local peaks
 * -> * cm_output_stride -> PAFScorer.predict -> (/input_scale + 0.5). */
typedef struct sb_bottomup_params {
  int32_t cms_buffer, pafs_buffer, offsets_buffer; /* op-list buffer ids (offsets: -1 if none) */
  int32_t cm_output_stride, paf_output_stride;
  float peak_threshold;
  int32_t refinement, integral_patch_size;
  int32_t n_nodes, n_edges;
  const int32_t* edges;            /* (n_edges, 2) node indices */
  const int32_t* sorted_edge_inds; /* BFS edge order from the topological root */
  int32_t n_sorted;
  int32_t n_line_points;
  float max_edge_length_ratio, dist_penalty_weight, min_line_scores;
  int32_t min_instance_peaks;
  float input_scale;
  int32_t max_peaks_per_sample, max_node_peaks, max_instances;
} sb_bottomup_params;

int sb_bottomup_configure(sb_handle_t h, int model_id, const sb_bottomup_params* params);

/* frames: (B,H,W,C_in) uint8.  *_host variant copies H2D / D2H inside; *_dev takes frames
 * already resident in HBM and leaves results on the device (pointers returned through
 * sb_bottomup_device_outputs).  Outputs: instance_peaks (B,max_instances,n_nodes,2),
 * instance_peak_vals (B,max_instances,n_nodes), instance_scores (B,max_instances),
 * n_valid (B), flags (B). */
int sb_infer_bottomup(sb_handle_t h, int model_id, const uint8_t* frames_host, int B,
                      float* out_instance_peaks, float* out_instance_peak_vals,
                      float* out_instance_scores, int32_t* out_n_valid, int32_t* out_flags);
int sb_infer_bottomup_dev(sb_handle_t h, int model_id, const uint8_t* frames_dev, int B);
/* Streaming form of sb_infer_bottomup for many batches (sleap/nn/inference.py:377-420, the
 * Predictor batch loop): sb_bottomup_submit queues the H2D copy (copy stream), the network and the
 * post-processing of one batch into slot 0/1 and returns; sb_bottomup_collect blocks until that
 * slot's results are in host memory.  Submitting batch i+1 before collecting batch i overlaps its
 * upload with the compute of batch i.  frames_host should be pinned for a truly asynchronous copy. */
int sb_bottomup_submit(sb_handle_t h, int model_id, const uint8_t* frames_host, int B, int slot);
int sb_bottomup_collect(sb_handle_t h, int model_id, int slot, int B, float* out_instance_peaks,
                        float* out_instance_peak_vals, float* out_instance_scores,
                        int32_t* out_n_valid, int32_t* out_flags);

/* sb_infer_bottomup_dev returns as soon as the work is queued: the network runs on the handle's
 * stream and the post-processing on a second stream, so that it overlaps the network of the next
 * call.  sb_bottomup_wait_results makes the handle's stream wait (device side) for the results of
 * the last call; sb_get_post_stream exposes the post-processing stream; sb_synchronize joins both. */
int sb_bottomup_wait_results(sb_handle_t h, int model_id);
int sb_get_post_stream(sb_handle_t h, void** out_stream);
int sb_bottomup_device_outputs(sb_handle_t h, int model_id, float** instance_peaks_dev,
                               float** instance_peak_vals_dev, float** instance_scores_dev,
                               int32_t** n_valid_dev, int32_t** flags_dev);
/* Device pointer of the contiguous per-frame result records the grouping kernel writes ([max_batch][width] float32,
 * width = ceil4(max_instances*n_nodes*3 + max_instances + 2): peaks | peak values | instance scores | n_valid | flags). */
int sb_bottomup_device_records(sb_handle_t h, int model_id, float** records_dev);
/* PAF graph of the last bottom-up call (return_paf_graph, inference.py:2995-3001); same layout
 * as sb_find_local_peaks / sb_score_paf_lines_batch outputs. */
int sb_bottomup_fetch_graph(sb_handle_t h, int model_id, int B, int cap_peaks, float* peaks,
                            float* peak_vals, int32_t* peak_channel_inds, int32_t* peak_offsets,
                            int cap_cands, int32_t* edge_inds, int32_t* edge_peak_inds,
                            float* line_scores, int32_t* cand_offsets);

/* The same post-processing chain on caller-supplied maps (no network): cms (B,H,W,n_nodes),
 * pafs (B,Hp,Wp,2*n_edges), optional learned offsets (B,H,W,2*n_nodes).  This is the entry the
 * parity tests drive (identical cms/pafs in -> bit-exact peaks / assignments out).  The PAF-graph
 * outputs are optional (peaks == NULL skips them).  params->*_buffer fields are ignored. */
int sb_bottomup_from_maps(sb_handle_t h, const sb_bottomup_params* params, const float* cms_host,
                          int B, int H, int W, const float* pafs_host, int Hp, int Wp,
                          const float* offsets_host, float* out_instance_peaks,
                          float* out_instance_peak_vals, float* out_instance_scores,
                          int32_t* out_n_valid, int32_t* out_flags, int cap_peaks, float* peaks,
                          float* peak_vals, int32_t* peak_channel_inds, int32_t* peak_offsets,
                          int cap_cands, int32_t* edge_inds, int32_t* edge_peak_inds,
                          float* line_scores, int32_t* cand_offsets);

/* ---- multi-GPU: exchange of the per-frame result records over NVLink peer memory -----------------
 * The reference runs on one GPU (sleap/nn/system.py:29-46 rejects more than one visible device); its consumer of the
 * per-frame results is Predictor._make_labeled_frames_from_generator (sleap/nn/inference.py:3230-3343).  Here frames
 * are sharded over one process per GPU and the grouping kernel's epilogue writes every frame's fixed-size record
 * [max_instances*n_nodes*2 peaks | max_instances*n_nodes values | max_instances scores | n_valid | flags] (float32)
 * directly into a gather window in EVERY rank's HBM (CUDA-IPC peer mappings over NVLink / NVSwitch) -- no collective
 * call, no rank waits for another inside a step.  Windows are `generations` deep; a consumer acknowledges a step to
 * all producers, and a producer only ever waits when it is `generations` steps ahead of the slowest consumer.
 *   sb_gather_init     allocate this rank's window, return its 64-byte CUDA IPC handle (exchange the handles of all
 *                      ranks out of band, e.g. torch.distributed.all_gather_object)
 *   sb_gather_connect  map every peer's window (all_ipc_handles: world x 64 bytes, rank order); from now on every
 *                      sb_infer_bottomup* / sb_bottomup_submit call pushes its records (one "step" per call)
 *   sb_gather_collect  host consumer: records of `step` from all ranks -> out_records_host [world][B][width]
 *                      (rank-major = frame order for contiguous shards), out_counts[r] = frames rank r pushed
 *   sb_gather_consume_dev  device consumer: wait + acknowledge on the post-processing stream (step < 0: the next
 *                      unconsumed step; results stay in the window returned by sb_gather_window until `generations`
 *                      later steps have been pushed)
 * All waits are bounded (5 s): a dead peer produces an error from sb_gather_collect / sb_gather_status, not a hang. */
#define SB_IPC_HANDLE_BYTES 64
int sb_gather_init(sb_handle_t h, int model_id, int rank, int world, int generations, void* out_ipc_handle);
int sb_gather_connect(sb_handle_t h, int model_id, const void* all_ipc_handles);
int sb_gather_enabled(sb_handle_t h, int model_id);
int sb_gather_consume_dev(sb_handle_t h, int model_id, int64_t step);
int sb_gather_window(sb_handle_t h, int model_id, int64_t step, float** out_dev_ptr, int64_t* out_floats);
int sb_gather_collect(sb_handle_t h, int model_id, int64_t step, int B, float* out_records_host, int32_t* out_counts);
int sb_gather_status(sb_handle_t h, int model_id, int32_t* out_status, int64_t* out_steps_pushed, int64_t* out_steps_consumed);
/* With the exchange connected, sb_infer_bottomup / sb_bottomup_submit wait (on the device, behind the post-processing) for
 * every rank's records of their step and bring the WHOLE gather window to the host in the one result copy they do anyway
 * (their own outputs are the rank's slice of it).  sb_bottomup_gathered returns that copy: slot 0 / 1 after
 * sb_bottomup_collect(slot), slot -1 after sb_infer_bottomup.  out_records [world][B][width], out_counts [world]. */
int sb_bottomup_gathered(sb_handle_t h, int model_id, int slot, int B, float* out_records, int32_t* out_counts);
int sb_gather_close(sb_handle_t h, int model_id);

/* sleap/nn/inference.py:1229-1380 SingleInstanceInferenceLayer.call and :1969-2200
 * FindInstancePeaks.call: net -> global peaks -> * output_stride -> (/input_scale + 0.5)
 * (+ crop_offsets / input_scale when crop_offsets_host != NULL).
 * out_points (B,n_nodes,2), out_vals (B,n_nodes). */
typedef struct sb_global_params {
  int32_t cms_buffer, offsets_buffer;
  int32_t output_stride;
  float peak_threshold;
  int32_t refinement, integral_patch_size;
  float input_scale;
} sb_global_params;
int sb_global_configure(sb_handle_t h, int model_id, const sb_global_params* params);
int sb_infer_global(sb_handle_t h, int model_id, const void* images_host, int images_are_u8,
                    int B, const float* crop_offsets_host, float* out_points, float* out_vals);

/* sleap/nn/inference.py:1638-1966 CentroidCrop.call: net -> local peaks -> * output_stride ->
 * /input_scale + 0.5; returns flat centroid list ordered by (sample, y, x) like the reference
 * (max_instances / top_k selection and cropping are separate calls: sb_crop_centered). */
typedef struct sb_centroid_params {
  int32_t cms_buffer, offsets_buffer;
  int32_t output_stride;
  float peak_threshold;
  int32_t refinement, integral_patch_size;
  float input_scale;
  int32_t max_peaks_per_sample;
} sb_centroid_params;
int sb_centroid_configure(sb_handle_t h, int model_id, const sb_centroid_params* params);
int sb_infer_centroids(sb_handle_t h, int model_id, const void* images_host, int images_are_u8,
                       int B, float* out_centroids, float* out_vals, int32_t* out_sample_inds,
                       int32_t* out_n, int32_t* out_flags);

/* sleap/nn/inference.py:2273-2311 TopDownInferenceModel.call = CentroidCrop.call (:1747-1966) -> FindInstancePeaks.call
 * (:2059-2200), as ONE device pipeline: frames are uploaded once, the centroid peaks, the per-frame top-k
 * (tf.math.top_k(max_instances), :1879-1894), the crops (crop_bboxes on the resident frames, :1918-1927) and the
 * centered-instance network + global peaks (+ crop offsets) never leave the GPU; results come back in one copy.
 * Outputs are dense and NaN padded: centroids (B,K,2), centroid_vals (B,K), instance_peaks (B,K,n_nodes,2),
 * instance_peak_vals (B,K,n_nodes) with K = max_centroids_per_frame; n_valid (B); flags (B).  The instance network is
 * configured for max_crops_per_call crops of crop_size x crop_size and runs as often as the batch's crop count needs.
 * Not covered (use the stage-level calls): instance models trained at an input scale != 1 (pre-crop resize). */
typedef struct sb_topdown_params {
  int32_t centroid_model, instance_model;
  sb_centroid_params centroid;      /* as sb_centroid_configure */
  sb_global_params instance;        /* as sb_global_configure */
  int32_t crop_size;
  int32_t max_instances;            /* top-k per frame by centroid confidence; <= 0: keep every centroid */
  int32_t max_centroids_per_frame;  /* K */
  int32_t max_crops_per_call;       /* batch the instance network is planned for */
} sb_topdown_params;
int sb_topdown_configure(sb_handle_t h, const sb_topdown_params* params, int max_batch, int H, int W, int C_in);
int sb_infer_topdown(sb_handle_t h, int centroid_model_id, const void* frames_host, int frames_are_u8, int B,
                     float* out_centroids, float* out_centroid_vals, float* out_instance_peaks,
                     float* out_instance_peak_vals, int32_t* out_n_valid, int32_t* out_flags);

#ifdef __cplusplus
}
#endif
#endif /* SLEAP_B200_H_ */
