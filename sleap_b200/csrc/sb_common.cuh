// Shared definitions for libsleapb200 (sm_100a).  Internal header; the public C-ABI is
// include/sleap_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/sleap_b200.h"

struct SbModel;

// Device workspace for the post-processing stages (capacity-bounded, per sample).
struct SbPostWs {
  int B = 0, H = 0, W = 0, C = 0;          // confidence-map shape the workspace was sized for
  int rows_per_chunk = 0, n_chunks = 0, chunk_cap = 0;
  int max_peaks = 0, max_node_peaks = 0, max_instances = 0, n_edges = 0;
  // local peaks
  int* chunk_cnt = nullptr;                 // [B][n_chunks]
  uint2* chunk_items = nullptr;             // [B][n_chunks][chunk_cap]  (flat idx, val bits)
  float* peaks = nullptr;                   // [B][max_peaks][2]   (x, y) already scaled
  float* peak_vals = nullptr;               // [B][max_peaks]
  int* peak_ch = nullptr;                   // [B][max_peaks]
  int* n_peaks = nullptr;                   // [B]
  int* total_peaks = nullptr;               // [B]   (uncapped count)
  int* node_cnt = nullptr;                  // [B][C]  (uncapped count per node)
  int* node_peaks = nullptr;                // [B][C][max_node_peaks]  peak index within sample
  // scoring / matching
  float* score_mat = nullptr;               // [B][E][K*K]
  int* match_cnt = nullptr;                 // [B][E]
  int* match_src = nullptr;                 // [B][E][K]
  int* match_dst = nullptr;                 // [B][E][K]
  float* match_score = nullptr;             // [B][E][K]
  // grouping output
  float* inst_peaks = nullptr;              // [B][max_instances][C][2]
  float* inst_vals = nullptr;               // [B][max_instances][C]
  float* inst_scores = nullptr;             // [B][max_instances]
  int* n_inst = nullptr;                    // [B]
  int* flags = nullptr;                     // [B]  overflow bit flags
  // contiguous per-frame result records written by the grouping kernel's epilogue:
  // [B][I*C*2 peaks | I*C vals | I scores | n_valid | flags]  (sb_record_width floats per frame)
  float* records = nullptr;
  uint2* sorted_items = nullptr;            // [B][max_peaks]  scanned items in tf.where order (k_local_emit scratch)
  int* edges_dev = nullptr;                 // [E][2]
  int* sorted_edges_dev = nullptr;          // [n_sorted]
  int n_sorted = 0;
  size_t bytes = 0;
};

// ---- peer-memory record exchange (sb_gather.cu; pushed from k_group's epilogue) ----
#define SB_GATHER_MAX_WORLD 8
#define SB_GATHER_TIMEOUT_ARRIVE 1
#define SB_GATHER_TIMEOUT_ACK 2
struct SbGatherDev {                        // by-value kernel argument: one step's view of the exchange
  int on, rank, world, G, Bmax;
  size_t width;
  unsigned long long step, timeout_ns;
  float* data[SB_GATHER_MAX_WORLD];               // every rank's window (own rank: local memory, others: NVLink peer mappings)
  unsigned long long* arrive[SB_GATHER_MAX_WORLD];
  unsigned long long* ack[SB_GATHER_MAX_WORLD];
  unsigned int* done;
  int* status;
};
struct SbGather {
  int rank = 0, world = 0, G = 0, Bmax = 0;
  size_t width = 0;
  void* local = nullptr;
  void* peer[SB_GATHER_MAX_WORLD] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool connected = false;
  long long step = 0;                              // steps pushed so far
  long long consumed = 0;                          // steps acknowledged so far (acks are cumulative)
  unsigned long long timeout_ns = 0;
  int *status_host = nullptr, *status_dev = nullptr, *counts_host = nullptr, *counts_dev = nullptr;
};

static __host__ __device__ inline size_t sb_record_width(int max_instances, int n_nodes) {
  // peaks | values | scores | n_valid | flags, padded to a multiple of 4 floats (16-byte rows for vector copies)
  return (((size_t)max_instances * n_nodes * 3 + max_instances + 2) + 3) & ~(size_t)3;
}

struct sb_handle_s {
  int device = 0;
  cudaStream_t stream = nullptr;       // stream all work is issued on
  cudaStream_t own_stream = nullptr;   // created by sb_create
  cudaStream_t aux_stream[3] = {nullptr, nullptr, nullptr};  // fork/join branches (tconv phases)
  cudaEvent_t fork_ev = nullptr, join_ev[3] = {nullptr, nullptr, nullptr};
  // post-processing of step i runs on its own stream so that it overlaps the network of step i+1
  cudaStream_t post_stream = nullptr;
  cudaEvent_t fwd_done_ev = nullptr, post_done_ev = nullptr;
  bool post_pending = false;
  std::string last_error;
  std::vector<void*> owned;                 // generic device allocations freed at destroy
  std::vector<SbModel*> models;
  int gpu_launches = 0;                     // kernels launched by this handle (bench: gpu_launches)
  int sm_count = 148;
};

extern thread_local std::string g_sb_last_error;

int sb_fail(sb_handle_s* h, int code, const char* fmt, ...);
void sb_models_free(sb_handle_s* h);

#define SB_CUDA(h, expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess)                                                            \
      return sb_fail((h), SB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                     cudaGetErrorString(_e), __FILE__, __LINE__);                     \
  } while (0)

#define SB_CHECK_LAUNCH(h) do { (h)->gpu_launches++; SB_CUDA((h), cudaGetLastError()); } while (0)

template <typename T>
static inline int sb_dev_alloc(sb_handle_s* h, T** p, size_t n) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, n * sizeof(T) + 16);
  if (e != cudaSuccess) return sb_fail(h, SB_ERR_CUDA, "cudaMalloc(%zu) failed: %s", n * sizeof(T), cudaGetErrorString(e));
  *p = (T*)q;
  return 0;
}

// ---- post-processing launchers (sb_post.cu) ----
struct SbPeakParams {
  float threshold;
  int refinement;      // SB_REFINE_*
  int patch;           // integral patch size (odd)
  float scale;         // multiply refined peaks (cm output stride); 1 for the stage-level API
  float input_scale;   // != 1: then / input_scale + 0.5 (centroid / single-instance paths)
};

int sbk_local_peaks(sb_handle_s* h, const void* cms, int cms_is_half, const float* offsets,
                    int B, int H, int W, int C, const SbPeakParams& p, SbPostWs& ws);
int sbk_global_peaks(sb_handle_s* h, const void* cms, int cms_is_half, const float* offsets,
                     int B, int H, int W, int C, const SbPeakParams& p,
                     const float* crop_off_dev, float* part_buf, int n_chunks, int rows_per_chunk,
                     float* out_points, float* out_vals);
int sbk_score_match(sb_handle_s* h, const float* pafs, int B, int Hp, int Wp, int C2,
                    int n_points, int pafs_stride, float max_edge_length, float dist_penalty_weight,
                    SbPostWs& ws);
int sbk_group(sb_handle_s* h, int B, int n_nodes, int min_instance_peaks, float min_line_scores,
              float input_scale, SbPostWs& ws, const SbGatherDev* gather = nullptr);
int sbk_lsap_batch(sb_handle_s* h, const float* scores, const int* n_src, const int* n_dst,
                   const int* offsets, int n_problems, int max_k, int* out_rows, int* out_cols,
                   float* out_scores, int* out_counts);
int sbk_crop(sb_handle_s* h, const void* images, int img_is_u8, int B, int H, int W, int C,
             const float* centroids, const int* sample_inds, int n, int crop_h, int crop_w,
             void* out, int out_is_u8_trunc);

int sbk_lines(sb_handle_s* h, const float* pafs, int Hp, int Wp, int C2, const float* lines_in,
              const float* peaks, const int* edge_peak_inds, const int* edge_inds, int n, int P,
              float pafs_stride, float max_edge_length, float dist_w, int* out_subs, float* out_lines,
              float* out_scores);
int sbk_integral(sb_handle_s* h, const float* cms, int N, int Hh, int Ww, int C, const float* xv,
                 const float* yv, float* x_hat, float* y_hat);
int sbk_local_dir(sb_handle_s* h, const float* patches, int N, float delta, float* out);

int sb_post_ws_alloc(sb_handle_s* h, SbPostWs& ws, int B, int H, int W, int C, int max_peaks,
                     int max_node_peaks, int max_instances, int n_edges);
void sb_post_ws_free(SbPostWs& ws);
