// Internal model structures (op-list, buffers) shared by sb_model.cu and sb_conv_tc.cu.
// The int32 record layout must match sleap_b200/nn/oplist.py.
#pragma once
#include <vector>

#include "sb_common.cuh"

enum {
  SB_OPK_BUFFER = 0,
  SB_OPK_CONV = 1,
  SB_OPK_TCONV = 2,
  SB_OPK_POOL = 3,
  SB_OPK_UPSAMPLE = 4,
  SB_OPK_ADD = 5,
  SB_OPK_PREPROCESS = 6,
  SB_OPK_COPY = 7,
};
enum { SB_OPF_RELU = 1, SB_OPF_BN = 2, SB_OPF_BILINEAR = 8, SB_OPF_FUSED_POOL = 16 };

struct SbOp {
  int32_t w[SB_OP_WORDS];
  int kind() const { return w[0]; }
  int in_buf() const { return w[1]; }
  int in_coff() const { return w[2]; }
  int in_C() const { return w[3]; }
  int in2_buf() const { return w[4]; }
  int in2_coff() const { return w[5]; }
  int out_buf() const { return w[6]; }
  int out_coff() const { return w[7]; }
  int out_C() const { return w[8]; }
  int k() const { return w[9]; }
  int stride() const { return w[10]; }
  int flags() const { return w[11]; }
  int w_off() const { return w[12]; }
  int b_off() const { return w[13]; }
  int bn_scale_off() const { return w[14]; }
  int bn_shift_off() const { return w[15]; }
  // CONV: w[18] = buffer of a fused 2x2 max-pool output (-1: none), w[19] = its channel offset;
  //       the POOL op that follows carries flag SB_OPF_FUSED_POOL and is skipped when the conv
  //       ran on the tensor-core path.
  int pool_buf() const { return w[18]; }
  int pool_coff() const { return w[19]; }
  // PREPROCESS: w[16] = float bits of input_scale, w[17] = pad_to_stride
  float input_scale() const { float f; memcpy(&f, &w[16], 4); return f; }
  int pad_stride() const { return w[17]; }
};

struct SbBuffer {
  int id = 0, stride_den = 1, C = 0, f32 = 0, is_input = 0;
  int H = 0, W = 0;
  void* dev = nullptr;
};

struct SbConvTcPlan;  // sb_conv_tc.cu
struct SbConv01Plan;  // sb_conv01.cu
struct SbTopdown;     // sb_topdown.cu

struct SbModel {
  int precision = 0;  // 0: fp16 activations + tensor-core convs; 1: fp32 CUDA-core path
  std::vector<SbOp> ops;
  std::vector<SbBuffer> buffers;
  std::vector<float> weights_host;
  float* weights_dev = nullptr;
  void* weights_tc_dev = nullptr;   // fp16 [tap][Cout][Cin] copies for the tensor-core path
  int64_t n_weights = 0;
  bool configured = false;
  int B = 0, Hin = 0, Win = 0, Cin = 0, Hres = 0, Wres = 0, Hnet = 0, Wnet = 0;
  size_t act_bytes = 0;
  void* frames_dev = nullptr;
  std::vector<SbConvTcPlan*> tc_plans;  // per op (nullptr = direct path)
  std::vector<char> skip_op;            // POOL ops fused into the producing tensor-core conv
  std::vector<cudaEvent_t> prof_events; // non-empty only inside sb_model_profile_ops
  std::vector<cudaEvent_t> fwd_events;  // sb_model_forward_times: (start, end) pairs around every forward pass
  int fwd_n = 0;                        // pairs recorded since the last read
  bool fwd_timing = false;
  // predictors
  SbPostWs ws;
  sb_bottomup_params bu{};
  std::vector<int> bu_edges;
  bool bu_configured = false;
  int guard_op = -1;
  // double-buffered asynchronous pipeline (sb_bottomup_submit / sb_bottomup_collect)
  void* frames_slot[2] = {nullptr, nullptr};
  float* stage_host[2] = {nullptr, nullptr};     // pinned result staging (per-frame records)
  int slot_B[2] = {0, 0}, rec_B = 0;            // frames of the batch last staged in each slot / in rec_host
  float* rec_host = nullptr;                     // pinned staging of the synchronous sb_infer_bottomup
  cudaEvent_t h2d_done_ev[2] = {nullptr, nullptr}, frames_free_ev[2] = {nullptr, nullptr}, result_ev[2] = {nullptr, nullptr};
  bool slot_used[2] = {false, false};
  cudaStream_t copy_stream = nullptr;      // first op that overwrites a head buffer the post-processing stream may still read
  sb_global_params gl{};
  bool gl_configured = false;
  float *gpart = nullptr, *gpoints = nullptr, *gvals = nullptr, *crop_off_dev = nullptr;
  int g_rpc = 1, g_chunks = 1;
  sb_centroid_params ce{};
  bool ce_configured = false;
  bool td_configured = false;              // fused top-down pipeline (sb_topdown_configure); state lives on the centroid model
  SbTopdown* td = nullptr;
  SbConv01Plan* conv01 = nullptr;          // fused first encoder block (frame -> conv0 -> conv1 -> pool), sb_conv01.cu
  bool conv01_enabled = false;             // the autotuner measured it faster than the two separate launches
  SbGather gather;                         // peer-memory exchange of the result records (sb_gather.cu)
  bool keep_dead_stores = false;           // sb_model_forward asked for a tensor whose stores are normally elided
};

int sb_run_ops(sb_handle_s* h, SbModel* m, const void* frames_dev, int frames_are_u8, int B);

void sb_topdown_free(SbModel* m);        // sb_topdown.cu

// record exchange (sb_gather.cu)
SbGatherDev sb_gather_dev(const SbModel* m, unsigned long long step);
void sb_gather_free(SbModel* m);
// queues on `s`: wait for every rank's records of `step`, copy the [world][B][width] window to host_dst, acknowledge
int sb_gather_queue_collect(sb_handle_s* h, SbModel* m, long long step, int B, float* host_dst, int* counts_dev, cudaStream_t s);
void sb_pipeline_slots_free(SbModel* m);

// tensor-core conv path (sb_conv_tc.cu)
int sb_conv_tc_prepare(sb_handle_s* h, SbModel* m);      // after buffers are allocated
void sb_conv_tc_release(SbModel* m);
bool sb_conv_tc_can(const SbModel* m, int op_index);
bool sb_conv_tc_out_dead(const SbModel* m, int buffer_id);   // its stores are elided unless keep_dead_stores is set
int sb_conv_tc_launch(sb_handle_s* h, SbModel* m, int op_index, int B);

// first conv fused with the PREPROCESS op before it (sb_model.cu): conv op index or -1
int sb_first_fusion_op(const SbModel* m, size_t pre_index);
// first layer as a Toeplitz GEMM on the stock tcgen05 conv kernels (sb_conv_tc.cu)
bool sb_first_view_can(const SbModel* m, int op_index);
int sb_first_view_launch(sb_handle_s* h, SbModel* m, int op_index, const void* frames_dev, int frames_are_u8, int B);
bool sb_first_buffer_view_can(const SbModel* m, int op_index);
int sb_first_buffer_view_launch(sb_handle_s* h, SbModel* m, int op_index, int B);
int sb_first_direct_launch(sb_handle_s* h, SbModel* m, int op_index, const void* frames_dev, int frames_are_u8, int B);

// 7x7 stride-2 stem through a space-to-depth view of the frame (sb_conv_tc.cu); conv op index or -1 (sb_model.cu)
int sb_stem_fusion_op(const SbModel* m, size_t pre_index);
bool sb_stem_view_can(const SbModel* m, int op_index);
int sb_stem_view_launch(sb_handle_s* h, SbModel* m, int op_index, const void* frames_dev, int frames_are_u8, int B);

// fused first encoder block (sb_conv01.cu)
int sb_conv01_prepare(sb_handle_s* h, SbModel* m, int conv0_op, int conv1_op, bool conv1_out_dead);
void sb_conv01_release(SbModel* m);
bool sb_conv01_can(const SbModel* m, int conv0_op);
int sb_conv01_conv1_op(const SbModel* m);
int sb_conv01_launch(sb_handle_s* h, SbModel* m, const void* frames_dev, int frames_are_u8, int B);

// first layer fused with preprocessing on the tensor cores (sb_conv_tc.cu)
bool sb_conv_first_tc_ok(const SbModel* m, const SbOp& conv);
int sb_conv_first_tc_launch(sb_handle_s* h, SbModel* m, const SbOp& op, const void* frames_dev, int frames_are_u8, int B);
