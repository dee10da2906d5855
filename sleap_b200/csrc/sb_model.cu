// Model engine: executes the flat op-list compiled from the reference's backbone + heads graph,
// and the fused predictors (bottom-up / single-instance / centered-instance / centroid) on top.
//
// Reference: sleap/nn/model.py:312-364 (graph topology), sleap/nn/inference.py:2864-3003
// (BottomUpInferenceLayer), :1319-1380 (SingleInstanceInferenceLayer), :1747-1966 (CentroidCrop),
// :2059-2200 (FindInstancePeaks).
#include <algorithm>

#include "sb_common.cuh"
#include "sb_kernels_direct.cuh"
#include "sb_model.h"

using namespace sbd;

namespace {

template <typename T> T* buf_ptr(const SbBuffer& b) { return (T*)b.dev; }

size_t elem_size(const SbModel* m, const SbBuffer& b) { return (b.f32 || m->precision == 1) ? 4 : 2; }

int grid_for(size_t total, int sm) {
  size_t g = (total + 255) / 256;
  size_t cap = (size_t)sm * 16;
  return (int)std::max<size_t>(1, std::min(g, cap));
}

}  // namespace

void sb_models_free(sb_handle_s* h) {
  for (SbModel* m : h->models) {
    if (!m) continue;
    for (auto& b : m->buffers) if (b.dev) cudaFree(b.dev);
    if (m->weights_dev) cudaFree(m->weights_dev);
    if (m->weights_tc_dev) cudaFree(m->weights_tc_dev);
    if (m->frames_dev) cudaFree(m->frames_dev);
    if (m->crop_off_dev) cudaFree(m->crop_off_dev);
    if (m->gpart) cudaFree(m->gpart);
    if (m->gpoints) cudaFree(m->gpoints);
    if (m->gvals) cudaFree(m->gvals);
    if (m->rec_host) cudaFreeHost(m->rec_host);
    for (int i = 0; i < 2; ++i) {
      if (m->frames_slot[i]) cudaFree(m->frames_slot[i]);
      if (m->stage_host[i]) cudaFreeHost(m->stage_host[i]);
      if (m->h2d_done_ev[i]) cudaEventDestroy(m->h2d_done_ev[i]);
      if (m->frames_free_ev[i]) cudaEventDestroy(m->frames_free_ev[i]);
      if (m->result_ev[i]) cudaEventDestroy(m->result_ev[i]);
    }
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    for (auto& e : m->fwd_events) cudaEventDestroy(e);
    sb_post_ws_free(m->ws);
    sb_gather_free(m);
    sb_topdown_free(m);
    sb_conv01_release(m);
    sb_conv_tc_release(m);
    delete m;
  }
  h->models.clear();
}

// pinned staging + device frame slots of the submit/collect pipeline are sized from (B, H, W, C, max_instances,
// n_nodes) at first use: any configure call that may change one of those drops them (re-created lazily)
void sb_pipeline_slots_free(SbModel* m) {
  for (int i = 0; i < 2; ++i) {
    if (m->frames_slot[i]) { cudaFree(m->frames_slot[i]); m->frames_slot[i] = nullptr; }
    if (m->stage_host[i]) { cudaFreeHost(m->stage_host[i]); m->stage_host[i] = nullptr; }
    m->slot_used[i] = false;
  }
  if (m->rec_host) { cudaFreeHost(m->rec_host); m->rec_host = nullptr; }
}

static SbModel* get_model(sb_handle_s* h, int id) {
  if (!h || id < 0 || id >= (int)h->models.size()) return nullptr;
  return h->models[id];
}

extern "C" {

int sb_load_model(sb_handle_t h, const int32_t* ops, int n_ops, const float* weights, int64_t n_weights,
                  int precision, int* out_model_id) {
  if (!h) return sb_fail(nullptr, SB_ERR_INVALID, "null handle");
  if (!ops || n_ops <= 0 || !weights || n_weights <= 0 || !out_model_id)
    return sb_fail(h, SB_ERR_INVALID, "sb_load_model: bad arguments");
  if (precision < 0 || precision > 2) return sb_fail(h, SB_ERR_INVALID, "precision must be 0 (fp16), 1 (fp32 CUDA cores) or 2 (split fp16 on the tensor cores)");
  SB_CUDA(h, cudaSetDevice(h->device));
  SbModel* m = new SbModel();
  m->precision = precision;
  m->n_weights = n_weights;
  for (int i = 0; i < n_ops; ++i) {
    const int32_t* w = ops + (size_t)i * SB_OP_WORDS;
    if (w[0] == SB_OPK_BUFFER) {
      SbBuffer b;
      b.id = w[1]; b.stride_den = w[2]; b.C = w[3]; b.f32 = w[4]; b.is_input = w[5];
      if (b.id != (int)m->buffers.size()) { delete m; return sb_fail(h, SB_ERR_INVALID, "buffer ids must be dense and ordered"); }
      if (b.stride_den <= 0 || b.C <= 0) { delete m; return sb_fail(h, SB_ERR_INVALID, "bad buffer record"); }
      m->buffers.push_back(b);
    } else {
      SbOp op;
      memcpy(op.w, w, sizeof(op.w));
      const int nb = (int)m->buffers.size();
      auto okbuf = [&](int id) { return id >= 0 && id < nb; };
      if (op.kind() < SB_OPK_CONV || op.kind() > SB_OPK_COPY) { delete m; return sb_fail(h, SB_ERR_INVALID, "op %d: unknown kind %d", i, op.kind()); }
      if (!okbuf(op.out_buf()) || (op.kind() != SB_OPK_PREPROCESS && !okbuf(op.in_buf()))) { delete m; return sb_fail(h, SB_ERR_INVALID, "op %d: bad buffer id", i); }
      if (op.kind() == SB_OPK_ADD && !okbuf(op.in2_buf())) { delete m; return sb_fail(h, SB_ERR_INVALID, "op %d: bad second input", i); }
      auto okoff = [&](int off, int64_t n) { return off < 0 ? true : (int64_t)off + n <= n_weights; };
      if (op.kind() == SB_OPK_CONV || op.kind() == SB_OPK_TCONV) {
        const int64_t nw = (int64_t)op.k() * op.k() * op.in_C() * op.out_C();
        if (op.w_off() < 0 || !okoff(op.w_off(), nw) || !okoff(op.b_off(), op.out_C()) ||
            !okoff(op.bn_scale_off(), op.out_C()) || !okoff(op.bn_shift_off(), op.out_C())) {
          delete m; return sb_fail(h, SB_ERR_INVALID, "op %d: weight offsets out of range", i);
        }
      }
      m->ops.push_back(op);
    }
  }
  if (m->buffers.empty() || m->ops.empty()) { delete m; return sb_fail(h, SB_ERR_INVALID, "empty model"); }
  cudaError_t e = cudaMalloc((void**)&m->weights_dev, (size_t)n_weights * sizeof(float));
  if (e != cudaSuccess) { delete m; return sb_fail(h, SB_ERR_CUDA, "cudaMalloc weights: %s", cudaGetErrorString(e)); }
  e = cudaMemcpy(m->weights_dev, weights, (size_t)n_weights * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(m->weights_dev); delete m; return sb_fail(h, SB_ERR_CUDA, "copy weights: %s", cudaGetErrorString(e)); }
  m->weights_host.assign(weights, weights + n_weights);
  h->models.push_back(m);
  *out_model_id = (int)h->models.size() - 1;
  return SB_OK;
}

int sb_model_configure(sb_handle_t h, int model_id, int max_batch, int H, int W, int C_in) {
  SbModel* m = get_model(h, model_id);
  if (!m) return sb_fail(h, SB_ERR_INVALID, "bad model id");
  if (max_batch <= 0 || H <= 0 || W <= 0 || (C_in != 1 && C_in != 3)) return sb_fail(h, SB_ERR_INVALID, "sb_model_configure: bad shape");
  SB_CUDA(h, cudaSetDevice(h->device));
  // preprocess op defines the net input size
  const SbOp* pre = nullptr;
  for (auto& op : m->ops) if (op.kind() == SB_OPK_PREPROCESS) { pre = &op; break; }
  if (!pre) return sb_fail(h, SB_ERR_INVALID, "model has no preprocess op");
  const float input_scale = pre->input_scale();
  const int pad_stride = std::max(1, pre->pad_stride());
  int Hres = H, Wres = W;
  if (input_scale != 1.0f) { Wres = (int)((float)W * input_scale); Hres = (int)((float)H * input_scale); }
  const int Hnet = ((Hres + pad_stride - 1) / pad_stride) * pad_stride;
  const int Wnet = ((Wres + pad_stride - 1) / pad_stride) * pad_stride;
  for (auto& b : m->buffers) {
    if (Hnet % b.stride_den || Wnet % b.stride_den)
      return sb_fail(h, SB_ERR_INVALID, "net input %dx%d not divisible by stride %d (pad_to_stride too small)", Hnet, Wnet, b.stride_den);
  }
  // a reconfigure invalidates everything sized from the old shape: drain the device first, then drop the
  // activation buffers, the double-buffer pipeline slots and the predictor workspaces (their configure
  // calls must be repeated; the C-ABI refuses to run a predictor on a stale workspace)
  SB_CUDA(h, cudaDeviceSynchronize());
  h->post_pending = false;
  for (auto& b : m->buffers) { if (b.dev) { cudaFree(b.dev); b.dev = nullptr; } }
  if (m->frames_dev) { cudaFree(m->frames_dev); m->frames_dev = nullptr; }
  sb_pipeline_slots_free(m);
  sb_gather_free(m);                             // window sizes depend on (B, max_instances, n_nodes): re-init after a reconfigure
  m->configured = false;
  m->bu_configured = false; m->gl_configured = false; m->ce_configured = false; m->td_configured = false;
  sb_conv01_release(m);
  m->conv01_enabled = false;
  sb_conv_tc_release(m);
  m->B = max_batch; m->Hin = H; m->Win = W; m->Cin = C_in; m->Hres = Hres; m->Wres = Wres; m->Hnet = Hnet; m->Wnet = Wnet;
  size_t total = 0;
  for (auto& b : m->buffers) {
    b.H = Hnet / b.stride_den; b.W = Wnet / b.stride_den;
    const size_t bytes = (size_t)max_batch * b.H * b.W * b.C * elem_size(m, b);
    cudaError_t e = cudaMalloc(&b.dev, bytes + 256);
    if (e != cudaSuccess) return sb_fail(h, SB_ERR_CUDA, "cudaMalloc buffer %d (%zu B): %s", b.id, bytes, cudaGetErrorString(e));
    total += bytes;
  }
  m->act_bytes = total;
  SB_CUDA(h, cudaMalloc(&m->frames_dev, (size_t)max_batch * H * W * C_in * sizeof(float)));
  int rc = sb_conv_tc_prepare(h, m);
  if (rc) return rc;
  m->configured = true;
  return SB_OK;
}

}  // extern "C"

// First conv fused with preprocessing (fp16 path): returns the conv op index or -1.
static int first_fusion_op(const SbModel* m, size_t pre_index) { return sb_first_fusion_op(m, pre_index); }
int sb_first_fusion_op(const SbModel* m, size_t pre_index) {
  if (m->precision == 1 || getenv("SB_DISABLE_FIRST_FUSION")) return -1;
  if (pre_index + 1 >= m->ops.size()) return -1;
  const SbOp& pre = m->ops[pre_index];
  const SbOp& cv = m->ops[pre_index + 1];
  if (cv.kind() != SB_OPK_CONV || cv.in_buf() != pre.out_buf() || cv.k() != 3 || cv.stride() != 1) return -1;
  if (pre.input_scale() != 1.0f) return -1;
  const SbBuffer& ib = m->buffers[pre.out_buf()];
  const SbBuffer& ob = m->buffers[cv.out_buf()];
  if (m->Cin != ib.C || (ib.C != 1 && ib.C != 3) || cv.in_C() != ib.C) return -1;
  if (ob.f32 || (cv.flags() & SB_OPF_BN) || ob.C % 8 || cv.out_coff() % 8) return -1;
  const int co = cv.out_C();
  if (!(co == 8 || co == 16 || co == 24 || co == 32 || co == 64)) return -1;
  for (size_t i = pre_index + 2; i < m->ops.size(); ++i)      // nobody else may read the preprocessed frame
    if (m->ops[i].kind() != SB_OPK_PREPROCESS && (m->ops[i].in_buf() == pre.out_buf() ||
        (m->ops[i].kind() == SB_OPK_ADD && m->ops[i].in2_buf() == pre.out_buf()))) return -1;
  return (int)pre_index + 1;
}

// 7x7 stride-2 stem right after PREPROCESS (hourglass.py:49-100) with 1 / 3 input channels and no resize: the conv op
// index when the tensor-core space-to-depth form can take it (the PREPROCESS op is then fused into the view kernel).
int sb_stem_fusion_op(const SbModel* m, size_t pre_index) {
  if (m->precision != 0 || pre_index + 1 >= m->ops.size()) return -1;
  const SbOp& pre = m->ops[pre_index];
  const SbOp& cv = m->ops[pre_index + 1];
  if (cv.kind() != SB_OPK_CONV || cv.in_buf() != pre.out_buf() || cv.k() != 7 || cv.stride() != 2) return -1;
  if (pre.input_scale() != 1.0f) return -1;
  const SbBuffer& ib = m->buffers[pre.out_buf()];
  if (m->Cin != ib.C || (ib.C != 1 && ib.C != 3) || cv.in_C() != ib.C) return -1;
  for (size_t i = pre_index + 2; i < m->ops.size(); ++i)      // nobody else may read the preprocessed frame
    if (m->ops[i].kind() != SB_OPK_PREPROCESS && (m->ops[i].in_buf() == pre.out_buf() ||
        (m->ops[i].kind() == SB_OPK_ADD && m->ops[i].in2_buf() == pre.out_buf()))) return -1;
  return (int)pre_index + 1;
}

template <typename TI, int CIN>
static void launch_first(int co, int B, cudaStream_t s, const TI* img, int Hin, int Win, int Hnet, int Wnet, __half* out,
                         int Ctot, int coff, const float* w, const float* b, int relu, int is_u8, int split) {
  dim3 blk(32, 8);
  auto grid = [&](int px) { return dim3((Wnet + 32 * px - 1) / (32 * px), (Hnet + 7) / 8, B); };
  switch (co) {
    case 8: k_conv_first<TI, CIN, 8, 4><<<grid(4), blk, 0, s>>>(img, Hin, Win, Hnet, Wnet, out, Ctot, coff, w, b, relu, is_u8, split); break;
    case 16: k_conv_first<TI, CIN, 16, 4><<<grid(4), blk, 0, s>>>(img, Hin, Win, Hnet, Wnet, out, Ctot, coff, w, b, relu, is_u8, split); break;
    case 24: k_conv_first<TI, CIN, 24, 2><<<grid(2), blk, 0, s>>>(img, Hin, Win, Hnet, Wnet, out, Ctot, coff, w, b, relu, is_u8, split); break;
    case 32: k_conv_first<TI, CIN, 32, 2><<<grid(2), blk, 0, s>>>(img, Hin, Win, Hnet, Wnet, out, Ctot, coff, w, b, relu, is_u8, split); break;
    default: k_conv_first<TI, CIN, 64, 1><<<grid(1), blk, 0, s>>>(img, Hin, Win, Hnet, Wnet, out, Ctot, coff, w, b, relu, is_u8, split); break;
  }
}

// CUDA-core first layer fused with preprocessing (k_conv_first): the fallback of, and the timing rival
// to, the Toeplitz tensor-core form (sb_first_view_launch).
int sb_first_direct_launch(sb_handle_s* h, SbModel* m, int op_index, const void* frames_dev, int frames_are_u8, int B) {
  const SbOp& op = m->ops[op_index];
  SbBuffer& ob = m->buffers[op.out_buf()];
  cudaStream_t s = h->stream;
  const float* Wt = m->weights_dev + op.w_off();
  const float* bias = op.b_off() >= 0 ? m->weights_dev + op.b_off() : nullptr;
  const int g = B;
  const int relu = (op.flags() & SB_OPF_RELU) ? 1 : 0;
  const int split = m->precision == 2 ? op.out_C() : 0;
  // rows/cols beyond the resized frame (Hres, Wres) are the bottom/right zero padding
  if (frames_are_u8) {
    if (m->Cin == 1) launch_first<unsigned char, 1>(op.out_C(), g, s, (const unsigned char*)frames_dev, m->Hin, m->Win, ob.H, ob.W, (__half*)ob.dev, ob.C, op.out_coff(), Wt, bias, relu, 1, split);
    else launch_first<unsigned char, 3>(op.out_C(), g, s, (const unsigned char*)frames_dev, m->Hin, m->Win, ob.H, ob.W, (__half*)ob.dev, ob.C, op.out_coff(), Wt, bias, relu, 1, split);
  } else {
    if (m->Cin == 1) launch_first<float, 1>(op.out_C(), g, s, (const float*)frames_dev, m->Hin, m->Win, ob.H, ob.W, (__half*)ob.dev, ob.C, op.out_coff(), Wt, bias, relu, 0, split);
    else launch_first<float, 3>(op.out_C(), g, s, (const float*)frames_dev, m->Hin, m->Win, ob.H, ob.W, (__half*)ob.dev, ob.C, op.out_coff(), Wt, bias, relu, 0, split);
  }
  SB_CHECK_LAUNCH(h);
  return 0;
}

// ------------------------------------------------------------------------------------------
template <typename T>
static int run_ops_t(sb_handle_s* h, SbModel* m, const void* frames_dev, int frames_are_u8, int B) {
  cudaStream_t s = h->stream;
  const bool split = m->precision == 2;         // split-fp16 activations: [lo | hi | hi] channel planes (sb_kernels_direct.cuh)
  int fused_first = -1, fused_conv1 = -1, fused_stem = -1;
  for (size_t oi = 0; oi < m->ops.size(); ++oi) {
    const SbOp& op = m->ops[oi];
    if (!m->prof_events.empty()) cudaEventRecord(m->prof_events[oi], s);
    SbBuffer& ob = m->buffers[op.out_buf()];
    if ((int)oi == m->guard_op && h->post_pending) SB_CUDA(h, cudaStreamWaitEvent(s, h->post_done_ev, 0));
    if (oi < m->skip_op.size() && m->skip_op[oi]) continue;     // 2x2 max-pool fused into the producing conv
    if ((int)oi == fused_conv1) continue;                       // ran inside the fused first block
    if ((int)oi == fused_stem) {                                // 7x7 s2 stem: frame -> space-to-depth view -> tcgen05
      int rc = sb_stem_view_launch(h, m, (int)oi, frames_dev, frames_are_u8, B);
      if (rc) return rc;
      continue;
    }
    if ((int)oi == fused_first && sb_conv01_can(m, (int)oi) && !m->keep_dead_stores) {
      int rc = sb_conv01_launch(h, m, frames_dev, frames_are_u8, B);
      if (rc) return rc;
      fused_conv1 = sb_conv01_conv1_op(m);
      continue;
    }
    if ((int)oi == fused_first && sb_first_view_can(m, (int)oi)) {
      int rc = sb_first_view_launch(h, m, (int)oi, frames_dev, frames_are_u8, B);
      if (rc) return rc;
      continue;
    }
    if ((int)oi == fused_first && sb_conv_first_tc_ok(m, op)) {
      int rc = sb_conv_first_tc_launch(h, m, op, frames_dev, frames_are_u8, B);
      if (rc) return rc;
      continue;
    }
    if ((int)oi == fused_first) {
      int rc = sb_first_direct_launch(h, m, (int)oi, frames_dev, frames_are_u8, B);
      if (rc) return rc;
      continue;
    }
    switch (op.kind()) {
      case SB_OPK_PREPROCESS: {
        if (sizeof(T) == 2 && (fused_first = first_fusion_op(m, oi)) >= 0) break;
        if (sizeof(T) == 2 && sb_stem_view_can(m, sb_stem_fusion_op(m, oi))) { fused_stem = sb_stem_fusion_op(m, oi); break; }
        const size_t total = (size_t)B * ob.H * ob.W * ob.C;
        const int resize = op.input_scale() != 1.0f;
        int mode_ch = 0;
        if (m->Cin == 3 && ob.C == 1) mode_ch = 1;
        if (m->Cin == 1 && ob.C == 3) mode_ch = 2;
        if (ob.f32 && sizeof(T) == 2) {          // precision 2 keeps the preprocessed frame in fp32
          if (frames_are_u8)
            k_preprocess<unsigned char, float><<<grid_for(total, h->sm_count), 256, 0, s>>>(
                (const unsigned char*)frames_dev, m->Hin, m->Win, m->Cin, (float*)ob.dev, ob.H, ob.W, ob.C, m->Hres, m->Wres, resize, mode_ch, 1, total);
          else
            k_preprocess<float, float><<<grid_for(total, h->sm_count), 256, 0, s>>>(
                (const float*)frames_dev, m->Hin, m->Win, m->Cin, (float*)ob.dev, ob.H, ob.W, ob.C, m->Hres, m->Wres, resize, mode_ch, 0, total);
        } else if (frames_are_u8)
          k_preprocess<unsigned char, T><<<grid_for(total, h->sm_count), 256, 0, s>>>(
              (const unsigned char*)frames_dev, m->Hin, m->Win, m->Cin, (T*)ob.dev, ob.H, ob.W, ob.C, m->Hres, m->Wres, resize, mode_ch, 1, total);
        else
          k_preprocess<float, T><<<grid_for(total, h->sm_count), 256, 0, s>>>(
              (const float*)frames_dev, m->Hin, m->Win, m->Cin, (T*)ob.dev, ob.H, ob.W, ob.C, m->Hres, m->Wres, resize, mode_ch, 0, total);
        SB_CHECK_LAUNCH(h);
        break;
      }
      case SB_OPK_CONV: {
        SbBuffer& ib = m->buffers[op.in_buf()];
        if (m->precision == 0 && sb_first_buffer_view_can(m, (int)oi)) {
          int rc = sb_first_buffer_view_launch(h, m, (int)oi, B);
          if (rc) return rc;
          break;
        }
        if (m->precision != 1 && sb_conv_tc_can(m, (int)oi)) {
          int rc = sb_conv_tc_launch(h, m, (int)oi, B);
          if (rc) return rc;
          break;
        }
        const int k = op.k(), st = op.stride();
        const int osplit = (split && !ob.f32) ? op.out_C() : 0;
        const int Hout = ob.H, Wout = ob.W;
        const int tot_h = std::max((Hout - 1) * st + k - ib.H, 0), tot_w = std::max((Wout - 1) * st + k - ib.W, 0);
        const int pad_top = tot_h / 2, pad_left = tot_w / 2;
        const float* W = m->weights_dev + op.w_off();
        const float* bias = op.b_off() >= 0 ? m->weights_dev + op.b_off() : nullptr;
        const float* bs = (op.flags() & SB_OPF_BN) ? m->weights_dev + op.bn_scale_off() : nullptr;
        const float* bh = (op.flags() & SB_OPF_BN) ? m->weights_dev + op.bn_shift_off() : nullptr;
        const int in_tile = (DC_TILE - 1) * st + k;
        const size_t sm = ((size_t)in_tile * in_tile * DC_CK + (size_t)k * k * DC_CK * DC_CO) * sizeof(float);
        dim3 g(((Wout + DC_TILE - 1) / DC_TILE) * ((Hout + DC_TILE - 1) / DC_TILE), (op.out_C() + DC_CO - 1) / DC_CO, B);
        const int relu = (op.flags() & SB_OPF_RELU) ? 1 : 0;
        if (ib.f32 && sizeof(T) == 2) {              // precision 2: first conv straight from the fp32 preprocessed frame
          if (ob.f32) return sb_fail(h, SB_ERR_INVALID, "conv from the fp32 input buffer to an fp32 head is not supported in precision 2");
          auto kern = k_conv_direct<float, __half>;
          if (sm > 48 * 1024) SB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
          kern<<<g, 256, sm, s>>>((const float*)ib.dev, ib.H, ib.W, ib.C, op.in_coff(), op.in_C(), (__half*)ob.dev, Hout, Wout,
                                  ob.C, op.out_coff(), op.out_C(), W, bias, bs, bh, k, st, pad_top, pad_left, relu, osplit);
        } else if (ob.f32 && sizeof(T) == 2) {
          auto kern = k_conv_direct<T, float>;
          if (sm > 48 * 1024) SB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
          kern<<<g, 256, sm, s>>>((const T*)ib.dev, ib.H, ib.W, ib.C, op.in_coff(), op.in_C(), (float*)ob.dev, Hout, Wout,
                                  ob.C, op.out_coff(), op.out_C(), W, bias, bs, bh, k, st, pad_top, pad_left, relu, 0);
        } else {
          auto kern = k_conv_direct<T, T>;
          if (sm > 48 * 1024) SB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
          kern<<<g, 256, sm, s>>>((const T*)ib.dev, ib.H, ib.W, ib.C, op.in_coff(), op.in_C(), (T*)ob.dev, Hout, Wout,
                                  ob.C, op.out_coff(), op.out_C(), W, bias, bs, bh, k, st, pad_top, pad_left, relu, osplit);
        }
        SB_CHECK_LAUNCH(h);
        break;
      }
      case SB_OPK_TCONV: {
        SbBuffer& ib = m->buffers[op.in_buf()];
        if (m->precision != 1 && sb_conv_tc_can(m, (int)oi)) {
          int rc = sb_conv_tc_launch(h, m, (int)oi, B);
          if (rc) return rc;
          break;
        }
        const float* W = m->weights_dev + op.w_off();
        const float* bias = op.b_off() >= 0 ? m->weights_dev + op.b_off() : nullptr;
        dim3 g((ob.H * ob.W + 255) / 256, (op.out_C() + DC_CO - 1) / DC_CO, B);
        k_tconv_direct<T, T><<<g, 256, 0, s>>>((const T*)ib.dev, ib.H, ib.W, ib.C, op.in_coff(), op.in_C(), (T*)ob.dev, ob.C,
                                               op.out_coff(), op.out_C(), W, bias, (op.flags() & SB_OPF_RELU) ? 1 : 0, split ? op.out_C() : 0);
        SB_CHECK_LAUNCH(h);
        break;
      }
      case SB_OPK_POOL: {
        SbBuffer& ib = m->buffers[op.in_buf()];
        // precision 2: the record carries the physical channel count (3C); the split kernels work on logical channels
        const size_t total = (size_t)B * ob.H * ob.W * (split ? op.in_C() / 3 : op.in_C());
        if (split)
          k_maxpool2_split<<<grid_for(total, h->sm_count), 256, 0, s>>>((const __half*)ib.dev, ib.H, ib.W, ib.C, op.in_coff(), op.in_C() / 3,
                                                                        (__half*)ob.dev, ob.H, ob.W, ob.C, op.out_coff(), total);
        else
        k_maxpool2<T><<<grid_for(total, h->sm_count), 256, 0, s>>>((const T*)ib.dev, ib.H, ib.W, ib.C, op.in_coff(), op.in_C(),
                                                                   (T*)ob.dev, ob.H, ob.W, ob.C, op.out_coff(), total);
        SB_CHECK_LAUNCH(h);
        break;
      }
      case SB_OPK_UPSAMPLE: {
        SbBuffer& ib = m->buffers[op.in_buf()];
        const size_t total = (size_t)B * ob.H * ob.W * (split ? op.in_C() / 3 : op.in_C());
        if (split)
          k_upsample2_split<<<grid_for(total, h->sm_count), 256, 0, s>>>((const __half*)ib.dev, ib.H, ib.W, ib.C, op.in_coff(), op.in_C() / 3,
                                                                         (__half*)ob.dev, ob.C, op.out_coff(),
                                                                         (op.flags() & SB_OPF_BILINEAR) ? 1 : 0, total);
        else
        k_upsample2<T><<<grid_for(total, h->sm_count), 256, 0, s>>>((const T*)ib.dev, ib.H, ib.W, ib.C, op.in_coff(), op.in_C(),
                                                                    (T*)ob.dev, ob.C, op.out_coff(),
                                                                    (op.flags() & SB_OPF_BILINEAR) ? 1 : 0, total);
        SB_CHECK_LAUNCH(h);
        break;
      }
      case SB_OPK_ADD: {
        SbBuffer& ib = m->buffers[op.in_buf()];
        SbBuffer& ib2 = m->buffers[op.in2_buf()];
        const size_t npix = (size_t)B * ob.H * ob.W;
        if (split)
          k_add_split<<<grid_for(npix * (op.in_C() / 3), h->sm_count), 256, 0, s>>>((const __half*)ib.dev, ib.C, op.in_coff(), (const __half*)ib2.dev,
                                                                                     ib2.C, op.in2_coff(), (__half*)ob.dev, ob.C, op.out_coff(),
                                                                                     op.in_C() / 3, npix);
        else
        k_add<T><<<grid_for(npix * op.in_C(), h->sm_count), 256, 0, s>>>((const T*)ib.dev, ib.C, op.in_coff(), (const T*)ib2.dev, ib2.C,
                                                                         op.in2_coff(), (T*)ob.dev, ob.C, op.out_coff(), op.in_C(), npix);
        SB_CHECK_LAUNCH(h);
        break;
      }
      case SB_OPK_COPY: {
        SbBuffer& ib = m->buffers[op.in_buf()];
        const size_t npix = (size_t)B * ob.H * ob.W;
        k_copy<T><<<grid_for(npix * op.in_C(), h->sm_count), 256, 0, s>>>((const T*)ib.dev, ib.C, op.in_coff(), (T*)ob.dev, ob.C,
                                                                          op.out_coff(), op.in_C(), npix);
        SB_CHECK_LAUNCH(h);
        break;
      }
      default:
        return sb_fail(h, SB_ERR_INVALID, "unknown op kind %d", op.kind());
    }
  }
  if (!m->prof_events.empty()) cudaEventRecord(m->prof_events[m->ops.size()], s);
  return 0;
}

int sb_run_ops(sb_handle_s* h, SbModel* m, const void* frames_dev, int frames_are_u8, int B) {
  if (!m->configured) return sb_fail(h, SB_ERR_INVALID, "model not configured");
  if (B <= 0 || B > m->B) return sb_fail(h, SB_ERR_INVALID, "batch %d exceeds configured max %d", B, m->B);
  const bool timed = m->fwd_timing && 2 * (m->fwd_n + 1) <= (int)m->fwd_events.size();
  if (timed) cudaEventRecord(m->fwd_events[2 * m->fwd_n], h->stream);
  const int rc = m->precision == 1 ? run_ops_t<float>(h, m, frames_dev, frames_are_u8, B)
                                   : run_ops_t<__half>(h, m, frames_dev, frames_are_u8, B);
  if (timed) { cudaEventRecord(m->fwd_events[2 * m->fwd_n + 1], h->stream); ++m->fwd_n; }
  return rc;
}

__global__ void k_half_to_float(const __half* __restrict__ in, float* __restrict__ out, size_t n) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x)
    out[t] = __half2float(in[t]);
}

static int upload_frames(sb_handle_s* h, SbModel* m, const void* images_host, int is_u8, int B) {
  const size_t bytes = (size_t)B * m->Hin * m->Win * m->Cin * (is_u8 ? 1 : 4);
  SB_CUDA(h, cudaMemcpyAsync(m->frames_dev, images_host, bytes, cudaMemcpyHostToDevice, h->stream));
  return 0;
}

extern "C" {

int sb_model_forward(sb_handle_t h, int model_id, const void* images_host, int images_are_u8, int B,
                     int n_outputs, const int32_t* output_buffer_ids, float** out_host_ptrs) {
  SbModel* m = get_model(h, model_id);
  if (!m) return sb_fail(h, SB_ERR_INVALID, "bad model id");
  SB_CUDA(h, cudaSetDevice(h->device));
  if (!m->configured) return sb_fail(h, SB_ERR_INVALID, "model not configured");
  if (B <= 0 || B > m->B) return sb_fail(h, SB_ERR_INVALID, "bad batch");
  int rc = upload_frames(h, m, images_host, images_are_u8, B);
  if (rc) return rc;
  for (int i = 0; i < n_outputs; ++i) {                  // a tensor nobody reads inside the graph is only written on request
    if (output_buffer_ids[i] >= 0 && sb_conv_tc_out_dead(m, output_buffer_ids[i])) m->keep_dead_stores = true;
    if (m->conv01 && output_buffer_ids[i] >= 0 && output_buffer_ids[i] < (int)m->buffers.size() &&
        (output_buffer_ids[i] == m->ops[sb_conv01_conv1_op(m)].in_buf() || output_buffer_ids[i] == m->ops[sb_conv01_conv1_op(m)].out_buf()))
      m->keep_dead_stores = true;                          // tensors internal to the fused first block
  }
  rc = sb_run_ops(h, m, m->frames_dev, images_are_u8, B);
  m->keep_dead_stores = false;
  if (rc) return rc;
  for (int i = 0; i < n_outputs; ++i) {
    const int id = output_buffer_ids[i];
    if (id < 0 || id >= (int)m->buffers.size()) return sb_fail(h, SB_ERR_INVALID, "bad output buffer id %d", id);
    SbBuffer& b = m->buffers[id];
    const size_t n = (size_t)B * b.H * b.W * b.C;
    if (elem_size(m, b) == 4) {
      SB_CUDA(h, cudaMemcpyAsync(out_host_ptrs[i], b.dev, n * 4, cudaMemcpyDeviceToHost, h->stream));
    } else {
      float* tmp = nullptr;
      SB_CUDA(h, cudaMalloc((void**)&tmp, n * 4));
      k_half_to_float<<<grid_for(n, h->sm_count), 256, 0, h->stream>>>((const __half*)b.dev, tmp, n);
      h->gpu_launches++;
      cudaError_t e = cudaMemcpyAsync(out_host_ptrs[i], tmp, n * 4, cudaMemcpyDeviceToHost, h->stream);
      cudaStreamSynchronize(h->stream);
      cudaFree(tmp);
      if (e != cudaSuccess) return sb_fail(h, SB_ERR_CUDA, "copy out: %s", cudaGetErrorString(e));
    }
  }
  SB_CUDA(h, cudaStreamSynchronize(h->stream));
  return SB_OK;
}

// Per-op device timing of one forward pass (CUDA events on the launching stream); used by
// bench.py for the roofline of the conv kernels.  out_ms[i] = duration of op i; out_kind[i] =
// 0 other, 1 tensor-core conv, 2 CUDA-core conv; out_flops[i] = 2*MACs of the op for batch B.
int sb_model_profile_ops(sb_handle_t h, int model_id, const uint8_t* frames_dev, int B, int cap, float* out_ms,
                         int32_t* out_kind, double* out_flops, int32_t* out_n_ops) {
  SbModel* m = get_model(h, model_id);
  if (!m || !m->configured) return sb_fail(h, SB_ERR_INVALID, "model not configured");
  SB_CUDA(h, cudaSetDevice(h->device));
  const int n = (int)m->ops.size();
  if (cap < n) return sb_fail(h, SB_ERR_INVALID, "capacity %d < n_ops %d", cap, n);
  m->prof_events.resize(n + 1);
  for (auto& e : m->prof_events) SB_CUDA(h, cudaEventCreate(&e));
  int rc = sb_run_ops(h, m, frames_dev, 1, B);
  cudaStreamSynchronize(h->stream);
  for (int i = 0; i < n && !rc; ++i) {
    cudaEventElapsedTime(&out_ms[i], m->prof_events[i], m->prof_events[i + 1]);
    const SbOp& op = m->ops[i];
    out_kind[i] = 0; out_flops[i] = 0.0;
    if (op.kind() == SB_OPK_CONV || op.kind() == SB_OPK_TCONV) {
      out_kind[i] = sb_conv_tc_can(m, i) ? 1 : 2;
      const SbBuffer& ib = m->buffers[op.in_buf()];
      const SbBuffer& ob = m->buffers[op.out_buf()];
      const double pix = op.kind() == SB_OPK_TCONV ? (double)ib.H * ib.W : (double)ob.H * ob.W;
      out_flops[i] = 2.0 * op.k() * op.k() * op.in_C() * op.out_C() * pix * B;
    }
  }
  for (auto& e : m->prof_events) cudaEventDestroy(e);
  m->prof_events.clear();
  *out_n_ops = n;
  return rc;
}

int sb_model_forward_times(sb_handle_t h, int model_id, int enable, int cap, float* out_ms, int32_t* out_n) {
  SbModel* m = get_model(h, model_id);
  if (!m) return sb_fail(h, SB_ERR_INVALID, "bad model id");
  SB_CUDA(h, cudaSetDevice(h->device));
  SB_CUDA(h, cudaStreamSynchronize(h->stream));
  int n = 0;
  for (; n < m->fwd_n && n < cap && out_ms; ++n)
    SB_CUDA(h, cudaEventElapsedTime(&out_ms[n], m->fwd_events[2 * n], m->fwd_events[2 * n + 1]));
  if (out_n) *out_n = n;
  m->fwd_n = 0;
  if (enable && m->fwd_events.empty()) {
    m->fwd_events.resize(2 * 1024);
    for (auto& e : m->fwd_events) SB_CUDA(h, cudaEventCreate(&e));
  }
  m->fwd_timing = enable != 0;
  return SB_OK;
}

// ---------------------------------- bottom-up ------------------------------------------------
int sb_bottomup_configure(sb_handle_t h, int model_id, const sb_bottomup_params* p) {
  SbModel* m = get_model(h, model_id);
  if (!m || !p) return sb_fail(h, SB_ERR_INVALID, "bad model id / params");
  if (!m->configured) return sb_fail(h, SB_ERR_INVALID, "call sb_model_configure first");
  SB_CUDA(h, cudaSetDevice(h->device));
  const int nb = (int)m->buffers.size();
  if (p->cms_buffer < 0 || p->cms_buffer >= nb || p->pafs_buffer < 0 || p->pafs_buffer >= nb)
    return sb_fail(h, SB_ERR_INVALID, "cms/pafs buffer ids out of range");
  SbBuffer& cb = m->buffers[p->cms_buffer];
  SbBuffer& pb = m->buffers[p->pafs_buffer];
  if (!cb.f32 || !pb.f32) return sb_fail(h, SB_ERR_INVALID, "cms / pafs buffers must be f32 head outputs");
  if (cb.C != p->n_nodes || pb.C != 2 * p->n_edges) return sb_fail(h, SB_ERR_INVALID, "head channels do not match skeleton");
  if (p->offsets_buffer >= 0 && (p->offsets_buffer >= nb || !m->buffers[p->offsets_buffer].f32 || m->buffers[p->offsets_buffer].C != 2 * p->n_nodes))
    return sb_fail(h, SB_ERR_INVALID, "bad offsets buffer");
  if (p->n_sorted > p->n_edges || p->max_peaks_per_sample <= 0 || p->max_node_peaks <= 0 || p->max_instances <= 0)
    return sb_fail(h, SB_ERR_INVALID, "bad capacities");
  for (int e = 0; e < 2 * p->n_edges; ++e)
    if (p->edges[e] < 0 || p->edges[e] >= p->n_nodes) return sb_fail(h, SB_ERR_INVALID, "edge node index out of range");
  SB_CUDA(h, cudaDeviceSynchronize());          // in-flight post-processing / result copies still use the old workspace
  h->post_pending = false;
  m->bu_configured = false;
  sb_pipeline_slots_free(m);                      // staging records are sized from max_instances / n_nodes
  sb_gather_free(m);
  sb_post_ws_free(m->ws);
  int rc = sb_post_ws_alloc(h, m->ws, m->B, cb.H, cb.W, cb.C, p->max_peaks_per_sample, p->max_node_peaks, p->max_instances, p->n_edges);
  if (rc) return rc;
  SB_CUDA(h, cudaMemcpy(m->ws.edges_dev, p->edges, (size_t)p->n_edges * 2 * sizeof(int), cudaMemcpyHostToDevice));
  if (p->n_sorted > 0)
    SB_CUDA(h, cudaMemcpy(m->ws.sorted_edges_dev, p->sorted_edge_inds, (size_t)p->n_sorted * sizeof(int), cudaMemcpyHostToDevice));
  m->ws.n_sorted = p->n_sorted;
  m->bu = *p;
  m->bu.edges = nullptr; m->bu.sorted_edge_inds = nullptr;
  m->bu_edges.assign(p->edges, p->edges + 2 * p->n_edges);
  m->guard_op = -1;
  if (!getenv("SB_DISABLE_POST_OVERLAP"))
    for (size_t i = 0; i < m->ops.size(); ++i) {
      const int ob = m->ops[i].out_buf();
      if (ob == p->cms_buffer || ob == p->pafs_buffer || (p->offsets_buffer >= 0 && ob == p->offsets_buffer)) { m->guard_op = (int)i; break; }
    }
  m->bu_configured = true;
  return SB_OK;
}

static size_t stage_floats(const SbModel* m) {          // [world][B][width] when the exchange is connected
  return (size_t)(m->gather.connected ? m->gather.world : 1) * m->B * sb_record_width(m->bu.max_instances, m->bu.n_nodes);
}

// One D2H copy of a batch's results into pinned `dst`: the rank's own records, or -- exchange connected -- the whole gather
// window of the step just pushed (device-side wait for the peers, copy, acknowledge; all on stream rs).
static int queue_result_copy(sb_handle_s* h, SbModel* m, int B, cudaStream_t rs, float* dst, int counts_slot) {
  const size_t w = sb_record_width(m->bu.max_instances, m->bu.n_nodes);
  if (m->gather.connected)
    return sb_gather_queue_collect(h, m, m->gather.step - 1, B, dst, m->gather.counts_dev + counts_slot * SB_GATHER_MAX_WORLD, rs);
  SB_CUDA(h, cudaMemcpyAsync(dst, m->ws.records, (size_t)B * w * sizeof(float), cudaMemcpyDeviceToHost, rs));
  return 0;
}
static const float* own_slice(const SbModel* m, const float* staged, int B) {
  return staged + (m->gather.connected ? (size_t)m->gather.rank * B * sb_record_width(m->bu.max_instances, m->bu.n_nodes) : 0);
}
static int check_exchange(sb_handle_s* h, SbModel* m) {
  if (m->gather.connected && *m->gather.status_host != 0) {
    const int st = *m->gather.status_host;
    *m->gather.status_host = 0;
    return sb_fail(h, SB_ERR_CUDA, "record exchange timed out (%s): a peer rank stopped pushing or consuming",
                   st == SB_GATHER_TIMEOUT_ARRIVE ? "waiting for arrivals" : "waiting for acknowledgements");
  }
  return 0;
}

static void unpack_records(const SbModel* m, const float* rec, int B, float* out_instance_peaks, float* out_instance_peak_vals,
                           float* out_instance_scores, int32_t* out_n_valid, int32_t* out_flags) {
  const size_t I = m->bu.max_instances, C = m->bu.n_nodes, w = sb_record_width((int)I, (int)C);
  for (int b = 0; b < B; ++b) {
    const float* r = rec + (size_t)b * w;
    memcpy(out_instance_peaks + (size_t)b * I * C * 2, r, I * C * 2 * sizeof(float));
    memcpy(out_instance_peak_vals + (size_t)b * I * C, r + I * C * 2, I * C * sizeof(float));
    memcpy(out_instance_scores + (size_t)b * I, r + I * C * 3, I * sizeof(float));
    out_n_valid[b] = (int32_t)r[I * C * 3 + I];
    if (out_flags) out_flags[b] = (int32_t)r[I * C * 3 + I + 1];
  }
}

static int bottomup_post_kernels(sb_handle_s* h, SbModel* m, int B) {
  const sb_bottomup_params& p = m->bu;
  SbBuffer& cb = m->buffers[p.cms_buffer];
  SbBuffer& pb = m->buffers[p.pafs_buffer];
  const float* off = p.offsets_buffer >= 0 ? (const float*)m->buffers[p.offsets_buffer].dev : nullptr;
  SbPeakParams pp{p.peak_threshold, p.refinement, p.integral_patch_size, (float)p.cm_output_stride, 1.0f};
  int rc = sbk_local_peaks(h, cb.dev, 0, off, B, cb.H, cb.W, cb.C, pp, m->ws);
  if (rc) return rc;
  const float max_len = p.max_edge_length_ratio * (float)std::max(std::max(pb.H, pb.W), pb.C) * (float)p.paf_output_stride;
  if ((rc = sbk_score_match(h, (const float*)pb.dev, B, pb.H, pb.W, pb.C, p.n_line_points, p.paf_output_stride, max_len,
                            p.dist_penalty_weight, m->ws))) return rc;
  if (m->gather.connected) {               // fused exchange: k_group's epilogue pushes the records to every peer
    const SbGatherDev gx = sb_gather_dev(m, (unsigned long long)m->gather.step);
    rc = sbk_group(h, B, p.n_nodes, p.min_instance_peaks, p.min_line_scores, p.input_scale, m->ws, &gx);
    if (!rc) m->gather.step++;
    return rc;
  }
  return sbk_group(h, B, p.n_nodes, p.min_instance_peaks, p.min_line_scores, p.input_scale, m->ws);
}

// Peak finding / PAF scoring / matching / grouping of this batch on the handle's post-processing
// stream: it only depends on the head outputs, so it overlaps the network of the next batch
// (run_ops waits on post_done_ev right before it overwrites a head buffer).
static int bottomup_post(sb_handle_s* h, SbModel* m, int B) {
  if (m->guard_op < 0) return bottomup_post_kernels(h, m, B);
  cudaStream_t main_stream = h->stream;
  SB_CUDA(h, cudaEventRecord(h->fwd_done_ev, main_stream));
  SB_CUDA(h, cudaStreamWaitEvent(h->post_stream, h->fwd_done_ev, 0));
  h->stream = h->post_stream;
  int rc = bottomup_post_kernels(h, m, B);
  cudaError_t e = cudaEventRecord(h->post_done_ev, h->post_stream);
  h->stream = main_stream;
  if (rc) return rc;
  if (e != cudaSuccess) return sb_fail(h, SB_ERR_CUDA, "event record: %s", cudaGetErrorString(e));
  h->post_pending = true;
  return 0;
}

int sb_infer_bottomup_dev(sb_handle_t h, int model_id, const uint8_t* frames_dev, int B) {
  SbModel* m = get_model(h, model_id);
  if (!m || !m->bu_configured) return sb_fail(h, SB_ERR_INVALID, "bottom-up predictor not configured");
  SB_CUDA(h, cudaSetDevice(h->device));
  int rc = sb_run_ops(h, m, frames_dev, 1, B);
  if (rc) return rc;
  return bottomup_post(h, m, B);
}

int sb_infer_bottomup(sb_handle_t h, int model_id, const uint8_t* frames_host, int B, float* out_instance_peaks,
                      float* out_instance_peak_vals, float* out_instance_scores, int32_t* out_n_valid,
                      int32_t* out_flags) {
  SbModel* m = get_model(h, model_id);
  if (!m || !m->bu_configured) return sb_fail(h, SB_ERR_INVALID, "bottom-up predictor not configured");
  SB_CUDA(h, cudaSetDevice(h->device));
  if (B <= 0 || B > m->B) return sb_fail(h, SB_ERR_INVALID, "bad batch");
  int rc = upload_frames(h, m, frames_host, 1, B);
  if (rc) return rc;
  if ((rc = sb_run_ops(h, m, m->frames_dev, 1, B))) return rc;
  if ((rc = bottomup_post(h, m, B))) return rc;
  cudaStream_t rs = h->post_pending ? h->post_stream : h->stream;
  if (!m->rec_host) SB_CUDA(h, cudaHostAlloc((void**)&m->rec_host, stage_floats(m) * sizeof(float), cudaHostAllocDefault));
  if ((rc = queue_result_copy(h, m, B, rs, m->rec_host, 3))) return rc;
  SB_CUDA(h, cudaStreamSynchronize(rs));
  h->post_pending = false;
  m->rec_B = B;
  if ((rc = check_exchange(h, m))) return rc;
  unpack_records(m, own_slice(m, m->rec_host, B), B, out_instance_peaks, out_instance_peak_vals, out_instance_scores, out_n_valid, out_flags);
  return SB_OK;
}

// Makes the handle's main stream wait (on the device, no host sync) for the post-processing of the
// last sb_infer_bottomup_dev call, e.g. before recording an end-of-work event or reading the results.
int sb_bottomup_wait_results(sb_handle_t h, int model_id) {
  SbModel* m = get_model(h, model_id);
  if (!m || !m->bu_configured) return sb_fail(h, SB_ERR_INVALID, "bottom-up predictor not configured");
  if (h->post_pending) SB_CUDA(h, cudaStreamWaitEvent(h->stream, h->post_done_ev, 0));
  return SB_OK;
}

// The stream post-processing runs on (consumers such as the NCCL gather can be queued behind it).
int sb_get_post_stream(sb_handle_t h, void** out_stream) {
  if (!h || !out_stream) return sb_fail(h, SB_ERR_INVALID, "null argument");
  *out_stream = (void*)h->post_stream;
  return SB_OK;
}

// Splits the per-frame result records (written by k_group's epilogue, ONE D2H copy per batch) into the caller's arrays.
// Asynchronous, double-buffered variant of sb_infer_bottomup for streaming many batches: submit
// batch i+1 (its H2D copy runs on a copy stream) while batch i computes, then collect batch i.
// Layout of the pinned staging record per slot: peaks | vals | scores | n_valid | flags.

int sb_bottomup_submit(sb_handle_t h, int model_id, const uint8_t* frames_host, int B, int slot) {
  SbModel* m = get_model(h, model_id);
  if (!m || !m->bu_configured) return sb_fail(h, SB_ERR_INVALID, "bottom-up predictor not configured");
  if (slot < 0 || slot > 1 || B <= 0 || B > m->B) return sb_fail(h, SB_ERR_INVALID, "bad slot / batch");
  SB_CUDA(h, cudaSetDevice(h->device));
  if (!m->copy_stream) {
    SB_CUDA(h, cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      SB_CUDA(h, cudaEventCreateWithFlags(&m->h2d_done_ev[i], cudaEventDisableTiming));
      SB_CUDA(h, cudaEventCreateWithFlags(&m->frames_free_ev[i], cudaEventDisableTiming));
      SB_CUDA(h, cudaEventCreateWithFlags(&m->result_ev[i], cudaEventDisableTiming));
    }
  }
  const size_t fbytes = (size_t)m->B * m->Hin * m->Win * m->Cin;
  for (int i = 0; i < 2; ++i) {
    if (!m->frames_slot[i]) SB_CUDA(h, cudaMalloc(&m->frames_slot[i], fbytes));
    if (!m->stage_host[i]) SB_CUDA(h, cudaHostAlloc((void**)&m->stage_host[i], stage_floats(m) * sizeof(float), cudaHostAllocDefault));
  }
  // H2D on the copy stream (after the network that last read this slot's frames has finished)
  if (m->slot_used[slot]) SB_CUDA(h, cudaStreamWaitEvent(m->copy_stream, m->frames_free_ev[slot], 0));
  SB_CUDA(h, cudaMemcpyAsync(m->frames_slot[slot], frames_host, (size_t)B * m->Hin * m->Win * m->Cin, cudaMemcpyHostToDevice, m->copy_stream));
  SB_CUDA(h, cudaEventRecord(m->h2d_done_ev[slot], m->copy_stream));
  SB_CUDA(h, cudaStreamWaitEvent(h->stream, m->h2d_done_ev[slot], 0));
  int rc = sb_run_ops(h, m, m->frames_slot[slot], 1, B);
  if (rc) return rc;
  SB_CUDA(h, cudaEventRecord(m->frames_free_ev[slot], h->stream));
  if ((rc = bottomup_post(h, m, B))) return rc;
  cudaStream_t rs = h->post_pending ? h->post_stream : h->stream;
  if ((rc = queue_result_copy(h, m, B, rs, m->stage_host[slot], 1 + slot))) return rc;
  m->slot_B[slot] = B;
  SB_CUDA(h, cudaEventRecord(m->result_ev[slot], rs));
  m->slot_used[slot] = true;
  return SB_OK;
}

int sb_bottomup_collect(sb_handle_t h, int model_id, int slot, int B, float* out_instance_peaks,
                        float* out_instance_peak_vals, float* out_instance_scores, int32_t* out_n_valid,
                        int32_t* out_flags) {
  SbModel* m = get_model(h, model_id);
  if (!m || !m->bu_configured) return sb_fail(h, SB_ERR_INVALID, "bottom-up predictor not configured");
  if (slot < 0 || slot > 1 || !m->slot_used[slot] || B <= 0 || B > m->B) return sb_fail(h, SB_ERR_INVALID, "bad slot / batch");
  SB_CUDA(h, cudaEventSynchronize(m->result_ev[slot]));
  int rc = check_exchange(h, m);
  if (rc) return rc;
  unpack_records(m, own_slice(m, m->stage_host[slot], B), B, out_instance_peaks, out_instance_peak_vals, out_instance_scores, out_n_valid, out_flags);
  return SB_OK;
}

int sb_bottomup_gathered(sb_handle_t h, int model_id, int slot, int B, float* out_records, int32_t* out_counts) {
  SbModel* m = get_model(h, model_id);
  if (!m || !m->bu_configured || !m->gather.connected) return sb_fail(h, SB_ERR_INVALID, "sb_bottomup_gathered: exchange not connected");
  if (slot < -1 || slot > 1 || !out_records) return sb_fail(h, SB_ERR_INVALID, "sb_bottomup_gathered: bad slot");
  const float* src = slot < 0 ? m->rec_host : m->stage_host[slot];
  const int have = slot < 0 ? m->rec_B : m->slot_B[slot];
  if (!src || have != B) return sb_fail(h, SB_ERR_INVALID, "sb_bottomup_gathered: no collected batch of %d frames in that slot", B);
  memcpy(out_records, src, (size_t)m->gather.world * B * sb_record_width(m->bu.max_instances, m->bu.n_nodes) * sizeof(float));
  if (out_counts)
    for (int r = 0; r < m->gather.world; ++r) out_counts[r] = m->gather.counts_host[(slot < 0 ? 3 : 1 + slot) * SB_GATHER_MAX_WORLD + r];
  return SB_OK;
}

int sb_bottomup_device_outputs(sb_handle_t h, int model_id, float** instance_peaks_dev, float** instance_peak_vals_dev,
                               float** instance_scores_dev, int32_t** n_valid_dev, int32_t** flags_dev) {
  SbModel* m = get_model(h, model_id);
  if (!m || !m->bu_configured) return sb_fail(h, SB_ERR_INVALID, "bottom-up predictor not configured");
  if (instance_peaks_dev) *instance_peaks_dev = m->ws.inst_peaks;
  if (instance_peak_vals_dev) *instance_peak_vals_dev = m->ws.inst_vals;
  if (instance_scores_dev) *instance_scores_dev = m->ws.inst_scores;
  if (n_valid_dev) *n_valid_dev = m->ws.n_inst;
  if (flags_dev) *flags_dev = m->ws.flags;
  return SB_OK;
}

int sb_bottomup_device_records(sb_handle_t h, int model_id, float** records_dev) {
  SbModel* m = get_model(h, model_id);
  if (!m || !m->bu_configured || !records_dev) return sb_fail(h, SB_ERR_INVALID, "bottom-up predictor not configured");
  *records_dev = m->ws.records;
  return SB_OK;
}

static int fetch_graph_ws(sb_handle_s* h, SbPostWs& ws, const int* edges_host, int B, int cap_peaks, float* peaks,
                          float* peak_vals, int32_t* peak_channel_inds, int32_t* peak_offsets, int cap_cands,
                          int32_t* edge_inds, int32_t* edge_peak_inds, float* line_scores, int32_t* cand_offsets) {
  const int C = ws.C, K = ws.max_node_peaks, E = ws.n_edges, MP = ws.max_peaks;
  std::vector<int> np(B), ncnt((size_t)B * C), nlist((size_t)B * C * K);
  std::vector<float> mat((size_t)B * E * K * K);
  SB_CUDA(h, cudaStreamSynchronize(h->stream));
  SB_CUDA(h, cudaStreamSynchronize(h->post_stream));
  SB_CUDA(h, cudaMemcpy(np.data(), ws.n_peaks, (size_t)B * 4, cudaMemcpyDeviceToHost));
  SB_CUDA(h, cudaMemcpy(ncnt.data(), ws.node_cnt, ncnt.size() * 4, cudaMemcpyDeviceToHost));
  SB_CUDA(h, cudaMemcpy(nlist.data(), ws.node_peaks, nlist.size() * 4, cudaMemcpyDeviceToHost));
  SB_CUDA(h, cudaMemcpy(mat.data(), ws.score_mat, mat.size() * 4, cudaMemcpyDeviceToHost));
  int tp = 0, tc = 0;
  for (int b = 0; b < B; ++b) {
    peak_offsets[b] = tp;
    cand_offsets[b] = tc;
    if (tp + np[b] > cap_peaks) return sb_fail(h, SB_ERR_INVALID, "peak capacity exceeded");
    if (np[b] > 0) {
      SB_CUDA(h, cudaMemcpy(peaks + 2 * (size_t)tp, ws.peaks + (size_t)b * MP * 2, (size_t)np[b] * 8, cudaMemcpyDeviceToHost));
      SB_CUDA(h, cudaMemcpy(peak_vals + tp, ws.peak_vals + (size_t)b * MP, (size_t)np[b] * 4, cudaMemcpyDeviceToHost));
      SB_CUDA(h, cudaMemcpy(peak_channel_inds + tp, ws.peak_ch + (size_t)b * MP, (size_t)np[b] * 4, cudaMemcpyDeviceToHost));
    }
    tp += np[b];
    for (int e = 0; e < E; ++e) {
      const int sn = edges_host[2 * e], dn = edges_host[2 * e + 1];
      const int ns = std::min(ncnt[(size_t)b * C + sn], K), nd = std::min(ncnt[(size_t)b * C + dn], K);
      for (int i = 0; i < ns; ++i)
        for (int j = 0; j < nd; ++j) {
          if (tc >= cap_cands) return sb_fail(h, SB_ERR_INVALID, "candidate capacity exceeded");
          edge_inds[tc] = e;
          edge_peak_inds[2 * tc] = nlist[((size_t)b * C + sn) * K + i];
          edge_peak_inds[2 * tc + 1] = nlist[((size_t)b * C + dn) * K + j];
          line_scores[tc] = mat[((size_t)b * E + e) * K * K + (size_t)i * nd + j];
          ++tc;
        }
    }
  }
  peak_offsets[B] = tp;
  cand_offsets[B] = tc;
  return SB_OK;
}

int sb_bottomup_fetch_graph(sb_handle_t h, int model_id, int B, int cap_peaks, float* peaks, float* peak_vals,
                            int32_t* peak_channel_inds, int32_t* peak_offsets, int cap_cands, int32_t* edge_inds,
                            int32_t* edge_peak_inds, float* line_scores, int32_t* cand_offsets) {
  SbModel* m = get_model(h, model_id);
  if (!m || !m->bu_configured) return sb_fail(h, SB_ERR_INVALID, "bottom-up predictor not configured");
  SB_CUDA(h, cudaSetDevice(h->device));
  return fetch_graph_ws(h, m->ws, m->bu_edges.data(), B, cap_peaks, peaks, peak_vals, peak_channel_inds, peak_offsets,
                        cap_cands, edge_inds, edge_peak_inds, line_scores, cand_offsets);
}

int sb_bottomup_from_maps(sb_handle_t h, const sb_bottomup_params* p, const float* cms_host, int B, int H, int W,
                          const float* pafs_host, int Hp, int Wp, const float* offsets_host,
                          float* out_instance_peaks, float* out_instance_peak_vals, float* out_instance_scores,
                          int32_t* out_n_valid, int32_t* out_flags, int cap_peaks, float* peaks, float* peak_vals,
                          int32_t* peak_channel_inds, int32_t* peak_offsets, int cap_cands, int32_t* edge_inds,
                          int32_t* edge_peak_inds, float* line_scores, int32_t* cand_offsets) {
  if (!h || !p) return sb_fail(h, SB_ERR_INVALID, "null handle / params");
  if (B <= 0 || H <= 0 || W <= 0 || Hp <= 0 || Wp <= 0 || p->n_nodes <= 0 || p->n_edges <= 0)
    return sb_fail(h, SB_ERR_INVALID, "sb_bottomup_from_maps: bad shape");
  SB_CUDA(h, cudaSetDevice(h->device));
  const int C = p->n_nodes, C2 = 2 * p->n_edges;
  SbPostWs ws;
  struct Guard { SbPostWs& w; std::vector<void*> bufs; ~Guard() { sb_post_ws_free(w); for (void* q : bufs) cudaFree(q); } } guard{ws, {}};
  int rc = sb_post_ws_alloc(h, ws, B, H, W, C, p->max_peaks_per_sample, p->max_node_peaks, p->max_instances, p->n_edges);
  if (rc) return rc;
  auto dalloc = [&](void** q, size_t bytes) -> int {
    cudaError_t e = cudaMalloc(q, bytes + 16);
    if (e != cudaSuccess) return sb_fail(h, SB_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
    guard.bufs.push_back(*q);
    return 0;
  };
  void *d_cms = nullptr, *d_pafs = nullptr, *d_off = nullptr;
  const size_t ncm = (size_t)B * H * W * C, npf = (size_t)B * Hp * Wp * C2;
  if ((rc = dalloc(&d_cms, ncm * 4)) || (rc = dalloc(&d_pafs, npf * 4))) return rc;
  SB_CUDA(h, cudaMemcpyAsync(d_cms, cms_host, ncm * 4, cudaMemcpyHostToDevice, h->stream));
  SB_CUDA(h, cudaMemcpyAsync(d_pafs, pafs_host, npf * 4, cudaMemcpyHostToDevice, h->stream));
  if (offsets_host) {
    if ((rc = dalloc(&d_off, 2 * ncm * 4))) return rc;
    SB_CUDA(h, cudaMemcpyAsync(d_off, offsets_host, 2 * ncm * 4, cudaMemcpyHostToDevice, h->stream));
  }
  SB_CUDA(h, cudaMemcpyAsync(ws.edges_dev, p->edges, (size_t)p->n_edges * 8, cudaMemcpyHostToDevice, h->stream));
  if (p->n_sorted > 0) SB_CUDA(h, cudaMemcpyAsync(ws.sorted_edges_dev, p->sorted_edge_inds, (size_t)p->n_sorted * 4, cudaMemcpyHostToDevice, h->stream));
  ws.n_sorted = p->n_sorted;
  SbPeakParams pp{p->peak_threshold, p->refinement, p->integral_patch_size, (float)p->cm_output_stride, 1.0f};
  if ((rc = sbk_local_peaks(h, d_cms, 0, (const float*)d_off, B, H, W, C, pp, ws))) return rc;
  const float max_len = p->max_edge_length_ratio * (float)std::max(std::max(Hp, Wp), C2) * (float)p->paf_output_stride;
  if ((rc = sbk_score_match(h, (const float*)d_pafs, B, Hp, Wp, C2, p->n_line_points, p->paf_output_stride, max_len,
                            p->dist_penalty_weight, ws))) return rc;
  if ((rc = sbk_group(h, B, C, p->min_instance_peaks, p->min_line_scores, p->input_scale, ws))) return rc;
  const size_t I = p->max_instances;
  SB_CUDA(h, cudaMemcpyAsync(out_instance_peaks, ws.inst_peaks, (size_t)B * I * C * 8, cudaMemcpyDeviceToHost, h->stream));
  SB_CUDA(h, cudaMemcpyAsync(out_instance_peak_vals, ws.inst_vals, (size_t)B * I * C * 4, cudaMemcpyDeviceToHost, h->stream));
  SB_CUDA(h, cudaMemcpyAsync(out_instance_scores, ws.inst_scores, (size_t)B * I * 4, cudaMemcpyDeviceToHost, h->stream));
  SB_CUDA(h, cudaMemcpyAsync(out_n_valid, ws.n_inst, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
  if (out_flags) SB_CUDA(h, cudaMemcpyAsync(out_flags, ws.flags, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
  SB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (peaks)
    return fetch_graph_ws(h, ws, p->edges, B, cap_peaks, peaks, peak_vals, peak_channel_inds, peak_offsets, cap_cands,
                          edge_inds, edge_peak_inds, line_scores, cand_offsets);
  return SB_OK;
}

// ---------------------------------- global peaks (single / centered instance) ----------------
int sb_global_configure(sb_handle_t h, int model_id, const sb_global_params* p) {
  SbModel* m = get_model(h, model_id);
  if (!m || !p) return sb_fail(h, SB_ERR_INVALID, "bad model id / params");
  if (!m->configured) return sb_fail(h, SB_ERR_INVALID, "call sb_model_configure first");
  SB_CUDA(h, cudaSetDevice(h->device));
  const int nb = (int)m->buffers.size();
  if (p->cms_buffer < 0 || p->cms_buffer >= nb || !m->buffers[p->cms_buffer].f32) return sb_fail(h, SB_ERR_INVALID, "bad cms buffer");
  if (p->offsets_buffer >= nb) return sb_fail(h, SB_ERR_INVALID, "bad offsets buffer");
  SbBuffer& cb = m->buffers[p->cms_buffer];
  if (cb.C > 256) return sb_fail(h, SB_ERR_UNSUPPORTED, "more than 256 confidence-map channels");
  int target = (2 * h->sm_count + m->B - 1) / m->B;
  m->g_rpc = std::max(1, (cb.H + target - 1) / target);
  m->g_chunks = (cb.H + m->g_rpc - 1) / m->g_rpc;
  if (m->gpart) cudaFree(m->gpart);
  if (m->gpoints) cudaFree(m->gpoints);
  if (m->gvals) cudaFree(m->gvals);
  if (m->crop_off_dev) cudaFree(m->crop_off_dev);
  SB_CUDA(h, cudaMalloc((void**)&m->gpart, (size_t)m->B * m->g_chunks * cb.C * 3 * 4));
  SB_CUDA(h, cudaMalloc((void**)&m->gpoints, (size_t)m->B * cb.C * 2 * 4));
  SB_CUDA(h, cudaMalloc((void**)&m->gvals, (size_t)m->B * cb.C * 4));
  SB_CUDA(h, cudaMalloc((void**)&m->crop_off_dev, (size_t)m->B * 2 * 4));
  m->gl = *p;
  m->gl_configured = true;
  return SB_OK;
}

int sb_infer_global(sb_handle_t h, int model_id, const void* images_host, int images_are_u8, int B,
                    const float* crop_offsets_host, float* out_points, float* out_vals) {
  SbModel* m = get_model(h, model_id);
  if (!m || !m->gl_configured) return sb_fail(h, SB_ERR_INVALID, "global-peak predictor not configured");
  SB_CUDA(h, cudaSetDevice(h->device));
  if (B <= 0 || B > m->B) return sb_fail(h, SB_ERR_INVALID, "bad batch");
  int rc = upload_frames(h, m, images_host, images_are_u8, B);
  if (rc) return rc;
  if (crop_offsets_host) SB_CUDA(h, cudaMemcpyAsync(m->crop_off_dev, crop_offsets_host, (size_t)B * 8, cudaMemcpyHostToDevice, h->stream));
  if ((rc = sb_run_ops(h, m, m->frames_dev, images_are_u8, B))) return rc;
  const sb_global_params& p = m->gl;
  SbBuffer& cb = m->buffers[p.cms_buffer];
  const float* off = p.offsets_buffer >= 0 ? (const float*)m->buffers[p.offsets_buffer].dev : nullptr;
  SbPeakParams pp{p.peak_threshold, p.refinement, p.integral_patch_size, (float)p.output_stride, p.input_scale};
  if ((rc = sbk_global_peaks(h, cb.dev, 0, off, B, cb.H, cb.W, cb.C, pp, crop_offsets_host ? m->crop_off_dev : nullptr,
                             m->gpart, m->g_chunks, m->g_rpc, m->gpoints, m->gvals))) return rc;
  SB_CUDA(h, cudaMemcpyAsync(out_points, m->gpoints, (size_t)B * cb.C * 8, cudaMemcpyDeviceToHost, h->stream));
  SB_CUDA(h, cudaMemcpyAsync(out_vals, m->gvals, (size_t)B * cb.C * 4, cudaMemcpyDeviceToHost, h->stream));
  SB_CUDA(h, cudaStreamSynchronize(h->stream));
  return SB_OK;
}

// ---------------------------------- centroids (top-down stage 1) ------------------------------
int sb_centroid_configure(sb_handle_t h, int model_id, const sb_centroid_params* p) {
  SbModel* m = get_model(h, model_id);
  if (!m || !p) return sb_fail(h, SB_ERR_INVALID, "bad model id / params");
  if (!m->configured) return sb_fail(h, SB_ERR_INVALID, "call sb_model_configure first");
  SB_CUDA(h, cudaSetDevice(h->device));
  const int nb = (int)m->buffers.size();
  if (p->cms_buffer < 0 || p->cms_buffer >= nb || !m->buffers[p->cms_buffer].f32) return sb_fail(h, SB_ERR_INVALID, "bad cms buffer");
  SbBuffer& cb = m->buffers[p->cms_buffer];
  sb_post_ws_free(m->ws);
  int rc = sb_post_ws_alloc(h, m->ws, m->B, cb.H, cb.W, cb.C, p->max_peaks_per_sample, 1, 1, 0);
  if (rc) return rc;
  m->ce = *p;
  m->ce_configured = true;
  return SB_OK;
}

int sb_infer_centroids(sb_handle_t h, int model_id, const void* images_host, int images_are_u8, int B,
                       float* out_centroids, float* out_vals, int32_t* out_sample_inds, int32_t* out_n,
                       int32_t* out_flags) {
  SbModel* m = get_model(h, model_id);
  if (!m || !m->ce_configured) return sb_fail(h, SB_ERR_INVALID, "centroid predictor not configured");
  SB_CUDA(h, cudaSetDevice(h->device));
  if (B <= 0 || B > m->B) return sb_fail(h, SB_ERR_INVALID, "bad batch");
  int rc = upload_frames(h, m, images_host, images_are_u8, B);
  if (rc) return rc;
  if ((rc = sb_run_ops(h, m, m->frames_dev, images_are_u8, B))) return rc;
  const sb_centroid_params& p = m->ce;
  SbBuffer& cb = m->buffers[p.cms_buffer];
  const float* off = p.offsets_buffer >= 0 ? (const float*)m->buffers[p.offsets_buffer].dev : nullptr;
  SbPeakParams pp{p.peak_threshold, p.refinement, p.integral_patch_size, (float)p.output_stride, p.input_scale};
  if ((rc = sbk_local_peaks(h, cb.dev, 0, off, B, cb.H, cb.W, cb.C, pp, m->ws))) return rc;
  std::vector<int> cnt(B), fl(B);
  SB_CUDA(h, cudaMemcpyAsync(cnt.data(), m->ws.n_peaks, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
  SB_CUDA(h, cudaMemcpyAsync(fl.data(), m->ws.flags, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
  SB_CUDA(h, cudaStreamSynchronize(h->stream));
  int total = 0;
  for (int b = 0; b < B; ++b) {
    if (cnt[b] > 0) {
      SB_CUDA(h, cudaMemcpyAsync(out_centroids + 2 * (size_t)total, m->ws.peaks + (size_t)b * m->ws.max_peaks * 2, (size_t)cnt[b] * 8, cudaMemcpyDeviceToHost, h->stream));
      SB_CUDA(h, cudaMemcpyAsync(out_vals + total, m->ws.peak_vals + (size_t)b * m->ws.max_peaks, (size_t)cnt[b] * 4, cudaMemcpyDeviceToHost, h->stream));
      for (int i = 0; i < cnt[b]; ++i) out_sample_inds[total + i] = b;
    }
    total += cnt[b];
    if (out_flags) out_flags[b] = fl[b];
  }
  SB_CUDA(h, cudaStreamSynchronize(h->stream));
  *out_n = total;
  return SB_OK;
}

}  // extern "C"
