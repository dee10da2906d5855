// CUDA-core ("direct") layer kernels: the fp32 strict-parity path, the first layers with C_in of
// 1 or 3, and every non-GEMM op (pool / upsample / add / preprocess).  Activations are NHWC with a
// channel-slice view (buffer row pitch Ctot, slice offset) so that skip connections are
// concatenated by construction (producers write straight into their slice of the concat buffer).
//
// Reference ops restated (file:line under /root/reference):
//   conv + bias + ReLU (+ BN affine after ReLU): architectures/encoder_decoder.py:117-131,
//        369-389; hourglass.py:36-45; heads.py:55-63 (1x1 linear)
//   Conv2DTranspose k3 s2 SAME:  encoder_decoder.py:304-310
//   MaxPool2D 2x2 s2 SAME:       encoder_decoder.py:109-114; unet.py:36-41; hourglass.py:94-98,133
//   UpSampling2D bilinear/nearest: encoder_decoder.py:335-339; hourglass.py:185-187
//   Add:                         hourglass.py:190
//   preprocess:                  inference.py:940-967; data/normalization.py:34-114;
//                                data/resizing.py:34-106
#pragma once
#include "sb_common.cuh"

namespace sbd {

template <typename T> __device__ __forceinline__ float ld(const T* p);
template <> __device__ __forceinline__ float ld<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld<__half>(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ void st(float* p, float v) { *p = v; }
__device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
// Split-fp16 activations (precision 2): a tensor of C logical channels occupies 3C fp16 channels [lo | hi | hi] with
// hi = fp16(v), lo = fp16(v - hi); a consumer conv whose weight rows are [Wh | Wl | Wh] (Wh = fp16(W), Wl = fp16(W - Wh))
// then accumulates lo*Wh + hi*Wl + hi*Wh = v*W up to the dropped lo*Wl term (2^-22 relative) on the fp16 tensor cores.
// The two correction planes come FIRST in channel (= K) order: the tensor core's fp32 accumulator truncates, each MMA
// step losing ~ulp(accumulator), so the 2^-11-sized terms are added while the accumulator is still small.
// `split` = C (the distance between the three planes), 0 = plain store.
__device__ __forceinline__ void st_split(float* p, float v, int) { *p = v; }
__device__ __forceinline__ void st_split(__half* p, float v, int split) {
  const __half hi = __float2half_rn(v);
  if (split) { p[0] = __float2half_rn(v - __half2float(hi)); p[split] = hi; p[2 * split] = hi; }
  else *p = hi;
}
__device__ __forceinline__ float ld_split(const __half* p, int split) { return __half2float(p[0]) + __half2float(p[split]); }

struct View {       // channel-slice view of an NHWC buffer
  void* ptr;
  int H, W, Ctot, coff;
};

constexpr int DC_TILE = 16;   // output tile 16x16 pixels, one pixel per thread
constexpr int DC_CO = 16;     // output channels per thread
constexpr int DC_CK = 8;      // input-channel chunk staged in shared memory

// Generic k x k, stride s, TF-"SAME" convolution.  weights: [k*k][Cin][Cout] fp32.
// Epilogue: + bias, ReLU (flag), * bn_scale + bn_shift (flag).
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) k_conv_direct(
    const TI* __restrict__ in, int Hin, int Win, int in_Ctot, int in_coff, int Cin,
    TO* __restrict__ out, int Hout, int Wout, int out_Ctot, int out_coff, int Cout,
    const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ bn_scale,
    const float* __restrict__ bn_shift, int k, int stride, int pad_top, int pad_left, int relu, int split = 0) {
  extern __shared__ float smem[];
  const int in_tile = (DC_TILE - 1) * stride + k;
  float* s_in = smem;                                   // [in_tile][in_tile][DC_CK]
  float* s_w = smem + in_tile * in_tile * DC_CK;        // [k*k][DC_CK][DC_CO]
  const int tiles_x = (Wout + DC_TILE - 1) / DC_TILE;
  const int tx0 = (blockIdx.x % tiles_x) * DC_TILE, ty0 = (blockIdx.x / tiles_x) * DC_TILE;
  const int co0 = blockIdx.y * DC_CO;
  const int b = blockIdx.z;
  const int lx = threadIdx.x % DC_TILE, ly = threadIdx.x / DC_TILE;
  const int ox = tx0 + lx, oy = ty0 + ly;
  const int iy0 = ty0 * stride - pad_top, ix0 = tx0 * stride - pad_left;
  const TI* in_b = in + (size_t)b * Hin * Win * in_Ctot + in_coff;
  float acc[DC_CO];
#pragma unroll
  for (int c = 0; c < DC_CO; ++c) acc[c] = 0.f;
  for (int c0 = 0; c0 < Cin; c0 += DC_CK) {
    __syncthreads();
    for (int t = threadIdx.x; t < in_tile * in_tile * DC_CK; t += 256) {
      const int c = t % DC_CK;
      const int xx = (t / DC_CK) % in_tile, yy = t / (DC_CK * in_tile);
      const int gy = iy0 + yy, gx = ix0 + xx;
      float v = 0.f;
      if (gy >= 0 && gy < Hin && gx >= 0 && gx < Win && c0 + c < Cin)
        v = ld<TI>(in_b + ((size_t)gy * Win + gx) * in_Ctot + c0 + c);
      s_in[t] = v;
    }
    for (int t = threadIdx.x; t < k * k * DC_CK * DC_CO; t += 256) {
      const int co = t % DC_CO;
      const int c = (t / DC_CO) % DC_CK;
      const int tap = t / (DC_CO * DC_CK);
      float v = 0.f;
      if (c0 + c < Cin && co0 + co < Cout) v = w[((size_t)tap * Cin + c0 + c) * Cout + co0 + co];
      s_w[t] = v;
    }
    __syncthreads();
    for (int ky = 0; ky < k; ++ky)
      for (int kx = 0; kx < k; ++kx) {
        const float* pin = s_in + ((ly * stride + ky) * in_tile + lx * stride + kx) * DC_CK;
        const float* pw = s_w + (ky * k + kx) * DC_CK * DC_CO;
#pragma unroll
        for (int c = 0; c < DC_CK; ++c) {
          const float v = pin[c];
          const float4* w4 = reinterpret_cast<const float4*>(pw + c * DC_CO);
#pragma unroll
          for (int q = 0; q < DC_CO / 4; ++q) {
            const float4 ww = w4[q];
            acc[4 * q + 0] = fmaf(v, ww.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(v, ww.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v, ww.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(v, ww.w, acc[4 * q + 3]);
          }
        }
      }
  }
  if (ox < Wout && oy < Hout) {
    TO* po = out + (((size_t)b * Hout + oy) * Wout + ox) * out_Ctot + out_coff + co0;
#pragma unroll
    for (int c = 0; c < DC_CO; ++c) {
      if (co0 + c < Cout) {
        float v = acc[c] + (bias ? bias[co0 + c] : 0.f);
        if (relu) v = fmaxf(v, 0.f);
        if (bn_scale) v = v * bn_scale[co0 + c] + bn_shift[co0 + c];
        st_split(po + c, v, split);
      }
    }
  }
}

// First layer fused with InferenceLayer.preprocess (inference.py:940-967) for the common case
// "no channel conversion, no resize": raw uint8 (or float) frame -> * (1/255) -> zero pad to the
// net size -> 3x3 SAME conv + bias + ReLU -> fp16 NHWC.  One output pixel per thread, all COUT
// channels in registers; the frame is read once from HBM (neighbour re-reads hit L1), the output
// is written once, so the kernel runs at HBM speed instead of paying a separate preprocess pass.
template <typename TI, int CIN, int COUT, int PX>
__global__ void __launch_bounds__(256) k_conv_first(const TI* __restrict__ img, int Hin, int Win, int Hnet, int Wnet,
                                                    __half* __restrict__ out, int out_Ctot, int out_coff,
                                                    const float* __restrict__ w /*[9][CIN][COUT]*/,
                                                    const float* __restrict__ bias, int relu, int in_is_u8, int split = 0) {
  // each thread: PX horizontally adjacent output pixels x COUT channels (weights read once from
  // shared memory per PX pixels; the thread's PX*COUT fp16 outputs are contiguous in NHWC)
  __shared__ __align__(16) float s_w[9 * CIN * COUT];
  __shared__ float s_b[COUT];
  for (int t = threadIdx.y * 32 + threadIdx.x; t < 9 * CIN * COUT; t += 256) s_w[t] = w[t];
  for (int t = threadIdx.y * 32 + threadIdx.x; t < COUT; t += 256) s_b[t] = bias ? bias[t] : 0.f;
  __syncthreads();
  const int ox0 = (blockIdx.x * 32 + threadIdx.x) * PX, oy = blockIdx.y * 8 + threadIdx.y, b = blockIdx.z;
  if (ox0 >= Wnet || oy >= Hnet) return;
  const TI* im = img + (size_t)b * Hin * Win * CIN;
  const float sc = in_is_u8 ? (1.0f / 255.0f) : 1.0f;
  float acc[PX][COUT];
#pragma unroll
  for (int p = 0; p < PX; ++p)
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[p][c] = s_b[c];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy + ky - 1;
    if (iy < 0 || iy >= Hin) continue;                       // SAME padding / bottom zero pad
    float in[PX + 2][CIN];
#pragma unroll
    for (int j = 0; j < PX + 2; ++j) {
      const int ix = ox0 + j - 1;
      const bool ok = ix >= 0 && ix < Win;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci)
        in[j][ci] = ok ? __fmul_rn((float)im[((size_t)iy * Win + ix) * CIN + ci], sc) : 0.f;
    }
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float4* w4 = reinterpret_cast<const float4*>(s_w + ((ky * 3 + kx) * CIN + ci) * COUT);
#pragma unroll
        for (int q = 0; q < COUT / 4; ++q) {
          const float4 ww = w4[q];
#pragma unroll
          for (int p = 0; p < PX; ++p) {
            const float v = in[p + kx][ci];
            acc[p][4 * q + 0] = fmaf(v, ww.x, acc[p][4 * q + 0]);
            acc[p][4 * q + 1] = fmaf(v, ww.y, acc[p][4 * q + 1]);
            acc[p][4 * q + 2] = fmaf(v, ww.z, acc[p][4 * q + 2]);
            acc[p][4 * q + 3] = fmaf(v, ww.w, acc[p][4 * q + 3]);
          }
        }
      }
  }
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    if (ox0 + p >= Wnet) break;
    __half* po = out + (((size_t)b * Hnet + oy) * Wnet + ox0 + p) * out_Ctot + out_coff;
    if (split) {                                             // precision 2: [lo | hi | hi] planes, COUT channels apart
#pragma unroll
      for (int q = 0; q < COUT / 8; ++q) {
        __align__(16) __half hh[8], ll[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = acc[p][8 * q + j];
          if (relu) a = fmaxf(a, 0.f);
          hh[j] = __float2half_rn(a);
          ll[j] = __float2half_rn(a - __half2float(hh[j]));
        }
        reinterpret_cast<uint4*>(po)[q] = *reinterpret_cast<uint4*>(ll);
        reinterpret_cast<uint4*>(po + COUT)[q] = *reinterpret_cast<uint4*>(hh);
        reinterpret_cast<uint4*>(po + 2 * COUT)[q] = *reinterpret_cast<uint4*>(hh);
      }
      continue;
    }
    const bool wide = (COUT % 16 == 0) && (((out_Ctot | out_coff) & 15) == 0);     // 32-byte aligned rows
#pragma unroll
    for (int q = 0; q < COUT / 8; ++q) {
      __align__(16) __half2 h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = acc[p][8 * q + 2 * j], c2 = acc[p][8 * q + 2 * j + 1];
        if (relu) { a = fmaxf(a, 0.f); c2 = fmaxf(c2, 0.f); }
        h[j] = __floats2half2_rn(a, c2);
      }
      if (wide && (q & 1) == 0 && q + 1 < COUT / 8) {
        __align__(16) __half2 h2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = acc[p][8 * (q + 1) + 2 * j], c2 = acc[p][8 * (q + 1) + 2 * j + 1];
          if (relu) { a = fmaxf(a, 0.f); c2 = fmaxf(c2, 0.f); }
          h2[j] = __floats2half2_rn(a, c2);
        }
        const uint32_t* w0 = reinterpret_cast<const uint32_t*>(h);
        const uint32_t* w1 = reinterpret_cast<const uint32_t*>(h2);
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(po + 8 * q), "r"(w0[0]), "r"(w0[1]),
                     "r"(w0[2]), "r"(w0[3]), "r"(w1[0]), "r"(w1[1]), "r"(w1[2]), "r"(w1[3]) : "memory");
      } else if (!(wide && (q & 1) == 1)) {
        reinterpret_cast<uint4*>(po)[q] = *reinterpret_cast<uint4*>(h);
      }
    }
  }
}

// Conv2DTranspose(k=3, strides=2, padding="same"): out (2H, 2W).  Per axis:
//   out[2i] = in[i]*W[0] + in[i-1]*W[2];  out[2i+1] = in[i]*W[1].   weights [9][Cin][Cout].
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) k_tconv_direct(
    const TI* __restrict__ in, int Hin, int Win, int in_Ctot, int in_coff, int Cin,
    TO* __restrict__ out, int out_Ctot, int out_coff, int Cout, const float* __restrict__ w,
    const float* __restrict__ bias, int relu, int split = 0) {
  const int Hout = 2 * Hin, Wout = 2 * Win;
  const int b = blockIdx.z;
  const int co0 = blockIdx.y * DC_CO;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  if (pix >= Hout * Wout) return;
  const int oy = pix / Wout, ox = pix - oy * Wout;
  const TI* in_b = in + (size_t)b * Hin * Win * in_Ctot + in_coff;
  float acc[DC_CO];
#pragma unroll
  for (int c = 0; c < DC_CO; ++c) acc[c] = 0.f;
  int kys[2], iys[2], nky = 0, kxs[2], ixs[2], nkx = 0;
  if (oy & 1) { kys[0] = 1; iys[0] = oy >> 1; nky = 1; }
  else { kys[0] = 0; iys[0] = oy >> 1; nky = 1; if ((oy >> 1) - 1 >= 0) { kys[1] = 2; iys[1] = (oy >> 1) - 1; nky = 2; } }
  if (ox & 1) { kxs[0] = 1; ixs[0] = ox >> 1; nkx = 1; }
  else { kxs[0] = 0; ixs[0] = ox >> 1; nkx = 1; if ((ox >> 1) - 1 >= 0) { kxs[1] = 2; ixs[1] = (ox >> 1) - 1; nkx = 2; } }
  for (int a = 0; a < nky; ++a)
    for (int bb = 0; bb < nkx; ++bb) {
      const TI* pin = in_b + ((size_t)iys[a] * Win + ixs[bb]) * in_Ctot;
      const float* pw = w + (size_t)(kys[a] * 3 + kxs[bb]) * Cin * Cout + co0;
      for (int c = 0; c < Cin; ++c) {
        const float v = ld<TI>(pin + c);
#pragma unroll
        for (int q = 0; q < DC_CO; ++q)
          if (co0 + q < Cout) acc[q] = fmaf(v, pw[(size_t)c * Cout + q], acc[q]);
      }
    }
  TO* po = out + (((size_t)b * Hout + oy) * Wout + ox) * out_Ctot + out_coff + co0;
#pragma unroll
  for (int c = 0; c < DC_CO; ++c)
    if (co0 + c < Cout) {
      float v = acc[c] + (bias ? bias[co0 + c] : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      st_split(po + c, v, split);
    }
}

template <typename T>
__global__ void k_maxpool2(const T* __restrict__ in, int Hin, int Win, int in_Ctot, int in_coff, int C,
                           T* __restrict__ out, int Hout, int Wout, int out_Ctot, int out_coff, size_t total) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const int ox = (int)((t / C) % Wout);
    const int oy = (int)((t / ((size_t)C * Wout)) % Hout);
    const int b = (int)(t / ((size_t)C * Wout * Hout));
    const T* pin = in + (size_t)b * Hin * Win * in_Ctot + in_coff + c;
    float m = -INFINITY;
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) {
        const int iy = 2 * oy + dy, ix = 2 * ox + dx;
        if (iy < Hin && ix < Win) m = fmaxf(m, ld<T>(pin + ((size_t)iy * Win + ix) * in_Ctot));
      }
    st(out + (((size_t)b * Hout + oy) * Wout + ox) * out_Ctot + out_coff + c, m);
  }
}

// x2 upsampling: bilinear with half-pixel centres (weights 1/4, 3/4, edge clamp) or nearest.
template <typename T>
__global__ void k_upsample2(const T* __restrict__ in, int Hin, int Win, int in_Ctot, int in_coff, int C,
                            T* __restrict__ out, int out_Ctot, int out_coff, int bilinear, size_t total) {
  const int Hout = 2 * Hin, Wout = 2 * Win;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const int ox = (int)((t / C) % Wout);
    const int oy = (int)((t / ((size_t)C * Wout)) % Hout);
    const int b = (int)(t / ((size_t)C * Wout * Hout));
    const T* pin = in + (size_t)b * Hin * Win * in_Ctot + in_coff + c;
    float v;
    if (!bilinear) {
      v = ld<T>(pin + ((size_t)(oy >> 1) * Win + (ox >> 1)) * in_Ctot);
    } else {
      const float sy = ((float)oy + 0.5f) * 0.5f - 0.5f, sx = ((float)ox + 0.5f) * 0.5f - 0.5f;
      const float fy = floorf(sy), fx = floorf(sx);
      const int y0 = max((int)fy, 0), y1 = min((int)ceilf(sy), Hin - 1);
      const int x0 = max((int)fx, 0), x1 = min((int)ceilf(sx), Win - 1);
      const float ly = sy - fy, lx = sx - fx;
      const float tl = ld<T>(pin + ((size_t)y0 * Win + x0) * in_Ctot), tr = ld<T>(pin + ((size_t)y0 * Win + x1) * in_Ctot);
      const float bl = ld<T>(pin + ((size_t)y1 * Win + x0) * in_Ctot), br = ld<T>(pin + ((size_t)y1 * Win + x1) * in_Ctot);
      const float tp = tl + (tr - tl) * lx, bt = bl + (br - bl) * lx;
      v = tp + (bt - tp) * ly;
    }
    st(out + (((size_t)b * Hout + oy) * Wout + ox) * out_Ctot + out_coff + c, v);
  }
}

template <typename T>
__global__ void k_add(const T* __restrict__ a, int a_Ctot, int a_coff, const T* __restrict__ bsrc, int b_Ctot,
                      int b_coff, T* __restrict__ out, int out_Ctot, int out_coff, int C, size_t npix) {
  const size_t total = npix * C;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const size_t p = t / C;
    st(out + p * out_Ctot + out_coff + c, ld<T>(a + p * a_Ctot + a_coff + c) + ld<T>(bsrc + p * b_Ctot + b_coff + c));
  }
}

template <typename T>
__global__ void k_copy(const T* __restrict__ a, int a_Ctot, int a_coff, T* __restrict__ out, int out_Ctot,
                       int out_coff, int C, size_t npix) {
  const size_t total = npix * C;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const size_t p = t / C;
    out[p * out_Ctot + out_coff + c] = a[p * a_Ctot + a_coff + c];
  }
}

// Preprocess (InferenceLayer.preprocess): gray<->rgb, u8 -> float * (1/255), bilinear resize by
// input_scale (half-pixel centres, no antialias), zero pad bottom/right to the net input size.
//   mode_ch: 0 keep, 1 rgb->gray (u8: truncating round trip like tf.image.rgb_to_grayscale), 2 gray->rgb
template <typename TI, typename TO>
__global__ void k_preprocess(const TI* __restrict__ in, int Hin, int Win, int Cin, TO* __restrict__ out,
                             int Hnet, int Wnet, int Cnet, int Hres, int Wres, int resize, int mode_ch,
                             int in_is_u8, size_t total) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % Cnet);
    const int ox = (int)((t / Cnet) % Wnet);
    const int oy = (int)((t / ((size_t)Cnet * Wnet)) % Hnet);
    const int b = (int)(t / ((size_t)Cnet * Wnet * Hnet));
    float v = 0.f;
    if (oy < Hres && ox < Wres) {
      const TI* img = in + (size_t)b * Hin * Win * Cin;
      auto fetch = [&](int y, int x) -> float {
        const TI* p = img + ((size_t)y * Win + x) * Cin;
        float f;
        if (mode_ch == 1) {
          const float sc = in_is_u8 ? (1.0f / 255.0f) : 1.0f;
          const float g = __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn((float)p[0], sc), 0.2989f),
                                              __fmul_rn(__fmul_rn((float)p[1], sc), 0.5870f)),
                                    __fmul_rn(__fmul_rn((float)p[2], sc), 0.1140f));
          if (in_is_u8) f = __fmul_rn(truncf(fminf(fmaxf(__fmul_rn(g, 255.5f), 0.f), 255.f)), 1.0f / 255.0f);
          else f = g;
        } else {
          const int cc = (mode_ch == 2) ? 0 : c;
          f = (float)p[cc];
          if (in_is_u8) f = __fmul_rn(f, 1.0f / 255.0f);
        }
        return f;
      };
      if (!resize) {
        v = fetch(oy, ox);
      } else {
        const float scy = (float)Hin / (float)Hres, scx = (float)Win / (float)Wres;
        const float sy = __fadd_rn(__fmul_rn(__fadd_rn((float)oy, 0.5f), scy), -0.5f);
        const float sx = __fadd_rn(__fmul_rn(__fadd_rn((float)ox, 0.5f), scx), -0.5f);
        const float fy = floorf(sy), fx = floorf(sx);
        const int y0 = max((int)fy, 0), y1 = min((int)ceilf(sy), Hin - 1);
        const int x0 = max((int)fx, 0), x1 = min((int)ceilf(sx), Win - 1);
        const float ly = sy - fy, lx = sx - fx;
        const float tl = fetch(y0, x0), tr = fetch(y0, x1), bl = fetch(y1, x0), br = fetch(y1, x1);
        const float tp = __fadd_rn(tl, __fmul_rn(__fadd_rn(tr, -tl), lx));
        const float bt = __fadd_rn(bl, __fmul_rn(__fadd_rn(br, -bl), lx));
        v = __fadd_rn(tp, __fmul_rn(__fadd_rn(bt, -tp), ly));
      }
    }
    st(out + t, v);
  }
}

// ---- precision 2 (split-fp16 activations, see st_split): the elementwise ops act on v = hi + lo ----
// C is the LOGICAL channel count; the tensors hold 3C channels [lo | hi | hi] from their channel offset.
__global__ void k_maxpool2_split(const __half* __restrict__ in, int Hin, int Win, int in_Ctot, int in_coff, int C,
                                 __half* __restrict__ out, int Hout, int Wout, int out_Ctot, int out_coff, size_t total) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const int ox = (int)((t / C) % Wout);
    const int oy = (int)((t / ((size_t)C * Wout)) % Hout);
    const int b = (int)(t / ((size_t)C * Wout * Hout));
    const __half* pin = in + (size_t)b * Hin * Win * in_Ctot + in_coff + c;
    float m = -INFINITY;
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) {
        const int iy = 2 * oy + dy, ix = 2 * ox + dx;
        if (iy < Hin && ix < Win) m = fmaxf(m, ld_split(pin + ((size_t)iy * Win + ix) * in_Ctot, C));
      }
    st_split(out + (((size_t)b * Hout + oy) * Wout + ox) * out_Ctot + out_coff + c, m, C);
  }
}

__global__ void k_upsample2_split(const __half* __restrict__ in, int Hin, int Win, int in_Ctot, int in_coff, int C,
                                  __half* __restrict__ out, int out_Ctot, int out_coff, int bilinear, size_t total) {
  const int Hout = 2 * Hin, Wout = 2 * Win;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const int ox = (int)((t / C) % Wout);
    const int oy = (int)((t / ((size_t)C * Wout)) % Hout);
    const int b = (int)(t / ((size_t)C * Wout * Hout));
    const __half* pin = in + (size_t)b * Hin * Win * in_Ctot + in_coff + c;
    float v;
    if (!bilinear) {
      v = ld_split(pin + ((size_t)(oy >> 1) * Win + (ox >> 1)) * in_Ctot, C);
    } else {
      const float sy = ((float)oy + 0.5f) * 0.5f - 0.5f, sx = ((float)ox + 0.5f) * 0.5f - 0.5f;
      const float fy = floorf(sy), fx = floorf(sx);
      const int y0 = max((int)fy, 0), y1 = min((int)ceilf(sy), Hin - 1);
      const int x0 = max((int)fx, 0), x1 = min((int)ceilf(sx), Win - 1);
      const float ly = sy - fy, lx = sx - fx;
      const float tl = ld_split(pin + ((size_t)y0 * Win + x0) * in_Ctot, C), tr = ld_split(pin + ((size_t)y0 * Win + x1) * in_Ctot, C);
      const float bl = ld_split(pin + ((size_t)y1 * Win + x0) * in_Ctot, C), br = ld_split(pin + ((size_t)y1 * Win + x1) * in_Ctot, C);
      const float tp = tl + (tr - tl) * lx, bt = bl + (br - bl) * lx;
      v = tp + (bt - tp) * ly;
    }
    st_split(out + (((size_t)b * Hout + oy) * Wout + ox) * out_Ctot + out_coff + c, v, C);
  }
}

__global__ void k_add_split(const __half* __restrict__ a, int a_Ctot, int a_coff, const __half* __restrict__ bsrc, int b_Ctot,
                            int b_coff, __half* __restrict__ out, int out_Ctot, int out_coff, int C, size_t npix) {
  const size_t total = npix * C;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const size_t p = t / C;
    st_split(out + p * out_Ctot + out_coff + c, ld_split(a + p * a_Ctot + a_coff + c, C) + ld_split(bsrc + p * b_Ctot + b_coff + c, C), C);
  }
}

}  // namespace sbd
