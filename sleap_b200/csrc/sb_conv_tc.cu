// tcgen05 implicit-GEMM convolution for sm_100a: NHWC fp16 activations, fp32 accumulation in TMEM.
//
// Replaces cuDNN's Conv2D / Conv2DTranspose as dispatched by the reference's Keras graph
// (sleap/nn/architectures/encoder_decoder.py:117-131, 304-310, 369-389; hourglass.py:36-45;
// heads.py:55-63).
//
// Mapping (one CTA = one 8x16-pixel output tile x one N tile of output channels):
//   GEMM M = 128 output pixels, N = C_out tile (16..256), K = taps x C_in.
//   A operand: TMA (cp.async.bulk.tensor.4d) loads an NHWC box [KC ch, 16 px, 8+halo rows] with
//     hardware zero fill outside the image (= TF "SAME" padding) into 128B/64B/32B-swizzled
//     shared memory, one pixel per row, i.e. exactly the canonical K-major UMMA layout.  The
//     x shift of a filter tap is baked into the TMA coordinate (one load per distinct dx); the
//     y shift is a whole number of 16-pixel rows = a swizzle-atom-aligned start-address offset,
//     so the three ky taps of one dx share one staged tile.
//   B operand: weights pre-arranged [tap][C_out][C_in] fp16 (K-major), TMA box [KC, N, 1].
//   D: TMEM accumulator (128 lanes x N fp32 columns), tcgen05.mma.cta_group::1.kind::f16 issued
//     by one thread; epilogue tcgen05.ld 32x32b -> bias / ReLU / BN affine -> fp16 (or fp32 for
//     head outputs) NHWC stores into the consumer's channel slice.
//   Conv2DTranspose(k3,s2) = four sub-pixel phase GEMMs over the input grid with strided stores.
// Warp roles: warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer, all 4 warps = epilogue.
#include <cuda.h>

#include <algorithm>

#include "sb_model.h"

namespace {

constexpr int TW = 16, TH = 8;          // output tile (pixels); M = 128
constexpr int MAX_GROUPS = 3, MAX_TAPS = 3;

struct TcTap { int row_off, w_tap; };
struct TcGroup { int dx, n_taps; TcTap taps[MAX_TAPS]; };

struct TcParams {
  int H, W;                    // iteration grid (input grid for tconv phases, output grid for convs)
  int tiles_x;
  int n_chunks, KC;
  int n_groups;
  TcGroup groups[MAX_GROUPS];
  int dy0, box_rows;
  int N, Cout;                 // UMMA N of this launch, valid output channels
  uint32_t idesc;
  int tmem_cols;
  void* out;
  int out_f32, out_H, out_W, out_Ctot, out_coff;
  int oy_mul, oy_add, ox_mul, ox_add;
  const float* bias;
  const float* bn_scale;
  const float* bn_shift;
  int relu;
  void* pool_out;               // optional fused MaxPool2D(2,2) output (same dtype as out)
  int pool_H, pool_W, pool_Ctot, pool_coff;
  int a_slot_bytes, b_slot_bytes, n_a_slots, n_b_slots;
  int a_tx_bytes, b_tx_bytes;
  int layout_type;             // UMMA LayoutType: 2 = SW128, 4 = SW64, 6 = SW32
  int row_bytes;               // KC * 2
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it) {
    if (it > (1u << 24)) {   // never hang the GPU: a lost arrival becomes a launch error
      printf("[sb_conv_tc] mbarrier timeout tag=%d block=(%d,%d,%d) thread=%d\n", tag, blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// K-major swizzled UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4,
//   [46,48) version = 1 (Blackwell), [61,64) layout type.  SBO = 8 rows x row_bytes.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int row_bytes, int layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(((8 * row_bytes) >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

__global__ void __launch_bounds__(128) k_conv_tc(const __grid_constant__ CUtensorMap mapA,
                                                 const __grid_constant__ CUtensorMap mapB,
                                                 const __grid_constant__ TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve shared memory (ring slots are 1024-aligned: required by the 128B swizzle atoms)
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = base;
  uint8_t* b_ring = a_ring + (size_t)P.n_a_slots * P.a_slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)P.n_b_slots * P.b_slot_bytes);
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + P.n_a_slots;
  uint64_t* fullB = emptyA + P.n_a_slots;
  uint64_t* emptyB = fullB + P.n_b_slots;
  uint64_t* accum = emptyB + P.n_b_slots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int x0 = (tile % P.tiles_x) * TW, y0 = (tile / P.tiles_x) * TH;
  const int n0 = blockIdx.y * P.N;
  const int b = blockIdx.z;

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.n_a_slots; ++i) { mbar_init(smem_u32(fullA + i), 1); mbar_init(smem_u32(emptyA + i), 1); }
    for (int i = 0; i < P.n_b_slots; ++i) { mbar_init(smem_u32(fullB + i), 1); mbar_init(smem_u32(emptyB + i), 1); }
    mbar_init(smem_u32(accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(P.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------ TMA producer ------------------------------
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    for (int ch = 0; ch < P.n_chunks; ++ch) {
      for (int g = 0; g < P.n_groups; ++g) {
        mbar_wait(smem_u32(emptyA + sa), pha ^ 1, 1);
        mbar_expect_tx(smem_u32(fullA + sa), (uint32_t)P.a_tx_bytes);
        tma_load_4d(smem_u32(a_ring + (size_t)sa * P.a_slot_bytes), &mapA, smem_u32(fullA + sa), ch * P.KC,
                    x0 + P.groups[g].dx, y0 + P.dy0, b);
        for (int t = 0; t < P.groups[g].n_taps; ++t) {
          mbar_wait(smem_u32(emptyB + sb), phb ^ 1, 2);
          mbar_expect_tx(smem_u32(fullB + sb), (uint32_t)P.b_tx_bytes);
          tma_load_3d(smem_u32(b_ring + (size_t)sb * P.b_slot_bytes), &mapB, smem_u32(fullB + sb), ch * P.KC, n0,
                      P.groups[g].taps[t].w_tap);
          if (++sb == P.n_b_slots) { sb = 0; phb ^= 1; }
        }
        if (++sa == P.n_a_slots) { sa = 0; pha ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------ MMA issuer --------------------------------
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    uint32_t first = 1;
    const int ksteps = P.KC / 16;
    for (int ch = 0; ch < P.n_chunks; ++ch) {
      for (int g = 0; g < P.n_groups; ++g) {
        mbar_wait(smem_u32(fullA + sa), pha, 3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_base = smem_u32(a_ring + (size_t)sa * P.a_slot_bytes);
        for (int t = 0; t < P.groups[g].n_taps; ++t) {
          mbar_wait(smem_u32(fullB + sb), phb, 4);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t b_base = smem_u32(b_ring + (size_t)sb * P.b_slot_bytes);
          const uint32_t a_tap = a_base + (uint32_t)(P.groups[g].taps[t].row_off * TW * P.row_bytes);
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t da = make_desc(a_tap + k * 32, P.row_bytes, P.layout_type);
            const uint64_t db = make_desc(b_base + k * 32, P.row_bytes, P.layout_type);
            tc_mma_f16(tmem_base, da, db, P.idesc, first ? 0u : 1u);
            first = 0;
          }
          tc_commit(smem_u32(emptyB + sb));
          if (++sb == P.n_b_slots) { sb = 0; phb ^= 1; }
        }
        tc_commit(smem_u32(emptyA + sa));
        if (++sa == P.n_a_slots) { sa = 0; pha ^= 1; }
      }
    }
    tc_commit(smem_u32(accum));
  }
  __syncwarp();

  // ------------------------------ epilogue (all 4 warps) ----------------------
  mbar_wait(smem_u32(accum), 0, 5);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int m = warp * 32 + lane;                     // accumulator row = TMEM lane
  const int iy = y0 + m / TW, ix = x0 + m % TW;
  const bool valid = (iy < P.H) && (ix < P.W);
  const int oy = iy * P.oy_mul + P.oy_add, ox = ix * P.ox_mul + P.ox_add;
  const size_t pix = ((size_t)b * P.out_H + oy) * P.out_W + ox;
  for (int c0 = 0; c0 < P.N; c0 += 16) {
    uint32_t r[16];
    tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int co = n0 + c0 + j;
      float x = __uint_as_float(r[j]);
      if (co < P.Cout) {
        if (P.bias) x += P.bias[co];
        if (P.relu) x = fmaxf(x, 0.f);
        if (P.bn_scale) x = x * P.bn_scale[co] + P.bn_shift[co];
      }
      v[j] = x;
    }
    if (P.pool_out != nullptr) {
      // fused MaxPool2D(2, strides=2): lanes hold pixels (ty = 2*warp + lane/16, tx = lane%16) of the
      // tile, so the 2x2 partners are lane^1 (x) and lane^16 (y); even-x lanes of the upper row store.
      float pv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float a = v[j];
        a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, 1));
        a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, 16));
        pv[j] = a;
      }
      if (valid && lane < 16 && (lane & 1) == 0 && n0 + c0 + 16 <= P.Cout) {
        const int py = (y0 >> 1) + warp, px = (x0 >> 1) + (lane >> 1);
        __half* pp = reinterpret_cast<__half*>(P.pool_out) + (((size_t)b * P.pool_H + py) * P.pool_W + px) * P.pool_Ctot +
                     P.pool_coff + n0 + c0;
        __half2 h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = __floats2half2_rn(pv[2 * j], pv[2 * j + 1]);
        reinterpret_cast<uint4*>(pp)[0] = *reinterpret_cast<uint4*>(&h[0]);
        reinterpret_cast<uint4*>(pp)[1] = *reinterpret_cast<uint4*>(&h[4]);
      }
    }
    if (valid) {
      if (P.out_f32) {
        float* po = reinterpret_cast<float*>(P.out) + pix * P.out_Ctot + P.out_coff + n0 + c0;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n0 + c0 + j < P.Cout) po[j] = v[j];
      } else {
        __half* po = reinterpret_cast<__half*>(P.out) + pix * P.out_Ctot + P.out_coff + n0 + c0;
        if (n0 + c0 + 16 <= P.Cout) {
          uint4 q0, q1;
          __half2 h[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
          q0 = *reinterpret_cast<uint4*>(&h[0]);
          q1 = *reinterpret_cast<uint4*>(&h[4]);
          reinterpret_cast<uint4*>(po)[0] = q0;
          reinterpret_cast<uint4*>(po)[1] = q1;
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c0 + j < P.Cout) po[j] = __float2half_rn(v[j]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols) : "memory");
  }
}

// ------------------------------- host side ---------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

struct TcLaunch {
  CUtensorMap mapA, mapB;
  TcParams P;
  dim3 grid;
  size_t smem;
};

}  // namespace

struct SbConvTcPlan {
  std::vector<TcLaunch> launches;   // 1 for conv, 4 phases for tconv
  __half* w16 = nullptr;            // [taps][Cout_pad][Cin]
  int Cout_pad = 0;
};

static CUtensorMapSwizzle swz_for(int KC) {
  return KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

static bool tc_eligible(const SbModel* m, const SbOp& op) {
  if (getenv("SB_DISABLE_TC")) return false;
  if (op.kind() != SB_OPK_CONV && op.kind() != SB_OPK_TCONV) return false;
  const int Cin = op.in_C();
  if (!(Cin == 16 || Cin == 32 || Cin % 64 == 0)) return false;
  if (op.kind() == SB_OPK_CONV && !((op.k() == 1 || op.k() == 3) && op.stride() == 1)) return false;
  const SbBuffer& ib = m->buffers[op.in_buf()];
  const SbBuffer& ob = m->buffers[op.out_buf()];
  if (ib.f32) return false;
  if (ib.W < TW || ib.H < TH + 2) return false;   // TMA box must fit inside the tensor
  if (ib.C % 8 || op.in_coff() % 8) return false;
  if (!ob.f32 && (ob.C % 8 || op.out_coff() % 8)) return false;
  return true;
}

void sb_conv_tc_release(SbModel* m) {
  for (SbConvTcPlan* p : m->tc_plans)
    if (p) { if (p->w16) cudaFree(p->w16); delete p; }
  m->tc_plans.clear();
}

bool sb_conv_tc_can(const SbModel* m, int op_index) {
  return op_index < (int)m->tc_plans.size() && m->tc_plans[op_index] != nullptr;
}

static int make_launch(sb_handle_s* h, SbModel* m, const SbOp& op, SbConvTcPlan* plan, int n_groups,
                       const TcGroup* groups, int dy0, int extra_rows, int n_wtaps, int oy_mul, int oy_add,
                       int ox_mul, int ox_add) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return sb_fail(h, SB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const SbBuffer& ib = m->buffers[op.in_buf()];
  const SbBuffer& ob = m->buffers[op.out_buf()];
  const int Cin = op.in_C(), Cout = op.out_C();
  const int KC = Cin >= 64 ? 64 : Cin;
  TcLaunch L;
  memset(&L, 0, sizeof(L));
  TcParams& P = L.P;
  P.H = ib.H; P.W = ib.W;
  P.tiles_x = (ib.W + TW - 1) / TW;
  const int tiles_y = (ib.H + TH - 1) / TH;
  P.n_chunks = Cin / KC; P.KC = KC;
  P.n_groups = n_groups;
  for (int g = 0; g < n_groups; ++g) P.groups[g] = groups[g];
  P.dy0 = dy0; P.box_rows = TH + extra_rows;
  const int N = std::min(plan->Cout_pad, 256);
  P.N = N; P.Cout = Cout;
  // cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format F16 (0) @7/@10, K-major (0) @15/@16,
  // n_dim = N>>3 @17, m_dim = M>>4 @24.
  P.idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  int cols = 32;
  while (cols < N) cols <<= 1;
  P.tmem_cols = cols;
  P.out = ob.dev; P.out_f32 = ob.f32; P.out_H = ob.H; P.out_W = ob.W; P.out_Ctot = ob.C; P.out_coff = op.out_coff();
  P.oy_mul = oy_mul; P.oy_add = oy_add; P.ox_mul = ox_mul; P.ox_add = ox_add;
  P.bias = op.b_off() >= 0 ? m->weights_dev + op.b_off() : nullptr;
  P.bn_scale = (op.flags() & SB_OPF_BN) ? m->weights_dev + op.bn_scale_off() : nullptr;
  P.bn_shift = (op.flags() & SB_OPF_BN) ? m->weights_dev + op.bn_shift_off() : nullptr;
  P.relu = (op.flags() & SB_OPF_RELU) ? 1 : 0;
  P.pool_out = nullptr;
  if (op.kind() == SB_OPK_CONV && op.pool_buf() >= 0 && !ob.f32 && Cout % 16 == 0 &&
      m->buffers[op.pool_buf()].C % 8 == 0 && op.pool_coff() % 8 == 0 && ib.H % 2 == 0 && ib.W % 2 == 0) {
    const SbBuffer& pb = m->buffers[op.pool_buf()];
    P.pool_out = pb.dev; P.pool_H = pb.H; P.pool_W = pb.W; P.pool_Ctot = pb.C; P.pool_coff = op.pool_coff();
  }
  P.row_bytes = KC * 2;
  P.layout_type = KC == 64 ? 2 : (KC == 32 ? 4 : 6);
  P.a_tx_bytes = P.box_rows * TW * KC * 2;
  P.b_tx_bytes = N * KC * 2;
  P.a_slot_bytes = (P.a_tx_bytes + 1023) / 1024 * 1024;
  P.b_slot_bytes = (P.b_tx_bytes + 1023) / 1024 * 1024;
  int total_steps = 0;
  for (int g = 0; g < n_groups; ++g) total_steps += groups[g].n_taps;
  P.n_a_slots = std::min(3, P.n_chunks * n_groups);
  P.n_b_slots = std::min(4, P.n_chunks * total_steps);
  while ((size_t)P.n_a_slots * P.a_slot_bytes + (size_t)P.n_b_slots * P.b_slot_bytes > 200 * 1024 && P.n_b_slots > 2) P.n_b_slots--;
  while ((size_t)P.n_a_slots * P.a_slot_bytes + (size_t)P.n_b_slots * P.b_slot_bytes > 200 * 1024 && P.n_a_slots > 2) P.n_a_slots--;
  L.smem = (size_t)P.n_a_slots * P.a_slot_bytes + (size_t)P.n_b_slots * P.b_slot_bytes + 1024 /*align slack*/ +
           (size_t)(2 * P.n_a_slots + 2 * P.n_b_slots + 1) * 8 + 16;
  L.grid = dim3(P.tiles_x * tiles_y, plan->Cout_pad / N, 1 /* z = batch, set at launch */);

  // A: NHWC view (slice channels, W, H, batch)
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)ib.W, (cuuint64_t)ib.H, (cuuint64_t)m->B};
    cuuint64_t strides[3] = {(cuuint64_t)ib.C * 2, (cuuint64_t)ib.W * ib.C * 2, (cuuint64_t)ib.H * ib.W * ib.C * 2};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)TW, (cuuint32_t)P.box_rows, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    void* gptr = (void*)((__half*)ib.dev + op.in_coff());
    CUresult r = enc(&L.mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, gptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz_for(KC), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return sb_fail(h, SB_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: %d", (int)r);
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)plan->Cout_pad, (cuuint64_t)n_wtaps};
    cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)plan->Cout_pad * Cin * 2};
    cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)N, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&L.mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)plan->w16, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz_for(KC), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return sb_fail(h, SB_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
  }
  plan->launches.push_back(L);
  return 0;
}

int sb_conv_tc_prepare(sb_handle_s* h, SbModel* m) {
  m->tc_plans.assign(m->ops.size(), nullptr);
  m->skip_op.assign(m->ops.size(), 0);
  if (m->precision != 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    SB_CUDA(h, cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    attr_set = true;
  }
  for (size_t oi = 0; oi < m->ops.size(); ++oi) {
    const SbOp& op = m->ops[oi];
    if (!tc_eligible(m, op)) continue;
    const int Cin = op.in_C(), Cout = op.out_C(), k = op.k(), taps = k * k;
    SbConvTcPlan* plan = new SbConvTcPlan();
    int cp = (Cout + 15) / 16 * 16;
    if (cp > 256) cp = (cp + 255) / 256 * 256;
    plan->Cout_pad = cp;
    // weights: fp32 blob [tap][Cin][Cout] -> fp16 [tap][Cout_pad][Cin] (K-major B operand)
    std::vector<__half> w16((size_t)taps * cp * Cin, __float2half(0.f));
    const float* w = m->weights_host.data() + op.w_off();
    for (int t = 0; t < taps; ++t)
      for (int ci = 0; ci < Cin; ++ci)
        for (int co = 0; co < Cout; ++co)
          w16[((size_t)t * cp + co) * Cin + ci] = __float2half_rn(w[((size_t)t * Cin + ci) * Cout + co]);
    cudaError_t e = cudaMalloc((void**)&plan->w16, w16.size() * sizeof(__half));
    if (e != cudaSuccess) { delete plan; return sb_fail(h, SB_ERR_CUDA, "cudaMalloc w16: %s", cudaGetErrorString(e)); }
    e = cudaMemcpy(plan->w16, w16.data(), w16.size() * sizeof(__half), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(plan->w16); delete plan; return sb_fail(h, SB_ERR_CUDA, "copy w16: %s", cudaGetErrorString(e)); }
    int rc = 0;
    if (op.kind() == SB_OPK_CONV && k == 3) {
      TcGroup g[3];
      for (int kx = 0; kx < 3; ++kx) {
        g[kx].dx = kx - 1; g[kx].n_taps = 3;
        for (int ky = 0; ky < 3; ++ky) g[kx].taps[ky] = TcTap{ky, ky * 3 + kx};
      }
      rc = make_launch(h, m, op, plan, 3, g, -1, 2, 9, 1, 0, 1, 0);
    } else if (op.kind() == SB_OPK_CONV) {   // 1x1
      TcGroup g[1];
      g[0].dx = 0; g[0].n_taps = 1; g[0].taps[0] = TcTap{0, 0};
      rc = make_launch(h, m, op, plan, 1, g, 0, 0, 1, 1, 0, 1, 0);
    } else {
      // Conv2DTranspose k3 s2: out[2i+a] gets (ky, iy) = a==0 ? {(0,i),(2,i-1)} : {(1,i)}; same along x.
      for (int a = 0; a < 2 && !rc; ++a)
        for (int bx = 0; bx < 2 && !rc; ++bx) {
          TcGroup g[2];
          int ng = 0;
          const int kxs[2] = {bx == 0 ? 0 : 1, 2}, dxs[2] = {0, -1};
          const int nkx = bx == 0 ? 2 : 1;
          const int extra = a == 0 ? 1 : 0;               // rows y0-1 .. y0+TH-1 when a == 0
          for (int q = 0; q < nkx; ++q) {
            g[ng].dx = dxs[q];
            if (a == 0) {
              g[ng].n_taps = 2;
              g[ng].taps[0] = TcTap{1, 0 * 3 + kxs[q]};     // ky=0 reads row i   (box row 1)
              g[ng].taps[1] = TcTap{0, 2 * 3 + kxs[q]};     // ky=2 reads row i-1 (box row 0)
            } else {
              g[ng].n_taps = 1;
              g[ng].taps[0] = TcTap{0, 1 * 3 + kxs[q]};
            }
            ++ng;
          }
          rc = make_launch(h, m, op, plan, ng, g, a == 0 ? -1 : 0, extra, 9, 2, a, 2, bx);
        }
    }
    if (rc) { cudaFree(plan->w16); delete plan; return rc; }
    m->tc_plans[oi] = plan;
    if (op.kind() == SB_OPK_CONV && op.pool_buf() >= 0 && oi + 1 < m->ops.size() &&
        m->ops[oi + 1].kind() == SB_OPK_POOL && (m->ops[oi + 1].flags() & SB_OPF_FUSED_POOL))
      if (plan->launches[0].P.pool_out != nullptr) m->skip_op[oi + 1] = 1;
  }
  return 0;
}

int sb_conv_tc_launch(sb_handle_s* h, SbModel* m, int op_index, int B) {
  SbConvTcPlan* plan = m->tc_plans[op_index];
  for (TcLaunch& L : plan->launches) {
    dim3 g = L.grid;
    g.z = B;
    k_conv_tc<<<g, 128, L.smem, h->stream>>>(L.mapA, L.mapB, L.P);
    SB_CHECK_LAUNCH(h);
  }
  return 0;
}
