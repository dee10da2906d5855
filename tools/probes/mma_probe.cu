// Issue-rate probe for SS-mode tcgen05.mma (kind::f16, K = 16 per instruction) on sm_100a:
// how many clocks one MMA of shape M x N takes when both operands sit in shared memory and
// nothing else runs, for cta_group::1 (M = 128) and cta_group::2 (M = 256 over a CTA pair), by N
// and by swizzle mode.  The convolution kernels' ">= 128-channel layers are bound by the operand
// feed, not by L2" diagnosis rests on these numbers.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && timeout 120 ./mma_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int row_bytes, int layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(((8 * row_bytes) >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
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

struct Args {
  int n;          // MMA N
  int nmma;       // MMAs per measurement
  int row_bytes;  // 128 / 64 / 32: swizzle span = bytes of K per staged row
  int layout;     // UMMA layout type: 2 = SW128, 4 = SW64, 6 = SW32
  int stages;     // distinct operand tiles cycled through
  int a_off_rows; // start address of the A tile shifted by this many staged rows (the halo kernels' tap offsets)
  int random;     // 1: pseudo-random fp16 operands in [-2, 2) (switching activity of real data) instead of near-constant ones
  int a_pitch;    // rows between consecutive 8-row groups of A (8 = dense tile; 10 = the 8+2-pixel halo box of the conv kernels)
  unsigned long long* cycles;  // per CTA
};

template <int CG, int KS>
__global__ void __launch_bounds__(128) k_probe(Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t cta_rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));

  const int b_rows = a.n / CG;
  const int a_bytes = ((16 * a.a_pitch + a.a_off_rows + 8) * a.row_bytes + 1023) / 1024 * 1024, b_bytes = b_rows * a.row_bytes;
  // operands: small non-trivial fp16 values (0x2c00 = 0.0625)
  for (int i = threadIdx.x; i < a.stages * (a_bytes + b_bytes) / 2; i += blockDim.x) {
    uint32_t hsh = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    hsh ^= hsh >> 15; hsh *= 2246822519u; hsh ^= hsh >> 13;
    // random: sign + exponent 0x38..0x3f (0.5 .. 2) + random mantissa; else small near-constant values (0x2c00 = 0.0625)
    ((uint16_t*)smem)[i] = a.random ? (uint16_t)((hsh & 0x8000u) | 0x3800u | (hsh & 0x07ffu)) : (uint16_t)(0x2c00 + (i & 7));
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;

  // instruction descriptor: D = f32, A = B = f16, K-major both, N >> 3 at [17,23), M >> 4 at [24,29)
  const uint32_t idesc = (1u << 4) | ((uint32_t)(a.n >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
  const uint32_t a0 = smem_u32(smem), b0 = a0 + a.stages * a_bytes;

  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (cta_rank == 0) {
      t0 = clock64();
      if (elect_one()) {
        // descriptors differ only in the start-address field (low word): one add per stage, an immediate per k step,
        // so the loop costs a few instructions per MMA (earlier versions measured their own address arithmetic:
        // 240 and 127 clocks per iteration whatever N)
        uint64_t ad0 = make_desc(a0 + (uint32_t)(a.a_off_rows * a.row_bytes), a.row_bytes, a.layout);
        ad0 = (ad0 & ~((uint64_t)0x3FFF << 32)) | ((uint64_t)(((a.a_pitch * a.row_bytes) >> 4) & 0x3FFF) << 32);   // SBO = pitch rows
        const uint64_t bd0 = make_desc(b0, a.row_bytes, a.layout);
        const uint32_t a_hi = (uint32_t)(ad0 >> 32), b_hi = (uint32_t)(bd0 >> 32);
        const uint32_t a_lo0 = (uint32_t)ad0, b_lo0 = (uint32_t)bd0;
        const uint32_t a_st = (uint32_t)a_bytes >> 4, b_st = (uint32_t)b_bytes >> 4;
#pragma unroll 1
        for (int i = 0; i < a.nmma / KS; ++i) {
          const uint32_t st = (uint32_t)i & (uint32_t)(a.stages - 1);
          const uint32_t a_lo = a_lo0 + st * a_st, b_lo = b_lo0 + st * b_st;
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            const uint64_t ad = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + 2 * k);
            const uint64_t bd = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + 2 * k);
            const uint32_t acc = (i | k) ? 1u : 0u;
            if (CG == 1) {
              asm volatile(
                  "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                  ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
            } else {
              asm volatile(
                  "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                  ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
            }
          }
        }
      }
      __syncwarp();
      {
        if (elect_one()) {
        if (CG == 1) {
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        } else {
          const uint16_t mask = 3;
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                       ::"r"(smem_u32(&bar)), "h"(mask) : "memory");
        }
        }
      }
      __syncwarp();
    }
    // both CTAs of a pair wait for the commit (the peer's barrier gets the multicast arrival)
    uint32_t it = 0;
    while (!mbar_try_wait(smem_u32(&bar), 0)) {
      if (++it > (1u << 26)) { if (lane == 0) printf("probe timeout block %d\n", blockIdx.x); __trap(); }
    }
    t1 = clock64();
    if (lane == 0 && cta_rank == 0) a.cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (lane == 0 && cta_rank != 0) a.cycles[blockIdx.x] = 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

static int g_a_off_rows = 0, g_a_pitch = 8, g_random = 0;
template <int CG, int KS>
static double run_ks(int grid, int n, int nmma, float* ms_out) {
  const int row_bytes = 32 * KS;
  Args a;
  a.n = n; a.nmma = nmma; a.row_bytes = row_bytes;
  a.layout = row_bytes == 128 ? 2 : row_bytes == 64 ? 4 : 6;
  a.stages = 4;
  a.a_off_rows = g_a_off_rows; a.a_pitch = g_a_pitch; a.random = g_random;
  const int a_tile = ((16 * a.a_pitch + a.a_off_rows + 8) * row_bytes + 1023) / 1024 * 1024;
  const int smem = a.stages * (a_tile + (n / CG) * row_bytes) + 1024;
  CK(cudaMalloc(&a.cycles, grid * sizeof(unsigned long long)));
  CK(cudaMemset(a.cycles, 0, grid * sizeof(unsigned long long)));
  CK(cudaFuncSetAttribute(k_probe<CG, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, k_probe<CG, KS>, a));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  CK(cudaEventElapsedTime(ms_out, e0, e1));
  unsigned long long* h = (unsigned long long*)malloc(grid * sizeof(unsigned long long));
  CK(cudaMemcpy(h, a.cycles, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  double sum = 0; int cnt = 0;
  for (int i = 0; i < grid; ++i) if (h[i]) { sum += (double)h[i]; ++cnt; }
  free(h);
  CK(cudaFree(a.cycles));
  return cnt ? sum / cnt / nmma : 0.0;
}

template <int CG>
static double run(int grid, int n, int row_bytes, int nmma, float* ms_out) {
  if (row_bytes == 128) return run_ks<CG, 4>(grid, n, nmma, ms_out);
  if (row_bytes == 64) return run_ks<CG, 2>(grid, n, nmma, ms_out);
  return run_ks<CG, 1>(grid, n, nmma, ms_out);
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("# %s, %d SMs\n", p.name, p.multiProcessorCount);
  printf("# clocks per tcgen05.mma (K = 16, f16 x f16 -> f32, both operands in shared memory); TF/s = whole chip from the event time\n");
  printf("%-10s %5s %5s %9s %6s %12s %10s %10s\n", "cta_group", "M", "N", "row_bytes", "grid", "clk_per_mma", "floor_clk", "TFLOP/s");
  const int nmma = 8192;
  const int grids[2] = {2, p.multiProcessorCount & ~1};
  // A-operand addressing of the halo conv kernels: start offsets that are not a multiple of the 8-row swizzle atom, and
  // 8-row groups 10 rows apart (the [8+2]-pixel staged box).  Full chip, 128-byte rows.
  printf("# A start offset / group pitch (rows of 128 B), grid = %d: clocks per MMA\n", grids[1]);
  printf("%-10s %5s %8s %8s %12s\n", "cta_group", "N", "off_rows", "pitch", "clk_per_mma");
  const int offs[5][2] = {{0, 8}, {1, 8}, {0, 10}, {1, 10}, {11, 10}};
  for (int oi = 0; oi < 5; ++oi) {
    g_a_off_rows = offs[oi][0]; g_a_pitch = offs[oi][1];
    for (int n = 32; n <= 256; n <<= 1) {
      float ms;
      const double c1 = run<1>(grids[1], n, 128, nmma, &ms);
      printf("%-10d %5d %8d %8d %12.1f\n", 1, n, g_a_off_rows, g_a_pitch, c1);
      if (n == 64 || n == 128) {
        const double c2 = run<2>(grids[1], n, 128, nmma, &ms);
        printf("%-10d %5d %8d %8d %12.1f\n", 2, n, g_a_off_rows, g_a_pitch, c2);
      }
    }
  }
  g_a_off_rows = 0; g_a_pitch = 8;
  // operand data: the tensor pipe's power draw, hence the clock the chip grants it, depends on the switching activity of
  // the operands.  Same MMA stream (262144 MMAs per CTA, all SMs), near-constant vs pseudo-random fp16 data: clocks per MMA
  // from clock64 (SM cycles) and TFLOP/s from the CUDA-event time -> effective SM clock = cycles / time.
  printf("# operand data vs effective clock (grid = %d, 262144 MMAs per CTA)\n", grids[1]);
  printf("%-10s %5s %8s %12s %10s %12s\n", "cta_group", "N", "data", "clk_per_mma", "TFLOP/s", "eff_clk_GHz");
  for (int rnd = 0; rnd < 2; ++rnd) {
    g_random = rnd;
    for (int n = 64; n <= 256; n <<= 1) {
      float ms;
      const int nm = 262144;
      const double c1 = run<1>(grids[1], n, 128, nm, &ms);
      const double tf = 2.0 * 128 * n * 16 * (double)nm * grids[1] / (ms * 1e-3) / 1e12;
      printf("%-10d %5d %8s %12.1f %10.1f %12.3f\n", 1, n, rnd ? "random" : "const", c1, tf, c1 * nm / (ms * 1e-3) / 1e9);
    }
  }
  g_random = 0;
  if (getenv("MMA_PROBE_OFFSETS_ONLY")) return 0;
  for (int gi = 0; gi < 2; ++gi) {
    for (int rb = 128; rb >= 32; rb >>= 1) {
      for (int n = 32; n <= 256; n <<= 1) {
        float ms;
        double c1 = run<1>(grids[gi], n, rb, nmma, &ms);
        double tf = 2.0 * 128 * n * 16 * (double)nmma * grids[gi] / (ms * 1e-3) / 1e12;
        printf("%-10d %5d %5d %9d %6d %12.1f %10.1f %10.1f\n", 1, 128, n, rb, grids[gi], c1, 128.0 * n / 256.0, tf);
        if (n >= 32) {
          double c2 = run<2>(grids[gi], n, rb, nmma, &ms);
          double tf2 = 2.0 * 256 * n * 16 * (double)nmma * (grids[gi] / 2) / (ms * 1e-3) / 1e12;
          printf("%-10d %5d %5d %9d %6d %12.1f %10.1f %10.1f\n", 2, 256, n, rb, grids[gi], c2, 128.0 * n / 256.0, tf2);
        }
      }
    }
  }
  return 0;
}
