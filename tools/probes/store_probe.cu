// Store-path micro-benchmark: how fast can 268 MB (the C4 first-layer output, 8 x 1024^2 x 16 fp16) be WRITTEN
// from registers, by store shape?  Settles the roofline of the write-bound layers (conv0 / conv1).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probes/store_probe tools/probes/store_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void st256(void* p, unsigned v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(p), "r"(v) : "memory");
}

// mode 0: 16 B per lane, lanes consecutive (512 B contiguous per warp instruction) - what a memset does
// mode 1: 32 B per lane (st.v8), lanes consecutive (1 KB contiguous per warp instruction) - conv epilogue, 1 px / lane
// mode 2: 32 B per lane, lanes 128 B apart (k_conv_first: 4 px per thread)
// mode 3: 32 B per lane, lanes 256 B apart (Toeplitz first layer: 8 px per thread)
template <int MODE>
__global__ void k_store(char* out, size_t bytes) {
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const size_t chunk = MODE == 0 ? 512 : (MODE == 1 ? 1024 : (MODE == 2 ? 4096 : 8192));
  for (size_t base = warp * chunk; base + chunk <= bytes; base += n_warps * chunk) {
    if (MODE == 0) {
      *reinterpret_cast<uint4*>(out + base + lane * 16) = make_uint4(1, 2, 3, 4);
    } else if (MODE == 1) {
      st256(out + base + lane * 32, 7u);
    } else if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) st256(out + base + lane * 128 + j * 32, 7u);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) st256(out + base + lane * 256 + j * 32, 7u);
    }
  }
}

template <int MODE>
float run(char* buf, size_t bytes, int blocks, int threads) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) k_store<MODE><<<blocks, threads>>>(buf, bytes);
  cudaEventRecord(e0);
  const int reps = 10;
  for (int i = 0; i < reps; ++i) k_store<MODE><<<blocks, threads>>>(buf, bytes);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps * 1e3f;
}

int main() {
  const size_t bytes = (size_t)8 * 1024 * 1024 * 32;     // 268 MB > 126 MB L2
  char* buf;
  cudaMalloc(&buf, bytes);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int cfgs[4][2] = {{sms * 2, 256}, {sms * 4, 256}, {sms * 8, 256}, {sms * 16, 128}};
  for (auto& c : cfgs) {
    float t0 = run<0>(buf, bytes, c[0], c[1]), t1 = run<1>(buf, bytes, c[0], c[1]), t2 = run<2>(buf, bytes, c[0], c[1]),
          t3 = run<3>(buf, bytes, c[0], c[1]);
    printf("grid %5d x %3d: 16B/lane contiguous %.1f us (%.2f TB/s) | 32B/lane contiguous %.1f us (%.2f TB/s) | 32B @128B stride %.1f us (%.2f TB/s) | 32B @256B stride %.1f us (%.2f TB/s)\n",
           c[0], c[1], t0, bytes / t0 / 1e6, t1, bytes / t1 / 1e6, t2, bytes / t2 / 1e6, t3, bytes / t3 / 1e6);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
