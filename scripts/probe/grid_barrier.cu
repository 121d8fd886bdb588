// Probe: what does a grid-wide barrier cost on B200 (148 CTAs x 320 threads, one per SM), and which part of it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o grid_barrier.so grid_barrier.cu
#include <cuda_runtime.h>
#include <stdint.h>

template <int VARIANT>
__global__ void __launch_bounds__(320, 1) barrier_loop(unsigned* bar, int iters, float* sink) {
  extern __shared__ uint8_t smem[];  // 200 KB: one CTA per SM like the GEMM kernels
  unsigned epoch = 0;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    if (VARIANT == 0 || VARIANT == 1) asm volatile("fence.proxy.async;" ::: "memory");
    if (VARIANT == 0 || VARIANT == 1 || VARIANT == 2) __threadfence();
    __syncthreads();
    ++epoch;
    if (threadIdx.x == 0) {
      const unsigned target = epoch * gridDim.x;
      if (VARIANT == 3) {
        atomicAdd(bar, 1u);
        while (*reinterpret_cast<volatile unsigned*>(bar) < target) {
        }
      } else {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        while (true) {
          unsigned v;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
          if (v >= target) break;
          if (VARIANT == 4) __nanosleep(100);
        }
      }
      if (VARIANT != 3) __threadfence();
    }
    __syncthreads();
    if (VARIANT == 0) asm volatile("fence.proxy.async;" ::: "memory");
    acc += 1.f;
  }
  if (acc < 0.f) sink[0] = acc + smem[0];
}

extern "C" float run_variant(int variant, int iters, unsigned* bar, float* sink) {
  cudaMemset(bar, 0, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int smem = 200 * 1024;
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
#define RUN(V)                                                                                   \
  cudaFuncSetAttribute(barrier_loop<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);      \
  cudaEventRecord(e0);                                                                           \
  barrier_loop<V><<<nsm, 320, smem>>>(bar, iters, sink);                                         \
  cudaEventRecord(e1);
  switch (variant) {
    case 0: { RUN(0) } break;
    case 1: { RUN(1) } break;
    case 2: { RUN(2) } break;
    case 3: { RUN(3) } break;
    default: { RUN(4) } break;
  }
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  return cudaGetLastError() == cudaSuccess ? ms : -1.f;
}
