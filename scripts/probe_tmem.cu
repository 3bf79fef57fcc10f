// Microbenchmark: tcgen05.ld (TMEM -> registers) bandwidth per SM on sm_100a, by shape / packing /
// warps per SM.  Not part of the product; its numbers size the full-sort epilogue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_tmem scripts/probe_tmem.cu && ./probe_tmem
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(R, ADDR)                                                                                              \
  asm volatile(                                                                                                    \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                     \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                     \
      : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]), "=r"(R[8]),  \
        "=r"(R[9]), "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15]), "=r"(R[16]),      \
        "=r"(R[17]), "=r"(R[18]), "=r"(R[19]), "=r"(R[20]), "=r"(R[21]), "=r"(R[22]), "=r"(R[23]), "=r"(R[24]),     \
        "=r"(R[25]), "=r"(R[26]), "=r"(R[27]), "=r"(R[28]), "=r"(R[29]), "=r"(R[30]), "=r"(R[31])                   \
      : "r"(ADDR)                                                                                                  \
      : "memory")
#define LD32P(R, ADDR)                                                                                             \
  asm volatile(                                                                                                    \
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "                                                          \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                     \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                     \
      : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]), "=r"(R[8]),  \
        "=r"(R[9]), "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15]), "=r"(R[16]),      \
        "=r"(R[17]), "=r"(R[18]), "=r"(R[19]), "=r"(R[20]), "=r"(R[21]), "=r"(R[22]), "=r"(R[23]), "=r"(R[24]),     \
        "=r"(R[25]), "=r"(R[26]), "=r"(R[27]), "=r"(R[28]), "=r"(R[29]), "=r"(R[30]), "=r"(R[31])                   \
      : "r"(ADDR)                                                                                                  \
      : "memory")
#define LD16x256(R, ADDR)                                                                                          \
  asm volatile(                                                                                                    \
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "                                                                    \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                     \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                     \
      : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]), "=r"(R[8]),  \
        "=r"(R[9]), "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15]), "=r"(R[16]),      \
        "=r"(R[17]), "=r"(R[18]), "=r"(R[19]), "=r"(R[20]), "=r"(R[21]), "=r"(R[22]), "=r"(R[23]), "=r"(R[24]),     \
        "=r"(R[25]), "=r"(R[26]), "=r"(R[27]), "=r"(R[28]), "=r"(R[29]), "=r"(R[30]), "=r"(R[31])                   \
      : "r"(ADDR)                                                                                                  \
      : "memory")

// mode 0: 32x32b.x32 (32 columns of this lane's row, 4 KB per warp instruction)
// mode 1: 32x32b.x32.pack::16b (64 columns, low halves packed)
// mode 2: 16x256b.x8
template <int MODE>
__global__ void __launch_bounds__(1024, 1) ld_kernel(int iters, unsigned long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const unsigned long long c0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r0[32], r1[32];
    const uint32_t col = (uint32_t)((i * 64 + (warp >> 2) * 128) & 255);
    if (MODE == 0) {
      LD32(r0, base + col);
      LD32(r1, base + col + 32);
    } else if (MODE == 1) {
      LD32P(r0, base + col);
      LD32P(r1, base + col + 64);
    } else {
      LD16x256(r0, base + col);
      LD16x256(r1, base + col + 64);
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= r0[j] ^ r1[j];
  }
  const unsigned long long c1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = c1 - c0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

template <int MODE>
void run(const char* name, int warps, int iters) {
  unsigned long long* cyc;
  uint32_t* sink;
  cudaMalloc(&cyc, 148 * 8);
  cudaMalloc(&sink, 4);
  ld_kernel<MODE><<<148, warps * 32>>>(16, cyc, sink);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  ld_kernel<MODE><<<148, warps * 32>>>(iters, cyc, sink);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  // register-side bytes per warp iteration: 2 loads x 32 regs x 32 lanes x 4 B = 8 KB
  const double bytes = (double)warps * iters * 8192.0;
  printf("%-28s warps/SM=%2d  %s  cycles=%llu  reg-side B/clk/SM=%.1f  (%.3f ms)\n", name, warps, cudaGetErrorString(err),
         h[0], bytes / (double)h[0], ms);
  cudaFree(cyc);
  cudaFree(sink);
}

int main() {
  for (int w : {4, 8, 16, 32}) run<0>("32x32b.x32", w, 4096);
  for (int w : {4, 8, 16}) run<1>("32x32b.x32.pack::16b", w, 4096);
  for (int w : {4, 8, 16}) run<2>("16x256b.x8", w, 4096);
  return 0;
}
