// In-switch all-reduce of the gradient accumulators over NVLink multicast (NVLS): the dense route of the
// data-parallel train step (hopwise_b200/distributed.py), replacing DDP's dense all-reduce of every embedding
// table (trainer/trainer.py:82-112) and, on this route, NCCL.
//
// Every rank maps the same symmetric buffer and holds a multicast address that names all N copies at once.
// Rank r owns the r-th slice of the buffer: one multimem.ld_reduce per 16 bytes makes the NVSwitch fetch the N
// copies and return their sum, one multimem.st writes the sum back into all N copies.  Each element is reduced
// exactly once, by one rank, and broadcast: the replicas receive bit-identical sums, and a GPU sends and
// receives ~one buffer per step instead of the 2(N-1)/N buffers a ring moves through its links.  The caller
// brackets the launch with a cross-rank barrier on each side (all gradients written before / all slices
// reduced after).
#include "common.cuh"

namespace {

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(256) multimem_all_reduce_kernel(float* __restrict__ mc, int64_t q_lo, int64_t q_hi) {
  // q = index of a 16-byte quad; four quads in flight per thread
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t q = q_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; q + 3 * stride < q_hi; q += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = multimem_ld_reduce_add(mc + 4 * (q + u * stride));
#pragma unroll
    for (int u = 0; u < 4; ++u) multimem_st(mc + 4 * (q + u * stride), v[u]);
  }
  for (; q < q_hi; q += stride) multimem_st(mc + 4 * q, multimem_ld_reduce_add(mc + 4 * q));
  __threadfence_system();
}

}  // namespace

extern "C" int kge_multimem_all_reduce_f32(void* multicast_ptr, int64_t n_floats, int32_t rank, int32_t world,
                                           kge_stream_t stream) {
  KGE_REQUIRE(multicast_ptr && n_floats >= 0 && world >= 1 && rank >= 0 && rank < world, KGE_E_ARG,
              "bad multimem all-reduce arguments");
  KGE_REQUIRE((reinterpret_cast<uintptr_t>(multicast_ptr) & 15) == 0 && (n_floats & 3) == 0, KGE_E_ARG,
              "multimem all-reduce needs a 16-byte aligned buffer of a multiple of 4 floats");
  const int64_t quads = n_floats / 4;
  const int64_t per = (quads + world - 1) / world;
  const int64_t lo = per * rank, hi = lo + per < quads ? lo + per : quads;
  if (hi <= lo) return 0;
  const int threads = 256;
  int64_t grid = (hi - lo + 4 * threads - 1) / (4 * threads);
  const int64_t cap = (int64_t)kge_num_sms() * 4;
  if (grid > cap) grid = cap;
  multimem_all_reduce_kernel<<<(unsigned)grid, threads, 0, (cudaStream_t)stream>>>(reinterpret_cast<float*>(multicast_ptr),
                                                                                 lo, hi);
  KGE_LAUNCH_CHECK();
  return 0;
}
