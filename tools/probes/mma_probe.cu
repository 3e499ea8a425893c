// Rate of tcgen05.mma kind::tf32 (M = 128) on one SM: clocks per instruction for N = 64/128/256, A operand
// from shared memory (SS) or from tensor memory (TS), operands resident (no data movement in the loop).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && ./mma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t a) {
  return (uint64_t)((a & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t p; asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p)); return p != 0;
}
// sw > 0: switch to the other accumulator (TMEM column 0 <-> 128) every sw iterations of four MMAs, the first MMA of
// a group overwriting (accumulate = 0) and a commit closing it — the pattern of a tiled GEMM
template <int TS>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int n, int iters, int sw) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)n >> 3) << 17) | ((128u >> 4) << 24);
  const uint64_t da = umma_desc(smem_u32(smem)), db = umma_desc(smem_u32(smem) + 16384);
  long long t0 = 0, t1 = 0;
  if (threadIdx.x < 32) {
    t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = sw > 0 ? (((uint32_t)(it / sw) & 1u) * 128u) : 0u;
        const bool first = sw > 0 && it % sw == 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t accum = (first && k == 0) ? 0u : 1u;
          if (TS)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(d), "r"(256u + 8u * k), "l"(db + 2ull * k), "r"(idesc), "r"(accum) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(d), "l"(da + 2ull * k), "l"(db + 2ull * k), "r"(idesc), "r"(accum) : "memory");
        }
        if (sw > 0 && (it + 1) % sw == 0 && it + 1 < iters)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    t1 = clock64();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(0u), "r"(512u));
  }
}
int main() {
  long long* d; cudaMalloc(&d, 4096);
  const int iters = 2000, smem = 80 * 1024;
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int sw : {0, 18, 6, 1}) {
    probe<1><<<148, 128, smem>>>(d, 128, iters, sw);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("TS N=128 grid=148, accumulator switch every %2d x 4 MMAs: %.1f clk per MMA  [%s]\n", sw,
           (double)c / (iters * 4.0), cudaGetErrorString(e));
  }
  for (int ts = 0; ts < 2; ++ts)
    for (int n : {64, 128, 256})
      for (int grid : {1, 148}) {
        if (ts) probe<1><<<grid, 128, smem>>>(d, n, iters, 0); else probe<0><<<grid, 128, smem>>>(d, n, iters, 0);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("%s N=%3d grid=%3d: %.1f clk per MMA (128xNx8 tf32)  -> %.0f MAC/clk/SM  [%s]\n", ts ? "TS" : "SS", n, grid,
               (double)c / (iters * 4.0), 128.0 * n * 8 / ((double)c / (iters * 4.0)), cudaGetErrorString(e));
      }
  return 0;
}
