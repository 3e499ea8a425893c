// What a packed fp32x2 instruction costs in the operand pattern of scan_rp_kernel's inner loop, on one SM
// sub-partition with 1, 2 or 3 resident warps:
//   mode 0: the loop as it is — h = fma2(P, h, B.bcast); acc = fma2(h, C.bcast, acc); P = mul2(P, RQ)  (per state)
//   mode 1: only the state updates     h = fma2(P, h, B.bcast)
//   mode 2: only the output updates    acc = fma2(h, C.bcast, acc)
//   mode 3: only the power chain       P = mul2(P, RQ)
//   mode 4: mode 0 with B and C as full packed operands (no scalar broadcast)
//   mode 5: mode 0 re-ordered per 16-state quad m: the 8 state updates of both row pairs, then the 8 power
//           multiplies, then the 8 output updates into 8 separate accumulators (4 per row pair)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scan_mix_probe scan_mix_probe.cu && ./scan_mix_probe
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int MODE>
__global__ void probe(long long* out, const float* __restrict__ src, int iters) {
  u64 H[2][16], P[2][4], acc[2][4], RQ[2];
  float bv[16], cv[16];
  for (int p = 0; p < 2; ++p) {
    for (int k = 0; k < 16; ++k) H[p][k] = pack2(src[k + p], src[k + 2 + p]);
    for (int q = 0; q < 4; ++q) P[p][q] = pack2(0.999f + src[q], 0.998f + src[q + 1]);
    acc[p][0] = acc[p][1] = acc[p][2] = acc[p][3] = 0ull;
    RQ[p] = pack2(0.9999f + src[p], 0.9998f);
  }
  for (int k = 0; k < 16; ++k) { bv[k] = src[32 + k + threadIdx.x % 3]; cv[k] = src[48 + k + threadIdx.x % 5]; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 5) {
#pragma unroll
      for (int m = 0; m < 4; ++m) {
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
          for (int q = 0; q < 4; ++q) H[p][4 * m + q] = fma2(P[p][q], H[p][4 * m + q], pack2(bv[4 * m + q], bv[4 * m + q]));
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
          for (int q = 0; q < 4; ++q) P[p][q] = mul2(P[p][q], RQ[p]);
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[p][q] = fma2(H[p][4 * m + q], pack2(cv[4 * m + q], cv[4 * m + q]), acc[p][q]);
      }
      continue;
    }
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = 4 * m + q;
          const u64 bb = MODE == 4 ? pack2(bv[k], cv[k]) : pack2(bv[k], bv[k]);
          const u64 cc = MODE == 4 ? pack2(cv[k], bv[k]) : pack2(cv[k], cv[k]);
          if (MODE == 0 || MODE == 1 || MODE == 4) H[p][k] = fma2(P[p][q], H[p][k], bb);
          if (MODE == 0 || MODE == 2 || MODE == 4) acc[p][q & 1] = fma2(H[p][k], cc, acc[p][q & 1]);
          if (MODE == 6) acc[p][q & 1] = fma2(cc, H[p][k], acc[p][q & 1]);           // scalar operand first
          if (MODE == 7) acc[p][q] = fma2(H[p][k], cc, acc[p][q]);                   // four accumulators per pair
          if (MODE == 8) { H[p][k] = fma2(H[p][k], P[p][q], bb); acc[p][q] = fma2(cc, H[p][k], acc[p][q]); P[p][q] = mul2(RQ[p], P[p][q]); }
          if (MODE == 0 || MODE == 3 || MODE == 4) P[p][q] = mul2(P[p][q], RQ[p]);
        }
  }
  long long t1 = clock64();
  u64 s = 0;
  for (int p = 0; p < 2; ++p) { for (int k = 0; k < 16; ++k) s ^= H[p][k]; for (int q = 0; q < 4; ++q) s ^= P[p][q] ^ acc[p][q]; }
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (s == 0x1234) out[1000] = 1;
}
template <int MODE>
void run(long long* d, const float* src, int per_iter, const char* name) {
  const int iters = 4000;
  for (int warps : {4, 8, 12}) {
    probe<MODE><<<1, warps * 32>>>(d, src, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("%-40s warps/SMSP=%d: %.2f clk per packed instruction per SMSP\n", name, warps / 4,
           (double)c / ((double)iters * per_iter * (warps / 4)));
  }
}
int main() {
  long long* d; cudaMalloc(&d, 8192 * 8);
  float h[128]; for (int i = 0; i < 128; ++i) h[i] = 0.001f * i;
  float* src; cudaMalloc(&src, sizeof(h)); cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>(d, src, 96, "full mix (broadcast B/C)");
  run<1>(d, src, 32, "state updates only");
  run<2>(d, src, 32, "output updates only");
  run<3>(d, src, 32, "power chain only");
  run<4>(d, src, 96, "full mix, packed B/C operands");
  run<5>(d, src, 96, "full mix, grouped order, 8 accumulators");
  run<6>(d, src, 32, "output updates only, scalar operand first");
  run<7>(d, src, 32, "output updates only, 4 accumulators per pair");
  run<8>(d, src, 96, "full mix, operands swapped, 8 accumulators");
  return 0;
}
