// gemm.cu — fp32 CUDA-core projection kernel: C = epi(A · Wᵀ), both operands K-contiguous.
//
// The operator behind vasr_linear (include/vasr.h): a general F.linear for views the tensor-core kernel
// (gemm_tc.cu) cannot take.  The model's own launch sequence runs every projection on gemm_tc.cu and never
// comes here.  Full fp32 (FFMA2: two fp32 FMAs per issue slot, paired along k so no
// operand duplication is needed); results sit within a few ulp of the reference's SGEMM,
// which is what keeps argmax token ids stable.  128x64x16 tiles, 256 threads, 8x4 outputs per
// thread, double-buffered shared memory with register-staged global prefetch.
#include "common.cuh"
#include "kernels.cuh"

namespace vasr {

namespace {

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int LDT = BK + 2;  // 18-float rows: 8-byte LDS of 16 rows hit 32 distinct banks

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__global__ void __launch_bounds__(256, 2) gemm_tn_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][BM * LDT];
  __shared__ __align__(16) float Bs[2][BN * LDT];

  const int tid = threadIdx.x;
  const int tx = tid & 15;   // n group: outputs n0 + 4*tx + j
  const int ty = tid >> 4;   // m group: outputs m0 + ty + 16*i
  // 1-D grid, n tile fastest: CTAs that share an A tile are scheduled together (L2 reuse)
  const unsigned n_tiles = (unsigned)((g.N + BN - 1) / BN);
  const int64_t m0 = (int64_t)(blockIdx.x / n_tiles) * BM;
  const int64_t n0 = (int64_t)(blockIdx.x % n_tiles) * BN;
  const int K = (int)g.K;

  // ---- global -> register staging (A: 2 float4 per thread, W: 1 float4 per thread)
  const int lrow = tid >> 2;        // 0..63
  const int lc4 = (tid & 3) * 4;    // 0,4,8,12
  const float* a_ptr[2];
  bool a_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int64_t m = m0 + lrow + 64 * i;
    a_ok[i] = m < g.M;
    int64_t mm = a_ok[i] ? m : 0;
    int64_t off = g.rows_per_batch > 0
                      ? (mm / g.rows_per_batch) * g.batch_stride + (mm % g.rows_per_batch) * g.lda
                      : mm * g.lda;
    a_ptr[i] = g.A + off + lc4;
  }
  const int64_t nb = n0 + lrow;
  const bool b_ok = nb < g.N;
  const float* b_ptr = g.W + (b_ok ? nb : 0) * (int64_t)K + lc4;
  // W row n_local is stored at smem row (n_local >> 2) + 16 * (n_local & 3) so that thread tx
  // finds its four output columns 4*tx + j at rows tx + 16*j (conflict-free 8-byte reads).
  const int b_srow = (lrow >> 2) + 16 * (lrow & 3);

  float4 ra[2], rb;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) ra[i] = a_ok[i] ? ldg4(a_ptr[i] + k0) : zero4;
    rb = b_ok ? ldg4(b_ptr + k0) : zero4;
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float* d = &As[buf][(lrow + 64 * i) * LDT + lc4];
      *reinterpret_cast<float2*>(d) = make_float2(ra[i].x, ra[i].y);
      *reinterpret_cast<float2*>(d + 2) = make_float2(ra[i].z, ra[i].w);
    }
    float* d = &Bs[buf][b_srow * LDT + lc4];
    *reinterpret_cast<float2*>(d) = make_float2(rb.x, rb.y);
    *reinterpret_cast<float2*>(d + 2) = make_float2(rb.z, rb.w);
  };

  u64 acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[i][j] = 0ull;
    }

  const int nk = K / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) gload((kt + 1) * BK);
    const float* as = &As[buf][ty * LDT];
    const float* bs = &Bs[buf][tx * LDT];
#pragma unroll
    for (int kk = 0; kk < BK; kk += 2) {
      u64 a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const u64*>(as + 16 * i * LDT + kk);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const u64*>(bs + 16 * j * LDT + kk);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma2(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  // ---- epilogue
  const int64_t nbase = n0 + 4 * tx;
  if (nbase >= g.N) return;
  const bool vec_ok = (nbase + 3 < g.N) && ((g.ldc & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  float pef[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int64_t n = nbase + j;
    if (n < g.N) {
      if (g.bias) bias[j] = g.bias[n];
      if (g.pe_time && n >= g.pe_half) pef[j] = g.pe_freq[n - g.pe_half];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty + 16 * i;
    if (m >= g.M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = nbase + j;
      float x = hsum2(acc[i][j]) + bias[j];
      if (g.q_scale && n < g.N) x = fake_quant_u8(x, g.q_scale[n], g.q_zp[n]);
      if (n >= g.act_from) x = apply_act(x, g.act);
      if (g.pe_time) {
        if (n < g.pe_half) {
          if (n < g.N) x += g.pe_time[(m % g.pe_rows) * g.pe_half + n];
        } else {
          x += pef[j];
        }
      }
      if (g.resid && n < g.N) x += g.resid[m * g.ldr + n];
      v[j] = x;
    }
    float* c = g.C + m * g.ldc + nbase;
    if (vec_ok) {
      *reinterpret_cast<float4*>(c) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (nbase + j < g.N) c[j] = v[j];
    }
  }
}

}  // namespace

cudaError_t launch_gemm(const GemmArgs& g, cudaStream_t s, int64_t* launches) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  if (g.K <= 0 || (g.K % BK) != 0 || (g.lda & 3) != 0 || (g.batch_stride & 3) != 0 ||
      (reinterpret_cast<uintptr_t>(g.A) & 15) != 0 || (reinterpret_cast<uintptr_t>(g.W) & 15) != 0)
    return cudaErrorInvalidValue;
  const int64_t tiles = ((g.N + BN - 1) / BN) * ((g.M + BM - 1) / BM);
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  gemm_tn_kernel<<<(unsigned)tiles, 256, 0, s>>>(g);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace vasr
