// norm_conv.cu — LayerNorm and the SSM block prologue (LayerNorm -> causal depthwise conv).
//
// Reference: F.layer_norm (eps 1e-5, biased variance) at model.py:200,224, ssm.py:408,423,504,
// attention.py:306-307; depthwise Conv1d(k, padding=k-1)[..., :L] at ssm.py:411-414.
// Both are streaming, HBM-bound kernels: one warp owns one token (C <= 1024 channels held in
// registers), rows are read and written once with fully coalesced 128-byte warp accesses.
#include "common.cuh"
#include "kernels.cuh"

namespace vasr {

namespace {

constexpr int LN_MAX_PER_LANE = 8;  // C <= 256 ... the model uses 192 (6 per lane)

// normalises one row held as v[i] = x[lane + 32*i]; returns through v.
template <int PER>
__device__ __forceinline__ void warp_layer_norm(float (&v)[PER], int C, int lane, const float* __restrict__ gamma,
                                                const float* __restrict__ beta) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) s += (lane + 32 * i < C) ? v[i] : 0.f;
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    float d = (lane + 32 * i < C) ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    int c = lane + 32 * i;
    if (c < C) v[i] = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
  }
}

__global__ void __launch_bounds__(256) layer_norm_kernel(const float* __restrict__ x, int64_t ldx,
                                                         float* __restrict__ y, int64_t ldy,
                                                         const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int64_t M, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  float v[LN_MAX_PER_LANE];
  const float* xr = x + m * ldx;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    int c = lane + 32 * i;
    v[i] = c < C ? xr[c] : 0.f;
  }
  warp_layer_norm<LN_MAX_PER_LANE>(v, C, lane, gamma, beta);
  float* yr = y + m * ldy;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    int c = lane + 32 * i;
    if (c < C) yr[c] = v[i];
  }
}

// One CTA = TT consecutive tokens of one utterance (+ k-1 halo tokens on the left).
constexpr int TT = 32;
constexpr int MAXK = 8;

__global__ void __launch_bounds__(256) ln_dwconv_kernel(const float* __restrict__ x, float* __restrict__ u,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta,
                                                        const float* __restrict__ w,
                                                        const float* __restrict__ bias, int64_t L, int C,
                                                        int k) {
  extern __shared__ float sx[];  // (TT + k - 1) x C, normalised rows; rows before t = 0 are zero
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = blockIdx.y;
  const int64_t t0 = (int64_t)blockIdx.x * TT;
  const int halo = k - 1;
  const int rows = TT + halo;
  for (int r = warp; r < rows; r += 8) {
    const int64_t t = t0 - halo + r;
    float v[LN_MAX_PER_LANE];
    const bool live = t >= 0 && t < L;
    if (live) {
      const float* xr = x + (b * L + t) * C;
#pragma unroll
      for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
        int c = lane + 32 * i;
        v[i] = c < C ? xr[c] : 0.f;
      }
      warp_layer_norm<LN_MAX_PER_LANE>(v, C, lane, gamma, beta);
    }
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
      int c = lane + 32 * i;
      if (c < C) sx[r * C + c] = live ? v[i] : 0.f;
    }
  }
  __syncthreads();
  // out[t, c] = bias[c] + sum_j w[c, j] * xn[t + j - (k-1), c]  ==  sx[(t - t0) + j][c]
  for (int idx = threadIdx.x; idx < TT * C; idx += 256) {
    const int tt = idx / C, c = idx - tt * C;
    const int64_t t = t0 + tt;
    if (t >= L) break;
    float acc = __ldg(bias + c);
    for (int j = 0; j < k; ++j) acc = fmaf(__ldg(w + c * k + j), sx[(tt + j) * C + c], acc);
    u[(b * L + t) * C + c] = acc;
  }
}

}  // namespace

cudaError_t launch_layer_norm(const float* x, int64_t ldx, float* y, int64_t ldy, const float* gamma,
                              const float* beta, int64_t M, int C, cudaStream_t s, int64_t* launches) {
  if (M <= 0) return cudaSuccess;
  if (C > 32 * LN_MAX_PER_LANE) return cudaErrorInvalidValue;
  layer_norm_kernel<<<(unsigned)((M + 7) / 8), 256, 0, s>>>(x, ldx, y, ldy, gamma, beta, M, C);
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_ln_dwconv(const float* x, float* u, const float* gamma, const float* beta, const float* w,
                             const float* bias, int64_t B, int64_t L, int C, int k, cudaStream_t s,
                             int64_t* launches) {
  if (B <= 0 || L <= 0) return cudaSuccess;
  if (C > 32 * LN_MAX_PER_LANE || k < 1 || k > MAXK || B > 65535) return cudaErrorInvalidValue;
  dim3 grid((unsigned)((L + TT - 1) / TT), (unsigned)B);
  size_t smem = (size_t)(TT + k - 1) * C * sizeof(float);
  ln_dwconv_kernel<<<grid, 256, smem, s>>>(x, u, gamma, beta, w, bias, L, C, k);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace vasr
