// context.cu — streaming kernels of the hierarchical global context
// (velocity_asr/attention.py): adaptive average pooling (:63-78), the softmax(QK^T/sqrt(hd))V
// core of the 4-head cross-attention over K2 <= 64 pooled keys (:143-161), and the gated mix
// g*local_t + (1-g)*global_t (:208-215).  The projections around them run through gemm.cu.
#include "common.cuh"
#include "kernels.cuh"

namespace vasr {

namespace {

// F.adaptive_avg_pool1d: window i = [floor(i L / K), ceil((i+1) L / K))
__global__ void __launch_bounds__(256) adaptive_pool_kernel(const float* __restrict__ x, int64_t ldx,
                                                            float* __restrict__ out, int64_t L, int64_t K,
                                                            int C, const int32_t* __restrict__ rag, int f_in,
                                                            int f_out) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = blockIdx.x, b = blockIdx.y;
  const int64_t Lb = rag ? rag[b * RAG_STRIDE + f_in] : L;      // rows keep the strides of L and K
  const int64_t Kb = rag ? rag[b * RAG_STRIDE + f_out] : K;
  if (i >= Kb) {
    for (int c = threadIdx.x; c < C; c += 256) out[(b * K + i) * C + c] = 0.f;
    return;
  }
  const int64_t s = (i * Lb) / Kb;
  const int64_t e = ((i + 1) * Lb + Kb - 1) / Kb;
  const float inv = 1.0f / (float)(e - s);
  for (int c = threadIdx.x; c < C; c += 256) {
    float acc = 0.f;
    for (int64_t t = s; t < e; ++t) acc += x[(b * L + t) * ldx + c];
    out[(b * K + i) * C + c] = acc * inv;
  }
}

// One thread per (token, head).  K/V rows of the utterance sit in shared memory; the K2 <= 64
// scores are recomputed in the second pass instead of being kept (hd = 12 -> 12 FMAs each).
constexpr int ATT_TOK = 64;
constexpr int ATT_MAX_HD = 16;
__global__ void __launch_bounds__(256) attention_kernel(const float* __restrict__ q, int64_t ldq,
                                                        const float* __restrict__ kv, float* __restrict__ o,
                                                        int64_t ldo, int64_t L, int64_t Kk, int heads, int hd,
                                                        const int32_t* __restrict__ rag) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float skv[];  // Kk x (2*A)
  const int A = heads * hd;
  const int64_t b = blockIdx.y;
  const float* kvb = kv + b * Kk * 2 * A;
  if (rag) Kk = rag[b * RAG_STRIDE + RAG_K2];    // ragged: this utterance has fewer pooled keys
  for (int idx = threadIdx.x; idx < Kk * 2 * A; idx += blockDim.x) skv[idx] = kvb[idx];
  __syncthreads();
  const int tok = threadIdx.x / heads, hh = threadIdx.x % heads;
  const int64_t t = (int64_t)blockIdx.x * ATT_TOK + tok;
  if (tok >= ATT_TOK || t >= L) return;
  const float scale = 1.0f / sqrtf((float)hd);
  float qv[ATT_MAX_HD];
  const float* qr = q + (b * L + t) * ldq + hh * hd;
#pragma unroll
  for (int d = 0; d < ATT_MAX_HD; ++d) qv[d] = d < hd ? qr[d] : 0.f;
  float mx = -INFINITY;
  for (int64_t kx = 0; kx < Kk; ++kx) {
    const float* kr = skv + kx * 2 * A + hh * hd;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < ATT_MAX_HD; ++d)
      if (d < hd) s = fmaf(qv[d], kr[d], s);
    mx = fmaxf(mx, s * scale);
  }
  float den = 0.f, acc[ATT_MAX_HD];
#pragma unroll
  for (int d = 0; d < ATT_MAX_HD; ++d) acc[d] = 0.f;
  for (int64_t kx = 0; kx < Kk; ++kx) {
    const float* kr = skv + kx * 2 * A + hh * hd;
    const float* vr = kr + A;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < ATT_MAX_HD; ++d)
      if (d < hd) s = fmaf(qv[d], kr[d], s);
    const float p = expf(s * scale - mx);
    den += p;
#pragma unroll
    for (int d = 0; d < ATT_MAX_HD; ++d)
      if (d < hd) acc[d] = fmaf(p, vr[d], acc[d]);
  }
  const float inv = 1.0f / den;
  float* orow = o + (b * L + t) * ldo + hh * hd;
#pragma unroll
  for (int d = 0; d < ATT_MAX_HD; ++d)
    if (d < hd) orow[d] = acc[d] * inv;
}

__global__ void __launch_bounds__(256) gate_mix_kernel(const float* __restrict__ f3, float* __restrict__ out,
                                                       int64_t M, int C) {
  pdl_trigger();
  pdl_wait();
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= M * C) return;
  const int64_t m = idx / C;
  const int c = (int)(idx - m * C);
  const float* r = f3 + m * 3 * C;
  const float g = sigmoid_f(r[c]);
  out[idx] = g * r[C + c] + (1.0f - g) * r[2 * C + c];
}

// calibration of an output FakeQuantize node: one CTA, fixed reduction order -> run-to-run identical
__global__ void __launch_bounds__(1024) minmax_kernel(const float* __restrict__ x, int64_t ldx, int64_t M, int c0,
                                                      int nc, float* __restrict__ mm) {
  __shared__ float smin[32], smax[32];
  float lo = INFINITY, hi = -INFINITY;
  const int64_t total = M * nc;
  for (int64_t i = threadIdx.x; i < total; i += 1024) {
    const float v = x[(i / nc) * ldx + c0 + (int)(i % nc)];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    lo = smin[threadIdx.x];
    hi = smax[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (threadIdx.x == 0) { mm[0] = lo; mm[1] = hi; }
  }
}

__global__ void set_qparams_kernel(const float* __restrict__ mm, float* __restrict__ qs, float* __restrict__ qz,
                                   int c0, int nc) {
  // scale = (max - min) / (qmax - qmin), clamped to >= 1e-10; zp = qmin - min / scale  (quantize.py:116-121)
  const float scale = fmaxf(__fdiv_rn(__fsub_rn(mm[1], mm[0]), 255.f), 1e-10f);
  const float zp = __fsub_rn(0.f, __fdiv_rn(mm[0], scale));
  for (int c = threadIdx.x; c < nc; c += blockDim.x) {
    qs[c0 + c] = scale;
    qz[c0 + c] = zp;
  }
}

}  // namespace

cudaError_t launch_minmax(const float* x, int64_t ldx, int64_t M, int c0, int nc, float* mm, cudaStream_t s,
                          int64_t* launches) {
  if (M <= 0 || nc <= 0) return cudaErrorInvalidValue;
  minmax_kernel<<<1, 1024, 0, s>>>(x, ldx, M, c0, nc, mm);
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_set_qparams(const float* mm, float* q_scale, float* q_zp, int c0, int nc, cudaStream_t s,
                               int64_t* launches) {
  set_qparams_kernel<<<1, 128, 0, s>>>(mm, q_scale, q_zp, c0, nc);
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_adaptive_pool(const float* x, int64_t ldx, float* out, int64_t B, int64_t L, int64_t K,
                                 int C, cudaStream_t s, int64_t* launches, const int32_t* rag, int f_in,
                                 int f_out) {
  if (B <= 0 || K <= 0) return cudaSuccess;
  if (B > 65535) return cudaErrorInvalidValue;
  dim3 grid((unsigned)K, (unsigned)B);
  const cudaError_t e = launch_k(adaptive_pool_kernel, grid, dim3(256), 0, s, x, ldx, out, L, K, C, rag, f_in, f_out);
  if (launches) ++*launches;
  return e;
}

cudaError_t launch_attention(const float* q, int64_t ldq, const float* kv, float* o, int64_t ldo, int64_t B,
                             int64_t L, int64_t Kk, int heads, int hd, cudaStream_t s, int64_t* launches,
                             const int32_t* rag) {
  if (B <= 0 || L <= 0) return cudaSuccess;
  if (hd > ATT_MAX_HD || heads * ATT_TOK > 1024 || B > 65535) return cudaErrorInvalidValue;
  const size_t smem = (size_t)Kk * 2 * heads * hd * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  dim3 grid((unsigned)((L + ATT_TOK - 1) / ATT_TOK), (unsigned)B);
  const cudaError_t e = launch_k(attention_kernel, grid, dim3(heads * ATT_TOK), smem, s, q, ldq, kv, o, ldo, L, Kk, heads, hd, rag);
  if (launches) ++*launches;
  return e;
}

cudaError_t launch_gate_mix(const float* f3, float* out, int64_t M, int C, cudaStream_t s, int64_t* launches) {
  if (M <= 0) return cudaSuccess;
  const cudaError_t e = launch_k(gate_mix_kernel, dim3((unsigned)((M * C + 255) / 256)), dim3(256), 0, s, f3, out, M, C);
  if (launches) ++*launches;
  return e;
}

}  // namespace vasr
