// common.cuh — shared device helpers for libvasr (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libvasr is written for sm_100a (B200) only"
#endif

namespace vasr {

typedef unsigned long long u64;

// ---- packed fp32x2 math (FFMA2/FMUL2/FADD2 on sm_100): two lanes per issue slot ----------
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float hsum2(u64 v) {
  float lo, hi;
  unpack2(v, lo, hi);
  return lo + hi;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- programmatic dependent launch -----------------------------------------------------------------
// Every kernel of the step is launched with programmatic stream serialisation (launch_k below): its CTAs may
// become resident while the previous kernel is still running and park in pdl_wait() until that kernel has
// completed and its writes are visible, so the ~1-2 us of launch latency between dependent kernels (104 per
// step) and the projection kernel's prologue (barrier init, TMEM allocation, cluster sync) overlap the previous
// kernel's tail.  Rules: pdl_trigger() lets the NEXT kernel's CTAs queue up once every CTA of this one has
// called it or exited — first thing in the short multi-wave kernels, but only near the END of the single-wave
// kernels that run long (scan: before its last block; projections: at a CTA's last tile), because CTAs parked
// in griddepcontrol.wait next to a running kernel slow it down; pdl_wait() before the first read of anything an
// earlier kernel wrote and before
// the first global write; every kernel launched this way executes pdl_wait() (that is what makes "previous
// kernel complete" imply "everything before it complete").  Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- activations with the reference's ATen semantics --------------
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// The same function for the projection epilogues, where erff's ~27 instructions per element made the GELU of a
// tile cost as many issue slots as its MMAs:  gelu(x) = relu(x) - |x| * 0.5 erfc(|x| / sqrt 2), and
// 0.5 erfc(a / sqrt 2) = 2^(a q(a) - 1) with q a degree-9 fit on [0, 6] (beyond 6 the term is < 6e-9 |x|).
// Absolute error <= 2.4e-7 over [-12, 12] (half an ulp of the result at |x| ~ 4; the fp32 erf form has 4.5e-7).
#define VASR_GELU_Q(F)                                                                                        \
  F(1.037785637e-06f) F(-1.249625166e-05f) F(8.313555008e-05f) F(-2.954076626e-04f) F(1.900888492e-05f)       \
  F(6.938213017e-03f) F(-5.244373530e-02f) F(-4.592177570e-01f) F(-1.151104808e+00f)
__device__ __forceinline__ float gelu_poly(float x) {
  const float a = fminf(fabsf(x), 6.0f);
  float q = -3.721141795e-08f;
#define VASR_STEP(c) q = fmaf(q, a, c);
  VASR_GELU_Q(VASR_STEP)
#undef VASR_STEP
  const float e = ex2_approx(fmaf(q, a, -1.0f));
  return fmaxf(x, 0.0f) - fabsf(x * e);
}
// two elements at a time: the Horner chain on the packed fp32x2 pipe
__device__ __forceinline__ void gelu_poly2(float& x0, float& x1) {
  const float a0 = fminf(fabsf(x0), 6.0f), a1 = fminf(fabsf(x1), 6.0f);
  const u64 a = pack2(a0, a1);
  u64 q = pack2(-3.721141795e-08f, -3.721141795e-08f);
#define VASR_STEP(c) q = fma2(q, a, pack2(c, c));
  VASR_GELU_Q(VASR_STEP)
#undef VASR_STEP
  q = fma2(q, a, pack2(-1.0f, -1.0f));
  float p0, p1;
  unpack2(q, p0, p1);
  x0 = fmaxf(x0, 0.0f) - fabsf(x0 * ex2_approx(p0));
  x1 = fmaxf(x1, 0.0f) - fabsf(x1 * ex2_approx(p1));
}

// F.softplus(beta=1, threshold=20) = x if x > 20 else log1p(exp(x)), evaluated as
// max(x,0) + log1p(exp(-|x|)) with MUFU ex2/lg2 (series for tiny arguments): abs error ~1e-7.
__device__ __forceinline__ float softplus_t20(float x) {
  if (x > 20.0f) return x;
  const float e = __expf(-fabsf(x));
  const float l = e < 0.01f ? e * (1.0f - e * (0.5f - e * 0.33333334f)) : __logf(1.0f + e);
  return fmaxf(x, 0.0f) + l;
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// FakeQuantize of an output activation (quantize.py:86-133, asymmetric per-tensor uint8):
//   q = clamp(round_half_even(x / scale + zp), 0, 255);  dq = (q - zp) * scale;  result x + (dq - x)
// with the reference's own operation order (IEEE division, separately rounded add), so values land on the
// same grid point unless x itself sits within rounding noise of a boundary.  scale <= 0 -> pass-through.
__device__ __forceinline__ float fake_quant_u8(float x, float scale, float zp) {
  if (!(scale > 0.f)) return x;
  float q = rintf(__fadd_rn(__fdiv_rn(x, scale), zp));
  q = fminf(fmaxf(q, 0.f), 255.f);
  const float dq = __fmul_rn(__fsub_rn(q, zp), scale);
  return __fadd_rn(x, __fsub_rn(dq, x));
}

enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_SOFTPLUS = 2, ACT_SIGMOID = 3 };

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_GELU: return gelu_erf(v);
    case ACT_SOFTPLUS: return softplus_t20(v);
    case ACT_SIGMOID: return sigmoid_f(v);
    default: return v;
  }
}

// Tuning / debugging switches read from the environment exist only in a -DVASR_DEBUG build; the shipped library
// has ONE code path per operation (the defaults below are compiled in).  VASR_PDL (pure scheduling change, kept
// with a bit-identity test) and VASR_LIB (which build the Python binding loads) are the only run-time switches.
#ifdef VASR_DEBUG
inline int debug_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
#else
inline int debug_env_int(const char*, int dflt) { return dflt; }
#endif

// host: launch with the programmatic-serialisation attribute (VASR_PDL=0 turns it off)
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("VASR_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace vasr
