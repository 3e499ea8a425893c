// scan.cu — selective scan (SelectiveSSM._sequential_scan / _mamba_scan / _parallel_scan,
// velocity_asr/ssm.py:134-337), with the D skip (ssm.py:170,213) and the silu(z) gate
// (ssm.py:129) fused into the same pass.
//
//   h[d,n] <- exp(dt[t,d] A[n]) h[d,n] + x[t,d] dt[t,d] B[t,n];   y[t,d] = sum_n h[d,n] C[t,n] + x[t,d] D[d]
//
// Mapping.  A (b, d) row is sequential in t and is owned by LPR = N/8 lanes of one warp; each
// lane keeps 8 states in registers as four packed fp32x2 values, so the whole update runs on
// FFMA2/FMUL2 (two state updates per issue slot).  The partial dot products <h, C> of eight
// consecutive timesteps are reduced across the LPR lanes with one transpose-reduce (7 shuffles
// per 8 steps for N = 64 instead of 24), which leaves each lane holding the finished y of one
// (or two) timesteps: that lane applies the D skip and the gate.  All rows of a CTA belong to
// one utterance, so the B/C rows of a 32-step chunk are staged once in shared memory and read
// by every row as broadcast 16-byte loads; x/dt/z/y tiles go through shared memory too so
// global accesses are contiguous runs per timestep.
//
// Structured A.  The reference never re-initialises A_log (model.py:305-318), so A[n] = -(n+1)
// in every random-init model (ssm.py:83-84).  Then exp(dt A[n]) = r^(n+1) with r = exp(-dt):
// two MUFU ops per lane-step plus a packed multiply chain instead of one MUFU per state.
// The generic path (one ex2 per state) is kept for trained checkpoints.
//
// 'parallel' quirk.  The reference's default scan (ssm.py:228-295) is an exclusive scan with a
// mis-ordered down-sweep combine.  It equals (oracle/velocity_oracle.py
// scan_parallel_streaming), with H[e] the true state after e steps, S[e] = sum_{s<e} dt_s and
// parent(t) = t & (t-1):
//     hP[0] = 0;  hP[t] = hP[parent] + exp(A S[t]) * (H[t] - exp(A (S[t]-S[parent])) H[parent])
// which is evaluated left to right with O(log L) saved ancestors per state.
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"

namespace vasr {

namespace {

constexpr int SPL = 8;             // states per lane
constexpr int SCAN_THREADS = 128;
constexpr int NLEV = 26;           // ancestor levels of the quirk mode (L < 2^24)
constexpr float LOG2E = 1.4426950408889634f;

struct State8 {
  u64 v[4];
};

// p[k] = r^(n0 + k + 1), r = 2^s  (s = -dt * log2 e)
__device__ __forceinline__ void powers_structured(float s, float n0p1, State8& p) {
  const float r1 = ex2_approx(s);
  const float e1 = ex2_approx(s * n0p1);
  const float r2 = r1 * r1;
  const u64 rr = pack2(r2, r2);
  p.v[0] = pack2(e1, e1 * r1);
  p.v[1] = mul2(p.v[0], rr);
  p.v[2] = mul2(p.v[1], rr);
  p.v[3] = mul2(p.v[2], rr);
}
// p[k] = 2^(s * al2[k]),  al2[k] = A[n0+k] * log2 e  (negative), s = dt >= 0
__device__ __forceinline__ void powers_generic(float s, const float (&al2)[SPL], State8& p) {
#pragma unroll
  for (int k = 0; k < 4; ++k) p.v[k] = pack2(ex2_approx(s * al2[2 * k]), ex2_approx(s * al2[2 * k + 1]));
}

// ------------------------------------------------------------------------------------------
// The reference's 'parallel' scan (quirk).  One row per lane group (8 states per lane).  Steps are
// taken eight at a time, aligned to multiples of 8, so the ancestor an index needs is known at
// compile time for seven of the eight steps:  parent(8q+i) = 8q + (i & (i-1)).  Ancestor levels
// 1..3 ("most recent index with at least l trailing zeros") therefore live in registers; only the
// step at 8q touches the higher levels, which sit in local memory (one read, ~one write per 8 steps).
// ------------------------------------------------------------------------------------------
constexpr int TCQ = 16;

struct Anc {
  State8 hp, H;
  double S;
};

template <int LPR, bool STRUCT>
__global__ void __launch_bounds__(SCAN_THREADS) scan_quirk_kernel(ScanArgs a) {
  constexpr int N = LPR * SPL;
  constexpr int ROWS = SCAN_THREADS / LPR;
  constexpr int NR = (LPR == 8) ? 3 : (LPR == 4 ? 2 : (LPR == 2 ? 1 : 0));
  constexpr int TPL = 8 >> NR;
  constexpr int HL = NLEV - 4;   // levels >= 4

  __shared__ __align__(16) float sB[TCQ][N];
  __shared__ __align__(16) float sC[TCQ][N];
  __shared__ float sx[TCQ][ROWS];
  __shared__ float sdt[TCQ][ROWS];
  __shared__ float sz[TCQ][ROWS];
  __shared__ float sy[TCQ][ROWS];

  const int tid = threadIdx.x;
  const int rl = tid / LPR, j = tid % LPR;
  const int n0 = j * SPL;
  const int d0 = blockIdx.x * ROWS;
  const int64_t b = blockIdx.y;
  const int64_t L = a.L;
  const bool gate = a.z != nullptr;

  float al2[SPL];
#pragma unroll
  for (int k = 0; k < SPL; ++k) al2[k] = a.A[n0 + k] * LOG2E;
  const float n0p1 = (float)(n0 + 1);
  const float Dd = a.D ? a.D[d0 + rl] : 0.f;

  Anc cur, a1, a2, a3;
#pragma unroll
  for (int k = 0; k < 4; ++k) cur.hp.v[k] = cur.H.v[k] = 0ull;
  cur.S = 0.0;
  a1 = a2 = a3 = cur;
  float hi_hp[HL][SPL], hi_H[HL][SPL];
  double hi_S[HL];
  for (int l = 0; l < HL; ++l) {
    hi_S[l] = 0.0;
#pragma unroll
    for (int k = 0; k < SPL; ++k) hi_hp[l][k] = hi_H[l][k] = 0.f;
  }

  for (int64_t tc0 = 0; tc0 < L; tc0 += TCQ) {
    const int tcn = (int)((L - tc0) < TCQ ? (L - tc0) : TCQ);
    for (int idx = tid; idx < TCQ * (N / 4); idx += SCAN_THREADS) {
      const int t = idx / (N / 4), f = idx % (N / 4);
      float4 vb = make_float4(0.f, 0.f, 0.f, 0.f), vc = vb;
      if (t < tcn) {
        const int64_t row = b * L + tc0 + t;
        vb = __ldg(reinterpret_cast<const float4*>(a.Bm + row * a.ldb + 4 * f));
        vc = __ldg(reinterpret_cast<const float4*>(a.Cm + row * a.ldc + 4 * f));
      }
      const int slot = (f & 1) * LPR + (f >> 1);   // conflict-free 16-byte phases (see below)
      *reinterpret_cast<float4*>(&sB[t][4 * slot]) = vb;
      *reinterpret_cast<float4*>(&sC[t][4 * slot]) = vc;
    }
    for (int idx = tid; idx < TCQ * ROWS; idx += SCAN_THREADS) {
      const int t = idx / ROWS, r = idx % ROWS;
      float vx = 0.f, vd = 0.f, vz = 0.f;
      if (t < tcn) {
        const int64_t row = b * L + tc0 + t;
        vx = __ldg(a.x + row * a.ldx + d0 + r);
        vd = __ldg(a.dt + row * a.lddt + d0 + r);
        if (gate) vz = __ldg(a.z + row * a.ldz + d0 + r);
      }
      sx[t][r] = vx;
      sdt[t][r] = vd;
      sz[t][r] = vz;
    }
    __syncthreads();

#pragma unroll 1
    for (int g8 = 0; g8 < TCQ; g8 += 8) {
      if (g8 >= tcn) break;
      float yp[8];
      // one step: hP from `par`, output partial, then the true-recurrence update of cur.H / cur.S
      auto step = [&](int i, const Anc& par, bool has_parent) {
        const int t = g8 + i;
        const float dtv = sdt[t][rl];
        const float xv = sx[t][rl];
        const ulonglong2 c01 = *reinterpret_cast<const ulonglong2*>(&sC[t][4 * j]);
        const ulonglong2 c23 = *reinterpret_cast<const ulonglong2*>(&sC[t][4 * (LPR + j)]);
        if (has_parent) {
          const float Sf = (float)cur.S;
          const float dS = (float)(cur.S - par.S);
          State8 q, pd;
          if (STRUCT) {
            powers_structured(-Sf * LOG2E, n0p1, q);
            powers_structured(-dS * LOG2E, n0p1, pd);
          } else {
            powers_generic(Sf, al2, q);
            powers_generic(dS, al2, pd);
          }
          const u64 neg1 = pack2(-1.f, -1.f);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const u64 inner = fma2(mul2(pd.v[k], neg1), par.H.v[k], cur.H.v[k]);   // H - pd * H[parent]
            cur.hp.v[k] = fma2(q.v[k], inner, par.hp.v[k]);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) cur.hp.v[k] = 0ull;
        }
        u64 acc = fma2(cur.hp.v[1], c01.y, mul2(cur.hp.v[0], c01.x));
        u64 acc2 = fma2(cur.hp.v[3], c23.y, mul2(cur.hp.v[2], c23.x));
        yp[i] = hsum2(add2(acc, acc2));
      };
      auto advance = [&](int i) {
        const int t = g8 + i;
        const float dtv = sdt[t][rl];
        State8 p;
        if (STRUCT) powers_structured(-dtv * LOG2E, n0p1, p);
        else powers_generic(dtv, al2, p);
        const float u = sx[t][rl] * dtv;
        const u64 uu = pack2(u, u);
        const ulonglong2 b01 = *reinterpret_cast<const ulonglong2*>(&sB[t][4 * j]);
        const ulonglong2 b23 = *reinterpret_cast<const ulonglong2*>(&sB[t][4 * (LPR + j)]);
        cur.H.v[0] = fma2(p.v[0], cur.H.v[0], mul2(uu, b01.x));
        cur.H.v[1] = fma2(p.v[1], cur.H.v[1], mul2(uu, b01.y));
        cur.H.v[2] = fma2(p.v[2], cur.H.v[2], mul2(uu, b23.x));
        cur.H.v[3] = fma2(p.v[3], cur.H.v[3], mul2(uu, b23.y));
        cur.S += (double)dtv;
      };

      // ---- i = 0 : index 8q, ancestor level >= 4 (dynamic)
      {
        const int64_t tg = tc0 + g8;
        if (tg == 0) {
          step(0, cur, false);
        } else {
          const int z = 3 + (__ffsll((long long)(tg >> 3)) - 1);
          const int pl = z + 1 - 4;
          Anc par;
          par.S = hi_S[pl];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            par.hp.v[k] = pack2(hi_hp[pl][2 * k], hi_hp[pl][2 * k + 1]);
            par.H.v[k] = pack2(hi_H[pl][2 * k], hi_H[pl][2 * k + 1]);
          }
          step(0, par, true);
          for (int l = 4; l <= z; ++l) {
            hi_S[l - 4] = cur.S;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              unpack2(cur.hp.v[k], hi_hp[l - 4][2 * k], hi_hp[l - 4][2 * k + 1]);
              unpack2(cur.H.v[k], hi_H[l - 4][2 * k], hi_H[l - 4][2 * k + 1]);
            }
          }
          a1 = a2 = a3 = cur;
        }
        advance(0);
      }
      step(1, a1, true); advance(1);
      step(2, a2, true); a1 = cur; advance(2);
      step(3, a1, true); advance(3);
      step(4, a3, true); a1 = a2 = cur; advance(4);
      step(5, a1, true); advance(5);
      step(6, a2, true); a1 = cur; advance(6);
      step(7, a1, true); advance(7);

#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int lane_bit = LPR >> (r + 1);
        const int cnt = 4 >> r;
        const bool hi = (j & lane_bit) != 0;
#pragma unroll
        for (int i = 0; i < cnt; ++i) {
          const float mine = hi ? yp[i + cnt] : yp[i];
          const float other = hi ? yp[i] : yp[i + cnt];
          yp[i] = mine + __shfl_xor_sync(0xffffffffu, other, lane_bit);
        }
      }
#pragma unroll
      for (int i = 0; i < TPL; ++i) {
        const int t = g8 + j * TPL + i;
        float yv = yp[i] + sx[t][rl] * Dd;
        if (gate) {
          const float zv = sz[t][rl];
          yv *= zv / (1.0f + __expf(-zv));
        }
        sy[t][rl] = yv;
      }
    }
    __syncthreads();
    for (int idx = tid; idx < tcn * ROWS; idx += SCAN_THREADS) {
      const int t = idx / ROWS, r = idx % ROWS;
      a.y[(b * L + tc0 + t) * a.ldy + d0 + r] = sy[t][r];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Fast path of the true recurrence (scan_mode sequential / mamba).
//
// ncu on the first version showed the scan is bound by shared-memory wavefronts, not by the FMA
// pipe: a warp-wide 16-byte load runs in four 8-lane phases, a phase cannot broadcast across
// phases, and B/C traffic per state update falls only with the number of rows a lane serves.
// Here each lane keeps 8 states of TWO rows (16 state registers), so one B/C load feeds two
// updates, and the B/C rows are stored permuted ([first 16 B of every lane | second 16 B of every
// lane]) so that the 8 lanes of a phase read 128 contiguous bytes.  The staging threads also
// pre-compute s = -dt*log2(e) and u = x*dt once per (t, row) instead of once per lane.
// ------------------------------------------------------------------------------------------
constexpr int TC2 = 8;    // timesteps per staged chunk

template <int LPR, int WARPS, bool STRUCT>
__global__ void __launch_bounds__(WARPS * 32, 672 / (WARPS * 32)) scan_rows2_kernel(ScanArgs a) {
  constexpr int N = LPR * SPL;
  constexpr int GROUPS = 32 / LPR;             // lane groups per warp
  constexpr int ROWS = WARPS * GROUPS * 2;     // rows per CTA
  constexpr int THREADS = WARPS * 32;
  constexpr int NR = (LPR == 8) ? 3 : (LPR == 4 ? 2 : 1);
  constexpr int TPL = 8 >> NR;
  constexpr int NBC = TC2 * (N / 4);           // float4 per chunk of B (and of C)
  constexpr int NXD = TC2 * ROWS;              // elements per chunk of x (dt, z)
  constexpr int PBC = (NBC + THREADS - 1) / THREADS;
  constexpr int PXD = (NXD + THREADS - 1) / THREADS;

  // two buffers: chunk c+1 is committed into the other buffer while chunk c is being consumed,
  // so one __syncthreads per chunk suffices
  __shared__ __align__(16) float sB[2][TC2][N];
  __shared__ __align__(16) float sC[2][TC2][N];
  __shared__ __align__(16) float2 ssu[2][TC2][ROWS];  // (s, u)
  __shared__ float sx[2][TC2][ROWS];
  __shared__ float sz[2][TC2][ROWS];
  __shared__ float sy[2][TC2][ROWS];

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane / LPR, j = lane % LPR;
  const int rl0 = (warp * GROUPS + g) * 2;      // first of this lane's two rows (within the CTA)
  const int n0 = j * SPL;
  const int d0 = blockIdx.x * ROWS;
  const int64_t b = blockIdx.y;
  const int64_t L = a.L;
  const bool gate = a.z != nullptr;
  const int nchunks = (int)((L + TC2 - 1) / TC2);

  float al2[SPL];
#pragma unroll
  for (int k = 0; k < SPL; ++k) al2[k] = a.A[n0 + k] * LOG2E;
  const float n0p1 = (float)(n0 + 1);
  float Dv[2] = {0.f, 0.f};
  if (a.D) { Dv[0] = __ldg(a.D + d0 + rl0); Dv[1] = __ldg(a.D + d0 + rl0 + 1); }

  State8 H0, H1;
#pragma unroll
  for (int k = 0; k < 4; ++k) H0.v[k] = H1.v[k] = 0ull;

  // ---- software pipeline: the global loads of chunk c+1 are in flight while chunk c computes
  float4 rb[PBC], rc[PBC];
  float rx[PXD], rd[PXD], rz[PXD];
  auto prefetch = [&](int c) {
    const int64_t tc0 = (int64_t)c * TC2;
    const int tcn = (int)((L - tc0) < TC2 ? (L - tc0) : TC2);
#pragma unroll
    for (int i = 0; i < PBC; ++i) {
      const int idx = tid + i * THREADS;
      const int t = idx / (N / 4), f = idx % (N / 4);
      rb[i] = rc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < NBC && t < tcn) {
        const int64_t row = b * L + tc0 + t;
        rb[i] = __ldg(reinterpret_cast<const float4*>(a.Bm + row * a.ldb + 4 * f));
        rc[i] = __ldg(reinterpret_cast<const float4*>(a.Cm + row * a.ldc + 4 * f));
      }
    }
#pragma unroll
    for (int i = 0; i < PXD; ++i) {
      const int idx = tid + i * THREADS;
      const int t = idx / ROWS, r = idx % ROWS;
      rx[i] = rd[i] = rz[i] = 0.f;
      if (idx < NXD && t < tcn) {
        const int64_t row = b * L + tc0 + t;
        rx[i] = __ldg(a.x + row * a.ldx + d0 + r);
        rd[i] = __ldg(a.dt + row * a.lddt + d0 + r);
        if (gate) rz[i] = __ldg(a.z + row * a.ldz + d0 + r);
      }
    }
  };
  // B/C rows are stored permuted: float4 f = 2*jj + c of a row lands at slot c*LPR + jj, so the
  // 8 lanes of a 16-byte load phase read 128 contiguous bytes
  auto commit = [&](int buf) {
#pragma unroll
    for (int i = 0; i < PBC; ++i) {
      const int idx = tid + i * THREADS;
      if (idx < NBC) {
        const int t = idx / (N / 4), f = idx % (N / 4);
        const int slot = (f & 1) * LPR + (f >> 1);
        *reinterpret_cast<float4*>(&sB[buf][t][4 * slot]) = rb[i];
        *reinterpret_cast<float4*>(&sC[buf][t][4 * slot]) = rc[i];
      }
    }
#pragma unroll
    for (int i = 0; i < PXD; ++i) {
      const int idx = tid + i * THREADS;
      if (idx < NXD) {
        const int t = idx / ROWS, r = idx % ROWS;
        ssu[buf][t][r] = make_float2(STRUCT ? -rd[i] * LOG2E : rd[i], rx[i] * rd[i]);
        sx[buf][t][r] = rx[i];
        sz[buf][t][r] = rz[i];
      }
    }
  };

  prefetch(0);
  commit(0);
  __syncthreads();

  for (int c = 0; c < nchunks; ++c) {
    const int64_t tc0 = (int64_t)c * TC2;
    const int tcn = (int)((L - tc0) < TC2 ? (L - tc0) : TC2);
    const int buf = c & 1;
    if (c + 1 < nchunks) prefetch(c + 1);

#pragma unroll 1
    for (int g4 = 0; g4 < TC2; g4 += 4) {
      float yp[8];   // index 2*i + r : step i, row r
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = g4 + i;
        const float4 su = *reinterpret_cast<const float4*>(&ssu[buf][t][rl0]);   // (s0,u0,s1,u1)
        const ulonglong2 b01 = *reinterpret_cast<const ulonglong2*>(&sB[buf][t][4 * j]);
        const ulonglong2 b23 = *reinterpret_cast<const ulonglong2*>(&sB[buf][t][4 * (LPR + j)]);
        const ulonglong2 c01 = *reinterpret_cast<const ulonglong2*>(&sC[buf][t][4 * j]);
        const ulonglong2 c23 = *reinterpret_cast<const ulonglong2*>(&sC[buf][t][4 * (LPR + j)]);
        State8 p;
        u64 uu, acc, acc2;
        if (STRUCT) powers_structured(su.x, n0p1, p); else powers_generic(su.x, al2, p);
        uu = pack2(su.y, su.y);
        H0.v[0] = fma2(p.v[0], H0.v[0], mul2(uu, b01.x));
        H0.v[1] = fma2(p.v[1], H0.v[1], mul2(uu, b01.y));
        H0.v[2] = fma2(p.v[2], H0.v[2], mul2(uu, b23.x));
        H0.v[3] = fma2(p.v[3], H0.v[3], mul2(uu, b23.y));
        acc = fma2(H0.v[1], c01.y, mul2(H0.v[0], c01.x));
        acc2 = fma2(H0.v[3], c23.y, mul2(H0.v[2], c23.x));
        yp[2 * i] = hsum2(add2(acc, acc2));
        if (STRUCT) powers_structured(su.z, n0p1, p); else powers_generic(su.z, al2, p);
        uu = pack2(su.w, su.w);
        H1.v[0] = fma2(p.v[0], H1.v[0], mul2(uu, b01.x));
        H1.v[1] = fma2(p.v[1], H1.v[1], mul2(uu, b01.y));
        H1.v[2] = fma2(p.v[2], H1.v[2], mul2(uu, b23.x));
        H1.v[3] = fma2(p.v[3], H1.v[3], mul2(uu, b23.y));
        acc = fma2(H1.v[1], c01.y, mul2(H1.v[0], c01.x));
        acc2 = fma2(H1.v[3], c23.y, mul2(H1.v[2], c23.x));
        yp[2 * i + 1] = hsum2(add2(acc, acc2));
      }
      // transpose-reduce the 8 partials over the LPR lanes of the group
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int lane_bit = LPR >> (r + 1);
        const int cnt = 4 >> r;
        const bool hi = (j & lane_bit) != 0;
#pragma unroll
        for (int i = 0; i < cnt; ++i) {
          const float mine = hi ? yp[i + cnt] : yp[i];
          const float other = hi ? yp[i] : yp[i + cnt];
          yp[i] = mine + __shfl_xor_sync(0xffffffffu, other, lane_bit);
        }
      }
      // lane j owns values v = j*TPL + i: step v >> 1, row v & 1
#pragma unroll
      for (int i = 0; i < TPL; ++i) {
        const int v = j * TPL + i;
        const int t = g4 + (v >> 1), rr = v & 1;
        float yv = yp[i] + sx[buf][t][rl0 + rr] * (rr ? Dv[1] : Dv[0]);
        if (gate) {
          const float zv = sz[buf][t][rl0 + rr];
          yv *= zv / (1.0f + __expf(-zv));
        }
        sy[buf][t][rl0 + rr] = yv;
      }
    }
    if (c + 1 < nchunks) commit(buf ^ 1);   // other buffer: last read during chunk c-1, before the previous barrier
    __syncthreads();                        // sy[buf] complete, chunk c+1 visible
    for (int idx = tid; idx < tcn * ROWS; idx += THREADS) {
      const int t = idx / ROWS, r = idx % ROWS;
      a.y[(b * L + tc0 + t) * a.ldy + d0 + r] = sy[buf][t][r];   // sy[buf] is next written during chunk c+2
    }
  }
}

template <int LPR, int WARPS>
cudaError_t launch_rows2(const ScanArgs& a, cudaStream_t s) {
  constexpr int ROWS = WARPS * (32 / LPR) * 2;
  dim3 grid((unsigned)(a.Di / ROWS), (unsigned)a.B);
  if (a.structured_a) scan_rows2_kernel<LPR, WARPS, true><<<grid, WARPS * 32, 0, s>>>(a);
  else scan_rows2_kernel<LPR, WARPS, false><<<grid, WARPS * 32, 0, s>>>(a);
  return cudaGetLastError();
}

// picks the CTA width: 3 warps (24 rows at N = 64) tiles 384 channels x 64 utterances onto
// 148 SMs almost evenly (1024 CTAs, 6.9 per SM); other widths are fallbacks for other Di.
template <int LPR>
cudaError_t launch_recurrence(const ScanArgs& a, cudaStream_t s) {
  constexpr int RPW = (32 / LPR) * 2;
  if (a.Di % (3 * RPW) == 0) return launch_rows2<LPR, 3>(a, s);
  if (a.Di % (4 * RPW) == 0) return launch_rows2<LPR, 4>(a, s);
  if (a.Di % (2 * RPW) == 0) return launch_rows2<LPR, 2>(a, s);
  if (a.Di % RPW == 0) return launch_rows2<LPR, 1>(a, s);
  return cudaErrorInvalidValue;
}

template <int LPR>
cudaError_t launch_lpr(const ScanArgs& a, cudaStream_t s) {
  constexpr int ROWS = SCAN_THREADS / LPR;
  if (a.Di % ROWS != 0) return cudaErrorInvalidValue;
  dim3 grid((unsigned)(a.Di / ROWS), (unsigned)a.B);
  if (!a.parallel_quirk) return launch_recurrence<LPR>(a, s);
  if (a.structured_a) scan_quirk_kernel<LPR, true><<<grid, SCAN_THREADS, 0, s>>>(a);
  else scan_quirk_kernel<LPR, false><<<grid, SCAN_THREADS, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_selective_scan(const ScanArgs& a, cudaStream_t s, int64_t* launches) {
  if (a.B <= 0 || a.L <= 0) return cudaSuccess;
  if (a.B > 65535 || a.L >= (1LL << (NLEV - 2))) return cudaErrorInvalidValue;
  if ((a.ldb & 3) || (a.ldc & 3) || (reinterpret_cast<uintptr_t>(a.Bm) & 15) ||
      (reinterpret_cast<uintptr_t>(a.Cm) & 15))
    return cudaErrorInvalidValue;
  cudaError_t e;
  switch (a.N) {
    case 64: e = launch_lpr<8>(a, s); break;
    case 32: e = launch_lpr<4>(a, s); break;
    case 16: e = launch_lpr<2>(a, s); break;
    default: return cudaErrorInvalidValue;
  }
  if (launches && e == cudaSuccess) ++*launches;
  return e;
}

}  // namespace vasr
