// beam.cu — CTC prefix beam search on the device (ctc_beam_search, velocity_asr/decode.py:128-217, the
// lm_scorer=None path).  The reference walks every (beam, token) pair of every frame in Python; its rule is
//   new_beams[key] = max over candidates of (score + log_prob), first-inserted wins a tie,
//   key = prefix (blank, or the beam's own last token again) | prefix + (token,)
//   then the `beam_width` best keys, stable in insertion order.
// Two facts make that cheap without changing a single result:
//  * inside one beam the extension candidates are ranked by the frame's log-probs alone, so a candidate that
//    makes the global top-W is among the frame's W+1 best non-blank tokens (one of them may be the beam's
//    last token, which does not extend).  beam_rows_kernel finds those per frame, all frames in parallel;
//  * two candidates share a key only when an extension recreates a prefix that is already a beam (the child
//    of the extending beam) — found through a per-utterance trie of prefixes (parent, token, depth).
// beam_search_kernel then runs the frames in order, one warp per utterance, one beam per lane, scores in
// fp64 (the reference adds Python floats), ties broken by the reference's insertion order
// (beam rank, then blank, then token id).
#include <float.h>

#include "common.cuh"
#include "kernels.cuh"

namespace vasr {

namespace {

// One warp per (utterance, frame) row: max, log-sum-exp, and the K best non-blank tokens by
// log-prob (ties -> lower token id).  log-prob = (x - max) - log(sum exp(x - max)), torch's order.
__global__ void __launch_bounds__(256) beam_rows_kernel(const float* __restrict__ logits, float2* __restrict__ stats,
                                                        int32_t* __restrict__ top_tok, int64_t M, int V, int K,
                                                        int blank) {
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  const float* x = logits + m * V;
  float mx = -INFINITY;
  for (int n = lane; n < V; n += 32) mx = fmaxf(mx, x[n]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int n = lane; n < V; n += 32) sum += expf(x[n] - mx);
  sum = warp_sum(sum);
  const float ls = logf(sum);
  if (lane == 0) stats[m] = make_float2(mx, ls);
  // K passes; pass k takes the best entry strictly after the previous winner in (lp desc, token asc) order
  float pv = INFINITY;
  int pi = -1;
  for (int k = 0; k < K; ++k) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int n = lane; n < V; n += 32) {
      if (n == blank) continue;
      const float v = (x[n] - mx) - ls;
      const bool after = v < pv || (v == pv && n > pi);
      if (after && (bi == 0x7fffffff || v > bv)) {  // ascending n inside a lane: first seen wins ties
        bv = v;
        bi = n;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) top_tok[m * K + k] = bi == 0x7fffffff ? -1 : bi;
    if (bi == 0x7fffffff) {
      for (int kk = k + 1; kk < K; ++kk)
        if (lane == 0) top_tok[m * K + kk] = -1;
      break;
    }
    pv = bv;
    pi = bi;
  }
}

constexpr long long kNoOrder = 0x7fffffffffffffffLL;

__global__ void __launch_bounds__(32) beam_search_kernel(
    const float* __restrict__ logits, const float2* __restrict__ stats, const int32_t* __restrict__ top_tok, int K,
    int64_t L, int V, int W, int blank, int32_t* trie_parent, int32_t* trie_token, int32_t* trie_depth,
    int64_t trie_cap, int32_t* __restrict__ out_tokens, int32_t* __restrict__ out_lens,
    double* __restrict__ out_scores) {
  const int64_t b = blockIdx.x;
  const int lane = threadIdx.x;
  int32_t* parent = trie_parent + b * trie_cap;
  int32_t* token = trie_token + b * trie_cap;
  int32_t* depth = trie_depth + b * trie_cap;

  __shared__ int s_node[32], s_last[32], s_pl[32], s_tk[32], n_node[32], n_last[32];
  __shared__ double s_score[32], n_score[32];

  if (lane == 0) {
    parent[0] = -1;
    token[0] = -1;
    depth[0] = 0;
  }
  int nb = 1, n_nodes = 1;
  int node = 0, last = -1;  // last = -1 is the reference's None
  double score = 0.0;
  const long long Vp = (long long)V + 1;
  __syncwarp();

  for (int64_t t = 0; t < L; ++t) {
    const int64_t row = b * L + t;
    const float* x = logits + row * V;
    const float2 st = stats[row];
    const int32_t* tt = top_tok + row * K;
    auto lp = [&](int tok) -> double { return (double)((x[tok] - st.x) - st.y); };
    const bool valid = lane < nb;

    s_node[lane] = valid ? node : -2;
    const int par = valid ? parent[node] : -3;
    const int tk = valid ? token[node] : -1;
    __syncwarp();
    int pl = -1;
    if (valid && par >= 0)
      for (int j = 0; j < nb; ++j)
        if (s_node[j] == par) pl = j;
    s_pl[lane] = pl;
    s_tk[lane] = tk;
    s_last[lane] = last;
    s_score[lane] = score;
    __syncwarp();

    // the key this beam already owns: its own blank / repeat candidates and, when its parent prefix is a
    // beam too, the parent's extension by this beam's last prefix token
    double e_score = -INFINITY;
    int e_last = blank;
    long long e_order = kNoOrder;
    bool e_avail = valid;
    if (valid) {
      double cs[3];
      int cl[3];
      long long co[3];
      bool ch[3];
      const bool h1 = pl >= 0 && s_last[pl] != tk;
      const double sc1 = h1 ? s_score[pl] + lp(tk) : 0.0;
      const long long o1 = (long long)pl * Vp + 1 + tk;
      const bool h3 = last >= 0 && last != blank;
      const double sc3 = h3 ? score + lp(last) : 0.0;
      const int first = (h1 && pl < lane) ? 0 : 2;  // where the parent's candidate sits in insertion order
      const int i2 = first == 0 ? 1 : 0, i3 = i2 + 1;
      cs[first] = sc1, cl[first] = tk, co[first] = o1, ch[first] = h1;
      cs[i2] = score + lp(blank), cl[i2] = blank, co[i2] = (long long)lane * Vp, ch[i2] = true;
      cs[i3] = sc3, cl[i3] = last, co[i3] = (long long)lane * Vp + 1 + last, ch[i3] = h3;
      bool none = true;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (!ch[i]) continue;
        if (none || cs[i] > e_score) {
          e_score = cs[i];
          e_last = cl[i];
        }
        if (none) e_order = co[i];
        none = false;
      }
    }

    // this beam's extensions to prefixes that are not beams yet, best first
    int p = valid ? 0 : K;
    int x_tok = -1;
    double x_score = -INFINITY;
    auto advance = [&]() {
      while (p < K) {
        const int tok = tt[p];
        if (tok < 0) {
          p = K;
          break;
        }
        bool skip = tok == last;
        for (int j = 0; j < nb && !skip; ++j) skip = s_pl[j] == lane && s_tk[j] == tok;
        if (!skip) {
          x_tok = tok;
          x_score = score + lp(tok);
          return;
        }
        ++p;
      }
    };
    advance();

    int count = 0;
    for (int r = 0; r < W; ++r) {
      const bool hx = p < K;
      const long long x_order = (long long)lane * Vp + 1 + x_tok;
      bool take_e = e_avail;
      if (e_avail && hx) take_e = e_score > x_score || (e_score == x_score && e_order < x_order);
      double ps = take_e ? e_score : (hx ? x_score : -INFINITY);
      long long po = take_e ? e_order : (hx ? x_order : kNoOrder);
      double bs = ps;
      long long bo = po;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, bs, o);
        const long long oo = __shfl_xor_sync(0xffffffffu, bo, o);
        if (oo != kNoOrder && (bo == kNoOrder || os > bs || (os == bs && oo < bo))) {
          bs = os;
          bo = oo;
        }
      }
      if (bo == kNoOrder) break;
      const bool mine = po == bo;
      const bool is_new = mine && !take_e;
      const unsigned newm = __ballot_sync(0xffffffffu, is_new);
      if (mine) {
        if (take_e) {
          n_node[r] = node;
          n_score[r] = e_score;
          n_last[r] = e_last;
          e_avail = false;
        } else {
          parent[n_nodes] = node;
          token[n_nodes] = x_tok;
          depth[n_nodes] = depth[node] + 1;
          n_node[r] = n_nodes;
          n_score[r] = x_score;
          n_last[r] = x_tok;
          ++p;
          advance();
        }
      }
      n_nodes += newm != 0;
      ++count;
    }
    __syncwarp();
    nb = count;
    if (lane < nb) {
      node = n_node[lane];
      score = n_score[lane];
      last = n_last[lane];
    }
    __syncwarp();
  }

  if (lane < W) {
    int32_t* out = out_tokens + (b * W + lane) * L;
    if (lane < nb) {
      const int len = depth[node];
      out_scores[b * W + lane] = score;
      out_lens[b * W + lane] = len;
      int nd = node;
      for (int i = len - 1; i >= 0; --i) {
        out[i] = token[nd];
        nd = parent[nd];
      }
      for (int64_t i = len; i < L; ++i) out[i] = -1;
    } else {
      out_scores[b * W + lane] = -INFINITY;
      out_lens[b * W + lane] = -1;
      for (int64_t i = 0; i < L; ++i) out[i] = -1;
    }
  }
}

}  // namespace

cudaError_t launch_beam_rows(const float* logits, float2* stats, int32_t* top_tok, int64_t M, int V, int K,
                             int blank, cudaStream_t s, int64_t* launches) {
  if (M <= 0) return cudaSuccess;
  beam_rows_kernel<<<(unsigned)((M + 7) / 8), 256, 0, s>>>(logits, stats, top_tok, M, V, K, blank);
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_beam_search(const float* logits, const float2* stats, const int32_t* top_tok, int K, int64_t B,
                               int64_t L, int V, int W, int blank, int32_t* trie, int64_t trie_cap,
                               int32_t* out_tokens, int32_t* out_lens, double* out_scores, cudaStream_t s,
                               int64_t* launches) {
  if (B <= 0) return cudaSuccess;
  beam_search_kernel<<<(unsigned)B, 32, 0, s>>>(logits, stats, top_tok, K, L, V, W, blank, trie,
                                                trie + B * trie_cap, trie + 2 * B * trie_cap, trie_cap, out_tokens,
                                                out_lens, out_scores);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace vasr
