// mel.cu — log-mel front end (compute_mel_spectrogram, velocity_asr/audio.py:65-143).
//
// Stage order of the reference: reflect-pad 200 (audio.py:100-101) -> frames of 400 every 160
// with a periodic Hann window -> one-sided 400-point DFT (torch.stft, audio.py:104-112) ->
// power (:115) -> 80 x 201 HTK-mel filterbank (:118-126) -> log(. + 1e-10) (:129) ->
// per (utterance, bin) (x - mean_T) / (std_T unbiased + 1e-10) (:132-135) -> (B, T, 80).
//
// Round-1 structure: the windowed DFT is one strided fp32 projection (gemm.cu) over the
// padded PCM — frame t of utterance b is the 400 floats at xp + b*ldp + 160 t, so the frame
// matrix is never materialised — against a (402 x 400) table [hann*cos | hann*sin]; this file
// holds the streaming kernels on either side of it.  A 400-point DFT, not a zero-padded
// 512-point one: the bin spacing has to be 40 Hz to match the reference.
#include "common.cuh"
#include "kernels.cuh"

namespace vasr {

namespace {

__global__ void __launch_bounds__(256) reflect_pad_kernel(const float* __restrict__ pcm, float* __restrict__ xp,
                                                          int64_t S, int pad, int64_t ldp) {
  const int64_t b = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= ldp) return;
  float v = 0.f;
  if (i < S + 2 * pad) {
    int64_t src = i - pad;
    if (src < 0) src = -src;                       // reflect without repeating the edge sample
    if (src >= S) src = 2 * (S - 1) - src;
    v = pcm[b * S + src];
  }
  xp[b * ldp + i] = v;
}

// 8 frames per CTA; power spectrum in shared memory, then one thread per (frame, mel bin).
constexpr int MF = 8;
__global__ void __launch_bounds__(256) mel_log_kernel(const float* __restrict__ spec, int64_t lds,
                                                      float* __restrict__ raw, int64_t M, int nf, int n_mels,
                                                      const int* __restrict__ fb_lo,
                                                      const int* __restrict__ fb_off,
                                                      const float* __restrict__ fb_w) {
  extern __shared__ float pw[];  // MF x nf
  const int64_t m0 = (int64_t)blockIdx.x * MF;
  for (int idx = threadIdx.x; idx < MF * nf; idx += 256) {
    const int f = idx / nf, k = idx - f * nf;
    const int64_t m = m0 + f;
    float p = 0.f;
    if (m < M) {
      const float re = spec[m * lds + k], im = spec[m * lds + nf + k];
      p = re * re + im * im;
    }
    pw[idx] = p;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < MF * n_mels; idx += 256) {
    const int f = idx / n_mels, j = idx - f * n_mels;
    const int64_t m = m0 + f;
    if (m >= M) break;
    const int lo = fb_lo[j], o0 = fb_off[j], o1 = fb_off[j + 1];
    float acc = 0.f;
    for (int o = o0; o < o1; ++o) acc = fmaf(fb_w[o], pw[f * nf + lo + (o - o0)], acc);
    raw[m * n_mels + j] = logf(acc + 1e-10f);
  }
}

// One CTA per utterance; thread (slice, bin) strides over frames; two passes (mean, then
// squared deviations about it) with double accumulation so T = 60,001 frames loses nothing.
constexpr int STAT_SLICES = 8;
__global__ void __launch_bounds__(1024) mel_stats_kernel(const float* __restrict__ raw, float* __restrict__ mean,
                                                         float* __restrict__ rstd, int64_t T, int n_mels) {
  extern __shared__ double red[];  // STAT_SLICES x n_mels
  const int64_t b = blockIdx.x;
  const int j = threadIdx.x % n_mels, sl = threadIdx.x / n_mels;
  const bool live = sl < STAT_SLICES;
  const float* base = raw + b * T * n_mels;
  double s = 0.0;
  if (live)
    for (int64_t t = sl; t < T; t += STAT_SLICES) s += (double)base[t * n_mels + j];
  if (live) red[sl * n_mels + j] = s;
  __syncthreads();
  double mu = 0.0;
  for (int q = 0; q < STAT_SLICES; ++q) mu += red[q * n_mels + j];
  mu /= (double)T;
  __syncthreads();
  double v = 0.0;
  if (live)
    for (int64_t t = sl; t < T; t += STAT_SLICES) {
      const double d = (double)base[t * n_mels + j] - mu;
      v += d * d;
    }
  if (live) red[sl * n_mels + j] = v;
  __syncthreads();
  if (sl == 0) {
    double var = 0.0;
    for (int q = 0; q < STAT_SLICES; ++q) var += red[q * n_mels + j];
    // torch.std(unbiased) of a single frame is NaN; keep that behaviour (0/0).
    const double sd = sqrt(var / (double)(T - 1));
    mean[b * n_mels + j] = (float)mu;
    rstd[b * n_mels + j] = (float)(1.0 / (sd + 1e-10));
  }
}

__global__ void __launch_bounds__(256) mel_finish_kernel(const float* __restrict__ raw,
                                                         const float* __restrict__ mean,
                                                         const float* __restrict__ rstd, float* __restrict__ out,
                                                         int64_t T, int n_mels, int64_t fpu, int front) {
  const int64_t b = blockIdx.y;
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= fpu * n_mels) return;
  const int64_t p = idx / n_mels;
  const int j = (int)(idx - p * n_mels);
  const int64_t t = p - front;
  float v = 0.f;
  if (t >= 0 && t < T) {
    v = raw[(b * T + t) * n_mels + j];
    if (mean) v = (v - mean[b * n_mels + j]) * rstd[b * n_mels + j];
  }
  out[b * fpu * n_mels + idx] = v;
}

}  // namespace

cudaError_t launch_reflect_pad(const float* pcm, float* xp, int64_t B, int64_t S, int pad, int64_t ldp,
                               cudaStream_t s, int64_t* launches) {
  if (B <= 0) return cudaSuccess;
  if (S <= pad || B > 65535) return cudaErrorInvalidValue;
  dim3 grid((unsigned)((ldp + 255) / 256), (unsigned)B);
  reflect_pad_kernel<<<grid, 256, 0, s>>>(pcm, xp, S, pad, ldp);
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_mel_log(const float* spec, int64_t lds, float* raw, int64_t M, int nf, int n_mels,
                           const int* fb_lo, const int* fb_off, const float* fb_w, cudaStream_t s,
                           int64_t* launches) {
  if (M <= 0) return cudaSuccess;
  mel_log_kernel<<<(unsigned)((M + MF - 1) / MF), 256, (size_t)MF * nf * sizeof(float), s>>>(
      spec, lds, raw, M, nf, n_mels, fb_lo, fb_off, fb_w);
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_mel_stats(const float* raw, float* mean, float* rstd, int64_t B, int64_t T, int n_mels,
                             cudaStream_t s, int64_t* launches) {
  if (B <= 0) return cudaSuccess;
  if (n_mels * STAT_SLICES > 1024) return cudaErrorInvalidValue;
  const int threads = ((n_mels * STAT_SLICES + 31) / 32) * 32;
  mel_stats_kernel<<<(unsigned)B, threads, (size_t)STAT_SLICES * n_mels * sizeof(double), s>>>(raw, mean, rstd, T,
                                                                                               n_mels);
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_mel_finish(const float* raw, const float* mean, const float* rstd, float* out, int64_t B,
                              int64_t T, int n_mels, int64_t frames_per_utt, int front, cudaStream_t s,
                              int64_t* launches) {
  if (B <= 0) return cudaSuccess;
  if (B > 65535) return cudaErrorInvalidValue;
  dim3 grid((unsigned)((frames_per_utt * n_mels + 255) / 256), (unsigned)B);
  mel_finish_kernel<<<grid, 256, 0, s>>>(raw, mean, rstd, out, T, n_mels, frames_per_utt, front);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace vasr
