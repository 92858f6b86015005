// STFT / fused mask + iSTFT kernels (HBM-bound stage of the hot path).
//
// Replaces torch.stft / torch.istft at models/bs_roformer/bs_roformer.py:485,575,
// models/bs_roformer/mel_band_roformer.py:516,622 and models/mdx23c_tfc_tdf_v3.py:19-26,42.
// Stereo pairs are transformed with ONE complex FFT per frame (z = xL + i xR), so a frame costs one
// N-point transform instead of two real ones.
#include "common.cuh"
#include "fft.cuh"
#include "sesa_b200.h"

// n_fft = 2048 fast paths (stft2048.cu)
int sesa_launch_stft2048(const float* audio, float* spec, const float* window, const float* twiddle, int n_signals,
                         int channels, int64_t length, int hop, int T, cudaStream_t stream);
int sesa_launch_mask_istft2048(const float* spec, const float* mask, const int* inv, const float* cnt, float* out,
                               const float* window, const float* env, const float* twiddle, int batch, int nstems,
                               int channels, int hop, int T, int64_t out_len, int mode, int n_gathered,
                               cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// STFT.  grid = (T, NS) ; one CTA per frame of one signal group (all C<=2 channels of one chunk).
// layout 0 (RoFormer): spec[(ns*T + t)][f][c][re/im]      ('b t (f s c)', bs_roformer.py:497)
// layout 1 (MDX23C)  : spec[ns][c][re/im][f < dim_f][t]    (mdx23c_tfc_tdf_v3.py:27-30)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stft_kernel(const float* __restrict__ audio, float* __restrict__ spec,
                                                   const float* __restrict__ window,
                                                   const float2* __restrict__ tw, int C, int64_t L, int N,
                                                   int hop, int T, int layout, int dim_f) {
  extern __shared__ float2 smem_fft[];
  float2* b0 = smem_fft;
  float2* b1 = smem_fft + N;
  const int t = blockIdx.x;
  const int ns = blockIdx.y;
  const float* a0 = audio + (int64_t)ns * C * L;
  const float* a1 = a0 + L;
  const int64_t base = (int64_t)t * hop - N / 2;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int64_t j = reflect_index(base + i, L);
    const float w = window[i];
    float2 z;
    z.x = a0[j] * w;
    z.y = (C == 2) ? a1[j] * w : 0.f;
    b0[i] = z;
  }
  __syncthreads();
  const float2* Z = block_fft<false>(b0, b1, N, tw);
  const int F = N / 2 + 1;
  if (layout == 0) {
    if (C == 2) {
      float4* out = reinterpret_cast<float4*>(spec + ((int64_t)ns * T + t) * (int64_t)F * 4);
      for (int k = threadIdx.x; k < F; k += blockDim.x) {
        const float2 zk = Z[k];
        const float2 zn = Z[(N - k) & (N - 1)];
        // XL = (Z[k] + conj(Z[N-k]))/2 ; XR = (Z[k] - conj(Z[N-k]))/(2i)
        float4 o;
        o.x = 0.5f * (zk.x + zn.x);
        o.y = 0.5f * (zk.y - zn.y);
        o.z = 0.5f * (zk.y + zn.y);
        o.w = 0.5f * (zn.x - zk.x);
        out[k] = o;
      }
    } else {
      float2* out = reinterpret_cast<float2*>(spec + ((int64_t)ns * T + t) * (int64_t)F * 2);
      for (int k = threadIdx.x; k < F; k += blockDim.x) out[k] = Z[k];
    }
  } else {
    // MDX23C: time-contiguous planes; one frame writes a strided column (coalescing is recovered by
    // neighbouring CTAs writing neighbouring t).
    float* o = spec + (int64_t)ns * C * 2 * dim_f * T + t;
    for (int k = threadIdx.x; k < dim_f; k += blockDim.x) {
      const float2 zk = Z[k];
      if (C == 2) {
        const float2 zn = Z[(N - k) & (N - 1)];
        o[((int64_t)0 * dim_f + k) * T] = 0.5f * (zk.x + zn.x);
        o[((int64_t)1 * dim_f + k) * T] = 0.5f * (zk.y - zn.y);
        o[((int64_t)2 * dim_f + k) * T] = 0.5f * (zk.y + zn.y);
        o[((int64_t)3 * dim_f + k) * T] = 0.5f * (zn.x - zk.x);
      } else {
        o[((int64_t)0 * dim_f + k) * T] = zk.x;
        o[((int64_t)1 * dim_f + k) * T] = zk.y;
      }
    }
  }
}

extern "C" int sesa_stft(const float* audio, float* spec, const float* window, const float* twiddle,
                         int n_signals, int channels, int64_t length, int n_fft, int hop, int layout,
                         int dim_f, void* stream) {
  SESA_CHECK_ARG(channels == 1 || channels == 2, "sesa_stft: channels must be 1 or 2, got %d", channels);
  SESA_CHECK_ARG(n_fft >= 64 && n_fft <= 8192 && (n_fft & (n_fft - 1)) == 0,
                 "sesa_stft: n_fft must be a power of two in [64, 8192], got %d", n_fft);
  SESA_CHECK_ARG(length > n_fft / 2, "sesa_stft: signal (%lld) shorter than reflect pad (%d)",
                 (long long)length, n_fft / 2);
  SESA_CHECK_ARG(layout == 0 || layout == 1, "sesa_stft: bad layout %d", layout);
  if (n_signals == 0) return SESA_OK;
  const int T = 1 + (int)(length / hop);
  if (n_fft == 2048 && layout == 0 && (channels == 1 || (reinterpret_cast<uintptr_t>(spec) & 15) == 0))
    return sesa_launch_stft2048(audio, spec, window, twiddle, n_signals, channels, length, hop, T, (cudaStream_t)stream);
  const size_t smem = (size_t)2 * n_fft * sizeof(float2);
  SESA_CUDA(cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(T, n_signals);
  stft_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(audio, spec, window,
                                                         reinterpret_cast<const float2*>(twiddle), channels,
                                                         length, n_fft, hop, T, layout,
                                                         layout == 1 ? dim_f : n_fft / 2 + 1);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

// ---------------------------------------------------------------------------------------------
// Fused complex-mask multiply + iSTFT (+ window, overlap-add over frames, /window-envelope, trim).
// One CTA produces `seg` consecutive output samples of one (chunk, stem): it inverse-transforms
// every frame overlapping the segment and accumulates in shared memory, so there are no global
// atomics and the result is deterministic.
//
// spec  : layout 0 -> [b][t][f][c][2]                       (stereo: float4 per bin)
// mask  : mode 0 (BS)  [n][b*T+t][f][c][2]  same feature order as spec
//         mode 1 (Mel) [n][b*T+t][J][2]     J gathered (f,s) rows; inv[(f*C+c)*2 + {0,1}] gives the (up
//                      to two) source rows j of bin (f,c) (-1 = none), cnt[f*C+c] the divisor
//         mode 2 (none) spec already holds the stem's spectrum: spec[(b*NST+n)][t][f][c][2]
// out   : [b][n][c][out_len]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mask_istft_kernel(
    const float* __restrict__ spec, const float* __restrict__ mask, const int* __restrict__ inv,
    const float* __restrict__ cnt, float* __restrict__ out, const float* __restrict__ window,
    const float* __restrict__ env, const float2* __restrict__ tw, int C, int N, int hop, int T,
    int64_t out_len, int seg, int nstems, int mode, int J) {
  extern __shared__ float2 smem_fft[];
  float2* b0 = smem_fft;
  float2* b1 = smem_fft + N;
  float* acc = reinterpret_cast<float*>(smem_fft + 2 * N);  // [C][seg]
  const int g = blockIdx.x;
  const int n = blockIdx.y;
  const int b = blockIdx.z;
  const int F = N / 2 + 1;
  const int64_t i0 = (int64_t)g * seg;                       // first output sample of this CTA
  const int64_t q0 = i0 + N / 2;                             // in centre-padded coordinates
  const int64_t q1 = min(q0 + seg, (int64_t)N / 2 + out_len);
  for (int i = threadIdx.x; i < C * seg; i += blockDim.x) acc[i] = 0.f;
  int64_t t_lo = (q0 - N) / hop + 1;
  if (q0 - N < 0) t_lo = 0;
  int64_t t_hi = (q1 - 1) / hop;
  if (t_hi > T - 1) t_hi = T - 1;
  const float scale = 1.0f / (float)N;
  for (int64_t t = t_lo; t <= t_hi; ++t) {
    __syncthreads();
    const int64_t row = (int64_t)b * T + t;
    // build the Hermitian-extended spectrum of z = yL + i yR
    for (int k = threadIdx.x; k < F; k += blockDim.x) {
      float2 yl, yr = make_float2(0.f, 0.f);
      if (C == 2) {
        float4 sp;
        if (mode == 2) sp = reinterpret_cast<const float4*>(spec)[(((int64_t)b * nstems + n) * T + t) * F + k];
        else sp = reinterpret_cast<const float4*>(spec)[row * F + k];
        if (mode == 0) {
          const float4 m = reinterpret_cast<const float4*>(mask)[((int64_t)n * gridDim.z * T + row) * F + k];
          yl = cmul(make_float2(sp.x, sp.y), make_float2(m.x, m.y));
          yr = cmul(make_float2(sp.z, sp.w), make_float2(m.z, m.w));
        } else if (mode == 1) {
          const float2* mrow = reinterpret_cast<const float2*>(mask) + ((int64_t)n * gridDim.z * T + row) * J;
          float2 ml = make_float2(0.f, 0.f), mr = make_float2(0.f, 0.f);
          const int* iv = inv + (k * 2) * 2;
          if (iv[0] >= 0) ml = cadd(ml, mrow[iv[0]]);
          if (iv[1] >= 0) ml = cadd(ml, mrow[iv[1]]);
          if (iv[2] >= 0) mr = cadd(mr, mrow[iv[2]]);
          if (iv[3] >= 0) mr = cadd(mr, mrow[iv[3]]);
          const float cl = cnt[k * 2], cr = cnt[k * 2 + 1];
          ml.x /= cl; ml.y /= cl; mr.x /= cr; mr.y /= cr;
          yl = cmul(make_float2(sp.x, sp.y), ml);
          yr = cmul(make_float2(sp.z, sp.w), mr);
        } else {
          yl = make_float2(sp.x, sp.y);
          yr = make_float2(sp.z, sp.w);
        }
      } else {
        float2 sp;
        if (mode == 2) sp = reinterpret_cast<const float2*>(spec)[(((int64_t)b * nstems + n) * T + t) * F + k];
        else sp = reinterpret_cast<const float2*>(spec)[row * F + k];
        if (mode == 0) {
          yl = cmul(sp, reinterpret_cast<const float2*>(mask)[((int64_t)n * gridDim.z * T + row) * F + k]);
        } else if (mode == 1) {
          const float2* mrow = reinterpret_cast<const float2*>(mask) + ((int64_t)n * gridDim.z * T + row) * J;
          float2 ml = make_float2(0.f, 0.f);
          const int* iv = inv + k * 2;
          if (iv[0] >= 0) ml = cadd(ml, mrow[iv[0]]);
          if (iv[1] >= 0) ml = cadd(ml, mrow[iv[1]]);
          const float cl = cnt[k];
          ml.x /= cl; ml.y /= cl;
          yl = cmul(sp, ml);
        } else {
          yl = sp;
        }
      }
      if (k == 0 || k == N / 2) {  // c2r ignores the imaginary part of DC and Nyquist
        yl.y = 0.f; yr.y = 0.f;
      }
      // Z[k] = YL + i YR ; Z[N-k] = conj(YL) + i conj(YR)
      b0[k] = make_float2(yl.x - yr.y, yl.y + yr.x);
      if (k != 0 && k != N / 2) b0[N - k] = make_float2(yl.x + yr.y, yr.x - yl.y);
    }
    __syncthreads();
    const float2* z = block_fft<true>(b0, b1, N, tw);
    // windowed overlap-add of the part of this frame that falls inside the segment
    const int64_t fq = t * hop;  // frame start in padded coordinates
    int lo = (int)max((int64_t)0, q0 - fq);
    int hi = (int)min((int64_t)N, q1 - fq);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const float2 v = z[i];
      const float w = window[i] * scale;
      const int o = (int)(fq + i - q0);
      acc[o] += v.x * w;
      if (C == 2) acc[seg + o] += v.y * w;
    }
  }
  __syncthreads();
  const int nvalid = (int)(q1 - q0);
  for (int i = threadIdx.x; i < nvalid; i += blockDim.x) {
    const float e = env[i0 + i];
    for (int c = 0; c < C; ++c)
      out[(((int64_t)b * nstems + n) * C + c) * out_len + i0 + i] = acc[c * seg + i] / e;
  }
}

extern "C" int sesa_mask_istft(const float* spec, const float* mask, const int* inv_index,
                               const float* inv_count, float* out, const float* window,
                               const float* envelope, const float* twiddle, int batch, int nstems,
                               int channels, int n_fft, int hop, int n_frames, int64_t out_len, int mode,
                               int n_gathered, void* stream) {
  SESA_CHECK_ARG(channels == 1 || channels == 2, "sesa_mask_istft: channels must be 1 or 2");
  SESA_CHECK_ARG(n_fft >= 64 && n_fft <= 8192 && (n_fft & (n_fft - 1)) == 0,
                 "sesa_mask_istft: n_fft must be a power of two in [64, 8192], got %d", n_fft);
  SESA_CHECK_ARG(mode >= 0 && mode <= 2, "sesa_mask_istft: bad mode %d", mode);
  SESA_CHECK_ARG(out_len > 0 && out_len <= (int64_t)hop * (n_frames - 1) + n_fft / 2,
                 "sesa_mask_istft: out_len %lld not covered by %d frames", (long long)out_len, n_frames);
  if (batch == 0 || nstems == 0) return SESA_OK;
  if (n_fft == 2048 && hop >= 128 && (reinterpret_cast<uintptr_t>(spec) & 15) == 0 &&
      (mask == nullptr || (reinterpret_cast<uintptr_t>(mask) & 15) == 0))
    return sesa_launch_mask_istft2048(spec, mask, inv_index, inv_count, out, window, envelope, twiddle, batch, nstems,
                                      channels, hop, n_frames, out_len, mode, n_gathered, (cudaStream_t)stream);
  int groups_of = 12;  // frames' worth of hops per CTA: (G + n_fft/hop - 1)/G redundant transforms
  while (groups_of > 1 &&
         (size_t)2 * n_fft * sizeof(float2) + (size_t)channels * groups_of * hop * sizeof(float) > 200 * 1024)
    --groups_of;   // n_fft 8192 / hop 1024 (MDX23C): 128 KB of FFT buffers leave room for 8 hops
  int seg = groups_of * hop;
  const int ngroups = (int)ceil_div64(out_len, seg);
  const size_t smem = (size_t)2 * n_fft * sizeof(float2) + (size_t)channels * seg * sizeof(float);
  SESA_CHECK_ARG(smem <= 200 * 1024, "sesa_mask_istft: segment does not fit in shared memory");
  SESA_CUDA(cudaFuncSetAttribute(mask_istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ngroups, nstems, batch);
  mask_istft_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(
      spec, mask, inv_index, inv_count, out, window, envelope, reinterpret_cast<const float2*>(twiddle),
      channels, n_fft, hop, n_frames, out_len, seg, nstems, mode, n_gathered);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}
