// Streaming windowed overlap-add of the demix() loop, slice framing for chunk-range shards, and the small
// elementwise combiners either side of demix (test-time augmentation, waveform ensembling).
//
// overlap_accumulate is the product path of utils.py:439-464: it is called once per engine batch with the model outputs
// of chunks [k0, k0 + nb) and folds them into the track result WITHOUT keeping every chunk's output resident.  The
// padded mix is cut into `step`-long regions; all samples of region r are covered by the same chunks
// [max(0, r - span + 1), min(r, n_chunks - 1)], so the chunk list is block-uniform.  Per sample the chunks are added in
// ascending order with separate multiply and add (the order and rounding of `result += x * window`):
//   * a region whose earlier chunks were folded by a previous call (or by the previous RANK, whose raw sums arrive as
//     the halo) starts from `partial`; otherwise it starts from 0 and `partial` is never read;
//   * a region whose later chunks are still to come writes its raw sums back to `partial`; a region whose last chunk is
//     in this batch is finished here: divided by the window sum of the GLOBAL schedule, NaN -> 0, cropped, written to
//     `out`.
// Memory is O(result + one batch) instead of O(all chunk outputs), and the result is bit-identical to the one-shot
// gather kernels of misc.cu (tests/test_gpu_kernels.py).
#include "common.cuh"
#include "sesa_b200.h"

namespace {

__device__ __forceinline__ float fade_window(const float* __restrict__ w, int o, int L, int fade, int kind) {
  if (kind == 1 && o < fade) return 1.0f;
  if (kind == 2 && o >= L - fade) return 1.0f;
  return w[o];
}

struct AccArgs {
  const float* y;          // [nb][nc][L]
  const int64_t* starts;   // global schedule (device)
  const int64_t* lens;
  const int32_t* kinds;
  const float* window;
  float* partial;          // [nc][part_ld], padded positions [part_p0, part_p0 + part_ld)
  float* out;              // [nc][out_ld], cropped positions [out_q0, out_q0 + out_cols)
  int64_t part_ld, part_p0, out_ld, out_q0, out_cols, crop, out_len, padded_len;
  int k0, nb, n_chunks, step, L, fade, nc, span, r_begin, blocks_per_region;
};

template <int NK, int V>
__global__ void __launch_bounds__(256) overlap_accumulate_kernel(const AccArgs a) {
  const int r = a.r_begin + blockIdx.x / a.blocks_per_region;
  const int i0 = ((blockIdx.x % a.blocks_per_region) * 256 + threadIdx.x) * V;   // offset inside the region
  if (i0 >= a.step) return;
  const int64_t p = (int64_t)r * a.step + i0;
  if (p >= a.padded_len) return;
  const int kf = max(0, r - a.span + 1);
  const int kl = min(r, a.n_chunks - 1);
  const bool seeded = kf < a.k0;                 // earlier chunks of this region are already in `partial`
  const bool complete = kl < a.k0 + a.nb;        // no later chunk will touch this region
  float cnt[V];
  float wq[NK][V];     // window value per covering chunk and sample; < 0: sample not covered
  int oq[NK];          // offset of the first sample inside chunk kf + j, or -1
#pragma unroll
  for (int e = 0; e < V; ++e) cnt[e] = 0.f;
#pragma unroll
  for (int j = 0; j < NK; ++j) {
    const int k = kf + j;
    oq[j] = -1;
#pragma unroll
    for (int e = 0; e < V; ++e) wq[j][e] = -1.0f;
    if (k <= kl) {
      const int64_t o64 = p - a.starts[k];
      const int n = (int)a.lens[k];
      const int kind = a.kinds[k];
      if (o64 >= 0 && o64 < n) {
        const int o = (int)o64;
        oq[j] = o;
        if (V == 4 && o + 3 < n) {     // whole quad inside the chunk: one 128-bit window load (o, L and fade positions: o % 4 == 0)
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.window + o));
          float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int e = 0; e < V; ++e) {
            if ((kind == 1 && o + e < a.fade) || (kind == 2 && o + e >= a.L - a.fade)) w[e] = 1.0f;
            wq[j][e] = w[e];
            cnt[e] = __fadd_rn(cnt[e], w[e]);
          }
        } else {
#pragma unroll
          for (int e = 0; e < V; ++e) {
            if (o + e < n) {
              const float w = fade_window(a.window, o + e, a.L, a.fade, kind);
              wq[j][e] = w;
              cnt[e] = __fadd_rn(cnt[e], w);
            }
          }
        }
      }
    }
  }
  const int64_t q = p - a.crop;                   // cropped coordinate of the first sample
  for (int sc = 0; sc < a.nc; ++sc) {
    float acc[V];
    float* pp = a.partial + (int64_t)sc * a.part_ld + (p - a.part_p0);
    if (seeded) {
      if (V == 4) {
        const float4 v = *reinterpret_cast<const float4*>(pp);
        acc[0] = v.x; acc[1 % V] = v.y; acc[2 % V] = v.z; acc[3 % V] = v.w;
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = pp[e];
      }
    } else {
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < NK; ++j) {
      const int k = kf + j;
      if (oq[j] < 0 || k < a.k0 || k >= a.k0 + a.nb) continue;
      const float* yp = a.y + ((int64_t)(k - a.k0) * a.nc + sc) * a.L + oq[j];
      if (V == 4 && wq[j][V - 1] >= 0.f) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(yp));     // streamed once
        acc[0] = __fadd_rn(acc[0], __fmul_rn(v.x, wq[j][0]));
        acc[1 % V] = __fadd_rn(acc[1 % V], __fmul_rn(v.y, wq[j][1 % V]));
        acc[2 % V] = __fadd_rn(acc[2 % V], __fmul_rn(v.z, wq[j][2 % V]));
        acc[3 % V] = __fadd_rn(acc[3 % V], __fmul_rn(v.w, wq[j][3 % V]));
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e)
          if (wq[j][e] >= 0.f) acc[e] = __fadd_rn(acc[e], __fmul_rn(yp[e], wq[j][e]));
      }
    }
    if (!complete) {
      if (V == 4) *reinterpret_cast<float4*>(pp) = make_float4(acc[0], acc[1 % V], acc[2 % V], acc[3 % V]);
      else {
#pragma unroll
        for (int e = 0; e < V; ++e) pp[e] = acc[e];
      }
      continue;
    }
    float res[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      res[e] = acc[e] / cnt[e];
      if (res[e] != res[e]) res[e] = 0.f;          // nan_to_num(nan=0) of 0/0 (utils.py:459)
    }
    const int64_t c0 = q - a.out_q0;               // column inside `out`
    float* op = a.out + (int64_t)sc * a.out_ld + c0;
    if (V == 4 && q >= 0 && q + 3 < a.out_len && c0 >= 0 && c0 + 3 < a.out_cols && p + 3 < a.padded_len) {
      __stcs(reinterpret_cast<float4*>(op), make_float4(res[0], res[1 % V], res[2 % V], res[3 % V]));
    } else {
#pragma unroll
      for (int e = 0; e < V; ++e)
        if (q + e >= 0 && q + e < a.out_len && c0 + e >= 0 && c0 + e < a.out_cols && p + e < a.padded_len) op[e] = res[e];
    }
  }
}

template <int NK, int V>
int launch_acc(const AccArgs& a, int regions, cudaStream_t st) {
  overlap_accumulate_kernel<NK, V><<<(unsigned)((int64_t)regions * a.blocks_per_region), 256, 0, st>>>(a);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

}  // namespace

extern "C" int sesa_overlap_accumulate(const float* y, int k0, int nb, const int64_t* starts, const int64_t* lens,
                                       const int32_t* kinds, int n_chunks, int64_t step, int64_t chunk_size, int fade,
                                       const float* window, int nstems, int channels, int64_t padded_len,
                                       int r_begin, int r_end, float* partial, int64_t part_ld, int64_t part_p0,
                                       int64_t crop, int64_t out_len, float* out, int64_t out_ld, int64_t out_q0,
                                       int64_t out_cols, void* stream) {
  SESA_CHECK_ARG(step > 0 && chunk_size >= step && chunk_size < (1ll << 30), "sesa_overlap_accumulate: bad step %lld / chunk %lld",
                 (long long)step, (long long)chunk_size);
  SESA_CHECK_ARG(k0 >= 0 && nb >= 0 && k0 + nb <= n_chunks, "sesa_overlap_accumulate: chunk range [%d, %d) outside the schedule",
                 k0, k0 + nb);
  const int64_t span = ceil_div64(chunk_size, step);
  SESA_CHECK_ARG(span <= 32, "sesa_overlap_accumulate: num_overlap above 32 is not supported (chunk %lld / step %lld)",
                 (long long)chunk_size, (long long)step);
  const int64_t n_regions = ceil_div64(padded_len, step);
  if (r_end > n_regions) r_end = (int)n_regions;
  if (r_begin < 0) r_begin = 0;
  if (r_end <= r_begin) return SESA_OK;
  SESA_CHECK_ARG(part_p0 <= (int64_t)r_begin * step && part_p0 + part_ld >= min((int64_t)r_end * step, (int64_t)((padded_len + 3) / 4 * 4)),
                 "sesa_overlap_accumulate: partial-sum slab [%lld, %lld) does not cover regions [%d, %d)", (long long)part_p0,
                 (long long)(part_p0 + part_ld), r_begin, r_end);
  AccArgs a;
  a.y = y; a.starts = starts; a.lens = lens; a.kinds = kinds; a.window = window; a.partial = partial; a.out = out;
  a.part_ld = part_ld; a.part_p0 = part_p0; a.out_ld = out_ld; a.out_q0 = out_q0; a.out_cols = out_cols; a.crop = crop;
  a.out_len = out_len; a.padded_len = padded_len; a.k0 = k0; a.nb = nb; a.n_chunks = n_chunks; a.step = (int)step;
  a.L = (int)chunk_size; a.fade = fade; a.nc = nstems * channels; a.span = (int)span; a.r_begin = r_begin;
  const bool vec = (step & 3) == 0 && (chunk_size & 3) == 0 && (crop & 3) == 0 && (part_ld & 3) == 0 && (part_p0 & 3) == 0 &&
                   (out_ld & 3) == 0 && (out_q0 & 3) == 0 && span <= 8 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(partial) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(window) & 15) == 0;
  const int regions = r_end - r_begin;
  cudaStream_t st = (cudaStream_t)stream;
  if (vec) {
    a.blocks_per_region = (int)ceil_div64(ceil_div64(step, 4), 256);
    return span <= 4 ? launch_acc<4, 4>(a, regions, st) : launch_acc<8, 4>(a, regions, st);
  }
  a.blocks_per_region = (int)ceil_div64(step, 256);
  if (span <= 4) return launch_acc<4, 1>(a, regions, st);
  if (span <= 8) return launch_acc<8, 1>(a, regions, st);
  return launch_acc<32, 1>(a, regions, st);
}

// ---------------------------------------------------------------------------------------------
// dst[c][i] = mix[c][reflect(p0 + i - left)] for i in [0, count), reading the mix through a window that holds only
// mix[:, src_off : src_off + src_cols): the border reflect pad of utils.py:391-393 for ONE slice of the padded mix (a
// chunk-range shard uploads just the samples its chunks touch).
__global__ void pad_reflect_slice_kernel(const float* __restrict__ src, int64_t src_ld, int64_t src_off, float* __restrict__ dst,
                                         int64_t len, int64_t left, int64_t p0, int64_t count) {
  const int c = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    dst[c * count + i] = src[c * src_ld + (reflect_index(p0 + i - left, len) - src_off)];
}

extern "C" int sesa_pad_reflect_slice(const float* src, int64_t src_cols, int64_t src_off, float* dst, int channels,
                                      int64_t len, int64_t left, int64_t p0, int64_t count, void* stream) {
  if (count <= 0) return SESA_OK;
  SESA_CHECK_ARG(left < len || left == 0, "sesa_pad_reflect_slice: pad %lld must be < len %lld", (long long)left, (long long)len);
  // host-side range check of the window (reflect_index is monotone on each side of the borders)
  int64_t lo = len, hi = -1;
  const int64_t probe[4] = {p0 - left, p0 + count - 1 - left, 0, len - 1};
  for (int t = 0; t < 4; ++t) {
    int64_t j = probe[t];
    if (t >= 2 && !(p0 - left <= j && j <= p0 + count - 1 - left)) continue;
    if (j < 0) j = -j;
    if (j >= len) j = 2 * (len - 1) - j;
    lo = j < lo ? j : lo;
    hi = j > hi ? j : hi;
  }
  SESA_CHECK_ARG(lo >= src_off && hi < src_off + src_cols, "sesa_pad_reflect_slice: slice needs mix[%lld, %lld] but the window holds [%lld, %lld)",
                 (long long)lo, (long long)hi, (long long)src_off, (long long)(src_off + src_cols));
  dim3 grid((unsigned)min((int64_t)4096, ceil_div64(count, 256)), channels);
  pad_reflect_slice_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_cols, src_off, dst, len, left, p0, count);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

// ---------------------------------------------------------------------------------------------
// Test-time augmentation (utils.py:241-292): the two augmented mixes (channel order reversed; polarity inverted) and
// the combination of the three estimates, in the reference's order of operations:
//   orig += swapped_est[::-1];  orig -= negated_est;  orig /= 3
__global__ void tta_variants_kernel(const float* __restrict__ mix, float* __restrict__ swapped, float* __restrict__ negated,
                                    int channels, int64_t len) {
  const int c = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = mix[c * len + i];
    swapped[(int64_t)(channels - 1 - c) * len + i] = v;
    negated[c * len + i] = -1.0f * v;
  }
}

extern "C" int sesa_tta_variants(const float* mix, float* swapped, float* negated, int channels, int64_t len, void* stream) {
  if (len <= 0) return SESA_OK;
  dim3 grid((unsigned)min((int64_t)2048, ceil_div64(len, 256)), channels);
  tta_variants_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mix, swapped, negated, channels, len);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

__global__ void tta_combine_kernel(const float* __restrict__ orig, const float* __restrict__ swapped, const float* __restrict__ negated,
                                   float* __restrict__ out, int channels, int64_t len, float count) {
  const int n = blockIdx.z, c = blockIdx.y;
  const int64_t row = ((int64_t)n * channels + c) * len, srow = ((int64_t)n * channels + (channels - 1 - c)) * len;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
    float v = __fadd_rn(orig[row + i], swapped[srow + i]);
    v = __fsub_rn(v, negated[row + i]);
    out[row + i] = __fdiv_rn(v, count);
  }
}

extern "C" int sesa_tta_combine(const float* orig, const float* swapped_est, const float* negated_est, float* out, int nstems,
                                int channels, int64_t len, void* stream) {
  if (len <= 0 || nstems <= 0) return SESA_OK;
  dim3 grid((unsigned)min((int64_t)2048, ceil_div64(len, 256)), channels, nstems);
  tta_combine_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(orig, swapped_est, negated_est, out, channels, len, 3.0f);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

// ---------------------------------------------------------------------------------------------
// Waveform-domain ensembling (ensemble.py:172-183 process_waveform): the reference reads its inputs as float64 and
// reduces over the model axis with numpy, i.e. sequentially in float64; so does this kernel (method 0: mean = sum / M, or
// sum(x*w) / sum(w) with weights; 1: median (mean of the two middle values for even M); 2: max; 3: min).
#define SESA_ENSEMBLE_MAX_INPUTS 16
struct EnsembleArgs {
  const float* in[SESA_ENSEMBLE_MAX_INPUTS];
  double w[SESA_ENSEMBLE_MAX_INPUTS];
  double wsum;
  int m, method, weighted;
};

__global__ void __launch_bounds__(256) ensemble_wave_kernel(const EnsembleArgs a, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double r;
    if (a.method == 0) {
      double acc = 0.0;
      if (a.weighted) {
        for (int m = 0; m < a.m; ++m) acc = __dadd_rn(acc, __dmul_rn((double)a.in[m][i], a.w[m]));
        r = acc / a.wsum;
      } else {
        for (int m = 0; m < a.m; ++m) acc = __dadd_rn(acc, (double)a.in[m][i]);
        r = acc / (double)a.m;
      }
    } else if (a.method == 1) {
      float v[SESA_ENSEMBLE_MAX_INPUTS];
      for (int m = 0; m < a.m; ++m) {       // insertion sort (M is a handful of models)
        float x = a.in[m][i];
        int j = m;
        while (j > 0 && v[j - 1] > x) { v[j] = v[j - 1]; --j; }
        v[j] = x;
      }
      r = (a.m & 1) ? (double)v[a.m / 2] : ((double)v[a.m / 2 - 1] + (double)v[a.m / 2]) / 2.0;
    } else {
      float x = a.in[0][i];
      for (int m = 1; m < a.m; ++m) x = a.method == 2 ? fmaxf(x, a.in[m][i]) : fminf(x, a.in[m][i]);
      r = (double)x;
    }
    out[i] = (float)r;
  }
}

extern "C" int sesa_ensemble_wave(const float* const* inputs_host, int n_inputs, const double* weights_host, int method, float* out,
                                  int64_t n, void* stream) {
  SESA_CHECK_ARG(n_inputs >= 1 && n_inputs <= SESA_ENSEMBLE_MAX_INPUTS, "sesa_ensemble_wave: 1..%d inputs supported, got %d",
                 SESA_ENSEMBLE_MAX_INPUTS, n_inputs);
  SESA_CHECK_ARG(method >= 0 && method <= 3, "sesa_ensemble_wave: method must be 0 (avg) 1 (median) 2 (max) 3 (min)");
  if (n <= 0) return SESA_OK;
  EnsembleArgs a;
  a.m = n_inputs; a.method = method; a.weighted = weights_host != nullptr; a.wsum = 0.0;
  for (int m = 0; m < n_inputs; ++m) {
    a.in[m] = inputs_host[m];
    a.w[m] = weights_host ? weights_host[m] : 1.0;
    a.wsum += a.w[m];
  }
  ensemble_wave_kernel<<<(unsigned)min((int64_t)148 * 8, ceil_div64(n, 256)), 256, 0, (cudaStream_t)stream>>>(a, out, n);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}
