// Framing, overlap-add and small elementwise kernels of the demix() loop.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"
#include "sesa_b200.h"

static thread_local char g_err[512] = "";

void sesa_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* sesa_last_error(void) { return g_err; }
extern "C" int sesa_abi_version(void) { return SESA_B200_ABI_VERSION; }

extern "C" int sesa_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem) {
  int dev = 0;
  SESA_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  SESA_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (total_mem) *total_mem = (int64_t)p.totalGlobalMem;
  return SESA_OK;
}

// ---------------------------------------------------------------------------------------------
__global__ void pad_reflect_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t len,
                                   int64_t left, int64_t total) {
  const int c = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    dst[c * total + i] = src[c * len + reflect_index(i - left, len)];
}

extern "C" int sesa_pad_reflect(const float* src, float* dst, int channels, int64_t len, int64_t left,
                                int64_t right, void* stream) {
  SESA_CHECK_ARG(left < len && right < len, "sesa_pad_reflect: pad (%lld,%lld) must be < len %lld",
                 (long long)left, (long long)right, (long long)len);
  const int64_t total = len + left + right;
  dim3 grid((unsigned)min((int64_t)4096, ceil_div64(total, 256)), channels);
  pad_reflect_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, len, left, total);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

// 128-bit form of frame_chunks_kernel for 4-aligned geometry (chunk starts, lengths' interior and L multiples of 4).
__global__ void __launch_bounds__(256) frame_chunks_vec4_kernel(const float* __restrict__ mix, int64_t mix_len, int channels,
                                                                const int64_t* __restrict__ starts,
                                                                const int64_t* __restrict__ lens,
                                                                const int32_t* __restrict__ modes, int64_t L,
                                                                float* __restrict__ chunks) {
  const int k = blockIdx.z, c = blockIdx.y;
  const int64_t s = starts[k], n = lens[k];
  const int mode = modes[k];
  const float* src = mix + c * mix_len + s;
  float* dst = chunks + ((int64_t)k * channels + c) * L;
  const bool aligned = ((s & 3) == 0) && ((mix_len & 3) == 0);
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < L; i += (int64_t)gridDim.x * blockDim.x * 4) {
    float4 v;
    if (aligned && i + 3 < n) {
      v = *reinterpret_cast<const float4*>(src + i);
    } else {
      float e[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t ii = i + j;
        e[j] = ii < n ? src[ii] : (mode == 1 ? src[2 * (n - 1) - ii] : 0.f);
      }
      v = make_float4(e[0], e[1], e[2], e[3]);
    }
    *reinterpret_cast<float4*>(dst + i) = v;
  }
}

__global__ void frame_chunks_kernel(const float* __restrict__ mix, int64_t mix_len, int channels,
                                    const int64_t* __restrict__ starts, const int64_t* __restrict__ lens,
                                    const int32_t* __restrict__ modes, int64_t L, float* __restrict__ chunks) {
  const int k = blockIdx.z, c = blockIdx.y;
  const int64_t s = starts[k], n = lens[k];
  const int mode = modes[k];
  const float* src = mix + c * mix_len + s;
  float* dst = chunks + ((int64_t)k * channels + c) * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (int64_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < n) v = src[i];
    else if (mode == 1) v = src[2 * (n - 1) - i];
    dst[i] = v;
  }
}

extern "C" int sesa_frame_chunks(const float* mix, int64_t mix_len, int channels, const int64_t* starts,
                                 const int64_t* lens, const int32_t* modes, int n_chunks, int64_t chunk_size,
                                 float* chunks, void* stream) {
  if (n_chunks == 0) return SESA_OK;
  SESA_CHECK_ARG(n_chunks <= 65535, "sesa_frame_chunks: too many chunks in one call (%d)", n_chunks);
  if ((chunk_size & 3) == 0 && (reinterpret_cast<uintptr_t>(mix) & 15) == 0 && (reinterpret_cast<uintptr_t>(chunks) & 15) == 0) {
    dim3 gridv((unsigned)min((int64_t)128, ceil_div64(chunk_size, 1024)), channels, n_chunks);
    frame_chunks_vec4_kernel<<<gridv, 256, 0, (cudaStream_t)stream>>>(mix, mix_len, channels, starts, lens, modes,
                                                                     chunk_size, chunks);
    SESA_LAUNCH_CHECK();
    return SESA_OK;
  }
  dim3 grid((unsigned)min((int64_t)256, ceil_div64(chunk_size, 256)), channels, n_chunks);
  frame_chunks_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mix, mix_len, channels, starts, lens, modes,
                                                              chunk_size, chunks);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

// ---------------------------------------------------------------------------------------------
// Overlap-add as a GATHER: each output sample sums the chunks that cover it in ascending chunk order
// with separate multiply and add, which is the order and rounding of the reference's
// `result[...] += x * window; counter[...] += window` loop (utils.py:439-442).
__device__ __forceinline__ float demix_window(const float* __restrict__ w, int64_t o, int64_t L, int fade, int kind) {
  if (kind == 1 && o < fade) return 1.0f;
  if (kind == 2 && o >= L - fade) return 1.0f;
  return w[o];
}

__global__ void overlap_add_kernel(const float* __restrict__ y, const int64_t* __restrict__ starts,
                                   const int64_t* __restrict__ lens, const int32_t* __restrict__ kinds,
                                   int n_chunks, int64_t step, int64_t L, int fade,
                                   const float* __restrict__ window, int nstems, int channels,
                                   int64_t padded_len, int64_t crop, int64_t out_len,
                                   float* __restrict__ result, float* __restrict__ counter) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = counter ? padded_len : out_len;
  if (i >= total) return;
  const int64_t p = counter ? i : i + crop;  // padded coordinate
  int64_t k_hi = p / step;
  if (k_hi > n_chunks - 1) k_hi = n_chunks - 1;
  int64_t k_lo = (p - L + step) / step;  // ceil((p - L + 1)/step) for p-L+1 > 0
  if (p - L + 1 <= 0) k_lo = 0;
  float cnt = 0.f;
  for (int64_t k = k_lo; k <= k_hi; ++k) {
    const int64_t o = p - starts[k];
    if (o < 0 || o >= lens[k]) continue;
    cnt = __fadd_rn(cnt, demix_window(window, o, L, fade, kinds[k]));
  }
  if (counter) counter[p] = cnt;
  const int64_t io = p - crop;
  if (io < 0 || io >= out_len) return;
  const int nc = nstems * channels;
  for (int sc = 0; sc < nc; ++sc) {
    float acc = 0.f;
    for (int64_t k = k_lo; k <= k_hi; ++k) {
      const int64_t o = p - starts[k];
      if (o < 0 || o >= lens[k]) continue;
      const float w = demix_window(window, o, L, fade, kinds[k]);
      acc = __fadd_rn(acc, __fmul_rn(y[((int64_t)k * nc + sc) * L + o], w));
    }
    float r = acc / cnt;
    if (r != r) r = 0.f;  // nan_to_num(nan=0) of 0/0 (utils.py:459)
    result[(int64_t)sc * out_len + io] = r;
  }
}

// Vectorised form of overlap_add_kernel: one thread finishes 4 consecutive output samples with 128-bit loads of the
// chunk outputs and of the window (same ascending-chunk accumulation per sample, separate multiply and add, so the
// result and the counter stay bit-identical to the scalar kernel).  Requires step, L, crop and the chunk starts to be
// multiples of 4 and 16-byte aligned rows (checked by the host entry); ragged chunk ends fall back to per-sample code.
__global__ void __launch_bounds__(256) overlap_add_vec4_kernel(
    const float* __restrict__ y, const int64_t* __restrict__ starts, const int64_t* __restrict__ lens,
    const int32_t* __restrict__ kinds, int n_chunks, int64_t step, int64_t L, int fade, const float* __restrict__ window,
    int nc, int64_t padded_len, int64_t crop, int64_t out_len, float* __restrict__ result, float* __restrict__ counter) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int64_t total = counter ? padded_len : out_len;
  if (i4 >= total) return;
  const int64_t p = counter ? i4 : i4 + crop;   // padded coordinate of the first of the 4 samples (multiple of 4)
  int64_t k_hi = (p + 3) / step;
  if (k_hi > n_chunks - 1) k_hi = n_chunks - 1;
  int64_t k_lo = (p - L + step) / step;
  if (p - L + 1 <= 0) k_lo = 0;
  float cnt[4] = {0.f, 0.f, 0.f, 0.f};
  float wq[8][4];   // window values of the up to 8 covering chunks k_lo + j (statically indexed: stays in registers)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t k = k_lo + j;
#pragma unroll
    for (int e = 0; e < 4; ++e) wq[j][e] = -1.0f;   // sentinel: sample not covered by this chunk
    if (k <= k_hi) {
      const int64_t o = p - starts[k];
      const int64_t n = lens[k];
      const int kind = kinds[k];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t oe = o + e;
        if (oe >= 0 && oe < n) {
          const float w = demix_window(window, oe, L, fade, kind);
          wq[j][e] = w;
          cnt[e] = __fadd_rn(cnt[e], w);
        }
      }
    }
  }
  if (counter) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (p + e < padded_len) counter[p + e] = cnt[e];
  }
  const int64_t io = p - crop;
  if (io + 3 < 0 || io >= out_len) return;
  for (int sc = 0; sc < nc; ++sc) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t k = k_lo + j;
      if (k > k_hi) continue;
      const int64_t o = p - starts[k];
      const float* yp = y + ((int64_t)k * nc + sc) * L + o;
      if (o >= 0 && o + 3 < lens[k]) {
        const float4 v = *reinterpret_cast<const float4*>(yp);
        acc[0] = __fadd_rn(acc[0], __fmul_rn(v.x, wq[j][0]));
        acc[1] = __fadd_rn(acc[1], __fmul_rn(v.y, wq[j][1]));
        acc[2] = __fadd_rn(acc[2], __fmul_rn(v.z, wq[j][2]));
        acc[3] = __fadd_rn(acc[3], __fmul_rn(v.w, wq[j][3]));
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (wq[j][e] >= 0.f) acc[e] = __fadd_rn(acc[e], __fmul_rn(yp[e], wq[j][e]));
      }
    }
    float r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      r[e] = acc[e] / cnt[e];
      if (r[e] != r[e]) r[e] = 0.f;
    }
    float* rp = result + (int64_t)sc * out_len + io;
    if (io >= 0 && io + 3 < out_len) *reinterpret_cast<float4*>(rp) = make_float4(r[0], r[1], r[2], r[3]);
    else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (io + e >= 0 && io + e < out_len) rp[e] = r[e];
    }
  }
}

// Region form of the vectorised gather (the product path: no counter output).  All samples of one `step`-long region
// [r * step, (r + 1) * step) of the padded mix are covered by the same chunks, so a block works inside ONE region and the
// chunk list (starts, lengths, window kinds) is block-uniform: no per-thread 64-bit divisions or schedule loads, only the
// 128-bit loads of the chunk outputs and of the window.  Arithmetic per sample is unchanged (ascending chunk order, separate
// multiply and add, divide by the window sum), so the result is bit-identical to the kernels above.
template <int NK>
__global__ void __launch_bounds__(256) overlap_add_region_kernel(
    const float* __restrict__ y, const int64_t* __restrict__ starts, const int64_t* __restrict__ lens,
    const int32_t* __restrict__ kinds, int n_chunks, int step, int L, int fade, const float* __restrict__ window, int nc,
    int64_t crop, int64_t out_len, float* __restrict__ result, int blocks_per_region, int span) {
  const int r = blockIdx.x / blocks_per_region;
  const int i4 = ((blockIdx.x - r * blocks_per_region) * 256 + threadIdx.x) * 4;   // offset inside the region
  if (i4 >= step) return;
  const int64_t p = (int64_t)r * step + i4;
  const int64_t io = p - crop;
  if (io < 0 || io >= out_len) return;    // crop and out_len are multiples of 4
  const int k_hi = min(r, n_chunks - 1);
  const int k_lo = max(0, r - span + 1);
  float cnt[4] = {0.f, 0.f, 0.f, 0.f};
  float wq[NK][4];
  int oq[NK];          // offset of this thread's first sample inside chunk k_lo + j; -1: chunk absent or not covering
  bool full[NK];
#pragma unroll
  for (int j = 0; j < NK; ++j) {
    const int k = k_lo + j;
    oq[j] = -1;
    full[j] = false;
#pragma unroll
    for (int e = 0; e < 4; ++e) wq[j][e] = -1.0f;
    if (k <= k_hi) {
      const int64_t o64 = p - starts[k];
      const int n = (int)lens[k];
      const int kind = kinds[k];
      if (o64 >= 0 && o64 < n) {
        const int o = (int)o64;
        oq[j] = o;
        if (o + 3 < n) {
          full[j] = true;
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(window + o));
          float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if ((kind == 1 && o + e < fade) || (kind == 2 && o + e >= L - fade)) w[e] = 1.0f;
            wq[j][e] = w[e];
            cnt[e] = __fadd_rn(cnt[e], w[e]);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (o + e < n) {
              const float w = demix_window(window, o + e, L, fade, kind);
              wq[j][e] = w;
              cnt[e] = __fadd_rn(cnt[e], w);
            }
          }
        }
      }
    }
  }
#pragma unroll 2
  for (int sc = 0; sc < nc; ++sc) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NK; ++j) {
      if (oq[j] < 0) continue;
      const float* yp = y + ((int64_t)(k_lo + j) * nc + sc) * L + oq[j];
      if (full[j]) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(yp));   // streamed once: do not keep in L2
        acc[0] = __fadd_rn(acc[0], __fmul_rn(v.x, wq[j][0]));
        acc[1] = __fadd_rn(acc[1], __fmul_rn(v.y, wq[j][1]));
        acc[2] = __fadd_rn(acc[2], __fmul_rn(v.z, wq[j][2]));
        acc[3] = __fadd_rn(acc[3], __fmul_rn(v.w, wq[j][3]));
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (wq[j][e] >= 0.f) acc[e] = __fadd_rn(acc[e], __fmul_rn(yp[e], wq[j][e]));
      }
    }
    float4 o4;
    o4.x = acc[0] / cnt[0]; o4.y = acc[1] / cnt[1]; o4.z = acc[2] / cnt[2]; o4.w = acc[3] / cnt[3];
    if (o4.x != o4.x) o4.x = 0.f;   // nan_to_num(nan=0) of 0/0 (utils.py:459)
    if (o4.y != o4.y) o4.y = 0.f;
    if (o4.z != o4.z) o4.z = 0.f;
    if (o4.w != o4.w) o4.w = 0.f;
    __stcs(reinterpret_cast<float4*>(result + (int64_t)sc * out_len + io), o4);
  }
}

extern "C" int sesa_overlap_add(const float* chunk_out, const int64_t* starts, const int64_t* lens,
                                const int32_t* kinds, int n_chunks, int64_t step, int64_t chunk_size, int fade,
                                const float* window, int nstems, int channels, int64_t padded_len,
                                int64_t crop, int64_t out_len, float* result, float* counter, void* stream) {
  SESA_CHECK_ARG(step > 0 && chunk_size >= step, "sesa_overlap_add: bad step %lld", (long long)step);
  SESA_CHECK_ARG(crop >= 0 && crop + out_len <= padded_len, "sesa_overlap_add: crop range outside the padded mix");
  const int64_t total = counter ? padded_len : out_len;
  if (total == 0) return SESA_OK;
  const bool vec = (step & 3) == 0 && (chunk_size & 3) == 0 && (crop & 3) == 0 && (out_len & 3) == 0 &&
                   chunk_size <= 8 * step && (reinterpret_cast<uintptr_t>(chunk_out) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(result) & 15) == 0;   // chunk starts are multiples of step
  const int64_t span64 = ceil_div64(chunk_size, step);   // chunks covering one region
  if (vec && counter == nullptr && chunk_size < (1ll << 30) && span64 <= 8) {
    const int bpr = (int)ceil_div64(ceil_div64(step, 4), 256);
    const int64_t regions = ceil_div64(padded_len, step);
    const int64_t blocks = regions * bpr;
    if (blocks < (1ll << 31)) {
      if (span64 <= 4)
        overlap_add_region_kernel<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            chunk_out, starts, lens, kinds, n_chunks, (int)step, (int)chunk_size, fade, window, nstems * channels, crop,
            out_len, result, bpr, (int)span64);
      else
        overlap_add_region_kernel<8><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            chunk_out, starts, lens, kinds, n_chunks, (int)step, (int)chunk_size, fade, window, nstems * channels, crop,
            out_len, result, bpr, (int)span64);
      SESA_LAUNCH_CHECK();
      return SESA_OK;
    }
  }
  if (vec) {
    overlap_add_vec4_kernel<<<(unsigned)ceil_div64(ceil_div64(total, 4), 256), 256, 0, (cudaStream_t)stream>>>(
        chunk_out, starts, lens, kinds, n_chunks, step, chunk_size, fade, window, nstems * channels, padded_len, crop,
        out_len, result, counter);
    SESA_LAUNCH_CHECK();
    return SESA_OK;
  }
  overlap_add_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
      chunk_out, starts, lens, kinds, n_chunks, step, chunk_size, fade, window, nstems, channels, padded_len,
      crop, out_len, result, counter);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}


// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                      float* __restrict__ y, int64_t rows, int dim, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * dim;
  float ss = 0.f;
  for (int i = lane; i < dim; i += 32) ss += xr[i] * xr[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  float* yr = y + row * dim;
  for (int i = lane; i < dim; i += 32) yr[i] = xr[i] * inv * scale * gamma[i];
}

extern "C" int sesa_rmsnorm(const float* x, const float* gamma, float* y, int64_t rows, int dim, void* stream) {
  if (rows == 0) return SESA_OK;
  rmsnorm_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, gamma, y, rows, dim,
                                                                                  sqrtf((float)dim));
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

__global__ void add_inplace_kernel(float* __restrict__ y, const float* __restrict__ x, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] += x[i];
}

extern "C" int sesa_add_inplace(float* y, const float* x, int64_t n, void* stream) {
  if (n == 0) return SESA_OK;
  add_inplace_kernel<<<(unsigned)min((int64_t)148 * 16, ceil_div64(n, 256)), 256, 0, (cudaStream_t)stream>>>(y, x, n);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

__global__ void gather_rows_kernel(const float* __restrict__ in, const int32_t* __restrict__ idx,
                                   float* __restrict__ out, int n_in, int n_out, int width) {
  const int64_t r = blockIdx.y;
  const float* src = in + r * (int64_t)n_in * width;
  float* dst = out + r * (int64_t)n_out * width;
  const int total = n_out * width;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    dst[i] = src[idx[i / width] * width + (i % width)];
}

extern "C" int sesa_gather_rows(const float* in, const int32_t* idx, float* out, int64_t rows, int n_in,
                                int n_out, int width, void* stream) {
  if (rows == 0 || n_out == 0) return SESA_OK;
  SESA_CHECK_ARG(rows <= 65535LL * 32768, "sesa_gather_rows: too many rows");
  int64_t done = 0;
  while (done < rows) {  // gridDim.y limit
    const int64_t nr = min((int64_t)65535, rows - done);
    dim3 grid((unsigned)min(8, (n_out * width + 255) / 256), (unsigned)nr);
    gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in + done * (int64_t)n_in * width, idx,
                                                                out + done * (int64_t)n_out * width, n_in,
                                                                n_out, width);
    SESA_LAUNCH_CHECK();
    done += nr;
  }
  return SESA_OK;
}
