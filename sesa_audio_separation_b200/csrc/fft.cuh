// Shared-memory Stockham autosort FFT (radix-4 passes, one radix-2 pass when log2(N) is odd).
// One CTA transforms one length-N complex sequence held in shared memory; N is a power of two.
// tw[k] = exp(-2*pi*i*k/N) (forward table, N entries, built on the host in double precision).
#pragma once
#include "common.cuh"

// Transforms buf0 (N float2) using buf1 as the ping-pong partner.  Returns the buffer holding the
// result.  INVERSE uses conjugated twiddles and does NOT scale by 1/N.
template <bool INVERSE>
__device__ __forceinline__ float2* block_fft(float2* buf0, float2* buf1, int N, const float2* __restrict__ tw) {
  float2* x = buf0;
  float2* y = buf1;
  int n = N;
  int s = 1;
  int log_s = 0;
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  while (n > 1) {
    if ((n & 3) == 0) {
      const int q4 = N >> 2;
      for (int idx = tid; idx < q4; idx += nthr) {
        const int p = idx >> log_s;
        const int q = idx & (s - 1);
        const float2 a = x[idx], b = x[idx + q4], c = x[idx + 2 * q4], d = x[idx + 3 * q4];
        const int k1 = p << log_s;
        float2 w1 = __ldg(tw + k1), w2 = __ldg(tw + 2 * k1), w3 = __ldg(tw + 3 * k1);
        if (INVERSE) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
        const float2 apc = cadd(a, c), amc = csub(a, c), bpd = cadd(b, d), bmd = csub(b, d);
        // J*(b-d): forward J = -i -> (y, -x); inverse J = +i -> (-y, x)
        const float2 jbmd = INVERSE ? make_float2(-bmd.y, bmd.x) : make_float2(bmd.y, -bmd.x);
        const int o = q + ((4 * p) << log_s);
        y[o] = cadd(apc, bpd);
        y[o + s] = cmul(w1, cadd(amc, jbmd));
        y[o + 2 * s] = cmul(w2, csub(apc, bpd));
        y[o + 3 * s] = cmul(w3, csub(amc, jbmd));
      }
      n >>= 2; s <<= 2; log_s += 2;
    } else {
      const int h = N >> 1;
      for (int idx = tid; idx < h; idx += nthr) {
        const int p = idx >> log_s;
        const int q = idx & (s - 1);
        const float2 a = x[idx], b = x[idx + h];
        float2 w = __ldg(tw + (p << log_s));
        if (INVERSE) w.y = -w.y;
        const int o = q + ((2 * p) << log_s);
        y[o] = cadd(a, b);
        y[o + s] = cmul(csub(a, b), w);
      }
      n >>= 1; s <<= 1; log_s += 1;
    }
    __syncthreads();
    float2* t = x; x = y; y = t;
  }
  return x;
}
