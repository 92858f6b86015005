// fp32 SIMT grouped GEMM with fused epilogues:  C = epi( A[M,K] . W[N,K]^T ).
//
// This is the exact-fp32 arithmetic path of the library (true IEEE fp32 FMA accumulation): it serves the
// ragged per-band GEMMs (BandSplit, bs_roformer.py:241-249; MaskEstimator, :301-310), MDX23C's small
// contractions and is the in-library cross-check for the tcgen05 kernels in gemm_tc.cu.
// One launch runs `n_groups` independent problems (gridDim.z) described by sesa_gemm_group records in
// device memory, so the 62 per-band Linears of one stage are ONE launch instead of 62.
#include "common.cuh"
#include "sesa_b200.h"

#define BM 128
#define BN 128
#define BK 8
#define PADM (BM + 4)

__device__ __forceinline__ float4 load4_guard(const float* __restrict__ p, bool ok, bool vec, int k, int K) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) {
    if (vec && k + 3 < K) {
      v = *reinterpret_cast<const float4*>(p + k);
    } else {
      if (k < K) v.x = p[k];
      if (k + 1 < K) v.y = p[k + 1];
      if (k + 2 < K) v.z = p[k + 2];
      if (k + 3 < K) v.w = p[k + 3];
    }
  }
  return v;
}

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == SESA_ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  if (act == SESA_ACT_TANH) return tanhf(v);
  if (act == SESA_ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
  return v;
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(const sesa_gemm_group* __restrict__ groups,
                                                        sesa_gemm_epilogue ep) {
  const sesa_gemm_group g = groups[blockIdx.z];
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  if (m0 >= g.M || n0 >= g.N) return;
  __shared__ __align__(16) float As[2][BK][PADM];
  __shared__ __align__(16) float Ws[2][BK][PADM];
  __shared__ float rowscale[BM];
  const int tid = threadIdx.x;
  const int lrow = tid >> 1;
  const int lk = (tid & 1) * 4;
  const bool a_ok = (m0 + lrow) < g.M;
  const bool w_ok = (n0 + lrow) < g.N;
  const float* Ap = g.A + (int64_t)(m0 + lrow) * g.lda;
  const float* Wp = g.W + (int64_t)(n0 + lrow) * g.ldw;
  const bool vecA = ((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
  const bool vecW = ((g.ldw & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.W) & 15) == 0);
  const int K = g.K;
  const int nk = (K + BK - 1) / BK;
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float ss = 0.f;

  float4 ra = load4_guard(Ap, a_ok, vecA, lk, K);
  float4 rw = load4_guard(Wp, w_ok, vecW, lk, K);
  ss += ra.x * ra.x + ra.y * ra.y + ra.z * ra.z + ra.w * ra.w;
  As[0][lk + 0][lrow] = ra.x; As[0][lk + 1][lrow] = ra.y; As[0][lk + 2][lrow] = ra.z; As[0][lk + 3][lrow] = ra.w;
  Ws[0][lk + 0][lrow] = rw.x; Ws[0][lk + 1][lrow] = rw.y; Ws[0][lk + 2][lrow] = rw.z; Ws[0][lk + 3][lrow] = rw.w;
  __syncthreads();

  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) {
      ra = load4_guard(Ap, a_ok, vecA, (kt + 1) * BK + lk, K);
      rw = load4_guard(Wp, w_ok, vecW, (kt + 1) * BK + lk, K);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      const int nb = buf ^ 1;
      ss += ra.x * ra.x + ra.y * ra.y + ra.z * ra.z + ra.w * ra.w;
      As[nb][lk + 0][lrow] = ra.x; As[nb][lk + 1][lrow] = ra.y; As[nb][lk + 2][lrow] = ra.z; As[nb][lk + 3][lrow] = ra.w;
      Ws[nb][lk + 0][lrow] = rw.x; Ws[nb][lk + 1][lrow] = rw.y; Ws[nb][lk + 2][lrow] = rw.z; Ws[nb][lk + 3][lrow] = rw.w;
    }
    __syncthreads();
  }

  if (ep.rownorm) {
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    if ((tid & 1) == 0) rowscale[lrow] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    __syncthreads();
  }

  // ---- epilogue: two row halves x two column quads per thread
  const bool vecC = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) && !ep.glu;
#pragma unroll
  for (int ih = 0; ih < 2; ++ih) {
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int rl = ih * 64 + ty * 4 + ii;
      const int m = m0 + rl;
      if (m >= g.M) continue;
      const float rs = ep.rownorm ? rowscale[rl] : 1.0f;
      int pos = 0;
      if (ep.rot_cols > 0) pos = (m / ep.pos_div) % ep.pos_mod;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        const int c = n0 + jh * 64 + tx * 4;
        if (c >= g.N) continue;
        float v[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float x = acc[ih * 4 + ii][jh * 4 + jj] * rs;
          if (g.bias != nullptr && c + jj < g.N) x += g.bias[c + jj];
          v[jj] = act_apply(x, ep.act);
        }
        if (c < ep.rot_cols) {  // rot_cols is a multiple of 4 columns (pairs never straddle a quad)
          const int half = ep.rot_dim >> 1;
#pragma unroll
          for (int pp = 0; pp < 2; ++pp) {
            const int d = ((c + 2 * pp) % ep.rot_dim) >> 1;
            const float2 cs = reinterpret_cast<const float2*>(ep.rot)[(int64_t)pos * half + d];
            const float x1 = v[2 * pp], x2 = v[2 * pp + 1];
            v[2 * pp] = x1 * cs.x - x2 * cs.y;
            v[2 * pp + 1] = x2 * cs.x + x1 * cs.y;
          }
        }
        if (ep.glu) {
          // weight rows were interleaved (value, gate) at load time: out[:, c/2 + p] = v[2p]*sigmoid(v[2p+1])
          float* crow = g.C + (int64_t)m * g.ldc + (c >> 1);
          if (c + 1 < g.N) crow[0] = v[0] * (1.0f / (1.0f + expf(-v[1])));
          if (c + 3 < g.N) crow[1] = v[2] * (1.0f / (1.0f + expf(-v[3])));
        } else {
          float* crow = g.C + (int64_t)m * g.ldc + c;
          if (vecC && c + 3 < g.N) {
            float4 o = make_float4(v[0], v[1], v[2], v[3]);
            if (ep.residual) {
              const float4 r = *reinterpret_cast<const float4*>(crow);
              o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
            }
            *reinterpret_cast<float4*>(crow) = o;
          } else {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              if (c + jj < g.N) crow[jj] = ep.residual ? crow[jj] + v[jj] : v[jj];
          }
        }
      }
    }
  }
}

extern "C" int sesa_gemm_simt(const sesa_gemm_group* groups_dev, int n_groups, int max_m, int max_n,
                              const sesa_gemm_epilogue* ep, void* stream) {
  SESA_CHECK_ARG(n_groups >= 0 && n_groups <= 65535, "sesa_gemm_simt: bad group count %d", n_groups);
  SESA_CHECK_ARG(ep != nullptr, "sesa_gemm_simt: null epilogue");
  SESA_CHECK_ARG(ep->rot_cols == 0 || (ep->rot != nullptr && ep->rot_dim > 0 && (ep->rot_dim & 3) == 0 &&
                                       (ep->rot_cols & 3) == 0 && ep->pos_div > 0 && ep->pos_mod > 0),
                 "sesa_gemm_simt: bad rotary parameters");
  SESA_CHECK_ARG(!(ep->glu && ep->residual), "sesa_gemm_simt: glu and residual are exclusive");
  if (n_groups == 0 || max_m <= 0 || max_n <= 0) return SESA_OK;
  dim3 grid((max_n + BN - 1) / BN, (max_m + BM - 1) / BM, n_groups);
  SESA_CHECK_ARG(grid.y <= 65535, "sesa_gemm_simt: M too large (%d rows)", max_m);
  gemm_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(groups_dev, *ep);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}
