// MDX23C (TFC_TDF_net, models/mdx23c_tfc_tdf_v3.py:100-242) support kernels.
//
// The convolutions and Linears of the U-Net run as tensor-core GEMMs (gemm_tc.cu: implicit-GEMM taps through TMA,
// plain GEMMs for 1x1 convs and the TDF Linears).  Activations live channels-last, x[b][t][f][c] fp32, so every
// "norm -> act -> conv" prologue of the reference (get_norm/get_act, :47-71) is ONE elementwise pass that applies
// InstanceNorm2d statistics + affine + GELU and writes the bf16 hi/lo planes the GEMM's TMA loads consume; the
// zero padding of the 3x3 convolutions then falls out of TMA's out-of-bounds fill on the already-activated planes.
#include "common.cuh"
#include "sesa_b200.h"
#include "tc_common.cuh"

namespace {

__device__ __forceinline__ float gelu_exact(float x) {
  // same 14-instruction exact-erf GELU as the GEMM epilogue (gemm_tc.cu: gelu_fast)
  const float u = fminf(fabsf(x), 6.0f);
  float r = -2.834913403e-06f;
  r = fmaf(r, u, 3.937759539e-05f);
  r = fmaf(r, u, -1.861794008e-04f);
  r = fmaf(r, u, -1.369391393e-04f);
  r = fmaf(r, u, 7.063424215e-03f);
  r = fmaf(r, u, -5.249617994e-02f);
  r = fmaf(r, u, -4.592081904e-01f);
  r = fmaf(r, u, -1.151105165e+00f);
  const float e = exp2f(u * r);
  return x * fmaf(0.5f, copysignf(1.0f - e, x), 0.5f);
}
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == SESA_ACT_GELU) return gelu_exact(v);
  if (act == 4) return fmaxf(v, 0.f);  // relu
  return v;
}

// ---- InstanceNorm2d statistics: per (b, c) over all positions ---------------------------------------------
// layout 0 (channels-last): element (b, i, c) at x[(b*n1 + i)*ld + c]
// layout 1 (channel-major rows): element (b, i, c, j) at x[((b*n1 + i)*C + c)*n2 + j]
__global__ void __launch_bounds__(256) instnorm_acc_kernel(const float* __restrict__ x, int layout, int B, int64_t n1, int C,
                                                           int n2, int64_t ld, int64_t chunk, double* __restrict__ acc) {
  const int b = blockIdx.y;
  const int64_t i0 = (int64_t)blockIdx.x * chunk;
  const int64_t i1 = min(n1, i0 + chunk);
  if (layout == 0) {
    // threads = (channel quads) x (position lanes): float4 loads, 512 contiguous bytes per warp and position; four
    // positions in flight per thread (the loop is latency-bound otherwise: deep layers have few positions per block)
    __shared__ float red[256][8];
    const int cq = C >> 2;                               // C % 4 == 0 (checked by the host entry)
    const int nq = cq < 256 ? cq : 256;                  // channel quads handled per pass
    const int lanes = 256 / nq;                          // position lanes
    const int tq = threadIdx.x % nq, tl = threadIdx.x / nq;
    for (int q0 = 0; q0 < cq; q0 += nq) {
      const int qd = q0 + tq;
      float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
      if (qd < cq && tl < lanes) {
        const float* p = x + ((int64_t)b * n1 + i0 + tl) * ld + 4 * qd;
        const int64_t stride = (int64_t)lanes * ld;
        int64_t i = i0 + tl;
        for (; i + 3 * lanes < i1; i += 4 * lanes, p += 4 * stride) {
          const float4 v0 = __ldg(reinterpret_cast<const float4*>(p));
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(p + stride));
          const float4 v2 = __ldg(reinterpret_cast<const float4*>(p + 2 * stride));
          const float4 v3 = __ldg(reinterpret_cast<const float4*>(p + 3 * stride));
          s[0] += (v0.x + v1.x) + (v2.x + v3.x); s[1] += (v0.y + v1.y) + (v2.y + v3.y);
          s[2] += (v0.z + v1.z) + (v2.z + v3.z); s[3] += (v0.w + v1.w) + (v2.w + v3.w);
          ss[0] += fmaf(v0.x, v0.x, v1.x * v1.x) + fmaf(v2.x, v2.x, v3.x * v3.x);
          ss[1] += fmaf(v0.y, v0.y, v1.y * v1.y) + fmaf(v2.y, v2.y, v3.y * v3.y);
          ss[2] += fmaf(v0.z, v0.z, v1.z * v1.z) + fmaf(v2.z, v2.z, v3.z * v3.z);
          ss[3] += fmaf(v0.w, v0.w, v1.w * v1.w) + fmaf(v2.w, v2.w, v3.w * v3.w);
        }
        for (; i < i1; i += lanes, p += stride) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p));
          s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
          ss[0] = fmaf(v.x, v.x, ss[0]); ss[1] = fmaf(v.y, v.y, ss[1]);
          ss[2] = fmaf(v.z, v.z, ss[2]); ss[3] = fmaf(v.w, v.w, ss[3]);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { red[threadIdx.x][e] = s[e]; red[threadIdx.x][4 + e] = ss[e]; }
      __syncthreads();
      if (tl == 0 && qd < cq) {
        for (int l = 1; l < lanes; ++l)
#pragma unroll
          for (int e = 0; e < 4; ++e) { s[e] += red[l * nq + tq][e]; ss[e] += red[l * nq + tq][4 + e]; }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          atomicAdd(&acc[((int64_t)b * C + 4 * qd + e) * 2], (double)s[e]);
          atomicAdd(&acc[((int64_t)b * C + 4 * qd + e) * 2 + 1], (double)ss[e]);
        }
      }
      __syncthreads();
    }
  } else {
    // rows of n2 contiguous values, one per (position, channel).  A row is read by the smallest power-of-two lane group
    // that covers it with 128-bit loads (deep layers have rows of 8-64 values: 2-16 lanes), so a warp sums 32/g rows at a
    // time; each row sum goes to its channel's fp64 accumulator.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const bool vec = (n2 & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    const int n4 = vec ? n2 >> 2 : n2;          // loads per row
    int g = 1;
    while (g < n4 && g < 32) g <<= 1;
    const int rpw = 32 / g, sub = lane / g, lj = lane % g;
    // layout 1: `chunk` counts ROWS (position-major, channel-minor) per block, so deep layers with a handful of positions
    // still spread over the machine
    const int64_t row_end = min(n1 * C, i0 + chunk);
    const int64_t per_pass = (int64_t)nw * rpw;
    for (int64_t r0 = i0; r0 < row_end; r0 += per_pass) {     // block-uniform trip count (full-mask shuffles below)
      const int64_t r = r0 + warp * rpw + sub;
      const bool ok = r < row_end;
      const int c = ok ? (int)(r % C) : 0;
      const int64_t i = ok ? r / C : 0;
      const float* p = x + (((int64_t)b * n1 + i) * C + c) * n2;
      float s = 0.f, ss = 0.f;
      if (ok) {
        if (vec) {
          const float4* p4 = reinterpret_cast<const float4*>(p);
          int j = lj;
          for (; j + g < n4; j += 2 * g) {
            const float4 v0 = __ldg(p4 + j), v1 = __ldg(p4 + j + g);
            s += (v0.x + v0.y) + (v0.z + v0.w) + (v1.x + v1.y) + (v1.z + v1.w);
            ss += fmaf(v0.x, v0.x, v0.y * v0.y) + fmaf(v0.z, v0.z, v0.w * v0.w) + fmaf(v1.x, v1.x, v1.y * v1.y) +
                  fmaf(v1.z, v1.z, v1.w * v1.w);
          }
          for (; j < n4; j += g) {
            const float4 v0 = __ldg(p4 + j);
            s += (v0.x + v0.y) + (v0.z + v0.w);
            ss += fmaf(v0.x, v0.x, v0.y * v0.y) + fmaf(v0.z, v0.z, v0.w * v0.w);
          }
        } else {
          for (int j = lj; j < n2; j += g) {
            const float v = p[j];
            s += v;
            ss = fmaf(v, v, ss);
          }
        }
      }
      for (int o = g >> 1; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
      }
      if (ok && lj == 0) {
        atomicAdd(&acc[((int64_t)b * C + c) * 2], (double)s);
        atomicAdd(&acc[((int64_t)b * C + c) * 2 + 1], (double)ss);
      }
    }
  }
}

__global__ void instnorm_finalize_kernel(const double* __restrict__ acc, int n, double inv_count, float eps,
                                         float2* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mean = acc[2 * i] * inv_count;
  double var = acc[2 * i + 1] * inv_count - mean * mean;  // biased variance, as nn.InstanceNorm2d
  if (var < 0) var = 0;
  stats[i] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
}

// ---- norm + affine + activation + bf16 split -------------------------------------------------------------
// mode 0: x[(b*n1+i)*ld + c]          -> planes[(b*n1+i)*ldp + c]             (channels-last, same layout)
// mode 2: x[((b*n1+i)*C + c)*n2 + j]  -> planes[((b*n1+i)*C + c)*ldp + j]     (channel-major rows, same layout)
__global__ void __launch_bounds__(256) norm_act_split_kernel(const float* __restrict__ x, int mode, int64_t rows, int C,
                                                             int n2, int64_t n1, int64_t ld, const float2* __restrict__ stats,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             int act, __nv_bfloat16* __restrict__ planes, int64_t ldp,
                                                             int64_t p_plane) {
  // 4 consecutive inner elements per thread
  const int inner = mode == 0 ? C : n2;
  const int64_t quads = (int64_t)(inner >> 2);
  const int64_t total = rows * quads;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / quads;
    const int q = (int)(idx - r * quads) * 4;
    float4 v;
    float o[4];
    if (mode == 0) {
      const int64_t b = r / n1;
      v = *reinterpret_cast<const float4*>(x + r * ld + q);
      float in[4] = {v.x, v.y, v.z, v.w};
      if (stats != nullptr) {   // (mean, rstd) pairs, gamma and beta of the quad's 4 channels: four 16-byte loads
        const float4 s01 = __ldg(reinterpret_cast<const float4*>(stats + b * C + q));
        const float4 s23 = __ldg(reinterpret_cast<const float4*>(stats + b * C + q + 2));
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + q));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + q));
        in[0] = (in[0] - s01.x) * s01.y * g4.x + b4.x;
        in[1] = (in[1] - s01.z) * s01.w * g4.y + b4.y;
        in[2] = (in[2] - s23.x) * s23.y * g4.z + b4.z;
        in[3] = (in[3] - s23.z) * s23.w * g4.w + b4.w;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = apply_act(in[e], act);
    } else {
      const int64_t bc = r;                 // row index = (b*n1 + i)*C + c
      const int c = (int)(bc % C);
      const int64_t b = bc / ((int64_t)C * n1);
      v = *reinterpret_cast<const float4*>(x + r * n2 + q);
      const float in[4] = {v.x, v.y, v.z, v.w};
      float2 st = make_float2(0.f, 1.f);
      float g = 1.f, be = 0.f;
      if (stats != nullptr) {
        st = stats[b * C + c];
        g = gamma[c];
        be = beta[c];
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = apply_act((in[e] - st.x) * st.y * g + be, act);
    }
    uint32_t h0, l0, h1, l1;
    tc::split_bf16x2(o[0], o[1], h0, l0);
    tc::split_bf16x2(o[2], o[3], h1, l1);
    __nv_bfloat16* pr = planes + r * ldp + q;
    *reinterpret_cast<uint2*>(pr) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(pr + p_plane) = make_uint2(l0, l1);
  }
}

// Mode 0 fast path (channels-last rows of C = 4 * quads channels, quads a divisor of 256): grid (slices, batch); a thread
// keeps ONE channel quad for the whole launch, so the (mean, rstd, gamma, beta) of its four channels live in registers and
// the row loop has no divisions; two rows in flight per thread.
__global__ void __launch_bounds__(256) norm_act_split_cl_kernel(const float* __restrict__ x, int64_t n1, int C, int64_t ld,
                                                                int64_t rows_per_block, const float2* __restrict__ stats,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                int act, __nv_bfloat16* __restrict__ planes, int64_t ldp,
                                                                int64_t p_plane) {
  const int quads = C >> 2;
  const int rpp = 256 / quads;                       // rows per pass
  const int tq = threadIdx.x % quads, tr = threadIdx.x / quads;
  const int b = blockIdx.y;
  const int q = 4 * tq;
  float mean[4] = {0.f, 0.f, 0.f, 0.f}, scale[4] = {1.f, 1.f, 1.f, 1.f}, shift[4] = {0.f, 0.f, 0.f, 0.f};
  if (stats != nullptr) {
    const float4 s01 = __ldg(reinterpret_cast<const float4*>(stats + (int64_t)b * C + q));
    const float4 s23 = __ldg(reinterpret_cast<const float4*>(stats + (int64_t)b * C + q + 2));
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + q));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + q));
    mean[0] = s01.x; mean[1] = s01.z; mean[2] = s23.x; mean[3] = s23.z;
    scale[0] = s01.y; scale[1] = s01.w; scale[2] = s23.y; scale[3] = s23.w;
    shift[0] = b4.x; shift[1] = b4.y; shift[2] = b4.z; shift[3] = b4.w;
    // keep the reference's evaluation order: ((x - mean) * rstd) * gamma + beta
    const float gm[4] = {g4.x, g4.y, g4.z, g4.w};
    auto finish = [&](const float4 v, int64_t r) {
      float in[4] = {v.x, v.y, v.z, v.w}, o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = apply_act((in[e] - mean[e]) * scale[e] * gm[e] + shift[e], act);
      uint32_t h0, l0, h1, l1;
      tc::split_bf16x2(o[0], o[1], h0, l0);
      tc::split_bf16x2(o[2], o[3], h1, l1);
      __nv_bfloat16* pr = planes + r * ldp + q;
      *reinterpret_cast<uint2*>(pr) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(pr + p_plane) = make_uint2(l0, l1);
    };
    const int64_t r_begin = (int64_t)b * n1 + (int64_t)blockIdx.x * rows_per_block;
    const int64_t r_end = min((int64_t)(b + 1) * n1, r_begin + rows_per_block);
    int64_t r = r_begin + tr;
    for (; r + rpp < r_end; r += 2 * rpp) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(x + r * ld + q));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(x + (r + rpp) * ld + q));
      finish(v0, r);
      finish(v1, r + rpp);
    }
    if (r < r_end) finish(__ldg(reinterpret_cast<const float4*>(x + r * ld + q)), r);
    return;
  }
  // no statistics: activation + split only
  const int64_t r_begin = (int64_t)b * n1 + (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min((int64_t)(b + 1) * n1, r_begin + rows_per_block);
  for (int64_t r = r_begin + tr; r < r_end; r += rpp) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ld + q));
    uint32_t h0, l0, h1, l1;
    tc::split_bf16x2(apply_act(v.x, act), apply_act(v.y, act), h0, l0);
    tc::split_bf16x2(apply_act(v.z, act), apply_act(v.w, act), h1, l1);
    __nv_bfloat16* pr = planes + r * ldp + q;
    *reinterpret_cast<uint2*>(pr) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(pr + p_plane) = make_uint2(l0, l1);
  }
}

// mode 1: channels-last x[(bt*F + f)*ld + c] -> channel-major planes[(bt*C + c)*ldp + f]   (TDF input, :117-119)
__global__ void __launch_bounds__(256) norm_act_split_tr_kernel(const float* __restrict__ x, int64_t BT, int F, int C,
                                                                int64_t T, int64_t ld, const float2* __restrict__ stats,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, int act,
                                                                __nv_bfloat16* __restrict__ planes, int64_t ldp,
                                                                int64_t p_plane) {
  // 32 channels x 64 frequencies per block: channel-contiguous 128-byte loads, frequency-contiguous bf16x2 stores
  // (128 bytes per warp and plane)
  __shared__ float tile[32][66];
  const int64_t bt = blockIdx.z;
  const int f0 = blockIdx.y * 64, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows per pass
  const int64_t b = bt / T;
  const int c = c0 + tx;
  float2 st = make_float2(0.f, 1.f);
  float ga = 1.f, be = 0.f;
  if (stats != nullptr && c < C) {
    st = stats[b * C + c];
    ga = gamma[c];
    be = beta[c];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int f = f0 + ty + 8 * k;
    float y = 0.f;
    if (f < F && c < C) {
      y = __ldg(x + (bt * F + f) * ld + c);
      if (stats != nullptr) y = (y - st.x) * st.y * ga + be;
      y = apply_act(y, act);
    }
    tile[tx][ty + 8 * k] = y;
  }
  __syncthreads();
  const bool pair_ok = (ldp & 1) == 0 && (p_plane & 1) == 0 && (reinterpret_cast<uintptr_t>(planes) & 3) == 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int cc = c0 + ty + 8 * k, f = f0 + 2 * tx;
    if (cc < C && f < F) {
      const float2 v = *reinterpret_cast<const float2*>(&tile[ty + 8 * k][2 * tx]);
      const int64_t o = (bt * C + cc) * ldp + f;
      if (pair_ok && f + 1 < F) {
        uint32_t h, l;
        tc::split_bf16x2(v.x, v.y, h, l);
        *reinterpret_cast<uint32_t*>(planes + o) = h;
        *reinterpret_cast<uint32_t*>(planes + o + p_plane) = l;
      } else {
        __nv_bfloat16 h, l;
        tc::split_bf16(v.x, h, l);
        planes[o] = h;
        planes[o + p_plane] = l;
        if (f + 1 < F) {
          tc::split_bf16(v.y, h, l);
          planes[o + 1] = h;
          planes[o + 1 + p_plane] = l;
        }
      }
    }
  }
}

// x[(bt*F + f)*ld + c] += g[(bt*C + c)*F + f]      ("x = x + tdf(x)", :134, with the TDF result channel-major)
__global__ void __launch_bounds__(256) transpose_add_kernel(float* __restrict__ x, const float* __restrict__ g, int F,
                                                            int C, int64_t ld) {
  __shared__ float tile[32][33];
  const int64_t bt = blockIdx.z;
  const int f0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, f = f0 + tx;
    tile[ty + 8 * k][tx] = (c < C && f < F) ? g[(bt * C + c) * F + f] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int f = f0 + ty + 8 * k, c = c0 + tx;
    if (f < F && c < C) x[(bt * F + f) * ld + c] += tile[tx][ty + 8 * k];
  }
}

// The same update, and the InstanceNorm statistics of its RESULT on the way (the next thing the reference does with
// x + tdf(x) is tfc2's norm, :135): a block walks `fs` consecutive 32-frequency tiles of one (bt, 32-channel) column, keeps
// per-thread partial sums of the new values (fp32 over <= 4*fs values), reduces them over the block and adds one fp64
// (sum, sum of squares) pair per channel to acc[b][c] — the accumulation scheme of instnorm_acc_kernel, without reading
// the tensor a second time.
__global__ void __launch_bounds__(256) transpose_add_stats_kernel(float* __restrict__ x, const float* __restrict__ g, int F,
                                                                  int C, int64_t ld, int64_t T, int fs,
                                                                  double* __restrict__ acc) {
  __shared__ float tile[32][33];
  __shared__ float red[2][8][32];
  const int64_t bt = blockIdx.z;
  const int c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float s = 0.f, ss = 0.f;
  for (int ft = 0; ft < fs; ++ft) {
    const int f0 = (blockIdx.y * fs + ft) * 32;
    if (f0 >= F) break;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + ty + 8 * k, f = f0 + tx;
      tile[ty + 8 * k][tx] = (c < C && f < F) ? g[(bt * C + c) * F + f] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int f = f0 + ty + 8 * k, c = c0 + tx;
      if (f < F && c < C) {
        float* xp = x + (bt * F + f) * ld + c;
        const float v = *xp + tile[tx][ty + 8 * k];
        *xp = v;
        s += v;
        ss = fmaf(v, v, ss);
      }
    }
    __syncthreads();
  }
  red[0][ty][tx] = s;
  red[1][ty][tx] = ss;
  __syncthreads();
  if (ty == 0 && c0 + tx < C) {
#pragma unroll
    for (int l = 1; l < 8; ++l) { s += red[0][l][tx]; ss += red[1][l][tx]; }
    const int64_t b = bt / T;
    atomicAdd(&acc[(b * C + c0 + tx) * 2], (double)s);
    atomicAdd(&acc[(b * C + c0 + tx) * 2 + 1], (double)ss);
  }
}

// spec (sesa_stft layout 0) [bt][f_full][c2] -> mix[bt][f'][c2*k + kk], f_full = kk*Fs + f'   (cac2cws, :191-196)
__global__ void mdx_pack_kernel(const float* __restrict__ spec, int64_t BT, int F_full, int Fs, int k, int c2,
                                float* __restrict__ mix) {
  const int ch = c2 * k;
  const int64_t total = BT * Fs * ch;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % ch);
    const int64_t r = i / ch;
    const int fs = (int)(r % Fs);
    const int64_t bt = r / Fs;
    const int co = cc / k, kk = cc - co * k;
    mix[i] = spec[(bt * F_full + (int64_t)kk * Fs + fs) * c2 + co];
  }
}

// planes[r][0:ch] = split(mix[r]), planes[r][ch:ch+C] = split(x[r] * first[r])   ("x * first_conv_out", cat(mix, x): :228-230)
__global__ void mdx_final_concat_kernel(const float* __restrict__ mix, int ch, const float* __restrict__ x, int64_t ldx,
                                        const float* __restrict__ first, int64_t ldf, int C, int64_t rows,
                                        __nv_bfloat16* __restrict__ planes, int64_t ldp, int64_t p_plane) {
  const int w = ch + C;
  const int64_t total = rows * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % w);
    const int64_t r = i / w;
    const float v = j < ch ? mix[r * ch + j] : x[r * ldx + (j - ch)] * first[r * ldf + (j - ch)];
    __nv_bfloat16 h, l;
    tc::split_bf16(v, h, l);
    planes[r * ldp + j] = h;
    planes[r * ldp + j + p_plane] = l;
  }
}

// y[bt][f'][n*ch + c2o*k + kk] -> out[(b*nt + n)][t][f_full][c2o], zero for f_full >= k*Fs   (cws2cac :198-203 + the
// zero padding of STFT.inverse :36-38), i.e. the layout sesa_mask_istft mode 2 reads.
__global__ void mdx_unpack_kernel(const float* __restrict__ y, int64_t B, int64_t T, int Fs, int k, int c2, int nt,
                                  int F_full, float* __restrict__ out) {
  const int64_t total = B * nt * T * F_full * c2;
  const int ch = c2 * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % c2);
    int64_t r = i / c2;
    const int ff = (int)(r % F_full);
    r /= F_full;
    const int64_t t = r % T;
    r /= T;
    const int n = (int)(r % nt);
    const int64_t b = r / nt;
    float v = 0.f;
    if (ff < k * Fs) {
      const int kk = ff / Fs, fs = ff - kk * Fs;
      v = y[((b * T + t) * Fs + fs) * (int64_t)(nt * ch) + n * ch + co * k + kk];
    }
    out[i] = v;
  }
}

}  // namespace

extern "C" int sesa_instnorm_stats(const float* x, int layout, int batch, int64_t n1, int channels, int n2, int64_t ld,
                                   double* scratch, float* stats, float eps, void* stream) {
  SESA_CHECK_ARG(layout == 0 || layout == 1, "sesa_instnorm_stats: layout must be 0 or 1");
  SESA_CHECK_ARG(batch > 0 && n1 > 0 && channels > 0 && n2 > 0, "sesa_instnorm_stats: empty tensor");
  SESA_CHECK_ARG(layout != 0 || ((channels & 3) == 0 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0),
                 "sesa_instnorm_stats: channels-last statistics need channels % 4 == 0 and 16-byte aligned rows");
  cudaStream_t st = (cudaStream_t)stream;
  SESA_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * (size_t)batch * channels, st));
  // work per block: enough blocks to cover the machine about four times whatever the layer's extent (deep layers have
  // few positions and many channels), bounded so that fp32 partial sums stay short
  // layout 0: positions per block; layout 1: rows (position x channel) per block
  const int64_t units = layout == 0 ? n1 : n1 * channels;
  int64_t chunk = ceil_div64(units * batch, 148 * 4);
  const int64_t cmin = layout == 0 ? 8 : 64, cmax = layout == 0 ? 256 : 1024;
  if (chunk < cmin) chunk = cmin;
  if (chunk > cmax) chunk = cmax;
  dim3 grid((unsigned)ceil_div64(units, chunk), batch);
  instnorm_acc_kernel<<<grid, 256, 0, st>>>(x, layout, batch, n1, channels, n2, ld, chunk, scratch);
  SESA_LAUNCH_CHECK();
  const int n = batch * channels;
  instnorm_finalize_kernel<<<(n + 255) / 256, 256, 0, st>>>(scratch, n, 1.0 / ((double)n1 * n2), eps,
                                                            reinterpret_cast<float2*>(stats));
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

extern "C" int sesa_norm_act_split(const float* x, int mode, int batch, int64_t n1, int channels, int n2, int64_t ld,
                                   const float* stats, const float* gamma, const float* beta, int act, void* planes,
                                   int64_t ldp, int64_t p_plane, void* stream) {
  SESA_CHECK_ARG(mode >= 0 && mode <= 2, "sesa_norm_act_split: mode must be 0, 1 or 2");
  SESA_CHECK_ARG(stats == nullptr || (gamma != nullptr && beta != nullptr), "sesa_norm_act_split: stats need gamma and beta");
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(planes);
  const float2* s2 = reinterpret_cast<const float2*>(stats);
  if (mode == 1) {
    // n1 = T frames, n2 = F: x channels-last [b][t][f][c] -> planes [b][t][c][f]
    const int64_t BT = (int64_t)batch * n1;
    SESA_CHECK_ARG(BT <= 65535, "sesa_norm_act_split: too many (b, t) slices for one launch");
    dim3 grid((channels + 31) / 32, (n2 + 63) / 64, (unsigned)BT);
    norm_act_split_tr_kernel<<<grid, 256, 0, st>>>(x, BT, n2, channels, n1, ld, s2, gamma, beta, act, pl, ldp, p_plane);
  } else {
    const int inner = mode == 0 ? channels : n2;
    SESA_CHECK_ARG((inner & 3) == 0 && (ldp & 3) == 0 && (p_plane & 3) == 0 && (mode != 0 || (ld & 3) == 0),
                   "sesa_norm_act_split: inner extent and strides must be multiples of 4");
    const int64_t rows = mode == 0 ? (int64_t)batch * n1 : (int64_t)batch * n1 * channels;
    const int cq = channels >> 2;
    if (mode == 0 && cq >= 1 && cq <= 256 && 256 % cq == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
        (stats == nullptr || ((reinterpret_cast<uintptr_t>(stats) & 15) == 0 && (reinterpret_cast<uintptr_t>(gamma) & 15) == 0 &&
                              (reinterpret_cast<uintptr_t>(beta) & 15) == 0)) && batch <= 65535) {
      // rows per block: about eight blocks per SM over the whole launch, a multiple of the rows a block covers per pass
      const int rpp = 256 / cq;
      int64_t rpb = ceil_div64(n1 * batch, 148 * 8);
      rpb = ceil_div64(rpb, 2 * rpp) * 2 * rpp;
      dim3 grid((unsigned)ceil_div64(n1, rpb), batch);
      norm_act_split_cl_kernel<<<grid, 256, 0, st>>>(x, n1, channels, ld, rpb, s2, gamma, beta, act, pl, ldp, p_plane);
      SESA_LAUNCH_CHECK();
      return SESA_OK;
    }
    const int64_t total = rows * (inner >> 2);
    norm_act_split_kernel<<<(unsigned)min((int64_t)148 * 16, ceil_div64(total, 256)), 256, 0, st>>>(
        x, mode, rows, channels, n2, n1, ld, s2, gamma, beta, act, pl, ldp, p_plane);
  }
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

extern "C" int sesa_transpose_add(float* x, const float* g, int64_t bt, int F, int channels, int64_t ld, void* stream) {
  if (bt <= 0) return SESA_OK;
  SESA_CHECK_ARG(bt <= 65535, "sesa_transpose_add: too many (b, t) slices for one launch");
  dim3 grid((channels + 31) / 32, (F + 31) / 32, (unsigned)bt);
  transpose_add_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, g, F, channels, ld);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

extern "C" int sesa_transpose_add_stats(float* x, const float* g, int batch, int64_t frames, int F, int channels, int64_t ld,
                                        double* scratch, float* stats, float eps, void* stream) {
  const int64_t bt = (int64_t)batch * frames;
  if (bt <= 0) return SESA_OK;
  SESA_CHECK_ARG(bt <= 65535, "sesa_transpose_add_stats: too many (b, t) slices for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  SESA_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * (size_t)batch * channels, st));
  const int ftiles = (F + 31) / 32;
  const int fs = ftiles >= 8 ? 8 : ftiles;       // frequency tiles per block: 8x fewer atomics than one tile per block
  dim3 grid((channels + 31) / 32, (ftiles + fs - 1) / fs, (unsigned)bt);
  transpose_add_stats_kernel<<<grid, 256, 0, st>>>(x, g, F, channels, ld, frames, fs, scratch);
  SESA_LAUNCH_CHECK();
  const int n = batch * channels;
  instnorm_finalize_kernel<<<(n + 255) / 256, 256, 0, st>>>(scratch, n, 1.0 / ((double)frames * F), eps,
                                                            reinterpret_cast<float2*>(stats));
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

extern "C" int sesa_mdx_pack(const float* spec, int64_t bt, int f_full, int fs, int k, int c2, float* mix, void* stream) {
  SESA_CHECK_ARG(k * fs <= f_full, "sesa_mdx_pack: dim_f exceeds the spectrogram");
  const int64_t total = bt * fs * c2 * k;
  if (total == 0) return SESA_OK;
  mdx_pack_kernel<<<(unsigned)min((int64_t)148 * 16, ceil_div64(total, 256)), 256, 0, (cudaStream_t)stream>>>(spec, bt, f_full,
                                                                                                           fs, k, c2, mix);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

extern "C" int sesa_mdx_final_concat(const float* mix, int ch, const float* x, int64_t ldx, const float* first, int64_t ldf,
                                     int channels, int64_t rows, void* planes, int64_t ldp, int64_t p_plane, void* stream) {
  const int64_t total = rows * (ch + channels);
  if (total == 0) return SESA_OK;
  mdx_final_concat_kernel<<<(unsigned)min((int64_t)148 * 16, ceil_div64(total, 256)), 256, 0, (cudaStream_t)stream>>>(
      mix, ch, x, ldx, first, ldf, channels, rows, reinterpret_cast<__nv_bfloat16*>(planes), ldp, p_plane);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

extern "C" int sesa_mdx_unpack(const float* y, int64_t batch, int64_t frames, int fs, int k, int c2, int nt, int f_full,
                               float* out, void* stream) {
  const int64_t total = batch * nt * frames * f_full * c2;
  if (total == 0) return SESA_OK;
  mdx_unpack_kernel<<<(unsigned)min((int64_t)148 * 16, ceil_div64(total, 256)), 256, 0, (cudaStream_t)stream>>>(
      y, batch, frames, fs, k, c2, nt, f_full, out);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}
