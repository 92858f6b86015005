// fp32 flash-style attention over strided sequences (exact-fp32 path; the tcgen05 version lives in
// attention_tc.cu).  Replaces Attend.forward (models/bs_roformer/attend.py:89-93,113-126) plus the sigmoid
// gating and head merge of Attention.forward (models/bs_roformer/bs_roformer.py:115-120).
//
// The residual stream stays in ONE token-major layout [(b t f), d]; the axial "rearranges" of
// bs_roformer.py:526-543 become a stride choice: a sequence is (base row, row stride).
//   time attention: sequence (b,f): rows (b*T + t)*F + f, t = 0..T-1  -> stride F
//   band attention: sequence (b,t): rows (b*T + t)*F + f, f = 0..F-1  -> stride 1
// qkv row layout: [q(h d) | k(h d) | v(h d) | gate logits(h)], q pre-scaled by dh^-0.5 and q,k already
// rotated (both fused into the producing GEMM's epilogue).
#include "common.cuh"
#include "sesa_b200.h"

#define DH 64
#define BKV 64

template <int RPT>  // rows per thread: q tile = 16*RPT rows, 128 threads as 16 (row groups) x 8 (col groups)
__global__ void __launch_bounds__(128) attention_simt_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                             int ld, int ldo, int heads, int seq_len,
                                                             int inner_cnt, int64_t outer_stride,
                                                             int64_t inner_stride, int64_t pos_stride,
                                                             int q_tiles) {
  constexpr int BQ = 16 * RPT;
  extern __shared__ __align__(16) float smem_att[];
  float* Qt = smem_att;                 // [DH][BQ]   (d-major)
  float* Kt = Qt + DH * BQ;             // [DH][BKV]
  float* Vs = Kt + DH * BKV;            // [BKV][DH]
  float* Pt = Vs + BKV * DH;            // [BKV][BQ]  (key-major)
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 7, ty = tid >> 3;
  const int seq = blockIdx.x / q_tiles;
  const int qt = blockIdx.x % q_tiles;
  const int h = blockIdx.y;
  const int64_t base = (int64_t)(seq / inner_cnt) * outer_stride + (int64_t)(seq % inner_cnt) * inner_stride;
  const int inner = heads * DH;
  const int q0 = qt * BQ;

  // load Q tile transposed: lane = row within a 32-row slab, each thread moves one float4 of d
  for (int it = warp; it < (BQ / 32) * 16; it += 4) {
    const int r = (it / 16) * 32 + lane;
    const int d4 = it % 16;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < seq_len)
      v = *reinterpret_cast<const float4*>(qkv + (base + (int64_t)(q0 + r) * pos_stride) * ld + h * DH + d4 * 4);
    Qt[(d4 * 4 + 0) * BQ + r] = v.x; Qt[(d4 * 4 + 1) * BQ + r] = v.y;
    Qt[(d4 * 4 + 2) * BQ + r] = v.z; Qt[(d4 * 4 + 3) * BQ + r] = v.w;
  }

  float o[RPT][8];
  float mrow[RPT], lrow[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    mrow[i] = -INFINITY; lrow[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[i][j] = 0.f;
  }

  for (int k0 = 0; k0 < seq_len; k0 += BKV) {
    __syncthreads();  // previous tile fully consumed (and Q stores visible on the first pass)
    for (int it = warp; it < (BKV / 32) * 16; it += 4) {
      const int r = (it / 16) * 32 + lane;
      const int d4 = it % 16;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + r < seq_len) {
        const float* rowp = qkv + (base + (int64_t)(k0 + r) * pos_stride) * ld + h * DH + d4 * 4;
        kv = *reinterpret_cast<const float4*>(rowp + inner);
        vv = *reinterpret_cast<const float4*>(rowp + 2 * inner);
      }
      Kt[(d4 * 4 + 0) * BKV + r] = kv.x; Kt[(d4 * 4 + 1) * BKV + r] = kv.y;
      Kt[(d4 * 4 + 2) * BKV + r] = kv.z; Kt[(d4 * 4 + 3) * BKV + r] = kv.w;
      *reinterpret_cast<float4*>(Vs + r * DH + d4 * 4) = vv;
    }
    __syncthreads();

    // S = Q K^T : thread owns rows ty*RPT..+RPT, keys {tx*4..+3} U {32+tx*4..+3}
    float s[RPT][8];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < DH; ++d) {
      float a[RPT];
#pragma unroll
      for (int i4 = 0; i4 < RPT / 4; ++i4) {
        const float4 t4 = *reinterpret_cast<const float4*>(Qt + d * BQ + ty * RPT + i4 * 4);
        a[i4 * 4 + 0] = t4.x; a[i4 * 4 + 1] = t4.y; a[i4 * 4 + 2] = t4.z; a[i4 * 4 + 3] = t4.w;
      }
      const float4 b0 = *reinterpret_cast<const float4*>(Kt + d * BKV + tx * 4);
      const float4 b1 = *reinterpret_cast<const float4*>(Kt + d * BKV + 32 + tx * 4);
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s[i][j] = fmaf(a[i], b[j], s[i][j]);
    }
    // mask keys beyond the sequence, online softmax
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = k0 + (j < 4 ? tx * 4 + j : 32 + tx * 4 + (j - 4));
      if (key >= seq_len) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) s[i][j] = -INFINITY;
      }
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      float mx = s[i][0];
#pragma unroll
      for (int j = 1; j < 8; ++j) mx = fmaxf(mx, s[i][j]);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
      const float mnew = fmaxf(mrow[i], mx);   // finite: every tile holds at least one valid key
      const float corr = expf(mrow[i] - mnew);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[i][j] = expf(s[i][j] - mnew);
        sum += s[i][j];
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      lrow[i] = lrow[i] * corr + sum;
      mrow[i] = mnew;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[i][j] *= corr;
    }
    // P -> shared (key-major) so the PV product reads row quads
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = (j < 4 ? tx * 4 + j : 32 + tx * 4 + (j - 4));
#pragma unroll
      for (int i4 = 0; i4 < RPT / 4; ++i4)
        *reinterpret_cast<float4*>(Pt + key * BQ + ty * RPT + i4 * 4) =
            make_float4(s[i4 * 4 + 0][j], s[i4 * 4 + 1][j], s[i4 * 4 + 2][j], s[i4 * 4 + 3][j]);
    }
    __syncthreads();
    // O += P V : thread owns rows ty*RPT..+RPT, dims {tx*4..+3} U {32+tx*4..+3}
#pragma unroll 4
    for (int j = 0; j < BKV; ++j) {
      float p[RPT];
#pragma unroll
      for (int i4 = 0; i4 < RPT / 4; ++i4) {
        const float4 t4 = *reinterpret_cast<const float4*>(Pt + j * BQ + ty * RPT + i4 * 4);
        p[i4 * 4 + 0] = t4.x; p[i4 * 4 + 1] = t4.y; p[i4 * 4 + 2] = t4.z; p[i4 * 4 + 3] = t4.w;
      }
      const float4 v0 = *reinterpret_cast<const float4*>(Vs + j * DH + tx * 4);
      const float4 v1 = *reinterpret_cast<const float4*>(Vs + j * DH + 32 + tx * 4);
      const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) o[i][jj] = fmaf(p[i], v[jj], o[i][jj]);
    }
  }

  // normalise, gate with sigmoid(gate logit), merge heads
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int r = q0 + ty * RPT + i;
    if (r >= seq_len) continue;
    const int64_t row = base + (int64_t)r * pos_stride;
    const float gl = qkv[row * ld + 3 * inner + h];
    const float sc = (1.0f / (1.0f + expf(-gl))) / lrow[i];
    float* op = out + row * ldo + h * DH;
    *reinterpret_cast<float4*>(op + tx * 4) = make_float4(o[i][0] * sc, o[i][1] * sc, o[i][2] * sc, o[i][3] * sc);
    *reinterpret_cast<float4*>(op + 32 + tx * 4) = make_float4(o[i][4] * sc, o[i][5] * sc, o[i][6] * sc, o[i][7] * sc);
  }
}

template <int RPT>
static int launch_attention(const float* qkv, float* out, int ld, int ldo, int heads, int n_seq, int seq_len,
                            int inner_cnt, int64_t outer_stride, int64_t inner_stride, int64_t pos_stride,
                            cudaStream_t stream) {
  constexpr int BQ = 16 * RPT;
  const int q_tiles = (seq_len + BQ - 1) / BQ;
  const size_t smem = sizeof(float) * (size_t)(DH * BQ + DH * BKV + BKV * DH + BKV * BQ);
  SESA_CUDA(cudaFuncSetAttribute(attention_simt_kernel<RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((int64_t)n_seq * q_tiles), heads);
  attention_simt_kernel<RPT><<<grid, 128, smem, stream>>>(qkv, out, ld, ldo, heads, seq_len, inner_cnt,
                                                          outer_stride, inner_stride, pos_stride, q_tiles);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

extern "C" int sesa_attention_simt(const float* qkv, float* out, int ld, int ldo, int heads, int dim_head,
                                   int n_seq, int seq_len, int inner_cnt, int64_t outer_stride,
                                   int64_t inner_stride, int64_t pos_stride, void* stream) {
  SESA_CHECK_ARG(dim_head == DH, "sesa_attention: dim_head must be 64, got %d", dim_head);
  SESA_CHECK_ARG((ld & 3) == 0 && (ldo & 3) == 0, "sesa_attention: row strides must be multiples of 4 floats");
  SESA_CHECK_ARG(inner_cnt > 0 && seq_len > 0, "sesa_attention: bad sequence geometry");
  if (n_seq == 0) return SESA_OK;
  if (seq_len <= 64)
    return launch_attention<4>(qkv, out, ld, ldo, heads, n_seq, seq_len, inner_cnt, outer_stride,
                               inner_stride, pos_stride, (cudaStream_t)stream);
  return launch_attention<8>(qkv, out, ld, ldo, heads, n_seq, seq_len, inner_cnt, outer_stride, inner_stride,
                             pos_stride, (cudaStream_t)stream);
}
