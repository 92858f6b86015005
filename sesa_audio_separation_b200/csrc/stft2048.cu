// Register-resident 2048-point FFT kernels for the STFT front end and the fused mask + iSTFT output stage
// (n_fft = 2048 is what every RoFormer config of the reference uses: bs_roformer.py:343, mel_band_roformer.py:341).
//
// The generic kernels in stft.cu run six radix-4 Stockham passes through shared memory with a CTA-wide barrier after
// each; that makes them shared-memory- and barrier-bound (0.6-0.9 TB/s of HBM traffic).  Here one frame is owned by 128
// threads that hold 16 complex points each and the transform is 16 x 16 x 8:
//   forward  : radix-16 (registers) -> smem exchange -> radix-16 -> smem exchange -> radix-8, after which thread t holds
//              Z[q + 256 k] for q = t and q = 256 - t, i.e. BOTH members Z[f], Z[N - f] of every bin pair it needs to
//              separate the two real channels packed as z = xL + i xR: the spectrum goes straight from registers to HBM;
//   inverse  : the transposed network (radix-8 -> exchange -> radix-16 -> exchange -> radix-16), so the masked spectrum is
//              read exactly once, straight into the registers of the thread that needs bins f and N - f, and the time
//              samples come out in natural order for the windowed overlap-add.
// Two shared-memory exchanges per frame instead of six passes, two barriers instead of seven, twiddles and window in
// registers for the whole CTA lifetime.
#include "common.cuh"
#include "sesa_b200.h"

namespace {

constexpr int FN = 2048;      // transform length
constexpr int FT = 128;       // threads per frame
constexpr int FPAD = FN + FN / 16;   // exchange buffer with one pad element per 16

__device__ __forceinline__ int padi(int i) { return i + (i >> 4); }

constexpr float kC1 = 0.92387953251128674f;   // cos(pi/8)
constexpr float kS1 = 0.38268343236508977f;   // sin(pi/8)
constexpr float kC2 = 0.70710678118654752f;   // cos(pi/4)

// multiply by (wr, wi) for the forward transform, by its conjugate for the inverse
template <bool INV>
__device__ __forceinline__ float2 cmulw(float2 a, float wr, float wi) {
  return INV ? make_float2(a.x * wr + a.y * wi, a.y * wr - a.x * wi) : make_float2(a.x * wr - a.y * wi, a.y * wr + a.x * wi);
}

// complex add / subtract as ONE packed fp32x2 instruction each (FADD2 / FFMA2 with a (-1, -1) multiplier: b * -1 is exact,
// so a - b is rounded once, exactly like a scalar subtraction): the butterflies are what bounds these kernels (instruction
// issue, not bytes), and two thirds of their arithmetic is complex additions
__device__ __forceinline__ float2 padd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 psub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }

template <bool INV>
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
  const float2 apc = padd(a, c), amc = psub(a, c), bpd = padd(b, d), bmd = psub(b, d);
  a = padd(apc, bpd);
  c = psub(apc, bpd);
  // b, d = amc +/- (-/+ i) bmd: component-swapped operands, cheaper as four scalar additions than as packed ones
  if (INV) {
    b = make_float2(amc.x - bmd.y, amc.y + bmd.x);
    d = make_float2(amc.x + bmd.y, amc.y - bmd.x);
  } else {
    b = make_float2(amc.x + bmd.y, amc.y - bmd.x);
    d = make_float2(amc.x - bmd.y, amc.y + bmd.x);
  }
}

// 16-point DFT in registers.  In: v[m] natural order.  Out: X[k] is left in v[4 * (k & 3) + (k >> 2)].
#define O16(k) (4 * ((k) & 3) + ((k) >> 2))
template <bool INV>
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
#pragma unroll
  for (int m0 = 0; m0 < 4; ++m0) dft4<INV>(v[m0], v[4 + m0], v[8 + m0], v[12 + m0]);
  // v[4 * k0 + m0] *= W16^(m0 * k0), forward W16^j = (cos(pi j / 8), -sin(pi j / 8))
  v[4 + 1] = cmulw<INV>(v[4 + 1], kC1, -kS1);    // j = 1
  v[4 + 2] = cmulw<INV>(v[4 + 2], kC2, -kC2);    // j = 2
  v[4 + 3] = cmulw<INV>(v[4 + 3], kS1, -kC1);    // j = 3
  v[8 + 1] = cmulw<INV>(v[8 + 1], kC2, -kC2);    // j = 2
  v[8 + 2] = INV ? make_float2(-v[8 + 2].y, v[8 + 2].x) : make_float2(v[8 + 2].y, -v[8 + 2].x);   // j = 4: -/+ i
  v[8 + 3] = cmulw<INV>(v[8 + 3], -kC2, -kC2);   // j = 6
  v[12 + 1] = cmulw<INV>(v[12 + 1], kS1, -kC1);  // j = 3
  v[12 + 2] = cmulw<INV>(v[12 + 2], -kC2, -kC2); // j = 6
  v[12 + 3] = cmulw<INV>(v[12 + 3], -kC1, kS1);  // j = 9
#pragma unroll
  for (int k0 = 0; k0 < 4; ++k0) dft4<INV>(v[4 * k0], v[4 * k0 + 1], v[4 * k0 + 2], v[4 * k0 + 3]);
}

// 8-point DFT in registers.  In: v[m] natural order.  Out: X[k] is left in v[2 * (k & 3) + (k >> 2)].
#define O8(k) (2 * ((k) & 3) + ((k) >> 2))
template <bool INV>
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  dft4<INV>(v[0], v[2], v[4], v[6]);   // E[k0] -> v[2 k0]
  dft4<INV>(v[1], v[3], v[5], v[7]);   // O[k0] -> v[2 k0 + 1]
  v[3] = cmulw<INV>(v[3], kC2, -kC2);
  v[5] = INV ? make_float2(-v[5].y, v[5].x) : make_float2(v[5].y, -v[5].x);
  v[7] = cmulw<INV>(v[7], -kC2, -kC2);
#pragma unroll
  for (int k0 = 0; k0 < 4; ++k0) {
    const float2 e = v[2 * k0], o = v[2 * k0 + 1];
    v[2 * k0] = padd(e, o);
    v[2 * k0 + 1] = psub(e, o);
  }
}

// the two residues q (mod 256) whose radix-8 columns a thread owns in the last forward / first inverse pass
__device__ __forceinline__ void pair_residues(int tid, int& qa, int& qb) {
  qa = tid;
  qb = tid == 0 ? 128 : 256 - tid;
}

// ---------------------------------------------------------------------------------------------------------------------
// STFT: grid (frame groups, signals); a CTA transforms `fpb` consecutive frames of one signal group.
// spec[(ns * T + t)][f][c][re/im]  (layout 0 of sesa_stft)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FT, 3) stft2048_kernel(const float* __restrict__ audio, float* __restrict__ spec,
                                                         const float* __restrict__ window, const float2* __restrict__ tw,
                                                         int C, int64_t L, int hop, int T, int fpb) {
  extern __shared__ float2 smem2k[];
  float2* y1 = smem2k;
  float2* y2 = y1 + FPAD;
  float2* tw1s = y2 + FPAD;        // [FT][17]: W_2048^(tid k), row per thread (padded like y1); in shared memory rather than
  float2* tw2s = tw1s + FT * 17;   // [8][16]:  W_128^(p k)      registers so that four CTAs fit on an SM
  const int tid = threadIdx.x;
  const int ns = blockIdx.y;
  const int p = tid >> 4, q = tid & 15;
#pragma unroll
  for (int k = 1; k < 16; ++k) tw1s[17 * tid + k] = __ldg(tw + tid * k);
  tw2s[tid] = __ldg(tw + 16 * (tid >> 4) * (tid & 15));
  float win[16];
#pragma unroll
  for (int m = 0; m < 16; ++m) win[m] = __ldg(window + tid + FT * m);
  __syncthreads();
  int qa, qb;
  pair_residues(tid, qa, qb);
  const float* a0 = audio + (int64_t)ns * C * L;
  const float* a1 = a0 + L;
  const int t_end = min(T, (int)(blockIdx.x + 1) * fpb);
  // the samples of frame t + 1 are requested while frame t is transformed (one CTA has only four warps to hide the
  // latency of its own loads)
  float2 raw[16];
  auto load_frame = [&](int t) {
    const int64_t base = (int64_t)t * hop - FN / 2 + tid;
    if (base - tid >= 0 && base - tid + FN <= L) {
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const int64_t j = base + FT * m;
        raw[m].x = __ldg(a0 + j);
        raw[m].y = C == 2 ? __ldg(a1 + j) : 0.f;
      }
    } else {
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const int64_t j = reflect_index(base + FT * m, L);
        raw[m].x = a0[j];
        raw[m].y = C == 2 ? a1[j] : 0.f;
      }
    }
  };
  const int t_begin = blockIdx.x * fpb;
  if (t_begin < t_end) load_frame(t_begin);
  for (int t = t_begin; t < t_end; ++t) {
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = make_float2(raw[m].x * win[m], raw[m].y * win[m]);
    if (t + 1 < t_end) load_frame(t + 1);
    // pass 1: radix 16, stride 1
    dft16<false>(v);
    y1[17 * tid] = v[0];
#pragma unroll
    for (int k = 1; k < 16; ++k) y1[17 * tid + k] = cmul(v[O16(k)], tw1s[17 * tid + k]);
    __syncthreads();
    // pass 2: radix 16, stride 16
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = y1[padi(tid + FT * m)];
    dft16<false>(v);
    y2[padi(q + 256 * p)] = v[0];
#pragma unroll
    for (int k = 1; k < 16; ++k) y2[padi(q + 256 * p + 16 * k)] = cmul(v[O16(k)], tw2s[16 * p + k]);
    __syncthreads();
    // pass 3: radix 8, stride 256, on the residues qa and qb = -qa (mod 256)
    float2 a[8], b[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      a[m] = y2[padi(qa + 256 * m)];
      b[m] = y2[padi(qb + 256 * m)];
    }
    dft8<false>(a);
    dft8<false>(b);
    // a[O8(k)] = Z[qa + 256 k], b[O8(k)] = Z[qb + 256 k]
    const int64_t row = (int64_t)ns * T + t;
    if (C == 2) {
      float4* out = reinterpret_cast<float4*>(spec) + row * (FN / 2 + 1);
      // XL = (Z[f] + conj(Z[N-f])) / 2 ; XR = (Z[f] - conj(Z[N-f])) / (2i)
      auto emit = [&](int f, float2 zk, float2 zn) {
        out[f] = make_float4(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y), 0.5f * (zk.y + zn.y), 0.5f * (zn.x - zk.x));
      };
      if (tid != 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          emit(qa + 256 * k, a[O8(k)], b[O8(7 - k)]);
          emit(qb + 256 * k, b[O8(k)], a[O8(7 - k)]);
        }
      } else {   // residues 0 and 128 are their own mirrors
        emit(0, a[O8(0)], a[O8(0)]);
        emit(256, a[O8(1)], a[O8(7)]);
        emit(512, a[O8(2)], a[O8(6)]);
        emit(768, a[O8(3)], a[O8(5)]);
        emit(1024, a[O8(4)], a[O8(4)]);
#pragma unroll
        for (int k = 0; k < 4; ++k) emit(128 + 256 * k, b[O8(k)], b[O8(7 - k)]);
      }
    } else {
      float2* out = reinterpret_cast<float2*>(spec) + row * (FN / 2 + 1);
      if (tid != 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          out[qa + 256 * k] = a[O8(k)];
          out[qb + 256 * k] = b[O8(k)];
        }
      } else {
#pragma unroll
        for (int k = 0; k < 5; ++k) out[256 * k] = a[O8(k)];
#pragma unroll
        for (int k = 0; k < 4; ++k) out[128 + 256 * k] = b[O8(k)];
      }
    }
    // the next frame's first exchange writes y1, whose readers all passed the second barrier; its second exchange
    // writes y2 after the next first barrier, which every reader of y2 above reaches first
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Fused complex mask multiply + iSTFT (+ synthesis window, overlap-add over frames, / window envelope, trim).
// One CTA (128 threads) inverse-transforms `G` consecutive frames of one (chunk, stem) — every frame exactly once — and
// overlap-adds them in a circular shared-memory accumulator of 2 x n_fft samples per channel.  The `hop` samples that no
// later frame of the CTA touches are flushed after each frame: divided by the window envelope and added to the
// zero-initialised output with red.global.add.  An output sample receives contributions from at most two CTAs (G is at
// least (n_fft - hop) / hop), and 0 + a + b == 0 + b + a exactly, so the result does not depend on the order in which
// the two arrive: the output is bitwise reproducible.  Layouts as in stft.cu.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int ACC = 2 * FN;   // circular accumulator length per channel

template <int MODE, int C>
__global__ void __launch_bounds__(FT, 3) mask_istft2048_kernel(
    const float* __restrict__ spec, const float* __restrict__ mask, const int* __restrict__ inv,
    const float* __restrict__ cnt, float* __restrict__ out, const float* __restrict__ window,
    const float* __restrict__ env, const float2* __restrict__ tw, int hop, int T, int64_t out_len, int G, int nstems,
    int J) {
  extern __shared__ float2 smem2k[];
  float2* y1 = smem2k;
  float2* y2 = smem2k + FPAD;
  float* acc = reinterpret_cast<float*>(smem2k + 2 * FPAD);   // [C][ACC]
  constexpr int F = FN / 2 + 1;
  const int tid = threadIdx.x;
  const int n = blockIdx.y, b = blockIdx.z;
  const int nb = gridDim.z;
  const int p = tid >> 4, q = tid & 15;
  float2 tw1[16], tw2[16];
  float win[16];
#pragma unroll
  for (int k = 1; k < 16; ++k) {
    tw1[k] = __ldg(tw + tid * k);
    tw2[k] = __ldg(tw + 16 * p * k);
  }
#pragma unroll
  for (int m = 0; m < 16; ++m) win[m] = __ldg(window + tid + FT * m) * (1.0f / (float)FN);
  int qa, qb;
  pair_residues(tid, qa, qb);
  for (int i = tid; i < C * ACC; i += FT) acc[i] = 0.f;
  const int t_lo = blockIdx.x * G;
  const int t_hi = min(T, t_lo + G) - 1;
  float* outp = out + ((int64_t)b * nstems + n) * C * out_len;
  // flush padded positions [P0, P1): out[P - n_fft/2] += acc / envelope, and clear them for reuse
  auto flush = [&](int64_t P0, int64_t P1) {
    for (int64_t P = P0 + tid; P < P1; P += FT) {
      const int64_t io = P - FN / 2;
      const int a = (int)(P & (ACC - 1));
      if (io >= 0 && io < out_len) {
        const float e = __ldg(env + io);
        atomicAdd(outp + io, acc[a] / e);
        if (C == 2) atomicAdd(outp + out_len + io, acc[ACC + a] / e);
      }
      acc[a] = 0.f;
      if (C == 2) acc[ACC + a] = 0.f;
    }
  };

  // masked spectrum of bin f for both channels: (YL, YR)
  auto load_bin = [&](int64_t row, int f, float2& yl, float2& yr) {
    yr = make_float2(0.f, 0.f);
    if (C == 2) {
      float4 sp;
      if (MODE == 2) sp = __ldg(reinterpret_cast<const float4*>(spec) + (((int64_t)b * nstems + n) * T + (row - (int64_t)b * T)) * F + f);
      else sp = __ldg(reinterpret_cast<const float4*>(spec) + row * F + f);
      if (MODE == 0) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(mask) + ((int64_t)n * nb * T + row) * F + f);
        yl = cmul(make_float2(sp.x, sp.y), make_float2(m.x, m.y));
        yr = cmul(make_float2(sp.z, sp.w), make_float2(m.z, m.w));
      } else if (MODE == 1) {
        const float2* mrow = reinterpret_cast<const float2*>(mask) + ((int64_t)n * nb * T + row) * J;
        float2 ml = make_float2(0.f, 0.f), mr = make_float2(0.f, 0.f);
        const int4 iv = __ldg(reinterpret_cast<const int4*>(inv) + f);
        if (iv.x >= 0) ml = cadd(ml, __ldg(mrow + iv.x));
        if (iv.y >= 0) ml = cadd(ml, __ldg(mrow + iv.y));
        if (iv.z >= 0) mr = cadd(mr, __ldg(mrow + iv.z));
        if (iv.w >= 0) mr = cadd(mr, __ldg(mrow + iv.w));
        const float2 cc = __ldg(reinterpret_cast<const float2*>(cnt) + f);
        ml.x /= cc.x; ml.y /= cc.x; mr.x /= cc.y; mr.y /= cc.y;
        yl = cmul(make_float2(sp.x, sp.y), ml);
        yr = cmul(make_float2(sp.z, sp.w), mr);
      } else {
        yl = make_float2(sp.x, sp.y);
        yr = make_float2(sp.z, sp.w);
      }
    } else {
      float2 sp;
      if (MODE == 2) sp = __ldg(reinterpret_cast<const float2*>(spec) + (((int64_t)b * nstems + n) * T + (row - (int64_t)b * T)) * F + f);
      else sp = __ldg(reinterpret_cast<const float2*>(spec) + row * F + f);
      if (MODE == 0) {
        yl = cmul(sp, __ldg(reinterpret_cast<const float2*>(mask) + ((int64_t)n * nb * T + row) * F + f));
      } else if (MODE == 1) {
        const float2* mrow = reinterpret_cast<const float2*>(mask) + ((int64_t)n * nb * T + row) * J;
        float2 ml = make_float2(0.f, 0.f);
        const int2 iv = __ldg(reinterpret_cast<const int2*>(inv) + f);
        if (iv.x >= 0) ml = cadd(ml, __ldg(mrow + iv.x));
        if (iv.y >= 0) ml = cadd(ml, __ldg(mrow + iv.y));
        const float cl = __ldg(cnt + f);
        ml.x /= cl; ml.y /= cl;
        yl = cmul(sp, ml);
      } else {
        yl = sp;
      }
    }
    if (f == 0 || f == FN / 2) {   // c2r ignores the imaginary part of DC and Nyquist
      yl.y = 0.f;
      yr.y = 0.f;
    }
  };
  // Z[f] = YL + i YR ; Z[N - f] = conj(YL) + i conj(YR)
  auto z_lo = [](float2 yl, float2 yr) { return make_float2(yl.x - yr.y, yl.y + yr.x); };
  auto z_hi = [](float2 yl, float2 yr) { return make_float2(yl.x + yr.y, yr.x - yl.y); };

  // 128-byte lines of one frame's spectrum row and mask row (for the L2 prefetch of the next frame below)
  const int spec_lines = (F * C * 8 + 127) / 128;
  const int mask_lines = MODE == 0 ? spec_lines : MODE == 1 ? (J * 8 + 127) / 128 : 0;
  for (int t = t_lo; t <= t_hi; ++t) {
    const int64_t row = (int64_t)b * T + t;
    if (t < t_hi) {
      // the NEXT frame's rows are requested into L2 now (no registers held), so that its 16 loads per thread, issued
      // right after this frame's transform, pay the L2 latency instead of the HBM latency: the CTA has only four warps
      // and nothing else to run while a frame's operands are in flight
      const char* sp_next = reinterpret_cast<const char*>(
          spec + (MODE == 2 ? (((int64_t)b * nstems + n) * T + (t + 1)) : (row + 1)) * (int64_t)F * C * 2);
      for (int l = tid; l < spec_lines; l += FT) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp_next + (int64_t)l * 128));
      if (MODE != 2) {
        const char* mk_next = reinterpret_cast<const char*>(
            mask + ((int64_t)n * nb * T + row + 1) * (MODE == 0 ? (int64_t)F * C * 2 : (int64_t)J * 2));
        for (int l = tid; l < mask_lines; l += FT) asm volatile("prefetch.global.L2 [%0];" ::"l"(mk_next + (int64_t)l * 128));
      }
    }
    // a[k] = Z[qa + 256 k], b[k] = Z[qb + 256 k], k = 0..7, built from bins f <= 1024 only
    float2 a[8], bb[8];
    if (tid != 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 yl, yr;
        load_bin(row, qa + 256 * k, yl, yr);      // f = qa + 256 k ; N - f = qb + 256 (7 - k)
        a[k] = z_lo(yl, yr);
        bb[7 - k] = z_hi(yl, yr);
        load_bin(row, qb + 256 * k, yl, yr);      // f = qb + 256 k ; N - f = qa + 256 (7 - k)
        bb[k] = z_lo(yl, yr);
        a[7 - k] = z_hi(yl, yr);
      }
    } else {
      float2 yl, yr;
      load_bin(row, 0, yl, yr);
      a[0] = z_lo(yl, yr);
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        load_bin(row, 256 * k, yl, yr);
        a[k] = z_lo(yl, yr);
        a[8 - k] = z_hi(yl, yr);
      }
      load_bin(row, 1024, yl, yr);
      a[4] = z_lo(yl, yr);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        load_bin(row, 128 + 256 * k, yl, yr);
        bb[k] = z_lo(yl, yr);
        bb[7 - k] = z_hi(yl, yr);
      }
    }
    // transposed pass 3: radix 8 over k, scattered to where the forward pass gathered from
    dft8<true>(a);
    dft8<true>(bb);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      y2[padi(qa + 256 * m)] = a[O8(m)];
      y2[padi(qb + 256 * m)] = bb[O8(m)];
    }
    __syncthreads();
    // the previous frame's hop is final for this CTA (its accumulation precedes the barrier above, no later frame reaches
    // back before t * hop)
    if (t > t_lo) flush((int64_t)(t - 1) * hop, (int64_t)t * hop);
    // transposed pass 2: gather from the forward scatter positions, conj twiddle, radix 16, scatter to tid + 128 m
    float2 v[16];
    v[0] = y2[padi(q + 256 * p)];
#pragma unroll
    for (int k = 1; k < 16; ++k) {
      const float2 x = y2[padi(q + 256 * p + 16 * k)];
      v[k] = cmulw<true>(x, tw2[k].x, tw2[k].y);
    }
    dft16<true>(v);
#pragma unroll
    for (int m = 0; m < 16; ++m) y1[padi(tid + FT * m)] = v[O16(m)];
    __syncthreads();
    // transposed pass 1
    v[0] = y1[17 * tid];
#pragma unroll
    for (int k = 1; k < 16; ++k) {
      const float2 x = y1[17 * tid + k];
      v[k] = cmulw<true>(x, tw1[k].x, tw1[k].y);
    }
    dft16<true>(v);
    // v[O16(m)] = z[tid + 128 m]: windowed overlap-add into the circular accumulator
    const int fq = (int)(((int64_t)t * hop) & (ACC - 1));
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int a = (fq + tid + FT * m) & (ACC - 1);
      const float2 z = v[O16(m)];
      acc[a] += z.x * win[m];
      if (C == 2) acc[ACC + a] += z.y * win[m];
    }
    // no barrier here: the next frame's y2 scatter follows this frame's second barrier (every y2 gather precedes it), its
    // y1 scatter, its flush and its accumulation follow its own barriers, which every thread reaches only after this frame
  }
  __syncthreads();
  if (t_hi >= t_lo) flush((int64_t)t_hi * hop, (int64_t)t_hi * hop + FN);
}

}  // namespace

int sesa_launch_stft2048(const float* audio, float* spec, const float* window, const float* twiddle, int n_signals,
                         int channels, int64_t length, int hop, int T, cudaStream_t stream) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    SESA_CUDA(cudaGetDevice(&dev));
    SESA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  // enough CTAs to fill every SM three times over (the residency limit), at most 16 frames per CTA (consecutive frames of a CTA re-read
  // most of their samples from L1)
  int fpb = (int)ceil_div64((int64_t)T * n_signals, (int64_t)sms * 3);
  if (fpb < 1) fpb = 1;
  if (fpb > 16) fpb = 16;
  dim3 grid((unsigned)ceil_div64(T, fpb), n_signals);
  const size_t smem = (size_t)(2 * FPAD + FT * 17 + 8 * 16) * sizeof(float2);
  static bool configured = false;
  if (!configured) {
    SESA_CUDA(cudaFuncSetAttribute(stft2048_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  stft2048_kernel<<<grid, FT, smem, stream>>>(audio, spec, window, reinterpret_cast<const float2*>(twiddle), channels,
                                           length, hop, T, fpb);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

template <int MODE, int C>
static int launch_istft_mc(const float* spec, const float* mask, const int* inv, const float* cnt, float* out,
                           const float* window, const float* env, const float* tw, int batch, int nstems, int hop, int T,
                           int64_t out_len, int n_gathered, int G, size_t smem, cudaStream_t stream) {
  SESA_CUDA(cudaFuncSetAttribute(mask_istft2048_kernel<MODE, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div64(T, G), nstems, batch);
  mask_istft2048_kernel<MODE, C><<<grid, FT, smem, stream>>>(spec, mask, inv, cnt, out, window, env,
                                                            reinterpret_cast<const float2*>(tw), hop, T, out_len, G,
                                                            nstems, n_gathered);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

int sesa_launch_mask_istft2048(const float* spec, const float* mask, const int* inv, const float* cnt, float* out,
                               const float* window, const float* env, const float* twiddle, int batch, int nstems,
                               int channels, int hop, int T, int64_t out_len, int mode, int n_gathered,
                               cudaStream_t stream) {
  // frames per CTA: a constant of the transform geometry only — NOT of the batch — so that the grouping of frames, and
  // with it every rounding, is the same whatever batch a chunk is launched in (batch invariance, sharded == unsharded);
  // never fewer than (n_fft - hop) / hop, so that an output sample is shared by at most two CTAs (see the kernel comment)
  const int g_min = (FN - 1) / hop;
  const int G = g_min > 8 ? g_min : 8;
  const size_t smem = (size_t)2 * FPAD * sizeof(float2) + (size_t)channels * ACC * sizeof(float);
  SESA_CUDA(cudaMemsetAsync(out, 0, (size_t)batch * nstems * channels * out_len * sizeof(float), stream));
#define SESA_ISTFT_CASE(M, CC)                                                                                       \
  if (mode == M && channels == CC)                                                                                   \
    return launch_istft_mc<M, CC>(spec, mask, inv, cnt, out, window, env, twiddle, batch, nstems, hop, T, out_len,    \
                                  n_gathered, G, smem, stream);
  SESA_ISTFT_CASE(0, 1) SESA_ISTFT_CASE(0, 2) SESA_ISTFT_CASE(1, 1) SESA_ISTFT_CASE(1, 2) SESA_ISTFT_CASE(2, 1)
  SESA_ISTFT_CASE(2, 2)
#undef SESA_ISTFT_CASE
  sesa_set_error("sesa_mask_istft: bad mode/channels");
  return SESA_ERR_ARG;
}
