// Grouped GEMM on the Blackwell 5th-generation tensor cores:  C = epi( A[M,K] . W[N,K]^T ).
//
// Serves the nn.Linear call sites of the RoFormer forward (bs_roformer.py:63,67 FeedForward; :99,104 to_qkv / to_out;
// :264 MaskEstimator MLP) and, through the implicit-GEMM mode, the convolutions and Linears of MDX23C
// (mdx23c_tfc_tdf_v3.py:74-138).  Design (B200-first, not a translation of anything in the reference, which only
// calls cuBLAS / cuDNN through PyTorch):
//   * persistent CTAs walk a tile list that spans all problems of a grouped launch (per-problem TMA maps in HBM);
//   * CG = 2 (default): a cluster of two CTAs owns a 256 x 256 tile with tcgen05.mma.cta_group::2 — each CTA stages its
//     own 128 rows of A and half of the W tile, the leader issues the MMAs, commits are multicast to both CTAs;
//   * warp 0 = TMA producer (cp.async.bulk.tensor, 128B-swizzled 64-wide K slabs, mbarrier ring), warp 1 =
//     single-thread tcgen05.mma issuer with fp32 accumulators in TMEM (two BN-column buffers: the epilogue of tile i
//     overlaps the main loop of tile i+1), warps 2-9 = epilogue: for planes-only outputs the thread that reads an
//     accumulator row finishes it and TMA stores the bf16 boxes (row-owner path); outputs that carry the fp32 residual
//     stream go tcgen05.ld -> smem transpose -> coalesced phase;
//   * fp32 parity on bf16 tensor cores: operands are bf16 hi/lo planes and each K slab issues
//     Ahi.Whi + Ahi.Wlo + Alo.Whi into the same fp32 TMEM accumulator (NSPLIT = 3); NSPLIT = 1 is plain bf16;
//   * the epilogue fuses row scale (RMSNorm, from per-row sum-of-squares slots), bias, GELU/tanh/sigmoid, rotary
//     embedding, GLU, residual add, the next layer's row sums of squares, and writes fp32 and/or the bf16 planes the
//     next tensor-core op consumes; convolution taps are shifted TMA boxes, ConvTranspose scatters rows (row_map).
#include <string.h>

#include "common.cuh"
#include "sesa_b200.h"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;            // one 128-byte swizzle row of bf16
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;      // two per TMEM lane quadrant, each owning half of the tile's columns (16 warps at
                                  // <= 112 registers spill ~400 bytes per thread and run 40 % slower: measured, round 1)
constexpr int NCH = EPI_WARPS / 4; // column slices per tile (one per epilogue warp of a quadrant)
constexpr int NUM_THREADS = 64 + EPI_WARPS * 32;  // warp 0 TMA, warp 1 MMA, warps 2.. epilogue
constexpr int EPI_COLS = 16;      // accumulator columns per epilogue step
constexpr int EPI_STAGE_BYTES = 32 * EPI_COLS * 4;  // per-warp transpose buffer (2 KB)
constexpr int MAX_GROUPS = 1024;

struct alignas(128) TcGroup {
  CUtensorMap mapA;  // dims (K, M, planes), box (64, 128, 1)
  CUtensorMap mapW;  // dims (K, N, planes), box (64, BN, 1)
  CUtensorMap mapP;  // output planes, dims (columns, M, planes), box (16, 32, 1), unswizzled: the epilogue's TMA stores
  const float* bias;
  const float* rowscale;
  float* C;
  __nv_bfloat16* P;
  int64_t ldc, ldp, p_plane;
  int32_t M, N, K, tile_begin;
  int32_t n_blocks, k_blocks, tile_end, _r;
  // implicit-GEMM convolution (taps == 0: plain GEMM; mapA is then 3-D, else 5-D (c, f, t, b, plane))
  int32_t taps, kb_per_tap, conv_T, conv_F, conv_stride;
  int32_t row_map, rm_F, rm_dt, rm_df;
  int32_t tap_dt[9], tap_df[9];
  // fused RMSNorm bookkeeping: per-row partial sums of squares, one slot per (n-block, column half) of the producer
  const float* rowss;   // consumer: rs = 1 / max(sqrt(sum of ss_slots partials), 1e-12)
  float* ss_out;        // producer: ss_out[row * (2*n_blocks) + 2*nb + half] = sum of squares of the stored values
  int32_t ss_slots;     // slots per row in rowss
  int32_t p_cols;       // planes are written for columns < p_cols (0 = all)
  int32_t c_col0;       // C is written for columns >= c_col0, at column n - c_col0
  int32_t ss_ld;        // floats between consecutive rows of ss_out (>= slots; lets band-sliced launches interleave)
  int32_t p_tma;        // mapP is valid (aligned, unscattered output planes)
};
static_assert(sizeof(TcGroup) % 128 == 0, "table entries must keep the tensor maps 128-byte aligned");

template <int BN, int NSPLIT, int CG = 1>
struct Cfg {
  static constexpr int NP = NSPLIT == 3 ? 2 : 1;           // planes staged per operand
  static constexpr int A_BYTES = BM * BK * 2;               // 16 KB per plane
  static constexpr int B_BYTES = (BN / CG) * BK * 2;        // a CTA pair (CG = 2) splits the W tile's rows
  static constexpr int STAGE_BYTES = NP * (A_BYTES + B_BYTES);
  static constexpr int STAGES = (196 * 1024) / STAGE_BYTES >= 6 ? 6 : (196 * 1024) / STAGE_BYTES;
  static constexpr int EPI_OFF = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFF = EPI_OFF + EPI_WARPS * EPI_STAGE_BYTES;   // per-warp bias slice of the tile (BN / NCH floats)
  static constexpr int BAR_OFF = BIAS_OFF + BN * 4 * 4;
  static constexpr int GRP_OFF = BAR_OFF + 256;                // copy of the group record of single-problem launches
  static constexpr int TILE_OFF = GRP_OFF + (int)sizeof(TcGroup);
  static constexpr int SMEM_BYTES = TILE_OFF + MAX_GROUPS * 4 + 1024 /*alignment slack*/;
  static constexpr int TMEM_COLS = 2 * BN;                  // 256 or 512: power of two
};

// Exact-erf GELU (bs_roformer.py:66, nn.GELU()) in ~14 instructions, branch-free:
//   erf(u/sqrt2) = 1 - 2^(u r(u)) for u = min(|x|, 6), r a degree-7 polynomial fitted (weighted minimax on [0, 6]) to
//   log2(erfc(u/sqrt2))/u.  Max abs error of the erf term 5.7e-8 (fp32 rounding level), of GELU 6.7e-7 at |x| ~ 4.5
//   (1.5e-7 relative); verified against math.erf over [-12, 12].
__device__ __forceinline__ float gelu_fast(float x) {
  const float u = fminf(fabsf(x), 6.0f);
  float r = -2.834913403e-06f;
  r = fmaf(r, u, 3.937759539e-05f);
  r = fmaf(r, u, -1.861794008e-04f);
  r = fmaf(r, u, -1.369391393e-04f);
  r = fmaf(r, u, 7.063424215e-03f);
  r = fmaf(r, u, -5.249617994e-02f);
  r = fmaf(r, u, -4.592081904e-01f);
  r = fmaf(r, u, -1.151105165e+00f);
  const float e = tc::ex2_approx(u * r);   // argument in [-60, 0]: no denormal handling needed
  const float ef = copysignf(1.0f - e, x);
  return x * fmaf(0.5f, ef, 0.5f);
}

// tanh (MaskEstimator MLPs, bs_roformer.py:271) in ~15 branch-free instructions: |x| < 0.5 -> x + x^3 P(x^2) (degree-3
// minimax fit, max relative error 7.4e-8 evaluated in fp32); otherwise 1 - 2 / (2^(2 log2(e) |x|) + 1) with ex2.approx and
// rcp.approx (max relative error 1.7e-7).  tanhf() costs about twice as much and made the K = 384 mask layer of the
// 4-stem Mel model epilogue-bound (41 % tensor-pipe activity, profiles/r2_mel_launches.md).
__device__ __forceinline__ float tanh_fast(float x) {
  const float a = fabsf(x);
  const float u = x * x;
  float p = 0.017544856294989586f;
  p = fmaf(p, u, -0.05318477749824524f);
  p = fmaf(p, u, 0.1332780420780182f);
  p = fmaf(p, u, -0.33333221077919006f);
  const float small = fmaf(x * u, p, x);
  const float e = tc::ex2_approx(a * 2.8853900817779268f);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  const float big = copysignf(fmaf(-2.0f, r, 1.0f), x);
  return a < 0.5f ? small : big;
}

// The same polynomial over 16 values, written breadth-first (each Horner step over all 16 before the next) so that the
// epilogue warps always have independent instructions to issue: two warps per scheduler cannot hide a serial chain.
__device__ __forceinline__ void gelu_fast16(float4 (&o)[4]) {
  float x[16], u[16], r[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) { x[4 * i] = o[i].x; x[4 * i + 1] = o[i].y; x[4 * i + 2] = o[i].z; x[4 * i + 3] = o[i].w; }
#pragma unroll
  for (int i = 0; i < 16; ++i) u[i] = fminf(fabsf(x[i]), 6.0f);
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = fmaf(-2.834913403e-06f, u[i], 3.937759539e-05f);
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], u[i], -1.861794008e-04f);
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], u[i], -1.369391393e-04f);
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], u[i], 7.063424215e-03f);
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], u[i], -5.249617994e-02f);
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], u[i], -4.592081904e-01f);
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], u[i], -1.151105165e+00f);
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = tc::ex2_approx(u[i] * r[i]);
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = x[i] * fmaf(0.5f, copysignf(1.0f - r[i], x[i]), 0.5f);
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
}

__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ int find_group(const int* tile_end, int n_groups, int tile) {
  int g = 0;
  while (g + 1 < n_groups && tile >= tile_end[g]) ++g;
  return g;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// Epilogue flavours are compile-time so that each kernel's epilogue loop stays small enough for the
// instruction cache (one epilogue warp per scheduler cannot hide instruction-fetch misses).
enum { F_PLAIN = 0, F_ROT = 1, F_GELU = 2, F_TANH = 3, F_GLU = 4, F_GENERIC = 5 };

// CG = 2: two CTAs of a cluster (one TPC) work on one 256 x BN tile with cta_group::2 MMAs: each CTA stages its own 128
// rows of A and HALF of the W tile, the leader CTA issues one M = 256 MMA over both shared memories, and each CTA's
// TMEM receives its 128 accumulator rows.  That halves the W traffic from L2 and through shared memory.
template <int BN, int NSPLIT, int FLAVOR, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const TcGroup* __restrict__ groups, int n_groups, int total_tiles, sesa_gemm_epilogue ep,
               int out_planes) {
  using C = Cfg<BN, NSPLIT, CG>;
  const int cta_rank = CG == 2 ? (int)tc::cluster_ctarank() : 0;
  const bool leader = cta_rank == 0;
  const int unit = blockIdx.x / CG;         // CTA (pair) index: the tile scheduler's granularity
  int n_units = gridDim.x / CG;
  // Static round-robin (unit u takes tiles u, u + n_units, ...) resonates with the column-block count when the two share
  // a factor: with N = 384 (256 + 128 columns) and 74 units, every odd unit would only ever get the half-width tiles and
  // sit idle for half of the launch.  When a single problem has a ragged last column block, the stride is reduced until it
  // is coprime with the column-block count (one or two units go without work: 1-3 % of the machine instead of 25 %).
  if (n_groups == 1) {
    const int nbk = groups[0].n_blocks;
    if (nbk > 1 && groups[0].N % BN != 0) {
      auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
      while (n_units > 1 && gcd(n_units, nbk) != 1) --n_units;
    }
  }
  const int first_tile = unit < n_units ? unit : total_tiles;   // units beyond the stride own no tiles
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* full_bar = bars;                      // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;         // [STAGES]
  uint64_t* tmem_full = bars + 2 * C::STAGES;     // [2]
  uint64_t* tmem_empty = bars + 2 * C::STAGES + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
  int* tile_end = reinterpret_cast<int*>(smem + C::TILE_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < n_groups; i += NUM_THREADS) tile_end[i] = groups[i].tile_end;
  // single-problem launches (every transformer GEMM): the epilogue warps read the problem record from shared memory
  // instead of chasing it through L2 at the start of every tile
  const TcGroup* g_local = reinterpret_cast<const TcGroup*>(smem + C::GRP_OFF);
  if (n_groups == 1)
    for (int i = threadIdx.x; i < (int)(sizeof(TcGroup) / 4); i += NUM_THREADS)
      reinterpret_cast<uint32_t*>(smem + C::GRP_OFF)[i] = reinterpret_cast<const uint32_t*>(groups)[i];
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(&tmem_full[a], 1);
      tc::mbar_init(&tmem_empty[a], EPI_WARPS * CG);   // the leader's copy collects both CTAs' epilogue warps
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) {
      tc::tmem_alloc_2sm(tmem_ptr, C::TMEM_COLS);
      tc::tmem_relinquish_2sm();
    } else {
      tc::tmem_alloc(tmem_ptr, C::TMEM_COLS);
      tc::tmem_relinquish();
    }
  }
  tc::tc_fence_before();
  if (CG == 2) tc::cluster_sync(); else __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================= TMA producer =================
    if (tc::elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = first_tile; tile < total_tiles; tile += n_units) {
        const TcGroup* g = &groups[find_group(tile_end, n_groups, tile)];
        const int t = tile - g->tile_begin;
        const int mb = (t / g->n_blocks) * CG + cta_rank, nb = t % g->n_blocks;
        const int kbs = g->k_blocks;
        const int taps = g->taps;
        // a pair splits the columns the MMA really multiplies (see n_eff in the issuer): CTA r stages W rows
        // [r * n_eff / 2, (r + 1) * n_eff / 2) of the tile
        const int n_left_p = g->N - nb * BN;
        const int n_eff_p = n_left_p >= BN ? BN : ((n_left_p + 16 * CG - 1) & ~(16 * CG - 1));
        const int w_row0 = nb * BN + cta_rank * (n_eff_p / 2);
        int cf0 = 0, ct0 = 0, cb = 0, kbpt = 1;
        if (taps > 0) {   // output pixel (b, t0, f0) of the tile's first row
          const int m0 = mb * BM;
          const int q = m0 / g->conv_F;
          cf0 = (m0 - q * g->conv_F) * g->conv_stride;
          cb = q / g->conv_T;
          ct0 = (q - cb * g->conv_T) * g->conv_stride;
          kbpt = g->kb_per_tap;
        }
        for (int kb = 0; kb < kbs; ++kb) {
          tc::mbar_wait(&empty_bar[s], ph ^ 1);
          if (leader) tc::mbar_expect_tx(&full_bar[s], C::STAGE_BYTES * CG);   // bytes of both CTAs land on the leader's barrier
          uint8_t* sa = stage_base + s * C::STAGE_BYTES;
          uint8_t* sb = sa + C::NP * C::A_BYTES;
          if (taps > 0) {
            // implicit im2col: the tap is a shifted TMA box over the channels-last activation; coordinates outside
            // the grid are zero-filled by TMA, which is exactly the convolution's zero padding
            const int tap = kb / kbpt;
            const int c0 = (kb - tap * kbpt) * BK;
#pragma unroll
            for (int p = 0; p < C::NP; ++p) {
              if (CG == 2) tc::tma_load_5d_2sm(sa + p * C::A_BYTES, &g->mapA, &full_bar[s], c0, cf0 + g->tap_df[tap], ct0 + g->tap_dt[tap], cb, p);
              else tc::tma_load_5d(sa + p * C::A_BYTES, &g->mapA, &full_bar[s], c0, cf0 + g->tap_df[tap], ct0 + g->tap_dt[tap], cb, p);
            }
          } else {
#pragma unroll
            for (int p = 0; p < C::NP; ++p) {
              if (CG == 2) tc::tma_load_3d_2sm(sa + p * C::A_BYTES, &g->mapA, &full_bar[s], kb * BK, mb * BM, p);
              else tc::tma_load_3d(sa + p * C::A_BYTES, &g->mapA, &full_bar[s], kb * BK, mb * BM, p);
            }
          }
#pragma unroll
          for (int p = 0; p < C::NP; ++p) {
            if (CG == 2) tc::tma_load_3d_2sm(sb + p * C::B_BYTES, &g->mapW, &full_bar[s], kb * BK, w_row0, p);
            else tc::tma_load_3d(sb + p * C::B_BYTES, &g->mapW, &full_bar[s], kb * BK, nb * BN, p);
          }
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (leader && tc::elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int tile = first_tile; tile < total_tiles; tile += n_units) {
        const TcGroup* g = &groups[find_group(tile_end, n_groups, tile)];
        const int kbs = g->k_blocks;
        // the last column block of a problem only multiplies the columns that exist (N granularity 16; 32 for a pair)
        const int n_left = g->N - ((tile - g->tile_begin) % g->n_blocks) * BN;
        const int n_eff = n_left >= BN ? BN : ((n_left + 16 * CG - 1) & ~(16 * CG - 1));
        const uint32_t idesc = tc::make_idesc_bf16(BM * CG, n_eff, 0, 0);
        tc::mbar_wait(&tmem_empty[as], aph ^ 1);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < kbs; ++kb) {
          tc::mbar_wait(&full_bar[s], ph);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(stage_base + s * C::STAGE_BYTES);
          const uint32_t sb = sa + C::NP * C::A_BYTES;
#pragma unroll
          for (int prod = 0; prod < NSPLIT; ++prod) {
            // products: (Ahi,Whi), (Ahi,Wlo), (Alo,Whi)
            const uint32_t a_addr = sa + (prod == 2 ? C::A_BYTES : 0);
            const uint32_t b_addr = sb + (prod == 1 ? C::B_BYTES : 0);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t ad = tc::make_smem_desc_sw128(a_addr + k * UMMA_K * 2);
              const uint64_t bd = tc::make_smem_desc_sw128(b_addr + k * UMMA_K * 2);
              if (CG == 2) tc::umma_f16_2sm(d_tmem, ad, bd, idesc, (kb | prod | k) != 0 ? 1u : 0u);
              else tc::umma_f16(d_tmem, ad, bd, idesc, (kb | prod | k) != 0 ? 1u : 0u);
            }
          }
          // frees the smem stage (in both CTAs of a pair) when these MMAs retire
          if (CG == 2) tc::umma_commit_2sm(&empty_bar[s], 3); else tc::umma_commit(&empty_bar[s]);
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (CG == 2) tc::umma_commit_2sm(&tmem_full[as], 3); else tc::umma_commit(&tmem_full[as]);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ================= epilogue warps =================
    // Phase A (thread = accumulator row): TMEM -> registers -> XOR-swizzled shared memory (raw accumulators).
    // Phase B (4 lanes per row, 4 columns per lane): row scale, bias, activation, rotary, GLU, residual, row sum of
    // squares, and the fp32 / bf16-plane stores — every global access is sector-complete and coalesced, and the
    // per-element arithmetic works on float4s with per-lane constants (bias quad, rotary quad).
    const int ew = warp - 2;
    const int q = warp & 3;          // TMEM lane quadrant this warp may access
    int two_planes;                  // kept in a register: re-reading the kernel parameter inside the store loop stalls on LDC
    asm volatile("mov.u32 %0, %1;" : "=r"(two_planes) : "r"(out_planes > 1 ? 1 : 0));
    const int ch = ew >> 2;          // which half of the tile's columns
    constexpr int HALF = BN / NCH;   // columns per epilogue warp
    const uint32_t stage = tc::smem_u32(smem + C::EPI_OFF + ew * EPI_STAGE_BYTES);   // byte address in shared space
    const uint32_t bias_s = tc::smem_u32(smem + C::BIAS_OFF + ew * (BN / NCH) * 4);
    constexpr bool kGeneric = FLAVOR == F_GENERIC;
    const bool use_glu = FLAVOR == F_GLU || (kGeneric && ep.glu);
    const int act = FLAVOR == F_GELU ? SESA_ACT_GELU : FLAVOR == F_TANH ? SESA_ACT_TANH : kGeneric ? ep.act : SESA_ACT_NONE;
    const int rot_cols = (FLAVOR == F_ROT || kGeneric) ? ep.rot_cols : 0;
    const bool residual = (FLAVOR == F_PLAIN || kGeneric) ? ep.residual != 0 : false;   // only the plain flavour carries the residual stream
    const int c4 = lane & 3;         // phase B: this lane's column quad inside a 16-column step
    const int rb = lane >> 2;        // phase B: row (within each group of 8) handled by this lane
    int as = 0;
    uint32_t aph = 0;
    for (int tile = first_tile; tile < total_tiles; tile += n_units) {
      const int gi = n_groups == 1 ? 0 : find_group(tile_end, n_groups, tile);
      const TcGroup* g = n_groups == 1 ? g_local : &groups[gi];
      const int t = tile - g->tile_begin;
      const int mb = (t / g->n_blocks) * CG + cta_rank, nb = t % g->n_blocks;
      const int M = g->M, N = g->N;
      const int m_base = mb * BM + q * 32;
      const int n_half = nb * BN + ch * HALF;
      const float* __restrict__ bias = g->bias;
      float* __restrict__ Cp = g->C;
      __nv_bfloat16* __restrict__ Pp = g->P;
      const int64_t ldc = g->ldc, ldp = g->ldp, p_plane = g->p_plane;
      float* __restrict__ ss_out = g->ss_out;
      const int p_cols = g->p_cols > 0 ? g->p_cols : (use_glu ? N >> 1 : N);
      const int c_col0 = g->c_col0;
      const int row_map = g->row_map, rm_F = g->rm_F, rm_dt = g->rm_dt, rm_df = g->rm_df;
      // ---------------------------------------------------------------------------------------------------------------
      // Row-owner path (planes-only outputs: ff1, to_qkv, MaskEstimator hidden layers): the thread that reads an
      // accumulator row from TMEM finishes it — row scale and sequence position are per-thread scalars, bias comes as
      // broadcast ld.shared, no transpose, no per-row address arithmetic or predicates — and leaves the bf16 planes in
      // shared memory as dense [32 rows][16 columns] boxes that one lane hands to TMA (cp.async.bulk.tensor store, rows
      // beyond M clipped by the tensor map).
      constexpr bool kRowOwnerFlavor = FLAVOR == F_GELU || FLAVOR == F_ROT || FLAVOR == F_TANH;
      const bool row_owner = kRowOwnerFlavor && g->p_tma != 0 && Pp != nullptr && row_map == 0 && ss_out == nullptr &&
                             g->rowscale == nullptr && (g->rowss == nullptr || g->ss_slots == 4) && n_half + HALF <= N &&
                             n_half + HALF <= p_cols && (Cp == nullptr || n_half + HALF <= c_col0) &&
                             (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0);
      if (row_owner) {
        const int row = m_base + lane;
        const int rowc = min(row, M - 1);
        float rs = 1.0f;
        if (g->rowss != nullptr) {   // fused RMSNorm: F.normalize(x, dim=-1) of the GEMM input (bs_roformer.py:49)
          const float4 s4 = __ldg(reinterpret_cast<const float4*>(g->rowss) + rowc);
          rs = 1.0f / fmaxf(sqrtf(((s4.x + s4.y) + s4.z) + s4.w), 1e-12f);
        }
        // rotary table, quad-major: float4 (cos, sin, cos, sin) of column-pair quad qd at position pos sits at index
        // qd * pos_mod + pos, so the 32 rows of a warp (consecutive positions on the band axis, one or two positions on
        // the time axis) read consecutive or identical 16-byte words: 4 cache lines per warp load instead of 32
        int64_t rot_row = 0;   // this row's position
        if (FLAVOR == F_ROT && rot_cols > 0) rot_row = (int64_t)((row / ep.pos_div) % ep.pos_mod);
        {   // bias slice of the warp's columns -> shared memory (lane -> 4 consecutive columns)
          float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias != nullptr && lane * 4 < HALF) bq = __ldg(reinterpret_cast<const float4*>(bias + n_half) + lane);   // interior tile: in range
          if (lane * 4 < HALF) sts128(bias_s + lane * 16, bq);
          __syncwarp();
        }
        const void* mapP = &groups[gi].mapP;
        tc::mbar_wait(&tmem_full[as], aph);
        tc::tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + ch * HALF;
        float v[EPI_COLS];
        tmem_ld16(t_row, v);
        // (cos, sin) pairs of this row for the step's 8 column pairs: requested one step ahead, right after the previous
        // step's rotation has consumed the registers
        float4 cs[4];
        auto load_cs = [&](int cstep) {
          const int nn = n_half + cstep * EPI_COLS;
          if (FLAVOR == F_ROT && cstep < HALF / EPI_COLS && nn < rot_cols) {
            const float4* rp = reinterpret_cast<const float4*>(ep.rot) + (int64_t)((nn % ep.rot_dim) >> 2) * ep.pos_mod + rot_row;
#pragma unroll
            for (int k = 0; k < 4; ++k) cs[k] = __ldg(rp + (int64_t)k * ep.pos_mod);
          }
        };
        load_cs(0);
#pragma unroll 1
        for (int c = 0; c < HALF / EPI_COLS; ++c) {
          const int n = n_half + c * EPI_COLS;
          const bool do_rot = FLAVOR == F_ROT && n < rot_cols;
          float4 o[4];
          tc::tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 b4 = lds128(bias_s + (c * EPI_COLS + 4 * k) * 4);
            o[k].x = fmaf(v[4 * k], rs, b4.x);
            o[k].y = fmaf(v[4 * k + 1], rs, b4.y);
            o[k].z = fmaf(v[4 * k + 2], rs, b4.z);
            o[k].w = fmaf(v[4 * k + 3], rs, b4.w);
          }
          if (c + 1 < HALF / EPI_COLS) tmem_ld16(t_row + (c + 1) * EPI_COLS, v);
          if (FLAVOR == F_GELU) gelu_fast16(o);
          if (FLAVOR == F_TANH) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { o[k].x = tanh_fast(o[k].x); o[k].y = tanh_fast(o[k].y); o[k].z = tanh_fast(o[k].z); o[k].w = tanh_fast(o[k].w); }
          }
          if (do_rot) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float x1 = o[k].x, x2 = o[k].y, x3 = o[k].z, x4 = o[k].w;
              o[k].x = x1 * cs[k].x - x2 * cs[k].y;
              o[k].y = x2 * cs[k].x + x1 * cs[k].y;
              o[k].z = x3 * cs[k].z - x4 * cs[k].w;
              o[k].w = x4 * cs[k].z + x3 * cs[k].w;
            }
          }
          load_cs(c + 1);
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tc::split_bf16x2(o[k].x, o[k].y, hi[2 * k], lo[2 * k]);
            tc::split_bf16x2(o[k].z, o[k].w, hi[2 * k + 1], lo[2 * k + 1]);
          }
          // the previous step's boxes must have been read by TMA before the staging buffer is rewritten
          if (c > 0) {
            if (lane == 0) tc::bulk_wait_read0();
            __syncwarp();
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + lane * 32), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + lane * 32 + 16), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]) : "memory");
          if (two_planes) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + 1024 + lane * 32), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + 1024 + lane * 32 + 16), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]) : "memory");
          }
          tc::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tc::tma_store_3d(mapP, stage, n, m_base, 0);
            if (two_planes) tc::tma_store_3d(mapP, stage + 1024, n, m_base, 1);
            tc::bulk_commit();
          }
        }
        if (lane == 0) tc::bulk_wait_read0();   // the staging buffer is free for the next tile (either path)
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) tc::mbar_arrive_leader(&tmem_empty[as]); else tc::mbar_arrive(&tmem_empty[as]);
        }
      } else {
      // per-lane constants of the 4 rows this lane finishes in phase B (rows it*8 + rb of the warp's 32)
      float rs4[4];
      int64_t orow4[4];
      int pos4[4];
      bool ok4[4];
      float ssq[4] = {0.f, 0.f, 0.f, 0.f};
      const int mm0 = m_base + rb;
      // sequence position of row mm0 (rotary): one division per tile, the other three rows follow incrementally
      int pos_q = 0, pos_r = 0;
      if (rot_cols > 0) {
        const int qd = mm0 / ep.pos_div;
        pos_r = mm0 - qd * ep.pos_div;
        pos_q = qd % ep.pos_mod;
      }
      const float* __restrict__ rowss = g->rowss;
      const int ss_slots = g->ss_slots;
      // row scales of the lane's 4 rows: all loads are issued before the first use (rows past M read row M-1's slot)
      float rsum4[4];
      {
        float4 s4[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int64_t mc = min(mm0 + it * 8, M - 1);
          rsum4[it] = 1.0f;
          if (g->rowscale != nullptr) rsum4[it] = __ldg(g->rowscale + mc);
          if (rowss != nullptr) {
            if (ss_slots == 4) s4[it] = __ldg(reinterpret_cast<const float4*>(rowss) + mc);
          }
        }
        if (rowss != nullptr) {   // fused RMSNorm: F.normalize(x, dim=-1) of the GEMM input (bs_roformer.py:49)
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            float ssum;
            if (ss_slots == 4) ssum = ((s4[it].x + s4[it].y) + s4[it].z) + s4[it].w;
            else {
              const int64_t mc = min(mm0 + it * 8, M - 1);
              ssum = 0.f;
              for (int k = 0; k < ss_slots; ++k) ssum += rowss[mc * ss_slots + k];
            }
            rsum4[it] = 1.0f / fmaxf(sqrtf(ssum), 1e-12f);
          }
        }
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int mm = mm0 + it * 8;
        ok4[it] = mm < M;
        const float r = rsum4[it];
        rs4[it] = r;
        if (row_map == 0) orow4[it] = mm;
        else {
          const int qq = mm / rm_F;
          orow4[it] = (int64_t)(2 * qq + rm_dt) * (2 * rm_F) + 2 * (mm - qq * rm_F) + rm_df;
        }
        int pp = 0;
        if (rot_cols > 0) {
          int rr = pos_r + it * 8, qa = pos_q;
          if (ep.pos_div == 1) { qa += rr; rr = 0; }
          while (rr >= ep.pos_div) { rr -= ep.pos_div; ++qa; }
          while (qa >= ep.pos_mod) qa -= ep.pos_mod;
          pp = qa;
        }
        pos4[it] = pp;
      }
      const bool c_vec = Cp == nullptr || ((ldc & 3) == 0 && (c_col0 & 3) == 0 && (reinterpret_cast<uintptr_t>(Cp) & 15) == 0);
      const bool p_vec = Pp == nullptr || ((ldp & 3) == 0 && (p_plane & 3) == 0 && (reinterpret_cast<uintptr_t>(Pp) & 7) == 0);
      const bool vec_ok = c_vec && p_vec;

      // this warp's bias slice goes to shared memory once per tile (lane -> 4 consecutive columns); the steps read their
      // quad back with one broadcast ld.shared instead of a global load per step
      {
        float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias != nullptr) {
          const int cb = n_half + lane * 4;
          if (cb < N) bq.x = __ldg(bias + cb);
          if (cb + 1 < N) bq.y = __ldg(bias + cb + 1);
          if (cb + 2 < N) bq.z = __ldg(bias + cb + 2);
          if (cb + 3 < N) bq.w = __ldg(bias + cb + 3);
        }
        if (lane * 4 < HALF) sts128(bias_s + lane * 16, bq);   // ordered before the first read by the in-loop __syncwarp
      }

      tc::mbar_wait(&tmem_full[as], aph);
      tc::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + ch * HALF;
      // the accumulator columns of step c + 1 are requested from TMEM while step c is being finished
      float v[EPI_COLS];
      if (n_half < N) tmem_ld16(t_row, v);
      float4 cs4[4];
      auto load_rot = [&](int cstep) {   // (cos, sin) pairs of this lane's two column pairs at its four rows' positions
        const int nn = n_half + cstep * EPI_COLS;
        if ((FLAVOR == F_ROT || kGeneric) && cstep < HALF / EPI_COLS && nn < rot_cols && nn < N) {
          const int rd = ((nn + c4 * 4) % ep.rot_dim) >> 1;
#pragma unroll
          for (int it = 0; it < 4; ++it)
            cs4[it] = __ldg(reinterpret_cast<const float4*>(ep.rot) + ((int64_t)(rd >> 1) * ep.pos_mod + pos4[it]));
        }
      };
      load_rot(0);
#pragma unroll 1
      for (int c = 0; c < HALF / EPI_COLS; ++c) {
        const int n = n_half + c * EPI_COLS;
        if (n >= N) break;  // warp-uniform
        const int colb = n + c4 * 4;
        // rotary (cos, sin) quads of this lane's 4 rows were requested during the previous step (after its rotation)
        const bool do_rot = (FLAVOR == F_ROT || kGeneric) && n < rot_cols;
        // ---- phase A
        tc::tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
          sts128(stage + lane * (EPI_COLS * 4) + ((j4 ^ ((lane >> 1) & 3)) << 4),
                 make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]));
        // (the rotary / generic flavours are short of registers: they request the next columns after phase B instead)
        constexpr bool kEarlyLd = FLAVOR != F_ROT && FLAVOR != F_GENERIC;
        const bool more = c + 1 < HALF / EPI_COLS && n + EPI_COLS < N;
        if (kEarlyLd && more) tmem_ld16(t_row + (c + 1) * EPI_COLS, v);
        const bool interior = vec_ok && n + EPI_COLS <= N;   // warp-uniform: no ragged right edge in this step
        float4 res[4];
        if (residual && interior) {   // issue the residual reads before the shared-memory round trip completes
#pragma unroll
          for (int it = 0; it < 4; ++it)
            res[it] = ok4[it] ? *reinterpret_cast<const float4*>(Cp + orow4[it] * ldc + (colb - c_col0)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
        // ---- phase B
        const float4 b4 = lds128(bias_s + (c * EPI_COLS + c4 * 4) * 4);
        if (interior && !use_glu) {
          // fast path: straight-line code over the lane's 4 rows (16 independent element chains for the scheduler);
          // only the stores are predicated
          float4 o[4];
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + rb;
            o[it] = lds128(stage + r * (EPI_COLS * 4) + ((c4 ^ ((r >> 1) & 3)) << 4));
          }
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            o[it].x = fmaf(o[it].x, rs4[it], b4.x);
            o[it].y = fmaf(o[it].y, rs4[it], b4.y);
            o[it].z = fmaf(o[it].z, rs4[it], b4.z);
            o[it].w = fmaf(o[it].w, rs4[it], b4.w);
          }
          if (act == SESA_ACT_GELU) {
            gelu_fast16(o);
          } else if (act == SESA_ACT_TANH) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              o[it].x = tanh_fast(o[it].x); o[it].y = tanh_fast(o[it].y); o[it].z = tanh_fast(o[it].z); o[it].w = tanh_fast(o[it].w);
            }
          } else if (act == SESA_ACT_SIGMOID) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              o[it].x = 1.0f / (1.0f + expf(-o[it].x)); o[it].y = 1.0f / (1.0f + expf(-o[it].y));
              o[it].z = 1.0f / (1.0f + expf(-o[it].z)); o[it].w = 1.0f / (1.0f + expf(-o[it].w));
            }
          }
          if (do_rot) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const float4 cs = cs4[it];
              const float x1 = o[it].x, x2 = o[it].y, x3 = o[it].z, x4 = o[it].w;
              o[it].x = x1 * cs.x - x2 * cs.y;
              o[it].y = x2 * cs.x + x1 * cs.y;
              o[it].z = x3 * cs.z - x4 * cs.w;
              o[it].w = x4 * cs.z + x3 * cs.w;
            }
          }
          load_rot(c + 1);   // cs4 is dead from here on: the next step's constants travel under the stores and phase A
          if (residual) {
#pragma unroll
            for (int it = 0; it < 4; ++it) { o[it].x += res[it].x; o[it].y += res[it].y; o[it].z += res[it].z; o[it].w += res[it].w; }
          }
          const bool wc = Cp != nullptr && colb >= c_col0;
          const bool wp = Pp != nullptr && colb < p_cols;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            if (ss_out != nullptr) ssq[it] += o[it].x * o[it].x + o[it].y * o[it].y + o[it].z * o[it].z + o[it].w * o[it].w;
            if (wc && ok4[it]) *reinterpret_cast<float4*>(Cp + orow4[it] * ldc + (colb - c_col0)) = o[it];
            if (wp) {
              uint32_t h0, l0, h1, l1;
              tc::split_bf16x2(o[it].x, o[it].y, h0, l0);
              tc::split_bf16x2(o[it].z, o[it].w, h1, l1);
              if (ok4[it]) {
                __nv_bfloat16* pr = Pp + orow4[it] * ldp + colb;
                *reinterpret_cast<uint2*>(pr) = make_uint2(h0, h1);
                if (two_planes) *reinterpret_cast<uint2*>(pr + p_plane) = make_uint2(l0, l1);
              }
            }
          }
        } else {
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + rb;
            float4 o = lds128(stage + r * (EPI_COLS * 4) + ((c4 ^ ((r >> 1) & 3)) << 4));
            if (!ok4[it]) continue;
            o.x = fmaf(o.x, rs4[it], b4.x);
            o.y = fmaf(o.y, rs4[it], b4.y);
            o.z = fmaf(o.z, rs4[it], b4.z);
            o.w = fmaf(o.w, rs4[it], b4.w);
            if (act == SESA_ACT_GELU) {
              o.x = gelu_fast(o.x); o.y = gelu_fast(o.y); o.z = gelu_fast(o.z); o.w = gelu_fast(o.w);
            } else if (act == SESA_ACT_TANH) {
              o.x = tanh_fast(o.x); o.y = tanh_fast(o.y); o.z = tanh_fast(o.z); o.w = tanh_fast(o.w);
            } else if (act == SESA_ACT_SIGMOID) {
              o.x = 1.0f / (1.0f + expf(-o.x)); o.y = 1.0f / (1.0f + expf(-o.y));
              o.z = 1.0f / (1.0f + expf(-o.z)); o.w = 1.0f / (1.0f + expf(-o.w));
            }
            if (do_rot) {
              const float4 cs = cs4[it];
              const float x1 = o.x, x2 = o.y, x3 = o.z, x4 = o.w;
              o.x = x1 * cs.x - x2 * cs.y;
              o.y = x2 * cs.x + x1 * cs.y;
              o.z = x3 * cs.z - x4 * cs.w;
              o.w = x4 * cs.z + x3 * cs.w;
            }
            const int64_t orow = orow4[it];
            if (use_glu) {
              // W rows interleaved (value, gate): out[:, colb/2 + {0,1}] = (o.x * sigmoid(o.y), o.z * sigmoid(o.w))
              const float g0 = o.x * (1.0f / (1.0f + expf(-o.y)));
              const float g1 = o.z * (1.0f / (1.0f + expf(-o.w)));
              const int oc = colb >> 1;
              if (Cp != nullptr) {
                float* cp = Cp + orow * ldc + oc;
                if (colb + 3 < N && ((reinterpret_cast<uintptr_t>(cp) & 7) == 0)) {
                  *reinterpret_cast<float2*>(cp) = make_float2(g0, g1);   // 4 lanes x 8 bytes: one full sector per row
                } else {
                  if (colb + 1 < N) cp[0] = g0;
                  if (colb + 3 < N) cp[1] = g1;
                }
              }
              if (Pp != nullptr) {
                __nv_bfloat16 h, l;
                if (colb + 1 < N) { tc::split_bf16(g0, h, l); Pp[orow * ldp + oc] = h; if (two_planes) Pp[orow * ldp + p_plane + oc] = l; }
                if (colb + 3 < N) { tc::split_bf16(g1, h, l); Pp[orow * ldp + oc + 1] = h; if (two_planes) Pp[orow * ldp + p_plane + oc + 1] = l; }
              }
              continue;
            }
            // ragged right edge of the problem, or unaligned outputs
            const float ov[4] = {o.x, o.y, o.z, o.w};
            for (int e = 0; e < 4 && colb + e < N; ++e) {
              float val = ov[e];
              if (Cp != nullptr && colb + e >= c_col0) {
                float* cp = Cp + orow * ldc + colb + e - c_col0;
                if (residual) val += *cp;
                *cp = val;
              }
              ssq[it] += val * val;
              if (Pp != nullptr && colb + e < p_cols) {
                __nv_bfloat16 h, l;
                tc::split_bf16(val, h, l);
                Pp[orow * ldp + colb + e] = h;
                if (two_planes) Pp[orow * ldp + p_plane + colb + e] = l;
              }
            }
          }
          load_rot(c + 1);
        }
        if (!kEarlyLd && more) tmem_ld16(t_row + (c + 1) * EPI_COLS, v);
        __syncwarp();
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) tc::mbar_arrive_leader(&tmem_empty[as]); else tc::mbar_arrive(&tmem_empty[as]);
      }
      if (ss_out != nullptr) {   // this warp's slot of the row sums of squares (4 lanes per row)
        const int ss_ld = g->ss_ld;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          float v = ssq[it];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          if (c4 == 0 && ok4[it]) ss_out[orow4[it] * ss_ld + NCH * nb + ch] = v;
        }
      }
      }
      if (++as == 2) { as = 0; aph ^= 1; }
    }
    if (lane == 0) tc::bulk_wait0();   // every TMA store of this warp has completed before the CTA may exit
  }

  tc::tc_fence_before();
  if (CG == 2) tc::cluster_sync(); else __syncthreads();
  if (warp == 1) {
    if (CG == 2) tc::tmem_dealloc_2sm(tmem_base, C::TMEM_COLS); else tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, int NSPLIT, int FLAVOR, int CG>
int launch_gemm_tc_f(const TcGroup* table, int n_groups, int total_tiles, const sesa_gemm_epilogue& ep, int out_planes,
                     cudaStream_t stream) {
  using C = Cfg<BN, NSPLIT, CG>;
  static_assert(C::STAGES >= 2, "pipeline needs at least two stages");
  static bool configured = false;
  if (!configured) {
    SESA_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, NSPLIT, FLAVOR, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   C::SMEM_BYTES));
    configured = true;
  }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    SESA_CUDA(cudaGetDevice(&dev));
    SESA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int units = sms / CG;   // persistent CTAs (CTA pairs): one per SM (TPC)
  const int grid = (total_tiles < units ? total_tiles : units) * CG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SESA_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, NSPLIT, FLAVOR, CG>, table, n_groups, total_tiles, ep, out_planes));
  return SESA_OK;
}

template <int BN, int NSPLIT, int CG>
int launch_gemm_tc(const TcGroup* table, int n_groups, int total_tiles, const sesa_gemm_epilogue& ep, int out_planes,
                   cudaStream_t stream) {
  int flavor = F_GENERIC;
  if (ep.glu && ep.act == SESA_ACT_NONE && ep.rot_cols == 0) flavor = F_GLU;
  else if (!ep.glu && ep.rot_cols > 0 && ep.act == SESA_ACT_NONE) flavor = F_ROT;
  else if (!ep.glu && ep.rot_cols == 0 && ep.act == SESA_ACT_NONE) flavor = F_PLAIN;
  else if (!ep.glu && ep.rot_cols == 0 && ep.act == SESA_ACT_GELU) flavor = F_GELU;
  else if (!ep.glu && ep.rot_cols == 0 && ep.act == SESA_ACT_TANH) flavor = F_TANH;
  switch (flavor) {
    case F_PLAIN: return launch_gemm_tc_f<BN, NSPLIT, F_PLAIN, CG>(table, n_groups, total_tiles, ep, out_planes, stream);
    case F_ROT: return launch_gemm_tc_f<BN, NSPLIT, F_ROT, CG>(table, n_groups, total_tiles, ep, out_planes, stream);
    case F_GELU: return launch_gemm_tc_f<BN, NSPLIT, F_GELU, CG>(table, n_groups, total_tiles, ep, out_planes, stream);
    case F_TANH: return launch_gemm_tc_f<BN, NSPLIT, F_TANH, CG>(table, n_groups, total_tiles, ep, out_planes, stream);
    case F_GLU: return launch_gemm_tc_f<BN, NSPLIT, F_GLU, CG>(table, n_groups, total_tiles, ep, out_planes, stream);
    default: return launch_gemm_tc_f<BN, NSPLIT, F_GENERIC, CG>(table, n_groups, total_tiles, ep, out_planes, stream);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host side
sesa_encode_tiled_fn sesa_get_encode_tiled() {
  static sesa_encode_tiled_fn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<sesa_encode_tiled_fn>(p);
  }
  return fn;
}

static int make_tmap_bf16_impl(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                               const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                               CUtensorMapSwizzle swizzle) {
  sesa_encode_tiled_fn enc = sesa_get_encode_tiled();
  if (enc == nullptr) {
    sesa_set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return SESA_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides != nullptr ? elem_strides[i] : 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sesa_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu, stride0 %llu B)", (int)r,
                   rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                   (unsigned long long)(rank > 1 ? strides_bytes[0] : 0));
    return SESA_ERR_CUDA;
  }
  return SESA_OK;
}

int sesa_make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  return make_tmap_bf16_impl(map, base, rank, dims, strides_bytes, box, elem_strides, CU_TENSOR_MAP_SWIZZLE_128B);
}
int sesa_make_tmap_bf16_plain(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                              const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap_bf16_impl(map, base, rank, dims, strides_bytes, box, nullptr, CU_TENSOR_MAP_SWIZZLE_NONE);
}

extern "C" int64_t sesa_gemm_tc_table_bytes(int n_groups) { return (int64_t)sizeof(TcGroup) * (n_groups > 0 ? n_groups : 0); }

extern "C" int sesa_gemm_tc_build(const sesa_tc_problem* pr, int n_groups, int block_n, int cta_group, void* table_host,
                                  int* total_tiles) {
  SESA_CHECK_ARG(cta_group == 1 || (cta_group == 2 && block_n == 256), "sesa_gemm_tc_build: cta_group must be 1, or 2 with block_n 256");
  SESA_CHECK_ARG(pr != nullptr && table_host != nullptr && total_tiles != nullptr, "sesa_gemm_tc_build: null argument");
  SESA_CHECK_ARG(n_groups > 0 && n_groups <= MAX_GROUPS, "sesa_gemm_tc_build: group count %d out of range", n_groups);
  SESA_CHECK_ARG(block_n == 128 || block_n == 256, "sesa_gemm_tc_build: block_n must be 128 or 256");
  TcGroup* tab = reinterpret_cast<TcGroup*>(table_host);
  int tiles = 0;
  for (int i = 0; i < n_groups; ++i) {
    const sesa_tc_problem& p = pr[i];
    SESA_CHECK_ARG(p.M > 0 && p.N > 0 && p.K > 0, "sesa_gemm_tc_build: empty problem %d", i);
    SESA_CHECK_ARG((p.lda & 7) == 0 && (p.ldw & 7) == 0 && (p.a_plane & 7) == 0 && (p.w_plane & 7) == 0,
                   "sesa_gemm_tc_build: problem %d: operand strides must be multiples of 8 bf16 elements", i);
    SESA_CHECK_ARG((reinterpret_cast<uintptr_t>(p.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.W) & 15) == 0,
                   "sesa_gemm_tc_build: problem %d: operands must be 16-byte aligned", i);
    SESA_CHECK_ARG(p.ldw >= p.K, "sesa_gemm_tc_build: problem %d: weight row stride smaller than K", i);
    TcGroup g;
    memset(&g, 0, sizeof(g));
    int rc;
    if (p.conv_taps > 0) {
      const int F = p.conv_F, T = p.conv_T, s = p.conv_stride;
      SESA_CHECK_ARG(p.conv_taps <= 9 && p.conv_cin > 0 && p.conv_B > 0 && T > 0 && F > 0 && (s == 1 || s == 2),
                     "sesa_gemm_tc_build: problem %d: bad convolution geometry", i);
      SESA_CHECK_ARG((int64_t)p.conv_B * T * F == p.M, "sesa_gemm_tc_build: problem %d: M != B*T*F", i);
      const int box_f = F < BM ? F : BM;
      const int box_t = BM / box_f;
      SESA_CHECK_ARG((F >= BM ? F % BM == 0 : BM % F == 0) && T % box_t == 0,
                     "sesa_gemm_tc_build: problem %d: a 128-pixel tile must cover whole rows of the %d x %d grid", i, T, F);
      const int kbpt = (p.conv_cin + BK - 1) / BK;
      SESA_CHECK_ARG(p.K == p.conv_taps * kbpt * BK, "sesa_gemm_tc_build: problem %d: K must be taps*round_up(cin,64)", i);
      SESA_CHECK_ARG(p.lda >= p.conv_cin, "sesa_gemm_tc_build: problem %d: channel stride smaller than cin", i);
      const uint64_t dimsA[5] = {(uint64_t)p.conv_cin, (uint64_t)p.conv_inF, (uint64_t)p.conv_inT, (uint64_t)p.conv_B, 2};
      const uint64_t strA[4] = {(uint64_t)p.lda * 2, (uint64_t)p.lda * 2 * p.conv_inF,
                                (uint64_t)p.lda * 2 * p.conv_inF * p.conv_inT,
                                (uint64_t)(p.a_plane > 0 ? p.a_plane : p.lda * (int64_t)p.conv_inF * p.conv_inT * p.conv_B) * 2};
      const uint32_t boxA[5] = {BK, (uint32_t)(box_f * s), (uint32_t)(box_t * s), 1, 1};
      const uint32_t esA[5] = {1, (uint32_t)s, (uint32_t)s, 1, 1};
      rc = sesa_make_tmap_bf16(&g.mapA, p.A, 5, dimsA, strA, boxA, esA);
      if (rc != SESA_OK) return rc;
      g.taps = p.conv_taps;
      g.kb_per_tap = kbpt;
      g.conv_T = T;
      g.conv_F = F;
      g.conv_stride = s;
      for (int k = 0; k < 9; ++k) {
        g.tap_dt[k] = p.conv_dt[k];
        g.tap_df[k] = p.conv_df[k];
      }
    } else {
      SESA_CHECK_ARG(p.lda >= p.K, "sesa_gemm_tc_build: problem %d: row stride smaller than K", i);
      const uint64_t dimsA[3] = {(uint64_t)p.K, (uint64_t)p.M, 2};
      const uint64_t strA[2] = {(uint64_t)p.lda * 2, (uint64_t)(p.a_plane > 0 ? p.a_plane : p.lda * (int64_t)p.M) * 2};
      const uint32_t boxA[3] = {BK, BM, 1};
      rc = sesa_make_tmap_bf16(&g.mapA, p.A, 3, dimsA, strA, boxA);
      if (rc != SESA_OK) return rc;
    }
    SESA_CHECK_ARG(p.row_map == 0 || (p.row_map == 1 && p.rm_F > 0), "sesa_gemm_tc_build: problem %d: bad row_map", i);
    g.row_map = p.row_map;
    g.rm_F = p.rm_F;
    g.rm_dt = p.rm_dt;
    g.rm_df = p.rm_df;
    const uint64_t dimsW[3] = {(uint64_t)p.K, (uint64_t)p.N, 2};
    const uint64_t strW[2] = {(uint64_t)p.ldw * 2, (uint64_t)(p.w_plane > 0 ? p.w_plane : p.ldw * (int64_t)p.N) * 2};
    const uint32_t boxW[3] = {BK, (uint32_t)(block_n / cta_group), 1};   // each CTA of a pair stages half of the W tile
    rc = sesa_make_tmap_bf16(&g.mapW, p.W, 3, dimsW, strW, boxW);
    if (rc != SESA_OK) return rc;
    g.bias = p.bias;
    g.rowscale = p.rowscale;
    g.rowss = p.rowss;
    g.ss_out = p.ss_out;
    g.ss_slots = p.ss_slots;
    g.p_cols = p.p_cols;
    g.c_col0 = p.c_col0;
    g.ss_ld = p.ss_ld > 0 ? p.ss_ld : NCH * ((p.N + block_n - 1) / block_n);
    SESA_CHECK_ARG(p.rowss == nullptr || (p.ss_slots > 0 && p.ss_slots <= 16), "sesa_gemm_tc_build: problem %d: bad ss_slots", i);
    SESA_CHECK_ARG(p.ss_out == nullptr || (p.C == nullptr || ((p.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0)),
                   "sesa_gemm_tc_build: problem %d: ss_out needs 16-byte aligned fp32 output rows", i);
    SESA_CHECK_ARG(p.c_col0 >= 0 && (p.c_col0 & 3) == 0 && p.p_cols >= 0 && (p.p_cols & 3) == 0,
                   "sesa_gemm_tc_build: problem %d: p_cols / c_col0 must be multiples of 4", i);
    g.p_tma = 0;
    if (p.P != nullptr && p.row_map == 0 && (p.ldp & 7) == 0 && (p.p_plane & 7) == 0 &&
        (reinterpret_cast<uintptr_t>(p.P) & 15) == 0) {
      const uint64_t pc = (uint64_t)(p.p_cols > 0 ? p.p_cols : p.N);
      const uint64_t dimsP[3] = {pc, (uint64_t)p.M, (uint64_t)(p.p_plane > 0 ? 2 : 1)};
      const uint64_t strP[2] = {(uint64_t)p.ldp * 2, (uint64_t)(p.p_plane > 0 ? p.p_plane : p.ldp * (int64_t)p.M) * 2};
      const uint32_t boxP[3] = {EPI_COLS, 32, 1};
      if (pc >= EPI_COLS && sesa_make_tmap_bf16_plain(&g.mapP, p.P, 3, dimsP, strP, boxP) == SESA_OK) g.p_tma = 1;
    }
    g.C = p.C;
    g.P = reinterpret_cast<__nv_bfloat16*>(p.P);
    g.ldc = p.ldc;
    g.ldp = p.ldp;
    g.p_plane = p.p_plane;
    g.M = p.M;
    g.N = p.N;
    g.K = p.K;
    g.n_blocks = (p.N + block_n - 1) / block_n;
    g.k_blocks = (p.K + BK - 1) / BK;
    g.tile_begin = tiles;
    tiles += ((p.M + BM * cta_group - 1) / (BM * cta_group)) * g.n_blocks;
    g.tile_end = tiles;
    memcpy(&tab[i], &g, sizeof(g));
  }
  *total_tiles = tiles;
  return SESA_OK;
}

extern "C" int sesa_gemm_tc(const void* table_dev, int n_groups, int total_tiles, int block_n, int cta_group, int nsplit,
                            int out_planes, const sesa_gemm_epilogue* ep, void* stream) {
  SESA_CHECK_ARG(cta_group == 1 || (cta_group == 2 && block_n == 256), "sesa_gemm_tc: cta_group must be 1, or 2 with block_n 256");
  SESA_CHECK_ARG(table_dev != nullptr && ep != nullptr, "sesa_gemm_tc: null argument");
  SESA_CHECK_ARG(n_groups > 0 && n_groups <= MAX_GROUPS, "sesa_gemm_tc: group count %d out of range", n_groups);
  SESA_CHECK_ARG(nsplit == 1 || nsplit == 3, "sesa_gemm_tc: nsplit must be 1 or 3");
  SESA_CHECK_ARG(out_planes == 1 || out_planes == 2, "sesa_gemm_tc: out_planes must be 1 or 2");
  SESA_CHECK_ARG(ep->rot_cols == 0 || (ep->rot != nullptr && ep->rot_dim > 0 && (ep->rot_dim & 15) == 0 &&
                                       (ep->rot_cols & 15) == 0 && ep->pos_div > 0 && ep->pos_mod > 0 &&
                                       (reinterpret_cast<uintptr_t>(ep->rot) & 15) == 0),
                 "sesa_gemm_tc: bad rotary parameters (rot_cols and rot_dim must be multiples of 16)");
  SESA_CHECK_ARG(!(ep->glu && ep->residual), "sesa_gemm_tc: glu and residual are exclusive");
  if (total_tiles <= 0) return SESA_OK;
  const TcGroup* tab = reinterpret_cast<const TcGroup*>(table_dev);
  cudaStream_t st = (cudaStream_t)stream;
  if (cta_group == 2 && nsplit == 3) return launch_gemm_tc<256, 3, 2>(tab, n_groups, total_tiles, *ep, out_planes, st);
  if (cta_group == 2 && nsplit == 1) return launch_gemm_tc<256, 1, 2>(tab, n_groups, total_tiles, *ep, out_planes, st);
  if (block_n == 256 && nsplit == 3) return launch_gemm_tc<256, 3, 1>(tab, n_groups, total_tiles, *ep, out_planes, st);
  if (block_n == 256 && nsplit == 1) return launch_gemm_tc<256, 1, 1>(tab, n_groups, total_tiles, *ep, out_planes, st);
  if (block_n == 128 && nsplit == 3) return launch_gemm_tc<128, 3, 1>(tab, n_groups, total_tiles, *ep, out_planes, st);
  if (block_n == 128 && nsplit == 1) return launch_gemm_tc<128, 1, 1>(tab, n_groups, total_tiles, *ep, out_planes, st);
  sesa_set_error("sesa_gemm_tc: unsupported block_n %d", block_n);
  return SESA_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------ row prep
// One warp per row: optional L2 normalisation, bf16 hi/lo planes, optional gate logits.
__global__ void __launch_bounds__(256) prep_rows_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int dim,
                                                        int normalize, __nv_bfloat16* __restrict__ planes, int64_t ldp,
                                                        int64_t p_plane, int out_planes,
                                                        const float* __restrict__ gate_w,
                                                        const float* __restrict__ gate_b, int n_gates,
                                                        float* __restrict__ gates, int64_t ldg,
                                                        float* __restrict__ rowinv, int ss_slots) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * ldx;
  const int nv = dim >> 2;  // dim % 4 == 0 (checked by the host entry)
  float inv = 1.0f;
  if (normalize) {
    float ss = 0.f;
    for (int i = lane; i < nv; i += 32) {
      const float4 v = *reinterpret_cast<const float4*>(xr + 4 * i);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (normalize == 2) {   // planes stay raw; hand the sum of squares to the consuming GEMM (slot 0, others 0)
      if (rowinv != nullptr && lane < ss_slots) rowinv[row * ss_slots + lane] = lane == 0 ? ss : 0.f;
    } else {
      inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    }
  }
  if (rowinv != nullptr && normalize != 2 && lane == 0) rowinv[row] = inv;
  float gacc[8];
#pragma unroll
  for (int h = 0; h < 8; ++h) gacc[h] = 0.f;
  for (int i = lane; i < nv; i += 32) {
    float4 v = *reinterpret_cast<const float4*>(xr + 4 * i);
    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    if (planes != nullptr) {
      __nv_bfloat16 h0, l0, h1, l1, h2, l2, h3, l3;
      tc::split_bf16(v.x, h0, l0); tc::split_bf16(v.y, h1, l1);
      tc::split_bf16(v.z, h2, l2); tc::split_bf16(v.w, h3, l3);
      __nv_bfloat16* pr = planes + row * ldp + 4 * i;
      *reinterpret_cast<uint2*>(pr) = make_uint2(tc::pack_bf16(h0, h1), tc::pack_bf16(h2, h3));
      if (out_planes > 1) *reinterpret_cast<uint2*>(pr + p_plane) = make_uint2(tc::pack_bf16(l0, l1), tc::pack_bf16(l2, l3));
    }
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      if (h < n_gates) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(gate_w + (int64_t)h * dim + 4 * i));
        gacc[h] += v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w;
      }
    }
  }
  if (n_gates > 0) {
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      if (h < n_gates) {
        float a = gacc[h];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) gates[row * ldg + h] = a + (gate_b != nullptr ? gate_b[h] : 0.f);
      }
    }
  }
}

extern "C" int sesa_prep_rows(const float* x, int64_t ldx, int64_t rows, int dim, int normalize, void* planes,
                              int64_t ldp, int64_t p_plane, int out_planes, const float* gate_w, const float* gate_b,
                              int n_gates, float* gates, int64_t ldg, float* rowinv, int ss_slots, void* stream) {
  SESA_CHECK_ARG(dim > 0 && (dim & 3) == 0 && (ldx & 3) == 0, "sesa_prep_rows: dim and ldx must be multiples of 4");
  SESA_CHECK_ARG(planes == nullptr || ((ldp & 3) == 0 && (p_plane & 3) == 0), "sesa_prep_rows: plane strides must be multiples of 4");
  SESA_CHECK_ARG(n_gates >= 0 && n_gates <= 8, "sesa_prep_rows: at most 8 gate outputs");
  SESA_CHECK_ARG(n_gates == 0 || (gate_w != nullptr && gates != nullptr), "sesa_prep_rows: gates need weights and an output");
  SESA_CHECK_ARG(out_planes == 1 || out_planes == 2, "sesa_prep_rows: out_planes must be 1 or 2");
  SESA_CHECK_ARG(normalize >= 0 && normalize <= 2 && (normalize != 2 || (ss_slots >= 1 && ss_slots <= 32)),
                 "sesa_prep_rows: normalize must be 0, 1 or 2 (with 1 <= ss_slots <= 32)");
  if (rows <= 0) return SESA_OK;
  const int wpb = 8;
  prep_rows_kernel<<<(unsigned)ceil_div64(rows, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      x, ldx, rows, dim, normalize, reinterpret_cast<__nv_bfloat16*>(planes), ldp, p_plane, out_planes, gate_w, gate_b,
      n_gates, gates, ldg, rowinv, ss_slots);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

// Mel-Band-RoFormer ends every Transformer with an RMSNorm (mel_band_roformer.py:218,226) whose output IS the new residual
// stream.  One warp per row: y = x / max(||x||, 1e-12) * sqrt(dim) * gamma written back in place (the arithmetic and order of
// rmsnorm_kernel, misc.cu), then — from the row the warp has just written — the bf16 hi/lo planes of y and its sum of
// squares (slot 0) for the fused RMSNorm of the consuming GEMM (the arithmetic and order of prep_rows_kernel, normalize 2).
// One read of x and one launch instead of two reads and two launches; bit-identical to the two-kernel sequence.
__global__ void __launch_bounds__(256) rmsnorm_planes_kernel(float* __restrict__ x, const float* __restrict__ gamma, int64_t rows,
                                                             int dim, float scale, __nv_bfloat16* __restrict__ planes, int64_t ldp,
                                                             int64_t p_plane, int out_planes, float* __restrict__ ss_out,
                                                             int ss_slots) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float* xr = x + row * dim;
  float ss = 0.f;
  for (int i = lane; i < dim; i += 32) ss += xr[i] * xr[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  for (int i = lane; i < dim; i += 32) xr[i] = xr[i] * inv * scale * gamma[i];
  __syncwarp();
  const int nv = dim >> 2;
  float s2 = 0.f;
  for (int i = lane; i < nv; i += 32) {
    const float4 v = *reinterpret_cast<const float4*>(xr + 4 * i);
    s2 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  if (ss_out != nullptr && lane < ss_slots) ss_out[row * ss_slots + lane] = lane == 0 ? s2 : 0.f;
  for (int i = lane; i < nv; i += 32) {
    const float4 v = *reinterpret_cast<const float4*>(xr + 4 * i);
    __nv_bfloat16 h0, l0, h1, l1, h2, l2, h3, l3;
    tc::split_bf16(v.x, h0, l0); tc::split_bf16(v.y, h1, l1);
    tc::split_bf16(v.z, h2, l2); tc::split_bf16(v.w, h3, l3);
    __nv_bfloat16* pr = planes + row * ldp + 4 * i;
    *reinterpret_cast<uint2*>(pr) = make_uint2(tc::pack_bf16(h0, h1), tc::pack_bf16(h2, h3));
    if (out_planes > 1) *reinterpret_cast<uint2*>(pr + p_plane) = make_uint2(tc::pack_bf16(l0, l1), tc::pack_bf16(l2, l3));
  }
}

extern "C" int sesa_rmsnorm_planes(float* x, const float* gamma, int64_t rows, int dim, void* planes, int64_t ldp, int64_t p_plane,
                                   int out_planes, float* ss_out, int ss_slots, void* stream) {
  SESA_CHECK_ARG(dim > 0 && (dim & 3) == 0, "sesa_rmsnorm_planes: dim must be a multiple of 4");
  SESA_CHECK_ARG(planes != nullptr && (ldp & 3) == 0 && (p_plane & 3) == 0, "sesa_rmsnorm_planes: planes with strides that are multiples of 4");
  SESA_CHECK_ARG((out_planes == 1 || out_planes == 2) && ss_slots >= 1 && ss_slots <= 32, "sesa_rmsnorm_planes: bad out_planes / ss_slots");
  if (rows <= 0) return SESA_OK;
  rmsnorm_planes_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, (cudaStream_t)stream>>>(
      x, gamma, rows, dim, sqrtf((float)dim), reinterpret_cast<__nv_bfloat16*>(planes), ldp, p_plane, out_planes, ss_out, ss_slots);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

// BandSplit prologue (bs_roformer.py:241-249; Mel :250-258): per (row, band) L2-normalise the band's slice of the feature
// row (F.normalize of RMSNorm; gamma*sqrt(dim_in) is folded into the band's weight) and write it as bf16 hi/lo planes at
// the band's 16-byte aligned plane column — the A operand of the grouped tensor-core GEMM.  One warp per (row, band).
__global__ void __launch_bounds__(256) band_prep_kernel(const float* __restrict__ feat, int64_t ld_feat, int64_t rows, int nb,
                                                        const int32_t* __restrict__ offs, const int32_t* __restrict__ poffs,
                                                        __nv_bfloat16* __restrict__ planes, int64_t ldp, int64_t p_plane,
                                                        int out_planes) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= rows * nb) return;
  const int64_t r = w / nb;
  const int b = (int)(w - r * nb);
  const int c0 = offs[b], n = offs[b + 1] - c0;
  const int p0 = poffs[b], np = poffs[b + 1] - p0;
  const float* x = feat + r * ld_feat + c0;
  float ss = 0.f;
  for (int i = lane; i < n; i += 32) ss = fmaf(x[i], x[i], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  __nv_bfloat16* pr = planes + r * ldp + p0;
  for (int i = lane; i < np; i += 32) {
    __nv_bfloat16 h, l;
    tc::split_bf16(i < n ? x[i] * inv : 0.f, h, l);
    pr[i] = h;
    if (out_planes > 1) pr[i + p_plane] = l;
  }
}

extern "C" int sesa_band_prep(const float* feat, int64_t ld_feat, int64_t rows, int n_bands, const int32_t* offs,
                              const int32_t* plane_offs, void* planes, int64_t ldp, int64_t p_plane, int out_planes,
                              void* stream) {
  SESA_CHECK_ARG(n_bands > 0 && (out_planes == 1 || out_planes == 2), "sesa_band_prep: bad arguments");
  if (rows <= 0) return SESA_OK;
  const int wpb = 8;
  band_prep_kernel<<<(unsigned)ceil_div64(rows * n_bands, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      feat, ld_feat, rows, n_bands, offs, plane_offs, reinterpret_cast<__nv_bfloat16*>(planes), ldp, p_plane, out_planes);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

__global__ void split_weight_kernel(const float* __restrict__ w, int64_t rows, int64_t cols,
                                    __nv_bfloat16* __restrict__ planes, int64_t ldp) {
  const int64_t total = rows * ldp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ldp, c = i % ldp;
    const float v = c < cols ? w[r * cols + c] : 0.f;
    __nv_bfloat16 h, l;
    tc::split_bf16(v, h, l);
    planes[i] = h;
    planes[total + i] = l;
  }
}

extern "C" int sesa_split_weight(const float* w, int64_t rows, int64_t cols, void* planes, int64_t ldp, void* stream) {
  SESA_CHECK_ARG(ldp >= cols && (ldp & 7) == 0, "sesa_split_weight: ldp must be a multiple of 8 and >= cols");
  if (rows <= 0) return SESA_OK;
  const int64_t total = rows * ldp;
  split_weight_kernel<<<(unsigned)min((int64_t)2048, ceil_div64(total, 256)), 256, 0, (cudaStream_t)stream>>>(
      w, rows, cols, reinterpret_cast<__nv_bfloat16*>(planes), ldp);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}
