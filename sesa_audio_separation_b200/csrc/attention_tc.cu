// Flash-style attention on the Blackwell tensor cores (tcgen05 + TMEM), fed by TMA.
//
// Replaces Attend.forward (models/bs_roformer/attend.py:89-93,113-126: softmax(q k^T) v, no mask, non-causal)
// together with the sigmoid gating and head merge of Attention.forward (models/bs_roformer/bs_roformer.py:115-120).
// q/k/v arrive as bf16 hi/lo planes written by the to_qkv GEMM epilogue (q pre-scaled by dh^-0.5, q and k already
// rotated); the gated output leaves as bf16 hi/lo planes, i.e. directly as the A operand of the to_out GEMM.
//
// The residual stream is token-major [(b t f), d], so the axial "rearranges" of bs_roformer.py:526-543 are TMA
// tensor-map strides: a time sequence (b, f) is a strided walk over t, a band sequence (b, t) is contiguous.
//   STRIDED mode: one sequence per CTA tile: 128 query rows x KV blocks of 64 keys (time attention, seq 801).
//   PACKED  mode: short contiguous sequences (seq_len <= 64) are packed floor(128/seq_len) per 128-row tile with a
//                 block-diagonal mask (band attention, seq 62 -> 2 sequences per tile).
// Per KV block: S = Q K^T (UMMA 128x64x64) -> TMEM; the 128 softmax threads (thread = row) read S, do the online
// softmax in fp32, write P as packed bf16 hi/lo back into TENSOR MEMORY (tcgen05.st); O_blk = P V (UMMA 128x64x64 with
// A read from TMEM and V as the MN-major shared-memory operand, K/V blocks double-buffered by TMA) accumulates into ONE
// fp32 O tile that stays in tensor memory for the whole sequence.  The exponent offset of a row only follows the running
// maximum when it grew by more than 2^8 ("lazy rescaling"): then, and only then, the threads of that warp pair rescale
// their O columns in TMEM; in the common case a block costs the softmax threads no accumulator work at all.
// fp32 parity uses the same three-product split as the GEMM (hi.hi + hi.lo + lo.hi).  Two CTAs are resident per SM so one CTA's softmax
// overlaps the other's MMAs.
#include "common.cuh"
#include "sesa_b200.h"
#include "tc_common.cuh"

namespace {

constexpr int DH = 64;
constexpr int BQ = 128;
constexpr int BKV = 64;
constexpr int SM_THREADS = 256;   // warps 0-7: softmax; thread (row = (warp&3)*32+lane, half = warp>>2)
constexpr int ATT_THREADS = SM_THREADS + 64;  // warp 8: MMA issue, warp 9: TMA producer
constexpr float LOG2E = 1.4426950408889634f;

struct AttnParams {
  const float* gates;   // [rows][ldg] gate logits
  __nv_bfloat16* out;   // [planes][rows][ldo]
  int64_t ldg, ldo, out_plane;
  int heads, inner;     // inner = heads * DH: k at column inner + h*DH, v at 2*inner + h*DH
  int mode;             // 0 strided, 1 packed
  int seq_len, n_seq;
  int inner_cnt;        // strided: sequence s -> (b = s / inner_cnt, f = s % inner_cnt)
  int64_t outer_stride, inner_stride, pos_stride;  // strided: row = b*outer + f*inner + pos*pos_stride
  int spt;              // packed: sequences per tile
  int seq_group;        // packed: sequences are packed within groups of this many (one group per chunk), so a
  int tiles_per_group;  //         sequence's arithmetic never depends on the batch it is launched in
  int q_tiles;          // strided: 128-row query tiles per sequence
  int out_planes;
  int total_work;        // work items (row tiles x heads) of the launch
};

// 64-thread named barrier of the warp pair (w, w+4) that shares a TMEM lane quadrant (ids 1..4, compile-time ids keep
// the kernel's barrier count at 5)
__device__ __forceinline__ void pair_barrier(int quadrant) {
  switch (quadrant) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}

template <int NSPLIT>
struct ACfg {
  static constexpr int NP = NSPLIT == 3 ? 2 : 1;
  static constexpr int Q_BYTES = BQ * 128;    // one plane of the Q tile (128 rows x 64 bf16)
  static constexpr int KV_BYTES = BKV * 128;  // one plane of a K or V block
  static constexpr int KV_STAGES = 2;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + NP * Q_BYTES;
  static constexpr int OFF_V = OFF_K + KV_STAGES * NP * KV_BYTES;
  static constexpr int OFF_X = OFF_V + KV_STAGES * NP * KV_BYTES;   // float xch[3][2][128]: row max (2 slots) / row sum exchange
  static constexpr int OFF_BAR = OFF_X + 3 * 2 * BQ * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  // TMEM columns: SP[0] [0,64) | SP[1] [64,128) | O [128,192).
  // SP[b] holds the fp32 scores of block j (b = j&1) and is overwritten IN PLACE by P (bf16 pairs: hi in the first
  // 32 columns, lo in the last 32) once both threads of a row have read their scores.
  static constexpr int TMEM_COLS = 256;
};

// work item w -> (head, row tile): heads fastest, so that the CTAs working on the 8 heads of one row tile run together and
// the 256-byte L2 fetches around each 128-byte head slice are shared instead of being fetched twice from HBM
struct WorkItem {
  int h, item, q0, c1, c3;
  int64_t row_base;
};
__device__ __forceinline__ WorkItem decode_work(const AttnParams& p, int w) {
  WorkItem it;
  it.h = w % p.heads;
  it.item = w / p.heads;
  it.q0 = 0; it.c1 = 0; it.c3 = 0; it.row_base = 0;
  if (p.mode == 0) {
    const int seq = it.item / p.q_tiles;
    it.q0 = (it.item - seq * p.q_tiles) * BQ;
    it.c1 = seq % p.inner_cnt;
    it.c3 = seq / p.inner_cnt;
  } else {
    const int grp = it.item / p.tiles_per_group;
    const int lt = it.item - grp * p.tiles_per_group;
    it.row_base = ((int64_t)grp * p.seq_group + (int64_t)lt * p.spt) * p.seq_len;
  }
  return it;
}

// Persistent CTAs (two per SM): each walks work items w = blockIdx.x, blockIdx.x + gridDim.x, ... and all mbarrier phases run
// on a block counter g that continues across items, so the Q/K/V loads and the first two score MMAs of item n+1 are issued
// while the softmax threads finish item n: the tensor pipe does not drain at item boundaries and the barrier / TMEM set-up
// is paid once per CTA instead of once per tile.
template <int NSPLIT>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap map, AttnParams p) {
  using C = ACfg<NSPLIT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem + C::OFF_Q;
  uint8_t* sK = smem + C::OFF_K;
  uint8_t* sV = smem + C::OFF_V;
  float* xch = reinterpret_cast<float*>(smem + C::OFF_X);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars + 0;    // TMA -> MMA, once per item
  uint64_t* q_empty = bars + 1;   // MMA (commit after the item's last S) -> TMA
  uint64_t* k_full = bars + 2;    // [2] TMA -> MMA
  uint64_t* k_empty = bars + 4;   // [2] MMA (commit) -> TMA
  uint64_t* v_full = bars + 6;    // [2]
  uint64_t* v_empty = bars + 8;   // [2]
  uint64_t* s_full = bars + 10;   // [2] MMA (commit) -> softmax
  uint64_t* p_full = bars + 12;   // [2] softmax (256 arrivals) -> MMA
  uint64_t* o_full = bars + 14;   // [2] MMA (commit) -> softmax
  uint64_t* o_empty = bars + 16;  // softmax (256 arrivals: O of the item has been read) -> MMA, once per item
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 17);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;

  if (tid == SM_THREADS) {
    tc::mbar_init(q_full, 1);
    tc::mbar_init(q_empty, 1);
    tc::mbar_init(o_empty, SM_THREADS);
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&k_full[b], 1);
      tc::mbar_init(&k_empty[b], 1);
      tc::mbar_init(&v_full[b], 1);
      tc::mbar_init(&v_empty[b], 1);
      tc::mbar_init(&s_full[b], 1);
      tc::mbar_init(&p_full[b], SM_THREADS);
      tc::mbar_init(&o_full[b], 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 8) {
    tc::tmem_alloc(tmem_ptr, C::TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_SP = tmem_base;         // + 64*b
  const uint32_t tmem_O = tmem_base + 128;

  const int nblk = p.mode == 0 ? (p.seq_len + BKV - 1) / BKV : BQ / BKV;   // KV blocks per item
  const int total = p.total_work;

  if (warp == 9) {
    // ================= TMA producer =================
    if (tc::elect_one()) {
      uint32_t G = 0;   // block counter of the item's first block
      int n = 0;        // items done by this CTA
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++n, G += nblk) {
        const WorkItem it = decode_work(p, w);
        auto load_rows = [&](uint8_t* dst, uint64_t* bar, int col, int r0, int plane) {
          // 64 rows x 64 columns of one plane (rows beyond the tensor are zero-filled by TMA)
          if (p.mode == 0) tc::tma_load_5d(dst, &map, bar, col, it.c1, r0, it.c3, plane);
          else tc::tma_load_5d(dst, &map, bar, col, (int)(it.row_base + r0), 0, 0, plane);
        };
        const int colq = it.h * DH, colk = p.inner + it.h * DH, colv = 2 * p.inner + it.h * DH;
        auto load_k = [&](int j) {
          const uint32_t g = G + j;
          const int st = g & 1;
          tc::mbar_wait(&k_empty[st], ((g >> 1) & 1) ^ 1);   // S_{g-2} retired (passes at once for g < 2)
          tc::mbar_expect_tx(&k_full[st], C::NP * C::KV_BYTES);
#pragma unroll
          for (int pl = 0; pl < C::NP; ++pl) load_rows(sK + (st * C::NP + pl) * C::KV_BYTES, &k_full[st], colk, j * BKV, pl);
        };
        auto load_v = [&](int j) {
          const uint32_t g = G + j;
          const int st = g & 1;
          tc::mbar_wait(&v_empty[st], ((g >> 1) & 1) ^ 1);   // P_{g-2} V_{g-2} retired
          tc::mbar_expect_tx(&v_full[st], C::NP * C::KV_BYTES);
#pragma unroll
          for (int pl = 0; pl < C::NP; ++pl) load_rows(sV + (st * C::NP + pl) * C::KV_BYTES, &v_full[st], colv, j * BKV, pl);
        };
        if (n > 0) tc::mbar_wait(q_empty, (n - 1) & 1);       // every S of the previous item has read Q
        tc::mbar_expect_tx(q_full, C::NP * C::Q_BYTES);
#pragma unroll
        for (int pl = 0; pl < C::NP; ++pl) {
          load_rows(sQ + pl * C::Q_BYTES, q_full, colq, it.q0, pl);
          load_rows(sQ + pl * C::Q_BYTES + C::KV_BYTES, q_full, colq, it.q0 + 64, pl);
        }
        // issue order follows the order in which the tensor pipe frees the stages:
        // S_0 S_1 PV_0 S_2 PV_1 S_3 ...  =>  K_0 V_0 K_1 V_1 K_2 | K_3 V_2 | K_4 V_3 | ...
        load_k(0);
        load_v(0);
        if (nblk > 1) { load_k(1); load_v(1); }
        if (nblk > 2) load_k(2);
        for (int j = 2; j < nblk; ++j) {
          if (j + 1 < nblk) load_k(j + 1);
          load_v(j);
        }
      }
    }
  } else if (warp == 8) {
    // ================= MMA issuer =================
    if (tc::elect_one()) {
      constexpr uint32_t idesc_s = tc::make_idesc_bf16(BQ, BKV, 0, 0);  // S = Q K^T : both K-major
      constexpr uint32_t idesc_o = tc::make_idesc_bf16(BQ, DH, 0, 1);   // O = P V   : P from TMEM, V MN-major
      const uint32_t aQ = tc::smem_u32(sQ), aK = tc::smem_u32(sK), aV = tc::smem_u32(sV);
      // The CTA's blocks are numbered g = 0, 1, ... across its items (item n = g / nblk, in-item index j = g % nblk).  The
      // tensor pipe runs S_0 S_1 | PV_0 S_2 | PV_1 S_3 | ...: scores run two blocks ahead of P V, also across item
      // boundaries (nblk may be 1).
      const int n_items = blockIdx.x < total ? (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
      const uint32_t GB = (uint32_t)n_items * (uint32_t)nblk;
      // Q lives in tensor memory for the whole item (columns [192, 256): hi plane 32 columns, lo plane 32): it is copied
      // there once per item (tcgen05.cp, in issue order with the MMAs), so the score MMAs read their A operand from TMEM
      // like P V does and only K crosses the shared-memory port (an SS-form 128x64x16 MMA fetches 6 KB per 32 cycles:
      // more than the 128 B/clk the port delivers)
      const uint32_t tmem_Q = tmem_base + 192;
      auto issue_s = [&](uint32_t g) {
        const int n = (int)(g / (uint32_t)nblk);
        const int j = (int)(g - (uint32_t)n * (uint32_t)nblk);
        const int st = g & 1;
        if (j == 0) {
          tc::mbar_wait(q_full, n & 1);
          tc::tc_fence_after();
#pragma unroll
          for (int pl = 0; pl < C::NP; ++pl)
#pragma unroll
            for (int ks = 0; ks < DH / 16; ++ks)
              tc::tmem_cp_128x256b(tmem_Q + pl * 32 + ks * 8, tc::make_smem_desc_sw128(aQ + pl * C::Q_BYTES + ks * 32));
          tc::umma_commit(q_empty);   // the Q tile in shared memory may be overwritten once the copies retire
        }
        tc::mbar_wait(&k_full[st], (g >> 1) & 1);
        tc::tc_fence_after();
#pragma unroll
        for (int prod = 0; prod < NSPLIT; ++prod) {
          const uint32_t qa = tmem_Q + (prod == 2 ? 32 : 0);
          const uint32_t ka = aK + (st * C::NP + (prod == 1 ? 1 : 0)) * C::KV_BYTES;
#pragma unroll
          for (int ks = 0; ks < DH / 16; ++ks)
            tc::umma_f16_ts(tmem_SP + st * 64, qa + ks * 8, tc::make_smem_desc_sw128(ka + ks * 32), idesc_s,
                            (prod | ks) != 0 ? 1u : 0u);
        }
        tc::umma_commit(&s_full[st]);
        tc::umma_commit(&k_empty[st]);
      };
      if (GB > 0) issue_s(0);
      if (GB > 1) issue_s(1);
      int n = 0, j = 0;
      for (uint32_t g = 0; g < GB; ++g) {
        const int st = g & 1;
        const uint32_t ph = (g >> 1) & 1;
        tc::mbar_wait(&p_full[st], ph);       // P_g is in TMEM (S_g consumed); any rescale of O is complete
        tc::mbar_wait(&v_full[st], ph);
        if (j == 0 && n > 0) tc::mbar_wait(o_empty, (n - 1) & 1);   // the previous item's O tile has been read
        tc::tc_fence_after();
#pragma unroll
        for (int prod = 0; prod < NSPLIT; ++prod) {
          const uint32_t pa = tmem_SP + st * 64 + (prod == 2 ? 32 : 0);
          const uint32_t va = aV + (st * C::NP + (prod == 1 ? 1 : 0)) * C::KV_BYTES;
#pragma unroll
          for (int ks = 0; ks < BKV / 16; ++ks)   // 16 keys = 8 packed TMEM columns of P, 16 rows (2 KB) of V
            tc::umma_f16_ts(tmem_O, pa + ks * 8, tc::make_smem_desc_sw128(va + ks * 2048, 8192), idesc_o,
                            (j | prod | ks) != 0 ? 1u : 0u);
        }
        tc::umma_commit(&o_full[st]);
        tc::umma_commit(&v_empty[st]);
        // scores two blocks ahead overwrite SP[st] after P_g V_g (the tensor pipe runs in issue order)
        if (g + 2 < GB) issue_s(g + 2);
        if (++j == nblk) { j = 0; ++n; }
      }
    }
  } else {
    // ================= softmax threads =================
    // Two threads per query row: thread (row i, half hf) owns keys [32 hf, 32 hf + 32) of every 64-key block and
    // output dims [32 hf, 32 hf + 32).  The row maximum is agreed through shared memory once per block.  O accumulates in
    // tensor memory; the threads touch it only to rescale (rare) and to read the finished tile.
    const int hf = warp >> 2;
    const int i = (warp & 3) * 32 + (tid & 31);   // row of the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    constexpr int HC = BKV / 2;   // 32 score columns / output dims per thread
    uint32_t G = 0;
    int n = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++n, G += nblk) {
      const WorkItem it = decode_work(p, w);
      int lo = 0, hi = p.seq_len;                   // valid key range (tile-relative key index)
      bool row_valid;
      int64_t out_row;
      if (p.mode == 0) {
        row_valid = it.q0 + i < p.seq_len;
        out_row = (int64_t)it.c3 * p.outer_stride + (int64_t)it.c1 * p.inner_stride + (int64_t)(it.q0 + i) * p.pos_stride;
      } else {
        const int sl = i / p.seq_len;
        const int grp = it.item / p.tiles_per_group;
        const int lseq = (it.item - grp * p.tiles_per_group) * p.spt + sl;   // sequence index inside its group
        row_valid = sl < p.spt && lseq < p.seq_group && (int64_t)grp * p.seq_group + lseq < p.n_seq;
        lo = row_valid ? sl * p.seq_len : 0;
        hi = row_valid ? lo + p.seq_len : 0;
        out_row = it.row_base + i;
      }
      // the gate logit of this (row, head) is only needed at the very end: request it now so its latency is never exposed
      const float gl = row_valid ? __ldg(p.gates + out_row * p.ldg + it.h) : 0.f;
      float mref = -INFINITY, lrun = 0.f;   // exponent reference of the row (lags the true maximum by at most 2^8), row sum
      // a warp whose 32 rows all lie beyond the sequence (the last query tile of 801 = 6 x 128 + 33 rows) only keeps the
      // barrier protocol going: its P rows are never stored, so their content does not matter
      const bool warp_active = __any_sync(0xffffffffu, row_valid);

      for (int j = 0; j < nblk; ++j) {
        const uint32_t g = G + j;
        const int st = g & 1;
        tc::mbar_wait(&s_full[st], (g >> 1) & 1);
        if (!warp_active) {
          tc::mbar_arrive(&p_full[st]);
          continue;
        }
        tc::tc_fence_after();
        float s[HC];
        tc::tmem_ld32(tmem_SP + st * 64 + lane_off + hf * HC, s);
        tc::tmem_ld_wait();
        const int k0 = j * BKV + hf * HC;
        float mx = -INFINITY;     // row maximum of the RAW scores; the log2(e) factor is folded into the exp2 argument FMA
        if (__all_sync(0xffffffffu, k0 >= lo && k0 + HC <= hi)) {
#pragma unroll
          for (int c = 0; c < HC; ++c) mx = fmaxf(mx, s[c]);
        } else {
#pragma unroll
          for (int c = 0; c < HC; ++c) {
            const int kj = k0 + c;
            s[c] = (kj >= lo && kj < hi) ? s[c] : -INFINITY;
            mx = fmaxf(mx, s[c]);
          }
        }
        float* xm = xch + (st * 2) * BQ;
        xm[hf * BQ + i] = mx;
        tc::tc_fence_before();
        // the two threads of a row sit in warps w and w+4: a 64-thread named barrier per warp pair (also orders
        // "both halves of every row have read their scores" before P overwrites them in place)
        pair_barrier(warp & 3);
        tc::tc_fence_after();
        const float mnew = fmaxf(mref, fmaxf(mx, xm[(hf ^ 1) * BQ + i]));
        // lazy rescaling: keep the old reference while exp2 arguments stay below 8 (p < 256: harmless for the bf16 split
        // and the fp32 accumulators).  Both threads of a row take the same decision from the same numbers.
        const bool grow = (mnew - mref) * LOG2E > 8.0f;   // also true for the first finite maximum (mref = -inf)
        if (__any_sync(0xffffffffu, grow)) {
          const float corr = grow ? tc::ex2_approx((mref - mnew) * LOG2E) : 1.0f;   // exp2(-inf) = 0 on the first block
          if (j > 0) {
            // rows of this warp moved their reference: rescale our 32 columns of O once every earlier P V has retired
            tc::mbar_wait(&o_full[(g - 1) & 1], ((g - 1) >> 1) & 1);
            tc::tc_fence_after();
            float ob[HC];
            tc::tmem_ld32(tmem_O + lane_off + hf * HC, ob);
            tc::tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < HC; ++d) ob[d] *= corr;
            tc::tmem_st32(tmem_O + lane_off + hf * HC, reinterpret_cast<const uint32_t*>(ob));
          }
          lrun *= corr;
          if (grow) mref = mnew;
        }
        const float moff = (mref == -INFINITY ? 0.f : mref) * LOG2E;
        float2 sum2 = make_float2(0.f, 0.f);
        // P leaves for tensor memory in two halves of 16 keys (8 packed registers per plane): keeps the live registers of
        // the loop under the 96 the two-CTA residency allows
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t phi[HC / 4], plo[HC / 4];
#pragma unroll
          for (int e = 0; e < HC / 4; ++e) {
            // packed fp32x2 arithmetic (FFMA2 / FADD2): exponent argument, row sum and hi/lo residual of a key pair
            const float2 a = __ffma2_rn(make_float2(s[hh * (HC / 2) + 2 * e], s[hh * (HC / 2) + 2 * e + 1]),
                                        make_float2(LOG2E, LOG2E), make_float2(-moff, -moff));
            const float2 pp = make_float2(tc::ex2_approx(a.x), tc::ex2_approx(a.y));
            sum2 = __fadd2_rn(sum2, pp);
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(phi[e]) : "f"(pp.y), "f"(pp.x));
            const float2 hf = make_float2(__uint_as_float(phi[e] << 16), __uint_as_float(phi[e] & 0xffff0000u));
            const float2 r = __ffma2_rn(hf, make_float2(-1.0f, -1.0f), pp);      // p - hi(p), exact
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(plo[e]) : "f"(r.y), "f"(r.x));
          }
          tc::tmem_st8(tmem_SP + st * 64 + lane_off + hf * (HC / 2) + hh * (HC / 4), phi);
          if (NSPLIT == 3) tc::tmem_st8(tmem_SP + st * 64 + lane_off + 32 + hf * (HC / 2) + hh * (HC / 4), plo);
        }
        tc::tmem_st_wait();
        lrun += sum2.x + sum2.y;
        tc::tc_fence_before();
        tc::mbar_arrive(&p_full[st]);
      }
      // the finished O tile (every warp waits for the last P V, so that no thread can run a whole item ahead of o_empty)
      const uint32_t gl_last = G + nblk - 1;
      float o[HC];
      tc::mbar_wait(&o_full[gl_last & 1], (gl_last >> 1) & 1);
      tc::tc_fence_after();
      tc::tmem_ld32(tmem_O + lane_off + hf * HC, o);   // unconditional: a conditionally written o[] would stay live
      tc::tmem_ld_wait();                              // across the whole item loop (32 registers)
      tc::tc_fence_before();
      tc::mbar_arrive(o_empty);     // the next item's first P V may overwrite O

      // combine the two halves' row sums (own exchange slot: slots 0/1 belong to the per-block row maxima), gate, split to
      // planes, store this thread's 32 output dims
      float* xs = xch + 4 * BQ;
      xs[hf * BQ + i] = lrun;
      pair_barrier(warp & 3);
      const float ltot = lrun + xs[(hf ^ 1) * BQ + i];
      if (row_valid) {
        const float sc = (1.0f / (1.0f + expf(-gl))) / ltot;
        __nv_bfloat16* op = p.out + out_row * p.ldo + it.h * DH + hf * HC;
#pragma unroll
        for (int d8 = 0; d8 < HC / 8; ++d8) {
          uint32_t hh[4], ll[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) tc::split_bf16x2(o[d8 * 8 + 2 * e] * sc, o[d8 * 8 + 2 * e + 1] * sc, hh[e], ll[e]);
          *reinterpret_cast<uint4*>(op + d8 * 8) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
          if (p.out_planes > 1) *reinterpret_cast<uint4*>(op + p.out_plane + d8 * 8) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
        }
      }
      // the slot is rewritten by the next item's exchange only after its own pair barrier sequence: every reader of this
      // item's sums has passed the barrier above and at least one more (per-block) barrier by then
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int NSPLIT>
int launch_attention_tc(const CUtensorMap& map, const AttnParams& p, cudaStream_t stream) {
  using C = ACfg<NSPLIT>;
  static bool configured = false;
  static int sms = 0;
  if (!configured) {
    SESA_CUDA(cudaFuncSetAttribute(attention_tc_kernel<NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    int dev = 0;
    SESA_CUDA(cudaGetDevice(&dev));
    SESA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    configured = true;
  }
  const int resident = 2 * sms;   // persistent CTAs: two per SM (TMEM 2 x 256 columns, ~97 KB of shared memory each)
  const unsigned grid = (unsigned)(p.total_work < resident ? p.total_work : resident);
  attention_tc_kernel<NSPLIT><<<grid, ATT_THREADS, C::SMEM_BYTES, stream>>>(map, p);
  SESA_LAUNCH_CHECK();
  return SESA_OK;
}

}  // namespace

extern "C" int sesa_attention_tc(const void* qkv_planes, int64_t ld, int64_t plane_stride, const float* gates,
                                 int64_t ldg, void* out_planes_ptr, int64_t ldo, int64_t out_plane_stride, int heads,
                                 int dim_head, int n_seq, int seq_len, int inner_cnt, int64_t outer_stride,
                                 int64_t inner_stride, int64_t pos_stride, int seq_group, int nsplit, int out_planes,
                                 void* stream) {
  SESA_CHECK_ARG(dim_head == DH, "sesa_attention_tc: dim_head must be 64, got %d", dim_head);
  SESA_CHECK_ARG((ld & 7) == 0 && (ldo & 7) == 0 && (plane_stride & 7) == 0 && (out_plane_stride & 7) == 0,
                 "sesa_attention_tc: plane strides must be multiples of 8 bf16 elements");
  SESA_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv_planes) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_planes_ptr) & 15) == 0,
                 "sesa_attention_tc: plane buffers must be 16-byte aligned");
  SESA_CHECK_ARG(inner_cnt > 0 && seq_len > 0 && heads > 0, "sesa_attention_tc: bad sequence geometry");
  SESA_CHECK_ARG(nsplit == 1 || nsplit == 3, "sesa_attention_tc: nsplit must be 1 or 3");
  SESA_CHECK_ARG(out_planes == 1 || out_planes == 2, "sesa_attention_tc: out_planes must be 1 or 2");
  if (n_seq == 0) return SESA_OK;
  const int inner = heads * DH;
  AttnParams p;
  p.gates = gates;
  p.out = reinterpret_cast<__nv_bfloat16*>(out_planes_ptr);
  p.ldg = ldg;
  p.ldo = ldo;
  p.out_plane = out_plane_stride;
  p.heads = heads;
  p.inner = inner;
  p.seq_len = seq_len;
  p.n_seq = n_seq;
  p.inner_cnt = inner_cnt;
  p.outer_stride = outer_stride;
  p.inner_stride = inner_stride;
  p.pos_stride = pos_stride;
  p.out_planes = out_planes;
  p.spt = 1;
  p.seq_group = n_seq;
  p.tiles_per_group = 1;
  p.q_tiles = 1;
  CUtensorMap map;
  const bool packed = pos_stride == 1 && seq_len <= BKV && inner_cnt == 1 && outer_stride == seq_len;
  const uint32_t box[5] = {64, 1, 64, 1, 1};
  if (packed) {
    // rows are one contiguous run of n_seq*seq_len tokens
    p.mode = 1;
    p.spt = BQ / seq_len;
    p.seq_group = seq_group > 0 ? seq_group : n_seq;
    SESA_CHECK_ARG(n_seq % p.seq_group == 0, "sesa_attention_tc: n_seq must be a multiple of seq_group");
    p.tiles_per_group = (p.seq_group + p.spt - 1) / p.spt;
    const uint64_t rows = (uint64_t)n_seq * seq_len;
    const uint64_t dims[5] = {(uint64_t)3 * inner, rows, 1, 1, 2};
    const uint64_t str[4] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * rows, (uint64_t)ld * 2 * rows, (uint64_t)plane_stride * 2};
    const uint32_t boxp[5] = {64, 64, 1, 1, 1};
    int rc = sesa_make_tmap_bf16(&map, qkv_planes, 5, dims, str, boxp);
    if (rc != SESA_OK) return rc;
    SESA_CHECK_ARG((int64_t)(n_seq / p.seq_group) * p.tiles_per_group * heads < (1LL << 31), "sesa_attention_tc: too many tiles for one launch");
    p.total_work = (int)((int64_t)(n_seq / p.seq_group) * p.tiles_per_group * heads);
  } else {
    // sequence s = (b, f): row = b*outer + f*inner + pos*pos_stride
    p.mode = 0;
    const int n_outer = n_seq / inner_cnt;
    SESA_CHECK_ARG(n_outer * inner_cnt == n_seq, "sesa_attention_tc: n_seq must be a multiple of inner_cnt");
    const uint64_t dims[5] = {(uint64_t)3 * inner, (uint64_t)inner_cnt, (uint64_t)seq_len, (uint64_t)n_outer, 2};
    const uint64_t str[4] = {(uint64_t)(inner_cnt > 1 ? inner_stride : 1) * ld * 2, (uint64_t)pos_stride * ld * 2,
                             (uint64_t)outer_stride * ld * 2, (uint64_t)plane_stride * 2};
    int rc = sesa_make_tmap_bf16(&map, qkv_planes, 5, dims, str, box);
    if (rc != SESA_OK) return rc;
    p.q_tiles = (seq_len + BQ - 1) / BQ;
    SESA_CHECK_ARG((int64_t)n_seq * p.q_tiles * heads < (1LL << 31), "sesa_attention_tc: too many tiles for one launch");
    p.total_work = (int)((int64_t)n_seq * p.q_tiles * heads);
  }
  if (nsplit == 3) return launch_attention_tc<3>(map, p, (cudaStream_t)stream);
  return launch_attention_tc<1>(map, p, (cudaStream_t)stream);
}
