/*
 * sesa_b200 — C ABI of the B200-native chunked-separation hot path.
 *
 * Drop-in boundary for the path behind SESA's demix() (reference: /root/reference/utils.py:330-477,
 * inference_pytorch.py:55-186) and the BS-RoFormer / Mel-Band-RoFormer / MDX23C forward passes
 * (models/bs_roformer/bs_roformer.py:447-587, models/bs_roformer/mel_band_roformer.py:480-633,
 * models/mdx23c_tfc_tdf_v3.py:205-242).  The reference is pure Python on top of PyTorch library
 * kernels, so each entry point replaces a LIBRARY CALL SITE of the reference (cited per function);
 * the Python host in sesa_audio_separation_b200/ mirrors the reference's operator surface and binds
 * these symbols with ctypes.
 *
 * Conventions: every pointer is a DEVICE pointer unless its name ends in _host; sizes are elements;
 * `stream` is a cudaStream_t passed as void*; nothing is allocated, freed or synchronised by the
 * library; every function returns 0 on success or a SESA_ERR_* code, with a message retrievable
 * from sesa_last_error().  There is no CPU fallback: without a CUDA device every compute entry
 * point fails with SESA_ERR_CUDA.
 */
#ifndef SESA_B200_H
#define SESA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SESA_B200_ABI_VERSION 3
#define SESA_TC_SS_SLOTS_PER_BLOCK 2 /* row-sum-of-squares slots a GEMM writes per column block (one per epilogue warp of a TMEM quadrant) */

enum { SESA_ACT_NONE = 0, SESA_ACT_GELU = 1, SESA_ACT_TANH = 2, SESA_ACT_SIGMOID = 3 };

/* One problem of a grouped GEMM  C[M,N] = epi(A[M,K] . W[N,K]^T)  (W in nn.Linear layout). */
typedef struct sesa_gemm_group {
  const float* A;
  const float* W;
  const float* bias; /* [N] or NULL */
  float* C;
  int32_t M, N, K, _pad;
  int64_t lda, ldw, ldc;
} sesa_gemm_group;

/* Epilogue shared by all groups of a launch. */
typedef struct sesa_gemm_epilogue {
  int32_t rownorm;  /* scale row m by 1/max(||A[m,:]||_2, 1e-12): fused RMSNorm (bs_roformer.py:43-50),
                       gamma*sqrt(K) being folded into W by the host */
  int32_t act;      /* SESA_ACT_*: GELU(erf) bs_roformer.py:66, tanh :271, sigmoid */
  int32_t residual; /* C += result (bs_roformer.py:214-215) */
  int32_t glu;      /* nn.GLU (bs_roformer.py:296): W rows interleaved (value,gate); writes N/2 columns */
  int32_t rot_cols; /* rotary embedding on column pairs n < rot_cols (bs_roformer.py:112-113) */
  int32_t rot_dim;  /* head dim */
  int32_t pos_div, pos_mod; /* sequence position of row m = (m / pos_div) % pos_mod */
  const float* rot; /* sesa_gemm_simt: [pos_mod][rot_dim/2][2] = (cos, sin) per position and column pair;
                       sesa_gemm_tc: the same values quad-major, [rot_dim/4][pos_mod][4] = (cos, sin, cos, sin) of the two
                       pairs of a column quad, so that rows at consecutive positions read consecutive 16-byte words */
} sesa_gemm_epilogue;

int sesa_abi_version(void);
const char* sesa_last_error(void);
/* 0 and *sm_count etc filled when a CUDA device is usable. */
int sesa_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem);

/* ---- framing -------------------------------------------------------------------------------- */
/* dst[c][i] = src[c][reflect(i-left)], i in [0, len+left+right): the border reflect pad of
 * utils.py:391-393 (nn.functional.pad(mix, (border, border), mode="reflect")). */
int sesa_pad_reflect(const float* src, float* dst, int channels, int64_t len, int64_t left, int64_t right,
                     void* stream);
/* chunks[k][c][0..L) = mix[c][start_k .. start_k+len_k) right-padded to L by reflection (mode 1) or zeros
 * (mode 0): utils.py:413-425.  starts/lens/modes are device arrays of n_chunks entries. */
int sesa_frame_chunks(const float* mix, int64_t mix_len, int channels, const int64_t* starts,
                      const int64_t* lens, const int32_t* modes, int n_chunks, int64_t chunk_size,
                      float* chunks, void* stream);

/* ---- STFT / iSTFT --------------------------------------------------------------------------- */
/* torch.stft(center=True, pad_mode='reflect', onesided) of n_signals groups of `channels` signals
 * (bs_roformer.py:485, mel_band_roformer.py:516, mdx23c_tfc_tdf_v3.py:19-26).
 * layout 0: spec[(sig*T+t)][f][c][re,im]; layout 1: spec[sig][c][re,im][f<dim_f][t].
 * window: [n_fft]; twiddle: [n_fft][2] = exp(-2 pi i k / n_fft). */
int sesa_stft(const float* audio, float* spec, const float* window, const float* twiddle, int n_signals,
              int channels, int64_t length, int n_fft, int hop, int layout, int dim_f, void* stream);
/* complex mask multiply (bs_roformer.py:556-567; Mel scatter-average mel_band_roformer.py:603-616) fused with
 * torch.istft (bs_roformer.py:575): irFFT, window, overlap-add over frames, / window envelope, trim.
 * mode 0: mask[n][b*T+t][f][c][2]; mode 1: mask[n][b*T+t][J][2] + inv_index[(f*C+c)*2+{0,1}], inv_count[f*C+c];
 * mode 2: no mask, spec[(b*nstems+n)][t][f][c][2].  out[b][n][c][out_len] (contiguous, fully overwritten; for
 * n_fft 2048 it is zero-filled on the stream first and accumulated with order-independent two-term atomics, so the
 * result is bitwise reproducible and independent of the batch); envelope[out_len]. */
int sesa_mask_istft(const float* spec, const float* mask, const int32_t* inv_index, const float* inv_count,
                    float* out, const float* window, const float* envelope, const float* twiddle, int batch,
                    int nstems, int channels, int n_fft, int hop, int n_frames, int64_t out_len, int mode,
                    int n_gathered, void* stream);

/* ---- dense math ----------------------------------------------------------------------------- */
/* Exact-fp32 grouped GEMM (nn.Linear call sites bs_roformer.py:63,67,99,101,104,237,264). */
int sesa_gemm_simt(const sesa_gemm_group* groups_dev, int n_groups, int max_m, int max_n,
                   const sesa_gemm_epilogue* ep_host, void* stream);
/* softmax(q k^T) v, sigmoid gating, head merge (attend.py:89-93,113-126; bs_roformer.py:115-120).
 * Sequence s covers rows base(s) + p*pos_stride, base(s) = (s / inner_cnt)*outer_stride + (s % inner_cnt)*inner_stride.
 * qkv row: [q | k | v | gate logits], ld floats; out row: [h d], ldo floats. */
int sesa_attention_simt(const float* qkv, float* out, int ld, int ldo, int heads, int dim_head, int n_seq,
                        int seq_len, int inner_cnt, int64_t outer_stride, int64_t inner_stride,
                        int64_t pos_stride, void* stream);
/* y[r,:] = x[r,:]/max(||x[r,:]||,1e-12) * sqrt(dim) * gamma   (RMSNorm, bs_roformer.py:43-50). In place ok. */
int sesa_rmsnorm(const float* x, const float* gamma, float* y, int64_t rows, int dim, void* stream);
/* y += x (skip connections, bs_roformer.py:521-524) */
int sesa_add_inplace(float* y, const float* x, int64_t n, void* stream);
/* out[r][j][0..w) = in[r][idx[j]][0..w)  (Mel band gather, mel_band_roformer.py:530); w floats per entry */
int sesa_gather_rows(const float* in, const int32_t* idx, float* out, int64_t rows, int n_in, int n_out,
                     int width, void* stream);

/* ---- tensor-core (tcgen05 / TMEM / TMA) dense math ------------------------------------------ */
/* Operands of the tensor-core kernels are bf16 "planes": plane 0 = bf16(x) ("hi"), plane 1 = bf16(x - hi) ("lo").
 * nsplit = 3 evaluates A.W^T as Ahi.Whi + Ahi.Wlo + Alo.Whi with fp32 accumulation in TMEM (~16 significant
 * bits per operand: the fp32-parity mode of BASELINE.json); nsplit = 1 uses the hi planes only (bf16 mode). */
typedef struct sesa_tc_problem {
  const void* A;         /* bf16 [planes][M][lda]; lda % 8 == 0, 16-byte aligned */
  const void* W;         /* bf16 [planes][N][ldw] (nn.Linear layout); ldw % 8 == 0 */
  const float* bias;     /* [N] or NULL */
  const float* rowscale; /* [M] or NULL: acc[m,:] *= rowscale[m] before the bias (fused RMSNorm) */
  float* C;              /* fp32 output [M][ldc] or NULL */
  void* P;               /* bf16 plane output [out_planes][M][ldp] or NULL (feeds the next tensor-core op) */
  int64_t lda, a_plane, ldw, w_plane, ldc, ldp, p_plane; /* element strides; *_plane = plane-to-plane distance */
  int32_t M, N, K, _pad;
  /* Implicit-GEMM convolution over a channels-last activation (conv_taps == 0: plain GEMM, fields ignored).
   * A = bf16 planes [planes][B][inT][inF][lda] (lda >= conv_cin channels); output row m = (b*T + t)*F + f;
   * tap i reads input pixel (t*stride + conv_dt[i], f*stride + conv_df[i]), zero outside the grid (the padding of
   * nn.Conv2d, mdx23c_tfc_tdf_v3.py:111,123); W columns are [tap][round_up(conv_cin, 64)] (zero padded), so
   * K = conv_taps * round_up(conv_cin, 64).  F must divide 128 or be a multiple of 128 (tile = 128 pixels). */
  int32_t conv_taps, conv_cin, conv_B, conv_T, conv_F, conv_inT, conv_inF, conv_stride;
  int32_t conv_dt[9], conv_df[9];
  /* Output row remap: 0 = row m; 1 = 2x up-sampling scatter of nn.ConvTranspose2d(kernel = stride = 2)
   * (mdx23c_tfc_tdf_v3.py:80): row = (2*(m / rm_F) + rm_dt) * 2*rm_F + 2*(m % rm_F) + rm_df. */
  int32_t row_map, rm_F, rm_dt, rm_df;
  /* Fused RMSNorm bookkeeping (bs_roformer.py:43-50 applied to the residual stream without a separate pass):
   * ss_out (producer, optional): ss_out[row * SESA_TC_SS_SLOTS_PER_BLOCK*ceil(N/block_n) + slot] receives this launch's per-row partial sums of
   *   squares of the stored values, one slot per (column block, column half) — deterministic, no atomics;
   * rowss (consumer, optional): acc[m,:] *= 1 / max(sqrt(sum_k rowss[m*ss_slots + k]), 1e-12). */
  const float* rowss;
  float* ss_out;
  int32_t ss_slots;
  int32_t p_cols;   /* planes P are written for columns n < p_cols (0 = all columns) */
  int32_t c_col0;   /* C is written for columns n >= c_col0, at column n - c_col0 (to_gates logits next to to_qkv) */
  int32_t ss_ld;    /* floats between consecutive rows of ss_out (0 = SESA_TC_SS_SLOTS_PER_BLOCK * ceil(N/block_n)) */
} sesa_tc_problem;

/* Bytes of the device-side group table for n_groups problems. */
int64_t sesa_gemm_tc_table_bytes(int n_groups);
/* Encode the TMA tensor maps and tile ranges of n_groups problems into table_host (host memory of
 * sesa_gemm_tc_table_bytes(n_groups) bytes); the caller copies the table to the device and keeps it while
 * the pointers stay valid.  *total_tiles receives the number of 128 x block_n output tiles. */
int sesa_gemm_tc_build(const sesa_tc_problem* problems_host, int n_groups, int block_n, int cta_group, void* table_host,
                       int* total_tiles);
/* Grouped GEMM on the 5th-gen tensor cores (nn.Linear call sites bs_roformer.py:63,67,99,104,237,264):
 * TMA-fed, tcgen05.mma into TMEM, persistent over tiles.  Epilogue fields of sesa_gemm_epilogue apply except
 * `rownorm` (use sesa_tc_problem.rowscale).  block_n in {128, 256}; nsplit in {1, 3}; out_planes in {1, 2};
 * cta_group 1 = one CTA per 128 x block_n tile, 2 = a CTA pair (thread-block cluster of 2, tcgen05 cta_group::2) per
 * 256 x 256 tile, each CTA staging half of the W tile (must match the value given to sesa_gemm_tc_build). */
int sesa_gemm_tc(const void* table_dev, int n_groups, int total_tiles, int block_n, int cta_group, int nsplit,
                 int out_planes, const sesa_gemm_epilogue* ep_host, void* stream);
/* Tensor-core attention (attend.py:89-93; gating and head merge bs_roformer.py:115-120).  qkv_planes: bf16
 * [planes][rows][ld] with row layout [q(h d) | k(h d) | v(h d)] (q pre-scaled, q/k rotated); gates: fp32 logits
 * [rows][ldg]; out_planes: bf16 [out_planes][rows][ldo] = softmax(q k^T) v * sigmoid(gate), heads merged.
 * Sequence geometry as sesa_attention_simt.  Short contiguous sequences (seq_len <= 64) are packed per tile; they are
 * packed within groups of seq_group consecutive sequences (0 = all), so that with seq_group = frames per chunk a
 * chunk's result does not depend on the batch it is launched in. */
int sesa_attention_tc(const void* qkv_planes, int64_t ld, int64_t plane_stride, const float* gates, int64_t ldg,
                      void* out_planes_ptr, int64_t ldo, int64_t out_plane_stride, int heads, int dim_head, int n_seq,
                      int seq_len, int inner_cnt, int64_t outer_stride, int64_t inner_stride, int64_t pos_stride,
                      int seq_group, int nsplit, int out_planes, void* stream);
/* Row preparation for the tensor-core GEMMs: per row r of x[rows][dim] (row stride ldx)
 *   inv = normalize == 1 ? 1/max(||x_r||_2, 1e-12) : 1            (F.normalize of RMSNorm, bs_roformer.py:49)
 *   planes[p][r][:] = bf16 split of x_r * inv                      (p < out_planes, row stride ldp)
 *   gates[r][h] = (x_r * inv) . gate_w[h] + gate_b[h], h < n_gates (to_gates logits, bs_roformer.py:117; optional)
 *   rowinv[r] = inv                                                (optional; normalize == 2: planes stay
 *                                                                   un-normalised and rowinv[r*ss_slots + 0] = ||x_r||^2,
 *                                                                   slots 1.. = 0: the rowss input of sesa_gemm_tc) */
int sesa_prep_rows(const float* x, int64_t ldx, int64_t rows, int dim, int normalize, void* planes, int64_t ldp,
                   int64_t p_plane, int out_planes, const float* gate_w, const float* gate_b, int n_gates,
                   float* gates, int64_t ldg, float* rowinv, int ss_slots, void* stream);
/* BandSplit prologue (bs_roformer.py:241-249): planes[p][r][plane_offs[b] + i] = split(feat[r][offs[b] + i] / max(||feat[r][offs[b]:offs[b+1]]||, 1e-12)),
 * zero padded up to plane_offs[b+1] (multiples of 8 so that every band is a 16-byte aligned TMA operand). */
int sesa_band_prep(const float* feat, int64_t ld_feat, int64_t rows, int n_bands, const int32_t* offs,
                   const int32_t* plane_offs, void* planes, int64_t ldp, int64_t p_plane, int out_planes, void* stream);
/* bf16 hi/lo planes of a weight matrix w[rows][cols] -> planes[2][rows][ldp] (zero padded to ldp). */
int sesa_split_weight(const float* w, int64_t rows, int64_t cols, void* planes, int64_t ldp, void* stream);

/* ---- MDX23C (TFC_TDF_net) support: channels-last activations x[b][t][f][c] ----------------------------------- */
/* InstanceNorm2d statistics (get_norm 'InstanceNorm', mdx23c_tfc_tdf_v3.py:47-59: biased variance, eps 1e-5) per
 * (b, c): stats[b][c] = (mean, 1/sqrt(var+eps)).  layout 0: x[(b*n1+i)*ld + c] (n2 = 1); layout 1:
 * x[((b*n1+i)*C + c)*n2 + j].  scratch: 2*batch*channels doubles. */
int sesa_instnorm_stats(const float* x, int layout, int batch, int64_t n1, int channels, int n2, int64_t ld,
                        double* scratch, float* stats, float eps, void* stream);
/* y = act((x - mean) * rstd * gamma + beta) split into bf16 hi/lo planes (the "norm -> act" prologue of every conv /
 * Linear of TFC_TDF, :104-128).  stats == NULL skips the normalisation; act: SESA_ACT_NONE, SESA_ACT_GELU or 4 (ReLU).
 * mode 0: channels-last in and out (n2 = 1): planes[(b*n1+i)*ldp + c];
 * mode 1: channels-last in (n1 = T, n2 = F, x[((b*T+t)*F + f)*ld + c]) -> channel-major planes[((b*T+t)*C + c)*ldp + f];
 * mode 2: channel-major in and out: x[((b*n1+i)*C + c)*n2 + j] -> planes[((b*n1+i)*C + c)*ldp + j]. */
int sesa_norm_act_split(const float* x, int mode, int batch, int64_t n1, int channels, int n2, int64_t ld,
                        const float* stats, const float* gamma, const float* beta, int act, void* planes, int64_t ldp,
                        int64_t p_plane, void* stream);
/* x[(bt*F + f)*ld + c] += g[(bt*C + c)*F + f]: "x = x + tdf(x)" (:134) with the TDF output channel-major. */
int sesa_transpose_add(float* x, const float* g, int64_t bt, int F, int channels, int64_t ld, void* stream);
/* The same update fused with the InstanceNorm2d statistics of its result (the norm of tfc2, :124,135): stats[b][c] =
 * (mean, 1/sqrt(var + eps)) of the updated x over (frames, F); scratch: 2*batch*channels doubles. */
int sesa_transpose_add_stats(float* x, const float* g, int batch, int64_t frames, int F, int channels, int64_t ld,
                             double* scratch, float* stats, float eps, void* stream);
/* cac2cws (:191-196): spec (sesa_stft layout 0, [bt][f_full][c2]) -> mix[bt][fs][c2*k + kk], f_full = kk*fs + f'. */
int sesa_mdx_pack(const float* spec, int64_t bt, int f_full, int fs, int k, int c2, float* mix, void* stream);
/* planes[r] = split([mix[r] | x[r] * first[r]]): "x * first_conv_out" and cat([mix, x]) (:228-230). */
int sesa_mdx_final_concat(const float* mix, int ch, const float* x, int64_t ldx, const float* first, int64_t ldf,
                          int channels, int64_t rows, void* planes, int64_t ldp, int64_t p_plane, void* stream);
/* cws2cac (:198-203) + zero padding of STFT.inverse (:36-38): y[bt][fs][nt*c2*k] -> out[(b*nt+n)][t][f_full][c2]
 * (the layout sesa_mask_istft mode 2 reads). */
int sesa_mdx_unpack(const float* y, int64_t batch, int64_t frames, int fs, int k, int c2, int nt, int f_full,
                    float* out, void* stream);

/* ---- windowed overlap-add of chunk outputs (utils.py:432-464) ------------------------------- */
/* chunk_out[k][n][c][L]; result[n][c][out_len] = sum_k (ascending) y*w / sum_k w over padded positions
 * p = crop + i, NaN -> 0.  window[L] is _getWindowingArray (utils.py:295-327); kinds[k]: 0 both ramps,
 * 1 no fade-in, 2 no fade-out (utils.py:432-437).  counter (optional, [padded_len]) receives sum_k w. */
int sesa_overlap_add(const float* chunk_out, const int64_t* starts, const int64_t* lens, const int32_t* kinds,
                     int n_chunks, int64_t step, int64_t chunk_size, int fade, const float* window,
                     int nstems, int channels, int64_t padded_len, int64_t crop, int64_t out_len,
                     float* result, float* counter, void* stream);

/* RMSNorm at the end of a Mel-Band Transformer (mel_band_roformer.py:218,226) fused with the operand preparation of the
 * next GEMM: x[rows][dim] <- x / max(||x||, 1e-12) * sqrt(dim) * gamma in place, planes <- bf16 hi/lo of the new rows,
 * ss_out[row][ss_slots] <- (sum of squares of the new row, 0, ...).  Bit-identical to sesa_rmsnorm followed by
 * sesa_prep_rows(normalize = 2). */
int sesa_rmsnorm_planes(float* x, const float* gamma, int64_t rows, int dim, void* planes, int64_t ldp, int64_t p_plane,
                        int out_planes, float* ss_out, int ss_slots, void* stream);

/* ---- streaming overlap-add (the product path of utils.py:439-464) ---------------------------- */
/* Folds the model outputs y[nb][n][c][L] of chunks [k0, k0+nb) into the track result, one call per engine batch, in
 * ascending chunk order per sample (result += x * window).  The padded mix is cut into step-long regions; regions
 * [r_begin, r_end) are processed.  A region whose earlier chunks were folded before (by an earlier call, or by the
 * previous rank of a chunk-range shard whose raw sums were copied in as the halo) continues from
 * partial[n*c][part_ld] (padded positions part_p0..); a region with chunks still to come stores its raw sums there;
 * a region whose last chunk is in this batch is finished: / sum_k w of the GLOBAL schedule, NaN -> 0
 * (utils.py:457-459), cropped by `crop` (:462-464) and written to out[n*c][out_ld] at column (p - crop) - out_q0 when
 * that lies in [0, out_cols) and p - crop in [0, out_len).  partial is never read before it was written, so it needs
 * no initialisation. */
int sesa_overlap_accumulate(const float* y, int k0, int nb, const int64_t* starts, const int64_t* lens,
                            const int32_t* kinds, int n_chunks, int64_t step, int64_t chunk_size, int fade,
                            const float* window, int nstems, int channels, int64_t padded_len, int r_begin, int r_end,
                            float* partial, int64_t part_ld, int64_t part_p0, int64_t crop, int64_t out_len, float* out,
                            int64_t out_ld, int64_t out_q0, int64_t out_cols, void* stream);
/* dst[c][i] = mix[c][reflect(p0 + i - left)], i in [0, count): the border reflect pad of utils.py:391-393 for one
 * slice [p0, p0+count) of the padded mix, reading a window src[c][src_cols] that holds mix[:, src_off : src_off+src_cols)
 * (a chunk-range shard uploads only the samples its chunks touch). */
int sesa_pad_reflect_slice(const float* src, int64_t src_cols, int64_t src_off, float* dst, int channels, int64_t len,
                           int64_t left, int64_t p0, int64_t count, void* stream);

/* ---- test-time augmentation (utils.py:241-292) ----------------------------------------------- */
/* swapped[c] = mix[C-1-c] (mix[::-1].copy(), :271), negated = -1.0 * mix (:271). */
int sesa_tta_variants(const float* mix, float* swapped, float* negated, int channels, int64_t len, void* stream);
/* out[n][c] = ((orig[n][c] + swapped_est[n][C-1-c]) - negated_est[n][c]) / 3 — the += / -= / /= sequence of :283-290. */
int sesa_tta_combine(const float* orig, const float* swapped_est, const float* negated_est, float* out, int nstems,
                     int channels, int64_t len, void* stream);

/* ---- waveform ensembling (ensemble.py:172-183 process_waveform) -------------------------------- */
/* out[i] = reduce over the n_inputs device arrays inputs_host[m][i] (a HOST array of device pointers), accumulated in
 * float64 in input order like numpy on the reference's float64 buffers.  method 0: mean, or sum(x*w)/sum(w) when
 * weights_host (host, float64, already normalised like ensemble.py:293-295) is given; 1: median; 2: max; 3: min. */
int sesa_ensemble_wave(const float* const* inputs_host, int n_inputs, const double* weights_host, int method, float* out,
                       int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SESA_B200_H */
