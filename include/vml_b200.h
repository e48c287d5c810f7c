/*
 * vml_b200.h -- C ABI of libvml_b200.so: the B200 (sm_100a) implementation of the
 * SMIN cross-modal proposal-scoring hot path of ChanukyaVardhan/Video-Moment-Localization.
 *
 * The reference has no FFI / plugin registry (it is pure PyTorch); its boundary for this
 * path is the Python contract   models.SMIN.forward  (models.py:367-377),
 * main.loss_fn (main.py:110-116) and utils.compute_ious (utils.py:10-31).  Each entry
 * point below names the reference function(s) it replaces.  The Python host side
 * (video-moment-localization_b200/smin.py etc.) binds these with ctypes and mirrors the
 * reference interface; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - no allocation inside, no global state except a cached driver entry point;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 = launched, negative = error (vml_last_error() has the text);
 *   - `prec`: VML_FP32 = validation mode (CUDA-core fp32 math, fp32 activations),
 *             VML_BF16 = fast mode (tcgen05 bf16 MMA, fp32 accumulate, bf16 activations);
 *   - "act" pointers are float* in VML_FP32 and __nv_bfloat16* in VML_BF16.
 *
 * Packed moment map.  The reference carries dense (B,L,L,...) tensors whose invalid cells
 * are exactly zero (invariant 3).  This library stores only the valid
 * cells: vml_build_cells() compacts moment_mask into a cell list sorted by (b,i,j);
 * fc is [n_cells, C, D], fm is [n_cells, D].  n_cells lives on the device
 * (cells->n_cells[0]); buffers are sized for `capacity` cells by the caller.
 */
#ifndef VML_B200_H
#define VML_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VML_API __attribute__((visibility("default")))
#else
#define VML_API
#endif

#define VML_FP32 0
#define VML_BF16 1
/* fp32 tensors, dense products on the tensor cores as TF32 (10-bit mantissa inputs, fp32 accumulation): accepted by the
 * GEMM-backed entry points (vml_linear, vml_clip_projection, vml_content_out, vml_moment_out) for the TRAINING forward;
 * everything else treats it as VML_FP32. */
#define VML_TF32 2

#define VML_OK 0
#define VML_ERR_ARG (-1)
#define VML_ERR_CUDA (-2)
#define VML_ERR_UNSUPPORTED (-3)

/* Model dimensions: config/*.yml:5-13 (T, L, C, d, dl, num_smi_layers, input_video_dim,
 * max_query_length, lstm_hidden_size).  D must equal 2*H (models.py:81). */
typedef struct {
  int32_t T, L, C, D, dl, layers, d0, Nq, H;
} vml_dims_t;

/* Compacted list of valid moment-map cells (device memory, caller-allocated). */
typedef struct {
  int32_t* code;      /* [capacity]  (b << 16) | (i << 8) | j, sorted by (b,i,j)            */
  int32_t* row_start; /* [B*L + 1]   first cell of map row (b,i); row_start[B*L] = n_cells  */
  int32_t* n_cells;   /* [1]         number of valid cells                                  */
  int32_t* status;    /* [1]         bit0: capacity overflow                                */
  int32_t capacity;
} vml_cells_t;

VML_API const char* vml_last_error(void);
VML_API int vml_version(void);
/* Number of kernels this library has launched since it was loaded (bench evidence). */
VML_API int64_t vml_launch_count(void);
/* Names of all kernels compiled into the library, '\n'-separated (for smoke/bench reports). */
VML_API const char* vml_kernel_names(void);

/* ---- layout ------------------------------------------------------------------------- */

/* moment_mask[B,L,L] (bool/u8) -> cell list.  Replaces the dense masking of
 * models.py:117,244,290,337 by compaction. */
VML_API int vml_build_cells(const uint8_t* moment_mask, int B, int L, vml_cells_t cells, void* stream);

/* packed [n_cells, inner] <-> dense [B, L, L, inner] (invalid cells zero-filled). */
VML_API int vml_unpack_cells(const void* packed, void* dense, vml_cells_t cells, int B, int L, int inner, int prec,
                     void* stream);
VML_API int vml_pack_cells(const void* dense, void* packed, vml_cells_t cells, int B, int L, int inner, int prec,
                   void* stream);

/* fp32 [rows, k] -> bf16 [rows, k_pad] (zero padded), for TMA-legal operand rows. */
VML_API int vml_cast_pad_bf16(const float* src, void* dst_bf16, int64_t rows, int k, int k_pad, void* stream);

/* One launch that takes the caller's forward() arguments (models.py:367: video_features[B,T,d0],
 * query_features[B,Nq,300] float; video/query/length/moment masks as bytes, dataset.py:165-176)
 * into library-owned operand buffers: features -> bf16 rows zero-padded to v_kpad / q_kpad
 * (VML_BF16) or float copies (VML_FP32, kpad == k); masks -> 0/1 byte copies; sm (optional,
 * the IoU map compute_ious reads) -> copy; qlen[B] = sum(query_mask) (models.py:50, without
 * the D2H copy of :52).  Any output pointer may be NULL (skipped).  Everything after this
 * call reads only library-owned buffers, so the rest of a step can be a replayed CUDA graph. */
VML_API int vml_ingest(const float* video_features, const float* query_features, const uint8_t* video_mask,
                       const uint8_t* query_mask, const uint8_t* length_mask, const uint8_t* moment_mask, const float* sm,
                       void* v_out, void* q_out, uint8_t* vmask_out, uint8_t* qmask_out, uint8_t* lmask_out,
                       uint8_t* mmask_out, float* sm_out, int32_t* qlen, int B, vml_dims_t d, int v_kpad, int q_kpad,
                       int prec, void* stream);
/* Same, for callers that keep the clip features / word vectors as bf16 on the host (half the H2D bytes of a step;
 * in VML_BF16 mode the operands are bit-identical to those vml_ingest makes from the fp32 tensors, because the
 * fp32 -> bf16 rounding is round-to-nearest on either side of the copy). */
VML_API int vml_ingest_bf16(const void* video_features, const void* query_features, const uint8_t* video_mask,
                            const uint8_t* query_mask, const uint8_t* length_mask, const uint8_t* moment_mask,
                            const float* sm, void* v_out, void* q_out, uint8_t* vmask_out, uint8_t* qmask_out,
                            uint8_t* lmask_out, uint8_t* mmask_out, float* sm_out, int32_t* qlen, int B, vml_dims_t d,
                            int v_kpad, int q_kpad, int prec, void* stream);
/* Same, with the clip features PACKED: video_rows holds only the first min(nfeats[b], T) rows of every sample, back to
 * back ([sum_b min(nfeats[b], T), d0], float or -- src_bf16 != 0 -- bf16); the rows dataset.py:69-73
 * (get_fixed_length_features: `out = np.zeros((T, d)); out[:nfeats] = cur_feat`) leaves at zero are re-created by this
 * launch instead of crossing PCIe.  nfeats: device int64 [B] (dataset.py:72).  B <= 4096.  The operands written are
 * bit-identical to vml_ingest's on the padded [B, T, d0] tensor.
 * src_bf16: bit 0 = the feature sources are bf16; bit 1 = the word vectors are packed too: query_features is ignored and the
 * first qlen[b] = sum(query_mask[b]) rows of every sample follow the packed clip rows in the same buffer, at the next
 * 256-byte boundary (video_rows then has to be 256-byte aligned); the other word rows are written as zeros (dataset.py:36
 * pads them with the <pad> vector, which models.py:50-54 never reads: the sequence is packed to its length). */
VML_API int vml_ingest_packed(const void* video_rows, const void* query_features, const uint8_t* video_mask,
                              const uint8_t* query_mask, const uint8_t* length_mask, const uint8_t* moment_mask,
                              const float* sm, const int64_t* nfeats, void* v_out, void* q_out, uint8_t* vmask_out,
                              uint8_t* qmask_out, uint8_t* lmask_out, uint8_t* mmask_out, float* sm_out, int32_t* qlen,
                              int B, vml_dims_t d, int v_kpad, int q_kpad, int prec, int src_bf16, void* stream);

/* ---- dense contractions --------------------------------------------------------------- */

/* out[M,N] = A[M,K] . W[N,K]^T + bias[N].  VML_FP32: A,W,out float.  VML_BF16: A,W bf16
 * (K multiple of 8), out bf16 or float (out_fp32 != 0); tcgen05/TMEM kernel.
 * m_dev, if not NULL, is a device int32; the live row count is min(M, *m_dev * m_scale)
 * (e.g. n_cells * C).  Replaces the nn.Linear / 1x1-conv call sites of
 * models.py:21,134-135,204-205,236-239,285-286. */
VML_API int vml_linear(const void* A, const void* W, const float* bias, void* out, int M, int N, int K, int ldo,
                       const int32_t* m_dev, int m_scale, int prec, int out_fp32, void* stream);

/* Strided, batched fp32 contraction (CUDA cores), the workhorse of the backward path:
 *   C[b][m][n] (=|+=) alpha * sum_k A[b][m][k] * B[b][n][k],  element strides sam/sak/sab etc.
 * so that dX = dY.W, dW = dY^T.X and the per-sample attention products need no transposed copies.
 * accumulate != 0: add into C.  splits > 1: K is cut into ranges combined with fp32 atomics (requires
 * accumulate).  m_dev / k_dev (optional): device int32 live counts, live M = min(M, *m_dev*m_scale), same for K.
 * Replaces the autograd-generated mm/bmm calls behind loss.backward() (main.py:150). */
VML_API int vml_gemm_strided(const float* A, int64_t sam, int64_t sak, int64_t sab, const float* B, int64_t sbn, int64_t sbk,
                             int64_t sbb, float* C, int64_t scm, int64_t scn, int64_t scb, int M, int N, int K, int batch,
                             float alpha, int accumulate, int splits, const int32_t* m_dev, int m_scale,
                             const int32_t* k_dev, int k_scale, void* stream);

/* a1: VideoEncoder.forward (models.py:25-36).  fv = (v.W^T + b)*mask + pe[t]*mask.
 * v: float [B*T,d0] (VML_FP32) or bf16 [B*T,k_pad] (VML_BF16).  fv: act [B*T, D]. */
VML_API int vml_clip_projection(const void* v, const void* W, const float* bias, const float* pe,
                        const uint8_t* video_mask, void* fv, int B, vml_dims_t d, int k_pad, int prec,
                        void* stream);

/* ---- a2: QueryEncoder (models.py:48-64) -------------------------------------------------- */

/* One bi-LSTM layer's recurrence.  gin[B*Nq, 2*4H] float = x.W_ih^T + b_ih + b_hh for
 * (forward | reverse) directions (gate order i,f,g,o); whh_t[2][H][4H] float = W_hh^T per
 * direction; qlen[B] int32 valid words.  y[B,Nq,2H] float (zero for t >= qlen, as
 * pad_packed_sequence); y_bf16 optional copy; fs[B,2H] optional = [h_fwd(len-1) | h_bwd(0)]
 * (+ optional bf16 copy).  One 8-CTA cluster per (direction, 8 samples); W_hh stays in smem. */
VML_API int vml_lstm_layer(const float* gin, const float* whh_t, const int32_t* qlen, float* y, void* y_bf16,
                           float* fs, void* fs_bf16, int B, int Nq, int H, void* stream);

/* Fast-mode variant of the recurrence (H = 256): h.W_hh^T as bf16 mma.sync with the weights resident in
 * registers, fp32 cell state; one 4-CTA cluster per (direction, 16 samples).  whh_frag: the bf16
 * fragment-ordered copy of W_hh, uint32 [2 dirs][4 CTAs][8 warps][32 chunks][32 lanes][4]
 * (register 4*chunk+w of a lane holds the pair W_hh[q*H + 64*cta + 8*warp + lane/4][16*ks + 2*(lane%4) + 8*j + {0,1}],
 * (ks, q, j) = (reg/8, (reg/2)%4, reg%2); see smin.pack_lstm_fragments).  Other arguments as vml_lstm_layer. */
VML_API int vml_lstm_layer_tc(const float* gin, const void* whh_frag, const int32_t* qlen, float* y, void* y_bf16,
                              float* fs, void* fs_bf16, int B, int Nq, int H, void* stream);

/* query_mask[B,Nq] u8 -> qlen[B] int32  (models.py:50, without the D2H copy of :52). */
VML_API int vml_query_lengths(const uint8_t* query_mask, int32_t* qlen, int B, int Nq, void* stream);

/* ---- a3+a4: Backbone fusion + ProposalGeneration (models.py:81,88-98,115-126) ----------- */

/* f = fv*fs fused with span pooling; only valid cells are written.
 * fv act [B,T,D]; fs float [B,D]; fc act [cap,C,D]; fm act [cap,D]; fb float [B,L,D]. */
VML_API int vml_span_pool_fuse(const void* fv, const float* fs, vml_cells_t cells, void* fc, void* fm, float* fb,
                       int B, vml_dims_t d, int prec, void* stream);

/* ---- input pipeline (main.py:118-133: 13 blocking, un-pinned copies per step in the reference) ---- */

/* One asynchronous host-to-device copy of a pinned host blob on `stream` (the whole collated batch travels as
 * ONE blob, see pipeline.pack_host_batch); no allocation, no synchronisation. */
VML_API int vml_copy_h2d_async(void* dst, const void* src_pinned, int64_t bytes, void* stream);

/* ---- labels and masks of a batch from the annotation scalars (dataset.py:95-127,139-158) ------- */

/* times double [B,2] = ground-truth (start, end) seconds, duration double [B], nfeats int64 [B] (clips kept of T).
 * Outputs (any may be NULL): sm float [B,L,L] IoU map, ym u8 = sm > 0.5; ss / se float [B,L] Gaussian boundary
 * penalties, ys / ye u8 = . > 0.5; ya u8 [B,L] snippet inside the moment; length_mask u8 [B,L];
 * moment_mask u8 [B,L,L] (upper triangle of valid snippets); video_mask u8 [B,T].  The u8 outputs hold 0/1 and can
 * be viewed as bool.  sm / ym / ya / masks are bit-exact with the reference's float32 CPU arithmetic; ss / se up
 * to the exp implementation (<= 2 ulp). */
VML_API int vml_make_labels(const double* times, const double* duration, const int64_t* nfeats, int B, int T, int L,
                            float* sm, uint8_t* ym, float* ss, uint8_t* ys, float* se, uint8_t* ye, uint8_t* ya,
                            uint8_t* length_mask, uint8_t* moment_mask, uint8_t* video_mask, void* stream);

/* Fixed-length clip sampling on the device (dataset.py:40-74 get_fixed_length_features): raw [sum nfeats, d0] holds the
 * videos' own clip features back to back, offsets [B+1] their row ranges.  frame_idx = round-half-even(spos + i*stride),
 * stride = 1 (nfeats <= T) or nfeats/T, cut to T; video_features [B,T,d0] rows >= min(nfeats, T) are zero-filled,
 * video_mask [B,T] (optional) marks the live rows, nfeats [B] = min(nfeats, T); start_index / end_index [B] are the
 * sampled-clip indices of the normalised ground-truth positions (dataset.py:57-62).  spos [B] (optional; NULL = 0, the
 * non-training splits) is the caller's random start offset (dataset.py:44-49).  status (device int32, OR-ed): bit 1 = a
 * sample whose index list fits neither nfeats nor T -- where the reference raises its AssertionError. */
VML_API int vml_sample_clips(const float* raw, const int64_t* offsets, const int32_t* spos, const double* start_pos,
                             const double* end_pos, int B, int T, int d0, float* video_features, uint8_t* video_mask,
                             int64_t* nfeats, int32_t* start_index, int32_t* end_index, int32_t* status, void* stream);

/* ---- a5+a6: ContentUnit (models.py:207-226,242-276) ----------------------------------------- */

/* middle of the unit: c_hat act [n*C, dl] -> cc_hat act [n*C, dl]
 * (content-word attention, gate, CxC self-attention).  Query-side tensors come from the folded
 * query projection qproj float [B*Nq, ld]: w_hat (models.py:249) at column off_what, the
 * W_q-folded keys ktil at off_ktil and their bias term beta at off_beta, so that
 * Q.K^T (models.py:209-211) = c_hat.ktil^T + beta; s_hat float [B, s_ld] (models.py:251). */
VML_API int vml_content_attention(const void* c_hat, const float* qproj, int ld, int off_what, int off_ktil,
                                  int off_beta, const float* s_hat, int s_ld, const uint8_t* query_mask,
                                  vml_cells_t cells, void* cc_hat, int B, vml_dims_t d, int prec, void* stream);

/* Fused front half of the unit (VML_BF16, dl = 128, C = 4, Nq <= 24): c_hat = fc.W^T + bias on
 * tcgen05, then attention / gate / clip self-attention in the same kernel; c_hat never leaves
 * the SM.  fc bf16 [cap*4, D], W bf16 [128, D], cc_hat bf16 [cap*4, 128]. */
VML_API int vml_content_in_attention(const void* fc, const void* W, const float* bias, const float* qproj, int ld,
                                     int off_what, int off_ktil, int off_beta, const float* s_hat, int s_ld,
                                     const uint8_t* query_mask, vml_cells_t cells, void* cc_hat, int B,
                                     vml_dims_t d, void* stream);

/* The WHOLE unit in one kernel (VML_BF16, dl = 128, C = 4, D % 128 == 0, D <= 512, Nq <= 31; query with
 * vml_content_unit_supported): a tile's 128 (cell, clip) rows of fc stay resident in shared memory from the
 * c_hat contraction to the residual add, so a layer reads fc once and writes cu once:
 *   cu = ContentUnit(fc) = cc_hat.Wc^T + bc + fc + fbar   (models.py:242-276; fbar = sigmoid(fm*fs)*fm per cell,
 *   see vml_boundary_unit),  mu_operand[n, D:2D] = mean_c cu (models.py:297).
 * store_cu = 0 skips the cu store (last SMI layer: only mean_c cu is consumed downstream; cu may alias fc).
 * fc, cu bf16 [cap*4, D]; W_chat bf16 [128, D]; Wc bf16 [D, 128]; fbar bf16 [cap, D]; mu_operand bf16 [cap, 2D]. */
VML_API int vml_content_unit(const void* fc, const void* W_chat, const float* b_chat, const float* qproj, int ld,
                             int off_what, int off_ktil, int off_beta, const float* s_hat, int s_ld,
                             const uint8_t* query_mask, vml_cells_t cells, const void* Wc, const float* bc,
                             const void* fbar, void* cu, void* mu_operand, int B, vml_dims_t d, int store_cu,
                             void* stream);
VML_API int vml_content_unit_supported(vml_dims_t d);

/* cu = cc_hat.Wc^T + bc + fc + sigmoid(fm*fs)*fm   (models.py:269-276).  VML_BF16 with fbar and
 * mu_operand non-NULL selects the fused epilogue: the gate term is read from fbar (see
 * vml_boundary_unit) and mean_c cu (models.py:297) is written to mu_operand[n, D:2D]. */
VML_API int vml_content_out(const void* cc_hat, const void* Wc, const float* bc, const void* fc, const void* fm,
                            const float* fs, const void* fbar, void* mu_operand, vml_cells_t cells, void* cu,
                            vml_dims_t d, int prec, void* stream);

/* ---- a7: BoundaryUnit (models.py:137-154,164-196) ------------------------------------------- */

/* Boundary-word scores use the W_q-folded keys kbt (column off_kbt of qproj) and their bias
 * term beta_b (column off_betab): (fb.Wq^T+bq).(fw.Wk^T+bk)^T = fb.kbt^T + beta_b.
 * Scratch: g_scratch float [B,L,D] (the gated rows G), ab_scratch float [B,L,L] (the attention rows A_b).
 * bu float [B,L,D] = f_bb + f_b + f_bm.
 * fbar (optional, act [n, D]) receives sigmoid(fm*fs)*fm per cell (models.py:191 == :272-274),
 * which the fused content-out epilogue reuses instead of recomputing it per clip.  fbar_bias (optional, float [D]) is
 * added to the STORED fbar before its rounding (not to f_bm): the content unit's output bias b_c then travels with the gate
 * term, and vml_content_unit called with bc == NULL adds fbar only (its residual add runs on the tensor cores).
 * Three launches: gate and rows (warp-level TF32 mma; 3xTF32 split in VML_FP32), then a streaming pass over
 * the map cells (one CTA per map row). */
VML_API int vml_boundary_unit(const float* qproj, int ld, int off_kbt, int off_betab, const float* fw, const float* fs,
                      const float* fb, const void* fm, const uint8_t* query_mask,
                      const uint8_t* length_mask, vml_cells_t cells, float* g_scratch, float* ab_scratch, float* bu,
                      void* fbar, const float* fbar_bias, float* prob_out, float* u_out, int B, vml_dims_t d, int prec,
                      void* stream);
/* prob_out (optional, float [B,L,Nq]) and u_out (optional, float [B,L,D] = Aq*lmask + fs) are the saved
 * activations the backward pass needs (training path only). */

/* a7 + the first half of a8's operand in one call: as vml_boundary_unit, and the per-sample streaming kernel also writes
 * operand[n, 0:D] = bu[b,i] * bu[b,j] (models.py:292-295; what vml_moment_pair computes) while the sample's boundary rows are
 * still on the SM.  operand: act [n, 2D].  Only where vml_boundary_pair_fused(d, prec) returns 1 (fast mode, L <= 16,
 * D = 256 or 512); bit-identical to vml_boundary_unit followed by vml_moment_pair. */
VML_API int vml_boundary_pair_fused(vml_dims_t d, int prec);
VML_API int vml_boundary_unit_pair(const float* qproj, int ld, int off_kbt, int off_betab, const float* fw, const float* fs,
                                   const float* fb, const void* fm, const uint8_t* query_mask, const uint8_t* length_mask,
                                   vml_cells_t cells, float* g_scratch, float* ab_scratch, float* bu, void* fbar,
                                   const float* fbar_bias, void* operand, int B, vml_dims_t d, int prec, void* stream);

/* ---- a8: MomentUnit (models.py:288-303) ------------------------------------------------------ */

/* operand[n, 2D] = [ bu[b,i]*bu[b,j] | mean_c cu[n,c,:] ]  (act) */
VML_API int vml_moment_operand(const void* cu, const float* bu, vml_cells_t cells, void* operand, vml_dims_t d,
                       int prec, void* stream);
/* operand[n, 0:D] = bu[b,i]*bu[b,j] only (the other half comes from the fused content-out epilogue) */
VML_API int vml_moment_pair(const float* bu, vml_cells_t cells, void* operand, vml_dims_t d, int prec, void* stream);
/* mu = operand.[Wfb|Wfc]^T + (bfb+bfc) + fm */
VML_API int vml_moment_out(const void* operand, const void* Wcat, const float* bias_sum, const void* fm,
                   vml_cells_t cells, void* mu, vml_dims_t d, int prec, void* stream);
/* Same product with the bu[b,i] * bu[b,j] half of the operand generated INSIDE the GEMM from the boundary rows bu float [B, L, D]
 * (four extra warps write those k-blocks of the A operand straight into the pipeline stages, rounded to the activation type as
 * vml_moment_pair rounds them): operand[n, 0:D] is neither written nor read, only operand[n, D:2D] = mean_c cu is loaded.
 * Only where vml_moment_gen_supported(d, prec) returns 1 (VML_BF16, D a multiple of 64); bit-identical to vml_moment_pair +
 * vml_moment_out. */
VML_API int vml_moment_gen_supported(vml_dims_t d, int prec);
VML_API int vml_moment_out_gen(const void* operand, const void* Wcat, const float* bias_sum, const void* fm, vml_cells_t cells,
                               const float* bu, void* mu, int B, vml_dims_t d, int prec, void* stream);

/* ---- a9: Localization (models.py:335-344) ------------------------------------------------------ */

/* w4 float [4,D] = (pm,ps,pe,pa) weights, b4 float[4].  Outputs dense float, masked to 0. */
VML_API int vml_localize(const void* fm, const float* fb, const float* w4, const float* b4, vml_cells_t cells,
                 const uint8_t* length_mask, float* pm, float* ps, float* pe, float* pa, int B,
                 vml_dims_t d, int prec, void* stream);

/* ---- a10: loss (main.py:89-116, with reduction=None read as 'none') -------------------------- */

/* loss[0] = L_m + L_s + L_e + 0.5 L_a ; parts[4] = the four terms; scratch float [4*B].
 * Optional gradients d loss / d{pm,ps,pe,pa} (all NULL to skip). */
VML_API int vml_scaled_iou_bce(const float* pm, const uint8_t* ym, const float* sm, const uint8_t* moment_mask,
                       const float* ps, const uint8_t* ys, const float* ss, const float* pe,
                       const uint8_t* ye, const float* se, const float* pa, const uint8_t* ya,
                       const uint8_t* length_mask, int B, int L, float* loss, float* parts, float* scratch,
                       float* g_pm, float* g_ps, float* g_pe, float* g_pa, void* stream);

/* ---- a11: compute_ious (utils.py:10-31) --------------------------------------------------------- */

/* score = ((pm*sqrt(ps_i))*sqrt(pe_j))*mask ; top-k (k <= 8) by score, ties -> lowest flat
 * index ; top_iou = sm[idx] ; counts[n_idx*4 + m_idx] += any(top_iou[:n] > m) for
 * n in {1,5}, m in {.1,.3,.5,.7} (ACCUMULATES into counts, like the reference's running
 * sums, main.py:155-156).  nms_num/nms_den: temporal-NMS IoU threshold as a rational;
 * nms_num >= nms_den disables NMS (the reference has none, utils.py:14).  step_counts (optional)
 * is a second accumulator of the same hits for per-step read-back: sample b adds into
 * step_counts[(b / step_group) * 8 + ...] (step_group <= 0: one group), so a pass over several
 * coalesced batches still reports each batch's own hits. */
VML_API int vml_score_topk_recall(const float* pm, const float* ps, const float* pe, const uint8_t* moment_mask,
                          const float* sm, int B, int L, int k, int nms_num, int nms_den, int32_t* top_idx,
                          float* top_score, float* top_iou, int64_t* counts, int64_t* step_counts, int step_group,
                          void* stream);
/* The same with the caller's own lists, as utils.py:10 allows (`n`, `m` are arbitrary Python lists there): ns[n_n]
 * (n_n <= 8, each >= 1) and ms[n_m] (n_m <= 8) are HOST arrays; k = max(ns) <= 32 is the reference's topk(max(n));
 * counts / step_counts are laid out [n_n][n_m] (step group stride n_n*n_m). */
VML_API int vml_score_topk_recall_nm(const float* pm, const float* ps, const float* pe, const uint8_t* moment_mask,
                          const float* sm, int B, int L, int k, int nms_num, int nms_den, int32_t* top_idx,
                          float* top_score, float* top_iou, int64_t* counts, int64_t* step_counts, int step_group,
                          const int32_t* ns, int n_n, const float* ms, int n_m, void* stream);

/* ---- backward / training path (fp32): adjoints behind loss.backward() (main.py:150) -----------------------
 * All tensors float.  Outputs documented "+=" accumulate (the caller zeroes gradient buffers once per step);
 * reductions over cells use fp32 atomics.  Dense products of the backward pass are vml_gemm_strided calls. */

/* out[b][n] += alpha * sum_m X[b][m][n]  (bias gradients, per-sample sums); rows may be device-counted. */
VML_API int vml_colsum(const float* X, int64_t row_stride, int64_t batch_stride, float* out, int64_t out_batch_stride, int M, int N,
                       int batch, const int32_t* m_dev, int m_scale, float alpha, void* stream);
/* a9 backward (models.py:335-344): d_fm [n,D], d_fb [B,L,D] written; dw4 [4,D], db4 [4] += . */
VML_API int vml_localize_bwd(const float* fm, const float* fb, const float* w4, const float* pm, const float* ps, const float* pe,
                             const float* pa, const float* g_pm, const float* g_ps, const float* g_pe, const float* g_pa,
                             const uint8_t* length_mask, vml_cells_t cells, float* d_fm, float* d_fb, float* dw4, float* db4, int B,
                             vml_dims_t d, void* stream);
/* a8 backward of bu_i*bu_j (models.py:292-295): d_bu [B,L,D] += gather over the row and the column of every snippet. */
VML_API int vml_pair_bwd(const float* d_operand, int ld_operand, const float* bu, vml_cells_t cells, float* d_bu, int B, vml_dims_t d,
                         void* stream);
/* a6 tail backward, elementwise: dY = d_cu_next (may be NULL) + d_operand[:, D:]/C;  d_gbar = sum_c dY. */
VML_API int vml_cu_tail_bwd(const float* d_cu_next, const float* d_operand, int ld_operand, vml_cells_t cells, float* dY, float* d_gbar,
                            vml_dims_t d, void* stream);
/* a5+a6 attention block backward (models.py:207-226,253-266): d_chat [n*C, dl] written; dq (gradient of the folded
 * query projection, same layout as qproj) and d_shat [B, s_ld] += . */
VML_API int vml_content_attn_bwd(const float* c_hat, const float* d_cc, const float* qproj, int ld, int off_what, int off_ktil,
                                 int off_beta, const float* s_hat, int s_ld, const uint8_t* query_mask, vml_cells_t cells,
                                 float* d_chat, float* dq, float* d_shat, int B, vml_dims_t d, void* stream);
/* gate term gbar = sigmoid(fm*fs)*fm (models.py:191-194 and 272-274): d_fm [n,D] = d_mu + ..., d_ab [B,L,L] +=, d_fs [B,D] += . */
VML_API int vml_gbar_bwd(const float* fm, const float* fs, const float* ab, const float* d_bu, const float* d_gbar_cu, const float* d_mu,
                         vml_cells_t cells, float* d_ab, float* d_fm, float* d_fs, int B, vml_dims_t d, void* stream);
/* row softmax backward: dS = P*(dP - sum(P*dP))*scale*colmask[b,w]; P, dP, dS [batch*R, W]. */
VML_API int vml_softmax_bwd(const float* P, const float* dP, const uint8_t* colmask, float* dS, int batch, int R, int W, float scale,
                            void* stream);
/* a7 gate backward: G = fb*U: d_fb += dG*U, d_Aq = dG*fb*lmask, tmp = dG*fb (its per-sample column sum is d_fs). */
VML_API int vml_gate_bwd(const float* dG, const float* fb, const float* U, const uint8_t* length_mask, float* d_fb, float* d_Aq,
                         float* tmp, int B, vml_dims_t d, void* stream);
/* Y[r,:] (=|+=) X[r,:] * mask[r] */
VML_API int vml_mask_rows(const float* X, const uint8_t* mask, float* Y, int64_t rows, int D, int accumulate, void* stream);
/* a3+a4 backward (models.py:81,88-98,115-126): d_fv [B,T,D] written, d_fs [B,D] += . */
VML_API int vml_span_pool_bwd(const float* d_fc, const float* d_fm, const float* d_fb, const float* fv, const float* fs,
                              vml_cells_t cells, float* d_fv, float* d_fs, int B, vml_dims_t d, void* stream);
/* Fused Adam step over a flat buffer (main.py:83: torch.optim.Adam defaults); step counts from 1; grad_scale multiplies g. */
VML_API int vml_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                          int step, float grad_scale, void* stream);
/* a2 training path: packed bi-LSTM layer forward that also saves acts [B,Nq,2,5,H] = (i,f,g,o,c), and its BPTT:
 * dgin [B*Nq,8H] and dgin_rec (zero at each sequence's first processed step) written. whh_t [2][H][4H], whh [2][4H][H]. */
VML_API int vml_lstm_train_fwd(const float* gin, const float* whh_t, const int32_t* qlen, float* y, float* fs, float* acts, int B,
                               int Nq, int H, void* stream);
VML_API int vml_lstm_train_bwd(const float* dy, const float* dfs, const float* whh, const float* acts, const int32_t* qlen, float* dgin,
                               float* dgin_rec, int B, int Nq, int H, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VML_B200_H */
