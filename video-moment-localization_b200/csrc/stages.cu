// Non-GEMM stages of the SMIN hot path: cell compaction, fused span pooling + fusion,
// content-word attention, boundary unit, moment-unit operand build, localization.
// Each kernel is templated on the activation storage type (float = validation mode,
// bf16 = fast mode); all arithmetic is fp32.
#include "common.cuh"

namespace vml {

// =====================================================================================
// cell compaction:  moment_mask[B,L,L] -> sorted list of valid cells
// =====================================================================================
__global__ void cells_count_kernel(const uint8_t* __restrict__ mask, int rows, int L, int32_t* __restrict__ cnt) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (warp >= rows) return;
  const uint8_t* m = mask + (size_t)warp * L;
  int c = 0;
  for (int j = lane; j < L; j += 32) c += m[j] ? 1 : 0;
  c = (int)warp_sum((float)c);  // L <= 256: exact in fp32
  if (lane == 0) cnt[warp] = c;
}

// single block: exclusive scan of cnt[rows] -> row_start[rows+1]; n_cells = min(total, capacity)
__global__ void cells_scan_kernel(int32_t* __restrict__ row_start, int rows, int32_t* __restrict__ n_cells,
                                  int32_t* __restrict__ status, int capacity) {
  __shared__ int part[1024];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int per = (rows + nt - 1) / nt;
  const int lo = min(tid * per, rows), hi = min(lo + per, rows);
  int s = 0;
  for (int r = lo; r < hi; ++r) s += row_start[r];
  part[tid] = s;
  __syncthreads();
  for (int off = 1; off < nt; off <<= 1) {  // Hillis-Steele inclusive scan
    int v = tid >= off ? part[tid - off] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  int run = part[tid] - s;
  for (int r = lo; r < hi; ++r) { int c = row_start[r]; row_start[r] = run; run += c; }
  if (tid == nt - 1) {
    const int total = part[tid];
    row_start[rows] = total;
    n_cells[0] = min(total, capacity);
    status[0] = total > capacity ? 1 : 0;
  }
}

__global__ void cells_fill_kernel(const uint8_t* __restrict__ mask, int rows, int L,
                                  const int32_t* __restrict__ row_start, int32_t* __restrict__ code, int capacity) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (warp >= rows) return;
  const uint8_t* m = mask + (size_t)warp * L;
  const int b = warp / L, i = warp % L;
  int base = row_start[warp];
  for (int j0 = 0; j0 < L; j0 += 32) {
    const int j = j0 + lane;
    const bool on = j < L && m[j];
    const unsigned bal = __ballot_sync(0xffffffffu, on);
    if (on) {
      const int pos = base + __popc(bal & ((1u << lane) - 1u));
      if (pos < capacity) code[pos] = (b << 16) | (i << 8) | j;
    }
    base += __popc(bal);
  }
}

int build_cells(const uint8_t* mask, int B, int L, vml_cells_t cells, cudaStream_t st) {
  VML_CHECK_ARG(B > 0 && B < 32768 && L > 0 && L <= 256 && cells.capacity > 0);
  static bool reg = (register_kernel("cells_count_kernel"), register_kernel("cells_scan_kernel"),
                     register_kernel("cells_fill_kernel"), true);
  (void)reg;
  const int rows = B * L;
  const int blocks = ceil_div(rows * 32, 256);
  cells_count_kernel<<<blocks, 256, 0, st>>>(mask, rows, L, cells.row_start);
  cells_scan_kernel<<<1, 1024, 0, st>>>(cells.row_start, rows, cells.n_cells, cells.status, cells.capacity);
  cells_fill_kernel<<<blocks, 256, 0, st>>>(mask, rows, L, cells.row_start, cells.code, cells.capacity);
  VML_LAUNCHED(3);
  return VML_OK;
}

// packed <-> dense helpers ---------------------------------------------------------------
template <typename T>
__global__ void unpack_cells_kernel(const T* __restrict__ packed, T* __restrict__ dense, const int32_t* __restrict__ code,
                                    const int32_t* __restrict__ n_cells, int L, int inner) {
  const int n = *n_cells;
  for (int c = blockIdx.x; c < n; c += gridDim.x) {
    int b, i, j; decode_cell(code[c], b, i, j);
    const T* s = packed + (size_t)c * inner;
    T* d = dense + (((size_t)b * L + i) * L + j) * inner;
    for (int e = threadIdx.x; e < inner; e += blockDim.x) d[e] = s[e];
  }
}
template <typename T>
__global__ void pack_cells_kernel(const T* __restrict__ dense, T* __restrict__ packed, const int32_t* __restrict__ code,
                                  const int32_t* __restrict__ n_cells, int L, int inner) {
  const int n = *n_cells;
  for (int c = blockIdx.x; c < n; c += gridDim.x) {
    int b, i, j; decode_cell(code[c], b, i, j);
    T* d = packed + (size_t)c * inner;
    const T* s = dense + (((size_t)b * L + i) * L + j) * inner;
    for (int e = threadIdx.x; e < inner; e += blockDim.x) d[e] = s[e];
  }
}

int unpack_cells(const void* packed, void* dense, vml_cells_t cells, int B, int L, int inner, int prec, cudaStream_t st) {
  static bool reg = (register_kernel("unpack_cells_kernel"), true); (void)reg;
  const size_t esz = prec == VML_BF16 ? 2 : 4;
  VML_CUDA(cudaMemsetAsync(dense, 0, (size_t)B * L * L * inner * esz, st));
  const int grid = min(cells.capacity, kNumSMs * 16);
  if (prec == VML_BF16) unpack_cells_kernel<bf16><<<grid, 128, 0, st>>>((const bf16*)packed, (bf16*)dense, cells.code, cells.n_cells, L, inner);
  else unpack_cells_kernel<float><<<grid, 128, 0, st>>>((const float*)packed, (float*)dense, cells.code, cells.n_cells, L, inner);
  VML_LAUNCHED(1);
  return VML_OK;
}
int pack_cells(const void* dense, void* packed, vml_cells_t cells, int B, int L, int inner, int prec, cudaStream_t st) {
  static bool reg = (register_kernel("pack_cells_kernel"), true); (void)reg;
  const int grid = min(cells.capacity, kNumSMs * 16);
  if (prec == VML_BF16) pack_cells_kernel<bf16><<<grid, 128, 0, st>>>((const bf16*)dense, (bf16*)packed, cells.code, cells.n_cells, L, inner);
  else pack_cells_kernel<float><<<grid, 128, 0, st>>>((const float*)dense, (float*)packed, cells.code, cells.n_cells, L, inner);
  VML_LAUNCHED(1);
  return VML_OK;
}

// fp32 -> bf16 with zero padding of the row to k_pad -------------------------------------
__global__ void cast_pad_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t rows, int k, int k_pad) {
  const int64_t total = rows * (k_pad / 4);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / (k_pad / 4);
    const int c = (int)(e % (k_pad / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c + 3 < k) v = *reinterpret_cast<const float4*>(src + r * k + c);
    else {
      if (c < k) v.x = src[r * k + c];
      if (c + 1 < k) v.y = src[r * k + c + 1];
      if (c + 2 < k) v.z = src[r * k + c + 2];
    }
    st4(dst + r * k_pad + c, v);
  }
}
int cast_pad(const float* src, void* dst, int64_t rows, int k, int k_pad, cudaStream_t st) {
  VML_CHECK_ARG(k % 4 == 0 && k_pad % 4 == 0 && k_pad >= k);
  static bool reg = (register_kernel("cast_pad_kernel"), true); (void)reg;
  if (rows <= 0) return VML_OK;
  const int64_t total = rows * (k_pad / 4);
  const int64_t want = ceil_div64(total, 256), cap = (int64_t)kNumSMs * 16;
  const int grid = (int)(want < cap ? want : cap);
  cast_pad_kernel<<<grid, 256, 0, st>>>(src, (bf16*)dst, rows, k, k_pad);
  VML_LAUNCHED(1);
  return VML_OK;
}

// =====================================================================================
// ingest: ONE launch that takes the caller's tensors into the library's operand buffers --
// clip features / word vectors fp32 -> bf16 rows zero-padded to a TMA-legal stride (or fp32
// copies), the four masks -> u8 copies, sm copy, qlen = sum(query_mask) (models.py:50).
// Everything downstream reads only library-owned buffers, so the rest of the step can be a
// replayed CUDA graph.
// =====================================================================================
struct IngestArgs {
  const float* src[2];   // video_features [rows0, k0], query_features [rows1, k1]
  void* dst[2];          // bf16 [rows, kpad] or float [rows, k]; nullptr = skip
  int64_t rows[2];
  int k[2], kpad[2];
  const uint8_t* msrc[4];  // video, query, length, moment masks
  uint8_t* mdst[4];
  int64_t mbytes[4];
  const float* sm_src; float* sm_dst; int64_t sm_n;
  const uint8_t* qmask; int32_t* qlen; int B, Nq;
};

template <bool BF16>
__global__ void __launch_bounds__(256)
ingest_kernel(IngestArgs a) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (!a.dst[s]) continue;
    const int k = a.k[s], kp4 = a.kpad[s] / 4;
    const int64_t total = a.rows[s] * kp4;
    const float* src = a.src[s];
    for (int64_t e = tid; e < total; e += nth) {
      const int64_t r = e / kp4;
      const int c = (int)(e - r * kp4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c + 3 < k) v = __ldg(reinterpret_cast<const float4*>(src + r * k + c));
      else {
        if (c < k) v.x = src[r * k + c];
        if (c + 1 < k) v.y = src[r * k + c + 1];
        if (c + 2 < k) v.z = src[r * k + c + 2];
      }
      if (BF16) st4(reinterpret_cast<bf16*>(a.dst[s]) + r * a.kpad[s] + c, v);
      else st4(reinterpret_cast<float*>(a.dst[s]) + r * a.kpad[s] + c, v);
    }
  }
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    if (!a.mdst[s]) continue;
    for (int64_t e = tid; e < a.mbytes[s]; e += nth) a.mdst[s][e] = a.msrc[s][e] ? 1 : 0;
  }
  if (a.sm_dst)
    for (int64_t e = tid; e < a.sm_n; e += nth) a.sm_dst[e] = a.sm_src[e];
  if (a.qlen)
    for (int64_t b = tid; b < a.B; b += nth) {
      int n = 0;
      for (int w = 0; w < a.Nq; ++w) n += a.qmask[b * a.Nq + w] ? 1 : 0;
      a.qlen[b] = n;
    }
}

int ingest(const float* vf, const float* qf, const uint8_t* vmask, const uint8_t* qmask, const uint8_t* lmask,
           const uint8_t* mmask, const float* sm, void* v_out, void* q_out, uint8_t* vmask_out, uint8_t* qmask_out,
           uint8_t* lmask_out, uint8_t* mmask_out, float* sm_out, int32_t* qlen, int B, vml_dims_t d, int v_kpad,
           int q_kpad, int prec, cudaStream_t st) {
  VML_CHECK_ARG(B > 0 && d.d0 % 4 == 0 && v_kpad % 4 == 0 && q_kpad % 4 == 0 && v_kpad >= d.d0 && q_kpad >= 300);
  VML_CHECK_ARG(prec == VML_BF16 || (v_kpad == d.d0 && q_kpad == 300));
  static bool reg = (register_kernel("ingest_kernel"), true); (void)reg;
  IngestArgs a;
  a.src[0] = vf; a.src[1] = qf; a.dst[0] = v_out; a.dst[1] = q_out;
  a.rows[0] = (int64_t)B * d.T; a.rows[1] = (int64_t)B * d.Nq;
  a.k[0] = d.d0; a.k[1] = 300; a.kpad[0] = v_kpad; a.kpad[1] = q_kpad;
  a.msrc[0] = vmask; a.msrc[1] = qmask; a.msrc[2] = lmask; a.msrc[3] = mmask;
  a.mdst[0] = vmask_out; a.mdst[1] = qmask_out; a.mdst[2] = lmask_out; a.mdst[3] = mmask_out;
  a.mbytes[0] = (int64_t)B * d.T; a.mbytes[1] = (int64_t)B * d.Nq; a.mbytes[2] = (int64_t)B * d.L;
  a.mbytes[3] = (int64_t)B * d.L * d.L;
  a.sm_src = sm; a.sm_dst = sm ? sm_out : nullptr; a.sm_n = (int64_t)B * d.L * d.L;
  a.qmask = qmask; a.qlen = qlen; a.B = B; a.Nq = d.Nq;
  const int64_t total = (v_out ? a.rows[0] * (v_kpad / 4) : 0) + (q_out ? a.rows[1] * (q_kpad / 4) : 0) + a.mbytes[3];
  const int64_t want = ceil_div64(total, 256 * 4), cap = (int64_t)kNumSMs * 8;
  const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
  if (prec == VML_BF16) ingest_kernel<true><<<grid, 256, 0, st>>>(a);
  else ingest_kernel<false><<<grid, 256, 0, st>>>(a);
  VML_LAUNCHED(1);
  return VML_OK;
}

// =====================================================================================
// a3+a4  fused  f = fv*fs  +  span pooling over clips   (models.py:81,88-98,115-126)
//
// One CTA per (sample, D-slice).  The slice of the fused clip sequence is turned into an
// inclusive prefix sum over t in shared memory (fp32); every pooled clip of every valid cell
// is then a difference of two prefix rows times fp32(1/clip_size) -- exactly the non-zero
// pattern of the reference's dense Wc -- and only valid (b,i,j) cells are stored, with
// 128-bit stores.  fm = mean_c fc (always / C), fb = average pool over T/L clips.
// =====================================================================================
template <typename ActT>
__global__ void __launch_bounds__(256)
span_pool_kernel(const ActT* __restrict__ fv, const float* __restrict__ fs, const int32_t* __restrict__ code,
                 const int32_t* __restrict__ row_start, ActT* __restrict__ fc, ActT* __restrict__ fm,
                 float* __restrict__ fb, int T, int L, int C, int D, int dslice, int capacity) {
  extern __shared__ __align__(16) float P[];  // [(T+1)][dslice]
  const int b = blockIdx.x, d0 = blockIdx.y * dslice;
  const int tid = threadIdx.x;
  const int r = T / L;

  // prefix sums of fv*fs along t, one column per thread (coalesced across the slice)
  for (int d = tid; d < dslice; d += blockDim.x) {
    const float s = fs[(size_t)b * D + d0 + d];
    const ActT* col = fv + (size_t)b * T * D + d0 + d;
    float run = 0.f;
    P[d] = 0.f;
    int t = 0;
    for (; t + 8 <= T; t += 8) {
      float x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) x[u] = to_f(col[(size_t)(t + u) * D]);
#pragma unroll
      for (int u = 0; u < 8; ++u) { run += x[u] * s; P[(size_t)(t + u + 1) * dslice + d] = run; }
    }
    for (; t < T; ++t) { run += to_f(col[(size_t)t * D]) * s; P[(size_t)(t + 1) * dslice + d] = run; }
  }
  __syncthreads();

  const int groups = dslice / 8;             // 8 columns per thread
  const int g = tid % groups, lane_cell = tid / groups, cells_per_iter = blockDim.x / groups;
  const int dd = g * 8;

  // fb: unmasked average pool (models.py:120-125)
  const float inv_r = 1.0f / (float)r;
  for (int l = lane_cell; l < L; l += cells_per_iter) {
    const float* hi = P + (size_t)((l + 1) * r) * dslice + dd;
    const float* lo = P + (size_t)(l * r) * dslice + dd;
    f8 a = ld8(hi), c = ld8(lo), o;
#pragma unroll
    for (int e = 0; e < 8; ++e) o.v[e] = (a.v[e] - c.v[e]) * inv_r;
    st8(fb + ((size_t)b * L + l) * D + d0 + dd, o);
  }

  const int n_lo = row_start[b * L], n_hi = min(row_start[(b + 1) * L], capacity);
  const float inv_C = 1.0f / (float)C;
  for (int n = n_lo + lane_cell; n < n_hi; n += cells_per_iter) {
    const int cd = code[n];
    const int i = (cd >> 8) & 0xff, j = cd & 0xff;
    const int nf = (j - i + 1) * r;
    const int cs = max(1, nf / C);
    const int nclips = min(C, nf);           // <= 0 below the diagonal -> all-zero cell
    const float w = 1.0f / (float)cs;        // fp32(1/clip_size), as the reference's Wc
    f8 mean;
#pragma unroll
    for (int e = 0; e < 8; ++e) mean.v[e] = 0.f;
    const int s0 = i * r;
    ActT* fc_row = fc + ((size_t)n * C) * D + d0 + dd;
    for (int c = 0; c < C; ++c) {
      f8 o;
      if (c < nclips) {
        f8 a = ld8(P + (size_t)(s0 + (c + 1) * cs) * dslice + dd);
        f8 z = ld8(P + (size_t)(s0 + c * cs) * dslice + dd);
#pragma unroll
        for (int e = 0; e < 8; ++e) { o.v[e] = (a.v[e] - z.v[e]) * w; mean.v[e] += o.v[e]; }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) o.v[e] = 0.f;
      }
      st8(fc_row + (size_t)c * D, o);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) mean.v[e] *= inv_C;
    st8(fm + (size_t)n * D + d0 + dd, mean);
  }
}

int span_pool_fuse(const void* fv, const float* fs, vml_cells_t cells, void* fc, void* fm, float* fb, int B,
                   vml_dims_t d, int prec, cudaStream_t st) {
  VML_CHECK_ARG(d.T % d.L == 0 && d.D % 8 == 0 && d.C >= 1);
  static bool reg = (register_kernel("span_pool_kernel"), true); (void)reg;
  // largest D-slice (multiple of 8 dividing D) whose prefix table fits in 200 KB and that
  // still yields >= 2 CTAs per SM
  int dslice = d.D;
  auto fits = [&](int s) { return (size_t)(d.T + 1) * s * 4 <= 200 * 1024; };
  while (dslice > 8 && (dslice % 16 == 0) && (!fits(dslice) || (int64_t)B * (d.D / dslice) < 2 * kNumSMs)) dslice /= 2;
  VML_CHECK_ARG(fits(dslice) && d.D % dslice == 0 && dslice % 8 == 0 && dslice / 8 <= 256);
  const size_t smem = (size_t)(d.T + 1) * dslice * 4;
  dim3 grid(B, d.D / dslice);
  if (prec == VML_BF16) {
    VML_CUDA(cudaFuncSetAttribute(span_pool_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    span_pool_kernel<bf16><<<grid, 256, smem, st>>>((const bf16*)fv, fs, cells.code, cells.row_start, (bf16*)fc,
                                                    (bf16*)fm, fb, d.T, d.L, d.C, d.D, dslice, cells.capacity);
  } else {
    VML_CUDA(cudaFuncSetAttribute(span_pool_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    span_pool_kernel<float><<<grid, 256, smem, st>>>((const float*)fv, fs, cells.code, cells.row_start, (float*)fc,
                                                     (float*)fm, fb, d.T, d.L, d.C, d.D, dslice, cells.capacity);
  }
  VML_LAUNCHED(1);
  return VML_OK;
}

// =====================================================================================
// a5+a6 (middle)  content-word attention, gate, CxC self-attention
//   (ContentAttention.forward models.py:207-226; ContentUnit.forward models.py:253-266)
// One warp per cell (C = 4 clips x dl features); lane k owns query word k in the score phase
// and dl/32 feature columns elsewhere.  Q.K^T is evaluated as c_hat.ktil^T + beta, where the
// per-word ktil / beta / w_hat (and the per-sentence s_hat) are columns of the single folded
// query projection `qproj` (weights folded at pack time, see smin.py:pack_weights).
// =====================================================================================
template <typename ActT, int DPL>
__global__ void __launch_bounds__(256)
content_attention_kernel(const ActT* __restrict__ c_hat, const float* __restrict__ qproj, int ld, int off_what,
                         int off_ktil, int off_beta, const float* __restrict__ s_hat, int s_ld,
                         const uint8_t* __restrict__ qmask, const int32_t* __restrict__ row_start,
                         ActT* __restrict__ cc_hat, int L, int Nq, int capacity) {
  constexpr int C = 4, DL = DPL * 32, KS = DL + 1;
  extern __shared__ __align__(16) float sm[];
  float* s_k = sm;                      // [Nq][DL+1]   ktil (padded: lane k reads row k)
  float* s_w = s_k + Nq * KS;           // [Nq][DL]     w_hat
  float* s_s = s_w + Nq * DL;           // [DL]         s_hat
  float* s_b = s_s + DL;                // [32]         beta
  float* s_m = s_b + 32;                // [32]         qmask as float
  float* s_c = s_m + 32;                // [warps][C][DL] c_hat staging
  float* s_p = s_c + 8 * C * DL;        // [warps][C][32] attention probabilities
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid / 32, lane = tid % 32;

  for (int e = tid; e < Nq * DL; e += blockDim.x) {
    const int k = e / DL, dcol = e % DL;
    const float* qr = qproj + ((size_t)b * Nq + k) * ld;
    s_k[k * KS + dcol] = qr[off_ktil + dcol];
    s_w[e] = qr[off_what + dcol];
  }
  for (int e = tid; e < DL; e += blockDim.x) s_s[e] = s_hat[(size_t)b * s_ld + e];
  if (tid < 32) {
    s_b[tid] = tid < Nq ? qproj[((size_t)b * Nq + tid) * ld + off_beta] : 0.f;
    s_m[tid] = tid < Nq ? (qmask[(size_t)b * Nq + tid] ? 1.f : 0.f) : 0.f;
  }
  __syncthreads();

  const int n_lo = row_start[b * L], n_hi = min(row_start[(b + 1) * L], capacity);
  float* my_c = s_c + warp * C * DL;
  float* my_p = s_p + warp * C * 32;
  const float sqrt_dl = sqrtf((float)DL);
  const int dcol = lane * DPL;

  for (int n = n_lo + blockIdx.x * 8 + warp; n < n_hi; n += gridDim.x * 8) {
    float ch[C][DPL];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const ActT* src = c_hat + ((size_t)n * C + c) * DL + dcol;
#pragma unroll
      for (int e = 0; e < DPL; ++e) { ch[c][e] = to_f(src[e]); my_c[c * DL + dcol + e] = ch[c][e]; }
    }
    __syncwarp();
    // scores: lane k <-> word k
    float sc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) sc[c] = 0.f;
    if (lane < Nq) {
      const float* kr = s_k + lane * KS;
      for (int dd = 0; dd < DL; ++dd) {
        const float kv = kr[dd];
#pragma unroll
        for (int c = 0; c < C; ++c) sc[c] = fmaf(my_c[c * DL + dd], kv, sc[c]);
      }
    }
    const float mk = s_m[lane], bt = s_b[lane];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float s = (sc[c] + bt) / sqrt_dl;
      s = s * mk;
      if (mk == 0.f) s = -1e9f;                      // masked_fill(mask == 0, -1e9)
      if (lane >= Nq) s = -INFINITY;                 // not a word at all
      const float mx = warp_max(s);
      const float ex = lane < Nq ? expf(s - mx) : 0.f;
      const float den = warp_sum(ex);
      my_p[c * 32 + lane] = ex / den;
    }
    __syncwarp();
    // attended words + gate
    float g[C][DPL];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int e = 0; e < DPL; ++e) g[c][e] = 0.f;
    for (int k = 0; k < Nq; ++k) {
      float wv[DPL];
#pragma unroll
      for (int e = 0; e < DPL; ++e) wv[e] = s_w[k * DL + dcol + e];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float p = my_p[c * 32 + k];
#pragma unroll
        for (int e = 0; e < DPL; ++e) g[c][e] = fmaf(p, wv[e], g[c][e]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int e = 0; e < DPL; ++e) g[c][e] = ch[c][e] * (g[c][e] + s_s[dcol + e]);
    // CxC self-attention over the clips of this cell
    float a[C][C];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int c2 = c; c2 < C; ++c2) {
        float p = 0.f;
#pragma unroll
        for (int e = 0; e < DPL; ++e) p = fmaf(g[c][e], g[c2][e], p);
        p = warp_sum(p) / sqrt_dl;
        a[c][c2] = p; a[c2][c] = p;
      }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float mx = a[c][0];
#pragma unroll
      for (int c2 = 1; c2 < C; ++c2) mx = fmaxf(mx, a[c][c2]);
      float den = 0.f;
#pragma unroll
      for (int c2 = 0; c2 < C; ++c2) { a[c][c2] = expf(a[c][c2] - mx); den += a[c][c2]; }
#pragma unroll
      for (int c2 = 0; c2 < C; ++c2) a[c][c2] /= den;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      ActT* dst = cc_hat + ((size_t)n * C + c) * DL + dcol;
#pragma unroll
      for (int e = 0; e < DPL; ++e) {
        float o = 0.f;
#pragma unroll
        for (int c2 = 0; c2 < C; ++c2) o = fmaf(a[c][c2], ch[c2][e], o);
        dst[e] = from_f<ActT>(o);
      }
    }
    __syncwarp();
  }
}

template <typename ActT, int DPL>
static int launch_content_attention(const void* c_hat, const float* qproj, int ld, int off_what, int off_ktil,
                                    int off_beta, const float* s_hat, int s_ld, const uint8_t* qmask,
                                    vml_cells_t cells, void* cc_hat, int B, vml_dims_t d, cudaStream_t st) {
  constexpr int DL = DPL * 32;
  const size_t smem = sizeof(float) * ((size_t)d.Nq * (DL + 1) + (size_t)d.Nq * DL + DL + 64 + 8 * 4 * DL + 8 * 4 * 32);
  VML_CUDA(cudaFuncSetAttribute(content_attention_kernel<ActT, DPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int vmax = d.L * (d.L + 1) / 2;
  int chunks = ceil_div(vmax, 8 * 4);             // ~4 cells per warp
  while ((int64_t)chunks * B > (int64_t)kNumSMs * 16 && chunks > 1) chunks = (chunks + 1) / 2;
  dim3 grid(chunks, B);
  content_attention_kernel<ActT, DPL><<<grid, 256, smem, st>>>((const ActT*)c_hat, qproj, ld, off_what, off_ktil, off_beta,
                                                               s_hat, s_ld, qmask, cells.row_start, (ActT*)cc_hat, d.L,
                                                               d.Nq, cells.capacity);
  VML_LAUNCHED(1);
  return VML_OK;
}

int content_attention(const void* c_hat, const float* qproj, int ld, int off_what, int off_ktil, int off_beta,
                      const float* s_hat, int s_ld, const uint8_t* qmask, vml_cells_t cells, void* cc_hat, int B,
                      vml_dims_t d, int prec, cudaStream_t st) {
  VML_CHECK_ARG(d.C == 4 && d.Nq <= 32 && (d.dl == 32 || d.dl == 64 || d.dl == 128));
  static bool reg = (register_kernel("content_attention_kernel"), true); (void)reg;
#define VML_CA(T, P) \
  return launch_content_attention<T, P>(c_hat, qproj, ld, off_what, off_ktil, off_beta, s_hat, s_ld, qmask, cells, cc_hat, B, d, st)
  if (prec == VML_BF16) { if (d.dl == 128) VML_CA(bf16, 4); if (d.dl == 64) VML_CA(bf16, 2); VML_CA(bf16, 1); }
  if (d.dl == 128) VML_CA(float, 4); if (d.dl == 64) VML_CA(float, 2); VML_CA(float, 1);
#undef VML_CA
}

// =====================================================================================
// a7  BoundaryUnit  (Attention.forward models.py:137-154; BoundaryUnit.forward :164-196)
// =====================================================================================
// Both kernels: one CTA per (sample, tile of 8 map rows), one WARP per row; a lane owns the 16-byte
// column groups {128*i + 4*lane}, so every global/shared access of a row is a coalesced 512 B.
constexpr int BU_RT = 8;        // rows (= warps) per CTA
constexpr int BU_MAXG = 8;      // D <= 128 * BU_MAXG (kernels are instantiated for NG = 1, 2, 4, 8 column groups)

// gate:  G[b,l,:] = fb * (softmax(q.k^T/sqrt(D)) . fw * lmask + fs), with q.k^T = fb.kbt^T + beta_b
// (W_q folded into the per-word keys kbt at pack time, so no per-layer projection of fb).  The
// sample's keys and word states are staged once per CTA in shared memory.
template <int NG>
__global__ void __launch_bounds__(BU_RT * 32)
boundary_gate_kernel(const float* __restrict__ qproj, int ld, int off_kbt, int off_betab,
                     const float* __restrict__ fw, const float* __restrict__ fs, const float* __restrict__ fb,
                     const uint8_t* __restrict__ qmask, const uint8_t* __restrict__ lmask, float* __restrict__ G,
                     int L, int Nq, int D) {
  extern __shared__ __align__(16) float sg[];
  float* s_k = sg;                 // [Nq][D]  kbt
  float* s_w = s_k + Nq * D;       // [Nq][D]  fw
  float* s_bm = s_w + Nq * D;      // [32] beta_b, [32] mask
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int dq = D / 4;
  for (int e = tid; e < Nq * dq; e += blockDim.x) {
    const int k = e / dq, c4 = (e % dq) * 4;
    *reinterpret_cast<float4*>(s_k + k * D + c4) = __ldg(reinterpret_cast<const float4*>(qproj + ((size_t)b * Nq + k) * ld + off_kbt + c4));
    *reinterpret_cast<float4*>(s_w + k * D + c4) = __ldg(reinterpret_cast<const float4*>(fw + ((size_t)b * Nq + k) * D + c4));
  }
  if (tid < 32) {
    s_bm[tid] = tid < Nq ? qproj[((size_t)b * Nq + tid) * ld + off_betab] : 0.f;
    s_bm[32 + tid] = (tid < Nq && qmask[(size_t)b * Nq + tid]) ? 1.f : 0.f;
  }
  __syncthreads();
  const int l = blockIdx.x * BU_RT + warp;
  if (l >= L) return;
  const int row = b * L + l;
  float4 x[NG];
#pragma unroll
  for (int i = 0; i < NG; ++i) x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < NG; ++i)
    if (i * 128 + lane * 4 < D) x[i] = __ldg(reinterpret_cast<const float4*>(fb + (size_t)row * D + i * 128 + lane * 4));
  // scores over the words (lane k keeps word k's score)
  float my_s = -INFINITY;
  for (int k = 0; k < Nq; ++k) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < NG; ++i)
      if (i * 128 + lane * 4 < D) {
        const float4 w = *reinterpret_cast<const float4*>(s_k + k * D + i * 128 + lane * 4);
        acc = fmaf(x[i].x, w.x, acc); acc = fmaf(x[i].y, w.y, acc); acc = fmaf(x[i].z, w.z, acc); acc = fmaf(x[i].w, w.w, acc);
      }
    acc = warp_sum(acc);
    if (lane == k) my_s = acc;
  }
  const float mk = s_bm[32 + lane];
  float sv = lane < Nq ? (my_s + s_bm[lane]) / sqrtf((float)D) : 0.f;
  sv = sv * mk;
  if (mk == 0.f) sv = -1e9f;                       // masked_fill(mask == 0, -1e9)
  if (lane >= Nq) sv = -INFINITY;
  const float mx = warp_max(sv);
  const float ex = lane < Nq ? expf(sv - mx) : 0.f;
  const float p_mine = ex / warp_sum(ex);
  // attended words, row mask, gate
  float4 a[NG];
#pragma unroll
  for (int i = 0; i < NG; ++i) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < Nq; ++k) {
    const float p = __shfl_sync(0xffffffffu, p_mine, k);
#pragma unroll
    for (int i = 0; i < NG; ++i)
      if (i * 128 + lane * 4 < D) {
        const float4 w = *reinterpret_cast<const float4*>(s_w + k * D + i * 128 + lane * 4);
        a[i].x = fmaf(p, w.x, a[i].x); a[i].y = fmaf(p, w.y, a[i].y); a[i].z = fmaf(p, w.z, a[i].z); a[i].w = fmaf(p, w.w, a[i].w);
      }
  }
  const float lm = lmask[row] ? 1.f : 0.f;
#pragma unroll
  for (int i = 0; i < NG; ++i)
    if (i * 128 + lane * 4 < D) {
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(fs + (size_t)b * D + i * 128 + lane * 4));
      float4 g;
      g.x = x[i].x * (a[i].x * lm + s4.x); g.y = x[i].y * (a[i].y * lm + s4.y);
      g.z = x[i].z * (a[i].z * lm + s4.z); g.w = x[i].w * (a[i].w * lm + s4.w);
      *reinterpret_cast<float4*>(G + (size_t)row * D + i * 128 + lane * 4) = g;
    }
}

// row:  A_b[i,:] = softmax_j(G_i.G_j/sqrt(D)) (masked) ;
//       bu[i] = A_b[i,:].fb + fb[i] + sum_j A_b[i,j] sigmoid(fm_ij*fs)*fm_ij   (also written out as fbar)
template <typename ActT, int NG>
__global__ void __launch_bounds__(BU_RT * 32)
boundary_row_kernel(const float* __restrict__ G, const float* __restrict__ fb, const float* __restrict__ fs,
                    const ActT* __restrict__ fm, const uint8_t* __restrict__ lmask, const int32_t* __restrict__ code,
                    const int32_t* __restrict__ row_start, float* __restrict__ bu, ActT* __restrict__ fbar, int L, int D,
                    int capacity) {
  extern __shared__ float s_all[];  // [BU_RT][L] attention rows
  const int b = blockIdx.y, warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int i_row = blockIdx.x * BU_RT + warp;
  if (i_row >= L) return;
  float* s_a = s_all + warp * L;
  const int row = b * L + i_row;
  const bool row_on = lmask[row] != 0;
  const int n_lo = row_start[row], n_hi = min(row_start[row + 1], capacity);
  float4 s4[NG], acc[NG];
#pragma unroll
  for (int i = 0; i < NG; ++i) { s4[i] = make_float4(0.f, 0.f, 0.f, 0.f); acc[i] = s4[i]; }
#pragma unroll
  for (int i = 0; i < NG; ++i)
    if (i * 128 + lane * 4 < D) {
      s4[i] = __ldg(reinterpret_cast<const float4*>(fs + (size_t)b * D + i * 128 + lane * 4));
      acc[i] = __ldg(reinterpret_cast<const float4*>(fb + (size_t)row * D + i * 128 + lane * 4));   // + f_b
    }
  if (row_on) {
    float4 gi[NG];
#pragma unroll
    for (int i = 0; i < NG; ++i) gi[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NG; ++i)
      if (i * 128 + lane * 4 < D) gi[i] = __ldg(reinterpret_cast<const float4*>(G + (size_t)row * D + i * 128 + lane * 4));
    for (int j0 = 0; j0 < L; j0 += 4) {            // 4 key rows per round trip: all loads first, then the math
      float4 w[4][NG];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* gj = G + ((size_t)b * L + min(j0 + u, L - 1)) * D;
#pragma unroll
        for (int i = 0; i < NG; ++i)
          w[u][i] = (i * 128 + lane * 4 < D) ? __ldg(reinterpret_cast<const float4*>(gj + i * 128 + lane * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float mk4 = (lane < 4 && j0 + lane < L && lmask[b * L + j0 + lane]) ? 1.f : 0.f;
      float d[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < NG; ++i) {
          t = fmaf(gi[i].x, w[u][i].x, t); t = fmaf(gi[i].y, w[u][i].y, t); t = fmaf(gi[i].z, w[u][i].z, t); t = fmaf(gi[i].w, w[u][i].w, t);
        }
        d[u] = warp_sum(t);
      }
      if (lane < 4 && j0 + lane < L) {
        float sc = (lane == 0 ? d[0] : lane == 1 ? d[1] : lane == 2 ? d[2] : d[3]) / sqrtf((float)D);
        sc = sc * mk4;
        if (mk4 == 0.f) sc = -1e9f;
        s_a[j0 + lane] = sc;
      }
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < L; j += 32) mx = fmaxf(mx, s_a[j]);
    mx = warp_max(mx);
    float den = 0.f;
    for (int j = lane; j < L; j += 32) { const float ex = expf(s_a[j] - mx); s_a[j] = ex; den += ex; }
    den = warp_sum(den);
    for (int j = lane; j < L; j += 32) s_a[j] = s_a[j] / den;
    __syncwarp();
    // f_bb = A_b[i,:] . fb  (added to the f_b already in acc: (f_bb + f_b) as the reference orders it)
    float4 bb[NG];
#pragma unroll
    for (int i = 0; i < NG; ++i) bb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int j = 0; j < L; ++j) {
      const float a = s_a[j];
      const float* fj = fb + ((size_t)b * L + j) * D;
#pragma unroll
      for (int i = 0; i < NG; ++i)
        if (i * 128 + lane * 4 < D) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(fj + i * 128 + lane * 4));
          bb[i].x = fmaf(a, w.x, bb[i].x); bb[i].y = fmaf(a, w.y, bb[i].y); bb[i].z = fmaf(a, w.z, bb[i].z); bb[i].w = fmaf(a, w.w, bb[i].w);
        }
    }
#pragma unroll
    for (int i = 0; i < NG; ++i)
      if (i * 128 + lane * 4 < D) { acc[i].x += bb[i].x; acc[i].y += bb[i].y; acc[i].z += bb[i].z; acc[i].w += bb[i].w; }
  }
  // f_bm over the valid cells of this map row; the gated map value is also what the content unit adds
  float4 bm[NG];
#pragma unroll
  for (int i = 0; i < NG; ++i) bm[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int n0 = n_lo; n0 < n_hi; n0 += 4) {        // 4 cells per round trip
    float4 m[4][NG];
    float a4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int n = min(n0 + u, n_hi - 1);
      a4[u] = (row_on && n0 + u < n_hi) ? s_a[code[n] & 0xff] : 0.f;
#pragma unroll
      for (int i = 0; i < NG; ++i)
        m[u][i] = (i * 128 + lane * 4 < D) ? ld4(fm + (size_t)n * D + i * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (n0 + u < n_hi) {
#pragma unroll
        for (int i = 0; i < NG; ++i)
          if (i * 128 + lane * 4 < D) {
            const float4 mm = m[u][i];
            float4 g;
            g.x = sigmoidf_(mm.x * s4[i].x) * mm.x; g.y = sigmoidf_(mm.y * s4[i].y) * mm.y;
            g.z = sigmoidf_(mm.z * s4[i].z) * mm.z; g.w = sigmoidf_(mm.w * s4[i].w) * mm.w;
            if (fbar) st4(fbar + (size_t)(n0 + u) * D + i * 128 + lane * 4, g);
            bm[i].x = fmaf(a4[u], g.x, bm[i].x); bm[i].y = fmaf(a4[u], g.y, bm[i].y);
            bm[i].z = fmaf(a4[u], g.z, bm[i].z); bm[i].w = fmaf(a4[u], g.w, bm[i].w);
          }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NG; ++i)
    if (i * 128 + lane * 4 < D) {
      float4 o;
      o.x = acc[i].x + bm[i].x; o.y = acc[i].y + bm[i].y; o.z = acc[i].z + bm[i].z; o.w = acc[i].w + bm[i].w;
      *reinterpret_cast<float4*>(bu + (size_t)row * D + i * 128 + lane * 4) = o;
    }
}

template <int NG>
static int launch_boundary(const float* qproj, int ld, int off_kbt, int off_betab, const float* fw, const float* fs,
                           const float* fb, const void* fm, const uint8_t* qmask, const uint8_t* lmask, vml_cells_t cells,
                           float* g_scratch, float* bu, void* fbar, int B, vml_dims_t d, int prec, cudaStream_t st) {
  dim3 grid(ceil_div(d.L, BU_RT), B);
  const size_t smem_g = sizeof(float) * (2 * (size_t)d.Nq * d.D + 64);
  VML_CUDA(cudaFuncSetAttribute(boundary_gate_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
  boundary_gate_kernel<NG><<<grid, BU_RT * 32, smem_g, st>>>(qproj, ld, off_kbt, off_betab, fw, fs, fb, qmask, lmask, g_scratch,
                                                             d.L, d.Nq, d.D);
  const size_t smem = sizeof(float) * BU_RT * d.L;
  if (prec == VML_BF16)
    boundary_row_kernel<bf16, NG><<<grid, BU_RT * 32, smem, st>>>(g_scratch, fb, fs, (const bf16*)fm, lmask, cells.code,
                                                                  cells.row_start, bu, (bf16*)fbar, d.L, d.D, cells.capacity);
  else
    boundary_row_kernel<float, NG><<<grid, BU_RT * 32, smem, st>>>(g_scratch, fb, fs, (const float*)fm, lmask, cells.code,
                                                                   cells.row_start, bu, (float*)fbar, d.L, d.D, cells.capacity);
  VML_LAUNCHED(2);
  return VML_OK;
}

int boundary_unit(const float* qproj, int ld, int off_kbt, int off_betab, const float* fw, const float* fs,
                  const float* fb, const void* fm, const uint8_t* qmask, const uint8_t* lmask, vml_cells_t cells,
                  float* g_scratch, float* bu, void* fbar, int B, vml_dims_t d, int prec, cudaStream_t st) {
  VML_CHECK_ARG(d.Nq <= 32 && d.D % 4 == 0 && d.D <= 128 * BU_MAXG && ld % 4 == 0 && off_kbt % 4 == 0);
  static bool reg = (register_kernel("boundary_gate_kernel"), register_kernel("boundary_row_kernel"), true); (void)reg;
#define VML_BU(NG) return launch_boundary<NG>(qproj, ld, off_kbt, off_betab, fw, fs, fb, fm, qmask, lmask, cells, g_scratch, bu, fbar, B, d, prec, st)
  const int ng = ceil_div(d.D, 128);
  if (ng <= 1) VML_BU(1);
  if (ng <= 2) VML_BU(2);
  if (ng <= 4) VML_BU(4);
  VML_BU(8);
#undef VML_BU
}

// =====================================================================================
// a8 operand:  [ bu_i * bu_j | mean_c cu ]   (MomentUnit.forward models.py:292-301)
// =====================================================================================
template <typename ActT>
__global__ void __launch_bounds__(128)
moment_operand_kernel(const ActT* __restrict__ cu, const float* __restrict__ bu, const int32_t* __restrict__ code,
                      const int32_t* __restrict__ n_cells, ActT* __restrict__ op, int L, int C, int D) {
  const int n_total = *n_cells;
  const int per = D / 8;  // 8-column groups per half
  for (int n = blockIdx.x; n < n_total; n += gridDim.x) {
    int b, i, j; decode_cell(code[n], b, i, j);
    for (int t = threadIdx.x; t < 2 * per; t += blockDim.x) {
      f8 o;
      if (t < per) {
        const int dd = t * 8;
        f8 x = ld8(bu + ((size_t)b * L + i) * D + dd), y = ld8(bu + ((size_t)b * L + j) * D + dd);
#pragma unroll
        for (int e = 0; e < 8; ++e) o.v[e] = x.v[e] * y.v[e];
        st8(op + (size_t)n * 2 * D + dd, o);
      } else {
        const int dd = (t - per) * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) o.v[e] = 0.f;
        for (int c = 0; c < C; ++c) {
          f8 x = ld8(cu + ((size_t)n * C + c) * D + dd);
#pragma unroll
          for (int e = 0; e < 8; ++e) o.v[e] += x.v[e];
        }
        const float inv = 1.0f / (float)C;
#pragma unroll
        for (int e = 0; e < 8; ++e) o.v[e] *= inv;
        st8(op + (size_t)n * 2 * D + D + dd, o);
      }
    }
  }
}

// first half only: operand[n, 0:D] = bu_i * bu_j (the mean_c cu half is written by the fused content-out epilogue)
template <typename ActT>
__global__ void __launch_bounds__(256)
moment_pair_kernel(const float* __restrict__ bu, const int32_t* __restrict__ code, const int32_t* __restrict__ n_cells,
                   ActT* __restrict__ op, int L, int D) {
  const int n_total = *n_cells;
  const int per = D / 8, cells_per_blk = blockDim.x / per;
  const int t = threadIdx.x % per, sub = threadIdx.x / per;
  for (int n = blockIdx.x * cells_per_blk + sub; n < n_total; n += gridDim.x * cells_per_blk) {
    int b, i, j; decode_cell(code[n], b, i, j);
    const int dd = t * 8;
    f8 x = ld8(bu + ((size_t)b * L + i) * D + dd), y = ld8(bu + ((size_t)b * L + j) * D + dd), o;
#pragma unroll
    for (int e = 0; e < 8; ++e) o.v[e] = x.v[e] * y.v[e];
    st8(op + (size_t)n * 2 * D + dd, o);
  }
}

int moment_pair(const float* bu, vml_cells_t cells, void* op, vml_dims_t d, int prec, cudaStream_t st) {
  VML_CHECK_ARG(d.D % 8 == 0 && d.D / 8 <= 256 && 256 % (d.D / 8) == 0);
  static bool reg = (register_kernel("moment_pair_kernel"), true); (void)reg;
  const int cpb = 256 / (d.D / 8);
  const int grid = min(ceil_div(cells.capacity, cpb), kNumSMs * 8);
  if (prec == VML_BF16) moment_pair_kernel<bf16><<<grid, 256, 0, st>>>(bu, cells.code, cells.n_cells, (bf16*)op, d.L, d.D);
  else moment_pair_kernel<float><<<grid, 256, 0, st>>>(bu, cells.code, cells.n_cells, (float*)op, d.L, d.D);
  VML_LAUNCHED(1);
  return VML_OK;
}

int moment_operand(const void* cu, const float* bu, vml_cells_t cells, void* op, vml_dims_t d, int prec, cudaStream_t st) {
  VML_CHECK_ARG(d.D % 8 == 0);
  static bool reg = (register_kernel("moment_operand_kernel"), true); (void)reg;
  const int grid = min(cells.capacity, kNumSMs * 16);
  if (prec == VML_BF16)
    moment_operand_kernel<bf16><<<grid, 128, 0, st>>>((const bf16*)cu, bu, cells.code, cells.n_cells, (bf16*)op, d.L, d.C, d.D);
  else
    moment_operand_kernel<float><<<grid, 128, 0, st>>>((const float*)cu, bu, cells.code, cells.n_cells, (float*)op, d.L, d.C, d.D);
  VML_LAUNCHED(1);
  return VML_OK;
}

// =====================================================================================
// a9  Localization (models.py:335-344): sigmoid(1x1 conv) heads, masked
// =====================================================================================
template <typename ActT>
__global__ void __launch_bounds__(256)
localize_pm_kernel(const ActT* __restrict__ fm, const float* __restrict__ w, const float* __restrict__ bias,
                   const int32_t* __restrict__ code, const int32_t* __restrict__ n_cells, float* __restrict__ pm, int L, int D) {
  const int n_total = *n_cells;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = blockDim.x / 32;
  for (int n = blockIdx.x * nw + warp; n < n_total; n += gridDim.x * nw) {
    float acc = 0.f;
    for (int e = lane * 4; e < D; e += 128) {
      float4 x = ld4(fm + (size_t)n * D + e), ww = ld4(w + e);
      acc = fmaf(x.x, ww.x, acc); acc = fmaf(x.y, ww.y, acc); acc = fmaf(x.z, ww.z, acc); acc = fmaf(x.w, ww.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      int b, i, j; decode_cell(code[n], b, i, j);
      pm[((size_t)b * L + i) * L + j] = sigmoidf_(acc + bias[0]);
    }
  }
}

__global__ void __launch_bounds__(128)
localize_boundary_kernel(const float* __restrict__ fb, const float* __restrict__ w4, const float* __restrict__ b4,
                         const uint8_t* __restrict__ lmask, float* __restrict__ ps, float* __restrict__ pe,
                         float* __restrict__ pa, int rows, int D) {
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int row = blockIdx.x * 4 + warp;
  if (row >= rows) return;
  float a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int e = lane; e < D; e += 32) {
    const float x = fb[(size_t)row * D + e];
    a1 = fmaf(x, w4[D + e], a1); a2 = fmaf(x, w4[2 * D + e], a2); a3 = fmaf(x, w4[3 * D + e], a3);
  }
  a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
  if (lane == 0) {
    const float m = lmask[row] ? 1.f : 0.f;
    ps[row] = sigmoidf_(a1 + b4[1]) * m;
    pe[row] = sigmoidf_(a2 + b4[2]) * m;
    pa[row] = sigmoidf_(a3 + b4[3]) * m;
  }
}

int localize(const void* fm, const float* fb, const float* w4, const float* b4, vml_cells_t cells, const uint8_t* lmask,
             float* pm, float* ps, float* pe, float* pa, int B, vml_dims_t d, int prec, cudaStream_t st) {
  VML_CHECK_ARG(d.D % 4 == 0);
  static bool reg = (register_kernel("localize_pm_kernel"), register_kernel("localize_boundary_kernel"), true); (void)reg;
  VML_CUDA(cudaMemsetAsync(pm, 0, sizeof(float) * (size_t)B * d.L * d.L, st));
  const int grid = min(ceil_div(cells.capacity, 8), kNumSMs * 8);
  if (prec == VML_BF16)
    localize_pm_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)fm, w4, b4, cells.code, cells.n_cells, pm, d.L, d.D);
  else
    localize_pm_kernel<float><<<grid, 256, 0, st>>>((const float*)fm, w4, b4, cells.code, cells.n_cells, pm, d.L, d.D);
  localize_boundary_kernel<<<ceil_div(B * d.L, 4), 128, 0, st>>>(fb, w4, b4, lmask, ps, pe, pa, B * d.L, d.D);
  VML_LAUNCHED(2);
  return VML_OK;
}

}  // namespace vml
