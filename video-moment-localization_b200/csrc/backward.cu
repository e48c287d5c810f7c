// Backward of the SMIN hot path (fp32 training path): the hand-written adjoints of the stages in
// stages.cu / boundary_mma.cu / query_loss_eval.cu.  Every dense product of the backward pass is a
// vml_gemm_strided call (gemm_strided.cuh); this file holds the non-GEMM adjoints:
//   localization heads, moment-unit pair product, content-unit tail / attention block, boundary-unit
//   streaming part and softmaxes, span pooling, clip-projection epilogue, column sums.
// Replaces what autograd generates behind loss.backward() (main.py:150) for models.py:25-344.
// Conventions: fp32 everywhere; "+=" outputs accumulate (the caller zeroes gradient buffers once per
// step); reductions over cells use fp32 atomics (sum order may differ in the last ulp run to run).
#include "common.cuh"

namespace vml {

static inline int grid_for(int64_t n) {                 // 256-thread CTAs for a grid-stride loop over n elements
  const int64_t want = ceil_div64(n, 256), cap = (int64_t)kNumSMs * 8;
  return (int)(want < cap ? want : cap);
}

// ------------------------------------------------------------------------------------------------
// out[b][n] += sum_m X[b][m][n]        (bias gradients, per-sample sums); rows may be device-counted
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ X, int64_t sxm, int64_t sxb, float* __restrict__ out, int64_t sob, int M, int N,
              const int32_t* __restrict__ m_dev, int m_scale, float alpha) {
  if (m_dev) M = min(M, *m_dev * m_scale);
  const int b = blockIdx.z;
  const int n = blockIdx.x * 64 + (threadIdx.x & 63), sub = threadIdx.x >> 6;   // 4 row-lanes per column
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int lo = blockIdx.y * rows_per, hi = min(M, lo + rows_per);
  __shared__ float red[4][64];
  float acc = 0.f;
  if (n < N)
    for (int m = lo + sub; m < hi; m += 4) acc += X[(int64_t)b * sxb + (int64_t)m * sxm + n];
  red[sub][threadIdx.x & 63] = acc;
  __syncthreads();
  if (sub == 0 && n < N && lo < hi) {
    const float s = (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
    atomicAdd(out + (int64_t)b * sob + n, alpha * s);
  }
}
int colsum(const float* X, int64_t sxm, int64_t sxb, float* out, int64_t sob, int M, int N, int batch, const int32_t* m_dev,
           int m_scale, float alpha, cudaStream_t st) {
  static bool reg = (register_kernel("colsum_kernel"), true); (void)reg;
  if (M <= 0 || N <= 0) return VML_OK;
  const int chunks = max(1, min(64, M / 64));
  dim3 grid(ceil_div(N, 64), chunks, batch);
  colsum_kernel<<<grid, 256, 0, st>>>(X, sxm, sxb, out, sob, M, N, m_dev, m_scale, alpha);
  VML_LAUNCHED(1);
  return VML_OK;
}

// ------------------------------------------------------------------------------------------------
// a9 backward (models.py:335-344):  p = sigmoid(w.x + b) * mask
//   cells:  d_fm[n,:] = g*w_pm,  dW_pm += g*fm[n,:],  db_pm += g,   g = g_pm * p (1 - p)
//   rows :  d_fb[r,:] = sum_k g_k w_k,  dW_k += g_k fb[r,:],  db_k += g_k   (k = ps, pe, pa; masked rows: 0)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
localize_bwd_cells_kernel(const float* __restrict__ fm, const float* __restrict__ w4, const float* __restrict__ pm,
                          const float* __restrict__ g_pm, const int32_t* __restrict__ code, const int32_t* __restrict__ n_cells,
                          float* __restrict__ d_fm, float* __restrict__ dw4, float* __restrict__ db4, int L, int D) {
  extern __shared__ float accw[];                       // [D] per-CTA dW_pm
  const int n_total = *n_cells;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = blockDim.x / 32;
  for (int e = threadIdx.x; e < D; e += blockDim.x) accw[e] = 0.f;
  __syncthreads();
  float gb = 0.f;
  for (int n = blockIdx.x * nw + warp; n < n_total; n += gridDim.x * nw) {
    int b, i, j; decode_cell(code[n], b, i, j);
    const size_t o = ((size_t)b * L + i) * L + j;
    const float p = pm[o], g = g_pm[o] * p * (1.f - p);
    gb += g;
    for (int e = lane; e < D; e += 32) {
      d_fm[(size_t)n * D + e] = g * w4[e];
      atomicAdd(&accw[e], g * fm[(size_t)n * D + e]);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < D; e += blockDim.x) if (accw[e] != 0.f) atomicAdd(dw4 + e, accw[e]);
  if (lane == 0 && gb != 0.f) atomicAdd(db4, gb);
}
__global__ void __launch_bounds__(128)
localize_bwd_rows_kernel(const float* __restrict__ fb, const float* __restrict__ w4, const float* __restrict__ ps,
                         const float* __restrict__ pe, const float* __restrict__ pa, const float* __restrict__ g_ps,
                         const float* __restrict__ g_pe, const float* __restrict__ g_pa, const uint8_t* __restrict__ lmask,
                         float* __restrict__ d_fb, float* __restrict__ dw4, float* __restrict__ db4, int rows, int D) {
  const int row = blockIdx.x;
  if (row >= rows) return;
  float g1 = 0.f, g2 = 0.f, g3 = 0.f;
  if (lmask[row]) {
    g1 = g_ps[row] * ps[row] * (1.f - ps[row]);
    g2 = g_pe[row] * pe[row] * (1.f - pe[row]);
    g3 = g_pa[row] * pa[row] * (1.f - pa[row]);
  }
  for (int e = threadIdx.x; e < D; e += blockDim.x) {
    d_fb[(size_t)row * D + e] = g1 * w4[D + e] + g2 * w4[2 * D + e] + g3 * w4[3 * D + e];
    if (g1 != 0.f || g2 != 0.f || g3 != 0.f) {
      const float x = fb[(size_t)row * D + e];
      atomicAdd(dw4 + D + e, g1 * x); atomicAdd(dw4 + 2 * D + e, g2 * x); atomicAdd(dw4 + 3 * D + e, g3 * x);
    }
  }
  if (threadIdx.x == 0) { atomicAdd(db4 + 1, g1); atomicAdd(db4 + 2, g2); atomicAdd(db4 + 3, g3); }
}
int localize_bwd(const float* fm, const float* fb, const float* w4, const float* pm, const float* ps, const float* pe,
                 const float* pa, const float* g_pm, const float* g_ps, const float* g_pe, const float* g_pa,
                 const uint8_t* lmask, vml_cells_t cells, float* d_fm, float* d_fb, float* dw4, float* db4, int B, vml_dims_t d,
                 cudaStream_t st) {
  static bool reg = (register_kernel("localize_bwd_cells_kernel"), register_kernel("localize_bwd_rows_kernel"), true); (void)reg;
  const int grid = min(ceil_div(cells.capacity, 8), kNumSMs * 4);
  localize_bwd_cells_kernel<<<grid, 256, sizeof(float) * d.D, st>>>(fm, w4, pm, g_pm, cells.code, cells.n_cells, d_fm, dw4, db4, d.L, d.D);
  localize_bwd_rows_kernel<<<B * d.L, 128, 0, st>>>(fb, w4, ps, pe, pa, g_ps, g_pe, g_pa, lmask, d_fb, dw4, db4, B * d.L, d.D);
  VML_LAUNCHED(2);
  return VML_OK;
}

// ------------------------------------------------------------------------------------------------
// a8 backward, pair product (models.py:292-295): operand[n, 0:D] = bu_i * bu_j
//   d_bu[b,l,:] += sum_{cells (l,j)} d_op[n,:D]*bu[b,j]  +  sum_{cells (i,l)} d_op[n,:D]*bu[b,i]
// one CTA per (b,l); the column cells are found by binary search in each (sorted) map row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pair_bwd_kernel(const float* __restrict__ d_op, int ld_op, const float* __restrict__ bu, const int32_t* __restrict__ code,
                const int32_t* __restrict__ row_start, float* __restrict__ d_bu, int L, int D, int capacity) {
  const int grow = blockIdx.x, b = grow / L, l = grow % L;
  for (int e0 = threadIdx.x * 4; e0 < D; e0 += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int n_lo = row_start[grow], n_hi = min(row_start[grow + 1], capacity);
    for (int n = n_lo; n < n_hi; ++n) {                       // row cells (l, j)
      const int j = code[n] & 0xff;
      const float4 g = *reinterpret_cast<const float4*>(d_op + (size_t)n * ld_op + e0);
      const float4 x = *reinterpret_cast<const float4*>(bu + ((size_t)b * L + j) * D + e0);
      acc.x = fmaf(g.x, x.x, acc.x); acc.y = fmaf(g.y, x.y, acc.y); acc.z = fmaf(g.z, x.z, acc.z); acc.w = fmaf(g.w, x.w, acc.w);
    }
    for (int i = 0; i < L; ++i) {                             // column cells (i, l)
      int lo = row_start[b * L + i], hi = min(row_start[b * L + i + 1], capacity);
      int found = -1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1, j = code[mid] & 0xff;
        if (j == l) { found = mid; break; }
        if (j < l) lo = mid + 1; else hi = mid;
      }
      if (found < 0) continue;
      const float4 g = *reinterpret_cast<const float4*>(d_op + (size_t)found * ld_op + e0);
      const float4 x = *reinterpret_cast<const float4*>(bu + ((size_t)b * L + i) * D + e0);
      acc.x = fmaf(g.x, x.x, acc.x); acc.y = fmaf(g.y, x.y, acc.y); acc.z = fmaf(g.z, x.z, acc.z); acc.w = fmaf(g.w, x.w, acc.w);
    }
    float4* o = reinterpret_cast<float4*>(d_bu + (size_t)grow * D + e0);
    float4 c = *o;
    c.x += acc.x; c.y += acc.y; c.z += acc.z; c.w += acc.w;
    *o = c;
  }
}
int pair_bwd(const float* d_op, int ld_op, const float* bu, vml_cells_t cells, float* d_bu, int B, vml_dims_t d, cudaStream_t st) {
  VML_CHECK_ARG(d.D % 4 == 0 && ld_op % 4 == 0);
  static bool reg = (register_kernel("pair_bwd_kernel"), true); (void)reg;
  pair_bwd_kernel<<<B * d.L, 128, 0, st>>>(d_op, ld_op, bu, cells.code, cells.row_start, d_bu, d.L, d.D, cells.capacity);
  VML_LAUNCHED(1);
  return VML_OK;
}

// ------------------------------------------------------------------------------------------------
// a6 tail backward, elementwise part (models.py:269-276 + :297):
//   dY[n,c,:] = d_cu_next[n,c,:] (or 0) + d_op[n, D + :] / C ;   d_gbar[n,:] = sum_c dY[n,c,:]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
cu_tail_bwd_kernel(const float* __restrict__ d_cu_next, const float* __restrict__ d_op, int ld_op, const int32_t* __restrict__ n_cells,
                   float* __restrict__ dY, float* __restrict__ d_gbar, int C, int D) {
  const int n_total = *n_cells;
  const float inv_c = 1.0f / (float)C;
  for (int n = blockIdx.x; n < n_total; n += gridDim.x)
    for (int e = threadIdx.x; e < D; e += blockDim.x) {
      const float m = d_op[(size_t)n * ld_op + D + e] * inv_c;
      float s = 0.f;
      for (int c = 0; c < C; ++c) {
        const size_t o = ((size_t)n * C + c) * D + e;
        const float v = (d_cu_next ? d_cu_next[o] : 0.f) + m;
        dY[o] = v;
        s += v;
      }
      d_gbar[(size_t)n * D + e] = s;
    }
}
int cu_tail_bwd(const float* d_cu_next, const float* d_op, int ld_op, vml_cells_t cells, float* dY, float* d_gbar, vml_dims_t d,
                cudaStream_t st) {
  static bool reg = (register_kernel("cu_tail_bwd_kernel"), true); (void)reg;
  cu_tail_bwd_kernel<<<min(cells.capacity, kNumSMs * 16), 128, 0, st>>>(d_cu_next, d_op, ld_op, cells.n_cells, dY, d_gbar, d.C, d.D);
  VML_LAUNCHED(1);
  return VML_OK;
}

// ------------------------------------------------------------------------------------------------
// a5+a6 (middle) backward: content-word attention, gate, CxC self-attention (models.py:207-226,253-266)
// One warp per cell (C = 4), lane owns DPL = dl/32 feature columns; the forward quantities are recomputed
// from c_hat.  Query-side gradients (w_hat, ktil, beta per word; s_hat per sample) are accumulated per CTA
// in shared memory and flushed with atomics into the gradient of the folded query projection `dq`.
// ------------------------------------------------------------------------------------------------
template <int DPL>
__global__ void __launch_bounds__(256)
content_attn_bwd_kernel(const float* __restrict__ c_hat, const float* __restrict__ d_cc, const float* __restrict__ qproj, int ld,
                        int off_what, int off_ktil, int off_beta, const float* __restrict__ s_hat, int s_ld,
                        const uint8_t* __restrict__ qmask, const int32_t* __restrict__ row_start, float* __restrict__ d_chat,
                        float* __restrict__ dq, float* __restrict__ d_shat, int L, int Nq, int capacity) {
  constexpr int C = 4, DL = DPL * 32, KS = DL + 1;
  extern __shared__ __align__(16) float sm[];
  float* s_k = sm;                      // [Nq][KS] ktil
  float* s_w = s_k + Nq * KS;           // [Nq][KS] w_hat
  float* a_k = s_w + Nq * KS;           // [Nq][DL] d ktil
  float* a_w = a_k + Nq * DL;           // [Nq][DL] d w_hat
  float* a_s = a_w + Nq * DL;           // [DL]     d s_hat
  float* a_b = a_s + DL;                // [32]     d beta
  float* s_s = a_b + 32;                // [DL]     s_hat
  float* s_b = s_s + DL;                // [32] beta, [32] mask
  float* s_c = s_b + 64;                // [warps][C][DL] c_hat staging
  float* s_d = s_c + 8 * C * DL;        // [warps][C][DL] dA staging
  float* s_p = s_d + 8 * C * DL;        // [warps][C][32] probabilities / d(raw score)
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  for (int e = tid; e < Nq * DL; e += blockDim.x) {
    const int k = e / DL, dcol = e % DL;
    const float* qr = qproj + ((size_t)b * Nq + k) * ld;
    s_k[k * KS + dcol] = qr[off_ktil + dcol];
    s_w[k * KS + dcol] = qr[off_what + dcol];
    a_k[e] = 0.f; a_w[e] = 0.f;
  }
  for (int e = tid; e < DL; e += blockDim.x) { s_s[e] = s_hat[(size_t)b * s_ld + e]; a_s[e] = 0.f; }
  if (tid < 32) {
    s_b[tid] = tid < Nq ? qproj[((size_t)b * Nq + tid) * ld + off_beta] : 0.f;
    s_b[32 + tid] = tid < Nq ? (qmask[(size_t)b * Nq + tid] ? 1.f : 0.f) : 0.f;
    a_b[tid] = 0.f;
  }
  __syncthreads();
  const int n_lo = row_start[b * L], n_hi = min(row_start[(b + 1) * L], capacity);
  float* my_c = s_c + warp * C * DL;
  float* my_d = s_d + warp * C * DL;
  float* my_p = s_p + warp * C * 32;
  const float sqrt_dl = sqrtf((float)DL), inv_sqrt = 1.0f / sqrt_dl;
  const int dcol = lane * DPL;
  for (int n = n_lo + blockIdx.x * 8 + warp; n < n_hi; n += gridDim.x * 8) {
    float ch[C][DPL], dcc[C][DPL];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int e = 0; e < DPL; ++e) {
        ch[c][e] = c_hat[((size_t)n * C + c) * DL + dcol + e];
        dcc[c][e] = d_cc[((size_t)n * C + c) * DL + dcol + e];
        my_c[c * DL + dcol + e] = ch[c][e];
      }
    __syncwarp();
    // ---- recompute forward: word softmax (lane = word), attended words, gate, clip attention -----------
    float pw[C];                                         // p[c][lane]
    {
      float sc[C] = {0.f, 0.f, 0.f, 0.f};
      if (lane < Nq) {
        const float* kr = s_k + lane * KS;
        for (int dd = 0; dd < DL; ++dd) {
          const float kv = kr[dd];
#pragma unroll
          for (int c = 0; c < C; ++c) sc[c] = fmaf(my_c[c * DL + dd], kv, sc[c]);
        }
      }
      const float mk = s_b[32 + lane], bt = s_b[lane];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float s = (sc[c] + bt) / sqrt_dl;
        s = s * mk;
        if (mk == 0.f) s = -1e9f;
        if (lane >= Nq) s = -INFINITY;
        const float mx = warp_max(s);
        const float ex = lane < Nq ? expf(s - mx) : 0.f;
        pw[c] = ex / warp_sum(ex);
        my_p[c * 32 + lane] = pw[c];
      }
    }
    __syncwarp();
    float A[C][DPL], G[C][DPL];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int e = 0; e < DPL; ++e) A[c][e] = 0.f;
    for (int k = 0; k < Nq; ++k) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float p = my_p[c * 32 + k];
#pragma unroll
        for (int e = 0; e < DPL; ++e) A[c][e] = fmaf(p, s_w[k * KS + dcol + e], A[c][e]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int e = 0; e < DPL; ++e) { A[c][e] += s_s[dcol + e]; G[c][e] = ch[c][e] * A[c][e]; }   // A now holds (A + s_hat)
    float Ac[C][C];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int c2 = c; c2 < C; ++c2) {
        float p = 0.f;
#pragma unroll
        for (int e = 0; e < DPL; ++e) p = fmaf(G[c][e], G[c2][e], p);
        p = warp_sum(p) / sqrt_dl;
        Ac[c][c2] = p; Ac[c2][c] = p;
      }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float mx = Ac[c][0];
#pragma unroll
      for (int c2 = 1; c2 < C; ++c2) mx = fmaxf(mx, Ac[c][c2]);
      float den = 0.f;
#pragma unroll
      for (int c2 = 0; c2 < C; ++c2) { Ac[c][c2] = expf(Ac[c][c2] - mx); den += Ac[c][c2]; }
#pragma unroll
      for (int c2 = 0; c2 < C; ++c2) Ac[c][c2] /= den;
    }
    // ---- backward ------------------------------------------------------------------------------------------
    float dch[C][DPL], dT[C][C];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int e = 0; e < DPL; ++e) dch[c][e] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float dAc[C], dot = 0.f;
#pragma unroll
      for (int c2 = 0; c2 < C; ++c2) {
        float p = 0.f;
#pragma unroll
        for (int e = 0; e < DPL; ++e) { p = fmaf(dcc[c][e], ch[c2][e], p); dch[c2][e] = fmaf(Ac[c][c2], dcc[c][e], dch[c2][e]); }
        dAc[c2] = warp_sum(p);
        dot = fmaf(Ac[c][c2], dAc[c2], dot);
      }
#pragma unroll
      for (int c2 = 0; c2 < C; ++c2) dT[c][c2] = Ac[c][c2] * (dAc[c2] - dot) * inv_sqrt;
    }
    float dA[C][DPL];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int e = 0; e < DPL; ++e) {
        float dG = 0.f;
#pragma unroll
        for (int c2 = 0; c2 < C; ++c2) dG = fmaf(dT[c][c2] + dT[c2][c], G[c2][e], dG);
        dch[c][e] = fmaf(dG, A[c][e], dch[c][e]);
        dA[c][e] = dG * ch[c][e];
        my_d[c * DL + dcol + e] = dA[c][e];
      }
    {                                                    // d s_hat += sum_c dA
#pragma unroll
      for (int e = 0; e < DPL; ++e) atomicAdd(&a_s[dcol + e], (dA[0][e] + dA[1][e]) + (dA[2][e] + dA[3][e]));
    }
    __syncwarp();
    // dp[c][lane] = dA[c] . w_hat[lane]  (lane = word), softmax backward -> d(raw score)
    float dr[C];
    {
      float dp[C] = {0.f, 0.f, 0.f, 0.f};
      if (lane < Nq) {
        const float* wr = s_w + lane * KS;
        for (int dd = 0; dd < DL; ++dd) {
          const float wv = wr[dd];
#pragma unroll
          for (int c = 0; c < C; ++c) dp[c] = fmaf(my_d[c * DL + dd], wv, dp[c]);
        }
      }
      const float mk = s_b[32 + lane];
      float db = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float dot = warp_sum(pw[c] * dp[c]);
        dr[c] = pw[c] * (dp[c] - dot) * inv_sqrt * mk;    // the masked score is a constant (-1e9)
        db += dr[c];
        my_p[c * 32 + lane] = dr[c];
      }
      if (lane < Nq && db != 0.f) atomicAdd(&a_b[lane], db);
    }
    __syncwarp();
    // d c_hat += sum_w dr[c][w] ktil[w];   d ktil[w] += sum_c dr[c][w] c_hat[c];   d w_hat[w] += sum_c p[c][w] dA[c]
    for (int k = 0; k < Nq; ++k) {
      float r4[C];
#pragma unroll
      for (int c = 0; c < C; ++c) r4[c] = my_p[c * 32 + k];
      float p4[C];
#pragma unroll
      for (int c = 0; c < C; ++c) p4[c] = __shfl_sync(0xffffffffu, pw[c], k);
#pragma unroll
      for (int e = 0; e < DPL; ++e) {
        const float kv = s_k[k * KS + dcol + e];
        float gk = 0.f, gw = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          dch[c][e] = fmaf(r4[c], kv, dch[c][e]);
          gk = fmaf(r4[c], ch[c][e], gk);
          gw = fmaf(p4[c], dA[c][e], gw);
        }
        atomicAdd(&a_k[k * DL + dcol + e], gk);
        atomicAdd(&a_w[k * DL + dcol + e], gw);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int e = 0; e < DPL; ++e) d_chat[((size_t)n * C + c) * DL + dcol + e] = dch[c][e];
    __syncwarp();
  }
  __syncthreads();
  if (n_lo + blockIdx.x * 8 >= n_hi) return;            // this CTA had no cells: nothing to flush
  for (int e = tid; e < Nq * DL; e += blockDim.x) {
    const int k = e / DL, dc = e % DL;
    float* qr = dq + ((size_t)b * Nq + k) * ld;
    atomicAdd(qr + off_ktil + dc, a_k[e]);
    atomicAdd(qr + off_what + dc, a_w[e]);
  }
  for (int e = tid; e < DL; e += blockDim.x) atomicAdd(d_shat + (size_t)b * s_ld + e, a_s[e]);
  if (tid < Nq) atomicAdd(dq + ((size_t)b * Nq + tid) * ld + off_beta, a_b[tid]);
}
int content_attn_bwd(const float* c_hat, const float* d_cc, const float* qproj, int ld, int off_what, int off_ktil, int off_beta,
                     const float* s_hat, int s_ld, const uint8_t* qmask, vml_cells_t cells, float* d_chat, float* dq,
                     float* d_shat, int B, vml_dims_t d, cudaStream_t st) {
  VML_CHECK_ARG(d.C == 4 && d.Nq <= 32 && (d.dl == 32 || d.dl == 64 || d.dl == 128));
  static bool reg = (register_kernel("content_attn_bwd_kernel"), true); (void)reg;
  const int DL = d.dl;
  const size_t smem = sizeof(float) * ((size_t)2 * d.Nq * (DL + 1) + (size_t)2 * d.Nq * DL + 2 * DL + 32 + 64 + 2 * 8 * 4 * DL + 8 * 4 * 32);
  const int vmax = d.L * (d.L + 1) / 2;
  int chunks = ceil_div(vmax, 8 * 4);
  while ((int64_t)chunks * B > (int64_t)kNumSMs * 16 && chunks > 1) chunks = (chunks + 1) / 2;
  dim3 grid(chunks, B);
#define VML_CAB(P)                                                                                                    \
  do {                                                                                                                \
    VML_CUDA(ensure_dyn_smem((const void*)content_attn_bwd_kernel<P>, smem));                                         \
    content_attn_bwd_kernel<P><<<grid, 256, smem, st>>>(c_hat, d_cc, qproj, ld, off_what, off_ktil, off_beta, s_hat, s_ld, qmask, \
                                                        cells.row_start, d_chat, dq, d_shat, d.L, d.Nq, cells.capacity); \
  } while (0)
  if (DL == 128) VML_CAB(4); else if (DL == 64) VML_CAB(2); else VML_CAB(1);
#undef VML_CAB
  VML_LAUNCHED(1);
  return VML_OK;
}

// ------------------------------------------------------------------------------------------------
// a7 backward, streaming part + the gate term of a6 (models.py:191-194, 272-274):  gbar = sigmoid(fm*fs)*fm
//   d_gbar_tot = d_gbar_cu[n] + A_b[i,j] * d_bu[i]                (content-unit use + f_bm use)
//   d_Ab[i,j] += d_bu[i] . gbar_ij
//   d_fm[n]    = d_mu[n] (residual) + d_gbar_tot * (s + fm*s*(1-s)*fs),   s = sigmoid(fm*fs)
//   d_fs[b]   += sum_n d_gbar_tot * fm^2 * s*(1-s)
// one CTA (128 threads) per map row (b, i).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gbar_bwd_kernel(const float* __restrict__ fm, const float* __restrict__ fs, const float* __restrict__ ab,
                const float* __restrict__ d_bu, const float* __restrict__ d_gbar_cu, const float* __restrict__ d_mu,
                const int32_t* __restrict__ code, const int32_t* __restrict__ row_start, float* __restrict__ d_ab,
                float* __restrict__ d_fm, float* __restrict__ d_fs, int L, int D, int capacity) {
  __shared__ float red[4];
  const int grow = blockIdx.x, b = grow / L;
  const int n_lo = row_start[grow], n_hi = min(row_start[grow + 1], capacity);
  if (n_lo >= n_hi) return;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int n = n_lo; n < n_hi; ++n) {
    const int j = code[n] & 0xff;
    const float a = ab[(size_t)grow * L + j];
    float dot = 0.f;
    for (int e = threadIdx.x; e < D; e += blockDim.x) {
      const float m = fm[(size_t)n * D + e], s = fs[(size_t)b * D + e];
      const float sg = sigmoidf_(m * s);
      const float gbar = sg * m;
      const float dbu = d_bu[(size_t)grow * D + e];
      dot = fmaf(dbu, gbar, dot);
      const float dg = d_gbar_cu[(size_t)n * D + e] + a * dbu;
      d_fm[(size_t)n * D + e] = d_mu[(size_t)n * D + e] + dg * (sg + m * sg * (1.f - sg) * s);
      atomicAdd(d_fs + (size_t)b * D + e, dg * m * m * sg * (1.f - sg));
    }
    dot = warp_sum(dot);
    __syncthreads();
    if (lane == 0) red[warp] = dot;
    __syncthreads();
    if (threadIdx.x == 0) d_ab[(size_t)grow * L + j] += (red[0] + red[1]) + (red[2] + red[3]);
  }
}
int gbar_bwd(const float* fm, const float* fs, const float* ab, const float* d_bu, const float* d_gbar_cu, const float* d_mu,
             vml_cells_t cells, float* d_ab, float* d_fm, float* d_fs, int B, vml_dims_t d, cudaStream_t st) {
  static bool reg = (register_kernel("gbar_bwd_kernel"), true); (void)reg;
  gbar_bwd_kernel<<<B * d.L, 128, 0, st>>>(fm, fs, ab, d_bu, d_gbar_cu, d_mu, cells.code, cells.row_start, d_ab, d_fm, d_fs, d.L, d.D,
                                          cells.capacity);
  VML_LAUNCHED(1);
  return VML_OK;
}

// ------------------------------------------------------------------------------------------------
// row softmax backward:  dS[r,w] = P[r,w] * (dP[r,w] - sum_w' P[r,w'] dP[r,w']) * scale * colmask[b,w]
// (the masked score is a constant, so its gradient is dropped).  One warp per row; rows = batch * R.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const float* __restrict__ P, const float* __restrict__ dP, const uint8_t* __restrict__ colmask, float* __restrict__ dS,
                   int rows, int R, int W, float scale) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (row >= rows) return;
  const int b = row / R;
  float dot = 0.f;
  for (int w = lane; w < W; w += 32) dot = fmaf(P[(size_t)row * W + w], dP[(size_t)row * W + w], dot);
  dot = warp_sum(dot);
  for (int w = lane; w < W; w += 32) {
    const float mk = (!colmask || colmask[(size_t)b * W + w]) ? 1.f : 0.f;
    dS[(size_t)row * W + w] = P[(size_t)row * W + w] * (dP[(size_t)row * W + w] - dot) * scale * mk;
  }
}
int softmax_bwd(const float* P, const float* dP, const uint8_t* colmask, float* dS, int batch, int R, int W, float scale, cudaStream_t st) {
  static bool reg = (register_kernel("softmax_bwd_kernel"), true); (void)reg;
  const int rows = batch * R;
  softmax_bwd_kernel<<<ceil_div(rows * 32, 256), 256, 0, st>>>(P, dP, colmask, dS, rows, R, W, scale);
  VML_LAUNCHED(1);
  return VML_OK;
}

// ------------------------------------------------------------------------------------------------
// a7 backward, gate (models.py:166-174):  G = fb * U,  U = Aq*lmask + fs
//   d_fb += dG * U ;   d_Aq = dG * fb * lmask ;   tmp = dG * fb   (its per-sample column sum is d_fs)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gate_bwd_kernel(const float* __restrict__ dG, const float* __restrict__ fb, const float* __restrict__ U, const uint8_t* __restrict__ lmask,
                float* __restrict__ d_fb, float* __restrict__ d_Aq, float* __restrict__ tmp, int64_t total, int D) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / D;
    const float g = dG[e], x = fb[e];
    d_fb[e] += g * U[e];
    const float t = g * x;
    tmp[e] = t;
    d_Aq[e] = lmask[row] ? t : 0.f;
  }
}
int gate_bwd(const float* dG, const float* fb, const float* U, const uint8_t* lmask, float* d_fb, float* d_Aq, float* tmp, int B,
             vml_dims_t d, cudaStream_t st) {
  static bool reg = (register_kernel("gate_bwd_kernel"), true); (void)reg;
  const int64_t total = (int64_t)B * d.L * d.D;
  gate_bwd_kernel<<<grid_for(total), 256, 0, st>>>(dG, fb, U, lmask, d_fb, d_Aq, tmp, total, d.D);
  VML_LAUNCHED(1);
  return VML_OK;
}

// rows of X scaled by a 0/1 byte mask (and optionally accumulated): Y[r,:] (=|+=) X[r,:] * mask[r]
__global__ void __launch_bounds__(256)
mask_rows_kernel(const float* __restrict__ X, const uint8_t* __restrict__ mask, float* __restrict__ Y, int64_t total, int D, int accumulate) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const float v = mask[e / D] ? X[e] : 0.f;
    Y[e] = accumulate ? Y[e] + v : v;
  }
}
int mask_rows(const float* X, const uint8_t* mask, float* Y, int64_t rows, int D, int accumulate, cudaStream_t st) {
  static bool reg = (register_kernel("mask_rows_kernel"), true); (void)reg;
  const int64_t total = rows * D;
  if (total <= 0) return VML_OK;
  mask_rows_kernel<<<grid_for(total), 256, 0, st>>>(X, mask, Y, total, D, accumulate);
  VML_LAUNCHED(1);
  return VML_OK;
}

// ------------------------------------------------------------------------------------------------
// a3+a4 backward (models.py:81,88-98,115-126): adjoint of span pooling + fusion.
//   g[n,c] = d_fc[n,c] + d_fm[n]/C ;  clip (n,c) covers t in [s, s+cs): difference array E[s] += w g, E[s+cs] -= w g
//   fb: snippet l covers [l r, (l+1) r): E[l r] += d_fb/r, E[(l+1) r] -= d_fb/r ;  d_f[t] = prefix sum of E
//   d_fv = d_f * fs ;  d_fs[b] += sum_t d_f[t] * fv[t]
// One CTA per (sample, 32-column slice); E lives in shared memory, one column per thread-group.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
span_pool_bwd_kernel(const float* __restrict__ d_fc, const float* __restrict__ d_fm, const float* __restrict__ d_fb,
                     const float* __restrict__ fv, const float* __restrict__ fs, const int32_t* __restrict__ code,
                     const int32_t* __restrict__ row_start, float* __restrict__ d_fv, float* __restrict__ d_fs, int T, int L, int C,
                     int D, int capacity) {
  constexpr int DSL = 32;
  extern __shared__ float E[];                           // [(T+1)][DSL]
  const int b = blockIdx.x, d0 = blockIdx.y * DSL, r = T / L;
  const int col = threadIdx.x % DSL, sub = threadIdx.x / DSL, nsub = blockDim.x / DSL;
  for (int e = threadIdx.x; e < (T + 1) * DSL; e += blockDim.x) E[e] = 0.f;
  __syncthreads();
  const float inv_r = 1.0f / (float)r, inv_C = 1.0f / (float)C;
  for (int l = sub; l < L; l += nsub) {
    const float g = d_fb[((size_t)b * L + l) * D + d0 + col] * inv_r;
    atomicAdd(&E[(l * r) * DSL + col], g);
    atomicAdd(&E[((l + 1) * r) * DSL + col], -g);
  }
  const int n_lo = row_start[b * L], n_hi = min(row_start[(b + 1) * L], capacity);
  for (int n = n_lo + sub; n < n_hi; n += nsub) {
    const int cd = code[n];
    const int i = (cd >> 8) & 0xff, j = cd & 0xff;
    const int nf = (j - i + 1) * r;
    const int cs = max(1, nf / C), nclips = min(C, nf);
    const float w = 1.0f / (float)cs;
    const float gm = d_fm[(size_t)n * D + d0 + col] * inv_C;
    for (int c = 0; c < nclips; ++c) {
      const float g = (d_fc[((size_t)n * C + c) * D + d0 + col] + gm) * w;
      const int s = i * r + c * cs;
      atomicAdd(&E[s * DSL + col], g);
      atomicAdd(&E[(s + cs) * DSL + col], -g);
    }
  }
  __syncthreads();
  if (sub == 0) {
    float run = 0.f, acc_s = 0.f;
    const float s = fs[(size_t)b * D + d0 + col];
    for (int t = 0; t < T; ++t) {
      run += E[t * DSL + col];
      const size_t o = ((size_t)b * T + t) * D + d0 + col;
      d_fv[o] = run * s;
      acc_s = fmaf(run, fv[o], acc_s);
    }
    d_fs[(size_t)b * D + d0 + col] += acc_s;
  }
}
int span_pool_bwd(const float* d_fc, const float* d_fm, const float* d_fb, const float* fv, const float* fs, vml_cells_t cells,
                  float* d_fv, float* d_fs, int B, vml_dims_t d, cudaStream_t st) {
  VML_CHECK_ARG(d.D % 32 == 0 && d.T % d.L == 0);
  static bool reg = (register_kernel("span_pool_bwd_kernel"), true); (void)reg;
  dim3 grid(B, d.D / 32);
  span_pool_bwd_kernel<<<grid, 256, sizeof(float) * (d.T + 1) * 32, st>>>(d_fc, d_fm, d_fb, fv, fs, cells.code, cells.row_start, d_fv,
                                                                          d_fs, d.T, d.L, d.C, d.D, cells.capacity);
  VML_LAUNCHED(1);
  return VML_OK;
}

// ------------------------------------------------------------------------------------------------
// fused Adam step over one flat parameter buffer (main.py:83 torch.optim.Adam defaults: no weight decay,
// no amsgrad):  m = b1 m + (1-b1) g ;  v = b2 v + (1-b2) g^2 ;  p -= lr * (m / bc1) / (sqrt(v / bc2) + eps)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
            float b1, float b2, float eps, float bc1, float bc2, float gscale) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const float gr = g[e] * gscale;
    const float mm = b1 * m[e] + (1.f - b1) * gr;
    const float vv = b2 * v[e] + (1.f - b2) * gr * gr;
    m[e] = mm; v[e] = vv;
    const float denom = sqrtf(vv) / sqrtf(bc2) + eps;
    p[e] = p[e] - (lr / bc1) * (mm / denom);
  }
}
int adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, int step, float gscale,
              cudaStream_t st) {
  static bool reg = (register_kernel("adam_kernel"), true); (void)reg;
  if (n <= 0) return VML_OK;
  const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
  adam_kernel<<<grid_for(n), 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, bc1, bc2, gscale);
  VML_LAUNCHED(1);
  return VML_OK;
}

// ------------------------------------------------------------------------------------------------
// a2 training path (models.py:48-64): packed bi-LSTM layer, forward with saved activations and BPTT.
// One CTA per (sample, direction), thread u = hidden unit u (blockDim.x == H).  fp32 throughout.
//   acts [B,Nq,2,5,H]: i, f, g, o (after the non-linearities) and c, for every processed step.
//   backward writes dgin [B*Nq, 8H] (gradient of the input projections; zero for t >= len) and dgin_rec,
//   the same but zero at each sequence's first processed step, so that dW_hh = dgin_rec^T . y_shifted is a
//   plain strided GEMM over all (sample, step) rows.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
lstm_train_fwd_kernel(const float* __restrict__ gin, const float* __restrict__ whh_t, const int32_t* __restrict__ qlen,
                      float* __restrict__ y, float* __restrict__ fs, float* __restrict__ acts, int B, int Nq, int H) {
  extern __shared__ float hs[];                          // [H]
  const int b = blockIdx.x, dir = blockIdx.y, u = threadIdx.x;
  const int len = min(qlen[b], Nq);
  const float* W = whh_t + (size_t)dir * H * 4 * H;      // [H][4H] = W_hh^T
  float c = 0.f, h = 0.f;
  hs[u] = 0.f;
  __syncthreads();
  for (int step = 0; step < len; ++step) {
    const int t = dir == 0 ? step : len - 1 - step;
    const float* gp = gin + ((size_t)b * Nq + t) * 8 * H + (size_t)dir * 4 * H + u;
    float g0 = gp[0], g1 = gp[H], g2 = gp[2 * H], g3 = gp[3 * H];
    for (int k = 0; k < H; ++k) {
      const float hv = hs[k];
      const float* wr = W + (size_t)k * 4 * H + u;
      g0 = fmaf(hv, wr[0], g0); g1 = fmaf(hv, wr[H], g1); g2 = fmaf(hv, wr[2 * H], g2); g3 = fmaf(hv, wr[3 * H], g3);
    }
    const float ig = sigmoidf_(g0), fg = sigmoidf_(g1), gg = tanhf(g2), og = sigmoidf_(g3);
    c = fg * c + ig * gg;
    h = og * tanhf(c);
    float* a = acts + (((size_t)b * Nq + t) * 2 + dir) * 5 * H + u;
    a[0] = ig; a[H] = fg; a[2 * H] = gg; a[3 * H] = og; a[4 * H] = c;
    y[((size_t)b * Nq + t) * 2 * H + (size_t)dir * H + u] = h;
    __syncthreads();
    hs[u] = h;
    __syncthreads();
  }
  for (int t = len; t < Nq; ++t) y[((size_t)b * Nq + t) * 2 * H + (size_t)dir * H + u] = 0.f;
  if (fs) fs[(size_t)b * 2 * H + (size_t)dir * H + u] = h;
}

__global__ void __launch_bounds__(1024)
lstm_train_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dfs, const float* __restrict__ whh,
                      const float* __restrict__ acts, const int32_t* __restrict__ qlen, float* __restrict__ dgin,
                      float* __restrict__ dgin_rec, int B, int Nq, int H) {
  extern __shared__ float dg_s[];                        // [4H] gate gradients of the current step
  const int b = blockIdx.x, dir = blockIdx.y, u = threadIdx.x;
  const int len = min(qlen[b], Nq);
  const float* W = whh + (size_t)dir * 4 * H * H;        // [4H][H] = W_hh
  float dh_rec = 0.f, dc_next = 0.f;
  for (int t = len; t < Nq; ++t) {
    float* o = dgin + ((size_t)b * Nq + t) * 8 * H + (size_t)dir * 4 * H + u;
    float* o2 = dgin_rec + ((size_t)b * Nq + t) * 8 * H + (size_t)dir * 4 * H + u;
#pragma unroll
    for (int q = 0; q < 4; ++q) { o[q * H] = 0.f; o2[q * H] = 0.f; }
  }
  for (int step = len - 1; step >= 0; --step) {
    const int t = dir == 0 ? step : len - 1 - step;
    const float* a = acts + (((size_t)b * Nq + t) * 2 + dir) * 5 * H + u;
    const float ig = a[0], fg = a[H], gg = a[2 * H], og = a[3 * H], c = a[4 * H];
    float c_prev = 0.f;
    if (step > 0) {
      const int tp = dir == 0 ? t - 1 : t + 1;
      c_prev = acts[(((size_t)b * Nq + tp) * 2 + dir) * 5 * H + 4 * H + u];
    }
    float dh = dy[((size_t)b * Nq + t) * 2 * H + (size_t)dir * H + u] + dh_rec;
    if (dfs && step == len - 1) dh += dfs[(size_t)b * 2 * H + (size_t)dir * H + u];   // fs = h of the last processed step
    const float tc = tanhf(c);
    const float d_o = dh * tc * og * (1.f - og);
    const float dc = dc_next + dh * og * (1.f - tc * tc);
    const float d_i = dc * gg * ig * (1.f - ig);
    const float d_g = dc * ig * (1.f - gg * gg);
    const float d_f = dc * c_prev * fg * (1.f - fg);
    dc_next = dc * fg;
    float* o = dgin + ((size_t)b * Nq + t) * 8 * H + (size_t)dir * 4 * H + u;
    float* o2 = dgin_rec + ((size_t)b * Nq + t) * 8 * H + (size_t)dir * 4 * H + u;
    o[0] = d_i; o[H] = d_f; o[2 * H] = d_g; o[3 * H] = d_o;
    const float z = step == 0 ? 0.f : 1.f;               // the first processed step has no recurrent input
    o2[0] = d_i * z; o2[H] = d_f * z; o2[2 * H] = d_g * z; o2[3 * H] = d_o * z;
    __syncthreads();
    dg_s[u] = d_i; dg_s[H + u] = d_f; dg_s[2 * H + u] = d_g; dg_s[3 * H + u] = d_o;
    __syncthreads();
    float acc = 0.f;                                     // dh_{prev}[u] = sum_r dgates[r] * W_hh[r][u]
    for (int r = 0; r < 4 * H; ++r) acc = fmaf(dg_s[r], W[(size_t)r * H + u], acc);
    dh_rec = acc;
  }
}
int lstm_train_fwd(const float* gin, const float* whh_t, const int32_t* qlen, float* y, float* fs, float* acts, int B, int Nq, int H,
                   cudaStream_t st) {
  VML_CHECK_ARG(H <= 1024 && H % 32 == 0);
  static bool reg = (register_kernel("lstm_train_fwd_kernel"), true); (void)reg;
  lstm_train_fwd_kernel<<<dim3(B, 2), H, sizeof(float) * H, st>>>(gin, whh_t, qlen, y, fs, acts, B, Nq, H);
  VML_LAUNCHED(1);
  return VML_OK;
}
int lstm_train_bwd(const float* dy, const float* dfs, const float* whh, const float* acts, const int32_t* qlen, float* dgin,
                   float* dgin_rec, int B, int Nq, int H, cudaStream_t st) {
  VML_CHECK_ARG(H <= 1024 && H % 32 == 0);
  static bool reg = (register_kernel("lstm_train_bwd_kernel"), true); (void)reg;
  lstm_train_bwd_kernel<<<dim3(B, 2), H, sizeof(float) * 4 * H, st>>>(dy, dfs, whh, acts, qlen, dgin, dgin_rec, B, Nq, H);
  VML_LAUNCHED(1);
  return VML_OK;
}

}  // namespace vml
