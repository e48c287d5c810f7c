// Query encoder recurrence (a2), scaled-IoU BCE loss (a10) and R@n,IoU=m evaluation (a11).
#include "common.cuh"

namespace vml {

// =====================================================================================
// a2  bi-LSTM layer recurrence with packed-sequence semantics (models.py:50-62)
// grid = (ceil(B/BT), 2 directions); thread u owns hidden unit u of BT samples.
// =====================================================================================
__global__ void query_lengths_kernel(const uint8_t* __restrict__ qmask, int32_t* __restrict__ qlen, int B, int Nq) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int s = 0;
  for (int k = 0; k < Nq; ++k) s += qmask[(size_t)b * Nq + k];
  qlen[b] = s;
}

template <int BT>
__global__ void lstm_layer_kernel(const float* __restrict__ gin, const float* __restrict__ whh_t,
                                  const int32_t* __restrict__ qlen, float* __restrict__ y, bf16* __restrict__ y16,
                                  float* __restrict__ fs, int B, int Nq, int H) {
  extern __shared__ float sh[];  // [BT][H]
  const int dir = blockIdx.y, b0 = blockIdx.x * BT, u = threadIdx.x;
  const float* W = whh_t + (size_t)dir * H * 4 * H;
  int len[BT];
  int maxlen = 0;
  float c[BT], h[BT];
#pragma unroll
  for (int s = 0; s < BT; ++s) {
    len[s] = (b0 + s < B) ? min(qlen[b0 + s], Nq) : 0;
    maxlen = max(maxlen, len[s]);
    c[s] = 0.f; h[s] = 0.f;
    sh[s * H + u] = 0.f;
  }
  __syncthreads();
  for (int step = 0; step < maxlen; ++step) {
    float acc[BT][4];
#pragma unroll
    for (int s = 0; s < BT; ++s) {
      const bool act = step < len[s];
      const int t = dir == 0 ? step : len[s] - 1 - step;
      const float* g = gin + ((size_t)(b0 + s) * Nq + (act ? t : 0)) * 8 * H + (size_t)dir * 4 * H + u;
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[s][q] = act ? g[q * H] : 0.f;
    }
    for (int k = 0; k < H; ++k) {
      const float* wr = W + (size_t)k * 4 * H + u;
      const float w0 = wr[0], w1 = wr[H], w2 = wr[2 * H], w3 = wr[3 * H];
#pragma unroll
      for (int s = 0; s < BT; ++s) {
        const float hv = sh[s * H + k];
        acc[s][0] = fmaf(hv, w0, acc[s][0]); acc[s][1] = fmaf(hv, w1, acc[s][1]);
        acc[s][2] = fmaf(hv, w2, acc[s][2]); acc[s][3] = fmaf(hv, w3, acc[s][3]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < BT; ++s) {
      if (step < len[s]) {
        const int t = dir == 0 ? step : len[s] - 1 - step;
        const float ig = sigmoidf_(acc[s][0]), fg = sigmoidf_(acc[s][1]), gg = tanhf(acc[s][2]), og = sigmoidf_(acc[s][3]);
        c[s] = fg * c[s] + ig * gg;
        h[s] = og * tanhf(c[s]);
        sh[s * H + u] = h[s];
        const size_t o = ((size_t)(b0 + s) * Nq + t) * 2 * H + (size_t)dir * H + u;
        y[o] = h[s];
        if (y16) y16[o] = __float2bfloat16_rn(h[s]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int s = 0; s < BT; ++s) {
    if (b0 + s >= B) continue;
    for (int t = len[s]; t < Nq; ++t) {   // pad_packed_sequence: zeros past the length
      const size_t o = ((size_t)(b0 + s) * Nq + t) * 2 * H + (size_t)dir * H + u;
      y[o] = 0.f;
      if (y16) y16[o] = __float2bfloat16_rn(0.f);
    }
    if (fs) fs[(size_t)(b0 + s) * 2 * H + (size_t)dir * H + u] = h[s];  // fwd: h(len-1); bwd: h(0)
  }
}

int query_lengths(const uint8_t* qmask, int32_t* qlen, int B, int Nq, cudaStream_t st) {
  static bool reg = (register_kernel("query_lengths_kernel"), true); (void)reg;
  query_lengths_kernel<<<ceil_div(B, 128), 128, 0, st>>>(qmask, qlen, B, Nq);
  VML_LAUNCHED(1);
  return VML_OK;
}

int lstm_layer(const float* gin, const float* whh_t, const int32_t* qlen, float* y, void* y16, float* fs, int B, int Nq,
               int H, cudaStream_t st) {
  VML_CHECK_ARG(H % 32 == 0 && H <= 1024);
  static bool reg = (register_kernel("lstm_layer_kernel"), true); (void)reg;
  constexpr int BT = 4;
  dim3 grid(ceil_div(B, BT), 2);
  lstm_layer_kernel<BT><<<grid, H, sizeof(float) * BT * H, st>>>(gin, whh_t, qlen, y, (bf16*)y16, fs, B, Nq, H);
  VML_LAUNCHED(1);
  return VML_OK;
}

// =====================================================================================
// a10  scaled-IoU BCE loss (main.py:89-116; BCELoss(reduction=None) read as 'none')
// block b: per-sample masked means of the four terms -> scratch[4][B]; a one-block
// finalize sums samples in index order (deterministic) and forms L_m+L_s+L_e+0.5 L_a.
// =====================================================================================
__device__ __forceinline__ float bce_elem(float p, float y) {
  const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);
  return -(y * lp + (1.f - y) * l1p);
}
// weighted two-layer form of main.py:92-94
__device__ __forceinline__ float scaled_bce(float p, float y, float s) {
  return (s * y) * bce_elem(p, y) + ((1.f - s) * (1.f - y)) * bce_elem(1.f - p, 1.f - y);
}
// d/dp, PyTorch's binary_cross_entropy_backward: (p - y) / max(p (1-p), 1e-12) * weight
__device__ __forceinline__ float bce_grad(float p, float y, float w) {
  return (p - y) / fmaxf((1.f - p) * p, 1e-12f) * w;
}

__device__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (warp == 0) {
    t = lane < (blockDim.x / 32) ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

__global__ void __launch_bounds__(256)
loss_sample_kernel(const float* __restrict__ pm, const uint8_t* __restrict__ ym, const float* __restrict__ sm,
                   const uint8_t* __restrict__ mmask, const float* __restrict__ ps, const uint8_t* __restrict__ ys,
                   const float* __restrict__ ss, const float* __restrict__ pe, const uint8_t* __restrict__ ye,
                   const float* __restrict__ se, const float* __restrict__ pa, const uint8_t* __restrict__ ya,
                   const uint8_t* __restrict__ lmask, int B, int L, float* __restrict__ scratch,
                   float* __restrict__ g_pm, float* __restrict__ g_ps, float* __restrict__ g_pe, float* __restrict__ g_pa) {
  __shared__ float red[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const size_t mo = (size_t)b * L * L, lo = (size_t)b * L;
  float lm_sum = 0.f, lm_cnt = 0.f;
  for (int e = tid; e < L * L; e += blockDim.x) {
    const float k = mmask[mo + e] ? 1.f : 0.f, y = ym[mo + e] ? 1.f : 0.f;
    lm_sum += scaled_bce(pm[mo + e], y, sm[mo + e]) * k;
    lm_cnt += k;
  }
  float ls = 0.f, le = 0.f, la = 0.f, lc = 0.f;
  for (int e = tid; e < L; e += blockDim.x) {
    const float k = lmask[lo + e] ? 1.f : 0.f;
    ls += scaled_bce(ps[lo + e], ys[lo + e] ? 1.f : 0.f, ss[lo + e]) * k;
    le += scaled_bce(pe[lo + e], ye[lo + e] ? 1.f : 0.f, se[lo + e]) * k;
    la += bce_elem(pa[lo + e], ya[lo + e] ? 1.f : 0.f) * k;
    lc += k;
  }
  lm_sum = block_sum(lm_sum, red); lm_cnt = block_sum(lm_cnt, red);
  ls = block_sum(ls, red); le = block_sum(le, red); la = block_sum(la, red); lc = block_sum(lc, red);
  if (tid == 0) {
    scratch[0 * B + b] = lm_sum / lm_cnt;
    scratch[1 * B + b] = ls / lc;
    scratch[2 * B + b] = le / lc;
    scratch[3 * B + b] = la / lc;
  }
  if (g_pm) {
    const float sc_m = 1.f / (lm_cnt * (float)B), sc_l = 1.f / (lc * (float)B);
    for (int e = tid; e < L * L; e += blockDim.x) {
      const float k = mmask[mo + e] ? 1.f : 0.f, y = ym[mo + e] ? 1.f : 0.f, s = sm[mo + e];
      g_pm[mo + e] = k != 0.f ? bce_grad(pm[mo + e], y, s * y + (1.f - s) * (1.f - y)) * sc_m : 0.f;
    }
    for (int e = tid; e < L; e += blockDim.x) {
      const float k = lmask[lo + e] ? 1.f : 0.f;
      const float y1 = ys[lo + e] ? 1.f : 0.f, y2 = ye[lo + e] ? 1.f : 0.f, y3 = ya[lo + e] ? 1.f : 0.f;
      const float s1 = ss[lo + e], s2 = se[lo + e];
      g_ps[lo + e] = k != 0.f ? bce_grad(ps[lo + e], y1, s1 * y1 + (1.f - s1) * (1.f - y1)) * sc_l : 0.f;
      g_pe[lo + e] = k != 0.f ? bce_grad(pe[lo + e], y2, s2 * y2 + (1.f - s2) * (1.f - y2)) * sc_l : 0.f;
      g_pa[lo + e] = k != 0.f ? 0.5f * bce_grad(pa[lo + e], y3, 1.f) * sc_l : 0.f;
    }
  }
}

__global__ void loss_finalize_kernel(const float* __restrict__ scratch, int B, float* __restrict__ loss, float* __restrict__ parts) {
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += scratch[threadIdx.x * B + b];
    parts[threadIdx.x] = s / (float)B;
  }
  __syncwarp();
  if (threadIdx.x == 0) loss[0] = ((parts[0] + parts[1]) + parts[2]) + 0.5f * parts[3];
}

int scaled_iou_bce(const float* pm, const uint8_t* ym, const float* sm, const uint8_t* mmask, const float* ps,
                   const uint8_t* ys, const float* ss, const float* pe, const uint8_t* ye, const float* se,
                   const float* pa, const uint8_t* ya, const uint8_t* lmask, int B, int L, float* loss, float* parts,
                   float* scratch, float* g_pm, float* g_ps, float* g_pe, float* g_pa, cudaStream_t st) {
  VML_CHECK_ARG(B > 0 && L > 0 && scratch != nullptr);
  VML_CHECK_ARG((g_pm == nullptr) == (g_ps == nullptr) && (g_pm == nullptr) == (g_pe == nullptr) && (g_pm == nullptr) == (g_pa == nullptr));
  static bool reg = (register_kernel("loss_sample_kernel"), register_kernel("loss_finalize_kernel"), true); (void)reg;
  loss_sample_kernel<<<B, 256, 0, st>>>(pm, ym, sm, mmask, ps, ys, ss, pe, ye, se, pa, ya, lmask, B, L, scratch, g_pm, g_ps, g_pe, g_pa);
  loss_finalize_kernel<<<1, 32, 0, st>>>(scratch, B, loss, parts);
  VML_LAUNCHED(2);
  return VML_OK;
}

// =====================================================================================
// a11  compute_ious (utils.py:10-31): score, top-k, IoU gather, R@n counts.
// One CTA per sample.  Scores are cached in shared memory; each of the k rounds is a block
// arg-max on 64-bit keys (score bits << 32 | ~flat index), i.e. ties -> lowest flat index;
// warp ballots compact the per-warp winners.  Optional greedy temporal NMS on the integer
// (i,j) grid: a picked proposal suppresses every proposal whose IoU with it is
// > nms_num/nms_den (exact integer arithmetic).
// =====================================================================================
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int o) {
  unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
  lo = __shfl_xor_sync(0xffffffffu, lo, o); hi = __shfl_xor_sync(0xffffffffu, hi, o);
  return ((unsigned long long)hi << 32) | lo;
}

__global__ void __launch_bounds__(256)
score_topk_kernel(const float* __restrict__ pm, const float* __restrict__ ps, const float* __restrict__ pe,
                  const uint8_t* __restrict__ mmask, const float* __restrict__ sm, int L, int k, int nms_num, int nms_den,
                  int32_t* __restrict__ top_idx, float* __restrict__ top_score, float* __restrict__ top_iou,
                  unsigned long long* __restrict__ counts) {
  extern __shared__ __align__(8) unsigned char smem_raw[];
  float* sc = reinterpret_cast<float*>(smem_raw);                       // [L*L] scores; < 0 marks taken/suppressed
  __shared__ unsigned long long wbest[8];
  __shared__ int picked[8];
  __shared__ float picked_iou[8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
  const int n = L * L;
  const size_t mo = (size_t)b * n;
  for (int e = tid; e < n; e += blockDim.x) {
    const int i = e / L, j = e % L;
    // same op order as utils.py:17-19 (fp32, IEEE sqrt): ((pm*sqrt(ps_i))*sqrt(pe_j))*mask
    float s = __fmul_rn(__fmul_rn(pm[mo + e], __fsqrt_rn(ps[(size_t)b * L + i])), __fsqrt_rn(pe[(size_t)b * L + j]));
    s = __fmul_rn(s, mmask[mo + e] ? 1.f : 0.f);
    sc[e] = s;
  }
  __syncthreads();
  const bool use_nms = nms_num < nms_den;
  for (int r = 0; r < k; ++r) {
    unsigned long long best = 0ull;
    bool any = false;
    for (int e = tid; e < n; e += blockDim.x) {
      const float s = sc[e];
      if (s >= 0.f) {  // alive (scores are >= +0)
        const unsigned long long key = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned)(0xffffffffu - (unsigned)e);
        if (!any || key > best) best = key;
        any = true;
      }
    }
    // ballot: which lanes hold a candidate at all
    const unsigned have = __ballot_sync(0xffffffffu, any);
    if (!any) best = 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long ot = shfl_xor_u64(best, o); best = ot > best ? ot : best; }
    if (lane == 0) wbest[warp] = have ? (best | 1ull << 63) : 0ull;   // bit63: valid flag (scores < 2^31 as bits)
    __syncthreads();
    if (tid == 0) {
      unsigned long long m = 0ull;
      for (int w = 0; w < (int)(blockDim.x / 32); ++w) m = wbest[w] > m ? wbest[w] : m;
      if (m == 0ull) { picked[r] = -1; picked_iou[r] = 0.f; }
      else {
        const int e = (int)(0xffffffffu - (unsigned)(m & 0xffffffffull));
        picked[r] = e;
        picked_iou[r] = sm[mo + e];
        top_idx[(size_t)b * k + r] = e;
        top_score[(size_t)b * k + r] = sc[e];
        top_iou[(size_t)b * k + r] = picked_iou[r];
      }
      if (m == 0ull) { top_idx[(size_t)b * k + r] = -1; top_score[(size_t)b * k + r] = 0.f; top_iou[(size_t)b * k + r] = 0.f; }
    }
    __syncthreads();
    const int pe_ = picked[r];
    if (pe_ < 0) continue;
    if (tid == 0) sc[pe_] = -1.f;
    if (use_nms) {
      const int bi = pe_ / L, bj = pe_ % L;
      for (int e = tid; e < n; e += blockDim.x) {
        const int i = e / L, j = e % L;
        const int inter = max(0, min(j, bj) + 1 - max(i, bi));
        const int uni = max(j, bj) + 1 - min(i, bi);
        if (uni > 0 && (long long)inter * nms_den > (long long)nms_num * uni) sc[e] = -1.f;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    const float thr[4] = {0.1f, 0.3f, 0.5f, 0.7f};
    const int ns[2] = {1, 5};
    for (int a = 0; a < 2; ++a)
      for (int t = 0; t < 4; ++t) {
        bool hit = false;
        for (int r = 0; r < min(ns[a], k); ++r) hit = hit || (picked[r] >= 0 && picked_iou[r] > thr[t]);
        if (hit) atomicAdd(&counts[a * 4 + t], 1ull);
      }
  }
}

int score_topk_recall(const float* pm, const float* ps, const float* pe, const uint8_t* mmask, const float* sm, int B,
                      int L, int k, int nms_num, int nms_den, int32_t* top_idx, float* top_score, float* top_iou,
                      int64_t* counts, cudaStream_t st) {
  VML_CHECK_ARG(B > 0 && L > 0 && k >= 1 && k <= 8 && nms_den > 0 && (size_t)L * L * 4 <= 200 * 1024);
  static bool reg = (register_kernel("score_topk_kernel"), true); (void)reg;
  const size_t smem = sizeof(float) * L * L;
  VML_CUDA(cudaFuncSetAttribute(score_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  score_topk_kernel<<<B, 256, smem, st>>>(pm, ps, pe, mmask, sm, L, k, nms_num, nms_den, top_idx, top_score, top_iou,
                                          (unsigned long long*)counts);
  VML_LAUNCHED(1);
  return VML_OK;
}

}  // namespace vml
