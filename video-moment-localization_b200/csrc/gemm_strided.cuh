// Strided, batched fp32 CUDA-core GEMM used by the backward path (and any small contraction that has no
// dedicated kernel):
//
//     C[b][m][n]  (=|+=)  alpha * sum_k A[b][m][k] * B[b][n][k]
//
// with arbitrary element strides for every operand, so that the four products of a linear layer's
// backward (dX = dY.W, dW = dY^T.X, and their batched per-sample forms) are all the same kernel with
// different strides -- no transposed copies.  The contraction length K and/or the row count M may be
// device-resident (live cell count).  `splits` > 1 cuts K into ranges handled by different CTAs whose
// partial tiles are combined with fp32 atomics (gradient accumulation; order-dependent in the last ulp).
// 64x64 tile, BK = 16, 256 threads, 4x4 micro-tile; fp32 FFMA throughout (validation-grade numerics).
#pragma once
#include "common.cuh"

namespace vml {

struct SGemm {
  const float* A; int64_t sam, sak, sab;
  const float* B; int64_t sbn, sbk, sbb;
  float* C; int64_t scm, scn, scb;
  int M, N, K, batch;
  float alpha;
  int accumulate;            // 0: C = ..., 1: C += ...   (forced to atomic add when splits > 1)
  int splits;
  const int32_t* m_dev; int m_scale;   // live M = min(M, *m_dev * m_scale)
  const int32_t* k_dev; int k_scale;   // live K = min(K, *k_dev * k_scale)
};

constexpr int SS_BM = 64, SS_BN = 64, SS_BK = 16, SS_THREADS = 256;

__global__ void __launch_bounds__(SS_THREADS)
gemm_strided_kernel(SGemm g) {
  __shared__ float As[SS_BK][SS_BM + 4];
  __shared__ float Bs[SS_BK][SS_BN + 4];
  int M = g.M, K = g.K;
  if (g.m_dev) M = min(M, *g.m_dev * g.m_scale);
  if (g.k_dev) K = min(K, *g.k_dev * g.k_scale);
  const int tiles_n = (g.N + SS_BN - 1) / SS_BN, tiles_m = (M + SS_BM - 1) / SS_BM;
  const int split = blockIdx.y, b = blockIdx.z;
  const int kper = (((K + g.splits - 1) / g.splits) + SS_BK - 1) / SS_BK * SS_BK;
  const int k_lo = split * kper, k_hi = min(K, k_lo + kper);
  if (k_lo >= k_hi && !(split == 0 && !g.accumulate && g.splits == 1)) return;
  const float* A = g.A + (int64_t)b * g.sab;
  const float* Bp = g.B + (int64_t)b * g.sbb;
  float* C = g.C + (int64_t)b * g.scb;
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  // loader mapping: along k when k is the contiguous index, along the row index otherwise
  const bool a_kfast = g.sak == 1, b_kfast = g.sbk == 1;
  for (int tile = blockIdx.x; tile < tiles_m * tiles_n; tile += gridDim.x) {
    const int m0 = (tile / tiles_n) * SS_BM, n0 = (tile % tiles_n) * SS_BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = k_lo; k0 < k_hi; k0 += SS_BK) {
      float av[4], bv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = tid + e * SS_THREADS;                  // 0..1023 over the 64 x 16 tile
        const int ar = a_kfast ? idx / SS_BK : idx % SS_BM, ak = a_kfast ? idx % SS_BK : idx / SS_BM;
        const int br = b_kfast ? idx / SS_BK : idx % SS_BN, bk = b_kfast ? idx % SS_BK : idx / SS_BN;
        av[e] = (m0 + ar < M && k0 + ak < k_hi) ? __ldg(A + (int64_t)(m0 + ar) * g.sam + (int64_t)(k0 + ak) * g.sak) : 0.f;
        bv[e] = (n0 + br < g.N && k0 + bk < k_hi) ? __ldg(Bp + (int64_t)(n0 + br) * g.sbn + (int64_t)(k0 + bk) * g.sbk) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = tid + e * SS_THREADS;
        const int ar = a_kfast ? idx / SS_BK : idx % SS_BM, ak = a_kfast ? idx % SS_BK : idx / SS_BM;
        const int br = b_kfast ? idx / SS_BK : idx % SS_BN, bk = b_kfast ? idx % SS_BK : idx / SS_BN;
        As[ak][ar] = av[e];
        Bs[bk][br] = bv[e];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SS_BK; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
        const float br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = m0 + ty * 4 + i;
      if (row >= M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + tx * 4 + j;
        if (col >= g.N) continue;
        float* c = C + (int64_t)row * g.scm + (int64_t)col * g.scn;
        const float v = g.alpha * acc[i][j];
        if (g.splits > 1) atomicAdd(c, v);
        else if (g.accumulate) *c += v;
        else *c = v;
      }
    }
  }
}

inline int launch_gemm_strided(SGemm g, cudaStream_t st) {
  VML_CHECK_ARG(g.M >= 0 && g.N > 0 && g.K >= 0 && g.batch > 0 && g.splits >= 1);
  VML_CHECK_ARG(g.splits == 1 || g.accumulate);      // split-K partial tiles are added into C: the caller pre-zeroes or accumulates
  if (g.M == 0) return VML_OK;
  static bool reg = (register_kernel("gemm_strided_kernel"), true);
  (void)reg;
  const int64_t tiles = (int64_t)ceil_div(g.M, SS_BM) * ceil_div(g.N, SS_BN);
  const int gx = (int)(tiles < 4096 ? tiles : 4096);
  dim3 grid(gx, g.splits, g.batch);
  gemm_strided_kernel<<<grid, SS_THREADS, 0, st>>>(g);
  VML_LAUNCHED(1);
  return VML_OK;
}

}  // namespace vml
