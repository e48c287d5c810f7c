// CUDA-core fp32 GEMM:  out = A[M,K] . W[N,K]^T  (+ epilogue).  This is the fp32 VALIDATION
// path (north_star: scores within 1e-5 of the reference; TF32/bf16 tensor-core math cannot
// meet that) and the small-M path.  64x64 tile, BK=16, 256 threads, 4x4 micro-tile.
#pragma once
#include "common.cuh"

namespace vml {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

template <typename Epi>
__global__ void __launch_bounds__(SG_THREADS)
gemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ W, int M, int N, int K,
                 const int32_t* __restrict__ m_dev, int m_scale, Epi epi) {
  __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Ws[SG_BK][SG_BN + 4];
  if (m_dev) M = min(M, *m_dev * m_scale);
  const int n_tiles_n = (N + SG_BN - 1) / SG_BN;
  const int n_tiles_m = (M + SG_BM - 1) / SG_BM;
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;
  const int lrow = tid / 4, lk = (tid % 4) * 4;  // loader: 64 rows x 4 float4

  for (int tile = blockIdx.x; tile < n_tiles_m * n_tiles_n; tile += gridDim.x) {
    const int m0 = (tile / n_tiles_n) * SG_BM, n0 = (tile % n_tiles_n) * SG_BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += SG_BK) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), w = a;
      if (m0 + lrow < M && k0 + lk < K) a = *reinterpret_cast<const float4*>(A + (size_t)(m0 + lrow) * K + k0 + lk);
      if (n0 + lrow < N && k0 + lk < K) w = *reinterpret_cast<const float4*>(W + (size_t)(n0 + lrow) * K + k0 + lk);
      __syncthreads();
      As[lk][lrow] = a.x; As[lk + 1][lrow] = a.y; As[lk + 2][lrow] = a.z; As[lk + 3][lrow] = a.w;
      Ws[lk][lrow] = w.x; Ws[lk + 1][lrow] = w.y; Ws[lk + 2][lrow] = w.z; Ws[lk + 3][lrow] = w.w;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SG_BK; ++k) {
        float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        float4 wv = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
        const float ar[4] = {av.x, av.y, av.z, av.w};
        const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
      }
    }
    const int col = n0 + tx * 4;
    if (col < N) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = m0 + ty * 4 + i;
        if (row < M) epi.template apply<4>(row, col, acc[i]);
      }
    }
  }
}

// M is the capacity (upper bound); live rows = *m_dev * m_scale when m_dev != nullptr.
template <typename Epi>
int launch_gemm_simt(const float* A, const float* W, int M, int N, int K, const int32_t* m_dev, int m_scale,
                     const Epi& epi, cudaStream_t stream) {
  VML_CHECK_ARG(K % 4 == 0 && N % 4 == 0);
  if (M <= 0) return VML_OK;
  static bool reg = (register_kernel("gemm_simt_kernel"), true);
  (void)reg;
  int64_t tiles = (int64_t)ceil_div(M, SG_BM) * ceil_div(N, SG_BN);
  int grid = (int)(tiles < (int64_t)kNumSMs * 8 ? tiles : (int64_t)kNumSMs * 8);
  gemm_simt_kernel<Epi><<<grid, SG_THREADS, 0, stream>>>(A, W, M, N, K, m_dev, m_scale, epi);
  VML_LAUNCHED(1);
  return VML_OK;
}

}  // namespace vml
