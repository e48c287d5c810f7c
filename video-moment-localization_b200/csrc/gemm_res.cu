// tcgen05 GEMM with a shared-memory (TMA in / TMA out) residual epilogue, bf16 fast mode:
//
//     out[M,N] = A[M,K] . W[N,K]^T + bias[N] + R1[M,N] (+ R2[M/4,N] broadcast over 4 consecutive rows)
//     side[M/4, N] = mean over the 4 rows of a group of out            (optional, with R2)
//
// used for   a6 tail  cu = cc_hat.Wc^T + bc + fc + fbar,  mean_c cu -> moment operand   (models.py:269-276,297)
//            a8 tail  mu = operand.[Wfb|Wfc]^T + (bfb+bfc) + fm                         (models.py:299-303)
//
// Both stages move far more bytes through the epilogue than through the main loop (K = 128 for a6), so
// the residual tiles travel like operands: the producer warp TMA-loads R1 (and R2) of a tile into a
// 2-deep ring of 128B-swizzled shared-memory tiles while the previous tile is being finished; the 8
// epilogue warps read TMEM + shared memory only, write the result IN PLACE into the same swizzled tile,
// and one thread TMA-stores it.  No epilogue thread ever waits on global memory.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer (128 x 128 x 16 tcgen05.mma, two
// TMEM accumulators), warps 2..9 epilogue (thread == row of the tile, two warps per TMEM lane quadrant,
// 64 columns each).
#include "common.cuh"
#include "gemm_umma.cuh"
#include "sm100.cuh"

namespace vml {

constexpr int GR_BN = 128, GR_STAGES = 3;
constexpr int GR_A_BYTES = UG_BM * UG_BK * 2, GR_B_BYTES = GR_BN * UG_BK * 2, GR_STAGE_BYTES = GR_A_BYTES + GR_B_BYTES;
constexpr int GR_R1_BYTES = UG_BM * GR_BN * 2;            // two 64-column boxes of 128 rows (16 KB each)
constexpr int GR_R2_BYTES = (UG_BM / 4) * GR_BN * 2;      // two 64-column boxes of 32 rows (4 KB each)
constexpr int GR_RES_BYTES = GR_R1_BYTES + GR_R2_BYTES;
constexpr int GR_SMEM = GR_STAGES * GR_STAGE_BYTES + 2 * GR_RES_BYTES + 1024 + 256;

template <bool HAS_R2>
__global__ void __launch_bounds__(UG_GEMM_THREADS, 1)
gemm_res_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmR1, const __grid_constant__ CUtensorMap tmR2,
                const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmSide, int M, int N, int K,
                const int32_t* __restrict__ m_dev, int m_scale, const float* __restrict__ bias) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array: keeps the shared address space visible to the
  // compiler (LDS/STS instead of generic LD/ST, which an integer round-trip of the pointer would force)
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* res = smem + GR_STAGES * GR_STAGE_BYTES;                  // [2][R1 | R2]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(res + 2 * GR_RES_BYTES);
  uint64_t* empty_bar = full_bar + GR_STAGES;
  uint64_t* tfull_bar = empty_bar + GR_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* rfull_bar = tempty_bar + 2;
  uint64_t* rempty_bar = rfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rempty_bar + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (m_dev) M = min(M, *m_dev * m_scale);
  const int tiles_m = (M + UG_BM - 1) / UG_BM, tiles_n = N / GR_BN;
  const int num_tiles = tiles_m * tiles_n;
  const int k_blocks = (K + UG_BK - 1) / UG_BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA); ptx::prefetch_tensormap(&tmB); ptx::prefetch_tensormap(&tmR1); ptx::prefetch_tensormap(&tmOut);
    if (HAS_R2) { ptx::prefetch_tensormap(&tmR2); ptx::prefetch_tensormap(&tmSide); }
    for (int s = 0; s < GR_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], UG_EPI_WARPS);
      ptx::mbar_init(&rfull_bar[a], 1); ptx::mbar_init(&rempty_bar[a], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<256>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: residual tiles first, then the operand ring =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int rb = 0; uint32_t rphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * UG_BM, n0 = (tile % tiles_n) * GR_BN;
        ptx::mbar_wait(&rempty_bar[rb], rphase ^ 1);               // the tile's previous contents have been stored
        unsigned char* r1 = res + rb * GR_RES_BYTES;
        ptx::mbar_arrive_expect_tx(&rfull_bar[rb], HAS_R2 ? GR_RES_BYTES : GR_R1_BYTES);
        ptx::tma_load_2d(r1, &tmR1, &rfull_bar[rb], n0, m0);
        ptx::tma_load_2d(r1 + GR_R1_BYTES / 2, &tmR1, &rfull_bar[rb], n0 + 64, m0);
        if (HAS_R2) {
          ptx::tma_load_2d(r1 + GR_R1_BYTES, &tmR2, &rfull_bar[rb], n0, m0 / 4);
          ptx::tma_load_2d(r1 + GR_R1_BYTES + GR_R2_BYTES / 2, &tmR2, &rfull_bar[rb], n0 + 64, m0 / 4);
        }
        if (++rb == 2) { rb = 0; rphase ^= 1; }
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = smem + stage * GR_STAGE_BYTES;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], GR_STAGE_BYTES);
          ptx::tma_load_2d(sa, &tmA, &full_bar[stage], kb * UG_BK, m0);
          ptx::tma_load_2d(sa + GR_A_BYTES, &tmB, &full_bar[stage], kb * UG_BK, n0);
          if (++stage == GR_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(UG_BM, GR_BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * GR_BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + stage * GR_STAGE_BYTES);
          const uint64_t adesc = ptx::umma_desc_sw128(a_addr), bdesc = ptx::umma_desc_sw128(a_addr + GR_A_BYTES);
#pragma unroll
          for (int k = 0; k < UG_BK / 16; ++k)
            ptx::umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == GR_STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quad = warp % 4;                             // TMEM lane quadrant this warp may access
    const int half = (warp - 2) / 4;                       // which 64-column box of the tile
    const int r = quad * 32 + lane;                        // row of the tile
    const bool storer = threadIdx.x == 64;
    int acc = 0; uint32_t acc_phase = 0;
    int rb = 0; uint32_t rphase = 0;
    int pending = -1;                                      // ring slot whose TMA store has been issued but not yet read out
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * UG_BM, n0 = (tile % tiles_n) * GR_BN;
      if (storer && pending >= 0) {                        // hand the previous tile's slot back to the producer
        ptx::bulk_wait_read<0>();
        ptx::mbar_arrive(&rempty_bar[pending]);
        pending = -1;
      }
      unsigned char* r1 = res + rb * GR_RES_BYTES + half * (GR_R1_BYTES / 2);
      unsigned char* r2 = res + rb * GR_RES_BYTES + GR_R1_BYTES + half * (GR_R2_BYTES / 2);
      const float* bcol = bias + n0 + half * 64;
      ptx::mbar_wait(&rfull_bar[rb], rphase);
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * GR_BN + half * 64);
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        float v[32];
        ptx::tmem_ld32(t_addr + (uint32_t)c, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          const int piece = (c + e) >> 3;
          uint4* px = reinterpret_cast<uint4*>(r1 + ptx::sw128_off(r, piece));
          const f8 xv = unpack8(*px);
          f8 o;
          if (HAS_R2) {
            const f8 fv = unpack8(*reinterpret_cast<const uint4*>(r2 + ptx::sw128_off(r >> 2, piece)));
#pragma unroll
            for (int q = 0; q < 8; ++q) o.v[q] = (v[e + q] + __ldg(bcol + c + e + q)) + xv.v[q] + fv.v[q];
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) o.v[q] = (v[e + q] + __ldg(bcol + c + e + q)) + xv.v[q];
          }
          uint4 packed;
          {
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&packed);
#pragma unroll
            for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(o.v[2 * q], o.v[2 * q + 1]);
          }
          *px = packed;                                    // result in place of the residual
          if (HAS_R2) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {                  // mean over the group's 4 rows (adjacent lanes)
              float s = o.v[q];
              s += __shfl_xor_sync(0xffffffffu, s, 1);
              s += __shfl_xor_sync(0xffffffffu, s, 2);
              o.v[q] = s * 0.25f;
            }
            if ((lane & 3) == 0) {                         // all 4 lanes have read this piece of R2 (same instruction stream)
              uint4 pm;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pm);
#pragma unroll
              for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(o.v[2 * q], o.v[2 * q + 1]);
              *reinterpret_cast<uint4*>(r2 + ptx::sw128_off(r >> 2, piece)) = pm;
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);   // accumulator drained
      ptx::fence_proxy_async();                            // shared-memory writes -> visible to the TMA store
      asm volatile("bar.sync 1, 256;" ::: "memory");       // all 8 epilogue warps have finished the tile
      if (storer) {
        unsigned char* base = res + rb * GR_RES_BYTES;
        ptx::tma_store_2d(&tmOut, base, n0, m0);
        ptx::tma_store_2d(&tmOut, base + GR_R1_BYTES / 2, n0 + 64, m0);
        if (HAS_R2) {
          ptx::tma_store_2d(&tmSide, base + GR_R1_BYTES, n0, m0 / 4);
          ptx::tma_store_2d(&tmSide, base + GR_R1_BYTES + GR_R2_BYTES / 2, n0 + 64, m0 / 4);
        }
        ptx::bulk_commit();
        pending = rb;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (++rb == 2) { rb = 0; rphase ^= 1; }
    }
    if (storer && pending >= 0) ptx::bulk_wait_read<0>();         // shared memory must outlive the last store's reads
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc<256>(tmem_base); }
}

// A bf16 [M,K] (lda), W bf16 [N,K], R1/out bf16 [M,N] (ld_out), R2 bf16 [M/4,N] (ld N), side bf16 [M/4, N] with
// row stride ld_side (may point into a wider matrix).  M = capacity; live rows = *m_dev * m_scale.
int gemm_res(const void* A, const void* W, const float* bias, const void* R1, const void* R2, void* out, void* side, int M, int N,
             int K, int lda, int ld_side, const int32_t* m_dev, int m_scale, cudaStream_t st) {
  VML_CHECK_ARG(N % GR_BN == 0 && K % 8 == 0 && lda % 8 == 0 && M > 0 && (R2 == nullptr) == (side == nullptr));
  VML_CHECK_ARG(R2 == nullptr || (M % 4 == 0 && ld_side % 8 == 0));
  static bool reg = (register_kernel("gemm_res_kernel"), true); (void)reg;
  CUtensorMap tmA, tmB, tmR1, tmR2, tmOut, tmSide;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, UG_BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)K, GR_BN))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmR1, R1, (uint64_t)M, (uint64_t)N, (uint64_t)N, UG_BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmOut, out, (uint64_t)M, (uint64_t)N, (uint64_t)N, UG_BM))) return rc;
  tmR2 = tmR1; tmSide = tmOut;
  if (R2) {
    if ((rc = make_tmap_bf16_2d(&tmR2, R2, (uint64_t)(M / 4), (uint64_t)N, (uint64_t)N, UG_BM / 4))) return rc;
    if ((rc = make_tmap_bf16_2d(&tmSide, side, (uint64_t)(M / 4), (uint64_t)N, (uint64_t)ld_side, UG_BM / 4))) return rc;
  }
  const int64_t tiles = (int64_t)ceil_div(M, UG_BM) * (N / GR_BN);
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  if (R2) {
    VML_CUDA(ensure_dyn_smem((const void*)(gemm_res_kernel<true>), (size_t)(GR_SMEM)));
    gemm_res_kernel<true><<<grid, UG_GEMM_THREADS, GR_SMEM, st>>>(tmA, tmB, tmR1, tmR2, tmOut, tmSide, M, N, K, m_dev, m_scale, bias);
  } else {
    VML_CUDA(ensure_dyn_smem((const void*)(gemm_res_kernel<false>), (size_t)(GR_SMEM)));
    gemm_res_kernel<false><<<grid, UG_GEMM_THREADS, GR_SMEM, st>>>(tmA, tmB, tmR1, tmR2, tmOut, tmSide, M, N, K, m_dev, m_scale, bias);
  }
  VML_LAUNCHED(1);
  return VML_OK;
}

}  // namespace vml
