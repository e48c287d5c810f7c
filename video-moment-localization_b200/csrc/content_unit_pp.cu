// a5+a6, PING-PONG schedule of the one-kernel content unit (content_unit.cu, V = 4): the same arithmetic on the same
// operands in the same order per element (bit-identical outputs), but the 16 row warps form TWO groups of 8 that take
// alternate tiles.  A tile's life in the row warps is an ATTENTION half (c_hat out of TMEM, word scores, softmax,
// attended words, gate, Gram, 4x4 clip attention: ~5 us of dependent phases) and a TAIL half (four column blocks: wait
// for the tensor cores, add fbar, round, store: ~6 us, most of it waiting).  With one group the halves of consecutive
// tiles run back to back (11.5-12.5 us per tile, ncu: 65 % of the cycles no warp is eligible); with two groups the tail
// of tile t runs under the attention half of tile t + 1.
//
//   reference: ContentUnit.forward models.py:242-276, ContentAttention.forward models.py:207-226, :297
//
// What is shared and how it is handed over:
//   * attention operands U = Ks | Wt | Ps, the S and A accumulators, the c_hat accumulator: ONE copy; attention halves are
//     mutually exclusive -- group g starts tile t's only after cc_ready(t - 1) (the other group's cc_hat is complete);
//   * Cs (c_hat, then cc_hat, read by the tail MMAs): one per group;
//   * Y: two 64-column accumulators (a column block = two output boxes, the MMAs of one run under the epilogue of the
//     other), mean_c accumulator: 64 columns (one output box at a time);
//   * the box ring (8 boxes beside two Cs); tile t + 2's first contraction is issued between the column blocks of tile
//     t's tail (see the program order in the kernel).
// TMEM columns: [0,128) c_hat, [128,256) A, [256,320) Y0, [320,384) Y1, [384,416) S, [416,480) mean_c.
//
// Warp roles (672 threads): 0 TMA producer, 1 MMA issuer of (1) and (4), 2..9 row group 0, 10..17 row group 1 (thread
// pair == tile row: TMEM lane quadrant warp % 4, column half (warp - 2) % 8 / 4), 18 store + mean_c MMAs, 20 mean_c drain.
#include <stdlib.h>

#include "common.cuh"
#include "gemm_umma.cuh"
#include "sm100.cuh"

namespace vml {

#ifdef VML_CU_TIMING
__device__ long long g_pp_dbg[8 * 64];
__device__ __forceinline__ long long pp_now() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define PP_T(t, slot) do { if (blockIdx.x == 0 && (t) < 8) g_pp_dbg[(t) * 64 + (slot)] = pp_now(); } while (0)
int pp_debug_read(long long* host, int n) { return (int)cudaMemcpyFromSymbol(host, g_pp_dbg, sizeof(long long) * n); }
#else
#define PP_T(t, slot) do { } while (0)
#endif

constexpr int PP_DL = 128;
constexpr int PP_GROUP_THREADS = 256, PP_ROW_WARPS = 16;
constexpr int PP_STORE_WARP = 2 + PP_ROW_WARPS, PP_DRAIN_WARP = 20, PP_THREADS = 32 * (PP_DRAIN_WARP + 1);
constexpr int PP_BOX = UG_BM * UG_BK * 2;                 // one 128 x 64 bf16 box (16 KB)
constexpr int PP_CS_BYTES = UG_BM * PP_DL * 2;
constexpr int PP_TM_A = 128, PP_TM_Y = 256, PP_TM_S = 384, PP_TM_SIDE = 416;
constexpr int PP_M4_ROWS = 252, PP_M4_BYTES = 2 * PP_M4_ROWS * 16;
constexpr int PP_BAR_BYTES = 512;
constexpr int PP_COLS = 64;                               // columns of a 128-column block a row thread owns

template <int NQP, int GS>
struct PpCfg {
  static constexpr int NW = GS * NQP, KG = NW / 8;
  static constexpr int KS_BYTES = NW * PP_DL * 2, WT_BYTES = NW * PP_DL * 2, PS_BYTES = UG_BM * NW * 2;
  static constexpr int GG_BYTES = 2 * UG_BM * 16;
  static constexpr int U_RAW = KS_BYTES + WT_BYTES + (PS_BYTES > GG_BYTES ? PS_BYTES : GG_BYTES);
  static constexpr int U_BYTES = (U_RAW + 1023) / 1024 * 1024;
  static constexpr int SIDE_FLOATS = 2 * NW + PP_DL + 128 + PP_M4_BYTES / 4;      // beta | mask | b1 | I_16 | M4
  static constexpr int FIXED = 2 * PP_CS_BYTES + U_BYTES + SIDE_FLOATS * 4 + 1024 + PP_BAR_BYTES;
  static constexpr int RB = (232448 - FIXED) / PP_BOX;
  static constexpr int SMEM = RB * PP_BOX + FIXED;
  static_assert(RB >= 6 && RB <= 12, "box ring");
};

__device__ __forceinline__ uint4 pp_pack8(const float* v) {
  uint4 u; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return u;
}

template <int NQP, int GS>
__global__ void __launch_bounds__(PP_THREADS, 1)
content_unit_pp_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                       const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut, int D,
                       const bf16* __restrict__ fbar, bf16* __restrict__ side, int ld_side, const float* __restrict__ bias1,
                       const float* __restrict__ qproj, int ld, int off_what, int off_ktil, int off_beta,
                       const float* __restrict__ s_hat, int s_ld, const uint8_t* __restrict__ qmask,
                       const int32_t* __restrict__ code, const int32_t* __restrict__ n_cells, int Nq, int B, int store_cu) {
  using Cfg = PpCfg<NQP, GS>;
  constexpr int NW = Cfg::NW, KG = Cfg::KG, RB = Cfg::RB;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* Ring = smem;                                  // RB boxes
  unsigned char* Cs0 = Ring + RB * PP_BOX;                     // c_hat / cc_hat of group 0's tile, then group 1's
  unsigned char* U = Cs0 + 2 * PP_CS_BYTES;
  unsigned char* Ks = U;
  unsigned char* Wt = Ks + Cfg::KS_BYTES;
  unsigned char* Ps = Wt + Cfg::WT_BYTES;
  float* s_beta = reinterpret_cast<float*>(U + Cfg::U_BYTES);  // [NW]
  float* s_mask = s_beta + NW;                                 // [NW]
  float* s_b1 = s_mask + NW;                                   // [128]
  unsigned char* I16 = reinterpret_cast<unsigned char*>(s_b1 + PP_DL);      // bf16 16x16 identity (512 B)
  unsigned char* M4 = I16 + 512;                               // master pattern of the mean_c operand (content_unit.cu)
  uint64_t* bars = reinterpret_cast<uint64_t*>(M4 + PP_M4_BYTES);
  uint64_t* chat_full = bars;              // the first contraction of a tile has completed
  uint64_t* chat_free = bars + 1;          // ... and has been read out of TMEM by the tile's group
  uint64_t* cc_ready = bars + 2;           // cc_hat(t) is in Cs: the tail MMAs may start AND the other group may start its attention half
  uint64_t* sfull_bar = bars + 3;
  uint64_t* afull_bar = bars + 4;
  uint64_t* yfull = bars + 5;              // [2] one per 64-column half accumulator
  uint64_t* yempty = bars + 7;             // [2]
  uint64_t* sready = bars + 9;             // [2] an output box is finished in shared memory -> store warp
  uint64_t* side_full = bars + 11;
  uint64_t* side_empty = bars + 12;
  uint64_t* tail_done = bars + 13;         // a group has finished the epilogue of its tile: the other group's epilogue may start.
                                           // (yfull / sready are waited on by parity: a waiter must never be two phases behind)
  uint64_t* bfull = bars + 14;             // [RB]
  uint64_t* bempty = bfull + 12;           // [RB]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bempty + 12);
  static_assert((14 + 12 + 12) * 8 + 8 <= PP_BAR_BYTES, "barrier block");

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int M = *n_cells * 4;
  const int num_tiles = (M + UG_BM - 1) / UG_BM;
  const int KB = D / UG_BK, NB = D / 128;
  const int tiles_per_cta = (num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile_begin = min((int)blockIdx.x * tiles_per_cta, num_tiles), tile_end = min(tile_begin + tiles_per_cta, num_tiles);
  const int n_my = tile_end - tile_begin, PER = 2 * KB;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmX); ptx::prefetch_tensormap(&tmW1); ptx::prefetch_tensormap(&tmW2);
    ptx::prefetch_tensormap(&tmOut);
    ptx::mbar_init(chat_full, 1);
    ptx::mbar_init(chat_free, PP_GROUP_THREADS);
    ptx::mbar_init(cc_ready, PP_GROUP_THREADS);
    ptx::mbar_init(sfull_bar, 1);
    ptx::mbar_init(afull_bar, 1);
    for (int j = 0; j < 2; ++j) {
      ptx::mbar_init(&yfull[j], 1);
      ptx::mbar_init(&yempty[j], PP_ROW_WARPS / 2);
      ptx::mbar_init(&sready[j], PP_GROUP_THREADS);
    }
    ptx::mbar_init(side_full, 1);
    ptx::mbar_init(side_empty, 1);
    ptx::mbar_init(tail_done, PP_GROUP_THREADS);
    for (int i = 0; i < 12; ++i) { ptx::mbar_init(&bfull[i], 1); ptx::mbar_init(&bempty[i], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Program order of the ring's boxes, the same in the producer, the MMA issuer, the store warp and the row threads:
  //   main(0), main(1), then for t = 0 .. n-1:  tail(t) block 0, block 1, [main(t+2) k-blocks], block 2, [k-blocks], ...
  // i.e. tile t + 2's first contraction is issued BETWEEN the column blocks of tile t's tail (its accumulator is free as
  // soon as group (t + 1) % 2 has parked c_hat(t + 1), right after tail(t) starts), so that c_hat(t + 2) is ready when
  // tile t's group comes back for its next attention half.  main boxes: X(kb), W1(kb); tail block: W2(nb, 0), W2(nb, 1),
  // X(2nb), X(2nb + 1).
  const int n_first = min(n_my, 2);
  auto kb_hi = [&](int nb) { return nb < 0 ? 0 : (NB == 1 ? KB : KB * nb / (NB - 1)); };      // main k-blocks issued after blocks 0 .. nb
  auto has2 = [&](int t) { return t + 2 < n_my; };
  auto base_iter = [&](int t) { return PER * n_first + 4 * NB * t + PER * min(t, max(n_my - 2, 0)); };
  auto box_tail = [&](int t, int nb) { return base_iter(t) + 4 * nb + (has2(t) ? 2 * kb_hi(nb - 1) : 0); };   // first of the block's 4 boxes
  auto box_main2 = [&](int t, int nb, int kb) { return base_iter(t) + 4 * (nb + 1) + 2 * kb; };  // k-block kb of main(t + 2), issued after block nb of tail(t)

  if (warp == 0) {
    if (lane == 0) {                                    // ===================== TMA producer =====================
      const uint64_t keep = ptx::l2_policy_evict_last(), once = ptx::l2_policy_evict_first();
      auto load = [&](int bi, const CUtensorMap* tm, int c0, int c1, uint64_t policy) {
        const int slot = bi % RB;
        VML_DBG_ASSERT(bi >= 0 && slot < RB && c1 >= 0);
        ptx::mbar_wait(&bempty[slot], ((bi / RB) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&bfull[slot], PP_BOX);
        ptx::tma_load_2d_hint(Ring + slot * PP_BOX, tm, &bfull[slot], c0, c1, policy);
      };
      auto main_kb = [&](int t, int kb, int bi) {
        const int m0 = (tile_begin + t) * UG_BM;
        load(bi, &tmX, kb * UG_BK, m0, keep);
        if (kb == 0) PP_T(t, 40);
        load(bi + 1, &tmW1, kb * UG_BK, 0, keep);
        if (kb == KB - 1) PP_T(t, 41);
      };
      for (int t = 0; t < n_first; ++t)
        for (int kb = 0; kb < KB; ++kb) main_kb(t, kb, PER * t + 2 * kb);
      for (int t = 0; t < n_my; ++t) {
        const int m0 = (tile_begin + t) * UG_BM;
        for (int nb = 0; nb < NB; ++nb) {
          const int bt = box_tail(t, nb);
          load(bt, &tmW2, 0, nb * 128, keep);
          if (nb == 0) PP_T(t, 42);
          load(bt + 1, &tmW2, UG_BK, nb * 128, keep);
          load(bt + 2, &tmX, (2 * nb) * UG_BK, m0, once);
          load(bt + 3, &tmX, (2 * nb + 1) * UG_BK, m0, once);
          if (nb == NB - 1) PP_T(t, 43);
          if (has2(t))
            for (int kb = kb_hi(nb - 1); kb < kb_hi(nb); ++kb) main_kb(t + 2, kb, box_main2(t, nb, kb));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                    // ===================== MMA issuer: (1) and (4) =====================
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(UG_BM, PP_DL);
      constexpr uint32_t idesc_h = ptx::umma_idesc_bf16(UG_BM, 64);
      constexpr uint32_t idesc_r = ptx::umma_idesc_bf16(UG_BM, 16);
      const uint64_t idn = ptx::umma_desc_nosw(ptx::smem_u32(I16), 128, 256);
      auto wait_box = [&](int bi) {
        const int slot = bi % RB;
        ptx::mbar_wait(&bfull[slot], (bi / RB) & 1);
        return slot;
      };
      auto main_kb = [&](int t, int kb, int bi) {
        if (kb == 0) {
          if (t > 0) ptx::mbar_wait(chat_free, (uint32_t)(t - 1) & 1);   // c_hat(t - 1) has left the accumulator
          ptx::tc_fence_after();
          PP_T(t, 16);
        }
        const int sx = wait_box(bi), sw = wait_box(bi + 1);
        ptx::tc_fence_after();
        const uint64_t adesc = ptx::umma_desc_sw128(ptx::smem_u32(Ring + sx * PP_BOX));
        const uint64_t bdesc = ptx::umma_desc_sw128(ptx::smem_u32(Ring + sw * PP_BOX));
#pragma unroll
        for (int k = 0; k < UG_BK / 16; ++k)
          ptx::umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        ptx::umma_commit(&bempty[sx]);
        ptx::umma_commit(&bempty[sw]);
        if (kb == KB - 1) { ptx::umma_commit(chat_full); PP_T(t, 17); }
      };
      for (int t = 0; t < n_first; ++t)
        for (int kb = 0; kb < KB; ++kb) main_kb(t, kb, PER * t + 2 * kb);
      for (int t = 0; t < n_my; ++t) {
        ptx::mbar_wait(cc_ready, (uint32_t)t & 1);       // cc_hat(t) is in its group's Cs
        ptx::tc_fence_after();
        PP_T(t, 18);
        const uint32_t c0 = ptx::smem_u32(Cs0 + (t & 1) * PP_CS_BYTES);
        for (int nb = 0; nb < NB; ++nb) {
          const uint32_t c = (uint32_t)(t * NB + nb);    // use counter of each half accumulator
          const int bt = box_tail(t, nb);
          int sw[2];
          for (int j = 0; j < 2; ++j) sw[j] = wait_box(bt + j);
#pragma unroll 1
          for (int j = 0; j < 2; ++j) {                  // output columns [128 nb + 64 j, + 64): X box 2nb + j, W2 rows 64 j .. of the block
            ptx::mbar_wait(&yempty[j], (c & 1) ^ 1);     // the previous block's half j has been read out of the accumulator
            if (j == 0) PP_T(t, 19 + 3 * nb);
            const int sx = wait_box(bt + 2 + j);
            if (j == 0) PP_T(t, 20 + 3 * nb);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + PP_TM_Y + (uint32_t)(64 * j);
            for (int kh = 0; kh < 2; ++kh) {
              const uint64_t bdesc = ptx::umma_desc_sw128(ptx::smem_u32(Ring + sw[kh] * PP_BOX + j * (64 * 128)));
#pragma unroll
              for (int k = 0; k < UG_BK / 16; ++k)
                ptx::umma_bf16(d_tmem, ptx::umma_desc_nosw(c0 + (uint32_t)((kh * 4 + k) * 256), 128, 2048), bdesc + (uint64_t)(k * 2),
                               idesc_h, (kh | k) != 0);
            }
            {                                            // residual: Y[:, 16k ..+16) += X_box[:, 16k ..+16) . I_16^T
              const uint64_t xdesc = ptx::umma_desc_sw128(ptx::smem_u32(Ring + sx * PP_BOX));
#pragma unroll
              for (int k = 0; k < UG_BK / 16; ++k)
                ptx::umma_bf16(d_tmem + (uint32_t)(16 * k), xdesc + (uint64_t)(k * 2), idn, idesc_r, true);
            }
            if (j == 1) { ptx::umma_commit(&bempty[sw[0]]); ptx::umma_commit(&bempty[sw[1]]); }
            ptx::umma_commit(&yfull[j]);
            if (j == 1) PP_T(t, 21 + 3 * nb);
          }
          if (has2(t))
            for (int kb = kb_hi(nb - 1); kb < kb_hi(nb); ++kb) main_kb(t + 2, kb, box_main2(t, nb, kb));
        }
      }
    }
  } else if (warp == PP_STORE_WARP) {
    if (lane == 0) {                                    // ===================== store warp + mean_c MMAs =====================
      uint32_t sc = 0;                                   // 64-column box counter
      constexpr uint32_t idesc_m = ptx::umma_idesc_bf16_bmn(UG_BM, 64);
      const uint32_t m4 = ptx::smem_u32(M4);
      for (int t = 0; t < n_my; ++t) {
        const int m0 = (tile_begin + t) * UG_BM;
        for (int nb = 0; nb < NB; ++nb) {
          const uint32_t c = (uint32_t)(t * NB + nb);
#pragma unroll 1
          for (int j = 0; j < 2; ++j, ++sc) {
            ptx::mbar_wait(&sready[j], c & 1);
            if (j == 0) PP_T(t, 32 + 2 * nb);
            const int sx = (box_tail(t, nb) + 2 + j) % RB;
            const unsigned char* bx = Ring + sx * PP_BOX;
            if (store_cu) {
              ptx::tma_store_2d_hint(&tmOut, bx, nb * 128 + 64 * j, m0, ptx::l2_policy_evict_first());
              ptx::bulk_commit();
            }
            // mean over the cell's 4 clips of the finished box on the tensor cores
            ptx::mbar_wait(side_empty, (sc & 1) ^ 1);
            ptx::tc_fence_after();
            const uint32_t bb = ptx::smem_u32(bx);
#pragma unroll
            for (int ks = 0; ks < UG_BM / 16; ++ks)
              ptx::umma_bf16(tmem_base + PP_TM_SIDE, ptx::umma_desc_nosw(m4 + (uint32_t)((124 - 4 * ks) * 16), PP_M4_ROWS * 16, 128),
                             ptx::umma_desc_sw128_mn(bb + (uint32_t)(ks * 2048)), idesc_m, ks != 0);
            ptx::umma_commit(side_full);
            if (store_cu) ptx::bulk_wait_read<0>();
            ptx::mbar_wait(side_full, sc & 1);           // the MMAs have read the box too
            ptx::mbar_arrive(&bempty[sx]);
            if (j == 1) PP_T(t, 33 + 2 * nb);
          }
        }
      }
    }
  } else if (warp == PP_DRAIN_WARP) {
    // ===================== mean_c drain warp: TMEM lanes 0..31 = the tile's 32 cells, 64 columns per round =====================
    uint32_t sc = 0;
    for (int t = 0; t < n_my; ++t) {
      const int cell = (tile_begin + t) * (UG_BM / 4) + lane;
      const bool live = cell * 4 < M;
      VML_DBG_ASSERT(!live || cell < *n_cells);
      for (int nb = 0; nb < NB; ++nb) {
#pragma unroll 1
        for (int j = 0; j < 2; ++j, ++sc) {
          ptx::mbar_wait(side_full, sc & 1);
          ptx::tc_fence_after();
          uint4* dst = reinterpret_cast<uint4*>(side + (size_t)cell * ld_side + nb * 128 + j * 64);
#pragma unroll 1
          for (int q = 0; q < 2; ++q) {
            float v[32];
            ptx::tmem_ld32(tmem_base + PP_TM_SIDE + (uint32_t)(32 * q), v);
            ptx::tmem_ld_wait();
            if (q == 1) {                                        // accumulator drained: the next box's MMAs may overwrite it
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(side_empty);
            }
            if (live) {
#pragma unroll
              for (int e = 0; e < 4; ++e) dst[4 * q + e] = pp_pack8(v + 8 * e);
            }
          }
        }
      }
    }
  } else if (warp >= 2 && warp < 2 + PP_ROW_WARPS) {
    // ===================== row groups: thread pair == tile row == (cell, clip) =====================
    const int gidx = (warp - 2) >> 3;                     // group 0 takes even local tiles, group 1 odd ones
    const int wg = (warp - 2) & 7;
    const int quad = warp % 4;                            // TMEM lane quadrant this warp may read
    const int grp = wg >> 2;                              // column half of a row this thread owns
    const int r = quad * 32 + lane;                       // row within the tile == TMEM lane
    const int at = wg * 32 + lane;                        // 0..255 within the group
    const bool issuer = at == 0;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t row_off = (uint32_t)((r & 7) * 16 + (r >> 3) * 2048);        // Cs: own row, chunk 0
    const uint32_t prow_off = (uint32_t)((r & 7) * 16 + (r >> 3) * (KG * 128)); // Ps: own row, chunk 0
    const float inv_sqrt_dl = 1.0f / sqrtf((float)PP_DL);
    unsigned char* Cs = Cs0 + gidx * PP_CS_BYTES;
    float4* s_gg = reinterpret_cast<float4*>(Ps);         // partial Grams of a row's two threads (Ps is dead by then)
    auto group_bar = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + gidx), "n"(PP_GROUP_THREADS) : "memory"); };
    if (gidx == 0) {                                      // constants; group 1 first looks at them after cc_ready(0)
      if (at < PP_DL) s_b1[at] = bias1[at];
      {                                                   // I_16, K-major core-matrix layout (lbo 128, sbo 256): element (n, k) = (n == k)
        const int n = at >> 4, k = at & 15;
        *reinterpret_cast<uint16_t*>(I16 + (n & 7) * 16 + (n >> 3) * 256 + (k >> 3) * 128 + (k & 7) * 2) = n == k ? (uint16_t)0x3F80 : (uint16_t)0;
      }
      for (int e = at; e < PP_M4_BYTES / 4; e += PP_GROUP_THREADS) reinterpret_cast<uint32_t*>(M4)[e] = 0u;
      group_bar();
      if (at < 16) {                                      // P[m', k] = 0.25 for k / 4 == m' (m' = 0..3), row index m' + 124
        const int mp = at >> 2, k = at;
        *reinterpret_cast<uint16_t*>(M4 + (mp + 124) * 16 + (k >> 3) * (PP_M4_ROWS * 16) + (k & 7) * 2) = (uint16_t)0x3E80;
      }
      ptx::fence_proxy_async();                           // I_16, M4 -> visible to the MMAs (ordered before this group's first cc_ready arrive)
    }
    auto groups_of = [&](int tile, int& first) {          // sample groups a tile's rows span
      const int mm0 = tile * UG_BM;
      first = code[tile * (UG_BM / 4)] >> 16;
      const int last = code[(min(mm0 + UG_BM, M) - 1) >> 2] >> 16;
      return (last - first) / GS + 1;
    };
    uint32_t sa_count = 0;                                // completed phases of sfull_bar / afull_bar (both advance once per sample group)
    for (int t = gidx; t < n_my; t += 2) {
      const int tile = tile_begin + t;
      const int m0 = tile * UG_BM;
      const int row = m0 + r;
      const bool valid = row < M;
      const int b = valid ? (code[row >> 2] >> 16) : -1;
      VML_DBG_ASSERT(!valid || (b >= 0 && b < B));
      int b_first;
      const int ngroups = groups_of(tile, b_first);
      int staged_bg = -1;                                 // sample group whose query operands the previous tile left in Ks / Wt
      if (issuer) PP_T(t, 0);
      if (t > 0) {
        int pf;
        const int png = groups_of(tile - 1, pf);
        staged_bg = pf + GS * (png - 1);
        sa_count += (uint32_t)png;                        // the other group's tile in between
        // ---- this group's turn: the other group's attention half (U, S, A) is over ----
        ptx::mbar_wait(cc_ready, (uint32_t)(t - 1) & 1);
        ptx::tc_fence_after();
      }
      if (issuer) PP_T(t, 1);
      for (int g = 0; g < ngroups; ++g) {
        const int bg = b_first + GS * g;
        const int sl = b - bg;                              // 0 .. GS-1: this row's sample is in the group
        const bool mine = valid && sl >= 0 && sl < GS;
        group_bar();                                        // this group's previous users of U are done
        if (bg != staged_bg) {
          for (int e = at; e < NW * (PP_DL / 8); e += PP_GROUP_THREADS) {
            const int w = e / (PP_DL / 8), ch = e % (PP_DL / 8);   // word slot, 8-feature chunk
            const int s2 = w / NQP, k = w % NQP;
            const int bb = min(bg + s2, B - 1);
            float kt[8], wh[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) { kt[q] = 0.f; wh[q] = 0.f; }
            if (k < Nq) {
              const float* src = qproj + ((size_t)bb * Nq + k) * ld;
              const float4 k0 = __ldg(reinterpret_cast<const float4*>(src + off_ktil + ch * 8));
              const float4 k1 = __ldg(reinterpret_cast<const float4*>(src + off_ktil + ch * 8 + 4));
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(src + off_what + ch * 8));
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(src + off_what + ch * 8 + 4));
              kt[0] = k0.x; kt[1] = k0.y; kt[2] = k0.z; kt[3] = k0.w; kt[4] = k1.x; kt[5] = k1.y; kt[6] = k1.z; kt[7] = k1.w;
              wh[0] = w0.x; wh[1] = w0.y; wh[2] = w0.z; wh[3] = w0.w; wh[4] = w1.x; wh[5] = w1.y; wh[6] = w1.z; wh[7] = w1.w;
            } else if (k == Nq) {                             // spare slot: s_hat, taken with probability 1
              const float* sh = s_hat + (size_t)bb * s_ld + ch * 8;
#pragma unroll
              for (int q = 0; q < 8; ++q) wh[q] = sh[q];
            }
            *reinterpret_cast<uint4*>(Ks + (w & 7) * 16 + (w >> 3) * 2048 + ch * 128) = pp_pack8(kt);
            *reinterpret_cast<uint4*>(Wt + (w & 7) * 16 + ch * (KG * 128) + (w >> 3) * 128) = pp_pack8(wh);
          }
          if (at < NW) {
            const int s2 = at / NQP, k = at % NQP;
            const int bb = min(bg + s2, B - 1);
            s_beta[at] = k < Nq ? qproj[((size_t)bb * Nq + k) * ld + off_beta] : 0.f;
            s_mask[at] = (k < Nq && qmask[(size_t)bb * Nq + k]) ? 1.f : 0.f;
          }
        }
        staged_bg = bg;
        if (g == 0) {
          // ---- c_hat row out of TMEM: + bias, round to bf16 (what the unfused path stores), park in Cs ----
          ptx::mbar_wait(chat_full, (uint32_t)t & 1);
          ptx::tc_fence_after();
          if (issuer) PP_T(t, 2);
          const uint32_t t_addr = tmem_base + lane_base;
#pragma unroll 1
          for (int c = grp * PP_COLS; c < grp * PP_COLS + PP_COLS; c += 32) {
            float v[32];
            ptx::tmem_ld32(t_addr + (uint32_t)c, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; e += 8) {
              float tt[8];
#pragma unroll
              for (int q = 0; q < 8; q += 2) {
                tt[q] = v[e + q]; tt[q + 1] = v[e + q + 1];
                ptx::add2(tt[q], tt[q + 1], s_b1[c + e + q], s_b1[c + e + q + 1]);
              }
              *reinterpret_cast<uint4*>(Cs + row_off + ((c + e) >> 3) * 128) = pp_pack8(tt);
            }
          }
          ptx::tc_fence_before();                              // accumulator read: the next tile's first contraction may overwrite it
          ptx::mbar_arrive(chat_free);
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        group_bar();
        if (issuer) {                                          // ---- (2) S = c_hat . ktil^T ----
          ptx::tc_fence_after();
          constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(UG_BM, NW);
          const uint32_t a0 = ptx::smem_u32(Cs), b0 = ptx::smem_u32(Ks);
#pragma unroll
          for (int k = 0; k < PP_DL / 16; ++k)
            ptx::umma_bf16(tmem_base + PP_TM_S, ptx::umma_desc_nosw(a0 + k * 256, 128, 2048),
                           ptx::umma_desc_nosw(b0 + k * 256, 128, 2048), idesc_s, k != 0);
          ptx::umma_commit(sfull_bar);
        }
        if (grp == 0) {
          ptx::mbar_wait(sfull_bar, sa_count & 1);
          ptx::tc_fence_after();
          // ---- masked softmax over this row's words (models.py:211-220), P row -> Ps ----------------------
          float sv[NW];
#pragma unroll
          for (int c = 0; c < NW; c += 16) ptx::tmem_ld16(tmem_base + lane_base + PP_TM_S + (uint32_t)c, sv + c);
          ptx::tmem_ld_wait();
          float p[NQP];
          float mx = -INFINITY;
#pragma unroll
          for (int k = 0; k < NQP; ++k) {
            const float raw = (GS == 2 && sl == 1) ? sv[(GS - 1) * NQP + k] : sv[k];
            const int slot = ((GS == 2 && sl == 1) ? NQP : 0) + k;
            const float mk = mine ? s_mask[slot] : 0.f;
            float s = (raw + (mine ? s_beta[slot] : 0.f)) * inv_sqrt_dl;
            s = s * mk;
            if (mk == 0.f) s = -1e9f;
            p[k] = s;
            if (k < Nq) mx = fmaxf(mx, s);
          }
          float den = 0.f;
#pragma unroll
          for (int k = 0; k < NQP; ++k) {
            const float ex = k < Nq ? __expf(p[k] - mx) : 0.f;
            p[k] = ex; den += ex;
          }
          const float inv_den = mine ? __fdividef(1.0f, den) : 0.f;
#pragma unroll
          for (int k = 0; k < NQP; ++k) p[k] = k == Nq ? (mine ? 1.0f : 0.f) : p[k] * inv_den;
          const uint4 zero4 = make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int kc = 0; kc < NQP / 8; ++kc) {
            const uint4 pk = pp_pack8(p + kc * 8);
            if (GS == 2) {
              *reinterpret_cast<uint4*>(Ps + prow_off + kc * 128) = sl == 0 ? pk : zero4;
              *reinterpret_cast<uint4*>(Ps + prow_off + (NQP / 8 + kc) * 128) = sl == 1 ? pk : zero4;
            } else {
              *reinterpret_cast<uint4*>(Ps + prow_off + kc * 128) = pk;      // rows of other samples carry zeros (inv_den = 0)
            }
          }
          ptx::fence_proxy_async();
        }
        ptx::tc_fence_before();
        group_bar();
        if (issuer) {                                          // ---- (3) A (+)= P . [w_hat ; s_hat] ----
          ptx::tc_fence_after();
          constexpr uint32_t idesc_a = ptx::umma_idesc_bf16_bmn(UG_BM, PP_DL);
          const uint32_t a0 = ptx::smem_u32(Ps), b0 = ptx::smem_u32(Wt);
#pragma unroll
          for (int k = 0; k < NW / 16; ++k)
            ptx::umma_bf16(tmem_base + PP_TM_A, ptx::umma_desc_nosw(a0 + k * 256, 128, KG * 128),
                           ptx::umma_desc_nosw(b0 + k * 256, 128, KG * 128), idesc_a, (g | k) != 0);
          ptx::umma_commit(afull_bar);
        }
        ptx::mbar_wait(afull_bar, sa_count & 1);
        ++sa_count;
        ptx::tc_fence_after();
      }
      // ---- gate G = c_hat * (A + s_hat), Gram of the cell's 4 clips (adjacent lanes); this thread: 64 columns ----
      float gg[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = grp * PP_COLS; c < grp * PP_COLS + PP_COLS; c += 32) {
        float a[32];
        ptx::tmem_ld32(tmem_base + lane_base + PP_TM_A + (uint32_t)c, a);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          const f8 ch = unpack8(*reinterpret_cast<const uint4*>(Cs + row_off + ((c + e) >> 3) * 128));
#pragma unroll
          for (int q = 0; q < 8; q += 2) {
            ptx::mul2(a[e + q], a[e + q + 1], ch.v[q], ch.v[q + 1]);
            if (!valid) { a[e + q] = 0.f; a[e + q + 1] = 0.f; }
          }
        }
        float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
        {
          const bool h1 = (lane & 1) != 0, h2 = (lane & 2) != 0;
#pragma unroll
          for (int e = 0; e < 32; ++e) g0 = fmaf(a[e], a[e], g0);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float mine1 = h1 ? a[16 + i] : a[i], send1 = h1 ? a[i] : a[16 + i];
            const float mine2 = h2 ? a[16 + i] : a[i], send2 = h2 ? a[i] : a[16 + i];
            g1 = fmaf(mine1, __shfl_xor_sync(0xffffffffu, send1, 1), g1);
            g2 = fmaf(mine2, __shfl_xor_sync(0xffffffffu, send2, 2), g2);
            g3 = fmaf(mine1, __shfl_xor_sync(0xffffffffu, send1, 3), g3);
          }
          g1 += __shfl_xor_sync(0xffffffffu, g1, 1);
          g2 += __shfl_xor_sync(0xffffffffu, g2, 2);
          g3 += __shfl_xor_sync(0xffffffffu, g3, 3);
        }
        gg[0] += g0; gg[1] += g1; gg[2] += g2; gg[3] += g3;
      }
      ptx::tc_fence_before();
      s_gg[grp * 128 + r] = make_float4(gg[0], gg[1], gg[2], gg[3]);
      group_bar();
      {
        // (columns 0..63) + (columns 64..127), the same order in every thread of the row (and in content_unit.cu / content_tc.cu)
        const float4 lo = s_gg[r], hi = s_gg[128 + r];
        gg[0] = lo.x + hi.x; gg[1] = lo.y + hi.y; gg[2] = lo.z + hi.z; gg[3] = lo.w + hi.w;
      }
      // ---- 4x4 clip self-attention (models.py:259-266): softmax over the cell's clips, mix c_hat rows ----
      float am = -INFINITY;
#pragma unroll
      for (int m = 0; m < 4; ++m) { gg[m] = gg[m] * inv_sqrt_dl; am = fmaxf(am, gg[m]); }
      float ad = 0.f;
#pragma unroll
      for (int m = 0; m < 4; ++m) { gg[m] = __expf(gg[m] - am); ad += gg[m]; }
      const float inv_ad = __fdividef(1.0f, ad);
#pragma unroll
      for (int m = 0; m < 4; ++m) gg[m] *= inv_ad;
      uint32_t sib[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) { const int rr = r ^ m; sib[m] = (uint32_t)((rr & 7) * 16 + (rr >> 3) * 2048); }
#pragma unroll 4
      for (int c = grp * PP_COLS; c < grp * PP_COLS + PP_COLS; c += 8) {
        f8 o;
#pragma unroll
        for (int q = 0; q < 8; ++q) o.v[q] = 0.f;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const f8 sv = unpack8(*reinterpret_cast<const uint4*>(Cs + sib[m] + (c >> 3) * 128));
#pragma unroll
          for (int q = 0; q < 8; q += 2) ptx::fma2(o.v[q], o.v[q + 1], gg[m], gg[m], sv.v[q], sv.v[q + 1]);
        }
        if (!valid) {
#pragma unroll
          for (int q = 0; q < 8; ++q) o.v[q] = 0.f;
        }
        __syncwarp();
        *reinterpret_cast<uint4*>(Cs + row_off + (c >> 3) * 128) = pp_pack8(o.v);
      }
      ptx::fence_proxy_async();                              // cc_hat -> visible to the tail MMAs
      ptx::mbar_arrive(cc_ready);                            // ... which may start, and so may the other group's attention half
      if (issuer) PP_T(t, 3);

      // ---- (4) epilogue: the accumulator holds cc_hat.W2^T + X (tensor cores), fbar carries the output bias.  A block is two
      //      64-column halves (= output boxes) on alternating accumulators; this thread: 32 columns of each ---------------
      const bf16* frow = fbar + (size_t)(row >> 2) * D + grp * 32;
      uint4 fq[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) fq[i] = valid ? __ldg(reinterpret_cast<const uint4*>(frow) + i) : make_uint4(0, 0, 0, 0);
      if (t > 0) ptx::mbar_wait(tail_done, (uint32_t)(t - 1) & 1);      // the previous tile's boxes are all through yfull / sready
      if (issuer) PP_T(t, 4);
      for (int nb = 0; nb < NB; ++nb) {
        const uint32_t c = (uint32_t)(t * NB + nb);
        const int bt = box_tail(t, nb);
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          unsigned char* xb = Ring + ((bt + 2 + j) % RB) * PP_BOX;
          ptx::mbar_wait_relaxed(&yfull[j], c & 1);              // TMEM data: ordered by the tcgen05 fence below
          ptx::tc_fence_after();
          if (issuer && j == 0) PP_T(t, 5 + 2 * nb);
          float acc[32];
          ptx::tmem_ld32(tmem_base + lane_base + PP_TM_Y + (uint32_t)(64 * j + 32 * grp), acc);
          ptx::tmem_ld_wait();
          ptx::tc_fence_before();                                // accumulator drained: the next block's half j may overwrite it
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&yempty[j]);
          if (!valid) {                // rows past the live count feed the mean_c MMA (0 x NaN = NaN): keep them finite
#pragma unroll
            for (int e = 0; e < 32; ++e) acc[e] = 0.f;
          }
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            const f8 fv = unpack8(fq[pc]);
            float* a = acc + pc * 8;
#pragma unroll
            for (int q = 0; q < 8; q += 2) ptx::add2(a[q], a[q + 1], fv.v[q], fv.v[q + 1]);
            *reinterpret_cast<uint4*>(xb + ptx::sw128_off(r, 4 * grp + pc)) = pp_pack8(a);      // result in place of the residual
          }
          const int nxt = 2 * nb + j + 1;                        // the next box's fbar
          if (nxt < 2 * NB) {
#pragma unroll
            for (int i = 0; i < 4; ++i) fq[i] = valid ? __ldg(reinterpret_cast<const uint4*>(frow + nxt * 64) + i) : make_uint4(0, 0, 0, 0);
          }
          ptx::fence_proxy_async();                              // shared-memory writes -> visible to the TMA store / mean_c MMAs
          ptx::mbar_arrive(&sready[j]);                          // the store warp takes it from here
          if (issuer && j == 1) PP_T(t, 6 + 2 * nb);
        }
      }
      ptx::mbar_arrive(tail_done);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tmem_base); }
}

template <int NQP, int GS>
static int launch_content_unit_pp(const CUtensorMap* tm, int grid, const bf16* fbar, bf16* side, int ld_side, const float* b1,
                                  const float* qproj, int ld, int off_what, int off_ktil, int off_beta, const float* s_hat,
                                  int s_ld, const uint8_t* qmask, vml_cells_t cells, int B, vml_dims_t d, int store_cu,
                                  cudaStream_t st) {
  using Cfg = PpCfg<NQP, GS>;
  static_assert(Cfg::SMEM <= 232448, "content_unit_pp_kernel exceeds the 227 KB shared-memory limit");
  VML_CUDA(ensure_dyn_smem((const void*)(content_unit_pp_kernel<NQP, GS>), (size_t)(Cfg::SMEM)));
  content_unit_pp_kernel<NQP, GS><<<grid, PP_THREADS, Cfg::SMEM, st>>>(tm[0], tm[1], tm[2], tm[3], d.D, fbar, side, ld_side, b1, qproj,
                                                                       ld, off_what, off_ktil, off_beta, s_hat, s_ld, qmask, cells.code,
                                                                       cells.n_cells, d.Nq, B, store_cu);
  VML_LAUNCHED(1);
  return VML_OK;
}

// Same contract as content_unit() with bias_in_fbar (content_unit.cu); called from there for variant 5.
int content_unit_pp(const CUtensorMap* tm, int grid, const void* fbar, void* side, int ld_side, const float* b1, const float* qproj,
                    int ld, int off_what, int off_ktil, int off_beta, const float* s_hat, int s_ld, const uint8_t* qmask,
                    vml_cells_t cells, int B, vml_dims_t d, int store_cu, cudaStream_t st) {
  static bool reg = (register_kernel("content_unit_pp_kernel"), true); (void)reg;
#define VML_PP(NQP, GS) return launch_content_unit_pp<NQP, GS>(tm, grid, (const bf16*)fbar, (bf16*)side, ld_side, b1, qproj, ld, off_what, off_ktil, off_beta, s_hat, s_ld, qmask, cells, B, d, store_cu, st)
  if (d.Nq + 1 <= 8) VML_PP(8, 2);
  if (d.Nq + 1 <= 16) VML_PP(16, 2);
  VML_PP(32, 1);
#undef VML_PP
}

}  // namespace vml

#ifdef VML_CU_TIMING
extern "C" __attribute__((visibility("default"))) int vml_debug_pp_timing(long long* host, int n) { return vml::pp_debug_read(host, n); }
#endif
