// a5+a6: the WHOLE content unit of one SMI layer as ONE persistent tcgen05 kernel (bf16 fast mode,
// dl = 128, C = 4, D <= 512):
//
//   c_hat = fc.W_c_hat^T + b -> content-word attention -> gate -> CxC clip self-attention -> cc_hat
//   cu    = cc_hat.W_c^T + b_c + fc + fbar,      side = mean_c cu  (the moment unit's operand half)
//   reference: ContentUnit.forward models.py:242-276, ContentAttention.forward models.py:207-226, :297
//
// The two-kernel version (content_tc.cu + gemm_res.cu) reads the fc tile twice from HBM and round-trips
// cc_hat; here a tile's 128 (cell, clip) rows of fc stay RESIDENT in shared memory (D/64 TMA boxes of
// 128 x 64, 128B-swizzled = 128 KB at D = 512) from the first contraction to the residual add, so the
// layer moves fc exactly once in and once out:
//   (1) main    c_hat[128 x 128]  = X[128 x D] . W1^T        A = the resident boxes, B = W1 K-blocks (ring)
//   (2) scores  S[128 x NW]       = c_hat_bf16 . ktil^T      as content_tc.cu
//   (3) attend  A[128 x 128]      = P[128 x NW] . [w_hat ; s_hat]
//   (4) tail    Y_nb[128 x 128]   = cc_hat[128 x 128] . W2[nb]^T   for the D/128 column blocks; A = cc_hat
//               written in place of c_hat (un-swizzled core-matrix layout), B = W2 boxes through the ring
//   epilogue(4) out = Y + b_c + X + fbar  written IN PLACE into the resident boxes and TMA-stored; fbar is
//               prefetched into registers, mean_c(out) (4 adjacent lanes) goes straight to global memory.
// The next tile's boxes are re-loaded as soon as the store of their column block has left shared memory,
// so its HBM reads and its main loop overlap the tail epilogue of the current tile.
//
// Warp roles (608 threads): warp 0 TMA producer, warp 1 MMA issuer for (1) and (4), warps 2..17 row warps: four
// threads (same TMEM lane, warps w, w + 4, w + 8, w + 12) == tile row == (cell, clip), each owning 32 of a row's 128
// columns in every per-row loop; thread 64 issues (2) and (3); the last warp issues the TMA stores and waits for
// them.  The row warps bound the kernel (measured per tile at the bench shape: 24.5 us with 4 row warps, 17.5 us
// with 8, 15.3 us with 16 at 96 registers per thread), the tensor cores and HBM are far from their roofs.
//
// TMEM columns: [0,128) c_hat accumulator; [128,256) A, later Y (even column blocks); [256,256+NW) S,
// [256,384) later Y (odd column blocks) -- the aliases are phase-exclusive (Y MMAs are issued only after
// every attention thread has finished reading S and A; the next tile's S/A only after Y was drained).
#include <stdlib.h>

#include "common.cuh"
#include "gemm_umma.cuh"
#include "sm100.cuh"

namespace vml {

#ifdef VML_CU_TIMING
__device__ long long g_cu_dbg[512];
__device__ __forceinline__ long long cu_now() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define CU_T(slot) do { if (issuer && blockIdx.x == 0 && it < 4) g_cu_dbg[it * 48 + (slot)] = cu_now(); } while (0)
#define CU_TX(slot) do { if (blockIdx.x == 0 && it < 4) g_cu_dbg[it * 48 + (slot)] = cu_now(); } while (0)
int cu_debug_read(long long* host, int n) { return (int)cudaMemcpyFromSymbol(host, g_cu_dbg, sizeof(long long) * n); }
#else
#define CU_T(slot) do { } while (0)
#define CU_TX(slot) do { } while (0)
#endif

constexpr int CU_DL = 128, CU_WST = 2, CU_MAXKB = 8;
#ifndef VML_CU_ROW_WARPS
#define VML_CU_ROW_WARPS 16
#endif
constexpr int CU_ROW_WARPS = VML_CU_ROW_WARPS, CU_ROW_THREADS = 32 * CU_ROW_WARPS, CU_THREADS = 64 + CU_ROW_THREADS + 32;   // + store warp
// V2: + an unused filler warp + the mean_c drain warp (a warp reads TMEM lanes 32*(warp % 4)..+31: the drain needs lanes 0..31)
constexpr int CU_DRAIN_WARP = (2 + CU_ROW_WARPS + 1 + 3) / 4 * 4, CU_THREADS_V2 = 32 * (CU_DRAIN_WARP + 1);
constexpr int CU_TMEM_SIDE = 384;                    // V2: mean_c accumulator, lanes 0..31 x 128 columns
constexpr int CU_M4_ROWS = 252, CU_M4_BYTES = 2 * CU_M4_ROWS * 16;   // V2: master pattern of the mean_c operand (see kernel)
constexpr int CU_SPLIT = CU_ROW_WARPS / 4;           // threads per tile row (one per TMEM-lane-sharing warp)
constexpr int CU_COLS = 128 / CU_SPLIT;              // columns of a 128-column block each of them owns
static_assert(CU_ROW_WARPS == 8 || CU_ROW_WARPS == 16, "2 or 4 row warps per TMEM lane quadrant");
#define CU_ROW_BAR() asm volatile("bar.sync 1, %0;" ::"n"(CU_ROW_THREADS) : "memory")
constexpr int CU_BOX = UG_BM * UG_BK * 2;            // one 128 x 64 bf16 box (16 KB)
constexpr int CU_CS_BYTES = UG_BM * CU_DL * 2;       // c_hat / cc_hat tile
constexpr int CU_TMEM_A = 128, CU_TMEM_S = 256, CU_TMEM_Y0 = 128, CU_TMEM_Y1 = 256;
constexpr int CU_BAR_BYTES = 512;                    // mbarriers + the TMEM slot

// V = 1: round-1 epilogue (bias, residual and fbar added in registers, mean_c by a shuffle reduce-scatter).
// V = 4: V3's epilogue, but NO resident tile: the 128 KB it occupied become a ring of ~10 boxes through which the tile's X and
//        W1 blocks stream for the first contraction and -- per output column block -- the W2 boxes plus the SAME X boxes
//        again (an L2 hit: the SM read them microseconds ago) for the residual MMA; the finished block is written in place of
//        those re-read boxes and stored from there.  With only two 16 KB weight stages in flight (all that fit beside a
//        resident tile) the kernel was bound by the weight stream: 16 box round trips of ~1 us per tile, serialised with the
//        row warps' work.  The deeper ring also lets tile t+1's first contraction run under tile t's attention phase.
// V = 3: V2 with mean_c on the tensor cores as well (side = M4 . out, out read back as an MN-major operand straight from the
//        finished tile; a dedicated warp drains the 32 x 128 result) -- the row warps only add fbar, round and store.
// V = 2: the output bias arrives folded into fbar (the boundary unit adds it before rounding), the residual X is added by
//        the tensor cores (Y += X_box . I_16, one M=128,N=16,K=16 MMA per 16 columns: exact products, fp32 accumulation) and
//        mean_c is read back from the finished bf16 tile in shared memory by one lane per (cell, 8-column piece) -- no
//        shuffles, ~2.3x fewer instructions per output element in the phase that bounds the kernel.
template <int NQP, int GS, int V = 3>
struct CuCfg {
  static constexpr int NW = GS * NQP;                // word slots of a GS-sample group
  static constexpr int KG = NW / 8;
  static constexpr int KS_BYTES = NW * CU_DL * 2;    // keys,   K-major (rows = word slots), lbo 128, sbo 2048
  static constexpr int WT_BYTES = NW * CU_DL * 2;    // values, MN-major, lbo 128, sbo KG*128
  static constexpr int PS_BYTES = UG_BM * NW * 2;    // probabilities, K-major, lbo 128, sbo KG*128
  static constexpr int GG_BYTES = CU_SPLIT * UG_BM * 16;   // partial Grams (float4 per row thread), aliased onto Ps
  static constexpr int U_RAW = KS_BYTES + WT_BYTES + (PS_BYTES > GG_BYTES ? PS_BYTES : GG_BYTES);
  static constexpr int U_BYTES = (U_RAW + 1023) / 1024 * 1024;
  static constexpr int SIDE_FLOATS = 2 * NW + CU_DL + (V == 1 ? CU_MAXKB * 64 : 128 + (V >= 3 ? CU_M4_BYTES / 4 : 0));   // beta | mask | b1 | b2 (V1) or I_16 + M4 (V2)
  static constexpr int FIXED = CU_CS_BYTES + U_BYTES + SIDE_FLOATS * 4 + 1024 + CU_BAR_BYTES;
  // V4: no resident tile -- ONE ring of 16 KB boxes (as many as fit) carries X, W1, W2 and the re-read X of the residual
  static constexpr int RB = (232448 - FIXED) / CU_BOX;
  static constexpr int X_BYTES = V == 4 ? RB * CU_BOX : CU_MAXKB * CU_BOX;
  static constexpr int SMEM = X_BYTES + FIXED + (V == 4 ? 0 : CU_WST * CU_BOX);
};

__device__ __forceinline__ uint4 cu_pack8(const float* v) {
  uint4 u; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return u;
}

template <int NQP, int GS, int V>
__global__ void __launch_bounds__(V >= 3 ? CU_THREADS_V2 : CU_THREADS, 1)
content_unit_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                    const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut, int D,
                    const bf16* __restrict__ fbar, bf16* __restrict__ side, int ld_side,
                    const float* __restrict__ bias1, const float* __restrict__ bias2, const float* __restrict__ qproj, int ld,
                    int off_what, int off_ktil, int off_beta, const float* __restrict__ s_hat, int s_ld,
                    const uint8_t* __restrict__ qmask, const int32_t* __restrict__ code, const int32_t* __restrict__ n_cells,
                    int Nq, int B, int store_cu) {
  using Cfg = CuCfg<NQP, GS, V>;
  constexpr int NW = Cfg::NW, KG = Cfg::KG;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array: keeps the shared address space visible to the
  // compiler (LDS/STS instead of generic LD/ST, which an integer round-trip of the pointer would force)
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* Xs = smem;                                    // resident fc tile: KB boxes
  unsigned char* Cs = Xs + Cfg::X_BYTES;                       // c_hat, then cc_hat
  unsigned char* U = Cs + CU_CS_BYTES;                         // Ks | Wt | Ps: the attention operands
  unsigned char* Ks = U;
  unsigned char* Wt = Ks + Cfg::KS_BYTES;
  unsigned char* Ps = Wt + Cfg::WT_BYTES;
  unsigned char* Wr = U + Cfg::U_BYTES;                        // weight ring
  float* s_beta = reinterpret_cast<float*>(Wr + (V == 4 ? 0 : CU_WST * CU_BOX));   // [NW]  (V4 has no separate weight ring)
  float* s_mask = s_beta + NW;                                      // [NW]
  float* s_b1 = s_mask + NW;                                        // [128]
  float* s_b2 = s_b1 + CU_DL;                                       // V1: [D] output bias;  V2: I_16 (bf16 16x16 identity, 512 B)
  unsigned char* I16 = reinterpret_cast<unsigned char*>(s_b2);
  // V2: mean over a cell's 4 clips on the tensor cores, side[c, n] = sum_r M4[c, r] out[r, n] with M4[c, r] = 0.25 [r / 4 == c].
  // Per K step s (tile rows 16s..16s+15) the A operand is A_s[m, k] = 0.25 [m == 4s + k / 4] = P[m - 4s, k]: ONE master
  // pattern P[m', k] = 0.25 [m' == k / 4], m' in [-124, 127], stored K-major with its rows 16 bytes apart (sbo 128) and its two
  // 8-column groups CU_M4_ROWS * 16 bytes apart; the descriptor of step s just starts 4s rows earlier.
  unsigned char* M4 = I16 + 512;
  uint64_t* xfull = reinterpret_cast<uint64_t*>(s_b2 + (V == 1 ? CU_MAXKB * 64 : 128 + (V >= 3 ? CU_M4_BYTES / 4 : 0)));
  uint64_t* xfree = xfull + CU_MAXKB;      // [4] per column block
  uint64_t* wfull = xfree + 4;
  uint64_t* wempty = wfull + CU_WST;
  uint64_t* chat_full = wempty + CU_WST;
  uint64_t* cc_ready = chat_full + 1;
  uint64_t* sfull_bar = cc_ready + 1;
  uint64_t* afull_bar = sfull_bar + 1;
  uint64_t* yfull = afull_bar + 1;         // [2]
  uint64_t* yempty = yfull + 2;            // [2]
  uint64_t* sready = yempty + 2;           // [2] column block finished in shared memory -> store warp
  uint64_t* cs_free = sready + 2;          // V2: the last column block's output (staged in Cs) has been stored
  uint64_t* side_full = cs_free + 1;       // V2: mean_c MMAs of a column block have completed
  uint64_t* side_empty = side_full + 1;    // V2: the drain warp has read the mean_c accumulator
  uint64_t* chat_free = side_empty + 1;    // V4: the row warps have parked c_hat -> the next tile's first contraction may start
  uint64_t* bfull = chat_free + 1;         // V4: [RB] box landed
  uint64_t* bempty = bfull + 12;           // V4: [RB] box consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bempty + 12);
  static_assert(Cfg::RB <= 12 && (24 + 12 + 12 + 12) * 8 + 8 <= CU_BAR_BYTES, "barrier block");
  unsigned char* Ring = smem;              // V4: RB boxes (V < 4: this is Xs)
  constexpr int RB = Cfg::RB;

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int M = *n_cells * 4;
  const int num_tiles = (M + UG_BM - 1) / UG_BM;
  const int KB = D / UG_BK, NB = D / 128;
  // a CTA walks a contiguous range of tiles: consecutive tiles mostly belong to the same sample(s), whose query
  // operands then stay staged in shared memory
  const int tiles_per_cta = (num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile_begin = min((int)blockIdx.x * tiles_per_cta, num_tiles), tile_end = min(tile_begin + tiles_per_cta, num_tiles);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmX); ptx::prefetch_tensormap(&tmW1); ptx::prefetch_tensormap(&tmW2);
    ptx::prefetch_tensormap(&tmOut);
    for (int i = 0; i < CU_MAXKB; ++i) ptx::mbar_init(&xfull[i], 1);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&xfree[i], 1);
    for (int i = 0; i < CU_WST; ++i) { ptx::mbar_init(&wfull[i], 1); ptx::mbar_init(&wempty[i], 1); }
    ptx::mbar_init(chat_full, 1);
    ptx::mbar_init(cc_ready, CU_ROW_THREADS);
    ptx::mbar_init(sfull_bar, 1);
    ptx::mbar_init(afull_bar, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&yfull[i], 1); ptx::mbar_init(&yempty[i], CU_ROW_WARPS);
      ptx::mbar_init(&sready[i], CU_ROW_THREADS);
    }
    ptx::mbar_init(cs_free, 1);
    ptx::mbar_init(side_full, 1);
    ptx::mbar_init(side_empty, 1);
    ptx::mbar_init(chat_free, CU_ROW_THREADS);
    for (int i = 0; i < 12; ++i) { ptx::mbar_init(&bfull[i], 1); ptx::mbar_init(&bempty[i], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // V4: program order of the ring's boxes.  PER = 2 * KB boxes per phase (main: X(kb), W1(kb);  tail: per column block
  // W2(nb, 0), W2(nb, 1), X(2nb), X(2nb + 1)); a CTA's sequence is main(0), then for t = 0..n-1: main(t + 1) (if any), tail(t)
  const int n_my = tile_end - tile_begin, PER = 2 * KB;
  auto base_main = [&](int t) { return t == 0 ? 0 : PER + 2 * PER * (t - 1); };
  auto base_tail = [&](int t) { return PER + 2 * PER * t + (t + 1 < n_my ? PER : 0); };

  if (V == 4 && warp == 0) {
    if (lane == 0) {                                    // ===================== TMA producer (V4) =====================
      // L2 policy: the tile's first read and the weights stay (evict_last: the tile is read again ~8 us later by this SM,
      // the weights by every SM all the time); the second read is the last use (evict_first).  Without the hints 55 % of the
      // re-reads missed L2 on the ActivityNet pass (ncu: 2.38 GB read against 1.64 GB algorithmic).
      const uint64_t keep = ptx::l2_policy_evict_last(), once = ptx::l2_policy_evict_first();
      auto load = [&](int bi, const CUtensorMap* tm, int c0, int c1, uint64_t policy) {
        const int slot = bi % RB;
        VML_DBG_ASSERT(bi >= 0 && slot < RB && (slot + 1) * CU_BOX <= Cfg::X_BYTES && c1 >= 0);
        ptx::mbar_wait(&bempty[slot], ((bi / RB) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&bfull[slot], CU_BOX);
        ptx::tma_load_2d_hint(Ring + slot * CU_BOX, tm, &bfull[slot], c0, c1, policy);
      };
      auto do_main = [&](int t) {
        const int base = base_main(t), m0 = (tile_begin + t) * UG_BM;
        for (int kb = 0; kb < KB; ++kb) { load(base + 2 * kb, &tmX, kb * UG_BK, m0, keep); load(base + 2 * kb + 1, &tmW1, kb * UG_BK, 0, keep); }
      };
      auto do_tail = [&](int t) {
        const int base = base_tail(t), m0 = (tile_begin + t) * UG_BM;
        for (int nb = 0; nb < NB; ++nb) {
          load(base + 4 * nb, &tmW2, 0, nb * 128, keep);
          load(base + 4 * nb + 1, &tmW2, UG_BK, nb * 128, keep);
          load(base + 4 * nb + 2, &tmX, (2 * nb) * UG_BK, m0, once);
          load(base + 4 * nb + 3, &tmX, (2 * nb + 1) * UG_BK, m0, once);
        }
      };
      if (n_my > 0) do_main(0);
      for (int t = 0; t < n_my; ++t) {
        if (t + 1 < n_my) do_main(t + 1);
        do_tail(t);
      }
    }
  } else if (V == 4 && warp == 1) {
    if (lane == 0) {                                    // ===================== MMA issuer (V4): (1) and (4) =====================
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(UG_BM, CU_DL);
      constexpr uint32_t idesc_r = ptx::umma_idesc_bf16(UG_BM, 16);
      const uint64_t idn = ptx::umma_desc_nosw(ptx::smem_u32(I16), 128, 256);
      auto wait_box = [&](int bi) {
        const int slot = bi % RB;
        ptx::mbar_wait(&bfull[slot], (bi / RB) & 1);
        return slot;
      };
      uint32_t yi = 0;
      auto do_main = [&](int t) {
        if (t > 0) ptx::mbar_wait(chat_free, (uint32_t)(t - 1) & 1);     // c_hat(t - 1) has left the accumulator
        ptx::tc_fence_after();
        const int base = base_main(t);
        for (int kb = 0; kb < KB; ++kb) {
          const int sx = wait_box(base + 2 * kb), sw = wait_box(base + 2 * kb + 1);
          ptx::tc_fence_after();
          const uint64_t adesc = ptx::umma_desc_sw128(ptx::smem_u32(Ring + sx * CU_BOX));
          const uint64_t bdesc = ptx::umma_desc_sw128(ptx::smem_u32(Ring + sw * CU_BOX));
#pragma unroll
          for (int k = 0; k < UG_BK / 16; ++k)
            ptx::umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          ptx::umma_commit(&bempty[sx]);
          ptx::umma_commit(&bempty[sw]);
        }
        ptx::umma_commit(chat_full);
      };
      auto do_tail = [&](int t) {
        ptx::mbar_wait(cc_ready, (uint32_t)t & 1);       // cc_hat is in Cs; S and A have been read out of TMEM
        ptx::tc_fence_after();
        const uint32_t c0 = ptx::smem_u32(Cs);
        const int base = base_tail(t);
        for (int nb = 0; nb < NB; ++nb, ++yi) {
          const uint32_t yb = yi & 1;
          ptx::mbar_wait(&yempty[yb], ((yi >> 1) & 1) ^ 1);
          const uint32_t d_tmem = tmem_base + (yb ? CU_TMEM_Y1 : CU_TMEM_Y0);
          VML_DBG_ASSERT((d_tmem & 0xffffu) + 128u <= (tmem_base & 0xffffu) + 512u);      // accumulator inside the 512 allocated columns
          int sw[2], sx[2];
          for (int j = 0; j < 2; ++j) sw[j] = wait_box(base + 4 * nb + j);
          for (int j = 0; j < 2; ++j) sx[j] = wait_box(base + 4 * nb + 2 + j);
          ptx::tc_fence_after();
          for (int kh = 0; kh < 2; ++kh) {
            const uint64_t bdesc = ptx::umma_desc_sw128(ptx::smem_u32(Ring + sw[kh] * CU_BOX));
#pragma unroll
            for (int k = 0; k < UG_BK / 16; ++k)
              ptx::umma_bf16(d_tmem, ptx::umma_desc_nosw(c0 + (uint32_t)((kh * 4 + k) * 256), 128, 2048), bdesc + (uint64_t)(k * 2),
                             idesc, (kh | k) != 0);
            ptx::umma_commit(&bempty[sw[kh]]);
          }
          for (int j = 0; j < 2; ++j) {                  // residual: Y[:, 64j + 16k ..+16) += X_box(2nb + j)[:, 16k ..+16) . I_16^T
            const uint64_t xdesc = ptx::umma_desc_sw128(ptx::smem_u32(Ring + sx[j] * CU_BOX));
#pragma unroll
            for (int k = 0; k < UG_BK / 16; ++k)
              ptx::umma_bf16(d_tmem + (uint32_t)(64 * j + 16 * k), xdesc + (uint64_t)(k * 2), idn, idesc_r, true);
          }
          ptx::umma_commit(&yfull[yb]);
        }
      };
      if (n_my > 0) do_main(0);
      for (int t = 0; t < n_my; ++t) {
        if (t + 1 < n_my) do_main(t + 1);
        do_tail(t);
      }
    }
  } else if (V == 4 && warp == 2 + CU_ROW_WARPS) {
    if (lane == 0) {                                    // ===================== store warp (V4) =====================
      uint32_t yi = 0;
      for (int t = 0; t < n_my; ++t) {
        const int m0 = (tile_begin + t) * UG_BM, base = base_tail(t);
        for (int nb = 0; nb < NB; ++nb, ++yi) {
          const uint32_t yb = yi & 1;
          ptx::mbar_wait(&sready[yb], (yi >> 1) & 1);
          const int sx0 = (base + 4 * nb + 2) % RB, sx1 = (base + 4 * nb + 3) % RB;
          const unsigned char* b0 = Ring + sx0 * CU_BOX;
          const unsigned char* b1 = Ring + sx1 * CU_BOX;
          if (store_cu) {               // the next reader of cu is a later kernel: do not let it push the tiles being re-read out of L2
            ptx::tma_store_2d_hint(&tmOut, b0, nb * 128, m0, ptx::l2_policy_evict_first());
            ptx::tma_store_2d_hint(&tmOut, b1, nb * 128 + 64, m0, ptx::l2_policy_evict_first());
            ptx::bulk_commit();
          }
          // mean over the cell's 4 clips of the finished block on the tensor cores (see V3)
          ptx::mbar_wait(side_empty, (yi & 1) ^ 1);
          ptx::tc_fence_after();
          constexpr uint32_t idesc_m = ptx::umma_idesc_bf16_bmn(UG_BM, 64);
          const uint32_t m4 = ptx::smem_u32(M4);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t bb = ptx::smem_u32(j ? b1 : b0);
#pragma unroll
            for (int ks = 0; ks < UG_BM / 16; ++ks)
              ptx::umma_bf16(tmem_base + CU_TMEM_SIDE + (uint32_t)(64 * j),
                             ptx::umma_desc_nosw(m4 + (uint32_t)((124 - 4 * ks) * 16), CU_M4_ROWS * 16, 128),
                             ptx::umma_desc_sw128_mn(bb + (uint32_t)(ks * 2048)), idesc_m, ks != 0);
          }
          ptx::umma_commit(side_full);
          if (store_cu) ptx::bulk_wait_read<0>();
          ptx::mbar_wait(side_full, yi & 1);             // the MMAs have read the boxes too
          ptx::mbar_arrive(&bempty[sx0]);
          ptx::mbar_arrive(&bempty[sx1]);
        }
      }
    }
  } else if (warp == 0) {
    if (lane == 0) {                                    // ===================== TMA producer =====================
      int wst = 0; uint32_t wph = 0;
      uint32_t it = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
        const uint32_t ph = it & 1;
        const int m0 = tile * UG_BM;
        for (int kq = 0; kq < KB; ++kq) {
          // V2: the last column block's boxes are released first (by the tensor cores, see the MMA issuer), so the next
          // tile's main loop starts with them; the in-place blocks follow in the order their stores complete
          const int kb = V >= 2 ? (kq + KB - 2) % KB : kq;
          if ((kb & 1) == 0) ptx::mbar_wait(&xfree[kb >> 1], ph ^ 1);      // previous tile's column block has been stored
          ptx::mbar_arrive_expect_tx(&xfull[kb], CU_BOX);
          ptx::tma_load_2d(Xs + kb * CU_BOX, &tmX, &xfull[kb], kb * UG_BK, m0);
          ptx::mbar_wait(&wempty[wst], wph ^ 1);
          ptx::mbar_arrive_expect_tx(&wfull[wst], CU_BOX);
          ptx::tma_load_2d(Wr + wst * CU_BOX, &tmW1, &wfull[wst], kb * UG_BK, 0);
          if (++wst == CU_WST) { wst = 0; wph ^= 1; }
        }
        for (int nb = 0; nb < NB; ++nb) {
          for (int kh = 0; kh < 2; ++kh) {
            ptx::mbar_wait(&wempty[wst], wph ^ 1);
            if (kh == 0) CU_TX(30 + nb);
            ptx::mbar_arrive_expect_tx(&wfull[wst], CU_BOX);
            ptx::tma_load_2d(Wr + wst * CU_BOX, &tmW2, &wfull[wst], kh * UG_BK, nb * 128);
            if (++wst == CU_WST) { wst = 0; wph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                    // ===================== MMA issuer: (1) and (4) =====================
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(UG_BM, CU_DL);
      int wst = 0; uint32_t wph = 0;
      uint32_t yi = 0, it = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
        const uint32_t ph = it & 1;
        for (int kq = 0; kq < KB; ++kq) {
          const int kb = V >= 2 ? (kq + KB - 2) % KB : kq;       // same order as the producer
          ptx::mbar_wait(&xfull[kb], ph);
          ptx::mbar_wait(&wfull[wst], wph);
          ptx::tc_fence_after();
          const uint64_t adesc = ptx::umma_desc_sw128(ptx::smem_u32(Xs + kb * CU_BOX));
          const uint64_t bdesc = ptx::umma_desc_sw128(ptx::smem_u32(Wr + wst * CU_BOX));
#pragma unroll
          for (int k = 0; k < UG_BK / 16; ++k)
            ptx::umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kq | k) != 0);
          ptx::umma_commit(&wempty[wst]);
          if (++wst == CU_WST) { wst = 0; wph ^= 1; }
        }
        ptx::umma_commit(chat_full);
        ptx::mbar_wait(cc_ready, ph);                    // cc_hat is in Cs; S and A have been read out of TMEM
        ptx::tc_fence_after();
        const uint32_t c0 = ptx::smem_u32(Cs);
        for (int nb = 0; nb < NB; ++nb, ++yi) {
          const uint32_t yb = yi & 1;
          ptx::mbar_wait(&yempty[yb], ((yi >> 1) & 1) ^ 1);
          CU_TX(18 + 3 * nb);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + (yb ? CU_TMEM_Y1 : CU_TMEM_Y0);
          for (int kh = 0; kh < 2; ++kh) {
            ptx::mbar_wait(&wfull[wst], wph);
            if (kh == 1) CU_TX(19 + 3 * nb);
            ptx::tc_fence_after();
            const uint64_t bdesc = ptx::umma_desc_sw128(ptx::smem_u32(Wr + wst * CU_BOX));
#pragma unroll
            for (int k = 0; k < UG_BK / 16; ++k)
              ptx::umma_bf16(d_tmem, ptx::umma_desc_nosw(c0 + (uint32_t)((kh * 4 + k) * 256), 128, 2048), bdesc + (uint64_t)(k * 2),
                             idesc, (kh | k) != 0);
            ptx::umma_commit(&wempty[wst]);
            if (++wst == CU_WST) { wst = 0; wph ^= 1; }
          }
          if (V >= 2) {
            // residual on the tensor cores: Y[:, 64j + 16s ..+16) += X_box(2nb + j)[:, 16s ..+16) . I_16^T
            constexpr uint32_t idesc_r = ptx::umma_idesc_bf16(UG_BM, 16);
            const uint64_t idn = ptx::umma_desc_nosw(ptx::smem_u32(I16), 128, 256);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint64_t xdesc = ptx::umma_desc_sw128(ptx::smem_u32(Xs + (2 * nb + j) * CU_BOX));
#pragma unroll
              for (int k = 0; k < UG_BK / 16; ++k)
                ptx::umma_bf16(d_tmem + (uint32_t)(64 * j + 16 * k), xdesc + (uint64_t)(k * 2), idn, idesc_r, true);
            }
            // The LAST column block's output is staged in Cs (dead once these MMAs have read cc_hat), not in place: its X
            // boxes are free for the next tile's loads as soon as the tensor cores have consumed them
            if (nb == NB - 1) ptx::umma_commit(&xfree[nb]);
          }
          ptx::umma_commit(&yfull[yb]);
          CU_TX(20 + 3 * nb);
        }
      }
    }
  } else if (warp == 2 + CU_ROW_WARPS) {
    if (lane == 0) {                                    // ===================== store warp =====================
      // Waiting for a TMA store to finish READING shared memory takes ~1 us; done here, off the row warps' path.
      uint32_t yi = 0;
      // V3: mean over the cell's 4 clips of the finished block: 8 K steps (16 tile rows) x 2 boxes, M = 128 (32 rows used),
      // N = 64, A = the shifted master pattern, B = the box itself read MN-major.  Block counter c: phase of side_full / side_empty.
      auto side_mma = [&](const unsigned char* boxes, uint32_t c) {
        ptx::mbar_wait(side_empty, (c & 1) ^ 1);               // the drain warp has read the previous block's result
        ptx::tc_fence_after();
        constexpr uint32_t idesc_m = ptx::umma_idesc_bf16_bmn(UG_BM, 64);
        const uint32_t m4 = ptx::smem_u32(M4);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t b0 = ptx::smem_u32(boxes + j * CU_BOX);
#pragma unroll
          for (int ks = 0; ks < UG_BM / 16; ++ks)
            ptx::umma_bf16(tmem_base + CU_TMEM_SIDE + (uint32_t)(64 * j),
                           ptx::umma_desc_nosw(m4 + (uint32_t)((124 - 4 * ks) * 16), CU_M4_ROWS * 16, 128),
                           ptx::umma_desc_sw128_mn(b0 + (uint32_t)(ks * 2048)), idesc_m, ks != 0);
        }
        ptx::umma_commit(side_full);
      };
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int m0 = tile * UG_BM;
        for (int nb = 0; nb < NB; ++nb, ++yi) {
          const uint32_t yb = yi & 1;
          ptx::mbar_wait(&sready[yb], (yi >> 1) & 1);
          const bool in_cs = V >= 2 && nb == NB - 1;             // V2: the last block was staged in Cs
          const unsigned char* src = in_cs ? Cs : Xs + (2 * nb) * CU_BOX;
          if (store_cu) {
            ptx::tma_store_2d(&tmOut, src, nb * 128, m0);
            ptx::tma_store_2d(&tmOut, src + CU_BOX, nb * 128 + 64, m0);
            ptx::bulk_commit();
            if (V >= 2) {
              // this warp has nothing else to do until the next block is ready (>= 1 us away): wait for its own store to
              // have read shared memory and hand the boxes back at once (V1 released a block one block late)
              if (V == 3) side_mma(src, yi);
              ptx::bulk_wait_read<0>();
              if (V == 3) ptx::mbar_wait(side_full, yi & 1);     // ... and the mean_c MMAs to have read them too
              ptx::mbar_arrive(in_cs ? cs_free : &xfree[nb]);
            } else {
              if (nb > 0) {                                    // the previous column block has left shared memory
                ptx::bulk_wait_read<1>();
                ptx::mbar_arrive(&xfree[nb - 1]);
              }
              if (nb == NB - 1) {
                ptx::bulk_wait_read<0>();
                ptx::mbar_arrive(&xfree[nb]);
              }
            }
          } else {
            if (V == 3) { side_mma(src, yi); ptx::mbar_wait(side_full, yi & 1); }
            ptx::mbar_arrive(in_cs ? cs_free : &xfree[nb]);    // nothing to store: the boxes are free at once
          }
        }
      }
    }
  } else if (V >= 3 && warp == CU_DRAIN_WARP) {
    // ===================== mean_c drain warp (V3): TMEM lanes 0..31 = the tile's 32 cells, 128 columns per block =====================
    uint32_t c = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int cell = tile * (UG_BM / 4) + lane;
      const bool live = cell * 4 < M;
      VML_DBG_ASSERT(!live || cell < *n_cells);
      for (int nb = 0; nb < NB; ++nb, ++c) {
        ptx::mbar_wait(side_full, c & 1);
        ptx::tc_fence_after();
        uint4* dst = reinterpret_cast<uint4*>(side + (size_t)cell * ld_side + nb * 128);
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {                          // 32 columns at a time (register budget of the whole CTA)
          float v[32];
          ptx::tmem_ld32(tmem_base + CU_TMEM_SIDE + (uint32_t)(32 * q), v);
          ptx::tmem_ld_wait();
          if (q == 3) {                                        // accumulator drained: the next block's MMAs may overwrite it
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(side_empty);
          }
          if (live) {
#pragma unroll
            for (int e = 0; e < 4; ++e) dst[4 * q + e] = cu_pack8(v + 8 * e);
          }
        }
      }
    }
  } else if (warp >= 2 + CU_ROW_WARPS) {
    // (V3: the filler warp between the store warp and the drain warp has no work)
  } else {
    // ===================== row warps (8): thread pair == tile row == (cell, clip) =====================
    // warp w serves TMEM lane quadrant w % 4; the two warps of a quadrant (grp 0 / 1) split every per-row loop by
    // columns, so each SM scheduler interleaves two of these latency-bound warps.
    const int quad = warp % 4;
    const int grp = (warp - 2) >> 2;                      // which CU_COLS-column share of a row this thread owns
    const int r = quad * 32 + lane;                       // row within the tile == TMEM lane
    const int at = threadIdx.x - 64;                      // 0..255
    const bool issuer = at == 0;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t row_off = (uint32_t)((r & 7) * 16 + (r >> 3) * 2048);        // Cs: own row, chunk 0
    const uint32_t prow_off = (uint32_t)((r & 7) * 16 + (r >> 3) * (KG * 128)); // Ps: own row, chunk 0
    const float inv_sqrt_dl = 1.0f / sqrtf((float)CU_DL);
    float4* s_gg = reinterpret_cast<float4*>(Ps);         // partial Grams, exchanged between a row's two threads (Ps is dead by then)
    if (at < CU_DL) s_b1[at] = bias1[at];
    if (V == 1) {
      for (int e = at; e < D; e += CU_ROW_THREADS) s_b2[e] = bias2[e];
    } else if (at < 256) {        // I_16, K-major core-matrix layout (lbo 128, sbo 256): element (n, k) = (n == k)
      const int n = at >> 4, k = at & 15;
      *reinterpret_cast<uint16_t*>(I16 + (n & 7) * 16 + (n >> 3) * 256 + (k >> 3) * 128 + (k & 7) * 2) = n == k ? (uint16_t)0x3F80 : (uint16_t)0;
    }
    if (V >= 3) {
      for (int e = at; e < CU_M4_BYTES / 4; e += CU_ROW_THREADS) reinterpret_cast<uint32_t*>(M4)[e] = 0u;
      CU_ROW_BAR();
      if (at < 16) {                                         // P[m', k] = 0.25 for k / 4 == m' (m' = 0..3), row index m' + 124
        const int mp = at >> 2, k = at;
        *reinterpret_cast<uint16_t*>(M4 + (mp + 124) * 16 + (k >> 3) * (CU_M4_ROWS * 16) + (k & 7) * 2) = (uint16_t)0x3E80;
      }
    }
    if (V >= 2) ptx::fence_proxy_async();     // I_16, M4 -> visible to the MMAs (ordered before the first cc_ready / sready arrive)
    uint32_t s_phase = 0, a_phase = 0, yi = 0, it = 0;
    int staged_bg = -1;                                   // sample group whose query operands sit in Ks / Wt / s_beta / s_mask
    for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const uint32_t ph = it & 1;
      const int m0 = tile * UG_BM;
      const int row = m0 + r;
      const bool valid = row < M;
      const int b = valid ? (code[row >> 2] >> 16) : -1;
      VML_DBG_ASSERT(!valid || (b >= 0 && b < B));
      const int b_first = code[tile * (UG_BM / 4)] >> 16;
      const int b_last = code[(min(m0 + UG_BM, M) - 1) >> 2] >> 16;
      const int ngroups = (b_last - b_first) / GS + 1;
      for (int g = 0; g < ngroups; ++g) {
        const int bg = b_first + GS * g;
        const int sl = b - bg;                              // 0 .. GS-1: this row's sample is in the group
        const bool mine = valid && sl >= 0 && sl < GS;
        if (g == 0) CU_T(0);
        CU_ROW_BAR();                                       // previous users of U / side data / Cs are done
        // ---- stage the group's query-side operands (a CTA walks consecutive tiles, so the group often repeats) ----
        if (bg != staged_bg) {
          for (int e = at; e < NW * (CU_DL / 8); e += CU_ROW_THREADS) {
            const int w = e / (CU_DL / 8), ch = e % (CU_DL / 8);   // word slot, 8-feature chunk
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
            *reinterpret_cast<uint4*>(Ks + (w & 7) * 16 + (w >> 3) * 2048 + ch * 128) = cu_pack8(kt);
            *reinterpret_cast<uint4*>(Wt + (w & 7) * 16 + ch * (KG * 128) + (w >> 3) * 128) = cu_pack8(wh);
          }
          if (at < NW) {
            const int s2 = at / NQP, k = at % NQP;
            const int bb = min(bg + s2, B - 1);
            s_beta[at] = k < Nq ? qproj[((size_t)bb * Nq + k) * ld + off_beta] : 0.f;
            s_mask[at] = (k < Nq && qmask[(size_t)bb * Nq + k]) ? 1.f : 0.f;
          }
        }
        staged_bg = bg;
        if (g == 0) CU_T(1);
        if (g == 0) {
          // ---- c_hat row out of TMEM: + bias, round to bf16 (what the unfused path stores), park in Cs ----
          ptx::mbar_wait(chat_full, ph);
          if (V == 2 || V == 3) ptx::mbar_wait(cs_free, ph ^ 1);        // the previous tile's last column block has left Cs
          CU_T(2);
          ptx::tc_fence_after();
          const uint32_t t_addr = tmem_base + lane_base;
#pragma unroll 1
          for (int c = grp * CU_COLS; c < grp * CU_COLS + CU_COLS; c += 32) {
            float v[32];
            ptx::tmem_ld32(t_addr + (uint32_t)c, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; e += 8) {
              float t[8];
#pragma unroll
              for (int q = 0; q < 8; q += 2) {
                t[q] = v[e + q]; t[q + 1] = v[e + q + 1];
                ptx::add2(t[q], t[q + 1], s_b1[c + e + q], s_b1[c + e + q + 1]);
              }
              *reinterpret_cast<uint4*>(Cs + row_off + ((c + e) >> 3) * 128) = cu_pack8(t);
            }
          }
          if (V == 4) {                                        // accumulator read: the next tile's first contraction may overwrite it
            ptx::tc_fence_before();
            ptx::mbar_arrive(chat_free);
          }
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        CU_ROW_BAR();
        if (g == 0) CU_T(3);
        if (issuer) {                                          // ---- (2) S = c_hat . ktil^T ----
          ptx::tc_fence_after();
          constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(UG_BM, NW);
          const uint32_t a0 = ptx::smem_u32(Cs), b0 = ptx::smem_u32(Ks);
#pragma unroll
          for (int k = 0; k < CU_DL / 16; ++k)
            ptx::umma_bf16(tmem_base + CU_TMEM_S, ptx::umma_desc_nosw(a0 + k * 256, 128, 2048),
                           ptx::umma_desc_nosw(b0 + k * 256, 128, 2048), idesc_s, k != 0);
          ptx::umma_commit(sfull_bar);
        }
        if (grp == 0) {
          ptx::mbar_wait(sfull_bar, s_phase);
          if (g == 0) CU_T(4);
          ptx::tc_fence_after();
          // ---- masked softmax over this row's words (models.py:211-220), P row -> Ps ----------------------
          float sv[NW];
#pragma unroll
          for (int c = 0; c < NW; c += 16) ptx::tmem_ld16(tmem_base + lane_base + CU_TMEM_S + (uint32_t)c, sv + c);
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
            const uint4 pk = cu_pack8(p + kc * 8);
            if (GS == 2) {
              *reinterpret_cast<uint4*>(Ps + prow_off + kc * 128) = sl == 0 ? pk : zero4;
              *reinterpret_cast<uint4*>(Ps + prow_off + (NQP / 8 + kc) * 128) = sl == 1 ? pk : zero4;
            } else {
              *reinterpret_cast<uint4*>(Ps + prow_off + kc * 128) = pk;      // rows of other samples carry zeros (inv_den = 0)
            }
          }
          ptx::fence_proxy_async();
        }
        s_phase ^= 1;
        ptx::tc_fence_before();
        CU_ROW_BAR();
        if (g == 0) CU_T(5);
        if (issuer) {                                          // ---- (3) A (+)= P . [w_hat ; s_hat] ----
          ptx::tc_fence_after();
          constexpr uint32_t idesc_a = ptx::umma_idesc_bf16_bmn(UG_BM, CU_DL);
          const uint32_t a0 = ptx::smem_u32(Ps), b0 = ptx::smem_u32(Wt);
#pragma unroll
          for (int k = 0; k < NW / 16; ++k)
            ptx::umma_bf16(tmem_base + CU_TMEM_A, ptx::umma_desc_nosw(a0 + k * 256, 128, KG * 128),
                           ptx::umma_desc_nosw(b0 + k * 256, 128, KG * 128), idesc_a, (g | k) != 0);
          ptx::umma_commit(afull_bar);
        }
        ptx::mbar_wait(afull_bar, a_phase); a_phase ^= 1;
        ptx::tc_fence_after();
      }
      CU_T(6);
      // ---- gate G = c_hat * (A + s_hat), Gram of the cell's 4 clips (adjacent lanes); this thread: 64 columns ----
      float gg[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = grp * CU_COLS; c < grp * CU_COLS + CU_COLS; c += 32) {
        float a[32];
        ptx::tmem_ld32(tmem_base + lane_base + CU_TMEM_A + (uint32_t)c, a);
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
        // Gram of the cell's 4 clips: the products with a partner row are computed once per pair -- the lane with the
        // selector bit clear takes columns [0,16) of the chunk, its partner [16,32), then the two halves are exchanged
        // and added (48 shuffles per chunk instead of 96; both lanes of a pair end with the same bits)
        float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
        {
          const bool h1 = (lane & 1) != 0, h2 = (lane & 2) != 0;
#pragma unroll
          for (int e = 0; e < 32; ++e) g0 = fmaf(a[e], a[e], g0);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float mine1 = h1 ? a[16 + i] : a[i], send1 = h1 ? a[i] : a[16 + i];      // partners r ^ 1 and r ^ 3 differ in bit 0
            const float mine2 = h2 ? a[16 + i] : a[i], send2 = h2 ? a[i] : a[16 + i];      // partner r ^ 2 differs in bit 1
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
      CU_ROW_BAR();
      {
        // (columns 0..63) + (columns 64..127), the same order in every thread of the row (and in content_tc.cu)
        float4 lo = s_gg[r], hi = s_gg[(CU_SPLIT / 2) * 128 + r];
        if (CU_SPLIT == 4) {
          const float4 l1 = s_gg[128 + r], h1 = s_gg[3 * 128 + r];
          lo = make_float4(lo.x + l1.x, lo.y + l1.y, lo.z + l1.z, lo.w + l1.w);
          hi = make_float4(hi.x + h1.x, hi.y + h1.y, hi.z + h1.z, hi.w + h1.w);
        }
        gg[0] = lo.x + hi.x; gg[1] = lo.y + hi.y; gg[2] = lo.z + hi.z; gg[3] = lo.w + hi.w;
      }
      CU_T(7);
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
      // cc_hat row written IN PLACE of the c_hat row: a row's siblings (r ^ 1..3) are lanes of the same warp, and
      // every chunk is read by all four before any of them overwrites it (__syncwarp between read and write)
#pragma unroll 4
      for (int c = grp * CU_COLS; c < grp * CU_COLS + CU_COLS; c += 8) {
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
        *reinterpret_cast<uint4*>(Cs + row_off + (c >> 3) * 128) = cu_pack8(o.v);
      }
      ptx::fence_proxy_async();                              // cc_hat -> visible to the tail MMAs
      ptx::mbar_arrive(cc_ready);
      CU_T(8);

      // ---- (4) epilogue: out = Y + b_c + X + fbar in place, side = mean over the cell's 4 clips; this thread: the
      //      64-column box `grp` of every 128-column block.  fbar comes straight from global memory (L2: the boundary
      //      unit has just written it), fetched before the wait for the accumulator ------------------------------
      constexpr int NP = CU_COLS / 8;                        // 16-byte pieces per thread and column block
      const bf16* frow = fbar + (size_t)(row >> 2) * D + grp * CU_COLS;
      bf16* srow = side + (size_t)(row >> 2) * ld_side + grp * CU_COLS;
      const int box_of = (grp * CU_COLS) >> 6, pc0 = ((grp * CU_COLS) & 63) >> 3;   // 64-column box and first piece inside it
      // fbar is ~1 us away (L2 / HBM): block nb + 1's pieces are requested right after block nb has consumed its own,
      // and the first half of a block's arithmetic (accumulator + bias + residual) runs while they are in flight
      uint4 fq[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) fq[i] = valid ? __ldg(reinterpret_cast<const uint4*>(frow) + i) : make_uint4(0, 0, 0, 0);
      for (int nb = 0; nb < NB; ++nb, ++yi) {
        const uint32_t yb = yi & 1, yph = (yi >> 1) & 1;
        VML_DBG_ASSERT(V != 4 || (box_of >= 0 && box_of < 2 && pc0 + NP <= 8));
        unsigned char* xb = V == 4 ? Ring + ((base_tail((int)it) + 4 * nb + 2 + box_of) % RB) * CU_BOX
                            : (V >= 2 && nb == NB - 1) ? Cs + box_of * CU_BOX : Xs + (2 * nb + box_of) * CU_BOX;
        ptx::mbar_wait_relaxed(&yfull[yb], yph);              // TMEM data: ordered by the tcgen05 fence below
        CU_T(9 + 2 * nb);
        ptx::tc_fence_after();
        const uint32_t t_addr = tmem_base + lane_base + (yb ? CU_TMEM_Y1 : CU_TMEM_Y0) + (uint32_t)(grp * CU_COLS);
        float acc[CU_COLS];
        ptx::tmem_ld32(t_addr, acc);
        if (CU_COLS == 64) ptx::tmem_ld32(t_addr + 32u, acc + (CU_COLS - 32));
        ptx::tmem_ld_wait();
        if constexpr (V == 1) {
        const float* bcol = s_b2 + nb * 128 + grp * CU_COLS;
#pragma unroll
        for (int pc = 0; pc < NP; ++pc) {
          const f8 xv = unpack8(*reinterpret_cast<const uint4*>(xb + ptx::sw128_off(r, pc0 + pc)));
#pragma unroll
          for (int q = 0; q < 8; q += 2) {
            float* a = acc + pc * 8 + q;
            ptx::add2(a[0], a[1], bcol[pc * 8 + q], bcol[pc * 8 + q + 1]);       // (acc + bias)
            ptx::add2(a[0], a[1], xv.v[q], xv.v[q + 1]);                          //   + residual
          }
        }
#pragma unroll
        for (int i = 0; i < NP; ++i)     // fbar is consumed only from here on (keeps the unpacking below the loop above)
          asm volatile("" : "+r"(fq[i].x), "+r"(fq[i].y), "+r"(fq[i].z), "+r"(fq[i].w));
#pragma unroll
        for (int pc = 0; pc < NP; ++pc) {
          const f8 fv = unpack8(fq[pc]);
          f8 o;
#pragma unroll
          for (int q = 0; q < 8; q += 2) {
            o.v[q] = acc[pc * 8 + q]; o.v[q + 1] = acc[pc * 8 + q + 1];
            ptx::add2(o.v[q], o.v[q + 1], fv.v[q], fv.v[q + 1]);                  //   + fbar
          }
          *reinterpret_cast<uint4*>(xb + ptx::sw128_off(r, pc0 + pc)) = cu_pack8(o.v);      // result in place of the residual
          // mean over the cell's 4 rows (adjacent lanes) as a reduce-scatter: 6 shuffles per piece instead of 16, every
          // lane ends with 2 of the piece's 8 columns.  Same association as the butterfly ((r0 + r1) + (r2 + r3)).
          {
            const bool hi1 = (lane & 1) != 0, hi2 = (lane & 2) != 0;
            float k[4];
#pragma unroll
            for (int i = 0; i < 4; i += 2) {
              k[i] = hi1 ? o.v[4 + i] : o.v[i]; k[i + 1] = hi1 ? o.v[5 + i] : o.v[i + 1];
              const float s0 = hi1 ? o.v[i] : o.v[4 + i], s1 = hi1 ? o.v[i + 1] : o.v[5 + i];
              ptx::add2(k[i], k[i + 1], __shfl_xor_sync(0xffffffffu, s0, 1), __shfl_xor_sync(0xffffffffu, s1, 1));
            }
            float m0 = hi2 ? k[2] : k[0], m1 = hi2 ? k[3] : k[1];
            const float s0 = hi2 ? k[0] : k[2], s1 = hi2 ? k[1] : k[3];
            ptx::add2(m0, m1, __shfl_xor_sync(0xffffffffu, s0, 2), __shfl_xor_sync(0xffffffffu, s1, 2));
            ptx::mul2(m0, m1, 0.25f, 0.25f);
            if (valid)
              *reinterpret_cast<__nv_bfloat162*>(srow + nb * 128 + pc * 8 + (hi1 ? 4 : 0) + (hi2 ? 2 : 0)) = __floats2bfloat162_rn(m0, m1);
          }
        }
        } else {
          // V2: the accumulator already holds cc_hat.W2^T + X (tensor cores); fbar carries the output bias.
          if (V >= 3 && !valid) {            // rows past the live count feed the mean_c MMA (0 x NaN = NaN): keep them finite
#pragma unroll
            for (int e = 0; e < CU_COLS; ++e) acc[e] = 0.f;
          }
#pragma unroll
          for (int pc = 0; pc < NP; ++pc) {
            const f8 fv = unpack8(fq[pc]);
            float* a = acc + pc * 8;
#pragma unroll
            for (int q = 0; q < 8; q += 2) ptx::add2(a[q], a[q + 1], fv.v[q], fv.v[q + 1]);
            *reinterpret_cast<uint4*>(xb + ptx::sw128_off(r, pc0 + pc)) = cu_pack8(a);      // result in place of the residual
          }
          if constexpr (V == 2) {
          __syncwarp();                                        // the warp's 32 rows x CU_COLS columns are in shared memory
          // mean over a cell's 4 clips from the finished bf16 tile: one lane per (cell, 8-column piece); the four 16-byte
          // reads of a quarter warp hit 8 distinct bank groups of the 128B swizzle (conflict-free), no shuffles
#pragma unroll
          for (int item = lane; item < 8 * NP; item += 32) {
            const int c = item / NP, p = item % NP;            // cell within this warp's 8, piece within its CU_COLS columns
            const int r0 = quad * 32 + 4 * c;
            f8 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = unpack8(*reinterpret_cast<const uint4*>(xb + ptx::sw128_off(r0 + q, pc0 + p)));
            float o[8];
#pragma unroll
            for (int q = 0; q < 8; q += 2) {
              float a0 = v[0].v[q], a1 = v[0].v[q + 1], b0 = v[2].v[q], b1 = v[2].v[q + 1];
              ptx::add2(a0, a1, v[1].v[q], v[1].v[q + 1]);                         // (r0 + r1)
              ptx::add2(b0, b1, v[3].v[q], v[3].v[q + 1]);                         // (r2 + r3)
              ptx::add2(a0, a1, b0, b1);
              ptx::mul2(a0, a1, 0.25f, 0.25f);
              o[q] = a0; o[q + 1] = a1;
            }
            if (m0 + r0 < M)
              *reinterpret_cast<uint4*>(side + (size_t)((m0 + r0) >> 2) * ld_side + nb * 128 + grp * CU_COLS + p * 8) = cu_pack8(o);
          }
          }
        }
        CU_T(38 + nb);
        if (nb + 1 < NB) {
#pragma unroll
          for (int i = 0; i < NP; ++i)
            fq[i] = valid ? __ldg(reinterpret_cast<const uint4*>(frow + (nb + 1) * 128) + i) : make_uint4(0, 0, 0, 0);
        }
        CU_T(34 + nb);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&yempty[yb]);          // accumulator drained
        ptx::fence_proxy_async();                              // shared-memory writes -> visible to the TMA stores
        ptx::mbar_arrive(&sready[yb]);                         // the store warp takes it from here; no barrier among the row warps
        CU_T(10 + 2 * nb);
      }
      CU_T(17);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tmem_base); }
}

template <int NQP, int GS, int V>
static int launch_content_unit(const CUtensorMap* tm, int grid, const bf16* fbar, bf16* side, int ld_side, const float* b1, const float* b2, const float* qproj, int ld,
                               int off_what, int off_ktil, int off_beta, const float* s_hat, int s_ld, const uint8_t* qmask,
                               vml_cells_t cells, int B, vml_dims_t d, int store_cu, cudaStream_t st) {
  using Cfg = CuCfg<NQP, GS, V>;
  static_assert(Cfg::SMEM <= 232448, "content_unit_kernel exceeds the 227 KB shared-memory limit");
  VML_CUDA(ensure_dyn_smem((const void*)(content_unit_kernel<NQP, GS, V>), (size_t)(Cfg::SMEM)));
  content_unit_kernel<NQP, GS, V><<<grid, V >= 3 ? CU_THREADS_V2 : CU_THREADS, Cfg::SMEM, st>>>(tm[0], tm[1], tm[2], tm[3], d.D, fbar, side, ld_side, b1, b2, qproj, ld,
                                                                    off_what, off_ktil, off_beta, s_hat, s_ld, qmask, cells.code,
                                                                    cells.n_cells, d.Nq, B, store_cu);
  VML_LAUNCHED(1);
  return VML_OK;
}

int content_unit_pp(const CUtensorMap* tm, int grid, const void* fbar, void* side, int ld_side, const float* b1, const float* qproj,
                    int ld, int off_what, int off_ktil, int off_beta, const float* s_hat, int s_ld, const uint8_t* qmask,
                    vml_cells_t cells, int B, vml_dims_t d, int store_cu, cudaStream_t st);      // content_unit_pp.cu

bool content_unit_supported(vml_dims_t d) {
  return d.dl == CU_DL && d.C == 4 && d.D % 128 == 0 && d.D <= CU_MAXKB * 64 && d.Nq <= 31;
}

// fc, cu bf16 [cap*4, D]; W1 bf16 [128, D]; W2 bf16 [D, 128]; fbar bf16 [cap, D]; side bf16 [cap, D] with row
// stride ld_side (may point into a wider matrix).  store_cu = 0: only the side output is produced (last layer).
int content_unit(const void* fc, const void* W1, const float* b1, const float* qproj, int ld, int off_what, int off_ktil,
                 int off_beta, const float* s_hat, int s_ld, const uint8_t* qmask, vml_cells_t cells, const void* W2,
                 const float* b2, const void* fbar, void* cu, void* side, int ld_side, int B, vml_dims_t d, int store_cu,
                 int bias_in_fbar, cudaStream_t st) {
  VML_CHECK_ARG(content_unit_supported(d) && ld % 4 == 0 && off_what % 4 == 0 && off_ktil % 4 == 0 && s_ld % 4 == 0 &&
                ld_side % 8 == 0 && cells.capacity > 0);
  static bool reg = (register_kernel("content_unit_kernel"), true); (void)reg;
  CUtensorMap tm[4];
  const uint64_t M = (uint64_t)cells.capacity * 4, D = (uint64_t)d.D;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tm[0], fc, M, D, D, UG_BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tm[1], W1, CU_DL, D, D, CU_DL))) return rc;
  if ((rc = make_tmap_bf16_2d(&tm[2], W2, D, CU_DL, CU_DL, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tm[3], cu, M, D, D, UG_BM))) return rc;
  VML_CHECK_ARG((reinterpret_cast<uintptr_t>(fbar) & 15) == 0 && (reinterpret_cast<uintptr_t>(side) & 15) == 0);
  const int tiles = ceil_div((int)M, UG_BM);
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
#define VML_CU(NQP, GS, V) return launch_content_unit<NQP, GS, V>(tm, grid, (const bf16*)fbar, (bf16*)side, ld_side, b1, b2, qproj, ld, off_what, off_ktil, off_beta, s_hat, s_ld, qmask, cells, B, d, store_cu, st)
  if (bias_in_fbar) {            // V2 / V3: output bias folded into fbar by the boundary unit, residual on the tensor cores
    const int variant = getenv("VML_CU_VARIANT") != nullptr ? atoi(getenv("VML_CU_VARIANT")) : 4;   // (A/B knob)
    if (variant == 5)              // ping-pong schedule: two row-warp groups on alternate tiles (content_unit_pp.cu)
      return content_unit_pp(tm, grid, fbar, side, ld_side, b1, qproj, ld, off_what, off_ktil, off_beta, s_hat, s_ld, qmask, cells, B, d,
                             store_cu, st);
    if (variant == 2) {            // mean_c read back by the row warps
      if (d.Nq + 1 <= 8) VML_CU(8, 2, 2);
      if (d.Nq + 1 <= 16) VML_CU(16, 2, 2);
      VML_CU(32, 1, 2);
    }
    if (variant == 3) {            // resident tile, residual + mean_c on the tensor cores
      if (d.Nq + 1 <= 8) VML_CU(8, 2, 3);
      if (d.Nq + 1 <= 16) VML_CU(16, 2, 3);
      VML_CU(32, 1, 3);
    }
    if (d.Nq + 1 <= 8) VML_CU(8, 2, 4);          // streamed tile (box ring), the default
    if (d.Nq + 1 <= 16) VML_CU(16, 2, 4);
    VML_CU(32, 1, 4);
  }
  if (d.Nq + 1 <= 8) VML_CU(8, 2, 1);
  if (d.Nq + 1 <= 16) VML_CU(16, 2, 1);
  VML_CU(32, 1, 1);
#undef VML_CU
}

}  // namespace vml

#ifdef VML_CU_TIMING
extern "C" __attribute__((visibility("default"))) int vml_debug_cu_timing(long long* host, int n) { return vml::cu_debug_read(host, n); }
#endif
