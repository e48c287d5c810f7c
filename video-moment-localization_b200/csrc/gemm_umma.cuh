// tcgen05 / TMEM / TMA GEMM for sm_100a:  out = A[M,K] . W[N,K]^T (+ epilogue), bf16 operands,
// fp32 accumulation in tensor memory.
//
// Persistent, warp-specialised CTA (320 threads):
//   warp 0      TMA producer   (cp.async.bulk.tensor 2D, 128B swizzle, mbarrier complete_tx)
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma 128 x BN x 16, commits to mbarriers)
//   warps 2..9  epilogue       (tcgen05.ld TMEM -> registers -> fused epilogue -> global); two warps
//               per TMEM lane quadrant, each taking half of the tile's columns, so every SM
//               scheduler has two epilogue warps to interleave (one alone runs at IPC ~0.2)
// Pipelines: smem full/empty ring (STAGES), TMEM full/empty double buffer (2 accumulators),
// so the epilogue of tile i overlaps the mainloop of tile i+1.
#pragma once
#include <stdlib.h>
#include <type_traits>
#include <utility>

#include "common.cuh"
#include "sm100.cuh"

namespace vml {

constexpr int UG_BM = 128, UG_BK = 64, UG_THREADS = 192;
constexpr int UG_EPI_WARPS = 8, UG_GEMM_THREADS = 64 + 32 * UG_EPI_WARPS;   // generic GEMM: 8 epilogue warps

// epilogues that use warp shuffles across rows must be called by all 32 lanes
template <typename E, typename = void>
struct epi_is_collective : std::false_type {};
template <typename E>
struct epi_is_collective<E, std::enable_if_t<E::kWarpCollective>> : std::true_type {};

// epilogues with global-memory operands can expose them as a prefetchable register bundle:
//   Pre load(row, col0, valid)  /  apply_pre<32>(row, col0, acc, pre, valid)
// The kernel then issues the loads of chunk c+1 (and of the next tile's first chunk, before it
// waits for the MMA) while chunk c is being finished, hiding the L2/HBM latency that a single
// epilogue warp per scheduler cannot hide by itself.
// epilogues of GEMMs that are bound by L2 -> shared-memory operand traffic ask for 2-CTA clusters: the two CTAs of a
// cluster work on vertically adjacent tiles and each loads HALF of the shared B (weight) tile, multicast to both
template <typename E, typename = void>
struct epi_wants_cluster : std::false_type {};
template <typename E>
struct epi_wants_cluster<E, std::enable_if_t<E::kCluster2>> : std::true_type {};

// an epilogue may cap the tile width (e.g. one whose shared-memory scratch is sized for 8 epilogue warps)
template <typename E, typename = void>
struct epi_max_bn : std::integral_constant<int, 256> {};
template <typename E>
struct epi_max_bn<E, std::void_t<decltype(E::kMaxBN)>> : std::integral_constant<int, E::kMaxBN> {};

// an epilogue may ask for per-warp shared-memory scratch (kScratchPerWarp bytes): the kernel places it behind the barrier block,
// the pipeline gives up stages if need be, and the epilogue is called as apply_pre_scr(..., scratch) by ALL lanes of the warp
template <typename E, typename = void>
struct epi_scratch : std::integral_constant<int, 0> {};
template <typename E>
struct epi_scratch<E, std::void_t<decltype(E::kScratchPerWarp)>> : std::integral_constant<int, E::kScratchPerWarp> {};

// an epilogue may GENERATE the leading k-blocks of the A operand on the SM instead of having them loaded (kGenWarps extra
// warps write the 128 x 64 bf16 box of such a k-block in the 128B-swizzled layout TMA would have produced; the stage's full
// barrier then also counts one arrival per generator warp).  Single-CTA kernel only.
template <typename E, typename = void>
struct epi_gen : std::integral_constant<int, 0> {};
template <typename E>
struct epi_gen<E, std::void_t<decltype(E::kGenWarps)>> : std::integral_constant<int, E::kGenWarps> {};

template <typename E, typename = void>
struct epi_has_pre : std::false_type {};
template <typename E>
struct epi_has_pre<E, std::void_t<typename E::Pre>> : std::true_type {};
// an epilogue may carry a run-time switch `pf_a`: the producer then asks L2 for the NEXT tile's A boxes while the current
// tile is being multiplied (the A rows of a tile are one contiguous block of memory when K spans the whole row)
template <typename E, typename = void>
struct epi_has_pf : std::false_type {};
template <typename E>
struct epi_has_pf<E, std::void_t<decltype(std::declval<const E&>().pf_a)>> : std::true_type {};

template <int BN, int SCR = 0 /* epilogue scratch bytes per CTA */>
struct UmmaCfg {
  static constexpr int A_BYTES = UG_BM * UG_BK * 2;
  static constexpr int B_BYTES = BN * UG_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int DEF_STAGES = BN >= 256 ? 4 : (BN >= 128 ? 5 : 6);
  static constexpr int FIT_STAGES = (232448 - 1024 - 256 - SCR) / STAGE_BYTES;
  static constexpr int STAGES = DEF_STAGES < FIT_STAGES ? DEF_STAGES : FIT_STAGES;
  static_assert(STAGES >= 2, "pipeline depth");
  static constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : (2 * BN);   // power of two for BN in {16,32,64,128,256}
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + SCR;
};

// EW epilogue warps (8 or 16): EW / 4 of them share a TMEM lane quadrant and split the tile's columns.  With 128 x 256
// tiles 8 warps need ~2x the main loop's time for a residual epilogue (measured: the moment GEMM ran at 53 % of the
// tensor pipe whatever fed its operands), hence 16 for BN = 256.
// TF32 = true: the operands are fp32 tensors read as TF32 (kind::tf32; the training forward): a 128-byte box row then holds
// 32 elements instead of 64, an MMA covers K = 8 instead of 16 -- the same 32 bytes per step, so only the K arithmetic, the
// instruction descriptor and the instruction kind differ.
template <int BN, typename Epi, int CL = 1, int EW = UG_EPI_WARPS, bool TF32 = false>
__global__ void __launch_bounds__(64 + 32 * EW + 32 * epi_gen<Epi>::value, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                 const int32_t* __restrict__ m_dev, int m_scale, Epi epi) {
  static_assert(CL == 1 || CL == 2, "single CTA or 2-CTA cluster");
  constexpr int GW = epi_gen<Epi>::value;
  static_assert(GW == 0 || (GW == 4 && CL == 1 && !TF32), "operand generator: 4 warps = 128 tile rows, single CTA, bf16");
  using Cfg = UmmaCfg<BN, EW * epi_scratch<Epi>::value>;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array: keeps the shared address space visible to the
  // compiler (LDS/STS instead of generic LD/ST, which an integer round-trip of the pointer would force)
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (m_dev) M = min(M, *m_dev * m_scale);
  constexpr int BK = TF32 ? UG_BK / 2 : UG_BK;             // elements per 128-byte box row
  const int tiles_m = (M + UG_BM - 1) / UG_BM, tiles_n = N / BN;
  const int k_blocks = (K + BK - 1) / BK;
  // CL == 2: a cluster walks PAIRS of vertically adjacent tiles (same n); rank r takes m-tile 2*pair_m + r (a tile past
  // the end still takes part in the B multicast and the barriers: its A rows are out of bounds = zero-filled)
  const int crank = CL == 2 ? (int)ptx::cluster_ctarank() : 0;
  const int num_tiles = CL == 2 ? ((tiles_m + 1) / 2) * tiles_n : tiles_m * tiles_n;     // tiles, or tile pairs
  const int tile0 = CL == 2 ? (int)blockIdx.x / 2 : (int)blockIdx.x, tile_step = CL == 2 ? (int)gridDim.x / 2 : (int)gridDim.x;
  auto tile_m0 = [&](int t) { return CL == 2 ? (2 * (t / tiles_n) + crank) * UG_BM : (t / tiles_n) * UG_BM; };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1 + GW); ptx::mbar_init(&empty_bar[s], CL); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], EW); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  if (CL == 2) ptx::cluster_sync_all();                    // the peer's barriers exist before anything is sent to them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int m0 = tile_m0(tile), n0 = (tile % tiles_n) * BN;
        if constexpr (epi_has_pf<Epi>::value && CL == 1) {
          // one CTA per m-tile asks (the tiles of the other n columns of that m-tile run at the same time on its neighbours)
          const int nt = tile + tile_step;
          if (epi.pf_a && nt < num_tiles && (nt % tiles_n == 0 || tile_step % tiles_n != 0))
            for (int kb = 0; kb < k_blocks; ++kb) ptx::tma_prefetch_2d(&tmA, kb * BK, tile_m0(nt));
        }
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);      // CL == 2: both CTAs' MMAs have retired from this stage
          unsigned char* sa = smem + stage * Cfg::STAGE_BYTES;
          bool gen_a = false;
          if constexpr (GW > 0) gen_a = epi.gen_kblock(kb);  // this k-block's A box is written by the generator warps
          ptx::mbar_arrive_expect_tx(&full_bar[stage], gen_a ? Cfg::B_BYTES : Cfg::STAGE_BYTES);
          if (!gen_a) ptx::tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
          if (CL == 2)                                       // my half of the B tile, to both CTAs of the cluster
            ptx::tma_load_2d_mc(sa + Cfg::A_BYTES + crank * (Cfg::B_BYTES / 2), &tmB, &full_bar[stage], kb * BK,
                                n0 + crank * (BN / 2), (uint16_t)3);
          else
            ptx::tma_load_2d(sa + Cfg::A_BYTES, &tmB, &full_bar[stage], kb * BK, n0);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = TF32 ? ptx::umma_idesc_tf32(UG_BM, BN, false, false) : ptx::umma_idesc_bf16(UG_BM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = ptx::umma_desc_sw128(a_addr);
          const uint64_t bdesc = ptx::umma_desc_sw128(a_addr + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < UG_BK / 16; ++k) {  // +32 B per K step (16 bf16 / 8 tf32) inside the 128B swizzle atom
            if (TF32) ptx::umma_tf32(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            else ptx::umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          if (CL == 2) ptx::umma_commit_mc(&empty_bar[stage], (uint16_t)3);   // ... in BOTH CTAs: the stage is refilled by multicast
          else ptx::umma_commit(&empty_bar[stage]);        // frees the smem stage when the MMAs retire
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull_bar[acc]);                 // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (GW > 0 && warp >= 2 + EW) {
    // ===================== operand generator warps =====================
    if constexpr (GW > 0) {
      const int gw = warp - 2 - EW;
      int stage = 0; uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int m0 = tile_m0(tile);
        typename Epi::Gen g;
        epi.gen_setup(g, m0, gw, lane, M);
        for (int kb = 0; kb < k_blocks; ++kb) {
          // (always: an arrival for the stage's NEXT use must not land in the phase of its current one)
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);      // the MMAs that read this stage have retired
          if (epi.gen_kblock(kb)) {
            epi.gen_fill(g, smem + stage * Cfg::STAGE_BYTES, kb, gw, lane);
            ptx::fence_proxy_async();                        // generic-proxy writes -> visible to the tensor core
          }
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&full_bar[stage]);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quad = warp % 4;                             // TMEM lane quadrant this warp may access
    static_assert(EW == 8 || (EW == 16 && BN >= 128), "epilogue warps: 8, or 16 for wide tiles");
    constexpr int HALF = EW == 16 ? BN / 4 : (BN >= 64 ? BN / 2 : BN);   // columns handled by this warp: [c_lo, c_hi)
    const int c_lo = ((warp - 2) / 4) * HALF;              // (BN = 32: the second warp of a quadrant has no columns)
    const int c_hi = c_lo + HALF < BN ? c_lo + HALF : BN;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int m0 = tile_m0(tile), n0 = (tile % tiles_n) * BN;
      const int row = m0 + quad * 32 + lane;
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      if constexpr (epi_has_pre<Epi>::value) {
        typename Epi::Pre pre = epi.load(row, n0 + c_lo, row < M && c_lo < BN);   // in flight while the MMA finishes
        ptx::mbar_wait(&tfull_bar[acc], acc_phase);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 32) {
          float v[32];
          ptx::tmem_ld32(t_addr + (uint32_t)c, v);
          if constexpr (epi_scratch<Epi>::value > 0) {
            unsigned char* scr = smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256 + (warp - 2) * epi_scratch<Epi>::value;
            if constexpr (EW == 16 && sizeof(typename Epi::Pre) > 64) {      // register budget: one bundle at a time (see below)
              ptx::tmem_ld_wait();
              epi.template apply_pre_scr<32>(row, n0 + c, v, pre, row < M, scr);
              if (c + 32 < c_hi) pre = epi.load(row, n0 + c + 32, row < M);
            } else {
              typename Epi::Pre nxt = pre;
              if (c + 32 < c_hi) nxt = epi.load(row, n0 + c + 32, row < M);
              ptx::tmem_ld_wait();
              epi.template apply_pre_scr<32>(row, n0 + c, v, pre, row < M, scr);
              pre = nxt;
            }
          } else if constexpr (EW == 16 && sizeof(typename Epi::Pre) > 64) {
            // 640 threads leave 96 registers: one bundle at a time (the next one is requested right after this chunk's
            // arithmetic, under the next TMEM load)
            ptx::tmem_ld_wait();
            epi.template apply_pre<32>(row, n0 + c, v, pre, row < M);
            if (c + 32 < c_hi) pre = epi.load(row, n0 + c + 32, row < M);
          } else {
            typename Epi::Pre nxt = pre;
            if (c + 32 < c_hi) nxt = epi.load(row, n0 + c + 32, row < M);
            ptx::tmem_ld_wait();
            epi.template apply_pre<32>(row, n0 + c, v, pre, row < M);
            pre = nxt;
          }
        }
      } else {
        ptx::mbar_wait(&tfull_bar[acc], acc_phase);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 32) {
          float v[32];
          ptx::tmem_ld32(t_addr + (uint32_t)c, v);
          ptx::tmem_ld_wait();
          if constexpr (epi_is_collective<Epi>::value) epi.template apply_warp<32>(row, n0 + c, v, row < M);
          else if (row < M) epi.template apply<32>(row, n0 + c, v);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL == 2) ptx::cluster_sync_all();                    // no CTA leaves while its peer may still multicast into it
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base); }
}

// ---- host side ---------------------------------------------------------------------------
// 2D bf16 tensor map, inner dimension = K (contiguous), 128B swizzle, box = {64, box_rows}.
// ---- CTA-pair variant (tcgen05.mma.cta_group::2, UMMA_M = 256) ---------------------------------------------------
// A cluster of two CTAs works on a 256 x BN tile: each CTA TMA-loads its own 128 A rows and its own HALF of the B tile
// (BN/2 weight rows), the leader's single thread issues one MMA for both SMs, and each CTA's epilogue warps drain the
// 128 accumulator rows that live in its own tensor memory.  Per MMA an SM now reads 4 KB of A + BN/2 rows of B from its
// shared memory instead of 4 KB + BN rows: the 128 B/clk read port that capped the single-CTA 128 x 256 tile at 67 %
// of the tensor pipe is no longer the limit.  Barriers: all TMA bytes of a stage count on the LEADER's full barrier;
// the leader's commits arrive on the empty / accumulator-full barriers of BOTH CTAs; both CTAs' epilogue warps
// arrive on the leader's accumulator-empty barrier.
template <int BN>
struct UmmaPairCfg {
  static constexpr int STAGES = 6;
  static constexpr int A_BYTES = UG_BM * UG_BK * 2;
  static constexpr int BH_BYTES = (BN / 2) * UG_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + BH_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "two accumulators of BN columns");
};

constexpr int UG_PAIR_EPI_WARPS = 16, UG_PAIR_THREADS = 64 + 32 * UG_PAIR_EPI_WARPS;

template <int BN, typename Epi>
__global__ void __launch_bounds__(UG_PAIR_THREADS, 1)
gemm_umma_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh, int M, int N, int K,
                      const int32_t* __restrict__ m_dev, int m_scale, Epi epi) {
  using Cfg = UmmaPairCfg<BN>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (m_dev) M = min(M, *m_dev * m_scale);
  const int tiles_m = (M + UG_BM - 1) / UG_BM, tiles_n = N / BN;
  const int k_blocks = (K + UG_BK - 1) / UG_BK;
  const int crank = (int)ptx::cluster_ctarank();
  const bool leader = crank == 0;
  const int num_pairs = ((tiles_m + 1) / 2) * tiles_n;
  const int pair0 = (int)blockIdx.x / 2, pair_step = (int)gridDim.x / 2;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmBh);
    for (int s = 0; s < Cfg::STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], 2 * UG_PAIR_EPI_WARPS); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();                                 // both CTAs' barriers and tensor memory exist
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {                                       // ===================== TMA producer (both CTAs) =====================
      int stage = 0; uint32_t phase = 0;
      for (int pair = pair0; pair < num_pairs; pair += pair_step) {
        const int m0 = (2 * (pair / tiles_n) + crank) * UG_BM, n0 = (pair % tiles_n) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);    // the pair's MMAs have retired from this stage
          unsigned char* sa = smem + stage * Cfg::STAGE_BYTES;
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);   // bytes of BOTH CTAs
          ptx::tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * UG_BK, m0);
          ptx::tma_load_2d_2sm(sa + Cfg::A_BYTES, &tmBh, &full_bar[stage], kb * UG_BK, n0 + crank * (BN / 2));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {                             // ===================== MMA issuer (leader only) =====================
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(2 * UG_BM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int pair = pair0; pair < num_pairs; pair += pair_step) {
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);   // both CTAs' epilogues have drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = ptx::umma_desc_sw128(a_addr);
          const uint64_t bdesc = ptx::umma_desc_sw128(a_addr + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < UG_BK / 16; ++k)
            ptx::umma_bf16_2sm(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          ptx::umma_commit_2sm(&empty_bar[stage], (uint16_t)3);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_2sm(&tfull_bar[acc], (uint16_t)3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs: the 128 rows in their own tensor memory) =====================
    const int quad = warp % 4;
    constexpr int HALF = BN / (UG_PAIR_EPI_WARPS / 4);     // columns per warp: 4 warps share a TMEM lane quadrant
    const int c_lo = ((warp - 2) / 4) * HALF, c_hi = c_lo + HALF;
    int acc = 0; uint32_t acc_phase = 0;
    for (int pair = pair0; pair < num_pairs; pair += pair_step) {
      const int m0 = (2 * (pair / tiles_n) + crank) * UG_BM, n0 = (pair % tiles_n) * BN;
      const int row = m0 + quad * 32 + lane;
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      typename Epi::Pre pre = epi.load(row, n0 + c_lo, row < M);       // in flight while the MMA finishes
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = c_lo; c < c_hi; c += 32) {
        float v[32];
        ptx::tmem_ld32(t_addr + (uint32_t)c, v);
        typename Epi::Pre nxt = pre;
        if (c + 32 < c_hi) nxt = epi.load(row, n0 + c + 32, row < M);
        ptx::tmem_ld_wait();
        epi.template apply_pre<32>(row, n0 + c, v, pre, row < M);
        pre = nxt;
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) ptx::mbar_arrive(&tempty_bar[acc]);
        else ptx::mbar_arrive_cluster(&tempty_bar[acc], 0);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();                                 // nobody leaves while the pair's MMAs / arrivals may still target it
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base); }
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t k, uint64_t row_stride_elems, uint32_t box_rows);

template <int BN, typename Epi>
int launch_gemm_umma_bn(const void* A, const void* W, int M, int N, int K, int lda, int ldw, const int32_t* m_dev,
                        int m_scale, const Epi& epi, cudaStream_t stream) {
  using Cfg = UmmaCfg<BN, UG_EPI_WARPS * epi_scratch<Epi>::value>;       // 8 epilogue warps
  using Cfg16 = UmmaCfg<BN, 16 * epi_scratch<Epi>::value>;               // 16 epilogue warps (wide tiles)
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, UG_BM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, BN);
  if (rc) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    VML_CUDA(ensure_dyn_smem((const void*)(gemm_umma_kernel<BN, Epi>), (size_t)(Cfg::SMEM_BYTES)));
    attr_done = true;
  }
  const int64_t tiles = (int64_t)ceil_div(M, UG_BM) * (N / BN);
  if constexpr (epi_wants_cluster<Epi>::value && epi_has_pre<Epi>::value && BN == 256) {
    // Opt-in (VML_GEMM_PAIR=1).  Validated bit-identical and measured neutral on B200 (ActivityNet 286 vs 288 us per layer,
    // Charades 35.5 vs 34.7 us): with its register epilogue (16-byte accesses to 32 different rows per instruction) and
    // ~1 GB of operand + residual + result traffic per ActivityNet layer, the moment GEMM sits at 53 % of BOTH the tensor
    // and the HBM roof; the pair lifts the shared-memory read port, which is not what binds.  Kept for the round that
    // moves the epilogue to TMA and fuses the bu_i*bu_j operand half into the producer.
    if (ceil_div(M, UG_BM) >= 2 && getenv("VML_GEMM_PAIR") != nullptr) {
      using PCfg = UmmaPairCfg<BN>;
      CUtensorMap tmBh;
      rc = make_tmap_bf16_2d(&tmBh, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, BN / 2);
      if (rc) return rc;
      static bool attrp = false;
      if (!attrp) {
        VML_CUDA(ensure_dyn_smem((const void*)(gemm_umma_pair_kernel<BN, Epi>), (size_t)(PCfg::SMEM_BYTES)));
        attrp = true;
      }
      const int64_t pairs = (int64_t)ceil_div(ceil_div(M, UG_BM), 2) * (N / BN);
      const int clusters = (int)(pairs < kNumSMs / 2 ? pairs : kNumSMs / 2);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(UG_PAIR_THREADS); cfg.dynamicSmemBytes = PCfg::SMEM_BYTES; cfg.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      VML_CUDA(cudaLaunchKernelEx(&cfg, gemm_umma_pair_kernel<BN, Epi>, tmA, tmBh, M, N, K, m_dev, m_scale, epi));
      VML_LAUNCHED(1);
      return VML_OK;
    }
  }
  if constexpr (epi_wants_cluster<Epi>::value && BN == 256) {
    // Opt-in (VML_GEMM_CLUSTER=1): measured on B200 it changes nothing (ActivityNet 285 vs 288 us, Charades 36 vs 35 us) --
    // the 128 x 256 tile is bound by the 128 B/clk shared-memory READ port of the MMA (12 KB of operands per 64-clk
    // instruction), not by L2 -> smem traffic; only cta_group::2 (each SM reads half of B) lifts that roof.
    if (ceil_div(M, UG_BM) >= 2 && getenv("VML_GEMM_CLUSTER") != nullptr) {
      // 2-CTA clusters: B tile loaded as two multicast halves (box of BN/2 rows), grid = an even number of CTAs
      CUtensorMap tmBh;
      rc = make_tmap_bf16_2d(&tmBh, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, BN / 2);
      if (rc) return rc;
      static bool attr2 = false;
      if (!attr2) {
        VML_CUDA(ensure_dyn_smem((const void*)(gemm_umma_kernel<BN, Epi, 2>), (size_t)(Cfg::SMEM_BYTES)));
        attr2 = true;
      }
      const int64_t pairs = (int64_t)ceil_div(ceil_div(M, UG_BM), 2) * (N / BN);
      const int clusters = (int)(pairs < kNumSMs / 2 ? pairs : kNumSMs / 2);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(UG_GEMM_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      VML_CUDA(cudaLaunchKernelEx(&cfg, gemm_umma_kernel<BN, Epi, 2>, tmA, tmBh, M, N, K, m_dev, m_scale, epi));
      VML_LAUNCHED(1);
      return VML_OK;
    }
  }
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  if constexpr (BN == 256) {
    if (getenv("VML_GEMM_EPI8") == nullptr) {               // (A/B knob) 16 epilogue warps for the wide tile
      static bool attr16 = false;
      if (!attr16) {
        VML_CUDA(ensure_dyn_smem((const void*)(gemm_umma_kernel<BN, Epi, 1, 16>), (size_t)(Cfg16::SMEM_BYTES)));
        attr16 = true;
      }
      gemm_umma_kernel<BN, Epi, 1, 16><<<grid, 64 + 32 * 16 + 32 * epi_gen<Epi>::value, Cfg16::SMEM_BYTES, stream>>>(tmA, tmB, M, N, K, m_dev, m_scale, epi);
      VML_LAUNCHED(1);
      return VML_OK;
    }
  }
  gemm_umma_kernel<BN, Epi><<<grid, UG_GEMM_THREADS + 32 * epi_gen<Epi>::value, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, M, N, K, m_dev, m_scale, epi);
  VML_LAUNCHED(1);
  return VML_OK;
}

int make_tmap_2d(CUtensorMap* map, int dtype, const void* ptr, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer, int swizzle);

// fp32 operands as TF32 (training forward): A float [M, lda], W float [N, ldw], both K-major; any float epilogue.
template <int BN, typename Epi>
int launch_gemm_umma_tf32_bn(const void* A, const void* W, int M, int N, int K, int lda, int ldw, const int32_t* m_dev,
                             int m_scale, const Epi& epi, cudaStream_t stream) {
  using Cfg = UmmaCfg<BN, (BN == 256 ? 16 : UG_EPI_WARPS) * epi_scratch<Epi>::value>;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d(&tmA, 1, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 4, UG_BK / 2, UG_BM, 3);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, 1, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw * 4, UG_BK / 2, BN, 3);
  if (rc) return rc;
  constexpr int EW = BN == 256 ? 16 : UG_EPI_WARPS;
  VML_CUDA(ensure_dyn_smem((const void*)(gemm_umma_kernel<BN, Epi, 1, EW, true>), (size_t)(Cfg::SMEM_BYTES)));
  const int64_t tiles = (int64_t)ceil_div(M, UG_BM) * (N / BN);
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  gemm_umma_kernel<BN, Epi, 1, EW, true><<<grid, 64 + 32 * EW, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, M, N, K, m_dev, m_scale, epi);
  VML_LAUNCHED(1);
  return VML_OK;
}
template <typename Epi>
int launch_gemm_umma_tf32(const void* A, const void* W, int M, int N, int K, int lda, int ldw, const int32_t* m_dev,
                          int m_scale, const Epi& epi, cudaStream_t stream) {
  VML_CHECK_ARG(lda % 4 == 0 && ldw % 4 == 0 && N % 32 == 0);
  VML_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0);
  if (M <= 0) return VML_OK;
  static bool reg = (register_kernel("gemm_umma_kernel<tf32>"), true);
  (void)reg;
  if (N % 256 == 0 && (int64_t)ceil_div(M, UG_BM) * (N / 256) >= 2 * kNumSMs)
    return launch_gemm_umma_tf32_bn<256, Epi>(A, W, M, N, K, lda, ldw, m_dev, m_scale, epi, stream);
  if (N % 128 == 0) return launch_gemm_umma_tf32_bn<128, Epi>(A, W, M, N, K, lda, ldw, m_dev, m_scale, epi, stream);
  if (N % 64 == 0) return launch_gemm_umma_tf32_bn<64, Epi>(A, W, M, N, K, lda, ldw, m_dev, m_scale, epi, stream);
  return launch_gemm_umma_tf32_bn<32, Epi>(A, W, M, N, K, lda, ldw, m_dev, m_scale, epi, stream);
}

// A: bf16 [M, lda] (K <= lda), W: bf16 [N, ldw].  M is the capacity; live rows = *m_dev * m_scale.
template <typename Epi>
int launch_gemm_umma(const void* A, const void* W, int M, int N, int K, int lda, int ldw, const int32_t* m_dev,
                     int m_scale, const Epi& epi, cudaStream_t stream) {
  VML_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && N % 32 == 0);
  VML_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0);
  if (M <= 0) return VML_OK;
  static bool reg = (register_kernel("gemm_umma_kernel"), true);
  (void)reg;
  // 128 x 256 tiles halve the shared-memory operand traffic per flop (the 128 x 128 shape sits exactly on the
  // 128 B/clk smem roof); used once there are enough tiles to fill the machine twice over
  const char* bn_knob = getenv("VML_GEMM_BN");              // (A/B knob) force the tile width: 128 or 256
  const int bn_force = bn_knob ? atoi(bn_knob) : 0;
  // 128 x 256 tiles halve the operand traffic per flop but need >= ~4 waves of them to balance 148 persistent CTAs; below
  // that (Charades pass: 340 wide tiles = 2.3 waves -> 3 rounds) twice as many 128 x 128 tiles finish earlier (measured
  // 41.2 vs 43.2 us per layer; TACoS / ActivityNet with >= 5 waves keep the wide tile: 105 vs 110 us)
  if constexpr (epi_max_bn<Epi>::value >= 256) {
    if (N % 256 == 0 && bn_force != 128 && ((int64_t)ceil_div(M, UG_BM) * (N / 256) >= 4 * kNumSMs || bn_force == 256))
      return launch_gemm_umma_bn<256, Epi>(A, W, M, N, K, lda, ldw, m_dev, m_scale, epi, stream);
  }
  if (N % 128 == 0) return launch_gemm_umma_bn<128, Epi>(A, W, M, N, K, lda, ldw, m_dev, m_scale, epi, stream);
  if (N % 64 == 0) return launch_gemm_umma_bn<64, Epi>(A, W, M, N, K, lda, ldw, m_dev, m_scale, epi, stream);
  return launch_gemm_umma_bn<32, Epi>(A, W, M, N, K, lda, ldw, m_dev, m_scale, epi, stream);
}

}  // namespace vml
