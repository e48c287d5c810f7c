// a5+a6 (front half) fused for sm_100a:   c_hat = fc.W_c_hat^T + b   ->   content-word attention,
// gate, CxC self-attention   ->   cc_hat,   in ONE persistent kernel (bf16 fast mode, dl = 128, C = 4).
//   reference: ContentUnit.forward models.py:246-266, ContentAttention.forward models.py:207-226
//
// The 128 x 128 x D contraction runs on tcgen05 (TMA producer warp, one-thread MMA issuer, fp32
// accumulator in TMEM, two accumulators).  The four "epilogue" warps are the attention warps: thread r
// owns tile row r = (cell, clip); it reads its c_hat row from TMEM once, accumulates the word scores
// on the fly, parks the row (bf16) in shared memory, releases the TMEM accumulator immediately (so the
// MMA of the next tile overlaps the attention of this one), then finishes softmax / attended words /
// gate / 4x4 clip self-attention.  The 4 clips of a cell are 4 adjacent lanes: the Gram matrix uses
// warp shuffles, the clip mixing reads sibling rows from shared memory.  c_hat never goes to HBM.
// The query-side operands (ktil, w_hat, s_hat, beta, mask) of the first two samples a tile touches
// are staged in shared memory by the attention warps while the MMA of the tile is in flight (cells
// are sorted by sample, so a 32-cell tile rarely spans more); rows of further samples fall back to
// L1-cached global loads.
#include "common.cuh"
#include "gemm_umma.cuh"
#include "sm100.cuh"

namespace vml {

constexpr int CF_DL = 128, CF_STAGES = 4, CF_ROWB = 272;  // smem row pitch of the parked c_hat tile (bytes): 16 mod 128 -> conflict-free 16B accesses
constexpr int CF_A_BYTES = UG_BM * UG_BK * 2, CF_B_BYTES = CF_DL * UG_BK * 2, CF_STAGE_BYTES = CF_A_BYTES + CF_B_BYTES;
constexpr int CF_SLOTS = 2;
template <int NQM>
struct CfSlot {   // per-sample query-side operands, fp32
  static constexpr int FLOATS = 2 * NQM * CF_DL + CF_DL + 2 * NQM;   // ktil | w_hat | s_hat | beta | mask
};
template <int NQM>
constexpr int cf_smem() { return CF_STAGES * CF_STAGE_BYTES + UG_BM * CF_ROWB + CF_SLOTS * CfSlot<NQM>::FLOATS * 4 + 1024 + 256; }

// One tile row: c_hat row out of TMEM -> cc_hat row.  STAGED: operands in shared memory, all NQM
// words present (zero-filled past Nq) so the word loops are branch-free; otherwise global memory
// and bounded by Nq.  Releases the TMEM accumulator as soon as the row has been read.
template <int NQM, bool STAGED>
__device__ __forceinline__ void attend_row(uint32_t t_addr, unsigned char* park, int r_in_tile, const float* __restrict__ bias,
                                           const float* kt_base, const float* wh_base, int kv_stride, const float* sh_base,
                                           const float* beta_base, int beta_stride, const float* mask_f,
                                           const uint8_t* mask_u8, int Nq, bool valid, bf16* out, uint64_t* tempty, int lane) {
  unsigned char* my_park = park + r_in_tile * CF_ROWB;
  const float sqrt_dl = sqrtf((float)CF_DL);
  // ---- phase 1: c_hat row out of TMEM, word scores on the fly, row parked as bf16 ----------
  float sc[NQM];
#pragma unroll
  for (int k = 0; k < NQM; ++k) sc[k] = 0.f;
#pragma unroll 1
  for (int c = 0; c < CF_DL; c += 32) {
    float v[32];
    ptx::tmem_ld32(t_addr + (uint32_t)c, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 32; e += 8) {                    // + bias, round to bf16 (what the unfused path stores), park
      f8 t;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        t.v[q] = __bfloat162float(__float2bfloat16_rn(v[e + q] + bias[c + e + q]));
        v[e + q] = t.v[q];
      }
      st8(reinterpret_cast<bf16*>(my_park) + c + e, t);
    }
#pragma unroll
    for (int k = 0; k < NQM; ++k) {
      if (STAGED || k < Nq) {
        const float4* kt = reinterpret_cast<const float4*>(kt_base + (size_t)k * kv_stride + c);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;     // 4 short chains instead of one 32-long
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float4 w = kt[e];
          s0 = fmaf(v[4 * e], w.x, s0); s1 = fmaf(v[4 * e + 1], w.y, s1); s2 = fmaf(v[4 * e + 2], w.z, s2); s3 = fmaf(v[4 * e + 3], w.w, s3);
        }
        sc[k] += (s0 + s1) + (s2 + s3);
      }
    }
  }
  // accumulator drained: let the MMA warp start the next tile while we finish the attention
  ptx::tc_fence_before();
  __syncwarp();
  if (lane == 0) ptx::mbar_arrive(tempty);

  // ---- phase 2: masked softmax over the words (models.py:211-220) ---------------------------
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < NQM; ++k) {
    if (k < Nq) {
      const float mk = STAGED ? mask_f[k] : (mask_u8[k] ? 1.f : 0.f);
      float s = (sc[k] + beta_base[(size_t)k * beta_stride]) / sqrt_dl;
      s = s * mk;
      if (mk == 0.f) s = -1e9f;
      sc[k] = s;
      mx = fmaxf(mx, s);
    }
  }
  float den = 0.f;
#pragma unroll
  for (int k = 0; k < NQM; ++k) {
    if (k < Nq) { sc[k] = expf(sc[k] - mx); den += sc[k]; } else sc[k] = 0.f;
  }
  const float inv_den = 1.0f / den;
#pragma unroll
  for (int k = 0; k < NQM; ++k) sc[k] *= inv_den;

  // ---- phase 3 per 32-column chunk: attended words, gate, Gram partials ---------------------
  float gg[4] = {0.f, 0.f, 0.f, 0.f};                  // G_r . G_{r^m}, m = 0..3 (siblings = adjacent lanes)
  __syncwarp();
#pragma unroll 1
  for (int c = 0; c < CF_DL; c += 32) {
    float a[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) a[e] = 0.f;
#pragma unroll
    for (int k = 0; k < NQM; ++k) {
      if (STAGED || k < Nq) {
        const float p = sc[k];
        const float4* wv = reinterpret_cast<const float4*>(wh_base + (size_t)k * kv_stride + c);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float4 w = wv[e];
          a[4 * e] = fmaf(p, w.x, a[4 * e]); a[4 * e + 1] = fmaf(p, w.y, a[4 * e + 1]);
          a[4 * e + 2] = fmaf(p, w.z, a[4 * e + 2]); a[4 * e + 3] = fmaf(p, w.w, a[4 * e + 3]);
        }
      }
    }
    const float4* sh = reinterpret_cast<const float4*>(sh_base + c);
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      const f8 ch = ld8(reinterpret_cast<const bf16*>(my_park) + c + e * 4);
      const float4 sa = sh[e], sb = sh[e + 1];
      a[4 * e] = ch.v[0] * (a[4 * e] + sa.x); a[4 * e + 1] = ch.v[1] * (a[4 * e + 1] + sa.y);
      a[4 * e + 2] = ch.v[2] * (a[4 * e + 2] + sa.z); a[4 * e + 3] = ch.v[3] * (a[4 * e + 3] + sa.w);
      a[4 * e + 4] = ch.v[4] * (a[4 * e + 4] + sb.x); a[4 * e + 5] = ch.v[5] * (a[4 * e + 5] + sb.y);
      a[4 * e + 6] = ch.v[6] * (a[4 * e + 6] + sb.z); a[4 * e + 7] = ch.v[7] * (a[4 * e + 7] + sb.w);
    }
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const float g = a[e];
      g0 = fmaf(g, g, g0);
      g1 = fmaf(g, __shfl_xor_sync(0xffffffffu, g, 1), g1);
      g2 = fmaf(g, __shfl_xor_sync(0xffffffffu, g, 2), g2);
      g3 = fmaf(g, __shfl_xor_sync(0xffffffffu, g, 3), g3);
    }
    gg[0] += g0; gg[1] += g1; gg[2] += g2; gg[3] += g3;
  }
  // ---- 4x4 clip self-attention (models.py:259-266): softmax over the cell's clips ---------
  float am = -INFINITY;
#pragma unroll
  for (int m = 0; m < 4; ++m) { gg[m] = gg[m] / sqrt_dl; am = fmaxf(am, gg[m]); }
  float ad = 0.f;
#pragma unroll
  for (int m = 0; m < 4; ++m) { gg[m] = expf(gg[m] - am); ad += gg[m]; }
  const float inv_ad = 1.0f / ad;
#pragma unroll
  for (int m = 0; m < 4; ++m) gg[m] *= inv_ad;
  const unsigned char* sib[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) sib[m] = park + (r_in_tile ^ m) * CF_ROWB;
#pragma unroll 4
  for (int c = 0; c < CF_DL; c += 8) {
    f8 o;
#pragma unroll
    for (int q = 0; q < 8; ++q) o.v[q] = 0.f;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const f8 sv = ld8(reinterpret_cast<const bf16*>(sib[m]) + c);
#pragma unroll
      for (int q = 0; q < 8; ++q) o.v[q] = fmaf(gg[m], sv.v[q], o.v[q]);
    }
    if (valid) st8(out + c, o);
  }
  __syncwarp();                                         // park rows are rewritten by the next tile
}

template <int NQM>
__global__ void __launch_bounds__(UG_THREADS, 1)
content_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int K,
                     const float* __restrict__ bias, const float* __restrict__ qproj, int ld, int off_what, int off_ktil,
                     int off_beta, const float* __restrict__ s_hat, int s_ld, const uint8_t* __restrict__ qmask,
                     const int32_t* __restrict__ code, const int32_t* __restrict__ n_cells, int Nq, int B,
                     bf16* __restrict__ cc_hat) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* park = smem + CF_STAGES * CF_STAGE_BYTES;                 // [128][CF_ROWB] bf16 c_hat rows
  float* slots = reinterpret_cast<float*>(park + UG_BM * CF_ROWB);         // [CF_SLOTS] staged samples
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(slots + CF_SLOTS * CfSlot<NQM>::FLOATS);
  uint64_t* empty_bar = full_bar + CF_STAGES;
  uint64_t* tfull_bar = empty_bar + CF_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int M = *n_cells * 4;
  const int num_tiles = (M + UG_BM - 1) / UG_BM;
  const int k_blocks = (K + UG_BK - 1) / UG_BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < CF_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<256>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {                                    // ===== TMA producer =====
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = smem + stage * CF_STAGE_BYTES;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], CF_STAGE_BYTES);
          ptx::tma_load_2d(sa, &tmA, &full_bar[stage], kb * UG_BK, tile * UG_BM);
          ptx::tma_load_2d(sa + CF_A_BYTES, &tmB, &full_bar[stage], kb * UG_BK, 0);
          if (++stage == CF_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                    // ===== MMA issuer =====
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(UG_BM, CF_DL);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * CF_DL);
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + stage * CF_STAGE_BYTES);
          const uint64_t adesc = ptx::umma_desc_sw128(a_addr), bdesc = ptx::umma_desc_sw128(a_addr + CF_A_BYTES);
#pragma unroll
          for (int k = 0; k < UG_BK / 16; ++k)
            ptx::umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == CF_STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===== attention warps: thread == tile row == (cell, clip) =====
    const int quad = warp % 4;
    const int r_in_tile = quad * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int row = tile * UG_BM + r_in_tile;
      const bool valid = row < M;
      const int b = valid ? (code[row >> 2] >> 16) : 0;
      // ---- stage the query-side operands of the tile's first CF_SLOTS samples (overlaps the MMA) ----
      const int b_first = code[tile * (UG_BM / 4)] >> 16;
      const int at = threadIdx.x - 64;                      // 0..127 among the attention threads
      asm volatile("bar.sync 1, 128;" ::: "memory");        // every warp is done with the previous tile's slots
      for (int sl = 0; sl < CF_SLOTS; ++sl) {
        const int bb = min(b_first + sl, B - 1);
        float* dst = slots + sl * CfSlot<NQM>::FLOATS;
        const float* src = qproj + (size_t)bb * Nq * ld;
        for (int e = at; e < NQM * (CF_DL / 4); e += 128) {  // words >= Nq are zero-filled: the row loops need no bound
          const int k = e / (CF_DL / 4), c4 = (e % (CF_DL / 4)) * 4;
          float4 kt = make_float4(0.f, 0.f, 0.f, 0.f), wh = kt;
          if (k < Nq) {
            kt = __ldg(reinterpret_cast<const float4*>(src + (size_t)k * ld + off_ktil + c4));
            wh = __ldg(reinterpret_cast<const float4*>(src + (size_t)k * ld + off_what + c4));
          }
          *reinterpret_cast<float4*>(dst + k * CF_DL + c4) = kt;
          *reinterpret_cast<float4*>(dst + (NQM + k) * CF_DL + c4) = wh;
        }
        if (at < CF_DL) dst[2 * NQM * CF_DL + at] = s_hat[(size_t)bb * s_ld + at];
        if (at < NQM) {
          dst[2 * NQM * CF_DL + CF_DL + at] = at < Nq ? src[(size_t)at * ld + off_beta] : 0.f;
          dst[2 * NQM * CF_DL + CF_DL + NQM + at] = (at < Nq && qmask[(size_t)bb * Nq + at]) ? 1.f : 0.f;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * CF_DL);
      const int sl = b - b_first;
      bf16* out = cc_hat + (size_t)row * CF_DL;
      if (__all_sync(0xffffffffu, sl >= 0 && sl < CF_SLOTS)) {
        // operands of this row's sample come from its staged slot (shared memory)
        const float* slot = slots + sl * CfSlot<NQM>::FLOATS;
        attend_row<NQM, true>(t_addr, park, r_in_tile, bias, slot, slot + NQM * CF_DL, CF_DL, slot + 2 * NQM * CF_DL,
                              slot + 2 * NQM * CF_DL + CF_DL, 1, slot + 2 * NQM * CF_DL + CF_DL + NQM, nullptr, Nq, valid, out,
                              &tempty_bar[acc], lane);
      } else {
        // rare: the warp touches a third sample of the tile -> L1-cached global operands
        const float* qrow = qproj + (size_t)b * Nq * ld;
        attend_row<NQM, false>(t_addr, park, r_in_tile, bias, qrow + off_ktil, qrow + off_what, ld, s_hat + (size_t)b * s_ld,
                               qrow + off_beta, ld, nullptr, qmask + (size_t)b * Nq, Nq, valid, out, &tempty_bar[acc], lane);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc<256>(tmem_base); }
}

// fc bf16 [cap*4, D]; W bf16 [128, D]; cc_hat bf16 [cap*4, 128]
int content_fused(const void* fc, const void* W, const float* bias, const float* qproj, int ld, int off_what, int off_ktil,
                  int off_beta, const float* s_hat, int s_ld, const uint8_t* qmask, vml_cells_t cells, void* cc_hat,
                  int B, vml_dims_t d, cudaStream_t st) {
  VML_CHECK_ARG(d.dl == CF_DL && d.C == 4 && d.Nq <= 24 && d.D % 8 == 0 && ld % 4 == 0 && off_what % 4 == 0 && off_ktil % 4 == 0);
  static bool reg = (register_kernel("content_fused_kernel"), true); (void)reg;
  CUtensorMap tmA, tmB;
  const int M = cells.capacity * 4;
  int rc = make_tmap_bf16_2d(&tmA, fc, (uint64_t)M, (uint64_t)d.D, (uint64_t)d.D, UG_BM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, W, (uint64_t)CF_DL, (uint64_t)d.D, (uint64_t)d.D, CF_DL);
  if (rc) return rc;
  const int tiles = ceil_div(M, UG_BM);
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  if (d.Nq <= 16) {
    VML_CUDA(cudaFuncSetAttribute(content_fused_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, cf_smem<16>()));
    content_fused_kernel<16><<<grid, UG_THREADS, cf_smem<16>(), st>>>(tmA, tmB, d.D, bias, qproj, ld, off_what, off_ktil, off_beta,
                                                                     s_hat, s_ld, qmask, cells.code, cells.n_cells, d.Nq, B,
                                                                     (bf16*)cc_hat);
  } else {
    VML_CUDA(cudaFuncSetAttribute(content_fused_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, cf_smem<24>()));
    content_fused_kernel<24><<<grid, UG_THREADS, cf_smem<24>(), st>>>(tmA, tmB, d.D, bias, qproj, ld, off_what, off_ktil, off_beta,
                                                                     s_hat, s_ld, qmask, cells.code, cells.n_cells, d.Nq, B,
                                                                     (bf16*)cc_hat);
  }
  VML_LAUNCHED(1);
  return VML_OK;
}

}  // namespace vml
