// a5+a6 (front half) as ONE persistent tcgen05 kernel (bf16 fast mode, dl = 128, C = 4):
//   c_hat = fc.W_c_hat^T + b  ->  content-word attention  ->  gate  ->  CxC clip self-attention  ->  cc_hat
//   reference: ContentUnit.forward models.py:246-266, ContentAttention.forward models.py:207-226
//
// All three contractions of the stage run on the tensor cores; c_hat, the word scores and the
// attended words never leave the SM:
//   (1) main   c_hat[128 x 128] = fc_tile[128 x D] . W^T         TMA-fed ring, accumulator in TMEM (x2)
//   (2) scores S[128 x NW]      = c_hat_bf16 . ktil^T            A = c_hat tile parked in smem by the
//                                                                attention warps, B = the W_q-folded keys
//   (3) attend A[128 x 128]     = P[128 x NW] . [w_hat ; s_hat]  A = softmax probabilities (bf16) written by
//                                                                the attention warps, B = word values (MN-major)
// A tile's rows belong to consecutive samples (cells are sorted by sample).  Operands (2)/(3) are
// staged per GROUP of two samples: NW = 2*NQP columns, a row uses its own sample's NQP columns and has
// exact zeros in the other half of P.  Tiles that span more than two samples loop over groups and
// accumulate (3) in TMEM.  One spare word slot per sample carries s_hat with probability 1, so that
// (3) directly yields  A + s_hat  (models.py:255-256).
//
// Warp roles: warp 0 TMA producer, warp 1 main-loop MMA issuer, warps 2..5 attention warps (thread ==
// tile row == (cell, clip); thread 64 also issues the two small MMAs).  The main loop of tile i+1 runs
// while the attention warps finish tile i (two c_hat accumulators).
#include "common.cuh"
#include "gemm_umma.cuh"
#include "sm100.cuh"

namespace vml {

constexpr int CT_DL = 128, CT_STAGES = 4;
constexpr int CT_A_BYTES = UG_BM * UG_BK * 2, CT_B_BYTES = CT_DL * UG_BK * 2, CT_STAGE_BYTES = CT_A_BYTES + CT_B_BYTES;
constexpr int CT_TMEM_S = 256, CT_TMEM_A = 320;      // TMEM columns: [0,256) two c_hat accumulators, S, A

template <int NQP>
struct CtCfg {
  static constexpr int NW = 2 * NQP;                 // word slots of a two-sample group
  static constexpr int KG = NW / 8;                  // 8-word groups
  static constexpr int CS_BYTES = UG_BM * CT_DL * 2; // c_hat tile, K-major, lbo 128, sbo 2048
  static constexpr int KS_BYTES = NW * CT_DL * 2;    // keys,   K-major (rows = word slots), lbo 128, sbo 2048
  static constexpr int WT_BYTES = NW * CT_DL * 2;    // values, MN-major (n = feature, k = word slot), lbo 128, sbo KG*128
  static constexpr int PS_BYTES = UG_BM * NW * 2;    // probabilities, K-major, lbo 128, sbo KG*128
  static constexpr int SIDE_FLOATS = 2 * NW + CT_DL; // beta | mask | bias
  static constexpr int SMEM = CT_STAGES * CT_STAGE_BYTES + CS_BYTES + KS_BYTES + WT_BYTES + PS_BYTES + SIDE_FLOATS * 4 + 1024 + 256;
};

__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  uint4 u; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return u;
}

template <int NQP>
__global__ void __launch_bounds__(UG_THREADS, 1)
content_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int K,
                  const float* __restrict__ bias, const float* __restrict__ qproj, int ld, int off_what, int off_ktil,
                  int off_beta, const float* __restrict__ s_hat, int s_ld, const uint8_t* __restrict__ qmask,
                  const int32_t* __restrict__ code, const int32_t* __restrict__ n_cells, int Nq, int B,
                  bf16* __restrict__ cc_hat) {
  using Cfg = CtCfg<NQP>;
  constexpr int NW = Cfg::NW, KG = Cfg::KG;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array: keeps the shared address space visible to the
  // compiler (LDS/STS instead of generic LD/ST, which an integer round-trip of the pointer would force)
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* Cs = smem + CT_STAGES * CT_STAGE_BYTES;
  unsigned char* Ks = Cs + Cfg::CS_BYTES;
  unsigned char* Wt = Ks + Cfg::KS_BYTES;
  unsigned char* Ps = Wt + Cfg::WT_BYTES;
  float* s_beta = reinterpret_cast<float*>(Ps + Cfg::PS_BYTES);     // [NW]
  float* s_mask = s_beta + NW;                                      // [NW]
  float* s_bias = s_mask + NW;                                      // [128]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_bias + CT_DL);
  uint64_t* empty_bar = full_bar + CT_STAGES;
  uint64_t* tfull_bar = empty_bar + CT_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* sfull_bar = tempty_bar + 2;     // scores ready
  uint64_t* afull_bar = sfull_bar + 1;      // attended words ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(afull_bar + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int M = *n_cells * 4;
  const int num_tiles = (M + UG_BM - 1) / UG_BM;
  const int k_blocks = (K + UG_BK - 1) / UG_BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < CT_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], 4); }
    ptx::mbar_init(sfull_bar, 1);
    ptx::mbar_init(afull_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
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
          unsigned char* sa = smem + stage * CT_STAGE_BYTES;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], CT_STAGE_BYTES);
          ptx::tma_load_2d(sa, &tmA, &full_bar[stage], kb * UG_BK, tile * UG_BM);
          ptx::tma_load_2d(sa + CT_A_BYTES, &tmB, &full_bar[stage], kb * UG_BK, 0);
          if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                    // ===== main-loop MMA issuer =====
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(UG_BM, CT_DL);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * CT_DL);
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + stage * CT_STAGE_BYTES);
          const uint64_t adesc = ptx::umma_desc_sw128(a_addr), bdesc = ptx::umma_desc_sw128(a_addr + CT_A_BYTES);
#pragma unroll
          for (int k = 0; k < UG_BK / 16; ++k)
            ptx::umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===== attention warps: thread == tile row == (cell, clip) =====
    const int quad = warp % 4;
    const int r = quad * 32 + lane;                       // row within the tile == TMEM lane
    const int at = threadIdx.x - 64;                      // 0..127 among the attention threads
    const bool issuer = at == 0;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t row_off = (uint32_t)((r & 7) * 16 + (r >> 3) * 2048);        // Cs: own row, chunk 0
    const uint32_t prow_off = (uint32_t)((r & 7) * 16 + (r >> 3) * (KG * 128)); // Ps: own row, chunk 0
    const float inv_sqrt_dl = 1.0f / sqrtf((float)CT_DL);
    if (at < CT_DL) s_bias[at] = bias[at];
    int acc = 0; uint32_t acc_phase = 0, s_phase = 0, a_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int row = tile * UG_BM + r;
      const bool valid = row < M;
      const int b = valid ? (code[row >> 2] >> 16) : -1;
      const int b_first = code[tile * (UG_BM / 4)] >> 16;
      const int b_last = code[(min(tile * UG_BM + UG_BM, M) - 1) >> 2] >> 16;
      const int ngroups = ((b_last - b_first) >> 1) + 1;
      for (int g = 0; g < ngroups; ++g) {
        const int bg = b_first + 2 * g;
        const int sl = b - bg;                              // 0 / 1: this row's sample is in the group
        const bool mine = valid && (sl == 0 || sl == 1);
        asm volatile("bar.sync 1, 128;" ::: "memory");      // previous users of Ks / Wt / side data / Cs are done
        // ---- stage the group's query-side operands (overlaps the main-loop MMA when g == 0) ----------
        for (int e = at; e < NW * (CT_DL / 8); e += 128) {
          const int w = e / (CT_DL / 8), ch = e % (CT_DL / 8);   // word slot, 8-feature chunk
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
          *reinterpret_cast<uint4*>(Ks + (w & 7) * 16 + (w >> 3) * 2048 + ch * 128) = pack8_bf16(kt);
          *reinterpret_cast<uint4*>(Wt + (w & 7) * 16 + ch * (KG * 128) + (w >> 3) * 128) = pack8_bf16(wh);
        }
        if (at < NW) {
          const int s2 = at / NQP, k = at % NQP;
          const int bb = min(bg + s2, B - 1);
          s_beta[at] = k < Nq ? qproj[((size_t)bb * Nq + k) * ld + off_beta] : 0.f;
          s_mask[at] = (k < Nq && qmask[(size_t)bb * Nq + k]) ? 1.f : 0.f;
        }
        if (g == 0) {
          // ---- c_hat row out of TMEM: + bias, round to bf16 (what the unfused path stores), park in Cs ----
          ptx::mbar_wait(&tfull_bar[acc], acc_phase);
          ptx::tc_fence_after();
          const uint32_t t_addr = tmem_base + lane_base + (uint32_t)(acc * CT_DL);
#pragma unroll 1
          for (int c = 0; c < CT_DL; c += 32) {
            float v[32];
            ptx::tmem_ld32(t_addr + (uint32_t)c, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; e += 8) {
              float t[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) t[q] = v[e + q] + s_bias[c + e + q];
              *reinterpret_cast<uint4*>(Cs + row_off + ((c + e) >> 3) * 128) = pack8_bf16(t);
            }
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);   // accumulator drained: next tile's main loop may run
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {                                          // ---- (2) S = c_hat . ktil^T ----
          ptx::tc_fence_after();
          constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(UG_BM, NW);
          const uint32_t a0 = ptx::smem_u32(Cs), b0 = ptx::smem_u32(Ks);
#pragma unroll
          for (int k = 0; k < CT_DL / 16; ++k)
            ptx::umma_bf16(tmem_base + CT_TMEM_S, ptx::umma_desc_nosw(a0 + k * 256, 128, 2048),
                           ptx::umma_desc_nosw(b0 + k * 256, 128, 2048), idesc_s, k != 0);
          ptx::umma_commit(sfull_bar);
        }
        ptx::mbar_wait(sfull_bar, s_phase); s_phase ^= 1;
        ptx::tc_fence_after();
        // ---- masked softmax over this row's words (models.py:211-220), P row -> Ps ----------------------
        {
          float sv[NW];
#pragma unroll
          for (int c = 0; c < NW; c += 16) ptx::tmem_ld16(tmem_base + lane_base + CT_TMEM_S + (uint32_t)c, sv + c);
          ptx::tmem_ld_wait();
          float p[NQP];
          float mx = -INFINITY;
#pragma unroll
          for (int k = 0; k < NQP; ++k) {
            const float raw = sl == 1 ? sv[NQP + k] : sv[k];
            const int slot = (sl == 1 ? NQP : 0) + k;
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
            const uint4 pk = pack8_bf16(p + kc * 8);
            *reinterpret_cast<uint4*>(Ps + prow_off + kc * 128) = sl == 0 ? pk : zero4;
            *reinterpret_cast<uint4*>(Ps + prow_off + (NQP / 8 + kc) * 128) = sl == 1 ? pk : zero4;
          }
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {                                          // ---- (3) A (+)= P . [w_hat ; s_hat] ----
          ptx::tc_fence_after();
          constexpr uint32_t idesc_a = ptx::umma_idesc_bf16_bmn(UG_BM, CT_DL);
          const uint32_t a0 = ptx::smem_u32(Ps), b0 = ptx::smem_u32(Wt);
#pragma unroll
          for (int k = 0; k < NW / 16; ++k)
            ptx::umma_bf16(tmem_base + CT_TMEM_A, ptx::umma_desc_nosw(a0 + k * 256, 128, KG * 128),
                           ptx::umma_desc_nosw(b0 + k * 256, 128, KG * 128), idesc_a, (g | k) != 0);
          ptx::umma_commit(afull_bar);
        }
        ptx::mbar_wait(afull_bar, a_phase); a_phase ^= 1;
        ptx::tc_fence_after();
      }
      // ---- gate G = c_hat * (A + s_hat), Gram of the cell's 4 clips (adjacent lanes) -------------------
      // (summed as two 64-column halves, the order content_unit.cu's thread pairs use: the kernels stay bit-identical)
      float gg[4] = {0.f, 0.f, 0.f, 0.f}, gh[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 0; c < CT_DL; c += 32) {
        float a[32];
        ptx::tmem_ld32(tmem_base + lane_base + CT_TMEM_A + (uint32_t)c, a);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          const f8 ch = unpack8(*reinterpret_cast<const uint4*>(Cs + row_off + ((c + e) >> 3) * 128));
#pragma unroll
          for (int q = 0; q < 8; ++q) a[e + q] = valid ? ch.v[q] * a[e + q] : 0.f;
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
        if (c < 64) { gg[0] += g0; gg[1] += g1; gg[2] += g2; gg[3] += g3; }
        else { gh[0] += g0; gh[1] += g1; gh[2] += g2; gh[3] += g3; }
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) gg[m] += gh[m];
      ptx::tc_fence_before();
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
      bf16* out = cc_hat + (size_t)row * CT_DL;
#pragma unroll 4
      for (int c = 0; c < CT_DL; c += 8) {
        f8 o;
#pragma unroll
        for (int q = 0; q < 8; ++q) o.v[q] = 0.f;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const f8 sv = unpack8(*reinterpret_cast<const uint4*>(Cs + sib[m] + (c >> 3) * 128));
#pragma unroll
          for (int q = 0; q < 8; ++q) o.v[q] = fmaf(gg[m], sv.v[q], o.v[q]);
        }
        if (valid) st8(out + c, o);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tmem_base); }
}

template <int NQP>
static int launch_content_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, int grid, const float* bias, const float* qproj,
                             int ld, int off_what, int off_ktil, int off_beta, const float* s_hat, int s_ld,
                             const uint8_t* qmask, vml_cells_t cells, void* cc_hat, int B, vml_dims_t d, cudaStream_t st) {
  VML_CUDA(ensure_dyn_smem((const void*)(content_tc_kernel<NQP>), (size_t)(CtCfg<NQP>::SMEM)));
  content_tc_kernel<NQP><<<grid, UG_THREADS, CtCfg<NQP>::SMEM, st>>>(tmA, tmB, d.D, bias, qproj, ld, off_what, off_ktil, off_beta,
                                                                    s_hat, s_ld, qmask, cells.code, cells.n_cells, d.Nq, B,
                                                                    (bf16*)cc_hat);
  VML_LAUNCHED(1);
  return VML_OK;
}

// fc bf16 [cap*4, D]; W bf16 [128, D]; cc_hat bf16 [cap*4, 128]
int content_tc(const void* fc, const void* W, const float* bias, const float* qproj, int ld, int off_what, int off_ktil,
               int off_beta, const float* s_hat, int s_ld, const uint8_t* qmask, vml_cells_t cells, void* cc_hat, int B,
               vml_dims_t d, cudaStream_t st) {
  VML_CHECK_ARG(d.dl == CT_DL && d.C == 4 && d.Nq <= 31 && d.D % 8 == 0 && ld % 4 == 0 && off_what % 4 == 0 && off_ktil % 4 == 0 &&
                s_ld % 4 == 0);
  static bool reg = (register_kernel("content_tc_kernel"), true); (void)reg;
  CUtensorMap tmA, tmB;
  const int M = cells.capacity * 4;
  int rc = make_tmap_bf16_2d(&tmA, fc, (uint64_t)M, (uint64_t)d.D, (uint64_t)d.D, UG_BM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, W, (uint64_t)CT_DL, (uint64_t)d.D, (uint64_t)d.D, CT_DL);
  if (rc) return rc;
  const int tiles = ceil_div(M, UG_BM);
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
#define VML_CT(NQP) return launch_content_tc<NQP>(tmA, tmB, grid, bias, qproj, ld, off_what, off_ktil, off_beta, s_hat, s_ld, qmask, cells, cc_hat, B, d, st)
  if (d.Nq + 1 <= 8) VML_CT(8);
  if (d.Nq + 1 <= 16) VML_CT(16);
  if (d.Nq + 1 <= 24) VML_CT(24);
  VML_CT(32);
#undef VML_CT
}

}  // namespace vml
