// tcgen05 / TMEM / TMA GEMM on fp32 operands read as TF32 (kind::tf32, fp32 accumulation in tensor memory):
//
//     C[m][n]  (= | += | atomic +=)  alpha * sum_k A(m, k) * B(n, k)
//
// for the large dense products of the TRAINING step (forward with saved activations and the whole backward, main.py:150
// `loss.backward()`), which round 1 ran on CUDA cores (gemm_strided.cuh).  Operands stay the caller's fp32 tensors -- no
// casts, no transposed copies: each operand may be
//   * K-major  (the contraction index is contiguous: activations [rows, K], weights [N, K]), or
//   * MN-major (the row / column index is contiguous: X and dY in dW = dY^T . X, W in dX = dY . W),
// the four combinations a linear layer's forward and backward need.  TMA moves 128-byte-wide boxes (32 fp32) with the
// 128B swizzle straight into the canonical UMMA layouts:
//   K-major  tile [128 rows x 32 k]   : one box, rows 128 B apart, 8-row groups 1024 B apart; an MMA (K = 8) advances 32 B
//   MN-major tile [32 k x 128 rows]   : four boxes of [32 k-rows x 32 mn] (4 KB each = LBO), k-rows 128 B apart; 32-bit types
//                                       read MN-major use the SWIZZLE_128B_BASE32B layout (32-byte pieces XOR row % 4, TMA
//                                       swizzle 128B_ATOM_32B): K groups of 4 rows 512 B (= SBO) apart; an MMA (K = 8)
//                                       advances two groups = 1024 B
// The tensor maps use CU_TENSOR_MAP_DATA_TYPE_TFLOAT32: the copy engine rounds fp32 to TF32 (the MMA would otherwise
// truncate the low 13 mantissa bits, which biases long sums).
// Split-K (dW: K = number of live rows, up to ~10^5) runs the K ranges on different CTAs and adds the partial tiles with
// fp32 atomics.  K and M may be device-resident (live cell count); for a device-limited K the (< 32) rows between the
// live count and the next multiple of 32 are zeroed in BOTH operands first (they are undefined by contract, the
// contraction must not see them).
// One CTA (192 threads) per (128 x 128 tile, K range): warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue
// (tcgen05.ld -> registers -> global).
#include <stdlib.h>

#include "common.cuh"
#include "gemm_umma.cuh"
#include "sm100.cuh"

namespace vml {

constexpr int TF_BM = 128, TF_BN = 128, TF_BK = 32, TF_STAGES = 6, TF_THREADS = 192;
constexpr int TF_OP_BYTES = TF_BM * TF_BK * 4;                 // 16 KB per operand and stage
constexpr int TF_SMEM = TF_STAGES * 2 * TF_OP_BYTES + 1024 + 256;

struct TfArgs {
  float* C; int64_t ldc;
  int M, N, K;
  float alpha;
  int mode;                  // 0: C = ., 1: C += . (plain read-modify-write), 2: atomic += (split-K)
  int splits;
  const int32_t* m_dev; int m_scale;
  const int32_t* k_dev; int k_scale;
  int mn_lbo, mn_sbo, mn_kstep, mn_ltype;      // MN-major descriptor parameters: 4096 / 512 / 1024 bytes, layout type 1 (debug knobs VML_TF_*)
};


template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(TF_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TfArgs g) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + TF_STAGES * 2 * TF_OP_BYTES);
  uint64_t* empty_bar = full_bar + TF_STAGES;
  uint64_t* tfull_bar = empty_bar + TF_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  int M = g.M, K = g.K;
  if (g.m_dev) M = min(M, *g.m_dev * g.m_scale);
  if (g.k_dev) K = min(K, *g.k_dev * g.k_scale);
  const int tiles_n = (g.N + TF_BN - 1) / TF_BN;
  const int m0 = ((int)blockIdx.x / tiles_n) * TF_BM, n0 = ((int)blockIdx.x % tiles_n) * TF_BN;
  // K range of this split, in whole 32-wide blocks
  const int kb_total = (K + TF_BK - 1) / TF_BK;
  const int kb_per = (kb_total + g.splits - 1) / g.splits;
  const int kb_lo = (int)blockIdx.y * kb_per, kb_hi = min(kb_total, kb_lo + kb_per);
  const bool idle = m0 >= M || (kb_lo >= kb_hi && !(g.mode == 0 && blockIdx.y == 0));   // nothing to add (mode 0 still writes zeros)
  if (idle) return;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < TF_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    ptx::mbar_init(tfull_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<TF_BN>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nkb = kb_hi - kb_lo;

  if (warp == 0) {
    if (lane == 0) {                                       // ===================== TMA producer =====================
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_lo; kb < kb_hi; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        unsigned char* sa = smem + stage * 2 * TF_OP_BYTES;
        unsigned char* sb = sa + TF_OP_BYTES;
        ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * TF_OP_BYTES);
        const int k0 = kb * TF_BK;
        if (A_MN) {
#pragma unroll
          for (int a = 0; a < TF_BM / 32; ++a) ptx::tma_load_2d(sa + a * 4096, &tmA, &full_bar[stage], m0 + 32 * a, k0);
        } else {
          ptx::tma_load_2d(sa, &tmA, &full_bar[stage], k0, m0);
        }
        if (B_MN) {
#pragma unroll
          for (int a = 0; a < TF_BN / 32; ++a) ptx::tma_load_2d(sb + a * 4096, &tmB, &full_bar[stage], n0 + 32 * a, k0);
        } else {
          ptx::tma_load_2d(sb, &tmB, &full_bar[stage], k0, n0);
        }
        if (++stage == TF_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                       // ===================== MMA issuer =====================
      constexpr uint32_t idesc = ptx::umma_idesc_tf32(TF_BM, TF_BN, A_MN, B_MN);
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(smem + stage * 2 * TF_OP_BYTES), b_addr = a_addr + TF_OP_BYTES;
#pragma unroll
        for (int k = 0; k < TF_BK / 8; ++k) {
          const uint64_t adesc = A_MN ? ptx::umma_desc_sw128_mn(a_addr + (uint32_t)(k * g.mn_kstep), g.mn_lbo, g.mn_sbo, g.mn_ltype)
                                      : ptx::umma_desc_sw128(a_addr) + (uint64_t)(k * 2);
          const uint64_t bdesc = B_MN ? ptx::umma_desc_sw128_mn(b_addr + (uint32_t)(k * g.mn_kstep), g.mn_lbo, g.mn_sbo, g.mn_ltype)
                                      : ptx::umma_desc_sw128(b_addr) + (uint64_t)(k * 2);
          ptx::umma_tf32(tmem_base, adesc, bdesc, idesc, (i | k) != 0);
        }
        ptx::umma_commit(&empty_bar[stage]);
        if (++stage == TF_STAGES) { stage = 0; phase ^= 1; }
      }
      ptx::umma_commit(tfull_bar);
    }
  } else {
    // ===================== epilogue warps: TMEM lane quadrant = warp % 4, one tile row per thread =====================
    const int quad = warp % 4;
    const int row = m0 + quad * 32 + lane;
    const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    if (nkb > 0) {
      ptx::mbar_wait(tfull_bar, 0);
      ptx::tc_fence_after();
    }
    float* crow = g.C + (int64_t)row * g.ldc;
    const bool vec = (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
#pragma unroll 1
    for (int c = 0; c < TF_BN; c += 32) {
      float v[32];
      if (nkb > 0) {
        ptx::tmem_ld32(t_addr + (uint32_t)c, v);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0.f;
      }
      if (row >= M) continue;
      const int col0 = n0 + c;
      if (col0 >= g.N) continue;
      if (vec && col0 + 32 <= g.N) {
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          float4 o = make_float4(g.alpha * v[e], g.alpha * v[e + 1], g.alpha * v[e + 2], g.alpha * v[e + 3]);
          float4* p = reinterpret_cast<float4*>(crow + col0 + e);
          if (g.mode == 2) atomicAdd(p, o);
          else {
            if (g.mode == 1) { const float4 old = *p; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
            *p = o;
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          if (col0 + e < g.N) {
            const float o = g.alpha * v[e];
            if (g.mode == 2) atomicAdd(crow + col0 + e, o);
            else if (g.mode == 1) crow[col0 + e] += o;
            else crow[col0 + e] = o;
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc<TF_BN>(tmem_base); }
}

// rows [live, roundup32(live)) of a [rows x cols] fp32 matrix (row stride ld) := 0
__global__ void zero_tail_rows_kernel(float* __restrict__ p, int64_t ld, int cols, int cap_rows, const int32_t* __restrict__ n_dev,
                                      int scale) {
  const int live = min(cap_rows, *n_dev * scale);
  const int hi = min(cap_rows, (live + 31) / 32 * 32);
  const int64_t total = (int64_t)(hi - live) * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    p[(int64_t)(live + e / cols) * ld + e % cols] = 0.f;
}

int make_tmap_2d(CUtensorMap* map, int dtype, const void* ptr, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer, int swizzle);

// Returns VML_OK when the product ran on the tensor cores, 1 when the shape is not eligible (caller falls back to the
// CUDA-core kernel), < 0 on error.
int launch_gemm_tf32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk, float* C, int64_t scm,
                     int64_t scn, int M, int N, int K, float alpha, int accumulate, const int32_t* m_dev, int m_scale,
                     const int32_t* k_dev, int k_scale, cudaStream_t st) {
  static const bool off = getenv("VML_NO_TF32") != nullptr;          // (A/B knob) force the CUDA-core path
  if (off) return 1;
  const bool a_k = sak == 1 && sam % 4 == 0, a_mn = sam == 1 && sak % 4 == 0 && !a_k;
  const bool b_k = sbk == 1 && sbn % 4 == 0, b_mn = sbn == 1 && sbk % 4 == 0 && !b_k;
  if (!(a_k || a_mn) || !(b_k || b_mn) || scn != 1) return 1;
  if (((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) != 0) return 1;
  if ((int64_t)M * N < 128 * 128 || (double)M * N * K < 1.0e8) return 1;        // small products: launch-bound either way
  if (m_dev && a_mn) return 1;                                                     // (no caller needs it)
  if (k_dev && !(a_mn && b_mn)) return 1;
  if (M <= 0 || K <= 0) return 1;
  static bool reg = (register_kernel("gemm_tf32_kernel"), true); (void)reg;
  CUtensorMap tmA, tmB;
  int rc;
  // K-major: dims {K, rows}, box {32, 128}, 128B swizzle;   MN-major: dims {rows, K}, box {32, 32}, 128B swizzle with 32-byte atoms
  auto knob = [](const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; };
  const int swz_mn = knob("VML_TF_SWZ", 4);                // CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
  if (a_k) rc = make_tmap_2d(&tmA, 1, A, (uint64_t)K, (uint64_t)M, (uint64_t)sam * 4, TF_BK, TF_BM, 3);
  else rc = make_tmap_2d(&tmA, 1, A, (uint64_t)M, (uint64_t)K, (uint64_t)sak * 4, 32, TF_BK, swz_mn);
  if (rc) return rc;
  if (b_k) rc = make_tmap_2d(&tmB, 1, B, (uint64_t)K, (uint64_t)N, (uint64_t)sbn * 4, TF_BK, TF_BN, 3);
  else rc = make_tmap_2d(&tmB, 1, B, (uint64_t)N, (uint64_t)K, (uint64_t)sbk * 4, 32, TF_BK, swz_mn);
  if (rc) return rc;
  const int tiles = ceil_div(M, TF_BM) * ceil_div(N, TF_BN);
  int splits = 1;
  if (tiles < kNumSMs) {                                   // few output tiles (dW): cut K so that the machine is filled
    splits = (2 * kNumSMs) / tiles;
    const int max_splits = ceil_div(K, 8 * TF_BK);         // >= 256 rows of K per split
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  int mode = accumulate ? 1 : 0;
  if (splits > 1) {
    if (!accumulate) VML_CUDA(cudaMemset2DAsync(C, (size_t)scm * 4, 0, (size_t)N * 4, (size_t)M, st));
    mode = 2;
  }
  if (k_dev) {                                             // undefined rows inside the last 32-row block must not be contracted
    zero_tail_rows_kernel<<<8, 256, 0, st>>>(const_cast<float*>(A), sak, M, K, k_dev, k_scale);
    zero_tail_rows_kernel<<<8, 256, 0, st>>>(const_cast<float*>(B), sbk, N, K, k_dev, k_scale);
  }
  TfArgs g{C, scm, M, N, K, alpha, mode, splits, m_dev, m_scale, k_dev, k_scale,
           knob("VML_TF_LBO", 4096), knob("VML_TF_SBO", 512), knob("VML_TF_KSTEP", 1024), knob("VML_TF_LTYPE", 1)};
  dim3 grid(tiles, splits);
#define VML_TF(AMN, BMN)                                                                                              \
  do {                                                                                                                \
    VML_CUDA(ensure_dyn_smem((const void*)(gemm_tf32_kernel<AMN, BMN>), (size_t)TF_SMEM));                             \
    gemm_tf32_kernel<AMN, BMN><<<grid, TF_THREADS, TF_SMEM, st>>>(tmA, tmB, g);                                        \
  } while (0)
  if (a_mn && b_mn) VML_TF(true, true);
  else if (a_mn) VML_TF(true, false);
  else if (b_mn) VML_TF(false, true);
  else VML_TF(false, false);
#undef VML_TF
  VML_LAUNCHED(k_dev ? 3 : 1);
  return VML_OK;
}

}  // namespace vml
