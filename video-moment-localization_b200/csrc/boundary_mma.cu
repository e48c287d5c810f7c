// a7  BoundaryUnit  (Attention.forward models.py:137-154; BoundaryUnit.forward :164-196)
//
// Per sample the unit is four small contractions around two softmaxes,
//     scores = fb.kbt^T (+beta_b)      [L x Nq x D]     Aq   = softmax(scores).fw      [L x D x Nq]
//     S      = G.G^T / sqrt(D)         [L x L  x D]     f_bb = softmax(S).fb           [L x D x L ]
// with G = fb * (Aq*lmask + fs), plus one streaming pass over the sample's map cells (boundary_stream_kernel,
// one CTA per map row, HBM-bound)
//     f_bm[i] = sum_j A_b[i,j] * sigmoid(fm_ij*fs)*fm_ij     (also written out per cell as `fbar`).
// The contractions run as warp-level mma.sync m16n8k8 TF32 with fp32 accumulation: one pass in the
// fast mode, the 3xTF32 split (a_hi.b_hi + a_hi.b_lo + a_lo.b_hi, ~2^-21 relative) in the fp32
// validation mode, so both modes share this code.  One CTA (8 warps) per (sample, 16 map rows):
// contractions over D are split across the warps along K and reduced through shared memory,
// contractions producing D columns are split across the warps along N.
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

namespace vml {

constexpr int BMM_ROWS = 16, BMM_WARPS = 8, BMM_THREADS = BMM_WARPS * 32;
constexpr int BMM_MAXQ = 32;      // word slots (Nq <= 32)
constexpr int BMM_MAXD64 = 8;     // D <= 512: D/64 column tiles per warp

// fp32 -> tf32, round to nearest, ties away from zero (= cvt.rna.tf32.f32 for every finite input that does not round up to
// infinity): half an ulp of the 10-bit mantissa is added to the magnitude and the 13 dropped bits are cleared.  ptxas expands
// cvt.rna.tf32 on sm_100a into a ~5-instruction compare/select sequence; with six conversions per mma.m16n8k8 that sequence was
// a quarter of all instructions the gate / rows kernels issued (ncu source page, profiles/r02_boundary_gate_rows_hotspots.txt).
__device__ __forceinline__ uint32_t f2tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// d[16x8] += a[16x8] . b[8x8]; fragment layout of mma.m16n8k8 (g = lane/4, t = lane%4):
//   a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4);  b0 (k=t, n=g)  b1 (k=t+4, n=g);  d0,d1 (g, 2t..2t+1)  d2,d3 (g+8, ..)
// Fast mode (PRECISE = false): the fp32 bits go to the tensor core as they are -- mma.tf32 reads the upper 19 bits of each
// operand register, i.e. the operands are TRUNCATED to tf32 (<= 2^-10 relative, against 2^-11 with round-to-nearest; the
// unit's outputs are compared at 2e-2).  The six roundings per mma were 12 of the ~25 instructions per mma of these
// latency-bound kernels (4 resident warps per scheduler, ~12 cycles per issued instruction).  VML_TF32_RNA restores them.
template <bool PRECISE>
__device__ __forceinline__ void mma_16x8x8(float (&d)[4], const float (&a)[4], const float (&b)[2]) {
  uint32_t ah[4], bh[2];
#if defined(VML_TF32_RNA)
  constexpr bool kRound = true;
#else
  constexpr bool kRound = PRECISE;
#endif
#pragma unroll
  for (int e = 0; e < 4; ++e) ah[e] = kRound ? f2tf32(a[e]) : __float_as_uint(a[e]);
#pragma unroll
  for (int e = 0; e < 2; ++e) bh[e] = kRound ? f2tf32(b[e]) : __float_as_uint(b[e]);
  if (PRECISE) {
    uint32_t al[4], bl[2];
#pragma unroll
    for (int e = 0; e < 4; ++e) al[e] = f2tf32(a[e] - __uint_as_float(ah[e]));
#pragma unroll
    for (int e = 0; e < 2; ++e) bl[e] = f2tf32(b[e] - __uint_as_float(bh[e]));
    mma_tf32(d, al, bh);
    mma_tf32(d, ah, bl);
  }
  mma_tf32(d, ah, bh);
}

// ---- gate:  G[b,l,:] = fb * (softmax(fb.kbt^T + beta_b).fw * lmask + fs) ------------------------------
// The sample's W_q-folded keys and word states are staged once per CTA in shared memory (rows padded
// by 4 floats: the mma fragment loads are then bank-conflict free).
template <bool PRECISE>
__global__ void __launch_bounds__(BMM_THREADS)
boundary_gate_mma_kernel(const float* __restrict__ qproj, int ld, int off_kbt, int off_betab, const float* __restrict__ fw,
                         const float* __restrict__ fs, const float* __restrict__ fb, const uint8_t* __restrict__ qmask,
                         const uint8_t* __restrict__ lmask, float* __restrict__ G, float* __restrict__ prob_out,
                         float* __restrict__ u_out, int L, int Nq, int D) {
  extern __shared__ __align__(16) float sg[];
  const int DS = D + 4;
  float* Ks = sg;                                     // [Nq][DS]  kbt
  float* Ws = Ks + (size_t)Nq * DS;                   // [Nq][DS]  fw
  float* Xs = Ws + (size_t)Nq * DS;                   // [16][DS]  this tile's fb rows
  __shared__ float part[BMM_WARPS][BMM_ROWS][BMM_MAXQ + 1];
  __shared__ float prob[BMM_ROWS][BMM_MAXQ + 4];
  const int b = blockIdx.y, i0 = blockIdx.x * BMM_ROWS;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, g = lane >> 2, t = lane & 3;
  const int dq = D / 4;
  for (int e = tid; e < Nq * dq; e += BMM_THREADS) {
    const int k = e / dq, c4 = (e - k * dq) * 4;
    *reinterpret_cast<float4*>(Ks + (size_t)k * DS + c4) = __ldg(reinterpret_cast<const float4*>(qproj + ((size_t)b * Nq + k) * ld + off_kbt + c4));
    *reinterpret_cast<float4*>(Ws + (size_t)k * DS + c4) = __ldg(reinterpret_cast<const float4*>(fw + ((size_t)b * Nq + k) * D + c4));
  }
  for (int e = tid; e < BMM_ROWS * dq; e += BMM_THREADS) {
    const int rr = e / dq, c4 = (e - rr * dq) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i0 + rr < L) v = __ldg(reinterpret_cast<const float4*>(fb + ((size_t)b * L + i0 + rr) * D + c4));
    *reinterpret_cast<float4*>(Xs + (size_t)rr * DS + c4) = v;
  }
  __syncthreads();
  const int rA = i0 + g, rB = i0 + g + 8;
  const bool vA = rA < L, vB = rB < L;
  const int nq_tiles = (Nq + 7) / 8;
  const int dpw = D / BMM_WARPS;                       // D columns (or K range) owned by this warp
  // ---- scores, split over K --------------------------------------------------------------------------
  {
    float acc[BMM_MAXQ / 8][4];
#pragma unroll
    for (int n = 0; n < BMM_MAXQ / 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
    for (int k0 = warp * dpw; k0 < (warp + 1) * dpw; k0 += 8) {
      const float a[4] = {Xs[g * DS + k0 + t], Xs[(g + 8) * DS + k0 + t], Xs[g * DS + k0 + t + 4], Xs[(g + 8) * DS + k0 + t + 4]};
#pragma unroll
      for (int n = 0; n < BMM_MAXQ / 8; ++n) {
        if (n < nq_tiles) {
          const int w = n * 8 + g;
          float bf[2] = {0.f, 0.f};
          if (w < Nq) { bf[0] = Ks[(size_t)w * DS + k0 + t]; bf[1] = Ks[(size_t)w * DS + k0 + t + 4]; }
          mma_16x8x8<PRECISE>(acc[n], a, bf);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < BMM_MAXQ / 8; ++n) {
      part[warp][g][n * 8 + 2 * t] = acc[n][0]; part[warp][g][n * 8 + 2 * t + 1] = acc[n][1];
      part[warp][g + 8][n * 8 + 2 * t] = acc[n][2]; part[warp][g + 8][n * 8 + 2 * t + 1] = acc[n][3];
    }
  }
  __syncthreads();
  // ---- masked softmax over the words (models.py:143-150); warp w owns rows 2w, 2w+1, lane = word ---------
  {
    const float mk = (lane < Nq && qmask[(size_t)b * Nq + lane]) ? 1.f : 0.f;
    const float beta = lane < Nq ? qproj[((size_t)b * Nq + lane) * ld + off_betab] : 0.f;
    const float sqrt_d = sqrtf((float)D);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int row = 2 * warp + rr;
      float s = 0.f;
#pragma unroll
      for (int p = 0; p < BMM_WARPS; ++p) s += part[p][row][lane];
      float sv = (s + beta) / sqrt_d;
      sv = sv * mk;
      if (mk == 0.f) sv = -1e9f;                       // masked_fill(mask == 0, -1e9)
      if (lane >= Nq) sv = -INFINITY;                  // not a word at all
      const float mx = warp_max(sv);
      const float ex = lane < Nq ? expf(sv - mx) : 0.f;
      prob[row][lane] = ex / warp_sum(ex);
      if (prob_out && lane < Nq && i0 + row < L) prob_out[((size_t)b * L + i0 + row) * Nq + lane] = prob[row][lane];   // saved for backward
    }
  }
  __syncthreads();
  // ---- attended words for this warp's D/8 columns, gate, store G ------------------------------------------
  {
    float acc[BMM_MAXD64][4];
#pragma unroll
    for (int n = 0; n < BMM_MAXD64; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
    const int ntd = dpw / 8;
    for (int k0 = 0; k0 < nq_tiles * 8; k0 += 8) {
      const float a[4] = {prob[g][k0 + t], prob[g + 8][k0 + t], prob[g][k0 + t + 4], prob[g + 8][k0 + t + 4]};
#pragma unroll
      for (int n = 0; n < BMM_MAXD64; ++n) {
        if (n < ntd) {
          const int col = warp * dpw + n * 8 + g;
          float bf[2];
          bf[0] = (k0 + t < Nq) ? Ws[(size_t)(k0 + t) * DS + col] : 0.f;
          bf[1] = (k0 + t + 4 < Nq) ? Ws[(size_t)(k0 + t + 4) * DS + col] : 0.f;
          mma_16x8x8<PRECISE>(acc[n], a, bf);
        }
      }
    }
    const float lmA = (vA && lmask[(size_t)b * L + rA]) ? 1.f : 0.f, lmB = (vB && lmask[(size_t)b * L + rB]) ? 1.f : 0.f;
#pragma unroll
    for (int n = 0; n < BMM_MAXD64; ++n) {
      if (n < ntd) {
        const int col = warp * dpw + n * 8 + 2 * t;
        const float2 s2 = __ldg(reinterpret_cast<const float2*>(fs + (size_t)b * D + col));
        if (vA) {
          const float2 x = *reinterpret_cast<const float2*>(Xs + g * DS + col);
          *reinterpret_cast<float2*>(G + ((size_t)b * L + rA) * D + col) =
              make_float2(x.x * (acc[n][0] * lmA + s2.x), x.y * (acc[n][1] * lmA + s2.y));
          if (u_out) *reinterpret_cast<float2*>(u_out + ((size_t)b * L + rA) * D + col) = make_float2(acc[n][0] * lmA + s2.x, acc[n][1] * lmA + s2.y);
        }
        if (vB) {
          const float2 x = *reinterpret_cast<const float2*>(Xs + (g + 8) * DS + col);
          *reinterpret_cast<float2*>(G + ((size_t)b * L + rB) * D + col) =
              make_float2(x.x * (acc[n][2] * lmB + s2.x), x.y * (acc[n][3] * lmB + s2.y));
          if (u_out) *reinterpret_cast<float2*>(u_out + ((size_t)b * L + rB) * D + col) = make_float2(acc[n][2] * lmB + s2.x, acc[n][3] * lmB + s2.y);
        }
      }
    }
  }
}

// ---- rows:  A_b = softmax_j(G_i.G_j / sqrt(D)) (masked);  bu[i] = A_b[i,:].fb + fb[i] + sum_j A_b[i,j] fbar_ij ----
// Keys / values (G, then fb) of the sample are staged in shared memory in blocks of JR rows.  Writes the
// attention rows A_b and bu = f_bb + f_b; boundary_stream_kernel then adds f_bm.
constexpr int BMM_JB = 64;

template <bool PRECISE>
__global__ void __launch_bounds__(BMM_THREADS)
boundary_rows_mma_kernel(const float* __restrict__ G, const float* __restrict__ fb, const uint8_t* __restrict__ lmask,
                         float* __restrict__ bu, float* __restrict__ ab_out, int L, int D,
                         int JR /* rows per staged block: multiple of 8, <= BMM_JB */) {
  extern __shared__ __align__(16) float sm_all[];
  const int LP = (L + 7) & ~7;                         // keys padded to a multiple of 8
  const int LS = LP + 1, DS = D + 4;
  float* Js = sm_all;                                  // [JR][DS]      staged G / fb rows
  float* part = Js + (size_t)JR * DS;                  // [8][16][LS]   split-K partial scores
  float* Ab = part + BMM_WARPS * BMM_ROWS * LS;        // [16][LS]      attention rows
  float* outT = Ab + BMM_ROWS * LS;                    // [16][DS]      tile rows: G_i, then f_bb + f_b
  const int b = blockIdx.y, i0 = blockIdx.x * BMM_ROWS;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, g = lane >> 2, t = lane & 3;
  const int rA = i0 + g, rB = i0 + g + 8;
  const bool vA = rA < L, vB = rB < L;
  const int dpw = D / BMM_WARPS, dq = D / 4;
  const float* Gb = G + (size_t)b * L * D;
  const float* fbb = fb + (size_t)b * L * D;
  auto stage = [&](const float* src, int j0) {         // rows [j0, j0 + JR) of src -> Js (zero past L)
    for (int e = tid; e < JR * dq; e += BMM_THREADS) {
      const int rr = e / dq, c4 = (e - rr * dq) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + rr < L) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(j0 + rr) * D + c4));
      *reinterpret_cast<float4*>(Js + (size_t)rr * DS + c4) = v;
    }
  };
  for (int e = tid; e < BMM_ROWS * dq; e += BMM_THREADS) {   // the tile's own G rows
    const int rr = e / dq, c4 = (e - rr * dq) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i0 + rr < L) v = __ldg(reinterpret_cast<const float4*>(Gb + (size_t)(i0 + rr) * D + c4));
    *reinterpret_cast<float4*>(outT + (size_t)rr * DS + c4) = v;
  }
  // ---- S = G_tile . G^T, split over K, one block of keys at a time --------------------------------------------
  for (int j0 = 0; j0 < LP; j0 += JR) {
    __syncthreads();
    stage(Gb, j0);
    __syncthreads();
    const int ntl = min(JR, LP - j0) / 8;
    float acc[BMM_JB / 8][4];
#pragma unroll
    for (int n = 0; n < BMM_JB / 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
    for (int k0 = warp * dpw; k0 < (warp + 1) * dpw; k0 += 8) {
      const float a[4] = {outT[g * DS + k0 + t], outT[(g + 8) * DS + k0 + t], outT[g * DS + k0 + t + 4], outT[(g + 8) * DS + k0 + t + 4]};
#pragma unroll
      for (int n = 0; n < BMM_JB / 8; ++n) {
        if (n < ntl) {
          const float bf[2] = {Js[(size_t)(n * 8 + g) * DS + k0 + t], Js[(size_t)(n * 8 + g) * DS + k0 + t + 4]};
          mma_16x8x8<PRECISE>(acc[n], a, bf);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < BMM_JB / 8; ++n) {
      if (n < ntl) {
        float* p0 = part + ((size_t)warp * BMM_ROWS + g) * LS + j0 + n * 8 + 2 * t;
        float* p1 = part + ((size_t)warp * BMM_ROWS + g + 8) * LS + j0 + n * 8 + 2 * t;
        p0[0] = acc[n][0]; p0[1] = acc[n][1]; p1[0] = acc[n][2]; p1[1] = acc[n][3];
      }
    }
  }
  __syncthreads();
  // ---- masked softmax over the keys (models.py:176-184); warp w owns rows 2w, 2w+1 ------------------------------
  {
    const float sqrt_d = sqrtf((float)D);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int row = 2 * warp + rr, i = i0 + row;
      const bool row_on = i < L && lmask[(size_t)b * L + i] != 0;
      float* arow = Ab + row * LS;
      float mx = -INFINITY;
      for (int j = lane; j < LP; j += 32) {
        float s = 0.f;
#pragma unroll
        for (int p = 0; p < BMM_WARPS; ++p) s += part[((size_t)p * BMM_ROWS + row) * LS + j];
        float sc = -INFINITY;
        if (j < L) {
          const float mk = lmask[(size_t)b * L + j] ? 1.f : 0.f;
          sc = (s / sqrt_d) * mk;
          if (mk == 0.f) sc = -1e9f;
        }
        arow[j] = sc;
        mx = fmaxf(mx, sc);
      }
      mx = warp_max(mx);
      float den = 0.f;
      for (int j = lane; j < LP; j += 32) { const float ex = j < L ? expf(arow[j] - mx) : 0.f; arow[j] = ex; den += ex; }
      den = warp_sum(den);
      for (int j = lane; j < LP; j += 32) arow[j] = row_on ? arow[j] / den : 0.f;
    }
  }
  // ---- f_bb for this warp's D/8 columns, one block of values at a time ---------------------------------------------
  {
    float acc[BMM_MAXD64][4];
#pragma unroll
    for (int n = 0; n < BMM_MAXD64; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
    const int ntd = dpw / 8;
    for (int j0 = 0; j0 < LP; j0 += JR) {
      __syncthreads();
      stage(fbb, j0);
      __syncthreads();
      const int kn = min(JR, LP - j0);
      for (int k0 = 0; k0 < kn; k0 += 8) {
        const float a[4] = {Ab[g * LS + j0 + k0 + t], Ab[(g + 8) * LS + j0 + k0 + t], Ab[g * LS + j0 + k0 + t + 4],
                            Ab[(g + 8) * LS + j0 + k0 + t + 4]};
#pragma unroll
        for (int n = 0; n < BMM_MAXD64; ++n) {
          if (n < ntd) {
            const int col = warp * dpw + n * 8 + g;
            const float bf[2] = {Js[(size_t)(k0 + t) * DS + col], Js[(size_t)(k0 + t + 4) * DS + col]};
            mma_16x8x8<PRECISE>(acc[n], a, bf);
          }
        }
      }
    }
    __syncthreads();                                   // every warp is done reading the tile's G rows in outT
#pragma unroll
    for (int n = 0; n < BMM_MAXD64; ++n) {
      if (n < ntd) {
        const int col = warp * dpw + n * 8 + 2 * t;
        float2 xa = make_float2(0.f, 0.f), xb = xa;
        if (vA) xa = __ldg(reinterpret_cast<const float2*>(fbb + (size_t)rA * D + col));
        if (vB) xb = __ldg(reinterpret_cast<const float2*>(fbb + (size_t)rB * D + col));
        outT[g * DS + col] = acc[n][0] + xa.x; outT[g * DS + col + 1] = acc[n][1] + xa.y;           // (f_bb + f_b)
        outT[(g + 8) * DS + col] = acc[n][2] + xb.x; outT[(g + 8) * DS + col + 1] = acc[n][3] + xb.y;
      }
    }
  }
  __syncthreads();
  // ---- hand over: attention rows and (f_bb + f_b); the streaming kernel adds f_bm ------------------------------------
  for (int e = tid; e < BMM_ROWS * LP; e += BMM_THREADS) {
    const int rr = e / LP, j = e - rr * LP;
    if (i0 + rr < L && j < L) ab_out[((size_t)b * L + i0 + rr) * L + j] = Ab[rr * LS + j];
  }
  for (int e = tid; e < BMM_ROWS * dq; e += BMM_THREADS) {
    const int rr = e / dq, c4 = (e - rr * dq) * 4;
    if (i0 + rr < L)
      *reinterpret_cast<float4*>(bu + ((size_t)b * L + i0 + rr) * D + c4) = *reinterpret_cast<const float4*>(outT + (size_t)rr * DS + c4);
  }
}

// ---- gate + rows in ONE kernel for maps with L <= 16 (one 16-row block per sample, e.g. Charades) ---------------------
// Same arithmetic, in the same order, as boundary_gate_mma_kernel followed by boundary_rows_mma_kernel (JR = 16): the
// gated rows G stay in shared memory (over the dead key / word tiles) instead of a round trip through global memory,
// and one launch + one staging pass disappear.  One CTA per sample.
// (Measured dead end, round 2: 16 warps per sample with the word states fetched straight into MMA fragments -- 32 resident
// warps per SM instead of 16, a third less shared memory -- runs in the same 54 us per 640 samples as this version; so does
// the version before the cheaper tf32 rounding below at +25 % instructions.  The kernel's time is the chain of eight
// barrier-separated phases per sample times 2.2 waves of CTAs, not issue slots or bytes.)
template <bool PRECISE, int DT /* D at compile time (loops unroll, shared-memory offsets become immediates); 0 = run time */>
__global__ void __launch_bounds__(BMM_THREADS)
boundary_gate_rows_kernel(const float* __restrict__ qproj, int ld, int off_kbt, int off_betab, const float* __restrict__ fw,
                          const float* __restrict__ fs, const float* __restrict__ fb, const uint8_t* __restrict__ qmask,
                          const uint8_t* __restrict__ lmask, float* __restrict__ G, float* __restrict__ prob_out,
                          float* __restrict__ u_out, float* __restrict__ bu, float* __restrict__ ab_out, int L, int Nq, int D_rt,
                          int bulk_stage) {
  const int D = DT > 0 ? DT : D_rt;
  extern __shared__ __align__(16) float sg[];
  const int DS = D + 4;
  float* Ks = sg;                                     // [Nq][DS]  kbt          (later: G rows, [16][DS])
  float* Ws = Ks + (size_t)Nq * DS;                   // [Nq][DS]  fw
  float* Xs = sg + (size_t)max(2 * Nq, BMM_ROWS) * DS;   // [16][DS]  the sample's fb rows (later: f_bb + f_b)
  float* Gs = sg;
  __shared__ float part[BMM_WARPS][BMM_ROWS][BMM_MAXQ + 1];
  __shared__ float prob[BMM_ROWS][BMM_MAXQ + 4];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, g = lane >> 2, t = lane & 3;
  const int dq = D / 4;
  // staging, default: the copy engine moves the sample's rows (one cp.async.bulk per 2 KB row into the padded layout, issued
  // by warp 0, counted on one mbarrier).  Through registers the staging was ~15 k of the 36 k warp instructions a sample
  // costs (5.4 k LDG.128 + 5.4 k STS.128 + addressing; ncu source page, profiles/r02_boundary_gate_rows_hotspots.txt) in a
  // kernel running at IPC ~2 of 4.
  __shared__ __align__(8) uint64_t stage_bar;
  if (bulk_stage) {
    if (tid == 0) { ptx::mbar_init(&stage_bar, 1); ptx::fence_barrier_init(); }
    __syncthreads();
    const int nx = min(L, BMM_ROWS);
    if (warp == 0) {
      if (lane == 0) ptx::mbar_arrive_expect_tx(&stage_bar, (uint32_t)((2 * Nq + nx) * D) * 4u);
      __syncwarp();
      const uint32_t rb = (uint32_t)D * 4u;
      for (int r = lane; r < 2 * Nq + nx; r += 32) {
        if (r < Nq) ptx::bulk_load_1d(Ks + (size_t)r * DS, qproj + ((size_t)b * Nq + r) * ld + off_kbt, rb, &stage_bar);
        else if (r < 2 * Nq) ptx::bulk_load_1d(Ws + (size_t)(r - Nq) * DS, fw + ((size_t)b * Nq + (r - Nq)) * D, rb, &stage_bar);
        else ptx::bulk_load_1d(Xs + (size_t)(r - 2 * Nq) * DS, fb + ((size_t)b * L + (r - 2 * Nq)) * D, rb, &stage_bar);
      }
    } else {
      for (int e = tid - 32; e < (BMM_ROWS - nx) * dq; e += BMM_THREADS - 32) {      // map rows past L: zero
        const int rr = nx + e / dq, c4 = (e % dq) * 4;
        *reinterpret_cast<float4*>(Xs + (size_t)rr * DS + c4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    ptx::mbar_wait(&stage_bar, 0);
  } else
  // staging through registers: 12 independent 16-byte loads in flight per thread before the first shared-memory store (the unit is
  // latency-bound: one load per iteration meant ~20 dependent trips to L2 / HBM per CTA).  A thread keeps ONE column chunk
  // and walks rows (BMM_THREADS / dq rows per sweep): no per-element division by the run-time row length.
  if (BMM_THREADS % dq == 0) {
    constexpr int UB = 4;
    const int rstep = BMM_THREADS / dq, r0 = tid / dq, c4 = (tid - r0 * dq) * 4;
    const int nrow = max(Nq, BMM_ROWS);
    const float* kp = qproj + (size_t)b * Nq * ld + off_kbt + c4;
    const float* wp = fw + (size_t)b * Nq * D + c4;
    const float* xp = fb + (size_t)b * L * D + c4;
    for (int k0 = r0; k0 < nrow; k0 += UB * rstep) {
      float4 kv[UB], wv[UB], xv[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int k = k0 + u * rstep;
        kv[u] = wv[u] = xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < Nq) {
          kv[u] = __ldg(reinterpret_cast<const float4*>(kp + (size_t)k * ld));
          wv[u] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)k * D));
        }
        if (k < L) xv[u] = __ldg(reinterpret_cast<const float4*>(xp + (size_t)k * D));
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int k = k0 + u * rstep;
        if (k < Nq) {
          *reinterpret_cast<float4*>(Ks + (size_t)k * DS + c4) = kv[u];
          *reinterpret_cast<float4*>(Ws + (size_t)k * DS + c4) = wv[u];
        }
        if (k < BMM_ROWS) *reinterpret_cast<float4*>(Xs + (size_t)k * DS + c4) = xv[u];
      }
    }
  } else {
    constexpr int UB = 4;
    const int nkw = Nq * dq, nx = BMM_ROWS * dq;
    for (int e0 = tid; e0 < max(nkw, nx); e0 += UB * BMM_THREADS) {
      float4 kv[UB], wv[UB], xv[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int e = e0 + u * BMM_THREADS;
        kv[u] = wv[u] = xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < nkw) {
          const int k = e / dq, c4 = (e - k * dq) * 4;
          kv[u] = __ldg(reinterpret_cast<const float4*>(qproj + ((size_t)b * Nq + k) * ld + off_kbt + c4));
          wv[u] = __ldg(reinterpret_cast<const float4*>(fw + ((size_t)b * Nq + k) * D + c4));
        }
        if (e < nx) {
          const int rr = e / dq, c4 = (e - rr * dq) * 4;
          if (rr < L) xv[u] = __ldg(reinterpret_cast<const float4*>(fb + ((size_t)b * L + rr) * D + c4));
        }
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int e = e0 + u * BMM_THREADS;
        if (e < nkw) {
          const int k = e / dq, c4 = (e - k * dq) * 4;
          *reinterpret_cast<float4*>(Ks + (size_t)k * DS + c4) = kv[u];
          *reinterpret_cast<float4*>(Ws + (size_t)k * DS + c4) = wv[u];
        }
        if (e < nx) {
          const int rr = e / dq, c4 = (e - rr * dq) * 4;
          *reinterpret_cast<float4*>(Xs + (size_t)rr * DS + c4) = xv[u];
        }
      }
    }
  }
  __syncthreads();
  const int rA = g, rB = g + 8;
  const bool vA = rA < L, vB = rB < L;
  const int nq_tiles = (Nq + 7) / 8;
  const int dpw = D / BMM_WARPS;                       // D columns (or K range) owned by this warp
  const int ntd = dpw / 8;
  // ---- word scores, split over K ---------------------------------------------------------------------
  {
    float acc[BMM_MAXQ / 8][4];
#pragma unroll
    for (int n = 0; n < BMM_MAXQ / 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
    for (int k0 = warp * dpw; k0 < (warp + 1) * dpw; k0 += 8) {
      const float a[4] = {Xs[g * DS + k0 + t], Xs[(g + 8) * DS + k0 + t], Xs[g * DS + k0 + t + 4], Xs[(g + 8) * DS + k0 + t + 4]};
#pragma unroll
      for (int n = 0; n < BMM_MAXQ / 8; ++n) {
        if (n < nq_tiles) {
          const int w = n * 8 + g;
          float bf[2] = {0.f, 0.f};
          if (w < Nq) { bf[0] = Ks[(size_t)w * DS + k0 + t]; bf[1] = Ks[(size_t)w * DS + k0 + t + 4]; }
          mma_16x8x8<PRECISE>(acc[n], a, bf);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < BMM_MAXQ / 8; ++n) {
      part[warp][g][n * 8 + 2 * t] = acc[n][0]; part[warp][g][n * 8 + 2 * t + 1] = acc[n][1];
      part[warp][g + 8][n * 8 + 2 * t] = acc[n][2]; part[warp][g + 8][n * 8 + 2 * t + 1] = acc[n][3];
    }
  }
  __syncthreads();
  // ---- masked softmax over the words (models.py:143-150); warp w owns rows 2w, 2w+1, lane = word ---------
  {
    const float mk = (lane < Nq && qmask[(size_t)b * Nq + lane]) ? 1.f : 0.f;
    const float beta = lane < Nq ? qproj[((size_t)b * Nq + lane) * ld + off_betab] : 0.f;
    const float sqrt_d = sqrtf((float)D);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int row = 2 * warp + rr;
      float s = 0.f;
#pragma unroll
      for (int p = 0; p < BMM_WARPS; ++p) s += part[p][row][lane];
      float sv = (s + beta) / sqrt_d;
      sv = sv * mk;
      if (mk == 0.f) sv = -1e9f;                       // masked_fill(mask == 0, -1e9)
      if (lane >= Nq) sv = -INFINITY;                  // not a word at all
      const float mx = warp_max(sv);
      const float ex = lane < Nq ? expf(sv - mx) : 0.f;
      prob[row][lane] = ex / warp_sum(ex);
      if (prob_out && lane < Nq && row < L) prob_out[((size_t)b * L + row) * Nq + lane] = prob[row][lane];   // saved for backward
    }
  }
  __syncthreads();
  // ---- attended words for this warp's D/8 columns, gate; G to global now, to shared memory once Ws is dead ----
  float gA[BMM_MAXD64][2], gB[BMM_MAXD64][2];
  {
    float acc[BMM_MAXD64][4];
#pragma unroll
    for (int n = 0; n < BMM_MAXD64; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
    for (int k0 = 0; k0 < nq_tiles * 8; k0 += 8) {
      const float a[4] = {prob[g][k0 + t], prob[g + 8][k0 + t], prob[g][k0 + t + 4], prob[g + 8][k0 + t + 4]};
#pragma unroll
      for (int n = 0; n < BMM_MAXD64; ++n) {
        if (n < ntd) {
          const int col = warp * dpw + n * 8 + g;
          float bf[2];
          bf[0] = (k0 + t < Nq) ? Ws[(size_t)(k0 + t) * DS + col] : 0.f;
          bf[1] = (k0 + t + 4 < Nq) ? Ws[(size_t)(k0 + t + 4) * DS + col] : 0.f;
          mma_16x8x8<PRECISE>(acc[n], a, bf);
        }
      }
    }
    const float lmA = (vA && lmask[(size_t)b * L + rA]) ? 1.f : 0.f, lmB = (vB && lmask[(size_t)b * L + rB]) ? 1.f : 0.f;
#pragma unroll
    for (int n = 0; n < BMM_MAXD64; ++n) {
      gA[n][0] = gA[n][1] = gB[n][0] = gB[n][1] = 0.f;
      if (n < ntd) {
        const int col = warp * dpw + n * 8 + 2 * t;
        const float2 s2 = __ldg(reinterpret_cast<const float2*>(fs + (size_t)b * D + col));
        if (vA) {
          const float2 x = *reinterpret_cast<const float2*>(Xs + g * DS + col);
          gA[n][0] = x.x * (acc[n][0] * lmA + s2.x); gA[n][1] = x.y * (acc[n][1] * lmA + s2.y);
          *reinterpret_cast<float2*>(G + ((size_t)b * L + rA) * D + col) = make_float2(gA[n][0], gA[n][1]);
          if (u_out) *reinterpret_cast<float2*>(u_out + ((size_t)b * L + rA) * D + col) = make_float2(acc[n][0] * lmA + s2.x, acc[n][1] * lmA + s2.y);
        }
        if (vB) {
          const float2 x = *reinterpret_cast<const float2*>(Xs + (g + 8) * DS + col);
          gB[n][0] = x.x * (acc[n][2] * lmB + s2.x); gB[n][1] = x.y * (acc[n][3] * lmB + s2.y);
          *reinterpret_cast<float2*>(G + ((size_t)b * L + rB) * D + col) = make_float2(gB[n][0], gB[n][1]);
          if (u_out) *reinterpret_cast<float2*>(u_out + ((size_t)b * L + rB) * D + col) = make_float2(acc[n][2] * lmB + s2.x, acc[n][3] * lmB + s2.y);
        }
      }
    }
  }
  __syncthreads();                                     // every warp is done with Ks / Ws: G rows take their place
#pragma unroll
  for (int n = 0; n < BMM_MAXD64; ++n) {
    if (n < ntd) {
      const int col = warp * dpw + n * 8 + 2 * t;
      *reinterpret_cast<float2*>(Gs + g * DS + col) = make_float2(gA[n][0], gA[n][1]);
      *reinterpret_cast<float2*>(Gs + (g + 8) * DS + col) = make_float2(gB[n][0], gB[n][1]);
    }
  }
  __syncthreads();
  // ================= rows part (boundary_rows_mma_kernel with a single key block) =================
  const int LP = (L + 7) & ~7;                         // 8 or 16
  const int ntl = LP / 8;
  {
    float acc[2][4];
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
    for (int k0 = warp * dpw; k0 < (warp + 1) * dpw; k0 += 8) {
      const float a[4] = {Gs[g * DS + k0 + t], Gs[(g + 8) * DS + k0 + t], Gs[g * DS + k0 + t + 4], Gs[(g + 8) * DS + k0 + t + 4]};
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        if (n < ntl) {
          const float bf[2] = {Gs[(size_t)(n * 8 + g) * DS + k0 + t], Gs[(size_t)(n * 8 + g) * DS + k0 + t + 4]};
          mma_16x8x8<PRECISE>(acc[n], a, bf);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      if (n < ntl) {
        part[warp][g][n * 8 + 2 * t] = acc[n][0]; part[warp][g][n * 8 + 2 * t + 1] = acc[n][1];
        part[warp][g + 8][n * 8 + 2 * t] = acc[n][2]; part[warp][g + 8][n * 8 + 2 * t + 1] = acc[n][3];
      }
    }
  }
  __syncthreads();
  // ---- masked softmax over the keys (models.py:176-184); warp w owns rows 2w, 2w+1; prob[][] now holds A_b -------
  {
    const float sqrt_d = sqrtf((float)D);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int row = 2 * warp + rr;
      const bool row_on = row < L && lmask[(size_t)b * L + row] != 0;
      float sc = -INFINITY;
      if (lane < LP) {
        float s = 0.f;
#pragma unroll
        for (int p = 0; p < BMM_WARPS; ++p) s += part[p][row][lane];
        if (lane < L) {
          const float mk = lmask[(size_t)b * L + lane] ? 1.f : 0.f;
          sc = (s / sqrt_d) * mk;
          if (mk == 0.f) sc = -1e9f;
        }
      }
      const float mx = warp_max(sc);
      const float ex = lane < L ? expf(sc - mx) : 0.f;
      const float den = warp_sum(ex);
      if (lane < LP) prob[row][lane] = row_on ? ex / den : 0.f;
    }
  }
  __syncthreads();
  // ---- f_bb = A_b . fb for this warp's D/8 columns; bu = f_bb + f_b -------------------------------------------
  {
    float acc[BMM_MAXD64][4];
#pragma unroll
    for (int n = 0; n < BMM_MAXD64; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
    for (int k0 = 0; k0 < LP; k0 += 8) {
      const float a[4] = {prob[g][k0 + t], prob[g + 8][k0 + t], prob[g][k0 + t + 4], prob[g + 8][k0 + t + 4]};
#pragma unroll
      for (int n = 0; n < BMM_MAXD64; ++n) {
        if (n < ntd) {
          const int col = warp * dpw + n * 8 + g;
          const float bf[2] = {Xs[(size_t)(k0 + t) * DS + col], Xs[(size_t)(k0 + t + 4) * DS + col]};
          mma_16x8x8<PRECISE>(acc[n], a, bf);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < BMM_MAXD64; ++n) {
      if (n < ntd) {
        const int col = warp * dpw + n * 8 + 2 * t;
        if (vA) {
          const float2 x = *reinterpret_cast<const float2*>(Xs + g * DS + col);
          *reinterpret_cast<float2*>(bu + ((size_t)b * L + rA) * D + col) = make_float2(acc[n][0] + x.x, acc[n][1] + x.y);
        }
        if (vB) {
          const float2 x = *reinterpret_cast<const float2*>(Xs + (g + 8) * DS + col);
          *reinterpret_cast<float2*>(bu + ((size_t)b * L + rB) * D + col) = make_float2(acc[n][2] + x.x, acc[n][3] + x.y);
        }
      }
    }
  }
  for (int e = tid; e < BMM_ROWS * LP; e += BMM_THREADS) {
    const int rr = e / LP, j = e - rr * LP;
    if (rr < L && j < L) ab_out[((size_t)b * L + rr) * L + j] = prob[rr][j];
  }
}

// fbar gate of the fast mode for two adjacent columns: x * sigmoid(x * s) = x / (1 + 2^(x * ns)), ns = -s * log2(e) folded
// once per (sample, column).  ex2.approx.ftz / rcp.approx.ftz straight (no denormal / range fix-ups: a huge 2^(.) gives
// rcp(inf) = 0, a flushed one gives x itself) and the fp32 arithmetic as f32x2 instructions: 4.5 issue slots per element
// where `__fdividef(x, 1.0f + __expf(-z))` took ~14 (the per-sample stream kernel ran at 57 % issue-slot utilisation with
// 23 thread instructions per element, ncu).  Both stream kernels use it, so their outputs stay bit-identical.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void gate2_fast(float x0, float x1, float ns0, float ns1, float& g0, float& g1) {
  float z0 = x0, z1 = x1;
  ptx::mul2(z0, z1, ns0, ns1);
  float e0 = ex2_ftz(z0), e1 = ex2_ftz(z1);
  ptx::add2(e0, e1, 1.0f, 1.0f);
  g0 = rcp_ftz(e0); g1 = rcp_ftz(e1);
  ptx::mul2(g0, g1, x0, x1);
}
constexpr float kNegLog2e = -1.4426950408889634f;

// ---- stream:  fbar_ij = sigmoid(fm_ij*fs)*fm_ij for every valid cell;  bu[i] += sum_j A_b[i,j] fbar_ij ----------------
// One WARP per map row (b, i), four rows per CTA, no shared memory and no block barrier: a lane owns the 8
// consecutive columns {256*q + 8*lane} of every cell (one 16-byte load per cell and column group in fast mode),
// walks the row's cells four at a time (8 independent 16-byte loads in flight per lane) and keeps the row's sum in
// registers.  The row's attention weights A_b[i, j(cell)] are fetched once per 32 cells (one coalesced load of the
// cell codes, one gather) and handed out by shuffle, so the per-cell loop has no dependent loads.  Cells are
// summed in map order j: the result does not depend on scheduling or on the batch.
template <typename ActT, int NG, bool PRECISE>
__global__ void __launch_bounds__(128)
boundary_stream_kernel(const float* __restrict__ ab, const float* __restrict__ fs, const ActT* __restrict__ fm,
                       const int32_t* __restrict__ code, const int32_t* __restrict__ row_start, float* __restrict__ bu,
                       ActT* __restrict__ fbar, const float* __restrict__ fbar_bias, int n_rows, int L, int D, int capacity) {
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int grow = blockIdx.x * 4 + warp;              // b * L + i
  if (grow >= n_rows) return;
  const int b = grow / L;
  const int n_lo = row_start[grow], n_hi = min(row_start[grow + 1], capacity);
  if (n_lo >= n_hi) return;                            // empty row: bu stays f_bb + f_b (A_b row is all zero there)
  const float* arow = ab + (size_t)grow * L;
  // blockIdx.y: 256-column group (fast mode launches NG = 1 with D / 256 groups in grid.y: half the registers per thread,
  // 5 resident CTAs per SM instead of 3 -- the kernel is latency-bound, resident warps are what it needs)
  const int col_shift = (int)blockIdx.y * 256;
  fs += col_shift; fm += col_shift; bu += col_shift;
  if (fbar) fbar += col_shift;
  if (fbar_bias) fbar_bias += col_shift;
  const int ldD = D;                                   // row stride of fs / fm / fbar / bu
  const int Dc = min(D - col_shift, NG * 256);         // columns this CTA covers
  // fbar_bias (optional): added to the STORED fbar only (before its rounding) -- the consumer's output bias travels with it
  f8 s8[NG], bm[NG], fb8[NG];
#pragma unroll
  for (int q = 0; q < NG; ++q) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { s8[q].v[e] = 0.f; bm[q].v[e] = 0.f; fb8[q].v[e] = 0.f; }
    if (q * 256 + lane * 8 < Dc) {
      s8[q] = ld8(fs + (size_t)b * ldD + q * 256 + lane * 8);
      if (!PRECISE) {
#pragma unroll
        for (int e = 0; e < 8; ++e) s8[q].v[e] *= kNegLog2e;     // fast mode: see gate2_fast
      }
      if (fbar_bias) fb8[q] = ld8(fbar_bias + q * 256 + lane * 8);
    }
  }
  constexpr int CB = 4;
  for (int seg = n_lo; seg < n_hi; seg += 32) {        // 32 cells of the row per segment
    const int seg_n = min(32, n_hi - seg);
    VML_DBG_ASSERT(seg >= 0 && seg + seg_n <= capacity && (lane >= seg_n || (__ldg(code + seg + lane) & 0xff) < L));
    // the attention weight of a cell hangs on a dependent pair of loads (code -> A_b[i, j]); the first batch of map loads is
    // put in flight between the two so that the row pays two memory round trips, not three
    const int cd_lane = lane < seg_n ? __ldg(code + seg + lane) : 0;
    float a_lane = 0.f;
    for (int c0 = 0; c0 < seg_n; c0 += CB) {
      f8 m[CB][NG];
#pragma unroll
      for (int u = 0; u < CB; ++u) {
        const int n = min(seg + c0 + u, n_hi - 1);
#pragma unroll
        for (int q = 0; q < NG; ++q) {
#pragma unroll
          for (int e = 0; e < 8; ++e) m[u][q].v[e] = 0.f;
          if (q * 256 + lane * 8 < Dc) m[u][q] = ld8(fm + (size_t)n * ldD + q * 256 + lane * 8);
        }
      }
      if (c0 == 0) a_lane = lane < seg_n ? __ldg(arow + (cd_lane & 0xff)) : 0.f;
#pragma unroll
      for (int u = 0; u < CB; ++u) {
        const float a = __shfl_sync(0xffffffffu, a_lane, (c0 + u) & 31);
        if (c0 + u < seg_n) {                          // warp-uniform
#pragma unroll
          for (int q = 0; q < NG; ++q)
            if (q * 256 + lane * 8 < Dc) {
              f8 gv;
              if (PRECISE) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float x = m[u][q].v[e], z = x * s8[q].v[e];
                  gv.v[e] = sigmoidf_(z) * x;
                  bm[q].v[e] = fmaf(a, gv.v[e], bm[q].v[e]);
                }
              } else {
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                  gate2_fast(m[u][q].v[e], m[u][q].v[e + 1], s8[q].v[e], s8[q].v[e + 1], gv.v[e], gv.v[e + 1]);
                  ptx::fma2(bm[q].v[e], bm[q].v[e + 1], a, a, gv.v[e], gv.v[e + 1]);
                }
              }
              if (fbar) {
#pragma unroll
                for (int e = 0; e < 8; e += 2) ptx::add2(gv.v[e], gv.v[e + 1], fb8[q].v[e], fb8[q].v[e + 1]);
                st8(fbar + (size_t)(seg + c0 + u) * ldD + q * 256 + lane * 8, gv);
              }
            }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < NG; ++q)
    if (q * 256 + lane * 8 < Dc) {
      float* o = bu + (size_t)grow * ldD + q * 256 + lane * 8;
      f8 tot = ld8(o);
#pragma unroll
      for (int e = 0; e < 8; ++e) tot.v[e] = tot.v[e] + bm[q].v[e];
      st8(o, tot);
    }
}

// ---- stream, one CTA per SAMPLE (fast mode, L <= 16, D = 256 * {1, 2}) ------------------------------------------------
// Same arithmetic in the same order as boundary_stream_kernel (bit-identical fbar and bu), different schedule.  There a
// warp owns one map row of ~5 cells and pays the dependent chain row_start -> cell code -> A_b[i, j] for it; here the
// sample's attention rows, row offsets and cell codes are staged in shared memory once per CTA (two round trips for up to
// 136 cells x 1 KB), a warp owns the row PAIR (p, L - 1 - p) of one 256-column group -- L + 1 cells for a full-length
// video whatever p is -- and streams them four at a time with the weights read from shared memory.
constexpr int BSS_L = 16, BSS_CB = 4, BSS_LBIG = 64;

template <int D /* 256 or 512 */>
__global__ void __launch_bounds__(D, 1024 / D)
boundary_stream_sample_kernel(const float* __restrict__ ab, const float* __restrict__ fs, const bf16* __restrict__ fm,
                              const int32_t* __restrict__ code, const int32_t* __restrict__ row_start, float* __restrict__ bu,
                              bf16* __restrict__ fbar, const float* __restrict__ fbar_bias, int L, int capacity,
                              bf16* __restrict__ pair_out, int ld_pair) {
  __shared__ float s_ab[BSS_L * BSS_L];
  __shared__ int s_rs[BSS_L + 1];
  __shared__ int s_j[BSS_L * BSS_L];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int ng = D / 256, g = warp % ng, p = warp / ng;          // column group, row pair
  if (tid <= L) s_rs[tid] = min(__ldg(row_start + (size_t)b * L + tid), capacity);
  for (int e = tid; e < L * L; e += blockDim.x) s_ab[e] = __ldg(ab + (size_t)b * L * L + e);
  const int col = g * 256 + lane * 8;
  f8 s8 = ld8(fs + (size_t)b * D + col);
#pragma unroll
  for (int e = 0; e < 8; ++e) s8.v[e] *= kNegLog2e;              // see gate2_fast
  __shared__ __align__(16) float s_bias[D];                   // the stored fbar's bias (kept out of the registers: 64 per thread)
  for (int e = tid; e < D; e += blockDim.x) s_bias[e] = fbar_bias ? __ldg(fbar_bias + e) : 0.f;
  __syncthreads();
  const int n0 = s_rs[0], ncell = min(s_rs[L] - n0, BSS_L * BSS_L);
  const int row0 = p, row1 = L - 1 - p;
  const bool active = row0 <= row1;
  const int lo0 = active ? s_rs[row0] : 0, cnt0 = active ? s_rs[row0 + 1] - lo0 : 0;
  const int lo1 = (active && row1 != row0) ? s_rs[row1] : 0, cnt1 = (active && row1 != row0) ? s_rs[row1 + 1] - lo1 : 0;
  const int total = cnt0 + cnt1;
  auto cell_of = [&](int q) { q = min(q, total - 1); return q < cnt0 ? lo0 + q : lo1 + (q - cnt0); };
  // the first batch of map loads goes out before the cell codes have landed
  uint4 m[BSS_CB];
  if (total > 0) {
#pragma unroll
    for (int u = 0; u < BSS_CB; ++u) m[u] = __ldg(reinterpret_cast<const uint4*>(fm + (size_t)cell_of(u) * D + col));
  }
  for (int e = tid; e < ncell; e += blockDim.x) s_j[e] = __ldg(code + n0 + e) & 0xff;
  __syncthreads();
  float bm0[8], bm1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { bm0[e] = 0.f; bm1[e] = 0.f; }
  if (total > 0) {
  VML_DBG_ASSERT(lo0 >= n0 && lo0 + cnt0 <= n0 + ncell && (cnt1 == 0 || (lo1 >= n0 && lo1 + cnt1 <= n0 + ncell)));
  for (int q0 = 0; q0 < total; q0 += BSS_CB) {
    if (q0 > 0) {
#pragma unroll
      for (int u = 0; u < BSS_CB; ++u) m[u] = __ldg(reinterpret_cast<const uint4*>(fm + (size_t)cell_of(q0 + u) * D + col));
    }
#pragma unroll
    for (int u = 0; u < BSS_CB; ++u) {
      const int q = q0 + u;
      if (q < total) {                                   // warp-uniform
        const bool second = q >= cnt0;
        const int n = second ? lo1 + (q - cnt0) : lo0 + q;
        const float a = s_ab[(second ? row1 : row0) * L + s_j[n - n0]];
        const f8 x8 = unpack8(m[u]);
        f8 gv;
#pragma unroll
        for (int e = 0; e < 8; e += 2) gate2_fast(x8.v[e], x8.v[e + 1], s8.v[e], s8.v[e + 1], gv.v[e], gv.v[e + 1]);
        if (second) {
#pragma unroll
          for (int e = 0; e < 8; e += 2) ptx::fma2(bm1[e], bm1[e + 1], a, a, gv.v[e], gv.v[e + 1]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; e += 2) ptx::fma2(bm0[e], bm0[e + 1], a, a, gv.v[e], gv.v[e + 1]);
        }
        if (fbar) {
          const f8 fb8 = ld8(s_bias + col);
#pragma unroll
          for (int e = 0; e < 8; e += 2) ptx::add2(gv.v[e], gv.v[e + 1], fb8.v[e], fb8.v[e + 1]);
          st8(fbar + (size_t)n * D + col, gv);
        }
      }
    }
  }
  }
  // bu = (f_bb + f_b) + f_bm for this warp's rows; rows without cells keep what the gate + rows kernel wrote
  f8 own0, own1;
#pragma unroll
  for (int e = 0; e < 8; ++e) { own0.v[e] = 0.f; own1.v[e] = 0.f; }
  if (active && (cnt0 > 0 || pair_out)) {
    float* o = bu + ((size_t)b * L + row0) * D + col;
    own0 = ld8(o);
    if (cnt0 > 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) own0.v[e] = own0.v[e] + bm0[e];
      st8(o, own0);
    }
  }
  if (active && row1 != row0 && (cnt1 > 0 || pair_out)) {
    float* o = bu + ((size_t)b * L + row1) * D + col;
    own1 = ld8(o);
    if (cnt1 > 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) own1.v[e] = own1.v[e] + bm1[e];
      st8(o, own1);
    }
  }
  if (pair_out == nullptr) return;
  // ---- a8, first half of the moment unit's operand: operand[n, 0:D] = bu_i * bu_j (models.py:292-295) for the sample's
  // cells, while its boundary rows are still in this CTA (moment_pair_kernel re-reads two 2 KB rows per cell from L2: 285 MB
  // of L2 traffic for 57 MB written on the 640-query Charades pass).  Same product, same rounding.
  __shared__ __align__(16) float s_bu[BSS_L][D];
  if (active) {
    st8(&s_bu[row0][col], own0);
    if (row1 != row0) st8(&s_bu[row1][col], own1);
  }
  __syncthreads();
  if (!active) return;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const int lo = half ? lo1 : lo0, cnt = half ? cnt1 : cnt0;
    const f8 x = half ? own1 : own0;
    for (int q = 0; q < cnt; ++q) {
      const int n = lo + q;
      const f8 y = ld8(&s_bu[s_j[n - n0]][col]);
      f8 o;
#pragma unroll
      for (int e = 0; e < 8; e += 2) { o.v[e] = x.v[e]; o.v[e + 1] = x.v[e + 1]; ptx::mul2(o.v[e], o.v[e + 1], y.v[e], y.v[e + 1]); }
      st8(pair_out + (size_t)n * ld_pair + col, o);
    }
  }
}

// ---- the same per-sample schedule for larger maps (16 < L <= 64: TACoS, ActivityNet) -----------------------------------------
// Dynamic shared memory (attention rows L x L, cell columns as bytes, and -- for the pair products -- an L x D fp32 copy of
// the sample's final boundary rows: 128 KB at L = 64, one CTA per SM); a warp walks the row pairs (p, L - 1 - p),
// p = slot, slot + 8, ... of its 256-column group.  Same arithmetic in the same order as boundary_stream_kernel /
// moment_pair_kernel: bit-identical.
template <int D /* 256 or 512 */>
__global__ void __launch_bounds__(D, 1)
boundary_stream_sample_big_kernel(const float* __restrict__ ab, const float* __restrict__ fs, const bf16* __restrict__ fm,
                                  const int32_t* __restrict__ code, const int32_t* __restrict__ row_start, float* __restrict__ bu,
                                  bf16* __restrict__ fbar, const float* __restrict__ fbar_bias, int L, int capacity,
                                  bf16* __restrict__ pair_out, int ld_pair) {
  extern __shared__ __align__(16) unsigned char bsb_raw[];
  float* s_bu = reinterpret_cast<float*>(bsb_raw);                          // [L][D]   (pair_out only)
  float* s_bias = s_bu + (pair_out ? (size_t)L * D : 0);                    // [D]
  float* s_ab = s_bias + D;                                                 // [L][L]
  int* s_rs = reinterpret_cast<int*>(s_ab + L * L);                         // [L + 1]
  uint8_t* s_j = reinterpret_cast<uint8_t*>(s_rs + L + 1);                  // [L * L]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  constexpr int NG = D / 256;
  const int g = warp % NG, slot = warp / NG;                     // column group, first row pair; 8 pair slots per CTA
  for (int e = tid; e <= L; e += blockDim.x) s_rs[e] = min(__ldg(row_start + (size_t)b * L + e), capacity);
  for (int e = tid; e < L * L; e += blockDim.x) s_ab[e] = __ldg(ab + (size_t)b * L * L + e);
  for (int e = tid; e < D; e += blockDim.x) s_bias[e] = fbar_bias ? __ldg(fbar_bias + e) : 0.f;
  const int col = g * 256 + lane * 8;
  f8 s8 = ld8(fs + (size_t)b * D + col);
#pragma unroll
  for (int e = 0; e < 8; ++e) s8.v[e] *= kNegLog2e;              // see gate2_fast
  __syncthreads();
  const int n0 = s_rs[0], ncell = min(s_rs[L] - n0, L * L);
  for (int e = tid; e < ncell; e += blockDim.x) s_j[e] = (uint8_t)(__ldg(code + n0 + e) & 0xff);
  __syncthreads();
  const int npairs = (L + 1) / 2;
  for (int p = slot; p < npairs; p += 8) {
    const int row0 = p, row1 = L - 1 - p;
    const int lo0 = s_rs[row0], cnt0 = s_rs[row0 + 1] - lo0;
    const int lo1 = row1 != row0 ? s_rs[row1] : 0, cnt1 = row1 != row0 ? s_rs[row1 + 1] - lo1 : 0;
    const int total = cnt0 + cnt1;
    auto cell_of = [&](int q) { q = min(q, total - 1); return q < cnt0 ? lo0 + q : lo1 + (q - cnt0); };
    float bm0[8], bm1[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { bm0[e] = 0.f; bm1[e] = 0.f; }
    for (int q0 = 0; q0 < total; q0 += BSS_CB) {
      uint4 m[BSS_CB];
#pragma unroll
      for (int u = 0; u < BSS_CB; ++u) m[u] = __ldg(reinterpret_cast<const uint4*>(fm + (size_t)cell_of(q0 + u) * D + col));
#pragma unroll
      for (int u = 0; u < BSS_CB; ++u) {
        const int q = q0 + u;
        if (q < total) {                                   // warp-uniform
          const bool second = q >= cnt0;
          const int n = second ? lo1 + (q - cnt0) : lo0 + q;
          VML_DBG_ASSERT(n >= n0 && n - n0 < ncell);
          const float a = s_ab[(second ? row1 : row0) * L + s_j[n - n0]];
          const f8 x8 = unpack8(m[u]);
          f8 gv;
#pragma unroll
          for (int e = 0; e < 8; e += 2) gate2_fast(x8.v[e], x8.v[e + 1], s8.v[e], s8.v[e + 1], gv.v[e], gv.v[e + 1]);
          if (second) {
#pragma unroll
            for (int e = 0; e < 8; e += 2) ptx::fma2(bm1[e], bm1[e + 1], a, a, gv.v[e], gv.v[e + 1]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; e += 2) ptx::fma2(bm0[e], bm0[e + 1], a, a, gv.v[e], gv.v[e + 1]);
          }
          if (fbar) {
            const f8 fb8 = ld8(s_bias + col);
#pragma unroll
            for (int e = 0; e < 8; e += 2) ptx::add2(gv.v[e], gv.v[e + 1], fb8.v[e], fb8.v[e + 1]);
            st8(fbar + (size_t)n * D + col, gv);
          }
        }
      }
    }
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {                 // bu = (f_bb + f_b) + f_bm; rows without cells keep what the rows kernel wrote
      const int row = half ? row1 : row0, cnt = half ? cnt1 : cnt0;
      if (half && row1 == row0) break;
      if (cnt > 0 || pair_out) {
        float* o = bu + ((size_t)b * L + row) * D + col;
        f8 own = ld8(o);
        if (cnt > 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) own.v[e] = own.v[e] + (half ? bm1[e] : bm0[e]);
          st8(o, own);
        }
        if (pair_out) st8(s_bu + (size_t)row * D + col, own);
      }
    }
  }
  if (pair_out == nullptr) return;
  __syncthreads();
  // ---- operand[n, 0:D] = bu_i * bu_j for the sample's cells (models.py:292-295) ----
  for (int p = slot; p < npairs; p += 8) {
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int row = half ? L - 1 - p : p;
      if (half && row == p) break;
      const int lo = s_rs[row], cnt = s_rs[row + 1] - lo;
      const f8 x = ld8(s_bu + (size_t)row * D + col);
      for (int q = 0; q < cnt; ++q) {
        const int n = lo + q;
        const f8 y = ld8(s_bu + (size_t)s_j[n - n0] * D + col);
        f8 o;
#pragma unroll
        for (int e = 0; e < 8; e += 2) { o.v[e] = x.v[e]; o.v[e + 1] = x.v[e + 1]; ptx::mul2(o.v[e], o.v[e + 1], y.v[e], y.v[e + 1]); }
        st8(pair_out + (size_t)n * ld_pair + col, o);
      }
    }
  }
}

// ---- rows, one CTA per SAMPLE for larger maps (16 < L <= 64) ------------------------------------------------------------------
// boundary_rows_mma_kernel gives every 16 map rows their own CTA, each of which re-stages ALL keys and values of the sample --
// within ~100 KB, i.e. 8 rows at a time at D = 512: 16 barrier-separated staging rounds of 16 KB per CTA and four copies of
// the sample's 256 KB per layer (ActivityNet: ~0.4 ms per layer at 640 queries).  Here the sample's L gated rows sit in
// shared memory at once (cp.async.bulk, 132 KB at L = 64, D = 512), every warp owns whole 16 x 16 score tiles over the FULL
// contraction (no split-K partials, no reduction), and the same buffer is re-filled with the value rows f_b for the second
// product.  Same formulas as boundary_rows_mma_kernel; the score sums are associated differently (one chain over D instead
// of eight partial chains), so results agree to rounding, not bit for bit.
constexpr int BRB_THREADS = 512, BRB_WARPS = 16, BRB_L = 64;

// GATE = true: the gate (boundary_gate_mma_kernel: word attention of every map row, G = f_b * (Aq * lmask + f_s)) runs as a
// prologue on the same resident rows -- keys kbt and word states f_w pass through one 24-row buffer, G replaces f_b in place and
// never goes to global memory (inference only: the training path wants G, the word probabilities and u saved).
template <bool PRECISE, bool GATE>
__global__ void __launch_bounds__(BRB_THREADS, 1)
boundary_rows_big_kernel(const float* __restrict__ G, const float* __restrict__ fb, const uint8_t* __restrict__ lmask,
                         float* __restrict__ bu, float* __restrict__ ab_out, int L, int D,
                         const float* __restrict__ qproj, int ld, int off_kbt, int off_betab, const float* __restrict__ fw,
                         const float* __restrict__ fs, const uint8_t* __restrict__ qmask, int Nq) {
  extern __shared__ __align__(16) float brb[];
  const int DS = D + 4;
  const int LP = (L + 7) & ~7, MT = (L + 15) / 16, LS = BRB_L + 1;
  float* Rs = brb;                                  // [MT * 16][DS]  (f_b, then) gated rows G, later the value rows f_b
  float* Ab = Rs + (size_t)MT * 16 * DS;             // [MT * 16][LS]  scores, then attention rows
  const int NQ8 = (Nq + 7) & ~7;
  float* Kw = Ab + (size_t)MT * 16 * LS;             // GATE: [NQ8][DS] keys kbt, then word states f_w
  float* Pw = Kw + (size_t)(GATE ? NQ8 : 0) * DS;    // GATE: [MT * 16][BMM_MAXQ + 1] word scores, then probabilities
  __shared__ __align__(8) uint64_t bar[4];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, g = lane >> 2, t = lane & 3;
  const float* Gb = G + (size_t)b * L * D;
  const float* fbb = fb + (size_t)b * L * D;
  if (tid == 0) { for (int i = 0; i < 4; ++i) ptx::mbar_init(&bar[i], 1); ptx::fence_barrier_init(); }
  for (int e = tid; e < (MT * 16 - L) * (D / 4); e += BRB_THREADS) {       // rows past L: zero (they are MMA operands)
    const int rr = L + e / (D / 4), c4 = (e % (D / 4)) * 4;
    *reinterpret_cast<float4*>(Rs + (size_t)rr * DS + c4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  auto stage = [&](const float* src, uint64_t* mb) {                      // the sample's L rows -> Rs, one bulk copy per row
    if (warp == 0) {
      if (lane == 0) ptx::mbar_arrive_expect_tx(mb, (uint32_t)(L * D) * 4u);
      __syncwarp();
      for (int r = lane; r < L; r += 32) ptx::bulk_load_1d(Rs + (size_t)r * DS, src + (size_t)r * D, (uint32_t)D * 4u, mb);
    }
  };
  if constexpr (GATE) {
    auto stage_words = [&](const float* src, size_t row_stride, uint64_t* mb) {     // Nq rows of D floats -> Kw
      if (warp == 1) {
        if (lane == 0) ptx::mbar_arrive_expect_tx(mb, (uint32_t)(Nq * D) * 4u);
        __syncwarp();
        for (int r = lane; r < Nq; r += 32) ptx::bulk_load_1d(Kw + (size_t)r * DS, src + (size_t)r * row_stride, (uint32_t)D * 4u, mb);
      }
    };
    constexpr int PS = BMM_MAXQ + 1;
    stage(fbb, &bar[0]);
    stage_words(qproj + (size_t)b * Nq * ld + off_kbt, (size_t)ld, &bar[2]);
    ptx::mbar_wait(&bar[0], 0);
    ptx::mbar_wait(&bar[2], 0);
    // ---- word scores f_b . kbt^T over the full D: a warp owns (16-row tile, 8-word tile) items ----
    const int NWT = NQ8 / 8;
    for (int item = warp; item < MT * NWT; item += BRB_WARPS) {
      const int mt = item / NWT, nt = item % NWT;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const float* ra = Rs + (size_t)(mt * 16 + g) * DS;
      const int w = nt * 8 + g;
      const float* rb = Kw + (size_t)min(w, Nq - 1) * DS;
#pragma unroll 4
      for (int k0 = 0; k0 < D; k0 += 8) {
        const float a[4] = {ra[k0 + t], ra[8 * DS + k0 + t], ra[k0 + t + 4], ra[8 * DS + k0 + t + 4]};
        float bf[2] = {0.f, 0.f};
        if (w < Nq) { bf[0] = rb[k0 + t]; bf[1] = rb[k0 + t + 4]; }
        mma_16x8x8<PRECISE>(acc, a, bf);
      }
      float* p0 = Pw + (size_t)(mt * 16 + g) * PS + nt * 8 + 2 * t;
      p0[0] = acc[0]; p0[1] = acc[1]; p0[8 * PS] = acc[2]; p0[8 * PS + 1] = acc[3];
    }
    ptx::fence_proxy_async();
    __syncthreads();                                 // scores complete, kbt dead
    stage_words(fw + (size_t)b * Nq * D, (size_t)D, &bar[3]);
    // ---- masked softmax over the words (models.py:143-150): a warp per row, lane = word ----
    {
      const float mk = (lane < Nq && qmask[(size_t)b * Nq + lane]) ? 1.f : 0.f;
      const float beta = lane < Nq ? qproj[((size_t)b * Nq + lane) * ld + off_betab] : 0.f;
      const float sqrt_d = sqrtf((float)D);
      for (int i = warp; i < MT * 16; i += BRB_WARPS) {
        float sv = -INFINITY;
        if (lane < Nq) {
          sv = ((Pw[(size_t)i * PS + lane] + beta) / sqrt_d) * mk;
          if (mk == 0.f) sv = -1e9f;                     // masked_fill(mask == 0, -1e9)
        }
        const float mx = warp_max(sv);
        const float ex = lane < Nq ? expf(sv - mx) : 0.f;
        const float den = warp_sum(ex);
        if (lane < NQ8) Pw[(size_t)i * PS + lane] = lane < Nq ? ex / den : 0.f;
      }
    }
    __syncthreads();
    ptx::mbar_wait(&bar[3], 0);
    // ---- attended words for this warp's D / 16 columns, gate, G in place of f_b ----
    {
      const int cpw = D / BRB_WARPS, ntd = cpw / 8;
      for (int mt = 0; mt < MT; ++mt) {
        float acc[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
        const float* aa = Pw + (size_t)(mt * 16 + g) * PS;
        for (int k0 = 0; k0 < NQ8; k0 += 8) {
          const float a[4] = {aa[k0 + t], aa[8 * PS + k0 + t], aa[k0 + t + 4], aa[8 * PS + k0 + t + 4]};
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            if (n < ntd) {
              const int col = warp * cpw + n * 8 + g;
              float bf[2];
              bf[0] = (k0 + t < Nq) ? Kw[(size_t)(k0 + t) * DS + col] : 0.f;
              bf[1] = (k0 + t + 4 < Nq) ? Kw[(size_t)(k0 + t + 4) * DS + col] : 0.f;
              mma_16x8x8<PRECISE>(acc[n], a, bf);
            }
          }
        }
        const int rA = mt * 16 + g, rB = rA + 8;
        const float lmA = (rA < L && lmask[(size_t)b * L + rA]) ? 1.f : 0.f, lmB = (rB < L && lmask[(size_t)b * L + rB]) ? 1.f : 0.f;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          if (n < ntd) {
            const int col = warp * cpw + n * 8 + 2 * t;
            const float2 s2 = __ldg(reinterpret_cast<const float2*>(fs + (size_t)b * D + col));
            float2* xa = reinterpret_cast<float2*>(Rs + (size_t)rA * DS + col);
            float2* xb = reinterpret_cast<float2*>(Rs + (size_t)rB * DS + col);
            const float2 va = *xa, vb = *xb;
            *xa = make_float2(va.x * (acc[n][0] * lmA + s2.x), va.y * (acc[n][1] * lmA + s2.y));     // rows past L stay zero
            *xb = make_float2(vb.x * (acc[n][2] * lmB + s2.x), vb.y * (acc[n][3] * lmB + s2.y));
          }
        }
      }
    }
    __syncthreads();                                 // Rs holds G
  } else {
    stage(Gb, &bar[0]);
    ptx::mbar_wait(&bar[0], 0);
  }
  // ---- S = G . G^T over the full D: a warp owns (16-row tile, pair of 8-key tiles) items ----------------------------------
  const int NT = LP / 8, NP = (NT + 1) / 2;
  for (int item = warp; item < MT * NP; item += BRB_WARPS) {
    const int mt = item / NP, n0 = (item % NP) * 2;
    const bool two = n0 + 1 < NT;
    float acc[2][4];
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
    const float* ra = Rs + (size_t)(mt * 16 + g) * DS;
    const float* rb0 = Rs + (size_t)(n0 * 8 + g) * DS;
    const float* rb1 = Rs + (size_t)((two ? n0 + 1 : n0) * 8 + g) * DS;
#pragma unroll 4
    for (int k0 = 0; k0 < D; k0 += 8) {
      const float a[4] = {ra[k0 + t], ra[8 * DS + k0 + t], ra[k0 + t + 4], ra[8 * DS + k0 + t + 4]};
      const float b0[2] = {rb0[k0 + t], rb0[k0 + t + 4]};
      const float b1[2] = {rb1[k0 + t], rb1[k0 + t + 4]};
      mma_16x8x8<PRECISE>(acc[0], a, b0);
      mma_16x8x8<PRECISE>(acc[1], a, b1);
    }
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      if (n == 0 || two) {
        float* p0 = Ab + (size_t)(mt * 16 + g) * LS + (n0 + n) * 8 + 2 * t;
        p0[0] = acc[n][0]; p0[1] = acc[n][1]; p0[8 * LS] = acc[n][2]; p0[8 * LS + 1] = acc[n][3];
      }
    }
  }
  ptx::fence_proxy_async();                          // this thread's reads of G are ordered before the bulk copies that overwrite it
  __syncthreads();                                   // scores complete; every warp is done reading G
  stage(fbb, &bar[1]);                               // value rows over the gated rows, under the softmax
  // ---- masked softmax over the keys (models.py:176-184): a warp per row, lanes over keys -------------------------------
  {
    const float sqrt_d = sqrtf((float)D);
    for (int i = warp; i < L; i += BRB_WARPS) {
      const bool row_on = lmask[(size_t)b * L + i] != 0;
      float* arow = Ab + (size_t)i * LS;
      float sc[2], mx = -INFINITY;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;
        sc[h] = -INFINITY;
        if (j < L) {
          const float mk = lmask[(size_t)b * L + j] ? 1.f : 0.f;
          sc[h] = (arow[j] / sqrt_d) * mk;
          if (mk == 0.f) sc[h] = -1e9f;
        }
        mx = fmaxf(mx, sc[h]);
      }
      mx = warp_max(mx);
      float ex[2], den = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) { ex[h] = lane + 32 * h < L ? expf(sc[h] - mx) : 0.f; den += ex[h]; }
      den = warp_sum(den);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;
        if (j < LP) {
          const float v = (row_on && j < L) ? ex[h] / den : 0.f;
          arow[j] = v;
          if (j < L) ab_out[((size_t)b * L + i) * L + j] = v;
        }
      }
    }
    for (int e = tid; e < (MT * 16 - L) * LP; e += BRB_THREADS) Ab[(size_t)(L + e / LP) * LS + e % LP] = 0.f;   // padded rows
  }
  __syncthreads();
  ptx::mbar_wait(&bar[1], 0);
  // ---- f_bb = A_b . f_b, bu = f_bb + f_b: a warp owns D / 16 columns of every row tile ------------------------------------
  {
    const int cpw = D / BRB_WARPS, ntd = cpw / 8;      // columns per warp (32 at D = 512), 8-column tiles
    for (int mt = 0; mt < MT; ++mt) {
      float acc[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
      const float* aa = Ab + (size_t)(mt * 16 + g) * LS;
      for (int k0 = 0; k0 < LP; k0 += 8) {
        const float a[4] = {aa[k0 + t], aa[8 * LS + k0 + t], aa[k0 + t + 4], aa[8 * LS + k0 + t + 4]};
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          if (n < ntd) {
            const int col = warp * cpw + n * 8 + g;
            const float bf[2] = {Rs[(size_t)(k0 + t) * DS + col], Rs[(size_t)(k0 + t + 4) * DS + col]};
            mma_16x8x8<PRECISE>(acc[n], a, bf);
          }
        }
      }
      const int rA = mt * 16 + g, rB = rA + 8;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        if (n < ntd) {
          const int col = warp * cpw + n * 8 + 2 * t;
          if (rA < L) {
            const float2 x = *reinterpret_cast<const float2*>(Rs + (size_t)rA * DS + col);
            *reinterpret_cast<float2*>(bu + ((size_t)b * L + rA) * D + col) = make_float2(acc[n][0] + x.x, acc[n][1] + x.y);
          }
          if (rB < L) {
            const float2 x = *reinterpret_cast<const float2*>(Rs + (size_t)rB * DS + col);
            *reinterpret_cast<float2*>(bu + ((size_t)b * L + rB) * D + col) = make_float2(acc[n][2] + x.x, acc[n][3] + x.y);
          }
        }
      }
    }
  }
}

template <bool PRECISE>
static int launch_rows(const float* G, const float* fb, const uint8_t* lmask, float* bu, float* ab, int B, vml_dims_t d,
                       cudaStream_t st) {
  // larger maps: one CTA per sample with all of its rows resident (A/B knob: VML_ROWS_TILED=1 keeps the 16-row CTAs)
  if (d.L > BMM_ROWS && d.L <= BRB_L && d.D % (8 * BRB_WARPS) == 0 && d.D / BRB_WARPS <= 32 && getenv("VML_ROWS_TILED") == nullptr &&
      ((reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(fb)) & 15) == 0) {
    static bool regb = (register_kernel("boundary_rows_big_kernel"), true); (void)regb;
    const int MT = ceil_div(d.L, 16);
    const size_t smem = sizeof(float) * ((size_t)MT * 16 * (d.D + 4) + (size_t)MT * 16 * (BRB_L + 1));
    VML_CHECK_ARG(smem <= 226 * 1024);
    VML_CUDA(ensure_dyn_smem((const void*)(boundary_rows_big_kernel<PRECISE, false>), (size_t)((int)smem)));
    boundary_rows_big_kernel<PRECISE, false><<<B, BRB_THREADS, smem, st>>>(G, fb, lmask, bu, ab, d.L, d.D, nullptr, 0, 0, 0, nullptr,
                                                                           nullptr, nullptr, 0);
    return VML_OK;
  }
  const int LP = (d.L + 7) & ~7;
  int JR = LP < BMM_JB ? LP : BMM_JB;
  auto need = [&](int jr) { return sizeof(float) * ((size_t)jr * (d.D + 4) + (size_t)(BMM_WARPS + 1) * BMM_ROWS * (LP + 1) + (size_t)BMM_ROWS * (d.D + 4)); };
  while (JR > 8 && need(JR) > 100 * 1024) JR -= 8;      // <= 100 KB: two CTAs per SM
  const size_t smem = need(JR);
  VML_CHECK_ARG(smem <= 227 * 1024);
  VML_CUDA(ensure_dyn_smem((const void*)(boundary_rows_mma_kernel<PRECISE>), (size_t)((int)smem)));
  dim3 grid(ceil_div(d.L, BMM_ROWS), B);
  boundary_rows_mma_kernel<PRECISE><<<grid, BMM_THREADS, smem, st>>>(G, fb, lmask, bu, ab, d.L, d.D, JR);
  return VML_OK;
}

template <typename ActT, int NG, bool PRECISE>
static int launch_stream(const float* ab, const float* fs, const void* fm, vml_cells_t cells, float* bu, void* fbar,
                         const float* fbar_bias, int B, vml_dims_t d, cudaStream_t st) {
  // NG column groups per warp; when NG * 256 < D the remaining groups run as separate CTAs (grid.y)
  dim3 grid(ceil_div(B * d.L, 4), ceil_div(d.D, NG * 256));
  boundary_stream_kernel<ActT, NG, PRECISE><<<grid, 128, 0, st>>>(
      ab, fs, (const ActT*)fm, cells.code, cells.row_start, bu, (ActT*)fbar, fbar_bias, B * d.L, d.L, d.D, cells.capacity);
  return VML_OK;
}

// the per-sample streaming kernel (fast mode, small maps) can also write the moment operand's first half, bu_i * bu_j
bool boundary_pair_fused(vml_dims_t d, int prec) {
  const char* ss_env = getenv("VML_STREAM_SAMPLE");
  return prec != VML_FP32 && d.L <= BSS_LBIG && (d.D == 256 || d.D == 512) && (ss_env == nullptr || atoi(ss_env) != 0) &&
         getenv("VML_PAIR_SPLIT") == nullptr;
}

int boundary_unit(const float* qproj, int ld, int off_kbt, int off_betab, const float* fw, const float* fs,
                  const float* fb, const void* fm, const uint8_t* qmask, const uint8_t* lmask, vml_cells_t cells,
                  float* g_scratch, float* ab_scratch, float* bu, void* fbar, const float* fbar_bias, float* prob_out, float* u_out,
                  int B, vml_dims_t d, int prec, cudaStream_t st, void* pair_out, int ld_pair) {
  VML_CHECK_ARG(d.Nq <= BMM_MAXQ && d.D % 64 == 0 && d.D <= 64 * BMM_MAXD64 && d.L <= 248 && ld % 4 == 0 && off_kbt % 4 == 0);
  VML_CHECK_ARG(g_scratch != nullptr && ab_scratch != nullptr);
  static bool reg = (register_kernel("boundary_gate_mma_kernel"), register_kernel("boundary_rows_mma_kernel"),
                     register_kernel("boundary_stream_kernel"), true); (void)reg;
  int rc = VML_OK, n_launched = 3;
  if (d.L <= BMM_ROWS) {
    // one row block per sample: gate + rows in one launch, G never leaves the SM
    static bool reg2 = (register_kernel("boundary_gate_rows_kernel"), true); (void)reg2;
    const int rows_kw = 2 * d.Nq > BMM_ROWS ? 2 * d.Nq : BMM_ROWS;
    const size_t smem_f = sizeof(float) * (size_t)(rows_kw + BMM_ROWS) * (d.D + 4);
    // rows moved by cp.async.bulk need 16-byte aligned sources (A/B knob: VML_GATE_BULK=0 stages through registers)
    const char* gb_env = getenv("VML_GATE_BULK");
    const bool bulk_env = gb_env == nullptr || atoi(gb_env) != 0;
    const int bulk = bulk_env && d.D % 4 == 0 && ((reinterpret_cast<uintptr_t>(qproj) | reinterpret_cast<uintptr_t>(fw) |
                                                   reinterpret_cast<uintptr_t>(fb)) & 15) == 0;
#define VML_GR(P, DT)                                                                                                     \
  do {                                                                                                                    \
    VML_CUDA(ensure_dyn_smem((const void*)(boundary_gate_rows_kernel<P, DT>), (size_t)((int)smem_f)));                   \
    boundary_gate_rows_kernel<P, DT><<<B, BMM_THREADS, smem_f, st>>>(qproj, ld, off_kbt, off_betab, fw, fs, fb, qmask, lmask, \
                                                                    g_scratch, prob_out, u_out, bu, ab_scratch, d.L, d.Nq, d.D, bulk); \
  } while (0)
    const char* dt_env = getenv("VML_GATE_DT");                    // A/B knob: VML_GATE_DT=0 keeps D a run-time value
    const bool dt512 = d.D == 512 && (dt_env == nullptr || atoi(dt_env) != 0);
    if (prec == VML_FP32) { if (dt512) VML_GR(true, 512); else VML_GR(true, 0); }
    else { if (dt512) VML_GR(false, 512); else VML_GR(false, 0); }
#undef VML_GR
    n_launched = 2;
  } else if (d.L <= BRB_L && prob_out == nullptr && u_out == nullptr && d.D % (8 * BRB_WARPS) == 0 && d.D / BRB_WARPS <= 32 &&
             getenv("VML_ROWS_TILED") == nullptr && getenv("VML_GATE_SPLIT") == nullptr && ld % 4 == 0 && off_kbt % 4 == 0 &&
             ((reinterpret_cast<uintptr_t>(qproj) | reinterpret_cast<uintptr_t>(fw) | reinterpret_cast<uintptr_t>(fb)) & 15) == 0) {
    // larger maps, inference: gate + rows in one per-sample kernel (the gated rows never leave the SM)
    static bool regg = (register_kernel("boundary_rows_big_kernel"), true); (void)regg;
    const int MT = ceil_div(d.L, 16), NQ8 = (d.Nq + 7) & ~7;
    const size_t smem = sizeof(float) * ((size_t)MT * 16 * (d.D + 4) + (size_t)MT * 16 * (BRB_L + 1) + (size_t)NQ8 * (d.D + 4) +
                                         (size_t)MT * 16 * (BMM_MAXQ + 1));
    VML_CHECK_ARG(smem <= 226 * 1024);
    if (prec == VML_FP32) {
      VML_CUDA(ensure_dyn_smem((const void*)(boundary_rows_big_kernel<true, true>), (size_t)((int)smem)));
      boundary_rows_big_kernel<true, true><<<B, BRB_THREADS, smem, st>>>(nullptr, fb, lmask, bu, ab_scratch, d.L, d.D, qproj, ld, off_kbt,
                                                                         off_betab, fw, fs, qmask, d.Nq);
    } else {
      VML_CUDA(ensure_dyn_smem((const void*)(boundary_rows_big_kernel<false, true>), (size_t)((int)smem)));
      boundary_rows_big_kernel<false, true><<<B, BRB_THREADS, smem, st>>>(nullptr, fb, lmask, bu, ab_scratch, d.L, d.D, qproj, ld, off_kbt,
                                                                          off_betab, fw, fs, qmask, d.Nq);
    }
    n_launched = 2;
  } else {
  dim3 grid(ceil_div(d.L, BMM_ROWS), B);
  const size_t smem_g = sizeof(float) * (size_t)(2 * d.Nq + BMM_ROWS) * (d.D + 4);
  if (prec == VML_FP32) {
    VML_CUDA(ensure_dyn_smem((const void*)(boundary_gate_mma_kernel<true>), (size_t)((int)smem_g)));
    boundary_gate_mma_kernel<true><<<grid, BMM_THREADS, smem_g, st>>>(qproj, ld, off_kbt, off_betab, fw, fs, fb, qmask, lmask, g_scratch, prob_out, u_out, d.L, d.Nq, d.D);
  } else {
    VML_CUDA(ensure_dyn_smem((const void*)(boundary_gate_mma_kernel<false>), (size_t)((int)smem_g)));
    boundary_gate_mma_kernel<false><<<grid, BMM_THREADS, smem_g, st>>>(qproj, ld, off_kbt, off_betab, fw, fs, fb, qmask, lmask, g_scratch, prob_out, u_out, d.L, d.Nq, d.D);
  }
  rc = prec == VML_FP32 ? launch_rows<true>(g_scratch, fb, lmask, bu, ab_scratch, B, d, st)
                        : launch_rows<false>(g_scratch, fb, lmask, bu, ab_scratch, B, d, st);
  if (rc) return rc;
  }
  const int ng = ceil_div(d.D, 256);
  // fast mode, small maps: one CTA per sample (A/B knob: VML_STREAM_SAMPLE=0 selects the warp-per-row kernel)
  const char* ss_env = getenv("VML_STREAM_SAMPLE");
  VML_CHECK_ARG(pair_out == nullptr || boundary_pair_fused(d, prec));
  if (prec != VML_FP32 && d.L <= BSS_L && (d.D == 256 || d.D == 512) && (ss_env == nullptr || atoi(ss_env) != 0)) {
    static bool reg3 = (register_kernel("boundary_stream_sample_kernel"), true); (void)reg3;
    if (d.D == 512) boundary_stream_sample_kernel<512><<<B, 512, 0, st>>>(ab_scratch, fs, (const bf16*)fm, cells.code, cells.row_start, bu,
                                                                      (bf16*)fbar, fbar_bias, d.L, cells.capacity, (bf16*)pair_out, ld_pair);
    else boundary_stream_sample_kernel<256><<<B, 256, 0, st>>>(ab_scratch, fs, (const bf16*)fm, cells.code, cells.row_start, bu,
                                                           (bf16*)fbar, fbar_bias, d.L, cells.capacity, (bf16*)pair_out, ld_pair);
    VML_LAUNCHED(n_launched);
    return VML_OK;
  }
  if (prec != VML_FP32 && d.L <= BSS_LBIG && (d.D == 256 || d.D == 512) && (ss_env == nullptr || atoi(ss_env) != 0)) {
    static bool reg4 = (register_kernel("boundary_stream_sample_big_kernel"), true); (void)reg4;
    const size_t smem = sizeof(float) * ((pair_out ? (size_t)d.L * d.D : 0) + d.D + (size_t)d.L * d.L) + sizeof(int) * (d.L + 1) +
                        (size_t)d.L * d.L + 16;
    if (d.D == 512) {
      VML_CUDA(ensure_dyn_smem((const void*)(boundary_stream_sample_big_kernel<512>), (size_t)((int)smem)));
      boundary_stream_sample_big_kernel<512><<<B, 512, smem, st>>>(ab_scratch, fs, (const bf16*)fm, cells.code, cells.row_start, bu,
                                                                 (bf16*)fbar, fbar_bias, d.L, cells.capacity, (bf16*)pair_out, ld_pair);
    } else {
      VML_CUDA(ensure_dyn_smem((const void*)(boundary_stream_sample_big_kernel<256>), (size_t)((int)smem)));
      boundary_stream_sample_big_kernel<256><<<B, 256, smem, st>>>(ab_scratch, fs, (const bf16*)fm, cells.code, cells.row_start, bu,
                                                                 (bf16*)fbar, fbar_bias, d.L, cells.capacity, (bf16*)pair_out, ld_pair);
    }
    VML_LAUNCHED(n_launched);
    return VML_OK;
  }
  if (prec == VML_FP32) rc = ng <= 1 ? launch_stream<float, 1, true>(ab_scratch, fs, fm, cells, bu, fbar, fbar_bias, B, d, st)
                                     : launch_stream<float, 2, true>(ab_scratch, fs, fm, cells, bu, fbar, fbar_bias, B, d, st);
  else rc = (ng <= 1 || getenv("VML_STREAM_NG2") == nullptr)        // fast mode: one 256-column group per warp (A/B knob: both in one)
                ? launch_stream<bf16, 1, false>(ab_scratch, fs, fm, cells, bu, fbar, fbar_bias, B, d, st)
                : launch_stream<bf16, 2, false>(ab_scratch, fs, fm, cells, bu, fbar, fbar_bias, B, d, st);
  if (rc) return rc;
  VML_LAUNCHED(n_launched);
  return VML_OK;
}

}  // namespace vml
