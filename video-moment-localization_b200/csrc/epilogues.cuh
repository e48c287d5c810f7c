// GEMM epilogues shared by the CUDA-core fp32 GEMM (gemm_simt.cuh) and the tcgen05 bf16 GEMM
// (gemm_umma.cuh).  An epilogue receives N consecutive fp32 accumulators of one output row
// and finishes the reference op that follows the contraction.
#pragma once
#include "common.cuh"

namespace vml {

// out = acc + bias                                   (nn.Linear, models.py:134-135,236-239 ...)
// A thread of the tcgen05 epilogue owns ONE output row (the TMEM lane), so every global access of a warp touches 32
// different rows: the cost is the number of memory instructions (one L1 wavefront per row each), not bytes.  Hence
// 16-byte accesses everywhere: bias as float4 (one broadcast wavefront), bf16 results eight at a time.
template <typename OutT>
struct EpiBias {
  const float* bias;  // [N] or nullptr
  OutT* out;
  int ldo;
  template <int N>
  __device__ __forceinline__ void apply(int row, int col0, const float* acc) const {
    OutT* o = out + (size_t)row * ldo + col0;
    if constexpr (N % 8 == 0) {                // tcgen05 epilogue (N = 32): col0 and ldo are multiples of 8
#pragma unroll
      for (int c = 0; c < N; c += 8) {
        f8 v;
#pragma unroll
        for (int e = 0; e < 8; ++e) v.v[e] = acc[c + e];
        if (bias) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col0 + c));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col0 + c + 4));
          v.v[0] += b0.x; v.v[1] += b0.y; v.v[2] += b0.z; v.v[3] += b0.w;
          v.v[4] += b1.x; v.v[5] += b1.y; v.v[6] += b1.z; v.v[7] += b1.w;
        }
        st8(o + c, v);
      }
    } else {
#pragma unroll
      for (int c = 0; c < N; c += 4) {
        float4 v = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        if (bias) { v.x += bias[col0 + c]; v.y += bias[col0 + c + 1]; v.z += bias[col0 + c + 2]; v.w += bias[col0 + c + 3]; }
        st4(o + c, v);
      }
    }
  }
};

// out = acc + bias, fp32, stored through a per-warp shared-memory transpose (tcgen05 GEMM with 8 epilogue warps only).
// In the TMEM register layout a thread owns one output ROW, so a warp's 16-byte store touches 32 different 128-byte lines
// (32 LSU wavefronts per instruction, 256 per 32 x 32 chunk); the LSTM input projections and the folded query projection
// write 50-70 MB of fp32 per pass behind K = 300-512 and ran at 1-1.8 TB/s with the tensor pipe at 16-24 % (ncu).  Here the
// 32 x 32 chunk goes through a padded shared-memory tile (conflict-free both ways) and leaves as 8 stores of four full
// lines each: 32 wavefronts per chunk.  Same additions, same bits.
struct EpiBiasT {
  static constexpr bool kWarpCollective = true;
  static constexpr int kMaxBN = 128;          // 8 epilogue warps: the scratch below is sized for them
  const float* bias;  // [N] or nullptr
  float* out;
  int ldo;
  template <int N>
  __device__ __forceinline__ void apply_warp(int row, int col0, const float* acc, bool valid) const {
    static_assert(N == 32, "one 32 x 32 chunk per call");
    __shared__ __align__(16) float scratch[8][32][36];
    const int lane = threadIdx.x % 32, w = (threadIdx.x / 32 - 2) & 7;
    float (*s)[36] = scratch[w];
#pragma unroll
    for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(&s[lane][c]) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
    const uint32_t live = __ballot_sync(0xffffffffu, valid);
    __syncwarp();                                                 // tile written -> readable by the other lanes
    const int cc = (lane % 8) * 4, rsub = lane / 8, row0 = row - lane;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) b = __ldg(reinterpret_cast<const float4*>(bias + col0 + cc));
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r = 4 * k + rsub;
      float4 v = *reinterpret_cast<const float4*>(&s[r][cc]);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      if ((live >> r) & 1u) *reinterpret_cast<float4*>(out + (size_t)(row0 + r) * ldo + col0 + cc) = v;
    }
    __syncwarp();                                                 // the next chunk overwrites the tile
  }
};

// a1  fv = (acc + bias)*mask + pe[t]*mask             (VideoEncoder.forward, models.py:27-34)
template <typename OutT>
struct EpiClip {
  const float* bias;     // [D]
  const float* pe;       // [T, D]
  const uint8_t* vmask;  // [B*T]
  int T;
  OutT* out;             // [B*T, D]
  int ldo;
  template <int N>
  __device__ __forceinline__ void apply(int row, int col0, const float* acc) const {
    const float m = vmask[row] ? 1.0f : 0.0f;
    const float* p = pe + (size_t)(row % T) * ldo + col0;
    OutT* o = out + (size_t)row * ldo + col0;
#pragma unroll
    for (int c = 0; c < N; c += 4) {
      float4 v;
      v.x = (acc[c] + bias[col0 + c]) * m + p[c] * m;
      v.y = (acc[c + 1] + bias[col0 + c + 1]) * m + p[c + 1] * m;
      v.z = (acc[c + 2] + bias[col0 + c + 2]) * m + p[c + 2] * m;
      v.w = (acc[c + 3] + bias[col0 + c + 3]) * m + p[c + 3] * m;
      st4(o + c, v);
    }
  }
};

// bf16 variant for the tcgen05 GEMM: the positional-encoding row of a thread (32 columns = eight 16-byte loads) and its mask
// byte are fetched while the MMA of the tile is still running; the scalar version above issued 64 dependent 4-byte loads per
// 32 columns, each a wavefront per row: the epilogue, not the tensor pipe, set the time of the projection (ncu: 14 % tensor
// active, long-scoreboard stalls on the epilogue FADDs).
struct EpiClipPre {
  const float* bias;     // [D]
  const float* pe;       // [T, D]
  const uint8_t* vmask;  // [B*T]
  int T;
  bf16* out;             // [B*T, D]
  int ldo;
  struct Pre { float4 p[8]; float m; };
  __device__ __forceinline__ Pre load(int row, int col0, bool valid) const {
    Pre r;
    if (valid) {
      const float4* p = reinterpret_cast<const float4*>(pe + (size_t)(row % T) * ldo + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) r.p[i] = __ldg(p + i);
      r.m = vmask[row] ? 1.0f : 0.0f;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) r.p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      r.m = 0.f;
    }
    return r;
  }
  template <int N>
  __device__ __forceinline__ void apply_pre(int row, int col0, const float* acc, const Pre& r, bool valid) const {
    static_assert(N == 32, "prefetch bundle holds 32 columns");
    if (!valid) return;
    bf16* o = out + (size_t)row * ldo + col0;
    const float m = r.m;
#pragma unroll
    for (int c = 0; c < N; c += 8) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col0 + c));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col0 + c + 4));
      const float4 p0 = r.p[c / 4], p1 = r.p[c / 4 + 1];
      f8 v;                                   // same expression, term by term, as EpiClip::apply
      v.v[0] = (acc[c] + b0.x) * m + p0.x * m;     v.v[1] = (acc[c + 1] + b0.y) * m + p0.y * m;
      v.v[2] = (acc[c + 2] + b0.z) * m + p0.z * m; v.v[3] = (acc[c + 3] + b0.w) * m + p0.w * m;
      v.v[4] = (acc[c + 4] + b1.x) * m + p1.x * m; v.v[5] = (acc[c + 5] + b1.y) * m + p1.y * m;
      v.v[6] = (acc[c + 6] + b1.z) * m + p1.z * m; v.v[7] = (acc[c + 7] + b1.w) * m + p1.w * m;
      st8(o + c, v);
    }
  }
};

// a6 tail  cu = (acc + bias) + fc + sigmoid(fm*fs)*fm   (ContentUnit.forward, models.py:269-276)
// rows are (cell, clip) pairs: row = cell*C + c.
template <typename ActT>
struct EpiContentOut {
  const float* bias;   // [D]
  const ActT* fc;      // [n*C, D]
  const ActT* fm;      // [n, D]
  const float* fs;     // [B, D]
  const int32_t* code; // [n]
  int C;
  ActT* out;           // [n*C, D]
  int ldo;             // D
  template <int N>
  __device__ __forceinline__ void apply(int row, int col0, const float* acc) const {
    const int cell = row / C;
    const int b = code[cell] >> 16;
    const ActT* x = fc + (size_t)row * ldo + col0;
    const ActT* m = fm + (size_t)cell * ldo + col0;
    const float* s = fs + (size_t)b * ldo + col0;
    ActT* o = out + (size_t)row * ldo + col0;
#pragma unroll
    for (int c = 0; c < N; c += 4) {
      float4 xv = ld4(x + c), mv = ld4(m + c), sv = ld4(s + c), v;
      v.x = (acc[c] + bias[col0 + c]) + xv.x + sigmoidf_(mv.x * sv.x) * mv.x;
      v.y = (acc[c + 1] + bias[col0 + c + 1]) + xv.y + sigmoidf_(mv.y * sv.y) * mv.y;
      v.z = (acc[c + 2] + bias[col0 + c + 2]) + xv.z + sigmoidf_(mv.z * sv.z) * mv.z;
      v.w = (acc[c + 3] + bias[col0 + c + 3]) + xv.w + sigmoidf_(mv.w * sv.w) * mv.w;
      st4(o + c, v);
    }
  }
};

// a6 tail, fused variant for the tcgen05 GEMM (thread == row, C == 4 so one cell == 4 adjacent lanes):
//   cu = (acc + bias) + fc + fbar,   fbar = sigmoid(fm*fs)*fm precomputed per CELL by the boundary unit
//   (it needs the same quantity for f_bm, models.py:191-194 == :272-274), and the moment unit's
//   operand half  mean_c cu  (models.py:297) is reduced with two warp shuffles and stored here.
struct EpiContentOutFused {
  static constexpr bool kWarpCollective = true;
  const float* bias;   // [D]
  const bf16* fc;      // [n*4, D]
  const bf16* fbar;    // [n, D]
  bf16* out;           // [n*4, D]
  bf16* op;            // [n, 2D]: columns D.. receive mean_c cu
  int ldo;             // D
  struct Pre { uint4 x[4], f[4]; };     // 32 columns of fc and fbar (bf16)
  __device__ __forceinline__ Pre load(int row, int col0, bool valid) const {
    Pre p;
    if (valid) {
      const uint4* x = reinterpret_cast<const uint4*>(fc + (size_t)row * ldo + col0);
      const uint4* f = reinterpret_cast<const uint4*>(fbar + (size_t)(row >> 2) * ldo + col0);
#pragma unroll
      for (int i = 0; i < 4; ++i) { p.x[i] = __ldg(x + i); p.f[i] = __ldg(f + i); }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) { p.x[i] = make_uint4(0, 0, 0, 0); p.f[i] = p.x[i]; }
    }
    return p;
  }
  template <int N>
  __device__ __forceinline__ void apply_pre(int row, int col0, const float* acc, const Pre& p, bool valid) const {
    static_assert(N == 32, "prefetch bundle holds 32 columns");
    const int cell = row >> 2;
    bf16* o = out + (size_t)row * ldo + col0;
    const bool writer = valid && ((threadIdx.x & 3) == 0);
    bf16* m = op + (size_t)cell * 2 * ldo + ldo + col0;
#pragma unroll
    for (int c = 0; c < N; c += 8) {
      f8 v;
      if (valid) {
        const f8 xv = unpack8(p.x[c / 8]), fv = unpack8(p.f[c / 8]), bv = ld8(bias + col0 + c);
#pragma unroll
        for (int e = 0; e < 8; ++e) v.v[e] = (acc[c + e] + bv.v[e]) + xv.v[e] + fv.v[e];
        st8(o + c, v);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v.v[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float t = v.v[e];
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        v.v[e] = t * 0.25f;
      }
      if (writer) st8(m + c, v);
    }
  }
};

// a8 tail  mu = (acc + (b_fb + b_fc)) + fm                (MomentUnit.forward, models.py:299-303)
// bf16 variant with a prefetchable operand bundle (fm) for the tcgen05 GEMM
struct EpiMomentOutPre {
  static constexpr bool kCluster2 = true;   // K = 2D, BN = 256: L2 -> smem operand traffic is the roof
  const float* bias;  // [D] = b_fb + b_fc
  const bf16* fm;     // [n, D]
  bf16* out;          // [n, D]
  int ldo;
  int pf_a;           // L2 prefetch of the next tile's operand rows (gemm_umma.cuh)
  struct Pre { uint4 m[4]; };
  __device__ __forceinline__ Pre load(int row, int col0, bool valid) const {
    Pre p;
    const uint4* m = reinterpret_cast<const uint4*>(fm + (size_t)row * ldo + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) p.m[i] = valid ? __ldg(m + i) : make_uint4(0, 0, 0, 0);
    return p;
  }
  template <int N>
  __device__ __forceinline__ void apply_pre(int row, int col0, const float* acc, const Pre& p, bool valid) const {
    static_assert(N == 32, "prefetch bundle holds 32 columns");
    if (!valid) return;
    bf16* o = out + (size_t)row * ldo + col0;
#pragma unroll
    for (int c = 0; c < N; c += 8) {
      const f8 mv = unpack8(p.m[c / 8]);
      const f8 bv = ld8(bias + col0 + c);      // two 16-byte broadcast loads instead of eight scalar ones
      f8 v;
#pragma unroll
      for (int e = 0; e < 8; ++e) v.v[e] = (acc[c + e] + bv.v[e]) + mv.v[e];
      st8(o + c, v);
    }
  }
};

template <typename ActT>
struct EpiMomentOut {
  const float* bias;  // [D] = b_fb + b_fc
  const ActT* fm;     // [n, D]
  ActT* out;          // [n, D]
  int ldo;
  template <int N>
  __device__ __forceinline__ void apply(int row, int col0, const float* acc) const {
    const ActT* m = fm + (size_t)row * ldo + col0;
    ActT* o = out + (size_t)row * ldo + col0;
#pragma unroll
    for (int c = 0; c < N; c += 4) {
      float4 mv = ld4(m + c), v;
      v.x = (acc[c] + bias[col0 + c]) + mv.x;
      v.y = (acc[c + 1] + bias[col0 + c + 1]) + mv.y;
      v.z = (acc[c + 2] + bias[col0 + c + 2]) + mv.z;
      v.w = (acc[c + 3] + bias[col0 + c + 3]) + mv.w;
      st4(o + c, v);
    }
  }
};

// ---- transposing variants of the two big bf16 epilogues (shared-memory scratch, see EpiBiasT and gemm_umma.cuh) --------
// The fp32 32 x 32 chunk of a warp goes through a padded tile; afterwards lane l owns 8 consecutive columns
// (l % 4) * 8 .. of the rows 8k + l / 4, k = 0..3, so that every 16-byte access of a warp covers 8 rows x 64 contiguous
// bytes (full sectors) instead of 32 rows x 16 bytes -- for the result AND for the operands that are added to it (the
// residual / the positional encoding), which are prefetched in that layout.  Same operations in the same order: bit-identical.
constexpr int kEpiTile = 32 * 36 * 4;      // bytes of a padded fp32 32 x 32 tile (conflict-free both ways)

__device__ __forceinline__ void epi_tile_put(float* s, int lane, const float* acc) {
#pragma unroll
  for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(s + lane * 36 + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
  __syncwarp();
}
__device__ __forceinline__ f8 epi_tile_get(const float* s, int r, int cc) {
  const float4 a = *reinterpret_cast<const float4*>(s + r * 36 + cc), b = *reinterpret_cast<const float4*>(s + r * 36 + cc + 4);
  f8 v; v.v[0] = a.x; v.v[1] = a.y; v.v[2] = a.z; v.v[3] = a.w; v.v[4] = b.x; v.v[5] = b.y; v.v[6] = b.z; v.v[7] = b.w;
  return v;
}

// a8, bf16:  fm' = (acc + bias) + fm
struct EpiMomentOutT {            // (the opt-in cluster / CTA-pair GEMM variants keep the register epilogue: VML_EPI_DIRECT=1)
  static constexpr int kScratchPerWarp = kEpiTile;
  const float* bias;  // [D] = b_fb + b_fc
  const bf16* fm;     // [n, D]
  bf16* out;          // [n, D]
  int ldo;
  int pf_a;           // L2 prefetch of the next tile's operand rows (gemm_umma.cuh)
  int rows_cap;       // rows of fm / out that exist (capacity; live rows are decided per launch on the device)
  struct Pre { uint4 m[4]; };
  __device__ __forceinline__ Pre load(int row, int col0, bool) const {
    const int lane = threadIdx.x % 32, row0 = row - lane, cc = (lane % 4) * 8;
    Pre p;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = row0 + 8 * k + lane / 4;
      p.m[k] = r < rows_cap ? __ldg(reinterpret_cast<const uint4*>(fm + (size_t)r * ldo + col0 + cc)) : make_uint4(0, 0, 0, 0);
    }
    return p;
  }
  template <int N>
  __device__ __forceinline__ void apply_pre_scr(int row, int col0, const float* acc, const Pre& p, bool valid, unsigned char* scr) const {
    static_assert(N == 32, "one 32 x 32 chunk per call");
    float* s = reinterpret_cast<float*>(scr);
    const int lane = threadIdx.x % 32, row0 = row - lane, cc = (lane % 4) * 8;
    const uint32_t live = __ballot_sync(0xffffffffu, valid);
    epi_tile_put(s, lane, acc);
    const f8 bv = ld8(bias + col0 + cc);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = 8 * k + lane / 4;
      const f8 a = epi_tile_get(s, r, cc), mv = unpack8(p.m[k]);
      f8 v;
#pragma unroll
      for (int e = 0; e < 8; ++e) v.v[e] = (a.v[e] + bv.v[e]) + mv.v[e];
      if ((live >> r) & 1u) st8(out + (size_t)(row0 + r) * ldo + col0 + cc, v);
    }
    __syncwarp();
  }
};

// a1, bf16:  fv = (acc + bias)*mask + pe[t]*mask
struct EpiClipT {
  static constexpr int kScratchPerWarp = kEpiTile;
  const float* bias;     // [D]
  const float* pe;       // [T, D]
  const uint8_t* vmask;  // [B*T]
  int T;
  bf16* out;             // [B*T, D]
  int ldo;
  int rows_cap;          // B*T
  struct Pre { float4 p[8]; float m[4]; };
  __device__ __forceinline__ Pre load(int row, int col0, bool) const {
    const int lane = threadIdx.x % 32, row0 = row - lane, cc = (lane % 4) * 8;
    Pre q;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = row0 + 8 * k + lane / 4;
      if (r < rows_cap) {
        const float4* p = reinterpret_cast<const float4*>(pe + (size_t)(r % T) * ldo + col0 + cc);
        q.p[2 * k] = __ldg(p); q.p[2 * k + 1] = __ldg(p + 1);
        q.m[k] = vmask[r] ? 1.0f : 0.0f;
      } else {
        q.p[2 * k] = q.p[2 * k + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        q.m[k] = 0.f;
      }
    }
    return q;
  }
  template <int N>
  __device__ __forceinline__ void apply_pre_scr(int row, int col0, const float* acc, const Pre& q, bool valid, unsigned char* scr) const {
    static_assert(N == 32, "one 32 x 32 chunk per call");
    float* s = reinterpret_cast<float*>(scr);
    const int lane = threadIdx.x % 32, row0 = row - lane, cc = (lane % 4) * 8;
    const uint32_t live = __ballot_sync(0xffffffffu, valid);
    epi_tile_put(s, lane, acc);
    const f8 bv = ld8(bias + col0 + cc);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = 8 * k + lane / 4;
      const f8 a = epi_tile_get(s, r, cc);
      const float m = q.m[k];
      const float4 p0 = q.p[2 * k], p1 = q.p[2 * k + 1];
      f8 v;                                   // same expression, term by term, as EpiClip::apply
      v.v[0] = (a.v[0] + bv.v[0]) * m + p0.x * m; v.v[1] = (a.v[1] + bv.v[1]) * m + p0.y * m;
      v.v[2] = (a.v[2] + bv.v[2]) * m + p0.z * m; v.v[3] = (a.v[3] + bv.v[3]) * m + p0.w * m;
      v.v[4] = (a.v[4] + bv.v[4]) * m + p1.x * m; v.v[5] = (a.v[5] + bv.v[5]) * m + p1.y * m;
      v.v[6] = (a.v[6] + bv.v[6]) * m + p1.z * m; v.v[7] = (a.v[7] + bv.v[7]) * m + p1.w * m;
      if ((live >> r) & 1u) st8(out + (size_t)(row0 + r) * ldo + col0 + cc, v);
    }
    __syncwarp();
  }
};

// a8 with the bu_i * bu_j half of the operand GENERATED inside the GEMM (models.py:292-295 + :296-303): four extra warps write
// the A box of the first D/64 k-blocks straight into the pipeline stage -- element (r, k) = bu[b_r, i_r][64 kb + k] *
// bu[b_r, j_r][64 kb + k], rounded to bf16 exactly as vml_moment_pair rounds it -- so the pair tensor is neither written to
// nor read from HBM (57 MB each way per layer on the 640-query Charades pass).  A lane owns one 16-byte chunk (8 columns) of
// 8 rows of the tile: every load of a warp covers 4 rows x 128 contiguous bytes of the (L1-resident) boundary rows.
struct EpiMomentOutGen : EpiMomentOutT {
  static constexpr int kGenWarps = 4;
  const float* bu;        // [B, L, D] fp32
  const int32_t* code;    // [n] cell codes b<<16 | i<<8 | j
  int L, D;
  struct Gen { int oi[8], oj[8]; };            // element offsets of this lane's 8 rows' bu_i / bu_j rows (-1: row past the live count)
  __device__ __forceinline__ bool gen_kblock(int kb) const { return kb * 64 < D; }
  __device__ __forceinline__ void gen_setup(Gen& g, int m0, int gw, int lane, int M) const {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = m0 + gw * 32 + 4 * it + (lane >> 3);
      int oi = -1, oj = -1;
      if (row < M) {
        const int cd = __ldg(code + row);
        const int b = cd >> 16, i = (cd >> 8) & 0xff, j = cd & 0xff;
        oi = (b * L + i) * D; oj = (b * L + j) * D;
      }
      g.oi[it] = oi; g.oj[it] = oj;
    }
  }
  __device__ __forceinline__ void gen_fill(const Gen& g, unsigned char* sa, int kb, int gw, int lane) const {
    const int p = lane & 7, col = kb * 64 + p * 8;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = gw * 32 + 4 * it + (lane >> 3);
      f8 o;
      if (g.oi[it] >= 0) {
        const f8 x = ld8(bu + g.oi[it] + col), y = ld8(bu + g.oj[it] + col);
#pragma unroll
        for (int e = 0; e < 8; ++e) o.v[e] = x.v[e] * y.v[e];
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) o.v[e] = 0.f;
      }
      st8(reinterpret_cast<bf16*>(sa + r * 128 + ((p ^ (r & 7)) << 4)), o);
    }
  }
};

}  // namespace vml
