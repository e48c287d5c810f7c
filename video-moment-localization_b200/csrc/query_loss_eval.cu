// Query encoder recurrence (a2), scaled-IoU BCE loss (a10) and R@n,IoU=m evaluation (a11).
#include <cooperative_groups.h>

#include "common.cuh"
#include "sm100.cuh"

namespace vml {

// =====================================================================================
// a2  bi-LSTM layer recurrence with packed-sequence semantics (models.py:50-62)
// grid = (ceil(B/BT), 2 directions); thread u owns hidden unit u of BT samples.
// =====================================================================================
__global__ void query_lengths_kernel(const uint8_t* __restrict__ qmask, int32_t* __restrict__ qlen, int B, int Nq) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int s = 0;
  for (int k = 0; k < Nq; ++k) s += qmask[(size_t)b * Nq + k];
  qlen[b] = s;
}

// One thread-block CLUSTER of 8 CTAs per (direction, tile of 8 samples).  CTA r keeps the
// recurrent weights of hidden units [r*H/8, (r+1)*H/8) (all four gates) resident in shared
// memory for the whole sequence (H=256: 128 KB fp32), so a time step costs no global weight
// traffic.  Per step: (A) thread (u, ks) accumulates the 4 gates of unit u for all 8 samples
// over K-slice ks; (B) thread (u, s) reduces the 8 slices, applies the cell update for
// (unit u, sample s) and writes h_t into the h buffer of every CTA of the cluster through
// distributed shared memory with st.async, whose completion is counted (complete_tx) on an
// mbarrier of the DESTINATION CTA: a CTA starts step t+1 as soon as the 8 KB of h_t have landed
// in its own buffer -- no cluster-wide barrier, no release fence in the loop.  Outputs are
// staged in shared memory and written once at the end; the next step's input-projection values
// are prefetched one step ahead.
constexpr int LSTM_CL = 8;   // CTAs per cluster == K-slices == samples per tile

__global__ void __cluster_dims__(LSTM_CL, 1, 1)
lstm_cluster_kernel(const float* __restrict__ gin, const float* __restrict__ whh_t, const int32_t* __restrict__ qlen,
                    float* __restrict__ y, bf16* __restrict__ y16, float* __restrict__ fs, bf16* __restrict__ fs16,
                    float* __restrict__ acts, int B, int Nq, int H) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int BT = LSTM_CL;
  extern __shared__ __align__(16) float lsm[];
  const int UH = H / LSTM_CL, KS = H / LSTM_CL;
  float* Wsl = lsm;                       // [H][4][UH]
  float* hbuf = Wsl + (size_t)H * 4 * UH; // [2][BT][H]   (sample-major: a warp's DSMEM store is 128 contiguous bytes)
  float* part = hbuf + 2 * H * BT;        // [8 ks][BT][4][UH]
  float* ybuf = part + LSTM_CL * BT * 4 * UH;  // [Nq][BT][UH] staged outputs of this CTA's units
  __shared__ __align__(8) uint64_t hbar[2];    // one per h buffer: counts the bytes of h_t that have landed
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / LSTM_CL;   // cluster id
  const int dir = cid & 1, b0 = (cid >> 1) * BT;
  const int tid = threadIdx.x;            // blockDim.x == H == UH * 8
  const int u = tid % UH, ks = tid / UH;  // phase A role: (unit, K-slice); phase B role: (unit, sample = ks)
  const int unit = rank * UH + u;

  const float* W = whh_t + (size_t)dir * H * 4 * H;
  for (int k = ks; k < H; k += LSTM_CL)   // rows of W^T: 4 gate segments of UH contiguous floats
#pragma unroll
    for (int g = 0; g < 4; ++g) Wsl[((size_t)k * 4 + g) * UH + u] = W[(size_t)k * 4 * H + g * H + unit];
  for (int e = tid; e < 2 * H * BT; e += blockDim.x) hbuf[e] = 0.f;
  for (int e = tid; e < Nq * BT * UH; e += blockDim.x) ybuf[e] = 0.f;   // zeros past the length (pad_packed_sequence)

  const int s = ks;                        // phase-B sample slot
  const int bs = b0 + s;
  const int my_len = bs < B ? min(qlen[bs], Nq) : 0;
  int maxlen = 0;
  for (int t = 0; t < BT; ++t) maxlen = max(maxlen, (b0 + t < B) ? min(qlen[b0 + t], Nq) : 0);
  float c_state = 0.f, h_state = 0.f;
  // DSMEM destinations: element (s, unit) of every peer's h buffer (lanes of a warp hold
  // consecutive units of one sample) and the peer's mbarriers, as shared::cluster addresses
  uint32_t remote_h[LSTM_CL], remote_bar[LSTM_CL];
#pragma unroll
  for (int r = 0; r < LSTM_CL; ++r) {
    remote_h[r] = ptx::mapa(ptx::smem_u32(hbuf + (size_t)s * H + unit), r);
    remote_bar[r] = ptx::mapa(ptx::smem_u32(&hbar[0]), r);
  }
  if (tid == 0) { ptx::mbar_init(&hbar[0], 1); ptx::mbar_init(&hbar[1], 1); ptx::fence_barrier_init(); }
  const uint32_t step_bytes = (uint32_t)(H * BT * sizeof(float));

  auto load_gin = [&](int step, float (&g)[4]) {
    if (step < my_len) {
      const int t = dir == 0 ? step : my_len - 1 - step;
      const float* p = gin + ((size_t)bs * Nq + t) * 8 * H + (size_t)dir * 4 * H + unit;
#pragma unroll
      for (int q = 0; q < 4; ++q) g[q] = p[q * H];
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) g[q] = 0.f;
    }
  };
  float gnext[4];
  load_gin(0, gnext);
  cluster.sync();

  for (int step = 0; step < maxlen; ++step) {
    const int cur = step & 1, nxt = cur ^ 1;
    const bool act = step < my_len;
    float gpre[4] = {gnext[0], gnext[1], gnext[2], gnext[3]};
    load_gin(step + 1, gnext);              // prefetch: consumed one step later
    if (tid == 0) ptx::mbar_arrive_expect_tx(&hbar[nxt], step_bytes);   // arm this step's receive barrier
    // ---- phase A: partial gates over K-slice ks, all BT samples ---------------------------
    float acc[BT][4];
#pragma unroll
    for (int a = 0; a < BT; ++a)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[a][q] = 0.f;
    const float* hb = hbuf + (size_t)cur * H * BT;
    for (int kk = 0; kk < KS; kk += 4) {
      const int k0 = ks * KS + kk;
      float4 h4[BT];
#pragma unroll
      for (int a = 0; a < BT; ++a) h4[a] = *reinterpret_cast<const float4*>(hb + (size_t)a * H + k0);   // warp-uniform: broadcast
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float* wr = Wsl + (size_t)(k0 + j) * 4 * UH + u;
        const float w0 = wr[0], w1 = wr[UH], w2 = wr[2 * UH], w3 = wr[3 * UH];
#pragma unroll
        for (int a = 0; a < BT; ++a) {
          const float hv = j == 0 ? h4[a].x : j == 1 ? h4[a].y : j == 2 ? h4[a].z : h4[a].w;
          acc[a][0] = fmaf(hv, w0, acc[a][0]); acc[a][1] = fmaf(hv, w1, acc[a][1]);
          acc[a][2] = fmaf(hv, w2, acc[a][2]); acc[a][3] = fmaf(hv, w3, acc[a][3]);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < BT; ++a)
#pragma unroll
      for (int q = 0; q < 4; ++q) part[(((size_t)ks * BT + a) * 4 + q) * UH + u] = acc[a][q];
    __syncthreads();
    // ---- phase B: reduce slices, cell update for (unit, sample s), broadcast h -------------
    if (act) {
      float gate[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v = gpre[q];
#pragma unroll
        for (int p = 0; p < LSTM_CL; ++p) v += part[(((size_t)p * BT + s) * 4 + q) * UH + u];
        gate[q] = v;
      }
      const float ig = sigmoidf_(gate[0]), fg = sigmoidf_(gate[1]), gg = tanhf(gate[2]), og = sigmoidf_(gate[3]);
      c_state = fg * c_state + ig * gg;
      h_state = og * tanhf(c_state);
      const int t = dir == 0 ? step : my_len - 1 - step;
      ybuf[((size_t)t * BT + s) * UH + u] = h_state;
      if (acts) {            // training forward: i, f, g, o (after the non-linearities) and c of this step -> BPTT
        float* a = acts + (((size_t)bs * Nq + t) * 2 + dir) * 5 * H + unit;
        a[0] = ig; a[H] = fg; a[2 * H] = gg; a[3 * H] = og; a[4 * H] = c_state;
      }
    }
#pragma unroll
    for (int r = 0; r < LSTM_CL; ++r)
      ptx::st_async_f32(remote_h[r] + (uint32_t)(nxt * H * BT * sizeof(float)), h_state, remote_bar[r] + (uint32_t)(nxt * 8));
    ptx::mbar_wait(&hbar[nxt], (uint32_t)((step >> 1) & 1));   // all of h_t has landed in MY buffer
  }
  // ---- write the staged outputs: y[b, t, dir*H + unit] for this CTA's UH units --------------
  for (int e = tid; e < Nq * BT * UH; e += blockDim.x) {
    const int uu = e % UH, ss = (e / UH) % BT, t = e / (UH * BT);
    if (b0 + ss < B) {
      const size_t o = ((size_t)(b0 + ss) * Nq + t) * 2 * H + (size_t)dir * H + rank * UH + uu;
      const float v = ybuf[e];
      y[o] = v;
      if (y16) y16[o] = __float2bfloat16_rn(v);
    }
  }
  if (bs < B) {
    if (fs) fs[(size_t)bs * 2 * H + (size_t)dir * H + unit] = h_state;  // fwd: h(len-1); bwd: h(0)
    if (fs16) fs16[(size_t)bs * 2 * H + (size_t)dir * H + unit] = __float2bfloat16_rn(h_state);
  }
  cluster.sync();                           // no CTA may exit while peers can still address its smem
}

// -------------------------------------------------------------------------------------
// BPTT of one bi-LSTM layer (training path, behind loss.backward(), main.py:150), same decomposition as the forward:
// an 8-CTA cluster per (direction, tile of 8 samples), CTA r owns hidden units [32r, 32r + 32) and keeps the 4 x 32 rows
// of W_hh that produce THEIR gates in shared memory (128 KB).  A step is
//   (1) thread (unit, sample): gate gradients from the saved activations and dh = dy + dh_rec  -> dgin / dgin_rec rows;
//   (2) thread k: partial dh_prev[sample][k] = sum over the CTA's 128 gate rows of dg[sample][row] * W_hh[row][k];
//   (3) reduce-scatter over the cluster: the partial of column k goes to the CTA that owns unit k (st.async into a
//       per-source slot, counted on the destination's mbarrier); the owner sums its 8 slots -> dh_rec of the next step.
// Samples past their length contribute zeros, so ragged tiles need no special casing.  (The round-1 kernel ran one CTA
// per (sample, direction) and re-read all of W_hh from L2 every step: 1.0 ms per layer at B = 64 against ~0.1 ms.)
// -------------------------------------------------------------------------------------
__global__ void __cluster_dims__(LSTM_CL, 1, 1)
lstm_bwd_cluster_kernel(const float* __restrict__ dy, const float* __restrict__ dfs, const float* __restrict__ whh,
                        const float* __restrict__ acts, const int32_t* __restrict__ qlen, float* __restrict__ dgin,
                        float* __restrict__ dgin_rec, int B, int Nq, int H) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int BT = LSTM_CL;
  extern __shared__ __align__(16) float lsm[];
  const int UH = H / LSTM_CL, R = 4 * UH;          // R gate rows owned by this CTA
  float* Wsl = lsm;                                // [R][H]   row q * UH + u  <-  W_hh[q * H + rank * UH + u][:]
  float* dgs = Wsl + (size_t)R * H;                // [BT][R]  gate gradients of the current step
  float* recv = dgs + BT * R;                      // [2][LSTM_CL src][BT][UH]
  __shared__ __align__(8) uint64_t rbar[2];
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / LSTM_CL;
  const int dir = cid & 1, b0 = (cid >> 1) * BT;
  const int tid = threadIdx.x;                     // blockDim.x == H
  const int u = tid % UH, s = tid / UH;            // phase (1) / (3) role: (unit, sample)
  const int unit = rank * UH + u;
  const float* W = whh + (size_t)dir * 4 * H * H;  // [4H][H]
  for (int e = tid; e < R * H; e += blockDim.x) {
    const int row = e / H, k = e - row * H, q = row / UH, uu = row - q * UH;
    Wsl[e] = W[((size_t)q * H + rank * UH + uu) * H + k];
  }
  for (int e = tid; e < 2 * LSTM_CL * BT * UH; e += blockDim.x) recv[e] = 0.f;
  const int bs = b0 + s;
  const int my_len = bs < B ? min(qlen[bs], Nq) : 0;
  int maxlen = 0;
  for (int t = 0; t < BT; ++t) maxlen = max(maxlen, (b0 + t < B) ? min(qlen[b0 + t], Nq) : 0);
  if (bs < B)
    for (int t = my_len; t < Nq; ++t) {            // rows past the length: zero gradient of the input projections
      float* o = dgin + ((size_t)bs * Nq + t) * 8 * H + (size_t)dir * 4 * H + unit;
      float* o2 = dgin_rec + ((size_t)bs * Nq + t) * 8 * H + (size_t)dir * 4 * H + unit;
#pragma unroll
      for (int q = 0; q < 4; ++q) { o[q * H] = 0.f; o2[q * H] = 0.f; }
    }
  // phase (2) role: output column k = tid, owned by CTA k / UH; my slot there is indexed by MY rank
  const int k_owner = tid / UH, k_local = tid % UH;
  uint32_t remote_slot = ptx::mapa(ptx::smem_u32(recv + ((size_t)rank * BT) * UH + k_local), k_owner);
  uint32_t remote_bar = ptx::mapa(ptx::smem_u32(&rbar[0]), k_owner);
  if (tid == 0) { ptx::mbar_init(&rbar[0], 1); ptx::mbar_init(&rbar[1], 1); ptx::fence_barrier_init(); }
  const uint32_t step_bytes = (uint32_t)(LSTM_CL * BT * UH * sizeof(float));
  float dh_rec = 0.f, dc_next = 0.f;
  cluster.sync();

  for (int step = maxlen - 1; step >= 0; --step) {
    const int it = maxlen - 1 - step, cur = it & 1;
    if (tid == 0) ptx::mbar_arrive_expect_tx(&rbar[cur], step_bytes);
    // ---- (1) gate gradients of (unit, sample s) --------------------------------------------------------
    float dg[4] = {0.f, 0.f, 0.f, 0.f};
    if (step < my_len) {
      const int t = dir == 0 ? step : my_len - 1 - step;
      const float* a = acts + (((size_t)bs * Nq + t) * 2 + dir) * 5 * H + unit;
      const float ig = a[0], fg = a[H], gg = a[2 * H], og = a[3 * H], c = a[4 * H];
      float c_prev = 0.f;
      if (step > 0) {
        const int tp = dir == 0 ? t - 1 : t + 1;
        c_prev = acts[(((size_t)bs * Nq + tp) * 2 + dir) * 5 * H + 4 * H + unit];
      }
      float dh = dy[((size_t)bs * Nq + t) * 2 * H + (size_t)dir * H + unit] + dh_rec;
      if (dfs && step == my_len - 1) dh += dfs[(size_t)bs * 2 * H + (size_t)dir * H + unit];   // fs = h of the last processed step
      const float tc = tanhf(c);
      dg[3] = dh * tc * og * (1.f - og);
      const float dc = dc_next + dh * og * (1.f - tc * tc);
      dg[0] = dc * gg * ig * (1.f - ig);
      dg[2] = dc * ig * (1.f - gg * gg);
      dg[1] = dc * c_prev * fg * (1.f - fg);
      dc_next = dc * fg;
      float* o = dgin + ((size_t)bs * Nq + t) * 8 * H + (size_t)dir * 4 * H + unit;
      float* o2 = dgin_rec + ((size_t)bs * Nq + t) * 8 * H + (size_t)dir * 4 * H + unit;
      const float z = step == 0 ? 0.f : 1.f;         // the first processed step has no recurrent input
#pragma unroll
      for (int q = 0; q < 4; ++q) { o[q * H] = dg[q]; o2[q * H] = dg[q] * z; }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) dgs[(size_t)s * R + q * UH + u] = dg[q];
    __syncthreads();
    // ---- (2) partial dh_prev[a][k] over this CTA's gate rows, (3) scatter to the owner of column k ------------
    if (step > 0) {
      float acc[BT];
#pragma unroll
      for (int a = 0; a < BT; ++a) acc[a] = 0.f;
      for (int r0 = 0; r0 < R; r0 += 4) {
        float4 d4[BT];
#pragma unroll
        for (int a = 0; a < BT; ++a) d4[a] = *reinterpret_cast<const float4*>(dgs + (size_t)a * R + r0);   // warp-uniform: broadcast
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float w = Wsl[(size_t)(r0 + j) * H + tid];
#pragma unroll
          for (int a = 0; a < BT; ++a) {
            const float dv = j == 0 ? d4[a].x : j == 1 ? d4[a].y : j == 2 ? d4[a].z : d4[a].w;
            acc[a] = fmaf(dv, w, acc[a]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < BT; ++a)
        ptx::st_async_f32(remote_slot + (uint32_t)((cur * LSTM_CL * BT * UH + a * UH) * sizeof(float)), acc[a],
                          remote_bar + (uint32_t)(cur * 8));
      ptx::mbar_wait(&rbar[cur], (uint32_t)((it >> 1) & 1));      // the 8 partials of MY units have landed
      float v = 0.f;
#pragma unroll
      for (int p = 0; p < LSTM_CL; ++p) v += recv[((size_t)(cur * LSTM_CL + p) * BT + s) * UH + u];
      dh_rec = v;
    }
    __syncthreads();                                 // dgs is rewritten by the next step
  }
  cluster.sync();                                    // no CTA may exit while peers can still address its smem
}

// -------------------------------------------------------------------------------------
// Fast-mode recurrence (bf16 operands, fp32 state), H = 256: h_{t-1}.W_hh^T on the tensor cores.
// One 4-CTA cluster per (direction, tile of 16 samples).  CTA r owns hidden units [64r, 64r+64);
// warp w of it owns 8 units and keeps the 4 x 8 x 256 slice of W_hh (all four gates) as
// mma.sync B-fragments IN REGISTERS for the whole sequence (128 registers per thread, loaded once
// from a fragment-ordered bf16 copy packed on the host side).  A time step of a warp is
// 64 x mma.m16n8k16 (A = h_{t-1} of the 16 samples from shared memory) whose accumulator layout
// gives every thread the i,f,g,o pre-activations of 2 units x 2 samples, so the cell update is
// thread-local.  The new h (bf16 pairs) goes to the h buffer of all 4 CTAs with st.async counted on
// the destination's mbarrier; warps never meet at a CTA barrier inside the loop.
// -------------------------------------------------------------------------------------
constexpr int LT_CL = 4, LT_BT = 16, LT_H = 256, LT_HP = 264;   // h row pitch (bf16): conflict-free fragment loads

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void st_async_u32(uint32_t cluster_addr, uint32_t v, uint32_t cluster_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(cluster_addr), "r"(v), "r"(cluster_mbar) : "memory");
}

// single-MUFU activations for the fast mode (tanh.approx.f32: ~2^-11 relative error, far inside bf16's)
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

__global__ void __cluster_dims__(LT_CL, 1, 1) __launch_bounds__(256, 1)
lstm_tc_kernel(const float* __restrict__ gin, const uint4* __restrict__ wfrag, const int32_t* __restrict__ qlen,
               float* __restrict__ y, bf16* __restrict__ y16, float* __restrict__ fs, bf16* __restrict__ fs16, int B, int Nq) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int H = LT_H;
  __shared__ __align__(16) bf16 hbuf[2][LT_BT][LT_HP];
  __shared__ __align__(8) uint64_t hbar[2];
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / LT_CL;
  const int dir = cid & 1, b0 = (cid >> 1) * LT_BT;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int t4 = lane & 3, r = lane >> 2;
  const int U = rank * 64 + warp * 8 + 2 * t4;          // this thread's two adjacent hidden units: U, U+1

  // recurrent weights -> registers (fragment order: chunk-major, coalesced 16-byte loads)
  uint32_t wreg[128];
  {
    const uint4* src = wfrag + ((size_t)((dir * LT_CL + rank) * 8 + warp) * 32) * 32;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const uint4 v = __ldg(src + c * 32 + lane);
      wreg[4 * c] = v.x; wreg[4 * c + 1] = v.y; wreg[4 * c + 2] = v.z; wreg[4 * c + 3] = v.w;
    }
  }
  for (int e = tid; e < 2 * LT_BT * LT_HP / 2; e += blockDim.x) reinterpret_cast<uint32_t*>(&hbuf[0][0][0])[e] = 0u;
  if (tid == 0) { ptx::mbar_init(&hbar[0], 1); ptx::mbar_init(&hbar[1], 1); ptx::fence_barrier_init(); }

  const int smp[2] = {b0 + r, b0 + r + 8};
  int len[2];
#pragma unroll
  for (int a = 0; a < 2; ++a) len[a] = smp[a] < B ? min(qlen[smp[a]], Nq) : 0;
  int maxlen = 0;
  for (int tt = 0; tt < LT_BT; ++tt) maxlen = max(maxlen, (b0 + tt < B) ? min(qlen[b0 + tt], Nq) : 0);

  // DSMEM destinations of this thread's h pairs in every CTA of the cluster
  uint32_t remote_h[LT_CL][2], remote_bar[LT_CL];
#pragma unroll
  for (int d = 0; d < LT_CL; ++d) {
    remote_h[d][0] = ptx::mapa(ptx::smem_u32(&hbuf[0][r][U]), d);
    remote_h[d][1] = ptx::mapa(ptx::smem_u32(&hbuf[0][r + 8][U]), d);
    remote_bar[d] = ptx::mapa(ptx::smem_u32(&hbar[0]), d);
  }
  constexpr uint32_t BUF_BYTES = LT_BT * LT_HP * 2, STEP_BYTES = LT_BT * H * 2;

  float gpre[2][4][2];                       // input-projection values of the coming step: [sample][gate][unit]
  auto load_gin = [&](int step) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      if (step < len[a]) {
        const int t = dir == 0 ? step : len[a] - 1 - step;
        const float* p = gin + ((size_t)smp[a] * Nq + t) * 8 * H + (size_t)dir * 4 * H + U;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 v = __ldg(reinterpret_cast<const float2*>(p + q * H));
          gpre[a][q][0] = v.x; gpre[a][q][1] = v.y;
        }
      }
    }
  };
  float c_state[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, h_state[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  load_gin(0);
  cluster.sync();

  for (int step = 0; step < maxlen; ++step) {
    const int cur = step & 1, nxt = cur ^ 1;
    if (tid == 0) ptx::mbar_arrive_expect_tx(&hbar[nxt], STEP_BYTES);       // arm this step's receive barrier
    if (step > 0) ptx::mbar_wait(&hbar[cur], (uint32_t)(((step - 1) >> 1) & 1));   // h_{t-1} has landed here
    float acc[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[q][e] = 0.f;
    const bf16* hr0 = &hbuf[cur][r][2 * t4];
    const bf16* hr1 = &hbuf[cur][r + 8][2 * t4];
#pragma unroll
    for (int ks = 0; ks < H / 16; ++ks) {
      uint32_t a[4];
      a[0] = *reinterpret_cast<const uint32_t*>(hr0 + ks * 16);
      a[1] = *reinterpret_cast<const uint32_t*>(hr1 + ks * 16);
      a[2] = *reinterpret_cast<const uint32_t*>(hr0 + ks * 16 + 8);
      a[3] = *reinterpret_cast<const uint32_t*>(hr1 + ks * 16 + 8);
#pragma unroll
      for (int q = 0; q < 4; ++q) mma_bf16_16816(acc[q], a, wreg[(ks * 4 + q) * 2], wreg[(ks * 4 + q) * 2 + 1]);
    }
    // accumulator (q, e): gate q of (sample e/2, unit e%2)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      if (step < len[a]) {
        const int t = dir == 0 ? step : len[a] - 1 - step;
#pragma unroll
        for (int un = 0; un < 2; ++un) {
          const float ig = sigmoid_fast(acc[0][2 * a + un] + gpre[a][0][un]), fg = sigmoid_fast(acc[1][2 * a + un] + gpre[a][1][un]);
          const float gg = tanh_fast(acc[2][2 * a + un] + gpre[a][2][un]), og = sigmoid_fast(acc[3][2 * a + un] + gpre[a][3][un]);
          c_state[a][un] = fg * c_state[a][un] + ig * gg;
          h_state[a][un] = og * tanh_fast(c_state[a][un]);
        }
        const size_t o = ((size_t)smp[a] * Nq + t) * 2 * H + (size_t)dir * H + U;
        *reinterpret_cast<float2*>(y + o) = make_float2(h_state[a][0], h_state[a][1]);
        if (y16) *reinterpret_cast<__nv_bfloat162*>(y16 + o) = __floats2bfloat162_rn(h_state[a][0], h_state[a][1]);
      }
    }
    load_gin(step + 1);                       // consumed one step later: its latency hides behind the exchange
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const __nv_bfloat162 hp = __floats2bfloat162_rn(h_state[a][0], h_state[a][1]);
      const uint32_t bits = *reinterpret_cast<const uint32_t*>(&hp);
#pragma unroll
      for (int d = 0; d < LT_CL; ++d)
        st_async_u32(remote_h[d][a] + (uint32_t)nxt * BUF_BYTES, bits, remote_bar[d] + (uint32_t)(nxt * 8));
    }
  }
  // zeros past each sample's length (pad_packed_sequence), final states
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    if (smp[a] < B) {
      for (int t = len[a]; t < Nq; ++t) {
        const size_t o = ((size_t)smp[a] * Nq + t) * 2 * H + (size_t)dir * H + U;
        *reinterpret_cast<float2*>(y + o) = make_float2(0.f, 0.f);
        if (y16) *reinterpret_cast<__nv_bfloat162*>(y16 + o) = __floats2bfloat162_rn(0.f, 0.f);
      }
      const size_t o = (size_t)smp[a] * 2 * H + (size_t)dir * H + U;   // fwd: h(len-1); bwd: h(0)
      if (fs) *reinterpret_cast<float2*>(fs + o) = make_float2(h_state[a][0], h_state[a][1]);
      if (fs16) *reinterpret_cast<__nv_bfloat162*>(fs16 + o) = __floats2bfloat162_rn(h_state[a][0], h_state[a][1]);
    }
  }
  if (maxlen > 0) ptx::mbar_wait(&hbar[maxlen & 1], (uint32_t)(((maxlen - 1) >> 1) & 1));   // the last step's h has landed here too
  cluster.sync();                             // no CTA exits while peers may still address its shared memory
}

int lstm_layer_tc(const float* gin, const void* wfrag, const int32_t* qlen, float* y, void* y16, float* fs, void* fs16, int B,
                  int Nq, int H, cudaStream_t st) {
  VML_CHECK_ARG(H == LT_H && B > 0 && Nq > 0);
  static bool reg = (register_kernel("lstm_tc_kernel"), true); (void)reg;
  const int clusters = 2 * ceil_div(B, LT_BT);
  lstm_tc_kernel<<<clusters * LT_CL, 256, 0, st>>>(gin, (const uint4*)wfrag, qlen, y, (bf16*)y16, fs, (bf16*)fs16, B, Nq);
  VML_LAUNCHED(1);
  return VML_OK;
}

int query_lengths(const uint8_t* qmask, int32_t* qlen, int B, int Nq, cudaStream_t st) {
  static bool reg = (register_kernel("query_lengths_kernel"), true); (void)reg;
  query_lengths_kernel<<<ceil_div(B, 128), 128, 0, st>>>(qmask, qlen, B, Nq);
  VML_LAUNCHED(1);
  return VML_OK;
}

int lstm_layer(const float* gin, const float* whh_t, const int32_t* qlen, float* y, void* y16, float* fs, void* fs16, int B,
               int Nq, int H, cudaStream_t st, float* acts) {
  VML_CHECK_ARG(H % 32 == 0 && H <= 1024);
  static bool reg = (register_kernel("lstm_cluster_kernel"), true); (void)reg;
  const int UH = H / LSTM_CL;
  const size_t smem = sizeof(float) * ((size_t)H * 4 * UH + 2 * (size_t)H * LSTM_CL + (size_t)LSTM_CL * LSTM_CL * 4 * UH +
                                       (size_t)Nq * LSTM_CL * UH);
  VML_CHECK_ARG(smem <= 227 * 1024);
  VML_CUDA(ensure_dyn_smem((const void*)(lstm_cluster_kernel), (size_t)((int)smem)));
  const int clusters = 2 * ceil_div(B, LSTM_CL);
  lstm_cluster_kernel<<<clusters * LSTM_CL, H, smem, st>>>(gin, whh_t, qlen, y, (bf16*)y16, fs, (bf16*)fs16, acts, B, Nq, H);
  VML_LAUNCHED(1);
  return VML_OK;
}

int lstm_bwd_cluster(const float* dy, const float* dfs, const float* whh, const float* acts, const int32_t* qlen, float* dgin,
                     float* dgin_rec, int B, int Nq, int H, cudaStream_t st) {
  VML_CHECK_ARG(H % 32 == 0 && H <= 1024);
  static bool reg = (register_kernel("lstm_bwd_cluster_kernel"), true); (void)reg;
  const int UH = H / LSTM_CL;
  const size_t smem = sizeof(float) * ((size_t)4 * UH * H + (size_t)LSTM_CL * 4 * UH + 2 * (size_t)LSTM_CL * LSTM_CL * UH);
  if (smem > 227 * 1024) return 1;               // caller falls back to the one-CTA-per-sample kernel
  VML_CUDA(ensure_dyn_smem((const void*)(lstm_bwd_cluster_kernel), (size_t)((int)smem)));
  const int clusters = 2 * ceil_div(B, LSTM_CL);
  lstm_bwd_cluster_kernel<<<clusters * LSTM_CL, H, smem, st>>>(dy, dfs, whh, acts, qlen, dgin, dgin_rec, B, Nq, H);
  VML_LAUNCHED(1);
  return VML_OK;
}

// =====================================================================================
// a10  scaled-IoU BCE loss (main.py:89-116; BCELoss(reduction=None) read as 'none')
// block b: per-sample masked means of the four terms -> scratch[4][B]; a one-block
// finalize sums samples in index order (deterministic) and forms L_m+L_s+L_e+0.5 L_a.
// =====================================================================================
__device__ __forceinline__ float bce_elem(float p, float y) {
  const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);
  return -(y * lp + (1.f - y) * l1p);
}
// weighted two-layer form of main.py:92-94
__device__ __forceinline__ float scaled_bce(float p, float y, float s) {
  return (s * y) * bce_elem(p, y) + ((1.f - s) * (1.f - y)) * bce_elem(1.f - p, 1.f - y);
}
// d/dp, PyTorch's binary_cross_entropy_backward: (p - y) / max(p (1-p), 1e-12) * weight
__device__ __forceinline__ float bce_grad(float p, float y, float w) {
  return (p - y) / fmaxf((1.f - p) * p, 1e-12f) * w;
}

__device__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (warp == 0) {
    t = lane < (blockDim.x / 32) ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

__global__ void __launch_bounds__(256)
loss_sample_kernel(const float* __restrict__ pm, const uint8_t* __restrict__ ym, const float* __restrict__ sm,
                   const uint8_t* __restrict__ mmask, const float* __restrict__ ps, const uint8_t* __restrict__ ys,
                   const float* __restrict__ ss, const float* __restrict__ pe, const uint8_t* __restrict__ ye,
                   const float* __restrict__ se, const float* __restrict__ pa, const uint8_t* __restrict__ ya,
                   const uint8_t* __restrict__ lmask, int B, int L, float* __restrict__ scratch,
                   float* __restrict__ g_pm, float* __restrict__ g_ps, float* __restrict__ g_pe, float* __restrict__ g_pa) {
  __shared__ float red[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const size_t mo = (size_t)b * L * L, lo = (size_t)b * L;
  float lm_sum = 0.f, lm_cnt = 0.f;
  for (int e = tid; e < L * L; e += blockDim.x) {
    const float k = mmask[mo + e] ? 1.f : 0.f, y = ym[mo + e] ? 1.f : 0.f;
    lm_sum += scaled_bce(pm[mo + e], y, sm[mo + e]) * k;
    lm_cnt += k;
  }
  float ls = 0.f, le = 0.f, la = 0.f, lc = 0.f;
  for (int e = tid; e < L; e += blockDim.x) {
    const float k = lmask[lo + e] ? 1.f : 0.f;
    ls += scaled_bce(ps[lo + e], ys[lo + e] ? 1.f : 0.f, ss[lo + e]) * k;
    le += scaled_bce(pe[lo + e], ye[lo + e] ? 1.f : 0.f, se[lo + e]) * k;
    la += bce_elem(pa[lo + e], ya[lo + e] ? 1.f : 0.f) * k;
    lc += k;
  }
  lm_sum = block_sum(lm_sum, red); lm_cnt = block_sum(lm_cnt, red);
  ls = block_sum(ls, red); le = block_sum(le, red); la = block_sum(la, red); lc = block_sum(lc, red);
  if (tid == 0) {
    scratch[0 * B + b] = lm_sum / lm_cnt;
    scratch[1 * B + b] = ls / lc;
    scratch[2 * B + b] = le / lc;
    scratch[3 * B + b] = la / lc;
  }
  if (g_pm) {
    const float sc_m = 1.f / (lm_cnt * (float)B), sc_l = 1.f / (lc * (float)B);
    for (int e = tid; e < L * L; e += blockDim.x) {
      const float k = mmask[mo + e] ? 1.f : 0.f, y = ym[mo + e] ? 1.f : 0.f, s = sm[mo + e];
      g_pm[mo + e] = k != 0.f ? bce_grad(pm[mo + e], y, s * y + (1.f - s) * (1.f - y)) * sc_m : 0.f;
    }
    for (int e = tid; e < L; e += blockDim.x) {
      const float k = lmask[lo + e] ? 1.f : 0.f;
      const float y1 = ys[lo + e] ? 1.f : 0.f, y2 = ye[lo + e] ? 1.f : 0.f, y3 = ya[lo + e] ? 1.f : 0.f;
      const float s1 = ss[lo + e], s2 = se[lo + e];
      g_ps[lo + e] = k != 0.f ? bce_grad(ps[lo + e], y1, s1 * y1 + (1.f - s1) * (1.f - y1)) * sc_l : 0.f;
      g_pe[lo + e] = k != 0.f ? bce_grad(pe[lo + e], y2, s2 * y2 + (1.f - s2) * (1.f - y2)) * sc_l : 0.f;
      g_pa[lo + e] = k != 0.f ? 0.5f * bce_grad(pa[lo + e], y3, 1.f) * sc_l : 0.f;
    }
  }
}

__global__ void loss_finalize_kernel(const float* __restrict__ scratch, int B, float* __restrict__ loss, float* __restrict__ parts) {
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += scratch[threadIdx.x * B + b];
    parts[threadIdx.x] = s / (float)B;
  }
  __syncwarp();
  if (threadIdx.x == 0) loss[0] = ((parts[0] + parts[1]) + parts[2]) + 0.5f * parts[3];
}

int scaled_iou_bce(const float* pm, const uint8_t* ym, const float* sm, const uint8_t* mmask, const float* ps,
                   const uint8_t* ys, const float* ss, const float* pe, const uint8_t* ye, const float* se,
                   const float* pa, const uint8_t* ya, const uint8_t* lmask, int B, int L, float* loss, float* parts,
                   float* scratch, float* g_pm, float* g_ps, float* g_pe, float* g_pa, cudaStream_t st) {
  VML_CHECK_ARG(B > 0 && L > 0 && scratch != nullptr);
  VML_CHECK_ARG((g_pm == nullptr) == (g_ps == nullptr) && (g_pm == nullptr) == (g_pe == nullptr) && (g_pm == nullptr) == (g_pa == nullptr));
  static bool reg = (register_kernel("loss_sample_kernel"), register_kernel("loss_finalize_kernel"), true); (void)reg;
  loss_sample_kernel<<<B, 256, 0, st>>>(pm, ym, sm, mmask, ps, ys, ss, pe, ye, se, pa, ya, lmask, B, L, scratch, g_pm, g_ps, g_pe, g_pa);
  loss_finalize_kernel<<<1, 32, 0, st>>>(scratch, B, loss, parts);
  VML_LAUNCHED(2);
  return VML_OK;
}

// =====================================================================================
// a11  compute_ious (utils.py:10-31): score, top-k, IoU gather, R@n counts.
// One CTA per sample.  Scores are cached in shared memory; each of the k rounds is a block
// arg-max on 64-bit keys (score bits << 32 | ~flat index), i.e. ties -> lowest flat index;
// warp ballots compact the per-warp winners.  Optional greedy temporal NMS on the integer
// (i,j) grid: a picked proposal suppresses every proposal whose IoU with it is
// > nms_num/nms_den (exact integer arithmetic).
// =====================================================================================
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int o) {
  unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
  lo = __shfl_xor_sync(0xffffffffu, lo, o); hi = __shfl_xor_sync(0xffffffffu, hi, o);
  return ((unsigned long long)hi << 32) | lo;
}

// n in `ns` (top-n prefixes), m in `ms` (strict IoU thresholds): the reference's arbitrary lists (utils.py:10,24-29)
struct RecallSpec { int ns[8]; float ms[8]; int n_n, n_m; };
constexpr int kMaxTopK = 32;

__global__ void __launch_bounds__(256)
score_topk_kernel(const float* __restrict__ pm, const float* __restrict__ ps, const float* __restrict__ pe,
                  const uint8_t* __restrict__ mmask, const float* __restrict__ sm, int L, int k, int nms_num, int nms_den,
                  int32_t* __restrict__ top_idx, float* __restrict__ top_score, float* __restrict__ top_iou,
                  unsigned long long* __restrict__ counts, unsigned long long* __restrict__ counts2, int group,
                  const RecallSpec spec) {
  extern __shared__ __align__(8) unsigned char smem_raw[];
  float* sc = reinterpret_cast<float*>(smem_raw);                       // [L*L] scores; < 0 marks taken/suppressed
  __shared__ unsigned long long wbest[8];
  __shared__ int picked[kMaxTopK];
  __shared__ float picked_iou[kMaxTopK];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
  const int n = L * L;
  const size_t mo = (size_t)b * n;
  for (int e = tid; e < n; e += blockDim.x) {
    const int i = e / L, j = e % L;
    // same op order as utils.py:17-19 (fp32, IEEE sqrt): ((pm*sqrt(ps_i))*sqrt(pe_j))*mask
    float s = __fmul_rn(__fmul_rn(pm[mo + e], __fsqrt_rn(ps[(size_t)b * L + i])), __fsqrt_rn(pe[(size_t)b * L + j]));
    s = __fmul_rn(s, mmask[mo + e] ? 1.f : 0.f);
    sc[e] = s;
  }
  __syncthreads();
  const bool use_nms = nms_num < nms_den;
  for (int r = 0; r < k; ++r) {
    unsigned long long best = 0ull;
    bool any = false;
    for (int e = tid; e < n; e += blockDim.x) {
      const float s = sc[e];
      if (s >= 0.f) {  // alive (scores are >= +0)
        const unsigned long long key = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned)(0xffffffffu - (unsigned)e);
        if (!any || key > best) best = key;
        any = true;
      }
    }
    // ballot: which lanes hold a candidate at all
    const unsigned have = __ballot_sync(0xffffffffu, any);
    if (!any) best = 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long ot = shfl_xor_u64(best, o); best = ot > best ? ot : best; }
    if (lane == 0) wbest[warp] = have ? (best | 1ull << 63) : 0ull;   // bit63: valid flag (scores < 2^31 as bits)
    __syncthreads();
    if (tid == 0) {
      unsigned long long m = 0ull;
      for (int w = 0; w < (int)(blockDim.x / 32); ++w) m = wbest[w] > m ? wbest[w] : m;
      if (m == 0ull) { picked[r] = -1; picked_iou[r] = 0.f; }
      else {
        const int e = (int)(0xffffffffu - (unsigned)(m & 0xffffffffull));
        picked[r] = e;
        picked_iou[r] = sm[mo + e];
        top_idx[(size_t)b * k + r] = e;
        top_score[(size_t)b * k + r] = sc[e];
        top_iou[(size_t)b * k + r] = picked_iou[r];
      }
      if (m == 0ull) { top_idx[(size_t)b * k + r] = -1; top_score[(size_t)b * k + r] = 0.f; top_iou[(size_t)b * k + r] = 0.f; }
    }
    __syncthreads();
    const int pe_ = picked[r];
    if (pe_ < 0) continue;
    if (tid == 0) sc[pe_] = -1.f;
    if (use_nms) {
      const int bi = pe_ / L, bj = pe_ % L;
      for (int e = tid; e < n; e += blockDim.x) {
        const int i = e / L, j = e % L;
        const int inter = max(0, min(j, bj) + 1 - max(i, bi));
        const int uni = max(j, bj) + 1 - min(i, bi);
        if (uni > 0 && (long long)inter * nms_den > (long long)nms_num * uni) sc[e] = -1.f;
      }
    }
    __syncthreads();
  }
  if (tid < spec.n_n * spec.n_m) {
    const int a = tid / spec.n_m, t = tid % spec.n_m, per = spec.n_n * spec.n_m;
    bool hit = false;
    for (int r = 0; r < min(spec.ns[a], k); ++r) hit = hit || (picked[r] >= 0 && picked_iou[r] > spec.ms[t]);
    if (hit) {
      atomicAdd(&counts[tid], 1ull);
      if (counts2) atomicAdd(&counts2[(b / group) * per + tid], 1ull);
    }
  }
}

int score_topk_recall(const float* pm, const float* ps, const float* pe, const uint8_t* mmask, const float* sm, int B,
                      int L, int k, int nms_num, int nms_den, int32_t* top_idx, float* top_score, float* top_iou,
                      int64_t* counts, int64_t* counts2, int group, const int* ns, int n_n, const float* ms, int n_m,
                      cudaStream_t st) {
  VML_CHECK_ARG(B > 0 && L > 0 && k >= 1 && k <= kMaxTopK && k <= L * L && nms_den > 0 && (size_t)L * L * 4 <= 200 * 1024);
  RecallSpec spec{};
  if (ns == nullptr || ms == nullptr) {        // the reference's defaults (utils.py:10)
    spec.n_n = 2; spec.n_m = 4;
    spec.ns[0] = 1; spec.ns[1] = 5;
    spec.ms[0] = 0.1f; spec.ms[1] = 0.3f; spec.ms[2] = 0.5f; spec.ms[3] = 0.7f;
  } else {
    VML_CHECK_ARG(n_n >= 1 && n_n <= 8 && n_m >= 1 && n_m <= 8);
    spec.n_n = n_n; spec.n_m = n_m;
    for (int a = 0; a < n_n; ++a) { VML_CHECK_ARG(ns[a] >= 1); spec.ns[a] = ns[a]; }
    for (int t = 0; t < n_m; ++t) spec.ms[t] = ms[t];
  }
  static bool reg = (register_kernel("score_topk_kernel"), true); (void)reg;
  const size_t smem = sizeof(float) * L * L;
  VML_CUDA(ensure_dyn_smem((const void*)(score_topk_kernel), (size_t)((int)smem)));
  score_topk_kernel<<<B, 256, smem, st>>>(pm, ps, pe, mmask, sm, L, k, nms_num, nms_den, top_idx, top_score, top_iou,
                                          (unsigned long long*)counts, (unsigned long long*)counts2, group > 0 ? group : B,
                                          spec);
  VML_LAUNCHED(1);
  return VML_OK;
}

}  // namespace vml
