// Labels and masks of a collated batch, generated on the device from the annotation scalars
// (reference: dataset.py:95-127 get_iou / get_boundary_penalties / get_snippet_label and
// dataset.py:139-158 masks + thresholds).  Replaces 10 of the 13 host tensors main.py:118-133 copies per step.
//
// Arithmetic mirrors the reference's float32 tensor ops one by one (Python-float scalars enter float32 tensor ops
// as float32; sigma and 2*sigma^2 are Python doubles), with explicit round-to-nearest intrinsics so that no FMA
// contraction changes a rounding: sm, ym, ya and the masks are bit-exact; ss / se differ from the CPU only by the
// exp implementation (<= 2 ulp).
#include "common.cuh"

namespace vml {

__global__ void __launch_bounds__(256)
make_labels_kernel(const double* __restrict__ times, const double* __restrict__ duration, const int64_t* __restrict__ nfeats,
                   int T, int L, float* __restrict__ sm, uint8_t* __restrict__ ym, float* __restrict__ ss,
                   uint8_t* __restrict__ ys, float* __restrict__ se, uint8_t* __restrict__ ye, uint8_t* __restrict__ ya,
                   uint8_t* __restrict__ lmask, uint8_t* __restrict__ mmask, uint8_t* __restrict__ vmask) {
  const int b = blockIdx.x;
  const double ts_d = times[2 * b], te_d = times[2 * b + 1];
  const float dur = (float)duration[b], gs = (float)ts_d, ge = (float)te_d, Lf = (float)L;
  const double sigma = (te_d - ts_d) / 5.0;                       // dataset.py:116 (Python floats)
  const float den = (float)(2.0 * (sigma * sigma));               // 2.0 * sigma**2, then a float32 operand
  const int n = (int)nfeats[b], r = T / L;
  const int len = min(L, (n + r - 1) / r);                        // ceil(nfeats / (T / L)), dataset.py:146
  // s_times[l] = (float(l) * duration) / L ; e_times[l] = (float(l + 1) * duration) / L    (dataset.py:96-97)
  auto s_time = [&](int l) { return __fdiv_rn(__fmul_rn((float)l, dur), Lf); };
  auto e_time = [&](int l) { return __fdiv_rn(__fmul_rn((float)(l + 1), dur), Lf); };
  for (int idx = threadIdx.x; idx < L * L; idx += blockDim.x) {
    const int i = idx / L, j = idx % L;
    const float ps = s_time(i), pe = e_time(j);
    const float inter = fmaxf(0.0f, __fsub_rn(fminf(pe, ge), fmaxf(ps, gs)));
    const float hull = fmaxf(0.0f, __fsub_rn(fmaxf(pe, ge), fminf(ps, gs)));
    const float iou = __fdiv_rn(inter, hull);
    const size_t o = (size_t)b * L * L + idx;
    if (sm) sm[o] = iou;
    if (ym) ym[o] = iou > 0.5f;
    if (mmask) mmask[o] = (i <= j && j < len) ? 1 : 0;
  }
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float st = s_time(l), et = e_time(l);
    const float ds = __fsub_rn(st, gs), de = __fsub_rn(et, ge);
    const float vs = expf(__fdiv_rn(-__fmul_rn(ds, ds), den));    // exp(-(t - tau)**2 / (2 sigma^2)), dataset.py:118-119
    const float ve = expf(__fdiv_rn(-__fmul_rn(de, de), den));
    const size_t o = (size_t)b * L + l;
    if (ss) ss[o] = vs;
    if (ys) ys[o] = vs > 0.5f;
    if (se) se[o] = ve;
    if (ye) ye[o] = ve > 0.5f;
    if (ya) ya[o] = (st >= gs && et <= ge) ? 1 : 0;
    if (lmask) lmask[o] = l < len ? 1 : 0;
  }
  if (vmask)
    for (int t = threadIdx.x; t < T; t += blockDim.x) vmask[(size_t)b * T + t] = t < n ? 1 : 0;
}

int make_labels(const double* times, const double* duration, const int64_t* nfeats, int B, int T, int L, float* sm, uint8_t* ym,
                float* ss, uint8_t* ys, float* se, uint8_t* ye, uint8_t* ya, uint8_t* lmask, uint8_t* mmask, uint8_t* vmask,
                cudaStream_t st) {
  VML_CHECK_ARG(times && duration && nfeats && B >= 0 && L > 0 && T % L == 0);
  static bool reg = (register_kernel("make_labels_kernel"), true); (void)reg;
  if (B == 0) return VML_OK;
  make_labels_kernel<<<B, 256, 0, st>>>(times, duration, nfeats, T, L, sm, ym, ss, ys, se, ye, ya, lmask, mmask, vmask);
  VML_LAUNCHED(1);
  return VML_OK;
}

// ---- fixed-length clip sampling (dataset.py:40-74 get_fixed_length_features) ---------------------------------------
// frame_idx = np.round(np.arange(spos, nfeats - 0.5, stride)).astype(int), stride = 1 or nfeats / T; numpy builds arange
// values as start + i * step in double and its length as ceil((stop - start) / step); np.round is round-half-to-even
// (rint).  The list is cut to T when its length fits neither nfeats (< T) nor T; a length that still does not fit makes
// the reference raise an AssertionError -> here: status bit 1, the sample's clips are zeros.
struct ClipPlan { int n_idx; int n_out; double stride; double spos; };

__device__ __forceinline__ ClipPlan clip_plan(long long nfeats, int T, int spos) {
  ClipPlan p;
  p.stride = nfeats <= T ? 1.0 : (double)nfeats * 1.0 / (double)T;
  p.spos = (double)spos;
  const double span = ((double)nfeats - 0.5) - p.spos;
  long long n = span > 0.0 ? (long long)ceil(span / p.stride) : 0;
  const bool fits = (nfeats < T && n == nfeats) || (nfeats >= T && n == T);
  if (!fits && n > T) n = T;                                     // frame_idx[:T]  ("ignore last one")
  p.n_idx = (int)n;
  const bool ok = (nfeats < T && n == nfeats) || (nfeats >= T && n == T);
  p.n_out = ok ? (int)(nfeats < T ? nfeats : T) : -1;
  return p;
}
__device__ __forceinline__ long long clip_frame(const ClipPlan& p, int i) { return (long long)rint(p.spos + (double)i * p.stride); }

__global__ void __launch_bounds__(128)
sample_index_kernel(const int64_t* __restrict__ offsets, const int32_t* __restrict__ spos, const double* __restrict__ start_pos,
                    const double* __restrict__ end_pos, int T, int64_t* __restrict__ nfeats_out, int32_t* __restrict__ start_index,
                    int32_t* __restrict__ end_index, int32_t* __restrict__ status) {
  const int b = blockIdx.x;
  const long long nfeats = offsets[b + 1] - offsets[b];
  const ClipPlan p = clip_plan(nfeats, T, spos ? spos[b] : 0);
  __shared__ int s_start, s_end;
  if (threadIdx.x == 0) { s_start = 0; s_end = T - 1; }           // dataset.py:57
  __syncthreads();
  if (p.n_out < 0) {
    if (threadIdx.x == 0) { atomicOr(status, 2); nfeats_out[b] = 0; start_index[b] = 0; end_index[b] = T - 1; }
    return;
  }
  const double sp = ((double)nfeats - 1.0) * (start_pos ? start_pos[b] : 0.0);      // float(nfeats - 1.0) * start_pos: Python doubles
  const double ep = ((double)nfeats - 1.0) * (end_pos ? end_pos[b] : 0.0);
  for (int i = threadIdx.x; i + 1 < p.n_idx; i += blockDim.x) {   // frame_idx is strictly increasing: at most one hit each
    const double f0 = (double)clip_frame(p, i), f1 = (double)clip_frame(p, i + 1);
    if (f0 <= ep && ep < f1) s_end = i;
    if (f0 <= sp && sp < f1) s_start = i;
  }
  __syncthreads();
  if (threadIdx.x == 0) { nfeats_out[b] = p.n_out; start_index[b] = s_start; end_index[b] = s_end; }
}

__global__ void __launch_bounds__(256)
sample_gather_kernel(const float* __restrict__ raw, const int64_t* __restrict__ offsets, const int32_t* __restrict__ spos, int T,
                     int d0, float* __restrict__ out, uint8_t* __restrict__ vmask) {
  const int t = blockIdx.x, b = blockIdx.y;
  const long long base = offsets[b], nfeats = offsets[b + 1] - base;
  const ClipPlan p = clip_plan(nfeats, T, spos ? spos[b] : 0);
  const bool live = p.n_out >= 0 && t < p.n_out;
  float* dst = out + ((size_t)b * T + t) * d0;
  if (vmask && threadIdx.x == 0) vmask[(size_t)b * T + t] = live ? 1 : 0;
  const float* src = live ? raw + (size_t)(base + clip_frame(p, t)) * d0 : nullptr;
  if ((d0 & 3) == 0 && ((reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = threadIdx.x; e < d0 / 4; e += blockDim.x)
      reinterpret_cast<float4*>(dst)[e] = live ? __ldg(reinterpret_cast<const float4*>(src) + e) : z;
  } else {
    for (int e = threadIdx.x; e < d0; e += blockDim.x) dst[e] = live ? __ldg(src + e) : 0.f;
  }
}

int sample_clips(const float* raw, const int64_t* offsets, const int32_t* spos, const double* start_pos, const double* end_pos,
                 int B, int T, int d0, float* out, uint8_t* vmask, int64_t* nfeats_out, int32_t* start_index, int32_t* end_index,
                 int32_t* status, cudaStream_t st) {
  VML_CHECK_ARG(raw && offsets && out && nfeats_out && start_index && end_index && status && B >= 0 && T > 0 && d0 > 0 && B < 65536);
  static bool reg = (register_kernel("sample_index_kernel"), register_kernel("sample_gather_kernel"), true); (void)reg;
  if (B == 0) return VML_OK;
  sample_index_kernel<<<B, 128, 0, st>>>(offsets, spos, start_pos, end_pos, T, nfeats_out, start_index, end_index, status);
  sample_gather_kernel<<<dim3(T, B), 256, 0, st>>>(raw, offsets, spos, T, d0, out, vmask);
  VML_LAUNCHED(2);
  return VML_OK;
}

}  // namespace vml
