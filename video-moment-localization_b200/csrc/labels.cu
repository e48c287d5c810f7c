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

}  // namespace vml
