// extern "C" boundary of libvml_b200.so (see include/vml_b200.h) and the GEMM-backed stages.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>

#include "common.cuh"
#include "epilogues.cuh"
#include "gemm_simt.cuh"
#include "gemm_strided.cuh"
#include "gemm_umma.cuh"

namespace vml {

// ---- error / registry ---------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static std::mutex g_reg_mu;
static std::string g_kernels;
static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
cudaError_t ensure_dyn_smem(const void* func, size_t bytes) {
  static std::mutex mu;
  static std::unordered_map<const void*, size_t> cur;
  std::lock_guard<std::mutex> lk(mu);
  size_t& c = cur[func];
  if (bytes <= c) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) c = bytes;
  return e;
}
void register_kernel(const char* name) {
  std::lock_guard<std::mutex> lk(g_reg_mu);
  if (g_kernels.find(std::string(name) + "\n") == std::string::npos) g_kernels += std::string(name) + "\n";
}

// ---- TMA descriptor ------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t k, uint64_t row_stride_elems, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return VML_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_stride_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)UG_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d (rows=%llu k=%llu ld=%llu)", (int)r,
                                     (unsigned long long)rows, (unsigned long long)k, (unsigned long long)row_stride_elems);
    return VML_ERR_CUDA; }
  return VML_OK;
}

// generic 2D map with 128B swizzle: dtype 0 = bf16, 1 = fp32 rounded to TF32 by the copy engine (plain fp32 if the driver
// refuses that type); inner = contiguous dimension
int make_tmap_2d(CUtensorMap* map, int dtype, const void* ptr, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer, int swizzle) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return VML_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)outer_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType types[3] = {CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32};
  CUresult r = CUDA_ERROR_UNKNOWN;
  for (int t = dtype; t < (dtype == 1 ? 3 : dtype + 1); ++t) {
    r = enc(map, types[t], 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) break;
  }
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d (inner=%llu outer=%llu stride=%llu B)", (int)r,
                                     (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)outer_stride_bytes);
    return VML_ERR_CUDA; }
  return VML_OK;
}
int launch_gemm_tf32(const float*, int64_t, int64_t, const float*, int64_t, int64_t, float*, int64_t, int64_t, int, int, int, float, int,
                     const int32_t*, int, const int32_t*, int, cudaStream_t);

// stage launchers implemented in stages.cu / query_loss_eval.cu
int build_cells(const uint8_t*, int, int, vml_cells_t, cudaStream_t);
int unpack_cells(const void*, void*, vml_cells_t, int, int, int, int, cudaStream_t);
int pack_cells(const void*, void*, vml_cells_t, int, int, int, int, cudaStream_t);
int cast_pad(const float*, void*, int64_t, int, int, cudaStream_t);
int ingest(const void*, const void*, int, const int64_t*, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, const float*, void*,
           void*, uint8_t*, uint8_t*, uint8_t*, uint8_t*, float*, int32_t*, int, vml_dims_t, int, int, int, cudaStream_t);
int span_pool_fuse(const void*, const float*, vml_cells_t, void*, void*, float*, int, vml_dims_t, int, cudaStream_t);
int content_attention(const void*, const float*, int, int, int, int, const float*, int, const uint8_t*, vml_cells_t,
                      void*, int, vml_dims_t, int, cudaStream_t);
int boundary_unit(const float*, int, int, int, const float*, const float*, const float*, const void*,
                  const uint8_t*, const uint8_t*, vml_cells_t, float*, float*, float*, void*, const float*, float*, float*, int,
                  vml_dims_t, int, cudaStream_t, void* pair_out = nullptr, int ld_pair = 0);
bool boundary_pair_fused(vml_dims_t, int);
// backward.cu
int colsum(const float*, int64_t, int64_t, float*, int64_t, int, int, int, const int32_t*, int, float, cudaStream_t);
int localize_bwd(const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*,
                 const float*, const float*, const float*, const uint8_t*, vml_cells_t, float*, float*, float*, float*, int, vml_dims_t,
                 cudaStream_t);
int pair_bwd(const float*, int, const float*, vml_cells_t, float*, int, vml_dims_t, cudaStream_t);
int cu_tail_bwd(const float*, const float*, int, vml_cells_t, float*, float*, vml_dims_t, cudaStream_t);
int content_attn_bwd(const float*, const float*, const float*, int, int, int, int, const float*, int, const uint8_t*, vml_cells_t,
                     float*, float*, float*, int, vml_dims_t, cudaStream_t);
int gbar_bwd(const float*, const float*, const float*, const float*, const float*, const float*, vml_cells_t, float*, float*, float*,
             int, vml_dims_t, cudaStream_t);
int softmax_bwd(const float*, const float*, const uint8_t*, float*, int, int, int, float, cudaStream_t);
int gate_bwd(const float*, const float*, const float*, const uint8_t*, float*, float*, float*, int, vml_dims_t, cudaStream_t);
int mask_rows(const float*, const uint8_t*, float*, int64_t, int, int, cudaStream_t);
int span_pool_bwd(const float*, const float*, const float*, const float*, const float*, vml_cells_t, float*, float*, int, vml_dims_t,
                  cudaStream_t);
int adam_step(float*, const float*, float*, float*, int64_t, float, float, float, float, int, float, cudaStream_t);
int lstm_train_fwd(const float*, const float*, const int32_t*, float*, float*, float*, int, int, int, cudaStream_t);
int lstm_train_bwd(const float*, const float*, const float*, const float*, const int32_t*, float*, float*, int, int, int, cudaStream_t);
int moment_pair(const float*, vml_cells_t, void*, vml_dims_t, int, cudaStream_t);
int gemm_res(const void*, const void*, const float*, const void*, const void*, void*, void*, int, int, int, int, int,
             const int32_t*, int, cudaStream_t);
int content_tc(const void*, const void*, const float*, const float*, int, int, int, int, const float*, int, const uint8_t*,
               vml_cells_t, void*, int, vml_dims_t, cudaStream_t);
int content_unit(const void*, const void*, const float*, const float*, int, int, int, int, const float*, int, const uint8_t*,
                 vml_cells_t, const void*, const float*, const void*, void*, void*, int, int, vml_dims_t, int, int, cudaStream_t);
bool content_unit_supported(vml_dims_t);
int sample_clips(const float*, const int64_t*, const int32_t*, const double*, const double*, int, int, int, float*, uint8_t*,
                 int64_t*, int32_t*, int32_t*, int32_t*, cudaStream_t);
int make_labels(const double*, const double*, const int64_t*, int, int, int, float*, uint8_t*, float*, uint8_t*, float*, uint8_t*,
                uint8_t*, uint8_t*, uint8_t*, uint8_t*, cudaStream_t);
int moment_operand(const void*, const float*, vml_cells_t, void*, vml_dims_t, int, cudaStream_t);
int localize(const void*, const float*, const float*, const float*, vml_cells_t, const uint8_t*, float*, float*, float*,
             float*, int, vml_dims_t, int, cudaStream_t);
int query_lengths(const uint8_t*, int32_t*, int, int, cudaStream_t);
int lstm_layer(const float*, const float*, const int32_t*, float*, void*, float*, void*, int, int, int, cudaStream_t, float* acts = nullptr);
int lstm_bwd_cluster(const float*, const float*, const float*, const float*, const int32_t*, float*, float*, int, int, int, cudaStream_t);
int lstm_layer_tc(const float*, const void*, const int32_t*, float*, void*, float*, void*, int, int, int, cudaStream_t);
int scaled_iou_bce(const float*, const uint8_t*, const float*, const uint8_t*, const float*, const uint8_t*, const float*,
                   const float*, const uint8_t*, const float*, const float*, const uint8_t*, const uint8_t*, int, int,
                   float*, float*, float*, float*, float*, float*, float*, cudaStream_t);
int score_topk_recall(const float*, const float*, const float*, const uint8_t*, const float*, int, int, int, int, int,
                      int32_t*, float*, float*, int64_t*, int64_t*, int, const int*, int, const float*, int, cudaStream_t);

// generic dispatch: fp32 -> CUDA-core GEMM, bf16 -> tcgen05 GEMM
template <typename Epi>
static int gemm_dispatch(const void* A, const void* W, int M, int N, int K, int lda, const int32_t* m_dev, int m_scale,
                         const Epi& epi, int prec, cudaStream_t st) {
  if (prec == VML_BF16) return launch_gemm_umma(A, W, M, N, K, lda, K, m_dev, m_scale, epi, st);
  // VML_TF32 (training forward): fp32 tensors as they are, tcgen05 kind::tf32; small or oddly strided products stay on CUDA cores
  if (prec == VML_TF32 && lda % 4 == 0 && K % 4 == 0 && N % 32 == 0 && (double)M * N * K >= 5.0e7 &&
      ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) & 15) == 0)
    return launch_gemm_umma_tf32(A, W, M, N, K, lda, K, m_dev, m_scale, epi, st);
  VML_CHECK_ARG(lda == K);
  return launch_gemm_simt((const float*)A, (const float*)W, M, N, K, m_dev, m_scale, epi, st);
}

}  // namespace vml

using namespace vml;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define VML_PREC_OK(p) VML_CHECK_ARG((p) == VML_FP32 || (p) == VML_BF16 || (p) == VML_TF32)

extern "C" {

VML_API const char* vml_last_error(void) { return g_err; }
VML_API int vml_version(void) { return 1; }
VML_API int64_t vml_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }
VML_API const char* vml_kernel_names(void) {
  static thread_local std::string copy;
  std::lock_guard<std::mutex> lk(g_reg_mu);
  copy = g_kernels;
  return copy.c_str();
}

VML_API int vml_build_cells(const uint8_t* moment_mask, int B, int L, vml_cells_t cells, void* stream) {
  return build_cells(moment_mask, B, L, cells, ST(stream));
}
VML_API int vml_unpack_cells(const void* packed, void* dense, vml_cells_t cells, int B, int L, int inner, int prec, void* stream) {
  VML_PREC_OK(prec);
  return unpack_cells(packed, dense, cells, B, L, inner, prec, ST(stream));
}
VML_API int vml_pack_cells(const void* dense, void* packed, vml_cells_t cells, int B, int L, int inner, int prec, void* stream) {
  VML_PREC_OK(prec);
  return pack_cells(dense, packed, cells, B, L, inner, prec, ST(stream));
}
VML_API int vml_cast_pad_bf16(const float* src, void* dst, int64_t rows, int k, int k_pad, void* stream) {
  return cast_pad(src, dst, rows, k, k_pad, ST(stream));
}

VML_API int vml_ingest(const float* video_features, const float* query_features, const uint8_t* video_mask,
                       const uint8_t* query_mask, const uint8_t* length_mask, const uint8_t* moment_mask, const float* sm,
                       void* v_out, void* q_out, uint8_t* vmask_out, uint8_t* qmask_out, uint8_t* lmask_out,
                       uint8_t* mmask_out, float* sm_out, int32_t* qlen, int B, vml_dims_t d, int v_kpad, int q_kpad,
                       int prec, void* stream) {
  VML_PREC_OK(prec);
  return ingest(video_features, query_features, 0, nullptr, video_mask, query_mask, length_mask, moment_mask, sm, v_out, q_out, vmask_out,
                qmask_out, lmask_out, mmask_out, sm_out, qlen, B, d, v_kpad, q_kpad, prec, ST(stream));
}

VML_API int vml_ingest_bf16(const void* video_features, const void* query_features, const uint8_t* video_mask,
                    const uint8_t* query_mask, const uint8_t* length_mask, const uint8_t* moment_mask, const float* sm,
                    void* v_out, void* q_out, uint8_t* vmask_out, uint8_t* qmask_out, uint8_t* lmask_out,
                    uint8_t* mmask_out, float* sm_out, int32_t* qlen, int B, vml_dims_t d, int v_kpad, int q_kpad,
                    int prec, void* stream) {
  VML_PREC_OK(prec);
  return ingest(video_features, query_features, 1, nullptr, video_mask, query_mask, length_mask, moment_mask, sm, v_out, q_out, vmask_out,
                qmask_out, lmask_out, mmask_out, sm_out, qlen, B, d, v_kpad, q_kpad, prec, ST(stream));
}

VML_API int vml_ingest_packed(const void* video_rows, const void* query_features, const uint8_t* video_mask,
                              const uint8_t* query_mask, const uint8_t* length_mask, const uint8_t* moment_mask, const float* sm,
                              const int64_t* nfeats, void* v_out, void* q_out, uint8_t* vmask_out, uint8_t* qmask_out,
                              uint8_t* lmask_out, uint8_t* mmask_out, float* sm_out, int32_t* qlen, int B, vml_dims_t d,
                              int v_kpad, int q_kpad, int prec, int src_bf16, void* stream) {
  VML_PREC_OK(prec);
  if (nfeats == nullptr) { set_error("vml_ingest_packed: nfeats is NULL"); return VML_ERR_ARG; }
  return ingest(video_rows, query_features, src_bf16, nfeats, video_mask, query_mask, length_mask, moment_mask, sm, v_out,
                q_out, vmask_out, qmask_out, lmask_out, mmask_out, sm_out, qlen, B, d, v_kpad, q_kpad, prec, ST(stream));
}

VML_API int vml_gemm_strided(const float* A, int64_t sam, int64_t sak, int64_t sab, const float* B, int64_t sbn, int64_t sbk,
                             int64_t sbb, float* C, int64_t scm, int64_t scn, int64_t scb, int M, int N, int K, int batch,
                             float alpha, int accumulate, int splits, const int32_t* m_dev, int m_scale,
                             const int32_t* k_dev, int k_scale, void* stream) {
  if (batch == 1) {
    // large products: tcgen05 kind::tf32 straight from the caller's fp32 tensors (gemm_tf32.cu); 1 = not eligible
    const int rc = launch_gemm_tf32(A, sam, sak, B, sbn, sbk, C, scm, scn, M, N, K, alpha, accumulate, m_dev, m_scale, k_dev, k_scale,
                                    ST(stream));
    if (rc <= 0) return rc;
  }
  SGemm g{A, sam, sak, sab, B, sbn, sbk, sbb, C, scm, scn, scb, M, N, K, batch, alpha, accumulate, splits, m_dev, m_scale, k_dev, k_scale};
  return launch_gemm_strided(g, ST(stream));
}

VML_API int vml_linear(const void* A, const void* W, const float* bias, void* out, int M, int N, int K, int ldo,
               const int32_t* m_dev, int m_scale, int prec, int out_fp32, void* stream) {
  VML_PREC_OK(prec);
  VML_CHECK_ARG(ldo >= N && ldo % 4 == 0 && m_scale >= 1);
  if (prec == VML_BF16 && !out_fp32) {
    VML_CHECK_ARG(ldo % 8 == 0);            // 16-byte bf16 stores in the epilogue
    EpiBias<bf16> e{bias, (bf16*)out, ldo};
    return gemm_dispatch(A, W, M, N, K, K, m_dev, m_scale, e, prec, ST(stream));
  }
  if (prec == VML_BF16 && N % 32 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
      (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0) && getenv("VML_EPI_DIRECT") == nullptr) {
    EpiBiasT e{bias, (float*)out, ldo};          // fp32 result of the tcgen05 GEMM: coalesced stores (epilogues.cuh)
    return launch_gemm_umma(A, W, M, N, K, K, K, m_dev, m_scale, e, ST(stream));
  }
  EpiBias<float> e{bias, (float*)out, ldo};
  return gemm_dispatch(A, W, M, N, K, K, m_dev, m_scale, e, prec, ST(stream));
}

VML_API int vml_clip_projection(const void* v, const void* W, const float* bias, const float* pe, const uint8_t* video_mask,
                        void* fv, int B, vml_dims_t d, int k_pad, int prec, void* stream) {
  VML_PREC_OK(prec);
  const int M = B * d.T;
  if (prec == VML_BF16) {
    if (getenv("VML_EPI_DIRECT") == nullptr) {
      EpiClipT e{bias, pe, video_mask, d.T, (bf16*)fv, d.D, M};
      return launch_gemm_umma(v, W, M, d.D, k_pad, k_pad, k_pad, nullptr, 1, e, ST(stream));
    }
    EpiClipPre e{bias, pe, video_mask, d.T, (bf16*)fv, d.D};
    return launch_gemm_umma(v, W, M, d.D, k_pad, k_pad, k_pad, nullptr, 1, e, ST(stream));
  }
  EpiClip<float> e{bias, pe, video_mask, d.T, (float*)fv, d.D};
  return gemm_dispatch(v, W, M, d.D, d.d0, d.d0, nullptr, 1, e, prec, ST(stream));
}

VML_API int vml_lstm_layer(const float* gin, const float* whh_t, const int32_t* qlen, float* y, void* y_bf16, float* fs,
                   void* fs_bf16, int B, int Nq, int H, void* stream) {
  return lstm_layer(gin, whh_t, qlen, y, y_bf16, fs, fs_bf16, B, Nq, H, ST(stream));
}
VML_API int vml_lstm_layer_tc(const float* gin, const void* whh_frag, const int32_t* qlen, float* y, void* y_bf16, float* fs,
                              void* fs_bf16, int B, int Nq, int H, void* stream) {
  return lstm_layer_tc(gin, whh_frag, qlen, y, y_bf16, fs, fs_bf16, B, Nq, H, ST(stream));
}
VML_API int vml_query_lengths(const uint8_t* query_mask, int32_t* qlen, int B, int Nq, void* stream) {
  return query_lengths(query_mask, qlen, B, Nq, ST(stream));
}
VML_API int vml_span_pool_fuse(const void* fv, const float* fs, vml_cells_t cells, void* fc, void* fm, float* fb, int B,
                       vml_dims_t d, int prec, void* stream) {
  VML_PREC_OK(prec);
  return span_pool_fuse(fv, fs, cells, fc, fm, fb, B, d, prec, ST(stream));
}

VML_API int vml_content_attention(const void* c_hat, const float* qproj, int ld, int off_what, int off_ktil, int off_beta,
                          const float* s_hat, int s_ld, const uint8_t* query_mask, vml_cells_t cells, void* cc_hat, int B,
                          vml_dims_t d, int prec, void* stream) {
  VML_PREC_OK(prec);
  return content_attention(c_hat, qproj, ld, off_what, off_ktil, off_beta, s_hat, s_ld, query_mask, cells, cc_hat, B, d, prec,
                           ST(stream));
}

VML_API int vml_content_in_attention(const void* fc, const void* W, const float* bias, const float* qproj, int ld, int off_what,
                             int off_ktil, int off_beta, const float* s_hat, int s_ld, const uint8_t* query_mask,
                             vml_cells_t cells, void* cc_hat, int B, vml_dims_t d, void* stream) {
  return content_tc(fc, W, bias, qproj, ld, off_what, off_ktil, off_beta, s_hat, s_ld, query_mask, cells, cc_hat, B, d,
                    ST(stream));
}

VML_API int vml_copy_h2d_async(void* dst, const void* src_pinned, int64_t bytes, void* stream) {
  VML_CHECK_ARG(dst && src_pinned && bytes >= 0);
  VML_CUDA(cudaMemcpyAsync(dst, src_pinned, (size_t)bytes, cudaMemcpyHostToDevice, ST(stream)));
  return VML_OK;
}

VML_API int vml_sample_clips(const float* raw, const int64_t* offsets, const int32_t* spos, const double* start_pos,
                             const double* end_pos, int B, int T, int d0, float* video_features, uint8_t* video_mask,
                             int64_t* nfeats, int32_t* start_index, int32_t* end_index, int32_t* status, void* stream) {
  return sample_clips(raw, offsets, spos, start_pos, end_pos, B, T, d0, video_features, video_mask, nfeats, start_index,
                      end_index, status, ST(stream));
}

VML_API int vml_make_labels(const double* times, const double* duration, const int64_t* nfeats, int B, int T, int L, float* sm,
                    uint8_t* ym, float* ss, uint8_t* ys, float* se, uint8_t* ye, uint8_t* ya, uint8_t* length_mask,
                    uint8_t* moment_mask, uint8_t* video_mask, void* stream) {
  return make_labels(times, duration, nfeats, B, T, L, sm, ym, ss, ys, se, ye, ya, length_mask, moment_mask, video_mask, ST(stream));
}

VML_API int vml_content_unit_supported(vml_dims_t d) { return content_unit_supported(d) ? 1 : 0; }

VML_API int vml_content_unit(const void* fc, const void* W_chat, const float* b_chat, const float* qproj, int ld, int off_what,
                     int off_ktil, int off_beta, const float* s_hat, int s_ld, const uint8_t* query_mask, vml_cells_t cells,
                     const void* Wc, const float* bc, const void* fbar, void* cu, void* mu_operand, int B, vml_dims_t d,
                     int store_cu, void* stream) {
  VML_CHECK_ARG(fc && W_chat && Wc && fbar && cu && mu_operand);
  // bc == NULL: the output bias is already inside fbar (vml_boundary_unit's fbar_bias) -> the kernel variant whose residual
  // add runs on the tensor cores and whose epilogue only adds fbar
  return content_unit(fc, W_chat, b_chat, qproj, ld, off_what, off_ktil, off_beta, s_hat, s_ld, query_mask, cells, Wc, bc, fbar, cu,
                      (bf16*)mu_operand + d.D, 2 * d.D, B, d, store_cu, bc == nullptr ? 1 : 0, ST(stream));
}

VML_API int vml_content_out(const void* cc_hat, const void* Wc, const float* bc, const void* fc, const void* fm, const float* fs,
                    const void* fbar, void* mu_operand, vml_cells_t cells, void* cu, vml_dims_t d, int prec, void* stream) {
  VML_PREC_OK(prec);
  const int M = cells.capacity * d.C;
  if (prec == VML_BF16 && fbar && mu_operand) {
    VML_CHECK_ARG(d.C == 4 && d.D % 32 == 0);
    if (d.D % 128 == 0)   // residual tiles through shared memory (TMA in / TMA out)
      return gemm_res(cc_hat, Wc, bc, fc, fbar, cu, (bf16*)mu_operand + d.D, M, d.D, d.dl, d.dl, 2 * d.D, cells.n_cells, d.C,
                      ST(stream));
    EpiContentOutFused e{bc, (const bf16*)fc, (const bf16*)fbar, (bf16*)cu, (bf16*)mu_operand, d.D};
    return launch_gemm_umma(cc_hat, Wc, M, d.D, d.dl, d.dl, d.dl, cells.n_cells, d.C, e, ST(stream));
  }
  if (prec == VML_BF16) {
    EpiContentOut<bf16> e{bc, (const bf16*)fc, (const bf16*)fm, fs, cells.code, d.C, (bf16*)cu, d.D};
    return gemm_dispatch(cc_hat, Wc, M, d.D, d.dl, d.dl, cells.n_cells, d.C, e, prec, ST(stream));
  }
  EpiContentOut<float> e{bc, (const float*)fc, (const float*)fm, fs, cells.code, d.C, (float*)cu, d.D};
  return gemm_dispatch(cc_hat, Wc, M, d.D, d.dl, d.dl, cells.n_cells, d.C, e, prec, ST(stream));
}

VML_API int vml_boundary_unit(const float* qproj, int ld, int off_kbt, int off_betab, const float* fw, const float* fs,
                      const float* fb, const void* fm, const uint8_t* query_mask, const uint8_t* length_mask,
                      vml_cells_t cells, float* g_scratch, float* ab_scratch, float* bu, void* fbar, const float* fbar_bias,
                      float* prob_out, float* u_out, int B, vml_dims_t d, int prec, void* stream) {
  VML_PREC_OK(prec);
  return boundary_unit(qproj, ld, off_kbt, off_betab, fw, fs, fb, fm, query_mask, length_mask, cells, g_scratch, ab_scratch, bu,
                       fbar, fbar_bias, prob_out, u_out, B, d, prec, ST(stream));
}

VML_API int vml_boundary_pair_fused(vml_dims_t d, int prec) { return boundary_pair_fused(d, prec) ? 1 : 0; }

VML_API int vml_boundary_unit_pair(const float* qproj, int ld, int off_kbt, int off_betab, const float* fw, const float* fs,
                                   const float* fb, const void* fm, const uint8_t* query_mask, const uint8_t* length_mask,
                                   vml_cells_t cells, float* g_scratch, float* ab_scratch, float* bu, void* fbar,
                                   const float* fbar_bias, void* operand, int B, vml_dims_t d, int prec, void* stream) {
  VML_PREC_OK(prec);
  VML_CHECK_ARG(operand != nullptr && (reinterpret_cast<uintptr_t>(operand) & 15) == 0);
  return boundary_unit(qproj, ld, off_kbt, off_betab, fw, fs, fb, fm, query_mask, length_mask, cells, g_scratch, ab_scratch, bu,
                       fbar, fbar_bias, nullptr, nullptr, B, d, prec, ST(stream), operand, 2 * d.D);
}

VML_API int vml_moment_operand(const void* cu, const float* bu, vml_cells_t cells, void* operand, vml_dims_t d, int prec, void* stream) {
  VML_PREC_OK(prec);
  return moment_operand(cu, bu, cells, operand, d, prec, ST(stream));
}

VML_API int vml_moment_pair(const float* bu, vml_cells_t cells, void* operand, vml_dims_t d, int prec, void* stream) {
  VML_PREC_OK(prec);
  return moment_pair(bu, cells, operand, d, prec, ST(stream));
}

VML_API int vml_moment_out(const void* operand, const void* Wcat, const float* bias_sum, const void* fm, vml_cells_t cells,
                   void* mu, vml_dims_t d, int prec, void* stream) {
  VML_PREC_OK(prec);
  // (K = 2D makes this stage tensor-bound: the register epilogue with 128 x 256 tiles measured faster than the
  //  shared-memory residual epilogue of gemm_res.cu, which is used for the byte-bound a6 tail)
  if (prec == VML_BF16 && d.D % 128 == 0 && getenv("VML_MOMENT_RES") != nullptr)   // A/B knob: TMA-in / TMA-out residual epilogue, 128 x 128 tiles
    return gemm_res(operand, Wcat, bias_sum, fm, nullptr, mu, nullptr, cells.capacity, d.D, 2 * d.D, 2 * d.D, 0, cells.n_cells, 1,
                    ST(stream));
  if (prec == VML_BF16) {
    // (A/B knob) L2 prefetch of the next tile's operand rows: measured on B200 it SLOWS the stage (20.9 -> 23.5 us per step at
    // 640-query passes): the loop is not short of latency cover, the prefetches only add L2 request traffic
    static const int pf = getenv("VML_GEMM_PREFETCH") != nullptr;
    if (getenv("VML_EPI_DIRECT") == nullptr) {        // coalescing epilogue through a shared-memory transpose (epilogues.cuh)
      EpiMomentOutT e{bias_sum, (const bf16*)fm, (bf16*)mu, d.D, pf, cells.capacity};
      return launch_gemm_umma(operand, Wcat, cells.capacity, d.D, 2 * d.D, 2 * d.D, 2 * d.D, cells.n_cells, 1, e, ST(stream));
    }
    EpiMomentOutPre e{bias_sum, (const bf16*)fm, (bf16*)mu, d.D, pf};
    return launch_gemm_umma(operand, Wcat, cells.capacity, d.D, 2 * d.D, 2 * d.D, 2 * d.D, cells.n_cells, 1, e, ST(stream));
  }
  EpiMomentOut<float> e{bias_sum, (const float*)fm, (float*)mu, d.D};
  return gemm_dispatch(operand, Wcat, cells.capacity, d.D, 2 * d.D, 2 * d.D, cells.n_cells, 1, e, prec, ST(stream));
}

// Opt-in (VML_MOMENT_GEN=1).  Validated bit-identical on B200 (GPU test) and measured 2.4x SLOWER (20.9 -> 50.9 us per step on
// the 640-query Charades pass): with 222 KB of the SM's 256 KB in use as shared memory the L1 is ~30 KB, so the generator warps'
// boundary-row loads go to L2 (~700 cycles under load) and the 80-register budget of the 704-thread CTA lets a lane keep only
// one of its eight row chunks in flight: ~2 us per generated k-block against a 0.7 us stage cadence.  It would need the
// tile's boundary rows staged in shared memory (64 KB that the pipeline + epilogue scratch do not leave).
static bool moment_gen_ok(vml_dims_t d, int prec) {
  const char* e = getenv("VML_MOMENT_GEN");
  return prec == VML_BF16 && d.D % 64 == 0 && d.L <= 255 && e != nullptr && atoi(e) != 0;
}
VML_API int vml_moment_gen_supported(vml_dims_t d, int prec) { return moment_gen_ok(d, prec) ? 1 : 0; }

VML_API int vml_moment_out_gen(const void* operand, const void* Wcat, const float* bias_sum, const void* fm, vml_cells_t cells,
                               const float* bu, void* mu, int B, vml_dims_t d, int prec, void* stream) {
  VML_PREC_OK(prec);
  VML_CHECK_ARG(moment_gen_ok(d, prec) && bu != nullptr && (int64_t)B * d.L * d.D < (int64_t)1 << 31 &&
                (reinterpret_cast<uintptr_t>(bu) & 15) == 0);
  EpiMomentOutGen e{{bias_sum, (const bf16*)fm, (bf16*)mu, d.D, 0, cells.capacity}, bu, cells.code, d.L, d.D};
  return launch_gemm_umma(operand, Wcat, cells.capacity, d.D, 2 * d.D, 2 * d.D, 2 * d.D, cells.n_cells, 1, e, ST(stream));
}

VML_API int vml_localize(const void* fm, const float* fb, const float* w4, const float* b4, vml_cells_t cells,
                 const uint8_t* length_mask, float* pm, float* ps, float* pe, float* pa, int B, vml_dims_t d, int prec,
                 void* stream) {
  VML_PREC_OK(prec);
  return localize(fm, fb, w4, b4, cells, length_mask, pm, ps, pe, pa, B, d, prec, ST(stream));
}

VML_API int vml_scaled_iou_bce(const float* pm, const uint8_t* ym, const float* sm, const uint8_t* moment_mask, const float* ps,
                       const uint8_t* ys, const float* ss, const float* pe, const uint8_t* ye, const float* se,
                       const float* pa, const uint8_t* ya, const uint8_t* length_mask, int B, int L, float* loss,
                       float* parts, float* scratch, float* g_pm, float* g_ps, float* g_pe, float* g_pa, void* stream) {
  return scaled_iou_bce(pm, ym, sm, moment_mask, ps, ys, ss, pe, ye, se, pa, ya, length_mask, B, L, loss, parts, scratch,
                        g_pm, g_ps, g_pe, g_pa, ST(stream));
}

VML_API int vml_score_topk_recall(const float* pm, const float* ps, const float* pe, const uint8_t* moment_mask, const float* sm,
                          int B, int L, int k, int nms_num, int nms_den, int32_t* top_idx, float* top_score,
                          float* top_iou, int64_t* counts, int64_t* step_counts, int step_group, void* stream) {
  return score_topk_recall(pm, ps, pe, moment_mask, sm, B, L, k, nms_num, nms_den, top_idx, top_score, top_iou, counts,
                           step_counts, step_group, nullptr, 0, nullptr, 0, ST(stream));
}

VML_API int vml_score_topk_recall_nm(const float* pm, const float* ps, const float* pe, const uint8_t* moment_mask, const float* sm,
                          int B, int L, int k, int nms_num, int nms_den, int32_t* top_idx, float* top_score,
                          float* top_iou, int64_t* counts, int64_t* step_counts, int step_group, const int32_t* ns, int n_n,
                          const float* ms, int n_m, void* stream) {
  VML_CHECK_ARG(ns != nullptr && ms != nullptr);
  return score_topk_recall(pm, ps, pe, moment_mask, sm, B, L, k, nms_num, nms_den, top_idx, top_score, top_iou, counts,
                           step_counts, step_group, ns, n_n, ms, n_m, ST(stream));
}

// ---- backward / training path (fp32) ---------------------------------------------------------------------
VML_API int vml_colsum(const float* X, int64_t row_stride, int64_t batch_stride, float* out, int64_t out_batch_stride, int M, int N,
                       int batch, const int32_t* m_dev, int m_scale, float alpha, void* stream) {
  return colsum(X, row_stride, batch_stride, out, out_batch_stride, M, N, batch, m_dev, m_scale, alpha, ST(stream));
}
VML_API int vml_localize_bwd(const float* fm, const float* fb, const float* w4, const float* pm, const float* ps, const float* pe,
                             const float* pa, const float* g_pm, const float* g_ps, const float* g_pe, const float* g_pa,
                             const uint8_t* length_mask, vml_cells_t cells, float* d_fm, float* d_fb, float* dw4, float* db4, int B,
                             vml_dims_t d, void* stream) {
  return localize_bwd(fm, fb, w4, pm, ps, pe, pa, g_pm, g_ps, g_pe, g_pa, length_mask, cells, d_fm, d_fb, dw4, db4, B, d, ST(stream));
}
VML_API int vml_pair_bwd(const float* d_operand, int ld_operand, const float* bu, vml_cells_t cells, float* d_bu, int B, vml_dims_t d,
                         void* stream) {
  return pair_bwd(d_operand, ld_operand, bu, cells, d_bu, B, d, ST(stream));
}
VML_API int vml_cu_tail_bwd(const float* d_cu_next, const float* d_operand, int ld_operand, vml_cells_t cells, float* dY, float* d_gbar,
                            vml_dims_t d, void* stream) {
  return cu_tail_bwd(d_cu_next, d_operand, ld_operand, cells, dY, d_gbar, d, ST(stream));
}
VML_API int vml_content_attn_bwd(const float* c_hat, const float* d_cc, const float* qproj, int ld, int off_what, int off_ktil,
                                 int off_beta, const float* s_hat, int s_ld, const uint8_t* query_mask, vml_cells_t cells,
                                 float* d_chat, float* dq, float* d_shat, int B, vml_dims_t d, void* stream) {
  return content_attn_bwd(c_hat, d_cc, qproj, ld, off_what, off_ktil, off_beta, s_hat, s_ld, query_mask, cells, d_chat, dq, d_shat, B,
                          d, ST(stream));
}
VML_API int vml_gbar_bwd(const float* fm, const float* fs, const float* ab, const float* d_bu, const float* d_gbar_cu, const float* d_mu,
                         vml_cells_t cells, float* d_ab, float* d_fm, float* d_fs, int B, vml_dims_t d, void* stream) {
  return gbar_bwd(fm, fs, ab, d_bu, d_gbar_cu, d_mu, cells, d_ab, d_fm, d_fs, B, d, ST(stream));
}
VML_API int vml_softmax_bwd(const float* P, const float* dP, const uint8_t* colmask, float* dS, int batch, int R, int W, float scale,
                            void* stream) {
  return softmax_bwd(P, dP, colmask, dS, batch, R, W, scale, ST(stream));
}
VML_API int vml_gate_bwd(const float* dG, const float* fb, const float* U, const uint8_t* length_mask, float* d_fb, float* d_Aq,
                         float* tmp, int B, vml_dims_t d, void* stream) {
  return gate_bwd(dG, fb, U, length_mask, d_fb, d_Aq, tmp, B, d, ST(stream));
}
VML_API int vml_mask_rows(const float* X, const uint8_t* mask, float* Y, int64_t rows, int D, int accumulate, void* stream) {
  return mask_rows(X, mask, Y, rows, D, accumulate, ST(stream));
}
VML_API int vml_span_pool_bwd(const float* d_fc, const float* d_fm, const float* d_fb, const float* fv, const float* fs,
                              vml_cells_t cells, float* d_fv, float* d_fs, int B, vml_dims_t d, void* stream) {
  return span_pool_bwd(d_fc, d_fm, d_fb, fv, fs, cells, d_fv, d_fs, B, d, ST(stream));
}
VML_API int vml_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                          int step, float grad_scale, void* stream) {
  return adam_step(p, g, m, v, n, lr, beta1, beta2, eps, step, grad_scale, ST(stream));
}
VML_API int vml_lstm_train_fwd(const float* gin, const float* whh_t, const int32_t* qlen, float* y, float* fs, float* acts, int B,
                               int Nq, int H, void* stream) {
  // the 8-CTA-cluster recurrence of the fp32 inference path (W_hh slices in shared memory, h exchanged through DSMEM), saving
  // the gate activations on the way; the one-CTA-per-sample kernel only for shapes the cluster kernel does not take
  const size_t UH = (size_t)H / 8;
  const size_t smem = sizeof(float) * ((size_t)H * 4 * UH + 2 * (size_t)H * 8 + 8 * 8 * 4 * UH + (size_t)Nq * 8 * UH);
  if (H % 32 == 0 && H <= 1024 && smem <= 227 * 1024 && getenv("VML_LSTM_TRAIN_NAIVE") == nullptr)
    return lstm_layer(gin, whh_t, qlen, y, nullptr, fs, nullptr, B, Nq, H, ST(stream), acts);
  return lstm_train_fwd(gin, whh_t, qlen, y, fs, acts, B, Nq, H, ST(stream));
}
VML_API int vml_lstm_train_bwd(const float* dy, const float* dfs, const float* whh, const float* acts, const int32_t* qlen, float* dgin,
                               float* dgin_rec, int B, int Nq, int H, void* stream) {
  if (getenv("VML_LSTM_TRAIN_NAIVE") == nullptr) {
    const int rc = lstm_bwd_cluster(dy, dfs, whh, acts, qlen, dgin, dgin_rec, B, Nq, H, ST(stream));   // 1: shape not taken
    if (rc <= 0) return rc;
  }
  return lstm_train_bwd(dy, dfs, whh, acts, qlen, dgin, dgin_rec, B, Nq, H, ST(stream));
}

}  // extern "C"
