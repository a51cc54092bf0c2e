// mmgclip_b200 -- extern "C" boundary (see include/mmgclip_b200.h).  Argument validation, error reporting and the
// block loops of the fused InfoNCE; all arithmetic lives in simt_kernels.cu / gemm_tc.cuh.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "../../include/mmgclip_b200.h"
#include "kernels.h"

namespace mmg {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  return set_error(MMG_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

// There is no CPU path: anything that is not device (or managed) memory is rejected.
static int require_device(const void* p, const char* name) {
  if (p == nullptr) return set_error(MMG_ERR_BAD_ARG, "%s is NULL", name);
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(MMG_ERR_NO_DEVICE, "%s: cannot query pointer (%s) -- is a CUDA device present?", name,
                     cudaGetErrorString(e));
  }
  if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)
    return set_error(MMG_ERR_BAD_ARG, "%s must be device memory (mmgclip_b200 has no CPU fallback)", name);
  return 0;
}

static inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

static const int kDefaultBlockBf16 = 8192;  // g block 8192 x 8192 bf16 = 128 MiB (measured best; 4096^2 = 32 MiB stays in L2)
static const int kDefaultBlockFp32 = 2048;  // S block 2048 x 2048 fp32 = 16 MiB

}  // namespace mmg

using namespace mmg;

#define MMG_REQ(p)                                      \
  do {                                                  \
    int rc__ = require_device((p), #p);                 \
    if (rc__ != 0) return rc__;                         \
  } while (0)
#define MMG_TRY(expr)              \
  do {                             \
    int rc__ = (expr);             \
    if (rc__ != 0) return rc__;    \
  } while (0)

extern "C" {

int mmg_version(void) { return 100; }

long long mmg_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* mmg_last_error_string(void) { return g_err; }

int mmg_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(MMG_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(MMG_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return 0;
}

int mmg_gemm(int prec, const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, float* C,
             long long ldc, int M, int N, int K, float alpha, const float* alpha_dev, const float* bias, int relu, int mode,
             int k_splits, mmg_stream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_gemm: empty problem %dx%dx%d", M, N, K);
  if (mode < 0 || mode > 2) return set_error(MMG_ERR_BAD_ARG, "mmg_gemm: bad store mode %d", mode);
  MMG_REQ(A);
  MMG_REQ(B);
  MMG_REQ(C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (prec == MMG_PREC_BF16) {
    TcOperand a{A, lda, a_mn}, b{B, ldb, b_mn};
    return tc_gemm_store(a, b, C, ldc, M, N, K, alpha, alpha_dev, bias, relu, mode, k_splits, st);
  }
  if (prec == MMG_PREC_FP32)
    return simt_gemm(static_cast<const float*>(A), lda, a_mn, static_cast<const float*>(B), ldb, b_mn, C, ldc, M, N, K,
                     alpha, alpha_dev, bias, relu, mode, k_splits, st);
  return set_error(MMG_ERR_BAD_ARG, "mmg_gemm: unknown precision %d", prec);
}

int mmg_gemm_split(const void* A_hi, const void* A_lo, long long lda, int a_mn, const void* B_hi, const void* B_lo,
                   long long ldb, int b_mn, float* C, long long ldc, int M, int N, int K, float alpha, const float* bias,
                   int relu, int mode, int k_splits, mmg_stream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_gemm_split: empty problem %dx%dx%d", M, N, K);
  if (mode < 0 || mode > 2) return set_error(MMG_ERR_BAD_ARG, "mmg_gemm_split: bad store mode %d", mode);
  MMG_REQ(A_hi);
  MMG_REQ(B_hi);
  MMG_REQ(C);
  TcOperand a0{A_hi, lda, a_mn}, a1{A_lo, lda, a_mn}, b0{B_hi, ldb, b_mn}, b1{B_lo, ldb, b_mn};
  // segments: hi.hi, then hi.lo (if B has a low part), then lo.hi (if A has one)
  int nseg = 1, seg_a = 0, seg_b = 0;
  if (B_lo != nullptr) { seg_b |= 1 << nseg; ++nseg; }
  if (A_lo != nullptr) { seg_a |= 1 << nseg; ++nseg; }
  return tc_gemm_store_seg(a0, A_lo ? &a1 : nullptr, b0, B_lo ? &b1 : nullptr, nseg, seg_a, seg_b, C, ldc, M, N, K, alpha,
                           nullptr, bias, relu, mode, k_splits, static_cast<cudaStream_t>(stream));
}

int mmg_cast_f32_to_bf16(const float* x, void* y_bf16, long long n, mmg_stream_t stream) {
  if (n < 0) return set_error(MMG_ERR_BAD_ARG, "mmg_cast: negative length");
  if (n == 0) return 0;
  MMG_REQ(x);
  MMG_REQ(y_bf16);
  return simt_cast_bf16(x, y_bf16, n, static_cast<cudaStream_t>(stream));
}

int mmg_cast_f32_to_f16(const float* x, void* y_f16, long long n, mmg_stream_t stream) {
  if (n < 0) return set_error(MMG_ERR_BAD_ARG, "mmg_cast_f32_to_f16: negative length");
  if (n == 0) return 0;
  MMG_REQ(x);
  MMG_REQ(y_f16);
  return simt_cast_f16(x, y_f16, n, static_cast<cudaStream_t>(stream));
}

int mmg_cast_f32_to_bf16_split(const float* x, void* hi_bf16, void* lo_bf16, long long n, mmg_stream_t stream) {
  if (n < 0) return set_error(MMG_ERR_BAD_ARG, "mmg_cast_split: negative length");
  if (n == 0) return 0;
  MMG_REQ(x);
  MMG_REQ(hi_bf16);
  MMG_REQ(lo_bf16);
  return simt_cast_split(x, hi_bf16, lo_bf16, n, static_cast<cudaStream_t>(stream));
}

int mmg_push_rows(const void* src, long long bytes, void* const* dst_ptrs, int n_dst, long long dst_offset_bytes,
                  mmg_stream_t stream) {
  if (bytes < 0 || n_dst < 1 || n_dst > 8 || dst_offset_bytes < 0 || dst_ptrs == nullptr)
    return set_error(MMG_ERR_BAD_ARG, "mmg_push_rows: bad arguments (bytes %lld, %d destinations)", bytes, n_dst);
  if (bytes == 0) return 0;
  MMG_REQ(src);
  if ((bytes & 15) != 0 || (dst_offset_bytes & 15) != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0)
    return set_error(MMG_ERR_BAD_ALIGN, "mmg_push_rows: source, offset and length must be multiples of 16 bytes");
  for (int i = 0; i < n_dst; ++i)
    if (dst_ptrs[i] == nullptr || (reinterpret_cast<uintptr_t>(dst_ptrs[i]) & 15) != 0)
      return set_error(MMG_ERR_BAD_ALIGN, "mmg_push_rows: destination %d is NULL or not 16-byte aligned", i);
  return simt_push_rows(src, bytes, dst_ptrs, n_dst, dst_offset_bytes, static_cast<cudaStream_t>(stream));
}

int mmg_l2norm_fwd(const float* u, int B, int D, float* y, float* inv_norm, void* y_16, int y16_is_f16,
                   mmg_stream_t stream) {
  if (B < 0 || D <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_l2norm_fwd: bad shape %dx%d", B, D);
  if (B == 0) return 0;
  MMG_REQ(u);
  MMG_REQ(y);
  return simt_l2norm_fwd(u, B, D, y, inv_norm, y_16, y16_is_f16 ? 1 : 0, static_cast<cudaStream_t>(stream));
}

int mmg_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, int B, int D, float* du, void* du_bf16,
                   void* du_bf16_lo, float* zero_fill, long long zero_floats, mmg_stream_t stream) {
  if (B < 0 || D <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_l2norm_bwd: bad shape %dx%d", B, D);
  if (zero_fill != nullptr && ((reinterpret_cast<uintptr_t>(zero_fill) & 15) != 0 || zero_floats < 0 || (zero_floats & 3) != 0))
    return set_error(MMG_ERR_BAD_ALIGN, "mmg_l2norm_bwd: zero_fill must be 16-byte aligned, a multiple of 4 floats");
  if (B == 0) {
    if (zero_fill != nullptr && zero_floats > 0)
      return check_cuda(cudaMemsetAsync(zero_fill, 0, (size_t)zero_floats * 4, static_cast<cudaStream_t>(stream)),
                        "cudaMemsetAsync(zero_fill)");
    return 0;
  }
  MMG_REQ(dy);
  MMG_REQ(y);
  MMG_REQ(inv_norm);
  if (du == nullptr && du_bf16 == nullptr) return set_error(MMG_ERR_BAD_ARG, "mmg_l2norm_bwd: no output");
  if (du_bf16_lo != nullptr && du_bf16 == nullptr)
    return set_error(MMG_ERR_BAD_ARG, "mmg_l2norm_bwd: du_bf16_lo needs du_bf16");
  return simt_l2norm_bwd(dy, y, inv_norm, B, D, du, du_bf16, du_bf16_lo, zero_fill, zero_floats,
                         static_cast<cudaStream_t>(stream));
}

int mmg_dropout_apply(float* y, const uint8_t* mask, float keep_scale, long long n, mmg_stream_t stream) {
  if (n <= 0) return 0;
  MMG_REQ(y);
  MMG_REQ(mask);
  return simt_dropout_apply(y, mask, keep_scale, n, static_cast<cudaStream_t>(stream));
}

int mmg_dropout_draw_apply(float* y, uint8_t* mask_out, float p, long long n, unsigned long long* state,
                           mmg_stream_t stream) {
  if (n < 0 || !(p >= 0.f && p <= 1.f)) return set_error(MMG_ERR_BAD_ARG, "mmg_dropout_draw_apply: bad n or p");
  if (n == 0) return 0;
  MMG_REQ(y);
  MMG_REQ(mask_out);
  MMG_REQ(state);
  if ((reinterpret_cast<uintptr_t>(state) & 7) != 0)
    return set_error(MMG_ERR_BAD_ALIGN, "mmg_dropout_draw_apply: state must be 8-byte aligned");
  return simt_dropout_draw_apply(y, mask_out, p, n, state, static_cast<cudaStream_t>(stream));
}

int mmg_relu_dropout_bwd(const float* dy, const float* y, const uint8_t* mask, float keep_scale, float* dz,
                         long long n, mmg_stream_t stream) {
  if (n <= 0) return 0;
  MMG_REQ(dy);
  MMG_REQ(dz);
  return simt_relu_dropout_bwd(dy, y, mask, keep_scale, dz, n, static_cast<cudaStream_t>(stream));
}

int mmg_colsum(const float* x, int rows, int cols, float* out, mmg_stream_t stream) {
  if (rows < 0 || cols <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_colsum: bad shape");
  MMG_REQ(x);
  MMG_REQ(out);
  return simt_colsum(x, rows, cols, out, static_cast<cudaStream_t>(stream));
}

int mmg_add(const float* x, const float* y, float* out, long long n, mmg_stream_t stream) {
  if (n <= 0) return 0;
  MMG_REQ(x);
  MMG_REQ(y);
  MMG_REQ(out);
  return simt_add(x, y, out, n, static_cast<cudaStream_t>(stream));
}

int mmg_gelu_fwd(const float* x, float* y, long long n, mmg_stream_t stream) {
  if (n <= 0) return 0;
  MMG_REQ(x);
  MMG_REQ(y);
  return simt_gelu_fwd(x, y, n, static_cast<cudaStream_t>(stream));
}

int mmg_gelu_bwd(const float* dy, const float* x, float* dx, long long n, mmg_stream_t stream) {
  if (n <= 0) return 0;
  MMG_REQ(dy);
  MMG_REQ(x);
  MMG_REQ(dx);
  return simt_gelu_bwd(dy, x, dx, n, static_cast<cudaStream_t>(stream));
}

int mmg_layernorm_fwd(const float* x, const float* gamma, const float* beta, int rows, int cols, float eps, float* y,
                      float* mean, float* rstd, mmg_stream_t stream) {
  if (rows <= 0 || cols <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_layernorm_fwd: bad shape");
  MMG_REQ(x);
  MMG_REQ(gamma);
  MMG_REQ(beta);
  MMG_REQ(y);
  MMG_REQ(mean);
  MMG_REQ(rstd);
  return simt_layernorm_fwd(x, gamma, beta, rows, cols, eps, y, mean, rstd, static_cast<cudaStream_t>(stream));
}

int mmg_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                      int rows, int cols, float* dx, float* dgamma, float* dbeta, mmg_stream_t stream) {
  if (rows <= 0 || cols <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_layernorm_bwd: bad shape");
  MMG_REQ(dy);
  MMG_REQ(x);
  MMG_REQ(dx);
  MMG_REQ(dgamma);
  MMG_REQ(dbeta);
  return simt_layernorm_bwd(dy, x, gamma, mean, rstd, rows, cols, dx, dgamma, dbeta,
                            static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------------------
// fused InfoNCE
// ---------------------------------------------------------------------------------------------------------
size_t mmg_infonce_workspace_bytes(int prec, int rows, int cols, int D) {
  (void)D;
  if (rows <= 0 || cols <= 0) return 0;
  if (prec == MMG_PREC_BF16 || prec == MMG_PREC_F16) {
    const long long rb = rows < kDefaultBlockBf16 ? rows : kDefaultBlockBf16;
    const long long cb = round_up(cols < kDefaultBlockBf16 ? cols : kDefaultBlockBf16, 64);
    const size_t loop = (size_t)(rb * cb * 2 + 256);
    const size_t fused = tc_infonce_bwd_fused_workspace(rows, cols, D);
    return loop > fused ? loop : fused;
  }
  const long long rb = rows < kDefaultBlockFp32 ? rows : kDefaultBlockFp32;
  const long long cb = round_up(cols < kDefaultBlockFp32 ? cols : kDefaultBlockFp32, 4);
  return (size_t)(rb * cb * 4 + 256);
}

static int check_infonce_args(const char* fn, int prec, const void* a_hat, const void* b_hat, int rows, int cols, int D,
                              int diag_offset, const float* scale) {
  if (prec != MMG_PREC_BF16 && prec != MMG_PREC_FP32 && prec != MMG_PREC_F16)
    return set_error(MMG_ERR_BAD_ARG, "%s: unknown precision", fn);
  if (rows <= 0 || cols <= 0 || D <= 0) return set_error(MMG_ERR_BAD_ARG, "%s: bad shape %dx%dx%d", fn, rows, cols, D);
  if (diag_offset < 0 || diag_offset + rows > cols)
    return set_error(MMG_ERR_BAD_ARG, "%s: rows [%d, %d) have no matching columns in [0, %d)", fn, diag_offset,
                     diag_offset + rows, cols);
  if (prec != MMG_PREC_FP32 && (D % 8) != 0)
    return set_error(MMG_ERR_UNSUPPORTED_SHAPE, "%s: tensor-core path needs D %% 8 == 0 (TMA pitch), got %d", fn, D);
  int rc;
  if ((rc = require_device(a_hat, "a_hat")) != 0) return rc;
  if ((rc = require_device(b_hat, "b_hat")) != 0) return rc;
  if ((rc = require_device(scale, "scale")) != 0) return rc;
  return 0;
}

int mmg_infonce_fwd(int prec, const void* a_hat, const void* b_hat, int rows, int cols, int D, int diag_offset,
                    const float* scale, float* rowsum, float* colsum, float* diag, void* workspace,
                    size_t workspace_bytes, mmg_stream_t stream) {
  MMG_TRY(check_infonce_args("mmg_infonce_fwd", prec, a_hat, b_hat, rows, cols, D, diag_offset, scale));
  MMG_REQ(rowsum);
  MMG_REQ(colsum);
  MMG_REQ(diag);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (prec != MMG_PREC_FP32) {
    // One persistent launch over all logit tiles; the tile lives only in TMEM.
    return tc_infonce_fwd(a_hat, b_hat, rows, cols, D, diag_offset, scale, rowsum, colsum, diag, nullptr, 0, st,
                          prec == MMG_PREC_F16);
  }
  if (workspace_bytes < mmg_infonce_workspace_bytes(prec, rows, cols, D))
    return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_fwd: workspace too small");
  MMG_REQ(workspace);
  const float* a = static_cast<const float*>(a_hat);
  const float* b = static_cast<const float*>(b_hat);
  float* S = static_cast<float*>(workspace);
  const int Rb = rows < kDefaultBlockFp32 ? rows : kDefaultBlockFp32;
  const int Cb = cols < kDefaultBlockFp32 ? cols : kDefaultBlockFp32;
  const long long lds = round_up(Cb, 4);
  for (int r0 = 0; r0 < rows; r0 += Rb) {
    const int rb = rows - r0 < Rb ? rows - r0 : Rb;
    for (int c0 = 0; c0 < cols; c0 += Cb) {
      const int cb = cols - c0 < Cb ? cols - c0 : Cb;
      MMG_TRY(simt_gemm(a + (long long)r0 * D, D, 0, b + (long long)c0 * D, D, 0, S, lds, rb, cb, D, 1.f, nullptr, nullptr, 0,
                        0, 1, st));
      MMG_TRY(simt_lse_block(S, lds, rb, cb, r0, c0, diag_offset, scale, rowsum, colsum, diag, st));
    }
  }
  return 0;
}

int mmg_infonce_stored_supported(int rows, int cols, int D, int n_owners, int n_parts) {
  if (rows <= 0 || cols <= 0 || D <= 0) return 0;
  return tc_infonce_stored_supported(rows, cols, D, n_owners, n_parts);
}

int mmg_infonce_fwd_store(int prec, const void* a_hat, const void* b_hat, int rows, int cols, int D, int diag_offset,
                          const float* scale, float* rowsum, float* colsum, float* diag, void* e_out, long long lde,
                          mmg_stream_t stream) {
  if (prec != MMG_PREC_BF16 && prec != MMG_PREC_F16)
    return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_fwd_store: tensor-core path only (MMG_PREC_BF16 / MMG_PREC_F16)");
  MMG_TRY(check_infonce_args("mmg_infonce_fwd_store", prec, a_hat, b_hat, rows, cols, D, diag_offset, scale));
  MMG_REQ(rowsum);
  MMG_REQ(colsum);
  MMG_REQ(diag);
  MMG_REQ(e_out);
  if (lde < cols || (lde & 7) != 0 || (reinterpret_cast<uintptr_t>(e_out) & 15) != 0)
    return set_error(MMG_ERR_BAD_ALIGN, "mmg_infonce_fwd_store: E needs a pitch >= cols that is a multiple of 8 and a "
                                        "16-byte aligned base");
  return tc_infonce_fwd(a_hat, b_hat, rows, cols, D, diag_offset, scale, rowsum, colsum, diag, e_out, lde,
                        static_cast<cudaStream_t>(stream), prec == MMG_PREC_F16);
}

int mmg_infonce_bwd_stored(int prec, const void* a_hat, const void* b_hat, const void* e_stored, long long lde, int rows,
                           int cols, int D, int diag_offset, const float* scale, const float* rinv, const float* cinv,
                           const float* scal, float* dA, float* const* dB_owners, int n_owners, int n_parts, int part,
                           void* workspace, size_t workspace_bytes, mmg_stream_t stream) {
  if (prec != MMG_PREC_BF16)
    return set_error(MMG_ERR_UNSUPPORTED_SHAPE, "mmg_infonce_bwd_stored: the stored-E transform writes bf16 coefficients "
                                                "(MMG_PREC_BF16 operands only)");
  MMG_TRY(check_infonce_args("mmg_infonce_bwd_stored", prec, a_hat, b_hat, rows, cols, D, diag_offset, scale));
  MMG_REQ(e_stored);
  MMG_REQ(rinv);
  MMG_REQ(cinv);
  MMG_REQ(scal);
  MMG_REQ(dA);
  MMG_REQ(workspace);
  if (dB_owners == nullptr || n_owners < 1 || n_owners > 8)
    return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_bwd_stored: 1..8 owners (got %d)", n_owners);
  for (int i = 0; i < n_owners; ++i)
    if (dB_owners[i] == nullptr) return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_bwd_stored: owner %d is NULL", i);
  if (n_parts < 1 || part < 0 || part >= n_parts)
    return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_bwd_stored: bad column part %d of %d", part, n_parts);
  int used = 0;
  MMG_TRY(tc_infonce_bwd_fused(a_hat, b_hat, rows, cols, D, diag_offset, scale, rinv, cinv, scal, dA, dB_owners,
                               n_owners, n_parts, part, nullptr, workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream), &used, e_stored, lde, 0));
  if (!used)
    return set_error(MMG_ERR_UNSUPPORTED_SHAPE,
                     "mmg_infonce_bwd_stored: shape not covered (see mmg_infonce_stored_supported), misaligned E or "
                     "outputs, or workspace smaller than mmg_infonce_workspace_bytes()");
  return 0;
}

int mmg_infonce_loss(const float* rowsum, const float* colsum, const float* diag, int n, const float* scale,
                     float inv_two_b, float* loss_out, mmg_stream_t stream) {
  if (n <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_loss: n must be positive");
  MMG_REQ(rowsum);
  MMG_REQ(colsum);
  MMG_REQ(diag);
  MMG_REQ(scale);
  MMG_REQ(loss_out);
  return simt_infonce_loss(rowsum, colsum, diag, n, scale, inv_two_b, loss_out, static_cast<cudaStream_t>(stream));
}

int mmg_infonce_row_part(const float* rowsum, const float* diag, int rows, float* part_out, mmg_stream_t stream) {
  if (rows <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_row_part: rows must be positive");
  MMG_REQ(rowsum);
  MMG_REQ(diag);
  MMG_REQ(part_out);
  return simt_infonce_row_part(rowsum, diag, rows, part_out, static_cast<cudaStream_t>(stream));
}

int mmg_infonce_loss_cols(const float* colsum, int cols, const float* scale, const float* row_part, float inv_two_b,
                          float* loss_out, mmg_stream_t stream) {
  if (cols <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_loss_cols: cols must be positive");
  MMG_REQ(colsum);
  MMG_REQ(scale);
  MMG_REQ(row_part);
  MMG_REQ(loss_out);
  return simt_infonce_loss_cols(colsum, cols, scale, row_part, inv_two_b, loss_out, static_cast<cudaStream_t>(stream));
}

int mmg_infonce_bwd_prep_diag(int prec, const float* rowsum, int rows, const float* colsum, int cols, int diag_offset,
                              const float* scale, const float* grad_loss, float inv_two_b, float* rinv, float* cinv,
                              float* scal, const float* a32, const float* b32, int D, const float* diag, float* dA,
                              float* dB_matching, float* dlogscale_acc, mmg_stream_t stream) {
  if (prec != MMG_PREC_BF16 && prec != MMG_PREC_F16)
    return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_bwd_prep_diag: tensor-core path only (MMG_PREC_BF16 / MMG_PREC_F16)");
  if (rows <= 0 || cols <= 0 || D <= 0 || diag_offset < 0 || diag_offset + rows > cols)
    return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_bwd_prep_diag: bad shape");
  MMG_REQ(rowsum);
  MMG_REQ(colsum);
  MMG_REQ(scale);
  MMG_REQ(grad_loss);
  MMG_REQ(rinv);
  MMG_REQ(cinv);
  MMG_REQ(scal);
  MMG_REQ(a32);
  MMG_REQ(b32);
  MMG_REQ(diag);
  MMG_REQ(dA);
  MMG_REQ(dB_matching);
  return simt_infonce_bwd_prep_diag(rowsum, rows, colsum, cols, diag_offset, scale, grad_loss, inv_two_b,
                                    prec == MMG_PREC_F16, rinv, cinv, scal, a32, b32, D, diag, dA, dB_matching,
                                    dlogscale_acc, static_cast<cudaStream_t>(stream));
}

int mmg_infonce_bwd_prep(int prec, const float* rowsum, int rows, const float* colsum, int cols, const float* scale,
                         const float* grad_loss, float inv_two_b, int diag_in_fp32, float* rinv, float* cinv,
                         float* scal, mmg_stream_t stream) {
  if (prec != MMG_PREC_BF16 && prec != MMG_PREC_F16 && prec != MMG_PREC_FP32)
    return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_bwd_prep: unknown precision");
  if (rows <= 0 || cols <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_bwd_prep: bad shape");
  MMG_REQ(rowsum);
  MMG_REQ(colsum);
  MMG_REQ(scale);
  MMG_REQ(grad_loss);
  MMG_REQ(rinv);
  MMG_REQ(cinv);
  MMG_REQ(scal);
  return simt_infonce_bwd_prep(rowsum, rows, colsum, cols, scale, grad_loss, inv_two_b, diag_in_fp32,
                               prec == MMG_PREC_F16, rinv, cinv, scal, static_cast<cudaStream_t>(stream));
}

int mmg_infonce_bwd_diag(const float* a32, const float* b32, int rows, int D, const float* diag, const float* scale,
                         const float* rinv, const float* cinv_paired, const float* scal, float* dA, float* dB,
                         float* dlogscale_acc, int init, mmg_stream_t stream) {
  if (rows <= 0 || D <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_bwd_diag: bad shape");
  MMG_REQ(a32);
  MMG_REQ(b32);
  MMG_REQ(diag);
  MMG_REQ(scale);
  MMG_REQ(rinv);
  MMG_REQ(cinv_paired);
  MMG_REQ(scal);
  MMG_REQ(dA);
  MMG_REQ(dB);
  if (dlogscale_acc != nullptr) MMG_REQ(dlogscale_acc);  // NULL: d/d logit_scale not wanted
  return simt_infonce_bwd_diag(a32, b32, rows, D, diag, scale, rinv, cinv_paired, scal, dA, dB, dlogscale_acc,
                               init != 0, static_cast<cudaStream_t>(stream));
}

int mmg_infonce_bwd(int prec, const void* a_hat, const void* b_hat, int rows, int cols, int D, int diag_offset,
                    const float* scale, const float* rinv, const float* cinv, const float* scal, float* dA, float* dB,
                    float* dlogscale_acc, int block_rows, int block_cols, void* workspace, size_t workspace_bytes,
                    mmg_stream_t stream) {
  MMG_TRY(check_infonce_args("mmg_infonce_bwd", prec, a_hat, b_hat, rows, cols, D, diag_offset, scale));
  MMG_REQ(rinv);
  MMG_REQ(cinv);
  MMG_REQ(scal);
  MMG_REQ(dA);
  MMG_REQ(dB);
  if (dlogscale_acc != nullptr) MMG_REQ(dlogscale_acc);  // NULL: d/d logit_scale not wanted (skips sum g*cos)
  MMG_REQ(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool tc = prec != MMG_PREC_FP32;  // tensor-core path (bf16 or fp16 embedding operands)
  const int emb_f16 = prec == MMG_PREC_F16 ? 1 : 0;
  const int dflt = tc ? kDefaultBlockBf16 : kDefaultBlockFp32;
  const int esz = tc ? 2 : 4;
  const int pad = tc ? 64 : 4;
  int Rb = block_rows > 0 ? block_rows : dflt;
  int Cb = block_cols > 0 ? block_cols : dflt;
  if (Rb > rows) Rb = rows;
  if (Cb > cols) Cb = cols;
  long long ldg = round_up(Cb, pad);
  // shrink the block until it fits the caller's workspace
  while ((size_t)((long long)Rb * ldg * esz) > workspace_bytes && Rb > 128) Rb = (Rb + 1) / 2;
  if ((size_t)((long long)Rb * ldg * esz) > workspace_bytes)
    return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_bwd: workspace too small (%zu bytes)", workspace_bytes);

  int phases = 3;
#ifdef MMG_MEASURE
  // measurement builds only (tests/gpu_epi_probe.py): MMG_BWD_PHASES=1 runs only the coefficient launches, =2 only the
  // gradient-GEMM launches (on whatever the scratch holds -- wrong results); unset / 3 = the real thing
  if (const char* e = getenv("MMG_BWD_PHASES")) phases = atoi(e);
  if (phases < 1 || phases > 3) phases = 3;
#endif

  // One persistent launch for the whole backward when the shape allows it (bwd_fused.cuh); explicit block shapes and
  // the phase hook select the block loop below.
  if (tc && phases == 3 && block_rows <= 0 && block_cols <= 0) {
    int used = 0;
    float* owners[1] = {dB};
    MMG_TRY(tc_infonce_bwd_fused(a_hat, b_hat, rows, cols, D, diag_offset, scale, rinv, cinv, scal, dA, owners, 1, 1, 0,
                                 dlogscale_acc, workspace, workspace_bytes, st, &used, nullptr, 0, emb_f16));
    if (used) return 0;
  }

  for (int r0 = 0; r0 < rows; r0 += Rb) {
    const int rb = rows - r0 < Rb ? rows - r0 : Rb;
    for (int c0 = 0; c0 < cols; c0 += Cb) {
      const int cb = cols - c0 < Cb ? cols - c0 : Cb;
      const int doff = r0 + diag_offset - c0;
      if (tc) {
        const char* a = static_cast<const char*>(a_hat) + (long long)r0 * D * 2;
        const char* b = static_cast<const char*>(b_hat) + (long long)c0 * D * 2;
        // (1) recompute the cosine block on tensor cores; epilogue turns it into bf16 gradient coefficients g
        if (phases & 1)
          MMG_TRY(tc_infonce_grad_block(a, b, rb, cb, D, doff, scale, rinv + r0, cinv + c0, scal, workspace, ldg,
                                        dlogscale_acc, st, emb_f16));
        if (!(phases & 2)) continue;
        // (2) dA[r0:, :] += g . b_blk   and   dB[c0:, :] += g^T . a_blk   in one launch
        // scal[3] = factor that turns the stored coefficients back into true ones (1 for bf16, coef * 2^-14 for fp16)
        TcOperand A0{workspace, ldg, 0, emb_f16}, B0{b, D, 1, emb_f16};
        TcOperand A1{workspace, ldg, 1, emb_f16}, B1{a, D, 1, emb_f16};
        MMG_TRY(tc_gemm_dual_accumulate(A0, B0, dA + (long long)r0 * D, D, rb, D, cb, A1, B1, dB + (long long)c0 * D,
                                        D, cb, D, rb, st, scal + 3));
      } else {
        const float* a = static_cast<const float*>(a_hat) + (long long)r0 * D;
        const float* b = static_cast<const float*>(b_hat) + (long long)c0 * D;
        float* S = static_cast<float*>(workspace);
        MMG_TRY(simt_gemm(a, D, 0, b, D, 0, S, ldg, rb, cb, D, 1.f, nullptr, nullptr, 0, 0, 1, st));
        MMG_TRY(simt_grad_block(S, ldg, rb, cb, 0, 0, doff, scale, rinv + r0, cinv + c0, scal, dlogscale_acc, st));
        MMG_TRY(simt_gemm(S, ldg, 0, b, D, 1, dA + (long long)r0 * D, D, rb, D, cb, 1.f, nullptr, nullptr, 0, 1, 1, st));
        MMG_TRY(simt_gemm(S, ldg, 1, a, D, 1, dB + (long long)c0 * D, D, cb, D, rb, 1.f, nullptr, nullptr, 0, 1, 1, st));
      }
    }
  }
  return 0;
}

int mmg_infonce_bwd_owners(int prec, const void* a_hat, const void* b_hat, int rows, int cols, int D, int diag_offset,
                           const float* scale, const float* rinv, const float* cinv, const float* scal, float* dA,
                           float* const* dB_owners, int n_owners, int n_parts, int part, float* dlogscale_acc,
                           void* workspace, size_t workspace_bytes, mmg_stream_t stream) {
  if (prec != MMG_PREC_BF16 && prec != MMG_PREC_F16)
    return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_bwd_owners: tensor-core path only (MMG_PREC_BF16 / MMG_PREC_F16)");
  MMG_TRY(check_infonce_args("mmg_infonce_bwd_owners", prec, a_hat, b_hat, rows, cols, D, diag_offset, scale));
  MMG_REQ(rinv);
  MMG_REQ(cinv);
  MMG_REQ(scal);
  MMG_REQ(dA);
  MMG_REQ(workspace);
  if (dB_owners == nullptr || n_owners < 1 || n_owners > 8)
    return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_bwd_owners: 1..8 owners (got %d)", n_owners);
  for (int i = 0; i < n_owners; ++i)
    if (dB_owners[i] == nullptr) return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_bwd_owners: owner %d is NULL", i);
  int used = 0;
  if (n_parts < 1 || part < 0 || part >= n_parts)
    return set_error(MMG_ERR_BAD_ARG, "mmg_infonce_bwd_owners: bad column part %d of %d", part, n_parts);
  MMG_TRY(tc_infonce_bwd_fused(a_hat, b_hat, rows, cols, D, diag_offset, scale, rinv, cinv, scal, dA, dB_owners,
                               n_owners, n_parts, part, dlogscale_acc, workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream), &used, nullptr, 0, prec == MMG_PREC_F16));
  if (!used)
    return set_error(MMG_ERR_UNSUPPORTED_SHAPE,
                     "mmg_infonce_bwd_owners: needs rows, cols / owners and D to be multiples of 256, column parts made of "
                     "whole column blocks, 16-byte aligned outputs and a workspace of mmg_infonce_workspace_bytes() (use "
                     "mmg_infonce_bwd + a reduce-scatter)");
  return 0;
}

int mmg_debug_fused_trace_region(int rows, int cols, int D, size_t* offset, size_t* bytes, int* records_per_role,
                                 int* roles) {
  if (offset == nullptr || bytes == nullptr || records_per_role == nullptr || roles == nullptr) return 0;
  return tc_fused_trace_region(rows, cols, D, offset, bytes, records_per_role, roles);
}

int mmg_tune(const char* key, int value) { return tc_tune(key, value); }

int mmg_fused_bwd_schedule(int rows, int cols, int D, int n_owners, int n_parts, int part, int pairs, int pair, int* items,
                           int max_items, int* info) {
  return tc_fused_bwd_schedule(rows, cols, D, n_owners, n_parts, part, pairs, pair, items, max_items, info);
}

int mmg_eos_pool(const float* hidden, const long long* attention_mask, int n, int seq, int H, float* out,
                 long long* idx_out, mmg_stream_t stream) {
  if (n < 0 || seq <= 0 || H <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_eos_pool: bad shape %d x %d x %d", n, seq, H);
  if (n == 0) return 0;
  MMG_REQ(hidden);
  MMG_REQ(attention_mask);
  MMG_REQ(out);
  if (idx_out != nullptr) MMG_REQ(idx_out);
  return simt_eos_pool(hidden, attention_mask, n, seq, H, out, idx_out, static_cast<cudaStream_t>(stream));
}

int mmg_eos_pool_bwd(const float* dout, const long long* idx, int n, int seq, int H, float* dhidden,
                     mmg_stream_t stream) {
  if (n < 0 || seq <= 0 || H <= 0)
    return set_error(MMG_ERR_BAD_ARG, "mmg_eos_pool_bwd: bad shape %d x %d x %d", n, seq, H);
  if (n == 0) return 0;
  MMG_REQ(dout);
  MMG_REQ(idx);
  MMG_REQ(dhidden);
  return simt_eos_pool_bwd(dout, idx, n, seq, H, dhidden, static_cast<cudaStream_t>(stream));
}

int mmg_adamw_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                   const long long* numel, int n_tensors, float lr, const float* lr_dev, float beta1, float beta2,
                   float eps, float weight_decay, long long* step_state, mmg_stream_t stream) {
  if (n_tensors < 0) return set_error(MMG_ERR_BAD_ARG, "mmg_adamw_step: n_tensors < 0");
  if (n_tensors == 0) return 0;
  if (params == nullptr || grads == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr || numel == nullptr)
    return set_error(MMG_ERR_BAD_ARG, "mmg_adamw_step: NULL pointer table");
  if (!(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f) || !(eps >= 0.f) || !(weight_decay >= 0.f) ||
      (lr_dev == nullptr && !(lr >= 0.f)))
    return set_error(MMG_ERR_BAD_ARG, "mmg_adamw_step: invalid hyper-parameter");
  MMG_REQ(step_state);
  if (lr_dev != nullptr) MMG_REQ(lr_dev);
  for (int i = 0; i < n_tensors; ++i) {
    if (numel[i] < 0) return set_error(MMG_ERR_BAD_ARG, "mmg_adamw_step: numel[%d] < 0", i);
    if (numel[i] == 0) continue;
    MMG_REQ(params[i]);
    MMG_REQ(grads[i]);
    MMG_REQ(exp_avg[i]);
    MMG_REQ(exp_avg_sq[i]);
  }
  return simt_adamw(params, grads, exp_avg, exp_avg_sq, numel, n_tensors, lr, lr_dev, beta1, beta2, eps, weight_decay,
                    step_state, static_cast<cudaStream_t>(stream));
}

int mmg_ce_fwd(const float* logits, long long ld, int n, int m, const long long* labels, float coef, float* lse,
               float* loss_out, mmg_stream_t stream) {
  if (n <= 0 || m <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_ce_fwd: bad shape %dx%d", n, m);
  if (labels == nullptr && m < n) return set_error(MMG_ERR_BAD_ARG, "mmg_ce_fwd: arange labels need n <= m (%d, %d)", n, m);
  MMG_REQ(logits);
  MMG_REQ(lse);
  MMG_REQ(loss_out);
  return simt_ce_fwd(logits, ld, n, m, labels, coef, lse, loss_out, static_cast<cudaStream_t>(stream));
}

int mmg_ce_bwd(const float* logits, long long ld, int n, int m, const long long* labels, const float* lse,
               const float* grad_loss, float coef, float* dlogits, long long ldd, mmg_stream_t stream) {
  if (n <= 0 || m <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_ce_bwd: bad shape %dx%d", n, m);
  if (labels == nullptr && m < n) return set_error(MMG_ERR_BAD_ARG, "mmg_ce_bwd: arange labels need n <= m (%d, %d)", n, m);
  MMG_REQ(logits);
  MMG_REQ(lse);
  MMG_REQ(grad_loss);
  MMG_REQ(dlogits);
  return simt_ce_bwd(logits, ld, n, m, labels, lse, grad_loss, coef, dlogits, ldd, static_cast<cudaStream_t>(stream));
}

int mmg_dot_sum(const float* x, const float* y, long long n, float* out, mmg_stream_t stream) {
  if (n < 0) return set_error(MMG_ERR_BAD_ARG, "mmg_dot_sum: negative length");
  MMG_REQ(out);
  if (n > 0) {
    MMG_REQ(x);
    MMG_REQ(y);
  }
  return simt_dot_sum(x, y, n, out, static_cast<cudaStream_t>(stream));
}

int mmg_zeroshot_score(const float* img, const float* txt, int N, int C, int D, const float* scale, float* logits_out,
                       float* probs_out, long long* argmax_out, int k, long long* topk_idx_out, float* topk_val_out,
                       mmg_stream_t stream) {
  if (N < 0 || C <= 0 || D <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_zeroshot_score: bad shape %dx%dx%d", N, C, D);
  if (k < 0 || k > 8 || k > C) return set_error(MMG_ERR_BAD_ARG, "mmg_zeroshot_score: need 0 <= k <= min(8, C)");
  if (N == 0) return 0;
  MMG_REQ(img);
  MMG_REQ(txt);
  MMG_REQ(scale);
  if (C > 64) {
    // more prompts than the tiled kernels take: one-warp-per-row kernel, which parks the row's logits in logits_out
    if (logits_out == nullptr)
      return set_error(MMG_ERR_BAD_ARG, "mmg_zeroshot_score: logits_out is required for more than 64 prompts (got %d)", C);
    return simt_zeroshot_wide(img, txt, N, C, D, scale, logits_out, probs_out, argmax_out, k, topk_idx_out, topk_val_out,
                              static_cast<cudaStream_t>(stream));
  }
  return simt_zeroshot(img, txt, N, C, D, scale, logits_out, probs_out, argmax_out, k, topk_idx_out, topk_val_out,
                       static_cast<cudaStream_t>(stream));
}


size_t mmg_zeroshot_workspace_bytes(int C, int D) { return tc_zeroshot_workspace_bytes(C, D); }

int mmg_zeroshot_score_tc(const float* img, const float* txt, int N, int C, int D, const float* scale, float* logits_out,
                          float* probs_out, long long* argmax_out, int k, long long* topk_idx_out, float* topk_val_out,
                          void* workspace, size_t workspace_bytes, mmg_stream_t stream) {
  if (N < 0 || C <= 0 || D <= 0) return set_error(MMG_ERR_BAD_ARG, "mmg_zeroshot_score_tc: bad shape %dx%dx%d", N, C, D);
  if (C > 64) return set_error(MMG_ERR_UNSUPPORTED_SHAPE, "mmg_zeroshot_score_tc: at most 64 prompts (got %d)", C);
  if (k < 0 || k > 8 || k > C) return set_error(MMG_ERR_BAD_ARG, "mmg_zeroshot_score_tc: need 0 <= k <= min(8, C)");
  if (N == 0) return 0;
  MMG_REQ(img);
  MMG_REQ(txt);
  MMG_REQ(scale);
  MMG_REQ(workspace);
  if (!tc_zeroshot_supported(img, N, C, D))
    return set_error(MMG_ERR_UNSUPPORTED_SHAPE,
                     "mmg_zeroshot_score_tc: needs D %% 4 == 0 and 16-byte aligned embeddings (use mmg_zeroshot_score)");
  return tc_zeroshot(img, txt, N, C, D, scale, logits_out, probs_out, argmax_out, k, topk_idx_out, topk_val_out,
                     workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
