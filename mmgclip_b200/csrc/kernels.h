// mmgclip_b200 -- internal launcher declarations shared by the .cu translation units.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace mmg {

// thread-local error reporting (c_api.cu)
int set_error(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
// every kernel launch of the library is counted (mmg_kernel_launch_count)
void count_launch();

// Programmatic dependent launch (PDL).  The kernels of a training step form a chain on one stream; with the attribute below
// a kernel may be SCHEDULED while its predecessor is still running: its CTAs run their prologue (barrier / TMEM set-up,
// descriptor prefetch -- nothing that touches global memory) and then block in griddepcontrol.wait until the predecessor
// grid has completed and its writes are visible.  Every kernel of the library that takes the attribute executes
// griddepcontrol.wait before its first global access and before it exits, so completion stays transitive along the chain.
// Captured into CUDA graphs as programmatic edges.  mmg_tune("pdl", 0) switches it off.
bool pdl_enabled();
struct PdlAttr {
  cudaLaunchAttribute a[2];
  unsigned n = 0;
  PdlAttr() {
    if (pdl_enabled()) {
      a[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      a[n].val.programmaticStreamSerializationAllowed = 1;
      ++n;
    }
  }
  void cluster(int x) {
    a[n].id = cudaLaunchAttributeClusterDimension;
    a[n].val.clusterDim.x = x;
    a[n].val.clusterDim.y = 1;
    a[n].val.clusterDim.z = 1;
    ++n;
  }
};

// <<<grid, block, smem, st>>> with the PDL attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  PdlAttr at;
  cfg.attrs = at.a;
  cfg.numAttrs = at.n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- tensor-core (bf16, tcgen05) launchers: tc_kernels.cu ----
struct TcOperand {
  const void* ptr;   // bf16
  long long ld;      // row pitch in elements
  int mn_major;      // 0: [rows, K]  1: [K, rows]
  int f16 = 0;       // element type: 0 bf16, 1 fp16
};

int tc_gemm_store(const TcOperand& A, const TcOperand& B, float* C, long long ldc, int M, int N, int K, float alpha,
                  const float* alpha_dev, const float* bias, int relu, int mode, int k_splits, cudaStream_t st);

// Same with up to three K segments (operand pairs chosen per segment by the bit masks), e.g. the split-precision heads'
// A_hi.B_hi + A_hi.B_lo + A_lo.B_hi as ONE contraction over 3K.  A1 / B1 may be NULL when no segment refers to them.
int tc_gemm_store_seg(const TcOperand& A0, const TcOperand* A1, const TcOperand& B0, const TcOperand* B1, int nseg,
                      int seg_a, int seg_b, float* C, long long ldc, int M, int N, int K, float alpha,
                      const float* alpha_dev, const float* bias, int relu, int mode, int k_splits, cudaStream_t st);

// Two independent accumulate-GEMMs in one launch (dA and dB of one logit block).
int tc_gemm_dual_accumulate(const TcOperand& A0, const TcOperand& B0, float* C0, long long ldc0, int M0, int N0, int K0,
                            const TcOperand& A1, const TcOperand& B1, float* C1, long long ldc1, int M1, int N1, int K1,
                            cudaStream_t st, const float* alpha_dev = nullptr);

int tc_infonce_fwd(const void* a_hat, const void* b_hat, int rows, int cols, int D, int diag_offset, const float* scale,
                   float* rowsum, float* colsum, float* diag, void* e_out, long long lde, cudaStream_t st, int emb_f16 = 0);

int tc_infonce_grad_block(const void* a_blk, const void* b_blk, int rb, int cb, int D, int diag_offset,
                          const float* scale, const float* rinv, const float* cinv, const float* scal, void* G,
                          long long ldg, float* dlogscale_acc, cudaStream_t st, int emb_f16 = 0);

// The whole bf16 backward as one persistent launch (bwd_fused.cuh).  *used = 0 (and nothing launched) when the shape is
// not covered (rows / cols / D not multiples of 256, workspace too small, MMG_BWD_FUSED=0): the caller falls back to the
// block loop.  tc_infonce_bwd_fused_workspace = bytes it needs (0 = not covered).
size_t tc_infonce_bwd_fused_workspace(int rows, int cols, int D);
// dB_owners[i] = fp32 buffer that owns column-side rows [i * cols / n_owners, (i+1) * ...): one owner = all of dB; in the
// row-sharded run the owners are the ranks (pointers into a staging buffer, or NVLink peer mappings).  n_parts > 1: the
// launch covers only part `part` of every owner's columns and dB_owners[i] is the [cols / n_owners / n_parts, D] buffer
// of that part.
int tc_infonce_bwd_fused(const void* a_hat, const void* b_hat, int rows, int cols, int D, int diag_offset,
                         const float* scale, const float* rinv, const float* cinv, const float* scal, float* dA,
                         float* const* dB_owners, int n_owners, int n_parts, int part, float* dlogscale_acc,
                         void* workspace, size_t workspace_bytes, cudaStream_t st, int* used,
                         const void* e_stored = nullptr, long long lde = 0, int emb_f16 = 0);
int tc_infonce_stored_supported(int rows, int cols, int D, int n_owners, int n_parts);
int tc_fused_trace_region(int rows, int cols, int D, size_t* off, size_t* bytes, int* per_role, int* roles);

// Zero-shot prompt scoring on the tensor pipe (zeroshot_tc.cu): 3xTF32, fp32-faithful, HBM-bound at large N.
size_t tc_zeroshot_workspace_bytes(int C, int D);
bool tc_zeroshot_supported(const float* img, int N, int C, int D);
int tc_zeroshot(const float* img, const float* txt, int N, int C, int D, const float* scale, float* logits_out,
                float* probs_out, long long* argmax_out, int k, long long* topk_idx, float* topk_val, void* workspace,
                size_t workspace_bytes, cudaStream_t st);

int tc_tune(const char* key, int value);
int tc_fused_bwd_schedule(int rows, int cols, int D, int n_owners, int n_parts, int part, int pairs, int pair, int* items,
                          int max_items, int* info);

// ---- SIMT (fp32) launchers: simt_kernels.cu ----
int simt_gemm(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn, float* C, long long ldc,
              int M, int N, int K, float alpha, const float* alpha_dev, const float* bias, int relu, int mode,
              int k_splits, cudaStream_t st);
int simt_cast_bf16(const float* x, void* y, long long n, cudaStream_t st);
int simt_cast_f16(const float* x, void* y, long long n, cudaStream_t st);
int simt_cast_split(const float* x, void* hi, void* lo, long long n, cudaStream_t st);
int simt_l2norm_fwd(const float* u, int B, int D, float* y, float* inv_norm, void* y_16, int y16_f16, cudaStream_t st);
int simt_push_rows(const void* src, long long bytes, void* const* dst_ptrs, int n_dst, long long dst_offset_bytes,
                   cudaStream_t st);
int simt_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, int B, int D, float* du, void* du_bf16,
                    void* du_bf16_lo, float* zero, long long zero_floats, cudaStream_t st);
int simt_dropout_apply(float* y, const uint8_t* mask, float keep_scale, long long n, cudaStream_t st);
int simt_dropout_draw_apply(float* y, uint8_t* mask, float p, long long n, unsigned long long* state, cudaStream_t st);
int simt_relu_dropout_bwd(const float* dy, const float* y, const uint8_t* mask, float keep_scale, float* dz,
                          long long n, cudaStream_t st);
int simt_colsum(const float* x, int rows, int cols, float* out, cudaStream_t st);
int simt_add(const float* x, const float* y, float* out, long long n, cudaStream_t st);
int simt_gelu_fwd(const float* x, float* y, long long n, cudaStream_t st);
int simt_gelu_bwd(const float* dy, const float* x, float* dx, long long n, cudaStream_t st);
int simt_layernorm_fwd(const float* x, const float* gamma, const float* beta, int rows, int cols, float eps, float* y,
                       float* mean, float* rstd, cudaStream_t st);
int simt_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                       int rows, int cols, float* dx, float* dgamma, float* dbeta, cudaStream_t st);

int simt_eos_pool(const float* hidden, const long long* mask, int n, int seq, int H, float* out, long long* idx_out,
                  cudaStream_t st);
int simt_eos_pool_bwd(const float* dout, const long long* idx, int n, int seq, int H, float* dhidden, cudaStream_t st);
constexpr int kAdamwMaxTensors = 64;  // tensors per launch (pointers travel as kernel parameters)
int simt_adamw(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
               const long long* numel, int n_tensors, float lr, const float* lr_dev, float beta1, float beta2,
               float eps, float weight_decay, long long* state, cudaStream_t st);

// fp32 InfoNCE block helpers operating on an fp32 cosine block S[rb, cb] (pitch lds) in scratch
int simt_lse_block(float* S, long long lds, int rb, int cb, int row0, int col0, int diag_offset, const float* scale,
                   float* rowsum, float* colsum, float* diag, cudaStream_t st);
int simt_grad_block(float* S, long long lds, int rb, int cb, int row0, int col0, int diag_offset, const float* scale,
                    const float* rinv, const float* cinv, const float* scal, float* dlogscale_acc, cudaStream_t st);

int simt_infonce_loss(const float* rowsum, const float* colsum, const float* diag, int n, const float* scale,
                      float inv_two_b, float* loss_out, cudaStream_t st);
int simt_infonce_row_part(const float* rowsum, const float* diag, int rows, float* part_out, cudaStream_t st);
int simt_infonce_loss_cols(const float* colsum, int cols, const float* scale, const float* row_part, float inv_two_b,
                           float* loss_out, cudaStream_t st);
int simt_infonce_bwd_prep(const float* rowsum, int rows, const float* colsum, int cols, const float* scale,
                          const float* grad_loss, float inv_two_b, int diag_in_fp32, int f16_scaled, float* rinv,
                          float* cinv, float* scal, cudaStream_t st);
int simt_infonce_bwd_diag(const float* a32, const float* b32, int rows, int D, const float* diag, const float* scale,
                          const float* rinv, const float* cinvm, const float* scal, float* dA, float* dB,
                          float* dlogscale_acc, int init, cudaStream_t st);
int simt_infonce_bwd_prep_diag(const float* rowsum, int rows, const float* colsum, int cols, int diag_offset,
                               const float* scale, const float* grad_loss, float inv_two_b, int f16_scaled, float* rinv,
                               float* cinv, float* scal, const float* a32, const float* b32, int D, const float* diag,
                               float* dA, float* dB, float* dlogscale_acc, cudaStream_t st);
int simt_dot_sum(const float* x, const float* y, long long n, float* out, cudaStream_t st);
int simt_ce_fwd(const float* logits, long long ld, int n, int m, const long long* labels, float coef, float* lse,
                float* loss_out, cudaStream_t st);
int simt_ce_bwd(const float* logits, long long ld, int n, int m, const long long* labels, const float* lse,
                const float* grad_loss, float coef, float* dlogits, long long ldd, cudaStream_t st);
int simt_zeroshot_wide(const float* img, const float* txt, int N, int C, int D, const float* scale, float* logits,
                       float* probs_out, long long* argmax_out, int k, long long* topk_idx, float* topk_val,
                       cudaStream_t st);
int simt_zeroshot(const float* img, const float* txt, int N, int C, int D, const float* scale, float* logits_out,
                  float* probs_out, long long* argmax_out, int k, long long* topk_idx, float* topk_val,
                  cudaStream_t st);

}  // namespace mmg
