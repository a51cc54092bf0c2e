// mmgclip_b200 -- fp32 / SIMT kernels.
//
//  * the reference-faithful fp32 arithmetic path (MMG_PREC_FP32): FFMA contraction + fp32 block helpers of the
//    fused InfoNCE, accurate expf/logf, deterministic reductions where it is cheap;
//  * the bandwidth-bound small kernels both precisions share: casts, row L2-normalise forward/backward, ReLU/dropout
//    backward, bias gradient, GELU, LayerNorm, loss finalisation, literal cross-entropy on materialised logits;
//  * zero-shot prompt scoring (fp32 logits, fused softmax / argmax / top-k with a defined tie rule).
//
// Warp-level primitives (shuffles) do the row reductions; loads are 16-byte vectorised where alignment allows.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace mmg {

#define MMG_LAUNCH_CHECK(what)                         \
  do {                                                 \
    cudaError_t e__ = cudaGetLastError();              \
    if (e__ != cudaSuccess) return check_cuda(e__, what); \
    count_launch();                                    \
  } while (0)

// launch with the programmatic-dependent-launch attribute (kernels.h); the kernel calls pdl_entry() first thing
#define MMG_LAUNCH_PDL(what, kern, grid, block, smem, st, ...)                                        \
  do {                                                                                                \
    cudaError_t e__ = launch_pdl(kern, dim3(grid), dim3(block), smem, st, __VA_ARGS__);               \
    if (e__ != cudaSuccess) return check_cuda(e__, what);                                             \
    count_launch();                                                                                   \
  } while (0)

// =====================================================================================================
// fp32 contraction  C[M,N] (op)= alpha * A . B^T (+bias)(ReLU)
// 64x64 block tile, 16-deep K slab, 256 threads x (4x4) register micro-tile.
// =====================================================================================================
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb,
             float* __restrict__ C, long long ldc, int M, int N, int K, float alpha_host,
             const float* __restrict__ alpha_dev, const float* __restrict__ bias, int relu, int mode,
             int k_per_split) {
  const float alpha = alpha_dev ? alpha_host * (*alpha_dev) : alpha_host;
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int m, k;
      if (A_MN) { k = e >> 6; m = e & 63; } else { m = e >> 4; k = e & 15; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < kend) v = A_MN ? A[(long long)gk * lda + gm] : A[(long long)gm * lda + gk];
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int n, k;
      if (B_MN) { k = e >> 6; n = e & 63; } else { n = e >> 4; k = e & 15; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < kend) v = B_MN ? B[(long long)gk * ldb + gn] : B[(long long)gn * ldb + gk];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float* dst = C + (long long)gm * ldc + gn;
      float v = acc[i][j] * alpha;
      if (mode == 1) v += *dst;  // accumulate first, then bias / activation: C = act(C + alpha*acc + bias)
      if (bias != nullptr && blockIdx.z == 0) v += bias[gn];
      if (relu) v = fmaxf(v, 0.f);
      if (mode == 2) atomicAdd(dst, v);
      else *dst = v;
    }
  }
}

int simt_gemm(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn, float* C, long long ldc,
              int M, int N, int K, float alpha, const float* alpha_dev, const float* bias, int relu, int mode,
              int k_splits, cudaStream_t st) {
  if (k_splits < 1) k_splits = 1;
  if (k_splits > 1 && mode != 2) return set_error(-1, "simt_gemm: k_splits > 1 needs MMG_ATOMIC_ADD");
  if (k_splits > 1 && relu) return set_error(-1, "simt_gemm: ReLU cannot be fused with split-K");
  int k_per_split = ((K + k_splits - 1) / k_splits + 15) / 16 * 16;
  if (k_per_split < 16) k_per_split = 16;
  k_splits = (K + k_per_split - 1) / k_per_split;
  if (k_splits < 1) k_splits = 1;
  dim3 grid((N + 63) / 64, (M + 63) / 64, k_splits);
  if (grid.y > 65535) return set_error(-3, "simt_gemm: M too large for the SIMT grid (%d)", M);
#define MMG_SGEMM(AM, BM) \
  sgemm_kernel<AM, BM><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, alpha, alpha_dev, bias, relu, \
                                             mode, k_per_split)
  if (a_mn && b_mn) MMG_SGEMM(true, true);
  else if (a_mn) MMG_SGEMM(true, false);
  else if (b_mn) MMG_SGEMM(false, true);
  else MMG_SGEMM(false, false);
#undef MMG_SGEMM
  MMG_LAUNCH_CHECK("sgemm_kernel");
  return 0;
}

// =====================================================================================================
// casts
// =====================================================================================================
__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  pdl_entry();
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(y) & 7) == 0);
  if (vec) {
    const long long n4 = n >> 2;
    for (long long j = i; j < n4; j += stride) {
      const float4 v = reinterpret_cast<const float4*>(x)[j];
      uint2 o;
      o.x = pack_bf16x2(v.x, v.y);
      o.y = pack_bf16x2(v.z, v.w);
      reinterpret_cast<uint2*>(y)[j] = o;
    }
    for (long long j = (n4 << 2) + i; j < n; j += stride) y[j] = __float2bfloat16_rn(x[j]);
  } else {
    for (long long j = i; j < n; j += stride) y[j] = __float2bfloat16_rn(x[j]);
  }
}

// fp32 -> fp16 (the operand copy of L2-normalised embeddings: |x| <= 1, so fp16's 11-bit significand costs no range)
__global__ void cast_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, long long n) {
  pdl_entry();
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(y) & 7) == 0);
  if (vec) {
    const long long n4 = n >> 2;
    for (long long j = i; j < n4; j += stride) {
      const float4 v = reinterpret_cast<const float4*>(x)[j];
      const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&lo);
      o.y = *reinterpret_cast<const uint32_t*>(&hi);
      reinterpret_cast<uint2*>(y)[j] = o;
    }
    for (long long j = (n4 << 2) + i; j < n; j += stride) y[j] = __float2half_rn(x[j]);
  } else {
    for (long long j = i; j < n; j += stride) y[j] = __float2half_rn(x[j]);
  }
}

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits survive, so the three-product contraction
// A_hi.B_hi + A_hi.B_lo + A_lo.B_hi on the bf16 tensor pipe is accurate to ~2^-17 per operand ("bf16x3").
__global__ void cast_split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, long long n) {
  pdl_entry();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(hi) & 7) == 0) &&
                   ((reinterpret_cast<uintptr_t>(lo) & 7) == 0);
  long long done = 0;
  if (vec) {
    const long long n4 = n >> 2;
    for (long long j = i0; j < n4; j += stride) {
      const float4 v = reinterpret_cast<const float4*>(x)[j];
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z),
                          h3 = __float2bfloat16_rn(v.w);
      uint2 oh, ol;
      oh.x = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
      oh.y = static_cast<uint32_t>(__bfloat16_as_ushort(h2)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h3)) << 16);
      ol.x = pack_bf16x2(v.x - __bfloat162float(h0), v.y - __bfloat162float(h1));
      ol.y = pack_bf16x2(v.z - __bfloat162float(h2), v.w - __bfloat162float(h3));
      reinterpret_cast<uint2*>(hi)[j] = oh;
      reinterpret_cast<uint2*>(lo)[j] = ol;
    }
    done = n4 << 2;
  }
  for (long long i = done + i0; i < n; i += stride) {
    const float v = x[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

static inline int ew_blocks(long long n, int per_thread) {
  long long b = (n / per_thread + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}

int simt_cast_split(const float* x, void* hi, void* lo, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  MMG_LAUNCH_PDL("cast_split_kernel", cast_split_kernel, ew_blocks(n, 4), 256, 0, st, x, reinterpret_cast<__nv_bfloat16*>(hi),
                 reinterpret_cast<__nv_bfloat16*>(lo), n);
  return 0;
}

int simt_cast_f16(const float* x, void* y, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  MMG_LAUNCH_PDL("cast_f16_kernel", cast_f16_kernel, ew_blocks(n, 4), 256, 0, st, x, reinterpret_cast<__half*>(y), n);
  return 0;
}

int simt_cast_bf16(const float* x, void* y, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  MMG_LAUNCH_PDL("cast_bf16_kernel", cast_bf16_kernel, ew_blocks(n, 4), 256, 0, st, x, reinterpret_cast<__nv_bfloat16*>(y), n);
  return 0;
}

// =====================================================================================================
// push all-gather: copy a contiguous block of rows into the same place of up to 8 destination buffers (this rank's and its
// peers' symmetric buffers, mapped over NVLink) -- 16-byte loads, one 16-byte store per destination.  Followed by a
// cross-rank barrier this IS the all-gather of the column-side embeddings: every rank receives world-1 shards at NVLink
// line rate with no protocol, and the copy is one launch inside the step's graph.
// =====================================================================================================
struct PushDst {
  void* p[8];
};

// A few CTAs per destination (NCCL-like footprint: the copy shares the GPU with the other head's projection, and a
// handful of SMs saturates a link): CTA b serves destination b % n_dst, stripe b / n_dst of kPushStripes.
constexpr int kPushStripes = 4;
constexpr int kPushThreads = 512;

__global__ void __launch_bounds__(kPushThreads)
push_rows_kernel(const uint4* __restrict__ src, long long n16, const __grid_constant__ PushDst dst, int n_dst,
                 long long dst_off16) {
  pdl_entry();
  const int d = blockIdx.x % n_dst;
  const int stripe = blockIdx.x / n_dst;
  const long long per = (n16 + kPushStripes - 1) / kPushStripes;
  const long long i0 = per * stripe;
  const long long i1 = i0 + per < n16 ? i0 + per : n16;
  uint4* __restrict__ out = reinterpret_cast<uint4*>(dst.p[d]) + dst_off16;
  long long i = i0 + threadIdx.x;
  // four independent 16-byte loads in flight per thread
  for (; i + 3 * kPushThreads < i1; i += 4 * kPushThreads) {
    const uint4 v0 = src[i], v1 = src[i + kPushThreads], v2 = src[i + 2 * kPushThreads], v3 = src[i + 3 * kPushThreads];
    out[i] = v0;
    out[i + kPushThreads] = v1;
    out[i + 2 * kPushThreads] = v2;
    out[i + 3 * kPushThreads] = v3;
  }
  for (; i < i1; i += kPushThreads) out[i] = src[i];
}

int simt_push_rows(const void* src, long long bytes, void* const* dst_ptrs, int n_dst, long long dst_offset_bytes,
                   cudaStream_t st) {
  PushDst d;
  for (int i = 0; i < 8; ++i) d.p[i] = i < n_dst ? dst_ptrs[i] : nullptr;
  const long long n16 = bytes / 16;
  MMG_LAUNCH_PDL("push_rows_kernel", push_rows_kernel, n_dst * kPushStripes, kPushThreads, 0, st,
                 reinterpret_cast<const uint4*>(src), n16, d, n_dst, dst_offset_bytes / 16);
  return 0;
}

// =====================================================================================================
// row L2 normalise (one warp per row, float4 loads, shuffle reduction)
// =====================================================================================================
__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const float* __restrict__ u, int B, int D, float* __restrict__ y, float* __restrict__ inv_norm,
                  uint16_t* __restrict__ yb, int yb_f16) {
  pdl_entry();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* ur = u + (long long)row * D;
  const bool vec = ((D & 3) == 0) && ((reinterpret_cast<uintptr_t>(u) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
  float ss = 0.f;
  if (vec) {
    for (int i = lane; i < (D >> 2); i += 32) {
      const float4 v = reinterpret_cast<const float4*>(ur)[i];
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  } else {
    for (int i = lane; i < D; i += 32) ss += ur[i] * ur[i];
  }
  ss = warp_sum(ss);
  // Reference: x / x.norm(dim=1, keepdim=True) with no epsilon (a zero row gives NaN there and here).
  const float nrm = sqrtf(ss);
  const float inv = 1.0f / nrm;
  if (lane == 0 && inv_norm != nullptr) inv_norm[row] = inv;
  float* yr = y + (long long)row * D;
  uint16_t* ybr = yb ? yb + (long long)row * D : nullptr;  // 16-bit operand copy: fp16 (yb_f16) or bf16
  if (vec) {
    for (int i = lane; i < (D >> 2); i += 32) {
      float4 v = reinterpret_cast<const float4*>(ur)[i];
      v.x = v.x / nrm; v.y = v.y / nrm; v.z = v.z / nrm; v.w = v.w / nrm;
      reinterpret_cast<float4*>(yr)[i] = v;
      if (ybr) {
        uint2 o;
        if (yb_f16) {
          const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
          o.x = *reinterpret_cast<const uint32_t*>(&lo);
          o.y = *reinterpret_cast<const uint32_t*>(&hi);
        } else {
          o.x = pack_bf16x2(v.x, v.y);
          o.y = pack_bf16x2(v.z, v.w);
        }
        reinterpret_cast<uint2*>(ybr)[i] = o;
      }
    }
  } else {
    for (int i = lane; i < D; i += 32) {
      const float v = ur[i] / nrm;
      yr[i] = v;
      if (ybr) ybr[i] = yb_f16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
    }
  }
}

int simt_l2norm_fwd(const float* u, int B, int D, float* y, float* inv_norm, void* y_16, int y16_f16, cudaStream_t st) {
  if (B <= 0) return 0;
  MMG_LAUNCH_PDL("l2norm_fwd_kernel", l2norm_fwd_kernel, (B + 7) / 8, 256, 0, st, u, B, D, y, inv_norm,
                 reinterpret_cast<uint16_t*>(y_16), y16_f16);
  return 0;
}

// `zero` (nullable, 16-byte aligned, zero_n4 float4 elements): a buffer this launch also clears -- the split-K output of
// the weight-gradient contraction that follows it accumulates with reduce-adds and would otherwise need a fill launch.
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ inv_norm, int B,
                  int D, float* __restrict__ du, __nv_bfloat16* __restrict__ dub, __nv_bfloat16* __restrict__ dul,
                  float4* __restrict__ zero, long long zero_n4) {
  pdl_entry();
  if (zero != nullptr) {
    const long long per = (zero_n4 + gridDim.x - 1) / gridDim.x;
    const long long z0 = per * blockIdx.x;
    const long long z1 = z0 + per < zero_n4 ? z0 + per : zero_n4;
    for (long long i = z0 + threadIdx.x; i < z1; i += 256) zero[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* dyr = dy + (long long)row * D;
  const float* yr = y + (long long)row * D;
  float dot = 0.f;
  for (int i = lane; i < D; i += 32) dot = fmaf(dyr[i], yr[i], dot);
  dot = warp_sum(dot);
  const float inv = inv_norm[row];
  for (int i = lane; i < D; i += 32) {
    const float v = (dyr[i] - yr[i] * dot) * inv;
    if (du) du[(long long)row * D + i] = v;
    if (dub) {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      dub[(long long)row * D + i] = h;
      if (dul) dul[(long long)row * D + i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

int simt_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, int B, int D, float* du, void* du_bf16,
                    void* du_bf16_lo, float* zero, long long zero_floats, cudaStream_t st) {
  if (B <= 0) return 0;
  MMG_LAUNCH_PDL("l2norm_bwd_kernel", l2norm_bwd_kernel, (B + 7) / 8, 256, 0, st, dy, y, inv_norm, B, D, du, reinterpret_cast<__nv_bfloat16*>(du_bf16),
                                                 reinterpret_cast<__nv_bfloat16*>(du_bf16_lo),
                                                 reinterpret_cast<float4*>(zero), zero ? zero_floats / 4 : 0);
  return 0;
}

// =====================================================================================================
// MultiLinearHead hidden-layer helpers
// =====================================================================================================
__global__ void dropout_apply_kernel(float* __restrict__ y, const uint8_t* __restrict__ mask, float keep_scale,
                                     long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = mask[i] ? y[i] * keep_scale : 0.f;
}

int simt_dropout_apply(float* y, const uint8_t* mask, float keep_scale, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  dropout_apply_kernel<<<ew_blocks(n, 4), 256, 0, st>>>(y, mask, keep_scale, n);
  MMG_LAUNCH_CHECK("dropout_apply_kernel");
  return 0;
}

// Inverted dropout with the keep mask drawn IN the kernel (the reference's nn.Dropout, projection.py:51,59,92,98): one launch
// draws, applies and records the mask.  Philox4x32-10 (counter-based, the generator family torch uses on CUDA): element i
// takes word i % 4 of the block with counter offset + i / 4 under the key `seed`; an element is kept when its 32-bit word
// >= p * 2^32.  state = {seed, offset, ticket} lives on the device: the last CTA to retire advances the offset by the
// blocks this launch consumed, so a replayed CUDA graph draws fresh masks with no host involvement.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void __launch_bounds__(256)
dropout_draw_apply_kernel(float* __restrict__ y, uint8_t* __restrict__ mask, float p, float keep_scale, long long n,
                          unsigned long long* __restrict__ state) {
  pdl_entry();
  const unsigned long long seed = state[0], offset = state[1];
  const uint32_t thr = p >= 1.f ? 0xFFFFFFFFu : static_cast<uint32_t>(static_cast<double>(p) * 4294967296.0);
  const bool drop_all = p >= 1.f;
  const long long nblk = (n + 3) >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nblk; b += stride) {
    const unsigned long long ctr = offset + static_cast<unsigned long long>(b);
    uint32_t r[4];
    philox4x32_10(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), 0u, 0u, static_cast<uint32_t>(seed),
                  static_cast<uint32_t>(seed >> 32), r);
    const long long i0 = b << 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = i0 + j;
      if (i < n) {
        const bool keep = !drop_all && r[j] >= thr;
        mask[i] = keep ? 1 : 0;
        y[i] = keep ? y[i] * keep_scale : 0.f;
      }
    }
  }
  // the last CTA to retire advances the stream position (every CTA has read `offset` before it could retire)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(&state[2], 1ull);
    if (t == gridDim.x - 1) {
      state[1] = offset + static_cast<unsigned long long>(nblk);
      state[2] = 0ull;
      __threadfence();
    }
  }
}

int simt_dropout_draw_apply(float* y, uint8_t* mask, float p, long long n, unsigned long long* state, cudaStream_t st) {
  if (n <= 0) return 0;
  const float keep_scale = p < 1.f ? 1.0f / (1.0f - p) : 0.f;
  MMG_LAUNCH_PDL("dropout_draw_apply_kernel", dropout_draw_apply_kernel, ew_blocks(n, 4), 256, 0, st, y, mask, p, keep_scale,
                 n, state);
  return 0;
}

// y is the layer output AFTER ReLU and dropout (y > 0 <=> pre-activation > 0 and kept); y == NULL means no ReLU.
__global__ void relu_dropout_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                        const uint8_t* __restrict__ mask, float keep_scale, float* __restrict__ dz,
                                        long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float g = (y == nullptr || y[i] > 0.f) ? dy[i] : 0.f;
    if (mask != nullptr) g = mask[i] ? g * keep_scale : 0.f;
    dz[i] = g;
  }
}

int simt_relu_dropout_bwd(const float* dy, const float* y, const uint8_t* mask, float keep_scale, float* dz,
                          long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  relu_dropout_bwd_kernel<<<ew_blocks(n, 4), 256, 0, st>>>(dy, y, mask, keep_scale, dz, n);
  MMG_LAUNCH_CHECK("relu_dropout_bwd_kernel");
  return 0;
}

// out[c] = sum_r x[r, c].  Block = 32 columns x 8 row groups; fixed-order shared-memory combine (deterministic).
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, int rows, int cols, float* __restrict__ out) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < cols)
    for (int r = ty; r < rows; r += 8) s += x[(long long)r * cols + c];
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][tx];
    out[c] = t;
  }
}

int simt_colsum(const float* x, int rows, int cols, float* out, cudaStream_t st) {
  if (cols <= 0) return 0;
  colsum_kernel<<<(cols + 31) / 32, 256, 0, st>>>(x, rows, cols, out);
  MMG_LAUNCH_CHECK("colsum_kernel");
  return 0;
}

// =====================================================================================================
// MLPProjectionHead helpers: exact-erf GELU, LayerNorm
// =====================================================================================================
__global__ void gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    y[i] = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
  }
}
__global__ void gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx,
                                long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * expf(-0.5f * v * v);
    dx[i] = dy[i] * (cdf + v * pdf);
  }
}
__global__ void add_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out,
                           long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = x[i] + y[i];
}
int simt_add(const float* x, const float* y, float* out, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  add_kernel<<<ew_blocks(n, 4), 256, 0, st>>>(x, y, out, n);
  MMG_LAUNCH_CHECK("add_kernel");
  return 0;
}
int simt_gelu_fwd(const float* x, float* y, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  gelu_fwd_kernel<<<ew_blocks(n, 4), 256, 0, st>>>(x, y, n);
  MMG_LAUNCH_CHECK("gelu_fwd_kernel");
  return 0;
}
int simt_gelu_bwd(const float* dy, const float* x, float* dx, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  gelu_bwd_kernel<<<ew_blocks(n, 4), 256, 0, st>>>(dy, x, dx, n);
  MMG_LAUNCH_CHECK("gelu_bwd_kernel");
  return 0;
}

__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     int rows, int cols, float eps, float* __restrict__ y, float* __restrict__ mean,
                     float* __restrict__ rstd) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (long long)row * cols;
  float s = 0.f;
  for (int i = lane; i < cols; i += 32) s += xr[i];
  const float mu = warp_sum(s) / cols;
  float v = 0.f;
  for (int i = lane; i < cols; i += 32) {
    const float d = xr[i] - mu;
    v = fmaf(d, d, v);
  }
  const float rs = rsqrtf(warp_sum(v) / cols + eps);
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
  for (int i = lane; i < cols; i += 32) y[(long long)row * cols + i] = (xr[i] - mu) * rs * gamma[i] + beta[i];
}

__global__ void __launch_bounds__(256)
layernorm_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                        const float* __restrict__ mean, const float* __restrict__ rstd, int rows, int cols,
                        float* __restrict__ dx) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (long long)row * cols;
  const float* dyr = dy + (long long)row * cols;
  const float mu = mean[row], rs = rstd[row];
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < cols; i += 32) {
    const float g = dyr[i] * gamma[i];
    const float xh = (xr[i] - mu) * rs;
    s1 += g;
    s2 = fmaf(g, xh, s2);
  }
  s1 = warp_sum(s1) / cols;
  s2 = warp_sum(s2) / cols;
  for (int i = lane; i < cols; i += 32) {
    const float g = dyr[i] * gamma[i];
    const float xh = (xr[i] - mu) * rs;
    dx[(long long)row * cols + i] = (g - s1 - xh * s2) * rs;
  }
}

__global__ void __launch_bounds__(256)
layernorm_bwd_params_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                            const float* __restrict__ rstd, int rows, int cols, float* __restrict__ dgamma,
                            float* __restrict__ dbeta) {
  __shared__ float pg[8][33], pb[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float sg = 0.f, sb = 0.f;
  if (c < cols)
    for (int r = ty; r < rows; r += 8) {
      const float d = dy[(long long)r * cols + c];
      sg = fmaf(d, (x[(long long)r * cols + c] - mean[r]) * rstd[r], sg);
      sb += d;
    }
  pg[ty][tx] = sg;
  pb[ty][tx] = sb;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a += pg[i][tx];
      b += pb[i][tx];
    }
    dgamma[c] = a;
    dbeta[c] = b;
  }
}

int simt_layernorm_fwd(const float* x, const float* gamma, const float* beta, int rows, int cols, float eps, float* y,
                       float* mean, float* rstd, cudaStream_t st) {
  if (rows <= 0) return 0;
  layernorm_fwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, gamma, beta, rows, cols, eps, y, mean, rstd);
  MMG_LAUNCH_CHECK("layernorm_fwd_kernel");
  return 0;
}
int simt_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                       int rows, int cols, float* dx, float* dgamma, float* dbeta, cudaStream_t st) {
  if (rows <= 0) return 0;
  layernorm_bwd_dx_kernel<<<(rows + 7) / 8, 256, 0, st>>>(dy, x, gamma, mean, rstd, rows, cols, dx);
  MMG_LAUNCH_CHECK("layernorm_bwd_dx_kernel");
  layernorm_bwd_params_kernel<<<(cols + 31) / 32, 256, 0, st>>>(dy, x, mean, rstd, rows, cols, dgamma, dbeta);
  MMG_LAUNCH_CHECK("layernorm_bwd_params_kernel");
  return 0;
}

// =====================================================================================================
// fp32 InfoNCE block helpers.  S holds cosines of logit block (rows row0.., cols col0..).
// =====================================================================================================
// rows: one warp per row -> E = exp(s*cos - s) written back in place, rowsum[row0+r] += sum_c E, diag.
__global__ void __launch_bounds__(256)
lse_rows_kernel(float* __restrict__ S, long long lds, int rb, int cb, int row0, int col0, int diag_offset,
                const float* __restrict__ scale, float* __restrict__ rowsum, float* __restrict__ diag) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rb) return;
  const float s = *scale;
  float* Sr = S + (long long)r * lds;
  const int dcol = row0 + r + diag_offset - col0;  // column inside this block holding the matching pair
  float acc = 0.f;
  for (int c = lane; c < cb; c += 32) {
    const float cosv = Sr[c];
    if (c == dcol) diag[row0 + r] = s * cosv;
    const float e = expf(s * cosv - s);
    Sr[c] = e;
    acc += e;
  }
  acc = warp_sum(acc);
  if (lane == 0) rowsum[row0 + r] += acc;
}

// columns: block = 32 columns x 8 row groups over the E block; fixed-order combine.
__global__ void __launch_bounds__(256)
lse_cols_kernel(const float* __restrict__ E, long long lds, int rb, int cb, int col0, float* __restrict__ colsum) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < cb)
    for (int r = ty; r < rb; r += 8) s += E[(long long)r * lds + c];
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cb) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][tx];
    colsum[col0 + c] += t;
  }
}

int simt_lse_block(float* S, long long lds, int rb, int cb, int row0, int col0, int diag_offset, const float* scale,
                   float* rowsum, float* colsum, float* diag, cudaStream_t st) {
  lse_rows_kernel<<<(rb + 7) / 8, 256, 0, st>>>(S, lds, rb, cb, row0, col0, diag_offset, scale, rowsum, diag);
  MMG_LAUNCH_CHECK("lse_rows_kernel");
  lse_cols_kernel<<<(cb + 31) / 32, 256, 0, st>>>(S, lds, rb, cb, col0, colsum);
  MMG_LAUNCH_CHECK("lse_cols_kernel");
  return 0;
}

// cos -> g in place (see EpiGrad in gemm_tc.cuh for the formula); accumulates sum g*cos.
__global__ void __launch_bounds__(256)
grad_block_kernel(float* __restrict__ S, long long lds, int rb, int cb, int row0, int col0, int diag_offset,
                  const float* __restrict__ scale, const float* __restrict__ rinv, const float* __restrict__ cinv,
                  const float* __restrict__ scal, float* __restrict__ dlogscale_acc) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  float dacc = 0.f;
  if (r < rb) {
    const float s = *scale;
    const float dcoef = scal[0];
    const float ri = rinv[row0 + r];
    float* Sr = S + (long long)r * lds;
    const int dcol = row0 + r + diag_offset - col0;
    for (int c = lane; c < cb; c += 32) {
      const float cosv = Sr[c];
      float g = expf(s * cosv - s) * (ri + cinv[col0 + c]);
      if (c == dcol) g -= dcoef;
      Sr[c] = g;
      dacc = fmaf(g, cosv, dacc);
    }
  }
  dacc = warp_sum(dacc);
  __shared__ float wsum[8];
  if (lane == 0) wsum[threadIdx.x >> 5] = dacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += wsum[i];
    if (t != 0.f && dlogscale_acc != nullptr) atomicAdd(dlogscale_acc, t);
  }
}

int simt_grad_block(float* S, long long lds, int rb, int cb, int row0, int col0, int diag_offset, const float* scale,
                    const float* rinv, const float* cinv, const float* scal, float* dlogscale_acc, cudaStream_t st) {
  grad_block_kernel<<<(rb + 7) / 8, 256, 0, st>>>(S, lds, rb, cb, row0, col0, diag_offset, scale, rinv, cinv, scal,
                                                  dlogscale_acc);
  MMG_LAUNCH_CHECK("grad_block_kernel");
  return 0;
}

// =====================================================================================================
// loss finalisation (single block, fixed-order tree => deterministic)
// =====================================================================================================
// Guard of the fixed softmax shift m = s (E = exp(logit - s), valid for unit-norm inputs and moderate s): a row or column
// whose terms ALL underflowed (sum == 0: s * (1 - max cos) > ~87) or overflowed (un-normalised inputs) cannot be
// normalised; instead of letting log(0) / coef/0 leak inf into the loss and the gradients silently, the loss is set to NaN
// here (the reference's per-row-max cross-entropy would still be finite -- mmgclip_b200.ops.info_nce documents the range
// and the materialised fallback).
// Deterministic sum over a thread-block CLUSTER of kLossCtas x 1024 threads: warp shuffles, one warp over the 32 warp sums,
// then CTA 0 adds the CTAs' partials in rank order out of distributed shared memory.  Returns the total in thread 0 of CTA 0
// (`bad` is OR-ed across the cluster the same way).  One launch, no scratch buffer, no atomics -- and eight SMs instead of
// one for the 3 x 32768 loads + logs of the loss (19.5 us -> a few us at B = 32768).
constexpr int kLossCtas = 8;

__device__ __forceinline__ double cluster_sum(double v, int& bad) {
  __shared__ double wpart[32];
  __shared__ double cta_part;
  __shared__ int cta_bad;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) wpart[threadIdx.x >> 5] = v;
  bad = __syncthreads_or(bad);
  if (threadIdx.x < 32) {
    double t = wpart[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) {
      cta_part = t;
      cta_bad = bad;
    }
  }
  cluster_sync_all();  // every CTA's partial is in its shared memory (release / acquire at cluster scope)
  double total = 0.0;
  if (cluster_ctarank() == 0 && threadIdx.x == 0) {
    int b = 0;
    for (uint32_t r = 0; r < kLossCtas; ++r) {
      uint32_t pa, ba;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(pa) : "r"(smem_u32(&cta_part)), "r"(r));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ba) : "r"(smem_u32(&cta_bad)), "r"(r));
      double pv;
      int bv;
      asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(pv) : "r"(pa) : "memory");
      asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(bv) : "r"(ba) : "memory");
      total += pv;
      b |= bv;
    }
    bad = b;
  }
  cluster_sync_all();  // nobody leaves (and frees its shared memory) before CTA 0 has read it
  return total;
}

// Each thread sums its strided share of f(i) in fp32 over four independent chains (a handful of terms of magnitude ~10:
// ~1e-7 relative); everything above the thread level is summed in double.
__global__ void __launch_bounds__(1024)
infonce_loss_kernel(const float* __restrict__ rowsum, const float* __restrict__ colsum, const float* __restrict__ diag,
                    int n, const float* __restrict__ scale, float inv_two_b, float* __restrict__ loss_out) {
  pdl_entry();
  const float s = *scale;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int bad = 0;
  int k = 0;
  for (int i = blockIdx.x * 1024 + threadIdx.x; i < n; i += kLossCtas * 1024, ++k) {
    const float rs = rowsum[i], cs = colsum[i];
    bad |= !(rs > 0.f && rs < INFINITY) || !(cs > 0.f && cs < INFINITY);
    acc[k & 3] += (logf(rs) - diag[i]) + (logf(cs) - diag[i]);
  }
  const double tot = cluster_sum((double)(acc[0] + acc[1]) + (double)(acc[2] + acc[3]), bad);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    loss_out[0] = bad ? __int_as_float(0x7fc00000) : (float)((tot + 2.0 * (double)n * (double)s) * (double)inv_two_b);
}

// launch of a loss kernel: one cluster of kLossCtas CTAs (+ the PDL attribute)
template <typename... KArgs, typename... Args>
static cudaError_t launch_loss_cluster(void (*kern)(KArgs...), cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kLossCtas);
  cfg.blockDim = dim3(1024);
  cfg.stream = st;
  PdlAttr at;
  at.cluster(kLossCtas);
  cfg.attrs = at.a;
  cfg.numAttrs = at.n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#define MMG_LAUNCH_LOSS(what, kern, st, ...)                        \
  do {                                                              \
    cudaError_t e__ = launch_loss_cluster(kern, st, __VA_ARGS__);   \
    if (e__ != cudaSuccess) return check_cuda(e__, what);           \
    count_launch();                                                 \
  } while (0)

int simt_infonce_loss(const float* rowsum, const float* colsum, const float* diag, int n, const float* scale,
                      float inv_two_b, float* loss_out, cudaStream_t st) {
  MMG_LAUNCH_LOSS("infonce_loss_kernel", infonce_loss_kernel, st, rowsum, colsum, diag, n, scale, inv_two_b, loss_out);
  return 0;
}

// Row-sharded loss in two parts, so that ONE cross-rank sum serves the column sums and the loss:
//   part_out[0] = sum_{local r} ( log rowsum[r] - 2*diag[r] )                 (before the exchange; rides in the same
//                                                                              all-reduce as the partial column sums)
//   loss        = inv_two_b * ( sum_ranks part + sum_{all c} log colsum[c] + 2*cols*s )   (after it, on every rank)
// which equals inv_two_b * sum_i (log rowsum_i + log colsum_i + 2s - 2 diag_i) over the global batch (the matching-pair
// logit of column c is the one of row c).  A vanished / overflowed sum poisons the part (NaN), and with it the loss.
__global__ void __launch_bounds__(1024)
infonce_row_part_kernel(const float* __restrict__ rowsum, const float* __restrict__ diag, int rows,
                        float* __restrict__ part_out) {
  pdl_entry();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int bad = 0;
  int k = 0;
  for (int i = blockIdx.x * 1024 + threadIdx.x; i < rows; i += kLossCtas * 1024, ++k) {
    const float rs = rowsum[i];
    bad |= !(rs > 0.f && rs < INFINITY);
    acc[k & 3] += logf(rs) - 2.0f * diag[i];
  }
  const double tot = cluster_sum((double)(acc[0] + acc[1]) + (double)(acc[2] + acc[3]), bad);
  if (blockIdx.x == 0 && threadIdx.x == 0) part_out[0] = bad ? __int_as_float(0x7fc00000) : (float)tot;
}

__global__ void __launch_bounds__(1024)
infonce_loss_cols_kernel(const float* __restrict__ colsum, int cols, const float* __restrict__ scale,
                         const float* __restrict__ row_part, float inv_two_b, float* __restrict__ loss_out) {
  pdl_entry();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int bad = 0;
  int k = 0;
  for (int i = blockIdx.x * 1024 + threadIdx.x; i < cols; i += kLossCtas * 1024, ++k) {
    const float cs = colsum[i];
    bad |= !(cs > 0.f && cs < INFINITY);
    acc[k & 3] += logf(cs);
  }
  const double tot = cluster_sum((double)(acc[0] + acc[1]) + (double)(acc[2] + acc[3]), bad);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double v = ((double)row_part[0] + tot + 2.0 * (double)cols * (double)(*scale)) * (double)inv_two_b;
    loss_out[0] = bad ? __int_as_float(0x7fc00000) : (float)v;  // a NaN part propagates by itself
  }
}

int simt_infonce_row_part(const float* rowsum, const float* diag, int rows, float* part_out, cudaStream_t st) {
  MMG_LAUNCH_LOSS("infonce_row_part_kernel", infonce_row_part_kernel, st, rowsum, diag, rows, part_out);
  return 0;
}

int simt_infonce_loss_cols(const float* colsum, int cols, const float* scale, const float* row_part, float inv_two_b,
                           float* loss_out, cudaStream_t st) {
  MMG_LAUNCH_LOSS("infonce_loss_cols_kernel", infonce_loss_cols_kernel, st, colsum, cols, scale, row_part, inv_two_b,
                  loss_out);
  return 0;
}

// fp16 coefficient scaling (MMG_PREC_F16).  The gradient coefficients g = coef * E * (1/rowsum + 1/colsum), coef =
// s*gl/(2B), are ~1/B^2 -- far below fp16's range -- but g / coef <= 2 (E <= rowsum and E <= colsum), so the tensor-core
// path stores g' = 2^14 * E * (1/rowsum + 1/colsum) in [0, 2^15] and the gradient epilogues multiply coef * 2^-14 back in
// (scal[3]): full 11-bit precision down to 2^-29 of the largest coefficient, no data-dependent scale to compute.
constexpr float kF16CoefScale = 16384.0f;

__global__ void infonce_bwd_prep_kernel(const float* __restrict__ rowsum, int rows, const float* __restrict__ colsum,
                                        int cols, const float* __restrict__ scale, const float* __restrict__ grad_loss,
                                        float inv_two_b, int diag_in_fp32, int f16_scaled, float* __restrict__ rinv,
                                        float* __restrict__ cinv, float* __restrict__ scal) {
  pdl_entry();
  const float coef = (*scale) * (*grad_loss) * inv_two_b;
  const float num = f16_scaled ? kF16CoefScale : coef;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) rinv[i] = num / rowsum[i];
  if (i < cols) cinv[i] = num / colsum[i];
  if (i == 0) {
    // [0] what the block kernels subtract on the diagonal (in the units g is stored in)
    scal[0] = diag_in_fp32 ? 0.f : (f16_scaled ? 2.0f * kF16CoefScale : 2.0f * coef);
    scal[1] = 2.0f * coef;                       // dcoef = s*gl/B
    scal[2] = diag_in_fp32 ? 1.f : 0.f;          // block kernels zero the diagonal element of g
    scal[3] = f16_scaled ? coef / kF16CoefScale : 1.0f;  // factor of the gradient epilogues (and of sum g*cos)
  }
}

int simt_infonce_bwd_prep(const float* rowsum, int rows, const float* colsum, int cols, const float* scale,
                          const float* grad_loss, float inv_two_b, int diag_in_fp32, int f16_scaled, float* rinv,
                          float* cinv, float* scal, cudaStream_t st) {
  const int n = rows > cols ? rows : cols;
  MMG_LAUNCH_PDL("infonce_bwd_prep_kernel", infonce_bwd_prep_kernel, (n + 255) / 256, 256, 0, st, rowsum, rows, colsum, cols,
                 scale, grad_loss, inv_two_b, diag_in_fp32, f16_scaled, rinv, cinv, scal);
  return 0;
}

// =====================================================================================================
// literal cross-entropy on materialised logits: F.cross_entropy(logits, labels), labels == NULL -> arange(n)
// (losses.py:39-43; AveragedMedicalCLIPLoss passes cluster labels, losses.py:207-212)
// =====================================================================================================
__global__ void __launch_bounds__(256)
ce_fwd_kernel(const float* __restrict__ logits, long long ld, int n, int m, const long long* __restrict__ labels,
              float coef, float* __restrict__ lse, float* __restrict__ loss_out) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  float term = 0.f;
  if (r < n) {
    const float* lr = logits + (long long)r * ld;
    float mx = -INFINITY;
    for (int c = lane; c < m; c += 32) mx = fmaxf(mx, lr[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < m; c += 32) s += expf(lr[c] - mx);
    s = warp_sum(s);
    const float l = mx + logf(s);
    if (lane == 0) {
      lse[r] = l;
      const long long lab = labels ? labels[r] : r;
      // an out-of-range class index (torch raises a device-side assert) poisons the loss instead of reading out of bounds
      term = (lab >= 0 && lab < m) ? l - lr[lab] : __int_as_float(0x7fc00000);
    }
  }
  __shared__ float wsum[8];
  if (lane == 0) wsum[threadIdx.x >> 5] = term;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += wsum[i];
    atomicAdd(loss_out, coef * t);
  }
}

__global__ void __launch_bounds__(256)
ce_bwd_kernel(const float* __restrict__ logits, long long ld, int n, int m, const long long* __restrict__ labels,
              const float* __restrict__ lse, const float* __restrict__ grad_loss, float coef,
              float* __restrict__ dlogits, long long ldd) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  const float k = coef * (*grad_loss);
  const float l = lse[r];
  const long long lab = labels ? labels[r] : r;
  for (int c = lane; c < m; c += 32) {
    float p = expf(logits[(long long)r * ld + c] - l);
    if (c == lab) p -= 1.0f;
    dlogits[(long long)r * ldd + c] = k * p;
  }
}

int simt_ce_fwd(const float* logits, long long ld, int n, int m, const long long* labels, float coef, float* lse,
                float* loss_out, cudaStream_t st) {
  if (n <= 0) return 0;
  ce_fwd_kernel<<<(n + 7) / 8, 256, 0, st>>>(logits, ld, n, m, labels, coef, lse, loss_out);
  MMG_LAUNCH_CHECK("ce_fwd_kernel");
  return 0;
}
int simt_ce_bwd(const float* logits, long long ld, int n, int m, const long long* labels, const float* lse,
                const float* grad_loss, float coef, float* dlogits, long long ldd, cudaStream_t st) {
  if (n <= 0) return 0;
  ce_bwd_kernel<<<(n + 7) / 8, 256, 0, st>>>(logits, ld, n, m, labels, lse, grad_loss, coef, dlogits, ldd);
  MMG_LAUNCH_CHECK("ce_bwd_kernel");
  return 0;
}

// =====================================================================================================
// InfoNCE backward, matching-pair (diagonal) element in fp32.
// g[r, r'] = E*(rinv + cinv) - dcoef carries almost all of the gradient's magnitude and cancels heavily against the
// off-diagonal sum; multiplying it by a bf16-rounded embedding would put the full 2^-9 operand rounding error straight
// into dA/dB.  The bf16 path therefore zeroes that element in the tensor-core contraction (scal[2] != 0) and applies
// it here from the fp32 embeddings:
//   g = exp(diag[r] - s) * (rinv[r] + cinvm[r]) - dcoef
//   dA[r,:] += g * b32m[r,:]    dBm[r,:] += g * a32[r,:]    dlogscale += g * diag[r] / s
// where b32m / dBm / cinvm are the column-side rows paired with the local rows.  One warp per pair; block-level
// fixed-order sum, one atomic per block.
// =====================================================================================================
__global__ void __launch_bounds__(256)
infonce_bwd_diag_kernel(const float* __restrict__ a32, const float* __restrict__ b32, int rows, int D,
                        const float* __restrict__ diag, const float* __restrict__ scale,
                        const float* __restrict__ rinv, const float* __restrict__ cinvm,
                        const float* __restrict__ scal, float* __restrict__ dA, float* __restrict__ dB,
                        float* __restrict__ dlogscale_acc, int init) {
  pdl_entry();
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const float s = *scale;
  const float dcoef = scal[1];
  float contrib = 0.f;
  if (r < rows) {
    const float lg = diag[r];
    const float g = expf(lg - s) * (rinv[r] + cinvm[r]) * scal[3] - dcoef;  // scal[3]: units of rinv / cinv -> true
    const float* ar = a32 + (long long)r * D;
    const float* br = b32 + (long long)r * D;
    float* dar = dA + (long long)r * D;
    float* dbr = dB + (long long)r * D;
    if (init) {
      // first writer of the gradient rows: no read-modify-write (the contraction kernels then accumulate on top)
      if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(ar) | reinterpret_cast<uintptr_t>(br) |
                            reinterpret_cast<uintptr_t>(dar) | reinterpret_cast<uintptr_t>(dbr)) & 15) == 0) {
        for (int i = lane * 4; i < D; i += 128) {
          const float4 av = *reinterpret_cast<const float4*>(ar + i), bv = *reinterpret_cast<const float4*>(br + i);
          *reinterpret_cast<float4*>(dar + i) = make_float4(g * bv.x, g * bv.y, g * bv.z, g * bv.w);
          *reinterpret_cast<float4*>(dbr + i) = make_float4(g * av.x, g * av.y, g * av.z, g * av.w);
        }
      } else {
        for (int i = lane; i < D; i += 32) {
          dar[i] = g * br[i];
          dbr[i] = g * ar[i];
        }
      }
    } else {
      for (int i = lane; i < D; i += 32) {
        const float av = ar[i], bv = br[i];
        dar[i] = fmaf(g, bv, dar[i]);
        dbr[i] = fmaf(g, av, dbr[i]);
      }
    }
    contrib = g * (lg / s);
  }
  __shared__ float wsum[8];
  if (lane == 0) wsum[threadIdx.x >> 5] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += wsum[i];
    if (dlogscale_acc != nullptr) atomicAdd(dlogscale_acc, t);
  }
}

int simt_infonce_bwd_diag(const float* a32, const float* b32, int rows, int D, const float* diag, const float* scale,
                          const float* rinv, const float* cinvm, const float* scal, float* dA, float* dB,
                          float* dlogscale_acc, int init, cudaStream_t st) {
  if (rows <= 0) return 0;
  MMG_LAUNCH_PDL("infonce_bwd_diag_kernel", infonce_bwd_diag_kernel, (rows + 7) / 8, 256, 0, st, a32, b32, rows, D, diag, scale, rinv, cinvm, scal, dA, dB,
                                                          dlogscale_acc, init);
  return 0;
}

// prep + matching pair in one launch (the bf16 path always runs both): every block first fills its grid-stride share of
// rinv[rows] / cinv[cols] / scal (what infonce_bwd_prep_kernel writes, for the contraction kernel that follows) and then
// handles its eight pairs, forming their rinv / cinv on the fly from rowsum / colsum -- so the second half does not depend
// on the first and needs no grid-wide ordering.  init = first-writer form only.
__global__ void __launch_bounds__(256)
infonce_bwd_prep_diag_kernel(const float* __restrict__ rowsum, int rows, const float* __restrict__ colsum, int cols,
                             int diag_offset, const float* __restrict__ scale, const float* __restrict__ grad_loss,
                             float inv_two_b, int f16_scaled, float* __restrict__ rinv, float* __restrict__ cinv,
                             float* __restrict__ scal,
                             const float* __restrict__ a32, const float* __restrict__ b32, int D,
                             const float* __restrict__ diag, float* __restrict__ dA, float* __restrict__ dB,
                             float* __restrict__ dlogscale_acc) {
  pdl_entry();
  const float s = *scale;
  const float coef = s * (*grad_loss) * inv_two_b;
  const int n = rows > cols ? rows : cols;
  const float num = f16_scaled ? kF16CoefScale : coef;  // see kF16CoefScale
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    if (i < rows) rinv[i] = num / rowsum[i];
    if (i < cols) cinv[i] = num / colsum[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scal[0] = 0.f;           // the contraction subtracts nothing on the diagonal ...
    scal[1] = 2.0f * coef;   // dcoef = s*gl/B
    scal[2] = 1.f;           // ... it zeroes the matching-pair element of g: applied here in fp32
    scal[3] = f16_scaled ? coef / kF16CoefScale : 1.0f;  // factor of the gradient epilogues (and of sum g*cos)
  }
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  float contrib = 0.f;
  if (r < rows) {
    const float lg = diag[r];
    const float g = expf(lg - s) * (coef / rowsum[r] + coef / colsum[diag_offset + r]) - 2.0f * coef;
    const float* ar = a32 + (long long)r * D;
    const float* br = b32 + (long long)r * D;
    float* dar = dA + (long long)r * D;
    float* dbr = dB + (long long)r * D;
    if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(ar) | reinterpret_cast<uintptr_t>(br) |
                          reinterpret_cast<uintptr_t>(dar) | reinterpret_cast<uintptr_t>(dbr)) & 15) == 0) {
      for (int i = lane * 4; i < D; i += 128) {
        const float4 av = *reinterpret_cast<const float4*>(ar + i), bv = *reinterpret_cast<const float4*>(br + i);
        *reinterpret_cast<float4*>(dar + i) = make_float4(g * bv.x, g * bv.y, g * bv.z, g * bv.w);
        *reinterpret_cast<float4*>(dbr + i) = make_float4(g * av.x, g * av.y, g * av.z, g * av.w);
      }
    } else {
      for (int i = lane; i < D; i += 32) {
        dar[i] = g * br[i];
        dbr[i] = g * ar[i];
      }
    }
    contrib = g * (lg / s);
  }
  if (dlogscale_acc != nullptr) {  // uniform across the grid
    __shared__ float wsum[8];
    if (lane == 0) wsum[threadIdx.x >> 5] = contrib;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += wsum[i];
      atomicAdd(dlogscale_acc, t);
    }
  }
}

int simt_infonce_bwd_prep_diag(const float* rowsum, int rows, const float* colsum, int cols, int diag_offset,
                               const float* scale, const float* grad_loss, float inv_two_b, int f16_scaled, float* rinv,
                               float* cinv, float* scal, const float* a32, const float* b32, int D, const float* diag,
                               float* dA, float* dB, float* dlogscale_acc, cudaStream_t st) {
  MMG_LAUNCH_PDL("infonce_bwd_prep_diag_kernel", infonce_bwd_prep_diag_kernel, (rows + 7) / 8, 256, 0, st, rowsum, rows, colsum, cols, diag_offset, scale, grad_loss,
                                                               inv_two_b, f16_scaled, rinv, cinv, scal, a32, b32, D, diag, dA, dB,
                                                               dlogscale_acc);
  return 0;
}

// out[0] = sum_i x[i]*y[i]  (single block, fixed order => deterministic); used for d logit_scale of materialised logits
__global__ void __launch_bounds__(1024)
dot_sum_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n, float* __restrict__ out) {
  __shared__ double part[1024];
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) acc += (double)x[i] * (double)y[i];
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)part[0];
}

int simt_dot_sum(const float* x, const float* y, long long n, float* out, cudaStream_t st) {
  dot_sum_kernel<<<1, 1024, 0, st>>>(x, y, n, out);
  MMG_LAUNCH_CHECK("dot_sum_kernel");
  return 0;
}

// =====================================================================================================
// zero-shot prompt scoring: logits = (s*img) . txt^T in fp32, softmax, argmax, top-k
// Block = 64 image rows; the 64x64 logit tile is accumulated with the same 4x4 FFMA micro-tile as sgemm, parked in
// shared memory, then one warp per 8 rows finishes softmax / argmax / top-k with shuffles.
// =====================================================================================================
__device__ __forceinline__ void warp_argmax(float& v, int& idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) {
      v = ov;
      idx = oi;
    }
  }
}

__global__ void __launch_bounds__(256)
zeroshot_kernel(const float* __restrict__ img, const float* __restrict__ txt, int N, int C, int D,
                const float* __restrict__ scale, float* __restrict__ logits_out, float* __restrict__ probs_out,
                long long* __restrict__ argmax_out, int k, long long* __restrict__ topk_idx,
                float* __restrict__ topk_val) {
  __shared__ float As[32][64 + 4];
  __shared__ float Bs[32][64 + 4];
  __shared__ float Ls[64][64 + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * 64;
  const float s = *scale;
  const bool vec = ((D & 3) == 0) && ((reinterpret_cast<uintptr_t>(img) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(txt) & 15) == 0);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < D; k0 += 32) {
    // 64 rows x 32 k = 512 float4 per operand, 2 per thread
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int e = tid + i * 256;
      const int r = e >> 3, kq = (e & 7) * 4;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
      const int gm = m0 + r, gk = k0 + kq;
      if (vec) {
        if (gm < N && gk < D) a = *reinterpret_cast<const float4*>(img + (long long)gm * D + gk);
        if (r < C && gk < D) b = *reinterpret_cast<const float4*>(txt + (long long)r * D + gk);
      } else {
        float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < 4; ++j) {
          if (gm < N && gk + j < D) av[j] = img[(long long)gm * D + gk + j];
          if (r < C && gk + j < D) bv[j] = txt[(long long)r * D + gk + j];
        }
        a = make_float4(av[0], av[1], av[2], av[3]);
        b = make_float4(bv[0], bv[1], bv[2], bv[3]);
      }
      // reference order of operations: (logit_scale * image_embeddings) @ text_embeddings.t()
      As[kq + 0][r] = s * a.x; As[kq + 1][r] = s * a.y; As[kq + 2][r] = s * a.z; As[kq + 3][r] = s * a.w;
      Bs[kq + 0][r] = b.x; Bs[kq + 1][r] = b.y; Bs[kq + 2][r] = b.z; Bs[kq + 3][r] = b.w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Ls[ty * 4 + i][tx * 4 + j] = acc[i][j];
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  for (int rr = 0; rr < 8; ++rr) {
    const int r = warp * 8 + rr;
    const long long gm = (long long)m0 + r;
    if (gm >= N) break;  // warp-uniform
    const float v0 = (lane < C) ? Ls[r][lane] : -INFINITY;
    const float v1 = (lane + 32 < C) ? Ls[r][lane + 32] : -INFINITY;
    if (logits_out != nullptr) {
      if (lane < C) logits_out[gm * C + lane] = v0;
      if (lane + 32 < C) logits_out[gm * C + lane + 32] = v1;
    }
    const float mx = warp_max(fmaxf(v0, v1));
    const float e0 = (lane < C) ? expf(v0 - mx) : 0.f;
    const float e1 = (lane + 32 < C) ? expf(v1 - mx) : 0.f;
    const float den = warp_sum(e0 + e1);
    const float p0 = e0 / den, p1 = e1 / den;
    if (probs_out != nullptr) {
      if (lane < C) probs_out[gm * C + lane] = p0;
      if (lane + 32 < C) probs_out[gm * C + lane + 32] = p1;
    }
    if (argmax_out != nullptr) {
      // the reference takes argmax of the PROBABILITIES (mmgclip_model.py:204,209); ties -> lowest index
      float bv = (lane < C) ? p0 : -INFINITY;
      int bi = lane;
      if (lane + 32 < C && p1 > bv) { bv = p1; bi = lane + 32; }
      warp_argmax(bv, bi);
      if (lane == 0) argmax_out[gm] = bi;
    }
    if (k > 0 && topk_idx != nullptr) {
      float w0 = v0, w1 = v1;  // top-k is defined on the logits: value desc, index asc
      for (int j = 0; j < k; ++j) {
        float bv = w0;
        int bi = lane;
        if (w1 > bv) { bv = w1; bi = lane + 32; }
        warp_argmax(bv, bi);
        if (lane == 0) {
          topk_idx[gm * k + j] = (j < C) ? bi : -1;
          if (topk_val != nullptr) topk_val[gm * k + j] = bv;
        }
        if (bi == lane) w0 = -INFINITY;
        if (bi == lane + 32) w1 = -INFINITY;
      }
    }
  }
}

int simt_zeroshot(const float* img, const float* txt, int N, int C, int D, const float* scale, float* logits_out,
                  float* probs_out, long long* argmax_out, int k, long long* topk_idx, float* topk_val,
                  cudaStream_t st) {
  if (N <= 0) return 0;
  zeroshot_kernel<<<(N + 63) / 64, 256, 0, st>>>(img, txt, N, C, D, scale, logits_out, probs_out, argmax_out, k,
                                                 topk_idx, topk_val);
  MMG_LAUNCH_CHECK("zeroshot_kernel");
  return 0;
}

// Any number of prompts (the tiled kernels above and in zeroshot_tc.cu stop at 64, PromptClassifier has no such limit):
// one warp per image row.  The scaled row sits in shared memory; lanes stride over the prompts' elements (coalesced,
// the prompt matrix stays in L1/L2), shuffle-reduce each logit, park the row's logits in global memory (logits_out or a
// caller-provided scratch row block) and finish softmax / argmax-of-probabilities / top-k from there with the same tie
// rules.  A fallback for rare shapes, not a tuned kernel.
__global__ void __launch_bounds__(256)
zeroshot_wide_kernel(const float* __restrict__ img, const float* __restrict__ txt, int N, int C, int D,
                     const float* __restrict__ scale, float* __restrict__ logits, float* __restrict__ probs_out,
                     long long* __restrict__ argmax_out, int k, long long* __restrict__ topk_idx,
                     float* __restrict__ topk_val) {
  extern __shared__ float rowbuf[];  // 8 warps x D
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gm = (long long)blockIdx.x * 8 + warp;
  if (gm >= N) return;
  const float s = *scale;
  float* a = rowbuf + warp * D;
  for (int i = lane; i < D; i += 32) a[i] = s * img[gm * D + i];  // (logit_scale * image_embeddings) first
  __syncwarp();
  float* lr = logits + gm * C;
  for (int c = 0; c < C; ++c) {
    const float* b = txt + (long long)c * D;
    float acc = 0.f;
    for (int i = lane; i < D; i += 32) acc = fmaf(a[i], b[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) lr[c] = acc;
  }
  __syncwarp();
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lr[c]);
  mx = warp_max(mx);
  float den = 0.f;
  for (int c = lane; c < C; c += 32) den += expf(lr[c] - mx);
  den = warp_sum(den);
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float p = expf(lr[c] - mx) / den;
    if (probs_out != nullptr) probs_out[gm * C + c] = p;
    if (p > bv) { bv = p; bi = c; }  // ascending c per lane: the first occurrence wins
  }
  warp_argmax(bv, bi);
  if (lane == 0 && argmax_out != nullptr) argmax_out[gm] = bi;
  if (k > 0 && topk_idx != nullptr) {
    // k selection passes over the logits (value desc, index asc); already-taken indices are skipped by comparing with
    // the previous winner in (value, index) order
    float pv = INFINITY;
    int pi = -1;
    for (int j = 0; j < k; ++j) {
      float v = -INFINITY;
      int vi = 0x7fffffff;
      for (int c = lane; c < C; c += 32) {
        const float x = lr[c];
        const bool after_prev = (x < pv) || (x == pv && c > pi);
        if (after_prev && (x > v || (x == v && c < vi))) { v = x; vi = c; }
      }
      warp_argmax(v, vi);
      if (lane == 0) {
        topk_idx[gm * k + j] = vi;
        if (topk_val != nullptr) topk_val[gm * k + j] = v;
      }
      pv = v; pi = vi;
    }
  }
}

int simt_zeroshot_wide(const float* img, const float* txt, int N, int C, int D, const float* scale, float* logits,
                       float* probs_out, long long* argmax_out, int k, long long* topk_idx, float* topk_val,
                       cudaStream_t st) {
  if (N <= 0) return 0;
  const size_t smem = (size_t)8 * D * sizeof(float);
  if (smem > 48 * 1024) return set_error(-3, "zeroshot (wide): D = %d exceeds 1536", D);
  zeroshot_wide_kernel<<<(N + 7) / 8, 256, smem, st>>>(img, txt, N, C, D, scale, logits, probs_out, argmax_out, k,
                                                       topk_idx, topk_val);
  MMG_LAUNCH_CHECK("zeroshot_wide_kernel");
  return 0;
}

// =====================================================================================================
// 'eos' text pooling (mmgclip_model.py:108-111): the hidden state of the last attended token of every sequence,
//   idx[r] = attention_mask[r, :].sum() - 1  (a negative index wraps like Python's: an all-zero mask picks seq - 1),
//   out[r, :] = hidden[r, idx[r], :].
// One CTA per sequence: mask row reduction, then a float4 row copy (the only HBM traffic that matters: 2*H*4 B per row;
// the other seq-1 hidden states of the sequence are never touched).
// =====================================================================================================
__global__ void __launch_bounds__(256) eos_pool_kernel(const float* __restrict__ hidden,
                                                       const long long* __restrict__ mask, int seq, int H,
                                                       float* __restrict__ out, long long* __restrict__ idx_out) {
  __shared__ long long s_part[8];
  __shared__ long long s_idx;
  const int r = blockIdx.x;
  long long acc = 0;
  for (int t = threadIdx.x; t < seq; t += blockDim.x) acc += mask[(long long)r * seq + t];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_part[w];
    long long idx = tot - 1;
    if (idx < 0) idx += seq;
    if (idx < 0 || idx >= seq) idx = -1;  // out of range: torch would raise; rows are written as NaN and idx = -1
    s_idx = idx;
    if (idx_out != nullptr) idx_out[r] = idx;
  }
  __syncthreads();
  const long long idx = s_idx;
  float* o = out + (long long)r * H;
  if (idx < 0) {
    for (int c = threadIdx.x; c < H; c += blockDim.x) o[c] = __int_as_float(0x7fc00000);
    return;
  }
  const float* src = hidden + ((long long)r * seq + idx) * H;
  if ((H & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(o)) & 15) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* o4 = reinterpret_cast<float4*>(o);
    for (int c = threadIdx.x; c < H / 4; c += blockDim.x) o4[c] = __ldg(s4 + c);
  } else {
    for (int c = threadIdx.x; c < H; c += blockDim.x) o[c] = src[c];
  }
}
// backward: dhidden (already zero-filled) [r, idx[r], :] = dout[r, :]
__global__ void __launch_bounds__(256) eos_pool_bwd_kernel(const float* __restrict__ dout,
                                                           const long long* __restrict__ idx, int seq, int H,
                                                           float* __restrict__ dhidden) {
  const int r = blockIdx.x;
  const long long i = idx[r];
  if (i < 0 || i >= seq) return;
  float* dst = dhidden + ((long long)r * seq + i) * H;
  const float* src = dout + (long long)r * H;
  for (int c = threadIdx.x; c < H; c += blockDim.x) dst[c] = src[c];
}
int simt_eos_pool(const float* hidden, const long long* mask, int n, int seq, int H, float* out, long long* idx_out,
                  cudaStream_t st) {
  if (n <= 0) return 0;
  eos_pool_kernel<<<n, 256, 0, st>>>(hidden, mask, seq, H, out, idx_out);
  MMG_LAUNCH_CHECK("eos_pool_kernel");
  return 0;
}
int simt_eos_pool_bwd(const float* dout, const long long* idx, int n, int seq, int H, float* dhidden, cudaStream_t st) {
  if (n <= 0) return 0;
  cudaError_t e = cudaMemsetAsync(dhidden, 0, (size_t)n * seq * H * sizeof(float), st);
  if (e != cudaSuccess) return check_cuda(e, "eos_pool_bwd memset");
  eos_pool_bwd_kernel<<<n, 256, 0, st>>>(dout, idx, seq, H, dhidden);
  MMG_LAUNCH_CHECK("eos_pool_bwd_kernel");
  return 0;
}

// =====================================================================================================
// Multi-tensor AdamW (ClassifierExperiment.py:74,118: torch.optim.AdamW over model.parameters(), i.e. the head weights).
// Decoupled weight decay, bias-corrected moments, the operation order of torch's implementation:
//   p *= 1 - lr*wd;  m += (1-b1)*(g - m);  v = b2*v + (1-b2)*g*g;
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// All tensors of a launch travel in the kernel parameters (no descriptor table in HBM, so a CUDA graph node owns its
// own pointers -- GraphedStep re-points the gradients per recorded graph).  The step counter lives on the device:
// every CTA reads t = state[0] + 1; the last CTA to retire (ticket in state[1]) publishes it, so the update is one launch
// and replays correctly inside a CUDA graph.  `lr_dev` (optional) is read at run time for the same reason (schedulers).
// =====================================================================================================
struct AdamwTensors {
  float* p[kAdamwMaxTensors];
  const float* g[kAdamwMaxTensors];
  float* m[kAdamwMaxTensors];
  float* v[kAdamwMaxTensors];
  int chunk0[kAdamwMaxTensors + 1];  // prefix sum of per-tensor chunk counts
  long long numel[kAdamwMaxTensors];
  int n;
};
constexpr int kAdamwChunk = 256 * 4 * 2;  // elements per CTA iteration: 256 threads x 2 float4

__global__ void __launch_bounds__(256) adamw_kernel(const __grid_constant__ AdamwTensors T, float lr_host,
                                                    const float* __restrict__ lr_dev, float beta1, float beta2,
                                                    float eps, float weight_decay, long long* __restrict__ state,
                                                    int advance) {
  __shared__ float s_coef[3];
  __shared__ long long s_t;
  if (threadIdx.x == 0) {
    const long long t0 = state[0] + 1;
    const float lr0 = lr_dev != nullptr ? *lr_dev : lr_host;
    // bias corrections in double, as torch's Python-scalar path does
    const double bc1 = 1.0 - pow((double)beta1, (double)t0);
    const double bc2 = 1.0 - pow((double)beta2, (double)t0);
    s_coef[0] = (float)(1.0 - (double)lr0 * (double)weight_decay);
    s_coef[1] = (float)((double)lr0 / bc1);
    s_coef[2] = (float)(1.0 / sqrt(bc2));
    s_t = t0;
  }
  __syncthreads();
  const long long t = s_t;
  const float decay = s_coef[0], step_size = s_coef[1], rsqrt_bc2 = s_coef[2];
  const float w1 = 1.0f - beta1, w2 = 1.0f - beta2;
  const int total = T.chunk0[T.n];
  for (int c = blockIdx.x; c < total; c += gridDim.x) {
    int ti = 0;
    while (ti + 1 < T.n && c >= T.chunk0[ti + 1]) ++ti;
    const long long base = (long long)(c - T.chunk0[ti]) * kAdamwChunk;
    const long long n = T.numel[ti];
    float* __restrict__ p = T.p[ti];
    const float* __restrict__ g = T.g[ti];
    float* __restrict__ m = T.m[ti];
    float* __restrict__ v = T.v[ti];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long i0 = base + ((long long)u * 256 + threadIdx.x) * 4;
      if (i0 >= n) continue;
      const bool vec = (i0 + 4 <= n) && ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
      float pv[4], gv[4], mv[4], vv[4];
      if (vec) {
        *reinterpret_cast<float4*>(pv) = *reinterpret_cast<const float4*>(p + i0);
        *reinterpret_cast<float4*>(gv) = *reinterpret_cast<const float4*>(g + i0);
        *reinterpret_cast<float4*>(mv) = *reinterpret_cast<const float4*>(m + i0);
        *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(v + i0);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = i0 + j < n;
          pv[j] = ok ? p[i0 + j] : 0.f; gv[j] = ok ? g[i0 + j] : 0.f;
          mv[j] = ok ? m[i0 + j] : 0.f; vv[j] = ok ? v[i0 + j] : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gj = gv[j];
        float pj = pv[j] * decay;
        const float mj = mv[j] + w1 * (gj - mv[j]);
        const float vj = vv[j] * beta2 + w2 * gj * gj;
        const float denom = sqrtf(vj) * rsqrt_bc2 + eps;
        pj -= step_size * (mj / denom);
        pv[j] = pj; mv[j] = mj; vv[j] = vj;
      }
      if (vec) {
        *reinterpret_cast<float4*>(p + i0) = *reinterpret_cast<float4*>(pv);
        *reinterpret_cast<float4*>(m + i0) = *reinterpret_cast<float4*>(mv);
        *reinterpret_cast<float4*>(v + i0) = *reinterpret_cast<float4*>(vv);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (i0 + j < n) { p[i0 + j] = pv[j]; m[i0 + j] = mv[j]; v[i0 + j] = vv[j]; }
      }
    }
  }
  // last CTA out publishes the new step count (every CTA has read state[0] before it takes a ticket)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd(reinterpret_cast<unsigned long long*>(state + 1), 1ull);
    if (ticket == (unsigned long long)gridDim.x - 1) {
      state[1] = 0;
      if (advance) state[0] = t;
      __threadfence();
    }
  }
}

int simt_adamw(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
               const long long* numel, int n_tensors, float lr, const float* lr_dev, float beta1, float beta2,
               float eps, float weight_decay, long long* state, cudaStream_t st) {
  int done = 0;
  while (done < n_tensors) {
    AdamwTensors T;
    memset(&T, 0, sizeof(T));
    int k = 0;
    long long chunks = 0;
    while (done + k < n_tensors && k < kAdamwMaxTensors) {
      const long long n = numel[done + k];
      const long long c = (n + kAdamwChunk - 1) / kAdamwChunk;
      if (chunks + c > 0x7fffffffLL) break;
      T.p[k] = params[done + k]; T.g[k] = grads[done + k];
      T.m[k] = exp_avg[done + k]; T.v[k] = exp_avg_sq[done + k];
      T.numel[k] = n;
      T.chunk0[k] = (int)chunks;
      chunks += c;
      ++k;
    }
    if (k == 0) return set_error(-1, "adamw: tensor %d too large", done);
    T.chunk0[k] = (int)chunks;
    T.n = k;
    done += k;
    int grid = (int)(chunks < 148 * 8 ? (chunks < 1 ? 1 : chunks) : 148 * 8);
    adamw_kernel<<<grid, 256, 0, st>>>(T, lr, lr_dev, beta1, beta2, eps, weight_decay, state, done == n_tensors ? 1 : 0);
    MMG_LAUNCH_CHECK("adamw_kernel");
  }
  return 0;
}

}  // namespace mmg
