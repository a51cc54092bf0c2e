// mmgclip_b200 -- zero-shot prompt scoring on the tensor pipe, fp32-faithful (3xTF32), HBM-bound.
//
//   logits[n, c] = (s * img[n, :]) . txt[c, :]     softmax over c, argmax of the probabilities, top-k of the logits
//   (mmgclip_model.py:201-209; evaluator.py:182-188, 282-299, 354-368)
//
// At BASELINE config 4 (N = 2^20 rows, C = 64 prompts, D = 512) the fp32 FFMA kernel (simt_kernels.cu) needs 6.9e10
// FMAs = ~1 ms of the FMA pipe while the 2 GiB of embeddings stream from HBM in ~0.34 ms.  Here the contraction runs
// as tcgen05.mma kind::tf32 with both operands split into two TF32 terms,
//     a = a_hi + a_lo,  a_hi = a with the 13 low mantissa bits cleared,  a_lo = rn_tf32(a - a_hi)
//     b' = fl(s * b) = b_hi + b_lo likewise   (the scale is folded into the 64 prompts by the prep kernel; the FFMA
//                                              kernel keeps the reference's scale-the-image-first rounding order)
//     a.b' ~= a_hi.b_hi + a_lo.b_hi + a_hi.b_lo       (products of TF32 terms are exact in the fp32 accumulator;
//                                                      what is dropped is O(2^-21) relative per product)
// The tensor core adds into its fp32 accumulator with truncation, a bias that grows with the number of accumulations
// (measured 5e-6 of the largest logit after 192 of them), so the dominant hi.hi term alternates between two partial
// accumulators, the small terms have their own columns, and the epilogue adds the four in fp32 -- logits then agree
// with an fp32 FFMA evaluation to ~1e-6 (max-abs over max-abs).  Shared-memory bandwidth is the scarce resource (TMA
// writes, splitter read + write, tensor-core operand reads), so the prompt hi and lo tiles sit back to back and one
// N = 128 MMA computes a_hi.[b_hi ; b_lo] (a_hi is read once), followed by an N = 64 MMA for a_lo.b_hi.  The image tile is split by four
// CUDA-core warps in shared memory right after the TMA lands it, so HBM sees the embeddings exactly once:
//
//   warp 0      image producer  : img tile [128 rows x 32 fp32] per k-block, 8 x 16 KB in flight
//   warp 3      prompt producer : prompt hi/lo tiles [64 x 32] per k-block (3 x 16 KB ring, from L2)
//   warps 4-11  splitter     : two groups of four warps take alternate k-blocks: a_lo = rn_tf32(a - trunc(a)) into a
//                              4-slot ring (the raw tile serves as a_hi); fence.proxy.async; arrive
//   warp 1      MMA issuer   : 4 k-steps x (M128 N128 K8 + M128 N64 K8) tcgen05.mma kind::tf32 per k-block
//   warp 2      TMEM alloc   : 2 accumulator stages x 2 pairs x 128 columns (all 512)
//   warps 12-15 epilogue     : one row per thread (64 logits in registers): top-k by k branch-free selection passes,
//                              softmax, argmax of the probabilities (ties -> lowest index), coalesced index stores
// Persistent grid (one CTA per SM), the prompts' hi/lo tiles (2 x 64 x D fp32, L2 resident) come from a small prep kernel.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace mmg {

namespace {

constexpr int kZsRows = 128;      // image rows per tile (UMMA M)
constexpr int kZsC = 64;          // prompt slots (UMMA N); prompts beyond C are zero rows
constexpr int kZsBK = 32;         // fp32 elements per stage along D (= one 128-byte swizzle span)
constexpr int kZsUK = 8;          // K per tcgen05.mma for TF32
constexpr int kZsA = 6;           // image-tile ring (HBM latency is hidden here)
constexpr int kZsL = 4;           // lo-term ring (splitter -> MMA), two slots per splitter group
constexpr int kZsB = 3;           // prompt-tile ring (L2)
constexpr int kZsABytes = kZsRows * kZsBK * 4;   // 16 KB
constexpr int kZsBBytes = kZsC * kZsBK * 4;      // 8 KB (hi or lo)
constexpr int kZsSmem = (kZsA + kZsL) * kZsABytes + kZsB * 2 * kZsBBytes + 512;  // 208.5 KB
constexpr int kZsThreads = 32 * 16;
constexpr int kZsAccCols = 4 * kZsC;  // per accumulator stage: two [hi.hi partial | small terms] pairs

__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
// round-to-nearest (ties away) to 10 explicit mantissa bits; the inputs here are tiny remainders, never near overflow
__device__ __forceinline__ float tf32_round(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x00001000u) & 0xFFFFE000u);
}

// prompts -> [2][64][D]: hi and lo TF32 terms of s * txt, rows >= C zero
__global__ void zeroshot_prep_kernel(const float* __restrict__ txt, int C, int D, const float* __restrict__ scale,
                                     float* __restrict__ out) {
  const long long n = (long long)kZsC * D;
  const float s = *scale;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = static_cast<int>(i / D);
    const float b = r < C ? s * txt[i] : 0.f;
    const float hi = tf32_trunc(b);
    out[i] = hi;
    out[n + i] = tf32_round(b - hi);
  }
}

// kind::tf32 instruction descriptor: c_format F32 = 1 at [4,6); a_format TF32 = 2 at [7,10); b_format TF32 = 2 at
// [10,13); both operands K-major; N>>3 at [17,23); M>>4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kZsThreads, 1)
zeroshot_tc_kernel(const __grid_constant__ CUtensorMap mImg, const __grid_constant__ CUtensorMap mThi,
                   const __grid_constant__ CUtensorMap mTlo, int N, int C, int D, int flags,
                   float* __restrict__ logits_out, float* __restrict__ probs_out, long long* __restrict__ argmax_out,
                   int k, long long* __restrict__ topk_idx, float* __restrict__ topk_val) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("[mmgclip_b200] dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  // Three rings of different depth: the image tiles are what HBM latency has to be hidden for (8 x 16 KB in flight per
  // SM); their low-order terms only live between the splitter and the MMA (2 slots); the prompt tiles come from L2.
  uint8_t* smem = smem_raw;
  uint8_t* sA = smem;                                   // [kZsA] image tile, raw fp32 -> hi term in place
  uint8_t* sL = sA + kZsA * kZsABytes;                  // [kZsL] lo term of the image tile
  uint8_t* sB = sL + kZsL * kZsABytes;                  // [kZsB] prompt tile: hi (8 KB) + lo (8 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kZsB * 2 * kZsBBytes);
  uint64_t* fullA = bars;                      // [kZsA] TMA -> splitter
  uint64_t* emptyA = fullA + kZsA;             // [kZsA] MMA -> image producer
  uint64_t* fullB = emptyA + kZsA;             // [kZsB] TMA -> MMA
  uint64_t* emptyB = fullB + kZsB;             // [kZsB] MMA -> prompt producer
  uint64_t* split_bar = emptyB + kZsB;         // [kZsL] splitter -> MMA (hi written in place, lo written)
  uint64_t* emptyL = split_bar + kZsL;         // [kZsL] MMA -> splitter
  uint64_t* tfull_bar = emptyL + kZsL;         // [2] MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;        // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ntiles = (N + kZsRows - 1) / kZsRows;
  const int nkb = (D + kZsBK - 1) / kZsBK;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&mImg);
  if (warp == 3 && lane == 0) {
    tma_prefetch_desc(&mThi);
    tma_prefetch_desc(&mTlo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kZsA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < kZsB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    for (int i = 0; i < kZsL; ++i) { mbar_init(&split_bar[i], 4); mbar_init(&emptyL[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 2 * kZsAccCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== image producer: runs up to kZsA k-blocks ahead =====================
    int sa = 0;
    uint32_t pa = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&emptyA[sa], pa ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&fullA[sa], kZsABytes);
          tma_load_2d(&mImg, &fullA[sa], sA + sa * kZsABytes, kb * kZsBK, t * kZsRows, kEvictFirst);  // streamed once
        }
        __syncwarp();
        if (++sa == kZsA) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===================== prompt producer (L2-resident hi / lo tiles) =====================
    int sb = 0;
    uint32_t pb = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&emptyB[sb], pb ^ 1);
        if (elect_one_sync()) {
          uint8_t* dst = sB + sb * 2 * kZsBBytes;
          mbar_arrive_expect_tx(&fullB[sb], 2 * kZsBBytes);
          tma_load_2d(&mThi, &fullB[sb], dst, kb * kZsBK, 0, kEvictLast);
          tma_load_2d(&mTlo, &fullB[sb], dst + kZsBBytes, kb * kZsBK, 0, kEvictLast);
        }
        __syncwarp();
        if (++sb == kZsB) { sb = 0; pb ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc128 = make_idesc_tf32(kZsRows, 2 * kZsC);
    const uint32_t idesc64 = make_idesc_tf32(kZsRows, kZsC);
    int sa = 0, sb = 0, sl = 0;
    uint32_t pb = 0, pl = 0;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int acc_stage = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc_stage], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + acc_stage * kZsAccCols;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&split_bar[sl], pl);   // image tile: hi in place (slot sa), lo in slot sl
        mbar_wait(&fullB[sb], pb);
        tcgen05_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_hi = smem_u32(sA + sa * kZsABytes);
          const uint32_t a_lo = smem_u32(sL + sl * kZsABytes);
          const uint32_t b_hi = smem_u32(sB + sb * 2 * kZsBBytes);
#pragma unroll
          for (int ks = 0; ks < kZsBK / kZsUK; ++ks) {
            const uint32_t off = ks * kZsUK * 4;
            const uint64_t dah = make_smem_desc_sw128(a_hi + off, 0, 1024);
            const uint64_t dal = make_smem_desc_sw128(a_lo + off, 0, 1024);
            const uint64_t dbh = make_smem_desc_sw128(b_hi + off, 0, 1024);  // 128 rows: hi prompts, then lo prompts
            // a_hi x [b_hi ; b_lo] (N = 128) -> pair kb & 1 = [hi.hi partial | hi.lo]; a_lo x b_hi (N = 64) -> the small-term
            // columns of pair 0 (after the N = 128 MMA of kb = 0 has initialised them)
            umma_tf32_ss(tmem_d + (kb & 1) * 2 * kZsC, dah, dbh, idesc128, (kb >= 2 || ks > 0) ? 1u : 0u);
            umma_tf32_ss(tmem_d + kZsC, dal, dbh, idesc64, 1u);
          }
          umma_commit(&emptyA[sa]);
          umma_commit(&emptyL[sl]);
          umma_commit(&emptyB[sb]);
          if (kb == nkb - 1) umma_commit(&tfull_bar[acc_stage]);
        }
        __syncwarp();
        if (++sa == kZsA) sa = 0;
        if (++sb == kZsB) { sb = 0; pb ^= 1; }
        if (++sl == kZsL) { sl = 0; pl ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== splitter: a -> (hi in place, lo in the lo ring).  Two groups of four warps take alternate
    // k-blocks, so two image tiles are being split at any time =====================
    const int grp = (warp - 4) >> 2;
    const int tid = threadIdx.x - 128 - grp * 128;
    int sa = 0, sl = 0, seq = 0;
    uint32_t pa = 0, pl = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      for (int kb = 0; kb < nkb; ++kb, ++seq) {
        if ((seq & 1) == grp) {
          mbar_wait(&fullA[sa], pa);
          mbar_wait(&emptyL[sl], pl ^ 1);
          float4* a = reinterpret_cast<float4*>(sA + sa * kZsABytes);
          float4* lo = reinterpret_cast<float4*>(sL + sl * kZsABytes);
#pragma unroll
          for (int i = 0; i < kZsABytes / 16 / 128; ++i) {
            const int idx = tid + i * 128;
            const float4 v = a[idx];
            // The tensor core reads TF32 operands as fp32 words with the 13 low mantissa bits ignored (checked on B200:
            // writing the truncated hi term back changes no result bit), so the raw tile IS the hi operand.
            const float4 h = make_float4(tf32_trunc(v.x), tf32_trunc(v.y), tf32_trunc(v.z), tf32_trunc(v.w));
            if (flags & 1) a[idx] = h;  // measurement hook: explicit write-back
            lo[idx] = make_float4(tf32_round(v.x - h.x), tf32_round(v.y - h.y), tf32_round(v.z - h.z),
                                  tf32_round(v.w - h.w));
          }
          fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&split_bar[sl]);
        }
        if (++sa == kZsA) { sa = 0; pa ^= 1; }
        if (++sl == kZsL) { sl = 0; pl ^= 1; }
      }
    }
  } else if (warp >= 12) {
    // ===================== epilogue (warps 12-15): one row per thread =====================
    const int q = warp & 3;
    const int nmain = nkb < 2 ? nkb : 2;  // accumulator pairs that were written
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int acc_stage = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc_stage], acc_phase);
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_stage * kZsAccCols;
      float l[64];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // (small0 [+ small1]) [+ main1] + main0, fp32 round-to-nearest adds; pair p = columns [128 p, 128 p + 128)
        float v[32], w[32];
        tmem_ld_32x32b_x32(tacc + kZsC + h * 32, v);
        tmem_ld_wait();
        if (nmain > 1) {
          tmem_ld_32x32b_x32(tacc + 3 * kZsC + h * 32, w);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += w[j];
          tmem_ld_32x32b_x32(tacc + 2 * kZsC + h * 32, w);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += w[j];
        }
        tmem_ld_32x32b_x32(tacc + h * 32, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) l[h * 32 + j] = v[j] + w[j];
      }
      // the accumulator stage is free as soon as it is in registers
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc_stage]);

      const long long gm = (long long)t * kZsRows + q * 32 + lane;
      if (gm < N) {  // (a block, not `continue`: the warp must be converged again for the next tile's tcgen05.ld)
      if (logits_out != nullptr) {
#pragma unroll
        for (int j = 0; j < 64; ++j)
          if (j < C) logits_out[gm * C + j] = l[j];
      }
#pragma unroll
      for (int j = 0; j < 64; ++j)
        if (j >= C) l[j] = -INFINITY;  // empty prompt slots never win
      if (k > 0 && topk_idx != nullptr) {
        // top-k of the logits, value descending / index ascending: k selection passes over the registers, each taking
        // the largest value not yet taken (strict > while scanning upwards => lowest index among equals).  Branch-free:
        // every lane does the same work whatever its row holds.
        uint32_t taken_lo = 0u, taken_hi = 0u;
        for (int i = 0; i < k; ++i) {
          float b0 = -INFINITY, b1 = -INFINITY;   // two interleaved scans (even / odd j) for instruction-level parallelism
          int i0 = -1, i1 = -1;
#pragma unroll
          for (int j = 0; j < 64; j += 2) {
            const uint32_t m0 = (j < 32 ? taken_lo : taken_hi) & (1u << (j & 31));
            const uint32_t m1 = ((j + 1) < 32 ? taken_lo : taken_hi) & (1u << ((j + 1) & 31));
            const bool c0 = (m0 == 0u) && (l[j] > b0);
            const bool c1 = (m1 == 0u) && (l[j + 1] > b1);
            b0 = c0 ? l[j] : b0;
            i0 = c0 ? j : i0;
            b1 = c1 ? l[j + 1] : b1;
            i1 = c1 ? (j + 1) : i1;
          }
          // merge: larger value wins, equal values -> lower index; -1 = nothing left (fewer than k finite logits)
          const bool pick1 = (i0 < 0) || (i1 >= 0 && (b1 > b0 || (b1 == b0 && i1 < i0)));
          const float bv = pick1 ? b1 : b0;
          const int bi = pick1 ? i1 : i0;
          if (bi >= 0) {
            if (bi < 32) taken_lo |= 1u << bi;
            else taken_hi |= 1u << (bi - 32);
          }
          topk_idx[gm * k + i] = (i < C) ? bi : -1;
          if (topk_val != nullptr) topk_val[gm * k + i] = (i < C && bi >= 0) ? bv : -INFINITY;
        }
      }
      if (probs_out != nullptr || argmax_out != nullptr) {
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 64; ++j) mx = fmaxf(mx, l[j]);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int j = 0; j < 64; j += 2) {
          // exp(x) = 2^(x log2 e): one FFMA + MUFU.EX2 (2 ulp), -inf slots give exactly 0
          l[j] = ex2_approx((l[j] - mx) * 1.4426950408889634f);
          l[j + 1] = ex2_approx((l[j + 1] - mx) * 1.4426950408889634f);
          d0 += l[j];
          d1 += l[j + 1];
        }
        const float rden = 1.f / (d0 + d1);
        // the reference takes argmax of the PROBABILITIES (mmgclip_model.py:204,209); ties -> lowest index
        float bv = -INFINITY;
        int bi = 0;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const float pj = l[j] * rden;
          if (j < C) {
            if (probs_out != nullptr) probs_out[gm * C + j] = pj;
            if (pj > bv) { bv = pj; bi = j; }
          }
        }
        if (argmax_out != nullptr) argmax_out[gm] = bi;
      }
      }
      __syncwarp();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 2 * kZsAccCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn zs_encode_fn() {
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

// fp32 row-major [rows, D] (pitch D): box = [box_rows x 32 floats] (one 128-byte swizzle span), zero fill out of bounds
int make_f32_tmap(CUtensorMap* m, const float* ptr, long long D, long long rows, int box_rows) {
  EncodeTiledFn fn = zs_encode_fn();
  if (fn == nullptr) return set_error(-4, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)D * 4};
  cuuint32_t box[2] = {(cuuint32_t)kZsBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-4, "cuTensorMapEncodeTiled (fp32 operand) failed (CUresult %d)", (int)r);
  return 0;
}

}  // namespace

size_t tc_zeroshot_workspace_bytes(int C, int D) {
  if (C <= 0 || C > kZsC || D <= 0) return 0;
  return (size_t)2 * kZsC * D * sizeof(float) + 256;
}

bool tc_zeroshot_supported(const float* img, int N, int C, int D) {
#ifdef MMG_MEASURE
  if (const char* e = getenv("MMG_ZEROSHOT_TC"))
    if (e[0] == '0') return false;
#endif
  return N >= 1 && C >= 1 && C <= kZsC && D >= 4 && (D % 4) == 0 && (reinterpret_cast<uintptr_t>(img) & 15) == 0;
}

int tc_zeroshot(const float* img, const float* txt, int N, int C, int D, const float* scale, float* logits_out,
                float* probs_out, long long* argmax_out, int k, long long* topk_idx, float* topk_val, void* workspace,
                size_t workspace_bytes, cudaStream_t st) {
  if (N <= 0) return 0;
  if (workspace == nullptr || workspace_bytes < tc_zeroshot_workspace_bytes(C, D))
    return set_error(-1, "zero-shot scoring: workspace too small (%zu bytes)", workspace_bytes);
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
    return set_error(-2, "zero-shot scoring: workspace must be 256-byte aligned");
  float* thl = static_cast<float*>(workspace);
  const long long nt = (long long)kZsC * D;
  zeroshot_prep_kernel<<<static_cast<int>((nt + 255) / 256 < 256 ? (nt + 255) / 256 : 256), 256, 0, st>>>(txt, C, D, scale,
                                                                                                      thl);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "zeroshot_prep_kernel");
  count_launch();

  CUtensorMap mImg, mThi, mTlo;
  int rc;
  if ((rc = make_f32_tmap(&mImg, img, D, N, kZsRows)) != 0) return rc;
  if ((rc = make_f32_tmap(&mThi, thl, D, kZsC, kZsC)) != 0) return rc;
  if ((rc = make_f32_tmap(&mTlo, thl + nt, D, kZsC, kZsC)) != 0) return rc;

  static bool configured = false;
  if (!configured) {
    e = cudaFuncSetAttribute(zeroshot_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kZsSmem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(zeroshot_tc_kernel)");
    configured = true;
  }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int ntiles = (N + kZsRows - 1) / kZsRows;
  const int grid = ntiles < sms ? ntiles : sms;
  int flags = 0;
#ifdef MMG_MEASURE
  // measurement builds only: bit 0 = write the truncated hi term back explicitly
  if (const char* f = getenv("MMG_ZEROSHOT_FLAGS")) flags = atoi(f);
#endif
  zeroshot_tc_kernel<<<grid, kZsThreads, kZsSmem, st>>>(mImg, mThi, mTlo, N, C, D, flags, logits_out, probs_out,
                                                        argmax_out, k, topk_idx, topk_val);
  e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "zeroshot_tc_kernel");
  count_launch();
  return 0;
}

}  // namespace mmg
