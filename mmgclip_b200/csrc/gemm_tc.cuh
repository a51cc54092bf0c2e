// mmgclip_b200 -- the one tensor-core mainloop every dense contraction on the hot path runs through.
//
//   acc[128 kCG x BN] (fp32, TMEM) = A[128 kCG x K] * B[BN x K]^T      (bf16 operands, tcgen05.mma kind::f16)
//
// Persistent, warp-specialised CTA (one per SM; kCG = 2: CTA pairs driving one 256-row MMA):
//   warp 0      TMA producer   : cp.async.bulk.tensor -> 128B-swizzled shared-memory ring (6-7 stages of 32 KB)
//   warp 1      MMA issuer     : one thread issues tcgen05.mma; tcgen05.commit frees ring slots / publishes tiles
//   warp 2      TMEM allocator : 2 accumulator stages x BN columns (BN = 256 -> all 512 TMEM columns)
//   warps 4..11 epilogue       : tcgen05.ld the accumulator (32 lanes x 32 columns at a time) and apply a fused
//                                epilogue while the MMA warp already works on the next tile in the other stage
//
// Each operand may be K-major (row-major [rows x K]) or MN-major ([K x rows], rows contiguous), chosen at run time
// per problem, so no transposed copies of embeddings / gradients are ever made.  A launch can carry two independent
// problems (dI and dT of one logit block in the block-loop backward), a split-K factor, or up to three K segments
// (operand pair per segment, one accumulator: the split-precision heads).  bwd_fused.cuh reuses the same roles and
// epilogues for the one-launch backward.
//
// The epilogues are where the CLIP-specific fusion lives (see EpiLse / EpiGrad below): the logit tile is consumed
// straight out of TMEM and never written to HBM.
#pragma once

#include "common.cuh"

namespace mmg {

constexpr int kBM = 128;        // UMMA M (rows per tile)
constexpr int kBK = 64;         // K elements per pipeline stage (= one 128-byte swizzle span of bf16)
constexpr int kUmmaK = 16;      // K per tcgen05.mma for 16-bit operands
// Epilogue warps come in groups of four (one per TMEM lane quarter); an epilogue declares kWarps = 8 or 16, i.e. each
// warp owns 32 accumulator rows and BN/2 or BN/4 of the tile's columns.  Everything instantiated uses 8: the 16-warp
// form (register file re-balanced with setmaxnreg) measured slower, its staging boxes cost an operand-ring stage.

struct GemmProblem {
  int M, N, K;           // logical extents (ragged edges are zero-filled by TMA and masked in the epilogue)
  int tiles_m, tiles_n;  // ceil(M/(128*kCG)), ceil(N/BN)
  int k_splits;          // >= 1; each split handles a contiguous range of 64-wide K blocks
  int a_mn, b_mn;        // 0 = K-major operand, 1 = MN-major operand
  int a_f16, b_f16;      // operand element type: 0 = bf16, 1 = fp16 (normalised embeddings travel as fp16)
  // K segments (single-problem launches only): the contraction runs over k_segs concatenated K ranges of K elements
  // each, segment i reading A from tensor map (seg_a >> i) & 1 and B from (seg_b >> i) & 1 -- one launch and one
  // accumulator for  A_hi.B_hi + A_hi.B_lo + A_lo.B_hi  (split-precision heads) instead of three accumulate passes.
  int k_segs, seg_a, seg_b;
  __host__ __device__ int num_tiles() const { return tiles_m * tiles_n * k_splits; }
  __host__ __device__ int seg_kb() const { return (K + 63) / 64; }
  __host__ __device__ int total_kb() const { return seg_kb() * (k_segs > 1 ? k_segs : 1); }
};

// kCG = 1: one CTA owns a 128 x BN tile.  kCG = 2: a CTA pair (cta_group::2) owns a 256 x BN tile -- each CTA stages
// its own 128 A rows and HALF of the B rows, which halves the B-operand shared-memory / L2 traffic per FLOP.
template <int BN, int kCG, int kStagingBytes = 0, int kScratchBytes = 0>
struct GemmSmem {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBRows = BN / kCG;
  static constexpr int kBBytes = kBRows * kBK * 2;
  // operand ring: whatever is left of the 227 KB a CTA may use after the epilogue's output staging boxes (TMA-store
  // epilogues), the barriers and the alignment slack -- 7 stages of 32 KB without staging, 6 with 32 KB of boxes
#ifndef MMG_MAX_STAGES
#define MMG_MAX_STAGES 64  // A/B builds cap the ring depth (-DMMG_MAX_STAGES=n) to measure its effect
#endif
  // The dynamic shared-memory window is declared 1024-byte aligned (the kernels trap if it is not), so there is no
  // alignment slack: 6 x 32 KB + 32 KB of boxes + 2 KB of scratch + 512 B of barriers = 226.5 KB.
  static constexpr int kSmemMax = 227 * 1024;
  static constexpr int kBarBytes = 512;   // mbarriers + TMEM slot
  static constexpr int kFit = (kSmemMax - kBarBytes - kStagingBytes - kScratchBytes) / (kABytes + kBBytes);
  static constexpr int kStages = kFit < MMG_MAX_STAGES ? kFit : MMG_MAX_STAGES;
  static constexpr int kScratch = kScratchBytes;
  static constexpr int kTotal = kStages * (kABytes + kBBytes) + kStagingBytes + kScratchBytes + kBarBytes;
};

struct TileCoord {
  int prob, m_blk, n_blk, kb_begin, kb_end;
  int atomic;  // this tile is a K-slice of an output tile shared with other CTAs: accumulate with red.add
};

// Tail balancing ("stream-K lite").  With T output tiles on G CTA groups the last round is only (T mod G)/G full.
// Tiles of that last round are cut into `rem_split` K-slices each, spread over all groups and accumulated with atomics,
// so the tail costs ~(T mod G)/G of a round instead of a whole one.  Only used for accumulate-mode launches.
struct TailSplit {
  int full_tiles;  // tiles [0, full_tiles) are whole; the rest are split
  int rem_split;   // >= 1
  int prefetch_kb; // L2 prefetch distance of the TMA producer in 64-wide K blocks (0 = off)
};

__device__ __forceinline__ TileCoord decode_tile(int t, const GemmProblem& p0, const GemmProblem& p1) {
  TileCoord c;
  const int t0 = p0.tiles_m * p0.tiles_n * p0.k_splits;
  c.prob = (t >= t0) ? 1 : 0;
  const GemmProblem& p = c.prob ? p1 : p0;
  if (c.prob) t -= t0;
  const int mn = p.tiles_m * p.tiles_n;
  const int split = t / mn;
  const int r = t - split * mn;
  c.m_blk = r / p.tiles_n;
  c.n_blk = r - c.m_blk * p.tiles_n;
  const int nkb = p.total_kb();
  const int per = (nkb + p.k_splits - 1) / p.k_splits;
  c.kb_begin = split * per;
  c.kb_end = min(nkb, c.kb_begin + per);
  c.atomic = 0;
  return c;
}

__device__ __forceinline__ TileCoord decode_virtual(int t, const GemmProblem& p0, const GemmProblem& p1,
                                                    const TailSplit& ts) {
  if (ts.rem_split <= 1 || t < ts.full_tiles) return decode_tile(t, p0, p1);
  const int r = t - ts.full_tiles;
  const int base = ts.full_tiles + r / ts.rem_split;
  const int part = r - (r / ts.rem_split) * ts.rem_split;
  TileCoord c = decode_tile(base, p0, p1);
  const int nkb = c.kb_end - c.kb_begin;
  const int per = (nkb + ts.rem_split - 1) / ts.rem_split;
  c.kb_begin = c.kb_begin + part * per;
  c.kb_end = min(c.kb_end, c.kb_begin + per);
  c.atomic = 1;
  return c;
}

// ---------------------------------------------------------------------------------------------------------
// Epilogues.  Every epilogue warp owns 32 accumulator rows (TMEM lanes 32q..32q+31, q = warp % 4, one row per
// thread) and one column group (cg = (warp-4)/4) of BN / (kWarps/4) columns, visited 32 columns at a time.
// ---------------------------------------------------------------------------------------------------------

// C = alpha * acc (+ bias[col]) (ReLU) stored / accumulated / atomically added to fp32 row-major C.
template <int kW>
struct EpiStoreF32T {
  struct Params {
    float* C;
    long long ldc;
    const float* bias;  // per output column, may be null
    float alpha;
    const float* alpha_ptr;  // optional device scalar multiplied into alpha (e.g. the logit scale)
    int mode;           // 0: C = v   1: C += v (exclusive owner)   2: red.add (split-K)
    int relu;
    int use_tma;        // output goes through TMA store / reduce-add (needs 16-byte aligned C and pitch)
  };
  static constexpr int kWarps = kW;
  static constexpr int kScratchBytes = 0;
  // one [32 rows x 32 fp32] (4 KB, 128B-swizzled) staging box per epilogue warp
  template <int BN> __host__ __device__ static constexpr int staging_bytes() { return kWarps * 4096; }

  static __device__ __forceinline__ void finish(const Params&, float, int lane) {
    if (lane == 0) tma_store_wait_all();
  }

  // TMA path: every warp stages its own 32 x 32 block of the tile and ships it with a bulk-tensor store (C = v) or
  // reduce-add (C += v, performed at L2 -- also what makes split-K safe).  Warp-local: no CTA-level barriers; full
  // 128-byte lines, no LSU traffic to global memory, ragged edges clipped by the hardware.
  template <int BN>
  static __device__ __forceinline__ void run_tma(const Params& P, uint32_t tacc, int m0, int n0, int N, int half, int q,
                                                 int lane, const CUtensorMap* cmap, uint8_t* box) {
    constexpr int kCols = BN / (kWarps / 4);
    const float alpha = P.alpha_ptr ? P.alpha * __ldg(P.alpha_ptr) : P.alpha;
    uint8_t* rowp = box + lane * 128;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
#pragma unroll 1
    for (int ch = 0; ch < kCols / 32; ++ch) {
      const int cl = half * kCols + ch * 32;
      const int c0 = n0 + cl;
      if (c0 >= N) break;  // warp-uniform
      float v[32];
      tmem_ld_32x32b_x32(tacc + cl, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = v[j] * alpha;
        if (P.bias != nullptr && c0 + j < N) x += __ldg(P.bias + c0 + j);
        if (P.relu) x = fmaxf(x, 0.f);
        v[j] = x;
      }
      if (lane == 0) tma_store_wait_read();  // this warp's previous box has left shared memory
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(rowp + ((static_cast<uint32_t>(j) ^ sw) << 4)) =
            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (P.mode == 0) tma_store_2d(cmap, box, c0, m0 + q * 32);
        else tma_reduce_add_2d(cmap, box, c0, m0 + q * 32);
        tma_store_commit();
      }
    }
  }

  template <int BN>
  static __device__ __forceinline__ void run(const Params& P, uint32_t tacc, int m0, int n0, int M, int N, int half,
                                             int q, int lane, int ewarp, float* smem, const CUtensorMap* cmap,
                                             uint8_t* staging, int force_atomic, float& carry) {
    (void)smem; (void)carry;
    if (P.use_tma) {
      run_tma<BN>(P, tacc, m0, n0, N, half, q, lane, cmap, staging + ewarp * 4096);
      return;
    }
    constexpr int kCols = BN / (kWarps / 4);
    const int mode = force_atomic ? 2 : P.mode;
    const int row = m0 + q * 32 + lane;
    float* crow = P.C + static_cast<long long>(row) * P.ldc;
    const bool vec_ok = ((P.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0);
    const float alpha = P.alpha_ptr ? P.alpha * __ldg(P.alpha_ptr) : P.alpha;
#pragma unroll 1
    for (int ch = 0; ch < kCols / 32; ++ch) {
      const int cl = half * kCols + ch * 32;
      const int c0 = n0 + cl;
      if (c0 >= N) break;  // warp-uniform
      float v[32];
      tmem_ld_32x32b_x32(tacc + cl, v);
      tmem_ld_wait();
      if (row < M) {
        const bool vec = vec_ok && c0 + 32 <= N;
        // C = act(C_old + alpha*acc + bias): accumulate first, then bias / activation (multi-pass contractions)
        if (mode == 1) {
          if (vec) {
            const float4* src = reinterpret_cast<const float4*>(crow + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 old = src[j];
              v[4 * j] = fmaf(v[4 * j], alpha, old.x);
              v[4 * j + 1] = fmaf(v[4 * j + 1], alpha, old.y);
              v[4 * j + 2] = fmaf(v[4 * j + 2], alpha, old.z);
              v[4 * j + 3] = fmaf(v[4 * j + 3], alpha, old.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < N) v[j] = fmaf(v[j], alpha, crow[c0 + j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= alpha;
        }
        if (P.bias != nullptr || P.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = v[j];
            if (P.bias != nullptr && c0 + j < N) x += __ldg(P.bias + c0 + j);
            if (P.relu) x = fmaxf(x, 0.f);
            v[j] = x;
          }
        }
        if (vec && mode != 2) {
          float4* dst = reinterpret_cast<float4*>(crow + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (c0 + j < N) {
              if (mode == 2) atomicAdd(crow + c0 + j, v[j]);
              else crow[c0 + j] = v[j];
            }
          }
        }
      }
    }
  }
};

using EpiStoreF32 = EpiStoreF32T<8>;

// Forward InfoNCE epilogue: the accumulator holds cosines cos[r][c] of a logit tile.  With the fixed shift m = s
// (|cos| <= 1 => logits in [-s, s]) E = exp(s*cos - s) feeds the row sums AND the column sums with a single exp per
// logit, and partial sums from different tiles / GPUs simply add.  Nothing of the tile is written to memory except
//   rowsum[r] += sum_c E      colsum[c] += sum_r E      diag[r] = s * cos[r][r]
// The tile is read in the 16x256b fragment layout (4 rows x 8 columns of every 32x32 block per thread): row sums
// accumulate in registers for the whole tile, column sums are pre-reduced over the thread's 4 rows and finished with a
// 3-stage / 7-shuffle transposing butterfly.  Interior tiles take a mask-free path; the diagonal is looked at only in
// tiles that contain it.
// kStoreE: additionally keep E (bf16, row-major [M, N]) for a backward that transforms it into the gradient coefficients
// instead of recomputing the cosines (bwd_fused.cuh, stored-E mode).  Each warp packs its 32 x 64 block of E into a
// 128B-swizzled staging box (conflict-free 4-byte stores straight from the 16x256b fragments) and ships it with a TMA
// bulk-tensor store, like EpiGrad does for g.
template <int kW, bool kStoreE = false>
struct EpiLseT {
  struct Params {
    float* rowsum;            // [M]   (atomic accumulate; zero-initialised by the caller)
    float* colsum;            // [N]
    float* diag;              // [M]   logit of the matching pair
    const float* scale_ptr;   // device scalar s = exp(logit_scale)
    int diag_offset;          // global column index of local row 0 (rank offset in the sharded case)
  };
  static constexpr int kWarps = kW;
  // kStoreE: one [32 rows x 64 bf16] (4 KB, 128B-swizzled) staging box per epilogue warp
  template <int BN> __host__ __device__ static constexpr int staging_bytes() { return kStoreE ? kWarps * 4096 : 0; }
  static constexpr int kScratchBytes = 0;
  static __device__ __forceinline__ void finish(const Params&, float, int lane) {
    if (kStoreE && lane == 0) tma_store_wait_all();
  }

  template <int BN, bool kMasked, bool kDiag>
  static __device__ __forceinline__ void tile(const Params& P, uint32_t tacc, int m0, int n0, int M, int N, int half,
                                              int q, int lane, float s, float sl2, const CUtensorMap* cmap,
                                              uint8_t* box) {
    constexpr int kCols = BN / (kWarps / 4);
    const int lr = lane >> 2;          // row within an 8-row group
    const int lc = (lane & 3) * 2;     // first of this thread's two columns within an 8-column group
    const int rbase = m0 + q * 32 + lr;  // rows rbase + 8*i, i = 0..3
    float racc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int ch = 0; ch < kCols / 32; ++ch) {
      const int cl = half * kCols + ch * 32;
      const int c0 = n0 + cl;
      if (kMasked && c0 >= N) break;  // warp-uniform
      float va[16], vb[16];
      tmem_ld_16x256b_x4(tacc + cl, va);                  // lanes 32q + 0..15  -> rows i = 0, 1
      tmem_ld_16x256b_x4(tacc + (16u << 16) + cl, vb);    // lanes 32q + 16..31 -> rows i = 2, 3
      tmem_ld_wait();
      if (kStoreE && (ch & 1) == 0) {
        if (lane == 0) tma_store_wait_read();  // this warp's previous box has left shared memory
        __syncwarp();
      }
      float cp[8];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float evs[2][4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = c0 + 8 * g + lc + e;
          float x[4] = {va[4 * g + e], va[4 * g + 2 + e], vb[4 * g + e], vb[4 * g + 2 + e]};
          float csum = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (kDiag) {
              const int row = rbase + 8 * i;
              if (row + P.diag_offset == col && (!kMasked || row < M)) P.diag[row] = x[i] * s;
            }
            float ev = ex2_approx(fmaf(x[i], sl2, -sl2));
            if (kMasked) ev = (rbase + 8 * i < M && col < N) ? ev : 0.f;
            racc[i] += ev;
            csum += ev;
            evs[e][i] = ev;
          }
          cp[2 * g + e] = csum;
        }
        if (kStoreE) {
          // rows lr + 8i (row & 7 == lr), columns 8g + lc, +1 of this 32-column chunk: 16-byte unit (ch & 1) * 4 + g of
          // the box row, 4 bytes at (lane & 3) * 4 inside it -- 32 distinct banks per store instruction
          const uint32_t unit = static_cast<uint32_t>((ch & 1) * 4 + g) ^ static_cast<uint32_t>(lr);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint32_t*>(box + (lr + 8 * i) * 128 + (unit << 4) + (lane & 3) * 4) =
                pack_bf16x2(evs[0][i], evs[1][i]);
        }
      }
      if (kStoreE && ((ch & 1) == 1 || ch + 1 == kCols / 32 || (kMasked && c0 + 32 >= N))) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          // columns / rows beyond the matrix are clipped by the hardware
          tma_store_2d(cmap, box, n0 + half * kCols + (ch & ~1) * 32, m0 + q * 32);
          tma_store_commit();
        }
      }
      // transposing butterfly over the 8 lanes that share lane%4: afterwards this lane holds the full 32-row sum of
      // column 8*g' + lc + e' with g' = 2*bit4 + bit3, e' = bit2 of the lane index
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool up = lane & 16;
        const float keep = up ? cp[j + 4] : cp[j];
        const float send = up ? cp[j] : cp[j + 4];
        cp[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const bool up = lane & 8;
        const float keep = up ? cp[j + 2] : cp[j];
        const float send = up ? cp[j] : cp[j + 2];
        cp[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      {
        const bool up = lane & 4;
        const float keep = up ? cp[1] : cp[0];
        const float send = up ? cp[0] : cp[1];
        cp[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      const int mycol = c0 + 8 * (2 * ((lane >> 4) & 1) + ((lane >> 3) & 1)) + lc + ((lane >> 2) & 1);
      if (!kMasked || mycol < N) atomicAdd(P.colsum + mycol, cp[0]);
    }
    // row sums: 4 values per thread, 4 lanes (same lane/4) share the rows -> 3-shuffle transposing butterfly
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const bool up = lane & 2;
      const float keep = up ? racc[i + 2] : racc[i];
      const float send = up ? racc[i] : racc[i + 2];
      racc[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    {
      const bool up = lane & 1;
      const float keep = up ? racc[1] : racc[0];
      const float send = up ? racc[0] : racc[1];
      racc[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    const int myrow = rbase + 8 * (lane & 3);
    if (!kMasked || myrow < M) atomicAdd(P.rowsum + myrow, racc[0]);
  }

  template <int BN>
  static __device__ __forceinline__ void run(const Params& P, uint32_t tacc, int m0, int n0, int M, int N, int half,
                                             int q, int lane, int ewarp, float* smem, const CUtensorMap* cmap,
                                             uint8_t* staging, int force_atomic, float& carry) {
    (void)smem; (void)force_atomic; (void)carry;
    const float s = __ldg(P.scale_ptr);
    const float sl2 = s * 1.4426950408889634f;
    const bool interior = (m0 + kBM <= M) && (n0 + BN <= N);
    // does this tile contain matching pairs?  columns of rows [m0, m0+128) are [m0+off, m0+off+128)
    const bool has_diag = (m0 + P.diag_offset < n0 + BN) && (m0 + P.diag_offset + kBM > n0);
    uint8_t* box = staging + ewarp * 4096;
    if (interior && !has_diag) tile<BN, false, false>(P, tacc, m0, n0, M, N, half, q, lane, s, sl2, cmap, box);
    else if (interior) tile<BN, false, true>(P, tacc, m0, n0, M, N, half, q, lane, s, sl2, cmap, box);
    else tile<BN, true, true>(P, tacc, m0, n0, M, N, half, q, lane, s, sl2, cmap, box);
  }
};

// Backward InfoNCE epilogue: recomputed cosines -> gradient coefficients of one logit tile,
//   g[r][c] = E[r][c] * (rinv[r] + cinv[c])           off the diagonal,
//   g[r][r'] = 0 (zero_diag: the matching pair is applied in fp32 by mmg_infonce_bwd_diag) or  E*(..) - dcoef,
//   rinv[r] = s*gl/(2B*rowsum[r]),  cinv[c] = s*gl/(2B*colsum[c]),  dcoef = s*gl/B        (gl = d loss)
// i.e. g = s * dloss/dlogit, so that dI = g . T and dT = g^T . I need no further scaling.  g goes out as bf16 into a
// block scratch (never the full B x B) that the two gradient GEMMs consume; sum(g * cos) accumulates d loss / d log(s).
// One accumulator row per thread (32x32b layout).  Everything is warp-local: each warp keeps a private copy of the
// column terms cinv of its columns in shared memory (broadcast 16-byte loads), assembles its 32 x 64 block of bf16
// coefficients in its own 128B-swizzled staging box and ships it with a TMA bulk-tensor store -- full 128-byte lines,
// no LSU traffic to global memory, no CTA-level barrier, ragged edges clipped by the hardware.  No masking of g is
// needed: out-of-range operand rows are zero-filled, so cos = 0 there and g*cos contributes nothing.
template <int kW>
struct EpiGradT {
  struct Params {
    const float* rinv;        // [M]
    const float* cinv;        // [N]
    const float* scale_ptr;   // device scalar s
    const float* scal;        // device scalars: [0] diagonal coefficient subtracted here, [2] != 0 -> zero the diagonal
    float* dlogscale_acc;     // device scalar accumulator: sum g * cos; NULL = not wanted (logit_scale is not trained)
    int diag_offset;          // (global column index of local row 0) - (global column index of block column 0)
    int g_row_off;            // row offset of this block inside the coefficient scratch (fused backward: buffer * Rb)
    int dbg;                  // measurement hook: bit 0 = skip the math (g = cos), bit 1 = skip staging + store
    int g_f16;                // coefficients go out as fp16 (MMG_PREC_F16: rinv / cinv then carry the 2^14 scaling and
                              // scal[3] the factor the gradient epilogues multiply back in) instead of bf16
  };
  static constexpr int kWarps = kW;
  // column terms of a tile's columns, one copy per accumulator stage: 2 x 256 floats
  static constexpr int kScratchBytes = 2048;
  // one [32 rows x 64 bf16] (4 KB, 128B-swizzled) staging box per epilogue warp
  template <int BN> __host__ __device__ static constexpr int staging_bytes() { return kWarps * 4096; }

  // 32 accumulator columns of this thread's row -> coefficients, packed to bf16 into the row's slot of the staging box
  // (cs: the column terms of these 32 columns in shared memory -- broadcast LDS.128)
  template <bool kDiag>
  static __device__ __forceinline__ void chunk(float (&v)[32], const float* cs, float ri, float sl2, int c0, int dcol,
                                               float dcoef, bool zero_diag, bool dls, float (&dacc)[2], uint8_t* rowp,
                                               uint32_t sw, uint32_t cbase, int dbg, int g_f16) {
    if (!(dbg & 1)) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 c4 = *reinterpret_cast<const float4*>(cs + 4 * j4);
        const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = 4 * j4 + jj;
          const float cosv = v[j];
          const float e = ex2_approx(fmaf(cosv, sl2, -sl2));
          float g = e * (ri + cv[jj]);
          if (kDiag && c0 + j == dcol) g = zero_diag ? 0.f : g - dcoef;
          if (dls) dacc[jj & 1] = fmaf(g, cosv, dacc[jj & 1]);  // two independent chains (warp-uniform predicate)
          v[j] = g;
        }
      }
    }
    if (!(dbg & 2)) {
      if (g_f16) {  // warp-uniform
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          o.x = pack_f16x2(v[8 * j + 0], v[8 * j + 1]);
          o.y = pack_f16x2(v[8 * j + 2], v[8 * j + 3]);
          o.z = pack_f16x2(v[8 * j + 4], v[8 * j + 5]);
          o.w = pack_f16x2(v[8 * j + 6], v[8 * j + 7]);
          *reinterpret_cast<uint4*>(rowp + (((cbase + j) ^ sw) << 4)) = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
          o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
          o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
          o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
          *reinterpret_cast<uint4*>(rowp + (((cbase + j) ^ sw) << 4)) = o;
        }
      }
    }
  }

  // One thread = one accumulator row, 32 columns at a time.  (Variants measured slower at 32768^2 x 512 and not kept,
  // coefficient launches 1.02-1.06 ms as written: column terms read from global memory instead of shared memory,
  // 1.23 ms; 16-column TMEM loads issued one step ahead, 1.28 ms; 32-column loads into two register buffers issued one
  // chunk ahead, 1.18 ms; one 64-column TMEM load per staging box, 1.29 ms.)
  template <int BN, bool kDiag>
  static __device__ __forceinline__ void tile(const Params& P, uint32_t tacc, int m0, int n0, int M, int N, int half,
                                              int q, int lane, const float* cs, float sl2, const CUtensorMap* cmap,
                                              uint8_t* box, float& carry) {
    constexpr int kCols = BN / (kWarps / 4);
    constexpr int kChunks = kCols / 32;
    const bool dls = P.dlogscale_acc != nullptr;
    const int row = m0 + q * 32 + lane;
    const int dcol = row + P.diag_offset;
    const float ri = (row < M) ? __ldg(P.rinv + row) : 0.f;
    float dcoef = 0.f;
    bool zero_diag = false;
    if (kDiag) {
      dcoef = __ldg(P.scal);
      zero_diag = __ldg(P.scal + 2) != 0.f;
    }
    uint8_t* rowp = box + lane * 128;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    const int cbeg = n0 + half * kCols;                       // first column of this warp
    const int nch = min(kChunks, (N - cbeg + 31) / 32);       // chunks that hold real columns (warp-uniform)
    const int dbg = P.dbg;
    float dacc[2] = {0.f, 0.f};
    float va[32];
#pragma unroll 1
    for (int ch = 0; ch < nch; ++ch) {
      tmem_ld_32x32b_x32(tacc + half * kCols + ch * 32, va);
      tmem_ld_wait();
      if ((ch & 1) == 0) {
        if (lane == 0) tma_store_wait_read();  // this warp's previous box has left shared memory
        __syncwarp();
      }
      chunk<kDiag>(va, cs + ch * 32, ri, sl2, cbeg + ch * 32, dcol, dcoef, zero_diag, dls, dacc, rowp, sw,
                   static_cast<uint32_t>((ch & 1) * 4), dbg, P.g_f16);
      if ((ch & 1) == 1 || ch + 1 == nch) {
        if (!(dbg & 2)) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            // columns beyond N are clipped by the hardware
            tma_store_2d(cmap, box, cbeg + (ch & ~1) * 32, P.g_row_off + m0 + q * 32);
            tma_store_commit();
          }
        }
      }
    }
    if (dls) carry += dacc[0] + dacc[1];  // per-thread partial; reduced and added once per warp in finish()
  }

  template <int BN>
  static __device__ __forceinline__ void run(const Params& P, uint32_t tacc, int m0, int n0, int M, int N, int half,
                                             int q, int lane, int ewarp, float* smem, const CUtensorMap* cmap,
                                             uint8_t* staging, int force_atomic, float& carry) {
    (void)force_atomic;
    constexpr int kCols = BN / (kWarps / 4);
    // Column terms of this warp's columns (zero beyond N) in shared memory.  `smem` is shared by the four warps that
    // work on the same columns of the same tile: each writes all the values it will read (identical values, so the
    // overlap is benign) and only synchronises with itself; the caller alternates two copies by accumulator stage, and
    // a warp cannot be two tiles ahead of another (the MMA of tile t+2 needs every warp's release of tile t).
    __syncwarp();
#pragma unroll
    for (int i = lane; i < kCols; i += 32) {
      const int c = n0 + half * kCols + i;
      smem[i] = (c < N) ? __ldg(P.cinv + c) : 0.f;
    }
    __syncwarp();
    const float s = __ldg(P.scale_ptr);
    const float sl2 = s * 1.4426950408889634f;
    const bool has_diag = (m0 + P.diag_offset < n0 + BN) && (m0 + P.diag_offset + kBM > n0);
    uint8_t* box = staging + ewarp * 4096;
    if (has_diag) tile<BN, true>(P, tacc, m0, n0, M, N, half, q, lane, smem, sl2, cmap, box, carry);
    else tile<BN, false>(P, tacc, m0, n0, M, N, half, q, lane, smem, sl2, cmap, box, carry);
  }

  // sum g*cos: one atomic per warp per launch (not per tile: 10^5 same-address atomics per launch serialise at L2)
  static __device__ __forceinline__ void finish(const Params& P, float carry, int lane) {
    if (P.dlogscale_acc != nullptr) {
      const float d = warp_sum(carry) * __ldg(P.scal + 3);  // scal[3]: 1, or the factor that undoes the fp16 scaling
      if (lane == 0 && d != 0.f) atomicAdd(P.dlogscale_acc, d);
    }
    if (lane == 0) tma_store_wait_all();
  }
};

// ---------------------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------------------
template <int BN, class Epi, int kCG>
__global__ void __launch_bounds__(32 * (4 + Epi::kWarps), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
               const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ CUtensorMap tmC0, const __grid_constant__ CUtensorMap tmC1,
               const GemmProblem p0, const GemmProblem p1, const TailSplit tail, const typename Epi::Params e0,
               const typename Epi::Params e1) {
  using S = GemmSmem<BN, kCG, Epi::template staging_bytes<BN>(), Epi::kScratchBytes>;
  constexpr int kStages = S::kStages;
  constexpr uint32_t kTmemCols = 2 * BN;  // 256 or 512: a power of two >= 32
  constexpr int kTileM = kBM * kCG;       // rows of one (pair) tile

  // Pointers are derived from the __shared__ array by pointer arithmetic only (an integer round trip would turn every
  // access below into a generic LD/ST instead of LDS/STS).  128B-swizzled TMA boxes need 1024-byte alignment.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("[mmgclip_b200] dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* smem = smem_raw;
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * S::kABytes;
  uint8_t* staging = sB + kStages * S::kBBytes;  // epilogue output tile (TMA-store epilogues), 1024-byte aligned
  float* epi_scratch = reinterpret_cast<float*>(staging + Epi::template staging_bytes<BN>());
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Epi::template staging_bytes<BN>() + Epi::kScratchBytes);
  uint64_t* full_bar = bars;                       // [kStages]  TMA -> MMA          (leader's copy is the live one)
  uint64_t* empty_bar = bars + kStages;            // [kStages]  MMA -> TMA          (multicast to both CTAs)
  uint64_t* tfull_bar = bars + 2 * kStages;        // [2]        MMA -> epilogue     (multicast to both CTAs)
  uint64_t* tempty_bar = bars + 2 * kStages + 2;   // [2]        epilogue -> MMA     (leader's copy is the live one)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kCG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int tile_first = blockIdx.x / kCG;
  const int tile_step = gridDim.x / kCG;

  if (kCG == 2) cluster_sync_all();  // both CTAs resident before the pair-wide TMEM allocation
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (p1.tiles_m > 0) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], kCG);   // one arrival per producer of the pair
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], Epi::kWarps * kCG);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (kCG == 2) tmem_alloc_cg2(tmem_slot, kTmemCols);
    else tmem_alloc(tmem_slot, kTmemCols);
  }
  tcgen05_fence_before();
  if (kCG == 2) cluster_sync_all();
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // set-up done (nothing above touches global memory): the next kernel may be scheduled; from here on this one reads what
  // its predecessor wrote
  pdl_entry();

  const int real_tiles = p0.num_tiles() + p1.num_tiles();
  const int total_tiles = tail.rem_split > 1 ? tail.full_tiles + (real_tiles - tail.full_tiles) * tail.rem_split
                                             : real_tiles;

  // 16 epilogue warps: 640 threads only get 96 registers each at launch.  The four non-epilogue warps (one warpgroup)
  // release 128 x (96 - 56) = 5120 registers (setmaxnreg.dec) and the four epilogue warpgroups take 512 x (104 - 96) = 4096
  // of them (setmaxnreg.inc blocks until the CTA's pool holds enough: asking for more than was released dead-locks).
  if (warp < 4) {
  if (Epi::kWarps == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ===================== TMA producer (one per CTA) =====================
    // The warp stays converged; one elected lane arms the barrier and issues the bulk-tensor copies.
    int stage = 0;
    uint32_t phase = 0;
    for (int t = tile_first; t < total_tiles; t += tile_step) {
      const TileCoord tc = decode_virtual(t, p0, p1, tail);
      const GemmProblem& p = tc.prob ? p1 : p0;
      const int m0 = tc.m_blk * kTileM + static_cast<int>(cta_rank) * kBM;       // this CTA's A rows
      const int n0 = tc.n_blk * BN + static_cast<int>(cta_rank) * S::kBRows;     // this CTA's share of the B rows
      const int seg_kb = p.seg_kb();
      for (int kb = tc.kb_begin; kb < tc.kb_end; ++kb) {
        // K segment of this block (problem 0 only): which of the two A / B tensor maps, and the K offset inside it
        const int seg = p.k_segs > 1 ? kb / seg_kb : 0;
        const CUtensorMap* mA = (tc.prob || ((p.seg_a >> seg) & 1)) ? &tmA1 : &tmA0;
        const CUtensorMap* mB = (tc.prob || ((p.seg_b >> seg) & 1)) ? &tmB1 : &tmB0;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], kCG * (S::kABytes + S::kBBytes));
          else mbar_arrive_cluster(&full_bar[stage], 0);
          uint8_t* a_dst = sA + stage * S::kABytes;
          uint8_t* b_dst = sB + stage * S::kBBytes;
          const int k0 = (kb - seg * seg_kb) * kBK;
          auto load = [&](const CUtensorMap* m, void* dst, int c0, int c1) {
            if (kCG == 2) tma_load_2d_cg2(m, &full_bar[stage], dst, c0, c1, kEvictNormal);
            else tma_load_2d(m, &full_bar[stage], dst, c0, c1, kEvictNormal);
          };
          if (!p.a_mn) {
            load(mA, a_dst, k0, m0);
          } else {
#pragma unroll
            for (int i = 0; i < kBM / 64; ++i) load(mA, a_dst + i * (kBK * 128), m0 + i * 64, k0);
          }
          if (!p.b_mn) {
            load(mB, b_dst, k0, n0);
          } else {
#pragma unroll
            for (int i = 0; i < S::kBRows / 64; ++i) load(mB, b_dst + i * (kBK * 128), n0 + i * 64, k0);
          }
          // long K loops stream at least one operand from HBM: pull the boxes kPrefetchKb blocks ahead into L2
          if (tail.prefetch_kb > 0 && p.k_segs <= 1 && kb + tail.prefetch_kb < tc.kb_end) {
            const int kp = k0 + tail.prefetch_kb * kBK;
            if (!p.a_mn) {
              tma_prefetch_l2_2d(mA, kp, m0);
            } else {
#pragma unroll
              for (int i = 0; i < kBM / 64; ++i) tma_prefetch_l2_2d(mA, m0 + i * 64, kp);
            }
            if (!p.b_mn) {
              tma_prefetch_l2_2d(mB, kp, n0);
            } else {
#pragma unroll
              for (int i = 0; i < S::kBRows / 64; ++i) tma_prefetch_l2_2d(mB, n0 + i * 64, kp);
            }
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = tile_first; t < total_tiles; t += tile_step, ++it) {
        const TileCoord tc = decode_virtual(t, p0, p1, tail);
        const GemmProblem& p = tc.prob ? p1 : p0;
        const int acc_stage = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc_stage], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t idesc = make_idesc_16(kTileM, BN, p.a_mn, p.b_mn, p.a_f16, p.b_f16);
        const uint32_t tmem_d = tmem_base + acc_stage * BN;
        // K-major: 8-row groups are 1024 B apart (SBO); one swizzle span along K, LBO unused.
        // MN-major: 8-k groups are 1024 B apart (SBO); 64-element MN atoms are kBK*128 B apart (LBO).
        const uint32_t a_lbo = p.a_mn ? kBK * 128 : 0, b_lbo = p.b_mn ? kBK * 128 : 0;
        const uint32_t a_kstep = p.a_mn ? kUmmaK * 128 : kUmmaK * 2;
        const uint32_t b_kstep = p.b_mn ? kUmmaK * 128 : kUmmaK * 2;
        for (int kb = tc.kb_begin; kb < tc.kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (elect_one_sync()) {
            const uint32_t a_addr = smem_u32(sA + stage * S::kABytes);
            const uint32_t b_addr = smem_u32(sB + stage * S::kBBytes);
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              const uint64_t da = make_smem_desc_sw128(a_addr + k * a_kstep, a_lbo, 1024);
              const uint64_t db = make_smem_desc_sw128(b_addr + k * b_kstep, b_lbo, 1024);
              const uint32_t acc = (kb > tc.kb_begin || k > 0) ? 1u : 0u;
              if (kCG == 2) umma_bf16_ss_cg2(tmem_d, da, db, idesc, acc);
              else umma_bf16_ss(tmem_d, da, db, idesc, acc);
            }
            if (kCG == 2) {
              umma_commit_cg2(&empty_bar[stage]);
              if (kb == tc.kb_end - 1) umma_commit_cg2(&tfull_bar[acc_stage]);
            } else {
              umma_commit(&empty_bar[stage]);
              if (kb == tc.kb_end - 1) umma_commit(&tfull_bar[acc_stage]);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (tc.kb_begin >= tc.kb_end) {
          // Degenerate split (no K blocks): nothing was issued; still publish the (stale) stage so the
          // epilogue does not dead-lock.  Host code never creates such splits; kept as a guard.
          if (elect_one_sync()) {
            if (kCG == 2) umma_commit_cg2(&tfull_bar[acc_stage]);
            else umma_commit(&tfull_bar[acc_stage]);
          }
          __syncwarp();
        }
      }
      if (kCG == 2 && it > 0) {
        // all epilogue arrivals (also the peer's remote ones) must have landed before the leader's barriers die
        const int last = it - 1;
        mbar_wait(&tempty_bar[last & 1], (last >> 1) & 1);
      }
    }
  }
  } else {
    if (Epi::kWarps == 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ===================== epilogue (both CTAs: 128 accumulator rows each) =====================
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;  // which half of the tile's columns
    int it = 0;
    float carry = 0.f;  // per-thread running value an epilogue may keep across tiles (EpiGrad: sum g*cos)
    for (int t = tile_first; t < total_tiles; t += tile_step, ++it) {
      const TileCoord tc = decode_virtual(t, p0, p1, tail);
      const GemmProblem& p = tc.prob ? p1 : p0;
      const int acc_stage = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc_stage], acc_phase);
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_stage * BN;
      if (tc.kb_begin < tc.kb_end)
        Epi::template run<BN>(tc.prob ? e1 : e0, tacc, tc.m_blk * kTileM + static_cast<int>(cta_rank) * kBM,
                              tc.n_blk * BN, p.M, p.N, half, q, lane, warp - 4,
                              epi_scratch + acc_stage * BN + half * (BN / (Epi::kWarps / 4)),
                              tc.prob ? &tmC1 : &tmC0, staging, tc.atomic, carry);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tempty_bar[acc_stage]);
        else mbar_arrive_cluster(&tempty_bar[acc_stage], 0);
      }
    }
    Epi::finish(e0, carry, lane);
  }

  tcgen05_fence_before();
  if (kCG == 2) cluster_sync_all();
  else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    if (kCG == 2) tmem_dealloc_cg2(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace mmg
