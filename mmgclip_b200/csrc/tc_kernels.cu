// mmgclip_b200 -- host-side launchers of the tcgen05 mainloop (gemm_tc.cuh): TMA tensor-map construction,
// tile-shape selection, persistent-grid sizing (one CTA per SM).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "bwd_fused.cuh"
#include "gemm_tc.cuh"
#include "kernels.h"

namespace mmg {

// ---- driver entry point for cuTensorMapEncodeTiled (no link-time libcuda dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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

// bf16 row-major matrix [outer, inner] with row pitch `ld` elements; box = [box_outer, 64] (64 bf16 = one 128-byte
// swizzle span).  Out-of-bounds box elements read as zero, which is what makes ragged edges free in the mainloop.
static int make_tmap(CUtensorMap* m, const void* ptr, long long inner, long long outer, long long ld, int box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error(-4, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return set_error(-2, "bf16 operand pointer must be 16-byte aligned");
  if ((ld & 7) != 0) return set_error(-2, "bf16 operand pitch must be a multiple of 8 elements (got %lld)", ld);
  if (inner <= 0 || outer <= 0) return set_error(-1, "empty operand (%lld x %lld)", outer, inner);
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-4, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return 0;
}

// fp32 row-major output matrix [rows, cols] with pitch ld: box = [32 rows x 32 columns] (one epilogue warp's block)
static int make_out_tmap_f32(CUtensorMap* m, const float* ptr, long long cols, long long rows, long long ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error(-4, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-4, "cuTensorMapEncodeTiled (fp32 output) failed (CUresult %d)", (int)r);
  return 0;
}
// Measured-and-rejected alternatives of the mainloop (A/B switches) only exist in measurement builds (make measure,
// -DMMG_MEASURE): MMG_TC_TMA_OUT=0 (direct stores instead of TMA stores), MMG_TC_DUAL_SPLIT=1, MMG_TC_PAIR=0 (single-CTA
// tiles), MMG_TC_TAIL=1 (tail balancing), MMG_TC_PREFETCH=n (L2 prefetch distance).  The product library is compiled with
// the defaults as constants and reads no environment variable.
static int measure_env(const char* name, int dflt) {
#ifdef MMG_MEASURE
  const char* e = getenv(name);
  return e != nullptr ? atoi(e) : dflt;
#else
  (void)name;
  return dflt;
#endif
}
static bool out_tma_ok(const float* C, long long ldc) {
  static const int en = measure_env("MMG_TC_TMA_OUT", 1);
  return en != 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 && (ldc & 3) == 0;
}

// operand with `rows` along M/N and `K` along the contraction
static int make_operand_map(CUtensorMap* m, const TcOperand& op, int rows, int K, int tile_rows) {
  if (!op.mn_major) return make_tmap(m, op.ptr, K, rows, op.ld, tile_rows);  // [rows, K]: box tile_rows x 64(K)
  return make_tmap(m, op.ptr, rows, K, op.ld, kBK);                          // [K, rows]: box 64(K) x 64(rows)
}

static int sm_count() {
  // per device: a process may drive several GPUs
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (n[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}

static GemmProblem make_problem(int M, int N, int K, int BN, int k_splits, int a_mn, int b_mn, int cg, int a_f16 = 0,
                                int b_f16 = 0) {
  GemmProblem p;
  p.M = M; p.N = N; p.K = K;
  p.tiles_m = (M + kBM * cg - 1) / (kBM * cg);
  p.tiles_n = (N + BN - 1) / BN;
  const int nkb = (K + kBK - 1) / kBK;  // (callers that add K segments re-derive the split afterwards)
  if (k_splits < 1) k_splits = 1;
  if (k_splits > nkb) k_splits = nkb;
  // no empty splits: shrink until the last split still owns a K block
  while (k_splits > 1) {
    const int per = (nkb + k_splits - 1) / k_splits;
    if ((k_splits - 1) * per < nkb) break;
    --k_splits;
  }
  p.k_splits = k_splits;
  p.a_mn = a_mn; p.b_mn = b_mn;
  p.a_f16 = a_f16; p.b_f16 = b_f16;
  p.k_segs = 1; p.seg_a = 0; p.seg_b = 0;
  return p;
}

static GemmProblem empty_problem() {
  GemmProblem p;
  memset(&p, 0, sizeof(p));
  p.k_splits = 1;
  p.k_segs = 1;
  return p;
}

static bool tail_balance_enabled();
static int prefetch_distance();
static bool dual_split_enabled() {
  static const int v = measure_env("MMG_TC_DUAL_SPLIT", 0);  // measured slower at 8192^2 blocks (1.84 vs 1.75 ms)
  return v == 1;
}

template <int BN, class Epi, int kCG>
static int launch(const CUtensorMap& a0, const CUtensorMap& b0, const CUtensorMap& a1, const CUtensorMap& b1,
                  const CUtensorMap& c0, const CUtensorMap& c1, const GemmProblem& p0, const GemmProblem& p1,
                  const typename Epi::Params& e0, const typename Epi::Params& e1, cudaStream_t st,
                  bool balance_tail = false) {
  auto kern = gemm_tc_kernel<BN, Epi, kCG>;
  using S = GemmSmem<BN, kCG, Epi::template staging_bytes<BN>(), Epi::kScratchBytes>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(gemm_tc_kernel)");
    configured = true;
  }
  const int total = p0.num_tiles() + p1.num_tiles();
  if (total <= 0) return 0;
  const int groups = sm_count() / kCG;  // CTAs (or CTA pairs) that can be resident: one per SM
  // tail balancing: cut the tiles of an under-filled last round into K-slices (see TailSplit in gemm_tc.cuh)
  TailSplit tail;
  tail.full_tiles = total;
  tail.rem_split = 1;
  tail.prefetch_kb = prefetch_distance();
  if (balance_tail && tail_balance_enabled() && total > groups && total % groups != 0 && p0.k_splits == 1 &&
      p1.k_splits <= 1) {
    const int rem = total % groups;
    int nkb = (p0.K + kBK - 1) / kBK;
    if (p1.num_tiles() > 0) { const int n1 = (p1.K + kBK - 1) / kBK; nkb = n1 < nkb ? n1 : nkb; }
    int best_s = 1;
    double best = 1.0;  // cost of the tail in units of a full round
    for (int sdiv = 2; sdiv <= 8; ++sdiv) {
      if (nkb / sdiv < 8) break;                                   // keep slices at least 8 K-blocks long
      if (((nkb + sdiv - 1) / sdiv) * (sdiv - 1) >= nkb) continue;  // would create an empty slice
      const int rounds = (rem * sdiv + groups - 1) / groups;
      const double cost = (double)rounds / sdiv + 0.02 * sdiv;     // small penalty per extra slice (atomics, ramp)
      if (cost < best - 1e-9) { best = cost; best_s = sdiv; }
    }
    if (best_s > 1) {
      tail.full_tiles = total - rem;
      tail.rem_split = best_s;
    }
  }
  const int virt = tail.full_tiles + (total - tail.full_tiles) * tail.rem_split;
  const int grid = (virt < groups ? virt : groups) * kCG;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(32 * (4 + Epi::kWarps));
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = st;
  PdlAttr at;
  at.cluster(kCG);
  cfg.attrs = at.a;
  cfg.numAttrs = at.n;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a0, b0, a1, b1, c0, c1, p0, p1, tail, e0, e1);
  if (e != cudaSuccess) return check_cuda(e, "gemm_tc_kernel launch");
  count_launch();
  return 0;
}

// Tile shape: 256-row CTA pairs (cta_group::2) when both extents fill them, else single-CTA 128 x {256,128} tiles.
static bool pair_enabled() {
  static const int v = measure_env("MMG_TC_PAIR", 1);
  return v != 0;
}
static bool tail_balance_enabled() {
  static const int v = measure_env("MMG_TC_TAIL", 0);  // measured slower (scalar atomics in the tail)
  return v == 1;
}
static int prefetch_distance() {
  static const int v = measure_env("MMG_TC_PREFETCH", 0);  // measured slower than no prefetch on B200
  return v < 0 ? 0 : v;
}
// The InfoNCE epilogues run on 8 epilogue warps.  A 16-warp variant (4 column groups, setmaxnreg re-balancing) was
// measured slower on B200 at 32768^2 x 512 (forward 0.754 vs 0.723 ms, coefficient launches 1.095 vs 1.019 ms): the
// staging boxes of 16 warps cost an operand-ring stage, which matters more than the extra latency hiding.
using EpiLse = EpiLseT<8>;
using EpiLseStore = EpiLseT<8, true>;
using EpiGrad = EpiGradT<8>;
struct TileCfg { int BN, cg; };
static TileCfg pick_tile(int M, int N, int work_items_per_tile = 0) {
  TileCfg c;
  c.BN = N > 128 ? 256 : 128;
  c.cg = (pair_enabled() && c.BN == 256 && M > 128) ? 2 : 1;
  if (c.cg == 2 && work_items_per_tile > 0) {
    // under-filled single-problem launch (e.g. a 4096-row head projection: 32 pair tiles on 74 pairs): twice as many
    // single-CTA tiles of half the work fill the SMs instead
    const long long pair_tiles = (long long)((M + 2 * kBM - 1) / (2 * kBM)) * ((N + c.BN - 1) / c.BN);
    if (pair_tiles * work_items_per_tile * 2 <= sm_count()) c.cg = 1;
  }
  return c;
}

#define MMG_DISPATCH(EPI, cfg, ...)                                              \
  do {                                                                           \
    if ((cfg).BN == 256 && (cfg).cg == 2) return launch<256, EPI, 2>(__VA_ARGS__); \
    if ((cfg).BN == 256) return launch<256, EPI, 1>(__VA_ARGS__);                \
    return launch<128, EPI, 1>(__VA_ARGS__);                                     \
  } while (0)

int tc_gemm_store(const TcOperand& A, const TcOperand& B, float* C, long long ldc, int M, int N, int K, float alpha,
                  const float* alpha_dev, const float* bias, int relu, int mode, int k_splits, cudaStream_t st) {
  return tc_gemm_store_seg(A, nullptr, B, nullptr, 1, 0, 0, C, ldc, M, N, K, alpha, alpha_dev, bias, relu, mode, k_splits,
                           st);
}

// C (op)= alpha * sum_i A[(seg_a >> i) & 1] . B[(seg_b >> i) & 1]^T over nseg K segments, one launch, one accumulator.
int tc_gemm_store_seg(const TcOperand& A0, const TcOperand* A1, const TcOperand& B0, const TcOperand* B1, int nseg,
                      int seg_a, int seg_b, float* C, long long ldc, int M, int N, int K, float alpha,
                      const float* alpha_dev, const float* bias, int relu, int mode, int k_splits, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(-1, "tc_gemm: empty problem %dx%dx%d", M, N, K);
  if (k_splits > 1 && mode != 2) return set_error(-1, "tc_gemm: k_splits > 1 needs MMG_ATOMIC_ADD");
  if (k_splits > 1 && (relu || bias)) return set_error(-1, "tc_gemm: bias/ReLU cannot be fused with split-K");
  if (nseg < 1 || nseg > 3) return set_error(-1, "tc_gemm: 1..3 K segments (got %d)", nseg);
  if (nseg > 1 && ((seg_a != 0 && A1 == nullptr) || (seg_b != 0 && B1 == nullptr)))
    return set_error(-1, "tc_gemm: a K segment refers to a missing operand");
  const TileCfg tcfg = pick_tile(M, N, k_splits < 1 ? 1 : k_splits);
  CUtensorMap ma, mb, ma1, mb1;
  int rc;
  if ((rc = make_operand_map(&ma, A0, M, K, kBM)) != 0) return rc;
  if ((rc = make_operand_map(&mb, B0, N, K, tcfg.BN / tcfg.cg)) != 0) return rc;
  ma1 = ma;
  mb1 = mb;
  if (A1 != nullptr && (rc = make_operand_map(&ma1, *A1, M, K, kBM)) != 0) return rc;
  if (B1 != nullptr && (rc = make_operand_map(&mb1, *B1, N, K, tcfg.BN / tcfg.cg)) != 0) return rc;
  GemmProblem p0 = make_problem(M, N, K, tcfg.BN, 1, A0.mn_major, B0.mn_major, tcfg.cg, A0.f16, B0.f16);
  p0.k_segs = nseg;
  p0.seg_a = nseg > 1 ? seg_a : 0;
  p0.seg_b = nseg > 1 ? seg_b : 0;
  {
    // split-K over the concatenated K range; no empty splits
    const int nkb = p0.total_kb();
    int ks = k_splits < 1 ? 1 : (k_splits > nkb ? nkb : k_splits);
    while (ks > 1 && (ks - 1) * ((nkb + ks - 1) / ks) >= nkb) --ks;
    p0.k_splits = ks;
  }
  GemmProblem p1 = empty_problem();
  EpiStoreF32::Params e;
  e.C = C; e.ldc = ldc; e.bias = bias; e.alpha = alpha; e.alpha_ptr = alpha_dev; e.mode = mode; e.relu = relu;
  // act(C_old + v) cannot be expressed as a reduction: accumulate + bias/ReLU keeps the direct read-modify-write path
  e.use_tma = out_tma_ok(C, ldc) && !(mode != 0 && (relu || bias != nullptr));
  CUtensorMap mc = ma;
  if (e.use_tma && (rc = make_out_tmap_f32(&mc, C, N, M, ldc)) != 0) return rc;
  MMG_DISPATCH(EpiStoreF32, tcfg, ma, mb, ma1, mb1, mc, mc, p0, p1, e, e, st);
}

int tc_gemm_dual_accumulate(const TcOperand& A0, const TcOperand& B0, float* C0, long long ldc0, int M0, int N0, int K0,
                            const TcOperand& A1, const TcOperand& B1, float* C1, long long ldc1, int M1, int N1, int K1,
                            cudaStream_t st, const float* alpha_dev) {
  if (N0 != N1) return set_error(-1, "tc_gemm_dual: both problems must share N");
  const TileCfg tcfg = pick_tile(M0 < M1 ? M0 : M1, N0);
  CUtensorMap ma0, mb0, ma1, mb1;
  int rc;
  if ((rc = make_operand_map(&ma0, A0, M0, K0, kBM)) != 0) return rc;
  if ((rc = make_operand_map(&mb0, B0, N0, K0, tcfg.BN / tcfg.cg)) != 0) return rc;
  if ((rc = make_operand_map(&ma1, A1, M1, K1, kBM)) != 0) return rc;
  if ((rc = make_operand_map(&mb1, B1, N1, K1, tcfg.BN / tcfg.cg)) != 0) return rc;
  EpiStoreF32::Params e0, e1;
  e0.C = C0; e0.ldc = ldc0; e0.bias = nullptr; e0.alpha = 1.f; e0.alpha_ptr = alpha_dev; e0.mode = 1; e0.relu = 0;
  e0.use_tma = out_tma_ok(C0, ldc0) && out_tma_ok(C1, ldc1);
  // Wave quantisation: the two problems of one 8192 x 8192 block are 128 pair tiles on 74 CTA pairs = 2 rounds at 86%.
  // With TMA reduce-add outputs (atomic at L2) every tile may be cut into K slices, so pick the split whose slice
  // count fills whole rounds best (128 x 4 = 512 slices = 6.92 rounds of 7 -> 99%); slices stay >= 16 K-blocks long.
  int ks = 1;
  if (e0.use_tma && dual_split_enabled()) {
    const int groups = sm_count() / tcfg.cg;
    const int t0 = ((M0 + kBM * tcfg.cg - 1) / (kBM * tcfg.cg)) * ((N0 + tcfg.BN - 1) / tcfg.BN);
    const int t1 = ((M1 + kBM * tcfg.cg - 1) / (kBM * tcfg.cg)) * ((N1 + tcfg.BN - 1) / tcfg.BN);
    const int nkb = ((K0 < K1 ? K0 : K1) + kBK - 1) / kBK;
    double best = 0.0;
    for (int sdiv = 1; sdiv <= 8; ++sdiv) {
      if (sdiv > 1 && nkb / sdiv < 16) break;
      const int total = (t0 + t1) * sdiv;
      const int rounds = (total + groups - 1) / groups;
      const double eff = (double)total / ((double)rounds * groups) - 0.004 * (sdiv - 1);  // slight bias to fewer slices
      if (eff > best + 1e-9) { best = eff; ks = sdiv; }
    }
  }
  GemmProblem p0 = make_problem(M0, N0, K0, tcfg.BN, ks, A0.mn_major, B0.mn_major, tcfg.cg, A0.f16, B0.f16);
  GemmProblem p1 = make_problem(M1, N1, K1, tcfg.BN, ks, A1.mn_major, B1.mn_major, tcfg.cg, A1.f16, B1.f16);
  e1 = e0;
  e1.C = C1; e1.ldc = ldc1;
  CUtensorMap mc0 = ma0, mc1 = ma0;
  if (e0.use_tma) {
    if ((rc = make_out_tmap_f32(&mc0, C0, N0, M0, ldc0)) != 0) return rc;
    if ((rc = make_out_tmap_f32(&mc1, C1, N1, M1, ldc1)) != 0) return rc;
  }
  MMG_DISPATCH(EpiStoreF32, tcfg, ma0, mb0, ma1, mb1, mc0, mc1, p0, p1, e0, e1, st, true);
}

int tc_infonce_fwd(const void* a_hat, const void* b_hat, int rows, int cols, int D, int diag_offset, const float* scale,
                   float* rowsum, float* colsum, float* diag, void* e_out, long long lde, cudaStream_t st, int emb_f16) {
  const TileCfg tcfg = pick_tile(rows, cols);
  TcOperand A{a_hat, D, 0}, B{b_hat, D, 0};
  CUtensorMap ma, mb;
  int rc;
  if ((rc = make_operand_map(&ma, A, rows, D, kBM)) != 0) return rc;
  if ((rc = make_operand_map(&mb, B, cols, D, tcfg.BN / tcfg.cg)) != 0) return rc;
  GemmProblem p0 = make_problem(rows, cols, D, tcfg.BN, 1, 0, 0, tcfg.cg, emb_f16, emb_f16);
  GemmProblem p1 = empty_problem();
  EpiLse::Params e;
  e.rowsum = rowsum; e.colsum = colsum; e.diag = diag; e.scale_ptr = scale; e.diag_offset = diag_offset;
  if (e_out != nullptr) {
    // only the CTA-pair 256 x 256 tile variant of the E-storing epilogue has been run on hardware (it is the one every
    // shape accepted by mmg_infonce_stored_supported takes)
    if (!(tcfg.BN == 256 && tcfg.cg == 2))
      return set_error(-3, "mmg_infonce_fwd_store: needs rows > 128 and cols > 128 (256 x 256 pair tiles)");
    // also keep E = exp(logit - s) as bf16 [rows, cols] for the stored-E backward: box = one epilogue warp's 32 x 64 block
    CUtensorMap mc;
    if ((rc = make_tmap(&mc, e_out, cols, rows, lde, 32)) != 0) return rc;
    EpiLseStore::Params es;
    es.rowsum = rowsum; es.colsum = colsum; es.diag = diag; es.scale_ptr = scale; es.diag_offset = diag_offset;
    return launch<256, EpiLseStore, 2>(ma, mb, ma, mb, mc, mc, p0, p1, es, es, st);
  }
  MMG_DISPATCH(EpiLse, tcfg, ma, mb, ma, mb, ma, ma, p0, p1, e, e, st);
}

int tc_infonce_grad_block(const void* a_blk, const void* b_blk, int rb, int cb, int D, int diag_offset,
                          const float* scale, const float* rinv, const float* cinv, const float* scal, void* G,
                          long long ldg, float* dlogscale_acc, cudaStream_t st, int emb_f16) {
  const TileCfg tcfg = pick_tile(rb, cb);
  TcOperand A{a_blk, D, 0}, B{b_blk, D, 0};
  CUtensorMap ma, mb;
  int rc;
  if ((rc = make_operand_map(&ma, A, rb, D, kBM)) != 0) return rc;
  if ((rc = make_operand_map(&mb, B, cb, D, tcfg.BN / tcfg.cg)) != 0) return rc;
  GemmProblem p0 = make_problem(rb, cb, D, tcfg.BN, 1, 0, 0, tcfg.cg, emb_f16, emb_f16);
  GemmProblem p1 = empty_problem();
  // output map: g block [rb, cb] bf16 (pitch ldg), box = 32 rows x 64 columns (one epilogue warp's block)
  CUtensorMap mc;
  if ((rc = make_tmap(&mc, G, cb, rb, ldg, 32)) != 0) return rc;
  const int dbg = measure_env("MMG_EPI_DBG", 0);  // measurement builds only (tests/gpu_epi_probe.py): wrong results when set
  EpiGrad::Params e;
  e.rinv = rinv; e.cinv = cinv; e.scale_ptr = scale;
  e.scal = scal; e.dlogscale_acc = dlogscale_acc; e.diag_offset = diag_offset; e.g_row_off = 0; e.dbg = dbg; e.g_f16 = emb_f16 ? 1 : 0;
  MMG_DISPATCH(EpiGrad, tcfg, ma, mb, ma, mb, mc, mc, p0, p1, e, e, st);
}


// ---------------------------------------------------------------------------------------------------------
// fused persistent backward (bwd_fused.cuh)
// ---------------------------------------------------------------------------------------------------------
// Plan knobs of the fused backward (fused on/off, block shape, scratch buffers, slice lengths, block order): process-wide
// overrides set through mmg_tune() -- the tests walk other schedules with it; unset = the measured defaults.  Every value
// gives correct results.  Measurement builds also honour the MMG_* environment variables of the same names.
enum TuneKey { kTuneFused, kTuneRb, kTuneCb, kTuneNbuf, kTuneKsl, kTuneKslT, kTuneSr, kTuneSc, kTuneRot, kTunePdl, kTuneCount };
static const char* const kTuneNames[kTuneCount] = {"fused", "fused_rb", "fused_cb", "fused_nbuf", "fused_ksl",
                                                   "fused_ksl_t", "fused_sr", "fused_sc", "fused_rot", "pdl"};
static const char* const kTuneEnv[kTuneCount] = {"MMG_BWD_FUSED", "MMG_FUSED_RB", "MMG_FUSED_CB", "MMG_FUSED_NBUF",
                                                 "MMG_FUSED_KSL", "MMG_FUSED_KSL_T", "MMG_FUSED_SR", "MMG_FUSED_SC",
                                                 "MMG_FUSED_ROT", "MMG_PDL"};
static std::mutex g_tune_mu;
static int g_tune[kTuneCount] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1};  // -1 = unset

int tc_tune(const char* key, int value) {
  if (key == nullptr) return set_error(-1, "mmg_tune: key is NULL");
  std::lock_guard<std::mutex> lock(g_tune_mu);
  if (strcmp(key, "reset") == 0) {
    for (int i = 0; i < kTuneCount; ++i) g_tune[i] = -1;
    return 0;
  }
  for (int i = 0; i < kTuneCount; ++i)
    if (strcmp(key, kTuneNames[i]) == 0) {
      g_tune[i] = value < 0 ? -1 : value;
      return 0;
    }
  return set_error(-1, "mmg_tune: unknown key '%s'", key);
}

static int tune_int(TuneKey k, int dflt) {
  int v;
  {
    std::lock_guard<std::mutex> lock(g_tune_mu);
    v = g_tune[k];
  }
  if (v >= 0) return v;
  return measure_env(kTuneEnv[k], dflt);
}

bool pdl_enabled() { return tune_int(kTunePdl, 1) != 0; }

struct FusedPlan {
  int ok, Rb, Cb, nbuf, kslI, kslT;
  size_t g_bytes, total_bytes;
  size_t trace_off, trace_bytes;  // MMG_FUSED_TRACE=1: debug timeline region at the end of the workspace
};

static int largest_divisor(int n, int cap) {
  for (int b = cap; b >= 256; b >>= 1)
    if (n % b == 0) return b;
  return 0;
}

static FusedPlan fused_plan(int rows, int cols, int D, int col_unit = 0, bool remote_owners = false) {
  FusedPlan f;
  memset(&f, 0, sizeof(f));
  if (tune_int(kTuneFused, 1) == 0) return f;
  if (rows < 256 || cols < 256 || D < 256 || (D % 256) != 0) return f;
  int Rb = tune_int(kTuneRb, 0), Cb = tune_int(kTuneCb, 0);
  if (Rb <= 0) Rb = largest_divisor(rows, 4096);
  // col_unit > 0: column blocks must also divide col_unit (the columns of one part of one owner)
  if (Cb <= 0) Cb = largest_divisor(col_unit > 0 ? col_unit : cols, 2048);
  if (col_unit > 0 && Cb >= 256 && (col_unit % Cb) != 0) return f;
  if (Rb < 256 || Cb < 256 || (Rb % 256) || (Cb % 256) || (rows % Rb) || (cols % Cb)) return f;
  // K blocks (of 64) per gradient slice: every slice ends in a 256 x 256 fp32 reduce-add at L2, so slices are long
  // (Remote owners -- the column-side gradient goes to other GPUs over NVLink -- take whole-K dB slices: every element
  // then crosses the link once per row block instead of once per slice; 8 x B200: 0.775 vs 0.945 ms per step.)
  int kslI = tune_int(kTuneKsl, 32);
  int kslT = tune_int(kTuneKslT, remote_owners ? Rb / kBK : kslI);
  while (kslI > 1 && ((Cb / kBK) % kslI) != 0) kslI >>= 1;
  while (kslT > 1 && ((Rb / kBK) % kslT) != 0) kslT >>= 1;
  if (kslI < 1 || kslT < 1) return f;
  int nbuf = tune_int(kTuneNbuf, 4);
  if (nbuf < 3) nbuf = 3;
  const int nblk = (rows / Rb) * (cols / Cb);
  if (nbuf > nblk) nbuf = nblk < 1 ? 1 : nblk;
  f.Rb = Rb; f.Cb = Cb; f.nbuf = nbuf; f.kslI = kslI; f.kslT = kslT;
  f.g_bytes = ((size_t)nbuf * Rb * Cb * 2 + 255) / 256 * 256;
  // counters: doneA[nblk], doneB[nblk]
  f.total_bytes = f.g_bytes + (size_t)nblk * 2 * sizeof(unsigned int) + 256;
  if (measure_env("MMG_FUSED_TRACE", 0) != 0) {
    f.trace_off = (f.total_bytes + 255) / 256 * 256;
    f.trace_bytes = (size_t)sm_count() * kTraceRoles * kTraceCap * 16;
    f.total_bytes = f.trace_off + f.trace_bytes;
  }
  f.ok = 1;
  return f;
}

// Everything of BwdFusedParams that defines the work-item schedule (no pointers).
static void fill_schedule(BwdFusedParams* pp, const FusedPlan& f, int rows, int cols, int D, int n_owners, int n_parts,
                          int part, int BN) {
  BwdFusedParams& p = *pp;
  memset(&p, 0, sizeof(p));
  const int owner_rows = cols / n_owners;
  p.rows = rows; p.cols = cols; p.D = D;
  p.Rb = f.Rb; p.Cb = f.Cb;
  p.nbc = cols / f.Cb / n_parts;   // column blocks this launch covers
  p.nblk = (rows / f.Rb) * p.nbc;
  p.nbuf = f.nbuf;
  p.tAm = f.Rb / 256; p.tAn = f.Cb / BN; p.tDn = D / BN;
  p.kslI = f.kslI; p.kslT = f.kslT;
  p.sI = (f.Cb / kBK) / f.kslI;
  p.sT = (f.Rb / kBK) / f.kslT;
  p.nA = p.tAm * p.tAn;
  p.nBI = p.tAm * p.tDn * p.sI;
  p.nB = p.nBI + p.tAn * p.tDn * p.sT;
  p.owner_rows = owner_rows;
  if (n_parts > 1) {
    p.blocks_per_owner = owner_rows / f.Cb;
    p.blocks_per_part = p.blocks_per_owner / n_parts;
  } else {
    p.blocks_per_owner = p.blocks_per_part = p.nbc;  // one part: block index == global block index
  }
  p.part = part;
  p.part_row0 = part * p.blocks_per_part * f.Cb;
  if (p.nbuf > p.nblk) p.nbuf = p.nblk < 1 ? 1 : p.nblk;
  // super-tile of the block order (BwdFusedParams::block_rc).  Default 1 x 1 = row-major: at 32768^2 x 512 super-tiles
  // of 2x2 .. 4x4 blocks measured the same or slightly slower (2.52 -> 2.54 ms; stored-E 2.16 -> 2.20 ms).  mmg_tune
  // "fused_sr" / "fused_sc" select others.
  int sr = tune_int(kTuneSr, 1), sc = tune_int(kTuneSc, 1);
  const int nbr = rows / f.Rb;
  if (sr <= 0) sr = 1;
  if (sc <= 0) sc = 1;
  while (sr > 1 && (nbr % sr) != 0) sr >>= 1;
  while (sc > 1 && (p.nbc % sc) != 0) sc >>= 1;
  if (sr < 1 || (nbr % sr) != 0) sr = 1;
  if (sc < 1 || (p.nbc % sc) != 0) sc = 1;
  p.sr = sr; p.sc = sc;
}

// Host-side enumeration of the fused backward's static schedule (no GPU needed): the items CTA pair `pair` of `pairs`
// walks, in order, as rows of {type, block, tm, tn, kb0, nkb, global column block, row block}.  info[8] = {Rb, Cb, nbuf, nA, nB,
// nblk, kslI, kslT}.  Returns the pair's item count (items beyond max_items are counted but not written), 0 when the
// shape is not covered by the fused kernel.  Used by tests/test_fused_schedule_cpu.py to check the dead-lock freedom
// argument of bwd_fused.cuh for arbitrary shapes.
int tc_fused_bwd_schedule(int rows, int cols, int D, int n_owners, int n_parts, int part, int pairs, int pair, int* items,
                          int max_items, int* info) {
  if (n_owners < 1 || n_parts < 1 || part < 0 || part >= n_parts || pairs < 1 || pair < 0 || pair >= pairs) return 0;
  if ((cols % n_owners) != 0) return 0;
  const int owner_rows = cols / n_owners;
  if (n_parts > 1 && ((owner_rows % n_parts) != 0 || ((owner_rows / n_parts) % 256) != 0)) return 0;
  const FusedPlan f = fused_plan(rows, cols, D, n_parts > 1 ? owner_rows / n_parts : 0, n_owners > 1);
  if (!f.ok) return 0;
  BwdFusedParams p;
  fill_schedule(&p, f, rows, cols, D, n_owners, n_parts, part, 256);
  if (info != nullptr) {
    info[0] = p.Rb; info[1] = p.Cb; info[2] = p.nbuf; info[3] = p.nA; info[4] = p.nB; info[5] = p.nblk;
    info[6] = p.kslI; info[7] = p.kslT;
  }
  BwdCursor cur;
  cur.init(p, pair, pairs);
  BwdItem it;
  int n = 0;
  while (cur.next(p, it)) {
    if (items != nullptr && n < max_items) {
      int* o = items + 8 * n;
      int rb, cbl;
      p.block_rc(it.blk, rb, cbl);
      o[0] = it.type; o[1] = it.blk; o[2] = it.tm; o[3] = it.tn; o[4] = it.kb0; o[5] = it.nkb;
      o[6] = p.global_cb(cbl); o[7] = rb;
    }
    ++n;
  }
  return n;
}

// Can the stored-E backward take this shape?  (Same conditions as the fused launch; the forward must then have been
// asked to keep E.)
int tc_infonce_stored_supported(int rows, int cols, int D, int n_owners, int n_parts) {
  if (n_owners < 1 || n_owners > kMaxOwners || (cols % n_owners) != 0 || ((cols / n_owners) % 256) != 0) return 0;
  if (n_parts < 1) return 0;
  const int owner_rows = cols / n_owners;
  if (n_parts > 1 && ((owner_rows % n_parts) != 0 || ((owner_rows / n_parts) % 256) != 0)) return 0;
  const FusedPlan f = fused_plan(rows, cols, D, n_parts > 1 ? owner_rows / n_parts : 0, n_owners > 1);
  const TileCfg t = pick_tile(rows, cols);
  return f.ok && t.BN == 256 && t.cg == 2 ? 1 : 0;  // the forward's E-storing epilogue exists for pair tiles only
}

// Debug: where the MMG_FUSED_TRACE=1 timeline of the last fused launch of this shape lives inside the workspace.
int tc_fused_trace_region(int rows, int cols, int D, size_t* off, size_t* bytes, int* per_role, int* roles) {
  const FusedPlan f = fused_plan(rows, cols, D);
  if (!f.ok || f.trace_bytes == 0) return 0;
  *off = f.trace_off; *bytes = f.trace_bytes; *per_role = kTraceCap; *roles = kTraceRoles;
  return 1;
}

size_t tc_infonce_bwd_fused_workspace(int rows, int cols, int D) {
  const FusedPlan f = fused_plan(rows, cols, D);
  return f.ok ? f.total_bytes : 0;
}

int tc_infonce_bwd_fused(const void* a_hat, const void* b_hat, int rows, int cols, int D, int diag_offset,
                         const float* scale, const float* rinv, const float* cinv, const float* scal, float* dA,
                         float* const* dB_owners, int n_owners, int n_parts, int part, float* dlogscale_acc,
                         void* workspace, size_t workspace_bytes, cudaStream_t st, int* used, const void* e_stored,
                         long long lde, int emb_f16) {
  *used = 0;
  if (e_stored != nullptr && (dlogscale_acc != nullptr || (reinterpret_cast<uintptr_t>(e_stored) & 15) != 0 ||
                              (lde & 7) != 0 || lde < cols))
    return 0;  // stored-E mode has no sum g*cos (the cosines are gone) and needs 16-byte aligned rows
  if (n_owners < 1 || n_owners > kMaxOwners || (cols % n_owners) != 0 || ((cols / n_owners) % 256) != 0) return 0;
  if (n_parts < 1 || part < 0 || part >= n_parts) return 0;
  const int owner_rows = cols / n_owners;
  if (n_parts > 1 && ((owner_rows % n_parts) != 0 || ((owner_rows / n_parts) % 256) != 0)) return 0;
  const FusedPlan f = fused_plan(rows, cols, D, n_parts > 1 ? owner_rows / n_parts : 0, n_owners > 1);  // parts = whole column blocks
  if (!f.ok || workspace_bytes < f.total_bytes) return 0;
  if (!out_tma_ok(dA, D)) return 0;
  for (int i = 0; i < n_owners; ++i)
    if (dB_owners[i] == nullptr || !out_tma_ok(dB_owners[i], D)) return 0;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return 0;
  constexpr int BN = 256;  // accumulator tile width (two TMEM stages)
  BwdFusedParams p;
  fill_schedule(&p, f, rows, cols, D, n_owners, n_parts, part, BN);
  p.E = e_stored; p.ldE = lde; p.G = workspace;
  p.trace = nullptr;
  if (f.trace_bytes > 0) {
    p.trace = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + f.trace_off);
    cudaError_t te = cudaMemsetAsync(p.trace, 0, f.trace_bytes, st);
    if (te != cudaSuccess) return check_cuda(te, "cudaMemsetAsync(fused backward trace)");
  }
  p.diag_offset = diag_offset;
  p.emb_f16 = emb_f16 ? 1 : 0;
  // several owners (row-sharded run): start the column walk at this rank's own columns (BwdFusedParams::global_cb)
  p.col_rot = 0;
  if (n_owners > 1 && n_parts == 1 && tune_int(kTuneRot, 1) != 0 && diag_offset > 0 && (diag_offset % f.Cb) == 0)
    p.col_rot = (diag_offset / f.Cb) % p.nbc;
  p.rinv = rinv; p.cinv = cinv; p.scale = scale; p.scal = scal; p.dlogscale_acc = dlogscale_acc;
  unsigned int* ctr = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + f.g_bytes);
  p.doneA = ctr;
  p.doneB = ctr + p.nblk;
  cudaError_t e = cudaMemsetAsync(ctr, 0, (size_t)p.nblk * 2 * sizeof(unsigned int), st);
  if (e != cudaSuccess) return check_cuda(e, "cudaMemsetAsync(fused backward counters)");

  CUtensorMap mAk, mBk, mAmn, mBmn, mGk, mGmn, mGst, mdA;
  BwdOwnerMaps mdB;
  int rc;
  const long long grow = (long long)f.nbuf * f.Rb;
  if ((rc = make_tmap(&mAk, a_hat, D, rows, D, kBM)) != 0) return rc;
  if ((rc = make_tmap(&mBk, b_hat, D, cols, D, BN / 2)) != 0) return rc;
  if ((rc = make_tmap(&mAmn, a_hat, D, rows, D, kBK)) != 0) return rc;
  if ((rc = make_tmap(&mBmn, b_hat, D, cols, D, kBK)) != 0) return rc;
  if ((rc = make_tmap(&mGk, workspace, f.Cb, grow, f.Cb, kBM)) != 0) return rc;
  if ((rc = make_tmap(&mGmn, workspace, f.Cb, grow, f.Cb, kBK)) != 0) return rc;
  if ((rc = make_tmap(&mGst, workspace, f.Cb, grow, f.Cb, 32)) != 0) return rc;
  if ((rc = make_out_tmap_f32(&mdA, dA, D, rows, D)) != 0) return rc;
  for (int i = 0; i < kMaxOwners; ++i) {
    const int o = i < n_owners ? i : 0;  // unused slots repeat owner 0 (never selected)
    if ((rc = make_out_tmap_f32(&mdB.m[i], dB_owners[o], D, owner_rows / n_parts, D)) != 0) return rc;
  }

  constexpr int kStoredTW = 8;  // transform warps of the stored-E kernel
  constexpr int ew = 8;         // epilogue warps
  const bool stored = e_stored != nullptr;
  auto kern = infonce_bwd_fused_kernel<256, 8, 0>;
  if (stored) kern = infonce_bwd_fused_kernel<256, 8, kStoredTW>;
  const int smem_bytes = FusedSmemT<256, 8>::kTotal;
  int slot = stored ? 1 : 0;
#ifdef MMG_MEASURE
  if (p.trace != nullptr) {  // debug timeline: separate instantiations
    if (stored) kern = infonce_bwd_fused_kernel<256, 8, kStoredTW, true>;
    else kern = infonce_bwd_fused_kernel<256, 8, 0, true>;
    slot += 2;
  }
#else
  p.trace = nullptr;
#endif
  static bool configured[4] = {false, false, false, false};
  if (!configured[slot]) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(infonce_bwd_fused_kernel)");
    configured[slot] = true;
  }
  const int pairs = sm_count() / 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(pairs * 2);
  cfg.blockDim = dim3(32 * (4 + ew + (stored ? kStoredTW : 0)));
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  PdlAttr at;
  at.cluster(2);
  cfg.attrs = at.a;
  cfg.numAttrs = at.n;
  // The kernel's CTAs wait for each other (doneA / doneB): every cluster of the grid must be resident at once.  Ask the
  // occupancy calculator (it accounts for MPS / green-context SM limits and the per-SM shared memory) once per device and
  // kernel variant; with fewer cluster slots than CTA pairs the caller's block loop runs instead.
  {
    static int max_clusters[64][4];
    static bool known[64][4];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!known[dev][slot]) {
      int n = 0;
      cudaError_t oe = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      if (oe != cudaSuccess) {
        cudaGetLastError();
        n = pairs;  // the query is unavailable: keep the one-CTA-per-SM assumption
      }
      max_clusters[dev][slot] = n;
      known[dev][slot] = true;
    }
    if (max_clusters[dev][slot] < pairs) return 0;  // *used stays 0
  }
  // (An L2 persisting access-policy window on the coefficient scratch was tried and is much slower -- 4.1 vs 2.65 ms at
  // 32768^2 -- and the device-wide set-aside also slows every later kernel of the process; not used.)
  e = cudaLaunchKernelEx(&cfg, kern, mAk, mBk, mAmn, mBmn, mGk, mGmn, mGst, mdA, mdB, p);
  if (e != cudaSuccess) return check_cuda(e, "infonce_bwd_fused_kernel launch");
  count_launch();
  *used = 1;
  return 0;
}

}  // namespace mmg
