// mmgclip_b200 -- the whole InfoNCE backward as ONE persistent launch.
//
// The block-by-block backward (c_api.cu: one coefficient launch + one gradient-GEMM launch per logit block) pays a
// launch / pipeline-fill / wave-quantisation cost per launch (measured ~6-10 us x 32 launches at B = 32768) and sends
// the coefficient block g through HBM.  This kernel keeps the same three contractions per block
//
//     (A)  g[Rb x Cb]   = EpiGrad( a[rows of the block] . b[cols of the block]^T )        K = D      -> bf16 scratch
//     (I)  dA[rows]    += g . b[cols]                                                     K = Cb     (slices of kslI x 64)
//     (T)  dB[cols]    += g^T . a[rows]                                                   K = Rb     (slices of kslT x 64)
//
// but runs them as a stream of work items over a persistent grid (one CTA pair per two SMs, the tcgen05 mainloop of
// gemm_tc.cuh).  Items of different blocks overlap: the coefficient tiles of block s are interleaved with the gradient
// slices of block s-2, so the MUFU-heavy coefficient epilogues hide behind the long-K gradient MMAs, the small
// coefficient blocks (nbuf x Rb x Cb bf16, default 4 x 16 MiB) stay in L2, and there is one prologue and one tail.
//
// Ordering.  Every item has a key (group, position); item lists are merged by key.  Coefficient tiles of block s have
// group s, gradient slices of block s have group s + 2.  CTA pair p owns the coefficient tiles with index = p mod P
// and the gradient slices with index = p mod P and walks them in key order; all three roles of a CTA (TMA producer,
// MMA issuer, epilogue warps) enumerate the same list independently.  Dependencies go through two global counters per
// block:
//     doneA[s] : (CTA, coefficient tile) pairs of block s whose stores are complete   (gradient loads of s wait for all)
//     doneB[s] : gradient slices of block s whose MMAs have completed               (block s + nbuf may then overwrite
//                                                                                    the scratch buffer)
// Stored-E mode (template flag kStoredE).  When the forward kept E = exp(logit - s) (bf16, rows x cols) the coefficient
// tiles need no tensor-core work at all: g = E * (rinv[r] + cinv[c]) is a pure streaming transform.  (A) items are then
// skipped by the TMA producer, the MMA issuer and the epilogue warps (they take no operand-ring slot and no accumulator
// stage); kTW extra *transform warps* per CTA read their 16 rows x 256 columns of E straight from global memory (eight
// independent 16-byte loads per thread in flight), scale them and store the bf16 coefficients into the scratch buffer,
// meet at a named barrier and publish doneA.  The
// backward executes 4 B^2 D FLOPs instead of 6 B^2 D, at the price of 2 B^2 bytes of HBM.  (Timeline, MMG_FUSED_TRACE=1
// at 32768^2 x 512: the transform warps are busy 76 % of the launch, ~7.6 us per tile, and the producers still wait
// ~0.46 ms for doneA.  Prefetching the next tile's E rows into L2 from the transform warps measured no gain -- 2.32 vs
// 2.62 ms recompute on that box, the same ratio as without -- and was not kept.)
//
// An item only ever waits for items with a strictly smaller key and every pair processes its items in key order, so the
// unfinished item with the smallest key can always run: no dead-lock as long as all CTAs are resident (grid <= #SMs).
#pragma once

#include "gemm_tc.cuh"

namespace mmg {

struct BwdFusedParams {
  int rows, cols, D;   // local rows, all columns, embedding width
  int Rb, Cb;          // block shape (multiples of 256 that divide rows / cols)
  int nbc, nblk;       // column blocks per block row; blocks in total (row-major over the block grid)
  int nbuf;            // coefficient scratch buffers
  int tAm, tAn, tDn;   // Rb/256, Cb/BN, D/BN
  int kslI, kslT;      // 64-wide K blocks per dA slice / per dB slice
  int sI, sT;          // slices per dA tile (Cb/64/kslI) / per dB tile (Rb/64/kslT)
  int nA, nBI, nB;     // items per block: coefficient tiles; dA slices; all gradient slices
  int diag_offset;     // global column paired with local row 0
  int emb_f16;         // 1: embedding operands AND coefficient scratch are fp16 (kind::f16 cannot mix the two 16-bit
                       // formats in one MMA: a bf16 x fp16 instruction traps); 0: all bf16
  const float* rinv;
  const float* cinv;
  const float* scale;
  const float* scal;
  float* dlogscale_acc;
  unsigned int* doneA;  // [nblk], zeroed by the host before the launch
  unsigned int* doneB;  // [nblk]
  int owner_rows;       // column-side gradient rows per owner: column c accumulates into dB map c / owner_rows at row
                        // c % owner_rows - part_row0 (one owner = the whole of dB on a single GPU; in the row-sharded
                        // run every rank owns cols / world rows)
  // Column parts: a launch may cover only the i-th of n equal parts of EVERY owner's columns (so that the reduce-scatter
  // of one part can travel while the next part is computed).  Block column index cb of this launch -> global block:
  int blocks_per_owner;   // column blocks per owner in the whole problem (owner_rows / Cb)
  int blocks_per_part;    // blocks_per_owner / n_parts
  int part;               // which part this launch covers
  int part_row0;          // part * blocks_per_part * Cb: first row of the part inside an owner's gradient rows
  // stored-E mode (kStoredE): E = exp(logit - s) of the local rows x all columns, bf16 row-major, written by the forward
  // (EpiLseT<.., true>); the coefficient tiles are then a transform of E instead of a recomputation of the cosines
  const void* E;
  long long ldE;
  void* G;                // coefficient scratch [nbuf * Rb, Cb] bf16 (the stored-E transform writes it with plain stores)
  // Debug timeline (MMG_FUSED_TRACE=1; NULL otherwise): per CTA and role kTraceCap records {globaltimer ns, tag}, see
  // trace_event().  Read back by tests/gpu_stored_e_probe.py trace.
  unsigned long long* trace;
  // Block order.  Block index blk (the order the blocks are walked in) -> (row block, column block of this launch).
  // Blocks are visited super-tile by super-tile (sr x sc blocks, row-major inside and across super-tiles) so that the
  // gradient rows a super-tile accumulates into and its operand rows can stay in L2.  sr = sc = 1 (the default: measured
  // no better with larger super-tiles, see fill_schedule) is plain row-major.
  int sr, sc;
  __host__ __device__ void block_rc(int blk, int& rb, int& cb) const {
    const int per = sr * sc;
    const int sidx = blk / per, w = blk - sidx * per;
    const int scols = nbc / sc;
    const int srow = sidx / scols, scol = sidx - srow * scols;
    const int wr = w / sc;
    rb = srow * sr + wr;
    cb = scol * sc + (w - wr * sc);
  }
  // Rank-rotated column order (several owners, one part): the walk starts at this rank's OWN columns and wraps around, so
  // that at any moment the ranks of a node reduce-add into eight DIFFERENT owners (in lockstep order every rank would
  // hit owner 0, then owner 1, ...: seven senders into one NVLink ingress while the other seven idle).
  int col_rot;            // column blocks to rotate by (0 on one GPU)
  __host__ __device__ int global_cb(int cb) const {
    if (col_rot != 0) {
      const int g = cb + col_rot;
      return g >= nbc ? g - nbc : g;
    }
    const int owner = cb / blocks_per_part;
    return owner * blocks_per_owner + part * blocks_per_part + (cb - owner * blocks_per_part);
  }
};

constexpr int kTraceCap = 1024;   // records per (CTA, role)
constexpr int kTraceRoles = 4;    // 0 TMA producer, 1 MMA issuer, 2 epilogue warp 0, 3 transform warp 0

// tag = role << 60 | event << 56 | item type << 52 | block << 32 | tm << 16 | tn;  events: 0 item picked up, 1 its
// dependency wait is over, 2 item finished (producer: loads issued; MMA: last commit issued; epilogue: stores issued;
// transform: published)
__device__ __forceinline__ void trace_event(unsigned long long* trace, int role, int& n, int event, int type, int blk,
                                            int tm, int tn) {
  if (trace == nullptr || n >= kTraceCap) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  unsigned long long* rec = trace + (static_cast<size_t>(blockIdx.x * kTraceRoles + role) * kTraceCap + n) * 2;
  rec[0] = t;
  rec[1] = (static_cast<unsigned long long>(role) << 60) | (static_cast<unsigned long long>(event) << 56) |
           (static_cast<unsigned long long>(type) << 52) | (static_cast<unsigned long long>(blk & 0xfffff) << 32) |
           (static_cast<unsigned long long>(tm & 0xffff) << 16) | static_cast<unsigned long long>(tn & 0xffff);
  ++n;
}

constexpr int kMaxOwners = 8;
struct BwdOwnerMaps {
  CUtensorMap m[kMaxOwners];  // fp32 [owner_rows / n_parts, D] output maps (box 32 x 32), one per owner
};

struct BwdItem {
  int type;      // 0 coefficient tile, 1 dA slice, 2 dB slice
  int blk;
  int tm, tn;
  int kb0, nkb;
};

// Merged, key-ordered walk over the items one CTA pair owns.
struct BwdCursor {
  int a, b, na_total, nb_total, step;
  __host__ __device__ __forceinline__ void init(const BwdFusedParams& p, int pair, int pairs) {
    a = pair; b = pair; step = pairs;
    na_total = p.nblk * p.nA;
    nb_total = p.nblk * p.nB;
  }
  __host__ __device__ __forceinline__ bool next(const BwdFusedParams& p, BwdItem& it) {
    const bool have_a = a < na_total, have_b = b < nb_total;
    if (!have_a && !have_b) return false;
    bool take_a;
    if (!have_b) take_a = true;
    else if (!have_a) take_a = false;
    else {
      const int sa = a / p.nA, ja = a - sa * p.nA;
      const int sb = b / p.nB, jb = b - sb * p.nB;
      const int ga = sa, gb = sb + 2;
      take_a = (ga != gb) ? (ga < gb) : ((2 * ja + 1) * p.nB <= (2 * jb + 1) * p.nA);
    }
    if (take_a) {
      it.type = 0;
      it.blk = a / p.nA;
      const int j = a - it.blk * p.nA;
      it.tm = j / p.tAn;
      it.tn = j - it.tm * p.tAn;
      it.kb0 = 0;
      it.nkb = p.D / kBK;
      a += step;
    } else {
      it.blk = b / p.nB;
      int j = b - it.blk * p.nB;
      int spt, ksl;  // slices per tile, K blocks per slice
      if (j < p.nBI) { it.type = 1; spt = p.sI; ksl = p.kslI; }
      else { it.type = 2; j -= p.nBI; spt = p.sT; ksl = p.kslT; }
      const int tile = j / spt;
      const int q = j - tile * spt;
      it.tm = tile / p.tDn;
      it.tn = tile - it.tm * p.tDn;
      it.kb0 = q * ksl;
      it.nkb = ksl;
      b += step;
    }
    return true;
  }
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned int atom_acq_rel_cta_add(unsigned int* smem_ctr, unsigned int v) {
  unsigned int old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(smem_ctr)), "r"(v) : "memory");
  return old;
}

// A coefficient tile is published (doneA) once per CTA: every epilogue warp waits for its own bulk stores, then bumps a
// shared-memory counter; whoever arrives last (acq_rel at CTA scope chains the others' completed stores in) pays for the
// one gpu-scope release.  Nobody waits for anybody.  Eight slots: warps are never more than three tiles apart (at most
// four TMEM accumulator stages), so a slot is never reused before all eight arrivals of its previous use.
template <int kEW>
__device__ __forceinline__ void publish_tile(unsigned int* pub_cnt, unsigned int seq, unsigned int* done_ctr, int lane) {
  if (lane == 0) {
    tma_store_wait_all();
    const unsigned int old = atom_acq_rel_cta_add(pub_cnt + (seq & 7u), 1u);
    if ((old & (kEW - 1)) == kEW - 1) {  // last of the CTA's epilogue warps
      fence_proxy_async_all();
      red_release_gpu_add(done_ctr, 1u);
    }
  }
  __syncwarp();
}

// one lane spins until *ctr >= want (with the same hang guard as mbar_wait), then the warp re-converges
__device__ __forceinline__ void wait_counter(const unsigned int* ctr, unsigned int want, int lane) {
  if (lane == 0) {
    if (ld_acquire_gpu(ctr) < want) {
      const long long t0 = clock64();
      while (ld_acquire_gpu(ctr) < want) {
        __nanosleep(64);
        if (clock64() - t0 > MMG_HANG_GUARD_CYCLES) {
          printf("[mmgclip_b200] fused backward: counter wait timed out (block %d thread %d want %u have %u)\n",
                 (int)blockIdx.x, (int)threadIdx.x, want, ld_acquire_gpu(ctr));
          __trap();
        }
      }
    }
    fence_proxy_async_all();  // the acquired data is touched next by TMA (async proxy)
  }
  __syncwarp();
}

// BN = columns of a work item's accumulator tile (256 rows x BN): 256 -> two TMEM accumulator stages, 128 -> four.  With
// four the MMA issuer could run three items ahead of the epilogue warps and absorb the bursts of coefficient tiles (their
// epilogue lasts ~2x their MMAs) -- but a 256 x 128 x 16 MMA reads 8 KB of operands per SM every 64 clocks, i.e. all of
// the 128 B/clk of shared-memory bandwidth, before TMA writes and epilogue traffic: measured 5.9 vs 2.6 ms at 32768^2.
// MMG_FUSED_BN=128 keeps the variant reachable; everything uses 256.
// kEW = epilogue warps per CTA (8 or 16).  With 16 a coefficient tile's epilogue takes about as long as its MMAs, but
// the staging boxes (16 x 4 KB) cost one operand-ring stage and every thread is held to 96 registers: measured 2.65 vs
// 2.61 ms at 32768^2 and 0.358 vs 0.353 ms at 4096 x 32768, i.e. no gain -- MMG_FUSED_EPI_WARPS=16 keeps it reachable.
template <int BN, int kEW>
using FusedSmemT = GemmSmem<BN, 2, kEW * 4096, EpiGradT<kEW>::kScratchBytes>;

// Stored-E coefficient tile, one transform warp's share: kRows rows x 256 columns of the CTA's 128 x 256 half tile.  A warp
// load covers one whole row (32 lanes x 16 bytes = 256 bf16); eight independent 16-byte loads per thread are in flight
// before the first use.  g = E * (rinv[row] + cinv[col]); the matching pair gets the EpiGrad treatment.
template <int kRows>
__device__ __forceinline__ void stored_e_rows(const BwdFusedParams& p, const uint8_t* e_rows, long long e_pitch,
                                              uint8_t* g_rows, long long g_pitch, const float* rinv_rows,
                                              const float* cinv_cols, int row_g0, int col_g0, int lane) {
  // e_rows / g_rows: first row of the share at the tile's first column; pitches in bytes
  // row_g0: global column paired with the share's first row; col_g0: global column of the tile's first column
  static_assert(kRows % 8 == 0, "rows are processed eight at a time");
  float cv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cv[j] = __ldg(cinv_cols + lane * 8 + j);
  const bool has_diag = (row_g0 < col_g0 + 256) && (row_g0 + kRows > col_g0);  // warp-uniform
  float dcoef = 0.f;
  bool zero_diag = false;
  if (has_diag) {
    dcoef = __ldg(p.scal);
    zero_diag = __ldg(p.scal + 2) != 0.f;
  }
#pragma unroll 1
  for (int r0 = 0; r0 < kRows; r0 += 8) {
    uint4 ev[8];
    float ri[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      ev[i] = __ldcs(reinterpret_cast<const uint4*>(e_rows + (r0 + i) * e_pitch + lane * 16));
      ri[i] = __ldg(rinv_rows + r0 + i);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t w[4] = {ev[i].x, ev[i].y, ev[i].z, ev[i].w};
      float g[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        g[2 * k] = __uint_as_float(w[k] << 16) * (ri[i] + cv[2 * k]);
        g[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u) * (ri[i] + cv[2 * k + 1]);
      }
      if (has_diag) {
        const int dj = row_g0 + r0 + i - (col_g0 + lane * 8);  // index of the matching column among this thread's 8
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j == dj) g[j] = zero_diag ? 0.f : g[j] - dcoef;
      }
      uint4 o;
      o.x = pack_bf16x2(g[0], g[1]);
      o.y = pack_bf16x2(g[2], g[3]);
      o.z = pack_bf16x2(g[4], g[5]);
      o.w = pack_bf16x2(g[6], g[7]);
      *reinterpret_cast<uint4*>(g_rows + (r0 + i) * g_pitch + lane * 16) = o;
    }
  }
}

// kTW = transform warps of the stored-E mode (0 = recompute mode): warps 4 + kEW .. 4 + kEW + kTW - 1.
// kTrace compiles the debug timeline in (measurement builds only, -DMMG_MEASURE + MMG_FUSED_TRACE=1; the product library
// carries none of it).
// L2 eviction-priority hints (evict_last on the coefficient scratch stores / loads, evict_first on its last reader and on the
// streamed operands, hints on the reduce-add targets) were swept in round 2 (profiles/r02b_fused_hint_plan_sweep.log): equal
// at best, evict_last on the scratch loads 2.74 -> 3.17-3.85 ms -- every load keeps the default policy.
// Variants that were built, measured on hardware and deleted because they did not win (numbers in DESIGN.md s4.2): per-panel
// instead of per-block doneA counters (2.60 vs 2.61 ms at 32768^2, 0.352 vs 0.355 ms at 4096 x 32768; stored-E 2.51 vs
// 2.41 ms), a deferred publish of the stored-E tiles (2.51 vs 2.41 ms), sixteen transform warps (2.44 vs 2.41 ms), 128-column
// accumulator tiles (5.9 vs 2.6 ms) and sixteen epilogue warps (2.65 vs 2.61 ms).
template <int BN, int kEW, int kTW, bool kTrace = false>
__global__ void __launch_bounds__(32 * (4 + kEW + kTW), 1)
infonce_bwd_fused_kernel(const __grid_constant__ CUtensorMap mAk, const __grid_constant__ CUtensorMap mBk,
                         const __grid_constant__ CUtensorMap mAmn, const __grid_constant__ CUtensorMap mBmn,
                         const __grid_constant__ CUtensorMap mGk, const __grid_constant__ CUtensorMap mGmn,
                         const __grid_constant__ CUtensorMap mGst, const __grid_constant__ CUtensorMap mdA,
                         const __grid_constant__ BwdOwnerMaps mdB, const BwdFusedParams p) {
  constexpr bool kStoredE = kTW > 0;
  static_assert(!kStoredE || BN == 256, "the stored-E transform is written for 256-column tiles");
  using S = FusedSmemT<BN, kEW>;
  using FusedGrad = EpiGradT<kEW>;
  using FusedStore = EpiStoreF32T<kEW>;
  constexpr int kStages = S::kStages;
  constexpr int kAcc = 512 / BN;          // accumulator stages: all 512 TMEM columns
  constexpr uint32_t kTmemCols = 512;
  constexpr int kBHalf = BN / 2;          // B-operand rows each CTA of the pair stages

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("[mmgclip_b200] dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* smem = smem_raw;
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * S::kABytes;
  uint8_t* staging = sB + kStages * S::kBBytes;
  float* epi_scratch = reinterpret_cast<float*>(staging + kEW * 4096);
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kEW * 4096 + S::kScratch);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tfull_bar = bars + 2 * kStages;
  uint64_t* tempty_bar = bars + 2 * kStages + kAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAcc);
  unsigned int* pub_cnt = reinterpret_cast<unsigned int*>(bars) + 96;  // [8] arrival counters of the publish slots

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int pair = blockIdx.x >> 1;
  const int pairs = gridDim.x >> 1;

  cluster_sync_all();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mAk); tma_prefetch_desc(&mBk); tma_prefetch_desc(&mAmn); tma_prefetch_desc(&mBmn);
    tma_prefetch_desc(&mGk); tma_prefetch_desc(&mGmn);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kAcc; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEW * 2);
    }
    for (int i = 0; i < 8; ++i) pub_cnt[i] = 0u;
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_cg2(tmem_slot, kTmemCols);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_entry();  // set-up done; see gemm_tc_kernel

  BwdCursor cur;
  cur.init(p, pair, pairs);
  BwdItem it;
  const unsigned int wantA = static_cast<unsigned int>(p.nA) * 2u;  // one arrival per CTA per coefficient tile
  const unsigned int wantB = static_cast<unsigned int>(p.nB);

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    int verified = -1;  // blocks [0, verified] are known to have all their coefficient tiles in the scratch
    int ntr = 0;
    while (cur.next(p, it)) {
      if (kStoredE && it.type == 0) continue;  // no tensor-core work: the epilogue warps transform E
      int rb, cbl;
      p.block_rc(it.blk, rb, cbl);
      const int col0 = p.global_cb(cbl) * p.Cb;  // first global column of the block
      const int buf = it.blk % p.nbuf;
      if constexpr (kTrace) if (lane == 0) trace_event(p.trace, 0, ntr, 0, it.type, it.blk, it.tm, it.tn);
      if (it.type != 0 && it.blk > verified) {
        wait_counter(p.doneA + it.blk, wantA, lane);
        verified = it.blk;
      }
      if constexpr (kTrace) if (lane == 0) trace_event(p.trace, 0, ntr, 1, it.type, it.blk, it.tm, it.tn);
      const int half_off = static_cast<int>(cta_rank) * kBM;     // this CTA's 128 of the tile's 256 rows (A operand)
      const int n_half = static_cast<int>(cta_rank) * kBHalf;    // this CTA's half of the tile's BN columns (B operand)
      for (int kb = 0; kb < it.nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (S::kABytes + S::kBBytes));
          else mbar_arrive_cluster(&full_bar[stage], 0);
          uint8_t* a_dst = sA + stage * S::kABytes;
          uint8_t* b_dst = sB + stage * S::kBBytes;
          const int k0 = (it.kb0 + kb) * kBK;
          if (it.type == 0) {
            tma_load_2d_cg2(&mAk, &full_bar[stage], a_dst, k0, rb * p.Rb + it.tm * 256 + half_off, kEvictNormal);
            tma_load_2d_cg2(&mBk, &full_bar[stage], b_dst, k0, col0 + it.tn * BN + n_half, kEvictNormal);
          } else if (it.type == 1) {
            // dA[rows] += g . b[cols]:  A = g (K-major, K = block columns),  B = b (MN-major: [K = column index][N = D])
            tma_load_2d_cg2(&mGk, &full_bar[stage], a_dst, k0, buf * p.Rb + it.tm * 256 + half_off, kEvictNormal);
            const int n0 = it.tn * BN + n_half;
#pragma unroll
            for (int i = 0; i < kBHalf / 64; ++i)
              tma_load_2d_cg2(&mBmn, &full_bar[stage], b_dst + i * (kBK * 128), n0 + i * 64, col0 + k0, kEvictNormal);
          } else {
            // dB[cols] += g^T . a[rows]:  A = g (MN-major: [K = block row][M = block column]),  B = a (MN-major)
            const int m0 = it.tm * 256 + half_off;
            const int n0 = it.tn * BN + n_half;
#pragma unroll
            for (int i = 0; i < 2; ++i)
              tma_load_2d_cg2(&mGmn, &full_bar[stage], a_dst + i * (kBK * 128), m0 + i * 64, buf * p.Rb + k0, kEvictNormal);
#pragma unroll
            for (int i = 0; i < kBHalf / 64; ++i)
              tma_load_2d_cg2(&mAmn, &full_bar[stage], b_dst + i * (kBK * 128), n0 + i * 64, rb * p.Rb + k0, kEvictNormal);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if constexpr (kTrace) if (lane == 0) trace_event(p.trace, 0, ntr, 2, it.type, it.blk, it.tm, it.tn);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int n = 0;
      int ntr = 0;
      while (cur.next(p, it)) {
        if (kStoredE && it.type == 0) continue;
        const int acc_stage = n % kAcc;
        const uint32_t acc_phase = (n / kAcc) & 1;
        ++n;
        if constexpr (kTrace) if (lane == 0) trace_event(p.trace, 1, ntr, 0, it.type, it.blk, it.tm, it.tn);
        mbar_wait(&tempty_bar[acc_stage], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t a_mn = it.type == 2 ? 1u : 0u, b_mn = it.type != 0 ? 1u : 0u;
        const uint32_t emb = static_cast<uint32_t>(p.emb_f16);  // fp16 mode: embeddings AND coefficients are fp16
        const uint32_t idesc = make_idesc_16(256, BN, a_mn, b_mn, emb, emb);
        const uint32_t tmem_d = tmem_base + acc_stage * BN;
        const uint32_t a_lbo = a_mn ? kBK * 128 : 0, b_lbo = b_mn ? kBK * 128 : 0;
        const uint32_t a_kstep = a_mn ? kUmmaK * 128 : kUmmaK * 2;
        const uint32_t b_kstep = b_mn ? kUmmaK * 128 : kUmmaK * 2;
        for (int kb = 0; kb < it.nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if constexpr (kTrace) if (kb == 0 && lane == 0) trace_event(p.trace, 1, ntr, 1, it.type, it.blk, it.tm, it.tn);
          if (elect_one_sync()) {
            const uint32_t a_addr = smem_u32(sA + stage * S::kABytes);
            const uint32_t b_addr = smem_u32(sB + stage * S::kBBytes);
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              const uint64_t da = make_smem_desc_sw128(a_addr + k * a_kstep, a_lbo, 1024);
              const uint64_t db = make_smem_desc_sw128(b_addr + k * b_kstep, b_lbo, 1024);
              umma_bf16_ss_cg2(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_cg2(&empty_bar[stage]);
            if (kb == it.nkb - 1) umma_commit_cg2(&tfull_bar[acc_stage]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if constexpr (kTrace) if (lane == 0) trace_event(p.trace, 1, ntr, 2, it.type, it.blk, it.tm, it.tn);
      }
      // every outstanding accumulator stage has been released (the peer's remote arrivals have landed) before the
      // leader's barriers die
      for (int last = n - 1; last >= 0 && last > n - 1 - kAcc; --last)
        mbar_wait(&tempty_bar[last % kAcc], (last / kAcc) & 1);
    }
  } else if (warp >= 4 && warp < 4 + kEW) {
    // ===================== epilogue warps (both CTAs) =====================
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int ewarp = warp - 4;
    int n = 0;
    float carry = 0.f;
    int pending = -1;     // block whose coefficient-tile stores of this warp are not yet published in doneA
    unsigned int a_seq = 0;  // coefficient tiles this warp has finished (identical across the CTA's epilogue warps)
    int verified = -1;    // scratch buffers of blocks [0, verified + nbuf] are known to be free
    int ntr = 0;
    typename FusedGrad::Params gp;
    gp.scale_ptr = p.scale; gp.scal = p.scal; gp.dlogscale_acc = p.dlogscale_acc; gp.dbg = 0;
    gp.g_f16 = p.emb_f16;
    typename FusedStore::Params sp;
    sp.C = nullptr; sp.ldc = 0; sp.bias = nullptr; sp.alpha = 1.f; sp.alpha_ptr = p.scal + 3; sp.mode = 1; sp.relu = 0;
    sp.use_tma = 1;
    while (cur.next(p, it)) {
      if (pending >= 0) {
        // publish the previous coefficient tile (deferred to here so the stores' latency is off the critical path, and
        // done BEFORE blocking on the next accumulator so it never depends on this item's progress)
        publish_tile<kEW>(pub_cnt, a_seq++, p.doneA + pending, lane);
        pending = -1;
      }
      if (kStoredE && it.type == 0) continue;  // the transform warps own the coefficient tiles
      const int acc_stage = n % kAcc;
      const uint32_t acc_phase = (n / kAcc) & 1;
      ++n;
      int rb, cbl;
      p.block_rc(it.blk, rb, cbl);
      const int col0 = p.global_cb(cbl) * p.Cb;  // first global column of the block
      const int buf = it.blk % p.nbuf;
      if constexpr (kTrace) if (ewarp == 0 && lane == 0) trace_event(p.trace, 2, ntr, 0, it.type, it.blk, it.tm, it.tn);
      mbar_wait(&tfull_bar[acc_stage], acc_phase);
      tcgen05_fence_after();
      if constexpr (kTrace) if (ewarp == 0 && lane == 0) trace_event(p.trace, 2, ntr, 1, it.type, it.blk, it.tm, it.tn);
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_stage * BN;
      const int half_off = static_cast<int>(cta_rank) * kBM;
      if (it.type == 0) {
        if (it.blk >= p.nbuf && it.blk - p.nbuf > verified) {
          wait_counter(p.doneB + (it.blk - p.nbuf), wantB, lane);  // the buffer's previous block has been consumed
          verified = it.blk - p.nbuf;
        }
        gp.rinv = p.rinv + rb * p.Rb;
        gp.cinv = p.cinv + col0;
        gp.diag_offset = rb * p.Rb + p.diag_offset - col0;
        gp.g_row_off = buf * p.Rb;
        FusedGrad::template run<BN>(gp, tacc, it.tm * 256 + half_off, it.tn * BN, p.Rb, p.Cb, half, q, lane, ewarp,
                           epi_scratch + acc_stage * BN + half * (BN / (kEW / 4)), &mGst, staging, 0, carry);
        pending = it.blk;
      } else {
        // all MMAs of this slice have completed => its TMA reads of the coefficient scratch are done
        if (ewarp == 0 && leader && lane == 0) red_release_gpu_add(p.doneB + it.blk, 1u);
        int m0 = (it.type == 1 ? rb * p.Rb : col0) + it.tm * 256 + half_off;
        const CUtensorMap* cmap = &mdA;
        if (it.type == 2) {
          // the owner of these 128 gradient rows (owner_rows is a multiple of 256: a tile never straddles owners); a
          // remote owner's buffer is reached by the same TMA reduce-add, over NVLink
          const int owner = m0 / p.owner_rows;
          m0 -= owner * p.owner_rows + p.part_row0;
          cmap = &mdB.m[owner];
        }
        FusedStore::template run_tma<BN>(sp, tacc, m0, it.tn * BN, p.D, half, q, lane, cmap, staging + ewarp * 4096);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tempty_bar[acc_stage]);
        else mbar_arrive_cluster(&tempty_bar[acc_stage], 0);
      }
      if constexpr (kTrace) if (ewarp == 0 && lane == 0) trace_event(p.trace, 2, ntr, 2, it.type, it.blk, it.tm, it.tn);
    }
    if (pending >= 0) publish_tile<kEW>(pub_cnt, a_seq++, p.doneA + pending, lane);
    FusedGrad::finish(gp, carry, lane);
  } else if (kStoredE && warp >= 4 + kEW) {
    // ===================== transform warps (stored-E mode, both CTAs) =====================
    constexpr int kRowsPerWarp = kBM / (kTW > 0 ? kTW : 1);
    const int tw = warp - (4 + kEW);
    int verified = -1;  // scratch buffers of blocks [0, verified + nbuf] are known to be free
    int ntr = 0;
    uint8_t* g_base = static_cast<uint8_t*>(p.G);
    while (cur.next(p, it)) {
      if (it.type != 0) continue;
      int rb, cbl;
      p.block_rc(it.blk, rb, cbl);
      const int col0 = p.global_cb(cbl) * p.Cb;
      const int buf = it.blk % p.nbuf;
      if constexpr (kTrace) if (tw == 0 && lane == 0) trace_event(p.trace, 3, ntr, 0, it.type, it.blk, it.tm, it.tn);
      if (it.blk >= p.nbuf && it.blk - p.nbuf > verified) {
        wait_counter(p.doneB + (it.blk - p.nbuf), wantB, lane);  // the buffer's previous block has been consumed
        verified = it.blk - p.nbuf;
      }
      if constexpr (kTrace) if (tw == 0 && lane == 0) trace_event(p.trace, 3, ntr, 1, it.type, it.blk, it.tm, it.tn);
      const int r0 = it.tm * 256 + static_cast<int>(cta_rank) * kBM + tw * kRowsPerWarp;  // first row inside the block
      const int c0 = it.tn * BN;                                                          // first column inside the block
      const long long grow0 = static_cast<long long>(rb) * p.Rb + r0;                     // local row
      stored_e_rows<kRowsPerWarp>(p, static_cast<const uint8_t*>(p.E) + (grow0 * p.ldE + col0 + c0) * 2, p.ldE * 2,
                                  g_base + ((static_cast<long long>(buf) * p.Rb + r0) * p.Cb + c0) * 2,
                                  static_cast<long long>(p.Cb) * 2, p.rinv + grow0, p.cinv + col0 + c0,
                                  static_cast<int>(grow0) + p.diag_offset, col0 + c0, lane);
      // all transform warps' stores -> one gpu-scope release per CTA (the consumers read the scratch through TMA)
      fence_proxy_async_all();
      asm volatile("bar.sync 1, %0;" ::"n"((kTW > 0 ? kTW : 1) * 32) : "memory");
      if (tw == 0 && lane == 0) {
        __threadfence();
        fence_proxy_async_all();
        red_release_gpu_add(p.doneA + it.blk, 1u);
        if constexpr (kTrace) trace_event(p.trace, 3, ntr, 2, it.type, it.blk, it.tm, it.tn);
      }
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc_cg2(tmem_base, kTmemCols);
  }
}

}  // namespace mmg
