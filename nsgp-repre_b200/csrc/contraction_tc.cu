// tcgen05 / TMEM / TMA contraction engine (sm_100a), 3xTF32 with fp32 accumulation.
//
//   out (+)= A * B^T      A: rows_a x K, B: rows_b x K, both K-major "virtual"
//                         operands over staged hi/lo planes (common.cuh).
//
// Every operand element was pre-split into hi = rna_tf32(x), lo = rna_tf32(x - hi)
// so the tensor core's tf32 truncation of the shared-memory words is a no-op and
//   A*B^T ~= A_hi*B_hi + A_hi*B_lo + A_lo*B_hi        (error ~2^-21 relative)
// is accumulated in fp32 in TMEM.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer   (cp.async.bulk.tensor.3d, SWIZZLE_128B, mbarrier tx)
//   warp 1      MMA issuer     (tcgen05.mma kind::tf32, M=128, N=BN, K=8), TMEM owner
//   warps 2..5  epilogue       (tcgen05.ld 32x32b -> registers -> global)
// Three pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty double
// buffer (MMA <-> epilogue), static round-robin work list (tile, K-split).
//
// Work item = (output tile 128 x BN, K range).  For the Gram epilogue only tiles
// touching the upper block-triangle are visited, K is split across items and the
// tile is red.add'ed into the running sum; the K chain per item is bounded so the
// fp32 accumulation chain inside the tensor core stays short.  For the GEMM
// epilogue tiles are exclusive and the tile is read-modify-written
// (out += alpha * acc).
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>
#include <unordered_map>

#include "common.cuh"
#include "tc_common.cuh"

namespace nsgp {

using namespace tc;

namespace {

// ---------------------------------------------------------------- work items
constexpr int BN = 128;            // tile columns (UMMA N)
constexpr int STAGES = 3;
constexpr uint32_t kABytes = BM * BK * 4;       // 16 KB, one operand plane of a stage
constexpr uint32_t kStageBytes = 4 * kABytes;   // A_hi | A_lo | B_hi | B_lo
// TMEM: 2 buffers x (main accumulator | correction accumulator) x 128 columns.
// hi*hi goes to `main`, the two 2^-11-sized cross terms to `corr`: the long chain
// of large addends is a third as long, and the small terms do not get truncated at
// the scale of the large sum.  The epilogue adds them in fp32 (round-to-nearest).
constexpr uint32_t kTmemCols = 512;

struct Item {
  const TcMaps* maps;
  const TcParams* p;
  int rb, cb, kb0, kb1;
};

// Single problem: items are (K split, tile) pairs enumerated arithmetically; Gram
// problems visit the upper block-triangle only.  Group: the list is in global memory.
__device__ __forceinline__ Item fetch_item(const TcMaps* pmaps, const TcParams* pp,
                                           const TcProblem* gprobs, const TcItem* gitems,
                                           int idx) {
  Item it;
  if (gitems != nullptr) {
    const int4 a = *reinterpret_cast<const int4*>(&gitems[idx]);
    const int kb1 = gitems[idx].kb1;
    it.maps = &gprobs[a.x].maps;
    it.p = &gprobs[a.x].p;
    it.rb = a.y; it.cb = a.z; it.kb0 = a.w; it.kb1 = kb1;
    return it;
  }
  it.maps = pmaps;
  it.p = pp;
  const int split = idx / pp->n_tiles;
  int tile = idx - split * pp->n_tiles;
  if (pp->gram) {
    int rb = 0;
    for (;; ++rb) {                       // tiles_m <= 36: short walk
      const int cnt = pp->tiles_n - rb;
      if (tile < cnt) break;
      tile -= cnt;
    }
    it.rb = rb; it.cb = rb + tile;
  } else {
    it.rb = tile / pp->tiles_n;
    it.cb = tile - it.rb * pp->tiles_n;
  }
  it.kb0 = (int)((long long)pp->nkb * split / pp->splits);
  it.kb1 = (int)((long long)pp->nkb * (split + 1) / pp->splits);
  return it;
}

// ---------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(kThreads, 1)
contraction_tc_kernel(const __grid_constant__ TcMaps pmaps, const __grid_constant__ TcParams pp,
                      const TcProblem* __restrict__ gprobs, const TcItem* __restrict__ gitems,
                      int n_items) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_base_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
        const Item it = fetch_item(&pmaps, &pp, gprobs, gitems, idx);
        const TcParams& p = *it.p;
        const int r0 = it.rb * BM, c0 = it.cb * BN;
        const bool share = p.gram && p.same_operand && it.rb == it.cb;
        // bytes landed per stage (full boxes always count, OOB parts are zero-filled)
        int segs_a = 0, segs_b = 0;
        for (int r = r0; r < r0 + BM && r < p.A.rows; r += p.A.br) ++segs_a;
        if (!share)
          for (int r = c0; r < c0 + BN && r < p.B.rows; r += p.B.br) ++segs_b;
        const uint32_t tx = 2u * (uint32_t)(segs_a * p.A.br + segs_b * p.B.br) * BK * 4u;
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], tx);
          const uint32_t sbase = smem_u32(smem + stage * kStageBytes);
          const int kx0 = kb * BK;
          for (int s = 0; s < segs_a; ++s) {
            const int r = r0 + s * p.A.br;
            const int t = r / p.A.Cs, c = r - t * p.A.Cs;
            const uint32_t off = (uint32_t)(s * p.A.br) * (BK * 4);
            tma_load_2d(sbase + off, &it.maps->a[0][t], &full_bar[stage], kx0, c);
            tma_load_2d(sbase + kABytes + off, &it.maps->a[1][t], &full_bar[stage], kx0, c);
          }
          for (int s = 0; s < segs_b; ++s) {
            const int r = c0 + s * p.B.br;
            const int t = r / p.B.Cs, c = r - t * p.B.Cs;
            const uint32_t off = (uint32_t)(s * p.B.br) * (BK * 4);
            tma_load_2d(sbase + 2 * kABytes + off, &it.maps->b[0][t], &full_bar[stage], kx0, c);
            tma_load_2d(sbase + 3 * kABytes + off, &it.maps->b[1][t], &full_bar[stage], kx0, c);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    uint32_t stage = 0, phase = 0;
    uint32_t local_item = 0;
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++local_item) {
      const Item it = fetch_item(&pmaps, &pp, gprobs, gitems, idx);
      const TcParams& p = *it.p;
      const bool share = p.gram && p.same_operand && it.rb == it.cb;
      // ragged right edge: issue only as many columns as are valid (N in steps of 16)
      int n_valid = p.n_cols - it.cb * BN;
      if (n_valid > BN) n_valid = BN;
      const uint32_t idesc = make_idesc_tf32(BM, (n_valid + 15) & ~15);
      const uint32_t buf = local_item & 1;
      const uint32_t use = local_item >> 1;
      mbar_wait_warp(&tmem_empty[buf], (use & 1) ^ 1, lane);     // epilogue drained this buffer
      tc_fence_after();
      const uint32_t d_main = tmem_base + buf * (2 * BN);
      const uint32_t d_corr = d_main + BN;
      for (int kb = it.kb0; kb < it.kb1; ++kb) {
        mbar_wait_warp(&full_bar[stage], phase, lane);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sbase = smem_u32(smem + stage * kStageBytes);
          const uint64_t a_hi = make_kmajor_sw128_desc(sbase);
          const uint64_t a_lo = make_kmajor_sw128_desc(sbase + kABytes);
          const uint64_t b_hi = share ? a_hi : make_kmajor_sw128_desc(sbase + 2 * kABytes);
          const uint64_t b_lo = share ? a_lo : make_kmajor_sw128_desc(sbase + 3 * kABytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);   // +32 bytes per K step
            const uint32_t first = (kb > it.kb0 || k > 0) ? 1u : 0u;
            tc_mma_tf32(d_corr, a_lo + adv, b_hi + adv, idesc, first);
            tc_mma_tf32(d_corr, a_hi + adv, b_lo + adv, idesc, 1u);
            tc_mma_tf32(d_main, a_hi + adv, b_hi + adv, idesc, first);
          }
          tc_commit(&empty_bar[stage]);                 // smem slot free when these retire
          if (kb == it.kb1 - 1) tc_commit(&tmem_full[buf]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (it.kb1 <= it.kb0 && lane == 0) tc_commit(&tmem_full[buf]);   // defensive: empty K range
      __syncwarp();
    }
  } else {
    // ============================ epilogue ============================
    const int quad = warp & 3;                         // TMEM lane quadrant of this warp
    uint32_t local_item = 0;
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++local_item) {
      const Item it = fetch_item(&pmaps, &pp, gprobs, gitems, idx);
      const TcParams& p = *it.p;
      const uint32_t buf = local_item & 1;
      const uint32_t use = local_item >> 1;
      mbar_wait_warp(&tmem_full[buf], use & 1, lane, 200);
      tc_fence_after();
      const int row = it.rb * BM + quad * 32 + lane;
      const bool row_ok = row < p.A.rows && it.kb1 > it.kb0;
      float* orow = p.out + (long long)row * p.ld;
      const float alpha = p.alpha;
      const int n_cols = p.n_cols;
      const bool vec = p.vec_red != 0;
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        uint32_t v[32], w[32];
        const uint32_t taddr =
            tmem_base + ((uint32_t)(quad * 32) << 16) + buf * (2 * BN) + chunk * 32;
        tc_ld32(taddr, v);
        tc_ld32(taddr + BN, w);
        tc_wait_ld();
        const int col0 = it.cb * BN + chunk * 32;
        // Gram: chunks strictly below the diagonal block of this row carry nothing needed
        const bool wanted = !p.gram || (col0 + 31 >= it.rb * BM);
        if (row_ok && col0 < n_cols && wanted) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int col = col0 + j;
            const float x0 = alpha * (__uint_as_float(v[j]) + __uint_as_float(w[j]));
            const float x1 = alpha * (__uint_as_float(v[j + 1]) + __uint_as_float(w[j + 1]));
            const float x2 = alpha * (__uint_as_float(v[j + 2]) + __uint_as_float(w[j + 2]));
            const float x3 = alpha * (__uint_as_float(v[j + 3]) + __uint_as_float(w[j + 3]));
            if (vec && col + 3 < n_cols) {
              red_add_v4(orow + col, x0, x1, x2, x3);
            } else {
              if (col < n_cols) atomicAdd(orow + col, x0);
              if (col + 1 < n_cols) atomicAdd(orow + col + 1, x1);
              if (col + 2 < n_cols) atomicAdd(orow + col + 2, x2);
              if (col + 3 < n_cols) atomicAdd(orow + col + 3, x3);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(kTmemCols)
                 : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int pick_box_rows(const Operand& o) {
  if (o.T == 1) {
    int br = 128;
    while (br > o.Cs) br >>= 1;
    return br < 8 ? 0 : br;
  }
  for (int br = 128; br >= 8; br >>= 1)
    if (o.Cs % br == 0) return br;
  return 0;
}

// One 2-D tensor map per (hi/lo, tap): dims (k: K valid columns, c: rows of the
// tap), box (32, br).  Columns >= K are out of bounds and read as zero - that is
// what masks the K tail.
int encode_operand(const Operand& o, int br, CUtensorMap (*dst)[kMaxTaps]) {
  EncodeTiledFn enc = get_encode_fn();
  NSGP_REQUIRE(enc != nullptr, "tcgen05 engine: cuTensorMapEncodeTiled is unavailable");
  for (int hl = 0; hl < 2; ++hl)
    for (int t = 0; t < o.T; ++t) {
      const float* base = o.base + (long long)hl * o.hl_stride + o.tap_off[t];
      NSGP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0,
                   "tcgen05 engine: operand base must be 16-byte aligned");
      cuuint64_t gdim[2] = {(cuuint64_t)o.K, (cuuint64_t)o.Cs};
      cuuint64_t gstr[1] = {(cuuint64_t)o.row_pitch * 4};
      cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)br};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = enc(&dst[hl][t], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr,
                       box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      NSGP_REQUIRE(r == CUDA_SUCCESS,
                   "cuTensorMapEncodeTiled failed (%d): K=%d Cs=%d pitch=%lld br=%d", (int)r, o.K,
                   o.Cs, o.row_pitch, br);
    }
  return 0;
}

void fill_operand(const Operand& o, int br, TcOperand* d) {
  d->T = o.T; d->Cs = o.Cs; d->rows = o.rows; d->br = br;
}

constexpr size_t kSmemBytes = (size_t)STAGES * kStageBytes + 1024 + 256;

int configure_kernel() {
  static bool configured = false;
  if (!configured) {
    NSGP_CHECK_CUDA(cudaFuncSetAttribute(contraction_tc_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemBytes));
    configured = true;
  }
  return 0;
}

// Validates one problem, encodes its tensor maps and fills every field of the
// parameter block except `splits`.  Returns 1 for an empty problem (nothing to do).
int build_problem(const ContractionArgs& a, TcProblem* out) {
  NSGP_REQUIRE(a.A.T >= 1 && a.A.T <= kMaxTaps && a.B.T >= 1 && a.B.T <= kMaxTaps,
               "tcgen05 engine: bad tap count");
  NSGP_REQUIRE(a.A.row_pitch % 4 == 0 && a.B.row_pitch % 4 == 0,
               "tcgen05 engine: row pitch must be 16-byte");
  NSGP_REQUIRE(ceil_div(a.A.K, BK) == ceil_div(a.B.K, BK),
               "contraction: operands disagree on K blocks");
  const int br_a = pick_box_rows(a.A), br_b = pick_box_rows(a.B);
  NSGP_REQUIRE(br_a > 0 && br_b > 0, "tcgen05 engine: operand rows per tap must be >= 8 "
               "(multiple of 8 when taps > 1)");
  TcParams& p = out->p;
  p = TcParams{};
  p.nkb = k_blocks(a.A);
  if (a.A.rows == 0 || a.n_cols == 0 || p.nkb == 0) return 1;
  int rc = encode_operand(a.A, br_a, out->maps.a);
  if (rc) return rc;
  rc = encode_operand(a.B, br_b, out->maps.b);
  if (rc) return rc;
  fill_operand(a.A, br_a, &p.A);
  fill_operand(a.B, br_b, &p.B);
  p.out = a.out; p.ld = a.ld; p.n_cols = a.n_cols; p.alpha = a.alpha;
  p.gram = (a.epi == kEpiGramAtomic) ? 1 : 0;
  p.same_operand = (p.gram && a.A.base == a.B.base && a.A.hl_stride == a.B.hl_stride &&
                    a.A.rows == a.B.rows) ? 1 : 0;
  p.tiles_m = ceil_div(a.A.rows, BM);
  p.tiles_n = ceil_div(a.n_cols, BN);
  p.n_tiles = p.gram ? p.tiles_m * (p.tiles_m + 1) / 2 : p.tiles_m * p.tiles_n;
  static const int vec_red = [] {
    const char* e = getenv("NSGP_VEC_RED");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  p.vec_red = (vec_red && a.ld % 4 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0) ? 1 : 0;
  p.splits = 1;
  return 0;
}

}  // namespace

namespace tc {
int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
}  // namespace tc

int contraction_tc(const ContractionArgs& a, cudaStream_t stream) {
  TcProblem prob;
  int rc = build_problem(a, &prob);
  if (rc == 1) return 0;
  if (rc) return rc;
  TcParams& p = prob.p;
  NSGP_REQUIRE(!p.gram || p.tiles_m == p.tiles_n, "Gram: operands must have the same rows");
  // CTA-pair kernel (256x256 tiles): measured (scripts/bench_gram.py) to run at the same
  // power-capped rate as the single-CTA kernel once barrier polling is cheap, so its
  // extra tile padding never pays; kept selectable for experiments.
  static const int force_pair = [] {
    const char* e = getenv("NSGP_PAIR_KERNEL");       // 1 = use it for every Gram
    return e ? atoi(e) : 0;
  }();
  const bool pair = p.gram && p.same_operand && force_pair == 1;
  if (pair) {
    const int t2 = ceil_div(a.A.rows, 256);
    p.tiles_m = p.tiles_n = t2;
    p.n_tiles = t2 * (t2 + 1) / 2;
  }
  // K splits.  The tensor core accumulates with truncation (~2^-25.6 relative per
  // accumulate step, measured), so the chain per TMEM accumulator is bounded; beyond
  // that choose the split count that minimises  waves x (K blocks per item +
  // per-item overhead) - the last wave of the static round-robin deal is the tail.
  // Partial tiles are red.add'ed, so splits need no workspace.
  // GEMM (W += update @ P): one chain per tile up to kChainGemm blocks (d <= 5120), so
  // every W element takes ONE fp32 rounding like the reference's add_ and the result
  // is deterministic.
  const int workers = pair ? sm_count() / 2 : sm_count();
  {
    const int s_min = ceil_div(p.nkb, p.gram ? (pair ? 32 : kChainGram) : kChainGemm);
    int s_max = p.nkb / 4 > s_min ? p.nkb / 4 : s_min;
    if (s_max > s_min + 64) s_max = s_min + 64;
    if (!p.gram) s_max = s_min;
    const int overhead = 4;            // K-block equivalents of pipeline fill + epilogue
    long long best = -1;
    int best_s = s_min;
    for (int sp = s_min; sp <= s_max; ++sp) {
      long long waves = ceil_div((long long)p.n_tiles * sp, workers);
      long long cost = waves * (ceil_div(p.nkb, sp) + overhead);
      if (best < 0 || cost < best) { best = cost; best_s = sp; }
    }
    p.splits = best_s;
  }
  if (pair) return launch_tc2_gram(prob.maps, p, stream);
  rc = configure_kernel();
  if (rc) return rc;
  const int items = p.n_tiles * p.splits;
  const int grid = items < sm_count() ? items : sm_count();
  ProfScope prof(p.gram ? kProfGram : kProfGemm, stream);
  contraction_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(prob.maps, p, nullptr, nullptr,
                                                               items);
  NSGP_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------- grouped launches
namespace {
int chain_splits(const TcParams& p) {
  return ceil_div(p.nkb, p.gram ? kChainGram : kChainGemm);
}
}  // namespace

size_t group_table_bytes(const ContractionArgs* probs, int n) {
  size_t items = 0;
  for (int i = 0; i < n; ++i) {
    const int tm = ceil_div(probs[i].A.rows, BM), tn = ceil_div(probs[i].n_cols, BN);
    const int nkb = k_blocks(probs[i].A);
    const bool gram = probs[i].epi == kEpiGramAtomic;
    const size_t tiles = gram ? (size_t)tm * (tm + 1) / 2 : (size_t)tm * tn;
    items += tiles * (size_t)ceil_div(nkb > 0 ? nkb : 1, gram ? kChainGram : kChainGemm);
  }
  return (size_t)n * sizeof(TcProblem) + items * sizeof(TcItem) + 256;
}

int group_table_build(const ContractionArgs* probs, int n, int kind, void* table_dev,
                      size_t table_bytes, GroupInfo* info, cudaStream_t stream) {
  NSGP_REQUIRE(probs && table_dev && info && n >= 0, "group_build: bad arguments");
  NSGP_REQUIRE((reinterpret_cast<uintptr_t>(table_dev) & 63) == 0,
               "group_build: table must be 64-byte aligned");
  std::vector<TcProblem> hp;
  hp.reserve(n);
  struct Cost { long long c; TcItem it; };
  std::vector<Cost> items;
  for (int i = 0; i < n; ++i) {
    TcProblem pr;
    int rc = build_problem(probs[i], &pr);
    if (rc == 1) continue;
    if (rc) return rc;
    const TcParams& p = pr.p;
    NSGP_REQUIRE(!p.gram || p.tiles_m == p.tiles_n, "Gram: operands must have the same rows");
    const int prob = (int)hp.size();
    const int splits = chain_splits(p);
    for (int sp = 0; sp < splits; ++sp) {
      const int kb0 = (int)((long long)p.nkb * sp / splits);
      const int kb1 = (int)((long long)p.nkb * (sp + 1) / splits);
      for (int rb = 0; rb < p.tiles_m; ++rb)
        for (int cb = p.gram ? rb : 0; cb < p.tiles_n; ++cb) {
          Cost c;
          c.it = TcItem{prob, rb, cb, kb0, kb1, 0, 0, 0};
          // cost ~ K blocks (+ fixed part); diagonal Gram tiles load half the bytes
          c.c = (long long)(kb1 - kb0) * 8 + 16;
          items.push_back(c);
        }
    }
    hp.push_back(pr);
  }
  // big items first; ties keep (problem, K range, tile) order so that tiles sharing
  // operand rows of one K range run at the same time (L2 reuse)
  std::stable_sort(items.begin(), items.end(),
                   [](const Cost& x, const Cost& y) { return x.c > y.c; });
  info->n_problems = (int)hp.size();
  info->n_items = (int)items.size();
  info->kind = kind;
  info->off_items = hp.size() * sizeof(TcProblem);
  info->bytes = info->off_items + items.size() * sizeof(TcItem);
  NSGP_REQUIRE(info->bytes <= table_bytes, "group_build: table too small (%zu < %zu)",
               table_bytes, info->bytes);
  if (info->n_items == 0) return 0;
  std::vector<TcItem> hi(items.size());
  for (size_t i = 0; i < items.size(); ++i) hi[i] = items[i].it;
  NSGP_CHECK_CUDA(cudaMemcpyAsync(table_dev, hp.data(), info->off_items, cudaMemcpyHostToDevice,
                                  stream));
  NSGP_CHECK_CUDA(cudaMemcpyAsync((char*)table_dev + info->off_items, hi.data(),
                                  hi.size() * sizeof(TcItem), cudaMemcpyHostToDevice, stream));
  return 0;
}

int group_launch(const void* table_dev, const GroupInfo& info, cudaStream_t stream) {
  if (info.n_items == 0) return 0;
  int rc = configure_kernel();
  if (rc) return rc;
  const TcProblem* probs = reinterpret_cast<const TcProblem*>(table_dev);
  const TcItem* items =
      reinterpret_cast<const TcItem*>((const char*)table_dev + info.off_items);
  const int grid = info.n_items < sm_count() ? info.n_items : sm_count();
  static const TcMaps dummy_maps{};
  TcParams dummy{};
  ProfScope prof(info.kind, stream);
  contraction_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(dummy_maps, dummy, probs, items,
                                                               info.n_items);
  NSGP_LAUNCHED();
  return 0;
}

}  // namespace nsgp
