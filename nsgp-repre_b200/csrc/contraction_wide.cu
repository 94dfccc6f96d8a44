// Wide-tile Gram kernel (tcgen05 / TMEM / TMA, 3xTF32): 128 x 256 output tiles for the
// covariance problems that are plain Grams X X^T of one staged operand (1x1 convs, the
// stride-2 3x3 convs, the stem).
//
// The generic kernel (contraction_tc.cu) moves 64 KB of operands per 12 MMAs and its
// 3-stage ring is latency-bound (~1 600 cycles per K block against 768 of MMA work,
// DESIGN.md 4).  Here one A tile (128 rows) meets TWO B tiles (256 rows): 96 KB per 24
// MMA-equivalents (12 instructions with N = 256), i.e. 2/3 of the bytes per unit of tensor
// work; the price is that the two accumulators (hi*hi | cross terms, 256 columns each) fill
// the 512 TMEM columns, so the epilogue of an item does not overlap the next item's MMAs,
// and that only two 96 KB stages fit.  Measured: slower than the generic kernel (see
// gram_wide_enabled) - kept as an opt-in experiment (NSGP_WIDE_KERNEL=1).
//
//   warp 0   TMA producer of B  (2 x [hi | lo] tiles of 128 rows x 32 columns)
//   warp 6   TMA producer of A  (skipped when the A block is the first B block: diagonal)
//   warp 1   MMA issuer: per K block 12 x tcgen05.mma.kind::tf32 (M = 128, N <= 256, K = 8)
//   warps 2..5  epilogue: main + corr -> red.global.add.v4.f32 (upper block-triangle only)
//
// Work item = (problem, row block rb, first column block cb0, 1 or 2 column blocks, K range);
// the column blocks of a row are paired from the diagonal to the right, the K range bounds
// the accumulation chain like in the generic kernel.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "geometry.h"
#include "tc_common.cuh"

namespace nsgp {

using namespace tc;

namespace {

constexpr int WS = 2;                           // ring stages
constexpr uint32_t kTileB = BM * BK * 4;        // 16 KB: one hi or lo plane of 128 rows
constexpr uint32_t kStageW = 6 * kTileB;        // B_hi(2) | B_lo(2) | A_hi | A_lo = 96 KB
constexpr int kThreadsW = 224;
constexpr size_t kSmemW = (size_t)WS * kStageW + 1024 + 256;
constexpr int kChainW = 64;                     // K blocks per accumulator chain

struct alignas(64) WProblem {
  CUtensorMap maps[2][kMaxTaps];                // [hi/lo][tap]: (K, Cs) row-major, box (32, 128)
  float* out;
  int ld, rows, Cs, T;
  int nkb, pad0, pad1, pad2;
};
struct alignas(16) WItem {
  int prob, rb, cb0, ncb;
  int kb0, kb1, pad0, pad1;
};

__global__ void __launch_bounds__(kThreadsW, 1)
gram_wide_kernel(const WProblem* __restrict__ probs, const WItem* __restrict__ items, int n_items,
                 unsigned long long* tl) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + WS * kStageW);
  uint64_t* b_full = a_full + WS;
  uint64_t* empty = b_full + WS;
  uint64_t* tmem_full = empty + WS;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  tl_begin(tl);
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < WS; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&b_full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_base_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0 || warp == 6) {
    // ============================ TMA producers ============================
    const bool is_b = warp == 0;
    uint32_t cnt = 0;
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
      const WItem it = items[idx];
      const WProblem& p = probs[it.prob];
      const bool share = it.rb == it.cb0;
      for (int kb = it.kb0; kb < it.kb1; ++kb, ++cnt) {
        const uint32_t s = cnt % WS, par = (cnt / WS) & 1;
        mbar_wait_warp(&empty[s], par ^ 1, lane);
        const uint32_t sbase = smem_u32(smem + s * kStageW);
        if (is_b) {
          mbar_expect_tx_elect(&b_full[s], (uint32_t)it.ncb * 2u * kTileB);
          for (int j = 0; j < it.ncb; ++j) {
            const int r = (it.cb0 + j) * BM;
            const int t = r / p.Cs, c = r - t * p.Cs;
            tma_load_2d_elect(sbase + j * kTileB, &p.maps[0][t], &b_full[s], kb * BK, c);
            tma_load_2d_elect(sbase + (2 + j) * kTileB, &p.maps[1][t], &b_full[s], kb * BK, c);
          }
        } else if (!share) {
          mbar_expect_tx_elect(&a_full[s], 2u * kTileB);
          const int r = it.rb * BM;
          const int t = r / p.Cs, c = r - t * p.Cs;
          tma_load_2d_elect(sbase + 4 * kTileB, &p.maps[0][t], &a_full[s], kb * BK, c);
          tma_load_2d_elect(sbase + 5 * kTileB, &p.maps[1][t], &a_full[s], kb * BK, c);
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    uint32_t cnt = 0, n_done = 0;
    uint32_t a_cnt0 = 0, a_cnt1 = 0;            // A-full phases consumed per stage
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++n_done) {
      const WItem it = items[idx];
      const WProblem& p = probs[it.prob];
      const bool share = it.rb == it.cb0;
      int n_valid = p.rows - it.cb0 * BM;
      if (n_valid > it.ncb * BM) n_valid = it.ncb * BM;
      const uint32_t idesc = make_idesc_tf32(BM, (n_valid + 15) & ~15);
      mbar_wait_warp(tmem_empty, (n_done & 1) ^ 1, lane);        // epilogue drained TMEM
      tc_fence_after();
      const uint32_t d_main = tmem_base, d_corr = tmem_base + 256;
      for (int kb = it.kb0; kb < it.kb1; ++kb, ++cnt) {
        const uint32_t s = cnt % WS, par = (cnt / WS) & 1;
        mbar_wait_warp(&b_full[s], par, lane);
        if (!share) {
          // the A barrier of a stage only completes for items that load A: its own phase count
          uint32_t& ac = s == 0 ? a_cnt0 : a_cnt1;
          mbar_wait_warp(&a_full[s], ac & 1, lane);
          ++ac;
        }
        tc_fence_after();
        const uint32_t sbase = smem_u32(smem + s * kStageW);
        const uint64_t b_hi = make_kmajor_sw128_desc(sbase);
        const uint64_t b_lo = make_kmajor_sw128_desc(sbase + 2 * kTileB);
        const uint64_t a_hi = share ? b_hi : make_kmajor_sw128_desc(sbase + 4 * kTileB);
        const uint64_t a_lo = share ? b_lo : make_kmajor_sw128_desc(sbase + 5 * kTileB);
        tc_mma_kblock_3xtf32(d_main, d_corr, a_hi, a_lo, b_hi, b_lo, idesc,
                             kb > it.kb0 ? 1u : 0u, 0u);
        tc_commit_elect(&empty[s]);
        if (kb == it.kb1 - 1) tc_commit_elect(tmem_full);
      }
    }
  } else {
    // ============================ epilogue ============================
    const int quad = warp & 3;
    uint32_t n_done = 0;
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++n_done) {
      const WItem it = items[idx];
      const WProblem& p = probs[it.prob];
      mbar_wait_warp(tmem_full, n_done & 1, lane, 200);
      tc_fence_after();
      const int rblk = it.rb * BM;
      const int row = rblk + quad * 32 + lane;
      const bool row_ok = row < p.rows;
      float* orow = p.out + (long long)row * p.ld;
      const int n_cols = p.rows;
      const bool vec = (p.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
#pragma unroll 1
      for (int chunk = 0; chunk < it.ncb * (BM / 32); ++chunk) {
        uint32_t v[32], w[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + chunk * 32;
        tc_ld32(taddr, v);
        tc_ld32(taddr + 256, w);
        tc_wait_ld();
        const int col0 = it.cb0 * BM + chunk * 32;
        if (row_ok && col0 < n_cols && col0 + 31 >= rblk) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int col = col0 + j;
            const float x0 = __uint_as_float(v[j]) + __uint_as_float(w[j]),
                        x1 = __uint_as_float(v[j + 1]) + __uint_as_float(w[j + 1]),
                        x2 = __uint_as_float(v[j + 2]) + __uint_as_float(w[j + 2]),
                        x3 = __uint_as_float(v[j + 3]) + __uint_as_float(w[j + 3]);
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
      if (lane == 0) mbar_arrive(tmem_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  tl_end(tl);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(512)
                 : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn wide_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(p);
    return (EncodeTiledFn) nullptr;
  }();
  return fn;
}

int wide_splits(int nkb) { return ceil_div(nkb, kChainW); }

}  // namespace

bool gram_wide_enabled() {
  static const bool on = [] {
    // measured at configs[1] (scripts/bench_cov.py, NSGP_TIMELINE=1): parity-green, but the
    // 2-stage ring loses more than the wider tile gains - 1.53 ms + 0.16 ms of left-over
    // generic problems against 1.34 ms for the generic kernel alone -> opt-in
    const char* e = nsgp_env("NSGP_WIDE_KERNEL");
    return e && e[0] == '1';
  }();
  return on;
}

// A same-operand Gram the wide kernel takes: whole 128-row boxes per tap, row-major operand
bool gram_wide_eligible(const ContractionArgs& a) {
  if (!gram_wide_enabled() || a.epi != kEpiGramAtomic) return false;
  if (a.A.base != a.B.base || a.A.hl_stride != a.B.hl_stride || a.A.rows != a.B.rows) return false;
  if (a.A.tile_nkb != 0 || a.alpha != 1.f || a.A.row_pitch % 4 != 0) return false;
  if (a.A.T == 1) return a.A.Cs >= 128 && a.A.rows > 128;      // one block: nothing to pair
  return a.A.Cs % 128 == 0;
}

size_t gram_wide_table_bytes(const ContractionArgs* probs, int n) {
  size_t items = 0;
  for (int i = 0; i < n; ++i) {
    const int tb = ceil_div(probs[i].A.rows, BM);
    items += (size_t)wide_splits(k_blocks(probs[i].A)) * tb * (tb / 2 + 1);
  }
  return (size_t)n * sizeof(WProblem) + items * sizeof(WItem) + 1024;
}

int gram_wide_table_build(const ContractionArgs* probs, int n, void* table_dev, size_t table_bytes,
                          SubGroup* sg, cudaStream_t stream) {
  EncodeTiledFn enc = wide_encode_fn();
  NSGP_REQUIRE(enc != nullptr, "wide kernel: cuTensorMapEncodeTiled is unavailable");
  std::vector<WProblem> hp(n);
  struct Key { long long cost; int prob, split; WItem it; };
  std::vector<Key> keys;
  for (int i = 0; i < n; ++i) {
    const Operand& o = probs[i].A;
    WProblem& p = hp[i];
    memset(&p, 0, sizeof(p));
    p.out = probs[i].out; p.ld = probs[i].ld; p.rows = o.rows; p.Cs = o.Cs; p.T = o.T;
    p.nkb = k_blocks(o);
    for (int hl = 0; hl < 2; ++hl)
      for (int t = 0; t < o.T; ++t) {
        const float* base = o.base + (long long)hl * o.hl_stride + o.tap_off[t];
        NSGP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0,
                     "wide kernel: operand base must be 16-byte aligned");
        cuuint64_t gdim[2] = {(cuuint64_t)o.K, (cuuint64_t)o.Cs};
        cuuint64_t gstr[1] = {(cuuint64_t)o.row_pitch * 4};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&p.maps[hl][t], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim,
                         gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        NSGP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (wide) failed (%d)", (int)r);
      }
    const int tb = ceil_div(o.rows, BM), splits = wide_splits(p.nkb);
    for (int sp = 0; sp < splits; ++sp) {
      const int kb0 = (int)((long long)p.nkb * sp / splits);
      const int kb1 = (int)((long long)p.nkb * (sp + 1) / splits);
      for (int rb = 0; rb < tb; ++rb)
        for (int cb = rb; cb < tb;) {
          const int ncb = tb - cb >= 2 ? 2 : 1;
          Key k;
          k.cost = (long long)(kb1 - kb0) * (ncb == 2 ? 12 : 8) + 24;
          k.prob = i; k.split = sp;
          k.it = WItem{i, rb, cb, ncb, kb0, kb1, 0, 0};
          keys.push_back(k);
          cb += ncb;
        }
    }
  }
  // big items first; ties keep (problem, K range) order so that the tiles of one K range
  // run together and share the operand rows in L2
  std::stable_sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) {
    if (a.cost != b.cost) return a.cost > b.cost;
    if (a.prob != b.prob) return a.prob < b.prob;
    return a.split < b.split;
  });
  sg->n_problems = n;
  sg->n_items = (int)keys.size();
  sg->off_probs = 0;
  sg->off_items = (size_t)n * sizeof(WProblem);
  const size_t need = sg->off_items + keys.size() * sizeof(WItem);
  NSGP_REQUIRE(need <= table_bytes, "wide table too small (%zu < %zu)", table_bytes, need);
  NSGP_REQUIRE((reinterpret_cast<uintptr_t>(table_dev) & 63) == 0,
               "wide table must be 64-byte aligned");
  std::vector<WItem> hi(keys.size());
  for (size_t i = 0; i < keys.size(); ++i) hi[i] = keys[i].it;
  if (n > 0) {
    NSGP_CHECK_CUDA(cudaMemcpyAsync(table_dev, hp.data(), (size_t)n * sizeof(WProblem),
                                    cudaMemcpyHostToDevice, stream));
    NSGP_CHECK_CUDA(cudaMemcpyAsync((char*)table_dev + sg->off_items, hi.data(),
                                    hi.size() * sizeof(WItem), cudaMemcpyHostToDevice, stream));
  }
  return 0;
}

int gram_wide_launch(const void* table_dev, const SubGroup& sg, cudaStream_t stream) {
  if (sg.n_items == 0) return 0;
  static bool configured = false;
  if (!configured) {
    NSGP_CHECK_CUDA(cudaFuncSetAttribute(gram_wide_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemW));
    configured = true;
  }
  const WProblem* probs = reinterpret_cast<const WProblem*>((const char*)table_dev + sg.off_probs);
  const WItem* items = reinterpret_cast<const WItem*>((const char*)table_dev + sg.off_items);
  const int grid = sg.n_items < sm_count() ? sg.n_items : sm_count();
  ProfScope prof(kProfGram, stream);
  gram_wide_kernel<<<grid, kThreadsW, kSmemW, stream>>>(probs, items, sg.n_items,
                                                        timeline_slot(11));
  NSGP_LAUNCHED();
  return 0;
}

}  // namespace nsgp
