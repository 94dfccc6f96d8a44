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
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace nsgp {

namespace {

constexpr int BM = 128;            // tile rows  (UMMA M)
constexpr int BK = 32;             // fp32 elements per K block = one 128-byte swizzle row
constexpr int UMMA_K = 8;          // tf32: 32 bytes per instruction
constexpr int kThreads = 192;      // 6 warps
constexpr int kEpiWarp0 = 2;
constexpr int kMaxChainBlocks = 32;   // K blocks accumulated in TMEM per work item

struct alignas(64) TcMaps {
  CUtensorMap a[2][kMaxTaps];   // [hi/lo][tap]
  CUtensorMap b[2][kMaxTaps];
};

struct TcOperand {
  int T, Cs, rows, br;            // taps, rows per tap, valid rows, rows per TMA box
  int nxc;                        // K blocks per staged row
  int tap_yoff[kMaxTaps];
  int tap_xoff[kMaxTaps];
};

struct TcParams {
  TcOperand A, B;
  float* out;
  int ld, n_cols;
  float alpha;
  int nkb;                        // K blocks in total
  int splits;                     // K splits
  int tiles_m, tiles_n, n_tiles;  // tile grid (Gram: n_tiles counts visited pairs)
  int same_operand;               // Gram: B is A
  int vec_red;                    // use red.global.add.v4.f32 in the epilogue
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar,
                                            int x, int y, int c) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(c)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b),
               "f"(c), "f"(d)
               : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor: 128-byte rows, 8-row
// swizzle atoms 1024 bytes apart (SBO), LBO unused (1), descriptor version 1.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::tf32, fp32 accumulate, both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// Gram tile enumeration over 128-row blocks and BN-column blocks: tile (rb, cb) is
// visited when its column range reaches the diagonal block or beyond.
template <int BN>
__device__ __forceinline__ bool gram_tile_needed(int rb, int cb) {
  return (cb + 1) * (BN / BM) - 1 >= rb;
}

struct TileCoord { int rb, cb; };

template <int BN, int EPI>
__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int tile) {
  TileCoord t;
  if (EPI == kEpiGramAtomic) {
    // row-major walk over the needed tiles; tiles_m <= 36 so the loop is short
    int rb = 0;
    for (;; ++rb) {
      int first = rb / (BN / BM);                 // first needed cb of this row
      int cnt = p.tiles_n - first;
      if (tile < cnt) { t.rb = rb; t.cb = first + tile; break; }
      tile -= cnt;
    }
  } else {
    t.rb = tile / p.tiles_n;
    t.cb = tile - t.rb * p.tiles_n;
  }
  return t;
}

// ---------------------------------------------------------------- the kernel
template <int BN, int EPI, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
contraction_tc_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stage][A_hi | A_lo | B_hi | B_lo] then barriers
  constexpr uint32_t kABytes = BM * BK * 4;       // 16 KB
  constexpr uint32_t kBBytes = BN * BK * 4;
  constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 2 * BN;          // double-buffered accumulator

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

  const int n_items = p.n_tiles * p.splits;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int split = item / p.n_tiles;
        const TileCoord tc = decode_tile<BN, EPI>(p, item - split * p.n_tiles);
        const int kb0 = (int)((long long)p.nkb * split / p.splits);
        const int kb1 = (int)((long long)p.nkb * (split + 1) / p.splits);
        const int r0 = tc.rb * BM, c0 = tc.cb * BN;
        const bool share = (EPI == kEpiGramAtomic) && p.same_operand && BN == BM && tc.rb == tc.cb;
        // bytes landed per stage (full boxes always count, OOB parts are zero-filled)
        int segs_a = 0, segs_b = 0;
        for (int r = r0; r < r0 + BM && r < p.A.rows; r += p.A.br) ++segs_a;
        if (!share)
          for (int r = c0; r < c0 + BN && r < p.B.rows; r += p.B.br) ++segs_b;
        const uint32_t tx = 2u * (uint32_t)(segs_a * p.A.br + segs_b * p.B.br) * BK * 4u;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], tx);
          const uint32_t sbase = smem_u32(smem + stage * kStageBytes);
          const int krow = kb / p.A.nxc, kx0 = (kb - krow * p.A.nxc) * BK;
          for (int s = 0; s < segs_a; ++s) {
            const int r = r0 + s * p.A.br;
            const int t = r / p.A.Cs, c = r - t * p.A.Cs;
            const uint32_t off = (uint32_t)(s * p.A.br) * (BK * 4);
            tma_load_3d(sbase + off, &maps.a[0][t], &full_bar[stage], kx0 + p.A.tap_xoff[t],
                        krow + p.A.tap_yoff[t], c);
            tma_load_3d(sbase + kABytes + off, &maps.a[1][t], &full_bar[stage],
                        kx0 + p.A.tap_xoff[t], krow + p.A.tap_yoff[t], c);
          }
          for (int s = 0; s < segs_b; ++s) {
            const int r = c0 + s * p.B.br;
            const int t = r / p.B.Cs, c = r - t * p.B.Cs;
            const uint32_t off = (uint32_t)(s * p.B.br) * (BK * 4);
            tma_load_3d(sbase + 2 * kABytes + off, &maps.b[0][t], &full_bar[stage],
                        kx0 + p.B.tap_xoff[t], krow + p.B.tap_yoff[t], c);
            tma_load_3d(sbase + 2 * kABytes + kBBytes + off, &maps.b[1][t], &full_bar[stage],
                        kx0 + p.B.tap_xoff[t], krow + p.B.tap_yoff[t], c);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
    uint32_t stage = 0, phase = 0;
    uint32_t local_item = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++local_item) {
      const int split = item / p.n_tiles;
      const TileCoord tc = decode_tile<BN, EPI>(p, item - split * p.n_tiles);
      const int kb0 = (int)((long long)p.nkb * split / p.splits);
      const int kb1 = (int)((long long)p.nkb * (split + 1) / p.splits);
      const bool share = (EPI == kEpiGramAtomic) && p.same_operand && BN == BM && tc.rb == tc.cb;
      const uint32_t buf = local_item & 1;
      const uint32_t use = local_item >> 1;
      mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);     // epilogue drained this buffer
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sbase = smem_u32(smem + stage * kStageBytes);
          const uint64_t a_hi = make_kmajor_sw128_desc(sbase);
          const uint64_t a_lo = make_kmajor_sw128_desc(sbase + kABytes);
          const uint64_t b_hi = share ? a_hi : make_kmajor_sw128_desc(sbase + 2 * kABytes);
          const uint64_t b_lo = share ? a_lo : make_kmajor_sw128_desc(sbase + 2 * kABytes + kBBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);   // +32 bytes per K step
            // small terms first, then the dominant one
            tc_mma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            tc_mma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
            tc_mma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
          }
          tc_commit(&empty_bar[stage]);                 // smem slot free when these retire
          if (kb == kb1 - 1) tc_commit(&tmem_full[buf]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (kb1 <= kb0 && lane == 0) tc_commit(&tmem_full[buf]);   // defensive: empty K range
      __syncwarp();
    }
  } else {
    // ============================ epilogue ============================
    const int quad = warp & 3;                         // TMEM lane quadrant of this warp
    uint32_t local_item = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++local_item) {
      const int split = item / p.n_tiles;
      const TileCoord tc = decode_tile<BN, EPI>(p, item - split * p.n_tiles);
      const int kb0 = (int)((long long)p.nkb * split / p.splits);
      const int kb1 = (int)((long long)p.nkb * (split + 1) / p.splits);
      const uint32_t buf = local_item & 1;
      const uint32_t use = local_item >> 1;
      mbar_wait(&tmem_full[buf], use & 1);
      tc_fence_after();
      const int row = tc.rb * BM + quad * 32 + lane;
      const bool row_ok = row < p.A.rows && kb1 > kb0;
      float* orow = p.out + (long long)row * p.ld;
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * BN + chunk * 32;
        tc_ld32(taddr, v);
        tc_wait_ld();
        const int col0 = tc.cb * BN + chunk * 32;
        // tiles strictly below the diagonal block of this row carry nothing needed
        const bool wanted = (EPI != kEpiGramAtomic) || (col0 + 31 >= tc.rb * BM);
        if (row_ok && col0 < p.n_cols && wanted) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int col = col0 + j;
            const float x0 = p.alpha * __uint_as_float(v[j]), x1 = p.alpha * __uint_as_float(v[j + 1]),
                        x2 = p.alpha * __uint_as_float(v[j + 2]), x3 = p.alpha * __uint_as_float(v[j + 3]);
            if (p.vec_red && col + 3 < p.n_cols) {
              red_add_v4(orow + col, x0, x1, x2, x3);
            } else {
              if (col < p.n_cols) atomicAdd(orow + col, x0);
              if (col + 1 < p.n_cols) atomicAdd(orow + col + 1, x1);
              if (col + 2 < p.n_cols) atomicAdd(orow + col + 2, x2);
              if (col + 3 < p.n_cols) atomicAdd(orow + col + 3, x3);
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

// One tensor map per (hi/lo, tap): dims (x: columns valid for this tap, y: plane
// rows, c: channel rows), box (32, 1, br).  Columns >= tap_ext are out of bounds
// and read as zero - that is what masks the K tail of a shifted window.
int encode_operand(const Operand& o, int br, CUtensorMap (*dst)[kMaxTaps]) {
  EncodeTiledFn enc = get_encode_fn();
  NSGP_REQUIRE(enc != nullptr, "tcgen05 engine: cuTensorMapEncodeTiled is unavailable");
  for (int hl = 0; hl < 2; ++hl)
    for (int t = 0; t < o.T; ++t) {
      const float* base = o.base + (long long)hl * o.hl_stride +
                          (long long)o.tap_plane[t] * o.plane_stride;
      NSGP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0,
                   "tcgen05 engine: operand base must be 16-byte aligned");
      cuuint64_t gdim[3] = {(cuuint64_t)o.tap_ext[t], (cuuint64_t)o.Hs, (cuuint64_t)o.Cs};
      cuuint64_t gstr[2] = {(cuuint64_t)o.Ws * 4, (cuuint64_t)o.Hs * o.Ws * 4};
      cuuint32_t box[3] = {(cuuint32_t)BK, 1u, (cuuint32_t)br};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = enc(&dst[hl][t], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstr,
                       box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      NSGP_REQUIRE(r == CUDA_SUCCESS,
                   "cuTensorMapEncodeTiled failed (%d): ext=%d Hs=%d Cs=%d Ws=%d br=%d", (int)r,
                   o.tap_ext[t], o.Hs, o.Cs, o.Ws, br);
    }
  return 0;
}

void fill_operand(const Operand& o, int br, TcOperand* d) {
  d->T = o.T; d->Cs = o.Cs; d->rows = o.rows; d->br = br;
  d->nxc = ceil_div(o.Kw, BK);
  for (int t = 0; t < o.T; ++t) { d->tap_yoff[t] = o.tap_yoff[t]; d->tap_xoff[t] = o.tap_xoff[t]; }
  if (getenv("NSGP_DBG_X0")) for (int t = 0; t < o.T; ++t) d->tap_xoff[t] = 0;
  if (getenv("NSGP_DBG_X4")) for (int t = 0; t < o.T; ++t) d->tap_xoff[t] *= 4;
  if (getenv("NSGP_DBG_Y0")) for (int t = 0; t < o.T; ++t) d->tap_yoff[t] = 0;
}

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

template <int BN, int EPI, int STAGES>
int launch_tc(const TcMaps& maps, const TcParams& p, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * (2 * BM * BK * 4 + 2 * BN * BK * 4) + 1024 + 256;
  static bool configured = false;
  auto kern = contraction_tc_kernel<BN, EPI, STAGES>;
  if (!configured) {
    NSGP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
    configured = true;
  }
  int items = p.n_tiles * p.splits;
  int grid = items < sm_count() ? items : sm_count();
  ProfScope prof(EPI == kEpiGramAtomic ? kProfGram : kProfGemm, stream);
  kern<<<grid, kThreads, smem, stream>>>(maps, p);
  NSGP_LAUNCHED();
  return 0;
}

}  // namespace

int contraction_tc(const ContractionArgs& a, cudaStream_t stream) {
  NSGP_REQUIRE(a.A.T >= 1 && a.A.T <= kMaxTaps && a.B.T >= 1 && a.B.T <= kMaxTaps,
               "tcgen05 engine: bad tap count");
  NSGP_REQUIRE(a.A.Ws % 4 == 0 && a.B.Ws % 4 == 0, "tcgen05 engine: row pitch must be 16-byte");
  NSGP_REQUIRE(a.A.Kh == a.B.Kh && ceil_div(a.A.Kw, BK) == ceil_div(a.B.Kw, BK),
               "contraction: operands disagree on K blocks");
  const int br_a = pick_box_rows(a.A), br_b = pick_box_rows(a.B);
  NSGP_REQUIRE(br_a > 0 && br_b > 0, "tcgen05 engine: operand rows per tap must be >= 8 "
               "(multiple of 8 when taps > 1)");
  if (a.A.rows == 0 || a.n_cols == 0) return 0;

  TcMaps maps;
  TcParams p{};
  int rc = encode_operand(a.A, br_a, maps.a);
  if (rc) return rc;
  const bool same = (a.epi == kEpiGramAtomic) && a.A.base == a.B.base &&
                    a.A.hl_stride == a.B.hl_stride && a.A.rows == a.B.rows;
  rc = encode_operand(a.B, br_b, maps.b);
  if (rc) return rc;
  fill_operand(a.A, br_a, &p.A);
  fill_operand(a.B, br_b, &p.B);
  p.out = a.out; p.ld = a.ld; p.n_cols = a.n_cols; p.alpha = a.alpha;
  p.nkb = k_blocks(a.A);
  p.same_operand = same ? 1 : 0;
  constexpr int BN = 128;
  p.tiles_m = ceil_div(a.A.rows, BM);
  p.tiles_n = ceil_div(a.n_cols, BN);
  if (p.nkb == 0) return 0;
  static const int vec_red = [] {
    const char* e = getenv("NSGP_VEC_RED");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  p.vec_red = (vec_red && a.ld % 4 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0) ? 1 : 0;
  if (a.epi == kEpiGramAtomic) {
    int n = 0;
    for (int rb = 0; rb < p.tiles_m; ++rb) n += p.tiles_n - rb / (BN / BM);
    p.n_tiles = n;
  } else {
    p.n_tiles = p.tiles_m * p.tiles_n;
  }
  // K splits: bound the in-TMEM accumulation chain (the tensor core accumulates
  // with truncation: ~2^-25.6 relative error per accumulate step, measured), then
  // fill the machine.  Partial tiles are red.add'ed, so splits need no workspace.
  int splits = ceil_div(p.nkb, kMaxChainBlocks);
  int fill = ceil_div(2 * sm_count(), p.n_tiles);
  int cap = p.nkb / 8 > 0 ? p.nkb / 8 : 1;
  if (fill > cap) fill = cap;
  if (splits < fill) splits = fill;
  p.splits = splits;
  if (a.epi == kEpiGramAtomic) return launch_tc<BN, kEpiGramAtomic, 3>(maps, p, stream);
  return launch_tc<BN, kEpiGemmRmw, 3>(maps, p, stream);
}

}  // namespace nsgp
