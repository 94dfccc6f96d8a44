// 2-CTA (cta_group::2) tcgen05 Gram kernel: a CTA pair on one TPC computes a 256 x 256
// output tile, 3xTF32, fp32 accumulation in TMEM.
//
// Why: with 128x128 tiles in SS mode every 64-cycle MMA reads 8 KB of shared memory
// (128 B/cycle, the whole smem bandwidth of the SM) while TMA writes another 85
// B/cycle - the tensor pipe stalls at ~55 % (profiles/ncu_gram_r01_summary.csv).
// In a pair each CTA owns 128 of the 256 tile rows (its A tile) and stages only half
// of the B tile; the MMA unit reads the other half from the peer.  Per SM that is
// 8 KB per 128-cycle MMA (64 B/cycle) plus 42 B/cycle of TMA writes.
//
//   per CTA:  warp 0  TMA producer  (own A rows + own half of the B rows, completion
//                                    bytes credited to the leader's full barrier)
//             warp 1  leader only: tcgen05.mma.cta_group::2 issuer; both: TMEM alloc
//             warps 2..5 epilogue of the CTA's own 128 accumulator rows
//   smem ring: 3 stages x [A_hi | A_lo | Bhalf_hi | Bhalf_lo] (64 KB)
//   TMEM: 2 accumulator buffers x 256 columns (all 512 columns)
//
// Same operand descriptors, staging layout and K-split rule as contraction_tc.cu.
// (Experimental: selected with NSGP_PAIR_KERNEL=1; see contraction_tc.cu.)
#include "common.cuh"
#include "tc_common.cuh"

namespace nsgp {
namespace tc {

namespace {

constexpr int BT = 256;            // pair tile edge
constexpr int STAGES = 3;
constexpr uint32_t kABytes = BM * BK * 4;                    // 16 KB, one operand plane
constexpr uint32_t kStageBytes = 4 * kABytes;                // 64 KB
constexpr uint32_t kTmemCols = 512;

struct Tile2 { int rb, cb; };

__device__ __forceinline__ Tile2 decode_tile2(int tiles_1d, int tile) {
  Tile2 t;
  int rb = 0;
  for (;; ++rb) {
    int cnt = tiles_1d - rb;
    if (tile < cnt) { t.rb = rb; t.cb = rb + tile; break; }
    tile -= cnt;
  }
  return t;
}

// number of br-row TMA boxes covering [r0, r0+128) below `rows`
__device__ __forceinline__ int seg_count(int r0, int rows, int br) {
  int n = 0;
  for (int r = r0; r < r0 + BM && r < rows; r += br) ++n;
  return n;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
contraction_tc2_gram_kernel(const __grid_constant__ TcMaps maps,
                            const __grid_constant__ TcParams p) {
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
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);      // leader's expect_tx arrive + the peer's arrive
      mbar_init(&empty_bar[s], 1);     // one multicast commit
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);     // one multicast commit
      mbar_init(&tmem_empty[b], 8);    // 4 epilogue warps x 2 CTAs (leader's copy is used)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_base_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int n_items = p.n_tiles * p.splits;

  if (warp == 0) {
    // ============================ TMA producer (both CTAs) ============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int split = item / p.n_tiles;
        const Tile2 tc = decode_tile2(p.tiles_m, item - split * p.n_tiles);
        const int kb0 = (int)((long long)p.nkb * split / p.splits);
        const int kb1 = (int)((long long)p.nkb * (split + 1) / p.splits);
        const bool share = tc.rb == tc.cb;           // diagonal: the B half IS the A tile
        const int r0 = tc.rb * BT + (int)rank * BM;
        const int c0 = tc.cb * BT + (int)rank * BM;
        const int segs_a = seg_count(r0, p.A.rows, p.A.br);
        const int segs_b = share ? 0 : seg_count(c0, p.B.rows, p.B.br);
        uint32_t tx_total = 0;
        if (rank == 0) {
          // bytes both CTAs will land on this barrier
          int sa = segs_a + seg_count(r0 + BM, p.A.rows, p.A.br);
          int sb = share ? 0 : segs_b + seg_count(c0 + BM, p.B.rows, p.B.br);
          tx_total = 2u * (uint32_t)(sa * p.A.br + sb * p.B.br) * BK * 4u;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], tx_total);
          else mbar_arrive_remote(&full_bar[stage], 0);
          const uint32_t sbase = smem_u32(smem + stage * kStageBytes);
          const int kx0 = kb * BK;
          for (int s = 0; s < segs_a; ++s) {
            const int r = r0 + s * p.A.br;
            const int t = r / p.A.Cs, c = r - t * p.A.Cs;
            const uint32_t off = (uint32_t)(s * p.A.br) * (BK * 4);
            tma_load_2d_2sm(sbase + off, &maps.a[0][t], &full_bar[stage], kx0, c);
            tma_load_2d_2sm(sbase + kABytes + off, &maps.a[1][t], &full_bar[stage], kx0, c);
          }
          for (int s = 0; s < segs_b; ++s) {
            const int r = c0 + s * p.B.br;
            const int t = r / p.B.Cs, c = r - t * p.B.Cs;
            const uint32_t off = (uint32_t)(s * p.B.br) * (BK * 4);
            tma_load_2d_2sm(sbase + 2 * kABytes + off, &maps.b[0][t], &full_bar[stage], kx0, c);
            tma_load_2d_2sm(sbase + 3 * kABytes + off, &maps.b[1][t], &full_bar[stage], kx0, c);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA only) ============================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BT, BT);      // M = 256 over the pair, N = 256
      uint32_t stage = 0, phase = 0;
      uint32_t local_item = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters, ++local_item) {
        const int split = item / p.n_tiles;
        const Tile2 tc = decode_tile2(p.tiles_m, item - split * p.n_tiles);
        const int kb0 = (int)((long long)p.nkb * split / p.splits);
        const int kb1 = (int)((long long)p.nkb * (split + 1) / p.splits);
        const bool share = tc.rb == tc.cb;
        const uint32_t buf = local_item & 1;
        const uint32_t use = local_item >> 1;
        mbar_wait_warp(&tmem_empty[buf], (use & 1) ^ 1, lane);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BT;
        for (int kb = kb0; kb < kb1; ++kb) {
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
              const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
              tc_mma_tf32_2sm(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              tc_mma_tf32_2sm(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
              tc_mma_tf32_2sm(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
            }
            tc_commit_2sm_mc(&empty_bar[stage], 0x3);          // free the slot in both CTAs
            if (kb == kb1 - 1) tc_commit_2sm_mc(&tmem_full[buf], 0x3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (kb1 <= kb0 && lane == 0) tc_commit_2sm_mc(&tmem_full[buf], 0x3);
        __syncwarp();
      }
    }
  } else {
    // ============================ epilogue (both CTAs, own 128 rows) ============================
    const int quad = warp & 3;
    uint32_t local_item = 0;
    for (int item = cluster_id; item < n_items; item += n_clusters, ++local_item) {
      const int split = item / p.n_tiles;
      const Tile2 tc = decode_tile2(p.tiles_m, item - split * p.n_tiles);
      const int kb0 = (int)((long long)p.nkb * split / p.splits);
      const int kb1 = (int)((long long)p.nkb * (split + 1) / p.splits);
      const uint32_t buf = local_item & 1;
      const uint32_t use = local_item >> 1;
      mbar_wait_warp(&tmem_full[buf], use & 1, lane, 200);
      tc_fence_after();
      const int rblk = tc.rb * BT + (int)rank * BM;
      const int row = rblk + quad * 32 + lane;
      const bool row_ok = row < p.A.rows && kb1 > kb0;
      float* orow = p.out + (long long)row * p.ld;
#pragma unroll 1
      for (int chunk = 0; chunk < BT / 32; ++chunk) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * BT + chunk * 32;
        tc_ld32(taddr, v);
        tc_wait_ld();
        const int col0 = tc.cb * BT + chunk * 32;
        // chunks strictly left of this CTA's diagonal block carry nothing needed
        if (row_ok && col0 < p.n_cols && col0 + 31 >= rblk) {
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
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tmem_empty[buf]);
        else mbar_arrive_remote(&tmem_empty[buf], 0);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();                       // the peer may still be reading this CTA's smem / barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(kTmemCols)
                 : "memory");
  }
}

}  // namespace

int launch_tc2_gram(const TcMaps& maps, const TcParams& p, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * kStageBytes + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    NSGP_CHECK_CUDA(cudaFuncSetAttribute(contraction_tc2_gram_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int items = p.n_tiles * p.splits;
  int clusters = sm_count() / 2;
  if (items < clusters) clusters = items;
  ProfScope prof(kProfGram, stream);
  contraction_tc2_gram_kernel<<<2 * clusters, kThreads, smem, stream>>>(maps, p);
  NSGP_LAUNCHED();
  return 0;
}

}  // namespace tc
}  // namespace nsgp
