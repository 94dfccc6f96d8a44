// Bring-up micro-benchmark: raw tcgen05.mma issue/execute rate on shared-memory-resident
// operands (no TMA, no epilogue), one CTA per SM so the whole chip is under load.
//   mode 0: 3xTF32 K-block pattern (12 MMAs: main + 2 cross terms), M=128 N=128 K=8
//   mode 1: hi*hi only (4 MMAs per K block)
//   mode 2: 3xTF32 pattern with N=256
//   mode 3: kind::f16 (bf16) M=128 N=128 K=16, 12 MMAs per block
#include "common.cuh"
#include "tc_common.cuh"

namespace nsgp {
using namespace tc;

__global__ void __launch_bounds__(128, 1)
mma_rate_kernel(int mode, int iters, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // fill 6 planes of 16 KB with small non-trivial tf32 values
  float* f = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < 6 * 4096; i += blockDim.x)
    f[i] = __uint_as_float((0x3f800000u + ((i * 2654435761u) & 0x007fe000u)));
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t sb = smem_u32(smem);
    const uint64_t a_hi = make_kmajor_sw128_desc(sb), a_lo = make_kmajor_sw128_desc(sb + 16384);
    const uint64_t b_hi = make_kmajor_sw128_desc(sb + 32768), b_lo = make_kmajor_sw128_desc(sb + 65536);
    const int n = (mode == 2) ? 256 : 128;
    const uint32_t idesc = make_idesc_tf32(128, n);
    // kind::f16, bf16 x bf16 -> f32: c_format 1, a_format 1, b_format 1
    const uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
    __syncwarp();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 0 || mode == 2) {
        tc_mma_kblock_3xtf32(tmem, tmem + 256, a_hi, a_lo, b_hi, b_lo, idesc, it > 0, 0u);
      } else if (mode == 1) {
        tc_mma_kblock_3xtf32(tmem, tmem + 256, a_hi, a_hi, b_hi, b_hi, idesc, it > 0, 3u);
      } else {
        asm volatile(
            "{\n\t.reg .pred pe, pt;\n\t.reg .b32 i;\n\t"
            "setp.eq.b32 pt, 0, 0;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "mov.b32 i, 0;\n\t"
            "L_%=: \n\t"
            "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pt;\n\t"
            "add.s32 i, i, 1;\n\t"
            "setp.lt.s32 pt, i, 12;\n\t"
            "@pt bra L_%=;\n\t"
            "setp.eq.b32 pt, 0, 0;\n\t}" ::"r"(tmem), "l"(a_hi), "l"(b_hi), "r"(idesc16)
            : "memory");
      }
    }
    tc_commit_elect(&bar);
    mbar_wait_warp(&bar, 0, lane);
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512)
                 : "memory");
  }
}

// TMA load latency / throughput probe: `depth` stages of 64 KB (4 boxes of 128 rows x 128 B)
// in flight per CTA, `iters` stages in total; the consumer only waits and releases.
//   out[cta] = total cycles.  map: 2-D tensor map (K, rows) of the probed pitch.
__global__ void __launch_bounds__(64, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap map, int iters, int depth, int kblocks,
                 int row_blocks, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[3], empty[3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0) {
    uint32_t st = 0, ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait_warp(&empty[st], ph ^ 1, lane);
      mbar_expect_tx_elect(&full[st], 4u * 16384u);
      const int kb = (it * 7 + blockIdx.x * 3) % kblocks;
      const int rb = (blockIdx.x + it) % row_blocks;
      for (int b = 0; b < 4; ++b)
        tma_load_2d_elect(smem_u32(smem + st * 65536 + b * 16384), &map, &full[st], kb * 32,
                          ((rb + b) % row_blocks) * 128);
      if (++st == (uint32_t)depth) { st = 0; ph ^= 1; }
    }
  } else {
    uint32_t st = 0, ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait_warp(&full[st], ph, lane);
      if (lane == 0) mbar_arrive(&empty[st]);
      if (++st == (uint32_t)depth) { st = 0; ph ^= 1; }
    }
    if (lane == 0) out[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
}

int debug_tma_probe(const float* base, long long pitch_elems, int K, int rows, int iters,
                    int depth, unsigned long long* out_dev, int n_ctas, cudaStream_t stream) {
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                          const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                          CUtensorMapFloatOOBfill);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  NSGP_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  NSGP_REQUIRE(fp && depth >= 1 && depth <= 3, "tma_probe: bad arguments");
  CUtensorMap map;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)pitch_elems * 4};
  cuuint32_t box[2] = {32, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((Enc)fp)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NSGP_REQUIRE(r == CUDA_SUCCESS, "tma_probe: encode failed (%d)", (int)r);
  const size_t smem = 3 * 65536 + 1024;
  NSGP_CHECK_CUDA(cudaFuncSetAttribute(tma_probe_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tma_probe_kernel<<<n_ctas, 64, smem, stream>>>(map, iters, depth, K / 32, rows / 128, out_dev);
  NSGP_LAUNCHED();
  return 0;
}

// Bulk-copy probe: every CTA streams its own contiguous region of `src` into shared memory
// with `depth` chunks of `chunk` bytes in flight (cp.async.bulk, mbarrier complete_tx) and
// drops the data.  How much HBM bandwidth can n_ctas SMs pull when the in-flight bytes live
// in shared memory instead of registers?  out[cta] = cycles.
__global__ void __launch_bounds__(64, 1)
bulk_probe_kernel(const uint8_t* __restrict__ src, long long bytes_per_cta, int chunk, int depth,
                  unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full[16], empty[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  const uint8_t* base = src + (long long)blockIdx.x * bytes_per_cta;
  const int iters = (int)(bytes_per_cta / chunk);
  if (warp == 0) {
    uint32_t st = 0, ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait_warp(&empty[st], ph ^ 1, lane);
      mbar_expect_tx_elect(&full[st], (uint32_t)chunk);
      if (lane == 0)
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"(smem_u32(smem_raw + (size_t)st * chunk)), "l"(base + (long long)it * chunk),
              "r"(chunk), "r"(smem_u32(&full[st]))
            : "memory");
      __syncwarp();
      if (++st == (uint32_t)depth) { st = 0; ph ^= 1; }
    }
  } else {
    uint32_t st = 0, ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait_warp(&full[st], ph, lane);
      if (lane == 0) mbar_arrive(&empty[st]);
      if (++st == (uint32_t)depth) { st = 0; ph ^= 1; }
    }
    if (lane == 0) out[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
}

// The same with ONE 3-D tensor-map box per stage: (256 floats, rows_per_box rows, B images)
// - the shape a batch-mean staging kernel needs (B images' chunks in one request).
__global__ void __launch_bounds__(64, 1)
tma3d_probe_kernel(const __grid_constant__ CUtensorMap map, int boxes_per_cta, int box_rows,
                   int box_bytes, int depth, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full[16], empty[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0) {
    uint32_t st = 0, ph = 0;
    for (int it = 0; it < boxes_per_cta; ++it) {
      mbar_wait_warp(&empty[st], ph ^ 1, lane);
      mbar_expect_tx_elect(&full[st], (uint32_t)box_bytes);
      const int row = (blockIdx.x * boxes_per_cta + it) * box_rows;
      if (lane == 0)
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
            "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_raw + (size_t)st * box_bytes)),
            "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&full[st])), "r"(0), "r"(row),
            "r"(0)
            : "memory");
      __syncwarp();
      if (++st == (uint32_t)depth) { st = 0; ph ^= 1; }
    }
  } else {
    uint32_t st = 0, ph = 0;
    for (int it = 0; it < boxes_per_cta; ++it) {
      mbar_wait_warp(&full[st], ph, lane);
      if (lane == 0) mbar_arrive(&empty[st]);
      if (++st == (uint32_t)depth) { st = 0; ph ^= 1; }
    }
    if (lane == 0) out[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
}

int debug_tma3d_probe(const float* base, long long img_elems, int B, int box_rows,
                      int boxes_per_cta, int depth, unsigned long long* out_dev, int n_ctas,
                      cudaStream_t stream) {
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                          const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                          CUtensorMapFloatOOBfill);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  NSGP_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  const int box_bytes = 256 * 4 * box_rows * B;
  NSGP_REQUIRE(fp && depth >= 1 && depth <= 16 && (size_t)box_bytes * depth <= 220 * 1024 &&
                   img_elems % 256 == 0, "tma3d_probe: bad arguments");
  CUtensorMap map;
  cuuint64_t gdim[3] = {256, (cuuint64_t)(img_elems / 256), (cuuint64_t)B};
  cuuint64_t gstr[2] = {1024, (cuuint64_t)img_elems * 4};
  cuuint32_t box[3] = {256, (cuuint32_t)box_rows, (cuuint32_t)B};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = ((Enc)fp)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstr, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NSGP_REQUIRE(r == CUDA_SUCCESS, "tma3d_probe: encode failed (%d)", (int)r);
  NSGP_CHECK_CUDA(cudaFuncSetAttribute(tma3d_probe_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  tma3d_probe_kernel<<<n_ctas, 64, (size_t)box_bytes * depth + 1024, stream>>>(
      map, boxes_per_cta, box_rows, box_bytes, depth, out_dev);
  NSGP_LAUNCHED();
  return 0;
}

int debug_bulk_probe(const void* src, long long bytes_per_cta, int chunk, int depth,
                     unsigned long long* out_dev, int n_ctas, cudaStream_t stream) {
  NSGP_REQUIRE(depth >= 1 && depth <= 16 && chunk % 16 == 0 && (size_t)chunk * depth <= 220 * 1024,
               "bulk_probe: bad arguments");
  const size_t smem = (size_t)chunk * depth;
  NSGP_CHECK_CUDA(cudaFuncSetAttribute(bulk_probe_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  bulk_probe_kernel<<<n_ctas, 64, smem, stream>>>((const uint8_t*)src, bytes_per_cta, chunk, depth,
                                                  out_dev);
  NSGP_LAUNCHED();
  return 0;
}

// Holds `threads` threads and `smem` bytes of shared memory on every SM for `cycles`
// clocks without touching memory: separates the occupancy / L1-carve-out cost that a
// persistent contraction kernel imposes on co-running HBM-bound kernels from the
// bandwidth contention (scripts/overlap_probe.py).
__global__ void occupy_kernel(long long cycles, int has_smem) {
  extern __shared__ uint8_t occ_smem[];
  if (has_smem && threadIdx.x == 0) occ_smem[0] = 1;
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) __nanosleep(200);
}

int debug_occupy(int threads, size_t smem, long long cycles, int n_ctas, cudaStream_t stream) {
  NSGP_CHECK_CUDA(cudaFuncSetAttribute(occupy_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
  occupy_kernel<<<n_ctas, threads, smem, stream>>>(cycles, smem > 0 ? 1 : 0);
  NSGP_LAUNCHED();
  return 0;
}

int debug_mma_rate(int mode, int iters, unsigned long long* out_dev, int n_ctas,
                   cudaStream_t stream) {
  const size_t smem = 6 * 16384 + 1024;
  NSGP_CHECK_CUDA(cudaFuncSetAttribute(mma_rate_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mma_rate_kernel<<<n_ctas, 128, smem, stream>>>(mode, iters, out_dev);
  NSGP_LAUNCHED();
  return 0;
}
}  // namespace nsgp
