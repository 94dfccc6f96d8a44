// SIMT fp32 contraction engine over K-major virtual operands.
// Bring-up / cross-check path: obviously-correct FFMA tiles that consume exactly
// the same staged planes as the tcgen05 engine (operand element = hi + lo, which
// reconstructs the fp32 value exactly).  Selected with nsgp_set_engine(1).
#include "common.cuh"

namespace nsgp {

__device__ __forceinline__ float fetch(const Operand& o, int row, int k) {
  if (row >= o.rows || k >= o.K) return 0.f;
  int t = row / o.Cs, c = row - t * o.Cs;
  const float* p = o.base + o.tap_off[t] + (long long)c * o.row_pitch + k;
  return __ldg(p) + __ldg(p + o.hl_stride);
}

template <int EPI>
__global__ void __launch_bounds__(256)
contraction_simt_kernel(const __grid_constant__ ContractionArgs a, int tiles_n) {
  __shared__ float As[32][65];
  __shared__ float Bs[32][65];
  const int tid = threadIdx.x;
  int rb, cb;
  if (EPI == kEpiGramAtomic) {
    // blockIdx.x enumerates pairs rb <= cb row by row
    int p = blockIdx.x, nb = tiles_n;
    rb = 0;
    while (p >= nb - rb) { p -= nb - rb; ++rb; }
    cb = rb + p;
  } else {
    rb = blockIdx.x / tiles_n;
    cb = blockIdx.x - rb * tiles_n;
  }
  const int r0 = rb * 64, c0 = cb * 64;
  const int nkb = (a.A.K + 31) / 32;
  const int kb0 = (int)((long long)nkb * blockIdx.y / gridDim.y);
  const int kb1 = (int)((long long)nkb * (blockIdx.y + 1) / gridDim.y);
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4] = {};
  for (int kb = kb0; kb < kb1; ++kb) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int lin = tid + e * 256;
      int row = lin >> 5, kk = lin & 31;
      As[kk][row] = fetch(a.A, r0 + row, kb * 32 + kk);
      Bs[kk][row] = fetch(a.B, c0 + row, kb * 32 + kk);
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = r0 + ty * 4 + i;
    if (r >= a.A.rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = c0 + tx * 4 + j;
      if (c >= a.n_cols) continue;
      float* o = a.out + (long long)r * a.ld + c;
      if (EPI == kEpiGramAtomic) atomicAdd(o, acc[i][j]);
      else *o += a.alpha * acc[i][j];
    }
  }
}

int contraction_simt(const ContractionArgs& a, cudaStream_t stream) {
  int tm = ceil_div(a.A.rows, 64), tn = ceil_div(a.n_cols, 64);
  int nkb = k_blocks(a.A);
  NSGP_REQUIRE(nkb == k_blocks(a.B), "contraction: operands disagree on K blocks");
  if (a.epi == kEpiGramAtomic) {
    int pairs = tm * (tm + 1) / 2;
    int splits = a.splits < 1 ? 1 : (a.splits > nkb ? (nkb > 0 ? nkb : 1) : a.splits);
    dim3 grid(pairs, splits);
    contraction_simt_kernel<kEpiGramAtomic><<<grid, 256, 0, stream>>>(a, tm);
  } else {
    dim3 grid(tm * tn, 1);
    contraction_simt_kernel<kEpiGemmRmw><<<grid, 256, 0, stream>>>(a, tn);
  }
  NSGP_LAUNCHED();
  return 0;
}

}  // namespace nsgp
