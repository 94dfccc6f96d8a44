// Fused multi-tensor SGD prologue of SGDNSCL.step (SGD_NSCL.py:387-415):
// weight decay (in place on the gradient, :399-400), momentum buffer (:402-411),
// update = -lr * buf.  Unprotected tensors are updated in place; protected ones
// get their update staged as tf32 hi/lo planes for the projection GEMM, whose
// epilogue performs W += update @ P (:85-95).  One launch for all tensors.
#include "common.cuh"
#include "sgd.h"

namespace nsgp {

constexpr int kChunk = 4096;   // elements per CTA (256 threads x 4 x float4)

__global__ void __launch_bounds__(256)
sgd_prologue_kernel(const SgdTensorDev* __restrict__ tensors,
                    const int* __restrict__ chunk_start, int n_tensors, float lr,
                    float momentum, float one_minus_damp, float wd, int nesterov) {
  // locate the tensor owning this chunk (binary search over the prefix table)
  int lo = 0, hi = n_tensors;
  const int chunk = blockIdx.x;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (chunk_start[mid] <= chunk) lo = mid; else hi = mid;
  }
  const SgdTensorDev t = tensors[lo];
  const long long beg = (long long)(chunk - chunk_start[lo]) * kChunk;
  const long long end = min(t.numel, beg + kChunk);
  for (long long i = beg + threadIdx.x; i < end; i += 256) {
    float w = t.w[i];
    float g = t.g[i];
    if (wd != 0.f) {
      g = g + wd * w;
      t.g[i] = g;                       // side effect kept: p.grad is modified
    }
    float src = g;
    if (momentum != 0.f) {
      float b = t.buf[i];
      b = t.first ? (b + g) : (b * momentum + one_minus_damp * g);
      t.buf[i] = b;
      if (nesterov) {
        g = g + momentum * b;
        t.g[i] = g;
        src = g;
      } else {
        src = b;
      }
    }
    float upd = -(lr * src);
    if (t.u_hi) {
      float h, l;
      tf32_split(upd, h, l);
      long long o = i;
      if (t.ldu != t.d) { long long r = i / t.d; o = r * t.ldu + (i - r * t.d); }
      t.u_hi[o] = h;
      t.u_lo[o] = l;
      if (t.apply != 0.f) t.w[i] = w + t.apply * upd;
    } else {
      t.w[i] = w + upd;
    }
  }
}

int launch_sgd_prologue(const SgdTensorDev* tensors_dev, const int* chunk_start_dev,
                        int n_tensors, int total_chunks, float lr, float momentum,
                        float one_minus_damp, float wd, int nesterov, cudaStream_t stream) {
  if (total_chunks == 0) return 0;
  ProfScope prof(kProfSgd, stream);
  sgd_prologue_kernel<<<total_chunks, 256, 0, stream>>>(
      tensors_dev, chunk_start_dev, n_tensors, lr, momentum, one_minus_damp, wd, nesterov);
  NSGP_LAUNCHED();
  return 0;
}

int sgd_chunks(long long numel) { return (int)((numel + kChunk - 1) / kChunk); }

}  // namespace nsgp
