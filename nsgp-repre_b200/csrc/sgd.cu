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
  // one element of the update (SGD_NSCL.py:399-411); returns -lr * (momentum buffer | grad)
  auto update = [&](float w, float& g, float& b) -> float {
    if (wd != 0.f) g = g + wd * w;                       // side effect kept: p.grad is modified
    float src = g;
    if (momentum != 0.f) {
      b = t.first ? (b + g) : (b * momentum + one_minus_damp * g);
      if (nesterov) { g = g + momentum * b; src = g; } else { src = b; }
    }
    return -(lr * src);
  };
  // 128-bit path: every pointer 16-byte aligned, the chunk a whole number of float4s and the
  // staged rows contiguous (pitch == d)
  const bool vec = ((reinterpret_cast<uintptr_t>(t.w) | reinterpret_cast<uintptr_t>(t.g) |
                     reinterpret_cast<uintptr_t>(t.buf) | reinterpret_cast<uintptr_t>(t.u_hi) |
                     reinterpret_cast<uintptr_t>(t.u_lo)) & 15) == 0 &&
                   (t.numel & 3) == 0 && (t.u_hi == nullptr || t.ldu == t.d);
  if (vec) {
    for (long long i = beg + threadIdx.x * 4; i < end; i += 256 * 4) {
      float4 w = *reinterpret_cast<const float4*>(t.w + i);
      float4 g = *reinterpret_cast<const float4*>(t.g + i);
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (momentum != 0.f) b = *reinterpret_cast<const float4*>(t.buf + i);
      float4 u;
      u.x = update(w.x, g.x, b.x); u.y = update(w.y, g.y, b.y);
      u.z = update(w.z, g.z, b.z); u.w = update(w.w, g.w, b.w);
      if (wd != 0.f || (momentum != 0.f && nesterov)) *reinterpret_cast<float4*>(t.g + i) = g;
      if (momentum != 0.f) *reinterpret_cast<float4*>(t.buf + i) = b;
      if (t.u_hi) {
        float4 h, l;
        tf32_split(u.x, h.x, l.x); tf32_split(u.y, h.y, l.y);
        tf32_split(u.z, h.z, l.z); tf32_split(u.w, h.w, l.w);
        *reinterpret_cast<float4*>(t.u_hi + i) = h;
        *reinterpret_cast<float4*>(t.u_lo + i) = l;
        if (t.apply != 0.f) {
          w.x += t.apply * u.x; w.y += t.apply * u.y; w.z += t.apply * u.z; w.w += t.apply * u.w;
          *reinterpret_cast<float4*>(t.w + i) = w;
        }
      } else {
        w.x += u.x; w.y += u.y; w.z += u.z; w.w += u.w;
        *reinterpret_cast<float4*>(t.w + i) = w;
      }
    }
    return;
  }
  for (long long i = beg + threadIdx.x; i < end; i += 256) {
    float w = t.w[i];
    float g = t.g[i];
    float b = momentum != 0.f ? t.buf[i] : 0.f;
    const float upd = update(w, g, b);
    if (wd != 0.f || (momentum != 0.f && nesterov)) t.g[i] = g;
    if (momentum != 0.f) t.buf[i] = b;
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
