// RoIAlign 256 x 7 x 7 extraction over the FPN levels, the step
// immediately before the RePRE statistics
// (mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:45-118; the pooling
// itself is mmcv.ops.RoIAlign, an un-vendored dependency - mmcv >= 2.0.0rc4 for mmdet 3.3 -
// whose published algorithm (Mask R-CNN / Detectron2 ROIAlign, `aligned=True`, average
// pooling, adaptive sampling grid ceil(roi_size / output_size) when sampling_ratio = 0) is
// restated here and checked against torchvision.ops.roi_align, the same algorithm).
//
// One launch covers all levels: the level of a RoI is map_roi_levels (:45-63)
//     floor(log2(sqrt(w * h) / finest_scale + 1e-6)) clamped to [0, L-1]
// evaluated per RoI, and the kernel can, in the same pass,
//   * write the (R, C*7*7) features (what cal_rois stores in rois_etc.pth), and / or
//   * add them into per-class sums (C_cls, C*7*7) + counts - the coarse prototypes
//     (standard_roi_replay_head.py:411-415) without the features ever going to HBM.
#include "../../include/nsgp_repre_b200.h"
#include "common.cuh"

namespace nsgp {

constexpr int kMaxLevels = 8;
struct RoiLevels {
  const float* feat[kMaxLevels];
  int H[kMaxLevels], W[kMaxLevels];
  float scale[kMaxLevels];
  int n;
};

__device__ __forceinline__ float bilinear(const float* __restrict__ f, int H, int W, float y,
                                          float x) {
  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) return 0.f;
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int y_low = (int)y, x_low = (int)x, y_high, x_high;
  if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else { y_high = y_low + 1; }
  if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else { x_high = x_low + 1; }
  const float ly = y - (float)y_low, lx = x - (float)x_low;
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float v1 = __ldg(f + y_low * W + x_low), v2 = __ldg(f + y_low * W + x_high);
  const float v3 = __ldg(f + y_high * W + x_low), v4 = __ldg(f + y_high * W + x_high);
  return hy * hx * v1 + hy * lx * v2 + ly * hx * v3 + ly * lx * v4;
}

__device__ __forceinline__ int roi_level(const float* __restrict__ roi, float finest_scale,
                                         int n_levels) {
  const float scale = sqrtf(__fmul_rn(roi[3] - roi[1], roi[4] - roi[2]));
  float l = floorf(log2f(__fadd_rn(__fdiv_rn(scale, finest_scale), 1e-6f)));
  if (!(l >= 0.f)) l = 0.f;                       // also NaN (negative area) -> level 0
  if (l > (float)(n_levels - 1)) l = (float)(n_levels - 1);
  return (int)l;
}

// Separable form of the bin average.  A sample at (y, x) contributes hy/ly (x) hx/lx to its
// four pixels and the validity test is a test on y OR a test on x, so the sum over the
// gh x gw samples of a bin factorises:
//     out = 1/count * sum_Y sum_X  Wy[ph][Y] * Wx[pw][X] * f[Y][X]
// with Wy[ph][Y] = sum over the bin's gh row samples of their weight on pixel row Y (and the
// same for columns).  The tables depend on the RoI only - built once per block in shared
// memory - and a bin touches ~(bin+2)^2 pixels instead of 4 * gh * gw (3-4x fewer loads).
constexpr int kMaxSpan = 66;         // pixel rows / columns one bin can touch through the table
constexpr int kMaxPooled = 14;

struct AxisTable {
  float w[kMaxPooled][kMaxSpan];
  int first[kMaxPooled], n[kMaxPooled];
};

// one thread per bin index along the axis: accumulate the samples in order (deterministic)
__device__ __forceinline__ bool build_axis(AxisTable& t, int p, float start, float bin, int g,
                                           int size) {
  int first = 0, n = 0;
  bool ok = true;
  for (int k = 0; k < kMaxSpan; ++k) t.w[p][k] = 0.f;
  for (int i = 0; i < g; ++i) {
    float v = start + (float)p * bin + ((float)i + .5f) * bin / (float)g;
    if (v < -1.0f || v > (float)size) continue;            // sample contributes nothing
    if (v <= 0.f) v = 0.f;
    int lo = (int)v, hi;
    if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; } else { hi = lo + 1; }
    const float l = v - (float)lo, h = 1.f - l;
    if (n == 0) first = lo;
    if (hi - first >= kMaxSpan) { ok = false; break; }
    t.w[p][lo - first] += h;
    t.w[p][hi - first] += l;
    n = hi - first + 1;
  }
  t.first[p] = first;
  t.n[p] = n;
  return ok;
}

// grid = (R, channel chunks); block = 256 threads over (channel within the chunk, bin)
__global__ void __launch_bounds__(256)
roi_align_kernel(RoiLevels lv, int channels, const float* __restrict__ rois, int pooled,
                 int sampling_ratio, int aligned, float finest_scale,
                 const long long* __restrict__ labels, int n_classes,
                 float* __restrict__ roi_feats, float* __restrict__ class_sums,
                 int* __restrict__ class_counts, int ch_per_block,
                 const int* __restrict__ levels) {
  __shared__ AxisTable ty, tx;
  __shared__ int s_fallback;
  const int r = blockIdx.x;
  const float* roi = rois + (long long)r * 5;
  const int b = (int)roi[0];
  const int l = levels != nullptr ? min(max(levels[r], 0), lv.n - 1)
                                  : (lv.n > 1 ? roi_level(roi, finest_scale, lv.n) : 0);
  const int H = lv.H[l], W = lv.W[l];
  const float sc = lv.scale[l], off = aligned ? 0.5f : 0.f;
  const float sw = roi[1] * sc - off, sh = roi[2] * sc - off;
  const float ew = roi[3] * sc - off, eh = roi[4] * sc - off;
  float rw = ew - sw, rh = eh - sh;
  if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
  const float bin_h = rh / (float)pooled, bin_w = rw / (float)pooled;
  const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)pooled);
  const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)pooled);
  const float count = fmaxf((float)(gh * gw), 1.f);
  const int bins = pooled * pooled;
  const int c0 = blockIdx.y * ch_per_block;
  const int c1 = min(channels, c0 + ch_per_block);
  if (threadIdx.x == 0) s_fallback = pooled > kMaxPooled ? 1 : 0;
  __syncthreads();
  if (pooled <= kMaxPooled) {
    bool ok = true;
    if (threadIdx.x < pooled) ok = build_axis(ty, threadIdx.x, sh, bin_h, gh, H);
    else if (threadIdx.x >= 32 && threadIdx.x < 32 + pooled)
      ok = build_axis(tx, threadIdx.x - 32, sw, bin_w, gw, W);
    if (!ok) s_fallback = 1;
  }
  __syncthreads();
  const bool direct = s_fallback != 0;
  long long label = -1;
  if (class_sums != nullptr) {
    label = labels[r];
    if (label < 0 || label >= n_classes) label = -1;
    if (label >= 0 && blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(class_counts + label, 1);
  }
  const long long D = (long long)channels * bins;
  for (int e = threadIdx.x; e < (c1 - c0) * bins; e += 256) {
    const int c = c0 + e / bins, bin = e - (e / bins) * bins;
    const int ph = bin / pooled, pw = bin - ph * pooled;
    const float* f = lv.feat[l] + ((long long)b * channels + c) * H * W;
    float acc = 0.f;
    if (!direct) {
      const int ny = ty.n[ph], nx = tx.n[pw];
      const float* f0 = f + (long long)ty.first[ph] * W + tx.first[pw];
      for (int j = 0; j < ny; ++j) {
        const float wy = ty.w[ph][j];
        float row = 0.f;
        for (int i = 0; i < nx; ++i) row += tx.w[pw][i] * __ldg(f0 + j * W + i);
        acc += wy * row;
      }
    } else {
      for (int iy = 0; iy < gh; ++iy) {
        const float y = sh + (float)ph * bin_h + ((float)iy + .5f) * bin_h / (float)gh;
        for (int ix = 0; ix < gw; ++ix) {
          const float x = sw + (float)pw * bin_w + ((float)ix + .5f) * bin_w / (float)gw;
          acc += bilinear(f, H, W, y, x);
        }
      }
    }
    const float v = acc / count;
    const long long o = (long long)c * bins + bin;
    if (roi_feats != nullptr) roi_feats[(long long)r * D + o] = v;
    if (label >= 0) atomicAdd(class_sums + label * D + o, v);
  }
}

// Backward of the same pooling: d feat[Y][X] += Wy[ph][Y] * Wx[pw][X] * d out[c][ph][pw] / count
// (the transpose of the separable forward; atomics because RoIs and bins overlap).
__global__ void __launch_bounds__(256)
roi_align_backward_kernel(RoiLevels lv, int channels, const float* __restrict__ rois, int pooled,
                          int sampling_ratio, int aligned, float finest_scale,
                          const float* __restrict__ grad_out, int ch_per_block,
                          const int* __restrict__ levels) {
  __shared__ AxisTable ty, tx;
  __shared__ int s_fallback;
  const int r = blockIdx.x;
  const float* roi = rois + (long long)r * 5;
  const int b = (int)roi[0];
  const int l = levels != nullptr ? min(max(levels[r], 0), lv.n - 1)
                                  : (lv.n > 1 ? roi_level(roi, finest_scale, lv.n) : 0);
  const int H = lv.H[l], W = lv.W[l];
  const float sc = lv.scale[l], off = aligned ? 0.5f : 0.f;
  const float sw = roi[1] * sc - off, sh = roi[2] * sc - off;
  const float ew = roi[3] * sc - off, eh = roi[4] * sc - off;
  float rw = ew - sw, rh = eh - sh;
  if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
  const float bin_h = rh / (float)pooled, bin_w = rw / (float)pooled;
  const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)pooled);
  const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)pooled);
  const float count = fmaxf((float)(gh * gw), 1.f);
  const int bins = pooled * pooled;
  const int c0 = blockIdx.y * ch_per_block;
  const int c1 = min(channels, c0 + ch_per_block);
  if (threadIdx.x == 0) s_fallback = pooled > kMaxPooled ? 1 : 0;
  __syncthreads();
  if (pooled <= kMaxPooled) {
    bool ok = true;
    if (threadIdx.x < pooled) ok = build_axis(ty, threadIdx.x, sh, bin_h, gh, H);
    else if (threadIdx.x >= 32 && threadIdx.x < 32 + pooled)
      ok = build_axis(tx, threadIdx.x - 32, sw, bin_w, gw, W);
    if (!ok) s_fallback = 1;
  }
  __syncthreads();
  const bool direct = s_fallback != 0;
  const long long D = (long long)channels * bins;
  for (int e = threadIdx.x; e < (c1 - c0) * bins; e += 256) {
    const int c = c0 + e / bins, bin = e - (e / bins) * bins;
    const int ph = bin / pooled, pw = bin - ph * pooled;
    float* f = const_cast<float*>(lv.feat[l]) + ((long long)b * channels + c) * H * W;
    const float g = grad_out[(long long)r * D + (long long)c * bins + bin] / count;
    if (g == 0.f) continue;
    if (!direct) {
      const int ny = ty.n[ph], nx = tx.n[pw];
      float* f0 = f + (long long)ty.first[ph] * W + tx.first[pw];
      for (int j = 0; j < ny; ++j) {
        const float wy = ty.w[ph][j] * g;
        if (wy == 0.f) continue;
        for (int i = 0; i < nx; ++i) {
          const float w = wy * tx.w[pw][i];
          if (w != 0.f) atomicAdd(f0 + j * W + i, w);
        }
      }
    } else {
      for (int iy = 0; iy < gh; ++iy) {
        float y = sh + (float)ph * bin_h + ((float)iy + .5f) * bin_h / (float)gh;
        for (int ix = 0; ix < gw; ++ix) {
          float x = sw + (float)pw * bin_w + ((float)ix + .5f) * bin_w / (float)gw;
          float yy = y;
          if (yy < -1.0f || yy > (float)H || x < -1.0f || x > (float)W) continue;
          if (yy <= 0.f) yy = 0.f;
          if (x <= 0.f) x = 0.f;
          int y_low = (int)yy, x_low = (int)x, y_high, x_high;
          if (y_low >= H - 1) { y_high = y_low = H - 1; yy = (float)y_low; } else { y_high = y_low + 1; }
          if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else { x_high = x_low + 1; }
          const float ly = yy - (float)y_low, lx = x - (float)x_low, hy = 1.f - ly, hx = 1.f - lx;
          atomicAdd(f + y_low * W + x_low, g * hy * hx);
          atomicAdd(f + y_low * W + x_high, g * hy * lx);
          atomicAdd(f + y_high * W + x_low, g * ly * hx);
          atomicAdd(f + y_high * W + x_high, g * ly * lx);
        }
      }
    }
  }
}

}  // namespace nsgp

using namespace nsgp;

extern "C" int repre_roi_align(const float* const* feats, const int32_t* heights,
                               const int32_t* widths, const float* spatial_scales, int n_levels,
                               int batch, int channels, const float* rois, int n_rois,
                               int pooled, int sampling_ratio, int aligned, float finest_scale,
                               const int32_t* levels, const int64_t* labels, int n_classes,
                               float* roi_feats, float* class_sums, int32_t* class_counts,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSGP_REQUIRE(feats && heights && widths && spatial_scales, "roi_align: null level tables");
  NSGP_REQUIRE(n_levels >= 1 && n_levels <= kMaxLevels, "roi_align: 1..%d levels", kMaxLevels);
  NSGP_REQUIRE(batch > 0 && channels > 0 && pooled > 0 && n_rois >= 0, "roi_align: bad sizes");
  NSGP_REQUIRE(roi_feats || class_sums, "roi_align: no output requested");
  NSGP_REQUIRE(!class_sums || (labels && class_counts && n_classes > 0),
               "roi_align: class sums need labels, counts and n_classes");
  RoiLevels lv{};
  lv.n = n_levels;
  for (int i = 0; i < n_levels; ++i) {
    NSGP_REQUIRE(feats[i] && heights[i] > 0 && widths[i] > 0, "roi_align: level %d is empty", i);
    lv.feat[i] = feats[i]; lv.H[i] = heights[i]; lv.W[i] = widths[i];
    lv.scale[i] = spatial_scales[i];
  }
  if (class_sums) {
    const size_t D = (size_t)channels * pooled * pooled;
    NSGP_CHECK_CUDA(cudaMemsetAsync(class_sums, 0, (size_t)n_classes * D * sizeof(float), stream));
    NSGP_CHECK_CUDA(cudaMemsetAsync(class_counts, 0, (size_t)n_classes * sizeof(int), stream));
  }
  if (n_rois == 0) return 0;
  NSGP_REQUIRE(rois != nullptr, "roi_align: null rois");
  // 5 channels x 49 bins = 245 of 256 threads busy for the 7 x 7 case
  int ch_per_block = 256 / (pooled * pooled);
  if (ch_per_block < 1) ch_per_block = 1;
  ch_per_block *= 4;
  dim3 grid(n_rois, ceil_div(channels, ch_per_block));
  ProfScope prof(kProfRepre, stream);
  roi_align_kernel<<<grid, 256, 0, stream>>>(lv, channels, rois, pooled, sampling_ratio, aligned,
                                             finest_scale, (const long long*)labels, n_classes,
                                             roi_feats, class_sums, class_counts, ch_per_block,
                                             levels);
  NSGP_LAUNCHED();
  return 0;
}

extern "C" int repre_roi_align_backward(float* const* grad_feats, const int32_t* heights,
                                        const int32_t* widths, const float* spatial_scales,
                                        int n_levels, int batch, int channels, const float* rois,
                                        int n_rois, int pooled, int sampling_ratio, int aligned,
                                        float finest_scale, const int32_t* levels,
                                        const float* grad_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NSGP_REQUIRE(grad_feats && heights && widths && spatial_scales, "roi_align_backward: null level tables");
  NSGP_REQUIRE(n_levels >= 1 && n_levels <= kMaxLevels, "roi_align_backward: 1..%d levels", kMaxLevels);
  NSGP_REQUIRE(batch > 0 && channels > 0 && pooled > 0 && n_rois >= 0, "roi_align_backward: bad sizes");
  if (n_rois == 0) return 0;
  NSGP_REQUIRE(rois && grad_out, "roi_align_backward: null pointer");
  RoiLevels lv{};
  lv.n = n_levels;
  for (int i = 0; i < n_levels; ++i) {
    NSGP_REQUIRE(grad_feats[i] && heights[i] > 0 && widths[i] > 0,
                 "roi_align_backward: level %d is empty", i);
    lv.feat[i] = grad_feats[i]; lv.H[i] = heights[i]; lv.W[i] = widths[i];
    lv.scale[i] = spatial_scales[i];
  }
  int ch_per_block = 256 / (pooled * pooled);
  if (ch_per_block < 1) ch_per_block = 1;
  ch_per_block *= 4;
  dim3 grid(n_rois, ceil_div(channels, ch_per_block));
  ProfScope prof(kProfRepre, stream);
  roi_align_backward_kernel<<<grid, 256, 0, stream>>>(lv, channels, rois, pooled, sampling_ratio,
                                                      aligned, finest_scale, grad_out,
                                                      ch_per_block, levels);
  NSGP_LAUNCHED();
  return 0;
}
