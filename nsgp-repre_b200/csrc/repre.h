// Launchers of the RePRE statistics kernels (repre.cu).
#pragma once
#include "common.cuh"

namespace nsgp {
int launch_class_index(const long long* labels, int M, int C, int* counts, int* offsets,
                       int* rows, cudaStream_t s);
int launch_segment_mean(const float* F, int D, const int* seg_off, const int* rows, int nseg,
                        int max_seg_rows, const float* mu, int mode, float* out, cudaStream_t s);
int launch_normalize_split(const float* F, int D, const int* rows, int n, float* hi, float* lo,
                           cudaStream_t s);
int launch_threshold_count(const float* S, int n, int ld, float thresh, unsigned char* mask,
                           int* counts, float* sim_out, cudaStream_t s);
struct ClassExtent {          // one class of a batched cosine-count call
  long long s_off;            // element offset of its Gram in the similarity buffer
  long long mask_off;         // byte offset of its (n x n) mask
  int row_off;                // first row in the concatenated row / count arrays
  int n, ld, pad;
};
int launch_threshold_count_batched(const float* S_all, const ClassExtent* ext_dev, int n_classes,
                                   int max_n, float thresh, unsigned char* mask_all,
                                   int* counts_all, cudaStream_t s);
int launch_replay_gather(const float* protos, const float* sigma, const long long* idx, int P,
                         int D, unsigned long long seed, float* out, cudaStream_t s);
int launch_row_sqnorm(const float* X, int n, int D, int ld, float* out, cudaStream_t s);
int launch_kmeans_argmin(const float* dots, int n, int k, int ld, const float* cnorm,
                         long long* labels, cudaStream_t s);
}  // namespace nsgp
