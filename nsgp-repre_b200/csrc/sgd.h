#pragma once
#include "common.cuh"

namespace nsgp {
struct SgdTensorDev {
  float* w;
  float* g;
  float* buf;
  float* u_hi;   // null -> unprotected: w += update in the prologue
  float* u_lo;
  long long numel;
  int first;     // state just created: buf = grad (SGD_NSCL.py:405-406)
  int d;         // protected: row length of update.view(cout, d); staged pitch is ldu
  int ldu;
  float apply;   // protected, low-rank form: also w += apply * update here (0: stage only)
};
int launch_sgd_prologue(const SgdTensorDev* tensors_dev, const int* chunk_start_dev,
                        int n_tensors, int total_chunks, float lr, float momentum,
                        float one_minus_damp, float wd, int nesterov, cudaStream_t s);
int sgd_chunks(long long numel);
}  // namespace nsgp
