// tcgen05 / TMA path (DLADMM_PREC_TF32X3, DLADMM_PREC_TF32).  Placeholder until the UMMA kernels land:
// tensor-core precisions report an error instead of silently running something else.
#pragma once
#include "common.cuh"

#define DLADMM_HAS_UMMA 0

namespace dladmm {
static inline size_t umma_workspace_bytes(const dladmm_problem*, int) { return 0; }
static inline int umma_forward(const dladmm_problem*, void*, cudaStream_t) {
  set_error("tensor-core precision requested but this build has no tcgen05 kernels");
  return DLADMM_ERR_INVALID;
}
static inline int umma_backward(const dladmm_problem*, const dladmm_cotangents*, void*, cudaStream_t) {
  set_error("tensor-core precision requested but this build has no tcgen05 kernels");
  return DLADMM_ERR_INVALID;
}
}  // namespace dladmm
