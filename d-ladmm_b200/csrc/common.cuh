// Shared device/host helpers for the D-LADMM kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/dladmm.h"

namespace dladmm {

typedef long long i64;

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define DL_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      dladmm::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return DLADMM_ERR_CUDA;                                                                 \
    }                                                                                         \
  } while (0)

#define DL_REQUIRE(cond, ...)                                                                 \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      dladmm::set_error(__VA_ARGS__);                                                         \
      return DLADMM_ERR_INVALID;                                                              \
    }                                                                                         \
  } while (0)

// ---- launch accounting / per-kind timers (definitions in dladmm_api.cu) -----------------------------
struct LaunchScope {
  int kind; cudaStream_t st; void* rec;
  LaunchScope(int kind, cudaStream_t st);
  ~LaunchScope();
};

// ---- broadcast parameter (device view of dladmm_bparam) ----------------------------------------
struct BP {
  const float* p;
  float* g;
  int rs;       // row stride
  int period;   // column period (0 = none)
};

__host__ inline BP make_bp(const dladmm_bparam& q) {
  BP b;
  b.p = q.ptr; b.g = q.grad; b.rs = q.row_stride; b.period = q.col_period;
  return b;
}

__device__ __forceinline__ float bp_at(const BP& q, int row, i64 col) {
  i64 off = (i64)row * q.rs;
  if (q.period) off += col % q.period;
  return __ldg(q.p + off);
}

// value for 4 consecutive columns of one row
__device__ __forceinline__ void bp_at4(const BP& q, int row, i64 col, float (&v)[4]) {
  if (q.period == 0) {
    float s = __ldg(q.p + (i64)row * q.rs);
    v[0] = v[1] = v[2] = v[3] = s;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = __ldg(q.p + (i64)row * q.rs + (col + j) % q.period);
  }
}

// ---- the reference's arithmetic, rounded op by op (ATen elementwise kernels never fuse) ----------
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }

// self_active (main_syn_l1l1_scalar.py:76-77): relu(x - th) - relu(-1.0*x - th); bit0=[x-th>0], bit1=[-x-th>0]
__device__ __forceinline__ float soft_act(float x, float th, unsigned& bits) {
  float a = fsub(x, th);
  float b = fsub(fmul(-1.0f, x), th);
  bits = (a > 0.f ? 1u : 0u) | (b > 0.f ? 2u : 0u);
  float ra = a > 0.f ? a : 0.f;
  float rb = b > 0.f ? b : 0.f;
  return fsub(ra, rb);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 4-wide row access helpers: vector path when the whole quad is in range and 16B aligned.
struct Quad {
  float v[4];
};

__device__ __forceinline__ Quad load4(const float* base, i64 off, int nvalid, bool vec) {
  Quad q;
  if (vec) {
    float4 t = *reinterpret_cast<const float4*>(base + off);
    q.v[0] = t.x; q.v[1] = t.y; q.v[2] = t.z; q.v[3] = t.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) q.v[j] = j < nvalid ? base[off + j] : 0.f;
  }
  return q;
}

__device__ __forceinline__ void store4(float* base, i64 off, const float (&v)[4], int nvalid, bool vec) {
  if (vec) {
    *reinterpret_cast<float4*>(base + off) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nvalid) base[off + j] = v[j];
  }
}

__device__ __forceinline__ void store4_u8(uint8_t* base, i64 off, const unsigned (&b)[4], int nvalid, bool vec) {
  if (vec) {
    *reinterpret_cast<uchar4*>(base + off) = make_uchar4((unsigned char)b[0], (unsigned char)b[1],
                                                          (unsigned char)b[2], (unsigned char)b[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nvalid) base[off + j] = (uint8_t)b[j];
  }
}

__device__ __forceinline__ void load4_u8(const uint8_t* base, i64 off, unsigned (&b)[4], int nvalid, bool vec) {
  if (vec) {
    uchar4 t = *reinterpret_cast<const uchar4*>(base + off);
    b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = j < nvalid ? base[off + j] : 0u;
  }
}

static inline int round_up(int x, int a) { return (x + a - 1) / a * a; }
static inline i64 round_up64(i64 x, i64 a) { return (x + a - 1) / a * a; }

}  // namespace dladmm
