// Fused epilogues for the tcgen05 kernels.  In the TMEM accumulator a thread owns ONE batch column and walks
// over feature rows (CH per tcgen05.ld), so every global access below is a coalesced 128-byte segment per warp
// and per-row parameters are warp-uniform.  The arithmetic is the reference's, op by op (see epilogues.cuh for
// the file:line map).  The only extra output is V_{k+1} = L_k + beta1_{k+1} T_{k+1}, the operand of the next W V
// product; the 3xTF32 operand split (x = trunc_tf32(x) + small) happens in shared memory inside the consumer.
//
// Each functor is used in two phases per CH-row chunk so that the global loads of the chunk are all in flight
// before the accumulator is read:  load(in, ...) -> tcgen05.ld -> apply(in, acc, ...).
#pragma once
#include "common.cuh"
#include "epilogues.cuh"

namespace dladmm {
namespace umma {

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// a broadcast parameter resolved once per kernel: scalars live in a register, per-row / per-slot ones are read
// through the read-only path (warp-uniform address when there is no column period)
// PSCALAR = every parameter of the call is a (1,1) scalar (scalar / tied / lasso variants): no loads, less code.
template <bool PSCALAR>
struct PV {
  const float* p; int rs; int period; float s;
  __device__ __forceinline__ void init(const BP& q) {
    p = q.p; rs = q.rs; period = q.period;
    s = (q.p && q.rs == 0 && q.period == 0) ? __ldg(q.p) : 0.f;
  }
  __device__ __forceinline__ float at(int row, i64 col) const {
    if (PSCALAR) return s;
    i64 off = (i64)row * rs;
    if (period) off += col % period;
    return __ldg(p + off);
  }
};

// T_0 = A Z0 + E0 - X, and V_0 = L0 + beta1_0 * T_0 for the first Z-step
template <bool PSCALAR>
struct UEpiT0 {
  static constexpr int CHUNK = CH;
  struct State { PV<PSCALAR> b1; };
  struct In { float e0[CH], x[CH], l0[CH]; };
  const float* __restrict__ E0; const float* __restrict__ X; const float* __restrict__ L0; float* __restrict__ T0;
  BP b1; float* __restrict__ V; i64 B;
  __device__ __forceinline__ void begin(State& st) const { st.b1.init(b1); }
  __device__ __forceinline__ void end(State&, int, int) const {}
  __device__ __forceinline__ void load(In& in, int row0, i64 b, bool valid, int n_feat) const {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const bool ok = valid && row0 + i < n_feat;
      const i64 off = (i64)(row0 + i) * B + b;
      in.e0[i] = ok ? __ldg(E0 + off) : 0.f;
      in.x[i] = ok ? __ldg(X + off) : 0.f;
      in.l0[i] = ok ? __ldg(L0 + off) : 0.f;
    }
  }
  __device__ __forceinline__ void apply(State& st, const In& in, int row0, i64 b, bool valid, const float (&v)[CH], int n_feat, i64) const {
    if (!valid) return;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int row = row0 + i;
      if (row >= n_feat) continue;
      const i64 off = (i64)row * B + b;
      const float t = fsub(fadd(v[i], in.e0[i]), in.x[i]);
      T0[off] = t;
      V[off] = fadd(in.l0[i], fmul(st.b1.at(row, b), t));
    }
  }
};

// Z_k = act(Z_{k-1} - [ss1*] acc, theta1)
template <bool PSCALAR>
struct UEpiZ {
  static constexpr int CHUNK = CH;
  struct State { PV<PSCALAR> th1; float s1; };
  struct In { float zp[CH]; };
  const float* __restrict__ Zp; float* __restrict__ Zk; uint8_t* __restrict__ maskZ;
  BP th1; BP ss1; i64 B;
  __device__ __forceinline__ void begin(State& st) const { st.th1.init(th1); st.s1 = ss1.p ? __ldg(ss1.p) : 1.f; }
  __device__ __forceinline__ void end(State&, int, int) const {}
  __device__ __forceinline__ void load(In& in, int row0, i64 b, bool valid, int n_feat) const {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const bool ok = valid && row0 + i < n_feat;
      in.zp[i] = ok ? __ldg(Zp + (i64)(row0 + i) * B + b) : 0.f;
    }
  }
  __device__ __forceinline__ void apply(State& st, const In& in, int row0, i64 b, bool valid, const float (&v)[CH], int n_feat, i64) const {
    if (!valid) return;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int row = row0 + i;
      if (row >= n_feat) continue;
      const i64 off = (i64)row * B + b;
      const float wv = ss1.p ? fmul(st.s1, v[i]) : v[i];
      unsigned bits;
      const float z = soft_act(fsub(in.zp[i], wv), st.th1.at(row, b), bits);
      Zk[off] = z;
      if (maskZ) maskZ[off] = (uint8_t)bits;
    }
  }
};

// E_k, T_{k+1}, L_k from acc = A Z_k; then V_{k+1} = L_k + beta1_{k+1} * T_{k+1} unless this is the last layer
template <int FAM, bool PSCALAR>
struct UEpiELT {
  static constexpr int CHUNK = CH;
  struct State { PV<PSCALAR> b2, ss2, ss2_2, th2, bL, b1n; };
  struct In { float x[CH], lp[CH], ep[CH]; };
  const float* __restrict__ X; const float* __restrict__ Ep; const float* __restrict__ Lp;
  float* __restrict__ Ek; float* __restrict__ Lk; float* __restrict__ Tn; uint8_t* __restrict__ maskE;
  BP b2, ss2, ss2_2, th2, bL;
  int has_next; BP b1n; float* __restrict__ V;
  i64 B;
  __device__ __forceinline__ void begin(State& st) const {
    st.b2.init(b2); st.ss2.init(ss2); st.ss2_2.init(ss2_2); st.th2.init(th2); st.bL.init(bL); st.b1n.init(b1n);
  }
  __device__ __forceinline__ void end(State&, int, int) const {}
  __device__ __forceinline__ void load(In& in, int row0, i64 b, bool valid, int n_feat) const {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const bool ok = valid && row0 + i < n_feat;
      const i64 off = (i64)(row0 + i) * B + b;
      in.x[i] = ok ? __ldg(X + off) : 0.f;
      in.lp[i] = ok ? __ldg(Lp + off) : 0.f;
      if (FAM == DLADMM_FAMILY_B) in.ep[i] = ok ? __ldg(Ep + off) : 0.f;
    }
  }
  __device__ __forceinline__ void apply(State& st, const In& in, int row0, i64 b, bool valid, const float (&v)[CH], int n_feat, i64) const {
    if (!valid) return;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int row = row0 + i;
      if (row >= n_feat) continue;
      const i64 off = (i64)row * B + b;
      const float x = in.x[i], lp = in.lp[i], acc = v[i];
      float e;
      unsigned bits = 0;
      if (FAM == DLADMM_FAMILY_B) {
        const float ep = in.ep[i];
        const float that = fsub(fadd(acc, ep), x);
        const float vvar = fadd(lp, fmul(st.b2.at(row, b), that));
        const float u = fsub(ep, fmul(st.ss2.at(row, b), vvar));
        e = soft_act(u, st.th2.at(row, b), bits);
      } else if (FAM == DLADMM_FAMILY_A) {
        const float u = fsub(fsub(x, acc), fmul(st.b2.at(row, b), lp));
        e = soft_act(u, st.th2.at(row, b), bits);
      } else {
        const float res = fsub(x, acc);
        e = fsub(fmul(st.ss2.at(row, b), res), fmul(st.ss2_2.at(row, b), lp));
      }
      const float t = fsub(fadd(acc, e), x);
      const float l = fadd(lp, fmul(st.bL.at(row, b), t));
      Ek[off] = e; Tn[off] = t; Lk[off] = l;
      if (FAM != DLADMM_FAMILY_C && maskE) maskE[off] = (uint8_t)bits;
      if (has_next) V[off] = fadd(l, fmul(st.b1n.at(row, b), t));
    }
  }
};

}  // namespace umma
}  // namespace dladmm
