// Fused epilogues for the tcgen05 kernels.  In the TMEM accumulator a thread owns ONE batch column and walks
// over feature rows (CH per tcgen05.ld), so every global access below is a coalesced 128-byte segment per warp
// and per-row parameters are warp-uniform.  The arithmetic is the reference's, op by op (see epilogues.cuh for
// the file:line map).  The only extra output is V_{k+1} = L_k + beta1_{k+1} T_{k+1}, the operand of the next W V
// product; the 3xTF32 operand split (x = trunc_tf32(x) + small) happens in shared memory inside the consumer.
//
// The (rows x B) arrays an epilogue reads elementwise are staged by TMA into a shared-memory ring (see
// umma_gemm.cuh): apply() receives `slot`, the staged inputs of its chunk -- present arrays back to back, each
// [CHUNK][128 columns] floats -- and reads them with conflict-free LDS (lane = column).  Only the 1-byte prox masks
// are register-prefetched one chunk ahead (prefetch()).
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>
#include "epilogues.cuh"

namespace dladmm {
namespace umma {

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// How a kernel instantiation reads the per-layer broadcast parameters (template argument PM of the epilogue functors):
//   PM_SCALAR  every parameter of the call is a (1,1) scalar (scalar / tied / lasso variants): one register each, no loads;
//   PM_ROWS    scalars and (rows,1) vectors only (full / ltheta): the epilogue warps copy the 256 rows of the current tile of every
//              parameter into a shared-memory table (fill_rowtab, umma_gemm.cuh) and at() is ONE unconditional broadcast read.
//              With ~220 KB of the SM configured as shared memory the L1 is a few KB: a __ldg per row and parameter missed it, and
//              the "table or __ldg" choice at run time cost a convergence barrier per access (measured: A Z epilogue of `full`
//              61 us per tile against 20 us scalar);
//   PM_GENERAL anything else, i.e. per-batch-slot (rows, bs) parameters (main_lena.py:35-36): __ldg with the column period.
enum { PM_GENERAL = 0, PM_SCALAR = 1, PM_ROWS = 2 };
template <int PM>
struct PV {
  const float* p; int rs; int period; float s; const float* tab;
  __device__ __forceinline__ void init(const BP& q) {
    p = q.p; rs = q.rs; period = q.period; tab = nullptr;
    s = (q.p && q.rs == 0 && q.period == 0) ? __ldg(q.p) : 0.f;
  }
  __device__ __forceinline__ float at(int row, i64 col) const {
    if (PM == PM_SCALAR) return s;
    if (PM == PM_ROWS) return tab[row];             // tab is pre-offset by the tile's first row; warp-uniform address
    i64 off = (i64)row * rs;
    if (period) off += col % period;
    return __ldg(p + off);
  }
};
__host__ __device__ constexpr int SUBF(int chunk) { return chunk * TILE_B; }   // floats of one staged array of one chunk

struct NoPre {};

// T_0 = A Z0 + E0 - X, and V_0 = L0 + beta1_0 * T_0 for the first Z-step
template <int PSCALAR>
struct UEpiT0 {
  static constexpr int WARPS = 8;
  static constexpr int CHUNK = 8;
  static constexpr int NIN = 3;                    // E0, X, L0
  struct State { PV<PSCALAR> b1; };
  typedef NoPre Pre;
  const float* __restrict__ E0; const float* __restrict__ X; const float* __restrict__ L0; float* __restrict__ T0;
  BP b1; float* __restrict__ V; i64 B; uint32_t in_mask;
  __nv_bfloat16* __restrict__ Vh; i64 ldh;         // bf16 mode: the W V operand as bf16 (pitch ldh)
  void host_inputs(const float* (&p)[MAX_EIN]) const { p[0] = E0; p[1] = X; p[2] = L0; }
  const uint8_t* host_mask() const { return nullptr; }
  static constexpr int NROWP = PSCALAR == PM_ROWS ? 1 : 0;
  __device__ __forceinline__ void row_params(BP (&q)[1]) const { q[0] = b1; }
  __device__ __forceinline__ void bind_rows(State& st, const float* tab, int n) const { st.b1.tab = tab; }
  __device__ __forceinline__ void begin(State& st) const { st.b1.init(b1); }
  __device__ __forceinline__ void end(State&, int, int) const {}
  __device__ __forceinline__ void prefetch(Pre&, int, i64, bool, int) const {}
  template <bool FULL>
  __device__ __forceinline__ void apply(State& st, const float* __restrict__ slot, int col, const Pre&, int row0, i64 b, bool valid,
                                        const float (&v)[CHUNK], int n_feat, i64) const {
    if (!valid) return;
    // one 64-bit offset per chunk, 32-bit row offsets inside it (B < 2^28, checked by the host): see UEpiBG1 in umma_bwd.cuh
    const i64 off0 = (i64)row0 * B + b;
    const unsigned Bu = (unsigned)B;
    float* const T0c = T0 + off0; float* const Vc = V + off0;
#pragma unroll
    for (int i = 0; i < CHUNK; ++i) {
      const int row = row0 + i;
      if (!FULL && row >= n_feat) continue;
      const unsigned ro = (unsigned)i * Bu;
      const float e0 = slot[i * TILE_B + col], x = slot[SUBF(CHUNK) + i * TILE_B + col], l0 = slot[2 * SUBF(CHUNK) + i * TILE_B + col];
      const float t = fsub(fadd(v[i], e0), x);
      T0c[ro] = t;
      if (V || Vh) {                                             // (K = 0: T_0 only, no layer to feed)
        const float vv = fadd(l0, fmul(st.b1.at(row, b), t));
        if (V) Vc[ro] = vv;
        if (Vh) Vh[(i64)row * ldh + b] = __float2bfloat16_rn(vv);
      }
    }
  }
};

// Z_k = act(Z_{k-1} - [ss1*] acc, theta1)
template <int PSCALAR>
struct UEpiZ {
  static constexpr int WARPS = 8;
  static constexpr int CHUNK = 16;
  static constexpr int NIN = 2;                    // Z_{k-1}, Z_label (metrics only)
  struct State { PV<PSCALAR> th1; float s1; float obj; float sq; };
  typedef NoPre Pre;
  const float* __restrict__ Zp; float* __restrict__ Zk; uint8_t* __restrict__ maskZ;
  BP th1; BP ss1; i64 B; uint32_t in_mask;
  float* obj_part;                                 // optional: per-warp partial sums of ||Z_k||_1 (fused objective / DLADMM_MET_L1_Z)
  const float* __restrict__ Zlabel; float* sq_part;  // optional: per-warp partial sums of (Z_label - Z_k)^2 (DLADMM_MET_SQERR_Z)
  __nv_bfloat16* __restrict__ Zh; i64 ldh;         // bf16 mode: Z_k as bf16, the operand of the A Z product
  void host_inputs(const float* (&p)[MAX_EIN]) const { p[0] = Zp; p[1] = sq_part ? Zlabel : nullptr; }
  const uint8_t* host_mask() const { return nullptr; }
  static constexpr int NROWP = PSCALAR == PM_ROWS ? 1 : 0;
  __device__ __forceinline__ void row_params(BP (&q)[1]) const { q[0] = th1; }
  __device__ __forceinline__ void bind_rows(State& st, const float* tab, int n) const { st.th1.tab = tab; }
  __device__ __forceinline__ void begin(State& st) const { st.th1.init(th1); st.s1 = ss1.p ? __ldg(ss1.p) : 1.f; st.obj = 0.f; st.sq = 0.f; }
  __device__ __forceinline__ void end(State& st, int entry, int lane) const {
    if (obj_part) {
      const float s = warp_sum(st.obj);
      if (lane == 0) obj_part[entry] = s;
    }
    if (sq_part) {
      const float s = warp_sum(st.sq);
      if (lane == 0) sq_part[entry] = s;
    }
  }
  __device__ __forceinline__ void prefetch(Pre&, int, i64, bool, int) const {}
  template <bool FULL>
  __device__ __forceinline__ void apply(State& st, const float* __restrict__ slot, int col, const Pre&, int row0, i64 b, bool valid,
                                        const float (&v)[CHUNK], int n_feat, i64) const {
    if (!valid) return;
    // one 64-bit offset per chunk; the stores go through byte pointers stepped by one 64-bit stride per row (two adds per pointer
    // and row; indexing base[i * B] compiled to IADD3 + IMAD.X + LEA + LEA.HI.X + a constant-bank load per store, SASS of round 2)
    const i64 off0 = (i64)row0 * B + b;
    char* zp = reinterpret_cast<char*>(Zk + off0); uint8_t* mp = maskZ + off0;
    const i64 sf = B * (i64)sizeof(float);
#pragma unroll
    for (int i = 0; i < CHUNK; ++i, zp += sf, mp += B) {
      const int row = row0 + i;
      if (!FULL && row >= n_feat) continue;
      const float wv = ss1.p ? fmul(st.s1, v[i]) : v[i];
      unsigned bits;
      const float z = soft_act(fsub(slot[i * TILE_B + col], wv), st.th1.at(row, b), bits);
      *reinterpret_cast<float*>(zp) = z;
      if (Zh) Zh[(i64)row * ldh + b] = __float2bfloat16_rn(z);
      if (maskZ) *mp = (uint8_t)(bits | (z > 0.f ? 4u : 0u) | (z < 0.f ? 8u : 0u));    // bits 2, 3: sign(Z_k) for the fused-loss backward
      if (obj_part) st.obj += fabsf(z);
      if (sq_part) { const float dl = slot[SUBF(CHUNK) + i * TILE_B + col] - z; st.sq += dl * dl; }   // warp-uniform branch
    }
  }
};

// E_k, T_{k+1}, L_k from acc = A Z_k; then V_{k+1} = L_k + beta1_{k+1} * T_{k+1} unless this is the last layer
// MET: also accumulate the per-layer metrics of dladmm_metrics that live on the (m x B) side (two more staged inputs, seven
// running sums per thread) -- a separate instantiation so that the plain forward keeps its registers and its ring depth.
constexpr int ELT_NMET = 7;                        // L1_RES, SQ_RES, SQERR_E, SQERR_AZ, L1_E, DOT_LX, DGAP_L
__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(__expf(x)); }   // F.softplus, threshold 20

template <int FAM, int PSCALAR, bool MET = false>
struct UEpiELT {
  static constexpr int WARPS = 8;
  static constexpr int CHUNK = 8;
  static constexpr int NIN = MET ? 5 : 3;          // X, L_{k-1}, E_{k-1} (family B only) [, E_label, X_clean]
  struct State { PV<PSCALAR> b2, ss2, ss2_2, th2, bL, b1n; float obj; float met[MET ? ELT_NMET : 1]; int o_el, o_xc; };
  typedef NoPre Pre;
  float* obj_part;                                 // optional: per-warp partial sums of the residual term of the objective
  int obj_kind;                                    // 2: 0.5*(X - A Z_k)^2 (LASSO), else |X - A Z_k| = |E_k - T_{k+1}|
  const float* __restrict__ Elabel; const float* __restrict__ Xclean;
  float* met_part; int met_stride;                 // MET: met_part[i * met_stride + entry], i < ELT_NMET
  const float* __restrict__ X; const float* __restrict__ Ep; const float* __restrict__ Lp;
  float* __restrict__ Ek; float* __restrict__ Lk; float* __restrict__ Tn; uint8_t* __restrict__ maskE;
  BP b2, ss2, ss2_2, th2, bL;
  int has_next; BP b1n; float* __restrict__ V;
  __nv_bfloat16* __restrict__ Vh; i64 ldh;         // bf16 mode: V_{k+1} as bf16 (V itself is then only kept for a backward)
  i64 B; uint32_t in_mask;
  void host_inputs(const float* (&p)[MAX_EIN]) const {
    p[0] = X; p[1] = Lp; p[2] = FAM == DLADMM_FAMILY_B ? Ep : nullptr;
    if (MET) { p[3] = Elabel; p[4] = Xclean; }
  }
  const uint8_t* host_mask() const { return nullptr; }
  static constexpr int NROWP = PSCALAR == PM_ROWS ? 6 : 0;
  __device__ __forceinline__ void row_params(BP (&q)[6]) const { q[0] = b2; q[1] = ss2; q[2] = ss2_2; q[3] = th2; q[4] = bL; q[5] = b1n; }
  __device__ __forceinline__ void bind_rows(State& st, const float* tab, int n) const {
    st.b2.tab = tab; st.ss2.tab = tab + n; st.ss2_2.tab = tab + 2 * n; st.th2.tab = tab + 3 * n; st.bL.tab = tab + 4 * n; st.b1n.tab = tab + 5 * n;
  }
  __device__ __forceinline__ void begin(State& st) const {
    st.b2.init(b2); st.ss2.init(ss2); st.ss2_2.init(ss2_2); st.th2.init(th2); st.bL.init(bL); st.b1n.init(b1n);
    st.obj = 0.f;
    if (MET) {
#pragma unroll
      for (int i = 0; i < ELT_NMET; ++i) st.met[i] = 0.f;
      st.o_el = ((in_mask >> 3) & 1u) ? __popc(in_mask & 7u) * SUBF(CHUNK) : -1;
      st.o_xc = ((in_mask >> 4) & 1u) ? __popc(in_mask & 15u) * SUBF(CHUNK) : -1;
    }
  }
  __device__ __forceinline__ void end(State& st, int entry, int lane) const {
    if (obj_part) {
      const float s = warp_sum(st.obj);
      if (lane == 0) obj_part[entry] = s;
    }
    if (MET && met_part) {
#pragma unroll
      for (int i = 0; i < ELT_NMET; ++i) {
        const float s = warp_sum(st.met[i]);
        if (lane == 0) met_part[(size_t)i * met_stride + entry] = s;
      }
    }
  }
  __device__ __forceinline__ void prefetch(Pre&, int, i64, bool, int) const {}
  template <bool FULL>
  __device__ __forceinline__ void apply(State& st, const float* __restrict__ slot, int col, const Pre&, int row0, i64 b, bool valid,
                                        const float (&v)[CHUNK], int n_feat, i64) const {
    if (!valid) return;
    const i64 off0 = (i64)row0 * B + b;              // one 64-bit offset per chunk, stepped byte pointers for the stores (see UEpiZ)
    char* ep_ = reinterpret_cast<char*>(Ek + off0); char* tp = reinterpret_cast<char*>(Tn + off0); char* lp_ = reinterpret_cast<char*>(Lk + off0);
    char* vp = reinterpret_cast<char*>(V + off0);
    uint8_t* mp = maskE + off0;
    const i64 sf = B * (i64)sizeof(float);
#pragma unroll
    for (int i = 0; i < CHUNK; ++i, ep_ += sf, tp += sf, lp_ += sf, vp += sf, mp += B) {
      const int row = row0 + i;
      if (!FULL && row >= n_feat) continue;
      const float x = slot[i * TILE_B + col], lp = slot[SUBF(CHUNK) + i * TILE_B + col], acc = v[i];
      float e;
      unsigned bits = 0;
      if (FAM == DLADMM_FAMILY_B) {
        const float ep = slot[2 * SUBF(CHUNK) + i * TILE_B + col];
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
      *reinterpret_cast<float*>(ep_) = e; *reinterpret_cast<float*>(tp) = t; *reinterpret_cast<float*>(lp_) = l;
      if (obj_part) { const float r = fsub(e, t); st.obj += obj_kind == 2 ? 0.5f * r * r : fabsf(r); }
      if (MET) {
        const float r = fsub(e, t);
        st.met[0] += fabsf(r); st.met[1] += r * r;
        if (st.o_el >= 0) { const float q = slot[st.o_el + i * TILE_B + col] - e; st.met[2] += q * q; }
        if (st.o_xc >= 0) { const float q = slot[st.o_xc + i * TILE_B + col] - acc; st.met[3] += q * q; }
        st.met[4] += fabsf(e); st.met[5] += l * x;
        st.met[6] += softplus_t(l - 1.f) + softplus_t(-l - 1.f);
      }
      if (FAM != DLADMM_FAMILY_C && maskE) *mp = (uint8_t)bits;
      if (has_next) {
        const float vv = fadd(l, fmul(st.b1n.at(row, b), t));
        if (V) *reinterpret_cast<float*>(vp) = vv;
        if (Vh) Vh[(i64)row * ldh + b] = __float2bfloat16_rn(vv);
      }
    }
  }
};

// sum over (d x B) of softplus(acc - a) + softplus(-acc - a) with acc = A^T L_k (main_lena.py:145-147, 224): reduction only
struct UEpiDgap {
  static constexpr int WARPS = 8;
  static constexpr int CHUNK = 16;
  static constexpr int NIN = 0;
  struct State { float s; };
  typedef NoPre Pre;
  float a; float* part; uint32_t in_mask;
  void host_inputs(const float* (&)[MAX_EIN]) const {}
  const uint8_t* host_mask() const { return nullptr; }
  static constexpr int NROWP = 0;
  __device__ __forceinline__ void begin(State& st) const { st.s = 0.f; }
  __device__ __forceinline__ void end(State& st, int entry, int lane) const {
    const float s = warp_sum(st.s);
    if (lane == 0) part[entry] = s;
  }
  __device__ __forceinline__ void prefetch(Pre&, int, i64, bool, int) const {}
  template <bool FULL>
  __device__ __forceinline__ void apply(State& st, const float* __restrict__, int, const Pre&, int row0, i64, bool valid,
                                        const float (&v)[CHUNK], int n_feat, i64) const {
    if (!valid) return;
#pragma unroll
    for (int i = 0; i < CHUNK; ++i) {
      if (!FULL && row0 + i >= n_feat) continue;
      st.s += softplus_t(v[i] - a) + softplus_t(-v[i] - a);
    }
  }
};

}  // namespace umma
}  // namespace dladmm
