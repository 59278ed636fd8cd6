// tcgen05 backward: epilogues of the two activation-side products (A^T dR and W^T dx1), the batch-reduction
// product dW -= s1 * dx1 * V^T (both operands K-major, split over CTAs along the batch), and the launch schedule.
// Recurrences: SURVEY.md 7.1 / oracle manual_backward; arithmetic identical to epilogues.cuh (m1_quad, EpiBG1/2).
#pragma once
#include "common.cuh"
#include "epilogues.cuh"
#include "umma_gemm.cuh"
#include "umma_epilogues.cuh"

namespace dladmm {
namespace umma {

// where row-reduced parameter-gradient partial sums go.
//   PS  (all parameters scalar): one entry per epilogue warp of the grid  -> part[slot * nentries + entry]
//   !PS: one entry per (32-column group, feature row)                     -> part[(slot * ngroups + group) * prow + row]
struct RedOut { float* part; int nentries; int ngroups; int prow; };

// Sums over the 32 lanes of N values per lane in N-1 + log2(32/N) shuffles (a warp_sum per value would take 5 N): at
// every stage a lane keeps one half of its values and hands the other half to its partner.  Returns, in lane l, the total
// of value index  l >> (5 - log2 N)  (N = 32: lane l owns value l; N = 8: four lanes share each value).  N a power of two.
template <int N>
__device__ __forceinline__ float warp_multi_sum(float (&x)[N], int lane) {
  int s = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < n / 2; ++j) {
      const float send = up ? x[j] : x[j + n / 2];
      const float keep = up ? x[j + n / 2] : x[j];
      x[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  float v = x[0];
#pragma unroll
  for (; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}
__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

// One parameter-gradient contribution of element (row i of the chunk, this lane's column).
//   PS  : register accumulation (st.red[IDX]), reduced once per kernel (red_finish).
//   !PS : stashed in st.rv[i * NRED + IDX]; red_flush reduces the chunk's CHUNK x NRED values over the warp's 32 columns with
//         one multi-value butterfly and writes one partial per (slot, 32-column group, row).
template <bool PS, int IDX, int NRED, class State>
__device__ __forceinline__ void red_put(State& st, const BP& q, int i, int row, i64 col, bool ok, float val) {
  if constexpr (PS) {                               // every parameter is a (1,1) scalar
    st.red[IDX] += ok ? val : 0.f;                  // (entries of parameters without a gradient are never read back)
  } else {
    if (q.g == nullptr) return;                     // warp-uniform
    if (q.period) {                                 // per-slot parameter (main_lena.py:35-36): direct accumulation
      if (ok) atomicAdd(q.g + (i64)row * q.rs + col % q.period, val);
      return;
    }
    st.rv[i * NRED + IDX] = ok ? val : 0.f;
  }
}
template <int NRED>
__device__ __forceinline__ uint32_t red_mask(const BP* const (&qs)[NRED]) {
  uint32_t m = 0;
#pragma unroll
  for (int r = 0; r < NRED; ++r) m |= (qs[r]->g != nullptr && qs[r]->p != nullptr && qs[r]->period == 0) ? (1u << r) : 0u;
  return m;
}
template <bool PS, int NRED, int CHUNK, int NPOW, class State>
__device__ __forceinline__ void red_flush(State& st, const RedOut& ro, const int (&slots)[NRED], uint32_t gmask, int row0, int n_feat,
                                          i64 group, int lane) {
  if constexpr (!PS) {
    if (gmask == 0) return;                         // warp-uniform: no per-row gradient is wanted
    const float v = warp_multi_sum<NPOW>(st.rv, lane);
    constexpr int SH = 5 - ilog2(NPOW);
    const int idx = lane >> SH;
    if ((lane & ((1 << SH) - 1)) != 0 || idx >= CHUNK * NRED) return;
    const int i = idx / NRED, r = idx % NRED, row = row0 + i;
    int slot = slots[0];
#pragma unroll
    for (int k = 1; k < NRED; ++k) slot = r == k ? slots[k] : slot;
    if (row < n_feat && ((gmask >> r) & 1u)) ro.part[((i64)slot * ro.ngroups + group) * ro.prow + row] = v;
  }
}

template <bool PS, int NRED>
__device__ __forceinline__ void red_finish(const RedOut& ro, const int (&slots)[NRED], float (&acc)[NRED], int entry, int lane) {
  if (!PS) return;
#pragma unroll
  for (int r = 0; r < NRED; ++r) {
    const float s = warp_sum(acc[r]);
    if (lane == 0) ro.part[(i64)slots[r] * ro.nentries + entry] = s;
  }
}

// rank of array i among the present ones (its position in the staged slot), -1 if absent
__device__ __forceinline__ int slot_rank(uint32_t mask, int i) { return (mask >> i) & 1u ? __popc(mask & ((1u << i) - 1u)) : -1; }

// dZ_k = gZ_k + carried + A^T dR ; dx1 = dZ_k * (m+ + m-) ; dtheta1 ; dx1 is the operand of the next two products
template <int PM>
struct UEpiBG1 {
  static constexpr bool PS = PM == PM_SCALAR;
  static constexpr int WARPS = 16;                 // measured: 1.97 -> 1.67 ms per 15 layers against 8 warps
  static constexpr int PAIR_MIN_K = 192;           // CTA-pair kernel from this reduction length on (C1: 1.58 -> 1.45 ms per 15 layers)
  static constexpr int CHUNK = 8;
  static constexpr int NIN = 2;                    // gZ_k, carried dZ -- each optional (the fused loss takes sign(Z_k) from bits 2, 3 of the mask)
  struct State { float red[1]; float rv[PS ? 1 : CHUNK]; uint32_t gmask; int lane; float lsc; int o_gz, o_cz, o_zk, o_mk; bool fast; };
  struct Pre { unsigned mk[CHUNK]; };
  const float* __restrict__ gZ; const float* cZin; const uint8_t* __restrict__ maskZ;
  BP th1; float* dx1; RedOut ro; i64 B;
  const float* __restrict__ Zk; float lz; const float* __restrict__ lscale;   // fused L1-L1 loss cotangent on Z_k
  uint32_t in_mask;
  void host_inputs(const float* (&p)[MAX_EIN]) const { p[0] = gZ; p[1] = cZin; }
  const uint8_t* host_mask() const { return maskZ; }
  static constexpr int NROWP = 0;
  __device__ __forceinline__ void begin(State& st) const {
    st.red[0] = 0.f; st.lane = threadIdx.x & 31;
    { const BP* const qs[1] = {&th1}; st.gmask = red_mask<1>(qs); }
    st.lsc = lscale ? lz * __ldg(lscale) : 0.f;
    st.o_gz = slot_rank(in_mask, 0) * SUBF(CHUNK); st.o_cz = slot_rank(in_mask, 1) * SUBF(CHUNK); st.o_zk = -1;
    st.o_mk = (in_mask & EIN_MASK_BIT) ? __popc(in_mask & ~EIN_MASK_BIT) * SUBF(CHUNK) * 4 : -1;   // byte offset of the staged mask
    // the fused-loss training step below the top layer: carried dZ and the mask bytes staged, no upstream stack, loss term on
    st.fast = lscale != nullptr && st.o_gz < 0 && st.o_cz >= 0 && st.o_mk >= 0;
  }
  __device__ __forceinline__ void end(State& st, int entry, int lane) const {
    const int slots[1] = {SL_TH1};
    if (th1.g && th1.period == 0) red_finish<PS, 1>(ro, slots, st.red, entry, lane);
  }
  __device__ __forceinline__ void prefetch(Pre& pre, int row0, i64 b, bool valid, int n_feat) const {
    if (in_mask & EIN_MASK_BIT) return;             // mask bytes come through the staging ring
#pragma unroll
    for (int i = 0; i < CHUNK; ++i) {
      const bool ok = valid && row0 + i < n_feat;
      pre.mk[i] = ok ? (unsigned)__ldg(maskZ + (i64)(row0 + i) * B + b) : 0u;
    }
  }
  template <bool FULL>
  __device__ __forceinline__ void apply(State& st, const float* __restrict__ slot, int col, const Pre& pre, int row0, i64 b, bool valid,
                                        const float (&v)[CHUNK], int n_feat, i64 group) const {
    // the common case (every row of the chunk exists, fused-loss step below the top layer) gets a copy of the row loop without the
    // warp-uniform tests for which inputs exist: this epilogue is issue-bound (ncu: 67 % of the issue slots), the four tests per
    // row were a fifth of its instructions
    if (FULL && st.fast) body<true, true>(st, slot, col, pre, row0, b, valid, v, n_feat, group);
    else body<FULL, false>(st, slot, col, pre, row0, b, valid, v, n_feat, group);
  }
  template <bool FULL, bool FAST>
  __device__ __forceinline__ void body(State& st, const float* __restrict__ slot, int col, const Pre& pre, int row0, i64 b, bool valid,
                                       const float (&v)[CHUNK], int n_feat, i64 group) const {
    if constexpr (!PS) {
#pragma unroll
      for (int i = 0; i < CHUNK; ++i) st.rv[i] = 0.f;
    }
    // one 64-bit address per chunk, 32-bit row offsets inside it (the host refuses batches of 2^28 columns and more): a 64-bit
    // row * B + b per row and array was 12 of the 41 instructions per element row of this issue-bound epilogue (ncu source page)
    // (round 2, SASS of this loop: indexing dx1c[i * B] still cost IADD3 + IMAD.X + LEA + LEA.HI.X + a constant-bank load per row;
    // a byte pointer stepped by one 64-bit stride is two adds.  Also tried: decoding the mask byte with PRMT table lookups instead of
    // predicates + FSEL -- ptxas re-materialises the table per PRMT, same instruction count, dropped.)
    char* dxp = reinterpret_cast<char*>(dx1 + ((i64)row0 * B + b));
    const i64 stride = B * (i64)sizeof(float);
#pragma unroll
    for (int i = 0; i < CHUNK; ++i, dxp += stride) {
      const int row = row0 + i;
      if (!FULL && row >= n_feat) continue;         // warp-uniform
      // FAST: a column past the batch has a zero accumulator, zero staged inputs and a zero mask byte (TMA fills out-of-bounds
      // elements with zeros), so its contribution to the reduction is an exact zero without a select
      const bool ok = FAST ? true : valid;
      float dz = v[i];
      if (!FAST && st.o_gz >= 0) dz += slot[st.o_gz + i * TILE_B + col];
      if (FAST || st.o_cz >= 0) dz += slot[st.o_cz + i * TILE_B + col];
      const unsigned mk = (FAST || st.o_mk >= 0) ? (unsigned)reinterpret_cast<const uint8_t*>(slot)[st.o_mk + i * TILE_B + col] : pre.mk[i];
      if (FAST || lscale) dz += st.lsc * ((mk & 4u) ? 1.f : ((mk & 8u) ? -1.f : 0.f));      // lsc * sign(Z_k)   (warp-uniform branch)
      const float mp = (mk & 1u) ? 1.f : 0.f, mn = (mk & 2u) ? 1.f : 0.f;
      const float o = dz * (mp + mn);
      red_put<PS, 0, 1>(st, th1, i, row, b, ok, dz * (mn - mp));
      if (valid) *reinterpret_cast<float*>(dxp) = o;
    }
    { const int slots[1] = {SL_TH1}; red_flush<PS, 1, CHUNK, CHUNK>(st, ro, slots, st.gmask, row0, n_feat, group, st.lane); }
  }
};

// dV = -s1 * W^T dx1 ; dbeta1, dss1 ; carried dL, dT ; then the elementwise cotangent flow of layer k-1 (m1):
// writes dR for the next A^T dR product, carried dE and dL.
template <int FAM, int PM>
struct UEpiBG2 {
  static constexpr bool PS = PM == PM_SCALAR;
  static constexpr int WARPS = 16;                 // 7 staged arrays: with the 120 KB ring (OP_STAGES = 2) four parts keep 2 slots each in flight
  static constexpr int CHUNK = 4;     // up to 7 staged arrays + mask per element: keep one ring slot small (depth >= EPI_PARTS)
  static constexpr int OP_STAGES = 2; // 2 operand stages, 120 KB staging ring (8 slots of 14.8 KB instead of 4): see op_stages_of
  static constexpr int NIN = 7;    // L_{k-1} (tied), T_k, cL | cE, E_{k-1}, L_{k-2}, E_{k-2}(B)   (the last four only below the top layer)
                                   // upstream cotangents gL, gE, gT (generic autograd path only) are read straight from global memory
  struct State { float red[6]; float rv[PS ? 1 : 32]; uint32_t gmask; PV<PM> b1, bL, b2, ss2, ss2_2; float s1; int lane; float lsc; int o[NIN]; int o_mk; bool fused; };
  struct Pre { unsigned mk[CHUNK]; };
  // layer k
  const float* __restrict__ Lp; const float* __restrict__ Tk; BP b1, ss1; const float* cLin; const float* cEin;
  int has_prev;
  // layer k-1 (m1)
  const float* __restrict__ Ek; const float* __restrict__ Ep; const float* __restrict__ Lpp; const uint8_t* __restrict__ maskE;
  const float* __restrict__ gE; const float* __restrict__ gL; const float* __restrict__ gT;
  BP bL, b2, ss2, ss2_2, th2;
  float* __restrict__ dR; float* cE; float* cL;
  RedOut ro; i64 B;
  float lw; const float* __restrict__ lscale;       // fused L1-L1 loss cotangents on E_{k-1} / T_k
  int lkind;                                        // 2: LASSO residual term (cotangent proportional to E_{k-1} - T_k itself)
  uint32_t in_mask;
  void host_inputs(const float* (&p)[MAX_EIN]) const {
    p[0] = ss1.p ? Lp : nullptr;                     // L_{k-1} only feeds V_k for d(ss1) (tied variant)
    p[1] = Tk; p[2] = cLin;
    p[3] = has_prev ? cEin : nullptr; p[4] = has_prev ? Ek : nullptr; p[5] = has_prev ? Lpp : nullptr;
    p[6] = (has_prev && FAM == DLADMM_FAMILY_B) ? Ep : nullptr;
  }
  const uint8_t* host_mask() const { return (has_prev && FAM != DLADMM_FAMILY_C) ? maskE : nullptr; }
  static constexpr int NROWP = PM == PM_ROWS ? 5 : 0;
  __device__ __forceinline__ void row_params(BP (&q)[5]) const { q[0] = b1; q[1] = bL; q[2] = b2; q[3] = ss2; q[4] = ss2_2; }
  __device__ __forceinline__ void bind_rows(State& st, const float* tab, int n) const {
    st.b1.tab = tab; st.bL.tab = tab + n; st.b2.tab = tab + 2 * n; st.ss2.tab = tab + 3 * n; st.ss2_2.tab = tab + 4 * n;
  }
  __device__ __forceinline__ void begin(State& st) const {
#pragma unroll
    for (int r = 0; r < 6; ++r) st.red[r] = 0.f;
    st.b1.init(b1); st.bL.init(bL); st.b2.init(b2); st.ss2.init(ss2); st.ss2_2.init(ss2_2);
    st.s1 = ss1.p ? __ldg(ss1.p) : 1.f;
    st.lane = threadIdx.x & 31;
    st.lsc = lscale ? lw * __ldg(lscale) : 0.f;
    {   // order of st.red / st.rv entries: bL, th2, ss2 (C: ss2_1), b2 (C: ss2_2), b1, ss1
      const BP* const qs[6] = {&bL, &th2, &ss2, FAM == DLADMM_FAMILY_C ? &ss2_2 : &b2, &b1, &ss1};
      st.gmask = red_mask<6>(qs);
      if (!has_prev) st.gmask &= 0x30u;             // top layer of the chain: only b1 / ss1 get a contribution here
    }
#pragma unroll
    for (int i = 0; i < NIN; ++i) st.o[i] = slot_rank(in_mask, i) * SUBF(CHUNK);
    st.o_mk = (in_mask & EIN_MASK_BIT) ? __popc(in_mask & ~EIN_MASK_BIT) * SUBF(CHUNK) * 4 : -1;
    // the fused-loss training step of the untied variants: no upstream stacks, mask bytes staged (or family C: none), no L_{k-1} input
    st.fused = lscale != nullptr && gL == nullptr && gE == nullptr && gT == nullptr && st.o[0] < 0 &&
               (FAM == DLADMM_FAMILY_C || st.o_mk >= 0);
  }
  __device__ __forceinline__ void end(State& st, int entry, int lane) const {
    const int slots[6] = {SL_BL, SL_TH2, SL_SS2, SL_B2, SL_B1, SL_SS1};
    red_finish<PS, 6>(ro, slots, st.red, entry, lane);
  }
  __device__ __forceinline__ void prefetch(Pre& pre, int row0, i64 b, bool valid, int n_feat) const {
    if (in_mask & EIN_MASK_BIT) return;             // mask bytes come through the staging ring
#pragma unroll
    for (int i = 0; i < CHUNK; ++i) {
      const bool ok = valid && has_prev && FAM != DLADMM_FAMILY_C && row0 + i < n_feat;
      pre.mk[i] = ok ? (unsigned)__ldg(maskE + (i64)(row0 + i) * B + b) : 0u;
    }
  }
  // SURE: the array is staged whenever control reaches the read (T_k and the carried dL always; the layer-below arrays behind
  // has_prev) -- no presence test per element (they were 7 of the hot loop's 43 ISETP + the selects behind them)
  template <bool SURE = false>
  __device__ __forceinline__ float in(const State& st, const float* slot, int a, int i, int col) const {
    if constexpr (SURE) return slot[st.o[a] + i * TILE_B + col];
    else return st.o[a] >= 0 ? slot[st.o[a] + i * TILE_B + col] : 0.f;
  }
  template <bool FULL>
  __device__ __forceinline__ void apply(State& st, const float* __restrict__ slot, int col, const Pre& pre, int row0, i64 b, bool valid,
                                        const float (&v)[CHUNK], int n_feat, i64 group) const {
    // the common case (all rows of the chunk exist, a layer below) gets a copy of the row loop without warp-uniform
    // branches so that the rows of the chunk are scheduled together
    // ... and, in the fused-loss training step, without the warp-uniform tests for upstream cotangent stacks either
    if (FULL && has_prev) {
      if (st.fused) rows<true, true>(st, slot, col, pre, row0, b, valid, v, n_feat, group);
      else rows<true, false>(st, slot, col, pre, row0, b, valid, v, n_feat, group);
    } else rows<false, false>(st, slot, col, pre, row0, b, valid, v, n_feat, group);
  }
  template <bool FAST, bool FUSED>
  __device__ __forceinline__ void rows(State& st, const float* __restrict__ slot, int col, const Pre& pre, int row0, i64 b, bool valid,
                                       const float (&v)[CHUNK], int n_feat, i64 group) const {
    if constexpr (!PS) {
#pragma unroll
      for (int i = 0; i < 32; ++i) st.rv[i] = 0.f;
    }
    rows_body<FAST, FUSED>(st, slot, col, pre, row0, b, valid, v, n_feat, group);
    const int slots[6] = {SL_BL, SL_TH2, SL_SS2, SL_B2, SL_B1, SL_SS1};
    red_flush<PS, 6, CHUNK, 32>(st, ro, slots, st.gmask, row0, n_feat, group, st.lane);
  }
  template <bool FAST, bool FUSED>
  __device__ __forceinline__ void rows_body(State& st, const float* __restrict__ slot, int col, const Pre& pre, int row0, i64 b, bool valid,
                                            const float (&v)[CHUNK], int n_feat, i64 group) const {
    const i64 off0 = (i64)row0 * B + b;              // one 64-bit offset per chunk, 32-bit row offsets inside it (see UEpiBG1)
    // byte pointers stepped by one 64-bit stride (see UEpiBG1::body)
    char* dRp = reinterpret_cast<char*>(dR + off0); char* cEp = reinterpret_cast<char*>(cE + off0); char* cLp = reinterpret_cast<char*>(cL + off0);
    const i64 stride = B * (i64)sizeof(float);
    const unsigned Bu = (unsigned)B;
#pragma unroll
    for (int i = 0; i < CHUNK; ++i, dRp += stride, cEp += stride, cLp += stride) {
      const int row = row0 + i;
      if (!FAST && row >= n_feat) continue;         // warp-uniform
      // fused fast path: a column past the batch has a zero accumulator and zero staged inputs (TMA fills out-of-bounds elements
      // with zeros; sgn(0) = 0), so every reduction term is an exact zero without a select; the stores stay guarded
      const bool ok = (FAST && FUSED) ? true : valid;
      const i64 off = off0 + (unsigned)i * Bu;
      const float vb1 = st.b1.at(row, b);
      const float tk = in<true>(st, slot, 1, i, col);
      const float dV = -st.s1 * v[i];
      red_put<PS, 4, 6>(st, b1, i, row, b, ok, dV * tk);
      if constexpr (!FUSED) {                         // d(ss1) exists in the tied variants only: V_k = L_{k-1} + beta1 T_k recomputed
        const float var = in(st, slot, 0, i, col) + vb1 * tk;
        red_put<PS, 5, 6>(st, ss1, i, row, b, ok, -var * v[i]);
      }
      float dL = in<true>(st, slot, 2, i, col) + dV;
      float dT = vb1 * dV;
      if (!FAST && !has_prev) continue;             // warp-uniform
      // ---- layer k-1: (dL, dT, dE) -> dR, carried dE, carried dL (m1_quad in epilogues.cuh) ----
      float dE = in<true>(st, slot, 3, i, col);
      if constexpr (!FUSED) {
        if (gL) dL += ok ? __ldg(gL + off) : 0.f;       // warp-uniform branches
        if (gE) dE += ok ? __ldg(gE + off) : 0.f;
        if (gT) dT += ok ? __ldg(gT + off) : 0.f;
      }
      const float tn = tk, ek = in<true>(st, slot, 4, i, col), lpp = in<true>(st, slot, 5, i, col);
      if (FUSED || lscale) { const float r = ek - tn; const float sl = st.lsc * (lkind == 2 ? r : sgn(r)); dE += sl; dT -= sl; }
      const float vbL = st.bL.at(row, b);
      red_put<PS, 0, 6>(st, bL, i, row, b, ok, dL * tn);
      const float dTt = dT + vbL * dL;
      const float dEt = dE + dTt;
      float dRv, nE, nL;
      const unsigned mk = (FAM != DLADMM_FAMILY_C && (FUSED || st.o_mk >= 0)) ? (unsigned)reinterpret_cast<const uint8_t*>(slot)[st.o_mk + i * TILE_B + col] : pre.mk[i];
      if (FAM == DLADMM_FAMILY_B) {
        const float ep = in<true>(st, slot, 6, i, col);
        const float vb2 = st.b2.at(row, b), vs2 = st.ss2.at(row, b);
        const float that = (tn - ek) + ep;
        const float q = lpp + vb2 * that;
        const float mp = (mk & 1u) ? 1.f : 0.f, mn = (mk & 2u) ? 1.f : 0.f;
        const float du = dEt * (mp + mn);
        red_put<PS, 1, 6>(st, th2, i, row, b, ok, dEt * (mn - mp));
        const float dQ = -vs2 * du;
        red_put<PS, 2, 6>(st, ss2, i, row, b, ok, -du * q);
        const float dThat = vb2 * dQ;
        red_put<PS, 3, 6>(st, b2, i, row, b, ok, dQ * that);
        dRv = dTt + dThat; nE = du + dThat; nL = dL + dQ;
      } else if (FAM == DLADMM_FAMILY_A) {
        const float vb2 = st.b2.at(row, b);
        const float mp = (mk & 1u) ? 1.f : 0.f, mn = (mk & 2u) ? 1.f : 0.f;
        const float du = dEt * (mp + mn);
        red_put<PS, 1, 6>(st, th2, i, row, b, ok, dEt * (mn - mp));
        red_put<PS, 3, 6>(st, b2, i, row, b, ok, -du * lpp);
        dRv = dTt - du; nE = 0.f; nL = dL - vb2 * du;
      } else {
        const float v1 = st.ss2.at(row, b), v2 = st.ss2_2.at(row, b);
        red_put<PS, 2, 6>(st, ss2, i, row, b, ok, dEt * (ek - tn));
        red_put<PS, 3, 6>(st, ss2_2, i, row, b, ok, -dEt * lpp);
        dRv = dTt - v1 * dEt; nE = 0.f; nL = dL - v2 * dEt;
      }
      if (valid) {
        *reinterpret_cast<float*>(dRp) = dRv;
        *reinterpret_cast<float*>(cEp) = nE;
        *reinterpret_cast<float*>(cLp) = nL;
      }
    }
  }
};

// ---- dW[i,j] += alpha * sum_b P[i,b] * Q[j,b]  (P = dx1 (d x B), Q = V_k (m x B)), K = batch ------------------
// Both operands are K-major (batch contiguous).  One 128 x 256 output tile per CTA over a slice of the batch;
// partial results are added to dW with fp32 reductions (red.global.add).
constexpr int NT_EPI_WARPS = 16;          // epilogue warps of the dW kernel (atomics of a 256 x 256 tile): 4 per quadrant
// Rows of dW per CTA: 256 = two tcgen05.mma of M = 128 per k-step into two TMEM accumulators, both reading the same Q tile.  The kernel
// is bound by shared-memory bandwidth (ncu: l1tex 71 %, tensor pipe 39 %: TMA writes + in-smem split + operand reads = 120 KB per
// 16-column chunk of a 128 x 256 tile); per flop the 256-row tile moves 96 KB -- Q is staged and split once for twice the products.
constexpr int NT_PROWS = 256;
// Mixed mode (NPASS == 4) of the dW product: P.Q^T = trunc(P).trunc(Q)^T on kind::tf32 (the raw words) + bf16(P_small).bf16(Q)^T +
// bf16(P).bf16(Q_small)^T on kind::f16.  Both operands are K-major rows of 16 floats (64 bytes, SWIZZLE_64B: 16-byte chunk c of
// row r sits at chunk c ^ ((r >> 1) & 3)); the split writes, for every row, a 64-byte row [bf16(x)[16] | bf16(x_small)[16]] with the
// same swizzle into the second half of the stage.  A thread owns half a row: 2 x 16 bytes in, 2 x 16 bytes out.
template <int ROWS, int NTHREADS>
__device__ __forceinline__ void split_rows_mix(const uint8_t* raw, uint8_t* packed, int tid) {
  constexpr int UNITS = ROWS * 2;
  constexpr int N = (UNITS + NTHREADS - 1) / NTHREADS;
  float4 va[N], vb[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int u = tid + i * NTHREADS;
    if (u < UNITS) {
      const int r = u >> 1, h = u & 1, sw = (r >> 1) & 3;
      va[i] = *reinterpret_cast<const float4*>(raw + r * 64 + (((2 * h) ^ sw) << 4));
      vb[i] = *reinterpret_cast<const float4*>(raw + r * 64 + (((2 * h + 1) ^ sw) << 4));
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int u = tid + i * NTHREADS;
    if (u < UNITS) {
      const int r = u >> 1, h = u & 1, sw = (r >> 1) & 3;
      const float4 x0 = va[i], x1 = vb[i];
      uint4 hi, lo;
      hi.x = pack_bf16x2(x0.x, x0.y); hi.y = pack_bf16x2(x0.z, x0.w); hi.z = pack_bf16x2(x1.x, x1.y); hi.w = pack_bf16x2(x1.z, x1.w);
      lo.x = pack_bf16x2(tf32_small(x0.x), tf32_small(x0.y)); lo.y = pack_bf16x2(tf32_small(x0.z), tf32_small(x0.w));
      lo.z = pack_bf16x2(tf32_small(x1.x), tf32_small(x1.y)); lo.w = pack_bf16x2(tf32_small(x1.z), tf32_small(x1.w));
      *reinterpret_cast<uint4*>(packed + r * 64 + ((h ^ sw) << 4)) = hi;
      *reinterpret_cast<uint4*>(packed + r * 64 + (((2 + h) ^ sw) << 4)) = lo;
    }
  }
}

template <int NPASS, int KC>
struct NtPlan {
  static constexpr int NOPS = NPASS >= 3 ? 2 : 1;
  static constexpr int A_BYTES = NT_PROWS * KC * 4;                   // P tile: two M = 128 halves that share the Q tile of the stage
  static constexpr int B_BYTES = TILE_N * KC * 4;
  static constexpr int STAGE_BYTES = NOPS * (A_BYTES + B_BYTES);     // [P raw | Q raw] [P small | Q small]
  static constexpr int RAW_BYTES = A_BYTES + B_BYTES;                 // what TMA delivers
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 256 + 1024;
};

struct NtShape {
  int M, N;            // valid rows of P (d) and of Q (m)
  i64 B;
  i64 chunk;           // batch columns per CTA (multiple of KC)
  int ldc;
  float acc_unit;      // ACC_RZ_BIAS_PER_MMA x MMAs per k-chunk (0: no compensation)
};

template <int NPASS, int KC>
__global__ void __launch_bounds__(roles_threads(NT_EPI_WARPS), 1)
umma_nt_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ, NtShape ns,
               const float* __restrict__ s1ptr, float sign, float* __restrict__ C) {
  using Plan = NtPlan<NPASS, KC>;
  constexpr int STAGES = Plan::STAGES;
  constexpr uint32_t LAYOUT = KC == 32 ? LAYOUT_SW128 : LAYOUT_SW64;
  constexpr uint32_t SBO = 8 * KC * 4;
  constexpr int EPI_PARTS = NT_EPI_WARPS / 4, SPLIT_WARP0 = EPI_WARP0 + NT_EPI_WARPS;
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment as an offset from the __shared__ symbol: keeps the address space visible to the compiler (LDS/STS, not generic)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + STAGES * Plan::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* ready = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * STAGES + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * NT_PROWS;
  const int halves = ns.M - i0 > 128 ? 2 : 1;         // (a last tile of <= 128 valid rows skips its second half)
  const int n0 = blockIdx.z * TILE_N;                 // tile of Q rows (columns of dW)
  const i64 b_begin = (i64)blockIdx.y * ns.chunk;
  i64 b_end = b_begin + ns.chunk;
  if (b_end > ns.B) b_end = ns.B;
  const int k_chunks = b_end > b_begin ? (int)((b_end - b_begin + KC - 1) / KC) : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmP); prefetch_tmap(&tmQ);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], NT_EPI_WARPS); }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (k_chunks > 0) {
    if (warp == 0) {
      int s = 0; uint32_t ph = 0;
      for (int kc = 0; kc < k_chunks; ++kc) {
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
          uint8_t* st = smem + s * Plan::STAGE_BYTES;
          mbar_expect_tx(&full[s], Plan::RAW_BYTES);
          const int bc = (int)(b_begin + (i64)kc * KC);
          tma_load_2d(st, &tmP, &full[s], bc, i0);
          tma_load_2d(st + Plan::A_BYTES, &tmQ, &full[s], bc, n0);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = make_idesc(128, TILE_N, 0, 0);
      constexpr uint32_t hi = desc_hi(SBO, LAYOUT);
      const uint32_t st0 = smem_u32(smem);
      const uint32_t a_lo0 = desc_lo(st0, 16), b_lo0 = desc_lo(st0 + Plan::A_BYTES, 16);
      int s = 0; uint32_t ph = 0;
      for (int kc = 0; kc < k_chunks; ++kc) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + s * (Plan::STAGE_BYTES >> 4), b_lo = b_lo0 + s * (Plan::STAGE_BYTES >> 4);
        // raw * raw (big * big) needs no split: issued while the splitters work on the stage
        constexpr uint32_t HALF = (128 * KC * 4) >> 4;        // descriptor offset of the second 128 rows of the P tile
        if (elect_one()) {
          for (int h = 0; h < halves; ++h) {
#pragma unroll
            for (int ks = 0; ks < KC / UMMA_K; ++ks)
              umma_tf32(tmem_base + h * TILE_N, desc_at(hi, a_lo + h * HALF + ks * 2), desc_at(hi, b_lo + ks * 2), idesc, (kc == 0 && ks == 0) ? 0u : 1u);
          }
        }
        __syncwarp();
        if (NPASS >= 3) {
          mbar_wait(&ready[s], ph);
          tc_fence_after();
        }
        if (elect_one()) {
          if (NPASS == 4) {
            // bf16 correction products (K = 16 = the chunk): packed rows [bf16(x) | bf16(x_small)], 32 bytes each
            constexpr uint32_t idesc16 = make_idesc(128, TILE_N, 0, 0, 1u);
            const uint32_t b16 = b_lo + (Plan::RAW_BYTES >> 4);
            for (int h = 0; h < halves; ++h) {
              const uint32_t a16 = a_lo + h * HALF + (Plan::RAW_BYTES >> 4);
              umma_f16(tmem_base + h * TILE_N, desc_at(hi, a16 + 2), desc_at(hi, b16), idesc16, 1u);      // P_small . Q
              umma_f16(tmem_base + h * TILE_N, desc_at(hi, a16), desc_at(hi, b16 + 2), idesc16, 1u);      // P . Q_small
            }
          }
          if (NPASS == 3) {
            for (int h = 0; h < halves; ++h) {
#pragma unroll
            for (int ks = 0; ks < KC / UMMA_K; ++ks) {
              const uint64_t da = desc_at(hi, a_lo + h * HALF + ks * 2), db = desc_at(hi, b_lo + ks * 2);
              const uint64_t das = desc_at(hi, a_lo + h * HALF + (Plan::RAW_BYTES >> 4) + ks * 2);
              const uint64_t dbs = desc_at(hi, b_lo + (Plan::RAW_BYTES >> 4) + ks * 2);
              umma_tf32(tmem_base + h * TILE_N, das, db, idesc, 1u);
              umma_tf32(tmem_base + h * TILE_N, da, dbs, idesc, 1u);
            }
            }
          }
          umma_commit(&empty[s]);
          if (kc == k_chunks - 1) umma_commit(tfull);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    } else if (warp >= SPLIT_WARP0) {
      // (the four dedicated splitter warps of the common role layout idle here: the split is done by the 16 epilogue warps, below)
    } else {
      // During the mainloop the epilogue warps have nothing to do (one accumulator, drained once at the end): all 16 of them
      // split the operand tiles of every stage in shared memory -- 24 KB per chunk spread over 512 threads instead of 128, which takes
      // the split off the stage turn-around (TMA -> split -> MMA -> free), the bound of this kernel's mainloop.
      if (NPASS >= 3) {
        const int tid = threadIdx.x - EPI_WARP0 * 32;
        int s = 0; uint32_t ph = 0;
        for (int kc = 0; kc < k_chunks; ++kc) {
          mbar_wait(&full[s], ph);
          uint8_t* st = smem + s * Plan::STAGE_BYTES;
          if (NPASS == 3) split_tile<Plan::RAW_BYTES, NT_EPI_WARPS * 32>(st, st + Plan::RAW_BYTES, tid);
          else split_rows_mix<(Plan::RAW_BYTES / 64), NT_EPI_WARPS * 32>(st, st + Plan::RAW_BYTES, tid);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ready[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
      const int q = warp & 3;
      const int half = (warp - EPI_WARP0) >> 2;
      // (acc_unit: accumulator compensation per k-chunk of this CTA's batch slice, see ACC_RZ_BIAS_PER_MMA in umma_gemm.cuh)
      const float alpha = sign * (s1ptr ? __ldg(s1ptr) : 1.f) * (1.0f + ns.acc_unit * (float)k_chunks);
      mbar_wait(tfull, 0);                              // every MMA has retired: the operand stages are free
      tc_fence_after();
      const uint32_t t00 = tmem_base + half * (TILE_N / EPI_PARTS) + ((uint32_t)(q * 32) << 16);
      // A TMEM lane is a ROW of dW, so a warp's 32 lanes hold 32 rows x (consecutive columns): adding straight from the
      // registers touches 32 sectors per instruction.  Each 32 x 32 block is transposed through shared memory (the free operand
      // stages, 33-float pitch) so that one reduction instruction covers 32 consecutive columns of one row (4-5 sectors): 8x
      // fewer L2 reduction operations for the same 148 x 32 K partial sums.
      float* tr = reinterpret_cast<float*>(smem) + (warp - EPI_WARP0) * (32 * 33);
      static_assert(NT_EPI_WARPS * 32 * 33 * 4 <= Plan::STAGES * Plan::STAGE_BYTES, "transpose scratch fits the operand ring");
      for (int hh = 0; hh < halves; ++hh) {
      const int row0 = i0 + hh * 128 + q * 32;
      const uint32_t t0 = t00 + hh * TILE_N;
#pragma unroll 1
      for (int c = 0; c < (TILE_N / EPI_PARTS) / 32; ++c) {
        const int j0 = n0 + half * (TILE_N / EPI_PARTS) + c * 32;
        if (j0 >= ns.N) break;
        float v[32];
        tmem_ld32(t0 + c * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) tr[lane * 33 + i] = alpha * v[i];
        __syncwarp();
        const bool colok = j0 + lane < ns.N;
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          if (row0 + r < ns.M && colok) atomicAdd(C + (i64)(row0 + r) * ns.ldc + j0 + lane, tr[r * 33 + lane]);
        }
        __syncwarp();
      }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace umma
}  // namespace dladmm
