// Fused epilogues of the D-LADMM layer.  Each functor consumes 4 consecutive batch columns of one
// feature row of a matrix product held in registers (from the FFMA kernel or from TMEM) and performs
// the reference's elementwise chain in the reference's own operation order.
//
//   EpiT0   : T_0 = A Z0 + E0 - X                         main_syn_l1l1_scalar.py:92
//   EpiZ    : Z_k = act(Z_{k-1} - [ss1*] W V_k, theta1)   :93-94 / :111-112 (tied: scalar_tied.py:113-114)
//   EpiELT  : E_k, T_{k+1}, L_k from R = A Z_k            :114-118 (B) / main_lena.py:87-89 (A) /
//                                                         main_syn_lasso_scalar.py:102-107 (C)
//   EpiBG1  : dZ_k = gZ_k + carried + A^T dR ; dx1 = dZ_k*(m+ + m-) ; dtheta1
//   EpiBG2  : dV = -s1 * W^T dx1 ; dbeta1, dss1 ; carried dL, dT ; then the elementwise part of layer
//             k-1's backward (m1_quad) so no standalone elementwise pass is needed between layers
//   m1_quad : cotangent flow through L/T/E-step of one layer (SURVEY.md 7.1 identities)
#pragma once
#include "common.cuh"

namespace dladmm {

// partial-sum slots (row-reduced parameter gradients), see reduce_partials_kernel
enum { SL_BL = 0, SL_TH2 = 1, SL_SS2 = 2, SL_B2 = 3, SL_B1 = 4, SL_SS1 = 5, SL_TH1 = 6, SL_COUNT = 7 };

// accumulate a parameter-gradient contribution: row-reduced params go to `red`, per-slot params
// (main_lena.py:35-36, (m x bs)) go straight to their gradient with an atomic.
__device__ __forceinline__ float sgn(float x) { return (x > 0.f ? 1.f : 0.f) - (x < 0.f ? 1.f : 0.f); }

__device__ __forceinline__ void pgrad(const BP& q, int row, i64 col, float val, float& red) {
  if (q.g == nullptr) return;
  if (q.period) atomicAdd(q.g + (i64)row * q.rs + col % q.period, val);
  else red += val;
}

struct EpiT0 {
  static constexpr int NRED = 0;
  static constexpr int SLOT0 = 0;
  const float* E0; const float* X; float* T0; i64 B;
  __device__ __forceinline__ void operator()(int row, i64 col, const float (&acc)[4], int nvalid, bool vec, float*) const {
    i64 off = (i64)row * B + col;
    Quad e = load4(E0, off, nvalid, vec), x = load4(X, off, nvalid, vec);
    float t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = fsub(fadd(acc[j], e.v[j]), x.v[j]);
    store4(T0, off, t, nvalid, vec);
  }
};

struct EpiZ {
  static constexpr int NRED = 0;
  static constexpr int SLOT0 = 0;
  const float* Zp; float* Zk; uint8_t* maskZ; BP th1; BP ss1; i64 B;
  __device__ __forceinline__ void operator()(int row, i64 col, const float (&acc)[4], int nvalid, bool vec, float*) const {
    i64 off = (i64)row * B + col;
    Quad zp = load4(Zp, off, nvalid, vec);
    float th[4]; bp_at4(th1, row, col, th);
    float s1 = ss1.p ? __ldg(ss1.p) : 1.f;
    float z[4]; unsigned bits[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float wv = ss1.p ? fmul(s1, acc[j]) : acc[j];
      z[j] = soft_act(fsub(zp.v[j], wv), th[j], bits[j]);
      bits[j] |= (z[j] > 0.f ? 4u : 0u) | (z[j] < 0.f ? 8u : 0u);      // sign(Z_k), see dladmm_problem.maskZ
    }
    store4(Zk, off, z, nvalid, vec);
    if (maskZ) store4_u8(maskZ, off, bits, nvalid, vec);
  }
};

template <int FAM>
struct EpiELT {
  static constexpr int NRED = 0;
  static constexpr int SLOT0 = 0;
  const float* X; const float* Ep; const float* Lp;
  float* Ek; float* Lk; float* Tn; uint8_t* maskE;
  BP b2, ss2, ss2_2, th2, bL; i64 B;
  __device__ __forceinline__ void operator()(int row, i64 col, const float (&acc)[4], int nvalid, bool vec, float*) const {
    i64 off = (i64)row * B + col;
    Quad x = load4(X, off, nvalid, vec), lp = load4(Lp, off, nvalid, vec);
    float e[4], t[4], l[4], vbL[4]; unsigned bits[4] = {0, 0, 0, 0};
    bp_at4(bL, row, col, vbL);
    if (FAM == DLADMM_FAMILY_B) {
      Quad ep = load4(Ep, off, nvalid, vec);
      float vb2[4], vs2[4], vth[4];
      bp_at4(b2, row, col, vb2); bp_at4(ss2, row, col, vs2); bp_at4(th2, row, col, vth);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float that = fsub(fadd(acc[j], ep.v[j]), x.v[j]);          // A Z + E - X
        float vvar = fadd(lp.v[j], fmul(vb2[j], that));            // L + beta2 * (...)
        float u = fsub(ep.v[j], fmul(vs2[j], vvar));               // E - ss2 * VVar
        e[j] = soft_act(u, vth[j], bits[j]);
      }
    } else if (FAM == DLADMM_FAMILY_A) {
      float vb2[4], vth[4];
      bp_at4(b2, row, col, vb2); bp_at4(th2, row, col, vth);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float u = fsub(fsub(x.v[j], acc[j]), fmul(vb2[j], lp.v[j]));  // X - A Z - beta2 * L
        e[j] = soft_act(u, vth[j], bits[j]);
      }
    } else {
      float v1[4], v2[4];
      bp_at4(ss2, row, col, v1); bp_at4(ss2_2, row, col, v2);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float res = fsub(x.v[j], acc[j]);                          // X - A Z
        e[j] = fsub(fmul(v1[j], res), fmul(v2[j], lp.v[j]));       // ss2_1*res - ss2_2*L
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      t[j] = fsub(fadd(acc[j], e[j]), x.v[j]);                     // T = A Z + E - X
      l[j] = fadd(lp.v[j], fmul(vbL[j], t[j]));                    // L = L + betaL * T
    }
    store4(Ek, off, e, nvalid, vec);
    store4(Tn, off, t, nvalid, vec);
    store4(Lk, off, l, nvalid, vec);
    if (FAM != DLADMM_FAMILY_C && maskE) store4_u8(maskE, off, bits, nvalid, vec);
  }
};

// ---- backward ------------------------------------------------------------------------------------
// Elementwise cotangent flow of layer j:  (dL_j, dT_{j+1}, dE_j) -> dR_j, carried dE_{j-1}, carried dL_{j-1}
// plus parameter-gradient contributions red[SL_BL..SL_B2].
struct M1Args {
  const float* Tn;      // T_{j+1}
  const float* Ek;      // E_j
  const float* Ep;      // E_{j-1} (E0 for j = 0)
  const float* Lp;      // L_{j-1} (L0 for j = 0)
  const uint8_t* maskE; // slab j (families A,B)
  const float* gE;      // upstream slabs for layer j (nullable)
  const float* gL;
  const float* gT;      // gT[j+1]
  BP bL, b2, ss2, ss2_2, th2;
  float* dR; float* cE; float* cL;   // outputs (m,B)
  i64 B;
  float lw; const float* lscale;     // fused L1-L1 loss: cotangent lw*scale*sign(E_j - T_{j+1}) on E_j, minus that on T_{j+1}
  int lkind;                         // 2: LASSO residual term 0.5*(E_j - T_{j+1})^2 -> cotangent lw*scale*(E_j - T_{j+1})
};

template <int FAM>
__device__ __forceinline__ void m1_quad(const M1Args& a, int row, i64 col, float (&dL)[4], float (&dE)[4],
                                        float (&dT)[4], int nvalid, bool vec, float* red) {
  i64 off = (i64)row * a.B + col;
  if (a.gL) { Quad g = load4(a.gL, off, nvalid, vec);
#pragma unroll
    for (int j = 0; j < 4; ++j) dL[j] += g.v[j]; }
  if (a.gE) { Quad g = load4(a.gE, off, nvalid, vec);
#pragma unroll
    for (int j = 0; j < 4; ++j) dE[j] += g.v[j]; }
  if (a.gT) { Quad g = load4(a.gT, off, nvalid, vec);
#pragma unroll
    for (int j = 0; j < 4; ++j) dT[j] += g.v[j]; }
  Quad tn = load4(a.Tn, off, nvalid, vec), lp = load4(a.Lp, off, nvalid, vec);
  Quad ek = load4(a.Ek, off, nvalid, vec);
  if (a.lscale) {
    const float sc = a.lw * __ldg(a.lscale);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float r = ek.v[j] - tn.v[j]; const float s = sc * (a.lkind == 2 ? r : sgn(r)); dE[j] += s; dT[j] -= s; }
  }
  float vbL[4]; bp_at4(a.bL, row, col, vbL);
  float dR[4], nE[4], nL[4];
  unsigned bits[4] = {0, 0, 0, 0};
  if (FAM != DLADMM_FAMILY_C) load4_u8(a.maskE, off, bits, nvalid, vec);
  if (FAM == DLADMM_FAMILY_B) {
    Quad ep = load4(a.Ep, off, nvalid, vec);
    float vb2[4], vs2[4];
    bp_at4(a.b2, row, col, vb2); bp_at4(a.ss2, row, col, vs2);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j >= nvalid) continue;
      pgrad(a.bL, row, col + j, dL[j] * tn.v[j], red[SL_BL]);
      float dTt = dT[j] + vbL[j] * dL[j];
      float dEt = dE[j] + dTt;
      float that = (tn.v[j] - ek.v[j]) + ep.v[j];            // A Z_j + E_{j-1} - X
      float q = lp.v[j] + vb2[j] * that;
      float mp = (bits[j] & 1u) ? 1.f : 0.f, mn = (bits[j] & 2u) ? 1.f : 0.f;
      float du = dEt * (mp + mn);
      pgrad(a.th2, row, col + j, dEt * (mn - mp), red[SL_TH2]);
      float dQ = -vs2[j] * du;
      pgrad(a.ss2, row, col + j, -du * q, red[SL_SS2]);
      float dThat = vb2[j] * dQ;
      pgrad(a.b2, row, col + j, dQ * that, red[SL_B2]);
      dR[j] = dTt + dThat;
      nE[j] = du + dThat;
      nL[j] = dL[j] + dQ;
    }
  } else if (FAM == DLADMM_FAMILY_A) {
    float vb2[4]; bp_at4(a.b2, row, col, vb2);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j >= nvalid) continue;
      pgrad(a.bL, row, col + j, dL[j] * tn.v[j], red[SL_BL]);
      float dTt = dT[j] + vbL[j] * dL[j];
      float dEt = dE[j] + dTt;
      float mp = (bits[j] & 1u) ? 1.f : 0.f, mn = (bits[j] & 2u) ? 1.f : 0.f;
      float du = dEt * (mp + mn);
      pgrad(a.th2, row, col + j, dEt * (mn - mp), red[SL_TH2]);
      pgrad(a.b2, row, col + j, -du * lp.v[j], red[SL_B2]);
      dR[j] = dTt - du;
      nE[j] = 0.f;
      nL[j] = dL[j] - vb2[j] * du;
    }
  } else {
    float v1[4], v2[4];
    bp_at4(a.ss2, row, col, v1); bp_at4(a.ss2_2, row, col, v2);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j >= nvalid) continue;
      pgrad(a.bL, row, col + j, dL[j] * tn.v[j], red[SL_BL]);
      float dTt = dT[j] + vbL[j] * dL[j];
      float dEt = dE[j] + dTt;
      pgrad(a.ss2, row, col + j, dEt * (ek.v[j] - tn.v[j]), red[SL_SS2]);   // X - A Z_j = E_j - T_{j+1}
      pgrad(a.ss2_2, row, col + j, -dEt * lp.v[j], red[SL_B2]);
      dR[j] = dTt - v1[j] * dEt;
      nE[j] = 0.f;
      nL[j] = dL[j] - v2[j] * dEt;
    }
  }
  store4(a.dR, off, dR, nvalid, vec);
  store4(a.cE, off, nE, nvalid, vec);
  store4(a.cL, off, nL, nvalid, vec);
}

struct EpiBG1 {
  static constexpr int NRED = 1;
  static constexpr int SLOT0 = SL_TH1;
  const float* gZ;        // upstream slab k (nullable)
  const float* cZin;      // carried from layer k+1 (nullable for the top layer)
  const uint8_t* maskZ;   // slab k
  BP th1;
  float* dx1;             // out (d,B): dZ_k * (m+ + m-), also the carried dZ_{k-1}
  i64 B;
  const float* Zk; float lz; const float* lscale;   // fused L1-L1 loss: cotangent lz*scale*sign(Z_k) on Z_k
  __device__ __forceinline__ void operator()(int row, i64 col, const float (&acc)[4], int nvalid, bool vec, float* red) const {
    i64 off = (i64)row * B + col;
    float dz[4] = {acc[0], acc[1], acc[2], acc[3]};
    if (lscale) { Quad z = load4(Zk, off, nvalid, vec); const float sc = lz * __ldg(lscale);
#pragma unroll
      for (int j = 0; j < 4; ++j) dz[j] += sc * sgn(z.v[j]); }
    if (gZ) { Quad g = load4(gZ, off, nvalid, vec);
#pragma unroll
      for (int j = 0; j < 4; ++j) dz[j] += g.v[j]; }
    if (cZin) { Quad g = load4(cZin, off, nvalid, vec);
#pragma unroll
      for (int j = 0; j < 4; ++j) dz[j] += g.v[j]; }
    unsigned bits[4]; load4_u8(maskZ, off, bits, nvalid, vec);
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float mp = (bits[j] & 1u) ? 1.f : 0.f, mn = (bits[j] & 2u) ? 1.f : 0.f;
      o[j] = dz[j] * (mp + mn);
      if (j < nvalid) pgrad(th1, row, col + j, dz[j] * (mn - mp), red[0]);
    }
    store4(dx1, off, o, nvalid, vec);
  }
};

template <int FAM>
struct EpiBG2 {
  static constexpr int NRED = 6;       // SL_BL..SL_B2 (layer k-1 via m1), SL_B1, SL_SS1 (layer k)
  static constexpr int SLOT0 = 0;
  const float* Lp;        // L_{k-1}
  const float* Tk;        // T_k
  BP b1, ss1;
  const float* cLin;      // carried dL_{k-1} so far (from m1 of layer k)
  const float* cEin;      // carried dE_{k-1}
  int has_prev;           // k > 0: run m1 for layer k-1
  M1Args prev;
  i64 B;
  __device__ __forceinline__ void operator()(int row, i64 col, const float (&acc)[4], int nvalid, bool vec, float* red) const {
    i64 off = (i64)row * B + col;
    Quad lp = load4(Lp, off, nvalid, vec), tk = load4(Tk, off, nvalid, vec);
    float vb1[4]; bp_at4(b1, row, col, vb1);
    float s1 = ss1.p ? __ldg(ss1.p) : 1.f;
    float dL[4], dT[4], dE[4];
    Quad cl = load4(cLin, off, nvalid, vec);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = lp.v[j] + vb1[j] * tk.v[j];            // V_k recomputed
      float dV = -s1 * acc[j];
      if (j < nvalid) {
        pgrad(b1, row, col + j, dV * tk.v[j], red[SL_B1]);
        if (ss1.g) red[SL_SS1] += -v * acc[j];
      }
      dL[j] = cl.v[j] + dV;
      dT[j] = vb1[j] * dV;
    }
    if (!has_prev) return;
    Quad ce = load4(cEin, off, nvalid, vec);
#pragma unroll
    for (int j = 0; j < 4; ++j) dE[j] = ce.v[j];
    m1_quad<FAM>(prev, row, col, dL, dE, dT, nvalid, vec, red);
  }
};

}  // namespace dladmm
