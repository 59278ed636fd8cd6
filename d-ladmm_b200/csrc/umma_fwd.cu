// tcgen05 forward schedule: T_0, then per layer W V + Z-prox and A Z + E/T/L (see umma_gemm.cuh / umma_epilogues.cuh).
#include "umma_host.cuh"

namespace dladmm {

bool umma_eligible(const dladmm_problem* p) {
  // TMA needs a 16-byte multiple pitch for the (rows x B) activations
  return p->precision != DLADMM_PREC_FP32 && (p->B % 4) == 0;
}

size_t umma_workspace_bytes(const dladmm_problem* p, int for_backward) {
  if (p->precision == DLADMM_PREC_FP32) return 0;
  return (for_backward ? ucarve_bwd(p, nullptr).bytes : ucarve(p, nullptr).bytes) + 1024;
}

// second stage of the fused metrics: out[k][i] = sum of the per-warp partial sums, one block per (layer, metric)
static __global__ void __launch_bounds__(256) metrics_reduce_kernel(const float* __restrict__ part, int nz, int ne, int ndg,
                                                                    uint32_t want, float* __restrict__ out) {
  const int k = blockIdx.x, i = blockIdx.y;
  if (!((want >> i) & 1u)) { if (threadIdx.x == 0) out[k * DLADMM_MET_COUNT + i] = 0.f; return; }
  const float* base = part + ((size_t)k * DLADMM_MET_COUNT + i) * OBJ_ENTRIES;
  const int n = (i == DLADMM_MET_L1_Z || i == DLADMM_MET_SQERR_Z) ? nz : (i == DLADMM_MET_DGAP_ATL ? ndg : ne);
  float s = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) s += base[j];
  s = warp_sum(s);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int j = 0; j < 8; ++j) v += sm[j];
    out[k * DLADMM_MET_COUNT + i] = v;
  }
}

// where metric i of layer k keeps its per-warp partial sums
static inline float* met_slot(const UWorkspace& w, int k, int i) { return w.metp + ((size_t)k * DLADMM_MET_COUNT + i) * OBJ_ENTRIES; }

template <int FAM, int NPASS, int PS, bool MET>
static int launch_elt(const dladmm_problem* p, const UWorkspace& w, const Slabs& s, int k, const void* Zk, float* Vnext,
                      __nv_bfloat16* Vnext_h, cudaStream_t st) {
  const dladmm_layer& l = p->layers[k];
  const dladmm_metrics* mt = p->metrics;
  umma::UEpiELT<FAM, PS, MET> epi;
  epi.X = p->X; epi.Ep = s.Ein(k); epi.Lp = s.Lin(k);
  epi.Ek = s.Eout(k); epi.Lk = s.Lout(k); epi.Tn = s.Tslab(k + 1); epi.maskE = s.mE(k);
  epi.b2 = make_bp(l.beta2); epi.ss2 = make_bp(l.ss2); epi.ss2_2 = make_bp(l.ss2_2); epi.th2 = make_bp(l.theta2);
  epi.bL = make_bp(betaL(p, l));
  epi.has_next = Vnext != nullptr || Vnext_h != nullptr;
  epi.b1n = make_bp(p->layers[k + 1 < p->K ? k + 1 : k].beta1);
  epi.V = Vnext; epi.B = p->B;
  epi.obj_part = p->objective ? w.objp + ((size_t)k * 2 + 1) * OBJ_ENTRIES : nullptr;
  epi.obj_kind = p->objective_kind;
  epi.Elabel = MET ? mt->E_label : nullptr; epi.Xclean = MET ? mt->X_clean : nullptr;
  epi.met_part = MET ? met_slot(w, k, DLADMM_MET_L1_RES) : nullptr;
  epi.met_stride = OBJ_ENTRIES;              // the ELT-side metrics are DLADMM_MET_L1_RES .. DLADMM_MET_DGAP_L, in order
  epi.Vh = (NPASS == 2 && Vnext_h) ? Vnext_h : nullptr; epi.ldh = w.ldh;
  if (NPASS == 2 && !p->Vsave) epi.V = nullptr;      // bf16 mode: the fp32 V is only kept for a backward
  const i64 nbt = (p->B + umma::TILE_B - 1) / umma::TILE_B;
  const int grid_e = (int)std::min<i64>(nbt * ((p->m + umma::TILE_N - 1) / umma::TILE_N), device_sm_count());
  return launch_umma<umma::UEpiELT<FAM, PS, MET>, NPASS>(DLADMM_KIND_GEMM_ELT, Zk, p->d, w.Ab, w.As, w.m256, w.dp, p->m, p->B, epi, st, grid_e, w.ldh);
}

template <int FAM, int NPASS, int PS>
static int forward_umma(const dladmm_problem* p, const UWorkspace& w, cudaStream_t st) {
  Slabs s(p);
  const int m = p->m, d = p->d, K = p->K;
  const i64 B = p->B;
  const dladmm_metrics* mt = p->metrics;
  const bool met = mt != nullptr && mt->want != 0;
  int rc;
  if ((rc = uprepare_weights<NPASS>(p, w, st, met && (mt->want >> DLADMM_MET_DGAP_ATL) & 1u))) return rc;
  // per-CTA partial sums of the fused objective / metrics: the reductions count grid x warps entries per layer; a CTA-pair launch
  // runs an even grid (one CTA fewer when that count is odd), so the entries start at zero
  if (p->objective && K > 0) DL_CUDA(cudaMemsetAsync(w.objp, 0, sizeof(float) * (size_t)K * 2 * OBJ_ENTRIES, st));
  if (met && K > 0) DL_CUDA(cudaMemsetAsync(w.metp, 0, sizeof(float) * (size_t)K * DLADMM_MET_COUNT * OBJ_ENTRIES, st));
  // V_k: one reused scratch slab, or every layer's kept for the backward (dladmm_problem.Vsave)
  auto Vslab = [&](int k) { return p->Vsave ? p->Vsave + s.ms * k : w.V; };
  const i64 nbt = (B + umma::TILE_B - 1) / umma::TILE_B;
  const int grid_z = (int)std::min<i64>(nbt * ((d + umma::TILE_N - 1) / umma::TILE_N), device_sm_count());
  const int grid_e = (int)std::min<i64>(nbt * ((m + umma::TILE_N - 1) / umma::TILE_N), device_sm_count());
  constexpr bool BF = NPASS == 2;
  auto to_bf16 = [&](const float* src, int rows, __nv_bfloat16* dst) -> int {
    { LaunchScope ls(DLADMM_KIND_PREP, st);
      to_bf16_kernel<<<(unsigned)(((i64)rows * B + 255) / 256), 256, 0, st>>>(src, rows, B, dst, w.ldh); }
    DL_CUDA(cudaGetLastError());
    return DLADMM_OK;
  };
  auto Zh = [&](int k) { return w.Zh + (size_t)(k & 1) * d * w.ldh; };      // bf16 mode: Z_k as the A Z operand (alternating)
  // T_0 = A Z0 + E0 - X (+ V_0), or T_0 given by the caller; with start_half the first Z-step does not exist and T_0 is unused
  if (!p->start_half) {
    if (p->T_init) {
      DL_CUDA(cudaMemcpyAsync(s.Tslab(0), p->T_init, sizeof(float) * (size_t)m * B, cudaMemcpyDeviceToDevice, st));
      if (K > 0) {
        const i64 quads = (B + 3) / 4;
        { LaunchScope ls(DLADMM_KIND_PREP, st);
          make_v_kernel<<<(unsigned)((quads * m + 255) / 256), 256, 0, st>>>(p->L0, p->T_init, make_bp(p->layers[0].beta1), m, B, Vslab(0)); }
        DL_CUDA(cudaGetLastError());
        if (BF && (rc = to_bf16(Vslab(0), m, w.Vh))) return rc;
      }
    } else {
      const void* z0op = p->Z0;
      if (BF) { if ((rc = to_bf16(p->Z0, d, Zh(1)))) return rc; z0op = Zh(1); }
      umma::UEpiT0<PS> epi{p->E0, p->X, p->L0, s.Tslab(0), K > 0 ? make_bp(p->layers[0].beta1) : BP{nullptr, nullptr, 0, 0},
                           (K > 0 && !(BF && !p->Vsave)) ? Vslab(0) : nullptr, B};
      epi.Vh = (BF && K > 0) ? w.Vh : nullptr; epi.ldh = w.ldh;
      if ((rc = launch_umma<umma::UEpiT0<PS>, NPASS>(DLADMM_KIND_GEMM_T0, z0op, d, w.Ab, w.As, w.m256, w.dp, m, B, epi, st, 0, w.ldh))) return rc;
    }
  } else if (BF) {
    if ((rc = to_bf16(p->Z0, d, Zh(0)))) return rc;                        // Z_0 := Z0 is the first A Z operand
  }
  for (int k = 0; k < K; ++k) {
    const dladmm_layer& l = p->layers[k];
    const size_t wi = (size_t)weight_index(p, k);
    const bool zstep = !(p->start_half && k == 0);
    const bool estep = !(p->stop_half && k == K - 1);
    // start_half: Z_0 is the caller's Z0 (read in place of Z slab 0, which is not written)
    const float* Zk = zstep ? s.Zout(k) : p->Z0;
    const float* Zprev = (k == 1 && p->start_half) ? p->Z0 : s.Zin(k);
    if (zstep) {
      umma::UEpiZ<PS> epi{Zprev, s.Zout(k), s.mZ(k), make_bp(l.theta1), make_bp(l.ss1), B};
      epi.obj_part = p->objective ? w.objp + (size_t)k * 2 * OBJ_ENTRIES : (met ? met_slot(w, k, DLADMM_MET_L1_Z) : nullptr);
      epi.Zlabel = met ? mt->Z_label : nullptr;
      epi.sq_part = (met && mt->Z_label) ? met_slot(w, k, DLADMM_MET_SQERR_Z) : nullptr;
      epi.Zh = (BF && estep) ? Zh(k) : nullptr; epi.ldh = w.ldh;
      // bf16 mode: weights are bf16 arrays of the same (d256 x mp) element count, at half the byte stride
      const void* wk = BF ? (const void*)((const __nv_bfloat16*)w.Wb + wi * w.d256 * w.mp) : (const void*)(w.Wb + wi * w.d256 * w.mp);
      const void* vop = BF ? (const void*)w.Vh : (const void*)Vslab(k);
      if ((rc = launch_umma<umma::UEpiZ<PS>, NPASS>(DLADMM_KIND_GEMM_Z, vop, m, wk, w.Ws + wi * w.d256 * w.mp, w.d256, w.mp, d, B, epi, st, grid_z, w.ldh)))
        return rc;
    }
    if (estep) {
      float* Vnext = k + 1 < K ? Vslab(k + 1) : nullptr;
      __nv_bfloat16* Vnext_h = (BF && k + 1 < K) ? w.Vh : nullptr;
      const void* zop = BF ? (const void*)Zh(k) : (const void*)Zk;
      rc = met ? launch_elt<FAM, NPASS, PS, true>(p, w, s, k, zop, Vnext, Vnext_h, st) : launch_elt<FAM, NPASS, PS, false>(p, w, s, k, zop, Vnext, Vnext_h, st);
      if (rc) return rc;
      if (met && ((mt->want >> DLADMM_MET_DGAP_ATL) & 1u)) {
        umma::UEpiDgap epi; epi.a = mt->dual_alpha; epi.part = met_slot(w, k, DLADMM_MET_DGAP_ATL);
        const void* lop = s.Lout(k);
        if (BF) { if ((rc = to_bf16(s.Lout(k), m, w.Lh))) return rc; lop = w.Lh; }
        if ((rc = launch_umma<umma::UEpiDgap, NPASS>(DLADMM_KIND_METRIC_GEMM, lop, m, w.Atb, w.Ats, w.d256, w.mp, d, B, epi, st, grid_z, w.ldh))) return rc;
      }
    }
  }
  if (p->objective && K > 0) {
    { LaunchScope ls(DLADMM_KIND_OBJECTIVE, st);
      objective_reduce_kernel<<<K, 256, 0, st>>>(w.objp, grid_z * umma::UEpiZ<PS>::WARPS, grid_e * umma::UEpiELT<FAM, PS>::WARPS, p->objective_alpha,
                                                p->objective); }
    DL_CUDA(cudaGetLastError());
  }
  if (met && K > 0) {
    { LaunchScope ls(DLADMM_KIND_OBJECTIVE, st);
      metrics_reduce_kernel<<<dim3(K, DLADMM_MET_COUNT), 256, 0, st>>>(w.metp, grid_z * umma::UEpiZ<PS>::WARPS, grid_e * umma::UEpiELT<FAM, PS>::WARPS,
                                                                       grid_z * umma::UEpiDgap::WARPS, mt->want, mt->out); }
    DL_CUDA(cudaGetLastError());
  }
  return DLADMM_OK;
}

int umma_forward(const dladmm_problem* p, void* ws_base, cudaStream_t st) {
  char* base = (char*)(((uintptr_t)ws_base + 1023) & ~(uintptr_t)1023);
  UWorkspace w = ucarve(p, base);
  const bool x3 = p->precision == DLADMM_PREC_TF32X3, bf = p->precision == DLADMM_PREC_BF16, mix = p->precision == DLADMM_PREC_TF32_BF16X2;
  const int pm = param_mode(p);
#define DL_PM(FN, F, NP)                                                                                  \
  (pm == umma::PM_SCALAR ? FN<F, NP, umma::PM_SCALAR>(p, w, st)                                           \
                         : pm == umma::PM_ROWS ? FN<F, NP, umma::PM_ROWS>(p, w, st) : FN<F, NP, umma::PM_GENERAL>(p, w, st))
  {
    const int rc = umma_forward_persistent(p, w, st);           // the all-layer schedule (umma_pfwd.cu), when the call is eligible
    if (rc != DLADMM_PF_NOT_TAKEN) return rc;
  }
#define DL_FWD(F) (x3 ? DL_PM(forward_umma, F, 3) : mix ? DL_PM(forward_umma, F, 4) : bf ? DL_PM(forward_umma, F, 2) : DL_PM(forward_umma, F, 1))
  switch (p->family) {
    case DLADMM_FAMILY_A: return DL_FWD(DLADMM_FAMILY_A);
    case DLADMM_FAMILY_B: return DL_FWD(DLADMM_FAMILY_B);
    default: return DL_FWD(DLADMM_FAMILY_C);
  }
#undef DL_FWD
}

}  // namespace dladmm
