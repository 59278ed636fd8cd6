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

template <int FAM, int NPASS, bool PS>
static int forward_umma(const dladmm_problem* p, const UWorkspace& w, cudaStream_t st) {
  Slabs s(p);
  const int m = p->m, d = p->d;
  const i64 B = p->B;
  int rc;
  if ((rc = uprepare_weights<NPASS>(p, w, st))) return rc;
  // V_k: one reused scratch slab, or every layer's kept for the backward (dladmm_problem.Vsave)
  auto Vslab = [&](int k) { return p->Vsave ? p->Vsave + s.ms * k : w.V; };
  const i64 nbt = (B + umma::TILE_B - 1) / umma::TILE_B;
  const int grid_z = (int)std::min<i64>(nbt * ((d + umma::TILE_N - 1) / umma::TILE_N), device_sm_count());
  const int grid_e = (int)std::min<i64>(nbt * ((m + umma::TILE_N - 1) / umma::TILE_N), device_sm_count());
  // T_0 = A Z0 + E0 - X (+ V_0), or T_0 given by the caller
  if (p->T_init) {
    DL_CUDA(cudaMemcpyAsync(s.Tslab(0), p->T_init, sizeof(float) * (size_t)m * B, cudaMemcpyDeviceToDevice, st));
    const i64 quads = (B + 3) / 4;
    { LaunchScope ls(DLADMM_KIND_PREP, st);
      make_v_kernel<<<(unsigned)((quads * m + 255) / 256), 256, 0, st>>>(p->L0, p->T_init, make_bp(p->layers[0].beta1), m, B, Vslab(0)); }
    DL_CUDA(cudaGetLastError());
  } else {
    umma::UEpiT0<PS> epi{p->E0, p->X, p->L0, s.Tslab(0), make_bp(p->layers[0].beta1), Vslab(0), B};
    if ((rc = launch_umma<umma::UEpiT0<PS>, NPASS>(DLADMM_KIND_GEMM_T0, p->Z0, d, w.Ab, w.As, w.m256, w.dp, m, B, epi, st))) return rc;
  }
  for (int k = 0; k < p->K; ++k) {
    const dladmm_layer& l = p->layers[k];
    const size_t wi = (size_t)weight_index(p, k);
    {
      umma::UEpiZ<PS> epi{s.Zin(k), s.Zout(k), s.mZ(k), make_bp(l.theta1), make_bp(l.ss1), B};
      epi.obj_part = p->objective ? w.objp + (size_t)k * 2 * OBJ_ENTRIES : nullptr;
      if ((rc = launch_umma<umma::UEpiZ<PS>, NPASS>(DLADMM_KIND_GEMM_Z, Vslab(k), m, w.Wb + wi * w.d256 * w.mp, w.Ws + wi * w.d256 * w.mp, w.d256, w.mp, d, B, epi, st)))
        return rc;
    }
    {
      umma::UEpiELT<FAM, PS> epi;
      epi.X = p->X; epi.Ep = s.Ein(k); epi.Lp = s.Lin(k);
      epi.Ek = s.Eout(k); epi.Lk = s.Lout(k); epi.Tn = s.Tslab(k + 1); epi.maskE = s.mE(k);
      epi.b2 = make_bp(l.beta2); epi.ss2 = make_bp(l.ss2); epi.ss2_2 = make_bp(l.ss2_2); epi.th2 = make_bp(l.theta2);
      epi.bL = make_bp(betaL(p, l));
      epi.has_next = k + 1 < p->K;
      epi.b1n = make_bp(p->layers[k + 1 < p->K ? k + 1 : k].beta1);
      epi.V = Vslab(k + 1 < p->K ? k + 1 : k); epi.B = B;
      epi.obj_part = p->objective ? w.objp + ((size_t)k * 2 + 1) * OBJ_ENTRIES : nullptr;
      if ((rc = launch_umma<umma::UEpiELT<FAM, PS>, NPASS>(DLADMM_KIND_GEMM_ELT, s.Zout(k), d, w.Ab, w.As, w.m256, w.dp, m, B, epi, st)))
        return rc;
    }
  }
  if (p->objective) {
    { LaunchScope ls(DLADMM_KIND_OBJECTIVE, st);
      objective_reduce_kernel<<<p->K, 256, 0, st>>>(w.objp, grid_z * umma::UEpiZ<PS>::WARPS, grid_e * umma::UEpiELT<FAM, PS>::WARPS, p->objective_alpha,
                                                   p->objective); }
    DL_CUDA(cudaGetLastError());
  }
  return DLADMM_OK;
}

int umma_forward(const dladmm_problem* p, void* ws_base, cudaStream_t st) {
  char* base = (char*)(((uintptr_t)ws_base + 1023) & ~(uintptr_t)1023);
  UWorkspace w = ucarve(p, base);
  const bool x3 = p->precision == DLADMM_PREC_TF32X3;
  const bool ps = all_params_scalar(p);
#define DL_FWD(F)                                                                                       \
  (x3 ? (ps ? forward_umma<F, 3, true>(p, w, st) : forward_umma<F, 3, false>(p, w, st))                  \
      : (ps ? forward_umma<F, 1, true>(p, w, st) : forward_umma<F, 1, false>(p, w, st)))
  switch (p->family) {
    case DLADMM_FAMILY_A: return DL_FWD(DLADMM_FAMILY_A);
    case DLADMM_FAMILY_B: return DL_FWD(DLADMM_FAMILY_B);
    default: return DL_FWD(DLADMM_FAMILY_C);
  }
#undef DL_FWD
}

}  // namespace dladmm
