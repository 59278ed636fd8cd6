// All-layer persistent forward (umma_persist.cuh): eligibility, tile shape, launch.  Its own translation unit: the kernel has
// 45 instantiations (3 families x 3 parameter modes x {tf32x3, tf32_bf16x2, tf32} x {256-row, 32-row tiles}) and compiles in parallel
// with the per-layer schedule of umma_fwd.cu.
#include "umma_host.cuh"

namespace dladmm {

// ---- all-layer persistent forward (umma_persist.cuh) --------------------------------------------------------------------
// out[k] = alpha * sum(Z partial sums of layer k) + sum(E/T/L partial sums of layer k); the per-unit entries of a layer are contiguous
static __global__ void __launch_bounds__(256) pf_objective_reduce_kernel(const float* __restrict__ part, long long units_t0, long long units_z,
                                                                         long long units_e, float alpha, float* __restrict__ out) {
  const long long u0 = units_t0 + (long long)blockIdx.x * (units_z + units_e);
  const float* pz = part + u0 * 8;
  const float* pe = pz + units_z * 8;
  float sz = 0.f, se = 0.f;
  for (long long i = threadIdx.x; i < units_z * 8; i += 256) sz += pz[i];
  for (long long i = threadIdx.x; i < units_e * 8; i += 256) se += pe[i];
  float s = warp_sum(alpha * sz + se);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += sm[i];
    out[blockIdx.x] = v;
  }
}

// Debugging aid (DLADMM_PF_TRACE=1): the persistent kernel's first CTAs stamp clock64 per unit into a buffer the library
// allocates once for that purpose (the one exception to "never allocates", and only in this mode); dladmm_debug_trace reads it.
static long long* g_pf_trace = nullptr;
static long long* pf_trace_buffer() {
  const char* e = getenv("DLADMM_PF_TRACE");
  if (!(e && e[0] == '1')) return nullptr;
  const size_t bytes = sizeof(long long) * umma::PF_TRACE_CTAS * umma::PF_TRACE_UNITS * 8;
  if (!g_pf_trace && cudaMalloc(&g_pf_trace, bytes) != cudaSuccess) { g_pf_trace = nullptr; return nullptr; }
  cudaMemset(g_pf_trace, 0, bytes);
  return g_pf_trace;
}
int pf_trace_read(long long* host_out, long long capacity) {
  const long long n = (long long)umma::PF_TRACE_CTAS * umma::PF_TRACE_UNITS * 8;
  if (!g_pf_trace || !host_out || capacity < n) return -1;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpy(host_out, g_pf_trace, sizeof(long long) * n, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int)n;
}

// DLADMM_NO_PERSISTENT=1: always run the per-layer kernels (A/B measurements, and the tests that cover both schedules)
static bool persistent_allowed() {
  const char* e = getenv("DLADMM_NO_PERSISTENT");
  return !(e && e[0] == '1');
}

static bool persistent_eligible(const dladmm_problem* p) {
  if (!persistent_allowed()) return false;
  if (p->precision != DLADMM_PREC_TF32X3 && p->precision != DLADMM_PREC_TF32 && p->precision != DLADMM_PREC_TF32_BF16X2) return false;
  if (p->K < 1 || p->K > umma::PF_MAX_LAYERS) return false;
  // Tensor-bound shapes gain nothing from the merged schedule (their per-layer launches already run ~90 % of the tensor peak and
  // the one-launch version measured 5 % slower at m 1000, d 2000: 50.3 vs 47.7 ms per 32 768-column forward on the same box);
  // DLADMM_PERSISTENT=1 forces it for any shape
  {
    const char* e = getenv("DLADMM_PERSISTENT");
    if (!(e && e[0] == '1') && (i64)p->m * p->d > (i64)512 * 1024) return false;
  }
  if (p->start_half || p->stop_half || p->T_init) return false;
  if (p->metrics && p->metrics->want) return false;
  {   // the kernel numbers its units with 32 bits
    const i64 nbt = (p->B + umma::TILE_B - 1) / umma::TILE_B;
    const i64 per = (p->d + 31) / 32 + (p->m + 31) / 32;       // (an upper bound: 32-row tiles)
    if (nbt * per * (p->K + 1) >= ((i64)1 << 31)) return false;
  }
  return true;
}

// Feature rows per tile of the persistent forward.  256 (one tcgen05.mma of N = 256 per k-step) is the throughput shape.  With few
// batch tiles a stage is one or two 256-row units: 146 SMs idle and each unit walks its 16-32 k-chunks through a 3-stage ring one
// TMA latency at a time (~26 us per stage whatever the batch).  32-row tiles spread a stage over 8x as many CTAs, each with a
// 7-stage ring (20 KB per stage) and an 8x shorter epilogue.  DLADMM_PF_TN=32|256 forces either.
static int pf_tile_rows(const dladmm_problem* p) {
  if (p->precision != DLADMM_PREC_TF32X3 && p->precision != DLADMM_PREC_TF32_BF16X2) return umma::TILE_N;
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("DLADMM_PF_TN"); forced = e ? atoi(e) : 0; }
  const i64 nbt = (p->B + umma::TILE_B - 1) / umma::TILE_B;
  const i64 units256 = nbt * ((p->m + umma::TILE_N - 1) / umma::TILE_N);     // units of an A Z stage at 256 rows per tile
  if (units256 * 4 > 1024) return umma::TILE_N;                                // (the workspace is sized for 32-row tiles up to here: ucarve)
  if (forced == 32 || forced == umma::TILE_N) return forced;
  return units256 * 4 <= device_sm_count() ? 32 : umma::TILE_N;
}

static umma::BPc to_bpc(const dladmm_bparam& q) { umma::BPc b; b.p = q.ptr; b.rs = q.row_stride; b.period = q.col_period; return b; }

template <int FAM, int NPASS, int PS, int TN>
static int forward_persistent_tn(const dladmm_problem* p, const UWorkspace& w, cudaStream_t st) {
  constexpr int KC = KChunk<NPASS>::value;
  using Plan = umma::SmemPlan<NPASS, KC, TN>;
  Slabs s(p);
  const int m = p->m, d = p->d, K = p->K;
  const i64 B = p->B;
  int rc;
  if ((rc = uprepare_weights<NPASS>(p, w, st, false))) return rc;
  DL_CUDA(cudaMemsetAsync(w.pf_flags, 0, w.pf_flag_bytes, st));

  umma::PfParams* pp = new umma::PfParams;          // ~9 KB: built on the heap, copied into the launch
  umma::PfMaps* mp = new umma::PfMaps;
  struct Guard { umma::PfParams* a; umma::PfMaps* b; ~Guard() { delete a; delete b; } } guard{pp, mp};
  memset(pp, 0, sizeof(*pp));
  pp->m = m; pp->d = d; pp->K = K; pp->last_only = p->last_only;
  pp->B = B; pp->n_btiles = (B + umma::TILE_B - 1) / umma::TILE_B;
  pp->tn = TN;
  pp->nt_z = (d + TN - 1) / TN; pp->nt_e = (m + TN - 1) / TN;
  pp->kc_z = (m + KC - 1) / KC; pp->kc_e = (d + KC - 1) / KC;
  pp->acc_scale_z = umma::acc_comp_scale(pp->kc_z * umma::mma_per_chunk(NPASS));
  pp->acc_scale_e = umma::acc_comp_scale(pp->kc_e * umma::mma_per_chunk(NPASS));
  pp->units_t0 = pp->n_btiles * pp->nt_e; pp->units_z = pp->n_btiles * pp->nt_z; pp->units_e = pp->n_btiles * pp->nt_e;
  pp->total_units = pp->units_t0 + (i64)K * (pp->units_z + pp->units_e);
  pp->X = p->X; pp->E0 = p->E0; pp->L0 = p->L0; pp->Z0 = p->Z0;
  pp->Z = p->Z; pp->E = p->E; pp->L = p->L; pp->T = p->T;
  pp->V = p->Vsave ? p->Vsave : w.V; pp->v_per_layer = p->Vsave ? 1 : 0;
  pp->maskZ = p->maskZ; pp->maskE = p->maskE;
  pp->zs = s.zs; pp->ms = s.ms;
  pp->obj_kind = p->objective_kind;
  pp->obj_part = p->objective ? w.pf_obj : nullptr;
  pp->flags = w.pf_flags;
  pp->spin_limit = 4000000000ll;                    // ~2 s of SM clocks
  // DLADMM_PF_PREFETCH=D: L2 prefetch distance of the staging producer in chunks (0 = off)
  { const char* e = getenv("DLADMM_PF_PREFETCH"); pp->prefetch = e ? atoi(e) : 0; }
  { const char* e = getenv("DLADMM_PF_SCOUT_SLEEP"); pp->scout_sleep_ns = e ? (unsigned)atoi(e) : 0u; }
  pp->trace = pf_trace_buffer();
  { const char* e = getenv("DLADMM_PF_XRESIDENT"); pp->x_resident = (e && e[0] == '0') ? 0 : ((i64)m * B * 4 <= (i64)96 << 20); }
  // two CTA sets half a layer period apart once every SM has several tiles of each half (see PfParams::nstreams)
  {
    static int skew_us = -2;
    if (skew_us == -2) { const char* e = getenv("DLADMM_PF_SKEW_US"); skew_us = e ? atoi(e) : -1; }
    const int sms = device_sm_count();
    // (measured on C1: two skewed sets are 3-6 % SLOWER than one list -- the E/T/L epilogue is bound per SM, not by chip-wide HBM
    //  contention, see profiles/r02_persistent_experiments.md -- so this stays an experiment switch: DLADMM_PF_SKEW_US > 0)
    const bool two = skew_us > 0 && (sms % 2) == 0 && pp->n_btiles >= 2 * (i64)sms;
    pp->nstreams = two ? 2 : 1;
    pp->split = (pp->n_btiles + 1) / 2;
    // default skew: the W V phase of one half on half the SMs, ~ tiles per CTA x 12 us, in SM clocks (1.9 GHz)
    const double tiles_per_cta = (double)(pp->split * pp->nt_z) / (sms / 2);
    const double us = skew_us > 0 ? (double)skew_us : tiles_per_cta * (NPASS == 3 ? 12.0 : 8.0) * ((double)m / 250.0);
    pp->skew = two ? (long long)(us * 1900.0) : 0;
  }
  const WeightMap wm(p);
  for (int k = 0; k < K; ++k) {
    const dladmm_layer& l = p->layers[k];
    umma::PfLayer& y = pp->layer[k];
    y.b1 = to_bpc(l.beta1); y.b2 = to_bpc(l.beta2); y.bL = to_bpc(betaL(p, l)); y.ss1 = to_bpc(l.ss1);
    y.ss2 = to_bpc(l.ss2); y.ss2_2 = to_bpc(l.ss2_2); y.th1 = to_bpc(l.theta1); y.th2 = to_bpc(l.theta2);
    y.widx = wm.idx[k];
  }
  const int depth = p->last_only ? 2 : K;
  const CUtensorMapSwizzle wsw = KC == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const CUtensorMapSwizzle asw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, nsw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if ((rc = umma::make_tmap_2d(&mp->A_big, w.Ab, w.m256, w.dp, w.dp, KC, TN, wsw))) return rc;
  if ((rc = umma::make_tmap_2d(&mp->A_small, NPASS >= 3 ? w.As : w.Ab, w.m256, w.dp, w.dp, KC, TN, wsw))) return rc;
  if ((rc = umma::make_tmap_3d(&mp->W_big, w.Wb, w.nW, w.d256, w.mp, w.mp, (i64)w.d256 * w.mp, KC, TN, wsw))) return rc;
  if ((rc = umma::make_tmap_3d(&mp->W_small, NPASS >= 3 ? w.Ws : w.Wb, w.nW, w.d256, w.mp, w.mp, (i64)w.d256 * w.mp, KC, TN, wsw))) return rc;
  if ((rc = umma::make_tmap_2d(&mp->actZ0, p->Z0, d, B, B, 32, KC, asw))) return rc;
  if ((rc = umma::make_tmap_3d(&mp->actZ, p->Z, depth, d, B, B, s.zs, 32, KC, asw))) return rc;
  if ((rc = umma::make_tmap_3d(&mp->actV, pp->V, p->Vsave ? K : 1, m, B, B, s.ms, 32, KC, asw))) return rc;
  if ((rc = umma::make_tmap_2d(&mp->sE0, p->E0, m, B, B, umma::TILE_B, 8, nsw))) return rc;
  if ((rc = umma::make_tmap_2d(&mp->sX, p->X, m, B, B, umma::TILE_B, 8, nsw))) return rc;
  if ((rc = umma::make_tmap_2d(&mp->sL0, p->L0, m, B, B, umma::TILE_B, 8, nsw))) return rc;
  if ((rc = umma::make_tmap_3d(&mp->sE, p->E, depth, m, B, B, s.ms, umma::TILE_B, 8, nsw))) return rc;
  if ((rc = umma::make_tmap_3d(&mp->sL, p->L, depth, m, B, B, s.ms, umma::TILE_B, 8, nsw))) return rc;
  if ((rc = umma::make_tmap_2d(&mp->sZ0, p->Z0, d, B, B, umma::TILE_B, 16, nsw))) return rc;
  if ((rc = umma::make_tmap_3d(&mp->sZ, p->Z, depth, d, B, B, s.zs, umma::TILE_B, 16, nsw))) return rc;

  auto kern = umma::umma_forward_persistent_kernel<FAM, PS, NPASS, KC, TN>;
  static bool attr_set[MAX_DEVICES] = {false};
  const int dev = current_device_index();
  if (!attr_set[dev]) {
    DL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (umma::pf_smem_total<NPASS, KC, TN>())));
    attr_set[dev] = true;
  }
  const int grid = pp->nstreams == 2 ? device_sm_count() : (int)std::min<i64>(std::max(pp->units_z, pp->units_e), device_sm_count());
  {
    LaunchScope ls(DLADMM_KIND_FWD_PERSISTENT, st);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(umma::roles_threads(8) + 32); cfg.dynamicSmemBytes = (umma::pf_smem_total<NPASS, KC, TN>()); cfg.stream = st;
    // the CTAs wait on one another's counters: a cooperative launch makes the driver guarantee (or refuse) that the whole grid is
    // resident at once, whatever else shares the device
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    DL_CUDA(cudaLaunchKernelEx(&cfg, kern, *mp, *pp));
  }
  DL_CUDA(cudaGetLastError());
  if (p->objective) {
    { LaunchScope ls(DLADMM_KIND_OBJECTIVE, st);
      pf_objective_reduce_kernel<<<K, 256, 0, st>>>(w.pf_obj, pp->units_t0, pp->units_z, pp->units_e, p->objective_alpha, p->objective); }
    DL_CUDA(cudaGetLastError());
  }
  return DLADMM_OK;
}

template <int FAM, int NPASS, int PS>
static int forward_persistent(const dladmm_problem* p, const UWorkspace& w, cudaStream_t st) {
  if constexpr (NPASS >= 3) {
    if (pf_tile_rows(p) == 32) return forward_persistent_tn<FAM, NPASS, PS, 32>(p, w, st);
  }
  return forward_persistent_tn<FAM, NPASS, PS, umma::TILE_N>(p, w, st);
}

// entry point of umma_fwd.cu: DLADMM_OK / error, or DLADMM_PF_NOT_TAKEN when the call is not eligible for the one-launch schedule
int umma_forward_persistent(const dladmm_problem* p, const UWorkspace& w, cudaStream_t st) {
  if (!persistent_eligible(p)) return DLADMM_PF_NOT_TAKEN;
  const bool x3 = p->precision == DLADMM_PREC_TF32X3, mix = p->precision == DLADMM_PREC_TF32_BF16X2;
  const int pm = param_mode(p);
#define DL_PM(FN, F, NP)                                                                                  \
  (pm == umma::PM_SCALAR ? FN<F, NP, umma::PM_SCALAR>(p, w, st)                                           \
                         : pm == umma::PM_ROWS ? FN<F, NP, umma::PM_ROWS>(p, w, st) : FN<F, NP, umma::PM_GENERAL>(p, w, st))
#define DL_PF(F) (x3 ? DL_PM(forward_persistent, F, 3) : mix ? DL_PM(forward_persistent, F, 4) : DL_PM(forward_persistent, F, 1))
  switch (p->family) {
    case DLADMM_FAMILY_A: return DL_PF(DLADMM_FAMILY_A);
    case DLADMM_FAMILY_B: return DL_PF(DLADMM_FAMILY_B);
    default: return DL_PF(DLADMM_FAMILY_C);
  }
#undef DL_PF
#undef DL_PM
}

}  // namespace dladmm
