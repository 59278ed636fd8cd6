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
  return launch_umma<umma::UEpiELT<FAM, PS, MET>, NPASS>(DLADMM_KIND_GEMM_ELT, Zk, p->d, w.Ab, w.As, w.m256, w.dp, p->m, p->B, epi, st, 0, w.ldh);
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
      if ((rc = launch_umma<umma::UEpiZ<PS>, NPASS>(DLADMM_KIND_GEMM_Z, vop, m, wk, w.Ws + wi * w.d256 * w.mp, w.d256, w.mp, d, B, epi, st, 0, w.ldh)))
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
        if ((rc = launch_umma<umma::UEpiDgap, NPASS>(DLADMM_KIND_METRIC_GEMM, lop, m, w.Atb, w.Ats, w.d256, w.mp, d, B, epi, st, 0, w.ldh))) return rc;
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

int umma_forward(const dladmm_problem* p, void* ws_base, cudaStream_t st) {
  char* base = (char*)(((uintptr_t)ws_base + 1023) & ~(uintptr_t)1023);
  UWorkspace w = ucarve(p, base);
  const bool x3 = p->precision == DLADMM_PREC_TF32X3, bf = p->precision == DLADMM_PREC_BF16, mix = p->precision == DLADMM_PREC_TF32_BF16X2;
  const int pm = param_mode(p);
#define DL_PM(FN, F, NP)                                                                                  \
  (pm == umma::PM_SCALAR ? FN<F, NP, umma::PM_SCALAR>(p, w, st)                                           \
                         : pm == umma::PM_ROWS ? FN<F, NP, umma::PM_ROWS>(p, w, st) : FN<F, NP, umma::PM_GENERAL>(p, w, st))
  if (persistent_eligible(p)) {
#define DL_PF(F) (x3 ? DL_PM(forward_persistent, F, 3) : mix ? DL_PM(forward_persistent, F, 4) : DL_PM(forward_persistent, F, 1))
    switch (p->family) {
      case DLADMM_FAMILY_A: return DL_PF(DLADMM_FAMILY_A);
      case DLADMM_FAMILY_B: return DL_PF(DLADMM_FAMILY_B);
      default: return DL_PF(DLADMM_FAMILY_C);
    }
#undef DL_PF
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
