// tcgen05 backward schedule: per layer A^T dR (+ Z mask), dW -= s1 dx1 V^T, W^T dx1 (+ the cotangent flow of the layer below).
#include "umma_host.cuh"
#include "umma_bwd.cuh"

namespace dladmm {

template <int NPASS>
static int launch_nt(const float* P, int M, const float* Q, int N, i64 B, const float* s1ptr, float* C, int ldc, cudaStream_t st) {
  constexpr int KC = NPASS >= 3 ? 16 : 32;
  using Plan = umma::NtPlan<NPASS, KC>;
  const CUtensorMapSwizzle sw = KC == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap tP, tQ;
  int rc;
  if ((rc = umma::make_tmap_2d(&tP, P, M, B, B, KC, umma::NT_PROWS, sw))) return rc;
  if ((rc = umma::make_tmap_2d(&tQ, Q, N, B, B, KC, umma::TILE_N, sw))) return rc;
  const int mt = (M + umma::NT_PROWS - 1) / umma::NT_PROWS, nt = (N + umma::TILE_N - 1) / umma::TILE_N;
  int split = std::max(1, device_sm_count() / (mt * nt));
  i64 chunk = round_up64((B + split - 1) / split, KC);
  split = (int)((B + chunk - 1) / chunk);
  umma::NtShape ns;
  ns.M = M; ns.N = N; ns.B = B; ns.chunk = chunk; ns.ldc = ldc;
  ns.acc_unit = umma::acc_comp_unit() * (float)umma::mma_per_chunk(NPASS);
  auto kern = umma::umma_nt_kernel<NPASS, KC>;
  static bool attr_set[MAX_DEVICES] = {false};
  const int dev = current_device_index();
  if (!attr_set[dev]) {
    DL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::TOTAL));
    attr_set[dev] = true;
  }
  {
    LaunchScope ls(DLADMM_KIND_BWD_GEMM_DW, st);
    kern<<<dim3(mt, split, nt), umma::roles_threads(umma::NT_EPI_WARPS), Plan::TOTAL, st>>>(tP, tQ, ns, s1ptr, -1.f, C);
  }
  DL_CUDA(cudaGetLastError());
  return DLADMM_OK;
}

template <int FAM, int NPASS, int PS>
static int backward_umma(const dladmm_problem* p, const dladmm_cotangents* g, const Workspace& sw, const UBwdWorkspace& w,
                         cudaStream_t st) {
  Slabs s(p);
  const int m = p->m, d = p->d, K = p->K;
  const i64 B = p->B;
  int rc;
  // transposed, split weights
  {
    SplitJobs jobs; jobs.n = 1;
    jobs.j[0].src = p->A; jobs.j[0].big = w.Atb; jobs.j[0].small = w.Ats;
    dim3 grid((w.mp + 31) / 32, (w.d256 + 31) / 32, 1);
    { LaunchScope ls(DLADMM_KIND_PREP, st); prep_split_t_kernel<NPASS><<<grid, 256, 0, st>>>(jobs, d, m, w.d256, w.mp); }
    DL_CUDA(cudaGetLastError());
    std::vector<const float*> uniq = WeightMap(p).uniq;
    for (size_t base = 0; base < uniq.size(); base += 32) {
      SplitJobs wj; wj.n = (int)std::min<size_t>(32, uniq.size() - base);
      for (int i = 0; i < wj.n; ++i) {
        size_t idx = base + i;
        wj.j[i].src = uniq[idx];
        wj.j[i].big = w.Wtb + idx * (size_t)w.m256 * w.dp;
        wj.j[i].small = w.Wts + idx * (size_t)w.m256 * w.dp;
      }
      dim3 g2((w.dp + 31) / 32, (w.m256 + 31) / 32, wj.n);
      { LaunchScope ls(DLADMM_KIND_PREP, st); prep_split_t_kernel<NPASS><<<g2, 256, 0, st>>>(wj, m, d, w.m256, w.dp); }
      DL_CUDA(cudaGetLastError());
    }
  }
  // top layer: elementwise cotangent flow with nothing carried (FFMA-path kernel), then split dR
  {
    M1Args a = make_m1(p, g, sw, K - 1);
    dim3 grid((sw.ncolTiles + 7) / 8, m);
    { LaunchScope ls(DLADMM_KIND_BWD_ELEM, st); m1_kernel<FAM><<<grid, 256, 0, st>>>(m, B, a, sw.part, sw.ncolTiles, sw.prow); }
    DL_CUDA(cudaGetLastError());
    ReduceJobs jobs; jobs.n = 0;
    add_m1_jobs(p, jobs, p->layers[K - 1]);
    if ((rc = launch_reduce(jobs, sw, st))) return rc;
  }
  // one grid for both activation-side products so that the per-warp partial entries line up
  const i64 nbt = (B + umma::TILE_B - 1) / umma::TILE_B;
  const i64 tiles_dz = nbt * ((d + umma::TILE_N - 1) / umma::TILE_N), tiles_dv = nbt * ((m + umma::TILE_N - 1) / umma::TILE_N);
  const int grid = (int)std::min<i64>(std::max(tiles_dz, tiles_dv), std::min(device_sm_count(), 256));
  umma::RedOut ro;
  // per-warp partial entries (all-scalar parameters): the two epilogues use different warp counts, so entries are laid
  // out for the larger one and the buffer is cleared once -- a kernel rewrites only its own entries every layer
  ro.part = w.part; ro.nentries = grid * umma::MAX_EPI_WARPS; ro.ngroups = w.ngroups; ro.prow = w.prow;
  // all-scalar parameters: every layer keeps its own block of entries and ONE launch reduces them all after the loop
  // (15 reduction launches of ~20 us each were 3 % of a training step)
  const size_t layer_entries = (size_t)SL_COUNT * ro.nentries;
  std::vector<ScalarJob> sjobs;
  constexpr bool SC = PS == umma::PM_SCALAR;      // all parameters scalar: per-warp entries, one reduction launch at the end
  if (SC) DL_CUDA(cudaMemsetAsync(w.part, 0, sizeof(float) * layer_entries * K, st));
  // per-row parameters: the partial blocks of up to red_group layers are kept side by side and summed by ONE launch (a launch per
  // layer was 20 latency-bound launches of ~160 blocks, 23 us each: 0.46 ms of the `full` K = 20 step)
  const i64 row_block = (i64)SL_COUNT * w.ngroups * w.prow;
  ReduceJobs jobs; jobs.n = 0;
  int pending = 0;
  for (int k = K - 1; k >= 0; --k) {
    const dladmm_layer& l = p->layers[k];
    const size_t wi = (size_t)weight_index(p, k);
    if (SC) ro.part = w.part + layer_entries * k;
    const i64 part_off = SC ? 0 : row_block * pending;
    if (!SC) ro.part = w.part + part_off;
    {
      umma::UEpiBG1<PS> epi;
      epi.gZ = g->gZ ? g->gZ + s.zs * k : nullptr;
      epi.cZin = k == K - 1 ? nullptr : sw.cZ;
      epi.maskZ = s.mZ(k);
      epi.th1 = make_bp(l.theta1);
      epi.dx1 = sw.cZ; epi.ro = ro; epi.B = B;
      const bool lossz = g->loss_kind != 0;
      epi.Zk = s.Zout(k); epi.lz = lossz ? g->loss_alpha * g->loss_layer_weight[k] : 0.f; epi.lscale = lossz ? g->loss_scale : nullptr;
      if ((rc = launch_umma<umma::UEpiBG1<PS>, NPASS>(DLADMM_KIND_BWD_GEMM_DZ, sw.dR, m, w.Atb, w.Ats, w.d256, w.mp, d, B, epi, st, grid)))
        return rc;
    }
    if (l.gW) {
      const float* Vk = p->Vsave ? p->Vsave + s.ms * k : w.V;    // kept by the forward, or recomputed here
      if (!p->Vsave) {
        const i64 quads = (B + 3) / 4;
        { LaunchScope ls(DLADMM_KIND_PREP, st);
          make_v_kernel<<<(unsigned)((quads * m + 255) / 256), 256, 0, st>>>(s.Lin(k), s.Tslab(k), make_bp(l.beta1), m, B, w.V); }
        DL_CUDA(cudaGetLastError());
      }
      if ((rc = launch_nt<NPASS>(sw.cZ, d, Vk, m, B, l.ss1.ptr, l.gW, m, st))) return rc;
    }
    // gW of layer k is complete up to here (dladmm_cotangents.layer_events): a data-parallel caller starts its allreduce now
    if (g->layer_events && g->layer_events[k]) DL_CUDA(cudaEventRecord((cudaEvent_t)g->layer_events[k], st));
    {
      umma::UEpiBG2<FAM, PS> epi;
      epi.Lp = s.Lin(k); epi.Tk = s.Tslab(k);
      epi.b1 = make_bp(l.beta1); epi.ss1 = make_bp(l.ss1);
      epi.cLin = sw.cL; epi.cEin = sw.cE;
      epi.has_prev = k > 0;
      const int j = k > 0 ? k - 1 : 0;
      const dladmm_layer& lj = p->layers[j];
      epi.Ek = s.Eout(j); epi.Ep = s.Ein(j); epi.Lpp = s.Lin(j); epi.maskE = s.mE(j);
      epi.gE = g->gE ? g->gE + s.ms * j : nullptr;
      epi.gL = g->gL ? g->gL + s.ms * j : nullptr;
      epi.gT = g->gT ? g->gT + s.ms * (j + 1) : nullptr;
      epi.bL = make_bp(betaL(p, lj)); epi.b2 = make_bp(lj.beta2); epi.ss2 = make_bp(lj.ss2); epi.ss2_2 = make_bp(lj.ss2_2);
      epi.th2 = make_bp(lj.theta2);
      epi.dR = sw.dR; epi.cE = sw.cE; epi.cL = sw.cL;
      epi.ro = ro; epi.B = B;
      epi.lw = g->loss_kind != 0 ? g->loss_layer_weight[j] : 0.f; epi.lscale = g->loss_kind != 0 ? g->loss_scale : nullptr;
      epi.lkind = g->loss_kind;
      if ((rc = launch_umma<umma::UEpiBG2<FAM, PS>, NPASS>(DLADMM_KIND_BWD_GEMM_DV, sw.cZ, d, w.Wtb + wi * w.m256 * w.dp, w.Wts + wi * w.m256 * w.dp, w.m256, w.dp, m, B, epi, st, grid)))
        return rc;
    }
    const int first = jobs.n;
    add_job(jobs, SL_TH1, l.theta1, d, part_off);
    add_job(jobs, SL_B1, l.beta1, m, part_off);
    add_job(jobs, SL_SS1, l.ss1, m, part_off);
    if (k > 0) add_m1_jobs(p, jobs, p->layers[k - 1], part_off);
    if (SC) {
      for (int i = first; i < jobs.n; ++i) sjobs.push_back(ScalarJob{ro.part + (size_t)jobs.j[i].slot * ro.nentries, jobs.j[i].grad});
      jobs.n = 0;
      continue;
    }
    if (++pending < w.red_group && k > 0 && jobs.n + 7 <= MAX_REDUCE_JOBS) continue;      // keep collecting
    pending = 0;
    if (jobs.n) {
      int ncol = ro.ngroups, prow = ro.prow;
      int maxrows = 1;
      for (int i = 0; i < jobs.n; ++i)
        if (!jobs.j[i].scalar) maxrows = std::max(maxrows, jobs.j[i].rows);
      dim3 rg((maxrows + 7) / 8, jobs.n);
      { LaunchScope ls(DLADMM_KIND_BWD_REDUCE, st); reduce_partials_kernel<<<rg, 256, 0, st>>>(jobs, w.part, ncol, prow); }
      DL_CUDA(cudaGetLastError());
      jobs.n = 0;
    }
  }
  for (size_t base = 0; base < sjobs.size(); base += 120) {
    ScalarJobs sj;
    sj.n = (int)std::min<size_t>(120, sjobs.size() - base);
    sj.count = ro.nentries;
    for (int i = 0; i < sj.n; ++i) sj.j[i] = sjobs[base + i];
    { LaunchScope ls(DLADMM_KIND_BWD_REDUCE, st); reduce_scalar_entries_kernel<<<sj.n, 256, 0, st>>>(sj); }
    DL_CUDA(cudaGetLastError());
  }
  return DLADMM_OK;
}

int umma_backward(const dladmm_problem* p, const dladmm_cotangents* g, const Workspace& sw, void* ws_base, cudaStream_t st) {
  char* base = (char*)(((uintptr_t)ws_base + 1023) & ~(uintptr_t)1023);
  UBwdWorkspace w = ucarve_bwd(p, base);
  const bool x3 = p->precision == DLADMM_PREC_TF32X3, mix = p->precision == DLADMM_PREC_TF32_BF16X2;
  const int pm = param_mode(p);
#define DL_BPM(F, NP)                                                                                                  \
  (pm == umma::PM_SCALAR ? backward_umma<F, NP, umma::PM_SCALAR>(p, g, sw, w, st)                                      \
                         : pm == umma::PM_ROWS ? backward_umma<F, NP, umma::PM_ROWS>(p, g, sw, w, st)                  \
                                               : backward_umma<F, NP, umma::PM_GENERAL>(p, g, sw, w, st))
#define DL_BWD(F) (x3 ? DL_BPM(F, 3) : mix ? DL_BPM(F, 4) : DL_BPM(F, 1))
  switch (p->family) {
    case DLADMM_FAMILY_A: return DL_BWD(DLADMM_FAMILY_A);
    case DLADMM_FAMILY_B: return DL_BWD(DLADMM_FAMILY_B);
    default: return DL_BWD(DLADMM_FAMILY_C);
  }
#undef DL_BWD
}

}  // namespace dladmm
