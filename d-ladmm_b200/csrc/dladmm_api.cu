// C-ABI entry points of libdladmm.so (see include/dladmm.h) and the per-layer launch schedule.
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <vector>

#include "common.cuh"
#include "epilogues.cuh"
#include "simt_gemm.cuh"
#include "host_common.cuh"

namespace dladmm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- launch accounting ---------------------------------------------------------------------------
struct ProfRec { int kind; cudaEvent_t a, b; };
static std::atomic<long long> g_launches{0};   // the profiling hooks below are single-threaded measurement aids
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;

LaunchScope::LaunchScope(int kind_, cudaStream_t st_) : kind(kind_), st(st_), rec(nullptr) {
  ++g_launches;
  if (g_prof_on) {
    ProfRec r; r.kind = kind;
    if (cudaEventCreate(&r.a) == cudaSuccess && cudaEventCreate(&r.b) == cudaSuccess) {
      cudaEventRecord(r.a, st);
      g_prof.push_back(r);
      rec = (void*)(size_t)g_prof.size();
    }
  }
}
LaunchScope::~LaunchScope() {
  if (rec) cudaEventRecord(g_prof[(size_t)rec - 1].b, st);
  static int dbg = -1;          // DLADMM_DEBUG_SYNC=1: synchronise and report after every launch (debugging aid)
  if (dbg < 0) { const char* e = getenv("DLADMM_DEBUG_SYNC"); dbg = (e && e[0] == '1') ? 1 : 0; }
  if (dbg) {
    cudaError_t e = cudaStreamSynchronize(st);
    fprintf(stderr, "[dladmm] launch kind %d -> %s\n", kind, cudaGetErrorString(e));
    fflush(stderr);
  }
}

static int validate(const dladmm_problem* p, int for_backward) {
  DL_REQUIRE(p != nullptr, "problem is NULL");
  DL_REQUIRE(p->abi_version == DLADMM_ABI_VERSION, "ABI version mismatch: got %d, library is %d", p->abi_version,
             DLADMM_ABI_VERSION);
  DL_REQUIRE(p->family >= 0 && p->family <= 2, "unknown family %d", p->family);
  DL_REQUIRE(p->B < ((int64_t)1 << 28), "batch of %lld columns: the kernels address rows inside a chunk with 32 bits (B < 2^28)", (long long)p->B);
  DL_REQUIRE(p->precision >= 0 && p->precision <= 4, "unknown precision %d", p->precision);
  DL_REQUIRE(p->m > 0 && p->d > 0 && p->K >= 0, "m, d must be positive and K non-negative (m=%d d=%d K=%d)", p->m, p->d, p->K);
  DL_REQUIRE(p->K > 0 || !for_backward, "backward needs K > 0");
  DL_REQUIRE(p->B >= 0, "B must be non-negative");
  DL_REQUIRE(p->layers != nullptr || p->K == 0, "layers is NULL");
  DL_REQUIRE((p->start_half == 0 || p->start_half == 1) && (p->stop_half == 0 || p->stop_half == 1), "start_half / stop_half must be 0 or 1");
  DL_REQUIRE(!(p->start_half || p->stop_half) || (!for_backward && !p->last_only && p->K > 0),
             "half-layer calls are forward-only, need K > 0 and every iterate (last_only = 0)");
  DL_REQUIRE(!(p->start_half && p->T_init), "start_half does not use T_init");
  DL_REQUIRE(p->objective_kind >= 0 && p->objective_kind <= 2, "unknown objective_kind %d", p->objective_kind);
  if (p->metrics && p->metrics->want) {
    const dladmm_metrics* mt = p->metrics;
    DL_REQUIRE(!for_backward, "metrics are computed by the forward");
    DL_REQUIRE(mt->out != nullptr && p->objective == nullptr, "metrics need `out` and exclude `objective`");
    DL_REQUIRE(mt->want < (1u << DLADMM_MET_COUNT), "unknown metric bits %u", mt->want);
    DL_REQUIRE(!((mt->want >> DLADMM_MET_SQERR_Z) & 1u) || mt->Z_label, "DLADMM_MET_SQERR_Z needs Z_label");
    DL_REQUIRE(!((mt->want >> DLADMM_MET_SQERR_E) & 1u) || mt->E_label, "DLADMM_MET_SQERR_E needs E_label");
    DL_REQUIRE(!((mt->want >> DLADMM_MET_SQERR_AZ) & 1u) || mt->X_clean, "DLADMM_MET_SQERR_AZ needs X_clean");
    DL_REQUIRE(!(p->start_half || p->stop_half), "metrics are not offered for half-layer calls");
  }
  DL_REQUIRE(p->A && (p->B == 0 || (p->X && p->Z0 && p->E0 && p->L0 && p->T && (p->K == 0 || (p->Z && p->E && p->L)))),
             "A, X, Z0, E0, L0, Z, E, L, T must be non-NULL");
  for (int k = 0; k < p->K; ++k) {
    const dladmm_layer& l = p->layers[k];
    DL_REQUIRE(l.W != nullptr, "layer %d: W is NULL", k);
    DL_REQUIRE(l.beta1.ptr && l.theta1.ptr, "layer %d: beta1/theta1 missing", k);
    if (p->family == DLADMM_FAMILY_A) DL_REQUIRE(l.beta2.ptr && l.theta2.ptr, "layer %d: family A needs beta2, theta2", k);
    if (p->family == DLADMM_FAMILY_B)
      DL_REQUIRE(l.beta2.ptr && l.beta3.ptr && l.ss2.ptr && l.theta2.ptr, "layer %d: family B needs beta2, beta3, ss2, theta2", k);
    if (p->family == DLADMM_FAMILY_C)
      DL_REQUIRE(l.beta3.ptr && l.ss2.ptr && l.ss2_2.ptr, "layer %d: family C needs beta3, ss2(_1), ss2_2", k);
    if (l.ss1.ptr) DL_REQUIRE(l.ss1.row_stride == 0 && l.ss1.col_period == 0, "layer %d: ss1 must be a (1,1) scalar", k);
  }
  if (for_backward) {
    DL_REQUIRE(p->last_only == 0, "backward needs every iterate (last_only must be 0)");
    DL_REQUIRE(p->maskZ != nullptr, "backward needs maskZ from the forward");
    DL_REQUIRE(p->family == DLADMM_FAMILY_C || p->maskE != nullptr, "backward needs maskE from the forward");
  }
  if (p->B > 0) {
    Workspace w = carve(p, for_backward);
    if (p->workspace == nullptr || p->workspace_bytes < w.bytes) {
      set_error("workspace too small: need %zu bytes, got %zu", w.bytes, p->workspace_bytes);
      return DLADMM_ERR_WORKSPACE;
    }
  }
  return DLADMM_OK;
}

static int check_device() {
  int dev = 0;
  DL_CUDA(cudaGetDevice(&dev));
  int major = 0;
  DL_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    set_error("device %d has compute capability %d.x; this library is sm_100a only and has no fallback", dev, major);
    return DLADMM_ERR_DEVICE;
  }
  return DLADMM_OK;
}

// launch helpers --------------------------------------------------------------------------------------
template <class BLoad, class Epi>
static int launch_simt(int kind, int M, i64 N, int Kd, const float* Aw, int lda, const BLoad& bl, const Epi& epi, float* part,
                       int ncolTiles, int prow, cudaStream_t st) {
  dim3 grid((unsigned)((N + SG_BN - 1) / SG_BN), (unsigned)((M + SG_BM - 1) / SG_BM));
  {
    LaunchScope ls(kind, st);
    simt_gemm_kernel<BLoad, Epi><<<grid, SG_THREADS, 0, st>>>(M, N, Kd, Aw, lda, bl, epi, part, ncolTiles, prow);
  }
  DL_CUDA(cudaGetLastError());
  return DLADMM_OK;
}

static int prepare_weights(const dladmm_problem* p, const Workspace& w, bool transposed, cudaStream_t st) {
  // A (m x d): normal copy -> Ap (m x dp); transposed -> Atp (d x mp)
  {
    PrepJobs jobs; jobs.n = 1;
    jobs.j[0].src = p->A;
    jobs.j[0].dst_n = transposed ? nullptr : w.Ap;
    jobs.j[0].dst_t = transposed ? w.Atp : nullptr;
    int R = p->m, C = p->d, ldn = w.dp, ldt = w.mp;
    dim3 grid((std::max(C, ldn) + 31) / 32, (std::max(R, ldt) + 31) / 32, 1);
    { LaunchScope ls(DLADMM_KIND_PREP, st); prep_weights_kernel<<<grid, 256, 0, st>>>(jobs, R, C, ldn, ldt); }
    DL_CUDA(cudaGetLastError());
  }
  // W_k (d x m): normal -> Wp (d x mp); transposed -> Wtp (m x dp)
  std::vector<const float*> uniq = WeightMap(p).uniq;
  for (size_t base = 0; base < uniq.size(); base += 32) {
    PrepJobs jobs; jobs.n = (int)std::min<size_t>(32, uniq.size() - base);
    for (int i = 0; i < jobs.n; ++i) {
      size_t idx = base + i;
      jobs.j[i].src = uniq[idx];
      jobs.j[i].dst_n = transposed ? nullptr : w.Wp + idx * (size_t)p->d * w.mp;
      jobs.j[i].dst_t = transposed ? w.Wtp + idx * (size_t)p->m * w.dp : nullptr;
    }
    int R = p->d, C = p->m, ldn = w.mp, ldt = w.dp;
    dim3 grid((std::max(C, ldn) + 31) / 32, (std::max(R, ldt) + 31) / 32, jobs.n);
    { LaunchScope ls(DLADMM_KIND_PREP, st); prep_weights_kernel<<<grid, 256, 0, st>>>(jobs, R, C, ldn, ldt); }
    DL_CUDA(cudaGetLastError());
  }
  return DLADMM_OK;
}

// ---- forward ----------------------------------------------------------------------------------------
template <int FAM>
static int forward_simt(const dladmm_problem* p, const Workspace& w, cudaStream_t st) {
  Slabs s(p);
  const int m = p->m, d = p->d;
  const i64 B = p->B;
  int rc;
  // T_0 = A Z0 + E0 - X (or given by the caller); a call that starts with an E/T/L-step (start_half) has no T_0
  if (p->start_half) {
  } else if (p->T_init) {
    DL_CUDA(cudaMemcpyAsync(s.Tslab(0), p->T_init, sizeof(float) * (size_t)m * B, cudaMemcpyDeviceToDevice, st));
  } else {
    BPlain bl{p->Z0, B};
    EpiT0 epi{p->E0, p->X, s.Tslab(0), B};
    if ((rc = launch_simt(DLADMM_KIND_GEMM_T0, m, B, d, w.Ap, w.dp, bl, epi, nullptr, 0, 0, st))) return rc;
  }
  for (int k = 0; k < p->K; ++k) {
    const dladmm_layer& l = p->layers[k];
    const float* Wk = w.Wp + (size_t)weight_index(p, k) * d * w.mp;
    const bool zstep = !(p->start_half && k == 0), estep = !(p->stop_half && k == p->K - 1);
    const float* Zk = zstep ? s.Zout(k) : p->Z0;                       // start_half: Z_0 := Z0 as given
    if (zstep) {
      BVar bl{s.Lin(k), s.Tslab(k), make_bp(l.beta1), B};
      EpiZ epi{(k == 1 && p->start_half) ? p->Z0 : s.Zin(k), s.Zout(k), s.mZ(k), make_bp(l.theta1), make_bp(l.ss1), B};
      if ((rc = launch_simt(DLADMM_KIND_GEMM_Z, d, B, m, Wk, w.mp, bl, epi, nullptr, 0, 0, st))) return rc;
    }
    if (estep) {
      BPlain bl{Zk, B};
      EpiELT<FAM> epi{p->X, s.Ein(k), s.Lin(k), s.Eout(k), s.Lout(k), s.Tslab(k + 1), s.mE(k),
                      make_bp(l.beta2), make_bp(l.ss2), make_bp(l.ss2_2), make_bp(l.theta2), make_bp(betaL(p, l)), B};
      if ((rc = launch_simt(DLADMM_KIND_GEMM_ELT, m, B, d, w.Ap, w.dp, bl, epi, nullptr, 0, 0, st))) return rc;
    }
  }
  return DLADMM_OK;
}

// ---- backward ---------------------------------------------------------------------------------------
template <int FAM>
static int backward_simt(const dladmm_problem* p, const dladmm_cotangents* g, const Workspace& w, cudaStream_t st) {
  Slabs s(p);
  const int m = p->m, d = p->d, K = p->K;
  const i64 B = p->B;
  int rc;
  // top layer: elementwise cotangent flow with nothing carried
  {
    M1Args a = make_m1(p, g, w, K - 1);
    dim3 grid((w.ncolTiles + 7) / 8, m);
    { LaunchScope ls(DLADMM_KIND_BWD_ELEM, st); m1_kernel<FAM><<<grid, 256, 0, st>>>(m, B, a, w.part, w.ncolTiles, w.prow); }
    DL_CUDA(cudaGetLastError());
    ReduceJobs jobs; jobs.n = 0;
    add_m1_jobs(p, jobs, p->layers[K - 1]);
    if ((rc = launch_reduce(jobs, w, st))) return rc;
  }
  for (int k = K - 1; k >= 0; --k) {
    const dladmm_layer& l = p->layers[k];
    const int wi = weight_index(p, k);
    // BG1: dZ_k = gZ_k + carried + A^T dR ; dx1 -> cZ
    {
      BPlain bl{w.dR, B};
      const bool lossz = g->loss_kind != 0 && g->loss_scale && g->loss_layer_weight;
      EpiBG1 epi{g->gZ ? g->gZ + s.zs * k : nullptr, k == K - 1 ? nullptr : w.cZ, s.mZ(k), make_bp(l.theta1), w.cZ, B,
                 s.Zout(k), lossz ? g->loss_alpha * g->loss_layer_weight[k] : 0.f, lossz ? g->loss_scale : nullptr};
      if ((rc = launch_simt(DLADMM_KIND_BWD_GEMM_DZ, d, B, m, w.Atp, w.mp, bl, epi, w.part, w.ncolTiles, w.prow, st))) return rc;
    }
    // BG3: gW -= s1 * dx1 * V_k^T
    if (l.gW) {
      BVar ql{s.Lin(k), s.Tslab(k), make_bp(l.beta1), B};
      int tiles = ((d + NT_BM - 1) / NT_BM) * ((m + NT_BN - 1) / NT_BN);
      int split = std::max(1, std::min<int>((592 + tiles - 1) / tiles, (int)((B + 255) / 256)));
      i64 chunk = round_up64((B + split - 1) / split, NT_BK);
      split = (int)((B + chunk - 1) / chunk);
      dim3 grid((d + NT_BM - 1) / NT_BM, (m + NT_BN - 1) / NT_BN, split);
      { LaunchScope ls(DLADMM_KIND_BWD_GEMM_DW, st);
        simt_gemm_nt_kernel<BVar><<<grid, NT_THREADS, 0, st>>>(d, m, B, chunk, w.cZ, ql, l.ss1.ptr, -1.f, l.gW, m); }
      DL_CUDA(cudaGetLastError());
    }
    // BG2: dV = -s1 W^T dx1 ; carried dL, dT ; fused elementwise part of layer k-1
    {
      BPlain bl{w.cZ, B};
      EpiBG2<FAM> epi;
      epi.Lp = s.Lin(k); epi.Tk = s.Tslab(k);
      epi.b1 = make_bp(l.beta1); epi.ss1 = make_bp(l.ss1);
      epi.cLin = w.cL; epi.cEin = w.cE;
      epi.has_prev = k > 0;
      epi.prev = make_m1(p, g, w, k > 0 ? k - 1 : 0);
      epi.B = B;
      const float* Wt = w.Wtp + (size_t)wi * m * w.dp;
      if ((rc = launch_simt(DLADMM_KIND_BWD_GEMM_DV, m, B, d, Wt, w.dp, bl, epi, w.part, w.ncolTiles, w.prow, st))) return rc;
    }
    ReduceJobs jobs; jobs.n = 0;
    add_job(jobs, SL_TH1, l.theta1, d);
    add_job(jobs, SL_B1, l.beta1, m);
    add_job(jobs, SL_SS1, l.ss1, m);
    if (k > 0) add_m1_jobs(p, jobs, p->layers[k - 1]);
    if ((rc = launch_reduce(jobs, w, st))) return rc;
  }
  return DLADMM_OK;
}

}  // namespace dladmm

namespace dladmm {

// ---- objective --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) objective_kernel(const float* __restrict__ Z, const float* __restrict__ E,
                                                        const float* __restrict__ T, i64 zs, i64 ms, float alpha, int kind,
                                                        float* __restrict__ out) {
  // grid.y = layer k; out[k] += alpha*sum|Z_k| + sum|E_k - T_{k+1}|   (kind 2: 0.5*(E_k - T_{k+1})^2 in place of the L1 residual)
  const int k = blockIdx.y;
  const float* z = Z + zs * k;
  const float* e = E + ms * k;
  const float* t = T + ms * (k + 1);
  float s = 0.f;
  for (i64 i = (i64)blockIdx.x * 256 + threadIdx.x; i < zs; i += (i64)gridDim.x * 256) s += alpha * fabsf(z[i]);
  for (i64 i = (i64)blockIdx.x * 256 + threadIdx.x; i < ms; i += (i64)gridDim.x * 256) {
    const float r = e[i] - t[i];
    s += kind == 2 ? 0.5f * r * r : fabsf(r);
  }
  __shared__ float sm[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = sm[threadIdx.x];
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) atomicAdd(out + k, v);
  }
}

}  // namespace dladmm

using namespace dladmm;

static int run_objective(const dladmm_problem* p, float alpha, int kind, float* out, cudaStream_t st);

extern "C" {

const char* dladmm_last_error(void) { return g_err; }

int64_t dladmm_launch_count(void) { return g_launches.load(); }

int dladmm_profile_start(void) {
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_prof_on = true;
  return DLADMM_OK;
}

int dladmm_profile_stop(double* ms_by_kind, int64_t* launches_by_kind) {
  g_prof_on = false;
  if (ms_by_kind) for (int i = 0; i < DLADMM_KIND_COUNT; ++i) ms_by_kind[i] = 0.0;
  if (launches_by_kind) for (int i = 0; i < DLADMM_KIND_COUNT; ++i) launches_by_kind[i] = 0;
  int rc = DLADMM_OK;
  for (auto& r : g_prof) {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(r.b);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.a, r.b);
    if (e != cudaSuccess) { set_error("profile event failed: %s", cudaGetErrorString(e)); rc = DLADMM_ERR_CUDA; }
    if (r.kind >= 0 && r.kind < DLADMM_KIND_COUNT) {
      if (ms_by_kind) ms_by_kind[r.kind] += ms;
      if (launches_by_kind) launches_by_kind[r.kind] += 1;
    }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof.clear();
  return rc;
}

int64_t dladmm_debug_trace(int64_t* host_out, int64_t capacity) { return pf_trace_read((long long*)host_out, (long long)capacity); }

int dladmm_query(int device, dladmm_caps* caps) {
  if (!caps) { set_error("caps is NULL"); return DLADMM_ERR_INVALID; }
  memset(caps, 0, sizeof(*caps));
  caps->abi_version = DLADMM_ABI_VERSION;
  cudaDeviceProp prop;
  DL_CUDA(cudaGetDeviceProperties(&prop, device));
  caps->cc_major = prop.major;
  caps->cc_minor = prop.minor;
  caps->sm_count = prop.multiProcessorCount;
  caps->supported = prop.major == 10;
  caps->has_tcgen05 = 1;
  caps->total_mem = (int64_t)prop.totalGlobalMem;
  return DLADMM_OK;
}

size_t dladmm_workspace_bytes(const dladmm_problem* p, int for_backward) {
  if (!p || p->K < 0 || (p->K > 0 && !p->layers)) return 0;
  dladmm_problem q = *p;
  q.workspace = nullptr;
  Workspace w = carve(&q, for_backward);
  size_t extra = umma_workspace_bytes(&q, for_backward);
  return w.bytes + extra;
}

int dladmm_forward(const dladmm_problem* p, void* stream) {
  int rc = validate(p, 0);
  if (rc) return rc;
  if (p->B == 0) {
    if (p->objective) DL_CUDA(cudaMemsetAsync(p->objective, 0, sizeof(float) * p->K, (cudaStream_t)stream));
    if (p->metrics && p->metrics->want && p->metrics->out)
      DL_CUDA(cudaMemsetAsync(p->metrics->out, 0, sizeof(float) * p->K * DLADMM_MET_COUNT, (cudaStream_t)stream));
    return DLADMM_OK;
  }
  if (p->workspace_bytes < dladmm_workspace_bytes(p, 0)) {
    set_error("workspace too small: need %zu bytes, got %zu", dladmm_workspace_bytes(p, 0), p->workspace_bytes);
    return DLADMM_ERR_WORKSPACE;
  }
  if ((rc = check_device())) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w = carve(p, 0);
  // tensor-core precisions need a 16-byte pitch for TMA (B % 4 == 0); other batch sizes run the FFMA kernels,
  // which are fp32 throughout (never less accurate than the precision asked for)
  if (umma_eligible(p)) return umma_forward(p, (char*)p->workspace + w.bytes, st);
  if (p->metrics && p->metrics->want) {
    set_error("metrics need a tensor-core precision (the FFMA path has no fused metric epilogues)");
    return DLADMM_ERR_INVALID;
  }
  if (p->objective && p->last_only) {
    set_error("objective with last_only needs a tensor-core precision and B %% 4 == 0 (the FFMA path reads the iterates back)");
    return DLADMM_ERR_INVALID;
  }
  if ((rc = prepare_weights(p, w, false, st))) return rc;
  switch (p->family) {
    case DLADMM_FAMILY_A: rc = forward_simt<DLADMM_FAMILY_A>(p, w, st); break;
    case DLADMM_FAMILY_B: rc = forward_simt<DLADMM_FAMILY_B>(p, w, st); break;
    default: rc = forward_simt<DLADMM_FAMILY_C>(p, w, st); break;
  }
  if (rc == DLADMM_OK && p->objective) rc = run_objective(p, p->objective_alpha, p->objective_kind, p->objective, st);   // Vsave is unused on this path
  return rc;
}

int dladmm_backward(const dladmm_problem* p, const dladmm_cotangents* g, void* stream) {
  int rc = validate(p, 1);
  if (rc) return rc;
  if (!g) { set_error("cotangents is NULL"); return DLADMM_ERR_INVALID; }
  if (g->loss_kind < 0 || g->loss_kind > 2) { set_error("unknown loss_kind %d", g->loss_kind); return DLADMM_ERR_INVALID; }
  if (g->loss_kind != 0 && (!g->loss_layer_weight || !g->loss_scale)) {
    set_error("a fused loss needs loss_layer_weight (host) and loss_scale (device)");
    return DLADMM_ERR_INVALID;
  }
  if (p->B == 0) return DLADMM_OK;
  if ((rc = check_device())) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w = carve(p, 1);
  if (p->workspace_bytes < dladmm_workspace_bytes(p, 1)) {
    set_error("workspace too small: need %zu bytes, got %zu", dladmm_workspace_bytes(p, 1), p->workspace_bytes);
    return DLADMM_ERR_WORKSPACE;
  }
  if (umma_eligible(p)) return umma_backward(p, g, w, (char*)p->workspace + w.bytes, st);
  if ((rc = prepare_weights(p, w, true, st))) return rc;
  switch (p->family) {
    case DLADMM_FAMILY_A: rc = backward_simt<DLADMM_FAMILY_A>(p, g, w, st); break;
    case DLADMM_FAMILY_B: rc = backward_simt<DLADMM_FAMILY_B>(p, g, w, st); break;
    default: rc = backward_simt<DLADMM_FAMILY_C>(p, g, w, st); break;
  }
  if (rc == DLADMM_OK && g->layer_events)       // this path makes no per-layer promise: everything is final after its last kernel
    for (int k = p->K - 1; k >= 0; --k)
      if (g->layer_events[k]) DL_CUDA(cudaEventRecord((cudaEvent_t)g->layer_events[k], st));
  return rc;
}

int dladmm_objective(const dladmm_problem* p, float alpha, float* out, void* stream) {
  DL_REQUIRE(p && out, "problem/out is NULL");
  DL_REQUIRE(p->last_only == 0, "objective needs every iterate (last_only must be 0)");
  DL_REQUIRE(p->Z && p->E && p->T && p->K > 0, "Z, E, T must be non-NULL");
  int rc;
  if ((rc = check_device())) return rc;
  return run_objective(p, alpha, p->objective_kind, out, (cudaStream_t)stream);
}

}  // extern "C"

static int run_objective(const dladmm_problem* p, float alpha, int kind, float* out, cudaStream_t st) {
  DL_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * p->K, st));
  if (p->B == 0) return DLADMM_OK;
  i64 zs = (i64)p->d * p->B, ms = (i64)p->m * p->B;
  int bx = (int)std::min<i64>((zs + 255) / 256, 592);
  { LaunchScope ls(DLADMM_KIND_OBJECTIVE, st); objective_kernel<<<dim3(bx, p->K), 256, 0, st>>>(p->Z, p->E, p->T, zs, ms, alpha, kind, out); }
  DL_CUDA(cudaGetLastError());
  return DLADMM_OK;
}
