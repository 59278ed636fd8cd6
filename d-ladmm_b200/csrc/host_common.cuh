// Host-side helpers shared by the translation units of libdladmm.so: workspace carving, slab addressing, the weight tie
// map and the parameter-gradient reduction jobs.  (dladmm_api.cu = C ABI + FFMA schedule, umma_fwd.cu / umma_bwd.cu = the
// tcgen05 schedules; split so that nvcc compiles them in parallel.)
#pragma once
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "epilogues.cuh"
#include "simt_gemm.cuh"

namespace dladmm {

// ---- workspace carving -------------------------------------------------------------------------
struct Workspace {
  // forward
  float* Ap;    // (m x dp)      A, K-major padded
  float* Wp;    // nW x (d x mp) W_k, K-major padded
  // backward
  float* Atp;   // (d x mp)      A^T
  float* Wtp;   // nW x (m x dp) W_k^T
  float* cZ;    // (d x B) carried dZ / dx1
  float* cE;    // (m x B)
  float* cL;    // (m x B)
  float* dR;    // (m x B)
  float* part;  // SL_COUNT x ncolTiles x prow
  size_t bytes;
  int mp, dp, nW, ncolTiles, prow;
};

// tied variant passes the same W pointer in every layer: map layer -> index of its unique weight
struct WeightMap {
  std::vector<const float*> uniq;
  std::vector<int> idx;
  explicit WeightMap(const dladmm_problem* p) : idx(p->K, 0) {
    for (int k = 0; k < p->K; ++k) {
      size_t j = 0;
      while (j < uniq.size() && uniq[j] != p->layers[k].W) ++j;
      if (j == uniq.size()) uniq.push_back(p->layers[k].W);
      idx[k] = (int)j;
    }
  }
};

static inline int unique_weights(const dladmm_problem* p) { return (int)WeightMap(p).uniq.size(); }
static inline int weight_index(const dladmm_problem* p, int k) { return WeightMap(p).idx[k]; }

static inline Workspace carve(const dladmm_problem* p, int for_backward) {
  Workspace w;
  memset(&w, 0, sizeof(w));
  w.mp = round_up(p->m, 32);
  w.dp = round_up(p->d, 32);
  w.nW = unique_weights(p);
  w.ncolTiles = (int)((p->B + SG_BN - 1) / SG_BN);
  w.prow = round_up(std::max(p->m, p->d), 32);
  char* base = (char*)p->workspace;
  size_t off = 0;
  auto take = [&](size_t nfloats) {
    float* r = (float*)(base + off);
    off += round_up64((i64)nfloats * 4, 256);
    return r;
  };
  w.Ap = take((size_t)p->m * w.dp);
  w.Wp = take((size_t)w.nW * p->d * w.mp);
  if (for_backward) {
    w.Atp = take((size_t)p->d * w.mp);
    w.Wtp = take((size_t)w.nW * p->m * w.dp);
    w.cZ = take((size_t)p->d * p->B);
    w.cE = take((size_t)p->m * p->B);
    w.cL = take((size_t)p->m * p->B);
    w.dR = take((size_t)p->m * p->B);
    w.part = take((size_t)SL_COUNT * w.ncolTiles * w.prow);
  }
  w.bytes = off;
  return w;
}

// slab addressing -----------------------------------------------------------------------------------
struct Slabs {
  const dladmm_problem* p;
  i64 zs, ms;   // slab sizes in elements
  explicit Slabs(const dladmm_problem* q) : p(q), zs((i64)q->d * q->B), ms((i64)q->m * q->B) {}
  int slot(int k) const { return p->last_only ? (k & 1) : k; }
  const float* Zin(int k) const { return k == 0 ? p->Z0 : p->Z + zs * slot(k - 1); }   // Z_{k-1}
  const float* Ein(int k) const { return k == 0 ? p->E0 : p->E + ms * slot(k - 1); }
  const float* Lin(int k) const { return k == 0 ? p->L0 : p->L + ms * slot(k - 1); }
  float* Zout(int k) const { return p->Z + zs * slot(k); }
  float* Eout(int k) const { return p->E + ms * slot(k); }
  float* Lout(int k) const { return p->L + ms * slot(k); }
  float* Tslab(int k) const { return p->T + ms * slot(k); }                            // T_k, k = 0..K
  uint8_t* mZ(int k) const { return p->maskZ ? p->maskZ + zs * slot(k) : nullptr; }
  uint8_t* mE(int k) const { return p->maskE ? p->maskE + ms * slot(k) : nullptr; }
};

static inline const dladmm_bparam& betaL(const dladmm_problem* p, const dladmm_layer& l) {
  return p->family == DLADMM_FAMILY_A ? l.beta1 : l.beta3;
}

// parameter-gradient reduction jobs -------------------------------------------------------------------
static inline void add_job(ReduceJobs& jobs, int slot, const dladmm_bparam& q, int rows, i64 part_off = 0) {
  if (q.grad == nullptr || q.ptr == nullptr || q.col_period != 0) return;   // per-slot params use atomics
  ReduceJob& j = jobs.j[jobs.n++];
  j.slot = slot;
  j.scalar = q.row_stride == 0;
  j.rows = rows;
  j.grad = q.grad;
  j.part_off = part_off;
}

static inline void add_m1_jobs(const dladmm_problem* p, ReduceJobs& jobs, const dladmm_layer& l, i64 part_off = 0) {
  add_job(jobs, SL_BL, betaL(p, l), p->m, part_off);
  if (p->family == DLADMM_FAMILY_B) {
    add_job(jobs, SL_TH2, l.theta2, p->m, part_off);
    add_job(jobs, SL_SS2, l.ss2, p->m, part_off);
    add_job(jobs, SL_B2, l.beta2, p->m, part_off);
  } else if (p->family == DLADMM_FAMILY_A) {
    add_job(jobs, SL_TH2, l.theta2, p->m, part_off);
    add_job(jobs, SL_B2, l.beta2, p->m, part_off);
  } else {
    add_job(jobs, SL_SS2, l.ss2, p->m, part_off);
    add_job(jobs, SL_B2, l.ss2_2, p->m, part_off);
  }
}

static inline int launch_reduce(const ReduceJobs& jobs, const Workspace& w, cudaStream_t st) {
  if (jobs.n == 0) return DLADMM_OK;
  int maxrows = 1;
  for (int i = 0; i < jobs.n; ++i)
    if (!jobs.j[i].scalar) maxrows = std::max(maxrows, jobs.j[i].rows);
  dim3 grid((maxrows + 7) / 8, jobs.n);
  { LaunchScope ls(DLADMM_KIND_BWD_REDUCE, st); reduce_partials_kernel<<<grid, 256, 0, st>>>(jobs, w.part, w.ncolTiles, w.prow); }
  DL_CUDA(cudaGetLastError());
  return DLADMM_OK;
}

static inline M1Args make_m1(const dladmm_problem* p, const dladmm_cotangents* g, const Workspace& w, int j) {
  Slabs s(p);
  const dladmm_layer& l = p->layers[j];
  M1Args a;
  a.Tn = s.Tslab(j + 1);
  a.Ek = s.Eout(j);
  a.Ep = s.Ein(j);
  a.Lp = s.Lin(j);
  a.maskE = s.mE(j);
  a.gE = g->gE ? g->gE + s.ms * j : nullptr;
  a.gL = g->gL ? g->gL + s.ms * j : nullptr;
  a.gT = g->gT ? g->gT + s.ms * (j + 1) : nullptr;
  a.bL = make_bp(betaL(p, l));
  a.b2 = make_bp(l.beta2);
  a.ss2 = make_bp(l.ss2);
  a.ss2_2 = make_bp(l.ss2_2);
  a.th2 = make_bp(l.theta2);
  a.dR = w.dR; a.cE = w.cE; a.cL = w.cL;
  a.B = p->B;
  a.lw = 0.f; a.lscale = nullptr; a.lkind = g->loss_kind;
  if (g->loss_kind != 0 && g->loss_scale && g->loss_layer_weight) { a.lw = g->loss_layer_weight[j]; a.lscale = g->loss_scale; }
  return a;
}

// ---- tcgen05 path (umma_fwd.cu / umma_bwd.cu) ----------------------------------------------------------------
bool umma_eligible(const dladmm_problem* p);
size_t umma_workspace_bytes(const dladmm_problem* p, int for_backward);
int umma_forward(const dladmm_problem* p, void* ws_base, cudaStream_t st);
int pf_trace_read(long long* host_out, long long capacity);
int umma_backward(const dladmm_problem* p, const dladmm_cotangents* g, const Workspace& sw, void* ws_base, cudaStream_t st);

}  // namespace dladmm
