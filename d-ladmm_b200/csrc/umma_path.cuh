// Host orchestration of the tcgen05 path (DLADMM_PREC_TF32X3 / DLADMM_PREC_TF32): weight preparation,
// tensor maps, and the per-layer launch schedule.  Included by dladmm_api.cu after its helper definitions.
#pragma once
#include "common.cuh"
#include "umma_gemm.cuh"
#include "umma_epilogues.cuh"

#define DLADMM_HAS_UMMA 1

namespace dladmm {

struct UWorkspace {
  float *Ab, *As;      // A  (m256 x dp): features of the A Z product on the N side, K = d
  float *Wb, *Ws;      // nW x (d256 x mp): features of the W V product on the N side, K = m
  float *Vb, *Vs;      // (m x B) split operand V_k = L_{k-1} + beta1_k T_k
  float *Zs;           // (d x B) small part of Z_k (or of Z0)
  size_t bytes;
  int m256, d256, mp, dp, nW;
};

static UWorkspace ucarve(const dladmm_problem* p, char* base) {
  UWorkspace w;
  memset(&w, 0, sizeof(w));
  w.m256 = round_up(p->m, umma::TILE_N);
  w.d256 = round_up(p->d, umma::TILE_N);
  w.mp = round_up(p->m, 32);
  w.dp = round_up(p->d, 32);
  w.nW = unique_weights(p);
  size_t off = 0;
  auto take = [&](size_t nfloats) {
    float* r = (float*)(base + off);
    off += round_up64((i64)nfloats * 4, 1024);
    return r;
  };
  w.Ab = take((size_t)w.m256 * w.dp);
  w.As = take((size_t)w.m256 * w.dp);
  w.Wb = take((size_t)w.nW * w.d256 * w.mp);
  w.Ws = take((size_t)w.nW * w.d256 * w.mp);
  w.Vb = take((size_t)p->m * p->B);
  w.Vs = take((size_t)p->m * p->B);
  w.Zs = take((size_t)p->d * p->B);
  w.bytes = off;
  return w;
}

static inline bool umma_eligible(const dladmm_problem* p) {
  // TMA needs a 16-byte multiple pitch for the (rows x B) activations
  return p->precision != DLADMM_PREC_FP32 && (p->B % 4) == 0;
}

static inline size_t umma_workspace_bytes(const dladmm_problem* p, int for_backward) {
  (void)for_backward;
  if (p->precision == DLADMM_PREC_FP32) return 0;
  return ucarve(p, nullptr).bytes + 1024;
}

// dense (R x C) -> zero padded (Rpad x Cpad) big/small (tf32 round-to-nearest split) or a plain padded copy
struct SplitJob { const float* src; float* big; float* small; };
struct SplitJobs { int n; SplitJob j[32]; };

template <int NPASS>
static __global__ void __launch_bounds__(256) prep_split_kernel(SplitJobs jobs, int R, int C, int Rpad, int Cpad) {
  const SplitJob jb = jobs.j[blockIdx.z];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int r0 = blockIdx.y * 32 + (threadIdx.x >> 5);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + 8 * i;
    if (r >= Rpad || c >= Cpad) continue;
    const float v = (r < R && c < C) ? jb.src[(i64)r * C + c] : 0.f;
    if (NPASS == 3) {
      const float b = umma::tf32_rna(v);
      jb.big[(i64)r * Cpad + c] = b;
      jb.small[(i64)r * Cpad + c] = v - b;
    } else {
      jb.big[(i64)r * Cpad + c] = v;
    }
  }
}

static __global__ void __launch_bounds__(256) split_trunc_kernel(const float* __restrict__ src, float* __restrict__ small, i64 n) {
  i64 i = ((i64)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 v = *reinterpret_cast<const float4*>(src + i);
    float4 s = make_float4(v.x - umma::tf32_trunc(v.x), v.y - umma::tf32_trunc(v.y), v.z - umma::tf32_trunc(v.z),
                           v.w - umma::tf32_trunc(v.w));
    *reinterpret_cast<float4*>(small + i) = s;
  } else {
    for (; i < n; ++i) small[i] = src[i] - umma::tf32_trunc(src[i]);
  }
}

template <int NPASS>
static int uprepare_weights(const dladmm_problem* p, const UWorkspace& w, cudaStream_t st) {
  {
    SplitJobs jobs; jobs.n = 1;
    jobs.j[0].src = p->A; jobs.j[0].big = w.Ab; jobs.j[0].small = w.As;
    dim3 grid((w.dp + 31) / 32, (w.m256 + 31) / 32, 1);
    { LaunchScope ls(DLADMM_KIND_PREP, st); prep_split_kernel<NPASS><<<grid, 256, 0, st>>>(jobs, p->m, p->d, w.m256, w.dp); }
    DL_CUDA(cudaGetLastError());
  }
  std::vector<const float*> uniq = WeightMap(p).uniq;
  for (size_t base = 0; base < uniq.size(); base += 32) {
    SplitJobs jobs; jobs.n = (int)std::min<size_t>(32, uniq.size() - base);
    for (int i = 0; i < jobs.n; ++i) {
      size_t idx = base + i;
      jobs.j[i].src = uniq[idx];
      jobs.j[i].big = w.Wb + idx * (size_t)w.d256 * w.mp;
      jobs.j[i].small = w.Ws + idx * (size_t)w.d256 * w.mp;
    }
    dim3 grid((w.mp + 31) / 32, (w.d256 + 31) / 32, jobs.n);
    { LaunchScope ls(DLADMM_KIND_PREP, st); prep_split_kernel<NPASS><<<grid, 256, 0, st>>>(jobs, p->d, p->m, w.d256, w.mp); }
    DL_CUDA(cudaGetLastError());
  }
  return DLADMM_OK;
}

static int device_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// C[j,b] = sum_k Wt[j,k] Act[k,b] with a fused epilogue.  act_* are (Kdim x B) batch-contiguous; w_* are prepared
// (n_pad x k_pad) K-major arrays.
template <class Epi, int NPASS>
static int launch_umma(int kind, const float* act_big, const float* act_small, int Kdim, const float* w_big, const float* w_small,
                       int n_pad, int k_pad, int n_feat, i64 B, const Epi& epi, cudaStream_t st) {
  constexpr int KC = NPASS == 3 ? 16 : 32;
  using Plan = umma::SmemPlan<NPASS, KC>;
  CUtensorMap tAb, tAs, tBb, tBs;
  int rc;
  if ((rc = umma::make_tmap_2d(&tAb, act_big, Kdim, B, B, 32, KC, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  const CUtensorMapSwizzle wsw = KC == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  if ((rc = umma::make_tmap_2d(&tBb, w_big, n_pad, k_pad, k_pad, KC, umma::TILE_N, wsw))) return rc;
  if (NPASS == 3) {
    if ((rc = umma::make_tmap_2d(&tAs, act_small, Kdim, B, B, 32, KC, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
    if ((rc = umma::make_tmap_2d(&tBs, w_small, n_pad, k_pad, k_pad, KC, umma::TILE_N, wsw))) return rc;
  } else {
    tAs = tAb; tBs = tBb;
  }
  umma::GemmShape gs;
  gs.n_feat = n_feat;
  gs.n_ntiles = (n_feat + umma::TILE_N - 1) / umma::TILE_N;
  gs.k_chunks = (Kdim + KC - 1) / KC;
  gs.B = B;
  gs.n_btiles = (B + umma::TILE_B - 1) / umma::TILE_B;
  auto kern = umma::umma_gemm_kernel<Epi, NPASS, KC>;
  static bool attr_set = false;     // per template instantiation
  if (!attr_set) {
    DL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::TOTAL));
    attr_set = true;
  }
  const i64 ntiles = gs.n_btiles * gs.n_ntiles;
  const int grid = (int)std::min<i64>(ntiles, device_sm_count());
  {
    LaunchScope ls(kind, st);
    kern<<<grid, umma::NUM_THREADS, Plan::TOTAL, st>>>(tAb, tAs, tBb, tBs, gs, epi);
  }
  DL_CUDA(cudaGetLastError());
  return DLADMM_OK;
}

template <int FAM, int NPASS, bool PS>
static int forward_umma(const dladmm_problem* p, const UWorkspace& w, cudaStream_t st) {
  Slabs s(p);
  const int m = p->m, d = p->d;
  const i64 B = p->B;
  int rc;
  if ((rc = uprepare_weights<NPASS>(p, w, st))) return rc;
  // T_0 = A Z0 + E0 - X (+ V_0)
  {
    if (NPASS == 3) {
      i64 n = (i64)d * B;
      { LaunchScope ls(DLADMM_KIND_PREP, st); split_trunc_kernel<<<(unsigned)((n / 4 + 256) / 256), 256, 0, st>>>(p->Z0, w.Zs, n); }
      DL_CUDA(cudaGetLastError());
    }
    umma::UEpiT0<NPASS, PS> epi{p->E0, p->X, p->L0, s.Tslab(0), make_bp(p->layers[0].beta1), w.Vb, w.Vs, B};
    if ((rc = launch_umma<umma::UEpiT0<NPASS, PS>, NPASS>(DLADMM_KIND_GEMM_T0, p->Z0, w.Zs, d, w.Ab, w.As, w.m256, w.dp, m, B, epi, st)))
      return rc;
  }
  for (int k = 0; k < p->K; ++k) {
    const dladmm_layer& l = p->layers[k];
    const size_t wi = (size_t)weight_index(p, k);
    {
      umma::UEpiZ<NPASS, PS> epi{s.Zin(k), s.Zout(k), w.Zs, s.mZ(k), make_bp(l.theta1), make_bp(l.ss1), B};
      if ((rc = launch_umma<umma::UEpiZ<NPASS, PS>, NPASS>(DLADMM_KIND_GEMM_Z, w.Vb, w.Vs, m, w.Wb + wi * w.d256 * w.mp,
                                                       w.Ws + wi * w.d256 * w.mp, w.d256, w.mp, d, B, epi, st)))
        return rc;
    }
    {
      umma::UEpiELT<FAM, NPASS, PS> epi;
      epi.X = p->X; epi.Ep = s.Ein(k); epi.Lp = s.Lin(k);
      epi.Ek = s.Eout(k); epi.Lk = s.Lout(k); epi.Tn = s.Tslab(k + 1); epi.maskE = s.mE(k);
      epi.b2 = make_bp(l.beta2); epi.ss2 = make_bp(l.ss2); epi.ss2_2 = make_bp(l.ss2_2); epi.th2 = make_bp(l.theta2);
      epi.bL = make_bp(betaL(p, l));
      epi.has_next = k + 1 < p->K;
      epi.b1n = make_bp(p->layers[k + 1 < p->K ? k + 1 : k].beta1);
      epi.Vb = w.Vb; epi.Vs = w.Vs; epi.B = B;
      if ((rc = launch_umma<umma::UEpiELT<FAM, NPASS, PS>, NPASS>(DLADMM_KIND_GEMM_ELT, s.Zout(k), w.Zs, d, w.Ab, w.As, w.m256, w.dp, m, B,
                                                              epi, st)))
        return rc;
    }
  }
  return DLADMM_OK;
}

static int umma_forward(const dladmm_problem* p, void* ws_base, cudaStream_t st) {
  char* base = (char*)(((uintptr_t)ws_base + 1023) & ~(uintptr_t)1023);
  UWorkspace w = ucarve(p, base);
  const bool x3 = p->precision == DLADMM_PREC_TF32X3;
  // are all broadcast parameters of the call (1,1) scalars?
  bool ps = true;
  for (int k = 0; k < p->K; ++k) {
    const dladmm_layer& l = p->layers[k];
    const dladmm_bparam* all[8] = {&l.beta1, &l.beta2, &l.beta3, &l.ss1, &l.ss2, &l.ss2_2, &l.theta1, &l.theta2};
    for (int i = 0; i < 8; ++i) ps = ps && (all[i]->ptr == nullptr || (all[i]->row_stride == 0 && all[i]->col_period == 0));
  }
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
