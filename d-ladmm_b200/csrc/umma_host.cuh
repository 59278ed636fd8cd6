// Host orchestration shared by the tcgen05 translation units: weight preparation, tensor maps, the product launcher.
#pragma once
#include "host_common.cuh"
#include "umma_gemm.cuh"
#include "umma_epilogues.cuh"
#include "umma_persist.cuh"
#include "umma_pair.cuh"

namespace dladmm {

static int device_sm_count();

// second stage of the all-scalar parameter-gradient reduction, for every layer in one launch: grad += sum(base[0..count))
struct ScalarJob { const float* base; float* grad; };
struct ScalarJobs { int n; int count; ScalarJob j[120]; };
static __global__ void __launch_bounds__(256) reduce_scalar_entries_kernel(ScalarJobs jobs) {
  const ScalarJob jb = jobs.j[blockIdx.x];
  float s = 0.f;
  for (int i = threadIdx.x; i < jobs.count; i += 256) s += jb.base[i];
  s = warp_sum(s);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += sm[i];
    atomicAdd(jb.grad, v);          // one add per (parameter, layer); atomic only because a caller may share a scalar between layers
  }
}

struct UWorkspace {
  float *Ab, *As;      // A  (m256 x dp) zero padded, tf32 big/small: features of the A Z product on the N side, K = d
  float *Wb, *Ws;      // nW x (d256 x mp): features of the W V product on the N side, K = m
  float *V;            // (m x B) operand V_k = L_{k-1} + beta1_k T_k of the next W V product
  float *objp;         // fused objective: [K][2][OBJ_ENTRIES] per-warp partial sums (Z part, E - T part)
  float *metp;         // fused metrics: [K][DLADMM_MET_COUNT][OBJ_ENTRIES] per-warp partial sums
  float *Atb, *Ats;    // A^T (d256 x mp), only for the dual-gap metric (DLADMM_MET_DGAP_ATL)
  __nv_bfloat16 *Vh, *Zh;   // bf16 mode: the operands of the two products, (m x ldh) and 2 x (d x ldh) (Z_k alternates)
  __nv_bfloat16 *Lh;        // bf16 mode + dual-gap metric: L_k as the operand of A^T L_k
  i64 ldh;             // pitch of the bf16 operand copies (B rounded up to 8: TMA needs a 16-byte multiple)
  unsigned* pf_flags;  // persistent forward: [2K+1][batch tiles] readiness counters
  float* pf_obj;       // persistent forward: [units][8] partial sums of the fused objective
  size_t pf_flag_bytes;
  size_t bytes;
  int m256, d256, mp, dp, nW;
};
// The all-layer persistent forward lives in its own translation unit (umma_pfwd.cu); returns DLADMM_PF_NOT_TAKEN when the call is
// not eligible for it (the caller then runs the per-layer schedule of umma_fwd.cu).
constexpr int DLADMM_PF_NOT_TAKEN = 1000;
int umma_forward_persistent(const dladmm_problem* p, const UWorkspace& w, cudaStream_t st);

static inline bool wants_metric(const dladmm_problem* p, int i) { return p->metrics && ((p->metrics->want >> i) & 1u); }
constexpr int OBJ_ENTRIES = 256 * umma::MAX_EPI_WARPS;   // upper bound of (grid x epilogue warps)

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
  w.V = take((size_t)p->m * p->B);
  w.objp = take(p->objective ? (size_t)p->K * 2 * OBJ_ENTRIES : 0);
  w.metp = take((p->metrics && p->metrics->want) ? (size_t)p->K * DLADMM_MET_COUNT * OBJ_ENTRIES : 0);
  const bool atl = wants_metric(p, DLADMM_MET_DGAP_ATL);
  w.Atb = take(atl ? (size_t)w.d256 * w.mp : 0);
  w.Ats = take(atl ? (size_t)w.d256 * w.mp : 0);
  w.ldh = round_up64(p->B, 8);
  const bool bf = p->precision == DLADMM_PREC_BF16;
  w.Vh = (__nv_bfloat16*)take(bf ? (size_t)(p->m * w.ldh + 1) / 2 : 0);
  w.Zh = (__nv_bfloat16*)take(bf ? (size_t)(2 * p->d * w.ldh + 1) / 2 : 0);
  w.Lh = (__nv_bfloat16*)take(bf && atl ? (size_t)(p->m * w.ldh + 1) / 2 : 0);
  {
    const i64 nbt = (p->B + umma::TILE_B - 1) / umma::TILE_B;
    // (sized for the small-batch variant's 32-row tiles when the batch is small enough for it: see pf_tile_rows, umma_fwd.cu)
    const int tn = nbt * ((p->m + umma::TILE_N - 1) / umma::TILE_N) * 4 <= 1024 ? 32 : umma::TILE_N;
    const i64 nt_z = (p->d + tn - 1) / tn, nt_e = (p->m + tn - 1) / tn;
    const i64 units = nbt * (nt_e + (i64)p->K * (nt_z + nt_e));
    w.pf_flag_bytes = (size_t)(2 * p->K + 1) * nbt * sizeof(unsigned);
    w.pf_flags = (unsigned*)take((size_t)(2 * p->K + 1) * nbt);
    w.pf_obj = take(p->objective ? (size_t)units * 8 : 0);
  }
  w.bytes = off;
  return w;
}

struct UBwdWorkspace {
  float *Atb, *Ats;    // A^T (d256 x mp) zero padded, tf32 big/small: features of A^T dR on the N side, K = m
  float *Wtb, *Wts;    // nW x (m256 x dp): W^T, features of W^T dx1 on the N side, K = d
  float *V;            // (m x B) V_k recomputed per layer
  float *part;         // parameter-gradient partial sums of the tcgen05 epilogues
  size_t bytes;
  int m256, d256, mp, dp, nW, ngroups, prow, nentries;
  int red_group;       // layers whose per-(column group, row) partial blocks are kept before one reduction launch sums them
};

static UBwdWorkspace ucarve_bwd(const dladmm_problem* p, char* base) {
  UBwdWorkspace w;
  memset(&w, 0, sizeof(w));
  w.m256 = round_up(p->m, umma::TILE_N);
  w.d256 = round_up(p->d, umma::TILE_N);
  w.mp = round_up(p->m, 32);
  w.dp = round_up(p->d, 32);
  w.nW = unique_weights(p);
  w.ngroups = (int)(((p->B + umma::TILE_B - 1) / umma::TILE_B) * (umma::TILE_B / 32));
  w.prow = round_up(std::max(p->m, p->d), 32);
  w.nentries = 256 * umma::MAX_EPI_WARPS;      // upper bound of (grid x epilogue warps)
  size_t off = 0;
  auto take = [&](size_t nfloats) {
    float* r = (float*)(base + off);
    off += round_up64((i64)nfloats * 4, 1024);
    return r;
  };
  w.Atb = take((size_t)w.d256 * w.mp);
  w.Ats = take((size_t)w.d256 * w.mp);
  w.Wtb = take((size_t)w.nW * w.m256 * w.dp);
  w.Wts = take((size_t)w.nW * w.m256 * w.dp);
  w.V = take((size_t)p->m * p->B);
  // per-(column group, row) partials of red_group layers (at most 4, at most 512 MB), or -- when every parameter is a scalar --
  // per-warp entries of ALL layers
  const size_t per_layer = (size_t)SL_COUNT * w.ngroups * w.prow;
  w.red_group = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(4, (size_t)std::max(p->K, 1)), ((size_t)128 << 20) / std::max<size_t>(per_layer, 1)));
  w.part = take(std::max(per_layer * w.red_group, (size_t)p->K * SL_COUNT * w.nentries));
  w.bytes = off;
  return w;
}
// dense (R x C) -> zero padded (Rpad x Cpad) K-major big/small parts (tf32 round-to-nearest split) or a plain padded copy
// (TMA needs a 16-byte multiple pitch; fc.weight has 4*m = 1000 B)
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
    } else if (NPASS == 4) {
      // mixed mode: tf32 part as above; the "small" array holds, per 16-element k-chunk, a 64-byte row of
      // [bf16(big)[16] | bf16(small)[16]] -- the K-major operands of the two kind::f16 correction products
      const float b = umma::tf32_rna(v);
      jb.big[(i64)r * Cpad + c] = b;
      __nv_bfloat16* pk = reinterpret_cast<__nv_bfloat16*>(jb.small) + ((i64)r * Cpad + (c & ~15)) * 2 + (c & 15);
      pk[0] = __float2bfloat16_rn(b);
      pk[16] = __float2bfloat16_rn(v - b);
    } else if (NPASS == 2) {
      reinterpret_cast<__nv_bfloat16*>(jb.big)[(i64)r * Cpad + c] = __float2bfloat16_rn(v);
    } else {
      jb.big[(i64)r * Cpad + c] = umma::tf32_rna(v);         // single pass: nearest tf32 (the tensor core would truncate)
    }
  }
}

// dst (Rpad x Cpad) = src^T, src is (C x R) dense; zero padded; tf32 split as above
template <int NPASS>
static __global__ void __launch_bounds__(256) prep_split_t_kernel(SplitJobs jobs, int R, int C, int Rpad, int Cpad) {
  __shared__ float tile[32][33];
  const SplitJob jb = jobs.j[blockIdx.z];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {          // read src[c][r] coalesced along r
    const int c = c0 + ty + 8 * i, r = r0 + tx;
    tile[ty + 8 * i][tx] = (c < C && r < R) ? jb.src[(i64)c * R + r] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {          // write dst[r][c] coalesced along c
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    if (r >= Rpad || c >= Cpad) continue;
    const float v = tile[tx][ty + 8 * i];
    if (NPASS == 3) {
      const float b = umma::tf32_rna(v);
      jb.big[(i64)r * Cpad + c] = b;
      jb.small[(i64)r * Cpad + c] = v - b;
    } else if (NPASS == 4) {
      // mixed mode: tf32 part as above; the "small" array holds, per 16-element k-chunk, a 64-byte row of
      // [bf16(big)[16] | bf16(small)[16]] -- the K-major operands of the two kind::f16 correction products
      const float b = umma::tf32_rna(v);
      jb.big[(i64)r * Cpad + c] = b;
      __nv_bfloat16* pk = reinterpret_cast<__nv_bfloat16*>(jb.small) + ((i64)r * Cpad + (c & ~15)) * 2 + (c & 15);
      pk[0] = __float2bfloat16_rn(b);
      pk[16] = __float2bfloat16_rn(v - b);
    } else if (NPASS == 2) {
      reinterpret_cast<__nv_bfloat16*>(jb.big)[(i64)r * Cpad + c] = __float2bfloat16_rn(v);
    } else {
      jb.big[(i64)r * Cpad + c] = umma::tf32_rna(v);         // single pass: nearest tf32 (the tensor core would truncate)
    }
  }
}

// (rows x B) fp32 -> bf16 with pitch ldh: operands of the bf16 products that no epilogue of ours produced (Z0, T_init's V)
static __global__ void __launch_bounds__(256) to_bf16_kernel(const float* __restrict__ src, int rows, i64 B, __nv_bfloat16* __restrict__ dst, i64 ldh) {
  const i64 idx = (i64)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (i64)rows * B) return;
  const i64 r = idx / B, c = idx % B;
  dst[r * ldh + c] = __float2bfloat16_rn(src[idx]);
}

template <int NPASS>
static int uprepare_weights(const dladmm_problem* p, const UWorkspace& w, cudaStream_t st, bool with_At = false) {
  if (with_At) {     // A^T (d256 x mp) for the dual-gap metric's A^T L_k product
    SplitJobs jobs; jobs.n = 1;
    jobs.j[0].src = p->A; jobs.j[0].big = w.Atb; jobs.j[0].small = w.Ats;
    dim3 grid((w.mp + 31) / 32, (w.d256 + 31) / 32, 1);
    { LaunchScope ls(DLADMM_KIND_PREP, st); prep_split_t_kernel<NPASS><<<grid, 256, 0, st>>>(jobs, p->d, p->m, w.d256, w.mp); }
    DL_CUDA(cudaGetLastError());
  }
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
      jobs.j[i].big = NPASS == 2 ? (float*)((__nv_bfloat16*)w.Wb + idx * (size_t)w.d256 * w.mp)      // bf16 array, element stride
                                 : w.Wb + idx * (size_t)w.d256 * w.mp;
      jobs.j[i].small = w.Ws + idx * (size_t)w.d256 * w.mp;
    }
    dim3 grid((w.mp + 31) / 32, (w.d256 + 31) / 32, jobs.n);
    { LaunchScope ls(DLADMM_KIND_PREP, st); prep_split_kernel<NPASS><<<grid, 256, 0, st>>>(jobs, p->d, p->m, w.d256, w.mp); }
    DL_CUDA(cudaGetLastError());
  }
  return DLADMM_OK;
}

// V_k = L_{k-1} + beta1_k * T_k (backward recomputation of the dW operand)
static __global__ void __launch_bounds__(256) make_v_kernel(const float* __restrict__ L, const float* __restrict__ T, BP b1, int rows,
                                                            i64 B, float* __restrict__ V) {
  const i64 quads = (B + 3) / 4;
  const i64 idx = (i64)blockIdx.x * 256 + threadIdx.x;
  if (idx >= quads * rows) return;
  const int row = (int)(idx / quads);
  const i64 col = (idx % quads) * 4;
  const i64 rem = B - col;
  const int nv = rem >= 4 ? 4 : (int)rem;
  const bool vec = (B & 3) == 0;
  const i64 off = (i64)row * B + col;
  Quad l = load4(L, off, nv, vec), t = load4(T, off, nv, vec);
  float b[4]; bp_at4(b1, row, col, b);
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = fadd(l.v[j], fmul(b[j], t.v[j]));
  store4(V, off, v, nv, vec);
}

// out[k] = alpha * sum(partZ[k][0..nz)) + sum(partE[k][0..ne)) : second stage of the fused objective, one block per layer
static __global__ void __launch_bounds__(256) objective_reduce_kernel(const float* __restrict__ part, int nz, int ne, float alpha,
                                                                      float* __restrict__ out) {
  const float* pz = part + (size_t)blockIdx.x * 2 * OBJ_ENTRIES;
  const float* pe = pz + OBJ_ENTRIES;
  float sz = 0.f, se = 0.f;
  for (int i = threadIdx.x; i < nz; i += 256) sz += pz[i];
  for (int i = threadIdx.x; i < ne; i += 256) se += pe[i];
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

// DLADMM_NO_PDL=1 launches the product kernels fully serialized (measurement / debugging)
static bool use_pdl() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DLADMM_NO_PDL"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

// the only mutable state of the library besides the measurement hooks: per-device caches of "attribute already set" and
// of the SM count (one process may drive several devices, one host thread each)
constexpr int MAX_DEVICES = 64;
static int current_device_index() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < MAX_DEVICES) ? dev : 0;
}

static int device_sm_count() {
  static int cache[MAX_DEVICES] = {0};
  const int dev = current_device_index();
  int& n = cache[dev];
  if (!n) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    const char* cap = getenv("DLADMM_GRID_CAP");     // experiments: persistent grids of at most this many CTAs
    if (cap && atoi(cap) > 0 && atoi(cap) < n) n = atoi(cap);
  }
  return n;
}

// C[j,b] = sum_k Wt[j,k] Act[k,b] with a fused epilogue.  act_* are (Kdim x B) batch-contiguous; w_* are prepared
// (n_pad x k_pad) K-major arrays.
// NPASS: 3 = 3xTF32, 1 = TF32, 2 = BF16 (act / w_big are then bf16 arrays; act has pitch `act_pitch` elements).
template <int NPASS> struct KChunk { static constexpr int value = NPASS >= 3 ? 16 : (NPASS == 2 ? 64 : 32); };

// DLADMM_PAIR=0: never use the CTA-pair kernel; DLADMM_PAIR=1: use it for every split-precision product with at least two batch
// tiles (tests); default: for reduction lengths >= PAIR_MIN_K (mainloop-bound products: the large-scale shape)
// (an epilogue functor may lower the threshold with `static constexpr int PAIR_MIN_K`: the dZ product of the backward gains 8 % from
//  the pair kernel already at K = 250 -- its 16-warp epilogue is issue-bound and the pair kernel's leaner stage loop leaves it more
//  issue slots -- while dV, HBM-bound on its staging ring, loses 30 % there)
template <class E, class = void> struct pair_min_k_of { static constexpr int value = umma::PAIR_MIN_K; };
template <class E> struct pair_min_k_of<E, decltype((void)E::PAIR_MIN_K)> { static constexpr int value = E::PAIR_MIN_K; };
static bool pair_wanted(int Kdim, i64 B, int min_k) {
  const char* e = getenv("DLADMM_PAIR");             // (read per launch: the tests switch it inside one process)
  const int mode = e ? (e[0] == '0' ? 0 : 1) : 2;
  if (mode == 0 || B <= umma::TILE_B) return false;
  return mode == 1 || Kdim >= min_k;
}

// the split-precision product on CTA pairs (umma_pair.cuh); same arguments as launch_umma below
template <class Epi>
static int launch_umma_pair(int kind, const void* act, int Kdim, const void* w_big, const void* w_pack, int n_pad, int k_pad, int n_feat,
                            i64 B, Epi epi, cudaStream_t st, int grid_override) {
  constexpr int KC = umma::PAIR_KC;
  CUtensorMap tA, tBb, tBp;
  int rc;
  if ((rc = umma::make_tmap_2d(&tA, (const float*)act, Kdim, B, B, 32, KC, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = umma::make_tmap_2d(&tBb, (const float*)w_big, n_pad, k_pad, k_pad, KC, umma::TILE_N / 2, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = umma::make_tmap_2d(&tBp, (const float*)w_pack, n_pad, k_pad, k_pad, KC, umma::TILE_N / 2, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  umma::EMaps em;
  {
    const float* ptrs[umma::MAX_EIN] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    epi.host_inputs(ptrs);
    epi.in_mask = 0;
    for (int i = 0; i < Epi::NIN; ++i) {
      if (!ptrs[i]) continue;
      epi.in_mask |= 1u << i;
      if ((rc = umma::make_tmap_2d(&em.m[i], ptrs[i], n_feat, B, B, umma::TILE_B, Epi::CHUNK, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
    }
    for (int i = 0; i < umma::MAX_EIN; ++i)
      if (!(epi.in_mask & (1u << i))) em.m[i] = tBb;
    em.mk = tBb;
    if (epi.host_mask() && epi.in_mask != 0 && (B % 16) == 0) {
      if ((rc = umma::make_tmap_2d_u8(&em.mk, epi.host_mask(), n_feat, B, B, umma::TILE_B, Epi::CHUNK))) return rc;
      epi.in_mask |= umma::EIN_MASK_BIT;
    }
    const int nin = __builtin_popcount(epi.in_mask & ~umma::EIN_MASK_BIT);
    const int slot_bytes = nin * Epi::CHUNK * umma::TILE_B * 4 + ((epi.in_mask & umma::EIN_MASK_BIT) ? Epi::CHUNK * umma::TILE_B : 0);
    if (nin > 0 && umma::PairPlan<Epi::WARPS>::RING / slot_bytes < Epi::WARPS / 4) {
      set_error("epilogue staging ring too small (pair kernel): %d bytes per slot, %d parts", slot_bytes, Epi::WARPS / 4);
      return DLADMM_ERR_INVALID;
    }
  }
  umma::GemmShape gs;
  memset(&gs, 0, sizeof(gs));
  gs.n_feat = n_feat;
  gs.n_ntiles = (n_feat + umma::TILE_N - 1) / umma::TILE_N;
  gs.k_chunks = (Kdim + KC - 1) / KC;
  gs.acc_scale = umma::acc_comp_scale(gs.k_chunks * umma::mma_per_chunk(4));
  gs.B = B;
  gs.n_btiles = (B + umma::TILE_B - 1) / umma::TILE_B;
  gs.n_full = gs.n_tiles = gs.n_btiles * gs.n_ntiles;
  auto kern = umma::umma_gemm_pair_kernel<Epi>;
  static bool attr_set[MAX_DEVICES] = {false};
  const int dev = current_device_index();
  if (!attr_set[dev]) {
    DL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, umma::PairPlan<Epi::WARPS>::SMEM));
    attr_set[dev] = true;
  }
  const i64 n_ptiles = ((gs.n_btiles + 1) / 2) * gs.n_ntiles;
  int grid = grid_override > 0 ? grid_override : (int)std::min<i64>(2 * n_ptiles, device_sm_count());
  grid &= ~1;
  if (grid < 2) grid = 2;
  {
    LaunchScope ls(kind, st);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(umma::roles_threads(Epi::WARPS)); cfg.dynamicSmemBytes = umma::PairPlan<Epi::WARPS>::SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = use_pdl() ? 1 : 0;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 2;
    DL_CUDA(cudaLaunchKernelEx(&cfg, kern, tA, tBb, tBp, em, gs, epi));
  }
  DL_CUDA(cudaGetLastError());
  return DLADMM_OK;
}

template <class Epi, int NPASS>
static int launch_umma(int kind, const void* act, int Kdim, const void* w_big, const void* w_small,
                       int n_pad, int k_pad, int n_feat, i64 B, Epi epi, cudaStream_t st, int grid_override = 0, i64 act_pitch = 0) {
  constexpr int KC = KChunk<NPASS>::value;
  using Plan = umma::SmemPlan<NPASS, KC>;
  if constexpr (NPASS == 4) {
    if (pair_wanted(Kdim, B, pair_min_k_of<Epi>::value)) return launch_umma_pair<Epi>(kind, act, Kdim, w_big, w_small, n_pad, k_pad, n_feat, B, epi, st, grid_override);
  }
  CUtensorMap tA, tBb, tBs;
  int rc;
  if (NPASS == 2) {
    if ((rc = umma::make_tmap_2d_bf16(&tA, act, Kdim, B, act_pitch ? act_pitch : B, 64, KC, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = umma::make_tmap_2d_bf16(&tBb, w_big, n_pad, k_pad, k_pad, KC, umma::TILE_N, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    tBs = tBb;
  } else {
    if ((rc = umma::make_tmap_2d(&tA, (const float*)act, Kdim, B, B, 32, KC, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
    const CUtensorMapSwizzle wsw = KC == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    if ((rc = umma::make_tmap_2d(&tBb, (const float*)w_big, n_pad, k_pad, k_pad, KC, umma::TILE_N, wsw))) return rc;
    if (NPASS >= 3) {
      if ((rc = umma::make_tmap_2d(&tBs, (const float*)w_small, n_pad, k_pad, k_pad, KC, umma::TILE_N, wsw))) return rc;
    } else {
      tBs = tBb;
    }
  }
  // epilogue inputs staged by TMA: one (CHUNK rows x 128 columns) box per present array
  umma::EMaps em;                          // per call, on the stack: the launch copies it into the kernel parameters
  {
    const float* ptrs[umma::MAX_EIN] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    epi.host_inputs(ptrs);
    epi.in_mask = 0;
    for (int i = 0; i < Epi::NIN; ++i) {
      if (!ptrs[i]) continue;
      epi.in_mask |= 1u << i;
      if ((rc = umma::make_tmap_2d(&em.m[i], ptrs[i], n_feat, B, B, umma::TILE_B, Epi::CHUNK, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
    }
    for (int i = 0; i < umma::MAX_EIN; ++i)
      if (!(epi.in_mask & (1u << i))) em.m[i] = tBb;       // placeholder, never dereferenced
    em.mk = tBb;
    // the 1-byte prox masks ride the same ring when TMA can address them (16-byte multiple pitch)
    if (epi.host_mask() && epi.in_mask != 0 && (B % 16) == 0) {
      if ((rc = umma::make_tmap_2d_u8(&em.mk, epi.host_mask(), n_feat, B, B, umma::TILE_B, Epi::CHUNK))) return rc;
      epi.in_mask |= umma::EIN_MASK_BIT;
    }
  }
  {
    // every epilogue part must own at least one slot of the staging ring (a consumer whose first slot is a wrapped one
    // would see that slot's barrier as already completed)
    const int nin = __builtin_popcount(epi.in_mask & ~umma::EIN_MASK_BIT);
    const int slot_bytes = nin * Epi::CHUNK * umma::TILE_B * 4 + ((epi.in_mask & umma::EIN_MASK_BIT) ? Epi::CHUNK * umma::TILE_B : 0);
    if (nin > 0 && umma::ring_bytes_of<Epi, NPASS, KC>() / slot_bytes < Epi::WARPS / 4) {
      set_error("epilogue staging ring too small: %d bytes per slot, %d parts", slot_bytes, Epi::WARPS / 4);
      return DLADMM_ERR_INVALID;
    }
  }
  umma::GemmShape gs;
  gs.n_feat = n_feat;
  gs.n_ntiles = (n_feat + umma::TILE_N - 1) / umma::TILE_N;
  gs.k_chunks = (Kdim + KC - 1) / KC;
  gs.acc_scale = umma::acc_comp_scale(gs.k_chunks * umma::mma_per_chunk(NPASS));
  gs.B = B;
  gs.n_btiles = (B + umma::TILE_B - 1) / umma::TILE_B;
  auto kern = umma::umma_gemm_kernel<Epi, NPASS, KC>;
  static bool attr_set[MAX_DEVICES] = {false};     // per template instantiation and per device
  const int dev = current_device_index();
  if (!attr_set[dev]) {
    DL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::TOTAL));
    attr_set[dev] = true;
  }
  const i64 ntiles = gs.n_btiles * gs.n_ntiles;
  const int grid = grid_override > 0 ? grid_override : (int)std::min<i64>(ntiles, device_sm_count());
  umma::plan_tiles(gs, grid);
  {
    LaunchScope ls(kind, st);
    // programmatic stream serialization: the kernel may be scheduled while the previous product kernel drains (it calls
    // griddepcontrol.launch_dependents after its prologue); its own loads and stores wait for that grid to complete
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(umma::roles_threads(Epi::WARPS)); cfg.dynamicSmemBytes = Plan::TOTAL; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = use_pdl() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    DL_CUDA(cudaLaunchKernelEx(&cfg, kern, tA, tBb, tBs, em, gs, epi));
  }
  DL_CUDA(cudaGetLastError());
  return DLADMM_OK;
}

// umma::PM_SCALAR / PM_ROWS / PM_GENERAL of a problem (see umma_epilogues.cuh)
static int param_mode(const dladmm_problem* p);
static bool all_params_scalar(const dladmm_problem* p) {
  bool ps = true;
  for (int k = 0; k < p->K; ++k) {
    const dladmm_layer& l = p->layers[k];
    const dladmm_bparam* all[8] = {&l.beta1, &l.beta2, &l.beta3, &l.ss1, &l.ss2, &l.ss2_2, &l.theta1, &l.theta2};
    for (int i = 0; i < 8; ++i) ps = ps && (all[i]->ptr == nullptr || (all[i]->row_stride == 0 && all[i]->col_period == 0));
  }
  return ps;
}

static int param_mode(const dladmm_problem* p) {
  if (all_params_scalar(p)) return umma::PM_SCALAR;
  for (int k = 0; k < p->K; ++k) {
    const dladmm_layer& l = p->layers[k];
    const dladmm_bparam* all[8] = {&l.beta1, &l.beta2, &l.beta3, &l.ss1, &l.ss2, &l.ss2_2, &l.theta1, &l.theta2};
    for (int i = 0; i < 8; ++i)
      if (all[i]->ptr && (all[i]->col_period != 0 || (all[i]->row_stride != 0 && all[i]->row_stride != 1))) return umma::PM_GENERAL;
  }
  return umma::PM_ROWS;
}

}  // namespace dladmm
