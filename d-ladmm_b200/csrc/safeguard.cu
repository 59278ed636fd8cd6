// Safeguarded-evaluation helpers (test_syn_l1l1_scalar.py:163-176, 238-274): the column norm of the safeguard
// operator S and the per-column selection between the learned and the classical iterate.  The two matrix products
// of every KM / learned step run through dladmm_forward with K = 1 and T_init.
#include "common.cuh"

namespace dladmm {

// Column norms of the safeguard test.  A block owns 32 columns (one warp-wide, coalesced row segment per load) and NORM_RG row
// groups: thread (cx, ry) sums rows ry, ry + NORM_RG, ... of column cx with four rows of loads in flight, the NORM_RG partial
// sums of a column are added in a fixed order through shared memory (deterministic).  One thread per column over all rows -- the
// first version -- is a chain of m dependent L2 round trips on 16 blocks: 132 us per call at 4 096 columns, 38 % of a safeguarded
// evaluation's kernel time (profiles/r02_launches_summary.md).
constexpr int NORM_RG = 16;
constexpr int NORM_THREADS = 32 * NORM_RG;

// out[b] = || [ beta * Tn[:,b] ; c * (En[:,b] - 2 Ek[:,b] + Ep[:,b]) ] ||_2      (test_syn_l1l1_scalar.py:173, two_norm)
static __global__ void __launch_bounds__(NORM_THREADS) sg_norm_kernel(int m, i64 B, float beta, float c, const float* __restrict__ Tn,
                                                                      const float* __restrict__ En, const float* __restrict__ Ek,
                                                                      const float* __restrict__ Ep, float* __restrict__ out) {
  __shared__ float part[NORM_RG][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const i64 b = (i64)blockIdx.x * 32 + cx;
  float s = 0.f;
  if (b < B) {
#pragma unroll 4
    for (int i = ry; i < m; i += NORM_RG) {
      const i64 off = (i64)i * B + b;
      const float t = beta * Tn[off];
      const float e = c * ((En[off] - 2.f * Ek[off]) + Ep[off]);
      s += t * t + e * e;
    }
  }
  part[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && b < B) {
    float a = 0.f;
#pragma unroll
    for (int r = 0; r < NORM_RG; ++r) a += part[r][cx];
    out[b] = sqrtf(a);
  }
}

// Snorm_ELZ (test_syn_l1l1_newS_Acols.py:174-192) with (Znn-Zn)^T P2 (Znn-Zn) = ||Znn-Zn||^2/(beta ss1) - ||A(Znn-Zn)||^2 and
// A(Znn-Zn) = Tnn - Tn: coalesced over the batch, no d x d operator.
static __global__ void __launch_bounds__(NORM_THREADS) sg_norm_elz_kernel(int m, int d, i64 B, float inv_bss1, const float* __restrict__ Tnn,
                                                                          const float* __restrict__ Tn, const float* __restrict__ Znn,
                                                                          const float* __restrict__ Zn, const float* __restrict__ Esub,
                                                                          const float* __restrict__ Eadd, float* __restrict__ out) {
  __shared__ float part[3][NORM_RG][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const i64 b = (i64)blockIdx.x * 32 + cx;
  float s1 = 0.f, s3 = 0.f, s2 = 0.f;
  if (b < B) {
#pragma unroll 4
    for (int i = ry; i < m; i += NORM_RG) {
      const i64 off = (i64)i * B + b;
      float t = Tnn[off];
      if (Esub) t = (t - Esub[off]) + Eadd[off];                     // Tnn given as A Znn + E' - X: swap E' for En
      const float dt = t - Tn[off];
      s1 += t * t;
      s3 += dt * dt;
    }
#pragma unroll 4
    for (int i = ry; i < d; i += NORM_RG) {
      const i64 off = (i64)i * B + b;
      const float dz = Znn[off] - Zn[off];
      s2 += dz * dz;
    }
  }
  part[0][ry][cx] = s1; part[1][ry][cx] = s2; part[2][ry][cx] = s3;
  __syncthreads();
  if (ry == 0 && b < B) {
    float a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int r = 0; r < NORM_RG; ++r) { a1 += part[0][r][cx]; a2 += part[1][r][cx]; a3 += part[2][r][cx]; }
    out[b] = sqrtf(fmaxf(a1 + (inv_bss1 * a2 - a3), 0.f));
  }
}

// keep + mu_k update + fallback count in one pass over the columns (mu_updater.py:18-72)
static __global__ void __launch_bounds__(256) sg_keep_update_kernel(i64 B, const float* __restrict__ snorm, float* __restrict__ mu,
                                                                    float one_minus_delta, int method, float param,
                                                                    float* __restrict__ keep, float* __restrict__ fallbacks) {
  const i64 b = (i64)blockIdx.x * 256 + threadIdx.x;
  float fb = 0.f;
  if (b < B) {
    const float s = snorm[b], m0 = mu[b];
    const bool k = s < one_minus_delta * m0;
    keep[b] = k ? 1.f : 0.f;
    fb = k ? 0.f : 1.f;
    if (k && method != 0) {
      float upd = s;                                               // RT
      if (method == 1) upd = param * s + (1.f - param) * m0;        // EMA
      else if (method == 2) upd = (1.f - param) * m0;               // GS
      if (method != 4) mu[b] = upd;
    }
    if (method == 4) mu[b] = 1e10f;                                 // "None" (BlankUpdater): no threshold after the first test
  }
  fb = warp_sum(fb);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = fb;
  __syncthreads();
  if (threadIdx.x == 0 && fallbacks) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += sm[i];
    if (v != 0.f) atomicAdd(fallbacks, v);                          // integer-valued: exact in any order
  }
}

// keep[b] = (snorm[b] < one_minus_delta * mu[b]) ; out = keep ? a : b, column-wise, for several (rows x B) arrays
struct SelJob { const float* a; const float* b; float* out; int rows; };
struct SelJobs { int n; SelJob j[6]; };

static __global__ void __launch_bounds__(256) sg_keep_kernel(i64 B, const float* __restrict__ snorm, const float* __restrict__ mu,
                                                             float one_minus_delta, float* __restrict__ keep) {
  const i64 b = (i64)blockIdx.x * 256 + threadIdx.x;
  if (b < B) keep[b] = snorm[b] < one_minus_delta * mu[b] ? 1.f : 0.f;
}

static __global__ void __launch_bounds__(256) sg_select_kernel(SelJobs jobs, i64 B, const float* __restrict__ keep) {
  const SelJob jb = jobs.j[blockIdx.y];
  const i64 n = (i64)jb.rows * B;
  for (i64 idx = (i64)blockIdx.x * 256 + threadIdx.x; idx < n; idx += (i64)gridDim.x * 256) {
    const i64 b = idx % B;
    jb.out[idx] = keep[b] != 0.f ? jb.a[idx] : jb.b[idx];
  }
}

}  // namespace dladmm

using namespace dladmm;

extern "C" {

int dladmm_sg_norm(int32_t m, int64_t B, float beta, float c, const float* Tn, const float* En, const float* Ek, const float* Ep,
                   float* out, void* stream) {
  DL_REQUIRE(m > 0 && B >= 0 && Tn && En && Ek && Ep && out, "sg_norm: bad arguments");
  if (B == 0) return DLADMM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  { LaunchScope ls(DLADMM_KIND_SAFEGUARD, st);
    sg_norm_kernel<<<(unsigned)((B + 31) / 32), NORM_THREADS, 0, st>>>(m, B, beta, c, Tn, En, Ek, Ep, out); }
  DL_CUDA(cudaGetLastError());
  return DLADMM_OK;
}

int dladmm_sg_norm_elz(int32_t m, int32_t d, int64_t B, float inv_beta_ss1, const float* Tnn, const float* Tn, const float* Znn,
                       const float* Zn, const float* E_sub, const float* E_add, float* out, void* stream) {
  DL_REQUIRE(m > 0 && d > 0 && B >= 0 && Tnn && Tn && Znn && Zn && out, "sg_norm_elz: bad arguments");
  DL_REQUIRE((E_sub == nullptr) == (E_add == nullptr), "sg_norm_elz: E_sub and E_add come together");
  if (B == 0) return DLADMM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  { LaunchScope ls(DLADMM_KIND_SAFEGUARD, st);
    sg_norm_elz_kernel<<<(unsigned)((B + 31) / 32), NORM_THREADS, 0, st>>>(m, d, B, inv_beta_ss1, Tnn, Tn, Znn, Zn, E_sub, E_add, out); }
  DL_CUDA(cudaGetLastError());
  return DLADMM_OK;
}

static int run_select(int32_t n_arrays, const dladmm_sg_pair* pairs, int64_t B, const float* keep, cudaStream_t st);

int dladmm_sg_select_update(int32_t n_arrays, const dladmm_sg_pair* pairs, int64_t B, const float* snorm, float* mu,
                            float one_minus_delta, int32_t method, float param, float* keep, float* fallbacks, void* stream) {
  DL_REQUIRE(n_arrays >= 0 && n_arrays <= 6 && (n_arrays == 0 || pairs) && snorm && mu && keep, "sg_select_update: bad arguments");
  DL_REQUIRE(method >= 0 && method <= 4, "sg_select_update: unknown mu updater %d", method);
  if (B == 0) return DLADMM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  { LaunchScope ls(DLADMM_KIND_SAFEGUARD, st);
    sg_keep_update_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(B, snorm, mu, one_minus_delta, method, param, keep, fallbacks); }
  DL_CUDA(cudaGetLastError());
  return run_select(n_arrays, pairs, B, keep, st);
}

int dladmm_sg_select(int32_t n_arrays, const dladmm_sg_pair* pairs, int64_t B, const float* snorm, const float* mu,
                     float one_minus_delta, float* keep, void* stream) {
  DL_REQUIRE(n_arrays >= 0 && n_arrays <= 6 && (n_arrays == 0 || pairs) && snorm && mu && keep, "sg_select: bad arguments");
  if (B == 0) return DLADMM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  { LaunchScope ls(DLADMM_KIND_SAFEGUARD, st);
    sg_keep_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(B, snorm, mu, one_minus_delta, keep); }
  DL_CUDA(cudaGetLastError());
  return run_select(n_arrays, pairs, B, keep, st);
}

}  // extern "C"

static int run_select(int32_t n_arrays, const dladmm_sg_pair* pairs, int64_t B, const float* keep, cudaStream_t st) {
  if (n_arrays == 0) return DLADMM_OK;
  SelJobs jobs; jobs.n = n_arrays;
  int maxrows = 1;
  for (int i = 0; i < n_arrays; ++i) {
    DL_REQUIRE(pairs[i].a && pairs[i].b && pairs[i].out && pairs[i].rows > 0, "sg_select: pair %d incomplete", i);
    jobs.j[i].a = pairs[i].a; jobs.j[i].b = pairs[i].b; jobs.j[i].out = pairs[i].out; jobs.j[i].rows = pairs[i].rows;
    maxrows = pairs[i].rows > maxrows ? pairs[i].rows : maxrows;
  }
  i64 blocks = ((i64)maxrows * B + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  { LaunchScope ls(DLADMM_KIND_SAFEGUARD, st);
    sg_select_kernel<<<dim3((unsigned)blocks, n_arrays), 256, 0, st>>>(jobs, B, keep); }
  DL_CUDA(cudaGetLastError());
  return DLADMM_OK;
}
