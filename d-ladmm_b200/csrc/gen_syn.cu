// On-device synthetic data generator: the device equivalent of gen_syn_data.py:12-47
//   A ~ N(0,1)^{m x d}, columns normalised to unit L2 norm            (gen_syn_data.py:14-16)
//   Z = Bernoulli(p) * N(mu, sigma)   (d x B)                         (:26-33)
//   E = Bernoulli(p) * N(mu, sigma)   (m x B)   [or dense N(0, sigma_e), gen_syn_unseen_data_lasso.py:41-42]
//   X = A Z + E                                                       (:46-47)
// and its out-of-distribution variants: amplitudes cos(g) / 1/(1+exp(g)) (gen_syn_unseen_data_cosine.py,
// gen_syn_unseen_data_logistic.py); replaced columns of A are a host-side edit of A (gen_syn.py::replace_A_columns).
// Random numbers come from Philox4x32-10 keyed by (seed) and countered by (stream, row, global column pair),
// so the data does not depend on how columns are sharded over GPUs.
#include "common.cuh"
#include "epilogues.cuh"
#include "simt_gemm.cuh"

namespace dladmm {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint64_t seed) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// uniform in (0,1]
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  float r = sqrtf(-2.0f * logf(u01(a)));
  float s, c;
  sincospif(2.0f * u01(b), &s, &c);
  n0 = r * c; n1 = r * s;
}

enum { STREAM_A = 0, STREAM_Z = 1, STREAM_E = 2 };

__global__ void __launch_bounds__(256) gen_A_kernel(float* __restrict__ A, i64 n, uint64_t seed) {
  i64 q = (i64)blockIdx.x * 256 + threadIdx.x;      // group of 4 elements
  if (q * 4 >= n) return;
  uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32), 0u, (uint32_t)STREAM_A};
  philox4x32_10(c, seed);
  float v[4];
  box_muller(c[0], c[1], v[0], v[1]);
  box_muller(c[2], c[3], v[2], v[3]);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (q * 4 + j < n) A[q * 4 + j] = v[j];
}

// one warp per column: A[:, j] /= ||A[:, j]||_2
__global__ void __launch_bounds__(256) normalize_cols_kernel(float* __restrict__ A, int m, int d) {
  int j = blockIdx.x * 8 + (threadIdx.x >> 5), l = threadIdx.x & 31;
  if (j >= d) return;
  float s = 0.f;
  for (int i = l; i < m; i += 32) { float v = A[(i64)i * d + j]; s += v * v; }
  s = warp_sum(s);
  float inv = 1.0f / sqrtf(s);
  for (int i = l; i < m; i += 32) A[(i64)i * d + j] *= inv;
}

// rows x B sparse (or dense) Gaussian field; each thread writes 4 consecutive columns of one row
__global__ void __launch_bounds__(256) gen_field_kernel(float* __restrict__ out, int rows, i64 B, i64 col_offset,
                                                        uint64_t seed, int stream, float p, float mu, float sigma,
                                                        int dense, int amplitude) {
  const i64 quads = (B + 3) / 4;
  i64 idx = (i64)blockIdx.x * 256 + threadIdx.x;
  if (idx >= quads * rows) return;
  int row = (int)(idx / quads);
  i64 col = (idx % quads) * 4;
  float v[4];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    // global column pair index: columns (2*pair, 2*pair+1)
    i64 gcol = col_offset + col + 2 * h;
    // pairs are aligned on the GLOBAL column grid so that shards agree; handle odd offsets elementwise
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      i64 gc = gcol + e;
      i64 pair = gc >> 1;
      uint32_t c[4] = {(uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)row, (uint32_t)stream};
      philox4x32_10(c, seed);
      float n0, n1;
      box_muller(c[0], c[1], n0, n1);
      float nrm = (gc & 1) ? n1 : n0;
      float u = u01((gc & 1) ? c[3] : c[2]);
      float val = mu + sigma * nrm;
      if (amplitude == 1) val = cosf(val);                       // gen_syn_unseen_data_cosine.py:33-34
      else if (amplitude == 2) val = 1.0f / (1.0f + expf(val));   // gen_syn_unseen_data_logistic.py:33-34
      v[2 * h + e] = dense ? val : (u <= p ? val : 0.f);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (col + j < B) out[(i64)row * B + col + j] = v[j];
}

struct EpiAddE {
  static constexpr int NRED = 0;
  static constexpr int SLOT0 = 0;
  const float* Es; float* X; i64 B;
  __device__ __forceinline__ void operator()(int row, i64 col, const float (&acc)[4], int nvalid, bool vec, float*) const {
    i64 off = (i64)row * B + col;
    Quad e = load4(Es, off, nvalid, vec);
    float x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = fadd(acc[j], e.v[j]);
    store4(X, off, x, nvalid, vec);
  }
};

}  // namespace dladmm

using namespace dladmm;

extern "C" {

size_t dladmm_gen_workspace_bytes(int32_t m, int32_t d) { return (size_t)m * round_up(d, 32) * sizeof(float) + 256; }

int dladmm_gen_syn(const dladmm_gen_desc* g, void* stream) {
  DL_REQUIRE(g != nullptr, "gen desc is NULL");
  DL_REQUIRE(g->m > 0 && g->d > 0 && g->B >= 0, "bad sizes m=%d d=%d", g->m, g->d);
  DL_REQUIRE(g->A && (g->B == 0 || (g->Zs && g->Es && g->X)), "A, Zs, Es, X must be non-NULL");
  DL_REQUIRE(g->p >= 0.f && g->p <= 1.f, "p must be in [0,1]");
  DL_REQUIRE(g->amplitude >= 0 && g->amplitude <= 2, "unknown amplitude %d", g->amplitude);
  cudaStream_t st = (cudaStream_t)stream;
  const int m = g->m, d = g->d;
  if (g->generate_A) {
    i64 n = (i64)m * d;
    { LaunchScope ls(DLADMM_KIND_GEN, st); gen_A_kernel<<<(unsigned)((n / 4 + 256) / 256), 256, 0, st>>>(g->A, n, g->seed); }
    DL_CUDA(cudaGetLastError());
    { LaunchScope ls(DLADMM_KIND_GEN, st); normalize_cols_kernel<<<(d + 7) / 8, 256, 0, st>>>(g->A, m, d); }
    DL_CUDA(cudaGetLastError());
  }
  if (g->B == 0) return DLADMM_OK;
  const int dp = round_up(d, 32);
  if (!g->workspace || g->workspace_bytes < dladmm_gen_workspace_bytes(m, d)) {
    set_error("gen workspace too small: need %zu bytes", dladmm_gen_workspace_bytes(m, d));
    return DLADMM_ERR_WORKSPACE;
  }
  const i64 quads = (g->B + 3) / 4;
  { LaunchScope ls(DLADMM_KIND_GEN, st);
    gen_field_kernel<<<(unsigned)((quads * d + 255) / 256), 256, 0, st>>>(g->Zs, d, g->B, g->col_offset, g->seed, STREAM_Z,
                                                                          g->p, g->mu, g->sigma, 0, g->amplitude); }
  DL_CUDA(cudaGetLastError());
  LaunchScope lsE(DLADMM_KIND_GEN, st);
  if (g->dense_noise)
    gen_field_kernel<<<(unsigned)((quads * m + 255) / 256), 256, 0, st>>>(g->Es, m, g->B, g->col_offset, g->seed, STREAM_E,
                                                                          1.f, 0.f, g->sigma_e, 1, 0);
  else
    gen_field_kernel<<<(unsigned)((quads * m + 255) / 256), 256, 0, st>>>(g->Es, m, g->B, g->col_offset, g->seed, STREAM_E,
                                                                          g->p, g->mu, g->sigma, 0, g->amplitude);
  DL_CUDA(cudaGetLastError());
  // X = A Zs + Es through the same fused product kernel
  float* Ap = (float*)g->workspace;
  PrepJobs jobs; jobs.n = 1;
  jobs.j[0].src = g->A; jobs.j[0].dst_n = Ap; jobs.j[0].dst_t = nullptr;
  dim3 pg((dp + 31) / 32, (m + 31) / 32, 1);
  { LaunchScope ls(DLADMM_KIND_PREP, st); prep_weights_kernel<<<pg, 256, 0, st>>>(jobs, m, d, dp, 0); }
  DL_CUDA(cudaGetLastError());
  BPlain bl{g->Zs, g->B};
  EpiAddE epi{g->Es, g->X, g->B};
  dim3 grid((unsigned)((g->B + SG_BN - 1) / SG_BN), (unsigned)((m + SG_BM - 1) / SG_BM));
  { LaunchScope ls(DLADMM_KIND_GEN, st);
    simt_gemm_kernel<BPlain, EpiAddE><<<grid, SG_THREADS, 0, st>>>(m, g->B, d, Ap, dp, bl, epi, nullptr, 0, 0); }
  DL_CUDA(cudaGetLastError());
  return DLADMM_OK;
}

}  // extern "C"
