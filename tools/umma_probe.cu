// Standalone probe of the tcgen05 GEMM skeleton (umma_gemm.cuh): C[j,b] = sum_k Wt[j,k] * Act[k,b]
// checked against a double-precision CPU product.  Development tool, not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_probe tools/umma_probe.cu
//   ./umma_probe [n_feat Kdim B]
#define UMMA_TRACE 1
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <vector>
#include "../d-ladmm_b200/csrc/umma_gemm.cuh"

namespace dladmm {
void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n");
}
LaunchScope::LaunchScope(int, cudaStream_t) {}
LaunchScope::~LaunchScope() {}
}
using namespace dladmm;
using namespace dladmm::umma;

struct EpiStore {
#ifndef PROBE_EPI_WARPS
#define PROBE_EPI_WARPS 8
#endif
  static constexpr int WARPS = PROBE_EPI_WARPS;
  static constexpr int CHUNK = CH;
  static constexpr int NIN = 0;
  static constexpr int NROWP = 0;
  struct State {};
  struct Pre {};
  float* C; i64 B; uint32_t in_mask;
  __device__ void begin(State&) const {}
  __device__ void end(State&, int, int) const {}
  __device__ void prefetch(Pre&, int, i64, bool, int) const {}
  template <bool FULL>
  __device__ void apply(State&, const float*, int, const Pre&, int row0, i64 b, bool valid, const float (&v)[CH], int n_feat, i64) const {
    if (!valid) return;
#pragma unroll
    for (int i = 0; i < CH; ++i)
      if (row0 + i < n_feat) C[(i64)(row0 + i) * B + b] = v[i];
  }
};

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

static uint16_t bf16_rn(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x7FFFu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); }
static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }
static float rna_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }

template <int NPASS, int KC>
static double run(int n_feat, int Kdim, i64 B, const std::vector<float>& W, const std::vector<float>& Act,
                  const std::vector<double>& ref, int reps) {
  const int npad = (n_feat + TILE_N - 1) / TILE_N * TILE_N, kpad = (Kdim + 31) / 32 * 32;
  std::vector<float> Wb((size_t)npad * kpad, 0.f), Ws((size_t)npad * kpad, 0.f), As((size_t)Kdim * B);
  for (int j = 0; j < n_feat; ++j)
    for (int k = 0; k < Kdim; ++k) {
      float w = W[(size_t)j * Kdim + k];
      float big = NPASS >= 3 ? rna_tf32(w) : w;
      Wb[(size_t)j * kpad + k] = big;
      if (NPASS == 4) {          // packed bf16 rows [bf16(big)[16] | bf16(small)[16]] per 16-element k-chunk
        uint16_t* pk = reinterpret_cast<uint16_t*>(Ws.data()) + ((size_t)j * kpad + (k & ~15)) * 2 + (k & 15);
        pk[0] = bf16_rn(big); pk[16] = bf16_rn(w - big);
      } else {
        Ws[(size_t)j * kpad + k] = w - big;
      }
    }
  for (size_t i = 0; i < As.size(); ++i) As[i] = Act[i] - trunc_tf32(Act[i]);
  float *dWb, *dWs, *dA, *dAs, *dC;
  CK(cudaMalloc(&dWb, Wb.size() * 4)); CK(cudaMalloc(&dWs, Ws.size() * 4));
  CK(cudaMalloc(&dA, Act.size() * 4)); CK(cudaMalloc(&dAs, As.size() * 4)); CK(cudaMalloc(&dC, (size_t)n_feat * B * 4));
  CK(cudaMemcpy(dWb, Wb.data(), Wb.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dWs, Ws.data(), Ws.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dA, Act.data(), Act.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAs, As.data(), As.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xFF, (size_t)n_feat * B * 4));
  CUtensorMap tAb, tBb, tBs;
  int rc = 0;
  rc |= make_tmap_2d(&tAb, dA, Kdim, B, B, 32, KC, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  CUtensorMapSwizzle wsw = KC == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  rc |= make_tmap_2d(&tBb, dWb, npad, kpad, kpad, KC, TILE_N, wsw);
  rc |= make_tmap_2d(&tBs, dWs, npad, kpad, kpad, KC, TILE_N, wsw);
  if (rc) { printf("tensor map creation failed\n"); exit(1); }
  GemmShape gs;
  long long* dtr; CK(cudaMalloc(&dtr, 256 * 8 * 8)); CK(cudaMemset(dtr, 0, 256 * 8 * 8));
  gs.trace = dtr;
  gs.n_feat = n_feat; gs.n_ntiles = npad / TILE_N; gs.k_chunks = (Kdim + KC - 1) / KC; gs.B = B;
  gs.acc_scale = acc_comp_scale(gs.k_chunks * mma_per_chunk(NPASS));
  gs.n_btiles = (B + TILE_B - 1) / TILE_B;
  EpiStore epi{dC, B, 0u};
  static EMaps em; for (int i = 0; i < MAX_EIN; ++i) em.m[i] = tBb; em.mk = tBb;
  auto kern = umma_gemm_kernel<EpiStore, NPASS, KC>;
  const int smem = SmemPlan<NPASS, KC>::TOTAL;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int nsm = 0; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  i64 ntiles = gs.n_btiles * gs.n_ntiles;
  int grid = (int)(ntiles < nsm ? ntiles : nsm);
  plan_tiles(gs, grid);
  printf("tiles: %lld full + %lld halves on %d CTAs\n", (long long)gs.n_full, (long long)(gs.n_tiles - gs.n_full), grid);
  kern<<<grid, roles_threads(EpiStore::WARPS), smem>>>(tAb, tBb, tBs, em, gs, epi);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> C((size_t)n_feat * B);
  CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  double num = 0, den = 0, maxabs = 0; size_t nbad = 0;
  for (size_t i = 0; i < C.size(); ++i) {
    double d = (double)C[i] - ref[i];
    if (!(fabs(d) < 1e30)) { ++nbad; continue; }
    num += d * d; den += ref[i] * ref[i]; if (fabs(d) > maxabs) maxabs = fabs(d);
  }
  double rel = sqrt(num / den);
  if (getenv("PROBE_TRACE")) {
    std::vector<long long> tr(256 * 8);
    CK(cudaMemcpy(tr.data(), dtr, tr.size() * 8, cudaMemcpyDeviceToHost));
    long long t0 = tr[0];
    printf("chunk: prod_got tma_issued | split_saw_full split_done | mma_saw_ready mma_issued   (cycles since first event)\n");
    for (int i = 0; i < 40; ++i)
      printf("%3d: %7lld %7lld | %7lld %7lld | %7lld %7lld\n", i, tr[i * 8 + 0] - t0, tr[i * 8 + 1] - t0, tr[i * 8 + 2] - t0,
             tr[i * 8 + 3] - t0, tr[i * 8 + 4] - t0, tr[i * 8 + 5] - t0);
  }
  float ms = 0;
  if (reps > 0) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) kern<<<grid, roles_threads(EpiStore::WARPS), smem>>>(tAb, tBb, tBs, em, gs, epi);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    // SM clock under this load: CTA 0's stamps of the LAST launch (first event of chunk 0 .. MMAs of its last traced chunk)
    std::vector<long long> tr(256 * 8);
    CK(cudaMemcpy(tr.data(), dtr, tr.size() * 8, cudaMemcpyDeviceToHost));
    const i64 my_tiles = (gs.n_tiles + grid - 1) / grid;
    const i64 nchunks = std::min<i64>(256, my_tiles * gs.k_chunks);
    const long long cyc = tr[(nchunks - 1) * 8 + 5] - tr[0];
    printf("  CTA 0: %lld chunks in %lld cycles = %.0f cycles per chunk; if that is the whole launch: %.0f MHz\n", (long long)nchunks, cyc,
           (double)cyc / nchunks, nchunks == my_tiles * gs.k_chunks ? cyc / (ms * 1e3) : 0.0);
  }
  printf("%s NPASS=%d KC=%d smem=%d grid=%d: rel_l2=%.3e maxabs=%.3e nonfinite=%zu  time=%.3f ms  %.1f TFLOP/s\n", "single", NPASS, KC,
         smem, grid, rel, maxabs, nbad, ms, ms > 0 ? 2.0 * n_feat * Kdim * B / ms / 1e9 : 0.0);
  cudaFree(dWb); cudaFree(dWs); cudaFree(dA); cudaFree(dAs); cudaFree(dC);
  return rel;
}

int main(int argc, char** argv) {
  int n_feat = argc > 1 ? atoi(argv[1]) : 500, Kdim = argc > 2 ? atoi(argv[2]) : 250;
  i64 B = argc > 3 ? atoll(argv[3]) : 1000;
  int reps = argc > 4 ? atoi(argv[4]) : 0;
  printf("probe n_feat=%d Kdim=%d B=%lld\n", n_feat, Kdim, B);
  std::vector<float> W((size_t)n_feat * Kdim), Act((size_t)Kdim * B);
  srand(1);
  for (auto& x : W) x = (float)rand() / RAND_MAX - 0.5f;
  for (auto& x : Act) x = (float)rand() / RAND_MAX - 0.5f;
  std::vector<double> ref((size_t)n_feat * B, 0.0);
  if (B * n_feat <= 8000000) {
    for (int j = 0; j < n_feat; ++j)
      for (int k = 0; k < Kdim; ++k) {
        double w = W[(size_t)j * Kdim + k];
        const float* a = &Act[(size_t)k * B];
        double* r = &ref[(size_t)j * B];
        for (i64 b = 0; b < B; ++b) r[b] += w * a[b];
      }
  } else {
    printf("(large case: reference skipped, timing only)\n");
  }
  run<1, 32>(n_feat, Kdim, B, W, Act, ref, reps);
  run<3, 16>(n_feat, Kdim, B, W, Act, ref, reps);
  run<4, 16>(n_feat, Kdim, B, W, Act, ref, reps);
  return 0;
}
