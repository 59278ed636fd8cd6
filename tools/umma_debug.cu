// Minimal single-tile tcgen05 debug: one CTA, Kdim=32 (one stage), prints smem words and TMEM results.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <vector>
#include "../d-ladmm_b200/csrc/umma_gemm.cuh"
namespace dladmm {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); }
LaunchScope::LaunchScope(int, cudaStream_t) {}
LaunchScope::~LaunchScope() {}
}
using namespace dladmm; using namespace dladmm::umma;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(128, 1) dbg_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                     float* out, int mode, int nk) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                 // 128 batch x 32 k  (16 KB)
  uint8_t* sb = smem + 16384;         // 256 rows x 32 k   (32 KB)
  uint64_t* bars = (uint64_t*)(smem + 49152);
  uint32_t* slot = (uint32_t*)(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = *slot;
  if (threadIdx.x == 0) {
    printf("tmem_base=0x%08x smem_a=0x%x\n", tbase, smem_u32(sa));
    mbar_expect_tx(&bars[0], 16384 + 32768);
    if (mode == 0 || mode == 2) { for (int g = 0; g < 4; ++g) tma_load_2d(sa + g * 4096, &tmA, &bars[0], g * 32, 0); }
    else tma_load_2d(sa, &tmA, &bars[0], 0, 0);
    tma_load_2d(sb, &tmB, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0);
    const float* fa = (const float*)sa; const float* fb = (const float*)sb;
    printf("sa[0..7]: %g %g %g %g %g %g %g %g | row1: %g %g\n", fa[0], fa[1], fa[2], fa[3], fa[4], fa[5], fa[6], fa[7], fa[32], fa[33]);
    printf("sb[0..7]: %g %g %g %g %g %g %g %g | row1: %g %g\n", fb[0], fb[1], fb[2], fb[3], fb[4], fb[5], fb[6], fb[7], fb[32], fb[33]);
    tc_fence_after();
    uint32_t idesc = make_idesc(128, 256, mode != 1 ? 1 : 0, 0);
    for (int ks = 0; ks < nk; ++ks) {
      uint64_t da = mode == 2 ? make_sdesc(smem_u32(sa) + ks * 1024, 4096, 512, 1)
                  : mode == 0 ? make_sdesc(smem_u32(sa) + ks * 1024, 4096, 1024, LAYOUT_SW128)
                              : make_sdesc(smem_u32(sa) + ks * 32, 16, 1024, LAYOUT_SW128);
      uint64_t db = make_sdesc(smem_u32(sb) + ks * 32, 16, 1024, LAYOUT_SW128);
      if (ks == 0) printf("idesc=0x%08x da=0x%016llx db=0x%016llx\n", idesc, (unsigned long long)da, (unsigned long long)db);
      umma_tf32(tbase, da, db, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(&bars[1]);
  }
  __syncthreads();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16), v);
  for (int i = 0; i < 32; ++i) out[(size_t)(warp * 32 + lane) * 32 + i] = v[i];
  if (warp == 0) {   // scan every column of lanes 0..31 for anything nonzero
    int nz = 0; float first = 0.f; int firstc = -1;
    for (int c = 0; c < 512; c += 32) {
      float w[32];
      tmem_ld32(tbase + c, w);
      for (int i = 0; i < 32; ++i) if (w[i] != 0.f) { if (firstc < 0) { firstc = c + i; first = w[i]; } ++nz; }
    }
    if (lane < 2) printf("lane %d: nonzero TMEM words=%d first col=%d val=%g\n", lane, nz, firstc, first);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main(int argc, char** argv) {
  int nk = argc > 1 ? atoi(argv[1]) : 1;
  int mode = argc > 2 ? atoi(argv[2]) : 0;
  const int K = 32, B = 128, N = 256;
  std::vector<float> Act((size_t)K * B), W((size_t)N * K);
  for (int k = 0; k < K; ++k) for (int b = 0; b < B; ++b) Act[(size_t)k * B + b] = (float)((k + 1) * 1000 + b);
  for (int j = 0; j < N; ++j) for (int k = 0; k < K; ++k) W[(size_t)j * K + k] = (float)(j * 100 + k);
  // simple check data: overwrite with small exactly representable values
  std::vector<float> A2(Act.size()), W2(W.size());
  for (int k = 0; k < K; ++k) for (int b = 0; b < B; ++b) A2[(size_t)k * B + b] = (float)((b % 7) - 3) * (float)(1 + (k % 3));
  for (int j = 0; j < N; ++j) for (int k = 0; k < K; ++k) W2[(size_t)j * K + k] = (float)((j % 5) - 2) * (float)(1 + (k % 2));
  float *dA, *dW, *dO;
  CK(cudaMalloc(&dA, A2.size() * 4)); CK(cudaMalloc(&dW, W2.size() * 4)); CK(cudaMalloc(&dO, 128 * 32 * 4));
  CK(cudaMemcpy(dA, A2.data(), A2.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, W2.data(), W2.size() * 4, cudaMemcpyHostToDevice));
  std::vector<float> A2T((size_t)B * K);
  for (int k = 0; k < K; ++k) for (int b = 0; b < B; ++b) A2T[(size_t)b * K + k] = A2[(size_t)k * B + b];
  float* dAT; CK(cudaMalloc(&dAT, A2T.size() * 4)); CK(cudaMemcpy(dAT, A2T.data(), A2T.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap tA, tB;
  if ((mode == 2 ? make_tmap_2d(&tA, dA, K, B, B, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
       : mode == 0 ? make_tmap_2d(&tA, dA, K, B, B, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B)
                 : make_tmap_2d(&tA, dAT, B, K, K, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B)) || make_tmap_2d(&tB, dW, N, K, K, 32, 256, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  CK(cudaFuncSetAttribute(dbg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  dbg_kernel<<<1, 128, 64 * 1024>>>(tA, tB, dO, mode, nk);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  std::vector<float> O(128 * 32);
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int b = 0; b < 128; ++b) for (int j = 0; j < 32; ++j) {
    double r = 0; for (int k = 0; k < nk * 8; ++k) r += (double)A2[(size_t)k * B + b] * W2[(size_t)j * K + k];
    if (fabs(r - O[b * 32 + j]) > 1e-3) { if (bad < 10) printf("mismatch b=%d j=%d got %g want %g\n", b, j, O[b * 32 + j], r); ++bad; }
  }
  printf("nk=%d mismatches=%d of %d; O[0][0..3]=%g %g %g %g O[1][0]=%g O[33][2]=%g\n", nk, bad, 128 * 32, O[0], O[1], O[2], O[3], O[32], O[33 * 32 + 2]);
  return 0;
}
