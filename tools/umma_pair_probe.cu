// Standalone probe: the split-precision product (tf32_bf16x2) with tcgen05.mma.cta_group::2 -- a CTA PAIR (two SMs of one TPC) computes
// M = 256 batch columns (128 per CTA) x N = 256 feature rows per instruction; each CTA stages its own activation tile and HALF of the
// weight tile, so per CTA the shared-memory traffic of a k-chunk drops from 104 KB to 72 KB and the L2 -> SM weight bytes halve.
// Store-only epilogue, checked against a double-precision CPU product.  Development tool, not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probe_pair tools/umma_pair_probe.cu -lcuda
//   timeout 30 ./probe_pair [n_feat Kdim B reps]
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
using namespace dladmm;
using namespace dladmm::umma;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int PKC = 16;                         // k-chunk
constexpr int P_A = TILE_B * PKC * 4;           // 8 KB raw activation tile (+ 8 KB of bf16 hi/lo)
constexpr int P_BH = (TILE_N / 2) * PKC * 4;    // 8 KB: this CTA's 128 rows of the tf32 weights (+ 8 KB packed bf16)
constexpr int P_STAGE = 2 * P_A + 2 * P_BH;     // 32 KB
constexpr int P_STAGES = 6;
constexpr int P_SMEM = P_STAGES * P_STAGE + 2048 + 1024;
constexpr int P_EPI_WARPS = 8;
constexpr int P_THREADS = 64 + 32 * P_EPI_WARPS + 32 * 4;     // producer, MMA, 8 epilogue, 4 splitters

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {      // acquire at cluster scope (remote arrivals)
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
// TMA load whose completion is signalled on a barrier given as a shared::cluster address (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {       // arrives on `bar` of BOTH CTAs when the MMAs issued so far have retired
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

struct PairShape { int n_feat, n_ntiles, k_chunks; i64 B, n_btiles, n_ptiles; float acc_scale; long long* trace; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB_big, const __grid_constant__ CUtensorMap tmB_pack,
            PairShape gs, float* __restrict__ C) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + P_STAGES * P_STAGE);
  uint64_t* afull = bars;                       // [S] local: this CTA's activation tile landed
  uint64_t* bfull = bars + P_STAGES;            // [S] leader: both weight halves landed
  uint64_t* ready = bars + 2 * P_STAGES;        // [S] leader: both CTAs' splitters are done with the stage
  uint64_t* empty = bars + 3 * P_STAGES;        // [S] local: the MMAs that read the stage retired (multicast commit)
  uint64_t* tfull = bars + 4 * P_STAGES;        // [2] local
  uint64_t* tempty = bars + 4 * P_STAGES + 2;   // [2] leader: both CTAs' epilogues drained the accumulator
  uint32_t* tmem_slot = (uint32_t*)(bars + 4 * P_STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(&afull[s], 1); mbar_init(&bfull[s], 1); mbar_init(&ready[s], 8); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 2 * P_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int npair = gridDim.x >> 1, pair = blockIdx.x >> 1;

  if (warp == 0) {
    // ===== TMA producer: own activation tile -> local afull; own half of the weights -> the leader's bfull =====
    int s = 0; uint32_t ph = 0; int trp = 0;
    for (i64 pt = pair; pt < gs.n_ptiles; pt += npair) {
      const i64 bt = 2 * (pt / gs.n_ntiles) + rank;
      const int j0 = (int)(pt % gs.n_ntiles) * TILE_N;
      const int b0 = (int)(bt * TILE_B);
      for (int kc = 0; kc < gs.k_chunks; ++kc) {
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
          if (blockIdx.x == 0 && gs.trace && trp < 256) gs.trace[trp * 8 + 0] = clock64();
          uint8_t* st = smem + s * P_STAGE;
          mbar_expect_tx(&afull[s], P_A);
#pragma unroll
          for (int g = 0; g < 4; ++g) tma_load_2d(st + g * (PKC * 128), &tmA, &afull[s], b0 + g * 32, kc * PKC);
          if (leader) mbar_expect_tx(&bfull[s], 4 * P_BH);            // both halves, tf32 + packed bf16
          const uint32_t bf = mapa(smem_u32(&bfull[s]), 0);
          tma_load_2d_pair(st + 2 * P_A, &tmB_big, bf, kc * PKC, j0 + (int)rank * (TILE_N / 2));
          tma_load_2d_pair(st + 2 * P_A + P_BH, &tmB_pack, bf, kc * PKC, j0 + (int)rank * (TILE_N / 2));
          if (blockIdx.x == 0 && gs.trace && trp < 256) gs.trace[trp * 8 + 1] = clock64();
        }
        ++trp;
        __syncwarp();
        if (++s == P_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===== MMA issuer (leader CTA only) =====
      constexpr uint32_t idesc32 = make_idesc(256, TILE_N, 1, 0, 2u), idesc16 = make_idesc(256, TILE_N, 1, 0, 1u);
      constexpr uint32_t a_hi = desc_hi(A_ATOM_BYTES, LAYOUT_SW128_BASE32B), a16_hi = desc_hi(1024, LAYOUT_SW128);
      constexpr uint32_t b_hi = desc_hi(8 * PKC * 4, LAYOUT_SW64);
      const uint32_t st0 = smem_u32(smem);
      const uint32_t a_lo0 = desc_lo(st0, PKC * 128), b_lo0 = desc_lo(st0 + 2 * P_A, 16);
      int s = 0; uint32_t ph = 0, acc = 0, aph = 0; int tr = 0;
      for (i64 pt = pair; pt < gs.n_ptiles; pt += npair) {
        mbar_wait_cluster(&tempty[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TILE_N;
        for (int kc = 0; kc < gs.k_chunks; ++kc, ++tr) {
          mbar_wait_cluster(&bfull[s], ph);
          if (blockIdx.x == 0 && lane == 0 && gs.trace && tr < 256) gs.trace[tr * 8 + 6] = clock64();
          mbar_wait_cluster(&ready[s], ph);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + s * (P_STAGE >> 4), b_lo = b_lo0 + s * (P_STAGE >> 4);
          if (elect_one()) {
            if (blockIdx.x == 0 && gs.trace && tr < 256) gs.trace[tr * 8 + 4] = clock64();
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma2_tf32(d_tmem, desc_at(a_hi, a_lo + ks * ((8 * 128) >> 4)), desc_at(b_hi, b_lo + ks * 2), idesc32, (kc == 0 && ks == 0) ? 0u : 1u);
            const uint32_t a16 = a_lo + (P_A >> 4), b16 = b_lo + (P_BH >> 4);
            umma2_f16(d_tmem, desc_at(a16_hi, a16), desc_at(b_hi, b16 + 2), idesc16, 1u);
            umma2_f16(d_tmem, desc_at(a16_hi, a16 + (P_A >> 5)), desc_at(b_hi, b16), idesc16, 1u);
            umma2_commit(&empty[s]);
            if (kc == gs.k_chunks - 1) umma2_commit(&tfull[acc]);
            if (blockIdx.x == 0 && gs.trace && tr < 256) gs.trace[tr * 8 + 5] = clock64();
          }
          __syncwarp();
          if (++s == P_STAGES) { s = 0; ph ^= 1; }
        }
        if (++acc == 2) { acc = 0; aph ^= 1; }
      }
    }
  } else if (warp >= 2 + P_EPI_WARPS) {
    // ===== splitters: bf16 hi / lo of this CTA's activation tile; every warp arrives on the LEADER's ready barrier =====
    const int tid = threadIdx.x - (2 + P_EPI_WARPS) * 32;
    int s = 0; uint32_t ph = 0; int trs = 0;
    for (i64 pt = pair; pt < gs.n_ptiles; pt += npair) {
      for (int kc = 0; kc < gs.k_chunks; ++kc) {
        mbar_wait(&afull[s], ph);
        if (blockIdx.x == 0 && tid == 0 && gs.trace && trs < 256) gs.trace[trs * 8 + 2] = clock64();
        uint8_t* st = smem + s * P_STAGE;
        split_tile_mix<128>(st, st + P_A, st + P_A + P_A / 2, tid);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa(smem_u32(&ready[s]), 0));
        if (blockIdx.x == 0 && tid == 0 && gs.trace && trs < 256) gs.trace[trs * 8 + 3] = clock64();
        ++trs;
        if (++s == P_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===== epilogue: store the accumulator (this CTA's 128 batch columns x 256 feature rows) =====
    const int q = warp & 3, half = (warp - 2) >> 2, col = q * 32 + lane;
    uint32_t acc = 0, aph = 0;
    for (i64 pt = pair; pt < gs.n_ptiles; pt += npair) {
      const i64 bt = 2 * (pt / gs.n_ntiles) + rank;
      const int j0 = (int)(pt % gs.n_ntiles) * TILE_N;
      const i64 b = bt * TILE_B + col;
      mbar_wait(&tfull[acc], aph);
      tc_fence_after();
      const uint32_t t0 = tmem_base + acc * TILE_N + half * 128 + ((uint32_t)(q * 32) << 16);
      for (int c = 0; c < 8; ++c) {
        float v[16];
        tmem_ld16(t0 + c * 16, v);
        const int row0 = j0 + half * 128 + c * 16;
        if (b < gs.B)
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (row0 + i < gs.n_feat) C[(i64)(row0 + i) * gs.B + b] = v[i] * gs.acc_scale;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa(smem_u32(&tempty[acc]), 0));
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static uint16_t bf16_rn(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x7FFFu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); }
static float rna_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }

int main(int argc, char** argv) {
  int n_feat = argc > 1 ? atoi(argv[1]) : 500, Kdim = argc > 2 ? atoi(argv[2]) : 250;
  i64 B = argc > 3 ? atoll(argv[3]) : 1000;
  int reps = argc > 4 ? atoi(argv[4]) : 0;
  printf("pair probe n_feat=%d Kdim=%d B=%lld\n", n_feat, Kdim, (long long)B);
  std::vector<float> W((size_t)n_feat * Kdim), Act((size_t)Kdim * B);
  srand(1);
  for (auto& x : W) x = (float)rand() / RAND_MAX - 0.5f;
  for (auto& x : Act) x = (float)rand() / RAND_MAX - 0.5f;
  const bool check = B * n_feat <= 8000000;
  std::vector<double> ref(check ? (size_t)n_feat * B : 0, 0.0);
  if (check)
    for (int j = 0; j < n_feat; ++j)
      for (int k = 0; k < Kdim; ++k) {
        const double w = W[(size_t)j * Kdim + k];
        for (i64 b = 0; b < B; ++b) ref[(size_t)j * B + b] += w * Act[(size_t)k * B + b];
      }
  const int npad = (n_feat + TILE_N - 1) / TILE_N * TILE_N, kpad = (Kdim + 31) / 32 * 32;
  std::vector<float> Wb((size_t)npad * kpad, 0.f), Wp((size_t)npad * kpad, 0.f);
  for (int j = 0; j < n_feat; ++j)
    for (int k = 0; k < Kdim; ++k) {
      const float w = W[(size_t)j * Kdim + k], big = rna_tf32(w);
      Wb[(size_t)j * kpad + k] = big;
      uint16_t* pk = reinterpret_cast<uint16_t*>(Wp.data()) + ((size_t)j * kpad + (k & ~15)) * 2 + (k & 15);
      pk[0] = bf16_rn(big); pk[16] = bf16_rn(w - big);
    }
  float *dWb, *dWp, *dA, *dC;
  CK(cudaMalloc(&dWb, Wb.size() * 4)); CK(cudaMalloc(&dWp, Wp.size() * 4)); CK(cudaMalloc(&dA, Act.size() * 4)); CK(cudaMalloc(&dC, (size_t)n_feat * B * 4));
  CK(cudaMemcpy(dWb, Wb.data(), Wb.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dWp, Wp.data(), Wp.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dA, Act.data(), Act.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(dC, 0xFF, (size_t)n_feat * B * 4));
  CUtensorMap tA, tBb, tBp;
  int rc = 0;
  rc |= make_tmap_2d(&tA, dA, Kdim, B, B, 32, PKC, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  rc |= make_tmap_2d(&tBb, dWb, npad, kpad, kpad, PKC, TILE_N / 2, CU_TENSOR_MAP_SWIZZLE_64B);
  rc |= make_tmap_2d(&tBp, dWp, npad, kpad, kpad, PKC, TILE_N / 2, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) { printf("tensor map creation failed\n"); return 1; }
  PairShape gs;
  gs.n_feat = n_feat; gs.n_ntiles = npad / TILE_N; gs.k_chunks = (Kdim + PKC - 1) / PKC; gs.B = B;
  gs.n_btiles = (B + TILE_B - 1) / TILE_B; gs.n_ptiles = ((gs.n_btiles + 1) / 2) * gs.n_ntiles;
  gs.acc_scale = acc_comp_scale(gs.k_chunks * 4);
  long long* dtr; CK(cudaMalloc(&dtr, 256 * 8 * 8)); CK(cudaMemset(dtr, 0, 256 * 8 * 8));
  gs.trace = dtr;
  CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
  int nsm = 0; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  int grid = (int)std::min<i64>(2 * gs.n_ptiles, nsm) & ~1;
  printf("pair tiles: %lld on %d CTAs (%d pairs), %d k-chunks\n", (long long)gs.n_ptiles, grid, grid / 2, gs.k_chunks);
  pair_kernel<<<grid, P_THREADS, P_SMEM>>>(tA, tBb, tBp, gs, dC);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  double rel = -1, maxabs = 0;
  if (check) {
    std::vector<float> Cc((size_t)n_feat * B);
    CK(cudaMemcpy(Cc.data(), dC, Cc.size() * 4, cudaMemcpyDeviceToHost));
    double num = 0, den = 0; size_t nbad = 0;
    for (size_t i = 0; i < Cc.size(); ++i) {
      const double d = (double)Cc[i] - ref[i];
      if (!(fabs(d) < 1e30)) { ++nbad; continue; }
      num += d * d; den += ref[i] * ref[i]; if (fabs(d) > maxabs) maxabs = fabs(d);
    }
    rel = sqrt(num / den);
    printf("check: rel_l2=%.3e maxabs=%.3e nonfinite=%zu\n", rel, maxabs, nbad);
  }
  if (reps > 0) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) pair_kernel<<<grid, P_THREADS, P_SMEM>>>(tA, tBb, tBp, gs, dC);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    std::vector<long long> tr(256 * 8);
    CK(cudaMemcpy(tr.data(), dtr, tr.size() * 8, cudaMemcpyDeviceToHost));
    const i64 my = (gs.n_ptiles + grid / 2 - 1) / (grid / 2);
    const i64 nch = std::min<i64>(256, my * gs.k_chunks);
    const long long cyc = tr[(nch - 1) * 8 + 5] - tr[4];
    if (getenv("PROBE_TRACE")) {
      const long long t0 = tr[0];
      printf("chunk: prod_got tma_issued | split_saw_afull split_arrived | mma_saw_bfull mma_saw_ready mma_issued\n");
      for (int i = 0; i < 36; ++i)
        printf("%3d: %7lld %7lld | %7lld %7lld | %7lld %7lld %7lld\n", i, tr[i * 8 + 0] - t0, tr[i * 8 + 1] - t0, tr[i * 8 + 2] - t0, tr[i * 8 + 3] - t0,
               tr[i * 8 + 6] - t0, tr[i * 8 + 4] - t0, tr[i * 8 + 5] - t0);
    }
    printf("pair NPASS=4: time=%.3f ms  %.1f TFLOP/s;  leader of pair 0: %lld chunks in %lld cycles = %.0f per chunk (%.0f MHz if that is the launch)\n", ms,
           2.0 * n_feat * Kdim * B / ms / 1e9, (long long)nch, cyc, (double)cyc / nch, nch == my * gs.k_chunks ? cyc / (ms * 1e3) : 0.0);
  }
  return 0;
}
