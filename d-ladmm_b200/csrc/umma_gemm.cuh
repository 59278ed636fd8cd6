// tcgen05 / TMEM / TMA matrix product with fused epilogue (sm_100a), the tensor-core arithmetic of the
// D-LADMM layer (DLADMM_PREC_TF32X3 and DLADMM_PREC_TF32).
//
//   C[j, b] = sum_k  Wt[j, k] * Act[k, b]          j: feature row, b: batch column (problem instance)
//
// Mapping onto one tcgen05.mma (cta_group::1, kind::tf32, fp32 accumulate in TMEM):
//   M (TMEM lanes, 128)      = batch columns        -> UMMA "A" = Act tile, MN-major (batch contiguous), TMA
//   N (TMEM columns, <=256)  = feature rows         -> UMMA "B" = prepared weights, K-major, TMA
//   K (8 per instruction)    = reduction index
// so that in the epilogue the 32 lanes of a warp hold 32 consecutive batch columns of one feature row:
// every global access of the fused epilogue is a coalesced 128-byte row segment, and per-row parameters are
// warp-uniform.
//
// 3xTF32: x = big + small with big = the tf32 the tensor core sees; three MMAs per k-step accumulate
// small*big + big*small + big*big in the same TMEM accumulator (fp32-level products).  Weights are split once per call
// (round-to-nearest) by the prep kernel.  Activations arrive as plain fp32 through TMA and are split IN SHARED MEMORY by
// four splitter warps (big = the raw word, which the tensor core truncates to tf32; small = x - trunc(x) written next to
// it), so no split copy of an activation ever exists in HBM.  (Splitting the weights in shared memory as well was measured
// slower: shared-memory bandwidth -- TMA writes + split + three operand reads per k-step -- is what bounds this kernel.)
//
// BF16 mode (NPASS == 2, DLADMM_PREC_BF16): kind::f16 on bf16 operands, one pass, K = 16 per instruction.  The activation
// operand is a bf16 copy (pitch multiple of 8) that the PRODUCING epilogue wrote next to its fp32 output, loaded as boxes of
// (64 batch columns x KC rows) with the plain 128-byte swizzle (MN-major canonical layout ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)),
// cute/atom/mma_traits_sm100.hpp); weights are converted once per call.  No splitter work.
//
// Warp roles (480 threads): warp 0 = TMA producer (operands), warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = epilogue (TMEM lane quadrant = warp_id % 4; two warps per quadrant split the feature rows),
// warps 10..13 = operand splitters (3xTF32 only), warp 14 = TMA producer of the epilogue-input staging ring.  Persistent over (batch tile, feature tile) pairs,
// two accumulators in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include <cuda.h>
#include <algorithm>
#include <cstdlib>
#include "common.cuh"

namespace dladmm {
namespace umma {

constexpr int TILE_B = 128;      // batch columns per tile (UMMA M)
constexpr int TILE_N = 256;      // feature rows per tile (UMMA N)
constexpr int UMMA_K = 8;        // tf32
// Epilogue warps: `parts` per TMEM lane quadrant, each owning 1/parts of the tile's feature rows.  Each epilogue functor
// picks its own count (Epi::WARPS = 8 or 16: the heavier elementwise chains want more warps to hide LDS/LDTM latency, the
// ones with many staged inputs want fewer parts so that each part keeps several ring slots in flight).
constexpr int MAX_EPI_WARPS = 16;
constexpr int roles_threads(int epi_warps) { return 64 + 32 * epi_warps + 32 * 4 + 32; }
constexpr int SPLIT_WARPS = 4;        // shared-memory tf32 splitters (3-pass mode)
constexpr int EPI_WARP0 = 2;             // warps [2, 2 + WARPS) = epilogue, then SPLIT_WARPS splitters, then the staging-ring producer
constexpr int CH = 16;            // feature rows per epilogue step (one tcgen05.ld.32x32b.x16)

// ---- PTX wrappers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait suspends the thread in hardware until the phase completes or a time limit passes; the explicit limit (ns) keeps a
// waiting role from coming back to re-issue the instruction every few hundred cycles (the kernels run under the power cap:
// issue slots spent on polling are clocks taken from the tensor pipe)
#ifndef DLADMM_MBAR_SUSPEND_NS
#define DLADMM_MBAR_SUSPEND_NS 100000
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
#if DLADMM_MBAR_SUSPEND_NS > 0
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#endif
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)DLADMM_MBAR_SUSPEND_NS)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ float tf32_small(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
// small = x - trunc_tf32(x) for BYTES of operand tiles (any layout: the transform is elementwise).  Fully unrolled: all
// the 16-byte loads of a thread are in flight before the first store (a rolled loop costs ~90 clk per 16 bytes and made
// the splitters the throughput limit of the whole pipeline -- measured with clock64 stamps in tools/umma_probe.cu).
template <int BYTES, int NTHREADS>
__device__ __forceinline__ void split_tile(const uint8_t* raw, uint8_t* small, int tid) {
  constexpr int N = BYTES / 16 / NTHREADS;
  static_assert(N * 16 * NTHREADS == BYTES, "tile bytes must be a multiple of 16 * threads");
  const float4* src = reinterpret_cast<const float4*>(raw);
  float4* dst = reinterpret_cast<float4*>(small);
  float4 v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = src[tid + i * NTHREADS];
#pragma unroll
  for (int i = 0; i < N; ++i)
    dst[tid + i * NTHREADS] = make_float4(tf32_small(v[i].x), tf32_small(v[i].y), tf32_small(v[i].z), tf32_small(v[i].w));
}
// Mixed mode (NPASS == 4, DLADMM_PREC_TF32_BF16X2): x*w = trunc_tf32(x)*rna_tf32(w) on kind::tf32 plus the two first-order correction
// products on kind::f16 at twice the rate -- bf16(x)*bf16(w - rna_tf32(w)) and bf16(x - trunc_tf32(x))*bf16(rna_tf32(w)); the
// correction operands only need ~8 significant bits because they are already 2^-11 of the product.  The splitter warps read the
// raw fp32 activation tile (four boxes of 32 batch columns x 16 k-rows, 128-byte rows whose 32-byte chunks are XOR-permuted by
// row % 4: the TMA SWIZZLE_128B_ATOM_32B image) and write the two bf16 tiles next to it in the MN-major SWIZZLE_128B image of
// the bf16 mode (two boxes of 64 batch columns x 16 k-rows, 128-byte rows whose 16-byte chunks are XOR-permuted by row % 8).
// A thread owns 8 consecutive batch columns of one k-row: 32 bytes in, 16 + 16 bytes out; the two 16-byte loads of a thread are
// issued in opposite order by neighbouring column groups so that a quarter warp covers all 32 banks.
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <int NTHREADS>
__device__ __forceinline__ void split_tile_mix(const uint8_t* raw, uint8_t* hi, uint8_t* lo, int tid) {
  constexpr int UNITS = 16 * (TILE_B / 8);               // (k-row, group of 8 batch columns)
  constexpr int N = UNITS / NTHREADS;
  static_assert(N * NTHREADS == UNITS, "units must divide over the splitter threads");
  float4 va[N], vb[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int u = tid + i * NTHREADS;
    const int c8 = u & 15, r = u >> 4;                   // column group 0..15, k-row 0..15
    const int g = c8 >> 2, c32 = c8 & 3;
    const float4* src = reinterpret_cast<const float4*>(raw + g * 2048 + r * 128 + ((c32 ^ (r & 3)) << 5));
    const int sw = g & 1;
    va[i] = src[sw]; vb[i] = src[sw ^ 1];
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int u = tid + i * NTHREADS;
    const int c8 = u & 15, r = u >> 4;
    const bool sw = ((c8 >> 2) & 1) != 0;
    const float4 x0 = sw ? vb[i] : va[i], x1 = sw ? va[i] : vb[i];      // columns 0..3, 4..7 of the group
    uint4 h, l;
    h.x = pack_bf16x2(x0.x, x0.y); h.y = pack_bf16x2(x0.z, x0.w); h.z = pack_bf16x2(x1.x, x1.y); h.w = pack_bf16x2(x1.z, x1.w);
    l.x = pack_bf16x2(tf32_small(x0.x), tf32_small(x0.y)); l.y = pack_bf16x2(tf32_small(x0.z), tf32_small(x0.w));
    l.z = pack_bf16x2(tf32_small(x1.x), tf32_small(x1.y)); l.w = pack_bf16x2(tf32_small(x1.z), tf32_small(x1.w));
    const int off = (c8 >> 3) * 2048 + r * 128 + (((c8 & 7) ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}
// Single-pass TF32: the tensor core TRUNCATES fp32 operands to tf32 -- a bias toward zero that is coherent from layer to layer
// (measured: 7.7e-4 per product, but 2.2e-2 on the iterates after 15 layers; bf16 operands rounded to nearest: 2.4e-3 per
// product, 5.6e-3 after 15 layers).  The splitter warps therefore round the activation tile to the nearest tf32 in place.
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
template <int BYTES, int NTHREADS>
__device__ __forceinline__ void round_tile(uint8_t* raw, int tid) {
  constexpr int N = BYTES / 16 / NTHREADS;
  static_assert(N * 16 * NTHREADS == BYTES, "tile bytes must be a multiple of 16 * threads");
  float4* p = reinterpret_cast<float4*>(raw);
  float4 v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = p[tid + i * NTHREADS];
#pragma unroll
  for (int i = 0; i < N; ++i) p[tid + i * NTHREADS] = make_float4(tf32_rn(v[i].x), tf32_rn(v[i].y), tf32_rn(v[i].z), tf32_rn(v[i].w));
}
// Programmatic dependent launch: a kernel launched with programmatic stream serialization may become resident while its
// predecessor drains; everything that reads or writes global memory sits behind grid_dep_wait().
__device__ __forceinline__ void grid_dep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <bool BF>
__device__ __forceinline__ void umma_op(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (BF) umma_f16(d_tmem, adesc, bdesc, idesc, accumulate);
  else umma_tf32(d_tmem, adesc, bdesc, idesc, accumulate);
}
// arrives on the mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[4]) { tmem_ld4(taddr, v); }
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[8]) { tmem_ld8(taddr, v); }
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[16]) { tmem_ld16(taddr, v); }

// one lane of a CONVERGED warp; lets ptxas emit the uniform-datapath instructions (UTCHMMA, UTMALDG, UTCBAR) straight-line
// instead of wrapping each of them in an elect/branch loop as it must inside divergent `if (lane == 0)` code
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// ---- descriptors ---------------------------------------------------------------------------------------
// instruction descriptor, kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor bit layout)
// a/b format field: kind::tf32 -> 2 (TF32); kind::f16 -> 0 (F16), 1 (BF16)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major, uint32_t ab_format = 2u) {
  return (1u << 4) | (ab_format << 7) | (ab_format << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): addresses/offsets in 16-byte units
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}
// descriptor = (hi << 32) | lo with lo = (addr >> 4) | (LBO >> 4) << 16 : advancing the start address is a 32-bit add on lo
__device__ __forceinline__ uint64_t desc_at(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | ((layout_type & 7u) << 29);
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16); }
constexpr uint32_t LAYOUT_SW128 = 2, LAYOUT_SW64 = 4;
// MN-major 32-bit operands have exactly one legal swizzled layout: 128-byte rows whose four 32-byte chunks are
// permuted by (row % 4) -- UMMA layout type 1 (SWIZZLE_128B_BASE32B), written by TMA with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  Atom = 32 elements (MN) x 4 rows (K) = 512 bytes.  (Plain SWIZZLE_128B
// with an MN-major tf32 operand silently multiplies by zero; measured on B200.)
constexpr uint32_t LAYOUT_SW128_BASE32B = 1;
constexpr uint32_t A_ATOM_BYTES = 512;

// ---- kernel ----------------------------------------------------------------------------------------------
#ifdef UMMA_TRACE
#define UMMA_TR(gs, idx, ev) do { if (blockIdx.x == 0 && (gs).trace && (idx) < 256) (gs).trace[(idx) * 8 + (ev)] = clock64(); } while (0)
#else
#define UMMA_TR(gs, idx, ev) do { } while (0)
#endif

// The tensor core adds every tcgen05.mma into its fp32 TMEM accumulator with round-toward-zero: each accumulating instruction
// shrinks the running sum by a fraction of an ulp.  Measured on B200 (profiles/r02_precision_table.md, mean signed error of one
// product against fp64): 1.7e-8 .. 1.9e-8 of the result per MMA instruction, the same for kind::tf32 and kind::f16 and for
// reduction lengths 250 .. 2000 -- a BIAS, and the dominant error of the fp32-class modes (2.3e-6 of 2.8e-6 at K = 500), which
// being coherent from layer to layer is what grows with depth.  The epilogues therefore scale the accumulator they read by
// 1 + ACC_RZ_BIAS_PER_MMA * (accumulating MMAs of the element): a first-order compensation of the expected loss (exact for no
// element; it removes the mean).  DLADMM_ACC_COMP=0 switches it off (scale 1.0f: bit-identical to the uncompensated kernels).
constexpr float ACC_RZ_BIAS_PER_MMA = 1.8e-8f;
__host__ __device__ constexpr int mma_per_chunk(int npass) { return npass == 3 ? 6 : 4; }   // 3xTF32: 2 k-steps x 3; the others: 4 per chunk
inline float acc_comp_unit() {         // the per-MMA constant, or 0 when the compensation is switched off
  static int on = -1;
  if (on < 0) { const char* e = getenv("DLADMM_ACC_COMP"); on = (e && e[0] == '0') ? 0 : 1; }
  return on ? ACC_RZ_BIAS_PER_MMA : 0.0f;
}
inline float acc_comp_scale(int n_mma) { return 1.0f + acc_comp_unit() * (float)n_mma; }

struct GemmShape {
  float acc_scale;   // accumulator compensation (see ACC_RZ_BIAS_PER_MMA); 1.0f = none
#ifdef UMMA_TRACE
  long long* trace;  // [256][8] clock64 stamps of CTA 0: 0 producer got the stage, 1 TMA issued, 2 splitter saw full, 3 split done,
                     //                                    4 MMA saw ready, 5 MMAs issued
#endif
  int n_feat;        // valid feature rows (epilogue bound); weights are zero padded to a multiple of TILE_N
  int n_ntiles;      // ceil(n_feat / TILE_N)
  int k_chunks;      // ceil(Kdim / KC)
  i64 B;             // batch columns
  i64 n_btiles;      // ceil(B / TILE_B)
  // Tile list: tiles [0, n_full) are (128 columns x 256 rows), numbered batch-tile major; tiles [n_full, n_tiles) are the
  // remaining batch tiles cut into (128 x 128) halves.  The host cuts the last, partly filled round of the persistent
  // grid this way when that shortens the longest CTA (512 tiles on 148 SMs: 4 rounds -> 3 rounds + a round of halves).
  i64 n_full, n_tiles;
};

struct TileInfo { i64 bt; int j0; int nrows; };
__device__ __forceinline__ TileInfo decode_tile(const GemmShape& gs, i64 tile) {
  TileInfo t;
  if (tile < gs.n_full) {
    t.bt = tile / gs.n_ntiles; t.j0 = (int)(tile % gs.n_ntiles) * TILE_N; t.nrows = TILE_N;
  } else {
    const i64 h = tile - gs.n_full;
    const int per = 2 * gs.n_ntiles;
    t.bt = gs.n_full / gs.n_ntiles + h / per; t.j0 = (int)(h % per) * (TILE_N / 2); t.nrows = TILE_N / 2;
  }
  return t;
}

// Epilogue inputs (the (rows x B) arrays the fused epilogue reads elementwise) are staged through a shared-memory
// ring by TMA: a dedicated producer thread streams (CHUNK rows x 128 columns) boxes of every input array, several
// chunks -- and tiles -- ahead of the epilogue warps, so the epilogue never waits a DRAM round trip in registers.
constexpr int MAX_EIN = 10;              // staged input arrays per epilogue
#ifndef UMMA_RING_BYTES
#define UMMA_RING_BYTES (72 * 1024)
#endif
#ifndef UMMA_STAGES
#define UMMA_STAGES 3
#endif
constexpr int RING_BYTES = UMMA_RING_BYTES;    // staging ring; depth = RING_BYTES / (present arrays * CHUNK * 512 B)

struct EMaps { CUtensorMap m[MAX_EIN]; CUtensorMap mk; };   // float inputs + the 1-byte prox mask of the chunk
constexpr uint32_t EIN_MASK_BIT = 1u << 31;                 // in_mask bit: the mask bytes are staged too

template <int NPASS, int KC, int TN = TILE_N>
struct SmemPlan {
  static constexpr int EB = NPASS == 2 ? 2 : 4;                      // operand element bytes (bf16 mode: 2)
  static constexpr int NOPS = NPASS >= 3 ? 2 : 1;                    // big (+ small; mixed mode: + the two bf16 correction operands)
  static constexpr int A_BYTES = TILE_B * KC * EB;                   // one operand part
  static constexpr int B_BYTES = TN * KC * EB;                       // TN feature rows per tile (256; the persistent kernel's small-batch variant: 32)
  static constexpr int STAGE_BYTES = NOPS * (A_BYTES + B_BYTES);     // [A raw | A small] [B big | B small]
  static constexpr int TX_BYTES = A_BYTES + NOPS * B_BYTES;          // what TMA delivers (A small is computed in place)
  static constexpr int STAGES = UMMA_STAGES;
  static constexpr int BAR_BYTES = 1024;
  static constexpr int ROWTAB = 8192;                                  // per-row parameter table of the epilogue (umma_epilogues.cuh)
  static constexpr int TOTAL = STAGES * STAGE_BYTES + RING_BYTES + BAR_BYTES + ROWTAB + 1024;   // + alignment slack
};
constexpr int MAX_RING_DEPTH = 36;

// Epi: functor with
//   CHUNK (feature rows per epilogue step), NIN (staged float input arrays), in_mask (bit i: array i present),
//   State / Pre types, begin/end,
//   prefetch(pre, row0, b, valid, n_feat)          register loads of small side inputs (prox masks), one chunk ahead
//   apply(state, slot, col, pre, row0, b, valid, acc[CHUNK], n_feat, group)
// where `slot` points at the staged inputs of this chunk: present arrays back to back, each [CHUNK][128] floats.
// position in the epilogue-input staging ring: slot index and mbarrier phase, advanced without integer division
struct RingPos {
  int s; uint32_t ph;
  __device__ __forceinline__ void init(int n, int depth) { s = 0; ph = 0; advance(n, depth); }
  __device__ __forceinline__ void advance(int by, int depth) {
    s += by;
    while (s >= depth) { s -= depth; ph ^= 1u; }
  }
};

// PM_ROWS: tabulate rows [j0, j0 + TILE_N) of the parameters q[0..NP) (scalars are broadcast, absent ones left alone: they are
// never read) -- called by all epilogue threads (etid of nthr); the caller synchronises the epilogue warps around it.
// Table layout: parameter i at tab + i * TILE_N, indexed by row - j0.
template <int NP>
__device__ __forceinline__ void fill_rowtab(const BP (&q)[NP], float* tab, int j0, int n_feat, int etid, int nthr, int nrows = TILE_N) {
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    if (q[i].p == nullptr) continue;
    for (int r = etid; r < nrows; r += nthr) {
      const int row = j0 + r;
      tab[i * TILE_N + r] = row < n_feat ? __ldg(q[i].p + (i64)row * q[i].rs) : 0.f;
    }
  }
}

// Shared memory is split between the operand stages of the mainloop and the staging ring of the epilogue inputs; an epilogue
// functor may trade one for the other with `static constexpr int OP_STAGES` (default: SmemPlan::STAGES = 3).  The dV epilogue of the
// backward is HBM-latency-bound on its seven staged arrays with the tensor pipe 19 % busy (ncu): it runs on 2 operand stages and a
// 120 KB ring.
template <class E, class = void> struct op_stages_of { static constexpr int value = UMMA_STAGES; };
template <class E> struct op_stages_of<E, decltype((void)E::OP_STAGES)> { static constexpr int value = E::OP_STAGES; };
template <class Epi, int NPASS, int KC>
constexpr int ring_bytes_of() { return RING_BYTES + (UMMA_STAGES - op_stages_of<Epi>::value) * SmemPlan<NPASS, KC>::STAGE_BYTES; }

template <class Epi, int NPASS, int KC>
__global__ void __launch_bounds__(roles_threads(Epi::WARPS), 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB_big,
                 const __grid_constant__ CUtensorMap tmB_small, const __grid_constant__ EMaps emaps, GemmShape gs, Epi epi) {
  using Plan = SmemPlan<NPASS, KC>;
  constexpr int STAGES = op_stages_of<Epi>::value;
  constexpr int RING_TOTAL = RING_BYTES + (Plan::STAGES - STAGES) * Plan::STAGE_BYTES;
  static_assert(STAGES >= 2 && STAGES <= Plan::STAGES, "operand stages");
  constexpr int EPI_WARPS = Epi::WARPS, EPI_PARTS = EPI_WARPS / 4;
  constexpr int SPLIT_WARP0 = EPI_WARP0 + EPI_WARPS, EIN_WARP = SPLIT_WARP0 + SPLIT_WARPS;
  static_assert(EPI_WARPS == 8 || EPI_WARPS == 16, "epilogue warps: 2 or 4 per TMEM lane quadrant");
  constexpr bool BF = NPASS == 2;                   // bf16 operands (kind::f16)
  constexpr int EB = Plan::EB;
  constexpr int MMA_K = 32 / EB;                    // 8 (tf32) or 16 (bf16): 32 bytes of K per instruction
  static_assert(KC * EB == 128 || KC * EB == 64, "weight rows of one chunk are 64 or 128 bytes");
  constexpr uint32_t B_LAYOUT = KC * EB == 128 ? LAYOUT_SW128 : LAYOUT_SW64;
  constexpr uint32_t B_SBO = 8 * KC * EB;           // 8 rows of KC elements
  constexpr int A_BOX_COLS = 128 / EB;              // batch columns of one activation box: 128-byte rows
  constexpr int CHK = Epi::CHUNK;
  constexpr int SUB_BYTES = CHK * TILE_B * 4;       // one staged array of one chunk
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment as an offset from the __shared__ symbol: keeps the address space visible to the compiler (LDS/STS, not generic)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem + STAGES * Plan::STAGE_BYTES;
  uint64_t* bars = (uint64_t*)(ring + RING_TOTAL);
  uint64_t* full = bars;                   // [STAGES]
  uint64_t* empty = bars + STAGES;         // [STAGES]
  uint64_t* ready = bars + 2 * STAGES;     // [STAGES] activation split done (3-pass mode)
  uint64_t* tfull = bars + 3 * STAGES;     // [2]
  uint64_t* tempty = bars + 3 * STAGES + 2;  // [2]
  uint64_t* efull = bars + 3 * STAGES + 4;                   // [MAX_RING_DEPTH]
  uint64_t* eempty = bars + 3 * STAGES + 4 + MAX_RING_DEPTH;  // [MAX_RING_DEPTH]
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * STAGES + 4 + 2 * MAX_RING_DEPTH);
  float* rowtab = (float*)(ring + RING_TOTAL + Plan::BAR_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const i64 ntiles = gs.n_tiles;
  // staging ring geometry (uniform over the CTA)
  const int nin = __popc(epi.in_mask & ~EIN_MASK_BIT);
  const bool mk_staged = (epi.in_mask & EIN_MASK_BIT) != 0;
  const int slot_bytes = nin * SUB_BYTES + (mk_staged ? CHK * TILE_B : 0);
  // The depth is a multiple of EPI_PARTS, so that a ring slot is always consumed by the same part: mbarrier waits only
  // see the phase PARITY, and a part that shared a slot with another part could reach the slot's wrap w+2 while wrap w+1
  // (the other part's) was still in flight -- its wait would pass on the stale phase.  (The host refuses depth < EPI_PARTS.)
  int depth = nin > 0 ? RING_TOTAL / slot_bytes : EPI_PARTS;
  if (depth > MAX_RING_DEPTH) depth = MAX_RING_DEPTH;
  depth -= depth % EPI_PARTS;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB_big);
    if (NPASS >= 3) prefetch_tmap(&tmB_small);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], SPLIT_WARPS); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], EPI_WARPS); }
    for (int s = 0; s < depth; ++s) { mbar_init(&efull[s], 1); mbar_init(&eempty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // The next product kernel of the stream may start its own prologue (barriers, TMEM, tensor-map prefetch) on SMs this
  // grid has left; this grid's global reads and writes all happen after the wait below (the MMA issuer and the splitters
  // touch shared memory and TMEM only, downstream of the producers' loads).
  grid_dep_launch_dependents();
  if (warp != 1 && !(warp >= SPLIT_WARP0 && warp < EIN_WARP)) grid_dep_wait();

  if (warp == 0) {
    // ===== TMA producer (MMA operands): the whole warp walks the schedule, one elected lane issues =====
    int s = 0; uint32_t ph = 0; int trp = 0; (void)trp;
    for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const TileInfo ti = decode_tile(gs, tile);
      const int b0 = (int)(ti.bt * TILE_B);          // batch column (TMA coordinates are 32-bit)
      const int j0 = ti.j0;                          // (a half tile still loads the 256-row weight box; its MMAs read 128 rows)
      for (int kc = 0; kc < gs.k_chunks; ++kc) {
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
          UMMA_TR(gs, trp, 0); 
          uint8_t* st = smem + s * Plan::STAGE_BYTES;
          mbar_expect_tx(&full[s], Plan::TX_BYTES);
          // activation (raw fp32 / bf16): boxes of (32 / 64 batch columns x KC rows), 128 B per row
#pragma unroll
          for (int g = 0; g < TILE_B / A_BOX_COLS; ++g)
            tma_load_2d(st + g * (KC * 128), &tmA, &full[s], b0 + g * A_BOX_COLS, kc * KC);
          // weights, pre-split by the prep kernel: one box of (KC k x 256 rows) per part
          uint8_t* b_dst = st + Plan::NOPS * Plan::A_BYTES;
          tma_load_2d(b_dst, &tmB_big, &full[s], kc * KC, j0);
          if (NPASS >= 3) tma_load_2d(b_dst + Plan::B_BYTES, &tmB_small, &full[s], kc * KC, j0);
          UMMA_TR(gs, trp, 1);
        }
        ++trp;
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: converged warp, one elected lane issues; descriptors are precomputed, only `lo` moves =====
    constexpr uint32_t idesc_full = make_idesc(TILE_B, TILE_N, /*A MN-major*/ 1, /*B K-major*/ 0, BF ? 1u : 2u);
    constexpr uint32_t idesc_half = make_idesc(TILE_B, TILE_N / 2, 1, 0, BF ? 1u : 2u);
    // A (MN-major, 128B swizzle / 32B atoms): rows of 32 batch columns (128 B); one k-step = 8 rows = 2 atoms (SBO = 512 B
    // apart); batch groups of 32 are KC*128 B apart (LBO).  B (K-major): rows of KC floats; 8-row groups B_SBO apart;
    // a k-step advances 32 B inside the swizzled row.
    // (bf16: 64 batch columns per 128-byte row, plain 128B swizzle, atoms of 8 rows = 1024 B; one k-step = 16 rows = 2 atoms)
    constexpr uint32_t a_hi = BF ? desc_hi(1024, LAYOUT_SW128) : desc_hi(A_ATOM_BYTES, LAYOUT_SW128_BASE32B);
    constexpr uint32_t A_KSTEP = MMA_K * 128;         // bytes between consecutive k-steps of the activation tile
    constexpr uint32_t b_hi = desc_hi(B_SBO, B_LAYOUT);
    const uint32_t st0 = smem_u32(smem);
    const uint32_t a_lo0 = desc_lo(st0, KC * 128);
    const uint32_t b_lo0 = desc_lo(st0 + Plan::NOPS * Plan::A_BYTES, 16);
    int s = 0; uint32_t ph = 0; int trm = 0; (void)trm;
    int acc = 0; uint32_t aph = 0;
    for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const uint32_t idesc = decode_tile(gs, tile).nrows == TILE_N ? idesc_full : idesc_half;
      mbar_wait(&tempty[acc], aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * TILE_N;
      for (int kc = 0; kc < gs.k_chunks; ++kc, ++trm) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + s * (Plan::STAGE_BYTES >> 4), b_lo = b_lo0 + s * (Plan::STAGE_BYTES >> 4);
        // The passes that read the raw activation words (big * w_small, big * w_big) need no split: they are issued as soon
        // as the TMA bytes land and run while the splitters work on the stage; only small * w_big waits for `ready`.
        // (Measured on the store-only probe: 178 -> 200 TFLOP/s at K=500, 190 -> 206 at K=250.)
        if (elect_one()) {
          UMMA_TR(gs, trm, 4);
#pragma unroll
          for (int ks = 0; ks < KC / MMA_K; ++ks) {
            const uint64_t da_big = desc_at(a_hi, a_lo + ks * (A_KSTEP >> 4));
            const uint64_t db_big = desc_at(b_hi, b_lo + ks * (32 >> 4));
            const uint32_t first = (kc == 0 && ks == 0) ? 0u : 1u;
            if (NPASS == 3) {
              const uint64_t db_small = desc_at(b_hi, b_lo + (Plan::B_BYTES >> 4) + ks * (32 >> 4));
              umma_tf32(d_tmem, da_big, db_small, idesc, first);
              umma_tf32(d_tmem, da_big, db_big, idesc, 1u);
            } else if (NPASS == 4) {
              umma_tf32(d_tmem, da_big, db_big, idesc, first);       // trunc_tf32(x) * rna_tf32(w): reads the raw words
            } else if (NPASS == 2) {
              umma_op<BF>(d_tmem, da_big, db_big, idesc, first);
            }
          }
        }
        __syncwarp();
        if (NPASS == 3 || NPASS == 1 || NPASS == 4) {
          mbar_wait(&ready[s], ph);                    // small part written (3 passes) / tile rounded to tf32 (1 pass) by the splitters
          tc_fence_after();
        }
        if (elect_one()) {
          if (NPASS == 4) {
            // the two correction products on kind::f16 (K = 16 = the whole chunk): bf16(x) * bf16(w_small), bf16(x_small) * bf16(w_big).
            // Operand images: [raw fp32 8 KB | bf16(x) 4 KB | bf16(x_small) 4 KB] and, behind the tf32 weights, 64-byte rows of
            // [bf16(w_big)[16] | bf16(w_small)[16]] written by the prep kernel (same box, same SWIZZLE_64B as the tf32 part).
            constexpr uint32_t a16_hi = desc_hi(1024, LAYOUT_SW128);
            constexpr uint32_t idesc16 = make_idesc(TILE_B, TILE_N, 1, 0, 1u), idesc16_half = make_idesc(TILE_B, TILE_N / 2, 1, 0, 1u);
            const uint32_t id16 = idesc == idesc_full ? idesc16 : idesc16_half;
            const uint32_t a16 = a_lo + (Plan::A_BYTES >> 4), b16 = b_lo + (Plan::B_BYTES >> 4);
            umma_f16(d_tmem, desc_at(a16_hi, a16), desc_at(b_hi, b16 + (32 >> 4)), id16, 1u);
            umma_f16(d_tmem, desc_at(a16_hi, a16 + (Plan::A_BYTES >> 5)), desc_at(b_hi, b16), id16, 1u);
          }
          if (NPASS == 1) {
#pragma unroll
            for (int ks = 0; ks < KC / MMA_K; ++ks)
              umma_tf32(d_tmem, desc_at(a_hi, a_lo + ks * (A_KSTEP >> 4)), desc_at(b_hi, b_lo + ks * (32 >> 4)), idesc,
                        (kc == 0 && ks == 0) ? 0u : 1u);
          }
          if (NPASS == 3) {
#pragma unroll
            for (int ks = 0; ks < KC / MMA_K; ++ks) {
              const uint64_t da_small = desc_at(a_hi, a_lo + (Plan::A_BYTES >> 4) + ks * (A_KSTEP >> 4));
              const uint64_t db_big = desc_at(b_hi, b_lo + ks * (32 >> 4));
              umma_tf32(d_tmem, da_small, db_big, idesc, 1u);
            }
          }
          umma_commit(&empty[s]);                      // frees the smem stage when these MMAs retire
          if (kc == gs.k_chunks - 1) umma_commit(&tfull[acc]);
          UMMA_TR(gs, trm, 5);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  } else if (warp == EIN_WARP) {
    // ===== TMA producer (epilogue inputs): chunk n = ((tile_iter * NCH + c) * EPI_PARTS + part) -> ring slot n % depth =====
    if (nin > 0) {
      RingPos rp; rp.init(0, depth);
      for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileInfo ti = decode_tile(gs, tile);
        const int b0 = (int)(ti.bt * TILE_B);
        const int j0 = ti.j0;
        const int rpw = ti.nrows / EPI_PARTS;         // feature rows per (tile, part)
        const int nch = rpw / CHK;
        for (int c = 0; c < nch; ++c) {
          for (int h = 0; h < EPI_PARTS; ++h, rp.advance(1, depth)) {
            const int s = rp.s;
            mbar_wait(&eempty[s], rp.ph ^ 1);
            const int row0 = j0 + h * rpw + c * CHK;
            if (elect_one()) {
              if (row0 >= gs.n_feat) {
                mbar_arrive(&efull[s]);                  // nothing to stage, keep the phases in step
              } else {
                uint8_t* dst = ring + s * slot_bytes;
                mbar_expect_tx(&efull[s], slot_bytes);
#pragma unroll
                for (int i = 0; i < Epi::NIN; ++i) {
                  if (epi.in_mask & (1u << i)) {
                    tma_load_2d(dst, &emaps.m[i], &efull[s], b0, row0);
                    dst += SUB_BYTES;
                  }
                }
                if (mk_staged) tma_load_2d(dst, &emaps.mk, &efull[s], b0, row0);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp >= SPLIT_WARP0) {
    // ===== operand splitters (3-pass mode): small = x - trunc_tf32(x) of the activation tile, in shared memory =====
    if (NPASS == 3 || NPASS == 1 || NPASS == 4) {
      const int tid = threadIdx.x - SPLIT_WARP0 * 32;
      int s = 0; uint32_t ph = 0; int trs = 0; (void)trs;
      for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int kc = 0; kc < gs.k_chunks; ++kc, ++trs) {
          mbar_wait(&full[s], ph);                       // TMA bytes landed; the stage was free (producer waited on empty)
          if (tid == 0) UMMA_TR(gs, trs, 2);
          uint8_t* st = smem + s * Plan::STAGE_BYTES;
          if (NPASS == 3) split_tile<Plan::A_BYTES, SPLIT_WARPS * 32>(st, st + Plan::A_BYTES, tid);
          else if (NPASS == 4) split_tile_mix<SPLIT_WARPS * 32>(st, st + Plan::A_BYTES, st + Plan::A_BYTES + Plan::A_BYTES / 2, tid);
          else round_tile<Plan::A_BYTES, SPLIT_WARPS * 32>(st, tid);
          fence_proxy_async();                           // generic-proxy smem writes -> visible to the tensor core
          __syncwarp();
          if (tid == 0) UMMA_TR(gs, trs, 3);
          if (lane == 0) mbar_arrive(&ready[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3;                            // TMEM lane quadrant this warp may access
    const int half = (warp - EPI_WARP0) >> 2;          // which part (1/EPI_PARTS) of the tile's feature rows
    const int col = q * 32 + lane;                     // column inside the tile
    int acc = 0; uint32_t aph = 0;
    typename Epi::State state;
    epi.begin(state);
    int tab_j0 = -1;                                   // PM_ROWS: the feature tile the parameter table currently holds
    RingPos rp; rp.init(half, depth);                  // this part's chunks are every EPI_PARTS-th slot of the staging ring
    for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const TileInfo ti = decode_tile(gs, tile);
      const i64 bt = ti.bt;
      const i64 b = bt * TILE_B + col;
      const bool valid = b < gs.B;
      const int rpw = ti.nrows / EPI_PARTS;            // feature rows per (tile, part); half as many in a half tile
      const int nch = rpw / CHK;
      const int jw = ti.j0 + half * rpw;               // first feature row of this warp
      if constexpr (Epi::NROWP > 0) {
        // per-row parameters of this tile's 256 feature rows -> shared memory (the epilogue warps only: named barrier 1);
        // once per launch when the product has a single feature tile
        static_assert(Epi::NROWP * TILE_N * 4 <= Plan::ROWTAB, "row-parameter table");
        if (ti.j0 != tab_j0) {                          // uniform over the epilogue warps
          asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");      // the previous tile's readers are done
          BP q[Epi::NROWP];
          epi.row_params(q);
          fill_rowtab<Epi::NROWP>(q, rowtab, ti.j0, gs.n_feat, threadIdx.x - EPI_WARP0 * 32, EPI_WARPS * 32);
          asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
          epi.bind_rows(state, rowtab - ti.j0, TILE_N);
          tab_j0 = ti.j0;
        }
      }
      typename Epi::Pre pre;
      epi.prefetch(pre, jw, b, valid, gs.n_feat);
      mbar_wait(&tfull[acc], aph);
      tc_fence_after();
      const uint32_t t0 = tmem_base + acc * TILE_N + half * rpw + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < nch; ++c, rp.advance(EPI_PARTS, depth)) {
        const int row0 = jw + c * CHK;
        typename Epi::Pre pre_next;
        epi.prefetch(pre_next, row0 + CHK, b, valid && c + 1 < nch, gs.n_feat);
        const int s = rp.s;
        if (nin > 0) mbar_wait(&efull[s], rp.ph);
        if (row0 < gs.n_feat) {                          // warp-uniform
          float v[CHK];
          tmem_ld(t0 + c * CHK, v);
#pragma unroll
          for (int i = 0; i < CHK; ++i) v[i] = __fmul_rn(v[i], gs.acc_scale);
          const float* slot = reinterpret_cast<const float*>(ring + s * slot_bytes);
          if (row0 + CHK <= gs.n_feat)                   // every row of the chunk exists: branch-free rows, interleaved by the compiler
            epi.template apply<true>(state, slot, col, pre, row0, b, valid, v, gs.n_feat, bt * (TILE_B / 32) + q);
          else
            epi.template apply<false>(state, slot, col, pre, row0, b, valid, v, gs.n_feat, bt * (TILE_B / 32) + q);
        }
        if (nin > 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&eempty[s]);
        }
        pre = pre_next;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
    epi.end(state, (int)blockIdx.x * EPI_WARPS + (warp - EPI_WARP0), lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ---- host side: tile list ---------------------------------------------------------------------------------------
// Persistent CTAs take tiles round-robin, so the kernel lasts as long as its busiest CTA: ceil(tiles / grid) tile times.
// When the last round is at most half full, its batch tiles are cut into 128-row halves (a half costs ~0.55 of a tile:
// half the MMAs, the same operand loads) so that twice as many CTAs share it.
inline void plan_tiles(GemmShape& gs, int grid) {
  const i64 ntiles = gs.n_btiles * gs.n_ntiles;
  gs.n_full = ntiles; gs.n_tiles = ntiles;
  if (grid <= 0 || ntiles <= grid) return;
  const i64 rounds = ntiles / grid;
  i64 n_full = rounds * grid;
  n_full -= n_full % gs.n_ntiles;                 // whole batch tiles only
  const i64 rest = ntiles - n_full;               // tiles that would form the last round
  if (rest == 0) return;
  auto cost = [&](i64 nf, i64 nh) {               // longest CTA, in tile times
    double worst = 0;
    for (int c = 0; c < grid; ++c) {
      const i64 f = nf > c ? (nf - c + grid - 1) / grid : 0;
      const i64 t = nf + nh > c ? (nf + nh - c + grid - 1) / grid : 0;
      worst = std::max(worst, (double)f + 0.55 * (double)(t - f));
    }
    return worst;
  };
  if (cost(n_full, 2 * rest) + 0.1 < cost(ntiles, 0)) { gs.n_full = n_full; gs.n_tiles = n_full + 2 * rest; }
}

// ---- host side: tensor maps ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2D fp32 row-major matrix (rows x cols, pitch in elements); box = (box_cols x box_rows); OOB reads give zeros
// 2D uint8 row-major matrix (prox masks), no swizzle
inline int make_tmap_2d_u8(CUtensorMap* out, const uint8_t* base, i64 rows, i64 cols, i64 pitch, int box_cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return DLADMM_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (u8) failed (%d) rows=%lld cols=%lld pitch=%lld", (int)r, rows, cols, pitch);
    return DLADMM_ERR_CUDA;
  }
  return DLADMM_OK;
}

inline int make_tmap_2d(CUtensorMap* out, const float* base, i64 rows, i64 cols, i64 pitch, int box_cols, int box_rows,
                        CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return DLADMM_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld pitch=%lld box=%dx%d", (int)r, rows, cols, pitch, box_cols,
              box_rows);
    return DLADMM_ERR_CUDA;
  }
  return DLADMM_OK;
}

// 3D fp32: `slabs` row-major (rows x cols) matrices `slab_stride` elements apart; box = (box_cols x box_rows x 1)
inline int make_tmap_3d(CUtensorMap* out, const float* base, i64 slabs, i64 rows, i64 cols, i64 pitch, i64 slab_stride, int box_cols,
                        int box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return DLADMM_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)slabs};
  cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)slab_stride * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3D) failed (%d) slabs=%lld rows=%lld cols=%lld pitch=%lld stride=%lld box=%dx%d", (int)r, slabs, rows,
              cols, pitch, slab_stride, box_cols, box_rows);
    return DLADMM_ERR_CUDA;
  }
  return DLADMM_OK;
}

// 2D bf16 row-major matrix (rows x cols, pitch in ELEMENTS, a multiple of 8)
inline int make_tmap_2d_bf16(CUtensorMap* out, const void* base, i64 rows, i64 cols, i64 pitch, int box_cols, int box_rows,
                             CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return DLADMM_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (bf16) failed (%d) rows=%lld cols=%lld pitch=%lld box=%dx%d", (int)r, rows, cols, pitch,
              box_cols, box_rows);
    return DLADMM_ERR_CUDA;
  }
  return DLADMM_OK;
}

}  // namespace umma
}  // namespace dladmm
