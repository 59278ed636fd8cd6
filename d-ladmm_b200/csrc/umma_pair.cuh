// CTA-pair variant of the fused product kernel (umma_gemm.cuh) for the split-precision mode: tcgen05.mma.cta_group::2.
//
// Two CTAs of one cluster (two SMs of a TPC) compute M = 256 batch columns (128 per CTA) x N = 256 feature rows per instruction.  Each
// CTA stages its OWN activation tile and HALF of the weight tile (128 of the 256 rows; the tensor cores read both halves), keeps its
// own accumulator in its own TMEM and runs the unchanged epilogue on its own 128 columns.  Per CTA and 16-deep k-chunk that is 72 KB of
// shared-memory traffic instead of 104 (TMA writes 24 instead of 40, operand reads 32 instead of 48) and half the L2 -> SM weight
// bytes -- the single-CTA mainloop is bound by exactly that.  Store-only probe (tools/umma_pair_probe.cu): 518-576
// cycles per 256-column chunk against 2 x 898-945, 374 against 284 TFLOP/s at the large-scale shape (K = 2000).
//
// Hand-offs (every CTA has the same barrier block; `leader` = cluster rank 0 issues all MMAs):
//   afull[s]  local   this CTA's activation tile of stage s has landed (TMA)            -> this CTA's splitter warps
//   bfull[s]  leader  both weight halves have landed: the peer's TMA signals the leader's barrier directly (cta_group::2 TMA)
//   ready[s]  leader  both CTAs' splitters are done with stage s (remote mbarrier.arrive from the peer, CTA-scope release -- the
//                     arrive of cutlass::arch::ClusterBarrier::arrive(cta_id); a cluster-scope release costs ~1500 cycles per arrive
//                     and made the splitters the bottleneck: 2200 cycles per chunk in the probe)
//   empty[s]  local   the MMAs that read stage s have retired: tcgen05.commit multicast to both CTAs  -> both producers
//   tfull[a]  local   accumulator a is complete (multicast commit)                     -> this CTA's epilogue warps
//   tempty[a] leader  both CTAs' epilogue warps have drained accumulator a (remote arrive)
// Used for mainloop-bound products (reduction length >= PAIR_MIN_K: the large-scale shape); the HBM-bound shapes keep the single-CTA
// kernels and the all-layer persistent schedule.
#pragma once
#include "umma_gemm.cuh"

namespace dladmm {
namespace umma {

constexpr int PAIR_KC = 16;
constexpr int PAIR_A = TILE_B * PAIR_KC * 4;              // 8 KB raw activation tile (+ 4 + 4 KB bf16 hi / lo)
constexpr int PAIR_BH = (TILE_N / 2) * PAIR_KC * 4;       // 8 KB: this CTA's 128 rows of the tf32 weights (+ 8 KB packed bf16)
constexpr int PAIR_STAGE = 2 * PAIR_A + 2 * PAIR_BH;      // 32 KB
constexpr int PAIR_BAR_BYTES = 1024, PAIR_ROWTAB = 8192;
// Operand stages against staging ring: the pair kernel serves mainloop-bound products, where the ring only has to hold two slots per
// epilogue part.  8 epilogue warps (the forward epilogues: 2 parts, slots of 8-20 KB): 40 KB ring, 5 stages; 16 warps (the backward
// epilogues: 4 parts, slots up to 14.8 KB): the 72 KB ring, 4 stages.
template <int EPI_WARPS> struct PairPlan {
  static constexpr int RING = EPI_WARPS == 8 ? 40 * 1024 : RING_BYTES;
  static constexpr int STAGES = EPI_WARPS == 8 ? 5 : 4;
  static constexpr int SMEM = STAGES * PAIR_STAGE + RING + PAIR_BAR_BYTES + PAIR_ROWTAB + 1024;
  static_assert(SMEM <= 227 * 1024, "shared memory of the pair kernel");
};
constexpr int PAIR_MIN_K = 768;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {      // acquire at cluster scope (remote arrivals)
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)DLADMM_MBAR_SUSPEND_NS) : "memory");
  } while (!done);
}
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

template <class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(roles_threads(Epi::WARPS), 1)
umma_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB_big,
                      const __grid_constant__ CUtensorMap tmB_pack, const __grid_constant__ EMaps emaps, GemmShape gs, Epi epi) {
  constexpr int EPI_WARPS = Epi::WARPS, EPI_PARTS = EPI_WARPS / 4;
  constexpr int STAGES = PairPlan<EPI_WARPS>::STAGES, RING = PairPlan<EPI_WARPS>::RING;
  constexpr int SPLIT_WARP0 = EPI_WARP0 + EPI_WARPS, EIN_WARP = SPLIT_WARP0 + SPLIT_WARPS;
  constexpr int CHK = Epi::CHUNK;
  constexpr int SUB_BYTES = CHK * TILE_B * 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem + STAGES * PAIR_STAGE;
  uint64_t* bars = (uint64_t*)(ring + RING);
  uint64_t* afull = bars;
  uint64_t* bfull = bars + STAGES;
  uint64_t* ready = bars + 2 * STAGES;
  uint64_t* empty = bars + 3 * STAGES;
  uint64_t* tfull = bars + 4 * STAGES;
  uint64_t* tempty = bars + 4 * STAGES + 2;
  uint64_t* efull = bars + 4 * STAGES + 4;
  uint64_t* eempty = bars + 4 * STAGES + 4 + MAX_RING_DEPTH;
  uint32_t* tmem_slot = (uint32_t*)(bars + 4 * STAGES + 4 + 2 * MAX_RING_DEPTH);
  static_assert((4 * STAGES + 4 + 2 * MAX_RING_DEPTH) * 8 + 8 <= PAIR_BAR_BYTES, "barrier block");
  float* rowtab = (float*)(ring + RING + PAIR_BAR_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const i64 n_ptiles = ((gs.n_btiles + 1) / 2) * gs.n_ntiles;     // (pair of batch tiles) x feature tile
  const int npair = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int nin = __popc(epi.in_mask & ~EIN_MASK_BIT);
  const bool mk_staged = (epi.in_mask & EIN_MASK_BIT) != 0;
  const int slot_bytes = nin * SUB_BYTES + (mk_staged ? CHK * TILE_B : 0);
  int depth = nin > 0 ? RING / slot_bytes : EPI_PARTS;
  if (depth > MAX_RING_DEPTH) depth = MAX_RING_DEPTH;
  depth -= depth % EPI_PARTS;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB_big); prefetch_tmap(&tmB_pack);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&afull[s], 1); mbar_init(&bfull[s], 1); mbar_init(&ready[s], 2 * SPLIT_WARPS); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 2 * EPI_WARPS); }
    for (int s = 0; s < depth; ++s) { mbar_init(&efull[s], 1); mbar_init(&eempty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // both CTAs' barriers exist before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_launch_dependents();
  if (warp != 1 && !(warp >= SPLIT_WARP0 && warp < EIN_WARP)) grid_dep_wait();

  if (warp == 0) {
    // ===== TMA producer: own activation tile -> local afull; own half of the weights -> the leader's bfull =====
    int s = 0; uint32_t ph = 0;
    for (i64 pt = pair; pt < n_ptiles; pt += npair) {
      const i64 bt = 2 * (pt / gs.n_ntiles) + rank;
      const int j0 = (int)(pt % gs.n_ntiles) * TILE_N;
      const int b0 = (int)(bt * TILE_B);               // (a batch tile past the end loads zeros: odd number of batch tiles)
      for (int kc = 0; kc < gs.k_chunks; ++kc) {
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
          uint8_t* st = smem + s * PAIR_STAGE;
          mbar_expect_tx(&afull[s], PAIR_A);
#pragma unroll
          for (int g = 0; g < 4; ++g) tma_load_2d(st + g * (PAIR_KC * 128), &tmA, &afull[s], b0 + g * 32, kc * PAIR_KC);
          if (leader) mbar_expect_tx(&bfull[s], 4 * PAIR_BH);          // both halves: tf32 + packed bf16 each
          const uint32_t bf = mapa_rank(smem_u32(&bfull[s]), 0);
          tma_load_2d_pair(st + 2 * PAIR_A, &tmB_big, bf, kc * PAIR_KC, j0 + (int)rank * (TILE_N / 2));
          tma_load_2d_pair(st + 2 * PAIR_A + PAIR_BH, &tmB_pack, bf, kc * PAIR_KC, j0 + (int)rank * (TILE_N / 2));
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===== MMA issuer (leader CTA only): M = 256 over the pair =====
      constexpr uint32_t idesc32 = make_idesc(256, TILE_N, 1, 0, 2u), idesc16 = make_idesc(256, TILE_N, 1, 0, 1u);
      constexpr uint32_t a_hi = desc_hi(A_ATOM_BYTES, LAYOUT_SW128_BASE32B), a16_hi = desc_hi(1024, LAYOUT_SW128);
      constexpr uint32_t b_hi = desc_hi(8 * PAIR_KC * 4, LAYOUT_SW64);
      const uint32_t st0 = smem_u32(smem);
      const uint32_t a_lo0 = desc_lo(st0, PAIR_KC * 128), b_lo0 = desc_lo(st0 + 2 * PAIR_A, 16);
      int s = 0; uint32_t ph = 0, acc = 0, aph = 0;
      for (i64 pt = pair; pt < n_ptiles; pt += npair) {
        mbar_wait_cluster(&tempty[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TILE_N;
        for (int kc = 0; kc < gs.k_chunks; ++kc) {
          mbar_wait_cluster(&bfull[s], ph);
          mbar_wait_cluster(&ready[s], ph);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + s * (PAIR_STAGE >> 4), b_lo = b_lo0 + s * (PAIR_STAGE >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma2_tf32(d_tmem, desc_at(a_hi, a_lo + ks * ((8 * 128) >> 4)), desc_at(b_hi, b_lo + ks * 2), idesc32, (kc == 0 && ks == 0) ? 0u : 1u);
            const uint32_t a16 = a_lo + (PAIR_A >> 4), b16 = b_lo + (PAIR_BH >> 4);
            umma2_f16(d_tmem, desc_at(a16_hi, a16), desc_at(b_hi, b16 + 2), idesc16, 1u);                    // bf16(x) * bf16(w_small)
            umma2_f16(d_tmem, desc_at(a16_hi, a16 + (PAIR_A >> 5)), desc_at(b_hi, b16), idesc16, 1u);        // bf16(x_small) * bf16(w_big)
            umma2_commit(&empty[s]);
            if (kc == gs.k_chunks - 1) umma2_commit(&tfull[acc]);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (++acc == 2) { acc = 0; aph ^= 1; }
      }
    }
  } else if (warp == EIN_WARP) {
    // ===== TMA producer (epilogue inputs): as in umma_gemm_kernel, for this CTA's batch tile =====
    if (nin > 0) {
      RingPos rp; rp.init(0, depth);
      for (i64 pt = pair; pt < n_ptiles; pt += npair) {
        const i64 bt = 2 * (pt / gs.n_ntiles) + rank;
        const int j0 = (int)(pt % gs.n_ntiles) * TILE_N;
        const int b0 = (int)(bt * TILE_B);
        const int rpw = TILE_N / EPI_PARTS;
        const int nch = rpw / CHK;
        for (int c = 0; c < nch; ++c) {
          for (int h = 0; h < EPI_PARTS; ++h, rp.advance(1, depth)) {
            const int s = rp.s;
            mbar_wait(&eempty[s], rp.ph ^ 1);
            const int row0 = j0 + h * rpw + c * CHK;
            if (elect_one()) {
              if (row0 >= gs.n_feat || bt >= gs.n_btiles) {
                mbar_arrive(&efull[s]);
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
    // ===== splitters: bf16 hi / lo of this CTA's activation tile; every warp arrives on the LEADER's ready barrier =====
    const int tid = threadIdx.x - SPLIT_WARP0 * 32;
    int s = 0; uint32_t ph = 0;
    for (i64 pt = pair; pt < n_ptiles; pt += npair) {
      for (int kc = 0; kc < gs.k_chunks; ++kc) {
        mbar_wait(&afull[s], ph);
        uint8_t* st = smem + s * PAIR_STAGE;
        split_tile_mix<SPLIT_WARPS * 32>(st, st + PAIR_A, st + PAIR_A + PAIR_A / 2, tid);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(mapa_rank(smem_u32(&ready[s]), 0));
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===== epilogue warps: the functors of umma_epilogues.cuh / umma_bwd.cuh, unchanged, on this CTA's 128 columns =====
    const int q = warp & 3;
    const int half = (warp - EPI_WARP0) >> 2;
    const int col = q * 32 + lane;
    uint32_t acc = 0, aph = 0;
    typename Epi::State state;
    epi.begin(state);
    int tab_j0 = -1;
    RingPos rp; rp.init(half, depth);
    for (i64 pt = pair; pt < n_ptiles; pt += npair) {
      const i64 bt = 2 * (pt / gs.n_ntiles) + rank;
      const int j0 = (int)(pt % gs.n_ntiles) * TILE_N;
      const i64 b = bt * TILE_B + col;
      const bool valid = b < gs.B;
      const int rpw = TILE_N / EPI_PARTS;
      const int nch = rpw / CHK;
      const int jw = j0 + half * rpw;
      if constexpr (Epi::NROWP > 0) {
        static_assert(Epi::NROWP * TILE_N * 4 <= PAIR_ROWTAB, "row-parameter table");
        if (j0 != tab_j0) {
          asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
          BP qv[Epi::NROWP];
          epi.row_params(qv);
          fill_rowtab<Epi::NROWP>(qv, rowtab, j0, gs.n_feat, threadIdx.x - EPI_WARP0 * 32, EPI_WARPS * 32);
          asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
          epi.bind_rows(state, rowtab - j0, TILE_N);
          tab_j0 = j0;
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
        if (row0 < gs.n_feat && bt < gs.n_btiles) {     // (the peer of an odd last batch tile has no columns: hand-shakes only --
          float v[CHK];                                 //  its column-group index would lie outside the gradient partials)
          tmem_ld(t0 + c * CHK, v);
#pragma unroll
          for (int i = 0; i < CHK; ++i) v[i] = __fmul_rn(v[i], gs.acc_scale);
          const float* slot = reinterpret_cast<const float*>(ring + s * slot_bytes);
          if (row0 + CHK <= gs.n_feat)
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
      if (lane == 0) mbar_arrive_remote(mapa_rank(smem_u32(&tempty[acc]), 0));
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
    epi.end(state, (int)blockIdx.x * EPI_WARPS + (warp - EPI_WARP0), lane);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // the peer may still read this CTA's shared memory / signal its barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace umma
}  // namespace dladmm
