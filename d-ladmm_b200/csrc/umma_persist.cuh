// All-layer persistent forward kernel (tcgen05 / TMEM / TMA, sm_100a): ONE launch runs T_0 and every layer's two fused
// products.  Same warp roles, operand ring, staging ring and epilogue arithmetic as umma_gemm_kernel (umma_gemm.cuh /
// umma_epilogues.cuh); what changes is the schedule:
//
//   * the work list is the concatenation of all stages -- T_0, then per layer the (batch tile x feature tile) units of
//     W V + Z-prox followed by those of A Z + E/T/L -- dealt round-robin to the persistent CTAs, so the MMAs of a CTA's next unit
//     overlap the epilogue of its previous one ACROSS stage and layer boundaries: the per-launch fill (a first mainloop with idle
//     epilogue warps) and drain (a last epilogue with an idle tensor pipe) of the 2K+1 per-layer launches are paid once;
//   * a unit of stage s only needs the outputs of stage s-1 for ITS batch tile, whichever CTA produced them: every epilogue warp
//     publishes "my rows of this tile are in global memory" on a per-(stage, batch tile) counter (release), and the two TMA
//     producers of a consumer unit wait for the counter of the stage below (acquire + async-proxy fence) before their first load.
//     Units are ordered stage-major, so a waiting CTA always waits for a unit that sits EARLIER in some resident CTA's list: no
//     deadlock as long as the grid is resident at once (grid <= SM count, one CTA per SM); a clock-bounded spin traps instead of
//     hanging if that is ever violated.
//
// Buffer reuse (the V scratch slab, the 2-slab ping-pong of last_only) is safe for the same reason: per batch tile the stages form
// a chain, and the overwriting unit sits at least one full link after the last reader.
#pragma once
#include "umma_gemm.cuh"
#include "umma_epilogues.cuh"

namespace dladmm {
namespace umma {

constexpr int PF_MAX_LAYERS = 64;
constexpr int PF_TRACE_CTAS = 4, PF_TRACE_UNITS = 512;
// stamp e of the i-th unit of this CTA: 0 unit type, 1 MMA warp got the accumulator, 2 first operands landed, 3 last MMA issued,
// 4 epilogue warp 0 saw the accumulator full, 5 epilogue warp 0 done, 6 operand producer passed the readiness wait, 7 stage | bt
#define PF_TR(i, e, val) do { if (p.trace && blockIdx.x < PF_TRACE_CTAS && (i) < PF_TRACE_UNITS) \
    p.trace[((size_t)blockIdx.x * PF_TRACE_UNITS + (i)) * 8 + (e)] = (val); } while (0)

struct BPc { const float* p; int rs; int period; };       // broadcast parameter without its gradient pointer
__device__ __forceinline__ BP to_bp(const BPc& q) { BP b; b.p = q.p; b.g = nullptr; b.rs = q.rs; b.period = q.period; return b; }

struct PfLayer { BPc b1, b2, bL, ss1, ss2, ss2_2, th1, th2; int widx; int pad; };

struct PfMaps {
  CUtensorMap A_big, A_small;          // (m256 x dp) prepared A, box (KC x 256)
  CUtensorMap W_big, W_small;          // 3D [nW][d256][mp] prepared weights, box (KC x 256 x 1)
  CUtensorMap actZ0;                   // 2D Z0 (d x B), operand box (32 x KC)
  CUtensorMap actZ;                    // 3D Z slabs [depth][d][B], operand box (32 x KC x 1)
  CUtensorMap actV;                    // 3D V slabs [1 or K][m][B]
  CUtensorMap sE0, sX, sL0;            // 2D (m x B), staging box (128 x 8)
  CUtensorMap sE, sL;                  // 3D slabs [depth][m][B], staging box (128 x 8 x 1)
  CUtensorMap sZ0;                     // 2D (d x B), staging box (128 x 16)
  CUtensorMap sZ;                      // 3D Z slabs, staging box (128 x 16 x 1)
};

struct PfParams {
  int m, d, K, last_only;
  i64 B, n_btiles;
  int tn;                              // feature rows per tile (the kernel's TN template argument)
  int nt_z, nt_e;                      // feature tiles of the W V product (ceil(d / tn)) and of the A Z product (ceil(m / tn))
  int kc_z, kc_e;                      // k-chunks: ceil(m / KC), ceil(d / KC)
  i64 units_t0, units_z, units_e, total_units;   // of the whole batch (objective bookkeeping: unit index u of the single-stream order)
  // Two CTA sets ("streams"): even CTAs own the batch tiles [0, split), odd CTAs [split, n_btiles); each set walks its own
  // stage-major list.  The second set starts `skew` clocks late, so that while one half of the SMs is in the HBM-bound A Z phase the
  // other half is in the tensor-bound W V phase (chip-wide HBM demand stays level instead of alternating).  nstreams = 1: one list.
  int nstreams; i64 split; long long skew;
  const float *X, *E0, *L0, *Z0;
  float *Z, *E, *L, *T, *V;
  uint8_t *maskZ, *maskE;
  i64 zs, ms;                          // slab strides in elements
  int v_per_layer;                     // 1: V slab k holds V_k (Vsave), 0: one scratch slab
  int obj_kind;
  float* obj_part;                     // [total_units][8 epilogue warps] partial sums of the fused objective, or NULL
  unsigned* flags;                     // [2K+1][n_btiles] readiness counters, zeroed by the host before the launch
  long long spin_limit;                // clock64 ticks a producer may wait for a counter before it traps
  int prefetch;                        // D > 0: the staging producer requests every chunk's boxes into L2 D chunks ahead
  int x_resident;                      // 1: loads of X carry an L2 evict_last policy
  unsigned scout_sleep_ns;             // pause between two polls of the scout lane's readiness spin
  float acc_scale_z, acc_scale_e;      // accumulator compensation of the W V / A Z products (umma_gemm.cuh, ACC_RZ_BIAS_PER_MMA)
  long long* trace;                    // debugging (DLADMM_PF_TRACE=1): [PF_TRACE_CTAS][PF_TRACE_UNITS][8] clock64 stamps, or NULL
  PfLayer layer[PF_MAX_LAYERS];
};

enum { PF_T0 = 0, PF_Z = 1, PF_E = 2 };
struct PfUnit { int type; int k; int stage; i64 bt; int j0; i64 u; };

// the walk of one CTA: its stream (set of batch tiles), its position in the stream's round-robin.  32-bit arithmetic (the host
// only takes this path when the unit count fits): the decode runs once per unit in every warp role, on the producers' critical path.
struct PfWalk {
  uint32_t bt0, nbt;     // batch tiles of this stream
  uint32_t u_t0, u_z, u_e, total;
  uint32_t first, stride;
  __device__ __forceinline__ void init(const PfParams& p) {
    if (p.nstreams == 2) {
      const int sidx = blockIdx.x & 1;
      bt0 = sidx ? (uint32_t)p.split : 0u; nbt = sidx ? (uint32_t)(p.n_btiles - p.split) : (uint32_t)p.split;
      first = blockIdx.x >> 1; stride = gridDim.x >> 1;
    } else {
      bt0 = 0; nbt = (uint32_t)p.n_btiles; first = blockIdx.x; stride = gridDim.x;
    }
    u_t0 = nbt * p.nt_e; u_z = nbt * p.nt_z; u_e = nbt * p.nt_e;
    total = u_t0 + (uint32_t)p.K * (u_z + u_e);
  }
  // v: index in this stream's stage-major list
  __device__ __forceinline__ PfUnit decode(const PfParams& p, uint32_t v) const {
    PfUnit r;
    const uint32_t nt_e = p.nt_e, nt_z = p.nt_z;
    if (v < u_t0) {
      const uint32_t q = v / nt_e, rem = v - q * nt_e;
      r.type = PF_T0; r.k = 0; r.stage = 0; r.bt = bt0 + q; r.j0 = (int)rem * p.tn;
      r.u = (i64)r.bt * nt_e + rem;
      return r;
    }
    const uint32_t y = v - u_t0, per = u_z + u_e;
    const uint32_t k = y / per, w = y - k * per;
    r.k = (int)k;
    const i64 layer0 = p.units_t0 + (i64)k * (p.units_z + p.units_e);       // first unit of layer k in the whole-batch numbering
    if (w < u_z) {
      const uint32_t q = w / nt_z, rem = w - q * nt_z;
      r.type = PF_Z; r.stage = 2 * r.k + 1; r.bt = bt0 + q; r.j0 = (int)rem * p.tn;
      r.u = layer0 + (i64)r.bt * nt_z + rem;
    } else {
      const uint32_t x = w - u_z, q = x / nt_e, rem = x - q * nt_e;
      r.type = PF_E; r.stage = 2 * r.k + 2; r.bt = bt0 + q; r.j0 = (int)rem * p.tn;
      r.u = layer0 + p.units_z + (i64)r.bt * nt_e + rem;
    }
    return r;
  }
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// X is the one array every layer re-reads (15 x 65 MB at C1, 7 % of the forward's HBM traffic) and it fits the 126 MB L2: its
// loads carry an evict_last policy so that the streaming iterates do not push it out between layers
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// non-blocking: has the stage below finished this unit's batch tile?  (the counters only grow: a "yes" stays true)
__device__ __forceinline__ bool pf_probe_ready(const PfParams& p, const PfUnit& un) {
  if (un.stage == 0) return true;
  const int below_tiles = (un.type == PF_Z) ? p.nt_e : p.nt_z;
  return ld_acquire_u32(p.flags + (i64)(un.stage - 1) * p.n_btiles + un.bt) >= (unsigned)(8 * below_tiles);
}
// wait until the stage below has finished this batch tile (every epilogue warp of every feature tile of it has published)
__device__ __forceinline__ void pf_wait_ready(const PfParams& p, const PfUnit& un, int lane, bool known_ready = false) {
  if (un.stage == 0) return;
  if (known_ready) {                   // observed earlier by this lane (acquire): only the proxy fence is left to do
    if (lane == 0) fence_proxy_async_all();
    __syncwarp();
    return;
  }
  if (lane == 0) {
    const int below_tiles = (un.type == PF_Z) ? p.nt_e : p.nt_z;     // a Z stage sits on T_0 / an E stage; an E stage sits on a Z stage
    const unsigned target = (unsigned)(8 * below_tiles);
    const unsigned* f = p.flags + (i64)(un.stage - 1) * p.n_btiles + un.bt;
    if (ld_acquire_u32(f) < target) {
      const long long t0 = clock64();
      while (ld_acquire_u32(f) < target) {
        if (clock64() - t0 > p.spin_limit) __trap();
      }
    }
    fence_proxy_async_all();           // the producing epilogues wrote through the generic proxy; TMA reads through the async proxy
  }
  __syncwarp();
}

// shared-memory split of the persistent kernel: operand stages vs staging ring (compile-time switches for A/B builds)
#ifndef PF_STAGES
#define PF_STAGES 3
#endif
#ifndef PF_RING_BYTES
#define PF_RING_BYTES RING_BYTES
#endif
constexpr int PF_SLOT_BYTES = 3 * 8 * TILE_B * 4;         // one staging-ring slot: three arrays x 8 rows (T_0, E/T/L) or one x 16 rows (Z)
constexpr int PF_DEPTH = PF_RING_BYTES / PF_SLOT_BYTES;   // 6
// operand stages of the kernel's (NPASS, KC, TN) variant: PF_STAGES of the 256-row tiles' 48 KB; the 32-row tiles of the small-batch
// variant are 20 KB per stage, so the same shared memory holds 7 of them -- with one batch tile per unit the mainloop is a latency
// chain (TMA -> split -> MMA -> free), and its rate is the number of chunks in flight
template <int NPASS, int KC, int TN>
__host__ __device__ constexpr int pf_stages() { return TN == TILE_N ? PF_STAGES : (PF_STAGES * SmemPlan<NPASS, KC, TILE_N>::STAGE_BYTES) / SmemPlan<NPASS, KC, TN>::STAGE_BYTES; }
template <int NPASS, int KC, int TN = TILE_N>
__host__ __device__ constexpr int pf_smem_total() { return pf_stages<NPASS, KC, TN>() * SmemPlan<NPASS, KC, TN>::STAGE_BYTES + PF_RING_BYTES + SmemPlan<NPASS, KC, TN>::BAR_BYTES + SmemPlan<NPASS, KC, TN>::ROWTAB + 1024; }
static_assert(PF_DEPTH % 2 == 0, "ring depth must be a multiple of the epilogue parts");

// the epilogue of ONE unit for one epilogue warp (rows [jw, jw + rpw) of the tile)
template <class Epi>
__device__ __forceinline__ void pf_run_epilogue(const Epi& epi, typename Epi::State& state, const PfParams& p, int n_feat, uint32_t t0,
                                                int jw, int rpw, i64 bt, int q, int col, i64 b, bool valid, uint8_t* ring,
                                                uint64_t* efull, uint64_t* eempty, RingPos& rp, int lane, float acc_scale) {
  constexpr int CHK = Epi::CHUNK;
  const int nch = rpw / CHK;
  typename Epi::Pre pre;
#pragma unroll 1
  for (int c = 0; c < nch; ++c, rp.advance(2, PF_DEPTH)) {
    const int row0 = jw + c * CHK;
    const int s = rp.s;
    mbar_wait(&efull[s], rp.ph);
    if (row0 < n_feat) {
      float v[CHK];
      tmem_ld(t0 + c * CHK, v);
#pragma unroll
      for (int i = 0; i < CHK; ++i) v[i] = __fmul_rn(v[i], acc_scale);
      const float* slot = reinterpret_cast<const float*>(ring + s * PF_SLOT_BYTES);
      if (row0 + CHK <= n_feat) epi.template apply<true>(state, slot, col, pre, row0, b, valid, v, n_feat, bt * (TILE_B / 32) + q);
      else epi.template apply<false>(state, slot, col, pre, row0, b, valid, v, n_feat, bt * (TILE_B / 32) + q);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&eempty[s]);
  }
}

template <int FAM, int PS, int NPASS, int KC, int TN>
__global__ void __launch_bounds__(roles_threads(8) + 32, 1)
umma_forward_persistent_kernel(const __grid_constant__ PfMaps maps, const __grid_constant__ PfParams p) {
  using Plan = SmemPlan<NPASS, KC, TN>;
  constexpr int STAGES = pf_stages<NPASS, KC, TN>();
  static_assert(TN == TILE_N || TN == 32 || TN == 64, "feature rows per tile");
  constexpr int EPI_WARPS = 8, EPI_PARTS = 2;
  constexpr int SPLIT_WARP0 = EPI_WARP0 + EPI_WARPS, EIN_WARP = SPLIT_WARP0 + SPLIT_WARPS, PUB_WARP = EIN_WARP + 1;
  constexpr int PUB_SLOTS = 4;
  static_assert(NPASS == 3 || NPASS == 1 || NPASS == 4, "tf32x3, single-pass tf32 or tf32 + two bf16 correction passes");
  static_assert(NPASS != 4 || KC == 16, "mixed mode: 16-row chunks");
  constexpr int MMA_K = 8;
  constexpr uint32_t B_LAYOUT = KC * 4 == 128 ? LAYOUT_SW128 : LAYOUT_SW64;
  constexpr uint32_t B_SBO = 8 * KC * 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem + STAGES * Plan::STAGE_BYTES;
  uint64_t* bars = (uint64_t*)(ring + PF_RING_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* ready = bars + 2 * STAGES;
  uint64_t* tfull = bars + 3 * STAGES;
  uint64_t* tempty = bars + 3 * STAGES + 2;
  uint64_t* efull = bars + 3 * STAGES + 4;
  uint64_t* eempty = bars + 3 * STAGES + 4 + MAX_RING_DEPTH;
  uint64_t* pfull = bars + 3 * STAGES + 4 + 2 * MAX_RING_DEPTH;         // [PUB_SLOTS] a unit's epilogue warps are done storing
  uint64_t* pempty = pfull + PUB_SLOTS;                                  // [PUB_SLOTS] ... and the publisher has published it
  uint64_t* rfull = pempty + PUB_SLOTS;                                  // [PUB_SLOTS] the scout saw a unit's inputs complete
  uint64_t* rempty = rfull + PUB_SLOTS;                                  // [PUB_SLOTS] ... and both TMA producers took note
  uint32_t* tmem_slot = (uint32_t*)(rempty + PUB_SLOTS);
  float* rowtab = (float*)(ring + PF_RING_BYTES + Plan::BAR_BYTES);
  static_assert((3 * STAGES + 4 + 2 * MAX_RING_DEPTH + 4 * PUB_SLOTS) * 8 + 8 <= Plan::BAR_BYTES, "barrier block");
  static_assert(2 * TN <= 512, "two accumulators in TMEM");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PfWalk walk; walk.init(p);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.A_big); prefetch_tmap(&maps.W_big); prefetch_tmap(&maps.actZ); prefetch_tmap(&maps.actV);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], SPLIT_WARPS); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], EPI_WARPS); }
    for (int s = 0; s < PF_DEPTH; ++s) { mbar_init(&efull[s], 1); mbar_init(&eempty[s], 4); }
    for (int s = 0; s < PUB_SLOTS; ++s) { mbar_init(&pfull[s], EPI_WARPS); mbar_init(&pempty[s], 1); mbar_init(&rfull[s], 1); mbar_init(&rempty[s], 2); }
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
  grid_dep_launch_dependents();
  if (warp != 1 && !(warp >= SPLIT_WARP0 && warp < EIN_WARP)) grid_dep_wait();
  if (p.nstreams == 2 && (blockIdx.x & 1) && p.skew > 0 && (warp == 0 || warp == EIN_WARP)) {
    // the second CTA set starts late (its producers hold everything else back)
    const long long t0 = clock64();
    while (clock64() - t0 < p.skew) __nanosleep(256);
  }

  if (warp == 0) {
    // ===== TMA producer (MMA operands) =====
    int s = 0; uint32_t ph = 0; int ui = 0;
    int rs = 0; uint32_t rph = 0;
    for (uint32_t v = walk.first; v < walk.total; v += walk.stride) {
      const PfUnit un = walk.decode(p, v);
      // the scout has seen this unit's inputs complete; order this warp's async-proxy loads after that
      mbar_wait(&rfull[rs], rph);
      if (lane == 0) { fence_proxy_async_all(); mbar_arrive(&rempty[rs]); }
      __syncwarp();
      if (++rs == PUB_SLOTS) { rs = 0; rph ^= 1; }
      if (lane == 0) { PF_TR(ui, 6, clock64()); PF_TR(ui, 7, ((long long)un.stage << 32) | un.bt); PF_TR(ui, 0, un.type); }
      ++ui;
      const int b0 = (int)(un.bt * TILE_B);
      const int kcs = un.type == PF_Z ? p.kc_z : p.kc_e;
      const int zslab = p.last_only ? (un.k & 1) : un.k;
      const int vslab = p.v_per_layer ? un.k : 0;
      const int widx = p.layer[un.k].widx;
      for (int kc = 0; kc < kcs; ++kc) {
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
          uint8_t* st = smem + s * Plan::STAGE_BYTES;
          mbar_expect_tx(&full[s], Plan::TX_BYTES);
          uint8_t* b_dst = st + Plan::NOPS * Plan::A_BYTES;
          if (un.type == PF_Z) {
#pragma unroll
            for (int g = 0; g < TILE_B / 32; ++g) tma_load_3d(st + g * (KC * 128), &maps.actV, &full[s], b0 + g * 32, kc * KC, vslab);
            tma_load_3d(b_dst, &maps.W_big, &full[s], kc * KC, un.j0, widx);
            if (NPASS >= 3) tma_load_3d(b_dst + Plan::B_BYTES, &maps.W_small, &full[s], kc * KC, un.j0, widx);
          } else {
            if (un.type == PF_T0) {
#pragma unroll
              for (int g = 0; g < TILE_B / 32; ++g) tma_load_2d(st + g * (KC * 128), &maps.actZ0, &full[s], b0 + g * 32, kc * KC);
            } else {
#pragma unroll
              for (int g = 0; g < TILE_B / 32; ++g) tma_load_3d(st + g * (KC * 128), &maps.actZ, &full[s], b0 + g * 32, kc * KC, zslab);
            }
            tma_load_2d(b_dst, &maps.A_big, &full[s], kc * KC, un.j0);
            if (NPASS >= 3) tma_load_2d(b_dst + Plan::B_BYTES, &maps.A_small, &full[s], kc * KC, un.j0);
          }
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc(TILE_B, TN, 1, 0, 2u);
    constexpr uint32_t a_hi = desc_hi(A_ATOM_BYTES, LAYOUT_SW128_BASE32B);
    constexpr uint32_t A_KSTEP = MMA_K * 128;
    constexpr uint32_t b_hi = desc_hi(B_SBO, B_LAYOUT);
    const uint32_t st0 = smem_u32(smem);
    const uint32_t a_lo0 = desc_lo(st0, KC * 128);
    const uint32_t b_lo0 = desc_lo(st0 + Plan::NOPS * Plan::A_BYTES, 16);
    int s = 0; uint32_t ph = 0; int ui = 0;
    int acc = 0; uint32_t aph = 0;
    for (uint32_t v = walk.first; v < walk.total; v += walk.stride) {
      const PfUnit un = walk.decode(p, v);
      const int kcs = un.type == PF_Z ? p.kc_z : p.kc_e;
      mbar_wait(&tempty[acc], aph ^ 1);
      tc_fence_after();
      if (lane == 0) PF_TR(ui, 1, clock64());
      const uint32_t d_tmem = tmem_base + acc * TN;
      for (int kc = 0; kc < kcs; ++kc) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (kc == 0 && lane == 0) PF_TR(ui, 2, clock64());
        const uint32_t a_lo = a_lo0 + s * (Plan::STAGE_BYTES >> 4), b_lo = b_lo0 + s * (Plan::STAGE_BYTES >> 4);
        if (NPASS == 3) {
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < KC / MMA_K; ++ks) {
              const uint64_t da_big = desc_at(a_hi, a_lo + ks * (A_KSTEP >> 4));
              const uint64_t db_big = desc_at(b_hi, b_lo + ks * (32 >> 4));
              const uint64_t db_small = desc_at(b_hi, b_lo + (Plan::B_BYTES >> 4) + ks * (32 >> 4));
              umma_tf32(d_tmem, da_big, db_small, idesc, (kc == 0 && ks == 0) ? 0u : 1u);
              umma_tf32(d_tmem, da_big, db_big, idesc, 1u);
            }
          }
          __syncwarp();
        }
        if (NPASS == 4) {
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < KC / MMA_K; ++ks)
              umma_tf32(d_tmem, desc_at(a_hi, a_lo + ks * (A_KSTEP >> 4)), desc_at(b_hi, b_lo + ks * (32 >> 4)), idesc, (kc == 0 && ks == 0) ? 0u : 1u);
          }
          __syncwarp();
        }
        mbar_wait(&ready[s], ph);
        tc_fence_after();
        if (elect_one()) {
          if (NPASS == 4) {         // the two bf16 correction products (umma_gemm.cuh, split_tile_mix)
            constexpr uint32_t a16_hi = desc_hi(1024, LAYOUT_SW128);
            constexpr uint32_t idesc16 = make_idesc(TILE_B, TN, 1, 0, 1u);
            const uint32_t a16 = a_lo + (Plan::A_BYTES >> 4), b16 = b_lo + (Plan::B_BYTES >> 4);
            umma_f16(d_tmem, desc_at(a16_hi, a16), desc_at(b_hi, b16 + (32 >> 4)), idesc16, 1u);
            umma_f16(d_tmem, desc_at(a16_hi, a16 + (Plan::A_BYTES >> 5)), desc_at(b_hi, b16), idesc16, 1u);
          }
#pragma unroll
          for (int ks = 0; ks < (NPASS == 4 ? 0 : KC / MMA_K); ++ks) {
            const uint64_t db_big = desc_at(b_hi, b_lo + ks * (32 >> 4));
            if (NPASS == 3) {
              umma_tf32(d_tmem, desc_at(a_hi, a_lo + (Plan::A_BYTES >> 4) + ks * (A_KSTEP >> 4)), db_big, idesc, 1u);
            } else {
              umma_tf32(d_tmem, desc_at(a_hi, a_lo + ks * (A_KSTEP >> 4)), db_big, idesc, (kc == 0 && ks == 0) ? 0u : 1u);
            }
          }
          umma_commit(&empty[s]);
          if (kc == kcs - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      if (lane == 0) PF_TR(ui, 3, clock64());
      ++ui;
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  } else if (warp == EIN_WARP) {
    // ===== TMA producer (epilogue inputs) =====
    RingPos rp; rp.init(0, PF_DEPTH);
    const uint64_t pol_x = l2_policy_evict_last();
    int rs = 0; uint32_t rph = 0;
    for (uint32_t v = walk.first; v < walk.total; v += walk.stride) {
      const PfUnit un = walk.decode(p, v);
      mbar_wait(&rfull[rs], rph);
      if (lane == 0) { fence_proxy_async_all(); mbar_arrive(&rempty[rs]); }
      __syncwarp();
      if (++rs == PUB_SLOTS) { rs = 0; rph ^= 1; }
      // The staging ring holds 72 KB -- a fraction of a tile's epilogue inputs (384 KB for an E/T/L unit) -- so with a loaded DRAM
      // latency of a few microseconds the ring alone keeps too few bytes in flight.  p.prefetch = D > 0: every chunk's boxes are
      // also requested into L2 D chunks ahead (cp.async.bulk.prefetch.tensor needs no shared memory), across the unit boundary
      // when the next unit's inputs are already complete; the ring loads then hit L2.  (Prefetching the WHOLE next unit was
      // measured 8-12 % slower: 57 MB per round do not survive in a 126 MB L2 that also streams 14 GB per forward.)
      PfUnit nx = un; bool nx_ok = false;
      if (p.prefetch > 0 && v + walk.stride < walk.total) {
        nx = walk.decode(p, v + walk.stride);
        nx_ok = true;
        if (nx.stage > 0) {
          const int below = (nx.type == PF_Z) ? p.nt_e : p.nt_z;
          nx_ok = ld_acquire_u32(p.flags + (i64)(nx.stage - 1) * p.n_btiles + nx.bt) >= (unsigned)(8 * below);
        }
      }
      auto prefetch_chunk = [&](const PfUnit& w, int c) {      // both parts of chunk c of unit w (called by one elected lane)
        const int wb0 = (int)(w.bt * TILE_B);
        const int wchk = w.type == PF_Z ? 16 : 8;
        const int wfeat = w.type == PF_Z ? p.d : p.m;
        const int wslab = p.last_only ? ((w.k - 1) & 1) : (w.k - 1);
#pragma unroll
        for (int h = 0; h < EPI_PARTS; ++h) {
          const int row0 = w.j0 + h * (TN / EPI_PARTS) + c * wchk;
          if (row0 >= wfeat) continue;
          if (w.type == PF_T0) {
            tma_prefetch_2d(&maps.sE0, wb0, row0); tma_prefetch_2d(&maps.sX, wb0, row0); tma_prefetch_2d(&maps.sL0, wb0, row0);
          } else if (w.type == PF_Z) {
            if (w.k == 0) tma_prefetch_2d(&maps.sZ0, wb0, row0); else tma_prefetch_3d(&maps.sZ, wb0, row0, wslab);
          } else {
            tma_prefetch_2d(&maps.sX, wb0, row0);
            if (w.k == 0) tma_prefetch_2d(&maps.sL0, wb0, row0); else tma_prefetch_3d(&maps.sL, wb0, row0, wslab);
            if (FAM == DLADMM_FAMILY_B) { if (w.k == 0) tma_prefetch_2d(&maps.sE0, wb0, row0); else tma_prefetch_3d(&maps.sE, wb0, row0, wslab); }
          }
        }
      };
      const int b0 = (int)(un.bt * TILE_B);
      const int chk = un.type == PF_Z ? 16 : 8;
      const int n_feat = un.type == PF_Z ? p.d : p.m;
      const int rpw = TN / EPI_PARTS;
      const int nch = rpw / chk;
      const int pslab = p.last_only ? ((un.k - 1) & 1) : (un.k - 1);      // slab of the previous layer's iterate (k >= 1)
      const bool has_ep = FAM == DLADMM_FAMILY_B;
      for (int c = 0; c < nch; ++c) {
        if (p.prefetch > 0 && elect_one()) {
          const int cp = c + p.prefetch;
          if (cp < nch) prefetch_chunk(un, cp);
          else if (nx_ok && cp - nch < (TN / EPI_PARTS) / (nx.type == PF_Z ? 16 : 8)) prefetch_chunk(nx, cp - nch);
        }
        __syncwarp();
        for (int h = 0; h < EPI_PARTS; ++h, rp.advance(1, PF_DEPTH)) {
          const int s = rp.s;
          mbar_wait(&eempty[s], rp.ph ^ 1);
          const int row0 = un.j0 + h * rpw + c * chk;
          if (elect_one()) {
            if (row0 >= n_feat) {
              mbar_arrive(&efull[s]);
            } else {
              uint8_t* dst = ring + s * PF_SLOT_BYTES;
              constexpr int SUB8 = 8 * TILE_B * 4, SUB16 = 16 * TILE_B * 4;
              if (un.type == PF_T0) {
                mbar_expect_tx(&efull[s], 3 * SUB8);
                tma_load_2d(dst, &maps.sE0, &efull[s], b0, row0);
                if (p.x_resident) tma_load_2d_hint(dst + SUB8, &maps.sX, &efull[s], b0, row0, pol_x);
                else tma_load_2d(dst + SUB8, &maps.sX, &efull[s], b0, row0);
                tma_load_2d(dst + 2 * SUB8, &maps.sL0, &efull[s], b0, row0);
              } else if (un.type == PF_Z) {
                mbar_expect_tx(&efull[s], SUB16);
                if (un.k == 0) tma_load_2d(dst, &maps.sZ0, &efull[s], b0, row0);
                else tma_load_3d(dst, &maps.sZ, &efull[s], b0, row0, pslab);
              } else {
                mbar_expect_tx(&efull[s], (has_ep ? 3 : 2) * SUB8);
                if (p.x_resident) tma_load_2d_hint(dst, &maps.sX, &efull[s], b0, row0, pol_x);
                else tma_load_2d(dst, &maps.sX, &efull[s], b0, row0);
                if (un.k == 0) tma_load_2d(dst + SUB8, &maps.sL0, &efull[s], b0, row0);
                else tma_load_3d(dst + SUB8, &maps.sL, &efull[s], b0, row0, pslab);
                if (has_ep) {
                  if (un.k == 0) tma_load_2d(dst + 2 * SUB8, &maps.sE0, &efull[s], b0, row0);
                  else tma_load_3d(dst + 2 * SUB8, &maps.sE, &efull[s], b0, row0, pslab);
                }
              }
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == PUB_WARP) {
    // ===== publisher: "this tile of this stage is in global memory" =====
    // The epilogue warps only ARRIVE on a shared-memory barrier when their stores are issued (release at CTA scope); this warp
    // acquires that, fences at GPU scope -- cumulative: it covers the epilogue warps' stores -- and bumps the readiness counter.
    // A __threadfence in every epilogue warp instead cost them 18 % of their time (ncu: CCTL.IVALL / ERRBAR / REDG under
    // stall_membar, profiles/r02_ncu_persistent_summary.md): the fence waits for ~500 stores per warp to drain.
    //
    // Lane 1 of the same warp is the SCOUT: it walks the same list up to PUB_SLOTS units ahead, spins (acquire, GPU scope) on
    // the counter each unit depends on and tells the two TMA producers through a shared-memory barrier -- an acquire load on
    // the producers' own path cost an L2 round trip (~1 us) of operand-pipeline stall at every unit boundary.
    // (two lanes of one warp in independent loops: independent thread scheduling interleaves them)
    if (lane == 0) {
      int ps = 0; uint32_t pph = 0;
      for (uint32_t v = walk.first; v < walk.total; v += walk.stride) {
        const PfUnit un = walk.decode(p, v);
        mbar_wait(&pfull[ps], pph);
        __threadfence();
        red_release_add_u32(p.flags + (i64)un.stage * p.n_btiles + un.bt, (unsigned)EPI_WARPS);
        mbar_arrive(&pempty[ps]);
        if (++ps == PUB_SLOTS) { ps = 0; pph ^= 1; }
      }
    } else if (lane == 1) {
      int rs = 0; uint32_t rph = 0;
      for (uint32_t v = walk.first; v < walk.total; v += walk.stride) {
        const PfUnit un = walk.decode(p, v);
        if (un.stage > 0) {
          const int below_tiles = (un.type == PF_Z) ? p.nt_e : p.nt_z;
          const unsigned target = (unsigned)(8 * below_tiles);
          const unsigned* f = p.flags + (i64)(un.stage - 1) * p.n_btiles + un.bt;
          if (ld_acquire_u32(f) < target) {
            const long long t0 = clock64();
            while (ld_acquire_u32(f) < target) {
              if (clock64() - t0 > p.spin_limit) __trap();
              // the scout runs up to PUB_SLOTS units ahead of the producers, so a coarse poll costs no latency; an unthrottled
              // acquire spin executed 6.9 M ld.acquire + CCTL.IVALL per forward (ncu), issue slots and power of a power-capped kernel
              if (p.scout_sleep_ns > 0) __nanosleep(p.scout_sleep_ns);
            }
          }
        }
        mbar_wait(&rempty[rs], rph ^ 1);
        mbar_arrive(&rfull[rs]);                     // release (CTA scope): the producers' acquire of it orders their loads after the counter
        if (++rs == PUB_SLOTS) { rs = 0; rph ^= 1; }
      }
    }
  } else if (warp >= SPLIT_WARP0) {
    // ===== operand splitters: small = x - trunc_tf32(x) (3 passes) / round to the nearest tf32 in place (1 pass) =====
    const int tid = threadIdx.x - SPLIT_WARP0 * 32;
    int s = 0; uint32_t ph = 0;
    for (uint32_t v = walk.first; v < walk.total; v += walk.stride) {
      const PfUnit un = walk.decode(p, v);
      const int kcs = un.type == PF_Z ? p.kc_z : p.kc_e;
      for (int kc = 0; kc < kcs; ++kc) {
        mbar_wait(&full[s], ph);
        uint8_t* st = smem + s * Plan::STAGE_BYTES;
        if (NPASS == 3) split_tile<Plan::A_BYTES, SPLIT_WARPS * 32>(st, st + Plan::A_BYTES, tid);
        else if (NPASS == 4) split_tile_mix<SPLIT_WARPS * 32>(st, st + Plan::A_BYTES, st + Plan::A_BYTES + Plan::A_BYTES / 2, tid);
        else round_tile<Plan::A_BYTES, SPLIT_WARPS * 32>(st, tid);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3;
    const int half = (warp - EPI_WARP0) >> 2;
    const int col = q * 32 + lane;
    const int ewarp = warp - EPI_WARP0;
    const int etid = threadIdx.x - EPI_WARP0 * 32;
    int acc = 0; uint32_t aph = 0; int ui = 0;
    int ps = 0; uint32_t pph = 0;
    int tab_key = -1;
    RingPos rp; rp.init(half, PF_DEPTH);
    const int rpw = TN / EPI_PARTS;
    for (uint32_t v = walk.first; v < walk.total; v += walk.stride) {
      const PfUnit un = walk.decode(p, v);
      const i64 b = un.bt * TILE_B + col;
      const bool valid = b < p.B;
      const int jw = un.j0 + half * rpw;
      const int k = un.k;
      const PfLayer& ly = p.layer[k];
      const int slab = p.last_only ? (k & 1) : k;
      const uint32_t t0 = tmem_base + acc * TN + half * rpw + ((uint32_t)(q * 32) << 16);
      float* objp = p.obj_part ? p.obj_part + un.u * EPI_WARPS : nullptr;
      const int key = (un.stage << 16) | (un.j0 / TN);    // PM_ROWS: which (stage, feature tile) the parameter table holds
      if (un.type == PF_T0) {
        UEpiT0<PS> epi;
        epi.E0 = p.E0; epi.X = p.X; epi.L0 = p.L0; epi.T0 = p.T;
        epi.b1 = p.K > 0 ? to_bp(p.layer[0].b1) : BP{nullptr, nullptr, 0, 0};
        epi.V = p.K > 0 ? p.V : nullptr; epi.B = p.B; epi.in_mask = 7u; epi.Vh = nullptr; epi.ldh = 0;
        typename UEpiT0<PS>::State state;
        epi.begin(state);
        if constexpr (PS == PM_ROWS) {
          if (tab_key != key) {                           // (uniform over the epilogue warps)
            asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");    // the previous unit's readers are done with the table
            BP qv[1]; epi.row_params(qv);
            fill_rowtab<1>(qv, rowtab, un.j0, p.m, etid, EPI_WARPS * 32, TN);
            asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
            tab_key = key;
          }
          epi.bind_rows(state, rowtab - un.j0, TILE_N);
        }
        mbar_wait(&tfull[acc], aph);
        tc_fence_after();
        if (ewarp == 0 && lane == 0) PF_TR(ui, 4, clock64());
        pf_run_epilogue(epi, state, p, p.m, t0, jw, rpw, un.bt, q, col, b, valid, ring, efull, eempty, rp, lane, p.acc_scale_e);
      } else if (un.type == PF_Z) {
        UEpiZ<PS> epi;
        epi.Zp = nullptr; epi.Zk = p.Z + p.zs * slab; epi.maskZ = p.maskZ ? p.maskZ + p.zs * slab : nullptr;
        epi.th1 = to_bp(ly.th1); epi.ss1 = to_bp(ly.ss1); epi.B = p.B; epi.in_mask = 1u;
        epi.obj_part = objp; epi.Zlabel = nullptr; epi.sq_part = nullptr; epi.Zh = nullptr; epi.ldh = 0;
        typename UEpiZ<PS>::State state;
        epi.begin(state);
        if constexpr (PS == PM_ROWS) {
          if (tab_key != key) {                           // (uniform over the epilogue warps)
            asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");    // the previous unit's readers are done with the table
            BP qv[1]; epi.row_params(qv);
            fill_rowtab<1>(qv, rowtab, un.j0, p.d, etid, EPI_WARPS * 32, TN);
            asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
            tab_key = key;
          }
          epi.bind_rows(state, rowtab - un.j0, TILE_N);
        }
        mbar_wait(&tfull[acc], aph);
        tc_fence_after();
        if (ewarp == 0 && lane == 0) PF_TR(ui, 4, clock64());
        pf_run_epilogue(epi, state, p, p.d, t0, jw, rpw, un.bt, q, col, b, valid, ring, efull, eempty, rp, lane, p.acc_scale_z);
        epi.end(state, ewarp, lane);
      } else {
        UEpiELT<FAM, PS, false> epi;
        epi.obj_part = objp; epi.obj_kind = p.obj_kind; epi.Elabel = nullptr; epi.Xclean = nullptr; epi.met_part = nullptr; epi.met_stride = 0;
        epi.X = p.X; epi.Ep = nullptr; epi.Lp = nullptr;
        epi.Ek = p.E + p.ms * slab; epi.Lk = p.L + p.ms * slab;
        epi.Tn = p.T + p.ms * (p.last_only ? ((k + 1) & 1) : (k + 1));
        epi.maskE = p.maskE ? p.maskE + p.ms * slab : nullptr;
        epi.b2 = to_bp(ly.b2); epi.ss2 = to_bp(ly.ss2); epi.ss2_2 = to_bp(ly.ss2_2); epi.th2 = to_bp(ly.th2); epi.bL = to_bp(ly.bL);
        epi.has_next = k + 1 < p.K;
        epi.b1n = to_bp(p.layer[k + 1 < p.K ? k + 1 : k].b1);
        epi.V = p.V + (p.v_per_layer ? p.ms * (k + 1 < p.K ? k + 1 : k) : 0);
        epi.Vh = nullptr; epi.ldh = 0; epi.B = p.B; epi.in_mask = FAM == DLADMM_FAMILY_B ? 7u : 3u;
        typename UEpiELT<FAM, PS, false>::State state;
        epi.begin(state);
        if constexpr (PS == PM_ROWS) {
          if (tab_key != key) {                           // (uniform over the epilogue warps)
            asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");    // the previous unit's readers are done with the table
            BP qv[6]; epi.row_params(qv);
            fill_rowtab<6>(qv, rowtab, un.j0, p.m, etid, EPI_WARPS * 32, TN);
            asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
            tab_key = key;
          }
          epi.bind_rows(state, rowtab - un.j0, TILE_N);
        }
        mbar_wait(&tfull[acc], aph);
        tc_fence_after();
        if (ewarp == 0 && lane == 0) PF_TR(ui, 4, clock64());
        pf_run_epilogue(epi, state, p, p.m, t0, jw, rpw, un.bt, q, col, b, valid, ring, efull, eempty, rp, lane, p.acc_scale_e);
        epi.end(state, ewarp, lane);
      }
      if (ewarp == 0 && lane == 0) PF_TR(ui, 5, clock64());
      ++ui;
      // the accumulator is free; this warp's rows of the tile are stored: hand the unit to the publisher warp
      // (lanes' stores -> __syncwarp -> lane 0's arrive, release at CTA scope)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tempty[acc]);
        mbar_wait(&pempty[ps], pph ^ 1);            // (the publisher is never more than a unit behind: free in practice)
        mbar_arrive(&pfull[ps]);
      }
      __syncwarp();
      if (++ps == PUB_SLOTS) { ps = 0; pph ^= 1; }
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace umma
}  // namespace dladmm
