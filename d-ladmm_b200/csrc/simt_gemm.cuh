// CUDA-core (FFMA) matrix products with fused epilogues: the DLADMM_PREC_FP32 arithmetic.
//
//   simt_gemm_kernel   : C(M x N) = Aw(M x Kd, K-major, zero padded pitch) * Bop(Kd x N, batch contiguous)
//                        -> epilogue functor on 4-column quads; N = batch (large), M = features
//   simt_gemm_nt_kernel: gW(M x N) += alpha * P(M x B) * Q(N x B)^T, reduction over the batch (split over CTAs)
//
// Tile 64 (features) x 256 (batch) x 16, 256 threads, 8x8 accumulators per thread; a warp owns 8 feature
// rows across the whole 256-column tile so that row-reduced parameter gradients are one warp shuffle tree.
#pragma once
#include "common.cuh"
#include "epilogues.cuh"

namespace dladmm {

constexpr int SG_BM = 64, SG_BN = 256, SG_BK = 16, SG_THREADS = 256;

// ---- B-operand loaders (activation side; row = reduction index, columns = batch) -----------------
struct BPlain {
  const float* p; i64 B;
  __device__ __forceinline__ float4 load(int k, i64 col, int nvalid, bool vec) const {
    Quad q = load4(p, (i64)k * B + col, nvalid, vec);
    return make_float4(q.v[0], q.v[1], q.v[2], q.v[3]);
  }
};
// V_k = L_{k-1} + beta1 * T_k, built while loading (main_syn_l1l1_scalar.py:93,111)
struct BVar {
  const float* L; const float* T; BP b1; i64 B;
  __device__ __forceinline__ float4 load(int k, i64 col, int nvalid, bool vec) const {
    i64 off = (i64)k * B + col;
    Quad l = load4(L, off, nvalid, vec), t = load4(T, off, nvalid, vec);
    float b[4]; bp_at4(b1, k, col, b);
    return make_float4(fadd(l.v[0], fmul(b[0], t.v[0])), fadd(l.v[1], fmul(b[1], t.v[1])),
                       fadd(l.v[2], fmul(b[2], t.v[2])), fadd(l.v[3], fmul(b[3], t.v[3])));
  }
};

template <class BLoad, class Epi>
__global__ void __launch_bounds__(SG_THREADS)
simt_gemm_kernel(int M, i64 N, int Kd, const float* __restrict__ Aw, int lda, BLoad bl, Epi epi,
                 float* __restrict__ part, int ncolTiles, int prow) {
  __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Bs[SG_BK][SG_BN];
  const int t = threadIdx.x, w = t >> 5, l = t & 31;
  const int m0 = blockIdx.y * SG_BM;
  const i64 n0 = (i64)blockIdx.x * SG_BN;
  const bool vec_ok = (N & 3) == 0;

  const int arow = t >> 2, ak = (t & 3) * 4;
  const int bk = t >> 6;
  const int bc = (t & 63) * 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra, rb[4];
  auto gload = [&](int k0) {
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + arow < M) ra = *reinterpret_cast<const float4*>(Aw + (i64)(m0 + arow) * lda + k0 + ak);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int k = k0 + r * 4 + bk;
      i64 col = n0 + bc;
      rb[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < Kd && col < N) {
        i64 rem = N - col;
        int nv = rem >= 4 ? 4 : (int)rem;
        rb[r] = bl.load(k, col, nv, vec_ok);
      }
    }
  };

  gload(0);
  for (int k0 = 0; k0 < Kd; k0 += SG_BK) {
    __syncthreads();
    As[ak + 0][arow] = ra.x; As[ak + 1][arow] = ra.y; As[ak + 2][arow] = ra.z; As[ak + 3][arow] = ra.w;
#pragma unroll
    for (int r = 0; r < 4; ++r) *reinterpret_cast<float4*>(&Bs[r * 4 + bk][bc]) = rb[r];
    __syncthreads();
    if (k0 + SG_BK < Kd) gload(k0 + SG_BK);
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][w * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][w * 8 + 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][l * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][128 + l * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }

  constexpr int NR = Epi::NRED > 0 ? Epi::NRED : 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + w * 8 + i;
    float red[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) red[r] = 0.f;
    if (row < M) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        i64 col = n0 + h * 128 + l * 4;
        if (col < N) {
          i64 rem = N - col;
          int nv = rem >= 4 ? 4 : (int)rem;
          float q[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
          epi(row, col, q, nv, vec_ok && nv == 4, red);
        }
      }
    }
    if (Epi::NRED > 0) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        float s = warp_sum(red[r]);
        if (l == 0 && row < M) part[((i64)(Epi::SLOT0 + r) * ncolTiles + blockIdx.x) * prow + row] = s;
      }
    }
  }
}

// ---- gW += alpha * P * Q^T, reduction over batch columns ------------------------------------------
constexpr int NT_BM = 128, NT_BN = 64, NT_BK = 32, NT_THREADS = 256;

template <class QLoad>
__global__ void __launch_bounds__(NT_THREADS)
simt_gemm_nt_kernel(int M, int N, i64 B, i64 chunk, const float* __restrict__ P, QLoad ql,
                    const float* __restrict__ s1ptr, float sign, float* __restrict__ C, int ldc) {
  __shared__ __align__(16) float Ps[NT_BK][NT_BM + 4];
  __shared__ __align__(16) float Qs[NT_BK][NT_BN + 4];
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * NT_BM, n0 = blockIdx.y * NT_BN;
  const i64 b_begin = (i64)blockIdx.z * chunk;
  i64 b_end = b_begin + chunk; if (b_end > B) b_end = B;
  const bool vec_ok = (B & 3) == 0;
  const int ty = t >> 4, tx = t & 15;      // 16 x 16 threads: 8 rows x 4 cols each

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loads: P tile 128 rows x 32 cols = 1024 quads -> 4 per thread; Q tile 64 x 32 = 512 quads -> 2 per thread
  const int lq = t & 7, lr = t >> 3;        // quad index within row (8 quads = 32 cols), row 0..31
  for (i64 b0 = b_begin; b0 < b_end; b0 += NT_BK) {
    float4 rp[4], rq[2];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int row = m0 + lr + 32 * r;
      i64 col = b0 + lq * 4;
      rp[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < M && col < b_end) {
        i64 rem = b_end - col; int nv = rem >= 4 ? 4 : (int)rem;
        Quad q = load4(P, (i64)row * B + col, nv, vec_ok && nv == 4);
        rp[r] = make_float4(q.v[0], q.v[1], q.v[2], q.v[3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      int row = n0 + lr + 32 * r;
      i64 col = b0 + lq * 4;
      rq[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < N && col < b_end) {
        i64 rem = b_end - col; int nv = rem >= 4 ? 4 : (int)rem;
        rq[r] = ql.load(row, col, nv, vec_ok && nv == 4);
        if (nv < 4) { if (nv < 2) rq[r].y = 0.f; if (nv < 3) rq[r].z = 0.f; rq[r].w = 0.f; }
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int row = lr + 32 * r;
      Ps[lq * 4 + 0][row] = rp[r].x; Ps[lq * 4 + 1][row] = rp[r].y;
      Ps[lq * 4 + 2][row] = rp[r].z; Ps[lq * 4 + 3][row] = rp[r].w;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      int row = lr + 32 * r;
      Qs[lq * 4 + 0][row] = rq[r].x; Qs[lq * 4 + 1][row] = rq[r].y;
      Qs[lq * 4 + 2][row] = rq[r].z; Qs[lq * 4 + 3][row] = rq[r].w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < NT_BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&Ps[kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&Ps[kk][ty * 8 + 4]);
      float4 b0v = *reinterpret_cast<const float4*>(&Qs[kk][tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[4] = {b0v.x, b0v.y, b0v.z, b0v.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  const float alpha = sign * (s1ptr ? __ldg(s1ptr) : 1.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int row = m0 + ty * 8 + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = n0 + tx * 4 + j;
      if (col < N) atomicAdd(C + (i64)row * ldc + col, alpha * acc[i][j]);
    }
  }
}

// ---- standalone elementwise backward of the top layer ----------------------------------------------
template <int FAM>
__global__ void __launch_bounds__(256, 3) m1_kernel(int M, i64 N, M1Args a, float* __restrict__ part, int ncolTiles, int prow) {
  // one warp per (row, 256-column tile): same partial layout as simt_gemm_kernel.  A block is one row x 8 consecutive tiles (8 KB
  // contiguous per array).  (Measured 0.21-0.22 ms for the 475 MB of the C1 top layer with this and with the former 8 rows x 1 KB
  // mapping alike: the kernel is bound by its dependent load -> compute -> store chain per thread, not by DRAM locality.)
  // Three blocks per SM (85 registers instead of the family-B instance's 88, which allowed two): the kernel's rate is its bytes in flight.
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int row = blockIdx.y;
  const int tile = blockIdx.x * 8 + w;
  const i64 n0 = (i64)tile * SG_BN;
  const bool vec_ok = (N & 3) == 0;
  float red[4] = {0.f, 0.f, 0.f, 0.f};
  if (tile >= ncolTiles) return;                     // (whole warp)
  if (row < M) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      i64 col = n0 + h * 128 + l * 4;
      if (col < N) {
        i64 rem = N - col; int nv = rem >= 4 ? 4 : (int)rem;
        float dL[4] = {0, 0, 0, 0}, dE[4] = {0, 0, 0, 0}, dT[4] = {0, 0, 0, 0};
        m1_quad<FAM>(a, row, col, dL, dE, dT, nv, vec_ok && nv == 4, red);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float s = warp_sum(red[r]);
    if (l == 0 && row < M) part[((i64)r * ncolTiles + tile) * prow + row] = s;
  }
}

// ---- second stage of the parameter-gradient reductions (deterministic) -----------------------------
// part_off: where the job's layer keeps its partial block inside `part` (the tcgen05 backward defers the reductions of a few layers
// into one launch: 20 launches of ~160 blocks were latency-bound at 23 us each, 0.46 ms of a `full` K = 20 training step)
struct ReduceJob { int slot; int rows; int scalar; float* grad; i64 part_off; };
constexpr int MAX_REDUCE_JOBS = 32;
struct ReduceJobs { int n; ReduceJob j[MAX_REDUCE_JOBS]; };

static __global__ void __launch_bounds__(256) reduce_partials_kernel(ReduceJobs jobs, const float* __restrict__ part,
                                                              int ncolTiles, int prow) {
  const ReduceJob jb = jobs.j[blockIdx.y];
  const float* base = part + jb.part_off + (i64)jb.slot * ncolTiles * prow;
  __shared__ float sm[256];
  if (jb.scalar) {
    if (blockIdx.x != 0) return;
    // thread t owns rows t, t + 256, ...: coalesced over the block, no division, four independent partial sums in flight (the former
    // flat loop with a 64-bit div/mod per element took 0.14 ms for the 64 K partials of the top layer: one dependent load at a time)
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int r = threadIdx.x; r < jb.rows; r += 256) {
      const float* q = base + r;
      int ct = 0;
      for (; ct + 4 <= ncolTiles; ct += 4) {
        s0 += q[(i64)ct * prow]; s1 += q[(i64)(ct + 1) * prow]; s2 += q[(i64)(ct + 2) * prow]; s3 += q[(i64)(ct + 3) * prow];
      }
      for (; ct < ncolTiles; ++ct) s0 += q[(i64)ct * prow];
    }
    const float s = (s0 + s1) + (s2 + s3);
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(jb.grad, sm[0]);   // one add per (parameter, launch); atomic because a C-ABI caller may alias grads between slots
  } else {
    // 8 consecutive rows per block: thread t reads row (t & 7) of column tiles (t >> 3), (t >> 3) + 32, ... -- a warp's load covers
    // four tiles x 8 rows = four full 32-byte sectors (one warp per row with the lanes over the tiles used 4 bytes of every sector it
    // touched: 0.66 ms per `full` K = 20 step); fixed summation order
    const int r = threadIdx.x & 7, c = threadIdx.x >> 3;
    const int row = blockIdx.x * 8 + r;
    float s0 = 0.f, s1 = 0.f;
    if (row < jb.rows) {
      for (int ct = c; ct < ncolTiles; ct += 64) {
        s0 += base[(i64)ct * prow + row];
        if (ct + 32 < ncolTiles) s1 += base[(i64)(ct + 32) * prow + row];
      }
    }
    sm[threadIdx.x] = s0 + s1;                       // [c][r]
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) {
      if (c < o) sm[c * 8 + r] += sm[(c + o) * 8 + r];
      __syncthreads();
    }
    if (c == 0 && row < jb.rows) atomicAdd(jb.grad + row, sm[r]);
  }
}

// ---- weight preparation: dense (R x C) -> zero-padded K-major copies ---------------------------------
// dst_n (R x ldn) = src ; dst_t (C x ldt) = src^T ; either may be NULL.  Padding is written as zeros.
struct PrepJob { const float* src; float* dst_n; float* dst_t; };
struct PrepJobs { int n; PrepJob j[32]; };

static __global__ void __launch_bounds__(256) prep_weights_kernel(PrepJobs jobs, int R, int C, int ldn, int ldt) {
  __shared__ float tile[32][33];
  const PrepJob jb = jobs.j[blockIdx.z];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = r0 + ty + 8 * i, c = c0 + tx;
    float v = (r < R && c < C) ? jb.src[(i64)r * C + c] : 0.f;
    tile[ty + 8 * i][tx] = v;
    if (jb.dst_n && r < R && c < ldn) jb.dst_n[(i64)r * ldn + c] = v;
  }
  __syncthreads();
  if (jb.dst_t) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = c0 + ty + 8 * i, r = r0 + tx;     // dst_t[c][r]
      if (c < C && r < ldt) jb.dst_t[(i64)c * ldt + r] = tile[tx][ty + 8 * i];
    }
  }
}

}  // namespace dladmm
