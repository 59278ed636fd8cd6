/*
 * dladmm.h -- C ABI of the B200-native D-LADMM hot path (libdladmm.so).
 *
 * The reference (xhchrn/D-LADMM) is pure Python with no FFI; its "plugin API" for this path is the
 * nn.Module contract of DLADMMNet.  The entry points below are what an autograd.Function behind that
 * module binds (ctypes), and each one cites the reference code it replaces:
 *
 *   dladmm_forward   <- DLADMMNet.forward, all K layers
 *                       main_syn_l1l1_scalar.py:80-127 (scalar), main_syn_l1l1_full.py:59-106 (full),
 *                       main_syn_l1l1_scalar_tied.py:82-129 (tied), main_syn_lasso_scalar.py:65-112 (lasso),
 *                       main_lena.py:61-98 (lena), main_syn_l1l1_ltheta.py:67-104 (ltheta)
 *   dladmm_backward  <- autograd of the above (total_loss.backward(), main_syn_l1l1_scalar.py:301)
 *   dladmm_gen_syn   <- gen_syn_data.py:12-47 (Gaussian unit-column A; Bernoulli*Gaussian Z, E; X = AZ+E)
 *   dladmm_objective <- per-layer L1-L1 objective, main_syn_l1l1_scalar.py:333-334
 *
 * Conventions
 *   - All matrices are float32, row-major.  Activations are (features x B) with the batch (problem
 *     instances, "columns") contiguous, exactly as the reference holds them (SURVEY.md preamble).
 *   - Every pointer in dladmm_problem / dladmm_layer is a DEVICE pointer owned by the caller (PyTorch's
 *     caching allocator).  The library never allocates or frees device memory; scratch is the caller's
 *     `workspace` (size from dladmm_workspace_bytes).  The dladmm_layer array itself lives on the HOST.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*); no host synchronisation inside.
 *   - Return value 0 on success, negative dladmm_status otherwise; dladmm_last_error() gives the message
 *     for the calling thread.  There is no CPU fallback: a device that is not sm_100 is an error.
 */
#ifndef DLADMM_H_
#define DLADMM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLADMM_ABI_VERSION 7

#if defined(__GNUC__)
#define DLADMM_API __attribute__((visibility("default")))
#else
#define DLADMM_API
#endif

typedef enum dladmm_status {
  DLADMM_OK = 0,
  DLADMM_ERR_INVALID = -1,      /* bad argument / unsupported shape */
  DLADMM_ERR_CUDA = -2,         /* a CUDA runtime call failed */
  DLADMM_ERR_WORKSPACE = -3,    /* workspace too small */
  DLADMM_ERR_DEVICE = -4        /* device is not sm_100 (no fallback) */
} dladmm_status;

/* E-step family.  A: lena/ltheta, B: scalar/full/tied, C: lasso. */
typedef enum dladmm_family { DLADMM_FAMILY_A = 0, DLADMM_FAMILY_B = 1, DLADMM_FAMILY_C = 2 } dladmm_family;

/* Arithmetic of the two matrix products per layer.  Epilogues are always fp32 with the reference's
 * operation order (no FMA contraction). */
typedef enum dladmm_precision {
  DLADMM_PREC_FP32 = 0,     /* CUDA-core FFMA, fp32 products and accumulation */
  DLADMM_PREC_TF32X3 = 1,   /* tcgen05 kind::tf32, 3 passes (big*big + big*small + small*big), fp32 accumulate in TMEM */
  DLADMM_PREC_TF32 = 2,     /* tcgen05 kind::tf32, single pass (stated-tolerance option: ~1e-3 relative per product) */
  DLADMM_PREC_BF16 = 3,     /* tcgen05 kind::f16 on bf16 operands, single pass, fp32 accumulate (stated-tolerance option:
                               ~4e-3 relative per product).  Iterates stay fp32 in HBM (the reference's API); operands are
                               rounded to bf16 in shared memory by the consumer, weights once per call. */
  DLADMM_PREC_TF32_BF16X2 = 4 /* split-precision, 2 tensor-pass equivalents instead of 3: trunc_tf32(x) * rna_tf32(w) on kind::tf32 plus
                               the two first-order corrections bf16(x) * bf16(w - rna_tf32(w)) and bf16(x - trunc_tf32(x)) *
                               bf16(rna_tf32(w)) on kind::f16 (twice the rate), all accumulated in the same fp32 TMEM accumulator.
                               The corrections are 2^-11 of the product, so their bf16 rounding is a ~2^-20 (unbiased) relative
                               error: fp32-level products like TF32X3 (measured table: profiles/r02_precision_table.md). */
} dladmm_precision;

/* Per-layer metrics accumulated inside the product epilogues (ABI v5): out[k * DLADMM_MET_COUNT + i] = sum over the
 * batch (and over the rows of the array named) for layer k.  What the reference's drivers compute script-side:
 *   training objectives  main_syn_l1l1_scalar.py:289-299 (L1_Z, L1_RES), main_syn_lasso_scalar.py:276-281 (L1_Z, SQ_RES),
 *                        main_lena.py:221-228 (L1_Z, L1_E, DGAP_ATL, DGAP_L, DOT_LX)
 *   evaluation metrics   test_syn_l1l1_scalar.py:486-546 (NMSE: SQERR_Z, SQERR_E), main_lena.py:262-267 (PSNR: SQERR_AZ) */
typedef enum dladmm_metric {
  DLADMM_MET_L1_Z = 0,      /* sum |Z_k|                                                        */
  DLADMM_MET_SQERR_Z = 1,   /* sum (Z_label - Z_k)^2     needs Z_label                          */
  DLADMM_MET_L1_RES = 2,    /* sum |X - A Z_k|           (taken as |E_k - T_{k+1}|)              */
  DLADMM_MET_SQ_RES = 3,    /* sum (X - A Z_k)^2                                                */
  DLADMM_MET_SQERR_E = 4,   /* sum (E_label - E_k)^2     needs E_label                          */
  DLADMM_MET_SQERR_AZ = 5,  /* sum (X_clean - A Z_k)^2   needs X_clean; A Z_k is the accumulator */
  DLADMM_MET_L1_E = 6,      /* sum |E_k|                                                        */
  DLADMM_MET_DOT_LX = 7,    /* sum L_k * X                                                      */
  DLADMM_MET_DGAP_L = 8,    /* sum softplus(L_k - 1) + softplus(-L_k - 1)                       */
  DLADMM_MET_DGAP_ATL = 9,  /* sum softplus(A^T L_k - a) + softplus(-A^T L_k - a), a = dual_alpha: one extra (d x m)(m x B)
                               product per layer, computed only when `want` has this bit (tensor-core precisions only) */
  DLADMM_MET_COUNT = 10
} dladmm_metric;

typedef struct dladmm_metrics {
  uint32_t want;            /* bit i set: metric i is needed (others may be left unspecified in `out`) */
  float dual_alpha;
  const float* Z_label;     /* (d,B) or NULL */
  const float* E_label;     /* (m,B) or NULL */
  const float* X_clean;     /* (m,B) or NULL */
  float* out;               /* device, K * DLADMM_MET_COUNT floats, overwritten by the forward (`objective` must be NULL) */
} dladmm_metrics;

/* A learnable parameter broadcast against a (rows x B) activation.
 *   value(row, col) = ptr[row * row_stride + (col_period ? col % col_period : 0)]
 *   (1,1) scalar: row_stride 0, col_period 0;  (rows,1): row_stride 1, col_period 0;
 *   (rows,bs) per batch slot (main_lena.py:35-36): row_stride bs, col_period bs.
 * `grad` (backward only; may be NULL) has the parameter's own shape, must be zero-initialised by the
 * caller, and is accumulated into.  ptr == NULL means "parameter absent" (e.g. ss1 outside the tied variant). */
typedef struct dladmm_bparam {
  const float* ptr;
  float* grad;
  int32_t row_stride;
  int32_t col_period;
} dladmm_bparam;

/* Per-layer parameters; reference names in comments (state_dict keys `<name>.<k>`). */
typedef struct dladmm_layer {
  dladmm_bparam beta1;    /* beta1.k  : V_k = L_{k-1} + beta1*T_k ; family A also L-update */
  dladmm_bparam beta2;    /* beta2.k  : family A,B E-step */
  dladmm_bparam beta3;    /* beta3.k  : family B,C L-update */
  dladmm_bparam ss1;      /* ss1.k    : tied only (main_syn_l1l1_scalar_tied.py:96) */
  dladmm_bparam ss2;      /* ss2.k (B) or ss2_1.k (C) */
  dladmm_bparam ss2_2;    /* ss2_2.k  : C only (main_syn_lasso_scalar.py:103) */
  dladmm_bparam theta1;   /* active_para.k  (Z threshold); lena: fixed 0.025 */
  dladmm_bparam theta2;   /* active_para1.k (E threshold, A,B); lena: fixed 0.06 */
  const float* W;         /* fc[k].weight (d,m) dense row-major; tied: same pointer in every layer */
  float* gW;              /* d(loss)/dW, (d,m), zero-initialised by caller, accumulated (backward only) */
} dladmm_layer;

typedef struct dladmm_problem {
  int32_t abi_version;    /* DLADMM_ABI_VERSION */
  int32_t family;         /* dladmm_family */
  int32_t precision;      /* dladmm_precision */
  int32_t m, d, K;        /* A is (m,d); K unrolled layers.  K = 0 (ABI v5): only T[0] = A Z0 + E0 - X is computed */
  int64_t B;              /* columns (problem instances) in this call; every (rows x B) array has pitch B */
  int32_t last_only;      /* 0: Z/E/L hold K slabs and T holds K+1 (reference API, a10);
                             1: inference only, slabs are reused modulo 2 (Z/E/L: 2 slabs, T: 2 slabs);
                                the final iterate is slab (K-1)%2, T_K is slab K%2 */
  int32_t reserved;
  const float* A;         /* (m,d) */
  const float* X;         /* (m,B) observation */
  const float* Z0;        /* (d,B) */
  const float* E0;        /* (m,B) */
  const float* L0;        /* (m,B) */
  const dladmm_layer* layers;  /* HOST array of K entries */
  float* Z;               /* out (K,d,B)   */
  float* E;               /* out (K,m,B)   */
  float* L;               /* out (K,m,B)   */
  float* T;               /* out (K+1,m,B), T[0] = A Z0 + E0 - X */
  uint8_t* maskZ;         /* out (K,d,B) or NULL: bit0 = [x-theta>0], bit1 = [-x-theta>0] of the Z prox (training); ABI v6: bit2 = [Z_k>0],
                             bit3 = [Z_k<0], so that the backward of the fused losses takes sign(Z_k) from the mask byte it reads anyway
                             instead of re-reading Z_k (theta may be negative: then bits 0 and 1 do not determine the sign) */
  uint8_t* maskE;         /* out (K,m,B) or NULL: same for the E prox (families A,B) */
  void* workspace;
  size_t workspace_bytes;
  const float* T_init;    /* optional (m,B): T_0 supplied by the caller instead of A Z0 + E0 - X (it is copied to T[0]).
                             Used to advance ONE layer from an arbitrary state (Z0,E0,L0,T_init), e.g. the classical KM
                             step and the safeguard operator of test_syn_l1l1_scalar.py:131-176 */
  /* ABI v4 */
  float* Vsave;           /* optional (K,m,B), training: the forward keeps every V_k = L_{k-1} + beta1_k T_k (the operand of
                             the W V product, main_syn_l1l1_scalar.py:93,111) here instead of one reused scratch slab, and a
                             backward given the same buffer reads it instead of recomputing V_k for the dW product.
                             Contents are defined only between a forward and the backward of the same problem. */
  float* objective;       /* optional out, K floats on the device, overwritten by the forward:
                             objective[k] = sum_b ( objective_alpha*||Z_k[:,b]||_1 + ||X[:,b] - A Z_k[:,b]||_1 )
                             (main_syn_l1l1_scalar.py:289-299, 333-334), accumulated inside the product epilogues from
                             Z_k and E_k - T_{k+1} -- the iterates are not read again.  Works with last_only = 1 too
                             (tensor-core precisions). */
  float objective_alpha;
  int32_t objective_kind; /* ABI v5: 0/1: L1 residual as above; 2: 0.5*||X - A Z_k||_2^2 in place of the L1 residual
                             (the LASSO objective, main_syn_lasso_scalar.py:276-281) */
  /* ABI v5: half-layer entry points (the E -> L -> Z ordering of main_syn_scalar_newS_layerwise.py:82-112 and its
   * safeguard, test_syn_l1l1_newS_Acols.py:136-277, advance "E/L-step of layer k-1, then Z-step of layer k"). */
  int32_t start_half;     /* 1: layer 0 has no Z-step: Z_0 := Z0 as given (Z slab 0 is not written, Z0 is read in its place);
                             the call starts with the E/T/L-step of layer 0.  T_init and T[0] are unused. */
  int32_t stop_half;      /* 1: the E/T/L-step of the last layer is not executed (E, L slab K-1 and T slab K are not written) */
  const dladmm_metrics* metrics;   /* optional (HOST struct), forward only */
} dladmm_problem;

/* Upstream cotangents for backward; each may be NULL (= zero).  Shapes as the forward outputs. */
typedef struct dladmm_cotangents {
  const float* gZ;        /* (K,d,B)   */
  const float* gE;        /* (K,m,B)   */
  const float* gL;        /* (K,m,B)   */
  const float* gT;        /* (K+1,m,B) */
  /* Optional fused training loss (adds its cotangents inside the backward kernels, nothing is materialised):
   *   loss = scale0 * sum_k w_k * sum_b ( alpha*||Z_k[:,b]||_1 + ||X[:,b] - A Z_k[:,b]||_1 )
   * i.e. the per-layer L1-L1 objective of main_syn_l1l1_scalar.py:289-299 with X - A Z_k taken as E_k - T_{k+1}.
   * loss_scale points to ONE device float = d(final loss)/d(loss) * scale0 (e.g. upstream gradient / B). */
  int32_t loss_kind;                /* 0: none, 1: L1-L1 objective, 2 (ABI v5): LASSO objective, the residual term being
                                       0.5*||X - A Z_k[:,b]||_2^2 (main_syn_lasso_scalar.py:276-281) */
  float loss_alpha;
  const float* loss_layer_weight;   /* HOST array of K floats w_k */
  const float* loss_scale;          /* DEVICE pointer to one float */
  /* ABI v7, optional: HOST array of K cudaEvent_t handles (NULL, or entries NULL).  Event k is recorded on `stream` right
   * after the weight-gradient product of layer k has been enqueued (layers run K-1 ... 0): once it fires, gW of layer k
   * holds the contributions of layers k ... K-1, i.e. it is final unless a layer below k shares the weight (tied / ptied).
   * A data-parallel caller waits for these on a side stream to allreduce finished weight gradients while the layers
   * below are still running (buckets in reverse layer order); every other gradient is final when the call's
   * last kernel retires.  The FFMA path records all of them after its last kernel. */
  void* const* layer_events;
} dladmm_cotangents;

typedef struct dladmm_caps {
  int32_t abi_version;
  int32_t cc_major, cc_minor;
  int32_t sm_count;
  int32_t supported;      /* 1 iff cc 10.x */
  int32_t has_tcgen05;    /* tensor-core precisions available in this build */
  int64_t total_mem;
} dladmm_caps;

typedef struct dladmm_gen_desc {
  int32_t m, d;
  int64_t B;              /* columns generated in this call */
  int64_t col_offset;     /* global index of the first column (shards reproduce the same stream) */
  uint64_t seed;
  float p;                /* Bernoulli keep probability (gen_syn_data.py -p) */
  float mu, sigma;        /* Gaussian amplitude (gen_syn_data.py:24-25) */
  int32_t dense_noise;    /* 0: E = Bern(p)*N(mu,sigma); 1: E ~ N(0, sigma_e) dense (gen_syn_unseen_data_lasso.py:41-42) */
  float sigma_e;
  int32_t generate_A;     /* 1: draw A ~ N(0,1) and normalise columns (gen_syn_data.py:14-16); 0: use A as given */
  float* A;               /* (m,d) in/out */
  float* Zs;              /* out (d,B) ground-truth sparse code */
  float* Es;              /* out (m,B) ground-truth noise */
  float* X;               /* out (m,B) = A Zs + Es */
  void* workspace;        /* scratch, dladmm_gen_workspace_bytes(m, d) bytes */
  size_t workspace_bytes;
  int32_t amplitude;      /* ABI v4: non-zero amplitudes of Zs and of a sparse Es, g ~ N(mu, sigma):
                             0: g (gen_syn_data.py:30-33); 1: cos(g) (gen_syn_unseen_data_cosine.py:33-38);
                             2: 1/(1+exp(g)) (gen_syn_unseen_data_logistic.py:33-38) */
  int32_t reserved;
} dladmm_gen_desc;

/* Bytes of scratch the caller must provide in problem->workspace. */
DLADMM_API size_t dladmm_workspace_bytes(const dladmm_problem* p, int for_backward);

DLADMM_API int dladmm_forward(const dladmm_problem* p, void* stream);

/* Requires the outputs and masks of a forward over the same problem (last_only == 0, masks non-NULL). */
DLADMM_API int dladmm_backward(const dladmm_problem* p, const dladmm_cotangents* g, void* stream);

DLADMM_API size_t dladmm_gen_workspace_bytes(int32_t m, int32_t d);

/* Counter-based (Philox4x32-10) generator: the value of element (row, global column) depends only on
 * (seed, row, col_offset + column), so column shards on different GPUs reproduce one global stream. */
DLADMM_API int dladmm_gen_syn(const dladmm_gen_desc* g, void* stream);

/* Per-layer L1-L1 objective summed over the batch: out[k] = sum_{b} ( alpha*||Z_k[:,b]||_1 + ||X[:,b] - A Z_k[:,b]||_1 ),
 * k = 0..K-1, from the iterates a forward (last_only == 0) left in p->Z/E/T, using X - A Z_k = E_k - T_{k+1}
 * (no extra matrix product).  `out` is K floats on the device, overwritten. */
DLADMM_API int dladmm_objective(const dladmm_problem* p, float alpha, float* out, void* stream);

/* Safeguarded evaluation (test_syn_l1l1_scalar.py:163-176, 238-274).
 * dladmm_sg_norm:   out[b] = || [ beta*Tn[:,b] ; c*(En[:,b] - 2 Ek[:,b] + Ep[:,b]) ] ||_2   (the norm of S(u), all (m,B))
 * dladmm_sg_select: keep[b] = snorm[b] < one_minus_delta * mu[b]; out = keep ? a : b column-wise for each pair. */
typedef struct dladmm_sg_pair { const float* a; const float* b; float* out; int32_t rows; } dladmm_sg_pair;
DLADMM_API int dladmm_sg_norm(int32_t m, int64_t B, float beta, float c, const float* Tn, const float* En, const float* Ek,
                              const float* Ep, float* out, void* stream);
DLADMM_API int dladmm_sg_select(int32_t n_arrays, const dladmm_sg_pair* pairs, int64_t B, const float* snorm, const float* mu,
                                float one_minus_delta, float* keep, void* stream);

/* Safeguard of the E -> L -> Z ordering (Snorm_ELZ, test_syn_l1l1_newS_Acols.py:174-192) without the d x d matrix
 * P2 = I/(beta ss1) - A^T A:  out[b] = sqrt( ||Tnn[:,b]||^2 + inv_beta_ss1 * ||Znn[:,b] - Zn[:,b]||^2 - ||Tnn[:,b] - Tn[:,b]||^2 )
 * with Tn = A Zn + En - X and Tnn = A Znn + En - X (so Tnn - Tn = A (Znn - Zn)).  Tn, Tnn are (m,B); Zn, Znn (d,B).
 * E_sub / E_add (both or neither, (m,B)): the caller holds A Znn + E' - X (the T output of a following E-step) instead of
 * Tnn; then Tnn is taken as (Tnn_given - E_sub) + E_add with E_sub = E', E_add = En. */
DLADMM_API int dladmm_sg_norm_elz(int32_t m, int32_t d, int64_t B, float inv_beta_ss1, const float* Tnn, const float* Tn,
                                  const float* Znn, const float* Zn, const float* E_sub, const float* E_add, float* out,
                                  void* stream);

/* mu_k updaters of mu_updater.py:18-72 fused with the selection: as dladmm_sg_select, then per column
 *   method 1 (EMA): upd = param*snorm + (1-param)*mu;  2 (GS): upd = (1-param)*mu;  3 (RT): upd = snorm;
 *   mu[b] = keep[b] ? upd : mu[b]   (in place; method 0 leaves mu alone; method 4, the reference's "None" updater
 *   mu_updater.py:97-108, sets mu[b] = 1e10 for every column), and *fallbacks (one device float, accumulated)
 *   += number of columns with keep == 0.  One launch for keep + mu + count, one for the copies. */
DLADMM_API int dladmm_sg_select_update(int32_t n_arrays, const dladmm_sg_pair* pairs, int64_t B, const float* snorm, float* mu,
                                       float one_minus_delta, int32_t method, float param, float* keep, float* fallbacks,
                                       void* stream);

DLADMM_API int dladmm_query(int device, dladmm_caps* caps);

/* Measurement hooks (bench.py).  Kernel kinds for the per-kind timers: */
typedef enum dladmm_kernel_kind {
  DLADMM_KIND_PREP = 0,        /* weight padding / transposition */
  DLADMM_KIND_GEMM_T0 = 1,     /* T_0 = A Z0 + E0 - X */
  DLADMM_KIND_GEMM_Z = 2,      /* W V product + Z prox epilogue */
  DLADMM_KIND_GEMM_ELT = 3,    /* A Z product + E/T/L epilogue */
  DLADMM_KIND_BWD_ELEM = 4,    /* top-layer elementwise cotangent flow */
  DLADMM_KIND_BWD_GEMM_DZ = 5, /* A^T dR + mask epilogue */
  DLADMM_KIND_BWD_GEMM_DW = 6, /* dx1 V^T (reduction over the batch) */
  DLADMM_KIND_BWD_GEMM_DV = 7, /* W^T dx1 + fused elementwise backward of the layer below */
  DLADMM_KIND_BWD_REDUCE = 8,  /* second stage of parameter-gradient reductions */
  DLADMM_KIND_GEN = 9,         /* synthetic data generator */
  DLADMM_KIND_OBJECTIVE = 10,
  DLADMM_KIND_METRIC_GEMM = 11, /* A^T L_k product of the dual-gap metric */
  DLADMM_KIND_SAFEGUARD = 12,   /* safeguard norms / selection */
  DLADMM_KIND_FWD_PERSISTENT = 13, /* all-layer persistent forward kernel */
  DLADMM_KIND_COUNT = 14
} dladmm_kernel_kind;

/* Total kernels this library has launched in this process (monotonic). */
DLADMM_API int64_t dladmm_launch_count(void);

/* While profiling is on, every kernel launch is bracketed by CUDA events on its stream (a measurement
 * pass only: the events cost a little).  dladmm_profile_stop synchronises those events and returns, per
 * kind, the summed device time in milliseconds and the number of launches (arrays of DLADMM_KIND_COUNT). */
DLADMM_API int dladmm_profile_start(void);
DLADMM_API int dladmm_profile_stop(double* ms_by_kind, int64_t* launches_by_kind);

/* Debugging aid: with DLADMM_PF_TRACE=1 in the environment the all-layer persistent forward kernel stamps clock64 per unit for
 * its first 4 CTAs ([4][512][8] values, layout in csrc/umma_persist.cuh); this copies the stamps of the last forward to the host
 * (synchronises the device).  Returns the number of values, or -1 when tracing is off / capacity too small. */
DLADMM_API int64_t dladmm_debug_trace(int64_t* host_out, int64_t capacity);

DLADMM_API const char* dladmm_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* DLADMM_H_ */
