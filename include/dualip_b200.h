/*
 * dualip_b200 — C-ABI of the B200-native DuaLip dual-ascent hot path.
 *
 * The reference (linkedin/DuaLip v5.0.1) has no FFI: its seam is the Python
 * protocol  f.calculate(dual_val, gamma, save_primal) -> ObjectiveResult
 * (reference src/dualip/objectives/matching.py:116-188) called once per
 * iteration by AcceleratedGradientDescent.maximize (optimizers/agd.py:150-160).
 * The entry points below are what a binding for that seam calls.  Plain
 * pointers and sizes only; no torch types.  All `*_dev` pointers are CUDA device
 * pointers on the plan's device; every launch goes to the `stream` argument
 * (a cudaStream_t passed as void*), performs no host synchronisation and no
 * allocation, and is therefore CUDA-graph capturable (except the *_host calls).
 *
 * Return value: 0 on success, negative DUALIP_E* on failure; the message is
 * available from dualip_last_error() (thread-local).
 */
#ifndef DUALIP_B200_H
#define DUALIP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DUALIP_B200_ABI_VERSION 2

/* error codes */
#define DUALIP_OK 0
#define DUALIP_EINVAL (-1)   /* bad argument / unsupported layout  (reference: ValueError)   */
#define DUALIP_ECUDA (-2)    /* CUDA runtime error                                            */
#define DUALIP_ENOMEM (-3)   /* allocation failure                                            */
#define DUALIP_ERANGE (-4)   /* size outside what this build supports                         */

/* projection kinds (reference src/dualip/projections/{box,cone,simplex}.py) */
#define DUALIP_PROJ_CLAMP 0       /* box(lower,upper), cone(lower) / cone(upper) / cone(): x = min(max(v,lo),hi); +-inf = open */
#define DUALIP_PROJ_SIMPLEX 1     /* "simplex":    {x>=0, sum x <= z}, batched Duchi with pre-clamp (simplex.py:126-236, :248-255) */
#define DUALIP_PROJ_SIMPLEX_EQ 2  /* "simplex_eq": {x>=0, sum x  = z}, same routine without the feasibility branch (:267-274)      */
#define DUALIP_PROJ_SIMPLEX_BISECT 3     /* "simplex",    method="bisection_search" (simplex.py:6-123, inequality=True): no pre-clamp,
                                            feasible needs every entry >= -1e-6, top-2 shortcut, 19 halvings of [-1, 0]       */
#define DUALIP_PROJ_SIMPLEX_EQ_BISECT 4  /* "simplex_eq", method="bisection_search" (inequality=False)                          */
/* Both bisection kinds depend on the padded length L of the reference's block (pad_len below): the zero padding takes part in
 * the maximum, the top-2 test and the sums.  They run in the fused kernel for columns of up to 1024 entries. */

/* flags of a projection class */
#define DUALIP_PROJ_FLAG_D1_UNPADDED 1u /* 1-nnz columns of this class sit in a bucket whose padded length L is 1, so the
                                           reference skips its top-2 shortcut for them (simplex.py:166 `if L > 1`)              */

/* One row per distinct (proj_type, proj_params) of the reference's projection_map
 * (projections/base.py:8-12).  Columns reference it through col_class[]. */
typedef struct dualip_proj_class {
  int32_t kind;   /* DUALIP_PROJ_*                                                     */
  float lo;       /* CLAMP: lower bound or -INFINITY                                    */
  float hi;       /* CLAMP: upper bound or +INFINITY                                    */
  float z;        /* SIMPLEX*: radius, fl32(z)                                          */
  float z_thr;    /* SIMPLEX : fl32(double(z) + 1e-6), the feasibility threshold (simplex.py:154) */
  uint32_t flags; /* DUALIP_PROJ_FLAG_*                                                 */
} dualip_proj_class;

/* Description of one (local shard of a) matching LP: A and c are CSC with ONE shared
 * sparsity pattern (reference MatchingInputArgs, matching.py:12-22). */
typedef struct dualip_csc_desc {
  int64_t n_cols;          /* entities (columns of A)                                   */
  int64_t nnz;             /* stored entries E                                          */
  int32_t n_rows;          /* dual dimension m                                          */
  int32_t index_bits;      /* 32 or 64: width of ccol_dev[] and row_dev[] entries       */
  const void* ccol_dev;    /* n_cols+1 column pointers, non-decreasing, ccol[0]=0       */
  const void* row_dev;     /* nnz row indices in [0,m)                                   */
  const float* a_dev;      /* nnz values of A                                            */
  const float* c_dev;      /* nnz values of c                                            */
                           /* all four arrays are COPIED into the plan's own layout; the caller may free them */
  const uint8_t* col_class_dev; /* n_cols class ids into classes[], or NULL = all columns class 0 */
  const dualip_proj_class* classes; /* host array                                        */
  int32_t n_classes;       /* 1..255                                                    */
  int32_t device;          /* CUDA device ordinal                                       */
  const int32_t* pad_len;  /* host array n_classes x DUALIP_PAD_BUCKETS, or NULL: the padded length L of the reference's
                              [L x K] block that a column of class k with d entries is projected in
                              (utils/sparse_utils.py:197,207), at pad_len[k*DUALIP_PAD_BUCKETS + ceil(log2(d))]; 0 = d.
                              "simplex_eq" depends on it (simplex.py:160-161): a column whose
                              clamped sum is below z gets (z - sum)/L added to every entry; so do both bisection kinds
                              (the padding is part of the block they search on).  The reference derives L from its length
                              buckets (objectives/matching.py:87-114). */
} dualip_csc_desc;
#define DUALIP_PAD_BUCKETS 32

typedef struct dualip_plan dualip_plan;

/* Scalars of one evaluation, all double (reference ObjectiveResult, types.py:32-41). */
typedef struct dualip_scalars {
  double dual_objective;      /* c.x + reg_penalty + lambda.(Ax-b)   (matching.py:33)   */
  double primal_objective;    /* c.x                                  (matching.py:160)  */
  double reg_penalty;         /* gamma/2 * ||x||^2                    (matching.py:157)  */
  double dual_val_times_grad; /* lambda.(Ax-b)                        (matching.py:167)  */
  double max_pos_slack;       /* max(max(Ax-b),0)                     (matching.py:168)  */
  double sum_pos_slack;       /* sum relu(Ax-b)                       (matching.py:169)  */
  double x_sq_norm;           /* ||x||^2                                                  */
  double grad_sq_norm;        /* ||Ax-b||^2                                               */
} dualip_scalars;

/* calc flags */
#define DUALIP_CALC_DEFAULT 0u

int dualip_abi_version(void);
const char* dualip_last_error(void);

/* Build the device-side slab layout for a shard (setup time; synchronises).  Replaces
 * MatchingSolverDualObjectiveFunction.__init__ / _compute_buckets (matching.py:43-114). */
int dualip_plan_create(dualip_plan** out, const dualip_csc_desc* desc);
void dualip_plan_destroy(dualip_plan* plan);

/* Re-cuts the per-CTA slab ranges of the plan from the per-CTA durations the hot kernel recorded in its last launch (the
 * plan starts from a fitted cost table; the cost of a simplex column depends on the iterate).  Synchronises `stream`.
 * Results do not change: columns are projected independently of the partition and the fixed-point gradient sums are exact
 * integers (their scale is kept unless a range's overflow bound forces it down).  Call it a few times early in a solve. */
int dualip_plan_rebalance(dualip_plan* plan, void* stream);

/* Introspection: fills up to `cap` int64 values:
 * [0] n_slabs [1] n_long_cols [2] n_ctas [3] threads/cta [4] smem bytes/cta [5] row index bits
 * [6] smem mode (0: lambda+grad in smem, 1: grad in smem, 2: neither) [7] stored slab elements (incl. padding)
 * [8] kernel launches per calc  [9] plan-owned device bytes [10] columns stored in slabs [11] nnz
 * [12] 1 if the gradient is accumulated in 32-bit fixed point (deterministic), 0 for fp32 atomics
 * [13] F: fixed-point fraction bits (value * 2^F)  [14] worst-row rounding-error estimate * 1e12
 * [15] longest column whose slab is staged through shared memory by the TMA engine (0: plain vector loads)
 * [16] 1 if the rows are stored scaled by per-row powers of two (fixed-point resolution per row; results unchanged)
 * [17] n_mid_cols: columns of 21..1024 entries kept column-contiguous and processed a warp per column inside the same
 *      launch ([1] counts only the columns beyond 1024 entries, which take the separate long-column kernel)
 * [18] 1 if this plan's own launches share the m-length tail among all CTAs (m >= 16384 or DUALIP_GRID_TAIL=1; sharded
 *      launches of 4 or more ranks also do so from m >= 8192 unless DUALIP_GRID_TAIL=0)
 * [19] 0, or 2 once a grid-wide barrier of the all-CTA tail timed out (the CTAs of a launch were not co-resident, e.g. another
 *      process held SMs for seconds): results of that launch are invalid.  Read from mapped host memory, no synchronisation;
 *      sharded launches report the same event through dualip_peer_status.
 * [20] 1 if the most recent launch of the hot kernel on this plan used the all-CTA tail */
int dualip_plan_info(const dualip_plan* plan, int64_t* out, int cap);

/* One evaluation of the dual at lambda on this shard, epilogue included (single device).
 * Replaces calculate() (matching.py:116-188).
 *   lambda_dev  m floats.            b_dev  m floats, or NULL (treated as 0: "local shard" mode, matching.py:56)
 *   grad_out_dev m floats  = A x*(lambda) - b
 *   scalars_out_dev          one dualip_scalars (device memory)
 *   x_out_dev   nnz floats or NULL (save_primal; CSC value order, matching.py:185-187)
 *   diag_out_dev nnz bytes or NULL: at the FIRST entry of every simplex column, branch|rho<<2 with
 *               branch 0=feasible 1=top-2 shortcut 2=Duchi (rho = support size, saturated at 63); tests only. */
int dualip_matching_calc(dualip_plan* plan, const float* lambda_dev, const float* b_dev, double gamma,
                         float* grad_out_dev, dualip_scalars* scalars_out_dev, float* x_out_dev,
                         uint8_t* diag_out_dev, uint32_t flags, void* stream);

/* Sharded evaluation, step 1: this shard's partial sums, packed for ONE all-reduce:
 * partial_out_dev[0..m) = sum_j a_rj x_rj over local columns, [m] = c.x, [m+1] = ||x||^2  (m+2 floats).
 * Replaces the local calculate() + three dist.reduce calls (matching.py:261-274). */
int dualip_matching_partial(dualip_plan* plan, const float* lambda_dev, double gamma, float* partial_out_dev,
                            float* x_out_dev, uint8_t* diag_out_dev, uint32_t flags, void* stream);

/* Sharded evaluation, step 2 (after the all-reduce; identical on every rank): m-length tail.
 * Replaces matching.py:280-299.  No plan needed. */
int dualip_matching_epilogue(const float* partial_sum_dev, int32_t m, const float* lambda_dev, const float* b_dev,
                             double gamma, float* grad_out_dev, dualip_scalars* scalars_out_dev, void* stream);

/* Host-buffer convenience used for the end-to-end measurement: copies lambda (and nothing else)
 * host->device, evaluates, copies grad and scalars device->host, synchronises the stream.
 * b_dev stays device-resident (it is part of the problem, like A and c). */
int dualip_matching_calc_host(dualip_plan* plan, const float* lambda_host, const float* b_dev, double gamma,
                              float* grad_out_host, dualip_scalars* scalars_out_host, void* stream);

/* ---- generic (non-block) LP objective: reference src/dualip/objectives/miplib.py:60-109 ----
 * A is given twice (CSR for A x, CSC for A^T lambda), int32 indices; all arrays are BORROWED device pointers that must
 * stay alive while the description is in use.  lo/hi: per-variable bounds, -INFINITY / +INFINITY for open sides (the
 * box / cone entries of the reference's projection_map, miplib.py:80-90).  row_scale: 1/||A_r||_2 per row when the
 * objective was built with use_jacobi_precondition (miplib.py:48-58,73-74,92-95), else NULL. */
typedef struct dualip_lp_desc {
  int32_t m;                      /* constraints (rows of A, length of lambda)            */
  int32_t n;                      /* variables                                             */
  int64_t nnz;
  const int32_t* csr_rowptr_dev;  /* m+1 */
  const int32_t* csr_col_dev;     /* nnz */
  const float* csr_val_dev;       /* nnz */
  const int32_t* csc_colptr_dev;  /* n+1 */
  const int32_t* csc_row_dev;     /* nnz */
  const float* csc_val_dev;       /* nnz */
  const float* c_dev;             /* n   */
  const float* b_dev;             /* m   */
  const float* lo_dev;            /* n   */
  const float* hi_dev;            /* n   */
  const float* row_scale_dev;     /* m or NULL */
  int32_t device;
} dualip_lp_desc;

/* One evaluation of the dual of the generic LP at lambda: x_out = clamp(-(A^T lambda' + c)/gamma, lo, hi),
 * grad_out = row_scale * (A x - b), scalars as for the matching objective (dual_val_times_grad = lambda'.(Ax-b) with
 * lambda' = row_scale * lambda).  scratch_dev: 8 doubles, zero on first use (the call leaves them zeroed).
 * Replaces MIPLIB2017ObjectiveFunction.calculate.  Asynchronous on `stream`, CUDA-graph capturable. */
int dualip_lp_calc(const dualip_lp_desc* desc, const float* lambda_dev, double gamma, float* x_out_dev,
                   float* grad_out_dev, dualip_scalars* scalars_out_dev, double* scratch_dev, void* stream);

/* ---- device-resident Maximizer state (reference optimizers/agd.py:121-229, agd_utils.py:4-89) ---- */
typedef struct dualip_agd dualip_agd;

/* Creates optimizer state for an m-vector on `device`: x = y = initial (device pointer, or NULL for zeros),
 * 15-deep history ring (agd_utils.py:71). */
int dualip_agd_create(dualip_agd** out, int32_t m, int32_t device, const float* initial_dev,
                      const uint8_t* equality_mask_dev /* m bytes or NULL */, double initial_step_size,
                      double max_step_size, int32_t history_len /* 15 */);
void dualip_agd_destroy(dualip_agd* agd);
/* Device pointer of the current evaluation point x (m floats) / last projected iterate y. */
const float* dualip_agd_x(const dualip_agd* agd);
const float* dualip_agd_y(const dualip_agd* agd);
/* Copies x and/or y (m floats each) into caller buffers on the same device (either may be NULL). */
int dualip_agd_get(dualip_agd* agd, float* x_out_dev, float* y_out_dev, void* stream);
/* One accelerated step from grad (m floats) evaluated at x, with momentum beta_i (agd.py:93-100):
 * step size from the Lipschitz history (agd_utils.py:65-89), y_new = proj(x + step*grad), x = y_new(1-beta)+y*beta.
 * If decay_now != 0: afterwards max_step_size = step*decay_factor (agd.py:102-109; gamma itself is host state).
 * Writes (dual_objective from scalars_dev, step) to log slot `iter_index` if 0 <= iter_index < capacity.  No host sync. */
int dualip_agd_step(dualip_agd* agd, const float* grad_dev, const dualip_scalars* scalars_dev, float beta,
                    int32_t decay_now, double decay_factor, int32_t iter_index, void* stream);
/* Sharded path: the same step taken directly from the all-reduced packed sums of dualip_matching_partial
 * ([sum_j a_rj x_rj (m) | c.x | ||x||^2]); also performs dualip_matching_epilogue's m-length tail at the evaluation point x
 * (grad_out_dev, scalars_out_dev are written), saving that launch.  Replaces matching.py:280-299 + agd.py:163-187. */
int dualip_agd_step_sharded(dualip_agd* agd, const float* partial_sum_dev, const float* b_dev, double gamma,
                            float* grad_out_dev, dualip_scalars* scalars_out_dev, float beta, int32_t decay_now,
                            double decay_factor, int32_t iter_index, void* stream);
/* ---- peer-memory exchange of the sharded path (one process per GPU, NVLink / NVSwitch) ----
 * Replaces the collective of the sharded path (the reference's three dist.reduce + barrier, matching.py:272-277; this
 * library's single NCCL all-reduce) by loads from the peers' memory inside the update kernel: every rank owns an exchange
 * window {arrival flags | two slots of m+2 floats}; dualip_matching_partial writes the shard's packed sums into the slot
 * of the upcoming step (dualip_peer_next_slot); dualip_agd_step_peer then (1) stores its step number into every peer's
 * flag array (st.release.sys over NVLink), (2) waits until all peers have arrived (ld.acquire.sys on its own flags),
 * (3) sums the W slots in rank order -- so every rank obtains bit-identical sums -- and (4) runs the objective's tail
 * and the accelerated update (as dualip_agd_step_sharded).  Two slots suffice: a slot is rewritten two steps later, after
 * a barrier that every reader of the old contents has passed.  The wait is bounded (20 s, or DUALIP_PEER_TIMEOUT_MS): on time-out the status word
 * is set (dualip_peer_status) and the kernel proceeds, it never hangs.  All ranks must take the same number of steps. */
typedef struct dualip_peer dualip_peer;
#define DUALIP_PEER_HANDLE_BYTES 64
#define DUALIP_PEER_MAX_WORLD 16
int dualip_peer_create(dualip_peer** out, int32_t m, int32_t rank, int32_t world, int32_t device);
void dualip_peer_destroy(dualip_peer* peer);
/* CUDA IPC handle of this rank's window (DUALIP_PEER_HANDLE_BYTES bytes) for the caller to all-gather. */
int dualip_peer_export(dualip_peer* peer, uint8_t* handle_out);
/* Opens the windows of all ranks: world x DUALIP_PEER_HANDLE_BYTES bytes in rank order (the own entry is ignored). */
int dualip_peer_connect_ipc(dualip_peer* peer, const uint8_t* handles);
/* Same with device pointers the caller obtained itself (windows of several ranks living in one process, symmetric-memory
 * allocators): world pointers in rank order, each a dualip_peer_window() of that rank. */
int dualip_peer_connect_ptrs(dualip_peer* peer, void* const* windows);
void* dualip_peer_window(dualip_peer* peer);
/* Where dualip_matching_partial must write the packed sums consumed by the NEXT dualip_agd_step_peer (m+2 floats). */
float* dualip_peer_next_slot(dualip_peer* peer);
/* 0 = fine, 1 = a wait timed out (results after that step are invalid).  Synchronises the stream. */
int dualip_peer_status(dualip_peer* peer, int32_t* status_out, void* stream);
int dualip_agd_step_peer(dualip_agd* agd, dualip_peer* peer, const float* b_dev, double gamma, float* grad_out_dev,
                         dualip_scalars* scalars_out_dev, float beta, int32_t decay_now, double decay_factor,
                         int32_t iter_index, void* stream);

/* 0 = fine, 1 = a wait timed out; reads a flag in mapped host memory, does not synchronise (poll it every few steps). */
int dualip_peer_status_nowait(dualip_peer* peer);

/* ---- one launch per iteration: evaluation at the optimizer's own evaluation point + the accelerated step ----
 * dualip_matching_ascent_step = dualip_matching_calc(lambda = dualip_agd_x(agd)) followed by dualip_agd_step, in ONE kernel
 * launch: the CTA of the fused kernel that finishes last runs the objective's m-length tail (grad_out_dev, scalars_out_dev
 * are written as by dualip_matching_calc) and then takes the step on the device-resident state.  Replaces one turn of the
 * loop of AcceleratedGradientDescent.maximize (reference optimizers/agd.py:150-206).  x_out_dev as for dualip_matching_calc. */
int dualip_matching_ascent_step(dualip_plan* plan, dualip_agd* agd, const float* b_dev, double gamma, float* grad_out_dev,
                                dualip_scalars* scalars_out_dev, float* x_out_dev, float beta, int32_t decay_now,
                                double decay_factor, int32_t iter_index, void* stream);
/* Sharded twin: dualip_matching_partial into this rank's exchange slot + dualip_agd_step_peer, in ONE launch.  The last CTA
 * publishes the shard's packed sums, waits for the peers' arrival flags, adds all slots in rank order (peer-memory loads
 * over NVLink), runs the tail and the step.  No collective call, no second launch (reference: three dist.reduce + barrier
 * + two broadcasts per iteration, objectives/matching.py:272-277, optimizers/agd.py:204-206). */
int dualip_matching_ascent_step_peer(dualip_plan* plan, dualip_agd* agd, dualip_peer* peer, const float* b_dev, double gamma,
                                     float* grad_out_dev, dualip_scalars* scalars_out_dev, float beta, int32_t decay_now,
                                     double decay_factor, int32_t iter_index, void* stream);

/* Sharded evaluation for a caller that keeps the dual iterate itself (the host-buffer path; a per-iteration callback):
 * dualip_matching_partial + the exchange through peer memory + dualip_matching_epilogue in ONE launch, no optimizer step.
 * Replaces the distributed calculate (objectives/matching.py:247-307) including its three dist.reduce + barrier.
 * COLLECTIVE: every rank must make the same sequence of exchange calls.  The _host variant copies lambda host->device and
 * grad / scalars device->host and synchronises the stream (and reports a timed-out exchange as an error). */
int dualip_matching_calc_peer(dualip_plan* plan, dualip_peer* peer, const float* lambda_dev, const float* b_dev, double gamma,
                              float* grad_out_dev, dualip_scalars* scalars_out_dev, void* stream);
int dualip_matching_calc_peer_host(dualip_plan* plan, dualip_peer* peer, const float* lambda_host, const float* b_dev,
                                   double gamma, float* grad_out_host, dualip_scalars* scalars_out_host, void* stream);

/* ---- scheduled launches and CUDA-graph replay (reference loop: optimizers/agd.py:150-206) ----
 * dualip_matching_ascent_step takes gamma, beta, the decay flag and the log slot as arguments, so every iteration is a
 * different launch.  With a device-resident schedule the kernel looks them up itself at the number of steps the state
 * has taken (a device counter): the launch is then IDENTICAL for every iteration, and `chunk` launches captured once in a
 * CUDA graph replay as one submission.  gamma_host[i] / beta_host[i]: values of iteration i (0-based) exactly as the
 * host loop would pass them (gamma already decayed, agd.py:102-109; beta from agd.py:93-100); decay_now_host[i] != 0
 * where max_step_size = step * decay_factor follows iteration i (may be NULL).  Synchronises the device. */
int dualip_agd_set_schedule(dualip_agd* agd, int32_t n_iters, const double* gamma_host, const float* beta_host,
                            const uint8_t* decay_now_host, double decay_factor);
/* Steps enqueued on this state so far (host-side count; the schedule index of the next launch). */
long long dualip_agd_steps_launched(const dualip_agd* agd);
/* One scheduled iteration: dualip_matching_ascent_step (peer == NULL) or dualip_matching_ascent_step_peer. */
int dualip_matching_ascent_step_scheduled(dualip_plan* plan, dualip_agd* agd, dualip_peer* peer, const float* b_dev,
                                          float* grad_out_dev, dualip_scalars* scalars_out_dev, void* stream);
/* `chunk` scheduled iterations as one CUDA graph (captured on a private stream; nothing is executed by create).  The
 * pointers are baked into the graph: plan, state, window and buffers must outlive it, and a dualip_plan_rebalance after
 * the capture invalidates it (build the graph once the plan has settled).  launch enqueues the whole chunk on `stream`. */
typedef struct dualip_ascent_graph dualip_ascent_graph;
int dualip_ascent_graph_create(dualip_ascent_graph** out, dualip_plan* plan, dualip_agd* agd, dualip_peer* peer,
                               const float* b_dev, float* grad_out_dev, dualip_scalars* scalars_out_dev, int32_t chunk);
int dualip_ascent_graph_launch(dualip_ascent_graph* graph, void* stream);
void dualip_ascent_graph_destroy(dualip_ascent_graph* graph);

/* ---- host-resident twin of the Maximizer state: for callers that keep the dual iterate in host memory and hand it to
 * dualip_matching_calc_host every iteration (the host-buffer path).  Same update as dualip_agd_step (agd.py:163-187,
 * agd_utils.py:4-89) in one call instead of ~25 tensor operations; needs no GPU.  x (the evaluation point) lives in
 * pinned memory when a CUDA device is present. */
typedef struct dualip_agd_host dualip_agd_host;
int dualip_agd_host_create(dualip_agd_host** out, int32_t m, const float* initial_host, const uint8_t* equality_mask_host,
                           double initial_step_size, double max_step_size, int32_t history_len /* 15, <= 64 */);
void dualip_agd_host_destroy(dualip_agd_host* agd);
float* dualip_agd_host_x(dualip_agd_host* agd);
float* dualip_agd_host_y(dualip_agd_host* agd);
int dualip_agd_host_step(dualip_agd_host* agd, const float* grad_host, float beta, int32_t decay_now, double decay_factor,
                         double* step_out);

/* One whole iteration for a dual iterate kept in host memory: the evaluation point of `agd` goes host->device, the dual is
 * evaluated there (peer != NULL: sharded, sums exchanged through peer memory), gradient and scalars come back into the
 * caller's (pinned) buffers, the stream is synchronised and the host state takes the accelerated step.  Replaces one turn of
 * AcceleratedGradientDescent.maximize with CPU tensors (optimizers/agd.py:150-206) by ONE native call. */
int dualip_matching_step_host(dualip_plan* plan, dualip_peer* peer, dualip_agd_host* agd, const float* b_dev, double gamma,
                              float beta, int32_t decay_now, double decay_factor, float* grad_out_host,
                              dualip_scalars* scalars_out_host, double* step_out, void* stream);

/* Copies log entries [0,count) to host: dual_objective and step size per iteration. Synchronises. */
int dualip_agd_read_log(dualip_agd* agd, int32_t count, double* dual_obj_host, double* step_host, void* stream);
int dualip_agd_reserve_log(dualip_agd* agd, int32_t capacity);

/* ---- setup-time helper kernels ---- */
/* In-place Jacobi row scaling (reference preprocessing/precondition.py:8-29): norms_out[r] = ||A_r||_2,
 * a[e] /= norms[row[e]], b[r] /= norms[r].  index_bits as above. Synchronises. */
int dualip_jacobi_precondition(float* a_dev, const void* row_dev, int32_t index_bits, int64_t nnz, float* b_dev,
                               int32_t m, float* norms_out_dev, int32_t device, void* stream);

/* Sharded Jacobi scaling: local squared row norms (double, m), to be all-reduced by the caller, then
 * a[e] *= scale[row[e]] with scale = 1/sqrt(sum).  Asynchronous on `stream`. */
int dualip_row_sq_norms(const float* a_dev, const void* row_dev, int32_t index_bits, int64_t nnz, int32_t m,
                        double* sq_out_dev, int32_t device, void* stream);
int dualip_scale_rows(float* a_dev, const void* row_dev, int32_t index_bits, int64_t nnz, const float* scale_dev,
                      int32_t device, void* stream);

/* ProjectionOperator.__call__ on a zero-padded dense block x[L][K] (row-major, one column per entity), the
 * contract of reference projections/base.py:30-36 as used by utils/sparse_utils.py:207-211.  out must not alias x. */
int dualip_project_block(const float* x_dev, float* out_dev, int64_t L, int64_t K, const dualip_proj_class* cls,
                         void* stream);

/* ---- stand-alone CSC operators: the reference's public extension recipe (docs/demo/matching_complex.rst:82-168) is
 * written with them (src/dualip/utils/sparse_utils.py).  The stock objective does not call them (its chain is fused into
 * one kernel); they exist so that user subclasses composed from these operators run on the device through this library.
 * All asynchronous on `stream`; index_bits = width of the row / ccol entries (32 or 64). */
/* left_multiply_sparse (sparse_utils.py:54-85): out[e] = vals[e] * v[row[e]]   (diag(v) @ M on the values) */
int dualip_csc_left_multiply(const float* vals_dev, const void* row_dev, int32_t index_bits, int64_t nnz, const float* v_dev,
                             float* out_dev, int32_t device, void* stream);
/* row_sums_csc (sparse_utils.py:223-243): out[r] = sum of the values stored in row r; out (m floats) is overwritten */
int dualip_csc_row_sums(const float* vals_dev, const void* row_dev, int32_t index_bits, int64_t nnz, int32_t m,
                        float* out_dev, int32_t device, void* stream);
/* The two halves of apply_F_to_columns (sparse_utils.py:133-220): the zero-padded row-major [L x K] block of columns
 * cols[0..K) (NULL: columns 0..K-1), block[i][k] = i-th stored value of column cols[k] or 0, and the write-back of a
 * projected block into a values array (positions of other columns are left untouched). */
int dualip_csc_gather_block(const void* ccol_dev, int32_t index_bits, const float* vals_dev, const int64_t* cols_dev, int64_t K,
                            int64_t L, float* block_dev, int32_t device, void* stream);
int dualip_csc_scatter_block(const void* ccol_dev, int32_t index_bits, const float* block_dev, const int64_t* cols_dev, int64_t K,
                             int64_t L, float* vals_out_dev, int32_t device, void* stream);

/* ---- the demo's fairness-row objective (docs/demo/matching_complex.rst:82-168) as one kernel ----
 * A matching LP whose constraint matrix carries two extra dense rows +-A_fairness (same sparsity pattern as A, values
 * f_dev); lambda, b and grad have n_rows + 2 entries.  desc: the caller's CSC arrays (borrowed, not copied; pad_len /
 * classes / col_class as for dualip_plan_create).  x_out_dev: nnz floats (always written: it is the kernel's scratch).
 * work_dev: dualip_fair_work_bytes(n_rows, n_classes) bytes of device memory, zeroed once by the caller. */
int dualip_fair_calc(const dualip_csc_desc* desc, const float* f_dev, const float* lambda_dev, const float* b_dev,
                     double gamma, float* grad_out_dev, dualip_scalars* scalars_out_dev, float* x_out_dev, float* work_dev,
                     void* stream);
int64_t dualip_fair_work_bytes(int32_t n_rows, int32_t n_classes);

#ifdef __cplusplus
}
#endif
#endif /* DUALIP_B200_H */
