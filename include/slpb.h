/* slpb.h — C ABI of the B200-native interior-point Newton step.
 *
 * This is the drop-in boundary for the hot path of SleipnirGroup/Sleipnir
 * (reference @ 0dacf975). In the reference the path sits behind
 *
 *   slp::interior_point<double>(const InteriorPointMatrixCallbacks<double>&,
 *       std::span<std::function<bool(const IterationInfo<double>&)>>,
 *       const Options&, Eigen::Vector<double, Eigen::Dynamic>& x)
 *       (include/sleipnir/optimization/solver/interior_point.hpp:63, exported
 *        at src/optimization/solver/interior_point.cpp:5-15)
 *
 * whose eight std::function callbacks return host Eigen objects
 * (interior_point_matrix_callbacks.hpp:18-250). A device-resident path cannot
 * sit behind host-matrix callbacks, so the cut is moved one step up: the host
 * side (Problem::solve, problem.hpp:512-668) hands over the flattened
 * expression graphs once, and then drives the IPM loop through the calls
 * below; only O(10) scalars per call come back.
 *
 * Conventions: plain pointers and sizes, no C++/torch types. Every function
 * returns 0 on success or a negative slpb_status; no exception crosses the
 * ABI. The caller owns every buffer it passes (all are HOST pointers unless
 * named dev_*); the library owns all device memory until slpb_destroy. A handle
 * is not thread-safe, but different handles may be used concurrently from
 * different threads (the reference solves different Problems concurrently,
 * multistart.hpp:55); there is no process-global mutable state.
 */
#ifndef SLPB_H_
#define SLPB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct slpb_solver slpb_solver;

enum slpb_status {
  SLPB_OK = 0,
  SLPB_ERR_CUDA = -1,        /* a CUDA runtime call failed (see slpb_last_error) */
  SLPB_ERR_ARGUMENT = -2,    /* inconsistent sizes / null pointers / bad enum  */
  SLPB_ERR_STATE = -3,       /* call order violated (e.g. eval before finalize) */
  SLPB_ERR_UNSUPPORTED = -4, /* graph exceeds what the kernels handle            */
  SLPB_ERR_NO_DEVICE = -5,   /* no CUDA device: the product has no CPU fallback  */
  SLPB_ERR_NCCL = -6
};

/* Opcodes of the node tape; one per Expression subclass of the reference
 * (autodiff/expression.hpp, line of each class in the comment). */
enum slpb_op {
  SLPB_OP_CONST = 0,      /* :576  */
  SLPB_OP_VAR = 1,        /* :594  */
  SLPB_OP_SUB = 2,        /* :444  */
  SLPB_OP_ADD = 3,        /* :481  */
  SLPB_OP_DIV = 4,        /* :616  */
  SLPB_OP_MUL = 5,        /* :656  */
  SLPB_OP_NEG = 6,        /* :696  */
  SLPB_OP_ABS = 7,        /* :773  */
  SLPB_OP_ACOS = 8,       /* :833  */
  SLPB_OP_ASIN = 9,       /* :887  */
  SLPB_OP_ATAN = 10,      /* :942  */
  SLPB_OP_ATAN2 = 11,     /* :996  */
  SLPB_OP_CBRT = 12,      /* :517  */
  SLPB_OP_COS = 13,       /* :1058 */
  SLPB_OP_COSH = 14,      /* :1112 */
  SLPB_OP_ERF = 15,       /* :1166 */
  SLPB_OP_EXP = 16,       /* :1222 */
  SLPB_OP_HYPOT = 17,     /* :1280 */
  SLPB_OP_IS_NONNEG = 18, /* :1351 */
  SLPB_OP_IS_POS = 19,    /* :1384 */
  SLPB_OP_LOG = 20,       /* :1417 */
  SLPB_OP_LOG10 = 21,     /* :1469 */
  SLPB_OP_MAX = 22,       /* :1525 */
  SLPB_OP_MIN = 23,       /* :1594 */
  SLPB_OP_POW = 24,       /* :1667 */
  SLPB_OP_SIGN = 25,      /* :1756 */
  SLPB_OP_SIN = 26,       /* :1805 */
  SLPB_OP_SINH = 27,      /* :1860 */
  SLPB_OP_SQRT = 28,      /* :1915 */
  SLPB_OP_TAN = 29,       /* :1971 */
  SLPB_OP_TANH = 30,      /* :2029 */
  SLPB_OP_COUNT = 31
};

/* The eight outputs of the IPM callback bundle
 * (interior_point_matrix_callbacks.hpp: f :54, g :75, H :110, H_c :145,
 *  c_e :168, A_e :196, c_i :217, A_i :245). */
enum slpb_output {
  SLPB_OUT_F = 0,
  SLPB_OUT_G = 1,
  SLPB_OUT_H_F = 2, /* Hessian of the cost, scaled by d_f inside H */
  SLPB_OUT_H_C = 3, /* Hessian of −yᵀc_e − zᵀc_i                   */
  SLPB_OUT_C_E = 4,
  SLPB_OUT_A_E = 5,
  SLPB_OUT_C_I = 6,
  SLPB_OUT_A_I = 7,
  SLPB_OUT_COUNT = 8
};

/* ---- lifetime ----------------------------------------------------------- */

/* One handle per Problem::solve (problem.hpp:281): owns a CUDA stream and all
 * workspaces. `device` is the CUDA ordinal. */
int slpb_create(int device, slpb_solver** out);
void slpb_destroy(slpb_solver* s);
/* Human-readable text of the last error on this handle (never NULL). */
const char* slpb_last_error(const slpb_solver* s);

/* ---- multi-GPU (optional) -------------------------------------------------
 * One process per GPU. The re-linearisation sweep (slpb_eval_current with
 * derivatives: g, A_e, A_i, H) is sharded over the time steps: every rank
 * evaluates its share of the tasks and ONE ncclAllGather per Newton iteration
 * exchanges the rows the ranks produced; assembly, factorisation and solve then
 * run redundantly (and bit-identically) on every rank, so all ranks take the
 * same decisions without further communication. Every rank must make the same
 * sequence of calls. NCCL is loaded at run time (libnccl.so.2).
 *   slpb_comm_unique_id: rank 0 creates the 128-byte ncclUniqueId, the caller
 *                        distributes it (MPI, torch.distributed, a file …);
 *   slpb_comm_init:      between slpb_create and slpb_finalize. */
int slpb_comm_unique_id(void* id_out_128_bytes);
int slpb_comm_init(slpb_solver* s, int rank, int world,
                   const void* id_128_bytes);
/* Sharded solves run the host loop on every rank in lockstep; a decision taken
 * from rank-local information (the wall-clock TIMEOUT of interior_point.hpp:
 * 860-862, a user callback asking to stop, :414-418) must be taken by all ranks
 * together or the others block in the next all-gather. *any = OR over the
 * ranks of local_flag (one tiny all-gather; world == 1: the flag itself). */
int slpb_comm_agree(slpb_solver* s, int32_t local_flag, int32_t* any);

/* ---- problem upload (once per solve) ------------------------------------ */

/* Flattened, de-duplicated node list in child-before-parent order: the union
 * of the graphs that Gradient/Hessian/Jacobian hold in the reference
 * (jacobian.hpp:159-177, hessian.hpp:160-178). lhs/rhs are node indices or −1.
 * `val` carries constants (ignored for other ops). leaf_x[i] / leaf_y[j] /
 * leaf_z[k] name the VAR nodes of decision variable i, of the y_ad multiplier
 * of equality row j and of the z_ad multiplier of inequality row k
 * (problem.hpp:315,519-520). */
int slpb_upload_tape(slpb_solver* s, int32_t n_nodes, const uint8_t* op,
                     const int32_t* lhs, const int32_t* rhs, const double* val,
                     int32_t n_x, const int32_t* leaf_x, int32_t n_y,
                     const int32_t* leaf_y, int32_t n_z,
                     const int32_t* leaf_z);

/* Row descriptors of one output, mirroring the members the reference keeps per
 * Jacobian/Hessian: per-row topological lists in parent→child order
 * (m_top_lists), (col,node) output lists (m_output_lists), the cached
 * triplets of LINEAR rows (m_cached_triplets) and which rows are re-swept
 * (m_nonlinear_rows). */
typedef struct slpb_rowset {
  int32_t n_rows;
  int32_t n_cols;            /* n for derivative outputs, 1 for value outputs */
  const int32_t* row_ptr;    /* n_rows+1 offsets into row_nodes               */
  const int32_t* row_nodes;  /* top-lists; row_nodes[row_ptr[r]] is the root  */
  const int32_t* out_ptr;    /* n_rows+1 offsets into out_col/out_node        */
  const int32_t* out_col;
  const int32_t* out_node;
  const uint8_t* row_swept;  /* 1: QUADRATIC/NONLINEAR row, swept every eval  */
  int32_t n_cached;          /* triplets of LINEAR rows, evaluated on the host */
  const int32_t* cached_row;
  const int32_t* cached_col;
  const double* cached_val;
} slpb_rowset;

/* For SLPB_OUT_F / C_E / C_I only row_ptr/row_nodes are read (value rows; an
 * empty top-list means "constant": `const_val[r]` is used instead). */
int slpb_upload_rows(slpb_solver* s, int which, const slpb_rowset* rows,
                     const double* const_val);

/* Compiles the uploaded graphs into device programs and fixes the sparsity
 * patterns of g, A_e, A_i, H and of the reduced KKT matrix. Must follow the
 * uploads. */
int slpb_finalize(slpb_solver* s);

/* ProblemScaling (solver/util/problem_scaling.hpp:21-115): d_f and the
 * per-row d_ce[m_e], d_ci[m_i]. Defaults to all ones. */
int slpb_set_scaling(slpb_solver* s, double d_f, const double* d_ce,
                     const double* d_ci);

/* Set to non-zero to make H ignore the H_C output (the behaviour of the
 * reference's feasibility-restoration Hessian callback,
 * feasibility_restoration.hpp:485-495). */
int slpb_set_ignore_constraint_hessian(slpb_solver* s, int ignore);

/* Arithmetic of the Schur-complement updates of the LDLᵀ factorisation.
 * REFERENCE (default): W(i,j) -= l_ik * w_jk with the product and the
 *   difference rounded separately, as Eigen::SimplicialLDLT does in the
 *   reference's x86-64 build (sparse_regularized_ldlt.hpp:74,105): whole solves
 *   follow the reference's regularisation decisions.
 * TENSOR: every term is one fused multiply-add and the rank-4 updates of dense
 *   frontal matrices of order >= 16 run on the FP64 tensor cores (BASELINE
 *   config 3; replaces the dense kernels behind
 *   solver/util/dense_regularized_ldlt.hpp:59-136). One rounding per term away
 *   from the reference; structurally singular pivots no longer cancel to an
 *   exact zero, so the iteration path of a whole solve may differ.
 * May be changed between factorisations; batches and groups created from a
 * solver inherit its mode. The environment variable SLPB_FACTOR_ARITH=tensor
 * makes TENSOR the initial mode of every new solver. */
enum slpb_factor_arithmetic {
  SLPB_ARITH_REFERENCE = 0,
  SLPB_ARITH_TENSOR = 1
};
int slpb_set_factor_arithmetic(slpb_solver* s, int mode);

enum slpb_ordering {
  SLPB_ORDER_NESTED_DISSECTION = 0, /* level-set bisection: log-depth tree     */
  SLPB_ORDER_AMD = 1,               /* approximate minimum degree (Eigen-like) */
  SLPB_ORDER_NATURAL = 2,
  SLPB_ORDER_CUSTOM = 3             /* perm[k] = index eliminated k-th         */
};

typedef struct slpb_symbolic_stats {
  int32_t dim;          /* n + m_e                                     */
  int64_t nnz_kkt;      /* lower-triangle nnz of the reduced KKT        */
  int64_t nnz_l;        /* strictly-lower nnz of L (simplicial count)   */
  int64_t nnz_l_stored; /* doubles stored in the supernodal panels      */
  int32_t n_supernodes;
  int32_t n_levels;     /* depth of the supernodal elimination tree     */
  int32_t max_front;    /* largest frontal matrix dimension             */
  int32_t etree_height; /* height of the column elimination tree        */
} slpb_symbolic_stats;

/* Symbolic analysis of the static KKT pattern (replaces
 * SimplicialLDLT::analyzePattern, sparse_regularized_ldlt.hpp:69-72): ordering,
 * elimination tree, supernodes, level schedule, assembly maps. */
int slpb_analyze(slpb_solver* s, int ordering, const int32_t* perm,
                 slpb_symbolic_stats* stats);
/* Copies out the permutation in use (dim entries). */
int slpb_get_permutation(const slpb_solver* s, int32_t* perm);

/* ---- iterate ------------------------------------------------------------- */

int slpb_set_iterate(slpb_solver* s, const double* x, const double* sl,
                     const double* y, const double* z);
int slpb_get_iterate(slpb_solver* s, double* x, double* sl, double* y,
                     double* z);

/* Finite-ness bits reported by slpb_eval_* . */
enum slpb_finite_bits {
  SLPB_FINITE_F = 1, SLPB_FINITE_C_E = 2, SLPB_FINITE_C_I = 4,
  SLPB_FINITE_G = 8, SLPB_FINITE_A_E = 16, SLPB_FINITE_A_I = 32,
  SLPB_FINITE_H = 64
};

typedef struct slpb_point_info {
  double f;          /* d_f · f(x)                                        */
  double ce_l1;      /* ‖c_e‖₁                                            */
  double cis_l1;     /* ‖c_i − s‖₁                                        */
  double log_s_sum;  /* Σ ln sᵢ (FilterEntry cost term, filter.hpp:53-57)  */
  int32_t finite;    /* OR of slpb_finite_bits that hold                   */
  int32_t ci_all_positive; /* all c_i > 0 (feasible_ipm test, :515)        */
} slpb_point_info;

/* Evaluates the current iterate: f, c_e, c_i always; with derivatives != 0
 * also A_e, A_i, g, H(x, y, z) (the re-linearisation of
 * interior_point.hpp:245-251 and :809-812). */
int slpb_eval_current(slpb_solver* s, int derivatives, slpb_point_info* info);

typedef struct slpb_kkt_stats {
  /* scaled problem (kkt_error.hpp:92-146) */
  double r_inf, r_l1;       /* ‖g − A_eᵀy − A_iᵀz‖                   */
  double y_l1, z_l1;
  double sz_min, sz_max;    /* extremes of sᵢzᵢ: ‖Sz − μe‖∞ for any μ */
  double sz_mu_l1;          /* ‖Sz − μe‖₁ for the μ passed in         */
  double ce_inf, ce_l1, cis_inf, cis_l1;
  /* unscaled problem (kkt_error.hpp:216-251) */
  double u_r_inf, u_y_l1, u_z_l1, u_sz_min, u_sz_max, u_ce_inf, u_cis_inf;
  /* is_locally_infeasible.hpp:17-60 */
  double aetce_l2, ce_l2, aitcip_l2, cip_l2;
  /* divergence guard, interior_point.hpp:405-408 */
  double x_inf, s_inf;
  int32_t xs_finite;
  int32_t pad;
} slpb_kkt_stats;

int slpb_kkt_stats_current(slpb_solver* s, double mu, slpb_kkt_stats* out);
/* Same quantities for the trial point, using the trial point's own g/A_e/A_i
 * (the α<α_min fallback, interior_point.hpp:692-706); evaluates them first. */
int slpb_kkt_stats_trial(slpb_solver* s, double mu, slpb_kkt_stats* out);
/* slpb_accept, slpb_eval_current(derivatives = 2) and slpb_kkt_stats_current in
 * ONE host round trip — the tail of an accepted iteration,
 * interior_point.hpp:779-832: x, s, y, z ← trial (z clamped), g, A_e, A_i, H
 * re-evaluated there, then the error reductions. *finite receives the
 * SLPB_FINITE_G/A_E/A_I/H bits. */
int slpb_accept_relinearize(slpb_solver* s, double mu, int32_t* finite,
                            slpb_kkt_stats* out);

/* ---- Newton system -------------------------------------------------------- */

typedef struct slpb_factor_info {
  int32_t n_pos, n_neg, n_zero; /* Inertia of D (inertia.hpp, ±DBL_EPSILON)    */
  int32_t zero_pivot;           /* a pivot was exactly 0 (Eigen NumericalIssue) */
  double min_abs_d;
} slpb_factor_info;

/* Assembles lhs = [H + tril(A_iᵀΣA_i); A_e] + diag(δ…δ, −γ…−γ)
 * (interior_point.hpp:426-440, sparse_regularized_ldlt.hpp:217-224) and factors
 * it P·lhs·Pᵀ = L·D·Lᵀ. With reassemble == 0 the previously assembled values
 * are re-used and only δ, γ change (the retry loop :104-151). */
int slpb_factor(slpb_solver* s, double delta, double gamma, int reassemble,
                slpb_factor_info* info);

/* Speculative pair: factors lhs + diag(δ₀, −γ₀) and lhs + diag(δ₁, −γ₁) in ONE
 * launch (two independent numeric factorisations over the same symbolic
 * structure). The reference's retry loop (sparse_regularized_ldlt.hpp:64-104)
 * always tries (0, 0) first and knows its second candidate (δ, γ_min) before
 * the first attempt returns, so both can run side by side; the host then keeps
 * the one the sequential algorithm would have kept (slpb_select_factor) and the
 * decision sequence is unchanged. info[v] describes variant v. */
int slpb_factor_pair(slpb_solver* s, const double delta[2],
                     const double gamma[2], int reassemble,
                     slpb_factor_info info[2]);
/* Chooses which variant (0 or 1) of the last factorisation slpb_solve /
 * slpb_soc_iterate / SLPB_ARR_D use. slpb_factor always selects 0. */
int slpb_select_factor(slpb_solver* s, int which);

typedef struct slpb_step_info {
  double alpha_max;  /* fraction-to-the-boundary on (s, p_s)            */
  double alpha_z;    /* fraction-to-the-boundary on (z, p_z)            */
  double g_dot_px;   /* gᵀpˣ                                            */
  double sinv_dot_ps;/* (S⁻¹e)ᵀpˢ ;  D_ϕ = g_dot_px − μ·sinv_dot_ps     */
  double px_inf, ps_inf, py_inf, pz_inf;
  int32_t finite;
  int32_t pad;
} slpb_step_info;

/* Optional: announces the barrier parameter of the next slpb_solve BEFORE the
 * factorisation. The right-hand side (interior_point.hpp:444-448) does not
 * depend on δ, γ, so it is built now and the following slpb_factor /
 * slpb_factor_pair carries its forward substitution through the elimination
 * (one extra column per front); slpb_solve(mu, …) with the same μ then only runs
 * the backward pass. Same arithmetic, same bits as the separate forward pass.
 * Any call that changes the iterate or the derivatives cancels it. */
int slpb_prepare_rhs(slpb_solver* s, double mu);

/* rhs (interior_point.hpp:444-448), p = lhs⁻¹ rhs, step recovery (:470-481),
 * both fraction-to-the-boundary rules (:488,497) and the pieces of D_ϕ (:508). */
int slpb_solve(slpb_solver* s, double mu, double tau, slpb_step_info* info);

/* Second-order correction (interior_point.hpp:561-664). begin: copies the
 * step and seeds c_e^soc, (c_i − s)^soc. iterate: accumulates with α_soc and
 * the CURRENT trial point, rebuilds the rhs (:611-616), solves into the SOC
 * step and returns its α's. */
int slpb_soc_begin(slpb_solver* s);
int slpb_soc_iterate(slpb_solver* s, double mu, double tau, double alpha_soc,
                     slpb_step_info* info);

/* trial = iterate + α·(p_x, p_s), + α_z·(p_y, p_z); evaluates f, c_e, c_i there
 * (interior_point.hpp:513-528, 626-633). which_step: 0 Newton step, 1 SOC
 * step. slack_from_ci != 0 sets trial_s = trial_c_i (feasible_ipm, :515-520). */
int slpb_trial(slpb_solver* s, double alpha, double alpha_z, int which_step,
               int slack_from_ci, slpb_point_info* info);

/* slpb_solve followed by slpb_trial at the full step (α = α_max; the dual step
 * α_z, or α_max too when dual_uses_primal_alpha != 0 — the SQP rule, sqp.hpp:346)
 * in ONE host round trip: the trial-point kernel reads the step lengths the
 * step reductions just left on the device (interior_point.hpp:470-528). The
 * caller ignores *trial when α_max sends it to feasibility restoration. */
int slpb_solve_trial(slpb_solver* s, double mu, double tau,
                     int dual_uses_primal_alpha, int slack_from_ci,
                     slpb_step_info* step, slpb_point_info* trial);

/* Evaluates f, c_e, c_i at a HOST-supplied point (x, s) without touching the
 * current iterate (it lands in the trial buffers): the acceptance test of
 * feasibility restoration (interior_point.hpp:738-756) probes the original
 * problem at the restoration iterate this way. */
int slpb_probe_point(slpb_solver* s, const double* x, const double* sl,
                     slpb_point_info* info);

/* Least-squares multiplier estimate at the current iterate
 * (lagrange_multiplier_estimate.hpp:55-131): y, z ← argmin ‖Âᵀ[y; z] − [∇f; −μe]‖,
 * Â = [A_e 0; A_i −S], z clamped to [μ/(κs), κμ/s]. Needs g, A_e, A_i of the
 * current iterate (slpb_eval_current with derivatives) and slpb_analyze. Solved
 * as the equivalent system with the Newton matrix's pattern (H → I, Σ → S⁻²),
 * so it reuses the symbolic factorisation; overwrites the assembled lhs and
 * the factor. info reports the inertia of that factorisation. */
int slpb_multiplier_estimate(slpb_solver* s, double mu, slpb_factor_info* info);

/* Commits the trial point: x,s,y,z,f,c_e,c_i ← trial, clamps z to
 * [μ/(κs), κμ/s], κ = 1e10 (interior_point.hpp:779-805). */
int slpb_accept(slpb_solver* s, double mu);

/* ---- readback (callbacks, final write-back, parity tests) ---------------- */

enum slpb_array {
  SLPB_ARR_X = 0, SLPB_ARR_S, SLPB_ARR_Y, SLPB_ARR_Z,
  SLPB_ARR_G,       /* dense, n                                   */
  SLPB_ARR_C_E, SLPB_ARR_C_I,
  SLPB_ARR_A_E_VAL, SLPB_ARR_A_I_VAL, SLPB_ARR_H_VAL, /* CSC values        */
  SLPB_ARR_KKT_VAL, /* lower CSC values of lhs (without δ, γ)     */
  SLPB_ARR_D,       /* diagonal of the factor, elimination order  */
  SLPB_ARR_RHS, SLPB_ARR_P_X, SLPB_ARR_P_S, SLPB_ARR_P_Y, SLPB_ARR_P_Z,
  SLPB_ARR_TRIAL_X, SLPB_ARR_TRIAL_S, SLPB_ARR_TRIAL_Y, SLPB_ARR_TRIAL_Z,
  SLPB_ARR_TRIAL_C_E, SLPB_ARR_TRIAL_C_I
};
int slpb_array_size(const slpb_solver* s, int which, int64_t* count);
int slpb_download(slpb_solver* s, int which, double* dst);

/* CSC patterns (Eigen::SparseMatrix<double, ColMajor, int> layout). which:
 * SLPB_OUT_A_E, SLPB_OUT_A_I, SLPB_OUT_H_C (= pattern of H, lower triangle) or
 * −1 for the reduced KKT (lower triangle, dim n+m_e). colptr may be NULL to
 * query nnz only. */
int slpb_pattern(const slpb_solver* s, int which, int32_t* rows, int32_t* cols,
                 int64_t* nnz, int32_t* colptr, int32_t* rowidx);

/* ---- many instances over one symbolic structure ---------------------------
 * slp::multistart (optimization/multistart.hpp:44-73) solves one problem from
 * many initial guesses: every start has the lhs pattern of
 * interior_point.hpp:426-440 and differs only in values. A batch factors and
 * solves `batch` such systems side by side — lane = instance, SoA storage
 * [entry][32] per group of 32 instances, one walk over the assembly tree for
 * all of them — with the arithmetic of slpb_factor / slpb_solve per instance
 * (bit-identical D and solutions). It borrows the symbolic analysis, the
 * stream and the device of `s`, which must outlive it. */
typedef struct slpb_batch slpb_batch;
int slpb_batch_create(slpb_solver* s, int32_t batch, slpb_batch** out);
void slpb_batch_destroy(slpb_batch* b);
int slpb_batch_size(const slpb_batch* b, int32_t* batch, int32_t* groups);
/* Instance i <- host arrays: lhs values in slpb_pattern(-1) order (without
 * delta/gamma) and/or a right-hand side of length n + m_e (either may be NULL). */
int slpb_batch_set_system(slpb_batch* b, int32_t instance,
                          const double* kkt_val, const double* rhs);
/* Instance i <- the assembled lhs values and the rhs resident in `src` (any
 * solver with the same KKT pattern on the same device, e.g. one start of a
 * multistart after slpb_factor / slpb_prepare_rhs): device-to-device. */
int slpb_batch_capture(slpb_batch* b, int32_t instance, slpb_solver* src);
/* Factors lhs_i + diag(delta_i, -gamma_i) for every instance in ONE launch
 * (RegularizedLDLT::compute's numeric step, sparse_regularized_ldlt.hpp:74,105);
 * info[i] as slpb_factor reports it. */
int slpb_batch_factor(slpb_batch* b, const double* delta, const double* gamma,
                      slpb_factor_info* info);
/* Forward and backward substitution of every instance's rhs in ONE launch
 * (sparse_regularized_ldlt.hpp:159-161). */
int slpb_batch_solve(slpb_batch* b);
enum slpb_batch_array { SLPB_BATCH_SOLUTION = 0, SLPB_BATCH_D = 1 };
/* Copies out instance i's solution (original order, n + m_e) or D
 * (elimination order). */
int slpb_batch_get(slpb_batch* b, int32_t instance, int what, double* dst);
/* Device time of the last slpb_batch_factor / slpb_batch_solve launch (CUDA
 * events on the stream, milliseconds). */
int slpb_batch_last_ms(const slpb_batch* b, float* factor_ms, float* solve_ms);
/* Doubles stored per instance: packed panels (= nnz(L) incl. supernodal
 * padding + dim) and packed update matrices. */
int slpb_batch_bytes(const slpb_batch* b, int64_t* panel_entries,
                     int64_t* update_entries);

/* ---- dynamic batching of concurrent solves (slp::multistart) ---------------
 * Starts of a multistart run on their own threads with their own handles; what
 * they share is the lhs pattern. Members of a group hand slpb_factor /
 * slpb_factor_pair / slpb_solve* / slpb_soc_iterate / slpb_multiplier_estimate
 * to ONE batched launch per round (slpb_batch_*): a call parks until every
 * active member has parked one or has left, then the last arriver runs the
 * round for all. Results per member are bit-identical to the ungrouped solve. */
typedef struct slpb_group slpb_group;
int slpb_group_create(int device, int32_t expected_members, slpb_group** out);
void slpb_group_destroy(slpb_group* g);
/* After slpb_analyze; blocks until all expected members have joined or been
 * abandoned. The first member's pattern and elimination order become the
 * group's; a solver with another pattern is refused (SLPB_ERR_ARGUMENT, counted
 * as abandoned) and carries on alone. */
int slpb_group_join(slpb_group* g, slpb_solver* s);
/* An expected member will not join (its start failed earlier, or it is not
 * an interior-point problem of this shape). */
int slpb_group_abandon(slpb_group* g);
/* The member stops / resumes taking part in the rounds (e.g. around feasibility
 * restoration, which runs on another handle for many iterations). */
int slpb_group_pause(slpb_solver* s);
int slpb_group_resume(slpb_solver* s);
/* Before slpb_destroy of a member. */
int slpb_group_leave(slpb_solver* s);
int slpb_group_stats(slpb_group* g, int64_t* rounds, int64_t* requests);

/* ---- instrumentation ------------------------------------------------------ */

typedef struct slpb_counters {
  int64_t kernel_launches;   /* kernels of this library launched so far    */
  int64_t factorizations, solves, evals_full, evals_values;
  int64_t tape_nodes, program_bytes, n_clusters, n_program_classes;
  int64_t h2d_bytes, d2h_bytes;
  /* numeric factorisations that ran through all fronts (a variant that meets
   * an exactly zero pivot is abandoned and not counted here) */
  int64_t factorizations_completed;
} slpb_counters;
int slpb_get_counters(const slpb_solver* s, slpb_counters* out);

/* Collectives of a sharded solve (slpb_comm_init with world > 1), per kind:
 * [0] derivative rows of the re-linearisation sweep, [1] roots of the locally
 * eliminated subtrees (update matrices, update vectors, inertia counts),
 * [2] solution pieces. bytes = payload this rank contributes per call × calls. */
typedef struct slpb_comm_stats {
  int64_t count[3];
  int64_t bytes[3];
  double total_ms[3];   /* pack + all-gather + unpack, CUDA events */
} slpb_comm_stats;
int slpb_get_comm_stats(slpb_solver* s, slpb_comm_stats* out);

/* Device time per phase since the handle was created, measured with CUDA
 * events on the handle's stream (milliseconds): [0] eval(full) [1] eval(values)
 * [2] assemble [3] factor [4] solve. Timing events disturb back-to-back
 * kernels, so the phases are SAMPLED (one run in SLPB_TIMER_EVERY, default 8;
 * SLPB_NO_TIMERS=1 switches them off): total_ms and count cover the sampled
 * runs (total_ms / count = mean device time of a run), launches counts every
 * run of the phase. */
typedef struct slpb_timers {
  double total_ms[5];
  int64_t count[5];
  int64_t launches[5];
} slpb_timers;
int slpb_get_timers(slpb_solver* s, slpb_timers* out);

/* Device time of the most recent SAMPLED factor / solve / eval (see
 * slpb_timers; SLPB_TIMER_EVERY=1 times every run), measured with CUDA events
 * on the handle's stream (milliseconds; 0 if not yet run). which: 0 eval(full),
 * 1 eval(values), 2 assemble, 3 factor, 4 solve. */
int slpb_last_device_ms(slpb_solver* s, int which, float* ms);
/* Benchmark hygiene: overwrites a 256 MiB scratch buffer (larger than the
 * 126 MB L2) on the handle's stream and waits, so that the next phase starts
 * with a cold L2. Not used by the solver itself. */
int slpb_flush_l2(slpb_solver* s);
/* The handle's CUDA stream (cudaStream_t as void*), for external timing. */
void* slpb_stream(slpb_solver* s);

#ifdef __cplusplus
}
#endif
#endif /* SLPB_H_ */
