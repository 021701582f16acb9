/* sdpcuda.h — C ABI of the B200-native SDP relaxation solver (libsdpcuda.so).
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no CUDA/torch types.  It is what the
 * SCIP-SDP side binding `sdpisolver_cuda.c` (our fifth implementation of src/sdpi/sdpisolver.h next to
 * sdpisolver_{dsdp.c,sdpa.cpp,mosek.c,none.c}) calls where the reference bindings call the vendor APIs:
 *   DSDPCreate/DSDPSetup/DSDPSolve/DSDPComputeX      sdpisolver_dsdp.c:991-1004,1489-1518
 *   SDPA::inputElement/initializeSolve/solve          sdpisolver_sdpa.cpp:981,1132-1412,1600-1670
 *   SDPA::getResultXVec/YMat/XMat, getPhaseValue      sdpisolver_sdpa.cpp:1906-3125
 * and what `lapack_cuda.c` calls where src/sdpi/lapack_interface.c calls DSYEVR (lapack_interface.c:178-603).
 *
 * Problem form handed over (SCIP-SDP's "dual", sdpisolver.h:34-43, after the binding has removed fixed variables,
 * empty rows/cols/blocks, split LP rows into one-sided rows and turned variable bounds into LP rows exactly like
 * sdpisolver_sdpa.cpp:1015-1412):
 *
 *      min  obj' y
 *      s.t. S^(k) = sum_j y_j A_j^(k) - C^(k)  >= 0 (psd)      k = 0..nblocks-1
 *           s     = D y - d                    >= 0            (nlp one-sided rows, CSR)
 *
 * with multipliers  X^(k) >= 0 (psd), x >= 0 :  sum_k A_j^(k).X^(k) + (D'x)_j = obj_j ,  max  sum_k C^(k).X^(k) + d'x.
 * All matrices are given by their lower triangle (row >= col), 0-based, doubles; indices are 32-bit ints.
 */
#ifndef SDPCUDA_H
#define SDPCUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDPCUDA_ABI_VERSION 5

/* return codes of every entry point */
#define SDPCUDA_OK            0
#define SDPCUDA_ERR_ARG       1   /* invalid argument / inconsistent problem */
#define SDPCUDA_ERR_NOMEM     2   /* host or device allocation failed */
#define SDPCUDA_ERR_CUDA      3   /* CUDA runtime error (no device, launch failure); never a silent CPU fallback */
#define SDPCUDA_ERR_STATE     4   /* result requested before a solve */

/* Solver phase after a solve, in SCIP-SDP naming (p = X-problem, d = y-problem), mirroring SDPA::PhaseType as it is
 * interpreted by sdpisolver_sdpa.cpp:1906-2317. */
typedef enum sdpcuda_phase
{
   SDPCUDA_NOINFO     = 0,  /* iteration/time limit or numerical trouble, neither side feasible */
   SDPCUDA_PFEAS      = 1,  /* X-side feasible only, not converged */
   SDPCUDA_DFEAS      = 2,  /* y-side feasible only, not converged */
   SDPCUDA_PDFEAS     = 3,  /* both feasible, gap not closed */
   SDPCUDA_PDINF      = 4,  /* both sides look infeasible */
   SDPCUDA_PFEAS_DINF = 5,  /* y-problem infeasible (certificate: X-ray) -> IsDualInfeasible / IsPrimalUnbounded */
   SDPCUDA_PINF_DFEAS = 6,  /* y-problem unbounded (certificate: y-ray) -> IsDualUnbounded / IsPrimalInfeasible */
   SDPCUDA_PDOPT      = 7,  /* converged to tolerances */
   SDPCUDA_PUNBD      = 8,  /* stopped: objective limit exceeded (IsObjlimExc) */
   SDPCUDA_DUNBD      = 9,  /* stopped: y-objective fell below the lower cut */
   SDPCUDA_DINF       = 10  /* y-problem infeasible by an X-ray, but the X-side itself is not (yet) feasible: IsDualInfeasible
                             * holds, primal feasibility is not claimed (both-infeasible case of checksdpi.c:656) */
} sdpcuda_phase;

/* why the iteration stopped (GetInternalStatus, sdpisolver.h:439-450) */
typedef enum sdpcuda_stop
{
   SDPCUDA_STOP_CONVERGED = 0,
   SDPCUDA_STOP_INFEASCERT= 1,
   SDPCUDA_STOP_NUMERICS  = 2,
   SDPCUDA_STOP_OBJLIMIT  = 3,
   SDPCUDA_STOP_ITERLIMIT = 4,
   SDPCUDA_STOP_TIMELIMIT = 5
} sdpcuda_stop;

typedef struct sdpcuda_problem
{
   int            m;           /* number of (active) variables y */
   const double*  obj;         /* [m] */
   int            nblocks;     /* SDP blocks */
   const int*     blocksizes;  /* [nblocks] */
   /* entries of the A_j^(k): variable j owns [varbeg[j], varbeg[j+1]) */
   const int*     varbeg;      /* [m+1] */
   const int*     entblk;      /* [nnz] */
   const int*     entrow;      /* [nnz] row >= col */
   const int*     entcol;      /* [nnz] */
   const double*  entval;      /* [nnz] */
   /* constant matrices C^(k) (the reference's A_0 after fixings) */
   int            cnnz;
   const int*     cblk;
   const int*     crow;
   const int*     ccol;
   const double*  cval;
   /* LP block: row l is  sum_{p in [lpbeg[l],lpbeg[l+1])} lpval[p]*y[lpind[p]] - lprhs[l] >= 0 */
   int            nlp;
   const int*     lpbeg;       /* [nlp+1] */
   const int*     lpind;
   const double*  lpval;
   const double*  lprhs;       /* [nlp] */
} sdpcuda_problem;

typedef struct sdpcuda_params
{
   double gaptol;        /* relative duality-gap tolerance (SCIP_SDPPAR_GAPTOL, handed on like setParameterEpsilonStar) */
   double feastol;       /* primal/dual feasibility tolerance (SCIP_SDPPAR_SDPSOLVERFEASTOL, setParameterEpsilonDash) */
   double objlimit;      /* stop when the lower bound (X-objective) exceeds this; >= 1e20 = off (SCIP_SDPPAR_OBJLIMIT) */
   double lambdastar;    /* scale of the initial point X = S = lambdastar*I; <= 0: computed from the data */
   double timelimit;     /* seconds; <= 0 or >= 1e20 = none.  Checked between the iterations of the multi-kernel path; the one-launch
                          * kernels (relaxations inside the single-CTA limits, frontier batches, packed solves) run to completion -
                          * milliseconds, bounded by maxiter - and never report SDPCUDA_STOP_TIMELIMIT */
   double absgaptol;     /* additionally require |pobj - dobj| <= absgaptol (the binding's post-check, sdpisolver_sdpa.cpp:449-451); <= 0 = off */
   int    maxiter;       /* <= 0: default (100) */
   int    setting;       /* 1 fast, 2 medium, 3 stable step-length/centering rules (SCIP_SDPSOLVERSETTING) */
   int    verbose;       /* iteration log to stdout (SCIP_SDPPAR_SDPINFO) */
   int    reserved;
   double preoptgap;     /* > 0: keep a copy of the first iterate whose relative gap and scaled infeasibilities are <= preoptgap
                            (SCIP_SDPPAR_WARMSTARTPOGAP; sdpisolver_sdpa.cpp:1611-1653); <= 0 = off */
} sdpcuda_params;

typedef struct sdpcuda_result
{
   int    phase;         /* sdpcuda_phase */
   int    stop;          /* sdpcuda_stop */
   int    iterations;
   int    launches;      /* CUDA kernel launches issued by this solve (0 for the CPU oracle) */
   double pobj;          /* X-side objective  sum C.X + d'x   (lower bound for the min problem) */
   double dobj;          /* y-side objective  obj'y */
   double relgap;
   double pinf;          /* scaled primal residual  ||obj - A(X) - D'x|| / (1+||obj||) */
   double dinf;          /* scaled dual residual */
   double mu;
   double seconds;       /* wall time of the solve including host<->device transfers */
   double device_ms;     /* device time between the first and last kernel of the solve (CUDA events) */
   double h2d_bytes;     /* host->device bytes moved by this call (problem upload + initial point) */
   double d2h_bytes;     /* device->host bytes moved by this call (per-iteration scalars; solution getters add their own) */
} sdpcuda_result;

typedef struct sdpcuda_handle sdpcuda_handle;

/* ---- life cycle (SCIPsdpiSolverCreate/Free, sdpisolver.h:126-137) ---- */
int  sdpcuda_abi_version(void);
const char* sdpcuda_backend_name(void);           /* "cuda-sm_100a" for the product library */
int  sdpcuda_create(sdpcuda_handle** h, int device /* -1: round-robin over visible devices, one stream per handle */);
int  sdpcuda_destroy(sdpcuda_handle* h);
void sdpcuda_default_params(sdpcuda_params* p);

/* ---- solve (SCIPsdpiSolverLoadAndSolveWithPenalty, sdpisolver.h:258-322) ----
 * Copies the problem to the device, runs the primal-dual predictor-corrector iteration there and keeps the
 * solution device-resident until the getters below fetch it.  start_y may be NULL. Blocking. */
int  sdpcuda_solve(sdpcuda_handle* h, const sdpcuda_problem* prob, const sdpcuda_params* par,
                   const double* start_y, sdpcuda_result* res);

/* ---- warm start (start point of SCIPsdpiSolverLoadAndSolve, sdpisolver.h:160-175; use in sdpisolver_sdpa.cpp:1481-1600) ----
 * Dense start matrices for the NEXT sdpcuda_solve on this handle (one shot, used together with its start_y):
 * which = 0: X_block (the reference's "startX", multiplier of the LMI), which = 1: S_block (its "startZ", the slack
 * sum A_j y_j - C).  A: n x n, full symmetric.  Both must be positive definite; if the first factorisation fails the
 * solve silently restarts from the default point.  All blocks and the LP part must be given, otherwise the point is ignored. */
int  sdpcuda_set_start_block(sdpcuda_handle* h, int which, int block, int n, const double* A);
int  sdpcuda_set_start_lp(sdpcuda_handle* h, int nlp, const double* xlp, const double* slp);
/* preoptimal point saved by the last solve when params.preoptgap > 0 (SCIPsdpiSolverGetPreoptimalSol, sdpisolver.h:492-515):
 * returns 1 / 0 in *exists; y [m], xlp [nlp] may be NULL; the X blocks come from sdpcuda_get_preopt_X (row-major, full) */
int  sdpcuda_get_preopt(sdpcuda_handle* h, int* exists, double* y, double* xlp);
int  sdpcuda_get_preopt_X(sdpcuda_handle* h, int block, double* X);

/* ---- one large SDP over several GPUs: one process and one handle per GPU, all ranks call sdpcuda_solve with
 * the SAME problem; every rank forms its share of the Schur complement (column strips of the entry path, chunks of the dense
 * path), one NCCL all-reduce over NVLink adds the disjoint shares, the remaining iteration runs replicated and bit-identical.
 * sdpcuda_dist_unique_id: called on rank 0, the 128 bytes travel to the other ranks over the caller's own channel
 * (torch.distributed broadcast in bench.py).  NCCL is bound with dlopen at the first call; without it these return ERR_STATE. */
int  sdpcuda_dist_unique_id(void* id128);
int  sdpcuda_dist_init(sdpcuda_handle* h, int nranks, int rank, const void* id128);
int  sdpcuda_dist_finalize(sdpcuda_handle* h);

/* Same iteration on the problem that the last sdpcuda_solve left resident in HBM (no host->device traffic); used to
 * measure the device-only throughput and for repeated solves with changed tolerances. */
int  sdpcuda_solve_resident(sdpcuda_handle* h, const sdpcuda_params* par, sdpcuda_result* res);

/* Reductions on the primal solution X that is resident on the device (computeConflictCut, relax_sdp.c:1030-1099, pulls
 * the dense X of every block to the host after every node, forms <A_j, X> and <A_0, X> there and calls LAPACK for lambda_min(X)):
 *   out[g] = sum over entries e in [groupbeg[g], groupbeg[g+1]) of  w_e val[e] X_{blk[e]}[row[e], col[e]],  w_e = 1 (row == col) or 2,
 * entries in the device's block numbering and reduced indices (row >= col); 12 bytes per entry go up, 8 bytes per group come back. */
int  sdpcuda_primal_products(sdpcuda_handle* h, int ngroups, const int* groupbeg, const int* blk, const int* row, const int* col,
   const double* val, double* out);
/* certified lower bound of min(lambda_min(X_block), 0): 0 when the Cholesky factorisation of X_block succeeds on the device (always
 * the case for an interior iterate), otherwise -sigma for the smallest sigma = 1e-14 |X|_max 10^k with X + sigma I positive definite */
int  sdpcuda_primal_mineig_bound(sdpcuda_handle* h, int block, double* bound);

/* Re-solve with a problem of the SAME STRUCTURE as the resident one (the >= 3 solves of a failing node in
 * sdpi.c:3437-3619 - tolerance tightening, FAST/MEDIUM/STABLE ladder, growing penalty parameter Gamma - and the post-check loop of
 * sdpisolver_sdpa.cpp:368-494 re-load everything although only tolerances, settings, the objective coefficient of r and right-hand
 * sides change).  The caller guarantees that P differs from the problem of the last sdpcuda_solve on this handle at most in `obj`
 * and `lprhs` (the binding compares a hash of all other arrays); only these two vectors travel to the device (8 (m + nlp) bytes),
 * the start point is staged like for sdpcuda_solve.  Without a resident problem (or after a packed/batched solve) it is a full
 * sdpcuda_solve. */
int  sdpcuda_solve_patched(sdpcuda_handle* h, const sdpcuda_problem* P, const sdpcuda_params* par, const double* start_y, sdpcuda_result* res);

/* ---- a frontier of independent node relaxations in one call (the open B&B nodes that SCIP-SDP's concurrent
 * solver threads would each hand to SCIPsdpiSolverLoadAndSolve, sdpi.c:3399) ----
 * Every relaxation that fits the single-CTA kernel (blocks of order <= 64, m <= 256, <= 16 blocks) is packed into one host image;
 * the device sees ONE host->device copy, ONE kernel launch (one CTA = one SM per node, cold start set up by the CTA itself) and
 * ONE device->host copy of the results and the y vectors for the whole batch.  Larger relaxations are solved one after the other
 * by sdpcuda_solve on the same handle.  res: [count] or NULL; y_out: NULL or [count] pointers (each NULL or room for m_i doubles).
 * The getters of the handle do not refer to batched nodes afterwards (the batch has its own device buffers).
 * device_ms / seconds of the batched nodes are those of the whole batch (the nodes run side by side); launches = 1 on the first. */
int  sdpcuda_solve_batch(sdpcuda_handle* h, int count, const sdpcuda_problem* const* probs, const sdpcuda_params* par,
                         sdpcuda_result* res, double* const* y_out,
                         const double* objlimits /* NULL, or [count]: per-node value for params.objlimit — the cutoff bound of the
                                                    node minus its fixed-variable objective (relaxing/SDP/objlimit, relax_sdp.c:4265) */);

/* ---- a frontier given as NODES of one mixed-integer SDP (the form SCIP-SDP's tree has them: one model, per node a pair of bound
 * vectors).  The model is SCIP-SDP's "dual" form (sdpi.h: min obj'y, sum_j A_j y_j - A_0 psd, lhs <= D y <= rhs, lb <= y <= ub):
 * entries (variable or -1 for A_0, block, row >= col, value) in the caller's order, rows in CSR.  sdpcuda_solve_nodes does for every
 * node what SCIPsdpiSolve does before the solver sees it (sdpi.c:3123-3399: rows without active variables decide or vanish, rows with
 * one active variable tighten its bounds, fixed variables move into the constant part, empty block rows/columns are removed; csrc/
 * node_marshal.hpp) and hands all remaining relaxations to sdpcuda_solve_batch.
 * lb, ub: [count*nvars].  status[i]: 0 = solved (res[i], bound[i] = objective incl. the fixed variables), 1 = infeasible by the
 * presolve, 2 = all variables fixed and feasible (bound[i] = its value).  y: NULL or [count*nvars] solution in MODEL variables;
 * lbout/ubout: NULL or the tightened bounds; cutoff: NULL or [count] objective cutoffs in model terms (see objlimits above). ---- */
typedef struct sdpcuda_model sdpcuda_model;
int  sdpcuda_model_create(sdpcuda_model** model, int nvars, const double* obj, int nblocks, const int* blocksizes, int nnz,
                          const int* entvar, const int* entblk, const int* entrow, const int* entcol, const double* entval,
                          int nrows, const int* rowbeg, const int* rowind, const double* rowval, const double* lhs, const double* rhs);
int  sdpcuda_model_destroy(sdpcuda_model* model);
int  sdpcuda_solve_nodes(sdpcuda_handle* h, const sdpcuda_model* model, int count, const double* lb, const double* ub,
                         const sdpcuda_params* par, const double* cutoff, int* status, sdpcuda_result* res, double* bound, double* y,
                         double* lbout, double* ubout);
/* test hook: the solver-form problem of one node as flat arrays (ibuf: blocksizes, varbeg, entblk, entrow, entcol, cblk, crow, ccol,
 * lpbeg, lpind, active; dbuf: obj, entval, cval, lpval, lprhs; sizes: m, nblocks, nnz, cnnz, nlp, lpnnz) */
int  sdpcuda_debug_node_problem(const sdpcuda_model* model, const double* lb, const double* ub, double feastol, int* status,
                                double* fixedobj, int* sizes, int* ibuf, size_t icap, double* dbuf, size_t dcap, double* lbout, double* ubout);

/* test hook (no device needed): packs ONE node exactly like sdpcuda_solve_batch does and returns the host image of its read-only
 * data, the length of its work space and its kernel descriptor bound to the given (fake) device addresses; *fits = 0 when the
 * relaxation is outside the single-CTA limits.  tests/test_batch_pack.py re-derives the operators from these arrays. */
int  sdpcuda_debug_pack_node(const sdpcuda_problem* prob, const sdpcuda_params* par, unsigned long long img_base,
                             unsigned long long work_base, unsigned long long y_base, unsigned char* image, size_t image_cap,
                             size_t* image_bytes, size_t* work_doubles, void* descriptor, size_t desc_cap, size_t* desc_bytes, int* fits);

/* the same for a whole batch: everything sdpcuda_solve_batch decides on the host (which problems are batched, one image, one work
 * space, one y buffer, descriptor order with the 256-thread relaxations first) for the given (fake or host) base addresses.
 * descriptors: nbatched kernel descriptors in launch order; result k (at res_base + k) belongs to input problem
 * problem_of_result[k], its y starts yoff_of_result[k] doubles behind y_base. */
int  sdpcuda_debug_pack_batch(int count, const sdpcuda_problem* const* probs, const sdpcuda_params* par, int flags /* 1: SDPCUDA_BATCH_TINY, 2: SDPCUDA_BATCH_SMEM */,
                              unsigned long long img_base, unsigned long long work_base, unsigned long long y_base,
                              unsigned long long res_base, unsigned char* image, size_t image_cap, size_t* image_bytes,
                              size_t* work_doubles, size_t* y_doubles, void* descriptors, size_t desc_cap, int* nbatched, int* ntiny,
                              int* problem_of_result, size_t* yoff_of_result,
                              size_t* stage_bytes /* [2] or NULL: shared memory of the two launches on top of the kernels' own */);

/* Per-kernel-class device timing of the NEXT solve (CUDA events around every launch of the class on the handle's
 * stream; adds a little overhead, so it is off by default).  After the solve sdpcuda_get_profile fills, for each class
 * c < SDPCUDA_NPROF, out[3*c+0] = launches, out[3*c+1] = device milliseconds, out[3*c+2] = algorithmic flops or bytes. */
#define SDPCUDA_NPROF 7
#define SDPCUDA_PROF_GEMM   0   /* FP64 DMMA GEMM, 64 x 64 CTA tiles (flops) */
#define SDPCUDA_PROF_GEMM_SMALL 6 /* the 32 x 32-tile instantiation used for the small products of the blocked factorisations (flops) */
#define SDPCUDA_PROF_DIAG   1   /* 64 x 64 diagonal-block Cholesky + inverse (flops) */
#define SDPCUDA_PROF_SCHUR  2   /* Schur complement assembly, entry/gather path + LP block (bytes) */
#define SDPCUDA_PROF_EIG    3   /* Jacobi / Lanczos step-length kernels (bytes) */
#define SDPCUDA_PROF_TRSV   4   /* blocked triangular solves with M (bytes) */
#define SDPCUDA_PROF_ELEM   5   /* element-wise / reduction sweeps over the arena (bytes) */
int  sdpcuda_set_profiling(sdpcuda_handle* h, int on);
int  sdpcuda_get_profile(sdpcuda_handle* h, double* out /* [3*SDPCUDA_NPROF] */);

/* ---- solution access (GetDualSol/GetPrimal*, sdpisolver.h:484-578) ---- */
int  sdpcuda_get_y(sdpcuda_handle* h, double* y /* [m] */);
int  sdpcuda_get_X(sdpcuda_handle* h, int block, double* X /* [n*n] full symmetric, row-major */);
int  sdpcuda_get_S(sdpcuda_handle* h, int block, double* S /* [n*n] */);
int  sdpcuda_get_xlp(sdpcuda_handle* h, double* x /* [nlp] multipliers */);
int  sdpcuda_get_slp(sdpcuda_handle* h, double* s /* [nlp] slacks */);

/* ---- symmetric eigen-decomposition, batched (replaces DSYEVR behind SCIPlapackCompute*, lapack_interface.c:178-603) ----
 * A: nbatch matrices n*n (symmetric, full storage), host memory, NOT destroyed.  w: [nbatch*n] ascending.
 * V: NULL or [nbatch*n*n], eigenvector k of matrix b at V[b*n*n + k*n .. +n) (i.e. "as rows", lapack_interface.c:507-603). */
int  sdpcuda_syev_batched(sdpcuda_handle* h, int n, int nbatch, const double* A, double* w, double* V);

/* is A + shift*I positive definite?  (A symmetric n x n, host memory, not modified.)  Device Cholesky; used by the GPU
 * version of SCIPsdpSolcheckerCheck (sdpsolchecker.c:58-270: lambda_min(Z(y)) >= -feastol  <=>  Z(y) + feastol*I psd). */
int  sdpcuda_psd_check(sdpcuda_handle* h, int n, const double* A, int lda, double shift, int* is_psd);

/* the same test for the problem that is RESIDENT on the device (the reduced problem of the last sdpcuda_solve): is
 * sum_j y_j A_j^(k) - C^(k) + shift*I positive definite for every block?  y: [m] host vector or NULL = the solution of the last
 * solve (already on the device).  Nothing but y travels to the device — no dense matrix is formed on the host (SURVEY.md 8f.3:
 * the block part of SCIPsdpSolcheckerCheck, sdpsolchecker.c:166-260, after every converged solve). */
int  sdpcuda_check_psd_resident(sdpcuda_handle* h, const double* y, double shift, int* is_psd);

/* ---- kernel-level entry points (parity tests and roofline measurement; host buffers, column-major like BLAS) ---- */
/* C(m x n) = alpha*op(A)*op(B) + beta*C ; transa/transb: 0 = N, 1 = T */
int  sdpcuda_dgemm(sdpcuda_handle* h, int transa, int transb, int m, int n, int k, double alpha,
                   const double* A, int lda, const double* B, int ldb, double beta, double* C, int ldc);
/* lower Cholesky A = L L' in place (strict upper part left untouched); info = 0 or index (1-based) of the failing pivot */
int  sdpcuda_dpotrf(sdpcuda_handle* h, int n, double* A, int lda, int* info);
/* the same factorisation together with the inverse factor (what the interior-point iteration computes for S and X):
 * A <- L, Linv <- L^-1 (lower triangle, strict upper part zero) */
int  sdpcuda_dpotrf_inv(sdpcuda_handle* h, int n, double* A, int lda, double* Linv, int ldi, int* info);
/* inverse of the lower-triangular factor, in place */
int  sdpcuda_dtrtri(sdpcuda_handle* h, int n, double* L, int ldl);
/* device-resident timing of the same kernels: runs `reps` launches on random n x n operands already in HBM and
 * returns the mean device time per launch in ms (CUDA events on the handle's stream). kind: 0 dgemm NN, 1 dgemm NT,
 * 2 dpotrf, 3 dtrtri, 4 DMMA register-resident peak probe (n ignored), 5 syrk-lower NT, 6 device copy (HBM probe) */
int  sdpcuda_time_kernel(sdpcuda_handle* h, int kind, int n, int reps, double* ms_per_launch, double* flops_or_bytes);

#ifdef __cplusplus
}
#endif
#endif
