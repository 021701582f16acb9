// ops.cuh — memory-bound kernels of the interior-point iteration: sparse operators A(.) and A'(.), LP-block operators,
// Schur-complement assembly (entry/gather path and LP block), element-wise updates and deterministic reductions.
// All dense SDP-block matrices live in one "arena" (block k at offset off[k], column-major, leading dimension ld[k]).
#pragma once
#include "common.cuh"

namespace sdpk {

struct DevEntries            // constraint-matrix entries, CSR over variables
{
   const int* varbeg;        // [m+1]
   const int* row;           // [nnz]  row >= col
   const int* col;
   const int* ld;            // leading dimension of the entry's block
   const long long* off;     // arena offset of the entry's block
   const double* val;
};

constexpr int RED_BLOCKS = 256;      // partial sums per reduction (stage 1), summed in fixed order (stage 2)

// T[pos] = T[mirror] = -c + sum_j y_j a_j  over the positions of the aggregate sparsity pattern (T zeroed before)
cudaError_t assemble_positions(cudaStream_t st, int npos, const int* posbeg, const long long* pos, const long long* mirror,
   const int* posvar, const double* posval, const double* posc, const double* y, double cscale, double* T);
// Rd = T - S over the arena; stats[0] += sum Rd^2, stats[1] += sum X.S   (partials -> finalize)
cudaError_t residual_matrix(cudaStream_t st, size_t arena, const double* T, const double* S, const double* X, double* Rd, double* partials);
// out[j] = sum_e val * (X[p1] + X[p2]) per variable
// skipcls (or nullptr): variables with skipcls[j] == 2 are left to apply_A_dense
cudaError_t apply_A(cudaStream_t st, int m, DevEntries E, const double* X, double* out, const int* skipcls = nullptr);
// dense constraint matrices of one block (expanded copies, matrix d at Ad + d*stride, stride = ld*n of the block):
// out[denselist[first + d]] = <A_d, Xk>  and  Tk += sum_d v[denselist[first + d]] A_d   -- streamed, no index arrays
cudaError_t apply_A_dense(cudaStream_t st, int count, int first, const int* denselist, const double* Ad, long long stride, const double* Xk, double* out);
cudaError_t assemble_dense(cudaStream_t st, int count, int first, const int* denselist, const double* Ad, long long stride, const double* v, double* Tk);
// sparse constant matrix: stats = sum c * (X[pos] (+ X[mirror])), same for a second matrix Y (or nullptr)
cudaError_t const_dots(cudaStream_t st, int cnnz, const long long* cpos, const long long* cmirror, const double* cval,
   const double* X, const double* Y, double* out2);
// LP rows: Dy, rdlp = Dy - d - s, partial statistics
cudaError_t lp_rows(cudaStream_t st, int nlp, const int* lpbeg, const int* lpind, const double* lpval, const double* lprhs,
   const double* y, const double* x, const double* s, double* Dy, double* rdlp, double* partials);
// out[j] (+)= sum over column j of D: val * x[row]
cudaError_t lp_cols(cudaStream_t st, int m, const int* colbeg, const int* colrow, const double* colval, const double* x,
   double* out, int accumulate);
// rp = b - AX - DTx and statistics
cudaError_t primal_residual(cudaStream_t st, int m, const double* b, const double* AX, const double* DTx, const double* y,
   double* rp, double* partials);
cudaError_t finalize_partials(cudaStream_t st, const double* partials, int nstats, double* out);

// Schur complement (lower triangle of M, ldm): entry/gather path over all variable pairs, LP block via atomics
// (nranks, rank): this launch only forms the share of `rank` out of `nranks` (column strips / partner variables dealt
// round-robin; every entry of M belongs to exactly one rank); (1, 0) = everything
cudaError_t schur_entries(cudaStream_t st, int m, DevEntries E, const int* heavy /* [m] 0/1 */, const int* heavylist, int nheavy,
   const double* X, const double* Z, double* M, int ldm, int nranks = 1, int rank = 0);
// dense path: variables whose constraint matrix is dense in one block.  scatter_dense expands them into full n x n matrices
// (Adense, matrix d at d*stride); schur_dense_dots then forms M_ij = A_i . U_j for every variable i and every dense j from
// U_j = X A_j S^-1 (two batched DMMA GEMMs in between), lower triangle only, each entry written exactly once.
cudaError_t scatter_dense(cudaStream_t st, int nd, const int* denselist, DevEntries E, int ld, long long stride, double* Adense);
cudaError_t schur_dense_dots(cudaStream_t st, int m, int nd, int d0, const int* denselist, const int* cls, DevEntries E, long long blockoff,
   const double* U, int ld, long long stride, double* M, int ldm);
// dense x dense pairs: C = Adense' U (count x cnt, from the DMMA GEMM, possibly as nslices partial products over k) scattered
// to M[i, j], i >= j
cudaError_t schur_dense_scatter(cudaStream_t st, int count, int cnt, int first, int first_j, const int* denselist, const double* C, int ldc,
   int nslices, long long slicestride, double* M, int ldm);
// colbeg/colrow/colval (column lists of the LP block, rows ascending): deterministic single-writer form; nullptr: atomic form
cudaError_t schur_lp(cudaStream_t st, int nlp, const int* lpbeg, const int* lpind, const double* lpval, const double* x,
   const double* s, double* M, int ldm, const int* colbeg = nullptr, const int* colrow = nullptr, const double* colval = nullptr,
   int maxcnt = 0);
cudaError_t schur_rank1_scatter(cudaStream_t st, int r, const int* var, const double* sig, const double* G1, const double* G2, int ldg,
   double* M, int ldm);
cudaError_t add_diagonal(cudaStream_t st, int n, double* A, int lda, double v);
cudaError_t symv_lower(cudaStream_t st, int n, const double* M, int ldm, const double* x, double* y);   // y = M x, M lower stored
// y = T x (trans = 0) or y = T' x (trans = 1) for a lower-triangular T (upper part never read)
cudaError_t trmv_lower(cudaStream_t st, int n, const double* T, int ldt, int trans, const double* x, double* y);
// y = U' x for an upper-triangular U (lower part never read): one warp per (contiguous) column
cudaError_t trmv_upper_t(cudaStream_t st, int n, const double* U, int ldu, const double* x, double* y);
// B = A' (n x n, out of place)
cudaError_t transpose(cudaStream_t st, int n, const double* A, int lda, double* B, int ldb);
// K = sym(P Z) - X on the pattern entries only, P given transposed (Pt); W: scratch matrix (pattern entries written)
cudaError_t sddmm_pattern_sym(cudaStream_t st, int n, const double* Pt, int ldp, const double* Z, int ldz, const double* X, int ldx,
   const int* colptr, const int* rowidx, double* W, int ldw, double* K, int ldk);

// Out = A * D for one block, D symmetric and sparse: its pattern is given column-wise (colptr[n+1], rowidx), its values
// are read from the dense array D itself.  One CTA per (256-row slab, column): gathers columns of A (coalesced).
cudaError_t spmm_pattern(cudaStream_t st, int n, const double* A, int lda, const double* D, int ldd, const int* colptr,
   const int* rowidx, double alpha, double* Out, int ldo);

// dense block helpers
cudaError_t sym_average(cudaStream_t st, int n, double* A, int lda, const double* subtract /* or nullptr */);   // A = (A+A')/2 - subtract
cudaError_t mirror_lower(cudaStream_t st, int n, double* A, int lda);                                            // upper = lower'
cudaError_t axpy(cudaStream_t st, size_t n, double a, const double* x, double* y);                               // y += a x
cudaError_t axpby_out(cudaStream_t st, size_t n, double a, const double* x, double b, const double* y, double* out);
// LP direction: klp = (sigmamu - dxa*dsa - x*rd)/s - x   (corr = 0 drops the first two terms)
cudaError_t lp_rhs(cudaStream_t st, int nlp, int corr, double sigmamu, const double* x, const double* s, const double* rd,
   const double* dxa, const double* dsa, double* klp);
// ds = D dy + rd (Ddy given), dx = klp - x/s * (D dy); also max step lengths to the boundary (partials: min ratios)
cudaError_t lp_direction(cudaStream_t st, int nlp, const double* x, const double* s, const double* rd, const double* klp,
   const double* Ddy, double* dx, double* ds, double* out2 /* alpha_p_max, alpha_d_max */);
// sum (x + ap dx)(s + ad ds) and sum (X + ap dX).(S + ad dS)
cudaError_t affine_mu(cudaStream_t st, size_t arena, const double* X, const double* dX, const double* S, const double* dS,
   int nlp, const double* x, const double* dx, const double* s, const double* ds, double ap, double ad, double* partials);
cudaError_t pick_value(cudaStream_t st, const double* src, double* dst);

} // namespace sdpk
