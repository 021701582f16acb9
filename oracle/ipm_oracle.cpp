/* oracle/ipm_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C++ + LAPACK/BLAS from the scipy-bundled OpenBLAS) of the algorithm the SCIP-SDP solver
 * bindings delegate to their vendor libraries: an infeasible-start primal-dual interior-point method with the HKM
 * search direction and Mehrotra predictor-corrector steps, i.e. what SDPA 7.4.4 runs behind
 * /root/reference/src/sdpi/sdpisolver_sdpa.cpp:1600-1670 (`sdpa->initializeSolve(); sdpa->solve()`), for the problem
 * form of sdpisolver_sdpa.cpp:40-57.  The interior-point arithmetic is NOT in /root/reference (DSDP 5.8 / SDPA 7.4.4 /
 * MOSEK >= 8.1 are un-vendored third-party libraries, INSTALL:8-12); the published algorithm is restated here
 * (Yamashita/Fujisawa/Kojima, "Implementation and evaluation of SDPA 6.0", and Toh/Todd/Tutuncu, SDPT3 user guide:
 * HKM direction, Schur complement M_ij = tr(A_i X A_j S^-1), Mehrotra corrector, step length from
 * lambda_min(L^-1 dX L^-T)).  Parity is pinned on the reference's own golden vectors for this path:
 * unittests/src/checksdpi.c:537-1094 (tests 1-4, 9, 10 reach the solver) and check/testset/short.solu:1-7 (B&B optima),
 * exercised through the reference's own sdpi.c compiled against this file (oracle/Makefile, tests/).
 *
 * It exports the same C ABI as the product library (include/sdpcuda.h) so the very same sdpisolver_cuda.c binding can be
 * linked against it for CPU-side tests and for the timed CPU baseline of bench.py.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 */
#include "../include/sdpcuda.h"
/* node marshalling of the product (plain host code, no numerics): shared so that drivers can be tested on the checker back end */
#define SDPNODE_EMIT_ABI 1
#include "../scip-sdp_b200/csrc/node_marshal.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" {
void scipy_dpotrf_(const char* uplo, const int* n, double* a, const int* lda, int* info);
void scipy_dtrtri_(const char* uplo, const char* diag, const int* n, double* a, const int* lda, int* info);
void scipy_dgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const double* alpha,
   const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c, const int* ldc);
void scipy_dtrsm_(const char* side, const char* uplo, const char* ta, const char* diag, const int* m, const int* n,
   const double* alpha, const double* a, const int* lda, double* b, const int* ldb);
void scipy_dsyev_(const char* jobz, const char* uplo, const int* n, double* a, const int* lda, double* w, double* work,
   const int* lwork, int* info);
void scipy_dsyevr_(const char* jobz, const char* range, const char* uplo, const int* n, double* a, const int* lda,
   const double* vl, const double* vu, const int* il, const int* iu, const double* abstol, int* m, double* w, double* z,
   const int* ldz, int* isuppz, double* work, const int* lwork, int* iwork, const int* liwork, int* info);
void scipy_dpotrs_(const char* uplo, const int* n, const int* nrhs, const double* a, const int* lda, double* b,
   const int* ldb, int* info);
void scipy_dsymv_(const char* uplo, const int* n, const double* alpha, const double* a, const int* lda, const double* x,
   const int* incx, const double* beta, double* y, const int* incy);
void scipy_dstev_(const char* jobz, const int* n, double* d, double* e, double* z, const int* ldz, double* work, int* info);
void scipy_openblas_set_num_threads(int nthreads);
int scipy_openblas_get_num_threads(void);
}

namespace {

typedef std::vector<double> vec;

struct Ent { int blk, row, col; double val; };

struct Problem
{
   int m = 0, nblocks = 0, nlp = 0, N = 0;   /* N = sum n_k + nlp */
   vec obj;
   std::vector<int> bs;
   std::vector<int> varbeg;
   std::vector<Ent> ent;
   std::vector<Ent> cent;
   std::vector<int> lpbeg, lpind;
   vec lpval, lprhs;
   std::vector<int> dense;   /* per variable: 1 = use the dense U_j = X A_j S^-1 route */
};

struct Iterate
{
   vec y;
   std::vector<vec> X, S;   /* dense n_k x n_k, symmetric, full storage */
   vec x, s;                /* LP multipliers / slacks */
};

void gemm(char ta, char tb, int n, const vec& A, const vec& B, vec& C, double alpha = 1.0, double beta = 0.0)
{
   if( n == 0 ) return;
   scipy_dgemm_(&ta, &tb, &n, &n, &n, &alpha, A.data(), &n, B.data(), &n, &beta, C.data(), &n);
}

/* lower Cholesky factor of symmetric A (full storage, column-major); returns false if not positive definite */
bool chol(int n, const vec& A, vec& L)
{
   L = A;
   if( n == 0 ) return true;
   int info = 0;
   scipy_dpotrf_("L", &n, L.data(), &n, &info);
   if( info != 0 ) return false;
   for( int c = 1; c < n; ++c )
      for( int r = 0; r < c; ++r )
         L[(size_t)c * n + r] = 0.0;
   return true;
}

/* Ainv = (L L')^-1 from the Cholesky factor */
void cholinv(int n, const vec& L, vec& Ainv, vec& Linv)
{
   Linv = L;
   if( n == 0 ) { Ainv.clear(); return; }
   int info = 0;
   scipy_dtrtri_("L", "N", &n, Linv.data(), &n, &info);
   Ainv.assign((size_t)n * n, 0.0);
   gemm('T', 'N', n, Linv, Linv, Ainv);
}

void symmetrize(int n, vec& A)
{
   for( int c = 0; c < n; ++c )
      for( int r = c + 1; r < n; ++r )
      {
         double v = 0.5 * (A[(size_t)c * n + r] + A[(size_t)r * n + c]);
         A[(size_t)c * n + r] = v;
         A[(size_t)r * n + c] = v;
      }
}

double dotm(const vec& A, const vec& B)
{
   double s = 0.0;
   for( size_t i = 0; i < A.size(); ++i ) s += A[i] * B[i];
   return s;
}

/* Smallest eigenvalue of a large symmetric matrix by Lanczos with full reorthogonalisation, the method SDPA itself uses for its
 * step lengths (Yamashita/Fujisawa/Kojima, SDPA 6.0, section on the step-size computation; Toh, "A note on the calculation of
 * step-lengths in interior-point methods for SDP").  The Ritz value is lowered by its residual bound |beta_k s_k|, so the
 * returned value is a lower bound of lambda_min up to the convergence tolerance (1e-10 relative): the step stays inside the cone. */
const int LANCZOS_MIN_ORDER = 256;
struct Prof { const char* name; std::chrono::steady_clock::time_point t0; static double acc[16]; static const char* names[16]; int id;
   Prof(int i, const char* n) : name(n), t0(std::chrono::steady_clock::now()), id(i) { names[i] = n; }
   ~Prof() { acc[id] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); } };
double Prof::acc[16]; const char* Prof::names[16];
static double g_lz_sec = 0.0; static long g_lz_steps = 0, g_lz_calls = 0;
double lanczos_lmin_impl(int n, const vec& B);
double lanczos_lmin(int n, const vec& B)
{
   auto t0 = std::chrono::steady_clock::now();
   double v = lanczos_lmin_impl(n, B);
   g_lz_sec += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); ++g_lz_calls;
   return v;
}
double lanczos_lmin_impl(int n, const vec& B)
{
   const int kmax = std::min(n, 300);
   std::vector<vec> Q;
   vec alpha, beta, q(n), w(n);
   double nrm = 0.0;
   for( int i = 0; i < n; ++i ) { q[i] = 1.0 + 0.37 * std::sin(1.0 + 0.7 * i); nrm += q[i] * q[i]; }
   nrm = std::sqrt(nrm);
   for( int i = 0; i < n; ++i ) q[i] /= nrm;
   const double one = 1.0, zero = 0.0;
   const int ione = 1;
   double theta = 0.0, bound = 1e300;
   for( int k = 0; k < kmax; ++k )
   {
      Q.push_back(q); ++g_lz_steps;
      scipy_dsymv_("L", &n, &one, B.data(), &n, q.data(), &ione, &zero, w.data(), &ione);
      double a = 0.0;
      for( int i = 0; i < n; ++i ) a += w[i] * q[i];
      alpha.push_back(a);
      for( int pass = 0; pass < 2; ++pass )                 /* full reorthogonalisation, twice */
         for( size_t j = 0; j < Q.size(); ++j )
         {
            double d = 0.0;
            for( int i = 0; i < n; ++i ) d += w[i] * Q[j][i];
            for( int i = 0; i < n; ++i ) w[i] -= d * Q[j][i];
         }
      double b = 0.0;
      for( int i = 0; i < n; ++i ) b += w[i] * w[i];
      b = std::sqrt(b);
      const int kk = k + 1;
      if( kk % 8 == 0 || kk == kmax || b <= 1e-14 * std::max(1.0, std::fabs(a)) )
      {
         vec d(alpha), e(beta), Z((size_t)kk * kk), work(std::max(1, 2 * kk));
         e.resize(std::max(1, kk));
         int info = 0;
         scipy_dstev_("V", &kk, d.data(), e.data(), Z.data(), &kk, work.data(), &info);
         theta = d[0];
         bound = std::fabs(b * Z[kk - 1]);                  /* |beta_k| * |last component of the Ritz vector| */
         if( bound <= 1e-10 * std::max(1.0, std::fabs(theta)) || b <= 1e-14 * std::max(1.0, std::fabs(a)) ) break;
      }
      beta.push_back(b);
      for( int i = 0; i < n; ++i ) q[i] = w[i] / b;
   }
   return theta - bound;
}

/* largest alpha in (0, inf] with  A + alpha*dA  psd, given the Cholesky factor L of A:  -1/lambda_min(L^-1 dA L^-T) */
double maxstep(int n, const vec& L, const vec& dA)
{
   if( n == 0 ) return 1e30;
   vec B = dA;
   double one = 1.0;
   scipy_dtrsm_("L", "L", "N", "N", &n, &n, &one, L.data(), &n, B.data(), &n);
   scipy_dtrsm_("R", "L", "T", "N", &n, &n, &one, L.data(), &n, B.data(), &n);
   symmetrize(n, B);
   double lmin;
   if( n >= LANCZOS_MIN_ORDER )
      lmin = lanczos_lmin(n, B);
   else
   {
   /* only the smallest eigenvalue: DSYEVR with index range [1,1], as lapack_interface.c:178-288 does for SCIP-SDP */
   vec w(n), work(std::max(1, 26 * n));
   std::vector<int> iwork(std::max(1, 10 * n)), isuppz(2);
   int lwork = (int)work.size(), liwork = (int)iwork.size(), info = 0, ione = 1, mfound = 0;
   double zero = 0.0, zdummy = 0.0;
   scipy_dsyevr_("N", "I", "L", &n, B.data(), &n, &zero, &zero, &ione, &ione, &zero, &mfound, w.data(), &zdummy, &ione,
      isuppz.data(), work.data(), &lwork, iwork.data(), &liwork, &info);
   lmin = w[0];
   }
   if( lmin >= -1e-300 ) return 1e30;
   return -1.0 / lmin;
}

struct Solver
{
   Problem P;
   Iterate it;
   sdpcuda_result res;
   bool solved = false;
   /* warm start staged by sdpcuda_set_start_* (one shot) and the preoptimal copy of the last solve */
   std::vector<vec> startX, startS;
   vec startx, starts;
   bool havestartlp = false;
   Iterate pre;
   bool preexists = false;

   /* --- operators --- */
   void AT(const vec& y, std::vector<vec>& out) const   /* out_k = sum_j y_j A_j^k (full symmetric) */
   {
      out.resize(P.nblocks);
      for( int k = 0; k < P.nblocks; ++k ) out[k].assign((size_t)P.bs[k] * P.bs[k], 0.0);
      for( int j = 0; j < P.m; ++j )
         for( int e = P.varbeg[j]; e < P.varbeg[j + 1]; ++e )
         {
            const Ent& t = P.ent[e];
            int n = P.bs[t.blk];
            out[t.blk][(size_t)t.col * n + t.row] += y[j] * t.val;
            if( t.row != t.col ) out[t.blk][(size_t)t.row * n + t.col] += y[j] * t.val;
         }
   }
   void Aop(const std::vector<vec>& X, vec& out) const   /* out_j = sum_k A_j^k . X^k */
   {
      out.assign(P.m, 0.0);
      for( int j = 0; j < P.m; ++j )
      {
         double s = 0.0;
         for( int e = P.varbeg[j]; e < P.varbeg[j + 1]; ++e )
         {
            const Ent& t = P.ent[e];
            int n = P.bs[t.blk];
            double v = X[t.blk][(size_t)t.col * n + t.row];
            if( t.row != t.col ) v += X[t.blk][(size_t)t.row * n + t.col];
            s += t.val * v;
         }
         out[j] = s;
      }
   }
   void Dmul(const vec& y, vec& out) const
   {
      out.assign(P.nlp, 0.0);
      for( int l = 0; l < P.nlp; ++l )
      {
         double s = 0.0;
         for( int p = P.lpbeg[l]; p < P.lpbeg[l + 1]; ++p ) s += P.lpval[p] * y[P.lpind[p]];
         out[l] = s;
      }
   }
   void DTmul(const vec& x, vec& out) const   /* out += D' x */
   {
      for( int l = 0; l < P.nlp; ++l )
         for( int p = P.lpbeg[l]; p < P.lpbeg[l + 1]; ++p ) out[P.lpind[p]] += P.lpval[p] * x[l];
   }
   void Cmat(std::vector<vec>& C) const
   {
      C.resize(P.nblocks);
      for( int k = 0; k < P.nblocks; ++k ) C[k].assign((size_t)P.bs[k] * P.bs[k], 0.0);
      for( const Ent& t : P.cent )
      {
         int n = P.bs[t.blk];
         C[t.blk][(size_t)t.col * n + t.row] += t.val;
         if( t.row != t.col ) C[t.blk][(size_t)t.row * n + t.col] += t.val;
      }
   }

   /* Schur complement  M_ij = sum_k tr(A_i X A_j S^-1) + (D' diag(x/s) D)_ij */
   void schur(const std::vector<vec>& X, const std::vector<vec>& Sinv, const vec& x, const vec& s, vec& M) const
   {
      const int m = P.m;
      M.assign((size_t)m * m, 0.0);
      std::vector<vec> U(P.nblocks);
      std::vector<vec> G(P.nblocks);
      std::vector<char> rowhit;
      for( int j = 0; j < m; ++j )
      {
         if( P.dense[j] )
         {
            /* U = X A_j Sinv per block; only rows of G = A_j Sinv touched by A_j are nonzero */
            std::vector<char> used(P.nblocks, 0);
            for( int e = P.varbeg[j]; e < P.varbeg[j + 1]; ++e )
            {
               const Ent& t = P.ent[e];
               int n = P.bs[t.blk];
               if( !used[t.blk] ) { used[t.blk] = 1; G[t.blk].assign((size_t)n * n, 0.0); U[t.blk].assign((size_t)n * n, 0.0); }
               for( int c = 0; c < n; ++c )
               {
                  G[t.blk][(size_t)c * n + t.row] += t.val * Sinv[t.blk][(size_t)c * n + t.col];
                  if( t.row != t.col ) G[t.blk][(size_t)c * n + t.col] += t.val * Sinv[t.blk][(size_t)c * n + t.row];
               }
            }
            for( int k = 0; k < P.nblocks; ++k )
               if( used[k] ) gemm('N', 'N', P.bs[k], X[k], G[k], U[k]);
            for( int i = 0; i < m; ++i )
            {
               if( P.dense[i] && i > j ) continue;
               double v = 0.0;
               for( int e = P.varbeg[i]; e < P.varbeg[i + 1]; ++e )
               {
                  const Ent& t = P.ent[e];
                  if( !used[t.blk] ) continue;
                  int n = P.bs[t.blk];
                  double u = U[t.blk][(size_t)t.col * n + t.row];
                  if( t.row != t.col ) u += U[t.blk][(size_t)t.row * n + t.col];
                  v += t.val * u;
               }
               M[(size_t)j * m + i] = v;
               M[(size_t)i * m + j] = v;
            }
         }
         else
         {
            /* entry formula: tr(A_i X A_j Sinv) = sum A_i(p,q) X(q,r) A_j(r,c) Sinv(c,p) over symmetric entries */
            for( int i = 0; i <= j; ++i )
            {
               if( P.dense[i] ) continue;
               double v = 0.0;
               for( int ei = P.varbeg[i]; ei < P.varbeg[i + 1]; ++ei )
               {
                  const Ent& a = P.ent[ei];
                  for( int ej = P.varbeg[j]; ej < P.varbeg[j + 1]; ++ej )
                  {
                     const Ent& b = P.ent[ej];
                     if( a.blk != b.blk ) continue;
                     int n = P.bs[a.blk];
                     const double* Xk = X[a.blk].data();
                     const double* Zk = Sinv[a.blk].data();
                     int p = a.row, q = a.col, r = b.row, c = b.col;
                     double t = Xk[(size_t)r * n + q] * Zk[(size_t)p * n + c];
                     if( r != c ) t += Xk[(size_t)c * n + q] * Zk[(size_t)p * n + r];
                     if( p != q )
                     {
                        t += Xk[(size_t)r * n + p] * Zk[(size_t)q * n + c];
                        if( r != c ) t += Xk[(size_t)c * n + p] * Zk[(size_t)q * n + r];
                     }
                     v += a.val * b.val * t;
                  }
               }
               M[(size_t)j * m + i] = v;
               M[(size_t)i * m + j] = v;
            }
         }
      }
      for( int l = 0; l < P.nlp; ++l )
      {
         double w = x[l] / s[l];
         for( int p = P.lpbeg[l]; p < P.lpbeg[l + 1]; ++p )
            for( int q = P.lpbeg[l]; q < P.lpbeg[l + 1]; ++q )
               M[(size_t)P.lpind[q] * m + P.lpind[p]] += w * P.lpval[p] * P.lpval[q];
      }
   }

   int solve(const sdpcuda_params& par, const double* starty);
};

int Solver::solve(const sdpcuda_params& par, const double* starty)
{
   auto t0 = std::chrono::steady_clock::now();
   const int m = P.m, nb = P.nblocks, nlp = P.nlp;
   const double gaptol = par.gaptol > 0 ? par.gaptol : 1e-6;
   const double feastol = par.feastol > 0 ? par.feastol : 1e-6;
   const int maxiter = par.maxiter > 0 ? par.maxiter : 100;
   const double inftol = 1e-8;
   double gammabase = par.setting >= 3 ? 0.7 : (par.setting == 2 ? 0.8 : 0.9);

   std::vector<vec> C; Cmat(C);
   double normb = 0, normC = 0;
   for( int j = 0; j < m; ++j ) normb += P.obj[j] * P.obj[j];
   normb = std::sqrt(normb);
   for( int k = 0; k < nb; ++k ) normC += dotm(C[k], C[k]);
   for( int l = 0; l < nlp; ++l ) normC += P.lprhs[l] * P.lprhs[l];
   normC = std::sqrt(normC);

   /* ---- initial point (SDPT3-style scaling unless lambdastar is prescribed) ---- */
   it.y.assign(m, 0.0);
   if( starty != NULL ) for( int j = 0; j < m; ++j ) it.y[j] = starty[j];
   it.X.resize(nb); it.S.resize(nb);
   {
      std::vector<vec> nrmA(nb, vec(m, 0.0));
      vec nrmD(m, 0.0);
      for( int j = 0; j < m; ++j )
         for( int e = P.varbeg[j]; e < P.varbeg[j + 1]; ++e )
            nrmA[P.ent[e].blk][j] += (P.ent[e].row == P.ent[e].col ? 1.0 : 2.0) * P.ent[e].val * P.ent[e].val;
      for( int l = 0; l < nlp; ++l )
         for( int p = P.lpbeg[l]; p < P.lpbeg[l + 1]; ++p ) nrmD[P.lpind[p]] += P.lpval[p] * P.lpval[p];
      for( int k = 0; k < nb; ++k )
      {
         int n = P.bs[k];
         double xi = std::max(10.0, std::sqrt((double)n)), eta = xi;
         double nc = std::sqrt(dotm(C[k], C[k]));
         eta = std::max(eta, nc);
         for( int j = 0; j < m; ++j )
         {
            double na = std::sqrt(nrmA[k][j]);
            if( na > 0 ) xi = std::max(xi, n * (1.0 + std::fabs(P.obj[j])) / (1.0 + na));
            eta = std::max(eta, na);
         }
         if( par.lambdastar > 0 ) xi = eta = par.lambdastar;
         it.X[k].assign((size_t)n * n, 0.0); it.S[k].assign((size_t)n * n, 0.0);
         for( int i = 0; i < n; ++i ) { it.X[k][(size_t)i * n + i] = xi; it.S[k][(size_t)i * n + i] = eta; }
      }
      double xi = std::max(10.0, std::sqrt((double)std::max(nlp, 1))), eta = xi;
      double sq = std::sqrt((double)std::max(nlp, 1));
      double nd = 0; for( int l = 0; l < nlp; ++l ) nd += P.lprhs[l] * P.lprhs[l];
      eta = std::max(eta, std::sqrt(nd));
      for( int j = 0; j < m; ++j )
      {
         double na = std::sqrt(nrmD[j]);
         if( na > 0 ) xi = std::max(xi, sq * (1.0 + std::fabs(P.obj[j])) / (1.0 + na));
         eta = std::max(eta, na);
      }
      if( par.lambdastar > 0 ) xi = eta = par.lambdastar;
      it.x.assign(nlp, xi); it.s.assign(nlp, eta);
   }
   Iterate cold = it;
   bool warm = false;
   if( starty != NULL && (int)startX.size() == nb && (int)startS.size() == nb && (nlp == 0 || (havestartlp && (int)startx.size() == nlp)) )
   {
      warm = true;
      for( int k = 0; k < nb; ++k )
         if( startX[k].size() != it.X[k].size() || startS[k].size() != it.S[k].size() ) warm = false;
      if( warm )
      {
         for( int k = 0; k < nb; ++k ) { it.X[k] = startX[k]; it.S[k] = startS[k]; }
         if( nlp > 0 ) { it.x = startx; it.s = starts; }
      }
   }
   startX.clear(); startS.clear(); startx.clear(); starts.clear(); havestartlp = false;
   preexists = false;

   std::vector<vec> ATy, Rd(nb), L(nb), Linv(nb), Sinv(nb), LX(nb), LXinv(nb), dXa(nb), dSa(nb), dX(nb), dS(nb), K(nb), T1(nb), T2(nb);
   vec rp(m), rdlp(nlp), Dy, AX, Mmat, Mfac, g(m), dy(m), dya(m), dxa(nlp), dsa(nlp), dx(nlp), ds(nlp), klp(nlp), tmpm(m);

   res = sdpcuda_result();
   res.phase = SDPCUDA_NOINFO; res.stop = SDPCUDA_STOP_ITERLIMIT;
   double mu = 0, pobj = 0, dobj = 0, relgap = 1e30, pinf = 1e30, dinf = 1e30, dinfabs = 1e30, pinfabs = 1e30;
   double bestmerit = 1e300; int stall = 0;
   bool pfeasever = false, dfeasever = false;
   int iter = 0;
   for( ; ; ++iter )
   {
      /* ---- residuals, gap, termination tests ---- */
      Prof* pr0 = new Prof(0, "residuals");
      AT(it.y, ATy);
      Aop(it.X, AX);
      Dmul(it.y, Dy);
      for( int j = 0; j < m; ++j ) rp[j] = P.obj[j] - AX[j];
      { vec t(m, 0.0); DTmul(it.x, t); for( int j = 0; j < m; ++j ) rp[j] -= t[j]; }
      double nrd = 0, nrp = 0, xs = 0, rayd = 0;
      pinfabs = 0; dinfabs = 0;
      for( int j = 0; j < m; ++j ) { nrp += rp[j] * rp[j]; pinfabs = std::max(pinfabs, std::fabs(rp[j])); }
      pobj = 0;
      for( int k = 0; k < nb; ++k )
      {
         size_t sz = it.S[k].size();
         Rd[k].resize(sz);
         double nk = 0, rk = 0;
         for( size_t i = 0; i < sz; ++i )
         {
            Rd[k][i] = ATy[k][i] - C[k][i] - it.S[k][i];
            nk += Rd[k][i] * Rd[k][i];
            double h = ATy[k][i] - it.S[k][i]; rk += h * h;
         }
         nrd += nk; rayd += rk; dinfabs = std::max(dinfabs, std::sqrt(nk));
         xs += dotm(it.X[k], it.S[k]);
         pobj += dotm(C[k], it.X[k]);
      }
      for( int l = 0; l < nlp; ++l )
      {
         rdlp[l] = Dy[l] - P.lprhs[l] - it.s[l];
         nrd += rdlp[l] * rdlp[l]; dinfabs = std::max(dinfabs, std::fabs(rdlp[l]));
         double h = Dy[l] - it.s[l]; rayd += h * h;
         xs += it.x[l] * it.s[l];
         pobj += P.lprhs[l] * it.x[l];
      }
      dobj = 0; for( int j = 0; j < m; ++j ) dobj += P.obj[j] * it.y[j];
      delete pr0;
      mu = P.N > 0 ? xs / P.N : 0.0;
      pinf = std::sqrt(nrp) / (1.0 + normb);
      dinf = std::sqrt(nrd) / (1.0 + normC);
      relgap = std::fabs(pobj - dobj) / std::max(1.0, 0.5 * (std::fabs(pobj) + std::fabs(dobj)));
      bool pfeas = pinf <= feastol && pinfabs <= std::max(feastol, 1e-9 * (1 + normb));
      bool dfeas = dinf <= feastol && dinfabs <= feastol;
      if( par.verbose )
         printf("  [oracle] it %3d  pobj % .10e  dobj % .10e  gap %.2e  pinf %.2e  dinf %.2e  mu %.2e\n", iter, pobj, dobj, relgap, pinf, dinf, mu);

      pfeasever = pfeasever || pfeas;   /* a feasible point seen once stays a proof of feasibility when a ray shows up later */
      dfeasever = dfeasever || dfeas;
      int ph = pfeas ? (dfeas ? SDPCUDA_PDFEAS : SDPCUDA_PFEAS) : (dfeas ? SDPCUDA_DFEAS : SDPCUDA_NOINFO);
      res.phase = ph;
      if( par.preoptgap > 0 && !preexists && relgap <= par.preoptgap && pinf <= std::max(feastol, par.preoptgap)
         && dinf <= std::max(feastol, par.preoptgap) )
      { pre = it; preexists = true; }
      if( pfeas && dfeas && relgap <= gaptol && (par.absgaptol <= 0 || std::fabs(pobj - dobj) <= par.absgaptol) )
      { res.phase = SDPCUDA_PDOPT; res.stop = SDPCUDA_STOP_CONVERGED; break; }
      /* y-problem infeasible: (X,x) >= 0 with A(X)+D'x -> 0 relative to C.X + d'x > 0 */
      if( pobj > 0 )
      {
         double na = 0; for( int j = 0; j < m; ++j ) { double h = P.obj[j] - rp[j]; na += h * h; }
         if( std::sqrt(na) / pobj < inftol ) { res.phase = pfeasever ? SDPCUDA_PFEAS_DINF : SDPCUDA_DINF; res.stop = SDPCUDA_STOP_INFEASCERT; break; }
      }
      /* y-problem unbounded: A'y - S -> 0, Dy - s -> 0 relative to -obj'y > 0 */
      if( dfeasever && dobj < 0 && std::sqrt(rayd) / (-dobj) < inftol )
      { res.phase = SDPCUDA_PINF_DFEAS; res.stop = SDPCUDA_STOP_INFEASCERT; break; }
      if( pfeas && par.objlimit < 1e20 && pobj > par.objlimit )
      { res.phase = SDPCUDA_PUNBD; res.stop = SDPCUDA_STOP_OBJLIMIT; break; }
      if( iter >= maxiter ) { res.stop = SDPCUDA_STOP_ITERLIMIT; break; }
      if( par.timelimit > 0 && par.timelimit < 1e20 &&
         std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > par.timelimit )
      { res.stop = SDPCUDA_STOP_TIMELIMIT; break; }
      {
         double merit = std::max(relgap, std::max(pinf, dinf));
         if( merit < 0.9 * bestmerit ) { bestmerit = merit; stall = 0; }
         else if( ++stall >= 15 ) { res.stop = SDPCUDA_STOP_NUMERICS; break; }
      }

      /* ---- factorizations ---- */
      bool ok = true;
      for( int k = 0; k < nb && ok; ++k )
      {
         { Prof p(1, "chol S,X"); ok = chol(P.bs[k], it.S[k], L[k]) && chol(P.bs[k], it.X[k], LX[k]); }
         if( ok ) { Prof p(2, "S^-1"); cholinv(P.bs[k], L[k], Sinv[k], Linv[k]); }
      }
      if( !ok && warm && iter == 0 )
      {
         /* the given start point is not interior: start again from the default point */
         vec keepy = it.y;
         it = cold; it.y = keepy;
         warm = false;
         --iter;
         continue;
      }
      if( !ok ) { res.stop = SDPCUDA_STOP_NUMERICS; break; }
      { Prof p(3, "schur"); schur(it.X, Sinv, it.x, it.s, Mmat); }
      {
         Prof p(4, "chol M");
         double reg = 0.0, maxd = 0.0;
         for( int j = 0; j < m; ++j ) maxd = std::max(maxd, Mmat[(size_t)j * m + j]);
         int info = 1, tries = 0;
         while( info != 0 && tries < 8 )
         {
            Mfac = Mmat;
            for( int j = 0; j < m; ++j ) Mfac[(size_t)j * m + j] += reg;
            if( m > 0 ) scipy_dpotrf_("L", &m, Mfac.data(), &m, &info); else info = 0;
            reg = (reg == 0.0) ? 1e-14 * std::max(maxd, 1e-300) : reg * 100.0;
            ++tries;
         }
         if( info != 0 ) { res.stop = SDPCUDA_STOP_NUMERICS; break; }
      }

      /* ---- predictor (sigma = 0) and corrector ---- */
      double sigma = 0.0, ap = 0, ad = 0;
      for( int pass = 0; pass < 2; ++pass )
      {
         /* K = sym((sigma mu I - dXa dSa - X Rd) S^-1) - X ;  klp likewise */
         Prof* pr5 = new Prof(5, "K");
         for( int k = 0; k < nb; ++k )
         {
            int n = P.bs[k];
            T1[k].assign((size_t)n * n, 0.0);
            gemm('N', 'N', n, it.X[k], Rd[k], T1[k], -1.0, 0.0);
            if( pass == 1 )
            {
               gemm('N', 'N', n, dXa[k], dSa[k], T1[k], -1.0, 1.0);
               for( int i = 0; i < n; ++i ) T1[k][(size_t)i * n + i] += sigma * mu;
            }
            K[k].assign((size_t)n * n, 0.0);
            gemm('N', 'N', n, T1[k], Sinv[k], K[k]);
            symmetrize(n, K[k]);
            for( size_t i = 0; i < K[k].size(); ++i ) K[k][i] -= it.X[k][i];
         }
         for( int l = 0; l < nlp; ++l )
         {
            double c = -it.x[l] * rdlp[l];
            if( pass == 1 ) c += sigma * mu - dxa[l] * dsa[l];
            klp[l] = c / it.s[l] - it.x[l];
         }
         delete pr5;
         Prof* pr6 = new Prof(6, "rhs+solve");
         Aop(K, g);
         DTmul(klp, g);
         for( int j = 0; j < m; ++j ) g[j] -= rp[j];
         dy = g;
         if( m > 0 )
         {
            int one = 1, info = 0;
            scipy_dpotrs_("L", &m, &one, Mfac.data(), &m, dy.data(), &m, &info);
            /* one step of iterative refinement against the unregularised M */
            vec r = g;
            for( int c = 0; c < m; ++c ) for( int rr = 0; rr < m; ++rr ) r[rr] -= Mmat[(size_t)c * m + rr] * dy[c];
            scipy_dpotrs_("L", &m, &one, Mfac.data(), &m, r.data(), &m, &info);
            for( int j = 0; j < m; ++j ) dy[j] += r[j];
         }
         /* dS = A'dy + Rd ; dX = K - sym(X (A'dy) S^-1) */
         delete pr6;
         Prof* pr7 = new Prof(7, "dS,dX");
         AT(dy, dS);
         for( int k = 0; k < nb; ++k )
         {
            int n = P.bs[k];
            T1[k].assign((size_t)n * n, 0.0); T2[k].assign((size_t)n * n, 0.0);
            gemm('N', 'N', n, it.X[k], dS[k], T1[k]);
            gemm('N', 'N', n, T1[k], Sinv[k], T2[k]);
            symmetrize(n, T2[k]);
            dX[k] = K[k];
            for( size_t i = 0; i < dX[k].size(); ++i ) { dX[k][i] -= T2[k][i]; dS[k][i] += Rd[k][i]; }
         }
         Dmul(dy, ds);
         for( int l = 0; l < nlp; ++l )
         {
            dx[l] = klp[l] - it.x[l] / it.s[l] * ds[l];
            ds[l] += rdlp[l];
         }
         /* step lengths */
         delete pr7;
         double apmax = 1e30, admax = 1e30;
         for( int k = 0; k < nb; ++k )
         {
            Prof p(8, "maxstep");
            apmax = std::min(apmax, maxstep(P.bs[k], LX[k], dX[k]));
            admax = std::min(admax, maxstep(P.bs[k], L[k], dS[k]));
         }
         for( int l = 0; l < nlp; ++l )
         {
            if( dx[l] < 0 ) apmax = std::min(apmax, -it.x[l] / dx[l]);
            if( ds[l] < 0 ) admax = std::min(admax, -it.s[l] / ds[l]);
         }
         if( pass == 0 )
         {
            ap = std::min(1.0, 0.98 * apmax); ad = std::min(1.0, 0.98 * admax);
            double xsa = 0;
            for( int k = 0; k < nb; ++k )
               for( size_t i = 0; i < it.X[k].size(); ++i )
                  xsa += (it.X[k][i] + ap * dX[k][i]) * (it.S[k][i] + ad * dS[k][i]);
            for( int l = 0; l < nlp; ++l ) xsa += (it.x[l] + ap * dx[l]) * (it.s[l] + ad * ds[l]);
            double mua = P.N > 0 ? xsa / P.N : 0.0;
            double ratio = mu > 0 ? std::max(0.0, mua / mu) : 0.0;
            double expo = (mu > 1e-6) ? std::max(1.0, 3.0 * std::min(ap, ad) * std::min(ap, ad)) : 1.0;
            sigma = std::min(1.0, std::pow(ratio, expo));
            if( par.setting >= 3 ) sigma = std::max(sigma, 0.1);
            dXa = dX; dSa = dS; dxa = dx; dsa = ds; dya = dy;
         }
         else
         {
            double gamma = gammabase + (0.99 - gammabase) * std::min(ap, ad);   /* from the predictor step lengths */
            ap = std::min(1.0, gamma * apmax); ad = std::min(1.0, gamma * admax);
         }
      }
      if( ap < 1e-8 && ad < 1e-8 ) { res.stop = SDPCUDA_STOP_NUMERICS; break; }
      for( int k = 0; k < nb; ++k )
         for( size_t i = 0; i < it.X[k].size(); ++i ) { it.X[k][i] += ap * dX[k][i]; it.S[k][i] += ad * dS[k][i]; }
      for( int l = 0; l < nlp; ++l ) { it.x[l] += ap * dx[l]; it.s[l] += ad * ds[l]; }
      for( int j = 0; j < m; ++j ) it.y[j] += ad * dy[j];
   }
   if( res.stop == SDPCUDA_STOP_NUMERICS || res.stop == SDPCUDA_STOP_ITERLIMIT || res.stop == SDPCUDA_STOP_TIMELIMIT )
   {
      /* phase already reflects feasibility of the last iterate */
   }
   res.iterations = iter; res.launches = 0;
   res.pobj = pobj; res.dobj = dobj; res.relgap = relgap; res.pinf = pinf; res.dinf = dinf; res.mu = mu;
   res.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
   res.device_ms = 0.0; res.h2d_bytes = 0.0; res.d2h_bytes = 0.0;
   solved = true;
   if( getenv("SDPORACLE_PROFILE") != NULL )
   {
      fprintf(stderr, "[oracle profile] %d iterations, %.3f s:", iter, res.seconds);
      for( int i = 0; i < 16; ++i ) if( Prof::names[i] != NULL ) { fprintf(stderr, "  %s %.3f", Prof::names[i], Prof::acc[i]); Prof::acc[i] = 0.0; }
      fprintf(stderr, "\n");
   }
   return SDPCUDA_OK;
}

} /* namespace */

struct sdpcuda_handle { Solver s; };

extern "C" {

int sdpcuda_abi_version(void) { return SDPCUDA_ABI_VERSION; }
/* checker-only: OpenBLAS thread count actually in use (torchrun exports OMP_NUM_THREADS=1, which OpenBLAS obeys) */
int sdporacle_set_threads(int n) { if( n > 0 ) scipy_openblas_set_num_threads(n); return scipy_openblas_get_num_threads(); }
int sdporacle_get_threads(void) { return scipy_openblas_get_num_threads(); }
const char* sdpcuda_backend_name(void) { return "cpu-oracle"; }

int sdpcuda_create(sdpcuda_handle** h, int device)
{
   (void)device;
   if( h == NULL ) return SDPCUDA_ERR_ARG;
   *h = new sdpcuda_handle();
   return SDPCUDA_OK;
}

int sdpcuda_destroy(sdpcuda_handle* h) { delete h; return SDPCUDA_OK; }

void sdpcuda_default_params(sdpcuda_params* p)
{
   memset(p, 0, sizeof(*p));
   p->gaptol = 1e-6; p->feastol = 1e-6; p->objlimit = 1e20; p->lambdastar = -1.0; p->timelimit = 1e20; p->preoptgap = -1.0;
   p->absgaptol = -1.0; p->maxiter = 100; p->setting = 1; p->verbose = 0;
}

int sdpcuda_solve(sdpcuda_handle* h, const sdpcuda_problem* pr, const sdpcuda_params* par, const double* start_y, sdpcuda_result* res)
{
   if( h == NULL || pr == NULL || par == NULL ) return SDPCUDA_ERR_ARG;
   Problem& P = h->s.P;
   P = Problem();
   P.m = pr->m; P.nblocks = pr->nblocks; P.nlp = pr->nlp;
   P.obj.assign(pr->obj, pr->obj + pr->m);
   P.bs.assign(pr->blocksizes, pr->blocksizes + pr->nblocks);
   P.varbeg.assign(pr->varbeg, pr->varbeg + pr->m + 1);
   int nnz = P.varbeg[P.m];
   P.ent.resize(nnz);
   for( int e = 0; e < nnz; ++e )
   {
      P.ent[e] = Ent{pr->entblk[e], pr->entrow[e], pr->entcol[e], pr->entval[e]};
      if( P.ent[e].blk < 0 || P.ent[e].blk >= P.nblocks || P.ent[e].row < P.ent[e].col || P.ent[e].row >= P.bs[P.ent[e].blk] || P.ent[e].col < 0 )
         return SDPCUDA_ERR_ARG;
   }
   P.cent.resize(pr->cnnz);
   for( int e = 0; e < pr->cnnz; ++e )
   {
      P.cent[e] = Ent{pr->cblk[e], pr->crow[e], pr->ccol[e], pr->cval[e]};
      if( P.cent[e].blk < 0 || P.cent[e].blk >= P.nblocks || P.cent[e].row < P.cent[e].col || P.cent[e].row >= P.bs[P.cent[e].blk] || P.cent[e].col < 0 )
         return SDPCUDA_ERR_ARG;
   }
   P.lpbeg.assign(pr->lpbeg, pr->lpbeg + (pr->nlp + 1) * (pr->nlp > 0 || pr->lpbeg != NULL ? 1 : 0));
   if( P.lpbeg.empty() ) P.lpbeg.assign(1, 0);
   int lnz = P.lpbeg[P.nlp];
   P.lpind.assign(pr->lpind, pr->lpind + lnz);
   P.lpval.assign(pr->lpval, pr->lpval + lnz);
   P.lprhs.assign(pr->lprhs, pr->lprhs + pr->nlp);
   P.N = P.nlp;
   for( int k = 0; k < P.nblocks; ++k ) P.N += P.bs[k];
   P.dense.assign(P.m, 0);
   for( int j = 0; j < P.m; ++j )
   {
      int cnt = P.varbeg[j + 1] - P.varbeg[j];
      int nmax = 0;
      for( int e = P.varbeg[j]; e < P.varbeg[j + 1]; ++e ) nmax = std::max(nmax, P.bs[P.ent[e].blk]);
      P.dense[j] = (cnt > std::max(4, nmax / 2)) ? 1 : 0;
   }
   int rc = h->s.solve(*par, start_y);
   if( res != NULL ) *res = h->s.res;
   return rc;
}

/* checker: the same full solve; the product library only ships obj and lprhs to the device */
int sdpcuda_solve_patched(sdpcuda_handle* h, const sdpcuda_problem* pr, const sdpcuda_params* par, const double* start_y, sdpcuda_result* res)
{
   int rc = sdpcuda_solve(h, pr, par, start_y, res);
   if( rc == SDPCUDA_OK && res != NULL ) res->h2d_bytes = 0.0;
   return rc;
}

int sdpcuda_solve_resident(sdpcuda_handle* h, const sdpcuda_params* par, sdpcuda_result* res)
{
   if( h == NULL || par == NULL ) return SDPCUDA_ERR_ARG;
   if( h->s.P.m <= 0 ) return SDPCUDA_ERR_STATE;
   int rc = h->s.solve(*par, NULL);
   if( res != NULL ) *res = h->s.res;
   return rc;
}
/* is sum_j y_j A_j - C + shift I positive definite for every block of the loaded problem?  (LAPACK Cholesky) */
int sdpcuda_check_psd_resident(sdpcuda_handle* h, const double* y, double shift, int* is_psd)
{
   if( h == NULL || is_psd == NULL ) return SDPCUDA_ERR_ARG;
   if( h->s.P.m <= 0 || (y == NULL && !h->s.solved) ) return SDPCUDA_ERR_STATE;
   const Problem& P = h->s.P;
   vec yy = (y != NULL) ? vec(y, y + P.m) : h->s.it.y;
   std::vector<vec> Z, C;
   h->s.AT(yy, Z);
   h->s.Cmat(C);
   *is_psd = 1;
   for( int k = 0; k < P.nblocks && *is_psd; ++k )
   {
      int n = P.bs[k], info = 0;
      for( size_t e = 0; e < Z[k].size(); ++e ) Z[k][e] -= C[k][e];
      for( int i = 0; i < n; ++i ) Z[k][(size_t)i * n + i] += shift;
      scipy_dpotrf_("L", &n, Z[k].data(), &n, &info);
      if( info != 0 ) *is_psd = 0;
   }
   return SDPCUDA_OK;
}
/* checker-side stand-in of the frontier batch: the nodes one after the other */
int sdpcuda_solve_batch(sdpcuda_handle* h, int count, const sdpcuda_problem* const* probs, const sdpcuda_params* par,
   sdpcuda_result* res, double* const* y_out, const double* objlimits)
{
   if( h == nullptr || count < 0 || par == nullptr || (count > 0 && probs == nullptr) ) return SDPCUDA_ERR_ARG;
   for( int i = 0; i < count; ++i ) if( probs[i] == nullptr || probs[i]->m <= 0 ) return SDPCUDA_ERR_ARG;
   for( int i = 0; i < count; ++i )
   {
      sdpcuda_params pi = *par;
      if( objlimits != nullptr ) pi.objlimit = objlimits[i];
      int rc = sdpcuda_solve(h, probs[i], &pi, nullptr, res != nullptr ? &res[i] : nullptr);
      if( rc != SDPCUDA_OK ) return rc;
      if( y_out != nullptr && y_out[i] != nullptr ) { rc = sdpcuda_get_y(h, y_out[i]); if( rc != SDPCUDA_OK ) return rc; }
   }
   return SDPCUDA_OK;
}
/* the packed image is a device-side layout of the product library: nothing to pack on the checker side */
int sdpcuda_debug_pack_node(const sdpcuda_problem* P, const sdpcuda_params* par, unsigned long long img_base, unsigned long long work_base,
   unsigned long long y_base, unsigned char* image, size_t image_cap, size_t* image_bytes, size_t* work_doubles,
   void* descriptor, size_t desc_cap, size_t* desc_bytes, int* fits)
{
   (void)P; (void)par; (void)img_base; (void)work_base; (void)y_base; (void)image; (void)image_cap; (void)image_bytes; (void)work_doubles;
   (void)descriptor; (void)desc_cap; (void)desc_bytes;
   if( fits != nullptr ) *fits = 0;
   return SDPCUDA_ERR_STATE;
}
int sdpcuda_debug_pack_batch(int count, const sdpcuda_problem* const* probs, const sdpcuda_params* par, int flags,
   unsigned long long img_base, unsigned long long work_base, unsigned long long y_base, unsigned long long res_base,
   unsigned char* image, size_t image_cap, size_t* image_bytes, size_t* work_doubles, size_t* y_doubles,
   void* descriptors, size_t desc_cap, int* nbatched, int* ntiny, int* problem_of_result, size_t* yoff_of_result, size_t* stage_bytes)
{
   (void)count; (void)probs; (void)par; (void)flags; (void)stage_bytes; (void)img_base; (void)work_base; (void)y_base; (void)res_base; (void)image; (void)image_cap;
   (void)image_bytes; (void)work_doubles; (void)y_doubles; (void)descriptors; (void)desc_cap; (void)ntiny; (void)problem_of_result; (void)yoff_of_result;
   if( nbatched != nullptr ) *nbatched = 0;
   return SDPCUDA_ERR_STATE;
}
int sdpcuda_set_profiling(sdpcuda_handle* h, int on) { (void)h; (void)on; return SDPCUDA_OK; }
int sdpcuda_get_profile(sdpcuda_handle* h, double* out)
{
   (void)h;
   for( int i = 0; i < 3 * SDPCUDA_NPROF; ++i ) out[i] = 0.0;
   return SDPCUDA_OK;
}

int sdpcuda_get_y(sdpcuda_handle* h, double* y)
{
   if( h == NULL || !h->s.solved ) return SDPCUDA_ERR_STATE;
   std::copy(h->s.it.y.begin(), h->s.it.y.end(), y);
   return SDPCUDA_OK;
}
int sdpcuda_get_X(sdpcuda_handle* h, int b, double* X)
{
   if( h == NULL || !h->s.solved ) return SDPCUDA_ERR_STATE;
   if( b < 0 || b >= h->s.P.nblocks ) return SDPCUDA_ERR_ARG;
   std::copy(h->s.it.X[b].begin(), h->s.it.X[b].end(), X);
   return SDPCUDA_OK;
}
int sdpcuda_get_S(sdpcuda_handle* h, int b, double* S)
{
   if( h == NULL || !h->s.solved ) return SDPCUDA_ERR_STATE;
   if( b < 0 || b >= h->s.P.nblocks ) return SDPCUDA_ERR_ARG;
   std::copy(h->s.it.S[b].begin(), h->s.it.S[b].end(), S);
   return SDPCUDA_OK;
}
/* the CPU restatement is a single-process checker: no sharded path */
int sdpcuda_primal_products(sdpcuda_handle* h, int ngroups, const int* groupbeg, const int* blk, const int* row, const int* col,
   const double* val, double* out)
{
   if( h == NULL || ngroups < 0 ) return SDPCUDA_ERR_ARG;
   Solver* S = &h->s;
   if( !S->solved ) return SDPCUDA_ERR_STATE;
   for( int g = 0; g < ngroups; ++g )
   {
      double acc = 0.0;
      for( int e = groupbeg[g]; e < groupbeg[g + 1]; ++e )
      {
         if( blk[e] < 0 || blk[e] >= S->P.nblocks || col[e] < 0 || row[e] < col[e] || row[e] >= S->P.bs[blk[e]] ) return SDPCUDA_ERR_ARG;
         const int n = S->P.bs[blk[e]];
         acc += (row[e] == col[e] ? 1.0 : 2.0) * val[e] * S->it.X[blk[e]][(size_t)col[e] * n + row[e]];
      }
      out[g] = acc;
   }
   return SDPCUDA_OK;
}

int sdpcuda_primal_mineig_bound(sdpcuda_handle* h, int block, double* bound)
{
   if( h == NULL || bound == NULL ) return SDPCUDA_ERR_ARG;
   Solver* S = &h->s;
   if( !S->solved ) return SDPCUDA_ERR_STATE;
   if( block < 0 || block >= S->P.nblocks ) return SDPCUDA_ERR_ARG;
   const int n = S->P.bs[block];
   double sigma = 0.0, scale = 1e-300;
   for( double v : S->it.X[block] ) scale = std::max(scale, std::fabs(v));
   for( int tries = 0; tries < 40; ++tries )
   {
      vec A = S->it.X[block], L;
      for( int i = 0; i < n; ++i ) A[(size_t)i * n + i] += sigma;
      if( chol(n, A, L) ) { *bound = -sigma; return SDPCUDA_OK; }
      sigma = (sigma == 0.0) ? 1e-14 * scale : sigma * 10.0;
   }
   *bound = -sigma;
   return SDPCUDA_OK;
}

int sdpcuda_dist_unique_id(void* id128) { (void)id128; return SDPCUDA_ERR_STATE; }
int sdpcuda_dist_init(sdpcuda_handle* h, int nranks, int rank, const void* id128)
{
   (void)id128;
   return (h != NULL && nranks == 1 && rank == 0) ? SDPCUDA_OK : SDPCUDA_ERR_STATE;
}
int sdpcuda_dist_finalize(sdpcuda_handle* h) { (void)h; return SDPCUDA_OK; }
int sdpcuda_set_start_block(sdpcuda_handle* h, int which, int block, int n, const double* A)
{
   if( h == NULL || A == NULL || block < 0 || n < 0 || which < 0 || which > 1 ) return SDPCUDA_ERR_ARG;
   std::vector<vec>& dst = (which == 0) ? h->s.startX : h->s.startS;
   if( (int)dst.size() <= block ) dst.resize(block + 1);
   dst[block].assign(A, A + (size_t)n * n);
   return SDPCUDA_OK;
}
int sdpcuda_set_start_lp(sdpcuda_handle* h, int nlp, const double* xlp, const double* slp)
{
   if( h == NULL || nlp < 0 || (nlp > 0 && (xlp == NULL || slp == NULL)) ) return SDPCUDA_ERR_ARG;
   h->s.startx.assign(xlp, xlp + nlp);
   h->s.starts.assign(slp, slp + nlp);
   h->s.havestartlp = true;
   return SDPCUDA_OK;
}
int sdpcuda_get_preopt(sdpcuda_handle* h, int* exists, double* y, double* xlp)
{
   if( h == NULL || exists == NULL ) return SDPCUDA_ERR_ARG;
   *exists = (h->s.solved && h->s.preexists) ? 1 : 0;
   if( *exists )
   {
      if( y != NULL ) std::copy(h->s.pre.y.begin(), h->s.pre.y.end(), y);
      if( xlp != NULL ) std::copy(h->s.pre.x.begin(), h->s.pre.x.end(), xlp);
   }
   return SDPCUDA_OK;
}
int sdpcuda_get_preopt_X(sdpcuda_handle* h, int b, double* X)
{
   if( h == NULL || !h->s.solved || !h->s.preexists ) return SDPCUDA_ERR_STATE;
   if( b < 0 || b >= h->s.P.nblocks ) return SDPCUDA_ERR_ARG;
   std::copy(h->s.pre.X[b].begin(), h->s.pre.X[b].end(), X);
   return SDPCUDA_OK;
}
int sdpcuda_get_xlp(sdpcuda_handle* h, double* x)
{
   if( h == NULL || !h->s.solved ) return SDPCUDA_ERR_STATE;
   std::copy(h->s.it.x.begin(), h->s.it.x.end(), x);
   return SDPCUDA_OK;
}
int sdpcuda_get_slp(sdpcuda_handle* h, double* s)
{
   if( h == NULL || !h->s.solved ) return SDPCUDA_ERR_STATE;
   std::copy(h->s.it.s.begin(), h->s.it.s.end(), s);
   return SDPCUDA_OK;
}

/* eigen-decomposition through LAPACK DSYEV: same convention as lapack_interface.c:507-603 (ascending, vectors as rows) */
int sdpcuda_syev_batched(sdpcuda_handle* h, int n, int nbatch, const double* A, double* w, double* V)
{
   (void)h;
   if( n <= 0 || nbatch < 0 ) return SDPCUDA_ERR_ARG;
   vec a((size_t)n * n), work(std::max(1, 34 * n));
   int lwork = (int)work.size(), info = 0;
   for( int b = 0; b < nbatch; ++b )
   {
      std::copy(A + (size_t)b * n * n, A + (size_t)(b + 1) * n * n, a.begin());
      scipy_dsyev_(V != NULL ? "V" : "N", "L", &n, a.data(), &n, w + (size_t)b * n, work.data(), &lwork, &info);
      if( info != 0 ) return SDPCUDA_ERR_ARG;
      if( V != NULL ) std::copy(a.begin(), a.end(), V + (size_t)b * n * n);   /* column k of a = eigenvector k = row k of V */
   }
   return SDPCUDA_OK;
}

int sdpcuda_psd_check(sdpcuda_handle* h, int n, const double* A, int lda, double shift, int* is_psd)
{
   (void)h;
   vec a((size_t)n * n);
   for( int c = 0; c < n; ++c )
      for( int r = 0; r < n; ++r ) a[(size_t)c * n + r] = A[(size_t)c * lda + r] + (r == c ? shift : 0.0);
   int info = 0;
   if( n > 0 ) scipy_dpotrf_("L", &n, a.data(), &n, &info);
   *is_psd = (info == 0);
   return SDPCUDA_OK;
}

int sdpcuda_dgemm(sdpcuda_handle* h, int ta, int tb, int m, int n, int k, double alpha, const double* A, int lda,
   const double* B, int ldb, double beta, double* C, int ldc)
{
   (void)h;
   scipy_dgemm_(ta ? "T" : "N", tb ? "T" : "N", &m, &n, &k, &alpha, A, &lda, B, &ldb, &beta, C, &ldc);
   return SDPCUDA_OK;
}
int sdpcuda_dpotrf(sdpcuda_handle* h, int n, double* A, int lda, int* info)
{
   (void)h;
   scipy_dpotrf_("L", &n, A, &lda, info);
   return SDPCUDA_OK;
}
int sdpcuda_dpotrf_inv(sdpcuda_handle* h, int n, double* A, int lda, double* Linv, int ldi, int* info)
{
   (void)h;
   scipy_dpotrf_("L", &n, A, &lda, info);
   if( *info != 0 ) return SDPCUDA_OK;
   for( int j = 0; j < n; ++j )
      for( int i = 0; i < n; ++i ) Linv[(size_t)j * ldi + i] = (i >= j) ? A[(size_t)j * lda + i] : 0.0;
   int info2 = 0;
   scipy_dtrtri_("L", "N", &n, Linv, &ldi, &info2);
   return info2 == 0 ? SDPCUDA_OK : SDPCUDA_ERR_ARG;
}
int sdpcuda_dtrtri(sdpcuda_handle* h, int n, double* L, int ldl)
{
   (void)h;
   int info = 0;
   scipy_dtrtri_("L", "N", &n, L, &ldl, &info);
   return info == 0 ? SDPCUDA_OK : SDPCUDA_ERR_ARG;
}
int sdpcuda_time_kernel(sdpcuda_handle* h, int kind, int n, int reps, double* ms, double* work)
{
   (void)h; (void)kind; (void)n; (void)reps; (void)ms; (void)work;
   return SDPCUDA_ERR_ARG;   /* device timing only exists in the product library */
}

} /* extern "C" */
