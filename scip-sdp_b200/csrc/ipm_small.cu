// ipm_small.cu — the whole interior-point solve of a small relaxation in ONE kernel launch.
//
// SCIP-SDP's branch-and-bound solves thousands of tiny relaxations (shipped instances: blocks of order 2..43, 3..105
// variables).  On that scale a kernel-per-operation pipeline is pure launch latency (hundreds of launches and three
// host round trips per iteration).  Here one CTA of 1024 threads runs the complete predictor-corrector iteration -
// residuals, Cholesky factors and inverses of S and X, Schur complement, its factorisation, both solves, the HKM
// directions, Lanczos step lengths, the update and all termination tests - with __syncthreads() as the only
// synchronisation.  Data stay in the same HBM/L2-resident buffers as in the multi-kernel path (ipm.cu), small blocks are
// factorised in shared memory.  The numerical recipe is identical to ipm.cu (same formulas, same status rules), so both
// paths are interchangeable behind sdpcuda_solve; other CTAs/SMs stay free for other solver handles (concurrent nodes).
#include "ipm_small.cuh"
#include "../../include/sdpcuda.h"

namespace sdpk {
namespace {

#ifdef SDPK_VARIANT_TINY
// second instantiation of this file (ipm_tiny.cu) for frontier batches of relaxations with blocks of order <= 16: 256 threads and
// 41 KB of shared memory per CTA, so that four CTAs (= four nodes) share an SM; the Schur factor still fits shared memory up to m = 64
constexpr int NT = 256;
constexpr int VMAXN = TINY_MAX_N;
constexpr int MSN = 64;
constexpr int MINB = 4;
#define ipm_small_batch_kernel ipm_tiny_batch_kernel
#else
constexpr int NT = 1024;
constexpr int VMAXN = SMALL_MAX_N;       // largest block
constexpr int MSN = SMALL_MAX_N;         // largest Schur complement whose factor lives in shared memory
constexpr int MINB = 1;
#endif
constexpr int LDS = VMAXN + 1;
constexpr int LDMS = MSN + 1;
#ifdef SDPK_VARIANT_TINY
constexpr int MPK = TINY_MPK;
#else
constexpr int MPK = SMALL_MPK;
#endif

struct Ctl                      // control block in shared memory, written by thread 0
{
   double st[40];
   double mu, pobj, dobj, relgap, pinf, dinf, sigma, ap, ad, apmax, admax, lastap, lastad, bestmerit, lam;
   int iter, stall, backtracks, phase, stop, done, rdzero, failS, failX, failM, pfeasever, dfeasever, xfail;
};

__device__ __forceinline__ double wsum(double v)
{
#pragma unroll
   for( int o = 16; o > 0; o >>= 1 ) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}

__device__ double bsum(double v, double* red)
{
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   v = wsum(v);
   __syncthreads();
   if( lane == 0 ) red[w] = v;
   __syncthreads();
   double s = 0.0;
#pragma unroll
   for( int q = 0; q < NT / 32; ++q ) s += red[q];
   return s;
}

__device__ double bmax(double v, double* red)
{
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
   for( int o = 16; o > 0; o >>= 1 ) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
   __syncthreads();
   if( lane == 0 ) red[w] = v;
   __syncthreads();
   double s = red[0];
#pragma unroll
   for( int q = 1; q < NT / 32; ++q ) s = fmax(s, red[q]);
   return s;
}

__device__ __forceinline__ double frsqrt(double x)
{
   double y = (double)rsqrtf((float)x);
   y = y * (1.5 - 0.5 * x * y * y);
   y = y * (1.5 - 0.5 * x * y * y);
   return y;
}

// T = sum_j v_j A_j - cscale * C  (full symmetric, zero elsewhere)
__device__ void assemble(const SmallArgs& a, const double* v, double cscale, double* T)
{
   for( long long i = threadIdx.x; i < a.arena; i += NT ) T[i] = 0.0;
   __syncthreads();
   for( int p = threadIdx.x; p < a.npos; p += NT )
   {
      double s = -cscale * a.posc[p];
      for( int e = a.posbeg[p]; e < a.posbeg[p + 1]; ++e ) s += v[a.posvar[e]] * a.posval[e];
      T[a.pos[p]] = s;
      T[a.mirror[p]] = s;
   }
   __syncthreads();
}

// out_j = A_j . Xm : one warp per variable
__device__ void applyA(const SmallArgs& a, const double* Xm, double* out)
{
   const int lane = threadIdx.x & 31;
   for( int j = threadIdx.x >> 5; j < a.m; j += NT / 32 )
   {
      double s = 0.0;
      for( int e = a.E.varbeg[j] + lane; e < a.E.varbeg[j + 1]; e += 32 )
      {
         int r = a.E.row[e], c = a.E.col[e], ld = a.E.ld[e];
         const double* Xk = Xm + a.E.off[e];
         double u = Xk[(size_t)c * ld + r];
         if( r != c ) u += Xk[(size_t)r * ld + c];
         s += a.E.val[e] * u;
      }
      s = wsum(s);
      if( lane == 0 ) out[j] = s;
   }
   __syncthreads();
}

__device__ void lpcols(const SmallArgs& a, const double* xv, double* out, bool accumulate)
{
   for( int j = threadIdx.x; j < a.m; j += NT )
   {
      double s = 0.0;
      for( int p = a.colbeg[j]; p < a.colbeg[j + 1]; ++p ) s += a.colval[p] * xv[a.colrow[p]];
      out[j] = accumulate ? out[j] + s : s;
   }
   __syncthreads();
}

__device__ void lprows(const SmallArgs& a, const double* yv, double* out)
{
   for( int l = threadIdx.x; l < a.nlp; l += NT )
   {
      double s = 0.0;
      for( int p = a.lpbeg[l]; p < a.lpbeg[l + 1]; ++p ) s += a.lpval[p] * yv[a.lpind[p]];
      out[l] = s;
   }
   __syncthreads();
}

// C = alpha * op(A) * op(B) + beta * C for one block (n <= 64), all matrices n x n with leading dimension ld
__device__ void gemmb(int n, int ld, bool ta, bool tb, double alpha, const double* A, const double* B, double beta, double* C)
{
   for( int e = threadIdx.x; e < n * n; e += NT )
   {
      const int i = e % n, j = e / n;
      double s0 = 0.0, s1 = 0.0;
      int k = 0;
      for( ; k + 2 <= n; k += 2 )
      {
         double a0 = ta ? A[(size_t)i * ld + k] : A[(size_t)k * ld + i];
         double a1 = ta ? A[(size_t)i * ld + k + 1] : A[(size_t)(k + 1) * ld + i];
         double b0 = tb ? B[(size_t)k * ld + j] : B[(size_t)j * ld + k];
         double b1 = tb ? B[(size_t)(k + 1) * ld + j] : B[(size_t)j * ld + k + 1];
         s0 += a0 * b0; s1 += a1 * b1;
      }
      if( k < n )
      {
         double a0 = ta ? A[(size_t)i * ld + k] : A[(size_t)k * ld + i];
         double b0 = tb ? B[(size_t)k * ld + j] : B[(size_t)j * ld + k];
         s0 += a0 * b0;
      }
      double v = alpha * (s0 + s1);
      if( beta != 0.0 ) v += beta * C[(size_t)j * ld + i];
      C[(size_t)j * ld + i] = v;
   }
   __syncthreads();
}

// A = (A + A')/2 - sub
__device__ void symavg(int n, int ld, double* A, const double* sub)
{
   for( int e = threadIdx.x; e < n * n; e += NT )
   {
      const int i = e % n, j = e / n;
      if( i < j ) continue;
      double v = 0.5 * (A[(size_t)j * ld + i] + A[(size_t)i * ld + j]);
      if( sub != nullptr ) v -= sub[(size_t)j * ld + i];
      A[(size_t)j * ld + i] = v;
      A[(size_t)i * ld + j] = v;
   }
   __syncthreads();
}

// Cholesky factor and its inverse of one block in shared memory; returns false (to all threads) on a non-positive pivot
__device__ bool cholinv(int n, int ld, const double* src, double* Lout, double* Linvout, double* sh, double* sh2, double* rdiag, int* flag)
{
   for( int e = threadIdx.x; e < n * n; e += NT )
   {
      const int i = e % n, j = e / n;
      sh[i * LDS + j] = (i >= j) ? src[(size_t)j * ld + i] : 0.0;
      sh2[i * LDS + j] = (i == j) ? 1.0 : 0.0;
   }
   if( threadIdx.x == 0 ) *flag = 0;
   __syncthreads();
   for( int k = 0; k < n; ++k )
   {
      if( threadIdx.x == 0 )
      {
         double d = sh[k * LDS + k];
         if( !(d > 0.0) ) { *flag = 1; d = 1.0; }
         double r = frsqrt(d);
         rdiag[k] = r;
         sh[k * LDS + k] = d * r;
      }
      __syncthreads();
      const double r = rdiag[k];
      for( int i = k + 1 + threadIdx.x; i < n; i += NT ) sh[i * LDS + k] *= r;
      __syncthreads();
      const int rem = n - k - 1;
      for( int e = threadIdx.x; e < rem * rem; e += NT )
      {
         const int i = k + 1 + e % rem, j = k + 1 + e / rem;
         if( i >= j ) sh[i * LDS + j] -= sh[i * LDS + k] * sh[j * LDS + k];
      }
      __syncthreads();
   }
   // W = L^-1 from R = I, row by row
   for( int k = 0; k < n; ++k )
   {
      const double r = rdiag[k];
      for( int j = threadIdx.x; j <= k; j += NT ) sh2[k * LDS + j] *= r;
      __syncthreads();
      const int below = n - k - 1;
      for( int e = threadIdx.x; e < below * (k + 1); e += NT )
      {
         const int i = k + 1 + e / (k + 1), j = e % (k + 1);
         sh2[i * LDS + j] -= sh[i * LDS + k] * sh2[k * LDS + j];
      }
      __syncthreads();
   }
   for( int e = threadIdx.x; e < n * n; e += NT )
   {
      const int i = e % n, j = e / n;
      Lout[(size_t)j * ld + i] = (i >= j) ? sh[i * LDS + j] : 0.0;
      Linvout[(size_t)j * ld + i] = (i >= j) ? sh2[i * LDS + j] : 0.0;
   }
   __syncthreads();
   const bool ok = (*flag == 0);
   __syncthreads();                 // the flag is reused by the next call
   return ok;
}

// Schur complement entry formula for one pair of variables (same as ops.cu)
__device__ __forceinline__ double pairterm(const DevEntries& E, int ei, int ej, const double* X, const double* Z)
{
   if( E.off[ei] != E.off[ej] ) return 0.0;
   const int ld = E.ld[ei];
   const double* Xk = X + E.off[ei];
   const double* Zk = Z + E.off[ei];
   const int p = E.row[ei], q = E.col[ei], r = E.row[ej], c = E.col[ej];
   double t = Xk[(size_t)r * ld + q] * Zk[(size_t)p * ld + c];
   if( r != c ) t += Xk[(size_t)c * ld + q] * Zk[(size_t)p * ld + r];
   if( p != q )
   {
      t += Xk[(size_t)r * ld + p] * Zk[(size_t)q * ld + c];
      if( r != c ) t += Xk[(size_t)c * ld + p] * Zk[(size_t)q * ld + r];
   }
   return E.val[ei] * E.val[ej] * t;
}

// C_d = A_d B_d for d < cnt, all n x n with leading dimension ld; 4 x 2 register tiles, rows up to ld (padding rows of C become zero),
// ends with a barrier.  One operand is the same matrix for all d (SHARED_A: A, else B): it is staged in the scratch tile at the head
// of the shared memory (free between the phases), the other one streams from global memory, 16 bytes per load.
__device__ __forceinline__ void ld2(const double* p, double& x, double& y)
{
#ifdef __CUDA_ARCH__
   const double2 v = *reinterpret_cast<const double2*>(p);
   x = v.x; y = v.y;
#else
   x = p[0]; y = p[1];
#endif
}
// (Measured and not kept: staging the streamed operand as well, two matrices at a time by the whole CTA - only 484 of the 1024
// threads have a tile then, and the 17 stages of example_CLS cost more in barriers than the loads from the L2 did: Schur phase
// 12.4 M -> 13.7 M cycles.)
template <bool SHARED_A>
__device__ __noinline__ void dense_product(int n, int ld, int cnt, long long stride, const double* A, const double* B, double* C)
{
   extern __shared__ __align__(16) double smem[];
   double* sh = smem;                                // VMAXN x (VMAXN + 1) doubles, n * ld fits
   {
      const double* S = SHARED_A ? A : B;
      for( int e = threadIdx.x; e < n * ld; e += NT ) sh[e] = S[e];
   }
   __syncthreads();
   const int rt = ld / 4, ct = (n + 1) / 2, per = rt * ct;
   for( int t = threadIdx.x; t < cnt * per; t += NT )
   {
      const int d = t / per, r = t % per, i0 = 4 * (r % rt), j0 = 2 * (r / rt);
      const bool two = j0 + 1 < n;
      const double* Ad = (SHARED_A ? sh : A + (size_t)d * stride) + i0;
      const double* B0 = (SHARED_A ? B + (size_t)d * stride : sh) + (size_t)j0 * ld;
      const double* B1 = B0 + (two ? ld : 0);
      double c00 = 0.0, c10 = 0.0, c20 = 0.0, c30 = 0.0, c01 = 0.0, c11 = 0.0, c21 = 0.0, c31 = 0.0;
      int k = 0;
      for( ; k + 2 <= n; k += 2 )                   // two steps of k per round: the columns of B come in pairs (ld and j0 * ld are even)
      {
         double a0, a1, a2, a3, e0, e1, e2, e3, b00, b01, b10, b11;
         ld2(Ad + (size_t)k * ld, a0, a1); ld2(Ad + (size_t)k * ld + 2, a2, a3);
         ld2(Ad + (size_t)(k + 1) * ld, e0, e1); ld2(Ad + (size_t)(k + 1) * ld + 2, e2, e3);
         ld2(B0 + k, b00, b01); ld2(B1 + k, b10, b11);
         c00 += a0 * b00; c10 += a1 * b00; c20 += a2 * b00; c30 += a3 * b00;
         c01 += a0 * b10; c11 += a1 * b10; c21 += a2 * b10; c31 += a3 * b10;
         c00 += e0 * b01; c10 += e1 * b01; c20 += e2 * b01; c30 += e3 * b01;
         c01 += e0 * b11; c11 += e1 * b11; c21 += e2 * b11; c31 += e3 * b11;
      }
      if( k < n )
      {
         double a0, a1, a2, a3;
         ld2(Ad + (size_t)k * ld, a0, a1); ld2(Ad + (size_t)k * ld + 2, a2, a3);
         const double b0 = B0[k], b1 = B1[k];
         c00 += a0 * b0; c10 += a1 * b0; c20 += a2 * b0; c30 += a3 * b0;
         c01 += a0 * b1; c11 += a1 * b1; c21 += a2 * b1; c31 += a3 * b1;
      }
      // the padding rows of C are written as zeros whatever the padding rows of A hold (they take part in the long dot products)
      if( i0 + 1 >= n ) { c10 = 0.0; c11 = 0.0; }
      if( i0 + 2 >= n ) { c20 = 0.0; c21 = 0.0; }
      if( i0 + 3 >= n ) { c30 = 0.0; c31 = 0.0; }
      double* C0 = C + (size_t)d * stride + (size_t)j0 * ld + i0;
      C0[0] = c00; C0[1] = c10; C0[2] = c20; C0[3] = c30;
      if( two ) { C0[ld] = c01; C0[ld + 1] = c11; C0[ld + 2] = c21; C0[ld + 3] = c31; }
   }
   __syncthreads();
}

// M[v_i, v_j] = <A_i, U_j> for the dense variables of one group (v = their variable indices, pairs with v_i >= v_j), A and U stored
// as cnt contiguous matrices of `stride` doubles with zero padding rows.  One warp per tile of 4 x 4 pairs, the lanes split the
// elements: every matrix is read once per tile instead of once per pair - with a whole frontier at work (one node per SM) the
// pair-by-pair version read 17 MB per node and iteration through the L2 and was bound by it.  No barrier at the end.
__device__ __noinline__ void dense_pair_dots(const int* vars, int cnt, long long stride, const double* A, const double* U, double* M, int ldm)
{
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int nt = (cnt + 3) / 4;
   for( int t = wid; t < nt * nt; t += NT / 32 )
   {
      const int d1 = 4 * (t / nt), d2 = 4 * (t % nt);
      int imax = -1, jmin = 0x7fffffff;
      for( int q = 0; q < 4; ++q )
      {
         if( d1 + q < cnt ) imax = max(imax, vars[d1 + q]);
         if( d2 + q < cnt ) jmin = min(jmin, vars[d2 + q]);
      }
      if( imax < jmin ) continue;                    // no pair of this tile belongs to the lower triangle
      // rows beyond cnt repeat the last matrix (computed, not stored)
      const double* Ab = A + (size_t)d1 * stride;
      const double* Ub = U + (size_t)d2 * stride;
      const int st = (int)stride;
      const int oa1 = (min(d1 + 1, cnt - 1) - d1) * st, oa2 = (min(d1 + 2, cnt - 1) - d1) * st, oa3 = (min(d1 + 3, cnt - 1) - d1) * st;
      const int ou1 = (min(d2 + 1, cnt - 1) - d2) * st, ou2 = (min(d2 + 2, cnt - 1) - d2) * st, ou3 = (min(d2 + 3, cnt - 1) - d2) * st;
      double c00 = 0.0, c10 = 0.0, c20 = 0.0, c30 = 0.0, c01 = 0.0, c11 = 0.0, c21 = 0.0, c31 = 0.0;
      double c02 = 0.0, c12 = 0.0, c22 = 0.0, c32 = 0.0, c03 = 0.0, c13 = 0.0, c23 = 0.0, c33 = 0.0;
#pragma unroll 1
      for( int e = lane; e < st; e += 32 )
      {
         const double* Ae = Ab + e;
         const double* Ue = Ub + e;
         const double a0 = Ae[0], a1 = Ae[oa1], a2 = Ae[oa2], a3 = Ae[oa3];
         double u = Ue[0];
         c00 += a0 * u; c10 += a1 * u; c20 += a2 * u; c30 += a3 * u;
         u = Ue[ou1];
         c01 += a0 * u; c11 += a1 * u; c21 += a2 * u; c31 += a3 * u;
         u = Ue[ou2];
         c02 += a0 * u; c12 += a1 * u; c22 += a2 * u; c32 += a3 * u;
         u = Ue[ou3];
         c03 += a0 * u; c13 += a1 * u; c23 += a2 * u; c33 += a3 * u;
      }
#define SDPK_PAIR_OUT(q1, q2, acc) do { const double sum_ = wsum(acc); \
         if( lane == 0 && d1 + q1 < cnt && d2 + q2 < cnt ) { const int vi_ = vars[d1 + q1], vj_ = vars[d2 + q2]; \
            if( vi_ >= vj_ ) M[(size_t)vj_ * ldm + vi_] = sum_; } } while( 0 )
      SDPK_PAIR_OUT(0, 0, c00); SDPK_PAIR_OUT(1, 0, c10); SDPK_PAIR_OUT(2, 0, c20); SDPK_PAIR_OUT(3, 0, c30);
      SDPK_PAIR_OUT(0, 1, c01); SDPK_PAIR_OUT(1, 1, c11); SDPK_PAIR_OUT(2, 1, c21); SDPK_PAIR_OUT(3, 1, c31);
      SDPK_PAIR_OUT(0, 2, c02); SDPK_PAIR_OUT(1, 2, c12); SDPK_PAIR_OUT(2, 2, c22); SDPK_PAIR_OUT(3, 2, c32);
      SDPK_PAIR_OUT(0, 3, c03); SDPK_PAIR_OUT(1, 3, c13); SDPK_PAIR_OUT(2, 3, c23); SDPK_PAIR_OUT(3, 3, c33);
#undef SDPK_PAIR_OUT
   }
}

#ifdef SDPK_SCHURTICKS
__device__ long long g_schur_ticks[8];
#define SCHURTICK(k) do { __syncthreads(); if( threadIdx.x == 0 ) { long long n_ = clock64(); g_schur_ticks[k] += n_ - tq_; tq_ = n_; } } while( 0 )
#else
#define SCHURTICK(k) do { } while( 0 )
#endif
__device__ void schur(const SmallArgs& a)
{
#ifdef SDPK_SCHURTICKS
   long long tq_ = clock64();
#endif
   const int m = a.m, ldm = a.ldm;
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   for( int e = threadIdx.x; e < ldm * m; e += NT ) a.M[e] = 0.0;
   __syncthreads();
   // pairs without a dense variable.  Light pairs (at most SCHUR_LIGHT entry-pair terms: the one-entry matrices of example_MkP, the
   // rank-one bars of example_TT) get ONE THREAD each - a warp per pair left 31 lanes idle and made the 5565 pairs of example_MkP
   // 174 rounds of dependent global loads per warp; heavier pairs keep one warp per pair, the lanes split the entry-pair product
   constexpr int SCHUR_LIGHT = 16;
   const int npairs = m * (m + 1) / 2;
   for( int p = threadIdx.x; p < npairs; p += NT )
   {
      int i = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
      while( (i + 1) * (i + 2) / 2 <= p ) ++i;
      while( i * (i + 1) / 2 > p ) --i;
      const int j = p - i * (i + 1) / 2;
      if( a.cls[i] == 2 || a.cls[j] == 2 ) continue;
      const int bi = a.E.varbeg[i], ni = a.E.varbeg[i + 1] - bi;
      const int bj = a.E.varbeg[j], nj = a.E.varbeg[j + 1] - bj;
      if( ni * nj > SCHUR_LIGHT ) continue;
      double v = 0.0;
      for( int t = 0; t < ni * nj; ++t ) v += pairterm(a.E, bi + t / nj, bj + t % nj, a.X, a.Sinv);
      a.M[(size_t)j * ldm + i] = v;
   }
   SCHURTICK(0);
   // heavier pairs: one warp per pair, the lanes split the entry-pair product.  A warp takes the rows i = wid, wid + 32 warps, ... and
   // looks at 32 columns j at a time (one per lane, ballot of the heavy ones): walking the pairs one by one cost a square root, six
   // dependent loads and a branch per pair and warp just to find out that a pair is light (example_MkP: all 4656 of them)
   for( int i = wid; i < m; i += NT / 32 )
   {
      if( a.cls[i] == 2 ) continue;
      const int bi = a.E.varbeg[i], ni = a.E.varbeg[i + 1] - bi;
      for( int j0 = 0; j0 <= i; j0 += 32 )
      {
         const int jl = j0 + lane;
         bool heavy = false;
         if( jl <= i && a.cls[jl] != 2 ) heavy = ni * (a.E.varbeg[jl + 1] - a.E.varbeg[jl]) > SCHUR_LIGHT;
         unsigned mask = __ballot_sync(0xffffffffu, heavy);
         while( mask != 0 )
         {
            const int j = j0 + __ffs(mask) - 1;
            mask &= mask - 1;
            const int bj = a.E.varbeg[j], nj = a.E.varbeg[j + 1] - bj;
            double v = 0.0;
            for( int t = lane; t < ni * nj; t += 32 ) v += pairterm(a.E, bi + t / nj, bj + t % nj, a.X, a.Sinv);
            v = wsum(v);
            if( lane == 0 ) a.M[(size_t)j * ldm + i] = v;
         }
      }
   }
   __syncthreads();
   SCHURTICK(1);
   // dense variables: U_d = X A_d S^-1 for all of them, then M_id = A_i . U_d
   for( int g = 0; g < a.ngroups; ++g )
   {
      const SmallBlock bk = a.blk[a.gblk[g]];
      const int n = bk.n, ld = bk.ld, cnt = a.gcount[g];
      const long long stride = (long long)ld * n;
      const double* Ad = a.Adense + a.gaoff[g];
      const double* Xk = a.X + bk.off;
      const double* Zk = a.Sinv + bk.off;
      // H_d = X A_d and U_d = H_d S^-1 with 4 x 2 register tiles, X and S^-1 staged in shared memory (one output per thread with both
      // operands from global memory made the two products of example_CLS, 33 matrices of order 43, the largest phase of the
      // iteration: bound by the load path of the SM).  Every output is still one sum over k in ascending order: the values are those of the
      // one-output loop, bit for bit.  Rows n .. ld - 1 of H_d and U_d are written as zeros (ld is a multiple of 4).
      dense_product<true>(n, ld, cnt, stride, Xk, Ad, a.Hd);
      SCHURTICK(2);
      dense_product<false>(n, ld, cnt, stride, a.Hd, Zk, a.Ud);
      SCHURTICK(3);
      // dense x dense pairs of the group: <A_i, U_d> as contiguous dot products over the expanded matrices (the entry list of a dense
      // matrix costs four index loads and two scattered loads of U per entry), 4 x 4 pairs per warp
      dense_pair_dots(a.denselist + a.gfirst[g], cnt, stride, Ad, a.Ud, a.M, ldm);
      SCHURTICK(4);
      // the other variables against the dense ones of the group: entry lists
      for( int p = wid; p < m * cnt; p += NT / 32 )
      {
         const int i = p / cnt, d = p % cnt;
         const int j = a.denselist[a.gfirst[g] + d];
         if( a.cls[i] == 2 && (i < j || a.E.off[a.E.varbeg[i]] == bk.off) ) continue;      // dense of this group: done above
         const double* Uj = a.Ud + (size_t)d * stride;
         double s0 = 0.0;
         for( int e = a.E.varbeg[i] + lane; e < a.E.varbeg[i + 1]; e += 32 )
         {
            if( a.E.off[e] != bk.off ) continue;
            int r = a.E.row[e], c = a.E.col[e];
            double u = Uj[(size_t)c * ld + r];
            if( r != c ) u += Uj[(size_t)r * ld + c];
            s0 += a.E.val[e] * u;
         }
         s0 = wsum(s0);
         if( lane == 0 ) a.M[(size_t)min(i, j) * ldm + max(i, j)] = s0;
      }
      __syncthreads();
   }
   SCHURTICK(5);
   // LP block: single-variable rows through the column view (one thread per variable), longer rows one after the other
   for( int j = threadIdx.x; j < m; j += NT )
   {
      double s0 = 0.0;
      for( int p = a.colbeg[j]; p < a.colbeg[j + 1]; ++p )
      {
         const int l = a.colrow[p];
         if( a.lpbeg[l + 1] - a.lpbeg[l] == 1 ) s0 += a.colval[p] * a.colval[p] * a.x[l] / a.s[l];
      }
      a.M[(size_t)j * ldm + j] += s0;
   }
   __syncthreads();
   // longer rows: the pair (i >= j) of row l belongs to the thread that meets it in the FIRST row containing both variables; that thread
   // walks the two column lists (rows ascending) and adds the terms of all common longer rows in row order.  One writer per entry:
   // no barrier between the rows (example_MkP: 30 rows of 15 variables were 30 barriers and read-modify-write round trips per
   // iteration) and the same summation order in every run.
   // Rows of up to 16 variables go one per warp, all rows side by side (example_MkP: 30 rows of 15 variables were 30 rounds of the
   // whole CTA with 225 busy threads each: 44 % of its Schur phase); longer rows take the whole CTA, one after the other (the
   // cardinality row of example_CLS, 33 variables, would keep one warp busy for 34 rounds).
   constexpr int LP_WARP_PAIRS = 256;
   for( int l = wid; l < a.nlp; l += NT / 32 )
   {
      const int b = a.lpbeg[l], cnt = a.lpbeg[l + 1] - b;
      if( cnt < 2 || cnt * cnt > LP_WARP_PAIRS ) continue;
      for( int t = lane; t < cnt * cnt; t += 32 )
      {
         const int i = a.lpind[b + t / cnt], j = a.lpind[b + t % cnt];
         if( i < j ) continue;
         int pa = a.colbeg[i], ea = a.colbeg[i + 1], pc = a.colbeg[j], ec = a.colbeg[j + 1];
         double total = 0.0;
         bool first = true, mine = false;
         while( pa < ea && pc < ec )
         {
            const int ra = a.colrow[pa], rc = a.colrow[pc];
            if( ra < rc ) ++pa;
            else if( rc < ra ) ++pc;
            else
            {
               if( a.lpbeg[ra + 1] - a.lpbeg[ra] >= 2 )
               {
                  if( first ) { first = false; mine = (ra == l); if( !mine ) break; }
                  total += (a.x[ra] / a.s[ra]) * a.colval[pa] * a.colval[pc];
               }
               ++pa; ++pc;
            }
         }
         if( mine ) a.M[(size_t)j * ldm + i] += total;
      }
   }
   for( int l = 0; l < a.nlp; ++l )
   {
      const int b = a.lpbeg[l], cnt = a.lpbeg[l + 1] - b;
      if( cnt * cnt <= LP_WARP_PAIRS ) continue;            // uniform branch
      for( int t = threadIdx.x; t < cnt * cnt; t += NT )
      {
         const int i = a.lpind[b + t / cnt], j = a.lpind[b + t % cnt];
         if( i < j ) continue;
         int pa = a.colbeg[i], ea = a.colbeg[i + 1], pc = a.colbeg[j], ec = a.colbeg[j + 1];
         double total = 0.0;
         bool first = true, mine = false;
         while( pa < ea && pc < ec )
         {
            const int ra = a.colrow[pa], rc = a.colrow[pc];
            if( ra < rc ) ++pa;
            else if( rc < ra ) ++pc;
            else
            {
               if( a.lpbeg[ra + 1] - a.lpbeg[ra] >= 2 )
               {
                  if( first ) { first = false; mine = (ra == l); if( !mine ) break; }
                  total += (a.x[ra] / a.s[ra]) * a.colval[pa] * a.colval[pc];
               }
               ++pa; ++pc;
            }
         }
         if( mine ) a.M[(size_t)j * ldm + i] += total;
      }
   }
   __syncthreads();
   SCHURTICK(6);
}

// Cholesky of M (lower, global memory) with diagonal regularisation `reg`; rdiag receives 1 / l_kk
__device__ bool cholM(const SmallArgs& a, double reg, double* rdiag, int* flag)
{
   const int m = a.m, ldm = a.ldm;
   for( int e = threadIdx.x; e < m * m; e += NT )
   {
      const int i = e % m, j = e / m;
      if( i >= j ) a.Mfac[(size_t)j * ldm + i] = a.M[(size_t)j * ldm + i] + ((i == j) ? reg : 0.0);
   }
   if( threadIdx.x == 0 ) *flag = 0;
   __syncthreads();
   double* F = a.Mfac;
   // one barrier per column (same arithmetic as the shared-memory variants: the column stays unscaled while it is used, the trailing
   // update carries 1 / d_k, all columns are scaled at the end)
   for( int k = 0; k < m; ++k )
   {
      double d = F[(size_t)k * ldm + k];
      if( !(d > 0.0) ) { if( threadIdx.x == 0 ) *flag = 1; d = 1.0; }
      const double r = frsqrt(d), rr = r * r;
      if( threadIdx.x == 0 ) rdiag[k] = r;
      for( int j = k + 1 + (threadIdx.x >> 5); j < m; j += NT / 32 )
      {
         const double lj = F[(size_t)k * ldm + j] * rr;
         for( int i = j + (threadIdx.x & 31); i < m; i += 32 ) F[(size_t)j * ldm + i] -= F[(size_t)k * ldm + i] * lj;
      }
      __syncthreads();
   }
   for( int e = threadIdx.x; e < m * m; e += NT )
   {
      const int i = e % m, k = e / m;
      if( i > k ) F[(size_t)k * ldm + i] *= rdiag[k];
      else if( i == k ) { const double d = F[(size_t)k * ldm + k]; F[(size_t)k * ldm + k] = (d > 0.0 ? d : 1.0) * rdiag[k]; }
   }
   __syncthreads();
   const bool ok = (*flag == 0);
   __syncthreads();
   return ok;
}

// v <- (L L')^-1 v with one warp: lane owns the rows r = lane, lane + 32, ... (m <= 256 -> 8 registers)
__device__ void solveM(const SmallArgs& a, const double* rdiag, double* v)
{
   __syncthreads();
   if( threadIdx.x < 32 )
   {
      const int lane = threadIdx.x, m = a.m, ldm = a.ldm;
      const double* F = a.Mfac;
      double r[SMALL_MAX_M / 32];
#pragma unroll
      for( int q = 0; q < SMALL_MAX_M / 32; ++q ) { int i = lane + 32 * q; r[q] = (i < m) ? v[i] : 0.0; }
      for( int k = 0; k < m; ++k )                 // forward substitution, column oriented
      {
         double xk = 0.0;
#pragma unroll
         for( int q = 0; q < SMALL_MAX_M / 32; ++q ) if( q == (k >> 5) ) xk = r[q];
         xk = __shfl_sync(0xffffffffu, xk, k & 31) * rdiag[k];
#pragma unroll
         for( int q = 0; q < SMALL_MAX_M / 32; ++q )
         {
            const int i = lane + 32 * q;
            if( i == k ) r[q] = xk;
            else if( i > k && i < m ) r[q] -= F[(size_t)k * ldm + i] * xk;
         }
      }
      for( int k = m - 1; k >= 0; --k )            // backward substitution with L'
      {
         double xk = 0.0;
#pragma unroll
         for( int q = 0; q < SMALL_MAX_M / 32; ++q ) if( q == (k >> 5) ) xk = r[q];
         xk = __shfl_sync(0xffffffffu, xk, k & 31) * rdiag[k];
#pragma unroll
         for( int q = 0; q < SMALL_MAX_M / 32; ++q )
         {
            const int i = lane + 32 * q;
            if( i == k ) r[q] = xk;
            else if( i < k ) r[q] -= F[(size_t)i * ldm + k] * xk;
         }
      }
#pragma unroll
      for( int q = 0; q < SMALL_MAX_M / 32; ++q ) { int i = lane + 32 * q; if( i < m ) v[i] = r[q]; }
   }
   __syncthreads();
}

// y = M x with M given by its lower triangle
__device__ void symvM(const SmallArgs& a, const double* x, double* y)
{
   const int lane = threadIdx.x & 31, m = a.m, ldm = a.ldm;
   for( int i = threadIdx.x >> 5; i < m; i += NT / 32 )
   {
      double s = 0.0;
      for( int k = lane; k < i; k += 32 ) s += a.M[(size_t)k * ldm + i] * x[k];
      for( int k = i + lane; k < m; k += 32 ) s += a.M[(size_t)i * ldm + k] * x[k];
      s = wsum(s);
      if( lane == 0 ) y[i] = s;
   }
   __syncthreads();
}

// ---- warp-level Lanczos: the matrix (n <= 64) and the Krylov vectors live in shared memory, one warp does everything ----
// Bs: n x n symmetric, row stride LDS.  Qs: (SMALL_LZ_STEPS + 2) vectors with stride LDS.  ab: alpha[32], beta[32], w[64] scratch.
// Returns (to all lanes) the safe estimate Ritz value - residual bound of the smallest eigenvalue.
// maxsteps: 8 for the predictor (its step lengths only steer the centring parameter), SMALL_LZ_STEPS for the corrector, where the run
// also ends after 8, 16 or 24 steps once the Ritz pair is good enough for a step length (residual bound below 1 % of the value, or
// the safe value above -0.5: the full step is taken anyway) - the stopping rule of lanczos_batched in eig.cu.
__device__ double lanczos_warp(int n, const double* Bs, double* Qs, double* ab, int maxsteps)
{
   const int lane = threadIdx.x & 31;
   const int r0 = lane, r1 = lane + 32;
   const bool h0 = r0 < n, h1 = r1 < n;
   double* al = ab;
   double* be = ab + SMALL_LZ_STEPS;
   double* ws = ab + 2 * SMALL_LZ_STEPS;            // 64 doubles: current w, shared between the lanes
   // (blocks of order <= 16 always run all n steps: exact values, iterate for iterate what the oracle's eigenvalue routine gives)
   const int steps = (n <= 16) ? n : min(n, min(maxsteps, SMALL_LZ_STEPS));
   double v0 = 0.0, v1 = 0.0;
   {
      unsigned h = (unsigned)r0 * 2654435761u + 12345u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      if( h0 ) v0 = 0.5 + (double)(h & 0xffffu) / 65536.0;
      h = (unsigned)r1 * 2654435761u + 12345u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      if( h1 ) v1 = 0.5 + (double)(h & 0xffffu) / 65536.0;
      double nr = wsum(v0 * v0 + v1 * v1);
      double inv = 1.0 / sqrt(nr);
      v0 *= inv; v1 *= inv;
      if( h0 ) Qs[r0] = v0;
      if( h1 ) Qs[r1] = v1;
   }
   __syncwarp();
   // smallest Ritz value of the k x k tridiagonal (32-way multisection on Sturm counts) and its residual bound
   auto ritz = [&](int k, double& th_out, double& rs_out)
   {
      double lo = 1e300, hi = -1e300;
      for( int i = 0; i < k; ++i )
      {
         double r = (i > 0 ? fabs(be[i - 1]) : 0.0) + (i < k - 1 ? fabs(be[i]) : 0.0);
         lo = fmin(lo, al[i] - r); hi = fmax(hi, al[i] + r);
      }
      const double width0 = hi - lo;
      for( int round = 0; round < 6 && (hi - lo) > 1e-9 * width0 + 1e-300; ++round )
      {
         const double xt = lo + (lane + 1) * (hi - lo) / 33.0;
         int cnt = 0;
         double dd = 1.0;
         for( int i = 0; i < k; ++i )
         {
            double b2 = (i > 0) ? be[i - 1] * be[i - 1] : 0.0;
            dd = al[i] - xt - (i > 0 ? b2 / dd : 0.0);
            if( dd == 0.0 ) dd = 1e-300;
            if( dd < 0.0 ) ++cnt;
         }
         double below = (cnt == 0) ? xt : lo;
         double above = (cnt >= 1) ? xt : hi;
   #pragma unroll
         for( int o = 16; o > 0; o >>= 1 )
         {
            below = fmax(below, __shfl_xor_sync(0xffffffffu, below, o));
            above = fmin(above, __shfl_xor_sync(0xffffffffu, above, o));
         }
         lo = below; hi = above;
      }
      const double theta = lo;
      double resid = 0.0;
      if( k < n && be[k - 1] > 1e-13 * (fabs(al[k - 1]) + 1e-300) )
      {
         double sm1 = 0.0, s0 = 1.0, nrm = 1.0, last = 1.0;
         for( int i = 0; i < k - 1; ++i )
         {
            double s1 = ((theta - al[i]) * s0 - (i > 0 ? be[i - 1] * sm1 : 0.0)) / be[i];
            sm1 = s0; s0 = s1; nrm += s1 * s1; last = s1;
            if( nrm > 1e200 ) { sm1 *= 1e-100; s0 *= 1e-100; last *= 1e-100; nrm *= 1e-200; }
         }
         resid = fabs(be[k - 1]) * fabs(last) / sqrt(nrm);
      }
         th_out = theta; rs_out = resid;
   };
   double p0 = 0.0, p1 = 0.0, bprev = 0.0;          // previous Lanczos vector (this lane's rows)
   int kdone = 0;
   for( int j = 0; j < steps; ++j )
   {
      const double* vj = Qs + j * LDS;
      double w0 = 0.0, w1 = 0.0;
      for( int k = 0; k < n; ++k )
      {
         const double vk = vj[k];
         if( h0 ) w0 += Bs[r0 * LDS + k] * vk;
         if( h1 ) w1 += Bs[r1 * LDS + k] * vk;
      }
      const double alpha = wsum(w0 * v0 + w1 * v1);
      w0 -= alpha * v0 + bprev * p0;
      w1 -= alpha * v1 + bprev * p1;
      // full re-orthogonalisation: lane q forms the inner product with vector q (q <= j <= 31), then all lanes update.  A second pass
      // only if the first one took away more than half of the squared norm ("twice is enough", Daniel-Gragg-Kaufman-Stewart: without
      // cancellation the second pass changes w by rounding errors only; it was a third of the cost of a step)
      double nrm2 = wsum(w0 * w0 + w1 * w1);
      double beta = 0.0;
      for( int pass = 0; pass < 2; ++pass )
      {
         if( h0 ) ws[r0] = w0;
         if( h1 ) ws[r1] = w1;
         __syncwarp();
         double cq = 0.0;
         if( lane <= j )
         {
            const double* vq = Qs + lane * LDS;
            for( int i = 0; i < n; ++i ) cq += ws[i] * vq[i];
         }
         // the j + 1 coefficients go through shared memory (w has been read by everybody): a broadcast load per term instead of a
         // shuffle in the dependent chain
         __syncwarp();
         if( lane <= j ) ws[lane] = cq;
         __syncwarp();
         for( int q = 0; q <= j; ++q )
         {
            const double c = ws[q];
            if( h0 ) w0 -= c * Qs[q * LDS + r0];
            if( h1 ) w1 -= c * Qs[q * LDS + r1];
         }
         __syncwarp();
         const double after2 = wsum(w0 * w0 + w1 * w1);
         beta = sqrt(after2);
         if( after2 >= 0.5 * nrm2 ) break;            // uniform over the warp (wsum leaves the same value in every lane)
         nrm2 = after2;
      }
      if( lane == 0 ) { al[j] = alpha; be[j] = beta; }
      kdone = j + 1;
      if( beta <= 1e-13 * (fabs(alpha) + bprev + 1e-300) ) break;
      const double cf = 1.0 / beta;
      p0 = v0; p1 = v1; bprev = beta;
      v0 = w0 * cf; v1 = w1 * cf;
      if( h0 ) Qs[(j + 1) * LDS + r0] = v0;
      if( h1 ) Qs[(j + 1) * LDS + r1] = v1;
      __syncwarp();
      if( n > 16 && kdone < steps && (kdone & 7) == 0 )
      {
         double th, rs;
         ritz(kdone, th, rs);
         if( rs <= 0.01 * fabs(th) || th - rs >= -0.5 ) return th - rs;
      }
   }
   __syncwarp();
   double theta, resid;
   ritz(kdone, theta, resid);
   return theta - resid;
}

// Cholesky of M + reg I entirely in shared memory (m <= 64, row stride LDMS); rdiag receives 1 / l_kk
__device__ bool cholM_smem(const SmallArgs& a, double reg, double* Ms, double* rdiag, int* flag)
{
   const int m = a.m, ldm = a.ldm;
   for( int e = threadIdx.x; e < m * m; e += NT )
   {
      const int i = e % m, j = e / m;
      if( i >= j ) Ms[i * LDMS + j] = a.M[(size_t)j * ldm + i] + ((i == j) ? reg : 0.0);
   }
   if( threadIdx.x == 0 ) *flag = 0;
   __syncthreads();
   for( int k = 0; k < m; ++k )
   {
      double d = Ms[k * LDMS + k];
      if( !(d > 0.0) ) { if( threadIdx.x == 0 ) *flag = 1; d = 1.0; }
      const double r = frsqrt(d), rr = r * r;
      if( threadIdx.x == 0 ) rdiag[k] = r;
      for( int j = k + 1 + (threadIdx.x >> 5); j < m; j += NT / 32 )
      {
         const double lj = Ms[j * LDMS + k] * rr;
         for( int i = j + (threadIdx.x & 31); i < m; i += 32 ) Ms[i * LDMS + j] -= Ms[i * LDMS + k] * lj;
      }
      __syncthreads();
   }
   for( int e = threadIdx.x; e < m * m; e += NT )
   {
      const int i = e % m, k = e / m;
      if( i > k ) Ms[i * LDMS + k] *= rdiag[k];
      else if( i == k ) { const double d = Ms[k * LDMS + k]; Ms[k * LDMS + k] = (d > 0.0 ? d : 1.0) * rdiag[k]; }
   }
   __syncthreads();
   const bool ok = (*flag == 0);
   __syncthreads();
   return ok;
}

// v <- (L L')^-1 v with L in shared memory (m <= 64): one warp, lane owns rows lane and lane + 32.
// The substitution is a chain of 2 m dependent steps (shuffle -> product -> FMA); everything that does not depend on the chain - the
// loads of column / row k and of 1 / l_kk - is written so that the compiler can issue it steps ahead (k loops split at 32: the source
// register of the shuffle is fixed inside a loop, unrolled fourfold).  Same operations in the same order as before: same bits.
__device__ __noinline__ void solveM_smem(int m, const double* Ms, const double* rdiag, double* v)
{
   __syncthreads();
   if( threadIdx.x < 32 )
   {
      const int lane = threadIdx.x;
      double r0 = (lane < m) ? v[lane] : 0.0, r1 = (lane + 32 < m) ? v[lane + 32] : 0.0;
      const double* row0 = Ms + min(lane, m - 1) * LDMS;
      const double* row1 = Ms + min(lane + 32, m - 1) * LDMS;
      const bool in0 = lane < m, in1 = lane + 32 < m;
      const int m0 = min(m, 32);
#pragma unroll 4
      for( int k = 0; k < m0; ++k )
      {
         const double xk = __shfl_sync(0xffffffffu, r0, k) * rdiag[k];
         const double l0 = row0[k], l1 = row1[k];
         if( lane == k ) r0 = xk; else if( lane > k && in0 ) r0 -= l0 * xk;
         if( in1 ) r1 -= l1 * xk;
      }
#pragma unroll 4
      for( int k = 32; k < m; ++k )
      {
         const double xk = __shfl_sync(0xffffffffu, r1, k - 32) * rdiag[k];
         const double l1 = row1[k];
         if( lane + 32 == k ) r1 = xk; else if( lane + 32 > k && in1 ) r1 -= l1 * xk;
      }
#pragma unroll 4
      for( int k = m - 1; k >= 32; --k )
      {
         const double xk = __shfl_sync(0xffffffffu, r1, k - 32) * rdiag[k];
         const double* rowk = Ms + k * LDMS;
         const double l0 = rowk[lane], l1 = rowk[min(lane + 32, k)];
         r0 -= l0 * xk;
         if( lane + 32 == k ) r1 = xk; else if( lane + 32 < k ) r1 -= l1 * xk;
      }
#pragma unroll 4
      for( int k = m0 - 1; k >= 0; --k )
      {
         const double xk = __shfl_sync(0xffffffffu, r0, k) * rdiag[k];
         const double l0 = Ms[k * LDMS + min(lane, k)];
         if( lane == k ) r0 = xk; else if( lane < k ) r0 -= l0 * xk;
      }
      if( lane < m ) v[lane] = r0;
      if( lane + 32 < m ) v[lane + 32] = r1;
   }
   __syncthreads();
}

// ---- packed variant for MSN < m <= MPK: row i of the lower triangle at Mp + i (i + 1) / 2 ----
__device__ bool cholM_pk(const SmallArgs& a, double reg, double* Mp, double* rdiag, int* flag)
{
   const int m = a.m, ldm = a.ldm;
   for( int e = threadIdx.x; e < m * m; e += NT )
   {
      const int i = e % m, j = e / m;
      if( i >= j ) Mp[i * (i + 1) / 2 + j] = a.M[(size_t)j * ldm + i] + ((i == j) ? reg : 0.0);
   }
   if( threadIdx.x == 0 ) *flag = 0;
   __syncthreads();
   // right-looking with ONE barrier per column: column k stays unscaled while it is used (every thread forms 1 / pivot itself), the
   // trailing update carries the factor 1 / d_k, and all columns are scaled by 1 / sqrt(d_k) at the end
   for( int k = 0; k < m; ++k )
   {
      double d = Mp[k * (k + 1) / 2 + k];
      if( !(d > 0.0) ) { if( threadIdx.x == 0 ) *flag = 1; d = 1.0; }
      const double r = frsqrt(d), rr = r * r;
      if( threadIdx.x == 0 ) rdiag[k] = r;
      // warp per column j of the trailing block, lanes over the rows i >= j: no integer division, lower triangle only
      for( int j = k + 1 + (threadIdx.x >> 5); j < m; j += NT / 32 )
      {
         const double lj = Mp[j * (j + 1) / 2 + k] * rr;
         for( int i = j + (threadIdx.x & 31); i < m; i += 32 ) Mp[i * (i + 1) / 2 + j] -= Mp[i * (i + 1) / 2 + k] * lj;
      }
      __syncthreads();
   }
   for( int e = threadIdx.x; e < m * m; e += NT )
   {
      const int i = e % m, k = e / m;
      if( i > k ) Mp[i * (i + 1) / 2 + k] *= rdiag[k];
      else if( i == k ) { const double d = Mp[k * (k + 1) / 2 + k]; Mp[k * (k + 1) / 2 + k] = (d > 0.0 ? d : 1.0) * rdiag[k]; }
   }
   __syncthreads();
   const bool ok = (*flag == 0);
   __syncthreads();
   return ok;
}

// v <- (L L')^-1 v with the packed factor: one warp, lane owns the rows lane + 32 q.  Written like solveM_smem: the k loops are split
// at the multiples of 32 (static source register of the shuffle), the row pointers are formed once, the loads run ahead of the chain.
__device__ __noinline__ void solveM_pk(int m, const double* Mp, const double* rdiag, double* v)
{
   __syncthreads();
   if( threadIdx.x < 32 )
   {
      constexpr int Q = (MPK + 31) / 32;
      const int lane = threadIdx.x;
      double r[Q];
      const double* row[Q];
      bool in[Q];
#pragma unroll
      for( int q = 0; q < Q; ++q )
      {
         const int i = lane + 32 * q, ic = min(i, m - 1);
         in[q] = (i < m);
         r[q] = in[q] ? v[i] : 0.0;
         row[q] = Mp + ic * (ic + 1) / 2;
      }
      // forward substitution, column k of L: entries (i, k), i > k
#pragma unroll
      for( int kq = 0; kq < Q; ++kq )
      {
         const int kend = min(32, m - 32 * kq);
#pragma unroll 4
         for( int kl = 0; kl < kend; ++kl )
         {
            const int k = 32 * kq + kl;
            const double xk = __shfl_sync(0xffffffffu, r[kq], kl) * rdiag[k];
#pragma unroll
            for( int q = kq; q < Q; ++q )
            {
               const double l = row[q][min(k, lane + 32 * q)];
               if( q == kq ) { if( lane == kl ) r[q] = xk; else if( lane > kl && in[q] ) r[q] -= l * xk; }
               else if( in[q] ) r[q] -= l * xk;
            }
         }
      }
      // backward substitution with L': row k of L, entries (k, i), i < k (contiguous)
#pragma unroll
      for( int kq = Q - 1; kq >= 0; --kq )
      {
         const int kend = min(32, m - 32 * kq);
#pragma unroll 4
         for( int kl = kend - 1; kl >= 0; --kl )
         {
            const int k = 32 * kq + kl;
            const double xk = __shfl_sync(0xffffffffu, r[kq], kl) * rdiag[k];
            const double* rowk = Mp + k * (k + 1) / 2;
#pragma unroll
            for( int q = 0; q <= kq; ++q )
            {
               const double l = rowk[min(lane + 32 * q, k)];
               if( q == kq ) { if( lane == kl ) r[q] = xk; else if( lane < kl ) r[q] -= l * xk; }
               else r[q] -= l * xk;
            }
         }
      }
#pragma unroll
      for( int q = 0; q < Q; ++q ) { const int i = lane + 32 * q; if( i < m ) v[i] = r[q]; }
   }
   __syncthreads();
}

// (Measured and NOT kept: inverting the shared-memory factor of M in place once per iteration and solving by two matrix-vector products.
// The substitution by one warp is a chain of 2 m dependent steps - 106 k cycles per solve at m = 96, four solves per iteration - and
// the products took 0.3 M instead of 5.7 M cycles per example_MkP relaxation; but the in-place inversion cost 4.9 M, a frontier of
// example_TT (m = 27, four CTAs per SM: throughput, not latency) lost 13 %, and the penalty formulations of the boundary tests
// (Gamma = 1e4 .. 1e6) stopped converging: the substitution is backward stable, the explicit inverse is not.)
// dynamic shared memory of the body (layout at the top of ipm_small_body)
constexpr size_t SMALL_CTL_BYTES = (8 * sizeof(double) + sizeof(Ctl) + 16 + 15) / 16 * 16;      // lam2, Ctl, flag
constexpr size_t SMALL_MSH_OFF = 2 * VMAXN * LDS + 32 + VMAXN + SMALL_MAX_M + 3 * SMALL_LZ_STEPS + 8 + 2 * (SMALL_LZ_STEPS + 2) * LDS
   + 2 * (2 * SMALL_LZ_STEPS + 64) + SMALL_CTL_BYTES / sizeof(double);                             // in doubles
constexpr size_t SMALL_SMEM = (SMALL_MSH_OFF + MSN * LDMS) * sizeof(double) + 64;

// the complete solve of ONE relaxation by the calling CTA (shared by the one-relaxation kernel and the frontier-batch kernel)
__device__ __forceinline__ void ipm_small_body(const SmallArgs& a)
{
   extern __shared__ __align__(16) double smem[];
   double* sh = smem;                               // 64 x 65
   double* sh2 = sh + VMAXN * LDS;                  // 64 x 65
   double* red = sh2 + VMAXN * LDS;                 // 32
   double* rdiag = red + 32;                        // 64
   double* rdiagM = rdiag + VMAXN;                  // 256
   double* al = rdiagM + SMALL_MAX_M;               // 32 + 32 + 40
   double* be = al + SMALL_LZ_STEPS;
   double* coef = be + SMALL_LZ_STEPS;
   double* lzq = coef + SMALL_LZ_STEPS + 8;         // 2 x (SMALL_LZ_STEPS + 2) x 65 : Krylov vectors of the two warp-level Lanczos runs
   double* lzab = lzq + 2 * (SMALL_LZ_STEPS + 2) * LDS;   // 2 x (32 + 32 + 64)
   double* lam2 = lzab + 2 * (2 * SMALL_LZ_STEPS + 64);   // 2 results
   Ctl* c = reinterpret_cast<Ctl*>(lam2 + 8);
   int* flag = reinterpret_cast<int*>(reinterpret_cast<char*>(lam2) + SMALL_CTL_BYTES - 16);
   // 64 x 65 : factor of the Schur complement when m <= 64.  LAST of the kernel's own buffers: a frontier launch that stages no work
   // space puts the packed factor of the nodes with m > 64 at the same place (a node needs one of the two), see batch_plan
   double* Msh = smem + SMALL_MSH_OFF;
   const bool msmall = (a.m <= MSN);
   double* const Mpk = smem + a.mpk_off;                                   // packed factor, reserved by the launch when mpk_off > 0
   const bool mpacked = !msmall && a.mpk_off > 0 && a.m <= MPK;
   const int tid = threadIdx.x;
   const int m = a.m, nb = a.nb, nlp = a.nlp;
   const double inftol = 1e-8;
   long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
   long long tq = clock64();
#define TICK(k) do { long long _n = clock64(); pc[k] += _n - tq; tq = _n; } while( 0 )
   // sub-phases of the directions: compiled in with -DSDPK_SUBTICKS only (eight more 64-bit counters cost the 256-thread instantiation,
   // which lives on 64 registers per thread, 10 % of its frontier throughput)
#ifdef SDPK_SUBTICKS
   long long ps[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tq2 = 0;
#define SUBTICK(k) do { long long _n = clock64(); ps[k] += _n - tq2; tq2 = _n; } while( 0 )
#define SUBTICK_START() do { tq2 = clock64(); } while( 0 )
#else
#define SUBTICK(k) do { } while( 0 )
#define SUBTICK_START() do { } while( 0 )
#endif

   if( tid == 0 )
   {
      c->iter = 0; c->stall = 0; c->backtracks = 0; c->phase = SDPCUDA_NOINFO; c->stop = SDPCUDA_STOP_ITERLIMIT; c->done = 0;
      c->pfeasever = 0; c->dfeasever = 0; c->bestmerit = 1e300; c->lastap = 0.0; c->lastad = 0.0; c->xfail = 0;
      c->mu = 0; c->pobj = 0; c->dobj = 0; c->relgap = 1e30; c->pinf = 1e30; c->dinf = 1e30;
   }
   if( a.selfinit )
   {
      // cold start and expanded dense matrices written by the CTA itself (what run_ipm / upload_problem do with memsets,
      // add_diagonal and scatter_dense launches for a single relaxation)
      for( long long i = tid; i < a.arena; i += NT ) { a.X[i] = 0.0; a.S[i] = 0.0; }
      double* AdW = const_cast<double*>(a.Adense);       // written only here, read-only for the iteration
      for( long long i = tid; i < a.adense_total; i += NT ) AdW[i] = 0.0;
      for( int l = tid; l < nlp; l += NT ) { a.x[l] = a.xil; a.s[l] = a.etal; }
      for( int j = tid; j < m; j += NT ) a.y[j] = 0.0;
      __syncthreads();
      for( int k = 0; k < nb; ++k )
      {
         const SmallBlock bk = a.blk[k];
         for( int i = tid; i < bk.n; i += NT )
         {
            a.X[bk.off + (long long)i * bk.ld + i] = a.xi[k];
            a.S[bk.off + (long long)i * bk.ld + i] = a.eta[k];
         }
      }
      for( int g = 0; g < a.ngroups; ++g )
      {
         const SmallBlock bk = a.blk[a.gblk[g]];
         const long long stride = (long long)bk.ld * bk.n;
         for( int d = 0; d < a.gcount[g]; ++d )
         {
            const int j = a.denselist[a.gfirst[g] + d];
            double* Ad = AdW + a.gaoff[g] + (long long)d * stride;
            for( int e = a.E.varbeg[j] + tid; e < a.E.varbeg[j + 1]; e += NT )
            {
               const int r = a.E.row[e], cc = a.E.col[e];
               const double v = a.E.val[e];
               Ad[(long long)cc * bk.ld + r] = v;
               Ad[(long long)r * bk.ld + cc] = v;
            }
         }
      }
      __syncthreads();
   }
   for( long long i = tid; i < a.arena; i += NT ) { a.dX[i] = 0.0; a.dS[i] = 0.0; }
   __syncthreads();

   for( ; ; )
   {
      // ---------------- residuals and statistics ----------------
      assemble(a, a.y, 1.0, a.K);
      double sRd = 0.0, sXS = 0.0;
      for( long long i = tid; i < a.arena; i += NT )
      {
         double sv = a.S[i], r = a.K[i] - sv;
         a.Rd[i] = r;
         sRd += r * r;
         sXS += a.X[i] * sv;
      }
      sRd = bsum(sRd, red); sXS = bsum(sXS, red);
      applyA(a, a.X, a.AX);
      lpcols(a, a.x, a.DTx, false);
      double s6 = 0, s7 = 0, s8 = 0, m17 = 0;
      for( int j = tid; j < m; j += NT )
      {
         double ax = a.AX[j] + a.DTx[j];
         double r = a.b[j] - ax;
         a.rp[j] = r;
         s6 += r * r; s7 += ax * ax; s8 += a.b[j] * a.y[j]; m17 = fmax(m17, fabs(r));
      }
      s6 = bsum(s6, red); s7 = bsum(s7, red); s8 = bsum(s8, red); m17 = bmax(m17, red);
      lprows(a, a.y, a.Dy);
      double s2 = 0, s3 = 0, s4 = 0, s5 = 0, m16 = 0;
      for( int l = tid; l < nlp; l += NT )
      {
         double d = a.Dy[l], r = d - a.lprhs[l] - a.s[l];
         a.rdlp[l] = r;
         s2 += r * r; s3 += a.x[l] * a.s[l]; s4 += a.lprhs[l] * a.x[l];
         double hh = d - a.s[l]; s5 += hh * hh; m16 = fmax(m16, fabs(r));
      }
      s2 = bsum(s2, red); s3 = bsum(s3, red); s4 = bsum(s4, red); s5 = bsum(s5, red); m16 = bmax(m16, red);
      double cx = 0.0, crd = 0.0;
      for( int e = tid; e < a.cnnz; e += NT )
      {
         long long p = a.cpos[e], q = a.cmirror[e];
         double cv = a.cval[e];
         cx += cv * (a.X[p] + (p != q ? a.X[q] : 0.0));
         crd += cv * (a.Rd[p] + (p != q ? a.Rd[q] : 0.0));
      }
      cx = bsum(cx, red); crd = bsum(crd, red);

      TICK(0);
      // ---------------- factorisations of S and X (their failure means the last step left the cone) ----------------
      bool okS = true, okX = true;
      for( int k = 0; k < nb; ++k )
      {
         const SmallBlock bk = a.blk[k];
         okS = cholinv(bk.n, bk.ld, a.S + bk.off, a.L + bk.off, a.Linv + bk.off, sh, sh2, rdiag, flag) && okS;
         okX = cholinv(bk.n, bk.ld, a.X + bk.off, a.LX + bk.off, a.LXinv + bk.off, sh, sh2, rdiag, flag) && okX;
      }
      if( !okS || !okX )
      {
         if( tid == 0 )
         {
            if( c->iter == 0 || c->backtracks >= 8 ) { c->stop = SDPCUDA_STOP_NUMERICS; c->done = 1; }
            ++c->backtracks;
         }
         __syncthreads();
         if( c->done ) break;
         if( !okX )
         {
            const double h = -0.5 * c->lastap;
            for( long long i = tid; i < a.arena; i += NT ) a.X[i] += h * a.dX[i];
            for( int l = tid; l < nlp; l += NT ) a.x[l] += h * a.dx[l];
         }
         if( !okS )
         {
            const double h = -0.5 * c->lastad;
            for( long long i = tid; i < a.arena; i += NT ) a.S[i] += h * a.dS[i];
            for( int l = tid; l < nlp; l += NT ) a.s[l] += h * a.ds[l];
            for( int j = tid; j < m; j += NT ) a.y[j] += h * a.dy[j];
         }
         __syncthreads();
         if( tid == 0 ) { if( !okX ) c->lastap *= 0.5; if( !okS ) c->lastad *= 0.5; }
         __syncthreads();
         continue;
      }

      TICK(1);
      // ---------------- termination tests (same rules as ipm.cu) ----------------
      if( tid == 0 )
      {
         c->backtracks = 0;
         const double nrd2 = sRd + s2, xs = sXS + s3;
         c->pobj = cx + s4;
         c->dobj = s8;
         c->mu = a.N > 0 ? xs / a.N : 0.0;
         c->pinf = sqrt(s6) / (1.0 + a.normb);
         c->dinf = sqrt(nrd2) / (1.0 + a.normC);
         const double dinfabs = fmax(sqrt(sRd), m16), pinfabs = m17;
         c->relgap = fabs(c->pobj - c->dobj) / fmax(1.0, 0.5 * (fabs(c->pobj) + fabs(c->dobj)));
         const double rayd = sqrt(fmax(0.0, sRd + 2.0 * crd + a.normCsdp2 + s5));
         const bool pfeas = c->pinf <= a.feastol && pinfabs <= fmax(a.feastol, 1e-9 * (1 + a.normb));
         const bool dfeas = c->dinf <= a.feastol && dinfabs <= a.feastol;
         c->pfeasever |= pfeas ? 1 : 0;
         c->dfeasever |= dfeas ? 1 : 0;
         c->phase = pfeas ? (dfeas ? SDPCUDA_PDFEAS : SDPCUDA_PFEAS) : (dfeas ? SDPCUDA_DFEAS : SDPCUDA_NOINFO);
         c->rdzero = (sRd <= 1e-28 * (1.0 + a.normCsdp2)) ? 1 : 0;
         if( a.verbose )
            printf("  [cuda-1cta] it %3d  pobj % .10e  dobj % .10e  gap %.2e  pinf %.2e  dinf %.2e  mu %.2e\n", c->iter, c->pobj, c->dobj, c->relgap, c->pinf, c->dinf, c->mu);
         if( pfeas && dfeas && c->relgap <= a.gaptol && (a.absgaptol <= 0 || fabs(c->pobj - c->dobj) <= a.absgaptol) )
         { c->phase = SDPCUDA_PDOPT; c->stop = SDPCUDA_STOP_CONVERGED; c->done = 1; }
         else if( c->pobj > 0 && sqrt(s7) / c->pobj < inftol )
         { c->phase = c->pfeasever ? SDPCUDA_PFEAS_DINF : SDPCUDA_DINF; c->stop = SDPCUDA_STOP_INFEASCERT; c->done = 1; }
         else if( c->dfeasever && c->dobj < 0 && rayd / (-c->dobj) < inftol )
         { c->phase = SDPCUDA_PINF_DFEAS; c->stop = SDPCUDA_STOP_INFEASCERT; c->done = 1; }
         else if( pfeas && a.objlimit < 1e20 && c->pobj > a.objlimit )
         { c->phase = SDPCUDA_PUNBD; c->stop = SDPCUDA_STOP_OBJLIMIT; c->done = 1; }
         else if( c->iter >= a.maxiter ) { c->stop = SDPCUDA_STOP_ITERLIMIT; c->done = 1; }
         else
         {
            double merit = fmax(c->relgap, fmax(c->pinf, c->dinf));
            if( merit < 0.9 * c->bestmerit ) { c->bestmerit = merit; c->stall = 0; }
            else if( ++c->stall >= 15 ) { c->stop = SDPCUDA_STOP_NUMERICS; c->done = 1; }
         }
      }
      __syncthreads();
      if( c->done ) break;
      const bool rdzero = c->rdzero != 0;
      const double mu = c->mu;

      // ---------------- S^-1, Schur complement, its factorisation ----------------
      for( int k = 0; k < nb; ++k )
      {
         const SmallBlock bk = a.blk[k];
         gemmb(bk.n, bk.ld, true, false, 1.0, a.Linv + bk.off, a.Linv + bk.off, 0.0, a.Sinv + bk.off);
      }
      TICK(2);
      schur(a);
      TICK(3);
      {
         double maxd = 0.0;
         for( int j = tid; j < m; j += NT ) maxd = fmax(maxd, a.M[(size_t)j * a.ldm + j]);
         maxd = bmax(maxd, red);
         double reg = 0.0;
         bool mok = false;
         for( int tries = 0; tries < 8 && !mok; ++tries )
         {
            mok = msmall ? cholM_smem(a, reg, Msh, rdiagM, flag) : (mpacked ? cholM_pk(a, reg, Mpk, rdiagM, flag) : cholM(a, reg, rdiagM, flag));
            if( !mok ) reg = (reg == 0.0) ? 1e-14 * fmax(maxd, 1e-300) : reg * 100.0;
         }
         if( !mok )
         {
            if( tid == 0 ) { c->stop = SDPCUDA_STOP_NUMERICS; c->done = 1; }
            __syncthreads();
            break;
         }
      }

      TICK(4);
      // ---------------- predictor and corrector ----------------
      bool failed = false;
      for( int pass = 0; pass < 2; ++pass )
      {
         double* oX = pass == 0 ? a.dXa : a.dX;
         double* oS = pass == 0 ? a.dSa : a.dS;
         double* ox = pass == 0 ? a.dxa : a.dx;
         double* os = pass == 0 ? a.dsa : a.ds;
         const double sigmamu = (pass == 1) ? c->sigma * mu : 0.0;
         SUBTICK_START();
         // K = sym((sigma mu I - dXa dSa - X Rd) S^-1) - X
         const bool haveT = (!rdzero) || pass == 1;
         for( int k = 0; k < nb; ++k )
         {
            const SmallBlock bk = a.blk[k];
            if( haveT )
            {
               if( !rdzero ) gemmb(bk.n, bk.ld, false, false, -1.0, a.X + bk.off, a.Rd + bk.off, 0.0, a.T1 + bk.off);
               if( pass == 1 )
               {
                  gemmb(bk.n, bk.ld, false, false, -1.0, a.dXa + bk.off, a.dSa + bk.off, rdzero ? 0.0 : 1.0, a.T1 + bk.off);
                  for( int i = tid; i < bk.n; i += NT ) a.T1[bk.off + (size_t)i * bk.ld + i] += sigmamu;
                  __syncthreads();
               }
               gemmb(bk.n, bk.ld, false, false, 1.0, a.T1 + bk.off, a.Sinv + bk.off, 0.0, a.K + bk.off);
               symavg(bk.n, bk.ld, a.K + bk.off, a.X + bk.off);
            }
            else
            {
               for( int e = tid; e < bk.n * bk.ld; e += NT ) a.K[bk.off + e] = -a.X[bk.off + e];
               __syncthreads();
            }
         }
         for( int l = tid; l < nlp; l += NT )
         {
            double cc = -a.x[l] * a.rdlp[l];
            if( pass == 1 ) cc += sigmamu - a.dxa[l] * a.dsa[l];
            a.klp[l] = cc / a.s[l] - a.x[l];
         }
         __syncthreads();
         SUBTICK(0);
         applyA(a, a.K, a.g);
         lpcols(a, a.klp, a.g, true);
         for( int j = tid; j < m; j += NT ) { a.g[j] -= a.rp[j]; a.dy[j] = a.g[j]; }
         SUBTICK(1);
         if( msmall ) solveM_smem(m, Msh, rdiagM, a.dy); else if( mpacked ) solveM_pk(m, Mpk, rdiagM, a.dy); else solveM(a, rdiagM, a.dy);
         SUBTICK(6);
         symvM(a, a.dy, a.tm1);
         SUBTICK(7);
         for( int j = tid; j < m; j += NT ) a.tm1[j] = a.g[j] - a.tm1[j];
         if( msmall ) solveM_smem(m, Msh, rdiagM, a.tm1); else if( mpacked ) solveM_pk(m, Mpk, rdiagM, a.tm1); else solveM(a, rdiagM, a.tm1);
         for( int j = tid; j < m; j += NT ) a.dy[j] += a.tm1[j];
         __syncthreads();
         SUBTICK(2);
         // dS = A'dy (+ Rd), dX = K - sym(X (A'dy) S^-1)
         assemble(a, a.dy, 0.0, oS);
         SUBTICK(3);
         for( int k = 0; k < nb; ++k )
         {
            const SmallBlock bk = a.blk[k];
            gemmb(bk.n, bk.ld, false, false, 1.0, a.X + bk.off, oS + bk.off, 0.0, a.T1 + bk.off);
            gemmb(bk.n, bk.ld, false, false, 1.0, a.T1 + bk.off, a.Sinv + bk.off, 0.0, a.T2 + bk.off);
            symavg(bk.n, bk.ld, a.T2 + bk.off, nullptr);
         }
         for( long long i = tid; i < a.arena; i += NT )
         {
            oX[i] = a.K[i] - a.T2[i];
            if( !rdzero ) oS[i] += a.Rd[i];
         }
         __syncthreads();
         SUBTICK(4);
         lprows(a, a.dy, a.Ddy);
         double rp1 = -1e300, rd1 = -1e300;
         for( int l = tid; l < nlp; l += NT )
         {
            const double ddy = a.Ddy[l];
            const double vx = a.klp[l] - a.x[l] / a.s[l] * ddy;
            const double vs = ddy + a.rdlp[l];
            ox[l] = vx; os[l] = vs;
            if( vx < 0.0 ) rp1 = fmax(rp1, a.x[l] / vx);
            if( vs < 0.0 ) rd1 = fmax(rd1, a.s[l] / vs);
         }
         rp1 = bmax(rp1, red); rd1 = bmax(rd1, red);
         double apmax = (rp1 > -1e299) ? -rp1 : 1e30, admax = (rd1 > -1e299) ? -rd1 : 1e30;
         SUBTICK(5);
         TICK(5);
         // SDP step lengths from lambda_min(LXinv dX LXinv') and lambda_min(Linv dS Linv')
         for( int k = 0; k < nb; ++k )
         {
            const SmallBlock bk = a.blk[k];
            gemmb(bk.n, bk.ld, false, false, 1.0, a.LXinv + bk.off, oX + bk.off, 0.0, a.T1 + bk.off);
            gemmb(bk.n, bk.ld, false, true, 1.0, a.T1 + bk.off, a.LXinv + bk.off, 0.0, a.T2 + bk.off);
            gemmb(bk.n, bk.ld, false, false, 1.0, a.Linv + bk.off, oS + bk.off, 0.0, a.T1 + bk.off);
            gemmb(bk.n, bk.ld, false, true, 1.0, a.T1 + bk.off, a.Linv + bk.off, 0.0, a.K + bk.off);
            // symmetrised copies into shared memory (sh: X side, sh2: S side), then one warp per matrix
            for( int e = tid; e < bk.n * bk.n; e += NT )
            {
               const int i = e % bk.n, j = e / bk.n;
               sh[i * LDS + j] = 0.5 * (a.T2[bk.off + (size_t)j * bk.ld + i] + a.T2[bk.off + (size_t)i * bk.ld + j]);
               sh2[i * LDS + j] = 0.5 * (a.K[bk.off + (size_t)j * bk.ld + i] + a.K[bk.off + (size_t)i * bk.ld + j]);
            }
            __syncthreads();
            if( tid < 64 )
            {
               const int wv = tid >> 5;
               const double lv = lanczos_warp(bk.n, wv == 0 ? sh : sh2, lzq + wv * (SMALL_LZ_STEPS + 2) * LDS, lzab + wv * (2 * SMALL_LZ_STEPS + 64), pass == 0 ? 8 : SMALL_LZ_STEPS);
               if( (tid & 31) == 0 ) lam2[wv] = lv;
            }
            __syncthreads();
            const double lx = lam2[0], ls = lam2[1];
            __syncthreads();
            if( !(lx == lx) || !(ls == ls) ) failed = true;
            if( lx < -1e-300 ) apmax = fmin(apmax, -1.0 / lx);
            if( ls < -1e-300 ) admax = fmin(admax, -1.0 / ls);
         }
         TICK(6);
         if( failed ) break;
         if( pass == 0 )
         {
            const double ap = fmin(1.0, 0.98 * apmax), ad = fmin(1.0, 0.98 * admax);
            double sa = 0.0;
            for( long long i = tid; i < a.arena; i += NT ) sa += (a.X[i] + ap * oX[i]) * (a.S[i] + ad * oS[i]);
            for( int l = tid; l < nlp; l += NT ) sa += (a.x[l] + ap * ox[l]) * (a.s[l] + ad * os[l]);
            sa = bsum(sa, red);
            if( tid == 0 )
            {
               const double mua = a.N > 0 ? sa / a.N : 0.0;
               const double ratio = mu > 0 ? fmax(0.0, mua / mu) : 0.0;
               const double mn = fmin(ap, ad);
               const double expo = (mu > 1e-6) ? fmax(1.0, 3.0 * mn * mn) : 1.0;
               double sg = fmin(1.0, pow(ratio, expo));
               if( a.setting >= 3 ) sg = fmax(sg, 0.1);
               c->sigma = sg; c->ap = ap; c->ad = ad;
            }
            __syncthreads();
         }
         else
         {
            if( tid == 0 )
            {
               const double gamma = a.gammabase + (0.99 - a.gammabase) * fmin(c->ap, c->ad);
               c->ap = fmin(1.0, gamma * apmax); c->ad = fmin(1.0, gamma * admax);
            }
            __syncthreads();
         }
      }
      if( failed || (c->ap < 1e-8 && c->ad < 1e-8) )
      {
         __syncthreads();
         if( tid == 0 ) { c->stop = SDPCUDA_STOP_NUMERICS; c->done = 1; }
         __syncthreads();
         break;
      }
      const double ap = c->ap, ad = c->ad;
      for( long long i = tid; i < a.arena; i += NT ) { a.X[i] += ap * a.dX[i]; a.S[i] += ad * a.dS[i]; }
      for( int l = tid; l < nlp; l += NT ) { a.x[l] += ap * a.dx[l]; a.s[l] += ad * a.ds[l]; }
      for( int j = tid; j < m; j += NT ) a.y[j] += ad * a.dy[j];
      __syncthreads();
      if( tid == 0 ) { c->lastap = ap; c->lastad = ad; ++c->iter; }
      __syncthreads();
   }
   __syncthreads();
   if( tid == 0 )
   {
      SmallResult r;
      r.phase = c->phase; r.stop = c->stop; r.iterations = c->iter; r.backtracks = c->backtracks;
      r.pobj = c->pobj; r.dobj = c->dobj; r.relgap = c->relgap; r.pinf = c->pinf; r.dinf = c->dinf; r.mu = c->mu;
      *a.out = r;
#ifdef SDPK_SCHURTICKS
      if( a.verbose >= 2 )
         printf("  [schur cycles] zero + light pairs %lld | heavy pairs %lld | X A_d %lld | H_d S^-1 %lld | dense pair dots %lld | entry lists x dense %lld | LP %lld\n",
            g_schur_ticks[0], g_schur_ticks[1], g_schur_ticks[2], g_schur_ticks[3], g_schur_ticks[4], g_schur_ticks[5], g_schur_ticks[6]);
#endif
      if( a.verbose >= 2 )
         printf("  [cuda-1cta cycles] resid %lld | fact S,X %lld | tests+Sinv %lld | schur %lld | chol M %lld | directions %lld | step lengths %lld\n",
            pc[0], pc[1], pc[2], pc[3], pc[4], pc[5], pc[6]);
#ifdef SDPK_SUBTICKS
      if( a.verbose >= 2 )
         printf("  [cuda-1cta cycles, directions] K %lld | A(K) + D'k %lld | first solve with M %lld | M dy %lld | second solve %lld | A'dy %lld | dX %lld | LP part %lld (m = %d, factor of M: %s)\n",
            ps[0], ps[1], ps[6], ps[7], ps[2], ps[3], ps[4], ps[5], a.m, msmall ? "shared memory" : (mpacked ? "packed in shared memory" : "global memory"));
#endif
   }
}

#ifndef SDPK_VARIANT_TINY
__global__ void __launch_bounds__(NT, MINB)
ipm_small_kernel(const SmallArgs a)
{
   ipm_small_body(a);
}
#endif

// frontier batch: CTA i solves the relaxation described by all[i] (its own device buffers); the descriptor is staged in shared
// memory once, the CTAs never communicate
__global__ void __launch_bounds__(NT, MINB)
ipm_small_batch_kernel(const SmallArgs* __restrict__ all)
{
   __shared__ SmallArgs sa;
   static_assert(sizeof(SmallArgs) % sizeof(int) == 0, "descriptor is copied in 4-byte words");
   const int* src = reinterpret_cast<const int*>(all + blockIdx.x);
   int* dst = reinterpret_cast<int*>(&sa);
   for( int i = threadIdx.x; i < (int)(sizeof(SmallArgs) / sizeof(int)); i += NT ) dst[i] = src[i];
   __syncthreads();
   __shared__ double* orig[4];                     // X, S, x, s as the host bound them (global memory)
   if( threadIdx.x == 0 ) { orig[0] = sa.X; orig[1] = sa.S; orig[2] = sa.x; orig[3] = sa.s; }
   if( sa.stage_doubles > 0 )
   {
      // the head of the node's work space moves into shared memory (behind the SMALL_SMEM bytes of the body): zero it like the host
      // zeroes the global work space, then point every array that lies completely inside the staged head to its shared copy
      extern __shared__ __align__(16) double smem[];
      double* stage = smem + (SMALL_SMEM + sizeof(double) - 1) / sizeof(double);
      for( long long i = threadIdx.x; i < sa.stage_doubles; i += NT ) stage[i] = 0.0;
      if( threadIdx.x == 0 )
      {
         double** ptrs[] = {&sa.X, &sa.S, &sa.Sinv, &sa.L, &sa.Linv, &sa.LX, &sa.LXinv, &sa.dX, &sa.dS, &sa.dXa, &sa.dSa, &sa.K, &sa.T1, &sa.T2, &sa.Rd,
                            &sa.dy, &sa.g, &sa.rp, &sa.AX, &sa.DTx, &sa.tm1, &sa.tm2,
                            &sa.x, &sa.s, &sa.dx, &sa.ds, &sa.dxa, &sa.dsa, &sa.klp, &sa.rdlp, &sa.Dy, &sa.Ddy, &sa.M, &sa.Mfac, &sa.Hd, &sa.Ud};
         for( double** pp : ptrs )
         {
            const long long off = *pp - sa.workbase;
            if( off >= 0 && off < sa.stage_doubles ) *pp = stage + off;
         }
         const long long offa = sa.Adense - sa.workbase;
         if( offa >= 0 && offa < sa.stage_doubles ) sa.Adense = stage + offa;
      }
      __syncthreads();
   }
   ipm_small_body(sa);
   if( sa.copyback && sa.stage_doubles > 0 )
   {
      __syncthreads();
      if( sa.X != orig[0] ) for( long long i = threadIdx.x; i < sa.arena; i += NT ) orig[0][i] = sa.X[i];
      if( sa.S != orig[1] ) for( long long i = threadIdx.x; i < sa.arena; i += NT ) orig[1][i] = sa.S[i];
      if( sa.x != orig[2] ) for( int l = threadIdx.x; l < sa.nlp; l += NT ) orig[2][l] = sa.x[l];
      if( sa.s != orig[3] ) for( int l = threadIdx.x; l < sa.nlp; l += NT ) orig[3][l] = sa.s[l];
   }
}

} // namespace

#ifndef SDPK_VARIANT_TINY
cudaError_t launch_ipm_small(cudaStream_t st, const SmallArgs& a)
{
   static bool configured[64] = {false};
   int dev = 0;
   SDPK_CUDA_CHECK( cudaGetDevice(&dev) );
   // Schur complements of order 65 .. SMALL_MPK: room for the packed factor behind the kernel's own buffers
   const size_t off = (SMALL_SMEM + 15) / 16 * 16;
   const bool pk = (a.m > MSN && a.m <= MPK);
   const size_t total = pk ? off + small_mpk_bytes(MPK) : SMALL_SMEM;
   if( !configured[dev & 63] )
   {
      SDPK_CUDA_CHECK( cudaFuncSetAttribute(ipm_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(off + small_mpk_bytes(MPK))) );
      configured[dev & 63] = true;
   }
   SmallArgs b = a;
   b.mpk_off = pk ? (long long)(off / sizeof(double)) : 0;
   ipm_small_kernel<<<1, NT, total, st>>>(b);
   count_launch();
   return cudaGetLastError();
}

size_t ipm_small_smem_bytes() { return SMALL_SMEM; }
size_t ipm_small_msh_offset_bytes() { return SMALL_MSH_OFF * sizeof(double); }

cudaError_t launch_ipm_small_batch(cudaStream_t st, int count, const SmallArgs* dev_args, size_t stage_bytes)
{
   static bool configured[64] = {false};
   int dev = 0;
   if( count <= 0 ) return cudaSuccess;
   SDPK_CUDA_CHECK( cudaGetDevice(&dev) );
   if( !configured[dev & 63] )
   {
      // up to the 227 KB a CTA may have (1.5 KB of them static: the staged descriptor)
      SDPK_CUDA_CHECK( cudaFuncSetAttribute(ipm_small_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024) );
      configured[dev & 63] = true;
   }
   if( SMALL_SMEM + stage_bytes > 225 * 1024 ) return cudaErrorInvalidValue;
   ipm_small_batch_kernel<<<count, NT, SMALL_SMEM + stage_bytes, st>>>(dev_args);
   count_launch();
   return cudaGetLastError();
}

#else

size_t ipm_tiny_smem_bytes() { return SMALL_SMEM; }
size_t ipm_tiny_msh_offset_bytes() { return SMALL_MSH_OFF * sizeof(double); }

cudaError_t launch_ipm_tiny_batch(cudaStream_t st, int count, const SmallArgs* dev_args, size_t stage_bytes)
{
   static bool configured[64] = {false};
   int dev = 0;
   if( count <= 0 ) return cudaSuccess;
   SDPK_CUDA_CHECK( cudaGetDevice(&dev) );
   if( !configured[dev & 63] )
   {
      SDPK_CUDA_CHECK( cudaFuncSetAttribute(ipm_tiny_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024) );
      configured[dev & 63] = true;
   }
   if( SMALL_SMEM + stage_bytes > 225 * 1024 ) return cudaErrorInvalidValue;
   ipm_tiny_batch_kernel<<<count, NT, SMALL_SMEM + stage_bytes, st>>>(dev_args);
   count_launch();
   return cudaGetLastError();
}

#endif

} // namespace sdpk
