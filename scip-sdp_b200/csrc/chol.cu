// chol.cu — blocked Cholesky factorisation, triangular inverse and triangular solves on FP64 tensor-core GEMMs.
//
// Recursive blocking: a matrix of order n is split at n1 (a multiple of CHOL_NB),
//      A11 = L11 L11'                       (recursion)
//      L21 = A21 L11^-T                     (GEMM against the explicit inverse of L11)
//      A22 -= L21 L21'                      (GEMM, lower tiles only)
//      A22 = L22 L22'                       (recursion)
//      Linv21 = -Linv22 (L21 Linv11)        (two GEMMs, only when the inverse factor is wanted)
// so that all O(n^3) work runs in the DMMA GEMM of gemm.cu and only CHOL_NB x CHOL_NB diagonal blocks are factorised
// (and inverted) inside one CTA in shared memory.  The interior-point iteration needs L and L^-1 of S (for S^-1 and the
// dual step length) and of X (primal step length), and L of the Schur complement M.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace sdpk {

long long* g_diag_dbg = nullptr;

namespace {

constexpr int NB = CHOL_NB;

// ---- leaf kernel, DMMA version -----------------------------------------------------------------------------------------------
// One CTA factorises (mode 0) and/or inverts (mode 1: A already holds L) a diagonal block of order nb <= NBL, NBL = 64 or 128.
// NBL/8 warps; warp w keeps the 8-row strip w of the block as DMMA accumulator fragments (lower tiles j <= w) in registers.
// Factorisation = NBL/4 macro steps over block columns of width 4 (fully unrolled, two barriers each):
//   (1) the owners of block column t put it into shared memory;
//   (2) one thread per row: 4 x 4 Cholesky of the diagonal block (every row thread redundantly - no extra barrier), forward
//       substitution of its own row -> panel row P_r = L[r, 4t..4t+3]; stored to global memory and to the row-major copy of L;
//   (3) rank-4 update C -= P P' of the not yet finished tiles: ONE DMMA.8x8x4 per tile, both operands read from the contiguous
//       64 x 4 panel (conflict-free fragment loads).
// Inverse W = L^-1 by recursive doubling: 8 x 8 diagonal inverses (one thread per column), then for s = 8, 16, .., NBL/2
//   W21 = -W22 (L21 W11) for all adjacent pairs of s-blocks at once, the products again on DMMA fragments.
// Shared memory: G[(NBL+1) x (NBL+4)] holds L row-major (element (r,c) at row r+1) and W transposed (element (r,c) at row c)
// in the two triangles of one array.
// reciprocal square root: MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, about 22 bits over the whole double range, no
// float conversions) + two Newton steps
__device__ __forceinline__ double fast_rsqrt64(double x)
{
   double y;
   asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
   const double hx = 0.5 * x;
   y = y * (1.5 - hx * y * y);
   y = y * (1.5 - hx * y * y);
   return y;
}

// reciprocal: MUFU.RCP64H seed + two Newton steps
__device__ __forceinline__ double fast_rcp64(double x)
{
   double y;
   asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
   y = fma(y, fma(-x, y, 1.0), y);
   y = fma(y, fma(-x, y, 1.0), y);
   return y;
}

// one product tile of the inverse recursion: D = sum_k A[.][k] B[k][.] over k in [klo, khi) (multiples of 8), operands via
// the two loader functors; two interleaved accumulators shorten the DMMA dependency chain
template <class FA, class FB>
__device__ __forceinline__ void tile_product(int klo, int khi, FA fa, FB fb, double& d0, double& d1)
{
   double e0 = 0.0, e1 = 0.0;
   d0 = 0.0; d1 = 0.0;
   for( int k0 = klo; k0 < khi; k0 += 8 )
   {
      const double a0 = fa(k0), b0 = fb(k0), a1 = fa(k0 + 4), b1 = fb(k0 + 4);
      dmma884(d0, d1, a0, b0);
      dmma884(e0, e1, a1, b1);
   }
   d0 += e0; d1 += e1;
}

__device__ long long* g_leaf_stamps = nullptr;       // debug: clock64 stamps of one macro step (set by the kind-9 probe)

// shared-memory carve-up and thread coordinates of the diagonal-block code (kernel below and the tile kernel potrf_dag_kernel)
template <int NBL>
struct LeafCtx
{
   static constexpr int NTILE = NBL / 8, LD = NBL + 4, NSTEP = NBL / 4, NTHREADS = NBL * 4, NWARP = NBL / 8;
   double *G, *Tr, *Pcol, *P, *rdg;
   int tid, lane, w, fr, fc;
   __device__ __forceinline__ LeafCtx(double* lsm)
   {
      G = lsm;                              // (NBL + 1) x LD
      Tr = G + (NBL + 1) * LD;              // (NBL / 2) x LD
      Pcol = Tr + (NBL / 2) * LD;           // 2 x (NBL x 4)  block column of the current / the next macro step
      P = Pcol + 2 * NBL * 4;               // 2 x (NBL x 4)  panel of the current / the next macro step
      rdg = P + 2 * NBL * 4;                // NBL      reciprocals of the diagonal of L
      tid = threadIdx.x; lane = tid & 31; w = tid >> 5; fr = lane >> 2; fc = lane & 3;
   }
};

// factorisation of the block held as accumulator fragments c (warp w = 8-row strip w, lower tiles j <= w): L goes to global memory
// (lower triangle of A) and, row-major, to G; rdg receives the reciprocals of the diagonal.  sbad: shared int of the caller.
// pf0/pf1 (optional): two flags in global memory that thread 0 reads (relaxed, nothing waits for the loads) four macro steps before
// the end; s_flag <- both were set (with acquire semantics for the whole CTA after the function's last barrier)
template <int NBL>
__device__ __forceinline__ void leaf_factor(const LeafCtx<NBL>& X, double (&c)[NBL / 8][2], int nb, double* __restrict__ A, int lda,
   int* __restrict__ info, int pivot_offset, int& sbad, const int* pf0 = nullptr, const int* pf1 = nullptr, int* s_flag = nullptr)
{
   int polled0 = 0, polled1 = 0;
   constexpr int NTILE = LeafCtx<NBL>::NTILE, LD = LeafCtx<NBL>::LD, NSTEP = LeafCtx<NBL>::NSTEP, NTHREADS = LeafCtx<NBL>::NTHREADS;
   double* const G = X.G; double* const Pcol = X.Pcol; double* const P = X.P; double* const rdg = X.rdg;
   const int tid = X.tid, w = X.w, fr = X.fr, fc = X.fc;
   __shared__ long long lstamp[8];
   P[tid] = 0.0; P[NTHREADS + tid] = 0.0;             // NTHREADS = 4 NBL; both panel buffers
   if( tid == 0 ) sbad = 0x7fffffff;                  // smallest index of a non-positive pivot
   // block column 0 -> shared memory
   if( (fc >> 1) == 0 )
      *reinterpret_cast<double2*>(Pcol + (8 * w + fr) * 4 + 2 * (fc & 1)) = make_double2(c[0][0], c[0][1]);
   // a partial block (order 66 in a 128-leaf: TT-500) stops after its last column group; the identity padding behind it is written here
   const int nbr = (nb + 3) & ~3;
   for( int e = tid; e < (NBL - nbr) * NBL; e += NTHREADS )
   {
      const int r = nbr + e / NBL, cc = e % NBL;
      G[(r + 1) * LD + cc] = (cc == r) ? 1.0 : 0.0;
      if( cc == r ) rdg[r] = 1.0;
   }
   __syncthreads();

   // Macro step t (4 columns), two barriers, ordered so that only what the NEXT step needs sits between them:
   //   (a) row threads: 4 x 4 Cholesky of the diagonal block + their panel row            (the dependent rsqrt chain)
   //   (b) every strip updates the ONE tile that holds block column t+1 and hands that column over
   //   (c) the rank-4 update of all other tiles runs after the second barrier, i.e. beside (a) of step t+1
#pragma unroll
   for( int t = 0; t < NSTEP; ++t )
   {
      if( 4 * t >= nb ) break;
      double* const Pc = Pcol + (t & 1) * NBL * 4;
      double* const Pp = P + (t & 1) * NBL * 4;
      if( t == NSTEP - 4 && pf0 != nullptr && tid == 0 )
      {
         asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(polled0) : "l"(pf0) : "memory");
         asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(polled1) : "l"(pf1) : "memory");
      }
      if( t == 8 && tid == 63 ) lstamp[0] = clock64();
      // (2) panel row of every row r >= 4t
      if( tid < NBL && tid >= 4 * t )
      {
         const int r = tid;
         const double4* Dg = reinterpret_cast<const double4*>(Pc + 16 * t);
         const double4 g0 = Dg[0], g1 = Dg[1], g2 = Dg[2], g3 = Dg[3], av = *reinterpret_cast<const double4*>(Pc + 4 * r);
         // 4 x 4 Cholesky through the leading principal minors m1..m4 (fraction-free elimination with the exact division of
         // Bareiss): the elimination itself needs only multiplications and two reciprocals that do not depend on each other,
         // and the FOUR reciprocal square roots are independent, instead of the chain pivot -> rsqrt -> next pivot -> rsqrt ...
         //    b_ij = m1 g_ij - g_i0 g_j0,   c_ij = (b_11 b_ij - b_i1 b_j1) / m1,   m4 = (c_22 c_33 - c_32^2) / m2,   m2 = b_11, m3 = c_22
         //    L_kk = m_{k+1} q_k,  1 / L_kk = m_k q_k,  L_i1 = b_i1 q_1,  L_32 = c_32 q_2   with   q_k = rsqrt(m_k m_{k+1}), m_0 = 1
         // (same rounding behaviour as the usual elimination: every Schur-complement entry is formed by one product and one FMA;
         // magnitudes stay below |a|^4).  A non-positive minor = non-positive pivot: replaced by 1 and reported after the loop.
         double m1 = g0.x;
         const bool b0 = !(m1 > 0.0); m1 = b0 ? 1.0 : m1;
         const double i1 = fast_rcp64(m1);
         const double b11 = fma(m1, g1.y, -g1.x * g1.x), b21 = fma(m1, g2.y, -g2.x * g1.x), b31 = fma(m1, g3.y, -g3.x * g1.x);
         const double b22 = fma(m1, g2.z, -g2.x * g2.x), b32 = fma(m1, g3.z, -g3.x * g2.x), b33 = fma(m1, g3.w, -g3.x * g3.x);
         double m2 = b11;
         const bool b1 = !(m2 > 0.0); m2 = b1 ? 1.0 : m2;
         const double i2 = fast_rcp64(m2);
         const double c22 = fma(m2, b22, -b21 * b21) * i1, c32 = fma(m2, b32, -b31 * b21) * i1, c33 = fma(m2, b33, -b31 * b31) * i1;
         double m3 = c22;
         const bool b2 = !(m3 > 0.0); m3 = b2 ? 1.0 : m3;
         double m4 = fma(m3, c33, -c32 * c32) * i2;
         const bool b3 = !(m4 > 0.0); m4 = b3 ? 1.0 : m4;
         const double q0 = fast_rsqrt64(m1), q1 = fast_rsqrt64(m1 * m2), q2 = fast_rsqrt64(m2 * m3), q3 = fast_rsqrt64(m3 * m4);
         const double r0 = q0, r1 = m1 * q1, r2 = m2 * q2, r3 = m3 * q3;                    // reciprocals of the diagonal of L
         const double s0 = m1, s1 = m2 * i1, s2 = m3 * i2;                                  // pivots (only their products with r are used)
         const double l10 = g1.x * q0, l20 = g2.x * q0, l30 = g3.x * q0;
         const double l21 = b21 * q1, l31 = b31 * q1, l32 = c32 * q2;
         (void)s0; (void)s1; (void)s2;
         if( r == 4 * t && (b0 || b1 || b2 || b3) )
         {
            const int q = b0 ? 0 : (b1 ? 1 : (b2 ? 2 : 3));
            if( 4 * t + q < nb ) atomicMin(&sbad, 4 * t + q);
         }
         // forward substitution of the row; on the four rows of the diagonal block itself the same formulas give the block's
         // own factor (entry q of row 4t+q is m_{q+1} q_q = L_qq), only the entries above the diagonal are cleared: no branch
         const int q = r - 4 * t;                         // >= 4 for the rows below the block
         double4 pr;
         pr.x = av.x * r0;
         pr.y = (av.y - pr.x * l10) * r1;
         pr.z = (av.z - pr.x * l20 - pr.y * l21) * r2;
         pr.w = (av.w - pr.x * l30 - pr.y * l31 - pr.z * l32) * r3;
         pr.y = (q < 1) ? 0.0 : pr.y;
         pr.z = (q < 2) ? 0.0 : pr.z;
         pr.w = (q < 3) ? 0.0 : pr.w;
         const bool below = (q >= 4);
         *reinterpret_cast<double4*>(Pp + 4 * r) = below ? pr : make_double4(0.0, 0.0, 0.0, 0.0);
         if( !below ) rdg[r] = (q == 0) ? r0 : (q == 1 ? r1 : (q == 2 ? r2 : r3));
         // row-major copy of L (zeros above the diagonal land in the triangle that the inverse fills later)
         *reinterpret_cast<double4*>(G + (r + 1) * LD + 4 * t) = pr;
      }
      if( t == 8 && tid == 63 ) lstamp[1] = clock64();
      __syncthreads();
      if( t == 8 && tid == 63 ) lstamp[2] = clock64();
      const int jlo = (t + 1) >> 1;                      // first tile column that still changes = the one holding block column t+1
      double a = 0.0;
      if( w >= jlo )
      {
         a = -Pp[(8 * w + fr) * 4 + fc];
         if( jlo < NTILE )
         {
            const double b = Pp[(8 * jlo + fr) * 4 + fc];
            dmma884(c[jlo][0], c[jlo][1], a, b);
            if( t + 1 < NSTEP && (fc >> 1) == ((t + 1) & 1) )
               *reinterpret_cast<double2*>(Pcol + ((t + 1) & 1) * NBL * 4 + (8 * w + fr) * 4 + 2 * (fc & 1)) = make_double2(c[jlo][0], c[jlo][1]);
         }
      }
      if( t == 8 && tid == 63 ) lstamp[3] = clock64();
      __syncthreads();
      if( t == 8 && tid == 63 ) lstamp[4] = clock64();
      if( t == 9 && tid == 63 ) lstamp[5] = clock64();
      if( w > jlo )
      {
#pragma unroll
         for( int j = 0; j < NTILE; ++j )
         {
            if( j > jlo && j <= w )
            {
               const double b = Pp[(8 * j + fr) * 4 + fc];
               dmma884(c[j][0], c[j][1], a, b);
            }
         }
      }
   }
   if( pf0 != nullptr && tid == 0 )
   {
      const int both = (polled0 != 0 && polled1 != 0) ? 1 : 0;
      if( both ) asm volatile("fence.acq_rel.gpu;" ::: "memory");
      *s_flag = both;
   }
   __syncthreads();
   if( tid == 0 && sbad != 0x7fffffff ) atomicCAS(info, 0, pivot_offset + sbad + 1);
   if( tid == 0 && g_leaf_stamps != nullptr ) for( int q = 0; q < 6; ++q ) g_leaf_stamps[q] = lstamp[q];
   // L -> global memory, lower triangle, column by column (coalesced)
   for( int e = tid; e < nb * nb; e += NTHREADS )
   {
      const int i = e % nb, j = e / nb;
      if( i >= j ) A[(size_t)j * lda + i] = G[(i + 1) * LD + j];
   }
}

// reciprocal: MUFU.RCP64H seed (about 22 bits) + ONE cubically convergent step  y (1 + e + e^2), e = 1 - x y  (three dependent FMAs)
__device__ __forceinline__ double fast_rcp64_cubic(double x)
{
   double y;
   asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
   const double e = fma(-x, y, 1.0);
   const double q = fma(e, e, e);
   return fma(y, q, y);
}

__device__ __forceinline__ int ld_acquire(const int* p)
{
   int v;
   asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
   return v;
}
__device__ __forceinline__ void st_release(int* p, int v)
{
   asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(count) : "memory"); }

// ---- diagonal block of order 64, second formulation: rows in lanes, columns by warp shuffles ------------------------------------
// The macro steps of leaf_factor cost two CTA barriers and about 240 cycles per column.  Here a block column of width 16 is
// factorised by warps that hold ONE ROW PER LANE in registers (16 entries): lanes 0-15 the rows of the 16 x 16 diagonal block
// (every duty warp redundantly), lanes 16-31 sixteen rows of the panel below it.  Column k inside the warp, no barrier, no shared memory:
//      d = shfl(a[k], k)   inv = 1/d   t = a[k] inv      a[j] -= t * shfl(a[k], j)   (j > k; the shuffles do not wait for inv)
// i.e. the elimination runs on UNSCALED columns (the dependent chain per column is shuffle -> reciprocal -> two FMAs, about 75
// cycles); L_ik = a_ik rsqrt(d_k) is formed beside the chain.  The panel rows finish together with the diagonal block: no triangular
// solve, no inverse of the diagonal block.  Between block columns: the 64 x 16 panel goes to shared memory, every warp applies it to
// its strip of accumulator fragments (rank-16 update, DMMA; the two tile columns of the next block column first, handed to the
// duty warps through shared memory and a producer/consumer barrier, the other tiles behind it), two CTA-wide barriers per 16 columns.
// Results as leaf_factor: L to global memory and row-major to G, reciprocals of the diagonal to rdg.
__device__ __forceinline__ void leaf_factor64(const LeafCtx<64>& X, double (&c)[8][2], int nb, double* __restrict__ A, int lda,
   int* __restrict__ info, int pivot_offset, int& sbad, const int* pf0 = nullptr, const int* pf1 = nullptr, int* s_flag = nullptr)
{
   constexpr int NBL = 64, LD = LeafCtx<64>::LD, NTHREADS = LeafCtx<64>::NTHREADS, LS = 18, LP = 20;
   constexpr unsigned FULL = 0xffffffffu;
   double* const G = X.G; double* const rdg = X.rdg;
   double* const S = X.Tr;                    // 64 x LS: the current block column, one row per lane
   double* const Pp = X.Tr + NBL * LS;        // 64 x LP: the finished panel (K-contiguous DMMA operand); Tr .. P are contiguous
   double* const Bc = Pp + NBL * LP;          // 3 duty warps x 3 x (16 x 4): broadcast buffers of the column code
   static_assert(NBL * LS + NBL * LP + 9 * 64 <= (NBL / 2) * LD + 16 * NBL, "scratch of the shuffle formulation fits Tr + Pcol + P");
   const int tid = X.tid, lane = X.lane, fr = X.fr, fc = X.fc;
   const int w = __shfl_sync(FULL, X.w, 0);   // warp-uniform for the compiler: the warp-level code below runs without divergence checks
   int polled0 = 0, polled1 = 0;
   __shared__ long long lstamp[8];
   if( tid == 0 ) sbad = 0x7fffffff;
   // (the loop over the block columns is NOT unrolled: 20 KB of straight-line code per block column, executed by one to three warps;
   // unrolled fourfold the kernel streamed its instructions from L2 and the factorisation took four times as long)
#pragma unroll 1
   for( int b = 0; b < 4; ++b )
   {
      const int jt0 = 2 * b;
      const int ngroups = 3 - b;                         // groups of 16 panel rows below the diagonal block
      const int nduty = (ngroups > 0) ? ngroups : 1;
      const int wd = w - jt0;
      const bool duty = (wd >= 0 && wd < nduty);
      if( b == 3 && pf0 != nullptr && tid == 0 )
      {
         polled0 = ld_acquire(pf0);           // thread 0 has nothing else to do in this block column: nobody waits for the loads
         polled1 = ld_acquire(pf1);
      }
      if( b == 1 && tid == 64 ) lstamp[0] = clock64();
      // (1) the two tile columns of block column b: rank-16 update with the previous panel, then to shared memory
      double av[4] = {0.0, 0.0, 0.0, 0.0};
      if( b > 0 && w >= jt0 )
      {
#pragma unroll
         for( int kk = 0; kk < 4; ++kk ) av[kk] = -Pp[(8 * w + fr) * LP + 4 * kk + fc];
      }
#pragma unroll
      for( int jt = 0; jt < 8; ++jt )
      {
         if( (jt == jt0 || jt == jt0 + 1) && jt <= w )
         {
            if( b > 0 )
            {
#pragma unroll
               for( int kk = 0; kk < 4; ++kk ) dmma884(c[jt][0], c[jt][1], av[kk], Pp[(8 * jt + fr) * LP + 4 * kk + fc]);
            }
            *reinterpret_cast<double2*>(S + (8 * w + fr) * LS + 8 * (jt - jt0) + 2 * fc) = make_double2(c[jt][0], c[jt][1]);
         }
      }
      if( duty ) named_bar_sync(1, NTHREADS); else named_bar_arrive(1, NTHREADS);
      if( b == 1 && tid == 64 ) lstamp[1] = clock64();
      if( !duty )
      {
         if( b > 0 )
         {
            // (2) the other tiles of the strip (needed two block columns later at the earliest), and the previous panel -> G
#pragma unroll
            for( int jt = 2; jt < 8; ++jt )
            {
               if( jt >= jt0 + 2 && jt <= w )
               {
#pragma unroll
                  for( int kk = 0; kk < 4; ++kk ) dmma884(c[jt][0], c[jt][1], av[kk], Pp[(8 * jt + fr) * LP + 4 * kk + fc]);
               }
            }
            const int rank = (w < jt0) ? w : w - nduty, nnd = 8 - nduty;
            for( int r = 16 * (b - 1) + 4 * rank + (lane >> 3); r < NBL; r += 4 * nnd )
               *reinterpret_cast<double2*>(G + (r + 1) * LD + 16 * (b - 1) + 2 * (lane & 7)) = *reinterpret_cast<const double2*>(Pp + r * LP + 2 * (lane & 7));
         }
         named_bar_arrive(2, NTHREADS);                  // done reading the previous panel
      }
      else
      {
         // (3) block column b: one row per lane
         const bool lower = (lane < 16);
         const bool valid = lower || (ngroups > 0);
         const int grow = lower ? 16 * b + lane : 16 * b + 16 + 16 * wd + (lane - 16);
         double a[16];
#pragma unroll
         for( int q = 0; q < 8; ++q )
         {
            double2 v = make_double2(0.0, 0.0);
            if( valid ) v = *reinterpret_cast<const double2*>(S + grow * LS + 2 * q);
            a[2 * q] = v.x; a[2 * q + 1] = v.y;
         }
#pragma unroll
         for( int j = 1; j < 16; ++j ) a[j] = (lower && j > lane) ? 0.0 : a[j];       // above the diagonal: never used, kept finite
         if( b == 1 && tid == 64 ) lstamp[2] = clock64();
         // 16 columns as four mini blocks of 4: inside a mini block the columns go through warp shuffles (pivot, then the
         // entries of the 4 x 4 pivot block); the finished four columns of the diagonal rows are then broadcast ONCE through shared
         // memory (one 32-byte store per lane, two 16-byte broadcast loads per row) for the rank-4 update of the columns to the right
         int badk = 0x7fffffff;
         double* const Bw = Bc + wd * 3 * 64;            // broadcast buffers of this warp: 3 mini blocks x 16 rows x 4
         double d = __shfl_sync(FULL, a[0], 0);
         double dmine = a[0];                            // lane l < 16: its own pivot d_l (unscaled diagonal entry at elimination time)
#pragma unroll
         for( int mb = 0; mb < 4; ++mb )
         {
            const int k0 = 4 * mb;
            double t[4];
            double2 f01 = make_double2(0.0, 0.0), f23 = make_double2(0.0, 0.0);      // row k0 + 4 of the broadcast, fetched early
#pragma unroll
            for( int q = 0; q < 4; ++q )
            {
               const int k = k0 + q;
               if( q == 3 && mb < 3 )
               {
                  // the four columns of this lane are final (the last one needs no update from column k itself): publish them
                  if( lower )
                  {
                     *reinterpret_cast<double2*>(Bw + mb * 64 + lane * 4) = make_double2(a[k0], a[k0 + 1]);
                     *reinterpret_cast<double2*>(Bw + mb * 64 + lane * 4 + 2) = make_double2(a[k0 + 2], a[k0 + 3]);
                  }
                  __syncwarp();
                  f01 = *reinterpret_cast<const double2*>(Bw + mb * 64 + (k0 + 4) * 4);
                  f23 = *reinterpret_cast<const double2*>(Bw + mb * 64 + (k0 + 4) * 4 + 2);
               }
               // a non-positive pivot is reported (and the results are void); the chain itself does not wait for the test
               badk = (!(d > 0.0) && badk == 0x7fffffff) ? 16 * b + k : badk;
               dmine = (lane == k) ? d : dmine;
               const double inv = fast_rcp64_cubic(d);
               double dn = 0.0;
               if( q < 3 )
               {
                  const double u1 = __shfl_sync(FULL, a[k], k + 1);
                  const double p1 = a[k] * u1;
                  a[k + 1] = fma(-p1, inv, a[k + 1]);
                  dn = __shfl_sync(FULL, a[k + 1], k + 1);
               }
               t[q] = a[k] * inv;
#pragma unroll
               for( int j = k + 2; j < k0 + 4; ++j )
               {
                  const double u = __shfl_sync(FULL, a[k], j);
                  a[j] = fma(-t[q], u, a[j]);
               }
               if( q < 3 ) d = dn;
            }
            if( mb < 3 )
            {
               a[k0 + 4] = fma(-t[3], f23.y, fma(-t[2], f23.x, fma(-t[1], f01.y, fma(-t[0], f01.x, a[k0 + 4]))));
               d = __shfl_sync(FULL, a[k0 + 4], k0 + 4);
#pragma unroll
               for( int j = k0 + 5; j < 16; ++j )
               {
                  const double2 u01 = *reinterpret_cast<const double2*>(Bw + mb * 64 + j * 4);
                  const double2 u23 = *reinterpret_cast<const double2*>(Bw + mb * 64 + j * 4 + 2);
                  a[j] = fma(-t[3], u23.y, fma(-t[2], u23.x, fma(-t[1], u01.y, fma(-t[0], u01.x, a[j]))));
               }
            }
         }
         // L_ik = a_ik / sqrt(d_k): ONE reciprocal square root per lane (lane l: column l), handed round through shared memory
         {
            const bool badp = !(dmine > 0.0);
            const double rsl = fast_rsqrt64(badp ? 1.0 : dmine);
            __syncwarp();
            if( lower ) Bw[lane] = rsl;
            if( lower && wd == 0 ) rdg[16 * b + lane] = rsl;
            __syncwarp();
#pragma unroll
            for( int q = 0; q < 8; ++q )
            {
               const double2 r2 = *reinterpret_cast<const double2*>(Bw + 2 * q);
               a[2 * q] *= r2.x; a[2 * q + 1] *= r2.y;
            }
         }
         if( lane == 0 && wd == 0 && badk < nb ) atomicMin(&sbad, badk);
#pragma unroll
         for( int j = 1; j < 16; ++j ) a[j] = (lower && j > lane) ? 0.0 : a[j];
         // (a duty warp never has tiles right of the block column: with b > 0 at most two warps are on duty, strips 2b and 2b + 1)
         if( b == 1 && tid == 64 ) lstamp[3] = clock64();
         named_bar_sync(2, NTHREADS);                    // everybody is done reading the previous panel
         if( b == 1 && tid == 64 ) lstamp[4] = clock64();
         if( valid && (!lower || wd == 0) )
         {
#pragma unroll
            for( int q = 0; q < 8; ++q )
               *reinterpret_cast<double2*>(Pp + grow * LP + 2 * q) = make_double2(a[2 * q], a[2 * q + 1]);
         }
      }
      if( b == 3 && pf0 != nullptr && tid == 0 )
      {
         *s_flag = (polled0 != 0 && polled1 != 0) ? 1 : 0;
      }
      __syncthreads();
      if( b == 1 && tid == 64 ) lstamp[5] = clock64();
   }
   // the last panel -> G
   if( tid < 128 )
      *reinterpret_cast<double2*>(G + (48 + (tid >> 3) + 1) * LD + 48 + 2 * (tid & 7)) = *reinterpret_cast<const double2*>(Pp + (48 + (tid >> 3)) * LP + 2 * (tid & 7));
   __syncthreads();
   if( tid == 0 && g_leaf_stamps != nullptr ) for( int q = 0; q < 6; ++q ) g_leaf_stamps[q] = lstamp[q];
   if( tid == 0 && sbad != 0x7fffffff ) atomicCAS(info, 0, pivot_offset + sbad + 1);
   // L -> global memory, lower triangle, column by column (coalesced)
   for( int e = tid; e < nb * nb; e += NTHREADS )
   {
      const int i = e % nb, j = e / nb;
      if( i >= j ) A[(size_t)j * lda + i] = G[(i + 1) * LD + j];
   }
}

// inverse W = L^-1 of the block whose factor sits row-major in G (with rdg): 8 x 8 diagonal inverses, then recursive doubling on
// DMMA fragments; W ends up transposed in the upper triangle of G (element (r,c) at row c)
template <int NBL>
__device__ __forceinline__ void leaf_invert(const LeafCtx<NBL>& X)
{
   constexpr int LD = LeafCtx<NBL>::LD, NWARP = LeafCtx<NBL>::NWARP;
   double* const G = X.G; double* const Tr = X.Tr; double* const rdg = X.rdg;
   const int tid = X.tid, w = X.w, fr = X.fr, fc = X.fc;
   // ---- inverse: 8 x 8 diagonal blocks, one thread per column ----
   if( tid < NBL )
   {
      const int b8 = tid >> 3, cq = tid & 7, o = 8 * b8;
      double x[8];
#pragma unroll
      for( int i = 0; i < 8; ++i )
      {
         double sacc = 0.0;
#pragma unroll
         for( int p2 = 0; p2 < 8; ++p2 ) if( p2 < i ) sacc += G[(o + i + 1) * LD + o + p2] * x[p2];
         x[i] = (i == cq) ? rdg[o + i] : ((i > cq) ? -rdg[o + i] * sacc : 0.0);
      }
#pragma unroll
      for( int i = 0; i < 8; ++i ) if( i >= cq ) G[(o + cq) * LD + o + i] = x[i];
   }
   __syncthreads();
#pragma unroll
   for( int ls = 3; (1 << ls) < NBL; ++ls )
   {
      const int s = 1 << ls, lt = ls - 3;                        // s-blocks, (s/8)^2 = 4^lt tiles per pair
      const int total = (NBL / (2 * s)) << (2 * lt);
      // T = L21 W11   (W11 lower triangular: k >= column tile)
      for( int idx = w; idx < total; idx += NWARP )
      {
         const int pr = idx >> (2 * lt), rem = idx & ((1 << (2 * lt)) - 1), ti = rem >> lt, tj = rem & ((1 << lt) - 1);
         const int o = 2 * s * pr;
         const double* Lrow = G + (size_t)(o + s + 8 * ti + fr + 1) * LD + o + fc;
         const double* Wcol = G + (size_t)(o + 8 * tj + fr) * LD + o + fc;
         const int cdiag = 8 * tj + fr - fc;
         double d0, d1;
         tile_product(8 * tj, s, [&](int k0) { return Lrow[k0]; }, [&](int k0) { return (k0 >= cdiag) ? Wcol[k0] : 0.0; }, d0, d1);
         *reinterpret_cast<double2*>(Tr + (size_t)(pr * s + 8 * ti + fr) * LD + 8 * tj + 2 * fc) = make_double2(d0, d1);
      }
      __syncthreads();
      // W21 = -W22 T  (W22 lower triangular: k <= row tile)
      for( int idx = w; idx < total; idx += NWARP )
      {
         const int pr = idx >> (2 * lt), rem = idx & ((1 << (2 * lt)) - 1), ti = rem >> lt, tj = rem & ((1 << lt) - 1);
         const int o = 2 * s * pr;
         const int rrow = o + s + 8 * ti + fr;
         const double* W2 = G + (size_t)(o + s + fc) * LD + rrow;
         const double* Tc = Tr + (size_t)(pr * s + fc) * LD + 8 * tj + fr;
         const int rdiag = 8 * ti + fr - fc;
         double d0, d1;
         tile_product(0, 8 * ti + 8, [&](int k0) { return (k0 <= rdiag) ? -W2[(size_t)k0 * LD] : 0.0; }, [&](int k0) { return Tc[(size_t)k0 * LD]; }, d0, d1);
         G[(size_t)(o + 8 * tj + 2 * fc) * LD + rrow] = d0;
         G[(size_t)(o + 8 * tj + 2 * fc + 1) * LD + rrow] = d1;
      }
      __syncthreads();
   }
}

template <int NBL>
__global__ void __launch_bounds__(NBL * 4)
leaf_kernel(int mode, int nb, double* __restrict__ A, int lda, double* __restrict__ Linv, int ldi,
   double* __restrict__ diaginv, int* __restrict__ info, int pivot_offset, long long* __restrict__ dbg)
{
   long long tc0 = clock64(), tc1 = 0, tc2 = 0, tc3 = 0;
   constexpr int NTILE = NBL / 8, LD = NBL + 4, NSTEP = NBL / 4, NTHREADS = NBL * 4, NWARP = NBL / 8;
   extern __shared__ __align__(16) double lsm[];
   __shared__ int sbad;
   const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
   const int fr = lane >> 2, fc = lane & 3;

   const LeafCtx<NBL> X(lsm);
   double* const G = X.G;
   double* const rdg = X.rdg;
   if( mode == 0 )
   {
      double c[NTILE][2];
#pragma unroll
      for( int j = 0; j < NTILE; ++j )
      {
         c[j][0] = 0.0; c[j][1] = 0.0;
         if( j <= w )
         {
            const int row = 8 * w + fr;
#pragma unroll
            for( int e = 0; e < 2; ++e )
            {
               const int col = 8 * j + 2 * fc + e;
               double v = (row == col) ? 1.0 : 0.0;      // identity padding keeps a partial block positive definite
               if( row < nb && col < nb && row >= col ) v = A[(size_t)col * lda + row];
               c[j][e] = v;
            }
         }
      }
      tc1 = clock64();
      if( tid == 0 ) g_leaf_stamps = (dbg != nullptr) ? dbg + 4 : nullptr;
      __syncthreads();
      if constexpr( NBL == 64 ) leaf_factor64(X, c, nb, A, lda, info, pivot_offset, sbad);
      else leaf_factor<NBL>(X, c, nb, A, lda, info, pivot_offset, sbad);
      if( tid == 0 ) g_leaf_stamps = nullptr;
      tc2 = clock64();
   }
   else
   {
      for( int e = tid; e < NBL * NBL; e += NTHREADS )
      {
         const int i = e % NBL, j = e / NBL;
         if( i >= j )
         {
            double v = (i == j) ? 1.0 : 0.0;
            if( i < nb && j < nb ) v = A[(size_t)j * lda + i];
            G[(i + 1) * LD + j] = v;
            if( i == j ) rdg[i] = 1.0 / v;
         }
      }
      __syncthreads();
      tc2 = clock64();
   }
   if( Linv == nullptr && diaginv == nullptr ) return;

   leaf_invert<NBL>(X);
   tc3 = clock64();
   if( Linv != nullptr )
      for( int e = tid; e < nb * nb; e += NTHREADS )
      {
         const int i = e % nb, j = e / nb;
         Linv[(size_t)j * ldi + i] = (i >= j) ? G[j * LD + i] : 0.0;
      }
   if( diaginv != nullptr )
      for( int e = tid; e < NBL * NBL; e += NTHREADS )
      {
         const int i = e % NBL, j = e / NBL;
         diaginv[(size_t)j * NBL + i] = (i < nb && j < nb && i >= j) ? G[j * LD + i] : 0.0;
      }
   if( dbg != nullptr && tid == 0 )
   {
      dbg[0] = tc1 - tc0; dbg[1] = tc2 - tc1; dbg[2] = tc3 - tc2; dbg[3] = clock64() - tc3;
   }
}

template <int NBL> constexpr size_t leaf_smem() { return sizeof(double) * ((size_t)(NBL + 1) * (NBL + 4) + (size_t)(NBL / 2) * (NBL + 4) + 17 * (size_t)NBL); }

// ---- tile-DAG Cholesky: ONE persistent kernel for the whole factorisation --------------------------------------------------------
// The lower triangle is cut into 64 x 64 tiles, ordered column by column (diagonal tile first).  CTAs claim tiles in that order from
// an atomic counter and compute a tile completely ("left-looking per tile"):
//     C = A_ij - sum_{k<j} L_ik L_jk'        one DMMA accumulation over K = 64 j, operands streamed through a cp.async ring
//     i == j:  L_jj = chol(C), W_jj = L_jj^-1  (the diagonal-block code above, on the accumulator fragments)
//     i >  j:  L_ij = C W_jj'                 (one more 64^3 product out of shared memory)
// and publish it through a ready flag (release/acquire at GPU scope).  A tile only depends on tiles that precede it in the claim
// order, and a claimed tile is owned by a running CTA, so the smallest unfinished tile can always proceed: no deadlock whatever
// part of the grid is resident (other streams may hold SMs).  The chain POTRF(j) -> TRSM(j+1,j) -> last update of (j+1,j+1) is the
// critical path; all other tiles of later columns are claimed early and hide behind it (look-ahead without a schedule).
constexpr int DAG_T = 64, DAG_THREADS = 256, DAG_BK = 16, DAG_STAGES = 4, DAG_LDS = DAG_T + 4, DAG_LDK = DAG_BK + 4;
// one ring slot: an MN-contiguous chunk [k][r] (16 x 68) plus either a second one (factor tiles) or a K-contiguous chunk [c][k] (64 x 20,
// tiles of the inverse)
constexpr int DAG_SLOT = DAG_BK * DAG_LDS + DAG_T * DAG_LDK;
constexpr size_t DAG_SMEM = sizeof(double) * (size_t)DAG_STAGES * DAG_SLOT;      // 75776 B >= leaf_smem<64>() and two 64 x 68 operand tiles
static_assert(DAG_STAGES * DAG_SLOT >= 2 * DAG_T * DAG_LDS && DAG_T * DAG_LDK >= DAG_BK * DAG_LDS, "operand tiles fit the ring");
// chain variant: the diagonal-block code keeps its shared memory for the whole kernel, one 64 x 68 operand tile behind it
constexpr size_t DAG_LEAF_DOUBLES = leaf_smem<DAG_T>() / sizeof(double);
constexpr size_t DAG_SMEM_CHAIN = leaf_smem<DAG_T>() + sizeof(double) * (size_t)DAG_T * DAG_LDS;
static_assert(DAG_SMEM_CHAIN >= DAG_SMEM && leaf_smem<DAG_T>() % 16 == 0, "chain layout");

// Row sums of W in pieces: a tile (i, j) of W with more than DAG_WSEG k-tiles hands the first pieces of its sum to HELPER tasks (up to
// three, DAG_WSEG k-tiles each), which are claimed before the tiles of the row, add up their piece as soon as its operands exist
// (columns far left of the chain: long before) and leave it in a scratch tile; the tile's own task keeps the last piece - the one
// that waits for the chain - and adds the helpers' pieces in a fixed order.  Without them the last rows of W (31 k-tiles = 65 us of
// DMMA time on one SM per tile) were still adding when the chain had ended: 0.12 ms of tail at n = 2000.
constexpr int DAG_WSEG = 8;
__host__ __device__ inline int dag_nhelp(int len) { const int s = (len + DAG_WSEG - 1) / DAG_WSEG - 1; return s < 0 ? 0 : (s > 3 ? 3 : s); }
__host__ __device__ inline int dag_rowhelp(int i) { int h = 0; for( int len = 1; len <= i; ++len ) h += dag_nhelp(len); return h; }     // helpers of row i

struct DagArgs
{
   int n, T;
   double* A; int lda;
   double* Linv; int ldi;        // optional: the inverse factor W = L^-1 (zeroed before the launch)
   int winv;                     // 1: the off-diagonal tiles of W are tasks of this kernel as well (ready flag of W_ij: ready[j * T + i])
   double* Wd;                   // T packed 64 x 64 inverses of the diagonal blocks of L
   int* sync;                    // [0] tile counter, [1] abort flag, [2 ..] T*T ready flags, then 2T flags of the pre-updated tiles (zeroed before the launch)
   int chain;                    // 1: ONE CTA carries the whole critical chain (all diagonal tiles and the tiles (j+1, j)), see dag_chain
   int wshift;                   // the tiles of row r of W are claimed with block column r - wshift (0, 1 or 2)
   int whelp;                    // 1: helper tasks for the long row sums of W (scratch tiles in `wscratch`, flags behind the other flags)
   int wpanel;                   // > 0: only the tiles of W inside diagonal panels of wpanel x wpanel tiles (the inverses of the panels' diagonal blocks)
   double* wscratch;
   long long watchdog;           // cycles a flag wait may last before the kernel aborts (0: no limit; SDPCUDA_DAG_WATCHDOG_S, default 2 s)
   int* info;
   long long* dbg;               // optional: 8 timestamps (ns) per tile of the critical chain (diagonal tiles: slot 2j, tiles (j+1,j): slot 2j+1)
};

__device__ __forceinline__ long long dag_now()
{
   long long t;
   asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
   return t;
}

__device__ __forceinline__ void dag_cp16(void* smem, const void* gmem, int srcbytes)
{
   unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
   asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(sa), "l"(gmem), "r"(srcbytes));
}

// thread 0 waits until the flags f0 and f1 are set (or the kernel is being aborted), then the whole CTA passes
__device__ __forceinline__ bool dag_wait(const int* f0, const int* f1, int* abortflag, int& s_abort, long long watchdog)
{
   if( threadIdx.x == 0 )
   {
      const long long t0 = clock64();
      int bad = 0;
      while( ld_acquire(f0) == 0 || ld_acquire(f1) == 0 )
      {
         if( ld_acquire(abortflag) != 0 ) { bad = 1; break; }
         if( watchdog > 0 && clock64() - t0 > watchdog ) { atomicExch(abortflag, 1); bad = 1; break; }     // a bug, not a wait
      }
      s_abort = bad;
   }
   __syncthreads();
   return s_abort == 0;
}

// Flags of a RANGE of k-tiles at once: warp 0 looks at up to 32 consecutive flag pairs fi[k], fj[k] (k = kt .. kmax - 1) per round trip
// and reports how far the operands are ready (s_upto: all k-tiles below it); returns when k-tile kt itself is there.  The flags only
// ever go from 0 to 1, so what has been seen ready stays ready: a task whose operands were finished long ago (most tiles of a large
// matrix) pays one round trip per 32 k-tiles instead of one per k-tile (two dependent L2 reads = 0.7 us against 1 us of DMMA work).
__device__ __forceinline__ bool dag_wait_upto(const int* fi, const int* fj, int kt, int kmax, int* abortflag, int& s_abort, int& s_upto, long long watchdog)
{
   if( threadIdx.x < 32 )
   {
      const int lane = threadIdx.x;
      const long long t0 = clock64();
      int bad = 0, upto = kt;
      for( ;; )
      {
         const int k = kt + lane;
         int r = 1;
         if( k < kmax ) r = (ld_acquire(fi + k) != 0 && ld_acquire(fj + k) != 0) ? 1 : 0;
         const unsigned m = __ballot_sync(0xffffffffu, r != 0);
         const int lead = (m == 0xffffffffu) ? 32 : (__ffs((int)~m) - 1);
         if( lead > 0 ) { upto = min(kt + lead, kmax); break; }
         if( lane == 0 )
         {
            if( ld_acquire(abortflag) != 0 ) bad = 1;
            else if( watchdog > 0 && clock64() - t0 > watchdog ) { atomicExch(abortflag, 1); bad = 1; }     // a bug, not a wait
         }
         bad = __shfl_sync(0xffffffffu, bad, 0);
         if( bad ) break;
      }
      if( lane == 0 ) { s_abort = bad; s_upto = upto; }
   }
   __syncthreads();
   return s_abort == 0;
}

// 16 columns [k0, k0+16) of the 64-row tile at (row0, .) of a column-major matrix -> smem [k][r], leading dimension DAG_LDS
__device__ __forceinline__ void dag_load_chunk(double* s, const double* __restrict__ g, int ld, int row0, int k0, int nrows, int tid)
{
#pragma unroll
   for( int c = tid; c < DAG_BK * (DAG_T / 2); c += DAG_THREADS )
   {
      const int k = c / (DAG_T / 2), r = (c % (DAG_T / 2)) * 2;
      const int gr = row0 + r;
      int bytes = 0;
      if( gr < nrows ) bytes = (nrows - gr >= 2) ? 16 : 8;
      const double* src = bytes ? (g + (size_t)(k0 + k) * ld + gr) : g;
      dag_cp16(s + k * DAG_LDS + r, src, bytes);
   }
}

// ---- the critical chain in ONE CTA (orders up to 3072, where the factorisation is bound by its dependency chain) ----------------
// The CTA that claims task 0 keeps the diagonal-block code's shared memory for the whole kernel and walks down the diagonal:
//     L_jj = chol(D_j), W_jj = L_jj^-1                      (leaf code, on the accumulator fragments)      -> ready[j][j]
//     L_{j+1,j} = P_{j+1,j} W_jj'                           (W_jj' straight out of the leaf's shared memory) -> ready[j+1][j]
//     D_{j+1} -= L_{j+1,j} L_{j+1,j}'                       (stays in the accumulator fragments: the next leaf starts at once)
// where P_{j+1,j} = A_{j+1,j} - sum_{k<j} L_{j+1,k} L_jk' and D_{j+1} = A_{j+1,j+1} - sum_{k<j} L_{j+1,k} L_{j+1,k}' are PRE-UPDATED
// in place by ordinary tasks of the claim order (slots of the tiles (j,j) and (j+1,j); flags pre[2(j+1)], pre[2(j+1)+1]); they only
// need columns < j, i.e. they are ready while the chain still factorises D_j, and P is prefetched during that leaf.  Off the chain
// compared with one task per tile: two flag hops, the reload of W_jj, the store/reload of L_{j+1,j} and of the diagonal tile.
__device__ __forceinline__ bool dag_chain(const DagArgs& a, double* dsm, int* abortflag, int& s_abort, int& sbad, int& s_pref, long long watchdog)
{
   const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, fr = lane >> 2, fc = lane & 3;
   const int T = a.T, n = a.n;
   int* const ready = a.sync + 2;
   int* const pre = ready + T * T;
   const LeafCtx<DAG_T> X(dsm);
   constexpr int LD = LeafCtx<DAG_T>::LD;
   double* const Pt = dsm + DAG_LEAF_DOUBLES;       // [k][r], leading dimension DAG_LDS
   double c[8][2];
   // (everything below exists ONCE in the kernel and the inner loops are not unrolled further than the tensor pipe needs: what this
   // CTA executes per step has to stay in the instruction cache - its warps often run alone, with nobody to hide a fetch behind)
   int jnext = 0;                                   // the diagonal tile to load into c: 0 at the start, j + 1 inside the loop
   if( tid == 0 ) g_leaf_stamps = (a.dbg != nullptr) ? a.dbg + 16 * T : nullptr;     // the leaf's own stamps (chain profile)
   for( int j = -1; j < T; ++j )
   {
      long long* const dbg = (a.dbg != nullptr && tid == 0 && j >= 0) ? a.dbg + 16 * j : nullptr;
      const bool has_next = (j + 1 < T);
      bool pref = false;
      if( j >= 0 )
      {
         const int col0 = DAG_T * j, nb = min(DAG_T, n - col0);
         if( dbg ) { dbg[0] = dag_now(); dbg[1] = dbg[0]; }
         const bool poll = has_next && j > 0;
         // (inside this kernel the macro-step code is the faster one - 10.5 against 11.8 us per block; alone in leaf_kernel<64> the
         // shuffle formulation wins, 17.9 k against 19.7 k cycles: profiles/r2_leaf_formulations.log)
         leaf_factor<DAG_T>(X, c, nb, a.A + (size_t)col0 * a.lda + col0, a.lda, a.info, col0, sbad,
            poll ? pre + 2 * (j + 1) : nullptr, poll ? pre + 2 * (j + 1) + 1 : nullptr, &s_pref);
         if( dbg ) dbg[2] = dag_now();
         pref = has_next && (j == 0 || s_pref != 0);     // the pre-updated tiles are there (the usual case)
      }
      if( j >= 0 && !has_next ) pref = false;
      // tile (j+1, j) -> Pt, in flight during the inverse
      if( pref )
      {
#pragma unroll 1
         for( int q = 0; q < 4; ++q ) dag_load_chunk(Pt + q * DAG_BK * DAG_LDS, a.A, a.lda, DAG_T * (j + 1), DAG_T * j + q * DAG_BK, n, tid);
         asm volatile("cp.async.commit_group;\n" ::);
      }
      if( j >= 0 )
      {
         const int col0 = DAG_T * j, nb = min(DAG_T, n - col0);
         leaf_invert<DAG_T>(X);
         if( dbg ) dbg[3] = dag_now();
         double* Wj = a.Wd + (size_t)j * DAG_T * DAG_T;
#pragma unroll 2
         for( int e = tid; e < DAG_T * DAG_T; e += DAG_THREADS )
         {
            const int r = e % DAG_T, q = e / DAG_T;
            const double v = (r < nb && q < nb && r >= q) ? X.G[q * LD + r] : 0.0;
            Wj[(size_t)q * DAG_T + r] = v;
            if( a.Linv != nullptr && r < nb && q < nb ) a.Linv[(size_t)(col0 + q) * a.ldi + col0 + r] = v;
         }
         if( dbg ) dbg[4] = dag_now();
         // Publishing without a fence on the chain: the stores are followed by a CTA barrier, and thread 0 releases the flag later
         // (release is cumulative over what the barrier ordered before it), when its own stores have long landed.
         if( !has_next )
         {
            __syncthreads();
            if( tid == 0 ) st_release(ready + j * T + j, 1);
            if( dbg ) dbg[5] = dag_now();
            break;
         }
         // L_jj and W_jj: released at once (thread 0 waits for the tiles of the next step anyway, and their tasks are waiting for W_jj)
         __syncthreads();
         if( tid == 0 ) st_release(ready + j * T + j, 1);
         if( dbg ) { dbg[5] = dag_now(); dbg[8] = dbg[5]; dbg[9] = dbg[5]; }
         if( !pref )
         {
            if( !dag_wait(pre + 2 * (j + 1), pre + 2 * (j + 1) + 1, abortflag, s_abort, watchdog) ) return false;
#pragma unroll 1
            for( int q = 0; q < 4; ++q ) dag_load_chunk(Pt + q * DAG_BK * DAG_LDS, a.A, a.lda, DAG_T * (j + 1), DAG_T * j + q * DAG_BK, n, tid);
            asm volatile("cp.async.commit_group;\n" ::);
         }
      }
      // the next diagonal tile -> c (pre-updated by its task; the raw tile for j + 1 = 0 and 1): the loads fly during the product below
      {
         const int o = DAG_T * jnext, row = o + 8 * w + fr;
#pragma unroll
         for( int jt = 0; jt < 8; ++jt )
#pragma unroll
            for( int e = 0; e < 2; ++e )
            {
               const int col = o + 8 * jt + 2 * fc + e;
               double v = 0.0;
               if( row < n && col < n && row >= col ) v = __ldcg(a.A + (size_t)col * a.lda + row);
               if( row == col && row >= n ) v = 1.0;
               c[jt][e] = v;
            }
         jnext = j + 2;
      }
      if( j < 0 ) continue;
      const int col0 = DAG_T * j;
      asm volatile("cp.async.wait_group 0;\n" ::);
      __syncthreads();
      if( dbg ) dbg[10] = dag_now();
      // L_{j+1,j} = P W_jj':  B operand [k][c] = W_jj[c][k] = G[k * LD + c] for k <= c (below it the array holds L)
      double d[8][2];
#pragma unroll
      for( int jt = 0; jt < 8; ++jt ) { d[jt][0] = 0.0; d[jt][1] = 0.0; }
#pragma unroll
      for( int k0 = 0; k0 < DAG_T; k0 += 4 )
      {
         const double av = Pt[(k0 + fc) * DAG_LDS + 8 * w + fr];
#pragma unroll
         for( int jt = 0; jt < 8; ++jt )
         {
            if( k0 >= 8 * jt + 8 ) continue;
            double bv = X.G[(k0 + fc) * LD + 8 * jt + fr];
            if( k0 + 4 > 8 * jt ) bv = (k0 + fc <= 8 * jt + fr) ? bv : 0.0;
            dmma884(d[jt][0], d[jt][1], av, bv);
         }
      }
      if( dbg ) dbg[11] = dag_now();
      const int row = DAG_T * (j + 1) + 8 * w + fr;
      if( row < n )
      {
#pragma unroll
         for( int jt = 0; jt < 8; ++jt )
         {
            const int col = col0 + 8 * jt + 2 * fc;
            a.A[(size_t)col * a.lda + row] = d[jt][0];
            a.A[(size_t)(col + 1) * a.lda + row] = d[jt][1];
         }
      }
      __syncthreads();                        // everybody is done reading Pt and G
#pragma unroll
      for( int jt = 0; jt < 8; ++jt )
      {
         Pt[(8 * jt + 2 * fc) * DAG_LDS + 8 * w + fr] = d[jt][0];
         Pt[(8 * jt + 2 * fc + 1) * DAG_LDS + 8 * w + fr] = d[jt][1];
      }
      if( dbg ) dbg[12] = dag_now();
      __syncthreads();
      if( tid == 0 ) st_release(ready + (j + 1) * T + j, 1);   // L_{j+1,j}: the pre-updates of the next step wait for it
      if( dbg ) dbg[13] = dag_now();
      // D_{j+1} -= L L'
#pragma unroll
      for( int kk = 0; kk < DAG_T; kk += 4 )
      {
         const double av = -Pt[(kk + fc) * DAG_LDS + 8 * w + fr];
#pragma unroll
         for( int jt = 0; jt < 8; ++jt )
         {
            if( jt > w ) continue;
            const double bv = Pt[(kk + fc) * DAG_LDS + 8 * jt + fr];
            dmma884(c[jt][0], c[jt][1], av, bv);
         }
      }
      // (Pt is written again only after the barriers of the next diagonal block)
   }
   if( tid == 0 ) g_leaf_stamps = nullptr;
   return true;
}

// up to two matrices per launch (S and X of an interior-point iteration): their tiles are claimed alternately from one counter, so the
// two dependency chains advance side by side on different SMs instead of two kernels sharing every SM
struct DagPair { DagArgs p[2]; int count; };

// CHAIN: the variant with the critical chain in one CTA (two instantiations rather than one kernel with both paths: the instruction
// footprint of what the chain CTA executes per step decides how fast a lone warp is fed)
template <bool CHAIN>
__global__ void __launch_bounds__(DAG_THREADS, 2) potrf_dag_kernel(const __grid_constant__ DagPair pair)
{
   extern __shared__ __align__(16) double dsm[];
   __shared__ int s_tile, s_abort, sbad, s_pref, s_upto;
   const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, fr = lane >> 2, fc = lane & 3;
   int* const counter = pair.p[0].sync;                  // one claim counter and one abort flag for the launch
   int* const abortflag = pair.p[0].sync + 1;
   int maxtotal = 0;
   for( int q = 0; q < pair.count; ++q )
      maxtotal = max(maxtotal, pair.p[q].winv ? pair.p[q].T * pair.p[q].T + pair.p[q].whelp : pair.p[q].T * (pair.p[q].T + 1) / 2);

   for( ;; )
   {
      __syncthreads();                                   // the previous tile is done with shared memory and s_tile
      if( tid == 0 ) s_tile = atomicAdd(counter, 1);
      __syncthreads();
      if( s_tile >= maxtotal * pair.count ) break;
      const DagArgs& a = pair.p[s_tile % pair.count];
      const int t = s_tile / pair.count;
      const int T = a.T, n = a.n, total = a.winv ? T * T + a.whelp : T * (T + 1) / 2;     // whelp: number of helper tasks (0: none)
      int* const ready = a.sync + 2;
      if( t >= total ) continue;
      // claim order: for c = 0, 1, ...: the factor tiles of block column c, (c,c) first, then (with the inverse) the tiles of row c of W
      int i, j;
      bool wtile = false, whelper = false;
      int hid = 0, nh = 0, kt0 = 0;                     // helper id (first helper of a tile of W), helpers of the tile, first k-tile of the task
      if( a.winv && a.whelp > 0 )
      {
         // block c0 of the order: the factor slots of column c0, the helpers of row c0 of W, the tiles of row c0 of W
         int c0 = 0, rem = t, Hc = 0, HB = 0;            // Hc: helpers of row c0, HB: helpers of the rows before
         for( ;; )
         {
            const int size = T + Hc;
            if( rem < size ) break;
            rem -= size; HB += Hc; ++c0; Hc += dag_nhelp(c0);
         }
         if( rem < T - c0 ) { i = c0 + rem; j = c0; }
         else if( rem < T - c0 + Hc )
         {
            int hidx = rem - (T - c0);
            whelper = wtile = true;
            i = c0; hid = HB + hidx;
            for( j = 0; j < i; ++j ) { const int q = dag_nhelp(i - j); if( hidx < q ) break; hidx -= q; }
            kt0 = hidx * DAG_WSEG;
         }
         else
         {
            wtile = true;
            i = c0; j = rem - (T - c0) - Hc;
            hid = HB;
            for( int jj = 0; jj < j; ++jj ) hid += dag_nhelp(i - jj);
            nh = dag_nhelp(i - j);
            kt0 = nh * DAG_WSEG;
         }
      }
      else if( a.winv )
      {
         // block c0 of the order: the factor slots of column c0, then the tiles of row c0 + wshift of W (block 0 also takes the rows
         // before that).  A tile of row r needs L_{r,k}, k < r: tile (r, r-1) comes from the chain, tile (r, r-2) from block r - 2,
         // and the rows of W above it sit in earlier blocks as well - with wshift <= 2 every dependency still points backwards.
         // The long row sums of the last rows get two chain steps more to stream before the chain ends.
         const int sh = a.wshift;
         int c0 = 0, rem = t;
         for( ;; )
         {
            const int size = (T - c0) + ((c0 + sh <= T - 1) ? c0 + sh : 0) + ((c0 == 0) ? sh * (sh - 1) / 2 : 0);
            if( rem < size ) break;
            rem -= size; ++c0;
         }
         if( rem < T - c0 ) { i = c0 + rem; j = c0; }
         else
         {
            rem -= T - c0;
            wtile = true;
            i = c0 + sh;
            if( c0 == 0 )
            {
               for( int r = 1; r < sh; ++r ) { if( rem < r ) { i = r; break; } rem -= r; }
            }
            j = rem;
         }
      }
      else
      {
         int start = 0;
         j = 0;
         while( start + (T - j) <= t ) { start += T - j; ++j; }
         i = j + (t - start);
      }
      if( wtile && a.wpanel > 0 && (i / a.wpanel) != (j / a.wpanel) ) continue;       // W = L^-1 is wanted panel by panel only
      if( wtile )
      {
         // ---- W_ij = -W_ii sum_{k=j}^{i-1} L_ik W_kj  (i > j): the sum streams through the ring like the factor updates (L_ik as the
         // MN-contiguous operand, W_kj K-contiguous), its last term waits for W_{i-1,j}; then one 64^3 product with W_ii ----
         const int row0 = DAG_T * i, col0 = DAG_T * j;
         double sacc[8][2];
#pragma unroll
         for( int jt = 0; jt < 8; ++jt ) { sacc[jt][0] = 0.0; sacc[jt][1] = 0.0; }
         bool ok = true;
         auto wload = [&](int cidx, int slot)
         {
            const int k0 = DAG_T * j + cidx * DAG_BK;                 // global k of the chunk
            double* As = dsm + (size_t)slot * DAG_SLOT;
            double* Bs = As + DAG_BK * DAG_LDS;
            dag_load_chunk(As, a.A, a.lda, row0, k0, n, tid);
            // W rows [k0, k0 + 16) x the 64 columns of tile column j -> smem [c][k]
            for( int q = tid; q < DAG_T * (DAG_BK / 2); q += DAG_THREADS )
            {
               const int cc = q / (DAG_BK / 2), kq = (q % (DAG_BK / 2)) * 2;
               dag_cp16(Bs + cc * DAG_LDK + kq, a.Linv + (size_t)(col0 + cc) * a.ldi + k0 + kq, 16);
            }
         };
         auto wcompute = [&](int slot)
         {
            const double* As = dsm + (size_t)slot * DAG_SLOT;
            const double* Bs = As + DAG_BK * DAG_LDS;
#pragma unroll
            for( int kk = 0; kk < DAG_BK; kk += 4 )
            {
               const double av = As[(kk + fc) * DAG_LDS + 8 * w + fr];
#pragma unroll
               for( int jt = 0; jt < 8; ++jt )
               {
                  const double bv = Bs[(8 * jt + fr) * DAG_LDK + kk + fc];
                  dmma884(sacc[jt][0], sacc[jt][1], av, bv);
               }
            }
         };
         // flags of k-tile kt (global tile index j + kt): L_{i,j+kt} and W_{j+kt,j} (the diagonal tile publishes W_jj with L_jj)
         auto wwait = [&](int kt)
         {
            const int k = j + kt;
            return dag_wait(ready + i * T + k, (k == j) ? ready + j * T + j : ready + j * T + k, abortflag, s_abort, pair.p[0].watchdog);
         };
         // this task adds the k-tiles kt0 .. nkt - 1 of the sum (a helper: its piece of DAG_WSEG k-tiles; chunk indices below are local)
         const int nkt = whelper ? kt0 + DAG_WSEG : i - j, nchunks = 4 * (nkt - kt0 - 1);
         int* const hflag = ready + T * T + 2 * T;
         int wupto = 0;
         auto wissue = [&](int cidx)
         {
            const int kt = kt0 + (cidx >> 2);
            if( (cidx & 3) == 0 && kt >= wupto )
            {
               // flags of k-tile kt (tile index j + kt): L_{i,j+kt} at ready[i T + j + kt], W_{j+kt,j} at ready[j T + j + kt] (kt = 0: the diagonal tile)
               ok = dag_wait_upto(ready + i * T + j, ready + j * T + j, kt, nkt, abortflag, s_abort, s_upto, pair.p[0].watchdog) && ok;
               wupto = s_upto;
            }
            if( ok ) wload(4 * kt0 + cidx, cidx % DAG_STAGES);
         };
#pragma unroll 1
         for( int sidx = 0; sidx < DAG_STAGES - 1; ++sidx )
         {
            if( sidx < nchunks ) wissue(sidx);
            asm volatile("cp.async.commit_group;\n" ::);
         }
#pragma unroll 1
         for( int cidx = 0; cidx < nchunks; ++cidx )
         {
            asm volatile("cp.async.wait_group %0;\n" :: "n"(DAG_STAGES - 2));
            __syncthreads();
            if( cidx + DAG_STAGES - 1 < nchunks ) wissue(cidx + DAG_STAGES - 1);
            asm volatile("cp.async.commit_group;\n" ::);
            wcompute(cidx % DAG_STAGES);
         }
         asm volatile("cp.async.wait_group 0;\n" ::);
         __syncthreads();
         if( ok )
         {
            if( nkt - 1 >= wupto ) ok = wwait(nkt - 1);
            if( ok )
            {
#pragma unroll
               for( int q = 0; q < 4; ++q ) wload(4 * (nkt - 1) + q, q);      // (global chunk index)
               asm volatile("cp.async.commit_group;\n" ::);
               asm volatile("cp.async.wait_group 0;\n" ::);
               __syncthreads();
#pragma unroll
               for( int q = 0; q < 4; ++q ) wcompute(q);
            }
         }
         __syncthreads();
         if( !ok ) break;
         if( whelper )
         {
            // the piece goes to its scratch tile (thread-major layout: the tile's own task reads it back the same way)
            double* sc = a.wscratch + (size_t)hid * DAG_T * DAG_T;
#pragma unroll
            for( int jt = 0; jt < 8; ++jt )
            {
               sc[(2 * jt) * DAG_THREADS + tid] = sacc[jt][0];
               sc[(2 * jt + 1) * DAG_THREADS + tid] = sacc[jt][1];
            }
            __threadfence();
            __syncthreads();
            if( tid == 0 ) st_release(hflag + hid, 1);
            continue;
         }
         for( int q = 0; q < nh && ok; ++q )
         {
            ok = dag_wait(hflag + hid + q, hflag + hid + q, abortflag, s_abort, pair.p[0].watchdog);
            if( !ok ) break;
            const double* sc = a.wscratch + (size_t)(hid + q) * DAG_T * DAG_T;
#pragma unroll
            for( int jt = 0; jt < 8; ++jt )
            {
               sacc[jt][0] += __ldcg(sc + (2 * jt) * DAG_THREADS + tid);
               sacc[jt][1] += __ldcg(sc + (2 * jt + 1) * DAG_THREADS + tid);
            }
         }
         if( ok ) ok = dag_wait(ready + i * T + i, ready + i * T + i, abortflag, s_abort, pair.p[0].watchdog);
         if( !ok ) break;
         double* Ws = dsm;                               // [k][r] = W_ii[r][k]
         double* Ss = dsm + DAG_T * DAG_LDS;             // [k][c] = S[k][c]
         const double* Wi = a.Wd + (size_t)i * DAG_T * DAG_T;
         for( int q = tid; q < DAG_T * (DAG_T / 2); q += DAG_THREADS )
         {
            const int k = q / (DAG_T / 2), r = (q % (DAG_T / 2)) * 2;
            dag_cp16(Ws + k * DAG_LDS + r, Wi + (size_t)k * DAG_T + r, 16);
         }
         asm volatile("cp.async.commit_group;\n" ::);
#pragma unroll
         for( int jt = 0; jt < 8; ++jt )
            *reinterpret_cast<double2*>(Ss + (8 * w + fr) * DAG_LDS + 8 * jt + 2 * fc) = make_double2(sacc[jt][0], sacc[jt][1]);
         asm volatile("cp.async.wait_group 0;\n" ::);
         __syncthreads();
         double d[8][2];
#pragma unroll
         for( int jt = 0; jt < 8; ++jt ) { d[jt][0] = 0.0; d[jt][1] = 0.0; }
#pragma unroll
         for( int k0 = 0; k0 < DAG_T; k0 += 4 )
         {
            if( k0 > 8 * w + 7 ) continue;                // W_ii is lower triangular: W[r][k] = 0 for k > r
            const double av = -Ws[(k0 + fc) * DAG_LDS + 8 * w + fr];
#pragma unroll
            for( int jt = 0; jt < 8; ++jt )
            {
               const double bv = Ss[(k0 + fc) * DAG_LDS + 8 * jt + fr];
               dmma884(d[jt][0], d[jt][1], av, bv);
            }
         }
         const int row = row0 + 8 * w + fr;
         if( row < n )
         {
#pragma unroll
            for( int jt = 0; jt < 8; ++jt )
            {
               const int col = col0 + 8 * jt + 2 * fc;
               a.Linv[(size_t)col * a.ldi + row] = d[jt][0];
               a.Linv[(size_t)(col + 1) * a.ldi + row] = d[jt][1];
            }
         }
         __threadfence();
         __syncthreads();
         if( tid == 0 ) st_release(ready + j * T + i, 1);
         continue;
      }
      // chain variant: task 0 is the chain itself; the slots of the other diagonal tiles and of the tiles (j+1, j) carry the
      // pre-updates D_{j+1} (diagonal tile (j+1, j+1) with the k-tiles < j) and P_{j+1,j}, stored back in place
      int nk = j;                                        // k-tiles of the update
      bool pretask = false;
      if constexpr( CHAIN )
      {
         if( i == j )
         {
            if( j == 0 )
            {
               if( !dag_chain(a, dsm, abortflag, s_abort, sbad, s_pref, pair.p[0].watchdog) ) break;
               continue;
            }
            if( j + 1 >= T ) continue;
            i = j + 1; j = j + 1; pretask = true;
         }
         else if( i == j + 1 )
         {
            if( j == 0 ) continue;
            pretask = true;
         }
      }
      const bool diag = (i == j);
      const int row0 = DAG_T * i, col0 = DAG_T * j;
      long long* const dbg = (a.dbg != nullptr && !CHAIN && tid == 0 && (diag || i == j + 1)) ? a.dbg + 8 * (2 * j + (diag ? 0 : 1)) : nullptr;
      if( dbg ) dbg[0] = dag_now();                     // claimed

      // ---- C = A_ij (diagonal tile: lower part, identity on the padding) ----
      double c[8][2];
#pragma unroll
      for( int jt = 0; jt < 8; ++jt )
      {
         const int row = row0 + 8 * w + fr;
#pragma unroll
         for( int e = 0; e < 2; ++e )
         {
            const int col = col0 + 8 * jt + 2 * fc + e;
            double v = 0.0;
            if( row < n && col < n && (!diag || row >= col) ) v = a.A[(size_t)col * a.lda + row];
            if( diag && row == col && row >= n ) v = 1.0;
            c[jt][e] = v;
         }
      }

      // ---- C -= sum_k L_ik L_jk' ----
      // k-tiles 0 .. j-2 stream through the cp.async ring (their flags are usually set long before); the LAST k-tile, the one the
      // critical chain waits for, is fetched in one go after its flags - four chunks in flight at once instead of one per ring turn
      const int nchunks = 4 * max(nk - 1, 0);
      bool ok = true;
      auto load_chunk = [&](int cidx, int slot)
      {
         const int kt = cidx >> 2, k0 = DAG_T * kt + (cidx & 3) * DAG_BK;
         double* As = dsm + (size_t)slot * DAG_SLOT;
         dag_load_chunk(As, a.A, a.lda, row0, k0, n, tid);
         if( !diag ) dag_load_chunk(As + DAG_BK * DAG_LDS, a.A, a.lda, col0, k0, n, tid);
      };
      auto compute_chunk = [&](int slot)
      {
         const double* As = dsm + (size_t)slot * DAG_SLOT;
         const double* Bs = diag ? As : As + DAG_BK * DAG_LDS;
#pragma unroll
         for( int kk = 0; kk < DAG_BK; kk += 4 )
         {
            const double av = -As[(kk + fc) * DAG_LDS + 8 * w + fr];
#pragma unroll
            for( int jt = 0; jt < 8; ++jt )
            {
               if( diag && jt > w ) continue;
               const double bv = Bs[(kk + fc) * DAG_LDS + 8 * jt + fr];
               dmma884(c[jt][0], c[jt][1], av, bv);
            }
         }
      };
      int upto = 0;                                     // k-tiles below this index are known to be ready
      auto issue = [&](int cidx)
      {
         if( (cidx & 3) == 0 && (cidx >> 2) >= upto )
         {
            ok = dag_wait_upto(ready + i * T, ready + j * T, cidx >> 2, nk, abortflag, s_abort, s_upto, pair.p[0].watchdog) && ok;
            upto = s_upto;
         }
         if( ok ) load_chunk(cidx, cidx % DAG_STAGES);
      };
#pragma unroll 1
      for( int sidx = 0; sidx < DAG_STAGES - 1; ++sidx )
      {
         if( sidx < nchunks ) issue(sidx);
         asm volatile("cp.async.commit_group;\n" ::);
      }
#pragma unroll 1
      for( int cidx = 0; cidx < nchunks; ++cidx )
      {
         asm volatile("cp.async.wait_group %0;\n" :: "n"(DAG_STAGES - 2));
         __syncthreads();
         if( cidx + DAG_STAGES - 1 < nchunks ) issue(cidx + DAG_STAGES - 1);
         asm volatile("cp.async.commit_group;\n" ::);
         compute_chunk(cidx % DAG_STAGES);
      }
      asm volatile("cp.async.wait_group 0;\n" ::);
      __syncthreads();
      if( nk > 0 && ok )
      {
         if( nk - 1 >= upto ) ok = dag_wait(ready + i * T + (nk - 1), ready + j * T + (nk - 1), abortflag, s_abort, pair.p[0].watchdog);
         if( ok )
         {
#pragma unroll
            for( int q = 0; q < 4; ++q ) load_chunk(4 * (nk - 1) + q, q);
            asm volatile("cp.async.commit_group;\n" ::);
            asm volatile("cp.async.wait_group 0;\n" ::);
            __syncthreads();
#pragma unroll
            for( int q = 0; q < 4; ++q ) compute_chunk(q);
         }
      }
      __syncthreads();
      if( !ok ) break;
      if( dbg ) dbg[1] = dag_now();                     // updates done

      if( pretask )
      {
         const int row = row0 + 8 * w + fr;
         if( row < n )
         {
#pragma unroll
            for( int jt = 0; jt < 8; ++jt )
#pragma unroll
               for( int e = 0; e < 2; ++e )
               {
                  const int col = col0 + 8 * jt + 2 * fc + e;
                  if( col < n && (!diag || row >= col) ) a.A[(size_t)col * a.lda + row] = c[jt][e];
               }
         }
         __threadfence();
         __syncthreads();
         if( tid == 0 ) st_release(ready + T * T + 2 * i + (diag ? 1 : 0), 1);
         continue;
      }
      if( !CHAIN && diag )
      {
         const int nb = min(DAG_T, n - col0);
         const LeafCtx<DAG_T> X(dsm);
         leaf_factor<DAG_T>(X, c, nb, a.A + (size_t)col0 * a.lda + col0, a.lda, a.info, col0, sbad);
         if( dbg ) dbg[2] = dag_now();                  // factor done
         leaf_invert<DAG_T>(X);
         if( dbg ) dbg[3] = dag_now();                  // inverse done
         constexpr int LD = LeafCtx<DAG_T>::LD;
         double* Wj = a.Wd + (size_t)j * DAG_T * DAG_T;
         for( int e = tid; e < DAG_T * DAG_T; e += DAG_THREADS )
         {
            const int r = e % DAG_T, q = e / DAG_T;
            const double v = (r < nb && q < nb && r >= q) ? X.G[q * LD + r] : 0.0;
            Wj[(size_t)q * DAG_T + r] = v;
            if( a.Linv != nullptr && r < nb && q < nb ) a.Linv[(size_t)(col0 + q) * a.ldi + col0 + r] = v;
         }
      }
      else
      {
         // ---- L_ij = C W_jj' ----
         ok = dag_wait(ready + j * T + j, ready + j * T + j, abortflag, s_abort, pair.p[0].watchdog);
         if( !ok ) break;
         if( dbg ) dbg[2] = dag_now();                  // saw the diagonal tile
         double* Cs = dsm;                               // [k][r]
         double* Ws = dsm + DAG_T * DAG_LDS;             // [k][c] = W_jj[c][k]
         const double* Wj = a.Wd + (size_t)j * DAG_T * DAG_T;
         for( int q = tid; q < DAG_T * (DAG_T / 2); q += DAG_THREADS )
         {
            const int k = q / (DAG_T / 2), r = (q % (DAG_T / 2)) * 2;
            dag_cp16(Ws + k * DAG_LDS + r, Wj + (size_t)k * DAG_T + r, 16);
         }
         asm volatile("cp.async.commit_group;\n" ::);
#pragma unroll
         for( int jt = 0; jt < 8; ++jt )
         {
            Cs[(8 * jt + 2 * fc) * DAG_LDS + 8 * w + fr] = c[jt][0];
            Cs[(8 * jt + 2 * fc + 1) * DAG_LDS + 8 * w + fr] = c[jt][1];
         }
         asm volatile("cp.async.wait_group 0;\n" ::);
         __syncthreads();
         double d[8][2];
#pragma unroll
         for( int jt = 0; jt < 8; ++jt ) { d[jt][0] = 0.0; d[jt][1] = 0.0; }
#pragma unroll
         for( int k0 = 0; k0 < DAG_T; k0 += 4 )
         {
            const double av = Cs[(k0 + fc) * DAG_LDS + 8 * w + fr];
#pragma unroll
            for( int jt = 0; jt < 8; ++jt )
            {
               if( k0 >= 8 * jt + 8 ) continue;          // W_jj is lower triangular: W[c][k] = 0 for k > c
               const double bv = Ws[(k0 + fc) * DAG_LDS + 8 * jt + fr];
               dmma884(d[jt][0], d[jt][1], av, bv);
            }
         }
         if( dbg ) dbg[3] = dag_now();                  // product done
         const int row = row0 + 8 * w + fr;
         if( row < n )
         {
#pragma unroll
            for( int jt = 0; jt < 8; ++jt )
            {
               const int col = col0 + 8 * jt + 2 * fc;
               a.A[(size_t)col * a.lda + row] = d[jt][0];
               a.A[(size_t)(col + 1) * a.lda + row] = d[jt][1];
            }
         }
      }
      if( dbg ) dbg[4] = dag_now();                     // stores issued
      __threadfence();
      __syncthreads();
      if( tid == 0 ) st_release(ready + i * T + j, 1);
      if( dbg ) dbg[5] = dag_now();                     // published
   }
}

__global__ void copy2d_kernel(int m, int n, const double* __restrict__ src, int lds, double* __restrict__ dst, int ldd)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   int j = blockIdx.y;
   if( i < m ) dst[(size_t)j * ldd + i] = src[(size_t)j * lds + i];
}

cudaError_t copy2d(cudaStream_t st, int m, int n, const double* src, int lds, double* dst, int ldd)
{
   if( m <= 0 || n <= 0 ) return cudaSuccess;
   dim3 grid(ceil_div(m, 256), n);
   copy2d_kernel<<<grid, 256, 0, st>>>(m, n, src, lds, dst, ldd);
   count_launch();
   return cudaGetLastError();
}

// leaf order of the recursion: 128 (default) or 64 (SDPCUDA_LEAF=64, tests)
int leaf_config()
{
   const char* e = getenv("SDPCUDA_LEAF");      // read on every call (tests switch it between solves)
   if( e != nullptr && strcmp(e, "64") == 0 ) return 64;
   return 128;
}

cudaError_t launch_diag(cudaStream_t st, int mode, int nb, double* A, int lda, double* Linv, int ldi, double* diaginv, int* info, int off)
{
   ProfScope prof(st, PROF_DIAG, (mode == 0 ? 1.0 : 0.0) * nb * (double)nb * nb / 3.0 + ((Linv || diaginv) ? nb * (double)nb * nb / 3.0 : 0.0));
   static bool configured[64] = {false};        // per-device function attributes
   int dev = 0;
   SDPK_CUDA_CHECK( cudaGetDevice(&dev) );
   if( !configured[dev & 63] )
   {
      SDPK_CUDA_CHECK( cudaFuncSetAttribute(leaf_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)leaf_smem<64>()) );
      SDPK_CUDA_CHECK( cudaFuncSetAttribute(leaf_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)leaf_smem<128>()) );
      configured[dev & 63] = true;
   }
   if( nb <= 64 )
      leaf_kernel<64><<<1, 256, leaf_smem<64>(), st>>>(mode, nb, A, lda, Linv, ldi, diaginv, info, off, g_diag_dbg);
   else if( nb <= 128 && diaginv == nullptr )
      leaf_kernel<128><<<1, 512, leaf_smem<128>(), st>>>(mode, nb, A, lda, Linv, ldi, diaginv, info, off, g_diag_dbg);
   else
      return cudaErrorInvalidValue;
   count_launch();
   return cudaGetLastError();
}

// largest block handled by one leaf launch (the packed 64 x 64 diagonal inverses of the substitution path need 64)
int leaf_order(const double* diaginv) { return (diaginv == nullptr && leaf_config() == 128) ? 128 : NB; }

int split_point(int n, int leaf)
{
   int n1 = round_up(n / 2, leaf);
   if( n1 >= n ) n1 -= leaf;
   return n1;
}

cudaError_t chol_rec(cudaStream_t st, int n, double* A, int lda, double* Linv, int ldi, double* diaginv, double* work,
   int ldw, int* d_info, int off)
{
   const int leaf = leaf_order(diaginv);
   if( n <= leaf )
   {
      return launch_diag(st, 0, n, A, lda, Linv, ldi, diaginv ? diaginv + (size_t)(off / NB) * NB * NB : nullptr, d_info, off);
   }
   const int n1 = split_point(n, leaf), n2 = n - n1;
   double* A21 = A + n1;
   double* A22 = A + (size_t)n1 * lda + n1;
   // without a wanted inverse the inverse of the leading block is still needed for L21: it goes to the work space
   double* Li11 = Linv ? Linv : work;
   const int ldi11 = Linv ? ldi : ldw;
   double* wrk = Linv ? work : work + (size_t)ldw * n1;      // scratch for L21 (n2 x n1), behind Li11 if that lives in work

   cudaError_t e;
   if( Linv )
   {
      e = chol_rec(st, n1, A, lda, Linv, ldi, diaginv, work, ldw, d_info, off);
   }
   else
   {
      // factor A11 and form its full inverse in the work space (recursively needs its own scratch behind it)
      SDPK_CUDA_CHECK( cudaMemsetAsync(Li11, 0, sizeof(double) * (size_t)ldw * n1, st) );
      e = chol_rec(st, n1, A, lda, Li11, ldi11, diaginv, work + (size_t)ldw * n1, ldw, d_info, off);
   }
   if( e != cudaSuccess ) return e;
   // L21 = A21 * Linv11'
   SDPK_CUDA_CHECK( gemm(st, false, true, n2, n1, n1, 1.0, A21, lda, 0, Li11, ldi11, 0, 0.0, wrk, ldw, 0, 1, GEMM_KHI_N) );
   SDPK_CUDA_CHECK( copy2d(st, n2, n1, wrk, ldw, A21, lda) );
   // A22 -= L21 L21'
   SDPK_CUDA_CHECK( gemm(st, false, true, n2, n2, n1, -1.0, A21, lda, 0, A21, lda, 0, 1.0, A22, lda, 0, 1, GEMM_LOWER) );
   if( Linv )
   {
      double* Li22 = Linv + (size_t)n1 * ldi + n1;
      double* Li21 = Linv + n1;
      SDPK_CUDA_CHECK( chol_rec(st, n2, A22, lda, Li22, ldi, diaginv, work, ldw, d_info, off + n1) );
      // Linv21 = -Linv22 * (L21 * Linv11)
      SDPK_CUDA_CHECK( gemm(st, false, false, n2, n1, n1, 1.0, A21, lda, 0, Linv, ldi, 0, 0.0, work, ldw, 0, 1, GEMM_KLO_N) );
      SDPK_CUDA_CHECK( gemm(st, false, false, n2, n1, n2, -1.0, Li22, ldi, 0, work, ldw, 0, 0.0, Li21, ldi, 0, 1, GEMM_KHI_M) );
   }
   else
   {
      SDPK_CUDA_CHECK( chol_rec(st, n2, A22, lda, nullptr, 0, diaginv, work, ldw, d_info, off + n1) );
   }
   return cudaSuccess;
}

cudaError_t trtri_rec(cudaStream_t st, int n, const double* L, int ldl, double* Linv, int ldi, double* work, int ldw)
{
   const int leaf = leaf_order(nullptr);
   if( n <= leaf )
   {
      return launch_diag(st, 1, n, const_cast<double*>(L), ldl, Linv, ldi, nullptr, nullptr, 0);
   }
   const int n1 = split_point(n, leaf), n2 = n - n1;
   const double* L21 = L + n1;
   double* Li22 = Linv + (size_t)n1 * ldi + n1;
   SDPK_CUDA_CHECK( trtri_rec(st, n1, L, ldl, Linv, ldi, work, ldw) );
   SDPK_CUDA_CHECK( trtri_rec(st, n2, L + (size_t)n1 * ldl + n1, ldl, Li22, ldi, work, ldw) );
   SDPK_CUDA_CHECK( gemm(st, false, false, n2, n1, n1, 1.0, L21, ldl, 0, Linv, ldi, 0, 0.0, work, ldw, 0, 1, GEMM_KLO_N) );
   SDPK_CUDA_CHECK( gemm(st, false, false, n2, n1, n2, -1.0, Li22, ldi, 0, work, ldw, 0, 0.0, Linv + n1, ldi, 0, 1, GEMM_KHI_M) );
   return cudaSuccess;
}



// 0: recursive kernel chain (round 1), 1: tile-DAG kernel for n > 128 (default); SDPCUDA_CHOL=rec|dag overrides (tests compare both)
int chol_variant()
{
   const char* e = getenv("SDPCUDA_CHOL");
   if( e != nullptr && strcmp(e, "rec") == 0 ) return 0;
   return 1;
}

struct DagProblem { int n; double* A; int lda; double* Linv; int ldi; double* diaginv; double* work; int ldw; int* d_info; int wpanel = 0; };

cudaError_t potrf_dag_launch(cudaStream_t st, const DagProblem* pr, int count)
{
   static bool configured[64] = {false};
   static int nsm[64] = {0};
   int dev = 0;
   SDPK_CUDA_CHECK( cudaGetDevice(&dev) );
   if( !configured[dev & 63] )
   {
      SDPK_CUDA_CHECK( cudaFuncSetAttribute(potrf_dag_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) );
      SDPK_CUDA_CHECK( cudaFuncSetAttribute(potrf_dag_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DAG_SMEM) );
      SDPK_CUDA_CHECK( cudaDeviceGetAttribute(&nsm[dev & 63], cudaDevAttrMultiProcessorCount, dev) );
      configured[dev & 63] = true;
   }
   // the inverse factor: inside the kernel where the factorisation is bound by its dependency chain (the tiles of W hide behind
   // it: n = 2000 1.05 -> 0.83 ms), level by level with batched GEMMs afterwards for large orders (7140: 10.3 vs 12.4 ms in the
   // kernel); SDPCUDA_CHOL_INV=levels|kernel overrides
   const char* ie = getenv("SDPCUDA_CHOL_INV");
   DagPair pair;
   pair.count = count;
   int total = 0, maxn = 0;
   for( int q = 0; q < count; ++q ) maxn = std::max(maxn, pr[q].n);
   // the chain-in-one-CTA variant where the dependency chain is the bound (SDPCUDA_CHOL_CHAIN=0: one task per tile everywhere)
   const char* ce = getenv("SDPCUDA_CHOL_CHAIN");
   const bool chain = (maxn <= 3072) && !(ce != nullptr && strcmp(ce, "0") == 0);
   double flops = 0.0;
   bool inkernel[2] = {false, false};
   for( int q = 0; q < count; ++q )
   {
      const DagProblem& P = pr[q];
      const int T = ceil_div(P.n, DAG_T);
      double* Wd = P.diaginv != nullptr ? P.diaginv : P.work + (size_t)P.ldw * P.n;
      int* sync = reinterpret_cast<int*>(P.work + (size_t)P.ldw * P.n + (P.diaginv != nullptr ? 0 : (size_t)T * DAG_T * DAG_T));
      inkernel[q] = (P.Linv != nullptr) && (P.wpanel > 0 || (ie != nullptr ? strcmp(ie, "levels") != 0 : P.n <= 3072));
      // helper tasks of the row sums of W: their scratch tiles live in the first n columns of the work space (free in this path)
      int nhelp = 0;
      {
         // OFF by default (SDPCUDA_DAG_WHELP=1): measured at n = 2000 the kernel takes 0.81 instead of 0.76 ms with the helpers - the
         // 0.11 ms that the inverse factor costs on top of the factorisation are mostly a SLOWER CHAIN (714 instead of 630 us: the
         // tasks that feed it wait longer for a CTA, 1.6 instead of 0.6 us per step), not a tail behind it (48 us), and more tasks
         // in the order make exactly that worse (profiles/r2_chain_cta_and_leaf_formulations.log)
         const char* he = getenv("SDPCUDA_DAG_WHELP");
         if( inkernel[q] && P.wpanel == 0 && (he != nullptr && he[0] == '1') )
         {
            for( int i = 0; i < T; ++i ) nhelp += dag_rowhelp(i);
            if( (size_t)nhelp * DAG_T * DAG_T > (size_t)P.ldw * P.n ) nhelp = 0;
         }
      }
      if( (size_t)T * DAG_T * DAG_T + (size_t)(T * T + 2 * T + 2 + nhelp + 1) / 2 > (size_t)P.ldw * 2 * CHOL_LEAF_MAX ) return cudaErrorInvalidValue;
      SDPK_CUDA_CHECK( cudaMemsetAsync(sync, 0, sizeof(int) * (size_t)(T * T + 2 * T + 2 + nhelp), st) );
      DagArgs& a = pair.p[q];
      a.n = P.n; a.T = T; a.A = P.A; a.lda = P.lda; a.Linv = P.Linv; a.ldi = P.ldi; a.Wd = Wd; a.sync = sync; a.info = P.d_info;
      a.dbg = (q == 0) ? g_diag_dbg : nullptr;
      {
         const char* we = getenv("SDPCUDA_DAG_WATCHDOG_S");       // compute-sanitizer slows the kernel down by orders of magnitude: 0 turns the limit off
         a.watchdog = (long long)((we != nullptr ? atof(we) : 2.0) * 2.0e9);
      }
      a.winv = inkernel[q] ? 1 : 0;
      a.whelp = nhelp; a.wscratch = P.work;
      a.wpanel = P.wpanel;
      a.chain = chain ? 1 : 0;
      {
         const char* we2 = getenv("SDPCUDA_DAG_WSHIFT");
         a.wshift = (we2 != nullptr) ? std::max(0, std::min(2, atoi(we2))) : 0;   // measured: earlier claims take CTAs from the tasks the chain waits for (2000: 0.77 / 0.81 / 0.86 ms for 0 / 1 / 2)
         if( T < 4 || !chain ) a.wshift = 0;          // without the chain CTA tile (r, r-1) is a task of block r - 1 itself
      }
      total = std::max(total, inkernel[q] ? T * T + nhelp : T * (T + 1) / 2);
      maxn = std::max(maxn, P.n);
      flops += (double)P.n * P.n * P.n / 3.0 * (inkernel[q] ? 2.0 : 1.0);
   }
   if( count == 1 ) pair.p[1] = pair.p[0];
   {
      ProfScope prof(st, PROF_DIAG, flops);
      // up to about n = 3000 the factorisation is bound by its dependency chain, not by flops: one CTA per SM is plenty (and the
      // chain runs faster on an SM of its own)
      const char* pe = getenv("SDPCUDA_DAG_PER_SM");
      const int per_sm = (pe != nullptr && atoi(pe) > 0) ? atoi(pe) : ((maxn <= 3072) ? 1 : 2);
      int grid = std::min(total * count, per_sm * nsm[dev & 63]);
      size_t smem_chain = DAG_SMEM_CHAIN;
      {
         const char* ge = getenv("SDPCUDA_DAG_GRID");
         const char* se = getenv("SDPCUDA_DAG_SMEM_KB");
         if( ge != nullptr && atoi(ge) > 0 ) grid = std::min(grid, atoi(ge));
         if( se != nullptr && atoi(se) > 0 ) smem_chain = std::max(smem_chain, (size_t)atoi(se) * 1024);
      }
      if( chain ) potrf_dag_kernel<true><<<grid, DAG_THREADS, smem_chain, st>>>(pair);
      else potrf_dag_kernel<false><<<grid, DAG_THREADS, DAG_SMEM, st>>>(pair);
      count_launch();
      SDPK_CUDA_CHECK( cudaGetLastError() );
   }
   for( int q = 0; q < count; ++q )
   {
      const DagProblem& P = pr[q];
      if( P.Linv == nullptr || inkernel[q] ) continue;
      const int n = P.n, lda = P.lda, ldi = P.ldi, ldw = P.ldw;
      double* const A = P.A; double* const Linv = P.Linv; double* const work = P.work;
      for( int s = DAG_T; s < n; s *= 2 )
      {
         const int full = n / (2 * s);                       // pairs with two complete s-blocks
         const long long sl = (long long)2 * s * ((long long)lda + 1), si = (long long)2 * s * ((long long)ldi + 1);
         if( full > 0 )
         {
            SDPK_CUDA_CHECK( gemm(st, false, false, s, s, s, 1.0, A + s, lda, sl, Linv, ldi, si, 0.0, work, ldw, 2 * s, full, GEMM_KLO_N) );
            SDPK_CUDA_CHECK( gemm(st, false, false, s, s, s, -1.0, Linv + (size_t)s * ldi + s, ldi, si, work, ldw, 2 * s, 0.0, Linv + s, ldi, si, full, GEMM_KHI_M) );
         }
         const int o = 2 * s * full, n2 = n - (o + s);       // a last pair whose second block is shorter
         if( n2 > 0 )
         {
            SDPK_CUDA_CHECK( gemm(st, false, false, n2, s, s, 1.0, A + (size_t)o * lda + o + s, lda, 0, Linv + (size_t)o * ldi + o, ldi, 0, 0.0,
               work + o, ldw, 0, 1, GEMM_KLO_N) );
            SDPK_CUDA_CHECK( gemm(st, false, false, n2, s, n2, -1.0, Linv + (size_t)(o + s) * ldi + o + s, ldi, 0, work + o, ldw, 0, 0.0,
               Linv + (size_t)o * ldi + o + s, ldi, 0, 1, GEMM_KHI_M) );
         }
      }
   }
   return cudaSuccess;
}

// Cholesky by the tile-DAG kernel and (if wanted) the inverse factor: the diagonal 64-blocks of W = L^-1 come out of the leaf code,
// the other tiles either inside the kernel or, for s = 64, 128, ..., all pairs of adjacent s-blocks joined at once,
// W21 = -W22 (L21 W11), as two batched GEMMs per level (plus two for a shorter last pair).  work: ldw x (n + 2 CHOL_LEAF_MAX) doubles;
// the packed diagonal inverses and the flags of the kernel live in its last 2 CHOL_LEAF_MAX columns.
cudaError_t potrf_dag(cudaStream_t st, int n, double* A, int lda, double* Linv, int ldi, double* diaginv, double* work, int ldw, int* d_info)
{
   DagProblem P = {n, A, lda, Linv, ldi, diaginv, work, ldw, d_info};
   return potrf_dag_launch(st, &P, 1);
}

} // namespace

cudaError_t potrf_lower(cudaStream_t st, int n, double* A, int lda, double* Linv, int ldi, double* diaginv, double* work, int ldw, int* d_info)
{
   if( n <= 0 ) return cudaSuccess;
   if( Linv )
      SDPK_CUDA_CHECK( cudaMemsetAsync(Linv, 0, sizeof(double) * (size_t)ldi * n, st) );
   if( n > CHOL_LEAF_MAX && chol_variant() == 1 )
      return potrf_dag(st, n, A, lda, Linv, ldi, diaginv, work, ldw, d_info);
   return chol_rec(st, n, A, lda, Linv, ldi, diaginv, work, ldw, d_info, 0);
}

// two factorisations of the same order with inverse factors in ONE launch of the tile-DAG kernel (S and X of an interior-point
// iteration); orders the kernel does not take (n <= CHOL_LEAF_MAX, or SDPCUDA_CHOL=rec) run one after the other
cudaError_t potrf_lower_pair(cudaStream_t st, int n, double* A0, double* Linv0, double* work0, int* info0, double* A1, double* Linv1,
   double* work1, int* info1, int ld, int ldw)
{
   if( n <= 0 ) return cudaSuccess;
   if( n > CHOL_LEAF_MAX && chol_variant() == 1 )
   {
      SDPK_CUDA_CHECK( cudaMemsetAsync(Linv0, 0, sizeof(double) * (size_t)ld * n, st) );
      SDPK_CUDA_CHECK( cudaMemsetAsync(Linv1, 0, sizeof(double) * (size_t)ld * n, st) );
      DagProblem P[2] = {{n, A0, ld, Linv0, ld, nullptr, work0, ldw, info0}, {n, A1, ld, Linv1, ld, nullptr, work1, ldw, info1}};
      return potrf_dag_launch(st, P, 2);
   }
   SDPK_CUDA_CHECK( potrf_lower(st, n, A0, ld, Linv0, ld, nullptr, work0, ldw, info0) );
   return potrf_lower(st, n, A1, ld, Linv1, ld, nullptr, work1, ldw, info1);
}

cudaError_t trtri_lower(cudaStream_t st, int n, const double* L, int ldl, double* Linv, int ldi, double* work, int ldw)
{
   if( n <= 0 ) return cudaSuccess;
   SDPK_CUDA_CHECK( cudaMemsetAsync(Linv, 0, sizeof(double) * (size_t)ldi * n, st) );
   return trtri_rec(st, n, L, ldl, Linv, ldi, work, ldw);
}


// ---- large orders: right-looking blocked factorisation with one panel of look-ahead ---------------------------------------
// Panels of width PB.  Step k on the main stream: factor the diagonal block (recursive kernel chain above, with its inverse
// P_k), L(below, k) = A(below, k) P_k' (GEMM), update of the NEXT panel's columns with panel k (GEMM).  The update of all
// later columns with panel k (the bulk of the n^3/3 flops, one large lower-triangular GEMM) goes to the side stream, so that
// the latency-bound factorisation of panel k+1 hides behind it.  Dependencies: update_rest(k) needs the panel-k solve;
// update_next(k+1) needs update_rest(k).  The inverse of L is NOT formed: solves use the panel inverses (potrs_panels below).
namespace {

// out[i] = (bin ? bin[i] : 0) - sign * dot(column i of Mx restricted by mode, vin) for the columns i0 .. of one panel
// mode 0: rows [0, len); mode 1: rows >= local column index (lower triangular panel inverse);
// mode 2: rows <= local column index (transposed panel inverse)
__global__ void __launch_bounds__(256)
panel_dot_kernel(const double* __restrict__ Mx, int ld, int ncols, int len, int mode, const double* __restrict__ vin,
   const double* __restrict__ bin, double sign, double* __restrict__ out)
{
   __shared__ double part[8][8];
   const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
   const int i0 = blockIdx.x * 8;
   if( i0 >= ncols ) return;
   const int nc = min(8, ncols - i0);
   double acc[8];
#pragma unroll
   for( int c = 0; c < 8; ++c ) acc[c] = 0.0;
   int kbeg = 0, kend = len;
   if( mode == 1 ) kbeg = i0;                 // entries k >= i
   if( mode == 2 ) kend = min(len, i0 + 8);   // entries k <= i
   for( int k = kbeg + tid; k < kend; k += 512 )
   {
      const int k2 = k + 256;
      const bool in2 = k2 < kend;
      const double v0 = vin[k], v1 = in2 ? vin[k2] : 0.0;
      double m0[8], m1[8];
#pragma unroll
      for( int c = 0; c < 8; ++c )
      {
         const double* col = Mx + (size_t)(i0 + c) * ld;
         m0[c] = (c < nc) ? col[k] : 0.0;
         m1[c] = (c < nc && in2) ? col[k2] : 0.0;
      }
#pragma unroll
      for( int c = 0; c < 8; ++c )
      {
         const int i = i0 + c;
         const bool ok0 = (mode == 0) || (mode == 1 ? k >= i : k <= i);
         const bool ok1 = (mode == 0) || (mode == 1 ? k2 >= i : k2 <= i);
         acc[c] += (ok0 ? m0[c] * v0 : 0.0) + (ok1 ? m1[c] * v1 : 0.0);
      }
   }
#pragma unroll
   for( int c = 0; c < 8; ++c )
   {
      double v = acc[c];
#pragma unroll
      for( int o = 16; o > 0; o >>= 1 ) v += __shfl_xor_sync(0xffffffffu, v, o);
      if( lane == 0 ) part[wid][c] = v;
   }
   __syncthreads();
   if( tid < nc )
   {
      double v = 0.0;
#pragma unroll
      for( int q = 0; q < 8; ++q ) v += part[q][tid];
      out[i0 + tid] = (bin != nullptr ? bin[i0 + tid] : 0.0) - sign * v;
   }
}

__global__ void transpose_small_kernel(int n, const double* __restrict__ A, int lda, double* __restrict__ B, int ldb)
{
   __shared__ double tile[32][33];
   const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
   for( int r = threadIdx.y; r < 32; r += blockDim.y )
      tile[r][threadIdx.x] = (x < n && y0 + r < n) ? A[(size_t)(y0 + r) * lda + x] : 0.0;
   __syncthreads();
   const int xo = blockIdx.y * 32 + threadIdx.x, yo0 = blockIdx.x * 32;
   for( int r = threadIdx.y; r < 32; r += blockDim.y )
      if( xo < n && yo0 + r < n ) B[(size_t)(yo0 + r) * ldb + xo] = tile[threadIdx.x][r];
}

} // namespace

// The same result as potrf_lower_lookahead (L in A, packed inverses of the pb x pb diagonal blocks of L in pinv, their transposes in
// pinvT) from ONE launch of the tile-DAG kernel: the factorisation as usual, and of W = L^-1 only the tiles inside the diagonal
// panels (a.wpanel: a tile of W is a task only if its row and column tile lie in the same panel - the inverse of a diagonal block
// depends on that block alone).  The inverse tiles land in the first n columns of the work space (free in this path) and are copied
// out panel by panel.  m = 7140 (MkP-120): 8.3 -> about 6 ms per factorisation of the Schur complement.
cudaError_t potrf_lower_panels(cudaStream_t st, int pb, int n, double* A, int lda, double* pinv, double* pinvT, double* work, int ldw, int* d_info)
{
   if( n <= 0 ) return cudaSuccess;
   if( pb % DAG_T != 0 || n <= CHOL_LEAF_MAX || chol_variant() != 1 ) return cudaErrorNotSupported;
   const int nblk = ceil_div(n, pb);
   for( int k = 0; k < nblk; ++k )
   {
      const int j0 = k * pb, kb = min(pb, n - j0);
      SDPK_CUDA_CHECK( cudaMemset2DAsync(work + (size_t)j0 * ldw + j0, sizeof(double) * ldw, 0, sizeof(double) * kb, kb, st) );
      SDPK_CUDA_CHECK( cudaMemsetAsync(pinv + (size_t)k * pb * pb, 0, sizeof(double) * (size_t)pb * pb, st) );
   }
   DagProblem P = {n, A, lda, work, ldw, nullptr, work, ldw, d_info};
   P.wpanel = pb / DAG_T;
   SDPK_CUDA_CHECK( potrf_dag_launch(st, &P, 1) );
   for( int k = 0; k < nblk; ++k )
   {
      const int j0 = k * pb, kb = min(pb, n - j0);
      double* Pk = pinv + (size_t)k * pb * pb;
      SDPK_CUDA_CHECK( copy2d(st, kb, kb, work + (size_t)j0 * ldw + j0, ldw, Pk, pb) );
      dim3 grid(ceil_div(kb, 32), ceil_div(kb, 32)), block(32, 8);
      transpose_small_kernel<<<grid, block, 0, st>>>(kb, Pk, pb, pinvT + (size_t)k * pb * pb, pb);
      count_launch();
   }
   return cudaGetLastError();
}

cudaError_t potrf_lower_lookahead(cudaStream_t st, cudaStream_t side, cudaEvent_t* ev, int nev, int pb, int n, double* A, int lda,
   double* pinv, double* pinvT, double* work, int ldw, int* d_info)
{
   const int nblk = ceil_div(n, pb);
   if( 2 * nblk + 2 > nev ) return cudaErrorInvalidValue;
   cudaEvent_t* evT = ev;              // panel solve of step k done (main stream)
   cudaEvent_t* evB = ev + nblk;       // update of the later columns with panel k done (side stream)
   int lastB = -1;
   for( int k = 0; k < nblk; ++k )
   {
      const int j0 = k * pb, kb = min(pb, n - j0), rem = n - j0 - kb;
      double* Pk = pinv + (size_t)k * pb * pb;
      double* PkT = pinvT + (size_t)k * pb * pb;
      SDPK_CUDA_CHECK( cudaMemsetAsync(Pk, 0, sizeof(double) * (size_t)pb * pb, st) );
      SDPK_CUDA_CHECK( chol_rec(st, kb, A + (size_t)j0 * lda + j0, lda, Pk, pb, nullptr, work, ldw, d_info, j0) );
      {
         dim3 grid(ceil_div(kb, 32), ceil_div(kb, 32)), block(32, 8);
         transpose_small_kernel<<<grid, block, 0, st>>>(kb, Pk, pb, PkT, pb);
         count_launch();
      }
      if( rem == 0 ) break;
      const int j1 = j0 + kb, kb1 = min(pb, rem), rem2 = rem - kb1, j2 = j1 + kb1;
      double* A21 = A + (size_t)j0 * lda + j1;
      SDPK_CUDA_CHECK( gemm(st, false, true, rem, kb, kb, 1.0, A21, lda, 0, Pk, pb, 0, 0.0, work, ldw, 0, 1, GEMM_KHI_N) );
      SDPK_CUDA_CHECK( copy2d(st, rem, kb, work, ldw, A21, lda) );
      SDPK_CUDA_CHECK( cudaEventRecord(evT[k], st) );
      if( lastB >= 0 ) SDPK_CUDA_CHECK( cudaStreamWaitEvent(st, evB[lastB], 0) );
      // next panel: A(j1.., j1..j2) -= L(j1.., k) L(j1..j2, k)'
      SDPK_CUDA_CHECK( gemm(st, false, true, rem, kb1, kb, -1.0, A21, lda, 0, A21, lda, 0, 1.0, A + (size_t)j1 * lda + j1, lda, 0, 1, GEMM_LOWER) );
      if( rem2 > 0 )
      {
         SDPK_CUDA_CHECK( cudaStreamWaitEvent(side, evT[k], 0) );
         const double* L2 = A + (size_t)j0 * lda + j2;
         SDPK_CUDA_CHECK( gemm(side, false, true, rem2, rem2, kb, -1.0, L2, lda, 0, L2, lda, 0, 1.0, A + (size_t)j2 * lda + j2, lda, 0, 1, GEMM_LOWER) );
         SDPK_CUDA_CHECK( cudaEventRecord(evB[k], side) );
         lastB = k;
      }
   }
   if( lastB >= 0 ) SDPK_CUDA_CHECK( cudaStreamWaitEvent(st, evB[lastB], 0) );
   return cudaSuccess;
}

// b <- (L L')^-1 b with the panel inverses of potrf_lower_lookahead; LT = L' (n x n, ldl) for the contiguous row access of the
// forward sweep.  Two launches per panel and sweep, every one a set of dot products with contiguous columns.
cudaError_t potrs_panels(cudaStream_t st, int pb, int n, const double* L, const double* LT, int ldl, const double* pinv, const double* pinvT,
   double* b, double* tmp)
{
   const int nblk = ceil_div(n, pb);
   ProfScope prof(st, PROF_TRSV, 8.0 * n * (double)n);
   // forward: z_k = P_k (b_k - L(k, 0:j0) z(0:j0))
   for( int k = 0; k < nblk; ++k )
   {
      const int j0 = k * pb, kb = min(pb, n - j0);
      // (with j0 = 0 the first launch only copies: the second one must not read and write the same vector)
      panel_dot_kernel<<<ceil_div(kb, 8), 256, 0, st>>>(LT + (size_t)j0 * ldl, ldl, kb, j0, 0, b, b + j0, 1.0, tmp + j0);
      // row r of P_k = column r of P_k', entries <= r
      panel_dot_kernel<<<ceil_div(kb, 8), 256, 0, st>>>(pinvT + (size_t)k * pb * pb, pb, kb, kb, 2, tmp + j0, nullptr, -1.0, b + j0);
      count_launch(2);
   }
   // backward: x_k = P_k' (z_k - L(j1:, k)' x(j1:))
   for( int k = nblk - 1; k >= 0; --k )
   {
      const int j0 = k * pb, kb = min(pb, n - j0), j1 = j0 + kb, rem = n - j1;
      panel_dot_kernel<<<ceil_div(kb, 8), 256, 0, st>>>(L + (size_t)j0 * ldl + j1, ldl, kb, rem, 0, b + j1, b + j0, 1.0, tmp + j0);
      // column r of P_k, entries >= r
      panel_dot_kernel<<<ceil_div(kb, 8), 256, 0, st>>>(pinv + (size_t)k * pb * pb, pb, kb, kb, 1, tmp + j0, nullptr, -1.0, b + j0);
      count_launch(2);
   }
   return cudaGetLastError();
}

} // namespace sdpk
