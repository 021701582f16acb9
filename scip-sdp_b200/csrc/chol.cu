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

template <int NBL>
__global__ void __launch_bounds__(NBL * 4)
leaf_kernel(int mode, int nb, double* __restrict__ A, int lda, double* __restrict__ Linv, int ldi,
   double* __restrict__ diaginv, int* __restrict__ info, int pivot_offset, long long* __restrict__ dbg)
{
   long long tc0 = clock64(), tc1 = 0, tc2 = 0, tc3 = 0;
   constexpr int NTILE = NBL / 8, LD = NBL + 4, NSTEP = NBL / 4, NTHREADS = NBL * 4, NWARP = NBL / 8;
   extern __shared__ __align__(16) double lsm[];
   __shared__ int sbad;
   double* G = lsm;                              // (NBL + 1) x LD
   double* Tr = G + (NBL + 1) * LD;              // (NBL / 2) x LD
   double* Pcol = Tr + (NBL / 2) * LD;           // NBL x 4  current block column
   double* P = Pcol + NBL * 4;                   // NBL x 4  panel
   double* rdg = P + NBL * 4;                    // NBL      reciprocals of the diagonal of L
   const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
   const int fr = lane >> 2, fc = lane & 3;

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
      P[tid] = 0.0;                                      // NTHREADS = 4 NBL
      if( tid == 0 ) sbad = 0x7fffffff;                  // smallest index of a non-positive pivot
      __syncthreads();
      tc1 = clock64();

#pragma unroll
      for( int t = 0; t < NSTEP; ++t )
      {
         const int jt = t >> 1, half = t & 1;
         // (1) block column t -> shared memory (rows of the strips w >= jt)
         if( w >= jt && (fc >> 1) == half )
            *reinterpret_cast<double2*>(Pcol + (8 * w + fr) * 4 + 2 * (fc & 1)) = make_double2(c[jt][0], c[jt][1]);
         __syncthreads();
         // (2) panel row of every row r >= 4t
         if( tid < NBL && tid >= 4 * t )
         {
            const int r = tid;
            const double4* Dg = reinterpret_cast<const double4*>(Pcol + 16 * t);
            const double4 g0 = Dg[0], g1 = Dg[1], g2 = Dg[2], g3 = Dg[3], av = *reinterpret_cast<const double4*>(Pcol + 4 * r);
            // 4 x 4 Cholesky, branch free (a non-positive pivot is replaced by 1 and reported after the loop)
            double s0 = g0.x;
            const bool b0 = !(s0 > 0.0); s0 = b0 ? 1.0 : s0;
            const double r0 = fast_rsqrt64(s0);
            const double l10 = g1.x * r0, l20 = g2.x * r0, l30 = g3.x * r0;
            double s1 = g1.y - l10 * l10;
            const bool b1 = !(s1 > 0.0); s1 = b1 ? 1.0 : s1;
            const double r1 = fast_rsqrt64(s1);
            const double l21 = (g2.y - l20 * l10) * r1, l31 = (g3.y - l30 * l10) * r1;
            double s2 = g2.z - l20 * l20 - l21 * l21;
            const bool b2 = !(s2 > 0.0); s2 = b2 ? 1.0 : s2;
            const double r2 = fast_rsqrt64(s2);
            const double l32 = (g3.z - l30 * l20 - l31 * l21) * r2;
            double s3 = g3.w - l30 * l30 - l31 * l31 - l32 * l32;
            const bool b3 = !(s3 > 0.0); s3 = b3 ? 1.0 : s3;
            const double r3 = fast_rsqrt64(s3);
            if( r == 4 * t && (b0 || b1 || b2 || b3) )
            {
               const int q = b0 ? 0 : (b1 ? 1 : (b2 ? 2 : 3));
               if( 4 * t + q < nb ) atomicMin(&sbad, 4 * t + q);
            }
            double4 pr;
            if( r >= 4 * t + 4 )
            {
               pr.x = av.x * r0;
               pr.y = (av.y - pr.x * l10) * r1;
               pr.z = (av.z - pr.x * l20 - pr.y * l21) * r2;
               pr.w = (av.w - pr.x * l30 - pr.y * l31 - pr.z * l32) * r3;
               *reinterpret_cast<double4*>(P + 4 * r) = pr;
            }
            else
            {
               const int q = r - 4 * t;
               const double d0 = s0 * r0, d1 = s1 * r1, d2 = s2 * r2, d3 = s3 * r3;
               pr.x = (q == 0) ? d0 : (q == 1 ? l10 : (q == 2 ? l20 : l30));
               pr.y = (q == 0) ? 0.0 : (q == 1 ? d1 : (q == 2 ? l21 : l31));
               pr.z = (q <= 1) ? 0.0 : (q == 2 ? d2 : l32);
               pr.w = (q <= 2) ? 0.0 : d3;
               *reinterpret_cast<double4*>(P + 4 * r) = make_double4(0.0, 0.0, 0.0, 0.0);
               rdg[r] = (q == 0) ? r0 : (q == 1 ? r1 : (q == 2 ? r2 : r3));
            }
            // row-major copy of L (zeros above the diagonal land in the triangle that the inverse fills later)
            *reinterpret_cast<double4*>(G + (r + 1) * LD + 4 * t) = pr;
         }
         __syncthreads();
         // (3) rank-4 update of the tiles that still change: rows >= 4t+4, columns >= 4t+4, lower tiles
         {
            const int jlo = (t + 1) >> 1;
            if( w >= jlo )
            {
               const double a = -P[(8 * w + fr) * 4 + fc];
#pragma unroll
               for( int j = 0; j < NTILE; ++j )
               {
                  if( j >= jlo && j <= w )
                  {
                     const double b = P[(8 * j + fr) * 4 + fc];
                     dmma884(c[j][0], c[j][1], a, b);
                  }
               }
            }
         }
      }
      __syncthreads();
      if( tid == 0 && sbad != 0x7fffffff ) atomicCAS(info, 0, pivot_offset + sbad + 1);
      tc2 = clock64();
      // L -> global memory, lower triangle, column by column (coalesced)
      for( int e = tid; e < nb * nb; e += NTHREADS )
      {
         const int i = e % nb, j = e / nb;
         if( i >= j ) A[(size_t)j * lda + i] = G[(i + 1) * LD + j];
      }
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

template <int NBL> constexpr size_t leaf_smem() { return sizeof(double) * ((size_t)(NBL + 1) * (NBL + 4) + (size_t)(NBL / 2) * (NBL + 4) + 9 * (size_t)NBL); }

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

} // namespace

cudaError_t potrf_lower(cudaStream_t st, int n, double* A, int lda, double* Linv, int ldi, double* diaginv, double* work, int ldw, int* d_info)
{
   if( n <= 0 ) return cudaSuccess;
   if( Linv )
      SDPK_CUDA_CHECK( cudaMemsetAsync(Linv, 0, sizeof(double) * (size_t)ldi * n, st) );
   return chol_rec(st, n, A, lda, Linv, ldi, diaginv, work, ldw, d_info, 0);
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
