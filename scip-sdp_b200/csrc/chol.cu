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

namespace sdpk {
namespace {

constexpr int NB = CHOL_NB;
constexpr int LDSM = NB + 1;

// fast reciprocal square root in double precision: float seed + two Newton steps (relative error ~1e-16)
__device__ __forceinline__ double fast_rsqrt(double x)
{
   double y = (double)rsqrtf((float)x);
   y = y * (1.5 - 0.5 * x * y * y);
   y = y * (1.5 - 0.5 * x * y * y);
   return y;
}

// mode 0: factorise A (nb x nb, lower) in place, optionally write inverse of L to Linv (upper part zeroed) / diaginv
// mode 1: A holds a lower-triangular factor already; only invert it
// One CTA of 128 threads; thread i < 64 owns row i (Cholesky-Crout: column k needs one dot product per row and two
// barriers), then thread j < 64 owns column j of the inverse (forward substitution with broadcast reads of L).
__global__ void __launch_bounds__(128)
diag_block_kernel(int mode, int nb, double* __restrict__ A, int lda, double* __restrict__ Linv, int ldi,
   double* __restrict__ diaginv, int* __restrict__ info, int pivot_offset)
{
   extern __shared__ __align__(16) double diag_smem[];
   double* L = diag_smem;
   double* W = diag_smem + NB * LDSM;
   double* dinv = W + NB * LDSM;              // reciprocals of the diagonal of L
   const int tid = threadIdx.x;

   for( int e = tid; e < NB * NB; e += blockDim.x )
   {
      int i = e % NB, j = e / NB;
      double v = 0.0;
      if( i < nb && j < nb && i >= j ) v = A[(size_t)j * lda + i];
      else if( i == j ) v = 1.0;                      // identity padding keeps the padded block positive definite
      L[i * LDSM + j] = v;
   }
   __syncthreads();

   if( mode == 0 )
   {
      const int i = tid;
      for( int k = 0; k < NB; ++k )
      {
         double s = 0.0;
         if( i < NB && i >= k )
         {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            const double* ri = L + i * LDSM;
            const double* rk = L + k * LDSM;
            int j = 0;
            for( ; j + 4 <= k; j += 4 )
            {
               s0 += ri[j] * rk[j]; s1 += ri[j + 1] * rk[j + 1]; s2 += ri[j + 2] * rk[j + 2]; s3 += ri[j + 3] * rk[j + 3];
            }
            for( ; j < k; ++j ) s0 += ri[j] * rk[j];
            s = ri[k] - ((s0 + s1) + (s2 + s3));
            if( i == k )
            {
               if( !(s > 0.0) )
               {
                  if( k < nb ) atomicCAS(info, 0, pivot_offset + k + 1);
                  s = 1.0;
               }
               double r = fast_rsqrt(s);
               L[k * LDSM + k] = s * r;
               dinv[k] = r;
            }
         }
         __syncthreads();
         if( i < NB && i > k )
            L[i * LDSM + k] = s * dinv[k];
         __syncthreads();
      }
      for( int e = tid; e < nb * nb; e += blockDim.x )
      {
         int r = e % nb, c = e / nb;
         if( r >= c ) A[(size_t)c * lda + r] = L[r * LDSM + c];
      }
   }
   else
   {
      if( tid < NB ) dinv[tid] = 1.0 / L[tid * LDSM + tid];
      __syncthreads();
   }

   if( Linv != nullptr || diaginv != nullptr )
   {
      // thread j: column j of W = L^-1; the k-loop starts at 0 for all threads (W is zero above the diagonal) so that the
      // reads of L[i][k] are warp-wide broadcasts and the reads of W[k][j] are conflict free
      const int j = tid;
      if( j < NB )
      {
         for( int i = 0; i < NB; ++i ) W[i * LDSM + j] = (i == j) ? dinv[j] : 0.0;
         for( int i = 1; i < NB; ++i )
         {
            const double* ri = L + i * LDSM;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int k = 0;
            for( ; k + 4 <= i; k += 4 )
            {
               s0 += ri[k] * W[k * LDSM + j]; s1 += ri[k + 1] * W[(k + 1) * LDSM + j];
               s2 += ri[k + 2] * W[(k + 2) * LDSM + j]; s3 += ri[k + 3] * W[(k + 3) * LDSM + j];
            }
            for( ; k < i; ++k ) s0 += ri[k] * W[k * LDSM + j];
            if( i > j ) W[i * LDSM + j] = -((s0 + s1) + (s2 + s3)) * dinv[i];
         }
      }
      __syncthreads();
      if( Linv != nullptr )
         for( int e = tid; e < nb * nb; e += blockDim.x )
         {
            int r = e % nb, c = e / nb;
            Linv[(size_t)c * ldi + r] = W[r * LDSM + c];
         }
      if( diaginv != nullptr )
         for( int e = tid; e < NB * NB; e += blockDim.x )
         {
            int r = e % NB, c = e / NB;
            diaginv[(size_t)c * NB + r] = (r < nb && c < nb) ? W[r * LDSM + c] : 0.0;
         }
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

constexpr size_t DIAG_SMEM = (2 * NB * LDSM + NB) * sizeof(double);

cudaError_t launch_diag(cudaStream_t st, int mode, int nb, double* A, int lda, double* Linv, int ldi, double* diaginv, int* info, int off)
{
   static bool configured = false;
   if( !configured )
   {
      SDPK_CUDA_CHECK( cudaFuncSetAttribute(diag_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM) );
      configured = true;
   }
   ProfScope prof(st, PROF_DIAG, (mode == 0 ? 1.0 : 0.0) * nb * (double)nb * nb / 3.0 + ((Linv || diaginv) ? nb * (double)nb * nb / 3.0 : 0.0));
   diag_block_kernel<<<1, 128, DIAG_SMEM, st>>>(mode, nb, A, lda, Linv, ldi, diaginv, info, off);
   count_launch();
   return cudaGetLastError();
}

int split_point(int n)
{
   int n1 = round_up(n / 2, NB);
   if( n1 >= n ) n1 -= NB;
   return n1;
}

cudaError_t chol_rec(cudaStream_t st, int n, double* A, int lda, double* Linv, int ldi, double* diaginv, double* work,
   int ldw, int* d_info, int off)
{
   if( n <= NB )
   {
      return launch_diag(st, 0, n, A, lda, Linv, ldi, diaginv ? diaginv + (size_t)(off / NB) * NB * NB : nullptr, d_info, off);
   }
   const int n1 = split_point(n), n2 = n - n1;
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
   SDPK_CUDA_CHECK( gemm(st, false, true, n2, n1, n1, 1.0, A21, lda, 0, Li11, ldi11, 0, 0.0, wrk, ldw, 0, 1, 0) );
   SDPK_CUDA_CHECK( copy2d(st, n2, n1, wrk, ldw, A21, lda) );
   // A22 -= L21 L21'
   SDPK_CUDA_CHECK( gemm(st, false, true, n2, n2, n1, -1.0, A21, lda, 0, A21, lda, 0, 1.0, A22, lda, 0, 1, GEMM_LOWER) );
   if( Linv )
   {
      double* Li22 = Linv + (size_t)n1 * ldi + n1;
      double* Li21 = Linv + n1;
      SDPK_CUDA_CHECK( chol_rec(st, n2, A22, lda, Li22, ldi, diaginv, work, ldw, d_info, off + n1) );
      // Linv21 = -Linv22 * (L21 * Linv11)
      SDPK_CUDA_CHECK( gemm(st, false, false, n2, n1, n1, 1.0, A21, lda, 0, Linv, ldi, 0, 0.0, work, ldw, 0, 1, 0) );
      SDPK_CUDA_CHECK( gemm(st, false, false, n2, n1, n2, -1.0, Li22, ldi, 0, work, ldw, 0, 0.0, Li21, ldi, 0, 1, 0) );
   }
   else
   {
      SDPK_CUDA_CHECK( chol_rec(st, n2, A22, lda, nullptr, 0, diaginv, work, ldw, d_info, off + n1) );
   }
   return cudaSuccess;
}

cudaError_t trtri_rec(cudaStream_t st, int n, const double* L, int ldl, double* Linv, int ldi, double* work, int ldw)
{
   if( n <= NB )
   {
      return launch_diag(st, 1, n, const_cast<double*>(L), ldl, Linv, ldi, nullptr, nullptr, 0);
   }
   const int n1 = split_point(n), n2 = n - n1;
   const double* L21 = L + n1;
   double* Li22 = Linv + (size_t)n1 * ldi + n1;
   SDPK_CUDA_CHECK( trtri_rec(st, n1, L, ldl, Linv, ldi, work, ldw) );
   SDPK_CUDA_CHECK( trtri_rec(st, n2, L + (size_t)n1 * ldl + n1, ldl, Li22, ldi, work, ldw) );
   SDPK_CUDA_CHECK( gemm(st, false, false, n2, n1, n1, 1.0, L21, ldl, 0, Linv, ldi, 0, 0.0, work, ldw, 0, 1, 0) );
   SDPK_CUDA_CHECK( gemm(st, false, false, n2, n1, n2, -1.0, Li22, ldi, 0, work, ldw, 0, 0.0, Linv + n1, ldi, 0, 1, 0) );
   return cudaSuccess;
}

// ---- blocked triangular solves with one right-hand side, one launch per diagonal block ------------------------------
// forward step for block b:  x_b = Linv_bb * rhs_b (every CTA redundantly), then rhs_r -= L[r, b-block] x_b for r below
__global__ void __launch_bounds__(256)
fwd_step_kernel(int n, int b, const double* __restrict__ L, int ldl, const double* __restrict__ diaginv, double* __restrict__ x,
   double* __restrict__ out)
{
   __shared__ double xb[NB];
   const int r0 = b * NB, nb = min(NB, n - r0);
   const double* Di = diaginv + (size_t)b * NB * NB;
   const int tid = threadIdx.x;
   if( tid < NB )
   {
      double s = 0.0;
      if( tid < nb )
         for( int k = 0; k <= tid; ++k ) s += Di[(size_t)k * NB + tid] * x[r0 + k];
      xb[tid] = s;
   }
   __syncthreads();
   int row = r0 + nb + blockIdx.x * blockDim.x + tid;
   if( row < n )
   {
      double s = 0.0;
      for( int k = 0; k < nb; ++k ) s += L[(size_t)(r0 + k) * ldl + row] * xb[k];
      x[row] -= s;
   }
   if( blockIdx.x == 0 && tid < nb )
      out[r0 + tid] = xb[tid];      // separate output: other CTAs may still be reading the right-hand side block
}

// backward step for block b:  x_b = Linv_bb' * rhs_b, then rhs_c -= L[b-block, c]' x_b for columns c left of the block
__global__ void __launch_bounds__(256)
bwd_step_kernel(int n, int b, const double* __restrict__ L, int ldl, const double* __restrict__ diaginv, double* __restrict__ x,
   double* __restrict__ out)
{
   __shared__ double xb[NB];
   const int r0 = b * NB, nb = min(NB, n - r0);
   const double* Di = diaginv + (size_t)b * NB * NB;
   const int tid = threadIdx.x;
   if( tid < NB )
   {
      double s = 0.0;
      if( tid < nb )
         for( int k = tid; k < nb; ++k ) s += Di[(size_t)tid * NB + k] * x[r0 + k];
      xb[tid] = s;
   }
   __syncthreads();
   // each warp handles columns c = blockIdx.x*8 + warp, ... ; column c of L is contiguous over the block rows
   const int lane = tid & 31, warp = tid >> 5;
   int c = blockIdx.x * 8 + warp;
   if( c < r0 )
   {
      double s = 0.0;
      for( int k = lane; k < nb; k += 32 ) s += L[(size_t)c * ldl + r0 + k] * xb[k];
#pragma unroll
      for( int o = 16; o > 0; o >>= 1 ) s += __shfl_xor_sync(0xffffffffu, s, o);
      if( lane == 0 ) x[c] -= s;
   }
   if( blockIdx.x == 0 && tid < nb )
      out[r0 + tid] = xb[tid];
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

cudaError_t potrs_vec(cudaStream_t st, int n, const double* L, int ldl, const double* diaginv, double* b, double* tmp)
{
   const int nblk = ceil_div(n, NB);
   ProfScope prof(st, PROF_TRSV, 8.0 * n * (double)n);      // reads the triangle of L twice
   for( int k = 0; k < nblk; ++k )
   {
      int below = n - min(n, (k + 1) * NB);
      fwd_step_kernel<<<max(1, ceil_div(below, 256)), 256, 0, st>>>(n, k, L, ldl, diaginv, b, tmp);
      count_launch();
   }
   for( int k = nblk - 1; k >= 0; --k )
   {
      bwd_step_kernel<<<max(1, ceil_div(k * NB, 8)), 256, 0, st>>>(n, k, L, ldl, diaginv, tmp, b);
      count_launch();
   }
   return cudaGetLastError();
}

} // namespace sdpk
