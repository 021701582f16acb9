/* lapack_cuda.c — GPU implementation of the eigenvalue entry points of src/sdpi/lapack_interface.h.
 *
 * Replaces the DSYEVR/DSYEVX calls of the reference's lapack_interface.c:178-603 by the batched Jacobi kernel of
 * libsdpcuda (sdpcuda_syev_batched): SCIPlapackComputeIthEigenvalue (cons_sdp.c:1687 separateSol, :701 feasibility check,
 * sdpsolchecker.c:247), SCIPlapackComputeEigenvectorsNegative (cons_sdp.c:1699, the default cut path) and
 * SCIPlapackComputeEigenvectorDecomposition (cons_sdp.c:7956, relax_sdp.c:1877,2735,3418).  Conventions kept:
 * eigenvalues ascending, i is 1-based, eigenvector k is row k of the output array, interval (-1e20, -tol].
 * The input matrix is NOT destroyed (the reference documents "will be destroyed"; callers copy before the call).
 * The three BLAS-level helpers (matrix-vector, matrix-matrix, linear solve) are not eigen kernels and operate on the
 * tiny host arrays of cons_sdp.c directly.
 */
#include <assert.h>
#include <math.h>
#include <string.h>

#include "sdpi/lapack_interface.h"
#include "blockmemshell/memory.h"
#include "scip/def.h"
#include "scip/pub_message.h"
#include "devctx.h"

#define MEM_CALL(x) do { if( NULL == (x) ) { SCIPerrorMessage("No memory in function call.\n"); return SCIP_NOMEMORY; } } while( FALSE )

static _Thread_local sdpcuda_handle* threadhandle = NULL;

sdpcuda_handle* sdpiCudaThreadHandle(void)
{
   if( threadhandle == NULL )
   {
      if( sdpcuda_create(&threadhandle, -1) != SDPCUDA_OK )
         threadhandle = NULL;
   }
   return threadhandle;
}

/** full decomposition on the device; w[n] ascending, V[n*n] with eigenvector k in row k (or V == NULL) */
static SCIP_RETCODE deviceEigen(int n, const SCIP_Real* A, SCIP_Real* w, SCIP_Real* V)
{
   sdpcuda_handle* h = sdpiCudaThreadHandle();
   if( h == NULL )
   {
      SCIPerrorMessage("no CUDA device handle available for the eigenvalue computation (there is no CPU fallback).\n");
      return SCIP_ERROR;
   }
   if( sdpcuda_syev_batched(h, n, 1, A, w, V) != SDPCUDA_OK )
   {
      SCIPerrorMessage("sdpcuda_syev_batched failed.\n");
      return SCIP_ERROR;
   }
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPlapackComputeIthEigenvalue(BMS_BUFMEM* bufmem, SCIP_Bool geteigenvectors, int n, SCIP_Real* A, int i,
   SCIP_Real* eigenvalue, SCIP_Real* eigenvector)
{
   SCIP_Real* w;
   SCIP_Real* V = NULL;
   SCIP_RETCODE rc;

   assert( bufmem != NULL );
   assert( n > 0 && i >= 1 && i <= n );
   assert( A != NULL && eigenvalue != NULL );
   assert( !geteigenvectors || eigenvector != NULL );

   MEM_CALL( BMSallocBufferMemoryArray(bufmem, &w, n) );
   if( geteigenvectors )
      MEM_CALL( BMSallocBufferMemoryArray(bufmem, &V, n * n) );
   rc = deviceEigen(n, A, w, V);
   if( rc == SCIP_OKAY )
   {
      *eigenvalue = w[i - 1];
      if( geteigenvectors )
         memcpy(eigenvector, V + (size_t)(i - 1) * n, sizeof(SCIP_Real) * (size_t)n);
   }
   BMSfreeBufferMemoryArrayNull(bufmem, &V);
   BMSfreeBufferMemoryArray(bufmem, &w);
   return rc;
}

SCIP_RETCODE SCIPlapackComputeIthEigenvalueAlternative(BMS_BUFMEM* bufmem, SCIP_Bool geteigenvectors, int n, SCIP_Real* A, int i,
   SCIP_Real* eigenvalue, SCIP_Real* eigenvector)
{
   return SCIPlapackComputeIthEigenvalue(bufmem, geteigenvectors, n, A, i, eigenvalue, eigenvector);
}

SCIP_RETCODE SCIPlapackComputeEigenvectorsNegative(BMS_BUFMEM* bufmem, int n, SCIP_Real* A, SCIP_Real tol, int* neigenvalues,
   SCIP_Real* eigenvalues, SCIP_Real* eigenvectors)
{
   SCIP_Real* w;
   SCIP_Real* V;
   SCIP_RETCODE rc;
   int k;

   assert( bufmem != NULL );
   assert( n > 0 );
   assert( A != NULL && neigenvalues != NULL && eigenvalues != NULL && eigenvectors != NULL );

   MEM_CALL( BMSallocBufferMemoryArray(bufmem, &w, n) );
   MEM_CALL( BMSallocBufferMemoryArray(bufmem, &V, n * n) );
   rc = deviceEigen(n, A, w, V);
   *neigenvalues = 0;
   if( rc == SCIP_OKAY )
   {
      /* eigenvalues in (-1e20, -tol], ascending, with their eigenvectors as rows */
      for( k = 0; k < n && w[k] <= -tol && w[k] > -1e20; ++k )
      {
         eigenvalues[k] = w[k];
         memcpy(eigenvectors + (size_t)k * n, V + (size_t)k * n, sizeof(SCIP_Real) * (size_t)n);
      }
      *neigenvalues = k;
   }
   BMSfreeBufferMemoryArray(bufmem, &V);
   BMSfreeBufferMemoryArray(bufmem, &w);
   return rc;
}

SCIP_RETCODE SCIPlapackComputeEigenvectorDecomposition(BMS_BUFMEM* bufmem, int n, SCIP_Real* A, SCIP_Real* eigenvalues,
   SCIP_Real* eigenvectors)
{
   (void) bufmem;
   assert( n > 0 );
   assert( A != NULL && eigenvalues != NULL && eigenvectors != NULL );
   return deviceEigen(n, A, eigenvalues, eigenvectors);
}

/** result = matrix * vector, matrix given row-wise (nrows x ncols) as in lapack_interface.c:607-650 */
SCIP_RETCODE SCIPlapackMatrixVectorMult(int nrows, int ncols, SCIP_Real* matrix, SCIP_Real* vector, SCIP_Real* result)
{
   int r;
   int c;
   for( r = 0; r < nrows; ++r )
   {
      SCIP_Real s = 0.0;
      for( c = 0; c < ncols; ++c )
         s += matrix[(size_t)r * ncols + c] * vector[c];
      result[r] = s;
   }
   return SCIP_OKAY;
}

/** result = op(A) * op(B) with all arrays column-major like the DGEMM call of lapack_interface.c:654-708
 *  (known answer of unittests/src/checklapack.c:83-91: A = [1 2; 3 4]', B = [5 6; 7 8]', A * B' -> 26 38 30 44) */
SCIP_RETCODE SCIPlapackMatrixMatrixMult(int nrowsA, int ncolsA, SCIP_Real* matrixA, SCIP_Bool transposeA, int nrowsB, int ncolsB,
   SCIP_Real* matrixB, SCIP_Bool transposeB, SCIP_Real* result)
{
   const int m = transposeA ? ncolsA : nrowsA;
   const int k = transposeA ? nrowsA : ncolsA;
   const int n = transposeB ? nrowsB : ncolsB;
   int i;
   int j;
   int l;

   assert( (transposeB ? ncolsB : nrowsB) == k );
   for( j = 0; j < n; ++j )
   {
      for( i = 0; i < m; ++i )
      {
         SCIP_Real s = 0.0;
         for( l = 0; l < k; ++l )
         {
            SCIP_Real a = transposeA ? matrixA[(size_t)i * nrowsA + l] : matrixA[(size_t)l * nrowsA + i];
            SCIP_Real b = transposeB ? matrixB[(size_t)l * nrowsB + j] : matrixB[(size_t)j * nrowsB + l];
            s += a * b;
         }
         result[(size_t)j * m + i] = s;
      }
   }
   return SCIP_OKAY;
}

/** minimum-norm least-squares solution of A x = b (A is m x n, column-major), the contract of the DGELSD call of
 *  lapack_interface.c:712-822.  One-sided Jacobi (Hestenes) SVD on the host: the columns of a copy U of A are rotated until they
 *  are mutually orthogonal, A V = U = Q diag(sigma), so x = V diag(1/sigma_i, sigma_i > rcond sigma_max) Q' b.  Unlike the normal
 *  equations this works at cond(A), not cond(A)^2; the arrays are the tiny host matrices of cons_sdp.c, so nothing goes to the device. */
SCIP_RETCODE SCIPlapackLinearSolve(BMS_BUFMEM* bufmem, int m, int n, SCIP_Real* A, SCIP_Real* b, SCIP_Real* x)
{
   SCIP_Real* U;
   SCIP_Real* V;
   SCIP_Real* sig;
   SCIP_Real smax = 0.0;
   const SCIP_Real rcond = 1e-13;
   int sweep;
   int p;
   int q;
   int i;

   assert( bufmem != NULL && A != NULL && b != NULL && x != NULL );
   MEM_CALL( BMSallocBufferMemoryArray(bufmem, &U, m * n) );
   MEM_CALL( BMSallocBufferMemoryArray(bufmem, &V, n * n) );
   MEM_CALL( BMSallocBufferMemoryArray(bufmem, &sig, n) );
   memcpy(U, A, sizeof(SCIP_Real) * (size_t)m * n);
   for( i = 0; i < n * n; ++i )
      V[i] = 0.0;
   for( i = 0; i < n; ++i )
      V[(size_t)i * n + i] = 1.0;
   for( sweep = 0; sweep < 60; ++sweep )
   {
      SCIP_Bool rotated = FALSE;
      for( p = 0; p < n - 1; ++p )
      {
         for( q = p + 1; q < n; ++q )
         {
            SCIP_Real alpha = 0.0;
            SCIP_Real beta = 0.0;
            SCIP_Real gamma = 0.0;
            SCIP_Real zeta;
            SCIP_Real t;
            SCIP_Real c;
            SCIP_Real sn;

            for( i = 0; i < m; ++i )
            {
               alpha += U[(size_t)p * m + i] * U[(size_t)p * m + i];
               beta += U[(size_t)q * m + i] * U[(size_t)q * m + i];
               gamma += U[(size_t)p * m + i] * U[(size_t)q * m + i];
            }
            if( REALABS(gamma) <= 1e-16 * sqrt(alpha * beta) || gamma == 0.0 )
               continue;
            rotated = TRUE;
            zeta = (beta - alpha) / (2.0 * gamma);
            t = (zeta >= 0.0 ? 1.0 : -1.0) / (REALABS(zeta) + sqrt(1.0 + zeta * zeta));
            c = 1.0 / sqrt(1.0 + t * t);
            sn = c * t;
            for( i = 0; i < m; ++i )
            {
               const SCIP_Real up = U[(size_t)p * m + i];
               const SCIP_Real uq = U[(size_t)q * m + i];
               U[(size_t)p * m + i] = c * up - sn * uq;
               U[(size_t)q * m + i] = sn * up + c * uq;
            }
            for( i = 0; i < n; ++i )
            {
               const SCIP_Real vp = V[(size_t)p * n + i];
               const SCIP_Real vq = V[(size_t)q * n + i];
               V[(size_t)p * n + i] = c * vp - sn * vq;
               V[(size_t)q * n + i] = sn * vp + c * vq;
            }
         }
      }
      if( !rotated )
         break;
   }
   for( p = 0; p < n; ++p )
   {
      SCIP_Real s2 = 0.0;
      for( i = 0; i < m; ++i )
         s2 += U[(size_t)p * m + i] * U[(size_t)p * m + i];
      sig[p] = sqrt(s2);
      smax = MAX(smax, sig[p]);
   }
   for( i = 0; i < n; ++i )
      x[i] = 0.0;
   for( p = 0; p < n; ++p )
   {
      SCIP_Real coef = 0.0;
      if( sig[p] <= rcond * smax )
         continue;
      for( i = 0; i < m; ++i )
         coef += U[(size_t)p * m + i] * b[i];
      coef /= sig[p] * sig[p];                           /* (u_p / sigma_p)' b / sigma_p */
      for( i = 0; i < n; ++i )
         x[i] += coef * V[(size_t)p * n + i];
   }
   BMSfreeBufferMemoryArray(bufmem, &sig);
   BMSfreeBufferMemoryArray(bufmem, &V);
   BMSfreeBufferMemoryArray(bufmem, &U);
   return SCIP_OKAY;
}

/* ---- one separation round in one device call -------------------------------------------------------------------------------
 * cons_sdp.c:1612-1797 (separateSol) is called once per SDP constraint and asks for the negative eigenpairs of that constraint's
 * matrix; the matrices of all constraints of a round are known before the first cut is formed, so a caller that collects them
 * (INTEGRATION.md shows the two-phase loop) gets them decomposed together: matrices of equal order share ONE host->device copy,
 * ONE launch of the batched Jacobi kernel and ONE copy back.  Same conventions as SCIPlapackComputeEigenvectorsNegative for every
 * matrix k: neigenvalues[k], eigenvalues[k][0..], eigenvector i of matrix k in row i of eigenvectors[k].  Not part of
 * lapack_interface.h. */
SCIP_RETCODE SCIPlapackComputeEigenvectorsNegativeBatch(BMS_BUFMEM* bufmem, int nmatrices, const int* sizes, SCIP_Real* const* matrices,
   SCIP_Real tol, int* neigenvalues, SCIP_Real* const* eigenvalues, SCIP_Real* const* eigenvectors)
{
   sdpcuda_handle* h = sdpiCudaThreadHandle();
   SCIP_Bool* done;
   SCIP_RETCODE rc = SCIP_OKAY;
   int k;

   assert( bufmem != NULL && nmatrices >= 0 );
   if( nmatrices == 0 )
      return SCIP_OKAY;
   assert( sizes != NULL && matrices != NULL && neigenvalues != NULL && eigenvalues != NULL && eigenvectors != NULL );
   if( h == NULL )
   {
      SCIPerrorMessage("no CUDA device handle available for the eigenvalue computation (there is no CPU fallback).\n");
      return SCIP_ERROR;
   }
   MEM_CALL( BMSallocClearBufferMemoryArray(bufmem, &done, nmatrices) );
   for( k = 0; k < nmatrices && rc == SCIP_OKAY; ++k )
   {
      SCIP_Real* Apack;
      SCIP_Real* w;
      SCIP_Real* V;
      const int n = sizes[k];
      int cnt = 0;
      int pos = 0;
      int l;

      if( done[k] )
         continue;
      for( l = k; l < nmatrices; ++l )
         cnt += (sizes[l] == n && !done[l]) ? 1 : 0;
      if( NULL == BMSallocBufferMemoryArray(bufmem, &Apack, (size_t)cnt * n * n) ) { rc = SCIP_NOMEMORY; break; }
      if( NULL == BMSallocBufferMemoryArray(bufmem, &w, (size_t)cnt * n) ) { BMSfreeBufferMemoryArray(bufmem, &Apack); rc = SCIP_NOMEMORY; break; }
      if( NULL == BMSallocBufferMemoryArray(bufmem, &V, (size_t)cnt * n * n) )
      {
         BMSfreeBufferMemoryArray(bufmem, &w); BMSfreeBufferMemoryArray(bufmem, &Apack); rc = SCIP_NOMEMORY; break;
      }
      for( l = k; l < nmatrices; ++l )
      {
         if( sizes[l] == n && !done[l] )
            memcpy(Apack + (size_t)(pos++) * n * n, matrices[l], sizeof(SCIP_Real) * (size_t)n * n);
      }
      if( sdpcuda_syev_batched(h, n, cnt, Apack, w, V) != SDPCUDA_OK )
      {
         SCIPerrorMessage("sdpcuda_syev_batched failed.\n");
         rc = SCIP_ERROR;
      }
      pos = 0;
      for( l = k; l < nmatrices && rc == SCIP_OKAY; ++l )
      {
         const SCIP_Real* wl;
         const SCIP_Real* Vl;
         int i;

         if( sizes[l] != n || done[l] )
            continue;
         wl = w + (size_t)pos * n;
         Vl = V + (size_t)pos * n * n;
         for( i = 0; i < n && wl[i] <= -tol && wl[i] > -1e20; ++i )
         {
            eigenvalues[l][i] = wl[i];
            memcpy(eigenvectors[l] + (size_t)i * n, Vl + (size_t)i * n, sizeof(SCIP_Real) * (size_t)n);
         }
         neigenvalues[l] = i;
         done[l] = TRUE;
         ++pos;
      }
      BMSfreeBufferMemoryArray(bufmem, &V);
      BMSfreeBufferMemoryArray(bufmem, &w);
      BMSfreeBufferMemoryArray(bufmem, &Apack);
   }
   BMSfreeBufferMemoryArray(bufmem, &done);
   return rc;
}
