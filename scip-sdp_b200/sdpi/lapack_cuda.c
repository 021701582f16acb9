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

/** minimum-norm least-squares solution of A x = b (A is m x n, column-major) through the normal equations and the
 *  eigen-decomposition of A'A on the device: x = V diag(1/lambda_i, lambda_i > eps) V' A' b  (pseudo-inverse, like DGELSD) */
SCIP_RETCODE SCIPlapackLinearSolve(BMS_BUFMEM* bufmem, int m, int n, SCIP_Real* A, SCIP_Real* b, SCIP_Real* x)
{
   SCIP_Real* G;
   SCIP_Real* w;
   SCIP_Real* V;
   SCIP_Real* atb;
   SCIP_RETCODE rc;
   SCIP_Real wmax = 0.0;
   int i;
   int j;
   int l;

   assert( bufmem != NULL && A != NULL && b != NULL && x != NULL );
   MEM_CALL( BMSallocBufferMemoryArray(bufmem, &G, n * n) );
   MEM_CALL( BMSallocBufferMemoryArray(bufmem, &V, n * n) );
   MEM_CALL( BMSallocBufferMemoryArray(bufmem, &w, n) );
   MEM_CALL( BMSallocBufferMemoryArray(bufmem, &atb, n) );
   for( i = 0; i < n; ++i )
   {
      SCIP_Real s = 0.0;
      for( l = 0; l < m; ++l )
         s += A[(size_t)i * m + l] * b[l];
      atb[i] = s;
      for( j = 0; j <= i; ++j )
      {
         s = 0.0;
         for( l = 0; l < m; ++l )
            s += A[(size_t)i * m + l] * A[(size_t)j * m + l];
         G[(size_t)i * n + j] = s;
         G[(size_t)j * n + i] = s;
      }
   }
   rc = deviceEigen(n, G, w, V);
   if( rc == SCIP_OKAY )
   {
      for( i = 0; i < n; ++i )
      {
         x[i] = 0.0;
         wmax = MAX(wmax, REALABS(w[i]));
      }
      for( l = 0; l < n; ++l )
      {
         SCIP_Real coef = 0.0;
         if( w[l] <= 1e-13 * wmax )
            continue;
         for( i = 0; i < n; ++i )
            coef += V[(size_t)l * n + i] * atb[i];
         coef /= w[l];
         for( i = 0; i < n; ++i )
            x[i] += coef * V[(size_t)l * n + i];
      }
   }
   BMSfreeBufferMemoryArray(bufmem, &atb);
   BMSfreeBufferMemoryArray(bufmem, &w);
   BMSfreeBufferMemoryArray(bufmem, &V);
   BMSfreeBufferMemoryArray(bufmem, &G);
   return rc;
}
