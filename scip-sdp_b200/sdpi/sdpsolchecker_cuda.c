/* sdpsolchecker_cuda.c — SCIPsdpSolcheckerCheck (src/sdpi/sdpsolchecker.h) with the SDP-block test on the device.
 *
 * Same contract as the reference's sdpsolchecker.c:58-270: a solution vector is infeasible if a variable bound or an LP
 * row is violated by more than feastol, or if for a kept block lambda_min(sum_j A_j y_j - A_0) < -feastol.  Bounds and
 * rows are O(nnz) host work; the eigenvalue test is done as a device Cholesky of Z(y) + feastol*I (sdpcuda_psd_check),
 * which scales to the 2000 x 2000 blocks of the max-cut configuration where a host DSYEVR per solve would dominate.
 * Linked into libsdpisolver_cuda.so; when the binding is compiled into SCIP-SDP itself either this file or the
 * reference's own sdpsolchecker.c can be used.
 */
#include <assert.h>

#include "sdpi/sdpsolchecker.h"
#include "blockmemshell/memory.h"
#include "scip/def.h"
#include "scip/pub_message.h"
#include "devctx.h"

#define MEM_CALL(x) do { if( NULL == (x) ) { SCIPerrorMessage("No memory in function call.\n"); return SCIP_NOMEMORY; } } while( FALSE )

SCIP_RETCODE SCIPsdpSolcheckerCheck(BMS_BUFMEM* bufmem, int nvars, const SCIP_Real* lb, const SCIP_Real* ub, int nsdpblocks,
   const int* sdpblocksizes, const int* sdpnblockvars, int sdpconstnnonz, const int* sdpconstnblocknonz, int* const* sdpconstrow,
   int* const* sdpconstcol, SCIP_Real* const* sdpconstval, int sdpnnonz, int* const* sdpnblockvarnonz, int* const* sdpvar,
   int** const* sdprow, int** const* sdpcol, SCIP_Real** const* sdpval, int* const* indchanges, const int* nremovedinds,
   const int* blockindchanges, int nlpcons, const int* lpindchanges, const SCIP_Real* lplhs, const SCIP_Real* lprhs, int lpnnonz,
   const int* lpbeg, const int* lpind, const SCIP_Real* lpval, const SCIP_Real* solvector, SCIP_Real feastol, SCIP_Real epsilon,
   SCIP_Bool* infeasible)
{
   SCIP_Real* Z = NULL;
   int maxsize = 0;
   int b;
   int i;
   int j;
   int v;

   (void) sdpconstnnonz; (void) sdpnnonz;
   assert( bufmem != NULL && lb != NULL && ub != NULL && solvector != NULL && infeasible != NULL );

   *infeasible = TRUE;
   for( i = 0; i < nvars; ++i )
   {
      if( solvector[i] < lb[i] - feastol || solvector[i] > ub[i] + feastol )
         return SCIP_OKAY;
   }
   for( i = 0; i < nlpcons; ++i )
   {
      SCIP_Real act = 0.0;
      int last = (i == nlpcons - 1) ? lpnnonz : lpbeg[i + 1];
      if( lpindchanges[i] < 0 )
         continue;
      for( j = lpbeg[i]; j < last; ++j )
      {
         if( lb[lpind[j]] < ub[lpind[j]] - epsilon )     /* fixed variables are already part of lhs/rhs */
            act += solvector[lpind[j]] * lpval[j];
      }
      if( act < lplhs[i] - feastol || act > lprhs[i] + feastol )
         return SCIP_OKAY;
   }

   for( b = 0; b < nsdpblocks; ++b )
   {
      if( blockindchanges[b] > -1 )
         maxsize = MAX(maxsize, sdpblocksizes[b] - nremovedinds[b]);
   }
   if( maxsize > 0 )
   {
      sdpcuda_handle* h = sdpiCudaThreadHandle();
      if( h == NULL )
      {
         SCIPerrorMessage("no CUDA device handle available for the solution check (there is no CPU fallback).\n");
         return SCIP_ERROR;
      }
      MEM_CALL( BMSallocBufferMemoryArray(bufmem, &Z, maxsize * maxsize) );
      for( b = 0; b < nsdpblocks; ++b )
      {
         int n = sdpblocksizes[b] - nremovedinds[b];
         int psd = 0;

         if( blockindchanges[b] < 0 )
            continue;
         for( i = 0; i < n * n; ++i )
            Z[i] = 0.0;
         for( v = 0; v < sdpnblockvars[b]; ++v )
         {
            int var = sdpvar[b][v];
            if( lb[var] < ub[var] - epsilon )
            {
               for( i = 0; i < sdpnblockvarnonz[b][v]; ++i )
               {
                  int r = sdprow[b][v][i] - indchanges[b][sdprow[b][v][i]];
                  int c = sdpcol[b][v][i] - indchanges[b][sdpcol[b][v][i]];
                  Z[(size_t)r * n + c] += solvector[var] * sdpval[b][v][i];
                  if( r != c )
                     Z[(size_t)c * n + r] += solvector[var] * sdpval[b][v][i];
               }
            }
         }
         if( sdpconstnblocknonz != NULL )
         {
            for( i = 0; i < sdpconstnblocknonz[b]; ++i )
            {
               int r = sdpconstrow[b][i] - indchanges[b][sdpconstrow[b][i]];
               int c = sdpconstcol[b][i] - indchanges[b][sdpconstcol[b][i]];
               Z[(size_t)r * n + c] -= sdpconstval[b][i];
               if( r != c )
                  Z[(size_t)c * n + r] -= sdpconstval[b][i];
            }
         }
         /* lambda_min(Z) >= -feastol  <=>  Z + feastol I is positive semidefinite; the tiny extra shift makes the
          * Cholesky test robust for matrices sitting exactly on the boundary */
         if( sdpcuda_psd_check(h, n, Z, n, feastol * (1.0 + 1e-6) + 1e-13, &psd) != SDPCUDA_OK )
         {
            BMSfreeBufferMemoryArray(bufmem, &Z);
            SCIPerrorMessage("sdpcuda_psd_check failed.\n");
            return SCIP_ERROR;
         }
         if( !psd )
         {
            BMSfreeBufferMemoryArray(bufmem, &Z);
            return SCIP_OKAY;
         }
      }
      BMSfreeBufferMemoryArray(bufmem, &Z);
   }
   *infeasible = FALSE;
   return SCIP_OKAY;
}
