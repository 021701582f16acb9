/* Minimal stand-in for SCIP's scip/def.h so that the SCIP-SDP `src/sdpi` layer and sdpisolver_cuda.c
 * compile without a SCIP installation (sdpisolver.h:47-48: the interface "can be used independently of any SCIP instance").
 * When building against a real SCIP, drop this directory from the include path. */
#ifndef SHIM_SCIP_DEF_H
#define SHIM_SCIP_DEF_H

#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <limits.h>
#include <float.h>
#include <assert.h>

#ifndef SCIP_EXPORT
#define SCIP_EXPORT __attribute__((visibility("default")))
#endif

#define SCIP_Bool unsigned int
#ifndef TRUE
#define TRUE  1
#define FALSE 0
#endif

#define SCIP_Real double
#define SCIP_REAL_MAX     (SCIP_Real)DBL_MAX
#define SCIP_REAL_MIN    -(SCIP_Real)DBL_MAX
#define SCIP_REAL_FORMAT  "lf"
#define SCIP_Longint long long
#define SCIP_LONGINT_FORMAT "lld"

#define SCIP_DEFAULT_INFINITY         1e+20
#define SCIP_DEFAULT_EPSILON          1e-09
#define SCIP_DEFAULT_MEM_ARRAYGROWFAC   1.2
#define SCIP_DEFAULT_MEM_ARRAYGROWINIT    4
#define SCIP_INVALID        (double)1e+99
#define SCIP_UNKNOWN        (double)1e+98
#define SCIP_MAXSTRLEN      1024

#define REALABS(x)        (fabs(x))
#define EPSEQ(x,y,eps)    (REALABS((x)-(y)) <= (eps))
#define EPSLT(x,y,eps)    ((x)-(y) < -(eps))
#define EPSLE(x,y,eps)    ((x)-(y) <= (eps))
#define EPSGT(x,y,eps)    ((x)-(y) > (eps))
#define EPSGE(x,y,eps)    ((x)-(y) >= -(eps))
#define EPSZ(x,eps)       (REALABS(x) <= (eps))
#define EPSP(x,eps)       ((x) > (eps))
#define EPSN(x,eps)       ((x) < -(eps))
#define EPSFLOOR(x,eps)   (floor((x)+(eps)))
#define EPSCEIL(x,eps)    (ceil((x)-(eps)))
#define EPSISINT(x,eps)   (EPSFLOOR(x,eps) - (x) >= -(eps))

#ifndef ABS
#define ABS(x)        ((x) >= 0 ? (x) : -(x))
#endif
#ifndef SQR
#define SQR(x)        ((x)*(x))
#endif
#ifndef MAX
#define MAX(x,y)      ((x) >= (y) ? (x) : (y))
#define MIN(x,y)      ((x) <= (y) ? (x) : (y))
#endif
#ifndef MAX3
#define MAX3(x,y,z) ((x) >= (y) ? MAX(x,z) : MAX(y,z))
#define MIN3(x,y,z) ((x) <= (y) ? MIN(x,z) : MIN(y,z))
#endif

#define SCIPABORT() assert(FALSE)

#include "scip/type_retcode.h"

#define SCIP_CALL_ABORT(x) do { SCIP_RETCODE _r_; if( (_r_ = (x)) != SCIP_OKAY ) { \
         fprintf(stderr, "[%s:%d] Error <%d> in function call\n", __FILE__, __LINE__, (int)_r_); abort(); } } while( FALSE )
#define SCIP_CALL(x) do { SCIP_RETCODE _restat_; if( (_restat_ = (x)) != SCIP_OKAY ) { \
         fprintf(stderr, "[%s:%d] Error <%d> in function call\n", __FILE__, __LINE__, (int)_restat_); return _restat_; } } while( FALSE )
#define SCIP_ALLOC(x) do { if( NULL == (x) ) { \
         fprintf(stderr, "[%s:%d] No memory in function call\n", __FILE__, __LINE__); return SCIP_NOMEMORY; } } while( FALSE )
#define SCIP_ALLOC_ABORT(x) do { if( NULL == (x) ) { fprintf(stderr, "[%s:%d] No memory\n", __FILE__, __LINE__); abort(); } } while( FALSE )

#endif
