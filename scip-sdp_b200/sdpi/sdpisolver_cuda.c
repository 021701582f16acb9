/* sdpisolver_cuda.c — SCIP-SDP solver binding for the B200-native interior-point solver (libsdpcuda).
 *
 * Fifth implementation of the solver interface src/sdpi/sdpisolver.h (53 SCIPsdpiSolver* functions), next to the
 * reference's sdpisolver_{dsdp.c,sdpa.cpp,mosek.c,none.c}.  Selected at build time like those
 * (CMakeLists.txt:28-29,146-176 `-DSDPS=cuda`, Makefile:46-121 `SDPS=cuda`).  The file is plain C99,
 * includes no CUDA header and talks to the device only through the C ABI of include/sdpcuda.h, so relax_sdp.c,
 * cons_sdp.c and sdpi.c drive it unchanged.
 *
 * What happens in a solve (which follows sdpisolver_sdpa.cpp:1015-1412):
 *   1. variables with ub - lb <= epsilon are fixed and leave the problem (objective -> fixedobj, value remembered);
 *   2. SDP blocks / rows / columns flagged by blockindchanges / indchanges are dropped, indices are compressed;
 *   3. every kept LP row becomes up to two one-sided rows (lhs first, then rhs), every finite bound of an active
 *      variable becomes a one-sided row, in this order; in penalty mode the variable r is appended with +1 on the
 *      diagonal of all SDP blocks and in all LP row sides (not in the variable bounds) and, if rbound, a row r >= 0;
 *   4. the flat problem goes to the device with one call; the interior-point iteration runs there;
 *   5. the returned point is re-checked against the caller's feastol with the reference's own checker
 *      (SCIPsdpSolcheckerCheck) and the absolute duality gap; tolerances are tightened and the solve repeated like
 *      checkFeastolAndResolve (sdpisolver_sdpa.cpp:368-494); if the result is still not acceptable the more conservative
 *      settings MEDIUM and STABLE are tried (sdpisolver_sdpa.cpp:1698-1795).
 */
#include <assert.h>
#include <stdlib.h>
#include <string.h>

#include "sdpi/sdpisolver.h"
#include "sdpi/sdpsolchecker.h"
#include "blockmemshell/memory.h"
#include "scip/def.h"
#include "scip/pub_message.h"

#include "sdpcuda.h"

#define CUDASDP_MIN_PENALTYPARAM     1e5    /* bounds and factors for the computed penalty parameter: same policy as */
#define CUDASDP_MAX_PENALTYPARAM     1e12   /* sdpisolver_dsdp.c:64-68 / sdpisolver_sdpa.cpp:111-115 so that the penalty */
#define CUDASDP_PENALTYPARAM_FACTOR  1e1    /* ladder of sdpi.c:3437-3619 behaves identically */
#define CUDASDP_MAX_MAXPENALTYPARAM  1e15
#define CUDASDP_MAXPENALTY_FACTOR    1e6
#define CUDASDP_PENALTYBOUNDTOL      1e-3   /* (Gamma - tr X)/Gamma below this: primal bound reached */
#define CUDASDP_TIGHTEN              0.1    /* factor for tolerances when the post-check fails */
#define CUDASDP_MINTOL               1e-9   /* do not tighten tolerances below this */
#define CUDASDP_INF                  1e20

#define MEM_CALL(x) do { if( NULL == (x) ) { SCIPerrorMessage("No memory in function call.\n"); return SCIP_NOMEMORY; } } while( FALSE )
/* inside LoadAndSolveWithPenalty: the buffer arrays allocated so far are released at TERMINATE */
#define MEM_GOTO(x) do { if( NULL == (x) ) { SCIPerrorMessage("No memory in function call.\n"); retcode = SCIP_NOMEMORY; goto TERMINATE; } } while( FALSE )
#define NEED_SOLVED(s) do { if( !(s)->solved ) { \
      SCIPerrorMessage("Tried to access solution information for SDP %d ahead of solving!\n", (s)->sdpcounter); return SCIP_LPERROR; } } while( FALSE )
#define NEED_SOLVED_BOOL(s) do { if( !(s)->solved ) { \
      SCIPerrorMessage("Tried to access solution information for SDP %d ahead of solving!\n", (s)->sdpcounter); return FALSE; } } while( FALSE )

/** solver interface data */
struct SCIP_SDPiSolver
{
   SCIP_MESSAGEHDLR*     messagehdlr;
   BMS_BLKMEM*           blkmem;
   BMS_BUFMEM*           bufmem;
   sdpcuda_handle*       dev;                /**< device-side solver object (owns stream and device memory) */

   /* mapping data of the last loaded problem */
   int                   nvars;              /**< variables of the input problem */
   int                   nactive;            /**< variables handed to the device (without r) */
   int                   nfixed;
   int                   varcap;             /**< allocated length of the per-variable arrays */
   int*                  var2act;            /**< [nvars] k >= 0: active index, -(k+1): k-th fixed variable */
   int*                  act2var;            /**< [nactive] */
   SCIP_Real*            fixedval;           /**< [nfixed] */
   SCIP_Real*            actobj;             /**< [nactive] objective handed to the device (zeros if !withobj) */
   SCIP_Real*            realobj;            /**< [nactive] true objective coefficients (for GetObjval) */
   int*                  boundslot;          /**< [2*nvars] LP-block position of lb (2j) / ub (2j+1) multiplier or -1 */
   int                   nsdpblocks;         /**< SDP blocks of the input problem */
   int                   blockcap;
   int*                  blk2dev;            /**< [nsdpblocks] device block index or -1 */
   int*                  devsize;            /**< [nsdpblocks] reduced size of the block on the device */
   int*                  origsize;           /**< [nsdpblocks] */
   int**                 red2orig;           /**< [nsdpblocks][devsize] original row/col index of a reduced index */
   int*                  red2origcap;
   int                   nlpcons;            /**< LP rows of the input problem */
   int                   rowcap;
   int*                  rowslot;            /**< [2*nlpcons] LP-block position of the lhs (2i) / rhs (2i+1) multiplier or -1 */
   int                   nrowsides;          /**< number of one-sided rows coming from LP rows */
   int                   nboundrows;         /**< number of one-sided rows coming from variable bounds (without r >= 0) */
   int                   nlprows;            /**< total rows of the device LP block */
   int                   ndevblocks;

   /* host copy of the last solution */
   SCIP_Real*            y;                  /**< [nactive (+1)] */
   int                   ycap;
   SCIP_Real*            xlp;                /**< [nlprows] multipliers of the device LP block (lazily fetched) */
   int                   xlpcap;
   SCIP_Bool             xlpvalid;
   /* preoptimal point of the last solve (host copies, reduced sizes; sdpisolver_sdpa.cpp:191-195) */
   SCIP_Bool             preoptexists;
   SCIP_Real*            preopty;            /**< [nactive] */
   int                   preoptycap;
   SCIP_Real**           preoptX;            /**< [nsdpblocks] dense reduced blocks */
   int*                  preoptXcap;
   SCIP_Real*            preoptxlp;          /**< [nlprows] */
   int                   preoptxlpcap;
   SCIP_Bool             wantpreopt;         /**< ask the device for a preoptimal point in the next run */
   SCIP_Real**           X;                  /**< [nsdpblocks] dense reduced multiplier blocks (lazily fetched) */
   int*                  Xcap;
   SCIP_Bool*            Xvalid;

   /* problem resident on the device: hash over every array handed to sdpcuda_solve except obj and lprhs */
   SCIP_Bool             residentvalid;
   unsigned long long    residenthash;
   int                   nuploads;           /**< full uploads (sdpcuda_solve) since creation */
   int                   npatched;           /**< re-solves on the resident problem (sdpcuda_solve_patched) since creation */
   SCIP_Real             h2dbytes;           /**< host->device bytes of all solves since creation */

   SCIP_Bool             devicecheck;        /**< post-check of the SDP blocks on the problem resident on the device (SDPCUDA_DEVICE_CHECK=1) */

   /* status */
   sdpcuda_result        res;
   SCIP_Bool             solved;
   SCIP_Bool             timelimit;
   SCIP_Bool             penalty;
   SCIP_Bool             rbound;
   SCIP_Bool             feasorig;
   SCIP_Real             fixedobj;
   SCIP_SDPSOLVERSETTING usedsetting;
   int                   sdpcounter;
   int                   niterations;
   int                   nsdpcalls;
   SCIP_Real             opttime;

   /* parameters */
   SCIP_Real             epsilon;
   SCIP_Real             gaptol;
   SCIP_Real             feastol;
   SCIP_Real             sdpsolverfeastol;
   SCIP_Real             objlimit;
   SCIP_Real             lambdastar;
   SCIP_Real             preoptimalgap;
   SCIP_Bool             sdpinfo;
   int                   nthreads;
};

/*
 * local helpers
 */

static SCIP_Bool isInf(SCIP_Real v) { return v <= -CUDASDP_INF || v >= CUDASDP_INF; }

static SCIP_Bool isFixedVar(const SCIP_SDPISOLVER* s, SCIP_Real lb, SCIP_Real ub) { return (ub - lb) <= s->epsilon; }

static SCIP_RETCODE growInt(BMS_BLKMEM* mem, int** arr, int* cap, int need)
{
   if( need > *cap )
   {
      int newcap = MAX(need, 2 * (*cap));
      if( *arr == NULL ) { MEM_CALL( BMSallocBlockMemoryArray(mem, arr, newcap) ); }
      else { MEM_CALL( BMSreallocBlockMemoryArray(mem, arr, *cap, newcap) ); }
      *cap = newcap;
   }
   return SCIP_OKAY;
}

static SCIP_RETCODE growReal(BMS_BLKMEM* mem, SCIP_Real** arr, int* cap, int need)
{
   if( need > *cap )
   {
      int newcap = MAX(need, 2 * (*cap));
      if( *arr == NULL ) { MEM_CALL( BMSallocBlockMemoryArray(mem, arr, newcap) ); }
      else { MEM_CALL( BMSreallocBlockMemoryArray(mem, arr, *cap, newcap) ); }
      *cap = newcap;
   }
   return SCIP_OKAY;
}

/** makes sure all per-variable, per-block and per-row mapping arrays are large enough */
static SCIP_RETCODE ensureMaps(SCIP_SDPISOLVER* s, int nvars, int nsdpblocks, const int* sdpblocksizes, int nlpcons)
{
   int b;

   if( nvars > s->varcap )
   {
      int oldcap = s->varcap;
      int newcap = MAX(nvars, 2 * oldcap);
      int c;
      c = oldcap; SCIP_CALL( growInt(s->blkmem, &s->var2act, &c, newcap) );
      c = oldcap; SCIP_CALL( growInt(s->blkmem, &s->act2var, &c, newcap) );
      c = oldcap; SCIP_CALL( growReal(s->blkmem, &s->fixedval, &c, newcap) );
      c = oldcap; SCIP_CALL( growReal(s->blkmem, &s->actobj, &c, newcap) );
      c = oldcap; SCIP_CALL( growReal(s->blkmem, &s->realobj, &c, newcap) );
      c = 2 * oldcap; SCIP_CALL( growInt(s->blkmem, &s->boundslot, &c, 2 * newcap) );
      s->varcap = newcap;
   }
   if( nsdpblocks > s->blockcap )
   {
      int oldcap = s->blockcap;
      int newcap = MAX(nsdpblocks, 2 * oldcap);
      int c;
      c = oldcap; SCIP_CALL( growInt(s->blkmem, &s->blk2dev, &c, newcap) );
      c = oldcap; SCIP_CALL( growInt(s->blkmem, &s->devsize, &c, newcap) );
      c = oldcap; SCIP_CALL( growInt(s->blkmem, &s->origsize, &c, newcap) );
      c = oldcap; SCIP_CALL( growInt(s->blkmem, &s->red2origcap, &c, newcap) );
      c = oldcap; SCIP_CALL( growInt(s->blkmem, &s->Xcap, &c, newcap) );
      c = oldcap; SCIP_CALL( growInt(s->blkmem, &s->preoptXcap, &c, newcap) );
      if( s->red2orig == NULL )
      {
         MEM_CALL( BMSallocBlockMemoryArray(s->blkmem, &s->red2orig, newcap) );
         MEM_CALL( BMSallocBlockMemoryArray(s->blkmem, &s->X, newcap) );
         MEM_CALL( BMSallocBlockMemoryArray(s->blkmem, &s->preoptX, newcap) );
         MEM_CALL( BMSallocBlockMemoryArray(s->blkmem, &s->Xvalid, newcap) );
      }
      else
      {
         MEM_CALL( BMSreallocBlockMemoryArray(s->blkmem, &s->red2orig, oldcap, newcap) );
         MEM_CALL( BMSreallocBlockMemoryArray(s->blkmem, &s->X, oldcap, newcap) );
         MEM_CALL( BMSreallocBlockMemoryArray(s->blkmem, &s->preoptX, oldcap, newcap) );
         MEM_CALL( BMSreallocBlockMemoryArray(s->blkmem, &s->Xvalid, oldcap, newcap) );
      }
      for( b = oldcap; b < newcap; ++b )
      {
         s->red2orig[b] = NULL; s->red2origcap[b] = 0;
         s->X[b] = NULL; s->Xcap[b] = 0; s->Xvalid[b] = FALSE;
         s->preoptX[b] = NULL; s->preoptXcap[b] = 0;
      }
      s->blockcap = newcap;
   }
   for( b = 0; b < nsdpblocks; ++b )
   {
      SCIP_CALL( growInt(s->blkmem, &s->red2orig[b], &s->red2origcap[b], sdpblocksizes[b]) );
   }
   if( nlpcons > s->rowcap )
   {
      int c = 2 * s->rowcap;
      SCIP_CALL( growInt(s->blkmem, &s->rowslot, &c, 2 * MAX(nlpcons, 2 * s->rowcap)) );
      s->rowcap = c / 2;
   }
   return SCIP_OKAY;
}

/** fetches the LP multipliers of the last solve from the device if not done yet */
static SCIP_RETCODE fetchXlp(SCIP_SDPISOLVER* s)
{
   if( !s->xlpvalid )
   {
      SCIP_CALL( growReal(s->blkmem, &s->xlp, &s->xlpcap, MAX(s->nlprows, 1)) );
      if( s->nlprows > 0 && sdpcuda_get_xlp(s->dev, s->xlp) != SDPCUDA_OK )
         return SCIP_LPERROR;
      s->xlpvalid = TRUE;
   }
   return SCIP_OKAY;
}

/** fetches one multiplier block (reduced size, dense) from the device if not done yet */
static SCIP_RETCODE fetchX(SCIP_SDPISOLVER* s, int b)
{
   assert( s->blk2dev[b] >= 0 );
   if( !s->Xvalid[b] )
   {
      int n = s->devsize[b];
      SCIP_CALL( growReal(s->blkmem, &s->X[b], &s->Xcap[b], n * n) );
      if( sdpcuda_get_X(s->dev, s->blk2dev[b], s->X[b]) != SDPCUDA_OK )
         return SCIP_LPERROR;
      s->Xvalid[b] = TRUE;
   }
   return SCIP_OKAY;
}

/** runs one device solve with the given tolerances/settings and pulls y */
/** FNV-1a over the structure of a solver-form problem: everything except the two vectors sdpcuda_solve_patched ships (obj, lprhs) */
static unsigned long long structureHash(const sdpcuda_problem* p)
{
   unsigned long long h = 1469598103934665603ULL;
   const int nnz = p->m > 0 ? p->varbeg[p->m] : 0;
   const int lnz = p->nlp > 0 ? p->lpbeg[p->nlp] : 0;
#define MIX(ptr, bytes) do { const unsigned char* q_ = (const unsigned char*)(ptr); size_t n_ = (size_t)(bytes); size_t t_; \
      for( t_ = 0; t_ < n_; ++t_ ) { h ^= q_[t_]; h *= 1099511628211ULL; } } while( FALSE )
   MIX(&p->m, sizeof(int)); MIX(&p->nblocks, sizeof(int)); MIX(&p->cnnz, sizeof(int)); MIX(&p->nlp, sizeof(int));
   MIX(p->blocksizes, sizeof(int) * p->nblocks);
   MIX(p->varbeg, sizeof(int) * (p->m + 1));
   MIX(p->entblk, sizeof(int) * nnz); MIX(p->entrow, sizeof(int) * nnz); MIX(p->entcol, sizeof(int) * nnz); MIX(p->entval, sizeof(double) * nnz);
   MIX(p->cblk, sizeof(int) * p->cnnz); MIX(p->crow, sizeof(int) * p->cnnz); MIX(p->ccol, sizeof(int) * p->cnnz); MIX(p->cval, sizeof(double) * p->cnnz);
   if( p->nlp > 0 )
   {
      MIX(p->lpbeg, sizeof(int) * (p->nlp + 1)); MIX(p->lpind, sizeof(int) * lnz); MIX(p->lpval, sizeof(double) * lnz);
   }
#undef MIX
   return h;
}

static SCIP_RETCODE runDevice(SCIP_SDPISOLVER* s, const sdpcuda_problem* prob, unsigned long long hash, SCIP_Real gaptol, SCIP_Real feastol,
   int setting, SCIP_Real timeleft, const SCIP_Real* starty)
{
   sdpcuda_params par;
   int b;
   int rc;

   sdpcuda_default_params(&par);
   par.gaptol = gaptol;
   par.feastol = feastol;
   par.absgaptol = s->penalty ? -1.0 : 0.5 * s->gaptol;   /* the post-check below is on the absolute gap */
   par.objlimit = s->objlimit;
   par.lambdastar = -1.0;
   par.timelimit = timeleft;
   par.setting = setting;
   par.verbose = s->sdpinfo ? 1 : 0;
   par.preoptgap = s->wantpreopt ? s->preoptimalgap : -1.0;

   /* the problem of the previous solve is still in HBM: tolerance tightening, the settings ladder, a growing penalty parameter
    * (only the objective coefficient of r changes) and repeated solves of one node ship obj and lprhs only (sdpi.c:3437-3619) */
   if( s->residentvalid && s->residenthash == hash )
   {
      rc = sdpcuda_solve_patched(s->dev, prob, &par, starty, &s->res);
      ++s->npatched;
   }
   else
   {
      s->residentvalid = FALSE;
      rc = sdpcuda_solve(s->dev, prob, &par, starty, &s->res);
      ++s->nuploads;
   }
   if( rc == SDPCUDA_OK )
   {
      s->residentvalid = TRUE;
      s->residenthash = hash;
      s->h2dbytes += s->res.h2d_bytes;
   }
   if( rc == SDPCUDA_ERR_NOMEM )
      return SCIP_NOMEMORY;
   if( rc != SDPCUDA_OK )
   {
      SCIPerrorMessage("sdpcuda_solve failed with code %d (backend %s) - no CPU fallback exists.\n", rc, sdpcuda_backend_name());
      return SCIP_LPERROR;
   }
   s->solved = TRUE;
   s->niterations += s->res.iterations;
   s->nsdpcalls += 1;
   s->opttime += s->res.seconds;
   if( s->res.stop == SDPCUDA_STOP_TIMELIMIT )
      s->timelimit = TRUE;
   s->xlpvalid = FALSE;
   for( b = 0; b < s->nsdpblocks; ++b )
      s->Xvalid[b] = FALSE;
   SCIP_CALL( growReal(s->blkmem, &s->y, &s->ycap, prob->m + 1) );
   if( prob->m > 0 && sdpcuda_get_y(s->dev, s->y) != SDPCUDA_OK )
      return SCIP_LPERROR;
   if( s->wantpreopt )
   {
      /* pull the preoptimal iterate right away: later runs of the settings ladder overwrite it on the device
       * (the reference keeps host copies as well, sdpisolver_sdpa.cpp:1622-1653) */
      int exists = 0;
      s->wantpreopt = FALSE;
      SCIP_CALL( growReal(s->blkmem, &s->preopty, &s->preoptycap, prob->m + 1) );
      SCIP_CALL( growReal(s->blkmem, &s->preoptxlp, &s->preoptxlpcap, MAX(s->nlprows, 1)) );
      if( sdpcuda_get_preopt(s->dev, &exists, s->preopty, s->preoptxlp) != SDPCUDA_OK )
         return SCIP_LPERROR;
      if( exists )
      {
         for( b = 0; b < s->nsdpblocks; ++b )
         {
            if( s->blk2dev[b] < 0 )
               continue;
            SCIP_CALL( growReal(s->blkmem, &s->preoptX[b], &s->preoptXcap[b], s->devsize[b] * s->devsize[b]) );
            if( sdpcuda_get_preopt_X(s->dev, s->blk2dev[b], s->preoptX[b]) != SDPCUDA_OK )
               return SCIP_LPERROR;
         }
         s->preoptexists = TRUE;
      }
   }
   return SCIP_OKAY;
}

/*
 * Miscellaneous Methods
 */

/** not part of sdpisolver.h: transfer statistics of this solver object for the tests of the resident re-solve path */
void SCIPsdpiSolverCudaGetTransferStats(SCIP_SDPISOLVER* sdpisolver, int* nuploads, int* npatched, SCIP_Real* h2dbytes)
{
   assert( sdpisolver != NULL );
   if( nuploads != NULL )
      *nuploads = sdpisolver->nuploads;
   if( npatched != NULL )
      *npatched = sdpisolver->npatched;
   if( h2dbytes != NULL )
      *h2dbytes = sdpisolver->h2dbytes;
}

/** not part of sdpisolver.h: what computeConflictCut (relax_sdp.c:1030-1099) takes from the primal solution, computed
 *  on the device-resident X instead of shipping every dense block to the host: for every block b and block variable v the inner
 *  product <A_v^b, X_b> (varproducts[b][v]), <A_0^b, X_b> (constproducts[b]) and a certified lower bound of min(lambda_min(X_b), 0)
 *  (mineigbounds[b]).  The SDP data are given like SCIPsdpiGetSDPdata returns them: original (unreduced) indices, lower triangle;
 *  entries in rows/columns that were removed before the solve meet zeros of X.  INTEGRATION.md shows the caller-side change. */
SCIP_RETCODE SCIPsdpiSolverGetPrimalInnerProducts(SCIP_SDPISOLVER* sdpisolver, int nsdpblocks, const int* sdpnblockvars,
   int* const* sdpnblockvarnonz, int** const* sdprow, int** const* sdpcol, SCIP_Real** const* sdpval, const int* sdpconstnblocknonz,
   int* const* sdpconstrow, int* const* sdpconstcol, SCIP_Real* const* sdpconstval, SCIP_Real* const* varproducts, SCIP_Real* constproducts,
   SCIP_Real* mineigbounds)
{
   SCIP_SDPISOLVER* s = sdpisolver;
   SCIP_RETCODE retcode = SCIP_OKAY;
   int* groupbeg = NULL;
   int* eblk = NULL;
   int* erow = NULL;
   int* ecol = NULL;
   int* orig2red = NULL;
   SCIP_Real* evalv = NULL;
   SCIP_Real* out = NULL;
   int ngroups = 0;
   int nent = 0;
   int maxsize = 0;
   int b;
   int v;
   int k;
   int g;

   assert( s != NULL );
   NEED_SOLVED(s);
   if( nsdpblocks != s->nsdpblocks )
      return SCIP_LPERROR;
   for( b = 0; b < nsdpblocks; ++b )
   {
      ngroups += sdpnblockvars[b] + 1;
      nent += sdpconstnblocknonz != NULL ? sdpconstnblocknonz[b] : 0;
      for( v = 0; v < sdpnblockvars[b]; ++v )
         nent += sdpnblockvarnonz[b][v];
      maxsize = MAX(maxsize, s->origsize[b]);
   }
   MEM_CALL( BMSallocBufferMemoryArray(s->bufmem, &groupbeg, ngroups + 1) );
   if( NULL == BMSallocBufferMemoryArray(s->bufmem, &eblk, nent + 1) || NULL == BMSallocBufferMemoryArray(s->bufmem, &erow, nent + 1)
      || NULL == BMSallocBufferMemoryArray(s->bufmem, &ecol, nent + 1) || NULL == BMSallocBufferMemoryArray(s->bufmem, &evalv, nent + 1)
      || NULL == BMSallocBufferMemoryArray(s->bufmem, &out, ngroups + 1) || NULL == BMSallocBufferMemoryArray(s->bufmem, &orig2red, maxsize + 1) )
   {
      retcode = SCIP_NOMEMORY;
      goto DONE;
   }
   nent = 0;
   g = 0;
   for( b = 0; b < nsdpblocks; ++b )
   {
      const int dev = s->blk2dev[b];

      for( k = 0; k < s->origsize[b]; ++k )
         orig2red[k] = -1;
      for( k = 0; dev >= 0 && k < s->devsize[b]; ++k )
         orig2red[s->red2orig[b][k]] = k;
      for( v = 0; v <= sdpnblockvars[b]; ++v )               /* v == nblockvars: the constant matrix */
      {
         const int cnt = v < sdpnblockvars[b] ? sdpnblockvarnonz[b][v] : (sdpconstnblocknonz != NULL ? sdpconstnblocknonz[b] : 0);
         const int* rws = v < sdpnblockvars[b] ? sdprow[b][v] : (cnt > 0 ? sdpconstrow[b] : NULL);
         const int* cls = v < sdpnblockvars[b] ? sdpcol[b][v] : (cnt > 0 ? sdpconstcol[b] : NULL);
         const SCIP_Real* vls = v < sdpnblockvars[b] ? sdpval[b][v] : (cnt > 0 ? sdpconstval[b] : NULL);

         groupbeg[g++] = nent;
         for( k = 0; dev >= 0 && k < cnt; ++k )
         {
            const int r = orig2red[rws[k]];
            const int c = orig2red[cls[k]];

            if( r < 0 || c < 0 )
               continue;
            eblk[nent] = dev;
            erow[nent] = MAX(r, c);
            ecol[nent] = MIN(r, c);
            evalv[nent] = vls[k];
            ++nent;
         }
      }
   }
   groupbeg[g] = nent;
   assert( g == ngroups );
   if( sdpcuda_primal_products(s->dev, ngroups, groupbeg, eblk, erow, ecol, evalv, out) != SDPCUDA_OK )
   {
      SCIPerrorMessage("sdpcuda_primal_products failed.\n");
      retcode = SCIP_LPERROR;
      goto DONE;
   }
   g = 0;
   for( b = 0; b < nsdpblocks; ++b )
   {
      for( v = 0; v < sdpnblockvars[b]; ++v )
         varproducts[b][v] = out[g++];
      constproducts[b] = out[g++];
      mineigbounds[b] = 0.0;
      if( s->blk2dev[b] >= 0 && sdpcuda_primal_mineig_bound(s->dev, s->blk2dev[b], &mineigbounds[b]) != SDPCUDA_OK )
      {
         SCIPerrorMessage("sdpcuda_primal_mineig_bound failed.\n");
         retcode = SCIP_LPERROR;
         goto DONE;
      }
   }

DONE:
   BMSfreeBufferMemoryArrayNull(s->bufmem, &orig2red);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &out);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &evalv);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &ecol);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &erow);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &eblk);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &groupbeg);
   return retcode;
}

const char* SCIPsdpiSolverGetSolverName(void)
{
   /* neither "DSDP" nor "SDPA" nor containing "Mosek": the caller then passes start settings of the parent node and
    * stores SettingsUsed per node (relax_sdp.c:4085-4120,4194-4203) but does not push lambdastar (relax_sdp.c:4823) */
   return "CUDA-IPM";
}

const char* SCIPsdpiSolverGetSolverDesc(void)
{
   return "B200-native primal-dual interior-point SDP solver (HKM direction, Mehrotra predictor-corrector; sm_100a CUDA kernels behind libsdpcuda)";
}

void* SCIPsdpiSolverGetSolverPointer(SCIP_SDPISOLVER* sdpisolver)
{
   assert( sdpisolver != NULL );
   return (void*) sdpisolver->dev;
}

int SCIPsdpiSolverGetDefaultSdpiSolverNpenaltyIncreases(void)
{
   return 8;
}

SCIP_Bool SCIPsdpiSolverDoesWarmstartNeedPrimal(void)
{
   return TRUE;
}

/*
 * SDPI Creation and Destruction Methods
 */

SCIP_RETCODE SCIPsdpiSolverCreate(SCIP_SDPISOLVER** sdpisolver, SCIP_MESSAGEHDLR* messagehdlr, BMS_BLKMEM* blkmem, BMS_BUFMEM* bufmem)
{
   SCIP_SDPISOLVER* s;
   int rc;

   assert( sdpisolver != NULL );
   assert( blkmem != NULL );
   assert( bufmem != NULL );

   MEM_CALL( BMSallocBlockMemory(blkmem, sdpisolver) );
   s = *sdpisolver;
   memset(s, 0, sizeof(*s));
   s->messagehdlr = messagehdlr;
   s->blkmem = blkmem;
   s->bufmem = bufmem;

   /* one device object per solver object: in concurrent mode every SCIP thread owns its own SDPI (relax_sdp.c:5041), the
    * device library assigns GPUs round-robin and gives each object a private stream (cf. sdpisolver_mosek.c:97-109) */
   rc = sdpcuda_create(&s->dev, -1);
   if( rc != SDPCUDA_OK )
   {
      SCIPerrorMessage("sdpcuda_create failed with code %d: no usable CUDA device / library (there is no CPU fallback).\n", rc);
      BMSfreeBlockMemory(blkmem, sdpisolver);
      return rc == SDPCUDA_ERR_NOMEM ? SCIP_NOMEMORY : SCIP_LPERROR;
   }

   s->solved = FALSE;
   s->timelimit = FALSE;
   s->sdpcounter = 0;
   s->residentvalid = FALSE;
   s->residenthash = 0ULL;
   s->nuploads = 0;
   s->npatched = 0;
   s->h2dbytes = 0.0;
   s->usedsetting = SCIP_SDPSOLVERSETTING_UNSOLVED;
   s->epsilon = 1e-9;
   s->gaptol = 1e-4;
   s->feastol = 1e-6;
   s->sdpsolverfeastol = 1e-6;
   s->objlimit = CUDASDP_INF;
   s->lambdastar = -1.0;
   s->preoptimalgap = -1.0;
   {
      /* the post-check of the SDP blocks on the problem that is resident on the device (measured on max-cut 2000: 134.1 -> 127.4 ms per
       * LoadAndSolve); SDPCUDA_DEVICE_CHECK=0 builds Z(y) on the host and ships it instead */
      const char* e = getenv("SDPCUDA_DEVICE_CHECK");
      s->devicecheck = !(e != NULL && e[0] == '0');
   }
   s->sdpinfo = FALSE;
   s->nthreads = -1;

   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverFree(SCIP_SDPISOLVER** sdpisolver)
{
   SCIP_SDPISOLVER* s;
   int b;

   assert( sdpisolver != NULL );
   assert( *sdpisolver != NULL );
   s = *sdpisolver;

   if( s->dev != NULL )
      (void) sdpcuda_destroy(s->dev);

   for( b = 0; b < s->blockcap; ++b )
   {
      BMSfreeBlockMemoryArrayNull(s->blkmem, &s->red2orig[b], s->red2origcap[b]);
      BMSfreeBlockMemoryArrayNull(s->blkmem, &s->X[b], s->Xcap[b]);
      BMSfreeBlockMemoryArrayNull(s->blkmem, &s->preoptX[b], s->preoptXcap[b]);
   }
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->red2orig, s->blockcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->X, s->blockcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->preoptX, s->blockcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->preoptXcap, s->blockcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->preopty, s->preoptycap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->preoptxlp, s->preoptxlpcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->Xvalid, s->blockcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->Xcap, s->blockcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->red2origcap, s->blockcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->origsize, s->blockcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->devsize, s->blockcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->blk2dev, s->blockcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->rowslot, 2 * s->rowcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->boundslot, 2 * s->varcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->realobj, s->varcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->actobj, s->varcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->fixedval, s->varcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->act2var, s->varcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->var2act, s->varcap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->y, s->ycap);
   BMSfreeBlockMemoryArrayNull(s->blkmem, &s->xlp, s->xlpcap);

   BMSfreeBlockMemory(s->blkmem, sdpisolver);
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverIncreaseCounter(SCIP_SDPISOLVER* sdpisolver)
{
   assert( sdpisolver != NULL );
   sdpisolver->sdpcounter++;
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverResetCounter(SCIP_SDPISOLVER* sdpisolver)
{
   assert( sdpisolver != NULL );
   sdpisolver->sdpcounter = 0;
   return SCIP_OKAY;
}

/*
 * Solving Methods
 */

SCIP_RETCODE SCIPsdpiSolverLoadAndSolve(
   SCIP_SDPISOLVER* sdpisolver, int nvars, const SCIP_Real* obj, const SCIP_Real* lb, const SCIP_Real* ub, int nsdpblocks,
   const int* sdpblocksizes, const int* sdpnblockvars, int sdpconstnnonz, const int* sdpconstnblocknonz, int* const* sdpconstrow,
   int* const* sdpconstcol, SCIP_Real* const* sdpconstval, int sdpnnonz, int* const* sdpnblockvarnonz, int* const* sdpvar,
   int** const* sdprow, int** const* sdpcol, SCIP_Real** const* sdpval, int* const* indchanges, const int* nremovedinds,
   const int* blockindchanges, int nremovedblocks, int nlpcons, const int* lpindchanges, const SCIP_Real* lplhs,
   const SCIP_Real* lprhs, int lpnnonz, const int* lpbeg, const int* lpind, const SCIP_Real* lpval, const SCIP_Real* starty,
   const int* startZnblocknonz, int* const* startZrow, int* const* startZcol, SCIP_Real* const* startZval,
   const int* startXnblocknonz, int* const* startXrow, int* const* startXcol, SCIP_Real* const* startXval,
   SCIP_SDPSOLVERSETTING startsettings, SCIP_Real timelimit, SDPI_CLOCK* usedsdpitime)
{
   return SCIPsdpiSolverLoadAndSolveWithPenalty(sdpisolver, 0.0, TRUE, FALSE, nvars, obj, lb, ub, nsdpblocks, sdpblocksizes,
      sdpnblockvars, sdpconstnnonz, sdpconstnblocknonz, sdpconstrow, sdpconstcol, sdpconstval, sdpnnonz, sdpnblockvarnonz, sdpvar,
      sdprow, sdpcol, sdpval, indchanges, nremovedinds, blockindchanges, nremovedblocks, nlpcons, lpindchanges, lplhs, lprhs,
      lpnnonz, lpbeg, lpind, lpval, starty, startZnblocknonz, startZrow, startZcol, startZval, startXnblocknonz, startXrow,
      startXcol, startXval, startsettings, timelimit, usedsdpitime, NULL, NULL);
}

SCIP_RETCODE SCIPsdpiSolverLoadAndSolveWithPenalty(
   SCIP_SDPISOLVER* sdpisolver, SCIP_Real penaltyparam, SCIP_Bool withobj, SCIP_Bool rbound, int nvars, const SCIP_Real* obj,
   const SCIP_Real* lb, const SCIP_Real* ub, int nsdpblocks, const int* sdpblocksizes, const int* sdpnblockvars, int sdpconstnnonz,
   const int* sdpconstnblocknonz, int* const* sdpconstrow, int* const* sdpconstcol, SCIP_Real* const* sdpconstval, int sdpnnonz,
   int* const* sdpnblockvarnonz, int* const* sdpvar, int** const* sdprow, int** const* sdpcol, SCIP_Real** const* sdpval,
   int* const* indchanges, const int* nremovedinds, const int* blockindchanges, int nremovedblocks, int nlpcons,
   const int* lpindchanges, const SCIP_Real* lplhs, const SCIP_Real* lprhs, int lpnnonz, const int* lpbeg, const int* lpind,
   const SCIP_Real* lpval, const SCIP_Real* starty, const int* startZnblocknonz, int* const* startZrow, int* const* startZcol,
   SCIP_Real* const* startZval, const int* startXnblocknonz, int* const* startXrow, int* const* startXcol,
   SCIP_Real* const* startXval, SCIP_SDPSOLVERSETTING startsettings, SCIP_Real timelimit, SDPI_CLOCK* usedsdpitime,
   SCIP_Bool* feasorig, SCIP_Bool* penaltybound)
{
   SCIP_SDPISOLVER* s = sdpisolver;
   sdpcuda_problem prob;
   unsigned long long probhash;
   SCIP_RETCODE retcode = SCIP_OKAY;
   SCIP_Real timeleft;
   SCIP_Real curgaptol;
   SCIP_Real curfeastol;
   SCIP_Real* devstart = NULL;
   SCIP_Bool withr;
   int* varbeg = NULL;
   int* fill = NULL;
   int* entblk = NULL;
   int* entrow = NULL;
   int* entcol = NULL;
   SCIP_Real* entval = NULL;
   int* cblk = NULL;
   int* crow = NULL;
   int* ccol = NULL;
   SCIP_Real* cval = NULL;
   int* devblocksizes = NULL;
   int* rowbeg = NULL;
   int* rowind = NULL;
   SCIP_Real* rowval = NULL;
   SCIP_Real* rowrhs = NULL;
   SCIP_Real* devobj = NULL;
   int m;
   int nnz;
   int cnnz;
   int lpcap;
   int lpnz;
   int nrows;
   int setting;
   int lastsetting;
   int i;
   int j;
   int b;
   int v;
   int k;

   assert( s != NULL );
   assert( penaltyparam > -1 * s->epsilon );
   assert( penaltyparam < s->epsilon || feasorig != NULL );
   assert( nvars > 0 );
   assert( obj != NULL && lb != NULL && ub != NULL );
   assert( nsdpblocks >= 0 );
   assert( nlpcons >= 0 );
   (void) sdpnnonz; (void) nremovedblocks;
   (void) startZnblocknonz; (void) startZrow; (void) startZcol; (void) startZval;
   s->preoptexists = FALSE;
   s->wantpreopt = FALSE;
   s->niterations = 0;
   s->nsdpcalls = 0;
   s->opttime = 0.0;
   s->feasorig = FALSE;
   s->solved = FALSE;

   /* remaining time (sdpisolver_dsdp.c:880-890: an exhausted limit is not an error) */
   timeleft = timelimit;
   if( !isInf(timeleft) )
      timeleft -= SDPIclockGetTime(usedsdpitime);
   if( timeleft <= 0.0 )
   {
      s->timelimit = TRUE;
      return SCIP_OKAY;
   }
   s->timelimit = FALSE;
   s->usedsetting = SCIP_SDPSOLVERSETTING_UNSOLVED;

   /* the counter is only increased for the original problem, a penalty formulation is still the same SDP */
   if( penaltyparam < s->epsilon )
      ++s->sdpcounter;

   s->penalty = (penaltyparam >= s->epsilon);
   s->rbound = rbound;
   withr = s->penalty;

   SCIP_CALL( ensureMaps(s, nvars, nsdpblocks, sdpblocksizes, nlpcons) );
   s->nvars = nvars;
   s->nsdpblocks = nsdpblocks;
   s->nlpcons = nlpcons;

   /* ---- 1. active and fixed variables ---- */
   s->nactive = 0;
   s->nfixed = 0;
   s->fixedobj = 0.0;
   for( j = 0; j < nvars; ++j )
   {
      if( isFixedVar(s, lb[j], ub[j]) )
      {
         s->fixedobj += obj[j] * lb[j];
         s->fixedval[s->nfixed] = lb[j];
         s->var2act[j] = -(++s->nfixed);
      }
      else
      {
         s->act2var[s->nactive] = j;
         s->realobj[s->nactive] = obj[j];
         s->actobj[s->nactive] = withobj ? obj[j] : 0.0;
         s->var2act[j] = s->nactive++;
      }
   }
   if( !withobj )
      s->fixedobj = 0.0;
   m = s->nactive + (withr ? 1 : 0);

   /* ---- 2. block and index compression ---- */
   s->ndevblocks = 0;
   for( b = 0; b < nsdpblocks; ++b )
   {
      s->origsize[b] = sdpblocksizes[b];
      if( blockindchanges[b] < 0 )
      {
         s->blk2dev[b] = -1;
         s->devsize[b] = 0;
         continue;
      }
      s->blk2dev[b] = s->ndevblocks++;
      assert( s->blk2dev[b] == b - blockindchanges[b] );
      k = 0;
      for( i = 0; i < sdpblocksizes[b]; ++i )
      {
         if( indchanges[b][i] >= 0 )
         {
            assert( i - indchanges[b][i] == k );
            s->red2orig[b][k++] = i;
         }
      }
      assert( k == sdpblocksizes[b] - nremovedinds[b] );
      s->devsize[b] = k;
   }

   /* ---- 3. constraint-matrix entries grouped by active variable (CSR over variables) ---- */
   MEM_GOTO( BMSallocClearBufferMemoryArray(s->bufmem, &varbeg, m + 2) );
   for( b = 0; b < nsdpblocks; ++b )
   {
      if( s->blk2dev[b] < 0 )
         continue;
      for( v = 0; v < sdpnblockvars[b]; ++v )
      {
         int act = s->var2act[sdpvar[b][v]];
         if( act >= 0 )
            varbeg[act + 1] += sdpnblockvarnonz[b][v];
      }
      if( withr )
         varbeg[s->nactive + 1] += s->devsize[b];
   }
   for( j = 0; j < m; ++j )
      varbeg[j + 1] += varbeg[j];
   nnz = varbeg[m];
   MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &fill, m + 1) );
   MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &entblk, nnz + 1) );
   MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &entrow, nnz + 1) );
   MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &entcol, nnz + 1) );
   MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &entval, nnz + 1) );
   for( j = 0; j < m; ++j )
      fill[j] = varbeg[j];
   for( b = 0; b < nsdpblocks; ++b )
   {
      int db = s->blk2dev[b];
      if( db < 0 )
         continue;
      for( v = 0; v < sdpnblockvars[b]; ++v )
      {
         int act = s->var2act[sdpvar[b][v]];
         if( act < 0 )
            continue;
         for( k = 0; k < sdpnblockvarnonz[b][v]; ++k )
         {
            int r = sdprow[b][v][k];
            int c = sdpcol[b][v][k];
            int p = fill[act]++;
            assert( indchanges[b][r] >= 0 && indchanges[b][c] >= 0 );
            r -= indchanges[b][r];
            c -= indchanges[b][c];
            entblk[p] = db;
            entrow[p] = MAX(r, c);
            entcol[p] = MIN(r, c);
            entval[p] = sdpval[b][v][k];
         }
      }
      if( withr )
      {
         for( i = 0; i < s->devsize[b]; ++i )
         {
            int p = fill[s->nactive]++;
            entblk[p] = db; entrow[p] = i; entcol[p] = i; entval[p] = 1.0;
         }
      }
   }

   /* ---- 4. constant part (may be absent: primal Slater check passes sdpconstnnonz = 0 and NULL arrays, sdpi.c:1661-1789) ---- */
   cnnz = 0;
   MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &cblk, sdpconstnnonz + 1) );
   MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &crow, sdpconstnnonz + 1) );
   MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &ccol, sdpconstnnonz + 1) );
   MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &cval, sdpconstnnonz + 1) );
   if( sdpconstnnonz > 0 )
   {
      for( b = 0; b < nsdpblocks; ++b )
      {
         int db = s->blk2dev[b];
         if( db < 0 )
            continue;
         for( k = 0; k < sdpconstnblocknonz[b]; ++k )
         {
            int r = sdpconstrow[b][k];
            int c = sdpconstcol[b][k];
            assert( indchanges[b][r] >= 0 && indchanges[b][c] >= 0 );
            r -= indchanges[b][r];
            c -= indchanges[b][c];
            assert( cnnz < sdpconstnnonz );
            cblk[cnnz] = db; crow[cnnz] = MAX(r, c); ccol[cnnz] = MIN(r, c); cval[cnnz] = sdpconstval[b][k];
            ++cnnz;
         }
      }
   }
   MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &devblocksizes, s->ndevblocks + 1) );
   for( b = 0; b < nsdpblocks; ++b )
   {
      if( s->blk2dev[b] >= 0 )
         devblocksizes[s->blk2dev[b]] = s->devsize[b];
   }

   /* ---- 5. LP block: row sides, then variable bounds, then r >= 0 ---- */
   lpcap = 2 * lpnnonz + 2 * nlpcons + 2 * s->nactive + 2;
   nrows = 2 * nlpcons + 2 * s->nactive + 1;
   MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &rowbeg, nrows + 1) );
   MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &rowrhs, nrows + 1) );
   MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &rowind, lpcap) );
   MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &rowval, lpcap) );
   nrows = 0;
   lpnz = 0;
   rowbeg[0] = 0;
   for( i = 0; i < nlpcons; ++i )
   {
      int first;
      int last;
      int side;

      s->rowslot[2 * i] = -1;
      s->rowslot[2 * i + 1] = -1;
      if( lpindchanges[i] < 0 )
         continue;
      first = lpbeg[i];
      last = (i == nlpcons - 1) ? lpnnonz : lpbeg[i + 1];    /* lpbeg has no sentinel entry */
      for( side = 0; side < 2; ++side )
      {
         SCIP_Real sgn = (side == 0) ? 1.0 : -1.0;
         SCIP_Real rhs = (side == 0) ? lplhs[i] : lprhs[i];
         if( (side == 0 && lplhs[i] <= -CUDASDP_INF) || (side == 1 && lprhs[i] >= CUDASDP_INF) )
            continue;
         for( k = first; k < last; ++k )
         {
            int act = s->var2act[lpind[k]];
            if( act >= 0 && lpval[k] != 0.0 )
            {
               rowind[lpnz] = act;
               rowval[lpnz] = sgn * lpval[k];
               ++lpnz;
            }
         }
         if( withr )
         {
            rowind[lpnz] = s->nactive;
            rowval[lpnz] = 1.0;
            ++lpnz;
         }
         rowrhs[nrows] = sgn * rhs;
         s->rowslot[2 * i + side] = nrows;
         rowbeg[++nrows] = lpnz;
      }
   }
   s->nrowsides = nrows;
   for( j = 0; j < nvars; ++j )
   {
      int act = s->var2act[j];
      s->boundslot[2 * j] = -1;
      s->boundslot[2 * j + 1] = -1;
      if( act < 0 )
         continue;
      if( !isInf(lb[j]) )
      {
         rowind[lpnz] = act; rowval[lpnz] = 1.0; ++lpnz;
         rowrhs[nrows] = lb[j];
         s->boundslot[2 * j] = nrows;
         rowbeg[++nrows] = lpnz;
      }
      if( !isInf(ub[j]) )
      {
         rowind[lpnz] = act; rowval[lpnz] = -1.0; ++lpnz;
         rowrhs[nrows] = -ub[j];
         s->boundslot[2 * j + 1] = nrows;
         rowbeg[++nrows] = lpnz;
      }
   }
   s->nboundrows = nrows - s->nrowsides;
   if( withr && rbound )
   {
      rowind[lpnz] = s->nactive; rowval[lpnz] = 1.0; ++lpnz;
      rowrhs[nrows] = 0.0;
      rowbeg[++nrows] = lpnz;
   }
   s->nlprows = nrows;
   assert( lpnz <= lpcap );

   /* ---- 6. objective and starting point ---- */
   MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &devobj, m + 1) );
   for( j = 0; j < s->nactive; ++j )
      devobj[j] = s->actobj[j];
   if( withr )
      devobj[s->nactive] = penaltyparam;
   if( starty != NULL && !withr )
   {
      MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &devstart, m + 1) );
      for( j = 0; j < s->nactive; ++j )
         devstart[j] = starty[s->act2var[j]];
   }
   /* full primal-dual start point (sdpisolver_sdpa.cpp:1481-1600): startZ = slack matrix, startX = multiplier matrix, both
    * sparse lower triangles in ORIGINAL indices, last block = LP block with the index convention of sdpisolver.h:171-173 */
   if( starty != NULL && startZnblocknonz != NULL && startXnblocknonz != NULL && !withr && penaltyparam < s->epsilon )
   {
      SCIP_Real* dense = NULL;
      SCIP_Real* lpx = NULL;
      SCIP_Real* lps = NULL;
      int which;
      int maxn = 1;

      for( b = 0; b < nsdpblocks; ++b )
         maxn = MAX(maxn, s->devsize[b]);
      MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &dense, maxn * maxn) );
      MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &lpx, nrows + 1) );
      MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &lps, nrows + 1) );
      for( which = 0; which < 2 && retcode == SCIP_OKAY; ++which )
      {
         const int* nnz = (which == 0) ? startXnblocknonz : startZnblocknonz;
         int* const* rws = (which == 0) ? startXrow : startZrow;
         int* const* cls = (which == 0) ? startXcol : startZcol;
         SCIP_Real* const* vls = (which == 0) ? startXval : startZval;
         SCIP_Real* lp = (which == 0) ? lpx : lps;

         for( b = 0; b < nsdpblocks; ++b )
         {
            int n = s->devsize[b];
            if( s->blk2dev[b] < 0 )
               continue;
            for( i = 0; i < n * n; ++i )
               dense[i] = 0.0;
            for( i = 0; i < nnz[b]; ++i )
            {
               int r = rws[b][i];
               int c = cls[b][i];
               /* the row/column may have been removed in the meantime */
               if( indchanges[b][r] > -1 && indchanges[b][c] > -1 )
               {
                  r -= indchanges[b][r];
                  c -= indchanges[b][c];
                  dense[r * n + c] = vls[b][i];
                  dense[c * n + r] = vls[b][i];
               }
            }
            if( sdpcuda_set_start_block(s->dev, which, s->blk2dev[b], n, dense) != SDPCUDA_OK )
               retcode = SCIP_LPERROR;
         }
         for( i = 0; i < nrows; ++i )
            lp[i] = 0.0;
         for( i = 0; i < nnz[nsdpblocks]; ++i )
         {
            int idx = rws[nsdpblocks][i];
            int slot = -1;
            assert( idx == cls[nsdpblocks][i] );
            if( idx < 2 * nlpcons )
               slot = (idx >= 0 && idx < 2 * s->nlpcons) ? s->rowslot[idx] : -1;
            else if( idx - 2 * nlpcons < 2 * nvars )
               slot = s->boundslot[idx - 2 * nlpcons];
            if( slot >= 0 )
               lp[slot] = vls[nsdpblocks][i];
         }
      }
      if( retcode == SCIP_OKAY && sdpcuda_set_start_lp(s->dev, nrows, lpx, lps) != SDPCUDA_OK )
         retcode = SCIP_LPERROR;
      BMSfreeBufferMemoryArrayNull(s->bufmem, &lps);
      BMSfreeBufferMemoryArrayNull(s->bufmem, &lpx);
      BMSfreeBufferMemoryArrayNull(s->bufmem, &dense);
      if( retcode != SCIP_OKAY )
         goto TERMINATE;
   }
   /* a preoptimal point is only recorded by the first run and only without penalty formulation (sdpisolver_sdpa.cpp:1612) */
   s->wantpreopt = (s->preoptimalgap >= 0.0 && !s->penalty
      && (startsettings == SCIP_SDPSOLVERSETTING_UNSOLVED || startsettings == SCIP_SDPSOLVERSETTING_FAST));

   prob.m = m;
   prob.obj = devobj;
   prob.nblocks = s->ndevblocks;
   prob.blocksizes = devblocksizes;
   prob.varbeg = varbeg;
   prob.entblk = entblk; prob.entrow = entrow; prob.entcol = entcol; prob.entval = entval;
   prob.cnnz = cnnz;
   prob.cblk = cblk; prob.crow = crow; prob.ccol = ccol; prob.cval = cval;
   prob.nlp = nrows;
   prob.lpbeg = rowbeg; prob.lpind = rowind; prob.lpval = rowval; prob.lprhs = rowrhs;

   probhash = structureHash(&prob);

   /* ---- 7. solve: settings ladder with post-check, like sdpisolver_sdpa.cpp:1416-1795 ---- */
   if( s->penalty || startsettings == SCIP_SDPSOLVERSETTING_STABLE || startsettings == SCIP_SDPSOLVERSETTING_PENALTY )
      setting = SCIP_SDPSOLVERSETTING_STABLE;
   else if( startsettings == SCIP_SDPSOLVERSETTING_MEDIUM )
      setting = SCIP_SDPSOLVERSETTING_MEDIUM;
   else if( startsettings == SCIP_SDPSOLVERSETTING_UNSOLVED || startsettings == SCIP_SDPSOLVERSETTING_FAST )
      setting = SCIP_SDPSOLVERSETTING_FAST;
   else
   {
      SCIPerrorMessage("Unknown setting for start-settings: %d!\n", (int) startsettings);
      retcode = SCIP_LPERROR;
      goto TERMINATE;
   }
   lastsetting = s->penalty ? setting : SCIP_SDPSOLVERSETTING_STABLE;

   for( ; setting <= lastsetting; ++setting )
   {
      curgaptol = s->gaptol;
      curfeastol = s->sdpsolverfeastol;

      retcode = runDevice(s, &prob, probhash, curgaptol, curfeastol, setting, timeleft, devstart);
      if( retcode != SCIP_OKAY )
         goto TERMINATE;

      /* post-check of the y-side solution against the caller's tolerances; tighten and repeat on failure */
      while( !s->penalty && s->solved && s->res.phase == SDPCUDA_PDOPT
         && curfeastol >= CUDASDP_MINTOL && curgaptol >= CUDASDP_MINTOL )
      {
         SCIP_Real* solvector;
         SCIP_Bool infeasible;
         SCIP_Bool again = FALSE;

         MEM_GOTO( BMSallocBufferMemoryArray(s->bufmem, &solvector, nvars) );
         retcode = SCIPsdpiSolverGetDualSol(s, NULL, solvector);
         if( retcode == SCIP_OKAY && s->devicecheck )
         {
            int devrc = SDPCUDA_OK;

            /* same contract as SCIPsdpSolcheckerCheck (sdpsolchecker.c:58-270): bounds and rows here (O(nnz)), the blocks as a
             * device Cholesky of Z(y) + feastol I assembled from the problem that is already resident in HBM */
            int psd = 1;

            infeasible = FALSE;
            for( i = 0; i < nvars && !infeasible; ++i )
               infeasible = (solvector[i] < lb[i] - s->feastol || solvector[i] > ub[i] + s->feastol);
            for( i = 0; i < nlpcons && !infeasible; ++i )
            {
               SCIP_Real act = 0.0;
               int last = (i == nlpcons - 1) ? lpnnonz : lpbeg[i + 1];
               int p;

               if( lpindchanges[i] < 0 )
                  continue;
               for( p = lpbeg[i]; p < last; ++p )
               {
                  if( lb[lpind[p]] < ub[lpind[p]] - s->epsilon )
                     act += solvector[lpind[p]] * lpval[p];
               }
               infeasible = (act < lplhs[i] - s->feastol || act > lprhs[i] + s->feastol);
            }
            if( !infeasible && s->ndevblocks > 0 )
            {
               devrc = sdpcuda_check_psd_resident(s->dev, NULL, s->feastol * (1.0 + 1e-6) + 1e-13, &psd);
               if( devrc != SDPCUDA_OK && devrc != SDPCUDA_ERR_STATE )
               {
                  BMSfreeBufferMemoryArray(s->bufmem, &solvector);
                  SCIPerrorMessage("sdpcuda_check_psd_resident failed.\n");
                  retcode = SCIP_LPERROR;
                  goto TERMINATE;
               }
               infeasible = !psd;
            }
            if( devrc == SDPCUDA_ERR_STATE )
            {
               /* the last solve left no problem resident in the form this test needs (packed single solve): ship Z(y) instead */
               retcode = SCIPsdpSolcheckerCheck(s->bufmem, nvars, lb, ub, nsdpblocks, sdpblocksizes, sdpnblockvars, sdpconstnnonz,
                  sdpconstnblocknonz, sdpconstrow, sdpconstcol, sdpconstval, sdpnnonz, sdpnblockvarnonz, sdpvar, sdprow, sdpcol, sdpval,
                  indchanges, nremovedinds, blockindchanges, nlpcons, lpindchanges, lplhs, lprhs, lpnnonz, lpbeg, lpind, lpval,
                  solvector, s->feastol, s->epsilon, &infeasible);
            }
         }
         else if( retcode == SCIP_OKAY )
         {
            retcode = SCIPsdpSolcheckerCheck(s->bufmem, nvars, lb, ub, nsdpblocks, sdpblocksizes, sdpnblockvars, sdpconstnnonz,
               sdpconstnblocknonz, sdpconstrow, sdpconstcol, sdpconstval, sdpnnonz, sdpnblockvarnonz, sdpvar, sdprow, sdpcol, sdpval,
               indchanges, nremovedinds, blockindchanges, nlpcons, lpindchanges, lplhs, lprhs, lpnnonz, lpbeg, lpind, lpval,
               solvector, s->feastol, s->epsilon, &infeasible);
         }
         BMSfreeBufferMemoryArray(s->bufmem, &solvector);
         if( retcode != SCIP_OKAY )
            goto TERMINATE;

         if( infeasible )
         {
            curfeastol *= CUDASDP_TIGHTEN;
            if( curfeastol >= CUDASDP_MINTOL )
               again = TRUE;
         }
         if( REALABS(s->res.dobj - s->res.pobj) >= s->gaptol )
         {
            /* the absolute gap is still open: tighten the gap tolerance by the same factor and solve again, like
             * sdpisolver_sdpa.cpp:449-460 (the device solver also gets the absolute rule itself, params.absgaptol) */
            infeasible = TRUE;
            curgaptol *= CUDASDP_TIGHTEN;
            if( curgaptol >= CUDASDP_MINTOL )
               again = TRUE;
         }
         if( again )
         {
            SCIPdebugMessage("post-check failed, solving again with feastol %g, gaptol %g\n", curfeastol, curgaptol);
            retcode = runDevice(s, &prob, probhash, curgaptol, curfeastol, setting, timeleft, devstart);
            if( retcode != SCIP_OKAY )
               goto TERMINATE;
         }
         else
         {
            if( infeasible )
            {
               s->solved = FALSE;
               SCIPmessagePrintInfo(s->messagehdlr, "CUDA-IPM failed to reach required feasibility tolerance!\n");
            }
            break;
         }
      }

      if( s->penalty )
         s->usedsetting = SCIP_SDPSOLVERSETTING_PENALTY;
      else if( SCIPsdpiSolverIsAcceptable(s) )
      {
         s->usedsetting = (SCIP_SDPSOLVERSETTING) setting;
         break;
      }
      if( s->timelimit )
         break;
   }

   /* ---- 8. penalty formulation: is the solution feasible for the original problem, was the primal bound hit ---- */
   if( s->penalty && s->solved )
   {
      SCIP_Real r = s->y[s->nactive];

      assert( feasorig != NULL );
      *feasorig = (r < s->feastol);
      if( withobj )
         s->feasorig = *feasorig;

      if( !(*feasorig) && penaltybound != NULL )
      {
         SCIP_Real trace = 0.0;

         retcode = fetchXlp(s);
         for( i = 0; i < s->nrowsides && retcode == SCIP_OKAY; ++i )
            trace += s->xlp[i];
         for( b = 0; b < nsdpblocks && retcode == SCIP_OKAY; ++b )
         {
            if( s->blk2dev[b] < 0 )
               continue;
            retcode = fetchX(s, b);
            for( i = 0; i < s->devsize[b] && retcode == SCIP_OKAY; ++i )
               trace += s->X[b][i * s->devsize[b] + i];
         }
         if( retcode != SCIP_OKAY )
            goto TERMINATE;
         *penaltybound = ((penaltyparam - trace) / penaltyparam < CUDASDP_PENALTYBOUNDTOL);
      }
      else if( penaltybound != NULL )
         *penaltybound = FALSE;
   }

TERMINATE:
   BMSfreeBufferMemoryArrayNull(s->bufmem, &devstart);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &devobj);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &rowval);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &rowind);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &rowrhs);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &rowbeg);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &devblocksizes);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &cval);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &ccol);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &crow);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &cblk);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &entval);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &entcol);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &entrow);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &entblk);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &fill);
   BMSfreeBufferMemoryArrayNull(s->bufmem, &varbeg);

   return retcode;
}

/*
 * Solution Information Methods
 */

SCIP_Bool SCIPsdpiSolverWasSolved(SCIP_SDPISOLVER* sdpisolver)
{
   assert( sdpisolver != NULL );
   return sdpisolver->solved;
}

SCIP_Bool SCIPsdpiSolverFeasibilityKnown(SCIP_SDPISOLVER* sdpisolver)
{
   int ph;
   assert( sdpisolver != NULL );
   NEED_SOLVED_BOOL( sdpisolver );
   ph = sdpisolver->res.phase;
   return !(ph == SDPCUDA_NOINFO || ph == SDPCUDA_PFEAS || ph == SDPCUDA_DFEAS || ph == SDPCUDA_PDINF);
}

SCIP_RETCODE SCIPsdpiSolverGetSolFeasibility(SCIP_SDPISOLVER* sdpisolver, SCIP_Bool* primalfeasible, SCIP_Bool* dualfeasible)
{
   assert( sdpisolver != NULL );
   assert( primalfeasible != NULL && dualfeasible != NULL );
   NEED_SOLVED( sdpisolver );

   switch( sdpisolver->res.phase )
   {
   case SDPCUDA_PDOPT:
   case SDPCUDA_PDFEAS:
      *primalfeasible = TRUE; *dualfeasible = TRUE;
      break;
   case SDPCUDA_PFEAS_DINF:
   case SDPCUDA_PUNBD:
      *primalfeasible = TRUE; *dualfeasible = FALSE;
      break;
   case SDPCUDA_PINF_DFEAS:
   case SDPCUDA_DUNBD:
      *primalfeasible = FALSE; *dualfeasible = TRUE;
      break;
   case SDPCUDA_DINF:
      *primalfeasible = FALSE; *dualfeasible = FALSE;   /* y-problem infeasible, nothing proven about the X-side */
      break;
   default:
      SCIPerrorMessage("CUDA-IPM doesn't know if primal and dual solutions are feasible\n");
      return SCIP_LPERROR;
   }
   return SCIP_OKAY;
}

SCIP_Bool SCIPsdpiSolverIsPrimalUnbounded(SCIP_SDPISOLVER* sdpisolver)
{
   assert( sdpisolver != NULL );
   NEED_SOLVED_BOOL( sdpisolver );
   return sdpisolver->res.phase == SDPCUDA_PFEAS_DINF || sdpisolver->res.phase == SDPCUDA_PUNBD;
}

SCIP_Bool SCIPsdpiSolverIsPrimalInfeasible(SCIP_SDPISOLVER* sdpisolver)
{
   assert( sdpisolver != NULL );
   NEED_SOLVED_BOOL( sdpisolver );
   return sdpisolver->res.phase == SDPCUDA_PINF_DFEAS || sdpisolver->res.phase == SDPCUDA_DUNBD;
}

SCIP_Bool SCIPsdpiSolverIsPrimalFeasible(SCIP_SDPISOLVER* sdpisolver)
{
   int ph;
   assert( sdpisolver != NULL );
   NEED_SOLVED_BOOL( sdpisolver );
   ph = sdpisolver->res.phase;
   return ph == SDPCUDA_PFEAS_DINF || ph == SDPCUDA_PDOPT || ph == SDPCUDA_PFEAS || ph == SDPCUDA_PDFEAS || ph == SDPCUDA_PUNBD;
}

SCIP_Bool SCIPsdpiSolverIsDualUnbounded(SCIP_SDPISOLVER* sdpisolver)
{
   assert( sdpisolver != NULL );
   NEED_SOLVED_BOOL( sdpisolver );
   return sdpisolver->res.phase == SDPCUDA_PINF_DFEAS || sdpisolver->res.phase == SDPCUDA_DUNBD;
}

SCIP_Bool SCIPsdpiSolverIsDualInfeasible(SCIP_SDPISOLVER* sdpisolver)
{
   assert( sdpisolver != NULL );
   NEED_SOLVED_BOOL( sdpisolver );
   return sdpisolver->res.phase == SDPCUDA_PFEAS_DINF || sdpisolver->res.phase == SDPCUDA_PUNBD || sdpisolver->res.phase == SDPCUDA_DINF;
}

SCIP_Bool SCIPsdpiSolverIsDualFeasible(SCIP_SDPISOLVER* sdpisolver)
{
   int ph;
   assert( sdpisolver != NULL );
   NEED_SOLVED_BOOL( sdpisolver );
   ph = sdpisolver->res.phase;
   return ph == SDPCUDA_PINF_DFEAS || ph == SDPCUDA_PDOPT || ph == SDPCUDA_DFEAS || ph == SDPCUDA_PDFEAS || ph == SDPCUDA_DUNBD;
}

SCIP_Bool SCIPsdpiSolverIsConverged(SCIP_SDPISOLVER* sdpisolver)
{
   assert( sdpisolver != NULL );
   NEED_SOLVED_BOOL( sdpisolver );
   return sdpisolver->res.phase == SDPCUDA_PDOPT;
}

SCIP_Bool SCIPsdpiSolverIsObjlimExc(SCIP_SDPISOLVER* sdpisolver)
{
   assert( sdpisolver != NULL );
   NEED_SOLVED_BOOL( sdpisolver );
   return sdpisolver->res.phase == SDPCUDA_PUNBD;
}

SCIP_Bool SCIPsdpiSolverIsIterlimExc(SCIP_SDPISOLVER* sdpisolver)
{
   assert( sdpisolver != NULL );
   NEED_SOLVED_BOOL( sdpisolver );
   return sdpisolver->res.stop == SDPCUDA_STOP_ITERLIMIT;
}

SCIP_Bool SCIPsdpiSolverIsTimelimExc(SCIP_SDPISOLVER* sdpisolver)
{
   assert( sdpisolver != NULL );
   return sdpisolver->timelimit;
}

/** -1 not started, 0 converged, 1 infeasible start, 2 numerical problems, 3 objective limit, 4 iteration limit,
 *  5 time limit, 6 user termination, 7 other (sdpisolver.h:439-450) */
int SCIPsdpiSolverGetInternalStatus(SCIP_SDPISOLVER* sdpisolver)
{
   assert( sdpisolver != NULL );
   if( !sdpisolver->solved )
      return -1;
   switch( sdpisolver->res.stop )
   {
   case SDPCUDA_STOP_CONVERGED:
   case SDPCUDA_STOP_INFEASCERT:
      return 0;
   case SDPCUDA_STOP_NUMERICS:
      return 2;
   case SDPCUDA_STOP_OBJLIMIT:
      return 3;
   case SDPCUDA_STOP_ITERLIMIT:
      return 4;
   case SDPCUDA_STOP_TIMELIMIT:
      return 5;
   default:
      return 7;
   }
}

SCIP_Bool SCIPsdpiSolverIsOptimal(SCIP_SDPISOLVER* sdpisolver)
{
   assert( sdpisolver != NULL );
   if( !sdpisolver->solved )
      return FALSE;
   return sdpisolver->res.phase == SDPCUDA_PDOPT;
}

SCIP_Bool SCIPsdpiSolverIsAcceptable(SCIP_SDPISOLVER* sdpisolver)
{
   int ph;
   assert( sdpisolver != NULL );
   if( sdpisolver->timelimit || !sdpisolver->solved )
      return FALSE;
   ph = sdpisolver->res.phase;
   return ph == SDPCUDA_PDOPT || ph == SDPCUDA_PUNBD || ph == SDPCUDA_PINF_DFEAS || ph == SDPCUDA_PFEAS_DINF || ph == SDPCUDA_DINF;
}

SCIP_RETCODE SCIPsdpiSolverIgnoreInstability(SCIP_SDPISOLVER* sdpisolver, SCIP_Bool* success)
{
   (void) sdpisolver;
   if( success != NULL )
      *success = FALSE;
   SCIPdebugMessage("Not implemented yet\n");
   return SCIP_LPERROR;
}

SCIP_RETCODE SCIPsdpiSolverGetObjval(SCIP_SDPISOLVER* sdpisolver, SCIP_Real* objval)
{
   SCIP_SDPISOLVER* s = sdpisolver;
   int j;

   assert( s != NULL );
   assert( objval != NULL );
   NEED_SOLVED( s );

   if( s->penalty && !s->feasorig )
   {
      /* objective of the penalty formulation itself (includes Gamma * r) */
      *objval = s->res.dobj;
   }
   else
   {
      /* recomputed from y, which is more accurate than the solver's running value (sdpisolver_sdpa.cpp:2365-2373) */
      *objval = 0.0;
      for( j = 0; j < s->nactive; ++j )
         *objval += s->y[j] * s->actobj[j];
   }
   *objval += s->fixedobj;
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverGetDualSol(SCIP_SDPISOLVER* sdpisolver, SCIP_Real* objval, SCIP_Real* dualsol)
{
   SCIP_SDPISOLVER* s = sdpisolver;
   int j;

   assert( s != NULL );
   NEED_SOLVED( s );

   if( objval != NULL )
   {
      SCIP_CALL( SCIPsdpiSolverGetObjval(s, objval) );
   }
   if( dualsol != NULL )
   {
      for( j = 0; j < s->nvars; ++j )
      {
         int act = s->var2act[j];
         dualsol[j] = (act >= 0) ? s->y[act] : s->fixedval[-act - 1];
      }
   }
   return SCIP_OKAY;
}

static SCIP_RETCODE sparsePrimal(SCIP_SDPISOLVER* s, SCIP_Bool preopt, int nblocks, int* nnonz, int** rows, int** cols, SCIP_Real** vals, SCIP_Bool* toosmall);

SCIP_RETCODE SCIPsdpiSolverGetPreoptimalPrimalNonzeros(SCIP_SDPISOLVER* sdpisolver, int nblocks, int* startXnblocknonz)
{
   SCIP_Bool toosmall;

   assert( sdpisolver != NULL );
   assert( nblocks > 0 );
   assert( startXnblocknonz != NULL );
   NEED_SOLVED( sdpisolver );

   /* no preoptimal point: signalled by -1 in the first entry (sdpisolver.h:492-499, sdpisolver_sdpa.cpp:2446-2451) */
   if( !sdpisolver->preoptexists )
   {
      startXnblocknonz[0] = -1;
      return SCIP_OKAY;
   }
   return sparsePrimal(sdpisolver, TRUE, nblocks, startXnblocknonz, NULL, NULL, NULL, &toosmall);
}

SCIP_RETCODE SCIPsdpiSolverGetPreoptimalSol(SCIP_SDPISOLVER* sdpisolver, SCIP_Bool* success, SCIP_Real* dualsol, int nblocks,
   int* startXnblocknonz, int** startXrow, int** startXcol, SCIP_Real** startXval)
{
   SCIP_SDPISOLVER* s = sdpisolver;
   SCIP_Bool toosmall;
   int j;

   assert( s != NULL );
   assert( success != NULL );
   assert( startXnblocknonz != NULL || nblocks == -1 );

   /* like sdpisolver_sdpa.cpp:2531-2536: no error, only *success = FALSE */
   if( !s->solved || !s->preoptexists )
   {
      *success = FALSE;
      if( startXnblocknonz != NULL && nblocks > 0 )
         startXnblocknonz[0] = -1;
      return SCIP_OKAY;
   }
   if( dualsol != NULL )
   {
      for( j = 0; j < s->nvars; ++j )
      {
         int act = s->var2act[j];
         dualsol[j] = (act >= 0) ? s->preopty[act] : s->fixedval[-act - 1];
      }
   }
   if( nblocks > -1 )
   {
      assert( startXrow != NULL && startXcol != NULL && startXval != NULL );
      SCIP_CALL( sparsePrimal(s, TRUE, nblocks, startXnblocknonz, startXrow, startXcol, startXval, &toosmall) );
      if( toosmall )
      {
         *success = FALSE;
         return SCIP_OKAY;
      }
   }
   *success = TRUE;
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverGetPrimalBoundVars(SCIP_SDPISOLVER* sdpisolver, SCIP_Real* lbvals, SCIP_Real* ubvals)
{
   SCIP_SDPISOLVER* s = sdpisolver;
   int j;

   assert( s != NULL );
   assert( lbvals != NULL && ubvals != NULL );
   NEED_SOLVED( s );

   SCIP_CALL( fetchXlp(s) );
   for( j = 0; j < s->nvars; ++j )
   {
      lbvals[j] = s->boundslot[2 * j] >= 0 ? s->xlp[s->boundslot[2 * j]] : 0.0;
      ubvals[j] = s->boundslot[2 * j + 1] >= 0 ? s->xlp[s->boundslot[2 * j + 1]] : 0.0;
   }
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverGetPrimalLPSides(SCIP_SDPISOLVER* sdpisolver, int nlpcons, int* lpindchanges, SCIP_Real* lplhs,
   SCIP_Real* lprhs, SCIP_Real* lhsvals, SCIP_Real* rhsvals)
{
   SCIP_SDPISOLVER* s = sdpisolver;
   int i;

   (void) lpindchanges; (void) lplhs; (void) lprhs;
   assert( s != NULL );
   NEED_SOLVED( s );
   if( nlpcons <= 0 )
      return SCIP_OKAY;
   assert( lhsvals != NULL && rhsvals != NULL );

   SCIP_CALL( fetchXlp(s) );
   for( i = 0; i < nlpcons; ++i )
   {
      lhsvals[i] = (i < s->nlpcons && s->rowslot[2 * i] >= 0) ? s->xlp[s->rowslot[2 * i]] : 0.0;
      rhsvals[i] = (i < s->nlpcons && s->rowslot[2 * i + 1] >= 0) ? s->xlp[s->rowslot[2 * i + 1]] : 0.0;
   }
   return SCIP_OKAY;
}

/** walks over the multiplier matrix X in sparse form; with rows == NULL only counts (shared by the two getters below) */
static SCIP_RETCODE sparsePrimal(SCIP_SDPISOLVER* s, SCIP_Bool preopt, int nblocks, int* nnonz, int** rows, int** cols, SCIP_Real** vals, SCIP_Bool* toosmall)
{
   SCIP_Real** Xsrc = preopt ? s->preoptX : s->X;
   SCIP_Real* xlpsrc;
   int b;
   int i;
   int j;
   int cnt;

   if( nblocks != s->nsdpblocks + 1 )
   {
      SCIPerrorMessage("expected nblocks = %d but got %d\n", s->nsdpblocks + 1, nblocks);
      return SCIP_LPERROR;
   }
   *toosmall = FALSE;
   for( b = 0; b < s->nsdpblocks; ++b )
   {
      int cap = nnonz[b];
      int n = s->devsize[b];

      cnt = 0;
      if( s->blk2dev[b] >= 0 )
      {
         if( !preopt )
            SCIP_CALL( fetchX(s, b) );
         for( i = 0; i < n; ++i )
         {
            for( j = 0; j <= i; ++j )
            {
               SCIP_Real val = Xsrc[b][i * n + j];
               if( REALABS(val) > s->epsilon )
               {
                  if( rows != NULL && cnt < cap )
                  {
                     rows[b][cnt] = s->red2orig[b][i];
                     cols[b][cnt] = s->red2orig[b][j];
                     vals[b][cnt] = val;
                  }
                  ++cnt;
               }
            }
         }
      }
      if( rows != NULL && cnt > cap )
         *toosmall = TRUE;
      nnonz[b] = cnt;
   }
   /* LP block: index 2i / 2i+1 for lhs / rhs of input row i, 2*nlpcons + 2j (+1) for lb (ub) of input variable j */
   {
      int cap = nnonz[nblocks - 1];
      cnt = 0;
      if( !preopt )
         SCIP_CALL( fetchXlp(s) );
      xlpsrc = preopt ? s->preoptxlp : s->xlp;
      for( i = 0; i < 2 * s->nlpcons; ++i )
      {
         if( s->rowslot[i] >= 0 && REALABS(xlpsrc[s->rowslot[i]]) > s->epsilon )
         {
            if( rows != NULL && cnt < cap )
            {
               rows[nblocks - 1][cnt] = i; cols[nblocks - 1][cnt] = i; vals[nblocks - 1][cnt] = xlpsrc[s->rowslot[i]];
            }
            ++cnt;
         }
      }
      for( i = 0; i < 2 * s->nvars; ++i )
      {
         if( s->boundslot[i] >= 0 && REALABS(xlpsrc[s->boundslot[i]]) > s->epsilon )
         {
            if( rows != NULL && cnt < cap )
            {
               rows[nblocks - 1][cnt] = 2 * s->nlpcons + i; cols[nblocks - 1][cnt] = 2 * s->nlpcons + i;
               vals[nblocks - 1][cnt] = xlpsrc[s->boundslot[i]];
            }
            ++cnt;
         }
      }
      if( rows != NULL && cnt > cap )
         *toosmall = TRUE;
      nnonz[nblocks - 1] = cnt;
   }
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverGetPrimalNonzeros(SCIP_SDPISOLVER* sdpisolver, int nblocks, int* startXnblocknonz)
{
   SCIP_Bool toosmall;
   assert( sdpisolver != NULL );
   assert( startXnblocknonz != NULL );
   NEED_SOLVED( sdpisolver );
   return sparsePrimal(sdpisolver, FALSE, nblocks, startXnblocknonz, NULL, NULL, NULL, &toosmall);
}

SCIP_RETCODE SCIPsdpiSolverGetPrimalMatrix(SCIP_SDPISOLVER* sdpisolver, int nblocks, int* startXnblocknonz, int** startXrow,
   int** startXcol, SCIP_Real** startXval)
{
   SCIP_Bool toosmall;
   assert( sdpisolver != NULL );
   assert( startXnblocknonz != NULL && startXrow != NULL && startXcol != NULL && startXval != NULL );
   NEED_SOLVED( sdpisolver );
   SCIP_CALL( sparsePrimal(sdpisolver, FALSE, nblocks, startXnblocknonz, startXrow, startXcol, startXval, &toosmall) );
   if( toosmall )
      SCIPdebugMessage("Insufficient memory for the primal matrix, needed sizes are returned in startXnblocknonz.\n");
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverGetPrimalSolutionMatrix(SCIP_SDPISOLVER* sdpisolver, int nsdpblocks, int* sdpblocksizes,
   int** indchanges, int* nremovedinds, int* blockindchanges, SCIP_Real** primalmatrices)
{
   SCIP_SDPISOLVER* s = sdpisolver;
   int b;
   int i;
   int j;

   (void) indchanges; (void) nremovedinds; (void) blockindchanges;
   assert( s != NULL );
   assert( nsdpblocks == 0 || (sdpblocksizes != NULL && primalmatrices != NULL) );
   NEED_SOLVED( s );
   if( nsdpblocks != s->nsdpblocks )
   {
      SCIPerrorMessage("expected nsdpblocks = %d but got %d\n", s->nsdpblocks, nsdpblocks);
      return SCIP_LPERROR;
   }

   /* dense, row-major, ORIGINAL block size, zeros in removed rows/columns (sdpisolver_sdpa.cpp:3050-3083) */
   for( b = 0; b < nsdpblocks; ++b )
   {
      int norig = sdpblocksizes[b];
      int n = s->devsize[b];

      assert( norig == s->origsize[b] );
      for( i = 0; i < norig * norig; ++i )
         primalmatrices[b][i] = 0.0;
      if( s->blk2dev[b] < 0 )
         continue;
      SCIP_CALL( fetchX(s, b) );
      for( i = 0; i < n; ++i )
         for( j = 0; j < n; ++j )
            primalmatrices[b][s->red2orig[b][i] * norig + s->red2orig[b][j]] = s->X[b][i * n + j];
   }
   return SCIP_OKAY;
}

SCIP_Real SCIPsdpiSolverGetMaxPrimalEntry(SCIP_SDPISOLVER* sdpisolver)
{
   SCIP_SDPISOLVER* s = sdpisolver;
   SCIP_Real maxentry = 0.0;
   int b;
   int i;

   assert( s != NULL );
   if( !s->solved )
      return 0.0;
   for( b = 0; b < s->nsdpblocks; ++b )
   {
      if( s->blk2dev[b] < 0 || fetchX(s, b) != SCIP_OKAY )
         continue;
      for( i = 0; i < s->devsize[b] * s->devsize[b]; ++i )
         maxentry = MAX(maxentry, REALABS(s->X[b][i]));
   }
   if( fetchXlp(s) == SCIP_OKAY )
   {
      for( i = 0; i < s->nlprows; ++i )
         maxentry = MAX(maxentry, REALABS(s->xlp[i]));
   }
   return maxentry;
}

SCIP_RETCODE SCIPsdpiSolverGetTime(SCIP_SDPISOLVER* sdpisolver, SCIP_Real* opttime)
{
   assert( sdpisolver != NULL );
   assert( opttime != NULL );
   *opttime = sdpisolver->opttime;
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverGetIterations(SCIP_SDPISOLVER* sdpisolver, int* iterations)
{
   assert( sdpisolver != NULL );
   assert( iterations != NULL );
   *iterations = sdpisolver->niterations;
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverGetSdpCalls(SCIP_SDPISOLVER* sdpisolver, int* calls)
{
   assert( sdpisolver != NULL );
   assert( calls != NULL );
   *calls = sdpisolver->nsdpcalls;
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverSettingsUsed(SCIP_SDPISOLVER* sdpisolver, SCIP_SDPSOLVERSETTING* usedsetting)
{
   assert( sdpisolver != NULL );
   assert( usedsetting != NULL );
   NEED_SOLVED( sdpisolver );
   *usedsetting = sdpisolver->usedsetting;
   return SCIP_OKAY;
}

/*
 * Numerical Methods
 */

SCIP_Real SCIPsdpiSolverInfinity(SCIP_SDPISOLVER* sdpisolver)
{
   (void) sdpisolver;
   return CUDASDP_INF;
}

SCIP_Bool SCIPsdpiSolverIsInfinity(SCIP_SDPISOLVER* sdpisolver, SCIP_Real val)
{
   (void) sdpisolver;
   return isInf(val);
}

SCIP_RETCODE SCIPsdpiSolverGetRealpar(SCIP_SDPISOLVER* sdpisolver, SCIP_SDPPARAM type, SCIP_Real* dval)
{
   assert( sdpisolver != NULL );
   assert( dval != NULL );
   switch( type )
   {
   case SCIP_SDPPAR_EPSILON:          *dval = sdpisolver->epsilon; break;
   case SCIP_SDPPAR_GAPTOL:           *dval = sdpisolver->gaptol; break;
   case SCIP_SDPPAR_FEASTOL:          *dval = sdpisolver->feastol; break;
   case SCIP_SDPPAR_SDPSOLVERFEASTOL: *dval = sdpisolver->sdpsolverfeastol; break;
   case SCIP_SDPPAR_PENALTYPARAM:     *dval = 0.0; break;
   case SCIP_SDPPAR_OBJLIMIT:         *dval = sdpisolver->objlimit; break;
   case SCIP_SDPPAR_LAMBDASTAR:       *dval = sdpisolver->lambdastar; break;
   case SCIP_SDPPAR_WARMSTARTPOGAP:   *dval = sdpisolver->preoptimalgap; break;
   default:
      return SCIP_PARAMETERUNKNOWN;
   }
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverSetRealpar(SCIP_SDPISOLVER* sdpisolver, SCIP_SDPPARAM type, SCIP_Real dval)
{
   assert( sdpisolver != NULL );
   switch( type )
   {
   case SCIP_SDPPAR_EPSILON:          sdpisolver->epsilon = dval; break;
   case SCIP_SDPPAR_GAPTOL:           sdpisolver->gaptol = dval; break;
   case SCIP_SDPPAR_FEASTOL:          sdpisolver->feastol = dval; break;
   case SCIP_SDPPAR_SDPSOLVERFEASTOL: sdpisolver->sdpsolverfeastol = dval; break;
   case SCIP_SDPPAR_PENALTYPARAM:     break;   /* handed over with every penalty solve */
   case SCIP_SDPPAR_OBJLIMIT:         sdpisolver->objlimit = dval; break;
   case SCIP_SDPPAR_LAMBDASTAR:       sdpisolver->lambdastar = dval; break;
   case SCIP_SDPPAR_WARMSTARTPOGAP:   sdpisolver->preoptimalgap = dval; break;
   default:
      return SCIP_PARAMETERUNKNOWN;
   }
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverGetIntpar(SCIP_SDPISOLVER* sdpisolver, SCIP_SDPPARAM type, int* ival)
{
   assert( sdpisolver != NULL );
   assert( ival != NULL );
   switch( type )
   {
   case SCIP_SDPPAR_SDPINFO:  *ival = (int) sdpisolver->sdpinfo; break;
   case SCIP_SDPPAR_NTHREADS: *ival = sdpisolver->nthreads; break;
   default:
      return SCIP_PARAMETERUNKNOWN;
   }
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverSetIntpar(SCIP_SDPISOLVER* sdpisolver, SCIP_SDPPARAM type, int ival)
{
   assert( sdpisolver != NULL );
   switch( type )
   {
   case SCIP_SDPPAR_SDPINFO:  sdpisolver->sdpinfo = (SCIP_Bool) ival; break;
   case SCIP_SDPPAR_NTHREADS: sdpisolver->nthreads = ival; break;   /* host threads are irrelevant for the device solver */
   default:
      return SCIP_PARAMETERUNKNOWN;
   }
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverComputeLambdastar(SCIP_SDPISOLVER* sdpisolver, SCIP_Real maxguess)
{
   /* the starting point is scaled from the problem data on the device; the guess is only remembered */
   assert( sdpisolver != NULL );
   sdpisolver->lambdastar = maxguess;
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverComputePenaltyparam(SCIP_SDPISOLVER* sdpisolver, SCIP_Real maxcoeff, SCIP_Real* penaltyparam)
{
   SCIP_Real val = CUDASDP_PENALTYPARAM_FACTOR * maxcoeff;
   (void) sdpisolver;
   assert( penaltyparam != NULL );
   *penaltyparam = MAX(CUDASDP_MIN_PENALTYPARAM, MIN(CUDASDP_MAX_PENALTYPARAM, val));
   return SCIP_OKAY;
}

SCIP_RETCODE SCIPsdpiSolverComputeMaxPenaltyparam(SCIP_SDPISOLVER* sdpisolver, SCIP_Real penaltyparam, SCIP_Real* maxpenaltyparam)
{
   (void) sdpisolver;
   assert( maxpenaltyparam != NULL );
   *maxpenaltyparam = MIN(CUDASDP_MAXPENALTY_FACTOR * penaltyparam, CUDASDP_MAX_MAXPENALTYPARAM);
   return SCIP_OKAY;
}

/*
 * File Interface Methods
 */

SCIP_RETCODE SCIPsdpiSolverReadSDP(SCIP_SDPISOLVER* sdpisolver, const char* fname)
{
   (void) sdpisolver; (void) fname;
   SCIPdebugMessage("Not implemented yet\n");
   return SCIP_LPERROR;
}

SCIP_RETCODE SCIPsdpiSolverWriteSDP(SCIP_SDPISOLVER* sdpisolver, const char* fname)
{
   (void) sdpisolver; (void) fname;
   SCIPdebugMessage("Not implemented yet\n");
   return SCIP_LPERROR;
}
