// ipm_small.cuh — single-launch interior-point solver for small relaxations (the B&B node regime of SCIP-SDP's shipped
// instances: blocks of order <= 64, <= 256 variables).  See ipm_small.cu.
#pragma once
#include "common.cuh"
#include "ops.cuh"

namespace sdpk {

constexpr int SMALL_MAX_N = 64;        // largest SDP block
constexpr int SMALL_MAX_M = 256;       // largest Schur complement
constexpr int TINY_MAX_N = 16;         // blocks of the second instantiation (ipm_tiny.cu): four relaxations per SM
constexpr int SMALL_MAX_BLOCKS = 16;
constexpr int SMALL_MAX_GROUPS = 8;
constexpr int SMALL_LZ_STEPS = 32;     // Lanczos steps per step-length matrix (exact when the block order is smaller)
// Schur complements of order 65 .. SMALL_MPK / TINY_MPK: the factor lives in shared memory as a packed lower triangle behind the
// kernel's own buffers (and behind a staged work space), if the launch reserves the bytes (SmallArgs::mpk_off > 0); example_MkP
// (m = 105) spent 57 % of its cycles in the substitutions and the factorisation of a factor that lived in global memory
constexpr int SMALL_MPK = 128, TINY_MPK = 112;
constexpr size_t small_mpk_bytes(int mmax) { return sizeof(double) * ((size_t)mmax * (mmax + 1) / 2 + 2); }

struct SmallBlock { int n, ld; long long off; long long lzoff; };

struct SmallResult
{
   int phase, stop, iterations, backtracks;
   double pobj, dobj, relgap, pinf, dinf, mu;
};

struct SmallArgs
{
   int m, nb, nlp, N, ldm, npos, cnnz, ndense, ngroups, maxiter, setting, verbose;
   long long arena;
   SmallBlock blk[SMALL_MAX_BLOCKS];
   DevEntries E;
   const int* cls;
   const int* posbeg; const long long* pos; const long long* mirror; const int* posvar; const double* posval; const double* posc;
   const long long* cpos; const long long* cmirror; const double* cval;
   const int* lpbeg; const int* lpind; const double* lpval; const double* lprhs;
   const int* colbeg; const int* colrow; const double* colval;
   const double* b;
   const int* denselist; const double* Adense;
   int gblk[SMALL_MAX_GROUPS], gfirst[SMALL_MAX_GROUPS], gcount[SMALL_MAX_GROUPS];
   long long gaoff[SMALL_MAX_GROUPS];
   double *X, *S, *Sinv, *L, *Linv, *LX, *LXinv, *dX, *dS, *dXa, *dSa, *K, *T1, *T2, *Rd;
   double *y, *dy, *g, *rp, *AX, *DTx, *tm1, *tm2;
   double *x, *s, *dx, *ds, *dxa, *dsa, *klp, *rdlp, *Dy, *Ddy;
   double *M, *Mfac, *Hd, *Ud, *lz;
   double gaptol, feastol, absgaptol, objlimit, normb, normC, normCsdp2, gammabase;
   SmallResult* out;
   // frontier batch (sdpcuda_solve_batch): the CTA sets up its own cold start X = xi I, S = eta I, x = xil, s = etal, y = 0 and
   // expands its dense constraint matrices into Adense itself, so that a node costs the host no launch and no copy of its own
   int selfinit;
   // frontier batch with SDPCUDA_BATCH_SMEM=1: the first stage_doubles doubles of the node's work space (starting at workbase:
   // vectors first, then the block matrices, then M and its factor - as many whole arrays as the launch has shared memory for) live in
   // the CTA's shared memory behind the kernel's own buffers; the entry kernel redirects the pointers, the body does not notice
   long long stage_doubles;
   double* workbase;
   int copyback;            // staged X, S, x, s are copied to their global addresses when the solve ends (a single packed solve whose
                            // multipliers the getters serve afterwards; a frontier batch only returns y and leaves this 0)
   long long adense_total;
   long long mpk_off;       // doubles from the start of dynamic shared memory to the packed Schur factor (0: not reserved)
   double xil, etal;
   double xi[SMALL_MAX_BLOCKS], eta[SMALL_MAX_BLOCKS];
};

cudaError_t launch_ipm_small(cudaStream_t st, const SmallArgs& a);
// one launch for a whole frontier of small relaxations: CTA i solves dev_args[i] (device array of `count` descriptors)
// (stage_bytes: everything the launch reserves behind the kernel's own buffers - staged work space and packed Schur factor)
cudaError_t launch_ipm_small_batch(cudaStream_t st, int count, const SmallArgs* dev_args, size_t stage_bytes = 0);
// the same for relaxations whose blocks all have order <= TINY_MAX_N: CTAs of 256 threads, four per SM (ipm_tiny.cu)
cudaError_t launch_ipm_tiny_batch(cudaStream_t st, int count, const SmallArgs* dev_args, size_t stage_bytes = 0);
// shared memory the two batch kernels use for themselves (the staged work space of a node comes on top)
size_t ipm_small_smem_bytes();
size_t ipm_small_msh_offset_bytes();      // where the 64 x 65 tile of the Schur factor (m <= 64) starts: the last of the kernel's own buffers
size_t ipm_tiny_msh_offset_bytes();
size_t ipm_tiny_smem_bytes();

} // namespace sdpk
