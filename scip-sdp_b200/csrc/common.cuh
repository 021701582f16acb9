// common.cuh — shared declarations of the libsdpcuda device library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

namespace sdpk {

// FP64 tensor-core instruction of sm_100a: D(8x8) += A(8x4) B(4x8); lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2(l%4)+{0,1}]
#ifdef __CUDACC__
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b)
{
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#endif


// every launch goes through this counter so that sdpcuda_result.launches / bench.py's gpu_launches are real counts
struct LaunchCounter { long long n = 0; };
extern thread_local LaunchCounter* g_counter;
inline void count_launch(int k = 1) { if( g_counter ) g_counter->n += k; }

// optional per-class device timing (CUDA events on the launching stream around each launch / launch group)
constexpr int NPROF = 7;
enum { PROF_GEMM = 0, PROF_DIAG = 1, PROF_SCHUR = 2, PROF_EIG = 3, PROF_TRSV = 4, PROF_ELEM = 5, PROF_GEMM_SMALL = 6 };
struct Profiler
{
   bool on = false;
   struct Rec { int cls; cudaEvent_t a, b; double work; int launches; };
   std::vector<Rec> recs;
   std::vector<cudaEvent_t> pool;
   size_t used = 0;
   double launches[NPROF] = {0}, ms[NPROF] = {0}, work[NPROF] = {0};
   cudaEvent_t get()
   {
      if( used == pool.size() ) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
      return pool[used++];
   }
   void reset() { recs.clear(); used = 0; for( int c = 0; c < NPROF; ++c ) launches[c] = ms[c] = work[c] = 0.0; }
   void collect()      // call after the stream has been synchronised
   {
      for( const Rec& r : recs )
      {
         float t = 0.f;
         cudaEventElapsedTime(&t, r.a, r.b);
         launches[r.cls] += r.launches; ms[r.cls] += t; work[r.cls] += r.work;
      }
      recs.clear(); used = 0;
   }
   ~Profiler() { for( cudaEvent_t e : pool ) cudaEventDestroy(e); }
};
extern thread_local Profiler* g_prof;
struct ProfScope
{
   Profiler* p; cudaStream_t st; size_t idx; long long n0;
   ProfScope(cudaStream_t s, int cls, double work) : p((g_prof && g_prof->on) ? g_prof : nullptr), st(s), idx(0), n0(0)
   {
      if( !p ) return;
      Profiler::Rec r; r.cls = cls; r.a = p->get(); r.b = p->get(); r.work = work; r.launches = 0;
      n0 = g_counter ? g_counter->n : 0;
      cudaEventRecord(r.a, st);
      idx = p->recs.size();
      p->recs.push_back(r);
   }
   ~ProfScope()
   {
      if( !p ) return;
      cudaEventRecord(p->recs[idx].b, st);
      p->recs[idx].launches = (int)((g_counter ? g_counter->n : 0) - n0);
   }
};

#define SDPK_CUDA_CHECK(expr) do { cudaError_t _e = (expr); if( _e != cudaSuccess ) { \
      fprintf(stderr, "[libsdpcuda] %s:%d CUDA error %s: %s\n", __FILE__, __LINE__, cudaGetErrorName(_e), cudaGetErrorString(_e)); \
      return _e; } } while( 0 )

inline int round_up(int x, int a) { return (x + a - 1) / a * a; }
inline int ceil_div(int x, int a) { return (x + a - 1) / a; }

// ---- gemm.cu -------------------------------------------------------------------------------------------------------
// C(m x n) = alpha * op(A) * op(B) + beta * C, column-major, FP64 DMMA tensor-core tiles fed by cp.async.
// flags: GEMM_LOWER computes only tiles that touch the lower triangle (row >= col) of C.
//        GEMM_KHI_M / GEMM_KHI_N restrict the k-range of a tile to k < m0+BM / k < n0+BN (op(A) lower triangular /
//        op(B) upper triangular), GEMM_KLO_M / GEMM_KLO_N to k >= m0 / k >= n0 (op(A) upper / op(B) lower triangular).
enum { GEMM_LOWER = 1, GEMM_KHI_M = 2, GEMM_KHI_N = 4, GEMM_KLO_M = 8, GEMM_KLO_N = 16, GEMM_BALANCED = 32 /* internal: set by the launcher */ };
cudaError_t gemm(cudaStream_t st, bool transa, bool transb, int m, int n, int k, double alpha,
   const double* A, int lda, long long strideA, const double* B, int ldb, long long strideB,
   double beta, double* C, int ldc, long long strideC, int batch, int flags);
cudaError_t dmma_peak_probe(cudaStream_t st, int iters, double* d_sink, double* flops, int blocks_per_sm = 4, int threads = 256);

// ---- chol.cu -------------------------------------------------------------------------------------------------------
// Cholesky A = L L' (lower, in place) by recursive blocking on DMMA GEMMs; optionally the inverse of L in Linv.
// d_info: device int, set to the 1-based index of the first non-positive pivot (0 = success; is NOT reset here).
// diaginv: optional workspace receiving the inverses of the NB x NB diagonal blocks of L (block b at b*NB*NB).
constexpr int CHOL_NB = 64;
constexpr int CHOL_LEAF_MAX = 128;   // largest diagonal block factorised by one CTA; work spaces are sized n + 2 * CHOL_LEAF_MAX columns
extern long long* g_diag_dbg;      // optional device buffer (4 x int64) receiving the phase cycle counts of the diagonal-block kernel
cudaError_t potrf_lower(cudaStream_t st, int n, double* A, int lda, double* Linv, int ldi, double* diaginv, double* work, int ldw, int* d_info);
cudaError_t trtri_lower(cudaStream_t st, int n, const double* L, int ldl, double* Linv, int ldi, double* work, int ldw);
// two factorisations of the same order (each with its inverse factor) in one launch of the tile kernel: the two dependency chains
// run side by side on different SMs
cudaError_t potrf_lower_pair(cudaStream_t st, int n, double* A0, double* Linv0, double* work0, int* info0, double* A1, double* Linv1,
   double* work1, int* info1, int ld, int ldw);
// solves L L' x = b for one right-hand side using the diagonal-block inverses (x overwrites b; tmp: n doubles)
// large orders: right-looking panels (width pb) with one panel of look-ahead on a side stream; pinv / pinvT receive the inverses
// of the diagonal panel blocks and their transposes (ceil(n/pb) blocks of pb x pb); ev: 2*ceil(n/pb)+2 events without timing
cudaError_t potrf_lower_panels(cudaStream_t st, int pb, int n, double* A, int lda, double* pinv, double* pinvT, double* work, int ldw, int* d_info);
cudaError_t potrf_lower_lookahead(cudaStream_t st, cudaStream_t side, cudaEvent_t* ev, int nev, int pb, int n, double* A, int lda,
   double* pinv, double* pinvT, double* work, int ldw, int* d_info);
cudaError_t potrs_panels(cudaStream_t st, int pb, int n, const double* L, const double* LT, int ldl, const double* pinv, const double* pinvT,
   double* b, double* tmp);

// ---- eig.cu --------------------------------------------------------------------------------------------------------
// batched symmetric eigen-decomposition by parallel cyclic Jacobi in shared memory (n <= JACOBI_MAX_N)
constexpr int JACOBI_MAX_N = 96;
cudaError_t jacobi_eig_batched(cudaStream_t st, int n, int nbatch, const double* A, int lda, long long strideA,
   double* w, double* V /* or nullptr */, int* d_sweeps);
// batched, adaptive variant: all matrices advance together, convergence is checked on the host every 8 steps
constexpr int LZB_MAXIT = 64;
struct LzDesc
{
   int n; int ld;
   const double* B;      // explicit symmetric matrix (full storage), or nullptr when the operator is implicit
   double* Q;            // (LZB_MAXIT+2)*n Lanczos vectors
   double* ab;           // 2*LZB_MAXIT recurrence coefficients
   double* out;          // 3 results: safe value, Ritz value, residual bound
   double* safe;         // extra destination of out[0], or nullptr
   // implicit operator  v -> W (D (W' v))  with W lower triangular, WT = W' stored, D symmetric (all n x n, leading dimension ld)
   const double* W; const double* WT; const double* D;
   double* t1; double* t2;
};
// blocks of order <= LZS_MAX_N: one CTA per matrix runs the whole recurrence out of shared memory (one launch, no host check)
constexpr int LZS_MAX_N = 128;
cudaError_t lanczos_small_batched(cudaStream_t st, int nmat, int maxn, const LzDesc* d_desc, int maxit);
// tickets: nmat zero-initialised counters; partials: nmat * pstride doubles with pstride >= ceil(max n / 8)
cudaError_t lanczos_batched(cudaStream_t st, int nmat, const LzDesc* h_desc, LzDesc* d_desc, int maxit, double* d_out3,
   double* h_out3, int* steps_done, unsigned* tickets, double* partials, int pstride);

} // namespace sdpk
