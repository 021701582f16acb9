// eig.cu — symmetric eigenvalue kernels.
//
// (1) jacobi_eig_batched: one CTA per matrix, parallel cyclic Jacobi (round-robin "chess tournament" ordering: n/2
//     disjoint rotations per step, applied first to the columns, then to the rows) with the matrix and the accumulated
//     eigenvectors resident in shared memory for n <= JACOBI_MAX_N.  It replaces the DSYEVR calls behind
//     SCIPlapackComputeIthEigenvalue / ComputeEigenvectorsNegative / ComputeEigenvectorDecomposition
//     (lapack_interface.c:178-603): eigenvalues ascending, eigenvector k stored as row k of the output.
//     Larger matrices run the same code on a global-memory scratch (correct, not fast; the cut-separation matrices of the
//     reference's instances have n = 10..43).
// (2) lanczos_small_batched / lanczos_batched: smallest eigenvalue of the step-length matrices L^-1 dX L^-T by Lanczos (plain
//     three-term recurrence), returning Ritz value minus residual bound, i.e. a safe value for the step length -1/lambda_min:
//     one CTA per matrix out of shared memory for orders <= 128, one launch per step over all matrices above, either on the
//     explicit matrix or on the operator v -> W (D (W' v)).
#include "common.cuh"
#include <algorithm>

namespace sdpk {
namespace {

__device__ __forceinline__ double block_sum(double v, double* red)
{
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
   for( int o = 16; o > 0; o >>= 1 ) v += __shfl_xor_sync(0xffffffffu, v, o);
   __syncthreads();
   if( lane == 0 ) red[warp] = v;
   __syncthreads();
   double s = 0.0;
   for( int w = 0; w < nw; ++w ) s += red[w];      // fixed order: deterministic
   return s;
}

// A: n x n symmetric (full), ld = lds;  V: eigenvector accumulator or nullptr
__global__ void __launch_bounds__(256)
jacobi_kernel(int n, const double* __restrict__ Ain, int lda, long long strideA, double* __restrict__ wout,
   double* __restrict__ Vout, double* __restrict__ gscratch, int use_smem, int* __restrict__ sweeps_out)
{
   extern __shared__ __align__(16) double jsm[];
   const int np = (n + 1) / 2;               // rotation pairs per step
   const int lds = n | 1;                    // odd leading dimension: conflict-free column and row walks
   const int tid = threadIdx.x, nt = blockDim.x;
   const int b = blockIdx.x;
   const bool wantV = (Vout != nullptr);

   double* A; double* V;
   double* aux = jsm;                        // [0,32) reduction scratch, then c[np], s[np], then int pairs
   double* cs = aux + 32;
   int* top = reinterpret_cast<int*>(cs + 2 * np);
   int* bot = top + np;
   if( use_smem )
   {
      A = reinterpret_cast<double*>(bot + np + (np & 1) * 1) ;
      A = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(A) + 15) & ~uintptr_t(15));
      V = A + (size_t)n * lds;
   }
   else
   {
      A = gscratch + (size_t)b * 2 * n * lds;
      V = A + (size_t)n * lds;
   }
   const double* Ab = Ain + (size_t)b * strideA;

   double fro = 0.0;
   for( int e = tid; e < n * n; e += nt )
   {
      int i = e % n, j = e / n;
      double v = 0.5 * (Ab[(size_t)j * lda + i] + Ab[(size_t)i * lda + j]);
      A[i * lds + j] = v;
      fro += v * v;
      if( wantV ) V[i * lds + j] = (i == j) ? 1.0 : 0.0;
   }
   for( int k = tid; k < np; k += nt ) { top[k] = 2 * k; bot[k] = 2 * k + 1; }     // index n (if n odd) is a dummy
   fro = block_sum(fro, aux);
   const double tol2 = fro * 1e-30 + 1e-300;

   int sweep = 0;
   for( ; sweep < 40; ++sweep )
   {
      double off = 0.0;
      for( int e = tid; e < n * n; e += nt )
      {
         int i = e % n, j = e / n;
         if( i != j ) off += A[i * lds + j] * A[i * lds + j];
      }
      off = block_sum(off, aux);
      if( off <= tol2 ) break;

      for( int step = 0; step < 2 * np - 1; ++step )
      {
         // rotation parameters
         for( int k = tid; k < np; k += nt )
         {
            int p = min(top[k], bot[k]), q = max(top[k], bot[k]);
            double c = 1.0, s = 0.0;
            if( q < n )
            {
               double apq = A[p * lds + q];
               if( fabs(apq) > 1e-300 )
               {
                  double tau = (A[q * lds + q] - A[p * lds + p]) / (2.0 * apq);
                  double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                  c = 1.0 / sqrt(1.0 + t * t);
                  s = t * c;
               }
            }
            cs[2 * k] = c; cs[2 * k + 1] = s;
         }
         __syncthreads();
         // columns: A <- A J, V <- V J
         for( int e = tid; e < np * n; e += nt )
         {
            int k = e / n, i = e % n;
            int p = min(top[k], bot[k]), q = max(top[k], bot[k]);
            if( q >= n ) continue;
            double c = cs[2 * k], s = cs[2 * k + 1];
            double aip = A[i * lds + p], aiq = A[i * lds + q];
            A[i * lds + p] = c * aip - s * aiq;
            A[i * lds + q] = s * aip + c * aiq;
            if( wantV )
            {
               double vip = V[i * lds + p], viq = V[i * lds + q];
               V[i * lds + p] = c * vip - s * viq;
               V[i * lds + q] = s * vip + c * viq;
            }
         }
         __syncthreads();
         // rows: A <- J' A
         for( int e = tid; e < np * n; e += nt )
         {
            int k = e / n, j = e % n;
            int p = min(top[k], bot[k]), q = max(top[k], bot[k]);
            if( q >= n ) continue;
            double c = cs[2 * k], s = cs[2 * k + 1];
            double apj = A[p * lds + j], aqj = A[q * lds + j];
            A[p * lds + j] = c * apj - s * aqj;
            A[q * lds + j] = s * apj + c * aqj;
         }
         __syncthreads();
         // next pairing: player top[0] stays, the others move one seat
         int nt_k = -1, nb_k = -1;
         if( tid < np )
         {
            int k = tid;
            nt_k = (k == 0) ? top[0] : ((k == 1) ? bot[0] : top[k - 1]);
            nb_k = (k == np - 1) ? top[np - 1] : bot[k + 1];
            if( np == 1 ) { nt_k = top[0]; nb_k = bot[0]; }
         }
         __syncthreads();
         if( tid < np ) { top[tid] = nt_k; bot[tid] = nb_k; }
         __syncthreads();
      }
   }
   if( sweeps_out != nullptr && tid == 0 ) sweeps_out[b] = sweep;

   // ascending order by rank counting (ties broken by index), eigenvector k written as row k
   for( int i = tid; i < n; i += nt )
   {
      double d = A[i * lds + i];
      int rank = 0;
      for( int j = 0; j < n; ++j )
      {
         double dj = A[j * lds + j];
         rank += (dj < d || (dj == d && j < i)) ? 1 : 0;
      }
      wout[(size_t)b * n + rank] = d;
   }
   if( wantV )
   {
      __syncthreads();
      for( int e = tid; e < n * n; e += nt )
      {
         int i = e / n, r = e % n;       // eigenvector of column i, component r
         double d = A[i * lds + i];
         int rank = 0;
         for( int j = 0; j < n; ++j )
         {
            double dj = A[j * lds + j];
            rank += (dj < d || (dj == d && j < i)) ? 1 : 0;
         }
         Vout[(size_t)b * n * n + (size_t)rank * n + r] = V[r * lds + i];
      }
   }
}

// ------------------------------------------------------------------------------------------------------------------
// ---- batched Lanczos: all step-length matrices of one predictor/corrector pass advance together ---------------------
__global__ void lzb_init_kernel(const LzDesc* __restrict__ D)
{
   __shared__ double red[32];
   const LzDesc d = D[blockIdx.x];
   double s = 0.0;
   for( int i = threadIdx.x; i < d.n; i += blockDim.x )
   {
      unsigned h = (unsigned)i * 2654435761u + 12345u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      double v = 0.5 + (double)(h & 0xffffu) / 65536.0;
      d.Q[i] = v;
      s += v * v;
   }
   s = block_sum(s, red);
   double inv = 1.0 / sqrt(s);
   for( int i = threadIdx.x; i < d.n; i += blockDim.x ) d.Q[i] *= inv;
}

// Dot products of 8 consecutive columns i0 .. i0+7 of a column-major matrix with a vector, by one CTA of 256 threads.
// mode 0: full column; mode 1: entries k >= column index (lower triangular W); mode 2: entries k <= column index (upper
// triangular W').  All 256 threads sweep every column together (16-byte loads, 8 columns x 2 chunks issued before the first
// use), so that a whole 8-column panel is in flight at once instead of one 16 KB column per warp; the 8 sums are reduced
// through shared memory in a fixed order.  Results in res[0..7] (valid after the trailing barrier).
__device__ __forceinline__ void coldot8(const double* __restrict__ Mx, int ld, int n, int i0, int mode,
   const double* __restrict__ vin, double* res /* shared, 8 */, double (*part)[8] /* shared, [8 warps][8] */)
{
   const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
   double acc[8];
#pragma unroll
   for( int c = 0; c < 8; ++c ) acc[c] = 0.0;
   const int ncol = min(8, n - i0);
   const int nev = n & ~1;
   // triangular operands: whole 1024-row chunks outside the non-zero range of these 8 columns are skipped
   const int kbeg = (mode == 1) ? (i0 / 1024) * 1024 : 0;
   const int kfin = (mode == 2) ? min(nev, i0 + 8) : nev;
   for( int k = kbeg + 2 * tid; k < kfin; k += 1024 )
   {
      const int k2 = k + 512;
      const bool in2 = k2 < kfin;
      // the vector may start at an odd element: scalar loads (cached), 16-byte loads only for the matrix columns
      const double2 v0 = make_double2(vin[k], vin[k + 1]);
      const double2 v1 = in2 ? make_double2(vin[k2], vin[k2 + 1]) : make_double2(0.0, 0.0);
      double2 m0[8], m1[8];
#pragma unroll
      for( int c = 0; c < 8; ++c )
      {
         const double* col = Mx + (size_t)(i0 + c) * ld;
         m0[c] = (c < ncol) ? *reinterpret_cast<const double2*>(col + k) : make_double2(0.0, 0.0);
         m1[c] = (c < ncol && in2) ? *reinterpret_cast<const double2*>(col + k2) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for( int c = 0; c < 8; ++c )
      {
         const int i = i0 + c;
         if( mode == 0 )
            acc[c] += m0[c].x * v0.x + m0[c].y * v0.y + m1[c].x * v1.x + m1[c].y * v1.y;
         else if( mode == 1 )
            acc[c] += (k >= i ? m0[c].x * v0.x : 0.0) + (k + 1 >= i ? m0[c].y * v0.y : 0.0)
                    + (k2 >= i ? m1[c].x * v1.x : 0.0) + (k2 + 1 >= i ? m1[c].y * v1.y : 0.0);
         else
            acc[c] += (k <= i ? m0[c].x * v0.x : 0.0) + (k + 1 <= i ? m0[c].y * v0.y : 0.0)
                    + (k2 <= i ? m1[c].x * v1.x : 0.0) + (k2 + 1 <= i ? m1[c].y * v1.y : 0.0);
      }
   }
   if( (n & 1) && tid == 0 )
   {
      const int k = n - 1;
#pragma unroll
      for( int c = 0; c < 8; ++c )
      {
         const int i = i0 + c;
         if( c < ncol && (mode == 0 || (mode == 1 && k >= i) || (mode == 2 && k <= i)) )
            acc[c] += Mx[(size_t)i * ld + k] * vin[k];
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
   if( tid < 8 )
   {
      double v = 0.0;
#pragma unroll
      for( int q = 0; q < 8; ++q ) v += part[q][tid];
      res[tid] = v;
   }
   __syncthreads();
}

// first two stages of the implicit operator v -> W (D (W' v)).  stage 0: t1 = W' v_j (column i of W, entries i..n-1);
// stage 1: t2 = D t1 (full column of D)
__global__ void __launch_bounds__(256)
lz_coldot_kernel(const LzDesc* __restrict__ D, int j, int stage)
{
   __shared__ double res[8];
   __shared__ double part[8][8];
   const LzDesc d = D[blockIdx.y];
   const int n = d.n, i0 = blockIdx.x * 8;
   if( i0 >= n || j >= n || d.B != nullptr ) return;
   const double* vin = (stage == 0) ? d.Q + (size_t)j * n : d.t1;
   coldot8(stage == 0 ? d.W : d.D, d.ld, n, i0, stage == 0 ? 1 : 0, vin, res, part);
   double* out = (stage == 0) ? d.t1 : d.t2;
   if( threadIdx.x < 8 && i0 + (int)threadIdx.x < n ) out[i0 + threadIdx.x] = res[threadIdx.x];
}

// one Lanczos step in ONE launch: every CTA forms 8 rows of u = B v_j (warp per row, contiguous column of the symmetric
// matrix) and its share of alpha = u.v_j; the CTA that finishes last (atomic ticket) completes the step for the whole vector:
// alpha, w = u - alpha v_j - beta_{j-1} v_{j-1}, beta_j = |w|, v_{j+1} = w / beta_j.  Plain three-term recurrence (no
// re-orthogonalisation: the extreme Ritz value and its residual bound do not need it); partial sums are added in a fixed order.
__global__ void __launch_bounds__(256)
lzb_step_kernel(const LzDesc* __restrict__ D, int j, int maxit, unsigned* __restrict__ tickets, double* __restrict__ partials, int pstride)
{
   __shared__ double red[32];
   __shared__ bool last;
   const LzDesc d = D[blockIdx.y];
   const int n = d.n, tid = threadIdx.x;
   const int nblk = (n + 7) / 8;
   if( (int)blockIdx.x >= nblk || j >= n ) return;
   __shared__ double res[8];
   __shared__ double part[8][8];
   double* Q = d.Q;
   const double* __restrict__ v = Q + (size_t)j * n;
   double* w = Q + (size_t)(j + 1) * n;
   // explicit matrix: u = B v_j.  Implicit operator: last stage u = W t2, row i of W = column i of WT, entries 0..i
   const bool implicit = (d.B == nullptr);
   const int i0 = blockIdx.x * 8;
   coldot8(implicit ? d.WT : d.B, d.ld, n, i0, implicit ? 2 : 0, implicit ? d.t2 : v, res, part);
   if( tid < 8 )
   {
      const int i = i0 + tid;
      double contrib = 0.0;
      if( i < n ) { w[i] = res[tid]; contrib = res[tid] * v[i]; }
      red[tid] = contrib;
   }
   __syncthreads();
   if( tid == 0 )
   {
      double a = 0.0;
      for( int q = 0; q < 8; ++q ) a += red[q];
      partials[(size_t)blockIdx.y * pstride + blockIdx.x] = a;
      __threadfence();
      unsigned t = atomicAdd(&tickets[blockIdx.y], 1u);
      last = (t == (unsigned)nblk - 1);
   }
   __syncthreads();
   if( !last ) return;
   __threadfence();
   // ---- tail: the whole vector, one CTA ----
   double a = 0.0;
   for( int b = tid; b < nblk; b += 256 ) a += __ldcg(&partials[(size_t)blockIdx.y * pstride + b]);
   a = block_sum(a, red);
   double* alpha = d.ab;
   double* beta = d.ab + maxit;
   const double bprev = (j > 0) ? beta[j - 1] : 0.0;
   const double* vprev = (j > 0) ? Q + (size_t)(j - 1) * n : nullptr;
   double nr = 0.0;
   for( int r = tid; r < n; r += 256 )
   {
      double x = __ldcg(&w[r]) - a * v[r] - (j > 0 ? bprev * vprev[r] : 0.0);
      w[r] = x;
      nr += x * x;
   }
   nr = sqrt(block_sum(nr, red));
   const double cf = (nr > 1e-300) ? 1.0 / nr : 0.0;
   for( int r = tid; r < n; r += 256 ) w[r] *= cf;
   if( tid == 0 ) { alpha[j] = a; beta[j] = nr; tickets[blockIdx.y] = 0u; }
}

// ---- several Lanczos steps in ONE cooperative launch (explicit matrices) ---------------------------------------------------------
// The step kernel above moves 2 x 32 MB in 22.7 us (max-cut 2000): launch, one memory round trip per CTA, the ticket and the tail of
// the last CTA add up to more than the streaming itself.  Here the grid stays resident for a whole chunk of steps (four CTAs per SM,
// cooperative launch); a step is two phases separated by grid-wide barriers (arrive counter + generation word, release/acquire):
//   A  r = B x_j for the CTA's groups of 8 rows (same coldot8 as above) and its share of x_j' B x_j
//   B  every CTA adds the shares in the same fixed order (alpha), forms x_{j+1} = s_j r - alpha s_j x_j - beta_{j-1} s_{j-1} x_{j-1} for
//      its rows and its share of |x_{j+1}|^2
// where v_j = s_j x_j, s_j = 1 / |x_j|: the vectors are kept UNSCALED, so that no third phase is needed for the normalisation
// (|x_{j+1}| = beta_j is known to every CTA after the second barrier).  Same recurrence and coefficients as the step kernel.
__device__ __forceinline__ void lz_grid_sync(unsigned* bar, unsigned& gen, unsigned nblocks)
{
   __syncthreads();
   if( threadIdx.x == 0 )
   {
      ++gen;
      __threadfence();
      const unsigned old = atomicAdd(bar, 1u);
      if( old == gen * nblocks - 1u )
         asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(bar + 1), "r"(gen) : "memory");
      else
      {
         unsigned seen, spins = 0;
         do
         {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar + 1) : "memory");
            if( ++spins > (1u << 25) ) asm volatile("trap;");      // a bug, not a wait (the launch is cooperative: all CTAs are resident)
         } while( seen < gen );
      }
   }
   __syncthreads();
}

__global__ void __launch_bounds__(256)
lzb_persist_kernel(const LzDesc* __restrict__ D, int nmat, int j0, int jend, int maxit, unsigned* __restrict__ bar,
   double* __restrict__ partials, int pstride)
{
   __shared__ double red[32];
   __shared__ double res[8];
   __shared__ double part[8][8];
   const int tid = threadIdx.x;
   unsigned gen = 0;
   double* const pA = partials;                                  // shares of x' B x   [mat][group]
   double* const pB = partials + (size_t)nmat * pstride;         // shares of |x_{j+1}|^2
   int maxgroups = 0;
   for( int q = 0; q < nmat; ++q ) maxgroups = max(maxgroups, (D[q].n + 7) / 8);
   const int total = nmat * maxgroups;
   for( int j = j0; j < jend; ++j )
   {
      // ---- phase A ----
      for( int g = blockIdx.x; g < total; g += gridDim.x )
      {
         const int q = g / maxgroups, grp = g - q * maxgroups;
         const LzDesc d = D[q];
         const int n = d.n, i0 = 8 * grp;
         if( i0 >= n || j >= n ) continue;
         const double* x = d.Q + (size_t)j * n;
         double* w = d.Q + (size_t)(j + 1) * n;
         coldot8(d.B, d.ld, n, i0, 0, x, res, part);
         if( tid < 8 )
         {
            const int i = i0 + tid;
            double contrib = 0.0;
            if( i < n ) { w[i] = res[tid]; contrib = res[tid] * x[i]; }
            red[tid] = contrib;
         }
         __syncthreads();
         if( tid == 0 )
         {
            double a = 0.0;
            for( int t = 0; t < 8; ++t ) a += red[t];
            pA[(size_t)q * pstride + grp] = a;
         }
         __syncthreads();
      }
      lz_grid_sync(bar, gen, gridDim.x);
      // ---- phase B ----
      for( int g = blockIdx.x; g < total; g += gridDim.x )
      {
         const int q = g / maxgroups, grp = g - q * maxgroups;
         const LzDesc d = D[q];
         const int n = d.n, i0 = 8 * grp, nblk = (n + 7) / 8;
         if( i0 >= n || j >= n ) continue;
         double* alpha = d.ab;
         double* beta = d.ab + maxit;
         double a = 0.0;
         for( int b = tid; b < nblk; b += 256 ) a += __ldcg(&pA[(size_t)q * pstride + b]);
         a = block_sum(a, red);                                 // the same order in every CTA
         const double bprev = (j > 0) ? __ldcg(&beta[j - 1]) : 0.0;                    // = |x_j|
         const double sj = (j > 0) ? ((bprev > 1e-300) ? 1.0 / bprev : 0.0) : 1.0;
         const double bpp = (j > 1) ? __ldcg(&beta[j - 2]) : 0.0;                      // = |x_{j-1}|
         const double sjm = (j > 1) ? ((bpp > 1e-300) ? 1.0 / bpp : 0.0) : 1.0;
         const double al = sj * sj * a;
         const double* x = d.Q + (size_t)j * n;
         const double* xp = d.Q + (size_t)(j > 0 ? j - 1 : 0) * n;
         double* w = d.Q + (size_t)(j + 1) * n;
         if( tid < 8 )
         {
            const int i = i0 + tid;
            double v = 0.0;
            if( i < n )
            {
               v = sj * __ldcg(&w[i]) - al * sj * __ldcg(&x[i]) - (j > 0 ? bprev * sjm * __ldcg(&xp[i]) : 0.0);
               w[i] = v;
            }
            red[tid] = v * v;
         }
         __syncthreads();
         if( tid == 0 )
         {
            double s2 = 0.0;
            for( int t = 0; t < 8; ++t ) s2 += red[t];
            pB[(size_t)q * pstride + grp] = s2;
            if( grp == 0 ) alpha[j] = al;
         }
         __syncthreads();
      }
      lz_grid_sync(bar, gen, gridDim.x);
      // ---- beta_j = |x_{j+1}|: written once per matrix (by the CTA of its first group) ----
      for( int g = blockIdx.x; g < total; g += gridDim.x )
      {
         const int q = g / maxgroups, grp = g - q * maxgroups;
         if( grp != 0 ) continue;
         const LzDesc d = D[q];
         const int n = d.n, nblk = (n + 7) / 8;
         if( j >= n ) continue;
         double s2 = 0.0;
         for( int b = tid; b < nblk; b += 256 ) s2 += __ldcg(&pB[(size_t)q * pstride + b]);
         s2 = block_sum(s2, red);
         if( tid == 0 ) d.ab[maxit + j] = sqrt(s2);      // read in phase B of the next step, i.e. behind its first barrier
      }
   }
}

// smallest eigenvalue of the kk x kk Lanczos tridiagonal (a, bt) by 32-way multisection on Sturm counts, one warp;
// theta = lower end of the final bracket, resid = |beta_kk s_kk| (0 when the Krylov space is exhausted or kk = n)
__device__ __forceinline__ void ritz_warp(const double* a, const double* bt, int kk, int n, int lane, double& theta, double& resid)
{
   double scale = 0.0;
   for( int i = 0; i < kk; ++i ) scale = fmax(scale, fabs(a[i]) + fabs(bt[i]));
   int keff = kk;
   for( int i = 0; i < kk - 1; ++i ) if( fabs(bt[i]) <= 1e-14 * scale ) { keff = i + 1; break; }
   double lo = 1e300, hi = -1e300;
   for( int i = 0; i < keff; ++i )
   {
      double r = (i > 0 ? fabs(bt[i - 1]) : 0.0) + (i < keff - 1 ? fabs(bt[i]) : 0.0);
      lo = fmin(lo, a[i] - r); hi = fmax(hi, a[i] + r);
   }
   const double width0 = hi - lo;
   for( int round = 0; round < 6 && (hi - lo) > 1e-9 * width0 + 1e-300; ++round )
   {
      const double xt = lo + (lane + 1) * (hi - lo) / 33.0;
      int cnt = 0;
      double dd = 1.0;
      for( int i = 0; i < keff; ++i )
      {
         double b2 = (i > 0) ? bt[i - 1] * bt[i - 1] : 0.0;
         dd = a[i] - xt - (i > 0 ? b2 / dd : 0.0);
         if( dd == 0.0 ) dd = 1e-300;
         if( dd < 0.0 ) ++cnt;
      }
      double below = (cnt == 0) ? xt : lo;
      double above = (cnt >= 1) ? xt : hi;
#pragma unroll
      for( int o = 16; o > 0; o >>= 1 )
      {
         below = fmax(below, __shfl_xor_sync(0xffffffffu, below, o));
         above = fmin(above, __shfl_xor_sync(0xffffffffu, above, o));
      }
      lo = below; hi = above;
   }
   theta = lo;            // lower end of the bracket: errs on the safe side
   resid = 0.0;
   if( keff == kk && kk < n )
   {
      double sm1 = 0.0, s0 = 1.0, nrm = 1.0, last = 1.0;
      for( int i = 0; i < kk - 1; ++i )
      {
         double s1 = ((theta - a[i]) * s0 - (i > 0 ? bt[i - 1] * sm1 : 0.0)) / bt[i];
         sm1 = s0; s0 = s1;
         nrm += s1 * s1;
         last = s1;
         if( nrm > 1e200 ) { sm1 *= 1e-100; s0 *= 1e-100; last *= 1e-100; nrm *= 1e-200; }
      }
      resid = fabs(bt[kk - 1]) * fabs(last) / sqrt(nrm);
   }
}

__global__ void lzb_ritz_kernel(const LzDesc* __restrict__ D, int k, int maxit)
{
   const LzDesc d = D[blockIdx.x];
   const int lane = threadIdx.x & 31;
   if( threadIdx.x >= 32 ) return;
   double theta, resid;
   ritz_warp(d.ab, d.ab + maxit, min(k, d.n), d.n, lane, theta, resid);
   if( lane == 0 )
   {
      d.out[0] = theta - resid;
      d.out[1] = theta;
      d.out[2] = resid;
      if( d.safe ) *d.safe = theta - resid;
   }
}

// ---- small blocks (n <= LZS_MAX_N): the whole Lanczos run of one matrix in ONE CTA, matrix and vectors in shared memory -----
// Same recurrence, start vector and stopping rule as the batched multi-launch version above; the Ritz value is checked
// (warp 0) after 8, 16, 24, 32, 48, 64, 96 and 128 steps.
__global__ void __launch_bounds__(256)
lz_smem_kernel(const LzDesc* __restrict__ D, int maxit)
{
   extern __shared__ __align__(16) double lsm[];
   __shared__ double red[32];
   __shared__ double res3[3];
   __shared__ int done;
   const LzDesc d = D[blockIdx.x];
   const int n = d.n, lds = n | 1, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
   double* Bs = lsm;
   double* v = Bs + (size_t)n * lds;
   double* vprev = v + LZS_MAX_N;
   double* w = vprev + LZS_MAX_N;
   double* alpha = w + LZS_MAX_N;
   double* beta = alpha + LZS_MAX_N + 1;
   for( int e = tid; e < n * n; e += 256 )
   {
      const int r = e % n, c = e / n;
      Bs[c * lds + r] = d.B[(size_t)c * d.ld + r];
   }
   double s = 0.0;
   if( tid < n )
   {
      unsigned h = (unsigned)tid * 2654435761u + 12345u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      const double x = 0.5 + (double)(h & 0xffffu) / 65536.0;
      v[tid] = x; vprev[tid] = 0.0;
      s = x * x;
   }
   if( tid == 0 ) done = 0;
   s = block_sum(s, red);
   if( tid < n ) v[tid] *= 1.0 / sqrt(s);
   __syncthreads();
   const int kmax = min(maxit, min(n, LZS_MAX_N));
   double bprev = 0.0;
   int j = 0;
   for( ; j < kmax; )
   {
      for( int i = wid; i < n; i += 8 )
      {
         const double* col = Bs + i * lds;        // symmetric: column i = row i
         double acc = 0.0;
         for( int k = lane; k < n; k += 32 ) acc += col[k] * v[k];
#pragma unroll
         for( int o = 16; o > 0; o >>= 1 ) acc += __shfl_xor_sync(0xffffffffu, acc, o);
         if( lane == 0 ) w[i] = acc;
      }
      __syncthreads();
      const double a = block_sum(tid < n ? w[tid] * v[tid] : 0.0, red);
      double x = 0.0;
      if( tid < n ) x = w[tid] - a * v[tid] - bprev * vprev[tid];
      const double nr = sqrt(block_sum(x * x, red));
      const double cf = (nr > 1e-300) ? 1.0 / nr : 0.0;
      if( tid < n ) { vprev[tid] = v[tid]; v[tid] = x * cf; }
      if( tid == 0 ) { alpha[j] = a; beta[j] = nr; }
      bprev = nr;
      ++j;
      __syncthreads();
      const bool check = (j == kmax) || (j >= 8 && ((j <= 32 && (j & 7) == 0) || (j <= 64 && (j & 15) == 0) || (j & 31) == 0));
      if( check )
      {
         if( wid == 0 )
         {
            double theta, resid;
            ritz_warp(alpha, beta, j, n, lane, theta, resid);
            if( lane == 0 )
            {
               res3[0] = theta - resid; res3[1] = theta; res3[2] = resid;
               done = (resid <= 0.01 * fabs(theta) || theta - resid >= -0.5) ? 1 : 0;
            }
         }
         __syncthreads();
         if( done ) break;
      }
   }
   if( tid < 3 ) d.out[tid] = res3[tid];
   if( tid == 0 && d.safe ) *d.safe = res3[0];
}

} // namespace

cudaError_t jacobi_eig_batched(cudaStream_t st, int n, int nbatch, const double* A, int lda, long long strideA,
   double* w, double* V, int* d_sweeps)
{
   if( n <= 0 || nbatch <= 0 ) return cudaSuccess;
   const int np = (n + 1) / 2, lds = n | 1;
   size_t aux = (32 + 2 * np) * sizeof(double) + 2 * np * sizeof(int) + 32;
   size_t mat = 2 * (size_t)n * lds * sizeof(double);
   static thread_local double* gscratch = nullptr;      // large-matrix fallback scratch (one per host thread / device use)
   static thread_local size_t gscratch_bytes = 0;
   static thread_local int gscratch_dev = -1;
   int use_smem = (n <= JACOBI_MAX_N) ? 1 : 0;
   size_t smem = aux + (use_smem ? mat : 0);
   if( !use_smem )
   {
      size_t need = mat * nbatch;
      int curdev = 0;
      SDPK_CUDA_CHECK( cudaGetDevice(&curdev) );
      if( need > gscratch_bytes || curdev != gscratch_dev )
      {
         gscratch_dev = curdev;
         if( gscratch ) cudaFree(gscratch);
         SDPK_CUDA_CHECK( cudaMalloc(&gscratch, need) );
         gscratch_bytes = need;
      }
   }
   static bool configured[64] = {false};        // per-device function attribute
   int dev = 0;
   SDPK_CUDA_CHECK( cudaGetDevice(&dev) );
   if( !configured[dev & 63] )
   {
      SDPK_CUDA_CHECK( cudaFuncSetAttribute(jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) );
      configured[dev & 63] = true;
   }
   ProfScope prof(st, PROF_EIG, (double)nbatch * (16.0 * n * n + 8.0 * n));
   jacobi_kernel<<<nbatch, 256, smem, st>>>(n, A, lda, strideA, w, V, gscratch, use_smem, d_sweeps);
   count_launch();
   return cudaGetLastError();
}

cudaError_t lanczos_small_batched(cudaStream_t st, int nmat, int maxn, const LzDesc* d_desc, int maxit)
{
   if( nmat <= 0 ) return cudaSuccess;
   if( maxn > LZS_MAX_N ) return cudaErrorInvalidValue;
   static bool configured[64] = {false};        // per-device function attribute
   int dev = 0;
   SDPK_CUDA_CHECK( cudaGetDevice(&dev) );
   if( !configured[dev & 63] )
   {
      SDPK_CUDA_CHECK( cudaFuncSetAttribute(lz_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) );
      configured[dev & 63] = true;
   }
   const size_t smem = sizeof(double) * ((size_t)maxn * (maxn | 1) + 5 * (size_t)LZS_MAX_N + 8);
   ProfScope prof(st, PROF_EIG, 8.0 * nmat * (double)maxn * maxn);
   lz_smem_kernel<<<nmat, 256, smem, st>>>(d_desc, maxit);
   count_launch();
   return cudaGetLastError();
}

} // namespace sdpk

namespace sdpk {

cudaError_t lanczos_batched(cudaStream_t st, int nmat, const LzDesc* h_desc, LzDesc* d_desc, int maxit, double* d_out3,
   double* h_out3, int* steps_done, unsigned* tickets, double* partials, int pstride)
{
   if( nmat <= 0 ) return cudaSuccess;
   int maxn = 0, minn = 1 << 30;
   double bytes = 0.0;
   bool any_implicit = false;
   for( int i = 0; i < nmat; ++i )
   {
      maxn = std::max(maxn, h_desc[i].n); minn = std::min(minn, h_desc[i].n);
      bytes += 8.0 * h_desc[i].n * (double)h_desc[i].n * (h_desc[i].B == nullptr ? 2.0 : 1.0);
      any_implicit = any_implicit || (h_desc[i].B == nullptr);
   }
   maxit = std::min(maxit, minn);
   SDPK_CUDA_CHECK( cudaMemcpyAsync(d_desc, h_desc, sizeof(LzDesc) * nmat, cudaMemcpyHostToDevice, st) );
   ProfScope prof(st, PROF_EIG, 0.0);
   lzb_init_kernel<<<nmat, 1024, 0, st>>>(d_desc);
   count_launch();
   // explicit matrices: a whole chunk of steps in one cooperative launch - OFF by default: measured on max-cut 2000 a step costs
   // about 19 us instead of 22.7 (two grid barriers over 500 CTAs eat most of what the saved launches and tails give), the solve
   // 5.65 instead of 5.81 ms per iteration, and the differently rounded recurrence needed 19 instead of 18 iterations on the
   // benchmark instance (107.3 vs 104.6 ms).  SDPCUDA_LZ_PERSIST=1 turns it on.
   static int coop_blocks[64] = {0};          // 0: not asked yet, -1: unavailable
   int dev = 0;
   SDPK_CUDA_CHECK( cudaGetDevice(&dev) );
   if( coop_blocks[dev & 63] == 0 )
   {
      int coop = 0, per_sm = 0, nsm = 0;
      cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
      if( coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lzb_persist_kernel, 256, 0) == cudaSuccess && per_sm > 0 )
         coop_blocks[dev & 63] = std::min(per_sm, 4) * nsm;
      else { coop_blocks[dev & 63] = -1; cudaGetLastError(); }
   }
   const char* pe = getenv("SDPCUDA_LZ_PERSIST");
   const bool persist = !any_implicit && coop_blocks[dev & 63] > 0 && (pe != nullptr && pe[0] == '1');
   int j = 0;
   const int chunk = 8;
   while( j < maxit )
   {
      int jend = std::min(maxit, j + chunk);
      if( persist )
      {
         unsigned* bar = tickets + nmat;                      // two words behind the tickets of the step kernel
         SDPK_CUDA_CHECK( cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned), st) );
         const int groups = nmat * ceil_div(maxn, 8);
         int grid = std::min(coop_blocks[dev & 63], std::max(groups, 1));
         int jj0 = j, mi = LZB_MAXIT;
         void* args[] = {(void*)&d_desc, (void*)&nmat, (void*)&jj0, (void*)&jend, (void*)&mi, (void*)&bar, (void*)&partials, (void*)&pstride};
         SDPK_CUDA_CHECK( cudaLaunchCooperativeKernel((const void*)lzb_persist_kernel, dim3(grid), dim3(256), args, 0, st) );
         count_launch();
         j = jend;
      }
      for( ; j < jend; ++j )
      {
         dim3 grid(ceil_div(maxn, 8), nmat);
         if( any_implicit )
         {
            lz_coldot_kernel<<<grid, 256, 0, st>>>(d_desc, j, 0);
            lz_coldot_kernel<<<grid, 256, 0, st>>>(d_desc, j, 1);
            count_launch(2);
         }
         lzb_step_kernel<<<grid, 256, 0, st>>>(d_desc, j, LZB_MAXIT, tickets, partials, pstride);
         count_launch();
      }
      lzb_ritz_kernel<<<nmat, 32, 0, st>>>(d_desc, j, LZB_MAXIT);
      count_launch();
      SDPK_CUDA_CHECK( cudaMemcpyAsync(h_out3, d_out3, sizeof(double) * 3 * nmat, cudaMemcpyDeviceToHost, st) );
      SDPK_CUDA_CHECK( cudaStreamSynchronize(st) );
      bool done = true;
      for( int i = 0; i < nmat; ++i )
      {
         double safe = h_out3[3 * i], theta = h_out3[3 * i + 1], resid = h_out3[3 * i + 2];
         // accurate enough for a step length: 1 % of the Ritz value, or clearly in the range where the full step is taken
         if( !(resid <= 0.01 * std::fabs(theta) || safe >= -0.5) ) done = false;
      }
      if( done ) break;
   }
   if( g_prof && g_prof->on && !g_prof->recs.empty() ) g_prof->recs.back().work = bytes * j;
   if( steps_done ) *steps_done = j;
   return cudaGetLastError();
}

} // namespace sdpk
