// ops.cu — see ops.cuh.  HBM-bound kernels: coalesced arena sweeps, grid-stride loops on a fixed grid of RED_BLOCKS CTAs
// for everything that reduces (partials are summed in a fixed order, so results are bit-reproducible run to run).
#include "ops.cuh"

namespace sdpk {

// layout of the statistics row: columns [0,16) are sums, [16,24) are maxima
constexpr int NSTAT = 24;
constexpr int NSUM = 16;

namespace {

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
   for( int o = 16; o > 0; o >>= 1 ) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
   for( int o = 16; o > 0; o >>= 1 ) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
   return v;
}
// block reductions in a fixed order; blockDim.x must be a multiple of 32 and <= 1024
__device__ double block_sum(double v, double* red)
{
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
   v = warp_sum(v);
   __syncthreads();
   if( lane == 0 ) red[warp] = v;
   __syncthreads();
   double s = 0.0;
   for( int w = 0; w < nw; ++w ) s += red[w];
   return s;
}
__device__ double block_max(double v, double* red)
{
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
   v = warp_max(v);
   __syncthreads();
   if( lane == 0 ) red[warp] = v;
   __syncthreads();
   double s = red[0];
   for( int w = 1; w < nw; ++w ) s = fmax(s, red[w]);
   return s;
}

__global__ void assemble_kernel(int npos, const int* __restrict__ posbeg, const long long* __restrict__ pos,
   const long long* __restrict__ mirror, const int* __restrict__ posvar, const double* __restrict__ posval,
   const double* __restrict__ posc, const double* __restrict__ y, double cscale, double* __restrict__ T)
{
   int p = blockIdx.x * blockDim.x + threadIdx.x;
   if( p >= npos ) return;
   double v = -cscale * posc[p];
   for( int e = posbeg[p]; e < posbeg[p + 1]; ++e ) v += y[posvar[e]] * posval[e];
   T[pos[p]] = v;
   T[mirror[p]] = v;
}

__global__ void __launch_bounds__(256)
residual_matrix_kernel(size_t arena, const double* __restrict__ T, const double* __restrict__ S, const double* __restrict__ X,
   double* __restrict__ Rd, double* __restrict__ partials)
{
   __shared__ double red[32];
   double s0 = 0.0, s1 = 0.0;
   for( size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < arena; i += (size_t)gridDim.x * blockDim.x )
   {
      double s = S[i], r = T[i] - s;
      Rd[i] = r;
      s0 += r * r;
      s1 += X[i] * s;
   }
   s0 = block_sum(s0, red);
   s1 = block_sum(s1, red);
   if( threadIdx.x == 0 ) { partials[blockIdx.x * NSTAT + 0] = s0; partials[blockIdx.x * NSTAT + 1] = s1; }
}

__global__ void apply_A_kernel(int m, DevEntries E, const int* __restrict__ skipcls, const double* __restrict__ X, double* __restrict__ out)
{
   int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
   if( warp >= m ) return;
   if( skipcls != nullptr && skipcls[warp] == 2 ) return;       // dense constraint matrices: apply_A_dense_kernel
   double s = 0.0;
   for( int e = E.varbeg[warp] + lane; e < E.varbeg[warp + 1]; e += 32 )
   {
      int r = E.row[e], c = E.col[e], ld = E.ld[e];
      const double* Xk = X + E.off[e];
      double v = Xk[(size_t)c * ld + r];
      if( r != c ) v += Xk[(size_t)r * ld + c];
      s += E.val[e] * v;
   }
   s = warp_sum(s);
   if( lane == 0 ) out[warp] = s;
}

__global__ void __launch_bounds__(1024)
const_dots_kernel(int cnnz, const long long* __restrict__ cpos, const long long* __restrict__ cmirror,
   const double* __restrict__ cval, const double* __restrict__ X, const double* __restrict__ Y, double* __restrict__ out2)
{
   __shared__ double red[32];
   double a = 0.0, b = 0.0;
   for( int e = threadIdx.x; e < cnnz; e += blockDim.x )
   {
      long long p = cpos[e], q = cmirror[e];
      double c = cval[e];
      a += c * (X[p] + (p != q ? X[q] : 0.0));
      if( Y != nullptr ) b += c * (Y[p] + (p != q ? Y[q] : 0.0));
   }
   a = block_sum(a, red);
   b = block_sum(b, red);
   if( threadIdx.x == 0 ) { out2[0] = a; out2[1] = b; }
}

__global__ void __launch_bounds__(256)
lp_rows_kernel(int nlp, const int* __restrict__ lpbeg, const int* __restrict__ lpind, const double* __restrict__ lpval,
   const double* __restrict__ lprhs, const double* __restrict__ y, const double* __restrict__ x, const double* __restrict__ s,
   double* __restrict__ Dy, double* __restrict__ rdlp, double* __restrict__ partials)
{
   __shared__ double red[32];
   double a2 = 0.0, axs = 0.0, adx = 0.0, aray = 0.0, amax = 0.0;
   for( int l = blockIdx.x * blockDim.x + threadIdx.x; l < nlp; l += gridDim.x * blockDim.x )
   {
      double d = 0.0;
      for( int p = lpbeg[l]; p < lpbeg[l + 1]; ++p ) d += lpval[p] * y[lpind[p]];
      double r = d - lprhs[l] - s[l];
      Dy[l] = d;
      rdlp[l] = r;
      a2 += r * r;
      axs += x[l] * s[l];
      adx += lprhs[l] * x[l];
      double h = d - s[l];
      aray += h * h;
      amax = fmax(amax, fabs(r));
   }
   a2 = block_sum(a2, red); axs = block_sum(axs, red); adx = block_sum(adx, red); aray = block_sum(aray, red);
   amax = block_max(amax, red);
   if( threadIdx.x == 0 )
   {
      double* p = partials + blockIdx.x * NSTAT;
      p[2] = a2; p[3] = axs; p[4] = adx; p[5] = aray; p[16] = amax;
   }
}

__global__ void lp_cols_kernel(int m, const int* __restrict__ colbeg, const int* __restrict__ colrow, const double* __restrict__ colval,
   const double* __restrict__ x, double* __restrict__ out, int accumulate)
{
   int j = blockIdx.x * blockDim.x + threadIdx.x;
   if( j >= m ) return;
   double s = 0.0;
   for( int p = colbeg[j]; p < colbeg[j + 1]; ++p ) s += colval[p] * x[colrow[p]];
   out[j] = accumulate ? out[j] + s : s;
}

__global__ void __launch_bounds__(256)
primal_residual_kernel(int m, const double* __restrict__ b, const double* __restrict__ AX, const double* __restrict__ DTx,
   const double* __restrict__ y, double* __restrict__ rp, double* __restrict__ partials)
{
   __shared__ double red[32];
   double a2 = 0.0, aray = 0.0, aby = 0.0, amax = 0.0;
   for( int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x )
   {
      double ax = AX[j] + DTx[j];
      double r = b[j] - ax;
      rp[j] = r;
      a2 += r * r;
      aray += ax * ax;
      aby += b[j] * y[j];
      amax = fmax(amax, fabs(r));
   }
   a2 = block_sum(a2, red); aray = block_sum(aray, red); aby = block_sum(aby, red); amax = block_max(amax, red);
   if( threadIdx.x == 0 )
   {
      double* p = partials + blockIdx.x * NSTAT;
      p[6] = a2; p[7] = aray; p[8] = aby; p[17] = amax;
   }
}

__global__ void finalize_kernel(const double* __restrict__ partials, int nblocks, double* __restrict__ out)
{
   int c = threadIdx.x;
   if( c >= NSTAT ) return;
   double v = partials[c];
   for( int b = 1; b < nblocks; ++b )
   {
      double p = partials[b * NSTAT + c];
      v = (c < NSUM) ? v + p : fmax(v, p);
   }
   out[c] = v;
}

// ---------------------------------------------------------------------------------------------------------------------
// Schur complement, entry/gather path:  M_ij = sum over entry pairs of A_i, A_j in the same block of
//     a_i(p,q) a_j(r,c) [ X(q,r) Z(c,p) + X(q,c) Z(r,p) + X(p,r) Z(c,q) + X(p,c) Z(r,q) ]   (terms dropped on diagonals)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double pair_term(const DevEntries& E, int ei, int ej, const double* __restrict__ X, const double* __restrict__ Z)
{
   if( E.off[ei] != E.off[ej] ) return 0.0;
   const int ld = E.ld[ei];
   const double* Xk = X + E.off[ei];
   const double* Zk = Z + E.off[ei];
   const int p = E.row[ei], q = E.col[ei], r = E.row[ej], c = E.col[ej];
   double t = Xk[(size_t)r * ld + q] * Zk[(size_t)p * ld + c];
   if( r != c ) t += Xk[(size_t)c * ld + q] * Zk[(size_t)p * ld + r];
   if( p != q )
   {
      t += Xk[(size_t)r * ld + p] * Zk[(size_t)q * ld + c];
      if( r != c ) t += Xk[(size_t)c * ld + p] * Zk[(size_t)q * ld + r];
   }
   return E.val[ei] * E.val[ej] * t;
}

// light pairs: one thread per (i, j), i >= j, both variables light
__global__ void __launch_bounds__(256)
schur_light_kernel(int m, DevEntries E, const int* __restrict__ heavy, const double* __restrict__ X, const double* __restrict__ Z,
   double* __restrict__ M, int ldm, int nranks, int rank)
{
   const int i = blockIdx.x * 32 + (threadIdx.x & 31);
   const int j = blockIdx.y * 8 + (threadIdx.x >> 5);
   if( blockIdx.x * 32 + 31 < blockIdx.y * 8 ) return;          // tile strictly above the diagonal
   if( (int)(blockIdx.y % nranks) != rank ) return;             // column strips of 8 are dealt round-robin to the ranks
   if( i >= m || j >= m || i < j ) return;
   const int ci = heavy[i], cj = heavy[j];
   if( ci == 1 || ci == 2 || cj == 1 || cj == 2 ) return;      // classes 1 (heavy) and 2 (dense) are handled elsewhere
   if( ci == 3 && cj == 3 ) return;                            // two rank-one variables: GEMM path (schur_rank1_scatter)
   double v = 0.0;
   for( int ei = E.varbeg[i]; ei < E.varbeg[i + 1]; ++ei )
      for( int ej = E.varbeg[j]; ej < E.varbeg[j + 1]; ++ej )
         v += pair_term(E, ei, ej, X, Z);
   M[(size_t)j * ldm + i] = v;
}

// heavy pairs: one CTA per (heavy variable h, other variable o); the CTA splits the entry-pair product
__global__ void __launch_bounds__(256)
schur_heavy_kernel(int m, DevEntries E, const int* __restrict__ heavy, const int* __restrict__ heavylist,
   const double* __restrict__ X, const double* __restrict__ Z, double* __restrict__ M, int ldm, int nranks, int rank)
{
   __shared__ double red[32];
   if( (int)(blockIdx.x % nranks) != rank ) return;               // partner variables are dealt round-robin to the ranks
   const int h = heavylist[blockIdx.y];
   const int o = blockIdx.x;
   if( heavy[o] == 2 ) return;                                    // pairs with a dense variable belong to the dense path
   if( heavy[o] == 1 && o > h ) return;                           // heavy-heavy pairs once
   const int bh = E.varbeg[h], nh = E.varbeg[h + 1] - bh;
   const int bo = E.varbeg[o], no = E.varbeg[o + 1] - bo;
   double v = 0.0;
   const long long total = (long long)nh * no;
   for( long long t = threadIdx.x; t < total; t += blockDim.x )
   {
      int eh = bh + (int)(t / no), eo = bo + (int)(t % no);
      v += pair_term(E, eh, eo, X, Z);
   }
   v = block_sum(v, red);
   if( threadIdx.x == 0 )
   {
      int i = max(h, o), j = min(h, o);
      M[(size_t)j * ldm + i] = v;
   }
}

// ---- dense path ----
__global__ void scatter_dense_kernel(int nd, const int* __restrict__ denselist, DevEntries E, int ld, long long stride, double* __restrict__ Adense)
{
   const int d = blockIdx.x;
   const int j = denselist[d];
   double* Ad = Adense + (size_t)d * stride;
   for( int e = E.varbeg[j] + threadIdx.x; e < E.varbeg[j + 1]; e += blockDim.x )
   {
      int r = E.row[e], c = E.col[e];
      double v = E.val[e];
      Ad[(size_t)c * ld + r] = v;
      Ad[(size_t)r * ld + c] = v;
   }
}

// one warp per (variable i, dense variable d): M[max(i,j), min(i,j)] = sum over entries of A_i in the block of a (U_j(r,c) + U_j(c,r))
__global__ void __launch_bounds__(256)
schur_dense_dots_kernel(int m, int nd, int d0, const int* __restrict__ denselist, const int* __restrict__ cls, DevEntries E, long long blockoff,
   const double* __restrict__ U, int ld, long long stride, double* __restrict__ M, int ldm)
{
   const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
   const int d = blockIdx.y;
   if( i >= m ) return;
   const int j = denselist[d0 + d];
   if( cls[i] == 2 ) return;                      // dense x dense pairs: tensor-core product + schur_dense_scatter
   const double* Uj = U + (size_t)d * stride;
   double s = 0.0;
   for( int e = E.varbeg[i] + lane; e < E.varbeg[i + 1]; e += 32 )
   {
      if( E.off[e] != blockoff ) continue;
      int r = E.row[e], c = E.col[e];
      double u = Uj[(size_t)c * ld + r];
      if( r != c ) u += Uj[(size_t)r * ld + c];
      s += E.val[e] * u;
   }
   s = warp_sum(s);
   if( lane == 0 )
   {
      int a = max(i, j), b = min(i, j);
      M[(size_t)b * ldm + a] = s;
   }
}

// C(a, b) = <A_i, U_j> for dense i = denselist[first + a], dense j = denselist[first_j + b] -> M[i, j] for i >= j (each pair once)
__global__ void schur_dense_scatter_kernel(int count, int cnt, int first, int first_j, const int* __restrict__ denselist,
   const double* __restrict__ C, int ldc, int nslices, long long slicestride, double* __restrict__ M, int ldm)
{
   const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
   if( a >= count || b >= cnt ) return;
   const int i = denselist[first + a], j = denselist[first_j + b];
   if( i < j ) return;
   double v = 0.0;
   for( int sl = 0; sl < nslices; ++sl ) v += C[(size_t)sl * slicestride + (size_t)b * ldc + a];      // k-slices in fixed order
   M[(size_t)j * ldm + i] = v;
}

// M[var_a, var_b] = sigma_a sigma_b G1(a, b) G2(a, b) for the rank-one variables a >= b (list ascending in the variable index)
__global__ void schur_rank1_scatter_kernel(int r, const int* __restrict__ var, const double* __restrict__ sig, const double* __restrict__ G1,
   const double* __restrict__ G2, int ldg, double* __restrict__ M, int ldm)
{
   const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
   if( a >= r || a < b ) return;
   M[(size_t)var[b] * ldm + var[a]] = sig[a] * sig[b] * G1[(size_t)b * ldg + a] * G2[(size_t)b * ldg + a];
}

__global__ void schur_lp_kernel(int nlp, const int* __restrict__ lpbeg, const int* __restrict__ lpind, const double* __restrict__ lpval,
   const double* __restrict__ x, const double* __restrict__ s, double* __restrict__ M, int ldm)
{
   int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
   if( warp >= nlp ) return;
   const int b = lpbeg[warp], cnt = lpbeg[warp + 1] - b;
   const double w = x[warp] / s[warp];
   for( int t = lane; t < cnt * cnt; t += 32 )
   {
      int p = b + t / cnt, q = b + t % cnt;
      int i = lpind[p], j = lpind[q];
      if( i >= j )
         atomicAdd(&M[(size_t)j * ldm + i], w * lpval[p] * lpval[q]);
   }
}

// Deterministic form of the same update (no atomics): the pair (i, j) of row r is handled by the thread of the FIRST row that contains
// both variables; it walks the two column lists (rows ascending, built at upload) once and adds the contributions of all common rows
// in row order.  Every entry of M has one writer, so the Schur complement is bit-reproducible from run to run and across the ranks
// of a sharded solve.  Needs row lists without repeated variables (checked at upload; otherwise the atomic kernel is used).
__global__ void schur_lp_det_kernel(int nlp, const int* __restrict__ lpbeg, const int* __restrict__ lpind, const int* __restrict__ colbeg,
   const int* __restrict__ colrow, const double* __restrict__ colval, const double* __restrict__ x, const double* __restrict__ s,
   double* __restrict__ M, int ldm)
{
   // one THREAD per pair of a row (grid.x = row, grid.y covers the cnt^2 pairs of the longest row): a row with 100 variables is 10^4
   // independent list walks, not 300 rounds of one warp (CLS-syn: 874 -> 20 us per launch)
   const int row = blockIdx.x;
   const int b = lpbeg[row], cnt = lpbeg[row + 1] - b;
   const int t = blockIdx.y * blockDim.x + threadIdx.x;
   if( t >= cnt * cnt ) return;
   const int i = lpind[b + t / cnt], j = lpind[b + t % cnt];
   if( i < j ) return;
   int a = colbeg[i], ae = colbeg[i + 1], c = colbeg[j], ce = colbeg[j + 1];
   double total = 0.0;
   bool first = true;
   while( a < ae && c < ce )
   {
      const int ra = colrow[a], rc = colrow[c];
      if( ra < rc ) ++a;
      else if( rc < ra ) ++c;
      else
      {
         if( first ) { first = false; if( ra != row ) return; }       // the pair belongs to the first row that holds both variables
         total += (x[ra] / s[ra]) * colval[a] * colval[c];
         ++a; ++c;
      }
   }
   if( !first ) M[(size_t)j * ldm + i] += total;
}

__global__ void add_diagonal_kernel(int n, double* __restrict__ A, int lda, double v)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if( i < n ) A[(size_t)i * lda + i] += v;
}

// y = M x with M given by its lower triangle: one warp per row (row part from the row, column part from the column)
__global__ void symv_lower_kernel(int n, const double* __restrict__ M, int ldm, const double* __restrict__ x, double* __restrict__ y)
{
   int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
   if( i >= n ) return;
   double s = 0.0;
   for( int k = lane; k < i; k += 32 ) s += M[(size_t)k * ldm + i] * x[k];          // row i, columns k < i  (strided)
   for( int k = i + lane; k < n; k += 32 ) s += M[(size_t)i * ldm + k] * x[k];      // column i, rows k >= i (contiguous)
   s = warp_sum(s);
   if( lane == 0 ) y[i] = s;
}

// y_i = sum_{k <= i} T[i + k ld] x_k : one thread per row, coalesced across rows for every k
__global__ void trmv_lower_n_kernel(int n, const double* __restrict__ T, int ldt, const double* __restrict__ x, double* __restrict__ y)
{
   __shared__ double xs[256];
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   const int imax = min(n, (blockIdx.x + 1) * (int)blockDim.x);        // largest row of this CTA + 1
   double s0 = 0.0, s1 = 0.0;
   for( int k0 = 0; k0 < imax; k0 += 256 )
   {
      __syncthreads();
      if( k0 + threadIdx.x < n ) xs[threadIdx.x] = x[k0 + threadIdx.x];
      __syncthreads();
      if( i < n )
      {
         int kend = min(256, i + 1 - k0);
         const double* p = T + (size_t)k0 * ldt + i;
         int k = 0;
         for( ; k + 2 <= kend; k += 2 ) { s0 += p[(size_t)k * ldt] * xs[k]; s1 += p[(size_t)(k + 1) * ldt] * xs[k + 1]; }
         if( k < kend ) s0 += p[(size_t)k * ldt] * xs[k];
      }
   }
   if( i < n ) y[i] = s0 + s1;
}

// y_k = sum_{i >= k} T[i + k ld] x_i : one warp per column (contiguous)
__global__ void trmv_lower_t_kernel(int n, const double* __restrict__ T, int ldt, const double* __restrict__ x, double* __restrict__ y)
{
   int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
   if( k >= n ) return;
   const double* col = T + (size_t)k * ldt;
   double s = 0.0;
   for( int i = k + lane; i < n; i += 32 ) s += col[i] * x[i];
   s = warp_sum(s);
   if( lane == 0 ) y[k] = s;
}

__global__ void __launch_bounds__(256)
spmm_pattern_kernel(int n, const double* __restrict__ A, int lda, const double* __restrict__ D, int ldd,
   const int* __restrict__ colptr, const int* __restrict__ rowidx, double alpha, double* __restrict__ Out, int ldo)
{
   const int c = blockIdx.y;
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   const int b = colptr[c], e = colptr[c + 1];
   double s0 = 0.0, s1 = 0.0;
   if( i < n )
   {
      int p = b;
      for( ; p + 2 <= e; p += 2 )
      {
         int r0 = rowidx[p], r1 = rowidx[p + 1];
         s0 += A[(size_t)r0 * lda + i] * D[(size_t)c * ldd + r0];
         s1 += A[(size_t)r1 * lda + i] * D[(size_t)c * ldd + r1];
      }
      if( p < e ) { int r0 = rowidx[p]; s0 += A[(size_t)r0 * lda + i] * D[(size_t)c * ldd + r0]; }
      Out[(size_t)c * ldo + i] = alpha * (s0 + s1);
   }
}

// Sampled product on the aggregate sparsity pattern: W(p, q) = sum_k Pt(k, p) Z(k, q) = (P Z)(p, q) with Pt = P' for the pattern
// entries (p, q) of column q only.  One CTA per column q, one warp per entry (both operand columns contiguous; the column of Z is
// shared by the warps of the CTA and stays in L1).  A(K) only reads K on the pattern, so the n^3 product P Z is not needed for it.
__global__ void __launch_bounds__(256)
sddmm_pattern_kernel(int n, const double* __restrict__ Pt, int ldp, const double* __restrict__ Z, int ldz, const int* __restrict__ colptr,
   const int* __restrict__ rowidx, double* __restrict__ W, int ldw)
{
   const int q = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int b = colptr[q], e = colptr[q + 1];
   const double* zc = Z + (size_t)q * ldz;
   for( int t = b + wid; t < e; t += 8 )
   {
      const int p = rowidx[t];
      const double* pc = Pt + (size_t)p * ldp;
      double s0 = 0.0, s1 = 0.0;
      int k = lane;
      for( ; k + 32 < n; k += 64 ) { s0 += pc[k] * zc[k]; s1 += pc[k + 32] * zc[k + 32]; }
      if( k < n ) s0 += pc[k] * zc[k];
      const double v = warp_sum(s0 + s1);
      if( lane == 0 ) W[(size_t)q * ldw + p] = v;
   }
}

// K(p, q) = (W(p, q) + W(q, p)) / 2 - X(p, q) on the (symmetric) pattern
__global__ void __launch_bounds__(256)
sym_pattern_kernel(int n, const double* __restrict__ W, int ldw, const double* __restrict__ X, int ldx, const int* __restrict__ colptr,
   const int* __restrict__ rowidx, double* __restrict__ K, int ldk)
{
   const int q = blockIdx.x;
   for( int t = colptr[q] + threadIdx.x; t < colptr[q + 1]; t += blockDim.x )
   {
      const int p = rowidx[t];
      K[(size_t)q * ldk + p] = 0.5 * (W[(size_t)q * ldw + p] + W[(size_t)p * ldw + q]) - X[(size_t)q * ldx + p];
   }
}

__global__ void trmv_upper_t_kernel(int n, const double* __restrict__ U, int ldu, const double* __restrict__ x, double* __restrict__ y)
{
   int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
   if( i >= n ) return;
   const double* col = U + (size_t)i * ldu;
   double s0 = 0.0, s1 = 0.0;
   int k = lane;
   for( ; k + 32 <= i; k += 64 ) { s0 += col[k] * x[k]; s1 += col[k + 32] * x[k + 32]; }
   for( ; k <= i; k += 32 ) s0 += col[k] * x[k];
   double s = warp_sum(s0 + s1);
   if( lane == 0 ) y[i] = s;
}

__global__ void transpose_kernel(int n, const double* __restrict__ A, int lda, double* __restrict__ B, int ldb)
{
   __shared__ double tile[32][33];
   int bx = blockIdx.x * 32, by = blockIdx.y * 32;
   for( int r = threadIdx.y; r < 32; r += 8 )
   {
      int i = bx + threadIdx.x, j = by + r;
      tile[r][threadIdx.x] = (i < n && j < n) ? A[(size_t)j * lda + i] : 0.0;
   }
   __syncthreads();
   for( int r = threadIdx.y; r < 32; r += 8 )
   {
      int i = by + threadIdx.x, j = bx + r;      // B(i, j) = A(j, i)
      if( i < n && j < n ) B[(size_t)j * ldb + i] = tile[threadIdx.x][r];
   }
}

__global__ void sym_average_kernel(int n, double* __restrict__ A, int lda, const double* __restrict__ sub)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   int j = blockIdx.y;
   if( i >= n || i < j ) return;
   double v = 0.5 * (A[(size_t)j * lda + i] + A[(size_t)i * lda + j]);
   if( sub != nullptr ) v -= sub[(size_t)j * lda + i];
   A[(size_t)j * lda + i] = v;
   A[(size_t)i * lda + j] = v;
}

__global__ void mirror_lower_kernel(int n, double* __restrict__ A, int lda)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   int j = blockIdx.y;
   if( i >= n || i <= j ) return;
   A[(size_t)i * lda + j] = A[(size_t)j * lda + i];
}

__global__ void axpy_kernel(size_t n, double a, const double* __restrict__ x, double* __restrict__ y)
{
   for( size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x )
      y[i] += a * x[i];
}

__global__ void axpby_kernel(size_t n, double a, const double* __restrict__ x, double b, const double* __restrict__ y, double* __restrict__ out)
{
   for( size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x )
      out[i] = a * x[i] + b * y[i];
}

__global__ void lp_rhs_kernel(int nlp, int corr, double sigmamu, const double* __restrict__ x, const double* __restrict__ s,
   const double* __restrict__ rd, const double* __restrict__ dxa, const double* __restrict__ dsa, double* __restrict__ klp)
{
   int l = blockIdx.x * blockDim.x + threadIdx.x;
   if( l >= nlp ) return;
   double c = -x[l] * rd[l];
   if( corr ) c += sigmamu - dxa[l] * dsa[l];
   klp[l] = c / s[l] - x[l];
}

__global__ void __launch_bounds__(1024)
lp_direction_kernel(int nlp, const double* __restrict__ x, const double* __restrict__ s, const double* __restrict__ rd,
   const double* __restrict__ klp, const double* __restrict__ Ddy, double* __restrict__ dx, double* __restrict__ ds, double* __restrict__ out2)
{
   __shared__ double red[32];
   double ap = -1e300, ad = -1e300;          // we reduce max of -ratio, i.e. min ratio
   for( int l = threadIdx.x; l < nlp; l += blockDim.x )
   {
      double ddy = Ddy[l];
      double vx = klp[l] - x[l] / s[l] * ddy;
      double vs = ddy + rd[l];
      dx[l] = vx; ds[l] = vs;
      if( vx < 0.0 ) ap = fmax(ap, x[l] / vx);       // x/vx is negative; the largest (closest to 0) is the binding ratio
      if( vs < 0.0 ) ad = fmax(ad, s[l] / vs);
   }
   ap = block_max(ap, red);
   ad = block_max(ad, red);
   if( threadIdx.x == 0 ) { out2[0] = (ap > -1e299) ? -ap : 1e30; out2[1] = (ad > -1e299) ? -ad : 1e30; }
}

__global__ void __launch_bounds__(256)
affine_mu_kernel(size_t arena, const double* __restrict__ X, const double* __restrict__ dX, const double* __restrict__ S,
   const double* __restrict__ dS, int nlp, const double* __restrict__ x, const double* __restrict__ dx, const double* __restrict__ s,
   const double* __restrict__ ds, double ap, double ad, double* __restrict__ partials)
{
   __shared__ double red[32];
   double a = 0.0, b = 0.0;
   for( size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < arena; i += (size_t)gridDim.x * blockDim.x )
      a += (X[i] + ap * dX[i]) * (S[i] + ad * dS[i]);
   for( int l = blockIdx.x * blockDim.x + threadIdx.x; l < nlp; l += gridDim.x * blockDim.x )
      b += (x[l] + ap * dx[l]) * (s[l] + ad * ds[l]);
   a = block_sum(a, red); b = block_sum(b, red);
   if( threadIdx.x == 0 ) { partials[blockIdx.x * NSTAT + 9] = a; partials[blockIdx.x * NSTAT + 10] = b; }
}

__global__ void pick_kernel(const double* __restrict__ src, double* __restrict__ dst) { dst[0] = src[0]; }

inline int grid_for(size_t n, int threads) { return (int)std::min<size_t>((n + threads - 1) / threads, 148 * 16); }

} // namespace

#define LAUNCH_END() do { count_launch(); return cudaGetLastError(); } while( 0 )

cudaError_t assemble_positions(cudaStream_t st, int npos, const int* posbeg, const long long* pos, const long long* mirror,
   const int* posvar, const double* posval, const double* posc, const double* y, double cscale, double* T)
{
   if( npos <= 0 ) return cudaSuccess;
   assemble_kernel<<<ceil_div(npos, 256), 256, 0, st>>>(npos, posbeg, pos, mirror, posvar, posval, posc, y, cscale, T);
   LAUNCH_END();
}

cudaError_t residual_matrix(cudaStream_t st, size_t arena, const double* T, const double* S, const double* X, double* Rd, double* partials)
{
   ProfScope prof(st, PROF_ELEM, 32.0 * arena);
   residual_matrix_kernel<<<RED_BLOCKS, 256, 0, st>>>(arena, T, S, X, Rd, partials);
   LAUNCH_END();
}

cudaError_t apply_A(cudaStream_t st, int m, DevEntries E, const double* X, double* out, const int* skipcls)
{
   if( m <= 0 ) return cudaSuccess;
   apply_A_kernel<<<ceil_div(m, 8), 256, 0, st>>>(m, E, skipcls, X, out);
   LAUNCH_END();
}

// dense constraint matrices (expanded n x n copies, matrix d at Ad + d*stride): streaming kernels instead of index gathers
// out[denselist[first + d]] = <A_d, Xk>   (one CTA per matrix, fixed-order reduction)
__global__ void __launch_bounds__(256)
apply_A_dense_kernel(int first, const int* __restrict__ denselist, const double* __restrict__ Ad, long long stride,
   const double* __restrict__ Xk, double* __restrict__ out)
{
   __shared__ double red[32];
   const double* A = Ad + (size_t)blockIdx.x * stride;
   double s0 = 0.0, s1 = 0.0;
   long long e = threadIdx.x;
   for( ; e + 256 < stride; e += 512 ) { s0 += A[e] * Xk[e]; s1 += A[e + 256] * Xk[e + 256]; }
   if( e < stride ) s0 += A[e] * Xk[e];
   const double s = block_sum(s0 + s1, red);
   if( threadIdx.x == 0 ) out[denselist[first + blockIdx.x]] = s;
}

// Tk[e] += sum_d v[denselist[first + d]] * A_d[e]   (one thread per element of the block, coalesced over e)
__global__ void __launch_bounds__(256)
assemble_dense_kernel(int count, int first, const int* __restrict__ denselist, const double* __restrict__ Ad, long long stride,
   const double* __restrict__ v, double* __restrict__ Tk)
{
   __shared__ double vs[256];
   const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   double a0 = 0.0, a1 = 0.0;
   for( int d0 = 0; d0 < count; d0 += 256 )
   {
      const int nd = min(256, count - d0);
      __syncthreads();
      if( (int)threadIdx.x < nd ) vs[threadIdx.x] = v[denselist[first + d0 + threadIdx.x]];
      __syncthreads();
      if( e < stride )
      {
         const double* A = Ad + (size_t)d0 * stride + e;
         int d = 0;
         for( ; d + 1 < nd; d += 2 ) { a0 += vs[d] * A[(size_t)d * stride]; a1 += vs[d + 1] * A[(size_t)(d + 1) * stride]; }
         if( d < nd ) a0 += vs[d] * A[(size_t)d * stride];
      }
   }
   if( e < stride ) Tk[e] += a0 + a1;
}

cudaError_t apply_A_dense(cudaStream_t st, int count, int first, const int* denselist, const double* Ad, long long stride, const double* Xk, double* out)
{
   if( count <= 0 ) return cudaSuccess;
   apply_A_dense_kernel<<<count, 256, 0, st>>>(first, denselist, Ad, stride, Xk, out);
   LAUNCH_END();
}

cudaError_t assemble_dense(cudaStream_t st, int count, int first, const int* denselist, const double* Ad, long long stride, const double* v, double* Tk)
{
   if( count <= 0 ) return cudaSuccess;
   assemble_dense_kernel<<<(unsigned)((stride + 255) / 256), 256, 0, st>>>(count, first, denselist, Ad, stride, v, Tk);
   LAUNCH_END();
}

cudaError_t const_dots(cudaStream_t st, int cnnz, const long long* cpos, const long long* cmirror, const double* cval,
   const double* X, const double* Y, double* out2)
{
   const_dots_kernel<<<1, 1024, 0, st>>>(cnnz, cpos, cmirror, cval, X, Y, out2);
   LAUNCH_END();
}

cudaError_t lp_rows(cudaStream_t st, int nlp, const int* lpbeg, const int* lpind, const double* lpval, const double* lprhs,
   const double* y, const double* x, const double* s, double* Dy, double* rdlp, double* partials)
{
   lp_rows_kernel<<<RED_BLOCKS, 256, 0, st>>>(nlp, lpbeg, lpind, lpval, lprhs, y, x, s, Dy, rdlp, partials);
   LAUNCH_END();
}

cudaError_t lp_cols(cudaStream_t st, int m, const int* colbeg, const int* colrow, const double* colval, const double* x,
   double* out, int accumulate)
{
   if( m <= 0 ) return cudaSuccess;
   lp_cols_kernel<<<ceil_div(m, 256), 256, 0, st>>>(m, colbeg, colrow, colval, x, out, accumulate);
   LAUNCH_END();
}

cudaError_t primal_residual(cudaStream_t st, int m, const double* b, const double* AX, const double* DTx, const double* y,
   double* rp, double* partials)
{
   primal_residual_kernel<<<RED_BLOCKS, 256, 0, st>>>(m, b, AX, DTx, y, rp, partials);
   LAUNCH_END();
}

cudaError_t finalize_partials(cudaStream_t st, const double* partials, int nstats, double* out)
{
   (void)nstats;
   finalize_kernel<<<1, 32, 0, st>>>(partials, RED_BLOCKS, out);
   LAUNCH_END();
}

cudaError_t schur_entries(cudaStream_t st, int m, DevEntries E, const int* heavy, const int* heavylist, int nheavy,
   const double* X, const double* Z, double* M, int ldm, int nranks, int rank)
{
   if( m <= 0 ) return cudaSuccess;
   ProfScope prof(st, PROF_SCHUR, 8.0 * m * (double)m / 2.0);
   dim3 grid(ceil_div(m, 32), ceil_div(m, 8));
   schur_light_kernel<<<grid, 256, 0, st>>>(m, E, heavy, X, Z, M, ldm, nranks, rank);
   count_launch();
   if( nheavy > 0 )
   {
      dim3 g2(m, nheavy);
      schur_heavy_kernel<<<g2, 256, 0, st>>>(m, E, heavy, heavylist, X, Z, M, ldm, nranks, rank);
      count_launch();
   }
   return cudaGetLastError();
}

cudaError_t scatter_dense(cudaStream_t st, int nd, const int* denselist, DevEntries E, int ld, long long stride, double* Adense)
{
   if( nd <= 0 ) return cudaSuccess;
   scatter_dense_kernel<<<nd, 256, 0, st>>>(nd, denselist, E, ld, stride, Adense);
   LAUNCH_END();
}

cudaError_t schur_dense_dots(cudaStream_t st, int m, int nd, int d0, const int* denselist, const int* cls, DevEntries E, long long blockoff,
   const double* U, int ld, long long stride, double* M, int ldm)
{
   if( nd <= 0 || m <= 0 ) return cudaSuccess;
   ProfScope prof(st, PROF_SCHUR, 8.0 * m * (double)nd);
   dim3 grid(ceil_div(m, 8), nd);
   schur_dense_dots_kernel<<<grid, 256, 0, st>>>(m, nd, d0, denselist, cls, E, blockoff, U, ld, stride, M, ldm);
   LAUNCH_END();
}

cudaError_t schur_dense_scatter(cudaStream_t st, int count, int cnt, int first, int first_j, const int* denselist, const double* C, int ldc,
   int nslices, long long slicestride, double* M, int ldm)
{
   if( count <= 0 || cnt <= 0 ) return cudaSuccess;
   dim3 grid(ceil_div(count, 256), cnt);
   schur_dense_scatter_kernel<<<grid, 256, 0, st>>>(count, cnt, first, first_j, denselist, C, ldc, nslices, slicestride, M, ldm);
   LAUNCH_END();
}

cudaError_t schur_rank1_scatter(cudaStream_t st, int r, const int* var, const double* sig, const double* G1, const double* G2, int ldg,
   double* M, int ldm)
{
   if( r <= 0 ) return cudaSuccess;
   ProfScope prof(st, PROF_SCHUR, 24.0 * r * (double)r / 2);
   dim3 grid(ceil_div(r, 256), r);
   schur_rank1_scatter_kernel<<<grid, 256, 0, st>>>(r, var, sig, G1, G2, ldg, M, ldm);
   LAUNCH_END();
}

cudaError_t schur_lp(cudaStream_t st, int nlp, const int* lpbeg, const int* lpind, const double* lpval, const double* x,
   const double* s, double* M, int ldm, const int* colbeg, const int* colrow, const double* colval, int maxcnt)
{
   if( nlp <= 0 ) return cudaSuccess;
   if( colbeg != nullptr && maxcnt > 0 && maxcnt <= 2048 )
   {
      const int threads = (maxcnt * maxcnt >= 128) ? 128 : 32;
      dim3 grid(nlp, ceil_div(maxcnt * maxcnt, threads));
      schur_lp_det_kernel<<<grid, threads, 0, st>>>(nlp, lpbeg, lpind, colbeg, colrow, colval, x, s, M, ldm);
   }
   else
      schur_lp_kernel<<<ceil_div(nlp, 8), 256, 0, st>>>(nlp, lpbeg, lpind, lpval, x, s, M, ldm);
   LAUNCH_END();
}

cudaError_t add_diagonal(cudaStream_t st, int n, double* A, int lda, double v)
{
   if( n <= 0 ) return cudaSuccess;
   add_diagonal_kernel<<<ceil_div(n, 256), 256, 0, st>>>(n, A, lda, v);
   LAUNCH_END();
}

cudaError_t symv_lower(cudaStream_t st, int n, const double* M, int ldm, const double* x, double* y)
{
   if( n <= 0 ) return cudaSuccess;
   symv_lower_kernel<<<ceil_div(n, 8), 256, 0, st>>>(n, M, ldm, x, y);
   LAUNCH_END();
}

cudaError_t trmv_lower(cudaStream_t st, int n, const double* T, int ldt, int trans, const double* x, double* y)
{
   if( n <= 0 ) return cudaSuccess;
   ProfScope prof(st, PROF_TRSV, 4.0 * n * (double)n);
   if( trans ) trmv_lower_t_kernel<<<ceil_div(n, 8), 256, 0, st>>>(n, T, ldt, x, y);
   else trmv_lower_n_kernel<<<ceil_div(n, 256), 256, 0, st>>>(n, T, ldt, x, y);
   LAUNCH_END();
}

cudaError_t spmm_pattern(cudaStream_t st, int n, const double* A, int lda, const double* D, int ldd, const int* colptr,
   const int* rowidx, double alpha, double* Out, int ldo)
{
   if( n <= 0 ) return cudaSuccess;
   ProfScope prof(st, PROF_ELEM, 16.0 * n * (double)n);
   dim3 grid(ceil_div(n, 256), n);
   spmm_pattern_kernel<<<grid, 256, 0, st>>>(n, A, lda, D, ldd, colptr, rowidx, alpha, Out, ldo);
   LAUNCH_END();
}

cudaError_t sddmm_pattern_sym(cudaStream_t st, int n, const double* Pt, int ldp, const double* Z, int ldz, const double* X, int ldx,
   const int* colptr, const int* rowidx, double* W, int ldw, double* K, int ldk)
{
   if( n <= 0 ) return cudaSuccess;
   ProfScope prof(st, PROF_ELEM, 0.0);
   sddmm_pattern_kernel<<<n, 256, 0, st>>>(n, Pt, ldp, Z, ldz, colptr, rowidx, W, ldw);
   sym_pattern_kernel<<<n, 64, 0, st>>>(n, W, ldw, X, ldx, colptr, rowidx, K, ldk);
   count_launch(2);
   return cudaGetLastError();
}

cudaError_t trmv_upper_t(cudaStream_t st, int n, const double* U, int ldu, const double* x, double* y)
{
   if( n <= 0 ) return cudaSuccess;
   ProfScope prof(st, PROF_TRSV, 4.0 * n * (double)n);
   trmv_upper_t_kernel<<<ceil_div(n, 8), 256, 0, st>>>(n, U, ldu, x, y);
   LAUNCH_END();
}

cudaError_t transpose(cudaStream_t st, int n, const double* A, int lda, double* B, int ldb)
{
   if( n <= 0 ) return cudaSuccess;
   ProfScope prof(st, PROF_ELEM, 16.0 * n * (double)n);
   dim3 grid(ceil_div(n, 32), ceil_div(n, 32)), block(32, 8);
   transpose_kernel<<<grid, block, 0, st>>>(n, A, lda, B, ldb);
   LAUNCH_END();
}

cudaError_t sym_average(cudaStream_t st, int n, double* A, int lda, const double* subtract)
{
   if( n <= 0 ) return cudaSuccess;
   ProfScope prof(st, PROF_ELEM, (subtract ? 24.0 : 16.0) * n * (double)n);
   dim3 grid(ceil_div(n, 256), n);
   sym_average_kernel<<<grid, 256, 0, st>>>(n, A, lda, subtract);
   LAUNCH_END();
}

cudaError_t mirror_lower(cudaStream_t st, int n, double* A, int lda)
{
   if( n <= 1 ) return cudaSuccess;
   ProfScope prof(st, PROF_ELEM, 8.0 * n * (double)n);
   dim3 grid(ceil_div(n, 256), n);
   mirror_lower_kernel<<<grid, 256, 0, st>>>(n, A, lda);
   LAUNCH_END();
}

cudaError_t axpy(cudaStream_t st, size_t n, double a, const double* x, double* y)
{
   if( n == 0 ) return cudaSuccess;
   ProfScope prof(st, PROF_ELEM, 24.0 * n);
   axpy_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, a, x, y);
   LAUNCH_END();
}

cudaError_t axpby_out(cudaStream_t st, size_t n, double a, const double* x, double b, const double* y, double* out)
{
   if( n == 0 ) return cudaSuccess;
   ProfScope prof(st, PROF_ELEM, 24.0 * n);
   axpby_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, a, x, b, y, out);
   LAUNCH_END();
}

cudaError_t lp_rhs(cudaStream_t st, int nlp, int corr, double sigmamu, const double* x, const double* s, const double* rd,
   const double* dxa, const double* dsa, double* klp)
{
   if( nlp <= 0 ) return cudaSuccess;
   lp_rhs_kernel<<<ceil_div(nlp, 256), 256, 0, st>>>(nlp, corr, sigmamu, x, s, rd, dxa, dsa, klp);
   LAUNCH_END();
}

cudaError_t lp_direction(cudaStream_t st, int nlp, const double* x, const double* s, const double* rd, const double* klp,
   const double* Ddy, double* dx, double* ds, double* out2)
{
   lp_direction_kernel<<<1, 1024, 0, st>>>(nlp, x, s, rd, klp, Ddy, dx, ds, out2);
   LAUNCH_END();
}

cudaError_t affine_mu(cudaStream_t st, size_t arena, const double* X, const double* dX, const double* S, const double* dS,
   int nlp, const double* x, const double* dx, const double* s, const double* ds, double ap, double ad, double* partials)
{
   ProfScope prof(st, PROF_ELEM, 32.0 * arena);
   affine_mu_kernel<<<RED_BLOCKS, 256, 0, st>>>(arena, X, dX, S, dS, nlp, x, dx, s, ds, ap, ad, partials);
   LAUNCH_END();
}

cudaError_t pick_value(cudaStream_t st, const double* src, double* dst)
{
   pick_kernel<<<1, 1, 0, st>>>(src, dst);
   LAUNCH_END();
}

} // namespace sdpk
