// gemm.cu — FP64 tensor-core GEMM for sm_100a.
//
// The FP64 tensor path on Blackwell is the warp-level DMMA.8x8x4 (mma.sync.aligned.m8n8k4.f64); tcgen05/TMEM have no
// f64 kind.  One CTA computes a BM x BN tile of C with 4 warps (2 x 2), each warp a
// (BM/2) x (BN/2) sub-tile out of 8x8 DMMA accumulators; operands are staged through a STAGES-deep cp.async
// (LDGSTS) ring in shared memory, padded so that the fragment loads are bank-conflict free:
//    "MN-contiguous" operand tile  [k][mn]  leading dimension BMN+4  (used for A in NN/NT and for B in NT/TT)
//    "K-contiguous"  operand tile  [mn][k]  leading dimension BK+4   (used for A in TN/TT and for B in NN/TN)
// Edges are handled by zero-filling cp.async (src-size < 16), so any m, n, k works as long as leading dimensions are
// even and base pointers 16-byte aligned.  All matrices column-major; batches via blockIdx.z and element strides.
#include "common.cuh"
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cstring>

namespace sdpk {

thread_local LaunchCounter* g_counter = nullptr;
thread_local Profiler* g_prof = nullptr;

namespace {

constexpr int BK = 16;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int srcbytes)
{
   unsigned s = (unsigned)__cvta_generic_to_shared(smem);
   asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(s), "l"(gmem), "r"(srcbytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }


// loads one operand tile (BMN x BK logical) of stage buffer `s` from global memory
// KCONTIG = false: global element (mn, k) at g[mn + k*ld]   -> smem [k][mn], lds = BMN + 4
// KCONTIG = true : global element (mn, k) at g[k + mn*ld]   -> smem [mn][k], lds = BK + 4
template <int BMN, bool KCONTIG, int NT>
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ g, int ld, int mn0, int k0, int MN, int K, int tid)
{
   if( !KCONTIG )
   {
      constexpr int LDS = BMN + 4;
      constexpr int CPR = BMN / 2;               // 16-byte chunks per k-row
      constexpr int TOTAL = BK * CPR;
#pragma unroll
      for( int c = tid; c < TOTAL; c += NT )
      {
         int k = c / CPR, mn = (c % CPR) * 2;
         int gk = k0 + k, gmn = mn0 + mn;
         int bytes = 0;
         if( gk < K && gmn < MN ) bytes = (MN - gmn >= 2) ? 16 : 8;
         const double* src = bytes ? (g + (size_t)gk * ld + gmn) : g;
         cp_async16(s + k * LDS + mn, src, bytes);
      }
   }
   else
   {
      constexpr int LDS = BK + 4;
      constexpr int CPR = BK / 2;
      constexpr int TOTAL = BMN * CPR;
#pragma unroll
      for( int c = tid; c < TOTAL; c += NT )
      {
         int mn = c / CPR, k = (c % CPR) * 2;
         int gk = k0 + k, gmn = mn0 + mn;
         int bytes = 0;
         if( gk < K && gmn < MN ) bytes = (K - gk >= 2) ? 16 : 8;
         const double* src = bytes ? (g + (size_t)gmn * ld + gk) : g;
         cp_async16(s + mn * LDS + k, src, bytes);
      }
   }
}

template <int BM, int BN, bool TA, bool TB, int STAGES>
__global__ void __launch_bounds__(128)
gemm_dmma_kernel(int M, int N, int K, double alpha, const double* __restrict__ A, int lda, long long strideA,
   const double* __restrict__ B, int ldb, long long strideB, double beta, double* __restrict__ C, int ldc,
   long long strideC, int flags)
{
   constexpr int NT = 128;
   constexpr int WM = BM / 2, WN = BN / 2;          // warp tile
   constexpr int MI = WM / 8, NI = WN / 8;
   // A is "MN-contiguous" when not transposed; B is "MN-contiguous" when transposed
   constexpr int A_ELEMS = TA ? BM * (BK + 4) : BK * (BM + 4);
   constexpr int B_ELEMS = TB ? BK * (BN + 4) : BN * (BK + 4);
   extern __shared__ __align__(16) double smem[];
   double* As = smem;
   double* Bs = smem + STAGES * A_ELEMS;

   // L2-aware rasterisation: the CTAs of a wave (launch order = blockIdx.x fastest) cover bands of GROUP_M tile rows instead of whole
   // tile columns, so that the operand panels a wave streams (GROUP_M row panels of A + the column panels of B it meets) stay
   // inside the 126 MB L2; matters once A alone is larger than L2 (4096^3: DRAM traffic 9x the algorithmic bytes before)
   constexpr int GROUP_M = 16;
   int tm = blockIdx.x, tn = blockIdx.y;
   if( gridDim.x > GROUP_M )
   {
      const int lin = blockIdx.x + gridDim.x * blockIdx.y;
      const int per = GROUP_M * gridDim.y;
      const int grp = lin / per, first = grp * GROUP_M;
      const int gsz = min((int)gridDim.x - first, GROUP_M);
      const int r = lin - grp * per;
      tm = first + r % gsz;
      tn = r / gsz;
   }
   // longest tiles first: with a triangular operand the k range of a tile grows with its row (KHI_M) or column (KHI_N); launched in
   // ascending order the longest tiles would start last and run alone at the end (Linv dX, 2000^3: 160 -> 116 units of tile time in
   // a list-scheduling model, tools/gemm_schedule_model.py)
   if( (flags & GEMM_KHI_M) && !(flags & GEMM_LOWER) ) tm = (int)gridDim.x - 1 - tm;
   else if( (flags & GEMM_KHI_N) && !(flags & GEMM_LOWER) ) tn = (int)gridDim.y - 1 - tn;
   const int m0 = tm * BM, n0 = tn * BN;
   if( (flags & GEMM_LOWER) && (m0 + BM <= n0) )
      return;
   A += (size_t)blockIdx.z * strideA;
   B += (size_t)blockIdx.z * strideB;
   C += (size_t)blockIdx.z * strideC;

   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int gid = lane >> 2, tig = lane & 3;
   const int wm = (warp & 1) * WM, wn = (warp >> 1) * WN;

   double acc[MI][NI][2];
#pragma unroll
   for( int i = 0; i < MI; ++i )
#pragma unroll
      for( int j = 0; j < NI; ++j ) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

   // k-range of this tile: triangular operands only contribute on part of the k axis
   int khi = K, klo = 0;
   if( flags & GEMM_KHI_M ) khi = min(khi, m0 + BM);
   if( flags & GEMM_KHI_N ) khi = min(khi, n0 + BN);
   if( flags & GEMM_KLO_M ) klo = max(klo, m0);
   if( flags & GEMM_KLO_N ) klo = max(klo, n0);
   const int KT0 = klo / BK;
   const int KT = max(KT0, (khi + BK - 1) / BK);
#pragma unroll
   for( int s = 0; s < STAGES - 1; ++s )
   {
      if( KT0 + s < KT )
      {
         load_tile<BM, TA, NT>(As + s * A_ELEMS, A, lda, m0, (KT0 + s) * BK, M, K, tid);
         load_tile<BN, !TB, NT>(Bs + s * B_ELEMS, B, ldb, n0, (KT0 + s) * BK, N, K, tid);
      }
      cp_async_commit();
   }

   for( int kt = KT0; kt < KT; ++kt )
   {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      {
         int nk = kt + STAGES - 1;
         if( nk < KT )
         {
            int s = (nk - KT0) % STAGES;
            load_tile<BM, TA, NT>(As + s * A_ELEMS, A, lda, m0, nk * BK, M, K, tid);
            load_tile<BN, !TB, NT>(Bs + s * B_ELEMS, B, ldb, n0, nk * BK, N, K, tid);
         }
         cp_async_commit();
      }
      const double* as = As + ((kt - KT0) % STAGES) * A_ELEMS;
      const double* bs = Bs + ((kt - KT0) % STAGES) * B_ELEMS;
#pragma unroll
      for( int kk = 0; kk < BK; kk += 4 )
      {
         double a[MI], b[NI];
#pragma unroll
         for( int i = 0; i < MI; ++i )
            a[i] = TA ? as[(wm + i * 8 + gid) * (BK + 4) + kk + tig] : as[(kk + tig) * (BM + 4) + wm + i * 8 + gid];
#pragma unroll
         for( int j = 0; j < NI; ++j )
            b[j] = TB ? bs[(kk + tig) * (BN + 4) + wn + j * 8 + gid] : bs[(wn + j * 8 + gid) * (BK + 4) + kk + tig];
#pragma unroll
         for( int i = 0; i < MI; ++i )
#pragma unroll
            for( int j = 0; j < NI; ++j )
               dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
   }
   cp_async_wait<0>();

#pragma unroll
   for( int i = 0; i < MI; ++i )
   {
      int row = m0 + wm + i * 8 + gid;
      if( row >= M ) continue;
#pragma unroll
      for( int j = 0; j < NI; ++j )
      {
#pragma unroll
         for( int e = 0; e < 2; ++e )
         {
            int col = n0 + wn + j * 8 + tig * 2 + e;
            if( col < N )
            {
               double* p = C + (size_t)col * ldc + row;
               double v = alpha * acc[i][j][e];
               if( beta != 0.0 ) v += beta * (*p);
               *p = v;
            }
         }
      }
   }
}

// ---- TMA-fed variant of the 64 x 64 kernel ------------------------------------------------------------------------------------------
// Same tiling and DMMA fragment code; the operand tiles are fetched by the TMA unit instead of 128 threads issuing cp.async: one
// elected thread arms an mbarrier with the byte count of the stage and issues two cp.async.bulk.tensor.3d (box BM x BK and BN x BK of
// a tensor map over the whole operand, third coordinate = batch index); out-of-range rows / k are zero-filled by the hardware, so
// there is no edge code; the other threads only wait on the mbarrier.  Shared-memory layouts:
//    "K-contiguous" operand (A in TN/TT, B in NN/TN): box [mn][BK], BK * 8 = 128 bytes per row, SWIZZLE_128B: the 16-byte chunk index
//        is XORed with (mn & 7), which makes the DMMA fragment loads (8 rows x 4 k) conflict free without padding;
//    "MN-contiguous" operand (A in NN/NT, B in NT/TT): no swizzle mode covers 512-byte rows, and a dense [k][64] tile would put the
//        four k-rows of a fragment load on the same banks (measured: SYRK 2000 25.8 -> 21.0 TFLOP/s).  The box is therefore 68 wide:
//        four more rows of the operand than the tile needs (zero-filled beyond the matrix), which lands as [k][68] — exactly the
//        padded, conflict-free layout of the cp.async kernel, made by the TMA unit itself.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
   // bounded: a transfer that never completes (a bug) ends in a trap, not in a hung GPU
   for( unsigned spins = 0; ; ++spins )
   {
      unsigned ok;
      asm volatile(
         "{\n"
         ".reg .pred p;\n"
         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
         "selp.u32 %0, 1, 0, p;\n"
         "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
      if( ok ) return;
      if( spins > (1u << 26) ) asm volatile("trap;\n");
   }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar)
{
   asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
      :: "r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

template <bool TA, bool TB, int STAGES>
__global__ void __launch_bounds__(128)
gemm_dmma_tma_kernel(int M, int N, int K, double alpha, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
   double beta, double* __restrict__ C, int ldc, long long strideC, int flags)
{
   constexpr int BM = 64, BN = 64, NT = 128;
   constexpr int WM = BM / 2, WN = BN / 2, MI = WM / 8, NI = WN / 8;
   constexpr int PADW = BM + 4;                            // row length of an MN-contiguous tile
   constexpr int TILE_A = TA ? BM * BK : BK * PADW, TILE_B = TB ? BK * PADW : BN * BK;      // doubles per operand tile
   extern __shared__ __align__(16) unsigned char tsm_raw[];
   __shared__ unsigned long long full[STAGES];
   // SWIZZLE_128B repeats every 1024 bytes of shared memory: the swizzled tiles (8 KB each) come first, on a 1024-byte boundary
   double* tsm = reinterpret_cast<double*>(tsm_raw + ((1024u - (smem_u32(tsm_raw) & 1023u)) & 1023u));
   double* As = (TA || TB) ? tsm : tsm + STAGES * TILE_B;          // A first unless only B is swizzled
   double* Bs = (TA || TB) ? tsm + STAGES * TILE_A : tsm;

   constexpr int GROUP_M = 16;
   int tm = blockIdx.x, tn = blockIdx.y;
   if( gridDim.x > GROUP_M )
   {
      const int lin = blockIdx.x + gridDim.x * blockIdx.y;
      const int per = GROUP_M * gridDim.y;
      const int grp = lin / per, first = grp * GROUP_M;
      const int gsz = min((int)gridDim.x - first, GROUP_M);
      const int r = lin - grp * per;
      tm = first + r % gsz;
      tn = r / gsz;
   }
   // Products of two triangular operands into a lower triangle (S^-1 = L^-T L^-1, the scaled step-length matrices): 528 tiles at
   // n = 2000 whose k ranges differ by a factor of 32, all resident at once - what counts is the SUM of the tiles an SM gets.  The
   // launch has 4 CTAs per SM; the tiles are sorted by work and dealt to the SMs in snake order (CTA b sits on SM b mod #SM when
   // the grid fits the device in one wave), so that every SM gets a long, two middle and a short tile (list-scheduling model:
   // 61 -> 45 tile times, ideal 40).
   if( flags & GEMM_BALANCED )
   {
      const int T = (M + BM - 1) / BM, nsm = (int)gridDim.x / 4;
      const int sidx = (int)blockIdx.x % nsm, q = (int)blockIdx.x / nsm;
      const int p = (q & 1) ? nsm * q + (nsm - 1 - sidx) : nsm * q + sidx;
      if( p >= T * (T + 1) / 2 ) return;
      int k = 1;
      while( k * (k + 1) / 2 <= p ) ++k;
      const int off = p - (k - 1) * k / 2;
      if( flags & GEMM_KHI_N ) { tn = T - k; tm = tn + off; }      // k range 64 (tn + 1): longest first
      else { tm = k - 1; tn = off; }                               // k range K - 64 tm: longest first
   }
   // longest tiles first: with a triangular operand the k range of a tile grows with its row (KHI_M) or column (KHI_N); launched in
   // ascending order the longest tiles would start last and run alone at the end (Linv dX, 2000^3: 160 -> 116 units of tile time in
   // a list-scheduling model, tools/gemm_schedule_model.py)
   if( (flags & GEMM_KHI_M) && !(flags & GEMM_LOWER) ) tm = (int)gridDim.x - 1 - tm;
   else if( (flags & GEMM_KHI_N) && !(flags & GEMM_LOWER) ) tn = (int)gridDim.y - 1 - tn;
   const int m0 = tm * BM, n0 = tn * BN;
   if( (flags & GEMM_LOWER) && (m0 + BM <= n0) )
      return;
   const int z = blockIdx.z;
   C += (size_t)z * strideC;

   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int gid = lane >> 2, tig = lane & 3;
   const int wm = (warp & 1) * WM, wn = (warp >> 1) * WN;

   if( tid == 0 )
   {
#pragma unroll
      for( int s = 0; s < STAGES; ++s ) mbar_init(&full[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
   }
   __syncthreads();

   double acc[MI][NI][2];
#pragma unroll
   for( int i = 0; i < MI; ++i )
#pragma unroll
      for( int j = 0; j < NI; ++j ) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

   int khi = K, klo = 0;
   if( flags & GEMM_KHI_M ) khi = min(khi, m0 + BM);
   if( flags & GEMM_KHI_N ) khi = min(khi, n0 + BN);
   if( flags & GEMM_KLO_M ) klo = max(klo, m0);
   if( flags & GEMM_KLO_N ) klo = max(klo, n0);
   const int KT0 = klo / BK;
   const int KT = max(KT0, (khi + BK - 1) / BK);

   auto issue = [&](int kt)
   {
      const int s = (kt - KT0) % STAGES, k0 = kt * BK;
      mbar_expect_tx(&full[s], (TILE_A + TILE_B) * (unsigned)sizeof(double));
      // A: K-contiguous (TA): map dims (K, M, batch); MN-contiguous: (M, K, batch).  B: K-contiguous (!TB): (K, N, batch); else (N, K, batch)
      if( TA ) tma_load_3d(As + s * TILE_A, &mapA, k0, m0, z, &full[s]); else tma_load_3d(As + s * TILE_A, &mapA, m0, k0, z, &full[s]);
      if( !TB ) tma_load_3d(Bs + s * TILE_B, &mapB, k0, n0, z, &full[s]); else tma_load_3d(Bs + s * TILE_B, &mapB, n0, k0, z, &full[s]);
   };
   if( tid == 0 )
   {
#pragma unroll 1
      for( int s = 0; s < STAGES - 1; ++s ) if( KT0 + s < KT ) issue(KT0 + s);
   }

   for( int kt = KT0; kt < KT; ++kt )
   {
      const int it = kt - KT0, s = it % STAGES;
      mbar_wait(&full[s], (unsigned)((it / STAGES) & 1));
      __syncthreads();                                      // everybody is done with the stage that is refilled next
      if( tid == 0 && kt + STAGES - 1 < KT ) issue(kt + STAGES - 1);
      const double* as = As + s * TILE_A;
      const double* bs = Bs + s * TILE_B;
#pragma unroll
      for( int kk = 0; kk < BK; kk += 4 )
      {
         double a[MI], b[NI];
         const int k = kk + tig;
#pragma unroll
         for( int i = 0; i < MI; ++i )
         {
            const int r = wm + i * 8 + gid;
            a[i] = TA ? as[r * BK + ((((k >> 1) ^ (r & 7)) << 1) | (k & 1))] : as[k * PADW + r];
         }
#pragma unroll
         for( int j = 0; j < NI; ++j )
         {
            const int c = wn + j * 8 + gid;
            b[j] = TB ? bs[k * PADW + c] : bs[c * BK + ((((k >> 1) ^ (c & 7)) << 1) | (k & 1))];
         }
#pragma unroll
         for( int i = 0; i < MI; ++i )
#pragma unroll
            for( int j = 0; j < NI; ++j )
               dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
   }

#pragma unroll
   for( int i = 0; i < MI; ++i )
   {
      int row = m0 + wm + i * 8 + gid;
      if( row >= M ) continue;
#pragma unroll
      for( int j = 0; j < NI; ++j )
      {
#pragma unroll
         for( int e = 0; e < 2; ++e )
         {
            int col = n0 + wn + j * 8 + tig * 2 + e;
            if( col < N )
            {
               double* p = C + (size_t)col * ldc + row;
               double v = alpha * acc[i][j][e];
               if( beta != 0.0 ) v += beta * (*p);
               *p = v;
            }
         }
      }
   }
}

// tensor map of a column-major operand with `inner` contiguous elements per column, `outer` columns, `batch` matrices
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled()
{
   static EncodeTiledFn fn = nullptr;
   static bool tried = false;
   if( !tried )
   {
      tried = true;
      void* p = nullptr;
      cudaDriverEntryPointQueryResult q;
      if( cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess )
         fn = reinterpret_cast<EncodeTiledFn>(p);
   }
   return fn;
}

bool make_map(CUtensorMap* map, const double* base, long long inner, long long outer, int ld, long long stride, int batch, int box_inner,
   int box_outer, bool swizzle128)
{
   EncodeTiledFn fn = encode_tiled();
   if( fn == nullptr ) return false;
   if( (reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld & 1) != 0 || ((stride & 1) != 0 && batch > 1) || (stride <= 0 && batch > 1) ) return false;
   cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)std::max(batch, 1)};
   cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(double), (cuuint64_t)(batch > 1 ? stride : (long long)ld * outer) * sizeof(double)};
   if( strides[1] == 0 ) strides[1] = 16;
   cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer, 1};
   cuuint32_t estr[3] = {1, 1, 1};
   return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 0: cp.async ring (round 1), 1: TMA-fed kernel for the 64 x 64 tiles (default when the tensor maps can be built); SDPCUDA_GEMM=cpasync|tma
int gemm_variant()
{
   const char* e = getenv("SDPCUDA_GEMM");
   if( e != nullptr && strcmp(e, "cpasync") == 0 ) return 0;
   return 1;
}

template <bool TA, bool TB>
cudaError_t launch_tma(cudaStream_t st, int m, int n, int k, double alpha, const double* A, int lda, long long sA,
   const double* B, int ldb, long long sB, double beta, double* C, int ldc, long long sC, int batch, int flags, bool* done)
{
   constexpr int STAGES = 3;          // 51.7 KB per CTA: four CTAs per SM (four stages: three CTAs, and SYRK 2000 loses a wave)
   constexpr size_t SMEM = (size_t)STAGES * ((TA ? 64 * BK : BK * 68) + (TB ? BK * 68 : 64 * BK)) * sizeof(double) + 1024;     // + alignment slack
   *done = false;
   CUtensorMap mapA, mapB;
   // A is (m x k) when not transposed: MN-contiguous, box 64 x BK; transposed: stored (k x m), K-contiguous, box BK x 64
   if( !(TA ? make_map(&mapA, A, k, m, lda, sA, batch, BK, 64, true) : make_map(&mapA, A, m, k, lda, sA, batch, 68, BK, false)) ) return cudaSuccess;
   // B is (k x n) when not transposed: K-contiguous; transposed: stored (n x k), MN-contiguous
   if( !(TB ? make_map(&mapB, B, n, k, ldb, sB, batch, 68, BK, false) : make_map(&mapB, B, k, n, ldb, sB, batch, BK, 64, true)) ) return cudaSuccess;
   auto kern = gemm_dmma_tma_kernel<TA, TB, STAGES>;
   static bool configured[64] = {false};
   int dev = 0;
   SDPK_CUDA_CHECK( cudaGetDevice(&dev) );
   if( !configured[dev & 63] )
   {
      SDPK_CUDA_CHECK( cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM) );
      configured[dev & 63] = true;
   }
   dim3 grid(ceil_div(m, 64), ceil_div(n, 64), batch);
   {
      // balanced dealing of the lower tiles of a triangular x triangular product (see the kernel)
      static int nsm[64] = {0};
      if( nsm[dev & 63] == 0 ) SDPK_CUDA_CHECK( cudaDeviceGetAttribute(&nsm[dev & 63], cudaDevAttrMultiProcessorCount, dev) );
      const int T = ceil_div(m, 64);
      const bool tri = (flags & GEMM_LOWER) && (((flags & GEMM_KHI_N) && !(flags & (GEMM_KHI_M | GEMM_KLO_M | GEMM_KLO_N)))
                                                || ((flags & GEMM_KLO_M) && (flags & GEMM_KLO_N) && !(flags & (GEMM_KHI_M | GEMM_KHI_N))));
      const char* be = getenv("SDPCUDA_GEMM_BALANCE");
      if( tri && batch == 1 && m == n && T >= 8 && T * (T + 1) / 2 <= 4 * nsm[dev & 63] && !(be != nullptr && be[0] == '0') )
      {
         grid = dim3(4 * nsm[dev & 63], 1, 1);
         flags |= GEMM_BALANCED;
      }
   }
   kern<<<grid, 128, SMEM, st>>>(m, n, k, alpha, mapA, mapB, beta, C, ldc, sC, flags);
   count_launch();
   *done = true;
   return cudaGetLastError();
}

template <int BM, int BN, bool TA, bool TB, int STAGES>
cudaError_t launch(cudaStream_t st, int m, int n, int k, double alpha, const double* A, int lda, long long sA,
   const double* B, int ldb, long long sB, double beta, double* C, int ldc, long long sC, int batch, int flags)
{
   constexpr int A_ELEMS = TA ? BM * (BK + 4) : BK * (BM + 4);
   constexpr int B_ELEMS = TB ? BK * (BN + 4) : BN * (BK + 4);
   constexpr size_t SMEM = (size_t)STAGES * (A_ELEMS + B_ELEMS) * sizeof(double);
   auto kern = gemm_dmma_kernel<BM, BN, TA, TB, STAGES>;
   // the opt-in shared-memory size is a per-device function attribute
   static bool configured[64] = {false};
   int dev = 0;
   SDPK_CUDA_CHECK( cudaGetDevice(&dev) );
   if( !configured[dev & 63] )
   {
      SDPK_CUDA_CHECK( cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM) );
      configured[dev & 63] = true;
   }
   dim3 grid(ceil_div(m, BM), ceil_div(n, BN), batch);
   kern<<<grid, 128, SMEM, st>>>(m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, flags);
   count_launch();
   return cudaGetLastError();
}

// ---- register-resident DMMA throughput probe: the measured FP64 tensor roofline denominator -------------------------
__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double* sink)
{
   double acc[16][2];
#pragma unroll
   for( int i = 0; i < 16; ++i ) { acc[i][0] = threadIdx.x * 1e-9; acc[i][1] = i * 1e-9; }
   double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
   for( int it = 0; it < iters; ++it )
   {
#pragma unroll
      for( int i = 0; i < 16; ++i )
         dmma884(acc[i][0], acc[i][1], a, b);
   }
   double s = 0.0;
#pragma unroll
   for( int i = 0; i < 16; ++i ) s += acc[i][0] + acc[i][1];
   if( s == 12345.678 ) sink[0] = s;
}

} // namespace

cudaError_t gemm(cudaStream_t st, bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int lda,
   long long sA, const double* B, int ldb, long long sB, double beta, double* C, int ldc, long long sC, int batch, int flags)
{
   if( m <= 0 || n <= 0 || batch <= 0 )
      return cudaSuccess;
   // algorithmic flop count: triangular operands / symmetric results only need part of the 2mnk product
   double frac = 1.0;
   const bool kflag = (flags & (GEMM_KHI_M | GEMM_KHI_N | GEMM_KLO_M | GEMM_KLO_N)) != 0;
   if( (flags & GEMM_LOWER) && kflag ) frac = 1.0 / 6.0;
   else if( (flags & GEMM_LOWER) || kflag ) frac = 0.5;
   // small problems use 32 x 32 tiles to fill more SMs
   const bool small = ((long long)ceil_div(m, 64) * ceil_div(n, 64) * batch) < 148;
   ProfScope prof(st, small ? PROF_GEMM_SMALL : PROF_GEMM, 2.0 * m * (double)n * k * batch * frac);
#define SDPK_GEMM_DISPATCH(BM, BN, ST) \
   do { \
      if( !ta && !tb ) return launch<BM, BN, false, false, ST>(st, m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch, flags); \
      if( !ta &&  tb ) return launch<BM, BN, false, true,  ST>(st, m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch, flags); \
      if(  ta && !tb ) return launch<BM, BN, true,  false, ST>(st, m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch, flags); \
      return launch<BM, BN, true, true, ST>(st, m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch, flags); \
   } while( 0 )
   if( small )
   {
      // few CTAs (the GEMMs inside the factorisation chains): latency bound, so keep the whole k range in flight with a deep
      // cp.async ring; many small CTAs (batched products): shallower ring, more CTAs per SM
      const long long ctas = (long long)ceil_div(m, 32) * ceil_div(n, 32) * batch;
      if( ctas <= 2 * 148 )
         SDPK_GEMM_DISPATCH(32, 32, 8);
      SDPK_GEMM_DISPATCH(32, 32, 4);
   }
   if( k > 0 && gemm_variant() == 1 )
   {
      bool done = false;
      cudaError_t e;
      if( !ta && !tb ) e = launch_tma<false, false>(st, m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch, flags, &done);
      else if( !ta && tb ) e = launch_tma<false, true>(st, m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch, flags, &done);
      else if( ta && !tb ) e = launch_tma<true, false>(st, m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch, flags, &done);
      else e = launch_tma<true, true>(st, m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch, flags, &done);
      if( e != cudaSuccess || done ) return e;          // otherwise (operands the tensor maps cannot describe): the cp.async kernel
   }
   SDPK_GEMM_DISPATCH(64, 64, 3);
#undef SDPK_GEMM_DISPATCH
}

cudaError_t dmma_peak_probe(cudaStream_t st, int iters, double* d_sink, double* flops, int blocks_per_sm, int threads)
{
   const int blocks = 148 * blocks_per_sm;
   dmma_peak_kernel<<<blocks, threads, 0, st>>>(iters, d_sink);
   count_launch();
   // per warp and iteration: 16 DMMA.8x8x4 = 16 * 8*8*4*2 flop
   *flops = (double)blocks * (threads / 32) * (double)iters * 16.0 * 512.0;
   return cudaGetLastError();
}

} // namespace sdpk
