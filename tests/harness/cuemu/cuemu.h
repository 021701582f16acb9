/* cuemu.h — TEST INFRASTRUCTURE: runs the source of a single-CTA CUDA kernel on the CPU.
 *
 * Every CUDA thread of a block is a fiber (own stack, hand-written context switch); __syncthreads() and the warp primitives
 * (__syncwarp, __shfl_sync, __shfl_xor_sync) are barriers of a cooperative scheduler, blocks run one after the other.  This is
 * how tests/test_cuemu_batch.py executes csrc/ipm_small.cu (the frontier-batch kernel and its 256-thread instantiation) without a
 * GPU: same source, same shared-memory layout, same descriptors as sdpcuda_solve_batch builds them.  It checks the LOGIC of the
 * kernel (indexing, in-kernel cold start, descriptor staging, work-space layout); it cannot find data races, and floating-point
 * results differ from the device in the last bits (no FMA contraction rules, host rsqrt).  Never linked into the product. */
#pragma once
#include <cuda_runtime.h>      /* host-side types only (cudaStream_t, cudaError_t); its qualifier macros are replaced below */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace cuemu {

struct Fiber
{
   void* sp;
   int tid;
   int state;              /* 0 runnable, 1 waiting at the block barrier, 2 waiting at a warp barrier, 3 done */
   char* stack;
   int shflphase;
};
extern Fiber* cur;
extern int cur_block;
void block_barrier();
void warp_barrier();
uint64_t* warp_slots(int phase);          /* 32 exchange slots of the calling fiber's warp */
double* dynamic_smem();
/* runs fn() on nthreads fibers for blocks 0..nblocks-1; smem_bytes of dynamic shared memory (guarded); returns 0, or 1 on deadlock /
 * shared-memory overrun */
int run_grid(void (*fn)(void*), void* arg, int nblocks, int nthreads, size_t smem_bytes);

template <class T> inline T shfl(T v, int src)
{
   static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
   Fiber* f = cur;
   const int ph = f->shflphase++ & 1;
   uint64_t bits = 0;
   memcpy(&bits, &v, sizeof(T));
   warp_slots(ph)[f->tid & 31] = bits;
   warp_barrier();
   bits = warp_slots(ph)[src & 31];
   T r;
   memcpy(&r, &bits, sizeof(T));
   return r;
}

/* every lane of the warp takes part (the kernels only use full masks) */
inline unsigned ballot(bool pred)
{
   Fiber* f = cur;
   const int ph = f->shflphase++ & 1;
   warp_slots(ph)[f->tid & 31] = pred ? 1u : 0u;
   warp_barrier();
   unsigned m = 0;
   for( int l = 0; l < 32; ++l ) if( warp_slots(ph)[l] != 0 ) m |= 1u << l;
   return m;
}

struct IdxProxy { int which; operator int() const { return which == 0 ? cur->tid : cur_block; } };
struct Idx3 { IdxProxy x; };

} // namespace cuemu

static const cuemu::Idx3 threadIdx = {{0}};
static const cuemu::Idx3 blockIdx = {{1}};

#undef __device__
#undef __global__
#undef __host__
#undef __forceinline__
#undef __noinline__
#undef __shared__
#undef __launch_bounds__
#undef __align__
#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __syncthreads() cuemu::block_barrier()
#define __syncwarp() cuemu::warp_barrier()
#define __shfl_sync(mask, v, src) cuemu::shfl((v), (src))
#define __shfl_xor_sync(mask, v, o) cuemu::shfl((v), (cuemu::cur->tid & 31) ^ (o))
#define __ballot_sync(mask, pred) cuemu::ballot((pred))
#define __ffs(x) __builtin_ffs((int)(x))
namespace cuemu { extern long long barrier_releases; }
/* the kernel's own phase profile (verbose >= 2) then counts block-barrier releases per phase instead of cycles */
static inline long long clock64() { return cuemu::barrier_releases; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
