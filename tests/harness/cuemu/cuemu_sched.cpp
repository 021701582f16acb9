/* cuemu_sched.cpp — TEST INFRASTRUCTURE: cooperative fiber scheduler behind cuemu.h (x86-64 System V only). */
#include "cuemu.h"
#include <cstdlib>
#include <vector>
#include <sys/mman.h>

extern "C" void cuemu_switch(void** from_sp, void* to_sp);
asm(R"(
.text
.globl cuemu_switch
.type cuemu_switch, @function
cuemu_switch:
   pushq %rbp
   pushq %rbx
   pushq %r12
   pushq %r13
   pushq %r14
   pushq %r15
   movq %rsp, (%rdi)
   movq %rsi, %rsp
   popq %r15
   popq %r14
   popq %r13
   popq %r12
   popq %rbx
   popq %rbp
   ret
.size cuemu_switch, .-cuemu_switch
)");

namespace cuemu {

Fiber* cur = nullptr;
int cur_block = 0;
long long barrier_releases = 0;
static void* main_sp = nullptr;
static void (*g_fn)(void*) = nullptr;
static void* g_arg = nullptr;
static std::vector<Fiber> fibers;
static std::vector<uint64_t> slots;           /* [nwarps][2][32] */
static double* g_smem = nullptr;
constexpr size_t STACK = 256 * 1024;
constexpr uint64_t CANARY = 0x7ff8dead7ff8beefULL;

double* dynamic_smem() { return g_smem; }
uint64_t* warp_slots(int phase) { return slots.data() + ((size_t)(cur->tid >> 5) * 2 + phase) * 32; }

static void yield_to_main() { Fiber* f = cur; cuemu_switch(&f->sp, main_sp); }
void block_barrier() { cur->state = 1; yield_to_main(); }
void warp_barrier() { cur->state = 2; yield_to_main(); }

static void trampoline()
{
   g_fn(g_arg);
   cur->state = 3;
   yield_to_main();
   abort();
}

static void prepare(Fiber& f, int tid)
{
   f.tid = tid; f.state = 0; f.shflphase = 0;
   uintptr_t top = ((uintptr_t)f.stack + STACK) & ~(uintptr_t)15;
   uint64_t* sp = (uint64_t*)top;
   *--sp = 0;                                  /* fake return address of the trampoline (keeps rsp = 8 mod 16 at its entry) */
   *--sp = (uint64_t)(uintptr_t)&trampoline;   /* popped by the ret of cuemu_switch */
   for( int r = 0; r < 6; ++r ) *--sp = 0;     /* rbp rbx r12 r13 r14 r15 */
   f.sp = sp;
}

int run_grid(void (*fn)(void*), void* arg, int nblocks, int nthreads, size_t smem_bytes)
{
   g_fn = fn; g_arg = arg;
   if( (int)fibers.size() < nthreads )
   {
      size_t old = fibers.size();
      fibers.resize(nthreads);
      for( size_t i = old; i < fibers.size(); ++i )
      {
         fibers[i].stack = (char*)mmap(nullptr, STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
         if( fibers[i].stack == (char*)MAP_FAILED ) { fprintf(stderr, "cuemu: cannot map a fiber stack\n"); return 1; }
      }
   }
   const int nwarps = (nthreads + 31) / 32;
   slots.assign((size_t)nwarps * 64, 0);
   const size_t nsm = (smem_bytes + 7) / 8, guard = 512;
   std::vector<uint64_t> smem(nsm + guard);
   g_smem = (double*)smem.data();
   int rc = 0;
   /* CUEMU_REVERSE=1: warps and lanes are visited in descending order.  A result that depends on the visiting order means that two
    * threads touch the same location without a barrier in between (the one kind of race this emulator can expose). */
   const char* rev = getenv("CUEMU_REVERSE");
   const bool reverse = (rev != nullptr && rev[0] == '1');
   for( int b = 0; b < nblocks && rc == 0; ++b )
   {
      cur_block = b;
      for( size_t i = 0; i < smem.size(); ++i ) smem[i] = CANARY;          /* uninitialised shared memory reads as NaN */
      for( int t = 0; t < nthreads; ++t ) prepare(fibers[t], t);
      int live = nthreads;
      while( live > 0 )
      {
         bool progress = false;
         for( int wi = 0; wi < nwarps; ++wi )
         {
            const int w = reverse ? nwarps - 1 - wi : wi;
            const int t0 = w * 32, t1 = (t0 + 32 < nthreads) ? t0 + 32 : nthreads;
            for( ; ; )                           /* let the warp run until all its lanes wait at the block barrier or are done */
            {
               bool ran = false;
               for( int ti = t0; ti < t1; ++ti )
               {
                  const int t = reverse ? t1 - 1 - (ti - t0) : ti;
                  if( fibers[t].state == 0 )
                  {
                     cur = &fibers[t];
                     cuemu_switch(&main_sp, cur->sp);
                     if( fibers[t].state == 3 ) --live;
                     ran = true; progress = true;
                  }
               }
               int waitw = 0, alive = 0;
               for( int t = t0; t < t1; ++t ) { if( fibers[t].state != 3 ) ++alive; if( fibers[t].state == 2 ) ++waitw; }
               if( waitw > 0 && waitw == alive ) { for( int t = t0; t < t1; ++t ) if( fibers[t].state == 2 ) fibers[t].state = 0; continue; }
               if( !ran ) break;
            }
         }
         int waitb = 0;
         for( int t = 0; t < nthreads; ++t ) if( fibers[t].state == 1 ) ++waitb;
         if( live > 0 && waitb == live ) { for( int t = 0; t < nthreads; ++t ) if( fibers[t].state == 1 ) fibers[t].state = 0; progress = true; ++barrier_releases; }
         if( !progress )
         {
            fprintf(stderr, "cuemu: deadlock in block %d (%d live threads, %d at the block barrier; divergent barrier or shuffle)\n", b, live, waitb);
            rc = 1;
            break;
         }
      }
      for( size_t i = nsm; i < smem.size(); ++i )
         if( smem[i] != CANARY ) { fprintf(stderr, "cuemu: block %d wrote behind its %zu bytes of dynamic shared memory\n", b, smem_bytes); rc = 1; break; }
   }
   g_smem = nullptr;
   return rc;
}

} // namespace cuemu
