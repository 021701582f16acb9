/* emu_ipm.cpp — TEST INFRASTRUCTURE: csrc/ipm_small.cu compiled for the CPU emulator (once per instantiation: -DSDPK_VARIANT_TINY for
 * the 256-thread one).  KERNEL_INC is the kernel source with its launch functions cut off and the dynamic shared-memory declaration
 * redirected (see the Makefile). */
#include "cuemu.h"
#include KERNEL_INC

#ifdef SDPK_VARIANT_TINY
#define ENTRY cuemu_run_tiny_batch
#else
#define ENTRY cuemu_run_small_batch
#endif

static void block_main(void* arg) { sdpk::ipm_small_batch_kernel(static_cast<const sdpk::SmallArgs*>(arg)); }

extern "C" int ENTRY(int count, const void* descriptors, size_t desc_bytes, size_t stage_bytes)
{
   if( desc_bytes != sizeof(sdpk::SmallArgs) ) return 2;
   return cuemu::run_grid(block_main, const_cast<void*>(descriptors), count, sdpk::NT, sdpk::SMALL_SMEM + stage_bytes);
}
