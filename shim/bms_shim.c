/* malloc-backed implementation of the BMS shim with byte accounting (header in front of each allocation) */
#include "blockmemshell/memory.h"
#include <stdio.h>
#include <stdint.h>

#define SHIM_HDR 16   /* keeps 16-byte alignment */
static long long shim_used = 0;   /* atomics via gcc builtins: concurrent SCIP threads own separate solvers but share the allocator */

void* shimBmsAlloc(size_t size, int clear)
{
   unsigned char* raw = (unsigned char*)(clear ? calloc(1, size + SHIM_HDR) : malloc(size + SHIM_HDR));
   if( raw == NULL )
      return NULL;
   *(size_t*)raw = size;
   __atomic_add_fetch(&shim_used, (long long)size, __ATOMIC_RELAXED);
   return raw + SHIM_HDR;
}

void shimBmsFree(void* ptr)
{
   unsigned char* raw;
   if( ptr == NULL )
      return;
   raw = (unsigned char*)ptr - SHIM_HDR;
   __atomic_sub_fetch(&shim_used, (long long)(*(size_t*)raw), __ATOMIC_RELAXED);
   free(raw);
}

void* shimBmsRealloc(void* ptr, size_t size)
{
   unsigned char* raw;
   size_t old;
   if( ptr == NULL )
      return shimBmsAlloc(size, 0);
   raw = (unsigned char*)ptr - SHIM_HDR;
   old = *(size_t*)raw;
   raw = (unsigned char*)realloc(raw, size + SHIM_HDR);
   if( raw == NULL )
      return NULL;
   *(size_t*)raw = size;
   __atomic_add_fetch(&shim_used, (long long)size - (long long)old, __ATOMIC_RELAXED);
   return raw + SHIM_HDR;
}

void* shimBmsDuplicate(const void* src, size_t size)
{
   void* p = shimBmsAlloc(size, 0);
   if( p != NULL && src != NULL )
      memcpy(p, src, size);
   return p;
}

long long BMSgetMemoryUsed(void) { return __atomic_load_n(&shim_used, __ATOMIC_RELAXED); }
long long BMSgetBlockMemoryUsed(const BMS_BLKMEM* blkmem) { (void)blkmem; return BMSgetMemoryUsed(); }
void BMScheckEmptyMemory(void) { if( BMSgetMemoryUsed() != 0 ) fprintf(stderr, "BMS shim: %lld bytes still allocated\n", BMSgetMemoryUsed()); }

/* handles are opaque tokens; they are not counted as used memory (SCIP's accounting does not count them either) */
BMS_BLKMEM* BMScreateBlockMemory(int initchunksize, int garbagefactor) { (void)initchunksize; (void)garbagefactor; return (BMS_BLKMEM*)malloc(8); }
void BMSdestroyBlockMemory(BMS_BLKMEM** blkmem) { if( blkmem != NULL && *blkmem != NULL ) { free(*blkmem); *blkmem = NULL; } }
BMS_BUFMEM* BMScreateBufferMemory(double f, int i, unsigned int c) { (void)f; (void)i; (void)c; return (BMS_BUFMEM*)malloc(8); }
void BMSdestroyBufferMemory(BMS_BUFMEM** bufmem) { if( bufmem != NULL && *bufmem != NULL ) { free(*bufmem); *bufmem = NULL; } }
