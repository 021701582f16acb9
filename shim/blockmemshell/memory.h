/* BMS memory macros of SCIP's blockmemshell/memory.h mapped onto malloc with byte accounting, so that the
 * reference unit tests' leak assertion (checksdpi.c:117, BMSgetMemoryUsed() == 0) can be kept. */
#ifndef SHIM_BLOCKMEMSHELL_MEMORY_H
#define SHIM_BLOCKMEMSHELL_MEMORY_H
#include <stdlib.h>
#include <string.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct BMS_BlkMem BMS_BLKMEM;
typedef struct BMS_BufMem BMS_BUFMEM;

/* implemented in shim/bms_shim.c */
void* shimBmsAlloc(size_t size, int clear);
void* shimBmsRealloc(void* ptr, size_t size);
void  shimBmsFree(void* ptr);
void* shimBmsDuplicate(const void* src, size_t size);
long long BMSgetMemoryUsed(void);
BMS_BLKMEM* BMScreateBlockMemory(int initchunksize, int garbagefactor);
void BMSdestroyBlockMemory(BMS_BLKMEM** blkmem);
BMS_BUFMEM* BMScreateBufferMemory(double arraygrowfac, int arraygrowinit, unsigned int clean);
void BMSdestroyBufferMemory(BMS_BUFMEM** bufmem);
long long BMSgetBlockMemoryUsed(const BMS_BLKMEM* blkmem);
void BMScheckEmptyMemory(void);

#define SHIM_ASSIGN(pp, val)  ( *(void**)(pp) = (val) )
#define SHIM_NBYTES(pp, num)  ( (size_t)((num) > 0 ? (num) : 1) * sizeof(**(pp)) )

#define BMSallocMemory(ptr)                      SHIM_ASSIGN((ptr), shimBmsAlloc(sizeof(**(ptr)), 0))
#define BMSallocClearMemory(ptr)                 SHIM_ASSIGN((ptr), shimBmsAlloc(sizeof(**(ptr)), 1))
#define BMSallocMemoryArray(ptr,num)             SHIM_ASSIGN((ptr), shimBmsAlloc(SHIM_NBYTES(ptr,num), 0))
#define BMSallocClearMemoryArray(ptr,num)        SHIM_ASSIGN((ptr), shimBmsAlloc(SHIM_NBYTES(ptr,num), 1))
#define BMSreallocMemoryArray(ptr,num)           SHIM_ASSIGN((ptr), shimBmsRealloc(*(ptr), SHIM_NBYTES(ptr,num)))
#define BMSduplicateMemory(ptr,source)           SHIM_ASSIGN((ptr), shimBmsDuplicate((source), sizeof(**(ptr))))
#define BMSduplicateMemoryArray(ptr,source,num)  SHIM_ASSIGN((ptr), shimBmsDuplicate((source), SHIM_NBYTES(ptr,num)))
#define BMSfreeMemory(ptr)                       do { shimBmsFree(*(ptr)); *(ptr) = NULL; } while( 0 )
#define BMSfreeMemoryNull(ptr)                   do { if( *(ptr) != NULL ) { shimBmsFree(*(ptr)); *(ptr) = NULL; } } while( 0 )
#define BMSfreeMemoryArray(ptr)                  BMSfreeMemory(ptr)
#define BMSfreeMemoryArrayNull(ptr)              BMSfreeMemoryNull(ptr)
#define BMSclearMemory(ptr)                      memset((void*)(ptr), 0, sizeof(*(ptr)))
#define BMSclearMemoryArray(ptr,num)             memset((void*)(ptr), 0, (size_t)(num) * sizeof(*(ptr)))
#define BMScopyMemory(ptr,source)                memcpy((void*)(ptr), (const void*)(source), sizeof(*(ptr)))
#define BMScopyMemoryArray(ptr,source,num)       do { if( (num) > 0 ) memcpy((void*)(ptr), (const void*)(source), (size_t)(num) * sizeof(*(ptr))); } while( 0 )
#define BMSmoveMemoryArray(ptr,source,num)       memmove((void*)(ptr), (const void*)(source), (size_t)(num) * sizeof(*(ptr)))

#define BMSallocBlockMemory(mem,ptr)                          ( (void)(mem), BMSallocMemory(ptr) )
#define BMSallocClearBlockMemory(mem,ptr)                     ( (void)(mem), BMSallocClearMemory(ptr) )
#define BMSallocBlockMemoryArray(mem,ptr,num)                 ( (void)(mem), BMSallocMemoryArray(ptr,num) )
#define BMSallocClearBlockMemoryArray(mem,ptr,num)            ( (void)(mem), BMSallocClearMemoryArray(ptr,num) )
#define BMSreallocBlockMemoryArray(mem,ptr,oldnum,newnum)     ( (void)(mem), (void)(oldnum), BMSreallocMemoryArray(ptr,newnum) )
#define BMSduplicateBlockMemory(mem,ptr,source)               ( (void)(mem), BMSduplicateMemory(ptr,source) )
#define BMSduplicateBlockMemoryArray(mem,ptr,source,num)      ( (void)(mem), BMSduplicateMemoryArray(ptr,source,num) )
#define BMSfreeBlockMemory(mem,ptr)                           do { (void)(mem); BMSfreeMemory(ptr); } while( 0 )
#define BMSfreeBlockMemoryNull(mem,ptr)                       do { (void)(mem); BMSfreeMemoryNull(ptr); } while( 0 )
#define BMSfreeBlockMemoryArray(mem,ptr,num)                  do { (void)(mem); (void)(num); BMSfreeMemory(ptr); } while( 0 )
#define BMSfreeBlockMemoryArrayNull(mem,ptr,num)              do { (void)(mem); (void)(num); BMSfreeMemoryNull(ptr); } while( 0 )

#define BMSallocBufferMemory(mem,ptr)                         ( (void)(mem), BMSallocMemory(ptr) )
#define BMSallocBufferMemoryArray(mem,ptr,num)                ( (void)(mem), BMSallocMemoryArray(ptr,num) )
#define BMSallocClearBufferMemoryArray(mem,ptr,num)           ( (void)(mem), BMSallocClearMemoryArray(ptr,num) )
#define BMSreallocBufferMemoryArray(mem,ptr,num)              ( (void)(mem), BMSreallocMemoryArray(ptr,num) )
#define BMSduplicateBufferMemoryArray(mem,ptr,source,num)     ( (void)(mem), BMSduplicateMemoryArray(ptr,source,num) )
#define BMSfreeBufferMemory(mem,ptr)                          do { (void)(mem); BMSfreeMemory(ptr); } while( 0 )
#define BMSfreeBufferMemoryNull(mem,ptr)                      do { (void)(mem); BMSfreeMemoryNull(ptr); } while( 0 )
#define BMSfreeBufferMemoryArray(mem,ptr)                     do { (void)(mem); BMSfreeMemory(ptr); } while( 0 )
#define BMSfreeBufferMemoryArrayNull(mem,ptr)                 do { (void)(mem); BMSfreeMemoryNull(ptr); } while( 0 )
#ifdef __cplusplus
}
#endif
#endif
