/* the two SCIP sorting routines the SCIP-SDP solver bindings use (pub_misc_sort.h); simple insertion/merge implementation */
#ifndef SHIM_SCIP_PUB_MISC_H
#define SHIM_SCIP_PUB_MISC_H
#include "scip/def.h"
#include "blockmemshell/memory.h"
static inline void SCIPsortIntReal(int* key, SCIP_Real* f1, int len)
{
   int i, j;
   for( i = 1; i < len; ++i )
   {
      int k = key[i]; SCIP_Real v = f1[i];
      for( j = i; j > 0 && key[j-1] > k; --j ) { key[j] = key[j-1]; f1[j] = f1[j-1]; }
      key[j] = k; f1[j] = v;
   }
}
static inline void SCIPsortIntIntReal(int* key, int* f1, SCIP_Real* f2, int len)
{
   int i, j;
   for( i = 1; i < len; ++i )
   {
      int k = key[i]; int a = f1[i]; SCIP_Real v = f2[i];
      for( j = i; j > 0 && key[j-1] > k; --j ) { key[j] = key[j-1]; f1[j] = f1[j-1]; f2[j] = f2[j-1]; }
      key[j] = k; f1[j] = a; f2[j] = v;
   }
}
#endif
