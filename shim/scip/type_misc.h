#ifndef SHIM_SCIP_TYPE_MISC_H
#define SHIM_SCIP_TYPE_MISC_H
#include "blockmemshell/memory.h"
#endif
