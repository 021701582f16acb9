/* devctx.h — one lazily created device handle per host thread for the entry points of the SCIP-SDP interfaces that carry
 * no solver object (SCIPlapack*, SCIPsdpSolcheckerCheck).  Concurrent SCIP threads get separate handles/streams. */
#ifndef SDPI_DEVCTX_H
#define SDPI_DEVCTX_H
#include "sdpcuda.h"
sdpcuda_handle* sdpiCudaThreadHandle(void);
#endif
