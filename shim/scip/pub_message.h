/* message macros: errors to stderr, info to stdout, debug output only with -DSCIP_DEBUG */
#ifndef SHIM_SCIP_PUB_MESSAGE_H
#define SHIM_SCIP_PUB_MESSAGE_H
#include <stdio.h>
#include <stdlib.h>
#include "scip/type_message.h"
#define SCIPerrorMessage(...) do { fprintf(stderr, "[%s:%d] ERROR: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); } while( 0 )
#ifdef SCIP_DEBUG
#define SCIPdebugMessage(...) do { printf("[%s:%d] debug: ", __FILE__, __LINE__); printf(__VA_ARGS__); } while( 0 )
#define SCIPdebugPrintf(...)  printf(__VA_ARGS__)
#define SCIPdebug(x) x
#else
#define SCIPdebugMessage(...) while( 0 ) printf(__VA_ARGS__)
#define SCIPdebugPrintf(...)  while( 0 ) printf(__VA_ARGS__)
#define SCIPdebug(x)
#endif
/* honour SHIM_QUIET=1 so that test runs stay readable */
static inline int shimMessagesQuiet(void) { static int q = -1; if( q < 0 ) { const char* e = getenv("SHIM_QUIET"); q = (e != NULL && e[0] == '1'); } return q; }
#define SCIPmessagePrintInfo(hdlr, ...) do { (void)(hdlr); if( !shimMessagesQuiet() ) printf(__VA_ARGS__); } while( 0 )
#define SCIPmessagePrintWarning(hdlr, ...) do { (void)(hdlr); if( !shimMessagesQuiet() ) printf(__VA_ARGS__); } while( 0 )
#endif
