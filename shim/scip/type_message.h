#ifndef SHIM_SCIP_TYPE_MESSAGE_H
#define SHIM_SCIP_TYPE_MESSAGE_H
typedef struct SCIP_Messagehdlr SCIP_MESSAGEHDLR;
#endif
