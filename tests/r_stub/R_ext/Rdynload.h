#ifndef SGM_R_STUB_RDYNLOAD_H
#define SGM_R_STUB_RDYNLOAD_H
#include "../Rinternals.h"
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct DllInfo DllInfo;
int R_registerRoutines(DllInfo* info, const void* c, const R_CallMethodDef* call, const void* f, const void* e);
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value);
extern const R_CallMethodDef* rstub_registered;   /* what R_init_<pkg> registered */
#endif
