/* ORACLE / TEST INFRASTRUCTURE ONLY. */
#ifndef SHIM_RDYNLOAD_H
#define SHIM_RDYNLOAD_H
#include "../Rinternals.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct shim_dllinfo DllInfo;
DL_FUNC R_GetCCallable(const char *package, const char *name);
#ifdef __cplusplus
}
#endif
#endif
