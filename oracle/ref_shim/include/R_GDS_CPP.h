/* ORACLE / TEST INFRASTRUCTURE ONLY.
 * Stand-in for gdsfmt's R_GDS_CPP.h: the error class, the TRY/CATCH macros and the
 * few GDS_* entry points the hot-path sources call, re-pointed at an in-memory
 * uint8 genotype matrix (oracle/ref_shim/runtime.cpp).  Contains no gdsfmt code. */
#ifndef SHIM_R_GDS_CPP_H
#define SHIM_R_GDS_CPP_H
#include "dType.h"
#include "Rinternals.h"
#ifdef __cplusplus
#include <cstdarg>
#include <cstdio>
#include <exception>
#include <string>

extern "C" {
#endif
typedef void *PdGDSObj;
typedef void *PdGDSFolder;
typedef void *PdAbstractArray;
typedef void *PdGDSFile;
typedef void *PdThreadMutex;
typedef void *PdThreadsSuspending;
typedef void *PdThread;
typedef struct { PdAbstractArray arr; C_Int64 pos; } CdIterator, *PdIterator;
enum C_SVType { svCustom = 0, svInt8 = 5, svUInt8 = 6, svInt16 = 7, svUInt16 = 8, svInt32 = 9,
                svUInt32 = 10, svInt64 = 11, svUInt64 = 12, svFloat32 = 13, svFloat64 = 14 };

int GDS_Array_DimCnt(PdAbstractArray obj);
void GDS_Array_GetDim(PdAbstractArray obj, C_Int32 *out, int n);
void GDS_Array_ReadData(PdAbstractArray obj, const C_Int32 *start, const C_Int32 *length,
                        void *out, enum C_SVType sv);
void GDS_Array_ReadDataEx(PdAbstractArray obj, const C_Int32 *start, const C_Int32 *length,
                          const C_BOOL *const sel[], void *out, enum C_SVType sv);
void GDS_Array_AppendData(PdAbstractArray obj, ssize_t cnt, const void *in, enum C_SVType sv);
int GDS_Attr_Name2Index(PdGDSObj obj, const char *name);
PdGDSObj GDS_Node_Path(PdGDSFolder folder, const char *path, C_BOOL must_exist);
PdGDSObj GDS_R_SEXP2Obj(SEXP obj, C_BOOL readonly);
PdGDSFolder GDS_R_SEXP2FileRoot(SEXP file);
void GDS_Iter_GetStart(PdAbstractArray obj, PdIterator out);
C_Float64 GDS_Iter_GetFloat(PdIterator it);
C_UInt64 GDS_Mach_GetCPULevelCache(int level);
PdThreadMutex GDS_Parallel_InitMutex(void);
void GDS_Parallel_DoneMutex(PdThreadMutex);
void GDS_Parallel_LockMutex(PdThreadMutex);
void GDS_Parallel_UnlockMutex(PdThreadMutex);
PdThreadsSuspending GDS_Parallel_InitSuspend(void);
void GDS_Parallel_DoneSuspend(PdThreadsSuspending);
void GDS_Parallel_Suspend(PdThreadsSuspending);
void GDS_Parallel_WakeUp(PdThreadsSuspending);
void GDS_Parallel_RunThreads(void (*proc)(PdThread, int, void *), void *param, int nthread);
void GDS_SetError(const char *msg);
const char *GDS_GetError(void);
#ifdef __cplusplus
}

namespace CoreArray {
class ErrCoreArray : public std::exception {
public:
    ErrCoreArray() {}
    ErrCoreArray(const char *fmt, ...) {
        va_list ap; va_start(ap, fmt); Init(fmt, ap); va_end(ap);
    }
    ErrCoreArray(const std::string &msg) : fMessage(msg) {}
    virtual const char *what() const throw() { return fMessage.c_str(); }
    virtual ~ErrCoreArray() throw() {}
protected:
    std::string fMessage;
    void Init(const char *fmt, va_list ap) {
        char buf[1024]; vsnprintf(buf, sizeof(buf), fmt, ap); fMessage = buf;
    }
};
struct TdAutoMutex {
    PdThreadMutex m;
    TdAutoMutex(PdThreadMutex x) : m(x) { if (m) GDS_Parallel_LockMutex(m); }
    ~TdAutoMutex() { if (m) GDS_Parallel_UnlockMutex(m); }
    void Reset(PdThreadMutex x) { if (m) GDS_Parallel_UnlockMutex(m); m = x; if (m) GDS_Parallel_LockMutex(m); }
};
}
#define _COREARRAY_ERRMACRO_(x) { va_list args; va_start(args, x); Init(x, args); va_end(args); }
using namespace CoreArray;

/* errors surface as a C++ exception of the shim (caught by the driver) */
struct shim_r_error : public std::exception {
    std::string msg;
    shim_r_error(const std::string &m) : msg(m) {}
    virtual const char *what() const throw() { return msg.c_str(); }
    virtual ~shim_r_error() throw() {}
};
#define COREARRAY_TRY SEXP rv_ans = R_NilValue; bool has_error = false; try {
#define COREARRAY_CATCH } \
    catch (std::exception &E) { GDS_SetError(E.what()); has_error = true; } \
    catch (const char *E) { GDS_SetError(E); has_error = true; } \
    catch (...) { GDS_SetError("unknown error!"); has_error = true; } \
    if (has_error) Rf_error("%s", GDS_GetError()); \
    return rv_ans;
#endif
#endif
