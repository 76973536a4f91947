/* ORACLE / TEST INFRASTRUCTURE ONLY.
 * Minimal stand-in for gdsfmt's dType.h (gdsfmt is not vendored by the reference and
 * not installed in this image): fixed-width typedefs, SIMD feature macros derived
 * from the compiler, empty export decorations.  Written from the names the
 * reference's hot-path sources use; contains no gdsfmt code. */
#ifndef SHIM_DTYPE_H
#define SHIM_DTYPE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <math.h>
#endif
typedef int8_t C_Int8;
typedef uint8_t C_UInt8;
typedef int16_t C_Int16;
typedef uint16_t C_UInt16;
typedef int32_t C_Int32;
typedef uint32_t C_UInt32;
typedef int64_t C_Int64;
typedef uint64_t C_UInt64;
typedef float C_Float32;
typedef double C_Float64;
typedef int8_t C_BOOL;
#define COREARRAY_DLL_LOCAL
#define COREARRAY_DLL_EXPORT
#define COREARRAY_DLL_DEFAULT
#define COREARRAY_INLINE inline
#define COREARRAY_CALL_ALIGN
#define COREARRAY_POSIX_THREAD 1
#define COREARRAY_PLATFORM_UNIX 1
#define COREARRAY_SIMD_ATTR_ALIGN __attribute__((aligned(32)))
#ifdef __SSE__
#define COREARRAY_SIMD_SSE 1
#endif
#ifdef __SSE2__
#define COREARRAY_SIMD_SSE2 1
#endif
#ifdef __SSE4_1__
#define COREARRAY_SIMD_SSE4_1 1
#endif
#ifdef __SSE4_2__
#define COREARRAY_SIMD_SSE4_2 1
#endif
#ifdef __AVX__
#define COREARRAY_SIMD_AVX 1
#endif
#ifdef __AVX2__
#define COREARRAY_SIMD_AVX2 1
#endif
#endif
