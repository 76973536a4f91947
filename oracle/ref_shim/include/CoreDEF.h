/* ORACLE / TEST INFRASTRUCTURE ONLY: stand-in for gdsfmt's CoreDEF.h. */
#ifndef SHIM_COREDEF_H
#define SHIM_COREDEF_H
#include "dType.h"
#endif
