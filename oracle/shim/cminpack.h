/* oracle/shim/cminpack.h -- declarations only, written from the two call sites
 * /root/reference/src/socp/shooting.cpp:803-826 (hybrd) and :830-851 (hybrj) and the callback
 * prototypes at shooting.hpp:282,293.  cminpack itself is an un-vendored, unpinned dependency
 * (src/socp/CMakeLists.txt:11-24); oracle/minpack.c supplies the two symbols.
 * TEST INFRASTRUCTURE ONLY. */
#ifndef SOCP_ORACLE_SHIM_CMINPACK_H
#define SOCP_ORACLE_SHIM_CMINPACK_H
#define __cminpack_func__(f) f
#include "../minpack.h"
#endif
