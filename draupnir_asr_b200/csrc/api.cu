// ABI bookkeeping for libdraupnir_b200.so (see include/draupnir_b200.h).
#include "common.cuh"

DRP_API int drp_abi_version(void) { return 1; }
