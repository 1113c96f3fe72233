// Capacity set "small" (door: 16 contacts, 64 rows, 16 environments in flight per SM) of the Sawyer engine.
#define MJ_CAPSET_SMALL 1
#include "earl_mj_rename_small.h"
#include "earl_mj_impl.inc"
