// Capacity set "large" (peg: 24 contacts, 96 rows, 12 environments in flight per SM) of the Sawyer engine.
#define MJ_CAPSET_LARGE 1
#define mj mjl  // the engine namespace of this translation unit (earl::mjl): no symbol is shared with the small set
#include "earl_mj_rename_large.h"
#include "earl_mj_impl.inc"
