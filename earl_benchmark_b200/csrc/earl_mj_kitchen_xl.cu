// earl_mj_kitchen_xl.cu -- the kitchen engine's extra-large capacity set (544 rows, 48 contacts; 2 environments per block):
// only the redo kernel and its launcher (earl_mjkx_redo_pass), see the header comment of earl_mj_kitchen.cu.
#define MJK_XL 1
#include "earl_mj_kitchen.cu"
