// mj_collide.cuh -- collision detection and contact constraint rows (stage 1: no contacts yet).
#pragma once

#include "mj_engine.cuh"

namespace earl {
namespace mj {

template <int NL>
MJ_HD void collide(const Model& m, const real* hull, Work& w, int lane) {
  (void)m; (void)hull; (void)lane;
  w.ncon = 0;
}

template <int NL>
MJ_HD void contact_rows(const Model& m, Work& w, int lane) {
  (void)m; (void)w; (void)lane;
}

}  // namespace mj
}  // namespace earl
