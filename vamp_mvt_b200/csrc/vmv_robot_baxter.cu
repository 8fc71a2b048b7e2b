// baxter: generated FK + tables (tools/robot_compiler.py, tools/pair_tables.py) and this robot's kernels.
#include "vmv_robot.cuh"

#include "gen/baxter_fk.cuh"
#include "gen/baxter_tables.h"
#include "gen/baxter_pairtab.h"

VMV_DEFINE_ROBOT(baxter, BAXTER, 1, 64)
