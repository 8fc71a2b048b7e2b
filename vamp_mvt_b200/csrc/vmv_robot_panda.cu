// panda: generated FK + tables (tools/robot_compiler.py, tools/pair_tables.py) and this robot's kernels.
#include "vmv_robot.cuh"

#include "gen/panda_fk.cuh"
#include "gen/panda_tables.h"
#include "gen/panda_pairtab.h"

VMV_DEFINE_ROBOT(panda, PANDA, 0, 128)
