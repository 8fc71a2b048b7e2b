// ur5: generated FK + tables (tools/robot_compiler.py, tools/pair_tables.py) and this robot's kernels.
#include "vmv_robot.cuh"

#include "gen/ur5_fk.cuh"
#include "gen/ur5_tables.h"
#include "gen/ur5_pairtab.h"

VMV_DEFINE_ROBOT(ur5, UR5, 0, 128)
