// fetch: generated FK + tables (tools/robot_compiler.py, tools/pair_tables.py) and this robot's kernels.
#include "vmv_robot.cuh"

#include "gen/fetch_fk.cuh"
#include "gen/fetch_tables.h"
#include "gen/fetch_pairtab.h"

VMV_DEFINE_ROBOT(fetch, FETCH, 0, 128)
