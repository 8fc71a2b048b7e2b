// Two-joint verdict tables for statically pruned self-collision pairs (sm_100a).
//
// The robot compiler proves, for link pairs separated by one or two joints, which fine sphere pairs
// can ever touch (gen/<robot>_tables.h: pair_lists).  The relative pose of such a pair depends on
// those joints only, so its verdict is a function of one or two joint values: it is tabulated once per
// robot and device (k_build_pair_table: at every cell centre the minimum clearance c over the group's
// listed sphere pairs, evaluated with the same FK routine the kernels run), with a Lipschitz band
//     band = sum_j lip_j * cell_width_j / 2 + 2e-5          (lip_j: gen/<robot>_pairtab.h)
// cell = 0 "free" (c > band: no listed pair can touch anywhere in the cell), 1 "collides" (c < -band:
// some listed pair overlaps everywhere in the cell), 2 "undecided".  The kernels then read one byte per
// group instead of evaluating the always-on pairs inside the FK routine (Panda: 46 sphere pairs and 23
// sphere poses per state) and instead of queueing records for the gated ones; only undecided states
// (Panda: 0.1 %) and states outside the joint box take the exact path.  Verdicts are unchanged.
#pragma once
#include <cstdint>

namespace vmv
{
    struct PairGroupHost
    {
        int dof[2];       // second = -1: the group depends on one joint
        float lo[2], hi[2];
        float lip[2];     // m / rad (or m / m for a prismatic joint)
        int first, count; // slice of <robot>_pair_group_pairs
    };

    static constexpr int kPairTabMaxGroups = 4;
    static constexpr int kPairTabWords = 11;  // 352 link pairs

    struct PairTabDev
    {
        int n_groups;
        int dof_a[kPairTabMaxGroups], dof_b[kPairTabMaxGroups];
        float lo_a[kPairTabMaxGroups], inv_a[kPairTabMaxGroups], lo_b[kPairTabMaxGroups], inv_b[kPairTabMaxGroups];
        int na[kPairTabMaxGroups], nb[kPairTabMaxGroups];
        const unsigned char *cells[kPairTabMaxGroups];
        uint32_t group_pairs[kPairTabMaxGroups][kPairTabWords];  // link pairs decided by group g
        uint32_t never_pairs[kPairTabWords];                     // link pairs that cannot touch inside the joint box
    };
}  // namespace vmv
