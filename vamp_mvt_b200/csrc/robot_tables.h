// Plain-data records shared by the generated robot tables (gen/*_tables.h), the kernels and the host.
#pragma once
#include <cstdint>

namespace vmv
{
    // One sphere to sweep against the environment.  Centre is expressed in the frame of rigid body
    // `body` (0 = world-fixed).  A bounding-sphere task has skip >= 0 (index of the task to continue
    // with when the bounding sphere is clear) and sphere = -1; a fine task has skip = -1.
    struct SphereTask
    {
        float cx, cy, cz, r;
        int body;
        int skip;
        int link;
        int sphere;
    };

    struct LinkInfo
    {
        int first_sphere;
        int n_spheres;
        int body;
        int bound_task;
    };

    struct LinkPair
    {
        int a, b;
    };

    // Statically pruned fine-sphere pairs of a link pair (robot compiler: pairs that can touch for
    // some joint values inside the joint limits); count < 0: no list, use the full cross product.
    struct PairInfo
    {
        int offset, count, inline_checked;
    };

    struct SpherePair
    {
        uint8_t task_a, task_b;
    };
}  // namespace vmv
