// TEST INFRASTRUCTURE ONLY: the reference's shape factories (src/impl/vamp/collision/factory.hh), compiled in
// place against the Eigen stand-in of shim/Eigen/Dense, behind C entry points.  tools/make_factory_golden.py
// turns their outputs into tests/golden/factory.npz, which pins vamp_mvt_b200/shapes.py.
#include <vamp/collision/factory.hh>

#include "ref_api.h"

namespace vf = vamp::collision::factory;

extern "C"
{
    // Cuboid(center, euler_xyz, half_extents) (bindings/environment.cc:75-104 -> factory.hh:26-61): the 15 stored fields
    void ref_factory_cuboid(const float *c, const float *e, const float *h, float *out15)
    {
        const auto b = vf::cuboid::flat(c[0], c[1], c[2], e[0], e[1], e[2], h[0], h[1], h[2]);
        const float f[15] = {b.x, b.y, b.z, b.axis_1_x, b.axis_1_y, b.axis_1_z, b.axis_2_x, b.axis_2_y, b.axis_2_z,
                             b.axis_3_x, b.axis_3_y, b.axis_3_z, b.axis_1_r, b.axis_2_r, b.axis_3_r};
        for (int i = 0; i < 15; ++i)
        {
            out15[i] = f[i];
        }
    }

    // Cylinder(center, euler_xyz, radius, length) (factory.hh:160-180): the 8 stored fields
    void ref_factory_cylinder_center(const float *c, const float *e, float radius, float length, float *out8)
    {
        const auto k = vf::cylinder::center::flat(c[0], c[1], c[2], e[0], e[1], e[2], radius, length);
        const float f[8] = {k.x1, k.y1, k.z1, k.xv, k.yv, k.zv, k.r, k.rdv};
        for (int i = 0; i < 8; ++i)
        {
            out8[i] = f[i];
        }
    }

    // Cylinder(endpoint1, endpoint2, radius) (factory.hh:113-124)
    void ref_factory_cylinder_endpoints(const float *p1, const float *p2, float radius, float *out8)
    {
        const auto k = vf::cylinder::endpoints::flat(p1[0], p1[1], p1[2], p2[0], p2[1], p2[2], radius);
        const float f[8] = {k.x1, k.y1, k.z1, k.xv, k.yv, k.zv, k.r, k.rdv};
        for (int i = 0; i < 8; ++i)
        {
            out8[i] = f[i];
        }
    }

    // make_heightfield(center, scaling, dims, data) (factory.hh:365-386): x y z and the INVERSE scales as stored
    void ref_factory_heightfield(const float *c, const float *s, float *out6)
    {
        const std::vector<float> data(4, 0.F);
        const auto h = vf::heightfield::flat(c[0], c[1], c[2], s[0], s[1], s[2], 2, 2, data);
        const float f[6] = {h.x, h.y, h.z, h.xs, h.ys, h.zs};
        for (int i = 0; i < 6; ++i)
        {
            out6[i] = f[i];
        }
    }
}
