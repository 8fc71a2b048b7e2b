// TEST INFRASTRUCTURE ONLY: C entry points of oracle/_ref/libvamp_ref.so.
// Environment construction mirrors the reference's Python-facing adders
// (src/impl/vamp/bindings/environment.cc:111-181): same containers, same z-aligned
// classification, re-sort after every insertion.  Shapes are built through the reference's own
// constructors (collision/shapes.hh) so min_distance is the reference's value.
#include "ref_robot.hh"
#include <vamp/collision/filter_centervox.hh>
#include "ref_api.h"

#include <cmath>
#include <cstring>
#include <memory>

namespace vc = vamp::collision;

extern const refh::RobotVTable ref_vt_panda;
extern const refh::RobotVTable ref_vt_ur5;
extern const refh::RobotVTable ref_vt_fetch;
extern const refh::RobotVTable ref_vt_baxter;

namespace
{
    struct RefEnv
    {
        refh::EnvF env;
        int next_id = 0;

        auto name() -> std::string
        {
            return std::to_string(next_id++);
        }
    };

    auto vt(int robot) -> const refh::RobotVTable &
    {
        switch (robot)
        {
            case 0:
                return ref_vt_panda;
            case 1:
                return ref_vt_ur5;
            case 2:
                return ref_vt_fetch;
            default:
                return ref_vt_baxter;
        }
    }

    auto E(void *p) -> RefEnv &
    {
        return *static_cast<RefEnv *>(p);
    }

    auto to_points(const float *pts, size_t n) -> std::vector<vc::Point>
    {
        std::vector<vc::Point> v(n);
        for (size_t i = 0; i < n; ++i)
        {
            v[i] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
        }
        return v;
    }
}  // namespace

extern "C"
{
    void *ref_env_create(void)
    {
        return new RefEnv();
    }

    void ref_env_destroy(void *env)
    {
        delete static_cast<RefEnv *>(env);
    }

    void ref_env_add_sphere(void *env, const float *f)
    {
        vc::Sphere<float> s(f[0], f[1], f[2], f[3]);
        s.name = E(env).name();
        E(env).env.spheres.emplace_back(s);
        E(env).env.sort();
    }

    void ref_env_add_cuboid(void *env, const float *f)
    {
        vc::Cuboid<float> s(
            f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8], f[9], f[10], f[11], f[12], f[13], f[14]);
        s.name = E(env).name();
        if (s.axis_3_z == 1.)
        {
            E(env).env.z_aligned_cuboids.emplace_back(s);
        }
        else
        {
            E(env).env.cuboids.emplace_back(s);
        }
        E(env).env.sort();
    }

    void ref_env_add_capsule(void *env, const float *f)
    {
        vc::Cylinder<float> s(f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7]);
        s.name = E(env).name();
        if (s.xv == 0. and s.yv == 0.)
        {
            E(env).env.z_aligned_capsules.emplace_back(s);
        }
        else
        {
            E(env).env.capsules.emplace_back(s);
        }
        E(env).env.sort();
    }

    void ref_env_add_heightfield(void *env, const float *f, size_t xd, size_t yd, const float *data)
    {
        std::vector<float> d(data, data + xd * yd);
        vc::HeightField<float> h(f[0], f[1], f[2], f[3], f[4], f[5], xd, yd, d);
        h.name = E(env).name();
        E(env).env.heightfields.emplace_back(h);
    }

    // the reference's own CAPT, opened up for tests/test_capt_predicate.py: depth, split values, the list of one leaf
    int ref_capt_nlog2(void *env, size_t which)
    {
        return E(env).env.capt_pointclouds[which].nlog2;
    }

    const float *ref_capt_tests(void *env, size_t which)
    {
        return E(env).env.capt_pointclouds[which].tests.data();
    }

    size_t ref_capt_leaf_list(void *env, size_t which, size_t leaf, float *out_xyz, size_t cap)
    {
        const auto &t = E(env).env.capt_pointclouds[which];
        size_t n = 0;
        for (uint32_t i = t.aff_starts[leaf]; i < t.aff_starts[leaf + 1]; ++i)
        {
            const auto xs = t.affordances[0][i].to_array(), ys = t.affordances[1][i].to_array(), zs = t.affordances[2][i].to_array();
            for (size_t l = 0; l < xs.size(); ++l)
            {
                if (!std::isfinite(xs[l]))
                {
                    continue;
                }
                if (n < cap)
                {
                    out_xyz[3 * n] = xs[l], out_xyz[3 * n + 1] = ys[l], out_xyz[3 * n + 2] = zs[l];
                }
                ++n;
            }
        }
        return n;
    }

    void ref_env_add_capt(void *env, const float *pts, size_t n, float r_min, float r_max, float r_point)
    {
        E(env).next_id++;
        E(env).env.capt_pointclouds.emplace_back(to_points(pts, n), r_min, r_max, r_point);
    }

    void ref_env_add_mvt(
        void *env,
        const float *pts,
        size_t n,
        float r_min,
        float r_max,
        const float *aabb_min,
        const float *aabb_max,
        float r_point)
    {
        E(env).next_id++;
        vc::Point lo{aabb_min[0], aabb_min[1], aabb_min[2]};
        vc::Point hi{aabb_max[0], aabb_max[1], aabb_max[2]};
        E(env).env.mvt_pointclouds.emplace_back(to_points(pts, n), r_min, r_max, lo, hi, r_point);
    }

    void ref_env_attach(void *env, const float *tf12, const float *s, size_t n)
    {
        Eigen::Isometry3f tf;
        for (int i = 0; i < 3; ++i)
        {
            tf.translation()[i] = tf12[i];
        }
        for (int i = 0; i < 9; ++i)
        {
            tf.linear().d[i] = tf12[3 + i];
        }

        vc::Attachment<float> a(tf);
        for (size_t i = 0; i < n; ++i)
        {
            a.spheres.emplace_back(vc::Sphere<float>(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]));
        }
        E(env).env.attachments.emplace(a);
    }

    void ref_env_detach(void *env)
    {
        E(env).env.attachments.reset();
    }

    size_t ref_env_dump(void *env, int kind, float *out, size_t cap)
    {
        const auto &e = E(env).env;
        size_t k = 0;
        auto put = [&](float v)
        {
            if (k < cap)
            {
                out[k] = v;
            }
            ++k;
        };

        switch (kind)
        {
            case 0:
                for (const auto &s : e.spheres)
                {
                    put(s.x), put(s.y), put(s.z), put(s.r), put(s.min_distance);
                }
                return e.spheres.size();
            case 1:
            case 2:
            {
                const auto &v = kind == 1 ? e.capsules : e.z_aligned_capsules;
                for (const auto &s : v)
                {
                    put(s.x1), put(s.y1), put(s.z1), put(s.xv), put(s.yv), put(s.zv), put(s.r), put(s.rdv),
                        put(s.min_distance);
                }
                return v.size();
            }
            default:
            {
                const auto &v = kind == 3 ? e.cuboids : e.z_aligned_cuboids;
                for (const auto &s : v)
                {
                    put(s.x), put(s.y), put(s.z);
                    put(s.axis_1_x), put(s.axis_1_y), put(s.axis_1_z);
                    put(s.axis_2_x), put(s.axis_2_y), put(s.axis_2_z);
                    put(s.axis_3_x), put(s.axis_3_y), put(s.axis_3_z);
                    put(s.axis_1_r), put(s.axis_2_r), put(s.axis_3_r);
                    put(s.min_distance);
                }
                return v.size();
            }
        }
    }

    int ref_robot_dim(int robot)
    {
        return vt(robot).dim;
    }

    int ref_robot_n_spheres(int robot)
    {
        return vt(robot).n_spheres;
    }

    int ref_robot_resolution(int robot)
    {
        return vt(robot).resolution;
    }

    void ref_validate_configs(int robot, void *env, const float *q, size_t n, uint8_t *out, int threads)
    {
        vt(robot).validate_configs(E(env).env, q, n, out, threads);
    }

    size_t ref_simplify(
        int robot,
        void *env,
        const float *path,
        size_t n,
        const int *ops,
        size_t n_ops,
        const float *settings12,
        const float *samples,
        size_t n_samples,
        float *out,
        size_t cap,
        size_t *iterations)
    {
        return vt(robot).simplify(E(env).env, path, n, ops, n_ops, settings12, samples, n_samples, out, cap, iterations);
    }

    // vamp::collision::filter_pointcloud_centervox (collision/filter_centervox.hh:302-333) on [n][3] points;
    // writes the kept points (up to cap) and returns their number, or (size_t)-1 where the reference throws
    size_t ref_filter_centervox(
        const float *pts,
        size_t n,
        float voxel_size,
        float max_range,
        const float *origin,
        const float *ws_min,
        const float *ws_max,
        float *out_xyz,
        size_t cap)
    {
        std::vector<vamp::collision::Point> pc(n);
        for (size_t i = 0; i < n; ++i)
        {
            pc[i] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
        }
        try
        {
            const auto kept = vamp::collision::filter_pointcloud_centervox(
                pc,
                voxel_size,
                max_range,
                vamp::collision::Point{origin[0], origin[1], origin[2]},
                vamp::collision::Point{ws_min[0], ws_min[1], ws_min[2]},
                vamp::collision::Point{ws_max[0], ws_max[1], ws_max[2]});
            for (size_t i = 0; i < kept.size() and i < cap; ++i)
            {
                out_xyz[3 * i] = kept[i][0], out_xyz[3 * i + 1] = kept[i][1], out_xyz[3 * i + 2] = kept[i][2];
            }
            return kept.size();
        }
        catch (const std::exception &)
        {
            return static_cast<size_t>(-1);
        }
    }

    size_t ref_path_op(int robot, int op, const float *path, size_t n, size_t arg, float *out, size_t cap, float *cost)
    {
        return vt(robot).path_op(op, path, n, arg, out, cap, cost);
    }

    size_t ref_planner(int robot, int which, void *env, const float *start, const float *goal, size_t max_iterations, size_t max_samples,
                       size_t batch_size, float *verts, size_t cap_v, uint32_t *edges, size_t cap_e, size_t *n_edges, size_t *iterations, float *cost)
    {
        return vt(robot).planner(which, E(env).env, start, goal, max_iterations, max_samples, batch_size, verts, cap_v, edges, cap_e, n_edges, iterations,
                                 cost);
    }

    void ref_halton(int robot, size_t skip, size_t n, float *out)
    {
        vt(robot).halton(skip, n, out);
    }

    void
    ref_validate_edges(int robot, void *env, const float *a, const float *b, size_t n, uint8_t *out, int threads)
    {
        vt(robot).validate_edges(E(env).env, a, b, n, out, threads);
    }

    void ref_sphere_fk(int robot, const float *q, size_t n, float *out)
    {
        vt(robot).sphere_fk(q, n, out);
    }

    void ref_eefk(int robot, const float *q, float *out16)
    {
        vt(robot).eefk(q, out16);
    }

    void ref_filter_points(int robot, void *env, const float *q, const float *pts, size_t n, float point_radius, uint8_t *keep)
    {
        vt(robot).filter_points(E(env).env, q, pts, n, point_radius, keep);
    }

    void ref_debug(
        int robot,
        void *env,
        const float *q,
        int32_t *env_hits,
        size_t cap_env,
        size_t *n_env,
        int32_t *self_hits,
        size_t cap_self,
        size_t *n_self)
    {
        std::vector<std::pair<int, int>> eh, sh;
        vt(robot).debug(E(env).env, q, eh, sh);
        *n_env = eh.size();
        *n_self = sh.size();
        for (size_t i = 0; i < eh.size() and i < cap_env; ++i)
        {
            env_hits[2 * i] = eh[i].first;
            env_hits[2 * i + 1] = eh[i].second;
        }
        for (size_t i = 0; i < sh.size() and i < cap_self; ++i)
        {
            self_hits[2 * i] = sh[i].first;
            self_hits[2 * i + 1] = sh[i].second;
        }
    }

    double ref_time_configs(int robot, void *env, const float *q, size_t n, int threads, int reps)
    {
        std::vector<uint8_t> out(n);
        double best = 1e300;
        for (int r = 0; r < reps; ++r)
        {
            const auto t0 = std::chrono::steady_clock::now();
            vt(robot).validate_configs(E(env).env, q, n, out.data(), threads);
            const auto t1 = std::chrono::steady_clock::now();
            best = std::min(best, std::chrono::duration<double>(t1 - t0).count());
        }
        return best;
    }

    double
    ref_time_edges(int robot, void *env, const float *a, const float *b, size_t n, int threads, int reps)
    {
        std::vector<uint8_t> out(n);
        double best = 1e300;
        for (int r = 0; r < reps; ++r)
        {
            const auto t0 = std::chrono::steady_clock::now();
            vt(robot).validate_edges(E(env).env, a, b, n, out.data(), threads);
            const auto t1 = std::chrono::steady_clock::now();
            best = std::min(best, std::chrono::duration<double>(t1 - t0).count());
        }
        return best;
    }
}
