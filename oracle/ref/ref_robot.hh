// Per-robot harness over the UNMODIFIED reference headers — TEST INFRASTRUCTURE ONLY.
// Each function calls the reference's own entry points:
//   validate_motion<Robot,8,res>   src/impl/vamp/planning/validate.hh:70
//   Robot::sphere_fk<8>            src/impl/vamp/robots/<robot>.hh
//   Robot::fkcc_debug<8>           src/impl/vamp/robots/<robot>.hh
//   Robot::eefk                    src/impl/vamp/robots/<robot>.hh
//   simplify<Robot,8,res>          src/impl/vamp/planning/simplify.hh:191 (shortcut, B-spline, reduce, perturb)
// exactly as the nanobind layer does (src/impl/vamp/bindings/robot_helper.hh:234-267).
#pragma once
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include <vamp/collision/environment.hh>
#include <vamp/collision/sphere_sphere.hh>
#include <vamp/collision/validity.hh>
#include <vamp/planning/validate.hh>
#include <vamp/planning/simplify.hh>
#include <vamp/random/halton.hh>
#include <vamp/planning/prm.hh>
#include <vamp/planning/fcit.hh>

// the product's host-side FCIT* restatement, driven here by the reference's own checks (CPU pin of its control flow)
#include "../../vamp_mvt_b200/csrc/vmv_fcit.hpp"
#include <vamp/vector.hh>

namespace refh
{
    using EnvF = vamp::collision::Environment<float>;
    using EnvV = vamp::collision::Environment<vamp::FloatVector<vamp::FloatVectorWidth>>;
    static constexpr std::size_t rake = vamp::FloatVectorWidth;

    struct RobotVTable
    {
        int dim;
        int n_spheres;
        int resolution;
        void (*validate_configs)(const EnvF &, const float *, std::size_t, std::uint8_t *, int);
        void (*validate_edges)(const EnvF &, const float *, const float *, std::size_t, std::uint8_t *, int);
        void (*sphere_fk)(const float *, std::size_t, float *);
        void (*eefk)(const float *, float *);
        void (*debug)(
            const EnvF &,
            const float *,
            std::vector<std::pair<int, int>> &,
            std::vector<std::pair<int, int>> &);
        void (*filter_points)(const EnvF &, const float *, const float *, std::size_t, float, std::uint8_t *);
        std::size_t (*simplify)(
            const EnvF &,
            const float *,
            std::size_t,
            const int *,
            std::size_t,
            const float *,
            const float *,
            std::size_t,
            float *,
            std::size_t,
            std::size_t *);
        void (*halton)(std::size_t, std::size_t, float *);
        std::size_t (*path_op)(int, const float *, std::size_t, std::size_t, float *, std::size_t, float *);
        // planner: 0 PRM::build_roadmap, 1 PRM::solve, 2 FCIT::solve (planning/prm.hh, fcit.hh) with the Halton sampler.
        // out: vertices [n][dim] (roadmap) or the path's waypoints (solve); edges as (from, to) pairs per adjacency entry
        std::size_t (*planner)(int, const EnvF &, const float *, const float *, std::size_t, std::size_t, std::size_t, float *, std::size_t,
                               std::uint32_t *, std::size_t, std::size_t *, std::size_t *, float *);
    };

    template <typename Fn>
    inline void parallel_for(std::size_t n, int threads, Fn &&fn)
    {
        if (threads <= 1 or n < 2)
        {
            fn(std::size_t(0), n);
            return;
        }

        std::vector<std::thread> pool;
        const std::size_t chunk = (n + threads - 1) / threads;
        for (int t = 0; t < threads; ++t)
        {
            const std::size_t lo = std::min(n, chunk * t);
            const std::size_t hi = std::min(n, lo + chunk);
            if (lo < hi)
            {
                pool.emplace_back([=, &fn] { fn(lo, hi); });
            }
        }

        for (auto &th : pool)
        {
            th.join();
        }
    }

    template <typename Robot>
    struct Harness
    {
        using Configuration = typename Robot::Configuration;
        using ConfigurationArray = typename Robot::ConfigurationArray;

        static auto load(const float *q) -> Configuration
        {
            ConfigurationArray a;
            for (std::size_t j = 0; j < Robot::dimension; ++j)
            {
                a[j] = q[j];
            }
            return Configuration(a);
        }

        static void
        validate_configs(const EnvF &env, const float *q, std::size_t n, std::uint8_t *out, int threads)
        {
            parallel_for(
                n,
                threads,
                [&](std::size_t lo, std::size_t hi)
                {
                    // one private vectorised environment per thread, as the bindings build per call
                    // (bindings/robot_helper.hh:266) but hoisted out of the loop
                    const EnvV ev(env);
                    for (std::size_t i = lo; i < hi; ++i)
                    {
                        const auto c = load(q + i * Robot::dimension);
                        out[i] = vamp::planning::validate_motion<Robot, rake, 1>(c, c, ev) ? 1 : 0;
                    }
                });
        }

        static void validate_edges(
            const EnvF &env,
            const float *a,
            const float *b,
            std::size_t n,
            std::uint8_t *out,
            int threads)
        {
            parallel_for(
                n,
                threads,
                [&](std::size_t lo, std::size_t hi)
                {
                    const EnvV ev(env);
                    for (std::size_t i = lo; i < hi; ++i)
                    {
                        const auto ca = load(a + i * Robot::dimension);
                        const auto cb = load(b + i * Robot::dimension);
                        out[i] =
                            vamp::planning::validate_motion<Robot, rake, Robot::resolution>(ca, cb, ev) ? 1 :
                                                                                                          0;
                    }
                });
        }

        static void sphere_fk(const float *q, std::size_t n, float *out)
        {
            for (std::size_t i = 0; i < n; ++i)
            {
                typename Robot::template ConfigurationBlock<rake> block;
                for (std::size_t j = 0; j < Robot::dimension; ++j)
                {
                    block[j] = q[i * Robot::dimension + j];
                }

                typename Robot::template Spheres<rake> s;
                Robot::template sphere_fk<rake>(block, s);
                for (std::size_t k = 0; k < Robot::n_spheres; ++k)
                {
                    float *o = out + (i * Robot::n_spheres + k) * 4;
                    o[0] = s.x[{k, 0}];
                    o[1] = s.y[{k, 0}];
                    o[2] = s.z[{k, 0}];
                    o[3] = s.r[{k, 0}];
                }
            }
        }

        static void eefk(const float *q, float *out16)
        {
            std::array<float, Robot::dimension> a;
            for (std::size_t j = 0; j < Robot::dimension; ++j)
            {
                a[j] = q[j];
            }

            const auto m = Robot::eefk(a).matrix();
            for (int r = 0; r < 4; ++r)
            {
                for (int c = 0; c < 4; ++c)
                {
                    out16[r * 4 + c] = m(r, c);
                }
            }
        }

        static void debug(
            const EnvF &env,
            const float *q,
            std::vector<std::pair<int, int>> &env_hits,
            std::vector<std::pair<int, int>> &self_hits)
        {
            typename Robot::template ConfigurationBlock<rake> block;
            for (std::size_t j = 0; j < Robot::dimension; ++j)
            {
                block[j] = q[j];
            }

            const auto dbg = Robot::template fkcc_debug<rake>(EnvV(env), block);
            for (std::size_t s = 0; s < dbg.first.size(); ++s)
            {
                for (const auto &name : dbg.first[s])
                {
                    env_hits.emplace_back(static_cast<int>(s), std::atoi(name.c_str()));
                }
            }

            for (const auto &p : dbg.second)
            {
                self_hits.emplace_back(static_cast<int>(p.first), static_cast<int>(p.second));
            }
        }

        // Helper::filter_self_from_pointcloud (bindings/robot_helper.hh:284-322): the binding itself
        // needs nanobind, so its loop is composed here from the same reference functions it calls
        // (Robot::sphere_fk<1>, sphere_sphere_sql2 with `< 0`, sphere_environment_in_collision).
        static void filter_points(
            const EnvF &env,
            const float *q,
            const float *pts,
            std::size_t n,
            float point_radius,
            std::uint8_t *keep)
        {
            const EnvV ev(env);
            typename Robot::template ConfigurationBlock<1> block;
            for (std::size_t j = 0; j < Robot::dimension; ++j)
            {
                block[j] = q[j];
            }

            typename Robot::template Spheres<1> out;
            Robot::template sphere_fk<1>(block, out);
            for (std::size_t k = 0; k < n; ++k)
            {
                const float x = pts[3 * k], y = pts[3 * k + 1], z = pts[3 * k + 2], r = point_radius;
                bool valid = true;
                for (auto i = 0U; i < Robot::n_spheres; ++i)
                {
                    if (vamp::collision::sphere_sphere_sql2(
                            out.x[{i, 0}], out.y[{i, 0}], out.z[{i, 0}], out.r[{i, 0}], x, y, z, r) < 0 or
                        vamp::sphere_environment_in_collision<>(ev, x, y, z, r))
                    {
                        valid = false;
                        break;
                    }
                }

                keep[k] = valid ? 1 : 0;
            }
        }

        // The reference's RNG interface (random/rng.hh:9-17) fed from a caller-supplied stream of unit-cube
        // samples, so that the product side can be driven with the same stream.  `dist` (std::mt19937
        // seeded 0, random/distribution.hh:11) is the reference's own.
        struct StreamRNG : public vamp::rng::RNG<Robot>
        {
            const float *samples{nullptr};
            std::size_t n{0}, at{0};

            inline void reset() noexcept override
            {
                at = 0;
                this->dist.reset();
            }

            inline auto next() noexcept -> vamp::FloatVector<Robot::dimension> override
            {
                const float *s = samples + (n ? (at++ % n) : 0) * Robot::dimension;
                return load(s);
            }
        };

        // simplify<Robot, rake, Robot::resolution> as bindings/robot_helper.hh:269-277 calls it.
        // ops: SimplifyRoutine values; sf = {max_iterations, interpolate, bspline.max_steps,
        // bspline.min_change, bspline.midpoint_interpolation, reduce.max_steps, reduce.max_empty_steps,
        // reduce.range_ratio, perturb.max_steps, perturb.max_empty_steps, perturb.perturbation_attempts,
        // perturb.range}.  Returns the number of waypoints (written up to `cap`).
        static std::size_t simplify(
            const EnvF &env,
            const float *path,
            std::size_t n,
            const int *ops,
            std::size_t n_ops,
            const float *sf,
            const float *samples,
            std::size_t n_samples,
            float *out,
            std::size_t cap,
            std::size_t *iterations)
        {
            namespace vp = vamp::planning;
            vp::Path<Robot> p;
            for (std::size_t i = 0; i < n; ++i)
            {
                p.emplace_back(load(path + i * Robot::dimension));
            }

            vp::SimplifySettings st;
            st.operations.clear();
            for (std::size_t i = 0; i < n_ops; ++i)
            {
                st.operations.emplace_back(static_cast<vp::SimplifyRoutine>(ops[i]));
            }

            st.max_iterations = static_cast<std::size_t>(sf[0]);
            st.interpolate = static_cast<std::size_t>(sf[1]);
            st.bspline.max_steps = static_cast<std::size_t>(sf[2]);
            st.bspline.min_change = sf[3];
            st.bspline.midpoint_interpolation = sf[4];
            st.reduce.max_steps = static_cast<std::size_t>(sf[5]);
            st.reduce.max_empty_steps = static_cast<std::size_t>(sf[6]);
            st.reduce.range_ratio = sf[7];
            st.perturb.max_steps = static_cast<std::size_t>(sf[8]);
            st.perturb.max_empty_steps = static_cast<std::size_t>(sf[9]);
            st.perturb.perturbation_attempts = static_cast<std::size_t>(sf[10]);
            st.perturb.range = sf[11];

            auto rng = std::make_shared<StreamRNG>();
            rng->samples = samples;
            rng->n = n_samples;

            const EnvV ev(env);
            const auto result = vp::simplify<Robot, rake, Robot::resolution>(p, ev, st, rng);
            for (std::size_t i = 0; i < result.path.size() and i < cap; ++i)
            {
                const auto a = result.path[i].to_array();
                for (std::size_t j = 0; j < Robot::dimension; ++j)
                {
                    out[i * Robot::dimension + j] = a[j];
                }
            }

            *iterations = result.iterations;
            return result.path.size();
        }

        // vamp::planning::Path<Robot> members (planning/plan.hh:12-153): op 0 cost, 1 subdivide,
        // 2 interpolate_to_resolution(arg), 3 interpolate_to_n_states(arg).  Returns the new size.
        static std::size_t
        path_op(int op, const float *path, std::size_t n, std::size_t arg, float *out, std::size_t cap, float *cost)
        {
            vamp::planning::Path<Robot> p;
            for (std::size_t i = 0; i < n; ++i)
            {
                p.emplace_back(load(path + i * Robot::dimension));
            }

            if (op == 1)
            {
                p.subdivide();
            }
            else if (op == 2)
            {
                p.interpolate_to_resolution(arg);
            }
            else if (op == 3)
            {
                p.interpolate_to_n_states(arg);
            }

            *cost = p.cost();
            for (std::size_t i = 0; i < p.size() and i < cap; ++i)
            {
                const auto a = p[i].to_array();
                for (std::size_t j = 0; j < Robot::dimension; ++j)
                {
                    out[i * Robot::dimension + j] = a[j];
                }
            }

            return p.size();
        }

        // vamp::rng::Halton<Robot> (random/halton.hh): samples skip .. skip + n - 1 of a fresh sequence,
        // scaled to the joint ranges as next() returns them.
        static void halton(std::size_t skip, std::size_t n, float *out)
        {
            vamp::rng::Halton<Robot> h;
            for (std::size_t i = 0; i < skip + n; ++i)
            {
                const auto a = h.next().to_array();
                if (i >= skip)
                {
                    for (std::size_t j = 0; j < Robot::dimension; ++j)
                    {
                        out[(i - skip) * Robot::dimension + j] = a[j];
                    }
                }
            }
        }

        // The reference's own roadmap planners (planning/prm.hh:43-300, fcit.hh:82-360), unmodified, over the exact
        // brute-force NN stand-in (shim/vamp/planning/nn.hh) and the reference's Halton sampler.
        static std::size_t planner(
            int which,
            const EnvF &env_f,
            const float *start_f,
            const float *goal_f,
            std::size_t max_iterations,
            std::size_t max_samples,
            std::size_t batch_size,
            float *verts,
            std::size_t cap_v,
            std::uint32_t *edges,
            std::size_t cap_e,
            std::size_t *n_edges,
            std::size_t *iterations,
            float *cost)
        {
            namespace vp = vamp::planning;
            const EnvV env(env_f);
            const Configuration start(load(start_f)), goal(load(goal_f));
            typename vamp::rng::RNG<Robot>::Ptr rng = std::make_shared<vamp::rng::Halton<Robot>>();
            std::size_t nv = 0;
            *n_edges = 0;
            auto put = [&](const Configuration &c)
            {
                if (nv < cap_v)
                {
                    const auto a = c.to_array();
                    for (std::size_t j = 0; j < Robot::dimension; ++j)
                    {
                        verts[nv * Robot::dimension + j] = a[j];
                    }
                }
                ++nv;
            };
            if (which == 0)
            {
                vp::RoadmapSettings<vp::PRMStarNeighborParams> settings(vp::PRMStarNeighborParams(Robot::dimension, Robot::space_measure()));
                settings.max_iterations = max_iterations, settings.max_samples = max_samples;
                const auto rm = vp::PRM<Robot, rake, Robot::resolution>::build_roadmap(start, goal, env, settings, rng);
                for (const auto &v : rm.vertices)
                {
                    put(v);
                }
                for (std::size_t i = 0; i < rm.edges.size(); ++i)
                {
                    for (const auto j : rm.edges[i])
                    {
                        if (*n_edges < cap_e)
                        {
                            edges[2 * *n_edges] = static_cast<std::uint32_t>(i);
                            edges[2 * *n_edges + 1] = static_cast<std::uint32_t>(j);
                        }
                        ++*n_edges;
                    }
                }
                *iterations = rm.iterations;
                *cost = 0.F;
            }
            else if (which == 1)
            {
                vp::RoadmapSettings<vp::PRMStarNeighborParams> settings(vp::PRMStarNeighborParams(Robot::dimension, Robot::space_measure()));
                settings.max_iterations = max_iterations, settings.max_samples = max_samples;
                const auto res = vp::PRM<Robot, rake, Robot::resolution>::solve(start, goal, env, settings, rng);
                for (const auto &v : res.path)
                {
                    put(v);
                }
                *iterations = res.iterations;
                *cost = res.cost;
            }
            else if (which == 3)
            {
                // vmvfcit::solve (the product's restatement) with the reference's fkcc / validate_motion / Halton as providers
                std::vector<float> states(start_f, start_f + Robot::dimension);
                states.insert(states.end(), goal_f, goal_f + Robot::dimension);
                auto dist = [](const float *a, const float *b) { return load(a).distance(load(b)); };
                auto samples = [&](std::size_t n, std::vector<float> &st)
                {
                    std::size_t got = 0;
                    typename Robot::template ConfigurationBlock<rake> block;
                    while (got < n)
                    {
                        auto c = rng->next();
                        for (auto i = 0U; i < Robot::dimension; ++i)
                        {
                            block[i] = c.broadcast(i);
                        }
                        if (not Robot::template fkcc<rake>(env, block))
                        {
                            continue;
                        }
                        const auto a = c.to_array();
                        st.insert(st.end(), a.begin(), a.begin() + Robot::dimension);
                        ++got;
                    }
                    return got;
                };
                auto edges = [&](std::uint32_t p, std::uint32_t c)
                {
                    return vp::validate_motion<Robot, rake, Robot::resolution>(
                        load(states.data() + p * Robot::dimension), load(states.data() + c * Robot::dimension), env);
                };
                const auto res = vmvfcit::solve(states, Robot::dimension, max_iterations, max_samples, batch_size, false, dist, samples, edges);
                for (const auto i : res.path)
                {
                    put(load(states.data() + i * Robot::dimension));
                }
                *iterations = res.iterations;
                *cost = res.cost;
            }
            else
            {
                vp::RoadmapSettings<vp::FCITStarNeighborParams> settings(vp::FCITStarNeighborParams(Robot::dimension, Robot::space_measure()));
                settings.max_iterations = max_iterations, settings.max_samples = max_samples, settings.batch_size = batch_size;
                const auto res = vp::FCIT<Robot, rake, Robot::resolution>::solve(start, goal, env, settings, rng);
                for (const auto &v : res.path)
                {
                    put(v);
                }
                *iterations = res.iterations;
                *cost = res.cost;
            }
            return nv;
        }

        static constexpr RobotVTable vtable{
            static_cast<int>(Robot::dimension),
            static_cast<int>(Robot::n_spheres),
            static_cast<int>(Robot::resolution),
            &validate_configs,
            &validate_edges,
            &sphere_fk,
            &eefk,
            &debug,
            &filter_points,
            &simplify,
            &halton,
            &path_op,
            &planner};
    };
}  // namespace refh
