/* C interface of oracle/_ref/libvamp_ref.so — TEST INFRASTRUCTURE ONLY.
 *
 * The library behind this header is the UNMODIFIED reference validation path, compiled from the
 * sources where they lie under /root/reference/src/impl (see oracle/ref/Makefile), wrapped in a
 * thin harness (ref_env.cc, ref_robot.hh).  It is the parity checker and the CPU baseline; the
 * product (vamp_mvt_b200) never links or loads it.
 *
 * Robot ids: 0 panda, 1 ur5, 2 fetch, 3 baxter.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

    void *ref_env_create(void);
    void ref_env_destroy(void *env);
    /* fields exactly as stored by the reference's shapes (collision/shapes.hh) */
    void ref_env_add_sphere(void *env, const float *xyzr);
    void ref_env_add_cuboid(void *env, const float *f15);
    void ref_env_add_capsule(void *env, const float *f8);
    void ref_env_add_heightfield(void *env, const float *f6, size_t xd, size_t yd, const float *data);
    void ref_env_add_capt(void *env, const float *pts, size_t n, float r_min, float r_max, float r_point);
    int ref_capt_nlog2(void *env, size_t which);
    const float *ref_capt_tests(void *env, size_t which);
    size_t ref_capt_leaf_list(void *env, size_t which, size_t leaf, float *out_xyz, size_t cap);
    void ref_env_add_mvt(
        void *env,
        const float *pts,
        size_t n,
        float r_min,
        float r_max,
        const float *aabb_min,
        const float *aabb_max,
        float r_point);
    void ref_env_attach(void *env, const float *tf12, const float *spheres_xyzr, size_t n);
    void ref_env_detach(void *env);
    /* kind: 0 spheres 1 capsules 2 z_capsules 3 cuboids 4 z_cuboids.  Writes, in the container's
       (sorted) order, the stored fields followed by min_distance; returns the object count. */
    size_t ref_env_dump(void *env, int kind, float *out, size_t cap_floats);

    int ref_robot_dim(int robot);
    int ref_robot_n_spheres(int robot);
    int ref_robot_resolution(int robot);

    /* out[i] = validate_motion<Robot,8,1>(q_i,q_i,env)  (what vamp.<robot>.validate runs) */
    void ref_validate_configs(int robot, void *env, const float *q, size_t n, uint8_t *out, int threads);
    /* out[i] = validate_motion<Robot,8,Robot::resolution>(a_i,b_i,env) */
    void
    ref_validate_edges(int robot, void *env, const float *a, const float *b, size_t n, uint8_t *out, int threads);
    /* out[i][s] = (x,y,z,r) of fine sphere s via Robot::sphere_fk<8> */
    void ref_sphere_fk(int robot, const float *q, size_t n, float *out_xyzr);
    /* Robot::eefk -> 4x4 row-major */
    void ref_eefk(int robot, const float *q, float *out16);
    /* fkcc_debug: env_hits = (sphere, object-id) pairs, self_hits = (sphere, sphere) pairs.
       object ids are the insertion order of the add_* calls. Returns counts through n_env/n_self. */
    void ref_debug(
        int robot,
        void *env,
        const float *q,
        int32_t *env_hits,
        size_t cap_env,
        size_t *n_env,
        int32_t *self_hits,
        size_t cap_self,
        size_t *n_self);
    /* keep[k] = point k survives Helper::filter_self_from_pointcloud (bindings/robot_helper.hh:284-322) */
    /* vamp::planning::simplify<Robot, 8, Robot::resolution> (planning/simplify.hh:191-258) on a path of n
     * waypoints.  ops: SimplifyRoutine values (0 BSPLINE, 1 REDUCE, 2 SHORTCUT, 3 PERTURB); settings12:
     * see ref_robot.hh; samples: [n_samples][dim] unit-cube stream behind RNG::next().  Returns the
     * number of waypoints of the result (written to `out` up to `cap`). */
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
        size_t *iterations);
    /* vamp::collision::filter_pointcloud_centervox (collision/filter_centervox.hh:302-333); returns the number
     * of kept points (written to out_xyz up to cap) or (size_t)-1 where the reference throws */
    size_t ref_filter_centervox(const float *pts, size_t n, float voxel_size, float max_range, const float *origin, const float *ws_min,
                                const float *ws_max, float *out_xyz, size_t cap);
    /* vamp::planning::Path<Robot> (planning/plan.hh:12-153): op 0 cost, 1 subdivide, 2 interpolate_to_resolution(arg),
     * 3 interpolate_to_n_states(arg); the resulting waypoints go to out (up to cap), its cost() to *cost */
    size_t ref_path_op(int robot, int op, const float *path, size_t n, size_t arg, float *out, size_t cap, float *cost);
    /* vamp::rng::Halton<Robot>::next() (random/halton.hh:76-107): samples skip .. skip+n-1, [n][dim] */
    /* the reference's planners over an exact brute-force NN stand-in: which = 0 PRM::build_roadmap, 1 PRM::solve, 2 FCIT::solve */
    size_t ref_planner(int robot, int which, void *env, const float *start, const float *goal, size_t max_iterations, size_t max_samples,
                       size_t batch_size, float *verts, size_t cap_v, uint32_t *edges, size_t cap_e, size_t *n_edges, size_t *iterations, float *cost);
    void ref_halton(int robot, size_t skip, size_t n, float *out);
    void ref_filter_points(int robot, void *env, const float *q, const float *pts, size_t n, float point_radius, uint8_t *keep);
    /* seconds for `reps` passes of ref_validate_configs / edges with `threads` threads (best pass) */
    double ref_time_configs(int robot, void *env, const float *q, size_t n, int threads, int reps);
    double
    ref_time_edges(int robot, void *env, const float *a, const float *b, size_t n, int threads, int reps);

#ifdef __cplusplus
}
#endif
