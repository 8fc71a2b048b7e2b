/* vamp_b200.h -- C ABI of the B200-native batched configuration / motion validation engine.
 *
 * Drop-in boundary for the validation path of chingchennn/vamp_mvt (reference paths below are
 * relative to /root/reference).  The reference has no FFI for this path -- the boundary there is a
 * set of C++ templates plus a nanobind surface -- so each entry point cites the
 * reference interface it replaces.  Plain pointers and sizes only; no exceptions cross the ABI;
 * every call returns VMV_OK (0) or a negative error code, with a message in vmv_last_error().
 *
 * Ownership: the caller owns every host/device buffer it passes; the library owns the
 * environment's device memory.  After vmv_env_commit() an environment is immutable and may be
 * shared by concurrent calls on distinct streams.
 *
 * There is no CPU fallback: every compute entry point fails with VMV_ERR_CUDA when no CUDA device
 * is usable.
 */
#ifndef VAMP_B200_H
#define VAMP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

    enum
    {
        VMV_OK = 0,
        VMV_ERR_ARG = -1,    /* bad argument (null pointer, unknown robot, uncommitted env ...) */
        VMV_ERR_CUDA = -2,   /* CUDA runtime error, see vmv_last_error() */
        VMV_ERR_STATE = -3,  /* call not legal in this state (e.g. add after commit) */
        VMV_ERR_LIMIT = -4   /* environment too large for the shared-memory staging area */
    };

    /* robots: src/impl/vamp/robots/{panda,ur5,fetch,baxter}.hh */
    enum
    {
        VMV_PANDA = 0,
        VMV_UR5 = 1,
        VMV_FETCH = 2,
        VMV_BAXTER = 3,
        VMV_N_ROBOTS = 4
    };

    const char *vmv_last_error(void);
    const char *vmv_version(void);

    /* number of usable CUDA devices (0 => every compute call fails) and device selection */
    int vmv_device_count(void);
    int vmv_set_device(int device);

    /* --- robot constants: Robot::dimension / n_spheres / resolution / s_a / s_m
     *     (robots/panda.hh:14-18,50-66) ------------------------------------------------------- */
    int vmv_robot_id(const char *name); /* "panda" -> VMV_PANDA, unknown -> VMV_ERR_ARG */
    const char *vmv_robot_name(int robot);
    int vmv_robot_dof(int robot);
    int vmv_robot_n_spheres(int robot);
    int vmv_robot_resolution(int robot);
    int vmv_robot_bounds(int robot, float *lower, float *range); /* dof floats each */

    /* --- environment: vamp.Environment and its adders (bindings/environment.cc:111-181) ---------
     * Shapes are passed with exactly the fields the reference stores (collision/shapes.hh):
     *   sphere  (4):  x y z r                                              shapes.hh:229-232
     *   cuboid  (15): x y z | axis_1 xyz | axis_2 xyz | axis_3 xyz | half extents 1 2 3   :34-50
     *   capsule (8):  x1 y1 z1 | xv yv zv | r | rdv = 1/|v|^2                            :135-148
     * min_distance, the z-aligned classification (cuboid: axis_3_z == 1; capsule: xv == 0 &&
     * yv == 0) and the ascending-min_distance sort happen inside (shapes.hh:52-67,165-189,238;
     * environment.cc:120-147; environment.hh:46-72). */
    typedef struct vmv_env vmv_env;
    vmv_env *vmv_env_create(void);
    void vmv_env_destroy(vmv_env *env);
    int vmv_env_add_spheres(vmv_env *env, const float *xyzr, size_t n);
    int vmv_env_add_cuboids(vmv_env *env, const float *f15, size_t n);
    int vmv_env_add_capsules(vmv_env *env, const float *f8, size_t n);
    /* HeightField(x,y,z, xs,ys,zs, xd,yd, data): shapes.hh:262-283; f6 = x y z xs ys zs where
     * xs,ys,zs are the INVERSE scales as stored (factory.hh:365-386) */
    int vmv_env_add_heightfield(vmv_env *env, const float *f6, size_t xd, size_t yd, const float *data);
    /* CAPT(points, r_min, r_max, r_point): capt.hh:299-369 (add_capt_pointcloud, environment.cc:150-160).
     * Builds the reference's k-d tree (median splits, capt.hh:106-119) and NOT its affordance lists: the queries
     * evaluate list membership per point from the tree (same verdicts, including the points the reference's lists
     * miss), and the tree is built on the GPU (segmented bitonic sorts): ~3.5 ms per 10^5 points instead of seconds.
     * At most 2^24 points (VMV_ERR_LIMIT). */
    int vmv_env_add_capt(vmv_env *env, const float *points_xyz, size_t n, float r_min, float r_max, float r_point);
    /* MVT(points, r_min, r_max, workspace_aabb_min, workspace_aabb_max, r_point): mvt.hh:146-170
     * (add_mvt_pointcloud, environment.cc:163-176).  The grid is cubic, floor(workspace x-width /
     * r_max) cells per axis (mvt.hh:437-446); points outside the workspace are clamped into it. */
    int vmv_env_add_mvt(vmv_env *env, const float *points_xyz, size_t n, float r_min, float r_max, const float *aabb_min_xyz, const float *aabb_max_xyz, float r_point);
    /* Attachment(tf) + add_spheres (attachments.hh:12-56, environment.cc:177-181).  tf12 =
     * translation(3) followed by the 3x3 rotation in column-major order (vector/math.hh:39-51). */
    int vmv_env_attach(vmv_env *env, const float *tf12, const float *spheres_xyzr, size_t n);
    int vmv_env_detach(vmv_env *env);
    /* sort + classify + pack + upload.  Adders may be called again afterwards; the next commit
     * re-packs.  Compute calls require a committed environment.  A commit after attach / detach alone
     * rewrites a few hundred bytes and keeps every device table (attach / detach are O(1) in the
     * reference: std::optional<Attachment>, collision/environment.hh:28). */
    int vmv_env_commit(vmv_env *env);
    /* introspection used by the tests: kind 0 spheres, 1 capsules, 2 z-capsules, 3 cuboids,
     * 4 z-cuboids; writes the stored fields followed by min_distance, in sorted order; returns count */
    long vmv_env_dump(const vmv_env *env, int kind, float *out, size_t cap_floats);

    /* --- batched validation -------------------------------------------------------------------
     * valid_bits: bit (i & 31) of word (i >> 5) is 1 iff unit i is valid; ceil(n/32) words.
     *
     * vmv_validate_configs*: unit i = vamp.<robot>.validate(q_i, env)
     *   = validate_motion<Robot,8,1>(q_i, q_i, env)   (bindings/robot_helper.hh:255-267,
     *   planning/validate.hh:70-77) = Robot::fkcc / fkcc_attach on the state (robots/panda.hh:5227).
     * vmv_validate_edges*: unit i = validate_motion<Robot,8,resolution>(a_i, b_i, env)
     *   (planning/validate.hh:24-77); resolution <= 0 selects Robot::resolution.
     * q, a, b: row-major [n][dof] float32.
     *
     * *_dev: all pointers are device pointers on the current device; asynchronous on `stream`
     * (a cudaStream_t, may be NULL).  The plain versions take host pointers and include the
     * host<->device copies and a synchronisation. */
    int vmv_validate_configs_dev(int robot, const vmv_env *env, const float *d_q, size_t n, uint32_t *d_valid_bits, void *stream);
    int vmv_validate_edges_dev(int robot, const vmv_env *env, const float *d_a, const float *d_b, size_t n, int resolution, uint32_t *d_valid_bits, void *stream);
    int vmv_validate_configs(int robot, const vmv_env *env, const float *q, size_t n, uint32_t *valid_bits);
    int vmv_validate_edges(int robot, const vmv_env *env, const float *a, const float *b, size_t n, int resolution, uint32_t *valid_bits);

    /* CenterVox pointcloud down-sampling: vamp.filter_pointcloud(..., filter_type="centervox")
     * (bindings/environment.cc:212-239) = vamp::collision::filter_pointcloud_centervox
     * (src/impl/vamp/collision/filter_centervox.hh:302-333): points inside the range sphere around `origin`
     * and inside the workspace box go to a grid of at most 255^3 voxels; a voxel keeps the point closest to
     * its centre (the earlier one on a tie); the result lists the occupied voxels in the order the
     * reference's sparse tables created them.  points: [n][3] float32 (host; _dev: device memory on the
     * current device).  kept_indices (host, up to `cap`; 32768 always suffices) receives the indices of the
     * kept points in the reference's output order, *n_kept their number.  VMV_ERR_LIMIT where the reference
     * throws "Voxel pool exhausted" (more occupied voxels than min((width/voxel_size)^3 * 0.05, 32768)). */
    int vmv_filter_pointcloud_centervox(const float *points, size_t n, float voxel_size, float max_range, const float *origin3,
                                        const float *workspace_min3, const float *workspace_max3, uint32_t *kept_indices, size_t cap, size_t *n_kept);
    int vmv_filter_pointcloud_centervox_dev(const float *d_points, size_t n, float voxel_size, float max_range, const float *origin3,
                                            const float *workspace_min3, const float *workspace_max3, uint32_t *kept_indices, size_t cap, size_t *n_kept);

    /* The reference's Halton configuration sampler (vamp::rng::Halton<Robot>::next(),
     * src/impl/vamp/random/halton.hh:76-107) evaluated on the device: sample s (0-based) of a fresh
     * sequence, scaled to the joint ranges exactly as next() returns it.  The planners draw every sample
     * from it (prm.hh:109-113, rrtc.hh:79, fcit.hh:322-330), so a batched front-end can have samples
     * first .. first+n-1 generated AND validated on the GPU and receive one bit per sample instead of
     * uploading 4*dof bytes per sample:
     *   vmv_halton_fill_dev   writes the configurations to device memory [n][dof];
     *   vmv_validate_halton   = fill + vmv_validate_configs_dev + D2H of the bits (q_out, if not NULL,
     *                           also receives the configurations, host memory [n][dof]).
     * Exact only while the reference's f32 recurrence is (first + n <= vmv_halton_exact_limit(robot):
     * 10^6 for up to 8 joints, 707280 for Baxter); beyond that VMV_ERR_LIMIT. */
    int vmv_halton_fill_dev(int robot, uint64_t first, size_t n, float *d_q, void *stream);
    int vmv_validate_halton(int robot, const vmv_env *env, uint64_t first, size_t n, uint32_t *valid_bits, float *q_out);
    uint64_t vmv_halton_exact_limit(int robot);

    /* PRM-style edge sets (prm.hh:136-146): edge i joins vertices pairs[2i], pairs[2i+1] of a
     * vertex table [n_vertices][dof]; 8 bytes per edge instead of 8*dof. */
    int vmv_validate_edges_indexed_dev(int robot, const vmv_env *env, const float *d_vertices, size_t n_vertices, const uint32_t *d_pairs, size_t n_edges, int resolution, uint32_t *d_valid_bits, void *stream);
    int vmv_validate_edges_indexed(int robot, const vmv_env *env, const float *vertices, size_t n_vertices, const uint32_t *pairs, size_t n_edges, int resolution, uint32_t *valid_bits);

    /* Robot::sphere_fk (robots/panda.hh:117-462; vamp.<robot>.fk, robot_helper.hh:234-247):
     * out[i][s] = (x, y, z, r) of fine sphere s for configuration i. */
    int vmv_sphere_fk_dev(int robot, const float *d_q, size_t n, float *d_xyzr, void *stream);
    int vmv_sphere_fk(int robot, const float *q, size_t n, float *xyzr);

    /* Helper::filter_self_from_pointcloud (bindings/robot_helper.hh:284-322;
     * vamp.<robot>.filter_self_from_pointcloud): bit k of keep_bits is 1 iff point k, as a sphere of
     * radius point_radius, neither overlaps a robot sphere at configuration q nor collides with the
     * environment. */
    int vmv_filter_points_dev(int robot, const vmv_env *env, const float *d_q, const float *d_points_xyz, size_t n, float point_radius, uint32_t *d_keep_bits, void *stream);
    int vmv_filter_self_from_pointcloud(int robot, const vmv_env *env, const float *q, const float *points_xyz, size_t n, float point_radius, uint32_t *keep_bits);

    /* Robot::fkcc_debug (robots/panda.hh:468-5224; vamp.<robot>.debug): for ONE configuration,
     * env_hits receives (fine sphere, object id) pairs -- object ids count the shapes in the
     * order they were added -- and self_hits (sphere, sphere) pairs.  Counts are returned through
     * n_env / n_self even when they exceed the capacities. */
    int vmv_debug(int robot, const vmv_env *env, const float *q, int32_t *env_hits, size_t cap_env, size_t *n_env, int32_t *self_hits, size_t cap_self, size_t *n_self);

    /* --- batched roadmap planner front-end: PRM<Robot, rake, resolution> (planning/prm.hh:43-300) -----------------
     * The reference's loop -- sample, fkcc, k nearest roadmap vertices within the PRM* radius (roadmap.hh:42-67),
     * validate_motion per neighbour, union-find, A* -- restated as bulk steps (csrc/vmv_planner.cu): Halton samples
     * generated and validated on the device, every vertex's exact neighbours among the earlier vertices by one
     * kernel, all candidate edges as one indexed edge batch, and the sequential bookkeeping replayed on the host
     * from the verdict bits.  The result is the reference's: same vertices, adjacency, iteration count and (solve) path.
     *   solve = 0: PRM::build_roadmap(start, goal, env, settings, rng)        (prm.hh:198-300)
     *   solve = 1: PRM::solve(start, goal, env, settings, rng)                (prm.hh:43-196): stops at the iteration that
     *              connects start and goal; vmv_roadmap_path then gives the waypoints start .. goal (utils::recover_path,
     *              planning/utils.hh:144-162), vmv_roadmap_iterations the reference's `iterations`.
     * Sampler: the reference's default, Halton (random/halton.hh); settings: max_iterations, max_samples and
     * PRMStarNeighborParams(dof, space_measure) -- space_measure = Robot::space_measure() (robots/panda.hh:111-114). */
    typedef struct vmv_roadmap vmv_roadmap;
    int vmv_prm(int robot, const vmv_env *env, const float *start, const float *goal, size_t max_iterations, size_t max_samples, double space_measure, int solve, vmv_roadmap **out);
    /* FCIT<Robot, rake, resolution>::solve (planning/fcit.hh:82-360): a lazy best-first search over the complete graph
     * of the samples, fed by batches.  The search itself is sequential and runs on the host (csrc/vmv_fcit.hpp, the
     * reference's control flow restated); what it consumes comes from the GPU in bulk: each batch of batch_size valid
     * samples from chunks of the Halton stream generated and validated on the device (fcit.hh:322-348), and each edge
     * verdict (fcit.hh:230-260) from a row cache -- the first question about a parent validates its edges to every
     * node in one indexed batch.  Path, cost and iteration count are the reference's.  vmv_roadmap_vertices then
     * lists start, goal and the samples; vmv_roadmap_path the solution (just the start when there is none). */
    int vmv_fcit(int robot, const vmv_env *env, const float *start, const float *goal, size_t max_iterations, size_t max_samples, size_t batch_size, int optimize, vmv_roadmap **out);
    void vmv_roadmap_destroy(vmv_roadmap *roadmap);
    size_t vmv_roadmap_vertices(const vmv_roadmap *roadmap, const float **q);                         /* count; [n][dof] */
    size_t vmv_roadmap_edges(const vmv_roadmap *roadmap, const uint32_t **pairs, const float **cost); /* adjacency entries (from, to), Roadmap::edges order */
    size_t vmv_roadmap_path(const vmv_roadmap *roadmap, const float **q, float *cost);                /* waypoints (0: no solution) */
    size_t vmv_roadmap_iterations(const vmv_roadmap *roadmap);
    size_t vmv_roadmap_work(const vmv_roadmap *roadmap, size_t *edges_checked);                       /* samples drawn */

    /* --- multi-GPU: one process per GPU of one node --------------------------------------------------
     * The reference is single-process and has no counterpart: a batch shards by unit, every rank
     * validates the contiguous word-aligned range it owns against its replica of the environment, and the verdict
     * words are gathered only where the planner needs the global mask (prm.hh:136-146 consumes one bit per
     * candidate edge).  NCCL (loaded at run time; never needed on one GPU) carries the plain collectives; the
     * gather itself can be fused into the validation kernels over CUDA-IPC windows (below).
     *
     * vmv_comm_unique_id   rank 0: 128 bytes to hand to every rank by any out-of-band means (ncclGetUniqueId)
     * vmv_comm_create      collective: rank `rank` of `world` (<= 8) on the CURRENT device
     * vmv_env_broadcast    collective: `root`'s environment (everything added so far, built CAPT / MVT tables
     *                      included -- no rank rebuilds them) replaces `env` on the other ranks; committed on return
     * vmv_allgather_bits   plain ncclAllGather of words_per_rank verdict words per rank into d_global
     *                      [world][words_per_rank], asynchronous on `stream` */
    typedef struct vmv_comm vmv_comm;
    int vmv_comm_unique_id(unsigned char *id128);
    int vmv_comm_create(vmv_comm **comm, const unsigned char *id128, int rank, int world);
    void vmv_comm_destroy(vmv_comm *comm);
    int vmv_comm_rank(const vmv_comm *comm);
    int vmv_comm_world(const vmv_comm *comm);
    int vmv_env_broadcast(vmv_comm *comm, vmv_env *env, int root);
    int vmv_allgather_bits(vmv_comm *comm, const uint32_t *d_local_words, size_t words_per_rank, uint32_t *d_global_words, void *stream);

    /* Gather fused into the validation kernels.  vmv_comm_window (collective, once) gives every rank a window
     * of `slots` (<= 4) global masks [world][stride] words (stride = words_per_rank rounded up to 4, see
     * vmv_comm_window_stride) and maps all windows into every process (CUDA IPC over NVLink / NVSwitch).
     * vmv_validate_*_gather_dev (collective per slot: every rank calls it for the same slot the same number of
     * times) validates this rank's shard and stores each verdict word, as it is produced, into row `rank` of
     * slot `slot` of EVERY rank's window -- one predicated store instruction per 32 units, no collective
     * kernel, no SM taken from the validation blocks; the last block publishes a sequence number behind a
     * system-scope fence.  vmv_comm_wait enqueues, on `stream`, a wait until the latest gather into `slot` has
     * landed from every rank; work enqueued behind it on that stream may read vmv_comm_window_ptr(comm, slot).
     * A slot may be written again once every rank is done reading its previous contents (the caller's
     * protocol; with two slots used alternately and a wait per step that holds by construction). */
    int vmv_comm_window(vmv_comm *comm, size_t words_per_rank, int slots);
    uint32_t *vmv_comm_window_ptr(vmv_comm *comm, int slot);
    size_t vmv_comm_window_stride(const vmv_comm *comm);
    int vmv_validate_configs_gather_dev(int robot, const vmv_env *env, vmv_comm *comm, int slot, const float *d_q, size_t n, void *stream);
    int vmv_validate_edges_indexed_gather_dev(int robot, const vmv_env *env, vmv_comm *comm, int slot, const float *d_vertices, size_t n_vertices, const uint32_t *d_pairs, size_t n_edges, int resolution, void *stream);
    int vmv_comm_wait(vmv_comm *comm, int slot, void *stream);
    /* The same gather without touching the kernels: a launch on `stream` writes this rank's verdict words into
     * its OWN window row (vmv_comm_local_row), then vmv_comm_publish (collective per slot, like the fused calls)
     * hands the row and its flag to the copy engines on a side stream that waits for `stream` -- no SM is used
     * and the copies overlap the next launch.  At most 4095 publications per slot. */
    uint32_t *vmv_comm_local_row(vmv_comm *comm, int slot);
    int vmv_comm_publish(vmv_comm *comm, int slot, size_t n_words, void *stream);
    /* before a launch on `stream` overwrites the local row of `slot`: a device-side wait until the copy engines
     * have sent its previous contents (no-op for a slot never published) */
    int vmv_comm_acquire(vmv_comm *comm, int slot, void *stream);

    /* --- device buffers for callers that do not bring their own allocator ---------------------- */
    void *vmv_dev_alloc(size_t bytes);
    void vmv_dev_free(void *p);
    void *vmv_host_alloc_pinned(size_t bytes);
    void vmv_host_free_pinned(void *p);
    int vmv_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
    int vmv_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
    int vmv_stream_sync(void *stream);

    /* --- measurement helpers -------------------------------------------------------------------
     * number of kernels this library has launched since load (bench.py reports it as gpu_launches) */
    uint64_t vmv_launch_count(void);
    /* development aid: 64 counters filled by builds compiled with -DVMV_C4_STATS (zeros otherwise) */
    int vmv_dev_stats(uint64_t *out64, int reset);
    /* development / tests: FNV-1a digests of the k-th CAPT's tables {nodes, leaf flags, grid points, grid starts}: the device
     * build (default) and the host build (environment variable VMV_CAPT_HOST_BUILD) must agree bit for bit */
    int vmv_env_capt_digest(const vmv_env *env, int k, uint64_t *out4);
    /* development / tests: the k-th CAPT's nodes in Eytzinger order, {split value, inherited bit} per node (the split values are
     * the reference's CAPT::tests); returns the number of floats (2 per node), copies at most cap_floats */
    long vmv_env_capt_nodes(const vmv_env *env, int k, float *out, size_t cap_floats);
    /* testing aid: 0 = automatic choice (default), 1 = force the generic per-thread kernel,
     * 2 = force the block-cooperative kernel, 3 = force the grid-culled kernel (2 and 3 fail with
     * VMV_ERR_LIMIT when they do not apply to the environment) */
    int vmv_force_kernel_path(int path);

#ifdef __cplusplus
}
#endif
#endif /* VAMP_B200_H */
