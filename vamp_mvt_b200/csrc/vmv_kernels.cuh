// Validation kernels (sm_100a).  One thread = one robot state.
//
//   k_validate_configs : one configuration per thread                 (reference Robot::fkcc,
//                        robots/panda.hh:5227, via validate_motion<Robot,8,1>)
//   k_validate_edges   : one warp per edge, one lane per raked state  (reference validate_vector,
//                        planning/validate.hh:24-67)
//   k_sphere_fk        : fine sphere centres                          (reference Robot::sphere_fk)
//
// Structure of a state check (phases):
//   A  straight-line FK of the moving rigid bodies (generated, gen/<robot>_fk.cuh); the 3x4 frames go
//      to a per-thread stash in shared memory, laid out [entry][thread] (bank-conflict free);
//   B  environment sweep as a task loop: every thread walks its own list of sphere tasks (bounding
//      sphere of a link -> on hit its fine spheres, else skip to the next link) but all threads run
//      the same loop body, so a warp stays converged no matter which link each lane is on;
//   C  self collision: bounding pair test on the bounding centres saved during B, fine pairs on hit;
//   D  attachment spheres (posed by the end-effector frame) against the environment and the links.
//
// The environment blob, the robot's task/link/pair tables and the stash live in dynamic shared
// memory; the blob arrives by one TMA bulk copy that overlaps phase A.
#pragma once
#include "vmv_device.cuh"

namespace vmv
{
    struct RobotDev
    {
        const SphereTask *tasks;  // kTasks
        const LinkInfo *links;    // kLinks
        const LinkPair *pairs;    // kPairs
        const int *attach_links;  // kAttachLinks
        const PairInfo *pair_info;    // kPairs
        const SpherePair *pair_lists; // statically pruned fine pairs
        int n_pair_lists;
    };

    struct LaunchEnv
    {
        const float *blob;    // device copy of the packed environment
        uint32_t blob_bytes;  // multiple of 16
        float attach_tf[12];  // end-effector frame (in its body) * attachment offset, row-major 3x4
        uint32_t n_objects;   // spheres + capsules + cuboids
        bool primitives_only; // no heightfield / pointcloud / attachment
        // World-fixed links (rigid body 0: Fetch's base and torso column, Baxter's torso, head and pedestal) meet the
        // environment the same way in every configuration: the host checks them ONCE per environment and robot
        // (mode 2, one launch of one state) and every later launch skips them (mode 1).  0 = check everything.
        uint32_t static_mode;
    };

    // A body frame is a rigid transform [R | t] (row-major 3x4, K = 4*row + col).  The stash keeps the
    // first two columns of R and t (9 floats); the third column is their cross product.
    static constexpr int kFrameFloats = 9;

    __host__ __device__ constexpr int frame_slot(int k)
    {
        return (k / 4) * 3 + ((k % 4) == 3 ? 2 : (k % 4));
    }

    // shared-memory carve-up, identical on host (size computation) and device
    template <typename M, int BLOCK>
    struct SmemLayout
    {
        static constexpr int kStashFrames = (M::kBodies - 1) * kFrameFloats;
        static constexpr int kStashBounds = M::kLinks * 3;

        __host__ __device__ static constexpr uint32_t align16(uint32_t v)
        {
            return (v + 15u) & ~15u;
        }

        uint32_t off_tasks, off_links, off_pairs, off_attach, off_stash, total;

        __host__ __device__ explicit SmemLayout(uint32_t blob_bytes)
        {
            uint32_t o = align16(blob_bytes);
            off_tasks = o;
            o += align16(M::kTasks * sizeof(SphereTask));
            off_links = o;
            o += align16(M::kLinks * sizeof(LinkInfo));
            off_pairs = o;
            o += align16((M::kPairs > 0 ? M::kPairs : 1) * sizeof(LinkPair));
            off_attach = o;
            o += align16((M::kAttachLinks > 0 ? M::kAttachLinks : 1) * sizeof(int));
            off_stash = o;
            o += (kStashFrames + kStashBounds) * BLOCK * sizeof(float);
            total = o;
        }
    };

    template <int BLOCK>
    struct StashSink
    {
        static constexpr bool kInlinePairs = true;
        float *base;  // already offset by threadIdx.x

        template <int BODY, int K>
        __device__ __forceinline__ void put(float v)
        {
            if (K % 4 != 2)
            {
                base[((BODY - 1) * kFrameFloats + frame_slot(K)) * BLOCK] = v;
            }
        }

        template <int LINK, int AXIS>
        __device__ __forceinline__ void bound(float)
        {
        }

        __device__ __forceinline__ void in_box(bool)
        {
        }

        __device__ __forceinline__ void reach_ok(bool)
        {
        }

        __device__ __forceinline__ void self_inline(bool)
        {
        }
    };

    struct RegSink
    {
        // used by k_sphere_fk: frames stay in registers/local for an immediate read-back
        static constexpr bool kInlinePairs = false;
        float *f;

        template <int BODY, int K>
        __device__ __forceinline__ void put(float v)
        {
            f[(BODY - 1) * 12 + K] = v;
        }

        template <int LINK, int AXIS>
        __device__ __forceinline__ void bound(float)
        {
        }

        __device__ __forceinline__ void in_box(bool)
        {
        }

        __device__ __forceinline__ void reach_ok(bool)
        {
        }

        __device__ __forceinline__ void self_inline(bool)
        {
        }
    };

    template <int BLOCK>
    __device__ __forceinline__ void
    task_centre(const SphereTask &t, const float *stash, float &x, float &y, float &z)
    {
        if (t.body == 0)
        {
            x = t.cx, y = t.cy, z = t.cz;
        }
        else
        {
            const float *F = stash + (t.body - 1) * kFrameFloats * BLOCK;
            const float r00 = F[0 * BLOCK], r01 = F[1 * BLOCK], tx = F[2 * BLOCK];
            const float r10 = F[3 * BLOCK], r11 = F[4 * BLOCK], ty = F[5 * BLOCK];
            const float r20 = F[6 * BLOCK], r21 = F[7 * BLOCK], tz = F[8 * BLOCK];
            // third column of a rotation = first x second
            const float r02 = fmaf(r10, r21, -(r20 * r11));
            const float r12 = fmaf(r20, r01, -(r00 * r21));
            const float r22 = fmaf(r00, r11, -(r10 * r01));
            x = fmaf(r00, t.cx, fmaf(r01, t.cy, fmaf(r02, t.cz, tx)));
            y = fmaf(r10, t.cx, fmaf(r11, t.cy, fmaf(r12, t.cz, ty)));
            z = fmaf(r20, t.cx, fmaf(r21, t.cy, fmaf(r22, t.cz, tz)));
        }
    }

    // Everything a thread needs to check states: pointers into shared memory.
    template <typename M, int BLOCK>
    struct BlockCtx
    {
        const float *env;
        const SphereTask *tasks;
        const LinkInfo *links;
        const LinkPair *pairs;
        const int *attach_links;
        float *stash;   // frames, offset by threadIdx.x
        float *bounds;  // bounding centres, offset by threadIdx.x
    };

    // Stage tables + environment.  Returns after issuing the copies; call ctx_wait() before phase B.
    template <typename M, int BLOCK>
    __device__ __forceinline__ BlockCtx<M, BLOCK>
    ctx_begin(unsigned char *smem, uint64_t *bar, const RobotDev &robot, const LaunchEnv &env)
    {
        const SmemLayout<M, BLOCK> L(env.blob_bytes);
        if (threadIdx.x == 0)
        {
            mbar_init(bar, 1);
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            tma_bulk_g2s(smem, env.blob, env.blob_bytes, bar);
        }

        // robot tables: a few KB, plain coalesced loads
        {
            uint32_t *dst = reinterpret_cast<uint32_t *>(smem + L.off_tasks);
            const uint32_t *src = reinterpret_cast<const uint32_t *>(robot.tasks);
            for (int i = threadIdx.x; i < M::kTasks * 8; i += BLOCK)
            {
                dst[i] = __ldg(src + i);
            }
            dst = reinterpret_cast<uint32_t *>(smem + L.off_links);
            src = reinterpret_cast<const uint32_t *>(robot.links);
            for (int i = threadIdx.x; i < M::kLinks * 4; i += BLOCK)
            {
                dst[i] = __ldg(src + i);
            }
            dst = reinterpret_cast<uint32_t *>(smem + L.off_pairs);
            src = reinterpret_cast<const uint32_t *>(robot.pairs);
            for (int i = threadIdx.x; i < M::kPairs * 2; i += BLOCK)
            {
                dst[i] = __ldg(src + i);
            }
            dst = reinterpret_cast<uint32_t *>(smem + L.off_attach);
            src = reinterpret_cast<const uint32_t *>(robot.attach_links);
            for (int i = threadIdx.x; i < M::kAttachLinks; i += BLOCK)
            {
                dst[i] = __ldg(src + i);
            }
        }

        BlockCtx<M, BLOCK> c;
        c.env = reinterpret_cast<const float *>(smem);
        c.tasks = reinterpret_cast<const SphereTask *>(smem + L.off_tasks);
        c.links = reinterpret_cast<const LinkInfo *>(smem + L.off_links);
        c.pairs = reinterpret_cast<const LinkPair *>(smem + L.off_pairs);
        c.attach_links = reinterpret_cast<const int *>(smem + L.off_attach);
        c.stash = reinterpret_cast<float *>(smem + L.off_stash) + threadIdx.x;
        c.bounds = c.stash + SmemLayout<M, BLOCK>::kStashFrames * BLOCK;
        return c;
    }

    __device__ __forceinline__ void ctx_wait(uint64_t *bar)
    {
        __syncthreads();  // robot tables visible
        mbar_wait(bar, 0);
    }

    // ------------------------------------------------------------------------------------------
    // phases B, C, D for one state whose frames are in the stash.  `Vote` lets the edge kernel stop
    // a whole warp as soon as any lane has found a collision.
    // ------------------------------------------------------------------------------------------
    struct NoVote
    {
        __device__ __forceinline__ bool stop(bool) const
        {
            return false;
        }
        __device__ __forceinline__ bool any(bool v) const
        {
            return v;
        }
        __device__ __forceinline__ bool rake_any(bool v) const
        {
            return v;
        }
        static constexpr bool kWarp = false;
        static constexpr bool kTogether = false;
    };

    // Configurations, one per lane: every lane is its own unit (no early stop on a neighbour's collision), but the
    // lanes of a warp stay in the loops TOGETHER until the last one is done, and a lane without work keeps calling
    // sphere_hits_env as a helper.  The pointcloud scans inside are cooperative over the lanes that arrive together:
    // with per-lane loops the late lanes -- the ones that went on to fine spheres -- worked through their pointcloud queries
    // in ever smaller groups, a lone lane reading 8 points per step from a list of thousands (measured on BASELINE
    // config 4: the stragglers were 2.3x the average SM time).
    struct LaneVote
    {
        __device__ __forceinline__ bool stop(bool) const
        {
            return false;
        }
        __device__ __forceinline__ bool any(bool v) const
        {
            return __any_sync(0xffffffffu, v);
        }
        __device__ __forceinline__ bool rake_any(bool v) const
        {
            return v;
        }
        static constexpr bool kWarp = false;
        static constexpr bool kTogether = true;
    };

    struct WarpVote
    {
        // every lane of the warp calls these together
        __device__ __forceinline__ bool stop(bool bad) const
        {
            return __any_sync(0xffffffffu, bad);
        }
        __device__ __forceinline__ bool any(bool v) const
        {
            return __any_sync(0xffffffffu, v);
        }
        // OR over the 8 lanes that hold the tines of one rake block
        __device__ __forceinline__ bool rake_any(bool v) const
        {
            const uint32_t b = __ballot_sync(0xffffffffu, v);
            return ((b >> ((threadIdx.x & 31) & 24)) & 0xffu) != 0u;
        }
        static constexpr bool kWarp = true;
        static constexpr bool kTogether = true;
    };

    template <typename M, int BLOCK, typename Vote>
    __device__ __forceinline__ bool
    check_state(const BlockCtx<M, BLOCK> &c, const LaunchEnv &env, bool has_state, const Vote &vote)
    {
        bool bad = false;

        // ---- B: environment -----------------------------------------------------------------
        // Pointcloud queries are not conservative for a bounding sphere (its radius exceeds the
        // tree's r_max, reference capt.hh:428-512 queried from robots/panda.hh:5633), so with a
        // pointcloud in the environment the hierarchy is part of the verdict and is reproduced as
        // the reference runs it: the un-inflated bounding radius, and "any lane of the rake block
        // hits" decides whether the block's fine spheres are swept (robots/panda.hh:5634-5645).
        const bool has_cloud = reinterpret_cast<const EnvHeader *>(c.env)->n_capts > 0 || reinterpret_cast<const EnvHeader *>(c.env)->n_mvts > 0;
        const bool collective = Vote::kWarp && has_cloud;
        {
            int ti = 0;
            bool active = has_state;
            while (vote.any(active))
            {
                // every lane reads a valid task (a lane without work re-reads its first one and contributes nothing):
                // no load in this loop depends on a condition, whichever lanes are still at work
                const SphereTask t = c.tasks[active ? ti : 0];
                float x, y, z;
                task_centre<BLOCK>(t, c.stash, x, y, z);
                bool hit = false;
                // static_mode 1: world-fixed links were checked when the environment was first used; 2: only they are
                const bool skip_task = (env.static_mode == 1u && t.body == 0) || (env.static_mode == 2u && t.body != 0);
                if ((Vote::kTogether && has_cloud) || (active && !skip_task))
                {
                    // (with a pointcloud in the environment every lane of the warp calls: one without a sphere helps the scans)
                    hit = sphere_hits_env(c.env, x, y, z, t.r, t.skip >= 0 ? t.r - 1e-6f : t.r, active && !skip_task);
                }
                if (active)
                {
                    VMV_STAT(t.skip >= 0 ? 12 : 13, 1);
                    VMV_STAT(t.skip >= 0 ? 14 : 15, hit ? 1 : 0);
                }
                if (collective)
                {
                    // lanes of a rake block walk the task list in lock step, so this is the block's bounding verdict
                    const bool block_hit = vote.rake_any(hit && t.skip >= 0);
                    hit = (active && t.skip >= 0) ? block_hit : hit;
                }
                if (active)
                {
                    if (t.skip >= 0)
                    {
                        float *b = c.bounds + t.link * 3 * BLOCK;
                        b[0] = x, b[BLOCK] = y, b[2 * BLOCK] = z;
                        ti = hit ? ti + 1 : t.skip;
                    }
                    else
                    {
                        bad = hit;
                        ti = ti + 1;
                    }
                    active = (ti < M::kTasks) && !bad;
                }
                if (vote.stop(bad))
                {
                    return false;
                }
            }
        }
        if (!Vote::kTogether && bad)
        {
            return false;
        }
        if (env.static_mode == 2u)
        {
            return !bad;  // the one-off check of the world-fixed links: environment only
        }

        // ---- C: self collision --------------------------------------------------------------
        {
            int pi = 0;
            bool active = has_state && !bad && M::kPairs > 0;
            while (vote.any(active))
            {
                if (active)
                {
                    const LinkPair p = c.pairs[pi];
                    const LinkInfo la = c.links[p.a], lb = c.links[p.b];
                    const float *ba = c.bounds + p.a * 3 * BLOCK, *bb = c.bounds + p.b * 3 * BLOCK;
                    const float dx = ba[0] - bb[0], dy = ba[BLOCK] - bb[BLOCK], dz = ba[2 * BLOCK] - bb[2 * BLOCK];
                    const float rs = c.tasks[la.bound_task].r + c.tasks[lb.bound_task].r;
                    if (sign_set((dx * dx + dy * dy + dz * dz) - rs * rs))
                    {
                        // fine pairs (rare): spheres of link a against spheres of link b
                        for (int i = 0; i < la.n_spheres && !bad; ++i)
                        {
                            const SphereTask ta = c.tasks[la.bound_task + 1 + i];
                            float ax, ay, az;
                            task_centre<BLOCK>(ta, c.stash, ax, ay, az);
                            {
                                // link b's bounding sphere encloses its spheres: a sphere of a outside it
                                // touches none of them (the inner loop re-poses every sphere of b; measured on
                                // heightfield scenes: Panda 2.2x, UR5 1.9x, Fetch 3.2x)
                                const float fx = ax - bb[0], fy = ay - bb[BLOCK], fz = az - bb[2 * BLOCK];
                                const float fr = ta.r + c.tasks[lb.bound_task].r;
                                if (!sign_set((fx * fx + fy * fy + fz * fz) - fr * fr))
                                {
                                    continue;
                                }
                            }
                            for (int j = 0; j < lb.n_spheres; ++j)
                            {
                                const SphereTask tb = c.tasks[lb.bound_task + 1 + j];
                                float bx, by, bz;
                                task_centre<BLOCK>(tb, c.stash, bx, by, bz);
                                const float ex = ax - bx, ey = ay - by, ez = az - bz;
                                const float rr = ta.r + tb.r;
                                if (sign_set((ex * ex + ey * ey + ez * ez) - rr * rr))
                                {
                                    bad = true;
                                    break;
                                }
                            }
                        }
                    }
                    pi = pi + 1;
                    active = (pi < M::kPairs) && !bad;
                }
                if (vote.stop(bad))
                {
                    return false;
                }
            }
        }
        if (!Vote::kTogether && bad)
        {
            return false;
        }

        // ---- D: attachment (reference fkcc_attach, robots/panda.hh:15308-15440) --------------
        const EnvHeader &H = *reinterpret_cast<const EnvHeader *>(c.env);
        if (H.n_attach > 0)
        {
            // T = F[ee_body] * attach_tf   (Attachment::pose, collision/attachments.hh:43-55)
            float T[12];
            {
                float F[12];
                if (M::kEeBody == 0 || !has_state)
                {
#pragma unroll
                    for (int k = 0; k < 12; ++k)
                    {
                        F[k] = (k % 5 == 0) ? 1.F : 0.F;
                    }
                }
                else
                {
#pragma unroll
                    for (int k = 0; k < 12; ++k)
                    {
                        F[k] = (k % 4 == 2) ? 0.F : c.stash[((M::kEeBody - 1) * kFrameFloats + frame_slot(k)) * BLOCK];
                    }
                    F[2] = fmaf(F[4], F[9], -(F[8] * F[5]));
                    F[6] = fmaf(F[8], F[1], -(F[0] * F[9]));
                    F[10] = fmaf(F[0], F[5], -(F[4] * F[1]));
                }
#pragma unroll
                for (int i = 0; i < 3; ++i)
                {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        float s = F[4 * i] * env.attach_tf[j] + F[4 * i + 1] * env.attach_tf[4 + j] +
                                  F[4 * i + 2] * env.attach_tf[8 + j];
                        T[4 * i + j] = (j == 3) ? s + F[4 * i + 3] : s;
                    }
                }
            }
            const float4 *S = reinterpret_cast<const float4 *>(c.env + H.off_attach + kAttachHdr);
            // attachment vs environment (validity.hh:259-276)
            for (uint32_t i = 0; i < H.n_attach && ((Vote::kTogether && has_cloud) ? vote.any(has_state && !bad) : (has_state && !bad)); ++i)
            {
                const float4 s = S[i];
                const float x = fmaf(T[0], s.x, fmaf(T[1], s.y, fmaf(T[2], s.z, T[3])));
                const float y = fmaf(T[4], s.x, fmaf(T[5], s.y, fmaf(T[6], s.z, T[7])));
                const float z = fmaf(T[8], s.x, fmaf(T[9], s.y, fmaf(T[10], s.z, T[11])));
                const bool live = has_state && !bad;
                bool h = false;
                if (has_cloud || live)
                {
                    h = sphere_hits_env(c.env, x, y, z, s.w, s.w, live);
                }
                bad = bad || (live && h);
            }
            // attachment vs links, bounding sphere first (validity.hh:278-301); per lane
            for (int k = 0; k < M::kAttachLinks && has_state && !bad; ++k)
            {
                const int l = c.attach_links[k];
                const LinkInfo li = c.links[l];
                const float *b = c.bounds + l * 3 * BLOCK;
                const float br = c.tasks[li.bound_task].r;
                bool near = false;
                for (uint32_t i = 0; i < H.n_attach && !near; ++i)
                {
                    const float4 s = S[i];
                    const float x = fmaf(T[0], s.x, fmaf(T[1], s.y, fmaf(T[2], s.z, T[3])));
                    const float y = fmaf(T[4], s.x, fmaf(T[5], s.y, fmaf(T[6], s.z, T[7])));
                    const float z = fmaf(T[8], s.x, fmaf(T[9], s.y, fmaf(T[10], s.z, T[11])));
                    const float dx = b[0] - x, dy = b[BLOCK] - y, dz = b[2 * BLOCK] - z;
                    const float rs = br + s.w;
                    near = sign_set((dx * dx + dy * dy + dz * dz) - rs * rs);
                }
                if (!near)
                {
                    continue;
                }
                for (int f = 0; f < li.n_spheres && !bad; ++f)
                {
                    const SphereTask tf = c.tasks[li.bound_task + 1 + f];
                    float fx, fy, fz;
                    task_centre<BLOCK>(tf, c.stash, fx, fy, fz);
                    for (uint32_t i = 0; i < H.n_attach; ++i)
                    {
                        const float4 s = S[i];
                        const float x = fmaf(T[0], s.x, fmaf(T[1], s.y, fmaf(T[2], s.z, T[3])));
                        const float y = fmaf(T[4], s.x, fmaf(T[5], s.y, fmaf(T[6], s.z, T[7])));
                        const float z = fmaf(T[8], s.x, fmaf(T[9], s.y, fmaf(T[10], s.z, T[11])));
                        const float dx = fx - x, dy = fy - y, dz = fz - z;
                        const float rs = tf.r + s.w;
                        if (sign_set((dx * dx + dy * dy + dz * dz) - rs * rs))
                        {
                            bad = true;
                            break;
                        }
                    }
                }
            }
            if (vote.stop(bad))
            {
                return false;
            }
        }

        return !bad;
    }

    // ------------------------------------------------------------------------------------------
    // kernels
    // ------------------------------------------------------------------------------------------
    template <typename R, int BLOCK>
    __global__ void __maxnreg__(128)
        k_validate_configs(RobotDev robot, LaunchEnv env, const float *__restrict__ q, size_t n, uint32_t *__restrict__ bits)
    {
        using M = typename R::Model;
        extern __shared__ __align__(128) unsigned char smem[];
        __shared__ uint64_t bar;
        BlockCtx<M, BLOCK> c = ctx_begin<M, BLOCK>(smem, &bar, robot, env);

        const size_t i = static_cast<size_t>(blockIdx.x) * BLOCK + threadIdx.x;
        const bool has = i < n;
        float cfg[M::kDof];
#pragma unroll
        for (int j = 0; j < M::kDof; ++j)
        {
            cfg[j] = has ? __ldg(q + i * M::kDof + j) : 0.F;
        }
        StashSink<BLOCK> sink{c.stash};
        R::frames(cfg, sink);

        ctx_wait(&bar);
        // (every thread calls: the lanes of a warp move through check_state together)
#ifdef VMV_LANEVOTE_OFF  // development A/B
        const bool clear = check_state<M, BLOCK>(c, env, has, NoVote{});
#else
        const bool clear = check_state<M, BLOCK>(c, env, has, LaneVote{});
#endif
        const bool valid = has && clear;
        const uint32_t word = __ballot_sync(0xffffffffu, valid);
        if ((threadIdx.x & 31) == 0 && has)
        {
            bits[i >> 5] = word;
        }
    }

    // FloatVector<dim>::l2_norm() of (b - a) with the reference's summation order
    // (vector/interface.hh:397-410, vector/avx.hh:441-452): rows of 8 lanes summed elementwise, then
    // (l4+l0 + l6+l2) + (l5+l1 + l7+l3).
    template <int DOF>
    __device__ __forceinline__ float ref_l2_norm(const float (&v)[DOF])
    {
        static_assert(DOF <= 16, "two rows of 8 lanes");
        // a second row (Baxter) is added the way GCC contracts row0*row0 + row1*row1 in the reference build:
        // the first product fused onto the rounded second one (read off the compiled reference)
        float lane[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
        {
            if (k + 8 < DOF)
            {
                lane[k] = __fmaf_rn(v[k], v[k], __fmul_rn(v[k + 8 < DOF ? k + 8 : 0], v[k + 8 < DOF ? k + 8 : 0]));
            }
            else
            {
                lane[k] = (k < DOF) ? __fmul_rn(v[k < DOF ? k : 0], v[k < DOF ? k : 0]) : 0.F;
            }
        }
        const float s0 = __fadd_rn(lane[4], lane[0]), s1 = __fadd_rn(lane[5], lane[1]);
        const float s2 = __fadd_rn(lane[6], lane[2]), s3 = __fadd_rn(lane[7], lane[3]);
        return __fsqrt_rn(__fadd_rn(__fadd_rn(s0, s2), __fadd_rn(s1, s3)));
    }

    // One warp per edge.  The reference checks 8 "tines" at (t+1)/8 along the edge, then steps all of
    // them back n-1 times by vector/(8n) (planning/validate.hh:31-64).  A warp covers 4 such steps
    // (32 states) per pass; the edge is invalid as soon as any state is.
    template <typename R, int BLOCK, bool INDEXED>
    __global__ void __maxnreg__(128) k_validate_edges(
        RobotDev robot,
        LaunchEnv env,
        const float *__restrict__ a,
        const float *__restrict__ b,
        const uint32_t *__restrict__ pairs,
        size_t n,
        float resolution,
        uint32_t *__restrict__ bits)
    {
        using M = typename R::Model;
        extern __shared__ __align__(128) unsigned char smem[];
        __shared__ uint64_t bar;
        BlockCtx<M, BLOCK> c = ctx_begin<M, BLOCK>(smem, &bar, robot, env);
        ctx_wait(&bar);

        constexpr int WARPS = BLOCK / 32;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int tine = lane & 7, sub = lane >> 3;
        const WarpVote vote{};

        // each warp owns 32 consecutive edges per round so that lane e can publish the verdict word
        for (size_t base = (static_cast<size_t>(blockIdx.x) * WARPS + warp) * 32; base < n;
             base += static_cast<size_t>(gridDim.x) * WARPS * 32)
        {
            uint32_t word = 0;
            const int count = static_cast<int>(min(static_cast<size_t>(32), n - base));
            for (int e = 0; e < count; ++e)
            {
                const size_t edge = base + e;
                const float *pa, *pb;
                if (INDEXED)
                {
                    pa = a + static_cast<size_t>(__ldg(pairs + 2 * edge)) * M::kDof;
                    pb = a + static_cast<size_t>(__ldg(pairs + 2 * edge + 1)) * M::kDof;
                }
                else
                {
                    pa = a + edge * M::kDof;
                    pb = b + edge * M::kDof;
                }
                float start[M::kDof], vec[M::kDof];
#pragma unroll
                for (int j = 0; j < M::kDof; ++j)
                {
                    start[j] = __ldg(pa + j);
                    vec[j] = __fsub_rn(__ldg(pb + j), start[j]);
                }
                const float dist = ref_l2_norm<M::kDof>(vec);
                // n = max(ceil(distance / rake * resolution), 1)   (validate.hh:41)
                const int steps = static_cast<int>(fmaxf(ceilf(__fmul_rn(__fdiv_rn(dist, 8.F), resolution)), 1.F));
                const float pct = static_cast<float>(tine + 1) / 8.F;
                const float denom = static_cast<float>(8 * steps);
                float cfg[M::kDof], back[M::kDof];
#pragma unroll
                for (int j = 0; j < M::kDof; ++j)
                {
                    cfg[j] = fmaf(vec[j], pct, start[j]);
                    back[j] = __fdiv_rn(vec[j], denom);
                }
                // this lane starts at step `sub`
                for (int s = 0; s < sub && s < steps; ++s)
                {
#pragma unroll
                    for (int j = 0; j < M::kDof; ++j)
                    {
                        cfg[j] = __fsub_rn(cfg[j], back[j]);
                    }
                }

                bool edge_ok = true;
                for (int step0 = 0; step0 < steps; step0 += 4)
                {
                    const bool has = (step0 + sub) < steps;
                    StashSink<BLOCK> sink{c.stash};
                    R::frames(cfg, sink);
                    if (!check_state<M, BLOCK>(c, env, has, vote))
                    {
                        edge_ok = false;
                        break;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                    {
#pragma unroll
                        for (int j = 0; j < M::kDof; ++j)
                        {
                            cfg[j] = __fsub_rn(cfg[j], back[j]);
                        }
                    }
                }
                word |= (edge_ok ? 1u : 0u) << e;
            }
            if (lane == 0)
            {
                bits[base >> 5] = word;
            }
        }
    }

    template <typename R, int BLOCK>
    __global__ void __launch_bounds__(BLOCK)
        k_sphere_fk(RobotDev robot, const float *__restrict__ q, size_t n, float *__restrict__ out)
    {
        using M = typename R::Model;
        const size_t i = static_cast<size_t>(blockIdx.x) * BLOCK + threadIdx.x;
        if (i >= n)
        {
            return;
        }
        float cfg[M::kDof];
#pragma unroll
        for (int j = 0; j < M::kDof; ++j)
        {
            cfg[j] = __ldg(q + i * M::kDof + j);
        }
        float F[(M::kBodies - 1) * 12];
        RegSink sink{F};
        R::frames(cfg, sink);
        for (int t = 0; t < M::kTasks; ++t)
        {
            const SphereTask tk = robot.tasks[t];
            if (tk.sphere < 0)
            {
                continue;
            }
            float x = tk.cx, y = tk.cy, z = tk.cz;
            if (tk.body > 0)
            {
                const float *f = F + (tk.body - 1) * 12;
                x = fmaf(f[0], tk.cx, fmaf(f[1], tk.cy, fmaf(f[2], tk.cz, f[3])));
                y = fmaf(f[4], tk.cx, fmaf(f[5], tk.cy, fmaf(f[6], tk.cz, f[7])));
                z = fmaf(f[8], tk.cx, fmaf(f[9], tk.cy, fmaf(f[10], tk.cz, f[11])));
            }
            float4 *o = reinterpret_cast<float4 *>(out + (i * M::kSpheres + tk.sphere) * 4);
            *o = make_float4(x, y, z, tk.r);
        }
    }

    // Helper::filter_self_from_pointcloud (reference bindings/robot_helper.hh:284-322): point k
    // survives unless its sphere (radius r_point) overlaps one of the robot's fine spheres at
    // configuration q (sphere_sphere_sql2 < 0) or collides with the environment.  Streaming: 12 B in,
    // 1 bit out per point; the robot is posed once per block into shared memory.
    template <typename R, int BLOCK>
    __global__ void __launch_bounds__(BLOCK) k_filter_points(
        RobotDev robot,
        LaunchEnv env,
        const float *__restrict__ q,
        const float *__restrict__ pts,
        size_t n,
        float r_point,
        uint32_t *__restrict__ keep_bits)
    {
        using M = typename R::Model;
        extern __shared__ __align__(128) unsigned char smem[];
        __shared__ uint64_t bar;
        float4 *spheres = reinterpret_cast<float4 *>(smem + ((env.blob_bytes + 15u) & ~15u));
        if (threadIdx.x == 0)
        {
            mbar_init(&bar, 1);
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            tma_bulk_g2s(smem, env.blob, env.blob_bytes, &bar);
        }
        {
            float cfg[M::kDof];
#pragma unroll
            for (int j = 0; j < M::kDof; ++j)
            {
                cfg[j] = __ldg(q + j);
            }
            float F[(M::kBodies - 1) * 12];
            RegSink sink{F};
            R::frames(cfg, sink);
            for (int t = threadIdx.x; t < M::kTasks; t += BLOCK)
            {
                const SphereTask tk = robot.tasks[t];
                if (tk.sphere < 0)
                {
                    continue;
                }
                float x = tk.cx, y = tk.cy, z = tk.cz;
                if (tk.body > 0)
                {
                    const float *f = F + (tk.body - 1) * 12;
                    x = fmaf(f[0], tk.cx, fmaf(f[1], tk.cy, fmaf(f[2], tk.cz, f[3])));
                    y = fmaf(f[4], tk.cx, fmaf(f[5], tk.cy, fmaf(f[6], tk.cz, f[7])));
                    z = fmaf(f[8], tk.cx, fmaf(f[9], tk.cy, fmaf(f[10], tk.cz, f[11])));
                }
                spheres[tk.sphere] = make_float4(x, y, z, tk.r);
            }
        }
        __syncthreads();
        mbar_wait(&bar, 0);
        const float *E = reinterpret_cast<const float *>(smem);

        const size_t n_tiles = (n + BLOCK - 1) / BLOCK;
        for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        {
            const size_t i = tile * BLOCK + threadIdx.x;
            const bool has = i < n;
            float x = 0.F, y = 0.F, z = 0.F;
            if (has)
            {
                x = __ldg(pts + 3 * i), y = __ldg(pts + 3 * i + 1), z = __ldg(pts + 3 * i + 2);
            }
            bool valid = has;
            for (int s = 0; s < M::kSpheres; ++s)
            {
                const float4 a = spheres[s];
                const float dx = a.x - x, dy = a.y - y, dz = a.z - z;
                const float rs = a.w + r_point;
                valid = valid && !((dx * dx + dy * dy + dz * dz) - rs * rs < 0.F);
            }
            // every lane calls: the pointcloud scans inside are warp-cooperative
            const bool hit = sphere_hits_env(E, x, y, z, r_point, r_point, valid);
            valid = valid && !hit;
            const uint32_t word = __ballot_sync(0xffffffffu, valid);
            if ((threadIdx.x & 31) == 0 && has)
            {
                keep_bits[i >> 5] = word;
            }
        }
    }

    // Robot::fkcc_debug for ONE configuration (reference robots/panda.hh:468-5224): every fine sphere
    // against every object with the early-outs of sphere_environment_get_collisions
    // (collision/validity.hh:160-257: primitives and heightfields, no pointclouds) and every allowed
    // sphere pair, non-hierarchical.  One block; thread s owns fine-sphere task s.
    template <typename R, int BLOCK>
    __global__ void __launch_bounds__(BLOCK) k_debug(
        RobotDev robot,
        LaunchEnv env,
        const float *__restrict__ q,
        const int *__restrict__ object_ids,
        int32_t *__restrict__ env_hits,
        uint32_t cap_env,
        int32_t *__restrict__ self_hits,
        uint32_t cap_self,
        uint32_t *__restrict__ counts)
    {
        using M = typename R::Model;
        extern __shared__ __align__(128) unsigned char smem[];
        __shared__ uint64_t bar;
        BlockCtx<M, BLOCK> c = ctx_begin<M, BLOCK>(smem, &bar, robot, env);
        float cfg[M::kDof];
#pragma unroll
        for (int j = 0; j < M::kDof; ++j)
        {
            cfg[j] = __ldg(q + j);
        }
        StashSink<BLOCK> sink{c.stash};
        R::frames(cfg, sink);  // every thread poses the same robot into its own stash column
        ctx_wait(&bar);
        const float *E = c.env;
        const EnvHeader &H = *reinterpret_cast<const EnvHeader *>(E);

        auto report = [&](int sphere, uint32_t packed_index)
        {
            const uint32_t slot = atomicAdd(&counts[0], 1u);
            if (slot < cap_env)
            {
                env_hits[2 * slot] = sphere;
                env_hits[2 * slot + 1] = object_ids[packed_index];
            }
        };

        for (int ti = threadIdx.x; ti < M::kTasks; ti += BLOCK)
        {
            const SphereTask t = c.tasks[ti];
            if (t.sphere < 0)
            {
                continue;
            }
            float x, y, z;
            task_centre<BLOCK>(t, c.stash, x, y, z);
            const float r = t.r, rsq = r * r;
            const float ext = fmaf(__fsqrt_rn(fmaf(x, x, fmaf(y, y, z * z))), 1.0000002f, r);
            uint32_t base = 0;
            const float4 *p = reinterpret_cast<const float4 *>(E + H.off_spheres);
            for (uint32_t i = 0; i < H.n_spheres && E[H.off_spheres + kSphereRec * i + 4] < ext; ++i)
            {
                const float4 a = p[2 * i];
                const float dx = a.x - x, dy = a.y - y, dz = a.z - z, rs = a.w + r;
                if (sign_set((dx * dx + dy * dy + dz * dz) - rs * rs))
                {
                    report(t.sphere, base + i);
                }
            }
            base += H.n_spheres;
            p = reinterpret_cast<const float4 *>(E + H.off_capsules);
            for (uint32_t i = 0; i < H.n_capsules && E[H.off_capsules + kCapsuleRec * i + 8] < ext; ++i)
            {
                const float4 a = p[3 * i], v = p[3 * i + 1];
                const float dot = (x - a.x) * v.x + (y - a.y) * v.y + (z - a.z) * v.z;
                const float cdf = fminf(fmaxf(dot * v.w, 0.F), 1.F);
                const float dx = x - (a.x + v.x * cdf), dy = y - (a.y + v.y * cdf), dz = z - (a.z + v.z * cdf), rs = r + a.w;
                if (sign_set((dx * dx + dy * dy + dz * dz) - rs * rs))
                {
                    report(t.sphere, base + i);
                }
            }
            base += H.n_capsules;
            p = reinterpret_cast<const float4 *>(E + H.off_zcapsules);
            for (uint32_t i = 0; i < H.n_zcapsules && p[2 * i + 1].z < ext; ++i)
            {
                const float4 a = p[2 * i], v = p[2 * i + 1];
                const float cdf = fminf(fmaxf((z - a.z) * v.x * v.y, 0.F), 1.F);
                const float dx = x - a.x, dy = y - a.y, dz = z - (a.z + v.x * cdf), rs = r + a.w;
                if (sign_set((dx * dx + dy * dy + dz * dz) - rs * rs))
                {
                    report(t.sphere, base + i);
                }
            }
            base += H.n_zcapsules;
            p = reinterpret_cast<const float4 *>(E + H.off_cuboids);
            for (uint32_t i = 0; i < H.n_cuboids && p[4 * i + 3].w < ext; ++i)
            {
                const float4 cc = p[4 * i], a1 = p[4 * i + 1], a2 = p[4 * i + 2], a3 = p[4 * i + 3];
                const float xs = x - cc.x, ys = y - cc.y, zs = z - cc.z;
                const float e1 = fmaxf(fabsf(a1.x * xs + a1.y * ys + a1.z * zs) - cc.w, 0.F);
                const float e2 = fmaxf(fabsf(a2.x * xs + a2.y * ys + a2.z * zs) - a1.w, 0.F);
                const float e3 = fmaxf(fabsf(a3.x * xs + a3.y * ys + a3.z * zs) - a2.w, 0.F);
                if (sign_set((e1 * e1 + e2 * e2 + e3 * e3) - rsq))
                {
                    report(t.sphere, base + i);
                }
            }
            base += H.n_cuboids;
            p = reinterpret_cast<const float4 *>(E + H.off_zcuboids);
            for (uint32_t i = 0; i < H.n_zcuboids && p[3 * i].w < ext; ++i)
            {
                const float4 cc = p[3 * i], ax = p[3 * i + 1], h = p[3 * i + 2];
                const float xs = x - cc.x, ys = y - cc.y, zs = z - cc.z;
                const float e1 = fmaxf(fabsf(ax.x * xs + ax.y * ys) - h.x, 0.F);
                const float e2 = fmaxf(fabsf(ax.z * xs + ax.w * ys) - h.y, 0.F);
                const float e3 = fmaxf(fabsf(zs) - h.z, 0.F);
                if (sign_set((e1 * e1 + e2 * e2 + e3 * e3) - rsq))
                {
                    report(t.sphere, base + i);
                }
            }
            base += H.n_zcuboids;
            for (uint32_t i = 0; i < H.n_heightfields; ++i)
            {
                const float4 *hp = reinterpret_cast<const float4 *>(E + H.off_heightfields + kHeightRec * i);
                const float4 a = hp[0], b = hp[1], cc = hp[2];
                const float xi = floorf(fminf(fmaxf(fmaf(a.w, a.x - x, cc.x), 0.F), b.z));
                const float yi = floorf(fminf(fmaxf(fmaf(b.x, a.y - y, cc.y), 0.F), b.w));
                const int index = __float2int_rn(fmaf(yi, b.z, xi));
                const float *data = reinterpret_cast<const float *>(
                    (static_cast<unsigned long long>(__float_as_uint(cc.w)) << 32) | __float_as_uint(cc.z));
                if (sign_set(z - r - fmaf(b.y, __ldg(data + index), a.z)))
                {
                    report(t.sphere, base + i);
                }
            }
        }

        // allowed sphere pairs, flattened over the link pairs
        for (int pi = 0; pi < M::kPairs; ++pi)
        {
            const LinkPair lp = c.pairs[pi];
            const LinkInfo A = c.links[lp.a], B = c.links[lp.b];
            for (int k = threadIdx.x; k < A.n_spheres * B.n_spheres; k += BLOCK)
            {
                const SphereTask ta = c.tasks[A.bound_task + 1 + k / B.n_spheres];
                const SphereTask tb = c.tasks[B.bound_task + 1 + k % B.n_spheres];
                float ax, ay, az, bx, by, bz;
                task_centre<BLOCK>(ta, c.stash, ax, ay, az);
                task_centre<BLOCK>(tb, c.stash, bx, by, bz);
                const float ex = ax - bx, ey = ay - by, ez = az - bz, rr = ta.r + tb.r;
                if (sign_set((ex * ex + ey * ey + ez * ez) - rr * rr))
                {
                    const uint32_t slot = atomicAdd(&counts[1], 1u);
                    if (slot < cap_self)
                    {
                        self_hits[2 * slot] = ta.sphere;
                        self_hits[2 * slot + 1] = tb.sphere;
                    }
                }
            }
        }
    }
}  // namespace vmv
