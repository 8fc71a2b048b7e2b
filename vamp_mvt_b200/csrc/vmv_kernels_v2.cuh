// Block-cooperative configuration kernel (sm_100a) for environments made of primitives
// (spheres, capsules, cuboids; at most 64 objects, no heightfield / pointcloud / attachment --
// those take the generic per-thread path in vmv_kernels.cuh).
//
// The reference walks, per configuration, link by link: bounding sphere vs every object, and on a hit
// every fine sphere of the link vs every object again (robots/panda.hh:5629-6010), then the
// self-collision pairs.  Per-thread that is badly divergent (a free configuration runs 11 sphere
// sweeps, a cluttered one 70).  Here a block of BLOCK configurations goes through the phases
// together and the data-dependent work is re-distributed over the threads of the block:
//
//   A   thread = configuration: straight-line FK; frames -> shared-memory stash, bounding-sphere
//       centres stay in registers.
//   B1  thread = configuration, loops are uniform: for every object (record loaded once), every
//       link's bounding sphere is tested (fully unrolled over links, centres/radii in registers).
//       Result: per link a 64-bit mask of the objects its bounding sphere touches.
//   B2  work items = (configuration, fine sphere) for every link with a non-empty mask, written to a
//       shared queue and processed BLOCK at a time.  A fine sphere is only tested against the
//       objects in its link's mask: a fine sphere lies inside its link's bounding sphere, so it can
//       touch nothing else.  Verdicts are unchanged (the reference's own early-outs are likewise
//       conservative culls).
//   C1  thread = configuration: the allowed link pairs, unrolled, on the register-resident bounding
//       centres -> records (configuration, pair) for the pairs whose bounding spheres overlap.
//   C2  one warp per record: the lanes pose the fine spheres of both links once, then split the
//       |A| x |B| sphere pairs.
#pragma once
#include <type_traits>

#include "vmv_kernels.cuh"

namespace vmv
{
    template <int I>
    using IC = std::integral_constant<int, I>;

    // ---- primitive margins (reference collision/sphere_*.hh); sign bit set <=> collision ----
    __device__ __forceinline__ float margin_sphere(const float4 a, float x, float y, float z, float r)
    {
        const float dx = a.x - x, dy = a.y - y, dz = a.z - z;
        const float rs = a.w + r;
        return (dx * dx + dy * dy + dz * dz) - rs * rs;
    }

    __device__ __forceinline__ float margin_capsule(const float4 a, const float4 v, float x, float y, float z, float r)
    {
        const float dot = (x - a.x) * v.x + (y - a.y) * v.y + (z - a.z) * v.z;
        const float cdf = fminf(fmaxf(dot * v.w, 0.F), 1.F);
        const float dx = x - (a.x + v.x * cdf), dy = y - (a.y + v.y * cdf), dz = z - (a.z + v.z * cdf);
        const float rs = r + a.w;
        return (dx * dx + dy * dy + dz * dz) - rs * rs;
    }

    __device__ __forceinline__ float margin_zcapsule(const float4 a, const float4 v, float x, float y, float z, float r)
    {
        const float dot = (z - a.z) * v.x;
        const float cdf = fminf(fmaxf(dot * v.y, 0.F), 1.F);
        const float dx = x - a.x, dy = y - a.y, dz = z - (a.z + v.x * cdf);
        const float rs = r + a.w;
        return (dx * dx + dy * dy + dz * dz) - rs * rs;
    }

    __device__ __forceinline__ float
    margin_cuboid(const float4 c, const float4 a1, const float4 a2, const float4 a3, float x, float y, float z, float rsq)
    {
        const float xs = x - c.x, ys = y - c.y, zs = z - c.z;
        const float e1 = fmaxf(fabsf(a1.x * xs + a1.y * ys + a1.z * zs) - c.w, 0.F);
        const float e2 = fmaxf(fabsf(a2.x * xs + a2.y * ys + a2.z * zs) - a1.w, 0.F);
        const float e3 = fmaxf(fabsf(a3.x * xs + a3.y * ys + a3.z * zs) - a2.w, 0.F);
        return (e1 * e1 + e2 * e2 + e3 * e3) - rsq;
    }

    __device__ __forceinline__ float
    margin_zcuboid(const float4 c, const float4 ax, const float4 h, float x, float y, float z, float rsq)
    {
        const float xs = x - c.x, ys = y - c.y, zs = z - c.z;
        const float e1 = fmaxf(fabsf(ax.x * xs + ax.y * ys) - h.x, 0.F);
        const float e2 = fmaxf(fabsf(ax.z * xs + ax.w * ys) - h.y, 0.F);
        const float e3 = fmaxf(fabsf(zs) - h.z, 0.F);
        return (e1 * e1 + e2 * e2 + e3 * e3) - rsq;
    }

    // exact test of one sphere against object `og` (index into the concatenation
    // spheres | capsules | z-capsules | cuboids | z-cuboids, each sorted by min_distance)
    __device__ __forceinline__ bool object_hit(const float *__restrict__ E, uint32_t og, float x, float y, float z, float r)
    {
        const EnvHeader &H = *reinterpret_cast<const EnvHeader *>(E);
        if (og < H.n_spheres)
        {
            const float4 *p = reinterpret_cast<const float4 *>(E + H.off_spheres) + 2 * og;
            return sign_set(margin_sphere(p[0], x, y, z, r));
        }
        og -= H.n_spheres;
        if (og < H.n_capsules)
        {
            const float4 *p = reinterpret_cast<const float4 *>(E + H.off_capsules) + 3 * og;
            return sign_set(margin_capsule(p[0], p[1], x, y, z, r));
        }
        og -= H.n_capsules;
        if (og < H.n_zcapsules)
        {
            const float4 *p = reinterpret_cast<const float4 *>(E + H.off_zcapsules) + 2 * og;
            return sign_set(margin_zcapsule(p[0], p[1], x, y, z, r));
        }
        og -= H.n_zcapsules;
        if (og < H.n_cuboids)
        {
            const float4 *p = reinterpret_cast<const float4 *>(E + H.off_cuboids) + 4 * og;
            return sign_set(margin_cuboid(p[0], p[1], p[2], p[3], x, y, z, r * r));
        }
        og -= H.n_cuboids;
        const float4 *p = reinterpret_cast<const float4 *>(E + H.off_zcuboids) + 3 * og;
        return sign_set(margin_zcuboid(p[0], p[1], p[2], x, y, z, r * r));
    }

    // Sink of phase A: frames -> shared stash, bounding centres -> registers.
    template <int BLOCK, int NLINKS, bool INLINE_PAIRS = true>
    struct StashBoundSink
    {
        static constexpr bool kInlinePairs = INLINE_PAIRS;
        float *base;  // stash, already offset by threadIdx.x
        float b[NLINKS][3];

        template <int BODY, int K>
        __device__ __forceinline__ void put(float v)
        {
            if (K % 4 != 2)
            {
                base[((BODY - 1) * kFrameFloats + frame_slot(K)) * BLOCK] = v;
            }
        }

        template <int LINK, int AXIS>
        __device__ __forceinline__ void bound(float v)
        {
            b[LINK][AXIS] = v;
        }

        bool inbox, self_hit, reach_valid;

        __device__ __forceinline__ void reach_ok(bool v)
        {
            reach_valid = v;
        }

        __device__ __forceinline__ void in_box(bool v)
        {
            inbox = v;
        }

        __device__ __forceinline__ void self_inline(bool v)
        {
            self_hit = v;
        }
    };

    template <typename M, int BLOCK>
    struct SmemLayoutV2
    {
        static constexpr int kStashFrames = (M::kBodies - 1) * kFrameFloats;
        // work queues sized for the worst case (every link of every configuration hit), 16 bits each
        static constexpr int kItemCap = BLOCK * M::kSpheres;
        static constexpr int kPairCap = BLOCK * (M::kPairs > 0 ? M::kPairs : 1);
        static_assert(BLOCK <= 128 && M::kPairs < 512 && M::kTasks < 256, "16-bit work-item encoding");
        static constexpr int kGroups = BLOCK / 8;     // phase C2 works in groups of 8 lanes
        static constexpr int kScratchPerGroup = 32;   // float4 slots for posed spheres (16 + 16)
        static constexpr int kMaxPairLists = 256;     // statically pruned sphere pairs kept in smem

        uint32_t off_tasks, off_links, off_pairs, off_pinfo, off_plists, off_stash, off_masks, off_items, off_pairq, off_scratch,
            off_flags, total;

        __host__ __device__ static constexpr uint32_t align16(uint32_t v)
        {
            return (v + 15u) & ~15u;
        }

        __host__ __device__ explicit SmemLayoutV2(uint32_t blob_bytes)
        {
            uint32_t o = align16(blob_bytes);
            off_tasks = o;
            o += align16(M::kTasks * sizeof(SphereTask));
            off_links = o;
            o += align16(M::kLinks * sizeof(LinkInfo));
            off_pairs = o;
            o += align16((M::kPairs > 0 ? M::kPairs : 1) * sizeof(LinkPair));
            off_pinfo = o;
            o += align16((M::kPairs > 0 ? M::kPairs : 1) * sizeof(PairInfo));
            off_plists = o;
            o += align16(kMaxPairLists * sizeof(SpherePair));
            off_stash = o;
            o += kStashFrames * BLOCK * sizeof(float);
            // the B-phase queues (masks, items) and the C-phase queues (pair records, scratch) are never
            // live at the same time: they share one region
            off_masks = o;
            uint32_t b_bytes = M::kLinks * BLOCK * sizeof(unsigned long long);
            off_items = o + b_bytes;
            b_bytes += align16(kItemCap * sizeof(uint16_t));
            off_pairq = o;
            uint32_t c_bytes = align16(kPairCap * sizeof(uint16_t));
            off_scratch = o + c_bytes;
            c_bytes += kGroups * kScratchPerGroup * sizeof(float4);
            o += b_bytes > c_bytes ? b_bytes : c_bytes;
            off_flags = o;
            o += 2 * align16(BLOCK * sizeof(uint32_t)) + 16;
            total = o;
        }
    };

    // Everything the phases need, as pointers into the block's shared memory.
    template <typename R, int BLOCK>
    struct V2Ctx
    {
        const float *E;
        const PairInfo *pinfos;
        const SpherePair *plists;
        const SphereTask *tasks;
        const LinkInfo *links;
        const LinkPair *pairs;
        float *stash;
        unsigned long long *masks;
        uint16_t *items;
        uint16_t *pairq;
        float4 *scratch;
        volatile uint32_t *invalid;  // [BLOCK] result of a pass: 1 = state in collision (or no state)
        volatile uint32_t *inbox;    // [BLOCK]
        uint32_t *counters;          // [0] fine items, [1] pair records
        uint64_t *bar;
    };

    // Issue the environment TMA copy and stage the robot tables.  Data is ready after a
    // __syncthreads() and mbar_wait(bar, 0) -- v2_pass does both after its FK phase.
    template <typename R, int BLOCK>
    __device__ __forceinline__ V2Ctx<R, BLOCK> v2_stage(unsigned char *smem, uint64_t *barp, const RobotDev &robot, const LaunchEnv &env)
    {
        using M = typename R::Model;
        using Lay = SmemLayoutV2<M, BLOCK>;
        const Lay L(env.blob_bytes);
        const int tid = threadIdx.x;
        uint64_t &bar = *barp;
        if (tid == 0)
        {
            mbar_init(&bar, 1);
        }
        __syncthreads();
        if (tid == 0)
        {
            tma_bulk_g2s(smem, env.blob, env.blob_bytes, &bar);
        }
        {
            uint32_t *dst = reinterpret_cast<uint32_t *>(smem + L.off_tasks);
            const uint32_t *src = reinterpret_cast<const uint32_t *>(robot.tasks);
            for (int i = tid; i < M::kTasks * 8; i += BLOCK)
            {
                dst[i] = __ldg(src + i);
            }
            dst = reinterpret_cast<uint32_t *>(smem + L.off_links);
            src = reinterpret_cast<const uint32_t *>(robot.links);
            for (int i = tid; i < M::kLinks * 4; i += BLOCK)
            {
                dst[i] = __ldg(src + i);
            }
            dst = reinterpret_cast<uint32_t *>(smem + L.off_pairs);
            src = reinterpret_cast<const uint32_t *>(robot.pairs);
            for (int i = tid; i < M::kPairs * 2; i += BLOCK)
            {
                dst[i] = __ldg(src + i);
            }
            dst = reinterpret_cast<uint32_t *>(smem + L.off_pinfo);
            src = reinterpret_cast<const uint32_t *>(robot.pair_info);
            for (int i = tid; i < M::kPairs * 3; i += BLOCK)
            {
                dst[i] = __ldg(src + i);
            }
            uint16_t *dl = reinterpret_cast<uint16_t *>(smem + L.off_plists);
            const uint16_t *sl = reinterpret_cast<const uint16_t *>(robot.pair_lists);
            for (int i = tid; i < min(robot.n_pair_lists, Lay::kMaxPairLists); i += BLOCK)
            {
                dl[i] = __ldg(sl + i);
            }
        }
        V2Ctx<R, BLOCK> X;
        X.E = reinterpret_cast<const float *>(smem);
        X.pinfos = reinterpret_cast<const PairInfo *>(smem + L.off_pinfo);
        X.plists = reinterpret_cast<const SpherePair *>(smem + L.off_plists);
        X.tasks = reinterpret_cast<const SphereTask *>(smem + L.off_tasks);
        X.links = reinterpret_cast<const LinkInfo *>(smem + L.off_links);
        X.pairs = reinterpret_cast<const LinkPair *>(smem + L.off_pairs);
        X.stash = reinterpret_cast<float *>(smem + L.off_stash);
        X.masks = reinterpret_cast<unsigned long long *>(smem + L.off_masks);
        X.items = reinterpret_cast<uint16_t *>(smem + L.off_items);
        X.pairq = reinterpret_cast<uint16_t *>(smem + L.off_pairq);
        X.scratch = reinterpret_cast<float4 *>(smem + L.off_scratch);
        X.invalid = reinterpret_cast<volatile uint32_t *>(smem + L.off_flags);
        X.inbox = X.invalid + BLOCK;
        X.counters = reinterpret_cast<uint32_t *>(smem + L.off_flags + 2 * Lay::align16(BLOCK * sizeof(uint32_t)));
        X.bar = barp;
        return X;
    }

    // One pass: up to BLOCK states (one per thread; `has` false = idle lane) through phases A..C2.
    // Must be called by every thread of the block.  On return X.invalid[tid] holds the verdict of
    // this thread's state and all shared-memory traffic of the pass is complete.
    template <typename R, int BLOCK>
    __device__ __forceinline__ void v2_pass(const V2Ctx<R, BLOCK> &X, const float (&cfg)[R::Model::kDof], const bool has)
    {
        using M = typename R::Model;
        using Lay = SmemLayoutV2<M, BLOCK>;
        const int tid = threadIdx.x;
        const float *E = X.E;
        const PairInfo *pinfos = X.pinfos;
        const SpherePair *plists = X.plists;
        const SphereTask *tasks = X.tasks;
        const LinkInfo *links = X.links;
        const LinkPair *pairs = X.pairs;
        float *stash = X.stash;
        unsigned long long *masks = X.masks;
        uint16_t *items = X.items;
        uint16_t *pairq = X.pairq;
        float4 *scratch = X.scratch;
        volatile uint32_t *invalid = X.invalid;
        volatile uint32_t *inbox = X.inbox;
        uint32_t *counters = X.counters;

        // ---- A: FK ------------------------------------------------------------------------------
        StashBoundSink<BLOCK, M::kLinks> sink;
        sink.base = stash + tid;
        R::frames(cfg, sink);
        // the ungated fine pairs of the always-overlapping link pairs were evaluated with the FK;
        // their (pruned) list is only valid inside the joint box
        invalid[tid] = (has && !(sink.inbox && sink.self_hit)) ? 0u : 1u;
        inbox[tid] = sink.inbox ? 1u : 0u;
        if (tid < 2)
        {
            counters[tid] = 0;
        }

        __syncthreads();  // tables, flags, counters
        mbar_wait(X.bar, 0);
        const EnvHeader &H = *reinterpret_cast<const EnvHeader *>(E);

        // ---- B1: bounding spheres vs every object, object-major, links unrolled ------------------
        uint32_t mlo[M::kLinks], mhi[M::kLinks];
#pragma unroll
        for (int l = 0; l < M::kLinks; ++l)
        {
            mlo[l] = 0u, mhi[l] = 0u;
        }
        if (has)
        {
            uint32_t og = 0;
            {
                const float4 *p = reinterpret_cast<const float4 *>(E + H.off_spheres);
                for (uint32_t k = 0; k < H.n_spheres; ++k, ++og)
                {
                    const float4 a = p[2 * k];
                    const uint32_t blo = og < 32 ? 1u << og : 0u, bhi = og < 32 ? 0u : 1u << (og - 32);
                    R::for_each_link(
                        [&](auto l, float br, int, int, float reach)
                        {
                            constexpr int li = decltype(l)::value;
                            if (!((E[H.off_spheres + kSphereRec * k + 4]) < reach) && sink.reach_valid)
                            {
                                return;  // out of this link's reach for every configuration
                            }
                            const bool hit = sign_set(margin_sphere(a, sink.b[li][0], sink.b[li][1], sink.b[li][2], br));
                            mlo[li] |= hit ? blo : 0u;
                            mhi[li] |= hit ? bhi : 0u;
                        });
                }
            }
            {
                const float4 *p = reinterpret_cast<const float4 *>(E + H.off_capsules);
                for (uint32_t k = 0; k < H.n_capsules; ++k, ++og)
                {
                    const float4 a = p[3 * k], v = p[3 * k + 1];
                    const uint32_t blo = og < 32 ? 1u << og : 0u, bhi = og < 32 ? 0u : 1u << (og - 32);
                    R::for_each_link(
                        [&](auto l, float br, int, int, float reach)
                        {
                            constexpr int li = decltype(l)::value;
                            if (!((E[H.off_capsules + kCapsuleRec * k + 8]) < reach) && sink.reach_valid)
                            {
                                return;  // out of this link's reach for every configuration
                            }
                            const bool hit = sign_set(margin_capsule(a, v, sink.b[li][0], sink.b[li][1], sink.b[li][2], br));
                            mlo[li] |= hit ? blo : 0u;
                            mhi[li] |= hit ? bhi : 0u;
                        });
                }
            }
            {
                const float4 *p = reinterpret_cast<const float4 *>(E + H.off_zcapsules);
                for (uint32_t k = 0; k < H.n_zcapsules; ++k, ++og)
                {
                    const float4 a = p[2 * k], v = p[2 * k + 1];
                    const uint32_t blo = og < 32 ? 1u << og : 0u, bhi = og < 32 ? 0u : 1u << (og - 32);
                    R::for_each_link(
                        [&](auto l, float br, int, int, float reach)
                        {
                            constexpr int li = decltype(l)::value;
                            if (!((v.z) < reach) && sink.reach_valid)
                            {
                                return;  // out of this link's reach for every configuration
                            }
                            const bool hit = sign_set(margin_zcapsule(a, v, sink.b[li][0], sink.b[li][1], sink.b[li][2], br));
                            mlo[li] |= hit ? blo : 0u;
                            mhi[li] |= hit ? bhi : 0u;
                        });
                }
            }
            {
                const float4 *p = reinterpret_cast<const float4 *>(E + H.off_cuboids);
                for (uint32_t k = 0; k < H.n_cuboids; ++k, ++og)
                {
                    const float4 c = p[4 * k], a1 = p[4 * k + 1], a2 = p[4 * k + 2], a3 = p[4 * k + 3];
                    const uint32_t blo = og < 32 ? 1u << og : 0u, bhi = og < 32 ? 0u : 1u << (og - 32);
                    R::for_each_link(
                        [&](auto l, float br, int, int, float reach)
                        {
                            constexpr int li = decltype(l)::value;
                            if (!((a3.w) < reach) && sink.reach_valid)
                            {
                                return;  // out of this link's reach for every configuration
                            }
                            const bool hit =
                                sign_set(margin_cuboid(c, a1, a2, a3, sink.b[li][0], sink.b[li][1], sink.b[li][2], br * br));
                            mlo[li] |= hit ? blo : 0u;
                            mhi[li] |= hit ? bhi : 0u;
                        });
                }
            }
            {
                const float4 *p = reinterpret_cast<const float4 *>(E + H.off_zcuboids);
                for (uint32_t k = 0; k < H.n_zcuboids; ++k, ++og)
                {
                    const float4 c = p[3 * k], ax = p[3 * k + 1], h = p[3 * k + 2];
                    const uint32_t blo = og < 32 ? 1u << og : 0u, bhi = og < 32 ? 0u : 1u << (og - 32);
                    R::for_each_link(
                        [&](auto l, float br, int, int, float reach)
                        {
                            constexpr int li = decltype(l)::value;
                            if (!((c.w) < reach) && sink.reach_valid)
                            {
                                return;  // out of this link's reach for every configuration
                            }
                            const bool hit =
                                sign_set(margin_zcuboid(c, ax, h, sink.b[li][0], sink.b[li][1], sink.b[li][2], br * br));
                            mlo[li] |= hit ? blo : 0u;
                            mhi[li] |= hit ? bhi : 0u;
                        });
                }
            }

            // enqueue fine-sphere items of the links whose bounding sphere touches something
            R::for_each_link(
                [&](auto l, float, int n_fine, int first_task, float)
                {
                    constexpr int li = decltype(l)::value;
                    if ((mlo[li] | mhi[li]) != 0u)
                    {
                        masks[li * BLOCK + tid] = (static_cast<unsigned long long>(mhi[li]) << 32) | mlo[li];
                        const uint32_t base = atomicAdd(&counters[0], static_cast<uint32_t>(n_fine));
                        for (int k = 0; k < n_fine; ++k)
                        {
                            items[base + k] = static_cast<uint16_t>((tid << 8) | (first_task + k));
                        }
                    }
                });
        }
        __syncthreads();

        // ---- B2: fine spheres, one item per thread per round --------------------------------------
        {
            const uint32_t n_items = counters[0];
            for (uint32_t it = tid; it < n_items; it += BLOCK)
            {
                const uint32_t item = items[it];
                const int c = item >> 8;
                if (invalid[c])
                {
                    continue;
                }
                const SphereTask t = tasks[item & 255u];
                float x, y, z;
                task_centre<BLOCK>(t, stash + c, x, y, z);
                unsigned long long m = masks[t.link * BLOCK + c];
                while (m != 0ull)
                {
                    const int o = __ffsll(static_cast<long long>(m)) - 1;
                    m &= m - 1ull;
                    if (object_hit(E, static_cast<uint32_t>(o), x, y, z, t.r))
                    {
                        invalid[c] = 1u;
                        break;
                    }
                }
            }
        }

        __syncthreads();  // the C-phase queues reuse the B-phase queues' memory

        // ---- C1: allowed link pairs on the bounding spheres ---------------------------------------
        if (has && !invalid[tid])
        {
            float brad[M::kLinks];
            R::for_each_link([&](auto l, float br, int, int, float) { brad[decltype(l)::value] = br; });
            R::for_each_pair(
                [&](auto pi, auto la, auto lb, auto inl)
                {
                    constexpr int a = decltype(la)::value, b = decltype(lb)::value;
                    if (decltype(inl)::value != 0 && sink.inbox)
                    {
                        return;  // already checked in phase A
                    }
                    const float dx = sink.b[a][0] - sink.b[b][0], dy = sink.b[a][1] - sink.b[b][1], dz = sink.b[a][2] - sink.b[b][2];
                    const float rs = brad[a] + brad[b];
                    if (sign_set((dx * dx + dy * dy + dz * dz) - rs * rs))
                    {
                        const uint32_t slot = atomicAdd(&counters[1], 1u);
                        pairq[slot] = static_cast<uint16_t>((tid << 9) | decltype(pi)::value);
                    }
                });
        }
        __syncthreads();

        // ---- C2: fine sphere pairs, one group of 8 lanes per (configuration, pair) record ---------
        // (four records in flight per warp: the work per record is a short chain of dependent
        // shared-memory reads, so concurrency across records is what hides the latency)
        {
            const uint32_t n_rec = counters[1];
            const int lane = tid & 31, gl = tid & 7, group = tid >> 3;
            const uint32_t gmask = 0xffu << (lane & 24);
            float4 *my = scratch + group * Lay::kScratchPerGroup;
            constexpr int HALF = Lay::kScratchPerGroup / 2;
            for (uint32_t r = group; r < n_rec; r += Lay::kGroups)
            {
                const uint32_t rec = pairq[r];
                const int c = rec >> 9;
                if (invalid[c])
                {
                    continue;  // group-uniform
                }
                const int pair = rec & 0x1ffu;
                const PairInfo pinfo = pinfos[pair];
                bool h = false;
                if (pinfo.count >= 0 && inbox[c] && pinfo.offset + pinfo.count <= Lay::kMaxPairLists)
                {
                    // statically pruned list: one feasible sphere pair per lane
                    for (int k = gl; k < pinfo.count; k += 8)
                    {
                        const SpherePair sp = plists[pinfo.offset + k];
                        const SphereTask ta = tasks[sp.task_a], tb = tasks[sp.task_b];
                        float ax, ay, az, bx, by, bz;
                        task_centre<BLOCK>(ta, stash + c, ax, ay, az);
                        task_centre<BLOCK>(tb, stash + c, bx, by, bz);
                        const float ex = ax - bx, ey = ay - by, ez = az - bz;
                        const float rr = ta.r + tb.r;
                        h |= sign_set((ex * ex + ey * ey + ez * ez) - rr * rr);
                    }
                }
                else
                {
                    // full cross product, tiled HALF x HALF through the group's scratch
                    const LinkPair p = pairs[pair];
                    const LinkInfo A = links[p.a], B = links[p.b];
                    for (int a0 = 0; a0 < A.n_spheres; a0 += HALF)
                    {
                        const int nac = min(A.n_spheres - a0, HALF);
                        for (int b0 = 0; b0 < B.n_spheres; b0 += HALF)
                        {
                            const int nbc = min(B.n_spheres - b0, HALF);
                            __syncwarp(gmask);
                            for (int sidx = gl; sidx < nac + nbc; sidx += 8)
                            {
                                const int ti = sidx < nac ? A.bound_task + 1 + a0 + sidx : B.bound_task + 1 + b0 + (sidx - nac);
                                const SphereTask t = tasks[ti];
                                float x, y, z;
                                task_centre<BLOCK>(t, stash + c, x, y, z);
                                my[sidx < nac ? sidx : HALF + (sidx - nac)] = make_float4(x, y, z, t.r);
                            }
                            __syncwarp(gmask);
                            for (int ia = 0; ia < nac; ++ia)
                            {
                                const float4 sa = my[ia];
                                for (int ib = gl; ib < nbc; ib += 8)
                                {
                                    const float4 sb = my[HALF + ib];
                                    const float ex = sa.x - sb.x, ey = sa.y - sb.y, ez = sa.z - sb.z;
                                    const float rr = sa.w + sb.w;
                                    h |= sign_set((ex * ex + ey * ey + ez * ez) - rr * rr);
                                }
                            }
                        }
                    }
                }
                if (h)
                {
                    invalid[c] = 1u;
                }
            }
        }
        __syncthreads();

    }

    template <typename R, int BLOCK>
    __global__ void __launch_bounds__(BLOCK)
        k_validate_configs_v2(RobotDev robot, LaunchEnv env, const float *__restrict__ q, size_t n, uint32_t *__restrict__ bits)
    {
        using M = typename R::Model;
        extern __shared__ __align__(128) unsigned char smem[];
        __shared__ uint64_t bar;
        const V2Ctx<R, BLOCK> X = v2_stage<R, BLOCK>(smem, &bar, robot, env);
        const int tid = threadIdx.x;
        const size_t i = static_cast<size_t>(blockIdx.x) * BLOCK + tid;
        const bool has = i < n;
        float cfg[M::kDof];
#pragma unroll
        for (int j = 0; j < M::kDof; ++j)
        {
            cfg[j] = has ? __ldg(q + i * M::kDof + j) : 0.F;
        }
        v2_pass<R, BLOCK>(X, cfg, has);
        const bool valid = has && !X.invalid[tid];
        const uint32_t word = __ballot_sync(0xffffffffu, valid);
        if ((tid & 31) == 0 && has)
        {
            bits[i >> 5] = word;
        }
    }

    // ------------------------------------------------------------------------------------------
    // Edges, block-cooperative.  A block takes 32 edges (one verdict word) at a time.  The states of
    // an edge are the reference's rake blocks (planning/validate.hh:31-64): block s holds the 8 tines
    // (t+1)/8 stepped back s times by vector/(8n).  Each pass checks 16 rake blocks = 128 states; blocks
    // are handed out round-robin over the edges still alive, so an edge found invalid drops its
    // remaining blocks -- the counterpart of the reference's early return.
    // ------------------------------------------------------------------------------------------
    template <typename R, int BLOCK, bool INDEXED>
    __global__ void __launch_bounds__(BLOCK) k_validate_edges_v2(
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
        static_assert(BLOCK == 128, "16 rake blocks of 8 tines per pass");
        extern __shared__ __align__(128) unsigned char smem[];
        __shared__ uint64_t bar;
        __shared__ float s_start[32][M::kDof], s_vec[32][M::kDof];
        __shared__ int s_steps[32], s_next[32], s_dead[32];
        __shared__ int s_slot_edge[16], s_slot_step[16], s_nslots;
        const V2Ctx<R, BLOCK> X = v2_stage<R, BLOCK>(smem, &bar, robot, env);
        const int tid = threadIdx.x, lane = tid & 31;
        const size_t n_chunks = (n + 31) / 32;

        for (size_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x)
        {
            __syncthreads();
            if (tid < 32)
            {
                const size_t edge = chunk * 32 + tid;
                int steps = 0;
                if (edge < n)
                {
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
                    float vec[M::kDof];
#pragma unroll
                    for (int j = 0; j < M::kDof; ++j)
                    {
                        const float st = __ldg(pa + j);
                        vec[j] = __fsub_rn(__ldg(pb + j), st);
                        s_start[tid][j] = st;
                        s_vec[tid][j] = vec[j];
                    }
                    const float dist = ref_l2_norm<M::kDof>(vec);
                    // n = max(ceil(distance / rake * resolution), 1)   (validate.hh:41)
                    steps = static_cast<int>(fmaxf(ceilf(__fmul_rn(__fdiv_rn(dist, 8.F), resolution)), 1.F));
                }
                s_steps[tid] = steps;
                s_next[tid] = 0;
                s_dead[tid] = 0;
            }
            __syncthreads();

            while (true)
            {
                // ---- hand out up to 16 rake blocks, round-robin over the live edges (warp 0) ------
                if (tid < 32)
                {
                    const int rem = s_dead[tid] ? 0 : s_steps[tid] - s_next[tid];
                    int base = 0, got = 0;
                    for (int r = 0; base < 16; ++r)
                    {
                        const uint32_t m = __ballot_sync(0xffffffffu, rem > r);
                        if (m == 0u)
                        {
                            break;
                        }
                        const int pos = base + __popc(m & ((1u << lane) - 1u));
                        if (rem > r && pos < 16)
                        {
                            s_slot_edge[pos] = tid;
                            s_slot_step[pos] = s_next[tid] + r;
                            ++got;
                        }
                        base += __popc(m);
                    }
                    s_next[tid] += got;
                    if (tid == 0)
                    {
                        s_nslots = min(base, 16);
                    }
                }
                __syncthreads();
                const int nslots = s_nslots;
                if (nslots == 0)
                {
                    break;
                }
                const int slot = tid >> 3, tine = tid & 7;
                const bool has = slot < nslots;
                const int e = has ? s_slot_edge[slot] : 0;
                const int step = has ? s_slot_step[slot] : 0;
                float cfg[M::kDof];
                {
                    const float pct = static_cast<float>(tine + 1) / 8.F;
                    const float denom = static_cast<float>(8 * s_steps[e]);
#pragma unroll
                    for (int j = 0; j < M::kDof; ++j)
                    {
                        const float v = s_vec[e][j];
                        const float back = __fdiv_rn(v, denom);
                        float c = fmaf(v, pct, s_start[e][j]);
                        for (int k = 0; k < step; ++k)
                        {
                            c = __fsub_rn(c, back);
                        }
                        cfg[j] = c;
                    }
                }
                v2_pass<R, BLOCK>(X, cfg, has);
                if (has && X.invalid[tid])
                {
                    s_dead[e] = 1;
                }
                __syncthreads();
            }

            if (tid < 32)
            {
                const bool ok = s_steps[tid] > 0 && !s_dead[tid];
                const uint32_t word = __ballot_sync(0xffffffffu, ok);
                if (tid == 0)
                {
                    bits[chunk] = word;
                }
            }
        }
    }
}  // namespace vmv
