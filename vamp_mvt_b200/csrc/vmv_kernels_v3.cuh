// Grid-culled block-cooperative kernels (sm_100a) for environments made of primitives
// (spheres, capsules, cuboids; at most 64 objects).
//
// vmv_kernels_v2.cuh tests every link's bounding sphere against every object in reach: ~130 exact
// tests per Panda configuration on the bench scene, of which fewer than one hits.  Here the
// environment is rasterised once, when it is first used with a robot, into a voxel table in HBM
// (L2-resident, a few MB): per voxel and per radius class a 64-bit mask of the objects that come
// within (class radius + voxel half diagonal) of the voxel centre.  One 8-byte load per link then
// replaces the object sweep; the mask is a conservative superset, and only its members (1-3 per
// configuration) go through the exact margin test.  Verdicts are those of the exact tests, i.e.
// unchanged (reference collision/validity.hh:46-158 is likewise "cull, then test").
//
// All primitives share one record, a rounded box {centre, 3 axes, 3 half extents, rounding
// radius}: margin = sum_i max(|a_i.(p-c)| - h_i, 0)^2 - (r + rho)^2.  A cuboid has rho = 0
// (reference collision/sphere_cuboid.hh:9-26), a sphere h = 0 (sphere_sphere.hh:10-23), a capsule is
// its segment (h = (len/2, 0, 0)) rounded by its radius -- the same distance as the clamped
// projection of sphere_capsule.hh:9-22.  One record type = no divergence on the object kind.
//
// Phases of a pass over BLOCK states:
//   A    thread = state: FK -> frame stash; bounding centres in registers; inline self pairs
//   B0   thread = state: one voxel-table load per link -> candidate mask; (state, link) items -> Q1
//   B1   Q1 items: bounding sphere vs its candidates (exact) -> hit mask; fine items -> Q2
//   B2   Q2 items: fine sphere vs its link's hit mask (exact)
//   C1   thread = state: allowed link pairs on the bounding spheres -> (state, pair) records
//   C2   thread = record: fine sphere pairs, the smaller link in registers
#pragma once
#include "vmv_kernels_v2.cuh"

namespace vmv
{
    static constexpr int kGridClasses = 4;
    static constexpr int kGridMaxLinks = 64;
    static constexpr int kObjRec = 16;  // floats: {cx cy cz rho}{a1 h1}{a2 h2}{a3 h3}

    struct GridDev
    {
        const unsigned long long *masks;  // [nz][ny][nx][kGridClasses]
        float x0, y0, z0, inv_h;
        int nx, ny, nz;
        unsigned long long all_mask;
        unsigned char link_class[kGridMaxLinks];
    };

    struct LaunchEnvV3
    {
        const float4 *objs;  // n_objects rounded-box records
        uint32_t n_objects;
        uint32_t max_fine;   // largest number of fine spheres of any link of the robot
        GridDev grid;
    };

    __device__ __forceinline__ float margin_obj(const float4 *__restrict__ o, float x, float y, float z, float r)
    {
        const float4 c = o[0], a1 = o[1], a2 = o[2], a3 = o[3];
        const float xs = x - c.x, ys = y - c.y, zs = z - c.z;
        const float e1 = fmaxf(fabsf(a1.x * xs + a1.y * ys + a1.z * zs) - a1.w, 0.F);
        const float e2 = fmaxf(fabsf(a2.x * xs + a2.y * ys + a2.z * zs) - a2.w, 0.F);
        const float e3 = fmaxf(fabsf(a3.x * xs + a3.y * ys + a3.z * zs) - a3.w, 0.F);
        const float rs = r + c.w;
        return (e1 * e1 + e2 * e2 + e3 * e3) - rs * rs;
    }

    // Rasterise the environment: thread = voxel.  An object is a candidate of class k in voxel v iff
    // its distance from the voxel centre is <= class_r[k] + slack, slack = half diagonal + 1e-4:
    // the distance function is 1-Lipschitz, so a sphere of radius <= class_r[k] centred anywhere in
    // the voxel cannot touch an object outside the mask.
    __global__ void __launch_bounds__(128) k_build_grid(
        const float4 *__restrict__ objs,
        uint32_t n_objects,
        float x0,
        float y0,
        float z0,
        float h,
        int nx,
        int ny,
        int nz,
        float r0,
        float r1,
        float r2,
        float r3,
        unsigned long long *__restrict__ out)
    {
        const size_t n_vox = static_cast<size_t>(nx) * ny * nz;
        const size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
        if (v >= n_vox)
        {
            return;
        }
        const int ix = static_cast<int>(v % nx), iy = static_cast<int>((v / nx) % ny), iz = static_cast<int>(v / (static_cast<size_t>(nx) * ny));
        const float x = x0 + (ix + 0.5F) * h, y = y0 + (iy + 0.5F) * h, z = z0 + (iz + 0.5F) * h;
        const float slack = 0.8660255F * h + 1e-4F;
        unsigned long long m0 = 0ull, m1 = 0ull, m2 = 0ull, m3 = 0ull;
        for (uint32_t k = 0; k < n_objects; ++k)
        {
            const float4 c = __ldg(objs + 4 * k), a1 = __ldg(objs + 4 * k + 1), a2 = __ldg(objs + 4 * k + 2), a3 = __ldg(objs + 4 * k + 3);
            const float xs = x - c.x, ys = y - c.y, zs = z - c.z;
            const float e1 = fmaxf(fabsf(a1.x * xs + a1.y * ys + a1.z * zs) - a1.w, 0.F);
            const float e2 = fmaxf(fabsf(a2.x * xs + a2.y * ys + a2.z * zs) - a2.w, 0.F);
            const float e3 = fmaxf(fabsf(a3.x * xs + a3.y * ys + a3.z * zs) - a3.w, 0.F);
            const float d = sqrtf(e1 * e1 + e2 * e2 + e3 * e3) - c.w - slack;
            const unsigned long long bit = 1ull << k;
            // !(d > r) also keeps an object whose record is not finite
            m0 |= !(d > r0) ? bit : 0ull;
            m1 |= !(d > r1) ? bit : 0ull;
            m2 |= !(d > r2) ? bit : 0ull;
            m3 |= !(d > r3) ? bit : 0ull;
        }
        ulonglong2 *o = reinterpret_cast<ulonglong2 *>(out + v * kGridClasses);
        o[0] = make_ulonglong2(m0, m1);
        o[1] = make_ulonglong2(m2, m3);
    }

    __device__ __forceinline__ unsigned long long grid_lookup(const GridDev &G, float x, float y, float z, int cls)
    {
        const int ix = __float2int_rd((x - G.x0) * G.inv_h);
        const int iy = __float2int_rd((y - G.y0) * G.inv_h);
        const int iz = __float2int_rd((z - G.z0) * G.inv_h);
        // outside the table = farther than the largest class radius from every object
        const bool in = (static_cast<unsigned>(ix) < static_cast<unsigned>(G.nx)) & (static_cast<unsigned>(iy) < static_cast<unsigned>(G.ny)) &
                        (static_cast<unsigned>(iz) < static_cast<unsigned>(G.nz));
        // a centre that is not finite fails `in` (its index saturates) -- give it every object
        if (!in)
        {
            return (x == x && y == y && z == z && fabsf(x) < 1e30F && fabsf(y) < 1e30F && fabsf(z) < 1e30F) ? 0ull : G.all_mask;
        }
        const size_t idx = ((static_cast<size_t>(iz) * G.ny + iy) * G.nx + ix) * kGridClasses + cls;
        return __ldg(G.masks + idx);
    }

    // One rigid-body frame read from the stash (9 floats; third rotation column = first x second).
    template <int BLOCK>
    struct BodyFrame
    {
        float r00, r01, r02, r10, r11, r12, r20, r21, r22, tx, ty, tz;

        __device__ __forceinline__ BodyFrame(int body, const float *stash)
        {
            if (body == 0)
            {
                r00 = 1.F, r01 = 0.F, r02 = 0.F, r10 = 0.F, r11 = 1.F, r12 = 0.F, r20 = 0.F, r21 = 0.F, r22 = 1.F;
                tx = 0.F, ty = 0.F, tz = 0.F;
            }
            else
            {
                const float *F = stash + (body - 1) * kFrameFloats * BLOCK;
                r00 = F[0 * BLOCK], r01 = F[1 * BLOCK], tx = F[2 * BLOCK];
                r10 = F[3 * BLOCK], r11 = F[4 * BLOCK], ty = F[5 * BLOCK];
                r20 = F[6 * BLOCK], r21 = F[7 * BLOCK], tz = F[8 * BLOCK];
                r02 = fmaf(r10, r21, -(r20 * r11));
                r12 = fmaf(r20, r01, -(r00 * r21));
                r22 = fmaf(r00, r11, -(r10 * r01));
            }
        }

        __device__ __forceinline__ void pose(float cx, float cy, float cz, float &x, float &y, float &z) const
        {
            x = fmaf(r00, cx, fmaf(r01, cy, fmaf(r02, cz, tx)));
            y = fmaf(r10, cx, fmaf(r11, cy, fmaf(r12, cz, ty)));
            z = fmaf(r20, cx, fmaf(r21, cy, fmaf(r22, cz, tz)));
        }
    };

    template <typename M, int BLOCK>
    struct SmemLayoutV3
    {
        static constexpr int kStashFrames = (M::kBodies - 1) * kFrameFloats;
        static constexpr int kPairCap = BLOCK * (M::kPairs > 0 ? M::kPairs : 1);
        static constexpr int kQ1Cap = BLOCK * M::kLinks;
        static_assert(BLOCK <= 128 && M::kPairs < 512 && M::kTasks < 512 && M::kLinks <= kGridMaxLinks, "16-bit work-item encoding");
        static constexpr int kGroups = BLOCK / 8;
        static constexpr int kScratchPerGroup = 32;
        static constexpr int kMaxPairLists = 256;

        uint32_t off_tasks, off_links, off_pairs, off_pinfo, off_plists, off_stash, off_masks, off_q1, off_q2, off_pairq, off_scratch,
            off_flags, q2_cap, total;

        __host__ __device__ static constexpr uint32_t align16(uint32_t v)
        {
            return (v + 15u) & ~15u;
        }

        __host__ __device__ SmemLayoutV3(uint32_t n_objects, uint32_t max_fine)
        {
            uint32_t o = align16(n_objects * kObjRec * 4);
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
            // B-phase (masks, Q1, Q2) and C-phase (pair records, scratch) regions alias
            off_masks = o;
            uint32_t b_bytes = M::kLinks * BLOCK * sizeof(unsigned long long);
            off_q1 = o + b_bytes;
            b_bytes += align16(kQ1Cap * sizeof(uint16_t));
            off_q2 = o + b_bytes;
            q2_cap = 2u * BLOCK * max_fine;  // two rounds of bounding items can always enqueue
            b_bytes += align16(q2_cap * sizeof(uint16_t));
            off_pairq = o;
            uint32_t c_bytes = align16(kPairCap * sizeof(uint16_t));
            off_scratch = o + c_bytes;
            o += b_bytes > c_bytes ? b_bytes : c_bytes;
            off_flags = o;
            o += 2 * align16(BLOCK * sizeof(uint32_t)) + 16;
            total = o;
        }
    };

    template <typename R, int BLOCK>
    struct V3Ctx
    {
        const float4 *objs;
        const PairInfo *pinfos;
        const SpherePair *plists;
        const SphereTask *tasks;
        const LinkInfo *links;
        const LinkPair *pairs;
        float *stash;
        unsigned long long *masks;
        uint16_t *q1, *q2;
        uint16_t *pairq;
        float4 *scratch;
        volatile uint32_t *invalid;
        volatile uint32_t *inbox;
        uint32_t *counters;  // [0] Q1, [1] Q2, [2] pair records
        uint64_t *bar;
        uint32_t max_fine, q2_cap;
    };

    template <typename R, int BLOCK>
    __device__ __forceinline__ V3Ctx<R, BLOCK> v3_stage(unsigned char *smem, uint64_t *barp, const RobotDev &robot, const LaunchEnvV3 &env)
    {
        using M = typename R::Model;
        using Lay = SmemLayoutV3<M, BLOCK>;
        const Lay L(env.n_objects, env.max_fine);
        const int tid = threadIdx.x;
        if (tid == 0)
        {
            mbar_init(barp, 1);
        }
        __syncthreads();
        if (tid == 0)
        {
            // n_objects >= 1 on this path (an empty environment has nothing to rasterise)
            tma_bulk_g2s(smem, env.objs, env.n_objects * kObjRec * 4, barp);
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
        V3Ctx<R, BLOCK> X;
        X.objs = reinterpret_cast<const float4 *>(smem);
        X.pinfos = reinterpret_cast<const PairInfo *>(smem + L.off_pinfo);
        X.plists = reinterpret_cast<const SpherePair *>(smem + L.off_plists);
        X.tasks = reinterpret_cast<const SphereTask *>(smem + L.off_tasks);
        X.links = reinterpret_cast<const LinkInfo *>(smem + L.off_links);
        X.pairs = reinterpret_cast<const LinkPair *>(smem + L.off_pairs);
        X.stash = reinterpret_cast<float *>(smem + L.off_stash);
        X.masks = reinterpret_cast<unsigned long long *>(smem + L.off_masks);
        X.q1 = reinterpret_cast<uint16_t *>(smem + L.off_q1);
        X.q2 = reinterpret_cast<uint16_t *>(smem + L.off_q2);
        X.pairq = reinterpret_cast<uint16_t *>(smem + L.off_pairq);
        X.scratch = reinterpret_cast<float4 *>(smem + L.off_scratch);
        X.invalid = reinterpret_cast<volatile uint32_t *>(smem + L.off_flags);
        X.inbox = X.invalid + BLOCK;
        X.counters = reinterpret_cast<uint32_t *>(smem + L.off_flags + 2 * Lay::align16(BLOCK * sizeof(uint32_t)));
        X.bar = barp;
        X.max_fine = env.max_fine;
        X.q2_cap = L.q2_cap;
        return X;
    }

    // fine spheres queued in Q2 against their link's hit mask
    template <typename R, int BLOCK>
    __device__ __forceinline__ void v3_flush_q2(const V3Ctx<R, BLOCK> &X)
    {
        const int tid = threadIdx.x;
        const uint32_t n2 = X.counters[1];
        for (uint32_t it = tid; it < n2; it += BLOCK)
        {
            const uint32_t item = X.q2[it];
            const int c = item & 127u;
            if (X.invalid[c])
            {
                continue;
            }
            const SphereTask t = X.tasks[item >> 7];
            float x, y, z;
            task_centre<BLOCK>(t, X.stash + c, x, y, z);
            unsigned long long m = X.masks[t.link * BLOCK + c];
            while (m != 0ull)
            {
                const int o = __ffsll(static_cast<long long>(m)) - 1;
                m &= m - 1ull;
                if (sign_set(margin_obj(X.objs + 4 * o, x, y, z, t.r)))
                {
                    X.invalid[c] = 1u;
                    break;
                }
            }
        }
        __syncthreads();
        if (tid == 0)
        {
            X.counters[1] = 0u;
        }
        __syncthreads();
    }

    template <typename R, int BLOCK>
    __device__ __forceinline__ void v3_pass(const V3Ctx<R, BLOCK> &X, const GridDev &G, const float (&cfg)[R::Model::kDof], const bool has)
    {
        using M = typename R::Model;
        using Lay = SmemLayoutV3<M, BLOCK>;
        const int tid = threadIdx.x;
        const PairInfo *pinfos = X.pinfos;
        const SpherePair *plists = X.plists;
        const SphereTask *tasks = X.tasks;
        const LinkInfo *links = X.links;
        const LinkPair *pairs = X.pairs;
        float *stash = X.stash;
        unsigned long long *masks = X.masks;
        uint16_t *pairq = X.pairq;
        volatile uint32_t *invalid = X.invalid;
        volatile uint32_t *inbox = X.inbox;
        uint32_t *counters = X.counters;

        // ---- A: FK ------------------------------------------------------------------------------
        StashBoundSink<BLOCK, M::kLinks> sink;
        sink.base = stash + tid;
        R::frames(cfg, sink);
        const bool live = has && !(sink.inbox && sink.self_hit);
        invalid[tid] = live ? 0u : 1u;
        inbox[tid] = sink.inbox ? 1u : 0u;
        if (tid < 3)
        {
            counters[tid] = 0;
        }

        // ---- B0: candidate masks from the voxel table ---------------------------------------------
        unsigned long long cand[M::kLinks];
        uint32_t link_bits_lo = 0u, link_bits_hi = 0u;
        if (live)
        {
            R::for_each_link(
                [&](auto l, float, int, int, float)
                {
                    constexpr int li = decltype(l)::value;
                    cand[li] = sink.reach_valid ? grid_lookup(G, sink.b[li][0], sink.b[li][1], sink.b[li][2], G.link_class[li]) : G.all_mask;
                });
            R::for_each_link(
                [&](auto l, float, int, int, float)
                {
                    constexpr int li = decltype(l)::value;
                    if (cand[li] != 0ull)
                    {
                        if (li < 32)
                        {
                            link_bits_lo |= 1u << (li & 31);
                        }
                        else
                        {
                            link_bits_hi |= 1u << (li & 31);
                        }
                    }
                });
        }
        __syncthreads();  // tables, flags, counters
        if ((link_bits_lo | link_bits_hi) != 0u)
        {
            uint32_t slot = atomicAdd(&counters[0], static_cast<uint32_t>(__popc(link_bits_lo) + __popc(link_bits_hi)));
            R::for_each_link(
                [&](auto l, float, int, int, float)
                {
                    constexpr int li = decltype(l)::value;
                    if (cand[li] != 0ull)
                    {
                        masks[li * BLOCK + tid] = cand[li];
                        X.q1[slot++] = static_cast<uint16_t>(tid | (li << 7));
                    }
                });
        }
        mbar_wait(X.bar, 0);
        __syncthreads();

        // ---- B1: bounding spheres vs their candidates; B2: fine spheres vs the hit masks ---------
        {
            const uint32_t n1 = counters[0];
            uint32_t q2_bound = 0u;  // upper bound of the Q2 fill, identical in every thread
            for (uint32_t base = 0; base < n1; base += BLOCK)
            {
                if (q2_bound + BLOCK * X.max_fine > X.q2_cap)
                {
                    v3_flush_q2<R, BLOCK>(X);
                    q2_bound = 0u;
                }
                q2_bound += BLOCK * X.max_fine;
                const uint32_t it = base + tid;
                if (it < n1)
                {
                    const uint32_t item = X.q1[it];
                    const int c = item & 127u, l = item >> 7;
                    if (!invalid[c])
                    {
                        const LinkInfo L = links[l];
                        const SphereTask t = tasks[L.bound_task];
                        float x, y, z;
                        task_centre<BLOCK>(t, stash + c, x, y, z);
                        unsigned long long m = masks[l * BLOCK + c], hit = 0ull;
                        while (m != 0ull)
                        {
                            const int o = __ffsll(static_cast<long long>(m)) - 1;
                            m &= m - 1ull;
                            if (sign_set(margin_obj(X.objs + 4 * o, x, y, z, t.r)))
                            {
                                hit |= 1ull << o;
                            }
                        }
                        if (hit != 0ull)
                        {
                            masks[l * BLOCK + c] = hit;
                            const uint32_t b2 = atomicAdd(&counters[1], static_cast<uint32_t>(L.n_spheres));
                            for (int k = 0; k < L.n_spheres; ++k)
                            {
                                X.q2[b2 + k] = static_cast<uint16_t>(c | ((L.bound_task + 1 + k) << 7));
                            }
                        }
                    }
                }
                __syncthreads();
            }
            v3_flush_q2<R, BLOCK>(X);  // ends with a barrier: the C-phase queues reuse this memory
        }

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
                        const uint32_t slot = atomicAdd(&counters[2], 1u);
                        pairq[slot] = static_cast<uint16_t>((tid << 9) | decltype(pi)::value);
                    }
                });
        }
        __syncthreads();

        // ---- C2: fine sphere pairs, one thread per (state, pair) record ---------------------------
        // The link with fewer spheres (A) is posed into registers four spheres at a time; the other
        // link's spheres (B) stream past, filtered by A's bounding sphere: a B sphere outside it
        // cannot touch any sphere of A.  Records of one pair sit next to each other in the queue
        // (C1 appends pair by pair), so the lanes of a warp mostly run the same trip counts.
        {
            const uint32_t n_rec = counters[2];
            for (uint32_t r = tid; r < n_rec; r += BLOCK)
            {
                const uint32_t rec = pairq[r];
                const int c = rec >> 9;
                if (invalid[c])
                {
                    continue;
                }
                const int pair = rec & 0x1ffu;
                const PairInfo pinfo = pinfos[pair];
                const float *st = stash + c;
                bool h = false;
                if (pinfo.count >= 0 && inbox[c] && pinfo.offset + pinfo.count <= Lay::kMaxPairLists)
                {
                    // statically pruned list of the sphere pairs that can touch inside the joint box
                    for (int k = 0; k < pinfo.count; ++k)
                    {
                        const SpherePair sp = plists[pinfo.offset + k];
                        const SphereTask ta = tasks[sp.task_a], tb = tasks[sp.task_b];
                        float ax, ay, az, bx, by, bz;
                        task_centre<BLOCK>(ta, st, ax, ay, az);
                        task_centre<BLOCK>(tb, st, bx, by, bz);
                        const float ex = ax - bx, ey = ay - by, ez = az - bz;
                        const float rr = ta.r + tb.r;
                        h |= sign_set((ex * ex + ey * ey + ez * ez) - rr * rr);
                    }
                }
                else
                {
                    const LinkPair p = pairs[pair];
                    LinkInfo A = links[p.a], B = links[p.b];
                    if (A.n_spheres > B.n_spheres)
                    {
                        const LinkInfo t = A;
                        A = B;
                        B = t;
                    }
                    const BodyFrame<BLOCK> FA(A.body, st), FB(B.body, st);
                    const SphereTask tba = tasks[A.bound_task];
                    float gx, gy, gz;
                    FA.pose(tba.cx, tba.cy, tba.cz, gx, gy, gz);
                    for (int a0 = 0; a0 < A.n_spheres && !h; a0 += 4)
                    {
                        float ax[4], ay[4], az[4], ar[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                        {
                            if (a0 + k < A.n_spheres)
                            {
                                const float4 ta = *reinterpret_cast<const float4 *>(&tasks[A.bound_task + 1 + a0 + k]);
                                FA.pose(ta.x, ta.y, ta.z, ax[k], ay[k], az[k]);
                                ar[k] = ta.w;
                            }
                            else
                            {
                                ax[k] = 1e18F, ay[k] = 1e18F, az[k] = 1e18F, ar[k] = 0.F;  // never touches
                            }
                        }
                        for (int jb = 0; jb < B.n_spheres; ++jb)
                        {
                            const float4 tb = *reinterpret_cast<const float4 *>(&tasks[B.bound_task + 1 + jb]);
                            float bx, by, bz;
                            FB.pose(tb.x, tb.y, tb.z, bx, by, bz);
                            const float fx = bx - gx, fy = by - gy, fz = bz - gz;
                            const float fr = tb.w + tba.r;
                            if (!sign_set((fx * fx + fy * fy + fz * fz) - fr * fr))
                            {
                                continue;
                            }
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                            {
                                const float ex = ax[k] - bx, ey = ay[k] - by, ez = az[k] - bz;
                                const float rr = ar[k] + tb.w;
                                h |= sign_set((ex * ex + ey * ey + ez * ez) - rr * rr);
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
        k_validate_configs_v3(RobotDev robot, const __grid_constant__ LaunchEnvV3 env, const float *__restrict__ q, size_t n, uint32_t *__restrict__ bits)
    {
        using M = typename R::Model;
        extern __shared__ __align__(128) unsigned char smem[];
        __shared__ uint64_t bar;
        const V3Ctx<R, BLOCK> X = v3_stage<R, BLOCK>(smem, &bar, robot, env);
        const int tid = threadIdx.x;
        const size_t i = static_cast<size_t>(blockIdx.x) * BLOCK + tid;
        const bool has = i < n;
        float cfg[M::kDof];
#pragma unroll
        for (int j = 0; j < M::kDof; ++j)
        {
            cfg[j] = has ? __ldg(q + i * M::kDof + j) : 0.F;
        }
        v3_pass<R, BLOCK>(X, env.grid, cfg, has);
        const bool valid = has && !X.invalid[tid];
        const uint32_t word = __ballot_sync(0xffffffffu, valid);
        if ((tid & 31) == 0 && has)
        {
            bits[i >> 5] = word;
        }
    }

    // Edges: same driver as k_validate_edges_v2 (32 edges per chunk, 16 rake blocks per pass,
    // round-robin over the live edges), passes through v3_pass.
    template <typename R, int BLOCK, bool INDEXED>
    __global__ void __launch_bounds__(BLOCK) k_validate_edges_v3(
        RobotDev robot,
        const __grid_constant__ LaunchEnvV3 env,
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
        const V3Ctx<R, BLOCK> X = v3_stage<R, BLOCK>(smem, &bar, robot, env);
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
                v3_pass<R, BLOCK>(X, env.grid, cfg, has);
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
