// Warp-autonomous grid-culled kernels (sm_100a) for environments made of primitives (at most 64
// objects).  Broad phase: the voxel table of candidate masks and the rounded-box records of
// vmv_grid.cuh; exact tests only on candidates.  The unit of cooperation is the warp, not the block:
//
//   * a warp owns 32 states per pass and a private slice of shared memory (frame stash, masks,
//     work queues); phases are separated by __syncwarp() only, so the 16-18 resident warps of an SM
//     drift apart and overlap each other's latencies instead of meeting at block barriers;
//   * queues are filled with ballot / prefix-sum positions (no shared-memory atomics); the verdict
//     mask of the 32 states is a warp-uniform register, merged with redux.sync after every round;
//   * blocks are persistent: the environment records (one TMA bulk copy) and the robot tables are
//     staged once per block, then each warp strides over its tiles;
//   * shared memory per warp is ~11-13 KB (Panda), which is what bounds residency.
//
// Phase order inside a pass: A (FK, lane = state) -> voxel-table loads issued -> C1 (link-pair bounding
// tests -> records) -> C2 (records, lane = record) -> B0 (consume the loads: candidate masks, Q1) ->
// B1 (lane = (state, link) item: exact bounding test, fine items -> Q2) -> B2 (lane = fine item).
// The self-collision phases sit between the issue and the use of the table loads and cover their
// L2 latency.
#pragma once
#include "vmv_grid.cuh"

namespace vmv
{
    static constexpr uint32_t kFullWarp = 0xffffffffu;

    // Multi-GPU verdict gather fused into the validation kernels (include/vamp_b200.h: vmv_comm_*).  Every rank
    // owns a window [world][words_per_rank] of verdict words; the windows of all ranks are mapped into every
    // process (CUDA IPC over NVLink / NVSwitch).  A kernel launched with world > 0 stores each verdict word it
    // produces into the launching rank's row of EVERY rank's window -- lane p of the warp stores to peer p, one
    // predicated 4-byte store instruction per 32 units -- so the all-gather costs no kernel of its own and no SM
    // is taken from the persistent validation blocks.  The last block to finish publishes `seq` in the flag
    // word (slot, rank) of every window, after a system-scope fence; consumers wait on their own flags.
    static constexpr int kMaxPeers = 8;
    struct GatherDev
    {
        int world;                      // 0: local launch, nothing below is read
        int rank;
        uint32_t seq;                   // value published when this launch's words have landed everywhere
        unsigned int *done;             // zeroed counter of finished blocks (local memory)
        uint32_t *peer_bits[kMaxPeers]; // this rank's row in peer p's window, for the slot being written
        uint32_t *peer_flag[kMaxPeers]; // flag word (slot, this rank) in peer p's window
    };

    // width = 4, 2 or 1 bytes (32, 16 or 8 units per verdict item)
    __device__ __forceinline__ void store_verdict(uint32_t *base, size_t index, uint32_t word, int width)
    {
        if (width == 4)
        {
            base[index] = word;
        }
        else if (width == 2)
        {
            reinterpret_cast<unsigned short *>(base)[index] = static_cast<unsigned short>(word);
        }
        else
        {
            reinterpret_cast<unsigned char *>(base)[index] = static_cast<unsigned char>(word);
        }
    }

    // called by every lane of a warp with the warp-uniform verdict item.  GATHER is a template parameter of the
    // kernels: carrying the peer table through a local launch costs registers (measured 3 % on the configuration
    // kernel), so local launches run an instantiation without it.
    template <bool GATHER>
    __device__ __forceinline__ void publish_verdict(uint32_t *bits, const GatherDev &g, size_t index, uint32_t word, int width)
    {
        const int lane = threadIdx.x & 31;
        if (!GATHER || g.world == 0)
        {
            if (lane == 0)
            {
                store_verdict(bits, index, word, width);
            }
        }
        else if (lane < g.world)
        {
            // lane == rank writes the local window row (peer_bits[rank] is the local mapping)
            store_verdict(g.peer_bits[lane], index, word, width);
        }
    }

    // end of a gathering kernel: called by every thread of the block after its last verdict store
    template <bool GATHER = true>
    __device__ __forceinline__ void gather_finish(const GatherDev &g)
    {
        if (!GATHER || g.world == 0)
        {
            return;
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            __threadfence_system();  // this block's remote stores are performed before the ticket below is visible
            const unsigned int t = atomicAdd(g.done, 1u);
            if (t == gridDim.x - 1)
            {
                __threadfence_system();
                for (int p = 0; p < g.world; ++p)
                {
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(g.peer_flag[p]), "r"(g.seq) : "memory");
                }
            }
        }
    }

    template <typename MaskT>
    __device__ __forceinline__ int mask_pop_lowest(MaskT &m);

    template <>
    __device__ __forceinline__ int mask_pop_lowest<uint32_t>(uint32_t &m)
    {
        const int o = __ffs(static_cast<int>(m)) - 1;
        m &= m - 1u;
        return o;
    }

    template <>
    __device__ __forceinline__ int mask_pop_lowest<unsigned long long>(unsigned long long &m)
    {
        const int o = __ffsll(static_cast<long long>(m)) - 1;
        m &= m - 1ull;
        return o;
    }

    // Append one item per set bit of every lane's `bits` to a warp-private queue: item = lane |
    // (base + bit) << 5, positions from an exclusive prefix sum over the lanes.  Returns the new fill.
    __device__ __forceinline__ uint32_t warp_append_bits(uint16_t *queue, uint32_t fill, uint32_t bits, uint32_t base)
    {
        const int lane = threadIdx.x & 31;
        const uint32_t cnt = static_cast<uint32_t>(__popc(bits));
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t v = __shfl_up_sync(kFullWarp, incl, d);
            incl += (lane >= d) ? v : 0u;
        }
        uint32_t at = fill + incl - cnt;
        while (bits != 0u)
        {
            const uint32_t b = static_cast<uint32_t>(__ffs(static_cast<int>(bits)) - 1);
            bits &= bits - 1u;
            queue[at++] = static_cast<uint16_t>(lane | ((base + b) << 5));
        }
        return fill + __shfl_sync(kFullWarp, incl, 31);
    }

    // Verdict table of one pair group (vmv_pairtab.cuh): thread = cell.  The other joints stay at 0 --
    // the relative pose of the group's links does not depend on them.
    template <typename R>
    __global__ void __launch_bounds__(128) k_build_pair_table(
        RobotDev robot,
        int dof_a,
        int dof_b,
        float lo_a,
        float step_a,
        float lo_b,
        float step_b,
        int na,
        int nb,
        float band,
        const int *__restrict__ group_pairs,
        int n_pairs,
        unsigned char *__restrict__ cells)
    {
        using M = typename R::Model;
        const size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
        if (v >= static_cast<size_t>(na) * nb)
        {
            return;
        }
        const int ia = static_cast<int>(v / nb), ib = static_cast<int>(v % nb);
        float cfg[M::kDof];
#pragma unroll
        for (int j = 0; j < M::kDof; ++j)
        {
            cfg[j] = (j == dof_a) ? lo_a + (ia + 0.5F) * step_a : ((j == dof_b) ? lo_b + (ib + 0.5F) * step_b : 0.F);
        }
        float F[(M::kBodies - 1) * 12];
        RegSink sink{F};
        R::frames(cfg, sink);
        auto pose = [&](const SphereTask &t, float &x, float &y, float &z)
        {
            x = t.cx, y = t.cy, z = t.cz;
            if (t.body > 0)
            {
                const float *f = F + (t.body - 1) * 12;
                x = fmaf(f[0], t.cx, fmaf(f[1], t.cy, fmaf(f[2], t.cz, f[3])));
                y = fmaf(f[4], t.cx, fmaf(f[5], t.cy, fmaf(f[6], t.cz, f[7])));
                z = fmaf(f[8], t.cx, fmaf(f[9], t.cy, fmaf(f[10], t.cz, f[11])));
            }
        };
        float cmin = 3.0e38F;
        for (int p = 0; p < n_pairs; ++p)
        {
            const PairInfo pi = robot.pair_info[group_pairs[p]];
            for (int k = 0; k < pi.count; ++k)
            {
                const SpherePair sp = robot.pair_lists[pi.offset + k];
                const SphereTask ta = robot.tasks[sp.task_a], tb = robot.tasks[sp.task_b];
                float ax, ay, az, bx, by, bz;
                pose(ta, ax, ay, az);
                pose(tb, bx, by, bz);
                const float ex = ax - bx, ey = ay - by, ez = az - bz;
                cmin = fminf(cmin, sqrtf(ex * ex + ey * ey + ez * ez) - (ta.r + tb.r));
            }
        }
        cells[v] = cmin > band ? 0 : (cmin < -band ? 1 : 2);
    }

    template <typename M, typename MaskT>
    struct SmemLayoutV4
    {
        static constexpr int kStashFrames = (M::kBodies - 1) * kFrameFloats;
        static constexpr int kPairChunk = 64;  // link pairs per C1 round (bounds the record queue)
        static constexpr int kPairCap = 32 * (M::kPairs < 1 ? 1 : (M::kPairs < kPairChunk ? M::kPairs : kPairChunk));
        static constexpr int kMaxPairLists = 256;
        static_assert(M::kPairs < 2048 && M::kTasks < 2048 && M::kLinks <= kGridMaxLinks, "16-bit work-item encoding");

        // block-shared part, then one private slice per warp
        uint32_t off_tasks, off_links, off_pairs, off_pinfo, off_plists, off_blob, shared_bytes;
        uint32_t w_stash, w_masks, w_q1, w_q2, w_pairq, warp_bytes, q2_cap;

        __host__ __device__ static constexpr uint32_t align16(uint32_t v)
        {
            return (v + 15u) & ~15u;
        }

        __host__ __device__ SmemLayoutV4(uint32_t n_objects, uint32_t max_fine, uint32_t q2_rounds, uint32_t blob_bytes = 0)
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
            off_blob = o;
            o += align16(blob_bytes);
            shared_bytes = (o + 127u) & ~127u;

            uint32_t w = 0;
            w_stash = w;
            w += kStashFrames * 32 * sizeof(float);
            // the record queue of C1/C2 is dead before B0 writes masks and Q1: same memory
            w_masks = w;
            uint32_t b = M::kLinks * 32 * sizeof(MaskT);
            w_q1 = w + b;
            b += align16(M::kLinks * 32 * sizeof(uint16_t));
            w_q2 = w + b;
            // one B1 round always fits; two (the common case) avoid a flush in between
            q2_cap = (q2_rounds < 1u ? 1u : q2_rounds) * 32u * max_fine;
            b += align16(q2_cap * sizeof(uint16_t));
            w_pairq = w;
            const uint32_t c = align16(kPairCap * sizeof(uint16_t));
            w += b > c ? b : c;
            warp_bytes = (w + 63u) & ~63u;  // 64 B: Panda lands on 11328 B, which lets a 20th warp fit
        }

        __host__ __device__ uint32_t total(uint32_t warps) const
        {
            return shared_bytes + warps * warp_bytes;
        }
    };

    template <typename R, typename MaskT>
    struct V4Ctx
    {
        const float4 *objs;
        const PairInfo *pinfos;
        const SpherePair *plists;
        const SphereTask *tasks;
        const LinkInfo *links;
        const LinkPair *pairs;
        float *stash;  // this warp's [entry][32]
        MaskT *masks;  // this warp's [link][32]
        uint16_t *q1, *q2, *pairq;
        uint32_t q2_cap, round_cap;  // Q2 capacity; most fine items one B1 round can add (32 * max_fine)
        const float *E;              // any-environment instantiation: the packed environment blob in shared memory
        const int *attach_links;     // the links an attachment is checked against (global memory, kAttachLinks entries)
        uint32_t n_objects;
    };

    // Block-level staging (once per persistent block) and this warp's slice.
    template <typename R, typename MaskT>
    __device__ __forceinline__ V4Ctx<R, MaskT> v4_stage(unsigned char *smem, uint64_t *barp, const RobotDev &robot, const GridEnv &env)
    {
        using M = typename R::Model;
        using Lay = SmemLayoutV4<M, MaskT>;
        const Lay L(env.n_objects, env.max_fine, env.q2_rounds, env.blob_bytes);
        const int tid = threadIdx.x, nthr = blockDim.x;
        if (tid == 0)
        {
            mbar_init(barp, 1);
        }
        __syncthreads();
        if (tid == 0 && env.n_objects > 0)
        {
            tma_bulk_g2s(smem, env.objs, env.n_objects * kObjRec * 4, barp);
        }
        {
            uint32_t *dst = reinterpret_cast<uint32_t *>(smem + L.off_tasks);
            const uint32_t *src = reinterpret_cast<const uint32_t *>(robot.tasks);
            for (int i = tid; i < M::kTasks * 8; i += nthr)
            {
                dst[i] = __ldg(src + i);
            }
            dst = reinterpret_cast<uint32_t *>(smem + L.off_links);
            src = reinterpret_cast<const uint32_t *>(robot.links);
            for (int i = tid; i < M::kLinks * 4; i += nthr)
            {
                dst[i] = __ldg(src + i);
            }
            dst = reinterpret_cast<uint32_t *>(smem + L.off_pairs);
            src = reinterpret_cast<const uint32_t *>(robot.pairs);
            for (int i = tid; i < M::kPairs * 2; i += nthr)
            {
                dst[i] = __ldg(src + i);
            }
            dst = reinterpret_cast<uint32_t *>(smem + L.off_pinfo);
            src = reinterpret_cast<const uint32_t *>(robot.pair_info);
            for (int i = tid; i < M::kPairs * 3; i += nthr)
            {
                dst[i] = __ldg(src + i);
            }
            uint16_t *dl = reinterpret_cast<uint16_t *>(smem + L.off_plists);
            const uint16_t *sl = reinterpret_cast<const uint16_t *>(robot.pair_lists);
            for (int i = tid; i < min(robot.n_pair_lists, Lay::kMaxPairLists); i += nthr)
            {
                dl[i] = __ldg(sl + i);
            }
            dst = reinterpret_cast<uint32_t *>(smem + L.off_blob);
            src = reinterpret_cast<const uint32_t *>(env.blob);
            for (int i = tid; i < static_cast<int>(env.blob_bytes / 4); i += nthr)
            {
                dst[i] = __ldg(src + i);
            }
        }
        __syncthreads();
        if (env.n_objects > 0)
        {
            mbar_wait(barp, 0);
        }

        unsigned char *mine = smem + L.shared_bytes + (tid >> 5) * L.warp_bytes;
        V4Ctx<R, MaskT> X;
        X.objs = reinterpret_cast<const float4 *>(smem);
        X.pinfos = reinterpret_cast<const PairInfo *>(smem + L.off_pinfo);
        X.plists = reinterpret_cast<const SpherePair *>(smem + L.off_plists);
        X.tasks = reinterpret_cast<const SphereTask *>(smem + L.off_tasks);
        X.links = reinterpret_cast<const LinkInfo *>(smem + L.off_links);
        X.pairs = reinterpret_cast<const LinkPair *>(smem + L.off_pairs);
        X.stash = reinterpret_cast<float *>(mine + L.w_stash);
        X.masks = reinterpret_cast<MaskT *>(mine + L.w_masks);
        X.q1 = reinterpret_cast<uint16_t *>(mine + L.w_q1);
        X.q2 = reinterpret_cast<uint16_t *>(mine + L.w_q2);
        X.pairq = reinterpret_cast<uint16_t *>(mine + L.w_pairq);
        X.q2_cap = L.q2_cap;
        X.round_cap = 32u * env.max_fine;
        X.E = reinterpret_cast<const float *>(smem + L.off_blob);
        X.attach_links = robot.attach_links;
        X.n_objects = env.n_objects;
        return X;
    }

    // C2: one lane per (state, pair) record.  Returns the states found in self collision.
    template <typename R, typename MaskT, bool SPLIT_ALWAYS>
    __device__ __forceinline__ uint32_t
    v4_pair_records(const V4Ctx<R, MaskT> &X, uint32_t n_rec, uint32_t invalid, uint32_t inbox_mask)
    {
        using Lay = SmemLayoutV4<typename R::Model, MaskT>;
        const int lane = threadIdx.x & 31;
        uint32_t hits = 0u;
        // A round lasts as long as its longest record (the hand's 18 spheres): lanes share records, each taking every
        // step-th sphere of the longer link (or entry of the pruned list).  Configuration kernel: two lanes per
        // record (measured +6 % over one; 4 and 8 lanes for rounds of few records: -0.7 % -- the kernel is bound
        // by the latency of its dependent steps, not by the instructions of this loop).  Edge kernel: its tiles are
        // bursty (the 8 tines of a rake block raise the same records), so the lanes are dealt out by the number of
        // records -- 8 per record up to 4, 4 up to 8, 2 up to 16, one above (measured +1.3 %).
        const int sh = SPLIT_ALWAYS ? 1 : (n_rec <= 4u ? 3 : (n_rec <= 8u ? 2 : (n_rec <= 16u ? 1 : 0)));
        const int half = lane & ((1 << sh) - 1), step = 1 << sh;
        const uint32_t first = lane >> sh, stride = 32u >> sh;
        for (uint32_t r = first; r < n_rec; r += stride)
        {
            const uint32_t rec = X.pairq[r];
            const int c = rec & 31u;
            if ((invalid >> c) & 1u)
            {
                continue;
            }
            const int pair = rec >> 5;
            const PairInfo pinfo = X.pinfos[pair];
            const float *st = X.stash + c;
            bool h = false;
            if (pinfo.count >= 0 && ((inbox_mask >> c) & 1u) && pinfo.offset + pinfo.count <= Lay::kMaxPairLists)
            {
                // statically pruned list of the sphere pairs that can touch inside the joint box
                for (int k = half; k < pinfo.count; k += step)
                {
                    const SpherePair sp = X.plists[pinfo.offset + k];
                    const SphereTask ta = X.tasks[sp.task_a], tb = X.tasks[sp.task_b];
                    float ax, ay, az, bx, by, bz;
                    task_centre<32>(ta, st, ax, ay, az);
                    task_centre<32>(tb, st, bx, by, bz);
                    const float ex = ax - bx, ey = ay - by, ez = az - bz;
                    const float rr = ta.r + tb.r;
                    h |= sign_set((ex * ex + ey * ey + ez * ez) - rr * rr);
                }
            }
            else
            {
                // the link with fewer spheres (A) is posed into registers four spheres at a time; the
                // other link's spheres stream past, filtered by A's bounding sphere
                const LinkPair p = X.pairs[pair];
                LinkInfo A = X.links[p.a], B = X.links[p.b];
                if (A.n_spheres > B.n_spheres)
                {
                    const LinkInfo t = A;
                    A = B;
                    B = t;
                }
                const BodyFrame<32> FA(A.body, st), FB(B.body, st);
                const float4 tba = *reinterpret_cast<const float4 *>(&X.tasks[A.bound_task]);
                float gx, gy, gz;
                FA.pose(tba.x, tba.y, tba.z, gx, gy, gz);
                for (int a0 = 0; a0 < A.n_spheres && !h; a0 += 4)
                {
                    float ax[4], ay[4], az[4], ar[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                    {
                        if (a0 + k < A.n_spheres)
                        {
                            const float4 ta = *reinterpret_cast<const float4 *>(&X.tasks[A.bound_task + 1 + a0 + k]);
                            FA.pose(ta.x, ta.y, ta.z, ax[k], ay[k], az[k]);
                            ar[k] = ta.w;
                        }
                        else
                        {
                            ax[k] = 1e18F, ay[k] = 1e18F, az[k] = 1e18F, ar[k] = 0.F;  // never touches
                        }
                    }
                    for (int jb = half; jb < B.n_spheres; jb += step)
                    {
                        const float4 tb = *reinterpret_cast<const float4 *>(&X.tasks[B.bound_task + 1 + jb]);
                        float bx, by, bz;
                        FB.pose(tb.x, tb.y, tb.z, bx, by, bz);
                        const float fx = bx - gx, fy = by - gy, fz = bz - gz;
                        const float fr = tb.w + tba.w;
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
            hits |= h ? (1u << c) : 0u;
        }
        return __reduce_or_sync(kFullWarp, hits);
    }

    // Any-environment instantiation: what a 32-bit candidate mask carries beside its (at most 30) object bits
    static constexpr uint32_t kAeObjMask = 0x3fffffffu;  // primitive candidates / exact hits
    static constexpr uint32_t kAeHfBit = 0x40000000u;    // the link's bounding sphere hits a heightfield (exact test, done with the FK)
    static constexpr uint32_t kAeCloudBit = 0x80000000u; // the nearest-point table could not rule the pointclouds out for the bounding sphere

    // B2: one lane per fine-sphere item, against its link's hit mask -- and, in the any-environment instantiation,
    // against the heightfields and the pointclouds (a link whose bounding sphere hit ANYTHING has its fine spheres
    // swept against EVERYTHING, reference robots/panda.hh:5633-5645; the pointcloud query is warp-cooperative, so the
    // whole round calls it together).  Returns the colliding states.
    template <typename R, typename MaskT, bool AE>
    __device__ __forceinline__ uint32_t v4_fine_items(const V4Ctx<R, MaskT> &X, uint32_t n2, uint32_t invalid)
    {
        const int lane = threadIdx.x & 31;
        bool has_hf = false, has_cloud = false;
        if constexpr (AE)
        {
            const EnvHeader &H = *reinterpret_cast<const EnvHeader *>(X.E);
            has_hf = H.n_heightfields > 0, has_cloud = H.n_capts + H.n_mvts > 0;
        }
        for (uint32_t base = 0; base < n2; base += 32)
        {
            const uint32_t it = base + lane;
            uint32_t hit = 0u;
            float x = 0.F, y = 0.F, z = 0.F, r = 0.F;
            bool ask_cloud = false;
            int c = 0;
            if (it < n2)
            {
                const uint32_t item = X.q2[it];
                c = item & 31u;
                if (!((invalid >> c) & 1u))
                {
                    const SphereTask t = X.tasks[item >> 5];
                    task_centre<32>(t, X.stash + c, x, y, z);
                    r = t.r;
                    MaskT m = X.masks[t.link * 32 + c];
                    while (m != 0)
                    {
                        const int o = mask_pop_lowest<MaskT>(m);
                        if (sign_set(margin_obj(X.objs + 4 * o, x, y, z, t.r)))
                        {
                            hit = 1u << c;
                            break;
                        }
                    }
                    if constexpr (AE)
                    {
                        if (hit == 0u && has_hf && sphere_hits_heightfields(X.E, x, y, z, r))
                        {
                            hit = 1u << c;
                        }
                        ask_cloud = hit == 0u;
                    }
                }
            }
            if constexpr (AE)
            {
                // (asked in two sub-rounds -- the first open question of every state, then the rest for the states that came
                // through -- the fine spheres raise a quarter of the pointcloud questions, but every sub-round is a full
                // descent-and-scan latency: measured Fetch 2.14e8 -> 1.89e8 configurations/s, UR5 4.33e8 -> 4.21e8; not kept)
                if (has_cloud && sphere_hits_clouds(X.E, x, y, z, r, ask_cloud))
                {
                    hit = 1u << c;
                }
            }
            invalid |= __reduce_or_sync(kFullWarp, hit);
        }
        return invalid;
    }

    // T = F[ee_body] * tf   (Attachment::pose, collision/attachments.hh:43-55); `stash` already offset by the lane
    template <typename M>
    __device__ __forceinline__ void attachment_frame(const float *stash, const AttachDev &A, float (&T)[12])
    {
        float F[12];
        if (M::kEeBody == 0)
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
                F[k] = (k % 4 == 2) ? 0.F : stash[((M::kEeBody - 1) * kFrameFloats + frame_slot(k)) * 32];
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
                const float s = F[4 * i] * A.tf[j] + F[4 * i + 1] * A.tf[4 + j] + F[4 * i + 2] * A.tf[8 + j];
                T[4 * i + j] = (j == 3) ? s + F[4 * i + 3] : s;
            }
        }
    }

    // D: the attachment (reference fkcc_attach, robots/panda.hh:15308-15440; validity.hh:259-301), lane = state: its spheres,
    // posed by the end-effector frame, against every primitive of the environment and against the links of attach_links
    // (bounding sphere first).  Primitive environments only (the any-environment batches with an attachment stay on the
    // per-thread kernel).  Returns true for a state in collision.
    // (Not inlined, and handed plain pointers: batches without an attachment must not carry this code through their passes --
    // inlined it cost the edge kernel 5 %, which is bound by instruction fetch)
    template <typename M>
    static __device__ __noinline__ bool v4_attachment(const float *stash, const float4 *objs, uint32_t n_objects, const SphereTask *tasks,
                                                      const LinkInfo *links, const int *attach_links, const AttachDev *Ap, bool alive)
    {
        const AttachDev &A = *Ap;
        struct
        {
            const float *stash;
            const float4 *objs;
            uint32_t n_objects;
            const SphereTask *tasks;
            const LinkInfo *links;
            const int *attach_links;
        } X{stash, objs, n_objects, tasks, links, attach_links};
        const int lane = 0;  // `stash` is already offset by the lane
        if (!alive)
        {
            return false;
        }
        float T[12];
        attachment_frame<M>(X.stash + lane, A, T);
        auto posed = [&](uint32_t i, float &x, float &y, float &z) -> float
        {
            const float4 s = __ldg(A.spheres + i);
            x = fmaf(T[0], s.x, fmaf(T[1], s.y, fmaf(T[2], s.z, T[3])));
            y = fmaf(T[4], s.x, fmaf(T[5], s.y, fmaf(T[6], s.z, T[7])));
            z = fmaf(T[8], s.x, fmaf(T[9], s.y, fmaf(T[10], s.z, T[11])));
            return s.w;
        };
        // attachment vs environment (validity.hh:259-276)
        for (uint32_t i = 0; i < A.n; ++i)
        {
            float x, y, z;
            const float r = posed(i, x, y, z);
            for (uint32_t o = 0; o < X.n_objects; ++o)
            {
                if (sign_set(margin_obj(X.objs + 4 * o, x, y, z, r)))
                {
                    return true;
                }
            }
        }
        // attachment vs links, bounding sphere first (validity.hh:278-301)
        for (int k = 0; k < M::kAttachLinks; ++k)
        {
            const LinkInfo li = X.links[__ldg(X.attach_links + k)];
            const SphereTask tb = X.tasks[li.bound_task];
            float bx, by, bz;
            task_centre<32>(tb, X.stash + lane, bx, by, bz);
            bool near = false;
            for (uint32_t i = 0; i < A.n && !near; ++i)
            {
                float x, y, z;
                const float r = posed(i, x, y, z);
                const float dx = bx - x, dy = by - y, dz = bz - z;
                const float rs = tb.r + r;
                near = sign_set((dx * dx + dy * dy + dz * dz) - rs * rs);
            }
            if (!near)
            {
                continue;
            }
            for (int f = 0; f < li.n_spheres; ++f)
            {
                const SphereTask tf = X.tasks[li.bound_task + 1 + f];
                float fx, fy, fz;
                task_centre<32>(tf, X.stash + lane, fx, fy, fz);
                for (uint32_t i = 0; i < A.n; ++i)
                {
                    float x, y, z;
                    const float r = posed(i, x, y, z);
                    const float dx = fx - x, dy = fy - y, dz = fz - z;
                    const float rs = tf.r + r;
                    if (sign_set((dx * dx + dy * dy + dz * dz) - rs * rs))
                    {
                        return true;
                    }
                }
            }
        }
        return false;
    }

    // Any environment: the attachment's spheres against everything (validity.hh:259-276) -- primitives and heightfields per lane,
    // the pointcloud query warp-cooperative like everywhere else (EVERY lane of the warp calls; `alive` false = the lane only
    // helps) -- then the links (the routine above, told there are no primitives).  Out of line for the same reason.
    template <typename M>
    static __device__ __noinline__ bool v4_attachment_any(const float *stash, const float4 *objs, uint32_t n_objects, const float *E, const SphereTask *tasks,
                                                          const LinkInfo *links, const int *attach_links, const AttachDev *Ap, bool alive)
    {
        const AttachDev &A = *Ap;
        const EnvHeader &H = *reinterpret_cast<const EnvHeader *>(E);
        float T[12];
        attachment_frame<M>(stash, A, T);
        bool bad = false;
        for (uint32_t i = 0; i < A.n; ++i)
        {
            const float4 sp = __ldg(A.spheres + i);
            const float x = fmaf(T[0], sp.x, fmaf(T[1], sp.y, fmaf(T[2], sp.z, T[3])));
            const float y = fmaf(T[4], sp.x, fmaf(T[5], sp.y, fmaf(T[6], sp.z, T[7])));
            const float z = fmaf(T[8], sp.x, fmaf(T[9], sp.y, fmaf(T[10], sp.z, T[11])));
            bool h = false;
            if (alive && !bad)
            {
                for (uint32_t o = 0; o < n_objects && !h; ++o)
                {
                    h = sign_set(margin_obj(objs + 4 * o, x, y, z, sp.w));
                }
                h = h || (H.n_heightfields > 0 && sphere_hits_heightfields(E, x, y, z, sp.w));
            }
            if (H.n_capts + H.n_mvts > 0)
            {
                h = sphere_hits_clouds(E, x, y, z, sp.w, alive && !bad && !h) || h;
            }
            bad = bad || h;
        }
        return bad || v4_attachment<M>(stash, objs, 0u, tasks, links, attach_links, Ap, alive && !bad);
    }

    template <int N, typename F>
    __device__ __forceinline__ void static_for(F &&f)
    {
        if constexpr (N > 0)
        {
            static_for<N - 1>(f);
            f(std::integral_constant<int, N - 1>{});
        }
    }

    // One pass over 32 states (lane = state in the dense phases).  Returns the warp-uniform mask of
    // invalid states (lanes without a state count as invalid).
    template <typename R, typename MaskT, bool TAB, bool SPLIT_ALWAYS, bool AE = false, bool ATT = false>
    __device__ __forceinline__ uint32_t
    v4_pass(const V4Ctx<R, MaskT> &X, const GridDev &G, const PairTabDev &T, const AttachDev &A, const float (&cfg)[R::Model::kDof], const bool has)
    {
        using M = typename R::Model;
        using Lay = SmemLayoutV4<M, MaskT>;
        const int lane = threadIdx.x & 31;
        constexpr int kWords = (M::kPairs + 31) / 32 > 0 ? (M::kPairs + 31) / 32 : 1;
        static_assert(kWords <= kPairTabWords, "pair masks of the verdict tables");

        // ---- two-joint verdict tables: one byte per group, loaded ahead of the FK ------------------
        // (the groups and their joints are compile-time facts of the robot: gen/<robot>_pairtab.h)
        constexpr int kGroups = TAB ? (R::PairTab::kGroups < kPairTabMaxGroups ? R::PairTab::kGroups : kPairTabMaxGroups) : 0;
        unsigned char cell[kPairTabMaxGroups] = {2, 2, 2, 2};
        static_for<kGroups>(
            [&](auto gc)
            {
                constexpr int g = decltype(gc)::value;
                constexpr int da = R::PairTab::dof_a(g), db = R::PairTab::dof_b(g);
                const float qa = cfg[da >= 0 ? da : 0];
                const float qb = db >= 0 ? cfg[db >= 0 ? db : 0] : 0.F;
                // clamped: the byte is only used for states inside the joint box
                const int ia = min(max(__float2int_rd((qa - T.lo_a[g]) * T.inv_a[g]), 0), T.na[g] - 1);
                const int ib = min(max(__float2int_rd((qb - T.lo_b[g]) * T.inv_b[g]), 0), T.nb[g] - 1);
                cell[g] = __ldg(T.cells[g] + static_cast<size_t>(ia) * T.nb[g] + ib);
            });

        // ---- A: FK ------------------------------------------------------------------------------
        StashBoundSink<32, M::kLinks, !TAB> sink;
        sink.base = X.stash + lane;
        R::frames(cfg, sink);
        // link pairs already decided for this state (inside the joint box): by a table, or because
        // their pruned list is empty
        uint32_t decided[kWords];
#pragma unroll
        for (int w = 0; w < kWords; ++w)
        {
            decided[w] = 0u;
        }
        if (TAB && sink.inbox)
        {
#pragma unroll
            for (int w = 0; w < kWords; ++w)
            {
                if (R::PairTab::word(w) != 0u)
                {
                    decided[w] = T.never_pairs[w];
                }
            }
#pragma unroll
            for (int g = 0; g < kPairTabMaxGroups; ++g)
            {
                if (g < kGroups)
                {
                    sink.self_hit = sink.self_hit || cell[g] == 1;
                    if (cell[g] != 2)
                    {
#pragma unroll
                        for (int w = 0; w < kWords; ++w)
                        {
                            if (R::PairTab::word(w) != 0u)
                            {
                                decided[w] |= T.group_pairs[g][w];
                            }
                        }
                    }
                }
            }
        }
        const bool live = has && !(sink.inbox && sink.self_hit);
        uint32_t invalid = ~__ballot_sync(kFullWarp, live);
        const uint32_t inbox_mask = __ballot_sync(kFullWarp, sink.inbox);

        // ---- voxel-table loads (consumed in B0, after the self-collision phases) -----------------
        MaskT cand[M::kLinks];
        R::for_each_link(
            [&](auto l, float, int, int, float)
            {
                constexpr int li = decltype(l)::value;
                cand[li] = !live ? static_cast<MaskT>(0)
                                 : (sink.reach_valid ? grid_lookup_t<MaskT>(G, sink.b[li][0], sink.b[li][1], sink.b[li][2], G.link_class[li])
                                                     : static_cast<MaskT>(G.all_mask));
            });
        __syncwarp();  // stash complete
        // ---- C1: allowed link pairs on the bounding spheres -> records; C2 per chunk of pairs -----
        if (M::kPairs > 0)
        {
            float brad[M::kLinks];
            R::for_each_link([&](auto l, float br, int, int, float) { brad[decltype(l)::value] = br; });
            uint32_t n_rec = 0u, word = 0u;
            R::for_each_pair(
                [&](auto pi, auto la, auto lb, auto inl)
                {
                    constexpr int a = decltype(la)::value, b = decltype(lb)::value, p = decltype(pi)::value;
                    const float dx = sink.b[a][0] - sink.b[b][0], dy = sink.b[a][1] - sink.b[b][1], dz = sink.b[a][2] - sink.b[b][2];
                    const float rs = brad[a] + brad[b];
                    // skipped: pairs checked inline in phase A (inside the joint box), pairs a table decided
                    constexpr bool tabled = ((R::PairTab::word(p >> 5) >> (p & 31)) & 1u) != 0u;
                    const bool skip = TAB ? (tabled && ((decided[p >> 5] >> (p & 31)) & 1u) != 0u) : (decltype(inl)::value != 0 && sink.inbox);
                    const bool hit = !skip && sign_set((dx * dx + dy * dy + dz * dz) - rs * rs);
                    word |= hit ? (1u << (p & 31)) : 0u;
                    if constexpr ((p & 31) == 31 || p + 1 == M::kPairs)
                    {
                        n_rec = warp_append_bits(X.pairq, n_rec, live ? word : 0u, p & ~31);
                        word = 0u;
                        if constexpr ((p + 1) % Lay::kPairChunk == 0 && p + 1 < M::kPairs)
                        {
                            __syncwarp();
                            invalid |= v4_pair_records<R, MaskT, SPLIT_ALWAYS>(X, n_rec, invalid, inbox_mask);
                            __syncwarp();
                            n_rec = 0u;
                        }
                    }
                });
            __syncwarp();
            invalid |= v4_pair_records<R, MaskT, SPLIT_ALWAYS>(X, n_rec, invalid, inbox_mask);
            __syncwarp();  // the record queue's memory is reused below
        }

        // ---- any environment: the bounding spheres against the heightfields (exact, one gather each) and against the
        //      nearest-point table of the pointclouds (one load each), lane = state.  A ROLLED loop over the links, each
        //      centre re-derived from the frame stash the way B1 does it (the kernel is bound by instruction fetch: unrolled
        //      over the centres in registers this phase was 7 KB for Fetch); the two flags wait in the link's mask slot
        //      (which shares its memory with the record queue of C1/C2: this phase runs after them, and only for the states
        //      the self-collision check has left).
        if constexpr (AE)
        {
            static_assert(sizeof(MaskT) == 4, "the any-environment instantiation keeps its link flags in 32-bit masks");
            const EnvHeader &H = *reinterpret_cast<const EnvHeader *>(X.E);
            const bool has_hf = H.n_heightfields > 0, has_cloud = H.n_capts + H.n_mvts > 0;
#pragma unroll 1
            for (int li = 0; li < M::kLinks; ++li)
            {
                uint32_t flags = 0u;
                if (live && !((invalid >> lane) & 1u))
                {
                    const SphereTask t = X.tasks[X.links[li].bound_task];
                    float x, y, z;
                    task_centre<32>(t, X.stash + lane, x, y, z);
                    if (has_hf && sphere_hits_heightfields(X.E, x, y, z, t.r))
                    {
                        flags |= kAeHfBit;
                    }
                    if (has_cloud)
                    {
                        bool undecided = true;
                        if (H.off_cloud_grid != 0)
                        {
                            const CloudGridRec &g = *reinterpret_cast<const CloudGridRec *>(X.E + H.off_cloud_grid);
                            undecided = !(cloud_clearance(g, x, y, z) > (t.r - 1e-6F) + g.r_point_max);
                        }
                        flags |= undecided ? kAeCloudBit : 0u;
                    }
                }
                X.masks[li * 32 + lane] = static_cast<MaskT>(flags);
            }
        }

        // ---- B0: candidate masks -> (state, link) items -------------------------------------------
        uint32_t n1 = 0u;
        {
            const bool alive = !((invalid >> lane) & 1u);
            uint32_t word = 0u;
            R::for_each_link(
                [&](auto l, float, int, int, float)
                {
                    constexpr int li = decltype(l)::value;
                    MaskT cm = cand[li];
                    if constexpr (AE)
                    {
                        cm = static_cast<MaskT>((static_cast<uint32_t>(cm) & kAeObjMask) | static_cast<uint32_t>(X.masks[li * 32 + lane]));
                    }
                    if (cm != 0)
                    {
                        X.masks[li * 32 + lane] = cm;
                        word |= 1u << (li & 31);
                    }
                    if constexpr ((li & 31) == 31 || li + 1 == M::kLinks)
                    {
                        n1 = warp_append_bits(X.q1, n1, alive ? word : 0u, li & ~31);
                        word = 0u;
                    }
                });
        }
        __syncwarp();

        // ---- B1: bounding spheres vs their candidates, fine items -> Q2; B2 whenever the next
        //      round might overflow Q2 and at the end ------------------------------------------------
        {
            uint32_t base = 0u;
            do
            {
                uint32_t n2 = 0u;
                while (base < n1 && n2 + X.round_cap <= X.q2_cap)
                {
                    const uint32_t it = base + lane;
                    uint32_t cnt = 0u, first = 0u;
                    int c = 0, l = 0;
                    bool have = false, bound_hit = false, ask_cloud = false;
                    float x = 0.F, y = 0.F, z = 0.F, rq = 0.F;
                    MaskT hit = 0;
                    LinkInfo L{};
                    if (it < n1)
                    {
                        const uint32_t item = X.q1[it];
                        c = item & 31u;
                        l = item >> 5;
                        if (!((invalid >> c) & 1u))
                        {
                            have = true;
                            L = X.links[l];
                            const SphereTask t = X.tasks[L.bound_task];
                            task_centre<32>(t, X.stash + c, x, y, z);
                            MaskT m = X.masks[l * 32 + c];
                            uint32_t flags = 0u;
                            if constexpr (AE)
                            {
                                flags = static_cast<uint32_t>(m) & ~kAeObjMask;
                                m = static_cast<MaskT>(static_cast<uint32_t>(m) & kAeObjMask);
                            }
                            while (m != 0)
                            {
                                const int o = mask_pop_lowest<MaskT>(m);
                                if (sign_set(margin_obj(X.objs + 4 * o, x, y, z, t.r)))
                                {
                                    hit |= static_cast<MaskT>(1) << o;
                                }
                            }
                            bound_hit = hit != 0 || (flags & kAeHfBit) != 0u;
                            // the pointclouds are queried with the bounding sphere's own, un-inflated radius
                            ask_cloud = AE && (flags & kAeCloudBit) != 0u && !bound_hit;
                            rq = t.r - 1e-6F;
                        }
                    }
                    if constexpr (AE)
                    {
                        // warp-cooperative: the whole round calls, lanes without a question help scanning
                        const EnvHeader &H = *reinterpret_cast<const EnvHeader *>(X.E);
                        if (H.n_capts + H.n_mvts > 0 && sphere_hits_clouds(X.E, x, y, z, rq, ask_cloud))
                        {
                            bound_hit = true;
                        }
                    }
                    if (have && bound_hit)
                    {
                        X.masks[l * 32 + c] = hit;  // what the fine spheres are tested against among the primitives
                        cnt = static_cast<uint32_t>(L.n_spheres);
                        first = static_cast<uint32_t>(L.bound_task + 1);
                    }
                    // exclusive prefix sum of cnt over the lanes
                    uint32_t incl = cnt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1)
                    {
                        const uint32_t v = __shfl_up_sync(kFullWarp, incl, d);
                        incl += (lane >= d) ? v : 0u;
                    }
                    const uint32_t at = n2 + incl - cnt;
                    for (uint32_t k = 0; k < cnt; ++k)
                    {
                        X.q2[at + k] = static_cast<uint16_t>(c | ((first + k) << 5));
                    }
                    n2 += __shfl_sync(kFullWarp, incl, 31);
                    base += 32u;
                }
                __syncwarp();
                invalid = v4_fine_items<R, MaskT, AE>(X, n2, invalid);
                __syncwarp();  // all reads of the queues (and, at the end, of this pass's stash) are done
            } while (base < n1);
        }
        // ---- D: attachment (only for the states everything else has left) --------------------------
        // (an instantiation of its own, launched only when something is attached: with this call in, the plain any-environment
        // kernel needed 158 registers instead of 121 and lost 12 % on BASELINE config 4)
        if constexpr (AE && ATT)
        {
            if (A.n > 0u)
            {
                const bool bad = v4_attachment_any<M>(X.stash + lane, X.objs, X.n_objects, X.E, X.tasks, X.links, X.attach_links, &A,
                                                      has && !((invalid >> lane) & 1u));
                invalid |= __ballot_sync(kFullWarp, bad);
                __syncwarp();  // this pass's stash has been read
            }
        }
        // (instantiations of their own here too: the untaken branch and its call cost the edge kernel 1.2 %, measured A/B)
        if constexpr (!AE && ATT)
        {
            const bool bad = v4_attachment<M>(X.stash + lane, X.objs, X.n_objects, X.tasks, X.links, X.attach_links, &A, has && !((invalid >> lane) & 1u));
            invalid |= __ballot_sync(kFullWarp, bad);
            __syncwarp();  // this pass's stash has been read
        }
        return invalid;
    }

    // Persistent blocks of blockDim.x / 32 autonomous warps; tile t = states [32 t, 32 t + 32).
    template <typename R, typename MaskT, bool TAB, bool GATHER, int MAXT, int MINB, bool AE = false, bool ATT = false>
    __global__ void __launch_bounds__(MAXT, MINB)
        k_validate_configs_v4(
            RobotDev robot,
            const __grid_constant__ GridEnv env,
            const float *__restrict__ q,
            size_t n,
            uint32_t *__restrict__ bits,
            unsigned int *__restrict__ next_tile,
            const __grid_constant__ GatherDev gather)
    {
        using M = typename R::Model;
        extern __shared__ __align__(128) unsigned char smem[];
        __shared__ uint64_t bar;
        const V4Ctx<R, MaskT> X = v4_stage<R, MaskT>(smem, &bar, robot, env);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
        const size_t n_tiles = (n + 31) / 32;
        const size_t stride = static_cast<size_t>(gridDim.x) * warps;
        // Every warp's first tile is its static one; further tiles are drawn from a device-wide counter
        // (zeroed by the host before the launch).  With a static stride the 2^15 tiles of the headline batch
        // are 11.07 per warp: most warps would wait for the ones that got a 12th, and tiles differ in cost
        // (invalid states drop out early).
        size_t tile = static_cast<size_t>(blockIdx.x) * warps + warp;
        while (tile < n_tiles)
        {
            const size_t i = tile * 32 + lane;
            const bool has = i < n;
            float cfg[M::kDof];
#pragma unroll
            for (int j = 0; j < M::kDof; ++j)
            {
                cfg[j] = has ? __ldg(q + i * M::kDof + j) : 0.F;
            }
            // (loading the next tile's configurations ahead of the pass was measured: 2 % slower, the
            // seven extra live registers cost more than the exposed load latency)
#ifndef VMV_V4_STATIC_TILES
            // issued ahead of the pass and consumed after it, unless the kernel is at its register limit
            // (Baxter: the ticket would be one more spilled value across the pass)
            constexpr bool kEarly = MAXT * MINB >= 512;  // register cap <= 128
            unsigned int ticket = 0u;
            if (kEarly && lane == 0)
            {
                ticket = atomicAdd(next_tile, 1u);
            }
#endif
#ifdef VMV_C4_STATS
            const long long t_tile = clock64();
#endif
            const uint32_t invalid = v4_pass<R, MaskT, TAB, true, AE, ATT>(X, env.grid, env.tab, env.att, cfg, has);
#ifdef VMV_C4_STATS
            if (lane == 0)
            {
                // tile durations: log2 buckets 2^13 .. 2^24 cycles at 48..59; total cycles at 60; tiles at 61
                const long long dt = clock64() - t_tile;
                int bkt = 0;
                while ((1ll << (bkt + 13)) < dt && bkt < 11)
                {
                    ++bkt;
                }
                VMV_STAT(48 + bkt, 1);
                VMV_STAT(60, dt);
                VMV_STAT(61, 1);
            }
#endif
            publish_verdict<GATHER>(bits, gather, tile, ~invalid, 4);
#ifndef VMV_V4_STATIC_TILES
            if (!kEarly && lane == 0)
            {
                ticket = atomicAdd(next_tile, 1u);
            }
            tile = stride + __shfl_sync(kFullWarp, ticket, 0);
#else
            tile += stride;
#endif
        }
        gather_finish<GATHER>(gather);
    }

    // Edges.  A warp owns a chunk of 32 edges (one verdict word): lane e keeps edge e's start, vector
    // and rake-step count in registers.  Every pass checks 4 rake blocks (8 tines each) of the
    // reference's schedule (planning/validate.hh:31-64), handed out round-robin over the edges still
    // alive, so an edge found invalid drops its remaining blocks -- the reference's early return.
    template <typename R, typename MaskT, bool TAB, bool INDEXED, bool GATHER, int MAXT, int MINB, bool ATT = false>
    __global__ void __launch_bounds__(MAXT, MINB) k_validate_edges_v4(
        RobotDev robot,
        const __grid_constant__ GridEnv env,
        const float *__restrict__ a,
        const float *__restrict__ b,
        const uint32_t *__restrict__ pairs,
        size_t n,
        float resolution,
        uint32_t *__restrict__ bits,
        unsigned int *__restrict__ next_chunk,
        int edges_per_chunk,
        const __grid_constant__ GatherDev gather)
    {
        using M = typename R::Model;
        extern __shared__ __align__(128) unsigned char smem[];
        __shared__ uint64_t bar;
        const V4Ctx<R, MaskT> X = v4_stage<R, MaskT>(smem, &bar, robot, env);
        const int lane = threadIdx.x & 31;
        // 8, 16 or 32 edges per warp: small batches are cut finer so that they still fill the GPU (a
        // chunk is worked off serially, four rake blocks per pass)
        // (whole verdict words are always written: chunks past the last edge store zeros)
        const size_t n_chunks = ((n + 31) / 32) * static_cast<size_t>(32 / edges_per_chunk);
        const int my_slot = lane >> 3, tine = lane & 7;
        const float pct = static_cast<float>(tine + 1) / 8.F;

        // chunks cost anything from one pass to dozens (edge lengths, early exits): warps draw them
        // from a device-wide counter (zeroed by the host before the launch)
        while (true)
        {
            unsigned int ticket = 0u;
            if (lane == 0)
            {
                ticket = atomicAdd(next_chunk, 1u);
            }
            const size_t chunk = __shfl_sync(kFullWarp, ticket, 0);
            if (chunk >= n_chunks)
            {
                break;
            }
            const size_t edge = lane < edges_per_chunk ? min(chunk * edges_per_chunk + lane, n) : n;
            float start[M::kDof], vec[M::kDof];
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
#pragma unroll
                for (int j = 0; j < M::kDof; ++j)
                {
                    start[j] = __ldg(pa + j);
                    vec[j] = __fsub_rn(__ldg(pb + j), start[j]);
                }
                const float dist = ref_l2_norm<M::kDof>(vec);
                // n = max(ceil(distance / rake * resolution), 1)   (validate.hh:41)
                steps = static_cast<int>(fmaxf(ceilf(__fmul_rn(__fdiv_rn(dist, 8.F), resolution)), 1.F));
            }
            else
            {
#pragma unroll
                for (int j = 0; j < M::kDof; ++j)
                {
                    start[j] = 0.F, vec[j] = 0.F;
                }
            }
            // backstep = vector / (rake * n), once per edge as in validate.hh:43-48
            float back[M::kDof];
#pragma unroll
            for (int j = 0; j < M::kDof; ++j)
            {
                back[j] = steps > 0 ? __fdiv_rn(vec[j], static_cast<float>(8 * steps)) : 0.F;
            }
            int next = 0;
            bool dead = false;

            while (true)
            {
                // hand out up to 4 rake blocks: round r takes the r-th pending block of every live edge
                const int rem = dead ? 0 : steps - next;
                int n_slots = 0, got = 0, owner = 0, round = 0;
#pragma unroll 1
                for (int r = 0; n_slots < 4; ++r)
                {
                    uint32_t m = __ballot_sync(kFullWarp, rem > r);
                    if (m == 0u)
                    {
                        break;
                    }
#pragma unroll 1
                    while (m != 0u && n_slots < 4)
                    {
                        const int o = __ffs(static_cast<int>(m)) - 1;
                        m &= m - 1u;
                        if (n_slots == my_slot)
                        {
                            owner = o, round = r;
                        }
                        got += (o == lane) ? 1 : 0;
                        ++n_slots;
                    }
                }
                if (n_slots == 0)
                {
                    break;
                }
                const bool has = my_slot < n_slots;
                const int step = __shfl_sync(kFullWarp, next, owner) + round;
                const int e_steps = __shfl_sync(kFullWarp, steps, owner);
                next += got;
                float cfg[M::kDof];
                {
                    float bk[M::kDof];
#pragma unroll
                    for (int j = 0; j < M::kDof; ++j)
                    {
                        const float v = __shfl_sync(kFullWarp, vec[j], owner);
                        const float st = __shfl_sync(kFullWarp, start[j], owner);
                        bk[j] = __shfl_sync(kFullWarp, back[j], owner);
                        cfg[j] = fmaf(v, pct, st);
                    }
                    // (one rolled loop over the steps: unrolled per joint this was 8 KB of a kernel that is bound by
                    // instruction fetch -- the body is past the 32 KB instruction cache, profiles/r2_edges_v1)
#pragma unroll 1
                    for (int k = 0; k < step; ++k)
                    {
#pragma unroll
                        for (int j = 0; j < M::kDof; ++j)
                        {
                            cfg[j] = __fsub_rn(cfg[j], bk[j]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < M::kDof; ++j)
                    {
                        cfg[j] = has ? cfg[j] : 0.F;
                    }
                }
                const uint32_t invalid = v4_pass<R, MaskT, TAB, false, false, ATT>(X, env.grid, env.tab, env.att, cfg, has);
#pragma unroll
                for (int p = 0; p < 4; ++p)
                {
                    const int o = __shfl_sync(kFullWarp, owner, 8 * p);
                    const bool bad = p < n_slots && ((invalid >> (8 * p)) & 0xffu) != 0u;
                    dead = dead || (bad && o == lane);
                }
            }

            const uint32_t word = __ballot_sync(kFullWarp, steps > 0 && !dead);
            // verdict bits are little-endian in their words: a chunk of 8 / 16 edges is a byte / half word
            publish_verdict<GATHER>(bits, gather, chunk, word, edges_per_chunk >> 3);
        }
        gather_finish<GATHER>(gather);
    }
}  // namespace vmv
