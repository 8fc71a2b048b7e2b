// Launchers of ONE robot's kernels; included by vmv_robot_<name>.cu after that robot's generated headers,
// then instantiated with VMV_DEFINE_ROBOT(name, NAME, tune).  See vmv_internal.h for why each robot is its own
// translation unit.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "vmv_internal.h"

namespace vmvh
{
    using vmv::kFullWarp;

#define VMV_X_LINK(l, r, n, t, reach) f(vmv::IC<l>{}, r, n, t, reach);
#define VMV_X_PAIR(p, a, b, inl) f(vmv::IC<p>{}, vmv::IC<a>{}, vmv::IC<b>{}, vmv::IC<inl>{});
#define VMV_X_MAXREACH(l, r, n, t, reach) m = (reach) > m ? (reach) : m;


    template <typename R, int BLOCK>
    int launch_configs_v2(const vmv::RobotDev &rd, const vmv::LaunchEnv &le, const float *q, size_t n, uint32_t *bits, cudaStream_t s)
    {
        using M = typename R::Model;
        const vmv::SmemLayoutV2<M, BLOCK> L(le.blob_bytes);
        if (L.total > kMaxSmem)
        {
            return fail(VMV_ERR_LIMIT, "environment too large for shared-memory staging");
        }
        auto kernel = vmv::k_validate_configs_v2<R, BLOCK>;
        VMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.total)));
        const unsigned grid = static_cast<unsigned>((n + BLOCK - 1) / BLOCK);
        kernel<<<grid, BLOCK, L.total, s>>>(rd, le, q, n, bits);
        g_launches++;
        VMV_CUDA(cudaGetLastError());
        return VMV_OK;
    }

    // Two-joint verdict tables (vmv_pairtab.cuh): built once per robot and device.
    struct PairTabCache
    {
        bool ready = false;
        vmv::PairTabDev dev{};
        float build_ms = 0.F;
    };
    constexpr size_t kPairTabBytes = size_t(4) << 20;  // all groups of a robot together

    template <typename R>
    int ensure_pair_tables(const RobotHost &r, const vmv::RobotDev &rd, vmv::PairTabDev &out)
    {
        int device = 0;
        VMV_CUDA(cudaGetDevice(&device));
        std::lock_guard<std::mutex> lock(g_mutex);
        static PairTabCache cache[kMaxDevices];  // this translation unit = this robot
        if (device >= kMaxDevices)
        {
            return fail(VMV_ERR_ARG, "device index too large");
        }
        PairTabCache &c = cache[device];
        if (!c.ready)
        {
            vmv::PairTabDev t{};
            for (int k = 0; k < r.n_pair_never; ++k)
            {
                t.never_pairs[r.pair_never[k] >> 5] |= 1u << (r.pair_never[k] & 31);
            }
            const int ng = std::min(r.n_pair_groups, vmv::kPairTabMaxGroups);
            t.n_groups = ng;
            cudaEvent_t e0, e1;
            VMV_CUDA(cudaEventCreate(&e0));
            VMV_CUDA(cudaEventCreate(&e1));
            VMV_CUDA(cudaEventRecord(e0, nullptr));
            for (int g = 0; g < ng; ++g)
            {
                const vmv::PairGroupHost &G = r.pair_groups[g];
                const size_t budget = kPairTabBytes / ng;
                int na, nb;
                if (G.dof[1] < 0)
                {
                    na = static_cast<int>(std::min<size_t>(budget, 1u << 16));
                    nb = 1;
                }
                else
                {
                    // cell widths in inverse proportion to the Lipschitz constants: equal band shares
                    const double wa = (G.hi[0] - G.lo[0]) * std::max(G.lip[0], 1e-3F), wb = (G.hi[1] - G.lo[1]) * std::max(G.lip[1], 1e-3F);
                    const double cells = static_cast<double>(budget);
                    na = std::max(16, static_cast<int>(std::sqrt(cells * wa / wb)));
                    nb = std::max(16, static_cast<int>(cells / na));
                }
                const float step_a = (G.hi[0] - G.lo[0]) / na, step_b = G.dof[1] < 0 ? 1.F : (G.hi[1] - G.lo[1]) / nb;
                const float band = 0.5F * (G.lip[0] * step_a + (G.dof[1] < 0 ? 0.F : G.lip[1] * step_b)) + 2e-5F;
                void *cells = nullptr, *d_pairs = nullptr;
                VMV_CUDA(cudaMalloc(&cells, static_cast<size_t>(na) * nb));
                VMV_CUDA(cudaMalloc(&d_pairs, G.count * sizeof(int)));
                VMV_CUDA(cudaMemcpy(d_pairs, r.pair_group_pairs + G.first, G.count * sizeof(int), cudaMemcpyHostToDevice));
                const size_t n_cells = static_cast<size_t>(na) * nb;
                vmv::k_build_pair_table<R><<<static_cast<unsigned>((n_cells + 127) / 128), 128>>>(
                    rd, G.dof[0], G.dof[1], G.lo[0], step_a, G.lo[1], step_b, na, nb, band, static_cast<const int *>(d_pairs), G.count,
                    static_cast<unsigned char *>(cells));
                g_launches++;
                VMV_CUDA(cudaGetLastError());
                VMV_CUDA(cudaDeviceSynchronize());
                cudaFree(d_pairs);
                t.dof_a[g] = G.dof[0], t.dof_b[g] = G.dof[1];
                t.lo_a[g] = G.lo[0], t.inv_a[g] = 1.F / step_a;
                t.lo_b[g] = G.lo[1], t.inv_b[g] = G.dof[1] < 0 ? 0.F : 1.F / step_b;
                t.na[g] = na, t.nb[g] = nb;
                t.cells[g] = static_cast<const unsigned char *>(cells);
                for (int k = 0; k < G.count; ++k)
                {
                    const int p = r.pair_group_pairs[G.first + k];
                    t.group_pairs[g][p >> 5] |= 1u << (p & 31);
                }
            }
            VMV_CUDA(cudaEventRecord(e1, nullptr));
            VMV_CUDA(cudaEventSynchronize(e1));
            VMV_CUDA(cudaEventElapsedTime(&c.build_ms, e0, e1));
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
            c.dev = t;
            c.ready = true;
        }
        out = c.dev;
        return VMV_OK;
    }

    // Launch geometry of the warp-autonomous kernels: the number of warps per block that maximises
    // the warps resident per SM (shared memory = block-shared tables + one slice per warp), and a
    // persistent grid of that many blocks per SM.
    // The search runs once per (kernel, shared-memory shape, device) and is cached: the occupancy
    // queries cost more host time than a launch.
    struct V4Geometry
    {
        const void *kernel;
        uint32_t shared_bytes, warp_bytes;
        int device, warps, blocks_per_sm;
    };
    inline std::vector<V4Geometry> g_v4_geometry;

    template <typename K>
    int v4_geometry(K kernel, uint32_t shared_bytes, uint32_t warp_bytes, int max_threads, size_t units, int &warps, unsigned &grid, uint32_t &smem)
    {
        int device = 0;
        VMV_CUDA(cudaGetDevice(&device));
        const void *key = reinterpret_cast<const void *>(kernel);
        int best_w = 0, best_blocks = 0;
        {
            std::lock_guard<std::mutex> lock(g_mutex);
            for (const auto &g : g_v4_geometry)
            {
                if (g.kernel == key && g.shared_bytes == shared_bytes && g.warp_bytes == warp_bytes && g.device == device)
                {
                    best_w = g.warps, best_blocks = g.blocks_per_sm;
                    break;
                }
            }
        }
        if (best_w == 0)
        {
            constexpr uint32_t kMaxDynamic = kMaxSmem - 1024;  // the opt-in limit counts static shared memory too
            VMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMaxDynamic)));
            for (int w = 1; w <= max_threads / 32; ++w)
            {
                const size_t bytes = static_cast<size_t>(shared_bytes) + static_cast<size_t>(w) * warp_bytes;
                if (bytes > kMaxDynamic)
                {
                    break;
                }
                int blocks = 0;
                VMV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, w * 32, bytes));
                if (blocks * w > best_blocks * best_w || (blocks * w == best_blocks * best_w && blocks > 0))
                {
                    best_w = w;
                    best_blocks = blocks;
                }
            }
            if (best_w == 0 || best_blocks == 0)
            {
                return fail(VMV_ERR_LIMIT, "robot tables too large for the grid-culled kernel");
            }
            std::lock_guard<std::mutex> lock(g_mutex);
            g_v4_geometry.push_back({key, shared_bytes, warp_bytes, device, best_w, best_blocks});
        }
        warps = best_w;
        smem = shared_bytes + static_cast<uint32_t>(best_w) * warp_bytes;
        const size_t tiles = (units + 31) / 32;
        const size_t blocks_needed = (tiles + best_w - 1) / best_w;
        grid = static_cast<unsigned>(std::min<size_t>(blocks_needed, static_cast<size_t>(sm_count()) * best_blocks));
        return VMV_OK;
    }
    // Launch bounds (threads per block, blocks per SM -> register cap) and fine-item queue depth of the
    // warp-autonomous kernels, measured per robot.  The configuration kernel needs ~90 registers
    // (Panda) and gains from 19-20 resident warps (one block, one-round queue: +9 %); the edge kernel
    // carries the edge state in registers and is best at 16 warps with the deeper queue.
    // TUNE 0: Panda / UR5 / Fetch.  TUNE 1 (Baxter): <= 255 registers; measured best of (128,2) (192,2) (256,2) (256,1).
    template <int TUNE>
    struct V4Tune
    {
#if defined(VMV_V4_MAXT) && defined(VMV_V4_MINB)
        static constexpr int kCfgThreads = VMV_V4_MAXT, kCfgBlocks = VMV_V4_MINB;
#else
        static constexpr int kCfgThreads = 640, kCfgBlocks = 1;  // <= 102 registers
#endif
        static constexpr int kCfgQ2Rounds = 1;
#if defined(VMV_V4_EDGE_MAXT) && defined(VMV_V4_EDGE_MINB)
        static constexpr int kEdgeThreads = VMV_V4_EDGE_MAXT, kEdgeBlocks = VMV_V4_EDGE_MINB;
#else
        static constexpr int kEdgeThreads = 256, kEdgeBlocks = 2;  // <= 128 registers
#endif
        static constexpr int kEdgeQ2Rounds = 2;
    };
    template <>
    struct V4Tune<1>
    {
        static constexpr int kCfgThreads = 256, kCfgBlocks = 1, kCfgQ2Rounds = 2;
        static constexpr int kEdgeThreads = 256, kEdgeBlocks = 1, kEdgeQ2Rounds = 2;
    };

    template <typename R, typename MaskT, bool GATHER, bool AE = false, bool ATT = false>
    int launch_configs_v4(const RobotHost &rh, const vmv::RobotDev &rd, vmv::GridEnv le, const float *q, size_t n, uint32_t *bits,
                          GatherDev gather, cudaStream_t s)
    {
        using M = typename R::Model;
        using Tune = V4Tune<R::kTune>;
        // the any-environment instantiation carries the pointcloud queries: blocks of 256 threads
        // (measured on BASELINE config 4, Fetch / UR5: 256 threads 0.522 / 0.348 ms, 320 0.522 / 0.352, 384 0.523 / 0.351,
        // 448 0.523 / 0.376, 512 0.560 / 0.401, 640 0.532 / 0.362 -- the kernel does not need more than 14 resident warps, and
        // under the 128-register cap of a 512-thread block the compiler spills; two blocks of 256 fit per SM)
#ifndef VMV_AE_THREADS
#define VMV_AE_THREADS 256
#endif
        constexpr int kThreads = AE ? (Tune::kCfgThreads > VMV_AE_THREADS ? VMV_AE_THREADS : Tune::kCfgThreads) : Tune::kCfgThreads;
        if (R::PairTab::kUseTables && !rh.inline_covered)
        {
            return fail(VMV_ERR_LIMIT, "robot's inline self-collision pairs are not covered by verdict tables");
        }
        {
            int rc = ensure_pair_tables<R>(rh, rd, le.tab);
            if (rc != VMV_OK)
            {
                return rc;
            }
        }
        le.q2_rounds = Tune::kCfgQ2Rounds;
        if (!AE)
        {
            le.blob = nullptr, le.blob_bytes = 0;
        }
        const vmv::SmemLayoutV4<M, MaskT> L(le.n_objects, le.max_fine, le.q2_rounds, le.blob_bytes);
        // (any environment with an attachment: two blocks of 256 per SM, so at most 128 registers -- left alone the compiler
        // takes 158 and only one block fits)
        constexpr int kBlocks = (AE && ATT && kThreads <= 256) ? 2 : Tune::kCfgBlocks;
        auto kernel = vmv::k_validate_configs_v4<R, MaskT, R::PairTab::kUseTables, GATHER, kThreads, kBlocks, AE, ATT>;
        int warps = 0;
        unsigned grid = 0;
        uint32_t smem = 0;
        int rc = v4_geometry(kernel, L.shared_bytes, L.warp_bytes, kThreads, n, warps, grid, smem);
        if (rc != VMV_OK)
        {
            return rc;
        }
#ifdef VMV_C4_STATS
        {
            unsigned long long *sp = stats_buffer();
            VMV_CUDA(cudaMemcpyToSymbol(vmv::g_stats, &sp, sizeof(sp)));
        }
#endif
        unsigned int *counter = nullptr;
        int slot = -1;
        rc = counter_acquire(s, counter, slot);
        if (rc != VMV_OK)
        {
            return rc;
        }
        gather.done = counter + 1;
        kernel<<<grid, warps * 32, smem, s>>>(rd, le, q, n, bits, counter, gather);
        g_launches++;
        const cudaError_t e = cudaGetLastError();
        rc = counter_release(s, slot);
        if (e != cudaSuccess)
        {
            return cuda_fail(e, "k_validate_configs_v4 launch");
        }
        return rc;
    }

    template <typename R, typename MaskT, bool GATHER, bool ATT = false>
    int launch_edges_v4(const RobotHost &rh, const vmv::RobotDev &rd, vmv::GridEnv le, const float *a, const float *b, const uint32_t *pairs, size_t n,
                        float resolution, uint32_t *bits, GatherDev gather, cudaStream_t s)
    {
        using M = typename R::Model;
        using Tune = V4Tune<R::kTune>;
        if (R::PairTab::kUseTables && !rh.inline_covered)
        {
            return fail(VMV_ERR_LIMIT, "robot's inline self-collision pairs are not covered by verdict tables");
        }
        {
            int rc = ensure_pair_tables<R>(rh, rd, le.tab);
            if (rc != VMV_OK)
            {
                return rc;
            }
        }
        le.q2_rounds = Tune::kEdgeQ2Rounds;
        const vmv::SmemLayoutV4<M, MaskT> L(le.n_objects, le.max_fine, le.q2_rounds);
        int warps = 0;
        unsigned grid = 0;
        uint32_t smem = 0;
        // finer chunks for batches that would otherwise leave warps idle (vmv_kernels_v4.cuh); measured: 2^18 edges
        // 16 per chunk 224 M/s, 32 per chunk 217 M/s (3.5 chunks per warp leave a long tail); 9.5 M edges 4.2e8 / 4.4e8
        static const int chunk_override = []
        {
            const char *e = std::getenv("VMV_EDGE_CHUNK");  // 8, 16 or 32 (tuning aid)
            const int v = e ? std::atoi(e) : 0;
            return (v == 8 || v == 16 || v == 32) ? v : 0;
        }();
        const int per_chunk = chunk_override ? chunk_override : (n >= (size_t(1) << 19) ? 32 : (n >= (size_t(1) << 15) ? 16 : 8));
        const size_t units = (n + 31) / 32 * 32 * (32 / per_chunk);
        unsigned int *counter = nullptr;
        int slot = -1;
        cudaError_t e = cudaSuccess;
        if (pairs != nullptr)
        {
            auto kernel = vmv::k_validate_edges_v4<R, MaskT, R::PairTab::kUseTables, true, GATHER, Tune::kEdgeThreads, Tune::kEdgeBlocks, ATT>;
            int rc = v4_geometry(kernel, L.shared_bytes, L.warp_bytes, Tune::kEdgeThreads, units, warps, grid, smem);
            rc = rc == VMV_OK ? counter_acquire(s, counter, slot) : rc;
            if (rc != VMV_OK)
            {
                return rc;
            }
            gather.done = counter + 1;
            kernel<<<grid, warps * 32, smem, s>>>(rd, le, a, b, pairs, n, resolution, bits, counter, per_chunk, gather);
            e = cudaGetLastError();
        }
        else
        {
            auto kernel = vmv::k_validate_edges_v4<R, MaskT, R::PairTab::kUseTables, false, GATHER, Tune::kEdgeThreads, Tune::kEdgeBlocks, ATT>;
            int rc = v4_geometry(kernel, L.shared_bytes, L.warp_bytes, Tune::kEdgeThreads, units, warps, grid, smem);
            rc = rc == VMV_OK ? counter_acquire(s, counter, slot) : rc;
            if (rc != VMV_OK)
            {
                return rc;
            }
            gather.done = counter + 1;
            kernel<<<grid, warps * 32, smem, s>>>(rd, le, a, b, pairs, n, resolution, bits, counter, per_chunk, gather);
            e = cudaGetLastError();
        }
        g_launches++;
        const int rc = counter_release(s, slot);
        if (e != cudaSuccess)
        {
            return cuda_fail(e, "k_validate_edges_v4 launch");
        }
        return rc;
    }

    template <typename R, int BLOCK>
    int launch_configs(const vmv::RobotDev &rd, const vmv::LaunchEnv &le, const float *q, size_t n, uint32_t *bits, cudaStream_t s)
    {
        using M = typename R::Model;
        const int force = g_force_path.load();
        if (force != 1 && le.primitives_only && le.n_objects <= 64)
        {
            const vmv::SmemLayoutV2<M, BLOCK> L2(le.blob_bytes);
            if (L2.total <= kMaxSmem)
            {
                return launch_configs_v2<R, BLOCK>(rd, le, q, n, bits, s);
            }
        }
        if (force == 2)
        {
            return fail(VMV_ERR_LIMIT, "block-cooperative kernel not applicable to this environment");
        }
        const vmv::SmemLayout<M, BLOCK> L(le.blob_bytes);
        if (L.total > kMaxSmem)
        {
            return fail(VMV_ERR_LIMIT, "environment too large for shared-memory staging");
        }
        auto kernel = vmv::k_validate_configs<R, BLOCK>;
        VMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.total)));
#ifdef VMV_C4_STATS
        {
            unsigned long long *sp = stats_buffer();
            VMV_CUDA(cudaMemcpyToSymbol(vmv::g_stats, &sp, sizeof(sp)));
        }
#endif
        const unsigned grid = static_cast<unsigned>((n + BLOCK - 1) / BLOCK);
        kernel<<<grid, BLOCK, L.total, s>>>(rd, le, q, n, bits);
        g_launches++;
        VMV_CUDA(cudaGetLastError());
        return VMV_OK;
    }

    template <typename R, int BLOCK>
    int launch_edges(
        const vmv::RobotDev &rd,
        const vmv::LaunchEnv &le,
        const float *a,
        const float *b,
        const uint32_t *pairs,
        size_t n,
        float resolution,
        uint32_t *bits,
        cudaStream_t s)
    {
        using M = typename R::Model;
        const vmv::SmemLayout<M, BLOCK> L(le.blob_bytes);
        if (L.total > kMaxSmem)
        {
            return fail(VMV_ERR_LIMIT, "environment too large for shared-memory staging");
        }
        const int force = g_force_path.load();
        if (BLOCK == 128 && force != 1 && le.primitives_only && le.n_objects <= 64)
        {
            const vmv::SmemLayoutV2<M, 128> L2(le.blob_bytes);
            if (L2.total <= kMaxSmem)
            {
                const size_t chunks = (n + 31) / 32;
                const int per_sm2 = std::max<int>(1, static_cast<int>(kMaxSmem / std::max<uint32_t>(L2.total + 4096, 1)));
                const unsigned grid2 = static_cast<unsigned>(std::min<size_t>(chunks, static_cast<size_t>(sm_count()) * per_sm2 * 8));
                if (pairs != nullptr)
                {
                    auto kernel = vmv::k_validate_edges_v2<R, 128, true>;
                    VMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L2.total)));
                    kernel<<<grid2, 128, L2.total, s>>>(rd, le, a, b, pairs, n, resolution, bits);
                }
                else
                {
                    auto kernel = vmv::k_validate_edges_v2<R, 128, false>;
                    VMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L2.total)));
                    kernel<<<grid2, 128, L2.total, s>>>(rd, le, a, b, pairs, n, resolution, bits);
                }
                g_launches++;
                VMV_CUDA(cudaGetLastError());
                return VMV_OK;
            }
        }
        if (force == 2)
        {
            return fail(VMV_ERR_LIMIT, "block-cooperative kernel not applicable to this environment");
        }
        const size_t rounds = (n + 31) / 32;  // one warp-round per verdict word
        const size_t blocks_needed = (rounds + BLOCK / 32 - 1) / (BLOCK / 32);
        const int per_sm = std::max<int>(1, static_cast<int>(kMaxSmem / std::max<uint32_t>(L.total, 1)));
        const unsigned grid = static_cast<unsigned>(std::min<size_t>(blocks_needed, static_cast<size_t>(sm_count()) * per_sm * 4));
        if (pairs != nullptr)
        {
            auto kernel = vmv::k_validate_edges<R, BLOCK, true>;
            VMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.total)));
            kernel<<<grid, BLOCK, L.total, s>>>(rd, le, a, b, pairs, n, resolution, bits);
        }
        else
        {
            auto kernel = vmv::k_validate_edges<R, BLOCK, false>;
            VMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.total)));
            kernel<<<grid, BLOCK, L.total, s>>>(rd, le, a, b, pairs, n, resolution, bits);
        }
        g_launches++;
        VMV_CUDA(cudaGetLastError());
        return VMV_OK;
    }

    template <typename R>
    int launch_fk(const vmv::RobotDev &rd, const float *q, size_t n, float *out, cudaStream_t s)
    {
        constexpr int BLOCK = 128;
        const unsigned grid = static_cast<unsigned>((n + BLOCK - 1) / BLOCK);
        vmv::k_sphere_fk<R, BLOCK><<<grid, BLOCK, 0, s>>>(rd, q, n, out);
        g_launches++;
        VMV_CUDA(cudaGetLastError());
        return VMV_OK;
    }

    template <typename R>
    int launch_filter(const vmv::RobotDev &rd, const vmv::LaunchEnv &le, const float *q, const float *pts, size_t n, float r_point, uint32_t *bits, cudaStream_t s)
    {
        using M = typename R::Model;
        constexpr int BLOCK = 128;
        const uint32_t smem = ((le.blob_bytes + 15u) & ~15u) + M::kSpheres * sizeof(float4);
        if (smem > kMaxSmem - 1024)
        {
            return fail(VMV_ERR_LIMIT, "environment too large for shared-memory staging");
        }
        auto kernel = vmv::k_filter_points<R, BLOCK>;
        VMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        const size_t tiles = (n + BLOCK - 1) / BLOCK;
        const unsigned grid = static_cast<unsigned>(std::min<size_t>(tiles, static_cast<size_t>(sm_count()) * 8));
        kernel<<<grid, BLOCK, smem, s>>>(rd, le, q, pts, n, r_point, bits);
        g_launches++;
        VMV_CUDA(cudaGetLastError());
        return VMV_OK;
    }

    template <typename R>
    int launch_debug(const vmv::RobotDev &rd, const vmv::LaunchEnv &le, const float *q, const int *object_ids, int32_t *env_hits, uint32_t cap_env,
                     int32_t *self_hits, uint32_t cap_self, uint32_t *counts)
    {
        constexpr int BLOCK = 64;
        const vmv::SmemLayout<typename R::Model, BLOCK> L(le.blob_bytes);
        if (L.total > kMaxSmem)
        {
            return fail(VMV_ERR_LIMIT, "environment too large for shared-memory staging");
        }
        auto kernel = vmv::k_debug<R, BLOCK>;
        VMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.total)));
        kernel<<<1, BLOCK, L.total, nullptr>>>(rd, le, q, object_ids, env_hits, cap_env, self_hits, cap_self, counts);
        g_launches++;
        VMV_CUDA(cudaGetLastError());
        return VMV_OK;
    }

    // ---- the robot's operation table ------------------------------------------------------------------
    template <typename R, int BLOCK>
    struct RobotOpsOf
    {
        static const RobotHost &host();

        static int configs_v4(int, bool wide, const vmv::RobotDev &rd, vmv::GridEnv le, const float *q, size_t n, uint32_t *bits, const GatherDev &g,
                              cudaStream_t s)
        {
            if (le.blob_bytes > 0)
            {
                // heightfields / pointclouds next to at most 30 primitives: the any-environment instantiation (local
                // launches; a gather is pushed by the caller afterwards)
                // (Baxter -- tuning class 1, 255 registers -- is faster on the per-thread kernel: measured 1.28e9 vs 0.63e9
                // configurations/s on a heightfield scene)
                if (wide || g.world > 0 || R::Model::kLinks > 64 || R::kTune != 0)
                {
                    return fail(VMV_ERR_LIMIT, "any-environment grid kernel: not applicable");
                }
                return le.att.n > 0 ? launch_configs_v4<R, uint32_t, false, true, true>(host(), rd, le, q, n, bits, g, s)
                                    : launch_configs_v4<R, uint32_t, false, true>(host(), rd, le, q, n, bits, g, s);
            }
            if (g.world > 0 && le.att.n > 0)
            {
                return fail(VMV_ERR_LIMIT, "grid kernel with fused gather: not applicable with an attachment");
            }
            if (g.world > 0)
            {
                return wide ? launch_configs_v4<R, unsigned long long, true>(host(), rd, le, q, n, bits, g, s)
                            : launch_configs_v4<R, uint32_t, true>(host(), rd, le, q, n, bits, g, s);
            }
            if (le.att.n > 0)
            {
                // something is attached: the instantiations that carry phase D (local launches; a gather is pushed afterwards)
                return wide ? launch_configs_v4<R, unsigned long long, false, false, true>(host(), rd, le, q, n, bits, g, s)
                            : launch_configs_v4<R, uint32_t, false, false, true>(host(), rd, le, q, n, bits, g, s);
            }
            return wide ? launch_configs_v4<R, unsigned long long, false>(host(), rd, le, q, n, bits, g, s)
                        : launch_configs_v4<R, uint32_t, false>(host(), rd, le, q, n, bits, g, s);
        }

        static int edges_v4(int, bool wide, const vmv::RobotDev &rd, vmv::GridEnv le, const float *a, const float *b, const uint32_t *pairs, size_t n,
                            float resolution, uint32_t *bits, const GatherDev &g, cudaStream_t s)
        {
            if (g.world > 0 && le.att.n > 0)
            {
                return fail(VMV_ERR_LIMIT, "grid kernel with fused gather: not applicable with an attachment");
            }
            if (g.world > 0)
            {
                return wide ? launch_edges_v4<R, unsigned long long, true>(host(), rd, le, a, b, pairs, n, resolution, bits, g, s)
                            : launch_edges_v4<R, uint32_t, true>(host(), rd, le, a, b, pairs, n, resolution, bits, g, s);
            }
            if (le.att.n > 0)
            {
                return wide ? launch_edges_v4<R, unsigned long long, false, true>(host(), rd, le, a, b, pairs, n, resolution, bits, g, s)
                            : launch_edges_v4<R, uint32_t, false, true>(host(), rd, le, a, b, pairs, n, resolution, bits, g, s);
            }
            return wide ? launch_edges_v4<R, unsigned long long, false>(host(), rd, le, a, b, pairs, n, resolution, bits, g, s)
                        : launch_edges_v4<R, uint32_t, false>(host(), rd, le, a, b, pairs, n, resolution, bits, g, s);
        }

        static int configs(const vmv::RobotDev &rd, const vmv::LaunchEnv &le, const float *q, size_t n, uint32_t *bits, cudaStream_t s)
        {
            return launch_configs<R, BLOCK>(rd, le, q, n, bits, s);
        }

        static int edges(const vmv::RobotDev &rd, const vmv::LaunchEnv &le, const float *a, const float *b, const uint32_t *pairs, size_t n,
                         float resolution, uint32_t *bits, cudaStream_t s)
        {
            return launch_edges<R, BLOCK>(rd, le, a, b, pairs, n, resolution, bits, s);
        }
    };
}  // namespace vmvh

// name: panda | ur5 | fetch | baxter; NAME: the same in capitals (the generated dispatch macros); TUNE: V4Tune
// class; BLOCK: threads per block of the per-thread / block-cooperative kernels
#define VMV_DEFINE_ROBOT(name, NAME, TUNE, BLOCK)                                                                  \
    namespace vmvh                                                                                                 \
    {                                                                                                              \
        struct name##_robot                                                                                        \
        {                                                                                                          \
            using Model = vmv::gen::name##_model;                                                                  \
            using PairTab = vmv::gen::name##_pairtab_traits;                                                       \
            static constexpr int kTune = TUNE;                                                                     \
            template <typename Sink>                                                                               \
            static __device__ __forceinline__ void frames(const float (&q)[Model::kDof], Sink &s)                  \
            {                                                                                                      \
                vmv::gen::name##_frames(q, s);                                                                     \
            }                                                                                                      \
            template <typename F>                                                                                  \
            static __device__ __forceinline__ void for_each_link(F &&f)                                            \
            {                                                                                                      \
                VMV_##NAME##_LINKS(VMV_X_LINK)                                                                     \
            }                                                                                                      \
            template <typename F>                                                                                  \
            static __device__ __forceinline__ void for_each_pair(F &&f)                                            \
            {                                                                                                      \
                VMV_##NAME##_PAIRS(VMV_X_PAIR)                                                                     \
            }                                                                                                      \
        };                                                                                                         \
        static float name##_max_reach()                                                                            \
        {                                                                                                          \
            float m = 0.F;                                                                                         \
            VMV_##NAME##_LINKS(VMV_X_MAXREACH)                                                                     \
            return m;                                                                                              \
        }                                                                                                          \
        static const RobotHost name##_host = {                                                                     \
            vmv::gen::name##_model::kName, vmv::gen::name##_model::kDof, vmv::gen::name##_model::kSpheres,         \
            vmv::gen::name##_model::kLinks, vmv::gen::name##_model::kPairs, vmv::gen::name##_model::kTasks,        \
            vmv::gen::name##_model::kResolution, vmv::gen::name##_model::kAttachLinks,                             \
            vmv::gen::name##_model::kEeBody, vmv::gen::name##_lower, vmv::gen::name##_range,                       \
            vmv::gen::name##_tasks_host, vmv::gen::name##_links_host, vmv::gen::name##_pairs_host,                 \
            vmv::gen::name##_attach_links_host, vmv::gen::name##_ee_tf_host, vmv::gen::name##_pair_info_host,      \
            vmv::gen::name##_pair_lists_host, vmv::gen::name##_pair_lists_count, name##_max_reach(),               \
            vmv::gen::name##_pair_groups, vmv::gen::name##_pair_group_count, vmv::gen::name##_pair_group_pairs,    \
            vmv::gen::name##_pair_never, vmv::gen::name##_pair_never_count, vmv::gen::name##_inline_covered};      \
        template <>                                                                                                \
        const RobotHost &RobotOpsOf<name##_robot, BLOCK>::host()                                                   \
        {                                                                                                          \
            return name##_host;                                                                                    \
        }                                                                                                          \
        const RobotOps ops_##name = {                                                                              \
            &name##_host,                                                                                          \
            &RobotOpsOf<name##_robot, BLOCK>::configs_v4,                                                          \
            &RobotOpsOf<name##_robot, BLOCK>::edges_v4,                                                            \
            &RobotOpsOf<name##_robot, BLOCK>::configs,                                                             \
            &RobotOpsOf<name##_robot, BLOCK>::edges,                                                               \
            &launch_fk<name##_robot>,                                                                              \
            &launch_filter<name##_robot>,                                                                          \
            &launch_debug<name##_robot>};                                                                          \
    }
