// Host-side declarations shared by the translation units of libvamp_b200.so: vmv_host.cu (environment,
// C ABI, robot-independent kernels) and one vmv_robot_<name>.cu per robot (that robot's kernels and their
// launchers).  One translation unit per robot because ptxas allocates registers per translation unit: the
// Panda kernels got 88 / 114 registers alone but 92 / 118 next to the other robots' kernels (measured 4-6 %).
#pragma once
#include <atomic>
#include <cstdint>
#include <mutex>
#include <string>

#include <cuda_runtime.h>

#include "../../include/vamp_b200.h"
#include "vmv_kernels_v4.cuh"

namespace vmvh
{
    int fail(int code, const std::string &msg);
    int cuda_fail(cudaError_t e, const char *what);

#define VMV_CUDA(call)                               \
    do                                               \
    {                                                \
        cudaError_t e_ = (call);                     \
        if (e_ != cudaSuccess)                       \
        {                                            \
            return ::vmvh::cuda_fail(e_, #call);     \
        }                                            \
    } while (0)

    constexpr int kMaxDevices = 16;
    constexpr uint32_t kMaxSmem = 227 * 1024;

    extern std::atomic<uint64_t> g_launches;
    extern std::mutex g_mutex;
    // 0: pick automatically; 1: generic per-thread kernel; 2: block-cooperative kernel; 3: grid-culled kernel
    extern std::atomic<int> g_force_path;

    int sm_count();
    unsigned long long *stats_buffer();  // 64 zero-initialised counters on the current device (development builds)

    // A zeroed work counter for one launch of a persistent kernel on stream s.  Slots come from a ring per
    // device; a slot is handed out again only after the event recorded behind its previous launch has
    // completed (counter_release records it), so concurrent launches on any number of streams never share one.
    int counter_acquire(cudaStream_t s, unsigned int *&counter, int &slot);
    int counter_release(cudaStream_t s, int slot);

    struct RobotHost
    {
        const char *name;
        int dof, n_spheres, n_links, n_pairs, n_tasks, resolution, n_attach_links, ee_body;
        const float *lower, *range;
        const vmv::SphereTask *tasks;
        const vmv::LinkInfo *links;
        const vmv::LinkPair *pairs;
        const int *attach_links;
        const float *ee_tf;
        const vmv::PairInfo *pair_info;
        const vmv::SpherePair *pair_lists;
        int n_pair_lists;
        float max_reach;  // farthest any link's bounding sphere extends from the world origin
        const vmv::PairGroupHost *pair_groups;  // two-joint verdict tables (vmv_pairtab.cuh)
        int n_pair_groups;
        const int *pair_group_pairs;
        const int *pair_never;
        int n_pair_never;
        bool inline_covered;
    };

    // Direct stores of verdict words into the peers' symmetric buffers (vmv_comm_*, multi-GPU): world = 0
    // when the launch writes local memory only.
    using GatherDev = vmv::GatherDev;

    // The launchers of one robot's kernels.  Every function returns VMV_OK or an error code.
    struct RobotOps
    {
        const RobotHost *host;
        int (*configs_v4)(int robot, bool wide, const vmv::RobotDev &, vmv::GridEnv, const float *q, size_t n, uint32_t *bits, const GatherDev &, cudaStream_t);
        int (*edges_v4)(int robot, bool wide, const vmv::RobotDev &, vmv::GridEnv, const float *a, const float *b, const uint32_t *pairs, size_t n,
                        float resolution, uint32_t *bits, const GatherDev &, cudaStream_t);
        int (*configs)(const vmv::RobotDev &, const vmv::LaunchEnv &, const float *q, size_t n, uint32_t *bits, cudaStream_t);
        int (*edges)(const vmv::RobotDev &, const vmv::LaunchEnv &, const float *a, const float *b, const uint32_t *pairs, size_t n, float resolution,
                     uint32_t *bits, cudaStream_t);
        int (*fk)(const vmv::RobotDev &, const float *q, size_t n, float *out, cudaStream_t);
        int (*filter)(const vmv::RobotDev &, const vmv::LaunchEnv &, const float *q, const float *pts, size_t n, float r_point, uint32_t *bits, cudaStream_t);
        int (*debug)(const vmv::RobotDev &, const vmv::LaunchEnv &, const float *q, const int *object_ids, int32_t *env_hits, uint32_t cap_env,
                     int32_t *self_hits, uint32_t cap_self, uint32_t *counts);
    };

    extern const RobotOps ops_panda, ops_ur5, ops_fetch, ops_baxter;
    const RobotOps &ops(int robot);
}  // namespace vmvh
