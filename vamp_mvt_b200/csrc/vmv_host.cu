// Host side of the C ABI (include/vamp_b200.h): environment construction, packing and upload, the
// robot-independent kernels (voxel tables, nearest-point table, CAPT build, Halton, CenterVox, gather), the multi-GPU
// communicator, and the entry points, which dispatch to the per-robot launchers (vmv_robot_<name>.cu)
// through vmvh::RobotOps.  No torch, no CPU fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <new>
#include <numeric>
#include <string>
#include <vector>

#include <dlfcn.h>

#include <cuda_runtime.h>

#include "vmv_internal.h"
#include "vmv_capt_build.cuh"
#include "vmv_comm.cuh"
#include "vmv_halton.cuh"
#include "vmv_filter.cuh"

namespace vmvh
{
    thread_local std::string g_error;
    std::atomic<uint64_t> g_launches{0};
    std::mutex g_mutex;
    std::atomic<int> g_force_path{0};

    int fail(int code, const std::string &msg)
    {
        g_error = msg;
        return code;
    }

    int cuda_fail(cudaError_t e, const char *what)
    {
        return fail(VMV_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    }

    const RobotOps &ops(int robot)
    {
        switch (robot)
        {
            case VMV_PANDA:
                return ops_panda;
            case VMV_UR5:
                return ops_ur5;
            case VMV_FETCH:
                return ops_fetch;
            default:
                return ops_baxter;
        }
    }
}  // namespace vmvh

using namespace vmvh;

namespace
{
    const RobotHost &robot_host(int robot)
    {
        return *ops(robot).host;
    }

    struct RobotDevTables
    {
        bool ready = false;
        vmv::RobotDev dev{};
    };
    RobotDevTables g_dev_tables[kMaxDevices][VMV_N_ROBOTS];
    std::mutex g_mutex;

    int robot_tables(int robot, vmv::RobotDev &out)
    {
        int device = 0;
        VMV_CUDA(cudaGetDevice(&device));
        if (device >= kMaxDevices)
        {
            return fail(VMV_ERR_ARG, "device index too large");
        }
        std::lock_guard<std::mutex> lock(g_mutex);
        RobotDevTables &t = g_dev_tables[device][robot];
        if (!t.ready)
        {
            const RobotHost &r = robot_host(robot);
            void *p = nullptr;
            VMV_CUDA(cudaMalloc(&p, r.n_tasks * sizeof(vmv::SphereTask)));
            VMV_CUDA(cudaMemcpy(p, r.tasks, r.n_tasks * sizeof(vmv::SphereTask), cudaMemcpyHostToDevice));
            t.dev.tasks = static_cast<const vmv::SphereTask *>(p);
            VMV_CUDA(cudaMalloc(&p, r.n_links * sizeof(vmv::LinkInfo)));
            VMV_CUDA(cudaMemcpy(p, r.links, r.n_links * sizeof(vmv::LinkInfo), cudaMemcpyHostToDevice));
            t.dev.links = static_cast<const vmv::LinkInfo *>(p);
            const int np = std::max(1, r.n_pairs);
            VMV_CUDA(cudaMalloc(&p, np * sizeof(vmv::LinkPair)));
            VMV_CUDA(cudaMemcpy(p, r.pairs, np * sizeof(vmv::LinkPair), cudaMemcpyHostToDevice));
            t.dev.pairs = static_cast<const vmv::LinkPair *>(p);
            const int na = std::max(1, r.n_attach_links);
            VMV_CUDA(cudaMalloc(&p, na * sizeof(int)));
            VMV_CUDA(cudaMemcpy(p, r.attach_links, na * sizeof(int), cudaMemcpyHostToDevice));
            t.dev.attach_links = static_cast<const int *>(p);
            VMV_CUDA(cudaMalloc(&p, np * sizeof(vmv::PairInfo)));
            VMV_CUDA(cudaMemcpy(p, r.pair_info, np * sizeof(vmv::PairInfo), cudaMemcpyHostToDevice));
            t.dev.pair_info = static_cast<const vmv::PairInfo *>(p);
            const int nl = std::max(1, r.n_pair_lists);
            VMV_CUDA(cudaMalloc(&p, nl * sizeof(vmv::SpherePair)));
            VMV_CUDA(cudaMemcpy(p, r.pair_lists, nl * sizeof(vmv::SpherePair), cudaMemcpyHostToDevice));
            t.dev.pair_lists = static_cast<const vmv::SpherePair *>(p);
            t.dev.n_pair_lists = r.n_pair_lists;
            t.ready = true;
        }
        out = t.dev;
        return VMV_OK;
    }

    // ---------------------------------------------------------------------------------------
    // environment (host restatement of collision/shapes.hh min_distance, environment.cc
    // classification, environment.hh sort; CAPT build per collision/capt.hh:106-369)
    // ---------------------------------------------------------------------------------------
    inline float dot3(float ax, float ay, float az, float bx, float by, float bz)
    {
        return (ax * bx) + (ay * by) + (az * bz);
    }

    inline float clamp_ref(float v, float lo, float hi)
    {
        // collision::clamp<float>, math.hh:52-55
        return std::max(std::min(v, hi), lo);
    }

    struct HSphere
    {
        float x, y, z, r, min_d;
        int id;
    };
    struct HCapsule
    {
        float x1, y1, z1, xv, yv, zv, r, rdv, min_d;
        int id;
    };
    struct HCuboid
    {
        float f[15];
        float min_d;
        int id;
    };
    struct HHeight
    {
        float f[6];
        size_t xd, yd;
        std::vector<float> data;
        int id;
        float *d_data = nullptr;
    };

    // std::vector whose resize() leaves new elements uninitialised (the CAPT point array is GBs and is
    // filled by parallel copies right after it is sized)
    template <typename T>
    struct NoInitAlloc : std::allocator<T>
    {
        template <typename U>
        struct rebind
        {
            using other = NoInitAlloc<U>;
        };
        template <typename U, typename... A>
        void construct(U *p, A &&...a)
        {
            if constexpr (sizeof...(A) == 0)
            {
                ::new (static_cast<void *>(p)) U;
            }
            else
            {
                ::new (static_cast<void *>(p)) U(std::forward<A>(a)...);
            }
        }
    };

    struct HCapt
    {
        float r_min, r_max, r_point;
        float list_reach_sq = 0.F;            // (r_max + r_point)^2, the build's own float value
        int nlog2 = 0;
        uint32_t n_points = 0;
        std::vector<float> nodes;             // per internal node (Eytzinger) {split, bit pattern 1 = the high half inherited the low half}
        std::vector<uint32_t> leafbits;       // two bits per leaf: 1 = list beyond the representative, 2 = representative finite
        std::vector<uint32_t> leaf_of_point;  // per input point, until the environment has taken it over
        float top_lo[3], top_hi[3];
        // enumeration grid over the finite points (vmv_device.cuh: capt_collides_warp)
        std::vector<float> gpts;              // {x y z leaf} in Morton order of the finest cell
        std::vector<uint32_t> gstart;         // kCaptGridLevels start tables
        float g_origin[3] = {0.F, 0.F, 0.F}, g_inv0 = 1.F, g_cell0 = 1.F;
        int id;
        float2 *d_nodes = nullptr;
        uint32_t *d_leafbits = nullptr;
        float4 *d_gpts = nullptr;
        uint32_t *d_gstart = nullptr;
        // device build: the four tables live in one allocation that belongs to this cloud (not to a commit), and the host
        // vectors above stay empty until somebody needs them (vmv_env_broadcast, vmv_env_capt_digest)
        void *d_block = nullptr;
        uint32_t n_grid_points = 0;
        bool host_valid = true;
    };

    // Multi-level Voxel Table, host build (reference collision/mvt.hh:146-170, 437-446, 531-604)
    struct HMvt
    {
        float r_min, r_max, r_point;
        float ws_min[3], ws_max[3];
        float g_min[3], g_max[3];
        float inv_scale = 0.F;
        uint32_t grid_width = 0;
        std::vector<uint32_t> cells;   // grid_width^3, 0xffffffff = empty
        std::vector<float4> voxels;    // 2 per voxel
        std::vector<float4> points;    // grouped by voxel
        int id;
        uint32_t *d_cells = nullptr;
        float4 *d_voxels = nullptr;
        float4 *d_points = nullptr;
    };

    inline int mvt_cell_host(float v)
    {
        return v <= 0.F ? 0 : (v >= 65535.F ? 65535 : static_cast<int>(v));
    }

    int mvt_build(HMvt &t, const float *pts, size_t n, float r_min, float r_max, const float *ws_min, const float *ws_max, float r_point)
    {
        t.r_min = r_min, t.r_max = r_max, t.r_point = r_point;
        for (int k = 0; k < 3; ++k)
        {
            t.ws_min[k] = ws_min[k], t.ws_max[k] = ws_max[k];
            // initialize_empty_bounds (mvt.hh:428-435): an empty table rejects every query
            t.g_min[k] = std::numeric_limits<float>::max();
            t.g_max[k] = std::numeric_limits<float>::lowest();
        }
        if (n == 0)
        {
            return VMV_OK;
        }
        const float width = ws_max[0] - ws_min[0];
        // configure_grid (mvt.hh:437-446)
        const float cells_f = std::floor(width / r_max);
        if (!(cells_f >= 1.F) || !std::isfinite(cells_f))
        {
            return -1;
        }
        const uint32_t gw = static_cast<uint32_t>(std::min(cells_f, 65535.F));
        if (static_cast<size_t>(gw) * gw * gw > (size_t(1) << 26))
        {
            return -2;
        }
        t.grid_width = gw;
        t.inv_scale = static_cast<float>(gw) / width;
        t.cells.assign(static_cast<size_t>(gw) * gw * gw, 0xffffffffu);
        std::vector<uint32_t> voxel_of(n), count;
        const float top = static_cast<float>(gw - 1);
        for (size_t i = 0; i < n; ++i)
        {
            // build_spatial_grid (mvt.hh:531-588): clamp into the grid, voxels numbered by first insertion
            const float *p = pts + 3 * i;
            const uint32_t vx = mvt_cell_host(std::min(std::max((p[0] - ws_min[0]) * t.inv_scale, 0.F), top));
            const uint32_t vy = mvt_cell_host(std::min(std::max((p[1] - ws_min[1]) * t.inv_scale, 0.F), top));
            const uint32_t vz = mvt_cell_host(std::min(std::max((p[2] - ws_min[2]) * t.inv_scale, 0.F), top));
            uint32_t &cell = t.cells[(static_cast<size_t>(vx) * gw + vy) * gw + vz];
            if (cell == 0xffffffffu)
            {
                cell = static_cast<uint32_t>(count.size());
                count.push_back(0);
            }
            voxel_of[i] = cell;
            count[cell]++;
        }
        const size_t nv = count.size();
        std::vector<uint32_t> start(nv + 1, 0), fill(nv, 0);
        for (size_t v = 0; v < nv; ++v)
        {
            start[v + 1] = start[v] + count[v];
        }
        std::vector<float> bb(6 * nv);
        t.points.assign(n, make_float4(0, 0, 0, 0));
        for (size_t i = 0; i < n; ++i)
        {
            const uint32_t v = voxel_of[i];
            const float *p = pts + 3 * i;
            float *b = bb.data() + 6 * v;
            if (fill[v] == 0)
            {
                b[0] = b[3] = p[0], b[1] = b[4] = p[1], b[2] = b[5] = p[2];
            }
            else
            {
                for (int k = 0; k < 3; ++k)
                {
                    b[k] = std::min(b[k], p[k]);
                    b[3 + k] = std::max(b[3 + k], p[k]);
                }
            }
            t.points[start[v] + fill[v]++] = make_float4(p[0], p[1], p[2], 0.F);
        }
        t.voxels.resize(2 * nv);
        for (size_t v = 0; v < nv; ++v)
        {
            const float *b = bb.data() + 6 * v;
            float s, c;
            std::memcpy(&s, &start[v], 4);
            std::memcpy(&c, &count[v], 4);
            t.voxels[2 * v] = make_float4(b[0], b[1], b[2], b[3]);
            t.voxels[2 * v + 1] = make_float4(b[4], b[5], s, c);
            for (int k = 0; k < 3; ++k)
            {
                // compute_global_bounds (mvt.hh:590-604)
                t.g_min[k] = std::min(t.g_min[k], b[k]);
                t.g_max[k] = std::max(t.g_max[k], b[3 + k]);
            }
        }
        return VMV_OK;
    }

    float finite_or_neg_inf(float v)
    {
        // a degenerate capsule gives NaN; the reference then never breaks out of the sweep at it
        return std::isnan(v) ? -std::numeric_limits<float>::infinity() : v;
    }

    // -- CAPT build -------------------------------------------------------------------------
    // The k-d tree of the reference (capt.hh:106-119: median splits on x, y, z in turn over 2^nlog2 points, the input
    // padded with +inf) and, per node, the ONE fact about the reference's affordance lists that is not a function of
    // the split planes: whether the high half was handed the low half's points (capt.hh:232-246 scans the sorted half
    // from its first element while the predicate holds, so it is all or nothing).  The lists themselves are never
    // built: vmv::capt_member evaluates membership per point at query time.
    struct Vol
    {
        float lo[3], hi[3];
    };

    struct CaptTask
    {
        uint32_t begin, count, i;
        Vol vol;
        int d;
    };

    constexpr int kCaptTaskDepth = 6;  // 64 subtrees

    struct CaptBuilder
    {
        HCapt &t;
        const std::vector<float> &pts;  // padded to 2^nlog2 points with +inf
        std::vector<uint32_t> &argsort;
        float min_l2;
        std::vector<CaptTask> *tasks;        // non-null: collect subtrees at kCaptTaskDepth
        std::vector<uint8_t> &leaf_flag;     // per leaf, bit 0: full list, bit 1: finite

        void subdivide(uint32_t begin, uint32_t count, uint32_t i, Vol vol, int d, int depth)
        {
            if (tasks != nullptr && depth == kCaptTaskDepth && count > 1)
            {
                tasks->push_back(CaptTask{begin, count, i, vol, d});
                return;
            }
            if (count == 1)
            {
                const float *rep = &pts[3 * argsort[begin]];
                if (std::isfinite(rep[0]))
                {
                    // leaves are numbered left to right, as the sorted range is
                    t.leaf_of_point[argsort[begin]] = begin;
                    // cell entirely inside the smallest query ball around its representative:
                    // the representative alone decides (capt.hh:39-46,150)
                    const float d0 = std::max(rep[0] - vol.lo[0], vol.hi[0] - rep[0]);
                    const float d1 = std::max(rep[1] - vol.lo[1], vol.hi[1] - rep[1]);
                    const float d2 = std::max(rep[2] - vol.lo[2], vol.hi[2] - rep[2]);
                    const bool rep_only = (d0 * d0 + d1 * d1 + d2 * d2) <= min_l2;
                    leaf_flag[begin] = static_cast<uint8_t>(2 | (rep_only ? 0 : 1));
                }
                return;
            }

            // median split on axis d (capt.hh:106-119)
            std::sort(
                argsort.begin() + begin,
                argsort.begin() + begin + count,
                [&](uint32_t a, uint32_t b) { return pts[3 * a + d] < pts[3 * b + d]; });
            const uint32_t half = count / 2, middle = begin + half;
            const float test = static_cast<float>((pts[3 * argsort[middle - 1] + d] + pts[3 * argsort[middle] + d]) / 2.0);
            // the high half inherits the low half iff the low half's smallest element is within r_max of the plane
            const float first = pts[3 * argsort[begin] + d];
            const uint32_t inherited = (first >= test - t.r_max && std::isfinite(first)) ? 1u : 0u;
            float inherited_f;
            std::memcpy(&inherited_f, &inherited, 4);
            t.nodes[2 * i] = test;
            t.nodes[2 * i + 1] = inherited_f;

            Vol lo_vol = vol, hi_vol = vol;
            lo_vol.hi[d] = test;
            hi_vol.lo[d] = test;
            const int nd = (d + 1) % 3;
            subdivide(begin, half, 2 * i + 1, lo_vol, nd, depth + 1);
            subdivide(begin + half, half, 2 * i + 2, hi_vol, nd, depth + 1);
        }
    };

    // The box of the points with a finite x (the reference's top-level reject, capt.hh:431-438), the fully finite points, and the
    // frame of the enumeration grid over them (origin, finest cell edge)
    void capt_frame(HCapt &t, const float *points, size_t n, std::vector<uint32_t> &finite)
    {
        const float inf = std::numeric_limits<float>::infinity();
        for (int k = 0; k < 3; ++k)
        {
            t.top_lo[k] = inf;
            t.top_hi[k] = -inf;
        }
        finite.clear();
        for (uint32_t k = 0; k < n; ++k)
        {
            const float *p = points + 3 * size_t(k);
            // (the tree's leaf step looks at x alone, capt.hh:150; a point with a non-finite y or z can never be within a
            // finite distance of anything, so it is left out of the grid as well)
            if (std::isfinite(p[0]))
            {
                for (int c = 0; c < 3; ++c)
                {
                    t.top_lo[c] = std::min(t.top_lo[c], p[c]);
                    t.top_hi[c] = std::max(t.top_hi[c], p[c]);
                }
                if (std::isfinite(p[1]) && std::isfinite(p[2]))
                {
                    finite.push_back(k);
                }
            }
        }
        constexpr int kDim = 1 << vmv::kCaptGridBits;
        float ext = 0.F;
        float glo[3] = {0.F, 0.F, 0.F};
        for (const uint32_t k : finite)
        {
            for (int c = 0; c < 3; ++c)
            {
                glo[c] = (k == finite[0]) ? points[3 * size_t(k) + c] : std::min(glo[c], points[3 * size_t(k) + c]);
            }
        }
        for (const uint32_t k : finite)
        {
            for (int c = 0; c < 3; ++c)
            {
                ext = std::max(ext, points[3 * size_t(k) + c] - glo[c]);
            }
        }
        t.g_cell0 = std::max(0.04F, ext / (kDim - 0.5F));
        t.g_inv0 = 1.F / t.g_cell0;
        for (int c = 0; c < 3; ++c)
        {
            t.g_origin[c] = glo[c];
        }
    }

    void capt_build(HCapt &t, const float *points, size_t n, float r_min, float r_max, float r_point)
    {
        t.r_min = r_min, t.r_max = r_max, t.r_point = r_point;
        const float max_l1 = r_max + r_point;
        t.list_reach_sq = max_l1 * max_l1;
        t.n_points = static_cast<uint32_t>(n);
        t.nlog2 = 0;
        while ((size_t(1) << t.nlog2) < n)
        {
            t.nlog2++;
        }
        const size_t pow2 = size_t(1) << t.nlog2;
        const float inf = std::numeric_limits<float>::infinity();
        std::vector<float> pts(points, points + 3 * n);
        pts.resize(3 * pow2, inf);
        t.nodes.assign(2 * (pow2 - 1), 0.F);
        t.leaf_of_point.assign(pow2, 0u);
        std::vector<uint8_t> leaf_flag(pow2, 0);
        std::vector<uint32_t> argsort(pow2);
        std::iota(argsort.begin(), argsort.end(), 0u);
        const float min_l2 = (r_min + r_point) * (r_min + r_point);
        // top levels here (a tree shallower than kCaptTaskDepth is done entirely by this call)
        std::vector<CaptTask> tasks;
        const auto tm0 = std::chrono::steady_clock::now();
        CaptBuilder top{t, pts, argsort, min_l2, &tasks, leaf_flag};
        top.subdivide(0, static_cast<uint32_t>(pow2), 0, Vol{{-inf, -inf, -inf}, {inf, inf, inf}}, 0, 0);
        {
            // the subtrees: disjoint ranges of `argsort`, `nodes`, `leaf_flag`
            std::atomic<size_t> next{0};
            std::atomic<bool> failed{false};  // an exception must not leave a worker thread (std::terminate)
            auto work = [&]()
            {
                try
                {
                    for (size_t k = next.fetch_add(1); k < tasks.size() && !failed.load(); k = next.fetch_add(1))
                    {
                        CaptBuilder b{t, pts, argsort, min_l2, nullptr, leaf_flag};
                        const CaptTask &task = tasks[k];
                        b.subdivide(task.begin, task.count, task.i, task.vol, task.d, kCaptTaskDepth);
                    }
                }
                catch (...)
                {
                    failed.store(true);
                }
            };
            const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
            const size_t n_threads = std::min<size_t>({tasks.size(), hw, 32});
            std::vector<std::thread> pool;
            for (size_t k = 1; k < n_threads; ++k)
            {
                pool.emplace_back(work);
            }
            work();
            for (auto &th : pool)
            {
                th.join();
            }
            if (failed.load())
            {
                throw std::bad_alloc();  // becomes an error code at the ABI (guarded)
            }
        }
        t.leaf_of_point.resize(n);
        t.leafbits.assign((pow2 + 15) / 16, 0u);
        for (size_t k = 0; k < pow2; ++k)
        {
            t.leafbits[k >> 4] |= static_cast<uint32_t>(leaf_flag[k]) << (2 * (k & 15));
        }

        std::vector<uint32_t> finite;
        capt_frame(t, points, n, finite);
        constexpr int kDim = 1 << vmv::kCaptGridBits;
        std::vector<std::pair<uint32_t, uint32_t>> keys(finite.size());
        for (size_t k = 0; k < finite.size(); ++k)
        {
            uint32_t cell[3];
            for (int c = 0; c < 3; ++c)
            {
                // the device derives cell ranges from the same expression (vmv_device.cuh)
                const float a = (points[3 * size_t(finite[k]) + c] - t.g_origin[c]) * t.g_inv0;
                cell[c] = static_cast<uint32_t>(std::max(0, std::min(kDim - 1, static_cast<int>(std::floor(a)))));
            }
            keys[k] = {vmv::morton_spread(cell[0]) | (vmv::morton_spread(cell[1]) << 1) | (vmv::morton_spread(cell[2]) << 2), finite[k]};
        }
        std::sort(keys.begin(), keys.end());
        t.gpts.resize(4 * keys.size());
        for (size_t k = 0; k < keys.size(); ++k)
        {
            const uint32_t id = keys[k].second;
            float leaf;
            std::memcpy(&leaf, &t.leaf_of_point[id], 4);
            t.gpts[4 * k] = points[3 * size_t(id)], t.gpts[4 * k + 1] = points[3 * size_t(id) + 1], t.gpts[4 * k + 2] = points[3 * size_t(id) + 2];
            t.gpts[4 * k + 3] = leaf;
        }
        t.gstart.assign(vmv::capt_grid_offset(vmv::kCaptGridLevels), 0u);
        for (int l = 0; l < vmv::kCaptGridLevels; ++l)
        {
            uint32_t *st = t.gstart.data() + vmv::capt_grid_offset(l);
            const uint32_t cells = 1u << (3 * (vmv::kCaptGridBits - l));
            size_t at = 0;
            for (uint32_t m = 0; m <= cells; ++m)
            {
                while (at < keys.size() && (keys[at].first >> (3 * l)) < m)
                {
                    ++at;
                }
                st[m] = static_cast<uint32_t>(at);
            }
        }
        if (std::getenv("VMV_CAPT_TIMING"))
        {
            std::fprintf(stderr, "capt_build: %zu points (%zu finite), %zu subtrees, %.3f s\n", n, finite.size(), tasks.size(),
                         std::chrono::duration<double>(std::chrono::steady_clock::now() - tm0).count());
        }
    }

    cudaError_t cuMemsetD32Async_or_kernel(float *dst, uint32_t pattern, size_t count, cudaStream_t st)
    {
        if (count == 0)
        {
            return cudaSuccess;
        }
        vmv::k_fill_u32<<<static_cast<unsigned>((count + 255) / 256), 256, 0, st>>>(reinterpret_cast<uint32_t *>(dst), pattern, count);
        return cudaGetLastError();
    }

    // the host-side copy of a device-built cloud's tables, on demand
    int capt_ensure_host(const HCapt &ct)
    {
        HCapt &t = const_cast<HCapt &>(ct);
        if (t.host_valid || t.d_block == nullptr)
        {
            return VMV_OK;
        }
        const size_t pow2 = size_t(1) << t.nlog2;
        t.nodes.assign(2 * (pow2 - 1), 0.F);
        t.leafbits.assign((pow2 + 15) / 16, 0u);
        t.gpts.assign(4 * size_t(t.n_grid_points), 0.F);
        t.gstart.assign(vmv::capt_grid_offset(vmv::kCaptGridLevels), 0u);
        if (!t.nodes.empty())
        {
            VMV_CUDA(cudaMemcpy(t.nodes.data(), t.d_nodes, t.nodes.size() * sizeof(float), cudaMemcpyDeviceToHost));
        }
        VMV_CUDA(cudaMemcpy(t.leafbits.data(), t.d_leafbits, t.leafbits.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (!t.gpts.empty())
        {
            VMV_CUDA(cudaMemcpy(t.gpts.data(), t.d_gpts, t.gpts.size() * sizeof(float), cudaMemcpyDeviceToHost));
        }
        VMV_CUDA(cudaMemcpy(t.gstart.data(), t.d_gstart, t.gstart.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        t.host_valid = true;
        return VMV_OK;
    }

    // The same tables built on the device (vmv_capt_build.cuh): one segmented bitonic sort per level of the tree, one more for
    // the Morton grid.  The results are copied back into the host-side description (the environment is re-packed from
    // it at every commit and serialised from it by vmv_env_broadcast); 10^5 points: ~3 ms instead of ~55 ms.
    // Points that are NaN are left to the host build (the sort's comparison needs a total order).
    int capt_build_device(HCapt &t, const float *points, size_t n, float r_min, float r_max, float r_point)
    {
        const auto tm0 = std::chrono::steady_clock::now();
        t.r_min = r_min, t.r_max = r_max, t.r_point = r_point;
        const float max_l1 = r_max + r_point;
        t.list_reach_sq = max_l1 * max_l1;
        t.n_points = static_cast<uint32_t>(n);
        t.nlog2 = 0;
        while ((size_t(1) << t.nlog2) < n)
        {
            t.nlog2++;
        }
        const uint32_t pow2 = 1u << t.nlog2;
        const uint32_t n_pad = std::max(pow2, vmv::kSortTile);
        const float inf = std::numeric_limits<float>::infinity();
        const float min_l2 = (r_min + r_point) * (r_min + r_point);
        std::vector<uint32_t> finite;
        capt_frame(t, points, n, finite);
        const uint32_t m = static_cast<uint32_t>(finite.size());

        // two allocations: the tables (they stay with the cloud) and the scratch of this build
        const uint32_t n_starts = vmv::capt_grid_offset(vmv::kCaptGridLevels);
        auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
        const size_t b_nodes = up(size_t(pow2) * sizeof(float2)), b_bits = up(size_t((pow2 + 15) / 16) * 4), b_gpts = up(size_t(std::max(m, 1u)) * sizeof(float4)),
                     b_starts = up(size_t(n_starts) * 4);
        unsigned char *block = nullptr;
        VMV_CUDA(cudaMalloc(reinterpret_cast<void **>(&block), b_nodes + b_bits + b_gpts + b_starts));
        t.d_block = block;
        float2 *d_nodes = reinterpret_cast<float2 *>(block);
        uint32_t *d_leafbits = reinterpret_cast<uint32_t *>(block + b_nodes);
        float4 *d_gpts = reinterpret_cast<float4 *>(block + b_nodes + b_bits);
        uint32_t *d_gstart = reinterpret_cast<uint32_t *>(block + b_nodes + b_bits + b_gpts);
        // the scratch of a build is kept per device between builds (grow-only): clouds arrive per frame, and a cudaMalloc /
        // cudaFree pair costs more than the 161 launches of the build
        static std::mutex scratch_mutex;
        static unsigned char *scratch_mem[kMaxDevices] = {};
        static size_t scratch_cap[kMaxDevices] = {};
        std::lock_guard<std::mutex> scratch_lock(scratch_mutex);
        int device = 0;
        VMV_CUDA(cudaGetDevice(&device));
        device %= kMaxDevices;
        const size_t b_pts = up(3 * size_t(n_pad) * 4), b_arr = up(size_t(n_pad) * 4);
        if (scratch_cap[device] < b_pts + 5 * b_arr)
        {
            if (scratch_mem[device] != nullptr)
            {
                cudaFree(scratch_mem[device]);
                scratch_mem[device] = nullptr;
                scratch_cap[device] = 0;
            }
            VMV_CUDA(cudaMalloc(reinterpret_cast<void **>(&scratch_mem[device]), b_pts + 5 * b_arr));
            scratch_cap[device] = b_pts + 5 * b_arr;
        }
        unsigned char *tmp = scratch_mem[device];
        float *d_pts = reinterpret_cast<float *>(tmp), *d_key = reinterpret_cast<float *>(tmp + b_pts);
        uint32_t *d_idx = reinterpret_cast<uint32_t *>(tmp + b_pts + b_arr), *d_leaf_of = reinterpret_cast<uint32_t *>(tmp + b_pts + 2 * b_arr),
                 *d_code = reinterpret_cast<uint32_t *>(tmp + b_pts + 3 * b_arr), *d_order = reinterpret_cast<uint32_t *>(tmp + b_pts + 4 * b_arr);
        cudaStream_t st = nullptr;
        if (n > 0)
        {
            VMV_CUDA(cudaMemcpyAsync(d_pts, points, 3 * n * sizeof(float), cudaMemcpyHostToDevice, st));
        }
        if (n_pad > n)
        {
            // +inf is 0x7f800000: the padding of the reference's tree (capt.hh:318-322)
            VMV_CUDA(cuMemsetD32Async_or_kernel(d_pts + 3 * n, 0x7f800000u, 3 * (size_t(n_pad) - n), st));
        }
        const auto tm1 = std::chrono::steady_clock::now();
        uint64_t launches = 0;
        const unsigned gb = (n_pad + 255) / 256;
        // (the codes kernel also writes the identity permutation: used here for the tree's order, below for the grid's)
        vmv::k_capt_codes<<<gb, 256, 0, st>>>(d_pts, static_cast<uint32_t>(n), n_pad, t.g_origin[0], t.g_origin[1], t.g_origin[2], t.g_inv0, d_code, d_idx);
        ++launches;
        for (int level = 0; level < t.nlog2; ++level)
        {
            const uint32_t seg = pow2 >> level;
            vmv::k_capt_gather_axis<<<gb, 256, 0, st>>>(d_pts, d_idx, n_pad, level % 3, d_key);
            VMV_CUDA(vmv::segmented_sort<float>(d_key, d_idx, n_pad, seg, st, launches));
            const uint32_t count = 1u << level;
            vmv::k_capt_level_nodes<<<(count + 127) / 128, 128, 0, st>>>(d_key, seg, count, count - 1u, r_max, d_nodes);
            launches += 2;
        }
        vmv::k_capt_leaves<<<((pow2 + 15) / 16 + 127) / 128, 128, 0, st>>>(d_pts, d_idx, d_nodes, static_cast<uint32_t>(t.nlog2), pow2, min_l2, d_leafbits, d_leaf_of);
        ++launches;
        // the Morton grid: the points by (cell code, index)
        vmv::k_capt_codes<<<gb, 256, 0, st>>>(d_pts, static_cast<uint32_t>(n), n_pad, t.g_origin[0], t.g_origin[1], t.g_origin[2], t.g_inv0, d_code, d_order);
        ++launches;
        VMV_CUDA(vmv::segmented_sort<uint32_t>(d_code, d_order, n_pad, n_pad, st, launches));
        if (m > 0)
        {
            vmv::k_capt_grid_points<<<(m + 255) / 256, 256, 0, st>>>(d_pts, d_order, d_leaf_of, m, d_gpts);
            ++launches;
        }
        vmv::k_capt_grid_starts<<<(n_starts + 255) / 256, 256, 0, st>>>(d_code, m, d_gstart);
        ++launches;
        VMV_CUDA(cudaGetLastError());
        g_launches += launches;

        if (std::getenv("VMV_CAPT_TIMING"))
        {
            cudaStreamSynchronize(st);
        }
        const auto tm2 = std::chrono::steady_clock::now();
        t.d_nodes = d_nodes, t.d_leafbits = d_leafbits, t.d_gpts = d_gpts, t.d_gstart = d_gstart;
        t.n_grid_points = m;
        t.host_valid = false;
        t.nodes.clear(), t.leafbits.clear(), t.gpts.clear(), t.gstart.clear();
        t.leaf_of_point.resize(n);
        if (n > 0)
        {
            VMV_CUDA(cudaMemcpyAsync(t.leaf_of_point.data(), d_leaf_of, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        }
        VMV_CUDA(cudaStreamSynchronize(st));
        if (std::getenv("VMV_CAPT_TIMING"))
        {
            const auto tm3 = std::chrono::steady_clock::now();
            auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            std::fprintf(stderr, "capt_build_device: %zu points: frame + alloc + upload %.2f ms, %llu launches %.2f ms, download %.2f ms\n", n, ms(tm0, tm1),
                         static_cast<unsigned long long>(launches), ms(tm1, tm2), ms(tm2, tm3));
        }
        return VMV_OK;
    }
}  // namespace

struct vmv_env
{
    std::vector<HSphere> spheres;
    std::vector<HCapsule> capsules, z_capsules;
    std::vector<HCuboid> cuboids, z_cuboids;
    std::vector<HHeight> heightfields;
    std::vector<HCapt> capts;
    std::vector<HMvt> mvts;
    std::vector<float> cloud_xyz;  // every point of every pointcloud (nearest-point table, vmv_device.cuh)
    std::vector<uint32_t> cloud_tag;  // per point: cloud << 24 | CAPT leaf (cloud 0xff: no list shortcut for this point)
    float cloud_r_point_max = 0.F;
    float cloud_grid_ms = 0.F;
    bool has_attachment = false;
    float attach_tf[12];  // row-major 3x4
    std::vector<float> attach_spheres;
    int next_id = 0;

    bool committed = false;
    // shapes_dirty: a primitive, heightfield or pointcloud changed since the last commit (full re-pack); an
    // attach / detach alone only rewrites the attachment tail of the blob (a few hundred bytes) and keeps every
    // device array, voxel table and the nearest-point table -- attach / detach are O(1) in the reference too
    // (std::optional<Attachment>, collision/environment.hh:28)
    bool shapes_dirty = true;
    int device = -1;
    std::vector<uint32_t> blob;
    float *d_blob = nullptr;
    size_t d_blob_cap = 0;   // bytes
    uint32_t attach_off = 0; // word offset of the attachment tail in the blob
    int *d_object_ids = nullptr;  // packed object index -> insertion id (vmv_debug)
    std::vector<void *> owned;  // device allocations referenced from the blob

    // grid-culled path (vmv_grid.cuh, vmv_kernels_v4.cuh): rounded-box records of all primitives and, per robot,
    // the voxel table of candidate masks (built on first use with that robot)
    std::vector<float> uobjs;  // kObjRec floats per object, packed order
    float4 *d_uobjs = nullptr;
    struct GridCache
    {
        bool ready = false;
        bool usable = false;
        bool wide = false;  // 64-bit masks (more than 32 objects), else 32-bit
        vmv::GridDev dev{};
        void *mem = nullptr;
        float build_ms = 0.F;
        size_t bytes = 0;
    };
    mutable std::mutex grid_mutex;
    mutable GridCache grids[VMV_N_ROBOTS];
    // verdict of the robot's world-fixed links against this environment (any-environment kernels): 0 not checked yet,
    // 1 clear, 2 in collision (then every configuration is invalid)
    mutable int static_state[VMV_N_ROBOTS] = {0, 0, 0, 0};

    // the device-built tables of the pointclouds (they belong to the clouds, not to a commit)
    void release_clouds()
    {
        for (auto &t : capts)
        {
            if (t.d_block)
            {
                cudaFree(t.d_block);
                t.d_block = nullptr;
            }
        }
    }

    void release_device()
    {
        for (auto &g : grids)
        {
            if (g.mem)
            {
                cudaFree(g.mem);
            }
            g = GridCache{};
        }
        for (int &st : static_state)
        {
            st = 0;
        }
        d_uobjs = nullptr;  // freed through `owned`
        for (void *p : owned)
        {
            cudaFree(p);
        }
        owned.clear();
        if (d_blob)
        {
            cudaFree(d_blob);
            d_blob = nullptr;
        }
        d_blob_cap = 0;
        shapes_dirty = true;
        committed = false;
    }
};

namespace
{
    template <typename T>
    void sort_by_min_distance(std::vector<T> &v)
    {
        std::stable_sort(v.begin(), v.end(), [](const T &a, const T &b) { return a.min_d < b.min_d; });
    }

    uint32_t f2u(float f)
    {
        uint32_t u;
        std::memcpy(&u, &f, 4);
        return u;
    }

    template <typename T, typename A>
    int upload(vmv_env *env, const std::vector<T, A> &host, T *&dev)
    {
        void *p = nullptr;
        const size_t bytes = std::max<size_t>(16, host.size() * sizeof(T));
        VMV_CUDA(cudaMalloc(&p, bytes));
        env->owned.push_back(p);
        if (!host.empty())
        {
            VMV_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
        }
        dev = static_cast<T *>(p);
        return VMV_OK;
    }

    // Every primitive as a rounded box {centre, rho}{axis_i, half extent_i} (vmv_grid.cuh), in
    // the packed order spheres | capsules | z-capsules | cuboids | z-cuboids.
    void build_rounded_boxes(vmv_env *env)
    {
        std::vector<float> &U = env->uobjs;
        U.clear();
        auto rec = [&](const double c[3], double rho, const double a[3][3], const double h[3])
        {
            U.push_back(static_cast<float>(c[0])), U.push_back(static_cast<float>(c[1])), U.push_back(static_cast<float>(c[2]));
            U.push_back(static_cast<float>(rho));
            for (int i = 0; i < 3; ++i)
            {
                U.push_back(static_cast<float>(a[i][0])), U.push_back(static_cast<float>(a[i][1])), U.push_back(static_cast<float>(a[i][2]));
                U.push_back(static_cast<float>(h[i]));
            }
        };
        const double I[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        const double Z[3] = {0, 0, 0};
        for (const auto &s : env->spheres)
        {
            const double c[3] = {s.x, s.y, s.z};
            rec(c, s.r, I, Z);
        }
        auto capsule = [&](const HCapsule &k)
        {
            const double v[3] = {k.xv, k.yv, k.zv};
            const double len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            if (!(len > 1e-12) || !std::isfinite(len))
            {
                // degenerate segment: the clamped projection of the reference stays at the first end point
                const double c[3] = {k.x1, k.y1, k.z1};
                rec(c, k.r, I, Z);
                return;
            }
            double a[3][3];
            for (int j = 0; j < 3; ++j)
            {
                a[0][j] = v[j] / len;
            }
            // any orthonormal complement of a[0]
            int m = 0;
            for (int j = 1; j < 3; ++j)
            {
                if (std::fabs(a[0][j]) < std::fabs(a[0][m]))
                {
                    m = j;
                }
            }
            double e[3] = {0, 0, 0};
            e[m] = 1.0;
            double d = a[0][m];
            double n2 = 0;
            for (int j = 0; j < 3; ++j)
            {
                a[1][j] = e[j] - d * a[0][j];
                n2 += a[1][j] * a[1][j];
            }
            n2 = std::sqrt(n2);
            for (int j = 0; j < 3; ++j)
            {
                a[1][j] /= n2;
            }
            a[2][0] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
            a[2][1] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
            a[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
            const double c[3] = {k.x1 + 0.5 * v[0], k.y1 + 0.5 * v[1], k.z1 + 0.5 * v[2]};
            const double h[3] = {0.5 * len, 0, 0};
            rec(c, k.r, a, h);
        };
        for (const auto &k : env->capsules)
        {
            capsule(k);
        }
        for (const auto &k : env->z_capsules)
        {
            capsule(k);
        }
        auto cuboid = [&](const HCuboid &b)
        {
            const float *f = b.f;
            const double c[3] = {f[0], f[1], f[2]};
            const double a[3][3] = {{f[3], f[4], f[5]}, {f[6], f[7], f[8]}, {f[9], f[10], f[11]}};
            const double h[3] = {f[12], f[13], f[14]};
            rec(c, 0.0, a, h);
        };
        for (const auto &b : env->cuboids)
        {
            cuboid(b);
        }
        for (const auto &b : env->z_cuboids)
        {
            // the z-aligned test reads only the xy parts of axes 1, 2 and takes axis 3 = z
            // (collision/sphere_cuboid.hh:35-52)
            HCuboid zb = b;
            zb.f[5] = 0.F, zb.f[8] = 0.F;
            zb.f[9] = 0.F, zb.f[10] = 0.F, zb.f[11] = 1.F;
            cuboid(zb);
        }
    }

    // Clearance grid of all pointcloud points (vmv_device.cuh: cloud_clearance).  The table covers the
    // points' bounding box grown by kCloudReach; outside it every cloud point is farther than that.
    constexpr float kCloudReach = 0.5F;
#ifndef VMV_CLOUD_MAX_VOXELS_LOG2
#define VMV_CLOUD_MAX_VOXELS_LOG2 24
#endif
    constexpr size_t kCloudMaxVoxels = size_t(1) << VMV_CLOUD_MAX_VOXELS_LOG2;

    int build_cloud_grid(vmv_env *env, vmv::CloudGridRec &g)
    {
        const auto tm_grid0 = std::chrono::steady_clock::now();
        const size_t n = env->cloud_xyz.size() / 3;
        float lo[3] = {3e38F, 3e38F, 3e38F}, hi[3] = {-3e38F, -3e38F, -3e38F};
        std::vector<float4> pts(n);
        for (size_t i = 0; i < n; ++i)
        {
            const float *p = env->cloud_xyz.data() + 3 * i;
            if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2]))
            {
                return VMV_OK;  // no grid: every query runs
            }
            for (int k = 0; k < 3; ++k)
            {
                lo[k] = std::min(lo[k], p[k]);
                hi[k] = std::max(hi[k], p[k]);
            }
            float tag;
            std::memcpy(&tag, &env->cloud_tag[i], 4);
            pts[i] = make_float4(p[0], p[1], p[2], tag);
        }
#ifndef VMV_CLOUD_VOXEL
#define VMV_CLOUD_VOXEL 0.0125
#endif
        double h = VMV_CLOUD_VOXEL;
        int dim[3];
        while (true)
        {
            size_t vox = 1;
            for (int k = 0; k < 3; ++k)
            {
                dim[k] = std::max(1, static_cast<int>(std::ceil((hi[k] - lo[k] + 2.0 * kCloudReach) / h)));
                vox *= static_cast<size_t>(dim[k]);
            }
            if (vox <= kCloudMaxVoxels)
            {
                break;
            }
            h *= 1.15;
        }
        const size_t n_vox = static_cast<size_t>(dim[0]) * dim[1] * dim[2];
        {
            // Morton order over the points' bounding box (10 bits per axis), then tiles of kCloudTile points
            // with their boxes: the build kernel skips tiles that cannot improve any voxel of a brick
            auto spread = [](uint32_t v)
            {
                v &= 0x3ffu;
                v = (v | (v << 16)) & 0x030000ffu;
                v = (v | (v << 8)) & 0x0300f00fu;
                v = (v | (v << 4)) & 0x030c30c3u;
                v = (v | (v << 2)) & 0x09249249u;
                return v;
            };
            std::vector<std::pair<uint32_t, uint32_t>> keys(n);
            for (size_t i = 0; i < n; ++i)
            {
                uint32_t q[3];
                for (int k = 0; k < 3; ++k)
                {
                    const float ext = std::max(hi[k] - lo[k], 1e-9F);
                    q[k] = static_cast<uint32_t>(std::min(1023.F, std::max(0.F, (pts[i].x * (k == 0) + pts[i].y * (k == 1) + pts[i].z * (k == 2) - lo[k]) / ext * 1023.F)));
                }
                keys[i] = {spread(q[0]) | (spread(q[1]) << 1) | (spread(q[2]) << 2), static_cast<uint32_t>(i)};
            }
            std::sort(keys.begin(), keys.end());
            std::vector<float4> sorted(n);
            for (size_t i = 0; i < n; ++i)
            {
                sorted[i] = pts[keys[i].second];
            }
            pts.swap(sorted);
        }
        const size_t n_tiles = (n + vmv::kCloudTile - 1) / vmv::kCloudTile;
        std::vector<float4> boxes(2 * std::max<size_t>(n_tiles, 1));
        for (size_t k = 0; k < n_tiles; ++k)
        {
            float blo[3] = {3e38F, 3e38F, 3e38F}, bhi[3] = {-3e38F, -3e38F, -3e38F};
            for (size_t i = k * vmv::kCloudTile; i < std::min(n, (k + 1) * vmv::kCloudTile); ++i)
            {
                const float c[3] = {pts[i].x, pts[i].y, pts[i].z};
                for (int a = 0; a < 3; ++a)
                {
                    blo[a] = std::min(blo[a], c[a]);
                    bhi[a] = std::max(bhi[a], c[a]);
                }
            }
            boxes[2 * k] = make_float4(blo[0], blo[1], blo[2], bhi[0]);
            boxes[2 * k + 1] = make_float4(bhi[1], bhi[2], 0.F, 0.F);
        }
        float4 *d_pts = nullptr, *d_boxes = nullptr;
        float4 *d_cells = nullptr;
        int rc = upload(env, pts, d_pts);
        if (rc == VMV_OK)
        {
            rc = upload(env, boxes, d_boxes);
        }
        if (rc != VMV_OK)
        {
            return rc;
        }
        {
            void *p = nullptr;
            VMV_CUDA(cudaMalloc(&p, n_vox * sizeof(float4)));
            env->owned.push_back(p);
            d_cells = static_cast<float4 *>(p);
        }
        cudaEvent_t e0, e1;
        VMV_CUDA(cudaEventCreate(&e0));
        VMV_CUDA(cudaEventCreate(&e1));
        VMV_CUDA(cudaEventRecord(e0, nullptr));
        {
            const size_t bricks = static_cast<size_t>((dim[0] + 7) / 8) * ((dim[1] + 7) / 8) * ((dim[2] + 3) / 4);
            vmv::k_build_cloud_grid<<<static_cast<unsigned>(bricks), 256>>>(
                d_pts, static_cast<uint32_t>(n), d_boxes, lo[0] - kCloudReach, lo[1] - kCloudReach, lo[2] - kCloudReach, static_cast<float>(h), dim[0],
                dim[1], dim[2], d_cells);
        }
        g_launches++;
        VMV_CUDA(cudaGetLastError());
        VMV_CUDA(cudaEventRecord(e1, nullptr));
        VMV_CUDA(cudaEventSynchronize(e1));
        VMV_CUDA(cudaEventElapsedTime(&env->cloud_grid_ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        if (std::getenv("VMV_CAPT_TIMING"))
        {
            std::fprintf(stderr, "build_cloud_grid: %zu points, %zu voxels (h = %.4f m): kernel %.2f ms, host side %.2f ms\n", n, n_vox, h, env->cloud_grid_ms,
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tm_grid0).count() - env->cloud_grid_ms);
        }
        g.x0 = lo[0] - kCloudReach, g.y0 = lo[1] - kCloudReach, g.z0 = lo[2] - kCloudReach;
        g.inv_h = static_cast<float>(1.0 / h);
        g.nx = dim[0], g.ny = dim[1], g.nz = dim[2];
        g.outside = kCloudReach - static_cast<float>(h) - 1e-3F;
        g.h = static_cast<float>(h);
        g.r_point_max = env->cloud_r_point_max;
        g.cells = d_cells;
        return VMV_OK;
    }

    // the attachment tail of the blob: {n 0 0 0} then n x {x y z r}; everything before attach_off stays
    void append_attachment(vmv_env *env)
    {
        std::vector<uint32_t> &B = env->blob;
        B.resize(env->attach_off);
        const uint32_t n_attach = env->has_attachment ? static_cast<uint32_t>(env->attach_spheres.size() / 4) : 0u;
        B.push_back(n_attach), B.push_back(0), B.push_back(0), B.push_back(0);
        if (env->has_attachment)
        {
            for (float v : env->attach_spheres)
            {
                B.push_back(f2u(v));
            }
        }
        while (B.size() % 4)
        {
            B.push_back(0u);
        }
        vmv::EnvHeader H;
        std::memcpy(&H, B.data(), sizeof(H));
        H.n_attach = n_attach;
        std::memcpy(B.data(), &H, sizeof(H));
    }

    int pack_and_upload(vmv_env *env)
    {
        env->release_device();
        VMV_CUDA(cudaGetDevice(&env->device));

        // device-side arrays first (their addresses go into the blob)
        for (auto &h : env->heightfields)
        {
            // one padding row + one element: the reference's clamp admits index == xd*yd + xd
            std::vector<float> padded(h.data);
            padded.resize(h.xd * (h.yd + 1) + h.xd + 1, 0.F);
            int rc = upload(env, padded, h.d_data);
            if (rc != VMV_OK)
            {
                return rc;
            }
        }
        for (auto &t : env->capts)
        {
            if (t.d_block != nullptr)
            {
                continue;  // built on the device: the tables are where the build left them
            }
            float *dn = nullptr, *dp = nullptr;
            int rc = upload(env, t.nodes, dn);
            if (rc == VMV_OK)
            {
                rc = upload(env, t.leafbits, t.d_leafbits);
            }
            if (rc == VMV_OK)
            {
                rc = upload(env, t.gpts, dp);
            }
            if (rc == VMV_OK)
            {
                rc = upload(env, t.gstart, t.d_gstart);
            }
            if (rc != VMV_OK)
            {
                return rc;
            }
            t.d_nodes = reinterpret_cast<float2 *>(dn);
            t.d_gpts = reinterpret_cast<float4 *>(dp);
        }
        for (auto &t : env->mvts)
        {
            int rc = upload(env, t.cells, t.d_cells);
            if (rc == VMV_OK)
            {
                rc = upload(env, t.voxels, t.d_voxels);
            }
            if (rc == VMV_OK)
            {
                rc = upload(env, t.points, t.d_points);
            }
            if (rc != VMV_OK)
            {
                return rc;
            }
        }

        std::vector<uint32_t> &B = env->blob;
        B.assign(sizeof(vmv::EnvHeader) / 4, 0u);
        auto align4 = [&]()
        {
            while (B.size() % 4)
            {
                B.push_back(0u);
            }
        };
        auto pf = [&](float v) { B.push_back(f2u(v)); };
        vmv::EnvHeader H{};

        H.n_spheres = static_cast<uint32_t>(env->spheres.size());
        H.off_spheres = static_cast<uint32_t>(B.size());
        for (const auto &s : env->spheres)
        {
            pf(s.x), pf(s.y), pf(s.z), pf(s.r), pf(s.min_d), pf(0), pf(0), pf(0);
        }
        H.n_capsules = static_cast<uint32_t>(env->capsules.size());
        H.off_capsules = static_cast<uint32_t>(B.size());
        for (const auto &c : env->capsules)
        {
            pf(c.x1), pf(c.y1), pf(c.z1), pf(c.r), pf(c.xv), pf(c.yv), pf(c.zv), pf(c.rdv), pf(c.min_d), pf(0), pf(0), pf(0);
        }
        H.n_zcapsules = static_cast<uint32_t>(env->z_capsules.size());
        H.off_zcapsules = static_cast<uint32_t>(B.size());
        for (const auto &c : env->z_capsules)
        {
            pf(c.x1), pf(c.y1), pf(c.z1), pf(c.r), pf(c.zv), pf(c.rdv), pf(c.min_d), pf(0);
        }
        H.n_cuboids = static_cast<uint32_t>(env->cuboids.size());
        H.off_cuboids = static_cast<uint32_t>(B.size());
        for (const auto &c : env->cuboids)
        {
            const float *f = c.f;
            pf(f[0]), pf(f[1]), pf(f[2]), pf(f[12]);
            pf(f[3]), pf(f[4]), pf(f[5]), pf(f[13]);
            pf(f[6]), pf(f[7]), pf(f[8]), pf(f[14]);
            pf(f[9]), pf(f[10]), pf(f[11]), pf(c.min_d);
        }
        H.n_zcuboids = static_cast<uint32_t>(env->z_cuboids.size());
        H.off_zcuboids = static_cast<uint32_t>(B.size());
        for (const auto &c : env->z_cuboids)
        {
            const float *f = c.f;
            pf(f[0]), pf(f[1]), pf(f[2]), pf(c.min_d);
            pf(f[3]), pf(f[4]), pf(f[6]), pf(f[7]);
            pf(f[12]), pf(f[13]), pf(f[14]), pf(0);
        }
        H.n_heightfields = static_cast<uint32_t>(env->heightfields.size());
        H.off_heightfields = static_cast<uint32_t>(B.size());
        for (const auto &h : env->heightfields)
        {
            const unsigned long long addr = reinterpret_cast<unsigned long long>(h.d_data);
            pf(h.f[0]), pf(h.f[1]), pf(h.f[2]), pf(h.f[3]);
            pf(h.f[4]), pf(h.f[5]), pf(static_cast<float>(h.xd)), pf(static_cast<float>(h.yd));
            pf(static_cast<float>(h.xd / 2)), pf(static_cast<float>(h.yd / 2));
            B.push_back(static_cast<uint32_t>(addr & 0xffffffffu));
            B.push_back(static_cast<uint32_t>(addr >> 32));
        }
        H.n_capts = static_cast<uint32_t>(env->capts.size());
        align4();
        H.off_capts = static_cast<uint32_t>(B.size());
        for (const auto &t : env->capts)
        {
            vmv::CaptRec r{};
            r.r_point = t.r_point;
            r.nlog2 = static_cast<uint32_t>(t.nlog2);
            r.n_tests = (1u << t.nlog2) - 1u;
            r.n_points = t.n_points;
            for (int k = 0; k < 3; ++k)
            {
                r.lo[k] = t.top_lo[k];
                r.hi[k] = t.top_hi[k];
                r.g_origin[k] = t.g_origin[k];
            }
            r.r_max = t.r_max;
            r.list_reach_sq = t.list_reach_sq;
            r.g_inv0 = t.g_inv0;
            r.g_cell0 = t.g_cell0;
            r.n_grid_points = t.d_block != nullptr ? t.n_grid_points : static_cast<uint32_t>(t.gpts.size() / 4);
            r.nodes = t.d_nodes;
            r.leafbits = t.d_leafbits;
            r.gpts = t.d_gpts;
            r.gstart = t.d_gstart;
            const uint32_t *w = reinterpret_cast<const uint32_t *>(&r);
            B.insert(B.end(), w, w + vmv::kCaptRec);
        }
        H.n_mvts = static_cast<uint32_t>(env->mvts.size());
        align4();
        H.off_mvts = static_cast<uint32_t>(B.size());
        for (const auto &t : env->mvts)
        {
            vmv::MvtRec r{};
            r.r_point = t.r_point;
            r.inv_scale = t.inv_scale;
            r.grid_width = t.grid_width;
            for (int k = 0; k < 3; ++k)
            {
                r.ws_min[k] = t.ws_min[k];
                r.g_min[k] = t.g_min[k];
                r.g_max[k] = t.g_max[k];
            }
            r.cells = t.d_cells;
            r.voxels = t.d_voxels;
            r.points = t.d_points;
            const uint32_t *w = reinterpret_cast<const uint32_t *>(&r);
            B.insert(B.end(), w, w + vmv::kMvtRec);
        }
        H.off_cloud_grid = 0;
        if (!env->cloud_xyz.empty())
        {
            vmv::CloudGridRec g{};
            int rc = build_cloud_grid(env, g);
            if (rc != VMV_OK)
            {
                return rc;
            }
            if (g.cells != nullptr)
            {
                align4();
                H.off_cloud_grid = static_cast<uint32_t>(B.size());
                const uint32_t *w = reinterpret_cast<const uint32_t *>(&g);
                B.insert(B.end(), w, w + vmv::kCloudGridRec);
            }
        }
        align4();
        H.off_attach = static_cast<uint32_t>(B.size());
        env->attach_off = H.off_attach;
        std::memcpy(B.data(), &H, sizeof(H));
        append_attachment(env);

        {
            std::vector<int> ids;
            for (const auto &o : env->spheres) ids.push_back(o.id);
            for (const auto &o : env->capsules) ids.push_back(o.id);
            for (const auto &o : env->z_capsules) ids.push_back(o.id);
            for (const auto &o : env->cuboids) ids.push_back(o.id);
            for (const auto &o : env->z_cuboids) ids.push_back(o.id);
            for (const auto &o : env->heightfields) ids.push_back(o.id);
            int rc = upload(env, ids, env->d_object_ids);
            if (rc != VMV_OK)
            {
                return rc;
            }
        }
        env->d_blob_cap = B.size() * 4 + 4096;  // room for an attachment to come and go without reallocating
        VMV_CUDA(cudaMalloc(reinterpret_cast<void **>(&env->d_blob), env->d_blob_cap));
        VMV_CUDA(cudaMemcpy(env->d_blob, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        {
            build_rounded_boxes(env);
            std::vector<float4> recs(env->uobjs.size() / 4);
            std::memcpy(recs.data(), env->uobjs.data(), env->uobjs.size() * sizeof(float));
            int rc = upload(env, recs, env->d_uobjs);
            if (rc != VMV_OK)
            {
                return rc;
            }
        }
        env->shapes_dirty = false;
        env->committed = true;
        return VMV_OK;
    }

    // attach / detach on an otherwise unchanged, committed environment: rewrite the blob's tail only
    int repack_attachment(vmv_env *env)
    {
        append_attachment(env);
        const size_t bytes = env->blob.size() * 4;
        if (bytes > env->d_blob_cap)
        {
            VMV_CUDA(cudaFree(env->d_blob));
            env->d_blob = nullptr;
            env->d_blob_cap = bytes + 4096;
            VMV_CUDA(cudaMalloc(reinterpret_cast<void **>(&env->d_blob), env->d_blob_cap));
        }
        VMV_CUDA(cudaMemcpy(env->d_blob, env->blob.data(), bytes, cudaMemcpyHostToDevice));
        env->committed = true;
        return VMV_OK;
    }

}  // namespace

namespace vmvh
{
    int sm_count()
    {
        static int cached[kMaxDevices] = {0};
        int device = 0;
        cudaGetDevice(&device);
        if (device < kMaxDevices && cached[device] == 0)
        {
            cudaDeviceGetAttribute(&cached[device], cudaDevAttrMultiProcessorCount, device);
        }
        return device < kMaxDevices ? std::max(1, cached[device]) : 148;
    }

    unsigned long long *stats_buffer()
    {
        static unsigned long long *buf[kMaxDevices] = {};
        int device = 0;
        cudaGetDevice(&device);
        device %= kMaxDevices;
        if (buf[device] == nullptr)
        {
            cudaMalloc(reinterpret_cast<void **>(&buf[device]), 64 * sizeof(unsigned long long));
            cudaMemset(buf[device], 0, 64 * sizeof(unsigned long long));
        }
        return buf[device];
    }

    // Work counters of the persistent kernels: a ring of slots per device, two words per slot (tile ticket,
    // finished-block count), zeroed on the launch's stream.  A slot carries the event recorded behind its last
    // launch and is handed out again only once that event has completed: however many streams and host threads
    // launch concurrently, no two live kernels share a counter.
    namespace
    {
        constexpr int kCounterSlots = 256;
        struct CounterRing
        {
            unsigned int *mem = nullptr;
            cudaEvent_t events[kCounterSlots] = {};
            bool used[kCounterSlots] = {};
            bool busy[kCounterSlots] = {};  // acquired, not yet released (the launch is being issued)
            unsigned next = 0;
        };
        CounterRing g_rings[kMaxDevices];
        std::mutex g_ring_mutex;
    }  // namespace

    int counter_acquire(cudaStream_t s, unsigned int *&counter, int &slot)
    {
        int device = 0;
        VMV_CUDA(cudaGetDevice(&device));
        if (device >= kMaxDevices)
        {
            return fail(VMV_ERR_ARG, "device index too large");
        }
        {
            std::lock_guard<std::mutex> lock(g_ring_mutex);
            CounterRing &r = g_rings[device];
            if (r.mem == nullptr)
            {
                VMV_CUDA(cudaMalloc(reinterpret_cast<void **>(&r.mem), kCounterSlots * 2 * sizeof(unsigned int)));
            }
            slot = -1;
            for (int tries = 0; tries < kCounterSlots && slot < 0; ++tries)
            {
                const int k = static_cast<int>(r.next++ % kCounterSlots);
                if (r.busy[k])
                {
                    continue;
                }
                if (r.used[k])
                {
                    const cudaError_t q = cudaEventQuery(r.events[k]);
                    if (q == cudaErrorNotReady)
                    {
                        continue;
                    }
                    if (q != cudaSuccess)
                    {
                        return cuda_fail(q, "cudaEventQuery(counter slot)");
                    }
                }
                slot = k;
            }
            if (slot < 0)
            {
                // every slot belongs to a kernel still in flight: wait for the oldest one that is not being issued
                for (int tries = 0; tries < kCounterSlots && slot < 0; ++tries)
                {
                    const int k = static_cast<int>(r.next++ % kCounterSlots);
                    if (!r.busy[k])
                    {
                        VMV_CUDA(cudaEventSynchronize(r.events[k]));
                        slot = k;
                    }
                }
                if (slot < 0)
                {
                    return fail(VMV_ERR_LIMIT, "more than 256 launches being issued concurrently");
                }
            }
            r.busy[slot] = true;
            if (!r.used[slot])
            {
                VMV_CUDA(cudaEventCreateWithFlags(&r.events[slot], cudaEventDisableTiming));
                r.used[slot] = true;
            }
            counter = r.mem + 2 * slot;
        }
        VMV_CUDA(cudaMemsetAsync(counter, 0, 2 * sizeof(unsigned int), s));
        return VMV_OK;
    }

    int counter_release(cudaStream_t s, int slot)
    {
        int device = 0;
        VMV_CUDA(cudaGetDevice(&device));
        std::lock_guard<std::mutex> lock(g_ring_mutex);
        CounterRing &r = g_rings[device % kMaxDevices];
        const cudaError_t e = cudaEventRecord(r.events[slot], s);
        r.busy[slot] = false;
        if (e != cudaSuccess)
        {
            return cuda_fail(e, "cudaEventRecord(counter slot)");
        }
        return VMV_OK;
    }
}  // namespace vmvh

namespace
{
    int make_launch_env(const RobotHost &r, const vmv_env *env, vmv::LaunchEnv &le)
    {
        if (env == nullptr || !env->committed)
        {
            return fail(VMV_ERR_STATE, "environment is not committed (call vmv_env_commit)");
        }
        int device = 0;
        VMV_CUDA(cudaGetDevice(&device));
        if (device != env->device)
        {
            return fail(VMV_ERR_STATE, "environment was committed on another device");
        }
        le.blob = env->d_blob;
        le.blob_bytes = static_cast<uint32_t>(env->blob.size() * 4);
        le.n_objects = static_cast<uint32_t>(
            env->spheres.size() + env->capsules.size() + env->z_capsules.size() + env->cuboids.size() + env->z_cuboids.size());
        le.primitives_only = env->heightfields.empty() && env->capts.empty() && env->mvts.empty() && !env->has_attachment;
        le.static_mode = 0;
        // attach_tf = ee_tf(robot) * attachment offset
        const float *E = r.ee_tf;
        float A[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
        if (env->has_attachment)
        {
            std::memcpy(A, env->attach_tf, sizeof(A));
        }
        for (int i = 0; i < 3; ++i)
        {
            for (int j = 0; j < 4; ++j)
            {
                float s = E[4 * i] * A[j] + E[4 * i + 1] * A[4 + j] + E[4 * i + 2] * A[8 + j];
                le.attach_tf[4 * i + j] = (j == 3) ? s + E[4 * i + 3] : s;
            }
        }
        return VMV_OK;
    }

    // ---- grid-culled path ---------------------------------------------------------------------
    constexpr size_t kGridMinConfigs = 1024;  // smaller batches do not amortise a table build
    constexpr size_t kGridMinEdges = 64;  // the edge kernel cuts small batches into 8-edge chunks and wins from here
    constexpr size_t kGridMaxVoxels = 1000000;

    struct RobotGridInfo
    {
        bool ready = false;
        float class_r[vmv::kGridClasses];
        unsigned char link_class[vmv::kGridMaxLinks];
        float reach;
        uint32_t max_fine;
    };
    RobotGridInfo g_grid_info[VMV_N_ROBOTS];

    // Radius classes of a robot's link bounding spheres: the partition of the sorted radii into at
    // most kGridClasses groups that minimises the total over-estimate (class radius - link radius).
    const RobotGridInfo &robot_grid_info(int robot)
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        RobotGridInfo &gi = g_grid_info[robot];
        if (gi.ready)
        {
            return gi;
        }
        const RobotHost &r = robot_host(robot);
        const int n = r.n_links;
        std::vector<int> order(n);
        std::iota(order.begin(), order.end(), 0);
        auto rad = [&](int l) { return r.tasks[r.links[l].bound_task].r; };
        std::sort(order.begin(), order.end(), [&](int a, int b) { return rad(a) < rad(b); });
        constexpr int K = vmv::kGridClasses;
        const double inf = 1e30;
        std::vector<std::vector<double>> dp(K + 1, std::vector<double>(n + 1, inf));
        std::vector<std::vector<int>> cut(K + 1, std::vector<int>(n + 1, 0));
        dp[0][0] = 0;
        for (int g = 1; g <= K; ++g)
        {
            for (int j = 1; j <= n; ++j)
            {
                for (int i = g - 1; i < j; ++i)
                {
                    if (dp[g - 1][i] >= inf)
                    {
                        continue;
                    }
                    double cost = 0;
                    for (int k = i; k < j; ++k)
                    {
                        cost += rad(order[j - 1]) - rad(order[k]);
                    }
                    if (dp[g - 1][i] + cost < dp[g][j])
                    {
                        dp[g][j] = dp[g - 1][i] + cost;
                        cut[g][j] = i;
                    }
                }
            }
        }
        int groups = std::min(K, n);
        for (int k = 0; k < K; ++k)
        {
            gi.class_r[k] = 0.F;
        }
        int j = n;
        for (int g = groups; g >= 1; --g)
        {
            const int i = cut[g][j];
            gi.class_r[g - 1] = rad(order[j - 1]);
            for (int k = i; k < j; ++k)
            {
                gi.link_class[order[k]] = static_cast<unsigned char>(g - 1);
            }
            j = i;
        }
        for (int k = groups; k < K; ++k)
        {
            gi.class_r[k] = gi.class_r[groups - 1];
        }
        gi.max_fine = 1;
        for (int l = 0; l < n; ++l)
        {
            gi.max_fine = std::max<uint32_t>(gi.max_fine, static_cast<uint32_t>(r.links[l].n_spheres));
        }
        gi.reach = r.max_reach;
        gi.ready = true;
        return gi;
    }

    // The voxel table of `env` for `robot` (built once, synchronously, on first use).  ok = false when
    // the path does not apply (no primitives, more than 64 of them, other content in the environment).
    int grid_launch_env(int robot, const vmv_env *env, vmv::GridEnv &out, bool &ok, bool &wide, bool configs = false)
    {
        ok = false;
        wide = false;
        const size_t n_obj = env->uobjs.size() / vmv::kObjRec;
        // any environment (configuration batches): heightfields and pointclouds next to at most 30 primitives go through
        // the grid-culled kernel's any-environment instantiation; attachments and edge batches stay on the per-thread kernel
        const bool extra = !env->heightfields.empty() || !env->capts.empty() || !env->mvts.empty();
        const bool ae = extra && configs && n_obj <= 30 && robot_host(robot).n_links <= 64 && std::getenv("VMV_NO_AE") == nullptr;
        // an attachment rides along (phase D of the grid-culled kernels; next to pointclouds for configuration batches only)
        if (((n_obj == 0 || extra) && !ae) || n_obj > 64 || (env->has_attachment && std::getenv("VMV_NO_V4_ATTACH") != nullptr))
        {
            return VMV_OK;
        }
        const RobotGridInfo &gi = robot_grid_info(robot);
        std::lock_guard<std::mutex> lock(env->grid_mutex);
        vmv_env::GridCache &gc = env->grids[robot];
        if (!gc.ready)
        {
            gc.ready = true;
            gc.usable = false;
            const float r_max = *std::max_element(gi.class_r, gi.class_r + vmv::kGridClasses);
            // region: bounding box of the objects grown by r_max (a centre outside it is farther than
            // r_max from every object), clipped to the cube the robot can reach
            const float grow = r_max + 1e-3F;
            double lo[3] = {1e30, 1e30, 1e30}, hi[3] = {-1e30, -1e30, -1e30};
            bool finite = true;
            for (size_t k = 0; k < n_obj; ++k)
            {
                const float *u = env->uobjs.data() + k * vmv::kObjRec;
                for (int j = 0; j < 3; ++j)
                {
                    const double ext = std::fabs(u[4 + j]) * u[7] + std::fabs(u[8 + j]) * u[11] + std::fabs(u[12 + j]) * u[15] + u[3] + grow;
                    lo[j] = std::min(lo[j], u[j] - ext);
                    hi[j] = std::max(hi[j], u[j] + ext);
                    finite = finite && std::isfinite(ext) && std::isfinite(u[j]);
                }
            }
            if (!finite)
            {
                return VMV_OK;
            }
            bool empty = false;
            for (int j = 0; j < 3; ++j)
            {
                lo[j] = std::max<double>(lo[j], -gi.reach);
                hi[j] = std::min<double>(hi[j], gi.reach);
                empty = empty || !(hi[j] > lo[j]);
            }
            if (empty || n_obj == 0)
            {
                // nothing within reach (or no primitive at all): a 1-voxel table of empty masks far away
                lo[0] = lo[1] = lo[2] = 1e6;
                hi[0] = hi[1] = hi[2] = 1e6 + 1;
            }
            double h = 0.025;
            int nx, ny, nz;
            while (true)
            {
                nx = std::max(1, static_cast<int>(std::ceil((hi[0] - lo[0]) / h)));
                ny = std::max(1, static_cast<int>(std::ceil((hi[1] - lo[1]) / h)));
                nz = std::max(1, static_cast<int>(std::ceil((hi[2] - lo[2]) / h)));
                if (static_cast<size_t>(nx) * ny * nz <= kGridMaxVoxels)
                {
                    break;
                }
                h *= 1.15;
            }
            const size_t n_vox = static_cast<size_t>(nx) * ny * nz;
            gc.wide = n_obj > 32;
            gc.bytes = n_vox * vmv::kGridClasses * (gc.wide ? sizeof(unsigned long long) : sizeof(uint32_t));
            VMV_CUDA(cudaMalloc(&gc.mem, gc.bytes));
            cudaEvent_t e0, e1;
            VMV_CUDA(cudaEventCreate(&e0));
            VMV_CUDA(cudaEventCreate(&e1));
            VMV_CUDA(cudaEventRecord(e0, nullptr));
            if (gc.wide)
            {
                vmv::k_build_grid_t<unsigned long long><<<static_cast<unsigned>((n_vox + 127) / 128), 128>>>(
                    env->d_uobjs, static_cast<uint32_t>(n_obj), static_cast<float>(lo[0]), static_cast<float>(lo[1]), static_cast<float>(lo[2]),
                    static_cast<float>(h), nx, ny, nz, gi.class_r[0], gi.class_r[1], gi.class_r[2], gi.class_r[3],
                    static_cast<unsigned long long *>(gc.mem));
            }
            else
            {
                vmv::k_build_grid_t<uint32_t><<<static_cast<unsigned>((n_vox + 127) / 128), 128>>>(
                    env->d_uobjs, static_cast<uint32_t>(n_obj), static_cast<float>(lo[0]), static_cast<float>(lo[1]), static_cast<float>(lo[2]),
                    static_cast<float>(h), nx, ny, nz, gi.class_r[0], gi.class_r[1], gi.class_r[2], gi.class_r[3],
                    static_cast<uint32_t *>(gc.mem));
            }
            g_launches++;
            VMV_CUDA(cudaGetLastError());
            VMV_CUDA(cudaEventRecord(e1, nullptr));
            VMV_CUDA(cudaEventSynchronize(e1));
            VMV_CUDA(cudaEventElapsedTime(&gc.build_ms, e0, e1));
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
            gc.dev.masks = static_cast<const unsigned long long *>(gc.mem);
            gc.dev.x0 = static_cast<float>(lo[0]), gc.dev.y0 = static_cast<float>(lo[1]), gc.dev.z0 = static_cast<float>(lo[2]);
            gc.dev.inv_h = static_cast<float>(1.0 / h);
            gc.dev.nx = nx, gc.dev.ny = ny, gc.dev.nz = nz;
            gc.dev.all_mask = n_obj >= 64 ? ~0ull : ((1ull << n_obj) - 1ull);
            std::memcpy(gc.dev.link_class, gi.link_class, sizeof(gc.dev.link_class));
            gc.usable = true;
        }
        if (!gc.usable)
        {
            return VMV_OK;
        }
        out.objs = env->d_uobjs;
        out.n_objects = static_cast<uint32_t>(n_obj);
        out.max_fine = gi.max_fine;
        out.grid = gc.dev;
        out.blob = ae ? env->d_blob : nullptr;
        out.blob_bytes = ae ? static_cast<uint32_t>(env->blob.size() * 4) : 0u;
        out.att = vmv::AttachDev{};
        if (env->has_attachment)
        {
            // the spheres sit at the tail of the device blob ({n 0 0 0} then n x {x y z r}); tf = ee_tf(robot) * attachment offset
            out.att.spheres = reinterpret_cast<const float4 *>(env->d_blob) + (env->attach_off + vmv::kAttachHdr) / 4;
            out.att.n = static_cast<uint32_t>(env->attach_spheres.size() / 4);
            const float *E = robot_host(robot).ee_tf;
            for (int i = 0; i < 3; ++i)
            {
                for (int j = 0; j < 4; ++j)
                {
                    const float s = E[4 * i] * env->attach_tf[j] + E[4 * i + 1] * env->attach_tf[4 + j] + E[4 * i + 2] * env->attach_tf[8 + j];
                    out.att.tf[4 * i + j] = (j == 3) ? s + E[4 * i + 3] : s;
                }
            }
        }
        wide = gc.wide;
        ok = true;
        return VMV_OK;
    }
    bool valid_robot(int robot)
    {
        return robot >= 0 && robot < VMV_N_ROBOTS;
    }
}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
namespace
{
    // No exception crosses the C ABI (include/vamp_b200.h): allocation failures and anything else thrown by the
    // host-side builds (CAPT / MVT construction, vectors, std::thread) become error codes.
    template <typename F>
    auto guarded(const char *what, F &&f) noexcept -> decltype(f())
    {
        try
        {
            return f();
        }
        catch (const std::bad_alloc &)
        {
            return fail(VMV_ERR_LIMIT, std::string(what) + ": out of host memory");
        }
        catch (const std::exception &e)
        {
            return fail(VMV_ERR_STATE, std::string(what) + ": " + e.what());
        }
        catch (...)
        {
            return fail(VMV_ERR_STATE, std::string(what) + ": unknown exception");
        }
    }
}  // namespace

extern "C"
{
    const char *vmv_last_error(void)
    {
        return g_error.c_str();
    }

    const char *vmv_version(void)
    {
        return "vamp_mvt_b200 0.1 (sm_100a)";
    }

    int vmv_device_count(void)
    {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess)
        {
            cudaGetLastError();
            return 0;
        }
        return n;
    }

    int vmv_set_device(int device)
    {
        VMV_CUDA(cudaSetDevice(device));
        return VMV_OK;
    }

    int vmv_robot_id(const char *name)
    {
        for (int i = 0; i < VMV_N_ROBOTS; ++i)
        {
            if (name != nullptr && std::strcmp(name, robot_host(i).name) == 0)
            {
                return i;
            }
        }
        return fail(VMV_ERR_ARG, "unknown robot");
    }

    const char *vmv_robot_name(int robot)
    {
        return valid_robot(robot) ? robot_host(robot).name : nullptr;
    }

    int vmv_robot_dof(int robot)
    {
        return valid_robot(robot) ? robot_host(robot).dof : fail(VMV_ERR_ARG, "unknown robot");
    }

    int vmv_robot_n_spheres(int robot)
    {
        return valid_robot(robot) ? robot_host(robot).n_spheres : fail(VMV_ERR_ARG, "unknown robot");
    }

    int vmv_robot_resolution(int robot)
    {
        return valid_robot(robot) ? robot_host(robot).resolution : fail(VMV_ERR_ARG, "unknown robot");
    }

    int vmv_robot_bounds(int robot, float *lower, float *range)
    {
        return guarded("vmv_robot_bounds", [&]() -> int
        {
            if (!valid_robot(robot) || lower == nullptr || range == nullptr)
            {
                return fail(VMV_ERR_ARG, "vmv_robot_bounds: bad argument");
            }
            std::memcpy(lower, robot_host(robot).lower, robot_host(robot).dof * sizeof(float));
            std::memcpy(range, robot_host(robot).range, robot_host(robot).dof * sizeof(float));
            return VMV_OK;
        });
    }

    vmv_env *vmv_env_create(void)
    {
        return new (std::nothrow) vmv_env();
    }

    void vmv_env_destroy(vmv_env *env)
    {
        if (env != nullptr)
        {
            env->release_device();
            env->release_clouds();
            delete env;
        }
    }

    int vmv_env_add_spheres(vmv_env *env, const float *f, size_t n)
    {
        return guarded("vmv_env_add_spheres", [&]() -> int
        {
            if (env == nullptr || (f == nullptr && n > 0))
            {
                return fail(VMV_ERR_ARG, "vmv_env_add_spheres: null argument");
            }
            for (size_t i = 0; i < n; ++i, f += 4)
            {
                HSphere s{f[0], f[1], f[2], f[3], 0.F, env->next_id++};
                s.min_d = std::sqrt(s.x * s.x + s.y * s.y + s.z * s.z) - s.r;  // shapes.hh:238
                env->spheres.push_back(s);
            }
            env->shapes_dirty = true;
            env->committed = false;
            return VMV_OK;
        });
    }

    int vmv_env_add_cuboids(vmv_env *env, const float *f, size_t n)
    {
        return guarded("vmv_env_add_cuboids", [&]() -> int
        {
            if (env == nullptr || (f == nullptr && n > 0))
            {
                return fail(VMV_ERR_ARG, "vmv_env_add_cuboids: null argument");
            }
            for (size_t i = 0; i < n; ++i, f += 15)
            {
                HCuboid c{};
                std::memcpy(c.f, f, sizeof(c.f));
                c.id = env->next_id++;
                // Cuboid::compute_min_distance, shapes.hh:52-67
                const float x = f[0], y = f[1], z = f[2];
                const float d1 = dot3(-x, -y, -z, f[3], f[4], f[5]);
                const float d2 = dot3(-x, -y, -z, f[6], f[7], f[8]);
                const float d3 = dot3(-x, -y, -z, f[9], f[10], f[11]);
                const float v1 = clamp_ref(d1, -f[12], f[12]);
                const float v2 = clamp_ref(d2, -f[13], f[13]);
                const float v3 = clamp_ref(d3, -f[14], f[14]);
                const float xn = x + f[3] * v1 + f[6] * v2 + f[9] * v3;
                const float yn = y + f[4] * v1 + f[7] * v2 + f[10] * v3;
                const float zn = z + f[5] * v1 + f[8] * v2 + f[11] * v3;
                c.min_d = finite_or_neg_inf(std::sqrt(xn * xn + yn * yn + zn * zn));
                // z-aligned iff axis_3_z == 1 (bindings/environment.cc:124)
                (f[11] == 1.F ? env->z_cuboids : env->cuboids).push_back(c);
            }
            env->shapes_dirty = true;
            env->committed = false;
            return VMV_OK;
        });
    }

    int vmv_env_add_capsules(vmv_env *env, const float *f, size_t n)
    {
        return guarded("vmv_env_add_capsules", [&]() -> int
        {
            if (env == nullptr || (f == nullptr && n > 0))
            {
                return fail(VMV_ERR_ARG, "vmv_env_add_capsules: null argument");
            }
            for (size_t i = 0; i < n; ++i, f += 8)
            {
                HCapsule c{f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], 0.F, env->next_id++};
                // Cylinder::compute_min_distance, shapes.hh:165-189
                const float dot = clamp_ref(dot3(-c.x1, -c.y1, -c.z1, c.xv, c.yv, c.zv) * c.rdv, 0.F, 1.F);
                const float xp = c.x1 + c.xv * dot, yp = c.y1 + c.yv * dot, zp = c.z1 + c.zv * dot;
                float xo = -xp, yo = -yp, zo = -zp;
                const float ol = std::sqrt(dot3(xo, yo, zo, xo, yo, zo));
                xo = xo / ol, yo = yo / ol, zo = zo / ol;
                const float ro = clamp_ref(ol, 0.F, c.r);
                const float xn = xp + ro * xo, yn = yp + ro * yo, zn = zp + ro * zo;
                c.min_d = finite_or_neg_inf(std::sqrt(xn * xn + yn * yn + zn * zn));
                // z-aligned iff xv == 0 and yv == 0 (bindings/environment.cc:138)
                ((c.xv == 0.F && c.yv == 0.F) ? env->z_capsules : env->capsules).push_back(c);
            }
            env->shapes_dirty = true;
            env->committed = false;
            return VMV_OK;
        });
    }

    int vmv_env_add_heightfield(vmv_env *env, const float *f6, size_t xd, size_t yd, const float *data)
    {
        return guarded("vmv_env_add_heightfield", [&]() -> int
        {
            if (env == nullptr || f6 == nullptr || data == nullptr || xd == 0 || yd == 0)
            {
                return fail(VMV_ERR_ARG, "vmv_env_add_heightfield: bad argument");
            }
            HHeight h;
            std::memcpy(h.f, f6, sizeof(h.f));
            h.xd = xd, h.yd = yd;
            h.data.assign(data, data + xd * yd);
            h.id = env->next_id++;
            env->heightfields.push_back(std::move(h));
            env->shapes_dirty = true;
            env->committed = false;
            return VMV_OK;
        });
    }

    int vmv_env_add_capt(vmv_env *env, const float *pts, size_t n, float r_min, float r_max, float r_point)
    {
        return guarded("vmv_env_add_capt", [&]() -> int
        {
            if (env == nullptr || (pts == nullptr && n > 0))
            {
                return fail(VMV_ERR_ARG, "vmv_env_add_capt: null argument");
            }
            if (n > (size_t(1) << 24))
            {
                return fail(VMV_ERR_LIMIT, "vmv_env_add_capt: more than 2^24 points in one pointcloud");
            }
            HCapt t;
            bool any_nan = false;
            for (size_t k = 0; k < 3 * n && !any_nan; ++k)
            {
                any_nan = std::isnan(pts[k]);
            }
            if (any_nan || std::getenv("VMV_CAPT_HOST_BUILD") != nullptr)
            {
                capt_build(t, pts, n, r_min, r_max, r_point);
            }
            else
            {
                const int rc = capt_build_device(t, pts, n, r_min, r_max, r_point);
                if (rc != VMV_OK)
                {
                    if (t.d_block != nullptr)
                    {
                        cudaFree(t.d_block);
                    }
                    return rc;
                }
            }
            t.id = env->next_id++;
            {
                const size_t cloud = env->capts.size();
                const bool taggable = cloud < vmv::kCloudTagNone;
                for (size_t k = 0; k < n; ++k)
                {
                    env->cloud_tag.push_back(taggable ? (static_cast<uint32_t>(cloud) << 24 | t.leaf_of_point[k]) : (vmv::kCloudTagNone << 24));
                }
                std::vector<uint32_t>().swap(t.leaf_of_point);
            }
            env->capts.push_back(std::move(t));
            env->cloud_xyz.insert(env->cloud_xyz.end(), pts, pts + 3 * n);
            env->cloud_r_point_max = std::max(env->cloud_r_point_max, r_point);
            env->shapes_dirty = true;
            env->committed = false;
            return VMV_OK;
        });
    }

    int vmv_env_add_mvt(vmv_env *env, const float *pts, size_t n, float r_min, float r_max, const float *aabb_min, const float *aabb_max, float r_point)
    {
        return guarded("vmv_env_add_mvt", [&]() -> int
        {
            if (env == nullptr || (n > 0 && pts == nullptr) || aabb_min == nullptr || aabb_max == nullptr || !(r_max > 0.F))
            {
                return fail(VMV_ERR_ARG, "vmv_env_add_mvt: bad argument");
            }
            HMvt t;
            const int rc = mvt_build(t, pts, n, r_min, r_max, aabb_min, aabb_max, r_point);
            if (rc == -1)
            {
                return fail(VMV_ERR_ARG, "vmv_env_add_mvt: workspace narrower than r_max");
            }
            if (rc == -2)
            {
                return fail(VMV_ERR_LIMIT, "vmv_env_add_mvt: more than 2^26 grid cells");
            }
            t.id = env->next_id++;
            env->mvts.push_back(std::move(t));
            env->cloud_tag.insert(env->cloud_tag.end(), n, vmv::kCloudTagNone << 24);
            env->cloud_xyz.insert(env->cloud_xyz.end(), pts, pts + 3 * n);
            env->cloud_r_point_max = std::max(env->cloud_r_point_max, r_point);
            env->shapes_dirty = true;
            env->committed = false;
            return VMV_OK;
        });
    }

    int vmv_env_attach(vmv_env *env, const float *tf12, const float *spheres, size_t n)
    {
        return guarded("vmv_env_attach", [&]() -> int
        {
            if (env == nullptr || tf12 == nullptr || (spheres == nullptr && n > 0))
            {
                return fail(VMV_ERR_ARG, "vmv_env_attach: null argument");
            }
            for (int r = 0; r < 3; ++r)
            {
                for (int c = 0; c < 3; ++c)
                {
                    env->attach_tf[4 * r + c] = tf12[3 + 3 * c + r];
                }
                env->attach_tf[4 * r + 3] = tf12[r];
            }
            env->attach_spheres.assign(spheres, spheres + 4 * n);
            env->has_attachment = true;
            env->committed = false;
            return VMV_OK;
        });
    }

    int vmv_env_detach(vmv_env *env)
    {
        return guarded("vmv_env_detach", [&]() -> int
        {
            if (env == nullptr)
            {
                return fail(VMV_ERR_ARG, "vmv_env_detach: null argument");
            }
            env->has_attachment = false;
            env->attach_spheres.clear();
            env->committed = false;
            return VMV_OK;
        });
    }

    int vmv_env_commit(vmv_env *env)
    {
        return guarded("vmv_env_commit", [&]() -> int
        {
            if (env == nullptr)
            {
                return fail(VMV_ERR_ARG, "vmv_env_commit: null argument");
            }
            if (!env->shapes_dirty && env->d_blob != nullptr)
            {
                int device = 0;
                VMV_CUDA(cudaGetDevice(&device));
                if (device == env->device)
                {
                    return env->committed ? VMV_OK : repack_attachment(env);
                }
            }
            sort_by_min_distance(env->spheres);
            sort_by_min_distance(env->capsules);
            sort_by_min_distance(env->z_capsules);
            sort_by_min_distance(env->cuboids);
            sort_by_min_distance(env->z_cuboids);
            return pack_and_upload(env);
        });
    }

    long vmv_env_dump(const vmv_env *env, int kind, float *out, size_t cap)
    {
        return guarded("vmv_env_dump", [&]() -> long
        {
            if (env == nullptr)
            {
                return fail(VMV_ERR_ARG, "vmv_env_dump: null argument");
            }
            // dump reflects the order a commit would produce
            vmv_env tmp;
            tmp.spheres = env->spheres, tmp.capsules = env->capsules, tmp.z_capsules = env->z_capsules;
            tmp.cuboids = env->cuboids, tmp.z_cuboids = env->z_cuboids;
            sort_by_min_distance(tmp.spheres), sort_by_min_distance(tmp.capsules), sort_by_min_distance(tmp.z_capsules);
            sort_by_min_distance(tmp.cuboids), sort_by_min_distance(tmp.z_cuboids);
            size_t k = 0;
            auto put = [&](float v)
            {
                if (k < cap && out != nullptr)
                {
                    out[k] = v;
                }
                ++k;
            };
            switch (kind)
            {
                case 0:
                    for (const auto &s : tmp.spheres)
                    {
                        put(s.x), put(s.y), put(s.z), put(s.r), put(s.min_d);
                    }
                    return static_cast<long>(tmp.spheres.size());
                case 1:
                case 2:
                {
                    const auto &v = kind == 1 ? tmp.capsules : tmp.z_capsules;
                    for (const auto &c : v)
                    {
                        put(c.x1), put(c.y1), put(c.z1), put(c.xv), put(c.yv), put(c.zv), put(c.r), put(c.rdv), put(c.min_d);
                    }
                    return static_cast<long>(v.size());
                }
                case 3:
                case 4:
                {
                    const auto &v = kind == 3 ? tmp.cuboids : tmp.z_cuboids;
                    for (const auto &c : v)
                    {
                        for (float f : c.f)
                        {
                            put(f);
                        }
                        put(c.min_d);
                    }
                    return static_cast<long>(v.size());
                }
                default:
                    return fail(VMV_ERR_ARG, "vmv_env_dump: unknown kind");
            }
        });
    }


    // gathers for the kernel generations without fused peer stores: local words -> every window, then the flags
    static int push_gather(const GatherDev &gather, const uint32_t *d_bits, size_t n_words, cudaStream_t s)
    {
        if (gather.world == 0)
        {
            return VMV_OK;
        }
        GatherDev g = gather;
        unsigned int *counter = nullptr;
        int slot = -1;
        int rc = counter_acquire(s, counter, slot);
        if (rc != VMV_OK)
        {
            return rc;
        }
        g.done = counter + 1;
        const unsigned grid = static_cast<unsigned>(std::min<size_t>((n_words + 255) / 256, static_cast<size_t>(sm_count())));
        vmv::k_comm_push<<<std::max(1u, grid), 256, 0, s>>>(d_bits, n_words, g);
        g_launches++;
        const cudaError_t e = cudaGetLastError();
        rc = counter_release(s, slot);
        return e != cudaSuccess ? cuda_fail(e, "k_comm_push launch") : rc;
    }

    // The robot's world-fixed links against the environment, once per environment (any-environment kernels; the
    // grid-culled kernels answer them with one table load).  state: 1 clear, 2 in collision.
    static int static_links_state(int robot, const vmv_env *env, const vmv::RobotDev &rd, vmv::LaunchEnv le, int &state)
    {
        const RobotHost &rh = robot_host(robot);
        bool any = false;
        for (int t = 0; t < rh.n_tasks; ++t)
        {
            any = any || rh.tasks[t].body == 0;
        }
        if (!any || le.primitives_only)
        {
            state = 0;  // nothing to gain: leave the launch as it is
            return VMV_OK;
        }
        std::lock_guard<std::mutex> lock(env->grid_mutex);
        if (env->static_state[robot] == 0)
        {
            std::vector<float> q(rh.dof);
            for (int j = 0; j < rh.dof; ++j)
            {
                q[j] = rh.lower[j] + 0.5F * rh.range[j];
            }
            void *dq = nullptr, *dw = nullptr;
            VMV_CUDA(cudaMalloc(&dq, rh.dof * sizeof(float)));
            VMV_CUDA(cudaMalloc(&dw, sizeof(uint32_t)));
            VMV_CUDA(cudaMemcpy(dq, q.data(), rh.dof * sizeof(float), cudaMemcpyHostToDevice));
            le.static_mode = 2;
            int rc = ops(robot).configs(rd, le, static_cast<const float *>(dq), 1, static_cast<uint32_t *>(dw), nullptr);
            uint32_t w = 0;
            if (rc == VMV_OK)
            {
                const cudaError_t e = cudaMemcpy(&w, dw, sizeof(w), cudaMemcpyDeviceToHost);
                rc = e == cudaSuccess ? VMV_OK : cuda_fail(e, "cudaMemcpy(static links verdict)");
            }
            cudaFree(dq);
            cudaFree(dw);
            if (rc != VMV_OK)
            {
                return rc;
            }
            env->static_state[robot] = (w & 1u) ? 1 : 2;
        }
        state = env->static_state[robot];
        return VMV_OK;
    }

    static int configs_common(int robot, const vmv_env *env, const float *d_q, size_t n, uint32_t *d_bits, const GatherDev &gather, void *stream)
    {
        vmv::LaunchEnv le{};
        int rc = make_launch_env(robot_host(robot), env, le);
        if (rc != VMV_OK || n == 0)
        {
            return rc;
        }
        vmv::RobotDev rd{};
        rc = robot_tables(robot, rd);
        if (rc != VMV_OK)
        {
            return rc;
        }
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const int force = g_force_path.load();
        if ((force == 0 && n >= kGridMinConfigs) || force == 3)
        {
            vmv::GridEnv l3{};
            bool ok = false, wide = false;
            rc = grid_launch_env(robot, env, l3, ok, wide, true);
            if (rc != VMV_OK)
            {
                return rc;
            }
            if (ok)
            {
                if (l3.blob_bytes > 0 || l3.att.n > 0)
                {
                    // any-environment / attachment instantiations: local launch, then the gather (if any) as a push
                    rc = ops(robot).configs_v4(robot, wide, rd, l3, d_q, n, d_bits, GatherDev{}, s);
                    if (rc == VMV_OK)
                    {
                        return push_gather(gather, d_bits, (n + 31) / 32, s);
                    }
                }
                else
                {
                    rc = ops(robot).configs_v4(robot, wide, rd, l3, d_q, n, d_bits, gather, s);
                }
                if (rc != VMV_ERR_LIMIT || force == 3)
                {
                    return rc;
                }
            }
            else if (force == 3)
            {
                return fail(VMV_ERR_LIMIT, "grid-culled kernel not applicable to this environment");
            }
        }
        if (force != 2)
        {
            int st = 0;
            rc = static_links_state(robot, env, rd, le, st);
            if (rc != VMV_OK)
            {
                return rc;
            }
            if (st == 2)
            {
                // a world-fixed link is in collision: no configuration is valid
                VMV_CUDA(cudaMemsetAsync(d_bits, 0, ((n + 31) / 32) * sizeof(uint32_t), s));
                return push_gather(gather, d_bits, (n + 31) / 32, s);
            }
            le.static_mode = st == 1 ? 1u : 0u;
        }
        rc = ops(robot).configs(rd, le, d_q, n, d_bits, s);
        return rc == VMV_OK ? push_gather(gather, d_bits, (n + 31) / 32, s) : rc;
    }

    int vmv_validate_configs_dev(int robot, const vmv_env *env, const float *d_q, size_t n, uint32_t *d_bits, void *stream)
    {
        return guarded("vmv_validate_configs_dev", [&]() -> int
        {
            if (!valid_robot(robot) || (n > 0 && (d_q == nullptr || d_bits == nullptr)))
            {
                return fail(VMV_ERR_ARG, "vmv_validate_configs_dev: bad argument");
            }
            return configs_common(robot, env, d_q, n, d_bits, GatherDev{}, stream);
        });
    }

    static int edges_common(
        int robot,
        const vmv_env *env,
        const float *d_a,
        const float *d_b,
        const uint32_t *d_pairs,
        size_t n,
        int resolution,
        uint32_t *d_bits,
        const GatherDev &gather,
        void *stream)
    {
        vmv::LaunchEnv le{};
        int rc = make_launch_env(robot_host(robot), env, le);
        if (rc != VMV_OK || n == 0)
        {
            return rc;
        }
        vmv::RobotDev rd{};
        rc = robot_tables(robot, rd);
        if (rc != VMV_OK)
        {
            return rc;
        }
        const float res = static_cast<float>(resolution > 0 ? resolution : robot_host(robot).resolution);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const int force = g_force_path.load();
        if ((force == 0 && n >= kGridMinEdges) || force == 3)
        {
            vmv::GridEnv l3{};
            bool ok = false, wide = false;
            rc = grid_launch_env(robot, env, l3, ok, wide);
            if (rc != VMV_OK)
            {
                return rc;
            }
            if (ok)
            {
                if (l3.att.n > 0 && gather.world > 0)
                {
                    // attachment instantiation: local launch, then the gather as a push
                    rc = ops(robot).edges_v4(robot, wide, rd, l3, d_a, d_b, d_pairs, n, res, d_bits, GatherDev{}, s);
                    if (rc == VMV_OK)
                    {
                        return push_gather(gather, d_bits, (n + 31) / 32, s);
                    }
                }
                else
                {
                    rc = ops(robot).edges_v4(robot, wide, rd, l3, d_a, d_b, d_pairs, n, res, d_bits, gather, s);
                }
                if (rc != VMV_ERR_LIMIT || force == 3)
                {
                    return rc;
                }
            }
            else if (force == 3)
            {
                return fail(VMV_ERR_LIMIT, "grid-culled kernel not applicable to this environment");
            }
        }
        if (force != 2)
        {
            int st = 0;
            rc = static_links_state(robot, env, rd, le, st);
            if (rc != VMV_OK)
            {
                return rc;
            }
            if (st == 2)
            {
                VMV_CUDA(cudaMemsetAsync(d_bits, 0, ((n + 31) / 32) * sizeof(uint32_t), s));
                return push_gather(gather, d_bits, (n + 31) / 32, s);
            }
            le.static_mode = st == 1 ? 1u : 0u;
        }
        rc = ops(robot).edges(rd, le, d_a, d_b, d_pairs, n, res, d_bits, s);
        return rc == VMV_OK ? push_gather(gather, d_bits, (n + 31) / 32, s) : rc;
    }

    int vmv_validate_edges_dev(int robot, const vmv_env *env, const float *d_a, const float *d_b, size_t n, int resolution, uint32_t *d_bits, void *stream)
    {
        return guarded("vmv_validate_edges_dev", [&]() -> int
        {
            if (!valid_robot(robot) || (n > 0 && (d_a == nullptr || d_b == nullptr || d_bits == nullptr)))
            {
                return fail(VMV_ERR_ARG, "vmv_validate_edges_dev: bad argument");
            }
            return edges_common(robot, env, d_a, d_b, nullptr, n, resolution, d_bits, GatherDev{}, stream);
        });
    }

    int vmv_validate_edges_indexed_dev(int robot, const vmv_env *env, const float *d_vertices, size_t n_vertices, const uint32_t *d_pairs, size_t n_edges, int resolution, uint32_t *d_bits, void *stream)
    {
        return guarded("vmv_validate_edges_indexed_dev", [&]() -> int
        {
            (void)n_vertices;
            if (!valid_robot(robot) || (n_edges > 0 && (d_vertices == nullptr || d_pairs == nullptr || d_bits == nullptr)))
            {
                return fail(VMV_ERR_ARG, "vmv_validate_edges_indexed_dev: bad argument");
            }
            return edges_common(robot, env, d_vertices, nullptr, d_pairs, n_edges, resolution, d_bits, GatherDev{}, stream);
        });
    }

    int vmv_sphere_fk_dev(int robot, const float *d_q, size_t n, float *d_xyzr, void *stream)
    {
        return guarded("vmv_sphere_fk_dev", [&]() -> int
        {
            if (!valid_robot(robot) || (n > 0 && (d_q == nullptr || d_xyzr == nullptr)))
            {
                return fail(VMV_ERR_ARG, "vmv_sphere_fk_dev: bad argument");
            }
            if (n == 0)
            {
                return VMV_OK;
            }
            vmv::RobotDev rd{};
            int rc = robot_tables(robot, rd);
            if (rc != VMV_OK)
            {
                return rc;
            }
            cudaStream_t s = static_cast<cudaStream_t>(stream);
            return ops(robot).fk(rd, d_q, n, d_xyzr, s);
        });
    }

    // ---- host-buffer versions: H2D + kernel + D2H + sync ------------------------------------
    // The batch is cut into chunks that alternate between two streams so that the upload of chunk
    // k+1 overlaps the kernel of chunk k (pinned host memory makes the copies truly asynchronous;
    // pageable memory still works, staged by the driver).  Device scratch is a grow-only pool per
    // device, so steady-state calls do not allocate.
    namespace
    {
        struct DevBuf
        {
            void *p = nullptr;
            ~DevBuf()
            {
                if (p)
                {
                    cudaFree(p);
                }
            }
            int alloc(size_t bytes)
            {
                VMV_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
                return VMV_OK;
            }
        };

        struct HostPathPool
        {
            std::mutex mutex;
            cudaStream_t streams[2] = {nullptr, nullptr};
            void *buf[3] = {nullptr, nullptr, nullptr};
            size_t cap[3] = {0, 0, 0};

            int ensure(int which, size_t bytes)
            {
                if (cap[which] < bytes)
                {
                    if (buf[which])
                    {
                        VMV_CUDA(cudaFree(buf[which]));
                        buf[which] = nullptr;
                        cap[which] = 0;
                    }
                    const size_t want = std::max<size_t>(bytes + bytes / 4, 1 << 20);
                    VMV_CUDA(cudaMalloc(&buf[which], want));
                    cap[which] = want;
                }
                return VMV_OK;
            }

            int init()
            {
                for (auto &st : streams)
                {
                    if (st == nullptr)
                    {
                        VMV_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
                    }
                }
                return VMV_OK;
            }
        };
        HostPathPool g_pools[kMaxDevices];

        // No return -- error or not -- leaves a copy in flight on the pool's streams: the caller may free or
        // reuse q / a / b / bits as soon as the call is back.
        struct PoolDrain
        {
            HostPathPool &pool;
            ~PoolDrain()
            {
                for (cudaStream_t st : pool.streams)
                {
                    if (st != nullptr)
                    {
                        cudaStreamSynchronize(st);
                    }
                }
            }
        };

        // multiple of 32: chunks own whole verdict words.  VMV_CHUNK_LOG2 overrides (tuning aid).
        size_t chunk_units()
        {
            static const size_t v = []
            {
                const char *e = std::getenv("VMV_CHUNK_LOG2");
                const int lg = e ? std::atoi(e) : 18;
                return size_t(1) << std::min(std::max(lg, 10), 26);
            }();
            return v;
        }
    }  // namespace

    int vmv_validate_configs(int robot, const vmv_env *env, const float *q, size_t n, uint32_t *bits)
    {
        return guarded("vmv_validate_configs", [&]() -> int
        {
            if (!valid_robot(robot) || (n > 0 && (q == nullptr || bits == nullptr)))
            {
                return fail(VMV_ERR_ARG, "vmv_validate_configs: bad argument");
            }
            if (n == 0)
            {
                vmv::LaunchEnv probe{};
                return make_launch_env(robot_host(robot), env, probe);
            }
            int device = 0;
            VMV_CUDA(cudaGetDevice(&device));
            HostPathPool &pool = g_pools[device % kMaxDevices];
            std::lock_guard<std::mutex> lock(pool.mutex);
            const PoolDrain drain{pool};
            const size_t dof = robot_host(robot).dof;
            int rc = pool.init();
            if (rc == VMV_OK)
            {
                rc = pool.ensure(0, n * dof * sizeof(float));
            }
            if (rc == VMV_OK)
            {
                rc = pool.ensure(2, ((n + 31) / 32) * 4);
            }
            if (rc != VMV_OK)
            {
                return rc;
            }
            float *dq = static_cast<float *>(pool.buf[0]);
            uint32_t *dw = static_cast<uint32_t *>(pool.buf[2]);
            int k = 0;
            const size_t chunk_cfg = chunk_units();
            for (size_t off = 0; off < n; off += chunk_cfg, ++k)
            {
                const size_t cnt = std::min(chunk_cfg, n - off);
                cudaStream_t st = pool.streams[k & 1];
                VMV_CUDA(cudaMemcpyAsync(dq + off * dof, q + off * dof, cnt * dof * sizeof(float), cudaMemcpyHostToDevice, st));
                rc = vmv_validate_configs_dev(robot, env, dq + off * dof, cnt, dw + off / 32, st);
                if (rc != VMV_OK)
                {
                    return rc;
                }
                VMV_CUDA(cudaMemcpyAsync(bits + off / 32, dw + off / 32, ((cnt + 31) / 32) * 4, cudaMemcpyDeviceToHost, st));
            }
            VMV_CUDA(cudaStreamSynchronize(pool.streams[0]));
            VMV_CUDA(cudaStreamSynchronize(pool.streams[1]));
            return VMV_OK;
        });
    }

    int vmv_validate_edges(int robot, const vmv_env *env, const float *a, const float *b, size_t n, int resolution, uint32_t *bits)
    {
        return guarded("vmv_validate_edges", [&]() -> int
        {
            if (!valid_robot(robot) || (n > 0 && (a == nullptr || b == nullptr || bits == nullptr)))
            {
                return fail(VMV_ERR_ARG, "vmv_validate_edges: bad argument");
            }
            if (n == 0)
            {
                vmv::LaunchEnv probe{};
                return make_launch_env(robot_host(robot), env, probe);
            }
            int device = 0;
            VMV_CUDA(cudaGetDevice(&device));
            HostPathPool &pool = g_pools[device % kMaxDevices];
            std::lock_guard<std::mutex> lock(pool.mutex);
            const PoolDrain drain{pool};
            const size_t dof = robot_host(robot).dof;
            int rc = pool.init();
            if (rc == VMV_OK)
            {
                rc = pool.ensure(0, n * dof * sizeof(float));
            }
            if (rc == VMV_OK)
            {
                rc = pool.ensure(1, n * dof * sizeof(float));
            }
            if (rc == VMV_OK)
            {
                rc = pool.ensure(2, ((n + 31) / 32) * 4);
            }
            if (rc != VMV_OK)
            {
                return rc;
            }
            float *da = static_cast<float *>(pool.buf[0]);
            float *db = static_cast<float *>(pool.buf[1]);
            uint32_t *dw = static_cast<uint32_t *>(pool.buf[2]);
            const size_t chunk = chunk_units() / 4;
            int k = 0;
            for (size_t off = 0; off < n; off += chunk, ++k)
            {
                const size_t cnt = std::min(chunk, n - off);
                cudaStream_t st = pool.streams[k & 1];
                VMV_CUDA(cudaMemcpyAsync(da + off * dof, a + off * dof, cnt * dof * sizeof(float), cudaMemcpyHostToDevice, st));
                VMV_CUDA(cudaMemcpyAsync(db + off * dof, b + off * dof, cnt * dof * sizeof(float), cudaMemcpyHostToDevice, st));
                rc = vmv_validate_edges_dev(robot, env, da + off * dof, db + off * dof, cnt, resolution, dw + off / 32, st);
                if (rc != VMV_OK)
                {
                    return rc;
                }
                VMV_CUDA(cudaMemcpyAsync(bits + off / 32, dw + off / 32, ((cnt + 31) / 32) * 4, cudaMemcpyDeviceToHost, st));
            }
            VMV_CUDA(cudaStreamSynchronize(pool.streams[0]));
            VMV_CUDA(cudaStreamSynchronize(pool.streams[1]));
            return VMV_OK;
        });
    }

    int vmv_validate_edges_indexed(int robot, const vmv_env *env, const float *vertices, size_t n_vertices, const uint32_t *pairs, size_t n_edges, int resolution, uint32_t *bits)
    {
        return guarded("vmv_validate_edges_indexed", [&]() -> int
        {
            if (!valid_robot(robot) || (n_edges > 0 && (vertices == nullptr || pairs == nullptr || bits == nullptr)))
            {
                return fail(VMV_ERR_ARG, "vmv_validate_edges_indexed: bad argument");
            }
            if (n_edges == 0)
            {
                vmv::LaunchEnv probe{};
                return make_launch_env(robot_host(robot), env, probe);
            }
            for (size_t i = 0; i < 2 * n_edges; ++i)
            {
                if (pairs[i] >= n_vertices)
                {
                    return fail(VMV_ERR_ARG, "vmv_validate_edges_indexed: vertex index out of range");
                }
            }
            int device = 0;
            VMV_CUDA(cudaGetDevice(&device));
            HostPathPool &pool = g_pools[device % kMaxDevices];
            std::lock_guard<std::mutex> lock(pool.mutex);
            const PoolDrain drain{pool};
            const size_t dof = robot_host(robot).dof;
            int rc = pool.init();
            if (rc == VMV_OK)
            {
                rc = pool.ensure(0, n_vertices * dof * sizeof(float));
            }
            if (rc == VMV_OK)
            {
                rc = pool.ensure(1, n_edges * 2 * sizeof(uint32_t));
            }
            if (rc == VMV_OK)
            {
                rc = pool.ensure(2, ((n_edges + 31) / 32) * 4);
            }
            if (rc != VMV_OK)
            {
                return rc;
            }
            cudaStream_t st = pool.streams[0];
            VMV_CUDA(cudaMemcpyAsync(pool.buf[0], vertices, n_vertices * dof * sizeof(float), cudaMemcpyHostToDevice, st));
            VMV_CUDA(cudaMemcpyAsync(pool.buf[1], pairs, n_edges * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            rc = vmv_validate_edges_indexed_dev(robot, env, static_cast<const float *>(pool.buf[0]), n_vertices, static_cast<const uint32_t *>(pool.buf[1]),
                                                n_edges, resolution, static_cast<uint32_t *>(pool.buf[2]), st);
            if (rc != VMV_OK)
            {
                return rc;
            }
            VMV_CUDA(cudaMemcpyAsync(bits, pool.buf[2], ((n_edges + 31) / 32) * 4, cudaMemcpyDeviceToHost, st));
            VMV_CUDA(cudaStreamSynchronize(st));
            return VMV_OK;
        });
    }

    // ---- CenterVox pointcloud filter (vmv_filter.cuh; reference collision/filter_centervox.hh) -------
    namespace
    {
        struct FilterPool
        {
            std::mutex mutex;
            cudaStream_t stream = nullptr;
            void *buf[4] = {nullptr, nullptr, nullptr, nullptr};  // points, keys, first indices, records + count
            size_t cap[4] = {0, 0, 0, 0};

            int ensure(int which, size_t bytes)
            {
                if (cap[which] < bytes)
                {
                    if (buf[which])
                    {
                        VMV_CUDA(cudaFree(buf[which]));
                        buf[which] = nullptr;
                        cap[which] = 0;
                    }
                    VMV_CUDA(cudaMalloc(&buf[which], bytes));
                    cap[which] = bytes;
                }
                return VMV_OK;
            }
        };
        FilterPool g_filter_pools[kMaxDevices];
        constexpr uint32_t kCenterVoxPoolMax = 32768;  // filter_centervox.hh:118

        int centervox_run(const float *pts, bool pts_on_device, size_t n, float voxel_size, float max_range, const float *origin, const float *ws_min,
                          const float *ws_max, uint32_t *out_indices, size_t cap_out, size_t *n_out)
        {
            if (n_out == nullptr || origin == nullptr || ws_min == nullptr || ws_max == nullptr || (n > 0 && pts == nullptr) ||
                (cap_out > 0 && out_indices == nullptr))
            {
                return fail(VMV_ERR_ARG, "vmv_filter_pointcloud_centervox: bad argument");
            }
            *n_out = 0;
            if (n == 0)
            {
                return VMV_OK;  // filter_centervox.hh:312-314
            }
            if (n >= 0xffffffffull)
            {
                return fail(VMV_ERR_LIMIT, "vmv_filter_pointcloud_centervox: more than 2^32 - 2 points");
            }
            // CenterSelectiveVoxelFilter's constructor (filter_centervox.hh:95-121), in its arithmetic
            const float width = std::max({ws_max[0] - ws_min[0], ws_max[1] - ws_min[1], ws_max[2] - ws_min[2]});
            if (!(voxel_size > 0.F) || !(width > 0.F) || !std::isfinite(width / voxel_size))
            {
                return fail(VMV_ERR_ARG, "vmv_filter_pointcloud_centervox: voxel size and workspace must be positive and finite");
            }
            const int grid_width = std::min(255, static_cast<int>(std::ceil(width / voxel_size)));
            vmv::CenterVoxParams P{};
            for (int k = 0; k < 3; ++k)
            {
                P.origin[k] = origin[k], P.ws_min[k] = ws_min[k], P.ws_max[k] = ws_max[k];
            }
            P.voxel_size = voxel_size;
            P.inv_scale = grid_width / width;
            P.max_range_sq = max_range * max_range;
            P.dim = std::min(grid_width + 1, 255);
            const float per_dim = width / voxel_size;
            const float estimate = std::pow(per_dim, 3.0F) * 0.05F;
            const size_t pool_size = estimate >= static_cast<float>(kCenterVoxPoolMax) ? kCenterVoxPoolMax : static_cast<size_t>(estimate);

            int device = 0;
            VMV_CUDA(cudaGetDevice(&device));
            FilterPool &pool = g_filter_pools[device % kMaxDevices];
            std::lock_guard<std::mutex> lock(pool.mutex);
            if (pool.stream == nullptr)
            {
                VMV_CUDA(cudaStreamCreateWithFlags(&pool.stream, cudaStreamNonBlocking));
            }
            cudaStream_t st = pool.stream;
            const size_t n_vox = static_cast<size_t>(P.dim) * P.dim * P.dim;
            const uint32_t rec_cap = kCenterVoxPoolMax;
            int rc = VMV_OK;
            if (!pts_on_device)
            {
                rc = pool.ensure(0, n * 3 * sizeof(float));
            }
            if (rc == VMV_OK)
            {
                rc = pool.ensure(1, n_vox * sizeof(unsigned long long));
            }
            if (rc == VMV_OK)
            {
                rc = pool.ensure(2, n_vox * sizeof(uint32_t));
            }
            if (rc == VMV_OK)
            {
                rc = pool.ensure(3, (3 * static_cast<size_t>(rec_cap) + 4) * sizeof(uint32_t));
            }
            if (rc != VMV_OK)
            {
                return rc;
            }
            const float *d_pts = pts;
            if (!pts_on_device)
            {
                VMV_CUDA(cudaMemcpyAsync(pool.buf[0], pts, n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
                d_pts = static_cast<const float *>(pool.buf[0]);
            }
            auto *d_best = static_cast<unsigned long long *>(pool.buf[1]);
            auto *d_first = static_cast<uint32_t *>(pool.buf[2]);
            auto *d_rec = static_cast<uint32_t *>(pool.buf[3]);
            uint32_t *d_count = d_rec + 3 * static_cast<size_t>(rec_cap);
            VMV_CUDA(cudaMemsetAsync(d_best, 0xff, n_vox * sizeof(unsigned long long), st));
            VMV_CUDA(cudaMemsetAsync(d_first, 0xff, n_vox * sizeof(uint32_t), st));
            VMV_CUDA(cudaMemsetAsync(d_count, 0, sizeof(uint32_t), st));
            {
                const size_t blocks = (n + 255) / 256;
                const unsigned grid = static_cast<unsigned>(std::min<size_t>(blocks, static_cast<size_t>(sm_count()) * 8));
                vmv::k_centervox_insert<<<grid, 256, 0, st>>>(P, d_pts, static_cast<uint32_t>(n), d_best, d_first);
                g_launches++;
                VMV_CUDA(cudaGetLastError());
                vmv::k_centervox_compact<<<static_cast<unsigned>((n_vox + 255) / 256), 256, 0, st>>>(P, d_pts, d_best, d_first, n_vox, rec_cap, d_rec, d_count);
                g_launches++;
                VMV_CUDA(cudaGetLastError());
            }
            uint32_t count = 0;
            const auto tm0 = std::chrono::steady_clock::now();
            VMV_CUDA(cudaMemcpyAsync(&count, d_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            VMV_CUDA(cudaStreamSynchronize(st));
            const auto tm1 = std::chrono::steady_clock::now();
            if (count > pool_size)
            {
                // the reference throws std::runtime_error("Voxel pool exhausted") (filter_centervox.hh:130-133)
                return fail(VMV_ERR_LIMIT, "vmv_filter_pointcloud_centervox: voxel pool exhausted (the reference throws here): more occupied voxels "
                                           "than min((width / voxel_size)^3 * 0.05, 32768)");
            }
            std::vector<uint32_t> rec(3 * static_cast<size_t>(count));
            if (count > 0)
            {
                VMV_CUDA(cudaMemcpy(rec.data(), d_rec, rec.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
            }
            // extract_points (filter_centervox.hh:167-184) walks the sparse tables in creation order: x slabs by
            // their first point, (x, y) columns inside a slab by theirs, voxels inside a column by theirs.
            // Ranks of the <= 255 slabs and of the occupied columns, then one LSD radix sort of the voxels on
            // (column rank, first point) -- comparison sorts with a decoding comparator cost milliseconds here.
            const uint32_t dim = static_cast<uint32_t>(P.dim);
            std::vector<uint32_t> first_x(dim, 0xffffffffu), first_xy(static_cast<size_t>(dim) * dim, 0xffffffffu), col_of(count);
            for (uint32_t k = 0; k < count; ++k)
            {
                const uint32_t v = rec[3 * k], f = rec[3 * k + 2];
                const uint32_t col = v / dim, vx = col / dim;
                col_of[k] = col;
                first_x[vx] = std::min(first_x[vx], f);
                first_xy[col] = std::min(first_xy[col], f);
            }
            std::vector<uint64_t> slabs, cols;
            for (uint32_t vx = 0; vx < dim; ++vx)
            {
                if (first_x[vx] != 0xffffffffu)
                {
                    slabs.push_back((static_cast<uint64_t>(first_x[vx]) << 32) | vx);
                }
            }
            std::sort(slabs.begin(), slabs.end());
            std::vector<uint32_t> rank_x(dim, 0), rank_col(static_cast<size_t>(dim) * dim, 0);
            for (size_t r = 0; r < slabs.size(); ++r)
            {
                rank_x[slabs[r] & 0xffffffffu] = static_cast<uint32_t>(r);
            }
            std::vector<uint32_t> col_ids;
            for (uint32_t c = 0; c < dim * dim; ++c)
            {
                if (first_xy[c] != 0xffffffffu)
                {
                    cols.push_back((static_cast<uint64_t>(rank_x[c / dim]) << 32) | first_xy[c]);
                    col_ids.push_back(c);
                }
            }
            {
                // first points are distinct across columns: sort the keys, find each column by its first point
                std::vector<uint64_t> sorted(cols);
                std::sort(sorted.begin(), sorted.end());
                for (size_t k = 0; k < cols.size(); ++k)
                {
                    rank_col[col_ids[k]] = static_cast<uint32_t>(std::lower_bound(sorted.begin(), sorted.end(), cols[k]) - sorted.begin());
                }
            }
            struct Item
            {
                uint64_t key;
                uint32_t keep;
            };
            std::vector<Item> items(count), scratch(count);
            for (uint32_t k = 0; k < count; ++k)
            {
                items[k] = Item{(static_cast<uint64_t>(rank_col[col_of[k]]) << 32) | rec[3 * k + 2], rec[3 * k + 1]};
            }
            std::vector<uint32_t> hist(65537);
            for (int pass = 0; pass < 3; ++pass)  // 48 significant bits: first point (32) + column rank (<= 16)
            {
                std::fill(hist.begin(), hist.end(), 0u);
                const int shift = 16 * pass;
                for (const Item &it : items)
                {
                    hist[((it.key >> shift) & 0xffffu) + 1]++;
                }
                for (int h = 0; h < 65536; ++h)
                {
                    hist[h + 1] += hist[h];
                }
                for (const Item &it : items)
                {
                    scratch[hist[(it.key >> shift) & 0xffffu]++] = it;
                }
                items.swap(scratch);
            }
            *n_out = count;
            for (uint32_t k = 0; k < count && k < cap_out; ++k)
            {
                out_indices[k] = items[k].keep;
            }
            if (std::getenv("VMV_FILTER_TIMING"))
            {
                const auto tm2 = std::chrono::steady_clock::now();
                std::fprintf(stderr, "centervox: device %.3f ms (wait), records + ordering %.3f ms, %u voxels, table %zu^3\n",
                             std::chrono::duration<double, std::milli>(tm1 - tm0).count(), std::chrono::duration<double, std::milli>(tm2 - tm1).count(),
                             count, static_cast<size_t>(P.dim));
            }
            return VMV_OK;
        }
    }  // namespace

    int vmv_filter_pointcloud_centervox(const float *points, size_t n, float voxel_size, float max_range, const float *origin, const float *workspace_min,
                                        const float *workspace_max, uint32_t *kept_indices, size_t cap, size_t *n_kept)
    {
        return guarded("vmv_filter_pointcloud_centervox", [&]() -> int
        {
            return centervox_run(points, false, n, voxel_size, max_range, origin, workspace_min, workspace_max, kept_indices, cap, n_kept);
        });
    }

    int vmv_filter_pointcloud_centervox_dev(const float *d_points, size_t n, float voxel_size, float max_range, const float *origin,
                                            const float *workspace_min, const float *workspace_max, uint32_t *kept_indices, size_t cap, size_t *n_kept)
    {
        return guarded("vmv_filter_pointcloud_centervox_dev", [&]() -> int
        {
            return centervox_run(d_points, true, n, voxel_size, max_range, origin, workspace_min, workspace_max, kept_indices, cap, n_kept);
        });
    }

    // ---- Halton sampler on the device (vmv_halton.cuh; reference random/halton.hh) ------------------
    extern "C++"
    {
    namespace
    {
        // largest sample count for which every joint's numerator / denominator stay below 2^24 (the regime
        // where the reference's f32 recurrence is exactly the radical inverse) inside the first epoch
        uint64_t halton_exact_limit(int dof)
        {
            uint64_t limit = 1000000;  // Halton::max_iterations (halton.hh:12): bases rotate afterwards
            for (int j = 0; j < dof; ++j)
            {
                const uint64_t b = vmv::halton_prime(j);
                uint64_t pw = 1;  // largest power of b below 2^24: indices up to pw - 1 keep den <= pw
                while (pw * b < (1ull << 24))
                {
                    pw *= b;
                }
                limit = std::min(limit, pw - 1);
            }
            return limit;
        }

        template <int DOF>
        int launch_halton(const vmv::HaltonScale &sc, uint64_t first, size_t n, float *d_q, cudaStream_t s)
        {
            const size_t blocks = (n + 255) / 256;
            const unsigned grid = static_cast<unsigned>(std::min<size_t>(blocks, static_cast<size_t>(sm_count()) * 8));
            vmv::k_halton_fill<DOF><<<grid, 256, 0, s>>>(sc, first, n, d_q);
            g_launches++;
            VMV_CUDA(cudaGetLastError());
            return VMV_OK;
        }
    }  // namespace
    }

    int vmv_halton_fill_dev(int robot, uint64_t first, size_t n, float *d_q, void *stream)
    {
        return guarded("vmv_halton_fill_dev", [&]() -> int
        {
            if (!valid_robot(robot) || (n > 0 && d_q == nullptr))
            {
                return fail(VMV_ERR_ARG, "vmv_halton_fill_dev: bad argument");
            }
            const int dof = robot_host(robot).dof;
            if (first + n > halton_exact_limit(dof))
            {
                return fail(VMV_ERR_LIMIT, "vmv_halton_fill_dev: samples beyond the exact range of the reference's f32 recurrence "
                                           "(first epoch, base^digits < 2^24); draw those on the host");
            }
            if (n == 0)
            {
                return VMV_OK;
            }
            vmv::HaltonScale sc{};
            for (int j = 0; j < dof; ++j)
            {
                sc.lower[j] = robot_host(robot).lower[j];
                sc.range[j] = robot_host(robot).range[j];
            }
            cudaStream_t s = static_cast<cudaStream_t>(stream);
            switch (dof)
            {
                case 6:
                    return launch_halton<6>(sc, first, n, d_q, s);
                case 7:
                    return launch_halton<7>(sc, first, n, d_q, s);
                case 8:
                    return launch_halton<8>(sc, first, n, d_q, s);
                case 14:
                    return launch_halton<14>(sc, first, n, d_q, s);
                default:
                    return fail(VMV_ERR_ARG, "vmv_halton_fill_dev: unsupported joint count");
            }
        });
    }

    uint64_t vmv_halton_exact_limit(int robot)
    {
        return valid_robot(robot) ? halton_exact_limit(robot_host(robot).dof) : 0;
    }

    // samples first .. first+n-1 generated and validated on the device; only the verdict bits (and, if
    // asked for, the configurations) cross PCIe
    int vmv_validate_halton(int robot, const vmv_env *env, uint64_t first, size_t n, uint32_t *bits, float *q_out)
    {
        return guarded("vmv_validate_halton", [&]() -> int
        {
            if (!valid_robot(robot) || (n > 0 && bits == nullptr))
            {
                return fail(VMV_ERR_ARG, "vmv_validate_halton: bad argument");
            }
            if (n == 0)
            {
                vmv::LaunchEnv probe{};
                return make_launch_env(robot_host(robot), env, probe);
            }
            int device = 0;
            VMV_CUDA(cudaGetDevice(&device));
            HostPathPool &pool = g_pools[device % kMaxDevices];
            std::lock_guard<std::mutex> lock(pool.mutex);
            const PoolDrain drain{pool};
            const size_t dof = robot_host(robot).dof;
            int rc = pool.init();
            if (rc == VMV_OK)
            {
                rc = pool.ensure(0, n * dof * sizeof(float));
            }
            if (rc == VMV_OK)
            {
                rc = pool.ensure(2, ((n + 31) / 32) * 4);
            }
            if (rc != VMV_OK)
            {
                return rc;
            }
            float *dq = static_cast<float *>(pool.buf[0]);
            uint32_t *dw = static_cast<uint32_t *>(pool.buf[2]);
            cudaStream_t st = pool.streams[0];
            rc = vmv_halton_fill_dev(robot, first, n, dq, st);
            if (rc == VMV_OK)
            {
                rc = vmv_validate_configs_dev(robot, env, dq, n, dw, st);
            }
            if (rc != VMV_OK)
            {
                return rc;
            }
            VMV_CUDA(cudaMemcpyAsync(bits, dw, ((n + 31) / 32) * 4, cudaMemcpyDeviceToHost, st));
            if (q_out != nullptr)
            {
                VMV_CUDA(cudaMemcpyAsync(q_out, dq, n * dof * sizeof(float), cudaMemcpyDeviceToHost, st));
            }
            VMV_CUDA(cudaStreamSynchronize(st));
            return VMV_OK;
        });
    }

    int vmv_filter_points_dev(int robot, const vmv_env *env, const float *d_q, const float *d_points, size_t n, float point_radius, uint32_t *d_keep_bits, void *stream)
    {
        return guarded("vmv_filter_points_dev", [&]() -> int
        {
            if (!valid_robot(robot) || d_q == nullptr || (n > 0 && (d_points == nullptr || d_keep_bits == nullptr)))
            {
                return fail(VMV_ERR_ARG, "vmv_filter_points_dev: bad argument");
            }
            vmv::LaunchEnv le{};
            int rc = make_launch_env(robot_host(robot), env, le);
            if (rc != VMV_OK || n == 0)
            {
                return rc;
            }
            vmv::RobotDev rd{};
            rc = robot_tables(robot, rd);
            if (rc != VMV_OK)
            {
                return rc;
            }
            cudaStream_t s = static_cast<cudaStream_t>(stream);
            return ops(robot).filter(rd, le, d_q, d_points, n, point_radius, d_keep_bits, s);
        });
    }

    int vmv_filter_self_from_pointcloud(int robot, const vmv_env *env, const float *q, const float *points, size_t n, float point_radius, uint32_t *keep_bits)
    {
        return guarded("vmv_filter_self_from_pointcloud", [&]() -> int
        {
            if (!valid_robot(robot) || q == nullptr || (n > 0 && (points == nullptr || keep_bits == nullptr)))
            {
                return fail(VMV_ERR_ARG, "vmv_filter_self_from_pointcloud: bad argument");
            }
            if (n == 0)
            {
                vmv::LaunchEnv probe{};
                return make_launch_env(robot_host(robot), env, probe);
            }
            const size_t dof = robot_host(robot).dof, words = (n + 31) / 32;
            DevBuf dq, dp, db;
            int rc = dq.alloc(dof * sizeof(float));
            if (rc == VMV_OK)
            {
                rc = dp.alloc(n * 3 * sizeof(float));
            }
            if (rc == VMV_OK)
            {
                rc = db.alloc(words * 4);
            }
            if (rc != VMV_OK)
            {
                return rc;
            }
            VMV_CUDA(cudaMemcpy(dq.p, q, dof * sizeof(float), cudaMemcpyHostToDevice));
            VMV_CUDA(cudaMemcpy(dp.p, points, n * 3 * sizeof(float), cudaMemcpyHostToDevice));
            rc = vmv_filter_points_dev(robot, env, static_cast<const float *>(dq.p), static_cast<const float *>(dp.p), n, point_radius,
                                       static_cast<uint32_t *>(db.p), nullptr);
            if (rc != VMV_OK)
            {
                return rc;
            }
            VMV_CUDA(cudaMemcpy(keep_bits, db.p, words * 4, cudaMemcpyDeviceToHost));
            return VMV_OK;
        });
    }

    int vmv_sphere_fk(int robot, const float *q, size_t n, float *xyzr)
    {
        return guarded("vmv_sphere_fk", [&]() -> int
        {
            if (!valid_robot(robot) || (n > 0 && (q == nullptr || xyzr == nullptr)))
            {
                return fail(VMV_ERR_ARG, "vmv_sphere_fk: bad argument");
            }
            if (n == 0)
            {
                return VMV_OK;
            }
            const size_t qbytes = n * robot_host(robot).dof * sizeof(float);
            const size_t obytes = n * robot_host(robot).n_spheres * 4 * sizeof(float);
            DevBuf dq, dout;
            int rc = dq.alloc(qbytes);
            if (rc == VMV_OK)
            {
                rc = dout.alloc(obytes);
            }
            if (rc != VMV_OK)
            {
                return rc;
            }
            VMV_CUDA(cudaMemcpyAsync(dq.p, q, qbytes, cudaMemcpyHostToDevice, nullptr));
            rc = vmv_sphere_fk_dev(robot, static_cast<const float *>(dq.p), n, static_cast<float *>(dout.p), nullptr);
            if (rc != VMV_OK)
            {
                return rc;
            }
            VMV_CUDA(cudaMemcpyAsync(xyzr, dout.p, obytes, cudaMemcpyDeviceToHost, nullptr));
            VMV_CUDA(cudaStreamSynchronize(nullptr));
            return VMV_OK;
        });
    }

    int vmv_debug(int robot, const vmv_env *env, const float *q, int32_t *env_hits, size_t cap_env, size_t *n_env, int32_t *self_hits, size_t cap_self, size_t *n_self)
    {
        return guarded("vmv_debug", [&]() -> int
        {
            if (!valid_robot(robot) || q == nullptr || n_env == nullptr || n_self == nullptr || (cap_env > 0 && env_hits == nullptr) ||
                (cap_self > 0 && self_hits == nullptr))
            {
                return fail(VMV_ERR_ARG, "vmv_debug: bad argument");
            }
            vmv::LaunchEnv le{};
            int rc = make_launch_env(robot_host(robot), env, le);
            if (rc != VMV_OK)
            {
                return rc;
            }
            vmv::RobotDev rd{};
            rc = robot_tables(robot, rd);
            if (rc != VMV_OK)
            {
                return rc;
            }
            const size_t dof = robot_host(robot).dof;
            DevBuf dq, de, ds, dc;
            rc = dq.alloc(dof * sizeof(float));
            rc = rc == VMV_OK ? de.alloc(2 * cap_env * sizeof(int32_t)) : rc;
            rc = rc == VMV_OK ? ds.alloc(2 * cap_self * sizeof(int32_t)) : rc;
            rc = rc == VMV_OK ? dc.alloc(2 * sizeof(uint32_t)) : rc;
            if (rc != VMV_OK)
            {
                return rc;
            }
            VMV_CUDA(cudaMemcpy(dq.p, q, dof * sizeof(float), cudaMemcpyHostToDevice));
            VMV_CUDA(cudaMemset(dc.p, 0, 2 * sizeof(uint32_t)));
            rc = ops(robot).debug(rd, le, static_cast<const float *>(dq.p), env->d_object_ids, static_cast<int32_t *>(de.p), static_cast<uint32_t>(cap_env),
                                  static_cast<int32_t *>(ds.p), static_cast<uint32_t>(cap_self), static_cast<uint32_t *>(dc.p));
            if (rc != VMV_OK)
            {
                return rc;
            }
            uint32_t counts[2] = {0, 0};
            VMV_CUDA(cudaMemcpy(counts, dc.p, sizeof(counts), cudaMemcpyDeviceToHost));
            *n_env = counts[0];
            *n_self = counts[1];
            if (cap_env > 0)
            {
                VMV_CUDA(cudaMemcpy(env_hits, de.p, 2 * std::min<size_t>(cap_env, counts[0]) * sizeof(int32_t), cudaMemcpyDeviceToHost));
            }
            if (cap_self > 0)
            {
                VMV_CUDA(cudaMemcpy(self_hits, ds.p, 2 * std::min<size_t>(cap_self, counts[1]) * sizeof(int32_t), cudaMemcpyDeviceToHost));
            }
            return VMV_OK;
        });
    }

    void *vmv_dev_alloc(size_t bytes)
    {
        void *p = nullptr;
        if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess)
        {
            g_error = "vmv_dev_alloc: cudaMalloc failed";
            cudaGetLastError();
            return nullptr;
        }
        return p;
    }

    void vmv_dev_free(void *p)
    {
        if (p)
        {
            cudaFree(p);
        }
    }

    void *vmv_host_alloc_pinned(size_t bytes)
    {
        void *p = nullptr;
        if (cudaMallocHost(&p, std::max<size_t>(bytes, 16)) != cudaSuccess)
        {
            g_error = "vmv_host_alloc_pinned: cudaMallocHost failed";
            cudaGetLastError();
            return nullptr;
        }
        return p;
    }

    void vmv_host_free_pinned(void *p)
    {
        if (p)
        {
            cudaFreeHost(p);
        }
    }

    int vmv_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream)
    {
        VMV_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
        return VMV_OK;
    }

    int vmv_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream)
    {
        VMV_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
        return VMV_OK;
    }

    int vmv_stream_sync(void *stream)
    {
        VMV_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
        return VMV_OK;
    }

    int vmv_dev_stats(uint64_t *out64, int reset)
    {
        unsigned long long *b = stats_buffer();
        VMV_CUDA(cudaDeviceSynchronize());
        if (out64 != nullptr)
        {
            VMV_CUDA(cudaMemcpy(out64, b, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        }
        if (reset)
        {
            VMV_CUDA(cudaMemset(b, 0, 64 * sizeof(unsigned long long)));
        }
        return VMV_OK;
    }

    int vmv_env_capt_digest(const vmv_env *env, int k, uint64_t *out4)
    {
        if (env == nullptr || out4 == nullptr || k < 0 || static_cast<size_t>(k) >= env->capts.size())
        {
            return fail(VMV_ERR_ARG, "vmv_env_capt_digest: bad argument");
        }
        const HCapt &t = env->capts[static_cast<size_t>(k)];
        {
            const int rc = capt_ensure_host(t);
            if (rc != VMV_OK)
            {
                return rc;
            }
        }
        auto fnv = [](const void *p, size_t bytes)
        {
            uint64_t h = 1469598103934665603ull;
            const unsigned char *b = static_cast<const unsigned char *>(p);
            for (size_t i = 0; i < bytes; ++i)
            {
                h = (h ^ b[i]) * 1099511628211ull;
            }
            return h;
        };
        out4[0] = fnv(t.nodes.data(), t.nodes.size() * sizeof(float));
        out4[1] = fnv(t.leafbits.data(), t.leafbits.size() * sizeof(uint32_t));
        out4[2] = fnv(t.gpts.data(), t.gpts.size() * sizeof(float));
        out4[3] = fnv(t.gstart.data(), t.gstart.size() * sizeof(uint32_t));
        return VMV_OK;
    }

    long vmv_env_capt_nodes(const vmv_env *env, int k, float *out, size_t cap_floats)
    {
        if (env == nullptr || k < 0 || static_cast<size_t>(k) >= env->capts.size())
        {
            return fail(VMV_ERR_ARG, "vmv_env_capt_nodes: bad argument");
        }
        const HCapt &t = env->capts[static_cast<size_t>(k)];
        const int rc = capt_ensure_host(t);
        if (rc != VMV_OK)
        {
            return rc;
        }
        if (out != nullptr)
        {
            std::memcpy(out, t.nodes.data(), std::min(cap_floats, t.nodes.size()) * sizeof(float));
        }
        return static_cast<long>(t.nodes.size());
    }

    uint64_t vmv_launch_count(void)
    {
        return g_launches.load();
    }

    int vmv_force_kernel_path(int path)
    {
        g_force_path.store(path);
        return VMV_OK;
    }
    // =========================================================================================
    // multi-GPU: one process per GPU of one node; NCCL for the plain collectives (loaded at run time, so
    // single-GPU users need no NCCL), CUDA IPC windows for the gather fused into the validation kernels
    // =========================================================================================
    extern "C++"
    {
    namespace
    {
        struct NcclId
        {
            char internal[128];
        };
        struct NcclApi
        {
            void *lib = nullptr;
            int (*GetUniqueId)(NcclId *) = nullptr;
            int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
            int (*CommDestroy)(void *) = nullptr;
            int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
            int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
            const char *(*GetErrorString)(int) = nullptr;
        };
        constexpr int kNcclUint8 = 1, kNcclUint32 = 3;  // ncclDataType_t (nccl.h)

        int nccl_api(const NcclApi *&out)
        {
            static NcclApi api;
            static bool tried = false;
            std::lock_guard<std::mutex> lock(g_mutex);
            if (!tried)
            {
                tried = true;
                // a process that already holds an NCCL (torch's bundled copy) gets that one: same soname
                void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
                if (lib == nullptr)
                {
                    lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
                }
                if (lib != nullptr)
                {
                    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
                    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
                    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
                    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(lib, "ncclAllGather"));
                    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(lib, "ncclBroadcast"));
                    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
                    if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Broadcast && api.GetErrorString)
                    {
                        api.lib = lib;
                    }
                }
            }
            if (api.lib == nullptr)
            {
                return fail(VMV_ERR_STATE, "NCCL is not available (libnccl.so.2 could not be loaded)");
            }
            out = &api;
            return VMV_OK;
        }

        int nccl_fail(const NcclApi &api, int code, const char *what)
        {
            return fail(VMV_ERR_CUDA, std::string(what) + ": " + api.GetErrorString(code));
        }

#define VMV_NCCL(api, call)                          \
    do                                               \
    {                                                \
        const int c_ = (call);                       \
        if (c_ != 0)                                 \
        {                                            \
            return nccl_fail(api, c_, #call);        \
        }                                            \
    } while (0)

        constexpr int kMaxSlots = 4;

        // cuStreamBatchMemOp of the driver API (libcuda, loaded at run time): the flag waits of vmv_comm_wait as stream
        // memory operations -- executed by the front end, no kernel launch and no SM -- where the driver offers them
        using BatchMemOpFn = int (*)(cudaStream_t, unsigned int, void *, unsigned int);
        BatchMemOpFn batch_mem_op()
        {
            static BatchMemOpFn fn = []() -> BatchMemOpFn
            {
                if (std::getenv("VMV_COMM_WAIT_KERNEL") != nullptr)
                {
                    return nullptr;  // tuning aid: force the kernel
                }
                void *lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
                if (lib == nullptr)
                {
                    return nullptr;
                }
                void *f = dlsym(lib, "cuStreamBatchMemOp_v2");
                return reinterpret_cast<BatchMemOpFn>(f);
            }();
            return fn;
        }
        // CUstreamBatchMemOpParams (cuda.h): a union padded to 6 x 8 bytes; the wait-value member
        struct WaitValueOp
        {
            unsigned int operation;  // CU_STREAM_MEM_OP_WAIT_VALUE_32 = 1
            unsigned int pad0;
            unsigned long long address;
            unsigned long long value;  // low 32 bits
            unsigned int flags;        // CU_STREAM_WAIT_VALUE_GEQ = 0: (int32_t)(*addr - value) >= 0
            unsigned int pad1;
            unsigned long long alias;
            unsigned long long pad2;
        };
        static_assert(sizeof(WaitValueOp) == 48, "CUstreamBatchMemOpParams is 6 x 8 bytes");

        // byte stream of an environment's host-side description (vmv_env_broadcast)
        struct Writer
        {
            std::vector<unsigned char> b;
            template <typename T>
            void pod(const T &v)
            {
                const unsigned char *p = reinterpret_cast<const unsigned char *>(&v);
                b.insert(b.end(), p, p + sizeof(T));
            }
            template <typename T, typename A>
            void vec(const std::vector<T, A> &v)
            {
                pod<uint64_t>(v.size());
                const unsigned char *p = reinterpret_cast<const unsigned char *>(v.data());
                b.insert(b.end(), p, p + v.size() * sizeof(T));
            }
        };
        struct Reader
        {
            const unsigned char *p, *end;
            bool ok = true;
            template <typename T>
            void pod(T &v)
            {
                if (static_cast<size_t>(end - p) < sizeof(T))
                {
                    ok = false;
                    return;
                }
                std::memcpy(&v, p, sizeof(T));
                p += sizeof(T);
            }
            template <typename T, typename A>
            void vec(std::vector<T, A> &v)
            {
                uint64_t n = 0;
                pod(n);
                if (!ok || static_cast<uint64_t>(end - p) < n * sizeof(T))
                {
                    ok = false;
                    return;
                }
                v.resize(n);
                std::memcpy(v.data(), p, n * sizeof(T));
                p += n * sizeof(T);
            }
        };
    }  // namespace
    }

}

struct vmv_comm
{
    int rank = 0, world = 1, device = 0;
    void *nccl = nullptr;
    size_t words_per_rank = 0;
    int slots = 0;
    size_t flags_off = 0, window_words = 0;  // in words from the window base
    uint32_t *window = nullptr;
    uint32_t *peer_window[vmv::kMaxPeers] = {};
    uint32_t seq[4] = {0, 0, 0, 0};
    // copy-engine publication (vmv_comm_publish): a side stream, an event per slot, and a device table of
    // sequence numbers the flag copies read from
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ready[4] = {};
    cudaEvent_t copied[4] = {};  // recorded on the copy stream behind a slot's publication
    bool published[4] = {false, false, false, false};
    uint32_t *d_seq = nullptr;
};
constexpr uint32_t kSeqTable = 4096;

namespace
{
    GatherDev gather_for(vmv_comm *c, int slot)
    {
        GatherDev g{};
        g.world = c->world;
        g.rank = c->rank;
        g.seq = ++c->seq[slot];
        const size_t row = static_cast<size_t>(slot) * c->world * c->words_per_rank + static_cast<size_t>(c->rank) * c->words_per_rank;
        for (int p = 0; p < c->world; ++p)
        {
            g.peer_bits[p] = c->peer_window[p] + row;
            g.peer_flag[p] = c->peer_window[p] + c->flags_off + static_cast<size_t>(slot) * vmv::kMaxPeers + c->rank;
        }
        return g;
    }

    int check_gather_call(const vmv_comm *c, int slot, size_t n_units)
    {
        if (c == nullptr || c->window == nullptr)
        {
            return fail(VMV_ERR_STATE, "no gather window (call vmv_comm_window)");
        }
        if (slot < 0 || slot >= c->slots)
        {
            return fail(VMV_ERR_ARG, "gather slot out of range");
        }
        if ((n_units + 31) / 32 > c->words_per_rank)
        {
            return fail(VMV_ERR_ARG, "batch larger than the window's words_per_rank * 32 units");
        }
        int device = 0;
        VMV_CUDA(cudaGetDevice(&device));
        if (device != c->device)
        {
            return fail(VMV_ERR_STATE, "communicator belongs to another device");
        }
        return VMV_OK;
    }
}  // namespace

extern "C"
{
    int vmv_comm_unique_id(unsigned char *id128)
    {
        return guarded("vmv_comm_unique_id", [&]() -> int
        {
            if (id128 == nullptr)
            {
                return fail(VMV_ERR_ARG, "vmv_comm_unique_id: null argument");
            }
            const NcclApi *api = nullptr;
            int rc = nccl_api(api);
            if (rc != VMV_OK)
            {
                return rc;
            }
            NcclId id{};
            VMV_NCCL(*api, api->GetUniqueId(&id));
            std::memcpy(id128, id.internal, sizeof(id.internal));
            return VMV_OK;
        });
    }

    int vmv_comm_create(vmv_comm **out, const unsigned char *id128, int rank, int world)
    {
        return guarded("vmv_comm_create", [&]() -> int
        {
            if (out == nullptr || id128 == nullptr || world < 1 || world > vmv::kMaxPeers || rank < 0 || rank >= world)
            {
                return fail(VMV_ERR_ARG, "vmv_comm_create: bad argument (1 <= world <= 8, 0 <= rank < world)");
            }
            const NcclApi *api = nullptr;
            int rc = nccl_api(api);
            if (rc != VMV_OK)
            {
                return rc;
            }
            vmv_comm *c = new vmv_comm();
            c->rank = rank, c->world = world;
            VMV_CUDA(cudaGetDevice(&c->device));
            NcclId id{};
            std::memcpy(id.internal, id128, sizeof(id.internal));
            const int code = api->CommInitRank(&c->nccl, world, id, rank);
            if (code != 0)
            {
                delete c;
                return nccl_fail(*api, code, "ncclCommInitRank");
            }
            *out = c;
            return VMV_OK;
        });
    }

    void vmv_comm_destroy(vmv_comm *c)
    {
        if (c == nullptr)
        {
            return;
        }
        for (int p = 0; p < c->world; ++p)
        {
            if (p != c->rank && c->peer_window[p] != nullptr)
            {
                cudaIpcCloseMemHandle(c->peer_window[p]);
            }
        }
        if (c->window != nullptr)
        {
            cudaFree(c->window);
        }
        if (c->copy_stream != nullptr)
        {
            cudaStreamDestroy(c->copy_stream);
        }
        for (cudaEvent_t e : c->ready)
        {
            if (e != nullptr)
            {
                cudaEventDestroy(e);
            }
        }
        for (cudaEvent_t e : c->copied)
        {
            if (e != nullptr)
            {
                cudaEventDestroy(e);
            }
        }
        if (c->d_seq != nullptr)
        {
            cudaFree(c->d_seq);
        }
        const NcclApi *api = nullptr;
        if (c->nccl != nullptr && nccl_api(api) == VMV_OK)
        {
            api->CommDestroy(c->nccl);
        }
        delete c;
    }

    int vmv_comm_rank(const vmv_comm *c)
    {
        return c ? c->rank : fail(VMV_ERR_ARG, "vmv_comm_rank: null argument");
    }

    int vmv_comm_world(const vmv_comm *c)
    {
        return c ? c->world : fail(VMV_ERR_ARG, "vmv_comm_world: null argument");
    }

    int vmv_allgather_bits(vmv_comm *c, const uint32_t *d_local, size_t words_per_rank, uint32_t *d_global, void *stream)
    {
        return guarded("vmv_allgather_bits", [&]() -> int
        {
            if (c == nullptr || (words_per_rank > 0 && (d_local == nullptr || d_global == nullptr)))
            {
                return fail(VMV_ERR_ARG, "vmv_allgather_bits: bad argument");
            }
            const NcclApi *api = nullptr;
            int rc = nccl_api(api);
            if (rc != VMV_OK)
            {
                return rc;
            }
            VMV_NCCL(*api, api->AllGather(d_local, d_global, words_per_rank, kNcclUint32, c->nccl, static_cast<cudaStream_t>(stream)));
            return VMV_OK;
        });
    }

    int vmv_comm_window(vmv_comm *c, size_t words_per_rank, int slots)
    {
        return guarded("vmv_comm_window", [&]() -> int
        {
            if (c == nullptr || words_per_rank == 0 || slots < 1 || slots > kMaxSlots)
            {
                return fail(VMV_ERR_ARG, "vmv_comm_window: bad argument (1 <= slots <= 4)");
            }
            if (c->window != nullptr)
            {
                return fail(VMV_ERR_STATE, "vmv_comm_window: the communicator already has a window");
            }
            const NcclApi *api = nullptr;
            int rc = nccl_api(api);
            if (rc != VMV_OK)
            {
                return rc;
            }
            // rows are 16-byte aligned so that the push kernel and consumers may use wide accesses
            c->words_per_rank = (words_per_rank + 3) & ~size_t(3);
            c->slots = slots;
            c->flags_off = static_cast<size_t>(slots) * c->world * c->words_per_rank;
            c->window_words = c->flags_off + static_cast<size_t>(slots) * vmv::kMaxPeers;
            VMV_CUDA(cudaMalloc(reinterpret_cast<void **>(&c->window), c->window_words * sizeof(uint32_t)));
            VMV_CUDA(cudaMemset(c->window, 0, c->window_words * sizeof(uint32_t)));
            c->peer_window[c->rank] = c->window;
            if (c->world > 1)
            {
                // exchange the IPC handles of the windows through the communicator, then map every peer's window
                cudaIpcMemHandle_t mine;
                VMV_CUDA(cudaIpcGetMemHandle(&mine, c->window));
                unsigned char *d_handles = nullptr;
                VMV_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_handles), sizeof(mine) * c->world));
                VMV_CUDA(cudaMemcpy(d_handles + sizeof(mine) * c->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice));
                const int code = api->AllGather(d_handles + sizeof(mine) * c->rank, d_handles, sizeof(mine), kNcclUint8, c->nccl, nullptr);
                if (code != 0)
                {
                    cudaFree(d_handles);
                    return nccl_fail(*api, code, "ncclAllGather(ipc handles)");
                }
                std::vector<cudaIpcMemHandle_t> all(c->world);
                VMV_CUDA(cudaMemcpy(all.data(), d_handles, sizeof(mine) * c->world, cudaMemcpyDeviceToHost));
                cudaFree(d_handles);
                for (int p = 0; p < c->world; ++p)
                {
                    if (p != c->rank)
                    {
                        void *ptr = nullptr;
                        VMV_CUDA(cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess));
                        c->peer_window[p] = static_cast<uint32_t *>(ptr);
                    }
                }
                // nobody writes into a window before every rank has mapped (and zeroed) its own: one more collective
                uint32_t *d_tok = nullptr;
                VMV_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_tok), sizeof(uint32_t) * c->world));
                const int code2 = api->AllGather(d_tok + c->rank, d_tok, 1, kNcclUint32, c->nccl, nullptr);
                VMV_CUDA(cudaStreamSynchronize(nullptr));
                cudaFree(d_tok);
                if (code2 != 0)
                {
                    return nccl_fail(*api, code2, "ncclAllGather(window barrier)");
                }
            }
            return VMV_OK;
        });
    }

    uint32_t *vmv_comm_window_ptr(vmv_comm *c, int slot)
    {
        if (c == nullptr || c->window == nullptr || slot < 0 || slot >= c->slots)
        {
            fail(VMV_ERR_ARG, "vmv_comm_window_ptr: bad argument");
            return nullptr;
        }
        return c->window + static_cast<size_t>(slot) * c->world * c->words_per_rank;
    }

    uint32_t *vmv_comm_local_row(vmv_comm *c, int slot)
    {
        uint32_t *base = vmv_comm_window_ptr(c, slot);
        return base ? base + static_cast<size_t>(c->rank) * c->words_per_rank : nullptr;
    }

    size_t vmv_comm_window_stride(const vmv_comm *c)
    {
        return c ? c->words_per_rank : 0;
    }

    int vmv_validate_configs_gather_dev(int robot, const vmv_env *env, vmv_comm *c, int slot, const float *d_q, size_t n, void *stream)
    {
        return guarded("vmv_validate_configs_gather_dev", [&]() -> int
        {
            if (!valid_robot(robot) || (n > 0 && d_q == nullptr))
            {
                return fail(VMV_ERR_ARG, "vmv_validate_configs_gather_dev: bad argument");
            }
            int rc = check_gather_call(c, slot, n);
            if (rc != VMV_OK)
            {
                return rc;
            }
            const GatherDev g = gather_for(c, slot);
            if (n == 0)
            {
                return push_gather(g, g.peer_bits[c->rank], 0, static_cast<cudaStream_t>(stream));
            }
            return configs_common(robot, env, d_q, n, g.peer_bits[c->rank], g, stream);
        });
    }

    int vmv_validate_edges_indexed_gather_dev(int robot, const vmv_env *env, vmv_comm *c, int slot, const float *d_vertices, size_t n_vertices,
                                              const uint32_t *d_pairs, size_t n_edges, int resolution, void *stream)
    {
        return guarded("vmv_validate_edges_indexed_gather_dev", [&]() -> int
        {
            (void)n_vertices;
            if (!valid_robot(robot) || (n_edges > 0 && (d_vertices == nullptr || d_pairs == nullptr)))
            {
                return fail(VMV_ERR_ARG, "vmv_validate_edges_indexed_gather_dev: bad argument");
            }
            int rc = check_gather_call(c, slot, n_edges);
            if (rc != VMV_OK)
            {
                return rc;
            }
            const GatherDev g = gather_for(c, slot);
            if (n_edges == 0)
            {
                return push_gather(g, g.peer_bits[c->rank], 0, static_cast<cudaStream_t>(stream));
            }
            return edges_common(robot, env, d_vertices, nullptr, d_pairs, n_edges, resolution, g.peer_bits[c->rank], g, stream);
        });
    }

    // Publication of this rank's row of `slot` -- already written into the LOCAL window by any kernel enqueued on
    // `stream` -- to every peer by the copy engines: a side stream waits for `stream`, then per peer one copy of
    // the row followed by one copy of the flag word (in-stream order: the data lands before the flag).  No SM
    // is involved, and the copies overlap whatever `stream` runs next.
    int vmv_comm_publish(vmv_comm *c, int slot, size_t n_words, void *stream)
    {
        return guarded("vmv_comm_publish", [&]() -> int
        {
            int rc = check_gather_call(c, slot, n_words * 32);
            if (rc != VMV_OK)
            {
                return rc;
            }
            if (c->copy_stream == nullptr)
            {
                VMV_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
                for (cudaEvent_t &e : c->ready)
                {
                    VMV_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                }
                for (cudaEvent_t &e : c->copied)
                {
                    VMV_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                }
                std::vector<uint32_t> table(kSeqTable);
                std::iota(table.begin(), table.end(), 0u);
                VMV_CUDA(cudaMalloc(reinterpret_cast<void **>(&c->d_seq), kSeqTable * sizeof(uint32_t)));
                VMV_CUDA(cudaMemcpy(c->d_seq, table.data(), kSeqTable * sizeof(uint32_t), cudaMemcpyHostToDevice));
            }
            const uint32_t seq = ++c->seq[slot];
            if (seq >= kSeqTable)
            {
                return fail(VMV_ERR_LIMIT, "vmv_comm_publish: more than 4095 publications into one slot (recreate the window)");
            }
            VMV_CUDA(cudaEventRecord(c->ready[slot], static_cast<cudaStream_t>(stream)));
            VMV_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ready[slot], 0));
            const size_t row = static_cast<size_t>(slot) * c->world * c->words_per_rank + static_cast<size_t>(c->rank) * c->words_per_rank;
            const size_t flag = c->flags_off + static_cast<size_t>(slot) * vmv::kMaxPeers + c->rank;
            for (int k = 1; k <= c->world; ++k)
            {
                const int p = (c->rank + k) % c->world;  // the local flag last; peers in a rotated order (no hot receiver)
                if (p != c->rank && n_words > 0)
                {
                    VMV_CUDA(cudaMemcpyAsync(c->peer_window[p] + row, c->window + row, n_words * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->copy_stream));
                }
                VMV_CUDA(cudaMemcpyAsync(c->peer_window[p] + flag, c->d_seq + seq, sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->copy_stream));
            }
            VMV_CUDA(cudaEventRecord(c->copied[slot], c->copy_stream));
            c->published[slot] = true;
            return VMV_OK;
        });
    }

    // Before a launch on `stream` overwrites this rank's row of `slot`: wait (on the device, no host block) until the
    // copy engines have finished sending the row's previous contents.
    int vmv_comm_acquire(vmv_comm *c, int slot, void *stream)
    {
        return guarded("vmv_comm_acquire", [&]() -> int
        {
            if (c == nullptr || c->window == nullptr || slot < 0 || slot >= c->slots)
            {
                return fail(VMV_ERR_ARG, "vmv_comm_acquire: bad argument");
            }
            if (c->published[slot])
            {
                VMV_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), c->copied[slot], 0));
            }
            return VMV_OK;
        });
    }

    int vmv_comm_wait(vmv_comm *c, int slot, void *stream)
    {
        return guarded("vmv_comm_wait", [&]() -> int
        {
            if (c == nullptr || c->window == nullptr || slot < 0 || slot >= c->slots)
            {
                return fail(VMV_ERR_ARG, "vmv_comm_wait: bad argument");
            }
            if (c->seq[slot] == 0)
            {
                return VMV_OK;  // nothing was gathered into this slot yet
            }
            const uint32_t *flags = c->window + c->flags_off + static_cast<size_t>(slot) * vmv::kMaxPeers;
            if (BatchMemOpFn op = batch_mem_op())
            {
                WaitValueOp ops[vmv::kMaxPeers] = {};
                for (int p = 0; p < c->world; ++p)
                {
                    ops[p].operation = 1;
                    ops[p].address = reinterpret_cast<unsigned long long>(flags + p);
                    ops[p].value = c->seq[slot];
                    ops[p].flags = 0;
                }
                if (op(static_cast<cudaStream_t>(stream), static_cast<unsigned int>(c->world), ops, 0) == 0)
                {
                    return VMV_OK;
                }
                // (not supported on this driver / device: fall through to the kernel)
            }
            vmv::k_comm_wait<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(flags, c->world, c->seq[slot]);
            g_launches++;
            VMV_CUDA(cudaGetLastError());
            return VMV_OK;
        });
    }

    int vmv_env_broadcast(vmv_comm *c, vmv_env *env, int root)
    {
        return guarded("vmv_env_broadcast", [&]() -> int
        {
            if (c == nullptr || env == nullptr || root < 0 || root >= c->world)
            {
                return fail(VMV_ERR_ARG, "vmv_env_broadcast: bad argument");
            }
            const NcclApi *api = nullptr;
            int rc = nccl_api(api);
            if (rc != VMV_OK)
            {
                return rc;
            }
            Writer w;
            if (c->rank == root)
            {
                w.vec(env->spheres), w.vec(env->capsules), w.vec(env->z_capsules), w.vec(env->cuboids), w.vec(env->z_cuboids);
                w.pod<uint64_t>(env->heightfields.size());
                for (const auto &h : env->heightfields)
                {
                    w.pod(h.f), w.pod<uint64_t>(h.xd), w.pod<uint64_t>(h.yd), w.pod(h.id), w.vec(h.data);
                }
                w.pod<uint64_t>(env->capts.size());
                for (const auto &t : env->capts)
                {
                    const int rch = capt_ensure_host(t);
                    if (rch != VMV_OK)
                    {
                        return rch;
                    }
                    w.pod(t.r_min), w.pod(t.r_max), w.pod(t.r_point), w.pod(t.list_reach_sq), w.pod(t.nlog2), w.pod(t.n_points), w.pod(t.top_lo), w.pod(t.top_hi);
                    w.pod(t.g_origin), w.pod(t.g_inv0), w.pod(t.g_cell0), w.pod(t.id);
                    w.vec(t.nodes), w.vec(t.leafbits), w.vec(t.gpts), w.vec(t.gstart);
                }
                w.pod<uint64_t>(env->mvts.size());
                for (const auto &t : env->mvts)
                {
                    w.pod(t.r_min), w.pod(t.r_max), w.pod(t.r_point), w.pod(t.ws_min), w.pod(t.ws_max), w.pod(t.g_min), w.pod(t.g_max);
                    w.pod(t.inv_scale), w.pod(t.grid_width), w.pod(t.id);
                    w.vec(t.cells), w.vec(t.voxels), w.vec(t.points);
                }
                w.vec(env->cloud_xyz), w.vec(env->cloud_tag), w.pod(env->cloud_r_point_max);
                w.pod<int>(env->has_attachment ? 1 : 0), w.pod(env->attach_tf), w.vec(env->attach_spheres), w.pod(env->next_id);
            }
            unsigned long long size = w.b.size();
            unsigned long long *d_size = nullptr;
            VMV_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_size), sizeof(size)));
            VMV_CUDA(cudaMemcpy(d_size, &size, sizeof(size), cudaMemcpyHostToDevice));
            int code = api->Broadcast(d_size, d_size, sizeof(size), kNcclUint8, root, c->nccl, nullptr);
            if (code == 0)
            {
                VMV_CUDA(cudaMemcpy(&size, d_size, sizeof(size), cudaMemcpyDeviceToHost));
            }
            cudaFree(d_size);
            if (code != 0)
            {
                return nccl_fail(*api, code, "ncclBroadcast(size)");
            }
            unsigned char *d_buf = nullptr;
            VMV_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_buf), std::max<size_t>(size, 16)));
            if (c->rank == root)
            {
                VMV_CUDA(cudaMemcpy(d_buf, w.b.data(), size, cudaMemcpyHostToDevice));
            }
            code = api->Broadcast(d_buf, d_buf, size, kNcclUint8, root, c->nccl, nullptr);
            if (code != 0)
            {
                cudaFree(d_buf);
                return nccl_fail(*api, code, "ncclBroadcast(environment)");
            }
            if (c->rank != root)
            {
                std::vector<unsigned char> b(size);
                VMV_CUDA(cudaMemcpy(b.data(), d_buf, size, cudaMemcpyDeviceToHost));
                cudaFree(d_buf);
                Reader r{b.data(), b.data() + b.size()};
                env->release_device();
                env->shapes_dirty = true;
                r.vec(env->spheres), r.vec(env->capsules), r.vec(env->z_capsules), r.vec(env->cuboids), r.vec(env->z_cuboids);
                uint64_t k = 0;
                r.pod(k);
                env->heightfields.assign(r.ok ? k : 0, HHeight{});
                for (auto &h : env->heightfields)
                {
                    uint64_t xd = 0, yd = 0;
                    r.pod(h.f), r.pod(xd), r.pod(yd), r.pod(h.id), r.vec(h.data);
                    h.xd = xd, h.yd = yd;
                }
                r.pod(k);
                env->release_clouds();
                env->capts.clear();
                env->capts.resize(r.ok ? k : 0);
                for (auto &t : env->capts)
                {
                    r.pod(t.r_min), r.pod(t.r_max), r.pod(t.r_point), r.pod(t.list_reach_sq), r.pod(t.nlog2), r.pod(t.n_points), r.pod(t.top_lo), r.pod(t.top_hi);
                    r.pod(t.g_origin), r.pod(t.g_inv0), r.pod(t.g_cell0), r.pod(t.id);
                    r.vec(t.nodes), r.vec(t.leafbits), r.vec(t.gpts), r.vec(t.gstart);
                }
                r.pod(k);
                env->mvts.clear();
                env->mvts.resize(r.ok ? k : 0);
                for (auto &t : env->mvts)
                {
                    r.pod(t.r_min), r.pod(t.r_max), r.pod(t.r_point), r.pod(t.ws_min), r.pod(t.ws_max), r.pod(t.g_min), r.pod(t.g_max);
                    r.pod(t.inv_scale), r.pod(t.grid_width), r.pod(t.id);
                    r.vec(t.cells), r.vec(t.voxels), r.vec(t.points);
                }
                int att = 0;
                r.vec(env->cloud_xyz), r.vec(env->cloud_tag), r.pod(env->cloud_r_point_max);
                r.pod(att), r.pod(env->attach_tf), r.vec(env->attach_spheres), r.pod(env->next_id);
                env->has_attachment = att != 0;
                if (!r.ok)
                {
                    return fail(VMV_ERR_STATE, "vmv_env_broadcast: truncated environment stream");
                }
            }
            else
            {
                cudaFree(d_buf);
            }
            return vmv_env_commit(env);
        });
    }

}