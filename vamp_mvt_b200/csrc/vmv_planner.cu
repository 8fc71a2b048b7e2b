// Batched front-end of the reference's roadmap planner (planning/prm.hh) behind the C ABI.
//
// The reference grows a PRM one sample at a time: draw a sample (Halton), fkcc, k nearest roadmap vertices within
// the PRM* radius (nigh k-d tree, exact), one validate_motion per neighbour, union-find, A* once start and goal
// meet (prm.hh:43-196; build_roadmap :198-300 is the same loop without the termination test).  Nothing in that
// loop feeds back into the samples or into the neighbour queries -- a vertex's neighbour set depends only on the
// vertices inserted before it -- so the same roadmap comes out of three bulk steps:
//   1. samples 0 .. max_iterations-1 of the Halton sequence generated AND validated on the device
//      (vmv_validate_halton: one bit per sample back); the valid ones, in order, are the vertices;
//   2. every vertex's k nearest EARLIER vertices within r, k and r as the reference computes them for the roadmap
//      size at that moment (roadmap.hh:42-67) -- k_causal_knn below, a warp per vertex, exact, with the distance
//      arithmetic of the reference's Space (nn.hh:48-52: (a - b).l2_norm(), its hsum order);
//   3. all candidate edges as ONE indexed edge batch (vmv_validate_edges_indexed_dev, 8 bytes per edge).
// The sequential part that is left -- adjacency lists in the reference's order, union-find, the iteration at which
// start and goal first share a component, A* on the roadmap truncated there -- is replayed on the host from the
// bits.  Result: the reference's vertices, edges, path and iteration count (tests/test_planner.py pins them against
// the reference's own prm.hh compiled in place).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <new>
#include <vector>

#include "vmv_internal.h"
#include "vmv_fcit.hpp"

using namespace vmvh;

namespace vmv
{
    static constexpr int kKnnMax = 64;  // PRM* k = ceil((e + e/d) ln n): 43 at n = 10^6, d = 7

    // A warp per query vertex i: its (at most) k_of[i] nearest vertices among 0 .. i-1 within r_of[i], ascending by
    // (distance, index).  The running best list lives in shared memory, sorted; a candidate that beats the current
    // worst is inserted by the whole warp (position by ballot, shift, store).
    template <int DOF>
    __global__ void __launch_bounds__(128) k_causal_knn(
        const float *__restrict__ V,
        uint32_t first,
        uint32_t n_total,
        const uint32_t *__restrict__ k_of,
        const float *__restrict__ r_of,
        uint32_t *__restrict__ nbr_idx,   // [n_total][kKnnMax]
        float *__restrict__ nbr_dist,     // [n_total][kKnnMax]
        uint32_t *__restrict__ nbr_cnt)   // [n_total]
    {
        __shared__ float s_d[4][kKnnMax];
        __shared__ uint32_t s_i[4][kKnnMax];
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        const uint32_t i = first + blockIdx.x * 4 + w;
        if (i >= n_total)
        {
            return;
        }
        float *ld = s_d[w];
        uint32_t *li = s_i[w];
        const uint32_t k = min(k_of[i], static_cast<uint32_t>(kKnnMax));
        const float r = r_of[i];
        float q[DOF];
#pragma unroll
        for (int j = 0; j < DOF; ++j)
        {
            q[j] = __ldg(V + static_cast<size_t>(i) * DOF + j);
        }
        uint32_t cnt = 0;
        for (uint32_t base = 0; base < i && k > 0; base += 32)
        {
            const uint32_t j = base + lane;
            float d = 3.0e38F;
            if (j < i)
            {
                float diff[DOF];
#pragma unroll
                for (int c = 0; c < DOF; ++c)
                {
                    diff[c] = __fsub_rn(__ldg(V + static_cast<size_t>(j) * DOF + c), q[c]);
                }
                d = ref_l2_norm<DOF>(diff);
            }
            const float worst = cnt == k ? ld[k - 1] : 3.0e38F;
            uint32_t m = __ballot_sync(kFullWarp, j < i && d <= r && (cnt < k || d < worst));
            while (m != 0u)
            {
                const int src = __ffs(static_cast<int>(m)) - 1;
                m &= m - 1u;
                const float cd = __shfl_sync(kFullWarp, d, src);
                const uint32_t cj = base + src;
                if (cnt == k && !(cd < ld[k - 1]))
                {
                    continue;
                }
                // entries lane and lane + 32 of the sorted list
                const float d0 = ld[lane], d1 = ld[lane + 32];
                const uint32_t i0 = li[lane], i1 = li[lane + 32];
                const bool in0 = static_cast<uint32_t>(lane) < cnt, in1 = static_cast<uint32_t>(lane + 32) < cnt;
                const bool less0 = in0 && (d0 < cd || (d0 == cd && i0 < cj));
                const bool less1 = in1 && (d1 < cd || (d1 == cd && i1 < cj));
                const uint32_t pos = __popc(__ballot_sync(kFullWarp, less0)) + __popc(__ballot_sync(kFullWarp, less1));
                __syncwarp();
                if (in0 && static_cast<uint32_t>(lane) >= pos && static_cast<uint32_t>(lane) + 1 < k)
                {
                    ld[lane + 1] = d0, li[lane + 1] = i0;
                }
                if (in1 && static_cast<uint32_t>(lane + 32) >= pos && static_cast<uint32_t>(lane + 32) + 1 < k)
                {
                    ld[lane + 33] = d1, li[lane + 33] = i1;
                }
                if (lane == 0)
                {
                    ld[pos] = cd, li[pos] = cj;
                }
                cnt = min(cnt + 1, k);
                __syncwarp();
            }
        }
        __syncwarp();
        if (lane == 0)
        {
            nbr_cnt[i] = cnt;
        }
        for (uint32_t e = lane; e < cnt; e += 32)
        {
            nbr_idx[static_cast<size_t>(i) * kKnnMax + e] = li[e];
            nbr_dist[static_cast<size_t>(i) * kKnnMax + e] = ld[e];
        }
    }

    // candidate edges of vertices first .. n-1 as index pairs (neighbour, vertex): validate_motion(neighbor, temp), prm.hh:138
    __global__ void k_knn_pairs(const uint32_t *__restrict__ nbr_idx, const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ nbr_cnt,
                                uint32_t first, uint32_t n_total, uint32_t *__restrict__ pairs)
    {
        const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n_total)
        {
            return;
        }
        const uint32_t o = offsets[i], c = nbr_cnt[i];
        for (uint32_t e = 0; e < c; ++e)
        {
            pairs[2 * (o + e)] = nbr_idx[static_cast<size_t>(i) * kKnnMax + e];
            pairs[2 * (o + e) + 1] = i;
        }
    }
}  // namespace vmv

struct vmv_roadmap
{
    int dof = 0;
    std::vector<float> vertices;         // [n][dof]
    std::vector<uint32_t> adj_pairs;     // adjacency entries (from, to) in the reference's enumeration (Roadmap::edges)
    std::vector<float> adj_cost;
    std::vector<float> path;             // solve: waypoints start .. goal (utils::recover_path)
    float cost = std::numeric_limits<float>::infinity();
    size_t iterations = 0;
    size_t samples_drawn = 0, edges_checked = 0;
};

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
        template <typename T>
        T *as()
        {
            return static_cast<T *>(p);
        }
    };

    // PRMStarNeighborParams (roadmap.hh:42-67), in the reference's double arithmetic
    constexpr double kE = 2.718281828459045235360287471352662498;  // vamp::utils::constants::e
    double unit_ball_measure(int dim)
    {
        // utils.hh: pi^(d/2) / Gamma(d/2 + 1)
        return std::pow(std::sqrt(3.141592653589793238462643383279502884), static_cast<double>(dim)) / std::tgamma(static_cast<double>(dim) / 2.0 + 1.0);
    }
    uint32_t prmstar_k(int dim, size_t n)
    {
        const double c = kE + (kE / static_cast<double>(dim));
        return static_cast<uint32_t>(std::ceil(c * std::log(static_cast<double>(n))));
    }
    float prmstar_r(int dim, double space_measure, size_t n)
    {
        const double inv = 1.0 / static_cast<double>(dim);
        const double ratio = space_measure / unit_ball_measure(dim);
        const double c = 2.0 * std::pow(1.0 + inv, inv) * std::pow(ratio, inv);
        return static_cast<float>(2.0 * c * std::pow(std::log(static_cast<double>(n)) / static_cast<double>(n), inv));
    }

    // Configuration::distance on the host: (a - b).l2_norm() with the AVX hsum order (vector/avx.hh:441-452)
    float ref_distance(const float *a, const float *b, int dof)
    {
        float lane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < 8; ++k)
        {
            if (k + 8 < dof)
            {
                const float d0 = a[k] - b[k], d1 = a[k + 8] - b[k + 8];
                lane[k] = std::fma(d0, d0, d1 * d1);
            }
            else if (k < dof)
            {
                const float d0 = a[k] - b[k];
                lane[k] = d0 * d0;
            }
        }
        const float s0 = lane[4] + lane[0], s1 = lane[5] + lane[1], s2 = lane[6] + lane[2], s3 = lane[7] + lane[3];
        return std::sqrt((s0 + s2) + (s1 + s3));
    }

    struct Graph
    {
        std::vector<std::vector<std::pair<uint32_t, float>>> nbrs;
    };

    template <int DOF>
    int launch_knn(const float *dV, uint32_t first, uint32_t n, const uint32_t *dk, const float *dr, uint32_t *di, float *dd, uint32_t *dc)
    {
        if (n > first)
        {
            vmv::k_causal_knn<DOF><<<(n - first + 3) / 4, 128>>>(dV, first, n, dk, dr, di, dd, dc);
            g_launches++;
            VMV_CUDA(cudaGetLastError());
        }
        return VMV_OK;
    }

    // which: 0 build_roadmap, 1 solve
    int prm_run(int robot, const vmv_env *env, const float *start, const float *goal, size_t max_iterations, size_t max_samples, double space_measure,
                int which, vmv_roadmap &R)
    {
        const RobotHost &rh = *ops(robot).host;
        const int dof = rh.dof;
        R.dof = dof;
        const int res = rh.resolution;

        if (which == 1)
        {
            // the straight line first (prm.hh:57-70)
            uint32_t w = 0;
            int rc = vmv_validate_edges(robot, env, start, goal, 1, res, &w);
            if (rc != VMV_OK)
            {
                return rc;
            }
            if (w & 1u)
            {
                R.path.assign(start, start + dof);
                R.path.insert(R.path.end(), goal, goal + dof);
                R.iterations = 0;
                R.vertices.assign(start, start + dof);
                R.vertices.insert(R.vertices.end(), goal, goal + dof);
                return VMV_OK;
            }
        }
        if (max_samples < 2)
        {
            return fail(VMV_ERR_ARG, "vmv_prm: max_samples must be at least 2 (start and goal)");
        }
        const uint64_t limit = vmv_halton_exact_limit(robot);
        if (max_iterations > limit)
        {
            return fail(VMV_ERR_LIMIT, "vmv_prm: max_iterations beyond the exact range of the device Halton sampler");
        }

        // ---- 1. samples: drawn and validated on the device, a chunk at a time, until the loop condition of the
        //         reference (iter++ < max_iterations and nodes.size() < max_samples) would stop --------------------
        std::vector<float> &V = R.vertices;
        V.assign(start, start + dof);
        V.insert(V.end(), goal, goal + dof);
        std::vector<uint32_t> sample_of;  // per vertex >= 2: index of its sample in the stream
        size_t drawn = 0;
        {
            const size_t chunk = 1 << 16;
            std::vector<uint32_t> bits((chunk + 31) / 32);
            std::vector<float> q(chunk * dof);
            bool full = false;
            while (!full && drawn < max_iterations)
            {
                const size_t n = std::min(chunk, max_iterations - drawn);
                int rc = vmv_validate_halton(robot, env, drawn, n, bits.data(), q.data());
                if (rc != VMV_OK)
                {
                    return rc;
                }
                size_t t = 0;
                for (; t < n; ++t)
                {
                    if (V.size() / dof >= max_samples)
                    {
                        full = true;
                        break;
                    }
                    if ((bits[t >> 5] >> (t & 31)) & 1u)
                    {
                        V.insert(V.end(), q.begin() + t * dof, q.begin() + (t + 1) * dof);
                        sample_of.push_back(static_cast<uint32_t>(drawn + t));
                    }
                }
                drawn += t;
            }
            if (V.size() / dof >= max_samples)
            {
                full = true;
            }
        }
        R.samples_drawn = drawn;
        // iter after the loop: one more than the samples consumed (the failing test still post-increments)
        R.iterations = drawn + 1;
        const uint32_t nv = static_cast<uint32_t>(V.size() / dof);

        // ---- 2. neighbours: k and r of the roadmap size at each insertion (= the vertex's own index) --------------
        std::vector<uint32_t> k_of(nv, 0);
        std::vector<float> r_of(nv, 0.F);
        for (uint32_t i = 2; i < nv; ++i)
        {
            k_of[i] = prmstar_k(dof, i);
            r_of[i] = prmstar_r(dof, space_measure, i);
            if (k_of[i] > static_cast<uint32_t>(vmv::kKnnMax))
            {
                return fail(VMV_ERR_LIMIT, "vmv_prm: more than 64 neighbours per vertex");
            }
        }
        DevBuf dV, dk, dr, di, dd, dc, doff, dpairs, dbits;
        int rc = dV.alloc(V.size() * 4);
        rc = rc == VMV_OK ? dk.alloc(nv * 4) : rc;
        rc = rc == VMV_OK ? dr.alloc(nv * 4) : rc;
        rc = rc == VMV_OK ? di.alloc(static_cast<size_t>(nv) * vmv::kKnnMax * 4) : rc;
        rc = rc == VMV_OK ? dd.alloc(static_cast<size_t>(nv) * vmv::kKnnMax * 4) : rc;
        rc = rc == VMV_OK ? dc.alloc(nv * 4) : rc;
        rc = rc == VMV_OK ? doff.alloc((nv + 1) * 4) : rc;
        if (rc != VMV_OK)
        {
            return rc;
        }
        VMV_CUDA(cudaMemcpy(dV.p, V.data(), V.size() * 4, cudaMemcpyHostToDevice));
        VMV_CUDA(cudaMemcpy(dk.p, k_of.data(), nv * 4, cudaMemcpyHostToDevice));
        VMV_CUDA(cudaMemcpy(dr.p, r_of.data(), nv * 4, cudaMemcpyHostToDevice));
        VMV_CUDA(cudaMemset(dc.p, 0, nv * 4));
        switch (dof)
        {
            case 6:
                rc = launch_knn<6>(dV.as<float>(), 2, nv, dk.as<uint32_t>(), dr.as<float>(), di.as<uint32_t>(), dd.as<float>(), dc.as<uint32_t>());
                break;
            case 7:
                rc = launch_knn<7>(dV.as<float>(), 2, nv, dk.as<uint32_t>(), dr.as<float>(), di.as<uint32_t>(), dd.as<float>(), dc.as<uint32_t>());
                break;
            case 8:
                rc = launch_knn<8>(dV.as<float>(), 2, nv, dk.as<uint32_t>(), dr.as<float>(), di.as<uint32_t>(), dd.as<float>(), dc.as<uint32_t>());
                break;
            case 14:
                rc = launch_knn<14>(dV.as<float>(), 2, nv, dk.as<uint32_t>(), dr.as<float>(), di.as<uint32_t>(), dd.as<float>(), dc.as<uint32_t>());
                break;
            default:
                rc = fail(VMV_ERR_ARG, "vmv_prm: unsupported joint count");
        }
        if (rc != VMV_OK)
        {
            return rc;
        }
        std::vector<uint32_t> cnt(nv), off(nv + 1, 0);
        VMV_CUDA(cudaMemcpy(cnt.data(), dc.p, nv * 4, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < nv; ++i)
        {
            off[i + 1] = off[i] + cnt[i];
        }
        const size_t n_cand = off[nv];
        R.edges_checked = n_cand;

        // ---- 3. all candidate edges in one indexed batch ---------------------------------------------------------
        std::vector<uint32_t> words((n_cand + 31) / 32, 0u);
        std::vector<uint32_t> nidx(static_cast<size_t>(nv) * vmv::kKnnMax);
        std::vector<float> ndist(static_cast<size_t>(nv) * vmv::kKnnMax);
        if (n_cand > 0)
        {
            rc = dpairs.alloc(n_cand * 8);
            rc = rc == VMV_OK ? dbits.alloc(words.size() * 4) : rc;
            if (rc != VMV_OK)
            {
                return rc;
            }
            VMV_CUDA(cudaMemcpy(doff.p, off.data(), (nv + 1) * 4, cudaMemcpyHostToDevice));
            vmv::k_knn_pairs<<<(nv - 2 + 127) / 128, 128>>>(di.as<uint32_t>(), doff.as<uint32_t>(), dc.as<uint32_t>(), 2, nv, dpairs.as<uint32_t>());
            g_launches++;
            VMV_CUDA(cudaGetLastError());
            rc = vmv_validate_edges_indexed_dev(robot, env, dV.as<float>(), nv, dpairs.as<uint32_t>(), n_cand, res, dbits.as<uint32_t>(), nullptr);
            if (rc != VMV_OK)
            {
                return rc;
            }
            VMV_CUDA(cudaMemcpy(words.data(), dbits.p, words.size() * 4, cudaMemcpyDeviceToHost));
            VMV_CUDA(cudaMemcpy(nidx.data(), di.p, nidx.size() * 4, cudaMemcpyDeviceToHost));
            VMV_CUDA(cudaMemcpy(ndist.data(), dd.p, ndist.size() * 4, cudaMemcpyDeviceToHost));
        }

        // ---- replay of the sequential part on the bits -------------------------------------------------------------
        Graph G;
        G.nbrs.resize(nv);
        std::vector<uint32_t> parent(nv), size(nv, 1), comp(nv);
        for (uint32_t i = 0; i < nv; ++i)
        {
            parent[i] = i;
            comp[i] = i;
        }
        // components are indexed as the reference numbers them (prm.hh:161-175): start 0, goal 1, then one per vertex
        // that arrives without a neighbour
        std::vector<uint32_t> c_parent = {0, 1}, c_size = {1, 1};
        auto find_root = [&](uint32_t c)
        {
            while (c != c_parent[c])
            {
                const uint32_t old = c_parent[c];
                c_parent[c] = c_parent[c_parent[c]];
                c = old;
            }
            return c;
        };
        auto merge = [&](uint32_t a, uint32_t b)
        {
            a = find_root(a), b = find_root(b);
            if (a == b)
            {
                return;
            }
            if (c_size[a] < c_size[b])
            {
                c_parent[a] = b;
                c_size[b] += c_size[a];
            }
            else
            {
                c_parent[b] = a;
                c_size[a] += c_size[b];
            }
        };
        comp[0] = 0, comp[1] = 1;
        uint32_t n_used = nv;
        bool solved = false;
        for (uint32_t i = 2; i < nv; ++i)
        {
            for (uint32_t e = 0; e < cnt[i]; ++e)
            {
                const size_t bit = off[i] + e;
                if ((words[bit >> 5] >> (bit & 31)) & 1u)
                {
                    const uint32_t j = nidx[static_cast<size_t>(i) * vmv::kKnnMax + e];
                    const float d = ndist[static_cast<size_t>(i) * vmv::kKnnMax + e];
                    G.nbrs[i].emplace_back(j, d);
                    G.nbrs[j].emplace_back(i, d);
                }
            }
            if (which == 1)
            {
                if (G.nbrs[i].empty())
                {
                    comp[i] = static_cast<uint32_t>(c_parent.size());
                    c_parent.push_back(comp[i]);
                    c_size.push_back(1);
                }
                else
                {
                    comp[i] = comp[G.nbrs[i].front().first];
                    for (const auto &nb : G.nbrs[i])
                    {
                        merge(comp[i], comp[nb.first]);
                    }
                }
                if (find_root(0) == find_root(1))
                {
                    n_used = i + 1;
                    solved = true;
                    R.iterations = static_cast<size_t>(sample_of[i - 2]) + 1;  // iter at the sample that closed the gap
                    break;
                }
            }
        }

        if (which == 1)
        {
            // the roadmap as it stood at that iteration
            V.resize(static_cast<size_t>(n_used) * dof);
            G.nbrs.resize(n_used);
            for (auto &l : G.nbrs)
            {
                l.erase(std::remove_if(l.begin(), l.end(), [&](const auto &nb) { return nb.first >= n_used; }), l.end());
            }
            if (solved)
            {
                // utils::astar (planning/utils.hh:73-137) on the roadmap
                const float inf = std::numeric_limits<float>::infinity();
                std::vector<float> g(n_used, inf);
                std::vector<uint32_t> par(n_used, std::numeric_limits<uint32_t>::max());
                std::vector<bool> expanded(n_used, false);
                g[0] = 0.F;
                par[0] = 0;
                struct QN
                {
                    uint32_t index;
                    float cost;
                };
                std::vector<QN> open;
                open.push_back({0, ref_distance(goal, start, dof)});
                while (!open.empty())
                {
                    const QN cur = open.back();
                    open.pop_back();
                    if (cur.index == 1)
                    {
                        break;
                    }
                    if (expanded[cur.index])
                    {
                        continue;
                    }
                    expanded[cur.index] = true;
                    const float cg = g[cur.index];
                    for (const auto &[ni, nd] : G.nbrs[cur.index])
                    {
                        if (cg + nd >= g[ni])
                        {
                            continue;
                        }
                        g[ni] = cg + nd;
                        par[ni] = cur.index;
                        open.push_back({ni, g[ni] + ref_distance(goal, V.data() + static_cast<size_t>(ni) * dof, dof)});
                    }
                    std::sort(open.begin(), open.end(), [](const QN &a, const QN &b) { return a.cost > b.cost; });
                }
                // recover_path (utils.hh:144-162): goal back to start, then reversed
                std::vector<uint32_t> chain;
                uint32_t c = 1;
                while (par[c] != c && par[c] != std::numeric_limits<uint32_t>::max())
                {
                    chain.push_back(c);
                    c = par[c];
                }
                chain.push_back(0);
                for (auto it = chain.rbegin(); it != chain.rend(); ++it)
                {
                    R.path.insert(R.path.end(), V.begin() + static_cast<size_t>(*it) * dof, V.begin() + static_cast<size_t>(*it + 1) * dof);
                }
                R.cost = g[1];
            }
        }
        // Roadmap::edges enumeration (prm.hh:283-292): per vertex, its adjacency list in insertion order
        for (uint32_t i = 0; i < G.nbrs.size(); ++i)
        {
            for (const auto &[j, d] : G.nbrs[i])
            {
                R.adj_pairs.push_back(i);
                R.adj_pairs.push_back(j);
                R.adj_cost.push_back(d);
            }
        }
        return VMV_OK;
    }

    // FCIT* (fcit.hh) with the GPU behind both providers of vmvfcit::solve
    int fcit_run(int robot, const vmv_env *env, const float *start, const float *goal, size_t max_iterations, size_t max_samples, size_t batch_size,
                 bool optimize, vmv_roadmap &R)
    {
        const RobotHost &rh = *ops(robot).host;
        const int dof = rh.dof;
        R.dof = dof;
        std::vector<float> &states = R.vertices;
        states.assign(start, start + dof);
        states.insert(states.end(), goal, goal + dof);
        int err = VMV_OK;

        // samples: the Halton stream generated and validated on the device a chunk at a time; what a batch does not
        // use stays in the look-ahead buffer for the next one
        const uint64_t limit = vmv_halton_exact_limit(robot);
        uint64_t pos = 0;  // next sample of the stream not yet fetched
        std::vector<uint32_t> bits;
        std::vector<float> q;
        size_t have = 0, at = 0;  // samples in the buffer / next one to look at
        auto samples = [&](size_t n, std::vector<float> &st) -> size_t
        {
            size_t got = 0;
            while (got < n && err == VMV_OK)
            {
                if (at == have)
                {
                    const size_t chunk = std::min<uint64_t>(std::max<size_t>(4096, 2 * (n - got)), limit > pos ? limit - pos : 0);
                    if (chunk == 0)
                    {
                        break;  // beyond the exact range of the device sampler
                    }
                    bits.assign((chunk + 31) / 32, 0u);
                    q.resize(chunk * dof);
                    err = vmv_validate_halton(robot, env, pos, chunk, bits.data(), q.data());
                    if (err != VMV_OK)
                    {
                        break;
                    }
                    pos += chunk;
                    have = chunk, at = 0;
                }
                for (; at < have && got < n; ++at)
                {
                    if ((bits[at >> 5] >> (at & 31)) & 1u)
                    {
                        st.insert(st.end(), q.begin() + at * dof, q.begin() + (at + 1) * dof);
                        ++got;
                    }
                }
            }
            R.samples_drawn = static_cast<size_t>(pos - (have - at));
            return got;
        };

        // edges: the first question about a parent validates its edges to EVERY node known so far in one indexed batch
        std::vector<std::vector<uint32_t>> rows;
        std::vector<uint32_t> rows_len;
        std::vector<uint32_t> pairs, words;
        auto edges = [&](uint32_t p, uint32_t c) -> bool
        {
            const uint32_t n_nodes = static_cast<uint32_t>(states.size() / dof);
            if (rows.size() < n_nodes)
            {
                rows.resize(n_nodes);
                rows_len.resize(n_nodes, 0);
            }
            if (c >= rows_len[p] && err == VMV_OK)
            {
                const uint32_t lo = rows_len[p];
                pairs.clear();
                for (uint32_t j = lo; j < n_nodes; ++j)
                {
                    pairs.push_back(p);
                    pairs.push_back(j);
                }
                words.assign((n_nodes - lo + 31) / 32, 0u);
                err = vmv_validate_edges_indexed(robot, env, states.data(), n_nodes, pairs.data(), n_nodes - lo, rh.resolution, words.data());
                R.edges_checked += n_nodes - lo;
                rows[p].resize((n_nodes + 31) / 32, 0u);
                for (uint32_t j = lo; j < n_nodes; ++j)
                {
                    if ((words[(j - lo) >> 5] >> ((j - lo) & 31)) & 1u)
                    {
                        rows[p][j >> 5] |= 1u << (j & 31);
                    }
                }
                rows_len[p] = n_nodes;
            }
            return err == VMV_OK && ((rows[p][c >> 5] >> (c & 31)) & 1u) != 0u;
        };
        auto dist = [dof](const float *a, const float *b) { return ref_distance(a, b, dof); };
        const vmvfcit::Result res = vmvfcit::solve(states, dof, max_iterations, max_samples, batch_size, optimize, dist, samples, edges);
        if (err != VMV_OK)
        {
            return err;
        }
        for (const uint32_t i : res.path)
        {
            R.path.insert(R.path.end(), states.begin() + static_cast<size_t>(i) * dof, states.begin() + static_cast<size_t>(i + 1) * dof);
        }
        R.cost = res.cost;
        R.iterations = res.iterations;
        return VMV_OK;
    }

    template <typename F>
    int guarded_planner(const char *what, F &&f) noexcept
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
    int vmv_prm(int robot, const vmv_env *env, const float *start, const float *goal, size_t max_iterations, size_t max_samples, double space_measure,
                int solve, vmv_roadmap **out)
    {
        return guarded_planner("vmv_prm", [&]() -> int
        {
            if (robot < 0 || robot >= VMV_N_ROBOTS || env == nullptr || start == nullptr || goal == nullptr || out == nullptr || !(space_measure > 0.0))
            {
                return fail(VMV_ERR_ARG, "vmv_prm: bad argument");
            }
            vmv_roadmap *R = new vmv_roadmap();
            const int rc = prm_run(robot, env, start, goal, max_iterations, max_samples, space_measure, solve ? 1 : 0, *R);
            if (rc != VMV_OK)
            {
                delete R;
                return rc;
            }
            *out = R;
            return VMV_OK;
        });
    }

    int vmv_fcit(int robot, const vmv_env *env, const float *start, const float *goal, size_t max_iterations, size_t max_samples, size_t batch_size,
                 int optimize, vmv_roadmap **out)
    {
        return guarded_planner("vmv_fcit", [&]() -> int
        {
            if (robot < 0 || robot >= VMV_N_ROBOTS || env == nullptr || start == nullptr || goal == nullptr || out == nullptr || max_samples < 2)
            {
                return fail(VMV_ERR_ARG, "vmv_fcit: bad argument");
            }
            vmv_roadmap *R = new vmv_roadmap();
            const int rc = fcit_run(robot, env, start, goal, max_iterations, max_samples, batch_size, optimize != 0, *R);
            if (rc != VMV_OK)
            {
                delete R;
                return rc;
            }
            *out = R;
            return VMV_OK;
        });
    }

    void vmv_roadmap_destroy(vmv_roadmap *r)
    {
        delete r;
    }

    size_t vmv_roadmap_vertices(const vmv_roadmap *r, const float **q)
    {
        if (r == nullptr)
        {
            return 0;
        }
        if (q != nullptr)
        {
            *q = r->vertices.data();
        }
        return r->dof > 0 ? r->vertices.size() / r->dof : 0;
    }

    size_t vmv_roadmap_edges(const vmv_roadmap *r, const uint32_t **pairs, const float **cost)
    {
        if (r == nullptr)
        {
            return 0;
        }
        if (pairs != nullptr)
        {
            *pairs = r->adj_pairs.data();
        }
        if (cost != nullptr)
        {
            *cost = r->adj_cost.data();
        }
        return r->adj_cost.size();
    }

    size_t vmv_roadmap_path(const vmv_roadmap *r, const float **q, float *cost)
    {
        if (r == nullptr)
        {
            return 0;
        }
        if (q != nullptr)
        {
            *q = r->path.data();
        }
        if (cost != nullptr)
        {
            *cost = r->cost;
        }
        return r->dof > 0 ? r->path.size() / r->dof : 0;
    }

    size_t vmv_roadmap_iterations(const vmv_roadmap *r)
    {
        return r ? r->iterations : 0;
    }

    size_t vmv_roadmap_work(const vmv_roadmap *r, size_t *edges_checked)
    {
        if (r != nullptr && edges_checked != nullptr)
        {
            *edges_checked = r->edges_checked;
        }
        return r ? r->samples_drawn : 0;
    }
}
