// CAPT build on the device (reference collision/capt.hh:106-119: median splits on x, y, z in turn over 2^nlog2 points, the
// input padded with +inf).  The tree is all that is built (vmv_device.cuh: capt_member evaluates the reference's affordance
// lists per point at query time), so the build is 17 rounds of "sort every node's range along one axis, read the split
// off the middle": one segmented bitonic sort per level over (coordinate, point index) pairs -- segments are aligned
// powers of two, so the network of a plain bitonic sort serves, with the last merge of every segment ascending.
// Ties are ordered by the point index: which half a point lands in when it ties with the median is the choice of the
// reference's sort (pdqsort_branchless, unstable) and unspecified there; the host build (std::sort) makes a third choice.
// Then: per node {split, inherited bit}, per leaf its two flags (walking its root path for the cell) and the leaf of every
// input point; the Morton grid of the raw points (one more sort, by cell code) and its per-level start tables.
#pragma once
#include <cstdint>

#include <cuda_runtime.h>

#include "vmv_device.cuh"

namespace vmv
{
    static constexpr uint32_t kSortTile = 2048;  // elements per block of the shared-memory stages (1024 threads)

    template <typename K>
    __device__ __forceinline__ bool sort_less(K ka, uint32_t ia, K kb, uint32_t ib)
    {
        return ka < kb || (ka == kb && ia < ib);
    }

    template <typename K>
    __device__ __forceinline__ void sort_cex(K &ka, uint32_t &ia, K &kb, uint32_t &ib, bool ascending)
    {
        // a sits at the lower position
        if (sort_less(kb, ib, ka, ia) == ascending)
        {
            const K tk = ka;
            ka = kb, kb = tk;
            const uint32_t ti = ia;
            ia = ib, ib = ti;
        }
    }

    // one compare-exchange stage (merge size k, distance j >= kSortTile) over the whole array; seg = segment size
    template <typename K>
    static __global__ void __launch_bounds__(256) k_sort_stage(K *__restrict__ key, uint32_t *__restrict__ idx, uint32_t n, uint32_t k, uint32_t j, uint32_t seg)
    {
        const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
        if (t >= n / 2)
        {
            return;
        }
        const uint32_t i = 2u * j * (t / j) + (t % j), p = i + j;
        const bool asc = (k == seg) || ((i & k) == 0u);
        K ka = key[i], kb = key[p];
        uint32_t ia = idx[i], ib = idx[p];
        sort_cex(ka, ia, kb, ib, asc);
        key[i] = ka, key[p] = kb;
        idx[i] = ia, idx[p] = ib;
    }

    // all stages with distance < kSortTile of the merge sizes k_first .. k_last (doubling), one tile per block in shared memory
    template <typename K>
    static __global__ void __launch_bounds__(1024) k_sort_tile(K *__restrict__ key, uint32_t *__restrict__ idx, uint32_t n, uint32_t k_first, uint32_t k_last, uint32_t seg)
    {
        __shared__ K sk[kSortTile];
        __shared__ uint32_t si[kSortTile];
        const uint32_t base = blockIdx.x * kSortTile, t = threadIdx.x;
        for (uint32_t u = t; u < kSortTile; u += blockDim.x)
        {
            sk[u] = key[base + u];
            si[u] = idx[base + u];
        }
        __syncthreads();
        for (uint32_t k = k_first; k <= k_last; k <<= 1)
        {
            for (uint32_t j = (k / 2 < kSortTile / 2 ? k / 2 : kSortTile / 2); j > 0u; j >>= 1)
            {
                const uint32_t i = 2u * j * (t / j) + (t % j), p = i + j;
                const bool asc = (k == seg) || (((base + i) & k) == 0u);
                sort_cex(sk[i], si[i], sk[p], si[p], asc);
                __syncthreads();
            }
        }
        for (uint32_t u = t; u < kSortTile; u += blockDim.x)
        {
            key[base + u] = sk[u];
            idx[base + u] = si[u];
        }
    }

    // every aligned segment of `seg` elements ascending by (key, idx); n a power of two >= kSortTile, seg a power of two
    template <typename K>
    inline cudaError_t segmented_sort(K *key, uint32_t *idx, uint32_t n, uint32_t seg, cudaStream_t s, uint64_t &launches)
    {
        if (seg < 2u)
        {
            return cudaSuccess;
        }
        const uint32_t local = seg < kSortTile ? seg : kSortTile;
        k_sort_tile<K><<<n / kSortTile, 1024, 0, s>>>(key, idx, n, 2u, local, seg);
        ++launches;
        for (uint32_t k = 2u * kSortTile; k <= seg && k != 0u; k <<= 1)
        {
            for (uint32_t j = k / 2; j >= kSortTile; j >>= 1)
            {
                k_sort_stage<K><<<(n / 2 + 255) / 256, 256, 0, s>>>(key, idx, n, k, j, seg);
                ++launches;
            }
            k_sort_tile<K><<<n / kSortTile, 1024, 0, s>>>(key, idx, n, k, k, seg);
            ++launches;
        }
        return cudaGetLastError();
    }

    static __global__ void k_fill_u32(uint32_t *__restrict__ dst, uint32_t pattern, size_t count)
    {
        const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
        if (i < count)
        {
            dst[i] = pattern;
        }
    }

    static __global__ void k_capt_gather_axis(const float *__restrict__ pts, const uint32_t *__restrict__ idx, uint32_t n, int d, float *__restrict__ key)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n)
        {
            key[i] = pts[3u * idx[i] + d];
        }
    }

    // the nodes of one level: split = (the two middle elements' sum, in float) / 2; the high half inherits the low half's
    // points iff the low half's smallest element is within r_max of the plane (capt.hh:232-246: all or nothing)
    static __global__ void k_capt_level_nodes(const float *__restrict__ key, uint32_t seg, uint32_t count, uint32_t first_node, float r_max, float2 *__restrict__ nodes)
    {
        const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
        if (m >= count)
        {
            return;
        }
        const uint32_t begin = m * seg, half = seg / 2;
        const float test = __fadd_rn(key[begin + half - 1], key[begin + half]) * 0.5F;
        const float first = key[begin];
        const bool inherited = (first >= __fsub_rn(test, r_max)) && (fabsf(first) <= 3.4028234664e38F);
        nodes[first_node + m] = make_float2(test, __uint_as_float(inherited ? 1u : 0u));
    }

    // one thread per word of leaf flags (16 leaves): the leaf's cell from its root path, the reference's test for a cell
    // inside the smallest query ball around its representative (capt.hh:39-46,150), the leaf of every input point
    static __global__ void k_capt_leaves(const float *__restrict__ pts, const uint32_t *__restrict__ order, const float2 *__restrict__ nodes, uint32_t nlog2,
                                         uint32_t n_leaves, float min_l2, uint32_t *__restrict__ leafbits, uint32_t *__restrict__ leaf_of_point)
    {
        const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
        if (w >= (n_leaves + 15u) / 16u)
        {
            return;
        }
        const float inf = __int_as_float(0x7f800000);
        uint32_t word = 0u;
        for (uint32_t u = 0; u < 16u; ++u)
        {
            const uint32_t leaf = 16u * w + u;
            if (leaf >= n_leaves)
            {
                break;
            }
            const uint32_t id = order[leaf];
            const float rep[3] = {pts[3u * id], pts[3u * id + 1u], pts[3u * id + 2u]};
            leaf_of_point[id] = leaf;
            if (!(fabsf(rep[0]) <= 3.4028234664e38F))
            {
                continue;
            }
            float lo[3] = {-inf, -inf, -inf}, hi[3] = {inf, inf, inf};
            uint32_t node = 0u;
            int d = 0;
            for (uint32_t l = 0; l < nlog2; ++l)
            {
                const float test = nodes[node].x;
                const uint32_t up = (leaf >> (nlog2 - 1u - l)) & 1u;
                if (d == 0)
                {
                    (up ? lo[0] : hi[0]) = test;
                }
                else if (d == 1)
                {
                    (up ? lo[1] : hi[1]) = test;
                }
                else
                {
                    (up ? lo[2] : hi[2]) = test;
                }
                node = 2u * node + 1u + up;
                d = d == 2 ? 0 : d + 1;
            }
            const float d0 = fmaxf(__fsub_rn(rep[0], lo[0]), __fsub_rn(hi[0], rep[0]));
            const float d1 = fmaxf(__fsub_rn(rep[1], lo[1]), __fsub_rn(hi[1], rep[1]));
            const float d2 = fmaxf(__fsub_rn(rep[2], lo[2]), __fsub_rn(hi[2], rep[2]));
            const bool rep_only = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)) <= min_l2;
            word |= (2u | (rep_only ? 0u : 1u)) << (2u * u);
        }
        leafbits[w] = word;
    }

    // Morton code of the finest grid cell of every point (vmv_device.cuh: the queries derive cell ranges from the same
    // expression); points with a non-finite coordinate, and the padding, sort last
    static __global__ void k_capt_codes(const float *__restrict__ pts, uint32_t n, uint32_t n_pad, float ox, float oy, float oz, float inv0,
                                        uint32_t *__restrict__ code, uint32_t *__restrict__ idx)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n_pad)
        {
            return;
        }
        idx[i] = i;
        uint32_t c = 0xffffffffu;
        if (i < n)
        {
            const float x = pts[3u * i], y = pts[3u * i + 1u], z = pts[3u * i + 2u];
            const float big = 3.4028234664e38F;
            if (fabsf(x) <= big && fabsf(y) <= big && fabsf(z) <= big)
            {
                constexpr int kDim = 1 << kCaptGridBits;
                const int cx = max(0, min(kDim - 1, __float2int_rd((x - ox) * inv0)));
                const int cy = max(0, min(kDim - 1, __float2int_rd((y - oy) * inv0)));
                const int cz = max(0, min(kDim - 1, __float2int_rd((z - oz) * inv0)));
                c = morton_spread(cx) | (morton_spread(cy) << 1) | (morton_spread(cz) << 2);
            }
        }
        code[i] = c;
    }

    static __global__ void k_capt_grid_points(const float *__restrict__ pts, const uint32_t *__restrict__ order, const uint32_t *__restrict__ leaf_of_point,
                                              uint32_t m, float4 *__restrict__ gpts)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i < m)
        {
            const uint32_t id = order[i];
            gpts[i] = make_float4(pts[3u * id], pts[3u * id + 1u], pts[3u * id + 2u], __uint_as_float(leaf_of_point[id]));
        }
    }

    // start tables: for level l and cell code c the first sorted point whose code >> 3 l is >= c (one entry past the last cell)
    static __global__ void k_capt_grid_starts(const uint32_t *__restrict__ code, uint32_t m, uint32_t *__restrict__ gstart)
    {
        const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
        if (e >= capt_grid_offset(kCaptGridLevels))
        {
            return;
        }
        int level = 0;
        while (level + 1 < kCaptGridLevels && e >= capt_grid_offset(level + 1))
        {
            ++level;
        }
        const uint32_t c = e - capt_grid_offset(level);
        uint32_t lo = 0u, hi = m;  // first index with (code >> 3 level) >= c
        while (lo < hi)
        {
            const uint32_t mid = (lo + hi) / 2u;
            if ((code[mid] >> (3 * level)) < c)
            {
                lo = mid + 1u;
            }
            else
            {
                hi = mid;
            }
        }
        gstart[e] = lo;
    }
}  // namespace vmv
