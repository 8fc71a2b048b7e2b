// CenterVox pointcloud down-sampling on the device (sm_100a).
//
// Reference: vamp::collision::filter_pointcloud_centervox / CenterSelectiveVoxelFilter,
// src/impl/vamp/collision/filter_centervox.hh:16-45 (Voxel::try_insert), :139-165 (try_insert_point),
// :167-184 (extract_points), :251-291 (insert_to_voxel), :310-330 (driver loop); Python surface
// vamp.filter_pointcloud(..., filter_type="centervox"), bindings/environment.cc:212-239.
//
// The reference walks the cloud once, point by point: a point inside the range sphere and the workspace
// box goes to the voxel (vx, vy, vz) of a grid of at most 255^3; the voxel keeps the point closest to its
// centre, the EARLIER point on a tie; the result lists the occupied voxels in the order its sparse
// three-level table created them.  None of that needs the sequential walk:
//
//   * "closest, earliest on a tie" = the minimum of the 64-bit key (distance bits << 32 | point index)
//     over the voxel's points (squared distances are >= 0, so their float bits order like the values):
//     one atomicMin per surviving point (k_centervox_insert, thread = point; 12 B read per point is the
//     only streaming traffic, the key table lives in L2);
//   * creation order = lexicographic order of (first point index seen in the voxel's x slab, first in
//     its (x, y) column, first in the voxel): a second atomicMin keeps the voxel's first index, the slab
//     and column minima follow from the <= 32768 occupied voxels (the reference's pool size) on the host.
//
// Arithmetic is written out with explicit FMAs where GCC contracts the reference's expressions under its
// -ffp-contract=fast, so keys compare as they do there: dx*dx + dy*dy + dz*dz becomes
// fma(dz, dz, fma(dx, dx, round(dy*dy))) (the first product is fused onto the rounded second one -- read off
// the compiled reference, oracle/_ref), min + (v + 0.5) * size becomes one FMA.
#pragma once
#include <cstdint>

namespace vmv
{
    struct CenterVoxParams
    {
        float origin[3], ws_min[3], ws_max[3];
        float voxel_size, inv_scale, max_range_sq;
        int dim;  // table edge: coordinates are clamped to [0, 254] and never exceed grid_width
    };

    static constexpr unsigned long long kVoxEmpty = ~0ull;

    __device__ __forceinline__ int centervox_coord(float p, float lo, float inv_scale)
    {
        // static_cast<int>((p - min) * inverse_scale_factor) clamped to [0, MAX_GRID_SIZE - 1]
        const int v = __float2int_rz(__fmul_rn(__fsub_rn(p, lo), inv_scale));
        return min(max(v, 0), 254);
    }

    // Culling, voxel and squared distance to the voxel centre of one point.  Returns false if the point is
    // culled (filter_centervox.hh:141-151).  NaN coordinates fail no test in the reference (every
    // comparison is false) and are inserted -- same here.
    __device__ __forceinline__ bool centervox_point(const CenterVoxParams &P, const float *__restrict__ pts, uint32_t i, size_t &id, float &d)
    {
        const float x = __ldg(pts + 3 * static_cast<size_t>(i)), y = __ldg(pts + 3 * static_cast<size_t>(i) + 1), z = __ldg(pts + 3 * static_cast<size_t>(i) + 2);
        const float dx = __fsub_rn(x, P.origin[0]), dy = __fsub_rn(y, P.origin[1]), dz = __fsub_rn(z, P.origin[2]);
        const float range_sq = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        if (range_sq >= P.max_range_sq)
        {
            return false;
        }
        if (x < P.ws_min[0] || x > P.ws_max[0] || y < P.ws_min[1] || y > P.ws_max[1] || z < P.ws_min[2] || z > P.ws_max[2])
        {
            return false;
        }
        const int vx = centervox_coord(x, P.ws_min[0], P.inv_scale);
        const int vy = centervox_coord(y, P.ws_min[1], P.inv_scale);
        const int vz = centervox_coord(z, P.ws_min[2], P.inv_scale);
        // voxel centre (set_voxel_center, :22-26) and squared distance to it (try_insert, :28-32)
        const float cx = __fmaf_rn(static_cast<float>(vx) + 0.5F, P.voxel_size, P.ws_min[0]);
        const float cy = __fmaf_rn(static_cast<float>(vy) + 0.5F, P.voxel_size, P.ws_min[1]);
        const float cz = __fmaf_rn(static_cast<float>(vz) + 0.5F, P.voxel_size, P.ws_min[2]);
        const float ex = __fsub_rn(x, cx), ey = __fsub_rn(y, cy), ez = __fsub_rn(z, cz);
        d = __fmaf_rn(ez, ez, __fmaf_rn(ex, ex, __fmul_rn(ey, ey)));
        id = (static_cast<size_t>(vx) * P.dim + vy) * P.dim + vz;
        return true;
    }

    __global__ void __launch_bounds__(256) k_centervox_insert(
        const __grid_constant__ CenterVoxParams P, const float *__restrict__ pts, uint32_t n, unsigned long long *__restrict__ best, uint32_t *__restrict__ first)
    {
        const uint32_t stride = gridDim.x * blockDim.x;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        {
            size_t id;
            float d;
            if (!centervox_point(P, pts, i, id, d))
            {
                continue;
            }
            // a NaN distance never wins `new < stored`; its key is above every real distance
            const uint32_t bits = (d == d) ? __float_as_uint(d) : 0xffffffffu;
            // most points lose against what their voxel already holds: a plain (L2) read first keeps them off
            // the atomic units (the value only ever decreases, so a stale read is merely conservative)
            const unsigned long long key = (static_cast<unsigned long long>(bits) << 32) | i;
            if (key < __ldcg(best + id))
            {
                atomicMin(best + id, key);
            }
            if (i < __ldcg(first + id))
            {
                atomicMin(first + id, i);
            }
        }
    }

    // thread = voxel: occupied voxels -> (voxel id, kept point, first point) records, any order
    __global__ void __launch_bounds__(256) k_centervox_compact(
        const __grid_constant__ CenterVoxParams P,
        const float *__restrict__ pts,
        const unsigned long long *__restrict__ best,
        const uint32_t *__restrict__ first,
        size_t n_vox,
        uint32_t cap,
        uint32_t *__restrict__ out,
        uint32_t *__restrict__ count)
    {
        const size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
        if (v >= n_vox)
        {
            return;
        }
        const unsigned long long key = best[v];
        if (key == kVoxEmpty)
        {
            return;
        }
        uint32_t keep = static_cast<uint32_t>(key & 0xffffffffull);
        const uint32_t f = first[v];
        {
            // a voxel whose FIRST point has a NaN distance keeps it for good: no later `new < NaN` holds
            size_t id;
            float d;
            if (centervox_point(P, pts, f, id, d) && !(d == d))
            {
                keep = f;
            }
        }
        const uint32_t at = atomicAdd(count, 1u);
        if (at < cap)
        {
            out[3 * at] = static_cast<uint32_t>(v);
            out[3 * at + 1] = keep;
            out[3 * at + 2] = f;
        }
    }
}  // namespace vmv
