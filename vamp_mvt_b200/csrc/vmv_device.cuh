// Device-side building blocks of the validation engine (sm_100a).
//
//  * sincos_f32         -- trigonometry with the reference's numerics (its cephes variant has up to
//                          4e-6 absolute error, which defines the reference's FK; we reproduce the
//                          same range reduction and polynomials so FK centres agree to ~5e-7 m)
//  * PackedEnv          -- layout of the environment blob staged into shared memory by one TMA bulk copy
//  * sphere_hits_env    -- one robot sphere against every container, sorted early-out
//                          (reference collision/validity.hh:46-158)
//  * capt_collides      -- CAPT descent + affordance scan (reference collision/capt.hh:428-512)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "robot_tables.h"

namespace vmv
{
    // ------------------------------------------------------------------------------------------
    // packed environment (all offsets in 4-byte words from the start of the blob; every record
    // array is 16-byte aligned so records are read with 128-bit shared-memory loads)
    // ------------------------------------------------------------------------------------------
    struct EnvHeader
    {
        uint32_t n_spheres, n_capsules, n_zcapsules, n_cuboids;
        uint32_t n_zcuboids, n_heightfields, n_capts, n_attach;
        uint32_t off_spheres, off_capsules, off_zcapsules, off_cuboids;
        uint32_t off_zcuboids, off_heightfields, off_capts, off_attach;
        uint32_t n_mvts, off_mvts, off_cloud_grid, pad1;  // off_cloud_grid = 0: no nearest-point table
    };

    static constexpr int kSphereRec = 8;     // {x y z r}{min_d 0 0 0}
    static constexpr int kCapsuleRec = 12;   // {x1 y1 z1 r}{xv yv zv rdv}{min_d 0 0 0}
    static constexpr int kZCapsuleRec = 8;   // {x1 y1 z1 r}{zv rdv min_d 0}
    static constexpr int kCuboidRec = 16;    // {x y z r1}{a1x a1y a1z r2}{a2x a2y a2z r3}{a3x a3y a3z min_d}
    static constexpr int kZCuboidRec = 12;   // {x y z min_d}{a1x a1y a2x a2y}{r1 r2 r3 0}
    static constexpr int kHeightRec = 12;    // {x y z xs}{ys zs xd yd}{xd2 yd2 ptr_lo ptr_hi}
    static constexpr int kCaptRec = 28;      // see CaptRec
    static constexpr int kAttachHdr = 4;     // {n 0 0 0} then n x {x y z r} (attachment-local)

    struct CaptRec
    {
        float r_point;
        uint32_t nlog2, n_tests, n_points;
        float lo[3];              // box of the finite points (top-level reject)
        float r_max;
        float hi[3];
        float list_reach_sq;      // (r_max + r_point)^2 as the build computes it
        float g_origin[3];        // enumeration grid: origin, cells per metre and cell edge at the finest level
        float g_inv0;
        float g_cell0;
        uint32_t n_grid_points;   // entries of gpts
        uint32_t pad[2];
        const float2 *nodes;      // Eytzinger, 2^nlog2 - 1: {split value, 1 = the high half inherited the low half's points}
        const uint32_t *leafbits; // two bits per leaf: 1 = carries a list beyond its representative, 2 = representative is finite
        const float4 *gpts;       // the finite points in Morton order of their finest cell: {x y z leaf}
        const uint32_t *gstart;   // per level: first point of every cell in Morton order (+ one end entry), capt_grid_offset()
    };
    static_assert(sizeof(CaptRec) == kCaptRec * 4, "CaptRec layout");

    // Multi-level Voxel Table (reference collision/mvt.hh).  The reference's three sparse pointer
    // levels (X -> Y -> Z -> voxel index) address a cube of grid_width^3 cells; in HBM that is one
    // dense u32 array (a missing Y or Z table is a run of empty cells), per voxel two float4
    // {bbox lo, bbox hi.x}{bbox hi.y, hi.z, start, count}, and the points as float4, grouped by voxel.
    static constexpr int kMvtRec = 24;
    struct MvtRec
    {
        float r_point, inv_scale;
        uint32_t grid_width, pad0;
        float ws_min[3];
        float pad1;
        float g_min[3];
        float pad2;
        float g_max[3];
        float pad3;
        const uint32_t *cells;
        const float4 *voxels;
        const float4 *points;
        const void *pad4;
    };
    static_assert(sizeof(MvtRec) == kMvtRec * 4, "MvtRec layout");

    // Nearest-point table of the pointclouds (CAPT and MVT together): per voxel the cloud point nearest to the voxel
    // CENTRE, {x, y, z, tag}; tag = cloud << 24 | leaf (the CAPT leaf the point is the representative of; cloud
    // 0xff = a point of an MVT or of a tree too large for the tag).  Two uses:
    //  * clearance: the distance function is 1-Lipschitz, so a position at offset d from the centre is at least
    //    (|nearest - centre| - d) away from every point.  A pointcloud query can only answer "collision" on an actual
    //    point within r + r_point of the centre (capt.hh:494-509, mvt.hh:383-397), so a sphere whose clearance bound
    //    exceeds that radius skips the tree descent / voxel walk altogether -- the verdict is the one the query would
    //    have returned;
    //  * a proven hit: when the table's point itself lies within r + r_point of the centre AND is on the affordance
    //    list of the leaf the centre descends to (capt_lists_point evaluates the reference's list construction for
    //    that one point), the reference's list scan would find it: "collision" without any list.
    static constexpr int kCloudGridRec = 12;
    static constexpr uint32_t kCloudTagNone = 0xffu;
    struct CloudGridRec
    {
        float x0, y0, z0, inv_h;
        int nx, ny, nz;
        float outside;      // clearance bound for positions outside the table
        float r_point_max;  // largest r_point of any cloud
        float h;            // voxel edge
        const float4 *cells;
    };
    static_assert(sizeof(CloudGridRec) == kCloudGridRec * 4, "CloudGridRec layout");

    // The table entry of the voxel (x, y, z) lies in and the clearance bound it gives; false = outside the table
    // (near: untouched, clearance: the table's bound for the outside, 0 for a centre that is not finite).
    // (not inlined, like the other per-link environment tests below: the any-environment kernels call them once per link
    // from straight-line code and are bound by instruction fetch)
    static __device__ __noinline__ bool cloud_probe(const CloudGridRec &g, float x, float y, float z, float4 &near, float &clearance)
    {
        const float fx = (x - g.x0) * g.inv_h, fy = (y - g.y0) * g.inv_h, fz = (z - g.z0) * g.inv_h;
        const int ix = __float2int_rd(fx), iy = __float2int_rd(fy), iz = __float2int_rd(fz);
        const bool in = (static_cast<unsigned>(ix) < static_cast<unsigned>(g.nx)) & (static_cast<unsigned>(iy) < static_cast<unsigned>(g.ny)) &
                        (static_cast<unsigned>(iz) < static_cast<unsigned>(g.nz));
        if (!in)
        {
            // a centre that is not finite lands here too: no claim about it
            clearance = (fabsf(x) < 1e30F && fabsf(y) < 1e30F && fabsf(z) < 1e30F) ? g.outside : 0.F;
            return false;
        }
        near = __ldg(g.cells + (static_cast<size_t>(iz) * g.ny + iy) * g.nx + ix);
        // offset of the position from the voxel centre, in metres, rounded up a little; the centre as the build kernel
        // computes it; the slack (1e-4 m absolute) covers the rounding of all of it
        const float cx = (static_cast<float>(ix) + 0.5F), cy = (static_cast<float>(iy) + 0.5F), cz = (static_cast<float>(iz) + 0.5F);
        const float ox = fx - cx, oy = fy - cy, oz = fz - cz;
        const float off = sqrtf(ox * ox + oy * oy + oz * oz) * g.h * 1.0001F;
        const float dx = near.x - (g.x0 + cx * g.h), dy = near.y - (g.y0 + cy * g.h), dz = near.z - (g.z0 + cz * g.h);
        clearance = sqrtf(dx * dx + dy * dy + dz * dz) * 0.99999F - 1e-4F - off;
        return true;
    }

    __device__ __forceinline__ float cloud_clearance(const CloudGridRec &g, float x, float y, float z)
    {
        float4 near;
        float c;
        cloud_probe(g, x, y, z, near, c);
        return c;
    }

    // Table build.  Block = a brick of 8 x 8 x 4 voxels, thread = voxel.  The points arrive sorted along a
    // Morton curve (host), in tiles of kCloudTile with their bounding boxes (256: measured 27 ms against 37 ms with 1024): a tile whose box is farther from every
    // voxel of the brick than what that voxel has already found is skipped without being loaded (the first
    // point of every tile seeds the bounds).  Without the cull this is |voxels| x |points| distance
    // evaluations -- 1.3 s for 2^24 voxels and 10^5 points.
#ifndef VMV_CLOUD_TILE
#define VMV_CLOUD_TILE 256
#endif
    static constexpr int kCloudTile = VMV_CLOUD_TILE;

    static __global__ void __launch_bounds__(256) k_build_cloud_grid(
        const float4 *__restrict__ points,
        uint32_t n_points,
        const float4 *__restrict__ tile_boxes,  // per tile {lo.xyz, hi.x}, {hi.yz, -, -}
        float x0,
        float y0,
        float z0,
        float h,
        int nx,
        int ny,
        int nz,
        float4 *__restrict__ out)
    {
        __shared__ float4 tile[kCloudTile];
        const int bxn = (nx + 7) / 8, byn = (ny + 7) / 8;
        const int brick = blockIdx.x;
        const int bx = brick % bxn, by = (brick / bxn) % byn, bz = brick / (bxn * byn);
        const int t = threadIdx.x;
        const int ix = bx * 8 + (t & 7), iy = by * 8 + ((t >> 3) & 7), iz = bz * 4 + (t >> 6);
        const float x = x0 + (ix + 0.5F) * h, y = y0 + (iy + 0.5F) * h, z = z0 + (iz + 0.5F) * h;
        const uint32_t n_tiles = (n_points + kCloudTile - 1) / kCloudTile;
        float best = 3.0e38F;
        uint32_t best_i = 0u;
        for (uint32_t k = 0; k < n_tiles; ++k)
        {
            const float4 p = __ldg(points + static_cast<size_t>(k) * kCloudTile);
            const float dx = p.x - x, dy = p.y - y, dz = p.z - z;
            const float d = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            best_i = d < best ? k * kCloudTile : best_i;
            best = fminf(best, d);
        }
        for (uint32_t k = 0; k < n_tiles; ++k)
        {
            const float4 b0 = __ldg(tile_boxes + 2 * k), b1 = __ldg(tile_boxes + 2 * k + 1);
            const float ex = x - fminf(fmaxf(x, b0.x), b0.w), ey = y - fminf(fmaxf(y, b0.y), b1.x), ez = z - fminf(fmaxf(z, b0.z), b1.y);
            // (0.999: the box distance must not exceed the distance to any point in it after rounding)
            const bool need = (ex * ex + ey * ey + ez * ez) * 0.999F < best;
            if (!__syncthreads_or(need))
            {
                continue;
            }
            const uint32_t base = k * kCloudTile;
            for (uint32_t i = t; i < kCloudTile; i += blockDim.x)
            {
                tile[i] = base + i < n_points ? __ldg(points + base + i) : make_float4(1e18F, 1e18F, 1e18F, 0.F);
            }
            __syncthreads();
#pragma unroll 8
            for (int i = 0; i < kCloudTile; ++i)
            {
                const float4 p = tile[i];
                const float dx = p.x - x, dy = p.y - y, dz = p.z - z;
                const float d = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                best_i = d < best ? base + i : best_i;
                best = fminf(best, d);
            }
        }
        if (ix < nx && iy < ny && iz < nz)
        {
            out[(static_cast<size_t>(iz) * ny + iy) * nx + ix] = __ldg(points + best_i);  // {x y z tag}
        }
    }

    // development counters (tools/c4_stats.py builds with -DVMV_C4_STATS): where the any-environment kernel spends its queries
#ifdef VMV_C4_STATS
    static __device__ unsigned long long *g_stats = nullptr;
#define VMV_STAT(i, v) atomicAdd(g_stats + (i), static_cast<unsigned long long>(v))
#else
#define VMV_STAT(i, v) ((void)0)
#endif

    __device__ __forceinline__ bool sign_set(float v)
    {
        return __float_as_int(v) < 0;
    }

    // ------------------------------------------------------------------------------------------
    // trigonometry (reference vector/avx.hh:454-548 and vector/interface.hh:451-458)
    // ------------------------------------------------------------------------------------------
    __device__ __forceinline__ float ref_sin(float x)
    {
        const uint32_t sign_in = __float_as_uint(x) & 0x80000000u;
        x = fabsf(x);
        // j = (rint(x * 4/pi) + 1) & ~1   (the reference rounds -- cvtps -- where cephes truncates)
        int j = __float2int_rn(x * 1.27323954473516f);
        j = (j + 1) & ~1;
        const float y = __int2float_rn(j);
        const uint32_t sign = sign_in ^ (static_cast<uint32_t>(j & 4) << 29);
        x = fmaf(y, -0.78515625f, x);
        x = fmaf(y, -2.4187564849853515625e-4f, x);
        x = fmaf(y, -3.77489497744594108e-8f, x);
        const float z = x * x;
        const bool use_sin = (j & 2) == 0;
        // both polynomials share the Horner shape ((k0 z + k1) z + k2) z
        const float k0 = use_sin ? -1.9515295891E-4f : 2.443315711809948E-005f;
        const float k1 = use_sin ? 8.3321608736E-3f : -1.388731625493765E-003f;
        const float k2 = use_sin ? -1.6666654611E-1f : 4.166664568298827E-002f;
        float p = fmaf(k0, z, k1);
        p = fmaf(p, z, k2);
        p = p * z;
        // sin: p*x + x ; cos: p*z - z/2 + 1
        const float a = use_sin ? x : z;
        const float b = use_sin ? x : fmaf(z, -0.5f, 1.0f);
        const float r = fmaf(p, a, b);
        return __uint_as_float(__float_as_uint(r) ^ sign);
    }

    // not inlined: a robot's FK calls it once per joint, and the straight-line FK bodies are already
    // large enough to pressure the instruction cache when warps run out of step
#ifdef VMV_SINCOS_INLINE
    __device__ __forceinline__ void sincos_f32(float x, float &s, float &c)
#else
    static __device__ __noinline__ void sincos_f32(float x, float &s, float &c)
#endif
    {
        s = ref_sin(x);
        // cos(x) = sin(x + pi/2), wrapped by 2 pi when the sum reaches pi (interface.hh:451-458)
        float v = x + 1.57079637050628662109375f;
        v = (v >= 3.1415927410125732421875f) ? v - 6.283185482025146484375f : v;
        c = ref_sin(v);
    }

    // ------------------------------------------------------------------------------------------
    // CAPT query (reference collision/capt.hh:428-512) WITHOUT the affordance lists.
    //
    // The reference answers "is any point of the list of the centre's leaf within r + r_point of the centre"; the
    // list of a leaf is a function of the k-d tree alone (capt.hh:106-119 median split, :150-170 leaf step, :221-246
    // what each half inherits), and for ONE point p and the leaf L of a centre c that function is cheap:
    //  * p is L's own representative, or L carries a list at all (a cell inside the smallest query ball around its
    //    representative keeps the representative alone -- one bit per leaf) and
    //  * at the node where the root paths of p and c part, p's half was handed to c's half: the low half always
    //    receives the high half's points within r_max of the plane, but the high half receives the low half's points
    //    only if the SMALLEST of them is within r_max of the plane (the reference scans the sorted half from its
    //    first element while the predicate holds) -- one bit per node, collected along c's descent into a mask that
    //    is indexed by the highest differing bit of the two leaf numbers;
    //  * below that node p is an inherited candidate, kept on c's side while within r_max of every plane: planes nest,
    //    so that is "within r_max of c's cell, axis by axis" (for the planes above the parting node it holds trivially);
    //  * at the leaf, p lies within r_max + r_point of the cell (the build's own float operations, bit for bit).
    // So the lists -- 7 GB and a 4 s host build for 10^5 points at r_max = 0.24 m -- are not stored.  Candidates come
    // from (a) the nearest-point table (cloud_probe): a sphere in contact usually contains the table's point, and
    // (b) a Morton-ordered uniform grid over the raw points, five levels of cell size, enumerated by the whole group
    // for one query at a time over the box (query ball) x (cell grown by r_max).
    // ------------------------------------------------------------------------------------------
    static constexpr int kCaptGridBits = 7;    // finest level: 128 cells per axis
    static constexpr int kCaptGridLevels = 5;  // cell edge g0 * 2^l
    // independent 128-bit loads per lane and step of the enumeration (measured on BASELINE config 4: 2 -> Fetch 0.578 /
    // UR5 0.452 ms, 4 -> 0.621 / 0.473, 8 -> 0.759 / 0.602: cells hold a few dozen points, wider steps load padding)
    static constexpr uint32_t kCaptEnumLoads = 2;

    __host__ __device__ __forceinline__ uint32_t capt_grid_offset(int level)
    {
        // levels are stored one after the other, 8^(bits - l) + 1 entries each
        uint32_t off = 0;
        for (int l = 0; l < level; ++l)
        {
            off += (1u << (3 * (kCaptGridBits - l))) + 1u;
        }
        return off;
    }

    __host__ __device__ __forceinline__ uint32_t morton_spread(uint32_t v)
    {
        v &= 0x3ffu;
        v = (v | (v << 16)) & 0x030000ffu;
        v = (v | (v << 8)) & 0x0300f00fu;
        v = (v | (v << 4)) & 0x030c30c3u;
        v = (v | (v << 2)) & 0x09249249u;
        return v;
    }

    struct CaptCell
    {
        uint32_t leaf, hmask;  // hmask bit b: c took the high side at the level whose leaf-number bit is b, and that node's high half did NOT inherit
        float lo[3], hi[3];    // the leaf's cell
        bool full, finite;
    };

    __device__ __forceinline__ void capt_descend(const CaptRec &t, float x, float y, float z, CaptCell &c)
    {
        const float inf = __int_as_float(0x7f800000);
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            c.lo[k] = -inf;
            c.hi[k] = inf;
        }
        c.hmask = 0u;
        uint32_t idx = 0u, bit = t.nlog2;
        for (uint32_t i = 0; i < t.nlog2; i += 3)
        {
#pragma unroll
            for (int k = 0; k < 3; ++k)
            {
                if (i + k < t.nlog2)
                {
                    const float2 node = __ldg(t.nodes + idx);  // {split, inherited}
                    const float v = (k == 0) ? x : ((k == 1) ? y : z);
                    const bool up = v >= node.x;
                    --bit;
                    if (up)
                    {
                        c.lo[k] = node.x;
                        c.hmask |= (__float_as_uint(node.y) ^ 1u) << bit;
                    }
                    else
                    {
                        c.hi[k] = node.x;
                    }
                    idx = 2u * idx + 1u + (up ? 1u : 0u);
                }
            }
        }
        c.leaf = idx - t.n_tests;
        const uint32_t w = __ldg(t.leafbits + (c.leaf >> 4)) >> (2u * (c.leaf & 15u));
        c.full = (w & 1u) != 0u;
        c.finite = (w & 2u) != 0u;
    }

    // Is the cloud point p = {x y z leaf} on the list of the leaf described by c?
    __device__ __forceinline__ bool capt_member(const CaptRec &t, const CaptCell &c, const float4 &p)
    {
        const uint32_t parted = c.leaf ^ (__float_as_uint(p.w) & 0xffffffu);
        if (parted == 0u)
        {
            return true;  // the leaf's representative
        }
        const bool handed = ((c.hmask >> (31 - __clz(parted))) & 1u) == 0u;
        const float pc[3] = {p.x, p.y, p.z};
        bool near = true;
        float dsq = 0.F;
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            near = near && (pc[k] >= __fsub_rn(c.lo[k], t.r_max)) && (pc[k] <= __fadd_rn(c.hi[k], t.r_max));
            const float d = __fsub_rn(pc[k], fminf(fmaxf(pc[k], c.lo[k]), c.hi[k]));
            dsq = __fadd_rn(dsq, __fmul_rn(d, d));
        }
        return c.full && handed && near && (dsq <= t.list_reach_sq);
    }

    // Called by whichever lanes of the warp currently have a query (the converged group, __activemask()); `active`
    // false = the lane has nothing to ask but helps.  `near` / `has_near`: this cloud's point from the nearest-point
    // table for the lane's position.
    // (not inlined: the any-environment kernels call it from two places and are bound by instruction fetch as it is)
    static __device__ __noinline__ bool
    capt_collides_warp(const CaptRec &t, float x, float y, float z, float r, bool active, const float4 &near, bool has_near)
    {
        const uint32_t group = __activemask();
        const int lane = threadIdx.x & 31;
        const int rank = __popc(group & ((1u << lane) - 1u));  // position of this lane inside the group
        const int gsize = __popc(group);
        CaptCell c = {};
        bool need = false, hit = false;
#ifdef VMV_C4_STATS
        const long long t_call = clock64();
#endif
        const float rr = r + t.r_point;
        const float rc_sq = rr * rr;
        if (active)
        {
            VMV_STAT(4, 1);
            // top-level AABB reject uses r, not r + r_point (capt.hh:431-438)
            const bool inb = (x + r >= t.lo[0]) & (x - r <= t.hi[0]) & (y + r >= t.lo[1]) & (y - r <= t.hi[1]) & (z + r >= t.lo[2]) &
                             (z - r <= t.hi[2]);
            if (inb)
            {
                capt_descend(t, x, y, z, c);
                if (c.finite)
                {
                    need = true;
                    if (has_near)
                    {
                        const float ex = near.x - x, ey = near.y - y, ez = near.z - z;
                        if ((ex * ex + ey * ey + ez * ez <= rc_sq) && capt_member(t, c, near))
                        {
                            hit = true;
                            need = false;
                            VMV_STAT(12, 1);
                        }
                    }
                }
            }
            VMV_STAT(5, need ? 1 : 0);
        }

        uint32_t pending = __ballot_sync(group, need);
#ifdef VMV_C4_STATS
        const long long t_enum = clock64();
        if (rank == 0)
        {
            VMV_STAT(11, 1);
            VMV_STAT(10, gsize);
        }
#endif
        while (pending != 0u)
        {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1u;
            // the query, for everybody
            CaptCell q;
            q.leaf = __shfl_sync(group, c.leaf, src);
            q.hmask = __shfl_sync(group, c.hmask, src);
            q.full = __shfl_sync(group, c.full ? 1 : 0, src) != 0;
            const float qc[3] = {__shfl_sync(group, x, src), __shfl_sync(group, y, src), __shfl_sync(group, z, src)};
            const float qrr = __shfl_sync(group, rr, src);
            const float qr = qrr * qrr;
            float blo[3], bhi[3];
            bool empty = false;
            float ext = 0.F;
#pragma unroll
            for (int k = 0; k < 3; ++k)
            {
                q.lo[k] = __shfl_sync(group, c.lo[k], src);
                q.hi[k] = __shfl_sync(group, c.hi[k], src);
                // where a listed point within the ball can be (a cull: the exact tests follow per point)
                blo[k] = fmaxf(qc[k] - qrr, q.lo[k] - t.r_max) - 1e-4F;
                bhi[k] = fminf(qc[k] + qrr, q.hi[k] + t.r_max) + 1e-4F;
                empty = empty || (blo[k] > bhi[k]);
                ext = fmaxf(ext, bhi[k] - blo[k]);
            }
            // the level at which the box is at most four (six) cells wide
            int level = 0;
            float inv = t.g_inv0;
                        // (wide boxes -- a link's bounding sphere in a big free-space cell -- take finer cells, six across: fewer
            // points outside the ball are read; measured against a fixed span of 3 and of 5 on BASELINE config 4)
            const float span = ext > 0.25F ? 5.F : 3.F;
            while (level < kCaptGridLevels - 1 && ext * inv > span)
            {
                ++level;
                inv *= 0.5F;
            }
            const int dim = (1 << kCaptGridBits) >> level;
            int i0[3], n[3];
#pragma unroll
            for (int k = 0; k < 3; ++k)
            {
                const float a = (blo[k] - t.g_origin[k]) * inv, b = (bhi[k] - t.g_origin[k]) * inv;
                empty = empty || (b < 0.F) || (a >= static_cast<float>(dim));
                const int ia = max(0, min(dim - 1, __float2int_rd(a))), ib = max(0, min(dim - 1, __float2int_rd(b)));
                i0[k] = ia;
                n[k] = ib - ia + 1;
            }
            const int n_cells = empty ? 0 : n[0] * n[1] * n[2];
            const uint32_t *starts = t.gstart + capt_grid_offset(level);
            const float g = t.g_cell0 * static_cast<float>(1 << level);
            bool found = false;
            VMV_STAT(6, rank == 0 ? 1 : 0);
            for (int cb = 0; cb < n_cells && !found; cb += gsize)
            {
                const int ci = cb + rank;
                uint32_t s = 0u, e = 0u;
                if (ci < n_cells)
                {
                    const int cx = i0[0] + ci % n[0], cy = i0[1] + (ci / n[0]) % n[1], cz = i0[2] + ci / (n[0] * n[1]);
                    // the cell against the ball (cell bounds grown by what their float arithmetic can be off)
                    const int cc[3] = {cx, cy, cz};
                    float dsq = 0.F;
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                    {
                        const float lo_k = t.g_origin[k] + static_cast<float>(cc[k]) * g - 1e-4F;
                        const float hi_k = lo_k + g + 2e-4F;
                        const float d = qc[k] - fminf(fmaxf(qc[k], lo_k), hi_k);
                        dsq += d * d;
                    }
                    if (dsq <= qr * 1.0001F)
                    {
                        const uint32_t m = (morton_spread(cx) | (morton_spread(cy) << 1) | (morton_spread(cz) << 2));
                        s = __ldg(starts + m);
                        e = __ldg(starts + m + 1);
#ifdef VMV_C4_STATS
                        // development build: every index this query derives must be inside its table
                        if (m + 1u >= capt_grid_offset(level + 1) - capt_grid_offset(level) || s > e || e > t.n_grid_points ||
                            cx >= dim || cy >= dim || cz >= dim)
                        {
                            VMV_STAT(63, 1);
                        }
#endif
                    }
                }
                uint32_t have = __ballot_sync(group, e > s);
                if (rank == 0)
                {
                    VMV_STAT(16, 1);              // cell-lookup rounds
                    VMV_STAT(17, __popc(have));   // cells with points, inside the ball's reach
                    VMV_STAT(18, n_cells);
                }
                while (have != 0u && !found)
                {
                    const int cl = __ffs(have) - 1;
                    have &= have - 1u;
                    const uint32_t cs = __shfl_sync(group, s, cl), ce = __shfl_sync(group, e, cl);
                    for (uint32_t base = cs; base < ce; base += kCaptEnumLoads * gsize)
                    {
                        // all loads first, every one from a valid address (a slot past the end re-reads the cell's first point)
                        float4 p[kCaptEnumLoads];
#pragma unroll
                        for (uint32_t u = 0; u < kCaptEnumLoads; ++u)
                        {
                            const uint32_t i = base + u * gsize + rank;
                            p[u] = __ldg(t.gpts + (i < ce ? i : cs));
                        }
                        bool h = false;
#pragma unroll
                        for (uint32_t u = 0; u < kCaptEnumLoads; ++u)
                        {
                            const float ex = p[u].x - qc[0], ey = p[u].y - qc[1], ez = p[u].z - qc[2];
                            h = h || ((ex * ex + ey * ey + ez * ez <= qr) && capt_member(t, q, p[u]));
                        }
                        if (rank == 0)
                        {
                            VMV_STAT(7, 1);
                            VMV_STAT(8, min(ce - base, kCaptEnumLoads * gsize));
                        }
                        if (__any_sync(group, h))
                        {
                            found = true;
                            break;
                        }
                    }
                }
            }
            if (rank == 0)
            {
                VMV_STAT(9, found ? 1 : 0);
            }
            hit = (lane == src) ? found : hit;
        }
#ifdef VMV_C4_STATS
        if (rank == 0)
        {
            const long long t_end = clock64();
            VMV_STAT(14, t_end - t_call);
            VMV_STAT(15, t_end - t_enum);
        }
#endif
        return hit;
    }

    // ------------------------------------------------------------------------------------------
    // MVT query (reference collision/mvt.hh:205-276; collides_simd :279-403 is this per lane): the
    // voxels within +-min(1, q / cell) of the centre's cell, bbox-culled, then every point.  The +-1
    // clamp is part of the semantics: a link's bounding sphere wider than a cell does not see points
    // two cells away, and then its fine spheres are not swept -- as in the reference.
    // ------------------------------------------------------------------------------------------
    __device__ __forceinline__ int mvt_cell(float v)
    {
        // static_cast<uint16_t>(float) on clamped arguments: truncation (negative stays 0)
        return v <= 0.F ? 0 : (v >= 65535.F ? 65535 : static_cast<int>(v));
    }

    static __device__ __noinline__ bool mvt_collides(const MvtRec &t, float x, float y, float z, float r)
    {
        const float q = r + t.r_point, q2 = q * q;
        if (x + q < t.g_min[0] || x - q > t.g_max[0] || y + q < t.g_min[1] || y - q > t.g_max[1] || z + q < t.g_min[2] ||
            z - q > t.g_max[2])
        {
            return false;
        }
        const float gq = fminf(1.F, q * t.inv_scale);
        const float top = static_cast<float>(t.grid_width - 1);
        const float gx = (x - t.ws_min[0]) * t.inv_scale, gy = (y - t.ws_min[1]) * t.inv_scale, gz = (z - t.ws_min[2]) * t.inv_scale;
        const int x0 = mvt_cell(fmaxf(0.F, gx - gq)), x1 = mvt_cell(fminf(top, gx + gq));
        const int y0 = mvt_cell(fmaxf(0.F, gy - gq)), y1 = mvt_cell(fminf(top, gy + gq));
        const int z0 = mvt_cell(fmaxf(0.F, gz - gq)), z1 = mvt_cell(fminf(top, gz + gq));
        const size_t gw = t.grid_width;
        for (int vx = x0; vx <= x1; ++vx)
        {
            for (int vy = y0; vy <= y1; ++vy)
            {
                for (int vz = z0; vz <= z1; ++vz)
                {
                    const uint32_t v = __ldg(t.cells + (static_cast<size_t>(vx) * gw + vy) * gw + vz);
                    if (v == 0xffffffffu)
                    {
                        continue;
                    }
                    const float4 b0 = __ldg(t.voxels + 2 * static_cast<size_t>(v));
                    const float4 b1 = __ldg(t.voxels + 2 * static_cast<size_t>(v) + 1);
                    if (x + q < b0.x || x - q > b0.w || y + q < b0.y || y - q > b1.x || z + q < b0.z || z - q > b1.y)
                    {
                        continue;
                    }
                    const uint32_t s = __float_as_uint(b1.z), e = s + __float_as_uint(b1.w);
                    for (uint32_t i = s; i < e; ++i)
                    {
                        const float4 p = __ldg(t.points + i);
                        const float dx = x - p.x, dy = y - p.y, dz = z - p.z;
                        if (dx * dx + dy * dy + dz * dz <= q2)
                        {
                            return true;
                        }
                    }
                }
            }
        }
        return false;
    }

    // ------------------------------------------------------------------------------------------
    // one robot sphere against the whole environment
    // ------------------------------------------------------------------------------------------
    // E points at the blob in shared memory.  Containers are swept in the reference's order; inside
    // a container objects are sorted by min_distance (distance of the object's nearest point from
    // the world origin) and the sweep stops at the first object that cannot be reached:
    // min_distance >= |p| + r.  The reference evaluates |p| with a 12-bit rsqrt and so sometimes
    // stops one object early or late; we use a (slightly inflated) accurate bound, so the set of
    // objects we test is a superset of every object that can actually touch the sphere.
    // r_pc is the radius used for pointcloud queries (see check_state: a link's bounding sphere is
    // queried with its exact, un-inflated radius there).
    // heightfields for one sphere (reference collision/sphere_heightfield.hh:9-30); no warp-level operations
    static __device__ __noinline__ bool sphere_hits_heightfields(const float *__restrict__ E, float x, float y, float z, float r)
    {
        const EnvHeader &H = *reinterpret_cast<const EnvHeader *>(E);
        for (uint32_t i = 0; i < H.n_heightfields; ++i)
        {
            const float4 *p = reinterpret_cast<const float4 *>(E + H.off_heightfields + kHeightRec * i);
            const float4 a = p[0], b = p[1], c = p[2];
            const float xo = a.x - x, yo = a.y - y;
            const float xi = floorf(fminf(fmaxf(fmaf(a.w, xo, c.x), 0.F), b.z));
            const float yi = floorf(fminf(fmaxf(fmaf(b.x, yo, c.y), 0.F), b.w));
            const int index = __float2int_rn(fmaf(yi, b.z, xi));
            const float *data = reinterpret_cast<const float *>(
                (static_cast<unsigned long long>(__float_as_uint(c.w)) << 32) | __float_as_uint(c.z));
            const float zh = __ldg(data + index);
            if (sign_set(z - r - fmaf(b.y, zh, a.z)))
            {
                return true;
            }
        }
        return false;
    }

    // spheres, capsules, cuboids and heightfields for one sphere (no warp-level operations)
    __device__ __forceinline__ bool sphere_hits_primitives(const float *__restrict__ E, float x, float y, float z, float r)
    {
        const EnvHeader &H = *reinterpret_cast<const EnvHeader *>(E);
        const float ext = fmaf(__fsqrt_rn(fmaf(x, x, fmaf(y, y, z * z))), 1.0000002f, r);

        {
            const float4 *p = reinterpret_cast<const float4 *>(E + H.off_spheres);
            for (uint32_t i = 0; i < H.n_spheres; ++i)
            {
                const float4 a = p[2 * i];
                const float min_d = E[H.off_spheres + kSphereRec * i + 4];
                if (!(min_d < ext))
                {
                    break;
                }
                const float dx = a.x - x, dy = a.y - y, dz = a.z - z;
                const float rs = a.w + r;
                if (sign_set((dx * dx + dy * dy + dz * dz) - rs * rs))
                {
                    return true;
                }
            }
        }

        {
            const float4 *p = reinterpret_cast<const float4 *>(E + H.off_capsules);
            for (uint32_t i = 0; i < H.n_capsules; ++i)
            {
                const float min_d = E[H.off_capsules + kCapsuleRec * i + 8];
                if (!(min_d < ext))
                {
                    break;
                }
                const float4 a = p[3 * i], v = p[3 * i + 1];
                const float dot = (x - a.x) * v.x + (y - a.y) * v.y + (z - a.z) * v.z;
                const float cdf = fminf(fmaxf(dot * v.w, 0.F), 1.F);
                const float dx = x - (a.x + v.x * cdf), dy = y - (a.y + v.y * cdf), dz = z - (a.z + v.z * cdf);
                const float rs = r + a.w;
                if (sign_set((dx * dx + dy * dy + dz * dz) - rs * rs))
                {
                    return true;
                }
            }
        }

        {
            const float4 *p = reinterpret_cast<const float4 *>(E + H.off_zcapsules);
            for (uint32_t i = 0; i < H.n_zcapsules; ++i)
            {
                const float4 a = p[2 * i], v = p[2 * i + 1];
                if (!(v.z < ext))
                {
                    break;
                }
                const float dot = (z - a.z) * v.x;
                const float cdf = fminf(fmaxf(dot * v.y, 0.F), 1.F);
                const float dx = x - a.x, dy = y - a.y, dz = z - (a.z + v.x * cdf);
                const float rs = r + a.w;
                if (sign_set((dx * dx + dy * dy + dz * dz) - rs * rs))
                {
                    return true;
                }
            }
        }

        const float rsq = r * r;
        {
            const float4 *p = reinterpret_cast<const float4 *>(E + H.off_cuboids);
            for (uint32_t i = 0; i < H.n_cuboids; ++i)
            {
                const float4 a3 = p[4 * i + 3];
                if (!(a3.w < ext))
                {
                    break;
                }
                const float4 c = p[4 * i], a1 = p[4 * i + 1], a2 = p[4 * i + 2];
                const float xs = x - c.x, ys = y - c.y, zs = z - c.z;
                const float e1 = fmaxf(fabsf(a1.x * xs + a1.y * ys + a1.z * zs) - c.w, 0.F);
                const float e2 = fmaxf(fabsf(a2.x * xs + a2.y * ys + a2.z * zs) - a1.w, 0.F);
                const float e3 = fmaxf(fabsf(a3.x * xs + a3.y * ys + a3.z * zs) - a2.w, 0.F);
                if (sign_set((e1 * e1 + e2 * e2 + e3 * e3) - rsq))
                {
                    return true;
                }
            }
        }

        {
            const float4 *p = reinterpret_cast<const float4 *>(E + H.off_zcuboids);
            for (uint32_t i = 0; i < H.n_zcuboids; ++i)
            {
                const float4 c = p[3 * i];
                if (!(c.w < ext))
                {
                    break;
                }
                const float4 ax = p[3 * i + 1], h = p[3 * i + 2];
                const float xs = x - c.x, ys = y - c.y, zs = z - c.z;
                const float e1 = fmaxf(fabsf(ax.x * xs + ax.y * ys) - h.x, 0.F);
                const float e2 = fmaxf(fabsf(ax.z * xs + ax.w * ys) - h.y, 0.F);
                const float e3 = fmaxf(fabsf(zs) - h.z, 0.F);
                if (sign_set((e1 * e1 + e2 * e2 + e3 * e3) - rsq))
                {
                    return true;
                }
            }
        }

        return sphere_hits_heightfields(E, x, y, z, r);
    }

    // The pointclouds (CAPT and MVT) for one sphere, warp-cooperative: every lane of the converged group calls,
    // `query` false = the lane has nothing to ask but helps.  The nearest-point table answers first.
    static __device__ __noinline__ bool sphere_hits_clouds(const float *__restrict__ E, float x, float y, float z, float r_pc, bool query)
    {
        const EnvHeader &H = *reinterpret_cast<const EnvHeader *>(E);
        bool hit = false;
        float4 near = make_float4(0.F, 0.F, 0.F, 0.F);
        uint32_t near_cloud = kCloudTagNone;  // the cloud the nearest-point table's entry belongs to
#ifdef VMV_C4_STATS
        const long long t_probe = clock64();
#endif
        if (H.off_cloud_grid != 0 && query)
        {
            const CloudGridRec &g = *reinterpret_cast<const CloudGridRec *>(E + H.off_cloud_grid);
            float clearance;
            const bool in = cloud_probe(g, x, y, z, near, clearance);
            // no cloud point within reach: every pointcloud query would answer "no"
            query = !(clearance > r_pc + g.r_point_max);
            VMV_STAT(2, 1);
            VMV_STAT(3, query ? 1 : 0);
            near_cloud = (in && query) ? (__float_as_uint(near.w) >> 24) : kCloudTagNone;
        }
#ifdef VMV_C4_STATS
        __syncwarp(__activemask());
        if ((threadIdx.x & 31) == __ffs(__activemask()) - 1)
        {
            VMV_STAT(62, clock64() - t_probe);
        }
#endif
        for (uint32_t i = 0; i < H.n_capts; ++i)
        {
            const CaptRec &t = *reinterpret_cast<const CaptRec *>(E + H.off_capts + kCaptRec * i);
            // lanes without a query keep the warp's loop structure (they help scanning)
            if (capt_collides_warp(t, x, y, z, r_pc, query && !hit, near, near_cloud == i))
            {
                hit = true;
            }
        }
        for (uint32_t i = 0; i < H.n_mvts; ++i)
        {
            const MvtRec &t = *reinterpret_cast<const MvtRec *>(E + H.off_mvts + kMvtRec * i);
            if (query && !hit && mvt_collides(t, x, y, z, r_pc))
            {
                hit = true;
            }
        }
        return hit;
    }

    // One sphere against everything.  May be called by any subset of a warp's lanes; lanes that arrive together
    // cooperate on the pointcloud scans (`active` false = lane has no sphere but helps scanning).
    __device__ __forceinline__ bool
    sphere_hits_env(const float *__restrict__ E, float x, float y, float z, float r, float r_pc, bool active)
    {
        bool hit = false;
        if (active)
        {
            hit = sphere_hits_primitives(E, x, y, z, r);
        }
        const EnvHeader &H = *reinterpret_cast<const EnvHeader *>(E);
        if (H.n_capts + H.n_mvts == 0)
        {
            return hit;
        }
        if (active)
        {
            VMV_STAT(0, 1);
            VMV_STAT(1, hit ? 1 : 0);
        }
        return sphere_hits_clouds(E, x, y, z, r_pc, active && !hit) || hit;
    }

    // ------------------------------------------------------------------------------------------
    // TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
    // ------------------------------------------------------------------------------------------
    __device__ __forceinline__ uint32_t smem_u32(const void *p)
    {
        return static_cast<uint32_t>(__cvta_generic_to_shared(p));
    }

    __device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    __device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
            "l"(src_gmem),
            "r"(bytes),
            "r"(smem_u32(bar))
            : "memory");
    }

    __device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
    {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_LOOP:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE;\n"
            "bra WAIT_LOOP;\n"
            "DONE:\n"
            "}\n" ::"r"(smem_u32(bar)),
            "r"(phase)
            : "memory");
    }
}  // namespace vmv
