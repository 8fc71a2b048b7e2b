// Voxel table of candidate object masks and the rounded-box object records (sm_100a): the broad
// phase of the grid-culled kernels in vmv_kernels_v4.cuh.
//
// The kernels of vmv_kernels_v2.cuh test every link's bounding sphere against every object in
// reach: ~130 exact tests per Panda configuration on the bench scene, of which fewer than one hits.
// Here the environment is rasterised once, when it is first used with a robot, into a voxel table in
// HBM (L2-resident, a few MB): per voxel and per radius class a bit mask of the objects that come
// within (class radius + voxel half diagonal) of the voxel centre.  One load per link then replaces
// the object sweep; the mask is a conservative superset, and only its members (1-3 per
// configuration) go through the exact margin test.  Verdicts are those of the exact tests, i.e.
// unchanged (reference collision/validity.hh:46-158 is likewise "cull, then test").
//
// All primitives share one record, a rounded box {centre, 3 axes, 3 half extents, rounding
// radius}: margin = sum_i max(|a_i.(p-c)| - h_i, 0)^2 - (r + rho)^2.  A cuboid has rho = 0
// (reference collision/sphere_cuboid.hh:9-26), a sphere h = 0 (sphere_sphere.hh:10-23), a capsule is
// its segment (h = (len/2, 0, 0)) rounded by its radius -- the same distance as the clamped
// projection of sphere_capsule.hh:9-22.  One record type = no divergence on the object kind.
#pragma once
#include "vmv_kernels_v2.cuh"
#include "vmv_pairtab.cuh"

namespace vmv
{
    static constexpr int kGridClasses = 4;
    static constexpr int kGridMaxLinks = 64;
    static constexpr int kObjRec = 16;  // floats: {cx cy cz rho}{a1 h1}{a2 h2}{a3 h3}

    struct GridDev
    {
        const void *masks;  // [nz][ny][nx][kGridClasses] of uint32_t (<= 32 objects) or unsigned long long
        float x0, y0, z0, inv_h;
        int nx, ny, nz;
        unsigned long long all_mask;
        unsigned char link_class[kGridMaxLinks];
    };

    // Attachment (reference fkcc_attach): spheres in the attachment's frame, the frame relative to the end-effector body
    struct AttachDev
    {
        const float4 *spheres;  // n x {x y z r}
        uint32_t n;
        float tf[12];           // ee_tf(robot) * attachment offset, row-major 3x4
    };

    struct GridEnv
    {
        const float4 *objs;  // n_objects rounded-box records
        uint32_t n_objects;
        uint32_t max_fine;   // largest number of fine spheres of any link of the robot
        uint32_t q2_rounds;  // fine-item queue capacity in B1 rounds (shared-memory budget of the kernel)
        GridDev grid;
        PairTabDev tab;      // two-joint verdict tables of the robot (n_groups = 0: none)
        // any-environment instantiation (heightfields / pointclouds next to at most 30 primitives): the packed
        // environment blob (vmv_device.cuh: EnvHeader + heightfield, CAPT, MVT and clearance-grid records), staged
        // into shared memory next to the rounded-box records; blob_bytes = 0 otherwise
        const float *blob;
        uint32_t blob_bytes;
        AttachDev att;       // n = 0: nothing attached
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
    template <typename MaskT>
    __global__ void __launch_bounds__(128) k_build_grid_t(
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
        MaskT *__restrict__ out)
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
        MaskT m0 = 0, m1 = 0, m2 = 0, m3 = 0;
        for (uint32_t k = 0; k < n_objects; ++k)
        {
            const float4 c = __ldg(objs + 4 * k), a1 = __ldg(objs + 4 * k + 1), a2 = __ldg(objs + 4 * k + 2), a3 = __ldg(objs + 4 * k + 3);
            const float xs = x - c.x, ys = y - c.y, zs = z - c.z;
            const float e1 = fmaxf(fabsf(a1.x * xs + a1.y * ys + a1.z * zs) - a1.w, 0.F);
            const float e2 = fmaxf(fabsf(a2.x * xs + a2.y * ys + a2.z * zs) - a2.w, 0.F);
            const float e3 = fmaxf(fabsf(a3.x * xs + a3.y * ys + a3.z * zs) - a3.w, 0.F);
            const float d = sqrtf(e1 * e1 + e2 * e2 + e3 * e3) - c.w - slack;
            const MaskT bit = static_cast<MaskT>(1) << k;
            // !(d > r) also keeps an object whose record is not finite
            m0 |= !(d > r0) ? bit : 0;
            m1 |= !(d > r1) ? bit : 0;
            m2 |= !(d > r2) ? bit : 0;
            m3 |= !(d > r3) ? bit : 0;
        }
        MaskT *o = out + v * kGridClasses;
        o[0] = m0, o[1] = m1, o[2] = m2, o[3] = m3;
    }

    template <typename MaskT>
    __device__ __forceinline__ MaskT grid_lookup_t(const GridDev &G, float x, float y, float z, int cls)
    {
        const int ix = __float2int_rd((x - G.x0) * G.inv_h);
        const int iy = __float2int_rd((y - G.y0) * G.inv_h);
        const int iz = __float2int_rd((z - G.z0) * G.inv_h);
        // outside the table = farther than the largest class radius from every object
        const bool in = (static_cast<unsigned>(ix) < static_cast<unsigned>(G.nx)) & (static_cast<unsigned>(iy) < static_cast<unsigned>(G.ny)) &
                        (static_cast<unsigned>(iz) < static_cast<unsigned>(G.nz));
        if (!in)
        {
            // a centre that is not finite also lands here (its index saturates): give it every object
            return (fabsf(x) < 1e30F && fabsf(y) < 1e30F && fabsf(z) < 1e30F) ? static_cast<MaskT>(0) : static_cast<MaskT>(G.all_mask);
        }
        const uint32_t idx = ((static_cast<uint32_t>(iz) * G.ny + iy) * G.nx + ix) * kGridClasses + cls;
        return __ldg(reinterpret_cast<const MaskT *>(G.masks) + idx);
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
}  // namespace vmv
