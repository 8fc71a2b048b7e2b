// Configurations of the reference's Halton sampler generated on the device (sm_100a), so that a planner
// front-end can have its samples validated without shipping 4*dof bytes per sample over PCIe: the host
// call takes (first, n) and returns one verdict bit per sample.
//
// Reference: vamp::rng::Halton<Robot>::next(), src/impl/vamp/random/halton.hh:76-107 -- a numerator /
// denominator recurrence per joint in f32 (bases 3, 5, 7, 11, ... halton.hh:17-33).  While both stay
// below 2^24 the recurrence is exact and sample i (1-based) is the radical inverse of i in the joint's
// base: numerator = digits of i reversed, denominator = base^(number of digits).  The kernel computes
// that directly (integer digits, division by a compile-time base = multiply-high), then n / d in f32 and
// Robot::scale_configuration = fma(u, range, lower) (robots/<robot>.hh; one FMA under the reference's
// -ffp-contract=fast).  The host refuses index ranges outside the exact regime (first epoch of 10^6
// samples; base^digits < 2^24).
//
// Thread = sample; a warp writes 32 * dof contiguous floats (28 B per Panda sample -- the only HBM
// traffic of this kernel).
#pragma once
#include <cstdint>

namespace vmv
{
    struct HaltonScale
    {
        float lower[16], range[16];
    };

    __host__ __device__ constexpr uint32_t halton_prime(int j)
    {
        constexpr uint32_t p[16] = {3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59};
        return p[j];
    }

    template <uint32_t B>
    __device__ __forceinline__ float halton_radical_inverse(uint32_t i)
    {
        uint32_t num = 0u, den = 1u;
        while (i != 0u)
        {
            const uint32_t qd = i / B;  // B is a compile-time constant: multiply-high + shift
            num = num * B + (i - qd * B);
            den *= B;
            i = qd;
        }
        return __fdiv_rn(__uint2float_rn(num), __uint2float_rn(den));
    }

    template <int DOF, int J = 0>
    __device__ __forceinline__ void halton_fill_row(uint32_t i, const HaltonScale &sc, float *row)
    {
        if constexpr (J < DOF)
        {
            row[J] = __fmaf_rn(halton_radical_inverse<halton_prime(J)>(i), sc.range[J], sc.lower[J]);
            halton_fill_row<DOF, J + 1>(i, sc, row);
        }
    }

    // (staging the rows in shared memory for fully coalesced stores was measured: 30.8 us instead of
    // 24.8 us per 10^6 Panda samples -- the kernel is bound by its integer digit loops, not by the stores)
    template <int DOF>
    __global__ void __launch_bounds__(256) k_halton_fill(const __grid_constant__ HaltonScale sc, uint64_t first, size_t n, float *__restrict__ q)
    {
        const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
        for (size_t s = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; s < n; s += stride)
        {
            float row[DOF];
            halton_fill_row<DOF>(static_cast<uint32_t>(first + s) + 1u, sc, row);
#pragma unroll
            for (int j = 0; j < DOF; ++j)
            {
                q[s * DOF + j] = row[j];
            }
        }
    }
}  // namespace vmv
