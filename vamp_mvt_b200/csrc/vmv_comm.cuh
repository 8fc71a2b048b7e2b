// Kernels of the multi-GPU verdict gather that do not belong to a robot (included by vmv_host.cu only).
#pragma once
#include "vmv_kernels_v4.cuh"

namespace vmv
{
    // consumer side: one thread per rank spins until that rank's launch `seq` (or a later one) has landed
    __global__ void k_comm_wait(const uint32_t *flags, int world, uint32_t seq)
    {
        if (static_cast<int>(threadIdx.x) < world)
        {
            uint32_t v;
            do
            {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
            } while (static_cast<int32_t>(v - seq) < 0);
        }
    }

    // gather for the kernel generations without fused stores: local words -> every window, then the flags
    __global__ void __launch_bounds__(256) k_comm_push(const uint32_t *__restrict__ local, size_t n_words, const __grid_constant__ GatherDev g)
    {
        for (int p = 0; p < g.world; ++p)
        {
            uint32_t *dst = g.peer_bits[p];
            for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_words; i += static_cast<size_t>(gridDim.x) * blockDim.x)
            {
                dst[i] = local[i];
            }
        }
        gather_finish(g);
    }
}  // namespace vmv
