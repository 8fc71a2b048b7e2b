// Micro-benchmark (dev tool): issue rate of FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, alone and
// mixed with an ALU-pipe instruction.  Prints warp-instructions per cycle per SM and flop/cycle/SM.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float s)
{
    float a[8];
    unsigned long long p[8];
    unsigned int m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        a[i] = threadIdx.x * 0.001f + i;
        p[i] = (static_cast<unsigned long long>(__float_as_uint(a[i])) << 32) | __float_as_uint(a[i] + 0.5f);
        m[i] = threadIdx.x + i;
    }
    const unsigned long long ps = (static_cast<unsigned long long>(__float_as_uint(s)) << 32) | __float_as_uint(s);
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            if (MODE == 0)
            {
                a[i] = fmaf(a[i], s, 0.25f);
            }
            else if (MODE == 1)
            {
                p[i] = ffma2(p[i], ps, ps);
            }
            else if (MODE == 2)
            {
                a[i] = fmaf(a[i], s, 0.25f);
                m[i] = __vimax_s32_relu(m[i], it) ^ m[(i + 1) & 7];
            }
            else
            {
                p[i] = ffma2(p[i], ps, ps);
                m[i] = __vimax_s32_relu(m[i], it) ^ m[(i + 1) & 7];
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        r += a[i] + __uint_as_float(static_cast<unsigned>(p[i])) + __uint_as_float(static_cast<unsigned>(p[i] >> 32)) + m[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char *name, int fp_per_iter, int lanes_per_fp, int other_per_iter)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out;
    cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    const int iters = 20000;
    k<MODE><<<sms * 8, 256>>>(out, iters, 0.999f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms * 8, 256>>>(out, iters, 0.999f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * khz * 1e3;
    const double warps = static_cast<double>(sms) * 8 * 8;
    const double fp_instr = warps * iters * fp_per_iter, other = warps * iters * other_per_iter;
    printf("%-22s %.3f ms  fp warp-instr/cycle/SM %.2f  flop/cycle/SM %.0f  total warp-instr/cycle/SM %.2f\n", name, ms,
           fp_instr / cycles / sms, fp_instr * 32 * lanes_per_fp * 2 / cycles / sms, (fp_instr + other) / cycles / sms);
    cudaFree(out);
}

int main()
{
    run<0>("FFMA", 8, 1, 0);
    run<1>("FFMA2", 8, 2, 0);
    run<2>("FFMA + 2 ALU", 8, 1, 16);
    run<3>("FFMA2 + 2 ALU", 8, 2, 16);
    return 0;
}
