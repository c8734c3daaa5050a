// Microbenchmark: FP32 FMA throughput of one B200, scalar FFMA vs packed FFMA2 (sm_100a).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32_peak fp32_peak.cu && ./fp32_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>  // 0: scalar FFMA (3 register operands), 1: FFMA2, 2: scalar FFMA with an immediate multiplier
__global__ void __launch_bounds__(256) k(float *out, float a, float b, int iters)
{
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], 0.999f, b);
        } else {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float2 v = __ffma2_rn(make_float2(x[i], x[i + 1]), make_float2(a, a), make_float2(b, b));
                x[i] = v.x;
                x[i + 1] = v.y;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 8, iters = 20000;
    float *out;
    cudaMalloc(&out, (size_t)grid * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, 0.999f, 1e-3f, iters);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, 0.999f, 1e-3f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)grid * 256 * 16 * iters;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", name, ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / 1.9e9);
    cudaFree(out);
}

int main()
{
    run<0>("scalar FFMA (reg, reg, reg)");
    run<2>("scalar FFMA (reg, imm, reg)");
    run<1>("packed FFMA2");
    return 0;
}
