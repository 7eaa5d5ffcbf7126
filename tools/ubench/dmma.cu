// Micro-benchmark: FP64 throughput of mma.sync.m8n8k4.f64 (DMMA) vs plain DFMA on one GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k_dmma(double *out, int iters, double a0, double b0)
{
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double *out, int iters, double a0, double b0)
{
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s SMs %d clock %d kHz\n", p.name, sms, p.clockRate);
    double *out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps = 1; warps <= 16; warps *= 2) {
        for (int ctas = 1; ctas <= 2; ctas++) {
            float ms;
            k_dmma<8><<<sms * ctas, 32 * warps>>>(out, 100, 1.0, 1.0);
            cudaEventRecord(e0);
            k_dmma<8><<<sms * ctas, 32 * warps>>>(out, iters, 1.0, 1.0);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            double fma = (double)sms * ctas * warps * iters * 8.0 * 256.0;
            printf("DMMA m8n8k4 warps/CTA %2d CTAs/SM %d : %.3f ms  %.2f TFLOP/s (%.1f FMA/clk/SM @1.965GHz)\n", warps, ctas, ms,
                   2 * fma / ms * 1e-9, fma / (ms * 1e-3) / sms / 1.965e9);
            k_dfma<8><<<sms * ctas, 32 * warps>>>(out, 100, 1.0, 1.0);
            cudaEventRecord(e0);
            k_dfma<8><<<sms * ctas, 32 * warps>>>(out, iters, 1.0, 1.0);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            fma = (double)sms * ctas * warps * iters * 8.0 * 32.0;
            printf("DFMA          warps/CTA %2d CTAs/SM %d : %.3f ms  %.2f TFLOP/s (%.1f FMA/clk/SM)\n", warps, ctas, ms,
                   2 * fma / ms * 1e-9, fma / (ms * 1e-3) / sms / 1.965e9);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
