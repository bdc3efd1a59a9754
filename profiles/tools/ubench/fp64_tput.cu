// FP64 throughput on B200: independent DFMA streams and mma.sync m8n8k4 f64 (DMMA), per SM and whole GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_tput fp64_tput.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048
template <int ACC>
__global__ void k_dfma(double* out, long long* cyc, double seed)
{
    double a[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) a[i] = seed + i + threadIdx.x;
    const double m = 1.0000001, c = seed * 0.25;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) a[i] = fma(a[i], m, c);
    }
    __syncthreads();
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_dmma(double* out, long long* cyc, double seed)
{
    double d0[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) d0[i][0] = d0[i][1] = seed + i;
    const double a = 1.0000001 + threadIdx.x * 1e-9, b = 0.9999999;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(d0[i][0]), "+d"(d0[i][1]) : "d"(a), "d"(b));
    }
    __syncthreads();
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += d0[i][0] + d0[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main()
{
    double* out;
    long long *cyc, h;
    cudaMalloc(&out, 148 * 8 * 1024 * 8);
    cudaMalloc(&cyc, 64);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int threads : {32, 128, 256, 512, 1024}) {
        for (int grid : {1, sms}) {
            k_dfma<16><<<grid, threads>>>(out, cyc, 1.25);
            cudaEventRecord(e0);
            k_dfma<16><<<grid, threads>>>(out, cyc, 1.25);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            const double fma_per_sm_clk = (double)threads * 16 * ITER / (double)h;
            printf("DFMA threads=%4d grid=%3d: %.2f FMA/clk/SM (%.2f clk per warp-instr per SMSP)  %.2f TFLOP/s (events)\n", threads, grid,
                   fma_per_sm_clk, 32.0 / (fma_per_sm_clk / (threads >= 128 ? 4 : 1)), 2.0 * threads * 16 * ITER * grid / (ms * 1e-3) * 1e-12);
        }
    }
    for (int threads : {32, 128, 256, 512}) {
        for (int grid : {1, sms}) {
            k_dmma<<<grid, threads>>>(out, cyc, 1.25);
            cudaEventRecord(e0);
            k_dmma<<<grid, threads>>>(out, cyc, 1.25);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            const double fma_per_sm_clk = (double)(threads / 32) * 4 * ITER * 256 / (double)h; // m8n8k4 = 256 FMA
            printf("DMMA threads=%4d grid=%3d: %.2f FMA/clk/SM  %.2f TFLOP/s (events)\n", threads, grid, fma_per_sm_clk,
                   2.0 * (threads / 32) * 4 * ITER * 256 * grid / (ms * 1e-3) * 1e-12);
        }
    }
    return 0;
}
