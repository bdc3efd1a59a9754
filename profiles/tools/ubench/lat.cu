// Micro-benchmarks of the latencies the small-N factor/solve kernels are bound by (single warp, clock64).
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

#define REP 256
__global__ void k_lat(double* out, long long* cyc, double seed)
{
    __shared__ double s[64];
    if (threadIdx.x < 64) s[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    double x = seed, y = seed * 0.5 + 1.0;
    long long t0, t1;
    // dependent DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; ++i) x = fma(x, 1.0000001, y);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // dependent FP64 division chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < REP; ++i) x = 1.0 / (x + 2.0);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // dependent 64-bit shuffle + DSETP/select (arg-max style)
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < REP; ++i) {
        const double o = __shfl_xor_sync(0xffffffffu, x, 1 + (i & 15));
        x = fabs(o) > fabs(x) ? o : x + 1e-30;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // REDUX chain
    unsigned u = (unsigned)__double2int_rn(x) + threadIdx.x;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < REP; ++i) u = __reduce_max_sync(0xffffffffu, u + threadIdx.x) + 1;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // dependent LDS chain (pointer chasing through shared memory)
    int idx = threadIdx.x & 31;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < REP; ++i) idx = (int)s[idx & 63] & 63;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // __syncthreads chain
    t0 = clock64();
    for (int i = 0; i < REP; ++i) __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // approx reciprocal + 2 Newton steps
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < REP; ++i) {
        const double d = x + 2.0;
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
        double e = fma(-d, r, 1.0);
        r = fma(r, e, r);
        e = fma(-d, r, 1.0);
        x = fma(r, e, r);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    out[threadIdx.x] = x + u + idx;
}

__global__ void k_cluster(long long* cyc)
{
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ int s_x[16];
    long long t0 = clock64();
    for (int i = 0; i < REP; ++i) cluster.sync();
    long long t1 = clock64();
    if (threadIdx.x == 0 && cluster.block_rank() == 0) cyc[0] = t1 - t0;
    // remote DSMEM read latency (dependent)
    if (threadIdx.x < 16) s_x[threadIdx.x] = (threadIdx.x + 1) & 15;
    cluster.sync();
    int idx = 0;
    const unsigned peer = (cluster.block_rank() + 1) % cluster.num_blocks();
    t0 = clock64();
    for (int i = 0; i < REP; ++i) idx = *cluster.map_shared_rank(&s_x[idx & 15], peer);
    t1 = clock64();
    if (threadIdx.x == 0 && cluster.block_rank() == 0) cyc[1] = (t1 - t0) + (idx & 1);
    cluster.sync();
}

int main()
{
    double* out;
    long long *cyc, h[8];
    cudaMalloc(&out, 1024 * 8);
    cudaMalloc(&cyc, 64);
    for (int threads : {32, 256}) {
        k_lat<<<1, threads>>>(out, cyc, 1.25);
        cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost);
        printf("threads=%d per-op cycles: DFMA %.1f  DDIV %.1f  shfl64+dsetp+sel %.1f  REDUX %.1f  LDS %.1f  syncthreads %.1f  rcp.approx+2NR %.1f\n",
               threads, h[0] / (double)REP, h[1] / (double)REP, h[2] / (double)REP, h[3] / (double)REP, h[4] / (double)REP,
               h[5] / (double)REP, h[6] / (double)REP);
    }
    cudaFuncSetAttribute(k_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs);
        cfg.blockDim = dim3(256);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_cluster, cyc);
        cudaError_t e2 = cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
        printf("cluster %2d: launch %s / %s  cluster.sync %.1f cycles  DSMEM dependent read %.1f cycles\n", cs,
               cudaGetErrorString(e), cudaGetErrorString(e2), h[0] / (double)REP, h[1] / (double)REP);
    }
    return 0;
}
