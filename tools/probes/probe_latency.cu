// Probe (measurement tool): latencies that bound the per-op cost of the shared-memory tile kernels on B200:
// dependent / independent FP64 FMA, FP64 pipe throughput per SM, __syncthreads, LDS.128 round trip.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/probe_latency tools/probes/probe_latency.cu
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void k_dfma_chain(double *out, int iters, long long *cyc) {
    double x = threadIdx.x * 1e-3, a = 1.0000001, b = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = fma(x, a, b);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dfma_ilp8(double *out, int iters, long long *cyc) {
    double x[8];
    for (int k = 0; k < 8; ++k) x[k] = threadIdx.x * 1e-3 + k;
    const double a = 1.0000001, b = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = fma(x[k], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int k = 0; k < 8; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_ffma_chain(float *out, int iters, long long *cyc) {
    float x = threadIdx.x * 1e-3f, a = 1.0000001f, b = 1e-9f;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = fmaf(x, a, b);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_sync(int iters, long long *cyc) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_lds_chain(int iters, long long *cyc, int *sink) {
    __shared__ int4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_int4((i * 7 + 3) & 1023, 0, 0, 0);
    __syncthreads();
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) idx = buf[idx].x;
    long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + threadIdx.x] = idx;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// one "op" of the tile kernel in isolation: LDS.128 x2 -> 8 DFMA/DMUL -> STS.128 x2 -> barrier, `iters` times
__global__ void k_op_like(int iters, long long *cyc, double *out) {
    __shared__ double2 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_double2(i * 1e-3, 1.0);
    __syncthreads();
    const double c = 0.8, s = 0.6;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int k = threadIdx.x; k < 512; k += blockDim.x) {
            const int i = (k * 4 + 1 + it) & 2047, j = i ^ 3;
            double2 a = buf[i], b = buf[j];
            double2 ra = make_double2(c * a.x + s * b.x, c * a.y + s * b.y);
            double2 rb = make_double2(-s * a.x + c * b.x, -s * a.y + c * b.y);
            buf[i] = ra;
            buf[j] = rb;
        }
        __syncthreads();
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = buf[threadIdx.x].x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    long long *d_cyc, h;
    double *d_out;
    float *f_out;
    int *i_out;
    cudaMalloc(&d_cyc, 8);
    cudaMalloc(&d_out, 8 * 148 * 1024);
    cudaMalloc(&f_out, 4 * 148 * 1024);
    cudaMalloc(&i_out, 4 * 148 * 1024);
    const int iters = 4096;
#define RUN(name, call, per)                                                    \
    do {                                                                        \
        call;                                                                   \
        call;                                                                   \
        cudaDeviceSynchronize();                                                \
        cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost);                       \
        printf("%-56s %10.2f cycles per %s\n", name, (double)h / iters, per);   \
    } while (0)
    RUN("dependent DFMA chain, 1 warp", (k_dfma_chain<<<1, 32>>>(d_out, iters, d_cyc)), "DFMA");
    RUN("dependent FFMA chain, 1 warp", (k_ffma_chain<<<1, 32>>>(f_out, iters, d_cyc)), "FFMA");
    RUN("8 independent DFMA per iteration, 1 warp", (k_dfma_ilp8<<<1, 32>>>(d_out, iters, d_cyc)), "8 DFMA");
    RUN("8 independent DFMA per iteration, 4 warps (1/SMSP)", (k_dfma_ilp8<<<1, 128>>>(d_out, iters, d_cyc)), "8 DFMA");
    RUN("8 independent DFMA per iteration, 16 warps (4/SMSP)", (k_dfma_ilp8<<<1, 512>>>(d_out, iters, d_cyc)), "8 DFMA");
    RUN("8 independent DFMA per iteration, 32 warps (8/SMSP)", (k_dfma_ilp8<<<1, 1024>>>(d_out, iters, d_cyc)), "8 DFMA");
    RUN("__syncthreads, 128 threads", (k_sync<<<1, 128>>>(iters, d_cyc)), "barrier");
    RUN("__syncthreads, 512 threads", (k_sync<<<1, 512>>>(iters, d_cyc)), "barrier");
    RUN("dependent LDS.128 chain, 1 warp", (k_lds_chain<<<1, 32>>>(iters, d_cyc, i_out)), "LDS");
    RUN("tile-op-like (512 pairs, real 2x2) + barrier, 128 threads", (k_op_like<<<1, 128>>>(iters, d_cyc, d_out)), "op");
    RUN("tile-op-like (512 pairs, real 2x2) + barrier, 512 threads", (k_op_like<<<1, 512>>>(iters, d_cyc, d_out)), "op");
    RUN("tile-op-like, 512 threads, 128 CTAs", (k_op_like<<<128, 512>>>(iters, d_cyc, d_out)), "op");
    // FP64 throughput of the whole chip: 148 x 8 CTAs x 256 threads
    {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        k_dfma_ilp8<<<148 * 8, 256>>>(d_out, iters, d_cyc);
        cudaEventRecord(e0);
        k_dfma_ilp8<<<148 * 8, 256>>>(d_out, iters, d_cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * iters * 148.0 * 8 * 256;
        printf("chip FP64 FMA throughput: %.2f TFLOP/s\n", flops / ms / 1e9);
    }
    return 0;
}
