// Microbenchmark (GPU box): FP64 throughput of register-resident DFMA chains vs the FP64 tensor-core
// instructions (mma.sync m8n8k4 / m16n8k8 / m16n8k16 f64) on B200.  Decides whether the weight-matrix
// product of the EM (vp_localisation.py:515-524, an (M x N)(N x N) float64 product) belongs on DMMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_bench fp64_bench.cu && ./fp64_bench
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double seed) {
    double a[16], b = seed + threadIdx.x * 1e-9, c = 1.0 - 1e-9 * threadIdx.x;
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = seed * (k + 1);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = fma(a[k], c, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>   // 0: m8n8k4, 1: m16n8k8, 2: m16n8k16
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double seed) {
    double a[8], b[4], c0[4], c1[4], c2[4], c3[4];
    for (int k = 0; k < 8; ++k) a[k] = seed + 1e-9 * (threadIdx.x + k);
    for (int k = 0; k < 4; ++k) { b[k] = 1.0 - 1e-9 * (threadIdx.x + k); c0[k] = c1[k] = c2[k] = c3[k] = seed * k; }
    for (int i = 0; i < iters; ++i) {
        if (SHAPE == 0) {
#define M884(C) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(C[0]), "+d"(C[1]) : "d"(a[0]), "d"(b[0]))
            M884(c0); M884(c1); M884(c2); M884(c3);
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[2]), "+d"(c0[3]) : "d"(a[1]), "d"(b[1]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[2]), "+d"(c1[3]) : "d"(a[1]), "d"(b[1]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[2]), "+d"(c2[3]) : "d"(a[1]), "d"(b[1]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[2]), "+d"(c3[3]) : "d"(a[1]), "d"(b[1]));
        } else if (SHAPE == 1) {
#define M1688(C) asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
        : "+d"(C[0]), "+d"(C[1]), "+d"(C[2]), "+d"(C[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]))
            M1688(c0); M1688(c1); M1688(c2); M1688(c3);
        } else {
#define M16816(C) asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};" \
        : "+d"(C[0]), "+d"(C[1]), "+d"(C[2]), "+d"(C[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]))
            M16816(c0); M16816(c1); M16816(c2); M16816(c3);
        }
    }
    double s = 0;
    for (int k = 0; k < 4; ++k) s += c0[k] + c1[k] + c2[k] + c3[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, iters = 20000;
    double* out;
    cudaMalloc(&out, sizeof(double) * 256 * sms * 8);
    for (int cps = 1; cps <= 4; cps *= 2) {
        const int grid = sms * cps;
        double ms = time_ms([&] { dfma_kernel<<<grid, 256>>>(out, iters, 0.5); });
        printf("DFMA   %d CTA/SM x 256 thr: %.2f TFLOP/s\n", cps, 2.0 * 16 * iters * 256.0 * grid / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { dmma_kernel<0><<<grid, 256>>>(out, iters, 0.5); });
        printf("m8n8k4   %d CTA/SM: %.2f TFLOP/s\n", cps, 2.0 * 8 * 256 * iters * 8.0 * grid / (ms * 1e-3) / 1e12);     // 8 warps x 8 mma x 256 FMA
        ms = time_ms([&] { dmma_kernel<1><<<grid, 256>>>(out, iters, 0.5); });
        printf("m16n8k8  %d CTA/SM: %.2f TFLOP/s\n", cps, 2.0 * 4 * 1024 * iters * 8.0 * grid / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { dmma_kernel<2><<<grid, 256>>>(out, iters, 0.5); });
        printf("m16n8k16 %d CTA/SM: %.2f TFLOP/s\n", cps, 2.0 * 4 * 2048 * iters * 8.0 * grid / (ms * 1e-3) / 1e12);
    }
    printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
