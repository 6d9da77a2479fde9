// Microbenchmark: how fast can CTAs stream contiguous regions of HBM into shared memory on B200?
//   mode 0: cp.async.bulk (1-D TMA) ring, one issuing thread, mbarrier completion
//   mode 1: LDG.128 by all threads, accumulate (no smem staging)
//   mode 2: cp.async (LDGSTS) 16 B per thread ring
// Each CTA streams `region` bytes starting at cta*region (or interleaved order), chunk bytes / stages configurable.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ bool mb_try(uint32_t b, uint32_t ph) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mb_wait(uint32_t b, uint32_t ph) { while (!mb_try(b, ph)) {} }
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int STAGES>
__global__ void __launch_bounds__(256) k_bulk(const double* src, size_t region_d, int chunk_d, int nsplit, double* out) {
    extern __shared__ __align__(128) unsigned char raw[];
    double* buf = reinterpret_cast<double*>(raw);
    __shared__ unsigned long long full[STAGES];
    const int tid = threadIdx.x;
    if (tid == 0) { for (int s = 0; s < STAGES; ++s) mb_init(s32(&full[s]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const double* base = src + (size_t)blockIdx.x * region_d;
    const int nch = (int)(region_d / chunk_d);
    const int piece = chunk_d / nsplit;
    auto issue = [&](int c) {
        const int s = c % STAGES;
        mb_expect(s32(&full[s]), chunk_d * 8);
        for (int q = 0; q < nsplit; ++q) bulk(s32(buf + (size_t)s * chunk_d + q * piece), base + (size_t)c * chunk_d + q * piece, piece * 8, s32(&full[s]));
    };
    if (tid == 0) for (int c = 0; c < STAGES && c < nch; ++c) issue(c);
    double acc = 0.0;
    for (int c = 0; c < nch; ++c) {
        const int s = c % STAGES;
        mb_wait(s32(&full[s]), (c / STAGES) & 1);
        const double* b = buf + (size_t)s * chunk_d;
        for (int i = tid; i < chunk_d; i += 256) acc += b[i];
        __syncthreads();
        if (tid == 0 && c + STAGES < nch) issue(c + STAGES);
    }
    if (acc == 1.2345) out[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(256) k_ldg(const double* src, size_t region_d, int unroll, double* out) {
    const double2* base = reinterpret_cast<const double2*>(src + (size_t)blockIdx.x * region_d);
    const size_t n2 = region_d / 2;
    double acc = 0.0;
    size_t i = threadIdx.x;
    for (; i + 7 * 256 < n2; i += 8 * 256) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = base[i + u * 256];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y;
    }
    for (; i < n2; i += 256) { double2 v = base[i]; acc += v.x + v.y; }
    if (acc == 1.2345) out[blockIdx.x] = acc;
}

template <int STAGES>
__global__ void __launch_bounds__(256) k_cpasync(const double* src, size_t region_d, int chunk_d, double* out) {
    extern __shared__ __align__(128) unsigned char raw[];
    double* buf = reinterpret_cast<double*>(raw);
    const int tid = threadIdx.x;
    const double* base = src + (size_t)blockIdx.x * region_d;
    const int nch = (int)(region_d / chunk_d);
    auto issue = [&](int c) {
        const int s = c % STAGES;
        for (int i = tid * 2; i < chunk_d; i += 512)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(buf + (size_t)s * chunk_d + i)), "l"(base + (size_t)c * chunk_d + i) : "memory");
    };
    for (int c = 0; c < STAGES - 1; ++c) { if (c < nch) issue(c); asm volatile("cp.async.commit_group;" ::: "memory"); }
    double acc = 0.0;
    for (int c = 0; c < nch; ++c) {
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
        __syncthreads();
        if (c + STAGES - 1 < nch) issue(c + STAGES - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const double* b = buf + (size_t)(c % STAGES) * chunk_d;
        for (int i = tid; i < chunk_d; i += 256) acc += b[i];
    }
    if (acc == 1.2345) out[blockIdx.x] = acc;
}

int main() {
    const size_t total = (size_t)1 << 30;            // 1 GiB of doubles region pool
    double* src; double* out;
    CK(cudaMalloc(&src, total));
    CK(cudaMalloc(&out, 1 << 20));
    CK(cudaMemset(src, 0, total));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    auto report = [&](const char* name, size_t bytes, float ms) { printf("%-44s %8.1f MB  %8.1f us  %7.2f TB/s\n", name, bytes / 1e6, ms * 1e3, bytes / (ms * 1e-3) / 1e12); };
    int ctas_list[] = {296, 592, 888, 1184, 2368};
    size_t region_list[] = {256 << 10, 512 << 10};
    for (size_t region : region_list)
        for (int ctas : ctas_list) {
            if ((size_t)ctas * region > total) continue;
            const size_t region_d = region / 8;
            char name[128];
            float ms;
#define RUN(label, launch) do { for (int w = 0; w < 2; ++w) { launch; } cudaEventRecord(a); for (int r = 0; r < 5; ++r) { launch; } cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms, a, b); snprintf(name, sizeof(name), "%s ctas=%d region=%zuK", label, ctas, region >> 10); report(name, (size_t)ctas * region, ms / 5); } while (0)
            CK(cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
            CK(cudaFuncSetAttribute(k_bulk<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
            CK(cudaFuncSetAttribute(k_cpasync<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
            RUN("bulk 4st x 16K x1", (k_bulk<4><<<ctas, 256, 4 * 16384>>>(src, region_d, 2048, 1, out)));
            RUN("bulk 4st x 16K x4pieces", (k_bulk<4><<<ctas, 256, 4 * 16384>>>(src, region_d, 2048, 4, out)));
            RUN("bulk 4st x 16K x16pieces", (k_bulk<4><<<ctas, 256, 4 * 16384>>>(src, region_d, 2048, 16, out)));
            RUN("bulk 8st x 8K", (k_bulk<8><<<ctas, 256, 8 * 8192>>>(src, region_d, 1024, 1, out)));
            RUN("bulk 4st x 32K (1 cta/sm)", (k_bulk<4><<<ctas, 256, 4 * 32768>>>(src, region_d, 4096, 1, out)));
            RUN("bulk 8st x 4K", (k_bulk<8><<<ctas, 256, 8 * 4096>>>(src, region_d, 512, 1, out)));
            RUN("ldg.128 unroll8", (k_ldg<<<ctas, 256>>>(src, region_d, 8, out)));
            RUN("cp.async 4st x 16K", (k_cpasync<4><<<ctas, 256, 4 * 16384>>>(src, region_d, 2048, out)));
        }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
