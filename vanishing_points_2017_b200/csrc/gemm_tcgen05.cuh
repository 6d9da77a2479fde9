// bf16 x bf16 -> fp32 GEMM for sm_100a: TMA-fed, tcgen05.mma with the
// accumulator in tensor memory, warp-specialised (TMA producer / MMA issuer /
// 4 epilogue warps), used for every dense contraction of cnn/deploy.prototxt.
//
//   D[m, n] = act( sum_k A[row(m, k), k] * B[n, k] + bias[n] )
//
// A and B are K-major (K contiguous).  A convolution is run as an *implicit*
// GEMM without materialising im2col: the activation tensor is stored NHWC with
// its zero padding in memory, the output rows m enumerate the PADDED grid
// (n, y, x) of the input, and kernel tap (kh, kw) of output row m is input row
// m + kh*Wp + kw -- a plain row shift of the 2-D TMA box.  Output rows that fall
// into the padding (x >= W or y >= H) are computed and dropped in the epilogue.
//
// Shared-memory operand tiles use the 128-byte swizzle written by TMA
// (CU_TENSOR_MAP_SWIZZLE_128B) and read by the UMMA descriptors below
// (layout_type 2, SBO = 1024 B); the K-slice of one tcgen05.mma (16 bf16 = 32 B)
// is selected by advancing the descriptor start address.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vpk {

static constexpr int kGemmThreads = 320;       // warp0 TMA, warp1 MMA/TMEM, warps 2-9 epilogue (two per TMEM lane quarter,
                                               // each taking half of the tile's columns)
static constexpr int kBM = 128;                // UMMA_M (cta_group::1)
static constexpr int kBK = 64;                 // bf16 per 128-byte swizzle row
static constexpr int kABytes = kBM * kBK * 2;  // 16 KiB per stage

struct GemmParams {
    // ---- operand addressing
    int m_total;       // rows of the (padded) output grid
    int k_blocks;      // number of 64-wide K blocks
    int cblocks;       // K blocks per kernel tap (channels/64)
    int taps_x;        // taps per kernel row (1 for plain GEMM)
    int row_pitch;     // Wp: A row shift per kernel row
    int a_col_group;   // A column offset per group
    int b_row_group;   // B row offset per group
    // ---- epilogue
    int hp_wp, wp, h_valid, w_valid;       // m -> (n, y, x); valid if y < h_valid and x < w_valid
    int out_hp_wp, out_wp, out_pad;        // output row = n*out_hp_wp + (y+out_pad)*out_wp + (x+out_pad)
    int ldc;                               // output row pitch (elements)
    int c_col_group;                       // output column offset per group
    int n_valid;                           // output columns per group
    int relu, out_f32;
    int lrn;                               // fused LRN (ACROSS_CHANNELS, size 5, alpha 1e-4, beta 0.75, k 1; cnn/deploy.prototxt:34-44)
                                           // after bias + ReLU; needs all n_valid channels in ONE tile (n_valid <= bn, groups == 1)
    int fold;                              // > 1: ONE CTA computes `fold` consecutive groups of its rows one after the other into
                                           // adjacent TMEM column ranges (grid.z = groups / fold), so that the fused LRN sees the
                                           // channels of all groups (conv2: 2 x 128; needs lrn, n_valid == bn, fold * bn <= 256)
    int bn;                                // N tile (multiple of 16, <= 256)
    const float* bias;                     // [groups * n_valid] or nullptr
    void* out;
    // ---- split-K (plain GEMMs with few output tiles: the K range is dealt to ksplit CTAs per tile,
    // each writes its raw fp32 partial tile to partial + split * partial_stride; splitk_finish_kernel
    // adds them in split order, then bias / ReLU / conversion).  ksplit <= 1: off.
    int ksplit;
    float* partial;
    long long partial_stride;
};

// ---- PTX wrappers -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps instead of hanging the GPU (the host then reports the failing launch by name,
// check_launch / KernelScope).  The bound is wall time on the device (20 s of %globaltimer once 2^22 polls have
// failed), not a poll count alone, so a time-sliced or preempted context is not mistaken for one.  No printf here: a call in
// this kernel costs registers in the pipeline loops (measured: conv2 0.46 -> 0.51 ms).
__device__ __forceinline__ unsigned long long mbar_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 22); ++i)
        if (mbar_try_wait(bar, parity)) return;
    // not reached in a run that is neither broken nor descheduled for long: now bound the wait by time
    const unsigned long long t0 = mbar_globaltimer();
    while (!mbar_try_wait(bar, parity))
        if (mbar_globaltimer() - t0 > 20000000000ull) __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, 128-byte swizzle operand descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);        // start address  [0,14)
    d |= (uint64_t)1 << 16;                        // LBO (unused for swizzled K-major) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;              // SBO = 8 rows * 128 B  [32,46)
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M x N
__device__ __forceinline__ uint32_t umma_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the same load without the wait: the registers may be used only after tmem_ld_wait(r) (which names them, so that the
// compiler keeps every use behind it); lets the epilogue work on one chunk while the next one is on its way
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

template <int kStages>
__global__ void __launch_bounds__(kGemmThreads)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[kStages], bar_empty[kStages], bar_acc;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_bytes = (uint32_t)p.bn * kBK * 2;
    const uint32_t sA = sbase, sB = sbase + kStages * kABytes;
    const int m0 = blockIdx.x * kBM;
    const int n0 = blockIdx.y * p.bn;
    const int nsplit = p.ksplit > 1 ? p.ksplit : 1;
    const int nfold = p.fold > 1 ? p.fold : 1;
    const int g = (blockIdx.z / nsplit) * nfold, ksp = blockIdx.z - (blockIdx.z / nsplit) * nsplit;      // first group of this CTA
    const int kb_per = (p.k_blocks + nsplit - 1) / nsplit;
    const int kb0 = ksp * kb_per, nkb = min(kb_per, p.k_blocks - kb0);       // host guarantees nkb >= 1
    uint32_t ncols = 32;
    while ((int)ncols < p.bn * nfold) ncols <<= 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        mbar_init(smem_u32(&bar_acc), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            for (int it = 0; it < nkb * nfold; ++it) {
                const int gi = it / nkb, i = it - gi * nkb, ge = g + gi;
                const int kb = kb0 + i;
                const int s = it % kStages;
                const uint32_t ph = (uint32_t)(it / kStages) & 1u;
                mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
                const uint32_t full = smem_u32(&bar_full[s]);
                mbar_expect_tx(full, kABytes + b_bytes);
                const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
                const int kh = tap / p.taps_x, kw = tap - kh * p.taps_x;
                tma_load_2d(sA + s * kABytes, &tmA, full, ge * p.a_col_group + cb * kBK, m0 + kh * p.row_pitch + kw);
                tma_load_2d(sB + s * b_bytes, &tmB, full, kb * kBK, ge * p.b_row_group + n0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = umma_idesc_bf16(kBM, p.bn);
        for (int it = 0; it < nkb * nfold; ++it) {
            const int gi = it / nkb, i = it - gi * nkb;
            const int s = it % kStages;
            const uint32_t ph = (uint32_t)(it / kStages) & 1u;
            mbar_wait(smem_u32(&bar_full[s]), ph);
            tcgen05_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                    uint64_t ad = umma_desc_sw128(sA + s * kABytes + k * 32);
                    uint64_t bd = umma_desc_sw128(sB + s * b_bytes + k * 32);
                    tcgen05_mma_bf16(tmem + (uint32_t)(gi * p.bn), ad, bd, idesc, (i | k) ? 1u : 0u);
                }
                tcgen05_commit(smem_u32(&bar_empty[s]));        // frees the smem stage when the MMAs retire
                if (it == nkb * nfold - 1) tcgen05_commit(smem_u32(&bar_acc));
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: TMEM -> registers -> bias/ReLU -> global =====
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may touch
        const int half = (warp - 2) >> 2;             // which half of the tile's columns this warp writes
        const int row = quarter * 32 + lane;
        const int m = m0 + row;
        bool valid = m < p.m_total;
        long long orow = 0;
        if (valid) {
            int n = m / p.hp_wp, r = m - n * p.hp_wp;
            int y = r / p.wp, x = r - y * p.wp;
            valid = (y < p.h_valid) && (x < p.w_valid);
            orow = (long long)n * p.out_hp_wp + (long long)(y + p.out_pad) * p.out_wp + (x + p.out_pad);
        }
        mbar_wait(smem_u32(&bar_acc), 0);
        tcgen05_fence_after();
        const int ccol0 = g * p.c_col_group + n0;
        if (p.lrn) {
            // Every thread owns one pixel with ALL its channels (the tile covers n_valid), so the cross-channel window is
            // thread-local: walk the channels in chunks of 16 with the neighbouring two of the previous / next chunk.
            const int nch = nfold * p.n_valid;                 // channels of the pixel (folded groups lie side by side in TMEM)
            const int cmid = ((nch / 16 + 1) / 2) * 16;        // this warp normalises the channels [cb, ce) and peeks two beyond each end
            const int cb = half ? cmid : 0, ce = half ? nch : cmid;
            float prev0 = 0.f, prev1 = 0.f, cur[16];
            uint32_t raw[16];
            const float* bias = p.bias ? p.bias + g * p.n_valid : nullptr;
            auto issue = [&](int c) { tmem_ld16_issue(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c, raw); };
            auto finish = [&](int c, float (&v)[16]) {           // bias + ReLU of the chunk that has arrived in raw
                tmem_ld_wait(raw);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float x = __uint_as_float(raw[j]);
                    if (bias) x += __ldg(bias + c + j);
                    x = p.relu ? fmaxf(x, 0.f) : x;
                    v[j] = (c + j < nch) ? x : 0.f;
                }
            };
            if (cb > 0) { issue(cb - 16); finish(cb - 16, cur); prev0 = cur[14]; prev1 = cur[15]; }
            issue(cb);
            finish(cb, cur);
            for (int c = cb; c < ce; c += 16) {
                const bool more = c + 16 < nch;
                if (more) issue(c + 16);                          // on its way while this chunk is squared
                float sq[20];
                sq[0] = prev0 * prev0; sq[1] = prev1 * prev1;
#pragma unroll
                for (int j = 0; j < 16; ++j) sq[2 + j] = cur[j] * cur[j];
                // channels 0 .. 13 of the chunk need nothing of the next one: normalise them while it arrives
                uint32_t pk[8];
                auto norm_pair = [&](int j) {
                    const float s0 = sq[2 * j] + sq[2 * j + 1] + sq[2 * j + 2] + sq[2 * j + 3] + sq[2 * j + 4];
                    const float s1 = sq[2 * j + 1] + sq[2 * j + 2] + sq[2 * j + 3] + sq[2 * j + 4] + sq[2 * j + 5];
                    const float y0 = cur[2 * j] * __powf(1.f + (1e-4f / 5.f) * s0, -0.75f);
                    const float y1 = cur[2 * j + 1] * __powf(1.f + (1e-4f / 5.f) * s1, -0.75f);
                    __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);
                    pk[j] = *reinterpret_cast<uint32_t*>(&h);
                };
#pragma unroll
                for (int j = 0; j < 7; ++j) norm_pair(j);
                float nxt[16];
                if (more) finish(c + 16, nxt);
                sq[18] = more ? nxt[0] * nxt[0] : 0.f;
                sq[19] = more ? nxt[1] * nxt[1] : 0.f;
                norm_pair(7);
                if (valid) {
                    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldc + ccol0 + c);
                    o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
                prev0 = cur[14]; prev1 = cur[15];
#pragma unroll
                for (int j = 0; j < 16; ++j) cur[j] = more ? nxt[j] : 0.f;
            }
        } else
        for (int c = half ? ((p.bn / 16 + 1) / 2) * 16 : 0; c < (half ? p.bn : ((p.bn / 16 + 1) / 2) * 16); c += 16) {
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c, r);
            if (!valid || n0 + c >= p.n_valid) continue;
            if (p.ksplit > 1) {
                float4* o = reinterpret_cast<float4*>(p.partial + (long long)ksp * p.partial_stride + orow * p.ldc + ccol0 + c);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    o[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                       __uint_as_float(r[4 * j + 3]));
                continue;
            }
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float x = __uint_as_float(r[j]);
                if (p.bias) x += __ldg(p.bias + g * p.n_valid + n0 + c + j);
                v[j] = p.relu ? fmaxf(x, 0.f) : x;
            }
            if (p.out_f32) {
                float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.ldc + ccol0 + c);
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                    pk[j] = *reinterpret_cast<uint32_t*>(&h);
                }
                uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldc + ccol0 + c);
                o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols) : "memory");
    }
}

// out[m, n] = act(bias[n] + sum_s partial[s][m][n]), splits added in order (deterministic)
__global__ void __launch_bounds__(256) splitk_finish_kernel(const float* __restrict__ partial, int ksplit, long long stride, int M,
                                                            int N, int ld, const float* __restrict__ bias, int relu, int out_f32,
                                                            void* __restrict__ out) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int nq = N >> 2;
    if (t >= (long long)M * nq) return;
    const int m = (int)(t / nq), n = (int)(t - (long long)m * nq) * 4;
    const float* src = partial + (long long)m * ld + n;
    float4 a = *reinterpret_cast<const float4*>(src);
    for (int s = 1; s < ksplit; ++s) {
        const float4 b = *reinterpret_cast<const float4*>(src + s * stride);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (bias) { a.x += bias[n]; a.y += bias[n + 1]; a.z += bias[n + 2]; a.w += bias[n + 3]; }
    if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
    if (out_f32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (long long)m * ld + n) = a;
    } else {
        __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(a.z, a.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&lo); o.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + (long long)m * ld + n) = o;
    }
}


// ---------------------------------------------------------------------------
// Persistent form: one CTA per SM slot walks the output tiles (m fastest, so neighbouring CTAs share the weight
// tile in the L2) with kAccBufs = TWO accumulators in tensor memory: while the epilogue warps drain tile j from one buffer,
// the MMA warp accumulates tile j + 1 into the other and the TMA warp is already filling the ring for it.  The
// shared-memory ring and its phases run on across tiles; barriers are initialised and tensor memory is allocated once
// per CTA.  kAccBufs = 1 (a tile that needs 256 columns, two CTAs per SM): the main loop waits for the epilogue, but the
// ring is refilled for the next tile meanwhile.  No split-K here (the FC layers keep the one-tile-per-CTA kernel above).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int kStages, int kEpiWarps, int kAccBufs>
__global__ void __launch_bounds__(64 + 32 * kEpiWarps)
gemm_bf16_tcgen05_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p,
                                    const int m_tiles, const int n_tiles, const int total_tiles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[kStages], bar_empty[kStages], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_bytes = (uint32_t)p.bn * kBK * 2;
    const uint32_t sA = sbase, sB = sbase + kStages * kABytes;
    const int nfold = p.fold > 1 ? p.fold : 1;
    const int nkb = p.k_blocks;
    const int per_tile = nkb * nfold;                      // K blocks streamed per tile
    uint32_t ncols1 = 32;                                  // columns of one accumulator
    while ((int)ncols1 < p.bn * nfold) ncols1 <<= 1;
    const uint32_t ncols = kAccBufs * ncols1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        for (int b = 0; b < kAccBufs; ++b) { mbar_init(smem_u32(&acc_full[b]), 1); mbar_init(smem_u32(&acc_empty[b]), 32 * kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            uint32_t it = 0;                               // K blocks issued by this CTA so far (ring position)
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int mt = t % m_tiles, rest = t / m_tiles, nt = rest % n_tiles, g = (rest / n_tiles) * nfold;
                const int m0 = mt * kBM, n0 = nt * p.bn;
                for (int q = 0; q < per_tile; ++q, ++it) {
                    const int gi = q / nkb, kb = q - gi * nkb, ge = g + gi;
                    const int s = (int)(it % kStages);
                    const uint32_t ph = (it / kStages) & 1u;
                    mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
                    const uint32_t full = smem_u32(&bar_full[s]);
                    mbar_expect_tx(full, kABytes + b_bytes);
                    const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
                    const int kh = tap / p.taps_x, kw = tap - kh * p.taps_x;
                    tma_load_2d(sA + s * kABytes, &tmA, full, ge * p.a_col_group + cb * kBK, m0 + kh * p.row_pitch + kw);
                    tma_load_2d(sB + s * b_bytes, &tmB, full, kb * kBK, ge * p.b_row_group + n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = umma_idesc_bf16(kBM, p.bn);
        uint32_t it = 0, j = 0;                            // ring position, tiles done by this CTA
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
            const uint32_t buf = j % kAccBufs, use = j / kAccBufs;
            mbar_wait(smem_u32(&acc_empty[buf]), (use & 1u) ^ 1u);            // the epilogue has drained this accumulator
            tcgen05_fence_after();
            const uint32_t tacc = tmem + buf * ncols1;
            for (int q = 0; q < per_tile; ++q, ++it) {
                const int gi = q / nkb, i = q - gi * nkb;
                const int s = (int)(it % kStages);
                const uint32_t ph = (it / kStages) & 1u;
                mbar_wait(smem_u32(&bar_full[s]), ph);
                tcgen05_fence_after();
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        uint64_t ad = umma_desc_sw128(sA + s * kABytes + k * 32);
                        uint64_t bd = umma_desc_sw128(sB + s * b_bytes + k * 32);
                        tcgen05_mma_bf16(tacc + (uint32_t)(gi * p.bn), ad, bd, idesc, (i | k) ? 1u : 0u);
                    }
                    tcgen05_commit(smem_u32(&bar_empty[s]));        // frees the smem stage when the MMAs retire
                    if (q == per_tile - 1) tcgen05_commit(smem_u32(&acc_full[buf]));
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> bias/ReLU(/LRN) -> global =====
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may touch
        constexpr int kParts = kEpiWarps / 4;         // warps per lane quarter: each takes a range of the tile's columns
        const int part = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        auto part_start = [&](int chunks, int q) { return ((q * chunks + kParts - 1) / kParts) * 16; };     // in columns
        uint32_t j = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
            const int mt = t % m_tiles, rest = t / m_tiles, nt = rest % n_tiles, g = (rest / n_tiles) * nfold;
            const int m0 = mt * kBM, n0 = nt * p.bn;
            const uint32_t buf = j % kAccBufs, use = j / kAccBufs;
            const uint32_t tacc = tmem + buf * ncols1 + ((uint32_t)(quarter * 32) << 16);
            const int m = m0 + row;
            bool valid = m < p.m_total;
            long long orow = 0;
            if (valid) {
                int n = m / p.hp_wp, r = m - n * p.hp_wp;
                int y = r / p.wp, x = r - y * p.wp;
                valid = (y < p.h_valid) && (x < p.w_valid);
                orow = (long long)n * p.out_hp_wp + (long long)(y + p.out_pad) * p.out_wp + (x + p.out_pad);
            }
            mbar_wait(smem_u32(&acc_full[buf]), use & 1u);
            tcgen05_fence_after();
            const int ccol0 = g * p.c_col_group + n0;
            if (p.lrn) {
                const int nch = nfold * p.n_valid;
                const int cb = part_start(nch / 16, part), ce = part + 1 < kParts ? part_start(nch / 16, part + 1) : nch;
                float prev0 = 0.f, prev1 = 0.f, cur[16];
                uint32_t raw[16];
                const float* bias = p.bias ? p.bias + g * p.n_valid : nullptr;
                auto issue = [&](int c) { tmem_ld16_issue(tacc + (uint32_t)c, raw); };
                auto finish = [&](int c, float (&v)[16]) {           // bias + ReLU of the chunk that has arrived in raw
                    tmem_ld_wait(raw);
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) {
                        float x = __uint_as_float(raw[jj]);
                        if (bias) x += __ldg(bias + c + jj);
                        x = p.relu ? fmaxf(x, 0.f) : x;
                        v[jj] = (c + jj < nch) ? x : 0.f;
                    }
                };
                if (cb > 0) { issue(cb - 16); finish(cb - 16, cur); prev0 = cur[14]; prev1 = cur[15]; }
                issue(cb);
                finish(cb, cur);
                for (int c = cb; c < ce; c += 16) {
                    const bool more = c + 16 < nch;
                    if (more) issue(c + 16);
                    float sq[20];
                    sq[0] = prev0 * prev0; sq[1] = prev1 * prev1;
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) sq[2 + jj] = cur[jj] * cur[jj];
                    uint32_t pk[8];
                    auto norm_pair = [&](int jj) {
                        const float s0 = sq[2 * jj] + sq[2 * jj + 1] + sq[2 * jj + 2] + sq[2 * jj + 3] + sq[2 * jj + 4];
                        const float s1 = sq[2 * jj + 1] + sq[2 * jj + 2] + sq[2 * jj + 3] + sq[2 * jj + 4] + sq[2 * jj + 5];
                        const float y0 = cur[2 * jj] * __powf(1.f + (1e-4f / 5.f) * s0, -0.75f);
                        const float y1 = cur[2 * jj + 1] * __powf(1.f + (1e-4f / 5.f) * s1, -0.75f);
                        __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);
                        pk[jj] = *reinterpret_cast<uint32_t*>(&h);
                    };
#pragma unroll
                    for (int jj = 0; jj < 7; ++jj) norm_pair(jj);
                    float nxt[16];
                    if (more) finish(c + 16, nxt);
                    if (c + 16 >= ce) {                    // last TMEM read of this thread for this tile: hand the accumulator back
                        tcgen05_fence_before();
                        mbar_arrive(smem_u32(&acc_empty[buf]));
                    }
                    sq[18] = more ? nxt[0] * nxt[0] : 0.f;
                    sq[19] = more ? nxt[1] * nxt[1] : 0.f;
                    norm_pair(7);
                    if (valid) {
                        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldc + ccol0 + c);
                        o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                    prev0 = cur[14]; prev1 = cur[15];
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) cur[jj] = more ? nxt[jj] : 0.f;
                }
                if (cb >= ce) { tcgen05_fence_before(); mbar_arrive(smem_u32(&acc_empty[buf])); }          // no channels for this half
            } else {
                const int c_lo = part_start(p.bn / 16, part), c_hi = part + 1 < kParts ? part_start(p.bn / 16, part + 1) : p.bn;
                for (int c = c_lo; c < c_hi; c += 16) {
                    uint32_t r[16];
                    tmem_ld16(tacc + (uint32_t)c, r);
                    if (c + 16 >= c_hi) {                  // last TMEM read of this thread for this tile
                        tcgen05_fence_before();
                        mbar_arrive(smem_u32(&acc_empty[buf]));
                    }
                    if (!valid || n0 + c >= p.n_valid) continue;
                    float v[16];
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) {
                        float x = __uint_as_float(r[jj]);
                        if (p.bias) x += __ldg(p.bias + g * p.n_valid + n0 + c + jj);
                        v[jj] = p.relu ? fmaxf(x, 0.f) : x;
                    }
                    if (p.out_f32) {
                        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.ldc + ccol0 + c);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) o[jj] = make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
                    } else {
                        uint32_t pk[8];
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * jj], v[2 * jj + 1]);
                            pk[jj] = *reinterpret_cast<uint32_t*>(&h);
                        }
                        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldc + ccol0 + c);
                        o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                }
                if (c_lo >= c_hi) { tcgen05_fence_before(); mbar_arrive(smem_u32(&acc_empty[buf])); }      // no columns for this half
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols) : "memory");
    }
}

}  // namespace vpk
