// Stage S0 + S1: line construction and sphere mapping.
//
// Replaces the Python loop of reference evaluation.py:158-168 and
// sphere_mapping.sphere_line_plot (sphere_mapping.py:36-72) with the index
// maps of coordinate_conversion.py:4-61.
//
// All geometry is float64 with explicitly rounded operations (__dmul_rn & co,
// never contracted into FMAs) in the same order as the numpy expressions the
// reference evaluates, so that histogram bin indices are bit-exact against the
// float64 CPU path; only asin/cos/atan/sincos can differ (by an ulp), which
// moves a vote across a cell border with probability ~1e-13.
//
// Accumulation is integer (uint32 counts, or Q.16 fixed point in 64-bit for
// weighted votes), so the result is independent of the order in which the
// atomics land: deterministic without ordered accumulation.
#include <math.h>
#include "vpk_internal.cuh"

namespace vpk {

static constexpr int kTile = 128;            // lines per tile of the pair space
static constexpr int kVoteThreads = 256;
static constexpr int kNumSamples = 10000;    // sphere_mapping.py:40
static constexpr int kLutMax = 4096;
static constexpr double kPi = 3.141592653589793;   // == numpy.pi

// ---------------------------------------------------------------------------
// S0
// ---------------------------------------------------------------------------
__global__ void lines_from_segments_kernel(const double* __restrict__ seg, int64_t n, double* __restrict__ lines) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2* s2 = reinterpret_cast<const double2*>(seg) + 2 * i;
    double2 p1 = s2[0], p2 = s2[1];
    // [x1,y1,1] x [x2,y2,1]   (evaluation.py:163-167, numpy.cross component order)
    lines[3 * i + 0] = __dsub_rn(p1.y, p2.y);
    lines[3 * i + 1] = __dsub_rn(p2.x, p1.x);
    lines[3 * i + 2] = __dsub_rn(__dmul_rn(p1.x, p2.y), __dmul_rn(p1.y, p2.x));
}

int lines_from_segments_dev(vpk_ctx* ctx, const double* d_seg, int64_t n, double* d_lines) {
    if (n <= 0) return VPK_OK;
    KernelScope ks(ctx, "lines_from_segments");
    lines_from_segments_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_seg, n, d_lines);
    return check_launch("lines_from_segments");
}

// Row N2 of SURVEY.md section 8(f): the normalisation of the raw LSD output that evaluation.detect_lsd_lines
// applies (reference evaluation.py:240-249): pixel coordinates -> origin at the image centre, divided by
// max(width, height) / 2, y pointing up; column 6 of the LSD rows (-log10 NFA) is passed through (:251).
// Every operation is a single correctly rounded float64 operation in the reference's order, so the
// result is bit-identical to numpy's.  One thread per segment; the image of a segment is found by
// bisection over the offsets.
__global__ void segments_from_lsd_kernel(const double* __restrict__ lsd, int ncols, const int32_t* __restrict__ offsets,
                                         const int32_t* __restrict__ widths, const int32_t* __restrict__ heights, int B, int64_t n,
                                         double* __restrict__ seg, double* __restrict__ nfa) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = B - 1;
    while (lo < hi) {                               // last image with offsets[b] <= i
        const int mid = (lo + hi + 1) >> 1;
        if ((int64_t)offsets[mid] <= i) lo = mid; else hi = mid - 1;
    }
    const double w = (double)widths[lo], h = (double)heights[lo];
    const double half_w = __ddiv_rn(w, 2.0), half_h = __ddiv_rn(h, 2.0);
    const double half_s = __ddiv_rn(fmax(w, h), 2.0);          // scale_w = scale_h = max(width, height) (:233-234)
    const double* r = lsd + i * ncols;
    const double x1 = __ddiv_rn(__dsub_rn(r[0], half_w), half_s), y1 = __ddiv_rn(__dsub_rn(r[1], half_h), half_s);
    const double x2 = __ddiv_rn(__dsub_rn(r[2], half_w), half_s), y2 = __ddiv_rn(__dsub_rn(r[3], half_h), half_s);
    seg[4 * i] = x1; seg[4 * i + 1] = __dmul_rn(y1, -1.0); seg[4 * i + 2] = x2; seg[4 * i + 3] = __dmul_rn(y2, -1.0);
    if (nfa) nfa[i] = ncols > 6 ? r[6] : 0.0;
}

int segments_from_lsd_dev(vpk_ctx* ctx, const double* d_lsd, int ncols, const int32_t* d_offsets, const int32_t* d_widths,
                          const int32_t* d_heights, int B, int64_t n, double* d_seg, double* d_nfa) {
    if (n <= 0) return VPK_OK;
    KernelScope ks(ctx, "segments_from_lsd");
    segments_from_lsd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_lsd, ncols, d_offsets, d_widths, d_heights, B, n,
                                                                                 d_seg, d_nfa);
    return check_launch("segments_from_lsd");
}

// ---------------------------------------------------------------------------
// shared index maps
// ---------------------------------------------------------------------------
// coordinate_conversion.py:29-30 followed by round-to-cell, clipped.
__device__ __forceinline__ int angle_bin(double angle, double half_over_s, double s) {
    double a = __dmul_rn(__dsub_rn(__dadd_rn(__ddiv_rn(angle, kPi), 0.5), half_over_s), s);
    double r = floor(__dadd_rn(a, 0.5));
    r = fmin(fmax(r, 0.0), s - 1.0);
    return (int)r;
}

// ---------------------------------------------------------------------------
// S1 votes
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool intersection_cell(double ax, double ay, double az, double bx, double by, double bz,
                                                  double half_over_s, double s, int S, int& row, int& col) {
    double px = __dsub_rn(__dmul_rn(ay, bz), __dmul_rn(az, by));
    double py = __dsub_rn(__dmul_rn(az, bx), __dmul_rn(ax, bz));
    double pz = __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
    double n = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)), __dmul_rn(pz, pz)));
    if (!(n > 0.0) || isinf(n)) return false;
    if (pz < 0.0) { px = -px; py = -py; }
    double y = fmin(fmax(__ddiv_rn(py, n), -1.0), 1.0);
    double beta = asin(y);
    double inner = __ddiv_rn(__ddiv_rn(px, n), cos(beta));
    if (isnan(inner)) return false;
    inner = fmin(fmax(inner, -1.0), 1.0);
    double alpha = asin(inner);
    col = angle_bin(alpha, half_over_s, s);
    row = (S - 1) - angle_bin(beta, half_over_s, s);
    return true;
}

// ---- FP32 pre-binning ---------------------------------------------------------------------------------
// The cell of an intersection is needed bit-exactly as the float64 expressions above give it, but a cell is
// pi / S wide (6.3e-3 rad at S = 500) while float32 resolves ~1e-6 rad: the cell is computed in float32 together
// with a rigorous bound of its own error, and only the pairs whose float32 position lies closer to a cell border
// than that bound (a few per cent) are re-evaluated with the float64 expressions.  Error budget of the fast path:
//   p = a x b from operands rounded to float32:   |dp_k| <= 2^-21 * mag,  mag = sum of the six |products|
//   direction on the sphere:                      eps_s  <= 2 * 2^-21 * mag / |p| + 2^-21   (normalisation, rsqrt)
//   beta  = asin(y):                              |dbeta|  <= eps_s / cos(beta) + 4e-7       (asinf: a few ulp of pi/2)
//   alpha = asin(x / cos(beta)):                  |dalpha| <= eps_s (1/c + 1/c^2) / sqrt(1 - inner^2) + 4e-7
//   index = angle * S / pi + S / 2 (float32):     1.5e-4 index units of rounding at magnitudes up to 512 (S <= 1024)
// and everything is doubled once more.  Pairs near the poles or the alpha = +-pi/2 seam (c or sqrt(1 - inner^2)
// below 0.05), degenerate pairs and images with S > 1024 always take the float64 path.
__device__ __forceinline__ bool fast_cell(float ax, float ay, float az, float bx, float by, float bz, float s_over_pi,
                                          float half_s, int S, int& row, int& col) {
    const float t0 = ay * bz, t1 = az * by, t2 = az * bx, t3 = ax * bz, t4 = ax * by, t5 = ay * bx;
    float px = t0 - t1, py = t2 - t3;
    const float pz = t4 - t5;
    const float mag = (fabsf(t0) + fabsf(t1)) + (fabsf(t2) + fabsf(t3)) + (fabsf(t4) + fabsf(t5));
    const float n2 = px * px + py * py + pz * pz;
    if (!(n2 > 1e-30f) || !(n2 < 1e30f)) return false;
    const float rn = rsqrtf(n2);
    if (pz < 0.0f) { px = -px; py = -py; }
    const float y = py * rn, x = px * rn;
    const float c2 = 1.0f - y * y;
    if (!(c2 > 0.0025f)) return false;                       // cos(beta) < 0.05
    const float rc = rsqrtf(c2);                             // 1 / cos(beta)
    const float inner = x * rc;
    const float q2 = 1.0f - inner * inner;
    if (!(q2 > 0.0025f)) return false;
    const float eps = 9.5367431640625e-7f * (mag * rn) + 4.76837158203125e-7f;          // 2^-20 mag / |p| + 2^-21
    const float dbeta = eps * rc + 4e-7f;
    const float dalpha = eps * (rc + rc * rc) * rsqrtf(q2) + 4e-7f;
    const float fa = asinf(inner) * s_over_pi + half_s;      // angle_to_index + 0.5: the cell is floor(fa)
    const float fb = asinf(y) * s_over_pi + half_s;
    const float ma = 2.0f * (dalpha * s_over_pi + 1.5e-4f), mb = 2.0f * (dbeta * s_over_pi + 1.5e-4f);
    const float ra = floorf(fa), rb = floorf(fb);
    if (fa - ra < ma || ra + 1.0f - fa < ma || fb - rb < mb || rb + 1.0f - fb < mb) return false;
    // away from the borders: the float64 cell is the same; the clip of angle_bin applies to both alike
    col = min(max((int)ra, 0), S - 1);
    row = (S - 1) - min(max((int)rb, 0), S - 1);
    return true;
}

// Work list of the vote kernel, built on the device: item k = (image, tile_i, tile_j), tile_i <= tile_j, of the
// upper-triangular tile-pair space of every image.  One block; images in chunks of blockDim.x with a running prefix.
__global__ void sphere_items_kernel(const int32_t* __restrict__ offsets, int B, int4* __restrict__ work) {
    __shared__ int s_warp[32], s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int b0 = 0; b0 < B; b0 += blockDim.x) {
        const int b = b0 + threadIdx.x;
        int T = 0;
        if (b < B) T = (offsets[b + 1] - offsets[b] + kTile - 1) / kTile;
        const int mine = T * (T + 1) / 2;
        int incl = mine;
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        int k = before + incl - mine;
        for (int ti = 0; ti < T; ++ti)
            for (int tj = ti; tj < T; ++tj) work[k++] = make_int4(b, ti, tj, 0);
        __syncthreads();
        if (threadIdx.x == 0) { int tot = 0; for (int w = 0; w < nw; ++w) tot += s_warp[w]; s_base += tot; }
        __syncthreads();
    }
}

// One CTA per (image, tile_i, tile_j) work item.  Tiles are staged in shared memory as SoA, in float32 for the
// fast path and float64 for the exact one.  Votes of a warp that fall into the same cell are merged
// (__match_any_sync) before the integer atomic; pairs the fast path cannot decide are queued in shared memory and
// evaluated in float64 afterwards by all threads (no divergence inside the pair loop).
constexpr int kQueueCap = 4096;
__global__ void __launch_bounds__(kVoteThreads)
sphere_votes_kernel(const double* __restrict__ lines, const int32_t* __restrict__ offsets,
                    const int4* __restrict__ work, int S, const double* __restrict__ weights,
                    uint32_t* __restrict__ hist, unsigned long long* __restrict__ whist, unsigned long long* __restrict__ stats) {
    __shared__ double sA[3][kTile], sB[3][kTile], sWA[kTile], sWB[kTile];
    __shared__ float fA[3][kTile], fB[3][kTile];
    __shared__ unsigned short s_queue[kQueueCap];
    __shared__ int s_nq;
    const int4 item = work[blockIdx.x];
    const int b = item.x, ti = item.y, tj = item.z;
    const int base = offsets[b];
    const int N = offsets[b + 1] - base;
    const int i0 = ti * kTile, j0 = tj * kTile;
    if (threadIdx.x == 0) s_nq = 0;
    for (int t = threadIdx.x; t < kTile; t += blockDim.x) {
        int gi = i0 + t, gj = j0 + t;
        bool vi = gi < N, vj = gj < N;
        const double* li = lines + 3 * (int64_t)(base + (vi ? gi : 0));
        const double* lj = lines + 3 * (int64_t)(base + (vj ? gj : 0));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double a = li[c], bb = lj[c];
            sA[c][t] = a; sB[c][t] = bb;
            fA[c][t] = (float)a; fB[c][t] = (float)bb;
        }
        if (weights) {
            sWA[t] = weights[base + (vi ? gi : 0)];
            sWB[t] = weights[base + (vj ? gj : 0)];
        }
    }
    __syncthreads();
    const double s = (double)S;
    const double half_over_s = __ddiv_rn(0.5, s);
    const float s_over_pi = (float)(s / kPi), half_s = 0.5f * (float)S;
    const bool fast_ok = S <= 1024;
    const int64_t plane = (int64_t)b * S * S;
    const int ni = min(kTile, N - i0), nj = min(kTile, N - j0);
    const int lane = threadIdx.x & 31;
    auto vote = [&](bool valid, int row, int col, int i, int j) {
        if (weights) {
            if (valid) {
                double q = floor(__dadd_rn(__dmul_rn(__dmul_rn(sWA[i], sWB[j]), 65536.0), 0.5));
                if (q > 0.0) atomicAdd(whist + plane + (int64_t)row * S + col, (unsigned long long)q);
            }
            return;
        }
        // lanes voting for the same cell elect a leader that adds their number
        const int cell = valid ? row * S + col : -1 - lane;
        const unsigned same = __match_any_sync(0xffffffffu, cell);
        if (valid && lane == __ffs(same) - 1) atomicAdd(hist + plane + cell, (unsigned)__popc(same));
    };
    const int total = ni * kTile;
    for (int idx0 = 0; idx0 < total; idx0 += blockDim.x) {         // warp-uniform trip count (match_any needs full warps)
        const int idx = idx0 + threadIdx.x;
        const int i = idx / kTile, j = idx % kTile;
        const bool pair = idx < total && j < nj && i0 + i < j0 + j;   // i < j only
        int row = 0, col = 0;
        bool done = false;
        if (pair) {
            done = fast_ok && fast_cell(fA[0][i], fA[1][i], fA[2][i], fB[0][j], fB[1][j], fB[2][j], s_over_pi, half_s, S, row, col);
            if (!done) {
                const int q = atomicAdd(&s_nq, 1);
                if (q < kQueueCap) s_queue[q] = (unsigned short)((i << 8) | j);
                else {
                    // queue full (a tile of near-degenerate pairs): decide in place
                    done = intersection_cell(sA[0][i], sA[1][i], sA[2][i], sB[0][j], sB[1][j], sB[2][j], half_over_s, s, S, row, col);
                }
            }
        }
        vote(pair && done, row, col, i, j);
    }
    __syncthreads();
    const int nq = min(s_nq, kQueueCap);
    if (stats && threadIdx.x == 0) { atomicAdd(stats + 0, (unsigned long long)nq); atomicAdd(stats + 1, 1ull); }
    for (int q0 = 0; q0 < nq; q0 += blockDim.x) {
        const int q = q0 + threadIdx.x;
        int row = 0, col = 0, i = 0, j = 0;
        bool ok = false;
        if (q < nq) {
            const int e = s_queue[q];
            i = e >> 8; j = e & 255;
            ok = intersection_cell(sA[0][i], sA[1][i], sA[2][i], sB[0][j], sB[1][j], sB[2][j], half_over_s, s, S, row, col);
        }
        vote(ok, row, col, i, j);
    }
}

// per-image maximum of the histogram (uint32 counts or Q.16 sums)
template <typename T>
__global__ void plane_max_kernel(const T* __restrict__ h, int64_t plane, unsigned long long* __restrict__ maxv) {
    const int b = blockIdx.y;
    const T* p = h + (int64_t)b * plane;
    unsigned long long m = 0;
    constexpr int V = 16 / sizeof(T);                      // elements per 16-byte load
    const bool vec = ((plane * sizeof(T)) % 16 == 0) && (reinterpret_cast<uintptr_t>(p) % 16 == 0);
    if (vec) {
        const int64_t nv = plane / V;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
            const uint4 q = reinterpret_cast<const uint4*>(p)[i];
            const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
            for (int k = 0; k < V; ++k) { const unsigned long long v = (unsigned long long)e[k]; m = v > m ? v : m; }
        }
    } else {
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < plane; i += (int64_t)gridDim.x * blockDim.x) {
            unsigned long long v = (unsigned long long)p[i];
            m = v > m ? v : m;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
        m = t > m ? t : m;
    }
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(maxv + b, m);
}

// floor(255 v / m) in exact integer arithmetic (32-bit division when it cannot overflow)
__device__ __forceinline__ uint8_t scale_to_u8(unsigned long long v, unsigned long long m) {
    if (m == 0) return 0;
    if (m <= 0x00ffffffull) return (uint8_t)(((uint32_t)v * 255u) / (uint32_t)m);      // v <= m < 2^24: 255 v < 2^32
    return (uint8_t)((v * 255ull) / m);
}

// image = floor(255*h/max) in exact integer arithmetic; optional float export
template <typename T>
__global__ void votes_image_kernel(const T* __restrict__ h, int64_t plane, const unsigned long long* __restrict__ maxv,
                                   uint8_t* __restrict__ img, float* __restrict__ wout) {
    const int b = blockIdx.y;
    const unsigned long long m = maxv[b];
    const T* p = h + (int64_t)b * plane;
    // uint32 counts, image only: 16 pixels per thread (four 16-byte loads, one 16-byte store)
    if (sizeof(T) == 4 && img && !wout && plane % 16 == 0 && reinterpret_cast<uintptr_t>(p) % 16 == 0 &&
        reinterpret_cast<uintptr_t>(img + (int64_t)b * plane) % 16 == 0) {
        const int64_t nv = plane / 16;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 v = reinterpret_cast<const uint4*>(p)[4 * i + q];
                o[q] = (uint32_t)scale_to_u8(v.x, m) | ((uint32_t)scale_to_u8(v.y, m) << 8) | ((uint32_t)scale_to_u8(v.z, m) << 16) |
                       ((uint32_t)scale_to_u8(v.w, m) << 24);
            }
            reinterpret_cast<uint4*>(img + (int64_t)b * plane)[i] = make_uint4(o[0], o[1], o[2], o[3]);
        }
        return;
    }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < plane; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long v = (unsigned long long)p[i];
        if (img) img[(int64_t)b * plane + i] = scale_to_u8(v, m);
        if (wout) wout[(int64_t)b * plane + i] = (float)((double)v * (1.0 / 65536.0));
    }
}

// ---------------------------------------------------------------------------
// S1 curves (reference geometry)
// ---------------------------------------------------------------------------
// a_k of numpy.linspace(-pi/2, pi/2, 10000): k*step + start, last = stop.
__device__ __forceinline__ double sample_alpha(int k, double step) {
    if (k == kNumSamples - 1) return 0.5 * kPi;
    return __dadd_rn(__dmul_rn((double)k, step), -0.5 * kPi);
}

// row of sample k for line (l0,l1,l2); -1 if beta is NaN   (sphere_mapping.py:61-63)
__device__ __forceinline__ int sample_row(int k, double step, double l0, double l1, double l2,
                                          double half_over_s, double s, int S) {
    double sa, ca;
    sincos(sample_alpha(k, step), &sa, &ca);
    double g = __ddiv_rn(__dsub_rn(__dmul_rn(-l0, sa), __dmul_rn(l2, ca)), l1);
    double beta = atan(g);
    if (isnan(beta)) return -1;
    return (S - 1) - angle_bin(beta, half_over_s, s);
}

// One CTA per line; thread c handles column c: the covered rows are the closed
// interval [min,max] of the rows of the reference's samples in that column
// plus the first sample of the next column.  beta(alpha) is monotone between
// the stationary point alpha* = atan(l0/l2) and the ends, so only the end
// samples and the samples bracketing alpha* need evaluating.  The interval is
// recorded in a per-image difference array (+1 at rmin, -1 at rmax+1).
__global__ void sphere_curves_kernel(const double* __restrict__ lines, const int32_t* __restrict__ offsets, int B, int S,
                                     const int32_t* __restrict__ first, double f, int32_t* __restrict__ diff, int64_t line0) {
    const int64_t line = line0 + blockIdx.x;
    __shared__ int s_img;
    if (threadIdx.x == 0) {
        int lo = 0, hi = B;           // largest b with offsets[b] <= line
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (offsets[mid] <= line) lo = mid; else hi = mid;
        }
        s_img = lo;
    }
    __syncthreads();
    const int b = s_img;
    const double l0 = __dmul_rn(lines[3 * line + 0], f);      // sphere_mapping.py:55-56
    const double l1 = __dmul_rn(lines[3 * line + 1], f);
    const double l2 = lines[3 * line + 2];
    const double s = (double)S;
    const double half_over_s = __ddiv_rn(0.5, s);
    const double step = __ddiv_rn(kPi, (double)(kNumSamples - 1));
    const double astar = atan(l0 / l2);
    int kstar = isnan(astar) ? -10 : (int)floor((astar + 0.5 * kPi) / step);
    int32_t* d = diff + (int64_t)b * (S + 1) * S;
    for (int c = threadIdx.x; c < S; c += blockDim.x) {
        int k0 = first[c];
        int k1 = min(first[c + 1], kNumSamples - 1);   // inclusive: next column's first sample
        if (first[c + 1] <= k0) continue;
        int rmin = S, rmax = -1;
        auto take = [&](int k) {
            int r = sample_row(k, step, l0, l1, l2, half_over_s, s, S);
            if (r >= 0) { rmin = min(rmin, r); rmax = max(rmax, r); }
        };
        take(k0);
        if (k1 > k0) take(k1);
#pragma unroll
        for (int e = -1; e <= 2; ++e) {
            int k = kstar + e;
            if (k > k0 && k < k1) take(k);
        }
        if (rmax >= rmin) {
            atomicAdd(d + (int64_t)rmin * S + c, 1);
            atomicAdd(d + (int64_t)(rmax + 1) * S + c, -1);
        }
    }
}

// ---- band form (S <= kBandMaxS): no difference array in HBM ----------------------------------------
// sin / cos of the first sample of every column (entry S: the last sample), computed once per S with
// the same device sincos as sample_row, so the rows below are those of sample_row bit for bit.
__global__ void curves_trig_kernel(const int32_t* __restrict__ first, int S, double2* __restrict__ trig) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > S) return;
    const double step = __ddiv_rn(kPi, (double)(kNumSamples - 1));
    double sa, ca;
    sincos(sample_alpha(min(first[c], kNumSamples - 1), step), &sa, &ca);
    trig[c] = make_double2(sa, ca);
}

__device__ __forceinline__ int trig_row(double2 sc, double l0, double l1, double l2, double half_over_s, double s, int S) {
    const double g = __ddiv_rn(__dsub_rn(__dmul_rn(-l0, sc.x), __dmul_rn(l2, sc.y)), l1);
    const double beta = atan(g);
    if (isnan(beta)) return -1;
    return (S - 1) - angle_bin(beta, half_over_s, s);
}

// The same row in float32 with a rigorous error bound (the north_star's FP32 pre-binning, applied to the
// curves): returns false when the cell coordinate is closer to a cell border than twice the bound, and the
// caller then evaluates the float64 expression.  With u = 2^-24: the two products and their difference are
// off by <= 4u A, A = |l0 sa| + |l2 ca| (conversions of l0, l2, sa, ca included); times 1/l1 (<= 3u relative)
// |dg| <= 7u A / |l1| (10u is used); atan has slope 1 / (1 + g^2) and atanf is good to 1 ulp (2^-22 is used);
// the fused multiply-add into cell units adds <= 2^-23 S.  `tools/check_fast_row.py` replays this in numpy
// float32 with noise on the device functions: no accepted sample disagrees in 137 M, 99.9 % are accepted.
__device__ __forceinline__ bool fast_row(float2 sc, float neg_l0, float l2, float inv_l1, float inv_l1_abs, float sop,
                                         float hs, float fma_err, int S, int& row) {
    const float p = __fmul_rn(neg_l0, sc.x), q = __fmul_rn(l2, sc.y);
    const float g = __fmul_rn(__fsub_rn(p, q), inv_l1);
    const float eg = (10.0f * 5.9604645e-8f) * ((fabsf(p) + fabsf(q)) * inv_l1_abs);
    const float gm = fmaxf(fabsf(g) - eg, 0.0f);
    const float dbeta = __fdividef(eg, 1.0f + gm * gm) + 2.3841858e-7f;
    const float u = fmaf(atanf(g), sop, hs);
    const float m = 2.0f * (dbeta * sop + fma_err) + 1e-6f;
    const float fl = floorf(u), fr = u - fl;
    if (!(fr > m && 1.0f - fr > m && fabsf(u) < 1e9f)) return false;
    row = (S - 1) - min(max((int)fl, 0), S - 1);
    return true;
}

// One CTA per (band of kBand columns, image); two threads take a line (16 columns each) and walk their
// columns starting at column (lane / 2 mod 16), so that the lanes of a warp are in different columns =
// different shared-memory banks at any time.  The row of the sample on a column border is evaluated
// once and used by both neighbours.  Intervals go into a (S+1) x kBand difference array in shared
// memory; the column-wise running sum and the coverage LUT follow in the same kernel.
constexpr int kBand = 32;
constexpr int kBandThreads = 256;
constexpr int kBandSplit = 2;                    // threads per line
constexpr int kBandPart = kBand / kBandSplit;    // columns per thread
constexpr int kBandMaxS = 1536;                  // (S + 1) * kBand * 4 bytes of shared memory
__global__ void __launch_bounds__(kBandThreads) sphere_curves_band_kernel(
    const double* __restrict__ lines, const int32_t* __restrict__ offsets, int S, const int32_t* __restrict__ first,
    const double2* __restrict__ trig, double f, const uint8_t* __restrict__ lut, uint32_t* __restrict__ counts,
    uint8_t* __restrict__ img, int b0, int use_f32) {
    extern __shared__ __align__(16) int32_t band_diff[];            // (S + 1) * kBand, then the segment sums
    __shared__ int s_first[kBand + 1];
    __shared__ double2 s_trig[kBand + 1];
    __shared__ float2 s_trigf[kBand + 1];
    const int b = b0 + blockIdx.y, c0 = blockIdx.x * kBand, nc = min(kBand, S - c0);
    const int tid = threadIdx.x, lane = tid & 31;
    for (int e = tid; e < (S + 1) * kBand; e += kBandThreads) band_diff[e] = 0;
    if (tid <= nc) {
        const double2 t = trig[c0 + tid];
        s_first[tid] = first[c0 + tid]; s_trig[tid] = t; s_trigf[tid] = make_float2((float)t.x, (float)t.y);
    }
    __syncthreads();
    const float sop = (float)((double)S / kPi), hs = 0.5f * (float)S, fma_err = 1.1920929e-7f * (float)S;
    const double s = (double)S;
    const double half_over_s = __ddiv_rn(0.5, s);
    const double step = __ddiv_rn(kPi, (double)(kNumSamples - 1));
    const int n0 = offsets[b], n1 = offsets[b + 1];
    const int p0 = (tid % kBandSplit) * kBandPart, np = min(kBandPart, nc - p0);      // this thread's columns p0 .. p0 + np
    for (int line = n0 + tid / kBandSplit; np > 0 && line < n1; line += kBandThreads / kBandSplit) {
        const double l0 = __dmul_rn(lines[3 * (int64_t)line + 0], f);      // sphere_mapping.py:55-56
        const double l1 = __dmul_rn(lines[3 * (int64_t)line + 1], f);
        const double l2 = lines[3 * (int64_t)line + 2];
        const double astar = atan(l0 / l2);
        const int kstar = isnan(astar) ? -10 : (int)floor((astar + 0.5 * kPi) / step);
        const float neg_l0 = -(float)l0, l2f = (float)l2, inv_l1 = 1.0f / (float)l1, inv_l1_abs = fabsf(inv_l1);
        auto border_row = [&](int q) {
            int r;
            if (use_f32 && fast_row(s_trigf[q], neg_l0, l2f, inv_l1, inv_l1_abs, sop, hs, fma_err, S, r)) return r;
            return trig_row(s_trig[q], l0, l1, l2, half_over_s, s, S);
        };
        int j = p0 + (lane / kBandSplit) % np;
        int row_lo = border_row(j);
        for (int it = 0; it < np; ++it) {
            const int row_hi = border_row(j + 1);
            const int k0 = s_first[j], kn = s_first[j + 1];
            if (kn > k0) {
                const int k1 = min(kn, kNumSamples - 1);
                int rmin = S, rmax = -1;
                if (row_lo >= 0) { rmin = row_lo; rmax = row_lo; }
                if (row_hi >= 0) { rmin = min(rmin, row_hi); rmax = max(rmax, row_hi); }
#pragma unroll
                for (int e = -1; e <= 2; ++e) {
                    const int k = kstar + e;
                    if (k > k0 && k < k1) {
                        const int r = sample_row(k, step, l0, l1, l2, half_over_s, s, S);
                        if (r >= 0) { rmin = min(rmin, r); rmax = max(rmax, r); }
                    }
                }
                if (rmax >= rmin) {
                    atomicAdd(&band_diff[rmin * kBand + j], 1);
                    atomicAdd(&band_diff[(rmax + 1) * kBand + j], -1);
                }
            }
            if (++j == p0 + np) {
                j = p0;
                if (it + 1 < np) row_lo = border_row(p0);
            } else {
                row_lo = row_hi;
            }
        }
    }
    __syncthreads();
    // running sum down the columns: thread (seg, j) owns rows [seg*rows_per, ...) of column j
    constexpr int kSegs = kBandThreads / kBand;
    const int seg = tid / kBand, jc = tid % kBand;
    const int rows_per = (S + kSegs - 1) / kSegs;
    const int r0 = seg * rows_per, r1 = min(S, r0 + rows_per);
    int32_t* seg_sum = band_diff + (S + 1) * kBand;                  // kSegs * kBand
    int tot = 0;
    for (int r = r0; r < r1; ++r) tot += band_diff[r * kBand + jc];
    seg_sum[seg * kBand + jc] = tot;
    __syncthreads();
    if (jc >= nc) return;
    int run = 0;
    for (int q = 0; q < seg; ++q) run += seg_sum[q * kBand + jc];
    for (int r = r0; r < r1; ++r) {
        run += band_diff[r * kBand + jc];
        const int64_t o = ((int64_t)b * S + r) * S + c0 + jc;
        if (counts) counts[o] = (uint32_t)run;
        if (img) img[o] = lut[min(run, kLutMax)];
    }
}

// column-wise running sum of the difference array -> counts, LUT -> uint8
__global__ void curves_scan_kernel(const int32_t* __restrict__ diff, int B, int S, const uint8_t* __restrict__ lut,
                                   uint32_t* __restrict__ counts, uint8_t* __restrict__ img) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * S) return;
    int b = (int)(t / S), c = (int)(t % S);
    const int32_t* d = diff + (int64_t)b * (S + 1) * S;
    int run = 0;
    for (int r = 0; r < S; ++r) {
        run += d[(int64_t)r * S + c];
        int64_t o = ((int64_t)b * S + r) * S + c;
        if (counts) counts[o] = (uint32_t)run;
        if (img) img[o] = lut[min(run, kLutMax)];
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int host_angle_bin(double angle, int S) {
    volatile double a = angle / kPi;
    a = a + 0.5;
    a = a - 0.5 / S;
    a = a * S;
    double r = floor(a + 0.5);
    if (r < 0) r = 0;
    if (r > S - 1) r = S - 1;
    return (int)r;
}

int sphere_map_dev(vpk_ctx* ctx, const double* d_lines, const int32_t* d_offsets, const int32_t* h_offsets,
                   int32_t B, int32_t S, int32_t mode, double alpha, const double* d_weights,
                   uint32_t* d_hist, unsigned long long* d_whist, uint8_t* d_img) {
    const int64_t plane = (int64_t)S * S;
    const int64_t sumN = h_offsets[B] - h_offsets[0];      // offsets may be a window of a larger batch
    if (mode == VPK_SPHERE_VOTES) {
        // work list of upper-triangular tile pairs
        int64_t items = 0;
        for (int b = 0; b < B; ++b) {
            int64_t T = (h_offsets[b + 1] - h_offsets[b] + kTile - 1) / kTile;
            items += T * (T + 1) / 2;
        }
        const bool weighted = d_weights != nullptr;
        if (weighted) VPK_CUDA(cudaMemsetAsync(d_whist, 0, sizeof(unsigned long long) * plane * B, ctx->stream));
        else VPK_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) * plane * B, ctx->stream));
        if (items > 0) {
            VPK_TRY(ctx->d_work.ensure(items * sizeof(int4)));
            {
                KernelScope ks(ctx, "sphere_items");
                sphere_items_kernel<<<1, 1024, 0, ctx->stream>>>(d_offsets, B, ctx->d_work.as<int4>());
                VPK_TRY(check_launch("sphere_items"));
            }
            {
                KernelScope ks(ctx, "sphere_votes");
                sphere_votes_kernel<<<(unsigned)items, kVoteThreads, 0, ctx->stream>>>(
                    d_lines, d_offsets, ctx->d_work.as<int4>(), S, d_weights, d_hist, d_whist, nullptr);
                VPK_TRY(check_launch("sphere_votes"));
            }
        }
        return VPK_OK;
    }
    if (mode == VPK_SPHERE_CURVES) {
        // column -> first sample index table (exactly as numpy.linspace + the bin map) and the coverage LUT: built once per
        // (S, alpha) and kept on the device, so a call needs no host work and no synchronisation
        const size_t tab_bytes = (S + 1) * sizeof(int32_t) + kLutMax + 1;
        const size_t trig_off = (tab_bytes + 15) / 16 * 16;             // then sin / cos of the column borders (band form)
        if (ctx->curves_tab_S != S || ctx->curves_tab_alpha != alpha || !ctx->d_curves_tab.p) {
            VPK_TRY(ctx->h_stage.ensure(tab_bytes));
            VPK_TRY(ctx->d_curves_tab.ensure(trig_off + (S + 1) * sizeof(double2)));
            int32_t* first = ctx->h_stage.as<int32_t>();
            uint8_t* lut = reinterpret_cast<uint8_t*>(first + S + 1);
            const double start = -0.5 * kPi, stop = 0.5 * kPi;
            volatile double step = (stop - start) / (double)(kNumSamples - 1);
            int c = 0;
            first[0] = 0;
            for (int k = 0; k < kNumSamples; ++k) {
                volatile double a = (double)k * step;
                a = a + start;
                if (k == kNumSamples - 1) a = stop;
                int col = host_angle_bin(a, S);
                while (c < col) first[++c] = k;
            }
            while (c < S) first[++c] = kNumSamples;
            for (int k = 0; k <= kLutMax; ++k)
                lut[k] = (uint8_t)floor(255.0 * (1.0 - pow(1.0 - alpha, (double)k)));
            VPK_CUDA(cudaMemcpyAsync(ctx->d_curves_tab.p, first, tab_bytes, cudaMemcpyHostToDevice, ctx->stream));
            curves_trig_kernel<<<(unsigned)(S / 128 + 1), 128, 0, ctx->stream>>>(
                ctx->d_curves_tab.as<int32_t>(), S, reinterpret_cast<double2*>(static_cast<char*>(ctx->d_curves_tab.p) + trig_off));
            VPK_TRY(check_launch("curves_trig"));
            VPK_CUDA(cudaStreamSynchronize(ctx->stream));          // once: the pinned staging buffer is reused by other calls
            ctx->curves_tab_S = S; ctx->curves_tab_alpha = alpha;
        }
        const int32_t* d_first = ctx->d_curves_tab.as<int32_t>();
        const uint8_t* d_lut = reinterpret_cast<const uint8_t*>(d_first + S + 1);
        if (S <= kBandMaxS && !getenv("VPK_CURVES_GLOBAL_DIFF")) {
            const size_t smem = sizeof(int32_t) * ((size_t)(S + 1) * kBand + kBandThreads);
            if (smem > ctx->curves_band_smem) {
                VPK_CUDA(cudaFuncSetAttribute(sphere_curves_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                ctx->curves_band_smem = smem;
            }
            for (int b0 = 0; b0 < B; b0 += 65535) {
                KernelScope ks(ctx, "sphere_curves");
                dim3 grid((unsigned)((S + kBand - 1) / kBand), (unsigned)std::min(B - b0, 65535));
                sphere_curves_band_kernel<<<grid, kBandThreads, smem, ctx->stream>>>(
                    d_lines, d_offsets, S, d_first, reinterpret_cast<const double2*>(reinterpret_cast<const char*>(d_first) + trig_off),
                    1.0, d_lut, d_hist, d_img, b0, getenv("VPK_CURVES_F64") ? 0 : 1);
                VPK_TRY(check_launch("sphere_curves"));
            }
            return VPK_OK;
        }
        // very large grids: difference array in HBM, one CTA per line, separate scan
        size_t diff_bytes = sizeof(int32_t) * (size_t)(S + 1) * S * B;
        VPK_TRY(ctx->d_misc.ensure(diff_bytes));
        VPK_CUDA(cudaMemsetAsync(ctx->d_misc.p, 0, diff_bytes, ctx->stream));
        if (sumN > 0) {
            KernelScope ks(ctx, "sphere_curves");
            int threads = S >= 512 ? 512 : ((S + 31) / 32) * 32;
            sphere_curves_kernel<<<(unsigned)sumN, threads, 0, ctx->stream>>>(d_lines, d_offsets, B, S, d_first, 1.0,
                                                                               ctx->d_misc.as<int32_t>(), (int64_t)h_offsets[0]);
            VPK_TRY(check_launch("sphere_curves"));
        }
        {
            KernelScope ks(ctx, "curves_scan");
            int64_t t = (int64_t)B * S;
            curves_scan_kernel<<<(unsigned)((t + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_misc.as<int32_t>(), B, S, d_lut,
                                                                                    d_hist, d_img);
            VPK_TRY(check_launch("curves_scan"));
        }
        return VPK_OK;
    }
    set_error("sphere_map: unknown mode %d", mode);
    return VPK_ERR_ARG;
}

// votes -> uint8 image (and optional float export of the Q.16 sums)
int sphere_votes_finish_dev(vpk_ctx* ctx, int32_t B, int32_t S, const uint32_t* d_hist,
                            const unsigned long long* d_whist, uint8_t* d_img, float* d_wout) {
    const int64_t plane = (int64_t)S * S;
    VPK_TRY(ctx->d_weights.ensure(sizeof(unsigned long long) * (size_t)B));
    unsigned long long* d_max = ctx->d_weights.as<unsigned long long>();
    VPK_CUDA(cudaMemsetAsync(d_max, 0, sizeof(unsigned long long) * B, ctx->stream));
    int64_t gx = (plane + 1023) / 1024;
    dim3 grid((unsigned)(gx < 64 ? gx : 64), (unsigned)B);
    {
        KernelScope ks(ctx, "plane_max");
        if (d_whist) plane_max_kernel<unsigned long long><<<grid, 256, 0, ctx->stream>>>(d_whist, plane, d_max);
        else plane_max_kernel<uint32_t><<<grid, 256, 0, ctx->stream>>>(d_hist, plane, d_max);
        VPK_TRY(check_launch("plane_max"));
    }
    if (d_img || d_wout) {
        KernelScope ks(ctx, "votes_image");
        if (d_whist) votes_image_kernel<unsigned long long><<<grid, 256, 0, ctx->stream>>>(d_whist, plane, d_max, d_img, d_wout);
        else votes_image_kernel<uint32_t><<<grid, 256, 0, ctx->stream>>>(d_hist, plane, d_max, d_img, nullptr);
        VPK_TRY(check_launch("votes_image"));
    }
    return VPK_OK;
}

}  // namespace vpk

using namespace vpk;

extern "C" {

int vpk_lines_from_segments(vpk_ctx* ctx, const double* segments, int64_t n, double* lines_out) {
    if (!ctx || (n > 0 && (!segments || !lines_out)) || n < 0) { set_error("vpk_lines_from_segments: bad argument"); return VPK_ERR_ARG; }
    if (n == 0) return VPK_OK;
    VPK_CUDA(cudaSetDevice(ctx->device));
    VPK_TRY(ctx->d_segments.ensure(n * 4 * sizeof(double)));
    VPK_TRY(ctx->d_lines.ensure(n * 3 * sizeof(double)));
    VPK_CUDA(cudaMemcpyAsync(ctx->d_segments.p, segments, n * 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    VPK_TRY(lines_from_segments_dev(ctx, ctx->d_segments.as<double>(), n, ctx->d_lines.as<double>()));
    VPK_CUDA(cudaMemcpyAsync(lines_out, ctx->d_lines.p, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    VPK_CUDA(cudaStreamSynchronize(ctx->stream));
    return VPK_OK;
}

int vpk_segments_from_lsd(vpk_ctx* ctx, const double* lsd, int32_t ncols, const int32_t* offsets, const int32_t* widths,
                          const int32_t* heights, int32_t B, double* segments_out, double* lines_out, double* nfa_out) {
    if (!ctx || !offsets || !widths || !heights || B < 0 || ncols < 4 || !segments_out) { set_error("vpk_segments_from_lsd: bad argument"); return VPK_ERR_ARG; }
    if (B == 0) return VPK_OK;
    if (offsets[0] != 0) { set_error("vpk_segments_from_lsd: offsets[0] must be 0"); return VPK_ERR_ARG; }
    for (int b = 0; b < B; ++b)
        if (offsets[b + 1] < offsets[b] || widths[b] <= 0 || heights[b] <= 0) { set_error("vpk_segments_from_lsd: bad offsets or image size"); return VPK_ERR_ARG; }
    const int64_t n = offsets[B];
    if (n == 0) return VPK_OK;
    if (!lsd) { set_error("vpk_segments_from_lsd: lsd is NULL"); return VPK_ERR_ARG; }
    if (nfa_out && ncols < 7) { set_error("vpk_segments_from_lsd: the NFA column needs 7-column LSD rows"); return VPK_ERR_ARG; }
    VPK_CUDA(cudaSetDevice(ctx->device));
    VPK_TRY(ctx->d_misc.ensure((size_t)n * ncols * sizeof(double)));
    VPK_TRY(ctx->d_segments.ensure((n + 1) * 4 * sizeof(double)));
    VPK_TRY(ctx->d_lines.ensure((n + 1) * 3 * sizeof(double)));
    VPK_TRY(ctx->d_offsets.ensure((size_t)(3 * (B + 1)) * sizeof(int32_t)));
    VPK_TRY(ctx->d_weights.ensure((size_t)(n + 1) * sizeof(double)));
    int32_t* d_off = ctx->d_offsets.as<int32_t>();
    int32_t* d_w = d_off + (B + 1);
    int32_t* d_h = d_w + (B + 1);
    cudaStream_t st = ctx->stream;
    VPK_CUDA(cudaMemcpyAsync(ctx->d_misc.p, lsd, (size_t)n * ncols * sizeof(double), cudaMemcpyHostToDevice, st));
    VPK_CUDA(cudaMemcpyAsync(d_off, offsets, (B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    VPK_CUDA(cudaMemcpyAsync(d_w, widths, B * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    VPK_CUDA(cudaMemcpyAsync(d_h, heights, B * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    VPK_TRY(segments_from_lsd_dev(ctx, ctx->d_misc.as<double>(), ncols, d_off, d_w, d_h, B, n, ctx->d_segments.as<double>(),
                                  nfa_out ? ctx->d_weights.as<double>() : nullptr));
    VPK_CUDA(cudaMemcpyAsync(segments_out, ctx->d_segments.p, n * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (lines_out) {
        VPK_TRY(lines_from_segments_dev(ctx, ctx->d_segments.as<double>(), n, ctx->d_lines.as<double>()));
        VPK_CUDA(cudaMemcpyAsync(lines_out, ctx->d_lines.p, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if (nfa_out) VPK_CUDA(cudaMemcpyAsync(nfa_out, ctx->d_weights.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    VPK_CUDA(cudaStreamSynchronize(st));
    return VPK_OK;
}

int vpk_sphere_map(vpk_ctx* ctx, const double* lines, const int32_t* offsets, int32_t B, int32_t S, int32_t mode,
                   double alpha, const double* weights, uint32_t* hist_out, float* whist_out, uint8_t* image_out) {
    if (!ctx || !offsets || B < 0 || S <= 0 || S > 4096) { set_error("vpk_sphere_map: bad argument"); return VPK_ERR_ARG; }
    if (B == 0) return VPK_OK;
    for (int b = 0; b < B; ++b)
        if (offsets[b + 1] < offsets[b] || offsets[0] != 0) { set_error("vpk_sphere_map: offsets must start at 0 and be non-decreasing"); return VPK_ERR_ARG; }
    const int64_t sumN = offsets[B];
    if (sumN > 0 && !lines) { set_error("vpk_sphere_map: lines is NULL"); return VPK_ERR_ARG; }
    if (mode == VPK_SPHERE_CURVES && weights) { set_error("vpk_sphere_map: weights apply to the votes mode only"); return VPK_ERR_ARG; }
    const bool weighted = weights != nullptr;
    if (hist_out && weighted) { set_error("vpk_sphere_map: hist_out is the unweighted output; use whist_out with weights"); return VPK_ERR_ARG; }
    if (whist_out && !weighted) { set_error("vpk_sphere_map: whist_out needs weights"); return VPK_ERR_ARG; }
    VPK_CUDA(cudaSetDevice(ctx->device));
    const int64_t plane = (int64_t)S * S;
    const size_t img_bytes = ((size_t)(plane * B + 15) / 16) * 16;
    VPK_TRY(ctx->d_lines.ensure((sumN + 1) * 3 * sizeof(double)));
    VPK_TRY(ctx->d_offsets.ensure((B + 1) * sizeof(int32_t)));
    VPK_TRY(ctx->d_hist.ensure((weighted ? sizeof(unsigned long long) : sizeof(uint32_t)) * plane * B));
    VPK_TRY(ctx->d_img.ensure(img_bytes + (whist_out ? sizeof(float) * plane * B : 0)));
    if (sumN) VPK_CUDA(cudaMemcpyAsync(ctx->d_lines.p, lines, sumN * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    VPK_CUDA(cudaMemcpyAsync(ctx->d_offsets.p, offsets, (B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    double* d_w = nullptr;
    if (weighted) {
        VPK_TRY(ctx->d_segments.ensure((sumN + 1) * sizeof(double)));
        d_w = ctx->d_segments.as<double>();
        VPK_CUDA(cudaMemcpyAsync(d_w, weights, sumN * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    uint32_t* d_hist = weighted ? nullptr : ctx->d_hist.as<uint32_t>();
    unsigned long long* d_whist = weighted ? ctx->d_hist.as<unsigned long long>() : nullptr;
    uint8_t* d_img = ctx->d_img.as<uint8_t>();
    float* d_wout = whist_out ? reinterpret_cast<float*>(d_img + img_bytes) : nullptr;
    VPK_TRY(sphere_map_dev(ctx, ctx->d_lines.as<double>(), ctx->d_offsets.as<int32_t>(), offsets, B, S, mode, alpha, d_w,
                           d_hist, d_whist, d_img));
    if (mode == VPK_SPHERE_VOTES && (image_out || whist_out))
        VPK_TRY(sphere_votes_finish_dev(ctx, B, S, d_hist, d_whist, image_out ? d_img : nullptr, d_wout));
    if (hist_out) VPK_CUDA(cudaMemcpyAsync(hist_out, d_hist, sizeof(uint32_t) * plane * B, cudaMemcpyDeviceToHost, ctx->stream));
    if (whist_out) VPK_CUDA(cudaMemcpyAsync(whist_out, d_wout, sizeof(float) * plane * B, cudaMemcpyDeviceToHost, ctx->stream));
    if (image_out) VPK_CUDA(cudaMemcpyAsync(image_out, d_img, plane * B, cudaMemcpyDeviceToHost, ctx->stream));
    VPK_CUDA(cudaStreamSynchronize(ctx->stream));
    return VPK_OK;
}

}  // extern "C"
