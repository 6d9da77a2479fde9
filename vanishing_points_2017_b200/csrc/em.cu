// Stage 3: EM vanishing-point localisation as ONE persistent kernel.
//
// Replaces vp_localisation.expectation_maximisation (reference
// vp_localisation.py:168-450) and everything it calls in
// probability_functions.py / coordinate_conversion.py.  One CTA owns one
// image at a time (work queue ordered by descending N); the VP state (<= 64
// hypotheses: directions, variances, priors, mixture parameters) lives in
// shared memory, the per-line arrays and the N x N segment-similarity matrix
// live in a per-CTA slice of an HBM workspace that stays L2-resident while the
// image is being processed.  All control flow of the reference (pruning,
// periodic split / merge, final hard-assignment refit) runs on the device.
//
// Arithmetic is float64 throughout: the reference is float64, its discrete
// decisions (counts < 3, argmax, err > 1.5, angle < thresh) sit on float values
// and the parity gate is 1e-4 rad on the refined VPs.  B200 keeps 64 FP64
// FMA/clk/SM, so the dominant (M x N)(N x N) weight-matrix product is bound by
// streaming the similarity matrix (8 N^2 bytes per product), not by the FMAs.
//
// SVD(diag(w) l) of calc_new_vanishing_point (vp_localisation.py:453-479) is
// replaced by the smallest eigenvector of the 3x3 matrix sum w^2 l l^T (Jacobi).
#include <math.h>
#include <algorithm>
#include "vpk_internal.cuh"

namespace vpk {

static constexpr int kEmThreads = 256;
static constexpr int kEmWarps = kEmThreads / 32;
static constexpr int kMaxM = VPK_MAX_VP;
static constexpr int kMaxComp = 100;       // probability_functions.py:87
static constexpr int kCells = VPK_GRID * VPK_GRID;
static constexpr int kJT = 64;             // j-tile of the weight-matrix product
static constexpr int kMCH = 16;            // VP rows per register block
static constexpr double kPi = 3.141592653589793;

struct EmParams {
    const double* lines;
    const double* segs;
    const int32_t* offsets;
    int B;
    const float* resp32;
    const double* resp64;
    const uint8_t* sphere;
    int S;
    const double* init_vp;
    const int32_t* init_off;
    vpk_em_config cfg;
    const int32_t* order;
    int* queue;
    double* ws;
    size_t ws_stride;      // doubles per CTA slot
    int nmax;
    double* overflow;      // shared clustering scratch for very large splits
    size_t overflow_cap;   // doubles
    int* overflow_lock;
    unsigned long long* phase_cycles;   // nullable: per-phase SM cycles summed over CTAs (profiling)
    EmDeviceOut out;
};

enum { PH_SETUP = 0, PH_PAIR, PH_RATING, PH_INIT, PH_ESTEP, PH_WMAT, PH_MSTEP, PH_MERGE, PH_SPLIT, PH_COUNTS, PH_OTHER, PH_N };
static const char* const kPhaseNames[PH_N] = {"em:setup", "em:pair_pass", "em:line_rating", "em:init", "em:estep", "em:wmat",
                                             "em:mstep", "em:merge", "em:split", "em:counts", "em:other"};

// per-image view of the slot workspace
struct Img {
    int N;
    const double* lp;      // (N,4) segments (input)
    double* ln;            // (N,3) unit lines
    double* lsim;          // (N,N)
    double* lweight;       // (N)
    double* colsum;        // (N)
    double* langle;        // (N)
    double* lvsq;          // (M,N)   -- lvsq|pvl|w are contiguous: split scratch
    double* pvl;           // (M,N)
    double* w;             // (M,N)
    int* assoc;            // (N)
    size_t scratch_cap;    // doubles available from lvsq on
};

struct EmShared {
    double cur[kMaxM][3], nxt[kMaxM][3], s[kMaxM], pv[kMaxM];
    double vx[kMaxM], vy[kMaxM], inv2s[kMaxM], coef[kMaxM];
    double nv[kMaxM][3], ns[kMaxM], err[kMaxM], cw[kMaxM], rowmax[kMaxM];
    int cnt[kMaxM], ok[kMaxM], rem[kMaxM];
    double pdf_a[kMaxComp], pdf_b[kMaxComp], pdf_w[kMaxComp];
    double resp[kCells];
    double tile[kJT * kMCH];
    double redv[kEmThreads];
    int redi[kEmThreads], redj[kEmThreads];
    double ang[kMaxM];       // generic per-VP scratch
    double sigma_prior;
    int M, npdf, img, flag, ia, ib;
    double da;
    unsigned long long phase[PH_N];   // profiling: SM cycles per phase of this CTA
    long long t_last;
    int timing;
};

// attribute the cycles since the previous lap to phase `ph` (thread 0 only, profiling runs only)
__device__ __forceinline__ void phase_lap(EmShared& sh, int ph) {
    if (sh.timing && threadIdx.x == 0) {
        long long t = clock64();
        sh.phase[ph] += (unsigned long long)(t - sh.t_last);
        sh.t_last = t;
    }
}

// ---------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max_nanprop(double v) {
    // max that propagates NaN like numpy.max
    for (int o = 16; o > 0; o >>= 1) {
        double t = __shfl_xor_sync(0xffffffffu, v, o);
        v = (isnan(v) || isnan(t)) ? nan("") : (t > v ? t : v);
    }
    return v;
}

// eigenvector of the smallest eigenvalue of the symmetric 3x3 matrix
// [g0 g1 g2; g1 g3 g4; g2 g4 g5] (cyclic Jacobi, float64).  false if not finite.
__device__ bool smallest_eigvec3(const double g[6], double out[3]) {
    double tr = g[0] + g[3] + g[5];
    if (!(tr > 0.0) || isinf(tr)) return false;
    double sc = 1.0 / tr;
    double a[3][3] = {{g[0] * sc, g[1] * sc, g[2] * sc}, {g[1] * sc, g[3] * sc, g[4] * sc}, {g[2] * sc, g[4] * sc, g[5] * sc}};
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            if (isnan(a[i][j])) return false;
    for (int sweep = 0; sweep < 24; ++sweep) {
        double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        if (off < 1e-40) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            double apq = a[p][q];
            if (apq == 0.0) continue;
            double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
            double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            if (isinf(theta)) t = 0.0;
            double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
            double app = a[p][p], aqq = a[q][q];
            a[p][p] = app - t * apq;
            a[q][q] = aqq + t * apq;
            a[p][q] = a[q][p] = 0.0;
            const int r = 3 - p - q;
            double arp = a[r][p], arq = a[r][q];
            a[r][p] = a[p][r] = c * arp - sn * arq;
            a[r][q] = a[q][r] = sn * arp + c * arq;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double vkp = v[k][p], vkq = v[k][q];
                v[k][p] = c * vkp - sn * vkq;
                v[k][q] = sn * vkp + c * vkq;
            }
        }
    }
    int m = 0;
    if (a[1][1] < a[m][m]) m = 1;
    if (a[2][2] < a[m][m]) m = 2;
    double x = v[0][m], y = v[1][m], z = v[2][m];
    double n = sqrt(x * x + y * y + z * z);
    if (!(n > 0.0)) return false;
    out[0] = x / n; out[1] = y / n; out[2] = z / n;
    return true;
}

__device__ __forceinline__ double sign_np(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }   // numpy.sign

// ---- segment-pair geometry (vp_localisation.py:700-762) ---------------------
struct Seg { double x1, y1, x2, y2; };
__device__ __forceinline__ Seg load_seg(const double* lp, int n) {
    const double2* p = reinterpret_cast<const double2*>(lp) + 2 * (size_t)n;
    double2 a = p[0], b = p[1];
    return {a.x, a.y, b.x, b.y};
}
// line_segment_point_distance (:743-758): note the squared *norm* of :747
__device__ __forceinline__ double psd(const Seg& s, double px, double py) {
    double dx = s.x2 - s.x1, dy = s.y2 - s.y1;
    double nrm = sqrt(dx * dx + dy * dy);
    double param = ((px - s.x1) * dx + (py - s.y1) * dy) / (nrm * nrm);
    double cx, cy;
    if (param < 0) { cx = s.x1; cy = s.y1; }
    else if (param > 1) { cx = s.x2; cy = s.y2; }
    else { cx = s.x1 + param * dx; cy = s.y1 + param * dy; }
    double ex = cx - px, ey = cy - py;
    return sqrt(ex * ex + ey * ey);
}
// line_distance_closest (:727-740)
__device__ __forceinline__ double seg_distance(const Seg& a, const Seg& b) {
    double d1 = psd(a, b.x1, b.y1), d2 = psd(a, b.x2, b.y2), d4 = psd(b, a.x1, a.y1), d5 = psd(b, a.x2, a.y2);
    return fmin(fmin(d1, d2), fmin(d4, d5));
}
__device__ __forceinline__ double seg_len(const Seg& a) {
    double dx = a.x1 - a.x2, dy = a.y1 - a.y2;
    return sqrt(dx * dx + dy * dy);
}
// lines_points_cosangle (:715-724)
__device__ __forceinline__ double cosangle(const Seg& a, const Seg& b, double f) {
    double v1x = a.x1 - a.x2, v1y = a.y1 - a.y2, v2x = b.x1 - b.x2, v2y = b.y1 - b.y2;
    double c = fabs((v1x * v2x + v1y * v2y) / (sqrt(v1x * v1x + v1y * v1y) * sqrt(v2x * v2x + v2y * v2y)));
    double dphi = fabs(acos(fmin(fmax(c, -1.0), 1.0)));
    if (isnan(c)) dphi = c;
    return cos(fmin(fmax(f * dphi, -0.5 * kPi), 0.5 * kPi));
}
// lines_proximity (:708-712), sigma = 1
__device__ __forceinline__ double proximity(const Seg& a, const Seg& b, double d) {
    double sg = fmin(seg_len(a), seg_len(b));
    return exp(-(d * d) / (2 * sg * sg));
}

// ---------------------------------------------------------------------------
// E3: lsim = cos9 * prox, symmetric, zero diagonal; column sums
// ---------------------------------------------------------------------------
__device__ void pair_pass(const Img& im, EmShared& sh) {
    phase_lap(sh, PH_OTHER);
    const int N = im.N;
    const int T = (N + 31) / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // upper-triangular 32x32 tiles, one warp per tile
    for (int t = warp; t < T * (T + 1) / 2; t += kEmWarps) {
        // decode (ti <= tj) from linear index
        int ti = 0, rem = t;
        while (rem >= T - ti) { rem -= T - ti; ++ti; }
        int tj = ti + rem;
        int j = tj * 32 + lane;
        Seg sj = load_seg(im.lp, j < N ? j : 0);
        for (int ii = 0; ii < 32; ++ii) {
            int i = ti * 32 + ii;
            if (i >= N) break;
            if (j >= N || (ti == tj && j < i)) continue;
            double val = 0.0;
            if (i != j) {
                Seg si = load_seg(im.lp, i);
                double d = seg_distance(si, sj);
                val = cosangle(si, sj, 9.0) * proximity(si, sj, d);
            }
            im.lsim[(size_t)i * N + j] = val;
            im.lsim[(size_t)j * N + i] = val;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += kEmThreads) {
        double acc = 0.0;
        for (int j = 0; j < N; ++j) acc += im.lsim[(size_t)j * N + k];
        im.colsum[k] = acc;
    }
    __syncthreads();
    phase_lap(sh, PH_PAIR);
}

// ---------------------------------------------------------------------------
// E4: kNN line rating (vp_localisation.py:34-84), one warp per line
// ---------------------------------------------------------------------------
__device__ void line_rating(const Img& im, EmShared& sh, bool use_weights) {
    phase_lap(sh, PH_OTHER);
    const int N = im.N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k1 = min(10, N), k2 = min(4, N);
    // per-warp candidate list in shared memory (reuses the GEMM tile)
    double* kd = sh.tile + warp * 32;          // sort distance (diag = 4)
    int* kj = reinterpret_cast<int*>(sh.tile + kEmWarps * 32) + warp * 32;
    for (int i = warp; i < N; i += kEmWarps) {
        Seg si = load_seg(im.lp, i);
        double llen = seg_len(si);
        if (!use_weights) { if (lane == 0) im.lweight[i] = 1.0; continue; }
        int cnt = 0;
        double thr = INFINITY;
        for (int j0 = 0; j0 < N; j0 += 32) {
            int j = j0 + lane;
            double d = INFINITY;
            if (j < N) d = (j == i) ? 4.0 : seg_distance(si, load_seg(im.lp, j));
            if (isnan(d)) d = INFINITY;
            unsigned mask = __ballot_sync(0xffffffffu, cnt < k1 ? (j < N) : (d < thr));
            while (mask) {
                int src = __ffs(mask) - 1;
                mask &= mask - 1;
                double dd = __shfl_sync(0xffffffffu, d, src);
                int jj = j0 + src;
                if (cnt >= k1 && !(dd < thr)) continue;
                if (lane == 0) {
                    int pos = cnt < k1 ? cnt : k1 - 1;
                    while (pos > 0 && kd[pos - 1] > dd) { kd[pos] = kd[pos - 1]; kj[pos] = kj[pos - 1]; --pos; }
                    kd[pos] = dd; kj[pos] = jj;
                }
                if (cnt < k1) ++cnt;
                __syncwarp();
                if (cnt >= k1) thr = kd[k1 - 1];
            }
        }
        __syncwarp();
        // cos9 to each neighbour, then the k2 largest (descending), sum prox*cos
        double c = -INFINITY, px = 0.0;
        if (lane < cnt) {
            int j = kj[lane];
            Seg sj = load_seg(im.lp, j);
            c = cosangle(si, sj, 9.0);
            double dtrue = (j == i) ? seg_distance(si, sj) : kd[lane];
            px = proximity(si, sj, dtrue);
            if (isnan(c)) c = -INFINITY;
        }
        double score = 0.0;
        for (int r = 0; r < k2; ++r) {
            double best = c;
            int who = lane;
            for (int o = 16; o > 0; o >>= 1) {
                double ob = __shfl_xor_sync(0xffffffffu, best, o);
                int ow = __shfl_xor_sync(0xffffffffu, who, o);
                if (ob > best || (ob == best && ow > who)) { best = ob; who = ow; }
            }
            double term = __shfl_sync(0xffffffffu, px * c, who);
            score += term;
            if (lane == who) c = -INFINITY;
        }
        score /= (double)k2;
        if (lane == 0) {
            double ls = fmin(fmax(score, 0.2), 1.0);      // :231
            if (isnan(score)) ls = score;
            im.lweight[i] = llen * ls;                     // :232-233
        }
        __syncwarp();
    }
    __syncthreads();
    phase_lap(sh, PH_RATING);
}

// ---------------------------------------------------------------------------
// E0/E1/E2: initial VPs and the prior mixture
// ---------------------------------------------------------------------------
__device__ void init_prior_and_vps(const EmParams& P, int b, EmShared& sh, bool have_init) {
    const int tid = threadIdx.x;
    const int G = VPK_GRID, S = P.S;
    for (int c = tid; c < kCells; c += kEmThreads)
        sh.resp[c] = P.resp64 ? P.resp64[(size_t)b * kCells + c] : (double)P.resp32[(size_t)b * kCells + c];
    __syncthreads();
    // --- E2 pdf_params (probability_functions.py:62-96): top-100 cells
    double* rank_w = sh.redv;          // reuse
    __shared__ double s_sum;
    __shared__ int s_keep[kCells];
    for (int c = tid; c < kCells; c += kEmThreads) {
        double v = sh.resp[c];
        int rank = 0;
        for (int o = 0; o < kCells; ++o) {
            double u = sh.resp[o];
            rank += (u > v) || (u == v && o > c);
        }
        s_keep[c] = rank < kMaxComp;
    }
    __syncthreads();
    if (tid == 0) {
        double sum = 0.0;
        for (int c = 0; c < kCells; ++c) if (s_keep[c]) sum += sh.resp[c];
        s_sum = sum;
        double sigma = kPi / (1.282 * G);
        sh.sigma_prior = sigma;
        int n = 0;
        for (int c = 0; c < kCells; ++c) {
            if (!s_keep[c]) continue;
            double w = sh.resp[c] / sum / (2 * kPi * sigma * sigma);
            if (!(w > 0)) continue;                    // calc_pdf skips weights <= 0 (:21)
            int a = c % G, bb = c / G;
            // numpy.linspace(-(G-1)/G*pi/2, (G-1)/G*pi/2, G)
            double lo = -(G - 1.0) / G * kPi / 2, hi = (G - 1.0) / G * kPi / 2, st = (hi - lo) / (G - 1);
            sh.pdf_a[n] = a == G - 1 ? hi : a * st + lo;
            sh.pdf_b[n] = bb == G - 1 ? hi : bb * st + lo;
            sh.pdf_w[n] = w;
            ++n;
        }
        sh.npdf = n;
    }
    (void)rank_w;
    __syncthreads();
    if (have_init) return;
    // --- E0 find_maxima (vp_localisation.py:13-31, border quirk included)
    __shared__ int s_max[kCells];
    for (int c = tid; c < kCells; c += kEmThreads) {
        int a = c % G, bb = c / G;
        double vm = sh.resp[c];
        double vu = a + 1 < G ? sh.resp[bb * G + a + 1] : 0.0;
        double vd = a - 1 > 0 ? sh.resp[bb * G + a - 1] : 0.0;
        double vl = bb - 1 > 0 ? sh.resp[(bb - 1) * G + a] : 0.0;
        double vr = bb + 1 < G ? sh.resp[(bb + 1) * G + a] : 0.0;
        s_max[c] = (vm > vu && vm > vd && vm > vl && vm > vr) ? 1 : 0;
    }
    __syncthreads();
    // keep the num_init_vp strongest maxima (:121-126)
    for (int c = tid; c < kCells; c += kEmThreads) {
        if (!s_max[c]) { s_keep[c] = 0; continue; }
        double v = sh.resp[c];
        int rank = 0;
        for (int o = 0; o < kCells; ++o)
            if (s_max[o]) { double u = sh.resp[o]; rank += (u > v) || (u == v && o > c); }
        s_keep[c] = rank < P.cfg.num_init_vp;
    }
    __syncthreads();
    // --- E1: per kept cell, mean index of the brightest pixels of the flipped sphere slice
    __shared__ double s_cand[kCells][3];
    __shared__ int s_has[kCells];
    const uint8_t* img = P.sphere + (size_t)b * S * S;
    const int warp = tid >> 5, lane = tid & 31;
    for (int c = warp; c < kCells; c += kEmWarps) {
        if (!s_keep[c]) { if (lane == 0) s_has[c] = 0; continue; }
        int ra = c / G, rb = c % G;      // ra: row of the response (beta), rb: column (alpha)
        int r0 = ra * S / G, r1 = (ra + 1) * S / G, c0 = rb * S / G, c1 = (rb + 1) * S / G;
        int w = c1 - c0, npx = (r1 - r0) * w;
        int mx = 0;
        for (int e = lane; e < npx; e += 32) {
            int r = r0 + e / w, cc = c0 + e % w;
            mx = max(mx, (int)img[(size_t)(S - 1 - r) * S + cc]);      // flipped vertically (:114)
        }
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        long long sr = 0, sc = 0;
        int cnt = 0;
        if (mx > 0)
            for (int e = lane; e < npx; e += 32) {
                int r = e / w, cc = e % w;
                if ((int)img[(size_t)(S - 1 - (r0 + r)) * S + c0 + cc] == mx) { sr += r; sc += cc; ++cnt; }
            }
        for (int o = 16; o > 0; o >>= 1) {
            sr += __shfl_xor_sync(0xffffffffu, sr, o);
            sc += __shfl_xor_sync(0xffffffffu, sc, o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if (lane == 0) {
            s_has[c] = cnt > 0;
            if (cnt > 0) {
                double idx0 = (double)sc / (double)cnt + c0;     // :158
                double idx1 = (double)sr / (double)cnt + r0;     // :157
                double alpha = (idx0 - 0.5 * S + 0.5) * kPi / S; // index_to_angle
                double beta = (idx1 - 0.5 * S + 0.5) * kPi / S;
                double x = sin(alpha) * cos(beta), y = sin(beta), z = cos(alpha) * cos(beta);
                double sg = sign_np(z);                          // angle_to_point :48
                s_cand[c][0] = x * sg; s_cand[c][1] = y * sg; s_cand[c][2] = z * sg;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        int M = 0;
        for (int c = 0; c < kCells && M < kMaxM; ++c)
            if (s_keep[c] && s_has[c]) {
                sh.cur[M][0] = s_cand[c][0]; sh.cur[M][1] = s_cand[c][1]; sh.cur[M][2] = s_cand[c][2];
                ++M;
            }
        sh.M = M;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// E5: E-step (probability_functions.py:99-147).  v = sh.cur or sh.nxt.
// ---------------------------------------------------------------------------
__device__ void estep(const Img& im, EmShared& sh, const double (*v)[3]) {
    phase_lap(sh, PH_OTHER);
    const int M = sh.M, N = im.N, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // prior at the VP angles: calc_angles (:252-259) + calc_pdf (:8-40)
    for (int m = warp; m < M; m += kEmWarps) {
        double beta = asin(v[m][1]);
        double inner = v[m][0] / cos(beta);
        inner = fmax(fmin(inner, 1.0), -1.0);
        if (isnan(v[m][0] / cos(beta))) inner = nan("");
        double x = asin(inner), y = beta;
        const double k = -0.5 / (sh.sigma_prior * sh.sigma_prior);
        double acc = 0.0;
        for (int n = lane; n < sh.npdf; n += 32) {
            double mx = sh.pdf_a[n], my = sh.pdf_b[n];
            double d1 = (x - mx) * (x - mx) + (y - my) * (y - my);
            double d2 = (x - mx + kPi) * (x - mx + kPi) + (y + my) * (y + my);
            double d3 = (x - mx - kPi) * (x - mx - kPi) + (y + my) * (y + my);
            double d4 = (x + mx) * (x + mx) + (y - my - kPi) * (y - my - kPi);
            double p = exp(d1 * k) + exp(d2 * k) + exp(d3 * k) + exp(d4 * k) + exp(d4 * k);   // 4th == 5th (:25-26)
            acc += p * sh.pdf_w[n];
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            sh.pv[m] = acc;
            sh.vx[m] = v[m][0] / v[m][2];
            sh.vy[m] = v[m][1] / v[m][2];
            double sm = sh.s[m] > 1e-200 ? sh.s[m] : 1e-200;    // calc_plv mutates s (:139)
            if (isnan(sh.s[m])) sm = 1e-200;
            sh.s[m] = sm;
            sh.inv2s[m] = 2.0 * sm;
            sh.coef[m] = 1.0 / sqrt(2.0 * kPi * sm);
        }
    }
    __syncthreads();
    for (int n = tid; n < N; n += kEmThreads) {
        Seg sg = load_seg(im.lp, n);
        double mx = 0.5 * (sg.x1 + sg.x2), my = 0.5 * (sg.y1 + sg.y2);
        double bx = sg.x1 - sg.x2, by = sg.y1 - sg.y2;
        double nb = sqrt(bx * bx + by * by);
        double pl = 0.0;
        for (int m = 0; m < M; ++m) {
            double ax = mx - sh.vx[m], ay = my - sh.vy[m];
            double c = (ax * bx + ay * by) / (sqrt(ax * ax + ay * ay) * nb);
            double q = 1.0 - fabs(c);
            double lvsq = q * q;                                            // calc_lvsq_angle (:174)
            double plv = exp(-(lvsq / sh.inv2s[m])) * sh.coef[m];           // calc_plv (:140-145)
            im.lvsq[(size_t)m * N + n] = lvsq;
            im.pvl[(size_t)m * N + n] = plv;
            pl += plv * sh.pv[m];
        }
        if (pl < 1e-12) pl = 1e-12;                                         // :117 (NaN stays NaN)
        for (int m = 0; m < M; ++m)
            im.pvl[(size_t)m * N + n] = im.pvl[(size_t)m * N + n] * sh.pv[m] / pl;   // calc_pvl (:128)
    }
    __syncthreads();
    phase_lap(sh, PH_ESTEP);
}

// ---------------------------------------------------------------------------
// E6: weight matrix (vp_localisation.py:515-524) as an (M x N)(N x N) product
// ---------------------------------------------------------------------------
__device__ void wmat(const Img& im, EmShared& sh, double bias, bool use_weights) {
    phase_lap(sh, PH_OTHER);
    const int M = sh.M, N = im.N, tid = threadIdx.x;
    for (int m0 = 0; m0 < M; m0 += kMCH) {
        const int mc = min(kMCH, M - m0);
        for (int k0 = 0; k0 < N; k0 += 2 * kEmThreads) {
            const int ka = k0 + tid, kb = ka + kEmThreads;
            double acc[kMCH][2];
#pragma unroll
            for (int i = 0; i < kMCH; ++i) acc[i][0] = acc[i][1] = 0.0;
            if (use_weights) {
                for (int j0 = 0; j0 < N; j0 += kJT) {
                    __syncthreads();
                    for (int e = tid; e < kJT * kMCH; e += kEmThreads) {
                        int j = e % kJT, mm = e / kJT;
                        double val = 0.0;
                        if (mm < mc && j0 + j < N) val = im.pvl[(size_t)(m0 + mm) * N + j0 + j] * im.lweight[j0 + j];
                        sh.tile[j * kMCH + mm] = val;
                    }
                    __syncthreads();
                    const int jn = min(kJT, N - j0);
#pragma unroll 4
                    for (int j = 0; j < jn; ++j) {
                        const double* row = im.lsim + (size_t)(j0 + j) * N;
                        double a0 = ka < N ? row[ka] : 0.0;
                        double a1 = kb < N ? row[kb] : 0.0;
                        const double2* t2 = reinterpret_cast<const double2*>(sh.tile + j * kMCH);
#pragma unroll
                        for (int q = 0; q < kMCH / 2; ++q) {
                            double2 t = t2[q];
                            acc[2 * q][0] += t.x * a0; acc[2 * q][1] += t.x * a1;
                            acc[2 * q + 1][0] += t.y * a0; acc[2 * q + 1][1] += t.y * a1;
                        }
                    }
                }
            }
#pragma unroll
            for (int mm = 0; mm < kMCH; ++mm) {
                if (mm >= mc) break;
                if (ka < N) {
                    double lw = im.lweight[ka], w_ = im.pvl[(size_t)(m0 + mm) * N + ka] * lw;
                    im.w[(size_t)(m0 + mm) * N + ka] = (w_ + bias * lw * acc[mm][0]) / (1 + bias * lw * im.colsum[ka]);
                }
                if (kb < N) {
                    double lw = im.lweight[kb], w_ = im.pvl[(size_t)(m0 + mm) * N + kb] * lw;
                    im.w[(size_t)(m0 + mm) * N + kb] = (w_ + bias * lw * acc[mm][1]) / (1 + bias * lw * im.colsum[kb]);
                }
            }
        }
    }
    __syncthreads();
    phase_lap(sh, PH_WMAT);
}

// assoc[n] = argmax_m w[m,n] (first maximum, numpy.argmax; NaN counts as max)
__device__ void argmax_assoc(const Img& im, EmShared& sh) {
    const int M = sh.M, N = im.N;
    for (int n = threadIdx.x; n < N; n += kEmThreads) {
        int best = 0;
        double bv = M > 0 ? im.w[n] : 0.0;
        for (int m = 1; m < M; ++m) {
            double x = im.w[(size_t)m * N + n];
            if (isnan(bv)) break;
            if (x > bv || isnan(x)) { bv = x; best = m; }
        }
        im.assoc[n] = best;
    }
    __syncthreads();
}

// E9: calc_vp_line_counts (vp_localisation.py:482-512).  lvsq must have been
// computed (estep) for the same VP set that is being counted.
__device__ void line_counts(const Img& im, EmShared& sh, double thresh) {
    phase_lap(sh, PH_OTHER);
    const int M = sh.M, N = im.N, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    argmax_assoc(im, sh);
    for (int n = tid; n < N; n += kEmThreads) {
        int m = im.assoc[n];
        double dist = im.lvsq[(size_t)m * N + n];
        if (dist > thresh * sqrt(sh.s[m]) || im.lweight[n] == 0.0) im.assoc[n] = -1;
    }
    __syncthreads();
    for (int m = warp; m < M; m += kEmWarps) {
        int c = 0;
        double cw = 0.0;
        for (int n = lane; n < N; n += 32)
            if (im.assoc[n] == m) { ++c; cw += im.lweight[n]; }
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        cw = warp_sum(cw);
        if (lane == 0) { sh.cnt[m] = c; sh.cw[m] = cw; }
    }
    __syncthreads();
    phase_lap(sh, PH_COUNTS);
}

// E7 + E8 for VP m by one warp: smallest eigenvector of sum (w/max w)^2 l l^T
// over the selected lines, and the variance update.  sel < 0: all lines
// (weights w[m,:]); sel >= 0: only lines with assoc == sel (final refit).
// extra: optional second weight row added to the first (merge: w[j]+w[k]).
__device__ bool refit_vp(const Img& im, const double* wrow, const double* wrow2, int sel, double out[3]) {
    const int N = im.N, lane = threadIdx.x & 31;
    double mx = -INFINITY;
    bool any = false;
    for (int n = lane; n < N; n += 32) {
        if (sel >= 0 && im.assoc[n] != sel) continue;
        double x = wrow[n] + (wrow2 ? wrow2[n] : 0.0);
        any = true;
        mx = (isnan(mx) || isnan(x)) ? nan("") : (x > mx ? x : mx);
    }
    mx = warp_max_nanprop(mx);
    any = __any_sync(0xffffffffu, any);
    if (!any || mx == 0.0 || isnan(mx) || isinf(mx)) return false;      // :456-460 / LinAlgError
    double g[6] = {0, 0, 0, 0, 0, 0};
    int rows = 0, only = -1;
    for (int n = lane; n < N; n += 32) {
        if (sel >= 0 && im.assoc[n] != sel) continue;
        double x = (wrow[n] + (wrow2 ? wrow2[n] : 0.0)) / mx;
        double a = x * im.ln[3 * (size_t)n], b = x * im.ln[3 * (size_t)n + 1], c = x * im.ln[3 * (size_t)n + 2];
        g[0] += a * a; g[1] += a * b; g[2] += a * c; g[3] += b * b; g[4] += b * c; g[5] += c * c;
        ++rows; only = n;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) g[k] = warp_sum(g[k]);
    for (int o = 16; o > 0; o >>= 1) {
        rows += __shfl_xor_sync(0xffffffffu, rows, o);
        only = max(only, __shfl_xor_sync(0xffffffffu, only, o));
    }
    double e[3];
    bool okv;
    if (rows == 1) {
        // A single 1x3 row has a 2-D null space; LAPACK's full SVD (what numpy.linalg.svd runs,
        // vp_localisation.py:466) completes V with the Householder reflector of dgelqf/dlarfg:
        // V[:,2] = row 3 of H = I - tau v v^T, v = (1, a2/(a1-beta), a3/(a1-beta)).
        double x = (wrow[only] + (wrow2 ? wrow2[only] : 0.0)) / mx;
        double a1 = x * im.ln[3 * (size_t)only], a2 = x * im.ln[3 * (size_t)only + 1], a3 = x * im.ln[3 * (size_t)only + 2];
        double nrm = sqrt(a1 * a1 + a2 * a2 + a3 * a3);
        double beta = -copysign(nrm, a1);
        double tau = (beta - a1) / beta;
        double v2 = a2 / (a1 - beta), v3 = a3 / (a1 - beta);
        e[0] = -tau * v3; e[1] = -tau * v3 * v2; e[2] = 1.0 - tau * v3 * v3;
        double n2 = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        okv = n2 > 0.0 && !isnan(n2);
        if (okv) { e[0] /= n2; e[1] /= n2; e[2] /= n2; }
    } else {
        okv = smallest_eigvec3(g, e);
    }
    if (!okv) return false;
    double sg = sign_np(e[2]);                                           // :474
    out[0] = e[0] * sg; out[1] = e[1] * sg; out[2] = e[2] * sg;
    return true;
}

// s = exp(log(sum lvsq*pvl) - log(sum pvl))   (vp_localisation.py:301-304)
__device__ double variance_update(const Img& im, int m, int m2) {
    const int N = im.N, lane = threadIdx.x & 31;
    double num = 0.0, den = 0.0;
    for (int n = lane; n < N; n += 32) {
        double p = im.pvl[(size_t)m * N + n];
        double q = im.lvsq[(size_t)m * N + n];
        if (m2 >= 0) { p += im.pvl[(size_t)m2 * N + n]; q = 0.5 * (im.lvsq[(size_t)m2 * N + n] + q); }   // merge (:663-664)
        num += q * p;
        den += p;
    }
    num = warp_sum(num);
    den = warp_sum(den);
    return exp(log(num) - log(den));
}

// remove the VPs flagged in sh.rem[] from cur / nxt / s (numpy.delete along the VP axis)
__device__ void compact_vps(EmShared& sh) {
    if (threadIdx.x == 0) {
        int k = 0;
        for (int m = 0; m < sh.M; ++m) {
            if (sh.rem[m]) continue;
            if (k != m) {
                for (int c = 0; c < 3; ++c) { sh.cur[k][c] = sh.cur[m][c]; sh.nxt[k][c] = sh.nxt[m][c]; }
                sh.s[k] = sh.s[m];
            }
            ++k;
        }
        sh.M = k;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// E10: merge_vps (vp_localisation.py:633-684) on the sh.nxt row set
// ---------------------------------------------------------------------------
__device__ void merge_vps(const Img& im, EmShared& sh, const EmParams& P, double thresh) {
    phase_lap(sh, PH_OTHER);
    const int tid = threadIdx.x, warp = tid >> 5;
    while (true) {
        __syncthreads();
        const int M = sh.M;
        if (M <= 1) break;
        // closest pair: first minimum in row-major order of the (M,M) angle matrix (diag = pi)
        if (tid == 0) {
            double best = INFINITY;
            int bj = 0, bk = 0;
            bool nanfound = false;
            for (int j = 0; j < M && !nanfound; ++j)
                for (int k = 0; k < M; ++k) {
                    double a;
                    if (j == k) a = kPi;
                    else {
                        double c = sh.nxt[k][0] * sh.nxt[j][0] + sh.nxt[k][1] * sh.nxt[j][1] + sh.nxt[k][2] * sh.nxt[j][2];
                        c = fmin(fmax(c, -1.0), 1.0);
                        a = fabs(acos(fmin(fmax(fabs(c), -1.0), 1.0)));
                        if (isnan(c)) a = c;
                    }
                    if (isnan(a)) { best = a; bj = j; bk = k; nanfound = true; break; }   // numpy.argmin returns the first NaN
                    if (a < best) { best = a; bj = j; bk = k; }
                }
            sh.ia = bj; sh.ib = bk; sh.da = best;
        }
        __syncthreads();
        if (!(sh.da < thresh)) break;
        const int j = sh.ia, k = sh.ib;
        estep(im, sh, sh.nxt);
        wmat(im, sh, P.cfg.wbias, P.cfg.use_weights != 0);
        if (warp == 0) {
            double nv[3];
            bool okv = refit_vp(im, im.w + (size_t)j * im.N, im.w + (size_t)k * im.N, -1, nv);
            double sk = variance_update(im, k, j);
            if ((tid & 31) == 0) {
                sh.s[k] = sk;                                  // assigned before the test (:666)
                sh.flag = (okv && !(sk > 0.01)) ? 1 : 0;       // max_stdd = 0.01 (:633, :668)
                if (sh.flag) { sh.nxt[k][0] = nv[0]; sh.nxt[k][1] = nv[1]; sh.nxt[k][2] = nv[2]; }
            }
        }
        __syncthreads();
        if (!sh.flag) break;
        for (int m = tid; m < M; m += kEmThreads) sh.rem[m] = (m == j);
        __syncthreads();
        compact_vps(sh);
    }
    __syncthreads();
    phase_lap(sh, PH_MERGE);
}

// ---------------------------------------------------------------------------
// E11: split_best_vp (vp_localisation.py:527-630)
// ---------------------------------------------------------------------------
// UPGMA down to two clusters on the n x n matrix D (what scikit-learn's
// AgglomerativeClustering(linkage='average', n_clusters=2) computes on a
// complete connectivity graph); labels follow _hc_cut: label 0 = the root's
// child with the larger node id.
__device__ void average_linkage_two(double* D, int n, int* rep, int* nodeid, double* csize, EmShared& sh) {
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += kEmThreads) { rep[i] = i; nodeid[i] = i; csize[i] = 1.0; }
    __syncthreads();
    for (int step = 0; step < n - 2; ++step) {
        // global minimum over active pairs a < b, lexicographic tie-break
        double bd = INFINITY;
        int ba = -1, bb = -1;
        for (int a = tid; a < n; a += kEmThreads) {
            if (nodeid[a] < 0) continue;
            const double* row = D + (size_t)a * n;
            for (int b = a + 1; b < n; ++b) {
                if (nodeid[b] < 0) continue;
                double d = row[b];
                if (d < bd) { bd = d; ba = a; bb = b; }
            }
        }
        sh.redv[tid] = bd; sh.redi[tid] = ba; sh.redj[tid] = bb;
        __syncthreads();
        for (int o = kEmThreads / 2; o > 0; o >>= 1) {
            if (tid < o) {
                double od = sh.redv[tid + o];
                int oa = sh.redi[tid + o], ob = sh.redj[tid + o];
                bool take = oa >= 0 && (sh.redi[tid] < 0 || od < sh.redv[tid] ||
                                        (od == sh.redv[tid] && (oa < sh.redi[tid] || (oa == sh.redi[tid] && ob < sh.redj[tid]))));
                if (take) { sh.redv[tid] = od; sh.redi[tid] = oa; sh.redj[tid] = ob; }
            }
            __syncthreads();
        }
        const int a = sh.redi[0], b = sh.redj[0];
        __syncthreads();
        if (a < 0) break;
        const double na = csize[a], nb = csize[b];
        for (int c = tid; c < n; c += kEmThreads) {
            if (c == a || c == b || nodeid[c] < 0) continue;
            double dn = (na * D[(size_t)a * n + c] + nb * D[(size_t)b * n + c]) / (na + nb);   // average_merge
            D[(size_t)a * n + c] = dn;
            D[(size_t)c * n + a] = dn;
        }
        for (int i = tid; i < n; i += kEmThreads)
            if (rep[i] == b) rep[i] = a;
        __syncthreads();
        if (tid == 0) { csize[a] = na + nb; nodeid[a] = n + step; nodeid[b] = -1; }
        __syncthreads();
    }
    // the two survivors; label 0 = larger node id
    if (tid == 0) {
        int c0 = -1, c1 = -1;
        for (int i = 0; i < n; ++i)
            if (nodeid[i] >= 0) { if (c0 < 0) c0 = i; else c1 = i; }
        if (c1 >= 0 && nodeid[c1] > nodeid[c0]) { int t = c0; c0 = c1; c1 = t; }
        sh.ia = c0; sh.ib = c1;
    }
    __syncthreads();
    const int c0 = sh.ia;
    for (int i = tid; i < n; i += kEmThreads) rep[i] = (rep[i] == c0) ? 0 : 1;
    __syncthreads();
}

__device__ void split_best_vp(const Img& im, EmShared& sh, const EmParams& P, double min_diff) {
    const int M = sh.M, N = im.N, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    argmax_assoc(im, sh);
    // global maximum of w (weightMatrix.max(), :539) only decides the sign of the greedy entries
    double lm = -INFINITY;
    for (size_t e = tid; e < (size_t)M * N; e += kEmThreads) {
        double x = im.w[e];
        lm = (isnan(lm) || isnan(x)) ? nan("") : (x > lm ? x : lm);
    }
    lm = warp_max_nanprop(lm);
    if (lane == 0) sh.redv[warp] = lm;
    __syncthreads();
    if (tid == 0) {
        double g = sh.redv[0];
        for (int k = 1; k < kEmWarps; ++k) { double t = sh.redv[k]; g = (isnan(g) || isnan(t)) ? nan("") : (t > g ? t : g); }
        sh.da = g;
    }
    __syncthreads();
    const double wmax = sh.da;
    // std of the segment angles of the lines greedily assigned to each VP (:541-544)
    for (int m = warp; m < M; m += kEmWarps) {
        double sum = 0.0;
        int c = 0;
        for (int n = lane; n < N; n += 32)
            if (im.assoc[n] == m && (im.w[(size_t)m * N + n] / wmax) > 0) { sum += im.langle[n]; ++c; }
        sum = warp_sum(sum);
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        double mean = sum / c, var = 0.0;
        for (int n = lane; n < N; n += 32)
            if (im.assoc[n] == m && (im.w[(size_t)m * N + n] / wmax) > 0) { double d = im.langle[n] - mean; var += d * d; }
        var = warp_sum(var);
        if (lane == 0) sh.ang[m] = c > 0 ? sqrt(var / c) : nan("");
    }
    // number of lines per VP
    for (int m = warp; m < M; m += kEmWarps) {
        int c = 0;
        for (int n = lane; n < N; n += 32) c += im.assoc[n] == m;
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) sh.cnt[m] = c;
    }
    __syncthreads();
    if (tid == 0) {
        // argsort(std)[::-1]: ascending, NaN last, stable; then reversed (:546-547)
        int ord[kMaxM];
        for (int m = 0; m < M; ++m) ord[m] = m;
        for (int i = 1; i < M; ++i) {
            int x = ord[i];
            double kx = isnan(sh.ang[x]) ? INFINITY : sh.ang[x];
            bool nx = isnan(sh.ang[x]);
            int p = i;
            while (p > 0) {
                int y = ord[p - 1];
                double ky = isnan(sh.ang[y]) ? INFINITY : sh.ang[y];
                bool ny = isnan(sh.ang[y]);
                bool greater = (ny && !nx) || (!ny && !nx && ky > kx);
                if (!greater) break;
                ord[p] = y; --p;
            }
            ord[p] = x;
        }
        int worst = -1;
        for (int m = 0; m < M; ++m) {
            int cand = ord[M - 1 - m];
            double px = sh.cur[m][0] / sh.cur[m][2], py = sh.cur[m][1] / sh.cur[m][2];     // row m, not cand (:557)
            if (sh.cnt[cand] > 8 && px > -1 && py > -1 && px < 1 && py < 1) { worst = cand; break; }
        }
        sh.ia = worst;
    }
    __syncthreads();
    const int worst = sh.ia;
    if (worst < 0) return;
    const int nw = sh.cnt[worst];
    // scratch layout (doubles): D[nw*nw] | csize[nw] | ints: idx[nw] rep[nw] nodeid[nw]
    size_t need = (size_t)nw * nw + nw + (3 * (size_t)nw + 1) / 2 + 4;
    double* scratch = im.lvsq;
    bool locked = false;
    // the line -> VP association and the weights are needed below: copy what we need first
    // (assoc lives outside the scratch region; lweight/ln too)
    if (need > im.scratch_cap) {
        if (need > P.overflow_cap) return;        // cannot split: leave the hypothesis set unchanged
        if (tid == 0) { while (atomicCAS(P.overflow_lock, 0, 1) != 0) __nanosleep(200); }
        __syncthreads();
        scratch = P.overflow;
        locked = true;
    }
    double* D = scratch;
    double* csize = D + (size_t)nw * nw;
    int* idx = reinterpret_cast<int*>(csize + nw);
    int* rep = idx + nw;
    int* nodeid = rep + nw;
    if (tid == 0) {
        int k = 0;
        for (int n = 0; n < N; ++n) if (im.assoc[n] == worst) idx[k++] = n;
    }
    __syncthreads();
    for (size_t e = tid; e < (size_t)nw * nw; e += kEmThreads) {
        int a = (int)(e / nw), b = (int)(e % nw);
        double d = 0.0;
        if (a != b) d = 1.0 - cosangle(load_seg(im.lp, idx[a]), load_seg(im.lp, idx[b]), 2.0);     // :572
        D[e] = d;
    }
    __syncthreads();
    average_linkage_two(D, nw, rep, nodeid, csize, sh);
    // per cluster: smallest right-singular vector of the lweight-scaled lines (:580-602)
    if (warp < 2) {
        const int c = warp;
        double g[6] = {0, 0, 0, 0, 0, 0};
        int cnt = 0;
        for (int q = lane; q < nw; q += 32) {
            if (rep[q] != c) continue;
            int n = idx[q];
            double lw = im.lweight[n];
            double a = im.ln[3 * (size_t)n] * lw, b = im.ln[3 * (size_t)n + 1] * lw, cc = im.ln[3 * (size_t)n + 2] * lw;
            g[0] += a * a; g[1] += a * b; g[2] += a * cc; g[3] += b * b; g[4] += b * cc; g[5] += cc * cc;
            ++cnt;
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) g[k] = warp_sum(g[k]);
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        double e[3] = {0, 0, 0};
        bool okv = cnt >= 3 && smallest_eigvec3(g, e);
        if (lane == 0) {
            sh.ok[c] = okv;
            double sg = e[2] < 0 ? -1.0 : 1.0;
            sh.nv[c][0] = e[0] * sg; sh.nv[c][1] = e[1] * sg; sh.nv[c][2] = e[2] * sg;
        }
    }
    __syncthreads();
    if (locked && tid == 0) { __threadfence(); atomicExch(P.overflow_lock, 0); }
    if (tid == 0) {
        if (sh.ok[0] && sh.ok[1]) {
            double c = sh.nv[0][0] * sh.nv[1][0] + sh.nv[0][1] * sh.nv[1][1] + sh.nv[0][2] * sh.nv[1][2];
            c = fmin(fmax(c, -1.0), 1.0);
            double ang = fabs(acos(fmin(fmax(fabs(c), -1.0), 1.0)));
            if (ang > min_diff && sh.M < kMaxM) {
                double stdd = sh.s[worst] / 2;
                for (int k = 0; k < 3; ++k) { sh.cur[worst][k] = sh.nv[0][k]; sh.cur[sh.M][k] = sh.nv[1][k]; sh.nxt[sh.M][k] = 0.0; }
                sh.s[worst] = stdd;
                sh.s[sh.M] = stdd;
                sh.M += 1;
            }
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// the persistent kernel
// ---------------------------------------------------------------------------
__device__ void write_result(const EmParams& P, int b, int base, const Img& im, EmShared& sh, int status, int iters,
                             bool have_vps) {
    const int tid = threadIdx.x, N = im.N;
    if (tid == 0) {
        P.out.status[b] = status;
        P.out.iterations[b] = iters;
        P.out.n_vp[b] = have_vps ? sh.M : 0;
    }
    for (int m = tid; m < kMaxM; m += kEmThreads) {
        bool live = have_vps && m < sh.M;
        for (int c = 0; c < 3; ++c) P.out.vp[((size_t)b * kMaxM + m) * 3 + c] = live ? sh.nxt[m][c] : 0.0;
        P.out.sigma[(size_t)b * kMaxM + m] = live ? sh.s[m] : 0.0;
        P.out.counts[(size_t)b * kMaxM + m] = live ? sh.cnt[m] : 0;
        P.out.counts_weighted[(size_t)b * kMaxM + m] = live ? sh.cw[m] : 0.0;
    }
    for (int n = tid; n < N; n += kEmThreads) P.out.vp_assoc[base + n] = have_vps ? im.assoc[n] : -1;
    if (P.out.decision_metric && have_vps) {
        double* dm = P.out.decision_metric + (size_t)kMaxM * base;
        for (size_t e = tid; e < (size_t)sh.M * N; e += kEmThreads) dm[e] = im.w[e];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kEmThreads, 2) em_kernel(EmParams P) {
    __shared__ __align__(16) EmShared sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const vpk_em_config& cfg = P.cfg;
    const double max_stdd = 1e-6;            // angle mode (:197)
    double* slot = P.ws + (size_t)blockIdx.x * P.ws_stride;
    if (tid == 0) {
        sh.timing = P.phase_cycles != nullptr;
        for (int k = 0; k < PH_N; ++k) sh.phase[k] = 0;
        sh.t_last = clock64();
    }

    while (true) {
        __syncthreads();
        if (tid == 0) sh.img = atomicAdd(P.queue, 1);
        __syncthreads();
        const int q = sh.img;
        if (q >= P.B) break;
        const int b = P.order[q];
        const int base = P.offsets[b];
        Img im;
        im.N = P.offsets[b + 1] - base;
        const int N = im.N;
        im.lp = P.segs + 4 * (size_t)base;
        {
            double* p = slot;
            im.lsim = p; p += (size_t)N * N;
            im.ln = p; p += 3 * (size_t)N;
            im.lweight = p; p += N;
            im.colsum = p; p += N;
            im.langle = p; p += N;
            im.assoc = reinterpret_cast<int*>(p); p += (N + 1) / 2 + 1;
            im.lvsq = p; p += (size_t)kMaxM * N;
            im.pvl = p; p += (size_t)kMaxM * N;
            im.w = p; p += (size_t)kMaxM * N;
            im.scratch_cap = 3 * (size_t)kMaxM * N;
        }
        if (tid == 0) sh.M = 0;
        __syncthreads();
        if (N == 0) { write_result(P, b, base, im, sh, VPK_EM_NO_INITIAL_VPS, 0, false); continue; }

        // ---- per-line constants: unit lines (:186/:226), segment angles (:765-776)
        for (int n = tid; n < N; n += kEmThreads) {
            const double* l = P.lines + 3 * (size_t)(base + n);
            double a = l[0], bb = l[1], c = l[2];
            double nr = sqrt(a * a + bb * bb + c * c);
            a /= nr; bb /= nr; c /= nr;
            nr = sqrt(a * a + bb * bb + c * c);               // the reference normalises twice
            im.ln[3 * (size_t)n] = a / nr; im.ln[3 * (size_t)n + 1] = bb / nr; im.ln[3 * (size_t)n + 2] = c / nr;
            Seg sg = load_seg(im.lp, n);
            double vx = sg.x1 - sg.x2, vy = sg.y1 - sg.y2;
            vx = vx / sqrt(vx * vx + vy * vy);
            double phi = fabs(acos(fmin(fmax(vx, -1.0), 1.0)));
            im.langle[n] = phi > 0.5 * kPi ? kPi - phi : phi;
        }
        __syncthreads();
        phase_lap(sh, PH_SETUP);
        if (cfg.use_weights) pair_pass(im, sh);
        else {
            for (int n = tid; n < N; n += kEmThreads) im.colsum[n] = 0.0;
        }
        __syncthreads();
        line_rating(im, sh, cfg.use_weights != 0);

        // ---- initial hypotheses and prior
        const bool have_init = P.init_vp != nullptr;
        phase_lap(sh, PH_OTHER);
        init_prior_and_vps(P, b, sh, have_init);
        phase_lap(sh, PH_INIT);
        if (have_init) {
            if (tid == 0) {
                int i0 = P.init_off[b], i1 = P.init_off[b + 1];
                int M = min(i1 - i0, kMaxM);
                for (int m = 0; m < M; ++m) {
                    const double* v = P.init_vp + 3 * (size_t)(i0 + m);
                    double nr = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
                    sh.cur[m][0] = v[0] / nr; sh.cur[m][1] = v[1] / nr; sh.cur[m][2] = v[2] / nr;
                }
                sh.M = M;
            }
            __syncthreads();
        }
        if (sh.M == 0) { write_result(P, b, base, im, sh, VPK_EM_NO_INITIAL_VPS, 0, false); continue; }
        for (int m = tid; m < kMaxM; m += kEmThreads) {
            sh.s[m] = sh.sigma_prior * 1e-6;                  // s_init (:219)
            sh.nxt[m][0] = sh.nxt[m][1] = sh.nxt[m][2] = 0.0;
        }
        __syncthreads();

        // ---- initial E-step, weights, counts; drop VPs with < 3 lines (:245-251)
        estep(im, sh, sh.cur);
        wmat(im, sh, cfg.wbias, cfg.use_weights != 0);
        line_counts(im, sh, cfg.outlier_thresh);
        for (int m = tid; m < sh.M; m += kEmThreads) sh.rem[m] = sh.cnt[m] < 3;
        __syncthreads();
        compact_vps(sh);

        int status = VPK_EM_NO_VPS_LEFT, iters = 0;
        bool done = false;
        for (int i = 0; i < cfg.num_iter && !done; ++i) {
            if (sh.M == 0) break;                              // "No VPs left!" (:258)
            if (i % cfg.split_merge_freq == 0 && i > 0 && i < 100 && cfg.do_split) {     // :262
                estep(im, sh, sh.cur);
                wmat(im, sh, cfg.wbias, cfg.use_weights != 0);
                phase_lap(sh, PH_OTHER);
                split_best_vp(im, sh, P, cfg.merge_thresh);
                phase_lap(sh, PH_SPLIT);
            }
            estep(im, sh, sh.cur);                             // :273
            wmat(im, sh, cfg.wbias, cfg.use_weights != 0);     // :282
            // ---- M-step (:284-322), one warp per VP
            const int M = sh.M;
            phase_lap(sh, PH_OTHER);
            for (int m = warp; m < M; m += kEmWarps) {
                if (!cfg.do_iterations) {
                    if (lane == 0) { sh.rem[m] = 0; sh.err[m] = 0.0; for (int c = 0; c < 3; ++c) sh.nxt[m][c] = sh.cur[m][c]; }
                    continue;
                }
                double nv[3];
                bool okv = refit_vp(im, im.w + (size_t)m * N, nullptr, -1, nv);
                double sv = okv ? variance_update(im, m, -1) : 0.0;
                if (lane == 0) {
                    int rem = 0;
                    double err = 0.0;
                    if (!okv) rem = 1;
                    else {
                        sh.nxt[m][0] = nv[0]; sh.nxt[m][1] = nv[1]; sh.nxt[m][2] = nv[2];
                        sv = isnan(sv) ? sv : fmin(sv, max_stdd);                  // :306
                        sv = isnan(sv) ? sv : fmax(sv, cfg.s_thresh);              // :307
                        sh.s[m] = sv;
                        if (isnan(sv)) rem = 1;
                        else {
                            double d = fabs(sh.cur[m][0] * nv[0] + sh.cur[m][1] * nv[1] + sh.cur[m][2] * nv[2]);
                            err = acos(fmin(d, 1.0));                               // :312
                            if (isnan(d)) err = d;
                            if (err > 1.5) rem = 1;
                        }
                    }
                    sh.rem[m] = rem;
                    sh.err[m] = err;
                }
            }
            __syncthreads();
            if (tid == 0) {
                double mx = 0.0;
                for (int m = 0; m < M; ++m) {
                    double e = sh.err[m];
                    if (isnan(e) || isnan(mx)) mx = nan("");      // numpy.maximum propagates NaN
                    else if (e > mx) mx = e;
                }
                sh.da = mx;
            }
            __syncthreads();
            const double max_err = sh.da;
            phase_lap(sh, PH_MSTEP);
            compact_vps(sh);
            estep(im, sh, sh.cur);                             // :332 (index i, with the new variances)

            if (max_err < cfg.final_convergence || i == cfg.num_iter - 1 || !cfg.do_iterations) {   // :335
                if (cfg.do_merge) merge_vps(im, sh, P, cfg.merge_thresh * 10);      // :339
                estep(im, sh, sh.cur);                         // :344 (index i, sic)
                wmat(im, sh, cfg.wbias, cfg.use_weights != 0);
                argmax_assoc(im, sh);
                // hard-assignment refit (:353-392)
                const int M2 = sh.M;
                for (int m = warp; m < M2; m += kEmWarps) {
                    int have = 0;
                    for (int n = lane; n < N; n += 32) have |= im.assoc[n] == m;
                    have = __any_sync(0xffffffffu, have);
                    if (!have) { if (lane == 0) sh.rem[m] = 0; continue; }
                    double nv[3];
                    bool okv = refit_vp(im, im.w + (size_t)m * N, nullptr, m, nv);
                    double sv = okv ? variance_update(im, m, -1) : 0.0;
                    if (lane == 0) {
                        int rem = 0;
                        if (!okv) rem = 1;
                        else {
                            sh.nxt[m][0] = nv[0]; sh.nxt[m][1] = nv[1]; sh.nxt[m][2] = nv[2];
                            sv = isnan(sv) ? sv : fmin(sv, max_stdd);               // :377
                            sh.s[m] = sv;
                            if (isnan(sv) || sv < cfg.s_thresh) rem = 1;            // :379
                            else {
                                double d = fabs(sh.cur[m][0] * nv[0] + sh.cur[m][1] * nv[1] + sh.cur[m][2] * nv[2]);
                                double err = acos(fmin(d, 1.0));
                                if (err > 1.5) rem = 1;
                            }
                        }
                        sh.rem[m] = rem;
                    }
                }
                __syncthreads();
                compact_vps(sh);
                estep(im, sh, sh.cur);                         // :398
                wmat(im, sh, cfg.wbias, cfg.use_weights != 0);
                if (sh.M == 0 || N == 0) { status = VPK_EM_NO_VPS_LEFT; break; }       // "decision metric is empty"
                // keep only the VPs that win at least one line (:406-413)
                argmax_assoc(im, sh);
                for (int m = tid; m < sh.M; m += kEmThreads) sh.rem[m] = 1;
                __syncthreads();
                for (int n = tid; n < N; n += kEmThreads) sh.rem[im.assoc[n]] = 0;
                __syncthreads();
                compact_vps(sh);
                estep(im, sh, sh.nxt);                         // :415 (index i+1)
                wmat(im, sh, cfg.wbias, cfg.use_weights != 0);
                line_counts(im, sh, cfg.outlier_thresh);
                // iteratively drop VPs with too few lines (:423-437)
                while (true) {
                    if (tid == 0) {
                        int v = -1;
                        for (int m = 0; m < sh.M; ++m) if (sh.cnt[m] < cfg.num_min_lines) { v = m; break; }
                        sh.ia = v;
                    }
                    __syncthreads();
                    const int v = sh.ia;
                    if (v < 0) break;
                    for (int m = tid; m < sh.M; m += kEmThreads) sh.rem[m] = (m == v);
                    __syncthreads();
                    compact_vps(sh);
                    if (sh.M == 0) break;                       // the reference crashes here (argmax of an empty axis)
                    estep(im, sh, sh.nxt);
                    wmat(im, sh, cfg.wbias, cfg.use_weights != 0);
                    line_counts(im, sh, cfg.outlier_thresh);
                }
                status = sh.M > 0 ? VPK_EM_OK : VPK_EM_NO_VPS_LEFT;
                iters = i;
                done = true;
                break;
            }
            if (i % cfg.split_merge_freq == 0 && i > 0 && i <= 100 + cfg.split_merge_freq && cfg.do_merge)   // :444
                merge_vps(im, sh, P, cfg.merge_thresh);
            // v[i+1] becomes the current set
            for (int m = tid; m < sh.M; m += kEmThreads)
                for (int c = 0; c < 3; ++c) { sh.cur[m][c] = sh.nxt[m][c]; sh.nxt[m][c] = 0.0; }
            __syncthreads();
        }
        write_result(P, b, base, im, sh, status, iters, done && status == VPK_EM_OK);
    }
    if (tid == 0 && sh.timing) {
        phase_lap(sh, PH_OTHER);
        for (int k = 0; k < PH_N; ++k) atomicAdd(P.phase_cycles + k, sh.phase[k]);
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct EmState {
    DBuf ws, order, queue, overflow, resp, out_small, out_assoc, out_dm, init_vp, init_off, sphere, phase;
};

void em_free(vpk_ctx* ctx) {
    if (!ctx->em) return;
    EmState* e = ctx->em;
    e->ws.release(); e->order.release(); e->queue.release(); e->overflow.release(); e->resp.release();
    e->out_small.release(); e->out_assoc.release(); e->out_dm.release(); e->init_vp.release(); e->init_off.release();
    e->sphere.release(); e->phase.release();
    delete e;
    ctx->em = nullptr;
}

static size_t slot_doubles(int nmax) {
    size_t N = (size_t)nmax;
    return N * N + 3 * N + 3 * N + (N + 1) / 2 + 1 + 3 * (size_t)kMaxM * N + 16;
}

int em_dev(vpk_ctx* ctx, const double* d_lines, const double* d_segments, const int32_t* d_offsets,
           const int32_t* h_offsets, int32_t B, const float* d_resp_f32, const double* d_resp_f64,
           const uint8_t* d_sphere, int32_t S, const double* d_init_vp, const int32_t* d_init_off,
           const vpk_em_config* cfg, const EmDeviceOut& out) {
    if (B <= 0) return VPK_OK;
    if (!ctx->em) ctx->em = new EmState();
    EmState* st = ctx->em;
    int nmax = 1;
    std::vector<int32_t> order(B);
    for (int b = 0; b < B; ++b) { order[b] = b; nmax = std::max(nmax, h_offsets[b + 1] - h_offsets[b]); }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return (h_offsets[a + 1] - h_offsets[a]) > (h_offsets[b + 1] - h_offsets[b]);
    });
    int grid = std::min(B, 2 * ctx->num_sms);
    size_t stride = slot_doubles(nmax);
    stride = (stride + 1) & ~(size_t)1;
    VPK_TRY(st->ws.ensure(stride * sizeof(double) * (size_t)grid));
    VPK_TRY(st->order.ensure(sizeof(int32_t) * (size_t)B));
    VPK_TRY(st->queue.ensure(2 * sizeof(int)));
    size_t ov = (size_t)nmax * nmax + 4 * (size_t)nmax + 16;
    VPK_TRY(st->overflow.ensure(ov * sizeof(double)));
    VPK_TRY(ctx->h_stage.ensure(sizeof(int32_t) * (size_t)B));
    memcpy(ctx->h_stage.p, order.data(), sizeof(int32_t) * (size_t)B);
    VPK_CUDA(cudaMemcpyAsync(st->order.p, ctx->h_stage.p, sizeof(int32_t) * (size_t)B, cudaMemcpyHostToDevice, ctx->stream));
    VPK_CUDA(cudaMemsetAsync(st->queue.p, 0, 2 * sizeof(int), ctx->stream));
    EmParams P;
    P.lines = d_lines; P.segs = d_segments; P.offsets = d_offsets; P.B = B;
    P.resp32 = d_resp_f32; P.resp64 = d_resp_f64; P.sphere = d_sphere; P.S = S;
    P.init_vp = d_init_vp; P.init_off = d_init_off;
    P.cfg = *cfg;
    P.order = st->order.as<int32_t>();
    P.queue = st->queue.as<int>();
    P.ws = st->ws.as<double>(); P.ws_stride = stride; P.nmax = nmax;
    P.overflow = st->overflow.as<double>(); P.overflow_cap = ov; P.overflow_lock = st->queue.as<int>() + 1;
    P.out = out;
    P.phase_cycles = nullptr;
    if (ctx->profiling) {
        VPK_TRY(st->phase.ensure(PH_N * sizeof(unsigned long long)));
        VPK_CUDA(cudaMemsetAsync(st->phase.p, 0, PH_N * sizeof(unsigned long long), ctx->stream));
        P.phase_cycles = st->phase.as<unsigned long long>();
    }
    {
        KernelScope ks(ctx, "em_persistent");
        em_kernel<<<grid, kEmThreads, 0, ctx->stream>>>(P);
        VPK_TRY(check_launch("em_persistent"));
    }
    VPK_CUDA(cudaStreamSynchronize(ctx->stream));      // h_stage (order) is reused by later calls
    if (P.phase_cycles) {
        // per-phase share of the kernel: CTA-cycles summed over CTAs, reported as "em:<phase>"
        // pseudo-entries in CTA-milliseconds at the SM clock (not additive with kernel times)
        unsigned long long h[PH_N];
        VPK_CUDA(cudaMemcpy(h, st->phase.p, sizeof(h), cudaMemcpyDeviceToHost));
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
        for (int k = 0; k < PH_N; ++k) {
            auto& e = ctx->prof[kPhaseNames[k]];
            e.total_ms += (double)h[k] / (khz > 0 ? (double)khz : 1.0e6);
            e.launches += 1;
        }
    }
    return VPK_OK;
}

}  // namespace vpk

using namespace vpk;

extern "C" {

void vpk_em_default_config(vpk_em_config* c) {
    if (!c) return;
    c->num_iter = 100; c->num_init_vp = 25; c->split_merge_freq = 10; c->num_min_lines = 3;
    c->do_merge = 1; c->do_split = 1; c->do_iterations = 1; c->use_weights = 1;
    c->wbias = 1.0; c->merge_thresh = 1e-3; c->outlier_thresh = 1.96 * 1.96; c->final_convergence = 5e-3;
    c->s_thresh = 1e-200;
}

int vpk_em(vpk_ctx* ctx, const double* lines, const double* segments, const int32_t* offsets, int32_t B,
           const double* responses, const uint8_t* sphere_images, int32_t S, const double* init_vp,
           const int32_t* init_vp_offsets, const vpk_em_config* cfg_in, vpk_em_result* out) {
    if (!ctx || !offsets || !out || B < 0 || S <= 0) { set_error("vpk_em: bad argument"); return VPK_ERR_ARG; }
    if (B == 0) return VPK_OK;
    if (!responses || (!sphere_images && !init_vp)) { set_error("vpk_em: responses and sphere_images (or init_vp) are required"); return VPK_ERR_ARG; }
    if (!out->status || !out->n_vp || !out->iterations || !out->vp || !out->sigma || !out->counts || !out->counts_weighted ||
        !out->vp_assoc) { set_error("vpk_em: result arrays must be allocated by the caller"); return VPK_ERR_ARG; }
    vpk_em_config cfg;
    if (cfg_in) cfg = *cfg_in; else vpk_em_default_config(&cfg);
    if (cfg.split_merge_freq <= 0 || cfg.num_iter < 0 || cfg.num_init_vp < 0) { set_error("vpk_em: bad config"); return VPK_ERR_ARG; }
    if (offsets[0] != 0) { set_error("vpk_em: offsets[0] must be 0"); return VPK_ERR_ARG; }
    for (int b = 0; b < B; ++b) if (offsets[b + 1] < offsets[b]) { set_error("vpk_em: offsets must be non-decreasing"); return VPK_ERR_ARG; }
    const int64_t sumN = offsets[B];
    if (sumN > 0 && (!lines || !segments)) { set_error("vpk_em: lines/segments are NULL"); return VPK_ERR_ARG; }
    VPK_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->em) ctx->em = new EmState();
    EmState* st = ctx->em;
    const size_t plane = (size_t)S * S;
    VPK_TRY(ctx->d_lines.ensure((sumN + 1) * 3 * sizeof(double)));
    VPK_TRY(ctx->d_segments.ensure((sumN + 1) * 4 * sizeof(double)));
    VPK_TRY(ctx->d_offsets.ensure((B + 1) * sizeof(int32_t)));
    VPK_TRY(st->resp.ensure((size_t)B * kCells * sizeof(double)));
    if (sphere_images) VPK_TRY(st->sphere.ensure(plane * B));
    // small outputs packed: status,n_vp,iterations (3B int32) | counts (B*64 int32) | vp | sigma | cw
    const size_t n_i32 = 3 * (size_t)B + (size_t)B * kMaxM;
    const size_t n_f64 = (size_t)B * kMaxM * 5;
    VPK_TRY(st->out_small.ensure(n_f64 * sizeof(double) + n_i32 * sizeof(int32_t) + 64));
    VPK_TRY(st->out_assoc.ensure((sumN + 1) * sizeof(int32_t)));
    if (out->decision_metric) VPK_TRY(st->out_dm.ensure(((size_t)kMaxM * sumN + 1) * sizeof(double)));
    if (sumN) {
        VPK_CUDA(cudaMemcpyAsync(ctx->d_lines.p, lines, sumN * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        VPK_CUDA(cudaMemcpyAsync(ctx->d_segments.p, segments, sumN * 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    VPK_CUDA(cudaMemcpyAsync(ctx->d_offsets.p, offsets, (B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    VPK_CUDA(cudaMemcpyAsync(st->resp.p, responses, (size_t)B * kCells * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (sphere_images) VPK_CUDA(cudaMemcpyAsync(st->sphere.p, sphere_images, plane * B, cudaMemcpyHostToDevice, ctx->stream));
    const double* d_init = nullptr;
    const int32_t* d_ioff = nullptr;
    if (init_vp) {
        if (!init_vp_offsets) { set_error("vpk_em: init_vp needs init_vp_offsets"); return VPK_ERR_ARG; }
        size_t nv = init_vp_offsets[B];
        VPK_TRY(st->init_vp.ensure((nv + 1) * 3 * sizeof(double)));
        VPK_TRY(st->init_off.ensure((B + 1) * sizeof(int32_t)));
        VPK_CUDA(cudaMemcpyAsync(st->init_vp.p, init_vp, nv * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        VPK_CUDA(cudaMemcpyAsync(st->init_off.p, init_vp_offsets, (B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        d_init = st->init_vp.as<double>();
        d_ioff = st->init_off.as<int32_t>();
    }
    EmDeviceOut d;
    double* f = st->out_small.as<double>();
    d.vp = f; f += (size_t)B * kMaxM * 3;
    d.sigma = f; f += (size_t)B * kMaxM;
    d.counts_weighted = f; f += (size_t)B * kMaxM;
    int32_t* ip = reinterpret_cast<int32_t*>(f);
    d.status = ip; ip += B;
    d.n_vp = ip; ip += B;
    d.iterations = ip; ip += B;
    d.counts = ip;
    d.vp_assoc = st->out_assoc.as<int32_t>();
    d.decision_metric = out->decision_metric ? st->out_dm.as<double>() : nullptr;
    VPK_TRY(em_dev(ctx, ctx->d_lines.as<double>(), ctx->d_segments.as<double>(), ctx->d_offsets.as<int32_t>(), offsets, B,
                   nullptr, st->resp.as<double>(), sphere_images ? st->sphere.as<uint8_t>() : nullptr, S, d_init, d_ioff,
                   &cfg, d));
    auto D2H = [&](void* dst, const void* src, size_t bytes) {
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    };
    VPK_CUDA(D2H(out->vp, d.vp, (size_t)B * kMaxM * 3 * sizeof(double)));
    VPK_CUDA(D2H(out->sigma, d.sigma, (size_t)B * kMaxM * sizeof(double)));
    VPK_CUDA(D2H(out->counts_weighted, d.counts_weighted, (size_t)B * kMaxM * sizeof(double)));
    VPK_CUDA(D2H(out->status, d.status, (size_t)B * sizeof(int32_t)));
    VPK_CUDA(D2H(out->n_vp, d.n_vp, (size_t)B * sizeof(int32_t)));
    VPK_CUDA(D2H(out->iterations, d.iterations, (size_t)B * sizeof(int32_t)));
    VPK_CUDA(D2H(out->counts, d.counts, (size_t)B * kMaxM * sizeof(int32_t)));
    if (sumN) VPK_CUDA(D2H(out->vp_assoc, d.vp_assoc, sumN * sizeof(int32_t)));
    if (out->decision_metric && sumN) VPK_CUDA(D2H(out->decision_metric, d.decision_metric, (size_t)kMaxM * sumN * sizeof(double)));
    VPK_CUDA(cudaStreamSynchronize(ctx->stream));
    return VPK_OK;
}

}  // extern "C"
