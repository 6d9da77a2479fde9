// TEST INFRASTRUCTURE ONLY.  Single-threaded host build of the EM state machine
// (vanishing_points_2017_b200/csrc/em_core.cuh, team = 1 thread) so that the
// control flow the CUDA kernels run can be checked against the golden vectors on
// a box without a GPU.  Never linked into libvpk.so, never imported by the package.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../vanishing_points_2017_b200/csrc/em_core.cuh"

using namespace vpk::em;

extern "C" int hostsim_em(const double* lines, const double* segs, int N, const double* resp, const uint8_t* sphere, int S,
                          const double* init_vp, int n_init, const vpk_em_config* cfg, int32_t* status, int32_t* n_vp,
                          int32_t* iterations, double* vp, double* sigma, int32_t* counts, double* cw, int32_t* assoc,
                          double* dm, int32_t* supersteps) {
    const Team T = make_team();
    std::vector<double> ws(slot_doubles(N) + 16, 0.0);
    Img im = make_img(N, ws.data(), segs);
    EmSlot* st = new EmSlot();
    memset(st, 0, sizeof(EmSlot));
    st->img = 0; st->N = N; st->base = 0; st->ws_off = 0; st->phase = PH_DONE;
    EmOut out = {status, n_vp, iterations, vp, sigma, counts, cw, assoc, dm};
    line_constants(im, lines, T);
    if (cfg->use_weights) {
        // em_pair_kernel, sequentially
        for (int k = 0; k < N; ++k) {
            const SegPre sk = seg_pre(load_seg(im.lp, k));
            double cd[kK1]; int cj[kK1]; int cnt = 0;
            double cs = 0.0;
            for (int j = 0; j < N; ++j) {
                double val = 0.0;
                if (j == k) knn_insert(cd, cj, 1, cnt, 16.0, j);
                else {
                    const SegPre sj = seg_pre(load_seg(im.lp, j));
                    const double d2 = seg_distance2(sj, sk);
                    val = similarity_pre(sj, sk, d2);
                    knn_insert(cd, cj, 1, cnt, d2, j);
                }
                im.lsim[lsim_index(N, j, k)] = val;
                cs += val;
            }
            im.colsum[k] = cs;
            im.lweight[k] = rate_line(im.lp, k, cj, cd, cnt, N);
        }
    } else {
        for (int n = 0; n < N; ++n) { im.lweight[n] = 1.0; im.colsum[n] = 0.0; }
    }
#if defined(VPK_HOST_TRACE)
    if (getenv("VPK_TRACE_LWEIGHT")) {
        FILE* fh = fopen(getenv("VPK_TRACE_LWEIGHT"), "wb");
        fwrite(im.lweight, sizeof(double), N, fh);
        fwrite(im.colsum, sizeof(double), N, fh);
        fwrite(im.lsim, sizeof(double), lsim_doubles(N), fh);
        fclose(fh);
    }
#endif
    InitScratch* isc = new InitScratch();
    for (int c = 0; c < kCells; ++c) isc->resp[c] = resp[c];
    PostScratch* sc = new PostScratch();
    std::vector<double> big((size_t)N * N + 4 * (size_t)N + 16);
    int lock = 0;
    bool active = init_slot(*st, *isc, im, out, *cfg, sphere, S, init_vp, n_init, T);
    int steps = 0;
    while (active && steps < 100000) {
        ++steps;
        if (st->run_e)
            for (int n = 0; n < N; ++n) estep_line(im, st->M, st->pv, st->vx, st->vy, st->inv2s, st->coef, n);
        if (st->run_w)
            for (int m = 0; m < st->M; ++m)
                for (int k = 0; k < N; ++k) {
                    double acc = 0.0;
                    if (cfg->use_weights)
                        for (int j = 0; j < N; ++j) acc += im.wt[wt_index(N, j, m, st->M)] * im.lsim[lsim_index(N, j, k)];
                    im.w[(size_t)m * N + k] = wmat_finish(im.wt[wt_index(N, k, m, st->M)], im.lweight[k], im.colsum[k], acc, cfg->wbias);
                }
        post_slot(*st, *sc, im, out, *cfg, big.data(), big.size(), &lock, T);
        active = !st->done;
    }
    *supersteps = steps;
    delete st; delete isc; delete sc;
    return active ? 1 : 0;
}

// One refit (E7, calc_new_vanishing_point): unit lines `ln` (N,3), weights `w` (N) -> vp (3).
// Returns 1 if a VP was produced, 0 otherwise; *refined = 1 if the ill-conditioned path ran.
extern "C" int hostsim_refit(const double* ln, const double* w, int N, double* vp, int* refined) {
    const Team T = make_team();
    std::vector<double> ws(slot_doubles(N) + 16, 0.0);
    std::vector<double> segs(4 * (size_t)N + 4, 0.0);
    Img im = make_img(N, ws.data(), segs.data());
    for (int n = 0; n < N; ++n) {
        for (int k = 0; k < 3; ++k) im.ln[3 * (size_t)n + k] = ln[3 * (size_t)n + k];
        im.w[n] = w[n];
        im.pvl[n] = 1.0; im.lvsq[n] = 0.0;
    }
    std::vector<RefitAcc> acc(kMaxM);
    refit_sums(im, im.w, nullptr, -1, 0, -1, acc[0], T);
    int any_refine = 0;
    refit_finish(acc.data(), 1, im, im.w, (size_t)N, nullptr, false, any_refine, T);
    for (int k = 0; k < 3; ++k) vp[k] = acc[0].nv[k];
    *refined = acc[0].refine;
    return acc[0].ok;
}
