/* vpk.h -- C ABI of the B200-native vanishing-point hot path (libvpk.so).
 *
 * Drop-in boundary for the lines -> sphere image -> CNN -> EM -> VPs path of
 * fkluger/vanishing_points_2017.  Each entry point names the reference
 * interface it replaces (file:line in the reference repository).  Plain
 * pointers and sizes only; every function returns an int status (VPK_OK == 0)
 * and never throws.  Unless a parameter is documented as a device pointer,
 * buffers are HOST memory owned by the caller; the library stages them through
 * device workspaces owned by the context.
 *
 * Ragged batches: image b owns rows offsets[b] .. offsets[b+1]-1 of the flat
 * (sumN, 4) segment / (sumN, 3) line arrays (row-major float64, the dtype the
 * reference's numpy arrays carry).
 */
#ifndef VPK_H
#define VPK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VPK_API __attribute__((visibility("default")))
#else
#define VPK_API
#endif

#define VPK_ABI_VERSION 1
#define VPK_MAX_VP 64   /* capacity of the per-image VP arrays in vpk_em_result   */
#define VPK_GRID 20     /* CNN response grid, cnn/deploy.prototxt:283-296         */
#define VPK_CNN_SIZE 500 /* CNN input side, cnn/deploy.prototxt:7                 */

/* function status */
enum { VPK_OK = 0, VPK_ERR_ARG = 1, VPK_ERR_CUDA = 2, VPK_ERR_STATE = 3, VPK_ERR_NOMEM = 4 };

/* per-image EM status (the reference returns a dict of None instead,
 * vp_localisation.py:205-206, 258-260, 402-404; no initial VP is an unhandled
 * np.vstack([]) ValueError at :165) */
enum { VPK_EM_OK = 0, VPK_EM_NO_INITIAL_VPS = 1, VPK_EM_NO_VPS_LEFT = 2, VPK_EM_CAPACITY = 3 };

/* sphere-mapping formulation */
enum { VPK_SPHERE_VOTES = 0,   /* pairwise-intersection vote histogram (north_star) */
       VPK_SPHERE_CURVES = 1   /* reference geometry: great-circle coverage raster   */ };

typedef struct vpk_ctx vpk_ctx;

/* ---- context ------------------------------------------------------------ */
/* replaces caffe.set_mode_gpu()/set_device(gpu_id), evaluation.py:20-21 */
VPK_API int vpk_create(int device, vpk_ctx** out);
VPK_API int vpk_destroy(vpk_ctx* ctx);
VPK_API int vpk_abi_version(void);
/* message of the last failing call on this thread ("" if none) */
VPK_API const char* vpk_last_error(void);
/* block until all work queued by this context has finished */
VPK_API int vpk_synchronize(vpk_ctx* ctx);
/* number of kernels this context has launched so far */
VPK_API int64_t vpk_launch_count(const vpk_ctx* ctx);

/* Device timestamps: vpk_mark records mark `slot` (0..3) on the context's stream; vpk_mark_elapsed returns the device
 * time from one recorded mark to another, which may belong to another context of the same device (several contexts,
 * each driven by its own host thread, keep several batches in flight; see pipeline.StreamedPipeline). */
VPK_API int vpk_mark(vpk_ctx* ctx, int32_t slot);
VPK_API int vpk_mark_elapsed(vpk_ctx* from, int32_t slot_from, vpk_ctx* to, int32_t slot_to, float* ms);

/* per-kernel device timing (CUDA events on the context's stream). */
VPK_API int vpk_profile_enable(vpk_ctx* ctx, int enable);
VPK_API int vpk_profile_reset(vpk_ctx* ctx);
/* Returns the number of distinct kernels seen; fills up to `cap` entries. */
VPK_API int vpk_profile_read(vpk_ctx* ctx, int cap, const char** names, double* total_ms, int64_t* launches);

/* ---- S0: line construction ---------------------------------------------- */
/* replaces the per-segment loop of evaluation.py:158-168 (and :199-209):
 * lines[n] = [x1,y1,1] x [x2,y2,1]. */
VPK_API int vpk_lines_from_segments(vpk_ctx* ctx, const double* segments, int64_t n, double* lines_out);

/* ---- S1: sphere mapping ------------------------------------------------- */
/* replaces sphere_mapping.sphere_line_plot (sphere_mapping.py:36-72) as
 * reached through evaluation.get_sphere_image (evaluation.py:12-14).
 *  mode VOTES : hist_out[b] (S*S uint32) = number of pairwise intersections
 *               per cell, or, with `weights` (sumN float64, nullable),
 *               whist_out[b] (S*S float32) = sum of Q.16-quantised w_i*w_j;
 *               image_out[b] = floor(255*h/max h).
 *  mode CURVES: hist_out[b] = number of lines whose sampled great circle covers
 *               the pixel; image_out[b] = floor(255*(1-(1-alpha)^k)).
 * Row 0 of every S*S plane is beta = +pi/2 (image orientation of the
 * reference canvas); column 0 is alpha = -pi/2.  Any output may be NULL. */
VPK_API int vpk_sphere_map(vpk_ctx* ctx, const double* lines, const int32_t* offsets, int32_t n_images,
                   int32_t size, int32_t mode, double alpha, const double* weights,
                   uint32_t* hist_out, float* whist_out, uint8_t* image_out);

/* ---- C0/C1: CNN --------------------------------------------------------- */
/* replaces evaluation.init_caffe (evaluation.py:17-22) + read_mean_blob
 * (:25-31).  weights[k]/biases[k], k = 0..7, are float32 host arrays in Caffe
 * blob layout for conv1..conv5, fc6, fc7, fc8_20x20 of cnn/deploy.prototxt
 * ((out, in/group, kh, kw) resp. (out, in) row-major); mean is 500*500
 * float32 or NULL (zeros). */
VPK_API int vpk_cnn_load(vpk_ctx* ctx, const float* const* weights, const float* const* biases, const float* mean);
/* replaces evaluation.caffe_forward (evaluation.py:34-38) for a batch:
 * images (n, 500, 500) uint8 -> sigout (n, 20, 20) float32; logits_out
 * (fc8 before the sigmoid) may be NULL. */
VPK_API int vpk_cnn_forward(vpk_ctx* ctx, const uint8_t* images, int32_t n_images, float* sigout, float* logits_out);

/* Diagnostics: one plain GEMM out = act(a * b^T + bias) through the tcgen05 kernel every
 * CNN layer uses.  a (m,k) and b (n,k) are bf16 bit patterns, out (m,n) float32;
 * k % 64 == 0, bn % 16 == 0, 16 <= bn <= 256, n % bn == 0.  ksplit > 1: the K range is dealt to
 * ksplit CTAs per output tile and a finishing pass adds the partial tiles in order (how the fully
 * connected layers run for batches of up to 256 images); ksplit = 1: one CTA per tile. */
VPK_API int vpk_debug_gemm(vpk_ctx* ctx, int32_t m, int32_t n, int32_t k, const uint16_t* a, const uint16_t* b,
                           const float* bias, int32_t relu, int32_t bn, int32_t ksplit, float* out);

/* ---- E0..E12: EM --------------------------------------------------------- */
/* keyword arguments of vp_localisation.expectation_maximisation
 * (vp_localisation.py:168-172), same names and defaults */
typedef struct vpk_em_config {
    int32_t num_iter;           /* 100  */
    int32_t num_init_vp;        /* 25   */
    int32_t split_merge_freq;   /* 10   */
    int32_t num_min_lines;      /* 3    */
    int32_t do_merge;           /* 1    */
    int32_t do_split;           /* 1    */
    int32_t do_iterations;      /* 1    */
    int32_t use_weights;        /* 1    */
    double wbias;               /* 1    */
    double merge_thresh;        /* 1e-3 */
    double outlier_thresh;      /* 1.96^2 */
    double final_convergence;   /* 5e-3 */
    double s_thresh;            /* 1e-200 */
} vpk_em_config;
VPK_API void vpk_em_default_config(vpk_em_config* cfg);

/* the reference's result dict (vp_localisation.py:441-442), flattened.
 * All pointers are caller-owned host arrays; decision_metric may be NULL. */
typedef struct vpk_em_result {
    int32_t* status;          /* (B)               VPK_EM_*                         */
    int32_t* n_vp;            /* (B)               rows of 'vp'                     */
    int32_t* iterations;      /* (B)               'iterations'                     */
    double* vp;               /* (B, VPK_MAX_VP, 3) 'vp'                            */
    double* sigma;            /* (B, VPK_MAX_VP)    'sigma'                         */
    int32_t* counts;          /* (B, VPK_MAX_VP)    'counts'                        */
    double* counts_weighted;  /* (B, VPK_MAX_VP)    'counts_weighted'               */
    int32_t* vp_assoc;        /* (sumN)             'vp_assoc' (-1 = outlier)       */
    double* decision_metric;  /* (VPK_MAX_VP*sumN) or NULL: image b's (n_vp, N_b)
                                 row-major block starts at VPK_MAX_VP*offsets[b]   */
} vpk_em_result;

/* replaces vp_localisation.expectation_maximisation (vp_localisation.py:
 * 168-450) called per image by evaluation.run_em_single (evaluation.py:
 * 344-347), for a ragged batch.  lines: (sumN,3) (normalised internally, the
 * caller's array is not mutated); segments: (sumN,4); responses: (B,20,20)
 * float64; sphere_images: (B,S,S) uint8; init_vp/init_vp_offsets: optional
 * per-image initial VPs (the reference's init_vp kwarg) or NULL. */
VPK_API int vpk_em(vpk_ctx* ctx, const double* lines, const double* segments, const int32_t* offsets,
           int32_t n_images, const double* responses, const uint8_t* sphere_images, int32_t size,
           const double* init_vp, const int32_t* init_vp_offsets,
           const vpk_em_config* cfg, vpk_em_result* out);

/* 'distribution' of the reference's result dict (vp_localisation.py:441-442): the PDF namedtuple of
 * probability_functions.py:5 that calc_probabilities (:99-120) returns for the LAST E-step, for image `image` of the
 * last vpk_em / vpk_pipeline_run call of this context.  Evaluated on the device from the planes that superstep left
 * in the EM workspace, so it must be asked for before the next EM call and only for batches that ran in one
 * workspace wave.  n_vp / n_lines: the image's 'vp' rows and lines (checked).  Caller-allocated float64 arrays:
 * p_v (n_vp), p_lv (n_lines, n_vp), p_vl (n_vp, n_lines), p_l (n_lines), lvsq (n_lines, n_vp), angles (n_vp, 2). */
VPK_API int vpk_em_distribution(vpk_ctx* ctx, int32_t image, int32_t n_vp, int32_t n_lines, double* p_v, double* p_lv,
                                double* p_vl, double* p_l, double* lvsq, double* angles);

/* Profiling runs only (vpk_profile_enable): accumulated algorithmic work of the weight-matrix
 * products (vp_localisation.py:515-524) since the last reset: out[0] = bytes of similarity
 * matrix consumed (8 N^2 per product), out[1] = flops (2 M N^2 per product), out[2] = number
 * of per-image products, out[3] = supersteps, out[4] = algorithmic bytes of the POST kernel
 * (the planes w, lvsq, pvl of every active image once per superstep: 8 (3 M N + 5 N)),
 * out[5] = of the E-step kernel (same figure: three planes written). */
VPK_API int vpk_em_stats(vpk_ctx* ctx, uint64_t out[6], int reset);
/* Profiling runs only: SM cycles the leading CTA of every image's cluster spent per phase of the fused
 * EM kernel since the last reset: out[0] E-step, [1] barrier after E, [2] weight-matrix product, [3] barrier
 * after W, [4] M-step sums spread over the cluster, [5] POST (state machine), [6] barrier + state mirror. */
VPK_API int vpk_em_phase_cycles(vpk_ctx* ctx, uint64_t out[8], int reset);

/* ---- whole path ---------------------------------------------------------- */
/* example.py:37-39 / benchmark.py:59-66 for a ragged batch, without the
 * per-image pickles: segments -> lines -> sphere image -> CNN -> EM, all
 * intermediates staying in HBM.
 *   vpk_pipeline_upload : host segments -> device-resident batch (H2D).
 *   vpk_pipeline_run    : run the path on the resident batch (no H2D).
 *   vpk_pipeline_fetch  : copy the results to host arrays (D2H).
 *   vpk_pipeline_host   : the three in sequence (the end-to-end call).      */
VPK_API int vpk_pipeline_upload(vpk_ctx* ctx, const double* segments, const int32_t* offsets, int32_t n_images);
VPK_API int vpk_pipeline_run(vpk_ctx* ctx, int32_t size, int32_t sphere_mode, double alpha, const vpk_em_config* cfg);
VPK_API int vpk_pipeline_fetch(vpk_ctx* ctx, vpk_em_result* out, float* sigout /* (B,20,20) or NULL */,
                       uint8_t* sphere_images /* (B,S,S) or NULL */);
VPK_API int vpk_pipeline_host(vpk_ctx* ctx, const double* segments, const int32_t* offsets, int32_t n_images,
                      int32_t size, int32_t sphere_mode, double alpha, const vpk_em_config* cfg,
                      vpk_em_result* out, float* sigout, uint8_t* sphere_images);
/* ---- N2: raw LSD output -> normalised segments (+ lines) ---------------------- */
/* The normalisation evaluation.detect_lsd_lines applies to lsd.detect_line_segments' rows (reference
 * evaluation.py:227-251) and the line construction of evaluation.py:158-168, for a ragged batch on the
 * device.  lsd: (sum N, ncols) float64 rows x1, y1, x2, y2, [width, p, -log10 NFA] in pixels
 * (ncols >= 4; 7 for LSD's own output); widths / heights: image sizes in pixels.  segments_out (sum N, 4):
 * origin at the image centre, divided by max(w, h) / 2, y up -- bit-identical to the numpy expressions;
 * lines_out (sum N, 3) = [x1,y1,1] x [x2,y2,1] or NULL; nfa_out (sum N) = column 6 or NULL. */
VPK_API int vpk_segments_from_lsd(vpk_ctx* ctx, const double* lsd, int32_t ncols, const int32_t* offsets, const int32_t* widths,
                                  const int32_t* heights, int32_t n_images, double* segments_out, double* lines_out,
                                  double* nfa_out);
/* vpk_pipeline_upload for raw LSD rows: they are normalised on the device into the resident batch. */
VPK_API int vpk_pipeline_upload_lsd(vpk_ctx* ctx, const double* lsd, int32_t ncols, const int32_t* offsets,
                                    const int32_t* widths, const int32_t* heights, int32_t n_images);

/* ---- N1: horizon line and orthogonal VP triplet --------------------------- */
/* calc_horizon.calculate_horizon_and_ortho_vp(em_result, maxbest=10, theta_vmin=pi/10, theta_z=pi/4)
 * (reference calc_horizon.py:19-225; callers example.py:65, benchmark.py:233) for a batch of EM
 * results.  vp (B, VPK_MAX_VP, 3) float64, counts (B, VPK_MAX_VP) int32 and n_vp (B) int32 are laid
 * out as vpk_em writes them (an image without VPs has n_vp = 0 and gets the reference's default
 * horizon, calc_horizon.py:207-212).  points (B, 5, 3) float64 receives the reference's return
 * values hP1, hP2, zVP, hVP1, hVP2 in that order; best_combo (B, 3) int32 the indices of the chosen
 * VPs (two entries and -1 when fewer than three VPs take part).
 * Optional (row N4, the evaluation of benchmark.py:247-253): with true_horizons (B,3) homogeneous
 * ground-truth horizon lines, scales (B) and heights (B) (benchmark.py's `scale` and `imageHeight`),
 * errors (B) receives max(|hP1.y - thP1.y|, |hP2.y - thP2.y|) / 2 * scale / imageHeight; all four NULL
 * otherwise. */
VPK_API int vpk_horizon(vpk_ctx* ctx, const double* vp, const int32_t* counts, const int32_t* n_vp, int32_t n_images,
                        int32_t maxbest, double theta_vmin, double theta_z, const double* true_horizons,
                        const double* scales, const double* heights, double* points, int32_t* best_combo, double* errors);
/* The same on the device-resident EM result of the last vpk_pipeline_run (no host round trip of the
 * EM result; only the 5 points + 3 indices per image come back). */
VPK_API int vpk_pipeline_horizon(vpk_ctx* ctx, int32_t maxbest, double theta_vmin, double theta_z,
                                 const double* true_horizons, const double* scales, const double* heights, double* points,
                                 int32_t* best_combo, double* errors);

/* device time of the last vpk_pipeline_run per stage: [lines+sphere, cnn, em, total] ms */
VPK_API int vpk_pipeline_stage_ms(vpk_ctx* ctx, float ms[4]);

#ifdef __cplusplus
}
#endif
#endif /* VPK_H */
