#include "vpk_internal.cuh"
namespace vpk { void em_free(vpk_ctx*) {} }
extern "C" {
void vpk_em_default_config(vpk_em_config* c) {
    if (!c) return;
    c->num_iter = 100; c->num_init_vp = 25; c->split_merge_freq = 10; c->num_min_lines = 3;
    c->do_merge = 1; c->do_split = 1; c->do_iterations = 1; c->use_weights = 1;
    c->wbias = 1.0; c->merge_thresh = 1e-3; c->outlier_thresh = 1.96 * 1.96; c->final_convergence = 5e-3;
    c->s_thresh = 1e-200;
}
int vpk_em(vpk_ctx*, const double*, const double*, const int32_t*, int32_t, const double*, const uint8_t*, int32_t,
           const double*, const int32_t*, const vpk_em_config*, vpk_em_result*) { vpk::set_error("vpk_em: not built yet"); return VPK_ERR_STATE; }
}
