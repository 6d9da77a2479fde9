#include "vpk_internal.cuh"
namespace vpk { void pipe_free(vpk_ctx*) {} }
extern "C" {
int vpk_pipeline_upload(vpk_ctx*, const double*, const int32_t*, int32_t) { vpk::set_error("pipeline: not built yet"); return VPK_ERR_STATE; }
int vpk_pipeline_run(vpk_ctx*, int32_t, int32_t, double, const vpk_em_config*) { vpk::set_error("pipeline: not built yet"); return VPK_ERR_STATE; }
int vpk_pipeline_fetch(vpk_ctx*, vpk_em_result*, float*, uint8_t*) { vpk::set_error("pipeline: not built yet"); return VPK_ERR_STATE; }
int vpk_pipeline_host(vpk_ctx*, const double*, const int32_t*, int32_t, int32_t, int32_t, double, const vpk_em_config*, vpk_em_result*, float*, uint8_t*) { vpk::set_error("pipeline: not built yet"); return VPK_ERR_STATE; }
int vpk_pipeline_stage_ms(vpk_ctx*, float*) { vpk::set_error("pipeline: not built yet"); return VPK_ERR_STATE; }
}
