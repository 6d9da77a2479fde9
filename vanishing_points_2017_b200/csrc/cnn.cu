#include "vpk_internal.cuh"
namespace vpk { void cnn_free(vpk_ctx*) {} }
extern "C" {
int vpk_cnn_load(vpk_ctx*, const float* const*, const float* const*, const float*) { vpk::set_error("vpk_cnn_load: not built yet"); return VPK_ERR_STATE; }
int vpk_cnn_forward(vpk_ctx*, const uint8_t*, int32_t, float*, float*) { vpk::set_error("vpk_cnn_forward: not built yet"); return VPK_ERR_STATE; }
}
