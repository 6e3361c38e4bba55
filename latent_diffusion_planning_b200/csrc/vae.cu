// VAE encoder handle (placeholder until the conv2d path lands): the entry points exist so that the ABI is complete
// and fail loudly with LDP_ERR_UNSUPPORTED.
#include "net_common.h"

using namespace ldp;

struct LdpVae {
  LdpVaeConfig cfg;
};

extern "C" {

int64_t ldp_vae_param_count(const LdpVaeConfig* cfg) {
  (void)cfg;
  return -1;
}

int ldp_vae_create(const LdpVaeConfig* cfg, const float* params_host, uint64_t n_params, LdpVae** out) {
  (void)cfg; (void)params_host; (void)n_params; (void)out;
  set_last_error("ldp_vae_create: VAE encoder path not built yet");
  return LDP_ERR_UNSUPPORTED;
}

int ldp_vae_destroy(LdpVae* h) {
  delete h;
  return LDP_OK;
}

int ldp_vae_encode(LdpVae* h, int precision, const void* images_dev, int pixel_format, int B, float lat_min, float lat_max,
                   float* latent_dev, void* cuda_stream) {
  (void)h; (void)precision; (void)images_dev; (void)pixel_format; (void)B; (void)lat_min; (void)lat_max; (void)latent_dev;
  (void)cuda_stream;
  set_last_error("ldp_vae_encode: VAE encoder path not built yet");
  return LDP_ERR_UNSUPPORTED;
}

}  // extern "C"
