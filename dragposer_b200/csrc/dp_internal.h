// dp_internal.h -- declarations shared by the engine translation units.
#pragma once
#include "dp_common.cuh"
#include "dp_temporal.cuh"

struct TpWork {
  float* enc;
  float* enc2;
  float* dec;
  float* dec2;
  float* dec_lat;
};

cudaError_t dp_frame_simt_launch(const DpFrameArgs& args, int num_sms, cudaStream_t stream);
cudaError_t dp_temporal_run(const float* blob, const TpLayout& L, const float* mu, const float* sigma,
                            const float* latent_buf, const float* disp_buf, const float* height_buf, int head, int B,
                            int window, float* target_buf, const TpWork& w, cudaStream_t st, long long* launches);
