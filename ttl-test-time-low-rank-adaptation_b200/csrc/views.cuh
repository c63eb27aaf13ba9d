// Host-side interface of the GPU view generator (views.cu).
#pragma once
#include "../../include/ttl_b200.h"
#include "kernels.cuh"

namespace ttl {

// One view, resolved on the host from a ttl_view_spec: source window of the uint8 image, size of the resized image it is
// resampled to, and where the S x S output window sits inside that resized image.
struct ViewDesc {
  long long img_off;       // byte offset of the view's image in the image buffer
  long long coef_h_off;    // int offsets into the coefficient buffer
  long long coef_v_off;
  long long tmp_off;       // byte offset of the view's horizontal-pass rows [h][S][3]
  int W;                   // row pitch of the image in pixels
  int x0, y0, w, h;        // source window
  int ow, oh;              // output size of the resize
  int ox, oy;              // first output column / row taken (CenterCrop); 0 for the random crops
  int ksh, ksv;            // filter taps per output column / row (Pillow ksize)
  int bicubic, flip;
};

// Fills out[0..n_views) for one image and advances the running coefficient / temp sizes.  Returns nullptr or an error.
const char* views_plan(const ttl_view_spec* specs, int n_views, int H, int W, int size, long long img_off, ViewDesc* out,
                       size_t* coef_ints, size_t* tmp_bytes);

// views (fp32 [n,3,S,S]) and/or patches (bf16 [n*T, Kp]) may be null.  mean/sd: 3 floats each (host).
void launch_views(const uint8_t* img, const ViewDesc* desc, int n_views, int max_h, int* coef, uint8_t* tmp, float* views,
                  bf16* patches, int size, int p, const float* mean, const float* sd, cudaStream_t st);

}  // namespace ttl
