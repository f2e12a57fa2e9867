// Episode construction on the device: the float32 bilinear resize that the
// reference's loaders run on the host for every query image and every support crop,
//   prep_im_for_blob           lib/model/utils/blob.py:35-52     im.astype(float32) - PIXEL_MEANS, cv2.resize(fx, fy)
//   support crop (training)    lib/roi_data_layer/fs_loader.py:113-138   crop box -> cv2.resize(dsize) -> zero-pad 320x320
//   support image (inference)  lib/roi_data_layer/inference_loader.py:95-109
// with the crop, the mean subtraction, the HWC -> CHW transposition (minibatch.py / permute(0,3,1,2)) and the
// zero padding of the canvas fused into one pass.
//
// Arithmetic follows cv2.resize(INTER_LINEAR) for CV_32F sources (OpenCV resizeGeneric_/HResizeLinear/VResizeLinear):
//   fx = (float)((dx + 0.5) * scale_x - 0.5)   (the product in double),  sx = floor(fx),  fx -= sx
//   sx < 0 -> (sx, fx) = (0, 0);   sx >= W - 1 -> (sx, fx) = (W - 1, 0);   the y axis clamps the two ROW indices instead
//   row pass first:  h(y, dx) = S[y][sx] * (1 - fx) + S[y][sx + 1] * fx,   then  D = h(sy) * (1 - fy) + h(sy + 1) * fy
// every product / sum rounded separately (OpenCV's SIMD path may fuse them: differences stay below 1 ulp of a pixel).
#pragma once
#include "api_common.cuh"

namespace dana {

struct EpisodeResizeParams {
  const void* src;         // HWC, 3 channels, u8 or f32
  int src_is_f32;
  long long row_pitch;     // elements between source rows
  int crop_x, crop_y, crop_w, crop_h;   // source window the resize sees as its whole image
  double scale_x, scale_y;              // source pixels per destination pixel
  int dst_w, dst_h;                     // resized extent
  float mean[3];                        // subtracted from the source before interpolation
  float* out;                           // [3][out_h][out_w], zero outside dst_h x dst_w
  int out_h, out_w;
};

__device__ __forceinline__ float episode_px(const EpisodeResizeParams& p, int y, int x, int c) {
  const long long i = static_cast<long long>(p.crop_y + y) * p.row_pitch + static_cast<long long>(p.crop_x + x) * 3 + c;
  const float v = p.src_is_f32 ? __ldg(static_cast<const float*>(p.src) + i)
                               : static_cast<float>(__ldg(static_cast<const unsigned char*>(p.src) + i));
  return __fsub_rn(v, p.mean[c]);
}

// thread per canvas pixel; the three channel planes are written with coalesced stores
__global__ void __launch_bounds__(256) episode_resize_kernel(const EpisodeResizeParams p) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x;
  const int dy = blockIdx.y;
  if (dx >= p.out_w) return;
  const long long plane = static_cast<long long>(p.out_h) * p.out_w;
  float* o = p.out + static_cast<long long>(dy) * p.out_w + dx;
  if (dx >= p.dst_w || dy >= p.dst_h) {
    o[0] = 0.0f;
    o[plane] = 0.0f;
    o[2 * plane] = 0.0f;
    return;
  }
  float fx = static_cast<float>((dx + 0.5) * p.scale_x - 0.5);
  int sx = static_cast<int>(floorf(fx));
  fx -= static_cast<float>(sx);
  if (sx < 0) {
    fx = 0.0f;
    sx = 0;
  }
  if (sx >= p.crop_w - 1) {
    fx = 0.0f;
    sx = p.crop_w - 1;
  }
  const int sx1 = min(sx + 1, p.crop_w - 1);
  float fy = static_cast<float>((dy + 0.5) * p.scale_y - 0.5);
  const int sy = static_cast<int>(floorf(fy));
  fy -= static_cast<float>(sy);
  const int y0 = min(max(sy, 0), p.crop_h - 1);
  const int y1 = min(max(sy + 1, 0), p.crop_h - 1);
  const float ax0 = 1.0f - fx, ax1 = fx, by0 = 1.0f - fy, by1 = fy;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float h0 = __fadd_rn(__fmul_rn(episode_px(p, y0, sx, c), ax0), __fmul_rn(episode_px(p, y0, sx1, c), ax1));
    const float h1 = __fadd_rn(__fmul_rn(episode_px(p, y1, sx, c), ax0), __fmul_rn(episode_px(p, y1, sx1, c), ax1));
    o[c * plane] = __fadd_rn(__fmul_rn(h0, by0), __fmul_rn(h1, by1));
  }
}

inline int episode_resize_run(const void* src, int src_is_f32, int src_h, int src_w, long long row_pitch, int crop_x,
                              int crop_y, int crop_w, int crop_h, double scale_x, double scale_y, int dst_w, int dst_h,
                              float mean0, float mean1, float mean2, float* out, int out_h, int out_w,
                              cudaStream_t stream) {
  if (!src || !out || src_h <= 0 || src_w <= 0 || out_h <= 0 || out_w <= 0) return DANA_EINVAL;
  if (crop_w <= 0 || crop_h <= 0 || crop_x < 0 || crop_y < 0 || crop_x + crop_w > src_w || crop_y + crop_h > src_h)
    return DANA_EINVAL;
  if (dst_w < 0 || dst_h < 0 || dst_w > out_w || dst_h > out_h || !(scale_x > 0.0) || !(scale_y > 0.0)) return DANA_EINVAL;
  if (row_pitch < 3LL * src_w) return DANA_EINVAL;
  EpisodeResizeParams p;
  p.src = src;
  p.src_is_f32 = src_is_f32;
  p.row_pitch = row_pitch;
  p.crop_x = crop_x;
  p.crop_y = crop_y;
  p.crop_w = crop_w;
  p.crop_h = crop_h;
  p.scale_x = scale_x;
  p.scale_y = scale_y;
  p.dst_w = dst_w;
  p.dst_h = dst_h;
  p.mean[0] = mean0;
  p.mean[1] = mean1;
  p.mean[2] = mean2;
  p.out = out;
  p.out_h = out_h;
  p.out_w = out_w;
  episode_resize_kernel<<<dim3((out_w + 255) / 256, out_h), 256, 0, stream>>>(p);
  DANA_LAUNCH_CHECK();
  return DANA_OK;
}

}  // namespace dana
