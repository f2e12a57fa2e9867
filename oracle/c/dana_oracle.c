/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or called by the product path
 * (dual-awareness-attention-for-few-shot-object-detection_b200/); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker.
 *
 * Plain-C restatement of the reference's CPU operators for the DAnA forward hot path:
 *   oracle_nms             <- lib/model/csrc/cpu/nms_cpu.cpp:6-67
 *   oracle_roi_align_fwd   <- lib/model/csrc/cpu/ROIAlign_cpu.cpp:18-219
 * Pinned against the reference's own compiled extension (oracle/_ref, built by oracle/build_ref.py)
 * in tests/test_oracle_pins.py, and against the golden vectors in tests/golden/.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared  (no FMA contraction: the reference x86 build
 * rounds every fp32 operation separately).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- NMS ------------------------------------------------------------------------------
 * nms_cpu.cpp:22     areas = (x2 - x1 + 1) * (y2 - y1 + 1)
 * nms_cpu.cpp:24     order = scores.sort(descending)   [torch sort is not guaranteed stable; this
 *                    restatement breaks ties by lower input index, and so does the CUDA path]
 * nms_cpu.cpp:37-62  greedy suppression, IoU >= threshold
 * nms_cpu.cpp:64     return nonzero(suppressed == 0)   -> kept input indices, ascending
 * Returns the number of kept boxes, indices in keep[]. */
typedef struct {
  float score;
  int64_t idx;
} oracle_key;

static int cmp_key_desc(const void* a, const void* b) {
  const oracle_key* ka = (const oracle_key*)a;
  const oracle_key* kb = (const oracle_key*)b;
  if (ka->score > kb->score) return -1;
  if (ka->score < kb->score) return 1;
  return (ka->idx < kb->idx) ? -1 : (ka->idx > kb->idx);
}

int64_t oracle_nms(const float* boxes, const float* scores, int64_t n, float threshold, int64_t* keep) {
  if (n <= 0) return 0;
  float* areas = (float*)malloc(sizeof(float) * (size_t)n);
  uint8_t* suppressed = (uint8_t*)calloc((size_t)n, 1);
  oracle_key* order = (oracle_key*)malloc(sizeof(oracle_key) * (size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    const float* b = boxes + 4 * i;
    areas[i] = (b[2] - b[0] + 1.0f) * (b[3] - b[1] + 1.0f);
    order[i].score = scores ? scores[i] : (float)(n - i);
    order[i].idx = i;
  }
  qsort(order, (size_t)n, sizeof(oracle_key), cmp_key_desc);
  for (int64_t _i = 0; _i < n; ++_i) {
    const int64_t i = order[_i].idx;
    if (suppressed[i]) continue;
    const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
    const float iarea = areas[i];
    for (int64_t _j = _i + 1; _j < n; ++_j) {
      const int64_t j = order[_j].idx;
      if (suppressed[j]) continue;
      const float xx1 = ix1 > boxes[4 * j] ? ix1 : boxes[4 * j];
      const float yy1 = iy1 > boxes[4 * j + 1] ? iy1 : boxes[4 * j + 1];
      const float xx2 = ix2 < boxes[4 * j + 2] ? ix2 : boxes[4 * j + 2];
      const float yy2 = iy2 < boxes[4 * j + 3] ? iy2 : boxes[4 * j + 3];
      float w = xx2 - xx1 + 1.0f;
      float h = yy2 - yy1 + 1.0f;
      if (w < 0.0f) w = 0.0f;
      if (h < 0.0f) h = 0.0f;
      const float inter = w * h;
      const float ovr = inter / (iarea + areas[j] - inter);
      if (ovr >= threshold) suppressed[j] = 1;
    }
  }
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i)
    if (!suppressed[i]) keep[m++] = i;
  free(areas);
  free(suppressed);
  free(order);
  return m;
}

/* ---- RoIAlign forward -----------------------------------------------------------------
 * ROIAlign_cpu.cpp:144-148  roi * spatial_scale, no rounding
 * :155-158  roi_w/h = max(end - start, 1); bin = roi / pooled
 * :161-165  grid = sampling_ratio > 0 ? sampling_ratio : ceil(roi / pooled)
 * :36-47    sample = start + p*bin + (i + .5) * bin / grid
 * :51-96    outside [-1, size] -> 0 ; clamp to 0 ; low >= size-1 -> low = high = size-1
 * :203-213  mean over the grid of the 4-tap bilinear value
 * input [B,C,H,W], rois [R,5], out [R,C,ph,pw]; all fp32. */
void oracle_roi_align_fwd(const float* input, const float* rois, int64_t num_rois, int channels, int height,
                          int width, int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio,
                          float* out) {
  for (int64_t n = 0; n < num_rois; ++n) {
    const float* roi = rois + 5 * n;
    const int batch_ind = (int)roi[0];
    const float start_w = roi[1] * spatial_scale;
    const float start_h = roi[2] * spatial_scale;
    const float end_w = roi[3] * spatial_scale;
    const float end_h = roi[4] * spatial_scale;
    float roi_w = end_w - start_w;
    float roi_h = end_h - start_h;
    if (roi_w < 1.0f) roi_w = 1.0f;
    if (roi_h < 1.0f) roi_h = 1.0f;
    const float bin_h = roi_h / (float)pooled_h;
    const float bin_w = roi_w / (float)pooled_w;
    const int grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceil(roi_h / pooled_h);
    const int grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceil(roi_w / pooled_w);
    const float count = (float)(grid_h * grid_w);
    for (int c = 0; c < channels; ++c) {
      const float* plane = input + ((int64_t)batch_ind * channels + c) * height * width;
      for (int ph = 0; ph < pooled_h; ++ph) {
        for (int pw = 0; pw < pooled_w; ++pw) {
          float acc = 0.0f;
          for (int iy = 0; iy < grid_h; ++iy) {
            const float yy = start_h + ph * bin_h + (float)(iy + .5f) * bin_h / (float)grid_h;
            for (int ix = 0; ix < grid_w; ++ix) {
              const float xx = start_w + pw * bin_w + (float)(ix + .5f) * bin_w / (float)grid_w;
              float x = xx, y = yy;
              if (y < -1.0 || y > height || x < -1.0 || x > width) continue;
              if (y <= 0) y = 0;
              if (x <= 0) x = 0;
              int y_low = (int)y, x_low = (int)x, y_high, x_high;
              if (y_low >= height - 1) {
                y_high = y_low = height - 1;
                y = (float)y_low;
              } else {
                y_high = y_low + 1;
              }
              if (x_low >= width - 1) {
                x_high = x_low = width - 1;
                x = (float)x_low;
              } else {
                x_high = x_low + 1;
              }
              const float ly = y - y_low, lx = x - x_low;
              const float hy = (float)(1. - ly), hx = (float)(1. - lx);
              const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
              acc += w1 * plane[y_low * width + x_low] + w2 * plane[y_low * width + x_high] +
                     w3 * plane[y_high * width + x_low] + w4 * plane[y_high * width + x_high];
            }
          }
          out[(((int64_t)n * channels + c) * pooled_h + ph) * pooled_w + pw] = acc / count;
        }
      }
    }
  }
}
