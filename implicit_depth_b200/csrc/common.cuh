// Shared definitions for the sm_100a plane-sweep kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define B200_MAX_VIEWS 8
#define B200_FEAT_C 16          // matching feature channels (options.py:138)
#define B200_CAM_STRIDE 32      // floats per (frame, view) camera record
#define B200_MLP_HID 128        // hidden width of both per-sample MLPs

// Camera record layout (floats), one per (b, k), written by volume_prepare_kernel:
//   [0..11]  P  = (K_src @ T_src<-cur)[:3, :4]   row-major     (geometry_utils.py:82-84)
//   [12..20] M  = P[:, :3] @ invK_cur[:3, :3]     row-major     (folds geometry_utils.py:60)
//   [21..23] t  = cur_T_src[:3, 3]                              (cost_volume.py:1088)
//   [24..26] pose distance, R measure, t measure                (geometry_utils.py:183-195)
//   [27..31] unused
#define CAM_P 0
#define CAM_M 12
#define CAM_T 21
#define CAM_POSE 24

// Matching features reach the volume kernels in one of two fp32 gather layouts (C = 16):
//   layout 0, texel records [image][N][16]: one 64-byte record per texel.  The quad-per-pixel gather of cv_dot
//     (lane j = channels 4j..4j+3) reads 8 neighbouring records = 512 contiguous bytes per LDG.128.
//   layout 1, quarter-planar [image][C/4][N][4]: four planes of float4 per image.  The thread-per-row gather of
//     the feature-volume kernels reads 32 neighbouring float4 of ONE plane per LDG.128 (512 contiguous bytes when
//     neighbouring pixels sample neighbouring texels, ~4 L1 wavefronts) where texel records cost it 16 lines.
// (Measured: cv_dot on layout 1 is 2x slower, fv_tc on layout 0 1.2x slower.)
#define FEAT_Q 4                       // channels per plane
#define FEAT_NQ (B200_FEAT_C / FEAT_Q)  // planes per image

// error plumbing for the C ABI: entry points return 0 or a negative code and keep a message.
extern "C" const char* b200_last_error(void);
void b200_set_error(const char* fmt, ...);

#define B200_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      b200_set_error(__VA_ARGS__);           \
      return -1;                             \
    }                                        \
  } while (0)

#define B200_CHECK_LAUNCH(name)                                                   \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      b200_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));     \
      return -2;                                                                  \
    }                                                                             \
  } while (0)

#define B200_CHECK_CUDA(expr)                                                     \
  do {                                                                            \
    cudaError_t e__ = (expr);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      b200_set_error("%s failed: %s", #expr, cudaGetErrorString(e__));            \
      return -2;                                                                  \
    }                                                                             \
  } while (0)

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float leaky(float x, float slope) { return fmaxf(x, x * slope); }  // 0 < slope < 1

// SFU exponential / reciprocal with flush-to-zero: ONE MUFU each.  (`__expf` / `__fdividef` without -use_fast_math
// compile to the non-ftz forms, whose denormal range handling costs ~5 extra predicated instructions per call -- half
// of the binary MLP's instruction stream before this.)  ex2.approx: 2^-22 relative error.
__device__ __forceinline__ float exp_fast(float v) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 1.4426950408889634f));
  return e;
}
__device__ __forceinline__ float rcp_fast(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// ELU(alpha = 1): exp through the SFU minus one: absolute error < 3e-7 on values of order one, far inside the 1e-3
// parity budget, at 4 instructions instead of expm1f's ~50.
__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : exp_fast(v) - 1.f; }

// SiLU x * sigmoid(x) (EfficientNetV2 image encoder): SFU exp and reciprocal, relative error ~1e-6.
__device__ __forceinline__ float silu(float v) { return v * rcp_fast(1.f + exp_fast(-v)); }
__device__ __forceinline__ float sigmoidf_fast(float v) { return rcp_fast(1.f + exp_fast(-v)); }

// Projection of the pixel-centre ray through plane depth zd into one source view.
// Mp = M @ (x+.5, y+.5, 1); returns source pixel coords and clamped depth
// (geometry_utils.py:84-89: z = max(c_z, 1e-5), xy / z).
__device__ __forceinline__ void project_plane(const float Mp[3], const float* __restrict__ cam, float zd,
                                              float& px, float& py, float& z) {
  float cx = fmaf(zd, Mp[0], cam[CAM_P + 3]);
  float cy = fmaf(zd, Mp[1], cam[CAM_P + 7]);
  float cz = fmaf(zd, Mp[2], cam[CAM_P + 11]);
  z = fmaxf(cz, 1e-5f);
  px = cx / z;
  py = cy / z;
}

// Bilinear tap set with ATen grid_sampler_2d semantics (bilinear, zeros padding,
// align_corners=False): ix = px - 0.5, weights from floor(), out-of-image taps dropped.
struct Taps {
  int idx[4];    // texel index y*w+x, or -1 when the tap is outside the image
  float wgt[4];  // nw, ne, sw, se
};

__device__ __forceinline__ Taps make_taps(float px, float py, int h, int w) {
  Taps t;
  float ix = px - 0.5f, iy = py - 0.5f;
  float fx = floorf(ix), fy = floorf(iy);
  float ax = ix - fx, ay = iy - fy;  // (ix - x_w), (iy - y_n)
  float bx = (fx + 1.f) - ix, by = (fy + 1.f) - iy;
  // saturating conversions: behind-camera points reach |coord| ~ 1e8
  int x0 = __float2int_rd(fminf(fmaxf(fx, -4.f), 1.0e6f));
  int y0 = __float2int_rd(fminf(fmaxf(fy, -4.f), 1.0e6f));
  int x1 = x0 + 1, y1 = y0 + 1;
  bool vx0 = (x0 >= 0) & (x0 < w), vx1 = (x1 >= 0) & (x1 < w);
  bool vy0 = (y0 >= 0) & (y0 < h), vy1 = (y1 >= 0) & (y1 < h);
  t.idx[0] = (vx0 & vy0) ? y0 * w + x0 : -1;
  t.idx[1] = (vx1 & vy0) ? y0 * w + x1 : -1;
  t.idx[2] = (vx0 & vy1) ? y1 * w + x0 : -1;
  t.idx[3] = (vx1 & vy1) ? y1 * w + x1 : -1;
  t.wgt[0] = bx * by;
  t.wgt[1] = ax * by;
  t.wgt[2] = bx * ay;
  t.wgt[3] = ax * ay;
  return t;
}
