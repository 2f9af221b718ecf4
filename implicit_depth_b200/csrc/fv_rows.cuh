// Per-sample input rows of the feature-volume MLP (mlp_feature_volume), shared by the SIMT
// and the tcgen05 kernels.
//
// The reference concatenates 26K+20 channels per (pixel, plane) in the order of
// modules/cost_volume.py:681-695.  Channels that do not depend on (pixel, plane) -- the K
// mask channels and the 3K pose-distance channels -- are folded into a per-frame bias by
// volume_prepare_kernel, and the remaining 22K+20 are re-ordered view-major (the first
// layer's weight columns are permuted to match when the weights are packed):
//
//   view block k (22 ch): warped_k[0..15] | z'_k | dot_k | angle_k | srcray_k[0..2]
//   tail block  (20 ch): cur[0..15] | curray[0..2] | z_d
#pragma once
#include "common.cuh"

#define FV_VIEW_CH 22
#define FV_TAIL_CH 20

struct PixelCtx {
  float pxc, pyc;     // pixel centre (geometry_utils.py:39)
  float ray[3];       // r = invK[:3,:3] @ (pxc, pyc, 1)   (geometry_utils.py:60)
  float curray[3];    // normalize(r): the current-view ray is plane-invariant (cost_volume.py:618)
  float inv_n1;       // 1 / max(|curray|, 1e-5): the current-ray factor of F.cosine_similarity (cost_volume.py:657)
};

__device__ __forceinline__ PixelCtx make_pixel_ctx(int x, int y, const float* __restrict__ invK /*3x3*/) {
  PixelCtx c;
  c.pxc = x + 0.5f;
  c.pyc = y + 0.5f;
#pragma unroll
  for (int r = 0; r < 3; ++r) c.ray[r] = fmaf(invK[3 * r], c.pxc, fmaf(invK[3 * r + 1], c.pyc, invK[3 * r + 2]));
  float n = sqrtf(c.ray[0] * c.ray[0] + c.ray[1] * c.ray[1] + c.ray[2] * c.ray[2]);
  float s = 1.f / fmaxf(n, 1e-12f);  // F.normalize eps
#pragma unroll
  for (int r = 0; r < 3; ++r) c.curray[r] = c.ray[r] * s;
  c.inv_n1 = 1.f / fmaxf(sqrtf(c.curray[0] * c.curray[0] + c.curray[1] * c.curray[1] + c.curray[2] * c.curray[2]), 1e-5f);
  return c;
}

// One 22-channel view block.  `srck` points at this view's quarter-planar features [4,N,4];
// `cur` are the 16 current-view channels of the pixel.  Returns whether the sample lies in
// the (2, w-2) x (2, h-2) window of get_mask (cost_volume.py:75-96).
__device__ __forceinline__ bool fv_view_block(const PixelCtx& pc, const float* __restrict__ cam,
                                              const float* __restrict__ srck, const float* cur, float zd, int h,
                                              int w, float out[FV_VIEW_CH]) {
  float Mp[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    Mp[r] = fmaf(cam[CAM_M + 3 * r], pc.pxc, fmaf(cam[CAM_M + 3 * r + 1], pc.pyc, cam[CAM_M + 3 * r + 2]));
  float px, py, z;
  project_plane(Mp, cam, zd, px, py, z);
  const Taps t = make_taps(px, py, h, w);
  const int n4 = h * w * FEAT_Q;
#pragma unroll
  for (int c = 0; c < 16; ++c) out[c] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (t.idx[i] >= 0) {
      const float* tp = srck + (size_t)t.idx[i] * FEAT_Q;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float4 s4 = ldg4(tp + (size_t)v * n4);  // plane v of the quarter-planar layout
        out[4 * v + 0] = fmaf(t.wgt[i], s4.x, out[4 * v + 0]);
        out[4 * v + 1] = fmaf(t.wgt[i], s4.y, out[4 * v + 1]);
        out[4 * v + 2] = fmaf(t.wgt[i], s4.z, out[4 * v + 2]);
        out[4 * v + 3] = fmaf(t.wgt[i], s4.w, out[4 * v + 3]);
      }
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < 16; ++c) dot = fmaf(out[c], cur[c], dot);
  // source ray: normalize(X - t_k), X = z_d * r   (geometry_utils.py:174-178, cost_volume.py:1085-1096)
  float vx = fmaf(zd, pc.ray[0], -cam[CAM_T + 0]);
  float vy = fmaf(zd, pc.ray[1], -cam[CAM_T + 1]);
  float vz = fmaf(zd, pc.ray[2], -cam[CAM_T + 2]);
  float n = sqrtf(vx * vx + vy * vy + vz * vz);
  float s = 1.f / fmaxf(n, 1e-12f);
  vx *= s;
  vy *= s;
  vz *= s;
  // F.cosine_similarity(eps=1e-5) of two already-normalised rays (cost_volume.py:657-659)
  float n1 = fmaxf(sqrtf(pc.curray[0] * pc.curray[0] + pc.curray[1] * pc.curray[1] + pc.curray[2] * pc.curray[2]), 1e-5f);
  float n2 = fmaxf(sqrtf(vx * vx + vy * vy + vz * vz), 1e-5f);
  float ang = (pc.curray[0] / n1) * (vx / n2) + (pc.curray[1] / n1) * (vy / n2) + (pc.curray[2] / n1) * (vz / n2);
  out[16] = z;    // clamped depth in the source view (cost_volume.py:589-594)
  out[17] = dot;  // dot * mask, mask == 1 (cost_volume.py:662-668)
  out[18] = ang;
  out[19] = vx;
  out[20] = vy;
  out[21] = vz;
  return (px > 2.f) & (px < (float)(w - 2)) & (py > 2.f) & (py < (float)(h - 2));
}

// ---- fast variant used by the tensor-core kernel -------------------------------------------------------------
// Same channels; reciprocals and reciprocal square roots through the SFU (rcp/rsqrt.approx + one Newton step,
// <= 1-2 ulp) instead of IEEE division/sqrt sequences: the row threads of fv_tc_kernel are issue/latency bound and
// the six divisions + three square roots of the strict version cost ~100 instructions per (row, view).
__device__ __forceinline__ float fv_rcp(float z) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
  return fmaf(r, fmaf(-z, r, 1.f), r);
}
__device__ __forceinline__ float fv_rsqrt(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * fmaf(-0.5f * x * r, r, 1.5f);  // one Newton step
}

__device__ __forceinline__ bool fv_view_block_fast(const PixelCtx& pc, const float* __restrict__ cam,
                                                   const float* __restrict__ srck, const float* cur, float zd, int h,
                                                   int w, bool exact_div, float out[FV_VIEW_CH]) {
  float Mp[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    Mp[r] = fmaf(cam[CAM_M + 3 * r], pc.pxc, fmaf(cam[CAM_M + 3 * r + 1], pc.pyc, cam[CAM_M + 3 * r + 2]));
  const float cx = fmaf(zd, Mp[0], cam[CAM_P + 3]);
  const float cy = fmaf(zd, Mp[1], cam[CAM_P + 7]);
  const float cz = fmaf(zd, Mp[2], cam[CAM_P + 11]);
  const float z = fmaxf(cz, 1e-5f);  // geometry_utils.py:84-89
  float px, py;
  if (exact_div) {  // tile-uniform: the plane whose in-bounds test becomes overall_mask keeps the IEEE division
    px = cx / z;
    py = cy / z;
  } else {
    const float rz = fv_rcp(z);
    px = cx * rz;
    py = cy * rz;
  }
  const Taps t = make_taps(px, py, h, w);
  const int n4 = h * w * FEAT_Q;
#pragma unroll
  for (int c = 0; c < 16; ++c) out[c] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (t.idx[i] >= 0) {
      const float* tp = srck + (size_t)t.idx[i] * FEAT_Q;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float4 s4 = ldg4(tp + (size_t)v * n4);  // plane v of the quarter-planar layout
        out[4 * v + 0] = fmaf(t.wgt[i], s4.x, out[4 * v + 0]);
        out[4 * v + 1] = fmaf(t.wgt[i], s4.y, out[4 * v + 1]);
        out[4 * v + 2] = fmaf(t.wgt[i], s4.z, out[4 * v + 2]);
        out[4 * v + 3] = fmaf(t.wgt[i], s4.w, out[4 * v + 3]);
      }
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < 16; ++c) dot = fmaf(out[c], cur[c], dot);
  // source ray: normalize(X - t_k), X = z_d * r   (geometry_utils.py:174-178, cost_volume.py:1085-1096)
  float vx = fmaf(zd, pc.ray[0], -cam[CAM_T + 0]);
  float vy = fmaf(zd, pc.ray[1], -cam[CAM_T + 1]);
  float vz = fmaf(zd, pc.ray[2], -cam[CAM_T + 2]);
  const float s = fv_rsqrt(fmaxf(vx * vx + vy * vy + vz * vz, 1e-24f));  // 1 / max(|v|, 1e-12)
  vx *= s;
  vy *= s;
  vz *= s;
  // F.cosine_similarity(eps=1e-5) of two already-normalised rays (cost_volume.py:657-659)
  const float inv_n2 = fv_rsqrt(fmaxf(vx * vx + vy * vy + vz * vz, 1e-10f));
  const float ang = (pc.curray[0] * vx + pc.curray[1] * vy + pc.curray[2] * vz) * (pc.inv_n1 * inv_n2);
  out[16] = z;    // clamped depth in the source view (cost_volume.py:589-594)
  out[17] = dot;  // dot * mask, mask == 1 (cost_volume.py:662-668)
  out[18] = ang;
  out[19] = vx;
  out[20] = vy;
  out[21] = vz;
  return (px > 2.f) & (px < (float)(w - 2)) & (py > 2.f) & (py < (float)(h - 2));
}
