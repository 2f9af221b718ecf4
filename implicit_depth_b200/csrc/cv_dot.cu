// Fused plane-sweep dot-product cost volume (simple_cost_volume):
//   per depth plane: homography warp of every source view's features into the current
//   frustum (bilinear, zero padding), dot with the current features, sum over views,
//   running arg-max over planes -- one kernel, nothing materialised.
// Replaces CostVolumeManager.build_cost_volume + forward (modules/cost_volume.py:221-358)
// and EfficientCostVolumeManager.build_cost_volume (:1245-1304).
//
// Mapping: a quad of 4 lanes owns one pixel; lane j of the quad owns channels 4j..4j+3, so one
// 64-byte texel record is fetched by one LDG.128 per lane and a warp-wide load instruction
// touches 8 records = 4 full 128-byte lines when neighbouring pixels sample neighbouring
// texels (100 % sector efficiency instead of 25 % for a thread-per-pixel gather).
//
// The kernel is issue-bound, so the per-(pixel, plane, view) projection is NOT repeated by the
// four lanes of a quad: lane j projects views j and j+4 only and publishes a 32-byte "tap record"
// (four separable bilinear weights with out-of-image taps zeroed and the four clamped texel offsets)
// in shared memory; the gather loop then reads one record per view with two quad-broadcast LDS.128
// and issues four LDG.128, skipping samples that lie entirely outside the source image.  Records of plane d+1 are
// produced while plane d is gathered (two buffers, one __syncwarp per plane).
// Zero-weight taps reproduce the reference's dropped taps exactly (x + 0*s == x for finite s).
#include <stdlib.h>

#include "common.cuh"

#define CVD_THREADS 256
#define CVD_WARPS (CVD_THREADS / 32)
#define CVD_PIX_PER_BLOCK (CVD_THREADS / 4)
#define CVD_SKIP 0xffffffffu  // o00 of a sample that lies entirely outside the source image
#define CVD_VIEW_STRIDE 10  // 16-byte chunks per view row (8 pixels + 2 pad): quarter-warp stores hit 8 distinct chunks

// Separable tap record of one projected sample (ATen grid_sampler_2d semantics: bilinear, zeros padding,
// align_corners=False; geometry_utils.py:84-89 for the projection).
struct TapRec {
  float wx0, wx1, wy0, wy1;
  unsigned o00, o01, o10, o11;  // float offsets of the (clamped) nw, ne, sw, se texel records
};

__device__ __forceinline__ float fast_rcp(float z) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
  const float e = fmaf(-z, r, 1.f);  // one Newton step: <= 1 ulp
  return fmaf(r, e, r);
}

__device__ __forceinline__ TapRec project_taps(float Mp0, float Mp1, float Mp2, float t0, float t1, float t2, float zd,
                                               int h, int w) {
  const float cx = fmaf(zd, Mp0, t0), cy = fmaf(zd, Mp1, t1), cz = fmaf(zd, Mp2, t2);
  const float z = fmaxf(cz, 1e-5f);
  const float rz = fast_rcp(z);
  const float ix = fmaf(cx, rz, -0.5f), iy = fmaf(cy, rz, -0.5f);
  const float fx = floorf(ix), fy = floorf(iy);
  const float ax = ix - fx, ay = iy - fy;
  const float bx = (fx + 1.f) - ix, by = (fy + 1.f) - iy;
  // saturating conversions: behind-camera points reach |coord| ~ 1e8
  const int x0 = __float2int_rd(fminf(fmaxf(fx, -4.f), (float)(w + 4)));
  const int y0 = __float2int_rd(fminf(fmaxf(fy, -4.f), (float)(h + 4)));
  TapRec r;
  r.wx0 = ((unsigned)x0 < (unsigned)w) ? bx : 0.f;
  r.wx1 = ((unsigned)(x0 + 1) < (unsigned)w) ? ax : 0.f;
  r.wy0 = ((unsigned)y0 < (unsigned)h) ? by : 0.f;
  r.wy1 = ((unsigned)(y0 + 1) < (unsigned)h) ? ay : 0.f;
  const int xa = min(max(x0, 0), w - 1), xb = min(max(x0 + 1, 0), w - 1);
  const int ya = min(max(y0, 0), h - 1), yb = min(max(y0 + 1, 0), h - 1);
  const int ra = ya * w, rb = yb * w;
  const bool inside = ((unsigned)(x0 + 1) <= (unsigned)w) & ((unsigned)(y0 + 1) <= (unsigned)h);
  // a sample entirely outside the image (a third of all samples at the near planes) is flagged and skipped by the
  // gather loop: every LDG.128 costs the L1 data pipe (this kernel's limiter) four write-back wavefronts
  r.o00 = inside ? (unsigned)(ra + xa) * B200_FEAT_C : CVD_SKIP;
  r.o01 = (unsigned)(ra + xb) * B200_FEAT_C;
  r.o10 = (unsigned)(rb + xa) * B200_FEAT_C;
  r.o11 = (unsigned)(rb + xb) * B200_FEAT_C;
  return r;
}

template <int NR, int MB>  // rounds of 4 views: K <= 4 * NR; MB = resident blocks per SM the registers are capped for
__global__ void __launch_bounds__(CVD_THREADS, MB)
cv_dot_kernel(const float* __restrict__ cur,     // [B, N, 16] pixel-major
              const float* __restrict__ src,     // [B, K, N, 16] pixel-major
              const float* __restrict__ cams,    // [B, K, 32]
              const float* __restrict__ planes,  // [B, D]
              float* __restrict__ cost,          // [B, D, N]
              float* __restrict__ lowest,        // [B, N] or null
              int* __restrict__ best_idx,        // [B, N] or null
              int K, int D, int h, int w) {
  __shared__ float s_cam[B200_MAX_VIEWS * B200_CAM_STRIDE];
  // tap records, structure of arrays: weights and offsets of (view k, pixel q) at chunk k * CVD_VIEW_STRIDE + q
  __shared__ float4 s_wgt[CVD_WARPS][2][4 * NR * CVD_VIEW_STRIDE];
  __shared__ uint4 s_off[CVD_WARPS][2][4 * NR * CVD_VIEW_STRIDE];
  extern __shared__ float s_planes[];
  const int N = h * w;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < K * B200_CAM_STRIDE; i += CVD_THREADS)
    s_cam[i] = cams[(size_t)b * K * B200_CAM_STRIDE + i];
  for (int i = threadIdx.x; i < D; i += CVD_THREADS) s_planes[i] = planes[b * D + i];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane >> 2;  // pixel slot in warp
  const int j = lane & 3;   // channel quarter / projected view within a round
  // a block covers an 8 x 8 pixel tile (warp = tile row, quad = pixel in the row): vertically adjacent pixels share
  // source texel rows, so the tile's taps hit in this SM's L1 instead of going to L2 from eight different SMs
  const int tiles_x = (w + 7) >> 3;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  int x = tx * 8 + q, y = ty * CVD_WARPS + warp;
  const bool live = (x < w) & (y < h);
  x = min(x, w - 1);  // dead lanes shadow a border pixel (keeps the quad convergent); result discarded
  y = min(y, h - 1);
  const int p = y * w + x;
  const float pxc = x + 0.5f, pyc = y + 0.5f;  // pixel centres, geometry_utils.py:39

  // per-(pixel, view) constants of the views this lane projects: Mp = M @ (x+.5, y+.5, 1), t = P[:, 3]
  float Mp[NR][3], tp[NR][3];
  bool act[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int k = 4 * r + j;
    act[r] = k < K;
    const float* cam = s_cam + (act[r] ? k : 0) * B200_CAM_STRIDE;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      Mp[r][i] = fmaf(cam[CAM_M + 3 * i], pxc, fmaf(cam[CAM_M + 3 * i + 1], pyc, cam[CAM_M + 3 * i + 2]));
      tp[r][i] = cam[CAM_P + 4 * i + 3];
    }
  }

  const float4 c4 = ldg4(cur + ((size_t)b * N + p) * B200_FEAT_C + 4 * j);
  const float* srcb = src + (size_t)b * K * N * B200_FEAT_C + 4 * j;  // lane base; record offsets stay 32-bit

  auto publish = [&](int d) {
    const float zd = s_planes[d];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      if (act[r]) {
        const TapRec t = project_taps(Mp[r][0], Mp[r][1], Mp[r][2], tp[r][0], tp[r][1], tp[r][2], zd, h, w);
        const int slot = (4 * r + j) * CVD_VIEW_STRIDE + q;
        s_wgt[warp][d & 1][slot] = make_float4(t.wx0, t.wx1, t.wy0, t.wy1);
        s_off[warp][d & 1][slot] = make_uint4(t.o00, t.o01, t.o10, t.o11);
      }
    }
  };

  publish(0);
  __syncwarp();
  float best = 0.f;
  int bi = 0;
  float keep = 0.f;  // lane j keeps the cost of plane d with (d & 3) == j for a 4-plane store
  for (int d = 0; d < D; ++d) {
    if (d + 1 < D) publish(d + 1);
    float acc = 0.f;
    const float* sk = srcb;
    for (int k = 0; k < K; ++k, sk += (size_t)N * B200_FEAT_C) {
      const float4 wv = s_wgt[warp][d & 1][k * CVD_VIEW_STRIDE + q];
      const uint4 ov = s_off[warp][d & 1][k * CVD_VIEW_STRIDE + q];
      if (ov.x != CVD_SKIP) {
        const float4 s00 = ldg4(sk + ov.x);
        const float4 s01 = ldg4(sk + ov.y);
        const float4 s10 = ldg4(sk + ov.z);
        const float4 s11 = ldg4(sk + ov.w);
        const float w00 = wv.x * wv.z, w01 = wv.y * wv.z, w10 = wv.x * wv.w, w11 = wv.y * wv.w;  // nw, ne, sw, se
        float d00 = s00.x * c4.x, d01 = s01.x * c4.x, d10 = s10.x * c4.x, d11 = s11.x * c4.x;
        d00 = fmaf(s00.y, c4.y, d00); d01 = fmaf(s01.y, c4.y, d01); d10 = fmaf(s10.y, c4.y, d10); d11 = fmaf(s11.y, c4.y, d11);
        d00 = fmaf(s00.z, c4.z, d00); d01 = fmaf(s01.z, c4.z, d01); d10 = fmaf(s10.z, c4.z, d10); d11 = fmaf(s11.z, c4.z, d11);
        d00 = fmaf(s00.w, c4.w, d00); d01 = fmaf(s01.w, c4.w, d01); d10 = fmaf(s10.w, c4.w, d10); d11 = fmaf(s11.w, c4.w, d11);
        float dotk = w00 * d00;
        dotk = fmaf(w01, d01, dotk);
        dotk = fmaf(w10, d10, dotk);
        dotk = fmaf(w11, d11, dotk);
        // mask = (z > 0) is identically 1 because z is clamped to 1e-5 (cost_volume.py:216)
        acc += dotk;
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if ((d & 3) == j) keep = acc;
    if ((d & 3) == 3 || d == D - 1) {
      const int dd = (d & ~3) + j;
      if (live && dd <= d) cost[((size_t)b * D + dd) * N + p] = keep;
    }
    if (d == 0 || acc > best) {  // strict '>' keeps the first maximum (torch.argmax, :354)
      best = acc;
      bi = d;
    }
    __syncwarp();
  }
  if (live && j == 0) {
    if (lowest) lowest[(size_t)b * N + p] = s_planes[bi];
    if (best_idx) best_idx[(size_t)b * N + p] = bi;
  }
}

extern "C" int b200_cv_dot(const float* cur, const float* src, const float* cams, const float* planes, float* cost,
                           float* lowest, int* best_idx, int B, int K, int C, int h, int w, int D, void* stream) {
  B200_CHECK_ARG(C == B200_FEAT_C, "cv_dot: only %d feature channels supported (got %d)", B200_FEAT_C, C);
  B200_CHECK_ARG(B > 0 && K > 0 && K <= B200_MAX_VIEWS && D > 0 && h > 0 && w > 0,
                 "cv_dot: bad sizes B=%d K=%d D=%d h=%d w=%d", B, K, D, h, w);
  // static shared memory (~42 KB for K > 4) + D floats of plane depths must fit the 48 KB default carve-out
  B200_CHECK_ARG(D <= 1024, "cv_dot: at most 1024 depth planes (got %d)", D);
  B200_CHECK_ARG((long long)h * w * B200_FEAT_C < (1ll << 30), "cv_dot: feature map too large (%d x %d)", h, w);
  B200_CHECK_ARG(cur && src && cams && planes && cost, "cv_dot: null pointer");
  B200_CHECK_ARG((((uintptr_t)cur | (uintptr_t)src) & 15) == 0, "cv_dot: feature pointers must be 16-byte aligned");
  const int N = h * w;
  dim3 grid(((w + 7) / 8) * ((h + CVD_WARPS - 1) / CVD_WARPS), B);
  cudaStream_t st = (cudaStream_t)stream;
#define CVD_LAUNCH(NR, MB) \
  cv_dot_kernel<NR, MB><<<grid, CVD_THREADS, D * sizeof(float), st>>>(cur, src, cams, planes, cost, lowest, best_idx, K, D, h, w)
  if (K <= 4) CVD_LAUNCH(1, 3); else CVD_LAUNCH(2, 3);  // 3 resident blocks per SM: measured optimum
  B200_CHECK_LAUNCH("cv_dot");
  return 0;
}
