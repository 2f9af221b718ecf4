// Fused plane-sweep dot-product cost volume, shared-memory-band version (BASELINE.json north_star: "stages feature tiles
// into shared memory via TMA ... bilinear-samples ... reduces the matching score across source views and depth
// hypotheses with warp-shuffle primitives").  Same arithmetic and the same results, bit for bit, as cv_dot_kernel
// (cv_dot.cu); what changes is where the taps come from.
//
// A block owns an 8 x 8 pixel tile of one frame (warp = tile row, quad = pixel, lane j of the quad = channels 4j..4j+3)
// and walks (super-group of 16 consecutive depth planes, source view) pairs.  For every pair warp 0 projects the tile's
// four corner pixels at the first and last plane -- a homography maps the tile to a convex quadrilateral and a pixel
// moves monotonically along its epipolar line between the two planes, so these 8 points bound every sample -- and, if
// the bounding box fits, ONE TMA box of 20 x 20 texel records (25.6 KB, fp32 x 16 channels; out-of-image texels
// zero-filled by the TMA unit = grid_sample's zeros padding) serves all 16 planes (the far, slowly moving planes: every
// staged texel is re-used by ~10 taps); otherwise the pair is split into four pieces of 4 planes, each with its own
// box.  Boxes land in a 3-stage shared-memory ring, two iterations ahead of their use.  The gather then is four LDS.128
// per (pixel, plane): no tag lookups, no per-tap clamping, 4 shared-memory wavefronts per warp instruction instead of
// the 8 L1 wavefronts of a scattered LDG.128 (the limiter of cv_dot_kernel, profiles/r02*_ncu_volume.md).
// Iterations whose box does not fit (near planes of wide baselines: > 3 texels of motion per plane) or whose corners
// fall behind the source camera gather from global memory exactly like cv_dot_kernel; iterations entirely outside
// the source image are skipped.
//
// Within a quad lane j projects plane 4g + j of the iteration (one projection per lane instead of four) and the
// quad shares (box offset, two fractions) through three shuffles per plane.  After the last view of a group two
// xor-shuffles per plane reduce the channel quarters; argmax is the strict '>' first maximum of torch.argmax
// (modules/cost_volume.py:354).
//
// Replaces CostVolumeManager.build_cost_volume + forward (modules/cost_volume.py:221-358) and
// EfficientCostVolumeManager.build_cost_volume (:1245-1304); warp = :134-219, grid_sample :192-198.
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "tc.cuh"
#include "tmap.cuh"

#define CVB_THREADS 256
#define CVB_WARPS 8
#define CVB_BW 20
#define CVB_BH 20
#define CVB_STAGES 3
#define CVB_BOX_FLOATS (CVB_BW * CVB_BH * B200_FEAT_C)
#define CVB_BOX_BYTES (CVB_BOX_FLOATS * 4)
#define CVB_G 4    // planes per piece (= lanes per quad: lane j projects plane j of a piece)
#define CVB_SG 16  // planes per super-group = accumulators per lane; one box when the whole group fits, else 4 pieces

enum { CVB_MODE_BOX = 0, CVB_MODE_GLOBAL = 1, CVB_MODE_SKIP = 2 };

struct CvbParams {
  CUtensorMap src_map;  // fp32 [B*K images][h][w][16], box {16, BW, BH, 1}, no swizzle, zero fill
  const float* cur;     // [B, N, 16] texel records
  const float* src;     // [B, K, N, 16]
  const float* cams;    // [B, K, 32]
  const float* planes;  // [B, D]
  float* cost;          // [B, D, N]
  float* lowest;        // [B, N] or null
  int* best_idx;        // [B, N] or null
  int K, D, h, w;
};

struct CvbMeta {
  int bx0, by0, mode, d0n;  // box origin, CVB_MODE_*, first plane * 32 + number of planes (4 | 16)
};

__device__ __forceinline__ float cvb_rcp(float z) {  // identical to cv_dot.cu: rcp.approx + one Newton step
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
  const float e = fmaf(-z, r, 1.f);
  return fmaf(r, e, r);
}

// sample position of pixel (Mp, t) at plane zd: fractional weights and the integer nw texel (saturated)
__device__ __forceinline__ void cvb_project(float Mp0, float Mp1, float Mp2, float t0, float t1, float t2, float zd, int h,
                                            int w, float& ax, float& ay, float& bx, float& by, int& x0, int& y0,
                                            float& cz) {
  const float cx = fmaf(zd, Mp0, t0), cy = fmaf(zd, Mp1, t1);
  cz = fmaf(zd, Mp2, t2);
  const float z = fmaxf(cz, 1e-5f);
  const float rz = cvb_rcp(z);
  const float ix = fmaf(cx, rz, -0.5f), iy = fmaf(cy, rz, -0.5f);
  const float fx = floorf(ix), fy = floorf(iy);
  ax = ix - fx;
  ay = iy - fy;
  bx = (fx + 1.f) - ix;
  by = (fy + 1.f) - iy;
  x0 = __float2int_rd(fminf(fmaxf(fx, -4.f), (float)(w + 4)));
  y0 = __float2int_rd(fminf(fmaxf(fy, -4.f), (float)(h + 4)));
}

__global__ void __launch_bounds__(CVB_THREADS, 2) cv_dot_band_kernel(const __grid_constant__ CvbParams prm) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* box = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  float* s_cam = box + CVB_STAGES * CVB_BOX_FLOATS;                 // [8][32]
  float* s_planes = s_cam + B200_MAX_VIEWS * B200_CAM_STRIDE;        // [D rounded up to 16, last plane repeated]
  const int Dp = (prm.D + CVB_SG - 1) / CVB_SG * CVB_SG;
  CvbMeta* s_meta = reinterpret_cast<CvbMeta*>(s_planes + Dp);       // [STAGES]
  uint64_t* full = reinterpret_cast<uint64_t*>(s_meta + CVB_STAGES);  // [STAGES]
  uint64_t* empty = full + CVB_STAGES;                                // [STAGES]

  const int K = prm.K, D = prm.D, h = prm.h, w = prm.w;
  const int N = h * w;
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < K * B200_CAM_STRIDE; i += CVB_THREADS)
    s_cam[i] = prm.cams[(size_t)b * K * B200_CAM_STRIDE + i];
  for (int i = threadIdx.x; i < Dp; i += CVB_THREADS) s_planes[i] = prm.planes[b * D + min(i, D - 1)];
  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&prm.src_map);
    for (int s = 0; s < CVB_STAGES; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], CVB_WARPS);
    }
    tc::mbar_fence_init();
  }
  __syncthreads();

  const int q = lane >> 2, j = lane & 3;
  const int tiles_x = (w + 7) >> 3;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  int x = tx * 8 + q, y = ty * 8 + warp;
  const bool live = (x < w) & (y < h);
  x = min(x, w - 1);  // dead lanes shadow a border pixel (keeps the quad convergent); result discarded
  y = min(y, h - 1);
  const int p = y * w + x;
  const float pxc = x + 0.5f, pyc = y + 0.5f;  // pixel centres, geometry_utils.py:39

  const float4 c4 = ldg4(prm.cur + ((size_t)b * N + p) * B200_FEAT_C + 4 * j);
  const float* srcb = prm.src + (size_t)b * K * N * B200_FEAT_C + 4 * j;

  const int n_sg = Dp / CVB_SG;  // super-groups of 16 planes = one accumulator set

  // ---- producer (warp 0, all lanes; every value below is warp-uniform) ---------------------------------------------
  // Bounding box of (view k, planes d_first..d_last) from the tile's corner pixels: lanes 0..7 = {x_lo, x_hi} x {y_lo,
  // y_hi} x {d_first, d_last}.  Returns the mode; (bx0, by0) = box origin with one texel of margin on every side.
  const int x_lo = tx * 8, x_hi = min(tx * 8 + 7, w - 1), y_lo = ty * 8, y_hi = min(ty * 8 + 7, h - 1);
  auto bound = [&](int k, int d_first, int d_last, int& bx0, int& by0) -> int {
    const float* cam = s_cam + k * B200_CAM_STRIDE;
    const int c = lane & 7;
    const float cxp = ((c & 1) ? x_hi : x_lo) + 0.5f, cyp = ((c & 2) ? y_hi : y_lo) + 0.5f;
    const float zd = s_planes[(c & 4) ? d_last : d_first];
    const float m0 = fmaf(cam[CAM_M + 0], cxp, fmaf(cam[CAM_M + 1], cyp, cam[CAM_M + 2]));
    const float m1 = fmaf(cam[CAM_M + 3], cxp, fmaf(cam[CAM_M + 4], cyp, cam[CAM_M + 5]));
    const float m2 = fmaf(cam[CAM_M + 6], cxp, fmaf(cam[CAM_M + 7], cyp, cam[CAM_M + 8]));
    float ax, ay, bx, by, cz;
    int x0, y0;
    cvb_project(m0, m1, m2, cam[CAM_P + 3], cam[CAM_P + 7], cam[CAM_P + 11], zd, h, w, ax, ay, bx, by, x0, y0, cz);
    int mnx = x0, mxx = x0, mny = y0, mxy = y0;
    int front = cz > 1e-4f ? 1 : 0;  // every corner well in front of the source camera: the bound holds
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
      mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
      mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
      mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
      front &= __shfl_xor_sync(0xffffffffu, front, o);
    }
    bx0 = mnx - 1;
    by0 = mny - 1;
    // needed texels: [mn - 1, mx + 2] (margin, + 1 for the se tap)
    if (!front || (mxx + 2 - bx0 + 1) > CVB_BW || (mxy + 2 - by0 + 1) > CVB_BH) return CVB_MODE_GLOBAL;
    if (mxx + 2 < 0 || mxy + 2 < 0 || bx0 >= w || by0 >= h) return CVB_MODE_SKIP;  // box entirely off the image
    return CVB_MODE_BOX;
  };
  // One call = one iteration into the ring: all 16 planes of (super-group, view) if one box bounds them -- the far,
  // slowly moving planes: 4x the re-use of every staged texel -- else 4 planes at a time.
  int p_sg = 0, p_k = 0, p_m = -1, p_it = 0;
  auto produce_next = [&]() {
    if (p_sg >= n_sg) return;
    const int stage = p_it % CVB_STAGES, use = p_it / CVB_STAGES;
    if (use > 0) tc::mbar_wait(&empty[stage], (use - 1) & 1u);
    int d0 = p_sg * CVB_SG, n = CVB_SG, bx0 = 0, by0 = 0, mode = CVB_MODE_GLOBAL;
    const int k = p_k;
    if (p_m < 0) {
      mode = bound(k, d0, d0 + CVB_SG - 1, bx0, by0);
      if (mode == CVB_MODE_GLOBAL) p_m = 0;  // does not fit (or behind the camera): four pieces
    }
    if (p_m >= 0) {
      d0 += CVB_G * p_m;
      n = CVB_G;
      mode = bound(k, d0, d0 + CVB_G - 1, bx0, by0);
      if (++p_m == CVB_SG / CVB_G) p_m = -1;
    }
    if (p_m < 0 && ++p_k == K) {
      p_k = 0;
      ++p_sg;
    }
    if (lane == 0) {
      s_meta[stage].bx0 = bx0;
      s_meta[stage].by0 = by0;
      s_meta[stage].mode = mode;
      s_meta[stage].d0n = d0 * 32 + n;
      if (mode == CVB_MODE_BOX) {
        tc::mbar_expect_tx(&full[stage], CVB_BOX_BYTES);
        tc::tma_load_4d(box + (size_t)stage * CVB_BOX_FLOATS, &prm.src_map, 0, bx0, by0, b * K + k, &full[stage]);
      } else {
        tc::mbar_arrive(&full[stage]);
      }
    }
    ++p_it;
    __syncwarp();
  };

  if (warp == 0) {
    produce_next();
    produce_next();
  }

  float best = 0.f;
  int bi = 0;
  int it = 0;
  for (int sg = 0; sg < n_sg; ++sg) {
    float acc[CVB_SG];
#pragma unroll
    for (int m = 0; m < CVB_SG; ++m) acc[m] = 0.f;
    for (int k = 0; k < K; ++k) {
      // projection constants of (this pixel, view k): Mp = M @ (x+.5, y+.5, 1), t = P[:, 3]
      const float* cam = s_cam + k * B200_CAM_STRIDE;
      float Mp[3], tp[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        Mp[i] = fmaf(cam[CAM_M + 3 * i], pxc, fmaf(cam[CAM_M + 3 * i + 1], pyc, cam[CAM_M + 3 * i + 2]));
        tp[i] = cam[CAM_P + 4 * i + 3];
      }
      const float* sk = srcb + (size_t)k * N * B200_FEAT_C;
      for (int done = 0; done < CVB_SG;) {
        if (warp == 0) produce_next();  // keeps the ring two iterations ahead
        const int stage = it % CVB_STAGES;
        tc::mbar_wait(&full[stage], (it / CVB_STAGES) & 1u);
        const int bx0 = s_meta[stage].bx0, by0 = s_meta[stage].by0, mode = s_meta[stage].mode;
        const int d0 = s_meta[stage].d0n >> 5, n = s_meta[stage].d0n & 31;
        const float* bs = box + (size_t)stage * CVB_BOX_FLOATS + 4 * j;
        if (mode != CVB_MODE_SKIP) {
#pragma unroll
          for (int mm = 0; mm < CVB_SG / CVB_G; ++mm) {
            if (n == CVB_SG || d0 == sg * CVB_SG + mm * CVB_G) {  // warp-uniform
              const float zd = s_planes[sg * CVB_SG + mm * CVB_G + j];  // this lane projects plane 4 mm + j of the group
              float ax, ay, bx, by, cz;
              int x0, y0;
              cvb_project(Mp[0], Mp[1], Mp[2], tp[0], tp[1], tp[2], zd, h, w, ax, ay, bx, by, x0, y0, cz);
              // offset of the nw texel inside the box (floats).  The corner bound (+ one texel of margin) guarantees
              // that every sample of a BOX iteration lies inside; should rounding ever put one outside, the whole warp
              // takes the global path for these planes (warp-uniform: the shuffles below stay convergent).
              const int rx = x0 - bx0, ry = y0 - by0;
              const bool inbox = ((unsigned)rx <= (unsigned)(CVB_BW - 2)) & ((unsigned)ry <= (unsigned)(CVB_BH - 2));
              const bool use_box = (mode == CVB_MODE_BOX) && __all_sync(0xffffffffu, inbox);
              const int off = (ry * CVB_BW + rx) * B200_FEAT_C;
#pragma unroll
              for (int i = 0; i < CVB_G; ++i) {
                const int src_lane = (lane & ~3) | i;
                const float axi = __shfl_sync(0xffffffffu, ax, src_lane), ayi = __shfl_sync(0xffffffffu, ay, src_lane);
                const float bxi = __shfl_sync(0xffffffffu, bx, src_lane), byi = __shfl_sync(0xffffffffu, by, src_lane);
                float4 s00, s01, s10, s11;
                float wx0 = bxi, wx1 = axi, wy0 = byi, wy1 = ayi;
                if (use_box) {
                  // out-of-image texels of the box are zeros (TMA fill) = grid_sample's zeros padding
                  const int o = __shfl_sync(0xffffffffu, off, src_lane);
                  s00 = *reinterpret_cast<const float4*>(bs + o);
                  s01 = *reinterpret_cast<const float4*>(bs + o + B200_FEAT_C);
                  s10 = *reinterpret_cast<const float4*>(bs + o + CVB_BW * B200_FEAT_C);
                  s11 = *reinterpret_cast<const float4*>(bs + o + CVB_BW * B200_FEAT_C + B200_FEAT_C);
                } else {
                  // global path (cv_dot_kernel's): clamped texels, weights of out-of-image taps zeroed
                  const int xs = __shfl_sync(0xffffffffu, x0, src_lane), ys = __shfl_sync(0xffffffffu, y0, src_lane);
                  if (!(((unsigned)(xs + 1) <= (unsigned)w) & ((unsigned)(ys + 1) <= (unsigned)h))) continue;  // all outside
                  wx0 = ((unsigned)xs < (unsigned)w) ? bxi : 0.f;
                  wx1 = ((unsigned)(xs + 1) < (unsigned)w) ? axi : 0.f;
                  wy0 = ((unsigned)ys < (unsigned)h) ? byi : 0.f;
                  wy1 = ((unsigned)(ys + 1) < (unsigned)h) ? ayi : 0.f;
                  const int xa = min(max(xs, 0), w - 1), xb = min(max(xs + 1, 0), w - 1);
                  const int ya = min(max(ys, 0), h - 1), yb = min(max(ys + 1, 0), h - 1);
                  s00 = ldg4(sk + (size_t)(ya * w + xa) * B200_FEAT_C);
                  s01 = ldg4(sk + (size_t)(ya * w + xb) * B200_FEAT_C);
                  s10 = ldg4(sk + (size_t)(yb * w + xa) * B200_FEAT_C);
                  s11 = ldg4(sk + (size_t)(yb * w + xb) * B200_FEAT_C);
                }
                const float w00 = wx0 * wy0, w01 = wx1 * wy0, w10 = wx0 * wy1, w11 = wx1 * wy1;  // nw, ne, sw, se
                float d00 = s00.x * c4.x, d01 = s01.x * c4.x, d10 = s10.x * c4.x, d11 = s11.x * c4.x;
                d00 = fmaf(s00.y, c4.y, d00); d01 = fmaf(s01.y, c4.y, d01); d10 = fmaf(s10.y, c4.y, d10); d11 = fmaf(s11.y, c4.y, d11);
                d00 = fmaf(s00.z, c4.z, d00); d01 = fmaf(s01.z, c4.z, d01); d10 = fmaf(s10.z, c4.z, d10); d11 = fmaf(s11.z, c4.z, d11);
                d00 = fmaf(s00.w, c4.w, d00); d01 = fmaf(s01.w, c4.w, d01); d10 = fmaf(s10.w, c4.w, d10); d11 = fmaf(s11.w, c4.w, d11);
                float dotk = w00 * d00;
                dotk = fmaf(w01, d01, dotk);
                dotk = fmaf(w10, d10, dotk);
                dotk = fmaf(w11, d11, dotk);
                acc[mm * CVB_G + i] += dotk;  // mask = (z > 0) is identically 1: z is clamped to 1e-5 (cost_volume.py:216)
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty[stage]);
        ++it;
        done += n;
      }
    }
    // ---- super-group done: reduce the channel quarters, store, running arg-max in plane order ----
#pragma unroll
    for (int mm = 0; mm < CVB_SG / CVB_G; ++mm) {
      float mine = 0.f;
#pragma unroll
      for (int i = 0; i < CVB_G; ++i) {
        float a = acc[mm * CVB_G + i];
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        const int d = sg * CVB_SG + mm * CVB_G + i;
        if (i == j) mine = a;
        if (d < D && (d == 0 || a > best)) {  // strict '>' keeps the first maximum (torch.argmax, :354)
          best = a;
          bi = d;
        }
      }
      const int dd = sg * CVB_SG + mm * CVB_G + j;
      if (live && dd < D) prm.cost[((size_t)b * D + dd) * N + p] = mine;
    }
  }
  if (live && j == 0) {
    if (prm.lowest) prm.lowest[(size_t)b * N + p] = s_planes[bi];
    if (prm.best_idx) prm.best_idx[(size_t)b * N + p] = bi;
  }
}

extern "C" int b200_cv_dot_band(const float* cur, const float* src, const float* cams, const float* planes, float* cost,
                                float* lowest, int* best_idx, int B, int K, int C, int h, int w, int D, void* stream) {
  B200_CHECK_ARG(C == B200_FEAT_C, "cv_dot_band: only %d feature channels supported (got %d)", B200_FEAT_C, C);
  B200_CHECK_ARG(B > 0 && K > 0 && K <= B200_MAX_VIEWS && D > 0 && h > 0 && w > 0,
                 "cv_dot_band: bad sizes B=%d K=%d D=%d h=%d w=%d", B, K, D, h, w);
  B200_CHECK_ARG(D <= 4096, "cv_dot_band: at most 4096 depth planes (got %d)", D);
  B200_CHECK_ARG((long long)h * w * B200_FEAT_C < (1ll << 30), "cv_dot_band: feature map too large (%d x %d)", h, w);
  B200_CHECK_ARG(cur && src && cams && planes && cost, "cv_dot_band: null pointer");
  B200_CHECK_ARG((((uintptr_t)cur | (uintptr_t)src) & 15) == 0, "cv_dot_band: feature pointers must be 16-byte aligned");
  EncodeTiledFn enc = get_encode();
  B200_CHECK_ARG(enc != nullptr, "cv_dot_band: cuTensorMapEncodeTiled not available from the driver");
  CvbParams p;
  memset(&p, 0, sizeof(p));
  cuuint64_t gdim[4] = {B200_FEAT_C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B * K};
  cuuint64_t gstr[3] = {B200_FEAT_C * 4, (cuuint64_t)w * B200_FEAT_C * 4, (cuuint64_t)h * w * B200_FEAT_C * 4};
  cuuint32_t bdim[4] = {B200_FEAT_C, CVB_BW, CVB_BH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&p.src_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(src), gdim, gstr, bdim, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b200_set_error("cv_dot_band: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return -2;
  }
  p.cur = cur;
  p.src = src;
  p.cams = cams;
  p.planes = planes;
  p.cost = cost;
  p.lowest = lowest;
  p.best_idx = best_idx;
  p.K = K;
  p.D = D;
  p.h = h;
  p.w = w;
  const int Dp = (D + CVB_SG - 1) / CVB_SG * CVB_SG;
  const size_t smem = 128 + (size_t)CVB_STAGES * CVB_BOX_BYTES + sizeof(float) * (B200_MAX_VIEWS * B200_CAM_STRIDE + Dp) +
                      CVB_STAGES * (sizeof(CvbMeta) + 16) + 16;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(cv_dot_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  dim3 grid(((w + 7) / 8) * ((h + 7) / 8), B);
  cv_dot_band_kernel<<<grid, CVB_THREADS, smem, (cudaStream_t)stream>>>(p);
  B200_CHECK_LAUNCH("cv_dot_band");
  return 0;
}
