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
// The per-view partial dots are accumulated per lane and reduced over the quad with two
// xor-shuffles once per (pixel, plane).
#include "common.cuh"

#define CVD_THREADS 128
#define CVD_PIX_PER_BLOCK (CVD_THREADS / 4)

__global__ void __launch_bounds__(CVD_THREADS)
cv_dot_kernel(const float* __restrict__ cur,     // [B, N, 16] pixel-major
              const float* __restrict__ src,     // [B, K, N, 16] pixel-major
              const float* __restrict__ cams,    // [B, K, 32]
              const float* __restrict__ planes,  // [B, D]
              float* __restrict__ cost,          // [B, D, N]
              float* __restrict__ lowest,        // [B, N] or null
              int* __restrict__ best_idx,        // [B, N] or null
              int K, int D, int h, int w) {
  __shared__ float s_cam[B200_MAX_VIEWS * B200_CAM_STRIDE];
  extern __shared__ float s_planes[];
  const int N = h * w;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < K * B200_CAM_STRIDE; i += CVD_THREADS)
    s_cam[i] = cams[(size_t)b * K * B200_CAM_STRIDE + i];
  for (int i = threadIdx.x; i < D; i += CVD_THREADS) s_planes[i] = planes[b * D + i];
  __syncthreads();

  const int q = threadIdx.x >> 2;  // pixel slot in block
  const int j = threadIdx.x & 3;   // channel quarter
  int p = blockIdx.x * CVD_PIX_PER_BLOCK + q;
  const bool live = p < N;
  if (!live) p = N - 1;  // keep the quad convergent for the shuffles; result discarded
  const int y = p / w, x = p - y * w;
  const float pxc = x + 0.5f, pyc = y + 0.5f;  // pixel centres, geometry_utils.py:39

  const float4 c4 = ldg4(cur + ((size_t)b * N + p) * B200_FEAT_C + 4 * j);
  const float* srcb = src + (size_t)b * K * N * B200_FEAT_C + 4 * j;

  float best = 0.f;
  int bi = 0;
  for (int d = 0; d < D; ++d) {
    const float zd = s_planes[d];
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      const float* cam = s_cam + k * B200_CAM_STRIDE;
      float Mp[3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
        Mp[r] = fmaf(cam[CAM_M + 3 * r], pxc, fmaf(cam[CAM_M + 3 * r + 1], pyc, cam[CAM_M + 3 * r + 2]));
      float px, py, z;
      project_plane(Mp, cam, zd, px, py, z);
      const Taps t = make_taps(px, py, h, w);
      const float* sk = srcb + (size_t)k * N * B200_FEAT_C;
      float dotk = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (t.idx[i] >= 0) {
          const float4 s4 = ldg4(sk + (size_t)t.idx[i] * B200_FEAT_C);
          float dt = s4.x * c4.x;
          dt = fmaf(s4.y, c4.y, dt);
          dt = fmaf(s4.z, c4.z, dt);
          dt = fmaf(s4.w, c4.w, dt);
          dotk = fmaf(t.wgt[i], dt, dotk);
        }
      }
      // mask = (z > 0) is identically 1 because z is clamped to 1e-5 (cost_volume.py:216)
      acc += dotk;
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (live && j == 0) cost[((size_t)b * D + d) * N + p] = acc;
    if (d == 0 || acc > best) {  // strict '>' keeps the first maximum (torch.argmax, :354)
      best = acc;
      bi = d;
    }
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
  B200_CHECK_ARG(D <= 4096, "cv_dot: at most 4096 depth planes (got %d)", D);
  B200_CHECK_ARG(cur && src && cams && planes && cost, "cv_dot: null pointer");
  B200_CHECK_ARG((((uintptr_t)cur | (uintptr_t)src) & 15) == 0, "cv_dot: feature pointers must be 16-byte aligned");
  const int N = h * w;
  dim3 grid((N + CVD_PIX_PER_BLOCK - 1) / CVD_PIX_PER_BLOCK, B);
  cv_dot_kernel<<<grid, CVD_THREADS, D * sizeof(float), (cudaStream_t)stream>>>(cur, src, cams, planes, cost, lowest,
                                                                                best_idx, K, D, h, w);
  B200_CHECK_LAUNCH("cv_dot");
  return 0;
}
