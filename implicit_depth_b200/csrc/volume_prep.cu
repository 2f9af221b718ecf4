// Small set-up kernels of the plane-sweep path: camera records, depth planes, hoisted
// MLP bias, layout change to the quarter-planar gather layout, arg-max over planes.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";
extern "C" const char* b200_last_error(void) { return g_err; }
void b200_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" int b200_abi_version(void) { return 2; }

// A dedicated non-blocking CUDA stream.  The host side keeps several streams that must be DISTINCT within one CUDA-graph
// capture (the plans' scheduler pools, the image-encoder stream, the pipeline's copy streams).  PyTorch hands its
// torch.cuda.Stream objects out of a fixed pool of 32 per device, round robin: a process that has created more than 32
// gets aliases, and two aliased branches of a captured step replay serially (measured: every 7th re-build of the
// plans in one process ran 8.3 instead of 6.8 ms).  Streams from here are wrapped as torch.cuda.ExternalStream.
extern "C" int b200_stream_create(int priority, void** stream_out) {
  B200_CHECK_ARG(stream_out != nullptr, "stream_create: null pointer");
  cudaStream_t s = nullptr;
  B200_CHECK_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, priority));
  *stream_out = (void*)s;
  return 0;
}
extern "C" int b200_stream_destroy(void* stream) {
  if (stream != nullptr) cudaStreamDestroy((cudaStream_t)stream);  // (work still queued on it completes first)
  return 0;
}
// sha256 of the sources this library was built from (passed by build.py as -DB200_SRC_DIGEST): `_abi.load()` compares
// it with the digest of the csrc/ it sits next to, so a stale binary is never loaded against newer ctypes signatures
#ifndef B200_SRC_DIGEST
#define B200_SRC_DIGEST "unknown"
#endif
extern "C" const char* b200_source_digest(void) { return B200_SRC_DIGEST; }

// ---------------------------------------------------------------------------------------
// One thread per (b, k): P = (K @ T)[:3], M = P3 @ invK3, t, pose distances.
// One thread per (b, d): z_d = exp(log zmin + log(zmax / zmin) * ramp_d)   (cost_volume.py:123-126)
// One thread per (b, n): hoisted first-layer bias of the feature-volume MLP: every input
// channel that is constant over (pixel, plane) -- the K mask channels (identically 1,
// SURVEY section 0.5) and the 3K pose-distance channels (cost_volume.py:690-692) -- is folded
// with b1 into bias_eff[b, n].
// ---------------------------------------------------------------------------------------
__global__ void volume_prepare_kernel(const float* __restrict__ src_Ks, const float* __restrict__ src_extr,
                                      const float* __restrict__ src_poses, const float* __restrict__ cur_invK,
                                      const float* __restrict__ min_depth, const float* __restrict__ max_depth,
                                      const float* __restrict__ planes_in, const float* __restrict__ W1,
                                      const float* __restrict__ b1, float* __restrict__ cams,
                                      float* __restrict__ planes, float* __restrict__ bias_eff, int B, int K, int D,
                                      int C, int range_stride) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  int nthreads = gridDim.x * blockDim.x;
  for (int i = tid; i < B * K; i += nthreads) {
    int b = i / K;
    const float* Km = src_Ks + (size_t)i * 16;
    const float* T = src_extr + (size_t)i * 16;
    const float* Ps = src_poses + (size_t)i * 16;
    const float* iK = cur_invK + (size_t)b * 16;
    float* cam = cams + (size_t)i * B200_CAM_STRIDE;
    float P[12];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) {
        float s = 0.f;
        for (int j = 0; j < 4; ++j) s = fmaf(Km[r * 4 + j], T[j * 4 + c], s);
        P[r * 4 + c] = s;
        cam[CAM_P + r * 4 + c] = s;
      }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        float s = 0.f;
        for (int j = 0; j < 3; ++j) s = fmaf(P[r * 4 + j], iK[j * 4 + c], s);
        cam[CAM_M + r * 3 + c] = s;
      }
    float tx = Ps[3], ty = Ps[7], tz = Ps[11];
    cam[CAM_T + 0] = tx;
    cam[CAM_T + 1] = ty;
    cam[CAM_T + 2] = tz;
    float tr = Ps[0] + Ps[5] + Ps[10];
    float r_meas = sqrtf(2.f * (1.f - fminf(3.f, tr) / 3.f));
    float t_meas = sqrtf(tx * tx + ty * ty + tz * tz);
    cam[CAM_POSE + 0] = sqrtf(t_meas * t_meas + r_meas * r_meas);
    cam[CAM_POSE + 1] = r_meas;
    cam[CAM_POSE + 2] = t_meas;
    for (int j = 27; j < B200_CAM_STRIDE; ++j) cam[j] = 0.f;
  }
  for (int i = tid; i < B * D; i += nthreads) {
    int d = i % D, b = i / D;
    if (planes_in) {
      planes[i] = planes_in[i];
    } else {
      // the reference broadcasts its min/max tensors over the batch (cost_volume.py:117-126): one range, or one per frame
      float zmin = min_depth[b * range_stride], zmax = max_depth[b * range_stride];
      float ramp = (D > 1) ? (float)d / (float)(D - 1) : 0.f;
      planes[i] = expf(logf(zmin) + logf(zmax / zmin) * ramp);
    }
  }
  if (W1 != nullptr) {
    const int cin = 26 * K + 4 + C;               // 16(K+1)+... with C=16 -> 26K+20
    const int off_mask = C * (K + 1);
    const int off_pose = C * (K + 1) + 4 * K + 1 + 3 * (K + 1);
    for (int i = tid; i < B * B200_MLP_HID; i += nthreads) {
      int b = i / B200_MLP_HID, n = i % B200_MLP_HID;
      const float* wrow = W1 + (size_t)n * cin;
      float s = b1[n];
      for (int k = 0; k < K; ++k) {
        const float* Ps = src_poses + (size_t)(b * K + k) * 16;
        float tx = Ps[3], ty = Ps[7], tz = Ps[11];
        float tr = Ps[0] + Ps[5] + Ps[10];
        float r_meas = sqrtf(2.f * (1.f - fminf(3.f, tr) / 3.f));
        float t_meas = sqrtf(tx * tx + ty * ty + tz * tz);
        float dist = sqrtf(t_meas * t_meas + r_meas * r_meas);
        s += wrow[off_mask + k];
        s = fmaf(wrow[off_pose + k], dist, s);
        s = fmaf(wrow[off_pose + K + k], r_meas, s);
        s = fmaf(wrow[off_pose + 2 * K + k], t_meas, s);
      }
      bias_eff[i] = s;
    }
  }
}

extern "C" int b200_volume_prepare(const float* src_Ks, const float* src_extrinsics, const float* src_poses,
                                   const float* cur_invK, const float* min_depth, const float* max_depth,
                                   const float* planes_in, const float* W1, const float* b1, float* cams,
                                   float* planes, float* bias_eff, int B, int K, int D, int C, int range_per_frame,
                                   void* stream) {
  B200_CHECK_ARG(B > 0 && K > 0 && K <= B200_MAX_VIEWS && D > 0, "volume_prepare: bad sizes B=%d K=%d D=%d", B, K, D);
  B200_CHECK_ARG(C == B200_FEAT_C, "volume_prepare: only %d feature channels supported (got %d)", B200_FEAT_C, C);
  B200_CHECK_ARG(src_Ks && src_extrinsics && src_poses && cur_invK && cams && planes, "volume_prepare: null pointer");
  B200_CHECK_ARG(planes_in || (min_depth && max_depth), "volume_prepare: need planes_in or min/max depth");
  B200_CHECK_ARG(!W1 || (b1 && bias_eff), "volume_prepare: W1 given without b1/bias_eff");
  volume_prepare_kernel<<<4, 256, 0, (cudaStream_t)stream>>>(src_Ks, src_extrinsics, src_poses, cur_invK, min_depth,
                                                            max_depth, planes_in, W1, b1, cams, planes, bias_eff, B,
                                                            K, D, C, range_per_frame ? 1 : 0);
  B200_CHECK_LAUNCH("volume_prepare");
  return 0;
}

// ---------------------------------------------------------------------------------------
// [n_img, C, HW] (strided planes) -> texel records [n_img, HW, C] (layout 0, cv_dot) or quarter-planar
// [n_img, C/4, HW, 4] (layout 1, feature-volume kernels; common.cuh), C = 16.  32x16 tile through shared memory so both sides are coalesced.
// ---------------------------------------------------------------------------------------
__global__ void nchw_to_pixel_major_kernel(const float* __restrict__ in, float* __restrict__ out, int HW,
                                           long long img_stride, long long ch_stride, int qplanar) {
  __shared__ float tile[B200_FEAT_C][33];
  int img = blockIdx.y;
  int p0 = blockIdx.x * 32;
  const float* src = in + (size_t)img * img_stride;
  for (int i = threadIdx.x; i < B200_FEAT_C * 32; i += blockDim.x) {
    int c = i / 32, p = i % 32;
    tile[c][p] = (p0 + p < HW) ? src[(size_t)c * ch_stride + p0 + p] : 0.f;
  }
  __syncthreads();
  float* dst = out + (size_t)img * HW * B200_FEAT_C;
  for (int i = threadIdx.x; i < B200_FEAT_C * 32; i += blockDim.x) {
    if (qplanar) {
      int q = i / (32 * FEAT_Q), p = (i / FEAT_Q) % 32, cc = i % FEAT_Q;
      if (p0 + p < HW) dst[((size_t)q * HW + p0 + p) * FEAT_Q + cc] = tile[q * FEAT_Q + cc][p];
    } else {
      int p = i / B200_FEAT_C, c = i % B200_FEAT_C;
      if (p0 + p < HW) dst[(size_t)(p0 + p) * B200_FEAT_C + c] = tile[c][p];
    }
  }
}

extern "C" int b200_feats_to_pixel_major(const float* in, float* out, int n_img, int C, int HW,
                                         long long img_stride, long long ch_stride, int layout, void* stream) {
  B200_CHECK_ARG(C == B200_FEAT_C, "feats_to_pixel_major: only %d channels supported (got %d)", B200_FEAT_C, C);
  B200_CHECK_ARG(in && out && n_img > 0 && HW > 0, "feats_to_pixel_major: bad arguments");
  dim3 grid((HW + 31) / 32, n_img);
  nchw_to_pixel_major_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, HW, img_stride, ch_stride, layout);
  B200_CHECK_LAUNCH("feats_to_pixel_major");
  return 0;
}

// ---------------------------------------------------------------------------------------
// argmax over planes with the first-maximum rule of torch.argmax (cost_volume.py:352-356)
// and the gather of the plane depth.
// ---------------------------------------------------------------------------------------
__global__ void volume_argmax_kernel(const float* __restrict__ vol, const float* __restrict__ planes,
                                     float* __restrict__ lowest, int* __restrict__ best_idx, int D, int N) {
  int b = blockIdx.y;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const float* v = vol + (size_t)b * D * N + p;
  float best = v[0];
  int bi = 0;
  for (int d = 1; d < D; ++d) {
    float x = v[(size_t)d * N];
    if (x > best || (x != x && best == best)) {  // NaN counts as maximal, like torch.argmax
      best = x;
      bi = d;
    }
  }
  lowest[(size_t)b * N + p] = planes[b * D + bi];
  if (best_idx) best_idx[(size_t)b * N + p] = bi;
}

extern "C" int b200_volume_argmax(const float* vol, const float* planes, float* lowest, int* best_idx, int B, int D,
                                  int N, void* stream) {
  B200_CHECK_ARG(vol && planes && lowest && B > 0 && D > 0 && N > 0, "volume_argmax: bad arguments");
  dim3 grid((N + 255) / 256, B);
  volume_argmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(vol, planes, lowest, best_idx, D, N);
  B200_CHECK_LAUNCH("volume_argmax");
  return 0;
}
