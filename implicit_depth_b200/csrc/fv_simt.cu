// Feature-volume MLP (mlp_feature_volume), exact-fp32 CUDA-core version.
// One CTA = 128 consecutive pixels of one frame at one depth plane:
//   phase 1  build the 128 x (22K+20) input rows in shared memory (warp, dot, rays, angle)
//   phase 2  H1 = lrelu(A @ W1p + bias_eff[b])           (register-tiled 8x8 FFMA GEMM)
//   phase 3  H2 = lrelu(H1 @ W2t + b2); out = H2 . w3 + b3
// Replaces FeatureVolumeManager.build_cost_volume (modules/cost_volume.py:437-706) and
// FastFeatureVolumeManager.build_cost_volume (:938-1146); MLP = modules/networks.py:218-233.
// This kernel is the strict-fp32 variant and the on-device cross-check for the tcgen05 one.
#include "fv_rows.cuh"

#define FVS_ROWS 128
#define FVS_THREADS 256
#define FVS_KC 16  // weight rows staged per step

// C[128x128] (8x8 per thread) += A_s[128 x kdim] * Wg[kdim x 128]; Wg streamed through w_s.
__device__ __forceinline__ void block_gemm_128(float acc[8][8], const float* __restrict__ a_s, int lda,
                                               const float* __restrict__ Wg, int kdim, float* __restrict__ w_s,
                                               int ty, int tx) {
  for (int k0 = 0; k0 < kdim; k0 += FVS_KC) {
    __syncthreads();  // previous chunk consumed (and A rows complete on first pass)
    for (int i = threadIdx.x; i < FVS_KC * 128 / 4; i += FVS_THREADS)
      reinterpret_cast<float4*>(w_s)[i] = __ldg(reinterpret_cast<const float4*>(Wg + (size_t)k0 * 128) + i);
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < FVS_KC; ++kk) {
      float a[8], wv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = a_s[(ty * 8 + i) * lda + k0 + kk];
      const float4 w0 = *reinterpret_cast<const float4*>(w_s + kk * 128 + tx * 4);
      const float4 w1 = *reinterpret_cast<const float4*>(w_s + kk * 128 + 64 + tx * 4);
      wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
      wv[4] = w1.x; wv[5] = w1.y; wv[6] = w1.z; wv[7] = w1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], wv[j], acc[i][j]);
    }
  }
}

__device__ __forceinline__ int col_of(int tx, int j) { return (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4); }

__global__ void __launch_bounds__(FVS_THREADS)
fv_simt_kernel(const float* __restrict__ cur, const float* __restrict__ src, const float* __restrict__ cams,
               const float* __restrict__ invK, const float* __restrict__ planes, const float* __restrict__ bias_eff,
               const float* __restrict__ W1p, const float* __restrict__ W2t, const float* __restrict__ b2,
               const float* __restrict__ w3, const float* __restrict__ b3, float* __restrict__ vol,
               unsigned char* __restrict__ mask_out, int K, int D, int h, int w, int KP) {
  extern __shared__ __align__(16) float smem[];
  const int lda = KP + 1;
  const int ldmax = (KP > B200_MLP_HID ? KP : B200_MLP_HID) + 1;
  float* a_s = smem;                          // [128][lda]  (later H1 as [128][129])
  float* w_s = a_s + FVS_ROWS * ldmax;        // [FVS_KC][128]
  float* cam_s = w_s + FVS_KC * 128;          // [K][32]
  float* invk_s = cam_s + B200_MAX_VIEWS * B200_CAM_STRIDE;  // 9 (+pad)
  int* inb_s = reinterpret_cast<int*>(invk_s + 12);          // [128]

  const int N = h * w;
  const int b = blockIdx.z, d = blockIdx.y;
  const int p0 = blockIdx.x * FVS_ROWS;
  const float zd = planes[b * D + d];
  const bool want_mask = (mask_out != nullptr) && (d == D - 1);

  for (int i = threadIdx.x; i < K * B200_CAM_STRIDE; i += FVS_THREADS)
    cam_s[i] = cams[(size_t)b * K * B200_CAM_STRIDE + i];
  if (threadIdx.x < 9) invk_s[threadIdx.x] = invK[b * 16 + (threadIdx.x / 3) * 4 + threadIdx.x % 3];
  if (threadIdx.x < FVS_ROWS) inb_s[threadIdx.x] = 0;
  __syncthreads();

  // ---- phase 1: input rows ----
  const int kin = FV_VIEW_CH * K + FV_TAIL_CH;
  for (int idx = threadIdx.x; idx < FVS_ROWS * (K + 1); idx += FVS_THREADS) {
    const int row = idx & (FVS_ROWS - 1);
    const int k = idx >> 7;
    int p = min(p0 + row, N - 1);
    const int y = p / w, x = p - y * w;
    const PixelCtx pc = make_pixel_ctx(x, y, invk_s);
    float c16[16];
    const float* cp = cur + (size_t)b * N * B200_FEAT_C + (size_t)p * FEAT_Q;  // quarter-planar (common.cuh)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const float4 t4 = ldg4(cp + (size_t)v * N * FEAT_Q);
      c16[4 * v] = t4.x; c16[4 * v + 1] = t4.y; c16[4 * v + 2] = t4.z; c16[4 * v + 3] = t4.w;
    }
    float* arow = a_s + row * lda;
    if (k < K) {
      float out[FV_VIEW_CH];
      const bool inb = fv_view_block(pc, cam_s + k * B200_CAM_STRIDE,
                                     src + ((size_t)b * K + k) * N * B200_FEAT_C, c16, zd, h, w, out);
#pragma unroll
      for (int c = 0; c < FV_VIEW_CH; ++c) arow[k * FV_VIEW_CH + c] = out[c];
      if (want_mask && inb) atomicOr(&inb_s[row], 1);
    } else {
      float* t = arow + K * FV_VIEW_CH;
#pragma unroll
      for (int c = 0; c < 16; ++c) t[c] = c16[c];
      t[16] = pc.curray[0];
      t[17] = pc.curray[1];
      t[18] = pc.curray[2];
      t[19] = zd;
      for (int c = kin; c < KP; ++c) arow[c] = 0.f;  // K padding (weights are zero there too)
    }
  }
  // (block_gemm_128 starts with a __syncthreads)

  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // ---- phase 2: layer 1 ----
  block_gemm_128(acc, a_s, lda, W1p, KP, w_s, ty, tx);
  __syncthreads();  // everyone done reading A before it is overwritten by H1
  if (want_mask && threadIdx.x < FVS_ROWS && p0 + threadIdx.x < N)
    mask_out[(size_t)b * N + p0 + threadIdx.x] = (unsigned char)(inb_s[threadIdx.x] != 0);
  const int ldh = B200_MLP_HID + 1;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = col_of(tx, j);
    const float bn = bias_eff[b * B200_MLP_HID + n];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a_s[(ty * 8 + i) * ldh + n] = leaky(acc[i][j] + bn, 0.01f);  // nn.LeakyReLU default slope, networks.py:225
      acc[i][j] = 0.f;
    }
  }

  // ---- phase 3: layer 2 + layer 3 ----
  block_gemm_128(acc, a_s, ldh, W2t, B200_MLP_HID, w_s, ty, tx);
  float part[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) part[i] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = col_of(tx, j);
    const float bn = b2[n], wn = w3[n];
#pragma unroll
    for (int i = 0; i < 8; ++i) part[i] = fmaf(leaky(acc[i][j] + bn, 0.01f), wn, part[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v = part[i];
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    const int p = p0 + ty * 8 + i;
    if (tx == 0 && p < N) vol[((size_t)b * D + d) * N + p] = v + b3[0];
  }
}

extern "C" int b200_fv_mlp_simt(const float* cur, const float* src, const float* cams, const float* cur_invK,
                                const float* planes, const float* bias_eff, const float* W1p, const float* W2t,
                                const float* b2, const float* w3, const float* b3, float* vol,
                                unsigned char* mask_out, int B, int K, int C, int h, int w, int D, int KP,
                                void* stream) {
  B200_CHECK_ARG(C == B200_FEAT_C, "fv_mlp_simt: only %d feature channels supported (got %d)", B200_FEAT_C, C);
  B200_CHECK_ARG(B > 0 && K > 0 && K <= B200_MAX_VIEWS && D > 0 && h > 0 && w > 0,
                 "fv_mlp_simt: bad sizes B=%d K=%d D=%d h=%d w=%d", B, K, D, h, w);
  B200_CHECK_ARG(KP % FVS_KC == 0 && KP >= FV_VIEW_CH * K + FV_TAIL_CH && KP <= 256,
                 "fv_mlp_simt: KP=%d must be a multiple of %d covering %d channels", KP, FVS_KC,
                 FV_VIEW_CH * K + FV_TAIL_CH);
  B200_CHECK_ARG(cur && src && cams && cur_invK && planes && bias_eff && W1p && W2t && b2 && w3 && b3 && vol,
                 "fv_mlp_simt: null pointer");
  const int N = h * w;
  const int lda = (KP > B200_MLP_HID ? KP : B200_MLP_HID) + 1;
  size_t smem = sizeof(float) * ((size_t)FVS_ROWS * lda + FVS_KC * 128 + B200_MAX_VIEWS * B200_CAM_STRIDE + 12) +
                sizeof(int) * FVS_ROWS;
  static bool attr_done = false;
  if (!attr_done) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(fv_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  B200_CHECK_ARG(smem <= 200 * 1024, "fv_mlp_simt: shared memory %zu too large", smem);
  dim3 grid((N + FVS_ROWS - 1) / FVS_ROWS, D, B);
  fv_simt_kernel<<<grid, FVS_THREADS, smem, (cudaStream_t)stream>>>(cur, src, cams, cur_invK, planes, bias_eff, W1p,
                                                                   W2t, b2, w3, b3, vol, mask_out, K, D, h, w, KP);
  B200_CHECK_LAUNCH("fv_mlp_simt");
  return 0;
}
