// Input side of the step: the small matrix algebra the reference does with per-tensor torch ops
// between the data loader and the plane sweep, as two tiny kernels over a staged batch.
//   * relative poses: BDModel.forward, experiment_modules/bd_model.py:196-204 (two batched 4x4 matmuls);
//   * intrinsics pyramid: datasets/scannet_dataset.py:479-484 (K[:2] / 2^i and its inverse for i = 0..4).
#include "common.cuh"

__device__ __forceinline__ void mat4_mul(const float* __restrict__ A, const float* __restrict__ Bm,
                                         float* __restrict__ out) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) s = fmaf(A[r * 4 + j], Bm[j * 4 + c], s);
      out[r * 4 + c] = s;
    }
}

// One thread per (frame, source view):
//   src_cam_T_cur_cam[b,k] = src_cam_T_world[b,k] @ cur_world_T_cam[b]      (bd_model.py:200)
//   cur_cam_T_src_cam[b,k] = cur_cam_T_world[b]   @ src_world_T_cam[b,k]    (bd_model.py:204)
__global__ void relative_poses_kernel(const float* __restrict__ src_cam_T_world,
                                      const float* __restrict__ src_world_T_cam,
                                      const float* __restrict__ cur_cam_T_world,
                                      const float* __restrict__ cur_world_T_cam,
                                      float* __restrict__ src_cam_T_cur_cam, float* __restrict__ cur_cam_T_src_cam,
                                      int B, int K) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K) return;
  int b = i / K;
  float A[16], Bm[16], C[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    A[j] = src_cam_T_world[(size_t)i * 16 + j];
    Bm[j] = cur_world_T_cam[(size_t)b * 16 + j];
  }
  mat4_mul(A, Bm, C);
#pragma unroll
  for (int j = 0; j < 16; ++j) src_cam_T_cur_cam[(size_t)i * 16 + j] = C[j];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    A[j] = cur_cam_T_world[(size_t)b * 16 + j];
    Bm[j] = src_world_T_cam[(size_t)i * 16 + j];
  }
  mat4_mul(A, Bm, C);
#pragma unroll
  for (int j = 0; j < 16; ++j) cur_cam_T_src_cam[(size_t)i * 16 + j] = C[j];
}

extern "C" int b200_relative_poses(const float* src_cam_T_world, const float* src_world_T_cam,
                                   const float* cur_cam_T_world, const float* cur_world_T_cam,
                                   float* src_cam_T_cur_cam, float* cur_cam_T_src_cam, int B, int K, void* stream) {
  B200_CHECK_ARG(B > 0 && K > 0, "relative_poses: bad sizes B=%d K=%d", B, K);
  B200_CHECK_ARG(src_cam_T_world && src_world_T_cam && cur_cam_T_world && cur_world_T_cam && src_cam_T_cur_cam &&
                     cur_cam_T_src_cam,
                 "relative_poses: null pointer");
  int n = B * K;
  relative_poses_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(src_cam_T_world, src_world_T_cam,
                                                                       cur_cam_T_world, cur_world_T_cam,
                                                                       src_cam_T_cur_cam, cur_cam_T_src_cam, B, K);
  B200_CHECK_LAUNCH("relative_poses");
  return 0;
}

// General 4x4 inverse by the adjugate in fp64, rounded once to fp32 (the reference inverts in fp32 LAPACK; the two
// agree to a few ulps).  A singular matrix yields inf/NaN -- nothing can be raised from a kernel.
__device__ void inv4_f64(const double* m, float* __restrict__ out) {
  double s0 = m[0] * m[5] - m[4] * m[1], s1 = m[0] * m[6] - m[4] * m[2], s2 = m[0] * m[7] - m[4] * m[3];
  double s3 = m[1] * m[6] - m[5] * m[2], s4 = m[1] * m[7] - m[5] * m[3], s5 = m[2] * m[7] - m[6] * m[3];
  double c5 = m[10] * m[15] - m[14] * m[11], c4 = m[9] * m[15] - m[13] * m[11], c3 = m[9] * m[14] - m[13] * m[10];
  double c2 = m[8] * m[15] - m[12] * m[11], c1 = m[8] * m[14] - m[12] * m[10], c0 = m[8] * m[13] - m[12] * m[9];
  double inv = 1.0 / (s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0);
  out[0] = (float)((m[5] * c5 - m[6] * c4 + m[7] * c3) * inv);
  out[1] = (float)((-m[1] * c5 + m[2] * c4 - m[3] * c3) * inv);
  out[2] = (float)((m[13] * s5 - m[14] * s4 + m[15] * s3) * inv);
  out[3] = (float)((-m[9] * s5 + m[10] * s4 - m[11] * s3) * inv);
  out[4] = (float)((-m[4] * c5 + m[6] * c2 - m[7] * c1) * inv);
  out[5] = (float)((m[0] * c5 - m[2] * c2 + m[3] * c1) * inv);
  out[6] = (float)((-m[12] * s5 + m[14] * s2 - m[15] * s1) * inv);
  out[7] = (float)((m[8] * s5 - m[10] * s2 + m[11] * s1) * inv);
  out[8] = (float)((m[4] * c4 - m[5] * c2 + m[7] * c0) * inv);
  out[9] = (float)((-m[0] * c4 + m[1] * c2 - m[3] * c0) * inv);
  out[10] = (float)((m[12] * s4 - m[13] * s2 + m[15] * s0) * inv);
  out[11] = (float)((-m[8] * s4 + m[9] * s2 - m[11] * s0) * inv);
  out[12] = (float)((-m[4] * c3 + m[5] * c1 - m[6] * c0) * inv);
  out[13] = (float)((m[0] * c3 - m[1] * c1 + m[2] * c0) * inv);
  out[14] = (float)((-m[12] * s3 + m[13] * s1 - m[14] * s0) * inv);
  out[15] = (float)((m[8] * s3 - m[9] * s1 + m[10] * s0) * inv);
}

// One thread per (level, matrix): K_s[i] = K with rows 0 and 1 divided by 2^i; invK_s[i] = inverse(K_s[i]).
__global__ void intrinsics_pyramid_kernel(const float* __restrict__ K0, float* __restrict__ Ks,
                                          float* __restrict__ invKs, int n, int levels) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * levels) return;
  int lvl = t / n, i = t % n;
  float scale = 1.f / (float)(1 << lvl);  // exact power of two: same bits as the reference's division
  float Kf[16];
  double Kd[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float v = K0[(size_t)i * 16 + j];
    if (j < 8) v *= scale;
    Kf[j] = v;
    Kd[j] = (double)v;
  }
  float* ko = Ks + ((size_t)lvl * n + i) * 16;
#pragma unroll
  for (int j = 0; j < 16; ++j) ko[j] = Kf[j];
  float inv[16];
  inv4_f64(Kd, inv);
  float* io = invKs + ((size_t)lvl * n + i) * 16;
#pragma unroll
  for (int j = 0; j < 16; ++j) io[j] = inv[j];
}

extern "C" int b200_intrinsics_pyramid(const float* K_s0, float* K_s, float* invK_s, int n, int levels,
                                       void* stream) {
  B200_CHECK_ARG(n > 0 && levels > 0 && levels <= 16, "intrinsics_pyramid: bad sizes n=%d levels=%d", n, levels);
  B200_CHECK_ARG(K_s0 && K_s && invK_s, "intrinsics_pyramid: null pointer");
  int total = n * levels;
  intrinsics_pyramid_kernel<<<(total + 63) / 64, 64, 0, (cudaStream_t)stream>>>(K_s0, K_s, invK_s, n, levels);
  B200_CHECK_LAUNCH("intrinsics_pyramid");
  return 0;
}
