// Memory-bound helper kernels of the conv path.  Activations are NHWC split-bf16 (hi, lo planes).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc.cuh"

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// load 8 consecutive channels of a split activation as fp32
__device__ __forceinline__ void load8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t off, float v[8]) {
  const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi + off));
  const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo + off));
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] = bf_lo(hw[e]) + bf_lo(lw[e]);
    v[2 * e + 1] = bf_hi(hw[e]) + bf_hi(lw[e]);
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t off, const float v[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) tc::split2(v[2 * e], v[2 * e + 1], h[e], l[e]);
  *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---------------------------------------------------------------------------------------
// fp32 [B,C,H,W] with arbitrary strides (NCHW or channels_last) -> NHWC split-bf16.
// 32 pixels x 32 channels through shared memory so both sides are coalesced.
// ---------------------------------------------------------------------------------------
__global__ void f32_to_split_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, int C, int HW, int W, long long sB, long long sC,
                                    long long sH, long long sW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 256 threads: 8 rows of 32
  const bool chan_fast = (sC == 1);
  for (int r = ty; r < 32; r += 8) {
    // read: make the fastest-varying input index follow tx
    const int c = chan_fast ? c0 + tx : c0 + r;
    const int p = chan_fast ? p0 + r : p0 + tx;
    float v = 0.f;
    if (c < C && p < HW) {
      const int y = p / W, x = p - y * W;
      v = in[(size_t)b * sB + (size_t)c * sC + (size_t)y * sH + (size_t)x * sW];
    }
    if (chan_fast) tile[r][tx] = v; else tile[tx][r] = v;  // tile[pixel][channel]
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    if (p < HW && c < C) {
      const float v = tile[r][tx];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const size_t o = ((size_t)b * HW + p) * C + c;
      hi[o] = h;
      lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

extern "C" int b200_f32_to_split(const float* in, void* hi, void* lo, int B, int C, int H, int W, long long sB,
                                 long long sC, long long sH, long long sW, void* stream) {
  B200_CHECK_ARG(in && hi && lo && B > 0 && C > 0 && H > 0 && W > 0, "f32_to_split: bad arguments");
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, B);
  f32_to_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, C, H * W, W,
                                                             sB, sC, sH, sW);
  B200_CHECK_LAUNCH("f32_to_split");
  return 0;
}

// NHWC split-bf16 -> fp32 NCHW (dense)
__global__ void split_to_nchw_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                     float* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    float v = 0.f;
    if (p < HW && c < C) {
      const size_t o = ((size_t)b * HW + p) * C + c;
      v = __bfloat162float(hi[o]) + __bfloat162float(lo[o]);
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    if (c < C && p < HW) out[((size_t)b * C + c) * HW + p] = tile[tx][r];
  }
}

extern "C" int b200_split_to_nchw(const void* hi, const void* lo, float* out, int B, int C, int H, int W,
                                  void* stream) {
  B200_CHECK_ARG(hi && lo && out && B > 0 && C > 0 && H > 0 && W > 0, "split_to_nchw: bad arguments");
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, B);
  split_to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)hi, (const __nv_bfloat16*)lo, out,
                                                              C, H * W);
  B200_CHECK_LAUNCH("split_to_nchw");
  return 0;
}

// ---------------------------------------------------------------------------------------
// x2 upsampling: mode 0 = bilinear, align_corners=False (utils/generic_utils.py:94-103, used by
// BDDecoderPP networks.py:72,75); mode 1 = nearest (networks_fast.py:42).  One thread = one INPUT pixel x 8
// channels = a 2x2 block of outputs: the four outputs blend the 3x3 input neighbourhood, 9 pixel loads instead of
// 16 (the kernel was issue-bound on unpacking its inputs).  Every output is formed by the same expression and
// weights as ATen's upsample_bilinear2d (src = max(0, (dst + 0.5) / 2 - 0.5)): the same values as the
// one-output-per-thread form it replaces.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
upsample2x_kernel(const __nv_bfloat16* __restrict__ ih, const __nv_bfloat16* __restrict__ il,
                  __nv_bfloat16* __restrict__ oh, __nv_bfloat16* __restrict__ ol, int B, int H, int W, int C, int mode) {
  const int cg = C >> 3;
  const size_t total = (size_t)B * H * W * cg;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cg);
    size_t r = i / cg;
    const int X = (int)(r % W);
    r /= W;
    const int Y = (int)(r % H);
    const int b = (int)(r / H);
    const size_t img = (size_t)b * H * W;
    const size_t obase = (((size_t)b * 2 * H + 2 * Y) * 2 * W + 2 * X) * C + c8 * 8;
    const size_t orow = (size_t)2 * W * C;
    if (mode == 1) {
      float v[8];
      load8(ih, il, (img + (size_t)Y * W + X) * C + c8 * 8, v);
      store8(oh, ol, obase, v);
      store8(oh, ol, obase + C, v);
      store8(oh, ol, obase + orow, v);
      store8(oh, ol, obase + orow + C, v);
      continue;
    }
    // rows / columns of the neighbourhood: index 0 = previous (clamped), 1 = centre, 2 = next (clamped)
    const int ys[3] = {max(Y - 1, 0), Y, min(Y + 1, H - 1)};
    const int xs[3] = {max(X - 1, 0), X, min(X + 1, W - 1)};
    float v[3][3][8];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int q = 0; q < 3; ++q) load8(ih, il, (img + (size_t)ys[a] * W + xs[q]) * C + c8 * 8, v[a][q]);
#pragma unroll
    for (int py = 0; py < 2; ++py) {
      const int oy = 2 * Y + py;
      const float sy = fmaxf(0.f, (oy + 0.5f) * 0.5f - 0.5f);
      const int y0 = (int)sy;
      const float ly = sy - y0;
      // taps: rows (Y-1, Y) for the even output row, (Y, Y+1) for the odd one, clamped.  At the top border ATen's
      // taps are rows (0, 1) with ly = 0; rows (0, 0) with ly = 0 give the same value.  Columns alike.
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        const int ox = 2 * X + px;
        const float sx = fmaxf(0.f, (ox + 0.5f) * 0.5f - 0.5f);
        const int x0 = (int)sx;
        const float lx = sx - x0;
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
          o[e] = (1.f - ly) * ((1.f - lx) * v[py][px][e] + lx * v[py][px + 1][e]) +
                 ly * ((1.f - lx) * v[py + 1][px][e] + lx * v[py + 1][px + 1][e]);
        store8(oh, ol, obase + (size_t)py * orow + (size_t)px * C, o);
      }
    }
  }
}

extern "C" int b200_upsample2x(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int B, int H, int W,
                               int C, int mode, void* stream) {
  B200_CHECK_ARG(in_hi && in_lo && out_hi && out_lo && B > 0 && H > 0 && W > 0, "upsample2x: bad arguments");
  B200_CHECK_ARG(C % 8 == 0, "upsample2x: C must be a multiple of 8 (got %d)", C);
  B200_CHECK_ARG(mode == 0 || mode == 1, "upsample2x: mode 0 (bilinear) or 1 (nearest)");
  const size_t total = (size_t)B * H * W * (C / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  upsample2x_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in_hi, (const __nv_bfloat16*)in_lo,
                                                             (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, B, H, W,
                                                             C, mode);
  B200_CHECK_LAUNCH("upsample2x");
  return 0;
}

// ---------------------------------------------------------------------------------------
// InstanceNorm2d (affine=False, eps=1e-5, biased variance; networks.py:277,283).
// Deterministic two-stage reduction (no atomics): results of one image never depend on what else
// is in the batch -- the property the reference protects by running its encoder unbatched
// (bd_model.py:149-160, depth_model.py:235-241).
// ---------------------------------------------------------------------------------------
#define IN_SLICES 32
__global__ void instnorm_partial_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                        double* __restrict__ partial, int HW, int C) {
  const int b = blockIdx.y, sl = blockIdx.x;
  const int c = threadIdx.x;  // blockDim.x == C
  const int p_begin = (int)((long long)HW * sl / IN_SLICES), p_end = (int)((long long)HW * (sl + 1) / IN_SLICES);
  double s = 0.0, ss = 0.0;
  for (int p = p_begin; p < p_end; ++p) {
    const size_t o = ((size_t)b * HW + p) * C + c;
    const float v = __bfloat162float(hi[o]) + __bfloat162float(lo[o]);
    s += v;
    ss += (double)v * v;
  }
  double* dst = partial + (((size_t)b * IN_SLICES + sl) * C + c) * 2;
  dst[0] = s;
  dst[1] = ss;
}
__global__ void instnorm_final_kernel(const double* __restrict__ partial, float* __restrict__ stats, int HW, int C,
                                      float eps) {
  const int b = blockIdx.x, c = threadIdx.x;
  double s = 0.0, ss = 0.0;
  for (int sl = 0; sl < IN_SLICES; ++sl) {
    const double* src = partial + (((size_t)b * IN_SLICES + sl) * C + c) * 2;
    s += src[0];
    ss += src[1];
  }
  const double mean = s / HW;
  const double var = fmax(ss / HW - mean * mean, 0.0);
  stats[((size_t)b * C + c) * 2] = (float)mean;
  stats[((size_t)b * C + c) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
}
// apply: y = (x - mean) * rstd, optional LeakyReLU; output either split NHWC with a replicated border of
// `pad` pixels (feeds the padding_mode="replicate" conv of networks.py:279-282 as a plain valid conv) or
// fp32 matching features in one of the two gather layouts of the volume kernels (common.cuh).
#define IN_ROWS 14  // output rows per thread of the apply kernel
__global__ void instnorm_apply_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                      const float* __restrict__ stats, __nv_bfloat16* __restrict__ oh,
                                      __nv_bfloat16* __restrict__ ol, float* __restrict__ of32, int B, int H, int W,
                                      int C, int pad, int act, float slope, int qplanar) {
  // one thread = 8 channels of one output column over a strip of IN_ROWS rows: the 16 statistics of its channels are
  // loaded once per strip (per element they were a third of the kernel's L1 traffic, which ran at 90 % of the LSU peak)
  const int cg = C >> 3;
  const int OH = H + 2 * pad, OW = W + 2 * pad;
  const int nstrip = (OH + IN_ROWS - 1) / IN_ROWS;
  const size_t total = (size_t)B * nstrip * OW * cg;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cg);
    size_t r = i / cg;
    const int ox = (int)(r % OW);
    r /= OW;
    const int strip = (int)(r % nstrip);
    const int b = (int)(r / nstrip);
    const int x = min(max(ox - pad, 0), W - 1);
    float mean[8], rstd[8];
    const float4* st4 = reinterpret_cast<const float4*>(stats + ((size_t)b * C + c8 * 8) * 2);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 t = __ldg(st4 + q);
      mean[2 * q] = t.x; rstd[2 * q] = t.y; mean[2 * q + 1] = t.z; rstd[2 * q + 1] = t.w;
    }
    const int oy_end = min((strip + 1) * IN_ROWS, OH);
    for (int oy = strip * IN_ROWS; oy < oy_end; ++oy) {
      const int y = min(max(oy - pad, 0), H - 1);
      float v[8];
      load8(hi, lo, (((size_t)b * H + y) * W + x) * C + c8 * 8, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float t = (v[e] - mean[e]) * rstd[e];
        if (act == 1) t = t >= 0.f ? t : t * slope;
        v[e] = t;
      }
      const size_t o = (((size_t)b * OH + oy) * OW + ox) * C + c8 * 8;
      if (oh) store8(oh, ol, o, v);
      if (of32 && !qplanar) {  // texel records [B, H*W, C]
        *reinterpret_cast<float4*>(of32 + o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(of32 + o + 4) = make_float4(v[4], v[5], v[6], v[7]);
      } else if (of32) {  // quarter-planar fp32 [B, C/4, H*W, 4] (common.cuh); pad == 0 here
        const size_t npix = (size_t)OH * OW, pix = (size_t)oy * OW + ox;
        float* q0 = of32 + (((size_t)b * (C >> 2) + 2 * c8) * npix + pix) * 4;
        *reinterpret_cast<float4*>(q0) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(q0 + npix * 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
}

extern "C" int b200_instance_norm(const void* in_hi, const void* in_lo, double* partial_ws, float* stats_ws,
                                  void* out_hi, void* out_lo, float* out_f32, int B, int H, int W, int C, int pad,
                                  int act, float slope, float eps, int f32_layout, void* stream) {
  B200_CHECK_ARG(in_hi && in_lo && partial_ws && stats_ws && (out_hi || out_f32), "instance_norm: null pointer");
  B200_CHECK_ARG(C % 8 == 0 && C <= 1024, "instance_norm: C must be a multiple of 8, at most 1024 (got %d)", C);
  B200_CHECK_ARG(!(out_f32 && pad != 0), "instance_norm: fp32 output has no border");
  cudaStream_t st = (cudaStream_t)stream;
  instnorm_partial_kernel<<<dim3(IN_SLICES, B), C, 0, st>>>((const __nv_bfloat16*)in_hi, (const __nv_bfloat16*)in_lo,
                                                            partial_ws, H * W, C);
  instnorm_final_kernel<<<B, C, 0, st>>>(partial_ws, stats_ws, H * W, C, eps);
  const size_t total = (size_t)B * ((H + 2 * pad + IN_ROWS - 1) / IN_ROWS) * (W + 2 * pad) * (C / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  instnorm_apply_kernel<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)in_hi, (const __nv_bfloat16*)in_lo, stats_ws,
                                               (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, out_f32, B, H, W, C,
                                               pad, act, slope, f32_layout);
  B200_CHECK_LAUNCH("instance_norm");
  return 0;
}
extern "C" int b200_instance_norm_ws_bytes(int B, int C, long long* partial_bytes, long long* stats_bytes) {
  *partial_bytes = (long long)B * IN_SLICES * C * 2 * sizeof(double);
  *stats_bytes = (long long)B * C * 2 * sizeof(float);
  return 0;
}

// ---------------------------------------------------------------------------------------
// Matching-encoder stem, antialiased ResNet-18 (antialiased-cnns 0.3, networks.py:250-270):
//   conv 7x7 stride 2 pad 3 (3 -> 64, BatchNorm folded into weights/bias) + ReLU
// fp32 NCHW image in, NHWC split-bf16 out.  CUDA-core kernel: 16x16 output pixels per CTA, input patch
// and all weights in shared memory, 4 output pixels x 16 channels per thread.
// ---------------------------------------------------------------------------------------
#define STEM_T 16
#define STEM_PATCH (2 * STEM_T + 5)  // 37
__global__ void __launch_bounds__(256)
stem_conv7_kernel(const float* __restrict__ img, const float* __restrict__ wt /*[147][64]*/,
                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ oh, __nv_bfloat16* __restrict__ ol, int H,
                  int W, int OH, int OW) {
  extern __shared__ float sm[];
  float* w_s = sm;                 // [147][64]
  float* p_s = w_s + 147 * 64;     // [3][37][38]
  const int n = blockIdx.z, oy0 = blockIdx.y * STEM_T, ox0 = blockIdx.x * STEM_T;
  for (int i = threadIdx.x; i < 147 * 64; i += 256) w_s[i] = wt[i];
  const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
  for (int i = threadIdx.x; i < 3 * STEM_PATCH * STEM_PATCH; i += 256) {
    const int c = i / (STEM_PATCH * STEM_PATCH), r = (i / STEM_PATCH) % STEM_PATCH, q = i % STEM_PATCH;
    const int y = iy0 + r, x = ix0 + q;
    float v = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W) v = img[((size_t)n * 3 + c) * H * W + (size_t)y * W + x];
    p_s[(c * STEM_PATCH + r) * (STEM_PATCH + 1) + q] = v;
  }
  __syncthreads();
  // thread -> (pixel group of 4 along x, channel group of 16): 64 pixel groups x 4 channel groups
  const int cgp = threadIdx.x & 3, pg = threadIdx.x >> 2;
  const int py = pg >> 2, px4 = (pg & 3) * 4;
  float acc[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
  for (int c = 0; c < 3; ++c)
    for (int dy = 0; dy < 7; ++dy) {
      const float* prow = p_s + (c * STEM_PATCH + 2 * py + dy) * (STEM_PATCH + 1) + 2 * px4;
#pragma unroll
      for (int dx = 0; dx < 7; ++dx) {
        const float* wp = w_s + ((c * 7 + dy) * 7 + dx) * 64 + cgp * 16;
        float wv[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 t = *reinterpret_cast<const float4*>(wp + 4 * q);
          wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float xin = prow[2 * i + dx];
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(xin, wv[j], acc[i][j]);
        }
      }
    }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int oy = oy0 + py, ox = ox0 + px4 + i;
    if (oy < OH && ox < OW) {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(acc[i][j] + bias[cgp * 16 + j], 0.f);
      const size_t o = (((size_t)n * OH + oy) * OW + ox) * 64 + cgp * 16;
      store8(oh, ol, o, v);
      store8(oh, ol, o + 8, v + 8);
    }
  }
}

extern "C" int b200_stem_conv7(const float* img, const float* wt, const float* bias, void* out_hi, void* out_lo,
                               int n_img, int H, int W, void* stream) {
  B200_CHECK_ARG(img && wt && bias && out_hi && out_lo && n_img > 0 && H > 0 && W > 0, "stem_conv7: bad arguments");
  const int OH = (H + 6 - 7) / 2 + 1, OW = (W + 6 - 7) / 2 + 1;
  const size_t smem = sizeof(float) * (147 * 64 + 3 * STEM_PATCH * (STEM_PATCH + 1));
  static bool done = false;
  if (!done) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(stem_conv7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done = true;
  }
  dim3 grid((OW + STEM_T - 1) / STEM_T, (OH + STEM_T - 1) / STEM_T, n_img);
  stem_conv7_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(img, wt, bias, (__nv_bfloat16*)out_hi,
                                                              (__nv_bfloat16*)out_lo, H, W, OH, OW);
  B200_CHECK_LAUNCH("stem_conv7");
  return 0;
}

// MaxPool2d(kernel 2, stride 1) followed by BlurPool(filt 4 = [1,3,3,1]^2/64, stride 2, reflect pad
// (1,2,1,2)) -- the anti-aliased "maxpool" of antialiased-cnns 0.3 -- fused; H,W -> H/2,W/2 for even sizes.
//   M[my][mx]  = max of the 2x2 input block at (my, mx)                 (max-pooled map, (H-1) x (W-1))
//   out[oy][ox] = sum_{a,q} f[a] f[q] / 64 * M[refl(2oy+a-1)][refl(2ox+q-1)],  f = [1,3,3,1]
// Two kernels: interior output rows slide down the image (below), the one or two rows whose window reflects at the
// top / bottom edge take the direct form.
__device__ __forceinline__ int mbp_reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

// direct form, one thread = one output pixel x 8 channels; rows = the output rows to compute (n_rows of them)
__global__ void maxblurpool_rows_kernel(const __nv_bfloat16* __restrict__ ih, const __nv_bfloat16* __restrict__ il,
                                        __nv_bfloat16* __restrict__ oh, __nv_bfloat16* __restrict__ ol, int B, int H, int W,
                                        int C, int OH, int OW, int row_a, int row_b_first, int n_rows) {
  const int cg = C >> 3;
  const int MH = H - 1, MW = W - 1;
  const size_t total = (size_t)B * n_rows * OW * cg;
  const float f[4] = {1.f, 3.f, 3.f, 1.f};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cg);
    size_t r = i / cg;
    const int ox = (int)(r % OW);
    r /= OW;
    const int ri = (int)(r % n_rows);
    const int b = (int)(r / n_rows);
    const int oy = ri == 0 ? row_a : row_b_first + ri - 1;  // row_a (the top row), then row_b_first, row_b_first+1, ...
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int a = 0; a < 4; ++a) {
      const int my = mbp_reflect(2 * oy + a - 1, MH);
      float hb[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) hb[e] = 0.f;
      for (int q = 0; q < 4; ++q) {
        const int mx = mbp_reflect(2 * ox + q - 1, MW);
        float m[8], t[8];
        const size_t base = ((size_t)b * H + my) * W + mx;
        load8(ih, il, base * C + c8 * 8, m);
        load8(ih, il, (base + 1) * C + c8 * 8, t);
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], t[e]);
        load8(ih, il, (base + W) * C + c8 * 8, t);
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], t[e]);
        load8(ih, il, (base + W + 1) * C + c8 * 8, t);
#pragma unroll
        for (int e = 0; e < 8; ++e) hb[e] = fmaf(f[q], fmaxf(m[e], t[e]), hb[e]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(f[a], hb[e], acc[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= (1.f / 64.f);
    store8(oh, ol, (((size_t)b * OH + oy) * OW + ox) * C + c8 * 8, acc);
  }
}

// Interior rows [oy_lo, oy_hi): one thread = one output column x 8 channels x MBP_SEG output rows, sliding down the
// image.  Per input row it loads the 5 (8 at a reflecting column) pixels of its window once, keeps the horizontal
// pair-maxima of the previous row and the horizontally blurred last four max-pooled rows in registers, and emits an
// output every second row: 2 x 5 pixel loads per output instead of 25, every reduction in a fixed order.
#define MBP_SEG 16
struct MbpRow { float v[4][8]; };
__device__ __forceinline__ void mbp_load_row(const __nv_bfloat16* ih, const __nv_bfloat16* il, size_t row_base, int C,
                                             int c8, bool fast, int x0, const int mc[4], MbpRow& hm) {
  if (fast) {  // columns x0 .. x0+4, neighbours share pixels
    float px[5][8];
#pragma unroll
    for (int q = 0; q < 5; ++q) load8(ih, il, (row_base + x0 + q) * C + c8 * 8, px[q]);
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int e = 0; e < 8; ++e) hm.v[q][e] = fmaxf(px[q][e], px[q + 1][e]);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float a[8], t[8];
      load8(ih, il, (row_base + mc[q]) * C + c8 * 8, a);
      load8(ih, il, (row_base + mc[q] + 1) * C + c8 * 8, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) hm.v[q][e] = fmaxf(a[e], t[e]);
    }
  }
}
__global__ void __launch_bounds__(256, 2)
maxblurpool_slide_kernel(const __nv_bfloat16* __restrict__ ih, const __nv_bfloat16* __restrict__ il,
                         __nv_bfloat16* __restrict__ oh, __nv_bfloat16* __restrict__ ol, int B, int H, int W, int C,
                         int OH, int OW, int oy_lo, int oy_hi) {
  const int cg = C >> 3;
  const int MW = W - 1;
  const int nseg = (oy_hi - oy_lo + MBP_SEG - 1) / MBP_SEG;
  const size_t total = (size_t)B * nseg * OW * cg;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cg);
    size_t r = i / cg;
    const int ox = (int)(r % OW);
    r /= OW;
    const int seg = (int)(r % nseg);
    const int b = (int)(r / nseg);
    const int oy0 = oy_lo + seg * MBP_SEG, oy1 = min(oy0 + MBP_SEG, oy_hi);
    int mc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) mc[q] = mbp_reflect(2 * ox + q - 1, MW);
    const bool fast = ox >= 1 && 2 * ox + 2 < MW;
    const int x0 = 2 * ox - 1;
    const int m0 = 2 * oy0 - 1;              // first max-pooled row of the segment (>= 1: interior rows only)
    const int jmax = 2 * (oy1 - oy0) + 1;    // last max-pooled row counter (row m0 + jmax)
    const size_t img = (size_t)b * H;
    MbpRow prev, cur;
    float hb[4][8];
    mbp_load_row(ih, il, (img + m0) * W, C, c8, fast, x0, mc, prev);
    // step(S): consume input row m0 + j + 1 -> max-pooled row j -> horizontally blurred row into slot S = j & 3
#define MBP_STEP(S)                                                                           \
    {                                                                                         \
      mbp_load_row(ih, il, (img + m0 + j + 1) * W, C, c8, fast, x0, mc, cur);                 \
      _Pragma("unroll") for (int e = 0; e < 8; ++e) {                                         \
        const float t0 = fmaxf(prev.v[0][e], cur.v[0][e]), t1 = fmaxf(prev.v[1][e], cur.v[1][e]); \
        const float t2 = fmaxf(prev.v[2][e], cur.v[2][e]), t3 = fmaxf(prev.v[3][e], cur.v[3][e]); \
        hb[S][e] = (t0 + t3) + 3.f * (t1 + t2);                                               \
      }                                                                                       \
      prev = cur;                                                                             \
    }
#define MBP_EMIT(S0, S1, S2, S3)                                                              \
    {                                                                                         \
      float o[8];                                                                             \
      _Pragma("unroll") for (int e = 0; e < 8; ++e)                                           \
        o[e] = ((hb[S0][e] + hb[S3][e]) + 3.f * (hb[S1][e] + hb[S2][e])) * (1.f / 64.f);      \
      const int oy = oy0 + ((j - 3) >> 1);                                                    \
      store8(oh, ol, (((size_t)b * OH + oy) * OW + ox) * C + c8 * 8, o);                      \
    }
    for (int j = 0; j <= jmax;) {  // four max-pooled rows per trip: slots 0..3, outputs after slots 1 (j >= 5) and 3
      MBP_STEP(0); ++j;
      if (j > jmax) break;
      MBP_STEP(1);
      if (j >= 5) MBP_EMIT(2, 3, 0, 1);
      ++j;
      if (j > jmax) break;
      MBP_STEP(2); ++j;
      if (j > jmax) break;
      MBP_STEP(3);
      MBP_EMIT(0, 1, 2, 3);
      ++j;
    }
#undef MBP_STEP
#undef MBP_EMIT
  }
}

extern "C" int b200_maxblurpool(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int B, int H, int W,
                                int C, void* stream) {
  B200_CHECK_ARG(in_hi && in_lo && out_hi && out_lo && B > 0 && H > 2 && W > 2, "maxblurpool: bad arguments");
  B200_CHECK_ARG(C % 8 == 0, "maxblurpool: C must be a multiple of 8");
  const int OH = (H - 1 + 3 - 4) / 2 + 1, OW = (W - 1 + 3 - 4) / 2 + 1;
  const int MH = H - 1;
  // interior output rows: the 4-row window 2oy-1 .. 2oy+2 of the max-pooled map needs no reflection
  int oy_hi = OH;
  while (oy_hi > 1 && 2 * (oy_hi - 1) + 2 >= MH) --oy_hi;
  const int oy_lo = 1;
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16 *ih = (const __nv_bfloat16*)in_hi, *il = (const __nv_bfloat16*)in_lo;
  __nv_bfloat16 *oh = (__nv_bfloat16*)out_hi, *ol = (__nv_bfloat16*)out_lo;
  if (oy_hi > oy_lo) {
    const int nseg = (oy_hi - oy_lo + MBP_SEG - 1) / MBP_SEG;
    const size_t total = (size_t)B * nseg * OW * (C / 8);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    maxblurpool_slide_kernel<<<blocks, 256, 0, st>>>(ih, il, oh, ol, B, H, W, C, OH, OW, oy_lo, oy_hi);
  }
  const int row_b_first = oy_hi > oy_lo ? oy_hi : 1;      // rows [row_b_first, OH) reflect at the bottom edge
  const int n_rows = 1 + (OH - row_b_first);              // + row 0 (top edge)
  const size_t total = (size_t)B * n_rows * OW * (C / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  maxblurpool_rows_kernel<<<blocks, 256, 0, st>>>(ih, il, oh, ol, B, H, W, C, OH, OW, 0, row_b_first, n_rows);
  B200_CHECK_LAUNCH("maxblurpool");
  return 0;
}

// ---------------------------------------------------------------------------------------
// Temporal prior (BDModel.sample_prior, experiment_modules/bd_model.py:395-410): back-project every pixel of
// the current frame at its rendered depth (BackprojectDepth, utils/geometry_utils.py:54-63), project it into
// the previous frame (Project3D, :76-89, z clamped to 1e-5), and fetch the previous prediction with
// F.grid_sample(mode="nearest", zeros padding, align_corners=False); pixels with rendered depth <= 0 get -1
// (the z > 0 half of the reference's mask is always true because of the clamp).
//   P    [B, 12] = (K @ prior_cam_T_world @ world_T_cam)[:3, :4] row-major
//   invK [B, 16] row-major 4x4 (only the upper 3x3 is read)
// ---------------------------------------------------------------------------------------
__global__ void sample_prior_kernel(const float* __restrict__ depth, const float* __restrict__ prior,
                                    const float* __restrict__ P, const float* __restrict__ invK,
                                    float* __restrict__ out, int B, int H, int W) {
  const int HW = H * W;
  const size_t total = (size_t)B * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW), p = (int)(i - (size_t)b * HW);
    const int y = p / W, x = p - y * W;
    const float* Kb = invK + b * 16;
    const float* Pb = P + b * 12;
    const float pxc = x + 0.5f, pyc = y + 0.5f;
    const float d = depth[i];
    float X[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) X[r] = d * (Kb[4 * r] * pxc + Kb[4 * r + 1] * pyc + Kb[4 * r + 2]);
    float c[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) c[r] = Pb[4 * r] * X[0] + Pb[4 * r + 1] * X[1] + Pb[4 * r + 2] * X[2] + Pb[4 * r + 3];
    const float z = fmaxf(c[2], 1e-5f);
    // normalise exactly like the reference ((p / size - 0.5) * 2), un-normalise like ATen (((g + 1) * size - 1) / 2)
    const float gx = (__fdiv_rn(__fdiv_rn(c[0], z), (float)W) - 0.5f) * 2.f;
    const float gy = (__fdiv_rn(__fdiv_rn(c[1], z), (float)H) - 0.5f) * 2.f;
    const float ix = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), -1.f), 0.5f);
    const float iy = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), -1.f), 0.5f);
    const float rx = nearbyintf(fminf(fmaxf(ix, -2.f), (float)W + 1.f));  // round-half-even like std::nearbyint
    const float ry = nearbyintf(fminf(fmaxf(iy, -2.f), (float)H + 1.f));
    float v = 0.f;
    if (rx >= 0.f && rx < (float)W && ry >= 0.f && ry < (float)H && ix == ix && iy == iy)
      v = prior[(size_t)b * HW + (int)ry * W + (int)rx];
    out[i] = d > 0.f ? v : -1.f;
  }
}

extern "C" int b200_sample_prior(const float* rendered_depth, const float* prior_prediction, const float* P,
                                 const float* invK, float* out, int B, int H, int W, void* stream) {
  B200_CHECK_ARG(rendered_depth && prior_prediction && P && invK && out && B > 0 && H > 0 && W > 0,
                 "sample_prior: bad arguments");
  int blocks = (int)(((size_t)B * H * W + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  sample_prior_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(rendered_depth, prior_prediction, P, invK, out, B, H, W);
  B200_CHECK_LAUNCH("sample_prior");
  return 0;
}
