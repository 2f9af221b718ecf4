// Memory-bound pieces of the EfficientNetV2 image encoder's MBConv blocks (torchvision efficientnet_v2_s layout,
// the stand-in for timm's tf_efficientnetv2_s_in21ft1k at the reference call site bd_model.py:46-51), on NHWC
// split-bf16 activations:
//   dwconv3x3_kernel   depthwise 3x3 (stride 1|2; pad (1,1), or (0,1) = TF "SAME" at stride 2) + folded BN bias + SiLU
//   se_pool_kernel     squeeze: per-(frame, channel) mean over the image, deterministic two-level reduction
//   se_fc_kernel       excitation: fc1 + SiLU + fc2 + sigmoid -> scale[b, c]  (fc2 weights transposed)
//   se_scale_kernel    x * scale[b, c]
//   dwconv3x3_pool_kernel / se_fc1_kernel / se_fc2_scale_kernel   the same chain in three launches: squeeze fused into
//                      the depthwise conv, fc1 spread over many blocks, fc2 + sigmoid + scale fused (b200_mbconv_dw_se:
//                      what the encoder plan uses)
// The 1x1 expand / project convolutions and the fused 3x3 convolutions run on the tensor-core conv kernels.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc.cuh"

__device__ __forceinline__ float mb_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float mb_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ void mb_load8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t off, float v[8]) {
  const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi + off));
  const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo + off));
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] = mb_lo(hw[e]) + mb_lo(lw[e]);
    v[2 * e + 1] = mb_hi(hw[e]) + mb_hi(lw[e]);
  }
}
__device__ __forceinline__ void mb_store8(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t off, const float v[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) tc::split2(v[2 * e], v[2 * e + 1], h[e], l[e]);
  *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---------------------------------------------------------------------------------------
// depthwise 3x3: one thread per (output pixel, 8 channels); weights [9][C] tap-major fp32.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dwconv3x3_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                 const float* __restrict__ wt, const float* __restrict__ bias, __nv_bfloat16* __restrict__ oh,
                 __nv_bfloat16* __restrict__ ol, int B, int H, int W, int C, int stride, int OH, int OW, int pad_lo) {
  const int cg = C >> 3;
  const size_t total = (size_t)B * OH * OW * cg;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cg);
    size_t r = i / cg;
    const int ox = (int)(r % OW);
    r /= OW;
    const int oy = (int)(r % OH);
    const int b = (int)(r / OH);
    float acc[8];
    {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8 + 4));
      acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
      acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    }
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int y = oy * stride + dy - pad_lo;
      if (y < 0 || y >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int x = ox * stride + dx - pad_lo;
        if (x < 0 || x >= W) continue;
        float v[8];
        mb_load8(hi, lo, (((size_t)b * H + y) * W + x) * C + c8 * 8, v);
        const float* wp = wt + (size_t)(dy * 3 + dx) * C + c8 * 8;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
        acc[0] = fmaf(v[0], w0.x, acc[0]); acc[1] = fmaf(v[1], w0.y, acc[1]);
        acc[2] = fmaf(v[2], w0.z, acc[2]); acc[3] = fmaf(v[3], w0.w, acc[3]);
        acc[4] = fmaf(v[4], w1.x, acc[4]); acc[5] = fmaf(v[5], w1.y, acc[5]);
        acc[6] = fmaf(v[6], w1.z, acc[6]); acc[7] = fmaf(v[7], w1.w, acc[7]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = silu(acc[e]);
    mb_store8(oh, ol, (((size_t)b * OH + oy) * OW + ox) * C + c8 * 8, acc);
  }
}

extern "C" int b200_dwconv3x3_silu(const void* in_hi, const void* in_lo, const float* wt, const float* bias,
                                   void* out_hi, void* out_lo, int B, int H, int W, int C, int stride, int pad_lo,
                                   void* stream) {
  B200_CHECK_ARG(in_hi && in_lo && wt && bias && out_hi && out_lo, "dwconv3x3: null pointer");
  B200_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && (stride == 1 || stride == 2),
                 "dwconv3x3: bad arguments (C %% 8 == 0, stride 1|2; got C=%d stride=%d)", C, stride);
  B200_CHECK_ARG(pad_lo == 0 || pad_lo == 1, "dwconv3x3: pad_lo is 0 (TF 'SAME' at stride 2 on even sizes) or 1");
  const int OH = (H + pad_lo + 1 - 3) / stride + 1, OW = (W + pad_lo + 1 - 3) / stride + 1;
  const size_t total = (size_t)B * OH * OW * (C / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  dwconv3x3_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in_hi, (const __nv_bfloat16*)in_lo,
                                                            wt, bias, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo,
                                                            B, H, W, C, stride, OH, OW, pad_lo);
  B200_CHECK_LAUNCH("dwconv3x3");
  return 0;
}

// ---------------------------------------------------------------------------------------
// squeeze: mean[b, c] over the HW pixels.  Block = (frame, 64 channels): 32 pixel lanes x 8 channel groups;
// per-thread serial sums over a fixed pixel stride, then a fixed-order tree in shared memory (deterministic).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
se_pool_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, float* __restrict__ mean,
               int HW, int C) {
  __shared__ float red[32][65];
  const int b = blockIdx.y, c0 = blockIdx.x * 64;
  const int cgp = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = c0 + cgp * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    for (int p = pl; p < HW; p += 32) {
      float v[8];
      mb_load8(hi, lo, ((size_t)b * HW + p) * C + c, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[pl][cgp * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
    if (c0 + threadIdx.x < C) mean[(size_t)b * C + c0 + threadIdx.x] = s / (float)HW;
  }
}

// excitation: block = (512-channel slice, frame), 16 warps.  w1 [S][C], b1 [S], w2t [S][C] (fc2 transposed:
// coalesced over channels), b2 [C]  (fp32, S <= 128).  Every block recomputes the small fc1 (S x C MACs) for its
// frame: a warp takes four rows at a time, eight channel steps unrolled = 32 independent weight loads in flight per
// lane (the kernel is pure load latency; it sits on the critical path of every MBConv block).
#define SE_FC_THREADS 512
__global__ void __launch_bounds__(SE_FC_THREADS)
se_fc_kernel(const float* __restrict__ mean, const float* __restrict__ w1, const float* __restrict__ b1,
             const float* __restrict__ w2t, const float* __restrict__ b2, float* __restrict__ scale, int C, int S) {
  __shared__ float s1[128];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* m = mean + (size_t)b * C;
  for (int j0 = warp * 4; j0 < S; j0 += 4 * (SE_FC_THREADS / 32)) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* r0 = w1 + (size_t)min(j0, S - 1) * C;
    const float* r1 = w1 + (size_t)min(j0 + 1, S - 1) * C;
    const float* r2 = w1 + (size_t)min(j0 + 2, S - 1) * C;
    const float* r3 = w1 + (size_t)min(j0 + 3, S - 1) * C;
    int c = lane;
    for (; c + 7 * 32 < C; c += 8 * 32) {
      float mv[8], a0[8], a1[8], a2[8], a3[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        mv[u] = __ldg(m + c + 32 * u);
        a0[u] = __ldg(r0 + c + 32 * u);
        a1[u] = __ldg(r1 + c + 32 * u);
        a2[u] = __ldg(r2 + c + 32 * u);
        a3[u] = __ldg(r3 + c + 32 * u);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc[0] = fmaf(a0[u], mv[u], acc[0]);
        acc[1] = fmaf(a1[u], mv[u], acc[1]);
        acc[2] = fmaf(a2[u], mv[u], acc[2]);
        acc[3] = fmaf(a3[u], mv[u], acc[3]);
      }
    }
    for (; c < C; c += 32) {
      const float mv = __ldg(m + c);
      acc[0] = fmaf(__ldg(r0 + c), mv, acc[0]);
      acc[1] = fmaf(__ldg(r1 + c), mv, acc[1]);
      acc[2] = fmaf(__ldg(r2 + c), mv, acc[2]);
      acc[3] = fmaf(__ldg(r3 + c), mv, acc[3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0 && j0 + i < S) s1[j0 + i] = silu(a + b1[j0 + i]);
    }
  }
  __syncthreads();
  const int c = blockIdx.x * SE_FC_THREADS + threadIdx.x;
  if (c < C) {
    float acc = b2[c];
#pragma unroll 16
    for (int j = 0; j < S; ++j) acc = fmaf(__ldg(w2t + (size_t)j * C + c), s1[j], acc);
    scale[(size_t)b * C + c] = sigmoidf_fast(acc);
  }
}

__global__ void __launch_bounds__(256)
se_scale_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                const float* __restrict__ scale, __nv_bfloat16* __restrict__ oh, __nv_bfloat16* __restrict__ ol,
                int B, int HW, int C) {
  const int cg = C >> 3;
  const size_t total = (size_t)B * HW * cg;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cg);
    const int b = (int)(i / ((size_t)HW * cg));
    float v[8];
    mb_load8(hi, lo, i * 8, v);
    const float* sp = scale + (size_t)b * C + c8 * 8;
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(sp));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(sp + 4));
    v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w;
    v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
    mb_store8(oh, ol, i * 8, v);
  }
}

// Squeeze-and-excitation of one MBConv block (torchvision SqueezeExcitation: avgpool, fc1, SiLU, fc2, sigmoid, scale).
// mean_ws / scale_ws: [B, C] fp32 workspaces.  In place when out == in.
extern "C" int b200_squeeze_excite(const void* in_hi, const void* in_lo, const float* w1, const float* b1,
                                   const float* w2t, const float* b2, float* mean_ws, float* scale_ws, void* out_hi,
                                   void* out_lo, int B, int HW, int C, int S, void* stream) {
  B200_CHECK_ARG(in_hi && in_lo && w1 && b1 && w2t && b2 && mean_ws && scale_ws && out_hi && out_lo,
                 "squeeze_excite: null pointer");
  B200_CHECK_ARG(B > 0 && HW > 0 && C > 0 && C % 8 == 0 && S > 0 && S <= 128,
                 "squeeze_excite: bad sizes (C %% 8 == 0, S <= 128; got C=%d S=%d)", C, S);
  cudaStream_t st = (cudaStream_t)stream;
  se_pool_kernel<<<dim3((C + 63) / 64, B), 256, 0, st>>>((const __nv_bfloat16*)in_hi, (const __nv_bfloat16*)in_lo,
                                                        mean_ws, HW, C);
  se_fc_kernel<<<dim3((C + SE_FC_THREADS - 1) / SE_FC_THREADS, B), SE_FC_THREADS, 0, st>>>(mean_ws, w1, b1, w2t, b2,
                                                                                          scale_ws, C, S);
  const size_t total = (size_t)B * HW * (C / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  se_scale_kernel<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)in_hi, (const __nv_bfloat16*)in_lo, scale_ws,
                                         (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, B, HW, C);
  B200_CHECK_LAUNCH("squeeze_excite");
  return 0;
}

// ---------------------------------------------------------------------------------------
// MBConv middle as one chain of four short kernels: depthwise 3x3 + SiLU with the squeeze fused in, fc1, fc2, scale.
// The separate pool pass (a second read of the expanded activation) is gone, and the excitation no longer recomputes
// fc1 in every block: fc1 is spread over S/4 blocks per frame, fc2 over C/128, so each block streams a few KB of
// weights instead of the whole S x C matrix (the old se_fc_kernel was pure load latency on the critical path of all
// 30 MBConv blocks).  Every reduction has a fixed order: results are deterministic and batch-invariant.
// ---------------------------------------------------------------------------------------
#define DWP_PIX 64  // output pixels per block of dwconv3x3_pool_kernel = pixels per partial sum (2 per thread)
extern "C" int b200_mbconv_pool_block(void) { return DWP_PIX; }

// Block = (64 output pixels, 64 channels, frame): thread = (pixel lane, 8 channels), two pixels per thread.
// partial[b][pixel block][c] = sum of the activation over the block's pixels (per-thread sum in pixel order, then a
// fixed-order tree over the 32 pixel lanes in shared memory).
__global__ void __launch_bounds__(256)
dwconv3x3_pool_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                      const float* __restrict__ wt, const float* __restrict__ bias, __nv_bfloat16* __restrict__ oh,
                      __nv_bfloat16* __restrict__ ol, float* __restrict__ partial, int H, int W, int C, int stride,
                      int OH, int OW, int pad_lo) {
  __shared__ float red[32][65];
  const int b = blockIdx.z, c0 = blockIdx.y * 64;
  const int cgp = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = c0 + cgp * 8;
  float pool[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int it = 0; it < DWP_PIX / 32; ++it) {
    const int p = blockIdx.x * DWP_PIX + it * 32 + pl;
    if (c < C && p < OH * OW) {
      const int oy = p / OW, ox = p - oy * OW;
      float acc[8];
      {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
        acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
        acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
      }
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int y = oy * stride + dy - pad_lo;
        if (y < 0 || y >= H) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int x = ox * stride + dx - pad_lo;
          if (x < 0 || x >= W) continue;
          float v[8];
          mb_load8(hi, lo, (((size_t)b * H + y) * W + x) * C + c, v);
          const float* wp = wt + (size_t)(dy * 3 + dx) * C + c;
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
          acc[0] = fmaf(v[0], w0.x, acc[0]); acc[1] = fmaf(v[1], w0.y, acc[1]);
          acc[2] = fmaf(v[2], w0.z, acc[2]); acc[3] = fmaf(v[3], w0.w, acc[3]);
          acc[4] = fmaf(v[4], w1.x, acc[4]); acc[5] = fmaf(v[5], w1.y, acc[5]);
          acc[6] = fmaf(v[6], w1.z, acc[6]); acc[7] = fmaf(v[7], w1.w, acc[7]);
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        acc[e] = silu(acc[e]);
        pool[e] += acc[e];
      }
      mb_store8(oh, ol, ((size_t)b * OH * OW + p) * C + c, acc);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[pl][cgp * 8 + e] = pool[e];  // zeros from threads outside the image / channels
  __syncthreads();
  if (threadIdx.x < 64 && c0 + threadIdx.x < C) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
    partial[((size_t)b * gridDim.x + blockIdx.x) * C + c0 + threadIdx.x] = s;
  }
}

// fc1 + SiLU: block = (4 rows of w1, frame), 8 warps = 4 rows x 2 interleaved halves of the channel range.  The block
// first finishes the squeeze (mean[c] = sum of the nPB partials / HW) into shared memory.
#define FC1_ROWS 4
__global__ void __launch_bounds__(256)
se_fc1_kernel(const float* __restrict__ partial, const float* __restrict__ w1, const float* __restrict__ b1,
              float* __restrict__ s1, int C, int S, int nPB, float inv_hw) {
  extern __shared__ float mean_s[];  // [C]
  __shared__ float part[8];
  const int b = blockIdx.y, j0 = blockIdx.x * FC1_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* pp = partial + (size_t)b * nPB * C;
  // two channels and two partials per step, unrolled x4: 16 independent loads in flight per thread (fixed order)
  for (int c = threadIdx.x; c < C; c += 512) {
    const int c2 = c + 256;
    const bool has2 = c2 < C;
    float a0 = 0.f, a1 = 0.f, e0 = 0.f, e1 = 0.f;
    int pb = 0;
#pragma unroll 4
    for (; pb + 1 < nPB; pb += 2) {
      a0 += __ldg(pp + (size_t)pb * C + c);
      a1 += __ldg(pp + (size_t)(pb + 1) * C + c);
      if (has2) {
        e0 += __ldg(pp + (size_t)pb * C + c2);
        e1 += __ldg(pp + (size_t)(pb + 1) * C + c2);
      }
    }
    if (pb < nPB) {
      a0 += __ldg(pp + (size_t)pb * C + c);
      if (has2) e0 += __ldg(pp + (size_t)pb * C + c2);
    }
    mean_s[c] = (a0 + a1) * inv_hw;
    if (has2) mean_s[c2] = (e0 + e1) * inv_hw;
  }
  __syncthreads();
  const int row = j0 + (warp >> 1), half = warp & 1;
  float a = 0.f;
  if (row < S) {
    const float* r = w1 + (size_t)row * C;
    int c = half * 32 + lane;
    for (; c + 7 * 64 < C; c += 8 * 64) {
      float wv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) wv[u] = __ldg(r + c + 64 * u);
#pragma unroll
      for (int u = 0; u < 8; ++u) a = fmaf(wv[u], mean_s[c + 64 * u], a);
    }
    for (; c < C; c += 64) a = fmaf(__ldg(r + c), mean_s[c], a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) part[warp] = a;
  __syncthreads();
  if (threadIdx.x < FC1_ROWS && j0 + threadIdx.x < S)
    s1[(size_t)b * S + j0 + threadIdx.x] =
        silu(part[2 * threadIdx.x] + part[2 * threadIdx.x + 1] + b1[j0 + threadIdx.x]);
}

// fc2 + sigmoid: block = (128 channels, frame), one channel per thread, w2t [S][C] coalesced over channels.
__global__ void __launch_bounds__(128)
se_fc2_kernel(const float* __restrict__ s1, const float* __restrict__ w2t, const float* __restrict__ b2,
              float* __restrict__ scale, int C, int S) {
  __shared__ float s1s[128];
  const int b = blockIdx.y;
  if (threadIdx.x < S) s1s[threadIdx.x] = s1[(size_t)b * S + threadIdx.x];
  __syncthreads();
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c < C) {
    float acc = b2[c];
#pragma unroll 16
    for (int j = 0; j < S; ++j) acc = fmaf(__ldg(w2t + (size_t)j * C + c), s1s[j], acc);
    scale[(size_t)b * C + c] = sigmoidf_fast(acc);
  }
}

// fc2 + sigmoid + scale in one launch: block = (strip of SEF_PIX pixels, 64 channels, frame).  The block first forms the
// 64 scales of its channels (4 threads per channel over interleaved quarters of the squeeze vector, summed in a fixed
// order), then multiplies its strip in place.  Re-computing the 64 x S dot products per strip is cheaper than a launch.
#define SEF_PIX 128
__global__ void __launch_bounds__(256)
se_fc2_scale_kernel(const float* __restrict__ s1, const float* __restrict__ w2t, const float* __restrict__ b2,
                    __nv_bfloat16* __restrict__ xh, __nv_bfloat16* __restrict__ xl, int HW, int C, int S) {
  __shared__ float s1s[128];
  __shared__ float red[4][64];
  __shared__ float scale_s[64];
  const int b = blockIdx.z, c0 = blockIdx.y * 64;
  if (threadIdx.x < S) s1s[threadIdx.x] = s1[(size_t)b * S + threadIdx.x];
  __syncthreads();
  {
    const int cl = threadIdx.x & 63, part = threadIdx.x >> 6;
    const int c = c0 + cl;
    float acc = 0.f;
    if (c < C) {
#pragma unroll 8
      for (int j = part; j < S; j += 4) acc = fmaf(__ldg(w2t + (size_t)j * C + c), s1s[j], acc);
    }
    red[part][cl] = acc;
  }
  __syncthreads();
  if (threadIdx.x < 64 && c0 + threadIdx.x < C) {
    const int cl = threadIdx.x;
    scale_s[cl] = sigmoidf_fast(((red[0][cl] + red[1][cl]) + (red[2][cl] + red[3][cl])) + b2[c0 + cl]);
  }
  __syncthreads();
  const int c8 = threadIdx.x & 7, prow = threadIdx.x >> 3;
  const int c = c0 + c8 * 8;
  if (c >= C) return;
  float sc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) sc[e] = scale_s[c8 * 8 + e];
  const int p_end = min((int)(blockIdx.x + 1) * SEF_PIX, HW);
  for (int p = blockIdx.x * SEF_PIX + prow; p < p_end; p += 32) {
    const size_t off = ((size_t)b * HW + p) * C + c;
    float v[8];
    mb_load8(xh, xl, off, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= sc[e];
    mb_store8(xh, xl, off, v);
  }
}

// Depthwise 3x3 (+ folded BatchNorm bias, SiLU) followed by the squeeze-and-excitation of one MBConv block
// (torchvision MBConv.block[1:3]): out = dw(x) * sigmoid(fc2(silu(fc1(mean_hw(dw(x)))))).
//   wt [9][C], bias [C]; w1 [S][C], b1 [S], w2t [S][C], b2 [C] (fp32, S <= 128);
//   workspaces: partial_ws [B][ceil(OH*OW / b200_mbconv_pool_block())][C], s1_ws [B][S], scale_ws [B][C].
extern "C" int b200_mbconv_dw_se(const void* in_hi, const void* in_lo, const float* wt, const float* bias,
                                 const float* w1, const float* b1, const float* w2t, const float* b2,
                                 float* partial_ws, float* s1_ws, float* scale_ws, void* out_hi, void* out_lo, int B,
                                 int H, int W, int C, int stride, int S, int pad_lo, void* stream) {
  B200_CHECK_ARG(in_hi && in_lo && wt && bias && w1 && b1 && w2t && b2 && partial_ws && s1_ws && scale_ws && out_hi &&
                     out_lo,
                 "mbconv_dw_se: null pointer");
  B200_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && (stride == 1 || stride == 2),
                 "mbconv_dw_se: bad arguments (C %% 8 == 0, stride 1|2; got C=%d stride=%d)", C, stride);
  B200_CHECK_ARG(S > 0 && S <= 128 && C <= 12288, "mbconv_dw_se: bad sizes (S <= 128, C <= 12288; got C=%d S=%d)", C, S);
  B200_CHECK_ARG(pad_lo == 0 || pad_lo == 1, "mbconv_dw_se: pad_lo is 0 (TF 'SAME' at stride 2 on even sizes) or 1");
  cudaStream_t st = (cudaStream_t)stream;
  const int OH = (H + pad_lo + 1 - 3) / stride + 1, OW = (W + pad_lo + 1 - 3) / stride + 1;
  const int OHW = OH * OW, nPB = (OHW + DWP_PIX - 1) / DWP_PIX;
  dwconv3x3_pool_kernel<<<dim3(nPB, (C + 63) / 64, B), 256, 0, st>>>(
      (const __nv_bfloat16*)in_hi, (const __nv_bfloat16*)in_lo, wt, bias, (__nv_bfloat16*)out_hi,
      (__nv_bfloat16*)out_lo, partial_ws, H, W, C, stride, OH, OW, pad_lo);
  se_fc1_kernel<<<dim3((S + FC1_ROWS - 1) / FC1_ROWS, B), 256, (size_t)C * sizeof(float), st>>>(
      partial_ws, w1, b1, s1_ws, C, S, nPB, 1.f / (float)OHW);
  (void)scale_ws;  // (the scales live in shared memory of the fused fc2 + scale kernel)
  se_fc2_scale_kernel<<<dim3((OHW + SEF_PIX - 1) / SEF_PIX, (C + 63) / 64, B), 256, 0, st>>>(
      s1_ws, w2t, b2, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, OHW, C, S);
  B200_CHECK_LAUNCH("mbconv_dw_se");
  return 0;
}

// ---------------------------------------------------------------------------------------
// Stem of the image-prior encoder: Conv2d(3, 24, 3, stride 2) + folded BatchNorm + SiLU straight from the fp32 NCHW
// image (timm `conv_stem` / `bn1`, TF "SAME" padding (0, 1) on even sizes; torchvision features[0], padding 1).
// K = 27: nothing for a tensor core -- the image -> split-bf16 conversion plus the per-tap tensor-core kernel on a
// channel-padded copy took 133 us, this takes ~10.  Thread = one output pixel x 8 output channels; the 27 x Cp weights
// sit in shared memory; fp32 FMAs in (c, dy, dx) order.
//   img [B, 3, H, W] fp32 (strides sB, sC, sH, sW in elements); w [27][Cp] (k = (c*3 + dy)*3 + dx), bias [Cp];
//   out: split NHWC [B, OH, OW, Cp], OH = (H + pad_lo + 1 - 3) / 2 + 1.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stem3x3_s2_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                  __nv_bfloat16* __restrict__ oh, __nv_bfloat16* __restrict__ ol, int B, int H, int W, int Cp, int OH,
                  int OW, int pad_lo, long long sB, long long sC, long long sH, long long sW) {
  extern __shared__ float w_s[];  // [27][Cp] then bias [Cp]
  for (int i = threadIdx.x; i < 28 * Cp; i += blockDim.x) w_s[i] = i < 27 * Cp ? w[i] : bias[i - 27 * Cp];
  __syncthreads();
  const int cg = Cp >> 3;
  const size_t total = (size_t)B * OH * OW * cg;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cg);
    size_t r = i / cg;
    const int ox = (int)(r % OW);
    r /= OW;
    const int oy = (int)(r % OH);
    const int b = (int)(r / OH);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = w_s[27 * Cp + c8 * 8 + e];
    const float* ib = img + (size_t)b * sB;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int y = 2 * oy + dy - pad_lo;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int x = 2 * ox + dx - pad_lo;
          float v = 0.f;
          if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(ib + c * sC + y * sH + x * sW);
          const float4 w0 = *reinterpret_cast<const float4*>(w_s + ((c * 3 + dy) * 3 + dx) * Cp + c8 * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(w_s + ((c * 3 + dy) * 3 + dx) * Cp + c8 * 8 + 4);
          acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]);
          acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
          acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
          acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
        }
      }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = silu(acc[e]);
    mb_store8(oh, ol, (((size_t)b * OH + oy) * OW + ox) * Cp + c8 * 8, acc);
  }
}

extern "C" int b200_stem3x3_s2_silu(const float* img, const float* w, const float* bias, void* out_hi, void* out_lo,
                                    int B, int H, int W, int Cp, int pad_lo, long long sB, long long sC, long long sH,
                                    long long sW, void* stream) {
  B200_CHECK_ARG(img && w && bias && out_hi && out_lo, "stem3x3_s2_silu: null pointer");
  B200_CHECK_ARG(B > 0 && H > 2 && W > 2 && Cp > 0 && Cp % 8 == 0 && Cp <= 256 && (pad_lo == 0 || pad_lo == 1),
                 "stem3x3_s2_silu: bad arguments (Cp %% 8 == 0, <= 256; pad_lo 0|1; got Cp=%d pad_lo=%d)", Cp, pad_lo);
  const int OH = (H + pad_lo + 1 - 3) / 2 + 1, OW = (W + pad_lo + 1 - 3) / 2 + 1;
  const size_t total = (size_t)B * OH * OW * (Cp / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  stem3x3_s2_kernel<<<blocks, 256, (size_t)28 * Cp * sizeof(float), (cudaStream_t)stream>>>(
      img, w, bias, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, B, H, W, Cp, OH, OW, pad_lo, sB, sC, sH, sW);
  B200_CHECK_LAUNCH("stem3x3_s2_silu");
  return 0;
}

// a + b on split activations (8 channels per thread).
__global__ void __launch_bounds__(256)
split_add_kernel(const __nv_bfloat16* __restrict__ ah, const __nv_bfloat16* __restrict__ al,
                 const __nv_bfloat16* __restrict__ bh, const __nv_bfloat16* __restrict__ bl,
                 __nv_bfloat16* __restrict__ oh, __nv_bfloat16* __restrict__ ol, size_t n8) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    float a[8], b[8];
    mb_load8(ah, al, i * 8, a);
    mb_load8(bh, bl, i * 8, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] += b[e];
    mb_store8(oh, ol, i * 8, a);
  }
}

extern "C" int b200_split_add(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, void* out_hi,
                              void* out_lo, long long n, void* stream) {
  B200_CHECK_ARG(a_hi && a_lo && b_hi && b_lo && out_hi && out_lo && n > 0 && n % 8 == 0, "split_add: bad arguments");
  const size_t n8 = (size_t)n / 8;
  int blocks = (int)((n8 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  split_add_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)a_hi, (const __nv_bfloat16*)a_lo, (const __nv_bfloat16*)b_hi, (const __nv_bfloat16*)b_lo,
      (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, n8);
  B200_CHECK_LAUNCH("split_add");
  return 0;
}

// ---------------------------------------------------------------------------------------
// 1x1 convolution to ONE channel + exp: the last layer of the regression heads of DepthDecoderPP
// (modules/networks.py:160-163: nn.Conv2d(C, 1, 1)) and SkipDecoderRegression (networks_fast.py:106-112), with the
// exp() of DepthModel.forward (depth_model.py:426-435).  8 lanes per pixel, fixed-order shuffle reduction.
//   in: split NHWC [n_pix, C]; w [C], bias [1]; out_log / out_exp: [n_pix] fp32 (= [B,1,H,W]).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
channel_dot_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                   const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out_log,
                   float* __restrict__ out_exp, long long n_pix, int C) {
  const int sub = threadIdx.x & 7;
  const long long pix = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3;
  const long long p = pix < n_pix ? pix : n_pix - 1;  // keep the 8-lane group convergent
  float acc = 0.f;
  for (int c = sub * 8; c < C; c += 64) {
    float v[8];
    mb_load8(hi, lo, (size_t)p * C + c, v);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + c + 4));
    acc = fmaf(v[0], w0.x, acc); acc = fmaf(v[1], w0.y, acc); acc = fmaf(v[2], w0.z, acc); acc = fmaf(v[3], w0.w, acc);
    acc = fmaf(v[4], w1.x, acc); acc = fmaf(v[5], w1.y, acc); acc = fmaf(v[6], w1.z, acc); acc = fmaf(v[7], w1.w, acc);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (sub == 0 && pix < n_pix) {
    const float v = acc + bias[0];
    out_log[pix] = v;
    if (out_exp) out_exp[pix] = expf(v);
  }
}

extern "C" int b200_channel_dot_exp(const void* in_hi, const void* in_lo, const float* w, const float* bias,
                                    float* out_log, float* out_exp, long long n_pix, int C, void* stream) {
  B200_CHECK_ARG(in_hi && in_lo && w && bias && out_log && n_pix > 0 && C > 0 && C % 8 == 0,
                 "channel_dot_exp: bad arguments (C %% 8 == 0; got C=%d)", C);
  const long long threads = n_pix * 8;
  channel_dot_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)in_hi, (const __nv_bfloat16*)in_lo, w, bias, out_log, out_exp, n_pix, C);
  B200_CHECK_LAUNCH("channel_dot_exp");
  return 0;
}

// ---------------------------------------------------------------------------------------
// Output side of the evaluation scripts: sigmoid_custom (modules/layers.py:138-139) followed by
// F.interpolate to the ground-truth size, bilinear (align_corners=False) or nearest (test_bd.py:225-243,
// inference/inference.py:159-162), one pass.  in [N, h, w] -> out [N, H, W] fp32 (N = B * planes).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sigmoid_resize_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int h, int w, int H, int W,
                      float multiplier, int nearest, int apply_sigmoid) {
  const size_t total = (size_t)N * H * W;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;  // ATen area_pixel_compute_scale (no align_corners)
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % W);
    const int Y = (int)((i / W) % H);
    const float* src = in + (i / ((size_t)W * H)) * (size_t)h * w;
    float v;
    if (nearest) {  // ATen nearest_neighbor_compute_source_index: min(floor(dst * scale), size - 1)
      const int y = min((int)floorf(Y * sy), h - 1), x = min((int)floorf(X * sx), w - 1);
      v = src[y * w + x];
      if (apply_sigmoid) v = 1.f / (1.f + expf(-multiplier * v));
    } else {        // ATen area_pixel_compute_source_index: max((dst + 0.5) * scale - 0.5, 0)
      const float fy = fmaxf((Y + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf((X + 0.5f) * sx - 0.5f, 0.f);
      const int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
      const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
      const float ly = fy - (float)y0, lx = fx - (float)x0;
      float t00 = src[y0 * w + x0], t01 = src[y0 * w + x1], t10 = src[y1 * w + x0], t11 = src[y1 * w + x1];
      if (apply_sigmoid) {  // the reference applies the sigmoid BEFORE interpolating
        t00 = 1.f / (1.f + expf(-multiplier * t00));
        t01 = 1.f / (1.f + expf(-multiplier * t01));
        t10 = 1.f / (1.f + expf(-multiplier * t10));
        t11 = 1.f / (1.f + expf(-multiplier * t11));
      }
      v = (1.f - ly) * ((1.f - lx) * t00 + lx * t01) + ly * ((1.f - lx) * t10 + lx * t11);
    }
    out[i] = v;
  }
}

extern "C" int b200_sigmoid_resize(const float* in, float* out, int N, int h, int w, int H, int W, float multiplier,
                                   int nearest, int apply_sigmoid, void* stream) {
  B200_CHECK_ARG(in && out && N > 0 && h > 0 && w > 0 && H > 0 && W > 0, "sigmoid_resize: bad arguments");
  const size_t total = (size_t)N * H * W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  sigmoid_resize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, out, N, h, w, H, W, multiplier, nearest,
                                                                 apply_sigmoid);
  B200_CHECK_LAUNCH("sigmoid_resize");
  return 0;
}
