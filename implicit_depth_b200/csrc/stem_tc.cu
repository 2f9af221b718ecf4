// Matching-encoder stem on the tensor cores: Conv2d(3, 64, 7, stride 2, pad 3) with BatchNorm folded + ReLU
// (modules/networks.py:250-266: conv1, bn1, relu of the antialiased ResNet-18), fp32 NCHW image in,
// NHWC split-bf16 out.
//
// Implicit GEMM with K = 3 channels x 7 kernel rows x 8 (7 kernel columns + 1 zero) = 168 (padded to 192):
// for output pixel (oy, ox) the K-slice of (channel c, kernel row dy) is 8 CONSECUTIVE input floats
// img[c, 2oy-3+dy, 2ox-3 .. 2ox+4], so a thread builds its row of the A operand with plain vector reads of a
// shared-memory image patch -- no im2col buffer.  Same structure as the feature-volume kernel (fv_tc.cu):
// two groups of 128 threads (thread = output pixel = TMEM lane) write their rows, split into bf16 hi/lo,
// STRAIGHT INTO TENSOR MEMORY through a 2-slot ring of 64-wide K chunks; warp 8 issues the MMAs with the weights
// resident in shared memory as [W_hi | W_lo] tiles:  A_hi x [W_hi | W_lo]  is one N = 128 instruction (columns
// [0,64) = hi*hi, [64,128) = hi*lo) and  A_lo x W_hi  one N = 64 instruction.  The epilogue adds the two
// column halves and the bias, applies ReLU, splits to hi/lo and writes 128-byte pixel records into a swizzled
// staging tile that one thread hands to the TMA unit (tensor store, border clipping included).
// TMEM per group: [256g, 256g+128) A ring (2 x (32 hi + 32 lo)), [256g+128, 256g+256) accumulator.
#include "common.cuh"
#include "tc.cuh"
#include "tmap.cuh"
#include <string.h>

#define ST_THREADS 288
#define ST_TY 8             // output rows per tile
#define ST_TX 16            // output columns per tile
#define ST_PH (2 * ST_TY + 5)   // 21 input rows
#define ST_PW 40            // 2*16 + 6 = 38 input columns, padded to a multiple of 4 floats
#define ST_KROWS 21         // (channel, kernel row) pairs = 8-wide K slices
#define ST_NCHUNK 3         // 64-wide K chunks (192 >= 168)

struct StemParams {
  CUtensorMap out_hi, out_lo;  // [64, OW, OH, n] bf16, box [64, 16, 8, 1], SWIZZLE_128B
  const float* img;            // [n, 3, H, W]
  const uint8_t* wimage;       // 3 chunks x [hi 64 x 64 | lo 64 x 64] bf16 SW128 = 48 KB
  const float* bias;           // [64]
  int n_img, H, W, OH, OW, tiles_x, tiles_y;
};

struct StemSync {
  uint64_t a_full[2];
  uint64_t a_empty[2];
  uint64_t acc_full;
};

__global__ void __launch_bounds__(ST_THREADS, 1) stem_tc_kernel(const __grid_constant__ StemParams prm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_s = base;                                    // 48 KB
  uint8_t* staging = base + 49152;                        // group g: hi 16 KB | lo 16 KB
  float* patch = reinterpret_cast<float*>(base + 49152 + 65536);  // group g: [3][21][40] floats
  float* bias_s = patch + 2 * 3 * ST_PH * ST_PW;
  StemSync* sync = reinterpret_cast<StemSync*>(bias_s + 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sync + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long total_tiles = (long long)prm.n_img * prm.tiles_y * prm.tiles_x;
  const int G = 2 * gridDim.x;

  for (int i = tid; i < 49152 / 16; i += ST_THREADS)
    reinterpret_cast<uint4*>(w_s)[i] = __ldg(reinterpret_cast<const uint4*>(prm.wimage) + i);
  if (tid < 64) bias_s[tid] = prm.bias[tid];
  if (warp == 8) {
    tc::tmem_alloc(tmem_slot, 512);
    if (lane == 0) {
      tc::prefetch_tmap(&prm.out_hi);
      tc::prefetch_tmap(&prm.out_lo);
      for (int g = 0; g < 2; ++g) {
        tc::mbar_init(&sync[g].a_full[0], 128);
        tc::mbar_init(&sync[g].a_full[1], 128);
        tc::mbar_init(&sync[g].a_empty[0], 1);
        tc::mbar_init(&sync[g].a_empty[1], 1);
        tc::mbar_init(&sync[g].acc_full, 1);
      }
      tc::mbar_fence_init();
    }
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // =============================== row groups ===============================
    const int g = warp >> 2;
    const int row = tid & 127;
    const int ry = row >> 4, rx = row & 15;
    StemSync* gs = &sync[g];
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t a_base = tmem + lane_base + g * 256;
    const uint32_t acc_base = a_base + 128;
    float* my_patch = patch + g * 3 * ST_PH * ST_PW;
    uint8_t* my_stage = staging + g * 32768;
    uint32_t nfill[2] = {0, 0};
    uint32_t tiles = 0;
    // The image patch of tile t+1 is fetched into REGISTERS (20 independent loads per thread) while tile t is built
    // and drained: the global-memory latency of the patch was a quarter of this kernel's stall samples when the loads
    // sat between the two barriers of the same tile.
    constexpr int PER_THREAD = (3 * ST_PH * ST_PW + 127) / 128;
    float pv[PER_THREAD];
    auto fetch_patch = [&](long long tt) {
      const int ftx = (int)(tt % prm.tiles_x);
      const int fty = (int)((tt / prm.tiles_x) % prm.tiles_y);
      const int fn = (int)(tt / ((long long)prm.tiles_x * prm.tiles_y));
      const int iy0 = fty * ST_TY * 2 - 3, ix0 = ftx * ST_TX * 2 - 3;
      const float* im = prm.img + (size_t)fn * 3 * prm.H * prm.W;
#pragma unroll
      for (int j = 0; j < PER_THREAD; ++j) {
        const int i = row + 128 * j;
        const int c = i / (ST_PH * ST_PW), r = (i / ST_PW) % ST_PH, q = i % ST_PW;
        const int y = iy0 + r, x = ix0 + q;
        pv[j] = 0.f;
        if (i < 3 * ST_PH * ST_PW && y >= 0 && y < prm.H && x >= 0 && x < prm.W)
          pv[j] = __ldg(im + ((size_t)c * prm.H + y) * prm.W + x);
      }
    };
    if ((long long)blockIdx.x * 2 + g < total_tiles) fetch_patch((long long)blockIdx.x * 2 + g);
    for (long long t = blockIdx.x * 2 + g; t < total_tiles; t += G, ++tiles) {
      const int tx = (int)(t % prm.tiles_x);
      const int ty = (int)((t / prm.tiles_x) % prm.tiles_y);
      const int n = (int)(t / ((long long)prm.tiles_x * prm.tiles_y));
      // ---- image patch (zero padded, fetched one tile ahead) -> shared memory ----
      tc::named_sync(1 + g, 128);  // everyone is done reading the previous patch
#pragma unroll
      for (int j = 0; j < PER_THREAD; ++j) {
        const int i = row + 128 * j;
        if (i < 3 * ST_PH * ST_PW) my_patch[i] = pv[j];
      }
      tc::named_sync(1 + g, 128);
      if (t + G < total_tiles) fetch_patch(t + G);
      // ---- build the row: 21 slices of 8 floats -> bf16 hi/lo -> TMEM, 64 K values (8 slices) per ring slot ----
#pragma unroll
      for (int ch = 0; ch < ST_NCHUNK; ++ch) {
        const uint32_t slot = ch & 1;
        if (nfill[slot] >= 1) tc::mbar_wait(&gs->a_empty[slot], (nfill[slot] - 1) & 1u);
        tc::fence_after_sync();
#pragma unroll
        for (int half = 0; half < 2; ++half) {  // 4 slices = 32 K values = 16 columns hi + 16 columns lo
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            const int kr = ch * 8 + half * 4 + s;  // (channel, kernel row) index
            if (kr < ST_KROWS) {
              const int c = kr / 7, dy = kr % 7;
              const float2* p = reinterpret_cast<const float2*>(my_patch + (c * ST_PH + 2 * ry + dy) * ST_PW + 2 * rx);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 v = p[e];
                tc::split2(v.x, v.y, hi[4 * s + e], lo[4 * s + e]);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) hi[4 * s + e] = lo[4 * s + e] = 0u;
            }
          }
          tc::tmem_st16(a_base + slot * 64 + half * 16, hi);
          tc::tmem_st16(a_base + slot * 64 + 32 + half * 16, lo);
        }
        tc::wait_st();
        tc::fence_before_sync();
        tc::mbar_arrive(&gs->a_full[slot]);
        ++nfill[slot];
      }
      // ---- epilogue: relu(main + cross + bias) -> split -> swizzled staging tile -> TMA store ----
      if (row == 0) tc::tma_store_wait_read<0>();  // the previous tile's store has finished reading the staging tile
      tc::named_sync(1 + g, 128);
      tc::mbar_wait(&gs->acc_full, tiles & 1u);
      tc::fence_after_sync();
      uint8_t* row_hi = my_stage + row * 128;
      uint8_t* row_lo = my_stage + 16384 + row * 128;
      const uint32_t swz = (uint32_t)(row & 7);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t rm[32], rc[32];
        tc::tmem_ld32(acc_base + 32 * q, rm);
        tc::tmem_ld32(acc_base + 64 + 32 * q, rc);
        tc::wait_ld();
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 8 * c8 + 2 * e;
            const float v0 = fmaxf(__uint_as_float(rm[j]) + __uint_as_float(rc[j]) + bias_s[32 * q + j], 0.f);
            const float v1 = fmaxf(__uint_as_float(rm[j + 1]) + __uint_as_float(rc[j + 1]) + bias_s[32 * q + j + 1], 0.f);
            tc::split2(v0, v1, hi[e], lo[e]);
          }
          const uint32_t off = ((uint32_t)(4 * q + c8) ^ swz) << 4;
          *reinterpret_cast<uint4*>(row_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(row_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      tc::fence_before_sync();  // order this tile's TMEM reads before the next tile's MMAs (via a_full)
      tc::fence_async_smem();
      tc::named_sync(1 + g, 128);
      if (row == 0) {
        tc::tma_store_4d(&prm.out_hi, my_stage, 0, tx * ST_TX, ty * ST_TY, n);
        tc::tma_store_4d(&prm.out_lo, my_stage + 16384, 0, tx * ST_TX, ty * ST_TY, n);
        tc::tma_store_commit();
      }
    }
    if (row == 0) tc::tma_store_wait_read<0>();  // (the writes complete with the grid)
  } else {
    // =============================== MMA issuer (whole warp loops, one elected lane issues) ==========
    constexpr uint32_t IDESC_MERGED = tc::idesc_bf16_f32(128, 128);
    constexpr uint32_t IDESC_HI = tc::idesc_bf16_f32(128, 64);
    long long t_cur[2];
    uint32_t nfill[2][2] = {{0, 0}, {0, 0}};
    int step[2] = {0, 0};
    for (int g = 0; g < 2; ++g) t_cur[g] = blockIdx.x * 2 + g;
    while (t_cur[0] < total_tiles || t_cur[1] < total_tiles) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (t_cur[g] >= total_tiles) continue;
        StemSync* gs = &sync[g];
        const uint32_t a_base = tmem + g * 256;
        const uint32_t acc = a_base + 128;
        const uint32_t slot = step[g] & 1u;
        if (!__all_sync(0xffffffffu, tc::mbar_try_wait(&gs->a_full[slot], nfill[g][slot] & 1u))) continue;
        tc::fence_after_sync();
        const int c = step[g];
        if (tc::elect_one()) {
          const uint64_t b_m = tc::smem_desc_sw128(tc::smem_u32(w_s + c * 16384));
          const uint32_t a_hi = a_base + slot * 64, a_lo = a_hi + 32;
          const int ksteps = (c == ST_NCHUNK - 1) ? 3 : 4;  // K = 168: the last chunk holds 40 (+8 zero) values
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k < ksteps) {
              tc::mma_ts(acc, a_hi + 8 * k, b_m + 2 * k, IDESC_MERGED, (c == 0 && k == 0) ? 0u : 1u);
              tc::mma_ts(acc, a_lo + 8 * k, b_m + 2 * k, IDESC_HI, 1u);
            }
          }
          tc::mma_commit(&gs->a_empty[slot]);
          if (c + 1 == ST_NCHUNK) tc::mma_commit(&gs->acc_full);
        }
        __syncwarp();
        ++nfill[g][slot];
        if (++step[g] == ST_NCHUNK) {
          step[g] = 0;
          t_cur[g] += G;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

static const size_t ST_SMEM = 1024 + 49152 + 65536 + sizeof(float) * (2 * 3 * ST_PH * ST_PW + 64) + 2 * sizeof(StemSync) + 16;

// img fp32 NCHW [n,3,H,W]; wimage: 48 KB from the packer (3 chunks x [hi 64x64 | lo 64x64] bf16, SW128,
// k = (c*7 + dy)*8 + dx, BatchNorm folded); bias [64]; out NHWC split [n, OH, OW, 64], OH = (H-1)/2+1.
extern "C" int b200_stem_conv7_tc(const float* img, const void* wimage, const float* bias, void* out_hi, void* out_lo,
                                  int n_img, int H, int W, int max_ctas, void* stream) {
  B200_CHECK_ARG(img && wimage && bias && out_hi && out_lo && n_img > 0 && H > 0 && W > 0, "stem_conv7_tc: bad arguments");
  B200_CHECK_ARG((((uintptr_t)wimage | (uintptr_t)out_hi | (uintptr_t)out_lo) & 15) == 0,
                 "stem_conv7_tc: buffers must be 16-byte aligned");
  EncodeTiledFn enc = get_encode();
  B200_CHECK_ARG(enc != nullptr, "stem_conv7_tc: cuTensorMapEncodeTiled not available from the driver");
  StemParams p;
  memset(&p, 0, sizeof(p));
  p.OH = (H + 6 - 7) / 2 + 1;
  p.OW = (W + 6 - 7) / 2 + 1;
  cuuint64_t gdim[4] = {64, (cuuint64_t)p.OW, (cuuint64_t)p.OH, (cuuint64_t)n_img};
  cuuint64_t gstr[3] = {128, (cuuint64_t)p.OW * 128, (cuuint64_t)p.OH * p.OW * 128};
  cuuint32_t box[4] = {64, ST_TX, ST_TY, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int part = 0; part < 2; ++part) {
    CUresult r = enc(part ? &p.out_lo : &p.out_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, part ? out_lo : out_hi, gdim,
                     gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      b200_set_error("stem_conv7_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
      return -2;
    }
  }
  p.img = img;
  p.wimage = (const uint8_t*)wimage;
  p.bias = bias;
  p.n_img = n_img;
  p.H = H;
  p.W = W;
  p.tiles_x = (p.OW + ST_TX - 1) / ST_TX;
  p.tiles_y = (p.OH + ST_TY - 1) / ST_TY;
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    B200_CHECK_CUDA(cudaGetDevice(&dev));
    B200_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    B200_CHECK_CUDA(cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
  }
  const long long total_tiles = (long long)n_img * p.tiles_x * p.tiles_y;
  int grid = n_sm;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  if ((long long)grid * 2 > total_tiles) grid = (int)((total_tiles + 1) / 2);
  stem_tc_kernel<<<grid, ST_THREADS, ST_SMEM, (cudaStream_t)stream>>>(p);
  B200_CHECK_LAUNCH("stem_conv7_tc");
  return 0;
}
