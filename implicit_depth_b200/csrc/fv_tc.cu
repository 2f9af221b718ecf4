// Feature-volume MLP (mlp_feature_volume) on the 5th-generation tensor cores.
//
// Persistent kernel, one CTA per SM, 9 warps:
//   warps 0-3 and 4-7  two independent "row groups"; thread i of a group owns row i of the group's
//                      current 128-row tile (= 128 consecutive pixels at one depth plane) and TMEM
//                      lane i.  It gathers/warps the source features, builds the row's 22K+20
//                      input channels in registers, splits them into bf16 hi/lo and stores them
//                      straight into TENSOR MEMORY as the A operand (tcgen05.st) -- the 2.5 GB
//                      MLP-input tensor of the reference never exists anywhere.
//   warp 8             one elected thread issues every tcgen05.mma: layer 1 as A(TMEM) x W1(smem),
//                      3 bf16 passes (hi*hi, hi*lo, lo*hi) per 16-wide k-step so the result is
//                      fp32-grade; layer 2 the same with H1 re-stored to TMEM by the row threads.
// Accumulators live in TMEM (128 columns per group); both weight matrices stay resident in shared
// memory as pre-swizzled split-bf16 images (160 KB) for the life of the CTA.  While one group is in
// an epilogue or gathering, the other group's MMAs keep the tensor pipe busy.
//
// TMEM map (512 columns): group g uses [256g, 256g+128) as A ring (2 slots x (32 hi + 32 lo)) and,
// after layer 1, as H1 (64 hi + 64 lo); [256g+128, 256g+256) is the fp32 accumulator.
//
// Replaces FeatureVolumeManager.build_cost_volume (modules/cost_volume.py:437-706) /
// FastFeatureVolumeManager.build_cost_volume (:938-1146) + MLP (modules/networks.py:218-233).
#include "fv_rows.cuh"
#include "tc.cuh"

#define FVT_THREADS 288
#define FVT_ROWS 128

struct FvTcParams {
  const float* cur;       // [B,N,16]
  const float* src;       // [B,K,N,16]
  const float* cams;      // [B,K,32]
  const float* invK;      // [B,4,4]
  const float* planes;    // [B,D]
  const float* bias_eff;  // [B,128]
  const uint8_t* wimage;  // smem image: W1 hi chunks | W1 lo chunks | W2 hi (2) | W2 lo (2), 16 KB each
  const float* b2;        // [128]
  const float* w3;        // [128]
  const float* b3;        // [1]
  float* vol;             // [B,D,N]
  unsigned char* mask;    // [B,N] or null
  int B, D, h, w;
};

struct GroupSync {
  uint64_t a_full[2];
  uint64_t a_empty[2];
  uint64_t acc_full;
  uint64_t h_full;
  uint64_t acc2_full;
};

template <int K>
struct FvCfg {
  static constexpr int KIN = FV_VIEW_CH * K + FV_TAIL_CH;
  static constexpr int NCHUNK = (KIN + 63) / 64;
  static constexpr int W_BYTES = (2 * NCHUNK + 4) * 16384;
};

// Flush one 64-channel chunk of the row into ring slot `slot` of the group's A region.
// `n` counts how often this slot has been filled before (by this group, over all tiles).
__device__ __forceinline__ void flush_chunk(const float (&buf)[64], uint32_t a_base, GroupSync* gs, uint32_t slot,
                                            uint32_t& n) {
  if (n >= 1) tc::mbar_wait(&gs->a_empty[slot], (n - 1) & 1u);
  tc::fence_after_sync();
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int j = 0; j < 16; ++j) tc::split2(buf[32 * half + 2 * j], buf[32 * half + 2 * j + 1], hi[j], lo[j]);
    tc::tmem_st16(a_base + slot * 64 + half * 16, hi);
    tc::tmem_st16(a_base + slot * 64 + 32 + half * 16, lo);
  }
  tc::wait_st();
  tc::fence_before_sync();
  tc::mbar_arrive(&gs->a_full[slot]);
  ++n;
}

template <int K>
__global__ void __launch_bounds__(FVT_THREADS, 1) fv_tc_kernel(const FvTcParams prm) {
  using Cfg = FvCfg<K>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* wbase = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w1_hi = wbase;
  uint8_t* w1_lo = w1_hi + Cfg::NCHUNK * 16384;
  uint8_t* w2_hi = w1_lo + Cfg::NCHUNK * 16384;
  uint8_t* w2_lo = w2_hi + 2 * 16384;
  float* fbase = reinterpret_cast<float*>(wbase + Cfg::W_BYTES);
  float* b2_s = fbase;                 // [128]
  float* w3_s = b2_s + 128;            // [128]
  float* bias_s = w3_s + 128;          // [2][128]
  float* cam_s = bias_s + 256;         // [2][8*32]
  float* invk_s = cam_s + 2 * B200_MAX_VIEWS * B200_CAM_STRIDE;  // [2][12]
  GroupSync* gsync = reinterpret_cast<GroupSync*>(invk_s + 24);  // [2], 8-byte aligned (offsets are multiples of 8 B)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gsync + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = prm.h * prm.w;
  const int NB = (N + FVT_ROWS - 1) / FVT_ROWS;
  const long long total_tiles = (long long)prm.B * NB * prm.D;
  const int n_groups = 2 * gridDim.x;

  // ---- one-time set-up ----
  for (int i = tid; i < Cfg::W_BYTES / 16; i += FVT_THREADS)
    reinterpret_cast<uint4*>(wbase)[i] = __ldg(reinterpret_cast<const uint4*>(prm.wimage) + i);
  if (tid < 128) {
    b2_s[tid] = prm.b2[tid];
    w3_s[tid] = prm.w3[tid];
  }
  if (warp == 8) {
    tc::tmem_alloc(tmem_slot, 512);
    if (lane == 0) {
      for (int g = 0; g < 2; ++g) {
        tc::mbar_init(&gsync[g].a_full[0], 128);
        tc::mbar_init(&gsync[g].a_full[1], 128);
        tc::mbar_init(&gsync[g].a_empty[0], 1);
        tc::mbar_init(&gsync[g].a_empty[1], 1);
        tc::mbar_init(&gsync[g].acc_full, 1);
        tc::mbar_init(&gsync[g].h_full, 128);
        tc::mbar_init(&gsync[g].acc2_full, 1);
      }
      tc::mbar_fence_init();
    }
  }
  tc::fence_async_smem();  // weight image written with generic stores, read by tcgen05.mma
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // =============================== row groups ===============================
    const int g = warp >> 2;
    const int row = tid & 127;
    const int gid = blockIdx.x * 2 + g;
    const long long t_begin = total_tiles * gid / n_groups;
    const long long t_end = total_tiles * (gid + 1) / n_groups;
    GroupSync* gs = &gsync[g];
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t a_base = tmem + lane_base + g * 256;
    const uint32_t acc_base = a_base + 128;
    float* my_bias = bias_s + g * 128;
    float* my_cam = cam_s + g * B200_MAX_VIEWS * B200_CAM_STRIDE;
    float* my_invk = invk_s + g * 12;
    const float b3v = prm.b3[0];

    uint32_t nfill0 = 0, nfill1 = 0;  // fills so far of ring slot 0 / 1 (chunk c of a tile uses slot c & 1)
    uint32_t tiles = 0;               // tiles done so far by this group
    int cur_b = -1;
    for (long long t = t_begin; t < t_end; ++t, ++tiles) {
      const int d = (int)(t % prm.D);
      const long long pbq = t / prm.D;
      const int pb = (int)(pbq % NB);
      const int b = (int)(pbq / NB);
      if (b != cur_b) {  // uniform over the group: per-frame tables
        tc::named_sync(1 + g, 128);
        my_bias[row] = prm.bias_eff[b * 128 + row];
        for (int i = row; i < K * B200_CAM_STRIDE; i += 128) my_cam[i] = prm.cams[(size_t)b * K * B200_CAM_STRIDE + i];
        if (row < 9) my_invk[row] = prm.invK[b * 16 + (row / 3) * 4 + row % 3];
        tc::named_sync(1 + g, 128);
        cur_b = b;
      }
      const int p_raw = pb * FVT_ROWS + row;
      const int p = min(p_raw, N - 1);
      const int y = p / prm.w, x = p - y * prm.w;
      const float zd = prm.planes[b * prm.D + d];
      const PixelCtx pc = make_pixel_ctx(x, y, my_invk);
      float c16[16];
      {
        const float* cp = prm.cur + ((size_t)b * N + p) * B200_FEAT_C;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 t4 = ldg4(cp + 4 * v);
          c16[4 * v] = t4.x; c16[4 * v + 1] = t4.y; c16[4 * v + 2] = t4.z; c16[4 * v + 3] = t4.w;
        }
      }
      // ---- build the row, 64 channels at a time, straight into tensor memory ----
      float buf[64];
      bool inb = false;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float out[FV_VIEW_CH];
        inb |= fv_view_block(pc, my_cam + k * B200_CAM_STRIDE, prm.src + ((size_t)b * K + k) * N * B200_FEAT_C, c16,
                             zd, prm.h, prm.w, out);
#pragma unroll
        for (int c = 0; c < FV_VIEW_CH; ++c) {
          const int ch = k * FV_VIEW_CH + c;
          buf[ch & 63] = out[c];
          if ((ch & 63) == 63) {
            if ((ch >> 6) & 1) flush_chunk(buf, a_base, gs, 1u, nfill1);
            else flush_chunk(buf, a_base, gs, 0u, nfill0);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < Cfg::NCHUNK * 64 - K * FV_VIEW_CH; ++c) {
        const int ch = K * FV_VIEW_CH + c;
        float v = 0.f;  // K padding
        if (c < 16) v = c16[c];
        else if (c < 19) v = pc.curray[c - 16];
        else if (c == 19) v = zd;
        buf[ch & 63] = v;
        if ((ch & 63) == 63) {
          if ((ch >> 6) & 1) flush_chunk(buf, a_base, gs, 1u, nfill1);
          else flush_chunk(buf, a_base, gs, 0u, nfill0);
        }
      }
      if (prm.mask != nullptr && d == prm.D - 1 && p_raw < N) prm.mask[(size_t)b * N + p] = inb ? 1 : 0;

      // ---- epilogue 1: H1 = lrelu(acc + bias_eff) -> split -> TMEM (over the A ring) ----
      tc::mbar_wait(&gs->acc_full, tiles & 1u);
      tc::fence_after_sync();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t r[32];
        tc::tmem_ld32(acc_base + 32 * q, r);
        tc::wait_ld();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = 32 * q + 2 * j;
          const float v0 = leaky(__uint_as_float(r[2 * j]) + my_bias[n], 0.01f);
          const float v1 = leaky(__uint_as_float(r[2 * j + 1]) + my_bias[n + 1], 0.01f);
          tc::split2(v0, v1, hi[j], lo[j]);
        }
        tc::tmem_st16(a_base + 16 * q, hi);
        tc::tmem_st16(a_base + 64 + 16 * q, lo);
      }
      tc::wait_st();
      tc::fence_before_sync();
      tc::mbar_arrive(&gs->h_full);

      // ---- epilogue 2: out = lrelu(acc2 + b2) . w3 + b3 ----
      tc::mbar_wait(&gs->acc2_full, tiles & 1u);
      tc::fence_after_sync();
      float o = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t r[32];
        tc::tmem_ld32(acc_base + 32 * q, r);
        tc::wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          o = fmaf(leaky(__uint_as_float(r[j]) + b2_s[32 * q + j], 0.01f), w3_s[32 * q + j], o);
      }
      if (p_raw < N) prm.vol[((size_t)b * prm.D + d) * N + p] = o + b3v;
      tc::fence_before_sync();  // order this tile's TMEM reads before the next tile's MMAs (via a_full)
    }
  } else {
    // =============================== MMA issuer (whole warp loops, one elected lane issues) ==========
    constexpr uint32_t IDESC = tc::idesc_bf16_f32(128, 128);
    long long t_cur[2], t_end[2];
    uint32_t nfill[2][2] = {{0, 0}, {0, 0}}, tiles[2] = {0, 0};
    int step[2] = {0, 0};
    for (int g = 0; g < 2; ++g) {
      const int gid = blockIdx.x * 2 + g;
      t_cur[g] = total_tiles * gid / n_groups;
      t_end[g] = total_tiles * (gid + 1) / n_groups;
    }
    while (t_cur[0] < t_end[0] || t_cur[1] < t_end[1]) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (t_cur[g] >= t_end[g]) continue;
        GroupSync* gs = &gsync[g];
        const uint32_t a_base = tmem + g * 256;
        const uint32_t acc = a_base + 128;
        if (step[g] < Cfg::NCHUNK) {
          const uint32_t slot = step[g] & 1u;
          if (!__all_sync(0xffffffffu, tc::mbar_try_wait(&gs->a_full[slot], nfill[g][slot] & 1u))) continue;
          tc::fence_after_sync();
          const int c = step[g];
          if (tc::elect_one()) {
            const uint64_t b_hi = tc::smem_desc_sw128(tc::smem_u32(w1_hi + c * 16384));
            const uint64_t b_lo = tc::smem_desc_sw128(tc::smem_u32(w1_lo + c * 16384));
            tc::mma_split_ts<4>(acc, a_base + slot * 64, a_base + slot * 64 + 32, b_hi, b_lo, IDESC, c == 0);
            tc::mma_commit(&gs->a_empty[slot]);
            if (c + 1 == Cfg::NCHUNK) tc::mma_commit(&gs->acc_full);
          }
          __syncwarp();
          ++nfill[g][slot];
          ++step[g];
        } else {
          if (!__all_sync(0xffffffffu, tc::mbar_try_wait(&gs->h_full, tiles[g] & 1u))) continue;
          tc::fence_after_sync();
          if (tc::elect_one()) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {  // K = 128 = 2 swizzle atoms of 64
              const uint64_t b_hi = tc::smem_desc_sw128(tc::smem_u32(w2_hi + half * 16384));
              const uint64_t b_lo = tc::smem_desc_sw128(tc::smem_u32(w2_lo + half * 16384));
              tc::mma_split_ts<4>(acc, a_base + half * 32, a_base + 64 + half * 32, b_hi, b_lo, IDESC, half == 0);
            }
            tc::mma_commit(&gs->acc2_full);
          }
          __syncwarp();
          step[g] = 0;
          ++tiles[g];
          ++t_cur[g];
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

template <int K>
static int launch_fv_tc(const FvTcParams& prm, int n_sm, cudaStream_t stream) {
  using Cfg = FvCfg<K>;
  const size_t smem = 1024 + Cfg::W_BYTES + sizeof(float) * (128 * 4 + 2 * B200_MAX_VIEWS * B200_CAM_STRIDE + 24) +
                      2 * sizeof(GroupSync) + 16;
  B200_CHECK_CUDA(cudaFuncSetAttribute(fv_tc_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int N = prm.h * prm.w;
  const long long total_tiles = (long long)prm.B * ((N + FVT_ROWS - 1) / FVT_ROWS) * prm.D;
  int grid = n_sm;
  if ((long long)grid * 2 > total_tiles) grid = (int)((total_tiles + 1) / 2);
  fv_tc_kernel<K><<<grid, FVT_THREADS, smem, stream>>>(prm);
  B200_CHECK_LAUNCH("fv_mlp_tc");
  return 0;
}

extern "C" int b200_fv_mlp_tc(const float* cur, const float* src, const float* cams, const float* cur_invK,
                              const float* planes, const float* bias_eff, const void* wimage, const float* b2,
                              const float* w3, const float* b3, float* vol, unsigned char* mask_out, int B, int K,
                              int C, int h, int w, int D, void* stream) {
  B200_CHECK_ARG(C == B200_FEAT_C, "fv_mlp_tc: only %d feature channels supported (got %d)", B200_FEAT_C, C);
  B200_CHECK_ARG(B > 0 && K > 0 && K <= B200_MAX_VIEWS && D > 0 && h > 0 && w > 0,
                 "fv_mlp_tc: bad sizes B=%d K=%d D=%d h=%d w=%d", B, K, D, h, w);
  B200_CHECK_ARG(cur && src && cams && cur_invK && planes && bias_eff && wimage && b2 && w3 && b3 && vol,
                 "fv_mlp_tc: null pointer");
  B200_CHECK_ARG(((uintptr_t)wimage & 15) == 0, "fv_mlp_tc: weight image must be 16-byte aligned");
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    B200_CHECK_CUDA(cudaGetDevice(&dev));
    B200_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  FvTcParams prm{cur, src, cams, cur_invK, planes, bias_eff, (const uint8_t*)wimage, b2, w3, b3, vol, mask_out,
                 B, D, h, w};
  cudaStream_t st = (cudaStream_t)stream;
  switch (K) {
    case 1: return launch_fv_tc<1>(prm, n_sm, st);
    case 2: return launch_fv_tc<2>(prm, n_sm, st);
    case 3: return launch_fv_tc<3>(prm, n_sm, st);
    case 4: return launch_fv_tc<4>(prm, n_sm, st);
    case 5: return launch_fv_tc<5>(prm, n_sm, st);
    case 6: return launch_fv_tc<6>(prm, n_sm, st);
    case 7: return launch_fv_tc<7>(prm, n_sm, st);
    default: return launch_fv_tc<8>(prm, n_sm, st);
  }
}

// size of the shared-memory weight image expected by b200_fv_mlp_tc for K source views
extern "C" int b200_fv_tc_wimage_bytes(int K) {
  const int kin = FV_VIEW_CH * K + FV_TAIL_CH;
  return (2 * ((kin + 63) / 64) + 4) * 16384;
}
