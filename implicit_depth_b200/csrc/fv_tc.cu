// Feature-volume MLP (mlp_feature_volume) on the 5th-generation tensor cores.
//
// Persistent kernel, one CTA per SM, 17 warps:
//   warps 0-15   two independent "row groups" of 256 threads.  A group works on one 128-row tile at a time
//                (= 128 consecutive pixels at one depth plane); row i of the tile is TMEM lane i and is built by
//                TWO threads -- role A (views 0..ceil(K/2)-1 and half of the current-view channels) and role B
//                (the remaining views, the rest of the tail) -- that sit in warps w and w+4 of the group, i.e. in
//                the same TMEM lane quarter.  Each gathers/warps its source views, builds its share of the row's
//                input channels in registers, splits them into bf16 hi/lo and stores them 32 channels at a time
//                straight into TENSOR MEMORY as the A operand (tcgen05.st) -- the 2.5 GB MLP-input tensor of the
//                reference never exists anywhere.  Two threads per row double the number of row
//                warps per scheduler (the row work is issue/latency bound, the tensor pipe waits for it).
//   warp 16      one elected thread issues every tcgen05.mma: layer 1 as A(TMEM) x W1(smem), 3 bf16 passes
//                (hi*hi, hi*lo, lo*hi) per 16-wide k-step so the result is fp32-grade; layer 2 the same with H1
//                re-stored to TMEM by the row threads (each role owns 64 of the 128 hidden units in both
//                epilogues; role B hands its partial output dot to role A through shared memory).
// Accumulators live in TMEM (128 columns per group); both weight matrices stay resident in shared memory as
// pre-swizzled split-bf16 images (160 KB) for the life of the CTA.  While one group is in an epilogue or
// gathering, the other group's MMAs keep the tensor pipe busy.
//
// TMEM map (512 columns): group g uses [256g, 256g+128) as A ring (2 slots x (32 hi + 32 lo)) and, after
// layer 1, as H1 (64 hi + 64 lo); [256g+128, 256g+256) is the fp32 accumulator.
//
// K-dimension layout of a row (first-layer weight columns are permuted on the host to match,
// implicit_depth_b200/cost_volume.py: tc_channel_layout): 32-channel "halves";
//   role A: view blocks 0..KA-1 (22 ch each, fv_rows.cuh) | cur[0..7]           | zero pad to HA halves
//   role B: view blocks KA..K-1                            | cur[8..15] curray zd | zero pad to HB halves
// with HA + HB even; the halves of the two roles are interleaved in K (A0 B0 A1 B1 ..., FvCfg::half_pos) and two
// consecutive halves form one 64-channel MMA chunk = one ring slot, so chunks complete while the build goes on.
//
// Replaces FeatureVolumeManager.build_cost_volume (modules/cost_volume.py:437-706) /
// FastFeatureVolumeManager.build_cost_volume (:938-1146) + MLP (modules/networks.py:218-233).
#include "fv_rows.cuh"
#include "tc.cuh"

#define FVT_ROWS 128
#define FVT_GROUP_THREADS 256
#define FVT_ROW_THREADS 512
#define FVT_THREADS 544

struct FvTcParams {
  const float* cur;       // [B,4,N,4] quarter-planar (common.cuh)
  const float* src;       // [B,K,4,N,4]
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
  static constexpr int KA = (K + 1) / 2, KB = K - KA;             // views built by role A / B
  static constexpr int CA = FV_VIEW_CH * KA + 8;                    // + cur[0..7]
  static constexpr int CB = FV_VIEW_CH * KB + 12;                   // + cur[8..15], curray[3], z_d
  static constexpr int HA = (CA + 31) / 32;
  static constexpr int NCHUNK = (HA + (CB + 31) / 32 + 1) / 2;
  static constexpr int HB = 2 * NCHUNK - HA;
  static constexpr int W_BYTES = (2 * NCHUNK + 4) * 16384;
  static constexpr int HMIN = HA < HB ? HA : HB;
  // Position of a role's h-th half in the row's K dimension: the two roles' halves are interleaved (A0 B0 A1 B1 ...)
  // so that 64-channel chunks complete progressively while both roles are still gathering -- with role A's halves
  // first and role B's last, two of the three chunks only completed at the very end of the build and their 24 MMAs
  // sat in every tile's critical path.
  __host__ __device__ static constexpr int half_pos(int role, int h) { return h < HMIN ? 2 * h + role : 2 * HMIN + (h - HMIN); }
};

// 16-byte shared-memory load through an explicit shared address (the tables are carved out of the dynamic shared
// memory block through generic pointers, which otherwise compile to generic LD.32, one per value).
__device__ __forceinline__ float4 lds4(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(tc::smem_u32(p)));
  return v;
}

// Flush one 32-channel half (global half index `hg` within the row) into the ring slot of its chunk.
// `g0` = chunks issued by this group before this tile: chunk G = g0 + hg/2 lives in slot G & 1 and is that slot's
// (G >> 1)-th use.
__device__ __forceinline__ void flush_half(const float (&buf)[32], uint32_t a_base, GroupSync* gs, uint32_t hg,
                                           uint32_t g0) {
  const uint32_t G = g0 + (hg >> 1);
  const uint32_t slot = G & 1u, use = G >> 1;
  if (use >= 1) tc::mbar_wait(&gs->a_empty[slot], (use - 1) & 1u);
  tc::fence_after_sync();
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) tc::split2(buf[2 * j], buf[2 * j + 1], hi[j], lo[j]);
  tc::tmem_st16(a_base + slot * 64 + (hg & 1u) * 16, hi);
  tc::tmem_st16(a_base + slot * 64 + 32 + (hg & 1u) * 16, lo);
  tc::wait_st();
  tc::fence_before_sync();
  tc::mbar_arrive(&gs->a_full[slot]);
}

// Role-specific part of one row: gathers this role's views and pushes its channels into tensor memory.
// Returns whether any of the role's views samples inside get_mask's window (cost_volume.py:75-96).
template <int K, int ROLE>
__device__ __forceinline__ bool build_row_share(const FvTcParams& prm, const PixelCtx& pc, const float (&c16)[16],
                                                const float* my_cam, int b, int N, float zd, bool exact_div,
                                                uint32_t a_base, GroupSync* gs, uint32_t g0) {
  using Cfg = FvCfg<K>;
  constexpr int V0 = ROLE ? Cfg::KA : 0, NV = ROLE ? Cfg::KB : Cfg::KA;
  constexpr int NH = ROLE ? Cfg::HB : Cfg::HA;
  constexpr int NT = ROLE ? 12 : 8;
  float buf[32];
  bool inb = false;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    float out[FV_VIEW_CH];
    const int k = V0 + v;
    inb |= fv_view_block_fast(pc, my_cam + k * B200_CAM_STRIDE, prm.src + ((size_t)b * K + k) * N * B200_FEAT_C, c16,
                              zd, prm.h, prm.w, exact_div, out);
#pragma unroll
    for (int c = 0; c < FV_VIEW_CH; ++c) {
      const int ch = v * FV_VIEW_CH + c;
      buf[ch & 31] = out[c];
      if ((ch & 31) == 31) flush_half(buf, a_base, gs, Cfg::half_pos(ROLE, ch >> 5), g0);
    }
  }
#pragma unroll
  for (int c = 0; c < NH * 32 - NV * FV_VIEW_CH; ++c) {
    const int ch = NV * FV_VIEW_CH + c;
    float v = 0.f;  // K padding
    if (c < NT) {
      if (ROLE == 0) v = c16[c];
      else if (c < 8) v = c16[8 + c];
      else if (c < 11) v = pc.curray[c - 8];
      else v = zd;
    }
    buf[ch & 31] = v;
    if ((ch & 31) == 31) flush_half(buf, a_base, gs, Cfg::half_pos(ROLE, ch >> 5), g0);
  }
  return inb;
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
  float2* part_s = reinterpret_cast<float2*>(invk_s + 24);       // [2 groups][2 tile parities][128]: B's partial, inb
  GroupSync* gsync = reinterpret_cast<GroupSync*>(part_s + 2 * 2 * FVT_ROWS);  // 8-byte aligned
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
  if (warp == 16) {
    tc::tmem_alloc(tmem_slot, 512);
    if (lane == 0) {
      for (int g = 0; g < 2; ++g) {
        tc::mbar_init(&gsync[g].a_full[0], 256);  // 2 halves x 128 rows
        tc::mbar_init(&gsync[g].a_full[1], 256);
        tc::mbar_init(&gsync[g].a_empty[0], 1);
        tc::mbar_init(&gsync[g].a_empty[1], 1);
        tc::mbar_init(&gsync[g].acc_full, 1);
        tc::mbar_init(&gsync[g].h_full, 256);
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

  if (warp < 16) {
    // =============================== row groups ===============================
    const int g = warp >> 3;
    const int role = (warp >> 2) & 1;
    const int qd = warp & 3;  // TMEM lane quarter
    const int row = qd * 32 + lane;
    const int gid = blockIdx.x * 2 + g;
    const long long t_begin = total_tiles * gid / n_groups;
    const long long t_end = total_tiles * (gid + 1) / n_groups;
    GroupSync* gs = &gsync[g];
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    const uint32_t a_base = tmem + lane_base + g * 256;
    const uint32_t acc_base = a_base + 128;
    float* my_bias = bias_s + g * 128;
    float* my_cam = cam_s + g * B200_MAX_VIEWS * B200_CAM_STRIDE;
    float* my_invk = invk_s + g * 12;
    float2* my_part = part_s + g * 2 * FVT_ROWS;
    const float b3v = prm.b3[0];
    const int gt = tid & (FVT_GROUP_THREADS - 1);  // thread index within the group

    uint32_t tiles = 0;  // tiles done so far by this group
    int cur_b = -1;
    for (long long t = t_begin; t < t_end; ++t, ++tiles) {
      const int d = (int)(t % prm.D);
      const long long pbq = t / prm.D;
      const int pb = (int)(pbq % NB);
      const int b = (int)(pbq / NB);
      if (b != cur_b) {  // uniform over the group: per-frame tables
        tc::named_sync(9 + g, FVT_GROUP_THREADS);
        if (gt < 128) my_bias[gt] = prm.bias_eff[b * 128 + gt];
        for (int i = gt; i < K * B200_CAM_STRIDE; i += FVT_GROUP_THREADS)
          my_cam[i] = prm.cams[(size_t)b * K * B200_CAM_STRIDE + i];
        if (gt >= 128 && gt < 137) my_invk[gt - 128] = prm.invK[b * 16 + ((gt - 128) / 3) * 4 + (gt - 128) % 3];
        tc::named_sync(9 + g, FVT_GROUP_THREADS);
        cur_b = b;
      }
      const int p_raw = pb * FVT_ROWS + row;
      const int p = min(p_raw, N - 1);
      const int y = p / prm.w, x = p - y * prm.w;
      const float zd = prm.planes[b * prm.D + d];
      const PixelCtx pc = make_pixel_ctx(x, y, my_invk);
      float c16[16];
      {
        const float* cp = prm.cur + (size_t)b * N * B200_FEAT_C + (size_t)p * FEAT_Q;  // quarter-planar
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 t4 = ldg4(cp + (size_t)v * N * FEAT_Q);
          c16[4 * v] = t4.x; c16[4 * v + 1] = t4.y; c16[4 * v + 2] = t4.z; c16[4 * v + 3] = t4.w;
        }
      }
      // ---- build this role's share of the row straight into tensor memory ----
      const bool last_plane = d == prm.D - 1;  // its in-bounds test is overall_mask (cost_volume.py:603-615)
      const uint32_t g0 = tiles * Cfg::NCHUNK;
      bool inb;
      if (role == 0) inb = build_row_share<K, 0>(prm, pc, c16, my_cam, b, N, zd, last_plane, a_base, gs, g0);
      else inb = build_row_share<K, 1>(prm, pc, c16, my_cam, b, N, zd, last_plane, a_base, gs, g0);

      // ---- epilogue 1: H1 = lrelu(acc + bias_eff) -> split -> TMEM (over the A ring); 64 units per role ----
      tc::mbar_wait(&gs->acc_full, tiles & 1u);
      tc::fence_after_sync();
#pragma unroll
      for (int qq = 0; qq < 2; ++qq) {
        const int q = 2 * role + qq;
        uint32_t r[32];
        tc::tmem_ld32(acc_base + 32 * q, r);
        tc::wait_ld();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 bb = lds4(my_bias + 32 * q + 4 * j4);
          const float v0 = leaky(__uint_as_float(r[4 * j4]) + bb.x, 0.01f);
          const float v1 = leaky(__uint_as_float(r[4 * j4 + 1]) + bb.y, 0.01f);
          const float v2 = leaky(__uint_as_float(r[4 * j4 + 2]) + bb.z, 0.01f);
          const float v3 = leaky(__uint_as_float(r[4 * j4 + 3]) + bb.w, 0.01f);
          tc::split2(v0, v1, hi[2 * j4], lo[2 * j4]);
          tc::split2(v2, v3, hi[2 * j4 + 1], lo[2 * j4 + 1]);
        }
        tc::tmem_st16(a_base + 16 * q, hi);
        tc::tmem_st16(a_base + 64 + 16 * q, lo);
      }
      tc::wait_st();
      tc::fence_before_sync();
      tc::mbar_arrive(&gs->h_full);

      // ---- epilogue 2: out = lrelu(acc2 + b2) . w3 + b3; role B's 64-unit partial travels through smem ----
      tc::mbar_wait(&gs->acc2_full, tiles & 1u);
      tc::fence_after_sync();
      float o = 0.f;
#pragma unroll
      for (int qq = 0; qq < 2; ++qq) {
        const int q = 2 * role + qq;
        uint32_t r[32];
        tc::tmem_ld32(acc_base + 32 * q, r);
        tc::wait_ld();
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 bb = lds4(b2_s + 32 * q + 4 * j4), ww = lds4(w3_s + 32 * q + 4 * j4);
          o = fmaf(leaky(__uint_as_float(r[4 * j4]) + bb.x, 0.01f), ww.x, o);
          o = fmaf(leaky(__uint_as_float(r[4 * j4 + 1]) + bb.y, 0.01f), ww.y, o);
          o = fmaf(leaky(__uint_as_float(r[4 * j4 + 2]) + bb.z, 0.01f), ww.z, o);
          o = fmaf(leaky(__uint_as_float(r[4 * j4 + 3]) + bb.w, 0.01f), ww.w, o);
        }
      }
      float2* slot = my_part + (tiles & 1u) * FVT_ROWS + row;
      if (role == 1) *slot = make_float2(o, inb ? 1.f : 0.f);
      // the two warps of this lane quarter meet: B's partial is visible to A, and A may go on to the next tile
      // (whose first MMA overwrites the accumulator) only after B has finished reading acc2
      tc::fence_before_sync();
      tc::named_sync(1 + g * 4 + qd, 64);
      tc::fence_after_sync();
      if (role == 0 && p_raw < N) {
        const float2 other = *slot;
        prm.vol[((size_t)b * prm.D + d) * N + p] = o + other.x + b3v;
        if (prm.mask != nullptr && last_plane) prm.mask[(size_t)b * N + p] = (inb || other.y != 0.f) ? 1 : 0;
      }
      tc::fence_before_sync();  // order this tile's TMEM reads before the next tile's MMAs (via a_full)
    }
  } else {
    // =============================== MMA issuer (whole warp loops, one elected lane issues) ==========
    constexpr uint32_t IDESC = tc::idesc_bf16_f32(128, 128);
    long long t_cur[2], t_end[2];
    uint32_t tiles[2] = {0, 0};
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
          const int c = step[g];
          const uint32_t G = tiles[g] * Cfg::NCHUNK + c;
          const uint32_t slot = G & 1u, use = G >> 1;
          if (!__all_sync(0xffffffffu, tc::mbar_try_wait(&gs->a_full[slot], use & 1u))) continue;
          tc::fence_after_sync();
          if (tc::elect_one()) {
            const uint64_t b_hi = tc::smem_desc_sw128(tc::smem_u32(w1_hi + c * 16384));
            const uint64_t b_lo = tc::smem_desc_sw128(tc::smem_u32(w1_lo + c * 16384));
            tc::mma_split_ts<4>(acc, a_base + slot * 64, a_base + slot * 64 + 32, b_hi, b_lo, IDESC, c == 0);
            tc::mma_commit(&gs->a_empty[slot]);
            if (c + 1 == Cfg::NCHUNK) tc::mma_commit(&gs->acc_full);
          }
          __syncwarp();
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
  if (warp == 16) tc::tmem_dealloc(tmem, 512);
}

template <int K>
static int launch_fv_tc(const FvTcParams& prm, int n_sm, int max_ctas, cudaStream_t stream) {
  using Cfg = FvCfg<K>;
  const size_t smem = 1024 + Cfg::W_BYTES + sizeof(float) * (128 * 4 + 2 * B200_MAX_VIEWS * B200_CAM_STRIDE + 24) +
                      sizeof(float2) * 2 * 2 * FVT_ROWS + 2 * sizeof(GroupSync) + 16;
  B200_CHECK_CUDA(cudaFuncSetAttribute(fv_tc_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int N = prm.h * prm.w;
  const long long total_tiles = (long long)prm.B * ((N + FVT_ROWS - 1) / FVT_ROWS) * prm.D;
  int grid = n_sm;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  if ((long long)grid * 2 > total_tiles) grid = (int)((total_tiles + 1) / 2);
  fv_tc_kernel<K><<<grid, FVT_THREADS, smem, stream>>>(prm);
  B200_CHECK_LAUNCH("fv_mlp_tc");
  return 0;
}

extern "C" int b200_fv_mlp_tc(const float* cur, const float* src, const float* cams, const float* cur_invK,
                              const float* planes, const float* bias_eff, const void* wimage, const float* b2,
                              const float* w3, const float* b3, float* vol, unsigned char* mask_out, int B, int K,
                              int C, int h, int w, int D, int max_ctas, void* stream) {
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
    case 1: return launch_fv_tc<1>(prm, n_sm, max_ctas, st);
    case 2: return launch_fv_tc<2>(prm, n_sm, max_ctas, st);
    case 3: return launch_fv_tc<3>(prm, n_sm, max_ctas, st);
    case 4: return launch_fv_tc<4>(prm, n_sm, max_ctas, st);
    case 5: return launch_fv_tc<5>(prm, n_sm, max_ctas, st);
    case 6: return launch_fv_tc<6>(prm, n_sm, max_ctas, st);
    case 7: return launch_fv_tc<7>(prm, n_sm, max_ctas, st);
    default: return launch_fv_tc<8>(prm, n_sm, max_ctas, st);
  }
}

// K-dimension layout of the first-layer weight image expected by b200_fv_mlp_tc for K source views:
// out[0] = views built by role A, out[1] = 32-channel halves of role A, out[2] = halves of role B,
// out[3] = 64-channel chunks.  Returns the image size in bytes.
extern "C" int b200_fv_tc_layout(int K, int* out) {
  const int KA = (K + 1) / 2, KB = K - KA;
  const int HA = (FV_VIEW_CH * KA + 8 + 31) / 32;
  const int NCHUNK = (HA + (FV_VIEW_CH * KB + 12 + 31) / 32 + 1) / 2;
  if (out) {
    out[0] = KA;
    out[1] = HA;
    out[2] = 2 * NCHUNK - HA;
    out[3] = NCHUNK;
  }
  return (2 * NCHUNK + 4) * 16384;
}

// size of the shared-memory weight image expected by b200_fv_mlp_tc for K source views
extern "C" int b200_fv_tc_wimage_bytes(int K) { return b200_fv_tc_layout(K, nullptr); }
