// Implicit-GEMM 2-D convolution on the tcgen05 tensor cores, NHWC split-bf16 activations.
//
//   out[b,oy,ox,:] = act( sum_seg sum_tap W_seg,tap @ in_seg[b, s*oy+dy-p, s*ox+dx-p, :] + bias (+ residual) )
//
// * A conv is a list of up to 4 K-SEGMENTS (input tensor, 1x1/3x3, stride 1/2).  Channel concatenation
//   (torch.cat before a conv: CVEncoder networks.py:208-215, BDDecoderPP :71-77, SkipDecoder
//   networks_fast.py:42-43) and the BasicBlock shortcut conv (layers.py:64-71, 89-92) are just extra
//   segments accumulating into the same TMEM tile -- nothing is concatenated or added in memory.
// * Activations are stored as two bf16 planes (hi, lo) with x = hi + lo (16 mantissa bits, same bytes as
//   fp32); weights are pre-split the same way.  Each K-chunk issues hi*hi, hi*lo, lo*hi MMAs into fp32
//   accumulators: fp32-grade results (the CPU-oracle parity bar) at bf16 MMA rates.
// * Two kernels behind one plan API:
//     conv_halo_kernel (conv_halo.cuh)  stride-1 convs with Cout % 64 == 0 -- 141 of the 147 conv launches of a
//                                        forward: halo patches, merged-N MMAs, TMA-store epilogue;
//     conv_tc_kernel (below)            everything else (stride 2, Cout = 16): one TMA box per tap, 64-channel
//                                        chunks, 128-byte swizzle, direct-store epilogue.
//   Both are warp-specialised and persistent: warp 0 = TMA producer, warp 1 = MMA issuer, remaining warps =
//   epilogue; two TMEM accumulator sets so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc.cuh"
#include "tmap.cuh"

#define CV_THREADS 320  // warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue (two per TMEM lane quarter)
#define CV_TH 8
#define CV_TW 16
#define CV_MAX_SEG 4

enum { ACT_NONE = 0, ACT_LRELU = 1, ACT_ELU = 2, ACT_RELU = 3, ACT_SILU = 4 };

struct ConvKParams {
  CUtensorMap maps[2 * CV_MAX_SEG];  // [2s] = hi plane, [2s+1] = lo plane of segment s
  CUtensorMap out_maps[2];           // halo kernel: output hi / lo (TMA stores)
  CUtensorMap res_maps[2];           // halo kernel: residual hi / lo (TMA loads into the staging tiles)
  int seg_C[CV_MAX_SEG], seg_ksize[CV_MAX_SEG], seg_stride[CV_MAX_SEG], seg_pad[CV_MAX_SEG];
  int nseg;
  const uint8_t* wimage;  // packed weights, layout depends on the kernel (see b200_conv_uses_halo)
  const float* bias;      // [Cout] or null
  const __nv_bfloat16* res_hi;  // optional residual, NHWC [B,OH,OW,Cout]
  const __nv_bfloat16* res_lo;
  __nv_bfloat16* out_hi;  // NHWC [B,OH,OW,Cout] (may be null if only out_f32 is wanted)
  __nv_bfloat16* out_lo;
  float* out_f32;         // optional NHWC fp32 copy (plain kernel only)
  int B, OH, OW, Cout, NT, n_ntiles, act, total_chunks, tiles_x, tiles_y, stages;
  float slope;
  int halo, sub, has_res, tps;
  int two;                     // halo kernel: two CTAs per SM (SUB = 1, NT = 64)
  int staged;                  // plain kernel: coalesced stores through a per-warp shared-memory transpose (NT 64 / 128)
  int ksplit, total_patches;   // halo kernel, split-K over a cluster of `ksplit` CTAs (1 = off); 32-channel patches per item
  int seg_chunk0[CV_MAX_SEG];  // index of each segment's first chunk in the weight image
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == ACT_LRELU) return fmaxf(v, v * slope);  // slope in [0, 1], checked at plan creation
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_ELU) return elu1(v);
  if (act == ACT_SILU) return silu(v);
  return v;
}

#include <type_traits>

// activation with the kind fixed at compile time (apply_act tests it per element)
template <int ACT>
__device__ __forceinline__ float act_t(float v, float slope) {
  if (ACT == ACT_LRELU) return fmaxf(v, v * slope);
  if (ACT == ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == ACT_ELU) return elu1(v);
  if (ACT == ACT_SILU) return silu(v);
  return v;
}
// calls f(integral_constant<int, act>, bool_constant<res>): one uniform branch per call instead of one per element
template <class F>
__device__ __forceinline__ void dispatch_act_res(int act, bool res, F&& f) {
#define B200_ACT_CASE(A)                                                          \
  case A:                                                                         \
    if (res) f(std::integral_constant<int, A>{}, std::true_type{});               \
    else f(std::integral_constant<int, A>{}, std::false_type{});                  \
    break;
  switch (act) {
    B200_ACT_CASE(ACT_LRELU)
    B200_ACT_CASE(ACT_RELU)
    B200_ACT_CASE(ACT_ELU)
    B200_ACT_CASE(ACT_SILU)
    default:
      if (res) f(std::integral_constant<int, ACT_NONE>{}, std::true_type{});
      else f(std::integral_constant<int, ACT_NONE>{}, std::false_type{});
  }
#undef B200_ACT_CASE
}

__global__ void __launch_bounds__(CV_THREADS, 1) conv_tc_kernel(const __grid_constant__ ConvKParams prm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NT = prm.NT;
  const uint32_t b_bytes = (uint32_t)NT * 256u;               // hi + lo weight tiles of one chunk
  const uint32_t stage_bytes = 32768u + ((b_bytes + 1023u) & ~1023u);
  const int S = prm.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)S * stage_bytes);
  uint64_t* empty = full + S;
  uint64_t* acc_full = empty + S;    // [2]
  uint64_t* acc_empty = acc_full + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint4* stage_out = reinterpret_cast<uint4*>(
      (reinterpret_cast<uintptr_t>(tmem_slot + 4) + 127) & ~uintptr_t(127));  // [8 warps][256] uint4 (staged stores)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m_tiles = prm.B * prm.tiles_y * prm.tiles_x;
  const int items = m_tiles * prm.n_ntiles;
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * NT) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 2 * prm.nseg; ++s) tc::prefetch_tmap(&prm.maps[s]);
    for (int s = 0; s < S; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(&acc_full[a], 1);
      tc::mbar_init(&acc_empty[a], 256);
    }
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, tmem_cols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  tc::grid_dependency_wait();  // PDL: inputs of the previous kernel are complete and visible from here on

  if (warp == 0) {
    {
      // =========================== TMA producer (whole warp loops, one elected lane issues) ===========
      const bool leader = tc::elect_one();
      uint32_t stage = 0, round = 0;  // stage ring position / wrap count
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int nt = item % prm.n_ntiles;
        const int mt = item / prm.n_ntiles;
        const int tx = mt % prm.tiles_x;
        const int ty = (mt / prm.tiles_x) % prm.tiles_y;
        const int b = mt / (prm.tiles_x * prm.tiles_y);
        const uint8_t* wsrc = prm.wimage + (size_t)nt * prm.total_chunks * b_bytes;
        for (int s = 0; s < prm.nseg; ++s) {
          const int ks = prm.seg_ksize[s], st = prm.seg_stride[s], pd = prm.seg_pad[s];
          const int cblocks = (prm.seg_C[s] + 63) >> 6;
          for (int tap = 0; tap < ks * ks; ++tap) {
            const int dy = tap / ks, dx = tap % ks;
            const int x0 = (tx * CV_TW) * st + dx - pd;
            const int y0 = (ty * CV_TH) * st + dy - pd;
            for (int cb = 0; cb < cblocks; ++cb) {
              if (round > 0) tc::mbar_wait(&empty[stage], (round - 1) & 1u);
              uint8_t* sa = base + (size_t)stage * stage_bytes;
              if (leader) {
                tc::mbar_expect_tx(&full[stage], 32768u + b_bytes);
                tc::tma_load_4d(sa, &prm.maps[2 * s], cb * 64, x0, y0, b, &full[stage]);
                tc::tma_load_4d(sa + 16384, &prm.maps[2 * s + 1], cb * 64, x0, y0, b, &full[stage]);
                tc::bulk_load(sa + 32768, wsrc, b_bytes, &full[stage]);
              }
              wsrc += b_bytes;
              if (++stage == (uint32_t)S) { stage = 0; ++round; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // =========================== MMA issuer (whole warp loops, one elected lane issues) ============
      const uint32_t idesc = tc::idesc_bf16_f32(128, NT);
      // (ring position kept incrementally, descriptors by addition to a pre-built base, one issuing lane for the whole
      // loop and no per-round warp sync: the tensor pipe queues only a few MMAs, so every cycle between the last MMA
      // of a round and the first of the next is idle pipe time -- see the halo kernel)
      const bool leader = tc::elect_one();
      const uint64_t a_base = tc::smem_desc_sw128(tc::smem_u32(base));
      const uint32_t stage_units = stage_bytes >> 4;
      uint32_t stage = 0, phase = 0, tile_i = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++tile_i) {
        const uint32_t a = tile_i & 1u;
        const uint32_t use = tile_i >> 1;  // how often accumulator a was used before
        if (use > 0) tc::mbar_wait(&acc_empty[a], (use - 1) & 1u);
        tc::fence_after_sync();
        const uint32_t acc = tmem + a * NT;
        uint32_t first = 1;
        for (int s = 0; s < prm.nseg; ++s) {
          const int taps = prm.seg_ksize[s] * prm.seg_ksize[s];
          const int C = prm.seg_C[s];
          const int cblocks = (C + 63) >> 6;
          for (int tap = 0; tap < taps; ++tap) {
            for (int cb = 0; cb < cblocks; ++cb) {
              tc::mbar_wait(&full[stage], phase);
              tc::fence_after_sync();
              const int ksteps = (min(64, C - cb * 64) + 15) >> 4;
              if (leader) {
                const uint64_t a_hi = a_base + (uint64_t)(stage * stage_units), b_hi = a_hi + (32768u >> 4);
                tc::mma_split_ss_n(ksteps, acc, a_hi, a_hi + (16384u >> 4), b_hi, b_hi + (((uint32_t)NT * 128u) >> 4),
                                   idesc, first);
                tc::mma_commit(&empty[stage]);
              }
              first = 0;
              if (++stage == (uint32_t)S) { stage = 0; phase ^= 1u; }
            }
          }
        }
        if (leader) tc::mma_commit(&acc_full[a]);
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue (warps 2..9) ===========================
    // a pixel (= TMEM lane) is drained by two threads: warps w and w + 4 read the same lane quarter and take the even /
    // the odd 16-channel chunks of the tile (1x1 and strided layers are drain-bound: few MMAs per output value)
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    uint32_t tile_i = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++tile_i) {
      const int nt = item % prm.n_ntiles;
      const int mt = item / prm.n_ntiles;
      const int tx = mt % prm.tiles_x;
      const int ty = (mt / prm.tiles_x) % prm.tiles_y;
      const int b = mt / (prm.tiles_x * prm.tiles_y);
      const uint32_t a = tile_i & 1u;
      tc::mbar_wait(&acc_full[a], (tile_i >> 1) & 1u);
      tc::fence_after_sync();
      const int oy = ty * CV_TH + (row >> 4), ox = tx * CV_TW + (row & 15);
      const bool live = (oy < prm.OH) && (ox < prm.OW);
      const size_t pix = ((size_t)b * prm.OH + oy) * prm.OW + ox;
      const int n_base = nt * NT;
      // activation / residual as compile-time cases of one body (the drain is instruction-bound)
      auto drain = [&](auto act_c, auto res_c) {
        constexpr int ACT = decltype(act_c)::value;
        constexpr bool RES = decltype(res_c)::value;
        for (int n0 = 16 * half; n0 < NT; n0 += 32) {
          const int n = n_base + n0;
          float4 bv[4];  // bias of the chunk: four 16-byte loads issued before the TMEM read is waited for
#pragma unroll
          for (int q = 0; q < 4; ++q)
            bv[q] = prm.bias ? __ldg(reinterpret_cast<const float4*>(prm.bias + n) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          uint32_t r[16];
          tc::tmem_ld16(tmem + lane_base + a * NT + n0, r);
          tc::wait_ld();
          if (live) {
            float v[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              v[4 * q] = __uint_as_float(r[4 * q]) + bv[q].x;
              v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bv[q].y;
              v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bv[q].z;
              v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bv[q].w;
            }
            if (RES) {
              const uint4* rh = reinterpret_cast<const uint4*>(prm.res_hi + pix * prm.Cout + n);
              const uint4* rl = reinterpret_cast<const uint4*>(prm.res_lo + pix * prm.Cout + n);
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const uint4 h4 = __ldg(rh + q), l4 = __ldg(rl + q);
                const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  v[8 * q + 2 * e] += __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
                  v[8 * q + 2 * e + 1] += __uint_as_float(hw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = act_t<ACT>(v[j], prm.slope);
            if (prm.out_hi) {
              uint32_t hi[8], lo[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) tc::split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
              uint4* oh = reinterpret_cast<uint4*>(prm.out_hi + pix * prm.Cout + n);
              uint4* ol = reinterpret_cast<uint4*>(prm.out_lo + pix * prm.Cout + n);
              oh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              oh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
              ol[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              ol[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            }
            if (prm.out_f32) {
              float4* of = reinterpret_cast<float4*>(prm.out_f32 + pix * prm.Cout + n);
#pragma unroll
              for (int q = 0; q < 4; ++q) of[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
          }
        }
      };
      // Staged drain (NT = 64 / 128, split-bf16 output only): the thread owns NT/2 CONTIGUOUS channels of its pixel and
      // drains them in rounds of 32.  A direct store puts every lane's 16 bytes into a different 128-byte line (pixel
      // stride = Cout * 2 bytes): 32 L1 tags per instruction, which bounded the 1x1 layers.  Here the warp transposes
      // its 32 pixels x 4 16-byte pieces per plane through shared memory (XOR-swizzled, conflict-free) so that 4 lanes
      // write one pixel's contiguous 64 bytes: 8 lines per store instruction.
      auto drain_staged = [&](auto act_c, auto res_c) {
        constexpr int ACT = decltype(act_c)::value;
        constexpr bool RES = decltype(res_c)::value;
        const int nch = NT >> 1;                       // channels per thread: 32 or 64
        uint4* st_hi = stage_out + (warp - 2) * 256;   // [32 pixels][4 pieces]
        uint4* st_lo = st_hi + 128;
        const int sw = (lane >> 1) & 3;
        for (int rd = 0; rd < (nch >> 5); ++rd) {
          const int c_base = n_base + half * nch + 32 * rd;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int n = c_base + 16 * c;
            float4 bv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
              bv[q] = prm.bias ? __ldg(reinterpret_cast<const float4*>(prm.bias + n) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t r[16];
            tc::tmem_ld16(tmem + lane_base + a * NT + half * nch + 32 * rd + 16 * c, r);
            tc::wait_ld();
            float v[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              v[4 * q] = __uint_as_float(r[4 * q]) + bv[q].x;
              v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bv[q].y;
              v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bv[q].z;
              v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bv[q].w;
            }
            if (RES && live) {
              const uint4* rh = reinterpret_cast<const uint4*>(prm.res_hi + pix * prm.Cout + n);
              const uint4* rl = reinterpret_cast<const uint4*>(prm.res_lo + pix * prm.Cout + n);
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const uint4 h4 = __ldg(rh + q), l4 = __ldg(rl + q);
                const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  v[8 * q + 2 * e] += __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
                  v[8 * q + 2 * e + 1] += __uint_as_float(hw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
                }
              }
            }
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              tc::split2(act_t<ACT>(v[2 * j], prm.slope), act_t<ACT>(v[2 * j + 1], prm.slope), hi[j], lo[j]);
            st_hi[lane * 4 + ((2 * c) ^ sw)] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            st_hi[lane * 4 + ((2 * c + 1) ^ sw)] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            st_lo[lane * 4 + ((2 * c) ^ sw)] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            st_lo[lane * 4 + ((2 * c + 1) ^ sw)] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          }
          __syncwarp();
          const int u = lane & 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int pp = j * 8 + (lane >> 2);
            const int prow = quarter * 32 + pp;
            const int poy = ty * CV_TH + (prow >> 4), pox = tx * CV_TW + (prow & 15);
            const int slot = pp * 4 + (u ^ ((pp >> 1) & 3));
            const uint4 wh = st_hi[slot], wl = st_lo[slot];
            if (poy < prm.OH && pox < prm.OW) {
              const size_t o = (((size_t)b * prm.OH + poy) * prm.OW + pox) * prm.Cout + c_base + 8 * u;
              *reinterpret_cast<uint4*>(prm.out_hi + o) = wh;
              *reinterpret_cast<uint4*>(prm.out_lo + o) = wl;
            }
          }
          __syncwarp();
        }
      };
      if (prm.staged) dispatch_act_res(prm.act, prm.res_hi != nullptr, drain_staged);
      else dispatch_act_res(prm.act, prm.res_hi != nullptr, drain);
      tc::fence_before_sync();
      tc::mbar_arrive(&acc_empty[a]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, tmem_cols);
}



#include "conv_halo.cuh"

// ------------------------------------------------------------------------------------------------
// host side: plan objects (tensor maps are encoded once, launches are cheap and graph-capturable)
// ------------------------------------------------------------------------------------------------
struct ConvPlan {
  ConvKParams k;
  size_t smem;
  int grid;
};

// C-ABI description of one segment / one conv (plain pointers and sizes)
struct b200_conv_seg {
  const void* in_hi;  // NHWC bf16 [B,H,W,C]
  const void* in_lo;
  int H, W, C, ksize, stride, pad;
  int pad_hi;  // padding at the bottom / right edge; `pad` is the top / left one (they differ for TF "SAME" padding
               // of stride-2 convs on even sizes: (0, 1))
};
struct b200_conv_desc {
  b200_conv_seg seg[CV_MAX_SEG];
  int nseg;
  const void* wimage;
  const float* bias;
  const void* res_hi;
  const void* res_lo;
  void* out_hi;
  void* out_lo;
  float* out_f32;
  int B, OH, OW, Cout, act;
  float slope;
  int max_ctas;  // upper bound on the CTAs of this (persistent) launch, 0 = all SMs: lets the caller keep SMs free
                 // for work on another stream
  int tile_hint; // halo kernel, 64-wide N tile: 0 = automatic, 1 = M = 128 items with two CTAs per SM, 2 = M = 256 items
                 // with one CTA per SM (tuning / tests; results do not depend on it)
};

extern "C" int b200_conv_ntile(int Cout) { return Cout <= 128 ? Cout : 128; }

extern "C" int b200_conv_uses_halo(const b200_conv_desc* d);

// N tile of a particular conv (decides the weight-image layout together with b200_conv_uses_halo).  The plain ring
// kernel on a low-resolution map has only a handful of 128-pixel M tiles, each streaming ALL of its weights: with
// fewer items than half the SMs a 64-wide N tile doubles the CTAs and halves the weight stream per CTA (the
// 1536 -> 256 project convs of the image encoder at 12x16: 16 CTAs x 24 chunks -> 32 CTAs).
extern "C" int b200_conv_ntile_for(const b200_conv_desc* d) {
  const int nt = b200_conv_ntile(d->Cout);
  if (nt < 128 || b200_conv_uses_halo(d)) return nt;
  const int m_tiles = d->B * ((d->OW + CV_TW - 1) / CV_TW) * ((d->OH + CV_TH - 1) / CV_TH);
  return (m_tiles * (d->Cout / 128) <= 74) ? 64 : 128;
}

// 1 if this conv runs on the halo kernel (weight image in 32-channel chunks, 64-byte swizzle, [hi | lo] per
// chunk), 0 for the plain kernel (64-channel chunks, 128-byte swizzle).  Pure function of the geometry.
extern "C" int b200_conv_uses_halo(const b200_conv_desc* d) {
  if (!d) return 0;
  const bool n16 = d->Cout == 16 || d->Cout == 32;  // narrow N tile: 3x3 stride-1 segments only, no residual
  if ((d->Cout % 64 != 0 && !n16) || d->out_f32 != nullptr || d->out_hi == nullptr) return 0;
  if (n16 && d->res_hi != nullptr) return 0;
  // a pure 1x1 conv uses every halo patch for a single tap, so the halo kernel's 2-deep patch ring exposes the TMA
  // latency of each 32-channel block (51 us for the 1536->256 project conv of the image encoder); the plain kernel
  // streams 64-channel boxes through a 3..6-stage ring instead
  bool all_1x1 = true;
  for (int s = 0; s < d->nseg; ++s) all_1x1 = all_1x1 && d->seg[s].ksize == 1;
  if (all_1x1) return 0;
  for (int s = 0; s < d->nseg; ++s) {
    const b200_conv_seg& sg = d->seg[s];
    const bool ok = sg.stride == 1 && sg.pad_hi == sg.pad &&
                    ((sg.ksize == 3 && (sg.pad == 0 || sg.pad == 1)) || (sg.ksize == 1 && sg.pad == 0));
    if (!ok || (n16 && sg.ksize != 3)) return 0;
  }
  return 1;
}

extern "C" long long b200_conv_wimage_bytes(const int* seg_C, const int* seg_ksize, int nseg, int Cout, int halo) {
  long long chunks = 0;
  const int cw = halo ? 32 : 64;
  for (int s = 0; s < nseg; ++s) chunks += (long long)seg_ksize[s] * seg_ksize[s] * ((seg_C[s] + cw - 1) / cw);
  const int NT = b200_conv_ntile(Cout);
  return chunks * ((Cout + NT - 1) / NT) * NT * (halo ? 128 : 256);
}

static int encode_nhwc(EncodeTiledFn enc, CUtensorMap* map, const void* ptr, int C, int W, int H, int B,
                       const cuuint32_t box[4], int estride, CUtensorMapSwizzle swz) {
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  return (int)enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

extern "C" int b200_conv_create(const b200_conv_desc* d, void** plan_out) {
  B200_CHECK_ARG(d && plan_out, "conv_create: null pointer");
  B200_CHECK_ARG(d->nseg >= 1 && d->nseg <= CV_MAX_SEG, "conv_create: 1..%d segments (got %d)", CV_MAX_SEG, d->nseg);
  B200_CHECK_ARG(d->Cout % 16 == 0 && d->Cout >= 16, "conv_create: Cout must be a multiple of 16 (got %d)", d->Cout);
  B200_CHECK_ARG(d->Cout <= 128 || d->Cout % 128 == 0, "conv_create: Cout > 128 must be a multiple of 128 (got %d)",
                 d->Cout);
  B200_CHECK_ARG(d->wimage && (d->out_hi || d->out_f32), "conv_create: missing weights or output");
  B200_CHECK_ARG((d->out_hi == nullptr) == (d->out_lo == nullptr), "conv_create: out_hi/out_lo must come together");
  B200_CHECK_ARG((d->res_hi == nullptr) == (d->res_lo == nullptr), "conv_create: res_hi/res_lo must come together");
  B200_CHECK_ARG(((uintptr_t)d->bias & 15) == 0, "conv_create: bias must be 16-byte aligned");
  B200_CHECK_ARG(d->act != ACT_LRELU || (d->slope >= 0.f && d->slope <= 1.f),
                 "conv_create: leaky-ReLU slope must be in [0, 1] (got %g)", (double)d->slope);
  EncodeTiledFn enc = get_encode();
  B200_CHECK_ARG(enc != nullptr, "conv_create: cuTensorMapEncodeTiled not available from the driver");
  ConvPlan* p = new ConvPlan();
  ConvKParams& k = p->k;
  memset(&k, 0, sizeof(k));
  k.nseg = d->nseg;
  k.total_chunks = 0;
  int dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  k.NT = b200_conv_ntile_for(d);
  k.n_ntiles = (d->Cout + k.NT - 1) / k.NT;
  const bool halo = b200_conv_uses_halo(d) != 0;
  k.halo = halo ? 1 : 0;
  k.sub = 1;
  if (halo) {
    // two M=128 sub-tiles per item share every weight chunk when the N tile is 64 wide and there is enough work
    k.sub = (k.NT <= 64) ? 2 : 1;
    if (k.sub == 2) {
      // wave quantisation on the persistent grid: an item of two sub-tiles costs ~1.7x an item of one (shared weight
      // chunks), so e.g. 192 double items on 148 SMs (2 waves = 3.4 units) lose to 384 single items (3 waves)
      const int rows = (d->OH + CVH_ROWS - 1) / CVH_ROWS;
      const int items2 = d->B * ((d->OW + 15) / 16) * rows * k.n_ntiles;
      const int items1 = d->B * ((d->OW + 7) / 8) * rows * k.n_ntiles;
      const int waves2 = (items2 + n_sm - 1) / n_sm, waves1 = (items1 + n_sm - 1) / n_sm;
      if (items2 < n_sm || 10 * waves1 < 17 * waves2) k.sub = 1;
      if (d->tile_hint == 1) k.sub = 1;
      if (d->tile_hint == 2) k.sub = 2;
    }
  }
  // Split-K over a thread-block cluster (conv_halo.cuh): for layers whose item count leaves most of the machine idle
  // (12x16 ... 24x32 maps: 16-64 items, each streaming all of its weights through one SM), `ksplit` CTAs share an item.
  // The split factor depends on the layer's geometry only (items PER IMAGE), never on the batch size: a frame's result
  // must not depend on what else is in the batch (the sum order changes with the factor).  Sized so that a batch of
  // four frames fills the machine (4 x items per image x S <= SMs); larger batches run their clusters in waves.
  k.ksplit = 1;
  k.total_patches = 0;
  if (halo && k.NT >= 64) {
    for (int s = 0; s < d->nseg; ++s) k.total_patches += (d->seg[s].C + 31) / 32;
    const int per_image = ((d->OW + 7) / 8) * ((d->OH + CVH_ROWS - 1) / CVH_ROWS) * k.n_ntiles;  // M = 128 items
    for (int S = 8; S >= 2; S >>= 1)
      if (4 * per_image * S <= n_sm && 2 * S <= k.total_patches && k.NT / S >= 8) {
        k.ksplit = S;
        k.sub = 1;
        break;
      }
  }
  const int cw = halo ? 32 : 64;  // channels per K chunk
  for (int s = 0; s < d->nseg; ++s) {
    const b200_conv_seg& sg = d->seg[s];
    if (!(sg.in_hi && sg.in_lo) || sg.C % 8 != 0 || sg.C < 8 || !(sg.ksize == 1 || sg.ksize == 3) ||
        !(sg.stride == 1 || sg.stride == 2) || ((uintptr_t)sg.in_hi & 15) || ((uintptr_t)sg.in_lo & 15)) {
      delete p;
      b200_set_error("conv_create: segment %d invalid (C=%d must be a multiple of 8, ksize 1|3, stride 1|2, "
                     "16-byte aligned planes)", s, sg.C);
      return -1;
    }
    // output size implied by this segment must match
    const int oh = (sg.H + sg.pad + sg.pad_hi - sg.ksize) / sg.stride + 1;
    const int ow = (sg.W + sg.pad + sg.pad_hi - sg.ksize) / sg.stride + 1;
    if (oh != d->OH || ow != d->OW) {
      delete p;
      b200_set_error("conv_create: segment %d gives %dx%d outputs, conv says %dx%d", s, oh, ow, d->OH, d->OW);
      return -1;
    }
    k.seg_C[s] = sg.C;
    k.seg_ksize[s] = sg.ksize;
    k.seg_stride[s] = sg.stride;
    k.seg_pad[s] = sg.pad;
    k.seg_chunk0[s] = k.total_chunks;
    k.total_chunks += sg.ksize * sg.ksize * ((sg.C + cw - 1) / cw);
    cuuint32_t box[4] = {64, (cuuint32_t)(CV_TW * sg.stride), (cuuint32_t)(CV_TH * sg.stride), 1};
    if (halo) {
      box[0] = 32;
      box[1] = (cuuint32_t)(8 * k.sub + 2);
      box[2] = 18;
    }
    for (int part = 0; part < 2; ++part) {
      const int r = encode_nhwc(enc, &k.maps[2 * s + part], part ? sg.in_lo : sg.in_hi, sg.C, sg.W, sg.H, d->B, box,
                                sg.stride, halo ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
      if (r != 0) {
        delete p;
        b200_set_error("conv_create: cuTensorMapEncodeTiled failed (%d) for segment %d (C=%d H=%d W=%d stride=%d)", r, s,
                       sg.C, sg.H, sg.W, sg.stride);
        return -2;
      }
    }
  }
  if (halo && k.NT >= 64) {
    const cuuint32_t obox[4] = {64, 8, CVH_ROWS, 1};
    for (int part = 0; part < 2; ++part) {
      int r = encode_nhwc(enc, &k.out_maps[part], part ? d->out_lo : d->out_hi, d->Cout, d->OW, d->OH, d->B, obox, 1,
                          CU_TENSOR_MAP_SWIZZLE_128B);
      if (r == 0 && d->res_hi)
        r = encode_nhwc(enc, &k.res_maps[part], part ? d->res_lo : d->res_hi, d->Cout, d->OW, d->OH, d->B, obox, 1,
                        CU_TENSOR_MAP_SWIZZLE_128B);
      if (r != 0) {
        delete p;
        b200_set_error("conv_create: cuTensorMapEncodeTiled failed (%d) for the output / residual map", r);
        return -2;
      }
    }
  }
  k.wimage = (const uint8_t*)d->wimage;
  k.bias = d->bias;
  k.res_hi = (const __nv_bfloat16*)d->res_hi;
  k.res_lo = (const __nv_bfloat16*)d->res_lo;
  k.has_res = d->res_hi != nullptr;
  k.out_hi = (__nv_bfloat16*)d->out_hi;
  k.out_lo = (__nv_bfloat16*)d->out_lo;
  k.out_f32 = d->out_f32;
  k.B = d->B;
  k.OH = d->OH;
  k.OW = d->OW;
  k.Cout = d->Cout;
  k.act = d->act;
  k.slope = d->slope;
  if (halo) {
    k.tiles_x = (d->OW + 8 * k.sub - 1) / (8 * k.sub);
    k.tiles_y = (d->OH + CVH_ROWS - 1) / CVH_ROWS;
    // weight stages: a whole kernel row (3 taps) per barrier round.  The tensor pipe queues only a few MMAs, so the
    // ~200-300 cycles the issuing warp spends between two rounds (barrier check, ring bookkeeping) are mostly idle
    // pipe time: fewer, longer rounds win even when the ring gets shallower (128-wide tile: 2 stages of 48 KB instead
    // of 7 of 16 KB measured 5-9 % faster on the three-segment layers, profiles/r02u_time_conv.log)
    k.tps = 3;
    int S = 4;
    // two CTAs per SM for the 64-channel layers that run M = 128 items on a full grid: half the shared memory each
    // (measured, profiles/r02k_time_conv.log: once the drain was specialised the one-CTA M = 256 form wins everywhere,
    // so this form is only taken on request)
    k.two = (k.NT == 64 && k.sub == 1 && k.ksplit == 1 && d->tile_hint == 1) ? 1 : 0;
    if (k.two) {
      k.tps = 1;
      S = 4;
    }
    while (S > 2 && conv_halo_smem(k.sub, k.NT, S, k.tps, k.n_ntiles) > (k.two ? 113 : 227) * 1024) --S;
    if (k.two && conv_halo_smem(k.sub, k.NT, S, k.tps, k.n_ntiles) > 113 * 1024) k.two = 0;
    k.stages = S;
    p->smem = conv_halo_smem(k.sub, k.NT, S, k.tps, k.n_ntiles);
  } else {
    k.tiles_x = (d->OW + CV_TW - 1) / CV_TW;
    k.tiles_y = (d->OH + CV_TH - 1) / CV_TH;
    const size_t stage_bytes = 32768 + (((size_t)k.NT * 256 + 1023) & ~(size_t)1023);
    int S = (int)((192 * 1024) / stage_bytes);
    if (S > 6) S = 6;
    if (S > k.total_chunks) S = k.total_chunks < 2 ? 2 : k.total_chunks;
    k.stages = S;
    k.staged = (k.NT == 64 || k.NT == 128) && d->out_hi != nullptr && d->out_f32 == nullptr;
    p->smem = 1024 + S * stage_bytes + (2 * S + 4) * 8 + 16 + 128 + 8 * 256 * 16;  // + store staging, 4 KB per drain warp
  }
  const int items = k.B * k.tiles_x * k.tiles_y * k.n_ntiles;
  // Split-K over a thread-block cluster (conv_halo.cuh): for layers whose item count leaves most of the machine idle
  // (12x16 ... 24x32 maps: 16-64 items, each streaming all of its weights through one SM), `ksplit` CTAs share an item.
  if (k.ksplit > 1) {
    p->grid = items * k.ksplit;
  } else {
  int cap = n_sm;
  if (d->max_ctas > 0 && cap > d->max_ctas) cap = d->max_ctas;
  if (k.two) cap *= 2;
  // (a "balanced" grid of ceil(items / rounds) CTAs was tried and loses: 384 items on 128 CTAs x 3 take 55 us, on 148
  // CTAs 46 us -- the items of the last, partial round run faster on the emptier machine)
  p->grid = items < cap ? items : cap;
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_halo_kernel<2, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_halo_kernel<1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_halo_kernel<1, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_halo_kernel<2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_halo_kernel<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_halo_kernel<2, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_halo_kernel<1, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_halo_kernel<1, 64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_halo_kernel<1, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_halo_kernel<1, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      delete p;
      b200_set_error("conv_create: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return -2;
    }
    attr_done = true;
  }
  *plan_out = p;
  return 0;
}

extern "C" int b200_conv_run(void* plan, void* stream) {
  B200_CHECK_ARG(plan, "conv_run: null plan");
  ConvPlan* p = (ConvPlan*)plan;
  cudaStream_t st = (cudaStream_t)stream;
  // Programmatic dependent launch: the kernels call griddepcontrol.wait after their prologue (barrier init, TMEM
  // allocation, tensor-map prefetch, bias staging), so on SMs the previous kernel has already left, the prologue
  // of this one overlaps the previous kernel's tail.  Captured into CUDA graphs as programmatic edges.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p->grid);
  cfg.blockDim = dim3(p->k.halo ? CVH_THREADS : CV_THREADS);
  cfg.dynamicSmemBytes = p->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (p->k.ksplit > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (unsigned)p->k.ksplit;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  cudaError_t le;
  if (!p->k.halo)
    le = cudaLaunchKernelEx(&cfg, conv_tc_kernel, p->k);
  else if (p->k.ksplit > 1 && p->k.NT == 128) le = cudaLaunchKernelEx(&cfg, conv_halo_kernel<1, 128, true>, p->k);
  else if (p->k.ksplit > 1) le = cudaLaunchKernelEx(&cfg, conv_halo_kernel<1, 64, true>, p->k);
  else if (p->k.NT == 128) le = cudaLaunchKernelEx(&cfg, conv_halo_kernel<1, 128>, p->k);
  else if (p->k.NT == 16 && p->k.sub == 2) le = cudaLaunchKernelEx(&cfg, conv_halo_kernel<2, 16>, p->k);
  else if (p->k.NT == 16) le = cudaLaunchKernelEx(&cfg, conv_halo_kernel<1, 16>, p->k);
  else if (p->k.NT == 32 && p->k.sub == 2) le = cudaLaunchKernelEx(&cfg, conv_halo_kernel<2, 32>, p->k);
  else if (p->k.NT == 32) le = cudaLaunchKernelEx(&cfg, conv_halo_kernel<1, 32>, p->k);
  else if (p->k.sub == 2) le = cudaLaunchKernelEx(&cfg, conv_halo_kernel<2, 64>, p->k);
  else if (p->k.two) le = cudaLaunchKernelEx(&cfg, conv_halo_kernel<1, 64, false, true>, p->k);
  else le = cudaLaunchKernelEx(&cfg, conv_halo_kernel<1, 64>, p->k);
  if (le != cudaSuccess) {
    b200_set_error("conv_run: launch failed: %s", cudaGetErrorString(le));
    return -2;
  }
  B200_CHECK_LAUNCH("conv_run");
  return 0;
}

extern "C" int b200_conv_destroy(void* plan) {
  delete (ConvPlan*)plan;
  return 0;
}
