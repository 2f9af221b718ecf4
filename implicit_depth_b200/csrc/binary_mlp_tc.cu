// Binary-occupancy MLP (BinaryMLPNetwork s0 + BDModel.run_mlp_val) fused into one persistent tcgen05 kernel.
//
//   logit(pixel, z) = w3 . elu( W2 . elu( W1f . feat(pixel) + b1 + w_d * z + w_p * prior(pixel) ) + b2 ) + b3
//
// Reference: per rendered plane `run_mlp_val` concatenates [depth, feature_s0 (64), (prior)] per pixel and runs
// Linear(65|66,128) -> ELU -> Linear(128,128) -> ELU -> Linear(128,1) (experiment_modules/bd_model.py:293-304,
// 412-442; modules/networks.py:87-115); `infer_depth` repeats that 12 times inside a per-pixel bisection
// (bd_model.py:273-292).  Here a 128-pixel tile of the 64-channel feature map is TMA-loaded ONCE into shared
// memory and every plane / bisection step of the tile is evaluated from it:
//   layer 1  feat (smem, split-bf16) x W1f (smem)  -> TMEM accumulator, 3 bf16 passes (fp32-grade);
//            the depth and prior columns of W1 are rank-1 terms added in fp32 by the row threads
//   layer 2  H1 re-split by the row threads into TENSOR MEMORY (A operand) x W2 (smem) -> same accumulator
//   layer 3  dot with w3 in registers; planes mode stores the logit, search mode updates the bisection bounds.
// Two row groups ping-pong so one group's MMAs overlap the other's ELU / split work.  A group is 8 warps: a pixel
// (= TMEM lane) is shared by TWO threads, each owning 64 of the 128 hidden channels (warps w and w+4 of a group read
// the same TMEM lane quarter), so 16 row warps hide the SFU / TMEM latencies of the two epilogues; the halves of the
// layer-3 dot product meet through shared memory.  Warp 16 issues all MMAs, warp 17 is the TMA producer.
// TMEM (512 columns): group g: accumulator [256g, 256g+128), H1 operand [256g+128, 256g+256) (64 hi | 64 lo).
#include "common.cuh"
#include "tc.cuh"
#include "tmap.cuh"

#define BM_THREADS 576
#define BM_MMA_WARP 16
#define BM_TMA_WARP 17
#define BM_ROWS 128
#define BM_FEAT_C 64

struct BmParams {
  CUtensorMap map_hi, map_lo;  // 2-D [64 channels, npix] bf16, box [64, 128], SWIZZLE_128B
  const uint8_t* wimage;       // W1f hi | W1f lo | W2c0 hi | W2c0 lo | W2c1 hi | W2c1 lo, 16 KB each
  const float* vecs;           // [6][128]: b1, w_depth, w_prior, b2, w3, {b3, ...}
  const float* depth;          // planes mode: [B, P, HW]
  const float* prior;          // [B, HW] or null
  float* pred;                 // planes: [B, P, HW]; search: [B, HW] (logit of the last step)
  float* search;               // search mode: [B, HW]
  long long npix;
  int HW, P, use_prior;
  float lo0, hi0, z0;          // search mode: initial bounds and first query depth
  const float* thr_bins;       // search mode: depth-dependent thresholds (Thresholder, binary_metrics_utils.py:42-52):
  const float* thr_vals;       //   threshold = thr_vals[#{i : thr_bins[i] < z}] (torch.bucketize); null = 0.5
  int n_thr;
};

struct BmSync {
  uint64_t feat_full, feat_empty, acc1_full, h_full, acc2_full, acc_free;
};

template <bool SEARCH, bool PRIOR>
__global__ void __launch_bounds__(BM_THREADS, 1) binary_mlp_tc_kernel(const __grid_constant__ BmParams prm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w1_hi = base;
  uint8_t* w1_lo = base + 16384;
  uint8_t* w2 = base + 32768;             // chunk c: hi at w2 + c*32768, lo at +16384
  uint8_t* feat = base + 98304;           // group g: hi at feat + g*32768, lo at +16384
  float4* e1_s = reinterpret_cast<float4*>(base + 163840);  // [64] {b1[c], w_d[c], b1[c+1], w_d[c+1]}, c = 2i
  float4* e2_s = e1_s + 64;                                 // [64] {b2[c], w3[c], b2[c+1], w3[c+1]}
  float2* wp_s = reinterpret_cast<float2*>(e2_s + 64);      // [64] {w_p[c], w_p[c+1]}
  float* part_s = reinterpret_cast<float*>(wp_s + 64);      // [2 groups][2 parities][2 halves][128] layer-3 partials
  BmSync* sync = reinterpret_cast<BmSync*>(part_s + 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sync + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long n_tiles = (prm.npix + BM_ROWS - 1) / BM_ROWS;
  const int G = 2 * gridDim.x;
  const int steps = prm.P;  // evaluations per tile (rendered planes, or bisection iterations)

  for (int i = tid; i < 98304 / 16; i += BM_THREADS)
    reinterpret_cast<uint4*>(base)[i] = __ldg(reinterpret_cast<const uint4*>(prm.wimage) + i);
  if (tid < 64) {
    const float* v = prm.vecs;
    const int c = 2 * tid;
    e1_s[tid] = make_float4(v[c], v[128 + c], v[c + 1], v[128 + c + 1]);
    e2_s[tid] = make_float4(v[384 + c], v[512 + c], v[384 + c + 1], v[512 + c + 1]);
    wp_s[tid] = make_float2(v[256 + c], v[256 + c + 1]);
  }
  const float b3v = __ldg(prm.vecs + 640);
  if (warp == BM_MMA_WARP) {
    tc::tmem_alloc(tmem_slot, 512);
    if (lane == 0) {
      for (int g = 0; g < 2; ++g) {
        tc::mbar_init(&sync[g].feat_full, 1);
        tc::mbar_init(&sync[g].feat_empty, 1);
        tc::mbar_init(&sync[g].acc1_full, 1);
        tc::mbar_init(&sync[g].h_full, 256);
        tc::mbar_init(&sync[g].acc2_full, 1);
        tc::mbar_init(&sync[g].acc_free, 256);
      }
      tc::mbar_fence_init();
    }
  }
  if (warp == BM_TMA_WARP && lane == 0) {
    tc::prefetch_tmap(&prm.map_hi);
    tc::prefetch_tmap(&prm.map_lo);
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < 16) {
    // =============================== row groups: two threads per pixel (= TMEM lane) ======================
    const int g = warp >> 3;
    const int quarter = warp & 3, ch = (warp >> 2) & 1;  // TMEM lane quarter; which 64 hidden channels
    const int row = quarter * 32 + lane;
    BmSync* gs = &sync[g];
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t acc = tmem + lane_base + g * 256;
    const uint32_t a_op = acc + 128;
    uint32_t n = 0;  // evaluation steps done so far by this group (barrier parity)
    for (long long t = blockIdx.x * 2 + g; t < n_tiles; t += G) {
      const long long pix = t * BM_ROWS + row;
      const bool live = pix < prm.npix;
      const long long pc = live ? pix : prm.npix - 1;
      const int b = (int)(pc / prm.HW), pin = (int)(pc - (long long)b * prm.HW);
      float pr = 0.f;
      if (PRIOR) pr = prm.prior ? __ldg(prm.prior + (size_t)b * prm.HW + pin) : -1.f;  // bd_model.py:433-434
      float lo = prm.lo0, hi = prm.hi0, z = prm.z0, logit = 0.f;
      const float* zp = SEARCH ? nullptr : prm.depth + (size_t)b * prm.P * prm.HW + pin;
      float z_next = SEARCH ? 0.f : __ldg(zp);
      for (int s = 0; s < steps; ++s, ++n) {
        if (!SEARCH) {  // the plane depth of the NEXT step is fetched now: a DRAM miss hides under this step
          z = z_next;
          if (s + 1 < steps) z_next = __ldg(zp + (size_t)(s + 1) * prm.HW);
        }
        // ---- epilogue 1: H1 = elu(acc + b1 + w_d z + w_p prior) -> split -> TMEM A operand ----
        tc::mbar_wait(&gs->acc1_full, n & 1u);
        tc::fence_after_sync();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t r[32];
          tc::tmem_ld32(acc + 64 * ch + 32 * q, r);
          tc::wait_ld();
          uint32_t h[16], l[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int i = 32 * ch + 16 * q + j;  // channel pair
            const float4 e = e1_s[i];
            float v0 = __uint_as_float(r[2 * j]) + fmaf(e.y, z, e.x);
            float v1 = __uint_as_float(r[2 * j + 1]) + fmaf(e.w, z, e.z);
            if (PRIOR) {
              const float2 w = wp_s[i];
              v0 = fmaf(w.x, pr, v0);
              v1 = fmaf(w.y, pr, v1);
            }
            tc::split2(elu1(v0), elu1(v1), h[j], l[j]);
          }
          tc::tmem_st16(a_op + 32 * ch + 16 * q, h);
          tc::tmem_st16(a_op + 64 + 32 * ch + 16 * q, l);
        }
        tc::wait_st();
        tc::fence_before_sync();
        tc::mbar_arrive(&gs->h_full);
        // ---- epilogue 2: logit = w3 . elu(acc + b2) + b3; each thread sums its 64 channels ----
        tc::mbar_wait(&gs->acc2_full, n & 1u);
        tc::fence_after_sync();
        float o = 0.f;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t r[32];
          tc::tmem_ld32(acc + 64 * ch + 32 * q, r);
          tc::wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 e = e2_s[32 * ch + 16 * q + j];
            o = fmaf(elu1(__uint_as_float(r[2 * j]) + e.x), e.y, o);
            o = fmaf(elu1(__uint_as_float(r[2 * j + 1]) + e.z), e.w, o);
          }
        }
        tc::fence_before_sync();
        tc::mbar_arrive(&gs->acc_free);
        // the two halves of the dot product meet in shared memory (double-buffered by step parity: the barrier of
        // step n+1 orders these reads before the writes of step n+2); both threads form the same sum
        float* part = part_s + ((g * 2 + (n & 1u)) * 2) * 128;
        part[ch * 128 + row] = o;
        asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
        logit = (part[row] + part[128 + row]) + b3v;
        if (SEARCH) {
          // bd_model.py:281-291: visible = sigmoid(pred) < threshold(z); max_bound[visible] = z; min_bound[~visible] = z
          float thr = 0.5f;
          if (prm.thr_bins != nullptr) {  // Thresholder.get_thresholds: torch.bucketize(z, bins), right = False
            int idx = 0;
            for (int i = 0; i < prm.n_thr; ++i) idx += (__ldg(prm.thr_bins + i) < z) ? 1 : 0;
            thr = __ldg(prm.thr_vals + min(idx, prm.n_thr - 1));
          }
          const bool visible = (1.f / (1.f + expf(-logit))) < thr;
          if (visible) hi = z; else lo = z;
          z = (hi + lo) / 2.f;
        } else if (live && ch == 0) {
          prm.pred[((size_t)b * prm.P + s) * prm.HW + pin] = logit;
        }
      }
      if (SEARCH && live && ch == 0) {
        prm.pred[(size_t)b * prm.HW + pin] = logit;
        prm.search[(size_t)b * prm.HW + pin] = z;
      }
    }
  } else if (warp == BM_MMA_WARP) {
    // =============================== MMA issuer (whole warp loops, one elected lane issues) ==========
    constexpr uint32_t IDESC = tc::idesc_bf16_f32(128, 128);
    long long t_cur[2];
    uint32_t n[2] = {0, 0}, tiles[2] = {0, 0};
    int s[2] = {0, 0}, stage[2] = {0, 0};
    for (int g = 0; g < 2; ++g) t_cur[g] = blockIdx.x * 2 + g;
    while (t_cur[0] < n_tiles || t_cur[1] < n_tiles) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (t_cur[g] >= n_tiles) continue;
        BmSync* gs = &sync[g];
        const uint32_t acc = tmem + g * 256;
        if (stage[g] == 0) {
          // layer 1 of evaluation n[g]: needs the feature tile (first step of a tile) and a free accumulator
          bool ok = true;
          if (s[g] == 0) ok = tc::mbar_try_wait(&gs->feat_full, tiles[g] & 1u);
          if (ok && n[g] > 0) ok = tc::mbar_try_wait(&gs->acc_free, (n[g] - 1) & 1u);
          if (!__all_sync(0xffffffffu, ok)) continue;
          tc::fence_after_sync();
          if (tc::elect_one()) {
            const uint64_t a_hi = tc::smem_desc_sw128(tc::smem_u32(feat + g * 32768));
            const uint64_t a_lo = tc::smem_desc_sw128(tc::smem_u32(feat + g * 32768 + 16384));
            const uint64_t b_hi = tc::smem_desc_sw128(tc::smem_u32(w1_hi));
            const uint64_t b_lo = tc::smem_desc_sw128(tc::smem_u32(w1_lo));
            tc::mma_split_ss<4>(acc, a_hi, a_lo, b_hi, b_lo, IDESC, 1u);
            tc::mma_commit(&gs->acc1_full);
            if (s[g] + 1 == steps) tc::mma_commit(&gs->feat_empty);
          }
          __syncwarp();
          stage[g] = 1;
        } else {
          if (!__all_sync(0xffffffffu, tc::mbar_try_wait(&gs->h_full, n[g] & 1u))) continue;
          tc::fence_after_sync();
          if (tc::elect_one()) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint64_t b_hi = tc::smem_desc_sw128(tc::smem_u32(w2 + half * 32768));
              const uint64_t b_lo = tc::smem_desc_sw128(tc::smem_u32(w2 + half * 32768 + 16384));
              tc::mma_split_ts<4>(acc, acc + 128 + half * 32, acc + 128 + 64 + half * 32, b_hi, b_lo, IDESC, half == 0);
            }
            tc::mma_commit(&gs->acc2_full);
          }
          __syncwarp();
          stage[g] = 0;
          ++n[g];
          if (++s[g] == steps) {
            s[g] = 0;
            ++tiles[g];
            t_cur[g] += G;
          }
        }
      }
    }
  } else {
    // =============================== TMA producer: one feature tile per (group, tile) ===============
    long long t_cur[2];
    uint32_t tiles[2] = {0, 0};
    for (int g = 0; g < 2; ++g) t_cur[g] = blockIdx.x * 2 + g;
    while (t_cur[0] < n_tiles || t_cur[1] < n_tiles) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (t_cur[g] >= n_tiles) continue;
        BmSync* gs = &sync[g];
        if (tiles[g] > 0 && !__all_sync(0xffffffffu, tc::mbar_try_wait(&gs->feat_empty, (tiles[g] - 1) & 1u))) continue;
        if (tc::elect_one()) {
          tc::mbar_expect_tx(&gs->feat_full, 32768u);
          const int p0 = (int)(t_cur[g] * BM_ROWS);
          tc::tma_load_2d(feat + g * 32768, &prm.map_hi, 0, p0, &gs->feat_full);
          tc::tma_load_2d(feat + g * 32768 + 16384, &prm.map_lo, 0, p0, &gs->feat_full);
        }
        __syncwarp();
        ++tiles[g];
        t_cur[g] += G;
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == BM_MMA_WARP) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
struct BmPlan {
  BmParams k;
  int grid;
};

struct b200_binary_mlp_desc {
  const void* feat_hi;  // NHWC bf16 [npix, 64]
  const void* feat_lo;
  long long npix;       // B * H * W
  int HW;               // pixels per frame
  const void* wimage;   // 96 KB: W1f hi|lo, W2 chunk0 hi|lo, W2 chunk1 hi|lo (SW128 K-major tiles of 128 x 64 bf16)
  const float* vecs;    // [6][128] fp32: b1, W1[:,0] (depth), W1[:,65] (prior) or 0, b2, w3, {b3, 0...}
  int use_prior;
};

static const size_t BM_SMEM = 1024 + 163840 + 2 * 64 * 16 + 64 * 8 + 1024 * 4 + 2 * sizeof(BmSync) + 16;

extern "C" int b200_binary_mlp_create(const b200_binary_mlp_desc* d, void** plan_out) {
  B200_CHECK_ARG(d && plan_out, "binary_mlp_create: null pointer");
  B200_CHECK_ARG(d->feat_hi && d->feat_lo && d->wimage && d->vecs, "binary_mlp_create: null buffer");
  B200_CHECK_ARG(d->npix > 0 && d->HW > 0 && d->npix % d->HW == 0, "binary_mlp_create: npix must be B * HW");
  B200_CHECK_ARG((((uintptr_t)d->feat_hi | (uintptr_t)d->feat_lo | (uintptr_t)d->wimage) & 15) == 0,
                 "binary_mlp_create: buffers must be 16-byte aligned");
  EncodeTiledFn enc = get_encode();
  B200_CHECK_ARG(enc != nullptr, "binary_mlp_create: cuTensorMapEncodeTiled not available from the driver");
  BmPlan* p = new BmPlan();
  memset(&p->k, 0, sizeof(p->k));
  cuuint64_t gdim[2] = {BM_FEAT_C, (cuuint64_t)d->npix};
  cuuint64_t gstr[1] = {BM_FEAT_C * 2};
  cuuint32_t box[2] = {BM_FEAT_C, BM_ROWS};
  cuuint32_t estr[2] = {1, 1};
  for (int part = 0; part < 2; ++part) {
    CUresult r = enc(part ? &p->k.map_lo : &p->k.map_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void*>(part ? d->feat_lo : d->feat_hi), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete p;
      b200_set_error("binary_mlp_create: cuTensorMapEncodeTiled failed (%d)", (int)r);
      return -2;
    }
  }
  p->k.wimage = (const uint8_t*)d->wimage;
  p->k.vecs = d->vecs;
  p->k.npix = d->npix;
  p->k.HW = d->HW;
  p->k.use_prior = d->use_prior;
  int dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const long long n_tiles = (d->npix + BM_ROWS - 1) / BM_ROWS;
  p->grid = (int)((n_tiles + 1) / 2 < n_sm ? (n_tiles + 1) / 2 : n_sm);
  cudaError_t e = cudaSuccess;
  const void* kernels[4] = {(const void*)binary_mlp_tc_kernel<false, false>, (const void*)binary_mlp_tc_kernel<false, true>,
                            (const void*)binary_mlp_tc_kernel<true, false>, (const void*)binary_mlp_tc_kernel<true, true>};
  for (int i = 0; i < 4 && e == cudaSuccess; ++i)
    e = cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BM_SMEM);
  if (e != cudaSuccess) {
    delete p;
    b200_set_error("binary_mlp_create: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    return -2;
  }
  *plan_out = p;
  return 0;
}

extern "C" int b200_binary_mlp_planes(void* plan, const float* depth, int P, const float* prior, float* pred,
                                      void* stream) {
  B200_CHECK_ARG(plan && depth && pred && P > 0, "binary_mlp_planes: bad arguments");
  BmPlan* p = (BmPlan*)plan;
  BmParams k = p->k;
  k.depth = depth;
  k.prior = prior;
  k.pred = pred;
  k.P = P;
  if (k.use_prior)
    binary_mlp_tc_kernel<false, true><<<p->grid, BM_THREADS, BM_SMEM, (cudaStream_t)stream>>>(k);
  else
    binary_mlp_tc_kernel<false, false><<<p->grid, BM_THREADS, BM_SMEM, (cudaStream_t)stream>>>(k);
  B200_CHECK_LAUNCH("binary_mlp_planes");
  return 0;
}

extern "C" int b200_binary_mlp_search(void* plan, const float* prior, int iters, float min_bound, float max_bound,
                                      float first_depth, const float* thr_bins, const float* thr_vals, int n_thr,
                                      float* search_out, float* pred_out, void* stream) {
  B200_CHECK_ARG(plan && search_out && pred_out && iters > 0, "binary_mlp_search: bad arguments");
  B200_CHECK_ARG((thr_bins == nullptr) == (thr_vals == nullptr) && (thr_bins == nullptr || n_thr > 0),
                 "binary_mlp_search: thr_bins / thr_vals come together with n_thr > 0");
  BmPlan* p = (BmPlan*)plan;
  BmParams k = p->k;
  k.prior = prior;
  k.pred = pred_out;
  k.search = search_out;
  k.P = iters;
  k.lo0 = min_bound;
  k.hi0 = max_bound;
  k.z0 = first_depth;
  k.thr_bins = thr_bins;
  k.thr_vals = thr_vals;
  k.n_thr = thr_bins ? n_thr : 0;
  if (k.use_prior)
    binary_mlp_tc_kernel<true, true><<<p->grid, BM_THREADS, BM_SMEM, (cudaStream_t)stream>>>(k);
  else
    binary_mlp_tc_kernel<true, false><<<p->grid, BM_THREADS, BM_SMEM, (cudaStream_t)stream>>>(k);
  B200_CHECK_LAUNCH("binary_mlp_search");
  return 0;
}

extern "C" int b200_binary_mlp_destroy(void* plan) {
  delete (BmPlan*)plan;
  return 0;
}
