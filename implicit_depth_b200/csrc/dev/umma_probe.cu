// Self-test of the tcgen05 building blocks used by the tensor-core kernels: one CTA computes
// D[128 x N] = A[128 x K] * B[N x K]^T with
//   mode 0: A and B as bf16 from shared memory (SS)            -> checks smem descriptors / swizzle
//   mode 1: A as bf16 from tensor memory (TS), B from smem     -> checks the TMEM A layout
//   mode 2: split-bf16 (hi*hi + hi*lo + lo*hi), A from TMEM    -> checks the fp32-grade number format
//   mode 3: split-bf16, A from shared memory
//   mode 4: bf16 SS with the A rows taken from inside a larger swizzled "halo patch": M-tile row r lives
//           at patch row (r/8)*10 + (r%8) + 11, i.e. descriptor start shifted by a non-multiple of 8 rows
//           and an 8-row-group stride of 1280 B -- the addressing an implicit-GEMM conv tap needs
//   mode 5: bf16 SS, 64-byte swizzle, K = 32: A rows inside a halo patch of 64-byte pixel records (patch row
//           (r/8)*10 + (r%8) + 11, 8-row-group stride 640 B), B as a dense [N x 32] SW64 tile -- the layout of
//           the 32-channel-block conv kernel
// The GPU tests compare D against a host matmul, so a layout mistake shows up as a numeric error
// in a 30-line kernel instead of inside the fused ones.
#include "../common.cuh"
#include "../tc.cuh"

#define PROBE_MAX_K 192
#define PROBE_MAX_N 128

__global__ void __launch_bounds__(128)
umma_probe_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ D, int K, int N,
                  int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: 1024-aligned tiles
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nchunk = K / 64;
  uint8_t* a_hi = base;                                  // nchunk x [128 x 64] bf16 (16 KB each)
  uint8_t* a_lo = a_hi + nchunk * 16384;
  uint8_t* b_hi = a_lo + nchunk * 16384;                 // nchunk x [N x 64] bf16
  uint8_t* b_lo = b_hi + nchunk * N * 128;
  // mode 5 (K = 32: nchunk = 0) lays its own patch / weight tiles out from `base`: keep the barrier and the TMEM slot
  // behind them (racecheck, profiles/r02h_sanitizer.md: they used to alias the patch that mode 5 zero-fills)
  uint8_t* tail = (mode == 5) ? base + 16384 + 8192 : b_lo + nchunk * N * 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tail);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 8);

  const int tid = threadIdx.x, warp = tid >> 5;
  const bool split = (mode == 2 || mode == 3);
  const bool halo = (mode == 4);

  const bool a_in_tmem = (mode == 1 || mode == 2);

  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    tc::mbar_init(bar, 1);
    tc::mbar_fence_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  if (mode == 5) {
    // ---- 64-byte-swizzle variant, self-contained ----
    uint8_t* a64 = base;           // patch: up to 171 pixel rows x 64 B
    uint8_t* b64 = base + 16384;   // [N x 32] bf16 SW64
    for (int i = tid; i < 16384 / 4; i += 128) reinterpret_cast<uint32_t*>(a64)[i] = 0u;
    __syncthreads();
    for (int i = tid; i < N * 16; i += 128) {
      const int n = i / 16, k = (i % 16) * 2;
      __nv_bfloat162 v = __floats2bfloat162_rn(Bm[n * K + k], Bm[n * K + k + 1]);
      *reinterpret_cast<__nv_bfloat162*>(b64 + tc::sw64_offset(n, k)) = v;
    }
    for (int k = 0; k < 32; k += 2) {
      const int srow = (tid / 8) * 10 + (tid % 8) + 11;
      __nv_bfloat162 v = __floats2bfloat162_rn(A[tid * K + k], A[tid * K + k + 1]);
      *reinterpret_cast<__nv_bfloat162*>(a64 + tc::sw64_offset(srow, k)) = v;
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      const uint32_t idesc = tc::idesc_bf16_f32(128, N);
      for (int ks = 0; ks < 2; ++ks)
        tc::mma_ss(tmem, tc::smem_desc_sw64(tc::smem_u32(a64) + 11 * 64 + ks * 32, 640),
                   tc::smem_desc_sw64(tc::smem_u32(b64) + ks * 32), idesc, ks);
      tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, 0);
    tc::fence_after_sync();
    for (int n0 = 0; n0 < N; n0 += 16) {
      uint32_t r[16];
      tc::tmem_ld16(tmem + lane_base + n0, r);
      tc::wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) D[tid * N + n0 + j] = __uint_as_float(r[j]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
    return;
  }

  // B tiles (all threads), pairs of k
  for (int i = tid; i < N * K / 2; i += 128) {
    int n = i / (K / 2), k = (i % (K / 2)) * 2;
    uint32_t hi, lo;
    tc::split2(Bm[n * K + k], Bm[n * K + k + 1], hi, lo);
    uint32_t off = (k / 64) * (N * 128) + tc::sw128_offset(n, k % 64);
    *reinterpret_cast<uint32_t*>(b_hi + off) = hi;
    *reinterpret_cast<uint32_t*>(b_lo + off) = lo;
  }
  // A: thread = row
  {
    const int row = tid;
    for (int k0 = 0; k0 < K; k0 += 32) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) tc::split2(A[row * K + k0 + 2 * j], A[row * K + k0 + 2 * j + 1], hi[j], lo[j]);
      if (a_in_tmem) {
        tc::tmem_st16(tmem + lane_base + 256 + k0 / 2, hi);
        tc::tmem_st16(tmem + lane_base + 256 + 96 + k0 / 2, lo);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          int k = k0 + 2 * j;
          const int srow = halo ? (row / 8) * 10 + (row % 8) + 11 : row;
          uint32_t off = (k / 64) * 16384 + tc::sw128_offset(srow, k % 64);
          *reinterpret_cast<uint32_t*>(a_hi + off) = hi[j];
          if (!halo) *reinterpret_cast<uint32_t*>(a_lo + off) = lo[j];  // the halo patch spills into a_lo's space
        }
      }
    }
    if (a_in_tmem) tc::wait_st();
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();

  if (tid == 0) {
    tc::fence_after_sync();
    const uint32_t idesc = tc::idesc_bf16_f32(128, N);
    uint32_t accum = 0;
    const int npass = split ? 3 : 1;
    for (int pass = 0; pass < npass; ++pass) {
      // pass 0: Ah*Bh, pass 1: Ah*Bl, pass 2: Al*Bh
      const uint8_t* as = (pass == 2) ? a_lo : a_hi;
      const uint8_t* bs = (pass == 1) ? b_lo : b_hi;
      const uint32_t a_t = tmem + 256 + ((pass == 2) ? 96 : 0);
      for (int ks = 0; ks < K / 16; ++ks) {
        const int chunk = ks / 4, sub = ks % 4;
        const uint64_t bdesc = tc::smem_desc_sw128(tc::smem_u32(bs + chunk * N * 128) + sub * 32);
        if (a_in_tmem) {
          tc::mma_ts(tmem, a_t + ks * 8, bdesc, idesc, accum);
        } else {
          const uint64_t adesc = halo ? tc::smem_desc_sw128(tc::smem_u32(as) + 11 * 128 + sub * 32, 1280)
                                      : tc::smem_desc_sw128(tc::smem_u32(as + chunk * 16384) + sub * 32);
          tc::mma_ss(tmem, adesc, bdesc, idesc, accum);
        }
        accum = 1;
      }
    }
    tc::mma_commit(bar);
  }
  tc::mbar_wait(bar, 0);
  tc::fence_after_sync();
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t r[16];
    tc::tmem_ld16(tmem + lane_base + n0, r);
    tc::wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[tid * N + n0 + j] = __uint_as_float(r[j]);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

extern "C" int b200_umma_probe(const float* A, const float* Bm, float* D, int K, int N, int mode, void* stream) {
  B200_CHECK_ARG(A && Bm && D, "umma_probe: null pointer");
  B200_CHECK_ARG(mode == 5 ? K == 32 : (K % 64 == 0 && K >= 64 && K <= PROBE_MAX_K),
                 "umma_probe: K must be 64, 128 or 192 (32 for mode 5)");
  B200_CHECK_ARG(N % 16 == 0 && N >= 16 && N <= PROBE_MAX_N, "umma_probe: N must be a multiple of 16 up to 128");
  B200_CHECK_ARG(mode >= 0 && mode <= 5, "umma_probe: mode 0..5");
  B200_CHECK_ARG(mode != 4 || K == 64, "umma_probe: mode 4 needs K == 64");
  const int nchunk = K >= 64 ? K / 64 : 1;
  size_t smem = 1024 + (size_t)nchunk * (2 * 16384 + 2 * N * 128) + 64;
  if (mode == 5) smem = 1024 + 16384 + 8192 + 64;  // patch (16 KB) + weight tile (<= 8 KB) + barrier / TMEM slot
  B200_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, Bm, D, K, N, mode);
  B200_CHECK_LAUNCH("umma_probe");
  return 0;
}
