// Dev probe: issue rate of tcgen05.mma (kind::f16, M=128, cta_group::1) as a function of N, operand source
// (A from shared memory vs tensor memory) and accumulator pattern.  One thread per CTA issues `iters`
// groups of 12 MMAs (the split-bf16 K=64 chunk pattern of the conv / feature-volume kernels) over
// uninitialised shared-memory tiles and reports elapsed SM clocks.  Not part of the product path.
#include "../common.cuh"
#include "../tc.cuh"

__global__ void __launch_bounds__(128) mma_rate_kernel(long long* out, int N, int mode, int iters) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (192 * 1024) / 16; i += 128) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = tc::idesc_bf16_f32(128, N);
    const uint32_t a0 = tc::smem_u32(base);               // 4 x (hi 16 KB | lo 16 KB) A chunks = 128 KB
    const uint32_t b0 = tc::smem_u32(base + 128 * 1024);  // B: hi N*128 B | lo N*128 B  (<= 64 KB)
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t a = a0 + (it & 3) * 32768u;
      const uint64_t a_hi = tc::smem_desc_sw128(a), a_lo = tc::smem_desc_sw128(a + 16384u);
      const uint64_t b_hi = tc::smem_desc_sw128(b0), b_lo = tc::smem_desc_sw128(b0 + N * 128u);
      const uint32_t acc = tmem + ((mode & 4) ? (it & 1) * 256u : 0u);
      if ((mode & 24) == 0 && (mode & 3) == 0) {
        tc::mma_split_ss<4>(acc, a_hi, a_lo, b_hi, b_lo, idesc, it == 0);
      } else if ((mode & 24) == 0 && (mode & 3) == 1) {  // A from TMEM columns 256.. (hi 32 cols | lo 32 cols)
        tc::mma_split_ts<4>(acc, tmem + 256 + 128, tmem + 256 + 160, b_hi, b_lo, idesc, it == 0);
      } else if ((mode & 24) == 0 && (mode & 3) == 2) {  // SS, plain bf16: 12 k-steps hi*hi only over 3 chunks (no operand reuse)
#pragma unroll
        for (int k = 0; k < 12; ++k)
          tc::mma_ss(acc, tc::smem_desc_sw128(a0 + ((it + k / 4) & 3) * 32768u) + 2 * (k & 3), b_hi + 2 * (k & 3), idesc,
                     (it == 0 && k == 0) ? 0u : 1u);
      } else if (mode & 8) {  // SS, 64-byte swizzle, conv-halo pattern: per k-step A_hi x [B_hi|B_lo] (N) + A_lo x B_hi (N/2)
        const uint32_t idh = tc::idesc_bf16_f32(128, N / 2);
        const uint64_t a64 = tc::smem_desc_sw64(a + 11 * 64, 1152), a64l = tc::smem_desc_sw64(a + 16384u + 11 * 64, 1152);
        const uint64_t b64 = tc::smem_desc_sw64(b0);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          tc::mma_ss(acc, a64 + 2 * (k & 1), b64 + 2 * (k & 1), idesc, (it == 0 && k == 0) ? 0u : 1u);
          tc::mma_ss(acc, a64l + 2 * (k & 1), b64 + 2 * (k & 1), idh, 1u);
        }
      } else if (mode & 16) {  // SS, 128-byte swizzle, same instruction mix as mode 8
        const uint32_t idh = tc::idesc_bf16_f32(128, N / 2);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          tc::mma_ss(acc, a_hi + 2 * (k & 3), b_hi + 2 * (k & 3), idesc, (it == 0 && k == 0) ? 0u : 1u);
          tc::mma_ss(acc, a_lo + 2 * (k & 3), b_hi + 2 * (k & 3), idh, 1u);
        }
      } else {  // SS, the same A and B k-slice every time (best-case operand locality)
#pragma unroll
        for (int k = 0; k < 12; ++k) tc::mma_ss(acc, a_hi, b_hi, idesc, (it == 0 && k == 0) ? 0u : 1u);
      }
    }
    tc::mma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

extern "C" int b200_mma_rate(long long* out_cycles, int N, int mode, int iters, int grid, void* stream) {
  B200_CHECK_ARG(out_cycles && N >= 16 && N <= 256 && N % 16 == 0 && iters > 0 && grid > 0, "mma_rate: bad arguments");
  const int smem = 193 * 1024 + 1024;
  B200_CHECK_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mma_rate_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(out_cycles, N, mode, iters);
  B200_CHECK_LAUNCH("mma_rate");
  return 0;
}
