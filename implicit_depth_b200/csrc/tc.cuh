// Thin inline-PTX layer over the sm_100a tensor-core path: mbarrier, TMEM allocation,
// tcgen05.mma (operands from shared memory or TMEM), tcgen05.ld/st, descriptors, and the
// split-bf16 ("bf16x3") number format used to get fp32-grade results out of bf16 MMAs.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// named barrier among a subset of warps (id 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void named_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Bounded wait: a protocol bug traps (launch error the host sees) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 26); ++it)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}

// generic-proxy writes (st.shared) -> visible to the async proxy (tcgen05.mma / TMA reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ programmatic dependent launch
// Wait until the grids this one depends on have completed and their memory is visible (no-op when the kernel was
// launched without the programmatic-serialization attribute), then let the next kernel in the stream begin launching:
// its CTAs become resident as ours exit and run their prologue up to their own wait.
__device__ __forceinline__ void grid_dependency_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ------------------------------------------------------------------ thread-block clusters / distributed shared memory
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// barrier over all threads of all CTAs of the cluster; orders shared-memory writes before it against DSMEM reads after
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `smem_addr` (a shared::cta address of this CTA's window) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_cluster_f4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(cluster_addr)
               : "memory");
  return v;
}

// ------------------------------------------------------------------ TMA
// 4-D tiled tensor-map load (innermost coordinate first), completion signalled on an mbarrier.
__device__ __forceinline__ void tma_load_4d(void* dst_smem, const void* tmap, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// Same box, but only into L2 (no shared-memory destination, no barrier): hides the DRAM latency of a box that will be
// loaded a little later.
__device__ __forceinline__ void tma_prefetch_4d(const void* tmap, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(tmap), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 4-D tiled tensor-map store shared -> global (out-of-bounds elements are dropped), bulk-group completion.
__device__ __forceinline__ void tma_store_4d(const void* tmap, const void* src_smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap),
               "r"(smem_u32(src_smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// 1-D bulk copy global -> shared (16-byte aligned, size multiple of 16).
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ------------------------------------------------------------------ TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32-bit, 16 consecutive columns: thread i of warp w touches TMEM lane 32*(w%4)+i.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
          taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// ------------------------------------------------------------------ descriptors
// Instruction descriptor (cute::UMMA::InstrDescriptor bit layout): bf16 x bf16 -> f32,
// both operands K-major, dense, no negate.
__host__ __device__ constexpr uint32_t idesc_bf16_f32(uint32_t M, uint32_t N) {
  return (1u << 4)      // c_format = F32
         | (1u << 7)    // a_format = BF16
         | (1u << 10)   // b_format = BF16
         | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// Shared-memory matrix descriptor, K-major operand in the 128-byte-swizzled canonical layout:
// rows are 128 B (64 bf16) apart, 8-row groups 1024 B apart (SBO); the 16-byte chunk index of a
// row is XORed with (row & 7).  The tile base must be 1024-byte aligned.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes = 1024) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// Same for the 64-byte-swizzled layout: rows are 64 B (32 bf16) apart, 8-row groups `sbo_bytes` apart (512 when
// dense); the 16-byte chunk index of a row is XORed with address bits 7..8 (= (row >> 1) & 3 for a 512-aligned tile).
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t smem_addr, uint32_t sbo_bytes = 512) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
         (4ull << 61);
}
// Byte offset of element (row, k) of a [rows x 32] bf16 K-major SW64 tile.
__device__ __forceinline__ uint32_t sw64_offset(uint32_t row, uint32_t k) {
  return row * 64u + ((((k >> 3) ^ ((row >> 1) & 3u)) << 4) | ((k & 7u) << 1));
}
// Byte offset of element (row, k) of a [rows x 64] bf16 K-major SW128 tile.
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t k) {
  return row * 128u + ((((k >> 3) ^ (row & 7u)) << 4) | ((k & 7u) << 1));
}

// ------------------------------------------------------------------ MMA issue (one thread)
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// One lane of a converged warp (elect.sync): the role loops are executed by the whole warp so that the
// compiler keeps descriptors in uniform registers; only the issue itself is predicated.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// Split-bf16 product of one 64-wide K chunk (4 k-steps of 16): D (+)= Ah*Bh + Ah*Bl + Al*Bh.
// Descriptors advance by 32 bytes (= 2 in the encoded start-address field) per k-step.
// `fresh` != 0 makes the very first MMA overwrite the accumulator.
template <int KSTEPS>
__device__ __forceinline__ void mma_split_ss(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                             uint32_t idesc, uint32_t fresh) {
#pragma unroll
  for (int k = 0; k < KSTEPS; ++k) mma_ss(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, (k == 0 && fresh) ? 0u : 1u);
#pragma unroll
  for (int k = 0; k < KSTEPS; ++k) mma_ss(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
#pragma unroll
  for (int k = 0; k < KSTEPS; ++k) mma_ss(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
}
__device__ __forceinline__ void mma_split_ss_n(int ksteps, uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi,
                                               uint64_t b_lo, uint32_t idesc, uint32_t fresh) {
  if (ksteps == 4) mma_split_ss<4>(d_tmem, a_hi, a_lo, b_hi, b_lo, idesc, fresh);
  else if (ksteps == 3) mma_split_ss<3>(d_tmem, a_hi, a_lo, b_hi, b_lo, idesc, fresh);
  else if (ksteps == 2) mma_split_ss<2>(d_tmem, a_hi, a_lo, b_hi, b_lo, idesc, fresh);
  else mma_split_ss<1>(d_tmem, a_hi, a_lo, b_hi, b_lo, idesc, fresh);
}
// Same with the A operand in tensor memory (8 columns per k-step).
template <int KSTEPS>
__device__ __forceinline__ void mma_split_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                             uint32_t idesc, uint32_t fresh) {
#pragma unroll
  for (int k = 0; k < KSTEPS; ++k) mma_ts(d_tmem, a_hi + 8 * k, b_hi + 2 * k, idesc, (k == 0 && fresh) ? 0u : 1u);
#pragma unroll
  for (int k = 0; k < KSTEPS; ++k) mma_ts(d_tmem, a_hi + 8 * k, b_lo + 2 * k, idesc, 1u);
#pragma unroll
  for (int k = 0; k < KSTEPS; ++k) mma_ts(d_tmem, a_lo + 8 * k, b_hi + 2 * k, idesc, 1u);
}

// ------------------------------------------------------------------ split-bf16 numbers
// x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits; the product of two such
// numbers is taken as hi*hi + hi*lo + lo*hi (three bf16 MMAs, fp32 accumulate), dropping only the
// lo*lo term (~2^-18 relative).
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);  // .x (low 16 bits) = x0
  float r0 = x0 - __bfloat162float(h.x);
  float r1 = x1 - __bfloat162float(h.y);
  __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

}  // namespace tc
