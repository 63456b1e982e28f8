// Inline-PTX wrappers for the Blackwell (sm_100a) async machinery used by gemm_tc.cu:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / fences).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace bnn {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Explicit shared-space accesses by 32-bit shared address (pointers derived through integer alignment
// arithmetic lose their address space and would compile to generic LD/ST).
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w)
                 : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint2 lds_v2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_b32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Bounded wait: a protocol bug must surface as a trapped launch (cudaErrorLaunchFailure), never
// as a hung GPU box.  10 s is far beyond any legitimate wait in these kernels.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3ffu) == 0 && globaltimer_ns() - t0 > 10000000000ull) {
            printf("bnn_b200: mbarrier wait timed out (block %d thread %d parity %u)\n",
                   (int)blockIdx.x, (int)threadIdx.x, parity);
            __trap();
        }
    }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 3-D tiled load (k, row, batch) -> swizzled smem tile; completes on `bar` with complete_tx.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// The same load delivered to the same shared-memory offset (and completing on the barrier at the same offset) in
// every CTA of the cluster named in cta_mask.
__device__ __forceinline__ void tma_load_3d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                               int c1, int c2, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
        : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(smem_result)), "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; one thread issues for the whole CTA.
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                        uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                       uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Block-scaled FP4: A and B hold E2M1 values, two per byte; every 32 of them along K are multiplied by a UE8M0
// scale factor read from TMEM (sf_a / sf_b: TMEM addresses).  K = 64 elements (32 bytes) per instruction, fp32
// accumulators, twice the rate of kind::i8.  The +-1 GEMMs keep EVERY scale byte at 0x7F (2^0), so the
// scale-factor layout does not matter and both operands point at the same block of constant columns.
__device__ __forceinline__ void mma_mxf4(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate, uint32_t sf_a, uint32_t sf_b) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(sf_a), "r"(sf_b)
        : "memory");
}
// 32 lanes x 1 column store: thread i of the warp writes lane (base + i).
__device__ __forceinline__ void tmem_st_32x32b_x1(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// Arrive on `bar` once every previously issued tcgen05.mma of this thread has completed
// (implicitly performs tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar))
                 : "memory");
}
// Single-CTA MMAs, arrival on the barrier at this offset in every CTA of cta_mask (cluster multicast form).
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives lane (base+i).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7])
        : "r"(taddr)
        : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[N]);
template <>
__device__ __forceinline__ void tmem_ld_cols<8>(uint32_t taddr, uint32_t (&v)[8]) {
    tmem_ld_32x32b_x8(taddr, v);
}
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t (&v)[32]) {
    tmem_ld_32x32b_x32(taddr, v);
}
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t (&v)[16]) {
    tmem_ld_32x32b_x16(taddr, v);
}
// Warp-group register re-allocation (all 4 warps of the group execute it).
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2)
// Two CTAs of a cluster on the SMs of one TPC execute one UMMA of M = 256: each holds its own 128 rows
// of A and HALF of the B tile in shared memory and its own 128 accumulator rows in TMEM; the leader
// (cluster rank 0) issues the instruction for both.  Barriers that the leader waits on live in the
// leader's shared memory and are reached from the peer through the shared::cluster window.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(rank)
        : "memory");
}
// TMA load issued by either CTA of a pair into ITS OWN shared memory; the transaction bytes are
// counted on the LEADER's barrier (peer bit of the shared::cluster address cleared, as CUTLASS'
// SM100_TMA_2SM_LOAD does).
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                                int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result) {  // the same warp of BOTH CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(smem_result)), "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
                 : "memory");
}
__device__ __forceinline__ void mma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_i8_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                           uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far are done.
__device__ __forceinline__ void mma_commit_2sm(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
}

// Compile-time selection between the single-CTA and the pair form, so that a kernel templated on the
// form keeps ONE statement per operation.
template <bool kPair>
__device__ __forceinline__ void mma_f16_t(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if constexpr (kPair) mma_f16_2sm(d, a, b, idesc, acc);
    else mma_f16(d, a, b, idesc, acc);
}
template <bool kPair>
__device__ __forceinline__ void mma_i8_t(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if constexpr (kPair) mma_i8_2sm(d, a, b, idesc, acc);
    else mma_i8(d, a, b, idesc, acc);
}
template <bool kPair>
__device__ __forceinline__ void mma_commit_t(uint64_t* bar) {
    if constexpr (kPair) mma_commit_2sm(bar);
    else mma_commit(bar);
}
template <bool kPair, uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_t(uint32_t* smem_result) {
    if constexpr (kPair) tmem_alloc_2sm<kCols>(smem_result);
    else tmem_alloc<kCols>(smem_result);
}
template <bool kPair, uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_t(uint32_t taddr) {
    if constexpr (kPair) tmem_dealloc_2sm<kCols>(taddr);
    else tmem_dealloc<kCols>(taddr);
}
// arrive on a barrier the LEADER waits on (single-CTA form: the CTA's own barrier)
template <bool kPair>
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    if constexpr (kPair) mbar_arrive_remote(bar, 0u);
    else mbar_arrive(bar);
}
template <bool kPair>
__device__ __forceinline__ void tma_load_3d_t(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                              int c1, int c2) {
    if constexpr (kPair) tma_load_3d_2sm(smem_dst, map, bar, c0, c1, c2);
    else tma_load_3d(smem_dst, map, bar, c0, c1, c2);
}
template <bool kPair>
__device__ __forceinline__ void cluster_sync_if() {
    if constexpr (kPair) cluster_sync();
}

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle (rows of 128 B, 8-row
// groups 1024 B apart).  Bit layout: [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4,
// [46,48) version=1 (sm_100), [61,64) layout type (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;               // LBO (ignored for swizzled K-major)
    d |= (uint64_t)(1024u >> 4) << 32;    // SBO: next 8-row group
    d |= (uint64_t)1 << 46;               // descriptor version
    d |= (uint64_t)2 << 61;               // SWIZZLE_128B
    return d;
}
// Instruction descriptor (upper 32 bits of CUTLASS' idescE): [4,6) D fmt, [7,10) A fmt,
// [10,13) B fmt, [15] A major (0=K), [16] B major, [17,23) N>>3, [24,29) M>>4.
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t d_fmt, uint32_t ab_fmt, uint32_t M,
                                                  uint32_t N) {
    return (d_fmt << 4) | (ab_fmt << 7) | (ab_fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// Block-scaled descriptor (CUTLASS InstrDescriptorBlockScaled): [4,6) B scale-factor id, [7,10) A format,
// [10,13) B format (E2M1 = 1), [17,23) N>>3, [23] scale format (1 = UE8M0), [24,29) M>>4, [29,31) A scale-factor
// id, [31] K size (0 = K64 dense).
__host__ __device__ constexpr uint32_t umma_idesc_mxf4(uint32_t M, uint32_t N) {
    return (1u << 7) | (1u << 10) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24);
}

}  // namespace ptx
}  // namespace bnn
