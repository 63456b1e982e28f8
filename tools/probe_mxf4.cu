// Stand-alone probe (not part of libbnn_b200): can the +-1 x +-1 product run on the FP4 tensor pipe?
//
//   +-1 is exact in E2M1 (+1 = 0x2, -1 = 0xA).  tcgen05.mma kind::mxf4.block_scale multiplies E2M1 operands
//   by per-32-element UE8M0 scale factors held in TMEM; with every scale byte = 0x7F (2^0) the scale-factor
//   layout is irrelevant and D = sum_k a_k * b_k accumulates exact integers in fp32 (|D| <= K < 2^24).
//   Nominal rate 2x kind::i8, operand bytes 1/2.  DESIGN.md section 9 item 3.
//
// What it does, per SM (one CTA, 128 threads): fills a 128 x 128-byte A tile and a 256 x 128-byte B tile in
// shared memory with ROW-CONSTANT sign patterns (invariant under the 128-byte swizzle, so the operand layout
// cannot be got wrong), fills TMEM columns 256..511 with 0x7F7F7F7F, issues `iters` K-steps of 4 MMAs
// (M128 N256 K64 each: 256 fp4 elements per 128-byte row) into columns 0..255, and
//   (1) checks D[m, n] == sign_a(m) * sign_b(n) * 256 * iters for every element,
//   (2) reports MMA throughput next to the same loop with kind::i8 (+-1 bytes, K32 per MMA).
//
// Build + run on a B200 (written in round 1 without GPU time left; never executed):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/probe_mxf4.cu -o gpurun_out/probe_mxf4
//   gpurun_out/probe_mxf4 [iters=4096]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t addr) {  // K-major, 128-byte swizzle (as ptx.cuh)
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// kind::i8: [4,6) D fmt = 2 (s32), [7,10) A = 1, [10,13) B = 1 (signed), N>>3 at 17, M>>4 at 24
__host__ __device__ constexpr uint32_t idesc_i8(uint32_t M, uint32_t N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// block-scaled (CUTLASS InstrDescriptorBlockScaled): [4,6) B sf id, [7,10) A fmt, [10,13) B fmt (E2M1 = 1),
// [17,23) N>>3, [23] scale format (1 = UE8M0), [24,29) M>>4, [29,31) A sf id, [31] K size (0 = K64)
__host__ __device__ constexpr uint32_t idesc_mxf4(uint32_t M, uint32_t N) {
    return (1u << 7) | (1u << 10) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24);
}

template <bool kFp4>
__global__ void __launch_bounds__(128, 1) probe(int iters, int* bad, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                 // 128 rows x 128 B
    uint8_t* sb = smem + 128 * 128;     // 256 rows x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // row r of A has sign (+ if r % 3 != 0), row n of B has sign (+ if n % 5 != 0): row-constant fills
    const uint8_t pos = kFp4 ? 0x22 : 0x01, neg = kFp4 ? 0xAA : 0xFF;
    for (int i = tid; i < 128 * 128; i += 128) sa[i] = ((i / 128) % 3 != 0) ? pos : neg;
    for (int i = tid; i < 256 * 128; i += 128) sb[i] = ((i / 128) % 5 != 0) ? pos : neg;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy fills -> async-proxy reads
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    // unit scale factors: every byte of columns 256..511 = 0x7F (UE8M0 2^0); warp w owns lanes 32w..32w+31
    if (kFp4) {
        const uint32_t one = 0x7F7F7F7Fu;
        for (int c = 256; c < 512; ++c)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};"
                         ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c), "r"(one) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint64_t da = desc_k_sw128(smem_u32(sa)), db = desc_k_sw128(smem_u32(sb));
        const uint32_t idesc = kFp4 ? idesc_mxf4(128, 256) : idesc_i8(128, 256);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // 4 x 32 bytes of K per 128-byte row
                const uint64_t adv = (uint64_t)((j * 32) >> 4);
                const uint32_t acc = (it | j) ? 1u : 0u;
                if (kFp4)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n"
                                 ::"r"(tmem), "l"(da + adv), "l"(db + adv), "r"(idesc), "r"(acc),
                                 "r"(tmem + 256u), "r"(tmem + 384u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                                 ::"r"(tmem), "l"(da + adv), "l"(db + adv), "r"(idesc), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the MMAs (bounded: ~2 s of polling, then give up and report)
    uint32_t ok = 0;
    for (long long spin = 0; spin < (1ll << 28) && !ok; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    if (tid == 0) { t1 = clock64(); cycles[blockIdx.x] = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok) { if (lane == 0) atomicAdd(bad, 1 << 20); }
    else {
        // D[m, n]: thread = row m (lane of TMEM), 256 columns; per MMA K = 64 (fp4) or 32 (i8) elements
        const int m = tid;
        const int per_iter = kFp4 ? 256 : 128;
        const int sa_sign = (m % 3 != 0) ? 1 : -1;
        int wrong = 0;
        for (int c0 = 0; c0 < 256; c0 += 8) {
            uint32_t v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 8; ++j) {
                const int n = c0 + j;
                const long long want = (long long)sa_sign * ((n % 5 != 0) ? 1 : -1) * per_iter * iters;
                const long long got = kFp4 ? (long long)__uint_as_float(v[j]) : (long long)(int)v[j];
                wrong += got != want;
            }
        }
        if (wrong) atomicAdd(bad, wrong);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <bool kFp4>
int run(const char* name, int iters, int sms, double ghz) {
    int* bad;
    long long* cyc;
    CK(cudaMalloc(&bad, sizeof(int)));
    CK(cudaMalloc(&cyc, sizeof(long long) * sms));
    CK(cudaMemset(bad, 0, sizeof(int)));
    const int smem = 128 * 128 + 256 * 128 + 1024;
    CK(cudaFuncSetAttribute(probe<kFp4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    probe<kFp4><<<sms, 128, smem>>>(8, bad, cyc);  // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(bad, 0, sizeof(int)));
    CK(cudaEventRecord(e0));
    probe<kFp4><<<sms, 128, smem>>>(iters, bad, cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    int hbad = 0;
    long long hc = 0;
    CK(cudaMemcpy(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&hc, cyc, sizeof(long long), cudaMemcpyDeviceToHost));
    const double k_per_mma = kFp4 ? 64.0 : 32.0;
    const double ops = 2.0 * 128 * 256 * k_per_mma * 4.0 * iters * sms;
    printf("%-10s iters %d: %.3f ms, %.1f TOP/s (kernel incl. fill/check), SM0 MMA loop %lld clk = %.1f TOP/s at %.2f GHz, wrong elements: %d%s\n",
           name, iters, ms, ops / (ms * 1e-3) / 1e12, hc, ops / (hc / (ghz * 1e9)) / 1e12, ghz, hbad & 0xFFFFF,
           (hbad >> 20) ? "  [MMA completion timed out]" : "");
    return hbad != 0;
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 4096;   // 256 * 4096 = 2^20 < 2^24: exact in fp32
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const double ghz = p.clockRate * 1e-6;
    printf("%s, %d SMs, %.2f GHz\n", p.name, p.multiProcessorCount, ghz);
    int rc = run<false>("kind::i8", iters, p.multiProcessorCount, ghz);
    rc |= run<true>("kind::mxf4", iters, p.multiProcessorCount, ghz);
    return rc;
}
