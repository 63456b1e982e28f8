// Batch-norm finalize kernels fused with their all-reduce over NVLink peer memory (data-parallel path).
//
// Per layer and direction the ranks must sum a [2, N] fp64 vector (forward: sum z, sum z^2; backward:
// sum delta, sum delta*(z-mu)) before a per-column finalize that every rank repeats identically.  The
// vectors are 64 KB: an NCCL all-reduce is pure launch + rendezvous latency (~45 us at 8 ranks, eight of
// them per step).  Here the producer kernel (colstats / bn_bwd_reduce) writes its partial sums straight
// into this rank's slot of a SYMMETRIC buffer (same offset on every rank, mapped into every peer over
// NVLink/NVSwitch), and the finalize kernel itself
//   1. signals "my slot is written" to every peer (st.release.sys into the peer's flag pad),
//   2. waits until every peer's flag for this epoch has arrived in its own pad (ld.acquire.sys, bounded),
//   3. reads all ranks' slots through the peer pointers in rank order (so every rank forms bit-identical
//      totals) and runs the finalize arithmetic on the totals.
// Slots and flags are double-buffered on the parity of a call counter that all ranks advance in lock
// step: a rank can run at most one call ahead of its slowest peer, so two buffers suffice.
#include "common.cuh"

namespace bnn {

namespace {

constexpr int kMaxRanks = 16;

struct PeerTable {
    const double* data[kMaxRanks];  // every rank's slot for this call (device-mapped peer pointers)
    uint32_t* flags[kMaxRanks];     // every rank's flag pad: uint32[2][kMaxRanks]
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_peer_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Cross-rank barrier executed by every block: block 0 publishes this rank's arrival to all peers, every
// block waits on the local pad.  Bounded: a missing peer is a trapped launch, not a hung GPU.
__device__ void peer_barrier(const PeerTable& t, int rank, int world, uint32_t epoch) {
    const int parity = epoch & 1u;
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(t.flags[threadIdx.x] + parity * kMaxRanks + rank, epoch);
    }
    if (threadIdx.x < world) {
        const uint32_t* f = t.flags[rank] + parity * kMaxRanks + threadIdx.x;
        unsigned long long t0 = 0;
        uint32_t spins = 0;
        while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
            if ((++spins & 0xfffu) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                if (now - t0 > 10000000000ull) {
                    printf("bnn_b200: peer barrier timed out (rank %d waiting for rank %d, epoch %u)\n",
                           rank, (int)threadIdx.x, epoch);
                    __trap();
                }
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(128)
bn_stats_finalize_p2p_kernel(const PeerTable t, int rank, int world, uint32_t epoch, int N, double count,
                             float* __restrict__ mu, float* __restrict__ var,
                             double* __restrict__ sums_total) {
    peer_barrier(t, rank, world, epoch);
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double s = 0.0, q = 0.0;
    for (int r = 0; r < world; ++r) {  // fixed order: identical totals on every rank
        s += ld_peer_f64(t.data[r] + n);
        q += ld_peer_f64(t.data[r] + N + n);
    }
    sums_total[n] = s;
    sums_total[N + n] = q;
    const double m = s / count;
    double v = q / count - m * m;
    if (v < 0) v = 0;
    mu[n] = (float)m;
    var[n] = (float)v;
}

__global__ void __launch_bounds__(128)
allreduce_f64_p2p_kernel(const PeerTable t, int rank, int world, uint32_t epoch, int n_total,
                         double* __restrict__ out) {
    peer_barrier(t, rank, world, epoch);
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_total) return;
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += ld_peer_f64(t.data[r] + n);
    out[n] = s;
}

}  // namespace
}  // namespace bnn

using namespace bnn;

static int fill_table(PeerTable& t, const uint64_t* peer_data_host, const uint64_t* peer_flags_host,
                      int world) {
    BNN_CHECK_ARG(world >= 1 && world <= kMaxRanks, "p2p: world %d out of range (max %d)", world,
                  kMaxRanks);
    for (int r = 0; r < kMaxRanks; ++r) {
        t.data[r] = r < world ? reinterpret_cast<const double*>(peer_data_host[r]) : nullptr;
        t.flags[r] = r < world ? reinterpret_cast<uint32_t*>(peer_flags_host[r]) : nullptr;
    }
    return BNN_OK;
}

extern "C" int bnn_bn_stats_finalize_p2p(const uint64_t* peer_data_host, const uint64_t* peer_flags_host,
                                         int rank, int world, uint32_t epoch, int N, double count,
                                         float* mu, float* var, double* sums_total,
                                         bnn_stream_t stream) {
    BNN_CHECK_ARG(peer_data_host && peer_flags_host && mu && var && sums_total && N > 0 && count > 0 &&
                      rank >= 0 && rank < world && epoch > 0,
                  "bnn_bn_stats_finalize_p2p: bad input");
    PeerTable t;
    int rc = fill_table(t, peer_data_host, peer_flags_host, world);
    if (rc) return rc;
    bn_stats_finalize_p2p_kernel<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(
        t, rank, world, epoch, N, count, mu, var, sums_total);
    BNN_LAUNCH_CHECK("bn_stats_finalize_p2p_kernel");
    return BNN_OK;
}

extern "C" int bnn_allreduce_f64_p2p(const uint64_t* peer_data_host, const uint64_t* peer_flags_host,
                                     int rank, int world, uint32_t epoch, int n, double* out,
                                     bnn_stream_t stream) {
    BNN_CHECK_ARG(peer_data_host && peer_flags_host && out && n > 0 && rank >= 0 && rank < world &&
                      epoch > 0,
                  "bnn_allreduce_f64_p2p: bad input");
    PeerTable t;
    int rc = fill_table(t, peer_data_host, peer_flags_host, world);
    if (rc) return rc;
    allreduce_f64_p2p_kernel<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(t, rank, world, epoch, n,
                                                                            out);
    BNN_LAUNCH_CHECK("allreduce_f64_p2p_kernel");
    return BNN_OK;
}
