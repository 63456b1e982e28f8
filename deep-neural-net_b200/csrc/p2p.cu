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
// Every exchange call site (layer x direction) has its own slot and its own flag row, and the epoch is
// the training-step counter that all ranks advance in lock step.  The counter lives in device memory
// (bnn_step_tick increments it) and the kernels read it through a pointer, so a whole training step can
// be captured into a CUDA graph and replayed.  A rank cannot overwrite a slot that a peer is still
// reading: between two uses of one call site every rank passes another site's barrier.
#include "common.cuh"

namespace bnn {

namespace {

constexpr int kMaxRanks = 16;

struct PeerTable {
    const double* data[kMaxRanks];  // every rank's slot for this call (device-mapped peer pointers)
    uint32_t* flags[kMaxRanks];     // every rank's flag row for this call site: uint32[kMaxRanks]
};

// epoch operand of every exchange kernel: the device-resident step counter when given, else immediate
__device__ __forceinline__ uint32_t resolve_epoch(const uint32_t* epoch_ptr, uint32_t epoch_imm) {
    return epoch_ptr ? *reinterpret_cast<const volatile uint32_t*>(epoch_ptr) : epoch_imm;
}

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
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(t.flags[threadIdx.x] + rank, epoch);
    }
    if (threadIdx.x < world) {
        const uint32_t* f = t.flags[rank] + threadIdx.x;
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
bn_stats_finalize_p2p_kernel(const PeerTable t, int rank, int world, uint32_t epoch_imm,
                             const uint32_t* __restrict__ epoch_ptr, int N, double count,
                             float* __restrict__ mu, float* __restrict__ var,
                             double* __restrict__ sums_total) {
    peer_barrier(t, rank, world, resolve_epoch(epoch_ptr, epoch_imm));
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
allreduce_f64_p2p_kernel(const PeerTable t, int rank, int world, uint32_t epoch_imm,
                         const uint32_t* __restrict__ epoch_ptr, int n_total, double* __restrict__ out) {
    peer_barrier(t, rank, world, resolve_epoch(epoch_ptr, epoch_imm));
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_total) return;
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += ld_peer_f64(t.data[r] + n);
    out[n] = s;
}

// ---------------------------------------------------------------------------------------------
// Data-parallel dW: the all-gather half.  The dW GEMM (gemm_tc.cu, BNN_EPI_SCATTER_F32) has pushed
// every rank's partial rows into the owner's staging buffer; the owner sums them in rank order,
// applies W -= lr * dW (layers.py:268) and stores the new rows into every rank's W.
// Flags are uint32[n_slots][kMaxRanks] of monotonically increasing epochs.
// ---------------------------------------------------------------------------------------------
struct FlagTable {
    uint32_t* flags[kMaxRanks];
};
struct WTable {
    float* w[kMaxRanks];
};

__device__ void spin_until(const uint32_t* f, uint32_t epoch, int rank, int peer, const char* what) {
    unsigned long long t0 = 0;
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
        if ((++spins & 0xfffu) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            if (now - t0 > 10000000000ull) {
                printf("bnn_b200: %s timed out (rank %d waiting for rank %d, epoch %u, saw %u)\n", what,
                       rank, peer, epoch, ld_acquire_sys(f));
                __trap();
            }
        }
    }
}

__global__ void __launch_bounds__(32)
peer_signal_kernel(const FlagTable t, int slot, int rank, int world, uint32_t epoch_imm,
                   const uint32_t* __restrict__ epoch_ptr) {
    const uint32_t epoch = resolve_epoch(epoch_ptr, epoch_imm);
    // every earlier kernel of this stream has completed, so its peer stores are performed; the fence
    // orders them before the flag for observers on other GPUs
    if (threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(t.flags[threadIdx.x] + slot * kMaxRanks + rank, epoch);
    }
}

// waits on n_slots flag rows: slot, slot + slot_stride, ...
__global__ void __launch_bounds__(32)
peer_wait_kernel(const uint32_t* __restrict__ flags_local, int slot, int slot_stride, int n_slots,
                 int rank, int world, uint32_t epoch_imm, const uint32_t* __restrict__ epoch_ptr) {
    const uint32_t epoch = resolve_epoch(epoch_ptr, epoch_imm);
    if (threadIdx.x < world)
        for (int i = 0; i < n_slots; ++i)
            spin_until(flags_local + (slot + i * slot_stride) * kMaxRanks + threadIdx.x, epoch, rank,
                       threadIdx.x, "peer wait");
}

__global__ void step_tick_kernel(uint32_t* __restrict__ step) { *step += 1u; }

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 v;
    // L2-only load: the staging rows were written by peers while this kernel may already have been
    // resident, so nothing may be served from a non-coherent path
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <int VEC>
struct Pack;
template <>
struct Pack<4> {
    float v[4];
    __device__ static Pack load_stream(const float* p) {
        const float4 f = ld_stream_f4(reinterpret_cast<const float4*>(p));
        return Pack{{f.x, f.y, f.z, f.w}};
    }
    __device__ static Pack load(const float* p) {
        const float4 f = *reinterpret_cast<const float4*>(p);
        return Pack{{f.x, f.y, f.z, f.w}};
    }
    __device__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <>
struct Pack<1> {
    float v[1];
    __device__ static Pack load_stream(const float* p) {
        float f;
        asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(f) : "l"(p));
        return Pack{{f}};
    }
    __device__ static Pack load(const float* p) { return Pack{{*p}}; }
    __device__ void store(float* p) const { *p = v[0]; }
};

template <int WORLD, int VEC>  // WORLD 0 = run-time world; VEC floats per thread and access
__global__ void __launch_bounds__(256)
dw_reduce_sgd_p2p_kernel(const float* __restrict__ stage, const WTable wt, const FlagTable ft,
                         int slot_pushed, int slot_landed, int rank, int world_rt, uint32_t epoch_imm,
                         const uint32_t* __restrict__ epoch_ptr, int rows, int N, int rows_per_owner,
                         float lr, int broadcast, uint32_t* __restrict__ counter,
                         uint32_t* __restrict__ epoch_word) {
    const int world = WORLD ? WORLD : world_rt;
    const uint32_t epoch = resolve_epoch(epoch_ptr, epoch_imm);
    if (epoch_word && blockIdx.x == 0 && threadIdx.x == 0) *epoch_word = epoch;
    if (broadcast & 2) {
        // "my partial rows are pushed" (the dW GEMM is the previous kernel of this stream, so its peer
        // stores are performed): block 0 tells every rank, which saves a separate signal launch
        if (blockIdx.x == 0 && threadIdx.x < world) {
            __threadfence_system();
            st_release_sys(ft.flags[threadIdx.x] + slot_pushed * kMaxRanks + rank, epoch);
        }
        broadcast &= 1;
    }
    // wait for every rank's partial rows of this epoch (bounded)
    if (threadIdx.x < world)
        spin_until(ft.flags[rank] + slot_pushed * kMaxRanks + threadIdx.x, epoch, rank, threadIdx.x,
                   "dW reduce");
    __syncthreads();
    const long long first = (long long)rank * rows_per_owner * N;  // first owned element of W
    const long long slice = (long long)rows_per_owner * N;  // one source rank's slice of the staging
    const long long total = (long long)rows * N;
    const long long stride = (long long)gridDim.x * blockDim.x * VEC;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < total; i += stride) {
        Pack<VEC> acc = Pack<VEC>::load_stream(stage + i);
#pragma unroll 4
        for (int r = 1; r < world; ++r) {  // fixed rank order: the owner alone forms the sum
            const Pack<VEC> v = Pack<VEC>::load_stream(stage + r * slice + i);
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc.v[j] = __fadd_rn(acc.v[j], v.v[j]);
        }
        Pack<VEC> w = Pack<VEC>::load(wt.w[rank] + first + i);
#pragma unroll
        for (int j = 0; j < VEC; ++j) w.v[j] = __fsub_rn(w.v[j], __fmul_rn(lr, acc.v[j]));
        if (broadcast) {
#pragma unroll 4
            for (int r = 0; r < world; ++r) {
                int dst = rank + r;  // start with the local copy, then walk the peers
                if (dst >= world) dst -= world;
                w.store(wt.w[dst] + first + i);
            }
        } else {
            w.store(wt.w[rank] + first + i);
        }
    }
    if (!broadcast) return;
    // the last block to finish publishes "my rows have landed" to every rank
    __threadfence_system();
    __syncthreads();
    __shared__ uint32_t last;
    if (threadIdx.x == 0) last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(ft.flags[threadIdx.x] + slot_landed * kMaxRanks + rank, epoch);
    }
}

}  // namespace
}  // namespace bnn

using namespace bnn;

static int fill_flags(FlagTable& t, const uint64_t* peer_flags_host, int world) {
    BNN_CHECK_ARG(world >= 1 && world <= kMaxRanks, "p2p: world %d out of range (max %d)", world,
                  kMaxRanks);
    for (int r = 0; r < kMaxRanks; ++r)
        t.flags[r] = r < world ? reinterpret_cast<uint32_t*>(peer_flags_host[r]) : nullptr;
    return BNN_OK;
}

extern "C" int bnn_step_tick(uint32_t* step_dev, bnn_stream_t stream) {
    BNN_CHECK_ARG(step_dev, "bnn_step_tick: null counter");
    step_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
    BNN_LAUNCH_CHECK("step_tick_kernel");
    return BNN_OK;
}

extern "C" int bnn_peer_signal(const uint64_t* peer_flags_host, int slot, int rank, int world,
                               uint32_t epoch, const uint32_t* epoch_ptr, bnn_stream_t stream) {
    BNN_CHECK_ARG(peer_flags_host && slot >= 0 && rank >= 0 && rank < world && (epoch > 0 || epoch_ptr),
                  "bnn_peer_signal: bad input");
    FlagTable t;
    int rc = fill_flags(t, peer_flags_host, world);
    if (rc) return rc;
    peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(t, slot, rank, world, epoch, epoch_ptr);
    BNN_LAUNCH_CHECK("peer_signal_kernel");
    return BNN_OK;
}

extern "C" int bnn_peer_wait(const uint32_t* flags_local, int slot, int slot_stride, int n_slots,
                             int rank, int world, uint32_t epoch, const uint32_t* epoch_ptr,
                             bnn_stream_t stream) {
    BNN_CHECK_ARG(flags_local && slot >= 0 && n_slots >= 1 && rank >= 0 && rank < world &&
                      world <= kMaxRanks && (epoch > 0 || epoch_ptr),
                  "bnn_peer_wait: bad input");
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags_local, slot, slot_stride, n_slots, rank,
                                                         world, epoch, epoch_ptr);
    BNN_LAUNCH_CHECK("peer_wait_kernel");
    return BNN_OK;
}

extern "C" int bnn_dw_reduce_sgd_p2p(const float* stage_local, const uint64_t* peer_w_host,
                                     const uint64_t* peer_flags_host, int slot_pushed, int slot_landed,
                                     int rank, int world, uint32_t epoch, const uint32_t* epoch_ptr,
                                     int K, int N, int rows_per_owner, float lr, int broadcast,
                                     uint32_t* counter, uint32_t* epoch_word, bnn_stream_t stream) {
    BNN_CHECK_ARG(stage_local && peer_w_host && peer_flags_host && counter && K > 0 && N > 0 &&
                      rank >= 0 && rank < world && (epoch > 0 || epoch_ptr) && slot_pushed >= 0 &&
                      slot_landed >= 0,
                  "bnn_dw_reduce_sgd_p2p: bad input");
    BNN_CHECK_ARG(rows_per_owner > 0 && (long long)rows_per_owner * world >= K,
                  "bnn_dw_reduce_sgd_p2p: rows_per_owner %d does not cover K=%d", rows_per_owner, K);
    FlagTable ft;
    int rc = fill_flags(ft, peer_flags_host, world);
    if (rc) return rc;
    WTable wt;
    bool vec = aligned16(stage_local) && ((long long)rows_per_owner * N) % 4 == 0;
    for (int r = 0; r < kMaxRanks; ++r)
        wt.w[r] = r < world ? reinterpret_cast<float*>(peer_w_host[r]) : nullptr;
    for (int r = 0; r < world; ++r) {
        BNN_CHECK_ARG(wt.w[r], "bnn_dw_reduce_sgd_p2p: W of rank %d is null", r);
        vec = vec && aligned16(wt.w[r]);
    }
    // rows this rank owns (the last owners of a short matrix may own fewer, or none)
    const long long first = (long long)rank * rows_per_owner;
    const int rows = (int)(first >= K ? 0 : (K - first < rows_per_owner ? K - first : rows_per_owner));
    vec = vec && ((long long)rows * N) % 4 == 0;
    const long long items = ((long long)rows * N + (vec ? 3 : 0)) / (vec ? 4 : 1);
    const long long want = (items + 255) / 256;
    const int cap = sm_count() * 8;
    const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    cudaStream_t s = (cudaStream_t)stream;
#define BNN_DW_LAUNCH(W, V)                                                                          \
    dw_reduce_sgd_p2p_kernel<W, V><<<grid, 256, 0, s>>>(stage_local, wt, ft, slot_pushed, slot_landed, \
                                                        rank, world, epoch, epoch_ptr, rows, N,        \
                                                        rows_per_owner,                               \
                                                        lr, broadcast, counter, epoch_word)
    if (vec) {
        switch (world) {
            case 2: BNN_DW_LAUNCH(2, 4); break;
            case 4: BNN_DW_LAUNCH(4, 4); break;
            case 8: BNN_DW_LAUNCH(8, 4); break;
            default: BNN_DW_LAUNCH(0, 4); break;
        }
    } else {
        BNN_DW_LAUNCH(0, 1);
    }
#undef BNN_DW_LAUNCH
    BNN_LAUNCH_CHECK("dw_reduce_sgd_p2p_kernel");
    return BNN_OK;
}

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
                                         int rank, int world, uint32_t epoch, const uint32_t* epoch_ptr,
                                         int N, double count, float* mu, float* var, double* sums_total,
                                         bnn_stream_t stream) {
    BNN_CHECK_ARG(peer_data_host && peer_flags_host && mu && var && sums_total && N > 0 && count > 0 &&
                      rank >= 0 && rank < world && (epoch > 0 || epoch_ptr),
                  "bnn_bn_stats_finalize_p2p: bad input");
    PeerTable t;
    int rc = fill_table(t, peer_data_host, peer_flags_host, world);
    if (rc) return rc;
    bn_stats_finalize_p2p_kernel<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(
        t, rank, world, epoch, epoch_ptr, N, count, mu, var, sums_total);
    BNN_LAUNCH_CHECK("bn_stats_finalize_p2p_kernel");
    return BNN_OK;
}

extern "C" int bnn_allreduce_f64_p2p(const uint64_t* peer_data_host, const uint64_t* peer_flags_host,
                                     int rank, int world, uint32_t epoch, const uint32_t* epoch_ptr, int n,
                                     double* out, bnn_stream_t stream) {
    BNN_CHECK_ARG(peer_data_host && peer_flags_host && out && n > 0 && rank >= 0 && rank < world &&
                      (epoch > 0 || epoch_ptr),
                  "bnn_allreduce_f64_p2p: bad input");
    PeerTable t;
    int rc = fill_table(t, peer_data_host, peer_flags_host, world);
    if (rc) return rc;
    allreduce_f64_p2p_kernel<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(t, rank, world, epoch,
                                                                            epoch_ptr, n, out);
    BNN_LAUNCH_CHECK("allreduce_f64_p2p_kernel");
    return BNN_OK;
}

extern "C" int bnn_peer_broadcast_rows(const uint64_t* peer_w_host, const uint64_t* peer_flags_host,
                                       int slot_landed, int rank, int world, int K, int N,
                                       int rows_per_owner, const uint32_t* epoch_word,
                                       bnn_stream_t stream) {
    BNN_CHECK_ARG(peer_w_host && peer_flags_host && epoch_word && K > 0 && N > 0 && rank >= 0 &&
                      rank < world && world <= kMaxRanks && slot_landed >= 0 && rows_per_owner > 0,
                  "bnn_peer_broadcast_rows: bad input");
    cudaStream_t s = (cudaStream_t)stream;
    const long long first = (long long)rank * rows_per_owner;
    const long long rows = first >= K ? 0 : (K - first < rows_per_owner ? K - first : rows_per_owner);
    const size_t off = (size_t)first * N * sizeof(float), bytes = (size_t)rows * N * sizeof(float);
    const char* src = reinterpret_cast<const char*>(peer_w_host[rank]) + off;
    for (int i = 1; i < world && bytes > 0; ++i) {
        const int p = (rank + i) % world;  // every rank starts with a different peer
        BNN_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(peer_w_host[p]) + off, src, bytes,
                                 cudaMemcpyDeviceToDevice, s));
    }
    for (int i = 0; i < world; ++i) {
        const int p = (rank + i) % world;
        uint32_t* flag = reinterpret_cast<uint32_t*>(peer_flags_host[p]) + slot_landed * kMaxRanks + rank;
        BNN_CUDA(cudaMemcpyAsync(flag, epoch_word, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    }
    return BNN_OK;
}
