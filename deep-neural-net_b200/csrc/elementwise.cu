// Output-layer activations and losses other than softmax + cross-entropy: the reference lets a
// binarized network end in any activation of mlpcode/activation.py and train against any loss of mlpcode/loss.py
// (network.py:236-262, 377-387).  All of them are element-wise or row-wise maps over [B, C] with C = the number of
// classes — HBM-trivial; what matters is the operation order, which follows the reference line by line.
#include "common.cuh"

namespace bnn {
namespace {

constexpr float kReluEpsilon = 0.01f;  // RELU_EPSILON, activation.py:6

// a = f(x)                                                          activation.py:8-121
__device__ __forceinline__ float act_fwd(int kind, float x) {
    switch (kind) {
        case BNN_ACT_SIGMOID: return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));           // :17-22
        case BNN_ACT_TANH: return tanhf(x);                                                 // :31-35
        case BNN_ACT_RELU: return x < 0.0f ? 0.0f : x;                                      // :65-69  a[x < 0] = 0
        case BNN_ACT_LEAKY_RELU: return x < 0.0f ? __fmul_rn(x, kReluEpsilon) : x;          // :80-84
        case BNN_ACT_HARD_SIGMOID: return fminf(fmaxf(__fdiv_rn(__fadd_rn(x, 1.0f), 2.0f), 0.0f), 1.0f);  // :95-99
        case BNN_ACT_HARD_TANH: return fminf(fmaxf(x, -1.0f), 1.0f);                         // :115-119
        default: return x;                                                                  // identity :8-9
    }
}

// delta = dA * f'(.) evaluated on the ACTIVATION a, as the reference does         activation.py:12-131
__device__ __forceinline__ float act_bwd(int kind, float dA, float a) {
    switch (kind) {
        case BNN_ACT_SIGMOID: return __fmul_rn(dA, __fmul_rn(a, __fsub_rn(1.0f, a)));       // :25-28
        case BNN_ACT_TANH: return __fmul_rn(dA, __fsub_rn(1.0f, __fmul_rn(a, a)));          // :38-41
        case BNN_ACT_RELU: return a <= 0.0f ? 0.0f : dA;                                    // :72-77
        case BNN_ACT_LEAKY_RELU: return a <= 0.0f ? __fmul_rn(dA, kReluEpsilon) : dA;       // :87-92
        case BNN_ACT_HARD_TANH: return (a < -1.0f || a > 1.0f) ? __fmul_rn(0.0f, dA) : dA;  // :122-131 (out *= dA)
        default: return dA;                                                                 // identity :12-14
    }
}

__global__ void __launch_bounds__(256)
activation_kernel(const float* __restrict__ x, long long n, int kind, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = act_fwd(kind, x[i]);
}

__global__ void __launch_bounds__(256)
activation_grad_kernel(const float* __restrict__ dA, const float* __restrict__ a, long long n, int kind,
                       float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = act_bwd(kind, dA[i], a[i]);
}

// softmax_derivative (activation.py:56-68): the reference applies softmax to the ALREADY soft-maxed activations
// again (a = softmax(a)) and contracts the Jacobian diag(s) - s s^T with dA:  delta_j = s_j * (dA_j - sum_k s_k dA_k).
// One warp per row, any C.
__global__ void __launch_bounds__(256)
softmax_grad_kernel(const float* __restrict__ dA, const float* __restrict__ a, int R, int C,
                    float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= R) return;
    const float* ar = a + (long long)row * C;
    const float* dr = dA + (long long)row * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, ar[c]);
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int c = lane; c < C; c += 32) sum += expf(ar[c] - mx);
    sum = warp_sum(sum);
    float dot = 0.0f;
    for (int c = lane; c < C; c += 32) dot = fmaf(__fdiv_rn(expf(ar[c] - mx), sum), dr[c], dot);
    dot = warp_sum(dot);
    for (int c = lane; c < C; c += 32) {
        const float s = __fdiv_rn(expf(ar[c] - mx), sum);
        out[(long long)row * C + c] = __fmul_rn(s, __fsub_rn(dr[c], dot));
    }
}

// Per-row losses, one thread per row (C = number of classes, small):                    loss.py:10-36, 71-119
//   CE     -sum_c y * log(clip(p, 1e-15, 1 - 1e-15))   (p is clipped IN PLACE like np.clip(..., out=ypred), :32;
//                                                       in fp32 the upper bound is 1.0)
//   MSE    0.5 * sum_c (p - y)^2                                                                     :87-89
//   HINGE  sum_c max(0, 1 - p * y)                      (labels in {-1, +1}: network.py:301-303)     :112-119
__global__ void __launch_bounds__(128)
loss_rows_kernel(float* __restrict__ ypred, const int8_t* __restrict__ y, int R, int C, int kind,
                 float* __restrict__ rows, double* __restrict__ loss_sum) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    float L = 0.0f;
    if (r < R) {
        for (int c = 0; c < C; ++c) {
            const long long i = (long long)r * C + c;
            const float yy = (float)y[i];
            float p = ypred[i];
            if (kind == BNN_LOSS_CROSS_ENTROPY) {
                p = fminf(fmaxf(p, 1e-15f), 1.0f);
                ypred[i] = p;
                L -= yy * logf(p);
            } else if (kind == BNN_LOSS_MSE) {
                const float d = __fsub_rn(p, yy);
                L += __fmul_rn(d, d);
            } else {
                L += fmaxf(0.0f, __fsub_rn(1.0f, __fmul_rn(p, yy)));
            }
        }
        if (kind == BNN_LOSS_MSE) L = __fmul_rn(0.5f, L);
        if (rows) rows[r] = L;
    }
    if (loss_sum) {
        double s = warp_sum((double)L);
        if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(loss_sum, s);
    }
}

// dL/dA, element-wise:                                                              loss.py:39-68, 92-131
//   CE     -(y / p - (1 - y) / (1 - p)) / count  on the clipped p (clipped in place, :62-63); p = 1 gives the
//          reference's own inf / nan for y = 0 — reproduced, not repaired
//   MSE    (p - y) / count                                                                         :107
//   HINGE  p * y < 1 ? -y / count : 0                                                              :127-129
__global__ void __launch_bounds__(256)
loss_grad_kernel(float* __restrict__ ypred, const int8_t* __restrict__ y, long long n, int kind, float count,
                 float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float yy = (float)y[i];
        float p = ypred[i];
        float g;
        if (kind == BNN_LOSS_CROSS_ENTROPY) {
            p = fminf(fmaxf(p, 1e-15f), 1.0f);
            ypred[i] = p;
            g = -__fsub_rn(__fdiv_rn(yy, p), __fdiv_rn(__fsub_rn(1.0f, yy), __fsub_rn(1.0f, p)));
            g = __fdiv_rn(g, count);
        } else if (kind == BNN_LOSS_MSE) {
            g = __fdiv_rn(__fsub_rn(p, yy), count);
        } else {
            g = __fmul_rn(p, yy) < 1.0f ? __fdiv_rn(-yy, count) : 0.0f;
        }
        out[i] = g;
    }
}

int ew_grid(long long n) {
    const long long blocks = (n + 255) / 256;
    return (int)(blocks < 148 * 8 ? (blocks > 0 ? blocks : 1) : 148 * 8);
}

}  // namespace
}  // namespace bnn

using namespace bnn;

extern "C" int bnn_activation(const float* x, int64_t n, int kind, float* out, bnn_stream_t stream) {
    BNN_CHECK_ARG(x && out && n > 0, "bnn_activation: bad input");
    BNN_CHECK_ARG(kind >= BNN_ACT_IDENTITY && kind <= BNN_ACT_HARD_TANH, "bnn_activation: bad kind %d", kind);
    activation_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, n, kind, out);
    BNN_LAUNCH_CHECK("activation_kernel");
    return BNN_OK;
}

extern "C" int bnn_activation_grad(const float* dA, const float* a, int64_t n, int kind, float* out,
                                   bnn_stream_t stream) {
    BNN_CHECK_ARG(dA && a && out && n > 0, "bnn_activation_grad: bad input");
    BNN_CHECK_ARG(kind >= BNN_ACT_IDENTITY && kind <= BNN_ACT_HARD_TANH && kind != BNN_ACT_HARD_SIGMOID,
                  "bnn_activation_grad: kind %d has no derivative in the reference", kind);
    activation_grad_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(dA, a, n, kind, out);
    BNN_LAUNCH_CHECK("activation_grad_kernel");
    return BNN_OK;
}

extern "C" int bnn_softmax_grad(const float* dA, const float* a, int R, int C, float* out, bnn_stream_t stream) {
    BNN_CHECK_ARG(dA && a && out && R > 0 && C > 0, "bnn_softmax_grad: bad input");
    softmax_grad_kernel<<<cdiv(R, 8), 256, 0, (cudaStream_t)stream>>>(dA, a, R, C, out);
    BNN_LAUNCH_CHECK("softmax_grad_kernel");
    return BNN_OK;
}

extern "C" int bnn_loss_rows(float* ypred, const int8_t* y, int R, int C, int kind, float* rows, double* loss_sum,
                             bnn_stream_t stream) {
    BNN_CHECK_ARG(ypred && y && R > 0 && C > 0 && (rows || loss_sum), "bnn_loss_rows: bad input");
    BNN_CHECK_ARG(kind >= BNN_LOSS_CROSS_ENTROPY && kind <= BNN_LOSS_HINGE, "bnn_loss_rows: bad kind %d", kind);
    cudaStream_t s = (cudaStream_t)stream;
    if (loss_sum) BNN_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double), s));
    loss_rows_kernel<<<cdiv(R, 128), 128, 0, s>>>(ypred, y, R, C, kind, rows, loss_sum);
    BNN_LAUNCH_CHECK("loss_rows_kernel");
    return BNN_OK;
}

extern "C" int bnn_loss_grad(float* ypred, const int8_t* y, int R, int C, int kind, double count, float* out,
                             bnn_stream_t stream) {
    BNN_CHECK_ARG(ypred && y && out && R > 0 && C > 0 && count > 0, "bnn_loss_grad: bad input");
    BNN_CHECK_ARG(kind >= BNN_LOSS_CROSS_ENTROPY && kind <= BNN_LOSS_HINGE, "bnn_loss_grad: bad kind %d", kind);
    const long long n = (long long)R * C;
    loss_grad_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(ypred, y, n, kind, (float)count, out);
    BNN_LAUNCH_CHECK("loss_grad_kernel");
    return BNN_OK;
}
