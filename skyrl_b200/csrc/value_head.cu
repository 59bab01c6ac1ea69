// Critic value head: values[r] = <hidden[r, :], w> (+ bias), forward and backward.  HBM-bound GEMV:
// 2*H bytes per token forward, 4*H backward (read hidden for d_w, write d_hidden).
//
// Reference: the critic's `value_head = nn.Linear(hidden_size, 1)` applied to the trunk's last hidden state,
// skyrl-train/skyrl_train/model_wrapper.py:472-489 (values = value_head(last_hidden_states).squeeze(-1)).
#include <cuda_bf16.h>

#include "common.cuh"

namespace skyb200 {

// one warp per row, 128-bit loads of 8 bf16
__global__ void __launch_bounds__(256) value_fwd_kernel(const __nv_bfloat16* __restrict__ hidden,
                                                        const float* __restrict__ w, float bias,
                                                        float* __restrict__ values, int64_t N, int H) {
    const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= N) return;
    const __nv_bfloat16* h = hidden + row * H;
    float acc = 0.f;
    const int nv = H / 8;
    for (int v = lane; v < nv; v += 32) {
        const uint4 q = *reinterpret_cast<const uint4*>(h + v * 8);
        const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&q);
        const float4 w0 = *reinterpret_cast<const float4*>(w + v * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(w + v * 8 + 4);
        const float2 a = __bfloat1622float2(p2[0]), b = __bfloat1622float2(p2[1]), c = __bfloat1622float2(p2[2]),
                     d = __bfloat1622float2(p2[3]);
        acc = fmaf(a.x, w0.x, acc);
        acc = fmaf(a.y, w0.y, acc);
        acc = fmaf(b.x, w0.z, acc);
        acc = fmaf(b.y, w0.w, acc);
        acc = fmaf(c.x, w1.x, acc);
        acc = fmaf(c.y, w1.y, acc);
        acc = fmaf(d.x, w1.z, acc);
        acc = fmaf(d.y, w1.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) values[row] = acc + bias;
}

// d_hidden[r, :] = g[r] * w   (bf16 out)
__global__ void __launch_bounds__(256) value_bwd_hidden_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                               __nv_bfloat16* __restrict__ d_hidden, int64_t N, int H) {
    const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= N) return;
    const float gr = g[row];
    __nv_bfloat16* o = d_hidden + row * H;
    const int nv = H / 8;
    for (int v = lane; v < nv; v += 32) {
        const float4 w0 = *reinterpret_cast<const float4*>(w + v * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(w + v * 8 + 4);
        uint4 q;
        __nv_bfloat162 t;
        t = __floats2bfloat162_rn(gr * w0.x, gr * w0.y);
        q.x = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(gr * w0.z, gr * w0.w);
        q.y = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(gr * w1.x, gr * w1.y);
        q.z = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(gr * w1.z, gr * w1.w);
        q.w = *reinterpret_cast<uint32_t*>(&t);
        *reinterpret_cast<uint4*>(o + v * 8) = q;
    }
}

constexpr int kVhRows = 512;  // rows folded by one block of the d_w reduction

// partial[blk, c] = sum_{r in block} g[r] * hidden[r, c]; thread = 8 consecutive columns (blockDim.x * 8 >= H)
__global__ void __launch_bounds__(512) value_bwd_w_partial_kernel(const __nv_bfloat16* __restrict__ hidden,
                                                                  const float* __restrict__ g,
                                                                  float* __restrict__ partial, int64_t N, int H) {
    const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
    if (c0 >= H) return;
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * kVhRows;
    const int64_t r1 = min(N, r0 + kVhRows);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t r = r0; r < r1; ++r) {
        const float gr = __ldg(g + r);
        const uint4 q = *reinterpret_cast<const uint4*>(hidden + r * H + c0);
        const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 a = __bfloat1622float2(p2[k]);
            acc[2 * k] = fmaf(gr, a.x, acc[2 * k]);
            acc[2 * k + 1] = fmaf(gr, a.y, acc[2 * k + 1]);
        }
    }
    float* o = partial + static_cast<int64_t>(blockIdx.x) * H + c0;
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = acc[k];
}
// d_w[c] (+)= sum_blk partial[blk, c]  (fixed order -> deterministic); d_bias (+)= sum g
__global__ void __launch_bounds__(256) value_bwd_w_fold_kernel(const float* __restrict__ partial, int nblk, int H,
                                                               float* __restrict__ d_w, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= H) return;
    float s = 0.f;
    for (int b = 0; b < nblk; ++b) s += partial[static_cast<int64_t>(b) * H + c];
    d_w[c] = accumulate ? d_w[c] + s : s;
}

}  // namespace skyb200

using namespace skyb200;

extern "C" int skyb200_value_head_forward(const void* hidden, const float* weight, float bias, float* values,
                                          int64_t N, int64_t H, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (N == 0) return SKYB200_OK;
    SKYB200_REQUIRE(hidden && weight && values, "null pointer");
    SKYB200_REQUIRE(H % 8 == 0 && H > 0, "H must be a multiple of 8");
    value_fwd_kernel<<<(unsigned)((N + 7) / 8), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(hidden), weight,
                                                                  bias, values, N, (int)H);
    return check_launch("value_fwd_kernel");
}

extern "C" size_t skyb200_value_head_backward_workspace(int64_t N, int64_t H) {
    return (size_t)(((N + kVhRows - 1) / kVhRows) * H * sizeof(float) + 256);
}

extern "C" int skyb200_value_head_backward(const void* hidden, const float* weight, const float* g_values,
                                           void* d_hidden, float* d_weight, int32_t accumulate, void* workspace,
                                           size_t workspace_bytes, int64_t N, int64_t H, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SKYB200_REQUIRE(hidden && weight && g_values && workspace, "null pointer");
    SKYB200_REQUIRE(H % 8 == 0 && H > 0, "H must be a multiple of 8");
    if (workspace_bytes < skyb200_value_head_backward_workspace(N, H))
        return fail(SKYB200_EWORKSPACE, "value_head_backward workspace too small");
    if (N == 0) {
        if (d_weight && !accumulate) SKYB200_CUDA_CHECK(cudaMemsetAsync(d_weight, 0, (size_t)H * 4, stream));
        return SKYB200_OK;
    }
    int rc;
    if (d_hidden != nullptr) {
        value_bwd_hidden_kernel<<<(unsigned)((N + 7) / 8), 256, 0, stream>>>(
            g_values, weight, static_cast<__nv_bfloat16*>(d_hidden), N, (int)H);
        rc = check_launch("value_bwd_hidden_kernel");
        if (rc) return rc;
    }
    if (d_weight != nullptr) {
        float* partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
        const int nblk = (int)((N + kVhRows - 1) / kVhRows);
        const int threads = 128;
        const dim3 grid((unsigned)nblk, (unsigned)((H / 8 + threads - 1) / threads));
        value_bwd_w_partial_kernel<<<grid, threads, 0, stream>>>(static_cast<const __nv_bfloat16*>(hidden), g_values,
                                                                 partial, N, (int)H);
        rc = check_launch("value_bwd_w_partial_kernel");
        if (rc) return rc;
        value_bwd_w_fold_kernel<<<(unsigned)((H + 255) / 256), 256, 0, stream>>>(partial, nblk, (int)H, d_weight,
                                                                                accumulate);
        rc = check_launch("value_bwd_w_fold_kernel");
        if (rc) return rc;
    }
    return SKYB200_OK;
}
