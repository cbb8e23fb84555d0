// Softmax-feature-map linear attention core (between the qkv and proj 1x1 convs).
// Replaces the body of LinearAttention.forward, nn/modules/block.py:3364-3372:
//   k = softmax(k, over head_dim);  q = softmax(q, over N);  ctx = k^T v (64x64);  y = q ctx
// with channel index t*C + head*64 + j in qkv and head*64 + j in y.  The Q softmax is split as
// P = exp(q - max_N q), s = sum_N P, y = P (diag(1/s) ctx), so Q is never normalised in memory.
//
// This file holds the fp32 CUDA-core kernel (1e-5 contract; also the fallback for odd shapes).
// One CTA per (image, head); K/V/Q token tiles are staged through shared memory with a
// layout-aware coalesced loader, so NCHW and NHWC inputs run the same math.
#include <cstdlib>

#include "el_common.cuh"

namespace el {

constexpr int kD = 64;    // head_dim (block.py:3474: heads = c // 64)
constexpr int kTN = 64;   // tokens per tile
constexpr int kLd = kTN + 1;

struct AttnArgs {
    const void* qkv; int64_t qb, qc, qn;
    void* y; int64_t yb, yc, yn;
    int heads, N;
};

// load a [64 channels][kTN tokens] tile as fp32 into smem[ch][kLd]; `fill` for tokens >= N
template <typename T, bool CH_FAST>
__device__ __forceinline__ void load_tile(float* __restrict__ sm, const T* __restrict__ base, int64_t sc, int64_t sn, int n0, int N, float fill) {
    for (int idx = threadIdx.x; idx < kD * kTN; idx += blockDim.x) {
        int ch, n;
        if (CH_FAST) { n = idx >> 6; ch = idx & 63; } else { ch = idx >> 6; n = idx & 63; }
        int tok = n0 + n;
        sm[ch * kLd + n] = tok < N ? to_f(base[(int64_t)ch * sc + (int64_t)tok * sn]) : fill;
    }
}

template <typename T, bool CH_FAST>
__global__ void __launch_bounds__(256) linattn_simt_kernel(const __grid_constant__ AttnArgs A) {
    extern __shared__ float smem[];
    float* sK = smem;                 // [64][kLd]  K tile, later P tile
    float* sV = sK + kD * kLd;        // [64][kLd]  V tile, later y tile
    float* sQ = sV + kD * kLd;        // [64][kLd]  Q tile (statistics pass)
    float* sC = sQ + kD * kLd;        // [64][65]   ctx / s
    float* sMax = sC + kD * 65;       // [64] running max of q over tokens
    float* sSum = sMax + kD;          // [64] running sum of exp(q - max)

    const int tid = threadIdx.x;
    const int b = blockIdx.x / A.heads, head = blockIdx.x % A.heads;
    const int C = A.heads * kD, N = A.N;
    const T* qkv = reinterpret_cast<const T*>(A.qkv) + (int64_t)b * A.qb;
    const T* gq = qkv + (int64_t)(0 * C + head * kD) * A.qc;
    const T* gk = qkv + (int64_t)(1 * C + head * kD) * A.qc;
    const T* gv = qkv + (int64_t)(2 * C + head * kD) * A.qc;

    if (tid < kD) { sMax[tid] = -INFINITY; sSum[tid] = 0.f; }
    const int ti = tid >> 4, tj = tid & 15;  // 4x4 register block of the 64x64 context
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

    // ---- pass 1 over tokens: ctx += softmax_d(K)^T V, and the Q column statistics
    for (int n0 = 0; n0 < N; n0 += kTN) {
        __syncthreads();
        load_tile<T, CH_FAST>(sK, gk, A.qc, A.qn, n0, N, 0.f);
        load_tile<T, CH_FAST>(sV, gv, A.qc, A.qn, n0, N, 0.f);  // zero V kills the padded tokens
        load_tile<T, CH_FAST>(sQ, gq, A.qc, A.qn, n0, N, -INFINITY);
        __syncthreads();
        if (tid < kTN) {  // softmax over the 64 channels of one token (column tid)
            float m = -INFINITY;
            for (int d = 0; d < kD; ++d) m = fmaxf(m, sK[d * kLd + tid]);
            float s = 0.f;
            for (int d = 0; d < kD; ++d) { float e = expf(sK[d * kLd + tid] - m); sK[d * kLd + tid] = e; s += e; }
            float inv = 1.f / s;
            for (int d = 0; d < kD; ++d) sK[d * kLd + tid] *= inv;
        } else if (tid < kTN + kD) {  // online max / sum of one q row over this tile
            const int row = tid - kTN;
            float m_old = sMax[row], m = m_old;
            for (int n = 0; n < kTN; ++n) m = fmaxf(m, sQ[row * kLd + n]);
            float s = 0.f;
            for (int n = 0; n < kTN; ++n) s += expf(sQ[row * kLd + n] - m);
            sSum[row] = sSum[row] * (m_old == -INFINITY ? 0.f : expf(m_old - m)) + s;
            sMax[row] = m;
        }
        __syncthreads();
#pragma unroll 4
        for (int n = 0; n < kTN; ++n) {
            float a[4], v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = sK[(4 * ti + r) * kLd + n];
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = sV[(4 * tj + c) * kLd + n];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] += a[r] * v[c];
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float inv = 1.f / sSum[4 * ti + r];  // fold the Q softmax denominator into ctx rows
#pragma unroll
        for (int c = 0; c < 4; ++c) sC[(4 * ti + r) * 65 + 4 * tj + c] = acc[r][c] * inv;
    }

    // ---- pass 2 over tokens: y[j][n] = sum_i exp(q[i][n] - max_i) * ctx'[i][j]
    T* gy = reinterpret_cast<T*>(A.y) + (int64_t)b * A.yb + (int64_t)(head * kD) * A.yc;
    const int tjb = tid >> 4, tnb = tid & 15;  // 4 output channels x 4 tokens per thread
    for (int n0 = 0; n0 < N; n0 += kTN) {
        __syncthreads();
        load_tile<T, CH_FAST>(sK, gq, A.qc, A.qn, n0, N, -INFINITY);
        __syncthreads();
        for (int idx = tid; idx < kD * kTN; idx += blockDim.x) {
            int i = idx >> 6, n = idx & 63;
            sK[i * kLd + n] = expf(sK[i * kLd + n] - sMax[i]);
        }
        __syncthreads();
        float o[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) o[r][c] = 0.f;
#pragma unroll 4
        for (int i = 0; i < kD; ++i) {
            float cj[4], pn[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) cj[r] = sC[i * 65 + 4 * tjb + r];
#pragma unroll
            for (int c = 0; c < 4; ++c) pn[c] = sK[i * kLd + 4 * tnb + c];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) o[r][c] += cj[r] * pn[c];
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) sV[(4 * tjb + r) * kLd + 4 * tnb + c] = o[r][c];
        __syncthreads();
        for (int idx = tid; idx < kD * kTN; idx += blockDim.x) {  // coalesced store in the output's layout
            int ch, n;
            if (CH_FAST) { n = idx >> 6; ch = idx & 63; } else { ch = idx >> 6; n = idx & 63; }
            int tok = n0 + n;
            if (tok < N) gy[(int64_t)ch * A.yc + (int64_t)tok * A.yn] = from_f<T>(sV[ch * kLd + n]);
        }
    }
}

constexpr size_t kAttnSimtSmem = (size_t)(3 * kD * kLd + kD * 65 + 2 * kD) * sizeof(float);

// tcgen05 path, linattn_tc.cu
int linattn_tc_launch(const AttnArgs& A, int B, int dtype, cudaStream_t s);
bool linattn_tc_supported(const AttnArgs& A, int dtype);
// TMA-fed, chunk-parallel tcgen05 path for N <= 512 tokens, linattn_tma.cu
int linattn_tma_launch(const AttnArgs& A, int B, int dtype, cudaStream_t s);
bool linattn_tma_supported(const AttnArgs& A, int dtype);

}  // namespace el

using namespace el;

extern "C" int el_linattn_fwd(const void* qkv, const int64_t qs[3], void* y, const int64_t ys[3], int B, int heads, int N, int dtype, void* stream) {
    if (!qkv || !y || !qs || !ys || B <= 0 || heads <= 0 || N <= 0) return EL_ERR_ARG;
    AttnArgs A{qkv, qs[0], qs[1], qs[2], y, ys[0], ys[1], ys[2], heads, N};
    cudaStream_t s = (cudaStream_t)stream;
    static const bool force_simt = getenv("EL_LINATTN_SIMT") != nullptr;  // A/B switch for tests and profiling
    static const bool no_tma = getenv("EL_LINATTN_NO_TMA") != nullptr;     // A/B switch: the round-1 tcgen05 kernel for every N
    if (!force_simt && !no_tma && linattn_tma_supported(A, dtype)) return linattn_tma_launch(A, B, dtype, s);
    if (!force_simt && linattn_tc_supported(A, dtype)) return linattn_tc_launch(A, B, dtype, s);
    const bool ch_fast = A.qc == 1 && A.yc == 1;
    EL_DISPATCH_DTYPE(dtype, {
        if (ch_fast) {
            cudaFuncSetAttribute(linattn_simt_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnSimtSmem);
            linattn_simt_kernel<T, true><<<B * heads, 256, kAttnSimtSmem, s>>>(A);
        } else {
            cudaFuncSetAttribute(linattn_simt_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnSimtSmem);
            linattn_simt_kernel<T, false><<<B * heads, 256, kAttnSimtSmem, s>>>(A);
        }
    });
    note_launches(1);
    return check_launch();
}
