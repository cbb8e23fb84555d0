// Backward of the linear-attention core (LinearAttention.forward, nn/modules/block.py:3364-3372), fp32 CUDA-core kernel
// for every dtype (training is not the throughput path; arithmetic is fp32 regardless of the storage type).
//
//   Ks = softmax_d(K)            Qs = softmax_N(Q) = P / s,  P = exp(Q - max_N Q)
//   ctx = Ks^T V                 Y  = Qs ctx
//   gQs = gY ctx^T               gctx = Qs^T gY            gKs = V gctx^T           gV = Ks gctx
//   gQ  = Qs * (gQs - sum_N Qs gQs)   (softmax over tokens, per channel)
//   gK  = Ks * (gKs - sum_d Ks gKs)   (softmax over channels, per token)
// One CTA per (image, head), four passes over the tokens in 64-token tiles (the operands of one head fit in L2):
//   0: column max / sum of Q          1: ctx and gctx (two 64x64 accumulations)
//   2: dq[i] = sum_N Qs gQs           3: gQ, gK, gV tiles
#include "el_common.cuh"

namespace el {

struct AttnBwdArgs {
    const void* qkv; int64_t qb, qc, qn;   // forward input (B, 3C, N)
    const void* gy; int64_t gb, gc, gn;    // upstream gradient (B, C, N)
    void* gqkv; int64_t ob, oc, on;        // output gradient (B, 3C, N)
    int heads, N;
};

namespace bwd {

constexpr int kD = 64, kTN = 64, kLd = kTN + 1;

template <typename T>
__device__ __forceinline__ void load_tile(float* __restrict__ sm, const T* __restrict__ base, int64_t sc, int64_t sn, int n0, int N, float fill, bool ch_fast) {
    for (int idx = threadIdx.x; idx < kD * kTN; idx += blockDim.x) {
        int ch, n;
        if (ch_fast) { n = idx >> 6; ch = idx & 63; } else { ch = idx >> 6; n = idx & 63; }
        const int tok = n0 + n;
        sm[ch * kLd + n] = tok < N ? to_f(base[(int64_t)ch * sc + (int64_t)tok * sn]) : fill;
    }
}
template <typename T>
__device__ __forceinline__ void store_tile(const float* __restrict__ sm, T* __restrict__ base, int64_t sc, int64_t sn, int n0, int N, bool ch_fast) {
    for (int idx = threadIdx.x; idx < kD * kTN; idx += blockDim.x) {
        int ch, n;
        if (ch_fast) { n = idx >> 6; ch = idx & 63; } else { ch = idx >> 6; n = idx & 63; }
        const int tok = n0 + n;
        if (tok < N) base[(int64_t)ch * sc + (int64_t)tok * sn] = from_f<T>(sm[ch * kLd + n]);
    }
}
// in-place softmax over the 64 channels of every token column of a [64][kLd] tile (threads 0..63, one column each)
__device__ __forceinline__ void softmax_channels(float* __restrict__ sm) {
    if (threadIdx.x < kTN) {
        const int t = threadIdx.x;
        float m = -INFINITY;
        for (int d = 0; d < kD; ++d) m = fmaxf(m, sm[d * kLd + t]);
        float s = 0.f;
        for (int d = 0; d < kD; ++d) { float e = expf(sm[d * kLd + t] - m); sm[d * kLd + t] = e; s += e; }
        const float inv = 1.f / s;
        for (int d = 0; d < kD; ++d) sm[d * kLd + t] *= inv;
    }
}
// acc[r][c] += sum_n A[(4*ti+r)][n] * B[(4*tj+c)][n]   (both tiles [64][kLd], contraction over the token axis)
__device__ __forceinline__ void outer_acc(float (&acc)[4][4], const float* __restrict__ A, const float* __restrict__ B, int ti, int tj) {
#pragma unroll 4
    for (int n = 0; n < kTN; ++n) {
        float a[4], b[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = A[(4 * ti + r) * kLd + n];
#pragma unroll
        for (int c = 0; c < 4; ++c) b[c] = B[(4 * tj + c) * kLd + n];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] += a[r] * b[c];
    }
}
// out[(4*tr+r)][4*tn+c] = sum_k M[k-major index] * X[k][4*tn+c]; M is a 64x64 matrix in smem (ld 65), transposed access selectable
template <bool M_TRANSPOSED>
__device__ __forceinline__ void mat_tile(float* __restrict__ out, const float* __restrict__ M, const float* __restrict__ X, int tr, int tn) {
    float o[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[r][c] = 0.f;
#pragma unroll 4
    for (int k = 0; k < kD; ++k) {
        float mr[4], xc[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) mr[r] = M_TRANSPOSED ? M[k * 65 + 4 * tr + r] : M[(4 * tr + r) * 65 + k];
#pragma unroll
        for (int c = 0; c < 4; ++c) xc[c] = X[k * kLd + 4 * tn + c];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) o[r][c] += mr[r] * xc[c];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) out[(4 * tr + r) * kLd + 4 * tn + c] = o[r][c];
}

template <typename T>
__global__ void __launch_bounds__(256) linattn_bwd_kernel(const __grid_constant__ AttnBwdArgs A) {
    extern __shared__ float smem[];
    float* sA = smem;                  // [64][kLd] tile buffers
    float* sB = sA + kD * kLd;
    float* sC = sB + kD * kLd;
    float* sD = sC + kD * kLd;
    float* sCtx = sD + kD * kLd;       // ctx   [i][j] (ld 65)
    float* sG = sCtx + kD * 65;        // gctx  [i][j] (ld 65)
    float* sMax = sG + kD * 65;        // per channel: max, 1/sum, dq
    float* sInv = sMax + kD;
    float* sDq = sInv + kD;

    const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
    const int b = blockIdx.x / A.heads, head = blockIdx.x % A.heads;
    const int C = A.heads * kD, N = A.N;
    const bool ch_fast = A.qc == 1 && A.gc == 1 && A.oc == 1;
    const T* qkv = reinterpret_cast<const T*>(A.qkv) + (int64_t)b * A.qb;
    const T* gq_in = qkv + (int64_t)(0 * C + head * kD) * A.qc;
    const T* gk_in = qkv + (int64_t)(1 * C + head * kD) * A.qc;
    const T* gv_in = qkv + (int64_t)(2 * C + head * kD) * A.qc;
    const T* gy = reinterpret_cast<const T*>(A.gy) + (int64_t)b * A.gb + (int64_t)(head * kD) * A.gc;
    T* out = reinterpret_cast<T*>(A.gqkv) + (int64_t)b * A.ob;
    T* oq = out + (int64_t)(0 * C + head * kD) * A.oc;
    T* ok = out + (int64_t)(1 * C + head * kD) * A.oc;
    T* ov = out + (int64_t)(2 * C + head * kD) * A.oc;

    // ---- pass 0: per-channel max and sum of exp over all tokens of Q
    if (tid < kD) { sMax[tid] = -INFINITY; sInv[tid] = 0.f; sDq[tid] = 0.f; }
    for (int n0 = 0; n0 < N; n0 += kTN) {
        __syncthreads();
        load_tile<T>(sA, gq_in, A.qc, A.qn, n0, N, -INFINITY, ch_fast);
        __syncthreads();
        if (tid < kD) {
            float m_old = sMax[tid], m = m_old;
            for (int n = 0; n < kTN; ++n) m = fmaxf(m, sA[tid * kLd + n]);
            float s = 0.f;
            for (int n = 0; n < kTN; ++n) s += expf(sA[tid * kLd + n] - m);
            sInv[tid] = sInv[tid] * (m_old == -INFINITY ? 0.f : expf(m_old - m)) + s;
            sMax[tid] = m;
        }
    }
    __syncthreads();
    if (tid < kD) sInv[tid] = 1.f / sInv[tid];

    // ---- pass 1: ctx[i][j] = sum_n Ks[n][i] V[n][j],  gctx[i][j] = sum_n Qs[n][i] gY[n][j]
    float acc_c[4][4], acc_g[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc_c[r][c] = 0.f; acc_g[r][c] = 0.f; }
    for (int n0 = 0; n0 < N; n0 += kTN) {
        __syncthreads();
        load_tile<T>(sA, gk_in, A.qc, A.qn, n0, N, 0.f, ch_fast);
        load_tile<T>(sB, gv_in, A.qc, A.qn, n0, N, 0.f, ch_fast);   // zero V / gY rows kill the padded tokens
        load_tile<T>(sC, gq_in, A.qc, A.qn, n0, N, -INFINITY, ch_fast);
        load_tile<T>(sD, gy, A.gc, A.gn, n0, N, 0.f, ch_fast);
        __syncthreads();
        softmax_channels(sA);
        for (int idx = tid; idx < kD * kTN; idx += 256) {
            const int i = idx >> 6, n = idx & 63;
            sC[i * kLd + n] = expf(sC[i * kLd + n] - sMax[i]) * sInv[i];
        }
        __syncthreads();
        outer_acc(acc_c, sA, sB, ti, tj);
        outer_acc(acc_g, sC, sD, ti, tj);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            sCtx[(4 * ti + r) * 65 + 4 * tj + c] = acc_c[r][c];
            sG[(4 * ti + r) * 65 + 4 * tj + c] = acc_g[r][c];
        }

    // ---- pass 2: dq[i] = sum_n Qs[n][i] * gQs[n][i],  gQs[n][i] = sum_j gY[n][j] ctx[i][j]
    for (int n0 = 0; n0 < N; n0 += kTN) {
        __syncthreads();
        load_tile<T>(sC, gq_in, A.qc, A.qn, n0, N, -INFINITY, ch_fast);
        load_tile<T>(sD, gy, A.gc, A.gn, n0, N, 0.f, ch_fast);
        __syncthreads();
        mat_tile<false>(sA, sCtx, sD, ti, tj);  // sA[i][n] = sum_j ctx[i][j] gY[j][n]
        __syncthreads();
        if (tid < kD) {
            float d = 0.f;
            for (int n = 0; n < kTN; ++n) d += expf(sC[tid * kLd + n] - sMax[tid]) * sInv[tid] * sA[tid * kLd + n];
            sDq[tid] += d;
        }
    }

    // ---- pass 3: gradients per token tile
    for (int n0 = 0; n0 < N; n0 += kTN) {
        __syncthreads();
        load_tile<T>(sC, gq_in, A.qc, A.qn, n0, N, -INFINITY, ch_fast);
        load_tile<T>(sD, gy, A.gc, A.gn, n0, N, 0.f, ch_fast);
        __syncthreads();
        mat_tile<false>(sA, sCtx, sD, ti, tj);  // gQs tile
        __syncthreads();
        for (int idx = tid; idx < kD * kTN; idx += 256) {  // gQ = Qs (gQs - dq)
            const int i = idx >> 6, n = idx & 63;
            const float qs = expf(sC[i * kLd + n] - sMax[i]) * sInv[i];
            sA[i * kLd + n] = qs * (sA[i * kLd + n] - sDq[i]);
        }
        __syncthreads();
        store_tile<T>(sA, oq, A.oc, A.on, n0, N, ch_fast);
        __syncthreads();
        load_tile<T>(sA, gk_in, A.qc, A.qn, n0, N, 0.f, ch_fast);
        load_tile<T>(sB, gv_in, A.qc, A.qn, n0, N, 0.f, ch_fast);
        __syncthreads();
        softmax_channels(sA);                    // Ks tile
        __syncthreads();
        mat_tile<true>(sC, sG, sA, ti, tj);      // gV[j][n]  = sum_i gctx[i][j] Ks[i][n]
        mat_tile<false>(sD, sG, sB, ti, tj);     // gKs[i][n] = sum_j gctx[i][j] V[j][n]
        __syncthreads();
        store_tile<T>(sC, ov, A.oc, A.on, n0, N, ch_fast);
        if (tid < kTN) {  // gK = Ks (gKs - sum_d Ks gKs) per token
            const int t = tid;
            float dot = 0.f;
            for (int d = 0; d < kD; ++d) dot += sA[d * kLd + t] * sD[d * kLd + t];
            for (int d = 0; d < kD; ++d) sD[d * kLd + t] = sA[d * kLd + t] * (sD[d * kLd + t] - dot);
        }
        __syncthreads();
        store_tile<T>(sD, ok, A.oc, A.on, n0, N, ch_fast);
    }
}

constexpr size_t kSmem = (size_t)(4 * kD * kLd + 2 * kD * 65 + 3 * kD) * sizeof(float);

}  // namespace bwd
}  // namespace el

using namespace el;

extern "C" int el_linattn_bwd(const void* qkv, const int64_t qs[3], const void* gy, const int64_t gs[3], void* gqkv, const int64_t os[3], int B, int heads,
                              int N, int dtype, void* stream) {
    if (!qkv || !gy || !gqkv || !qs || !gs || !os || B <= 0 || heads <= 0 || N <= 0) return EL_ERR_ARG;
    AttnBwdArgs A{qkv, qs[0], qs[1], qs[2], gy, gs[0], gs[1], gs[2], gqkv, os[0], os[1], os[2], heads, N};
    cudaStream_t s = (cudaStream_t)stream;
    EL_DISPATCH_DTYPE(dtype, {
        cudaFuncSetAttribute(bwd::linattn_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd::kSmem);
        bwd::linattn_bwd_kernel<T><<<B * heads, 256, bwd::kSmem, s>>>(A);
    });
    note_launches(1);
    return check_launch();
}
