// Depthwise k x k convolution (stride 1, "same" padding) over channel-contiguous (NHWC) views, with an optional fused
// bias + activation epilogue.  Engine-path replacement for the depthwise halves of the reference's separable blocks:
//   DSConv.dw  (nn/modules/conv.py:87-104: Conv2d(c, c, k, groups=c, bias=False), k = 3 and 7 in DSBottleneck,
//               nn/modules/block.py:1467-1503)           -> no epilogue, the result feeds the pointwise conv
//   DWConv     (nn/modules/conv.py:107-112, the Detect cls tower head.py:66-71) -> folded-BN bias + SiLU epilogue
// HBM-bound by design: every input element is read from DRAM once (the k x k window re-reads hit L1), every output
// written once; arithmetic is fp32 FMA on CUDA cores (a depthwise filter has no contraction for the tensor cores).
//
// Mapping: thread <-> (channel vector of 16 B, group of OC adjacent output columns, tile of RT rows).  For one output
// row the thread walks the k input rows, loads the OC+k-1 input vectors of each row once, unpacks them once and feeds
// OC x k x V FMAs; filter taps come from shared memory (fp32, one 16 B broadcast per 4 channels).
#include "el_common.cuh"

namespace el {

template <typename T> __device__ __forceinline__ float dw_silu(float v) {
    if constexpr (sizeof(T) == 2) return __fdividef(v, 1.f + __expf(-v));
    else return v / (1.f + expf(-v));
}

constexpr int kDwThreads = 128;
constexpr int kDwChanBlock = 64;  // channels per CTA (filter taps of one channel block live in shared memory)

template <typename T, int K, int OC, int RT>
__global__ void __launch_bounds__(kDwThreads) dwconv_kernel(const T* __restrict__ x, Strides4 xs, const float* __restrict__ w, const float* __restrict__ bias,
                                                            T* __restrict__ o, Strides4 os, int C, int H, int W, int act, int n_cg, int n_rt,
                                                            uint32_t items) {
    constexpr int V = Vec16<T>::N, HV = V / 4, P = K / 2, NIN = OC + K - 1;
    extern __shared__ __align__(16) float s_w[];  // [tap][HV][cvl][4]
    const int c0 = (int)blockIdx.y * kDwChanBlock;
    const int cb = min(kDwChanBlock, C - c0), CVL = cb / V;
    // caller layout: w[tap][C] fp32 (tap-major)
    for (int i = threadIdx.x; i < K * K * cb; i += kDwThreads) {
        const int tap = i / cb, c = i - tap * cb, cvl = c / V, e = c - cvl * V;
        s_w[((tap * HV + e / 4) * CVL + cvl) * 4 + (e & 3)] = __ldg(w + (int64_t)tap * C + c0 + c);
    }
    __syncthreads();
    const uint32_t idx = blockIdx.x * kDwThreads + threadIdx.x;
    if (idx >= items) return;
    uint32_t t = idx;
    const int cvl = (int)(t % (uint32_t)CVL); t /= (uint32_t)CVL;
    const int cg = (int)(t % (uint32_t)n_cg); t /= (uint32_t)n_cg;
    const int rt = (int)(t % (uint32_t)n_rt);
    const int64_t n = t / (uint32_t)n_rt;
    const int ch = c0 + cvl * V, x0 = cg * OC, y0 = rt * RT, y1 = min(y0 + RT, H);
    const T* px = x + n * xs.n + ch;
    T* po = o + n * os.n + ch;
    float bv[V];
#pragma unroll
    for (int e = 0; e < V; ++e) bv[e] = bias ? __ldg(bias + ch + e) : 0.f;
    const float4* sw4 = reinterpret_cast<const float4*>(s_w);
    for (int y = y0; y < y1; ++y) {
        float acc[OC][V];
#pragma unroll
        for (int oc = 0; oc < OC; ++oc)
#pragma unroll
            for (int e = 0; e < V; ++e) acc[oc][e] = bv[e];
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
            const int iy = y + ky - P;
            if (iy < 0 || iy >= H) continue;
            const T* prow = px + (int64_t)iy * xs.h;
            float v[NIN][V];
#pragma unroll
            for (int j = 0; j < NIN; ++j) {
                const int ix = x0 + j - P;
                uint4 raw = make_uint4(0, 0, 0, 0);
                if (ix >= 0 && ix < W) raw = ldg_cached(prow + (int64_t)ix * xs.w);
                unpack<T>(raw, v[j]);
            }
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                float wv[V];
#pragma unroll
                for (int h = 0; h < HV; ++h) {
                    const float4 w4 = sw4[((ky * K + kx) * HV + h) * CVL + cvl];
                    wv[4 * h] = w4.x; wv[4 * h + 1] = w4.y; wv[4 * h + 2] = w4.z; wv[4 * h + 3] = w4.w;
                }
#pragma unroll
                for (int oc = 0; oc < OC; ++oc)
#pragma unroll
                    for (int e = 0; e < V; ++e) acc[oc][e] = fmaf(v[oc + kx][e], wv[e], acc[oc][e]);
            }
        }
        T* q = po + (int64_t)y * os.h + (int64_t)x0 * os.w;
#pragma unroll
        for (int oc = 0; oc < OC; ++oc) {
            if (x0 + oc < W) {
                float r[V];
#pragma unroll
                for (int e = 0; e < V; ++e) r[e] = act == 1 ? dw_silu<T>(acc[oc][e]) : (act == 2 ? fmaxf(acc[oc][e], 0.f) : acc[oc][e]);
                *reinterpret_cast<uint4*>(q + (int64_t)oc * os.w) = pack<T>(r);
            }
        }
    }
}

template <typename T, int K, int OC, int RT>
static int launch_dw(const void* x, Strides4 xs, const float* w, const float* bias, void* out, Strides4 os, int B, int C, int H, int W, int act,
                     cudaStream_t st) {
    constexpr int V = Vec16<T>::N;
    const int n_cg = (int)ceil_div(W, OC), n_rt = (int)ceil_div(H, RT);
    const int n_cb = (int)ceil_div(C, kDwChanBlock);
    if (C % V || (C > kDwChanBlock && C % kDwChanBlock)) return EL_ERR_UNSUPPORTED;
    const int cvl = (C < kDwChanBlock ? C : kDwChanBlock) / V;
    const int64_t items = (int64_t)cvl * n_cg * n_rt * B;
    if (items >= (1ll << 31) || n_cb > 65535) return EL_ERR_UNSUPPORTED;
    dim3 grid((unsigned)ceil_div(items, kDwThreads), (unsigned)n_cb);
    const size_t smem = (size_t)K * K * kDwChanBlock * sizeof(float);
    dwconv_kernel<T, K, OC, RT><<<grid, kDwThreads, smem, st>>>((const T*)x, xs, w, bias, (T*)out, os, C, H, W, act, n_cg, n_rt, (uint32_t)items);
    return EL_OK;
}

}  // namespace el

using namespace el;

extern "C" int el_dwconv_fwd(const void* x, const int64_t xs_[4], const float* w, const float* bias, void* out, const int64_t os_[4], int B, int C,
                             int H, int W, int k, int act, int dtype, void* stream) {
    if (!x || !w || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0 || act < 0 || act > 2) return EL_ERR_ARG;
    if (k != 3 && k != 5 && k != 7) return EL_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const Strides4 xs = s4(xs_), os = s4(os_);
    int rc = EL_ERR_UNSUPPORTED;
    EL_DISPATCH_DTYPE(dtype, {
        if (!channel_vectorisable<T>(x, xs, C) || !channel_vectorisable<T>(out, os, C)) return EL_ERR_UNSUPPORTED;
        // small maps get shorter row tiles so that the grid still covers the 148 SMs several times
        const bool small = (int64_t)B * H * W * (C / Vec16<T>::N) < (int64_t)kSMs * kDwThreads * 4 * 4 * 8;
        if (k == 3) rc = small ? launch_dw<T, 3, 4, 2>(x, xs, w, bias, out, os, B, C, H, W, act, st) : launch_dw<T, 3, 4, 8>(x, xs, w, bias, out, os, B, C, H, W, act, st);
        else if (k == 5) rc = small ? launch_dw<T, 5, 4, 2>(x, xs, w, bias, out, os, B, C, H, W, act, st) : launch_dw<T, 5, 4, 8>(x, xs, w, bias, out, os, B, C, H, W, act, st);
        else rc = small ? launch_dw<T, 7, 4, 2>(x, xs, w, bias, out, os, B, C, H, W, act, st) : launch_dw<T, 7, 4, 8>(x, xs, w, bias, out, os, B, C, H, W, act, st);
    });
    if (rc != EL_OK) return rc;
    note_launches(1);
    return check_launch();
}
