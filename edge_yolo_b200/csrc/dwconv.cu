// Depthwise k x k convolution (stride 1, "same" padding) over channel-contiguous (NHWC) views, with an optional fused
// bias + activation epilogue.  Engine-path replacement for the depthwise halves of the reference's separable blocks:
//   DSConv.dw  (nn/modules/conv.py:87-104: Conv2d(c, c, k, groups=c, bias=False), k = 3 and 7 in DSBottleneck,
//               nn/modules/block.py:1467-1503)           -> no epilogue, the result feeds the pointwise conv
//   DWConv     (nn/modules/conv.py:107-112, the Detect cls tower head.py:66-71) -> folded-BN bias + SiLU epilogue
// HBM-bound by design for k = 3 (every input element read from DRAM once, every output written once); k = 7 is bounded by
// fp32 FMA issue (49 taps per output) and is written to keep the load/store pipe below it.
//
// A CTA owns a tile of TH output rows x TW output columns x up to 64 channels.  The input patch (tile + halo) is staged once in
// shared memory with 16-byte cp.async in the global (row, column, channel-vector) order -- coalesced reads, linear writes,
// zero fill outside the image.  A thread <-> (channel vector cv, output column x, strip of 4 rows): lanes run over (cv, x), so
// every LDS.128 / STG.128 of a warp is one contiguous 512-byte span.  The register tile is VERTICAL: for each filter column kx
// the thread loads the 4 + k - 1 input vectors of its column once, and every filter tap (one 16-byte broadcast per 4 channels)
// feeds 4 output rows x V FMAs.
#include "el_common.cuh"

namespace el {

template <typename T> __device__ __forceinline__ float dw_silu(float v) {
    if constexpr (sizeof(T) == 2) {  // x * sigmoid(x) = h + h * tanh(h), h = x / 2 (one MUFU; error below 16-bit rounding)
        const float h = 0.5f * v;
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
        return fmaf(h, t, h);
    } else {
        return v / (1.f + expf(-v));
    }
}

constexpr int kDwRT = 4;          // output rows per strip = vertical register tile
constexpr int kDwMaxThreads = 256;

struct DwGeom {
    int CVL, cvl_shift;   // channel vectors per CTA (power of two, <= 8)
    int TW;               // output columns per tile
    int NS;               // thread strips per CTA (threads = CVL * TW * NS)
    int TH;               // output rows per tile (multiple of kDwRT)
    int n_cb;             // channel blocks
};

__device__ __forceinline__ uint32_t dw_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <typename T, int K>
__global__ void __launch_bounds__(kDwMaxThreads) dwconv_tile_kernel(const T* __restrict__ x, Strides4 xs, const float* __restrict__ w,
                                                                    const float* __restrict__ bias, T* __restrict__ o, Strides4 os, int C, int H, int W,
                                                                    int act, DwGeom G) {
    constexpr int V = Vec16<T>::N, HV = V / 4, P = K / 2, NIN = kDwRT + K - 1;
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int CVL = G.CVL;
    float* s_w = reinterpret_cast<float*>(s_raw);                                   // [tap][HV][cvl][4]
    uint4* s_in = reinterpret_cast<uint4*>(s_raw + (size_t)K * K * CVL * V * 4);    // [row][column][cv]
    const int cbk = (int)blockIdx.z % G.n_cb;
    const int64_t n = (int)blockIdx.z / G.n_cb;
    const int c0 = cbk * CVL * V;
    const int tx0 = (int)blockIdx.x * G.TW, ty0 = (int)blockIdx.y * G.TH;
    const int TH = min(G.TH, H - ty0), rows_in = TH + K - 1, PWf = G.TW + K - 1;
    const int nthreads = blockDim.x, tid = threadIdx.x;
    pdl_launch_dependents();
    // filter taps first (constants), then wait for the producer of x.  caller layout: w[tap][C] fp32 (tap-major)
    for (int i = tid; i < K * K * CVL * V; i += nthreads) {
        const int tap = i / (CVL * V), c = i - tap * (CVL * V), cvl = c / V, e = c - cvl * V;
        s_w[((tap * HV + e / 4) * CVL + cvl) * 4 + (e & 3)] = __ldg(w + (int64_t)tap * C + c0 + c);
    }
    pdl_wait();
    {   // ---- stage the input patch: pieces (row r, column p, vector cv), cv fastest = contiguous in global memory and in the patch
        const T* xb = x + n * xs.n + c0;
        const int per_row = PWf << G.cvl_shift;
        for (int r = 0; r < rows_in; ++r) {
            const int iy = ty0 - P + r;
            const bool row_ok = iy >= 0 && iy < H;
            const T* xrow = xb + (int64_t)iy * xs.h;
            const uint32_t drow = dw_smem_addr(s_in + (size_t)r * per_row);
            for (int i = tid; i < per_row; i += nthreads) {
                const int cv = i & (CVL - 1), ix = tx0 - P + (i >> G.cvl_shift);
                const bool ok = row_ok && ix >= 0 && ix < W;
                const T* src = ok ? xrow + (int64_t)ix * xs.w + cv * V : xb;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(drow + (uint32_t)i * 16), "l"(src), "r"(ok ? 16u : 0u) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int cvl = tid & (CVL - 1), rest = tid >> G.cvl_shift;
    const int xl = rest % G.TW, strip0 = rest / G.TW;
    const int ox = tx0 + xl;
    if (strip0 >= G.NS || ox >= W) return;
    const int ch = c0 + cvl * V;
    float bv[V];
#pragma unroll
    for (int e = 0; e < V; ++e) bv[e] = bias ? __ldg(bias + ch + e) : 0.f;
    const float4* sw4 = reinterpret_cast<const float4*>(s_w);
    const int row_stride = PWf << G.cvl_shift;  // uint4 units
    for (int ly0 = strip0 * kDwRT; ly0 < TH; ly0 += G.NS * kDwRT) {
        // accumulators, inputs and taps are kept as fp32 pairs of adjacent channels: one FFMA2 per pair
        f32x2 acc2[kDwRT][V / 2];
#pragma unroll
        for (int r = 0; r < kDwRT; ++r)
#pragma unroll
            for (int e = 0; e < V / 2; ++e) acc2[r][e] = pack_f32x2(bv[2 * e], bv[2 * e + 1]);
#pragma unroll 1
        for (int kx = 0; kx < K; ++kx) {
            const uint4* scol = s_in + (size_t)ly0 * row_stride + ((xl + kx) << G.cvl_shift) + cvl;
            f32x2 v2[NIN][V / 2];
#pragma unroll
            for (int j = 0; j < NIN; ++j) {
                // rows below the tile's last output row + halo are not staged: clamp (their products land in rows that are never stored)
                const int rr = min(ly0 + j, rows_in - 1) - ly0;
                float f[V];
                unpack<T>(scol[(size_t)rr * row_stride], f);
#pragma unroll
                for (int e = 0; e < V / 2; ++e) v2[j][e] = pack_f32x2(f[2 * e], f[2 * e + 1]);
            }
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                f32x2 w2[V / 2];
#pragma unroll
                for (int h = 0; h < HV; ++h) {
                    const float4 w4 = sw4[((ky * K + kx) * HV + h) * CVL + cvl];
                    w2[2 * h] = pack_f32x2(w4.x, w4.y); w2[2 * h + 1] = pack_f32x2(w4.z, w4.w);
                }
#pragma unroll
                for (int r = 0; r < kDwRT; ++r)
#pragma unroll
                    for (int e = 0; e < V / 2; ++e) acc2[r][e] = fma_f32x2(v2[r + ky][e], w2[e], acc2[r][e]);
            }
        }
        float acc[kDwRT][V];
#pragma unroll
        for (int r = 0; r < kDwRT; ++r)
#pragma unroll
            for (int e = 0; e < V / 2; ++e) unpack_f32x2(acc2[r][e], acc[r][2 * e], acc[r][2 * e + 1]);
        T* q = o + n * os.n + (int64_t)(ty0 + ly0) * os.h + (int64_t)ox * os.w + ch;
#pragma unroll
        for (int r = 0; r < kDwRT; ++r) {
            if (ly0 + r < TH) {
                float f[V];
#pragma unroll
                for (int e = 0; e < V; ++e) f[e] = act == 1 ? dw_silu<T>(acc[r][e]) : (act == 2 ? fmaxf(acc[r][e], 0.f) : acc[r][e]);
                *reinterpret_cast<uint4*>(q + (int64_t)r * os.h) = pack<T>(f);
            }
        }
    }
}

template <typename T, int K>
static int launch_dw(const void* x, Strides4 xs, const float* w, const float* bias, void* out, Strides4 os, int B, int C, int H, int W, int act,
                     cudaStream_t st) {
    constexpr int V = Vec16<T>::N;
    if (C % V) return EL_ERR_UNSUPPORTED;
    const int CV = C / V;
    DwGeom G;
    G.CVL = CV < 8 ? CV : 8;
    if ((G.CVL & (G.CVL - 1)) || CV % G.CVL) return EL_ERR_UNSUPPORTED;  // 1, 2, 4 or 8 vectors per CTA
    G.cvl_shift = G.CVL == 1 ? 0 : (G.CVL == 2 ? 1 : (G.CVL == 4 ? 2 : 3));
    G.n_cb = CV / G.CVL;
    const int tw_max = kDwMaxThreads / G.CVL;
    const int col_tiles = (int)ceil_div(W, tw_max);
    G.TW = (int)ceil_div(W, col_tiles);
    G.NS = kDwMaxThreads / (G.CVL * G.TW);
    if (G.NS < 1) G.NS = 1;
    const size_t row_bytes = (size_t)(G.TW + K - 1) * G.CVL * 16, w_bytes = (size_t)K * K * G.CVL * V * 4;
    // tile height: up to ~72 KB of shared memory (3 CTAs per SM), no more than the map needs, and enough CTAs to fill the GPU twice
    int th = (int)(((72 * 1024 - w_bytes) / row_bytes - (K - 1)) / kDwRT) * kDwRT;
    const int th_need = (int)ceil_div(H, kDwRT) * kDwRT;
    if (th > th_need) th = th_need;
    if (th < kDwRT) return EL_ERR_UNSUPPORTED;
    while (th > kDwRT && (int64_t)col_tiles * ceil_div(H, th) * B * G.n_cb < 2 * kSMs) th -= kDwRT;
    G.TH = th;
    if (G.NS > th / kDwRT) G.NS = th / kDwRT;
    const size_t smem = w_bytes + (size_t)(th + K - 1) * row_bytes;
    const int threads = (int)ceil_div(G.CVL * G.TW * G.NS, 32) * 32;
    const int64_t gz = (int64_t)B * G.n_cb;
    if (gz > 65535 || ceil_div(H, th) > 65535 || threads > kDwMaxThreads) return EL_ERR_UNSUPPORTED;
    dim3 grid((unsigned)col_tiles, (unsigned)ceil_div(H, th), (unsigned)gz);
    cudaError_t e = cudaFuncSetAttribute(dwconv_tile_kernel<T, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return EL_ERR_CUDA; }
    e = launch_pdl(dwconv_tile_kernel<T, K>, grid, dim3(threads), smem, st, (const T*)x, xs, w, bias, (T*)out, os, C, H, W, act, G);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return EL_ERR_CUDA; }
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
        if (k == 3) rc = launch_dw<T, 3>(x, xs, w, bias, out, os, B, C, H, W, act, st);
        else if (k == 5) rc = launch_dw<T, 5>(x, xs, w, bias, out, os, B, C, H, W, act, st);
        else rc = launch_dw<T, 7>(x, xs, w, bias, out, os, B, C, H, W, act, st);
    });
    if (rc != EL_OK) return rc;
    note_launches(1);
    return check_launch();
}
