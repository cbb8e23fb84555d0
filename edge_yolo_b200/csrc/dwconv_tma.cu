// Depthwise 3 x 3 convolution (stride 1, padding 1, NHWC 16-bit) as a persistent, TMA-pipelined streaming kernel.
// Engine-path replacement for the k = 3 depthwise halves of DSConv / DWConv (nn/modules/conv.py:87-104, 124-130; DSBottleneck
// nn/modules/block.py:1467-1503; Detect class towers nn/modules/head.py:66-71) where dwconv_tile_kernel (dwconv.cu) runs today.
//
// Why: dwconv_tile_kernel stages a patch with per-thread cp.async, waits for all of it, computes, stores and exits -- nothing inside a CTA
// overlaps the fetch with the arithmetic, and the 20 k = 3 launch sites of EdgeLine-n sat at 1.2 - 2.4 TB/s (0.2 - 0.36 of the HBM roofline,
// 372 us of a 2.8 ms step; bench.py `kernels.dwconv`, round 1).  Every input element is needed once from DRAM and every output written once
// (2 * B * C * H * W * e bytes), so this is a pure streaming problem; the 9 FMAs per output are ~80 instructions per 16-byte output vector,
// far below what the SM issues while HBM delivers that vector.  Design:
//   * one producer lane issues 4-D TMA boxes (CB channels, 22 columns, TH + 2 rows, 1 image: tile + halo; out-of-image pixels are zero-filled
//     by the TMA unit = the convolution's padding) into a ring of shared-memory stages (full / empty mbarriers), several tiles ahead;
//   * 320 compute threads <-> (channel vector cv, output column x of 20, strip of 5 output rows): lanes run over (cv, x), so every 16-byte
//     shared load and every 16-byte global store of a warp is a set of contiguous 128-byte runs; the 7 x 3 input vectors of a strip are
//     read from the stage once and each feeds up to three filter rows (vertical register tile), arithmetic as packed fp32x2 FMAs;
//   * CTAs are persistent (one per SM), tiles = (image, channel block of <= 64 channels, 5 * groups rows, 20 columns): 20 divides every
//     EdgeLine map width (160 / 80 / 40 / 20 and their 1280-pixel doubles), so no column is computed twice or wasted.
#include <cuda.h>

#include "el_common.cuh"

namespace el {
namespace dwt {

constexpr int kTW = 20, kRT = 5;           // output columns per tile, output rows per thread strip
constexpr int kCompute = 320, kThreads = 352;
constexpr int kMaxStages = 6;
constexpr int kPW = kTW + 2;                // patch columns

struct Args {
    CUtensorMap src_map;
    const float* w;      // [9][C] fp32, tap-major (ops.pack_dw_weight)
    const float* bias;   // [C] or null
    void* out; int64_t osn, osh, osw;
    int C, H, W, B, act, CVL, NG, TH, tiles_x, tiles_y, n_cb, stages;   // CVL: 16-byte channel vectors per channel block (1..8, any value)
    int64_t n_tiles;
    uint32_t stage_bytes, box_bytes;   // ring pitch (128-byte multiple) and the bytes one box delivers
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst), "l"(map),
                 "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
template <typename T> __device__ __forceinline__ float silu16(float v) {  // x * sigmoid(x) = h + h * tanh(h), h = x / 2 (one MUFU)
    const float h = 0.5f * v;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) dwconv3_tma_kernel(const __grid_constant__ Args A) {
    extern __shared__ __align__(128) unsigned char sm_raw[];
    const uint32_t sbase = (smem_addr(sm_raw) + 127u) & ~127u;
    unsigned char* sm = sm_raw + (sbase - smem_addr(sm_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = A.stages, C = A.C;
    const uint32_t off_w = (uint32_t)S * A.stage_bytes;
    float* s_w = reinterpret_cast<float*>(sm + off_w);             // [9][C]
    float* s_b = s_w + 9 * C;                                      // [C]
    const uint32_t bar_full = sbase + off_w + (uint32_t)(10 * C) * 4;   // C % 16 == 0: 8-byte aligned
    const uint32_t bar_empty = bar_full + 8 * kMaxStages;

    const int64_t first = blockIdx.x;
    const int my_tiles = first < A.n_tiles ? (int)((A.n_tiles - first + gridDim.x - 1) / gridDim.x) : 0;
    const int CVL = A.CVL, CB = CVL * 8;

    pdl_launch_dependents();
    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&A.src_map) : "memory");
        for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, kCompute / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 9 * C; i += kThreads) s_w[i] = __ldg(A.w + i);
    for (int i = tid; i < C; i += kThreads) s_b[i] = A.bias ? __ldg(A.bias + i) : 0.f;
    __syncthreads();
    pdl_wait();

    auto tile_coords = [&](int64_t tile, int& img, int& cb, int& ty, int& tx) {
        tx = (int)(tile % A.tiles_x); tile /= A.tiles_x;
        ty = (int)(tile % A.tiles_y); tile /= A.tiles_y;
        cb = (int)(tile % A.n_cb);
        img = (int)(tile / A.n_cb);
    };

    if (warp == kCompute / 32) {
        // ------------------------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            for (int tl = 0; tl < my_tiles; ++tl) {
                int img, cb, ty, tx;
                tile_coords(first + (int64_t)tl * gridDim.x, img, cb, ty, tx);
                const int s = tl % S, use = tl / S;
                if (use > 0) mbar_wait(bar_empty + 8 * s, (uint32_t)(use - 1) & 1);
                mbar_expect_tx(bar_full + 8 * s, A.box_bytes);
                tma_load_4d(sbase + (uint32_t)s * A.stage_bytes, &A.src_map, cb * CB, tx * kTW - 1, ty * A.TH - 1, img, bar_full + 8 * s);
            }
        }
        return;
    }
    // ---------------------------------------------------------------------------------------- compute threads
    constexpr int V = 8;
    const int cvl = tid % CVL, rest = tid / CVL;
    const int xl = rest % kTW, grp = rest / kTW;
    const bool active = grp < A.NG;                  // CVL * 20 * NG <= kCompute: the threads past the last strip only take part in the barriers
    const int ly0 = grp * kRT;
    const uint32_t row_pitch = (uint32_t)kPW * CVL * 16;    // bytes per patch row
    const uint32_t my_off = (uint32_t)ly0 * row_pitch + (uint32_t)(xl * CVL + cvl) * 16;
    T* outp = reinterpret_cast<T*>(A.out);
    for (int tl = 0; tl < my_tiles; ++tl) {
        int img, cb, ty, tx;
        tile_coords(first + (int64_t)tl * gridDim.x, img, cb, ty, tx);
        const int s = tl % S;
        const int ch = cb * CB + cvl * V;
        const int ox = tx * kTW + xl, oy0 = ty * A.TH + ly0;
        mbar_wait(bar_full + 8 * s, (uint32_t)(tl / S) & 1);
        const uint32_t base = sbase + (uint32_t)s * A.stage_bytes + my_off;
        f32x2 acc[kRT][V / 2];
        if (active) {
        {
            const float4 b0 = *reinterpret_cast<const float4*>(s_b + ch), b1 = *reinterpret_cast<const float4*>(s_b + ch + 4);
#pragma unroll
            for (int r = 0; r < kRT; ++r) {
                acc[r][0] = pack_f32x2(b0.x, b0.y); acc[r][1] = pack_f32x2(b0.z, b0.w);
                acc[r][2] = pack_f32x2(b1.x, b1.y); acc[r][3] = pack_f32x2(b1.z, b1.w);
            }
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            f32x2 v[kRT + 2][V / 2];
#pragma unroll
            for (int j = 0; j < kRT + 2; ++j) {
                float f[V];
                unpack<T>(lds128(base + (uint32_t)j * row_pitch + (uint32_t)(kx * CVL) * 16), f);
#pragma unroll
                for (int e = 0; e < V / 2; ++e) v[j][e] = pack_f32x2(f[2 * e], f[2 * e + 1]);
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float4 w0 = *reinterpret_cast<const float4*>(s_w + (ky * 3 + kx) * C + ch);
                const float4 w1 = *reinterpret_cast<const float4*>(s_w + (ky * 3 + kx) * C + ch + 4);
                const f32x2 w2[V / 2] = {pack_f32x2(w0.x, w0.y), pack_f32x2(w0.z, w0.w), pack_f32x2(w1.x, w1.y), pack_f32x2(w1.z, w1.w)};
#pragma unroll
                for (int r = 0; r < kRT; ++r)
#pragma unroll
                    for (int e = 0; e < V / 2; ++e) acc[r][e] = fma_f32x2(v[r + ky][e], w2[e], acc[r][e]);
            }
        }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);  // this warp has read everything it needs from the stage
        if (active && ox < A.W) {
            T* q = outp + (int64_t)img * A.osn + (int64_t)oy0 * A.osh + (int64_t)ox * A.osw + ch;
#pragma unroll
            for (int r = 0; r < kRT; ++r) {
                if (oy0 + r < A.H) {
                    float f[V];
#pragma unroll
                    for (int e = 0; e < V / 2; ++e) unpack_f32x2(acc[r][e], f[2 * e], f[2 * e + 1]);
                    if (A.act == 1) {
#pragma unroll
                        for (int e = 0; e < V; ++e) f[e] = silu16<T>(f[e]);
                    } else if (A.act == 2) {
#pragma unroll
                        for (int e = 0; e < V; ++e) f[e] = fmaxf(f[e], 0.f);
                    }
                    stg_stream(q + (int64_t)r * A.osh, pack<T>(f));
                }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

}  // namespace dwt

// k = 3, 16-bit, channel-contiguous views, C a multiple of 16 (or of 8 with a divisor 2..8 of the vector count): returns EL_ERR_UNSUPPORTED otherwise (the caller falls back to
// dwconv_tile_kernel).  Launch counting / error collection stay with the caller.
int dwconv3_tma_launch(const void* x, Strides4 xs, const float* w, const float* bias, void* out, Strides4 os, int B, int C, int H, int W, int act,
                       int dtype, cudaStream_t st) {
    if (dtype != EL_BF16 && dtype != EL_F16) return EL_ERR_UNSUPPORTED;
    if (C <= 0 || C % 8) return EL_ERR_UNSUPPORTED;
    int CVL = 1;  // channel block = the largest divisor of the vector count that is <= 8 (80 channels: blocks of 5 vectors)
    for (int d = 2; d <= 8; ++d) if ((C / 8) % d == 0) CVL = d;
    if (CVL < 2) return EL_ERR_UNSUPPORTED;
    if (xs.c != 1 || os.c != 1 || xs.n % 8 || xs.h % 8 || xs.w % 8 || os.n % 8 || os.h % 8 || os.w % 8 || !aligned16(x) || !aligned16(out)) return EL_ERR_UNSUPPORTED;
    dwt::EncodeTiledFn fn = dwt::encode_fn();
    if (!fn) return EL_ERR_CUDA;
    dwt::Args A{};
    const int CB = CVL * 8;
    A.CVL = CVL; A.NG = dwt::kCompute / (CVL * dwt::kTW); A.TH = A.NG * dwt::kRT;
    A.n_cb = C / CB;
    A.box_bytes = (uint32_t)(A.TH + 2) * dwt::kPW * CVL * 16;
    A.stage_bytes = (A.box_bytes + 127u) & ~127u;
    const size_t fixed = 128 + (size_t)10 * C * 4 + 16 * dwt::kMaxStages + 16;
    int S = (int)(((size_t)200 * 1024 - fixed) / A.stage_bytes);
    if (S > dwt::kMaxStages) S = dwt::kMaxStages;
    if (S < 2) return EL_ERR_UNSUPPORTED;
    A.stages = S;
    A.w = w; A.bias = bias; A.out = out; A.osn = os.n; A.osh = os.h; A.osw = os.w;
    A.C = C; A.H = H; A.W = W; A.B = B; A.act = act;
    A.tiles_x = (int)ceil_div(W, dwt::kTW); A.tiles_y = (int)ceil_div(H, A.TH);
    A.n_tiles = (int64_t)B * A.n_cb * A.tiles_x * A.tiles_y;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)xs.w * 2, (cuuint64_t)xs.h * 2, (cuuint64_t)(B > 1 ? xs.n : (int64_t)H * xs.h) * 2};
    const cuuint32_t box[4] = {(cuuint32_t)CB, (cuuint32_t)dwt::kPW, (cuuint32_t)(A.TH + 2), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (fn(&A.src_map, dtype == EL_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return EL_ERR_CUDA;
    const size_t smem = fixed + (size_t)S * A.stage_bytes;
    const int64_t gx = A.n_tiles < kSMs ? A.n_tiles : kSMs;
    cudaError_t e;
    if (dtype == EL_BF16) {
        e = cudaFuncSetAttribute(dwt::dwconv3_tma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = launch_pdl(dwt::dwconv3_tma_kernel<__nv_bfloat16>, dim3((unsigned)gx), dim3(dwt::kThreads), smem, st, A);
    } else {
        e = cudaFuncSetAttribute(dwt::dwconv3_tma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = launch_pdl(dwt::dwconv3_tma_kernel<__half>, dim3((unsigned)gx), dim3(dwt::kThreads), smem, st, A);
    }
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return EL_ERR_CUDA; }
    return EL_OK;
}

}  // namespace el
