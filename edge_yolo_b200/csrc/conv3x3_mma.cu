// Dense 3 x 3 convolution (stride 1 / 2, padding 1) with NARROW channel counts (C_in = 16 / 32, N <= 64) for the inference engine:
//     out[b, y, x, n] = act( sum_{ky,kx,c} X[b, s*y+ky-1, s*x+kx-1, c] * W[n, c, ky, kx] + bias[n] )            NHWC 16-bit activations
// These are the shared high-band convs f_h of _WaveletEnhancer (nn/modules/block.py:3668-3673: Conv(c, c/2, 3)) at the early, large maps:
// 16 -> 8 on three stacked 80 x 80 bands and 32 -> 16 on 40 x 40 for EdgeLine-n (a 3B-image batch each, `Conv.forward_fuse`
// nn/modules/conv.py:58-60).  el_conv3x3_fwd runs them as tcgen05 implicit GEMMs with nine tap-shifted TMA boxes per 128-pixel tile; with
// N = 8 / 16 a tile is nine 4 KB loads, eighteen tiny MMAs and a 2 KB store, and the time is the per-tile fixed cost (bench.py, round 2:
// 62 us for 59 MB = 0.15 of the HBM roofline at 16 -> 8 @ 80 x 80 x 192; 2 x 25.5 us at 32 -> 16 @ 40 x 40 x 192).
// Here the problem is what it is -- a streaming kernel with a little tensor-core work per pixel:
//   * CTA tile = 16 x 16 output pixels; the haloed 18 x 18 x C input tile is staged ONCE by 16-byte cp.async (zero fill = the padding),
//     double buffered across the tiles of a persistent CTA; pixel pitch C*2 + 16 bytes so that ldmatrix rows fall into distinct banks;
//   * a warp owns two output rows (two m16 tiles of 16 pixels); per (tap, 16-channel chunk) ONE ldmatrix.x4 -- the A fragment of tap
//     (ky, kx) is the same staged tile addressed at pixel (y + ky, x + kx) -- and one mma.sync.m16n8k16 per 8 output channels, weights as
//     B fragments in shared memory (built once per CTA from the plain fp32 (N, C, 3, 3) weight: no host packing);
//   * epilogue in registers: bias, SiLU / ReLU, 16-bit pack, 4-byte stores (a warp store = 8 pixels x 16 contiguous bytes);
//   * stride 2 (layer 1 of the yaml, Conv(16, 32, 3, 2) on the 320 x 320 stem output): every ldmatrix row has its own address, so the
//     16 pixels of an m-tile simply sit two staged pixels apart; the staged tile is 33 x 33.
#include <type_traits>

#include "el_common.cuh"

namespace el {
namespace c3m {

constexpr int kTile = 16;            // output tile edge
constexpr int kThreads = 256;
constexpr int kMaxN = 64, kMaxC = 32;

struct Args {
    const void* x; int64_t xn, xh, xw;      // element strides; channels contiguous
    void* out; int64_t on, oh, ow;
    const float* w;                          // (N, C, 3, 3) fp32
    const float* bias;                       // (N) or null
    int B, H, W, Ho, Wo, C, N, act, tiles_x, tiles_y;   // H, W: input; Ho, Wo: output
    int64_t n_tiles;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled (the source is not read)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <typename T> __device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <> __device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <> __device__ __forceinline__ void mma16816<__half>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
    __half2 t = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float silu_fast(float v) {  // x * sigmoid(x) = h + h * tanh(h), h = x / 2 (one MUFU)
    const float h = 0.5f * v;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

// KC = C / 16 K chunks per tap, NT = N / 8 output-channel tiles
template <typename T, int KC, int NT, int S>
__global__ void __launch_bounds__(kThreads) conv3x3_mma_kernel(const __grid_constant__ Args A) {
    constexpr int C = 16 * KC, N = 8 * NT;
    constexpr int kHalo = (kTile - 1) * S + 3;                    // staged tile edge: 18 (stride 1) / 33 (stride 2)
    constexpr uint32_t kPitch = C * 2 + 16;                       // bytes per staged pixel
    constexpr uint32_t kStage = kHalo * kHalo * kPitch;           // bytes per staged tile
    constexpr int kCpp = C / 8;                                   // 16-byte chunks per pixel
    constexpr int kChunks = kHalo * kHalo * kCpp;
    constexpr int kNS = S == 1 ? 3 : 2;                           // staged tiles per CTA
    extern __shared__ __align__(16) unsigned char sm[];
    T* s_wf = reinterpret_cast<T*>(sm);                           // B fragments: [tap][kc][nt][lane][4]
    constexpr uint32_t kWfBytes = 9 * KC * NT * 32 * 4 * 2;
    float* s_bias = reinterpret_cast<float*>(sm + kWfBytes);      // [N]
    const uint32_t tiles = smem_addr(sm) + kWfBytes + ((N * 4 + 15) & ~15);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    pdl_launch_dependents();
    // ---- B fragments of mma.m16n8k16 (k16 x n8, "col"): lane (g = lane / 4, t = lane % 4) holds k = 2t, 2t+1, 2t+8, 2t+9 of column n = g.
    //      The plain (N, C, 3, 3) weight is read with coalesced, independent 16-byte loads (all of a thread's loads in flight at once) and
    //      every element scattered to its slot: a per-slot gather (index arithmetic + one dependent scalar load per element, 18 rounds at
    //      C = 32, N = 16) put ~5 us of pure latency in front of every launch.
    {
        constexpr int kW4 = N * C * 9 / 4;                       // float4s of the weight (N * C * 9 is a multiple of 4: C % 16 == 0)
        constexpr int kPer = (kW4 + kThreads - 1) / kThreads;
        float4 wv[kPer];
#pragma unroll
        for (int r = 0; r < kPer; ++r) {
            const int q = tid + r * kThreads;
            wv[r] = q < kW4 ? __ldg(reinterpret_cast<const float4*>(A.w) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < kPer; ++r) {
            const int q = tid + r * kThreads;
            if (q < kW4) {
                const float v4[4] = {wv[r].x, wv[r].y, wv[r].z, wv[r].w};
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                    const int idx = 4 * q + e4, tap = idx % 9, nc = idx / 9, c = nc % C, n = nc / C;   // element W[n][c][tap]
                    const int k = c & 15, kc = c >> 4, nt = n >> 3, ln = 4 * (n & 7) + ((k & 7) >> 1), e = (k & 1) + 2 * (k >> 3);
                    s_wf[((((tap * KC + kc) * NT + nt) * 32 + ln) << 2) + e] = from_f<T>(v4[e4]);
                }
            }
        }
    }
    for (int i = tid; i < N; i += kThreads) s_bias[i] = A.bias ? __ldg(A.bias + i) : 0.f;
    pdl_wait();  // the producer of x has completed from here on

    const T* xp = reinterpret_cast<const T*>(A.x);
    T* outp = reinterpret_cast<T*>(A.out);
    const int per_img = A.tiles_x * A.tiles_y;
    auto stage_tile = [&](int64_t tile, int buf) {
        const int img = (int)(tile / per_img), r = (int)(tile - (int64_t)img * per_img);
        const int y0 = (r / A.tiles_x) * kTile * S - 1, x0 = (r % A.tiles_x) * kTile * S - 1;
        const T* base = xp + (int64_t)img * A.xn;
        const uint32_t dst0 = tiles + (uint32_t)buf * kStage;
        for (int i = tid; i < kChunks; i += kThreads) {
            const int ch = i % kCpp, p = i / kCpp, py = p / kHalo, px = p - py * kHalo;
            const int y = y0 + py, x = x0 + px;
            const bool ok = y >= 0 && y < A.H && x >= 0 && x < A.W;
            const T* src = ok ? base + (int64_t)y * A.xh + (int64_t)x * A.xw + ch * 8 : xp;
            cp_async16_zfill(dst0 + (uint32_t)p * kPitch + ch * 16, src, ok);
        }
        cp_async_commit();
    };
    const int64_t first = blockIdx.x, stride = gridDim.x;
    // ring of kNS staged tiles, kNS - 1 of them in flight ahead of the one being computed: with one tile of look-ahead the kernel moved
    // (CTAs per SM) x one tile per memory latency, i.e. ~2 TB/s whatever else was tuned
#pragma unroll
    for (int p = 0; p < kNS - 1; ++p) {
        if (first + p * stride < A.n_tiles) stage_tile(first + p * stride, p); else cp_async_commit();
    }
    // per-lane ldmatrix row: matrix (lane >> 3) = [pixels 0-7 | 8-15] x [k 0-7 | 8-15], row (lane & 7)
    const uint32_t a_lane = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * S) * kPitch + (uint32_t)(lane >> 4) * 16;
    const int g = lane >> 2, t4 = lane & 3;
    // up to 36 B fragments (72 registers) stay in registers for the whole kernel: one LDS.64 per MMA otherwise
    constexpr int kFrags = 9 * KC * NT;
    constexpr bool kRegB = kFrags <= 36;
    uint2 bfr[kRegB ? kFrags : 1];
    if constexpr (kRegB) {
        __syncthreads();  // the fragments written by the whole CTA above
#pragma unroll
        for (int f = 0; f < kFrags; ++f) bfr[f] = *reinterpret_cast<const uint2*>(s_wf + (f * 32 + lane) * 4);
    }
    int it = 0;
    for (int64_t tile = first; tile < A.n_tiles; tile += stride, ++it) {
        const int buf = it % kNS;
        if (tile + (kNS - 1) * stride < A.n_tiles) stage_tile(tile + (kNS - 1) * stride, (it + kNS - 1) % kNS); else cp_async_commit();
        cp_async_wait<kNS - 1>();
        __syncthreads();  // this tile's pixels (every thread's copies) and, first time round, the weight fragments are visible
        const int img = (int)(tile / per_img), r = (int)(tile - (int64_t)img * per_img);
        const int y0 = (r / A.tiles_x) * kTile, x0 = (r % A.tiles_x) * kTile;
        const uint32_t tbase = tiles + (uint32_t)buf * kStage + a_lane;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int ly = 2 * warp + half;   // output row of the tile = one m16 tile of 16 pixels along x
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = 0.f; acc[nt][1] = 0.f; acc[nt][2] = 0.f; acc[nt][3] = 0.f; }
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const uint32_t arow = tbase + (uint32_t)((ly * S + tap / 3) * kHalo + tap % 3) * kPitch;
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) {
                    uint32_t a[4];
                    ldmatrix_x4(arow + kc * 32, a);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        uint2 bf;
                        if constexpr (kRegB) bf = bfr[(tap * KC + kc) * NT + nt];
                        else bf = *reinterpret_cast<const uint2*>(s_wf + (((tap * KC + kc) * NT + nt) * 32 + lane) * 4);
                        mma16816<T>(acc[nt], a, bf.x, bf.y);
                    }
                }
            }
            // ---- epilogue: d0,d1 = (pixel g, channels 2t, 2t+1), d2,d3 = (pixel g + 8, same channels)
            const int y = y0 + ly;
            if (y < A.Ho) {
                T* orow = outp + (int64_t)img * A.on + (int64_t)y * A.oh;
#pragma unroll
                for (int hx = 0; hx < 2; ++hx) {
                    const int x = x0 + g + 8 * hx;
                    if (x < A.Wo) {
                        T* o = orow + (int64_t)x * A.ow + 2 * t4;
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) {
                            float v0 = acc[nt][2 * hx] + s_bias[8 * nt + 2 * t4], v1 = acc[nt][2 * hx + 1] + s_bias[8 * nt + 2 * t4 + 1];
                            if (A.act == 1) { v0 = silu_fast(v0); v1 = silu_fast(v1); }
                            else if (A.act == 2) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                            *reinterpret_cast<uint32_t*>(o + 8 * nt) = pack2<T>(v0, v1);
                        }
                    }
                }
            }
        }
        __syncthreads();  // the buffer is refilled by the next iteration's prefetch
    }
    cp_async_wait<0>();
}

template <typename T, int KC, int NT, int S>
static cudaError_t launch(const Args& A, cudaStream_t st) {
    constexpr int C = 16 * KC, N = 8 * NT, kHalo = (kTile - 1) * S + 3;
    const size_t smem = (size_t)9 * KC * NT * 32 * 4 * 2 + ((N * 4 + 15) & ~15) + (size_t)(S == 1 ? 3 : 2) * kHalo * kHalo * (C * 2 + 16);
    auto kern = conv3x3_mma_kernel<T, KC, NT, S>;
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    // resident CTAs per SM: by shared memory and by registers (the occupancy query knows both); more CTAs than that would run as a second wave
    static const int per_sm = [&] {  // one query per instantiation (shared memory is a compile-time constant of it)
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kThreads, smem) != cudaSuccess || n < 1) n = 1;
        return n > 4 ? 4 : n;
    }();
    int64_t grid = (int64_t)kSMs * per_sm;
    if (grid > A.n_tiles) grid = A.n_tiles;
    return launch_pdl(kern, dim3((unsigned)grid), dim3(kThreads), smem, st, A);
}

}  // namespace c3m
}  // namespace el

using namespace el;

extern "C" int el_conv3x3_mma_ok(int C, int N, int stride) {
    return (C == 16 || C == 32) && N > 0 && N % 8 == 0 && N <= c3m::kMaxN && (stride == 1 || stride == 2);
}

extern "C" int el_conv3x3_mma_fwd(const void* x, const int64_t xs_[4], int C, const float* w, const float* bias, void* out, const int64_t os_[4], int B,
                                  int H, int W, int N, int stride, int act, int dtype, void* stream) {
    if (!x || !xs_ || !w || !out || !os_ || B <= 0 || C <= 0 || H <= 0 || W <= 0 || N <= 0 || act < 0 || act > 2) return EL_ERR_ARG;
    if (dtype != EL_BF16 && dtype != EL_F16) return EL_ERR_UNSUPPORTED;
    if (!el_conv3x3_mma_ok(C, N, stride) || xs_[1] != 1 || os_[1] != 1 || !aligned16(x) || !aligned16(w) || ((uintptr_t)out & 3)) return EL_ERR_UNSUPPORTED;
    for (int i = 0; i < 4; ++i) {
        if (i != 1 && xs_[i] % 8) return EL_ERR_UNSUPPORTED;   // 16-byte cp.async sources
        if (i != 1 && os_[i] % 2) return EL_ERR_UNSUPPORTED;   // 4-byte stores
    }
    c3m::Args A{};
    A.x = x; A.xn = xs_[0]; A.xh = xs_[2]; A.xw = xs_[3];
    A.out = out; A.on = os_[0]; A.oh = os_[2]; A.ow = os_[3];
    A.w = w; A.bias = bias;
    A.B = B; A.H = H; A.W = W; A.C = C; A.N = N; A.act = act;
    A.Ho = (H - 1) / stride + 1; A.Wo = (W - 1) / stride + 1;   // (H + 2 - 3) / s + 1
    A.tiles_x = (int)ceil_div(A.Wo, c3m::kTile); A.tiles_y = (int)ceil_div(A.Ho, c3m::kTile);
    A.n_tiles = (int64_t)B * A.tiles_x * A.tiles_y;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaErrorInvalidValue;
    const int kc = C / 16, nt = N / 8;
#define EL_C3M_NT(TT, KK, SS)                                                                                                   \
    switch (nt) {                                                                                                               \
        case 1: e = c3m::launch<TT, KK, 1, SS>(A, st); break; case 2: e = c3m::launch<TT, KK, 2, SS>(A, st); break;             \
        case 3: e = c3m::launch<TT, KK, 3, SS>(A, st); break; case 4: e = c3m::launch<TT, KK, 4, SS>(A, st); break;             \
        case 5: e = c3m::launch<TT, KK, 5, SS>(A, st); break; case 6: e = c3m::launch<TT, KK, 6, SS>(A, st); break;             \
        case 7: e = c3m::launch<TT, KK, 7, SS>(A, st); break; default: e = c3m::launch<TT, KK, 8, SS>(A, st); break;            \
    }
#define EL_C3M(TT)                                                                  \
    do {                                                                            \
        if (kc == 1) { if (stride == 1) { EL_C3M_NT(TT, 1, 1) } else { EL_C3M_NT(TT, 1, 2) } } \
        else { if (stride == 1) { EL_C3M_NT(TT, 2, 1) } else { EL_C3M_NT(TT, 2, 2) } }         \
    } while (0)
    if (dtype == EL_BF16) EL_C3M(__nv_bfloat16); else EL_C3M(__half);
#undef EL_C3M
#undef EL_C3M_NT
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return EL_ERR_CUDA; }
    note_launches(1);
    return check_launch();
}
