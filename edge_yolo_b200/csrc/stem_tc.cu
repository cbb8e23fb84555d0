// Fused uint8 stem on the tensor cores (16-bit activations): preprocess (engine/predictor.py:117-135) + layer 0 of the yaml
// (cfg/models/11/yolo11-test.yaml:21: Conv(3, C0, k=3, s=2, p=1) + BatchNorm + SiLU, nn/modules/conv.py:41-60) as an im2col GEMM
//     out[pixel, co] = SiLU( sum_{ky,kx,ci} img[2*oy-1+ky, 2*ox-1+kx, ci] * (W[co,ci,ky,kx] / 255) + bias[co] )
// with K = 27 padded to 32.  uint8 pixel values are exact in bf16 / fp16, so the only rounding is that of the (folded, /255)
// weights to 16 bits -- the same rounding the reference's half / bf16 model applies to its weights.
//
// Why tensor cores for a 3-channel conv: on CUDA cores the 27 x C0 FMAs per output pixel make the kernel FMA-issue bound
// (3-register FFMA issues every second cycle per scheduler: ~300 us for the B = 64, 640^2 batch however the register tiles are
// shaped, see stem_conv_u8_kernel).  Here a thread only BUILDS its pixel's 64-byte im2col row (27 byte loads, exact u8 -> 16-bit
// conversion, four 16-byte stores into a 64-byte-swizzled K-major tile); one tcgen05.mma pair per 128-pixel tile does the math,
// and the epilogue (tcgen05.ld, bias, SiLU, 16-bit pack) writes each pixel's C0 channels contiguously: HBM-bound.
#include <type_traits>

#include "el_common.cuh"

namespace el {
namespace stemtc {

constexpr int kTW = 64, kTH = 2;                   // output patch per tile: 64 x 2 pixels = 128 GEMM rows
constexpr int kInRows = 2 * kTH + 1;               // 5 input rows
constexpr int kInBytes = (2 * kTW + 1) * 3;        // 387 bytes of each input row
constexpr int kInWords = (kInBytes + 3) / 4 + 1;   // 98 aligned words cover them at any alignment
constexpr int kInPitch = 100;                      // words per staged row

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
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
// K-major, 64-byte rows, 64-byte swizzle: SBO = 8 rows x 64 B, layout type 4 (cute::UMMA::LayoutType::SWIZZLE_64B)
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ uint32_t umma_idesc(int fmt, int M, int N) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// two small non-negative integers (exact in 16 bits) / two floats -> packed 16-bit pair
template <typename T> __device__ __forceinline__ uint32_t pack2f(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2f<__nv_bfloat16>(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
template <> __device__ __forceinline__ uint32_t pack2f<__half>(float a, float b) {
    __half2 t = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float u8_to_f(uint32_t b) { return __uint_as_float(0x4B000000u | b) - 8388608.f; }

template <typename T, int C0>
__global__ void __launch_bounds__(128) stem_tc_kernel(const uint8_t* __restrict__ src, const float* __restrict__ w, const float* __restrict__ bias,
                                                      T* __restrict__ dst, Strides4 ds, int H, int W, int tiles_x, int tiles_y, int total) {
    __shared__ __align__(1024) unsigned char s_a[128 * 64];      // im2col tile: 128 rows x 32 K, 64-byte swizzle
    __shared__ __align__(1024) unsigned char s_wt[C0 * 64];      // weights: C0 rows x 32 K, same layout
    __shared__ __align__(16) uint32_t s_in[kInRows * kInPitch];  // raw bytes of the input patch
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
    constexpr uint32_t kCols = C0 < 32 ? 32 : C0;
    const uint32_t bar = smem_addr(&s_bar);

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem)), "n"(kCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid < C0) {
        // Operand form (round 2): the im2col rows are fp16 values 1024 + v (bit pattern 0x6400 | v: exact, and assembled from the raw bytes by
        // one PRMT per pair -- no u8 -> float -> 16-bit conversion chain: that chain was ~130 of the kernel's ~290 instructions per pixel), so
        // both MMA operands are fp16 whatever the activation type T.  K index = 10 * ky + (kx * 3 + ci), slot 10 * ky + 9 is padding (weight
        // 0), slots 30 / 31 carry the bias: the rows hold 1.0 there and the weight tile holds, as an fp16 hi + lo pair,
        //     bias' = bias / 2 - 1024 * sum_k W'_k        (W'_k = fp16(W_k / 2): the rounded weights, so the 1024 offsets cancel exactly)
        // The accumulator is then h = (conv + bias) / 2, the argument of SiLU(x) = h + h * tanh(h): one MUFU.TANH + one FFMA per channel.
        float k32[32];
        float wsum = 0.f;
#pragma unroll
        for (int k = 0; k < 30; ++k) {
            const int ky = k / 10, j = k - 10 * ky, kx = j / 3, ci = j - 3 * kx;
            float v = 0.f;
            if (j < 9) {
                v = __half2float(__float2half_rn(0.5f * __ldg(w + tid * 27 + ci * 9 + ky * 3 + kx)));
                wsum += v;
            }
            k32[k] = v;
        }
        const float hb = 0.5f * __ldg(bias + tid) - 1024.f * wsum;
        k32[30] = __half2float(__float2half_rn(hb));
        k32[31] = hb - k32[30];
        const uint32_t swz = ((uint32_t)tid >> 1) & 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack2f<__half>(k32[8 * j], k32[8 * j + 1]); o.y = pack2f<__half>(k32[8 * j + 2], k32[8 * j + 3]);
            o.z = pack2f<__half>(k32[8 * j + 4], k32[8 * j + 5]); o.w = pack2f<__half>(k32[8 * j + 6], k32[8 * j + 7]);
            *reinterpret_cast<uint4*>(s_wt + tid * 64 + ((j ^ swz) << 4)) = o;
        }
    }
    pdl_launch_dependents();
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = umma_idesc(0, 128, C0);  // A and B are fp16 (see the weight prologue), D fp32
    const int py = tid >> 6, px = tid & 63;        // this thread's pixel inside the 64 x 2 patch = GEMM row tid
    const uint32_t a_swz = ((uint32_t)tid >> 1) & 3;
    const int row_bytes = W * 3;
    constexpr int V = 8;

    // the input patch of tile t+1 is requested (4 words per thread, held in registers) right after tile t's patch has been staged,
    // so its global-memory latency hides behind tile t's im2col build, MMA and epilogue
    // Tile coordinates (image n, tile row ry, tile column rx) advance by gridDim.x tiles per iteration as a mixed-radix add: the divisions
    // by the run-time tile counts (t / per_img, r / tiles_x, r % tiles_x, in the prefetch and again in the loop body) were ~100 of the
    // kernel's 449 instructions per pixel (ncu source page, round 2: I2F.RP / MUFU.RCP / F2I / IABS sequences, issue slots 73 % busy).
    struct TileC { int n, ry, rx; };
    const int per_img = tiles_x * tiles_y;
    TileC cur, step;
    {
        const int t0 = blockIdx.x, d = gridDim.x;
        cur.n = t0 / per_img; const int r0 = t0 - cur.n * per_img; cur.ry = r0 / tiles_x; cur.rx = r0 - cur.ry * tiles_x;
        step.n = d / per_img; const int rd = d - step.n * per_img; step.ry = rd / tiles_x; step.rx = rd - step.ry * tiles_x;
    }
    auto advance = [&](TileC c) {
        c.rx += step.rx;
        if (c.rx >= tiles_x) { c.rx -= tiles_x; ++c.ry; }
        c.ry += step.ry;
        if (c.ry >= tiles_y) { c.ry -= tiles_y; ++c.n; }
        c.n += step.n;
        return c;
    };
    uint32_t pre[4];
    auto fetch = [&](const TileC& c) {
        const int n = c.n;
        const int oy0 = c.ry * kTH, ox0 = c.rx * kTW;
        const uint8_t* img = src + (int64_t)n * H * row_bytes;
        const int iy0 = 2 * oy0 - 1, seg0 = (2 * ox0 - 1) * 3;
        const int w_lo = (seg0 - (seg0 & 3)) >> 2;  // floor(seg0 / 4)
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // aligned 32-bit loads (W % 4 == 0: words never straddle image rows)
            const int i = tid + 128 * k;
            const int rr = i / kInWords, wi = i - rr * kInWords;
            const int iy = iy0 + rr, byte0 = 4 * (w_lo + wi);
            pre[k] = 0;
            if (i < kInRows * kInWords && iy >= 0 && iy < H && byte0 >= 0 && byte0 < row_bytes)
                pre[k] = __ldg(reinterpret_cast<const uint32_t*>(img + (int64_t)iy * row_bytes + byte0));
        }
    };
    if ((int)blockIdx.x < total) fetch(cur);
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int n = cur.n;
        const int oy0 = cur.ry * kTH, ox0 = cur.rx * kTW;
        const TileC nxt = advance(cur);
        const int seg0 = (2 * ox0 - 1) * 3;
        // ---- stage the 5 x 387-byte input patch
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = tid + 128 * k;
            if (i < kInRows * kInWords) s_in[(i / kInWords) * kInPitch + (i % kInWords)] = pre[k];
        }
        __syncthreads();
        if (t + (int)gridDim.x < total) fetch(nxt);
        // ---- im2col row of this thread's pixel: 3 x 9 bytes -> fp16 pairs (0x6400 | byte): per image row three aligned words, two funnel
        //      shifts to the byte the pixel starts at, five PRMTs
        {
            uint32_t pr[16];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const uint32_t boff = (uint32_t)((2 * py + ky) * (kInPitch * 4) + (seg0 & 3) + 6 * px);
                const uint32_t* pw = s_in + (boff >> 2);
                const uint32_t sh = (boff & 3) * 8;
                const uint32_t w0 = pw[0], w1 = pw[1], w2 = pw[2];
                const uint32_t r0 = __funnelshift_r(w0, w1, sh), r1 = __funnelshift_r(w1, w2, sh), r2 = w2 >> sh;  // bytes 0-3, 4-7, 8
                pr[5 * ky + 0] = __byte_perm(r0, 0x64646464u, 0x4140);
                pr[5 * ky + 1] = __byte_perm(r0, 0x64646464u, 0x4342);
                pr[5 * ky + 2] = __byte_perm(r1, 0x64646464u, 0x4140);
                pr[5 * ky + 3] = __byte_perm(r1, 0x64646464u, 0x4342);
                pr[5 * ky + 4] = __byte_perm(r2, 0x64646464u, 0x4440);  // (byte 8, padding slot: weight 0)
            }
            pr[15] = 0x3C003C00u;  // 1.0, 1.0: the two bias columns
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(s_a + tid * 64 + ((j ^ a_swz) << 4)) = make_uint4(pr[4 * j], pr[4 * j + 1], pr[4 * j + 2], pr[4 * j + 3]);
        }
        proxy_fence();  // generic-proxy writes of the A tile -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            umma(tmem, umma_desc_sw64(smem_addr(s_a)), umma_desc_sw64(smem_addr(s_wt)), idesc, 0u);
            umma(tmem, umma_desc_sw64(smem_addr(s_a) + 32), umma_desc_sw64(smem_addr(s_wt) + 32), idesc, 1u);
            umma_commit(bar);
        }
        mbar_wait(bar, (uint32_t)it & 1);
        tc_fence_after();
        // ---- epilogue: thread <-> pixel (TMEM lane = GEMM row = tid): C0 fp32 -> bias + SiLU -> 16-bit, contiguous per pixel
        const int oy = oy0 + py, ox = ox0 + px;
        const bool valid = oy < H / 2 && ox < W / 2;
        T* q = dst + (int64_t)n * ds.n + (int64_t)oy * ds.h + (int64_t)ox * ds.w;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c0 = 0; c0 < C0; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + c0, v);
            if (valid) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    float f[V];
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const float h = __uint_as_float(v[8 * g + e]);  // (conv + bias) / 2 straight from the accumulator
                        float th;
                        asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
                        f[e] = fmaf(h, th, h);  // SiLU(x) = h + h * tanh(h), h = x / 2
                    }
                    *reinterpret_cast<uint4*>(q + c0 + 8 * g) = pack<T>(f);
                }
            }
        }
        tc_fence_before();
        __syncthreads();  // TMEM accumulator, A tile and input patch are reused by the next tile
        cur = nxt;
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kCols));
}

}  // namespace stemtc

// 16-bit path of el_stem_conv_u8 (called from epilogue.cu)
int stem_tc_launch(const uint8_t* src, const float* w, const float* bias, void* dst, Strides4 ds, int B, int C0, int H, int W, int dtype, cudaStream_t st) {
    const int tiles_x = (int)ceil_div(W / 2, stemtc::kTW), tiles_y = (int)ceil_div(H / 2, stemtc::kTH);
    const int64_t total = (int64_t)B * tiles_x * tiles_y;
    if (total >= (1ll << 31)) return EL_ERR_UNSUPPORTED;
    const int grid = (int)(total < (int64_t)kSMs * 8 ? total : (int64_t)kSMs * 8);  // persistent: up to 8 CTAs (32 warps) per SM
#define EL_STEM_TC(TT, CC) stemtc::stem_tc_kernel<TT, CC><<<grid, 128, 0, st>>>(src, w, bias, (TT*)dst, ds, H, W, tiles_x, tiles_y, (int)total)
    if (dtype == EL_BF16) {
        if (C0 == 16) EL_STEM_TC(__nv_bfloat16, 16); else if (C0 == 32) EL_STEM_TC(__nv_bfloat16, 32); else if (C0 == 64) EL_STEM_TC(__nv_bfloat16, 64);
        else return EL_ERR_UNSUPPORTED;
    } else if (dtype == EL_F16) {
        if (C0 == 16) EL_STEM_TC(__half, 16); else if (C0 == 32) EL_STEM_TC(__half, 32); else if (C0 == 64) EL_STEM_TC(__half, 64);
        else return EL_ERR_UNSUPPORTED;
    } else {
        return EL_ERR_UNSUPPORTED;
    }
#undef EL_STEM_TC
    return EL_OK;
}

}  // namespace el
