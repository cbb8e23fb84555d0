// Depthwise 3 x 3 -> pointwise 1 x 1 in ONE kernel (inference engine, NHWC 16-bit):
//     d[p, c]   = dw_act( sum_{ky,kx} X[p + (ky-1, kx-1), c] * Wd[c, ky, kx] + dw_bias[c] )      rounded to the activation type
//     out[p, n] = act( sum_c d[p, c] * Wp[n, c] + bias[n] )
// Replaces the pair el_dwconv_fwd -> el_pwconv_fwd for
//   DSConv.forward (k = 3)          nn/modules/conv.py:100-104  (dw -> pw -> BatchNorm -> SiLU; cv1 of every DSBottleneck, block.py:1494-1503)
//   DWConv(x, x, 3) -> Conv(x, c, 1) nn/modules/head.py:66-71    (the two stages of every class tower of the Detect / GFL head)
// Why: the depthwise output is written to HBM by one kernel and read back by the next (2 * B * C * H * W * e bytes and a launch per site;
// the 17 k = 3 depthwise launches of EdgeLine-n cost ~0.3 ms of a 2.65 ms step, bench.py `kernels.dwconv`), although the GEMM that consumes it
// only needs it as its A operand in shared memory.  Here the depthwise result never leaves the SM:
//   warp 0              TMA producer (one elected lane issues; the loop is warp-uniform): per (tile, 64-channel chunk) ONE 4-D box (channels, 22 x, 8 y,
//                       1 image) = the 20 x 6 output tile + halo, out-of-image pixels zero-filled by the TMA unit (= the convolution's padding), into a
//                       ring of patch stages;
//   warps 10-19         depthwise producers: thread <-> (channel vector, output column, strip of 3 rows); 15 LDS.128 of the patch feed 27 taps as
//                       packed fp32x2 FMAs (same evaluation order as dwconv3_tma_kernel), optional bias + SiLU, one rounding to the activation
//                       type, and the 16-byte result goes straight into the K-major SWIZZLED A tile of the GEMM (row = pixel ly * 20 + lx,
//                       chunk cv ^ swizzle(row)): what el_pwconv_fwd would have fetched by TMA; fence.proxy.async + mbarrier hand-over.
//                       A partial last chunk (80 channels = 64 + 16) is mapped (vector, column, ONE row) over all warps;
//   warp 1              tcgen05.mma M128 x N x K16 (elected lane) from the A tile and the resident weight tiles, accumulators double buffered in TMEM;
//   warps 2-9           two epilogue groups of four warps (group g drains accumulator g, tiles alternate; as in conv3x3_halo.cu): tcgen05.ld, bias,
//                       SiLU / ReLU, 16-bit pack into a swizzled staging tile, and ONE 4-D TMA store per 64 output channels: box (channels, 20 x,
//                       6 y, 1 image), clipped at the image border.
// Rows 120..127 of the A tile are never written (zero-filled once) and never stored (an MMA row only depends on its own A row).
// Roofline: HBM, algorithmic bytes B * H * W * (C + N) * e (the depthwise tensor does not exist).
#include <cuda.h>

#include <type_traits>

#include "el_common.cuh"

namespace el {
namespace ds {

constexpr int kTW = 20, kTH = 6, kRT = 3;     // output tile (120 of the 128 accumulator rows), rows per depthwise thread
constexpr int kPW = kTW + 2, kPH = kTH + 2;   // patch = tile + halo
constexpr int kDwThreads = 320;               // (<= 8 channel vectors) x 20 columns x 2 strips
constexpr int kEpiGroups = 2;                 // epilogue groups of four warps; group g drains accumulator g (tiles alternate)
constexpr int kEpiThreads = 128 * kEpiGroups;
constexpr int kThreads = 64 + kEpiThreads + kDwThreads;    // TMA warp, MMA warp, 8 epilogue warps, 10 depthwise warps
constexpr int kMaxSP = 6, kMaxSA = 4;

struct Args {
    CUtensorMap src_map, out_map;
    const void* wpk;        // ops.pack_pw_weight(w, [C], n_tile = n_pad): per K chunk one swizzled K-major tile [n_pad][row bytes], padded to 1 KiB
    const float* bias;      // [N] or null (folded BatchNorm of the pointwise conv)
    const float* dw_w;      // [9][C] fp32, tap-major (ops.pack_dw_weight)
    const float* dw_bias;   // [C] or null
    int C, N, n_pad, chunks, rb, cvl_shift, ob, dw_act, SP, SA;
    int H, W, B, tiles_x, tiles_y;
    int64_t n_tiles;
    uint32_t w_bytes, w_tile_bytes, patch_bytes, a_bytes, tmem_cols;
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst), "l"(map),
                 "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2),
                 "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// K-major swizzled shared-memory matrix descriptor (see pwconv.cu): SBO = 8 rows x row bytes, layout 2 / 4 / 6 = 128 / 64 / 32-byte swizzle
__device__ __forceinline__ uint64_t umma_desc_sw(uint32_t saddr, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((8 * row_bytes) >> 4) << 32) | (1ull << 46) | (layout << 61);
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
// one lane of a converged warp: the loops around it stay warp-uniform (descriptors / coordinates in uniform registers, no waterfall around UTMALDG / UTCHMMA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier(int g) {  // the four warps of epilogue group g
    if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
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
__device__ __forceinline__ float silu_tanh(float v) {  // x * sigmoid(x) = h + h * tanh(h), h = x / 2 (one MUFU; 16-bit outputs)
    const float h = 0.5f * v;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

// Depthwise 3 x 3 of RT vertically adjacent outputs of one channel vector (8 channels): (RT + 2) x 3 LDS.128 of the staged patch, 9 * RT * 4 packed
// fp32x2 FMAs in the evaluation order of dwconv3_tma_kernel (kx outer, ky inner, bias first), optional activation, one rounding to T.
template <typename T, int RT>
__device__ __forceinline__ void dw3_strip(uint32_t base, uint32_t row_pitch, uint32_t kx_pitch, const float* s_dw, const float* s_dwb, int Cpad, int ch,
                                          bool silu, bool relu, uint4 (&o)[RT]) {
    constexpr int V = 8;
    f32x2 acc[RT][V / 2];
    {
        const float4 b0 = *reinterpret_cast<const float4*>(s_dwb + ch), b1 = *reinterpret_cast<const float4*>(s_dwb + ch + 4);
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            acc[r][0] = pack_f32x2(b0.x, b0.y); acc[r][1] = pack_f32x2(b0.z, b0.w);
            acc[r][2] = pack_f32x2(b1.x, b1.y); acc[r][3] = pack_f32x2(b1.z, b1.w);
        }
    }
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
        f32x2 v[RT + 2][V / 2];
#pragma unroll
        for (int j = 0; j < RT + 2; ++j) {
            float f[V];
            unpack<T>(lds128(base + (uint32_t)j * row_pitch + (uint32_t)kx * kx_pitch), f);
#pragma unroll
            for (int e = 0; e < V / 2; ++e) v[j][e] = pack_f32x2(f[2 * e], f[2 * e + 1]);
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            // taps are stored [tap][half][vector][4]: the lanes of a warp (8 vectors) read 128 contiguous bytes per half (no bank conflict)
            const float4 w0 = *reinterpret_cast<const float4*>(s_dw + (ky * 3 + kx) * Cpad + (ch >> 1));
            const float4 w1 = *reinterpret_cast<const float4*>(s_dw + (ky * 3 + kx) * Cpad + (Cpad >> 1) + (ch >> 1));
            const f32x2 w2[V / 2] = {pack_f32x2(w0.x, w0.y), pack_f32x2(w0.z, w0.w), pack_f32x2(w1.x, w1.y), pack_f32x2(w1.z, w1.w)};
#pragma unroll
            for (int r = 0; r < RT; ++r)
#pragma unroll
                for (int e = 0; e < V / 2; ++e) acc[r][e] = fma_f32x2(v[r + ky][e], w2[e], acc[r][e]);
        }
    }
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        float f[V];
#pragma unroll
        for (int e = 0; e < V / 2; ++e) unpack_f32x2(acc[r][e], f[2 * e], f[2 * e + 1]);
        if (silu) {
#pragma unroll
            for (int e = 0; e < V; ++e) f[e] = silu_tanh(f[e]);
        } else if (relu) {
#pragma unroll
            for (int e = 0; e < V; ++e) f[e] = fmaxf(f[e], 0.f);
        }
        o[r] = pack<T>(f);
    }
}

template <typename T, int ACT>
__global__ void __launch_bounds__(kThreads, 1) dsconv3_tc_kernel(const __grid_constant__ Args A) {
    extern __shared__ __align__(1024) unsigned char sm_raw[];
    const uint32_t sbase = (smem_addr(sm_raw) + 1023u) & ~1023u;
    unsigned char* sm = sm_raw + (sbase - smem_addr(sm_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int SP = A.SP, SA = A.SA, nch = A.chunks, rb = A.rb, n_pad = A.n_pad, ob = A.ob;
    const int CB = rb / 2, Cpad = nch * CB;
    // shared memory map: [weights][A ring][2 staging tiles 128 x ob][patch ring][depthwise taps + bias][bias][barriers]
    const uint32_t off_a = (A.w_bytes + 1023u) & ~1023u;
    const uint32_t off_stage = off_a + (uint32_t)SA * A.a_bytes;
    const uint32_t staging_bytes = 128u * ob * 2;
    const uint32_t off_patch = off_stage + 2 * kEpiGroups * staging_bytes;
    const uint32_t off_dw = off_patch + (uint32_t)SP * A.patch_bytes;
    float* s_dw = reinterpret_cast<float*>(sm + off_dw);   // [9][2 halves][Cpad / 8 vectors][4], zero past C
    float* s_dwb = s_dw + 9 * Cpad;                        // [Cpad]
    float* s_bias = s_dwb + Cpad;                          // [n_pad + 64]: the last store box may overhang n_pad
    const uint32_t off_bar = off_dw + (uint32_t)(10 * Cpad + n_pad + 64) * 4;
    const uint32_t bar_w = sbase + off_bar;
    const uint32_t bar_acc_full = bar_w + 8;                    // [2]
    const uint32_t bar_acc_empty = bar_w + 24;                  // [2]
    const uint32_t bar_pfull = bar_w + 40;                      // [kMaxSP] patch landed (TMA)
    const uint32_t bar_pempty = bar_pfull + 8 * kMaxSP;         // [kMaxSP] patch read by the depthwise warps
    const uint32_t bar_afull = bar_pempty + 8 * kMaxSP;         // [kMaxSA] A tile written
    const uint32_t bar_aempty = bar_afull + 8 * kMaxSA;         // [kMaxSA] A tile consumed by the MMAs
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + off_bar + 40 + 16 * kMaxSP + 16 * kMaxSA);

    const int64_t first = blockIdx.x;
    const int my_tiles = first < A.n_tiles ? (int)((A.n_tiles - first + gridDim.x - 1) / gridDim.x) : 0;
    const int per_img = A.tiles_x * A.tiles_y;
    constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;

    pdl_launch_dependents();
    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&A.src_map) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&A.out_map) : "memory");
        mbar_init(bar_w, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(bar_acc_full + 8 * b, 1); mbar_init(bar_acc_empty + 8 * b, 128); }
        for (int s = 0; s < SP; ++s) { mbar_init(bar_pfull + 8 * s, 1); mbar_init(bar_pempty + 8 * s, kDwThreads / 32); }
        for (int s = 0; s < SA; ++s) { mbar_init(bar_afull + 8 * s, kDwThreads / 32); mbar_init(bar_aempty + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_w, A.w_bytes);
        const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(A.wpk);
        for (uint32_t o = 0; o < A.w_bytes; o += 32768) bulk_g2s(sbase + o, wsrc + o, min(32768u, A.w_bytes - o), bar_w);
    }
    for (int i = tid; i < 9 * Cpad; i += kThreads) {
        const int tap = i / Cpad, c = i - tap * Cpad;
        s_dw[tap * Cpad + ((c >> 2) & 1) * (Cpad >> 1) + (c >> 3) * 4 + (c & 3)] = c < A.C ? __ldg(A.dw_w + tap * A.C + c) : 0.f;
    }
    for (int i = tid; i < Cpad; i += kThreads) s_dwb[i] = (A.dw_bias && i < A.C) ? __ldg(A.dw_bias + i) : 0.f;
    for (int i = tid; i < n_pad + 64; i += kThreads) s_bias[i] = (A.bias && i < A.N) ? __ldg(A.bias + i) : 0.f;
    // the A tiles start as zeros: the columns a partial last chunk never writes and the rows 120..127 must hold finite numbers
    for (uint32_t o = (uint32_t)tid * 16; o < (uint32_t)SA * A.a_bytes; o += kThreads * 16) sts128(sbase + off_a + o, make_uint4(0, 0, 0, 0));
    proxy_fence();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(s_tmem)), "r"(A.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    pdl_wait();  // everything above touched only parameters; the producer of x is complete from here on

    if (warp == 0) {
        // ------------------------------------------------------------------------------------ TMA producer: one haloed patch per (tile, chunk)
        {
            const bool leader = elect_one();
            int s = 0, use = 0;
            for (int tl = 0; tl < my_tiles; ++tl) {
                const int64_t tile = first + (int64_t)tl * gridDim.x;
                const int img = (int)(tile / per_img), r = (int)(tile % per_img);
                const int y0 = (r / A.tiles_x) * kTH - 1, x0 = (r % A.tiles_x) * kTW - 1;
                for (int c = 0; c < nch; ++c) {
                    if (use > 0) mbar_wait(bar_pempty + 8 * s, (uint32_t)(use - 1) & 1);
                    if (leader) {
                        mbar_expect_tx(bar_pfull + 8 * s, A.patch_bytes);
                        tma_load_4d(sbase + off_patch + (uint32_t)s * A.patch_bytes, &A.src_map, c * CB, x0, y0, img, bar_pfull + 8 * s);
                    }
                    __syncwarp();
                    if (++s == SP) { s = 0; ++use; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------------------------ MMA issuer
        {
            const bool leader = elect_one();
            const uint32_t idesc = umma_idesc(kFmt, 128, n_pad);
            mbar_wait(bar_w, 0);
            int a = 0;
            uint32_t a_par = 0;
            const int ksteps = rb / 32;
            for (int tl = 0; tl < my_tiles; ++tl) {
                const int b = tl & 1, ub = tl >> 1;
                if (ub > 0) mbar_wait(bar_acc_empty + 8 * b, (uint32_t)(ub - 1) & 1);
                tc_fence_after();
                const uint32_t d = tmem + (uint32_t)b * n_pad;
                for (int c = 0; c < nch; ++c) {
                    mbar_wait(bar_afull + 8 * a, a_par);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw(sbase + off_a + (uint32_t)a * A.a_bytes, rb), db = umma_desc_sw(sbase + (uint32_t)c * A.w_tile_bytes, rb);
                    if (leader) {
                        for (int ks = 0; ks < ksteps; ++ks)  // one MMA per 16 channels = 32 bytes along K inside the swizzle atom (+2 in the start field)
                            umma(d, da + (uint64_t)(2 * ks), db + (uint64_t)(2 * ks), idesc, (c > 0 || ks > 0) ? 1u : 0u);
                        umma_commit(bar_aempty + 8 * a);
                        if (c == nch - 1) umma_commit(bar_acc_full + 8 * b);
                    }
                    __syncwarp();
                    if (++a == SA) { a = 0; a_par ^= 1; }
                }
            }
        }
    } else if (warp < 2 + 4 * kEpiGroups) {
        // ------------------------------------------------------------------------------------ epilogue warps
        // two groups of four warps, group g owning accumulator g and the tiles tl = g (mod 2): a tile's epilogue is a latency chain, and with one
        // CTA per SM only another tile's epilogue can overlap it (the 3x3 halo kernel gained 1.25x from the same split)
        const int g = (warp - 2) >> 2, q = warp & 3, row = q * 32 + lane, et = (tid - 64) & 127;
        const int rbo = ob * 2;
        const uint32_t swz = ((uint32_t)(row * rbo) >> 7) & (uint32_t)(rbo / 16 - 1);
        int sub = 0;
        for (int tl = g; tl < my_tiles; tl += kEpiGroups) {
            const int b = tl & 1;
            const int64_t tile = first + (int64_t)tl * gridDim.x;
            const int img = (int)(tile / per_img), r = (int)(tile % per_img);
            const int y0 = (r / A.tiles_x) * kTH, x0 = (r % A.tiles_x) * kTW;
            mbar_wait(bar_acc_full + 8 * b, (uint32_t)(tl >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + (uint32_t)b * n_pad + ((uint32_t)(q * 32) << 16);
            for (int c0 = 0; c0 < A.N; c0 += ob, ++sub) {
                const uint32_t stg = sbase + off_stage + (uint32_t)(2 * g + (sub & 1)) * staging_bytes;
                if (et == 0) bulk_wait_read<1>();  // the store that last read this staging buffer (two uses ago) is done with it
                epi_barrier(g);
                const int jmax = min(ob, n_pad - c0);  // the last box of an 80-channel conv holds 16 real columns: skip the other 48
                for (int j0 = 0; j0 < jmax; j0 += 32) {  // two 16-column TMEM loads in flight per wait
                    uint32_t v[2][16];
                    tmem_ld16_nowait(taddr + c0 + j0, v[0]);
                    if (j0 + 16 < jmax) tmem_ld16_nowait(taddr + c0 + j0 + 16, v[1]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        if (j0 + jj * 16 >= jmax) break;
                        const int j = j0 + jj * 16;
                        float f[16];
                        const float4* b4 = reinterpret_cast<const float4*>(s_bias + c0 + j);
#pragma unroll
                        for (int e4 = 0; e4 < 4; ++e4) {
                            const float4 bb = b4[e4];
                            f[4 * e4] = __uint_as_float(v[jj][4 * e4]) + bb.x; f[4 * e4 + 1] = __uint_as_float(v[jj][4 * e4 + 1]) + bb.y;
                            f[4 * e4 + 2] = __uint_as_float(v[jj][4 * e4 + 2]) + bb.z; f[4 * e4 + 3] = __uint_as_float(v[jj][4 * e4 + 3]) + bb.w;
                        }
                        if (ACT == 1) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) f[e] = silu_tanh(f[e]);
                        } else if (ACT == 2) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) f[e] = fmaxf(f[e], 0.f);
                        }
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint4 o;
                            o.x = pack2<T>(f[8 * h], f[8 * h + 1]); o.y = pack2<T>(f[8 * h + 2], f[8 * h + 3]);
                            o.z = pack2<T>(f[8 * h + 4], f[8 * h + 5]); o.w = pack2<T>(f[8 * h + 6], f[8 * h + 7]);
                            const uint32_t chunk = (uint32_t)(j / 8 + h) ^ swz;
                            sts128(stg + (uint32_t)row * rbo + chunk * 16, o);
                        }
                    }
                }
                if (c0 + ob >= A.N) {  // last box: the accumulator has been read completely
                    tc_fence_before();
                    mbar_arrive(bar_acc_empty + 8 * b);
                }
                proxy_fence();  // generic-proxy writes of the staging tile -> visible to the TMA store (async proxy)
                epi_barrier(g);
                if (et == 0) {  // rows 0..119 of the staging tile are the 20 x 6 box; rows / columns past the image are clipped
                    tma_store_4d(&A.out_map, c0, x0, y0, img, stg);
                    bulk_commit();
                }
            }
        }
        if (et == 0) bulk_wait_read<0>();
    } else {
        // ------------------------------------------------------------------------------------ depthwise producers
        constexpr int V = 8;
        const int dt = tid - (64 + kEpiThreads);
        const int CVL = 1 << A.cvl_shift;
        const int cvl = dt & (CVL - 1), rest = dt >> A.cvl_shift;
        const int xl = rest % kTW, strip = rest / kTW;
        const bool active = strip < kTH / kRT;     // narrow chunks (16 / 32 channels) leave the upper warps without work: they only keep the barrier counts
        const uint32_t row_pitch = (uint32_t)kPW * CVL * 16, kx_pitch = (uint32_t)CVL * 16;
        const uint32_t my_off = (uint32_t)(strip * kRT) * row_pitch + (uint32_t)(xl * CVL + cvl) * 16;
        const int m0 = strip * kRT * kTW + xl;     // A-tile row of this thread's first output pixel
        // A partial last chunk (80 channels: 64 + 16) has nv < 8 real channel vectors.  Mapping its work like a full chunk would leave 8 - nv of every
        // 8 lanes computing zeros for a whole chunk's time; instead thread <-> (vector < nv, column, ONE row), so the chunk costs nv / 8 of a full one
        // spread over all warps.  Its unwritten A columns keep whatever finite values an earlier chunk (or the zero fill at kernel start) left
        // there: they meet zero weight columns.
        const int nv = (A.C - (nch - 1) * CB) / V;              // real vectors of the last chunk (== CVL when C is a multiple of the chunk)
        const bool tail_remap = nv < CVL;
        const int t_cv = dt % nv, t_rest = dt / nv, t_xl = t_rest % kTW, t_ly = t_rest / kTW;
        const bool t_active = t_ly < kTH;
        const uint32_t t_off = (uint32_t)t_ly * row_pitch + (uint32_t)(t_xl * CVL + t_cv) * 16;
        const int t_m = t_ly * kTW + t_xl;
        const bool dw_silu = A.dw_act == 1, dw_relu = A.dw_act == 2;
        int s = 0, a = 0;
        uint32_t p_par = 0, a_par = 1;   // parities: patch use, and the A slot's PREVIOUS use (what its empty barrier is waited on)
        bool a_wrapped = false;
        for (int tl = 0; tl < my_tiles; ++tl) {
            for (int c = 0; c < nch; ++c) {
                const bool tail = tail_remap && c == nch - 1;
                mbar_wait(bar_pfull + 8 * s, p_par);
                const uint32_t pbase = sbase + off_patch + (uint32_t)s * A.patch_bytes;
                uint4 o[kRT];
                if (!tail) {
                    if (active) dw3_strip<T, kRT>(pbase + my_off, row_pitch, kx_pitch, s_dw, s_dwb, Cpad, c * CB + cvl * V, dw_silu, dw_relu, o);
                } else if (t_active) {
                    uint4 o1[1];
                    dw3_strip<T, 1>(pbase + t_off, row_pitch, kx_pitch, s_dw, s_dwb, Cpad, c * CB + t_cv * V, dw_silu, dw_relu, o1);
                    o[0] = o1[0];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_pempty + 8 * s);  // this warp has read everything it needs from the patch
                if (a_wrapped) mbar_wait(bar_aempty + 8 * a, a_par);
                const uint32_t a_base = sbase + off_a + (uint32_t)a * A.a_bytes;
                if (!tail) {
                    if (active) {
#pragma unroll
                        for (int r = 0; r < kRT; ++r) {
                            const uint32_t m = (uint32_t)(m0 + r * kTW);
                            const uint32_t sw = ((m * (uint32_t)rb) >> 7) & (uint32_t)(rb / 16 - 1);
                            sts128(a_base + m * (uint32_t)rb + (((uint32_t)cvl ^ sw) << 4), o[r]);
                        }
                    }
                } else if (t_active) {
                    const uint32_t m = (uint32_t)t_m;
                    const uint32_t sw = ((m * (uint32_t)rb) >> 7) & (uint32_t)(rb / 16 - 1);
                    sts128(a_base + m * (uint32_t)rb + (((uint32_t)t_cv ^ sw) << 4), o[0]);
                }
                proxy_fence();  // generic-proxy writes of the A tile -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_afull + 8 * a);
                if (++s == SP) { s = 0; p_par ^= 1; }
                if (++a == SA) { a = 0; a_par ^= 1; a_wrapped = true; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(A.tmem_cols));
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
// (channels, x, y, image) view of an NHWC activation; swizzle 0 = none, else the box row bytes (32 / 64 / 128)
static bool make_map4(CUtensorMap* map, const void* base, int channels, int W, int H, int B, const int64_t st[4] /* n, c, h, w elements */, int box_c,
                      int box_x, int box_y, int swizzle_bytes, int dtype) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)channels, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)st[3] * 2, (cuuint64_t)st[2] * 2, (cuuint64_t)(B > 1 ? st[0] : (int64_t)H * st[2]) * 2};
    const cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_x, (cuuint32_t)box_y, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    return fn(map, dtype == EL_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box,
              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// shared-memory plan of a (C, N) site; false = not covered
static bool plan(int C, int N, Args& A, size_t& smem) {
    if (C <= 0 || N <= 0 || C % 8 || N % 8 || N > 256) return false;
    if (C == 16) { A.rb = 32; A.cvl_shift = 1; }
    else if (C == 32) { A.rb = 64; A.cvl_shift = 2; }
    else if (C >= 64) { A.rb = 128; A.cvl_shift = 3; }
    else return false;
    const int CB = A.rb / 2;
    A.chunks = (int)ceil_div(C, CB);
    A.n_pad = (int)ceil_div(N, 16) * 16;
    A.w_tile_bytes = ((uint32_t)A.n_pad * A.rb + 1023u) & ~1023u;
    A.w_bytes = (uint32_t)A.chunks * A.w_tile_bytes;
    A.patch_bytes = (uint32_t)kPH * kPW * CB * 2;   // 8 * 22 * CB * 2: a multiple of 128 for CB = 16 / 32 / 64
    A.a_bytes = 128u * A.rb;
    uint32_t cols = 32;
    while (cols < 2u * A.n_pad) cols <<= 1;
    if (cols > 512) return false;
    A.tmem_cols = cols;
    const int Cpad = A.chunks * CB;
    int ob0 = 64;
    while (ob0 > 16 && ob0 / 2 >= A.n_pad) ob0 >>= 1;
    const size_t budget = (size_t)227 * 1024;
    for (int ob = ob0; ob >= 16; ob >>= 1) {
        const size_t fixed = 1024 + ((A.w_bytes + 1023u) & ~1023u) + 2 * kEpiGroups * (size_t)128 * ob * 2 + (size_t)(10 * Cpad + A.n_pad + 64) * 4 + 40 + 16 * kMaxSP +
                             16 * kMaxSA + 16;
        if (fixed + 2 * (size_t)A.a_bytes + 2 * (size_t)A.patch_bytes > budget) continue;
        size_t used = fixed + 2 * (size_t)A.a_bytes + 2 * (size_t)A.patch_bytes;
        int SA = 2, SP = 2;
        auto grow = [&](int& v, int cap, size_t sz) { while (v < cap && used + sz <= budget) { ++v; used += sz; } };
        grow(SP, 4, A.patch_bytes); grow(SA, 3, A.a_bytes); grow(SP, kMaxSP, A.patch_bytes); grow(SA, kMaxSA, A.a_bytes);
        if (ob < ob0 && SP < 3) continue;  // narrower store boxes only when they buy a usable ring
        A.ob = ob; A.SA = SA; A.SP = SP;
        smem = used;
        return true;
    }
    return false;
}

}  // namespace ds
}  // namespace el

using namespace el;

extern "C" int el_dsconv3_ok(int C, int N) {
    ds::Args A{};
    size_t smem = 0;
    return ds::plan(C, N, A, smem) ? 1 : 0;
}

extern "C" int el_dsconv3_fwd(const void* x, const int64_t xs_[4], int C, const float* dw_w, const float* dw_bias, int dw_act, const void* wpk,
                              const float* bias, int act, void* out, const int64_t os_[4], int B, int H, int W, int N, int dtype, void* stream) {
    if (!x || !xs_ || !dw_w || !wpk || !out || !os_ || B <= 0 || C <= 0 || H <= 0 || W <= 0 || N <= 0 || act < 0 || act > 2 || dw_act < 0 || dw_act > 2)
        return EL_ERR_ARG;
    if (dtype != EL_BF16 && dtype != EL_F16) return EL_ERR_UNSUPPORTED;
    if (xs_[1] != 1 || os_[1] != 1 || !aligned16(x) || !aligned16(out) || !aligned16(wpk)) return EL_ERR_UNSUPPORTED;
    for (int i = 0; i < 4; ++i)
        if (i != 1 && (xs_[i] % 8 || os_[i] % 8)) return EL_ERR_UNSUPPORTED;
    ds::Args A{};
    size_t smem = 0;
    if (!ds::plan(C, N, A, smem)) return EL_ERR_UNSUPPORTED;
    A.wpk = wpk; A.bias = bias; A.dw_w = dw_w; A.dw_bias = dw_bias; A.dw_act = dw_act;
    A.C = C; A.N = N; A.H = H; A.W = W; A.B = B;
    A.tiles_x = (int)ceil_div(W, ds::kTW); A.tiles_y = (int)ceil_div(H, ds::kTH);
    A.n_tiles = (int64_t)B * A.tiles_x * A.tiles_y;
    if (!ds::make_map4(&A.src_map, x, C, W, H, B, xs_, A.rb / 2, ds::kPW, ds::kPH, 0, dtype)) return EL_ERR_CUDA;
    if (!ds::make_map4(&A.out_map, out, N, W, H, B, os_, A.ob, ds::kTW, ds::kTH, A.ob * 2, dtype)) return EL_ERR_CUDA;
    const int64_t gx = kSMs < A.n_tiles ? kSMs : A.n_tiles;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
#define EL_DS_LAUNCH(TT, ACT)                                                                                                                \
    {                                                                                                                                        \
        e = cudaFuncSetAttribute(ds::dsconv3_tc_kernel<TT, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);                   \
        if (e == cudaSuccess) e = launch_pdl(ds::dsconv3_tc_kernel<TT, ACT>, dim3((unsigned)gx), dim3(ds::kThreads), smem, st, A);            \
    }
    if (dtype == EL_BF16) {
        if (act == 0) EL_DS_LAUNCH(__nv_bfloat16, 0) else if (act == 1) EL_DS_LAUNCH(__nv_bfloat16, 1) else EL_DS_LAUNCH(__nv_bfloat16, 2)
    } else {
        if (act == 0) EL_DS_LAUNCH(__half, 0) else if (act == 1) EL_DS_LAUNCH(__half, 1) else EL_DS_LAUNCH(__half, 2)
    }
#undef EL_DS_LAUNCH
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return EL_ERR_CUDA; }
    note_launches(1);
    return check_launch();
}
