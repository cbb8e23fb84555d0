// Pointwise (1 x 1) convolution of the inference engine as a streaming tcgen05 GEMM with a fused epilogue:
//     out[p, n] = act( sum_k X[p, k] * W[n, k] + bias[n] ) (+ res[p, n])          p = pixel (B*H*W rows), NHWC activations
// Replaces, for 16-bit activations, cuDNN / cuBLAS 1x1 convs followed by el_bias_act_fwd:
//   Conv(k=1).forward_fuse      nn/modules/conv.py:58-60   (cv1 / cv2 / fuse / f_ll of DSC3K2_Wavelet block.py:3757-3788, C2PSA cv1/cv2/ffn
//                                                            block.py:3412-3497, SPPF block.py:204-223)
//   DSConv.pw + BatchNorm + SiLU nn/modules/conv.py:100-104 (+ the DSBottleneck shortcut, block.py:1500-1503)
//   LinearAttention.qkv / proj  nn/modules/block.py:3353-3373
// The K dimension may be the concatenation of up to 4 source tensors, so the C2f-style `torch.cat` of the reference
// (block.py:3783-3788) never materialises: cv2 reads a, b', m1.. in place.  The result can be split over two
// destinations (chunk(2, 1) of cv1's output).
//
// Roofline: HBM.  AI = 2*K*N / (2*(K+N)) flop/B <= 128 for every EdgeLine site, far below the bf16 ridge (~213 flop/B),
// so the design goal is bytes in flight, not tensor-pipe occupancy:
//   * A (activations): 128-pixel x 64-channel chunks stream through a ring of shared-memory stages with 16-byte
//     cp.async (LDGSTS, L2-only), written directly in the canonical no-swizzle K-major UMMA layout
//     [k-group of 8 channels][pixel row][16 B]; the k-group pitch is padded by 16 B so that the scattered 16 B
//     writes of a warp spread over all banks.  S-2 stages (>= 33 KB per CTA) are in flight while one is consumed.
//   * B (weights): pre-packed on the host into the same layout, loaded ONCE per CTA with 1-D bulk TMA copies and
//     kept resident; CTAs are persistent over pixel tiles.
//   * D: fp32 accumulators in TMEM, double buffered (2 x N columns); tcgen05.mma M128 x N x K16 issued by one
//     thread, completion via tcgen05.commit -> mbarrier; epilogue reads with tcgen05.ld (thread <-> pixel row),
//     adds the folded-BatchNorm bias, applies SiLU / ReLU, adds the shortcut, and writes 16-byte vectors.
#include <type_traits>

#include "el_common.cuh"

namespace el {
namespace pw {

constexpr int kThreads = 128;
constexpr int kTileM = 128;
constexpr uint32_t kLboA = 128 * 16 + 16;    // padded k-group pitch of the A stages (bytes)
constexpr int kMaxChunks = 28;
constexpr int kMaxSrc = 4;

struct Chunk {
    uint16_t g0;    // first channel group (of 8) inside the source
    uint16_t wg0;   // first k-group of this chunk in the packed weight tile
    uint8_t src;    // source tensor
    uint8_t ng;     // real channel groups in this chunk (1..8)
    uint8_t ngp;    // padded to even (k-groups fed to the MMAs)
    uint8_t pad_;
};

struct Args {
    const void* src[kMaxSrc];
    int64_t pitch[kMaxSrc];  // elements between consecutive pixels
    Chunk chunk[kMaxChunks];
    int nchunks;
    const void* wpk;         // [n_tiles][KGp][n_tile][8] 16-bit
    const float* bias;       // [N] or null
    const void* res; int64_t res_pitch;
    void* out; int64_t out_pitch;
    void* out2; int64_t out2_pitch;
    int split;               // channels >= split go to out2 (when out2 != null)
    int64_t M;
    int N, n_tile, kgp;      // kgp = padded k-groups in total
    int act, stages;
    uint32_t stage_bytes;    // widest chunk's k-groups x kLboA
    uint32_t tmem_cols;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
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
// 16-byte LDGSTS, L2 only; src_bytes = 0 zero-fills (rows past M, padded k-groups)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t umma_idesc(int fmt, int M, int N) {  // D = f32, A/B = fmt (0 f16, 1 bf16), both K-major
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

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {  // 32 lanes x 32 bit x 16 columns
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
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

template <typename T>
__global__ void __launch_bounds__(kThreads) pwconv_tc_kernel(const __grid_constant__ Args A) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = A.stages, n_tile = A.n_tile;
    const uint32_t w_bytes = (uint32_t)A.kgp * n_tile * 16;
    // shared memory map: [weights][A ring][bias][barriers]
    const uint32_t sbase = smem_addr(sm);
    const uint32_t off_ring = (w_bytes + 127) & ~127u;
    const uint32_t off_bias = off_ring + (uint32_t)S * A.stage_bytes;
    const uint32_t off_bar = off_bias + (((uint32_t)n_tile * 4 + 127) & ~127u);
    float* s_bias = reinterpret_cast<float*>(sm + off_bias);
    const uint32_t bar_w = sbase + off_bar;             // weights landed
    const uint32_t bar_acc = sbase + off_bar + 8;       // [2] accumulator buffer complete
    const uint32_t bar_free = sbase + off_bar + 24;     // [S] MMAs that read the stage are done
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + off_bar + 24 + 8 * 8);

    const int nt = blockIdx.y, n0 = nt * n_tile;
    const int64_t m_tiles = (A.M + kTileM - 1) / kTileM;
    constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;

    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_acc, 1);
        mbar_init(bar_acc + 8, 1);
        for (int s = 0; s < S; ++s) mbar_init(bar_free + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // resident weight tile of this CTA's output-channel block: bulk TMA copies, <= 32 KB each
        mbar_expect_tx(bar_w, w_bytes);
        const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(A.wpk) + (size_t)nt * w_bytes;
        for (uint32_t o = 0; o < w_bytes; o += 32768) bulk_g2s(sbase + o, wsrc + o, min(32768u, w_bytes - o), bar_w);
    }
    for (int i = tid; i < n_tile; i += kThreads) s_bias[i] = (A.bias && n0 + i < A.N) ? __ldg(A.bias + n0 + i) : 0.f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(s_tmem)), "r"(A.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t idesc = umma_idesc(kFmt, kTileM, n_tile);
    const uint32_t lbo_b = (uint32_t)n_tile * 16;

    // this CTA's tiles: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int64_t first = blockIdx.x;
    const int my_tiles = first < m_tiles ? (int)((m_tiles - first + gridDim.x - 1) / gridDim.x) : 0;
    const int nch = A.nchunks;
    const int total = my_tiles * nch;

    // producer: item j = (tile j / nch, chunk j % nch) -> stage j % S
    auto produce = [&](int j) {
        const int slot = j % S, use = j / S;
        if (use > 0) mbar_wait(bar_free + 8 * slot, (uint32_t)(use - 1) & 1);
        const int tl = j / nch, c = j - tl * nch;
        const Chunk ck = A.chunk[c];
        const int64_t m0 = (first + (int64_t)tl * gridDim.x) * kTileM;
        const T* base = reinterpret_cast<const T*>(A.src[ck.src]);
        const int64_t pitch = A.pitch[ck.src];
        const uint32_t stage = sbase + off_ring + (uint32_t)slot * A.stage_bytes;
        const int ngp = ck.ngp;
        int row = tid / ngp, g = tid - row * ngp;
        const int pieces = kTileM * ngp;
        const bool regular = (kThreads % ngp) == 0;  // 2, 4, 8 k-groups: (row, g) advance by a fixed row step
        const int rstep = kThreads / ngp;
        for (int p = tid; p < pieces; p += kThreads) {
            if (!regular) { row = p / ngp; g = p - row * ngp; }
            const int64_t m = m0 + row;
            const bool valid = m < A.M && g < ck.ng;
            const T* src = valid ? base + m * pitch + (int64_t)(ck.g0 + g) * 8 : base;
            cp_async16(stage + (uint32_t)g * kLboA + (uint32_t)row * 16, src, valid ? 16u : 0u);
            row += rstep;
        }
    };

    const int dist = S - 2;  // stages in flight ahead of the consumer
    for (int j = 0; j < dist; ++j) {
        if (j < total) produce(j);
        cp_async_commit();
    }
    if (tid == 0) mbar_wait(bar_w, 0);

    for (int j = 0; j < total; ++j) {
        if (j + dist < total) produce(j + dist);
        cp_async_commit();
        // groups are committed once per loop trip (empty ones included), so "all but the newest `dist`" = item j has landed
        switch (dist) {
            case 1: cp_async_wait<1>(); break;
            case 2: cp_async_wait<2>(); break;
            case 3: cp_async_wait<3>(); break;
            case 4: cp_async_wait<4>(); break;
            case 5: cp_async_wait<5>(); break;
            default: cp_async_wait<6>(); break;
        }
        proxy_fence();  // LDGSTS writes (generic proxy) -> visible to the tensor core (async proxy)
        __syncthreads();
        const int tl = j / nch, c = j - tl * nch;
        const int buf = tl & 1;
        const bool last = c == nch - 1;
        if (tid == 0) {
            tc_fence_after();
            const Chunk ck = A.chunk[c];
            const uint32_t stage = sbase + off_ring + (uint32_t)(j % S) * A.stage_bytes;
            const uint32_t d = tmem + (uint32_t)buf * n_tile;
            for (int ks = 0; ks < ck.ngp / 2; ++ks) {
                const uint64_t da = umma_desc(stage + (uint32_t)(2 * ks) * kLboA, kLboA, 128);
                const uint64_t db = umma_desc(sbase + (uint32_t)(ck.wg0 + 2 * ks) * lbo_b, lbo_b, 128);
                umma(d, da, db, idesc, (c > 0 || ks > 0) ? 1u : 0u);
            }
            umma_commit(bar_free + 8 * (j % S));
            if (last) umma_commit(bar_acc + 8 * buf);
        }
        if (last) {
            // ---------------------------------------------------------------- epilogue of tile tl (thread <-> pixel row)
            mbar_wait(bar_acc + 8 * buf, (uint32_t)(tl >> 1) & 1);
            tc_fence_after();
            const int64_t m = (first + (int64_t)tl * gridDim.x) * kTileM + tid;
            const bool valid = m < A.M;
            const uint32_t taddr = tmem + (uint32_t)buf * n_tile + ((uint32_t)(warp * 32) << 16);
            const T* rrow = A.res ? reinterpret_cast<const T*>(A.res) + m * A.res_pitch + n0 : nullptr;
            T* orow = reinterpret_cast<T*>(A.out) + m * A.out_pitch;
            T* orow2 = A.out2 ? reinterpret_cast<T*>(A.out2) + m * A.out2_pitch : nullptr;
            for (int c0 = 0; c0 < n_tile; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c0, v);  // warp-collective: executed by all lanes, stores predicated below
                if (!valid || n0 + c0 >= A.N) continue;
                float f[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float t = __uint_as_float(v[e]) + s_bias[c0 + e];
                    f[e] = A.act == 1 ? __fdividef(t, 1.f + __expf(-t)) : (A.act == 2 ? fmaxf(t, 0.f) : t);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int n = n0 + c0 + 8 * h;
                    if (n >= A.N) break;
                    if (rrow) {
                        float r8[8];
                        unpack<T>(ldg_stream(rrow + c0 + 8 * h), r8);
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[8 * h + e] += r8[e];
                    }
                    uint4 o;
                    o.x = pack2<T>(f[8 * h], f[8 * h + 1]); o.y = pack2<T>(f[8 * h + 2], f[8 * h + 3]);
                    o.z = pack2<T>(f[8 * h + 4], f[8 * h + 5]); o.w = pack2<T>(f[8 * h + 6], f[8 * h + 7]);
                    T* dst = (orow2 && n >= A.split) ? orow2 + (n - A.split) : orow + n;
                    *reinterpret_cast<uint4*>(dst) = o;
                }
            }
            tc_fence_before();  // TMEM reads of this buffer are ordered before the MMAs that reuse it (two tiles later)
        }
    }
    cp_async_wait<0>();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(A.tmem_cols));
}

}  // namespace pw
}  // namespace el

using namespace el;

extern "C" int el_pwconv_tile(int N, int k_groups) {
    // output channels per CTA: multiple of 16, <= 256, resident weight tile (k_groups * 16 B per row) <= 144 KB, as even a split as possible
    if (N <= 0 || k_groups <= 0) return 0;
    const int n16 = (int)ceil_div(N, 16) * 16;
    int max_rows = (int)((144 * 1024) / ((int64_t)k_groups * 16)) & ~15;
    if (max_rows > 256) max_rows = 256;
    if (max_rows < 16) return 0;
    const int tiles = (int)ceil_div(n16, max_rows);
    return (int)ceil_div(ceil_div(n16, tiles), 16) * 16;
}

extern "C" int el_pwconv_fwd(int nsrc, const void* const src[], const int64_t src_pitch[], const int32_t src_c[], const void* wpk, const float* bias,
                             const void* res, int64_t res_pitch, void* out, int64_t out_pitch, void* out2, int64_t out2_pitch, int split, int64_t M,
                             int N, int act, int dtype, void* stream) {
    if (nsrc < 1 || nsrc > pw::kMaxSrc || !src || !src_pitch || !src_c || !wpk || !out || M <= 0 || N <= 0 || act < 0 || act > 2) return EL_ERR_ARG;
    if (dtype != EL_BF16 && dtype != EL_F16) return EL_ERR_UNSUPPORTED;
    if (N % 8 || (out2 && (split % 8 || split <= 0 || split >= N))) return EL_ERR_UNSUPPORTED;
    pw::Args A{};
    int nch = 0, wg = 0;
    for (int i = 0; i < nsrc; ++i) {
        if (!src[i] || src_c[i] <= 0 || src_c[i] % 8 || src_pitch[i] % 8 || !aligned16(src[i])) return EL_ERR_UNSUPPORTED;
        A.src[i] = src[i];
        A.pitch[i] = src_pitch[i];
        const int kg = src_c[i] / 8;
        for (int g0 = 0; g0 < kg; g0 += 8) {
            if (nch >= pw::kMaxChunks) return EL_ERR_UNSUPPORTED;
            const int ng = kg - g0 < 8 ? kg - g0 : 8;
            pw::Chunk& c = A.chunk[nch++];
            c.g0 = (uint16_t)g0; c.wg0 = (uint16_t)wg; c.src = (uint8_t)i; c.ng = (uint8_t)ng; c.ngp = (uint8_t)((ng + 1) & ~1);
            wg += c.ngp;
        }
    }
    if (!aligned16(out) || out_pitch % 8 || (out2 && (!aligned16(out2) || out2_pitch % 8)) || (res && (!aligned16(res) || res_pitch % 8)))
        return EL_ERR_UNSUPPORTED;
    A.nchunks = nch;
    A.kgp = wg;
    A.wpk = wpk; A.bias = bias; A.res = res; A.res_pitch = res_pitch;
    A.out = out; A.out_pitch = out_pitch; A.out2 = out2; A.out2_pitch = out2_pitch; A.split = out2 ? split : N;
    A.M = M; A.N = N; A.act = act;
    A.n_tile = el_pwconv_tile(N, wg);
    if (A.n_tile <= 0) return EL_ERR_UNSUPPORTED;
    const int n_tiles = (int)ceil_div(N, A.n_tile);
    uint32_t cols = 32;
    while (cols < 2u * A.n_tile) cols <<= 1;
    if (cols > 512) return EL_ERR_UNSUPPORTED;
    A.tmem_cols = cols;
    const size_t w_bytes = (size_t)A.kgp * A.n_tile * 16;
    const size_t fixed = ((w_bytes + 127) & ~(size_t)127) + (((size_t)A.n_tile * 4 + 127) & ~(size_t)127) + 24 + 8 * 8 + 16;
    int max_ngp = 2;
    for (int i = 0; i < nch; ++i) max_ngp = A.chunk[i].ngp > max_ngp ? A.chunk[i].ngp : max_ngp;
    A.stage_bytes = (uint32_t)max_ngp * pw::kLboA;
    // shared memory per CTA: the smallest of 56 / 100 / 220 KB (4 / 2 / 1 CTAs per SM) that still holds a ring of >= 5 stages
    // (>= 3 in flight); with resident weights too large for that, whatever ring fits in 220 KB (>= 3 stages)
    int S = 0;
    for (size_t budget : {(size_t)56 * 1024, (size_t)100 * 1024, (size_t)220 * 1024}) {
        if (budget <= fixed) continue;
        S = (int)((budget - fixed) / A.stage_bytes);
        if (S >= 5) break;
    }
    if (S < 3) return EL_ERR_UNSUPPORTED;
    if (S > 8) S = 8;
    A.stages = S;
    const size_t smem = fixed + (size_t)S * A.stage_bytes;
    int per_sm = (int)(227 * 1024 / (smem + 1024));
    if (per_sm > (int)(512 / cols)) per_sm = (int)(512 / cols);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    const int64_t m_tiles = ceil_div(M, pw::kTileM);
    int64_t gx = (int64_t)kSMs * per_sm / n_tiles;
    if (gx < 1) gx = 1;
    if (gx > m_tiles) gx = m_tiles;
    dim3 grid((unsigned)gx, (unsigned)n_tiles);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (dtype == EL_BF16) {
        e = cudaFuncSetAttribute(pw::pwconv_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { g_last_cuda_error = (int)e; return EL_ERR_CUDA; }
        pw::pwconv_tc_kernel<__nv_bfloat16><<<grid, pw::kThreads, smem, st>>>(A);
    } else {
        e = cudaFuncSetAttribute(pw::pwconv_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { g_last_cuda_error = (int)e; return EL_ERR_CUDA; }
        pw::pwconv_tc_kernel<__half><<<grid, pw::kThreads, smem, st>>>(A);
    }
    note_launches(1);
    return check_launch();
}
