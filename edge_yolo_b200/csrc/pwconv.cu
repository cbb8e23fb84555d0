// Pointwise (1 x 1) convolution of the inference engine as a streaming TMA + tcgen05 GEMM with a fused epilogue:
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
// so the design goal is bytes in flight and few LSU wavefronts, not tensor-pipe occupancy.  Warp-specialised, persistent:
//   warp 0 (one lane)  TMA producer: 128-pixel x <=64-channel boxes of the activations (2-D tensor maps, 32/64/128-byte
//                      swizzle = row bytes, out-of-range pixels / channels zero-filled by the hardware) into a ring of
//                      shared-memory stages; full/empty mbarriers.
//   warp 1 (one lane)  tcgen05.mma M128 x N x K16, A and B from shared memory (K-major swizzled tiles), fp32 accumulators
//                      in TMEM, double buffered (2 x N columns); tcgen05.commit frees the stage / publishes the tile.
//   warps 2-5          epilogue: tcgen05.ld (thread <-> pixel row), + folded-BatchNorm bias, SiLU / ReLU, + shortcut,
//                      16-bit pack into a swizzled staging tile, TMA store (clips the ragged last pixel tile and channel
//                      tails); overlaps the loads and MMAs of the next tile.
// Weights are pre-packed on the host into the same swizzled tiles and stay resident in shared memory (1-D bulk copies,
// once per CTA).
#include <cuda.h>
#include <stdlib.h>

#include <type_traits>

#include "el_common.cuh"

namespace el {
namespace pw {

constexpr int kThreads = 320;   // TMA warp, MMA warp, one or two epilogue groups of four warps (the launch uses 64 + 128 * groups)
constexpr int kTileM = 128;
constexpr int kMaxChunks = 40;
constexpr int kMaxSrc = 4;
constexpr int kMaxStages = 8;

struct Chunk {
    uint32_t w_off;   // byte offset of this chunk's weight tile in the resident weight block
    uint16_t c0;      // first channel inside the source
    uint8_t src;      // source tensor (1x1) or filter tap ky * 3 + kx (3x3)
    uint8_t rb;       // row bytes of the box = swizzle span: 32, 64 or 128 (16, 32 or 64 channels)
};

struct Args {
    CUtensorMap src_map[kMaxSrc];
    CUtensorMap out_map, out2_map;
    Chunk chunk[kMaxChunks];
    int nchunks;
    const void* wpk;         // [n_tiles][w_bytes] swizzled weight tiles
    const float* bias;       // [N] or null
    const void* res; int64_t res_pitch; float res_scale;  // RES 1: out = res + res_scale * act(...);  RES 2: addend z (B, N, H/2, W/2), pitch = its pixel pitch
    int up_H, up_W;          // RES 2: output map height / width (the flat pixel index is decoded to find the low-resolution addend row)
    int has_out2, split;     // channels >= split go to out2
    int64_t M;
    int N, n_tile, ob;       // ob: channels per staging / store box (16, 32 or 64)
    int act, stages;
    int groups;              // epilogue groups: 1 (several CTAs per SM hide each other's epilogue) or 2 (one CTA per SM: group g drains accumulator g, tiles alternate)
    uint32_t w_bytes, stage_bytes, tmem_cols;
    int stream_w;            // 1: weight tiles are not resident: each ring stage holds [A box | weight tile of the chunk] (wide K)
    uint32_t a_stage_bytes;  // stream_w: offset of the weight tile inside a stage
    // 3x3 mode (el_conv3x3_fwd): a pixel tile is a tw x th patch of one image (tw * th = 128), src_map[0] / out_map are 4-D
    // (channel, x, y, image) maps, chunk.src is the filter tap and the box of tap (ky, kx) starts at (x0 * stride + kx - 1, y0 * stride + ky - 1)
    int spatial, tw, th, tiles_x, tiles_y, stride;
    int64_t n_tiles_m;
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
// 2-D tiled TMA load: box at (c0 = channel, c1 = pixel) -> shared, completion in bytes on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar),
                 "r"(c0), "r"(c1)
                 : "memory");
}
// 2-D tiled TMA store shared -> global (SASS: UTMASTG); out-of-range rows / channels are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
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

// K-major swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30) (unused: one swizzle
// atom along K), SBO>>4 [32,46) = 8 rows x row bytes, version=1 [46,48), layout [61,64): 2 = 128B, 4 = 64B, 6 = 32B swizzle
__device__ __forceinline__ uint64_t umma_desc_sw(uint32_t saddr, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((8 * row_bytes) >> 4) << 32) | (1ull << 46) | (layout << 61);
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
// one lane of a converged warp.  The producer / MMA loops stay warp-uniform around it, so coordinates and descriptors live in uniform registers
// and UTMALDG / UTCHMMA are not wrapped in a per-thread register -> uniform-register waterfall (what `if (lane == 0) { loop }` compiles to:
// ~19 instructions per MMA, measured on the 3x3 halo kernel where 36 MMAs per chunk made the issue thread the bottleneck)
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

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {  // 32 lanes x 32 bit x 16 columns
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
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
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// RES: 0 none, 1 out = res + res_scale * act(.), 2 out = act(. + nearest-2x-upsampled addend);  GROUPS: epilogue groups (threads = 64 + 128 * GROUPS).
// The one-group form must keep four CTAs of 192 threads resident per SM (<= 80 registers): that co-residency is what hides its epilogue.
template <typename T, int ACT, int RES, int GROUPS>
__global__ void __launch_bounds__(64 + 128 * GROUPS, GROUPS == 1 ? 4 : 1) pwconv_tc_kernel(const __grid_constant__ Args A) {
    extern __shared__ __align__(1024) unsigned char sm_raw[];
    // dynamic shared memory is only guaranteed 16-byte aligned: round up to the 1024 B the 128-byte swizzle atoms need
    const uint32_t sbase = (smem_addr(sm_raw) + 1023u) & ~1023u;
    unsigned char* sm = sm_raw + (sbase - smem_addr(sm_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = A.stages, n_tile = A.n_tile;
    // shared memory map: [weights][stage ring][2 staging tiles 128 x ob][bias][barriers]
    const uint32_t off_ring = A.stream_w ? 0u : ((A.w_bytes + 1023u) & ~1023u);
    const uint32_t off_stage = off_ring + (uint32_t)S * A.stage_bytes;
    const uint32_t staging_bytes = (uint32_t)kTileM * A.ob * 2;
    const uint32_t off_bias = off_stage + 2 * (uint32_t)GROUPS * staging_bytes;
    const uint32_t off_bar = off_bias + (((uint32_t)(n_tile + 64) * 4 + 127) & ~127u);  // + 64: the last store box may overhang n_tile
    float* s_bias = reinterpret_cast<float*>(sm + off_bias);
    const uint32_t bar_w = sbase + off_bar;                  // weights landed
    const uint32_t bar_acc_full = sbase + off_bar + 8;       // [2] accumulator buffer complete
    const uint32_t bar_acc_empty = sbase + off_bar + 24;     // [2] accumulator buffer drained by the epilogue
    const uint32_t bar_full = sbase + off_bar + 40;          // [S] stage filled by TMA
    const uint32_t bar_empty = bar_full + 8 * kMaxStages;    // [S] stage consumed by the MMAs
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + off_bar + 40 + 16 * kMaxStages);

    const int nt = blockIdx.y, n0 = nt * n_tile;
    const int64_t m_tiles = A.spatial ? A.n_tiles_m : (A.M + kTileM - 1) / kTileM;
    const int64_t first = blockIdx.x;
    const int my_tiles = first < m_tiles ? (int)((m_tiles - first + gridDim.x - 1) / gridDim.x) : 0;
    const int nch = A.nchunks;
    constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;

    pdl_launch_dependents();  // the next kernel may start its prologue; it waits for this grid before touching activations
    if (tid == 0) {
        for (int i = 0; i < kMaxSrc; ++i)
            if (i == 0 || (!A.spatial && A.chunk[nch - 1].src >= i)) asm volatile("prefetch.tensormap [%0];" ::"l"(&A.src_map[i]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&A.out_map) : "memory");
        mbar_init(bar_w, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(bar_acc_full + 8 * b, 1); mbar_init(bar_acc_empty + 8 * b, 128); }
        for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (!A.stream_w) {
            mbar_expect_tx(bar_w, A.w_bytes);
            const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(A.wpk) + (size_t)nt * A.w_bytes;
            for (uint32_t o = 0; o < A.w_bytes; o += 32768) bulk_g2s(sbase + o, wsrc + o, min(32768u, A.w_bytes - o), bar_w);
        }
    }
    for (int i = tid; i < n_tile + 64; i += (int)blockDim.x) s_bias[i] = (A.bias && n0 + i < A.N) ? __ldg(A.bias + n0 + i) : 0.f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(s_tmem)), "r"(A.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    pdl_wait();  // everything above touched only weights / bias; activations of the previous kernel are complete from here on

    if (warp == 0) {
        // ------------------------------------------------------------------------------------ TMA producer
        {
            const bool leader = elect_one();
            int it = 0;
            for (int tl = 0; tl < my_tiles; ++tl) {
                const int64_t tile = first + (int64_t)tl * gridDim.x;
                const int m0 = (int)(tile * kTileM);
                int img = 0, x0 = 0, y0 = 0;
                if (A.spatial) {
                    const int per_img = A.tiles_x * A.tiles_y, r = (int)(tile % per_img);
                    img = (int)(tile / per_img);
                    y0 = (r / A.tiles_x) * A.th * A.stride - 1;
                    x0 = (r % A.tiles_x) * A.tw * A.stride - 1;
                }
                for (int c = 0; c < nch; ++c, ++it) {
                    const int s = it % S, use = it / S;
                    if (use > 0) mbar_wait(bar_empty + 8 * s, (uint32_t)(use - 1) & 1);
                    const Chunk ck = A.chunk[c];
                    const uint32_t dst = sbase + off_ring + (uint32_t)s * A.stage_bytes;
                    const uint32_t wt_bytes = A.stream_w ? (uint32_t)n_tile * ck.rb : 0u;
                    if (leader) {
                        mbar_expect_tx(bar_full + 8 * s, (uint32_t)kTileM * ck.rb + wt_bytes);
                        if (A.stream_w)  // this chunk's weight tile rides in the same stage (L2-resident after the first tile)
                            bulk_g2s(dst + A.a_stage_bytes, reinterpret_cast<const unsigned char*>(A.wpk) + (size_t)nt * A.w_bytes + ck.w_off, wt_bytes,
                                     bar_full + 8 * s);
                        if (A.spatial) tma_load_4d(dst, &A.src_map[0], ck.c0, x0 + ck.src % 3, y0 + ck.src / 3, img, bar_full + 8 * s);
                        else tma_load_2d(dst, &A.src_map[ck.src], ck.c0, m0, bar_full + 8 * s);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------------------------ MMA issuer
        {
            const bool leader = elect_one();
            const uint32_t idesc = umma_idesc(kFmt, kTileM, n_tile);
            if (!A.stream_w) mbar_wait(bar_w, 0);
            int it = 0;
            for (int tl = 0; tl < my_tiles; ++tl) {
                const int b = tl & 1, ub = tl >> 1;
                if (ub > 0) mbar_wait(bar_acc_empty + 8 * b, (uint32_t)(ub - 1) & 1);
                tc_fence_after();
                const uint32_t d = tmem + (uint32_t)b * n_tile;
                for (int c = 0; c < nch; ++c, ++it) {
                    const int s = it % S;
                    mbar_wait(bar_full + 8 * s, (uint32_t)(it / S) & 1);
                    tc_fence_after();
                    const Chunk ck = A.chunk[c];
                    const uint32_t a_base = sbase + off_ring + (uint32_t)s * A.stage_bytes;
                    const uint32_t b_base = A.stream_w ? a_base + A.a_stage_bytes : sbase + ck.w_off;
                    const uint64_t da = umma_desc_sw(a_base, ck.rb), db = umma_desc_sw(b_base, ck.rb);
                    const int ksteps = ck.rb / 32;
                    if (leader) {
                        for (int ks = 0; ks < ksteps; ++ks)  // one MMA per 16 channels = 32 bytes along K inside the swizzle atom (+2 in the descriptor's start field)
                            umma(d, da + (uint64_t)(2 * ks), db + (uint64_t)(2 * ks), idesc, (c > 0 || ks > 0) ? 1u : 0u);
                        umma_commit(bar_empty + 8 * s);
                        if (c == nch - 1) umma_commit(bar_acc_full + 8 * b);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ------------------------------------------------------------------------------------ epilogue warps
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;          // pixel row inside the tile
        const int g = (warp - 2) >> 2;          // epilogue group: owns accumulator g and the tiles tl = g (mod groups) when there are two
        constexpr int G = GROUPS;
        const int et = (tid - 64) & 127;        // 0..127 inside the group
        const int ob = A.ob, rbo = ob * 2;      // staging row bytes = swizzle span of the store box
        const uint32_t swz = ((uint32_t)(row * rbo) >> 7) & (uint32_t)(rbo / 16 - 1);
        const int n_real = min(n_tile, A.N - n0);
        int sub = 0;  // staging buffer use counter
        for (int tl = g; tl < my_tiles; tl += G) {
            const int b = tl & 1;   // one group: accumulators alternate; two groups: b == g
            const int64_t m0 = (first + (int64_t)tl * gridDim.x) * kTileM;
            mbar_wait(bar_acc_full + 8 * b, (uint32_t)(tl >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + (uint32_t)b * n_tile + ((uint32_t)(q * 32) << 16);
            const bool rvalid = m0 + row < A.M;
            const T* rrow = nullptr;
            if (RES == 1) rrow = reinterpret_cast<const T*>(A.res) + (m0 + row) * A.res_pitch + n0;
            if (RES == 2 && rvalid) {  // pixel (b, y, x) of the output reads pixel (b, y / 2, x / 2) of the low-resolution addend
                const int64_t m = m0 + row;
                const int hw = A.up_H * A.up_W;
                const int64_t bimg = m / hw;
                const int r = (int)(m - bimg * hw), y = r / A.up_W, x = r - y * A.up_W;
                rrow = reinterpret_cast<const T*>(A.res) + ((bimg * (A.up_H / 2) + y / 2) * (A.up_W / 2) + x / 2) * A.res_pitch + n0;
            }
            for (int c0 = 0; c0 < n_real; c0 += ob, ++sub) {
                const uint32_t stg = sbase + off_stage + (uint32_t)(2 * g + (sub & 1)) * staging_bytes;
                if (et == 0) bulk_wait_read<1>();  // the store that last read this staging buffer (two uses ago) is done with it
                epi_barrier(g);
                const int jmax = min(ob, ((n_real - c0) + 15) & ~15);  // the last box of an 80-channel conv holds 16 real columns: skip the other 48
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
                    if (RES == 2) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (rvalid && n0 + c0 + j + 8 * h < A.N) {
                                float r8[8];
                                unpack<T>(ldg_cached(rrow + c0 + j + 8 * h), r8);  // each low-resolution pixel is read by 4 output pixels: keep it in L1
#pragma unroll
                                for (int e = 0; e < 8; ++e) f[8 * h + e] += r8[e];
                            }
                        }
                    }
                    if (ACT == 1) {  // SiLU: x * sigmoid(x) = h + h * tanh(h), h = x / 2 (one MUFU per element; 16-bit outputs)
#pragma unroll
                        for (int e = 0; e < 16; ++e) { const float h = 0.5f * f[e]; f[e] = fmaf(h, tanh_fast(h), h); }
                    } else if (ACT == 2) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) f[e] = fmaxf(f[e], 0.f);
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (RES == 1) {
                            if (rvalid && n0 + c0 + j + 8 * h < A.N) {
                                float r8[8];
                                unpack<T>(ldg_stream(rrow + c0 + j + 8 * h), r8);
#pragma unroll
                                for (int e = 0; e < 8; ++e) f[8 * h + e] = fmaf(A.res_scale, f[8 * h + e], r8[e]);
                            }
                        }
                        uint4 o;
                        o.x = pack2<T>(f[8 * h], f[8 * h + 1]); o.y = pack2<T>(f[8 * h + 2], f[8 * h + 3]);
                        o.z = pack2<T>(f[8 * h + 4], f[8 * h + 5]); o.w = pack2<T>(f[8 * h + 6], f[8 * h + 7]);
                        const uint32_t chunk = (uint32_t)(j / 8 + h) ^ swz;
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)row * rbo + chunk * 16), "r"(o.x), "r"(o.y), "r"(o.z),
                                     "r"(o.w)
                                     : "memory");
                    }
                }
                }
                proxy_fence();  // generic-proxy writes of the staging tile -> visible to the TMA store (async proxy)
                epi_barrier(g);
                if (et == 0) {
                    const int n = n0 + c0;
                    if (A.spatial) {
                        const int64_t tile = first + (int64_t)tl * gridDim.x;
                        const int per_img = A.tiles_x * A.tiles_y, r = (int)(tile % per_img);
                        tma_store_4d(&A.out_map, n, (r % A.tiles_x) * A.tw, (r / A.tiles_x) * A.th, (int)(tile / per_img), stg);
                    } else if (A.has_out2 && n >= A.split) {
                        tma_store_2d(&A.out2_map, n - A.split, (int)m0, stg);
                    } else {
                        tma_store_2d(&A.out_map, n, (int)m0, stg);
                    }
                    bulk_commit();
                }
            }
            tc_fence_before();
            mbar_arrive(bar_acc_empty + 8 * b);
        }
        if (et == 0) bulk_wait_read<0>();
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

// (channels, pixels) view with `pitch` elements between pixels; box = (row_bytes / 2 channels, 128 pixels), swizzle span = row bytes
static bool make_map(CUtensorMap* map, const void* base, int channels, int64_t M, int64_t pitch, int row_bytes, int dtype) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)channels, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
    const cuuint32_t box[2] = {(cuuint32_t)(row_bytes / 2), (cuuint32_t)kTileM};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    return fn(map, dtype == EL_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box,
              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// (channels, x, y, image) view of an NHWC activation; box = (row_bytes / 2 channels, tw pixels, th rows, 1 image) traversed with
// `stride` along x and y (a stride-2 convolution reads every second pixel of the tap-shifted window)
static bool make_map4(CUtensorMap* map, const void* base, int channels, int W, int H, int B, const int64_t st[4] /* n, c, h, w elements */, int row_bytes,
                      int tw, int th, int stride, int dtype) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)channels, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)st[3] * 2, (cuuint64_t)st[2] * 2, (cuuint64_t)st[0] * 2};
    const cuuint32_t box[4] = {(cuuint32_t)(row_bytes / 2), (cuuint32_t)(tw * stride), (cuuint32_t)(th * stride), 1};
    const cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    return fn(map, dtype == EL_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box,
              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static inline int box_bytes_for(int channels) { return channels <= 16 ? 32 : (channels <= 32 ? 64 : 128); }

template <typename T>
static cudaError_t launch(const Args& A, dim3 grid, size_t smem, int res_mode, cudaStream_t st) {
#define EL_PW_LAUNCH_G(ACT, RES, GR)                                                                                                         \
    {                                                                                                                                        \
        cudaError_t e = cudaFuncSetAttribute(pwconv_tc_kernel<T, ACT, RES, GR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);    \
        if (e != cudaSuccess) return e;                                                                                                      \
        return launch_pdl(pwconv_tc_kernel<T, ACT, RES, GR>, grid, dim3(64 + 128 * GR), smem, st, A);                                        \
    }
#define EL_PW_LAUNCH(ACT, RES)                                                                                                               \
    {                                                                                                                                        \
        if (A.groups == 2) EL_PW_LAUNCH_G(ACT, RES, 2) else EL_PW_LAUNCH_G(ACT, RES, 1)                                                      \
    }
#define EL_PW_BY_ACT(RES)                                                                                                               \
    {                                                                                                                                   \
        if (A.act == 0) EL_PW_LAUNCH(0, RES) else if (A.act == 1) EL_PW_LAUNCH(1, RES) else EL_PW_LAUNCH(2, RES)                        \
    }
    if (res_mode == 1) EL_PW_BY_ACT(1) else if (res_mode == 2) EL_PW_BY_ACT(2) else EL_PW_BY_ACT(0)
#undef EL_PW_BY_ACT
#undef EL_PW_LAUNCH
#undef EL_PW_LAUNCH_G
}

}  // namespace pw
}  // namespace el

using namespace el;

extern "C" int el_pwconv_tile(int N, int w_row_bytes, int64_t M) {
    // output channels per CTA: multiple of 16, <= 256, resident weight block (w_row_bytes per output channel) <= 128 KB, as even a split
    // as possible; few pixel tiles (small maps) split N further, down to 64 channels, until ~2 CTAs per SM exist
    if (N <= 0 || w_row_bytes <= 0 || M <= 0) return 0;
    const int n16 = (int)ceil_div(N, 16) * 16;
    int max_rows = (int)((128 * 1024) / (int64_t)w_row_bytes) & ~15;
    if (max_rows > 256) max_rows = 256;
    if (max_rows < 16) return 0;
    int tiles = (int)ceil_div(n16, max_rows);
    const int64_t m_tiles = ceil_div(M, pw::kTileM);
    while (m_tiles * tiles < 2 * kSMs && ceil_div(n16, tiles + 1) >= 64) ++tiles;
    return (int)ceil_div(ceil_div(n16, tiles), 16) * 16;
}

extern "C" int el_conv3x3_tile(int N, int C, int64_t M) {
    // C <= 32: the nine weight tiles stay resident (el_pwconv_tile rule).  Wider inputs stream each chunk's weight tile through the
    // ring next to its activation box, so N is only split by the UMMA limit (256) and, on small maps, to get ~2 CTAs per SM.
    if (N <= 0 || C <= 0 || M <= 0) return 0;
    const int rb = C <= 16 ? 32 : (C <= 32 ? 64 : 128);
    const int chunks = 9 * (int)ceil_div(C, rb / 2);
    if (C <= 32) return el_pwconv_tile(N, chunks * rb, M);
    const int n16 = (int)ceil_div(N, 16) * 16;
    int tiles = (int)ceil_div(n16, 256);
    const int64_t m_tiles = ceil_div(M, pw::kTileM);
    while (m_tiles * tiles < 2 * kSMs && ceil_div(n16, tiles + 1) >= 64) ++tiles;
    return (int)ceil_div(ceil_div(n16, tiles), 16) * 16;
}

extern "C" int el_pwconv_fwd(int nsrc, const void* const src[], const int64_t src_pitch[], const int32_t src_c[], const void* wpk, const float* bias,
                             const void* res, int64_t res_pitch, float res_scale, int up_H, int up_W, void* out, int64_t out_pitch, void* out2,
                             int64_t out2_pitch, int split, int64_t M, int N, int act, int dtype, void* stream) {
    if (nsrc < 1 || nsrc > pw::kMaxSrc || !src || !src_pitch || !src_c || !wpk || !out || M <= 0 || N <= 0 || act < 0 || act > 2) return EL_ERR_ARG;
    if (dtype != EL_BF16 && dtype != EL_F16) return EL_ERR_UNSUPPORTED;
    if (N % 8 || M >= (1ll << 31) - 256 || (out2 && (split % 16 || split <= 0 || split >= N))) return EL_ERR_UNSUPPORTED;
    if (!aligned16(out) || out_pitch % 8 || (out2 && (!aligned16(out2) || out2_pitch % 8)) || (res && (!aligned16(res) || res_pitch % 8)))
        return EL_ERR_UNSUPPORTED;
    for (int i = 0; i < nsrc; ++i)
        if (!src[i] || src_c[i] <= 0 || src_c[i] % 8 || src_pitch[i] % 8 || !aligned16(src[i])) return EL_ERR_UNSUPPORTED;
    pw::Args A{};
    // K chunks: per source, boxes of 64 channels (16 / 32 when the whole source is that narrow); tails are zero-filled by TMA
    int nch = 0, w_row_bytes = 0, max_rb = 32;
    for (int i = 0; i < nsrc; ++i) {
        const int rb = pw::box_bytes_for(src_c[i]);
        for (int c0 = 0; c0 < src_c[i]; c0 += rb / 2) {
            if (nch >= pw::kMaxChunks) return EL_ERR_UNSUPPORTED;
            pw::Chunk& c = A.chunk[nch++];
            c.c0 = (uint16_t)c0; c.src = (uint8_t)i; c.rb = (uint8_t)rb;
            w_row_bytes += rb;
        }
        max_rb = rb > max_rb ? rb : max_rb;
    }
    A.nchunks = nch;
    A.n_tile = el_pwconv_tile(N, w_row_bytes, M);
    if (A.n_tile <= 0) return EL_ERR_UNSUPPORTED;
    uint32_t w_off = 0;
    for (int i = 0; i < nch; ++i) {  // every weight tile starts on a 1024 B boundary (swizzle atom alignment)
        A.chunk[i].w_off = w_off;
        w_off += ((uint32_t)A.n_tile * A.chunk[i].rb + 1023u) & ~1023u;
    }
    A.w_bytes = w_off;
    A.stage_bytes = (uint32_t)pw::kTileM * max_rb;
    A.a_stage_bytes = A.stage_bytes;
    const int n_tiles = (int)ceil_div(N, A.n_tile);
    if (n_tiles > 1) {
        // In-place use (a destination aliasing a K source, e.g. the enhancer's `fuse` GEMM writing over b) is only safe with ONE output-channel
        // tile: the CTA of column tile nt loads ALL K channels of its pixel rows but stores only its own n_tile channels, so with the N axis
        // split CTA (m, 0) could overwrite channels that CTA (m, 1) has not loaded yet (cross-CTA write-after-read race).  Refuse it.
        auto aliases = [&](const void* o, int64_t o_pitch, int o_c) {
            const char* ob0 = (const char*)o;
            const char* ob1 = ob0 + ((M - 1) * o_pitch + o_c) * 2;
            for (int i = 0; i < nsrc; ++i) {
                const char* sb0 = (const char*)src[i];
                const char* sb1 = sb0 + ((M - 1) * src_pitch[i] + src_c[i]) * 2;
                if (ob1 <= sb0 || sb1 <= ob0) continue;            // disjoint byte ranges
                if (o_pitch == src_pitch[i]) {                      // channel slices of one pixel-major buffer: disjoint channel windows are fine
                    int64_t d = ((ob0 - sb0) / 2) % o_pitch;
                    if (d < 0) d += o_pitch;
                    if (d >= src_c[i] && d + o_c <= o_pitch) continue;
                }
                return true;
            }
            return false;
        };
        if (aliases(out, out_pitch, out2 ? split : N) || (out2 && aliases(out2, out2_pitch, N - split))) return EL_ERR_ARG;
    }
    // store box: up to 64 channels; with two destinations it must not straddle the split
    int ob = 64;
    if (out2) while (split % ob) ob >>= 1;
    while (ob > 16 && ob / 2 >= A.n_tile) ob >>= 1;
    if (n_tiles > 1) while (A.n_tile % ob) ob >>= 1;  // an overhanging last box is only harmless past N (clipped), not into the next tile
    if (ob < 16) return EL_ERR_UNSUPPORTED;
    A.ob = ob;
    A.wpk = wpk; A.bias = bias; A.res = res; A.res_pitch = res_pitch; A.res_scale = res_scale; A.up_H = up_H; A.up_W = up_W;
    A.has_out2 = out2 != nullptr; A.split = out2 ? split : N;
    A.M = M; A.N = N; A.act = act;
    uint32_t cols = 32;
    while (cols < 2u * A.n_tile) cols <<= 1;
    if (cols > 512) return EL_ERR_UNSUPPORTED;
    A.tmem_cols = cols;
    // CTAs per SM: the most (<= 4, limited by TMEM columns) whose share of shared memory still holds a ring of >= 3 stages (one being
    // consumed, two in flight per CTA); with resident weights too large for that, whatever ring fits one CTA (>= 2 stages).
    // Two or more co-resident CTAs are what overlaps one tile's epilogue with the other's loads and MMAs -- so when 64-channel
    // store boxes (2 x 16 KB of staging) leave room for a single CTA only, 32-channel boxes are used if they make room for two
    // with a ring that still covers a whole tile (measured: 128 -> 128 @ 80^2 51.7 -> 42.5 us).
    size_t fixed = 0;
    int S = 0, per_sm = 1;
    auto plan = [&](int box) {
        fixed = 1024 /* alignment slack */ + ((A.w_bytes + 1023u) & ~1023u) + 2 * (size_t)pw::kTileM * box * 2 +
                (((size_t)(A.n_tile + 64) * 4 + 127) & ~(size_t)127) + 40 + 16 * pw::kMaxStages + 16;
        S = 0; per_sm = 1;
        for (int ps = 4; ps >= 1; --ps) {
            if (ps > (int)(512 / cols)) continue;
            const size_t budget = (size_t)227 * 1024 / ps - 1024;
            if (budget <= fixed) continue;
            const int s_fit = (int)((budget - fixed) / A.stage_bytes);
            if (s_fit >= 3 || (ps == 1 && s_fit >= 2)) { S = s_fit; per_sm = ps; break; }
        }
    };
    plan(ob);
    if (per_sm == 1 && ob == 64) {
        const size_t fixed0 = fixed; const int S0 = S;
        plan(32);
        if (per_sm >= 2 && S >= nch + 1 && S >= 3) { ob = 32; A.ob = 32; }
        else { fixed = fixed0; S = S0; per_sm = 1; }
    }
    if (S < 2) return EL_ERR_UNSUPPORTED;
    // One CTA per SM (large resident weights): nothing overlaps a tile's epilogue latency chain but the next tile's loads and MMAs, so a second
    // epilogue group (four more warps, its own pair of staging tiles, accumulator g for tiles tl = g mod 2) keeps two epilogues in flight --
    // what made the 3x3 halo kernel 1.25x faster.  EL_PW_GROUPS=1 keeps the single group.
    A.groups = 1;
    static const int max_groups = [] { const char* v = getenv("EL_PW_GROUPS"); return v ? atoi(v) : 2; }();
    if (per_sm == 1 && max_groups >= 2) {
        const size_t fixed2 = fixed + 2 * (size_t)pw::kTileM * A.ob * 2;
        const size_t budget = (size_t)227 * 1024 - 1024;
        if (budget > fixed2 && (budget - fixed2) / A.stage_bytes >= 3 && ceil_div(M, pw::kTileM) * n_tiles >= 2 * kSMs) {
            A.groups = 2; fixed = fixed2; S = (int)((budget - fixed2) / A.stage_bytes);
        }
    }
    if (S > pw::kMaxStages) S = pw::kMaxStages;
    A.stages = S;
    const size_t smem = fixed + (size_t)S * A.stage_bytes;
    const int64_t m_tiles = ceil_div(M, pw::kTileM);
    int64_t gx = (int64_t)kSMs * per_sm / n_tiles;
    if (gx < 1) gx = 1;
    if (gx > m_tiles) gx = m_tiles;
    for (int i = 0; i < nsrc; ++i)
        if (!pw::make_map(&A.src_map[i], src[i], src_c[i], M, src_pitch[i], pw::box_bytes_for(src_c[i]), dtype)) return EL_ERR_CUDA;
    if (!pw::make_map(&A.out_map, out, out2 ? split : N, M, out_pitch, ob * 2, dtype)) return EL_ERR_CUDA;
    if (out2 && !pw::make_map(&A.out2_map, out2, N - split, M, out2_pitch, ob * 2, dtype)) return EL_ERR_CUDA;
    dim3 grid((unsigned)gx, (unsigned)n_tiles);
    cudaStream_t st = (cudaStream_t)stream;
    const int res_mode = !res ? 0 : (up_H > 0 ? 2 : 1);
    if (res_mode == 2 && ((up_H & 1) || (up_W & 1) || up_W <= 0 || M % ((int64_t)up_H * up_W))) return EL_ERR_ARG;
    const cudaError_t e = dtype == EL_BF16 ? pw::launch<__nv_bfloat16>(A, grid, smem, res_mode, st) : pw::launch<__half>(A, grid, smem, res_mode, st);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return EL_ERR_CUDA; }
    note_launches(1);
    return check_launch();
}

// Dense 3 x 3 convolution (padding 1, stride 1 or 2) + bias + activation as an implicit GEMM on the same kernel: the nine filter
// taps are nine K-chunk groups whose TMA boxes are the tap-shifted windows of the input (zero padding = TMA out-of-bounds fill).
extern "C" int el_conv3x3_fwd(const void* x, const int64_t xs_[4], int C, const void* wpk, const float* bias, void* out, const int64_t os_[4], int B,
                              int H, int W, int N, int stride, int act, int dtype, void* stream) {
    if (!x || !xs_ || !wpk || !out || !os_ || B <= 0 || C <= 0 || H <= 0 || W <= 0 || N <= 0 || act < 0 || act > 2) return EL_ERR_ARG;
    if (dtype != EL_BF16 && dtype != EL_F16) return EL_ERR_UNSUPPORTED;
    if ((stride != 1 && stride != 2) || C % 8 || N % 8 || xs_[1] != 1 || os_[1] != 1 || !aligned16(x) || !aligned16(out)) return EL_ERR_UNSUPPORTED;
    for (int i = 0; i < 4; ++i)
        if (i != 1 && (xs_[i] % 8 || os_[i] % 8)) return EL_ERR_UNSUPPORTED;
    const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
    pw::Args A{};
    const int rb = pw::box_bytes_for(C);
    int nch = 0, w_row_bytes = 0;
    for (int tap = 0; tap < 9; ++tap)
        for (int c0 = 0; c0 < C; c0 += rb / 2) {
            if (nch >= pw::kMaxChunks) return EL_ERR_UNSUPPORTED;
            pw::Chunk& c = A.chunk[nch++];
            c.c0 = (uint16_t)c0; c.src = (uint8_t)tap; c.rb = (uint8_t)rb;
            w_row_bytes += rb;
        }
    A.nchunks = nch;
    // pixel tile = tw x th patch with tw * th = 128: the widest power of two that wastes the least of the last column tile
    int tw = 128;
    {
        int best = -1;
        for (int t = 4; t <= 128; t <<= 1) {
            if (t * stride > 256 || (128 / t) * stride > 256) continue;
            const int64_t covered = ceil_div(Wo, t) * t * (ceil_div(Ho, 128 / t) * (128 / t));
            if (best < 0 || covered < best || (covered == best && t > tw)) { best = (int)covered; tw = t; }
        }
    }
    A.spatial = 1; A.tw = tw; A.th = 128 / tw; A.stride = stride;
    A.tiles_x = (int)ceil_div(Wo, A.tw); A.tiles_y = (int)ceil_div(Ho, A.th);
    A.n_tiles_m = (int64_t)B * A.tiles_x * A.tiles_y;
    const int64_t M = A.n_tiles_m * pw::kTileM;
    A.n_tile = el_conv3x3_tile(N, C, M);
    if (A.n_tile <= 0) return EL_ERR_UNSUPPORTED;
    A.stream_w = C > 32;
    uint32_t w_off = 0;
    for (int i = 0; i < nch; ++i) {
        A.chunk[i].w_off = w_off;
        w_off += ((uint32_t)A.n_tile * A.chunk[i].rb + 1023u) & ~1023u;
    }
    A.w_bytes = w_off;
    A.a_stage_bytes = (uint32_t)pw::kTileM * rb;
    A.stage_bytes = A.a_stage_bytes + (A.stream_w ? (((uint32_t)A.n_tile * rb + 1023u) & ~1023u) : 0u);
    const int n_tiles = (int)ceil_div(N, A.n_tile);
    int ob = 64;
    while (ob > 16 && ob / 2 >= A.n_tile) ob >>= 1;
    if (n_tiles > 1) while (A.n_tile % ob) ob >>= 1;
    if (ob < 16) return EL_ERR_UNSUPPORTED;
    A.ob = ob;
    A.wpk = wpk; A.bias = bias; A.res = nullptr; A.res_pitch = 0; A.res_scale = 1.f;
    A.has_out2 = 0; A.split = N; A.groups = 1;
    A.M = M; A.N = N; A.act = act;
    uint32_t cols = 32;
    while (cols < 2u * A.n_tile) cols <<= 1;
    if (cols > 512) return EL_ERR_UNSUPPORTED;
    A.tmem_cols = cols;
    const size_t fixed = 1024 + (A.stream_w ? 0 : ((A.w_bytes + 1023u) & ~1023u)) + 2 * (size_t)pw::kTileM * ob * 2 + (((size_t)(A.n_tile + 64) * 4 + 127) & ~(size_t)127) + 40 +
                         16 * pw::kMaxStages + 16;
    int S = 0, per_sm = 1;
    for (int ps = 4; ps >= 1; --ps) {
        if (ps > (int)(512 / cols)) continue;
        const size_t budget = (size_t)227 * 1024 / ps - 1024;
        if (budget <= fixed) continue;
        const int s_fit = (int)((budget - fixed) / A.stage_bytes);
        if (s_fit >= 4 || (ps == 1 && s_fit >= 2)) { S = s_fit; per_sm = ps; break; }
    }
    if (S < 2) return EL_ERR_UNSUPPORTED;
    if (S > pw::kMaxStages) S = pw::kMaxStages;
    A.stages = S;
    const size_t smem = fixed + (size_t)S * A.stage_bytes;
    int64_t gx = (int64_t)kSMs * per_sm / n_tiles;
    if (gx < 1) gx = 1;
    if (gx > A.n_tiles_m) gx = A.n_tiles_m;
    if (!pw::make_map4(&A.src_map[0], x, C, W, H, B, xs_, rb, A.tw, A.th, stride, dtype)) return EL_ERR_CUDA;
    if (!pw::make_map4(&A.out_map, out, N, Wo, Ho, B, os_, ob * 2, A.tw, A.th, 1, dtype)) return EL_ERR_CUDA;
    dim3 grid((unsigned)gx, (unsigned)n_tiles);
    cudaStream_t st = (cudaStream_t)stream;
    const cudaError_t e = dtype == EL_BF16 ? pw::launch<__nv_bfloat16>(A, grid, smem, 0, st) : pw::launch<__half>(A, grid, smem, 0, st);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return EL_ERR_CUDA; }
    note_launches(1);
    return check_launch();
}
