// Dense 3 x 3 convolution (stride 1, padding 1) with WIDE inputs (C_in a multiple of 64) for the inference engine, as an implicit GEMM
// whose nine filter taps all read ONE haloed shared-memory tile:
//     out[b, y, x, n] = act( sum_{ky,kx,c} X[b, y+ky-1, x+kx-1, c] * W[n, c, ky, kx] + bias[n] )                NHWC 16-bit activations
// Replaces cuDNN + el_bias_act_fwd for Conv(k=3).forward_fuse (nn/modules/conv.py:58-60) at the sites el_conv3x3_fwd loses: the shared
// high-band conv f_h of _WaveletEnhancer (nn/modules/block.py:3668-3673) from c = 64 up and the box towers of the Detect head
// (nn/modules/head.py:59-63: Conv(x, 64, 3), Conv(64, 64, 3)).  el_conv3x3_fwd gives every tap its own tap-shifted TMA box, i.e. pulls the
// input nine times through the L2 -> SM path, which is what saturates from C_in = 64 (measured round 1: 108 us against cuDNN's 47 us at
// 64 -> 64 @ 80 x 80, batch 64).  Here the input crosses that path 1.43 times:
//   * pixel tile = 8 rows x 14 columns of one image, laid out as M = 128 accumulator rows with a PITCH of 16 (r = oy * 16 + ox; the two
//     columns ox = 14, 15 of every row are junk that is computed and never stored);
//   * per 64-channel K chunk ONE 4-D TMA box (64 channels, 16 x, 10 y, 1 image) starting at (x0 - 1, y0 - 1) lands as 160 rows of 128 B
//     (128-byte swizzle; out-of-image pixels are the zero padding, filled by the TMA unit);
//   * the A operand of tap (ky, kx) is that same tile read through a descriptor whose start address is moved by (ky * 16 + kx) rows:
//     row r of the MMA then reads tile row r + ky * 16 + kx = pixel (oy + ky, ox + kx).  The shifted start is not 1 KiB aligned; the
//     swizzle is a function of the absolute shared-memory address, so the plain start-address shift (base_offset 0) addresses the rows
//     TMA wrote -- verified on a B200 for exactly these shifts by tools/exp_umma_row_shift.cu (profiles/r02_exp_umma_row_shift.log);
//   * weights: per (tap, chunk) one K-major SW128 tile [N][64], all resident in shared memory (9 * C_in * N * 2 B <= 144 KB);
//   * warp roles as in pwconv.cu: TMA producer, MMA issuer (36 MMAs M128 x N x K16 per chunk), two to four epilogue GROUPS of four warps, each
//     with its own TMEM accumulator and staging tile, tiles round-robin (tcgen05.ld, bias,
//     SiLU / ReLU, 16-bit pack into a swizzled staging tile, one TMA store per output row of the tile: box (channels, 14, 1, 1), clipped at
//     the image border); double-buffered TMEM accumulators; persistent CTAs; PDL.
#include <cuda.h>
#include <stdlib.h>

#include <type_traits>

#include "el_common.cuh"

namespace el {
namespace c3 {

constexpr int kMaxEpi = 4;                        // epilogue groups of four warps (one per TMEM lane quarter), each with its own TMEM accumulator
constexpr int kThreads = 64 + 128 * kMaxEpi;      // TMA warp, MMA warp, up to 16 epilogue warps (the launch uses 64 + 128 * groups)
constexpr int kTW = 14, kTH = 8, kPitch = 16;     // output tile, pitch of the M rows
constexpr int kInRows = (kTH + 2) * kPitch;       // 160 rows of 128 B per K chunk
constexpr uint32_t kStageBytes = kInRows * 128 + 512;  // + 4 rows: tap (2, 2) reads two rows past the tile for the junk columns of the last row
constexpr int kMaxStages = 6;
constexpr int kMaxGroup = 4;

struct Args {
    CUtensorMap src_map, out_map;
    const void* wpk;       // [9 taps][C/64 chunks] tiles of n_pad x 128 B (SW128, K-major), each padded to 1 KiB
    const float* bias;
    int B, H, W, N, n_pad, chunks, act, stages, ob, groups, nstg;  // groups: epilogue groups = TMEM accumulators; nstg: staging tiles per group
    int tiles_x, tiles_y;
    int64_t n_tiles;
    uint32_t w_bytes, tile_w_bytes, tmem_cols;
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
// K-major SW128 descriptor: start >> 4, SBO = 1024 B (8 rows of 128 B), version 1, layout 2 (SWIZZLE_128B); base_offset 0 also for starts
// that are not 1 KiB aligned (tools/exp_umma_row_shift.cu, mode 0)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// the same descriptor from its low word ((address & 0x3FFFF) >> 4, plus a byte offset >> 4): the high word is constant
__device__ __forceinline__ uint64_t umma_desc_lo(uint32_t lo) { return ((uint64_t)0x40004040u << 32) | lo; }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
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

template <typename T, int ACT>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_halo_kernel(const __grid_constant__ Args A) {
    extern __shared__ __align__(1024) unsigned char sm_raw[];
    const uint32_t sbase = (smem_addr(sm_raw) + 1023u) & ~1023u;
    unsigned char* sm = sm_raw + (sbase - smem_addr(sm_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = A.stages, N = A.N, n_pad = A.n_pad, nch = A.chunks, ob = A.ob;
    // shared memory map: [weights][stage ring][groups x nstg staging tiles 128 x ob][bias][barriers]
    const int G = A.groups, nstg = A.nstg;
    const uint32_t off_ring = (A.w_bytes + 1023u) & ~1023u;
    const uint32_t stage_bytes = (kStageBytes + 1023u) & ~1023u;
    const uint32_t off_stage = off_ring + (uint32_t)S * stage_bytes;
    const uint32_t staging_bytes = 128u * ob * 2;
    const uint32_t off_bias = off_stage + (uint32_t)(G * nstg) * staging_bytes;
    const uint32_t off_bar = off_bias + (((uint32_t)(n_pad + 64) * 4 + 127) & ~127u);
    float* s_bias = reinterpret_cast<float*>(sm + off_bias);
    const uint32_t bar_w = sbase + off_bar;
    const uint32_t bar_acc_full = sbase + off_bar + 8;                  // [kMaxEpi] accumulator g complete
    const uint32_t bar_acc_empty = bar_acc_full + 8 * kMaxEpi;          // [kMaxEpi] accumulator g drained by its epilogue group
    const uint32_t bar_full = bar_acc_empty + 8 * kMaxEpi;              // [kMaxStages]
    const uint32_t bar_empty = bar_full + 8 * kMaxStages;               // [kMaxStages]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + off_bar + 8 + 16 * kMaxEpi + 16 * kMaxStages);

    const int64_t first = blockIdx.x;
    const int my_tiles = first < A.n_tiles ? (int)((A.n_tiles - first + gridDim.x - 1) / gridDim.x) : 0;
    const int per_img = A.tiles_x * A.tiles_y;
    constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
    const int nthreads = blockDim.x;

    pdl_launch_dependents();
    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&A.src_map) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&A.out_map) : "memory");
        mbar_init(bar_w, 1);
        for (int b = 0; b < kMaxEpi; ++b) { mbar_init(bar_acc_full + 8 * b, 1); mbar_init(bar_acc_empty + 8 * b, 128); }
        for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_w, A.w_bytes);
        const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(A.wpk);
        for (uint32_t o = 0; o < A.w_bytes; o += 32768) bulk_g2s(sbase + o, wsrc + o, min(32768u, A.w_bytes - o), bar_w);
    }
    for (int i = tid; i < n_pad + 64; i += nthreads) s_bias[i] = (A.bias && i < N) ? __ldg(A.bias + i) : 0.f;
    // the four rows behind every stage (read by tap (2, 2) for the junk columns of the last tile row) must hold finite numbers
    for (int i = tid; i < S * 32; i += nthreads)
        *reinterpret_cast<uint4*>(sm + off_ring + (uint32_t)(i >> 5) * stage_bytes + kInRows * 128 + (i & 31) * 16) = make_uint4(0, 0, 0, 0);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(s_tmem)), "r"(A.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------------------------------ TMA producer: one haloed box per K chunk
        const bool leader = elect_one();
        int s = 0, use = 0;
        for (int tl = 0; tl < my_tiles; ++tl) {
            const int64_t tile = first + (int64_t)tl * gridDim.x;
            const int img = (int)(tile / per_img), r = (int)(tile % per_img);
            const int y0 = (r / A.tiles_x) * kTH - 1, x0 = (r % A.tiles_x) * kTW - 1;
            for (int c = 0; c < nch; ++c) {
                if (use > 0) mbar_wait(bar_empty + 8 * s, (uint32_t)(use - 1) & 1);
                if (leader) {
                    mbar_expect_tx(bar_full + 8 * s, kInRows * 128);
                    tma_load_4d(sbase + off_ring + (uint32_t)s * stage_bytes, &A.src_map, c * 64, x0, y0, img, bar_full + 8 * s);
                }
                __syncwarp();
                if (++s == S) { s = 0; ++use; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------------------------ MMA issuer: 9 taps x 4 K steps per chunk.
        // All MMAs of a tile accumulate into ONE TMEM accumulator, back to back; tile tl uses accumulator tl % G, drained by epilogue group tl % G.
        // (A variant that interleaved the MMAs of several tiles measured 3.5 - 4x slower and was dropped.)
        // Issue cost matters here (36 MMAs per chunk): the WHOLE warp walks the loop, so every address below is warp-uniform and lives in the
        // uniform datapath (UIADD3 on the low descriptor word, no register -> uniform-register moves, no per-thread waterfall around
        // UTCHMMA -- which is what `if (lane == 0) { loop }` compiles to: ~19 instructions per MMA); one elected lane issues MMAs and commits.
        {
            const uint32_t idesc = umma_idesc(kFmt, 128, n_pad);
            const bool leader = elect_one();
            const uint32_t b_tap_step = ((uint32_t)nch * A.tile_w_bytes) >> 4;
            mbar_wait(bar_w, 0);
            int s = 0, b = 0;
            uint32_t s_par = 0, ub = 0;
            for (int tl = 0; tl < my_tiles; ++tl) {
                if (ub > 0) mbar_wait(bar_acc_empty + 8 * b, (ub - 1) & 1);
                tc_fence_after();
                const uint32_t d = tmem + (uint32_t)b * n_pad;
                for (int c = 0; c < nch; ++c) {
                    mbar_wait(bar_full + 8 * s, s_par);
                    tc_fence_after();
                    const uint32_t a_lo = ((sbase + off_ring + (uint32_t)s * stage_bytes) & 0x3FFFF) >> 4;
                    const uint32_t b_lo = ((sbase + (uint32_t)c * A.tile_w_bytes) & 0x3FFFF) >> 4;
                    if (leader) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)  // tap (ky, kx): the tile rows shifted by ky * 16 + kx; 32 bytes along K per step
                                umma(d, umma_desc_lo(a_lo + (uint32_t)(((tap / 3) * kPitch + tap % 3) * 128 + 32 * ks) / 16),
                                     umma_desc_lo(b_lo + (uint32_t)tap * b_tap_step + 2 * ks), idesc, (c > 0 || tap > 0 || ks > 0) ? 1u : 0u);
                        }
                        umma_commit(bar_empty + 8 * s);
                        if (c == nch - 1) umma_commit(bar_acc_full + 8 * b);
                    }
                    __syncwarp();
                    if (++s == S) { s = 0; s_par ^= 1; }
                }
                if (++b == G) { b = 0; ++ub; }
            }
        }
    } else {
        // ------------------------------------------------------------------------------------ epilogue groups
        // A tile's epilogue is a latency chain (accumulator wait -> tcgen05.ld -> bias / SiLU -> pack -> staging -> TMA stores), and with resident
        // weights + ring only ONE CTA fits an SM, so nothing else hides it: ncu of the eight-warp, one-tile-at-a-time form showed the tensor pipe
        // 31 % active and the epilogue warps 37 % of their time at their barrier (profiles/r02f_ncu_halo_*.txt).  Hence G independent groups of four
        // warps (one per TMEM lane quarter), group g owning accumulator g and the tiles tl = g (mod G): up to G tiles' epilogues are in flight.
        const int g = (warp - 2) >> 2, q = warp & 3, row = q * 32 + lane, gw = (warp - 2) & 3;
        const int rbo = ob * 2;
        const uint32_t swz = ((uint32_t)(row * rbo) >> 7) & (uint32_t)(rbo / 16 - 1);
        const bool store_leader = gw == 0 && elect_one();   // issues (and therefore tracks) this group's TMA stores
        const uint32_t bar_id = 1 + g;
        int sub = 0;
        uint32_t use = 0;
        for (int tl = g; tl < my_tiles; tl += G, ++use) {
            mbar_wait(bar_acc_full + 8 * g, use & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + (uint32_t)g * n_pad + ((uint32_t)(q * 32) << 16);
            for (int c0 = 0; c0 < N; c0 += ob, ++sub) {
                const uint32_t stg = sbase + off_stage + (uint32_t)(g * nstg + (nstg == 2 ? (sub & 1) : 0)) * staging_bytes;
                // the stores that last read this staging tile must be done with it: one tile per group -> all of the group's earlier stores;
                // two tiles per group -> all but the latest commit
                if (store_leader) { if (nstg == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                const int jmax = min(ob, n_pad - c0);
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
                        if (ACT == 1) {  // SiLU: x * sigmoid(x) = h + h * tanh(h), h = x / 2
#pragma unroll
                            for (int e = 0; e < 16; ++e) { const float h = 0.5f * f[e]; f[e] = fmaf(h, tanh_fast(h), h); }
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
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)row * rbo + chunk * 16), "r"(o.x), "r"(o.y),
                                         "r"(o.z), "r"(o.w)
                                         : "memory");
                        }
                    }
                }
                if (c0 + ob >= N) {  // last box of the tile: the accumulator has been read completely
                    tc_fence_before();
                    mbar_arrive(bar_acc_empty + 8 * g);
                }
                proxy_fence();
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                if (gw == 0) {  // one store per output row of the tile: 14 of its 16 accumulator rows; rows / columns past the image are clipped
                    const int64_t tile = first + (int64_t)tl * gridDim.x;
                    const int img = (int)(tile / per_img), r = (int)(tile % per_img);
                    const int y0 = (r / A.tiles_x) * kTH, x0 = (r % A.tiles_x) * kTW;
                    if (store_leader) {
                        for (int oy = 0; oy < kTH; ++oy)
                            if (y0 + oy < A.H) tma_store_4d(&A.out_map, c0, x0, y0 + oy, img, stg + (uint32_t)(oy * kPitch) * rbo);
                        bulk_commit();
                    }
                    __syncwarp();
                }
            }
        }
        if (store_leader) bulk_wait_read<0>();
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
// (channels, x, y, image) view of an NHWC activation
static bool make_map4(CUtensorMap* map, const void* base, int channels, int W, int H, int B, const int64_t st[4] /* n, c, h, w elements */, int box_c,
                      int box_x, int box_y, int dtype) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)channels, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)st[3] * 2, (cuuint64_t)st[2] * 2, (cuuint64_t)st[0] * 2};
    const cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_x, (cuuint32_t)box_y, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const int rb = box_c * 2;
    const CUtensorMapSwizzle sw = rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    return fn(map, dtype == EL_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box,
              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// shared-memory / TMEM plan of a site: returns the ring depth (0 = does not fit).  groups = epilogue groups = TMEM accumulators: the most (<= 4)
// whose staging tiles still leave a ring of three stages, else two (accumulator double buffering, as before).
static int plan(int C, int N, int& n_pad, int& ob, uint32_t& w_bytes, uint32_t& tile_w_bytes, int& groups, int& nstg, uint32_t& tmem_cols, size_t& smem) {
    n_pad = (int)ceil_div(N, 16) * 16;
    tile_w_bytes = ((uint32_t)n_pad * 128 + 1023u) & ~1023u;
    w_bytes = 9u * (uint32_t)(C / 64) * tile_w_bytes;
    ob = 64;
    while (ob > 16 && ob / 2 >= n_pad) ob >>= 1;
    const int nstg_max = N > ob ? 2 : 1;   // several store boxes per tile: two staging tiles per group let box i + 1 be packed while box i is stored
    const uint32_t stage_bytes = (kStageBytes + 1023u) & ~1023u;
    const size_t budget = (size_t)227 * 1024;
    static const int g_max = [] { const char* v = getenv("EL_C3_GROUPS"); const int g = v ? atoi(v) : kMaxEpi; return g < 2 ? 2 : (g > kMaxEpi ? kMaxEpi : g); }();
    for (int pass = 0; pass < 2; ++pass)   // second pass: one staging tile per group even with several boxes per tile (wide N, large weights)
    for (int G = g_max; G >= 2; --G) {
        nstg = pass == 0 ? nstg_max : 1;
        if (pass == 1 && nstg_max == 1) break;
        uint32_t cols = 32;
        while (cols < (uint32_t)(G * n_pad)) cols <<= 1;
        if (cols > 512) continue;
        const size_t fixed = 1024 + ((w_bytes + 1023u) & ~1023u) + (size_t)(G * nstg) * 128 * ob * 2 + (((size_t)(n_pad + 64) * 4 + 127) & ~(size_t)127) + 8 +
                             16 * kMaxEpi + 16 * kMaxStages + 16;
        if (fixed + 2 * stage_bytes > budget) continue;
        int S = (int)((budget - fixed) / stage_bytes);
        if (S < 3 && G > 2) continue;
        if (S > kMaxStages) S = kMaxStages;
        groups = G; tmem_cols = cols;
        smem = fixed + (size_t)S * stage_bytes;
        return S;
    }
    return 0;
}

}  // namespace c3
}  // namespace el

using namespace el;

/* 1 if el_conv3x3_halo_fwd covers a (C_in, N) site: 16-bit, C_in a multiple of 64, N a multiple of 8 up to 256, nine weight tiles resident. */
extern "C" int el_conv3x3_halo_ok(int C, int N) {
    if (C < 64 || C % 64 || N <= 0 || N % 8 || N > 256) return 0;
    int n_pad, ob, groups, nstg; uint32_t wb, tb, cols; size_t smem;
    return c3::plan(C, N, n_pad, ob, wb, tb, groups, nstg, cols, smem) >= 2 ? 1 : 0;
}

extern "C" int el_conv3x3_halo_fwd(const void* x, const int64_t xs_[4], int C, const void* wpk, const float* bias, void* out, const int64_t os_[4], int B,
                                   int H, int W, int N, int act, int dtype, void* stream) {
    if (!x || !xs_ || !wpk || !out || !os_ || B <= 0 || C <= 0 || H <= 0 || W <= 0 || N <= 0 || act < 0 || act > 2) return EL_ERR_ARG;
    if (dtype != EL_BF16 && dtype != EL_F16) return EL_ERR_UNSUPPORTED;
    if (!el_conv3x3_halo_ok(C, N) || xs_[1] != 1 || os_[1] != 1 || !aligned16(x) || !aligned16(out)) return EL_ERR_UNSUPPORTED;
    for (int i = 0; i < 4; ++i)
        if (i != 1 && (xs_[i] % 8 || os_[i] % 8)) return EL_ERR_UNSUPPORTED;
    c3::Args A{};
    size_t smem = 0;
    A.stages = c3::plan(C, N, A.n_pad, A.ob, A.w_bytes, A.tile_w_bytes, A.groups, A.nstg, A.tmem_cols, smem);
    if (A.stages < 2) return EL_ERR_UNSUPPORTED;
    A.wpk = wpk; A.bias = bias; A.B = B; A.H = H; A.W = W; A.N = N; A.chunks = C / 64; A.act = act;
    A.tiles_x = (int)ceil_div(W, c3::kTW); A.tiles_y = (int)ceil_div(H, c3::kTH);
    A.n_tiles = (int64_t)B * A.tiles_x * A.tiles_y;
    if (!c3::make_map4(&A.src_map, x, C, W, H, B, xs_, 64, c3::kPitch, c3::kTH + 2, dtype)) return EL_ERR_CUDA;
    if (!c3::make_map4(&A.out_map, out, N, W, H, B, os_, A.ob, c3::kTW, 1, dtype)) return EL_ERR_CUDA;
    int64_t gx = kSMs < A.n_tiles ? kSMs : A.n_tiles;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
#define EL_C3_LAUNCH(TT, ACT)                                                                                                                  \
    {                                                                                                                                          \
        e = cudaFuncSetAttribute(c3::conv3x3_halo_kernel<TT, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);                   \
        if (e == cudaSuccess) e = launch_pdl(c3::conv3x3_halo_kernel<TT, ACT>, dim3((unsigned)gx), dim3(64 + 128 * A.groups), smem, st, A);            \
    }
    if (dtype == EL_BF16) {
        if (act == 0) EL_C3_LAUNCH(__nv_bfloat16, 0) else if (act == 1) EL_C3_LAUNCH(__nv_bfloat16, 1) else EL_C3_LAUNCH(__nv_bfloat16, 2)
    } else {
        if (act == 0) EL_C3_LAUNCH(__half, 0) else if (act == 1) EL_C3_LAUNCH(__half, 1) else EL_C3_LAUNCH(__half, 2)
    }
#undef EL_C3_LAUNCH
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return EL_ERR_CUDA; }
    note_launches(1);
    return check_launch();
}
