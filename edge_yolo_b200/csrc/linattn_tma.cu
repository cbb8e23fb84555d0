// Linear-attention core for 16-bit NHWC activations, resident form (N <= 512 tokens: every EdgeLine model at 640 x 640, where the
// C2PSA block sees a 20 x 20 map).  Same maths as linattn.cu / linattn_tc.cu (LinearAttention.forward, nn/modules/block.py:3364-3372):
//     ksm = softmax_d(k)   q' = softmax_N(q)   ctx = ksm^T v  (64 x 64)   y = q' ctx
// evaluated as  P_c = exp(q - m_c) per 128-token chunk c (m_c = per-channel max of the chunk), s = sum_c exp(m_c - m) colsum(P_c),
//               y_c = P_c * (diag(exp(m_c - m) / s) ctx),        m = max_c m_c
// so q is read ONCE (the chunk-local exponentials are kept; the global max / sum only rescale the 64 x 64 context).
//
// Why a second tcgen05 kernel: linattn_tc.cu ran one 4-warp CTA per (image, head) whose threads fetched every operand row with their own
// global loads and walked the chunks one after the other -- one warp per scheduler, so its time was instruction count x single-warp latency
// (ncu, round 1: 228 registers, 6 % occupancy, issue slots 20 % busy, 28 us for 26 MB = 0.14 of the HBM roofline).  Here:
//   * warp 16 (one lane)  issues ALL operand tiles of the problem with 3-D TMA loads up front (12 x 16 KB boxes of 64 channels x 128 tokens,
//                         128-byte swizzle, out-of-range tokens zero-filled): the whole 154 KB working set is in flight at once;
//   * warps 0-15          four warpgroups, one per token chunk, work concurrently (4 warps per scheduler instead of 1): softmax_d(K) in place in
//                         the swizzled tile (thread <-> token), then the chunk's column statistics and P_c in place in the Q tile (thread <->
//                         8 channels x 8 tokens: packed 16-bit max, 16-lane shuffles), the rescaled context tile, and the epilogue;
//   * warp 17 (one lane)  issues the MMAs: GEMM1 reads ksm^T and V straight from the TMA-written tiles as MN-major SW128 operands (V is never
//                         touched by a thread), GEMM2 reads P_c (K-major SW128, the same bytes the Q tile occupied) and the context tile.
// TMEM: D1 = ctx (64 lanes x 64 columns), D2_c = y of chunk c (128 lanes x 64 columns each): 320 of 512 columns.
#include <cuda.h>

#include <type_traits>

#include "el_common.cuh"

namespace el {

struct AttnArgs {
    const void* qkv; int64_t qb, qc, qn;
    void* y; int64_t yb, yc, yn;
    int heads, N;
};

namespace ta {

constexpr int kD = 64;
constexpr int kChunk = 128;
constexpr int kMaxChunks = 4;
constexpr uint32_t kTile = kD * kChunk * 2;            // 16 KiB
constexpr int kThreads = 18 * 32;                      // 16 compute warps + TMA warp + MMA warp
constexpr uint32_t kOffK = 0, kOffV = kMaxChunks * kTile, kOffQ = 2 * kMaxChunks * kTile;
constexpr uint32_t kOffStat = 3 * kMaxChunks * kTile;  // cmax[4][64], csum[4][64] fp32
constexpr uint32_t kOffBar = kOffStat + 2 * kMaxChunks * kD * 4;
constexpr uint32_t kSmemBytes = kOffBar + 256 + 1024;  // barriers + tmem address; + slack for the 1024 B alignment of the tiles
constexpr uint32_t kTmemCols = 512;
constexpr float kLog2e = 1.4426950408889634f;

struct Args {
    CUtensorMap qkv_map;  // (3C channels, N tokens, B images), box (64, 128, 1), SWIZZLE_128B
    void* y; int64_t yb, yn;
    int heads, N, C;
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(map),
                 "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// 128-byte-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version = 1
// [46,48), layout [61,64) = 2 (SWIZZLE_128B).  Rows (8 per swizzle atom) are 128 B apart, so SBO = 1024 B in both uses:
//   K-major  (GEMM2's A = P_c: rows = tokens, the 64 channels of a row are the K dim): a K step of 16 channels = +32 B on the start address;
//   MN-major (GEMM1's ksm^T and V, GEMM2's context tile: rows = K index, the 64 MN elements of a row contiguous): a K step of 16 rows = +2048 B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(kTile >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t umma_idesc(int fmt, int a_mn_major, int b_mn_major, int M, int N) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// one lane of a converged warp; the loops around it stay warp-uniform, so descriptors / coordinates live in uniform registers and the
// UTMALDG / UTCHMMA issue is not wrapped in a per-thread register -> uniform-register waterfall (what `if (lane == 0) { loop }` compiles to)
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// packed 16-bit pair maximum (HMNMX2): exact, and half the instructions of an fp32 max after unpacking
template <typename T> __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ uint32_t max2<__nv_bfloat16>(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}
template <> __device__ __forceinline__ uint32_t max2<__half>(uint32_t a, uint32_t b) {
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}
template <typename T> __device__ __forceinline__ uint4 max2x4(uint4 a, uint4 b) {
    return make_uint4(max2<T>(a.x, b.x), max2<T>(a.y, b.y), max2<T>(a.z, b.z), max2<T>(a.w, b.w));
}
template <typename T> __device__ __forceinline__ uint32_t neg_inf2() { return std::is_same<T, __nv_bfloat16>::value ? 0xFF80FF80u : 0xFC00FC00u; }

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) linattn_tma_kernel(const __grid_constant__ Args A) {
    extern __shared__ __align__(1024) unsigned char sm_raw[];
    const uint32_t sbase = (smem_addr(sm_raw) + 1023u) & ~1023u;  // swizzle atoms need 1024 B alignment
    unsigned char* sm = sm_raw + (sbase - smem_addr(sm_raw));
    float* s_cmax = reinterpret_cast<float*>(sm + kOffStat);
    float* s_csum = s_cmax + kMaxChunks * kD;
    const uint32_t bar_kq = sbase + kOffBar;           // [4] K + Q tiles of chunk c have landed
    const uint32_t bar_v = bar_kq + 32;                // [4] V tile
    const uint32_t bar_kready = bar_v + 32;            // [4] softmax_d(K) written back by the chunk's 128 threads
    const uint32_t bar_pready = bar_kready + 32;       // [4] P_c and the chunk's context tile written
    const uint32_t bar_d2 = bar_pready + 32;           // [4] GEMM2 of chunk c complete
    const uint32_t bar_g1 = bar_d2 + 32;               // GEMM1 (all chunks) complete
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + kOffBar + 176);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = (int)blockIdx.x / A.heads, head = (int)blockIdx.x % A.heads;
    const int N = A.N, C = A.C;
    const int n_chunks = (N + kChunk - 1) / kChunk;
    constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;

    pdl_launch_dependents();
    if (warp == 16) {
        // ------------------------------------------------------------------------------------ TMA producer: the whole problem at once, and first:
        // this lane initialises the load barriers itself and issues all twelve tile loads before the CTA-wide set-up barrier, so the operand
        // fetch (the ~3 us HBM phase of this kernel) overlaps the TMEM allocation and the other barriers' initialisation
        const bool leader = elect_one();
        if (leader) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&A.qkv_map) : "memory");
            for (int c = 0; c < kMaxChunks; ++c) { mbar_init(bar_kq + 8 * c, 1); mbar_init(bar_v + 8 * c, 1); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        pdl_wait();  // the qkv GEMM that produced our input has completed from here on
        for (int c = 0; c < n_chunks; ++c) {  // K and Q first (the compute warps start on them), V is only read by the tensor core
            if (leader) {
                mbar_expect_tx(bar_kq + 8 * c, 2 * kTile);
                tma_load_3d(sbase + kOffK + c * kTile, &A.qkv_map, C + head * kD, c * kChunk, b, bar_kq + 8 * c);
                tma_load_3d(sbase + kOffQ + c * kTile, &A.qkv_map, head * kD, c * kChunk, b, bar_kq + 8 * c);
            }
        }
        for (int c = 0; c < n_chunks; ++c) {
            if (leader) {
                mbar_expect_tx(bar_v + 8 * c, kTile);
                tma_load_3d(sbase + kOffV + c * kTile, &A.qkv_map, 2 * C + head * kD, c * kChunk, b, bar_v + 8 * c);
            }
        }
    } else if (tid == 0) {
        for (int c = 0; c < kMaxChunks; ++c) {
            mbar_init(bar_kready + 8 * c, 128); mbar_init(bar_pready + 8 * c, 128);
            mbar_init(bar_d2 + 8 * c, 1);
        }
        mbar_init(bar_g1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 17) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(s_tmem)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t tmem_d1 = tmem, tmem_d2 = tmem + kD;
    pdl_wait();  // every thread that stores to global memory orders itself behind the previous grid as well

    if (warp == 16) {
        // (loads already issued above)
    } else if (warp == 17) {
        // ------------------------------------------------------------------------------------ MMA issuer
        {
            const bool leader = elect_one();
            const uint32_t idesc1 = umma_idesc(kFmt, 1, 1, 64, 64);   // ctx[i][j] += ksm[n][i] v[n][j]: both operands MN-major, K = tokens
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(bar_kready + 8 * c, 0);
                mbar_wait(bar_v + 8 * c, 0);
                tc_fence_after();
                const int nvalid = min(kChunk, N - c * kChunk);
                const int ksteps = (nvalid + 15) >> 4;  // rows past N are zero in V (TMA fill): they add nothing
                const uint64_t dk = umma_desc_sw128(sbase + kOffK + c * kTile), dv = umma_desc_sw128(sbase + kOffV + c * kTile);
                if (leader) {
                    for (int ks = 0; ks < ksteps; ++ks)  // a K step of 16 token rows = +2048 B on the start address (>> 4 in the descriptor)
                        umma(tmem_d1, dk + (uint64_t)(ks * 128), dv + (uint64_t)(ks * 128), idesc1, (c > 0 || ks > 0) ? 1u : 0u);
                    if (c == n_chunks - 1) umma_commit(bar_g1);
                }
                __syncwarp();
            }
            const uint32_t idesc2 = umma_idesc(kFmt, 0, 1, 128, 64);  // y[n][j] = P_c[n][i] ctx_c[i][j]: A K-major, B MN-major, K = 64 channels
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(bar_pready + 8 * c, 0);
                tc_fence_after();
                const uint64_t dq = umma_desc_sw128(sbase + kOffQ + c * kTile), dk = umma_desc_sw128(sbase + kOffK + c * kTile);
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma(tmem_d2 + c * kD, dq + (uint64_t)(ks * 2), dk + (uint64_t)(ks * 128), idesc2, ks > 0 ? 1u : 0u);
                    umma_commit(bar_d2 + 8 * c);
                }
                __syncwarp();
            }
        }
    } else if ((warp >> 2) < n_chunks) {
        // ------------------------------------------------------------------------------------ compute: warpgroup wg owns token chunk wg
        const int wg = warp >> 2, t = tid & 127, quarter = warp & 3;
        const int n_base = wg * kChunk;
        const uint32_t kt = sbase + kOffK + wg * kTile, qt = sbase + kOffQ + wg * kTile;
        mbar_wait(bar_kq + 8 * wg, 0);
        {   // softmax over the 64 channels of token t, in place (thread <-> token row; the swizzle makes the row reads conflict-free)
            uint4 r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = lds128(kt + t * 128 + ((j ^ (t & 7)) << 4));
            uint4 pm = r[0];
#pragma unroll
            for (int j = 1; j < 8; ++j) pm = max2x4<T>(pm, r[j]);
            const uint32_t p2 = max2<T>(max2<T>(pm.x, pm.y), max2<T>(pm.z, pm.w));
            float m2[2];
            {
                float m8[8];
                unpack<T>(make_uint4(p2, p2, p2, p2), m8);
                m2[0] = m8[0]; m2[1] = m8[1];
            }
            const float ml = fmaxf(m2[0], m2[1]) * kLog2e;
            float f[kD];
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float t8[8];
                unpack<T>(r[j], t8);
#pragma unroll
                for (int e = 0; e < 8; ++e) { f[8 * j + e] = ex2_approx(fmaf(t8[e], kLog2e, -ml)); s += f[8 * j + e]; }
            }
            const float inv = 1.f / s;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float t8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) t8[e] = f[8 * j + e] * inv;
                sts128(kt + t * 128 + ((j ^ (t & 7)) << 4), pack<T>(t8));
            }
        }
        proxy_fence();  // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(bar_kready + 8 * wg);

        // column statistics of q over this chunk and P_c = exp(q - m_c), in place: thread <-> channel group g (8 channels = one 16 B unit) x
        // token subset s (tokens s, s + 16, ..): 8 x 8 values per thread
        const int g = t >> 4, sub = t & 15;
        const uint32_t qaddr = qt + sub * 128 + ((g ^ (sub & 7)) << 4);  // token k * 16 + sub: + k * 2048 (the swizzle phase only depends on sub)
        uint4 rq[8];
        const uint32_t ninf = neg_inf2<T>();
#pragma unroll
        for (int k = 0; k < 8; ++k)  // tokens past N count as -inf: they drop out of the max, and exp gives exact zeros for the sum
            rq[k] = (n_base + k * 16 + sub < N) ? lds128(qaddr + k * 2048) : make_uint4(ninf, ninf, ninf, ninf);
        uint4 pm = rq[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) pm = max2x4<T>(pm, rq[k]);
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            uint4 other;
            other.x = __shfl_xor_sync(0xffffffffu, pm.x, o); other.y = __shfl_xor_sync(0xffffffffu, pm.y, o);
            other.z = __shfl_xor_sync(0xffffffffu, pm.z, o); other.w = __shfl_xor_sync(0xffffffffu, pm.w, o);
            pm = max2x4<T>(pm, other);
        }
        float m8[8], ml8[8], sum8[8];
        unpack<T>(pm, m8);
#pragma unroll
        for (int e = 0; e < 8; ++e) { ml8[e] = m8[e] * kLog2e; sum8[e] = 0.f; }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float t8[8];
            unpack<T>(rq[k], t8);
#pragma unroll
            for (int e = 0; e < 8; ++e) { t8[e] = ex2_approx(fmaf(t8[e], kLog2e, -ml8[e])); sum8[e] += t8[e]; }
            sts128(qaddr + k * 2048, pack<T>(t8));
        }
#pragma unroll
        for (int o = 1; o < 16; o <<= 1)
#pragma unroll
            for (int e = 0; e < 8; ++e) sum8[e] += __shfl_xor_sync(0xffffffffu, sum8[e], o);
        if (sub == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { s_cmax[wg * kD + g * 8 + e] = m8[e]; s_csum[wg * kD + g * 8 + e] = sum8[e]; }
        }
        proxy_fence();
        asm volatile("bar.sync 1, %0;" ::"r"(n_chunks * 128) : "memory");  // every chunk's (m_c, colsum) is visible

        // context tile of this chunk: ctx_c[i][j] = ctx[i][j] * exp(m_c[i] - m[i]) / s[i]  (MN-major SW128 B operand of GEMM2: row = i, the 64 j
        // contiguous).  It replaces the K tile, which the tensor core has finished reading once GEMM1 has been committed.
        float fac = 0.f;
        const int i = quarter * 16 + lane;  // accumulator row of an M = 64 tile held by this lane (lanes 0-15 of each 32-lane quarter)
        if (lane < 16) {
            float m = -INFINITY;
            for (int c = 0; c < n_chunks; ++c) m = fmaxf(m, s_cmax[c * kD + i]);
            float s = 0.f;
            for (int c = 0; c < n_chunks; ++c) s += s_csum[c * kD + i] * ex2_approx((s_cmax[c * kD + i] - m) * kLog2e);
            fac = ex2_approx((s_cmax[wg * kD + i] - m) * kLog2e) / s;
        }
        mbar_wait(bar_g1, 0);
        tc_fence_after();
        {
            uint32_t v0[32], v1[32];
            const uint32_t taddr = tmem_d1 + ((uint32_t)(quarter * 32) << 16);
            tmem_ld32(taddr, v0);
            tmem_ld32(taddr + 32, v1);
            if (lane < 16) {
#pragma unroll
                for (int jc = 0; jc < 8; ++jc) {
                    float t8[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) t8[e] = __uint_as_float(jc < 4 ? v0[8 * jc + e] : v1[8 * (jc - 4) + e]) * fac;
                    sts128(kt + i * 128 + ((jc ^ (i & 7)) << 4), pack<T>(t8));
                }
            }
        }
        proxy_fence();
        tc_fence_before();
        mbar_arrive(bar_pready + 8 * wg);

        // epilogue: lane <-> token row 32 * quarter + lane of D2_c, 64 fp32 columns -> 128 B of 16-bit channels
        mbar_wait(bar_d2 + 8 * wg, 0);
        tc_fence_after();
        {
            const int n = n_base + quarter * 32 + lane;
            T* dst = reinterpret_cast<T*>(A.y) + (int64_t)b * A.yb + (int64_t)n * A.yn + head * kD;
            const uint32_t taddr = tmem_d2 + wg * kD + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t v[32];
                tmem_ld32(taddr + 32 * h, v);
                if (n < N) {
#pragma unroll
                    for (int jc = 0; jc < 4; ++jc) {
                        float t8[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) t8[e] = __uint_as_float(v[8 * jc + e]);
                        *reinterpret_cast<uint4*>(dst + 32 * h + 8 * jc) = pack<T>(t8);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 17) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
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

}  // namespace ta

bool linattn_tma_supported(const AttnArgs& A, int dtype) {
    if (dtype != EL_BF16 && dtype != EL_F16) return false;
    if (A.qc != 1 || A.yc != 1 || A.N > ta::kMaxChunks * ta::kChunk) return false;  // channel-contiguous (NHWC) views, all chunks resident
    return aligned16(A.qkv) && aligned16(A.y) && A.qn % 8 == 0 && A.qb % 8 == 0 && A.yn % 8 == 0 && A.yb % 8 == 0;
}

int linattn_tma_launch(const AttnArgs& A, int B, int dtype, cudaStream_t s) {
    ta::EncodeTiledFn fn = ta::encode_fn();
    if (!fn) return EL_ERR_CUDA;
    ta::Args P{};
    const int C = A.heads * ta::kD;
    const cuuint64_t dims[3] = {(cuuint64_t)(3 * C), (cuuint64_t)A.N, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)A.qn * 2, (cuuint64_t)(B > 1 ? A.qb : (int64_t)A.N * A.qn) * 2};
    const cuuint32_t box[3] = {(cuuint32_t)ta::kD, (cuuint32_t)ta::kChunk, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (fn(&P.qkv_map, dtype == EL_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(A.qkv), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return EL_ERR_CUDA;
    P.y = A.y; P.yb = A.yb; P.yn = A.yn; P.heads = A.heads; P.N = A.N; P.C = C;
    cudaError_t e;
    if (dtype == EL_BF16) {
        e = cudaFuncSetAttribute(ta::linattn_tma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ta::kSmemBytes);
        if (e == cudaSuccess) e = launch_pdl(ta::linattn_tma_kernel<__nv_bfloat16>, dim3((unsigned)(B * A.heads)), dim3(ta::kThreads), ta::kSmemBytes, s, P);
    } else {
        e = cudaFuncSetAttribute(ta::linattn_tma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ta::kSmemBytes);
        if (e == cudaSuccess) e = launch_pdl(ta::linattn_tma_kernel<__half>, dim3((unsigned)(B * A.heads)), dim3(ta::kThreads), ta::kSmemBytes, s, P);
    }
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return EL_ERR_CUDA; }
    note_launches(1);
    return check_launch();
}

}  // namespace el
