// tcgen05 / TMEM path of the linear-attention core for 16-bit activations (bf16 / fp16), NHWC.
// Same maths as linattn.cu (LinearAttention.forward, nn/modules/block.py:3364-3372); the two contractions run
// on the 5th-generation tensor cores with fp32 accumulators in tensor memory:
//   GEMM1  ctx[i][j]  = sum_n softmax_d(K)[n][i] * V[n][j]     M=64 (i), N=64 (j), K = tokens   -> TMEM cols [0,64)
//   GEMM2  y[n][j]    = sum_i exp(q[n][i]-max_i) * ctx[i][j]/s_i  M=128 tokens, N=64, K=64        -> TMEM cols [64,128)
// One CTA (4 warps) per (image, head) -- or (EL_LINATTN_CLUSTER=1, measured slower at the benchmark shape, see linattn_tc_launch)
// a CLUSTER of 4 CTAs per (image, head): CTA r takes the chunks r, r+4, ...; the partial ctx (64 x 64 fp32) and the per-channel
// (max, sum) of q are exchanged through distributed shared memory (mapa + ld.shared::cluster between two barrier.cluster phases), every
// CTA then normalises the summed ctx itself and runs GEMM2 on its own chunks.  Operands are written to shared memory by the CTA's own threads in the
// canonical no-swizzle UMMA layouts ("chunk-major": 16-byte chunks of 8 elements, [chunk of the 64-wide dim][token][8]):
//   softmax(K)^T and V  : MN-major (channels contiguous), K dim = tokens, LBO = 128 B, SBO = 2048 B
//   P = exp(q - max)    : K-major  (channels = K dim),                   LBO = 2048 B, SBO = 128 B
//   ctx' (as B of GEMM2): K-major  [chunk of i][j][8],                   LBO = 1024 B, SBO = 128 B
// so every warp-wide shared store is 512 contiguous bytes.  MMAs are issued by one thread, completion is tracked with
// tcgen05.commit -> mbarrier, accumulators come back with tcgen05.ld (32 lanes x 32 bit x 64 columns per warp).
#include <type_traits>

#include "el_common.cuh"

namespace el {

struct AttnArgs {
    const void* qkv; int64_t qb, qc, qn;
    void* y; int64_t yb, yc, yn;
    int heads, N;
};

namespace tc {

constexpr int kD = 64;
constexpr int kChunk = 128;                       // tokens per operand tile
constexpr uint32_t kTileBytes = kD * kChunk * 2;  // 16 KiB: one 64 x 128 bf16 operand tile
constexpr uint32_t kTmemCols = 128;               // D1 (64) + D2 (64)

// shared memory map (bytes)
constexpr uint32_t kOffA0 = 0, kOffB0 = kTileBytes, kOffA1 = 2 * kTileBytes, kOffB1 = 3 * kTileBytes;
constexpr uint32_t kOffQ = 4 * kTileBytes;            // raw q tile for the column statistics (128 rows x 128 B, chunk-swizzled)
constexpr uint32_t kOffCtx = kOffQ + kTileBytes;      // ctx' bf16, 8 KiB
constexpr uint32_t kOffStat = kOffCtx + kD * kD * 2;  // max[64], sum[64] fp32
constexpr uint32_t kOffBar = kOffStat + 6 * kD * 4;   // (max, sum, 4 x 64 partials) then 3 mbarriers + tmem address
constexpr uint32_t kSmemBytes = kOffBar + 64;
constexpr uint32_t kOffCtxP = (kSmemBytes + 15) & ~15u;        // cluster variant: this CTA's partial ctx, fp32 [64][64]
constexpr uint32_t kOffGStat = kOffCtxP + kD * kD * 4;          // cluster variant: global max[64], 1 / sum[64]
constexpr uint32_t kSmemBytesCluster = kOffGStat + 2 * kD * 4;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// address of the same shared-memory location in CTA `rank` of the cluster, and a load through it
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ float ld_cluster_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

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

// SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A/B format, majors, N>>3 at [17,23), M>>4 at [24,29)
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 lanes x 32 bit, 32 consecutive columns -> 32 registers per thread
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

// exp(a - b) as one FFMA + MUFU.EX2: ex2(a * log2e - b * log2e).  (__expf(a - b) costs FADD + FMUL + a denormal-range fix-up of three more
// instructions; with one warp per scheduler every instruction of these 64-element loops is ~5 cycles of latency.)
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
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

// one token row of a head: 64 contiguous 16-bit channels = 8 x 16 B
template <typename T> __device__ __forceinline__ void load_row(const T* p, bool valid, uint4 (&r)[8]) {
#pragma unroll
    for (int g = 0; g < 8; ++g) r[g] = valid ? ldg_cached(p + 8 * g) : make_uint4(0, 0, 0, 0);
}

template <typename T, int CS>
__global__ void __launch_bounds__(128) linattn_tc_kernel(const __grid_constant__ AttnArgs A) {
    extern __shared__ __align__(1024) unsigned char sm[];
    const uint32_t sbase = smem_addr(sm);
    float* s_max = reinterpret_cast<float*>(sm + kOffStat);
    float* s_sum = s_max + kD;
    const uint32_t bar_g1[2] = {sbase + kOffBar, sbase + kOffBar + 8};
    const uint32_t bar_g2 = sbase + kOffBar + 16;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + kOffBar + 32);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int prob = (int)blockIdx.x / CS, rank = CS > 1 ? (int)cluster_ctarank() : 0;  // 1-D clusters: rank == blockIdx.x % CS
    const int b = prob / A.heads, head = prob % A.heads;
    const int C = A.heads * kD, N = A.N;
    const T* qkv = reinterpret_cast<const T*>(A.qkv) + (int64_t)b * A.qb;
    const T* gq = qkv + (0 * C + head * kD);
    const T* gk = qkv + (1 * C + head * kD);
    const T* gv = qkv + (2 * C + head * kD);
    constexpr int kFmt = sizeof(T) == 2 && std::is_same<T, __nv_bfloat16>::value ? 1 : 0;

    if (tid == 0) {
        mbar_init(bar_g1[0], 1);
        mbar_init(bar_g1[1], 1);
        mbar_init(bar_g2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < kD) { s_max[tid] = -INFINITY; s_sum[tid] = 0.f; }
    if (warp == 0) {  // one warp owns the TMEM allocation (128 columns: two CTAs can share an SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(s_tmem)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t tmem_d1 = tmem, tmem_d2 = tmem + kD;

    const int n_chunks = (N + kChunk - 1) / kChunk;
    const int n_my = rank < n_chunks ? (n_chunks - rank + CS - 1) / CS : 0;  // this CTA's chunks: rank, rank + CS, ...
    const uint32_t idesc1 = umma_idesc(kFmt, 1, 1, 64, 64);    // both operands MN-major (channels contiguous)
    const uint32_t idesc2 = umma_idesc(kFmt, 0, 0, 128, 64);   // both operands K-major

    // ------------------------------------------------------------------ pass 1: ctx = softmax_d(K)^T V, q column statistics
    // The three token rows (K, V, Q: 24 x 16 B per thread) of chunk c+1 are requested while chunk c is being processed.
    uint4 rk[8], rv[8], rq[8];
    {
        const int n0 = rank * kChunk + tid;
        const bool v0 = n0 < N;
        load_row<T>(gk + (int64_t)n0 * A.qn, v0, rk);
        load_row<T>(gv + (int64_t)n0 * A.qn, v0, rv);
        load_row<T>(gq + (int64_t)n0 * A.qn, v0, rq);
    }
    for (int c = 0; c < n_my; ++c) {  // c counts this CTA's chunks; gc is the chunk's index in the sequence
        const int buf = c & 1, gc = rank + c * CS;
        unsigned char* sA = sm + (buf ? kOffA1 : kOffA0);
        unsigned char* sB = sm + (buf ? kOffB1 : kOffB0);
        const int n = gc * kChunk + tid;
        const bool valid = n < N;
        if (c >= 2) mbar_wait(bar_g1[buf], ((c >> 1) - 1) & 1);  // the MMAs that read this buffer two chunks ago are done
        float f[kD];
        // K: softmax over the 64 channels of this token, entirely in registers
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            float t8[8];
            unpack<T>(rk[g], t8);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[8 * g + e] = t8[e];
        }
        float m = f[0];
#pragma unroll
        for (int i = 1; i < kD; ++i) m = fmaxf(m, f[i]);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kD; ++i) { f[i] = ex2_approx(fmaf(f[i], kLog2e, -m * kLog2e)); s += f[i]; }
        const float inv = valid ? 1.f / s : 0.f;  // padded tokens contribute exact zeros
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            uint4 o;
            o.x = pack2<T>(f[8 * g] * inv, f[8 * g + 1] * inv); o.y = pack2<T>(f[8 * g + 2] * inv, f[8 * g + 3] * inv);
            o.z = pack2<T>(f[8 * g + 4] * inv, f[8 * g + 5] * inv); o.w = pack2<T>(f[8 * g + 6] * inv, f[8 * g + 7] * inv);
            *reinterpret_cast<uint4*>(sA + g * 2048 + tid * 16) = o;
        }
        // V: straight copy into the same layout; Q: raw tile for the per-channel max / sum over tokens (chunk index
        // XOR-swizzled by the row: conflict-free for the row writes here and the column reads below)
#pragma unroll
        for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4*>(sB + g * 2048 + tid * 16) = rv[g];
#pragma unroll
        for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4*>(sm + kOffQ + tid * 128 + ((g ^ (tid & 7)) << 4)) = rq[g];
        if (c + 1 < n_my) {  // prefetch the next chunk's rows; they land while the MMAs and the statistics run
            const int nn = n + CS * kChunk;
            const bool vn = nn < N;
            load_row<T>(gk + (int64_t)nn * A.qn, vn, rk);
            load_row<T>(gv + (int64_t)nn * A.qn, vn, rv);
            load_row<T>(gq + (int64_t)nn * A.qn, vn, rq);
        }
        proxy_fence();  // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const int nvalid = min(kChunk, N - gc * kChunk);
            const int ksteps = (nvalid + 15) >> 4;  // 16 tokens per MMA; the tail inside a step is zero-padded
            for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t da = umma_desc(smem_addr(sA) + ks * 256, 128, 2048);
                const uint64_t db = umma_desc(smem_addr(sB) + ks * 256, 128, 2048);
                umma(tmem_d1, da, db, idesc1, (c > 0 || ks > 0) ? 1u : 0u);
            }
            umma_commit(bar_g1[buf]);
        }
        {   // online max / sum of exp over this tile's tokens: thread = (channel, half of the tokens), halves merged through smem
            const int nvalid = min(kChunk, N - gc * kChunk);
            const T* col = reinterpret_cast<const T*>(sm + kOffQ);
            const int ch = tid & 63, hf = tid >> 6, g = ch >> 3, e = ch & 7;
            const int t0 = hf * 64, t1 = min(nvalid, t0 + 64);
            // fixed 64-iteration loops, unrolled and predicated, so that the independent shared-memory loads of a column pipeline
            // (measured neutral at N = 400: 45.4 us either way -- the kernel is bound by its barrier / MMA-completion chain, ncu source page)
            float mx = -INFINITY;
#pragma unroll 16
            for (int tt = 0; tt < 64; ++tt) {
                const int t = t0 + tt;
                const float v = to_f(col[t * 64 + ((g ^ (t & 7)) << 3) + e]);
                mx = t < t1 ? fmaxf(mx, v) : mx;
            }
            float* s_pm = reinterpret_cast<float*>(sm + kOffStat) + 2 * kD;  // [2][64] partial max, [2][64] partial sum
            s_pm[hf * 64 + ch] = mx;
            __syncthreads();
            const float mt = fmaxf(fmaxf(s_pm[ch], s_pm[64 + ch]), s_max[ch]);
            float acc = 0.f;
#pragma unroll 16
            for (int tt = 0; tt < 64; ++tt) {
                const int t = t0 + tt;
                const float v = ex2_approx(fmaf(to_f(col[t * 64 + ((g ^ (t & 7)) << 3) + e]), kLog2e, -mt * kLog2e));
                acc += t < t1 ? v : 0.f;
            }
            s_pm[128 + hf * 64 + ch] = acc;
            __syncthreads();
            if (tid < kD) {
                const float m_old = s_max[tid];
                s_sum[tid] = s_sum[tid] * (m_old == -INFINITY ? 0.f : __expf(m_old - mt)) + s_pm[128 + tid] + s_pm[192 + tid];
                s_max[tid] = mt;
            }
        }
        __syncthreads();  // the q tile and the partials are reused by the next chunk
    }
    // all of GEMM1 has landed in TMEM once the last commit fires (commits complete in order)
    if (n_my > 0) {
        const int last = n_my - 1;
        mbar_wait(bar_g1[last & 1], (last >> 1) & 1);
        tc_fence_after();
    }
    const float* s_cmax = s_max;  // per-channel max of q over ALL tokens, used by pass 2
    if constexpr (CS == 1) {
        // -------------------------------------------------------------- ctx' = diag(1/s) ctx -> shared (B operand of GEMM2, K-major)
        uint32_t r0[32], r1[32];
        const uint32_t taddr = tmem_d1 + ((uint32_t)(warp * 32) << 16);
        tmem_ld32(taddr, r0);        // columns j = 0..31
        tmem_ld32(taddr + 32, r1);   // columns j = 32..63
        if (lane < 16) {             // M = 64 accumulators live in lanes 0..15 of each 32-lane sub-partition: row i = 16*warp + lane
            const int i = warp * 16 + lane;
            const float inv = 1.f / s_sum[i];
            T* ctx = reinterpret_cast<T*>(sm + kOffCtx);
            // element (j, i) of the K-major B tile at (i/8)*1024 B + j*16 B + (i%8)*2 B
            T* dst = ctx + (i >> 3) * 512 + (i & 7);
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[j * 8] = from_f<T>(__uint_as_float(r0[j]) * inv);
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[(j + 32) * 8] = from_f<T>(__uint_as_float(r1[j]) * inv);
        }
    } else {
        // -------------------------------------------------------------- cluster: publish the partial ctx and the local (max, sum)
        float* s_ctxp = reinterpret_cast<float*>(sm + kOffCtxP);
        float* s_gmax = reinterpret_cast<float*>(sm + kOffGStat);
        float* s_ginv = s_gmax + kD;
        if (n_my > 0) {
            uint32_t r0[32], r1[32];
            const uint32_t taddr = tmem_d1 + ((uint32_t)(warp * 32) << 16);
            tmem_ld32(taddr, r0);
            tmem_ld32(taddr + 32, r1);
            if (lane < 16) {
                float* dst = s_ctxp + (warp * 16 + lane) * kD;
#pragma unroll
                for (int j = 0; j < 32; ++j) dst[j] = __uint_as_float(r0[j]);
#pragma unroll
                for (int j = 0; j < 32; ++j) dst[32 + j] = __uint_as_float(r1[j]);
            }
        } else {  // no tokens in this CTA (N <= 128 * rank): contributes zeros and (-inf, 0)
            for (int e = tid; e < kD * kD; e += 128) s_ctxp[e] = 0.f;
        }
        cluster_arrive();
        cluster_wait();  // every CTA's partials are visible cluster-wide
        const uint32_t a_ctxp = smem_addr(s_ctxp), a_max = smem_addr(s_max), a_sum = smem_addr(s_sum);
        if (tid < kD) {  // global statistics of channel tid: M = max_r m_r, S = sum_r s_r * exp(m_r - M)
            float mr[CS], sr[CS], M = -INFINITY;
#pragma unroll
            for (int r = 0; r < CS; ++r) {
                mr[r] = ld_cluster_f32(map_to_rank(a_max + tid * 4, r));
                sr[r] = ld_cluster_f32(map_to_rank(a_sum + tid * 4, r));
                M = fmaxf(M, mr[r]);
            }
            float S = 0.f;
#pragma unroll
            for (int r = 0; r < CS; ++r) S += mr[r] == -INFINITY ? 0.f : sr[r] * __expf(mr[r] - M);
            s_gmax[tid] = M;
            s_ginv[tid] = 1.f / S;
        }
        __syncthreads();
        {   // ctx' = diag(1/S) * sum_r ctx_r -> K-major B tile of GEMM2; consecutive threads read consecutive j (coalesced DSMEM rows)
            uint32_t base[CS];
#pragma unroll
            for (int r = 0; r < CS; ++r) base[r] = map_to_rank(a_ctxp, r);
            T* ctx = reinterpret_cast<T*>(sm + kOffCtx);
            for (int e = tid; e < kD * kD; e += 128) {
                const int i = e >> 6, j = e & 63;
                float v = 0.f;
#pragma unroll
                for (int r = 0; r < CS; ++r) v += ld_cluster_f32(base[r] + e * 4);
                ctx[(i >> 3) * 512 + j * 8 + (i & 7)] = from_f<T>(v * s_ginv[i]);
            }
        }
        cluster_arrive();  // second phase: this CTA no longer reads remote memory; matched by the wait before exit
        s_cmax = s_gmax;
    }
    float* s_cml = reinterpret_cast<float*>(sm + kOffStat) + 2 * kD;  // max * log2e per channel (reuses the pass-1 partials)
    if (tid < kD) s_cml[tid] = s_cmax[tid] * kLog2e;
    tc_fence_before();
    __syncthreads();

    // ------------------------------------------------------------------ pass 2: y = exp(q - max) ctx'
    T* gy = reinterpret_cast<T*>(A.y) + (int64_t)b * A.yb + head * kD;
    load_row<T>(gq + (int64_t)(rank * kChunk + tid) * A.qn, rank * kChunk + tid < N, rq);
    for (int c = 0; c < n_my; ++c) {
        unsigned char* sP = sm + ((c & 1) ? kOffA1 : kOffA0);
        const int n = (rank + c * CS) * kChunk + tid;
        const bool valid = n < N;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            float t8[8];
            unpack<T>(rq[g], t8);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float p = ex2_approx(fmaf(t8[e], kLog2e, -s_cml[8 * g + e]));  // branch-free; rows past N are zeroed by the select
                t8[e] = valid ? p : 0.f;
            }
            uint4 o;
            o.x = pack2<T>(t8[0], t8[1]); o.y = pack2<T>(t8[2], t8[3]); o.z = pack2<T>(t8[4], t8[5]); o.w = pack2<T>(t8[6], t8[7]);
            *reinterpret_cast<uint4*>(sP + g * 2048 + tid * 16) = o;  // row = token, 16 B chunk g of the K dim
        }
        if (c + 1 < n_my) load_row<T>(gq + (int64_t)(n + CS * kChunk) * A.qn, n + CS * kChunk < N, rq);  // next chunk's q rows (L2 hits)
        proxy_fence();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {  // K = 64 channels = 4 steps of 16 = 2 chunks of 8 each
                const uint64_t da = umma_desc(smem_addr(sP) + ks * 4096, 2048, 128);
                const uint64_t db = umma_desc(sbase + kOffCtx + ks * 2048, 1024, 128);
                umma(tmem_d2, da, db, idesc2, ks > 0 ? 1u : 0u);
            }
            umma_commit(bar_g2);
        }
        mbar_wait(bar_g2, c & 1);
        tc_fence_after();
        {   // epilogue: lane <-> token (row 32*warp + lane of the 128-row accumulator), 64 fp32 columns -> 128 B of bf16
            const uint32_t taddr = tmem_d2 + ((uint32_t)(warp * 32) << 16);
            uint32_t v[32];
            T* dst = gy + (int64_t)n * A.yn;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                tmem_ld32(taddr + 32 * h, v);
                if (valid) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 o;
                        o.x = pack2<T>(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1]));
                        o.y = pack2<T>(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3]));
                        o.z = pack2<T>(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5]));
                        o.w = pack2<T>(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7]));
                        *reinterpret_cast<uint4*>(dst + 32 * h + 8 * g) = o;
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();  // D2 has been drained: the next chunk's MMA may overwrite it
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
    if constexpr (CS > 1) cluster_wait();  // no CTA of the cluster exits while a peer may still read its shared memory
}

}  // namespace tc

bool linattn_tc_supported(const AttnArgs& A, int dtype) {
    if (dtype != EL_BF16 && dtype != EL_F16) return false;
    if (A.qc != 1 || A.yc != 1) return false;  // channel-contiguous (NHWC) views only
    const int C = A.heads * tc::kD;
    (void)C;
    return aligned16(A.qkv) && aligned16(A.y) && A.qn % 8 == 0 && A.qb % 8 == 0 && A.yn % 8 == 0 && A.yb % 8 == 0;
}

template <typename T, int CS>
static void launch_one(const AttnArgs& A, int B, cudaStream_t s) {
    const uint32_t smem = CS > 1 ? tc::kSmemBytesCluster : tc::kSmemBytes;
    cudaFuncSetAttribute(tc::linattn_tc_kernel<T, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * A.heads * CS));
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CS > 1 ? 1 : 0;
    cudaLaunchKernelEx(&cfg, tc::linattn_tc_kernel<T, CS>, A);
}

int linattn_tc_launch(const AttnArgs& A, int B, int dtype, cudaStream_t s) {
    // One CTA per (image, head).  EL_LINATTN_CLUSTER=1 selects the 4-CTA cluster variant that splits the tokens: built to fill the GPU when
    // B * heads < 148 (configs[1]: 128 problems of 4 chunks), parity-tested, but MEASURED SLOWER there (54.1 vs 45.4 us): four times the
    // CTAs each pay the fixed costs (TMEM allocation, barrier init, two cluster barriers, 64 KB of DSMEM reads for the ctx sum) that
    // already dominate this latency-bound kernel.
    static const int mode = [] { const char* v = getenv("EL_LINATTN_CLUSTER"); return v ? atoi(v) : 0; }();
    const bool cluster = mode != 0;
    if (dtype == EL_BF16) {
        if (cluster) launch_one<__nv_bfloat16, 4>(A, B, s);
        else launch_one<__nv_bfloat16, 1>(A, B, s);
    } else {
        if (cluster) launch_one<__half, 4>(A, B, s);
        else launch_one<__half, 1>(A, B, s);
    }
    note_launches(1);
    return check_launch();
}

}  // namespace el
