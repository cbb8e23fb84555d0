// Depthwise k x k convolution (stride 1, "same" padding) over channel-contiguous (NHWC) views, with an optional fused
// bias + activation epilogue.  Engine-path replacement for the depthwise halves of the reference's separable blocks:
//   DSConv.dw  (nn/modules/conv.py:87-104: Conv2d(c, c, k, groups=c, bias=False), k = 3 and 7 in DSBottleneck,
//               nn/modules/block.py:1467-1503)           -> no epilogue, the result feeds the pointwise conv
//   DWConv     (nn/modules/conv.py:124-130, the Detect cls tower head.py:66-71) -> folded-BN bias + SiLU epilogue
// HBM-bound by design for k = 3 (every input element read from DRAM once, every output written once); k = 7 is bounded by
// fp32 FMA issue (49 taps per output) and is written to keep the load/store pipe below it.
//
// A CTA owns a tile of TH output rows x TW output columns x up to 64 channels.  The input patch (tile + halo) is staged once in
// shared memory with 16-byte cp.async in the global (row, column, channel-vector) order -- coalesced reads, linear writes,
// zero fill outside the image.  A thread <-> (channel vector cv, output column x, strip of 4 rows): lanes run over (cv, x), so
// every LDS.128 / STG.128 of a warp is one contiguous 512-byte span.  The register tile is VERTICAL: for each filter column kx
// the thread loads the 4 + k - 1 input vectors of its column once, and every filter tap (one 16-byte broadcast per 4 channels)
// feeds 4 output rows x V FMAs.
#include <cstdlib>

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
    int CVL, cvl_shift;   // channel vectors per CTA (<= 8); cvl_shift = log2(CVL), or -1 when CVL is not a power of two (C = 48, 80, ...)
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
    // (index / CVL, index % CVL): shifts for the usual power-of-two vector counts, a division for e.g. the 80-channel class towers
    auto div_cvl = [&](int i) { return G.cvl_shift >= 0 ? i >> G.cvl_shift : i / CVL; };
    auto mod_cvl = [&](int i) { return G.cvl_shift >= 0 ? i & (CVL - 1) : i % CVL; };
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
        const int per_row = PWf * CVL;
        for (int r = 0; r < rows_in; ++r) {
            const int iy = ty0 - P + r;
            const bool row_ok = iy >= 0 && iy < H;
            const T* xrow = xb + (int64_t)iy * xs.h;
            const uint32_t drow = dw_smem_addr(s_in + (size_t)r * per_row);
            for (int i = tid; i < per_row; i += nthreads) {
                const int cv = mod_cvl(i), ix = tx0 - P + div_cvl(i);
                const bool ok = row_ok && ix >= 0 && ix < W;
                const T* src = ok ? xrow + (int64_t)ix * xs.w + cv * V : xb;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(drow + (uint32_t)i * 16), "l"(src), "r"(ok ? 16u : 0u) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int cvl = mod_cvl(tid), rest = div_cvl(tid);
    const int xl = rest % G.TW, strip0 = rest / G.TW;
    const int ox = tx0 + xl;
    if (strip0 >= G.NS || ox >= W) return;
    const int ch = c0 + cvl * V;
    float bv[V];
#pragma unroll
    for (int e = 0; e < V; ++e) bv[e] = bias ? __ldg(bias + ch + e) : 0.f;
    const float4* sw4 = reinterpret_cast<const float4*>(s_w);
    const int row_stride = PWf * CVL;  // uint4 units
    for (int ly0 = strip0 * kDwRT; ly0 < TH; ly0 += G.NS * kDwRT) {
        // accumulators, inputs and taps are kept as fp32 pairs of adjacent channels: one FFMA2 per pair
        f32x2 acc2[kDwRT][V / 2];
#pragma unroll
        for (int r = 0; r < kDwRT; ++r)
#pragma unroll
            for (int e = 0; e < V / 2; ++e) acc2[r][e] = pack_f32x2(bv[2 * e], bv[2 * e + 1]);
#pragma unroll 1
        for (int kx = 0; kx < K; ++kx) {
            const uint4* scol = s_in + (size_t)ly0 * row_stride + (xl + kx) * CVL + cvl;
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

// =====================================================================================================================
// k x k on the tensor cores (16-bit activations; written for k = 7, also used for 3 and 5).  The CUDA-core kernel above is bounded by FMA issue (49 taps per output:
// 0.15 - 0.4 of the HBM roofline, ncu: fma pipe 43 %, issue 72 %).  A depthwise row filter is a banded (Toeplitz) matrix:
//     out_c[y][x] = sum_dy  sum_x'  in_c[y + dy][x'] * T_dy[x'][x],     T_dy[x'][x] = w_c[dy][x' - x]  (0 <= x' - x < 7)
// so one channel's 16 (rows) x 8 (columns) output block is 7 mma.sync m16n8k16 (one per filter row dy) with the data as the
// A operand (16 rows x 16 patch columns, ldmatrix from a per-channel plane in shared memory) and the Toeplitz block as B
// (16 x 8, the same for every block of the channel: 14 registers per channel, built once per CTA).  44 % of the MMA's
// multiplies are useful, which is still ~6x the CUDA-core FFMA2 rate, and the kernel becomes bounded by shared-memory
// wavefronts instead (ldmatrix: 6 per (row block, dy) thanks to the column block the two 8-column outputs share).
//   CTA = (image, 16 channels = one 32-byte sector per pixel, 16 output columns); it walks down the image in 32-row tiles.
//   stage : patch (38 rows x 22 columns x 16 channels) -> 16 channel planes [row][24] (2-byte scatter; zero fill = padding)
//   MMA   : warp w owns channels 2w, 2w+1; per channel and 16-row block: 7 x (ldmatrix.x4 + ldmatrix.x2 + 2 MMA)
//   out   : bias + activation on the accumulators -> channel planes of the output tile -> 16-byte NHWC vectors -> global
// Measured and not kept: (a) pixel pairs per thread with 32-bit plane stores / loads (128 registers: 66 vs 56 us at 160^2); (b) splitting
// the rows of small maps over more CTAs (41 vs 36.5 us at 80^2); (c) a second generation with cp.async into a raw NHWC buffer and
// ldmatrix.trans for both transpositions (no prefetch registers, a third barrier per tile: 60.8 vs 57.0 us at 160^2, 50 vs 42 us for k = 3).
// Of the 56 us at 160^2 about 25 us scale with the filter rows (ldmatrix + MMA); the rest is staging, write-out and barriers at 16 warps / SM.
namespace dwtc {

constexpr int TX = 16, TY = 32, CH = 16;
constexpr int PC = 24;            // patch row pitch (elements): 48 B -> the 8 row addresses of an ldmatrix hit 8 distinct 16 B slots
constexpr int OC = 24;            // output tile row pitch: 12 words -> conflict-free accumulator stores
constexpr int PLANE_OUT = TY * OC;
constexpr int kThreads = 256;
template <int KS> struct Geo {    // filter size 3 / 5 / 7 (the same kernel: KS Toeplitz blocks per channel, KS MMAs per output block)
    static constexpr int PAD = KS / 2;
    static constexpr int PR = TY + KS - 1;   // patch rows (38 for k = 7)
    static constexpr int PCU = TX + KS - 1;  // patch columns in use (22 for k = 7); columns PCU..23 stay zero
    static constexpr int PLANE_IN = PR * PC;
    static constexpr size_t kSmem = (size_t)CH * PLANE_IN * 2 + (size_t)CH * PLANE_OUT * 2 + (size_t)CH * KS * KS * 4;
};

template <typename T> __device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1);
template <> __device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <> __device__ __forceinline__ void mma16816<__half>(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <typename T> __device__ __forceinline__ uint32_t pack2h(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2h<__nv_bfloat16>(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
template <> __device__ __forceinline__ uint32_t pack2h<__half>(float a, float b) {
    __half2 t = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

template <typename T, int KS>
__global__ void __launch_bounds__(kThreads, 2) dwconv_tc_kernel(const T* __restrict__ x, Strides4 xs, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, T* __restrict__ o, Strides4 os, int C, int H, int W, int act,
                                                                 int n_cb, int rows_per_cta) {
    constexpr int PAD = Geo<KS>::PAD, PR = Geo<KS>::PR, PCU = Geo<KS>::PCU, PLANE_IN = Geo<KS>::PLANE_IN;
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint16_t* s_in = reinterpret_cast<uint16_t*>(s_raw);                                    // [16 channels][38 rows][24]
    uint16_t* s_out = s_in + CH * PLANE_IN;                                                  // [16 channels][32 rows][24]
    float* s_wt = reinterpret_cast<float*>(s_raw + (size_t)CH * (PLANE_IN + PLANE_OUT) * 2);  // [16 channels][49 taps]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int tx0 = (int)blockIdx.x * TX, c0 = ((int)blockIdx.y % n_cb) * CH;  // blockIdx.y = row chunk * n_cb + channel block
    const int64_t n = blockIdx.z;
    pdl_launch_dependents();
    // constants: filter taps of the CTA's channels; zero the pitch columns 22, 23 once (ldmatrix reads them against zero Toeplitz rows: they must be finite; staging rewrites columns 0..21 of every row for every tile)
    for (int i = tid; i < CH * KS * KS; i += kThreads) {
        const int c = i / (KS * KS), tap = i - c * (KS * KS);
        s_wt[i] = __ldg(w + (int64_t)tap * C + c0 + c);
    }
    for (int i = tid; i < CH * PR * ((PC - PCU) / 2); i += kThreads) {  // (plane, row): columns PCU .. 23 (22, 23 for k = 7), one word per two columns
        const int row = i / ((PC - PCU) / 2), wd = i - row * ((PC - PCU) / 2);
        *reinterpret_cast<uint32_t*>(s_in + row * PC + PCU + 2 * wd) = 0u;
    }
    __syncthreads();
    // Toeplitz B fragments of this warp's two channels: element (k, n) = w[dy][k - n]; b0 = rows k = 2t, 2t+1, b1 = rows 2t+8, 2t+9, column n = g
    uint32_t bf[2][KS][2];
    float bv[2];
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        const float* wc = s_wt + (2 * warp + cc) * (KS * KS);
        bv[cc] = bias ? __ldg(bias + c0 + 2 * warp + cc) : 0.f;
#pragma unroll
        for (int dy = 0; dy < KS; ++dy) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k0 = 2 * t + 8 * h - g, k1 = k0 + 1;  // tap index dx = k - n
                const float w0 = (k0 >= 0 && k0 < KS) ? wc[dy * KS + k0] : 0.f;
                const float w1 = (k1 >= 0 && k1 < KS) ? wc[dy * KS + k1] : 0.f;
                bf[cc][dy][h] = pack2h<T>(w0, w1);
            }
        }
    }
    pdl_wait();
    const T* xb = x + n * xs.n + c0;
    T* ob = o + n * os.n + c0;
    const uint32_t s_in_addr = dw_smem_addr(s_in);
    // thread <-> patch pixel (both 16-byte vectors of its 16 channels).  The global loads of tile i+1 are issued before the MMA phase
    // of tile i (ncu: the un-prefetched version spent 3.6 of ~13 stall cycles per issue on long_scoreboard at 16 warps per SM).
    // (Pixel PAIRS per thread with 32-bit plane stores / loads halve the scatter wavefronts but need 128 registers: measured slower.)
    constexpr int NPX = (PR * PCU + kThreads - 1) / kThreads;
    uint4 pv[NPX][2];
    auto fetch = [&](int ty0) {
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
            const int p = tid + j * kThreads, r = p / PCU, c = p - r * PCU;
            const int iy = ty0 - PAD + r, ix = tx0 - PAD + c;
            pv[j][0] = make_uint4(0, 0, 0, 0);
            pv[j][1] = pv[j][0];
            if (p < PR * PCU && iy >= 0 && iy < H && ix >= 0 && ix < W) {
                const T* src = xb + (int64_t)iy * xs.h + (int64_t)ix * xs.w;
                pv[j][0] = ldg_l2(src);
                pv[j][1] = ldg_l2(src + 8);
            }
        }
    };
    const int ty_begin = (int)blockIdx.y / n_cb * rows_per_cta, ty_end = min(H, ty_begin + rows_per_cta);
    fetch(ty_begin);
    for (int ty0 = ty_begin; ty0 < ty_end; ty0 += TY) {
        // ---- stage: 2-byte scatter of the prefetched pixels into the channel planes
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
            const int p = tid + j * kThreads;
            if (p < PR * PCU) {
                const int r = p / PCU, c = p - r * PCU;
                uint16_t* d = s_in + r * PC + c;
                const uint32_t q[8] = {pv[j][0].x, pv[j][0].y, pv[j][0].z, pv[j][0].w, pv[j][1].x, pv[j][1].y, pv[j][1].z, pv[j][1].w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    d[(2 * e) * PLANE_IN] = (uint16_t)(q[e] & 0xffffu);
                    d[(2 * e + 1) * PLANE_IN] = (uint16_t)(q[e] >> 16);
                }
            }
        }
        __syncthreads();
        if (ty0 + TY < ty_end) fetch(ty0 + TY);
        // ---- MMA: per channel and 16-row block, 7 filter rows x 2 column blocks
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int ch = 2 * warp + cc;
            const uint32_t plane = s_in_addr + (uint32_t)ch * PLANE_IN * 2;
#pragma unroll
            for (int mt = 0; mt < TY / 16; ++mt) {
                if (ty0 + 16 * mt >= H) break;  // warp-uniform
                float acc[2][4];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[nt][e] = bv[cc];
                // ldmatrix row addresses: lane l supplies row (l & 15) of the 16-row block; x4: column block 0 / 1 by (l >> 4); x2: column block 2
                const uint32_t a_lo = plane + (uint32_t)((16 * mt + (lane & 15)) * PC + 8 * (lane >> 4)) * 2;
                const uint32_t a_hi = plane + (uint32_t)((16 * mt + (lane & 15)) * PC + 16) * 2;
#pragma unroll
                for (int dy = 0; dy < KS; ++dy) {
                    uint32_t m00, m10, m01, m11, m02, m12;
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(m00), "=r"(m10), "=r"(m01), "=r"(m11) : "r"(a_lo + dy * PC * 2));
                    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(m02), "=r"(m12) : "r"(a_hi + dy * PC * 2));
                    mma16816<T>(acc[0], m00, m10, m01, m11, bf[cc][dy][0], bf[cc][dy][1]);
                    mma16816<T>(acc[1], m01, m11, m02, m12, bf[cc][dy][0], bf[cc][dy][1]);
                }
                uint16_t* po = s_out + ch * PLANE_OUT + (16 * mt + g) * OC + 2 * t;
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    float f[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) f[e] = act == 1 ? dw_silu<T>(acc[nt][e]) : (act == 2 ? fmaxf(acc[nt][e], 0.f) : acc[nt][e]);
                    *reinterpret_cast<uint32_t*>(po + 8 * nt) = pack2h<T>(f[0], f[1]);            // row g,     columns 8 nt + 2t, +1
                    *reinterpret_cast<uint32_t*>(po + 8 * OC + 8 * nt) = pack2h<T>(f[2], f[3]);   // row g + 8
                }
            }
        }
        __syncthreads();
        // ---- write out: thread <-> output pixel, 16 channels gathered from the planes into two 16-byte NHWC vectors
#pragma unroll
        for (int j = 0; j < TX * TY / kThreads; ++j) {
            const int p = tid + j * kThreads, y = p >> 4, xx = p & 15;
            if (ty0 + y < H && tx0 + xx < W) {
                const uint16_t* ps = s_out + y * OC + xx;
                uint32_t q[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) q[e] = (uint32_t)ps[(2 * e) * PLANE_OUT] | ((uint32_t)ps[(2 * e + 1) * PLANE_OUT] << 16);
                T* dst = ob + (int64_t)(ty0 + y) * os.h + (int64_t)(tx0 + xx) * os.w;
                *reinterpret_cast<uint4*>(dst) = make_uint4(q[0], q[1], q[2], q[3]);
                *reinterpret_cast<uint4*>(dst + 8) = make_uint4(q[4], q[5], q[6], q[7]);
            }
        }
        // the next tile's staging writes only s_in (all MMAs of this tile are behind the barrier above); its epilogue writes s_out
        // after the barrier that follows the staging, i.e. after every thread has finished the write-out above
    }
}

template <typename T, int KS>
static int launch(const void* x, Strides4 xs, const float* w, const float* bias, void* out, Strides4 os, int B, int C, int H, int W, int act,
                  cudaStream_t st) {
    // a CTA walks down its rows in 32-row tiles (Toeplitz fragments built once, next tile prefetched).  Splitting the rows of small maps
    // over more CTAs was measured slower (80^2: 41 vs 36.5 us, 40^2: 32.7 vs 28.4 us -- the per-CTA prologue outweighs the fuller waves);
    // EL_DW_SPLIT=1 re-enables it for experiments.
    const int n_cb = C / CH, row_tiles = (int)ceil_div(H, TY);
    const int64_t base = ceil_div(W, TX) * n_cb * B;
    static const int split_mode = [] { const char* v = getenv("EL_DW_SPLIT"); return v ? atoi(v) : 0; }();
    int chunks = split_mode ? (int)ceil_div(4 * 2 * kSMs, base) : 1;
    if (chunks > row_tiles) chunks = row_tiles;
    if (chunks < 1) chunks = 1;
    const int rows_per_cta = (int)ceil_div(row_tiles, chunks) * TY;
    chunks = (int)ceil_div(H, rows_per_cta);
    if ((int64_t)n_cb * chunks > 65535) return EL_ERR_UNSUPPORTED;
    dim3 grid((unsigned)ceil_div(W, TX), (unsigned)(n_cb * chunks), (unsigned)B);
    cudaError_t e = cudaFuncSetAttribute(dwconv_tc_kernel<T, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<KS>::kSmem);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return EL_ERR_CUDA; }
    e = launch_pdl(dwconv_tc_kernel<T, KS>, grid, dim3(kThreads), Geo<KS>::kSmem, st, (const T*)x, xs, w, bias, (T*)out, os, C, H, W, act, n_cb, rows_per_cta);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return EL_ERR_CUDA; }
    return EL_OK;
}

}  // namespace dwtc

template <typename T, int K>
static int launch_dw(const void* x, Strides4 xs, const float* w, const float* bias, void* out, Strides4 os, int B, int C, int H, int W, int act,
                     cudaStream_t st) {
    constexpr int V = Vec16<T>::N;
    if (C % V) return EL_ERR_UNSUPPORTED;
    const int CV = C / V;
    DwGeom G;
    G.CVL = CV < 8 ? CV : 8;
    while (CV % G.CVL) --G.CVL;  // largest divisor of CV that is <= 8 (80 channels: 5 vectors per CTA)
    G.cvl_shift = G.CVL == 1 ? 0 : (G.CVL == 2 ? 1 : (G.CVL == 4 ? 2 : (G.CVL == 8 ? 3 : -1)));
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

// persistent TMA-pipelined k = 3 kernel, dwconv_tma.cu
int dwconv3_tma_launch(const void* x, Strides4 xs, const float* w, const float* bias, void* out, Strides4 os, int B, int C, int H, int W, int act,
                       int dtype, cudaStream_t st);

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
        if constexpr (sizeof(T) == 2) {
            // k = 7, 16-bit maps of at least 32 x 32 with C % 16 == 0: Toeplitz MMA kernel (smaller maps do not fill its 32 x 16 tiles; for
            // k = 3 it was measured equal or slower than the CUDA-core kernel: 41.7 vs 41.3 us at 160^2, 60 vs 44 us at 80^2 -- its staging
            // and write-out cost what the 9 taps cost).  EL_DW_TC = 0: never, 2: every map size, 3: every map size and k = 3 / 5 too
            static const int tc_mode = [] { const char* v = getenv("EL_DW_TC"); return v ? atoi(v) : 1; }();
            const bool k_ok = k == 7 || tc_mode == 3;
            if (k_ok && C % dwtc::CH == 0 && B <= 65535 && C / dwtc::CH <= 65535 && tc_mode && (tc_mode >= 2 || (H >= 32 && W >= 32))) {
                rc = k == 7 ? dwtc::launch<T, 7>(x, xs, w, bias, out, os, B, C, H, W, act, st)
                   : k == 5 ? dwtc::launch<T, 5>(x, xs, w, bias, out, os, B, C, H, W, act, st)
                            : dwtc::launch<T, 3>(x, xs, w, bias, out, os, B, C, H, W, act, st);
                if (rc != EL_OK) return rc;
                note_launches(1);
                return check_launch();
            }
        }
        if constexpr (sizeof(T) == 2) {
            // k = 3, C = 16 / 32 / 64 n: persistent TMA-pipelined streaming kernel (dwconv_tma.cu); EL_DW_TMA=0 keeps the cp.async tile kernel
            static const bool tma_on = [] { const char* v = getenv("EL_DW_TMA"); return !(v && v[0] == '0'); }();
            if (k == 3 && tma_on) {
                rc = dwconv3_tma_launch(x, xs, w, bias, out, os, B, C, H, W, act, dtype, st);
                if (rc == EL_OK) { note_launches(1); return check_launch(); }
                if (rc != EL_ERR_UNSUPPORTED) return rc;
            }
        }
        if (k == 3) rc = launch_dw<T, 3>(x, xs, w, bias, out, os, B, C, H, W, act, st);
        else if (k == 5) rc = launch_dw<T, 5>(x, xs, w, bias, out, os, B, C, H, W, act, st);
        else rc = launch_dw<T, 7>(x, xs, w, bias, out, os, B, C, H, W, act, st);
    });
    if (rc != EL_OK) return rc;
    note_launches(1);
    return check_launch();
}
