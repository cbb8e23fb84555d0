// Haar DWT split / adjoint, sub-band merge (bilinear up * w + concat) and gated residual for the
// DSC3K2_Wavelet neck.  Replaces _PywtDWT2D.forward (nn/modules/block.py:3619-3642) and the
// elementwise tail of _WaveletEnhancer.forward (block.py:3696-3710) of the reference.
//
// All kernels are HBM-bound streaming kernels: the fast path walks channel-contiguous (NHWC or
// NHWC channel-slice) views with 16-byte vectors, one vector per thread per tap; a generic
// strided kernel (thread order follows the fastest-moving stride) covers NCHW and odd shapes.
#include <cstdlib>

#include "el_common.cuh"

namespace el {

// float32(2^-1/2)^2 -- the value of every tap _PywtDWT2D builds (block.py:3597-3609), NOT 0.5.
__device__ __constant__ float kHaar = 0.49999997f;

// ------------------------------------------------------------------------------- DWT forward
template <typename T>
__global__ void __launch_bounds__(256) dwt_fwd_cvec(const T* __restrict__ x, Strides4 xs, T* __restrict__ o, int64_t ob, Strides4 os,
                                                    int CV, int H2, int W2, int64_t total) {
    constexpr int V = Vec16<T>::N;
    const float k = kHaar;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int cv = (int)(idx % CV);
        int64_t t = idx / CV;
        int j = (int)(t % W2); t /= W2;
        int i = (int)(t % H2);
        int64_t n = t / H2;
        const T* p = x + n * xs.n + (int64_t)(2 * i) * xs.h + (int64_t)(2 * j) * xs.w + cv * V;
        float a[V], b[V], c[V], d[V];
        uint4 ra = ldg_stream(p), rb = ldg_stream(p + xs.w), rc = ldg_stream(p + xs.h), rd = ldg_stream(p + xs.h + xs.w);
        unpack<T>(ra, a); unpack<T>(rb, b); unpack<T>(rc, c); unpack<T>(rd, d);
        float ll[V], lh[V], hl[V], hh[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            float s0 = a[e] + b[e], s1 = c[e] + d[e], d0 = a[e] - b[e], d1 = c[e] - d[e];
            ll[e] = k * (s0 + s1);
            lh[e] = k * (d0 + d1);
            hl[e] = k * (s0 - s1);
            hh[e] = k * (d0 - d1);
        }
        T* q = o + n * os.n + (int64_t)i * os.h + (int64_t)j * os.w + cv * V;
        *reinterpret_cast<uint4*>(q) = pack<T>(ll);
        *reinterpret_cast<uint4*>(q + ob) = pack<T>(lh);
        *reinterpret_cast<uint4*>(q + 2 * ob) = pack<T>(hl);
        *reinterpret_cast<uint4*>(q + 3 * ob) = pack<T>(hh);
    }
}

// decompose a flat index over (n, c, i, j) with either c or j moving fastest
template <bool CH_FAST>
__device__ __forceinline__ void split_ncij(int64_t idx, int C, int I, int J, int64_t& n, int& c, int& i, int& j) {
    if (CH_FAST) {
        c = (int)(idx % C); idx /= C;
        j = (int)(idx % J); idx /= J;
        i = (int)(idx % I); n = idx / I;
    } else {
        j = (int)(idx % J); idx /= J;
        i = (int)(idx % I); idx /= I;
        c = (int)(idx % C); n = idx / C;
    }
}

template <typename T, bool CH_FAST>
__global__ void __launch_bounds__(256) dwt_fwd_generic(const T* __restrict__ x, Strides4 xs, T* __restrict__ o, int64_t ob, Strides4 os,
                                                       int C, int H2, int W2, int64_t total) {
    const float k = kHaar;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t n; int c, i, j;
        split_ncij<CH_FAST>(idx, C, H2, W2, n, c, i, j);
        const T* p = x + n * xs.n + (int64_t)c * xs.c + (int64_t)(2 * i) * xs.h + (int64_t)(2 * j) * xs.w;
        float a = to_f(p[0]), b = to_f(p[xs.w]), cc = to_f(p[xs.h]), d = to_f(p[xs.h + xs.w]);
        float s0 = a + b, s1 = cc + d, d0 = a - b, d1 = cc - d;
        T* q = o + n * os.n + (int64_t)c * os.c + (int64_t)i * os.h + (int64_t)j * os.w;
        q[0] = from_f<T>(k * (s0 + s1));
        q[ob] = from_f<T>(k * (d0 + d1));
        q[2 * ob] = from_f<T>(k * (s0 - s1));
        q[3 * ob] = from_f<T>(k * (d0 - d1));
    }
}

// ------------------------------------------------------------------------------- DWT adjoint
// One thread per 2x2 output quad (ceil sizes so odd trailing rows/cols get their zeros).
template <typename T, bool CH_FAST>
__global__ void __launch_bounds__(256) dwt_bwd_generic(const T* __restrict__ g, int64_t gb, Strides4 gs, T* __restrict__ o, Strides4 os,
                                                       int C, int H, int W, int64_t total) {
    const float k = kHaar;
    const int H2 = H / 2, W2 = W / 2, HQ = (H + 1) / 2, WQ = (W + 1) / 2;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t n; int c, i, j;
        split_ncij<CH_FAST>(idx, C, HQ, WQ, n, c, i, j);
        T* q = o + n * os.n + (int64_t)c * os.c + (int64_t)(2 * i) * os.h + (int64_t)(2 * j) * os.w;
        if (i < H2 && j < W2) {
            const T* p = g + n * gs.n + (int64_t)c * gs.c + (int64_t)i * gs.h + (int64_t)j * gs.w;
            float ll = to_f(p[0]), lh = to_f(p[gb]), hl = to_f(p[2 * gb]), hh = to_f(p[3 * gb]);
            float s0 = ll + lh, s1 = hl + hh, d0 = ll - lh, d1 = hl - hh;
            q[0] = from_f<T>(k * (s0 + s1));
            q[os.w] = from_f<T>(k * (d0 + d1));
            q[os.h] = from_f<T>(k * (s0 - s1));
            q[os.h + os.w] = from_f<T>(k * (d0 - d1));
        } else {  // pixels the floor()ed analysis never read
            const T z = from_f<T>(0.f);
            q[0] = z;
            if (2 * j + 1 < W) q[os.w] = z;
            if (2 * i + 1 < H) {
                q[os.h] = z;
                if (2 * j + 1 < W) q[os.h + os.w] = z;
            }
        }
    }
}

// ----------------------------------------------------------------------------- merge forward
struct BandPtrs {
    const void* p[4];
    Strides4 s[4];
};
struct BandPtrsMut {
    void* p[4];
    Strides4 s[4];
};

__device__ __forceinline__ void band_weights(const float* __restrict__ alpha, float (&w)[4]) {
    // softplus(alpha) / (sum + 1e-6), block.py:3696-3697 (torch softplus: x > 20 -> x)
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float a = __ldg(alpha + i);
        w[i] = a > 20.f ? a : log1pf(expf(a));
        s += w[i];
    }
    s += 1e-6f;
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = w[i] / s;
}

// PyTorch upsample_bilinear2d source rule, align_corners=False
__device__ __forceinline__ void bilin_src(int dst, float scale, int n_in, int& i0, int& i1, float& lam) {
    float src = scale * (dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = (int)src;
    if (i0 > n_in - 1) i0 = n_in - 1;
    i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    lam = src - (float)i0;
}

template <typename T>
__global__ void __launch_bounds__(256) merge_fwd_cvec(const T* __restrict__ b, Strides4 bs, BandPtrs bands, const float* __restrict__ alpha,
                                                      T* __restrict__ out, Strides4 os, int c, int H, int W, int h, int w, int64_t total) {
    constexpr int V = Vec16<T>::N;
    __shared__ float sw[4];
    if (threadIdx.x == 0) {
        float wt[4];
        band_weights(alpha, wt);
        sw[0] = wt[0]; sw[1] = wt[1]; sw[2] = wt[2]; sw[3] = wt[3];
    }
    __syncthreads();
    const int CV = 3 * c / V, half = c / 2;
    const float sh = (float)h / (float)H, swd = (float)w / (float)W;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int cv = (int)(idx % CV);
        int64_t t = idx / CV;
        int x = (int)(t % W); t /= W;
        int y = (int)(t % H);
        int64_t n = t / H;
        int ch = cv * V;
        T* q = out + n * os.n + (int64_t)y * os.h + (int64_t)x * os.w + ch;
        if (ch < c) {  // pass-through copy of b
            stg_stream(q, ldg_stream(b + n * bs.n + (int64_t)y * bs.h + (int64_t)x * bs.w + ch));
            continue;
        }
        int seg = (ch - c) / half, cc = (ch - c) - seg * half;
        int y0, y1, x0, x1; float ly, lx;
        bilin_src(y, sh, h, y0, y1, ly);
        bilin_src(x, swd, w, x0, x1, lx);
        const Strides4 s = bands.s[seg];
        const T* p = reinterpret_cast<const T*>(bands.p[seg]) + n * s.n + cc;
        float v00[V], v01[V], v10[V], v11[V], r[V];
        // band pixels are re-read by ~4 output pixels each: keep them in L1
        unpack<T>(ldg_cached(p + (int64_t)y0 * s.h + (int64_t)x0 * s.w), v00);
        unpack<T>(ldg_cached(p + (int64_t)y0 * s.h + (int64_t)x1 * s.w), v01);
        unpack<T>(ldg_cached(p + (int64_t)y1 * s.h + (int64_t)x0 * s.w), v10);
        unpack<T>(ldg_cached(p + (int64_t)y1 * s.h + (int64_t)x1 * s.w), v11);
        const float wy0 = 1.f - ly, wx0 = 1.f - lx, wb = sw[seg];
#pragma unroll
        for (int e = 0; e < V; ++e) r[e] = (wy0 * (wx0 * v00[e] + lx * v01[e]) + ly * (wx0 * v10[e] + lx * v11[e])) * wb;
        stg_stream(q, pack<T>(r));
    }
}

template <typename T, bool CH_FAST>
__global__ void __launch_bounds__(256) merge_fwd_generic(const T* __restrict__ b, Strides4 bs, BandPtrs bands, const float* __restrict__ alpha,
                                                         T* __restrict__ out, Strides4 os, int c, int H, int W, int h, int w, int64_t total) {
    __shared__ float sw[4];
    if (threadIdx.x == 0) {
        float wt[4];
        band_weights(alpha, wt);
        sw[0] = wt[0]; sw[1] = wt[1]; sw[2] = wt[2]; sw[3] = wt[3];
    }
    __syncthreads();
    const int half = c / 2;
    const float sh = (float)h / (float)H, swd = (float)w / (float)W;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t n; int ch, y, x;
        split_ncij<CH_FAST>(idx, 3 * c, H, W, n, ch, y, x);
        T* q = out + n * os.n + (int64_t)ch * os.c + (int64_t)y * os.h + (int64_t)x * os.w;
        if (ch < c) {
            *q = b[n * bs.n + (int64_t)ch * bs.c + (int64_t)y * bs.h + (int64_t)x * bs.w];
            continue;
        }
        int seg = (ch - c) / half, cc = (ch - c) - seg * half;
        int y0, y1, x0, x1; float ly, lx;
        bilin_src(y, sh, h, y0, y1, ly);
        bilin_src(x, swd, w, x0, x1, lx);
        const Strides4 s = bands.s[seg];
        const T* p = reinterpret_cast<const T*>(bands.p[seg]) + n * s.n + (int64_t)cc * s.c;
        float v00 = to_f(p[(int64_t)y0 * s.h + (int64_t)x0 * s.w]), v01 = to_f(p[(int64_t)y0 * s.h + (int64_t)x1 * s.w]);
        float v10 = to_f(p[(int64_t)y1 * s.h + (int64_t)x0 * s.w]), v11 = to_f(p[(int64_t)y1 * s.h + (int64_t)x1 * s.w]);
        float r = ((1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11)) * sw[seg];
        *q = from_f<T>(r);
    }
}

// ---------------------------------------------------------------------------- merge backward
// (1) gb = gout[:, :c];  (2) gband_i = w_i * up^T(gout_i)  (gather form, no atomics);
// (3) galpha_w[i] += <gout_i, up(band_i)>.
template <typename T, bool CH_FAST>
__global__ void __launch_bounds__(256) merge_bwd_gb(const T* __restrict__ go, Strides4 gos, T* __restrict__ gb, Strides4 gbs, int c, int H, int W,
                                                    int64_t total) {
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t n; int ch, y, x;
        split_ncij<CH_FAST>(idx, c, H, W, n, ch, y, x);
        gb[n * gbs.n + (int64_t)ch * gbs.c + (int64_t)y * gbs.h + (int64_t)x * gbs.w] =
            go[n * gos.n + (int64_t)ch * gos.c + (int64_t)y * gos.h + (int64_t)x * gos.w];
    }
}

// range of destination indices whose bilinear footprint can touch source index s
__device__ __forceinline__ void dst_range(int s, float scale, int n_out, int& lo, int& hi) {
    float inv = 1.f / scale;
    lo = (int)floorf((s - 1 + 0.5f) * inv - 0.5f) - 1;
    hi = (int)ceilf((s + 1 + 0.5f) * inv - 0.5f) + 1;
    lo = lo < 0 ? 0 : lo;
    hi = hi > n_out - 1 ? n_out - 1 : hi;
}

template <typename T, bool CH_FAST>
__global__ void __launch_bounds__(256) merge_bwd_bands(const T* __restrict__ go, Strides4 gos, const float* __restrict__ alpha, BandPtrsMut gband,
                                                       int c, int H, int W, int h, int w, int64_t total) {
    __shared__ float sw[4];
    if (threadIdx.x == 0) {
        float wt[4];
        band_weights(alpha, wt);
        sw[0] = wt[0]; sw[1] = wt[1]; sw[2] = wt[2]; sw[3] = wt[3];
    }
    __syncthreads();
    const int half = c / 2;
    const float sh = (float)h / (float)H, swd = (float)w / (float)W;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t n; int ch, sy, sx;
        split_ncij<CH_FAST>(idx, 2 * c, h, w, n, ch, sy, sx);  // 4 bands x c/2 channels
        int seg = ch / half, cc = ch - seg * half;
        int ylo, yhi, xlo, xhi;
        dst_range(sy, sh, H, ylo, yhi);
        dst_range(sx, swd, W, xlo, xhi);
        const T* gp = go + n * gos.n + (int64_t)(c + ch) * gos.c;
        float acc = 0.f;
        for (int y = ylo; y <= yhi; ++y) {
            int y0, y1; float ly;
            bilin_src(y, sh, h, y0, y1, ly);
            float wy = (y0 == sy ? 1.f - ly : 0.f) + (y1 == sy ? ly : 0.f);
            if (wy == 0.f) continue;
            for (int x = xlo; x <= xhi; ++x) {
                int x0, x1; float lx;
                bilin_src(x, swd, w, x0, x1, lx);
                float wx = (x0 == sx ? 1.f - lx : 0.f) + (x1 == sx ? lx : 0.f);
                if (wx == 0.f) continue;
                acc += wy * wx * to_f(gp[(int64_t)y * gos.h + (int64_t)x * gos.w]);
            }
        }
        const Strides4 s = gband.s[seg];
        reinterpret_cast<T*>(gband.p[seg])[n * s.n + (int64_t)cc * s.c + (int64_t)sy * s.h + (int64_t)sx * s.w] = from_f<T>(acc * sw[seg]);
    }
}

template <typename T, bool CH_FAST>
__global__ void __launch_bounds__(256) merge_bwd_alpha(const T* __restrict__ go, Strides4 gos, BandPtrs bands, float* __restrict__ galpha_w, int c,
                                                       int H, int W, int h, int w, int64_t total) {
    const int half = c / 2;
    const float sh = (float)h / (float)H, swd = (float)w / (float)W;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t n; int ch, y, x;
        split_ncij<CH_FAST>(idx, 2 * c, H, W, n, ch, y, x);
        int seg = ch / half, cc = ch - seg * half;
        int y0, y1, x0, x1; float ly, lx;
        bilin_src(y, sh, h, y0, y1, ly);
        bilin_src(x, swd, w, x0, x1, lx);
        const Strides4 s = bands.s[seg];
        const T* p = reinterpret_cast<const T*>(bands.p[seg]) + n * s.n + (int64_t)cc * s.c;
        float v00 = to_f(p[(int64_t)y0 * s.h + (int64_t)x0 * s.w]), v01 = to_f(p[(int64_t)y0 * s.h + (int64_t)x1 * s.w]);
        float v10 = to_f(p[(int64_t)y1 * s.h + (int64_t)x0 * s.w]), v11 = to_f(p[(int64_t)y1 * s.h + (int64_t)x1 * s.w]);
        float up = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
        float g = to_f(go[n * gos.n + (int64_t)(c + ch) * gos.c + (int64_t)y * gos.h + (int64_t)x * gos.w]);
        float pr = up * g;
        acc[0] += seg == 0 ? pr : 0.f;
        acc[1] += seg == 1 ? pr : 0.f;
        acc[2] += seg == 2 ? pr : 0.f;
        acc[3] += seg == 3 ? pr : 0.f;
    }
    __shared__ float red[4][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float v = warp_sum(acc[i]);
        if (lane == 0) red[i][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float v = 0.f;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) v += red[threadIdx.x][k];
        atomicAdd(galpha_w + threadIdx.x, v);
    }
}

// --------------------------------------------------------------------------- gated residual
template <typename T>
__global__ void __launch_bounds__(256) gated_cvec(const T* b, Strides4 bs, const T* __restrict__ y, Strides4 ys, const float* __restrict__ gamma,
                                                  T* o, Strides4 os, int CV, int H, int W, int64_t total) {
    constexpr int V = Vec16<T>::N;
    const float g = tanhf(__ldg(gamma));
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int cv = (int)(idx % CV);
        int64_t t = idx / CV;
        int x = (int)(t % W); t /= W;
        int yy = (int)(t % H);
        int64_t n = t / H;
        float fb[V], fy[V], r[V];
        // b may alias o: plain (coherent) load
        unpack<T>(*reinterpret_cast<const uint4*>(b + n * bs.n + (int64_t)yy * bs.h + (int64_t)x * bs.w + cv * V), fb);
        unpack<T>(ldg_stream(y + n * ys.n + (int64_t)yy * ys.h + (int64_t)x * ys.w + cv * V), fy);
#pragma unroll
        for (int e = 0; e < V; ++e) r[e] = fb[e] + g * fy[e];
        *reinterpret_cast<uint4*>(o + n * os.n + (int64_t)yy * os.h + (int64_t)x * os.w + cv * V) = pack<T>(r);
    }
}

template <typename T, bool CH_FAST>
__global__ void __launch_bounds__(256) gated_generic(const T* b, Strides4 bs, const T* __restrict__ y, Strides4 ys, const float* __restrict__ gamma,
                                                     T* o, Strides4 os, int C, int H, int W, int64_t total) {
    const float g = tanhf(__ldg(gamma));
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t n; int c, yy, x;
        split_ncij<CH_FAST>(idx, C, H, W, n, c, yy, x);
        float r = to_f(b[n * bs.n + (int64_t)c * bs.c + (int64_t)yy * bs.h + (int64_t)x * bs.w]) +
                  g * to_f(y[n * ys.n + (int64_t)c * ys.c + (int64_t)yy * ys.h + (int64_t)x * ys.w]);
        o[n * os.n + (int64_t)c * os.c + (int64_t)yy * os.h + (int64_t)x * os.w] = from_f<T>(r);
    }
}

// backward of the gated residual: d/dy = tanh(gamma) g, d/dgamma = (1 - tanh^2) <g, y> (d/db = g is the caller's pass-through)
template <typename T, bool CH_FAST>
__global__ void __launch_bounds__(256) gated_bwd_generic(const T* __restrict__ g, Strides4 gs, const T* __restrict__ y, Strides4 ys,
                                                         const float* __restrict__ gamma, T* __restrict__ gy, Strides4 os, float* __restrict__ ggamma,
                                                         int C, int H, int W, int64_t total) {
    const float t = tanhf(__ldg(gamma));
    float acc = 0.f;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t n; int c, yy, x;
        split_ncij<CH_FAST>(idx, C, H, W, n, c, yy, x);
        const float gv = to_f(g[n * gs.n + (int64_t)c * gs.c + (int64_t)yy * gs.h + (int64_t)x * gs.w]);
        acc += gv * to_f(y[n * ys.n + (int64_t)c * ys.c + (int64_t)yy * ys.h + (int64_t)x * ys.w]);
        gy[n * os.n + (int64_t)c * os.c + (int64_t)yy * os.h + (int64_t)x * os.w] = from_f<T>(gv * t);
    }
    __shared__ float red[8];
    float v = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int k = 0; k < 8; ++k) s += red[k];
        atomicAdd(ggamma, s * (1.f - t * t));
    }
}

// uint8 HWC -> normalised activations (value / 255), any destination strides
template <typename T>
__global__ void __launch_bounds__(256) ingest_u8_kernel(const uint8_t* __restrict__ src, T* __restrict__ dst, Strides4 ds, int H, int W, int64_t total) {
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = idx;  // one thread per pixel: 3 bytes in, 3 elements out
        int x = (int)(t % W); t /= W;
        int y = (int)(t % H);
        int64_t n = t / H;
        const uint8_t* p = src + idx * 3;
        T* q = dst + n * ds.n + (int64_t)y * ds.h + (int64_t)x * ds.w;
        q[0] = from_f<T>((float)p[0] / 255.f);
        q[ds.c] = from_f<T>((float)p[1] / 255.f);
        q[2 * ds.c] = from_f<T>((float)p[2] / 255.f);
    }
}


// =============================================================================================
// Tiled fast paths.  Thread <-> (channel vector cv, column) is fixed for the whole kernel, so every
// division, bilinear column weight and band pointer is computed once; the thread then walks kRows
// rows with pointer increments only.  grid = (column tiles, row chunks, images).
// =============================================================================================
constexpr int kRowsMax = 8;  // rows walked per thread; the host shrinks it until the grid fills the 148 SMs several times over

struct TileMap {
    int cv, col;      // this thread's channel vector and column
    bool active;
};
__device__ __forceinline__ TileMap tile_map(int CV, int cols_per_block, int n_cols) {
    TileMap m;
    const int xi = (int)threadIdx.x / CV;
    m.cv = (int)threadIdx.x - xi * CV;
    m.col = (int)blockIdx.x * cols_per_block + xi;
    m.active = xi < cols_per_block && m.col < n_cols;
    return m;
}

template <typename T>
__global__ void __launch_bounds__(256) dwt_fwd_tiled(const T* __restrict__ x, Strides4 xs, T* __restrict__ o, int64_t ob, Strides4 os, int CV,
                                                     int cols_per_block, int rows, int H2, int W2) {
    constexpr int V = Vec16<T>::N;
    pdl_launch_dependents();
    pdl_wait();
    const TileMap m = tile_map(CV, cols_per_block, W2);
    if (!m.active) return;
    const float k = kHaar;
    const int i0 = (int)blockIdx.y * rows, i1 = min(i0 + rows, H2);
    const int64_t n = blockIdx.z;
    const T* p = x + n * xs.n + (int64_t)(2 * i0) * xs.h + (int64_t)(2 * m.col) * xs.w + m.cv * V;
    T* q = o + n * os.n + (int64_t)i0 * os.h + (int64_t)m.col * os.w + m.cv * V;
#pragma unroll 2
    for (int i = i0; i < i1; ++i, p += 2 * xs.h, q += os.h) {
        float a[V], b[V], c[V], d[V];
        uint4 ra = ldg_stream(p), rb = ldg_stream(p + xs.w), rc = ldg_stream(p + xs.h), rd = ldg_stream(p + xs.h + xs.w);
        unpack<T>(ra, a); unpack<T>(rb, b); unpack<T>(rc, c); unpack<T>(rd, d);
        float ll[V], lh[V], hl[V], hh[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            float s0 = a[e] + b[e], s1 = c[e] + d[e], d0 = a[e] - b[e], d1 = c[e] - d[e];
            ll[e] = k * (s0 + s1); lh[e] = k * (d0 + d1); hl[e] = k * (s0 - s1); hh[e] = k * (d0 - d1);
        }
        *reinterpret_cast<uint4*>(q) = pack<T>(ll);
        *reinterpret_cast<uint4*>(q + ob) = pack<T>(lh);
        *reinterpret_cast<uint4*>(q + 2 * ob) = pack<T>(hl);
        *reinterpret_cast<uint4*>(q + 3 * ob) = pack<T>(hh);
    }
}

constexpr int kMergeRows = 16;

template <typename T>
__global__ void __launch_bounds__(256) merge_fwd_tiled(const T* __restrict__ b, Strides4 bs, BandPtrs bands, const float* __restrict__ alpha,
                                                       T* __restrict__ out, Strides4 os, int c, int CV, int cols_per_block, int rows, int H, int W,
                                                       int h, int w, int c_off) {
    constexpr int V = Vec16<T>::N;
    // per-block tables: the four band weights and the vertical taps of this block's rows (both uniform over the block)
    __shared__ float s_w[4];
    __shared__ int s_ya[kMergeRows], s_yb[kMergeRows];
    __shared__ float s_ly[kMergeRows];
    const int y0b = (int)blockIdx.y * rows, y1b = min(y0b + rows, H);
    if (threadIdx.x == 0) {
        float wt[4];
        band_weights(alpha, wt);
        s_w[0] = wt[0]; s_w[1] = wt[1]; s_w[2] = wt[2]; s_w[3] = wt[3];
    }
    if ((int)threadIdx.x < y1b - y0b) {
        int ya, yb; float ly;
        bilin_src(y0b + threadIdx.x, (float)h / (float)H, h, ya, yb, ly);
        s_ya[threadIdx.x] = ya; s_yb[threadIdx.x] = yb; s_ly[threadIdx.x] = ly;
    }
    __syncthreads();
    const TileMap m = tile_map(CV, cols_per_block, W);
    if (!m.active) return;
    const int64_t n = blockIdx.z;
    const int ch = m.cv * V + c_off;  // c_off = c: bands-only output (the pass-through copy of b is skipped)
    T* q = out + n * os.n + (int64_t)y0b * os.h + (int64_t)m.col * os.w + (ch - c_off);
    if (ch < c) {  // pass-through copy of b
        const T* p = b + n * bs.n + (int64_t)y0b * bs.h + (int64_t)m.col * bs.w + ch;
#pragma unroll 4
        for (int y = y0b; y < y1b; ++y, p += bs.h, q += os.h) stg_stream(q, ldg_stream(p));
        return;
    }
    const int half = c / 2, seg = (ch - c) / half, cc = (ch - c) - seg * half;
    int x0, x1; float lx;
    bilin_src(m.col, (float)w / (float)W, w, x0, x1, lx);
    const Strides4 s = bands.s[seg];
    const T* p0 = reinterpret_cast<const T*>(bands.p[seg]) + n * s.n + cc + (int64_t)x0 * s.w;
    const T* p1 = reinterpret_cast<const T*>(bands.p[seg]) + n * s.n + cc + (int64_t)x1 * s.w;
    const float wx0 = 1.f - lx;
    // Horizontally blended source rows are cached in registers: consecutive output rows share their two source rows
    // (for the exact 2x case each source row pair serves two output rows), so a row costs one pair of 16 B loads
    // instead of four and half the unpack work.  The row index is block-uniform: no divergence.
    int ca = -1, cb = -1;
    float ha[V], hb[V];
    auto hrow = [&](int yy, float (&dst)[V]) {
        float v0[V], v1[V];
        unpack<T>(ldg_cached(p0 + (int64_t)yy * s.h), v0);  // each band pixel feeds ~4 output pixels: keep it in L1
        unpack<T>(ldg_cached(p1 + (int64_t)yy * s.h), v1);
#pragma unroll
        for (int e = 0; e < V; ++e) dst[e] = wx0 * v0[e] + lx * v1[e];
    };
    const float wb = s_w[seg];
    for (int y = y0b; y < y1b; ++y, q += os.h) {
        const int ya = s_ya[y - y0b], yb = s_yb[y - y0b];
        const float ly = s_ly[y - y0b];
        if (ya != ca) {
            if (ya == cb) {
#pragma unroll
                for (int e = 0; e < V; ++e) ha[e] = hb[e];
            } else {
                hrow(ya, ha);
            }
            ca = ya;
        }
        if (yb != cb) {
            if (yb == ca) {
#pragma unroll
                for (int e = 0; e < V; ++e) hb[e] = ha[e];
            } else {
                hrow(yb, hb);
            }
            cb = yb;
        }
        const float wy0 = 1.f - ly;
        float r[V];
#pragma unroll
        for (int e = 0; e < V; ++e) r[e] = (wy0 * ha[e] + ly * hb[e]) * wb;
        stg_stream(q, pack<T>(r));
    }
}

// Exact 2x case (H = 2h, W = 2w: every EdgeLine site at 640^2 / 1280^2).  Output rows 2k+1 and 2k+2 both blend source rows
// k and k+1 (weights .75/.25 and .25/.75), so a thread walks SOURCE rows: one pair of 16 B loads and one horizontal blend
// per source row feed two output rows; the band weight is folded into the horizontal taps.
// The kernel was load-latency bound (ncu: long_scoreboard 11-19 of ~20 stall cycles per issue, one dependent L2 round trip per
// source row): all 2*(SR+1) loads of a thread are now issued before the first use, and the band weights are evaluated per warp
// (softplus in lanes, shuffles) instead of thread 0 + a block barrier.
template <typename T, int SR>
__global__ void __launch_bounds__(256) merge_fwd_x2(const T* __restrict__ b, Strides4 bs, BandPtrs bands, const float* __restrict__ alpha,
                                                    T* __restrict__ out, Strides4 os, int c, int CV, int cols_per_block, int h, int w, int c_off) {
    constexpr int V = Vec16<T>::N;
    pdl_launch_dependents();
    const int H = 2 * h, W = 2 * w;
    const TileMap m = tile_map(CV, cols_per_block, W);
    const int ch = m.cv * V + c_off;  // c_off = c: bands-only output (the pass-through copy of b is skipped)
    const int half = c / 2;
    const int seg = ch >= c ? min((ch - c) / half, 3) : 0, cc = (ch - c) - seg * half;
    float wb;
    {   // softplus(alpha) / (sum + 1e-6), same operation order as band_weights()
        const float a = __ldg(alpha + (threadIdx.x & 3));
        const float sp = a > 20.f ? a : log1pf(expf(a));
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) s += __shfl_sync(0xffffffffu, sp, i);
        s += 1e-6f;
        wb = __shfl_sync(0xffffffffu, sp, seg) / s;
    }
    if (!m.active) return;
    const int k0 = (int)blockIdx.y * SR, k1 = min(k0 + SR, h);
    const int64_t n = blockIdx.z;
    // this block writes output rows 2*k0+1 .. 2*k1 (clipped to H-1), plus row 0 when k0 == 0
    const int y_first = k0 == 0 ? 0 : 2 * k0 + 1, y_last = min(2 * k1, H - 1);
    if (ch < c) {  // pass-through copy of b
        const T* p = b + n * bs.n + (int64_t)y_first * bs.h + (int64_t)m.col * bs.w + ch;
        T* q = out + n * os.n + (int64_t)y_first * os.h + (int64_t)m.col * os.w + (ch - c_off);
#pragma unroll 4
        for (int y = y_first; y <= y_last; ++y, p += bs.h, q += os.h) stg_stream(q, ldg_stream(p));
        return;
    }
    int x0, x1; float lx;
    bilin_src(m.col, 0.5f, w, x0, x1, lx);
    const float ax = (1.f - lx) * wb, bx = lx * wb;
    const Strides4 s = bands.s[seg];
    const T* pb = reinterpret_cast<const T*>(bands.p[seg]) + n * s.n + cc + (int64_t)x0 * s.w;
    const int64_t dx = (int64_t)(x1 - x0) * s.w;
    uint4 ra[SR + 1], rb[SR + 1];  // source rows k0 .. k0+SR (clamped at the bottom edge): every load in flight at once
#pragma unroll
    for (int i = 0; i <= SR; ++i) {
        const T* p = pb + (int64_t)min(k0 + i, h - 1) * s.h;
        ra[i] = ldg_cached(p);       // each band pixel feeds ~4 output pixels: keep it in L1
        rb[i] = ldg_cached(p + dx);
    }
    T* q = out + n * os.n + (int64_t)y_first * os.h + (int64_t)m.col * os.w + (ch - c_off);
    const int64_t oh = os.h;
    float prev[V], cur[V], r[V];
    auto hrow = [&](uint4 a0, uint4 a1, float (&dst)[V]) {
        float v0[V], v1[V];
        unpack<T>(a0, v0);
        unpack<T>(a1, v1);
#pragma unroll
        for (int e = 0; e < V; ++e) dst[e] = ax * v0[e] + bx * v1[e];
    };
    hrow(ra[0], rb[0], prev);
    if (k0 == 0) {  // output row 0 = source row 0
        stg_stream(q, pack<T>(prev));
        q += oh;
    }
#pragma unroll
    for (int i = 0; i < SR; ++i) {
        const int k = k0 + i;
        if (k < k1) {
            hrow(ra[i + 1], rb[i + 1], cur);
#pragma unroll
            for (int e = 0; e < V; ++e) r[e] = 0.75f * prev[e] + 0.25f * cur[e];
            stg_stream(q, pack<T>(r));
            q += oh;
            if (2 * k + 2 < H) {
#pragma unroll
                for (int e = 0; e < V; ++e) r[e] = 0.25f * prev[e] + 0.75f * cur[e];
                stg_stream(q, pack<T>(r));
                q += oh;
            }
#pragma unroll
            for (int e = 0; e < V; ++e) prev[e] = cur[e];
        }
    }
}

// Exact 2x case, shared-memory staged and software-pipelined (the default for channel-vectorised views).
// What the profiles of the earlier versions said (round 2, largest site, B = 64, c = 16, 160 x 160):
//   * merge_fwd_x2 (registers): 18.5 M warp instructions for 205 k warp-level stores -- 90 per store, of which ~30 are the blend; the rest
//     is per-thread set-up (softplus, 64-bit index products, ten address computations) amortised over eight stores, and the L1 data
//     pipe is 72 % busy because a 16 B-per-lane load whose lanes alternate between the four band tensors costs 16 wavefronts;
//   * a first staged version (one CTA = one row chunk: cp.async -> barrier -> blend -> exit) halved the instructions but every CTA of a
//     wave sat in the same phase: nobody stored while everybody waited for its loads (time = a whole number of ~3.9 us waves).
// This version: a CTA (channel vector x output column threads) owns a column tile and a RANGE of source rows and streams down it in
// steps of SR rows through a double buffer: the cp.async loads of step s+2 are in flight while step s is blended and stored, the
// horizontally blended last row of a step stays in registers (no halo re-read between steps), set-up is paid once per range.
//   thread (cv, y) stages the fixed slot (band, vector, source column y) of every row (no index arithmetic in the loop); the 4x re-use
//   of a band pixel is served by LDS.128; 16-bit maps blend with packed fp32x2 arithmetic (horizontal 2 ops / pair, vertical through
//   the shared difference cur - prev 3 ops / pair).  Tile row layout [band][source column][vector] with a band pitch of S 16-byte units
//   (S chosen on the host so the 16 B reads of a quarter warp fall into distinct banks).
//   DRAM reads = band bytes x (cpb/2+2)/(cpb/2) column halo x (RPC+1)/RPC row halo.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ f32x2 mul_f32x2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// 16 bytes of T as packed fp32 pairs (16-bit types: 4 pairs; fp32: 2 pairs)
template <typename T> struct Pairs { static constexpr int N = Vec16<T>::N / 2; };
template <typename T> __device__ __forceinline__ void unpack_pairs(uint4 raw, f32x2 (&p)[Pairs<T>::N]) {
    float f[Vec16<T>::N];
    unpack<T>(raw, f);
#pragma unroll
    for (int i = 0; i < Pairs<T>::N; ++i) p[i] = pack_f32x2(f[2 * i], f[2 * i + 1]);
}
template <typename T> __device__ __forceinline__ uint4 pack_pairs(const f32x2 (&p)[Pairs<T>::N]) {
    float f[Vec16<T>::N];
#pragma unroll
    for (int i = 0; i < Pairs<T>::N; ++i) unpack_f32x2(p[i], f[2 * i], f[2 * i + 1]);
    return pack<T>(f);
}

// blend state of one thread: the horizontally blended previous source row.  16-bit maps keep it as packed fp32 pairs.
template <typename T, bool FAST> struct MergeRow;
template <typename T> struct MergeRow<T, true> {
    static constexpr int NP = Pairs<T>::N;
    f32x2 v[NP];
    __device__ __forceinline__ void hblend(uint4 a0, uint4 a1, float ax, float bx) {
        const f32x2 ax2 = pack_f32x2(ax, ax), bx2 = pack_f32x2(bx, bx);
        f32x2 v0[NP], v1[NP];
        unpack_pairs<T>(a0, v0);
        unpack_pairs<T>(a1, v1);
#pragma unroll
        for (int e = 0; e < NP; ++e) v[e] = fma_f32x2(v1[e], bx2, mul_f32x2(v0[e], ax2));
    }
    __device__ __forceinline__ uint4 packed() const { return pack_pairs<T>(v); }
    // rows 2k+1 = .75 prev + .25 cur and 2k+2 = .25 prev + .75 cur through the shared difference d = cur - prev
    __device__ __forceinline__ void vblend(const MergeRow& cur, uint4& lo, uint4& hi) const {
        const f32x2 neg1 = pack_f32x2(-1.f, -1.f), q25 = pack_f32x2(0.25f, 0.25f), nq25 = pack_f32x2(-0.25f, -0.25f);
        f32x2 r0[NP], r1[NP];
#pragma unroll
        for (int e = 0; e < NP; ++e) {
            const f32x2 d = fma_f32x2(v[e], neg1, cur.v[e]);
            r0[e] = fma_f32x2(d, q25, v[e]);
            r1[e] = fma_f32x2(d, nq25, cur.v[e]);
        }
        lo = pack_pairs<T>(r0);
        hi = pack_pairs<T>(r1);
    }
};
template <typename T> struct MergeRow<T, false> {  // fp32 maps: the reference's operation order, scalar IEEE arithmetic
    static constexpr int V = Vec16<T>::N;
    float v[V];
    __device__ __forceinline__ void hblend(uint4 a0, uint4 a1, float ax, float bx) {
        float v0[V], v1[V];
        unpack<T>(a0, v0);
        unpack<T>(a1, v1);
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = ax * v0[e] + bx * v1[e];
    }
    __device__ __forceinline__ uint4 packed() const { return pack<T>(v); }
    __device__ __forceinline__ void vblend(const MergeRow& cur, uint4& lo, uint4& hi) const {
        float r0[V], r1[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            r0[e] = 0.75f * v[e] + 0.25f * cur.v[e];
            r1[e] = 0.25f * v[e] + 0.75f * cur.v[e];
        }
        lo = pack<T>(r0);
        hi = pack<T>(r1);
    }
};

template <typename T>
__global__ void __launch_bounds__(256) merge_fwd_x2p(const T* __restrict__ b, Strides4 bs, BandPtrs bands, const float* __restrict__ alpha,
                                                     T* __restrict__ out, Strides4 os, int c, int CVb, int hv, int hv_shift, int SR, int RPC, int h,
                                                     int w, int S) {
    constexpr int V = Vec16<T>::N;
    constexpr bool kFast = sizeof(T) == 2;  // 16-bit maps: SFU softplus, packed fp32x2 blends
    extern __shared__ __align__(16) unsigned char merge_smem[];
    pdl_launch_dependents();
    const int H = 2 * h, W = 2 * w;
    const int cv = (int)threadIdx.x, xi = (int)threadIdx.y, cpb = (int)blockDim.y;
    const int X0 = (int)blockIdx.x * cpb, X = X0 + xi;
    const int u = max(cv - CVb, 0);
    const int seg = min(hv_shift >= 0 ? u >> hv_shift : u / hv, 3), vi = u - seg * hv;
    float wb;
    {   // softplus(alpha) / (sum + 1e-6), same operation order as band_weights()
        const int lane = (cv + (int)blockDim.x * xi) & 31;
        const float a = __ldg(alpha + (lane & 3));
        const float sp = a > 20.f ? a : (kFast ? __logf(1.f + __expf(a)) : log1pf(expf(a)));
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) s += __shfl_sync(0xffffffffu, sp, i);
        s += 1e-6f;
        const float spb = __shfl_sync(0xffffffffu, sp, seg);
        wb = kFast ? __fdividef(spb, s) : spb / s;
    }
    const int ka = (int)blockIdx.y * RPC, kb = min(ka + RPC, h);  // source-row intervals [ka, kb) of this CTA: needs rows ka .. kb (clamped)
    const int y_first = ka == 0 ? 0 : 2 * ka + 1;                 // output rows y_first .. min(2 kb, H - 1)
    pdl_wait();  // everything above overlapped the producer's tail; activations are read (and `out` is written) only from here on
    if (cv < CVb) {  // pass-through copy of b (full-concat form only); these threads take no part in the staging
        if (X < W) {
            const int y_last = min(2 * kb, H - 1);
            const T* p = b + (int64_t)blockIdx.z * bs.n + (int64_t)y_first * bs.h + (int64_t)X * bs.w + cv * V;
            T* q = out + (int64_t)blockIdx.z * os.n + (int64_t)y_first * os.h + (int64_t)X * os.w + cv * V;
#pragma unroll 4
            for (int y = y_first; y <= y_last; ++y, p += bs.h, q += os.h) stg_stream(q, ldg_stream(p));
        }
        return;  // exited threads do not count at the CTA barrier
    }
    const int n_scols = cpb / 2 + 2, sc0 = max(X0 / 2 - 1, 0);
    const int row_units = 4 * S;
    uint4* const row0 = reinterpret_cast<uint4*>(merge_smem);      // source row ka
    uint4* const ring = row0 + row_units;                           // [2][SR] rows: buffer (s & 1) holds rows ka + s SR + 1 .. ka + (s+1) SR
    // ---- staging: this thread's slot (band seg, vector vi, source column sc0 + xi) of consecutive rows, one running source pointer
    const bool stager = xi < n_scols;
    const Strides4 sb = bands.s[seg];
    const T* src = reinterpret_cast<const T*>(bands.p[seg]) + (int64_t)blockIdx.z * sb.n + (int64_t)ka * sb.h + (int64_t)min(sc0 + xi, w - 1) * sb.w + vi * V;
    int src_row = ka;                                                // row `src` points at; rows past the bottom edge repeat row h - 1
    const int slot = seg * S + xi * hv + vi;
    auto stage_rows = [&](uint4* dst, int n) {                      // next n rows -> dst[0 .. n)
        if (stager) {
            for (int r = 0; r < n; ++r, dst += row_units) {
                cp_async16(dst + slot, src);
                if (src_row < h - 1) { src += sb.h; ++src_row; }
            }
        }
        cp_async_commit();
    };
    const int n_int = kb - ka, n_steps = (n_int + SR - 1) / SR;
    auto step_rows = [&](int s) { return min(SR, n_int - s * SR); };  // new rows (= intervals) of step s
    // group 0: row ka + the rows of step 0; group 1: the rows of step 1
    if (stager) {
        cp_async16(row0 + slot, src);
        if (src_row < h - 1) { src += sb.h; ++src_row; }
    }
    stage_rows(ring, step_rows(0));
    if (n_steps > 1) stage_rows(ring + SR * row_units, step_rows(1)); else cp_async_commit();

    const bool active = X < W;
    // PyTorch's source rule for exact 2x: X = 2m -> (m-1, m) with lambda .75 (X = 0: source 0 alone); X = 2m+1 -> (m, m+1) with lambda .25
    const int m = X >> 1, odd = X & 1;
    const int x0 = min(odd ? m : max(m - 1, 0), w - 1), x1 = min(x0 + 1, w - 1);
    const float lx = odd ? 0.25f : (m == 0 ? 0.f : 0.75f);
    const float ax = (1.f - lx) * wb, bx = lx * wb;
    const int rd0 = seg * S + (x0 - sc0) * hv + vi, dx = (x1 - x0) * hv;
    T* q = out + (int64_t)blockIdx.z * os.n + (int64_t)y_first * os.h + (int64_t)min(X, W - 1) * os.w + cv * V;
    const int64_t oh = os.h;
    MergeRow<T, kFast> prev, cur;
    int k = ka;
    for (int s = 0; s < n_steps; ++s) {
        cp_async_wait<1>();   // this thread's copies of step s (and row ka) have landed; step s+1 may still be in flight
        __syncthreads();      // ... and everybody else's
        const uint4* buf = ring + (s & 1) * SR * row_units + rd0;
        if (active) {
            if (s == 0) {
                prev.hblend(row0[rd0], row0[rd0 + dx], ax, bx);
                if (ka == 0) {  // output row 0 = source row 0
                    stg_stream(q, prev.packed());
                    q += oh;
                }
            }
            const int n = step_rows(s);
            for (int i = 0; i < n; ++i, ++k, buf += row_units) {
                cur.hblend(buf[0], buf[dx], ax, bx);
                uint4 lo, hi;
                prev.vblend(cur, lo, hi);
                stg_stream(q, lo);
                q += oh;
                if (k < h - 1) {  // the last interval of the image produces one row
                    stg_stream(q, hi);
                    q += oh;
                }
                prev = cur;
            }
        }
        __syncthreads();      // buffer (s & 1) is free again
        if (s + 2 < n_steps) stage_rows(ring + (s & 1) * SR * row_units, step_rows(s + 2)); else cp_async_commit();
    }
}

// host-side plan of merge_fwd_x2p: column tile that wastes the fewest lanes, row range that gives every SM several resident CTAs
struct MergePlan { int cpb, SR, RPC, S; dim3 grid, block; size_t smem; bool ok; };
static inline MergePlan merge_plan(int CVt, int hv, int W, int h, int B, int force_sr, int force_rpc) {
    MergePlan p{};
    double best = -1.0;
    for (int cpb = (256 / CVt) & ~1; cpb >= 4; cpb -= 2) {  // cpb >= 4: the cpb threads of a channel vector stage the cpb/2 + 2 source columns
        // useful lanes / launched lanes (ragged last column tile, last warp of the CTA); ties go to the wider tile (less column halo)
        const int tiles = (int)ceil_div(W, cpb);
        const double eff = (double)W * CVt / ((double)tiles * ceil_div(cpb * CVt, 32) * 32);
        if (eff > best + 1e-9) { best = eff; p.cpb = cpb; }
    }
    if (best < 0.0) return p;
    const int n_scols = p.cpb / 2 + 2;
    p.S = n_scols * hv;
    const int want = hv >= 8 ? 0 : (hv == 4 ? 4 : 2);  // band pitch mod 8 (16 B units): distinct banks for the lanes of a quarter warp
    if (hv < 8) while (p.S % 8 != want) ++p.S;
    const int64_t col_tiles = ceil_div(W, p.cpb);
    p.SR = force_sr > 0 ? force_sr : 4;
    auto bytes = [&](int SR) { return (size_t)(2 * SR + 1) * 4 * p.S * 16; };
    while (p.SR > 1 && bytes(p.SR) > 40 * 1024) p.SR >>= 1;
    p.smem = bytes(p.SR);
    if (p.smem > 96 * 1024) return p;
    // row ranges: ~8 CTAs per SM over the whole grid, at least two pipeline steps per CTA where the map is tall enough
    int64_t splits = ceil_div((int64_t)kSMs * 8, col_tiles * B);
    const int64_t max_splits = ceil_div(h, 2 * p.SR);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.RPC = (int)ceil_div(h, splits);  // balanced ranges; the last step of a range may be short
    if (force_rpc > 0) p.RPC = force_rpc;
    p.grid = dim3((unsigned)col_tiles, (unsigned)ceil_div(h, p.RPC), (unsigned)B);
    p.block = dim3((unsigned)CVt, (unsigned)p.cpb, 1);
    p.ok = true;
    return p;
}

template <typename T>
__global__ void __launch_bounds__(256) gated_tiled(const T* b, Strides4 bs, const T* __restrict__ y, Strides4 ys, const float* __restrict__ gamma, T* o,
                                                   Strides4 os, T* o2, Strides4 os2, int CV, int cols_per_block, int rows, int H, int W) {
    constexpr int V = Vec16<T>::N;
    const TileMap m = tile_map(CV, cols_per_block, W);
    if (!m.active) return;
    const float g = tanhf(__ldg(gamma));
    const int r0 = (int)blockIdx.y * rows, r1 = min(r0 + rows, H);
    const int64_t n = blockIdx.z;
    const T* pb = b + n * bs.n + (int64_t)r0 * bs.h + (int64_t)m.col * bs.w + m.cv * V;
    const T* py = y + n * ys.n + (int64_t)r0 * ys.h + (int64_t)m.col * ys.w + m.cv * V;
    T* po = o + n * os.n + (int64_t)r0 * os.h + (int64_t)m.col * os.w + m.cv * V;
    T* po2 = o2 ? o2 + n * os2.n + (int64_t)r0 * os2.h + (int64_t)m.col * os2.w + m.cv * V : nullptr;
#pragma unroll 4
    for (int r = r0; r < r1; ++r, pb += bs.h, py += ys.h, po += os.h) {
        float fb[V], fy[V], res[V];
        unpack<T>(*reinterpret_cast<const uint4*>(pb), fb);  // b may alias o: coherent load
        unpack<T>(ldg_stream(py), fy);
#pragma unroll
        for (int e = 0; e < V; ++e) res[e] = fb[e] + g * fy[e];
        const uint4 pk = pack<T>(res);
        *reinterpret_cast<uint4*>(po) = pk;
        if (po2) { *reinterpret_cast<uint4*>(po2) = pk; po2 += os2.h; }  // second copy, e.g. the block's concat slice
    }
}

static inline dim3 tile_grid(int CV, int n_cols, int n_rows, int B, int& cols_per_block, int& rows, int rows_max = kRowsMax, int rows_min = 1) {
    cols_per_block = 256 / CV < n_cols ? 256 / CV : n_cols;
    const int64_t col_tiles = ceil_div(n_cols, cols_per_block);
    rows = rows_max;
    while (rows > rows_min && col_tiles * ceil_div(n_rows, rows) * B < (int64_t)kSMs * 16) rows >>= 1;  // small maps: more, shorter threads
    return dim3((unsigned)col_tiles, (unsigned)ceil_div(n_rows, rows), (unsigned)B);
}

// grid for a streaming grid-stride kernel: enough CTAs to cover `total`, capped at a multiple
// of the SM count so the tail wave is full (148 SMs x 8 resident 256-thread CTAs)
static inline int stream_grid(int64_t total, int threads = 256) {
    int64_t need = ceil_div(total, threads);
    int64_t cap = (int64_t)kSMs * 16;
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace el

using namespace el;

extern "C" int el_dwt_haar_fwd(const void* x, const int64_t xs_[4], void* bands, const int64_t bs_[5], int B, int C, int H, int W, int dtype,
                               void* stream) {
    if (!x || !bands || !xs_ || !bs_ || B <= 0 || C <= 0 || H <= 0 || W <= 0) return EL_ERR_ARG;
    const int H2 = H / 2, W2 = W / 2;
    if (H2 == 0 || W2 == 0) return EL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    Strides4 xs = s4(xs_), os = s4(bs_ + 1);
    const int64_t ob = bs_[0];
    EL_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec16<T>::N;
        if (channel_vectorisable<T>(x, xs, C) && channel_vectorisable<T>(bands, os, C) && ob % V == 0) {
            if (C / V <= 256 && B <= 65535) {
                int cpb, rows;
                dim3 g = tile_grid(C / V, W2, H2, B, cpb, rows);
                launch_pdl(dwt_fwd_tiled<T>, g, dim3(256), 0, st, (const T*)x, xs, (T*)bands, ob, os, C / V, cpb, rows, H2, W2);
            } else {
                int64_t total = (int64_t)B * H2 * W2 * (C / V);
                dwt_fwd_cvec<T><<<stream_grid(total), 256, 0, st>>>((const T*)x, xs, (T*)bands, ob, os, C / V, H2, W2, total);
            }
        } else {
            int64_t total = (int64_t)B * C * H2 * W2;
            if (xs.c == 1)
                dwt_fwd_generic<T, true><<<stream_grid(total), 256, 0, st>>>((const T*)x, xs, (T*)bands, ob, os, C, H2, W2, total);
            else
                dwt_fwd_generic<T, false><<<stream_grid(total), 256, 0, st>>>((const T*)x, xs, (T*)bands, ob, os, C, H2, W2, total);
        }
    });
    note_launches(1);
    return check_launch();
}

extern "C" int el_dwt_haar_bwd(const void* g, const int64_t gs_[5], void* gx, const int64_t gxs_[4], int B, int C, int H, int W, int dtype,
                               void* stream) {
    if (!g || !gx || !gs_ || !gxs_ || B <= 0 || C <= 0 || H <= 0 || W <= 0) return EL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Strides4 gs = s4(gs_ + 1), os = s4(gxs_);
    int64_t total = (int64_t)B * C * ((H + 1) / 2) * ((W + 1) / 2);
    EL_DISPATCH_DTYPE(dtype, {
        if (os.c == 1)
            dwt_bwd_generic<T, true><<<stream_grid(total), 256, 0, st>>>((const T*)g, gs_[0], gs, (T*)gx, os, C, H, W, total);
        else
            dwt_bwd_generic<T, false><<<stream_grid(total), 256, 0, st>>>((const T*)g, gs_[0], gs, (T*)gx, os, C, H, W, total);
    });
    note_launches(1);
    return check_launch();
}

extern "C" int el_wave_merge_fwd(const void* b, const int64_t bs_[4], const void* const band[4], const int64_t band_s[16], const float* alpha,
                                 void* out, const int64_t os_[4], int B, int c, int H, int W, int h, int w, int dtype, void* stream) {
    if (!band || !band_s || !alpha || !out || (b && !bs_) || B <= 0 || c <= 0 || (c & 1) || H <= 0 || W <= 0 || h <= 0 || w <= 0) return EL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    // b == NULL: bands-only mode, out has 2c channels [LLu | LHu | HLu | HHu] (the consumer reads b in place)
    const int c_off = b ? 0 : c, co = b ? 3 * c : 2 * c;
    Strides4 bs = b ? s4(bs_) : Strides4{0, 1, 0, 0}, os = s4(os_);
    BandPtrs bp;
    for (int i = 0; i < 4; ++i) {
        if (!band[i]) return EL_ERR_ARG;
        bp.p[i] = band[i];
        bp.s[i] = s4(band_s + 4 * i);
        // the stride of a size-1 dimension is arbitrary (PyTorch reports 1 for a (B, C, 1, 1) tensor in either memory format): never stepped
        if (h == 1) bp.s[i].h = 0;
        if (w == 1) bp.s[i].w = 0;
        if (B == 1) bp.s[i].n = 0;
    }
    EL_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec16<T>::N;
        bool vec = (!b || channel_vectorisable<T>(b, bs, c)) && channel_vectorisable<T>(out, os, co) && (c / 2) % V == 0;
        for (int i = 0; i < 4; ++i) vec = vec && channel_vectorisable<T>(bp.p[i], bp.s[i], c / 2);
        if (!b && !(vec && co / V <= 256 && B <= 65535)) return EL_ERR_UNSUPPORTED;
        if (vec) {
            static const int use_smem = [] { const char* e = getenv("EL_MERGE_SMEM"); return e ? atoi(e) : 1; }();
            static const int force_sr = [] { const char* e = getenv("EL_MERGE_SR"); return e ? atoi(e) : 0; }();    // experiment knobs
            static const int force_rpc = [] { const char* e = getenv("EL_MERGE_RPC"); return e ? atoi(e) : 0; }();
            MergePlan mp{};
            if (use_smem && co / V <= 256 && B <= 65535 && H == 2 * h && W == 2 * w) mp = merge_plan(co / V, (c / 2) / V, W, h, B, force_sr, force_rpc);
            if (mp.ok) {
                auto kern = merge_fwd_x2p<T>;
                if (mp.smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mp.smem);
                const int hvv = (c / 2) / V;
                int hv_shift = -1;
                for (int sft = 0; sft < 16; ++sft) if ((1 << sft) == hvv) hv_shift = sft;
                launch_pdl(kern, mp.grid, mp.block, mp.smem, st, (const T*)b, bs, bp, alpha, (T*)out, os, c, b ? c / V : 0, hvv, hv_shift, mp.SR, mp.RPC, h, w, mp.S);
            } else if (co / V <= 256 && B <= 65535 && H == 2 * h && W == 2 * w) {
                int cpb, srows;
                dim3 g = tile_grid(co / V, W, h, B, cpb, srows, 4, 2);  // rows here = SOURCE rows per thread
                if (srows == 4) merge_fwd_x2<T, 4><<<g, 256, 0, st>>>((const T*)b, bs, bp, alpha, (T*)out, os, c, co / V, cpb, h, w, c_off);
                else merge_fwd_x2<T, 2><<<g, 256, 0, st>>>((const T*)b, bs, bp, alpha, (T*)out, os, c, co / V, cpb, h, w, c_off);
            } else if (co / V <= 256 && B <= 65535) {
                int cpb, rows;
                dim3 g = tile_grid(co / V, W, H, B, cpb, rows, kMergeRows, 4);
                merge_fwd_tiled<T><<<g, 256, 0, st>>>((const T*)b, bs, bp, alpha, (T*)out, os, c, co / V, cpb, rows, H, W, h, w, c_off);
            } else {
                int64_t total = (int64_t)B * H * W * (3 * c / V);
                merge_fwd_cvec<T><<<stream_grid(total), 256, 0, st>>>((const T*)b, bs, bp, alpha, (T*)out, os, c, H, W, h, w, total);
            }
        } else {
            int64_t total = (int64_t)B * 3 * c * H * W;
            if (os.c == 1)
                merge_fwd_generic<T, true><<<stream_grid(total), 256, 0, st>>>((const T*)b, bs, bp, alpha, (T*)out, os, c, H, W, h, w, total);
            else
                merge_fwd_generic<T, false><<<stream_grid(total), 256, 0, st>>>((const T*)b, bs, bp, alpha, (T*)out, os, c, H, W, h, w, total);
        }
    });
    note_launches(1);
    return check_launch();
}

extern "C" int el_wave_merge_bwd(const void* gout, const int64_t gos_[4], const void* const band[4], const int64_t band_s[16], const float* alpha,
                                 void* gb, const int64_t gbs_[4], void* const gband[4], const int64_t gband_s[16], float* galpha_w, int B, int c,
                                 int H, int W, int h, int w, int dtype, void* stream) {
    if (!gout || !alpha || B <= 0 || c <= 0 || (c & 1) || H <= 0 || W <= 0 || h <= 0 || w <= 0) return EL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Strides4 gos = s4(gos_);
    EL_DISPATCH_DTYPE(dtype, {
        if (gb) {
            Strides4 gbs = s4(gbs_);
            int64_t total = (int64_t)B * c * H * W;
            if (gos.c == 1)
                merge_bwd_gb<T, true><<<stream_grid(total), 256, 0, st>>>((const T*)gout, gos, (T*)gb, gbs, c, H, W, total);
            else
                merge_bwd_gb<T, false><<<stream_grid(total), 256, 0, st>>>((const T*)gout, gos, (T*)gb, gbs, c, H, W, total);
        }
        if (gband) {
            BandPtrsMut gp;
            for (int i = 0; i < 4; ++i) {
                if (!gband[i]) return EL_ERR_ARG;
                gp.p[i] = gband[i];
                gp.s[i] = s4(gband_s + 4 * i);
            }
            int64_t total = (int64_t)B * 2 * c * h * w;
            if (gp.s[0].c == 1)
                merge_bwd_bands<T, true><<<stream_grid(total), 256, 0, st>>>((const T*)gout, gos, alpha, gp, c, H, W, h, w, total);
            else
                merge_bwd_bands<T, false><<<stream_grid(total), 256, 0, st>>>((const T*)gout, gos, alpha, gp, c, H, W, h, w, total);
        }
        if (galpha_w) {
            if (!band) return EL_ERR_ARG;
            BandPtrs bp;
            for (int i = 0; i < 4; ++i) {
                bp.p[i] = band[i];
                bp.s[i] = s4(band_s + 4 * i);
            }
            int64_t total = (int64_t)B * 2 * c * H * W;
            int grid = stream_grid(total);
            if (grid > kSMs * 4) grid = kSMs * 4;  // fewer atomics
            if (gos.c == 1)
                merge_bwd_alpha<T, true><<<grid, 256, 0, st>>>((const T*)gout, gos, bp, galpha_w, c, H, W, h, w, total);
            else
                merge_bwd_alpha<T, false><<<grid, 256, 0, st>>>((const T*)gout, gos, bp, galpha_w, c, H, W, h, w, total);
        }
    });
    note_launches((gb ? 1 : 0) + (gband ? 1 : 0) + (galpha_w ? 1 : 0));
    return check_launch();
}

extern "C" int el_gated_residual_fwd(const void* b, const int64_t bs_[4], const void* y, const int64_t ys_[4], const float* gamma, void* out,
                                     const int64_t os_[4], void* out2, const int64_t os2_[4], int B, int C, int H, int W, int dtype, void* stream) {
    if (!b || !y || !gamma || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0 || (out2 && !os2_)) return EL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Strides4 bs = s4(bs_), ys = s4(ys_), os = s4(os_), os2 = out2 ? s4(os2_) : Strides4{0, 0, 0, 0};
    EL_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec16<T>::N;
        if (out2 && !(channel_vectorisable<T>(b, bs, C) && channel_vectorisable<T>(y, ys, C) && channel_vectorisable<T>(out, os, C) &&
                      channel_vectorisable<T>(out2, os2, C) && C / V <= 256 && B <= 65535))
            return EL_ERR_UNSUPPORTED;  // the second destination is an engine-only (NHWC) feature
        if (channel_vectorisable<T>(b, bs, C) && channel_vectorisable<T>(y, ys, C) && channel_vectorisable<T>(out, os, C)) {
            if (C / V <= 256 && B <= 65535) {
                int cpb, rows;
                dim3 g = tile_grid(C / V, W, H, B, cpb, rows);
                gated_tiled<T><<<g, 256, 0, st>>>((const T*)b, bs, (const T*)y, ys, gamma, (T*)out, os, (T*)out2, os2, C / V, cpb, rows, H, W);
            } else {
                int64_t total = (int64_t)B * H * W * (C / V);
                gated_cvec<T><<<stream_grid(total), 256, 0, st>>>((const T*)b, bs, (const T*)y, ys, gamma, (T*)out, os, C / V, H, W, total);
            }
        } else {
            int64_t total = (int64_t)B * C * H * W;
            if (os.c == 1)
                gated_generic<T, true><<<stream_grid(total), 256, 0, st>>>((const T*)b, bs, (const T*)y, ys, gamma, (T*)out, os, C, H, W, total);
            else
                gated_generic<T, false><<<stream_grid(total), 256, 0, st>>>((const T*)b, bs, (const T*)y, ys, gamma, (T*)out, os, C, H, W, total);
        }
    });
    note_launches(1);
    return check_launch();
}

extern "C" int el_ingest_u8(const uint8_t* src, void* dst, const int64_t ds_[4], int B, int H, int W, int dtype, void* stream) {
    if (!src || !dst || B <= 0 || H <= 0 || W <= 0) return EL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Strides4 ds = s4(ds_);
    int64_t total = (int64_t)B * H * W;
    EL_DISPATCH_DTYPE(dtype, { ingest_u8_kernel<T><<<stream_grid(total), 256, 0, st>>>(src, (T*)dst, ds, H, W, total); });
    note_launches(1);
    return check_launch();
}

extern "C" int el_gated_residual_bwd(const void* g, const int64_t gs_[4], const void* y, const int64_t ys_[4], const float* gamma, void* gy,
                                     const int64_t os_[4], float* ggamma, int B, int C, int H, int W, int dtype, void* stream) {
    if (!g || !y || !gamma || !gy || !ggamma || B <= 0 || C <= 0 || H <= 0 || W <= 0) return EL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Strides4 gs = s4(gs_), ys = s4(ys_), os = s4(os_);
    const int64_t total = (int64_t)B * C * H * W;
    int grid = stream_grid(total);
    if (grid > kSMs * 4) grid = kSMs * 4;  // fewer atomics on the scalar
    EL_DISPATCH_DTYPE(dtype, {
        if (os.c == 1)
            gated_bwd_generic<T, true><<<grid, 256, 0, st>>>((const T*)g, gs, (const T*)y, ys, gamma, (T*)gy, os, ggamma, C, H, W, total);
        else
            gated_bwd_generic<T, false><<<grid, 256, 0, st>>>((const T*)g, gs, (const T*)y, ys, gamma, (T*)gy, os, ggamma, C, H, W, total);
    });
    note_launches(1);
    return check_launch();
}
