// Quality Focal Loss and Distribution Focal Loss, forward and backward, one pass each.
// Replaces quality_focal_loss (utils/loss.py:22-70) and DFLoss.__call__ (loss.py:209-224) /
// distribution_focal_loss (loss.py:88-137) of the reference.  Elementwise / per-row streaming
// kernels: HBM-bound.
#include "el_common.cuh"

namespace el {

constexpr int kQflThreads = 256;
constexpr int kQflPerThread = 4;

// sigmoid and softplus share one exponential: e = exp(-|x|) in (0, 1];  sigmoid = 1/(1+e) (x >= 0) or e/(1+e);  softplus(-|x|) = log1p(e)
__device__ __forceinline__ void qfl_terms(float x, float t, float beta, float& p, float& bce, float& diff, float& scale) {
    const float e = expf(-fabsf(x));
    const float inv = 1.f / (1.f + e);
    p = x >= 0.f ? inv : e * inv;
    bce = fmaxf(x, 0.f) - x * t + log1pf(e);                // binary_cross_entropy_with_logits
    diff = t > 0.f ? fabsf(t - p) : p;                      // loss.py:57-61
    scale = beta == 2.f ? diff * diff : powf(diff, beta);
}

// four consecutive logits of type T as one 16-byte (fp32) or 8-byte (16-bit) access
template <typename T> __device__ __forceinline__ void load4(const T* p, float (&f)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&f)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[4]) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
}
template <> __device__ __forceinline__ void load4<__half>(const __half* p, float (&f)[4]) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
template <typename T> __device__ __forceinline__ void store4(T* p, const float (&f)[4]);
template <> __device__ __forceinline__ void store4<float>(float* p, const float (&f)[4]) { *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]); }
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}
template <> __device__ __forceinline__ void store4<__half>(__half* p, const float (&f)[4]) {
    __half2 a = __floats2half2_rn(f[0], f[1]), b = __floats2half2_rn(f[2], f[3]);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

// VEC: every pointer is 16-byte aligned -> groups of four elements per thread and iteration (128-bit target / loss accesses), the
// n % 4 tail goes through the scalar loop.
template <typename T, bool VEC>
__global__ void __launch_bounds__(kQflThreads) qfl_fwd_kernel(const T* __restrict__ pred, const float* __restrict__ target, int64_t n, float beta,
                                                              float* __restrict__ loss, float* __restrict__ partials) {
    float acc = 0.f;
    const int64_t tid0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    int64_t done = 0;
    if constexpr (VEC) {
        const int64_t n4 = n >> 2;
        for (int64_t q = tid0; q < n4; q += nthr) {
            float x[4], l[4];
            load4<T>(pred + 4 * q, x);
            const float4 t = *reinterpret_cast<const float4*>(target + 4 * q);
            const float tt[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float p, bce, diff, scale;
                qfl_terms(x[k], tt[k], beta, p, bce, diff, scale);
                l[k] = bce * scale;
            }
            if (loss) *reinterpret_cast<float4*>(loss + 4 * q) = make_float4(l[0], l[1], l[2], l[3]);
            acc += (l[0] + l[1]) + (l[2] + l[3]);
        }
        done = n4 << 2;
    }
    for (int64_t i = done + tid0; i < n; i += nthr) {
        float p, bce, diff, scale;
        qfl_terms(to_f(pred[i]), target[i], beta, p, bce, diff, scale);
        float l = bce * scale;
        if (loss) loss[i] = l;
        acc += l;
    }
    if (partials) {
        __shared__ float red[kQflThreads / 32];
        float v = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int k = 0; k < kQflThreads / 32; ++k) s += red[k];
            partials[blockIdx.x] = s;
        }
    }
}

__global__ void __launch_bounds__(256) sum_partials(const float* __restrict__ partials, int n, float* __restrict__ out) {
    __shared__ float red[8];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) acc += partials[i];  // fixed order: deterministic
    float v = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int k = 0; k < 8; ++k) s += red[k];
        *out = s;
    }
}

__device__ __forceinline__ float qfl_grad(float x, float t, float beta) {
    float p, bce, diff, scale;
    qfl_terms(x, t, beta, p, bce, diff, scale);
    // d scale / dx: the modulating factor is not detached in the reference (autograd flows through it)
    float sgn = 1.f;
    if (t > 0.f) { float d = t - p; sgn = d > 0.f ? -1.f : (d < 0.f ? 1.f : 0.f); }
    float dpow = beta == 2.f ? 2.f * diff : beta * powf(diff, beta - 1.f);
    return (p - t) * scale + bce * dpow * sgn * p * (1.f - p);
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(kQflThreads) qfl_bwd_kernel(const T* __restrict__ pred, const float* __restrict__ target, int64_t n, float beta,
                                                              const float* __restrict__ gout, const float* __restrict__ gscalar, T* __restrict__ gpred) {
    const float gs = gscalar ? __ldg(gscalar) : 1.f;
    const int64_t tid0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    int64_t done = 0;
    if constexpr (VEC) {
        const int64_t n4 = n >> 2;
        for (int64_t q = tid0; q < n4; q += nthr) {
            float x[4], g[4];
            load4<T>(pred + 4 * q, x);
            const float4 t = *reinterpret_cast<const float4*>(target + 4 * q);
            float4 go = make_float4(gs, gs, gs, gs);
            if (gout) go = *reinterpret_cast<const float4*>(gout + 4 * q);
            g[0] = qfl_grad(x[0], t.x, beta) * go.x; g[1] = qfl_grad(x[1], t.y, beta) * go.y;
            g[2] = qfl_grad(x[2], t.z, beta) * go.z; g[3] = qfl_grad(x[3], t.w, beta) * go.w;
            store4<T>(gpred + 4 * q, g);
        }
        done = n4 << 2;
    }
    for (int64_t i = done + tid0; i < n; i += nthr) gpred[i] = from_f<T>(qfl_grad(to_f(pred[i]), target[i], beta) * (gout ? gout[i] : gs));
}

// ------------------------------------------------------------------------------------- DFL
// one thread per (row, side): 16 logits; 4 consecutive lanes form a row.  per_side = 0: DFLoss (mean over the 4 sides of a row, loss.py:209-224);
// per_side = 1: distribution_focal_loss (one value per side, loss.py:88-137)
template <typename T, bool BWD>
__global__ void __launch_bounds__(256) dfl_kernel(const T* __restrict__ pred, const float* __restrict__ target, int64_t sides, float* __restrict__ loss,
                                                  const float* __restrict__ gout, T* __restrict__ gpred, int per_side) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool valid = i < sides;
    float lg[16];
    float l = 0.f;
    if (valid) {
        const T* p = pred + i * 16;
        if constexpr (sizeof(T) == 4) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                float f[4];
                unpack<T>(ldg_stream(p + 4 * v), f);
#pragma unroll
                for (int e = 0; e < 4; ++e) lg[4 * v + e] = f[e];
            }
        } else {
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                float f[8];
                unpack<T>(ldg_stream(p + 8 * v), f);
#pragma unroll
                for (int e = 0; e < 8; ++e) lg[8 * v + e] = f[e];
            }
        }
        float t = fminf(fmaxf(target[i], 0.f), 14.99f);  // clamp_(0, reg_max - 1 - 0.01), loss.py:216
        int tl = (int)t;                                  // target.long()
        float wl = (float)(tl + 1) - t, wr = 1.f - wl;
        float m = lg[0];
#pragma unroll
        for (int k = 1; k < 16; ++k) m = fmaxf(m, lg[k]);
        float s = 0.f, xl = 0.f, xr = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            xl = k == tl ? lg[k] : xl;
            xr = k == tl + 1 ? lg[k] : xr;
            lg[k] = expf(lg[k] - m);
            s += lg[k];
        }
        float lse = m + logf(s);
        l = (lse - xl) * wl + (lse - xr) * wr;  // CE(pred, tl)*wl + CE(pred, tr)*wr
        if (BWD) {
            float g = per_side ? __ldg(gout + i) : __ldg(gout + (i >> 2)) * 0.25f;  // .mean(-1) over the 4 sides
            float inv = 1.f / s;
            float o[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) o[k] = g * (lg[k] * inv - (k == tl ? wl : 0.f) - (k == tl + 1 ? wr : 0.f));
            T* q = gpred + i * 16;
            if constexpr (sizeof(T) == 4) {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    float f[4] = {o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]};
                    *reinterpret_cast<uint4*>(q + 4 * v) = pack<T>(f);
                }
            } else {
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = o[8 * v + e];
                    *reinterpret_cast<uint4*>(q + 8 * v) = pack<T>(f);
                }
            }
        }
    }
    if (!BWD) {
        if (per_side) {
            if (valid) loss[i] = l;
        } else {
            l += __shfl_xor_sync(0xffffffffu, l, 1);
            l += __shfl_xor_sync(0xffffffffu, l, 2);
            if (valid && (threadIdx.x & 3) == 0) loss[i >> 2] = l * 0.25f;
        }
    }
}

static inline int qfl_grid(int64_t n) {
    int64_t need = ceil_div(n, (int64_t)kQflThreads * kQflPerThread);
    int64_t cap = (int64_t)kSMs * 8;
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

}  // namespace el

using namespace el;

extern "C" int el_qfl_partials(int64_t n) { return qfl_grid(n); }

extern "C" int el_qfl_fwd(const void* pred, const float* target, int64_t n, float beta, int dtype, float* loss, float* loss_sum, float* partials,
                          void* stream) {
    if (!pred || !target || n <= 0 || (!loss && !loss_sum) || (loss_sum && !partials)) return EL_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = qfl_grid(n);
    const bool vec = aligned16(pred) && aligned16(target) && (!loss || aligned16(loss));
    EL_DISPATCH_DTYPE(dtype, {
        if (vec) qfl_fwd_kernel<T, true><<<grid, kQflThreads, 0, s>>>((const T*)pred, target, n, beta, loss, loss_sum ? partials : nullptr);
        else qfl_fwd_kernel<T, false><<<grid, kQflThreads, 0, s>>>((const T*)pred, target, n, beta, loss, loss_sum ? partials : nullptr);
    });
    if (loss_sum) sum_partials<<<1, 256, 0, s>>>(partials, grid, loss_sum);
    note_launches(loss_sum ? 2 : 1);
    return check_launch();
}

extern "C" int el_qfl_bwd(const void* pred, const float* target, int64_t n, float beta, int dtype, const float* gout, const float* gscalar, void* gpred,
                          void* stream) {
    if (!pred || !target || !gpred || n <= 0 || ((gout == nullptr) == (gscalar == nullptr))) return EL_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = aligned16(pred) && aligned16(target) && aligned16(gpred) && (!gout || aligned16(gout));
    EL_DISPATCH_DTYPE(dtype, {
        if (vec) qfl_bwd_kernel<T, true><<<qfl_grid(n), kQflThreads, 0, s>>>((const T*)pred, target, n, beta, gout, gscalar, (T*)gpred);
        else qfl_bwd_kernel<T, false><<<qfl_grid(n), kQflThreads, 0, s>>>((const T*)pred, target, n, beta, gout, gscalar, (T*)gpred);
    });
    note_launches(1);
    return check_launch();
}

extern "C" int el_dfl_fwd(const void* pred, const float* target, int64_t rows, int dtype, float* loss, void* stream) {
    if (!pred || !target || !loss || rows <= 0 || !aligned16(pred)) return EL_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t sides = rows * 4;
    EL_DISPATCH_DTYPE(dtype, { dfl_kernel<T, false><<<(unsigned)ceil_div(sides, 256), 256, 0, s>>>((const T*)pred, target, sides, loss, nullptr, nullptr, 0); });
    note_launches(1);
    return check_launch();
}

extern "C" int el_dfl_side_fwd(const void* pred, const float* target, int64_t sides, int dtype, float* loss, void* stream) {
    if (!pred || !target || !loss || sides <= 0 || !aligned16(pred)) return EL_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    EL_DISPATCH_DTYPE(dtype, { dfl_kernel<T, false><<<(unsigned)ceil_div(sides, 256), 256, 0, s>>>((const T*)pred, target, sides, loss, nullptr, nullptr, 1); });
    note_launches(1);
    return check_launch();
}

extern "C" int el_dfl_side_bwd(const void* pred, const float* target, int64_t sides, int dtype, const float* gout, void* gpred, void* stream) {
    if (!pred || !target || !gout || !gpred || sides <= 0 || !aligned16(pred) || !aligned16(gpred)) return EL_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    EL_DISPATCH_DTYPE(dtype, { dfl_kernel<T, true><<<(unsigned)ceil_div(sides, 256), 256, 0, s>>>((const T*)pred, target, sides, nullptr, gout, (T*)gpred, 1); });
    note_launches(1);
    return check_launch();
}

extern "C" int el_dfl_bwd(const void* pred, const float* target, int64_t rows, int dtype, const float* gout, void* gpred, void* stream) {
    if (!pred || !target || !gout || !gpred || rows <= 0 || !aligned16(pred) || !aligned16(gpred)) return EL_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t sides = rows * 4;
    EL_DISPATCH_DTYPE(dtype, { dfl_kernel<T, true><<<(unsigned)ceil_div(sides, 256), 256, 0, s>>>((const T*)pred, target, sides, nullptr, gout, (T*)gpred, 0); });
    note_launches(1);
    return check_launch();
}
