// Shared device/host helpers for libedgeline_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/edgeline_b200.h"

namespace el {

extern thread_local int g_last_cuda_error;
extern unsigned long long g_launches;  // kernels enqueued by this library (measurement aid, see el_launch_count)
inline void note_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

inline int check_launch() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_cuda_error = (int)e;
        return EL_ERR_CUDA;
    }
    return EL_OK;
}

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- scalar conversions ---------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- 16-byte vectors of T <-> float registers -------------------------------------------------
template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

// Activation loads.  Neither uses the non-coherent path (ld.global.nc / __ldg): a kernel launched with programmatic dependent launch
// becomes resident while its producer is still WRITING the tensors it will read after griddepcontrol.wait, which breaks the
// "read-only for the lifetime of the kernel" contract of .nc -- measured: the engine graph and the eager path disagreed in the last bits
// of a few pixels until the loads were made coherent (EL_PDL=0 also cured it).
// streaming (read-once) 128-bit load, L2 only: re-used tiles of other tensors stay resident in L1
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// 128-bit load of data that neighbouring threads / CTAs of the same SM read again (bilinear taps, 2x upsampling): allocate in L1
__device__ __forceinline__ uint4 ldg_cached(const void* p) {
    uint4 r;
    asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_l2(const void* p) { return ldg_stream(p); }
__device__ __forceinline__ void stg_stream(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename T> __device__ __forceinline__ void unpack(uint4 raw, float (&f)[Vec16<T>::N]);
template <> __device__ __forceinline__ void unpack<float>(uint4 raw, float (&f)[4]) {
    f[0] = __uint_as_float(raw.x); f[1] = __uint_as_float(raw.y); f[2] = __uint_as_float(raw.z); f[3] = __uint_as_float(raw.w);
}
template <> __device__ __forceinline__ void unpack<__nv_bfloat16>(uint4 raw, float (&f)[8]) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
template <> __device__ __forceinline__ void unpack<__half>(uint4 raw, float (&f)[8]) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}

template <typename T> __device__ __forceinline__ uint4 pack(const float (&f)[Vec16<T>::N]);
template <> __device__ __forceinline__ uint4 pack<float>(const float (&f)[4]) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
}
template <> __device__ __forceinline__ uint4 pack<__nv_bfloat16>(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&t);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
template <> __device__ __forceinline__ uint4 pack<__half>(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __half2 t = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&t);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- packed fp32x2 FMA (Blackwell FFMA2): two IEEE fp32 FMAs per instruction on a 64-bit register pair.  The 3-register scalar
// FFMA issues every second cycle per scheduler, so FMA-bound CUDA-core kernels (depthwise 7x7, decode MLP) double their
// arithmetic rate by keeping adjacent channels as pairs.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack_f32x2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma_f32x2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// ---- warp helpers -------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

struct Strides4 { int64_t n, c, h, w; };
inline Strides4 s4(const int64_t* s) { return Strides4{s[0], s[1], s[2], s[3]}; }

// true when a (ptr, strides) view can be walked with 16-byte channel vectors
template <typename T> inline bool channel_vectorisable(const void* p, const Strides4& s, int C) {
    constexpr int V = 16 / sizeof(T);
    return s.c == 1 && (C % V) == 0 && aligned16(p) && (s.n % V) == 0 && (s.h % V) == 0 && (s.w % V) == 0;
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------------
// Kernels of this library that follow one another in a stream are launched with programmatic stream serialization: a kernel
// calls pdl_launch_dependents() at entry, so the NEXT kernel's CTAs may become resident and run their prologue (barrier init,
// TMEM allocation, weight staging) while this one is still computing; the next kernel calls pdl_wait() before it touches any
// activation memory, which blocks until this grid has completed and flushed.  A kernel that follows a non-PDL kernel (cuDNN,
// PyTorch) simply starts after it, as usual.  Captured by CUDA graphs as programmatic edges.  EL_PDL=0 disables the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

#define EL_DISPATCH_DTYPE(dtype, ...)                                              \
    switch (dtype) {                                                               \
        case EL_F32: { using T = float; __VA_ARGS__; } break;                      \
        case EL_F16: { using T = __half; __VA_ARGS__; } break;                     \
        case EL_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break;             \
        default: return EL_ERR_ARG;                                                \
    }

}  // namespace el
