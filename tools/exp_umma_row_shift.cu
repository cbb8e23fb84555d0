// Experiment for DESIGN.md section 9 item 2b (wide 3x3 convolutions from ONE haloed shared-memory tile): can a tcgen05.mma A operand start
// at an arbitrary ROW of a 128-byte-swizzled K-major tile, and with which descriptor encoding?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/exp_umma_row_shift tools/exp_umma_row_shift.cu && /tmp/exp_umma_row_shift
//
// Set-up: A_full = 256 rows x 64 bf16 channels (row = pixel) written to shared memory the way a SWIZZLE_128B TMA box lands (16-byte chunk j
// of row r at chunk j ^ (r & 7), tile 1 KiB aligned); B = 64 x 64 weights, same layout.  For a row shift s and a group pitch P (rows between
// consecutive 8-row groups: 8 = dense, 16 = an image-row pitch of 16 pixels with an 8-pixel-wide output tile) the expected result is
//     D[m][n] = sum_k A_full[s + (m / 8) * P + (m % 8)][k] * B[n][k],        m < 128, n < 64.
// Three descriptor encodings are tried per (s, P):
//     mode 0: start address = tile + s * 128, base_offset = 0
//     mode 1: start address = tile + s * 128, base_offset = (start >> 7) & 7      (CUTLASS's rule for starts that are not 1 KiB aligned)
//     mode 2: start address = tile + (s & ~7) * 128, base_offset = s & 7
// and the program prints which of them reproduce the expected matrix exactly (small integers: exact in bf16 / fp32).
// Not part of libedgeline_b200.so; nothing in the product depends on it.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kRows = 256, kK = 64, kN = 64, kM = 128;

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
// K-major, 128-byte swizzle: start>>4 [0,14), LBO>>4 [16,30) (unused), SBO>>4 [32,46) = bytes between 8-row groups, version 1 [46,48),
// base_offset [49,52), layout 2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_offset) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | ((uint64_t)(base_offset & 7) << 49) | (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc_bf16(int M, int N) {  // D = f32, A / B = bf16, both K-major
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// one CTA of 128 threads; out[m][n] fp32
__global__ void __launch_bounds__(128) row_shift_kernel(const __nv_bfloat16* __restrict__ a_full, const __nv_bfloat16* __restrict__ b, int shift,
                                                        int pitch_rows, int mode, float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char sm_raw[];
    const uint32_t sbase = (smem_addr(sm_raw) + 1023u) & ~1023u;
    unsigned char* sm = sm_raw + (sbase - smem_addr(sm_raw));
    unsigned char* sA = sm;                       // kRows x 128 B
    unsigned char* sB = sm + kRows * 128;         // kN x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + kRows * 128 + kN * 128);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    // generic-proxy writes in the TMA SWIZZLE_128B pattern: chunk j of row r -> chunk j ^ (r & 7)
    for (int i = tid; i < kRows * 8; i += 128) {
        const int r = i >> 3, j = i & 7;
        *reinterpret_cast<uint4*>(sA + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(a_full + r * kK + j * 8);
    }
    for (int i = tid; i < kN * 8; i += 128) {
        const int r = i >> 3, j = i & 7;
        *reinterpret_cast<uint4*>(sB + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(b + r * kK + j * 8);
    }
    if (tid == 0) {
        mbar_init(smem_addr(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(s_tmem)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to the tensor core's async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;
    if (tid == 0) {
        const uint32_t a_tile = sbase, b_tile = sbase + kRows * 128;
        const uint32_t sbo = (uint32_t)pitch_rows * 128;
        uint32_t start = a_tile + (uint32_t)shift * 128, bo = 0;
        if (mode == 1) bo = (start >> 7) & 7;
        if (mode == 2) { start = a_tile + (uint32_t)(shift & ~7) * 128; bo = (uint32_t)shift & 7; }
        const uint32_t idesc = idesc_bf16(kM, kN);
        for (int ks = 0; ks < kK / 16; ++ks)  // 16 channels = 32 bytes along K inside the swizzle atom
            umma(tmem, desc_sw128(start + 32 * ks, sbo, bo), desc_sw128(b_tile + 32 * ks, 1024, 0), idesc, ks > 0 ? 1u : 0u);
        umma_commit(smem_addr(bar));
    }
    mbar_wait(smem_addr(bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = tid;  // warp w reads TMEM lanes 32w .. 32w + 31
    for (int c0 = 0; c0 < kN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int e = 0; e < 16; ++e) out[row * kN + c0 + e] = __uint_as_float(v[e]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

int main() {
    std::vector<__nv_bfloat16> hA(kRows * kK), hB(kN * kK);
    std::vector<float> fA(kRows * kK), fB(kN * kK);
    for (int r = 0; r < kRows; ++r)
        for (int k = 0; k < kK; ++k) { fA[r * kK + k] = (float)((r * 3 + k * 5) % 7 - 3); hA[r * kK + k] = __float2bfloat16(fA[r * kK + k]); }
    for (int n = 0; n < kN; ++n)
        for (int k = 0; k < kK; ++k) { fB[n * kK + k] = (float)((n * 2 + k) % 5 - 2); hB[n * kK + k] = __float2bfloat16(fB[n * kK + k]); }
    __nv_bfloat16 *dA, *dB;
    float* dOut;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dOut, kM * kN * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = kRows * 128 + kN * 128 + 64 + 1024;
    cudaFuncSetAttribute(row_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<float> got(kM * kN);
    const int shifts[] = {0, 1, 2, 3, 7, 8, 9, 17, 18, 34};
    for (int pitch : {8, 16}) {
        for (int s : shifts) {
            if (s + 15 * pitch + 8 > kRows) continue;
            printf("pitch %2d rows, shift %2d:", pitch, s);
            for (int mode = 0; mode < 3; ++mode) {
                cudaMemset(dOut, 0xff, kM * kN * 4);
                row_shift_kernel<<<1, 128, smem>>>(dA, dB, s, pitch, mode, dOut);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("  mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(got.data(), dOut, kM * kN * 4, cudaMemcpyDeviceToHost);
                int bad = 0;
                for (int m = 0; m < kM; ++m) {
                    const int r = s + (m / 8) * pitch + (m % 8);
                    for (int n = 0; n < kN; ++n) {
                        float want = 0.f;
                        for (int k = 0; k < kK; ++k) want += fA[r * kK + k] * fB[n * kK + k];
                        bad += got[m * kN + n] != want;
                    }
                }
                printf("  mode %d %s", mode, bad ? "MISMATCH" : "ok");
                if (bad) printf("(%d)", bad);
            }
            printf("\n");
        }
    }
    return 0;
}
