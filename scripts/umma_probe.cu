// Probe: may a K-major SWIZZLE_128B UMMA operand start at a row that is NOT a multiple of 8 (i.e. inside a
// 1024-byte swizzle atom)?  D[128 x N] = A[128 x 64] * B[rows r0 .. r0+N of a taller smem tile]^T for r0 = 0..9,
// with the descriptor's base_offset field either 0 or (start_addr >> 7) & 7.  Prints the max error of each
// variant against a CPU reference.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../object-tracking_b200/csrc/ptx.cuh"
using namespace b2t;

constexpr int N = 128, ROWS_B = 160;
#ifndef ROWB
#define ROWB 128          // bytes per K-major row: 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
#endif
constexpr int KEL = ROWB / 2;   // K elements per row

__device__ __forceinline__ uint64_t desc_with_base(uint32_t saddr, uint32_t base_off) {
    return umma_desc_kmajor(saddr, ROWB) | (uint64_t(base_off & 7) << 49);
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                       const __grid_constant__ CUtensorMap tmB, float *out, int r0,
                                                       int use_base) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sA = smem, *sB = smem + 128 * ROWB;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sB + ROWS_B * ROWB);
    uint64_t *done = bar + 1;
    uint32_t *slot = reinterpret_cast<uint32_t *>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<128>(slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tacc = *slot;
    if (warp == 0 && elect_one()) {
        mbar_expect_tx(bar, 128 * ROWB + ROWS_B * ROWB);
        tma_load_2d(&tmA, bar, sA, 0, 0, kEvictNormal);
        tma_load_2d(&tmB, bar, sB, 0, 0, kEvictNormal);
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_f16(128, N);
        for (int k = 0; k < ROWB / 32; ++k) {
            const uint32_t a = smem_u32(sA) + k * 32, b = smem_u32(sB) + r0 * ROWB + k * 32;
            const uint64_t da = umma_desc_kmajor(a, ROWB);
            const uint64_t db = desc_with_base(b, use_base ? ((b >> 7) & 7) : 0);
            umma_f16(tacc, da, db, idesc, k ? 1u : 0u);
        }
        umma_commit(done);
    }
    __syncwarp();
    mbar_wait(done, 0);
    tc_fence_after();
    for (int j = 0; j < N / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(tacc + (uint32_t(warp * 32) << 16) + j * 32, v);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * N + j * 32 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tacc);
}

int main() {
    std::vector<__half> A(128 * KEL), B(ROWS_B * KEL);
    srand(1);
    for (auto &x : A) x = __float2half((rand() % 17 - 8) / 8.f);
    for (auto &x : B) x = __float2half((rand() % 17 - 8) / 8.f);
    __half *dA, *dB;
    float *dO;
    cudaMalloc(&dA, A.size() * 2);
    cudaMalloc(&dB, B.size() * 2);
    cudaMalloc(&dO, 128 * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    CUtensorMap tA, tB;
    cuuint32_t es[2] = {1, 1};
    {
        cuuint64_t d[2] = {KEL, 128}, s[1] = {ROWB};
        cuuint32_t b[2] = {KEL, 128};
        enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            (ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    {
        cuuint64_t d[2] = {KEL, ROWS_B}, s[1] = {ROWB};
        cuuint32_t b[2] = {KEL, ROWS_B};
        CUresult r = enc(&tB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         (ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode B failed %d\n", (int)r); return 1; }
    }
    const int smem = 128 * ROWB + ROWS_B * ROWB + 1024 + 64;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> O(128 * N);
    for (int use_base = 0; use_base < 2; ++use_base)
        for (int r0 = 0; r0 <= 18; ++r0) {
            if (r0 > 9 && r0 < 16) continue;
            probe_kernel<<<1, 128, smem>>>(tA, tB, dO, r0, use_base);
            cudaError_t e = cudaDeviceSynchronize();
            if (e) { printf("r0=%d base=%d: CUDA error %s\n", r0, use_base, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
            double worst = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int k = 0; k < KEL; ++k) ref += (double)__half2float(A[m * KEL + k]) * __half2float(B[(r0 + n) * KEL + k]);
                    worst = fmax(worst, fabs(ref - O[m * N + n]));
                }
            printf("row_bytes=%d base_offset=%s r0=%2d  max|err| = %g  %s\n", ROWB, use_base ? "addr" : "0   ", r0, worst, worst < 1e-3 ? "OK" : "WRONG");
        }
    return 0;
}
