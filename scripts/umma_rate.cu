// Probe: issue rate of tcgen05.mma (kind::f16, fp32 accumulate) as the conv kernels use it -- three MMAs per 16-deep
// K step (w_lo*x_hi, w_hi*x_lo -> correction accumulator; w_hi*x_hi -> main accumulator), A = [128 x 64] K-major
// weight tile, B = N rows of a taller K-major activation patch whose start row is shifted per tap (kh*P + kw).
// Reports clocks per MMA for: N, aligned vs tap-shifted B starts, SWIZZLE_128B vs 64B rows, cta_group::1 vs ::2.
// Operand values are whatever the fill loop wrote; only timing matters.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate umma_rate.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "../object-tracking_b200/csrc/ptx.cuh"
using namespace b2t;

struct Args {
    int N;          // instruction N (cta_group::2: both CTAs together)
    int P;          // patch pitch in rows; P == 0 -> every tap starts at row 0 (aligned)
    int row_bytes;  // 128 or 64
    int iters;      // (tap, chunk) iterations of 9 taps each
    int mode;       // 0: 3 MMAs/kstep as shipped; 1: main MMA only; 2: A operand row shift instead of B
    int feat;       // loop structure of the real kernel: 1 = whole warp loops, per tap wait(done barrier)+fence+syncwarp;
                    // 2 = 3-stage full/empty ring with a producer warp and a tcgen05.commit per tap; 4 = 8 warps spin on a barrier
    long long *out;
};

__device__ __forceinline__ void mma2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void commit2(uint64_t *bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CG>
__global__ void __launch_bounds__(320, 1) rate_kernel(const Args a) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    // layout: A hi (16 KB) | A lo (16 KB) | B hi (64 KB) | B lo (64 KB) | barrier
    uint8_t *sA = smem, *sB = smem + 32768;
    uint64_t *done = reinterpret_cast<uint64_t *>(smem + 32768 + 131072);
    uint64_t *ready = done + 1, *spin = done + 2, *w_full = done + 3, *w_empty = done + 6;
    uint32_t *slot = reinterpret_cast<uint32_t *>(done + 9);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (32768 + 131072) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(smem)[i] = (i * 2654435761u) & 0x03ff03ffu;   // small positive fp16 values
    if (threadIdx.x == 0) {
        mbar_init(done, 1); mbar_init(ready, 1); mbar_init(spin, 1);
        for (int i = 0; i < 3; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        fence_barrier_init();
        mbar_arrive(ready);
    }
    fence_proxy_async();
    if (warp == 0) {
        if (CG == 1) tmem_alloc<512>(slot);
        else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *slot;
    const bool leader = CG == 1 || cluster_rank() == 0;
    long long t0 = 0, t1 = 0;
    const int lane = threadIdx.x & 31;
    if (warp == 1) {
        const int N = a.N;
        const uint32_t idesc = umma_idesc_f16(CG == 2 ? 256 : 128, N), dhi = umma_desc_hi(a.row_bytes);
        const uint32_t wh = umma_desc_lo(smem_u32(sA)), wl = wh + (16384 >> 4);
        const uint32_t xh = umma_desc_lo(smem_u32(sB)), xl = xh + (65536 >> 4);
        const uint32_t t_main = tmem, t_corr = tmem + N;
        const int ksteps = a.row_bytes / 32;
        const bool warp_loop = a.feat & 3;
        const bool issuer = leader && (warp_loop ? lane == 0 : elect_one());
        if (issuer) t0 = clock64();
        if (issuer || (warp_loop && leader)) {
            int ws = 0; uint32_t wphase = 0;
            for (int it = 0; it < a.iters; ++it)
                for (int tap = 0; tap < 9; ++tap) {
                    if (a.feat & 2) { mbar_wait(&w_full[ws], wphase); tc_fence_after(); }
                    else if (a.feat & 1) { mbar_wait(ready, 0); tc_fence_after(); }
                    if (issuer) {
                        const int kh = tap / 3, kw = tap - kh * 3;
                        const uint32_t s16 = a.P ? ((kh * a.P + kw) * a.row_bytes) >> 4 : 0;
                        const uint32_t sa = a.mode == 2 ? s16 : 0, sb = a.mode == 2 ? 0 : s16;
                        for (int k = 0; k < ksteps; ++k) {
                            const uint64_t dwh = umma_desc_make(wh + sa + 2 * k, dhi), dwl = umma_desc_make(wl + sa + 2 * k, dhi);
                            const uint64_t dxh = umma_desc_make(xh + sb + 2 * k, dhi), dxl = umma_desc_make(xl + sb + 2 * k, dhi);
                            const uint32_t acc = (it | tap | k) ? 1u : 0u;
                            if (CG == 1) {
                                if (a.mode != 1) { umma_f16(t_corr, dwl, dxh, idesc, acc); umma_f16(t_corr, dwh, dxl, idesc, 1u); }
                                umma_f16(t_main, dwh, dxh, idesc, acc);
                            } else {
                                if (a.mode != 1) { mma2(t_corr, dwl, dxh, idesc, acc); mma2(t_corr, dwh, dxl, idesc, 1u); }
                                mma2(t_main, dwh, dxh, idesc, acc);
                            }
                        }
                        if (a.feat & 2) umma_commit(&w_empty[ws]);
                    }
                    if (warp_loop) __syncwarp();
                    if (++ws == 3) { ws = 0; wphase ^= 1; }
                }
            if (issuer) { if (CG == 1) umma_commit(done); else commit2(done); }
        }
        __syncwarp();
        mbar_wait(done, 0);
        t1 = clock64();
        tc_fence_after();
        long long t0e = t0;                       // t0 lives in the issuing lane only
        for (int o = 16; o; o >>= 1) t0e = max(t0e, __shfl_xor_sync(0xffffffffu, t0e, o));
        if (lane == 0) a.out[blockIdx.x] = leader ? t1 - t0e : 0;
        if (lane == 0) mbar_arrive(spin);       // releases the spinning warps
    } else if (warp == 0) {
        if ((a.feat & 2) && leader && elect_one()) {     // "producer": refills a stage as soon as it is empty (no TMA)
            int ws = 0; uint32_t wphase = 0;
            for (int i = 0; i < a.iters * 9; ++i) {
                mbar_wait(&w_empty[ws], wphase ^ 1);
                mbar_arrive(&w_full[ws]);
                if (++ws == 3) { ws = 0; wphase ^= 1; }
            }
        }
    } else if (a.feat & 4) {
        mbar_wait(spin, 0);                             // 8 warps polling an mbarrier, like the waiting epilogue warps
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) {
        if (CG == 1) tmem_dealloc<512>(tmem);
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

template <int CG>
static void run(const char *tag, Args a, int n_cta) {
    const int smem = 32768 + 131072 + 128 + 1024;
    cudaFuncSetAttribute(rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_cta);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<CG>, a);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: CUDA error %s\n", tag, cudaGetErrorString(e)); exit(1); }
    }
    std::vector<long long> h(n_cta);
    cudaMemcpy(h.data(), a.out, n_cta * sizeof(long long), cudaMemcpyDeviceToHost);
    std::vector<long long> v;
    for (auto x : h) if (x > 0) v.push_back(x);
    std::sort(v.begin(), v.end());
    const double n_mma = (double)a.iters * 9 * (a.row_bytes / 32) * (a.mode == 1 ? 1 : 3);
    const double floor = 128.0 * a.N / 256.0 / 1.0;   // clocks per MMA per SM at the nominal rate (both cta_groups)
    printf("%-34s cg=%d N=%3d P=%2d row=%3d mode=%d feat=%d: clk/MMA min %.1f med %.1f max %.1f  (floor %.0f, eff %.2f)\n", tag, CG,
           a.N, a.P, a.row_bytes, a.mode, a.feat, v.front() / n_mma, v[v.size() / 2] / n_mma, v.back() / n_mma, floor,
           floor / (v[v.size() / 2] / n_mma));
}

int main() {
    long long *d;
    cudaMalloc(&d, 1024 * sizeof(long long));
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    const int it = 40;
    for (int n_cta : {1, n_sm}) {
        printf("---- %d CTAs\n", n_cta);
        for (int feat : {0, 1, 2, 3, 4, 5, 6, 7}) {
            run<1>("N=192 P=14", Args{192, 14, 128, it, 0, feat, d}, n_cta);
            run<1>("N=128 P=30 64B rows", Args{128, 30, 64, it, 0, feat, d}, n_cta);
        }
        run<1>("N=64", Args{64, 14, 128, it, 0, 0, d}, n_cta);
        run<1>("N=256", Args{256, 14, 128, it, 0, 0, d}, n_cta);
        run<1>("main only N=192", Args{192, 14, 128, it, 1, 0, d}, n_cta);
        if (n_cta >= 2) {
            run<2>("2-CTA N=208", Args{208, 14, 128, it, 0, 0, d}, n_cta & ~1);
            run<2>("2-CTA N=256", Args{256, 14, 128, it, 0, 0, d}, n_cta & ~1);
        }
    }
    return 0;
}
