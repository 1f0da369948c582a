// conv_halo_kernel -- the production convolution: implicit GEMM on tcgen05 with the activation patch held
// stationary in shared memory.
//
//   D[cout, pixel] = sum_{tap, c} Wt[cout, tap, c] * X[pixel + tap, c]
//
// * M = 128 output channels (A operand = one [128 x 64] weight tile per (tap, 64-channel chunk), K-major,
//   TMA 2-D box {64, 128} out of the packed [Cout][tap][Cin_pad] weights).
// * N = pixels (B operand).  Per 64-channel chunk ONE TMA 4-D box {64 ch, hP, h_rows, 1} brings the tile's
//   (hR x hC)-pixel patch plus its halo into shared memory, rows ordered x-fastest with pitch hP; image borders
//   are zero-filled by the TMA unit ('same' padding).  Output pixel n = r*hP + c needs, for tap (kh,kw), patch row
//   n + kh*hP + kw: a UNIFORM row shift, so every tap's B operand is the same patch addressed through a shared-
//   memory descriptor whose start is advanced by (kh*hP + kw)*128 bytes.  (UMMA's 128-byte swizzle is a pure
//   function of the address, scripts/umma_probe.cu verifies that operands may start at any 128-byte row.)
//   The patch is read from L2 once per chunk instead of once per tap: 9x less activation traffic, and the
//   kernel's L2->SM stream is then dominated by the weights, which a N~176-256 tile amortises to < 42 B/clk/SM.
// * precision: operands are fp16 (hi, lo) pairs; w_hi*x_hi goes to a main TMEM accumulator, w_lo*x_hi + w_hi*x_lo
//   to a separate correction accumulator (the tensor core truncates addends to the accumulator's exponent, so
//   the small terms must not share an accumulator with the big one); extra main accumulators are used
//   round-robin when TMEM has room (N <= 160).  The epilogue adds them in fp32 round-to-nearest.
// * epilogue: threads own one output channel each (TMEM lane), apply scale/bias (folded BN) + LeakyReLU and stage
//   the tile [pixel][channel] in the (now dead) operand buffers; then (pixel, 8-channel) items are written with
//   16-byte stores: optional 2x2 max-pool, hi/lo split, concat / space-to-depth addressing.
// * split-K over channel chunks (gridDim.z): raw fp32 partials, finished by splitk_epilogue_kernel.
#include "kernels.cuh"

namespace b2t {

constexpr int kWTileBytes = 128 * 64 * 2;            // one plane of a weight tile
constexpr int kWStageBytes = 2 * kWTileBytes;
constexpr int kHaloThreads = 192;

// Two resource shapes of the same kernel:
//  * big   -- one CTA per SM: two 64 KB patch buffers (double-buffered over channel chunks), 3 weight stages, all
//             512 TMEM columns (N up to 256).  For the long-K layers (26x26 and 13x13 grids, ConvLSTM).
//  * small -- two CTAs per SM: one 44 KB patch buffer, 2 weight stages, 256 TMEM columns (N <= 128).  For the
//             short-K, many-tile layers at the top of the network, where a CTA is mostly prologue + epilogue:
//             the co-resident CTA keeps the tensor core and the TMA unit busy meanwhile.
template <bool kSmall>
struct HaloCfg {
    static constexpr int kHaloBufs = kSmall ? 1 : 2;
    static constexpr int kHaloBufBytes = kSmall ? 45056 : 65536;      // hi + lo patch of one channel chunk
    static constexpr int kWStages = kSmall ? 2 : 3;
    static constexpr int kTmemCols = kSmall ? 256 : 512;
    static constexpr int kOperandBytes = kHaloBufs * kHaloBufBytes + kWStages * kWStageBytes;
    static constexpr int kSmemBytes = kOperandBytes + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*scale, bias*/;
    static constexpr int kMaxN = kSmall ? 128 : 256;
    static_assert(kMaxN * 132 * 4 <= kOperandBytes, "epilogue stage must fit in the operand buffers");
};

template <bool kSmall>
__global__ void __launch_bounds__(kHaloThreads, kSmall ? 2 : 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                 const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                 const ConvParams p) {
    using Cfg = HaloCfg<kSmall>;
    constexpr int kHaloBufs = Cfg::kHaloBufs, kHaloBufBytes = Cfg::kHaloBufBytes, kWStages = Cfg::kWStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *s_halo = smem;
    uint8_t *s_w = smem + kHaloBufs * kHaloBufBytes;
    uint8_t *tail = s_w + kWStages * kWStageBytes;
    uint64_t *halo_full = reinterpret_cast<uint64_t *>(tail);   // [2]
    uint64_t *halo_empty = halo_full + 2;                       // [2]
    uint64_t *w_full = halo_empty + 2;                          // [kWStages]
    uint64_t *w_empty = w_full + kWStages;                      // [kWStages]
    uint64_t *accum_bar = w_empty + kWStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);
    float *s_scale = reinterpret_cast<float *>(tail + 256);
    float *s_bias = s_scale + 128;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x;
    const int tx = t % p.h_tiles_x;  t /= p.h_tiles_x;
    const int ty = t % p.h_tiles_y;
    const int b = t / p.h_tiles_y;
    const int x0 = tx * p.hC, y0 = ty * p.hR;
    const int cout0 = blockIdx.y * 128;
    const int pad = p.ksize >> 1, taps = p.ksize * p.ksize;
    const int per = (p.cin_chunks + p.splits - 1) / p.splits;
    const int c_begin = blockIdx.z * per, c_end = min(p.cin_chunks, c_begin + per);
    const int n_chunks = c_end - c_begin;                      // host guarantees >= 1
    const int N = p.hN;
    const int n_main = max(1, min(min(p.n_main, Cfg::kTmemCols / N - 1), n_chunks * taps));

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX_hi); tma_prefetch_desc(&tmX_lo);
        tma_prefetch_desc(&tmW_hi); tma_prefetch_desc(&tmW_lo);
        for (int i = 0; i < kHaloBufs; ++i) { mbar_init(&halo_full[i], 1); mbar_init(&halo_empty[i], 1); }
        for (int i = 0; i < kWStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    if (warp >= 2) {
        const int i = threadIdx.x - 64, c = cout0 + i;
        s_scale[i] = (c < p.Cout) ? p.scale[c] : 0.f;
        s_bias[i] = (c < p.Cout) ? p.bias[c] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int ws = 0;
            uint32_t wphase = 0;
            const uint32_t halo_tx = 2u * p.h_rows * p.hP * p.kbytes, w_tx = 2u * 128u * p.kbytes;
            const int kelems = p.kbytes / 2;
            for (int it = 0; it < n_chunks; ++it) {
                const int ci = c_begin + it, hb = it % kHaloBufs;
                mbar_wait(&halo_empty[hb], ((it / kHaloBufs) & 1) ^ 1);
                mbar_expect_tx(&halo_full[hb], halo_tx);
                uint8_t *hdst = s_halo + hb * kHaloBufBytes;
                tma_load_4d(&tmX_hi, &halo_full[hb], hdst, ci * kelems, x0 - pad, y0 - pad, b, kEvictNormal);
                tma_load_4d(&tmX_lo, &halo_full[hb], hdst + p.h_plane_bytes, ci * kelems, x0 - pad, y0 - pad, b, kEvictNormal);
                for (int tap = 0; tap < taps; ++tap) {
                    mbar_wait(&w_empty[ws], wphase ^ 1);
                    mbar_expect_tx(&w_full[ws], w_tx);
                    uint8_t *wdst = s_w + ws * kWStageBytes;
                    const int kcoord = (tap * p.cin_chunks + ci) * kelems;
                    tma_load_2d(&tmW_hi, &w_full[ws], wdst, kcoord, cout0, kEvictLast);
                    tma_load_2d(&tmW_lo, &w_full[ws], wdst + kWTileBytes, kcoord, cout0, kEvictLast);
                    if (++ws == kWStages) { ws = 0; wphase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = umma_idesc_f16(128, N);
        const uint32_t t_corr = tmem_base + n_main * N;
        int ws = 0, mma_it = 0;
        uint32_t wphase = 0;
        for (int it = 0; it < n_chunks; ++it) {
            const int hb = it % kHaloBufs;
            mbar_wait(&halo_full[hb], (it / kHaloBufs) & 1);
            const uint32_t xh = smem_u32(s_halo + hb * kHaloBufBytes), xl = xh + p.h_plane_bytes;
            for (int tap = 0; tap < taps; ++tap, ++mma_it) {
                mbar_wait(&w_full[ws], wphase);
                tc_fence_after();
                if (elect_one()) {
                    const int kh = tap / p.ksize, kw = tap - kh * p.ksize;
                    const uint32_t shift = (kh * p.hP + kw) * p.kbytes;
                    const int ksteps = p.kbytes / 32;
                    const uint32_t wh = smem_u32(s_w + ws * kWStageBytes), wl = wh + kWTileBytes;
                    const uint32_t t_main = tmem_base + (mma_it % n_main) * N;
#pragma unroll 1
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t dwh = umma_desc_kmajor(wh + k * 32, p.kbytes), dwl = umma_desc_kmajor(wl + k * 32, p.kbytes);
                        const uint64_t dxh = umma_desc_kmajor(xh + shift + k * 32, p.kbytes),
                                       dxl = umma_desc_kmajor(xl + shift + k * 32, p.kbytes);
                        umma_f16(t_corr, dwl, dxh, idesc, (mma_it | k) ? 1u : 0u);
                        umma_f16(t_corr, dwh, dxl, idesc, 1u);
                        umma_f16(t_main, dwh, dxh, idesc, (mma_it >= n_main || k) ? 1u : 0u);
                    }
                    umma_commit(&w_empty[ws]);
                    if (tap == taps - 1) umma_commit(&halo_empty[hb]);
                    if (tap == taps - 1 && it == n_chunks - 1) umma_commit(accum_bar);
                }
                __syncwarp();
                if (++ws == kWStages) { ws = 0; wphase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue =====================
        // phase 1: each thread owns one output channel (TMEM lane): accumulators -> fp32 sum -> scale/bias/leaky
        //          -> shared-memory stage [pixel][channel] (the operand buffers are dead once accum_bar fires);
        // phase 2: the 128 threads walk (pixel, 8-channel group) items: coalesced 16-byte stores, 2x2 max-pool,
        //          hi/lo split, concat / space-to-depth addressing -- all through emit8().
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int q = warp & 3, et = threadIdx.x - 64;
        const int ch_local = q * 32 + lane;
        const float sc = s_scale[ch_local], bi = s_bias[ch_local];
        const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
        float *stage = reinterpret_cast<float *>(smem);
        constexpr int kLd = 132;                                  // floats per staged pixel (128 + pad, 16-byte aligned)
        const bool finish = p.splits == 1;
#pragma unroll 1
        for (int n0 = 0; n0 < N; n0 += 16) {
            uint32_t a[16];
            float v[16];
            tmem_ld16(lane_addr + n0, a);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(a[i]);
#pragma unroll 1
            for (int m = 1; m <= n_main; ++m) {
                tmem_ld16(lane_addr + m * N + n0, a);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += __uint_as_float(a[i]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float tv = v[i];
                if (finish) {
                    tv = fmaf(tv, sc, bi);
                    tv = p.act ? leaky(tv) : tv;
                }
                stage[(n0 + i) * kLd + ch_local] = tv;
            }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");           // epilogue warps only
        const int groups = min(16, (p.Cout - cout0 + 7) / 8);     // 8-channel groups that exist in this cout tile
        const long long mtot = (long long)p.B * p.H * p.W;
        const int rows_valid = min(p.hR, p.H - y0), cols_valid = min(p.hC, p.W - x0);
        if (!finish) {
            const int items = rows_valid * cols_valid * 16;
#pragma unroll 1
            for (int it = et; it < items; it += 128) {
                const int g = it & 15, px = it >> 4;
                const int r = px / cols_valid, c = px - r * cols_valid;
                const int chn = cout0 + g * 8;
                if (chn >= p.ldp) continue;
                const float4 *src = reinterpret_cast<const float4 *>(stage + (r * p.hP + c) * kLd + g * 8);
                float4 *dst = reinterpret_cast<float4 *>(
                    p.partial + ((long long)blockIdx.z * mtot + ((long long)b * p.H + y0 + r) * p.W + x0 + c) * p.ldp + chn);
                dst[0] = src[0];
                dst[1] = src[1];
            }
        } else {
            if (p.out.hi || p.out.f32) {
                const int items = rows_valid * cols_valid * groups;
#pragma unroll 1
                for (int it = et; it < items; it += 128) {
                    const int g = it % groups, px = it / groups;
                    const int r = px / cols_valid, c = px - r * cols_valid;
                    const float4 *src = reinterpret_cast<const float4 *>(stage + (r * p.hP + c) * kLd + g * 8);
                    const float4 lo4 = src[0], hi4 = src[1];
                    const float v8[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
                    emit8(p.out, b, y0 + r, x0 + c, cout0 + g * 8, p.Cout, v8);
                }
            }
            if (p.pool) {
                const int pr = rows_valid >> 1, pc = cols_valid >> 1;
                const int items = pr * pc * groups;
#pragma unroll 1
                for (int it = et; it < items; it += 128) {
                    const int g = it % groups, px = it / groups;
                    const int r = (px / pc) * 2, c = (px - (px / pc) * pc) * 2;
                    const float *s0 = stage + (r * p.hP + c) * kLd + g * 8;
                    float v8[8];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 a0 = reinterpret_cast<const float4 *>(s0)[h];
                        const float4 a1 = reinterpret_cast<const float4 *>(s0 + kLd)[h];
                        const float4 a2 = reinterpret_cast<const float4 *>(s0 + p.hP * kLd)[h];
                        const float4 a3 = reinterpret_cast<const float4 *>(s0 + (p.hP + 1) * kLd)[h];
                        v8[4 * h + 0] = fmaxf(fmaxf(a0.x, a1.x), fmaxf(a2.x, a3.x));
                        v8[4 * h + 1] = fmaxf(fmaxf(a0.y, a1.y), fmaxf(a2.y, a3.y));
                        v8[4 * h + 2] = fmaxf(fmaxf(a0.z, a1.z), fmaxf(a2.z, a3.z));
                        v8[4 * h + 3] = fmaxf(fmaxf(a0.w, a1.w), fmaxf(a2.w, a3.w));
                    }
                    emit8(p.pout, b, (y0 + r) >> 1, (x0 + c) >> 1, cout0 + g * 8, p.Cout, v8);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

int conv_halo_init() {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         HaloCfg<false>::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(conv_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloCfg<true>::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    // ask for the full shared-memory carve-out so that two small CTAs fit on an SM
    e = cudaFuncSetAttribute(conv_halo_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    return (int)e;
}

int launch_conv_halo(bool small, const CUtensorMap &x_hi, const CUtensorMap &x_lo, const CUtensorMap &w_hi,
                     const CUtensorMap &w_lo, const ConvParams &p, cudaStream_t st) {
    dim3 grid(p.B * p.h_tiles_x * p.h_tiles_y, (p.Cout + 127) / 128, p.splits);
    if (small)
        conv_halo_kernel<true><<<grid, kHaloThreads, HaloCfg<true>::kSmemBytes, st>>>(x_hi, x_lo, w_hi, w_lo, p);
    else
        conv_halo_kernel<false><<<grid, kHaloThreads, HaloCfg<false>::kSmemBytes, st>>>(x_hi, x_lo, w_hi, w_lo, p);
    return (int)cudaGetLastError();
}

}  // namespace b2t
