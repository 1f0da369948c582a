// conv_pm_kernel -- "pixel-major" persistent convolution for the narrow layers at the top of the network
// (conv_1: 3 -> 32, conv_2: 32 -> 64, conv_4: 128 -> 64 channels).
//
// conv_halo_kernel puts the output channels on the MMA's M dimension (128 TMEM lanes); with Cout = 32 or 64 half or
// three quarters of every tcgen05.mma would be wasted.  Here the roles are swapped:
//
//   D[pixel, cout] = sum_{tap, c} X[pixel + tap, c] * Wt[cout, tap, c]
//
// * M = 128 pixels of one image tile (A operand).  The tile's activation patch (with halo) sits in shared memory
//   exactly as in conv_halo_kernel -- rows ordered x-fastest with pitch hP -- and tap (kh,kw) is the same patch read
//   through a descriptor whose start address is advanced by (kh*hP + kw) rows.
// * N = Cout rounded up to 16 (B operand = the layer's weights, all (tap, chunk) tiles resident in shared memory,
//   loaded once per CTA).
// * mode 0 (conv_2, conv_4): fp16 (hi, lo) activations and weights, three MMAs per 16-deep K step as everywhere else.
// * mode 1 (conv_1 on uint8 frames): the frame is stored as fp16 integers 0..255 (exact, so there is no lo plane),
//   8 channels per pixel (3 used) = one 16-byte core-matrix row.  With the no-swizzle K-major layout
//   (row m at 16*m bytes: SBO = 128 B, K group j at +16*j bytes: LBO = 16 B) the operand row of pixel m covers the
//   pixels m .. m+3 of the patch: the kw taps of the 3x3 window are folded into K by the descriptor itself (an
//   im2col that costs nothing), one K = 32 step per kernel row kh.  1/255 is folded into the epilogue scale.
//   An item is a PAIR of image rows x up to 126 columns (M = columns): two accumulator pairs, one per row, so the
//   2x2 max-pool happens in registers (rows in-thread, columns by one shuffle), BN + LeakyReLU are applied after the
//   pool (monotonic per channel) and every thread writes 32 contiguous bytes per plane -- no shared-memory staging.
// * TMEM holds two accumulator sets (main + correction, N columns each) so the 8-warp epilogue of tile j overlaps
//   the MMAs of tile j+1; the epilogue thread owns one pixel (TMEM lane) and half of the channels, stages
//   [pixel][channel] in shared memory and shares the coalesced store phase (2x2 max-pool, hi/lo split) with
//   conv_halo_kernel.
#include "kernels.cuh"

namespace b2t {

constexpr int kPmThreads = 64 + 256;

__device__ __forceinline__ void umma_pm3(uint32_t t_main, uint32_t t_corr, uint32_t xh, uint32_t xl, uint32_t wh, uint32_t wl,
                                         uint32_t hi, uint32_t idesc, uint32_t acc) {
    const uint64_t dxh = umma_desc_make(xh, hi), dxl = umma_desc_make(xl, hi);
    const uint64_t dwh = umma_desc_make(wh, hi), dwl = umma_desc_make(wl, hi);
    umma_f16(t_corr, dxh, dwl, idesc, acc);
    umma_f16(t_corr, dxl, dwh, idesc, 1u);
    umma_f16(t_main, dxh, dwh, idesc, acc);
}

// The same three products in TWO instructions when the lo weight tile directly follows the hi tile in shared memory
// (w_rows == N): B = [w_hi ; w_lo] is then one K-major operand of 2N rows, and x_hi * [w_hi ; w_lo] fills the adjacent
// accumulators [main | corr] at once; x_lo * w_hi follows into corr.  Same addends in the same order per accumulator
// (bit-identical), but x_hi and w_hi are read from shared memory once instead of twice: 14 KB instead of 18 KB per K step
// at N = 64, where the kernel is bound by the tensor core's shared-memory reads (ncu: l1tex data pipe 66-72 % of peak
// with the tensor pipe 48 % active).
__device__ __forceinline__ void umma_pm2(uint32_t t_main, uint32_t N, uint32_t xh, uint32_t xl, uint32_t wh, uint32_t hi,
                                         uint32_t idesc_2n, uint32_t idesc_n, uint32_t acc) {
    const uint64_t dxh = umma_desc_make(xh, hi), dxl = umma_desc_make(xl, hi), dwh = umma_desc_make(wh, hi);
    umma_f16(t_main, dxh, dwh, idesc_2n, acc);
    umma_f16(t_main + N, dxl, dwh, idesc_n, 1u);
}

__global__ void __launch_bounds__(kPmThreads, 2)
conv_pm_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
               const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
               const ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    uint64_t *patch_full = reinterpret_cast<uint64_t *>(smem);   // [2]
    uint64_t *patch_empty = patch_full + 2;
    uint64_t *acc_full = patch_empty + 2;
    uint64_t *acc_empty = acc_full + 2;
    uint64_t *w_full = acc_empty + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(w_full + 1);
    float *s_scale = reinterpret_cast<float *>(smem + 256);       // [128]
    float *s_bias = s_scale + 128;
    uint8_t *s_patch = smem + 2048;
    float *stage = reinterpret_cast<float *>(s_patch + 2 * p.pw_patch_bytes);
    uint8_t *s_w = reinterpret_cast<uint8_t *>(stage) + 2 * p.pw_stage_bytes;      // two stage buffers (one per epilogue group)

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform
    const int pad = p.ksize >> 1, taps = p.ksize * p.ksize;
    const int N = p.pm_n;
    const int n_items = p.B * p.h_tiles_y * p.h_tiles_x;
    const int n_wtiles = taps * p.cin_chunks;
    const bool ints = p.pm_mode == 1;

    // ---- prologue: barriers, zeroed patch buffers (the MMA reads a few rows past the loaded box: finite zeros),
    //      scale/bias, conv_1's packed weights, TMEM
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmX_hi);
        if (!ints) { tma_prefetch_desc(&tmX_lo); tma_prefetch_desc(&tmW_hi); tma_prefetch_desc(&tmW_lo); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&patch_full[i], 1); mbar_init(&patch_empty[i], 1);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 1);
        }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 2 * p.pw_patch_bytes / 16; i += kPmThreads)
        reinterpret_cast<uint4 *>(s_patch)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 128) {
        s_scale[threadIdx.x] = threadIdx.x < p.Cout ? __ldg(p.scale + threadIdx.x) : 0.f;
        s_bias[threadIdx.x] = threadIdx.x < p.Cout ? __ldg(p.bias + threadIdx.x) : 0.f;
    }
    if (ints)
        for (int i = threadIdx.x; i < p.pm_w_bytes / 16; i += kPmThreads)
            reinterpret_cast<uint4 *>(s_w)[i] = __ldg(reinterpret_cast<const uint4 *>(p.pm_w) + i);
    fence_proxy_async();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.pm_tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    griddep_launch();

    // item -> (image, tile row, tile column) with host-made magic multipliers (exact for item < 2^32 / divisor)
    auto decode_item = [&](int item, int &b, int &y0, int &x0) {
        const int t = p.magic_tx ? (int)__umulhi((unsigned)item, p.magic_tx) : item;
        const int tx = item - t * p.h_tiles_x;
        b = p.magic_ty ? (int)__umulhi((unsigned)t, p.magic_ty) : t;
        const int ty = t - b * p.h_tiles_y;
        y0 = ty * p.hR; x0 = tx * p.hC;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        const int kelems = p.kbytes / 2;
        const uint32_t patch_tx = (ints ? 1u : 2u) * p.h_rows * p.hP * p.kbytes;
        if (!ints && elect_one()) {                       // every weight tile of the layer, once
            mbar_expect_tx(w_full, (uint32_t)n_wtiles * p.pw_tile_bytes);
            for (int ci = 0; ci < p.cin_chunks; ++ci)
                for (int tap = 0; tap < taps; ++tap) {
                    uint8_t *wdst = s_w + (ci * taps + tap) * p.pw_tile_bytes;
                    const int kcoord = (tap * p.cin_chunks + ci) * kelems;
                    tma_load_2d(&tmW_hi, w_full, wdst, kcoord, 0, kEvictLast);
                    tma_load_2d(&tmW_lo, w_full, wdst + p.pw_tile_bytes / 2, kcoord, 0, kEvictLast);
                }
        }
        __syncwarp();
        griddep_wait();          // the activations / the frame copy come from the previous kernel of the stream
        int g_chunk = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int b, y0, x0;
            decode_item(item, b, y0, x0);
            for (int ci = 0; ci < p.cin_chunks; ++ci, ++g_chunk) {
                const int hb = g_chunk & 1;
                mbar_wait(&patch_empty[hb], ((g_chunk >> 1) & 1) ^ 1);
                uint8_t *hdst = s_patch + hb * p.pw_patch_bytes;
                const bool skip_x = (B2T_DBG_BITS(p) & 2) && g_chunk >= 2;
                if (B2T_TRACE_PTR(p) && blockIdx.x == 0 && g_chunk < 64 && lane == 0) B2T_TRACE_PTR(p)[g_chunk * 16 + 0] = clock64();
                if (elect_one()) {
                    if (skip_x) mbar_arrive(&patch_full[hb]);
                    else {
                        mbar_expect_tx(&patch_full[hb], patch_tx);
                        tma_load_4d(&tmX_hi, &patch_full[hb], hdst, ci * kelems, x0 - pad, y0 - pad, b, kEvictNormal);
                        if (!ints)
                            tma_load_4d(&tmX_lo, &patch_full[hb], hdst + p.h_plane_bytes, ci * kelems, x0 - pad, y0 - pad, b, kEvictNormal);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform arithmetic, one elected lane issues) =====================
        const uint32_t idesc = umma_idesc_f16(128, N);
        const uint32_t p16_0 = umma_desc_lo(smem_u32(s_patch)), pb16 = p.pw_patch_bytes >> 4;
        const uint32_t w16 = (smem_u32(s_w) & 0x3FFFF) >> 4;
        int g_chunk = 0, j = 0;
        if (!ints) { mbar_wait(w_full, 0); tc_fence_after(); }
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++j) {
            const int ab = j & 1;
            const bool tr = B2T_TRACE_PTR(p) && blockIdx.x == 0 && j < 64 && lane == 0;
            if (tr) B2T_TRACE_PTR(p)[j * 16 + 1] = clock64();
            mbar_wait(&acc_empty[ab], ((j >> 1) & 1) ^ 1);       // the epilogue has drained this accumulator set
            tc_fence_after();
            if (tr) B2T_TRACE_PTR(p)[j * 16 + 2] = clock64();
            const uint32_t t_main = tmem_base + ab * 2 * N, t_corr = t_main + N;
            if (ints) {
                // conv_1: A rows = pixels (16 B each, SBO 128 B, LBO 16 B -> a row spans 4 pixels = K 32);
                // B = [32 cout][32 k] per kernel row, core matrices of 128 B: LBO 128 B (next K group), SBO 512 B
                const int hb = g_chunk & 1;
                mbar_wait(&patch_full[hb], (g_chunk >> 1) & 1);
                tc_fence_after();
                if (tr) B2T_TRACE_PTR(p)[j * 16 + 3] = clock64();
                const uint32_t a_hi = (128u >> 4) | (1u << 14), b_hi = (512u >> 4) | (1u << 14);
                const uint32_t xa = p16_0 + hb * pb16;                       // LBO field = 1 (16 B) from umma_desc_lo
                const uint32_t wb = w16 | ((128u >> 4) << 16);
                // The item is a pair of image rows (one pooled row): output row r uses patch rows r + kh and accumulates
                // into columns [r*N, r*N + N) of the item's TMEM buffer.  Patch rows 1 and 2 feed BOTH output rows (with
                // kernel rows j and j-1), so they are read once: the packed weights keep the kernel rows in DESCENDING
                // order per plane, and the N = 64 operand that starts at kernel row j is [W_j ; W_j-1] = the columns of
                // row 0 and row 1.  Patch rows 0 / 3 feed one output row each (N = 32).  16 MMAs and 88 KB of shared-memory
                // operand reads per item instead of 24 and 120 KB -- the kernel is bound by those reads (ncu: tensor
                // data pipe 72 % of peak with the tensor pipe 30 % active).  u*w_hi and u*w_lo share the accumulator: the
                // chain is 12 MMAs short, so the w_lo terms (2^-11 of the sum) keep 13 bits.
                const uint32_t t_buf = tmem_base + ab * 2 * N;
                const uint32_t idesc2 = umma_idesc_f16(128, 2 * N);
                auto wdesc = [&](int kh, int plane, int ks) {
                    return umma_desc_make(wb + plane * (6144 >> 4) + (2 - kh) * (2048 >> 4) + ks * (256 >> 4), b_hi);
                };
                if (elect_one()) {
                    if (!(B2T_DBG_BITS(p) & 8)) {
#pragma unroll
                        for (int pr = 1; pr <= 2; ++pr) {
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                const uint64_t da = umma_desc_make(xa + pr * p.hP + 2 * ks, a_hi);
                                umma_f16(t_buf, da, wdesc(pr, 0, ks), idesc2, (pr == 1 && ks == 0) ? 0u : 1u);
                                umma_f16(t_buf, da, wdesc(pr, 1, ks), idesc2, 1u);
                            }
                        }
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const uint64_t d0 = umma_desc_make(xa + 2 * ks, a_hi), d3 = umma_desc_make(xa + 3 * p.hP + 2 * ks, a_hi);
                            umma_f16(t_buf, d0, wdesc(0, 0, ks), idesc, 1u);          // patch row 0 -> output row 0, kernel row 0
                            umma_f16(t_buf, d0, wdesc(0, 1, ks), idesc, 1u);
                            umma_f16(t_buf + N, d3, wdesc(2, 0, ks), idesc, 1u);      // patch row 3 -> output row 1, kernel row 2
                            umma_f16(t_buf + N, d3, wdesc(2, 1, ks), idesc, 1u);
                        }
                    }
                    umma_commit(&patch_empty[hb]);
                    umma_commit(&acc_full[ab]);
                }
                __syncwarp();
                if (tr) B2T_TRACE_PTR(p)[j * 16 + 4] = clock64();
                ++g_chunk;
                continue;
            }
            const uint32_t dhi = umma_desc_hi(p.kbytes);
            const uint32_t kb16 = p.kbytes >> 4, row16 = p.hP * kb16 - (p.ksize - 1) * kb16;
            const uint32_t xl_off = p.h_plane_bytes >> 4, wl_off = p.pw_tile_bytes >> 5, wt16 = p.pw_tile_bytes >> 4;
            const bool k128 = p.kbytes == 128;
            // hi and lo weight tiles adjacent and 2N accumulator columns in one instruction: merged issue (umma_pm2)
            const bool merged = p.pw_tile_bytes == 2 * N * p.kbytes && 2 * N <= 256;
            const uint32_t idesc_2n = umma_idesc_f16(128, 2 * N);
            uint32_t acc = 0, wh = w16 | (1u << 16);
            for (int ci = 0; ci < p.cin_chunks; ++ci, ++g_chunk) {
                const int hb = g_chunk & 1;
                mbar_wait(&patch_full[hb], (g_chunk >> 1) & 1);
                tc_fence_after();
                uint32_t xh = p16_0 + hb * pb16;
                int kw = 0;
                for (int tap = 0; tap < taps; ++tap) {
                    const uint32_t xl = xh + xl_off, wl = wh + wl_off;
                    const bool last_tap = tap == taps - 1;
                    if (elect_one()) {
                        if (B2T_DBG_BITS(p) & 8) {
                        } else if (merged) {
                            umma_pm2(t_main, N, xh, xl, wh, dhi, idesc_2n, idesc, acc);
                            umma_pm2(t_main, N, xh + 2, xl + 2, wh + 2, dhi, idesc_2n, idesc, 1u);
                            if (k128) {
                                umma_pm2(t_main, N, xh + 4, xl + 4, wh + 4, dhi, idesc_2n, idesc, 1u);
                                umma_pm2(t_main, N, xh + 6, xl + 6, wh + 6, dhi, idesc_2n, idesc, 1u);
                            }
                        } else {
                            umma_pm3(t_main, t_corr, xh, xl, wh, wl, dhi, idesc, acc);
                            umma_pm3(t_main, t_corr, xh + 2, xl + 2, wh + 2, wl + 2, dhi, idesc, 1u);
                            if (k128) {
                                umma_pm3(t_main, t_corr, xh + 4, xl + 4, wh + 4, wl + 4, dhi, idesc, 1u);
                                umma_pm3(t_main, t_corr, xh + 6, xl + 6, wh + 6, wl + 6, dhi, idesc, 1u);
                            }
                        }
                        if (last_tap) umma_commit(&patch_empty[hb]);
                        if (last_tap && ci == p.cin_chunks - 1) umma_commit(&acc_full[ab]);
                    }
                    __syncwarp();
                    acc = 1;
                    wh += wt16;
                    if (++kw == p.ksize) { kw = 0; xh += row16; } else xh += kb16;
                }
            }
        }
    } else {
        // ===================== epilogue: two independent groups of 4 warps, group g owns accumulator set g and
        // stage buffer g and takes every other item, so the (latency-bound) epilogues of two tiles are in flight
        // while the MMAs of the following ones run =====================
        const int grp = (warp - 2) >> 2, q = warp & 3;
        const int n = q * 32 + lane;                              // this thread's pixel (TMEM lane)
        const int et = (threadIdx.x - 64) & 127;
        const int ld = p.pm_stage_ld;
        float *my_stage = stage + grp * (p.pw_stage_bytes >> 2);
        const uint32_t t_main = tmem_base + grp * 2 * N + (uint32_t(q * 32) << 16);
        int k = 0;
        for (int item = blockIdx.x + grp * gridDim.x; item < n_items; item += 2 * gridDim.x, ++k) {
            int b, y0, x0;
            decode_item(item, b, y0, x0);
            const int jj = 2 * k + grp;
            const bool tr = B2T_TRACE_PTR(p) && blockIdx.x == 0 && jj < 64 && et == 0;
            if (tr) B2T_TRACE_PTR(p)[jj * 16 + 5] = clock64();
            mbar_wait(&acc_full[grp], k & 1);
            tc_fence_after();
            if (tr) B2T_TRACE_PTR(p)[jj * 16 + 6] = clock64();
            if (ints) {
                // conv_1: the item is one pooled row.  TMEM lane = image column c; the item's buffer holds
                // [row 0 main | row 0 corr | row 1 main | row 1 corr] x 32 channels.  The 2x2 max-pool runs in
                // registers: vertical = the two rows of the same lane, horizontal = lane pairs (shuffle).  The host
                // negates the weights of channels whose folded-BN scale is negative and passes |scale|, so BN +
                // LeakyReLU are increasing in the raw sum for every channel and can be applied AFTER the pool to
                // max(raw) -- bit-identical to pooling the activated values.
                // Even lanes then finish channels 0..15, odd lanes 16..31, and write 32 contiguous bytes per plane.
                const uint32_t t_buf = tmem_base + grp * 2 * N + (uint32_t(q * 32) << 16);
                const int c = n, xq = (x0 + c) >> 1;
                const bool valid = c < p.hC && x0 + c < p.W;
                const int half_mine = lane & 1;
                float keep[16];
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {                 // 8 channels at a time keeps the live set small
                    uint32_t ua[8], ub[8];
                    tmem_ld8(t_buf + 8 * g8, ua);
                    tmem_ld8(t_buf + N + 8 * g8, ub);
                    tmem_ld_wait();
                    if (p.out.hi && valid) {              // full-resolution copy requested (keep_prepool): both rows
                        float v8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) { const float tv = fmaf(__uint_as_float(ua[i]), s_scale[8 * g8 + i], s_bias[8 * g8 + i]); v8[i] = fmaxf(tv, 0.1f * tv); }
                        emit8(p.out, b, y0, x0 + c, 8 * g8, p.Cout, v8);
#pragma unroll
                        for (int i = 0; i < 8; ++i) { const float tv = fmaf(__uint_as_float(ub[i]), s_scale[8 * g8 + i], s_bias[8 * g8 + i]); v8[i] = fmaxf(tv, 0.1f * tv); }
                        emit8(p.out, b, y0 + 1, x0 + c, 8 * g8, p.Cout, v8);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float mx = fmaxf(__uint_as_float(ua[i]), __uint_as_float(ub[i]));
                        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                        if ((g8 >> 1) == half_mine) keep[8 * (g8 & 1) + i] = mx;
                    }
                }
                tc_fence_before();
                if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
                if (et == 0) mbar_arrive(&acc_empty[grp]);             // TMEM buffer drained by all four warps
                if (tr) B2T_TRACE_PTR(p)[jj * 16 + 7] = clock64();
                if (valid && !(B2T_DBG_BITS(p) & 4)) {
                    const int ch0 = 16 * half_mine;
                    uint32_t hw[8], lw[8];                       // 16 channels as packed fp16 pairs (hi plane, lo plane)
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float t2[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const float tv = fmaf(keep[2 * i + e], s_scale[ch0 + 2 * i + e], s_bias[ch0 + 2 * i + e]);
                            t2[e] = fminf(fmaxf(fmaxf(tv, 0.1f * tv), -65504.f), 65504.f);
                        }
                        const __half2 h2 = __floats2half2_rn(t2[0], t2[1]);
                        const float2 hf = __half22float2(h2);
                        const __half2 l2 = __floats2half2_rn(t2[0] - hf.x, t2[1] - hf.y);
                        hw[i] = *reinterpret_cast<const uint32_t *>(&h2);
                        lw[i] = *reinterpret_cast<const uint32_t *>(&l2);
                    }
                    op_t *dst = p.pout.hi + (((long long)b * (p.H >> 1) + (y0 >> 1)) * (p.W >> 1) + xq) * p.pout.pix_stride_b +
                                p.pout.ch_off_b + ch0;
                    reinterpret_cast<uint4 *>(dst)[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    reinterpret_cast<uint4 *>(dst)[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                    reinterpret_cast<uint4 *>(dst + p.pout.plane_stride)[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    reinterpret_cast<uint4 *>(dst + p.pout.plane_stride)[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
                }
                if (tr) B2T_TRACE_PTR(p)[jj * 16 + 8] = clock64();
                continue;
            }
            for (int c0 = 0; c0 < N; c0 += 32) {               // N is a multiple of 32 here (32 or 64)
                uint32_t a[32], c2[32];
                tmem_ld32(t_main + c0, a);
                tmem_ld32(t_main + N + c0, c2);
                tmem_ld_wait();
                float4 *dst = reinterpret_cast<float4 *>(my_stage + n * ld + c0);
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 sc = *reinterpret_cast<const float4 *>(s_scale + c0 + 4 * i4);
                    const float4 bi = *reinterpret_cast<const float4 *>(s_bias + c0 + 4 * i4);
                    float4 v;
                    v.x = fmaf(__uint_as_float(a[4 * i4 + 0]) + __uint_as_float(c2[4 * i4 + 0]), sc.x, bi.x);
                    v.y = fmaf(__uint_as_float(a[4 * i4 + 1]) + __uint_as_float(c2[4 * i4 + 1]), sc.y, bi.y);
                    v.z = fmaf(__uint_as_float(a[4 * i4 + 2]) + __uint_as_float(c2[4 * i4 + 2]), sc.z, bi.z);
                    v.w = fmaf(__uint_as_float(a[4 * i4 + 3]) + __uint_as_float(c2[4 * i4 + 3]), sc.w, bi.w);
                    if (p.act) { v.x = fmaxf(v.x, 0.1f * v.x); v.y = fmaxf(v.y, 0.1f * v.y); v.z = fmaxf(v.z, 0.1f * v.z); v.w = fmaxf(v.w, 0.1f * v.w); }
                    dst[i4] = v;
                }
            }
            tc_fence_before();
            if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
            if (et == 0) mbar_arrive(&acc_empty[grp]);
            if (tr) B2T_TRACE_PTR(p)[jj * 16 + 7] = clock64();
            if (!(B2T_DBG_BITS(p) & 4)) epilogue_store(p, my_stage, ld, p.pm_glog, b, y0, x0, 0, 0, et, 128);
            if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
            if (tr) B2T_TRACE_PTR(p)[jj * 16 + 8] = clock64();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.pm_tmem_cols) : "memory");
}

// uint8 HWC3 frame -> fp16 [pixel][8] (channels 3..7 zero), the integer values 0..255 exactly.  The source is n
// segments of seg_pix pixels each, seg_stride bytes apart (S streams x T consecutive frames of longer clips: the
// window is gathered by this kernel, no staging copy); dense input: seg_pix = npix.
__global__ void __launch_bounds__(256) frames_to_c8_kernel(const uint8_t *__restrict__ src, uint4 *__restrict__ dst, long long npix,
                                                           long long seg_pix, long long seg_stride) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        const long long seg = i / seg_pix;
        const uint8_t *s = src + seg * seg_stride + 3 * (i - seg * seg_pix);
        const __half2 a = __floats2half2_rn((float)s[0], (float)s[1]);
        const __half2 b = __floats2half2_rn((float)s[2], 0.f);
        uint4 o;
        o.x = *reinterpret_cast<const uint32_t *>(&a);
        o.y = *reinterpret_cast<const uint32_t *>(&b);
        o.z = 0; o.w = 0;
        dst[i] = o;
    }
}

int conv_pm_init() {
    cudaError_t e = cudaFuncSetAttribute(conv_pm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaFuncSetAttribute(conv_pm_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}

int conv_pm_smem_bytes(const ConvParams &p) {
    const int w = p.pm_mode == 1 ? p.pm_w_bytes : p.ksize * p.ksize * p.cin_chunks * p.pw_tile_bytes;
    return 1024 /*align*/ + 2048 /*barriers, scale, bias*/ + 2 * p.pw_patch_bytes + 2 * p.pw_stage_bytes + w;
}

int launch_conv_pm(int n_sm, const CUtensorMap &x_hi, const CUtensorMap &x_lo, const CUtensorMap &w_hi,
                   const CUtensorMap &w_lo, const ConvParams &p, cudaStream_t st) {
    const int items = p.B * p.h_tiles_x * p.h_tiles_y;
    const int smem = conv_pm_smem_bytes(p);
    // two co-resident CTAs per SM when shared memory and TMEM allow (conv_1, conv_4): they hide each other's
    // per-tile barrier and store latencies
    const int per_sm = (2 * (smem + 1024) <= 227 * 1024 && 2 * p.pm_tmem_cols <= 512) ? 2 : 1;
    const int grid = items < n_sm * per_sm ? items : n_sm * per_sm;
    return (int)launch_pdl(conv_pm_kernel, dim3(grid), dim3(kPmThreads), smem, st, x_hi, x_lo, w_hi, w_lo, p);
}

int launch_frames_to_c8(const void *frames, void *dst, long long npix, long long seg_pix, long long seg_stride, cudaStream_t st) {
    const int blocks = (int)((npix + 255) / 256 < 148 * 16 ? (npix + 255) / 256 : 148 * 16);
    frames_to_c8_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint8_t *>(frames), reinterpret_cast<uint4 *>(dst), npix,
                                                seg_pix, seg_stride);
    return (int)cudaGetLastError();
}

}  // namespace b2t
