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
// * split-K over channel chunks: raw fp32 partials, finished by splitk_epilogue_kernel.
// * persistent: one CTA per SM walks a static list of (pixel tile, cout tile, K split) items; one elected thread per
//   role issues (TMA producer warp, MMA warp), all role arithmetic is warp-uniform so that ptxas keeps descriptors in
//   uniform registers and emits back-to-back UTCHMMA ('The issuing thread').
#include <string.h>

#include "kernels.cuh"

namespace b2t {

constexpr int kWTileBytes = 128 * 64 * 2;            // one plane of a weight tile
constexpr int kWStageBytes = 2 * kWTileBytes;
constexpr int kEpiThreads = 256;                     // 8 epilogue warps: two per TMEM lane quarter, each takes half the columns
constexpr int kHaloThreads = 64 + kEpiThreads;

// Epilogue of one tile, run by the 8 epilogue warps (CTA warps 2..9).
// phase 1: a thread owns one output channel (TMEM lane = warp%4 quarter) and half of the tile's pixel columns:
//          accumulators -> fp32 sum -> scale/bias/LeakyReLU -> shared-memory stage [pixel][channel];
// phase 2: threads walk (pixel, 8-channel group) items: 16-byte coalesced stores, 2x2 max-pool, hi/lo split,
//          concat / space-to-depth addressing through emit8(), or raw fp32 split-K partials.
// `acc_free` (may be NULL) is arrived on once every accumulator has been read (persistent kernel).
constexpr int kStageLd = 132;                            // floats per staged pixel (128 + pad, keeps 16-byte alignment)
__device__ __forceinline__ void halo_epilogue(const ConvParams &p, float *stage, uint32_t tmem_acc, int n_acc, int N,
                                              int b, int y0, int x0, int cout0, int zsplit, uint64_t *acc_free) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int et = threadIdx.x - 64;                     // 0..255
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int ch_local = q * 32 + lane, ch = cout0 + ch_local;
    const bool finish = p.splits == 1;
    const float sc = (finish && ch < p.Cout) ? __ldg(p.scale + ch) : 0.f;
    const float bi = (finish && ch < p.Cout) ? __ldg(p.bias + ch) : 0.f;
    const uint32_t lane_addr = tmem_acc + (uint32_t(q * 32) << 16);
#pragma unroll 1
    for (int n0 = half * 16; n0 < N; n0 += 32) {
        uint32_t a[16], c2[16];
        float v[16];
        tmem_ld16(lane_addr + n0, a);
        tmem_ld16(lane_addr + (n_acc - 1) * N + n0, c2);   // the correction accumulator is the last one
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(a[i]) + __uint_as_float(c2[i]);
#pragma unroll 1
        for (int m = 1; m < n_acc - 1; ++m) {              // extra main accumulators (round-robin scheme)
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
                tv = p.act ? fmaxf(tv, 0.1f * tv) : tv;    // LeakyReLU(0.1) == max(x, 0.1 x)
            }
            stage[(n0 + i) * kStageLd + ch_local] = tv;
        }
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (acc_free && et == 0) mbar_arrive(acc_free);
    if (B2T_TRACE_PTR(p) && blockIdx.x == 0 && et == 0) B2T_TRACE_PTR(p)[511] = clock64();   // last phase-1 completion (developer trace)
    if (B2T_DBG_BITS(p) & 4) return;
    epilogue_store(p, stage, kStageLd, 4, b, y0, x0, cout0, zsplit, et, kEpiThreads);
}

// Multi-pass variant (conv_halo_kernel<big>, p.k_passes > 1, N <= 192): the K range of an item is accumulated in
// k_passes separate tensor-core chains (the tensor core truncates every addend to the accumulator's exponent, so one
// chain must stay short: DESIGN.md section 4).  After each pass but the last the epilogue threads read the accumulators
// and PARK the running fp32 sums while the MMA warp restarts the accumulators: pixel columns [0, 128) in the 128 TMEM
// columns the two accumulators (2 x N <= 384) leave free (tcgen05.st), columns [128, 192) in 32 registers per thread.
// After the last pass the sums are added in pass order -- the same arithmetic as the split-K partial buffers + finish
// kernel this replaces, without their HBM round trip and second launch.
constexpr int kPassMaxN = 192, kPassTmemCols = 128;
template <bool kLast>
__device__ __forceinline__ void halo_pass_accumulate(const ConvParams &p, float (&keep)[2][16], float *stage, uint32_t tmem_acc,
                                                     int N, int cout0, bool first_pass) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int ch_local = q * 32 + lane, ch = cout0 + ch_local;
    const bool finish = p.splits == 1;                           // a K-split item stages its raw sum (split-K partial)
    const float sc = (kLast && finish && ch < p.Cout) ? __ldg(p.scale + ch) : 0.f;
    const float bi = (kLast && finish && ch < p.Cout) ? __ldg(p.bias + ch) : 0.f;
    const uint32_t lane_addr = tmem_acc + (uint32_t(q * 32) << 16);
    const uint32_t park = lane_addr + 2 * N;                     // first spare column
#pragma unroll
    for (int j = 0; j < kPassMaxN / 32; ++j) {
        const int n0 = half * 16 + 32 * j;                       // this thread's j-th group of 16 columns
        if (n0 < N) {
            uint32_t a[16], c2[16];
            tmem_ld16(lane_addr + n0, a);                        // main accumulator
            tmem_ld16(lane_addr + N + n0, c2);                   // correction accumulator
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __float_as_uint(__uint_as_float(a[i]) + __uint_as_float(c2[i]));
            if (!first_pass) {                                   // + the parked sum of the earlier passes, pass order
                if (j < kPassTmemCols / 32) {
                    tmem_ld16(park + n0, c2);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) a[i] = __float_as_uint(__uint_as_float(c2[i]) + __uint_as_float(a[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) a[i] = __float_as_uint(keep[j - kPassTmemCols / 32][i] + __uint_as_float(a[i]));
                }
            }
            if (kLast) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float tv = __uint_as_float(a[i]);
                    if (finish) {
                        tv = fmaf(tv, sc, bi);
                        tv = p.act ? fmaxf(tv, 0.1f * tv) : tv;
                    }
                    stage[(n0 + i) * kStageLd + ch_local] = tv;
                }
            } else if (j < kPassTmemCols / 32) {
                tmem_st16(park + n0, a);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) keep[j - kPassTmemCols / 32][i] = __uint_as_float(a[i]);
            }
        }
    }
    if (!kLast) tmem_st_wait();
}

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
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t *s_halo = smem;
    uint8_t *s_w = smem + kHaloBufs * kHaloBufBytes;
    uint8_t *tail = s_w + kWStages * kWStageBytes;
    uint64_t *halo_full = reinterpret_cast<uint64_t *>(tail);   // [2]
    uint64_t *halo_empty = halo_full + 2;                       // [2]
    uint64_t *w_full = halo_empty + 2;                          // [kWStages]
    uint64_t *w_empty = w_full + kWStages;                      // [kWStages]
    uint64_t *accum_bar = w_empty + kWStages;                   // MMA -> epilogue: the item's accumulators are complete
    uint64_t *acc_empty = accum_bar + 1;                        // epilogue -> MMA: TMEM has been read
    uint64_t *stage_free = acc_empty + 1;                       // epilogue -> producer: the staged tile has been stored
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(stage_free + 1);

    // warp-uniform role index: the shuffle lets the compiler prove uniformity, so the role branches are uniform
    // branches and the MMA issuer's descriptor arithmetic stays in uniform registers (no R2UR / waterfall loops
    // around UTCHMMA -- the single issuing thread is otherwise the kernel's bottleneck)
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int pad = p.ksize >> 1, taps = p.ksize * p.ksize;
    const int N = p.hN;
    // The CTA is persistent: it walks items (pixel tile, cout tile, K split) blockIdx.x, + gridDim.x, ...  Barriers,
    // TMEM and the weight ring live across items; the weights of the next item are prefetched during the epilogue
    // of the current one (the staged tile aliases the patch buffers only, unless N > 248).
    const int gx = p.B * p.h_tiles_x * p.h_tiles_y, gy = (p.Cout + 127) >> 7;
    const int n_items = gx * gy * p.splits;
    const int per = (p.cin_chunks + p.splits - 1) / p.splits;
    const bool stage_hits_w = N * kStageLd * 4 > kHaloBufs * kHaloBufBytes;
    auto decode_item = [&](int item, int &b, int &y0, int &x0, int &cout0, int &z, int &c_begin, int &n_chunks) {
        int t = item % gx;
        const int rest = item / gx;
        const int tx = t % p.h_tiles_x;  t /= p.h_tiles_x;
        const int ty = t % p.h_tiles_y;
        b = t / p.h_tiles_y;
        x0 = tx * p.hC; y0 = ty * p.hR;
        cout0 = (rest % gy) * 128;
        z = rest / gy;
        c_begin = z * per;
        n_chunks = min(p.cin_chunks, c_begin + per) - c_begin;         // host guarantees >= 1
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX_hi); tma_prefetch_desc(&tmX_lo);
        tma_prefetch_desc(&tmW_hi); tma_prefetch_desc(&tmW_lo);
        for (int i = 0; i < kHaloBufs; ++i) { mbar_init(&halo_full[i], 1); mbar_init(&halo_empty[i], 1); }
        for (int i = 0; i < kWStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        mbar_init(accum_bar, 1); mbar_init(acc_empty, 1); mbar_init(stage_free, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    griddep_launch();            // the next kernel of the stream may start its own prologue on SMs we have left

    if (warp == 0) {
        // ===================== TMA producer (whole warp runs the uniform loops, one elected lane issues) =====
        int ws = 0, g_it = 0, k = 0;
        uint32_t wphase = 0;
        const uint32_t halo_tx = 2u * p.h_rows * p.hP * p.kbytes, w_tx = 2u * 128u * p.kbytes;
        const int kelems = p.kbytes / 2;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
            int b, y0, x0, cout0, z, c_begin, n_chunks;
            decode_item(item, b, y0, x0, cout0, z, c_begin, n_chunks);
            // weight tiles are requested strictly in (chunk, tap) order; (w_it, w_tap) is the next one
            int w_it = 0, w_tap = 0, w_seq = 0;
            auto issue_next_w = [&]() {
                mbar_wait(&w_empty[ws], wphase ^ 1);
                uint8_t *wdst = s_w + ws * kWStageBytes;
                const int kcoord = (w_tap * p.cin_chunks + c_begin + w_it) * kelems;
                if (elect_one()) {
                    mbar_expect_tx(&w_full[ws], w_tx);
                    tma_load_2d(&tmW_hi, &w_full[ws], wdst, kcoord, cout0, kEvictLast);
                    tma_load_2d(&tmW_lo, &w_full[ws], wdst + kWTileBytes, kcoord, cout0, kEvictLast);
                }
                __syncwarp();
                ++w_seq;
                if (++w_tap == taps) { w_tap = 0; ++w_it; }
                if (++ws == kWStages) { ws = 0; wphase ^= 1; }
            };
            if (k > 0) {
                if (!stage_hits_w) {
                    const int pf = min(kWStages, n_chunks * taps);
                    while (w_seq < pf) issue_next_w();
                }
                mbar_wait(stage_free, (k - 1) & 1);          // the previous item's staged tile has left the patch buffers
            } else {
                // first item: the weights do not depend on the previous layer -- request them, then wait for the
                // previous kernel of the stream (programmatic dependent launch) before touching its output
                const int pf = min(kWStages, n_chunks * taps);
                while (w_seq < pf) issue_next_w();
                griddep_wait();
            }
            int seq = 0;
            for (int it = 0; it < n_chunks; ++it, ++g_it) {
                const int ci = c_begin + it, hb = g_it % kHaloBufs;
                mbar_wait(&halo_empty[hb], ((g_it / kHaloBufs) & 1) ^ 1);
                uint8_t *hdst = s_halo + hb * kHaloBufBytes;
                if (elect_one()) {
                    mbar_expect_tx(&halo_full[hb], halo_tx);
                    tma_load_4d(&tmX_hi, &halo_full[hb], hdst, ci * kelems, x0 - pad, y0 - pad, b + p.b_in_off, kEvictNormal);
                    tma_load_4d(&tmX_lo, &halo_full[hb], hdst + p.h_plane_bytes, ci * kelems, x0 - pad, y0 - pad, b + p.b_in_off, kEvictNormal);
                }
                __syncwarp();
                for (int tap = 0; tap < taps; ++tap, ++seq)
                    if (seq == w_seq) issue_next_w();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // Everything below is warp-uniform (kernel parameters, loop counters, shuffled values): descriptors are
        // built with uniform-datapath adds and one elected lane only executes the tcgen05 instructions themselves.
        const uint32_t idesc = umma_idesc_f16(128, N), dhi = umma_desc_hi(p.kbytes);
        const uint32_t kb16 = p.kbytes >> 4, row16 = p.hP * kb16 - (p.ksize - 1) * kb16;
        const uint32_t w16_0 = umma_desc_lo(smem_u32(s_w)), h16_0 = umma_desc_lo(smem_u32(s_halo));
        const uint32_t xl_off = p.h_plane_bytes >> 4;
        const bool k128 = p.kbytes == 128, issue = !(B2T_DBG_BITS(p) & 8);
        int ws = 0, g_it = 0, k = 0, acc_seq = 0;
        uint32_t wphase = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
            int b, y0, x0, cout0, z, c_begin, n_chunks;
            decode_item(item, b, y0, x0, cout0, z, c_begin, n_chunks);
            // accumulation passes: chunks [j*pc, (j+1)*pc) of the item form chain j (k_passes = 1: the whole item)
            int passes = 1, pc = n_chunks;
            if (!kSmall && p.k_passes > 1) { passes = min(p.k_passes, n_chunks); pc = (n_chunks + passes - 1) / passes; }
            // (several passes: one main accumulator, the spare TMEM columns hold the parked sums)
            const int n_main = passes > 1 ? 1 : max(1, min(min(p.n_main, Cfg::kTmemCols / N - 1), n_chunks * taps));
            const uint32_t t_corr = tmem_base + n_main * N;
            const bool tr = B2T_TRACE_PTR(p) && blockIdx.x == 0 && k < 64 && lane == 0;
            if (tr) B2T_TRACE_PTR(p)[k * 8 + 1] = clock64();
            int mi = 0, pass_left = 0;                            // chunks left in the current chain (no division in this loop:
            uint32_t first = 1, am = 0;                           //  one thread issues every MMA of the CTA)
            for (int it = 0; it < n_chunks; ++it, ++g_it) {
                if (pass_left == 0) {                            // a new chain: the epilogue has read the accumulators
                    if (acc_seq > 0) { mbar_wait(acc_empty, (acc_seq - 1) & 1); tc_fence_after(); }
                    if (tr && it == 0) B2T_TRACE_PTR(p)[k * 8 + 2] = clock64();
                    mi = 0; first = 1; am = 0;
                    pass_left = min(pc, n_chunks - it);
                }
                const bool pass_end = --pass_left == 0;
                const int hb = g_it % kHaloBufs;
                mbar_wait(&halo_full[hb], (g_it / kHaloBufs) & 1);
                if (tr && it == 0) B2T_TRACE_PTR(p)[k * 8 + 3] = clock64();
                uint32_t xh = h16_0 + hb * (kHaloBufBytes >> 4);
                int kw = 0;
                for (int tap = 0; tap < taps; ++tap) {
                    mbar_wait(&w_full[ws], wphase);
                    tc_fence_after();
                    const uint32_t wh = w16_0 + ws * (kWStageBytes >> 4), wl = wh + (kWTileBytes >> 4);
                    const uint32_t xl = xh + xl_off;
                    const uint32_t t_main = tmem_base + mi * N;
                    const uint32_t ac = first ^ 1u;
                    const bool last_tap = tap == taps - 1;
                    if (elect_one()) {
                        if (issue) {
                            umma_kstep(t_main, t_corr, wh, wl, xh, xl, dhi, idesc, am, ac);
                            umma_kstep(t_main, t_corr, wh + 2, wl + 2, xh + 2, xl + 2, dhi, idesc, 1u, 1u);
                            if (k128) {
                                umma_kstep(t_main, t_corr, wh + 4, wl + 4, xh + 4, xl + 4, dhi, idesc, 1u, 1u);
                                umma_kstep(t_main, t_corr, wh + 6, wl + 6, xh + 6, xl + 6, dhi, idesc, 1u, 1u);
                            }
                        }
                        umma_commit(&w_empty[ws]);
                        if (last_tap) umma_commit(&halo_empty[hb]);
                        if (last_tap && pass_end) umma_commit(accum_bar);
                    }
                    __syncwarp();
                    first = 0;
                    if (++mi == n_main) { mi = 0; am = 1; }          // every main accumulator has been written once
                    if (++kw == p.ksize) { kw = 0; xh += row16; } else xh += kb16;   // next tap: patch start shifted by (kh*hP + kw) rows
                    if (++ws == kWStages) { ws = 0; wphase ^= 1; }
                }
                if (pass_end) ++acc_seq;
            }
            if (tr) B2T_TRACE_PTR(p)[k * 8 + 7] = clock64();
        }
    } else {
        // ===================== epilogue (the patch buffers are dead once accum_bar fires) =====================
        int k = 0, acc_seq = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
            int b, y0, x0, cout0, z, c_begin, n_chunks;
            decode_item(item, b, y0, x0, cout0, z, c_begin, n_chunks);
            int passes = 1, pc = n_chunks, n_pass = 1;
            if (!kSmall && p.k_passes > 1) {
                passes = min(p.k_passes, n_chunks);
                pc = (n_chunks + passes - 1) / passes;
                n_pass = (n_chunks + pc - 1) / pc;
            }
            const int n_main = passes > 1 ? 1 : max(1, min(min(p.n_main, Cfg::kTmemCols / N - 1), n_chunks * taps));
            const bool tr = B2T_TRACE_PTR(p) && blockIdx.x == 0 && k < 64 && threadIdx.x == 64;
            if (tr) B2T_TRACE_PTR(p)[k * 8 + 4] = clock64();
            if (kSmall || n_pass == 1) {
                mbar_wait(accum_bar, acc_seq & 1);
                ++acc_seq;
                tc_fence_after();
                if (tr) B2T_TRACE_PTR(p)[k * 8 + 5] = clock64();
                halo_epilogue(p, reinterpret_cast<float *>(smem), tmem_base, n_main + 1, N, b, y0, x0, cout0, z, acc_empty);
            } else if constexpr (!kSmall) {
                float keep[2][16];
                float *stage = reinterpret_cast<float *>(smem);
                for (int ps = 0; ps < n_pass; ++ps) {
                    mbar_wait(accum_bar, acc_seq & 1);
                    ++acc_seq;
                    tc_fence_after();
                    if (ps + 1 < n_pass) {
                        halo_pass_accumulate<false>(p, keep, stage, tmem_base, N, cout0, ps == 0);
                        tc_fence_before();
                        asm volatile("bar.sync 1, 256;" ::: "memory");
                        if (threadIdx.x == 64) mbar_arrive(acc_empty);          // the next chain may restart the accumulators
                    } else {
                        if (tr) B2T_TRACE_PTR(p)[k * 8 + 5] = clock64();
                        halo_pass_accumulate<true>(p, keep, stage, tmem_base, N, cout0, false);
                        tc_fence_before();
                        asm volatile("bar.sync 1, 256;" ::: "memory");
                        if (threadIdx.x == 64) mbar_arrive(acc_empty);
                        epilogue_store(p, stage, kStageLd, 4, b, y0, x0, cout0, z, threadIdx.x - 64, kEpiThreads);
                    }
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 64) mbar_arrive(stage_free);
            if (tr) B2T_TRACE_PTR(p)[k * 8 + 6] = clock64();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Persistent variant for the short-K layers (splits == 1, N <= 128): one CTA per SM walks a static round-robin list
// of (pixel tile, cout tile) items.  TMEM holds two accumulator sets (main + correction each) and the epilogue has
// its own staging buffer, so the epilogue of item j overlaps the TMA + MMA main loop of item j+1 and the CTA
// prologue (barrier init, TMEM allocation, descriptor prefetch) is paid once per SM instead of once per tile.
// Shared memory is carved at run time: [2 patch buffers][epilogue stage N x 132 floats][weight ring].  The ring holds
// as many (tap, chunk) weight tiles as fit (p.pw_stages <= 16): enough bytes in flight to cover the L2 latency
// (an SM needs ~80 KB outstanding to pull 42 B/clk), and when ALL of a layer's weight tiles fit and there is a
// single cout tile (conv_2, conv_4) they are loaded once per CTA and stay resident.
constexpr int kPMaxWStages = 16;
constexpr int kPSmemBytes = 227 * 1024 - 1024;

__global__ void __launch_bounds__(kHaloThreads, 1)
conv_halo_persist_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                         const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                         const ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t *tail = smem;                                        // barriers first (fixed offsets)
    uint64_t *patch_full = reinterpret_cast<uint64_t *>(tail);   // [2]
    uint64_t *patch_empty = patch_full + 2;
    uint64_t *acc_full = patch_empty + 2;
    uint64_t *acc_empty = acc_full + 2;
    uint64_t *w_full = acc_empty + 2;                            // [kPMaxWStages]
    uint64_t *w_empty = w_full + kPMaxWStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(w_empty + kPMaxWStages);
    uint8_t *s_patch = smem + 1024;
    float *stage = reinterpret_cast<float *>(s_patch + 2 * p.pw_patch_bytes);
    uint8_t *s_w = reinterpret_cast<uint8_t *>(stage) + p.pw_stage_bytes;
    const int n_ws = p.pw_stages, w_tile = p.pw_tile_bytes, w_plane = w_tile / 2;
    const int kPPatchBytes = p.pw_patch_bytes;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform (see above)
    const int pad = p.ksize >> 1, taps = p.ksize * p.ksize;
    const int N = p.hN;
    const int n_ct = (p.Cout + 127) / 128;
    const int n_items = p.B * p.h_tiles_y * p.h_tiles_x * n_ct;
    const bool resident = n_ct == 1 && taps * p.cin_chunks <= n_ws;   // every weight tile of the layer stays in smem

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX_hi); tma_prefetch_desc(&tmX_lo);
        tma_prefetch_desc(&tmW_hi); tma_prefetch_desc(&tmW_lo);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&patch_full[i], 1); mbar_init(&patch_empty[i], 1);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 1);
        }
        for (int i = 0; i < kPMaxWStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    griddep_launch();

    auto decode_item = [&](int item, int &b, int &y0, int &x0, int &cout0) {
        const int ct = item % n_ct;  item /= n_ct;
        const int tx = item % p.h_tiles_x;  item /= p.h_tiles_x;
        const int ty = item % p.h_tiles_y;
        b = item / p.h_tiles_y;
        y0 = ty * p.hR; x0 = tx * p.hC; cout0 = ct * 128;
    };

    if (warp == 0) {
        // ===================== TMA producer (whole warp runs the uniform loops, one elected lane issues) =====
        const uint32_t patch_tx = 2u * p.h_rows * p.hP * p.kbytes, w_tx = (uint32_t)w_tile;
        const int kelems = p.kbytes / 2;
        int g_chunk = 0, ws = 0, n_loaded = 0;
        uint32_t wphase = 0;
        auto load_w = [&](int ws, int ci, int tap, int cout0) {
            mbar_expect_tx(&w_full[ws], w_tx);
            uint8_t *wdst = s_w + ws * w_tile;
            const int kcoord = (tap * p.cin_chunks + ci) * kelems;
            tma_load_2d(&tmW_hi, &w_full[ws], wdst, kcoord, cout0, kEvictLast);
            tma_load_2d(&tmW_lo, &w_full[ws], wdst + w_plane, kcoord, cout0, kEvictLast);
        };
        if (resident && elect_one())
            for (int ci = 0; ci < p.cin_chunks; ++ci)
                for (int tap = 0; tap < taps; ++tap) load_w(ci * taps + tap, ci, tap, 0);
        __syncwarp();
        griddep_wait();          // the activations come from the previous kernel of the stream (weights do not)
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int b, y0, x0, cout0;
            decode_item(item, b, y0, x0, cout0);
            for (int ci = 0; ci < p.cin_chunks; ++ci, ++g_chunk) {
                const int hb = g_chunk & 1;
                mbar_wait(&patch_empty[hb], ((g_chunk >> 1) & 1) ^ 1);
                if (B2T_TRACE_PTR(p) && blockIdx.x == 0 && ci == 0 && g_chunk < 64 && lane == 0) B2T_TRACE_PTR(p)[g_chunk * 8 + 0] = clock64();
                uint8_t *hdst = s_patch + hb * kPPatchBytes;
                const bool skip_x = (B2T_DBG_BITS(p) & 2) && g_chunk >= 2;
                if (elect_one()) {
                    if (skip_x) mbar_arrive(&patch_full[hb]);
                    else {
                        mbar_expect_tx(&patch_full[hb], patch_tx);
                        tma_load_4d(&tmX_hi, &patch_full[hb], hdst, ci * kelems, x0 - pad, y0 - pad, b + p.b_in_off, kEvictNormal);
                        tma_load_4d(&tmX_lo, &patch_full[hb], hdst + p.h_plane_bytes, ci * kelems, x0 - pad, y0 - pad, b + p.b_in_off, kEvictNormal);
                    }
                }
                __syncwarp();
                if (resident) continue;
                for (int tap = 0; tap < taps; ++tap) {
                    mbar_wait(&w_empty[ws], wphase ^ 1);
                    const bool skip_w = (B2T_DBG_BITS(p) & 1) && n_loaded >= n_ws;
                    if (elect_one()) {
                        if (skip_w) mbar_arrive(&w_full[ws]);
                        else load_w(ws, ci, tap, cout0);
                    }
                    __syncwarp();
                    ++n_loaded;
                    if (++ws == n_ws) { ws = 0; wphase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform arithmetic, lane 0 executes the tcgen05 instructions) =====
        const uint32_t idesc = umma_idesc_f16(128, N), dhi = umma_desc_hi(p.kbytes);
        const uint32_t kb16 = p.kbytes >> 4, row16 = p.hP * kb16 - (p.ksize - 1) * kb16;
        const uint32_t w16_0 = umma_desc_lo(smem_u32(s_w)), p16_0 = umma_desc_lo(smem_u32(s_patch));
        const uint32_t xl_off = p.h_plane_bytes >> 4, wl_off = w_plane >> 4, wt16 = w_tile >> 4, pb16 = kPPatchBytes >> 4;
        const bool k128 = p.kbytes == 128, issue = !(B2T_DBG_BITS(p) & 8);
        int g_chunk = 0, j = 0, ws = 0;
        uint32_t wphase = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++j) {
            const int ab = j & 1;
            if (B2T_TRACE_PTR(p) && blockIdx.x == 0 && j < 64 && lane == 0) B2T_TRACE_PTR(p)[j * 8 + 1] = clock64();
            mbar_wait(&acc_empty[ab], ((j >> 1) & 1) ^ 1);       // the epilogue has drained this accumulator set
            tc_fence_after();
            if (B2T_TRACE_PTR(p) && blockIdx.x == 0 && j < 64 && lane == 0) B2T_TRACE_PTR(p)[j * 8 + 2] = clock64();
            const uint32_t t_main = tmem_base + ab * 2 * N, t_corr = t_main + N;
            uint32_t acc = 0;                                     // first MMA of the item overwrites
            if (resident) ws = 0;
            for (int ci = 0; ci < p.cin_chunks; ++ci, ++g_chunk) {
                const int hb = g_chunk & 1;
                mbar_wait(&patch_full[hb], (g_chunk >> 1) & 1);
                if (B2T_TRACE_PTR(p) && blockIdx.x == 0 && j < 64 && lane == 0 && ci == 0) B2T_TRACE_PTR(p)[j * 8 + 3] = clock64();
                uint32_t xh = p16_0 + hb * pb16;
                int kw = 0;
                for (int tap = 0; tap < taps; ++tap) {
                    mbar_wait(&w_full[ws], resident ? 0 : wphase);
                    tc_fence_after();
                    const uint32_t wh = w16_0 + ws * wt16, wl = wh + wl_off, xl = xh + xl_off;
                    const bool last_tap = tap == taps - 1;
                    if (elect_one()) {
                        if (issue) {
                            umma_kstep(t_main, t_corr, wh, wl, xh, xl, dhi, idesc, acc, acc);
                            umma_kstep(t_main, t_corr, wh + 2, wl + 2, xh + 2, xl + 2, dhi, idesc, 1u, 1u);
                            if (k128) {
                                umma_kstep(t_main, t_corr, wh + 4, wl + 4, xh + 4, xl + 4, dhi, idesc, 1u, 1u);
                                umma_kstep(t_main, t_corr, wh + 6, wl + 6, xh + 6, xl + 6, dhi, idesc, 1u, 1u);
                            }
                        }
                        if (!resident) umma_commit(&w_empty[ws]);
                        if (last_tap) umma_commit(&patch_empty[hb]);
                        if (last_tap && ci == p.cin_chunks - 1) umma_commit(&acc_full[ab]);
                    }
                    __syncwarp();
                    acc = 1;
                    if (++kw == p.ksize) { kw = 0; xh += row16; } else xh += kb16;
                    if (++ws == n_ws && !resident) { ws = 0; wphase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue (overlaps the next item's main loop) =====================
        int j = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++j) {
            int b, y0, x0, cout0;
            decode_item(item, b, y0, x0, cout0);
            const int ab = j & 1;
            if (B2T_TRACE_PTR(p) && blockIdx.x == 0 && j < 64 && threadIdx.x == 64) B2T_TRACE_PTR(p)[j * 8 + 4] = clock64();
            mbar_wait(&acc_full[ab], (j >> 1) & 1);
            tc_fence_after();
            if (B2T_TRACE_PTR(p) && blockIdx.x == 0 && j < 64 && threadIdx.x == 64) B2T_TRACE_PTR(p)[j * 8 + 5] = clock64();
            halo_epilogue(p, stage, tmem_base + ab * 2 * N, 2, N, b, y0, x0, cout0, 0, &acc_empty[ab]);
            asm volatile("bar.sync 1, 256;" ::: "memory");       // stage buffer free for the next item
            if (B2T_TRACE_PTR(p) && blockIdx.x == 0 && j < 64 && threadIdx.x == 64) B2T_TRACE_PTR(p)[j * 8 + 6] = clock64();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// conv_chain_kernel -- a RUN of consecutive conv layers in ONE persistent cooperative launch, for small batches
// (one stream frame by frame, BASELINE configs 2 and 4: 1..8 frames per step).
//
// At batch 1 a layer is a few microseconds of MMAs on each SM; launched one kernel per layer (plus a split-K finish
// kernel) the step is the sum of 40 kernel boundaries: grid drain, launch, barrier init, TMEM allocation, descriptor
// fetch, first-TMA latency -- 20 us per layer against 2-5 us of work.  Here one grid of gridDim.x = #SM CTAs walks the
// layer list (kernel parameters carry every layer's ConvParams and TMA descriptors): the main loop of each layer is
// conv_halo_kernel<big>'s (same roles, barriers, TMEM and weight ring, alive across layers), and layers are
// separated by a grid barrier (monotonic arrival counter in global memory, release/acquire at gpu scope):
//   * the producer warp requests the first WEIGHT tiles of layer L+1 before it waits for layer L to complete (weights
//     do not depend on activations), so the weight stream from HBM -- the roofline of this regime -- keeps flowing
//     across the barrier;
//   * a split-K layer is finished in place: after the barrier that says "all partials are written" the 256 epilogue
//     threads of every CTA reduce a slice of the partial buffer in fixed order (splitk_finish_range, deterministic)
//     and a second barrier releases the next layer -- no second kernel.
// Activations written by other SMs' generic stores are read by TMA (async proxy) after an acquire of the counter and
// a fence.proxy.async; partials are read with ld.global.cg (see splitk_finish_range).
// The grid barrier needs every CTA resident: the launch is cooperative (refused otherwise), one CTA per SM.
constexpr int kChainMaxLayers = 22;               // conv_2 .. conv_23
struct alignas(64) ChainLayer {
    CUtensorMap x_hi, x_lo, w_hi, w_lo;
    ConvParams p;
};
struct alignas(64) ChainParams {
    ChainLayer L[kChainMaxLayers];
    unsigned int *counter;        // zeroed by the host before the launch
    int n_layers;
};
static_assert(sizeof(ChainParams) <= 32000, "kernel parameter space (32764 bytes on sm_70+ with CUDA >= 12.1)");

// Epilogue of a split-K item: raw fp32 sums go straight from registers to the partial buffer
// [split][pixel][channel].  A thread owns one output channel (its TMEM lane) and half of the tile's pixel columns; the
// 32 lanes of a warp are 32 consecutive channels, so every store instruction writes 128 contiguous bytes of one
// pixel -- no shared-memory staging and no second pass.
__device__ __forceinline__ void chain_partial_epilogue(const ConvParams &p, uint32_t tmem_acc, int n_acc, int N, int b, int y0,
                                                       int x0, int cout0, int zsplit, uint64_t *acc_free) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int ch = cout0 + q * 32 + lane;
    const bool ch_ok = ch < p.ldp;
    const uint32_t lane_addr = tmem_acc + (uint32_t(q * 32) << 16);
    const int rows_valid = min(p.hR, p.H - y0), cols_valid = min(p.hC, p.W - x0);
    const long long mtot = (long long)p.B * p.H * p.W;
    float *base = p.partial + ((long long)zsplit * mtot + ((long long)b * p.H + y0) * p.W + x0) * p.ldp + ch;
#pragma unroll 1
    for (int n0 = half * 16; n0 < N; n0 += 32) {
        uint32_t a[16], c2[16];
        float v[16];
        tmem_ld16(lane_addr + n0, a);
        tmem_ld16(lane_addr + (n_acc - 1) * N + n0, c2);   // the correction accumulator is the last one
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(a[i]) + __uint_as_float(c2[i]);
#pragma unroll 1
        for (int m = 1; m < n_acc - 1; ++m) {
            tmem_ld16(lane_addr + m * N + n0, a);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += __uint_as_float(a[i]);
        }
        int r = small_div(n0, p.hP), c = n0 - r * p.hP;     // column n = r*hP + c of the tile
        float *ptr = base + ((long long)r * p.W + c) * p.ldp;
        const long long wrap = (long long)(p.W - p.hP + 1) * p.ldp;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (ch_ok && r < rows_valid && c < cols_valid) *ptr = v[i];
            if (++c == p.hP) { c = 0; ++r; ptr += wrap; } else ptr += p.ldp;
        }
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (acc_free && threadIdx.x == 64) mbar_arrive(acc_free);
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void grid_arrive(unsigned int *ctr) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
// one thread: wait until `target` arrivals have been counted; traps instead of hanging the device if that never happens
__device__ __forceinline__ void grid_wait(const unsigned int *ctr, unsigned int target) {
    const long long t0 = clock64();
    while (ld_acquire_gpu(ctr) < target)
        if (clock64() - t0 > (1ll << 31)) __trap();
}

__global__ void __launch_bounds__(kHaloThreads, 1) conv_chain_kernel(const __grid_constant__ ChainParams cp) {
    using Cfg = HaloCfg<false>;
    constexpr int kHaloBufs = Cfg::kHaloBufs, kHaloBufBytes = Cfg::kHaloBufBytes, kWStages = Cfg::kWStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *s_halo = smem;
    uint8_t *s_w = smem + kHaloBufs * kHaloBufBytes;
    uint8_t *tail = s_w + kWStages * kWStageBytes;
    uint64_t *halo_full = reinterpret_cast<uint64_t *>(tail);   // [2]
    uint64_t *halo_empty = halo_full + 2;                       // [2]
    uint64_t *w_full = halo_empty + 2;                          // [kWStages]
    uint64_t *w_empty = w_full + kWStages;                      // [kWStages]
    uint64_t *accum_bar = w_empty + kWStages;
    uint64_t *acc_empty = accum_bar + 1;
    uint64_t *stage_free = acc_empty + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(stage_free + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform (see above)
    const unsigned int G = gridDim.x;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < kHaloBufs; ++i) { mbar_init(&halo_full[i], 1); mbar_init(&halo_empty[i], 1); }
        for (int i = 0; i < kWStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        mbar_init(accum_bar, 1); mbar_init(acc_empty, 1); mbar_init(stage_free, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    // item = (pixel tile, cout tile, K split) of one layer; every role walks the same list
    // The K dimension is split in units of one (channel chunk, tap) pair = 12 MMAs, chunk-major: finer than whole
    // chunks, so that a 13x13 layer at batch 1 (8 cout tiles x 72 units) spreads over every SM.  Split z owns units
    // [u0, u0 + n_units).
    struct Geo { int gx, gy, n_items, units; };
    auto geometry = [&](const ConvParams &p) {
        Geo g;
        g.gx = p.B * p.h_tiles_x * p.h_tiles_y; g.gy = (p.Cout + 127) >> 7;
        g.n_items = g.gx * g.gy * p.splits;
        g.units = p.cin_chunks * p.ksize * p.ksize;
        return g;
    };
    auto decode_item = [&](const ConvParams &p, const Geo &g, int item, int &b, int &y0, int &x0, int &cout0, int &z,
                           int &u0, int &n_units) {
        int t = item % g.gx;
        const int rest = item / g.gx;
        const int tx = t % p.h_tiles_x;  t /= p.h_tiles_x;
        const int ty = t % p.h_tiles_y;
        b = t / p.h_tiles_y;
        x0 = tx * p.hC; y0 = ty * p.hR;
        cout0 = (rest % g.gy) * 128;
        z = rest / g.gy;
        u0 = z * p.k_per_units;
        n_units = min(g.units, u0 + p.k_per_units) - u0;                // host guarantees >= 1
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        int ws = 0, g_it = 0, kk = 0;
        uint32_t wphase = 0;
        bool prev_hits_w = false;          // the previous item's staged tile reaches into the weight ring
        unsigned int ready = 0;            // arrivals that complete the previous layer
        for (int L = 0; L < cp.n_layers; ++L) {
            const ConvParams &p = cp.L[L].p;
            const CUtensorMap *tmX_hi = &cp.L[L].x_hi, *tmX_lo = &cp.L[L].x_lo, *tmW_hi = &cp.L[L].w_hi, *tmW_lo = &cp.L[L].w_lo;
            const Geo g = geometry(p);
            const int pad = p.ksize >> 1, taps = p.ksize * p.ksize, N = p.hN;
            const bool stage_hits_w = N * kStageLd * 4 > kHaloBufs * kHaloBufBytes;
            const uint32_t halo_tx = 2u * p.h_rows * p.hP * p.kbytes, w_tx = 2u * 128u * p.kbytes;
            const int kelems = p.kbytes / 2;
            bool input_ready = false;
            if (lane == 0 && (int)blockIdx.x < g.n_items) {
                tma_prefetch_desc(tmX_hi); tma_prefetch_desc(tmX_lo);
                tma_prefetch_desc(tmW_hi); tma_prefetch_desc(tmW_lo);
            }
            for (int item = blockIdx.x; item < g.n_items; item += G, ++kk) {
                int b, y0, x0, cout0, z, u0, n_units;
                decode_item(p, g, item, b, y0, x0, cout0, z, u0, n_units);
                // weight tiles are requested strictly in unit order; (w_ci, w_tap) is the next one
                int w_ci = u0 / taps, w_tap = u0 - w_ci * taps, w_seq = 0;
                auto issue_next_w = [&]() {
                    mbar_wait(&w_empty[ws], wphase ^ 1);
                    uint8_t *wdst = s_w + ws * kWStageBytes;
                    const int kcoord = (w_tap * p.cin_chunks + w_ci) * kelems;
                    if (elect_one()) {
                        mbar_expect_tx(&w_full[ws], w_tx);
                        tma_load_2d(tmW_hi, &w_full[ws], wdst, kcoord, cout0, kEvictLast);
                        tma_load_2d(tmW_lo, &w_full[ws], wdst + kWTileBytes, kcoord, cout0, kEvictLast);
                    }
                    __syncwarp();
                    ++w_seq;
                    if (++w_tap == taps) { w_tap = 0; ++w_ci; }
                    if (++ws == kWStages) { ws = 0; wphase ^= 1; }
                };
                // weights first: they do not depend on the previous layer (nor on the previous item's staged tile,
                // unless that tile reaches into the weight ring)
                if (kk == 0 || !prev_hits_w) {
                    const int pf = min(kWStages, n_units);
                    while (w_seq < pf) issue_next_w();
                }
                if (kk > 0) mbar_wait(stage_free, (kk - 1) & 1);     // the previous item's staged tile has left the patch buffers
                if (!input_ready) {
                    if (L > 0) {                                     // the previous layer is complete on every SM
                        if (lane == 0) grid_wait(cp.counter, ready);
                        __syncwarp();
                        asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy stores -> TMA reads
                    }
                    if (B2T_TRACE_PTR(cp.L[0].p) && blockIdx.x == 0 && lane == 0) B2T_TRACE_PTR(cp.L[0].p)[L * 8 + 0] = clock64();
                    input_ready = true;
                }
                int seq = 0;
                for (int u = u0; u < u0 + n_units; ++g_it) {
                    const int ci = u / taps, t0 = u - ci * taps, t1 = min(taps, t0 + (u0 + n_units - u));
                    const int hb = g_it % kHaloBufs;
                    mbar_wait(&halo_empty[hb], ((g_it / kHaloBufs) & 1) ^ 1);
                    uint8_t *hdst = s_halo + hb * kHaloBufBytes;
                    if (elect_one()) {
                        mbar_expect_tx(&halo_full[hb], halo_tx);
                        tma_load_4d(tmX_hi, &halo_full[hb], hdst, ci * kelems, x0 - pad, y0 - pad, b, kEvictNormal);
                        tma_load_4d(tmX_lo, &halo_full[hb], hdst + p.h_plane_bytes, ci * kelems, x0 - pad, y0 - pad, b, kEvictNormal);
                    }
                    __syncwarp();
                    for (int tap = t0; tap < t1; ++tap, ++seq)
                        if (seq == w_seq) issue_next_w();
                    u += t1 - t0;
                }
                // a split layer's epilogue writes its partials straight from registers: nothing is staged
                prev_hits_w = stage_hits_w && p.splits == 1;
            }
            ready += G * (p.splits > 1 ? 2u : 1u);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform arithmetic, one elected lane issues) =====================
        const uint32_t w16_0 = umma_desc_lo(smem_u32(s_w)), h16_0 = umma_desc_lo(smem_u32(s_halo));
        int ws = 0, g_it = 0, kk = 0;
        uint32_t wphase = 0;
        for (int L = 0; L < cp.n_layers; ++L) {
            const ConvParams &p = cp.L[L].p;
            const Geo g = geometry(p);
            const int taps = p.ksize * p.ksize, N = p.hN;
            const uint32_t idesc = umma_idesc_f16(128, N), dhi = umma_desc_hi(p.kbytes);
            const uint32_t kb16 = p.kbytes >> 4, row16 = p.hP * kb16 - (p.ksize - 1) * kb16;
            const uint32_t xl_off = p.h_plane_bytes >> 4;
            const bool k128 = p.kbytes == 128;
            for (int item = blockIdx.x; item < g.n_items; item += G, ++kk) {
                int b, y0, x0, cout0, z, u0, n_units;
                decode_item(p, g, item, b, y0, x0, cout0, z, u0, n_units);
                const int n_main = max(1, min(min(p.n_main, Cfg::kTmemCols / N - 1), n_units));
                const uint32_t t_corr = tmem_base + n_main * N;
                if (kk > 0) { mbar_wait(acc_empty, (kk - 1) & 1); tc_fence_after(); }   // the epilogue has read the accumulators
                int mi = 0;
                uint32_t first = 1, am = 0;
                for (int u = u0; u < u0 + n_units; ++g_it) {
                    const int ci = u / taps, t0 = u - ci * taps, t1 = min(taps, t0 + (u0 + n_units - u));
                    const int hb = g_it % kHaloBufs;
                    mbar_wait(&halo_full[hb], (g_it / kHaloBufs) & 1);
                    if (B2T_TRACE_PTR(cp.L[0].p) && blockIdx.x == 0 && lane == 0 && u == u0 && item == (int)blockIdx.x)
                        B2T_TRACE_PTR(cp.L[0].p)[L * 8 + 1] = clock64();
                    // tap t reads the patch shifted by (kh*hP + kw) rows
                    const int kh0 = t0 / p.ksize;
                    int kw = t0 - kh0 * p.ksize;
                    uint32_t xh = h16_0 + hb * (kHaloBufBytes >> 4) + (kh0 * p.hP + kw) * kb16;
                    for (int tap = t0; tap < t1; ++tap) {
                        mbar_wait(&w_full[ws], wphase);
                        tc_fence_after();
                        const uint32_t wh = w16_0 + ws * (kWStageBytes >> 4), wl = wh + (kWTileBytes >> 4);
                        const uint32_t xl = xh + xl_off;
                        const uint32_t t_main = tmem_base + mi * N;
                        const uint32_t ac = first ^ 1u;
                        const bool last_tap = tap == t1 - 1;
                        if (B2T_TRACE_PTR(cp.L[0].p) && blockIdx.x == 0 && lane == 0 && u == u0 && tap == t0 && item == (int)blockIdx.x)
                            B2T_TRACE_PTR(cp.L[0].p)[L * 8 + 2] = clock64();
                        if (elect_one()) {
                            umma_kstep(t_main, t_corr, wh, wl, xh, xl, dhi, idesc, am, ac);
                            umma_kstep(t_main, t_corr, wh + 2, wl + 2, xh + 2, xl + 2, dhi, idesc, 1u, 1u);
                            if (k128) {
                                umma_kstep(t_main, t_corr, wh + 4, wl + 4, xh + 4, xl + 4, dhi, idesc, 1u, 1u);
                                umma_kstep(t_main, t_corr, wh + 6, wl + 6, xh + 6, xl + 6, dhi, idesc, 1u, 1u);
                            }
                            umma_commit(&w_empty[ws]);
                            if (last_tap) umma_commit(&halo_empty[hb]);
                            if (last_tap && u + (t1 - t0) == u0 + n_units) umma_commit(accum_bar);
                        }
                        __syncwarp();
                        first = 0;
                        if (++mi == n_main) { mi = 0; am = 1; }
                        if (++kw == p.ksize) { kw = 0; xh += row16; } else xh += kb16;
                        if (++ws == kWStages) { ws = 0; wphase ^= 1; }
                    }
                    u += t1 - t0;
                }
            }
        }
    } else {
        // ===================== epilogue + grid barrier + split-K finish =====================
        const int et = threadIdx.x - 64;
        int kk = 0;
        unsigned int done = 0;
        long long *trace = B2T_TRACE_PTR(cp.L[0].p);            // developer builds: per-layer clock stamps of CTA 0
        const bool tr = trace && blockIdx.x == 0 && et == 0;
        for (int L = 0; L < cp.n_layers; ++L) {
            const ConvParams &p = cp.L[L].p;
            const Geo g = geometry(p);
            const int N = p.hN;
            bool first = true;
            for (int item = blockIdx.x; item < g.n_items; item += G, ++kk) {
                int b, y0, x0, cout0, z, u0, n_units;
                decode_item(p, g, item, b, y0, x0, cout0, z, u0, n_units);
                const int n_main = max(1, min(min(p.n_main, Cfg::kTmemCols / N - 1), n_units));
                mbar_wait(accum_bar, kk & 1);
                tc_fence_after();
                if (tr && first) trace[L * 8 + 3] = clock64();
                first = false;
                if (p.splits > 1)
                    chain_partial_epilogue(p, tmem_base, n_main + 1, N, b, y0, x0, cout0, z, acc_empty);
                else
                    halo_epilogue(p, reinterpret_cast<float *>(smem), tmem_base, n_main + 1, N, b, y0, x0, cout0, z, acc_empty);
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (threadIdx.x == 64) mbar_arrive(stage_free);
            }
            // This CTA's part of layer L is in global memory: arrive on the grid barrier, then WAIT for the phase to
            // complete before arriving anywhere else -- a CTA without items in the next layer would otherwise count
            // for that layer's phase while another CTA is still in this one, and the monotonic counter would release
            // waiters early.
            if (tr) trace[L * 8 + 4] = clock64();
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");
            done += G;
            if (et == 0) { grid_arrive(cp.counter); grid_wait(cp.counter, done); }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tr) trace[L * 8 + 5] = clock64();
            if (p.splits > 1) {                                   // every CTA's partials are written: finish a slice
                // work items are dealt to WARPS round-robin over the CTAs (warp w of CTA b = global warp w*G + b), 32
                // consecutive items per warp: a layer with fewer items than threads still loads every SM's L2 port
                splitk_finish_range(p, ((long long)(et >> 5) * G + blockIdx.x) * 32 + (et & 31), (long long)G * kEpiThreads);
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
                done += G;
                if (et == 0) { grid_arrive(cp.counter); grid_wait(cp.counter, done); }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            if (tr) trace[L * 8 + 6] = clock64();
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
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(conv_halo_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmemBytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloCfg<false>::kSmemBytes);
    return (int)e;
}

// ---- conv_chain_kernel host side: the parameter block is assembled layer by layer by api.cu
struct ChainBuilder { ChainParams cp; };
ChainBuilder *chain_new(unsigned int *counter) {
    ChainBuilder *b = new ChainBuilder();
    memset(&b->cp, 0, sizeof b->cp);
    b->cp.counter = counter;
    return b;
}
void chain_free(ChainBuilder *b) { delete b; }
int chain_add(ChainBuilder *b, const CUtensorMap &x_hi, const CUtensorMap &x_lo, const CUtensorMap &w_hi,
              const CUtensorMap &w_lo, const ConvParams &p) {
    if (b->cp.n_layers >= kChainMaxLayers) return -1;
    ChainLayer &l = b->cp.L[b->cp.n_layers++];
    l.x_hi = x_hi; l.x_lo = x_lo; l.w_hi = w_hi; l.w_lo = w_lo; l.p = p;
    return 0;
}
int chain_layers(const ChainBuilder *b) { return b->cp.n_layers; }
int launch_conv_chain(int n_sm, const ChainBuilder *b, cudaStream_t st) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv_chain_kernel, kHaloThreads,
                                                                  HaloCfg<false>::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) return (int)cudaErrorCooperativeLaunchTooLarge;
    e = cudaMemsetAsync(b->cp.counter, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_sm); cfg.blockDim = dim3(kHaloThreads); cfg.dynamicSmemBytes = HaloCfg<false>::kSmemBytes; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, conv_chain_kernel, b->cp);
}

int launch_conv_halo_persist(int n_sm, const CUtensorMap &x_hi, const CUtensorMap &x_lo, const CUtensorMap &w_hi,
                             const CUtensorMap &w_lo, const ConvParams &p, cudaStream_t st) {
    const int items = p.B * p.h_tiles_x * p.h_tiles_y * ((p.Cout + 127) / 128);
    const int smem = 1024 /*align*/ + 1024 /*barriers*/ + 2 * p.pw_patch_bytes + p.pw_stage_bytes + p.pw_stages * p.pw_tile_bytes;
    return (int)launch_pdl(conv_halo_persist_kernel, dim3(items < n_sm ? items : n_sm), dim3(kHaloThreads), smem, st, x_hi, x_lo, w_hi, w_lo, p);
}

int launch_conv_halo(int n_sm, bool small, const CUtensorMap &x_hi, const CUtensorMap &x_lo, const CUtensorMap &w_hi,
                     const CUtensorMap &w_lo, const ConvParams &p, cudaStream_t st) {
    const int items = p.B * p.h_tiles_x * p.h_tiles_y * ((p.Cout + 127) / 128) * p.splits;
    const int cap = n_sm * (small ? 2 : 1);
    const int grid = items < cap ? items : cap;
    if (small)
        return (int)launch_pdl(conv_halo_kernel<true>, dim3(grid), dim3(kHaloThreads), HaloCfg<true>::kSmemBytes, st, x_hi, x_lo, w_hi, w_lo, p);
    return (int)launch_pdl(conv_halo_kernel<false>, dim3(grid), dim3(kHaloThreads), HaloCfg<false>::kSmemBytes, st, x_hi, x_lo, w_hi, w_lo, p);
}

}  // namespace b2t
