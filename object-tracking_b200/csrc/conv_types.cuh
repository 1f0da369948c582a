// Shared parameter blocks and the output "emit" routine used by every conv epilogue.
//
// Activation layout in HBM ("split planes"): an activation tensor is two fp16 NHWC planes, hi and lo,
// with value = hi + lo (22 significant bits).  hi plane at `hi`, lo plane at `hi + plane_stride`.
// A destination may be a channel slice of a wider buffer (pix_stride > C, ch_off > 0): this is how the
// route/concat layers (KerasYOLO.py:391, MultiObjDetTracker.py:175) are written in place.
#pragma once
#include "ptx.cuh"

namespace b2t {

enum { DEST_PLAIN = 0, DEST_S2D_TF = 1 };   // darknet's reorg ordering is a separate gather kernel (conv_misc.cu)

struct Dest {
    op_t *hi;      // fp16 hi plane base (NULL = none)
    long long plane_stride; // elements between the hi and lo plane
    int pix_stride_b;       // channels per pixel of the split-plane buffer
    int ch_off_b;           // first channel of this tensor inside the split-plane buffer
    float *f32;             // fp32 NHWC base (NULL = none)
    int pix_stride_f;
    int ch_off_f;
    int accumulate_f;       // fp32 dest: add to what is there (ConvLSTM recurrent term)
    long long img_stride_f; // fp32 dest: pixels between consecutive images (0 = H*W, dense); the recurrent conv of
                            // time step t writes stream s to frame s*T + t of the (S*T)-frame gate buffer
    int H, W;               // spatial dims of the SOURCE tensor the coordinates refer to
    int mode;               // DEST_*
};

struct ConvParams {
    int B, H, W;            // output (= input) spatial size, stride-1 'same' conv
    int ksize;              // 1 or 3
    int cin_chunks;         // padded Cin / 64
    int Cout;
    int TW, TH;             // M tile = TW x TH pixels (TW*TH = 128)
    int tiles_x, tiles_y;
    int splits;             // split-K factor (gridDim.z)
    int chunks_total;       // ksize*ksize*cin_chunks
    int ldp;                // leading dim of the fp32 partial buffer (Cout rounded up to 32)
    int act;                // 1 = LeakyReLU(0.1), 0 = linear
    int pool;               // 1 = also emit the 2x2/2 max-pooled tensor to `pout`
    int n_main;             // TMEM accumulators for the hi*hi products (1..3), + 1 for the corrections
    const float *scale;     // per-Cout multiplier (folded BN) ...
    const float *bias;      // ... and offset (or conv bias)
    float *partial;         // [splits][B*H*W][ldp] fp32 raw accumulators (split-K or SIMT engine)
    Dest out;               // full-resolution destination (hi and/or f32 may be NULL)
    Dest pout;              // pooled destination
    // ---- halo engine (conv_halo.cu): pixel tile = hR rows x hC columns of one image, held in shared memory as
    // an (h_rows x hP)-pixel patch with its halo; MMA N = hN = round_up(hR*hP, 16)
    int hC, hP, hR, hN;
    int h_rows;             // patch rows loaded per channel chunk (hR + 2*pad [+1 when the row wrap trick is used])
    int h_plane_bytes;      // bytes between the hi and lo patch in shared memory
    int h_tiles_x, h_tiles_y;
    int kbytes;             // bytes of one K-chunk row: 128 (64 channels, SWIZZLE_128B) or 64 (32 channels, SWIZZLE_64B)
    // persistent variant: run-time shared-memory carve-up
    int pw_patch_bytes;     // one patch buffer (hi + lo)
    int pw_stage_bytes;     // epilogue stage
    int pw_tile_bytes;      // one weight tile (hi + lo) = 2 * w_rows * kbytes
    int pw_stages;          // weight tiles that fit in the ring (<= 16)
    // pixel-major variant (conv_pm.cu): pixels on the MMA's M dimension, all weight tiles resident in shared memory
    int pm_mode;            // 0 = fp16 (hi,lo) activations, swizzled K chunks; 1 = conv_1: fp16 integer frame, 8 ch / pixel
    int pm_n;               // MMA N = Cout rounded up to 16 (<= 128)
    int pm_tmem_cols;       // power of two >= 4 * pm_n
    int pm_stage_ld;        // floats per staged pixel (pm_n + 4)
    int pm_glog;            // log2 of the 8-channel groups per pixel in the store phase
    const unsigned char *pm_w;   // mode 1: packed weights [kh][plane][32 cout x 32 k, 128-byte core matrices]
    int pm_w_bytes;
    unsigned int magic_tx, magic_ty;   // ceil(2^32 / h_tiles_x), ceil(2^32 / h_tiles_y) (0 when the divisor is 1): fast item decode
    int b_in_off;           // added to the image index of the INPUT view only (state slot of the first stream)
    int k_passes;           // conv_halo_kernel<big>: accumulation passes per item (>= 1): the K range of an item is cut into
                            // k_passes chains, their fp32 sums are added in registers (no split-K partials in HBM)
    int k_per_units;        // conv_chain_kernel: (chunk, tap) units per K split, chunk-major (finer than whole chunks)
#ifdef B2T_DEV              // developer builds only (make DEV=1); the release library has neither field nor the code behind them
    int dbg;                // experiments: 1 = weight TMA only for the first ring pass, 2 = patch TMA only for the
                            // first buffers, 4 = no epilogue stores, 8 = no MMAs (results are wrong)
    long long *trace;       // per-item clock64 stamps of CTA 0
#endif
};

#ifdef B2T_DEV
#define B2T_DBG_BITS(p) ((p).dbg)
#define B2T_TRACE_PTR(p) ((p).trace)
#else
#define B2T_DBG_BITS(p) 0
#define B2T_TRACE_PTR(p) (static_cast<long long *>(nullptr))
#endif

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : 0.1f * v; }

// n / d for 0 <= n < 2^20, 1 <= d < 2^12, exact: the quotient of (n + 0.5) / d is at least 0.5 / d away from an integer,
// far more than the error of the approximate fp32 division
__device__ __forceinline__ int small_div(int n, int d) { return __float2int_rz(__fdividef((float)n + 0.5f, (float)d)); }

// Write up to 8 consecutive channels [c, c+8) of source pixel (b,y,x).  Cout = channels of the source tensor.
__device__ __forceinline__ void emit8(const Dest &d, int b, int y, int x, int c, int Cout, const float (&v)[8]) {
    const int nvalid = min(8, Cout - c);
    if (nvalid <= 0) return;
    long long pix;
    int cc = c;
    if (d.mode == DEST_S2D_TF) {  // tf.space_to_depth(2): out[b,y/2,x/2,((y&1)*2+(x&1))*C + c]
        pix = ((long long)b * (d.H / 2) + (y >> 1)) * (d.W / 2) + (x >> 1);
        cc = ((y & 1) * 2 + (x & 1)) * Cout + c;
    } else {
        pix = ((long long)b * d.H + y) * d.W + x;
    }
    if (d.hi) {
        op_t *p = d.hi + pix * d.pix_stride_b + d.ch_off_b + cc;
        __align__(16) op_t h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split_f16(v[i], h[i], l[i]);
        if (nvalid == 8 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && ((d.plane_stride & 7) == 0)) {
            *reinterpret_cast<uint4 *>(p) = *reinterpret_cast<const uint4 *>(h);
            *reinterpret_cast<uint4 *>(p + d.plane_stride) = *reinterpret_cast<const uint4 *>(l);
        } else {
            for (int i = 0; i < nvalid; ++i) {
                p[i] = h[i];
                p[d.plane_stride + i] = l[i];
            }
        }
    }
    if (d.f32) {
        if (d.img_stride_f) pix = (long long)b * d.img_stride_f + (long long)y * d.W + x;   // plain mode only
        float *p = d.f32 + pix * d.pix_stride_f + d.ch_off_f + cc;
        if (nvalid == 8 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
            float4 a = make_float4(v[0], v[1], v[2], v[3]), bq = make_float4(v[4], v[5], v[6], v[7]);
            if (d.accumulate_f) {
                const float4 o0 = reinterpret_cast<const float4 *>(p)[0], o1 = reinterpret_cast<const float4 *>(p)[1];
                a.x += o0.x; a.y += o0.y; a.z += o0.z; a.w += o0.w;
                bq.x += o1.x; bq.y += o1.y; bq.z += o1.z; bq.w += o1.w;
            }
            reinterpret_cast<float4 *>(p)[0] = a;
            reinterpret_cast<float4 *>(p)[1] = bq;
        } else {
            for (int i = 0; i < nvalid; ++i) p[i] = d.accumulate_f ? p[i] + v[i] : v[i];
        }
    }
}

// Store phase shared by the conv epilogues: the tile sits in shared memory as stage[pixel n = r*hP + c][channel]
// (ld floats per pixel); the epilogue threads walk (pixel, 8-channel group) items and
// write 16-byte pieces: raw fp32 split-K partials, or hi/lo split planes with optional 2x2 max-pool, concat /
// space-to-depth addressing through emit8().  glog = log2(channel groups per pixel) (4 for a 128-channel tile).
// `et` = index of the calling thread among the `nthr` threads that share this tile's store phase.
__device__ __forceinline__ void epilogue_store(const ConvParams &p, const float *stage, int ld, int glog, int b, int y0,
                                               int x0, int cout0, int zsplit, int et, int nthr) {
    const int g = et & ((1 << glog) - 1), ps = et >> glog, slots = nthr >> glog;
    const int rows_valid = min(p.hR, p.H - y0), cols_valid = min(p.hC, p.W - x0);
    if (p.splits != 1) {
        const int chn = cout0 + g * 8;
        if (chn < p.ldp) {
            const long long mtot = (long long)p.B * p.H * p.W;
            for (int r = 0; r < rows_valid; ++r) {
                float *rowp = p.partial + ((long long)zsplit * mtot + ((long long)b * p.H + y0 + r) * p.W + x0) * p.ldp + chn;
                for (int c = ps; c < cols_valid; c += slots) {
                    const float4 *src = reinterpret_cast<const float4 *>(stage + (r * p.hP + c) * ld + g * 8);
                    float4 *dst = reinterpret_cast<float4 *>(rowp + (long long)c * p.ldp);
                    dst[0] = src[0];
                    dst[1] = src[1];
                }
            }
        }
        return;
    }
    if (cout0 + g * 8 >= p.Cout) return;
    // (pixel, group) items are spread over all threads: item -> (row, column) by shift/mask (columns padded to a
    // power of two, the padding items are skipped)
    if (p.out.hi || p.out.f32) {
        const int cs = 32 - __clz(cols_valid - 1), n_it = rows_valid << cs;
        for (int it = ps; it < n_it; it += slots) {
            const int r = it >> cs, c = it & ((1 << cs) - 1);
            if (c >= cols_valid) continue;
            const float4 *src = reinterpret_cast<const float4 *>(stage + (r * p.hP + c) * ld + g * 8);
            const float4 lo4 = src[0], hi4 = src[1];
            const float v8[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
            emit8(p.out, b, y0 + r, x0 + c, cout0 + g * 8, p.Cout, v8);
        }
    }
    if (p.pool) {
        const int pr = rows_valid >> 1, pc = cols_valid >> 1;
        const int cs = 32 - __clz(max(pc, 1) - 1), n_it = pr << cs;
        for (int it = ps; it < n_it; it += slots) {
            const int r2 = it >> cs, c2 = it & ((1 << cs) - 1);
            if (c2 >= pc) continue;
            const float *s0 = stage + (2 * r2 * p.hP + 2 * c2) * ld + g * 8;
            float v8[8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 a0 = reinterpret_cast<const float4 *>(s0)[h];
                const float4 a1 = reinterpret_cast<const float4 *>(s0 + ld)[h];
                const float4 a2 = reinterpret_cast<const float4 *>(s0 + p.hP * ld)[h];
                const float4 a3 = reinterpret_cast<const float4 *>(s0 + (p.hP + 1) * ld)[h];
                v8[4 * h + 0] = fmaxf(fmaxf(a0.x, a1.x), fmaxf(a2.x, a3.x));
                v8[4 * h + 1] = fmaxf(fmaxf(a0.y, a1.y), fmaxf(a2.y, a3.y));
                v8[4 * h + 2] = fmaxf(fmaxf(a0.z, a1.z), fmaxf(a2.z, a3.z));
                v8[4 * h + 3] = fmaxf(fmaxf(a0.w, a1.w), fmaxf(a2.w, a3.w));
            }
            emit8(p.pout, b, ((y0 >> 1) + r2), ((x0 >> 1) + c2), cout0 + g * 8, p.Cout, v8);
        }
    }
}

// One channel `c` of source pixel (b,y,x) -- used by the halo engine, whose epilogue threads own a channel each.
__device__ __forceinline__ void emit1(const Dest &d, int b, int y, int x, int c, int Cout, float v) {
    long long pix;
    int cc = c;
    if (d.mode == DEST_S2D_TF) {
        pix = ((long long)b * (d.H / 2) + (y >> 1)) * (d.W / 2) + (x >> 1);
        cc = ((y & 1) * 2 + (x & 1)) * Cout + c;
    } else {
        pix = ((long long)b * d.H + y) * d.W + x;
    }
    if (d.hi) {
        op_t h, l;
        split_f16(v, h, l);
        op_t *p = d.hi + pix * d.pix_stride_b + d.ch_off_b + cc;
        p[0] = h;
        p[d.plane_stride] = l;
    }
    if (d.f32) {
        if (d.img_stride_f) pix = (long long)b * d.img_stride_f + (long long)y * d.W + x;   // plain mode only
        float *p = d.f32 + pix * d.pix_stride_f + d.ch_off_f + cc;
        *p = d.accumulate_f ? *p + v : v;
    }
}

// Finishes a split-K convolution: fixed-order sum of the raw fp32 partials, scale/bias/leaky, optional 2x2 max-pool,
// hi/lo split, same destinations as the fused epilogue.  One work item = 8 channels of one pixel (or of one 2x2 quad
// when pooling); the caller walks items start, start + step, ...  The partials are read with ld.global.cg: inside
// conv_chain_kernel they were written by other SMs of the SAME grid, and an L1 line of an earlier layer's partials at
// the same address would be stale.
__device__ __forceinline__ void splitk_finish_range(const ConvParams &p, long long start, long long step) {
    const int cgroups = (p.Cout + 7) / 8;
    const int Hq = p.pool ? p.H / 2 : p.H, Wq = p.pool ? p.W / 2 : p.W;
    const long long total = (long long)p.B * Hq * Wq * cgroups;
    const long long mtot = (long long)p.B * p.H * p.W;
    // item -> (image, row, column, channel group): 32-bit arithmetic and exact float divisions (small_div) whenever the
    // counts allow -- 64-bit integer divisions cost more than the reduction itself
    const bool small = total < (1ll << 31) && (long long)p.B * Hq * Wq < (1 << 20);
    for (long long t = start; t < total; t += step) {
        int cg, xq, yq, b;
        if (small) {
            const int qi = (int)t / cgroups;
            cg = (int)t - qi * cgroups;
            const int row = small_div(qi, Wq);
            xq = qi - row * Wq;
            b = small_div(row, Hq);
            yq = row - b * Hq;
        } else {
            cg = int(t % cgroups);
            long long q = t / cgroups;
            xq = int(q % Wq);  q /= Wq;
            yq = int(q % Hq);
            b = int(q / Hq);
        }
        const int c = cg * 8;
        float s8[8], b8[8], mx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool ok = c + i < p.Cout;
            s8[i] = ok ? __ldg(p.scale + c + i) : 0.f;
            b8[i] = ok ? __ldg(p.bias + c + i) : 0.f;
            mx[i] = -INFINITY;
        }
        const int npix = p.pool ? 4 : 1;
        for (int k = 0; k < npix; ++k) {
            const int y = p.pool ? 2 * yq + (k >> 1) : yq, x = p.pool ? 2 * xq + (k & 1) : xq;
            const long long pix = ((long long)b * p.H + y) * p.W + x;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            // fixed order z = 0, 1, 2, ...; eight splits' loads are in flight at a time (ldp is a multiple of 32, so
            // both float4 are in bounds)
            const float *base = p.partial + pix * p.ldp + c;
            const long long zs = mtot * p.ldp;
            int z = 0;
            for (; z + 8 <= p.splits; z += 8) {
                float4 a[8], bq[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 *src = reinterpret_cast<const float4 *>(base + (long long)(z + j) * zs);
                    a[j] = __ldcg(src); bq[j] = __ldcg(src + 1);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v[0] += a[j].x; v[1] += a[j].y; v[2] += a[j].z; v[3] += a[j].w;
                    v[4] += bq[j].x; v[5] += bq[j].y; v[6] += bq[j].z; v[7] += bq[j].w;
                }
            }
            for (; z < p.splits; ++z) {
                const float4 *src = reinterpret_cast<const float4 *>(base + (long long)z * zs);
                const float4 a = __ldcg(src), bq = __ldcg(src + 1);
                v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
                v[4] += bq.x; v[5] += bq.y; v[6] += bq.z; v[7] += bq.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float tt = fmaf(v[i], s8[i], b8[i]);
                v[i] = p.act ? leaky(tt) : tt;
                mx[i] = fmaxf(mx[i], v[i]);
            }
            if (p.out.hi || p.out.f32) emit8(p.out, b, y, x, c, p.Cout, v);
        }
        if (p.pool) emit8(p.pout, b, yq, xq, c, p.Cout, mx);
    }
}

}  // namespace b2t
