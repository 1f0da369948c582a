// DEVELOPER BUILD ONLY (make DEV=1 -> -DB2T_DEV): two cross-check engines that validate the production kernels on the
// device, layer by layer.  They are NOT compiled into the release library -- the product has one conv path
// (conv_halo.cu / conv_pm.cu).
//   * conv_umma_kernel  -- the first-generation tcgen05 tile kernel (one TMA box per tap)
//   * conv_simt_kernel  -- the same contraction with fp32 FMAs on the joined (hi + lo) operands
#ifdef B2T_DEV
// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, accumulator in TMEM), operands fed by TMA.
//
// Replaces, for every conv_k of the YOLOv2 graph except conv_1 (KerasYOLO.py:277-400; darknet
// convolutional_layer.c:445-485 = im2col_cpu + gemm_nn + batchnorm + leaky): Conv2D(k in {1,3}, stride 1, 'same')
// -> BatchNormalization (folded to scale/bias) -> LeakyReLU(0.1) -> optional MaxPooling2D(2,2), and the
// ConvLSTM2D / 1x1-head contractions of MultiObjDetTracker.py:176-183.
//
//   D[pixel, cout] = sum_{tap, c} A[pixel + tap, c] * Wt[cout, tap, c]
//
// * M tile  = 128 pixels = a TW x TH patch of one image (TW*TH = 128); one TMA 4-D box {64 ch, TW, TH, 1} per
//   (tap, 64-channel chunk) with the tap shift in the box coordinates: out-of-image pixels are zero-filled by
//   the TMA unit, which IS the 'same' padding -- no im2col buffer exists anywhere.
// * N tile  = BN output channels (64 or 128), weights K-major [Cout][tap][Cin_pad], one 2-D box {64, BN}.
// * precision: both operands are fp16 hi/lo pairs (x = hi + lo).  Each 64-deep K chunk issues
//   a_hi*w_hi + a_hi*w_lo + a_lo*w_hi into the same fp32 TMEM accumulator (single-pass
//   bf16/tf32 miss the 1e-3 bbox bar; the 3-term form on fp16 pairs carries ~22 bits per operand).
// * warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-5 = epilogue
//   (TMEM -> registers -> scale/bias/leaky -> 2x2 max via shuffles -> hi/lo split -> 16-byte stores).
// * split-K (gridDim.z > 1): raw fp32 accumulators go to a partial buffer, splitk_epilogue_kernel finishes
//   in a fixed order (deterministic).
#include "kernels.cuh"

namespace b2t {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // fp16 elements = one 128-byte swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2; // 16 KB
constexpr int kUmmaThreads = 192;

template <int BN>
struct UmmaCfg {
    static constexpr int kBTileBytes = BN * kBlockK * 2;
    static constexpr int kStageBytes = 2 * kATileBytes + 2 * kBTileBytes;
    static constexpr int kStages = (BN == 128) ? 3 : 4;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * BN * 4;
};

template <int BN>
__global__ void __launch_bounds__(kUmmaThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const ConvParams p) {
    using Cfg = UmmaCfg<BN>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t *tail = smem + kStages * Cfg::kStageBytes;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(tail);
    uint64_t *empty_bar = full_bar + kStages;
    uint64_t *accum_bar = empty_bar + kStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);
    float *s_scale = reinterpret_cast<float *>(tail + 256);
    float *s_bias = s_scale + BN;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- tile coordinates
    int mt = blockIdx.x;
    const int tx = mt % p.tiles_x;  mt /= p.tiles_x;
    const int ty = mt % p.tiles_y;
    const int b = mt / p.tiles_y;
    const int x0 = tx * p.TW, y0 = ty * p.TH;
    const int n0 = blockIdx.y * BN;
    const int per = (p.chunks_total + p.splits - 1) / p.splits;
    const int k_begin = blockIdx.z * per;
    const int k_end = min(p.chunks_total, k_begin + per);
    const int n_iter = k_end - k_begin;              // host guarantees >= 1 for every z
    // TMEM accumulators (BN fp32 columns each): n_main for the hi*hi products, used round-robin over the K
    // chunks, and one for the small hi*lo + lo*hi corrections.  The tensor core truncates every addend to the
    // accumulator's exponent, so short accumulation chains and magnitude-matched accumulators keep the
    // result at fp32 quality; the epilogue adds them up in fp32 round-to-nearest.
    const int n_main = min(p.n_main, n_iter);
    const int pad = p.ksize >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA_hi);
        tma_prefetch_desc(&tmA_lo);
        tma_prefetch_desc(&tmB_hi);
        tma_prefetch_desc(&tmB_lo);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<4 * BN>(tmem_slot);
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < BN; i += 128) {
            const int c = n0 + i;
            s_scale[i] = (c < p.Cout) ? p.scale[c] : 0.f;
            s_bias[i] = (c < p.Cout) ? p.bias[c] : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < n_iter; ++it) {
                const int kc = k_begin + it;
                const int tap = kc / p.cin_chunks, cc = kc - tap * p.cin_chunks;
                const int kh = tap / p.ksize, kw = tap - kh * p.ksize;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t *st = smem + stage * Cfg::kStageBytes;
                mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                tma_load_4d(&tmA_hi, &full_bar[stage], st, cc * kBlockK, x0 + kw - pad, y0 + kh - pad, b, kEvictNormal);
                tma_load_4d(&tmA_lo, &full_bar[stage], st + kATileBytes, cc * kBlockK, x0 + kw - pad, y0 + kh - pad, b,
                            kEvictNormal);
                tma_load_2d(&tmB_hi, &full_bar[stage], st + 2 * kATileBytes, kc * kBlockK, n0, kEvictNormal);
                tma_load_2d(&tmB_lo, &full_bar[stage], st + 2 * kATileBytes + Cfg::kBTileBytes, kc * kBlockK, n0,
                            kEvictNormal);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_f16(kBlockM, BN);
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < n_iter; ++it) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                const uint32_t a_hi = sa, a_lo = sa + kATileBytes;
                const uint32_t b_hi = sa + 2 * kATileBytes, b_lo = b_hi + Cfg::kBTileBytes;
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                    const uint32_t off = k * 32;  // 16 fp16 = 32 bytes inside the 128-byte swizzle row
                    const uint64_t dah = umma_desc_sw128(a_hi + off), dal = umma_desc_sw128(a_lo + off);
                    const uint64_t dbh = umma_desc_sw128(b_hi + off), dbl = umma_desc_sw128(b_lo + off);
                    const uint32_t t_main = tmem_acc + (it % n_main) * BN, t_corr = tmem_acc + n_main * BN;
                    umma_f16(t_corr, dal, dbh, idesc, (it | k) ? 1u : 0u);
                    umma_f16(t_corr, dah, dbl, idesc, 1u);
                    umma_f16(t_main, dah, dbh, idesc, (it >= n_main || k) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);               // frees the smem slot when these MMAs retire
                if (it == n_iter - 1) umma_commit(accum_bar); // accumulator complete
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
    } else {
        // ===================== epilogue =====================
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int r = q * 32 + lane;            // accumulator row = pixel inside the tile
        const int ly = r / p.TW, lx = r - ly * p.TW;
        const int y = y0 + ly, x = x0 + lx;
        const bool valid = (y < p.H) && (x < p.W);
        const long long pix = ((long long)b * p.H + y) * p.W + x;
        const long long mtot = (long long)p.B * p.H * p.W;
#pragma unroll 1
        for (int j = 0; j < BN / 32; ++j) {
            uint32_t acc[32];
            tmem_ld32(tmem_acc + (uint32_t(q * 32) << 16) + j * 32, acc);
            tmem_ld_wait();
            for (int a = 1; a <= n_main; ++a) {          // remaining main accumulators, then the corrections
                uint32_t more[32];
                tmem_ld32(tmem_acc + (uint32_t(q * 32) << 16) + a * BN + j * 32, more);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(more[i]));
            }
            const int c0 = n0 + j * 32;
            if (p.splits > 1) {
                if (valid && c0 < p.ldp) {
                    float4 *dst = reinterpret_cast<float4 *>(p.partial + ((long long)blockIdx.z * mtot + pix) * p.ldp + c0);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        dst[i] = make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]),
                                             __uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3]));
                }
                continue;
            }
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float t = fmaf(__uint_as_float(acc[i]), s_scale[j * 32 + i], s_bias[j * 32 + i]);
                v[i] = p.act ? leaky(t) : t;
            }
            if (valid && (p.out.hi || p.out.f32)) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float(&v8)[8] = *reinterpret_cast<const float(*)[8]>(&v[8 * g]);
                    emit8(p.out, b, y, x, c0 + 8 * g, p.Cout, v8);
                }
            }
            if (p.pool) {  // 2x2/2 max: the four pixels sit in lanes l, l^1, l^TW, l^TW^1 (TW <= 16, tiles even-aligned)
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float m = fmaxf(v[i], __shfl_xor_sync(0xffffffffu, v[i], 1));
                    v[i] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, p.TW));
                }
                if (valid && !(lx & 1) && !(ly & 1)) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float(&v8)[8] = *reinterpret_cast<const float(*)[8]>(&v[8 * g]);
                        emit8(p.pout, b, y >> 1, x >> 1, c0 + 8 * g, p.Cout, v8);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<4 * BN>(tmem_acc);
}

// ------------------------------------------------------------------------------------------------
// SIMT cross-check engine: the same contraction with fp32 FMAs on the joined (hi+lo) operands, raw result to
// p.partial[0]; splitk_epilogue_kernel (splits = 1) finishes it.  Not a product path -- it exists so the
// tcgen05 kernel can be validated layer by layer on the device.

__global__ void __launch_bounds__(256) conv_simt_kernel(const SimtView v, const ConvParams p) {
    __shared__ float sA[64][17];
    __shared__ float sB[64][17];
    const int tid = threadIdx.x;
    const long long mtot = (long long)p.B * p.H * p.W;
    const long long m0 = (long long)blockIdx.x * 64;
    const int n0 = blockIdx.y * 64;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;      // loader: row, first of 4 k
    const int ty = tid >> 4, tx = tid & 15;             // compute: 4 pixels x 4 couts
    float acc[4][4] = {};
    // decode the loader's pixel once
    const long long m = m0 + lrow;
    const bool m_ok = m < mtot;
    int pb = 0, py = 0, px = 0;
    if (m_ok) {
        px = int(m % p.W);
        py = int((m / p.W) % p.H);
        pb = int(m / ((long long)p.W * p.H));
    }
    const int pad = p.ksize >> 1;
    const int cin_pad = p.cin_chunks * (p.kbytes / 2);
    for (int tap = 0; tap < p.ksize * p.ksize; ++tap) {
        const int kh = tap / p.ksize, kw = tap % p.ksize;
        const int yy = py + kh - pad, xx = px + kw - pad;
        const bool in_ok = m_ok && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
        const long long apix = ((long long)pb * p.H + yy) * p.W + xx;
        for (int c0 = 0; c0 < cin_pad; c0 += 16) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float a = 0.f;
                if (in_ok) {
                    const op_t *q = v.a_hi + apix * v.a_pix_stride + c0 + lk + i;
                    a = join_f16(q[0], q[v.a_plane]);
                }
                sA[lrow][lk + i] = a;
                float w = 0.f;
                if (n0 + lrow < p.Cout) {
                    const op_t *q = v.w_hi + (long long)(n0 + lrow) * v.w_ld + tap * cin_pad + c0 + lk + i;
                    w = join_f16(q[0], q[v.w_plane]);
                }
                sB[lrow][lk + i] = w;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                float a[4], w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { a[i] = sA[ty * 4 + i][kk]; w[i] = sB[tx * 4 + i][kk]; }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    for (int i = 0; i < 4; ++i) {
        const long long mm = m0 + ty * 4 + i;
        if (mm >= mtot) continue;
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tx * 4 + j;
            if (c < p.ldp) p.partial[mm * p.ldp + c] = acc[i][j];
        }
    }
}

// ---------------------------------------------------------------- host-side launchers (used by api.cu)
int launch_conv_umma(int BN, const CUtensorMap &a_hi, const CUtensorMap &a_lo, const CUtensorMap &b_hi,
                     const CUtensorMap &b_lo, const ConvParams &p, cudaStream_t st) {
    dim3 grid(p.B * p.tiles_x * p.tiles_y, (p.Cout + BN - 1) / BN, p.splits);
    if (BN == 128) {
        conv_umma_kernel<128><<<grid, kUmmaThreads, UmmaCfg<128>::kSmemBytes, st>>>(a_hi, a_lo, b_hi, b_lo, p);
    } else {
        conv_umma_kernel<64><<<grid, kUmmaThreads, UmmaCfg<64>::kSmemBytes, st>>>(a_hi, a_lo, b_hi, b_lo, p);
    }
    return (int)cudaGetLastError();
}
int conv_umma_init() {
    cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         UmmaCfg<128>::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(conv_umma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             UmmaCfg<64>::kSmemBytes);
    return (int)e;
}
int launch_conv_simt(const SimtView &v, const ConvParams &p, cudaStream_t st) {
    const long long mtot = (long long)p.B * p.H * p.W;
    dim3 grid((unsigned)((mtot + 63) / 64), (p.ldp + 63) / 64);
    conv_simt_kernel<<<grid, 256, 0, st>>>(v, p);
    return (int)cudaGetLastError();
}

}  // namespace b2t
#endif  // B2T_DEV
