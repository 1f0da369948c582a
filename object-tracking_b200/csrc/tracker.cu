// Recurrent tracker steps.
//
// lstm_gates_kernel / dense_sigmoid_kernel  = TinyTracker.py:36-37 and TinyHeatmapTracker.py:43-44:
//     LSTM(units, implementation=2) -> Dense(n_out, sigmoid), Keras-2 gate order i,f,c,o,
//     z = [x,h]·[W;U] + b, i,f,o = hard_sigmoid(z) (or sigmoid), c' = f*c + i*tanh(z_c), h' = o*tanh(c').
// pool_features_kernel                      = TinyTracker.py:29-33 GlobalMaxPooling2D | MaxPooling2D(4,4)+Flatten
//                                             (optionally through the CHW-viewed-as-HWC reinterpretation of
//                                             preprocessing.py:419).
// heatmap kernels                           = utils.py:53-58 / :61-79.
// convlstm_gates_kernel                     = gate maths of ConvLSTM2D (MultiObjDetTracker.py:176).
//
// The LSTM is a latency-bound GEMV (12.6 MB of fp32 weights, a few streams): the weights are packed at load
// time so that the [4 gates x 4 units] columns one CTA owns are contiguous per input row (64-byte rows, fully
// coalesced), 128 CTAs stream disjoint slabs once per step for all S streams, and the reduction order is fixed.
#include "kernels.cuh"

namespace b2t {

constexpr int kUnitsPerBlock = 4;
constexpr int kLstmThreads = 256;
constexpr int kMaxStreams = 8;      // streams per launch group


__device__ __forceinline__ float hard_sigmoid_f(float x) { return fminf(fmaxf(fmaf(0.2f, x, 0.5f), 0.f), 1.f); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// Thread layout: 4 lanes cover the 16 columns of one packed weight row (one float4 each), a warp covers 8
// consecutive rows (512 contiguous bytes per load instruction), the CTA's 8 warps cover 64 rows per iteration and
// the row loop is unrolled x4 so that every thread keeps 4 independent 16-byte loads in flight (the kernel is a
// latency-bound GEMV: memory-level parallelism is what buys bandwidth).
__global__ void __launch_bounds__(kLstmThreads) lstm_gates_kernel(const LstmParams p) {
    __shared__ float red[8][16][kMaxStreams];
    const int ub = blockIdx.x;
    const int s0 = blockIdx.y * kMaxStreams;                   // first row (stream) of this CTA's group
    const int nS = min(kMaxStreams, p.S - s0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c4 = lane & 3, rq = threadIdx.x >> 2;            // column quad, row slot (0..63)
    const int n_x = p.n_feat + p.n_det, n_rows = n_x + p.units;
    const float4 *w = reinterpret_cast<const float4 *>(p.wp + (long long)ub * n_rows * 16) + c4;
    const int k_lo = p.mode == 2 ? n_x : 0, k_hi = p.mode == 1 ? n_x : n_rows;
    float acc[kMaxStreams][4];
#pragma unroll
    for (int s = 0; s < kMaxStreams; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[s][j] = 0.f;
    auto xval = [&](int s, int k) -> float {
        if (k < p.n_feat) return __ldg(p.fv + (long long)(s0 + s) * p.fv_stride + k);
        if (k < n_x) return __ldg(p.det + (long long)(s0 + s) * p.det_stride + (k - p.n_feat));
        return __ldg(p.h_in + (long long)(s0 + s) * p.units + (k - n_x));
    };
    for (int k0 = k_lo + rq; k0 < k_hi; k0 += 4 * 64) {
        float4 wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k0 + u * 64;
            wv[u] = k < k_hi ? __ldg(w + (long long)k * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k0 + u * 64;
            if (k < k_hi) {
#pragma unroll
                for (int s = 0; s < kMaxStreams; ++s) {
                    if (s < nS) {
                        const float xv = xval(s, k);
                        acc[s][0] = fmaf(xv, wv[u].x, acc[s][0]);
                        acc[s][1] = fmaf(xv, wv[u].y, acc[s][1]);
                        acc[s][2] = fmaf(xv, wv[u].z, acc[s][2]);
                        acc[s][3] = fmaf(xv, wv[u].w, acc[s][3]);
                    }
                }
            }
        }
    }
    // reduce the 8 row slots of a warp (lane bits 2..4), fixed order, then the 8 warps through shared memory
#pragma unroll
    for (int s = 0; s < kMaxStreams; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = acc[s][j];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (lane < 4) red[warp][c4 * 4 + j][s] = v;
        }
    __syncthreads();
    // 4 units x nS streams finish
    if (threadIdx.x < kUnitsPerBlock * nS) {
        const int uu = threadIdx.x % kUnitsPerBlock, s = s0 + threadIdx.x / kUnitsPerBlock;
        const int unit = ub * kUnitsPerBlock + uu;
        float z[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            float t = 0.f;
            for (int r = 0; r < 8; ++r) t += red[r][g * 4 + uu][s - s0];
            z[g] = t + (p.mode == 2 ? p.zx[(long long)s * p.zx_stride + g * p.units + unit] : p.bias[g * p.units + unit]);
        }
        if (p.mode == 1) {
#pragma unroll
            for (int g = 0; g < 4; ++g) p.zx[(long long)s * p.zx_stride + g * p.units + unit] = z[g];
            return;
        }
        const float i = p.hard_sigmoid ? hard_sigmoid_f(z[0]) : sigmoid_f(z[0]);
        const float f = p.hard_sigmoid ? hard_sigmoid_f(z[1]) : sigmoid_f(z[1]);
        const float o = p.hard_sigmoid ? hard_sigmoid_f(z[3]) : sigmoid_f(z[3]);
        const float cn = fmaf(f, p.c[(long long)s * p.units + unit], i * tanhf(z[2]));
        const float hn = o * tanhf(cn);
        p.c[(long long)s * p.units + unit] = cn;
        p.h_out[(long long)s * p.units + unit] = hn;
        if (p.h_seq) p.h_seq[(long long)s * p.h_seq_stride + unit] = hn;
    }
}

// ---------------------------------------------------------------- batched input projection
// zx[r] = x[r] * W + b for all R rows (frames) in ONE pass over the weights: a thread keeps its share of the CTA's
// weight slab (rows rq, rq+64, ... x one float4 of the 16 columns) in registers -- all loads in flight at once -- and
// walks the rows in groups of 8.  Same thread layout and summation order as lstm_gates_kernel mode 1 (which it
// replaces when n_feat + n_det <= 64 * kProjMaxK).
constexpr int kProjMaxK = 17;
__global__ void __launch_bounds__(kLstmThreads) lstm_proj_kernel(const LstmParams p) {
    __shared__ float red[8][16][kMaxStreams];
    __shared__ __align__(16) float xs[kMaxStreams][64 * kProjMaxK];      // the group's input rows (coalesced staging)
    const int ub = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c4 = lane & 3, rq = threadIdx.x >> 2;
    const int n_x = p.n_feat + p.n_det, n_rows = n_x + p.units;
    const float4 *w = reinterpret_cast<const float4 *>(p.wp + (long long)ub * n_rows * 16) + c4;
    float4 wv[kProjMaxK];
#pragma unroll
    for (int j = 0; j < kProjMaxK; ++j) {
        const int k = rq + 64 * j;
        wv[j] = k < n_x ? __ldg(w + (long long)k * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int s0 = 0; s0 < p.S; s0 += kMaxStreams) {
        const int nS = min(kMaxStreams, p.S - s0);
        if (((p.n_feat | p.fv_stride) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.fv) & 15) == 0) {
            // one float4 per thread and row, the group's loads all in flight before the first store
            for (int k = threadIdx.x * 4; k < p.n_feat; k += kLstmThreads * 4) {
                float4 v[kMaxStreams];
#pragma unroll
                for (int s = 0; s < kMaxStreams; ++s)
                    if (s < nS) v[s] = __ldg(reinterpret_cast<const float4 *>(p.fv + (long long)(s0 + s) * p.fv_stride + k));
#pragma unroll
                for (int s = 0; s < kMaxStreams; ++s)
                    if (s < nS) *reinterpret_cast<float4 *>(&xs[s][k]) = v[s];
            }
        } else {
            for (int i = threadIdx.x; i < nS * p.n_feat; i += kLstmThreads) {
                const int s = i / p.n_feat, k = i - s * p.n_feat;
                xs[s][k] = __ldg(p.fv + (long long)(s0 + s) * p.fv_stride + k);
            }
        }
        for (int i = threadIdx.x; i < nS * p.n_det; i += kLstmThreads) {
            const int s = i / p.n_det, k = i - s * p.n_det;
            xs[s][p.n_feat + k] = __ldg(p.det + (long long)(s0 + s) * p.det_stride + k);
        }
        __syncthreads();
        float acc[kMaxStreams][4];
#pragma unroll
        for (int s = 0; s < kMaxStreams; ++s)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[s][j] = 0.f;
#pragma unroll
        for (int j = 0; j < kProjMaxK; ++j) {
            const int k = rq + 64 * j;
            if (k < n_x) {
#pragma unroll
                for (int s = 0; s < kMaxStreams; ++s)
                    if (s < nS) {
                        const float xv = xs[s][k];
                        acc[s][0] = fmaf(xv, wv[j].x, acc[s][0]);
                        acc[s][1] = fmaf(xv, wv[j].y, acc[s][1]);
                        acc[s][2] = fmaf(xv, wv[j].z, acc[s][2]);
                        acc[s][3] = fmaf(xv, wv[j].w, acc[s][3]);
                    }
            }
        }
#pragma unroll
        for (int s = 0; s < kMaxStreams; ++s)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = acc[s][j];
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (lane < 4) red[warp][c4 * 4 + j][s] = v;
            }
        __syncthreads();
        if (threadIdx.x < kUnitsPerBlock * nS) {
            const int uu = threadIdx.x % kUnitsPerBlock, s = s0 + threadIdx.x / kUnitsPerBlock;
            const int unit = ub * kUnitsPerBlock + uu;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float t = 0.f;
                for (int r = 0; r < 8; ++r) t += red[r][g * 4 + uu][s - s0];
                p.zx[(long long)s * p.zx_stride + g * p.units + unit] = t + p.bias[g * p.units + unit];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- fused recurrent steps
// T sequential steps h_t = LSTM(zx_t + h_{t-1} U) for S streams in ONE launch: units/4 CTAs (all co-resident: one per
// SM), each keeps its U slab in registers for the whole sequence; between steps the CTAs meet at a counter barrier in
// global memory (h_{t-1} is read with ld.cg after it).  Same summation order as lstm_gates_kernel mode 2.
constexpr int kSeqMaxStreams = 16;
constexpr int kSeqMaxK = 8;              // units <= 512
__global__ void __launch_bounds__(kLstmThreads) lstm_seq_kernel(const LstmParams p, int T, float *h_a, float *h_b,
                                                                unsigned int *counter) {
    __shared__ float red[8][16][kSeqMaxStreams];
    __shared__ __align__(16) float hs[kSeqMaxStreams][64 * kSeqMaxK];     // h_{t-1} of every stream
    const int ub = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c4 = lane & 3, rq = threadIdx.x >> 2;
    const int n_x = p.n_feat + p.n_det, n_rows = n_x + p.units;
    const float4 *w = reinterpret_cast<const float4 *>(p.wp + ((long long)ub * n_rows + n_x) * 16) + c4;
    float4 wv[kSeqMaxK];
#pragma unroll
    for (int j = 0; j < kSeqMaxK; ++j) {
        const int k = rq + 64 * j;
        wv[j] = k < p.units ? __ldg(w + (long long)k * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int t = 0; t < T; ++t) {
        const float *h_in = (t & 1) ? h_b : h_a;
        float *h_out = (t & 1) ? h_a : h_b;
        {   // coalesced float4 loads from L2 (other CTAs wrote h), all in flight before the first store
            const int n4 = p.S * p.units / 4;                              // units is a multiple of 4
            float4 v[kSeqMaxStreams * 64 * kSeqMaxK / 4 / kLstmThreads];
#pragma unroll
            for (int r = 0; r < kSeqMaxStreams * 64 * kSeqMaxK / 4 / kLstmThreads; ++r) {
                const int i4 = threadIdx.x + r * kLstmThreads;
                if (i4 < n4) v[r] = __ldcg(reinterpret_cast<const float4 *>(h_in) + i4);
            }
#pragma unroll
            for (int r = 0; r < kSeqMaxStreams * 64 * kSeqMaxK / 4 / kLstmThreads; ++r) {
                const int i4 = threadIdx.x + r * kLstmThreads;
                if (i4 < n4) {
                    const int i = i4 * 4, sidx = i / p.units, k = i - sidx * p.units;
                    *reinterpret_cast<float4 *>(&hs[sidx][k]) = v[r];
                }
            }
        }
        __syncthreads();
        float acc[kSeqMaxStreams][4];
#pragma unroll
        for (int s = 0; s < kSeqMaxStreams; ++s)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[s][j] = 0.f;
#pragma unroll
        for (int j = 0; j < kSeqMaxK; ++j) {
            const int k = rq + 64 * j;
            if (k < p.units) {
#pragma unroll
                for (int s = 0; s < kSeqMaxStreams; ++s)
                    if (s < p.S) {
                        const float hv = hs[s][k];
                        acc[s][0] = fmaf(hv, wv[j].x, acc[s][0]);
                        acc[s][1] = fmaf(hv, wv[j].y, acc[s][1]);
                        acc[s][2] = fmaf(hv, wv[j].z, acc[s][2]);
                        acc[s][3] = fmaf(hv, wv[j].w, acc[s][3]);
                    }
            }
        }
#pragma unroll
        for (int s = 0; s < kSeqMaxStreams; ++s)
            if (s < p.S) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v = acc[s][j];
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (lane < 4) red[warp][c4 * 4 + j][s] = v;
                }
            }
        __syncthreads();
        if (threadIdx.x < kUnitsPerBlock * p.S) {
            const int uu = threadIdx.x % kUnitsPerBlock, s = threadIdx.x / kUnitsPerBlock;
            const int unit = ub * kUnitsPerBlock + uu;
            float z[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float tt = 0.f;
                for (int r = 0; r < 8; ++r) tt += red[r][g * 4 + uu][s];
                z[g] = tt + p.zx[((long long)s * T + t) * 4 * p.units + g * p.units + unit];
            }
            const float i = p.hard_sigmoid ? hard_sigmoid_f(z[0]) : sigmoid_f(z[0]);
            const float f = p.hard_sigmoid ? hard_sigmoid_f(z[1]) : sigmoid_f(z[1]);
            const float o = p.hard_sigmoid ? hard_sigmoid_f(z[3]) : sigmoid_f(z[3]);
            const float cn = fmaf(f, p.c[(long long)s * p.units + unit], i * tanhf(z[2]));
            const float hn = o * tanhf(cn);
            p.c[(long long)s * p.units + unit] = cn;
            h_out[(long long)s * p.units + unit] = hn;
            if (p.h_seq) p.h_seq[((long long)s * T + t) * p.units + unit] = hn;
        }
        if (t + 1 < T) {                     // every CTA has published h_t before anyone reads it
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                atomicAdd(counter, 1u);
                const unsigned int target = gridDim.x * (unsigned)(t + 1);
                unsigned int seen;
                const long long t_start = clock64();
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
                    // all CTAs are co-resident: launch_lstm_seq launches cooperatively (a runtime guarantee); should
                    // the barrier still not complete, fail loudly instead of hanging the device
                    if (seen < target && clock64() - t_start > (1ll << 31)) __trap();
                } while (seen < target);
            }
            __syncthreads();
        }
    }
}

// y[s][j] = sigmoid(sum_k h[s][k] * Wd[k][j] + bd[j]); one warp per output, lanes over k
__global__ void dense_sigmoid_kernel(const float *h, const float *wd, const float *bd, int units, int n_out, int S,
                                     float *y, int y_stride) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= S * n_out) return;
    const int s = gw / n_out, j = gw - s * n_out;
    float acc = 0.f;
    for (int k = lane; k < units; k += 32) acc = fmaf(h[(long long)s * units + k], __ldg(wd + (long long)k * n_out + j), acc);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[(long long)s * y_stride + j] = sigmoid_f(acc + bd[j]);
}

// ---------------------------------------------------------------- feature pooling

__device__ __forceinline__ float pool_read(const PoolParams &p, int b, int h, int w, int c) {
    int yy = h, xx = w, cc = c;
    if (p.chw_view) {   // viewed[h][w][c] = flat_chw[(h*W + w)*C + c]
        const int f = (h * p.W + w) * p.C + c;
        cc = f / (p.H * p.W);
        yy = (f / p.W) % p.H;
        xx = f % p.W;
    }
    const op_t *q = p.hi + (((long long)b * p.H + yy) * p.W + xx) * p.pix_stride + p.ch_off + cc;
    return join_f16(q[0], q[p.plane]);
}

__global__ void pool_features_kernel(const PoolParams p) {
    const int PH = p.H / 4, PW = p.W / 4;
    const int per = p.mode == 0 ? p.C : PH * PW * p.C;
    const long long total = (long long)p.B * per;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int b = int(t / per), r = int(t - (long long)b * per);
        float m = -INFINITY;
        if (p.mode == 0) {
            for (int h = 0; h < p.H; ++h)
                for (int w = 0; w < p.W; ++w) m = fmaxf(m, pool_read(p, b, h, w, r));
        } else {
            const int c = r % p.C, pw = (r / p.C) % PW, ph = r / (p.C * PW);
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) m = fmaxf(m, pool_read(p, b, ph * 4 + i, pw * 4 + j, c));
        }
        p.out[t] = m;
    }
}

// Global max, natural layout: one CTA per (frame, 64 channels).  A thread owns 8 consecutive channels (one 16-byte load
// per plane and pixel) and one of 32 pixel lanes; a warp reads 4 pixels x 128 contiguous bytes per load instruction and
// every thread has all its loads in flight at once (13x13 grid: 6 pixels x 2 planes).  max() is exact in any order.
__global__ void __launch_bounds__(256) pool_global_kernel(const PoolParams p) {
    __shared__ float red[32][64 + 1];
    const int b = blockIdx.x, cg = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const int c = blockIdx.y * 64 + cg * 8;
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
    if (c < p.C) {
        const op_t *q = p.hi + (long long)b * p.H * p.W * p.pix_stride + p.ch_off + c;
#pragma unroll 4
        for (int px = pl; px < p.H * p.W; px += 32) {
            const op_t *e = q + (long long)px * p.pix_stride;
            const uint4 h4 = *reinterpret_cast<const uint4 *>(e), l4 = *reinterpret_cast<const uint4 *>(e + p.plane);
            const op_t *h = reinterpret_cast<const op_t *>(&h4), *l = reinterpret_cast<const op_t *>(&l4);
#pragma unroll
            for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], join_f16(h[i], l[i]));
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[pl][cg * 8 + i] = m[i];
    __syncthreads();
    if (threadIdx.x < 64) {
        const int ch = blockIdx.y * 64 + threadIdx.x;
        float v = red[0][threadIdx.x];
#pragma unroll
        for (int r = 1; r < 32; ++r) v = fmaxf(v, red[r][threadIdx.x]);
        if (ch < p.C) p.out[(long long)b * p.C + ch] = v;
    }
}

// ---------------------------------------------------------------- heat maps (utils.py:53-79)
__device__ __forceinline__ void heatmap_fill(double x, double y, double w, double h, int size, float *heat) {
    // python int() truncates toward zero; numpy slices wrap negative bounds once and clamp to the array
    const int sx = (int)(x * size), sy = (int)(y * size), sh = (int)(h * size), sw = (int)(w * size);
    auto norm = [size](int v) { int t = v < 0 ? v + size : v; return t < 0 ? 0 : (t > size ? size : t); };
    const int y0 = norm(sy), y1 = norm(sy + sh + 1), x0 = norm(sx), x1 = norm(sx + sw + 1);
    for (int i = threadIdx.x; i < size * size; i += blockDim.x) {
        const int r = i / size, c = i % size;
        heat[i] = (r >= y0 && r < y1 && c >= x0 && c < x1) ? 1.f : 0.f;
    }
}

// utils.py:53-58 generate_heatmap_feat on (S,4) rows [x, y, w, h] = top-left corner and size, image-relative
__global__ void heatmap_from_box_kernel(const float *xywh, int n, int size, float *heat) {
    const int s = blockIdx.x;
    if (s >= n) return;
    heatmap_fill(xywh[4 * s], xywh[4 * s + 1], xywh[4 * s + 2], xywh[4 * s + 3], size,
                 heat + (long long)s * size * size);
}

// preprocessing.py:434-456 + YOLO.py:177-180: take the highest-probability detection whose class is allowed
// (rows are already sorted by -prob), normalise by the frame size in double like the python code, emit the
// LSTM's bbox input [cx/w, cy/h, bw/w, bh/h] (zeros if none) and, optionally, its heat-map.
__global__ void select_detection_kernel(const SelectParams p) {
    const int b = blockIdx.x;
    __shared__ int pick;
    if (threadIdx.x == 0) {
        pick = -1;
        const int n = max(0, min(p.counts[b], p.max_dets));
        for (int i = 0; i < n; ++i) {
            const int cls = (int)p.dets[((long long)b * p.max_dets + i) * 8 + 6];
            if (!p.class_mask || p.class_mask[cls]) { pick = i; break; }
        }
        if (p.chosen) p.chosen[b] = pick;
    }
    __syncthreads();
    double x = 0, y = 0, w = 0, h = 0;
    if (pick >= 0) {
        const float *d = p.dets + ((long long)b * p.max_dets + pick) * 8;
        x = (double)d[0] / p.frame_w;  y = (double)d[1] / p.frame_h;
        w = (double)d[2] / p.frame_w;  h = (double)d[3] / p.frame_h;
    }
    if (threadIdx.x == 0 && p.det_in) {
        float *o = p.det_in + 4 * b;
        o[0] = (float)x; o[1] = (float)y; o[2] = (float)w; o[3] = (float)h;
    }
    if (p.heat) heatmap_fill(x - w / 2.0, y - h / 2.0, w, h, p.heat_size, p.heat + (long long)b * p.heat_size * p.heat_size);
}

__global__ void box_from_heatmap_kernel(const float *heat, int n, int size, float thresh, int *rect) {
    const int s = blockIdx.x;
    if (s >= n) return;
    __shared__ int sm[4];
    if (threadIdx.x == 0) { sm[0] = size; sm[1] = size; sm[2] = -1; sm[3] = -1; }
    __syncthreads();
    int x1 = size, y1 = size, x2 = -1, y2 = -1;
    for (int i = threadIdx.x; i < size * size; i += blockDim.x) {
        if (heat[(long long)s * size * size + i] >= thresh) {
            const int r = i / size, c = i % size;
            x1 = min(x1, c); y1 = min(y1, r); x2 = max(x2, c); y2 = max(y2, r);
        }
    }
    for (int o = 16; o; o >>= 1) {
        x1 = min(x1, __shfl_xor_sync(0xffffffffu, x1, o));
        y1 = min(y1, __shfl_xor_sync(0xffffffffu, y1, o));
        x2 = max(x2, __shfl_xor_sync(0xffffffffu, x2, o));
        y2 = max(y2, __shfl_xor_sync(0xffffffffu, y2, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&sm[0], x1); atomicMin(&sm[1], y1); atomicMax(&sm[2], x2); atomicMax(&sm[3], y2);
    }
    __syncthreads();
    if (threadIdx.x < 4) rect[4 * s + threadIdx.x] = sm[threadIdx.x];
}

// ---------------------------------------------------------------- ConvLSTM2D gates

__global__ void convlstm_gates_kernel(const ConvLstmGateParams p) {
    // one thread = 8 units of one pixel of one stream; stream s of this call lives in state slot slot0 + s and its
    // time step t is frame s*T + t of the gate / h_seq buffers
    const int groups = p.units / 8;
    const long long total = (long long)p.S * p.M * groups;
    for (long long tt = blockIdx.x * (long long)blockDim.x + threadIdx.x; tt < total;
         tt += (long long)gridDim.x * blockDim.x) {
        const int u0 = int(tt % groups) * 8;
        const int sm = int(tt / groups), s = sm / p.M, m = sm - s * p.M;
        const float *g = p.g + ((long long)(s * p.T + p.t) * p.M + m) * 4 * p.units;
        float *c = p.c + ((long long)(p.slot0 + s) * p.M + m) * p.units;
        float h8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int u = u0 + i;
            const float zi = g[u], zf = g[p.units + u], zc = g[2 * p.units + u], zo = g[3 * p.units + u];
            const float ig = p.hard_sigmoid ? hard_sigmoid_f(zi) : sigmoid_f(zi);
            const float fg = p.hard_sigmoid ? hard_sigmoid_f(zf) : sigmoid_f(zf);
            const float og = p.hard_sigmoid ? hard_sigmoid_f(zo) : sigmoid_f(zo);
            const float cn = fmaf(fg, c[u], ig * tanhf(zc));
            c[u] = cn;
            h8[i] = og * tanhf(cn);
        }
        const int y = m / p.G, x = m - y * p.G;
        emit8(p.h_rec, p.slot0 + s, y, x, u0, p.units, h8);
        emit8(p.h_seq, s * p.T + p.t, y, x, u0, p.units, h8);
    }
}

// ---------------------------------------------------------------- launchers
int launch_lstm_gates(const LstmParams &p, cudaStream_t st) {
    dim3 grid(p.units / kUnitsPerBlock, (p.S + kMaxStreams - 1) / kMaxStreams);
    lstm_gates_kernel<<<grid, kLstmThreads, 0, st>>>(p);
    return (int)cudaGetLastError();
}
int launch_lstm_proj(const LstmParams &p, cudaStream_t st) {
    if (p.n_feat + p.n_det > 64 * kProjMaxK) return -1;          // caller falls back to lstm_gates_kernel mode 1
    lstm_proj_kernel<<<p.units / kUnitsPerBlock, kLstmThreads, 0, st>>>(p);
    return (int)cudaGetLastError();
}
int launch_lstm_seq(const LstmParams &p, int T, float *h_a, float *h_b, unsigned int *counter, int n_sm, cudaStream_t st) {
    if (p.S > kSeqMaxStreams || p.units > 64 * kSeqMaxK || p.units / kUnitsPerBlock > n_sm) return -1;   // fall back
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return (int)e;
    // The kernel synchronises its CTAs between time steps with a counter barrier, so every CTA must be resident at
    // once.  A cooperative launch makes that a guarantee of the runtime (it is refused when the grid cannot be
    // co-resident, and scheduled as a whole when other streams -- e.g. the next step's persistent conv kernels in
    // BaseTracker's pipelined mode -- occupy SMs), instead of an assumption about what else is running.
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_seq_kernel, kLstmThreads, 0);
    if (e != cudaSuccess) return (int)e;
    if (per_sm * n_sm < p.units / kUnitsPerBlock) return -1;     // fall back to one launch per step
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.units / kUnitsPerBlock); cfg.blockDim = dim3(kLstmThreads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, lstm_seq_kernel, p, T, h_a, h_b, counter);
}
int launch_dense_sigmoid(const float *h, const float *wd, const float *bd, int units, int n_out, int S, float *y,
                         int y_stride, cudaStream_t st) {
    const long long warps = (long long)S * n_out;
    dense_sigmoid_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(h, wd, bd, units, n_out, S, y, y_stride);
    return (int)cudaGetLastError();
}
int launch_pool_features(const PoolParams &p, cudaStream_t st) {
    // 16-byte loads: 8-channel groups aligned in both planes (every activation buffer of the engine is; anything else
    // takes the element-wise kernel below)
    const bool vec = (p.C & 7) == 0 && (p.pix_stride & 7) == 0 && (p.ch_off & 7) == 0 && (p.plane & 7) == 0 &&
                     (reinterpret_cast<uintptr_t>(p.hi) & 15) == 0;
    if (p.mode == 0 && !p.chw_view && vec) {
        pool_global_kernel<<<dim3(p.B, (p.C + 63) / 64), 256, 0, st>>>(p);
        return (int)cudaGetLastError();
    }
    const int per = p.mode == 0 ? p.C : (p.H / 4) * (p.W / 4) * p.C;
    const long long total = (long long)p.B * per;
    pool_features_kernel<<<(unsigned)min((long long)148 * 8, (total + 127) / 128), 128, 0, st>>>(p);
    return (int)cudaGetLastError();
}
int launch_heatmap_from_box(const float *xywh, int n, int size, float *heat, cudaStream_t st) {
    heatmap_from_box_kernel<<<n, 256, 0, st>>>(xywh, n, size, heat);
    return (int)cudaGetLastError();
}
int launch_select_detection(const SelectParams &p, cudaStream_t st) {
    select_detection_kernel<<<p.B, 256, 0, st>>>(p);
    return (int)cudaGetLastError();
}
int launch_box_from_heatmap(const float *heat, int n, int size, float thresh, int *rect, cudaStream_t st) {
    box_from_heatmap_kernel<<<n, 256, 0, st>>>(heat, n, size, thresh, rect);
    return (int)cudaGetLastError();
}
int launch_convlstm_gates(const ConvLstmGateParams &p, cudaStream_t st) {
    const long long total = (long long)p.S * p.M * (p.units / 8);
    convlstm_gates_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace b2t
