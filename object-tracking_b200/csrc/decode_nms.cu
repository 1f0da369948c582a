// Anchor decode + objectness threshold + NMS, one CTA per frame, entirely on the device; only the kept boxes
// ever cross PCIe.
//
// keras flavour  = utility/utils.py:208-257 decode_netout (+ sigmoid :259-260, softmax :262-270 with its GLOBAL
//                  max subtraction and global-min < -100 temperature rescale, bbox_iou :155-173,
//                  interval_overlap :175-188): per-class greedy NMS, IoU >= thr suppresses, output in row-major
//                  anchor order.
// darknet flavour = region_layer.c:158-185 (logistic / per-anchor softmax), :76-84 get_region_box,
//                  :364-437 get_region_detections, :336-362 correct_region_boxes (letterbox un-mapping,
//                  relative = 0), box.c:21-55 do_nms_obj (class-agnostic, sorted by objectness, IoU > thr),
//                  YOLO.py:152-159 (one row per (detection, class) with prob > 0, sorted by -prob).
//
// Arithmetic: every step the reference performs in float32 is performed in float32 here, in the same order
// (numpy's pairwise summation included); exp() is evaluated in double and rounded once, which is within
// 2 ulp of numpy's SIMD float32 exp (the only source of non-bit-exactness).
// Parallel structure: warp-shuffle reductions for the global max/min, ballot compaction of the anchors whose
// objectness can pass the threshold (p = conf*softmax <= conf), one warp per surviving anchor for the
// softmax, rank-by-counting sort, one warp per class segment for the greedy suppression.
#include "kernels.cuh"

namespace b2t {

constexpr int kDecThreads = 512;
constexpr int kMaxCand = 2048;     // anchors that pass the objectness gate (19*19*5 = 1805 fits)
constexpr int kMaxEntry = 4096;    // (anchor, class) pairs above threshold


struct DecodeSmem {
    float red_max[16], red_min[16];
    float gmax, gmin_z;
    int n_cand, n_entry, n_seg, n_out, overflow;
    int warp_counts[16];
    short cand_anchor[kMaxCand];
    float cand_conf[kMaxCand];
    float4 cand_box[kMaxCand];                 // x, y, w, h
    unsigned long long cand_best[kMaxCand];    // (score bits << 32) | (0xFFFFFFFF - class)
    float ent_p[kMaxEntry];
    short ent_cand[kMaxEntry], ent_cls[kMaxEntry];
    float srt_p[kMaxEntry];
    short srt_cand[kMaxEntry], srt_cls[kMaxEntry];
    unsigned char alive[kMaxEntry];
    short seg_start[kMaxEntry + 1];
    float scratch[16][161];                    // per-warp class-vector staging (C <= 160 fast path) + 1 stash
};

__device__ __forceinline__ float exp_rn(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float sigmoid_np(float x) {  // 1. / (1. + np.exp(-x)) in float32
    return __fdiv_rn(1.f, __fadd_rn(1.f, exp_rn(-x)));
}

// numpy's float32 add.reduce over a contiguous row: pairwise_sum (8 accumulators below 128 elements,
// recursive halving above), added to a zero initial value.
__device__ float np_pairwise_sum(const float *a, int n) {
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
        return r;
    }
    if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_pairwise_sum(a, n2), np_pairwise_sum(a + n2, n - n2));
}

// utils.py:175-188 interval_overlap / :155-173 bbox_iou, float32 step by step
__device__ __forceinline__ float overlap_np(float a_lo, float a_hi, float b_lo, float b_hi) {
    if (b_lo < a_lo) return (b_hi < a_lo) ? 0.f : __fsub_rn(fminf(a_hi, b_hi), a_lo);
    return (a_hi < b_lo) ? 0.f : __fsub_rn(fminf(a_hi, b_hi), b_lo);
}
__device__ __forceinline__ float iou_keras(const float4 p, const float4 q) {
    const float pw2 = p.z * 0.5f, ph2 = p.w * 0.5f, qw2 = q.z * 0.5f, qh2 = q.w * 0.5f;
    const float iw = overlap_np(__fsub_rn(p.x, pw2), __fadd_rn(p.x, pw2), __fsub_rn(q.x, qw2), __fadd_rn(q.x, qw2));
    const float ih = overlap_np(__fsub_rn(p.y, ph2), __fadd_rn(p.y, ph2), __fsub_rn(q.y, qh2), __fadd_rn(q.y, qh2));
    const float inter = __fmul_rn(iw, ih);
    const float uni = __fsub_rn(__fadd_rn(__fmul_rn(p.z, p.w), __fmul_rn(q.z, q.w)), inter);
    return __fdiv_rn(inter, uni);
}
// box.c:152-182
__device__ __forceinline__ float overlap_dk(float x1, float w1, float x2, float w2) {
    const float l1 = __fsub_rn(x1, w1 * 0.5f), l2 = __fsub_rn(x2, w2 * 0.5f);
    const float r1 = __fadd_rn(x1, w1 * 0.5f), r2 = __fadd_rn(x2, w2 * 0.5f);
    return __fsub_rn(r1 < r2 ? r1 : r2, l1 > l2 ? l1 : l2);
}
__device__ __forceinline__ float iou_darknet(const float4 a, const float4 b) {
    const float w = overlap_dk(a.x, a.z, b.x, b.z), h = overlap_dk(a.y, a.w, b.y, b.w);
    const float inter = (w < 0.f || h < 0.f) ? 0.f : __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(__fmul_rn(a.z, a.w), __fmul_rn(b.z, b.w)), inter);
    return __fdiv_rn(inter, uni);
}

__device__ __forceinline__ float warp_max(float v) {
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Ordered append of flagged items across the CTA (ballot + warp prefix + smem scan): returns the slot.
__device__ __forceinline__ int block_ordered_slot(bool flag, int *warp_counts, int *total_inout) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) warp_counts[warp] = __popc(m);
    __syncthreads();
    int base = *total_inout;
    for (int w = 0; w < warp; ++w) base += warp_counts[w];
    const int slot = base + __popc(m & ((1u << lane) - 1));
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kDecThreads / 32; ++w) t += warp_counts[w];
        *total_inout += t;
    }
    __syncthreads();
    return slot;
}

template <bool kDarknet>
__global__ void __launch_bounds__(kDecThreads) decode_nms_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) uint8_t dec_smem_raw[];
    DecodeSmem &s = *reinterpret_cast<DecodeSmem *>(dec_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.x;
    const int D = 5 + p.C, NA = p.GH * p.GW * p.A;
    const float *net = p.logits + (long long)frame * NA * D;

    if (tid == 0) { s.n_cand = 0; s.n_entry = 0; s.n_seg = 0; s.n_out = 0; s.overflow = 0; }

    // ---- (1) keras only: global max and min of every class logit (utils.py:263-266)
    if (!kDarknet) {
        float mx = -INFINITY, mn = INFINITY;
        const long long total = (long long)NA * p.C;
        for (long long t = tid; t < total; t += kDecThreads) {
            const int a = int(t / p.C), k = int(t - (long long)a * p.C);
            const float v = net[(long long)a * D + 5 + k];
            mx = fmaxf(mx, v);
            mn = fminf(mn, v);
        }
        mx = warp_max(mx);
        mn = warp_min(mn);
        if (lane == 0) { s.red_max[warp] = mx; s.red_min[warp] = mn; }
        __syncthreads();
        if (warp == 0) {
            mx = lane < kDecThreads / 32 ? s.red_max[lane] : -INFINITY;
            mn = lane < kDecThreads / 32 ? s.red_min[lane] : INFINITY;
            mx = warp_max(mx);
            mn = warp_min(mn);
            if (lane == 0) { s.gmax = mx; s.gmin_z = __fsub_rn(mn, mx); }   // min(x - max) = min(x) - max
        }
    }
    __syncthreads();

    // ---- (2) objectness gate, ordered compaction of the anchors that can still pass
    // keras order: (row, col, a) = memory order.  darknet order: index = a*G*G + row*G + col.
    for (int base = 0; base < NA; base += kDecThreads) {
        const int i = base + tid;
        bool flag = false;
        float conf = 0.f;
        int anchor = 0;
        if (i < NA) {
            if (kDarknet) {
                const int a = i / (p.GH * p.GW), cell = i - a * (p.GH * p.GW);
                anchor = cell * p.A + a;     // position in the NHWC logits
            } else {
                anchor = i;
            }
            const float to = net[(long long)anchor * D + 4];
            conf = kDarknet ? (float)(1.0 / (1.0 + exp(-(double)to))) : sigmoid_np(to);
            flag = conf > p.obj_thr;
        }
        const int slot = block_ordered_slot(flag, s.warp_counts, &s.n_cand);
        if (flag) {
            if (slot < kMaxCand) {
                s.cand_anchor[slot] = (short)anchor;
                s.cand_conf[slot] = conf;
                s.cand_best[slot] = 0ull;
            } else {
                s.overflow = 1;
            }
        }
    }
    __syncthreads();
    const int n_cand = min(s.n_cand, kMaxCand);

    // ---- (3) one warp per surviving anchor: class probabilities, box
    for (int ci = warp; ci < n_cand; ci += kDecThreads / 32) {
        const int anchor = s.cand_anchor[ci];
        const float conf = s.cand_conf[ci];
        const float *row = net + (long long)anchor * D;
        const int a = anchor % p.A, cell = anchor / p.A;
        const int gy = cell / p.GW, gx = cell - gy * p.GW;
        if (lane == 0) {
            float4 bx;
            if (!kDarknet) {
                bx.x = __fdiv_rn(__fadd_rn((float)gx, sigmoid_np(row[0])), (float)p.GW);
                bx.y = __fdiv_rn(__fadd_rn((float)gy, sigmoid_np(row[1])), (float)p.GH);
                bx.z = __fdiv_rn(__fmul_rn(p.anchors[2 * a], exp_rn(row[2])), (float)p.GW);
                bx.w = __fdiv_rn(__fmul_rn(p.anchors[2 * a + 1], exp_rn(row[3])), (float)p.GH);
            } else {
                const float sx = (float)(1.0 / (1.0 + exp(-(double)row[0])));
                const float sy = (float)(1.0 / (1.0 + exp(-(double)row[1])));
                float x = __fdiv_rn(__fadd_rn((float)gx, sx), (float)p.GW);
                float y = __fdiv_rn(__fadd_rn((float)gy, sy), (float)p.GH);
                float w = (float)(exp((double)row[2]) * (double)p.anchors[2 * a] / p.GW);
                float h = (float)(exp((double)row[3]) * (double)p.anchors[2 * a + 1] / p.GH);
                int new_w, new_h;
                if (((float)p.net_w / p.orig_w) < ((float)p.net_h / p.orig_h)) {
                    new_w = p.net_w;  new_h = (p.orig_h * p.net_w) / p.orig_w;
                } else {
                    new_h = p.net_h;  new_w = (p.orig_w * p.net_h) / p.orig_h;
                }
                x = (float)(((double)x - (p.net_w - new_w) / 2. / p.net_w) / (double)__fdiv_rn((float)new_w, (float)p.net_w));
                y = (float)(((double)y - (p.net_h - new_h) / 2. / p.net_h) / (double)__fdiv_rn((float)new_h, (float)p.net_h));
                w = __fmul_rn(w, __fdiv_rn((float)p.net_w, (float)new_w));
                h = __fmul_rn(h, __fdiv_rn((float)p.net_h, (float)new_h));
                bx.x = __fmul_rn(x, (float)p.orig_w);  bx.z = __fmul_rn(w, (float)p.orig_w);
                bx.y = __fmul_rn(y, (float)p.orig_h);  bx.w = __fmul_rn(h, (float)p.orig_h);
            }
            s.cand_box[ci] = bx;
        }
        // class vector e[k]
        float *e = s.scratch[warp];
        float sum;
        const bool fits = p.C <= 160;
        if (!kDarknet) {
            const float gmax = s.gmax, lo = s.gmin_z;
            if (fits) {
                for (int k = lane; k < p.C; k += 32) {
                    float z = __fsub_rn(row[5 + k], gmax);
                    if (lo < -100.f) z = __fmul_rn(__fdiv_rn(z, lo), -100.f);
                    e[k] = exp_rn(z);
                }
                __syncwarp();
                sum = (lane == 0) ? np_pairwise_sum(e, p.C) : 0.f;
                sum = __shfl_sync(0xffffffffu, sum, 0);
            } else {
                sum = 0.f;   // rare large-C path: sequential (order differs from numpy beyond 160 classes)
                for (int k = 0; k < p.C; ++k) {
                    float z = __fsub_rn(row[5 + k], gmax);
                    if (lo < -100.f) z = __fmul_rn(__fdiv_rn(z, lo), -100.f);
                    sum = __fadd_rn(sum, exp_rn(z));
                }
            }
        } else {
            float mx = -INFINITY;
            for (int k = lane; k < p.C; k += 32) mx = fmaxf(mx, row[5 + k]);
            mx = warp_max(mx);
            if (fits) {
                for (int k = lane; k < p.C; k += 32) e[k] = exp_rn(__fsub_rn(row[5 + k], mx));
                __syncwarp();
            }
            sum = 0.f;
            if (lane == 0)
                for (int k = 0; k < p.C; ++k) sum = __fadd_rn(sum, fits ? e[k] : exp_rn(__fsub_rn(row[5 + k], mx)));
            sum = __shfl_sync(0xffffffffu, sum, 0);
            if (lane == 0) e[160] = mx;   // row max, for the C > 160 recompute path below
            __syncwarp();
        }
        for (int k0 = 0; k0 < p.C; k0 += 32) {
            const int k = k0 + lane;
            float pk = 0.f;
            if (k < p.C) {
                float ek;
                if (fits) {
                    ek = e[k];
                } else if (!kDarknet) {
                    float z = __fsub_rn(row[5 + k], s.gmax);
                    if (s.gmin_z < -100.f) z = __fmul_rn(__fdiv_rn(z, s.gmin_z), -100.f);
                    ek = exp_rn(z);
                } else {
                    ek = exp_rn(__fsub_rn(row[5 + k], e[160]));
                }
                pk = __fmul_rn(conf, __fdiv_rn(ek, sum));
            }
            const bool hit = (k < p.C) && (pk > p.obj_thr);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            int base = 0;
            if (lane == 0 && m) base = atomicAdd(&s.n_entry, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit) {
                const int slot = base + __popc(m & ((1u << lane) - 1));
                if (slot < kMaxEntry) {
                    s.ent_p[slot] = pk;
                    s.ent_cand[slot] = (short)ci;
                    s.ent_cls[slot] = (short)k;
                } else {
                    s.overflow = 1;
                }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    const int n_ent = min(s.n_entry, kMaxEntry);

    // ---- (4) sort.  keras: (class asc, p desc, candidate index desc)  [reversed stable argsort per class]
    //              darknet: one segment, key = (objectness desc, candidate index asc) over CANDIDATES that
    //              own at least ... every gated anchor takes part in do_nms_obj, with or without a class hit.
    if (!kDarknet) {
        for (int i = tid; i < n_ent; i += kDecThreads) {
            const float pi = s.ent_p[i];
            const int ci = s.ent_cand[i], ki = s.ent_cls[i];
            int rank = 0;
            for (int j = 0; j < n_ent; ++j) {
                const float pj = s.ent_p[j];
                const int cj = s.ent_cand[j], kj = s.ent_cls[j];
                const bool before = (kj < ki) || (kj == ki && (pj > pi || (pj == pi && cj > ci)));
                rank += before;
            }
            s.srt_p[rank] = pi;
            s.srt_cand[rank] = (short)ci;
            s.srt_cls[rank] = (short)ki;
            s.alive[rank] = 1;
        }
        __syncthreads();
        for (int i = tid; i < n_ent; i += kDecThreads) {
            if (i == 0 || s.srt_cls[i] != s.srt_cls[i - 1]) {
                const int sg = atomicAdd(&s.n_seg, 1);
                s.seg_start[sg] = (short)i;   // unordered list of segment heads; the end is found by scanning
            }
        }
        __syncthreads();
        // ---- (5) greedy suppression, one warp per class segment (utils.py:236-249)
        for (int sg = warp; sg < s.n_seg; sg += kDecThreads / 32) {
            const int beg = s.seg_start[sg];
            const int cls = s.srt_cls[beg];
            int end = beg + 1;
            while (end < n_ent && s.srt_cls[end] == cls) ++end;
            for (int i = beg; i < end; ++i) {
                if (!s.alive[i]) continue;             // classes[c] == 0 -> skip (uniform across the warp)
                const float4 bi = s.cand_box[s.srt_cand[i]];
                for (int j = i + 1 + lane; j < end; j += 32)
                    if (s.alive[j] && iou_keras(bi, s.cand_box[s.srt_cand[j]]) >= p.nms_thr) s.alive[j] = 0;
                __syncwarp();
            }
        }
        __syncthreads();
        // ---- (6) per box: score = max surviving class prob, label = first argmax (utils.py:123-134)
        for (int i = tid; i < n_ent; i += kDecThreads) {
            if (!s.alive[i]) continue;
            const unsigned long long key =
                ((unsigned long long)__float_as_uint(s.srt_p[i]) << 32) | (0xFFFFFFFFu - (unsigned)s.srt_cls[i]);
            atomicMax(&s.cand_best[s.srt_cand[i]], key);
        }
        __syncthreads();
        // ---- (7) ordered output of boxes with score > obj_thr (utils.py:252-255)
        for (int base = 0; base < n_cand; base += kDecThreads) {
            const int ci = base + tid;
            bool keep = false;
            float score = 0.f;
            int label = 0;
            if (ci < n_cand) {
                const unsigned long long key = s.cand_best[ci];
                score = __uint_as_float((unsigned)(key >> 32));
                label = key ? (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFu)) : 0;
                keep = score > p.obj_thr;
            }
            const int slot = block_ordered_slot(keep, s.warp_counts, &s.n_out);
            if (keep && slot < p.max_boxes) {
                float *o = p.boxes + ((long long)frame * p.max_boxes + slot) * 8;
                const float4 bx = s.cand_box[ci];
                o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
                o[4] = s.cand_conf[ci]; o[5] = score; o[6] = (float)label; o[7] = (float)s.cand_anchor[ci];
            }
        }
        __syncthreads();
        if (tid == 0) p.counts[frame] = s.overflow ? -1 : min(s.n_out, p.max_boxes);
    } else {
        // darknet: rank the gated anchors by objectness (qsort in the reference: ties unordered; here index order)
        short *order = s.srt_cand;                       // order[rank] = candidate
        for (int i = tid; i < n_cand; i += kDecThreads) {
            const float oi = s.cand_conf[i];
            int rank = 0;
            for (int j = 0; j < n_cand; ++j) {
                const float oj = s.cand_conf[j];
                rank += (oj > oi) || (oj == oi && j < i);
            }
            order[rank] = (short)i;
            s.alive[i] = 1;                              // indexed by candidate
        }
        __syncthreads();
        if (warp == 0) {                                  // class-agnostic greedy, one warp (box.c:41-54)
            for (int a = 0; a < n_cand; ++a) {
                const int i = order[a];
                if (!s.alive[i]) continue;
                const float4 bi = s.cand_box[i];
                for (int bq = a + 1 + lane; bq < n_cand; bq += 32) {
                    const int j = order[bq];
                    if (s.alive[j] && iou_darknet(bi, s.cand_box[j]) > p.nms_thr) s.alive[j] = 0;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // rows = (detection, class) entries of surviving detections, sorted by (-prob, detection rank, class)
        // first mark dead entries, then rank-sort the live ones
        short *cand_rank = s.seg_start;                  // candidate -> rank in objectness order
        for (int a = tid; a < n_cand; a += kDecThreads) cand_rank[order[a]] = (short)a;
        __syncthreads();
        for (int base = 0; base < n_ent; base += kDecThreads) {
            const int i = base + tid;
            bool live = false;
            if (i < n_ent) live = s.alive[s.ent_cand[i]] != 0;
            int rank = 0;
            if (live) {
                const float pi = s.ent_p[i];
                const int ri = cand_rank[s.ent_cand[i]], ki = s.ent_cls[i];
                for (int j = 0; j < n_ent; ++j) {
                    if (!s.alive[s.ent_cand[j]]) continue;
                    const float pj = s.ent_p[j];
                    const int rj = cand_rank[s.ent_cand[j]], kj = s.ent_cls[j];
                    rank += (pj > pi) || (pj == pi && (rj < ri || (rj == ri && kj < ki)));
                }
                if (rank < p.max_boxes) {
                    float *o = p.boxes + ((long long)frame * p.max_boxes + rank) * 8;
                    const int ci = s.ent_cand[i];
                    const float4 bx = s.cand_box[ci];
                    o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
                    o[4] = s.cand_conf[ci]; o[5] = pi; o[6] = (float)ki;
                    const int anchor = s.cand_anchor[ci];          // NHWC position -> darknet index a*G*G + cell
                    o[7] = (float)((anchor % p.A) * (p.GH * p.GW) + anchor / p.A);
                }
                atomicAdd(&s.n_out, 1);
            }
        }
        __syncthreads();
        if (tid == 0) p.counts[frame] = s.overflow ? -1 : min(s.n_out, p.max_boxes);
    }
}

int launch_decode(bool darknet, const DecodeParams &p, cudaStream_t st) {
    // the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(decode_nms_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecodeSmem));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(decode_nms_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecodeSmem));
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    if (darknet)
        decode_nms_kernel<true><<<p.B, kDecThreads, sizeof(DecodeSmem), st>>>(p);
    else
        decode_nms_kernel<false><<<p.B, kDecThreads, sizeof(DecodeSmem), st>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace b2t
