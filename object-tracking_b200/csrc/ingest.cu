// Frame ingest: the resize the reference does on the host before every forward pass --
// cv2.resize(image, (IMAGE_W, IMAGE_H)) at KerasYOLO.py:526, MultiObjDetTracker.py:300 -- as a device kernel whose
// output is bit-identical to OpenCV's INTER_LINEAR for 8-bit images.
//
// OpenCV is a third-party dependency of the reference (README.md:12-18, unpinned); its algorithm (imgproc/resize.cpp,
// HResizeLinear / VResizeLinear for uchar with INTER_RESIZE_COEF_BITS = 11) is restated here:
//   fx = float((dx + 0.5) * scale_x - 0.5); sx = floor(fx); fx -= sx; clamp (sx < 0 -> sx = 0, fx = 0;
//   sx >= W-1 -> sx = W-1, fx = 0); a0 = short(rint((1 - fx) * 2048)), a1 = short(rint(fx * 2048));
//   rows the same way without the fx clamp, source rows clipped to [0, H-1];
//   h(row, dx) = S[row][sx] * a0 + S[row][sx+1] * a1                                  (int, 19 bits)
//   dst = (((b0 * (h(sy) >> 4)) >> 16) + ((b1 * (h(sy+1) >> 4)) >> 16) + 2) >> 2.
// The coefficient tables are built on the HOST with exactly these float/double expressions (tiny: W' + H' entries),
// so the device only does integer arithmetic.  tests/: oracle/ingest_oracle.py is pinned against cv2.resize itself.
#include "kernels.cuh"

namespace b2t {

__global__ void __launch_bounds__(256) resize_bilinear_u8_kernel(const ResizeParams p) {
    const long long total = (long long)p.B * p.dst_h * p.dst_w;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int dx = int(t % p.dst_w);
        long long q = t / p.dst_w;
        const int dy = int(q % p.dst_h);
        const int b = int(q / p.dst_h);
        const int4 cx = __ldg(p.xtab + dx);                    // sx, sx1, a0, a1
        const int4 cy = __ldg(p.ytab + dy);                    // sy0, sy1, b0, b1
        const uint8_t *img = p.src + (long long)b * p.src_h * p.src_w * 3;
        const uint8_t *r0 = img + (long long)cy.x * p.src_w * 3, *r1 = img + (long long)cy.y * p.src_w * 3;
        uint8_t *o = p.dst + ((long long)(b * p.dst_h + dy) * p.dst_w + dx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int h0 = int(r0[cx.x * 3 + c]) * cx.z + int(r0[cx.y * 3 + c]) * cx.w;
            const int h1 = int(r1[cx.x * 3 + c]) * cx.z + int(r1[cx.y * 3 + c]) * cx.w;
            const int v = (((cy.z * (h0 >> 4)) >> 16) + ((cy.w * (h1 >> 4)) >> 16) + 2) >> 2;
            o[c] = (uint8_t)min(max(v, 0), 255);
        }
    }
}

// darknet's ingest: load_image_color (image.c:1442-1482: RGB, float(u8)/255.) + letterbox_image (image.c:960-979 =
// resize_image :1347-1389, separable bilinear with the (n-1)/(m-1) scale and darknet's operation order, embedded in a
// 0.5-filled net-sized canvas) for a batch of (src_h, src_w, 3) uint8 frames -> (net_h, net_w, 3) float32 frames, the
// input network_predict_image (network.c:609-616) feeds the net.  Same arithmetic, step by step, as the float CHW
// kernel behind the compat layer's network_predict_image (darknet_compat.cu).
__global__ void __launch_bounds__(256) letterbox_u8_kernel(const LetterboxParams p) {
    const long long per = (long long)p.net_w * p.net_h * 3, total = per * p.B;
    const int ox = (p.net_w - p.new_w) / 2, oy = (p.net_h - p.new_h) / 2;
    const float w_scale = (float)(p.src_w - 1) / (p.new_w - 1), h_scale = (float)(p.src_h - 1) / (p.new_h - 1);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int b = int(t / per), i = int(t - (long long)b * per);
        const int k = i % 3, x = (i / 3) % p.net_w, y = i / (3 * p.net_w);
        const int rx = x - ox, ry = y - oy;
        float v = 0.5f;
        if (rx >= 0 && rx < p.new_w && ry >= 0 && ry < p.new_h) {
            const uint8_t *img = p.src + (long long)b * p.src_h * p.src_w * 3 + (p.swap_rb ? 2 - k : k);
            auto px = [&](int r, int c) { return (float)((double)img[((long long)r * p.src_w + c) * 3] / 255.0); };
            auto hrow = [&](int r) {           // horizontally resized pixel (rx, r)
                if (rx == p.new_w - 1 || p.src_w == 1) return px(r, p.src_w - 1);
                const float sx = rx * w_scale;
                const int ix = (int)sx;
                const float dx = sx - ix;
                return __fadd_rn(__fmul_rn(1.f - dx, px(r, ix)), __fmul_rn(dx, px(r, ix + 1)));
            };
            const float sy = ry * h_scale;
            const int iy = (int)sy;
            const float dy = sy - iy;
            v = __fmul_rn(1.f - dy, hrow(iy));
            if (!(ry == p.new_h - 1 || p.src_h == 1)) v = __fadd_rn(v, __fmul_rn(dy, hrow(iy + 1)));
        }
        p.dst[t] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Callers after the path.
//
// draw_boxes (utility/utils.py:190-206): cv2.rectangle(image, (xmin, ymin), (xmax, ymax), colour, 3) for every decoded
// box, on the device, in place.  OpenCV rasterises a thickness-3 segment as the pixel set
// { max(0, a - u, u - b) + |v - c| <= 2 }  (u along the segment [a, b], v across it at c): a 5-pixel band with
// diamond-cut ends; a rectangle is the union of its four edges.  Pinned against cv2.rectangle itself
// (tests/test_oracle_cpu.py for the rule, tests/test_gpu_parity_r2.py for the kernel).  The text label of
// cv2.putText stays on the host.  Box corners follow the reference's arithmetic under numpy >= 2 (float32):
// int((x -/+ w/2) * W), int((y -/+ h/2) * H), truncation toward zero.
__global__ void draw_boxes_kernel(unsigned char *frames, int B, int H, int W, const float *rows, const int *counts, int max_rows,
                                  unsigned char c0, unsigned char c1, unsigned char c2) {
    const int b = blockIdx.y, k = blockIdx.x;
    if (b >= B || k >= min(counts[b], max_rows)) return;
    const float *r = rows + ((long long)b * max_rows + k) * 8;
    const float hw = __fdiv_rn(r[2], 2.f), hh = __fdiv_rn(r[3], 2.f);
    int xa = (int)__fmul_rn(__fsub_rn(r[0], hw), (float)W), xb = (int)__fmul_rn(__fadd_rn(r[0], hw), (float)W);
    int ya = (int)__fmul_rn(__fsub_rn(r[1], hh), (float)H), yb = (int)__fmul_rn(__fadd_rn(r[1], hh), (float)H);
    if (xa > xb) { const int t = xa; xa = xb; xb = t; }
    if (ya > yb) { const int t = ya; ya = yb; yb = t; }
    unsigned char *img = frames + (long long)b * H * W * 3;
    auto put = [&](int x, int y) {
        if (x >= 0 && x < W && y >= 0 && y < H) {
            unsigned char *p = img + ((long long)y * W + x) * 3;
            p[0] = c0; p[1] = c1; p[2] = c2;
        }
    };
    // horizontal edges at y = ya, yb: u = x in [xa, xb]; vertical edges at x = xa, xb: u = y in [ya, yb]
    const int lh = xb - xa + 5, lv = yb - ya + 5;
    for (int i = threadIdx.x; i < 5 * lh; i += blockDim.x) {
        const int dv = i / lh - 2, x = xa - 2 + i % lh;
        const int out = max(0, max(xa - x, x - xb));
        if (out + abs(dv) <= 2) { put(x, ya + dv); put(x, yb + dv); }
    }
    for (int i = threadIdx.x; i < 5 * lv; i += blockDim.x) {
        const int dv = i / lv - 2, y = ya - 2 + i % lv;
        const int out = max(0, max(ya - y, y - yb));
        if (out + abs(dv) <= 2) { put(xa + dv, y); put(xb + dv, y); }
    }
}

int launch_draw_boxes(unsigned char *frames, int B, int H, int W, const float *rows, const int *counts, int max_rows,
                      int c0, int c1, int c2, cudaStream_t st) {
    draw_boxes_kernel<<<dim3(max_rows, B), 128, 0, st>>>(frames, B, H, W, rows, counts, max_rows, (unsigned char)c0,
                                                         (unsigned char)c1, (unsigned char)c2);
    return (int)cudaGetLastError();
}

// overlap_score / average_overlap_score (utility/utils.py:82-110) for n pairs of corner boxes (x1, y1, x2, y2) in
// float64, the reference's arithmetic (Python floats): |dx*dy| products without clamping an empty intersection, and
// the mean accumulated left to right.  One thread per pair, then thread 0 sums in order (bit-identical to the loop).
__global__ void overlap_scores_kernel(const double *t, const double *p, int n, double *scores, double *mean) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double *a = t + 4 * i, *b = p + 4 * i;
        const double x1 = fmax(a[0], b[0]), y1 = fmax(a[1], b[1]), x2 = fmin(a[2], b[2]), y2 = fmin(a[3], b[3]);
        const double inter = fabs(__dmul_rn(__dsub_rn(x1, x2), __dsub_rn(y1, y2)));
        const double uni = __dsub_rn(__dadd_rn(fabs(__dmul_rn(__dsub_rn(a[0], a[2]), __dsub_rn(a[1], a[3]))),
                                               fabs(__dmul_rn(__dsub_rn(b[0], b[2]), __dsub_rn(b[1], b[3])))), inter);
        scores[i] = __ddiv_rn(inter, uni);
    }
    __syncthreads();
    if (threadIdx.x == 0 && mean) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s = __dadd_rn(s, scores[i]);
        *mean = __ddiv_rn(s, (double)n);
    }
}

int launch_overlap_scores(const double *t, const double *p, int n, double *scores, double *mean, cudaStream_t st) {
    overlap_scores_kernel<<<1, 256, 0, st>>>(t, p, n, scores, mean);
    return (int)cudaGetLastError();
}

int launch_letterbox_u8(const LetterboxParams &p, cudaStream_t st) {
    const long long total = (long long)p.B * p.net_h * p.net_w * 3;
    const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    letterbox_u8_kernel<<<blocks, 256, 0, st>>>(p);
    return (int)cudaGetLastError();
}

int launch_resize_bilinear_u8(const ResizeParams &p, cudaStream_t st) {
    const long long total = (long long)p.B * p.dst_h * p.dst_w;
    const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    resize_bilinear_u8_kernel<<<blocks, 256, 0, st>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace b2t
