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
