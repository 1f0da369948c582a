// Small product kernels shared by every conv path: the split-K finish, conv_1 for float32 frames and the
// split-plane -> fp32 read-out.
#include "kernels.cuh"

namespace b2t {

// ------------------------------------------------------------------------------------------------
// Finishes a split-K convolution (splitk_finish_range, conv_types.cuh).
__global__ void __launch_bounds__(256) splitk_epilogue_kernel(const ConvParams p) {
    griddep_launch();
    griddep_wait();              // the partials come from the previous kernel of the stream
    splitk_finish_range(p, blockIdx.x * (long long)blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

// ------------------------------------------------------------------------------------------------
// conv_1 (3 -> 32, 3x3) fused with the input normalisation (image/255, utils.py:150-153), BN, LeakyReLU
// and the 2x2 max-pool.  K = 27 is no tensor-core shape and the layer is HBM-bound (0.5 MB in, 5.5 MB out
// per frame), so this is a direct fp32 convolution: one thread = one pooled pixel x 16 output channels.

// Block = 128 threads = 2 pooled rows x 32 pooled columns x 2 channel halves.  The (6 x 66)-pixel input patch is
// staged in shared memory as normalised floats (coalesced byte loads, LUT for u8), the 27x32 weights likewise; the
// tap loop over kh stays rolled so the kernel body fits the instruction cache (a fully unrolled 1728-FMA body
// thrashes it and runs 5x slower).
constexpr int kC1W = 32, kC1H = 2;                       // pooled pixels per block
constexpr int kC1PW = 2 * kC1W + 2, kC1PH = 2 * kC1H + 2; // input patch incl. halo
__global__ void __launch_bounds__(128, 4) conv1_direct_kernel(const Conv1Params p) {
    __shared__ __align__(16) float sw[27 * 32];
    __shared__ float sscale[32], sbias[32];
    __shared__ float spatch[kC1PH][kC1PW][3];
    const int Hq = p.H / 2, Wq = p.W / 2;
    const int tiles_x = (Wq + kC1W - 1) / kC1W, tiles_y = (Hq + kC1H - 1) / kC1H;
    int t = blockIdx.x;
    const int tx = t % tiles_x;  t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int xq0 = tx * kC1W, yq0 = ty * kC1H;
    for (int i = threadIdx.x; i < 27 * 32; i += 128) sw[i] = p.w[i];
    if (threadIdx.x < 32) { sscale[threadIdx.x] = p.scale[threadIdx.x]; sbias[threadIdx.x] = p.bias[threadIdx.x]; }
    {   // stage the patch: rows 2*yq0-1 .. 2*yq0+2*kC1H, cols 2*xq0-1 .. 2*xq0+2*kC1W, 3 channels (contiguous bytes)
        const int y_lo = 2 * yq0 - 1, x_lo = 2 * xq0 - 1;
        for (int i = threadIdx.x; i < kC1PH * kC1PW * 3; i += 128) {
            const int r = i / (kC1PW * 3), rem = i - r * (kC1PW * 3);
            const int col = rem / 3, ch = rem - col * 3;
            const int yy = y_lo + r, xx = x_lo + col;
            float v = 0.f;
            if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
                const long long off = (((long long)b * p.H + yy) * p.W + xx) * 3 + ch;
                // u8: image/255. -- fp32 division is bit-identical to numpy's float64 division rounded to fp32
                // for all 256 byte values (checked exhaustively in tests/test_oracle_cpu.py)
                v = p.dtype == 0 ? __fdiv_rn((float)__ldg(reinterpret_cast<const uint8_t *>(p.frames) + off), 255.f)
                                 : __ldg(reinterpret_cast<const float *>(p.frames) + off);
            }
            spatch[r][col][ch] = v;
        }
    }
    __syncthreads();
    // channel half is warp-uniform (warps 0,1 -> channels 0..15, warps 2,3 -> 16..31): the weight reads below are
    // shared-memory broadcasts (1 wavefront) instead of 4-way split loads
    const int half = threadIdx.x >> 6, lx = threadIdx.x % kC1W, ly = (threadIdx.x >> 5) & 1;
    const int xq = xq0 + lx, yq = yq0 + ly;
    // packed fp32x2 FMAs (sm_100 FFMA2): two output channels per instruction, each lane an IEEE fma
    float2 acc2[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc2[k][i] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 *wr = reinterpret_cast<const float4 *>(&sw[((kh * 3 + kw) * 3 + c) * 32 + half * 16]);
                float2 w2[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 f = wr[i];
                    w2[2 * i] = make_float2(f.x, f.y);
                    w2[2 * i + 1] = make_float2(f.z, f.w);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float a = spatch[2 * ly + (k >> 1) + kh][2 * lx + (k & 1) + kw][c];
                    const float2 aa = make_float2(a, a);
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc2[k][i] = __ffma2_rn(aa, w2[i], acc2[k][i]);
                }
            }
    }
    float acc[4][16];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[k][2 * i] = acc2[k][i].x; acc[k][2 * i + 1] = acc2[k][i].y; }
    if (xq >= Wq || yq >= Hq) return;
    float mx[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) mx[i] = -INFINITY;
    const bool want_full = p.out.hi || p.out.f32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            acc[k][i] = leaky(fmaf(acc[k][i], sscale[half * 16 + i], sbias[half * 16 + i]));
            mx[i] = fmaxf(mx[i], acc[k][i]);
        }
        if (want_full) {
            const float(&lo8)[8] = *reinterpret_cast<const float(*)[8]>(&acc[k][0]);
            const float(&hi8)[8] = *reinterpret_cast<const float(*)[8]>(&acc[k][8]);
            emit8(p.out, b, 2 * yq + (k >> 1), 2 * xq + (k & 1), half * 16, 32, lo8);
            emit8(p.out, b, 2 * yq + (k >> 1), 2 * xq + (k & 1), half * 16 + 8, 32, hi8);
        }
    }
    const float(&m0)[8] = *reinterpret_cast<const float(*)[8]>(&mx[0]);
    const float(&m1)[8] = *reinterpret_cast<const float(*)[8]>(&mx[8]);
    emit8(p.pout, b, yq, xq, half * 16, 32, m0);
    emit8(p.pout, b, yq, xq, half * 16 + 8, 32, m1);
}

// split planes -> fp32 NHWC (KerasYOLO.extract / network_extract_feat read-out)
__global__ void planes_to_f32_kernel(const op_t *hi, long long plane, int pix_stride, int ch_off, int C,
                                     long long npix, float *out) {
    const long long total = npix * C;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long pix = t / C;
        const int c = int(t - pix * C);
        const op_t *q = hi + pix * pix_stride + ch_off + c;
        out[t] = join_f16(q[0], q[plane]);
    }
}

// maxpool size 2, stride 1 of the tiny graph (darknet/src/parser.c:471-486: padding (size-1)/2 = 0;
// maxpool_layer.c:79-114: out[i][j] = max over in[i..i+1][j..j+1] inside the image) on split planes.  hi + lo is an
// exact fp32 value (22 significant bits), so joining, comparing and re-splitting is lossless.
__global__ void __launch_bounds__(256) pool_s1_kernel(const op_t *in, long long in_plane, op_t *out, long long out_plane, int B, int H,
                                                      int W, int C) {
    const int groups = C / 8;
    const long long total = (long long)B * H * W * groups;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int g = int(t % groups);
        const long long pix = t / groups;
        const int x = int(pix % W), y = int((pix / W) % H);
        float m[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
                if (y + dy >= H || x + dx >= W) continue;
                const op_t *q = in + (pix + (long long)dy * W + dx) * C + g * 8;
                const uint4 h4 = *reinterpret_cast<const uint4 *>(q), l4 = *reinterpret_cast<const uint4 *>(q + in_plane);
                const op_t *hh = reinterpret_cast<const op_t *>(&h4), *ll = reinterpret_cast<const op_t *>(&l4);
#pragma unroll
                for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], join_f16(hh[i], ll[i]));
            }
        __align__(16) op_t oh[8], ol[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split_f16(m[i], oh[i], ol[i]);
        op_t *d = out + pix * C + g * 8;
        *reinterpret_cast<uint4 *>(d) = *reinterpret_cast<const uint4 *>(oh);
        *reinterpret_cast<uint4 *>(d + out_plane) = *reinterpret_cast<const uint4 *>(ol);
    }
}

// darknet's reorg layer (reorg_layer.c:91-110 -> reorg_cpu(x, w, h, c, batch, stride = 2, forward = 0, out), blas.c:9-30) as
// a GATHER: the (C, H, W) source viewed flat; destination flat index d = i + W*(j + H*k) (k < C, j < H, i < W) takes
// source flat index s = w2 + 2W*(h2 + 2H*c2) with c2 = k % (C/4), off = k / (C/4), w2 = 2i + off % 2, h2 = 2j + off / 2;
// the destination buffer is then read as (4C, H/2, W/2).  One thread = 8 consecutive destination channels of one
// destination pixel (NHWC planes): 16-byte stores, 2-byte gathered loads out of an L2-resident tensor.
__global__ void __launch_bounds__(256) reorg_gather_kernel(const op_t *in, long long in_plane, int in_stride, op_t *out,
                                                           long long out_plane, int out_stride, int B, int H, int W, int C) {
    const int Hd = H / 2, Wd = W / 2, Cd = 4 * C, groups = Cd / 8, out_c = C / 4;
    const long long total = (long long)B * Hd * Wd * groups;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int g = int(t % groups);
        const int pix = int(t / groups);                    // (b, yd, xd)
        const int b = small_div(pix, Hd * Wd), rem = pix - b * Hd * Wd;
        const int yd = small_div(rem, Wd), xd = rem - yd * Wd;
        __align__(16) op_t oh[8], ol[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int cd = g * 8 + e;
            const int d = (cd * Hd + yd) * Wd + xd;         // flat index in the (4C, H/2, W/2) destination = (C, H, W) loop space
            const int row = small_div(d, W), i = d - row * W;
            const int k = small_div(row, H), j = row - k * H;
            const int off = small_div(k, out_c), c2 = k - off * out_c;
            const int w2 = 2 * i + (off & 1), h2 = 2 * j + (off >> 1);
            const int s = w2 + 2 * W * (h2 + 2 * H * c2);   // flat index in the (C, H, W) source
            const int sr = small_div(s, W), xs = s - sr * W;
            const int cs = small_div(sr, H), ys = sr - cs * H;
            const op_t *q = in + (((long long)b * H + ys) * W + xs) * in_stride + cs;
            oh[e] = q[0];
            ol[e] = q[in_plane];
        }
        op_t *dst = out + (long long)pix * out_stride + g * 8;
        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(oh);
        *reinterpret_cast<uint4 *>(dst + out_plane) = *reinterpret_cast<const uint4 *>(ol);
    }
}

// ---------------------------------------------------------------- host-side launchers (used by api.cu)
int launch_reorg_gather(const op_t *in, long long in_plane, int in_stride, op_t *out, long long out_plane, int out_stride, int B,
                        int H, int W, int C, cudaStream_t st) {
    const long long total = (long long)B * (H / 2) * (W / 2) * (4 * C / 8);
    const int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    reorg_gather_kernel<<<blocks, 256, 0, st>>>(in, in_plane, in_stride, out, out_plane, out_stride, B, H, W, C);
    return (int)cudaGetLastError();
}
int launch_pool_s1(const op_t *in, long long in_plane, op_t *out, long long out_plane, int B, int H, int W, int C, cudaStream_t st) {
    const long long total = (long long)B * H * W * (C / 8);
    const int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    pool_s1_kernel<<<blocks, 256, 0, st>>>(in, in_plane, out, out_plane, B, H, W, C);
    return (int)cudaGetLastError();
}
int launch_splitk_epilogue(const ConvParams &p, cudaStream_t st) {
    const int cgroups = (p.Cout + 7) / 8;
    const long long total = (long long)p.B * (p.pool ? p.H / 2 : p.H) * (p.pool ? p.W / 2 : p.W) * cgroups;
    const int blocks = (int)min((long long)148 * 16, (total + 255) / 256);
    return (int)launch_pdl(splitk_epilogue_kernel, dim3(blocks), dim3(256), 0, st, p);
}
int launch_conv1(const Conv1Params &p, cudaStream_t st) {
    const int Hq = p.H / 2, Wq = p.W / 2;
    const long long blocks = (long long)p.B * ((Wq + kC1W - 1) / kC1W) * ((Hq + kC1H - 1) / kC1H);
    conv1_direct_kernel<<<(unsigned)blocks, 128, 0, st>>>(p);
    return (int)cudaGetLastError();
}
int launch_planes_to_f32(const op_t *hi, long long plane, int pix_stride, int ch_off, int C, long long npix,
                         float *out, cudaStream_t st) {
    const long long total = npix * C;
    const int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    planes_to_f32_kernel<<<blocks, 256, 0, st>>>(hi, plane, pix_stride, ch_off, C, npix, out);
    return (int)cudaGetLastError();
}

}  // namespace b2t
