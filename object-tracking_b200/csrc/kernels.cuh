// Parameter blocks and host launchers of every kernel in the library (one definition, shared by the kernels'
// translation units and api.cu).
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "conv_types.cuh"

namespace b2t {

struct SimtView {
    const op_t *a_hi;  long long a_plane;  int a_pix_stride;   // activations (channel offset pre-applied)
    const op_t *w_hi;  long long w_plane;  int w_ld;           // weights [Cout][taps*cin_pad]
};

struct Conv1Params {
    const void *frames;   // (B,H,W,3) uint8 or float32
    int dtype;            // 0 = u8, 1 = f32
    int B, H, W;
    const float *w;       // [27][32] fp32 (tap-major: (kh*3+kw)*3 + cin)
    const float *scale, *bias;
    const float *lut;     // [256] = float(u / 255.0)
    Dest out;             // full-res (optional)
    Dest pout;            // pooled
};

struct DecodeParams {
    const float *logits;   // (B, G, G, A, D) fp32
    int B, GH, GW, A, C;
    float obj_thr, nms_thr;
    float anchors[32];
    float *boxes;          // (B, max_boxes, 8)
    int *counts;           // (B)
    int max_boxes;
    int orig_w, orig_h, net_w, net_h;   // darknet flavour only
};

struct LstmParams {
    const float *wp;      // [units/4][n_in + units][4 gates][4 units]
    const float *bias;    // [4*units] keras order
    const float *fv;      // (S, n_feat)
    const float *det;     // (S, n_det)
    const float *h_in;    // (S, units)
    float *h_out;         // (S, units)
    float *c;             // (S, units) in place
    int n_feat, n_det, units, S;
    int fv_stride, det_stride;   // elements between consecutive streams' rows
    int hard_sigmoid;
    int mode;             // 0 = full step; 1 = input projection only (zx_out = x*W + b, no state change);
                          // 2 = recurrent step on a precomputed projection (z = zx_in + h*U)
    float *zx;            // (rows, 4*units) projection buffer (mode 1: out, mode 2: in)
    int zx_stride;        // elements between consecutive streams' rows of zx
    float *h_seq;         // optional copy of h' (mode 0/2), row stride h_seq_stride
    int h_seq_stride;
};

struct PoolParams {
    const op_t *hi;  long long plane;  int pix_stride, ch_off;
    int B, H, W, C;
    int mode;       // 0 = global max -> (B,C); 1 = 4x4/4 max + flatten -> (B,(H/4)*(W/4)*C)
    int chw_view;   // 1 = read the tensor as the reference does: CHW buffer reshaped (H,W,C) without transpose
    float *out;
};

struct SelectParams {
    const float *dets;  const int *counts;  int max_dets, B;
    const unsigned char *class_mask;   // device, n_class bytes, or NULL = all classes
    int frame_w, frame_h;
    float *det_in;      // (B,4)
    int heat_size;  float *heat;       // (B, size*size) or NULL
    int *chosen;        // (B) row index or -1
};

struct ResizeParams {
    const unsigned char *src;   // (B, src_h, src_w, 3) uint8
    unsigned char *dst;         // (B, dst_h, dst_w, 3) uint8
    int B, src_h, src_w, dst_h, dst_w;
    const int4 *xtab;           // per destination column: source column, next source column, a0, a1 (2048 = 1.0)
    const int4 *ytab;           // per destination row:    source row,    next source row,    b0, b1
};

struct LetterboxParams {
    const unsigned char *src;   // (B, src_h, src_w, 3) uint8
    float *dst;                 // (B, net_h, net_w, 3) float32
    int B, src_h, src_w, net_h, net_w;
    int new_w, new_h;           // size of the embedded resized image (darknet's integer arithmetic, host-computed)
    int swap_rb;                // 1 = source is BGR (cv2.imread), emit RGB like load_image_color
};

struct ConvLstmGateParams {
    const float *g;       // (S*T, M, 4u) pre-activations (input conv + bias + recurrent conv), frame = s*T + t
    float *c;             // (slots, M, u) cell state, in place
    Dest h_rec;           // h' as split planes for the next recurrent conv (image = state slot)
    Dest h_seq;           // h' as split planes at frame s*T + t (input of the 1x1 head)
    int M, units, G;      // M = G*G pixels of one stream
    int S, T, t;          // streams of this call, steps per stream in the gate buffer, current step
    int slot0;            // state slot of stream 0
    int hard_sigmoid;
};

// Launch with programmatic stream serialization (the kernel MUST call griddep_wait() before touching anything its
// predecessor wrote).  Developer builds: B2T_PDL=0 falls back to a plain launch.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
#ifdef B2T_DEV
    static const bool pdl = !(getenv("B2T_PDL") && atoi(getenv("B2T_PDL")) == 0);
#else
    constexpr bool pdl = true;
#endif
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

int launch_conv_umma(int BN, const CUtensorMap &a_hi, const CUtensorMap &a_lo, const CUtensorMap &b_hi,
                     const CUtensorMap &b_lo, const ConvParams &p, cudaStream_t st);
int conv_umma_init();
int conv_halo_init();
int launch_conv_halo(int n_sm, bool small, const CUtensorMap &x_hi, const CUtensorMap &x_lo, const CUtensorMap &w_hi,
                     const CUtensorMap &w_lo, const ConvParams &p, cudaStream_t st);
int launch_conv_halo_persist(int n_sm, const CUtensorMap &x_hi, const CUtensorMap &x_lo, const CUtensorMap &w_hi,
                             const CUtensorMap &w_lo, const ConvParams &p, cudaStream_t st);
// conv_chain_kernel: a run of layers in one persistent cooperative launch (small batches); parameters built layer by layer
struct ChainBuilder;
ChainBuilder *chain_new(unsigned int *counter_dev);
void chain_free(ChainBuilder *b);
int chain_add(ChainBuilder *b, const CUtensorMap &x_hi, const CUtensorMap &x_lo, const CUtensorMap &w_hi,
              const CUtensorMap &w_lo, const ConvParams &p);      // -1 = full
int chain_layers(const ChainBuilder *b);
int launch_conv_chain(int n_sm, const ChainBuilder *b, cudaStream_t st);
int conv_pm_init();
int conv_pm_smem_bytes(const ConvParams &p);
int launch_conv_pm(int n_sm, const CUtensorMap &x_hi, const CUtensorMap &x_lo, const CUtensorMap &w_hi,
                   const CUtensorMap &w_lo, const ConvParams &p, cudaStream_t st);
int launch_frames_to_c8(const void *frames, void *dst, long long npix, long long seg_pix, long long seg_stride, cudaStream_t st);
int launch_splitk_epilogue(const ConvParams &p, cudaStream_t st);
int launch_conv_simt(const SimtView &v, const ConvParams &p, cudaStream_t st);
int launch_conv1(const Conv1Params &p, cudaStream_t st);
int launch_planes_to_f32(const op_t *hi, long long plane, int pix_stride, int ch_off, int C, long long npix,
                         float *out, cudaStream_t st);
int launch_reorg_gather(const op_t *in, long long in_plane, int in_stride, op_t *out, long long out_plane, int out_stride, int B,
                        int H, int W, int C, cudaStream_t st);
int launch_pool_s1(const op_t *in, long long in_plane, op_t *out, long long out_plane, int B, int H, int W, int C, cudaStream_t st);
int launch_decode(bool darknet, const DecodeParams &p, cudaStream_t st);
int launch_lstm_gates(const LstmParams &p, cudaStream_t st);
int launch_lstm_proj(const LstmParams &p, cudaStream_t st);      // -1 = shape not supported, use lstm_gates mode 1
int launch_lstm_seq(const LstmParams &p, int T, float *h_a, float *h_b, unsigned int *counter, int n_sm, cudaStream_t st);   // -1 = fall back
int launch_dense_sigmoid(const float *h, const float *wd, const float *bd, int units, int n_out, int S, float *y,
                         int y_stride, cudaStream_t st);
int launch_pool_features(const PoolParams &p, cudaStream_t st);
int launch_heatmap_from_box(const float *xywh, int n, int size, float *heat, cudaStream_t st);
int launch_select_detection(const SelectParams &p, cudaStream_t st);
int launch_box_from_heatmap(const float *heat, int n, int size, float thresh, int *rect, cudaStream_t st);
int launch_resize_bilinear_u8(const ResizeParams &p, cudaStream_t st);
int launch_letterbox_u8(const LetterboxParams &p, cudaStream_t st);
int launch_convlstm_gates(const ConvLstmGateParams &p, cudaStream_t st);
int launch_draw_boxes(unsigned char *frames, int B, int H, int W, const float *rows, const int *counts, int max_rows,
                      int c0, int c1, int c2, cudaStream_t st);
int launch_overlap_scores(const double *t, const double *p, int n, double *scores, double *mean, cudaStream_t st);

}  // namespace b2t
