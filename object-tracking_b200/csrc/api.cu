// C-ABI of libb200track.so (include/b200track.h): context, weight packing, TMA descriptors, launch plan.
// Host code only orchestrates; every numeric step runs in the kernels of conv_halo.cu / conv_pm.cu / decode_nms.cu /
// tracker.cu.  There is no CPU fallback: without a device every compute entry point fails with an error.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "../../include/b200track.h"
#include "kernels.cuh"

using namespace b2t;

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CK(expr)                                                                                           \
    do {                                                                                                   \
        cudaError_t e__ = (cudaError_t)(expr);                                                             \
        if (e__ != cudaSuccess) return fail(-2, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
    } while (0)

// shared with darknet_compat.cu
int b2t_fail_internal(int code, const char *msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}
extern "C" const char *b2t_last_error(void) { return g_err; }
extern "C" int b2t_version(void) { return 200; }

// Developer switches (tile geometry / split overrides, trace stamps, the SIMT and first-generation tile engines) exist
// only in `make DEV=1` builds.  The release library reads no environment variable and has one conv path.
#ifdef B2T_DEV
extern "C" int b2t_dev_build(void) { return 1; }
static int dev_env(const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; }
#else
extern "C" int b2t_dev_build(void) { return 0; }
static inline int dev_env(const char *, int dflt) { return dflt; }
#endif

// ------------------------------------------------------------------------------------------------ plan
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

struct ActBuf {              // split-plane activation tensor sized for max_batch frames
    size_t off = 0;          // byte offset of the hi plane inside the workspace
    int H = 0, W = 0, C = 0; // C = channels per pixel (pix_stride)
    long long plane = 0;     // elements between hi and lo plane
    op_t *hi = nullptr;
};

struct ConvLayer {
    int index = 0, k = 1, cin = 0, cin_pad = 0, cout = 0;
    int kchunk = 64;                              // channels per K chunk: 64 (128-byte rows) or 32 (64-byte rows)
    bool act = true, pool = false;
    int H = 0, W = 0;
    int in_buf = -1, in_ch_off = 0;               // input view
    int out_buf = -1, out_ch_off = 0, out_mode = DEST_PLAIN;   // full-res split-plane destination (-1 none)
    int pout_buf = -1;                            // pooled split-plane destination
    int f32_out = 0;                              // 1 = logits, 2 = convlstm gates, 3 = tracker logits
    int f32_accumulate = 0;
    size_t off_whi = 0, off_wlo = 0, off_scale = 0, off_bias = 0;
    int ldw = 0;
    int TW = 16, TH = 8, BN = 128;
    CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;      // developer tile engine (dev_engines.cu)
    int hC = 0, hP = 0, hR = 0, hN = 0, h_rows = 0, h_plane_bytes = 0;   // halo engine tile (conv_halo.cu)
    bool h_small = false;                         // two-CTAs-per-SM resource shape
    CUtensorMap tmX_hi, tmX_lo, tmW_hi, tmW_lo;
    CUtensorMap tmWp_hi, tmWp_lo;                 // persistent variant: weight box of w_rows = min(128, round_up(Cout, 64)) rows
    int w_rows = 128;
    // pixel-major persistent variant (conv_pm.cu) for the narrow layers
    bool pm = false, pm_flat = false;             // flat: 1x1 conv over the pixel list of the whole batch
    int pmC = 0, pmP = 0, pmR = 0, pm_rows = 0, pm_plane_bytes = 0;
    CUtensorMap tmXp_hi, tmXp_lo;
    bool flat1x1 = false;                         // 1x1 conv, plain destination, no pool: persistent kernel over the flat pixel list
    CUtensorMap tmXf_hi, tmXf_lo;                 // its patch view: [cin_pad][MB*H*W] with a box of 128 pixels
    bool have_weights = false;
    int ps1_buf = -1;                             // tiny graph: a 2x2 stride-1 max-pool of out_buf follows (maxpool_layer.c:79-114)
    int reorg_buf = -1;                           // darknet semantics, conv_21: out_buf is a plain tensor, reorg_gather_kernel
                                                  // permutes it into channels [0, 4*cout) of this buffer (blas.c:9-30)
};

struct b2t_ctx {
    b2t_config cfg;
    int G = 0, A = 5, D = 0;
    int keep_prepool = 0;
    std::vector<ActBuf> bufs;
    std::map<std::string, int> buf_by_name;       // "norm_k" -> buffer
    std::map<std::string, int> cout_by_name;      // true channel count of that tensor
    std::map<std::string, int> choff_by_name;
    std::vector<ConvLayer> conv;                  // [0] unused, 1..23, then convlstm layers
    int L_CIN = 0, L_CREC = 0, L_HEAD = 0;        // indices of the ConvLSTM layers (0 = absent)
    int n_conv = 23;                              // conv layers of the detector graph: 23 (cfg/yolov2*.cfg) or 9 (cfg/yolov2-tiny*.cfg)
    int conv1_cout = 32;                          // real output channels of conv_1 (the kernel computes 32; tiny: 16 + 16 zero)
    // conv_1
    size_t off_w1 = 0, off_s1 = 0, off_b1 = 0, off_lut = 0;
    size_t off_w1pm = 0, off_s1pm = 0;           // conv_1 on the tensor cores: packed fp16 (hi,lo) weights, scale / 255
    size_t off_c8 = 0;                            // workspace: frames as fp16 integers, 8 channels per pixel
    // memory
    size_t weight_bytes = 0, ws_bytes = 0;
    std::vector<uint8_t> host_blob;
    uint8_t *d_blob = nullptr, *d_ws = nullptr;
    bool own_blob = false, own_ws = false;
    size_t off_logits = 0, off_partial = 0, partial_bytes = 0;
    size_t off_gates = 0, off_cstate = 0;
    int buf_hrec = -1, buf_hseq = -1, buf_z = -1;
    bool finalized = false;
    long launches = 0;
    int n_sm = 148;
    int chain_max_batch = 1;                      // batches up to this size run conv_2..23 in conv_chain_kernel (0 = never)
    bool chain_forced = false;                    // reserved[1] given explicitly: no image-size condition
    unsigned int *d_chain_counter = nullptr;      // its grid-barrier arrival counter
    long capture_launches0 = 0;                   // b2t_graph_begin: launch counter at the start of the capture
    PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    // b2t_resize_frames: coefficient tables of the last (src, dst) geometry
    int rs_geom[4] = {0, 0, 0, 0};
    int4 *d_rs_tab = nullptr;
};

static int add_buf(b2t_ctx *c, const std::string &name, int H, int W, int C) {
    ActBuf b;
    b.H = H; b.W = W; b.C = C;
    b.plane = (long long)c->cfg.max_batch * H * W * C;
    c->bufs.push_back(b);
    const int id = (int)c->bufs.size() - 1;
    if (!name.empty()) c->buf_by_name[name] = id;
    return id;
}

static void choose_tile(int H, int W, bool pool, int &TW, int &TH) {
    double best = -1;
    for (int tw = 4; tw <= 128; tw *= 2) {
        const int th = 128 / tw;
        if (pool && (tw > 16 || (th & 1))) continue;
        const double eff = (double)H * W / ((double)((W + tw - 1) / tw * tw) * ((H + th - 1) / th * th));
        if (eff > best + 1e-9 || (fabs(eff - best) <= 1e-9 && tw > TW)) { best = eff; TW = tw; TH = th; }
    }
}

// Halo-engine tile: hR rows x hC columns of one image, patch pitch hP, MMA N = round_up(hR*hP, 16) <= 256,
// patch (with halo) <= 256 rows of 128 bytes per plane.  Picks the geometry with the least wasted MMA columns.
static void choose_halo_tile(int H, int W, int ksize, bool pool, ConvLayer &l) {
    const int pad = ksize / 2, taps = ksize * ksize;
    // short K and many tiles -> the small shape (two CTAs per SM overlap each other's prologue/epilogue)
    const int small_mode = dev_env("B2T_SMALL", -1);
    const int kbytes = l.kchunk * 2;
    l.h_small = small_mode >= 0 ? (small_mode != 0 && H >= 26) : (l.cin_pad / l.kchunk * taps <= 36 && H >= 52);
    const int max_rows = l.h_small ? 176 : 256, max_n = l.h_small ? 128 : 256;
    double best = 1e30;
    for (int nx = 1; nx <= 16; ++nx) {
        int C = (W + nx - 1) / nx;
        if (pool && (C & 1)) ++C;
        const bool wrap = pad && nx == 1;                  // full-width tile: right halo column = next row's left halo
        const int P = wrap ? W + 1 : C + 2 * pad;
        if (P > 256) continue;
        if (pool && (P & 1)) continue;
        for (int R = 1; R <= H; ++R) {
            if (pool && (R & 1)) continue;
            const int rows = R + 2 * pad + (wrap ? 1 : 0);
            const int N = round_up(R * P, 16);
            if (rows * P > max_rows || N > max_n || (pool && N > 240)) break;
            const int tiles_x = (W + C - 1) / C, tiles_y = (H + R - 1) / R;
            // cycles per (tap, chunk) of one tile: 12 MMAs of 128 x N x 16 vs. the L2->SM stream (~42 B/clk/SM)
            const double t_mma = 6.0 * N * kbytes / 128, t_l2 = (256.0 * kbytes + 2.0 * rows * P * kbytes / taps) / 42.0;
            const double cost = (t_mma > t_l2 ? t_mma : t_l2) * tiles_x * tiles_y;
            if (cost < best - 1e-9) {
                best = cost;
                l.hC = C; l.hP = P; l.hR = R; l.hN = N; l.h_rows = rows;
                l.h_plane_bytes = (int)align_up((size_t)rows * P * kbytes, 1024);
            }
        }
    }
}

// Pixel-major tile (conv_pm.cu): hR rows x hC columns with hR * hP <= 128 (the MMA's M), fewest tiles wins.
// row_bytes = bytes of one patch row (K chunk of one pixel); mode 1 (conv_1) rows cover 4 pixels.
static void choose_pm_tile(int H, int W, int ksize, bool pool, int row_bytes, bool ints, ConvLayer &l) {
    const int pad = ksize / 2;
    long best = -1;
    for (int C = 1; C <= W && C <= 128; ++C) {
        if (pool && (C & 1)) continue;
        const int P = C + 2 * pad;
        if (P > 128 || (pool && (P & 1))) continue;
        int R = 128 / P;
        if (R > H) R = H;
        if (pool) R &= ~1;
        if (R < 1) continue;
        const long tiles = (long)((W + C - 1) / C) * ((H + R - 1) / R);
        if (best < 0 || tiles < best) {
            best = tiles;
            l.pmC = C; l.pmP = P; l.pmR = R; l.pm_rows = R + 2 * pad;
            const int reach = 128 + 2 * pad * P + 2 * pad + (ints ? 2 : 0);      // rows the shifted operands can touch
            const int rows = l.pm_rows * P > reach ? l.pm_rows * P : reach;
            l.pm_plane_bytes = (int)align_up((size_t)rows * row_bytes, 1024);
        }
    }
}

static unsigned magic_for(int d) { return d <= 1 ? 0u : (unsigned)((0x100000000ull + (unsigned)d - 1) / (unsigned)d); }

static int choose_splits_halo(int ctas, int cin_chunks, int taps, int n_sm, int batch, bool capped = true) {
    const int force = dev_env("B2T_SPLITS", 0);
    int best_s = 1;
    double best_cost = 1e30;
    const int max_s = cin_chunks > 32 ? 32 : cin_chunks;
    for (int s = 1; s <= max_s; ++s) {
        const int per = (cin_chunks + s - 1) / s;
        if ((cin_chunks + per - 1) / per != s) continue;
        if (force > 0 && s != force && s != max_s) continue;
        // the tensor core truncates addends to the accumulator's exponent: keep one accumulation chain short
        static const int chain_cap = dev_env("B2T_CHAIN", 320);
        if (capped && per * taps * 4 > chain_cap && s < max_s) continue;
        const long waves = ((long)ctas * s + n_sm - 1) / n_sm;
        // split-K finish = a second kernel + one write and s reads of the partials: its bytes grow with the batch (3 units
        // per split fits 36 frames; at <= 8 frames the partials are a few MB and stay in L2)
        const double red = s > 1 ? (batch <= 8 ? 1.0 + 0.25 * s : 3.0 * s) : 0.0;
        const double cost = (double)waves * (per * taps + 10.0) + red;
        if (cost < best_cost - 1e-9) { best_cost = cost; best_s = s; }
    }
    return best_s;
}

// K split of conv_chain_kernel (small batches): in units of one (channel chunk, tap) pair = 12 MMAs.  The split-K
// reduction costs a grid barrier + an in-place finish there, not a second kernel, and idle SMs are the dominant loss at
// batch 1, so the split is much finer than choose_splits_halo's.  Cost in units: one wave of items = per + ~4 (first
// patch latency + epilogue), a reduction = ~6 + 0.15 per split.
static int choose_splits_chain(int ctas, int units, int n_sm) {
    static const int chain_cap = dev_env("B2T_CHAIN", 320);
    int best_s = 1;
    double best_cost = 1e30;
    for (int s = 1; s <= 32 && s <= units; ++s) {
        const int per = (units + s - 1) / s;
        if ((units + per - 1) / per != s) continue;
        if (per * 4 > chain_cap && s < 32 && s < units) continue;    // keep one accumulation chain short
        const long waves = ((long)ctas * s + n_sm - 1) / n_sm;
        // measured on B200 (profiles/r2_chain_trace_b1.txt), in units of 12 MMAs ~ 1.2k clocks: first-patch latency ~1.5,
        // staged epilogue of an unsplit item ~7 (few CTAs format the whole output), register-to-partial epilogue of a
        // split item ~4, then barrier + in-place finish (every SM formats a slice) + barrier ~9
        const double cost = s == 1 ? (double)waves * (per + 8.5) : (double)waves * (per + 5.5) + 9.0 + 0.1 * s;
        if (cost < best_cost - 1e-9) { best_cost = cost; best_s = s; }
    }
    return best_s;
}

#ifdef B2T_DEV
static int choose_splits(int tiles, int chunks, int n_sm) {
    const char *env = getenv("B2T_SPLITS");
    if (env && atoi(env) > 0) {
        int s = atoi(env) < chunks ? atoi(env) : chunks;
        const int per = (chunks + s - 1) / s;
        return (chunks + per - 1) / per;
    }
    int best_s = 1;
    double best_cost = 1e30;
    const int max_s = chunks / 2 < 1 ? 1 : (chunks / 2 > 32 ? 32 : chunks / 2);
    for (int s = 1; s <= max_s; ++s) {
        const int per = (chunks + s - 1) / s;
        const int s_eff = (chunks + per - 1) / per;
        if (s_eff != s) continue;
        const long ctas = (long)tiles * s;
        const long waves = (ctas + n_sm - 1) / n_sm;
        const double cost = (double)waves * (per + 8.0) + (s > 1 ? 2.0 * s : 0.0);   // + split-K reduce traffic
        if (cost < best_cost - 1e-9) { best_cost = cost; best_s = s; }
    }
    return best_s;
}
#endif

// ------------------------------------------------------------------------------------------------ create
static ConvLayer &new_conv(b2t_ctx *c, int index, int k, int cin, int cout, bool act, bool pool, int H, int W) {
    ConvLayer l;
    l.index = index; l.k = k; l.cin = cin; l.cout = cout;
    l.kchunk = (cin <= 32 && index != 1 && c->cfg.engine != 2) ? 32 : 64;   // conv_2: <= 32 input channels, SWIZZLE_64B rows
    l.cin_pad = round_up(cin, l.kchunk);
    l.act = act; l.pool = pool; l.H = H; l.W = W;
    l.ldw = k * k * l.cin_pad;
    l.BN = cout >= 128 ? 128 : 64;
    choose_tile(H, W, pool, l.TW, l.TH);
    choose_halo_tile(H, W, k, pool, l);
    if (cout <= 64 && c->cfg.engine == B2T_ENGINE_TCGEN05 && index != 1) choose_pm_tile(H, W, k, pool, l.kchunk * 2, false, l);
    if (index == 1 && c->cfg.engine == B2T_ENGINE_TCGEN05) {
        // conv_1 (conv_pm.cu mode 1): an item is two image rows x pmC columns (columns on the MMA's M), patch = 4 rows
        const int nt = (W + 125) / 126;
        l.pmC = ((W + nt - 1) / nt + 1) & ~1;
        l.pmP = l.pmC + 2; l.pmR = 2; l.pm_rows = 4;
        l.pm_plane_bytes = (int)align_up((size_t)(3 * l.pmP + 128 + 8) * 16, 1024);   // last operand row: pixels 3P+127 .. +3
    }
    if ((int)c->conv.size() <= index) c->conv.resize(index + 1);
    c->conv[index] = l;
    return c->conv[index];
}

static void blob_reserve(b2t_ctx *c, ConvLayer &l) {
    size_t o = c->weight_bytes;
    l.off_whi = o;  o += align_up((size_t)l.cout * l.ldw * 2, 256);
    l.off_wlo = o;  o += align_up((size_t)l.cout * l.ldw * 2, 256);
    l.off_scale = o;  o += align_up((size_t)l.cout * 4, 256);
    l.off_bias = o;  o += align_up((size_t)l.cout * 4, 256);
    c->weight_bytes = o;
}

extern "C" int b2t_create(const b2t_config *cfg, b2t_ctx **out) {
    if (!cfg || !out) return fail(-1, "b2t_create: null argument");
    if (cfg->image_h % 32 || cfg->image_w % 32 || cfg->image_h <= 0 || cfg->image_w <= 0)
        return fail(-1, "b2t_create: image size %dx%d must be a positive multiple of 32", cfg->image_h, cfg->image_w);
    if (cfg->image_h != cfg->image_w) return fail(-1, "b2t_create: square input expected (GRID_H == GRID_W)");
    if (cfg->n_class < 1 || cfg->max_batch < 1) return fail(-1, "b2t_create: bad n_class/max_batch");
#ifndef B2T_DEV
    if (cfg->engine != B2T_ENGINE_TCGEN05)
        return fail(-1, "b2t_create: engine %d is a developer cross-check engine (build with make DEV=1)", cfg->engine);
#endif
    b2t_ctx *c = new b2t_ctx();
    c->cfg = *cfg;
    c->keep_prepool = cfg->reserved[0];
    if (cfg->reserved[1]) { c->chain_max_batch = cfg->reserved[1] < 0 ? 0 : cfg->reserved[1]; c->chain_forced = true; }
    c->G = cfg->image_h / 32;
    c->D = 5 + cfg->n_class;
    const int AD = c->A * c->D;
    const int H0 = cfg->image_h;
    c->conv.resize(24);

    const int Gs = c->G;
    const bool lstm = cfg->convlstm_units > 0;
    const bool tiny = cfg->reserved[2] == 1;
    // conv_1 blob: fp32 [27][32], scale, bias, LUT
    c->off_w1 = 0;
    c->off_s1 = align_up(27 * 32 * 4, 256);
    c->off_b1 = c->off_s1 + 256;
    c->off_lut = c->off_b1 + 256;
    c->off_w1pm = c->off_lut + 1024;
    c->off_s1pm = c->off_w1pm + 3 * 4096;
    c->weight_bytes = c->off_s1pm + 512;
    if (tiny) {
        // cfg/yolov2-tiny-voc.cfg / yolov2-tiny.cfg: conv 16, 32, 64, 128, 256 (each 3x3 + BN + leaky + maxpool 2/2), 512
        // (+ maxpool size 2 STRIDE 1), 1024, F8 (1024 voc / 512 coco), 1x1 head.  conv_1's kernel computes 32 channels:
        // the upper 16 carry zero weights, and conv_2 reads 32 input channels with zero weights on them.
        if (cfg->convlstm_units) { delete c; return fail(-1, "b2t_create: the tiny graph has no ConvLSTM head"); }
        const int f8 = cfg->reserved[3] > 0 ? cfg->reserved[3] : 1024;
        const int TT[8][3] = {{3, 32, 1}, {16, 32, 1}, {32, 64, 1}, {64, 128, 1}, {128, 256, 1}, {256, 512, 2}, {512, 1024, 0}, {1024, f8, 0}};
        c->n_conv = 9;
        c->conv1_cout = 16;
        c->conv.resize(10);
        int H = H0, prev = -1;
        for (int i = 1; i <= 8; ++i) {
            const int *t = TT[i - 1];
            char nm[32];
            snprintf(nm, sizeof nm, "norm_%d", i);
            ConvLayer &l = new_conv(c, i, 3, t[0], t[1], true, t[2] == 1, H, H);
            l.in_buf = prev;
            if (t[2] == 1) {
                if (c->keep_prepool) l.out_buf = add_buf(c, nm, H, H, round_up(l.cout, 64));
                l.pout_buf = add_buf(c, std::string("pool_") + std::to_string(i), H / 2, H / 2, i == 1 ? 32 : round_up(l.cout, 64));
                prev = l.pout_buf;
                H /= 2;
            } else {
                l.out_buf = add_buf(c, nm, H, H, round_up(l.cout, 64));
                prev = l.out_buf;
                if (t[2] == 2) { l.ps1_buf = add_buf(c, std::string("pool_") + std::to_string(i), H, H, round_up(l.cout, 64)); prev = l.ps1_buf; }
            }
            c->cout_by_name[nm] = i == 1 ? 16 : l.cout;
            c->cout_by_name[std::string("pool_") + std::to_string(i)] = i == 1 ? 16 : l.cout;
            if (i > 1) blob_reserve(c, l);
        }
        c->buf_by_name["conv_feat"] = c->conv[8].out_buf; c->cout_by_name["conv_feat"] = f8;
        ConvLayer &l = new_conv(c, 9, 1, f8, AD, false, false, Gs, Gs);
        l.in_buf = prev;
        l.f32_out = 1;
        blob_reserve(c, l);
    } else {
    // conv_k table (KerasYOLO.py:277-400): k, cin, cout, pool-after
    static const int T[20][4] = {{3, 3, 32, 1},     {3, 32, 64, 1},    {3, 64, 128, 0},    {1, 128, 64, 0},
                                 {3, 64, 128, 1},   {3, 128, 256, 0},  {1, 256, 128, 0},   {3, 128, 256, 1},
                                 {3, 256, 512, 0},  {1, 512, 256, 0},  {3, 256, 512, 0},   {1, 512, 256, 0},
                                 {3, 256, 512, 1},  {3, 512, 1024, 0}, {1, 1024, 512, 0},  {3, 512, 1024, 0},
                                 {1, 1024, 512, 0}, {3, 512, 1024, 0}, {3, 1024, 1024, 0}, {3, 1024, 1024, 0}};

    const int zc = lstm ? 1024 + round_up(AD, 64) : 1024;        // channels of the conv_22 output buffer
    const int concat = add_buf(c, "concat", Gs, Gs, 1280);
    const int featz = add_buf(c, "norm_22", Gs, Gs, zc);
    c->buf_z = featz;

    int H = H0, prev = -1;
    int skip_buf = -1;
    for (int i = 1; i <= 20; ++i) {
        const int *t = T[i - 1];
        char nm[32];
        snprintf(nm, sizeof nm, "norm_%d", i);
        ConvLayer &l = new_conv(c, i, t[0], t[1], t[2], true, t[3] != 0, H, H);
        l.in_buf = prev;
        if (i == 20) {                         // writes straight into the concat buffer (route 27,24)
            l.out_buf = concat; l.out_ch_off = 256;
            c->buf_by_name[nm] = concat; c->choff_by_name[nm] = 256;
        } else if (l.pool) {
            if (c->keep_prepool || i == 13) {
                l.out_buf = add_buf(c, nm, H, H, round_up(l.cout, 64));
                if (i == 13) skip_buf = l.out_buf;
            }
            l.pout_buf = add_buf(c, std::string("pool_") + std::to_string(i), H / 2, H / 2,
                                 (i == 1 && cfg->engine != 2) ? 32 : round_up(l.cout, 64));
        } else {
            l.out_buf = add_buf(c, nm, H, H, round_up(l.cout, 64));
        }
        c->cout_by_name[nm] = l.cout;
        if (i > 1) blob_reserve(c, l);
        prev = l.pool ? l.pout_buf : l.out_buf;
        if (l.pool) H /= 2;
    }
    {   // conv_21 on the 26x26 skip, space_to_depth into concat[0:256]
        ConvLayer &l = new_conv(c, 21, 1, 512, 64, true, false, 2 * Gs, 2 * Gs);
        l.in_buf = skip_buf;
        if (cfg->semantics == B2T_SEM_DARKNET) {
            // darknet's reorg is a permutation that mixes positions and channels: as a per-element scatter in the conv
            // epilogue it was 54 us per 36 frames; conv_21 now writes a plain tensor (coalesced) and a gather kernel with
            // coalesced stores permutes it into the concat buffer
            l.out_buf = add_buf(c, "reorg_src", 2 * Gs, 2 * Gs, 64);
            l.reorg_buf = concat;
        } else {
            l.out_buf = concat; l.out_ch_off = 0;
            l.out_mode = DEST_S2D_TF;
        }
        blob_reserve(c, l);
        c->cout_by_name["norm_21"] = 64;
    }
    {
        ConvLayer &l = new_conv(c, 22, 3, 1280, 1024, true, false, Gs, Gs);
        l.in_buf = concat;
        l.out_buf = featz;
        blob_reserve(c, l);
        c->cout_by_name["norm_22"] = 1024;
        c->buf_by_name["conv_feat"] = featz; c->cout_by_name["conv_feat"] = 1024;
        c->cout_by_name["concat"] = 1280;
    }
    {
        ConvLayer &l = new_conv(c, 23, 1, 1024, AD, false, false, Gs, Gs);
        l.in_buf = featz;
        l.f32_out = 1;
        if (lstm) { l.out_buf = featz; l.out_ch_off = 1024; }
        blob_reserve(c, l);
    }
    if (lstm) {
        const int u = cfg->convlstm_units;
        if (u % 64) { delete c; return fail(-1, "convlstm_units must be a multiple of 64"); }
        c->buf_hrec = add_buf(c, "", Gs, Gs, u);
        c->buf_hseq = add_buf(c, "", Gs, Gs, u);
        c->L_CIN = 24; c->L_CREC = 25; c->L_HEAD = 26;
        ConvLayer &a = new_conv(c, 24, 3, zc, 4 * u, false, false, Gs, Gs);
        a.in_buf = featz; a.f32_out = 2;
        blob_reserve(c, a);
        ConvLayer &r = new_conv(c, 25, 3, u, 4 * u, false, false, Gs, Gs);
        r.in_buf = c->buf_hrec; r.f32_out = 2; r.f32_accumulate = 1;
        blob_reserve(c, r);
        ConvLayer &h = new_conv(c, 26, 1, u, AD, false, false, Gs, Gs);
        h.in_buf = c->buf_hseq; h.f32_out = 3;
        blob_reserve(c, h);
    }
    }
    c->host_blob.assign(c->weight_bytes, 0);
    {   // LUT[u] = float(u / 255.)  (utils.py:150-153: numpy true division in float64, cast to fp32 by Keras)
        float *lut = reinterpret_cast<float *>(c->host_blob.data() + c->off_lut);
        for (int i = 0; i < 256; ++i) lut[i] = (float)((double)i / 255.0);
    }

    // ---- workspace layout
    size_t o = 0;
    const int MB = cfg->max_batch;
    for (size_t i = 0; i < c->bufs.size(); ++i) {
        ActBuf &b = c->bufs[i];
        b.off = o;
        o += align_up((size_t)b.plane * 2 * 2, 1024);
    }
    c->off_logits = o;  o += align_up((size_t)MB * Gs * Gs * AD * 4, 1024);
    c->off_c8 = o;  o += align_up((size_t)MB * H0 * H0 * 16, 1024);
    // split-K / SIMT partials: worst case over layers and batch sizes
    size_t pb = 0;
    for (size_t i = 2; i < c->conv.size(); ++i) {
        const ConvLayer &l = c->conv[i];
        if (!l.index) continue;
        const size_t ldp = round_up(l.cout, 32);
#ifdef B2T_DEV
        const int tiles1 = ((l.W + l.TW - 1) / l.TW) * ((l.H + l.TH - 1) / l.TH) * ((l.cout + l.BN - 1) / l.BN);
#endif
        for (int bsz = 1; bsz <= MB; ++bsz) {
            size_t s = 1;
#ifdef B2T_DEV
            if (cfg->engine == 2) {
                s = choose_splits(tiles1 * bsz, l.k * l.k * l.cin_pad / 64, 148);
                if (s == 1) continue;
            } else
#endif
            if (cfg->engine == B2T_ENGINE_TCGEN05) {
                const int ctas = ((l.W + l.hC - 1) / l.hC) * ((l.H + l.hR - 1) / l.hR) * ((l.cout + 127) / 128) * bsz;
                s = choose_splits_halo(ctas, l.cin_pad / l.kchunk, l.k * l.k, 148, bsz);
                if (bsz <= c->chain_max_batch) {
                    const size_t s2 = choose_splits_chain(ctas, l.cin_pad / l.kchunk * l.k * l.k, 148);
                    if (s2 > s) s = s2;
                }
                if (s == 1) continue;
            }
            const size_t need = s * (size_t)bsz * l.H * l.W * ldp * 4;
            if (need > pb) pb = need;
        }
    }
    c->off_partial = o;  c->partial_bytes = pb;  o += align_up(pb, 1024);
    if (lstm) {
        const int u = cfg->convlstm_units;
        c->off_gates = o;   o += align_up((size_t)MB * Gs * Gs * 4 * u * 4, 1024);
        c->off_cstate = o;  o += align_up((size_t)MB * Gs * Gs * u * 4, 1024);    // one (h, c) state slot per frame of a batch
    }
    c->ws_bytes = o;
    *out = c;
    return 0;
}

extern "C" void b2t_destroy(b2t_ctx *c) {
    if (!c) return;
    if (c->own_blob && c->d_blob) cudaFree(c->d_blob);
    if (c->own_ws && c->d_ws) cudaFree(c->d_ws);
    if (c->d_rs_tab) cudaFree(c->d_rs_tab);
    if (c->d_chain_counter) cudaFree(c->d_chain_counter);
    delete c;
}

extern "C" size_t b2t_weight_bytes(const b2t_ctx *c) { return c ? c->weight_bytes : 0; }
extern "C" size_t b2t_workspace_bytes(const b2t_ctx *c) { return c ? c->ws_bytes : 0; }
extern "C" long b2t_launch_count(const b2t_ctx *c) { return c ? c->launches : 0; }

extern "C" int b2t_bind_memory(b2t_ctx *c, void *blob, void *ws) {
    if (!c) return fail(-1, "null ctx");
    if (c->finalized) return fail(-1, "b2t_bind_memory after b2t_finalize");
    if (((uintptr_t)blob & 1023) || ((uintptr_t)ws & 1023)) return fail(-1, "device pointers must be 1024-byte aligned");
    c->d_blob = (uint8_t *)blob;
    c->d_ws = (uint8_t *)ws;
    return 0;
}

// ------------------------------------------------------------------------------------------------ weights
// kernel_hwio (k,k,cin_src,cout) -> [cout][tap][cin_pad] hi/lo; perm[c_dst] = c_src or -1 (NULL = identity).
// Every output channel is pre-scaled by a power of two so that its largest |w| lands in [128, 256): small
// weights (|w| ~ 1e-2 is typical) would otherwise push the lo halves into fp16's subnormal range and lose
// the pair's 22 bits.  unscale[co] = 2^-shift is folded (exactly) into the epilogue's per-channel scale.
static void pack_conv(b2t_ctx *c, const ConvLayer &l, const float *ker, int cin_src, const int *perm,
                      std::vector<float> &unscale) {
    op_t *hi = reinterpret_cast<op_t *>(c->host_blob.data() + l.off_whi);
    op_t *lo = reinterpret_cast<op_t *>(c->host_blob.data() + l.off_wlo);
    const int taps = l.k * l.k;
    unscale.assign(l.cout, 1.f);
    for (int co = 0; co < l.cout; ++co) {
        float m = 0.f;
        for (int t = 0; t < taps; ++t)
            for (int ci = 0; ci < cin_src; ++ci) m = fmaxf(m, fabsf(ker[((size_t)t * cin_src + ci) * l.cout + co]));
        int shift = 0;
        if (m > 0.f && std::isfinite(m)) {
            int e;
            frexpf(m, &e);               // m = f * 2^e, f in [0.5, 1)
            shift = 8 - e;               // m * 2^shift in [128, 256)
            if (shift > 40) shift = 40;
            if (shift < -8) shift = -8;
        }
        const float up = ldexpf(1.f, shift);
        unscale[co] = ldexpf(1.f, -shift);
        for (int t = 0; t < taps; ++t)
            for (int ci = 0; ci < l.cin_pad; ++ci) {
                const int src = perm ? perm[ci] : (ci < cin_src ? ci : -1);
                float w = 0.f;
                if (src >= 0) w = ker[((size_t)t * cin_src + src) * l.cout + co] * up;
                const size_t d = (size_t)co * l.ldw + (size_t)t * l.cin_pad + ci;
                split_f16(w, hi[d], lo[d]);
            }
    }
}

static void fold_bn(const b2t_ctx *c, int n, const float *gamma, const float *beta, const float *mean, const float *var,
                    float *scale, float *bias) {
    for (int i = 0; i < n; ++i) {
        double s;
        if (c->cfg.semantics == B2T_SEM_DARKNET)
            s = (double)gamma[i] / (sqrt((double)var[i]) + 1e-6);                 // blas.c:156 normalize_cpu
        else
            s = (double)gamma[i] / sqrt((double)var[i] + (double)c->cfg.bn_eps);  // keras BN inference
        scale[i] = (float)s;
        bias[i] = (float)((double)beta[i] - (double)mean[i] * s);
    }
}

extern "C" int b2t_set_conv_weights(b2t_ctx *c, int idx, const float *ker, const float *gamma, const float *beta,
                                    const float *mean, const float *var, const float *bias) {
    if (!c || idx < 1 || idx > c->n_conv || !ker) return fail(-1, "b2t_set_conv_weights: bad arguments (conv_%d)", idx);
    ConvLayer &l = c->conv[idx];
    const bool bn = idx != c->n_conv;
    if (bn && !(gamma && beta && mean && var)) return fail(-1, "conv_%d needs gamma/beta/mean/var", idx);
    if (!bn && !bias) return fail(-1, "the head conv needs a bias");
    std::vector<float> k1, g1, b1v, m1, v1;
    if (idx == 1 && c->conv1_cout != 32) {
        // conv_1 with fewer than 32 real output channels: pad with zero weights and an identity BatchNorm (output 0)
        const int rc1 = c->conv1_cout;
        k1.assign(27 * 32, 0.f); g1.assign(32, 1.f); b1v.assign(32, 0.f); m1.assign(32, 0.f); v1.assign(32, 1.f);
        for (int t = 0; t < 27; ++t)
            for (int co = 0; co < rc1; ++co) k1[t * 32 + co] = ker[t * rc1 + co];
        for (int co = 0; co < rc1; ++co) { g1[co] = gamma[co]; b1v[co] = beta[co]; m1[co] = mean[co]; v1[co] = var[co]; }
        ker = k1.data(); gamma = g1.data(); beta = b1v.data(); mean = m1.data(); var = v1.data();
    }
    if (idx == 1) {
        float *w = reinterpret_cast<float *>(c->host_blob.data() + c->off_w1);
        memcpy(w, ker, 27 * 32 * 4);   // (kh,kw,cin,cout) is already [tap*3+cin][32]
        float *s1 = reinterpret_cast<float *>(c->host_blob.data() + c->off_s1);
        fold_bn(c, 32, gamma, beta, mean, var, s1, reinterpret_cast<float *>(c->host_blob.data() + c->off_b1));
        // tensor-core path (conv_pm.cu mode 1): per kernel row kh a [32 cout][32 k] fp16 tile, k = kw*8 + cin, stored
        // as 128-byte core matrices [cout/8][k/8][cout%8][k%8]; the three tiles of the hi plane in DESCENDING kernel-row
        // order (kh = 2, 1, 0), then the lo plane: two consecutive tiles [W_kh ; W_kh-1] form the 64-row operand that
        // serves both output rows of a row pair from one patch row; 1/255 goes into the scale
        op_t *wp = reinterpret_cast<op_t *>(c->host_blob.data() + c->off_w1pm);
        float *s1pm = reinterpret_cast<float *>(c->host_blob.data() + c->off_s1pm);
        memset(wp, 0, 3 * 4096);
        for (int co = 0; co < 32; ++co) {
            float m = 0.f;
            for (int t = 0; t < 27; ++t) m = fmaxf(m, fabsf(ker[t * 32 + co]));
            int shift = 0;
            if (m > 0.f && std::isfinite(m)) {
                int e;
                frexpf(m, &e);
                shift = 8 - e;
                if (shift > 40) shift = 40;
                if (shift < -8) shift = -8;
            }
            // channels with a negative folded-BN scale get negated weights and |scale|: BN + LeakyReLU are then
            // increasing in the raw sum for every channel, which lets the kernel max-pool before applying them
            const float up = s1[co] < 0.f ? -ldexpf(1.f, shift) : ldexpf(1.f, shift);
            for (int kh = 0; kh < 3; ++kh)
                for (int kw = 0; kw < 3; ++kw)
                    for (int ci = 0; ci < 3; ++ci) {
                        const size_t d = (size_t)(2 - kh) * 1024 + (size_t)((co / 8) * 4 + kw) * 64 + (co % 8) * 8 + ci;   // fp16 elements
                        split_f16(ker[((kh * 3 + kw) * 3 + ci) * 32 + co] * up, wp[d], wp[d + 3072]);
                    }
            s1pm[co] = (float)(fabs((double)s1[co]) * (double)ldexpf(1.f, -shift) / 255.0);
        }
    } else {
        std::vector<float> un;
        pack_conv(c, l, ker, l.cin, nullptr, un);
        float *s = reinterpret_cast<float *>(c->host_blob.data() + l.off_scale);
        float *b = reinterpret_cast<float *>(c->host_blob.data() + l.off_bias);
        if (bn) {
            fold_bn(c, l.cout, gamma, beta, mean, var, s, b);
        } else {
            for (int i = 0; i < l.cout; ++i) { s[i] = 1.f; b[i] = bias[i]; }
        }
        for (int i = 0; i < l.cout; ++i) s[i] *= un[i];          // exact: power of two
    }
    l.have_weights = true;
    return 0;
}

extern "C" int b2t_load_darknet_weights(b2t_ctx *c, const char *path) {
    if (!c || !path) return fail(-1, "null argument");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(-3, "cannot open %s", path);
    int32_t hdr[3];
    if (fread(hdr, 4, 3, f) != 3) { fclose(f); return fail(-3, "%s: truncated header", path); }
    // parser.c:1220-1226: "seen" is size_t from v0.2 on, int32 before
    const bool wide = (hdr[0] * 10 + hdr[1] >= 2) && hdr[0] < 1000 && hdr[1] < 1000;
    if (fseek(f, wide ? 8 : 4, SEEK_CUR)) { fclose(f); return fail(-3, "%s: truncated header", path); }
    std::vector<float> bn, raw, hwio;
    for (int idx = 1; idx <= c->n_conv; ++idx) {          // file order = cfg order (conv_21's skip conv precedes conv_22)
        ConvLayer l = c->conv[idx];
        if (idx == 1) l.cout = c->conv1_cout;             // the file holds the real channel count
        const bool isbn = idx != c->n_conv;
        const size_t nb = isbn ? 4 * (size_t)l.cout : (size_t)l.cout, nw = (size_t)l.cout * l.cin * l.k * l.k;
        bn.resize(nb);
        raw.resize(nw);
        hwio.resize(nw);
        if (fread(bn.data(), 4, nb, f) != nb || fread(raw.data(), 4, nw, f) != nw) {
            fclose(f);
            return fail(-3, "%s: truncated at conv_%d (wrong class count?)", path, idx);
        }
        // [Cout][Cin][kh][kw] -> (kh,kw,Cin,Cout)   (KerasYOLO.py:270-272)
        const int kk = l.k * l.k;
        for (int co = 0; co < l.cout; ++co)
            for (int ci = 0; ci < l.cin; ++ci)
                for (int t = 0; t < kk; ++t) hwio[((size_t)t * l.cin + ci) * l.cout + co] = raw[((size_t)co * l.cin + ci) * kk + t];
        int rc;
        if (isbn)   // file order: biases(beta), scales(gamma), rolling_mean, rolling_variance (parser.c:1165-1171)
            rc = b2t_set_conv_weights(c, idx, hwio.data(), bn.data() + l.cout, bn.data(), bn.data() + 2 * l.cout,
                                      bn.data() + 3 * l.cout, nullptr);
        else
            rc = b2t_set_conv_weights(c, idx, hwio.data(), nullptr, nullptr, nullptr, nullptr, bn.data());
        if (rc) { fclose(f); return rc; }
    }
    float extra;
    const bool trailing = fread(&extra, 4, 1, f) == 1;
    fclose(f);
    if (trailing) return fail(-3, "%s: trailing data after the last conv layer (wrong class count?)", path);
    return 0;
}

extern "C" int b2t_set_convlstm_weights(b2t_ctx *c, const float *kernel, const float *recurrent, const float *bias,
                                        const float *head_kernel, const float *head_bias) {
    if (!c || !c->L_CIN) return fail(-1, "context was created without convlstm_units");
    if (!kernel || !recurrent || !bias || !head_kernel || !head_bias) return fail(-1, "null weight pointer");
    const int AD = c->A * c->D, u = c->cfg.convlstm_units;
    ConvLayer &a = c->conv[c->L_CIN], &r = c->conv[c->L_CREC], &h = c->conv[c->L_HEAD];
    // keras channel order of z: [x_bbox (A*D), x_vis (1024)]  (MultiObjDetTracker.py:175); ours: [feat | logits | pad]
    std::vector<int> perm(a.cin_pad, -1);
    for (int i = 0; i < 1024; ++i) perm[i] = AD + i;
    for (int i = 0; i < AD; ++i) perm[1024 + i] = i;
    std::vector<float> ua, ur, uh;
    pack_conv(c, a, kernel, AD + 1024, perm.data(), ua);
    pack_conv(c, r, recurrent, u, nullptr, ur);
    pack_conv(c, h, head_kernel, u, nullptr, uh);
    auto fill = [&](ConvLayer &l, const float *b, const std::vector<float> &un) {
        float *s = reinterpret_cast<float *>(c->host_blob.data() + l.off_scale);
        float *bb = reinterpret_cast<float *>(c->host_blob.data() + l.off_bias);
        for (int i = 0; i < l.cout; ++i) { s[i] = un[i]; bb[i] = b ? b[i] : 0.f; }
        l.have_weights = true;
    };
    fill(a, bias, ua);
    fill(r, nullptr, ur);
    fill(h, head_bias, uh);
    return 0;
}

// ------------------------------------------------------------------------------------------------ finalize
static int make_tmap(b2t_ctx *c, CUtensorMap *tm, void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
                     const cuuint32_t *box, int swz = 0 /* 0 = 128B, 1 = 64B, 2 = none */) {
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = c->encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, base, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swz == 2 ? CU_TENSOR_MAP_SWIZZLE_NONE : swz == 1 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-2, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

extern "C" int b2t_finalize(b2t_ctx *c, int upload, void *stream) {
    if (!c) return fail(-1, "null ctx");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->cfg.device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, c->cfg.device));
    if (prop.major != 10) return fail(-2, "device %d is sm_%d%d; this library contains sm_100a code only", c->cfg.device,
                                      prop.major, prop.minor);
    c->n_sm = prop.multiProcessorCount;
    if (upload)
        for (size_t i = 1; i < c->conv.size(); ++i)
            if (c->conv[i].index && !c->conv[i].have_weights) return fail(-1, "conv layer %zu has no weights", i);
    if (!c->d_blob) { CK(cudaMalloc(&c->d_blob, c->weight_bytes)); c->own_blob = true; }
    if (!c->d_ws) { CK(cudaMalloc(&c->d_ws, c->ws_bytes)); c->own_ws = true; }
    if (!c->d_chain_counter) CK(cudaMalloc(&c->d_chain_counter, 256));
    if (!c->encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) return fail(-2, "cuTensorMapEncodeTiled not available in this driver");
        c->encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    }
    int rc = conv_halo_init();
#ifdef B2T_DEV
    if (!rc) rc = conv_umma_init();
#endif
    if (!rc) rc = conv_pm_init();
    if (rc) return fail(-2, "cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString((cudaError_t)rc));
    if (upload) CK(cudaMemcpyAsync(c->d_blob, c->host_blob.data(), c->weight_bytes, cudaMemcpyHostToDevice, st));
    // pad channels / never-written halo must be finite zeros
    CK(cudaMemsetAsync(c->d_ws, 0, c->ws_bytes, st));
    for (auto &b : c->bufs) b.hi = reinterpret_cast<op_t *>(c->d_ws + b.off);
    const int MB = c->cfg.max_batch;
    for (size_t i = 2; i < c->conv.size(); ++i) {
        ConvLayer &l = c->conv[i];
        if (!l.index) continue;
        const ActBuf &in = c->bufs[l.in_buf];
        const int nb = MB;
        cuuint64_t dims[4] = {(cuuint64_t)l.cin_pad, (cuuint64_t)l.W, (cuuint64_t)l.H, (cuuint64_t)nb};
        cuuint64_t strides[3] = {(cuuint64_t)in.C * 2, (cuuint64_t)in.C * 2 * l.W, (cuuint64_t)in.C * 2 * l.W * l.H};
        cuuint32_t box[4] = {64, (cuuint32_t)l.TW, (cuuint32_t)l.TH, 1};
        if (l.kchunk == 64) {
            if ((rc = make_tmap(c, &l.tmA_hi, in.hi + l.in_ch_off, 4, dims, strides, box))) return rc;
            if ((rc = make_tmap(c, &l.tmA_lo, in.hi + in.plane + l.in_ch_off, 4, dims, strides, box))) return rc;
        }
        cuuint64_t wd[2] = {(cuuint64_t)l.ldw, (cuuint64_t)l.cout};
        cuuint64_t wst[1] = {(cuuint64_t)l.ldw * 2};
        cuuint32_t wbox[2] = {64, (cuuint32_t)l.BN};
        if (l.kchunk == 64) {
            if ((rc = make_tmap(c, &l.tmB_hi, c->d_blob + l.off_whi, 2, wd, wst, wbox))) return rc;
            if ((rc = make_tmap(c, &l.tmB_lo, c->d_blob + l.off_wlo, 2, wd, wst, wbox))) return rc;
        }
        const bool sw64 = l.kchunk == 32;
        cuuint32_t xbox[4] = {(cuuint32_t)l.kchunk, (cuuint32_t)l.hP, (cuuint32_t)l.h_rows, 1};
        if ((rc = make_tmap(c, &l.tmX_hi, in.hi + l.in_ch_off, 4, dims, strides, xbox, sw64))) return rc;
        if ((rc = make_tmap(c, &l.tmX_lo, in.hi + in.plane + l.in_ch_off, 4, dims, strides, xbox, sw64))) return rc;
        cuuint32_t wbox2[2] = {(cuuint32_t)l.kchunk, 128};
        if ((rc = make_tmap(c, &l.tmW_hi, c->d_blob + l.off_whi, 2, wd, wst, wbox2, sw64))) return rc;
        if ((rc = make_tmap(c, &l.tmW_lo, c->d_blob + l.off_wlo, 2, wd, wst, wbox2, sw64))) return rc;
        l.w_rows = l.cout <= 64 ? 64 : 128;
        cuuint32_t wbox3[2] = {(cuuint32_t)l.kchunk, (cuuint32_t)l.w_rows};
        if ((rc = make_tmap(c, &l.tmWp_hi, c->d_blob + l.off_whi, 2, wd, wst, wbox3, sw64))) return rc;
        if ((rc = make_tmap(c, &l.tmWp_lo, c->d_blob + l.off_wlo, 2, wd, wst, wbox3, sw64))) return rc;
        // 1x1 layers without pooling and with a plain destination: tiles of exactly 128 pixels of the flat pixel list
        // of the whole batch, persistent kernel (TMEM double buffering: the epilogue overlaps the next tile's MMAs)
        static const int flat_mode = dev_env("B2T_FLAT", 1);
        // (measured: a win up to K = 512; longer K re-streams too many weight bytes per 128-pixel tile -- the
        // shared-memory port saturates -- and the N = 192 whole-image tiles of conv_halo_kernel stay faster)
        // (cout > 64: the persistent kernel's weight box is w_rows = 64 rows for narrower layers, its MMA reads 128)
        l.flat1x1 = flat_mode && l.k == 1 && !l.pool && l.out_mode == DEST_PLAIN && l.kchunk == 64 && nb == MB && l.cout > 64 &&
                    l.cin_pad <= 512 && c->cfg.engine == B2T_ENGINE_TCGEN05;
        if (l.flat1x1) {
            const cuuint64_t npx = (cuuint64_t)MB * l.H * l.W;
            cuuint64_t fd[4] = {(cuuint64_t)l.cin_pad, npx, 1, 1};
            cuuint64_t fs[3] = {(cuuint64_t)in.C * 2, (cuuint64_t)in.C * 2 * npx, (cuuint64_t)in.C * 2 * npx};
            cuuint32_t fb[4] = {64, 128, 1, 1};
            if ((rc = make_tmap(c, &l.tmXf_hi, in.hi + l.in_ch_off, 4, fd, fs, fb))) return rc;
            if ((rc = make_tmap(c, &l.tmXf_lo, in.hi + in.plane + l.in_ch_off, 4, fd, fs, fb))) return rc;
        }
        // pixel-major variant: narrow layers whose weights all stay resident in shared memory
        static const int pm_mode = dev_env("B2T_PM", 1);
        l.pm = false;
        if (pm_mode && l.pmC && c->cfg.engine == B2T_ENGINE_TCGEN05 && l.cout <= 64 && nb == MB) {
            l.pm_flat = l.k == 1 && !l.pool && l.out_mode == DEST_PLAIN;
            const int kbytes = l.kchunk * 2;
            const int plane = l.pm_flat ? 128 * kbytes : l.pm_plane_bytes;
            const int ld = round_up(l.cout, 16) + 4;
            const size_t smem = 1024 + 2048 + 2 * 2 * (size_t)plane + 2 * align_up((size_t)128 * ld * 4, 1024) +
                                (size_t)l.k * l.k * (l.cin_pad / l.kchunk) * 2 * l.w_rows * kbytes;
            if (smem <= 226 * 1024) {
                l.pm = true;
                if (l.pm_flat) {
                    const cuuint64_t npx = (cuuint64_t)MB * l.H * l.W;
                    cuuint64_t fd[4] = {(cuuint64_t)l.cin_pad, npx, 1, 1};
                    cuuint64_t fs[3] = {(cuuint64_t)in.C * 2, (cuuint64_t)in.C * 2 * npx, (cuuint64_t)in.C * 2 * npx};
                    cuuint32_t fb[4] = {(cuuint32_t)l.kchunk, 128, 1, 1};
                    if ((rc = make_tmap(c, &l.tmXp_hi, in.hi + l.in_ch_off, 4, fd, fs, fb, sw64))) return rc;
                    if ((rc = make_tmap(c, &l.tmXp_lo, in.hi + in.plane + l.in_ch_off, 4, fd, fs, fb, sw64))) return rc;
                } else {
                    cuuint32_t pbox[4] = {(cuuint32_t)l.kchunk, (cuuint32_t)l.pmP, (cuuint32_t)l.pm_rows, 1};
                    if ((rc = make_tmap(c, &l.tmXp_hi, in.hi + l.in_ch_off, 4, dims, strides, pbox, sw64))) return rc;
                    if ((rc = make_tmap(c, &l.tmXp_lo, in.hi + in.plane + l.in_ch_off, 4, dims, strides, pbox, sw64))) return rc;
                }
            }
        }
    }
    {   // conv_1 on the tensor cores: TMA view of the fp16-integer frame copy [MB][H][W][8]
        ConvLayer &l = c->conv[1];
        static const int pm_mode = dev_env("B2T_PM", 1);
        l.pm = false;
        if (pm_mode && l.pmC && c->cfg.engine == B2T_ENGINE_TCGEN05) {
            cuuint64_t dims[4] = {8, (cuuint64_t)l.W, (cuuint64_t)l.H, (cuuint64_t)MB};
            cuuint64_t strides[3] = {16, (cuuint64_t)16 * l.W, (cuuint64_t)16 * l.W * l.H};
            cuuint32_t box[4] = {8, (cuuint32_t)l.pmP, (cuuint32_t)l.pm_rows, 1};
            if ((rc = make_tmap(c, &l.tmXp_hi, c->d_ws + c->off_c8, 4, dims, strides, box, 2))) return rc;
            l.tmXp_lo = l.tmXp_hi;
            l.pm = true;
        }
    }
    c->finalized = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ forward
static Dest dest_planes(const b2t_ctx *c, int buf, int ch_off, int srcH, int srcW, int mode) {
    Dest d;
    memset(&d, 0, sizeof d);
    if (buf >= 0) {
        const ActBuf &b = c->bufs[buf];
        d.hi = b.hi;
        d.plane_stride = b.plane;
        d.pix_stride_b = b.C;
        d.ch_off_b = ch_off;
    }
    d.H = srcH; d.W = srcW; d.mode = mode;
    return d;
}

// f32_img_stride (pixels, 0 = dense) and in_img_off (images) serve the ConvLSTM recurrent conv: stream s of a step reads
// state slot in_img_off + s and accumulates into frame s*T + t of the gate buffer.
// layer-independent part of a conv launch's parameter block
static void base_params(b2t_ctx *c, ConvLayer &l, int B, float *f32_dst, long long f32_img_stride, int in_img_off,
                        ConvParams &p) {
    memset(&p, 0, sizeof p);
    p.B = B; p.H = l.H; p.W = l.W; p.ksize = l.k; p.cin_chunks = l.cin_pad / l.kchunk; p.Cout = l.cout;
    p.kbytes = l.kchunk * 2;
    p.TW = l.TW; p.TH = l.TH;
    p.tiles_x = (l.W + l.TW - 1) / l.TW; p.tiles_y = (l.H + l.TH - 1) / l.TH;
    p.chunks_total = l.k * l.k * p.cin_chunks;
    p.ldp = round_up(l.cout, 32);
    p.act = l.act; p.pool = l.pool;
    { static const int nm = dev_env("B2T_NMAIN", 3); p.n_main = nm < 1 ? 1 : nm > 3 ? 3 : nm; }
    p.scale = reinterpret_cast<const float *>(c->d_blob + l.off_scale);
    p.bias = reinterpret_cast<const float *>(c->d_blob + l.off_bias);
    p.partial = reinterpret_cast<float *>(c->d_ws + c->off_partial);
    p.out = dest_planes(c, l.out_buf, l.out_ch_off, l.H, l.W, l.out_mode);
    p.pout = dest_planes(c, l.pout_buf, 0, l.H / 2, l.W / 2, DEST_PLAIN);
    if (f32_dst) {
        p.out.f32 = f32_dst;
        p.out.pix_stride_f = l.cout;
        p.out.ch_off_f = 0;
        p.out.accumulate_f = l.f32_accumulate;
        p.out.img_stride_f = f32_img_stride;
    }
    p.b_in_off = in_img_off;
}

// tile geometry + K split of the halo engine (conv_halo_kernel / conv_chain_kernel) for `B` frames
static int halo_geometry(b2t_ctx *c, ConvLayer &l, int B, ConvParams &p, bool chain = false) {
    p.hC = l.hC; p.hP = l.hP; p.hR = l.hR; p.hN = l.hN; p.h_rows = l.h_rows; p.h_plane_bytes = l.h_plane_bytes;
    p.h_tiles_x = (l.W + l.hC - 1) / l.hC; p.h_tiles_y = (l.H + l.hR - 1) / l.hR;
    const int ctas = B * p.h_tiles_x * p.h_tiles_y * ((l.cout + 127) / 128);
    if (chain) {
        const int units = p.cin_chunks * l.k * l.k;
        p.splits = choose_splits_chain(ctas, units, c->n_sm);
        p.k_per_units = (units + p.splits - 1) / p.splits;
    } else if (!l.h_small && l.hN <= 192) {
        // long accumulation chains are cut into passes INSIDE the CTA (conv_halo_kernel<big>: the running sums are
        // parked in spare TMEM columns / registers), so the K split is chosen for parallelism only
        static const int chain_cap = dev_env("B2T_CHAIN", 320);
        p.splits = choose_splits_halo(ctas, p.cin_chunks, l.k * l.k, c->n_sm, B, false);
        const int per = (p.cin_chunks + p.splits - 1) / p.splits;
        p.k_passes = (per * l.k * l.k * 4 + chain_cap - 1) / chain_cap;
    } else {
        p.splits = choose_splits_halo(ctas, p.cin_chunks, l.k * l.k, c->n_sm, B);
    }
    if (p.splits > 1 && (size_t)p.splits * B * l.H * l.W * p.ldp * 4 > c->partial_bytes)
        return fail(-2, "internal: split-K workspace too small for conv %d", l.index);
    return ctas;
}

// Small batches: can this layer run inside conv_chain_kernel (the big resource shape of the halo engine)?
static bool chain_eligible(const b2t_ctx *c, const ConvLayer &l) {
    return c->cfg.engine == B2T_ENGINE_TCGEN05 && l.index >= 2 && l.index <= c->n_conv && l.h_rows * l.hP <= 256 && l.hN <= 256 &&
           2 * l.h_plane_bytes <= 65536;
}

static int run_conv(b2t_ctx *c, ConvLayer &l, int B, float *f32_dst, cudaStream_t st, long long f32_img_stride = 0,
                    int in_img_off = 0) {
    ConvParams p;
    base_params(c, l, B, f32_dst, f32_img_stride, in_img_off, p);
    int rc;
#ifdef B2T_DEV
    if (c->cfg.engine == B2T_ENGINE_SIMT) {
        p.splits = 1;
        const ActBuf &in = c->bufs[l.in_buf];
        SimtView v;
        v.a_hi = in.hi + l.in_ch_off + (long long)in_img_off * l.H * l.W * in.C; v.a_plane = in.plane; v.a_pix_stride = in.C;
        v.w_hi = reinterpret_cast<const op_t *>(c->d_blob + l.off_whi);
        v.w_plane = (long long)(l.off_wlo - l.off_whi) / 2; v.w_ld = l.ldw;
        if ((rc = launch_conv_simt(v, p, st))) return fail(-2, "conv_simt launch: %s", cudaGetErrorString((cudaError_t)rc));
        if ((rc = launch_splitk_epilogue(p, st))) return fail(-2, "epilogue launch: %s", cudaGetErrorString((cudaError_t)rc));
        c->launches += 2;
        return 0;
    }
    { static const int dbg = dev_env("B2T_DBG", 0); p.dbg = dbg; }
#endif
    if (c->cfg.engine == B2T_ENGINE_TCGEN05 && l.pm && !f32_dst) {
        p.splits = 1;
        p.pm_mode = 0;
        p.pm_n = round_up(l.cout, 32);
        p.pm_tmem_cols = 4 * p.pm_n <= 128 ? 128 : 4 * p.pm_n <= 256 ? 256 : 512;
        p.pm_stage_ld = p.pm_n + 4;
        p.pm_glog = p.pm_n <= 32 ? 2 : p.pm_n <= 64 ? 3 : 4;
        p.pw_stage_bytes = (int)align_up((size_t)128 * p.pm_stage_ld * 4, 1024);
        p.pw_tile_bytes = 2 * l.w_rows * p.kbytes;
        if (l.pm_flat) {
            p.B = 1; p.H = 1; p.W = B * l.H * l.W;
            p.hC = 128; p.hP = 128; p.hR = 1; p.h_rows = 1;
            p.h_plane_bytes = 128 * p.kbytes;
        } else {
            p.hC = l.pmC; p.hP = l.pmP; p.hR = l.pmR; p.h_rows = l.pm_rows;
            p.h_plane_bytes = l.pm_plane_bytes;
        }
        p.hN = 128;
        p.pw_patch_bytes = 2 * p.h_plane_bytes;
        p.h_tiles_x = (p.W + p.hC - 1) / p.hC; p.h_tiles_y = (p.H + p.hR - 1) / p.hR;
        p.magic_tx = magic_for(p.h_tiles_x); p.magic_ty = magic_for(p.h_tiles_y);
        if ((rc = launch_conv_pm(c->n_sm, l.tmXp_hi, l.tmXp_lo, l.tmWp_hi, l.tmWp_lo, p, st)))
            return fail(-2, "conv_pm launch (conv %d): %s", l.index, cudaGetErrorString((cudaError_t)rc));
        c->launches += 1;
        return 0;
    }
    if (c->cfg.engine == B2T_ENGINE_TCGEN05 && l.flat1x1 && !in_img_off) {
        const int npx = B * l.H * l.W;
        const int items = ((npx + 127) / 128) * ((l.cout + 127) / 128);
        if (items >= c->n_sm) {                 // enough tiles to fill the machine without a K split
            p.splits = 1;
            p.B = 1; p.H = 1; p.W = npx;
            p.hC = 128; p.hP = 128; p.hR = 1; p.hN = 128; p.h_rows = 1; p.h_plane_bytes = 128 * p.kbytes;
            p.h_tiles_x = (npx + 127) / 128; p.h_tiles_y = 1;
            p.pw_patch_bytes = 2 * p.h_plane_bytes;
            p.pw_stage_bytes = (int)align_up((size_t)128 * 132 * 4, 1024);
            p.pw_tile_bytes = 2 * l.w_rows * p.kbytes;
            const int room = 222 * 1024 - 1024 - 1024 - 2 * p.pw_patch_bytes - p.pw_stage_bytes;
            p.pw_stages = room / p.pw_tile_bytes > 16 ? 16 : room / p.pw_tile_bytes;
            if ((rc = launch_conv_halo_persist(c->n_sm, l.tmXf_hi, l.tmXf_lo, l.tmWp_hi, l.tmWp_lo, p, st)))
                return fail(-2, "conv_halo_persist launch (conv %d): %s", l.index, cudaGetErrorString((cudaError_t)rc));
            c->launches += 1;
            return 0;
        }
    }
    if (c->cfg.engine == B2T_ENGINE_TCGEN05) {
        const int ctas = halo_geometry(c, l, B, p);
        if (ctas < 0) return ctas;
        static const int persist_mode = dev_env("B2T_PERSIST", 2);
        // persistent variant (one CTA per SM, TMEM double buffering) for every short-K layer with enough tiles;
        // B2T_PERSIST=1 restricts it to 1x1 layers and layers whose weights stay resident, 0 disables it
        bool persist = persist_mode && l.h_small && p.splits == 1 && p.hN <= 128 && ctas > 2 * c->n_sm &&
                       (l.k == 1 || (l.cout <= 128 && l.k * l.k * p.cin_chunks * 2 * l.w_rows * p.kbytes <= 96 * 1024) ||
                        persist_mode > 1);
        if (persist) {
            p.pw_patch_bytes = 2 * l.h_plane_bytes;
            p.pw_stage_bytes = (int)align_up((size_t)p.hN * 132 * 4, 1024);
            p.pw_tile_bytes = 2 * l.w_rows * p.kbytes;
            // leave a few KB of the 228 KB SM array to L1
            const int room = 222 * 1024 - 1024 /*align*/ - 1024 /*barriers*/ - 2 * p.pw_patch_bytes - p.pw_stage_bytes;
            p.pw_stages = room / p.pw_tile_bytes > 16 ? 16 : room / p.pw_tile_bytes;
            if (p.pw_stages < 2) persist = false;
        }
#ifdef B2T_DEV
        static long long *d_trace = nullptr;
        const char *tr = getenv("B2T_TRACE_CONV");
        if (tr && atoi(tr) == l.index) {
            if (!d_trace) cudaMalloc(&d_trace, 64 * 8 * 8);
            cudaMemsetAsync(d_trace, 0, 64 * 8 * 8, st);
            p.trace = d_trace;
        }
#endif
        if (persist)
            rc = launch_conv_halo_persist(c->n_sm, l.tmX_hi, l.tmX_lo, l.tmWp_hi, l.tmWp_lo, p, st);
        else
            rc = launch_conv_halo(c->n_sm, l.h_small, l.tmX_hi, l.tmX_lo, l.tmW_hi, l.tmW_lo, p, st);
#ifdef B2T_DEV
        if (p.trace) {
            long long h[64 * 8];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, d_trace, sizeof h, cudaMemcpyDeviceToHost);
            const long long t0 = h[1];
            if (persist) {
                fprintf(stderr, "[trace conv_%d] item: patch_issue | mma: wait_acc_begin acc_ok patch_ok | epi: wait_begin acc_full done (cycles rel.)\n", l.index);
                for (int j = 0; j < 12; ++j)
                    fprintf(stderr, "  %2d: %7lld | %7lld %7lld %7lld | %7lld %7lld %7lld\n", j, h[j * 8] - t0, h[j * 8 + 1] - t0,
                            h[j * 8 + 2] - t0, h[j * 8 + 3] - t0, h[j * 8 + 4] - t0, h[j * 8 + 5] - t0, h[j * 8 + 6] - t0);
            } else {
                fprintf(stderr, "[trace conv_%d N=%d splits=%d] item: mma: start acc_ok patch_ok issue_end | epi: wait accum_ok done ; last phase1 %lld\n", l.index, p.hN, p.splits, h[511] - t0);
                for (int j = 0; j < 4; ++j)
                    fprintf(stderr, "  %2d: %7lld %7lld %7lld %7lld | %7lld %7lld %7lld\n", j, h[j * 8 + 1] - t0, h[j * 8 + 2] - t0,
                            h[j * 8 + 3] - t0, h[j * 8 + 7] - t0, h[j * 8 + 4] - t0, h[j * 8 + 5] - t0, h[j * 8 + 6] - t0);
            }
        }
#endif
        if (rc)
            return fail(-2, "conv_halo launch (conv %d): %s", l.index, cudaGetErrorString((cudaError_t)rc));
        c->launches += 1;
        if (p.splits > 1) {
            if ((rc = launch_splitk_epilogue(p, st))) return fail(-2, "epilogue launch: %s", cudaGetErrorString((cudaError_t)rc));
            c->launches += 1;
        }
        return 0;
    }
#ifdef B2T_DEV
    const int tiles = B * p.tiles_x * p.tiles_y * ((l.cout + l.BN - 1) / l.BN);
    p.splits = choose_splits(tiles, p.chunks_total, c->n_sm);
    if (p.splits > 1 && (size_t)p.splits * B * l.H * l.W * p.ldp * 4 > c->partial_bytes)
        return fail(-2, "internal: split-K workspace too small for conv %d", l.index);
    if ((rc = launch_conv_umma(l.BN, l.tmA_hi, l.tmA_lo, l.tmB_hi, l.tmB_lo, p, st)))
        return fail(-2, "conv_umma launch (conv %d): %s", l.index, cudaGetErrorString((cudaError_t)rc));
    c->launches += 1;
    if (p.splits > 1) {
        if ((rc = launch_splitk_epilogue(p, st))) return fail(-2, "epilogue launch: %s", cudaGetErrorString((cudaError_t)rc));
        c->launches += 1;
    }
    return 0;
#else
    return fail(-1, "engine %d is not part of the release build", c->cfg.engine);
#endif
}

static int run_conv1(b2t_ctx *c, const void *frames, int dtype, int B, cudaStream_t st) {
    ConvLayer &l = c->conv[1];
    if (l.pm && dtype == B2T_FRAME_U8) {
        // tensor-core path: frame -> fp16 integers (8 ch / pixel), then conv_pm_kernel mode 1
        int rc = 0;
        if (frames) {                                     // NULL: b2t_ingest_frames already filled the fp16 frame copy
            const long long npix = (long long)B * l.H * l.W;
            rc = launch_frames_to_c8(frames, c->d_ws + c->off_c8, npix, npix, 0, st);
            if (rc) return fail(-2, "frames_to_c8 launch: %s", cudaGetErrorString((cudaError_t)rc));
            c->launches += 1;
        }
        ConvParams p;
        memset(&p, 0, sizeof p);
        p.B = B; p.H = l.H; p.W = l.W; p.ksize = 3; p.cin_chunks = 1; p.Cout = 32; p.kbytes = 16;
        p.act = 1; p.pool = 1; p.splits = 1; p.ldp = 32;
        p.scale = reinterpret_cast<const float *>(c->d_blob + c->off_s1pm);
        p.bias = reinterpret_cast<const float *>(c->d_blob + c->off_b1);
        p.out = dest_planes(c, l.out_buf, 0, l.H, l.W, DEST_PLAIN);
        p.pout = dest_planes(c, l.pout_buf, 0, l.H / 2, l.W / 2, DEST_PLAIN);
        p.hC = l.pmC; p.hP = l.pmP; p.hR = l.pmR; p.h_rows = l.pm_rows; p.hN = 128;
        p.h_plane_bytes = l.pm_plane_bytes;
        p.h_tiles_x = (l.W + p.hC - 1) / p.hC; p.h_tiles_y = (l.H + p.hR - 1) / p.hR;
        p.pm_mode = 1; p.pm_n = 32; p.pm_tmem_cols = 128; p.pm_stage_ld = 36; p.pm_glog = 2;
        p.pw_patch_bytes = l.pm_plane_bytes;
        p.pw_stage_bytes = 0;                                  // mode 1 pools in registers: no staging buffers
        p.pm_w = c->d_blob + c->off_w1pm; p.pm_w_bytes = 3 * 4096;
        p.magic_tx = magic_for(p.h_tiles_x); p.magic_ty = magic_for(p.h_tiles_y);
#ifdef B2T_DEV
        { static const int dbg = dev_env("B2T_DBG", 0); p.dbg = dbg; }
        static long long *d_trace = nullptr;
        if (getenv("B2T_TRACE_CONV") && atoi(getenv("B2T_TRACE_CONV")) == 1) {
            if (!d_trace) cudaMalloc(&d_trace, 64 * 16 * 8);
            cudaMemsetAsync(d_trace, 0, 64 * 16 * 8, st);
            p.trace = d_trace;
        }
#endif
        if ((rc = launch_conv_pm(c->n_sm, l.tmXp_hi, l.tmXp_lo, l.tmXp_hi, l.tmXp_lo, p, st)))
            return fail(-2, "conv_pm launch (conv 1): %s", cudaGetErrorString((cudaError_t)rc));
#ifdef B2T_DEV
        if (p.trace) {
            long long h[64 * 16];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, d_trace, sizeof h, cudaMemcpyDeviceToHost);
            const long long t0 = h[1];
            fprintf(stderr, "[trace conv_1] item: tma_issue | mma: start acc_ok patch_ok issued | epi: wait acc_full phase1 stored (cycles rel.)\n");
            for (int j = 0; j < 24; ++j)
                fprintf(stderr, "  %2d: %7lld | %7lld %7lld %7lld %7lld | %7lld %7lld %7lld %7lld\n", j, h[j * 16] - t0, h[j * 16 + 1] - t0,
                        h[j * 16 + 2] - t0, h[j * 16 + 3] - t0, h[j * 16 + 4] - t0, h[j * 16 + 5] - t0, h[j * 16 + 6] - t0,
                        h[j * 16 + 7] - t0, h[j * 16 + 8] - t0);
        }
#endif
        c->launches += 1;
        return 0;
    }
    if (!frames) return fail(-1, "b2t_yolo_forward: frames == NULL needs uint8 frames ingested with b2t_ingest_frames");
    Conv1Params p;
    memset(&p, 0, sizeof p);
    p.frames = frames; p.dtype = dtype; p.B = B; p.H = l.H; p.W = l.W;
    p.w = reinterpret_cast<const float *>(c->d_blob + c->off_w1);
    p.scale = reinterpret_cast<const float *>(c->d_blob + c->off_s1);
    p.bias = reinterpret_cast<const float *>(c->d_blob + c->off_b1);
    p.lut = reinterpret_cast<const float *>(c->d_blob + c->off_lut);
    p.out = dest_planes(c, l.out_buf, 0, l.H, l.W, DEST_PLAIN);
    p.pout = dest_planes(c, l.pout_buf, 0, l.H / 2, l.W / 2, DEST_PLAIN);
    const int rc = launch_conv1(p, st);
    if (rc) return fail(-2, "conv1 launch: %s", cudaGetErrorString((cudaError_t)rc));
    c->launches += 1;
    return 0;
}

// tiny graph: maxpool size 2 stride 1 (parser.c:471-486: padding 0; maxpool_layer.c:79-114: window [i, i+1] x [j, j+1]
// clipped at the border) of layer l's output into its ps1 buffer, on split planes
static int run_pool_s1(b2t_ctx *c, const ConvLayer &l, int B, cudaStream_t st) {
    const ActBuf &in = c->bufs[l.out_buf], &out = c->bufs[l.ps1_buf];
    const int rc = launch_pool_s1(in.hi, in.plane, out.hi, out.plane, B, l.H, l.W, in.C, st);
    if (rc) return fail(-2, "pool_s1 launch: %s", cudaGetErrorString((cudaError_t)rc));
    c->launches += 1;
    return 0;
}

// darknet reorg (reorg_layer.c:91-110 -> blas.c:9-30, stride 2, forward = 0) of conv_21's plain output into channels
// [0, 4*cout) of the concat buffer
static int run_reorg(b2t_ctx *c, const ConvLayer &l, int B, cudaStream_t st) {
    const ActBuf &in = c->bufs[l.out_buf], &out = c->bufs[l.reorg_buf];
    const int rc = launch_reorg_gather(in.hi, in.plane, in.C, out.hi, out.plane, out.C, B, l.H, l.W, l.cout, st);
    if (rc) return fail(-2, "reorg launch: %s", cudaGetErrorString((cudaError_t)rc));
    c->launches += 1;
    return 0;
}

static int forward_impl(b2t_ctx *c, const void *frames, int dtype, int B, float *logits_user, cudaStream_t st,
                        cudaEvent_t *ev /* 24 events or NULL */, int first = 1, int last = -1) {
    if (!c || !c->finalized) return fail(-1, "b2t_yolo_forward: context not finalized");
    const int NC = c->n_conv;
    if (last < 0 || last == 23) last = last < 0 || NC < 23 ? NC : 23;      // 23 = "to the head" for callers written against the YOLOv2 graph
    if (first < 1 || last > NC || first > last) return fail(-1, "b2t_yolo_forward: bad layer range [%d, %d]", first, last);
    if (first == 1 && !frames && !(c->conv[1].pm && dtype == B2T_FRAME_U8))
        return fail(-1, "b2t_yolo_forward: null frames");
    if (B < 1 || B > c->cfg.max_batch) return fail(-1, "batch %d outside [1, max_batch=%d]", B, c->cfg.max_batch);
    if (dtype != B2T_FRAME_U8 && dtype != B2T_FRAME_F32) return fail(-1, "bad frame dtype %d", dtype);
    int rc;
    if (ev) cudaEventRecord(ev[0], st);
    if (first == 1) {
        if ((rc = run_conv1(c, frames, dtype, B, st))) return rc;
        if (ev) cudaEventRecord(ev[1], st);
    }
    float *logits = reinterpret_cast<float *>(c->d_ws + c->off_logits);
    // small batches: consecutive layers run inside ONE persistent cooperative launch (conv_chain_kernel); per-layer
    // event timing (b2t_profile_forward) keeps the one-kernel-per-layer schedule
    // (measured, profiles/r2_batch_sweep_*.txt: the chain wins for one 416x416 frame; at 608x608 the layers have enough
    // items for the per-layer kernels even at batch 1)
    const bool chain = !ev && c->chain_max_batch > 0 && B <= c->chain_max_batch && c->d_chain_counter &&
                       (c->cfg.image_h <= 448 || c->chain_forced);
    ChainBuilder *cb = nullptr;
#ifdef B2T_DEV
    static long long *chain_trace = nullptr;
    int chain_idx[32], chain_splits[32];
    if (dev_env("B2T_TRACE_CONV", 0) == 100 && !chain_trace) cudaMalloc(&chain_trace, 24 * 8 * 8);
#endif
    auto flush_chain = [&]() -> int {
        if (!cb) return 0;
        int r = 0;
        if (chain_layers(cb) > 0) {
            r = launch_conv_chain(c->n_sm, cb, st);
            c->launches += 1;
#ifdef B2T_DEV
            if (!r && chain_trace) {
                long long h[24 * 8];
                cudaStreamSynchronize(st);
                cudaMemcpy(h, chain_trace, sizeof h, cudaMemcpyDeviceToHost);
                fprintf(stderr, "[trace chain B=%d] layer: input_ready patch_ok w_ok | epi: acc_ok items_done barrier_A done (clocks rel. to the first stamp)\n", B);
                long long t0 = h[0];
                for (int j = 0; j < chain_layers(cb); ++j)
                    fprintf(stderr, "  conv_%-2d splits=%-2d: %8lld %8lld %8lld | %8lld %8lld %8lld %8lld\n", chain_idx[j], chain_splits[j], h[j * 8] - t0,
                            h[j * 8 + 1] - t0, h[j * 8 + 2] - t0, h[j * 8 + 3] - t0, h[j * 8 + 4] - t0, h[j * 8 + 5] - t0, h[j * 8 + 6] - t0);
            }
#endif
        }
        chain_free(cb);
        cb = nullptr;
        return r ? fail(-2, "conv_chain launch: %s", cudaGetErrorString((cudaError_t)r)) : 0;
    };
    for (int i = first < 2 ? 2 : first; i <= last; ++i) {
        ConvLayer &l = c->conv[i];
        if (chain && chain_eligible(c, l)) {
            ConvParams p;
            base_params(c, l, B, i == NC ? logits : nullptr, 0, 0, p);
            const int ctas = halo_geometry(c, l, B, p, true);
            if (ctas < 0) { if (cb) chain_free(cb); return ctas; }
            if (!cb) cb = chain_new(c->d_chain_counter);
#ifdef B2T_DEV
            if (chain_trace && chain_layers(cb) == 0) { cudaMemsetAsync(chain_trace, 0, 24 * 8 * 8, st); p.trace = chain_trace; }
            if (chain_layers(cb) < 32) { chain_idx[chain_layers(cb)] = i; chain_splits[chain_layers(cb)] = p.splits; }
#endif
            if (chain_add(cb, l.tmX_hi, l.tmX_lo, l.tmW_hi, l.tmW_lo, p)) {       // full: launch what we have, start anew
                if ((rc = flush_chain())) return rc;
                cb = chain_new(c->d_chain_counter);
                chain_add(cb, l.tmX_hi, l.tmX_lo, l.tmW_hi, l.tmW_lo, p);
            }
            if (l.ps1_buf >= 0 || l.reorg_buf >= 0) {      // a stride-1 pool / reorg is its own kernel: the chain ends here
                if ((rc = flush_chain())) return rc;
                if (l.ps1_buf >= 0 && (rc = run_pool_s1(c, l, B, st))) return rc;
                if (l.reorg_buf >= 0 && (rc = run_reorg(c, l, B, st))) return rc;
            }
            continue;
        }
        if ((rc = flush_chain())) return rc;
        if ((rc = run_conv(c, l, B, i == NC ? logits : nullptr, st))) return rc;
        if (l.ps1_buf >= 0 && (rc = run_pool_s1(c, l, B, st))) return rc;
        if (l.reorg_buf >= 0 && (rc = run_reorg(c, l, B, st))) return rc;
        if (ev) cudaEventRecord(ev[i], st);
    }
    if ((rc = flush_chain())) return rc;
    if (last == NC && logits_user && logits_user != logits)
        CK(cudaMemcpyAsync(logits_user, logits, (size_t)B * c->G * c->G * c->A * c->D * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int b2t_yolo_forward(b2t_ctx *c, const void *frames, int dtype, int B, float *logits_dev, void *stream) {
    return forward_impl(c, frames, dtype, B, logits_dev, (cudaStream_t)stream, nullptr);
}

// conv_first .. conv_last of the same forward pass (1 <= first <= last <= 23): lets a host pipeline put the layers
// whose outputs the tracker tail reads (conv_9 and later) behind an event while conv_1..8 of the next step already run.
extern "C" int b2t_yolo_forward_range(b2t_ctx *c, const void *frames, int dtype, int B, int first, int last,
                                      float *logits_dev, void *stream) {
    return forward_impl(c, frames, dtype, B, logits_dev, (cudaStream_t)stream, nullptr, first, last);
}

// Frame ingest without a staging copy: the uint8 frames of a step are n_seg segments of seg_frames consecutive frames,
// seg_stride_bytes apart (S streams x T frames of longer device-resident clips).  Converts them into the context's fp16
// frame buffer; a following b2t_yolo_forward[_range](frames_dev = NULL, B2T_FRAME_U8, batch = n_seg * seg_frames)
// starts at conv_1.  Lets a host keep the forward pass in a CUDA graph while the input pointer changes every step.
extern "C" int b2t_ingest_frames(b2t_ctx *c, const unsigned char *frames, int n_seg, int seg_frames, long long seg_stride_bytes,
                                 void *stream) {
    if (!c || !c->finalized || !frames) return fail(-1, "b2t_ingest_frames: bad arguments");
    if (!c->conv[1].pm) return fail(-1, "b2t_ingest_frames: the tensor-core conv_1 path is not active");
    const long long B = (long long)n_seg * seg_frames;
    if (n_seg < 1 || seg_frames < 1 || B > c->cfg.max_batch) return fail(-1, "b2t_ingest_frames: %d x %d frames outside [1, max_batch=%d]", n_seg, seg_frames, c->cfg.max_batch);
    const ConvLayer &l = c->conv[1];
    const long long seg_pix = (long long)seg_frames * l.H * l.W;
    if (n_seg > 1 && seg_stride_bytes < seg_pix * 3) return fail(-1, "b2t_ingest_frames: segments overlap");
    const int rc = launch_frames_to_c8(frames, c->d_ws + c->off_c8, B * l.H * l.W, seg_pix, seg_stride_bytes, (cudaStream_t)stream);
    if (rc) return fail(-2, "frames_to_c8 launch: %s", cudaGetErrorString((cudaError_t)rc));
    c->launches += 1;
    return 0;
}

extern "C" const float *b2t_logits(const b2t_ctx *c) {
    return (c && c->d_ws) ? reinterpret_cast<const float *>(c->d_ws + c->off_logits) : nullptr;
}

extern "C" int b2t_profile_forward(b2t_ctx *c, const void *frames, int dtype, int B, float *ms, double *bytes,
                                   void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t ev[24];
    for (auto &e : ev) CK(cudaEventCreate(&e));
    int rc = forward_impl(c, frames, dtype, B, nullptr, st, ev);
    if (!rc) {
        CK(cudaStreamSynchronize(st));
        for (int i = 1; i <= 23; ++i) {
            if (i > c->n_conv) { ms[i - 1] = 0.f; if (bytes) bytes[i - 1] = 0.0; continue; }
            CK(cudaEventElapsedTime(&ms[i - 1], ev[i - 1], ev[i]));
            if (bytes) {
                const ConvLayer &l = c->conv[i];
                const double px = (double)B * l.H * l.W;
                const double w = (double)l.k * l.k * l.cin * l.cout;
                const double in = px * l.cin, out = px * l.cout * (l.pool ? (i == 13 ? 1.25 : 0.25) : 1.0);
                bytes[i - 1] = 4.0 * (w + out) + (i == 1 ? (dtype == B2T_FRAME_U8 ? 1.0 : 4.0) : 4.0) * in;
            }
        }
    }
    for (auto &e : ev) cudaEventDestroy(e);
    return rc;
}

static int lookup(const b2t_ctx *c, const char *name, int *buf, int *choff, int *C) {
    if (!c || !name) return fail(-1, "null argument");
    auto it = c->buf_by_name.find(name);
    if (it == c->buf_by_name.end())
        return fail(-1, "layer '%s' is not kept by this context (pre-pool outputs need config.reserved[0]=1)", name);
    *buf = it->second;
    auto jt = c->choff_by_name.find(name);
    *choff = jt == c->choff_by_name.end() ? 0 : jt->second;
    auto kt = c->cout_by_name.find(name);
    *C = kt == c->cout_by_name.end() ? c->bufs[*buf].C : kt->second;
    return 0;
}

extern "C" int b2t_layer_dims(const b2t_ctx *c, const char *name, int *h, int *w, int *ch) {
    if (name && c && (!strcmp(name, "conv_23") || !strcmp(name, ("conv_" + std::to_string(c->n_conv)).c_str()))) {
        if (h) *h = c->G; if (w) *w = c->G; if (ch) *ch = c->A * c->D;
        return 0;
    }
    int buf, off, C;
    const int rc = lookup(c, name, &buf, &off, &C);
    if (rc) return rc;
    if (h) *h = c->bufs[buf].H;
    if (w) *w = c->bufs[buf].W;
    if (ch) *ch = C;
    return 0;
}

extern "C" long b2t_extract(b2t_ctx *c, const char *name, int B, float *out, void *stream) {
    if (!c || !c->finalized || !out) return fail(-1, "b2t_extract: bad arguments");
    if (B < 1 || B > c->cfg.max_batch) return fail(-1, "bad batch");
    cudaStream_t st = (cudaStream_t)stream;
    if (!strcmp(name, "conv_23") || !strcmp(name, ("conv_" + std::to_string(c->n_conv)).c_str())) {
        const size_t n = (size_t)c->G * c->G * c->A * c->D;
        CK(cudaMemcpyAsync(out, c->d_ws + c->off_logits, n * B * 4, cudaMemcpyDeviceToDevice, st));
        return (long)n;
    }
    int buf, off, C;
    const int rc = lookup(c, name, &buf, &off, &C);
    if (rc) return rc;
    const ActBuf &b = c->bufs[buf];
    const long long npix = (long long)B * b.H * b.W;
    const int e = launch_planes_to_f32(b.hi, b.plane, b.C, off, C, npix, out, st);
    if (e) return fail(-2, "planes_to_f32 launch: %s", cudaGetErrorString((cudaError_t)e));
    c->launches += 1;
    return (long)b.H * b.W * C;
}

// ------------------------------------------------------------------------------------------------ decode
static int fill_decode(DecodeParams &p, const float *logits, int B, int gh, int gw, int nb, int nc, float t1, float t2,
                       const float *anchors, float *boxes, int *counts, int maxb) {
    if (!logits || !boxes || !counts || !anchors) return fail(-1, "decode: null pointer");
    if (nb < 1 || nb > 16) return fail(-1, "decode: n_box %d outside [1,16]", nb);
    if (gh * gw * nb > 32767) return fail(-1, "decode: more than 32767 anchors per frame");
    if (B < 1 || nc < 1 || maxb < 1) return fail(-1, "decode: bad sizes");
    memset(&p, 0, sizeof p);
    p.logits = logits; p.B = B; p.GH = gh; p.GW = gw; p.A = nb; p.C = nc;
    p.obj_thr = t1; p.nms_thr = t2;
    for (int i = 0; i < 2 * nb; ++i) p.anchors[i] = anchors[i];
    p.boxes = boxes; p.counts = counts; p.max_boxes = maxb;
    return 0;
}

extern "C" int b2t_decode_nms(b2t_ctx *c, const float *logits, int B, int gh, int gw, int nb, int nc, float obj_thr,
                              float nms_thr, const float *anchors, float *boxes, int *counts, int maxb, void *stream) {
    DecodeParams p;
    int rc = fill_decode(p, logits, B, gh, gw, nb, nc, obj_thr, nms_thr, anchors, boxes, counts, maxb);
    if (rc) return rc;
    if ((rc = launch_decode(false, p, (cudaStream_t)stream))) return fail(-2, "decode launch: %s", cudaGetErrorString((cudaError_t)rc));
    if (c) c->launches += 1;
    return 0;
}

extern "C" int b2t_region_detect(b2t_ctx *c, const float *logits, int B, int gh, int gw, int nb, int nc, float thresh,
                                 float nms_thr, const float *anchors, int ow, int oh, int nw, int nh, float *dets,
                                 int *counts, int maxd, void *stream) {
    DecodeParams p;
    int rc = fill_decode(p, logits, B, gh, gw, nb, nc, thresh, nms_thr, anchors, dets, counts, maxd);
    if (rc) return rc;
    if (ow < 1 || oh < 1 || nw < 1 || nh < 1) return fail(-1, "region_detect: bad frame size");
    p.orig_w = ow; p.orig_h = oh; p.net_w = nw; p.net_h = nh;
    if ((rc = launch_decode(true, p, (cudaStream_t)stream))) return fail(-2, "decode launch: %s", cudaGetErrorString((cudaError_t)rc));
    if (c) c->launches += 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------ LSTM head
static const int kMaxT = 16;
struct b2t_lstm {
    b2t_ctx *ctx;
    int n_feat, n_det, units, n_out, max_streams;
    float *d_wp = nullptr, *d_bias = nullptr, *d_wd = nullptr, *d_bd = nullptr;
    float *d_h[2] = {nullptr, nullptr}, *d_c = nullptr;
    float *d_zx = nullptr, *d_hseq = nullptr;      // (max_streams*kMaxT, 4u) / (max_streams*kMaxT, u) sequence scratch
    unsigned int *d_counter = nullptr;             // grid barrier of lstm_seq_kernel
    int cur = 0;
    bool have = false;
};

extern "C" int b2t_lstm_create(b2t_ctx *ctx, int n_feat, int n_det, int units, int n_out, int max_streams, b2t_lstm **out) {
    if (!out || n_feat < 1 || n_det < 0 || units < 4 || units % 4 || n_out < 1 || max_streams < 1)
        return fail(-1, "b2t_lstm_create: bad arguments");
    b2t_lstm *l = new b2t_lstm();
    l->ctx = ctx; l->n_feat = n_feat; l->n_det = n_det; l->units = units; l->n_out = n_out; l->max_streams = max_streams;
    const size_t rows = (size_t)n_feat + n_det + units;
    if (cudaMalloc(&l->d_wp, rows * 4 * units * 4) || cudaMalloc(&l->d_bias, 4 * units * 4) ||
        cudaMalloc(&l->d_wd, (size_t)units * n_out * 4) || cudaMalloc(&l->d_bd, n_out * 4) ||
        cudaMalloc(&l->d_h[0], (size_t)max_streams * units * 4) || cudaMalloc(&l->d_h[1], (size_t)max_streams * units * 4) ||
        cudaMalloc(&l->d_c, (size_t)max_streams * units * 4) ||
        cudaMalloc(&l->d_zx, (size_t)max_streams * kMaxT * 4 * units * 4) ||
        cudaMalloc(&l->d_hseq, (size_t)max_streams * kMaxT * units * 4) || cudaMalloc(&l->d_counter, 256)) {
        b2t_lstm_destroy(l);
        return fail(-2, "b2t_lstm_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    cudaMemset(l->d_h[0], 0, (size_t)max_streams * units * 4);
    cudaMemset(l->d_h[1], 0, (size_t)max_streams * units * 4);
    cudaMemset(l->d_c, 0, (size_t)max_streams * units * 4);
    *out = l;
    return 0;
}

extern "C" void b2t_lstm_destroy(b2t_lstm *l) {
    if (!l) return;
    cudaFree(l->d_wp); cudaFree(l->d_bias); cudaFree(l->d_wd); cudaFree(l->d_bd);
    cudaFree(l->d_h[0]); cudaFree(l->d_h[1]); cudaFree(l->d_c); cudaFree(l->d_zx); cudaFree(l->d_hseq);
    cudaFree(l->d_counter);
    delete l;
}

extern "C" int b2t_lstm_set_weights(b2t_lstm *l, const float *kernel, const float *recurrent, const float *bias,
                                    const float *dk, const float *db, void *stream) {
    if (!l || !kernel || !recurrent || !bias || !dk || !db) return fail(-1, "b2t_lstm_set_weights: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int u = l->units, nx = l->n_feat + l->n_det, rows = nx + u;
    // [units/4][rows][gate][4]: the 16 columns one CTA needs are contiguous per input row
    std::vector<float> wp((size_t)rows * 4 * u);
    for (int ub = 0; ub < u / 4; ++ub)
        for (int k = 0; k < rows; ++k)
            for (int g = 0; g < 4; ++g)
                for (int uu = 0; uu < 4; ++uu) {
                    const int col = g * u + ub * 4 + uu;
                    const float v = k < nx ? kernel[(size_t)k * 4 * u + col] : recurrent[(size_t)(k - nx) * 4 * u + col];
                    wp[(((size_t)ub * rows + k) * 4 + g) * 4 + uu] = v;
                }
    CK(cudaMemcpyAsync(l->d_wp, wp.data(), wp.size() * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(l->d_bias, bias, 4 * u * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(l->d_wd, dk, (size_t)u * l->n_out * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(l->d_bd, db, l->n_out * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));   // wp is a host temporary
    l->have = true;
    return 0;
}

extern "C" int b2t_lstm_reset(b2t_lstm *l, int s, void *stream) {
    if (!l) return fail(-1, "null lstm");
    cudaStream_t st = (cudaStream_t)stream;
    if (s >= l->max_streams) return fail(-1, "stream index %d >= max_streams %d", s, l->max_streams);
    const size_t off = s < 0 ? 0 : (size_t)s * l->units, n = (s < 0 ? (size_t)l->max_streams : 1) * l->units;
    CK(cudaMemsetAsync(l->d_h[0] + off, 0, n * 4, st));
    CK(cudaMemsetAsync(l->d_h[1] + off, 0, n * 4, st));
    CK(cudaMemsetAsync(l->d_c + off, 0, n * 4, st));
    return 0;
}

// streams [slot0, slot0 + S) of the head's state are stepped; every other stream keeps its (h, c)
extern "C" int b2t_lstm_step_slots(b2t_lstm *l, int slot0, const float *fv, int fv_stride, const float *det, int det_stride,
                                   int S, float *y, int y_stride, int hard_sigmoid, void *stream) {
    if (!l || !l->have) return fail(-1, "b2t_lstm_step: weights not set");
    if (!fv || (!det && l->n_det) || !y) return fail(-1, "b2t_lstm_step: null pointer");
    if (S < 1 || slot0 < 0 || slot0 + S > l->max_streams)
        return fail(-1, "streams [%d,%d) outside [0,%d)", slot0, slot0 + S, l->max_streams);
    cudaStream_t st = (cudaStream_t)stream;
    if (fv_stride <= 0) fv_stride = l->n_feat;
    if (det_stride <= 0) det_stride = l->n_det;
    if (y_stride <= 0) y_stride = l->n_out;
    const int nxt = l->cur ^ 1;
    const size_t o = (size_t)slot0 * l->units;
    {
        LstmParams p;
        memset(&p, 0, sizeof p);
        p.wp = l->d_wp; p.bias = l->d_bias; p.fv = fv; p.det = det;
        p.fv_stride = fv_stride; p.det_stride = det_stride;
        p.h_in = l->d_h[l->cur] + o; p.h_out = l->d_h[nxt] + o; p.c = l->d_c + o;
        p.n_feat = l->n_feat; p.n_det = l->n_det; p.units = l->units; p.S = S;
        p.hard_sigmoid = hard_sigmoid;
        const int rc = launch_lstm_gates(p, st);
        if (rc) return fail(-2, "lstm launch: %s", cudaGetErrorString((cudaError_t)rc));
        if (l->ctx) l->ctx->launches += 1;
    }
    // streams not stepped keep their state: copy them across the double buffer
    if (slot0 > 0)
        CK(cudaMemcpyAsync(l->d_h[nxt], l->d_h[l->cur], o * 4, cudaMemcpyDeviceToDevice, st));
    if (slot0 + S < l->max_streams)
        CK(cudaMemcpyAsync(l->d_h[nxt] + o + (size_t)S * l->units, l->d_h[l->cur] + o + (size_t)S * l->units,
                           (size_t)(l->max_streams - slot0 - S) * l->units * 4, cudaMemcpyDeviceToDevice, st));
    l->cur = nxt;
    const int rc = launch_dense_sigmoid(l->d_h[l->cur] + o, l->d_wd, l->d_bd, l->units, l->n_out, S, y, y_stride, st);
    if (rc) return fail(-2, "dense launch: %s", cudaGetErrorString((cudaError_t)rc));
    if (l->ctx) l->ctx->launches += 1;
    return 0;
}

extern "C" int b2t_lstm_step(b2t_lstm *l, const float *fv, int fv_stride, const float *det, int det_stride, int S,
                             float *y, int y_stride, int hard_sigmoid, void *stream) {
    return b2t_lstm_step_slots(l, 0, fv, fv_stride, det, det_stride, S, y, y_stride, hard_sigmoid, stream);
}

extern "C" int b2t_lstm_sequence(b2t_lstm *l, const float *fv, const float *det, int S, int T, float *y, int reset,
                                 int hard_sigmoid, void *stream) {
    if (!l || !l->have) return fail(-1, "b2t_lstm_sequence: weights not set");
    if (!fv || (!det && l->n_det) || !y) return fail(-1, "b2t_lstm_sequence: null pointer");
    if (S < 1 || S > l->max_streams || T < 1 || T > kMaxT) return fail(-1, "b2t_lstm_sequence: S=%d T=%d out of range", S, T);
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (reset && (rc = b2t_lstm_reset(l, -1, stream))) return rc;
    // one step per stream (streams advancing frame by frame): the full-step kernel does projection + recurrence in one
    // pass over the weights -- two launches instead of three, none of them cooperative
    if (T == 1) return b2t_lstm_step_slots(l, 0, fv, l->n_feat, det, l->n_det, S, y, l->n_out, hard_sigmoid, stream);
    const int u = l->units, R = S * T;
    LstmParams p;
    memset(&p, 0, sizeof p);
    p.wp = l->d_wp; p.bias = l->d_bias;
    p.n_feat = l->n_feat; p.n_det = l->n_det; p.units = u; p.hard_sigmoid = hard_sigmoid;
    // (1) input projection of all S*T rows at once: it does not depend on the recurrent state
    p.mode = 1; p.fv = fv; p.det = det; p.fv_stride = l->n_feat; p.det_stride = l->n_det;
    p.S = R; p.zx = l->d_zx; p.zx_stride = 4 * u; p.h_in = l->d_h[l->cur]; p.h_out = l->d_h[l->cur]; p.c = l->d_c;
    rc = launch_lstm_proj(p, st);                                  // one pass over the weights for all rows
    if (rc == -1) rc = launch_lstm_gates(p, st);
    if (rc) return fail(-2, "lstm launch: %s", cudaGetErrorString((cudaError_t)rc));
    // (2) T sequential recurrent steps over the S streams (h*U + gates); h_t kept for the Dense head:
    //     one fused launch with a grid barrier between steps when the shape allows, else one launch per step
    p.mode = 2; p.S = S; p.zx = l->d_zx; p.h_seq = l->d_hseq;
    rc = launch_lstm_seq(p, T, l->d_h[l->cur], l->d_h[l->cur ^ 1], l->d_counter, l->ctx ? l->ctx->n_sm : 132, st);
    if (rc > 0) return fail(-2, "lstm launch: %s", cudaGetErrorString((cudaError_t)rc));
    const bool fused = rc == 0;
    if (fused) {
        if ((T & 1) && S < l->max_streams)
            CK(cudaMemcpyAsync(l->d_h[l->cur ^ 1] + (size_t)S * u, l->d_h[l->cur] + (size_t)S * u,
                               (size_t)(l->max_streams - S) * u * 4, cudaMemcpyDeviceToDevice, st));
        l->cur ^= (T & 1);
    }
    for (int t = 0; t < T && !fused; ++t) {
        const int nxt = l->cur ^ 1;
        p.mode = 2; p.S = S;
        p.zx = l->d_zx + (size_t)t * 4 * u; p.zx_stride = T * 4 * u;
        p.h_in = l->d_h[l->cur]; p.h_out = l->d_h[nxt];
        p.h_seq = l->d_hseq + (size_t)t * u; p.h_seq_stride = T * u;
        if ((rc = launch_lstm_gates(p, st))) return fail(-2, "lstm launch: %s", cudaGetErrorString((cudaError_t)rc));
        if (S < l->max_streams)
            CK(cudaMemcpyAsync(l->d_h[nxt] + (size_t)S * u, l->d_h[l->cur] + (size_t)S * u,
                               (size_t)(l->max_streams - S) * u * 4, cudaMemcpyDeviceToDevice, st));
        l->cur = nxt;
    }
    // (3) Dense(n_out, sigmoid) on all S*T hidden states
    if ((rc = launch_dense_sigmoid(l->d_hseq, l->d_wd, l->d_bd, u, l->n_out, R, y, l->n_out, st)))
        return fail(-2, "dense launch: %s", cudaGetErrorString((cudaError_t)rc));
    if (l->ctx) l->ctx->launches += fused ? 3 : T + 2;
    return 0;
}

// cv2.resize(frame, (dst_w, dst_h)) for (B, src_h, src_w, 3) uint8 frames on the device, bit-identical to OpenCV's
// INTER_LINEAR (KerasYOLO.py:526, MultiObjDetTracker.py:300); see csrc/ingest.cu for the algorithm.
extern "C" int b2t_resize_frames(b2t_ctx *c, const unsigned char *src_dev, int src_h, int src_w, int B,
                                 unsigned char *dst_dev, int dst_h, int dst_w, void *stream) {
    if (!c || !src_dev || !dst_dev) return fail(-1, "b2t_resize_frames: null argument");
    if (src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1 || B < 1) return fail(-1, "b2t_resize_frames: bad geometry");
    if (dst_h > 8192 || dst_w > 8192) return fail(-1, "b2t_resize_frames: destination larger than 8192");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->cfg.device));
    if (!c->d_rs_tab) CK(cudaMalloc(&c->d_rs_tab, 2 * 8192 * sizeof(int4)));
    if (c->rs_geom[0] != src_h || c->rs_geom[1] != src_w || c->rs_geom[2] != dst_h || c->rs_geom[3] != dst_w) {
        std::vector<int4> tab(dst_w + dst_h);
        const double scale_x = 1.0 / ((double)dst_w / src_w), scale_y = 1.0 / ((double)dst_h / src_h);
        for (int dx = 0; dx < dst_w; ++dx) {
            float fx = (float)((dx + 0.5) * scale_x - 0.5);
            int sx = (int)floorf(fx);
            fx -= sx;
            if (sx < 0) { fx = 0.f; sx = 0; }
            if (sx >= src_w - 1) { fx = 0.f; sx = src_w - 1; }
            const int a0 = (int)lrintf((1.f - fx) * 2048.f), a1 = (int)lrintf(fx * 2048.f);
            tab[dx] = make_int4(sx, sx + 1 < src_w ? sx + 1 : src_w - 1, a0, a1);
        }
        for (int dy = 0; dy < dst_h; ++dy) {
            float fy = (float)((dy + 0.5) * scale_y - 0.5);
            const int sy = (int)floorf(fy);
            fy -= sy;
            const int b0 = (int)lrintf((1.f - fy) * 2048.f), b1 = (int)lrintf(fy * 2048.f);
            const int s0 = sy < 0 ? 0 : (sy > src_h - 1 ? src_h - 1 : sy);
            const int s1 = sy + 1 < 0 ? 0 : (sy + 1 > src_h - 1 ? src_h - 1 : sy + 1);
            tab[dst_w + dy] = make_int4(s0, s1, b0, b1);
        }
        CK(cudaMemcpyAsync(c->d_rs_tab, tab.data(), tab.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));                  // `tab` is a host temporary
        c->rs_geom[0] = src_h; c->rs_geom[1] = src_w; c->rs_geom[2] = dst_h; c->rs_geom[3] = dst_w;
    }
    ResizeParams p;
    p.src = src_dev; p.dst = dst_dev; p.B = B; p.src_h = src_h; p.src_w = src_w; p.dst_h = dst_h; p.dst_w = dst_w;
    p.xtab = c->d_rs_tab; p.ytab = c->d_rs_tab + dst_w;
    const int rc = launch_resize_bilinear_u8(p, st);
    if (rc) return fail(-2, "resize launch: %s", cudaGetErrorString((cudaError_t)rc));
    c->launches += 1;
    return 0;
}

// load_image_color + letterbox_image (image.c:1442-1482, :960-979) for a batch of uint8 frames already on the device.
extern "C" int b2t_letterbox_frames(b2t_ctx *c, const unsigned char *src_dev, int src_h, int src_w, int B, int swap_rb,
                                    float *dst_dev, void *stream) {
    if (!c || !src_dev || !dst_dev) return fail(-1, "b2t_letterbox_frames: null argument");
    if (src_h < 1 || src_w < 1 || B < 1) return fail(-1, "b2t_letterbox_frames: bad geometry");
    LetterboxParams p;
    p.src = src_dev; p.dst = dst_dev; p.B = B; p.src_h = src_h; p.src_w = src_w;
    p.net_h = c->cfg.image_h; p.net_w = c->cfg.image_w; p.swap_rb = swap_rb ? 1 : 0;
    if (((float)p.net_w / src_w) < ((float)p.net_h / src_h)) { p.new_w = p.net_w; p.new_h = (src_h * p.net_w) / src_w; }
    else { p.new_h = p.net_h; p.new_w = (src_w * p.net_h) / src_h; }
    if (p.new_w < 2 || p.new_h < 2) return fail(-1, "b2t_letterbox_frames: frame aspect ratio too extreme");
    const int rc = launch_letterbox_u8(p, (cudaStream_t)stream);
    if (rc) return fail(-2, "letterbox launch: %s", cudaGetErrorString((cudaError_t)rc));
    c->launches += 1;
    return 0;
}

extern "C" int b2t_pool_features(b2t_ctx *c, const char *name, int B, int mode, int chw_view, float *fv, void *stream) {
    if (!c || !c->finalized || !fv) return fail(-1, "b2t_pool_features: bad arguments");
    if (B < 1 || B > c->cfg.max_batch) return fail(-1, "bad batch");
    int buf, off, C;
    int rc = lookup(c, name, &buf, &off, &C);
    if (rc) return rc;
    const ActBuf &b = c->bufs[buf];
    PoolParams p;
    p.hi = b.hi; p.plane = b.plane; p.pix_stride = b.C; p.ch_off = off;
    p.B = B; p.H = b.H; p.W = b.W; p.C = C; p.mode = mode; p.chw_view = chw_view; p.out = fv;
    if ((rc = launch_pool_features(p, (cudaStream_t)stream))) return fail(-2, "pool launch: %s", cudaGetErrorString((cudaError_t)rc));
    c->launches += 1;
    return 0;
}

extern "C" int b2t_heatmap_from_box(b2t_ctx *c, const float *xywh, int n, int size, float *heat, void *stream) {
    if (!xywh || !heat || n < 1 || size < 1) return fail(-1, "b2t_heatmap_from_box: bad arguments");
    const int rc = launch_heatmap_from_box(xywh, n, size, heat, (cudaStream_t)stream);
    if (rc) return fail(-2, "heatmap launch: %s", cudaGetErrorString((cudaError_t)rc));
    if (c) c->launches += 1;
    return 0;
}

extern "C" int b2t_box_from_heatmap(b2t_ctx *c, const float *heat, int n, int size, float thresh, int *rect, void *stream) {
    if (!heat || !rect || n < 1 || size < 1) return fail(-1, "b2t_box_from_heatmap: bad arguments");
    const int rc = launch_box_from_heatmap(heat, n, size, thresh, rect, (cudaStream_t)stream);
    if (rc) return fail(-2, "heatmap launch: %s", cudaGetErrorString((cudaError_t)rc));
    if (c) c->launches += 1;
    return 0;
}

extern "C" int b2t_select_detection(b2t_ctx *c, const float *dets, const int *counts, int max_dets, int B,
                                    const unsigned char *class_mask_dev, int frame_w, int frame_h, float *det_in,
                                    int heat_size, float *heat, int *chosen, void *stream) {
    if (!dets || !counts || B < 1 || frame_w < 1 || frame_h < 1) return fail(-1, "b2t_select_detection: bad arguments");
    SelectParams p;
    p.dets = dets; p.counts = counts; p.max_dets = max_dets; p.B = B; p.class_mask = class_mask_dev;
    p.frame_w = frame_w; p.frame_h = frame_h; p.det_in = det_in; p.heat_size = heat_size; p.heat = heat; p.chosen = chosen;
    const int rc = launch_select_detection(p, (cudaStream_t)stream);
    if (rc) return fail(-2, "select launch: %s", cudaGetErrorString((cudaError_t)rc));
    if (c) c->launches += 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------ callers after the path
extern "C" int b2t_draw_boxes(b2t_ctx *c, unsigned char *frames_dev, int B, int H, int W, const float *rows_dev,
                              const int *counts_dev, int max_rows, int c0, int c1, int c2, void *stream) {
    if (!frames_dev || !rows_dev || !counts_dev || B < 1 || H < 1 || W < 1 || max_rows < 1) return fail(-1, "b2t_draw_boxes: bad arguments");
    if (max_rows > 65535 || B > 65535) return fail(-1, "b2t_draw_boxes: more than 65535 boxes per frame / frames");
    const int rc = launch_draw_boxes(frames_dev, B, H, W, rows_dev, counts_dev, max_rows, c0, c1, c2, (cudaStream_t)stream);
    if (rc) return fail(-2, "draw_boxes launch: %s", cudaGetErrorString((cudaError_t)rc));
    if (c) c->launches += 1;
    return 0;
}

extern "C" int b2t_overlap_scores(b2t_ctx *c, const double *y_true_dev, const double *y_pred_dev, int n, double *scores_dev,
                                  double *mean_dev, void *stream) {
    if (!y_true_dev || !y_pred_dev || !scores_dev || n < 1) return fail(-1, "b2t_overlap_scores: bad arguments");
    const int rc = launch_overlap_scores(y_true_dev, y_pred_dev, n, scores_dev, mean_dev, (cudaStream_t)stream);
    if (rc) return fail(-2, "overlap_scores launch: %s", cudaGetErrorString((cudaError_t)rc));
    if (c) c->launches += 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------ CUDA graphs
// A step of this path is 30-45 kernel launches of a few microseconds each: replayed eagerly it is launch-latency
// bound.  These four calls let a plain-C host capture any sequence of b2t_* calls made on `stream` (every entry point is
// stream-ordered; the ones that synchronise -- b2t_finalize, b2t_lstm_set_weights, the first b2t_resize_frames of a
// geometry -- must run before the capture) and replay it with one launch.
struct b2t_graph {
    b2t_ctx *ctx;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    long kernels;                 // kernel nodes per replay (for b2t_launch_count)
};

extern "C" int b2t_graph_begin(b2t_ctx *c, void *stream) {
    if (!c || !c->finalized) return fail(-1, "b2t_graph_begin: context not finalized");
    if (!stream) return fail(-1, "b2t_graph_begin: capture needs a non-default stream");
    CK(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
    c->capture_launches0 = c->launches;
    return 0;
}

extern "C" int b2t_graph_end(b2t_ctx *c, void *stream, b2t_graph **out) {
    if (!c || !out) return fail(-1, "b2t_graph_end: null argument");
    cudaGraph_t g = nullptr;
    CK(cudaStreamEndCapture((cudaStream_t)stream, &g));
    if (!g) return fail(-2, "b2t_graph_end: the capture was invalidated (a call synchronised or failed inside it)");
    b2t_graph *bg = new b2t_graph();
    bg->ctx = c; bg->graph = g; bg->exec = nullptr;
    bg->kernels = c->launches - c->capture_launches0;
    c->launches = c->capture_launches0;               // captured launches did not run; replays are counted instead
    cudaError_t e = cudaGraphInstantiate(&bg->exec, g, 0);
    if (e != cudaSuccess) {
        cudaGraphDestroy(g);
        delete bg;
        return fail(-2, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
    *out = bg;
    return 0;
}

extern "C" int b2t_graph_launch(b2t_graph *g, void *stream) {
    if (!g || !g->exec) return fail(-1, "b2t_graph_launch: null graph");
    CK(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
    g->ctx->launches += g->kernels;
    return 0;
}

extern "C" void b2t_graph_destroy(b2t_graph *g) {
    if (!g) return;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
}

// ------------------------------------------------------------------------------------------------ weight broadcast
// The one collective of the path: the packed blob (backbone + ConvLSTM head) from `root` to every
// rank of an NCCL communicator, over NVLink / NVSwitch.  NCCL is resolved at run time (dlopen of the library the host
// process already uses), so the .so has no link-time dependency on it.  Ranks other than root then call
// b2t_finalize(ctx, upload = 0).
#include <dlfcn.h>
extern "C" int b2t_broadcast_weights(b2t_ctx *c, void *nccl_comm, int root, void *stream) {
    if (!c || !nccl_comm) return fail(-1, "b2t_broadcast_weights: null argument");
    if (!c->d_blob) return fail(-1, "b2t_broadcast_weights: no device blob yet (b2t_bind_memory or b2t_finalize first)");
    typedef int (*bcast_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
    static bcast_fn fn = nullptr;
    if (!fn) {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return fail(-2, "b2t_broadcast_weights: libnccl.so.2 not found (%s)", dlerror());
        fn = (bcast_fn)dlsym(h, "ncclBroadcast");
        if (!fn) return fail(-2, "b2t_broadcast_weights: ncclBroadcast not found in libnccl");
    }
    const int rc = fn(c->d_blob, c->d_blob, c->weight_bytes, 1 /* ncclUint8 */, root, nccl_comm, (cudaStream_t)stream);
    if (rc) return fail(-2, "ncclBroadcast failed with ncclResult %d", rc);
    return 0;
}

// ------------------------------------------------------------------------------------------------ ConvLSTM
static int convlstm_reset_slots(b2t_ctx *c, int slot0, int n, cudaStream_t st) {
    const ActBuf &h = c->bufs[c->buf_hrec];
    const size_t per = (size_t)c->G * c->G * c->cfg.convlstm_units;
    CK(cudaMemsetAsync(h.hi + (size_t)slot0 * per, 0, n * per * 2, st));
    CK(cudaMemsetAsync(h.hi + h.plane + (size_t)slot0 * per, 0, n * per * 2, st));
    CK(cudaMemsetAsync(c->d_ws + c->off_cstate + (size_t)slot0 * per * 4, 0, n * per * 4, st));
    return 0;
}

extern "C" int b2t_convlstm_reset(b2t_ctx *c, void *stream) {
    if (!c || !c->L_CIN || !c->finalized) return fail(-1, "convlstm not configured");
    return convlstm_reset_slots(c, 0, c->cfg.max_batch, (cudaStream_t)stream);
}

extern "C" int b2t_convlstm_reset_slots(b2t_ctx *c, int slot0, int n_slots, void *stream) {
    if (!c || !c->L_CIN || !c->finalized) return fail(-1, "convlstm not configured");
    if (slot0 < 0 || n_slots < 1 || slot0 + n_slots > c->cfg.max_batch) return fail(-1, "b2t_convlstm_reset_slots: slots [%d,%d) outside [0,%d)", slot0, slot0 + n_slots, c->cfg.max_batch);
    return convlstm_reset_slots(c, slot0, n_slots, (cudaStream_t)stream);
}

// S streams x T consecutive frames of the last b2t_yolo_forward (frame index s*T + t).  The input convolution W*z and
// the 1x1 head are one batched launch each over all S*T frames; only U*h + the gate maths are sequential, batched over
// the S streams (pixels of all streams share one pass over the recurrent weights).
extern "C" int b2t_convlstm_sequence(b2t_ctx *c, int S, int T, int slot0, int reset, float *trk_logits, int hard_sigmoid,
                                     void *stream) {
    if (!c || !c->L_CIN || !c->finalized) return fail(-1, "convlstm not configured");
    if (S < 1 || T < 1 || S * T > c->cfg.max_batch || !trk_logits) return fail(-1, "b2t_convlstm_sequence: S=%d x T=%d frames outside [1, max_batch=%d]", S, T, c->cfg.max_batch);
    if (slot0 < 0 || slot0 + S > c->cfg.max_batch) return fail(-1, "b2t_convlstm_sequence: state slots [%d,%d) outside [0,%d)", slot0, slot0 + S, c->cfg.max_batch);
    cudaStream_t st = (cudaStream_t)stream;
    const int u = c->cfg.convlstm_units, M = c->G * c->G;
    float *gates = reinterpret_cast<float *>(c->d_ws + c->off_gates);
    int rc;
    if (reset && (rc = convlstm_reset_slots(c, slot0, S, st))) return rc;
    // input convolution for all S*T frames at once (does not depend on h)
    if ((rc = run_conv(c, c->conv[c->L_CIN], S * T, gates, st))) return rc;
    for (int t = 0; t < T; ++t) {
        // += U * h_{t-1}: frame s*T + t of the gate buffer <- state slot slot0 + s.  After a reset h_{-1} = 0 and the
        // term vanishes exactly: the launch is skipped.
        if (!(reset && t == 0))
            if ((rc = run_conv(c, c->conv[c->L_CREC], S, gates + (size_t)t * M * 4 * u, st, (long long)T * M, slot0))) return rc;
        ConvLstmGateParams p;
        memset(&p, 0, sizeof p);
        p.g = gates; p.c = reinterpret_cast<float *>(c->d_ws + c->off_cstate);
        p.h_rec = dest_planes(c, c->buf_hrec, 0, c->G, c->G, DEST_PLAIN);
        p.h_seq = dest_planes(c, c->buf_hseq, 0, c->G, c->G, DEST_PLAIN);
        p.M = M; p.units = u; p.G = c->G; p.S = S; p.T = T; p.t = t; p.slot0 = slot0; p.hard_sigmoid = hard_sigmoid;
        if ((rc = launch_convlstm_gates(p, st))) return fail(-2, "gates launch: %s", cudaGetErrorString((cudaError_t)rc));
        c->launches += 1;
    }
    return run_conv(c, c->conv[c->L_HEAD], S * T, trk_logits, st);
}

// frames [0,batch) of the last forward = consecutive time steps of ONE stream (state slot 0), state carried over
extern "C" int b2t_convlstm_window(b2t_ctx *c, int B, float *trk_logits, int hard_sigmoid, void *stream) {
    return b2t_convlstm_sequence(c, 1, B, 0, 0, trk_logits, hard_sigmoid, stream);
}
