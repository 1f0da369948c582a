// The reference's existing C-ABI (include/darknet_compat.h): libdarknet.so's symbols as models_detection/YOLO.py
// binds them, implemented on top of the b2t_* context.  Host code here only parses cfg/.data files and marshals
// arrays; forward pass, letterbox, region layer, box decode and NMS are device kernels.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200track.h"
#include "../../include/darknet_compat.h"

int b2t_fail_internal(int code, const char *msg);
namespace b2t {
bool jpeg_decode_rgb(const uint8_t *data, size_t size, std::vector<uint8_t> &rgb, int &width, int &height, std::string &error);
}

namespace {

int g_device = 0;

struct Net {
    b2t_ctx *ctx = nullptr;
    int w = 0, h = 0, classes = 0, n_box = 5, grid = 0;
    float anchors[32] = {0};
    float *d_in = nullptr;        // (net_h, net_w, 3) float, what conv_1 reads
    float *d_src = nullptr;       // caller's CHW image on the device
    size_t src_cap = 0;
    float *d_region = nullptr;    // (A*D, G, G) region-layer output
    float *d_dets = nullptr;      // (max, 8)
    int *d_count = nullptr;
    float *d_feat = nullptr;
    size_t feat_cap = 0;
    std::vector<float> h_region, h_feat, h_tmp;
    float thresh = 0.5f;
    int det_w = 0, det_h = 0;
    bool tiny = false;            // cfg/yolov2-tiny*.cfg: 9 conv layers, region layer = darknet layer 15 (31 for yolov2)
    int region_idx = 31;
};

std::mutex g_mu;
std::map<detection *, Net *> g_owner;

int err(const std::string &m) { return b2t_fail_internal(-1, m.c_str()); }

// ---------------------------------------------------------------- tiny device kernels
__global__ void chw_to_hwc_kernel(const float *src, int w, int h, int c, float *dst) {
    const int n = w * h * c;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int k = i % c, x = (i / c) % w, y = i / (c * w);
        dst[i] = src[(k * h + y) * w + x];
    }
}
__global__ void hwc_to_chw_kernel(const float *src, int w, int h, int c, float *dst) {
    const int n = w * h * c;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int x = i % w, y = (i / w) % h, k = i / (w * h);
        dst[i] = src[(y * w + x) * c + k];
    }
}
// letterbox_image (image.c:960-979) = resize_image (:1347-1389, separable bilinear with darknet's (n-1)/(m-1)
// scale and its operation order) embedded in a 0.5-filled net-sized canvas; CHW in, HWC out.
__global__ void letterbox_kernel(const float *src, int sw, int sh, int c, int nw, int nh, int new_w, int new_h,
                                 float *dst) {
    const int n = nw * nh * c;
    const int ox = (nw - new_w) / 2, oy = (nh - new_h) / 2;
    const float w_scale = (float)(sw - 1) / (new_w - 1), h_scale = (float)(sh - 1) / (new_h - 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int k = i % c, x = (i / c) % nw, y = i / (c * nw);
        const int rx = x - ox, ry = y - oy;
        float v = 0.5f;
        if (rx >= 0 && rx < new_w && ry >= 0 && ry < new_h) {
            const float *pl = src + (size_t)k * sh * sw;
            auto hrow = [&](int r) {           // horizontally resized pixel (rx, r)
                if (rx == new_w - 1 || sw == 1) return pl[r * sw + sw - 1];
                const float sx = rx * w_scale;
                const int ix = (int)sx;
                const float dx = sx - ix;
                return __fadd_rn(__fmul_rn(1.f - dx, pl[r * sw + ix]), __fmul_rn(dx, pl[r * sw + ix + 1]));
            };
            const float sy = ry * h_scale;
            const int iy = (int)sy;
            const float dy = sy - iy;
            v = __fmul_rn(1.f - dy, hrow(iy));
            if (!(ry == new_h - 1 || sh == 1)) v = __fadd_rn(v, __fmul_rn(dy, hrow(iy + 1)));
        }
        dst[i] = v;
    }
}
// forward_region_layer (region_layer.c:158-185, softmax=1): logistic on x,y,objectness, per-anchor softmax;
// input NHWC logits (G,G,A,D), output darknet's CHW (A*D,G,G).
__global__ void region_activate_kernel(const float *logits, int G, int A, int C, float *out) {
    const int D = 5 + C, cells = G * G;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells * A) return;
    const int a = i % A, cell = i / A;
    const float *r = logits + ((size_t)cell * A + a) * D;
    float *o = out + (size_t)a * D * cells + cell;
    o[0] = (float)(1. / (1. + exp(-(double)r[0])));
    o[cells] = (float)(1. / (1. + exp(-(double)r[1])));
    o[2 * cells] = r[2];
    o[3 * cells] = r[3];
    o[4 * cells] = (float)(1. / (1. + exp(-(double)r[4])));
    float mx = -INFINITY, sum = 0.f;
    for (int k = 0; k < C; ++k) mx = fmaxf(mx, r[5 + k]);
    for (int k = 0; k < C; ++k) {
        const float e = (float)exp((double)(r[5 + k] - mx));
        sum += e;
        o[(size_t)(5 + k) * cells] = e;
    }
    for (int k = 0; k < C; ++k) o[(size_t)(5 + k) * cells] /= sum;
}

// ---------------------------------------------------------------- cfg / data parsing (host)
struct Section { std::string name; std::map<std::string, std::string> kv; };

bool read_sections(const char *path, std::vector<Section> &out) {
    FILE *f = fopen(path, "r");
    if (!f) return false;
    char line[4096];
    while (fgets(line, sizeof line, f)) {
        std::string s(line);
        size_t a = s.find_first_not_of(" \t\r\n");
        if (a == std::string::npos) continue;
        size_t b = s.find_last_not_of(" \t\r\n");
        s = s.substr(a, b - a + 1);
        if (s[0] == '#' || s[0] == ';') continue;
        if (s[0] == '[') { out.push_back(Section{s, {}}); continue; }
        const size_t eq = s.find('=');
        if (eq == std::string::npos) continue;
        auto trim = [](std::string t) {
            const size_t x = t.find_first_not_of(" \t"), y = t.find_last_not_of(" \t");
            return x == std::string::npos ? std::string() : t.substr(x, y - x + 1);
        };
        if (out.empty()) out.push_back(Section{"", {}});
        out.back().kv[trim(s.substr(0, eq))] = trim(s.substr(eq + 1));
    }
    fclose(f);
    return true;
}
int geti(const Section &s, const char *k, int d) {
    auto it = s.kv.find(k);
    return it == s.kv.end() ? d : atoi(it->second.c_str());
}

const int kFilters[22] = {32, 64, 128, 64, 128, 256, 128, 256, 512, 256, 512, 256, 512, 1024, 512, 1024, 512, 1024, 1024, 1024, 64, 1024};
const int kSizes[22] = {3, 3, 3, 1, 3, 3, 1, 3, 3, 1, 3, 1, 3, 3, 1, 3, 1, 3, 3, 3, 1, 3};

// cfg/yolov2-tiny-voc.cfg / yolov2-tiny.cfg: darknet layer index -> engine tensor
const char *tiny_layer_name(int idx) {
    static const char *names[15] = {"norm_1", "pool_1", "norm_2", "pool_2", "norm_3", "pool_3", "norm_4", "pool_4",
                                    "norm_5", "pool_5", "norm_6", "pool_6", "norm_7", "norm_8", "conv_9"};
    return (idx >= 0 && idx < 15) ? names[idx] : nullptr;
}

const char *layer_name(int idx) {      // darknet layer index -> engine tensor
    static const char *names[31] = {"norm_1", "pool_1", "norm_2", "pool_2", "norm_3", "norm_4", "norm_5", "pool_5",
                                    "norm_6", "norm_7", "norm_8", "pool_8", "norm_9", "norm_10", "norm_11", "norm_12",
                                    "norm_13", "pool_13", "norm_14", "norm_15", "norm_16", "norm_17", "norm_18",
                                    "norm_19", "norm_20", "norm_13", nullptr, nullptr, "concat", "norm_22", "conv_23"};
    return (idx >= 0 && idx < 31) ? names[idx] : nullptr;
}

int run_forward(Net *n) {
    if (b2t_yolo_forward(n->ctx, n->d_in, B2T_FRAME_F32, 1, nullptr, nullptr)) return -1;
    const int cells = n->grid * n->grid;
    region_activate_kernel<<<(cells * n->n_box + 127) / 128, 128>>>(b2t_logits(n->ctx), n->grid, n->n_box, n->classes,
                                                                    n->d_region);
    n->h_region.resize((size_t)cells * n->n_box * (5 + n->classes));
    if (cudaMemcpy(n->h_region.data(), n->d_region, n->h_region.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        return err("network_predict: device error: " + std::string(cudaGetErrorString(cudaGetLastError())));
    return 0;
}

}  // namespace

// ---------------------------------------------------------------- exported darknet symbols
extern "C" {

void cuda_set_device(int n) { g_device = n; cudaSetDevice(n); }

network *load_network(char *cfg, char *weights, int clear) {
    (void)clear;
    std::vector<Section> secs;
    if (!cfg || !read_sections(cfg, secs)) { err(std::string("load_network: cannot read cfg ") + (cfg ? cfg : "(null)")); return nullptr; }
    Net *n = new Net();
    int conv = 0, bad = 0, bad_tiny = 0, pools = 0, pool_s1 = 0, tiny_f8 = 0;
    static const int kTinyFilters[7] = {16, 32, 64, 128, 256, 512, 1024};
    for (const Section &s : secs) {
        if (s.name == "[net]" || s.name == "[network]") { n->w = geti(s, "width", 416); n->h = geti(s, "height", 416); }
        if (s.name == "[convolutional]") {
            const int f = geti(s, "filters", 1), k = geti(s, "size", 1);
            if (conv < 22 && (f != kFilters[conv] || k != kSizes[conv])) bad = 1;
            if (conv < 7 && (f != kTinyFilters[conv] || k != 3)) bad_tiny = 1;
            if (conv == 7) { tiny_f8 = f; if (k != 3) bad_tiny = 1; }
            if (conv == 8 && k != 1) bad_tiny = 1;
            ++conv;
        }
        if (s.name == "[maxpool]") {
            ++pools;
            if (geti(s, "stride", 1) == 1) { ++pool_s1; if (pools != 6 || geti(s, "size", 2) != 2) bad_tiny = 1; }
        }
        if (s.name == "[region]") {
            n->classes = geti(s, "classes", 20);
            n->n_box = geti(s, "num", 5);
            auto it = s.kv.find("anchors");
            if (it != s.kv.end()) {
                int i = 0;
                const char *p = it->second.c_str();
                while (*p && i < 32) {
                    n->anchors[i++] = (float)atof(p);
                    p = strchr(p, ',');
                    if (!p) break;
                    ++p;
                }
            }
        }
    }
    n->tiny = conv == 9 && !bad_tiny && pools == 6 && pool_s1 == 1;
    if ((!(conv == 23 && !bad) && !n->tiny) || n->n_box != 5 || n->classes < 1 || n->w != n->h || n->w % 32) {
        err("load_network: supported graphs are cfg/yolov2.cfg / yolov2-voc.cfg (23 conv layers) and cfg/yolov2-tiny.cfg / "
            "yolov2-tiny-voc.cfg (9 conv layers, sixth maxpool stride 1); 5 anchors, square input multiple of 32");
        delete n;
        return nullptr;
    }
    n->grid = n->w / 32;
    n->region_idx = n->tiny ? 15 : 31;
    b2t_config c;
    memset(&c, 0, sizeof c);
    if (n->tiny) { c.reserved[2] = 1; c.reserved[3] = tiny_f8; }
    c.image_h = n->h; c.image_w = n->w; c.n_class = n->classes; c.max_batch = 1;
    c.semantics = B2T_SEM_DARKNET; c.bn_eps = 1e-3f; c.engine = B2T_ENGINE_TCGEN05; c.device = g_device;
    if (b2t_create(&c, &n->ctx)) { delete n; return nullptr; }
    if (!weights || !weights[0]) { err("load_network: a weights file is required"); b2t_destroy(n->ctx); delete n; return nullptr; }
    if (b2t_load_darknet_weights(n->ctx, weights) || b2t_finalize(n->ctx, 1, nullptr)) { b2t_destroy(n->ctx); delete n; return nullptr; }
    const size_t cells = (size_t)n->grid * n->grid, nd = cells * n->n_box;
    if (cudaMalloc(&n->d_in, (size_t)n->w * n->h * 3 * 4) || cudaMalloc(&n->d_region, nd * (5 + n->classes) * 4) ||
        cudaMalloc(&n->d_dets, nd * 8 * 4) || cudaMalloc(&n->d_count, 4)) {
        err("load_network: cudaMalloc failed");
        free_network(reinterpret_cast<network *>(n));
        return nullptr;
    }
    cudaDeviceSynchronize();
    return reinterpret_cast<network *>(n);
}

void free_network(network *net) {
    Net *n = reinterpret_cast<Net *>(net);
    if (!n) return;
    cudaFree(n->d_in); cudaFree(n->d_src); cudaFree(n->d_region); cudaFree(n->d_dets); cudaFree(n->d_count); cudaFree(n->d_feat);
    if (n->ctx) b2t_destroy(n->ctx);
    delete n;
}

int network_width(network *net) { return net ? reinterpret_cast<Net *>(net)->w : 0; }
int network_height(network *net) { return net ? reinterpret_cast<Net *>(net)->h : 0; }

metadata get_metadata(char *file) {
    metadata m = {0, nullptr};
    std::vector<Section> secs;
    if (!file || !read_sections(file, secs) || secs.empty()) { err("get_metadata: cannot read data file"); return m; }
    const Section &s = secs[0];
    m.classes = geti(s, "classes", 2);
    auto it = s.kv.find("names");
    if (it == s.kv.end()) it = s.kv.find("labels");
    if (it == s.kv.end()) { fprintf(stderr, "No names or labels found\n"); return m; }
    FILE *f = fopen(it->second.c_str(), "r");
    if (!f) { err("get_metadata: cannot open names file " + it->second); return m; }
    std::vector<std::string> names;
    char line[1024];
    while (fgets(line, sizeof line, f)) {
        std::string t(line);
        while (!t.empty() && (t.back() == '\n' || t.back() == '\r')) t.pop_back();
        names.push_back(t);
    }
    fclose(f);
    m.names = (char **)calloc(names.size() + 1, sizeof(char *));
    for (size_t i = 0; i < names.size(); ++i) m.names[i] = strdup(names[i].c_str());
    return m;
}

image make_image(int w, int h, int c) {
    image im = {w, h, c, nullptr};
    im.data = (float *)calloc((size_t)w * h * c, sizeof(float));
    return im;
}

// image.c:1442-1482 load_image_color -> load_image_stb(filename, 3): decode, then CHW float RGB = byte / 255.
// JPEG (what YOLO.py:141 passes) is decoded by jpeg_decode.cu, bit-exactly like the reference's decoder; binary PPM (P6)
// is read directly; a (w, h) request resizes like load_image (resize_image, image.c:1347-1389).
// resize_image (image.c:1347-1389): separable bilinear with the (n-1)/(m-1) scale, horizontal pass first, in darknet's
// float operation order (the same arithmetic as the device letterbox kernels)
static image resize_image_host(image im, int w, int h) {
    image part = make_image(w, im.h, im.c), out = make_image(w, h, im.c);
    const float w_scale = (float)(im.w - 1) / (w - 1), h_scale = (float)(im.h - 1) / (h - 1);
    for (int k = 0; k < im.c; ++k)
        for (int r = 0; r < im.h; ++r) {
            const float *src = im.data + ((size_t)k * im.h + r) * im.w;
            float *dst = part.data + ((size_t)k * im.h + r) * w;
            for (int c = 0; c < w; ++c) {
                if (c == w - 1 || im.w == 1) { dst[c] = src[im.w - 1]; continue; }
                const float sx = c * w_scale;
                const int ix = (int)sx;
                const float dx = sx - ix;
                dst[c] = (1 - dx) * src[ix] + dx * src[ix + 1];
            }
        }
    for (int k = 0; k < im.c; ++k)
        for (int r = 0; r < h; ++r) {
            const float sy = r * h_scale;
            const int iy = (int)sy;
            const float dy = sy - iy;
            float *dst = out.data + ((size_t)k * h + r) * w;
            const float *p0 = part.data + ((size_t)k * im.h + iy) * w;
            for (int c = 0; c < w; ++c) dst[c] = (1 - dy) * p0[c];
            if (r == h - 1 || im.h == 1) continue;
            for (int c = 0; c < w; ++c) dst[c] += dy * p0[w + c];
        }
    free(part.data);
    return out;
}

image load_image_color(char *filename, int w, int h) {
    image im = {0, 0, 0, nullptr};
    FILE *f = filename ? fopen(filename, "rb") : nullptr;
    if (!f) { err(std::string("load_image_color: cannot open ") + (filename ? filename : "(null)")); return im; }
    std::vector<unsigned char> file;
    unsigned char chunk[65536];
    size_t got;
    while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) file.insert(file.end(), chunk, chunk + got);
    fclose(f);
    std::vector<unsigned char> buf;
    int iw = 0, ih = 0;
    if (file.size() >= 2 && file[0] == 0xFF && file[1] == 0xD8) {
        std::string e;
        if (!b2t::jpeg_decode_rgb(file.data(), file.size(), buf, iw, ih, e)) { err(std::string("load_image_color: ") + filename + ": " + e); return im; }
    } else if (file.size() >= 2 && file[0] == 'P' && file[1] == '6') {
        int maxv = 0, pos = 0;
        if (sscanf((const char *)file.data(), "P6 %d %d %d%n", &iw, &ih, &maxv, &pos) != 3 || maxv != 255 || iw < 1 || ih < 1 ||
            file.size() < (size_t)pos + 1 + (size_t)iw * ih * 3) {
            err("load_image_color: bad or truncated PPM (binary P6 with maxval 255 expected)");
            return im;
        }
        buf.assign(file.begin() + pos + 1, file.begin() + pos + 1 + (size_t)iw * ih * 3);
    } else {
        err(std::string("load_image_color: ") + filename + ": the compat layer decodes JPEG and binary PPM; "
            "decode other formats in the caller and use make_image()");
        return im;
    }
    im = make_image(iw, ih, 3);
    for (int k = 0; k < 3; ++k)           // CHW, RGB, (float)byte / 255.  (image.c load_image_stb)
        for (int y = 0; y < ih; ++y)
            for (int x = 0; x < iw; ++x)
                im.data[((size_t)k * ih + y) * iw + x] = (float)((double)buf[((size_t)y * iw + x) * 3 + k] / 255.);
    if (w && h && (w != iw || h != ih)) {
        image r = resize_image_host(im, w, h);
        free(im.data);
        im = r;
    }
    return im;
}

void rgbgr_image(image im) {
    if (!im.data || im.c < 3) return;
    const size_t n = (size_t)im.w * im.h;
    for (size_t i = 0; i < n; ++i) { const float t = im.data[i]; im.data[i] = im.data[i + 2 * n]; im.data[i + 2 * n] = t; }
}

void free_image(image m) { free(m.data); }

float *network_predict(network *net, float *input) {
    Net *n = reinterpret_cast<Net *>(net);
    if (!n || !input) { err("network_predict: null argument"); return nullptr; }
    const size_t bytes = (size_t)n->w * n->h * 3 * 4;
    if (n->src_cap < bytes) { cudaFree(n->d_src); if (cudaMalloc(&n->d_src, bytes)) { err("cudaMalloc failed"); return nullptr; } n->src_cap = bytes; }
    cudaMemcpy(n->d_src, input, bytes, cudaMemcpyHostToDevice);
    chw_to_hwc_kernel<<<296, 256>>>(n->d_src, n->w, n->h, 3, n->d_in);
    if (run_forward(n)) return nullptr;
    return n->h_region.data();
}

float *network_predict_image(network *net, image im) {
    Net *n = reinterpret_cast<Net *>(net);
    if (!n || !im.data || im.c != 3) { err("network_predict_image: bad image"); return nullptr; }
    const size_t bytes = (size_t)im.w * im.h * 3 * 4;
    if (n->src_cap < bytes) { cudaFree(n->d_src); if (cudaMalloc(&n->d_src, bytes)) { err("cudaMalloc failed"); return nullptr; } n->src_cap = bytes; }
    cudaMemcpy(n->d_src, im.data, bytes, cudaMemcpyHostToDevice);
    int new_w, new_h;
    if (((float)n->w / im.w) < ((float)n->h / im.h)) { new_w = n->w; new_h = (im.h * n->w) / im.w; }
    else { new_h = n->h; new_w = (im.w * n->h) / im.h; }
    letterbox_kernel<<<296, 256>>>(n->d_src, im.w, im.h, 3, n->w, n->h, new_w, new_h, n->d_in);
    if (run_forward(n)) return nullptr;
    return n->h_region.data();
}

// Fills the G*G*A detection array the way YOLO.py consumes it: entries with prob > 0 carry box / objectness / prob.
// nms <= 0 -> thresholded detections without suppression (get_network_boxes); nms > 0 -> after do_nms_obj, survivors
// first, ordered by objectness like darknet's qsort.
static int fill_dets(Net *n, detection *dets, int total, float nms, bool compact) {
    const int cells = n->grid * n->grid, nd = cells * n->n_box;
    if (b2t_region_detect(n->ctx, b2t_logits(n->ctx), 1, n->grid, n->grid, n->n_box, n->classes, n->thresh,
                          nms > 0 ? nms : 2.0f, n->anchors, n->det_w, n->det_h, n->w, n->h, n->d_dets, n->d_count, nd, nullptr))
        return -1;
    int cnt = 0;
    n->h_tmp.resize((size_t)nd * 8);
    cudaMemcpy(&cnt, n->d_count, 4, cudaMemcpyDeviceToHost);
    if (cnt < 0) return err("get_network_boxes: candidate overflow");
    cudaMemcpy(n->h_tmp.data(), n->d_dets, (size_t)cnt * 8 * 4, cudaMemcpyDeviceToHost);
    for (int i = 0; i < total; ++i) {
        dets[i].objectness = 0;
        dets[i].bbox = box{0, 0, 0, 0};
        memset(dets[i].prob, 0, sizeof(float) * n->classes);
    }
    // rows are (detection, class) pairs sorted by -prob; group them per detection
    std::map<int, int> slot_of;       // anchor index -> slot
    std::vector<std::pair<float, int>> order;   // (-objectness, anchor)
    for (int r = 0; r < cnt; ++r) {
        const float *row = &n->h_tmp[(size_t)r * 8];
        const int anchor = (int)row[7];
        if (!slot_of.count(anchor)) { slot_of[anchor] = 0; order.push_back({-row[4], anchor}); }
    }
    if (compact) {
        std::stable_sort(order.begin(), order.end());
        for (size_t i = 0; i < order.size(); ++i) slot_of[order[i].second] = (int)i;
    } else {
        for (auto &kv : slot_of) kv.second = kv.first;
    }
    for (int r = 0; r < cnt; ++r) {
        const float *row = &n->h_tmp[(size_t)r * 8];
        const int slot = slot_of[(int)row[7]];
        if (slot >= total) continue;
        dets[slot].bbox = box{row[0], row[1], row[2], row[3]};
        dets[slot].objectness = row[4];
        dets[slot].prob[(int)row[6]] = row[5];
    }
    return 0;
}

detection *get_network_boxes(network *net, int w, int h, float thresh, float hier, int *map, int relative, int *num) {
    (void)hier; (void)map;
    Net *n = reinterpret_cast<Net *>(net);
    if (!n) { err("get_network_boxes: null network"); return nullptr; }
    if (relative) { w = 1; h = 1; }                 // correct_region_boxes: relative boxes are not scaled to pixels
    const int nd = n->grid * n->grid * n->n_box;
    detection *dets = (detection *)calloc(nd, sizeof(detection));
    for (int i = 0; i < nd; ++i) {
        dets[i].classes = n->classes;
        dets[i].prob = (float *)calloc(n->classes, sizeof(float));
        dets[i].mask = nullptr;
    }
    n->thresh = thresh; n->det_w = w; n->det_h = h;
    if (num) *num = nd;
    { std::lock_guard<std::mutex> g(g_mu); g_owner[dets] = n; }
    fill_dets(n, dets, nd, 0.f, false);
    return dets;
}

void do_nms_obj(detection *dets, int total, int classes, float thresh) {
    (void)classes;
    Net *n = nullptr;
    { std::lock_guard<std::mutex> g(g_mu); auto it = g_owner.find(dets); if (it != g_owner.end()) n = it->second; }
    if (!n) { err("do_nms_obj: detections were not produced by get_network_boxes of this library"); return; }
    fill_dets(n, dets, total, thresh, true);
}

void free_detections(detection *dets, int nd) {
    if (!dets) return;
    { std::lock_guard<std::mutex> g(g_mu); g_owner.erase(dets); }
    for (int i = 0; i < nd; ++i) { free(dets[i].prob); free(dets[i].mask); }
    free(dets);
}

void free_ptrs(void **ptrs, int n) {
    if (!ptrs) return;
    for (int i = 0; i < n; ++i) free(ptrs[i]);
    free(ptrs);
}

dims layer_dims(network *net, int idx) {
    dims d = {0, 0, 0};
    Net *n = reinterpret_cast<Net *>(net);
    if (!n) return d;
    const char *name = n->tiny ? tiny_layer_name(idx - 1) : layer_name(idx - 1);
    if (idx - 1 == n->region_idx || idx - 1 == n->region_idx - 1) { d.w = d.h = n->grid; d.c = n->n_box * (5 + n->classes); return d; }
    if (!name) { err("layer_dims: layer is not kept by the B200 engine"); return d; }
    b2t_layer_dims(n->ctx, name, &d.h, &d.w, &d.c);
    return d;
}

feature network_extract_feat(network *net, int idx) {
    feature f = {0, nullptr};
    Net *n = reinterpret_cast<Net *>(net);
    if (!n) return f;
    if (idx - 1 == n->region_idx) {    // the region layer's output
        f.size = (int)n->h_region.size();
        f.feat = n->h_region.data();
        return f;
    }
    const char *name = n->tiny ? tiny_layer_name(idx - 1) : layer_name(idx - 1);
    int h = 0, w = 0, c = 0;
    if (!name || b2t_layer_dims(n->ctx, name, &h, &w, &c)) { err("network_extract_feat: layer is not kept by the B200 engine"); return f; }
    const size_t cnt = (size_t)h * w * c;
    if (n->feat_cap < 2 * cnt * 4) { cudaFree(n->d_feat); if (cudaMalloc(&n->d_feat, 2 * cnt * 4)) { err("cudaMalloc failed"); return f; } n->feat_cap = 2 * cnt * 4; }
    if (b2t_extract(n->ctx, name, 1, n->d_feat, nullptr) < 0) return f;
    hwc_to_chw_kernel<<<296, 256>>>(n->d_feat, w, h, c, n->d_feat + cnt);
    n->h_feat.resize(cnt);
    cudaMemcpy(n->h_feat.data(), n->d_feat + cnt, cnt * 4, cudaMemcpyDeviceToHost);
    f.size = (int)cnt;
    f.feat = n->h_feat.data();
    return f;
}

}  // extern "C"
