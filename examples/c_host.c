/*
 * c_host.c -- a plain C99 host of libb200track.so: no Python, no torch.  It drives the same path the plugin classes
 * drive (the reference's YOLO.detect flow, models_detection/YOLO.py:140-162, in batch form):
 *
 *   b2t_create -> cudaMalloc + b2t_bind_memory -> b2t_load_darknet_weights -> b2t_finalize
 *   frames (B,416,416,3) uint8 -> [CUDA graph: b2t_ingest_frames is outside, b2t_yolo_forward + b2t_region_detect +
 *   b2t_pool_features inside] -> detections + pooled feature on the host
 *
 * usage: c_host <darknet .weights> <raw uint8 frames file> <n_frames> <n_class>
 * output (stdout, one line per frame): "frame i: n detections; first: cx cy w h obj prob class; logit[0] fv[0]"
 * Build: gcc -std=c99 -Iinclude -I/usr/local/cuda/include examples/c_host.c -o c_host \
 *            -L object-tracking_b200 -lb200track -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,...
 * tests/test_c_host.py compiles it (CPU suite) and compares its output with the Python plugin path (GPU suite).
 */
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime_api.h>

#include "b200track.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        if ((call) < 0) {                                                        \
            fprintf(stderr, "%s failed: %s\n", #call, b2t_last_error());         \
            return 1;                                                            \
        }                                                                        \
    } while (0)
#define CUDA(call)                                                               \
    do {                                                                         \
        cudaError_t e_ = (call);                                                 \
        if (e_ != cudaSuccess) {                                                 \
            fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));          \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: %s weights frames.u8 n_frames n_class\n", argv[0]);
        return 2;
    }
    const int B = atoi(argv[3]), C = atoi(argv[4]), S = 416, G = S / 32, A = 5, MAXD = 64;
    static const float anchors[10] = {0.57273f, 0.677385f, 1.87446f, 2.06253f, 3.33843f, 5.47434f, 7.88282f, 3.52778f,
                                      9.77052f, 9.16828f};          /* KerasYOLO.py:45 / cfg/yolov2.cfg */
    b2t_config cfg = {0};
    cfg.image_h = cfg.image_w = S;
    cfg.n_class = C;
    cfg.max_batch = B;
    cfg.semantics = B2T_SEM_DARKNET;
    cfg.engine = B2T_ENGINE_TCGEN05;
    cfg.bn_eps = 1e-3f;
    b2t_ctx *ctx = NULL;
    CHECK(b2t_create(&cfg, &ctx));
    void *blob = NULL, *ws = NULL;
    CUDA(cudaMalloc(&blob, b2t_weight_bytes(ctx)));
    CUDA(cudaMalloc(&ws, b2t_workspace_bytes(ctx)));
    CHECK(b2t_bind_memory(ctx, blob, ws));
    CHECK(b2t_load_darknet_weights(ctx, argv[1]));
    cudaStream_t st;
    CUDA(cudaStreamCreate(&st));
    CHECK(b2t_finalize(ctx, 1, st));

    const size_t fbytes = (size_t)B * S * S * 3;
    unsigned char *h_frames = (unsigned char *)malloc(fbytes), *d_frames = NULL;
    FILE *f = fopen(argv[2], "rb");
    if (!f || fread(h_frames, 1, fbytes, f) != fbytes) {
        fprintf(stderr, "cannot read %zu bytes of frames from %s\n", fbytes, argv[2]);
        return 1;
    }
    fclose(f);
    CUDA(cudaMalloc((void **)&d_frames, fbytes));
    float *d_dets = NULL, *d_fv = NULL;
    int *d_counts = NULL;
    CUDA(cudaMalloc((void **)&d_dets, (size_t)B * MAXD * 8 * sizeof(float)));
    CUDA(cudaMalloc((void **)&d_counts, B * sizeof(int)));
    CUDA(cudaMalloc((void **)&d_fv, (size_t)B * 1024 * sizeof(float)));

    /* one step = ingest (outside the graph: its source pointer may change per step) + a captured graph */
    CUDA(cudaMemcpyAsync(d_frames, h_frames, fbytes, cudaMemcpyHostToDevice, st));
    CHECK(b2t_ingest_frames(ctx, d_frames, 1, B, 0, st));
    CHECK(b2t_yolo_forward(ctx, NULL, B2T_FRAME_U8, B, NULL, st));            /* warm-up outside the capture */
    CUDA(cudaStreamSynchronize(st));
    b2t_graph *g = NULL;
    CHECK(b2t_graph_begin(ctx, st));
    CHECK(b2t_yolo_forward(ctx, NULL, B2T_FRAME_U8, B, NULL, st));
    CHECK(b2t_region_detect(ctx, b2t_logits(ctx), B, G, G, A, C, 0.5f, 0.45f, anchors, S, S, S, S, d_dets, d_counts, MAXD, st));
    CHECK(b2t_pool_features(ctx, "norm_20", B, 0, 0, d_fv, st));               /* fv_layer 25 = conv_20's output, global max */
    CHECK(b2t_graph_end(ctx, st, &g));
    for (int rep = 0; rep < 2; ++rep) {                                          /* replay twice: same result */
        CHECK(b2t_ingest_frames(ctx, d_frames, 1, B, 0, st));
        CHECK(b2t_graph_launch(g, st));
    }
    float *dets = (float *)malloc((size_t)B * MAXD * 8 * sizeof(float)), *fv = (float *)malloc((size_t)B * 1024 * sizeof(float));
    float *logits = (float *)malloc((size_t)B * G * G * A * (5 + C) * sizeof(float));
    int *counts = (int *)malloc(B * sizeof(int));
    CUDA(cudaMemcpyAsync(dets, d_dets, (size_t)B * MAXD * 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA(cudaMemcpyAsync(counts, d_counts, B * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA(cudaMemcpyAsync(fv, d_fv, (size_t)B * 1024 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA(cudaMemcpyAsync(logits, b2t_logits(ctx), (size_t)B * G * G * A * (5 + C) * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < B; ++i) {
        const float *d = dets + (size_t)i * MAXD * 8;
        printf("frame %d: %d detections; first: %.9g %.9g %.9g %.9g %.9g %.9g %d; logit %.9g fv %.9g\n", i, counts[i],
               counts[i] > 0 ? d[0] : 0.f, counts[i] > 0 ? d[1] : 0.f, counts[i] > 0 ? d[2] : 0.f, counts[i] > 0 ? d[3] : 0.f,
               counts[i] > 0 ? d[4] : 0.f, counts[i] > 0 ? d[5] : 0.f, counts[i] > 0 ? (int)d[6] : -1,
               logits[(size_t)i * G * G * A * (5 + C)], fv[(size_t)i * 1024]);
    }
    fprintf(stderr, "kernels launched by the context: %ld\n", b2t_launch_count(ctx));
    b2t_graph_destroy(g);
    b2t_destroy(ctx);
    cudaFree(blob); cudaFree(ws); cudaFree(d_frames); cudaFree(d_dets); cudaFree(d_counts); cudaFree(d_fv);
    free(h_frames); free(dets); free(fv); free(logits); free(counts);
    return 0;
}
