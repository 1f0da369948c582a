/*
 * b200track.h -- C-ABI of libb200track.so: the per-frame detect-and-track hot path of
 * ktzsh/object-tracking on one B200 (sm_100a).
 *
 * Plain C: opaque handle, int status (0 = ok, <0 = error; text in b2t_last_error()), plain
 * pointers and sizes, no torch types.  The library never calls exit().  Device pointers are
 * raw CUDA addresses (e.g. torch.Tensor.data_ptr()); "stream" is a cudaStream_t passed as void*
 * (NULL = the legacy default stream).  One context per GPU; a context is not thread-safe.
 *
 * The reference has no FFI for its Keras path (it calls into a TensorFlow session); every
 * entry point below names the reference code it replaces.  The reference's *existing* FFI
 * (models_detection/YOLO.py:58-119 -> libdarknet.so) is declared in darknet_compat.h.
 */
#ifndef B200TRACK_H
#define B200TRACK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2t_ctx b2t_ctx;

/* semantics of BatchNorm / space-to-depth, which differ between the two reference detectors */
enum { B2T_SEM_KERAS = 0,     /* gamma*(x-mean)/sqrt(var+eps)+beta, tf.space_to_depth  (KerasYOLO.py:257-262,241) */
       B2T_SEM_DARKNET = 1 }; /* (x-mean)/(sqrt(var)+1e-6)*gamma+beta, reorg_cpu        (blas.c:147-158, :9-30)    */

/* which contraction kernel runs the convolutions.  The release library has ONE: B2T_ENGINE_TCGEN05.  The other two
 * values select developer cross-check engines that exist only in `make DEV=1` builds (libb200track_dev.so,
 * b2t_dev_build() == 1); b2t_create rejects them otherwise. */
enum { B2T_ENGINE_TCGEN05 = 0,   /* tcgen05.mma, fp16 hi/lo split operands, 3 MMAs per product, fp32 accumulate in TMEM;
                                    activation patch stationary in shared memory (conv_halo.cu / conv_pm.cu)          */
       B2T_ENGINE_SIMT = 1,      /* DEV: fp32 FMA on the same operands (slow)                                          */
       B2T_ENGINE_TCGEN05_TILE = 2 }; /* DEV: first-generation tcgen05 kernel, one TMA box per tap (dev_engines.cu)   */

enum { B2T_FRAME_U8 = 0,      /* HWC uint8, divided by 255 on load (utils.py:150-153 normalize) */
       B2T_FRAME_F32 = 1 };   /* HWC float32, already normalised                                 */

typedef struct {
    int image_h, image_w;     /* KerasYOLO.IMAGE_H/W (416) ; multiples of 32                     */
    int n_class;              /* KerasYOLO.CLASS                                                 */
    int max_batch;            /* frames per b2t_yolo_forward call                                */
    int semantics;            /* B2T_SEM_*                                                       */
    float bn_eps;             /* keras epsilon (1e-3); ignored for B2T_SEM_DARKNET               */
    int engine;               /* B2T_ENGINE_*                                                    */
    int device;               /* CUDA device ordinal                                             */
    int convlstm_units;       /* 0 = no MultiObjDetTracker head; else ConvLSTM2D filters (512)   */
    int reserved[7];          /* reserved[0] = 1: also keep the pre-pool outputs of conv_1,2,5,8 (KerasYOLO.extract)
                                 reserved[1]: largest batch that runs conv_2..23 as ONE persistent cooperative launch
                                 (conv_chain_kernel, the batch-1 schedule); 0 = default (1: measured faster only
                                 for a single frame, profiles/r2_batch_sweep_416.txt), -1 = never
                                 reserved[2] = 1: the tiny graph of cfg/yolov2-tiny*.cfg (9 conv layers, sixth maxpool
                                 stride 1) instead of cfg/yolov2.cfg's; reserved[3] = filters of its conv_8 (1024 voc,
                                 512 coco; 0 = 1024).  Weights: b2t_load_darknet_weights or b2t_set_conv_weights(1..9). */
} b2t_config;

const char *b2t_last_error(void);
int  b2t_version(void);
int  b2t_dev_build(void);     /* 1 = built with -DB2T_DEV (cross-check engines, B2T_* environment overrides, trace stamps) */

/* ---- lifetime ---------------------------------------------------------------------------- */
int  b2t_create(const b2t_config *cfg, b2t_ctx **out);
void b2t_destroy(b2t_ctx *ctx);

/* Device memory is the caller's (a torch tensor): ask for the sizes, allocate, bind.  If
 * b2t_bind_memory is never called the context cudaMalloc's its own at b2t_finalize. */
size_t b2t_weight_bytes(const b2t_ctx *ctx);      /* packed weight blob (what a weight broadcast moves)      */
size_t b2t_workspace_bytes(const b2t_ctx *ctx);   /* activations + split-K partials + decode scratch          */
int    b2t_bind_memory(b2t_ctx *ctx, void *weight_blob_dev, void *workspace_dev);

/* ---- weights: replaces KerasYOLO.init_weights (KerasYOLO.py:244-274) + WeightReader (utils.py:138-148) ---- */
/* conv_index 1..23.  kernel: Keras layout (kh,kw,Cin,Cout) float32.  BN layers pass gamma/beta/mean/var,
 * conv_23 passes bias (others NULL).  Host pointers; packed into a host staging copy of the blob. */
int  b2t_set_conv_weights(b2t_ctx *ctx, int conv_index, const float *kernel_hwio,
                          const float *gamma, const float *beta, const float *mean, const float *var,
                          const float *bias);
/* darknet .weights file (darknet/src/parser.c:1149-1230; v0.1 int32 and v0.2 size_t "seen" headers) */
int  b2t_load_darknet_weights(b2t_ctx *ctx, const char *path);
/* ConvLSTM2D + 1x1 head of MultiObjDetTracker (MultiObjDetTracker.py:176-183), Keras layouts:
 * kernel (3,3,Cz,4u) with Cz = 5*(5+C)+1024, recurrent (3,3,u,4u), bias (4u), head (1,1,u,5*(5+C)), head_bias */
int  b2t_set_convlstm_weights(b2t_ctx *ctx, const float *kernel, const float *recurrent, const float *bias,
                              const float *head_kernel, const float *head_bias);
/* upload the staged blob to the device, build TMA descriptors.  Call once after all weights are set
 * (or after a broadcast wrote the device blob: pass upload=0 to keep the device copy). */
int  b2t_finalize(b2t_ctx *ctx, int upload, void *stream);

/* ---- detector forward: replaces model.predict in KerasYOLO.predict (KerasYOLO.py:531) and
 *      forward_network in darknet (network.c:198-221) ---------------------------------------- */
/* frames_dev: (B,H,W,3) uint8 or float32 on the device.  logits_dev: (B,G,G,5,5+C) float32 or NULL
 * (the context keeps its own copy, see b2t_logits).  Asynchronous on `stream`. */
int  b2t_yolo_forward(b2t_ctx *ctx, const void *frames_dev, int frame_dtype, int batch,
                      float *logits_dev, void *stream);
/* conv_first .. conv_last (1..23) of the same pass; frames_dev is read only when conv_first == 1, logits_dev is
 * written only when conv_last == 23.  For host pipelines that overlap the tracker tail of step i with the first
 * layers of step i+1 (BaseTracker.track_windows(pipeline=True)). */
int  b2t_yolo_forward_range(b2t_ctx *ctx, const void *frames_dev, int frame_dtype, int batch, int conv_first,
                            int conv_last, float *logits_dev, void *stream);
/* Frame ingest without a staging copy: the step's uint8 frames are n_seg segments of seg_frames consecutive frames,
 * seg_stride_bytes apart (S streams x T frames of longer device-resident clips).  Fills the context's input buffer; a
 * following b2t_yolo_forward[_range](frames_dev = NULL, B2T_FRAME_U8, batch = n_seg * seg_frames) starts at conv_1, so
 * the forward pass can live in a CUDA graph while the input pointer changes every step. */
int  b2t_ingest_frames(b2t_ctx *ctx, const unsigned char *frames_dev, int n_seg, int seg_frames,
                       long long seg_stride_bytes, void *stream);
const float *b2t_logits(const b2t_ctx *ctx);          /* device (max_batch,G,G,5*(5+C)) of the last forward */
/* KerasYOLO.extract (KerasYOLO.py:509-520) / network_extract_feat (network.c:589-598): copy a layer's
 * post-activation output (pre-pool), NHWC float32, into out_dev.  name: "norm_1".."norm_22", "conv_23",
 * "conv_feat" (= norm_22), "concat".  Returns the number of floats per frame, <0 on error. */
long b2t_extract(b2t_ctx *ctx, const char *name, int batch, float *out_dev, void *stream);
int  b2t_layer_dims(const b2t_ctx *ctx, const char *name, int *h, int *w, int *c);

/* ---- anchor decode + threshold + per-class NMS: replaces decode_netout (utils.py:208-257) ---- */
/* logits_dev (B,G,G,A,5+C) fp32 -> boxes_dev (B,max_boxes,8) rows [x,y,w,h,conf,score,label,anchor_id]
 * in row-major anchor order, counts_dev (B) int32.  anchors: host, 2*A floats. */
int  b2t_decode_nms(b2t_ctx *ctx, const float *logits_dev, int batch, int grid_h, int grid_w, int n_box,
                    int n_class, float obj_threshold, float nms_threshold, const float *anchors,
                    float *boxes_dev, int *counts_dev, int max_boxes, void *stream);
/* darknet flavour: forward_region_layer + get_region_detections + do_nms_obj (region_layer.c:158-185,
 * :364-437; box.c:21-55) incl. letterbox un-mapping (correct_region_boxes :336-362, relative=0).
 * rows [cx,cy,w,h (pixels of the orig_w x orig_h frame), objectness, best prob, best class, anchor_id],
 * sorted by -best prob (YOLO.py:159), only rows with some prob > 0 after NMS. */
int  b2t_region_detect(b2t_ctx *ctx, const float *logits_dev, int batch, int grid_h, int grid_w, int n_box,
                       int n_class, float thresh, float nms_threshold, const float *anchors,
                       int orig_w, int orig_h, int net_w, int net_h,
                       float *dets_dev, int *counts_dev, int max_dets, void *stream);

/* ---- recurrent trackers ------------------------------------------------------------------- */
/* TinyTracker / TinyHeatmapTracker (TinyTracker.py:25-41, TinyHeatmapTracker.py:26-48): an opaque
 * LSTM head bound to the context.  pool: 0 = Global max (F = C_feat), 1 = MaxPooling2D(4,4)+Flatten. */
typedef struct b2t_lstm b2t_lstm;
int  b2t_lstm_create(b2t_ctx *ctx, int n_feat, int n_det, int units, int n_out, int max_streams,
                     b2t_lstm **out);
void b2t_lstm_destroy(b2t_lstm *l);
/* Keras layouts: kernel (n_feat+n_det, 4u), recurrent (u,4u), bias (4u) gate order i,f,c,o;
 * dense_kernel (u,n_out), dense_bias (n_out).  Host pointers. */
int  b2t_lstm_set_weights(b2t_lstm *l, const float *kernel, const float *recurrent, const float *bias,
                          const float *dense_kernel, const float *dense_bias, void *stream);
int  b2t_lstm_reset(b2t_lstm *l, int stream_index /* -1 = all */, void *stream);
/* fv_dev (S,n_feat) fp32 pooled features, det_dev (S,n_det) fp32 -> y_dev (S,n_out) fp32; state persists.
 * *_stride = elements between consecutive streams' rows (<= 0: dense), so that time step t of S windows
 * stored (S,T,F) can be stepped without a gather. */
int  b2t_lstm_step(b2t_lstm *l, const float *fv_dev, int fv_stride, const float *det_dev, int det_stride,
                   int n_streams, float *y_dev, int y_stride, int hard_sigmoid, void *stream);
/* the same for streams [slot0, slot0 + n_streams) of the head's max_streams state slots; the other streams keep
 * their (h, c): BaseTracker.step(frame, stream=k) for interleaved online streams */
int  b2t_lstm_step_slots(b2t_lstm *l, int slot0, const float *fv_dev, int fv_stride, const float *det_dev,
                         int det_stride, int n_streams, float *y_dev, int y_stride, int hard_sigmoid, void *stream);
/* A whole window in one call: fv_dev (S,T,n_feat), det_dev (S,T,n_det) dense -> y_dev (S,T,n_out).  The input
 * projection x*W of all S*T rows is one launch, only h*U + gates is sequential (T launches), the Dense head is one
 * launch.  reset != 0 zeroes (h,c) first = Keras' stateless windows. T <= 16. */
int  b2t_lstm_sequence(b2t_lstm *l, const float *fv_dev, const float *det_dev, int n_streams, int n_steps,
                       float *y_dev, int reset, int hard_sigmoid, void *stream);
/* Frame ingest: cv2.resize(frame, (dst_w, dst_h)) -- KerasYOLO.py:526, MultiObjDetTracker.py:300 -- for a batch of
 * (src_h, src_w, 3) uint8 frames on the device; bit-identical to OpenCV's INTER_LINEAR for 8-bit images.
 * Not capturable in a CUDA graph the first time a (src, dst) geometry is seen (coefficient tables are uploaded). */
int  b2t_resize_frames(b2t_ctx *ctx, const unsigned char *src_dev, int src_h, int src_w, int batch,
                       unsigned char *dst_dev, int dst_h, int dst_w, void *stream);
/* darknet's ingest for YOLO.detect (YOLO.py:141-145): load_image_color's float(u8)/255 RGB conversion + letterbox_image
 * (image.c:960-979; network_predict_image, network.c:609-616) for `batch` (src_h, src_w, 3) uint8 frames on the
 * device -> dst_dev (batch, image_h, image_w, 3) float32, ready for b2t_yolo_forward(B2T_FRAME_F32).
 * swap_rb != 0: the source is BGR (cv2.imread). */
int  b2t_letterbox_frames(b2t_ctx *ctx, const unsigned char *src_dev, int src_h, int src_w, int batch, int swap_rb,
                          float *dst_dev, void *stream);
/* pooled feature of the last forward's conv layer `name` for frames [0,batch): Global -> (B,C);
 * Max -> (B,(H/4)*(W/4)*C).  chw_view=1 reproduces preprocessing.py:419 (CHW buffer viewed as HWC). */
int  b2t_pool_features(b2t_ctx *ctx, const char *name, int batch, int pool_mode, int chw_view,
                       float *fv_dev, void *stream);
/* generate_heatmap_feat / generate_rectangle_from_heatmap (utils.py:53-79) on the device */
int  b2t_heatmap_from_box(b2t_ctx *ctx, const float *xywh_dev /* (S,4) top-left x,y,w,h rel. */, int n,
                          int size, float *heat_dev /* (S,size*size) */, void *stream);
/* preprocessing.py:434-456 + YOLO.py:177-180: first (highest-prob) row of b2t_region_detect whose class is
 * allowed -> LSTM bbox input [cx/w,cy/h,bw/w,bh/h] (zeros if none), optionally its heat-map, chosen row or -1 */
int  b2t_select_detection(b2t_ctx *ctx, const float *dets_dev, const int *counts_dev, int max_dets, int batch,
                          const unsigned char *class_mask_dev /* n_class bytes or NULL */, int frame_w, int frame_h,
                          float *det_in_dev /* (B,4) */, int heat_size, float *heat_dev /* or NULL */,
                          int *chosen_dev /* or NULL */, void *stream);
int  b2t_box_from_heatmap(b2t_ctx *ctx, const float *heat_dev, int n, int size, float thresh,
                          int *rect_dev /* (S,4) x1,y1,x2,y2 */, void *stream);

/* MultiObjDetTracker (MultiObjDetTracker.py:160-189): ConvLSTM2D over concat[conv_23 logits, conv_feat]
 * of the last b2t_yolo_forward, then the 1x1 head.  The context holds max_batch recurrent-state slots (h, c).
 * b2t_convlstm_sequence: frames [0, S*T) of the last forward are S streams x T consecutive time steps (frame index
 * s*T + t; TimeDistributed, :162-171); stream s uses state slot slot0 + s.  reset != 0 zeroes those slots first
 * (= Keras' stateless windows).  The input conv and the head run once over all S*T frames; the recurrent conv runs
 * once per time step over the S streams.  trk_logits_dev (S*T,G,G,5*(5+C)) fp32.
 * b2t_convlstm_window(batch) = sequence(S=1, T=batch, slot0=0, reset=0). */
int  b2t_convlstm_reset(b2t_ctx *ctx, void *stream);                               /* every slot */
int  b2t_convlstm_reset_slots(b2t_ctx *ctx, int slot0, int n_slots, void *stream);
int  b2t_convlstm_sequence(b2t_ctx *ctx, int n_streams, int n_steps, int slot0, int reset, float *trk_logits_dev,
                           int hard_sigmoid, void *stream);
int  b2t_convlstm_window(b2t_ctx *ctx, int batch, float *trk_logits_dev, int hard_sigmoid, void *stream);

/* ---- callers after the path ---------------------------------------------------- */
/* draw_boxes (utils.py:190-206): cv2.rectangle(image, (xmin,ymin), (xmax,ymax), colour, 3) for the first counts[b] rows
 * [x,y,w,h,...] (b2t_decode_nms layout, image-relative centre form) of every frame, in place on (B,H,W,3) uint8 frames;
 * pixel-identical to OpenCV's thickness-3 rectangle.  The cv2.putText label stays on the host. */
int  b2t_draw_boxes(b2t_ctx *ctx, unsigned char *frames_dev, int batch, int h, int w, const float *rows_dev,
                    const int *counts_dev, int max_rows, int c0, int c1, int c2, void *stream);
/* overlap_score / average_overlap_score (utils.py:82-110): n pairs of corner boxes (x1,y1,x2,y2), float64 like the
 * reference's Python floats -> scores_dev (n) and their left-to-right mean (mean_dev, may be NULL), bit-identical. */
int  b2t_overlap_scores(b2t_ctx *ctx, const double *y_true_dev, const double *y_pred_dev, int n, double *scores_dev,
                        double *mean_dev, void *stream);

/* ---- CUDA graphs: one launch per step for a plain-C host ---------------- */
/* Capture every b2t_* call made on `stream` between begin and end (all entry points are stream-ordered; the ones that
 * synchronise -- b2t_finalize, b2t_lstm_set_weights, the first b2t_resize_frames of a geometry -- must precede the
 * capture; device pointers passed during the capture are baked into the graph), then replay with b2t_graph_launch. */
typedef struct b2t_graph b2t_graph;
int  b2t_graph_begin(b2t_ctx *ctx, void *stream /* non-default */);
int  b2t_graph_end(b2t_ctx *ctx, void *stream, b2t_graph **out);
int  b2t_graph_launch(b2t_graph *g, void *stream);
void b2t_graph_destroy(b2t_graph *g);

/* ---- multi-GPU: the one collective of the path --------------------------------------- */
/* ncclBroadcast of the packed weight blob from rank `root` over `nccl_comm` (an ncclComm_t; NCCL is resolved with
 * dlopen at run time).  Every rank calls it after b2t_bind_memory / rank `root` after b2t_finalize(upload=1); the other
 * ranks then call b2t_finalize(ctx, 0, stream).  Streams are independent units: nothing else is exchanged. */
int  b2t_broadcast_weights(b2t_ctx *ctx, void *nccl_comm, int root, void *stream);

/* ---- introspection for bench.py ------------------------------------------------------------ */
long b2t_launch_count(const b2t_ctx *ctx);            /* kernels launched by this context so far */
/* per-conv timing with CUDA events: runs the forward once, fills ms[23] and (optional) the algorithmic
 * bytes[23] of each conv (weights + input + output, fp32-equivalent 4 B/element) */
int  b2t_profile_forward(b2t_ctx *ctx, const void *frames_dev, int frame_dtype, int batch, float *ms,
                         double *bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200TRACK_H */
