/*
 * darknet_compat.h -- the reference's EXISTING C-ABI for this path, re-exported by libb200track.so.
 *
 * models_detection/YOLO.py:58-119 binds these symbols of darknet's libdarknet.so with ctypes
 * (declared in darknet/include/darknet.h:586-775 and, for the fork's additions, darknet/src/network.h:11-20,
 * network.c:589-607).  libb200track.so exports them with the same names, argument lists and struct layouts
 * (darknet.h:35-40 metadata, :507-512 image, :514-525 box/detection; YOLO.py:6-37), so the reference's YOLO.py
 * works unchanged with  CDLL("…/libb200track.so")  in place of  CDLL("darknet/libdarknet.so")  -- see
 * INTEGRATION.md.  The forward pass, the region layer, box decoding, letterbox un-mapping and do_nms_obj run on
 * the B200; only cfg/.data parsing and array marshalling are host code.
 *
 * Differences from darknet, all deliberate:
 *  - errors never exit() the process: load_network returns NULL, the others return empty results, and the text is
 *    available from b2t_last_error() (darknet: error()/file_error() print and exit, utils.c:253-285);
 *  - accepted graphs: cfg/yolov2.cfg, cfg/yolov2-voc.cfg (23 conv layers) and cfg/yolov2-tiny.cfg, cfg/yolov2-tiny-voc.cfg
 *    (9 conv layers, sixth max-pool with stride 1); any class count, anchors and square input multiple of 32;
 *  - load_image_color decodes Huffman-coded JPEG (sequential and progressive; bit-exact with darknet's stb_image path,
 *    csrc/jpeg_decode.cu) and binary PPM (P6); PNG / BMP callers hold pixels and use make_image()/network_predict.
 */
#ifndef B200_DARKNET_COMPAT_H
#define B200_DARKNET_COMPAT_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y, w, h; } box;                              /* darknet.h:514-516 */
typedef struct detection {                                             /* darknet.h:518-525 */
    box bbox;
    int classes;
    float *prob;
    float *mask;
    float objectness;
    int sort_class;
} detection;
typedef struct { int w, h, c; float *data; } image;                    /* darknet.h:507-512 */
typedef struct { int classes; char **names; } metadata;                /* darknet.h:35-40   */
typedef struct { int size; float *feat; } feature;                     /* network.h:11-14 (fork) */
typedef struct { int w, h, c; } dims;                                  /* network.h:16-20 (fork) */
typedef struct network network;                                        /* opaque here */

void cuda_set_device(int n);                                           /* cuda.c:13-18      */
network *load_network(char *cfg, char *weights, int clear);            /* network.c:53-61   */
void free_network(network *net);
metadata get_metadata(char *file);                                     /* option_list.c:35-50 */
image make_image(int w, int h, int c);
image load_image_color(char *filename, int w, int h);                  /* image.c:1482 (JPEG, PPM) */
void rgbgr_image(image im);                                            /* image.c:515-525   */
void free_image(image m);
float *network_predict(network *net, float *input);                    /* network.c:507-518; CHW float, net-sized */
float *network_predict_image(network *net, image im);                  /* network.c:609-616; letterboxes to net size */
int network_width(network *net);
int network_height(network *net);
detection *get_network_boxes(network *net, int w, int h, float thresh, float hier, int *map, int relative,
                             int *num);                                /* network.c:572-577 */
void do_nms_obj(detection *dets, int total, int classes, float thresh);/* box.c:21-55       */
void free_detections(detection *dets, int n);                          /* network.c:579-587 */
void free_ptrs(void **ptrs, int n);                                    /* utils.c:237-242   */
feature network_extract_feat(network *net, int n);                     /* network.c:589-598: layers[n-1].output, CHW, borrowed */
dims layer_dims(network *net, int n);                                  /* network.c:600-607 */

#ifdef __cplusplus
}
#endif
#endif
