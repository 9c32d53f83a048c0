/*
 * b200sift.h -- C ABI of the B200-native (sm_100a) SIFT detect / describe /
 * match hot path.
 *
 * The reference (sapt36/VFX_Image_Stitching) has no FFI layer: the boundary
 * of its hot path is the Python module sift_impl (function-level API,
 * sift_impl.py:15-526) plus compute_shift_sift's matcher loop
 * (image_stitching_sift.py:52-83).  This header is what a ctypes shim with
 * those function names binds (see INTEGRATION.md and
 * vfx_image_stitching_b200/sift_impl.py); every entry point cites the
 * reference lines it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types in signatures
 *     (a stream is passed as void* = cudaStream_t);
 *   - every function returns 0 on success, a negative B200SIFT_E* code on
 *     failure; b200sift_last_error() returns a thread-local message;
 *   - "host" pointers may be pageable or pinned; functions with an
 *     `on_device` flag also accept device pointers of the context's GPU;
 *   - one context per (process, GPU); calls on one context are not
 *     thread-safe and are synchronous from the caller's point of view;
 *   - there is no CPU fallback: without a usable sm_100 device
 *     b200sift_create fails.
 */
#ifndef B200SIFT_H
#define B200SIFT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(B200SIFT_BUILD)
#define B200SIFT_API __attribute__((visibility("default")))
#else
#define B200SIFT_API
#endif

#define B200SIFT_OK 0
#define B200SIFT_EARG (-1)      /* bad argument */
#define B200SIFT_ECUDA (-2)     /* CUDA runtime / driver error */
#define B200SIFT_ECAPACITY (-3) /* fixed-capacity device buffer overflowed */
#define B200SIFT_ESTATE (-4)    /* call order violated (e.g. no detect before match_images) */

typedef struct b200sift_ctx b200sift_ctx;

/* The cv2.KeyPoint fields the reference reads and writes (sift_impl.py:206-210,
 * 290-291, 339-341); class_id is always -1 and is not carried. */
typedef struct {
    float x, y;     /* kp.pt */
    float size;     /* kp.size */
    float angle;    /* kp.angle, degrees in [0,360) */
    float response; /* kp.response */
    int32_t octave; /* kp.octave, packed: byte0 octave, byte1 layer, bits16+ sub-layer offset */
} b200sift_keypoint;

/* Keyword arguments / defaults of the reference, as one POD
 * (sift_impl.py:15,117,170,247,361-362). */
typedef struct {
    double sigma;               /* 1.6  */
    int32_t num_intervals;      /* 3    */
    double assumed_blur;        /* 0.5  */
    int32_t image_border_width; /* 5    */
    double contrast_threshold;  /* 0.04 */
    double eigen_ratio;         /* 10   */
    int32_t max_iter;           /* 5    */
    double radius_factor;       /* 3    */
    int32_t ori_bins;           /* 36   */
    double peak_ratio;          /* 0.8  */
    double scale_factor;        /* 1.5  */
    int32_t window_width;       /* 4   (any value with window_width^2 * desc_bins <= 1024) */
    int32_t desc_bins;          /* 8   (the 4 x 4 x 8 default has its own specialised kernel) */
    double scale_multiplier;    /* 3    */
    double descriptor_max_value;/* 0.2  */
} b200sift_params;

/* image element types accepted by detect_describe */
#define B200SIFT_U8 0
#define B200SIFT_F32 1

B200SIFT_API void b200sift_default_params(b200sift_params *p);
B200SIFT_API const char *b200sift_last_error(void);
/* version / build info, e.g. "b200sift 0.1 sm_100a" */
B200SIFT_API const char *b200sift_version(void);

B200SIFT_API int b200sift_create(int device, b200sift_ctx **ctx);
B200SIFT_API void b200sift_destroy(b200sift_ctx *ctx);
/* Run all work of this context on `cuda_stream` (a cudaStream_t; NULL = the
 * context's own stream).  Lets a torch caller order work on its stream. */
B200SIFT_API int b200sift_set_stream(b200sift_ctx *ctx, void *cuda_stream);

/* How the calling thread waits for the device inside the synchronous entry points: 0 = spin (lowest
 * latency), 1 = poll and sched_yield() between polls (for hosts with fewer cores than waiting threads,
 * e.g. eight ranks with several contexts each on 32 cores), 2 = sleep on a cudaEventBlockingSync event
 * (frees the core entirely; wake-ups cost tens of microseconds).  Default: 0, or 1 when the host has
 * fewer than 8 hardware threads per visible GPU. */
B200SIFT_API int b200sift_set_sync_mode(b200sift_ctx *ctx, int mode);

/* The CUDA stream (cudaStream_t) the context currently launches on: its own, or the one given to
 * b200sift_set_stream.  Lets a host framework order its own streams against the library's work. */
B200SIFT_API int b200sift_get_stream(b200sift_ctx *ctx, void **cuda_stream);
/* Device time in ms of the kernels of the last detect_describe / match call
 * (CUDA events on the context stream; host<->device copies excluded). */
B200SIFT_API int b200sift_last_kernel_ms(b200sift_ctx *ctx, float *ms);
/* Device time in ms of the descriptor kernel alone in the last detect_describe (CUDA events on the context
 * stream around its launch) and the number of oriented keypoints it described: the dominant kernel of the
 * step, reported by bench.py next to the blur's roofline line. */
B200SIFT_API int b200sift_last_describe_ms(b200sift_ctx *ctx, float *ms, int32_t *n_keypoints);
/* Number of kernel launches issued by this context since creation. */
B200SIFT_API int b200sift_launch_count(b200sift_ctx *ctx, long long *n);

/* ---------------------------------------------------------------------
 * Fused path: compute_keypoints_and_descriptors (sift_impl.py:15-39) for a
 * batch of n_images images of identical shape (h rows, w cols, `channels`
 * 1 or 3 = BGR, dtype u8 or f32 (f32 only with channels == 1)).
 * images[i] points at image i (row stride in bytes).  Results stay on the
 * device; n_keypoints[i] receives the keypoint count of image i.
 * ------------------------------------------------------------------- */
B200SIFT_API int b200sift_detect_describe(b200sift_ctx *ctx, const b200sift_params *params, int n_images,
                             const void *const *images, int h, int w, int channels, int dtype,
                             size_t row_stride, int on_device, int32_t *n_keypoints);

/* Copy the results of image `image` of the last detect_describe to the host.
 * Any of kps / desc_f32 / desc_u8 may be NULL.  Keypoints are in the order of
 * remove_duplicate_keypoints (sift_impl.py:299-327) and already converted to
 * input-image coordinates (:333-343); desc_* are (n,128) row-major
 * (:361-526; integer valued 0..255). */
B200SIFT_API int b200sift_get_keypoints(b200sift_ctx *ctx, int image, b200sift_keypoint *kps, float *desc_f32,
                           uint8_t *desc_u8);

/* Same for ALL images of the last detect_describe in two device->host copies: kps / desc_u8 have
 * room for `capacity` records, image-major in batch order (image i starts at the sum of the
 * counts before it); images added with b200sift_append_results are not included.  Either pointer
 * may be NULL.  B200SIFT_EARG (nothing copied) when capacity < sum(n_keypoints). */
B200SIFT_API int b200sift_get_all_keypoints(b200sift_ctx *ctx, b200sift_keypoint *kps, uint8_t *desc_u8,
                               int64_t capacity);

/* Stage counters of the last detect_describe for image `image`:
 * candidates that passed is_pixel_an_extremum (:143-163), candidates that
 * survived localize_extremum_via_quadratic_fit (:169-211), oriented keypoints
 * before de-duplication (:246-293). */
B200SIFT_API int b200sift_get_stats(b200sift_ctx *ctx, int image, int32_t *n_candidates, int32_t *n_localized,
                       int32_t *n_oriented);

/* Device-resident views of image `image`'s results (valid until the next
 * detect_describe on this context): used by the multi-GPU descriptor
 * all-gather and by callers that keep everything on the GPU. */
B200SIFT_API int b200sift_device_results(b200sift_ctx *ctx, int image, const uint8_t **d_desc_u8,
                            const b200sift_keypoint **d_kps, int32_t *n);

/* ---------------------------------------------------------------------
 * Matcher: the A->B nearest-neighbour loop of compute_shift_sift
 * (image_stitching_sift.py:63-73) on uint8 descriptors, exact integer
 * arithmetic, lowest j wins ties.  best_idx[i] = argmin_j |A_i-B_j|^2
 * (-1 if nB == 0), best_d2 / second_d2 = smallest and second smallest
 * squared distance (INT32_MAX if absent).  second_d2 may be NULL.
 * ------------------------------------------------------------------- */
B200SIFT_API int b200sift_match(b200sift_ctx *ctx, const uint8_t *A, int nA, const uint8_t *B, int nB,
                   int on_device, int32_t *best_idx, int32_t *best_d2, int32_t *second_d2);

/* Nearest / second-nearest ratio test (sift_visualizeUI.py:247-257: knnMatch(k = 2), keep m when
 * m.distance < 0.7 * n.distance) with EXACT neighbours (the reference asks approximate FLANN
 * KD-trees) and exact integers: row i of A is accepted iff it has two neighbours and
 * ratio_den^2 * best_d2 < ratio_num^2 * second_d2 (ratio = ratio_num / ratio_den, 7 / 10 in the
 * reference).  ia / ib (capacity nA) receive the accepted A rows in order and their nearest B rows,
 * *n_good their number; best_d2 / second_d2 (nA each, may be NULL) the squared distances of every row. */
B200SIFT_API int b200sift_ratio_match(b200sift_ctx *ctx, const uint8_t *A, int nA, const uint8_t *B, int nB,
                                      int on_device, int ratio_num, int ratio_den, int32_t *ia, int32_t *ib,
                                      int32_t *best_d2, int32_t *second_d2, int32_t *n_good);

/* Grid shape of the last tensor-core matcher launch of this context: B tiles (256 rows) per CTA
 * and the number of B chunks (diagnostic; the parity tests use it to prove that the shared-memory
 * ring and the TMEM double buffer wrapped). */
B200SIFT_API int b200sift_match_grid(b200sift_ctx *ctx, int32_t *tiles_per_chunk, int32_t *n_chunks);

/* compute_shift_sift's match list (image_stitching_sift.py:63-79) between two
 * images of the last detect_describe: accepted iff best_d2 < desc_thresh.
 * ia/ib (capacity = keypoints of imgA) receive the A / B keypoint indices in
 * A order, xyxy (n x 4) the (xA,yA,xB,yB) coordinates; *n_matches the count. */
B200SIFT_API int b200sift_match_images(b200sift_ctx *ctx, int imgA, int imgB, int desc_thresh, int32_t *ia,
                          int32_t *ib, float *xyxy, int32_t *n_matches);

/* The first loop of run_panorama (image_stitching_sift.py:312-327) for n_pairs
 * image pairs (pairs[2p] -> pairs[2p+1]) of the last detect_describe in ONE
 * device pass: matcher (:63-73), acceptance best < desc_thresh (:74-79) and the
 * ransac() vote (:86-111).  Per pair: shifts[2p..2p+1] = voted (dx,dy) ((0,0)
 * without matches), n_matches[p], best_index[p] (-1 without matches) and
 * best_xyxy[4p..] = the winning ((xA,yA),(xB,yB)).  Any output may be NULL. */
B200SIFT_API int b200sift_match_pairs(b200sift_ctx *ctx, int n_pairs, const int32_t *pairs, int desc_thresh,
                                      double dist_sq_thresh, double *shifts, int32_t *n_matches,
                                      int32_t *best_index, float *best_xyxy);

/* Match list of pair p of the last b200sift_match_pairs call (A order):
 * ia / ib keypoint indices and xyxy (n x 4); capacity n_matches[p]. */
B200SIFT_API int b200sift_get_pair_matches(b200sift_ctx *ctx, int p, int32_t *ia, int32_t *ib, float *xyxy);

/* Append an externally produced result set (n descriptors (n,128) uint8 + n (x,y) float32 pairs)
 * to the results of the last detect_describe as an extra image, so that pairs against it can be
 * part of b200sift_match_pairs.  This is how a rank matches its last image against the first
 * image of the next rank after the NCCL all-gather.  on_device: the pointers are
 * device pointers of this context's GPU.  *image_index receives the new image's index. */
B200SIFT_API int b200sift_append_results(b200sift_ctx *ctx, const uint8_t *desc, const float *xy, int n,
                                         int on_device, int32_t *image_index);

/* ---- multi-GPU neighbour exchange (the boundary pair of image_stitching_sift.py:312-327) ----
 * Wire format: rows of 136 bytes.  Row 0 is a header of 34 int32: header[0] = keypoints of the image,
 * header[1..n_tail] = caller-defined (copied from the host array `tail`, n_tail <= 33).  Rows 1..min(n, cap)
 * hold one keypoint each: 128 descriptor bytes + (x, y) float32.
 *
 * pack: image `image` of the last detect_describe -> dst (device, (cap+1)*136 bytes).  Stream-ordered on the
 * context stream; no host synchronisation. */
B200SIFT_API int b200sift_pack_exchange(b200sift_ctx *ctx, int image, const int32_t *tail, int n_tail, void *dst,
                                        int cap);

/* unpack, after the all-gather: `gathered` (device) = `world` blocks of (cap+1)*136 bytes.  Copies the `world`
 * headers to headers (host, world*34 int32; the one synchronisation of the exchange).  If src >= 0 and block
 * src holds no more than cap keypoints, its rows are appended as an extra image exactly like
 * b200sift_append_results and *image_index receives its index; otherwise *image_index = -1. */
B200SIFT_API int b200sift_unpack_exchange(b200sift_ctx *ctx, const void *gathered, int world, int cap, int src,
                                          int32_t *headers, int32_t *image_index);

/* The exchange without any host synchronisation: `wire` (device, (cap+1)*136 bytes in the format above,
 * e.g. the buffer a neighbour's b200sift_pack_exchange output was received into) is appended as an extra
 * image whose keypoint count min(header[0], cap) is read ON THE DEVICE: the rows are copied by a kernel
 * that reads the header, the image is laid out with cap rows, and b200sift_match_pairs[_device] patches
 * the real count into its tables before it packs.  `wire` must stay valid and unchanged until those calls
 * have run.  The caller learns header[0] > cap (truncation) from its own collective and repeats with a
 * larger cap.  Stream-ordered on the context stream. */
B200SIFT_API int b200sift_append_exchange(b200sift_ctx *ctx, const void *wire, int cap, int32_t *image_index);

/* b200sift_match_pairs whose result stays on the device: the voted (dx, dy) of pair p is written as two
 * doubles at dst + p*dst_stride bytes ((0,0) without matches) on the context stream, no host
 * synchronisation, so that it can feed a collective directly.  b200sift_get_pair_matches is not available
 * after this call. */
B200SIFT_API int b200sift_match_pairs_device(b200sift_ctx *ctx, int n_pairs, const int32_t *pairs, int desc_thresh,
                                             double dist_sq_thresh, void *dst, size_t dst_stride);

/* ransac() translation vote (image_stitching_sift.py:86-111) on the device:
 * matches n x 4 float64 (xA,yA,xB,yB) -- the reference votes on Python floats; any n; returns the
 * winning index in *best (first maximum; -1 when n == 0) and its (dx,dy) in move[2]. */
B200SIFT_API int b200sift_ransac(b200sift_ctx *ctx, const double *matches, int n, double dist_sq_thresh,
                    double *move, int32_t *best);

/* ---------------------------------------------------------------------
 * Stage API (what sift_visualizeUI.py:104-115 calls one by one).  Host
 * arrays in, host arrays out; the work runs on the GPU.
 * ------------------------------------------------------------------- */

/* cv2.GaussianBlur(src,(0,0),sigma) on float32 (sift_impl.py:56,91): h x w,
 * dense rows.  on_device: src/dst are device pointers (used by the roofline
 * bench; no copies, asynchronous until b200sift_sync). */
B200SIFT_API int b200sift_gaussian_blur(b200sift_ctx *ctx, const float *src, int h, int w, double sigma,
                           float *dst, int on_device);
B200SIFT_API int b200sift_sync(b200sift_ctx *ctx);

/* generate_base_image (sift_impl.py:45-56): image h x w float32 -> out 2h x 2w. */
B200SIFT_API int b200sift_base_image(b200sift_ctx *ctx, const float *image, int h, int w, double sigma,
                        double assumed_blur, float *out);

/* generate_gaussian_images (sift_impl.py:82-97): base h x w; sigmas[n_layers]
 * (index 0 unused, as in the reference); out_layers[o*n_layers+l] receives the
 * dense (h>>o) x (w>>o) layer. */
B200SIFT_API int b200sift_gaussian_pyramid(b200sift_ctx *ctx, const float *base, int h, int w, int n_octaves,
                              const double *sigmas, int n_layers, float *const *out_layers);

/* generate_DoG_images (sift_impl.py:100-111) for one octave-layer pair list:
 * layers as produced above; out_dogs[o*(n_layers-1)+l] = layer[l+1]-layer[l]. */
B200SIFT_API int b200sift_dog_pyramid(b200sift_ctx *ctx, const float *const *layers, int h, int w, int n_octaves,
                         int n_layers, float *const *out_dogs);

/* find_scale_space_extrema (sift_impl.py:117-140) on a caller-supplied
 * Gaussian pyramid.  dog_layers = the caller's dog_images ([o*(n_layers-1)+l],
 * extrema and the quadratic fit read them, as :124-129 do), or NULL: the DoG
 * is then the float32 difference of adjacent Gaussian layers (:109), never
 * materialised.  Keypoints come back in the reference's scan order, in
 * base-image coordinates, not de-duplicated.  *n receives the count
 * (<= capacity). */
B200SIFT_API int b200sift_find_extrema(b200sift_ctx *ctx, const b200sift_params *params,
                          const float *const *layers, const float *const *dog_layers, int h, int w,
                          int n_octaves, int n_layers, b200sift_keypoint *kps, int capacity, int32_t *n);

/* localize_extremum_via_quadratic_fit (sift_impl.py:169-211, with the gradient / Hessian of
 * :217-240) for n caller-supplied candidates, results in input order.  cand = n x (octave, layer,
 * y, x) int32 (the format of b200sift_extrema_candidates).  layers: is_dog != 0 -> DoG layers
 * (n_layers = num_intervals + 2; what the reference passes as dog_images[octave]), else Gaussian
 * layers (n_layers = num_intervals + 3; the cube is formed from their float32 differences).
 * octave_base < 0: layers hold a whole pyramid of n_octaves octaves (h x w = octave 0); otherwise
 * they hold ONLY octave `octave_base` (n_octaves = 1, h x w = that octave) and every candidate must
 * name it.  kps[i] receives pt / size / response / octave (angle = -1 like cv2.KeyPoint()),
 * final_layer[i] the layer the reference returns next to the keypoint, or -1 where it returns None. */
B200SIFT_API int b200sift_localize(b200sift_ctx *ctx, const b200sift_params *params, const float *const *layers,
                                   int is_dog, int h, int w, int n_octaves, int n_layers, int octave_base,
                                   const int32_t *cand, int n, b200sift_keypoint *kps, int32_t *final_layer);

/* compute_keypoints_with_orientations (sift_impl.py:246-293) for n keypoints on ONE Gaussian image
 * (h x w float32, the reference's gauss_img argument; `octave` is its octave argument).  Keypoint i
 * yields counts[i] oriented keypoints at out[i * ori_bins ...], ascending in histogram bin like the
 * list the reference returns; out holds n * params->ori_bins records. */
B200SIFT_API int b200sift_orientations(b200sift_ctx *ctx, const b200sift_params *params,
                                       const b200sift_keypoint *kps, int n, int octave, const float *gauss_img,
                                       int h, int w, b200sift_keypoint *out, int32_t *counts);

/* remove_duplicate_keypoints (sift_impl.py:314-327): sort by compare_keypoints
 * (:299-311) and drop repeats; in place, *n_out receives the new count. */
B200SIFT_API int b200sift_remove_duplicates(b200sift_ctx *ctx, b200sift_keypoint *kps, int n, int32_t *n_out);

/* generate_descriptors (sift_impl.py:361-526) for keypoints already converted
 * to input-image size, on a caller-supplied Gaussian pyramid.  desc_f32 is
 * (n, window_width^2 * desc_bins) row-major, 128 columns with the defaults. */
B200SIFT_API int b200sift_descriptors(b200sift_ctx *ctx, const b200sift_params *params,
                         const b200sift_keypoint *kps, int n, const float *const *layers, int h,
                         int w, int n_octaves, int n_layers, float *desc_f32);

/* Candidate stage alone (is_pixel_an_extremum, sift_impl.py:143-163): returns
 * (octave, layer, y, x) int32 quadruples in scan order. */
B200SIFT_API int b200sift_extrema_candidates(b200sift_ctx *ctx, const b200sift_params *params,
                                const float *const *layers, int h, int w, int n_octaves,
                                int n_layers, int32_t *cand, int capacity, int32_t *n);

/* cylindrical_projection (image_stitching_sift.py:117-136), uint8 h x w x ch. */
B200SIFT_API int b200sift_cylindrical_projection(b200sift_ctx *ctx, const uint8_t *src, int h, int w, int ch,
                                    double focal, uint8_t *dst);

/* blend_two_images (image_stitching_sift.py:156-202, with pad_image :139-153) on the device: imgA / imgB are
 * host uint8 BGR images (h x w x 3), (dx, dy) the shift and ref_match = {xA, yA, xB, yB} the matched pair
 * compute_shift_sift returned.  *out_h / *out_w receive the size of the result; with out == NULL the call
 * only reports that size.  alpha_float64 = 0 reproduces the reference CLI (ref_match holds Python floats,
 * numpy blends in float32), 1 the float64 arithmetic numpy uses when ref_match holds numpy float64 scalars. */
B200SIFT_API int b200sift_blend_two_images(b200sift_ctx *ctx, const uint8_t *imgA, int hA, int wA,
                                           const uint8_t *imgB, int hB, int wB, double dx, double dy,
                                           const double *ref_match, int alpha_float64, uint8_t *out,
                                           size_t out_capacity, int32_t *out_h, int32_t *out_w);

/* The reduction of rectangle_crop (image_stitching_sift.py:224-236): box = {y_min, y_max, x_min, x_max} of the
 * pixels whose cv2 BGR2GRAY value exceeds black_threshold; y_max = -1 when there is none. */
B200SIFT_API int b200sift_crop_bbox(b200sift_ctx *ctx, const uint8_t *img, int h, int w, int black_threshold,
                                    int32_t *box);

/* Measurement hook for the roofline line of bench.py: runs `iters` launches of
 * the Gaussian blur kernel for `sigma` on an internal n_img x h x w float32
 * batch (pitch = w rounded up to 8) that is already resident in HBM, and
 * returns the mean device time per launch in ms (CUDA events on the context
 * stream).  flush_l2 != 0 writes a 256 MiB scratch buffer before every timed
 * launch.  Algorithmic traffic of one launch: 8 * n_img * h * w bytes. */
B200SIFT_API int b200sift_bench_blur(b200sift_ctx *ctx, int n_img, int h, int w, double sigma, int iters,
                                     int flush_l2, float *ms_per_launch);

/* Measurement hook for the matcher sweep (BASELINE.json config 5): runs `iters` launches of the
 * tensor-core matcher kernel alone on nA x nB uint8 descriptors resident in HBM (already in the
 * packed layout) and returns the mean device time per launch in ms.  A / B = host (nA,128) / (nB,128)
 * descriptors, or both NULL for an on-device synthetic fill.  top2 != 0 selects the nearest +
 * second-nearest epilogue.  Algorithmic work: 2*128*nA*nB operations. */
B200SIFT_API int b200sift_bench_match(b200sift_ctx *ctx, const uint8_t *A, int nA, const uint8_t *B, int nB, int top2,
                                      int iters, float *ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* B200SIFT_H */
