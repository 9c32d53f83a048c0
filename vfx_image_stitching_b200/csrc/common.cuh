// Shared declarations of the b200sift CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <sched.h>
#include <time.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/b200sift.h"

namespace b200 {

void set_error(const char *fmt, ...);
// B200SIFT_TIMELINE=1: tl_mark() records an event on `s` after whatever was just queued there;
// tl_report() (end of detect_describe / match_pairs) prints all marks in microseconds since the first.
void tl_mark(cudaStream_t s, const char *fmt, ...);
void tl_report();

#define B200_CUDA(call)                                                              \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            b200::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,             \
                            cudaGetErrorString(e__));                                \
            return B200SIFT_ECUDA;                                                   \
        }                                                                            \
    } while (0)

#define B200_CHECK(expr)                 \
    do {                                 \
        int rc__ = (expr);               \
        if (rc__ != 0) return rc__;      \
    } while (0)

#define B200_ARG(cond)                                                   \
    do {                                                                 \
        if (!(cond)) {                                                   \
            b200::set_error("%s:%d bad argument: %s", __FILE__, __LINE__, #cond); \
            return B200SIFT_EARG;                                        \
        }                                                                \
    } while (0)

// device counters (b200sift_ctx::d_counters): stage totals, the work queues of the warp-per-item
// kernels, the sizes of the descriptor work classes, then CNT_PER_IMG counters per image
// (candidates, localized, oriented, output)
enum {
    CNT_CAND = 0, CNT_LOC = 1, CNT_RAW = 2, CNT_OUT = 3, CNT_WORK_DESC = 4, CNT_WORK_ORI = 5,
    CNT_CLASS = 8,   // kDescClasses entries
    CNT_HDR = 16, CNT_PER_IMG = 4
};
// Oriented keypoints are binned by descriptor window size when they are emitted, largest class
// first in the descriptor kernel's work queue: the windows differ 6x in area, and a large one picked
// up last would leave its warp running alone (ncu, round 2: SMs idle for a quarter of the kernel).
constexpr int kDescClasses = 5;

constexpr int kMaxOctaves = 24;
constexpr int kMaxLayers = 10;  // num_intervals + 3 <= 10
constexpr int kMaxBlurRadius = 64;

// Gaussian pyramid of a batch of same-shape images, resident in HBM.
// Layer-major inside an octave: [layer][image][row][pitch] float32, so that
// one blur launch covers a whole layer of all images; pitch is the row
// length rounded up to 8 floats (32 B sectors, float4-aligned rows).
struct Pyramid {
    int n_img = 0, n_oct = 0, n_layers = 0;
    int h[kMaxOctaves], w[kMaxOctaves], pitch[kMaxOctaves];
    size_t oct_off[kMaxOctaves];  // float offset of octave o
    float *base = nullptr;        // device allocation
    size_t floats = 0, capacity_floats = 0;
    __host__ __device__ size_t img_stride(int o) const { return (size_t)h[o] * pitch[o]; }
    __host__ __device__ float *layer(int o, int l, int img = 0) const {
        return base + oct_off[o] + ((size_t)l * n_img + img) * img_stride(o);
    }
};

// Device-side view handed to the detection kernels (no host pointers inside).
struct PyrView {
    int n_img, n_oct, n_layers;
    int h[kMaxOctaves], w[kMaxOctaves], pitch[kMaxOctaves];
    const float *oct[kMaxOctaves];  // base of octave o
    __device__ __forceinline__ const float *layer(int o, int l, int img) const {
        return oct[o] + ((size_t)l * n_img + img) * ((size_t)h[o] * pitch[o]);
    }
};

// Gaussian tap sets of one context: host mirror + sigma cache of b200sift_ctx::d_taps (pyramid.cu).
struct TapCache {
    double sigma[16] = {};
    int radius[16] = {};
    bool valid[16] = {};
    float taps[16][kMaxBlurRadius + 1] = {};
};

// 3x3x3 extremum that passed is_pixel_an_extremum.
struct Candidate {
    uint32_t img_o_l;  // img << 16 | octave << 8 | layer
    uint32_t yx;       // y << 16 | x
};

// Localized extremum (output of the quadratic fit), input of orientation.
struct Localized {
    float x, y, size, response;  // base-image coordinates (before the 0.5 conversion)
    int32_t octave_packed;
    uint32_t img_o_l;            // img << 16 | octave << 8 | final layer
    uint64_t order;              // scan-order key of the originating candidate
};

// Oriented keypoint before sorting.
struct RawKeypoint {
    float x, y, size, angle, response;
    int32_t octave_packed;
    uint32_t img;
    uint32_t pad;
    uint64_t order;  // (candidate scan order << 6) | peak bin
};

// one requested image pair of the batched matcher / its result
struct PairDesc { int offA, nA, offB, nB, imgA, imgB; };
// An image appended from a neighbour rank's exchange buffer (b200sift_append_exchange): its keypoint
// count is only known on the device (header word of the wire buffer); the host lays it out with
// `cap` rows and a kernel patches the real count into the matcher's tables.
struct RemoteImage { int image; const int32_t *d_count; int cap; };
struct PairResult { double dx, dy; int n_matches, best; float xyxy[4]; };

struct DetectParams {
    int num_intervals, border, max_iter, ori_bins;
    float dog_thresh;       // floor(0.5*contrast/num_intervals*255)
    float contrast_thr_f;   // (float)contrast_threshold
    float eigen_ratio_f;
    float sigma_f;
    float radius_factor_f;
    double scale_factor;
    double peak_ratio;
    double scale_multiplier_half;  // scale_multiplier*0.5
    float descriptor_max_value_f;
};

}  // namespace b200

struct b200sift_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_desc0 = nullptr, ev_desc1 = nullptr;   // around the descriptor kernel of the last detect_describe
    bool desc_timed = false;
    // host waits: spinning (lowest latency) or, when host cores are scarce (several ranks / contexts per
    // core), sleeping on an event created with cudaEventBlockingSync
    int sync_mode = 0;   // 0 spin, 1 poll + sched_yield, 2 sleep
    cudaEvent_t ev_sync = nullptr;
    // second stream: the extrema scan of octave o overlaps the blurs of octaves > o, and the keypoint
    // sort overlaps the descriptor kernel (dependencies by events, no host involvement)
    cudaStream_t side_stream = nullptr;
    cudaStream_t blur_side_stream = nullptr;  // non-seeding layers of each octave (build_octaves)
    void *h_pin = nullptr; size_t h_pin_cap = 0;  // pinned scratch for small device->host results
    bool sort_fast = false;                   // last run_sort_async took the per-image shared-memory path
    int sort_dedupe = 0;
    int pyr_o_tail = 0;                       // first octave produced by the pyramid tail kernel (0 = none)
    cudaStream_t blur_stream = nullptr;       // launch override used by build_octaves, else `stream`
    cudaEvent_t ev_seed = nullptr, ev_blur_side = nullptr;
    cudaEvent_t ev_oct[b200::kMaxOctaves] = {}, ev_side = nullptr, ev_main = nullptr;
    bool oct_events_valid = false;
    float last_ms = 0.f;
    long long launches = 0;

    b200::Pyramid pyr;
    b200::TapCache taps;            // Gaussian tap sets of this context (host mirror)
    float *d_taps = nullptr;        // [16][kMaxBlurRadius + 1] in global memory
    float *d_up = nullptr;  size_t up_cap = 0;     // upsampled (pre-blur) base, [img][2h][pitch0]
    uint8_t *d_in = nullptr; size_t in_cap = 0;    // uploaded input images
    void **d_ptrs = nullptr, **h_ptrs = nullptr; size_t ptrs_cap = 0;  // pointer table of device-resident inputs
    float *d_dog = nullptr; size_t dog_cap = 0;    // materialised DoG (stage API only)
    b200::Pyramid dog_pyr;                         // caller-supplied DoG layers (stage API only)

    // sparse stage
    b200::Candidate *d_cand = nullptr; int cand_cap = 0;
    b200::Localized *d_loc = nullptr;  int loc_cap = 0;
    b200::RawKeypoint *d_raw = nullptr; int raw_cap = 0;
    uint8_t *d_raw_desc = nullptr;                 // [raw_cap][128]
    int32_t *d_class_idx = nullptr;                // [kDescClasses][raw_cap] raw indices per work class
    uint32_t *d_sort_idx = nullptr; uint32_t *d_keep = nullptr; uint32_t *d_pos = nullptr;
    void *d_cub_tmp = nullptr; size_t cub_tmp_cap = 0;
    int *d_seg = nullptr; size_t seg_cap = 0; std::vector<int> h_seg;   // per-image segments of the sort
    b200sift_keypoint *d_kps = nullptr;            // final, compact, image-major
    uint8_t *d_desc = nullptr;                     // final [n][128]
    int out_cap = 0;
    // counters: [0]=n_cand [1]=n_loc [2]=n_raw [3]=n_out [4]=overflow flags ; then per image x4
    int32_t *d_counters = nullptr; int32_t *h_counters = nullptr; int counters_len = 0;
    std::vector<int> img_off;      // n_img+1 prefix of final keypoints per image
    std::vector<int> stat_cand, stat_loc, stat_raw;
    int n_img_last = 0;     // images addressable by index: detected + appended
    int n_img_detected = 0; // images of the last detect_describe (what get_all_keypoints returns)
    bool have_results = false;

    // matcher scratch
    uint8_t *d_mA = nullptr, *d_mB = nullptr; size_t mA_cap = 0, mB_cap = 0;
    int32_t *d_mout = nullptr; size_t mout_cap = 0;
    void *d_misc = nullptr; size_t misc_cap = 0;
    uint8_t *d_pair = nullptr; size_t pair_cap = 0;   // batched pair matching scratch
    uint8_t *d_tc = nullptr; size_t tc_cap = 0;       // tensor-core matcher: packed descriptors + norms
    std::vector<unsigned char> h_tc_tables;           // host copies of its small tables (kept alive for async copies)
    uint8_t *d_tcsrc = nullptr; size_t tcsrc_cap = 0; // generic match(): A and B side by side
    b200::PairResult *d_pair_res = nullptr; int32_t *d_pair_ia = nullptr, *d_pair_ib = nullptr;
    float *d_pair_xy = nullptr; int pair_rows_max = 0, pair_n = 0;
    int last_tiles_per_chunk = 0, last_n_chunks = 0;   // grid shape of the last tensor-core matcher launch
    std::vector<int> pair_counts;
    std::vector<b200::PairDesc> h_pair_desc;
    std::vector<b200::RemoteImage> remote;      // device-counted images of the current result set
};

namespace b200 {

// wait on the host until everything queued on the context's stream has run
inline cudaError_t ctx_sync(b200sift_ctx *c)
{
    if (c->sync_mode == 0 || !c->ev_sync) return cudaStreamSynchronize(c->stream);   // spin
    cudaError_t e = cudaEventRecord(c->ev_sync, c->stream);
    if (e != cudaSuccess) return e;
    if (c->sync_mode == 2) return cudaEventSynchronize(c->ev_sync);                   // sleep (cudaEventBlockingSync)
    // mode 1: poll; give the core away between polls once the wait is not a short one, and sleep in
    // short naps once it is a long one (several waiting threads per core must not starve the threads
    // that launch work)
    for (int spins = 0;; ++spins) {
        e = cudaEventQuery(c->ev_sync);
        if (e != cudaErrorNotReady) return e;
        if (spins > 256) {
            struct timespec ts = {0, 20000};
            nanosleep(&ts, nullptr);
        } else if (spins > 64) {
            sched_yield();
        }
    }
}

template <typename T>
int ensure(T **p, size_t *cap, size_t need)
{
    if (need <= *cap && *p) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    size_t n = need + need / 4 + 256;
    cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) -> %s", n * sizeof(T), cudaGetErrorString(e));
        *cap = 0;
        return B200SIFT_ECUDA;
    }
    *cap = n;
    return 0;
}

// one-time per-device kernel attributes (b200sift_create, under the init lock)
int pyramid_init_device();
int detect_init_device();
int match_init_device();
// pyramid.cu
int pyramid_layout(b200sift_ctx *c, int n_img, int h0, int w0, int n_oct, int n_layers);   // c->pyr
int pyramid_layout_into(Pyramid &p, int n_img, int h0, int w0, int n_oct, int n_layers);
int launch_gray_upsample(b200sift_ctx *c, const void *d_in, size_t img_stride_bytes, const void *const *d_ptrs,
                         size_t row_stride, int n_img, int h, int w, int channels, int dtype, float *d_out,
                         int out_pitch);
int launch_blur(b200sift_ctx *c, const float *src, float *dst, int n_img, int h, int w, int pitch,
                size_t img_stride, double sigma, float *dst2, int h2, int w2, int pitch2,
                size_t img_stride2);
int build_octaves(b200sift_ctx *c, const double *sigmas);
int base_blur(b200sift_ctx *c, const float *d_up, double sigma_diff);
int launch_dog(b200sift_ctx *c, const float *a, const float *b, float *out, size_t n);
// detect.cu
PyrView make_view(const b200sift_ctx *c);
DetectParams make_detect_params(const b200sift_params &p);
int run_detect(b200sift_ctx *c, const b200sift_params &p, int use_dog);
PyrView make_view(const Pyramid &p);
int run_localize_direct(b200sift_ctx *c, const b200sift_params &p, int use_dog, int single_octave,
                        const int32_t *h_cand, int n, b200sift_keypoint *h_kps, int32_t *h_final_layer);
int run_orient_direct(b200sift_ctx *c, const b200sift_params &p, const b200sift_keypoint *h_kps, int n, int octave,
                      b200sift_keypoint *h_out, int32_t *h_counts);
int run_describe(b200sift_ctx *c, const b200sift_params &p, const RawKeypoint *d_raw, int n, int converted,
                 uint8_t *d_out, int use_classes);
int run_sort_gather(b200sift_ctx *c, int n_raw, int n_img, int scan_order, int dedupe, int convert, int with_desc);
int run_sort_async(b200sift_ctx *c, int n_raw, int n_img, int scan_order, int dedupe);   // on the side stream
int run_gather(b200sift_ctx *c, int n_raw, int n_img, int dedupe, int convert, int with_desc);
int ensure_sparse_for(b200sift_ctx *c, int n_img, int n_raw);
int launch_ransac(b200sift_ctx *c, const double *d_matches, int n, double thr, double *move, int32_t *best);
int launch_cyl(b200sift_ctx *c, const uint8_t *d_src, int h, int w, int ch, double f, uint8_t *d_dst);
// match.cu
int run_match(b200sift_ctx *c, const uint8_t *dA, int nA, const uint8_t *dB, int nB, int32_t *d_best_idx,
              int32_t *d_best_d2, int32_t *d_second_d2);
int run_match_pairs(b200sift_ctx *c, int n_pairs, const int *h_pairs, int thresh, double vote_thr);
int run_ratio_accept(b200sift_ctx *c, const int32_t *d_idx, const int32_t *d_d1, const int32_t *d_d2, int nA, int num,
                     int den, int32_t *d_ia, int32_t *d_ib, int32_t *d_count);
int bench_match_tc(b200sift_ctx *c, const uint8_t *hA, int nA, const uint8_t *hB, int nB, int top2, int iters,
                   float *ms_kernel);
int run_accept(b200sift_ctx *c, const int32_t *d_idx, const int32_t *d_d2, int nA, int thresh,
               const b200sift_keypoint *kA, const b200sift_keypoint *kB, int32_t *d_ia, int32_t *d_ib,
               float *d_xyxy, int32_t *d_count);

}  // namespace b200
