// C ABI of libb200sift.so (include/b200sift.h): argument checking, host <->
// device staging, orchestration of the kernels in pyramid.cu / detect.cu /
// match.cu.  No arithmetic of the path happens on the host.
#include <math.h>
#include <algorithm>
#include <mutex>
#include <thread>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace b200 {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}


namespace {
struct TlMark {
    cudaEvent_t ev;
    std::string name;
};
std::vector<TlMark> g_tl;
bool tl_on()
{
    static const bool on = getenv("B200SIFT_TIMELINE") != nullptr;
    return on;
}
}  // namespace

void tl_mark(cudaStream_t s, const char *fmt, ...)
{
    if (!tl_on()) return;
    char buf[96];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    TlMark m;
    cudaEventCreate(&m.ev);
    cudaEventRecord(m.ev, s);
    m.name = buf;
    g_tl.push_back(m);
}

void tl_report()
{
    if (!tl_on() || g_tl.empty()) return;
    cudaDeviceSynchronize();
    for (size_t i = 0; i < g_tl.size(); ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, g_tl[0].ev, g_tl[i].ev);
        fprintf(stderr, "[b200sift timeline] %9.1f us  %s\n", ms * 1e3, g_tl[i].name.c_str());
    }
    for (auto &m : g_tl) cudaEventDestroy(m.ev);
    g_tl.clear();
}

struct Timer {
    b200sift_ctx *c;
    explicit Timer(b200sift_ctx *ctx) : c(ctx) { cudaEventRecord(c->ev0, c->stream); }
    void stop()
    {
        cudaEventRecord(c->ev1, c->stream);
        ctx_sync(c);
        cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1);
    }
};

// B200SIFT_TRACE=1: device time between phase boundaries of detect_describe (CUDA events on the
// context stream, so host gaps inside a phase count towards it), printed to stderr.
struct Trace {
    bool on;
    std::vector<cudaEvent_t> ev;
    std::vector<const char *> name;
    Trace() : on(getenv("B200SIFT_TRACE") != nullptr) {}
    void mark(b200sift_ctx *c, const char *n)
    {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, c->stream);
        ev.push_back(e);
        name.push_back(n);
    }
    void report()
    {
        if (!on || ev.empty()) return;
        cudaEventSynchronize(ev.back());
        for (size_t i = 1; i < ev.size(); ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
            fprintf(stderr, "[b200sift trace] %-22s %8.1f us\n", name[i], ms * 1e3);
        }
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        ev.clear();
        name.clear();
    }
};

// sift_impl.py:59-63  int(round(log(min(shape)) / log(2) - 1))
static int num_octaves(int h, int w)
{
    const int m = h < w ? h : w;
    return (int)rint(log((double)m) / log(2.0) - 1.0);
}

// sift_impl.py:66-79
static void gaussian_sigmas(double sigma, int num_intervals, double *out)
{
    const int n = num_intervals + 3;
    const double k = pow(2.0, 1. / num_intervals);
    out[0] = sigma;
    for (int i = 1; i < n; ++i) {
        const double prev = pow(k, (double)(i - 1)) * sigma, tot = k * prev;
        out[i] = sqrt(tot * tot - prev * prev);
    }
}

// Upload a caller-supplied Gaussian pyramid (dense host layers, octave-major,
// n_layers per octave) into the context pyramid (n_img = 1).
static int upload_pyramid_into(b200sift_ctx *c, Pyramid &p, const float *const *layers, int h, int w, int n_oct,
                               int n_layers)
{
    B200_ARG(layers != nullptr);
    B200_CHECK(pyramid_layout_into(p, 1, h, w, n_oct, n_layers));
    for (int o = 0; o < n_oct; ++o)
        for (int l = 0; l < n_layers; ++l) {
            const float *src = layers[o * n_layers + l];
            B200_ARG(src != nullptr);
            B200_CUDA(cudaMemcpy2DAsync(p.layer(o, l), (size_t)p.pitch[o] * 4, src, (size_t)p.w[o] * 4,
                                        (size_t)p.w[o] * 4, p.h[o], cudaMemcpyHostToDevice, c->stream));
        }
    return 0;
}

static int upload_pyramid(b200sift_ctx *c, const float *const *layers, int h, int w, int n_oct, int n_layers)
{
    c->oct_events_valid = false;
    return upload_pyramid_into(c, c->pyr, layers, h, w, n_oct, n_layers);
}

// Kernel function attributes (dynamic shared memory opt-in, carve-out) are per device and the
// library serves several devices / several contexts per device from concurrent host threads: they
// are set exactly once per device, here, and never from a launch path.
static std::mutex g_init_lock;
static bool g_device_ready[64] = {};

static int init_device_once(int device)
{
    std::lock_guard<std::mutex> guard(g_init_lock);
    if (device < 64 && g_device_ready[device]) return 0;
    B200_CHECK(pyramid_init_device());
    B200_CHECK(detect_init_device());
    B200_CHECK(match_init_device());
    if (device < 64) g_device_ready[device] = true;
    return 0;
}

static void fill_stats(b200sift_ctx *c)
{
    const int n = c->pyr.n_img;
    c->stat_cand.assign(n, 0);
    c->stat_loc.assign(n, 0);
    c->stat_raw.assign(n, 0);
    for (int i = 0; i < n; ++i) {
        c->stat_cand[i] = c->h_counters[CNT_HDR + i * CNT_PER_IMG + 0];
        c->stat_loc[i] = c->h_counters[CNT_HDR + i * CNT_PER_IMG + 1];
        c->stat_raw[i] = c->h_counters[CNT_HDR + i * CNT_PER_IMG + 2];
    }
}

}  // namespace b200

using namespace b200;

extern "C" {

void b200sift_default_params(b200sift_params *p)
{
    if (!p) return;
    p->sigma = 1.6;
    p->num_intervals = 3;
    p->assumed_blur = 0.5;
    p->image_border_width = 5;
    p->contrast_threshold = 0.04;
    p->eigen_ratio = 10;
    p->max_iter = 5;
    p->radius_factor = 3;
    p->ori_bins = 36;
    p->peak_ratio = 0.8;
    p->scale_factor = 1.5;
    p->window_width = 4;
    p->desc_bins = 8;
    p->scale_multiplier = 3;
    p->descriptor_max_value = 0.2;
}

const char *b200sift_last_error(void) { return g_err; }
const char *b200sift_version(void) { return "b200sift 0.1 sm_100a"; }

int b200sift_create(int device, b200sift_ctx **out)
{
    B200_ARG(out != nullptr);
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        set_error("no CUDA device (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
        return B200SIFT_ECUDA;
    }
    B200_ARG(device >= 0 && device < n_dev);
    B200_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libb200sift is built for sm_100a only", device, prop.major, prop.minor);
        return B200SIFT_ECUDA;
    }
    B200_CHECK(init_device_once(device));
    b200sift_ctx *c = new b200sift_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    // The context's own stream carries the critical path of a call (the blur chain that seeds the next
    // octave, refine -> orient -> describe); the two side streams carry work that only has to be done by the
    // time the main stream joins them (layers 4-5 of an octave, the extrema scan of finished octaves, the
    // keypoint sort).  Highest priority for the former, lowest for the latter: queued CTAs of the seed chain
    // are scheduled before the side work that would otherwise fill the SMs first.
    int prio_least = 0, prio_greatest = 0;
    B200_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    B200_CUDA(cudaStreamCreateWithPriority(&c->own_stream, cudaStreamNonBlocking, prio_greatest));
    c->stream = c->own_stream;
    B200_CUDA(cudaEventCreate(&c->ev0));
    B200_CUDA(cudaEventCreate(&c->ev1));
    B200_CUDA(cudaEventCreate(&c->ev_desc0));
    B200_CUDA(cudaEventCreate(&c->ev_desc1));
    B200_CUDA(cudaEventCreateWithFlags(&c->ev_sync, cudaEventBlockingSync | cudaEventDisableTiming));
    // default: spin; poll + yield when the host has fewer than 8 hardware threads per visible GPU (one
    // process per GPU with several contexts each then has more waiting threads than cores)
    c->sync_mode = std::thread::hardware_concurrency() < 8u * (unsigned)n_dev ? 1 : 0;
    B200_CUDA(cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, prio_least));
    B200_CUDA(cudaStreamCreateWithPriority(&c->blur_side_stream, cudaStreamNonBlocking, prio_least));
    B200_CUDA(cudaEventCreateWithFlags(&c->ev_seed, cudaEventDisableTiming));
    B200_CUDA(cudaEventCreateWithFlags(&c->ev_blur_side, cudaEventDisableTiming));
    for (int o = 0; o < kMaxOctaves; ++o) B200_CUDA(cudaEventCreateWithFlags(&c->ev_oct[o], cudaEventDisableTiming));
    B200_CUDA(cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming));
    B200_CUDA(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    *out = c;
    return 0;
}

void b200sift_destroy(b200sift_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    b200::ctx_sync(c);
    void *ptrs[] = {c->pyr.base, c->d_up, c->d_in, c->d_dog, c->d_cand, c->d_loc, c->d_raw, c->d_raw_desc,
                    c->d_sort_idx, c->d_keep, c->d_pos, c->d_class_idx, c->d_cub_tmp, c->d_kps, c->d_desc, c->d_counters,
                    c->d_mA, c->d_mB, c->d_mout, c->d_misc, c->d_pair, c->d_tc, c->d_tcsrc, c->d_seg, c->d_ptrs,
                    c->d_taps, c->dog_pyr.base};
    if (c->h_ptrs) cudaFreeHost(c->h_ptrs);
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaEventDestroy(c->ev_desc0);
    cudaEventDestroy(c->ev_desc1);
    if (c->ev_sync) cudaEventDestroy(c->ev_sync);
    for (int o = 0; o < kMaxOctaves; ++o) cudaEventDestroy(c->ev_oct[o]);
    cudaEventDestroy(c->ev_side);
    cudaEventDestroy(c->ev_main);
    cudaStreamSynchronize(c->side_stream);
    cudaStreamDestroy(c->side_stream);
    cudaStreamSynchronize(c->blur_side_stream);
    cudaStreamDestroy(c->blur_side_stream);
    cudaEventDestroy(c->ev_seed);
    cudaEventDestroy(c->ev_blur_side);
    cudaStreamDestroy(c->own_stream);
    delete c;
}

int b200sift_set_stream(b200sift_ctx *c, void *s)
{
    B200_ARG(c != nullptr);
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return 0;
}

int b200sift_set_sync_mode(b200sift_ctx *c, int mode)
{
    B200_ARG(c != nullptr && mode >= 0 && mode <= 2);
    c->sync_mode = mode;
    return 0;
}

int b200sift_get_stream(b200sift_ctx *c, void **s)
{
    B200_ARG(c && s);
    *s = (void *)c->stream;
    return 0;
}

int b200sift_last_kernel_ms(b200sift_ctx *c, float *ms)
{
    B200_ARG(c && ms);
    *ms = c->last_ms;
    return 0;
}

int b200sift_last_describe_ms(b200sift_ctx *c, float *ms, int32_t *n_keypoints)
{
    B200_ARG(c && ms);
    *ms = 0.f;
    if (n_keypoints) *n_keypoints = 0;
    if (!c->desc_timed || !c->have_results) return 0;
    B200_CUDA(cudaSetDevice(c->device));
    B200_CUDA(cudaEventSynchronize(c->ev_desc1));
    B200_CUDA(cudaEventElapsedTime(ms, c->ev_desc0, c->ev_desc1));
    if (n_keypoints) *n_keypoints = c->h_counters[CNT_RAW];
    return 0;
}

int b200sift_launch_count(b200sift_ctx *c, long long *n)
{
    B200_ARG(c && n);
    *n = c->launches;
    return 0;
}

int b200sift_sync(b200sift_ctx *c)
{
    B200_ARG(c != nullptr);
    B200_CUDA(cudaSetDevice(c->device));   // the event of sync modes 1 / 2 is recorded on the context's device
    B200_CUDA(b200::ctx_sync(c));
    return 0;
}

int b200sift_detect_describe(b200sift_ctx *c, const b200sift_params *params, int n_images,
                             const void *const *images, int h, int w, int channels, int dtype, size_t row_stride,
                             int on_device, int32_t *n_keypoints)
{
    B200_ARG(c && images && n_images >= 1 && h >= 2 && w >= 2);
    B200_ARG(channels == 1 || channels == 3);
    B200_ARG(dtype == B200SIFT_U8 || (dtype == B200SIFT_F32 && channels == 1));
    b200sift_params P;
    if (params) P = *params; else b200sift_default_params(&P);
    B200_ARG(P.num_intervals >= 1 && P.num_intervals + 3 <= kMaxLayers && P.sigma > 0);
    B200_CUDA(cudaSetDevice(c->device));
    c->have_results = false;
    c->remote.clear();
    const size_t esz = dtype == B200SIFT_F32 ? 4 : 1;
    const size_t min_stride = (size_t)w * channels * esz;
    if (row_stride == 0) row_stride = min_stride;
    B200_ARG(row_stride >= min_stride);

    // ---- inputs: host images are staged into one dense block per image; device-resident images
    // are read in place through a pointer table (no copies)
    const size_t img_bytes = ((min_stride * h) + 255) & ~(size_t)255;
    const uint8_t *d_in = nullptr;
    const void *const *d_ptrs = nullptr;
    size_t in_row_stride = min_stride, in_img_stride = img_bytes;
    for (int i = 0; i < n_images; ++i) B200_ARG(images[i] != nullptr);
    if (on_device) {
        if ((size_t)n_images > c->ptrs_cap) {
            if (c->d_ptrs) cudaFree(c->d_ptrs);
            if (c->h_ptrs) cudaFreeHost(c->h_ptrs);
            c->ptrs_cap = (size_t)n_images + 64;
            B200_CUDA(cudaMalloc((void **)&c->d_ptrs, c->ptrs_cap * sizeof(void *)));
            B200_CUDA(cudaMallocHost((void **)&c->h_ptrs, c->ptrs_cap * sizeof(void *)));
        }
        for (int i = 0; i < n_images; ++i) c->h_ptrs[i] = const_cast<void *>(images[i]);
        B200_CUDA(cudaMemcpyAsync(c->d_ptrs, c->h_ptrs, sizeof(void *) * n_images, cudaMemcpyHostToDevice, c->stream));
        d_ptrs = c->d_ptrs;
        in_row_stride = row_stride;
    } else {
        size_t cap = c->in_cap;
        B200_CHECK(ensure(&c->d_in, &cap, img_bytes * n_images));
        c->in_cap = cap;
        for (int i = 0; i < n_images; ++i)
            B200_CUDA(cudaMemcpy2DAsync(c->d_in + (size_t)i * img_bytes, min_stride, images[i], row_stride,
                                        min_stride, h, cudaMemcpyHostToDevice, c->stream));
        d_in = c->d_in;
    }

    const int H0 = 2 * h, W0 = 2 * w;
    const int n_oct = num_octaves(H0, W0);
    B200_ARG(n_oct >= 1);
    const int n_layers = P.num_intervals + 3;
    B200_CHECK(pyramid_layout(c, n_images, H0, W0, n_oct, n_layers));
    {
        size_t cap = c->up_cap;
        B200_CHECK(ensure(&c->d_up, &cap, (size_t)n_images * c->pyr.img_stride(0)));
        c->up_cap = cap;
    }
    double sig[kMaxLayers];
    gaussian_sigmas(P.sigma, P.num_intervals, sig);
    const double d2 = P.sigma * P.sigma - (2 * P.assumed_blur) * (2 * P.assumed_blur);
    const double sigma_diff = sqrt(d2 > 0.01 ? d2 : 0.01);  // sift_impl.py:55

    Timer tm(c);
    static Trace tr;
    tr.mark(c, "start");
    tl_mark(c->stream, "start");
    B200_CHECK(launch_gray_upsample(c, d_in, in_img_stride, d_ptrs, in_row_stride, n_images, h, w, channels, dtype, c->d_up,
                                    c->pyr.pitch[0]));
    tl_mark(c->stream, "main  upsample");
    B200_CHECK(base_blur(c, c->d_up, sigma_diff));
    tl_mark(c->stream, "main  base blur");
    tr.mark(c, "upsample+base blur");
    B200_CHECK(build_octaves(c, sig));
    tr.mark(c, "octaves (blur)");
    B200_CHECK(run_detect(c, P, 0));
    tr.mark(c, "extrema+refine+orient");
    fill_stats(c);
    const int n_raw = c->h_counters[CNT_RAW];
    B200_CHECK(run_sort_async(c, n_raw, n_images, 0, 1));         // side stream, overlaps the descriptors
    B200_CUDA(cudaEventRecord(c->ev_desc0, c->stream));
    B200_CHECK(run_describe(c, P, c->d_raw, n_raw, 0, c->d_raw_desc, 1));
    B200_CUDA(cudaEventRecord(c->ev_desc1, c->stream));
    c->desc_timed = n_raw > 0;
    tl_mark(c->stream, "main  describe");
    tr.mark(c, "describe (|| sort)");
    B200_CHECK(run_gather(c, n_raw, n_images, 1, 1, 1));
    tl_mark(c->stream, "main  gather + counts on host");
    tr.mark(c, "dedupe+gather");
    tm.stop();
    tr.report();
    tl_report();
    c->n_img_last = n_images;
    c->n_img_detected = n_images;
    c->have_results = true;
    if (n_keypoints)
        for (int i = 0; i < n_images; ++i) n_keypoints[i] = c->img_off[i + 1] - c->img_off[i];
    return 0;
}

int b200sift_get_keypoints(b200sift_ctx *c, int image, b200sift_keypoint *kps, float *desc_f32, uint8_t *desc_u8)
{
    B200_ARG(c != nullptr);
    if (!c->have_results) {
        set_error("get_keypoints before a successful detect_describe");
        return B200SIFT_ESTATE;
    }
    B200_ARG(image >= 0 && image < c->n_img_last);
    const int off = c->img_off[image], n = c->img_off[image + 1] - off;
    if (n == 0) return 0;
    B200_CUDA(cudaSetDevice(c->device));
    if (kps)
        B200_CUDA(cudaMemcpyAsync(kps, c->d_kps + off, sizeof(b200sift_keypoint) * n, cudaMemcpyDeviceToHost,
                                  c->stream));
    if (desc_u8)
        B200_CUDA(cudaMemcpyAsync(desc_u8, c->d_desc + (size_t)off * 128, (size_t)n * 128, cudaMemcpyDeviceToHost,
                                  c->stream));
    if (desc_f32) {
        // widen on the host: the device keeps (and ships) the compact uint8 form
        uint8_t *tmp = desc_u8;
        std::vector<uint8_t> buf;
        if (!tmp) {
            buf.resize((size_t)n * 128);
            tmp = buf.data();
            B200_CUDA(cudaMemcpyAsync(tmp, c->d_desc + (size_t)off * 128, (size_t)n * 128, cudaMemcpyDeviceToHost,
                                      c->stream));
        }
        B200_CUDA(b200::ctx_sync(c));
        for (size_t i = 0; i < (size_t)n * 128; ++i) desc_f32[i] = (float)tmp[i];
    }
    B200_CUDA(b200::ctx_sync(c));
    return 0;
}

int b200sift_get_all_keypoints(b200sift_ctx *c, b200sift_keypoint *kps, uint8_t *desc_u8, int64_t capacity)
{
    B200_ARG(c != nullptr);
    if (!c->have_results) {
        set_error("get_all_keypoints before a successful detect_describe");
        return B200SIFT_ESTATE;
    }
    const int n = c->img_off[c->n_img_detected];  // images appended later (append_results) are not part of it
    if (n == 0) return 0;
    if ((int64_t)n > capacity) {
        set_error("get_all_keypoints: capacity %lld < %d keypoints", (long long)capacity, n);
        return B200SIFT_EARG;
    }
    B200_CUDA(cudaSetDevice(c->device));
    if (kps)
        B200_CUDA(cudaMemcpyAsync(kps, c->d_kps, sizeof(b200sift_keypoint) * (size_t)n, cudaMemcpyDeviceToHost,
                                  c->stream));
    if (desc_u8)
        B200_CUDA(cudaMemcpyAsync(desc_u8, c->d_desc, (size_t)n * 128, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(b200::ctx_sync(c));
    return 0;
}

int b200sift_get_stats(b200sift_ctx *c, int image, int32_t *n_cand, int32_t *n_loc, int32_t *n_ori)
{
    B200_ARG(c != nullptr);
    B200_ARG(image >= 0 && image < (int)c->stat_cand.size());
    if (n_cand) *n_cand = c->stat_cand[image];
    if (n_loc) *n_loc = c->stat_loc[image];
    if (n_ori) *n_ori = c->stat_raw[image];
    return 0;
}

int b200sift_device_results(b200sift_ctx *c, int image, const uint8_t **d_desc, const b200sift_keypoint **d_kps,
                            int32_t *n)
{
    B200_ARG(c != nullptr);
    if (!c->have_results) {
        set_error("device_results before a successful detect_describe");
        return B200SIFT_ESTATE;
    }
    B200_ARG(image >= 0 && image < c->n_img_last);
    const int off = c->img_off[image];
    if (d_desc) *d_desc = c->d_desc + (size_t)off * 128;
    if (d_kps) *d_kps = c->d_kps + off;
    if (n) *n = c->img_off[image + 1] - off;
    return 0;
}

int b200sift_match(b200sift_ctx *c, const uint8_t *A, int nA, const uint8_t *B, int nB, int on_device,
                   int32_t *best_idx, int32_t *best_d2, int32_t *second_d2)
{
    B200_ARG(c && nA >= 0 && nB >= 0 && best_idx && best_d2);
    if (nA == 0) return 0;
    B200_ARG(A != nullptr && (nB == 0 || B != nullptr));
    B200_CUDA(cudaSetDevice(c->device));
    const uint8_t *dA = A, *dB = B;
    if (!on_device) {
        size_t cap = c->mA_cap;
        B200_CHECK(ensure(&c->d_mA, &cap, (size_t)nA * 128));
        c->mA_cap = cap;
        cap = c->mB_cap;
        B200_CHECK(ensure(&c->d_mB, &cap, (size_t)(nB > 0 ? nB : 1) * 128));
        c->mB_cap = cap;
        B200_CUDA(cudaMemcpyAsync(c->d_mA, A, (size_t)nA * 128, cudaMemcpyHostToDevice, c->stream));
        if (nB) B200_CUDA(cudaMemcpyAsync(c->d_mB, B, (size_t)nB * 128, cudaMemcpyHostToDevice, c->stream));
        dA = c->d_mA;
        dB = c->d_mB;
    }
    size_t cap = c->misc_cap;
    B200_CHECK(ensure((uint8_t **)&c->d_misc, &cap, (size_t)nA * 3 * sizeof(int32_t)));
    c->misc_cap = cap;
    int32_t *d_idx = (int32_t *)c->d_misc, *d_b1 = d_idx + nA, *d_b2 = d_b1 + nA;
    Timer tm(c);
    B200_CHECK(run_match(c, dA, nA, dB, nB, d_idx, d_b1, d_b2));
    tm.stop();
    B200_CUDA(cudaMemcpyAsync(best_idx, d_idx, sizeof(int32_t) * nA, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(cudaMemcpyAsync(best_d2, d_b1, sizeof(int32_t) * nA, cudaMemcpyDeviceToHost, c->stream));
    if (second_d2)
        B200_CUDA(cudaMemcpyAsync(second_d2, d_b2, sizeof(int32_t) * nA, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(b200::ctx_sync(c));
    return 0;
}

int b200sift_match_images(b200sift_ctx *c, int imgA, int imgB, int desc_thresh, int32_t *ia, int32_t *ib,
                          float *xyxy, int32_t *n_matches)
{
    B200_ARG(c && n_matches);
    if (!c->have_results) {
        set_error("match_images before a successful detect_describe");
        return B200SIFT_ESTATE;
    }
    B200_ARG(imgA >= 0 && imgA < c->n_img_last && imgB >= 0 && imgB < c->n_img_last);
    *n_matches = 0;
    const int offA = c->img_off[imgA], nA = c->img_off[imgA + 1] - offA;
    const int offB = c->img_off[imgB], nB = c->img_off[imgB + 1] - offB;
    if (nA == 0) return 0;
    B200_CUDA(cudaSetDevice(c->device));
    // scratch: idx, d1, d2, ia, ib (int32 x nA each), xyxy (float x 4nA), count
    size_t cap = c->misc_cap;
    B200_CHECK(ensure((uint8_t **)&c->d_misc, &cap, (size_t)(nA * 9 + 4) * sizeof(int32_t)));
    c->misc_cap = cap;
    int32_t *d_idx = (int32_t *)c->d_misc, *d_b1 = d_idx + nA, *d_b2 = d_b1 + nA, *d_ia = d_b2 + nA,
            *d_ib = d_ia + nA;
    float *d_xy = (float *)(d_ib + nA);
    int32_t *d_cnt = (int32_t *)(d_xy + 4 * (size_t)nA);
    Timer tm(c);
    B200_CHECK(run_match(c, c->d_desc + (size_t)offA * 128, nA, c->d_desc + (size_t)offB * 128, nB, d_idx, d_b1,
                         d_b2));
    B200_CHECK(run_accept(c, d_idx, d_b1, nA, desc_thresh, c->d_kps + offA, c->d_kps + offB, d_ia, d_ib, d_xy,
                          d_cnt));
    tm.stop();
    int32_t n = 0;
    B200_CUDA(cudaMemcpyAsync(&n, d_cnt, sizeof(n), cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(b200::ctx_sync(c));
    if (n > 0) {
        if (ia) B200_CUDA(cudaMemcpyAsync(ia, d_ia, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
        if (ib) B200_CUDA(cudaMemcpyAsync(ib, d_ib, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
        if (xyxy) B200_CUDA(cudaMemcpyAsync(xyxy, d_xy, sizeof(float) * 4 * n, cudaMemcpyDeviceToHost, c->stream));
        B200_CUDA(b200::ctx_sync(c));
    }
    *n_matches = n;
    return 0;
}

int b200sift_match_pairs(b200sift_ctx *c, int n_pairs, const int32_t *pairs, int desc_thresh, double vote_thr,
                         double *shifts, int32_t *n_matches, int32_t *best_index, float *best_xyxy)
{
    B200_ARG(c && n_pairs >= 0 && (n_pairs == 0 || pairs));
    if (!c->have_results) {
        set_error("match_pairs before a successful detect_describe");
        return B200SIFT_ESTATE;
    }
    c->pair_n = 0;
    if (n_pairs == 0) return 0;
    B200_CUDA(cudaSetDevice(c->device));
    for (int p = 0; p < n_pairs; ++p) {
        const int a = pairs[2 * p], b = pairs[2 * p + 1];
        B200_ARG(a >= 0 && a < c->n_img_last && b >= 0 && b < c->n_img_last);
    }
    Timer tm(c);
    tl_mark(c->stream, "match start");
    B200_CHECK(run_match_pairs(c, n_pairs, pairs, desc_thresh, vote_thr));
    tl_mark(c->stream, "main  pack + match + finalize");
    if (c->h_pin_cap < sizeof(PairResult) * n_pairs) {  // pinned: a pageable target would stage the copy
        if (c->h_pin) cudaFreeHost(c->h_pin);
        c->h_pin = nullptr;
        c->h_pin_cap = 0;
        B200_CUDA(cudaMallocHost(&c->h_pin, sizeof(PairResult) * n_pairs * 2));
        c->h_pin_cap = sizeof(PairResult) * n_pairs * 2;
    }
    PairResult *res = static_cast<PairResult *>(c->h_pin);
    B200_CUDA(cudaMemcpyAsync(res, c->d_pair_res, sizeof(PairResult) * n_pairs, cudaMemcpyDeviceToHost, c->stream));
    tm.stop();
    B200_CUDA(b200::ctx_sync(c));
    tl_mark(c->stream, "main  pair results on host");
    tl_report();
    c->pair_counts.assign(n_pairs, 0);
    for (int p = 0; p < n_pairs; ++p) {
        c->pair_counts[p] = res[p].n_matches;
        if (shifts) { shifts[2 * p] = res[p].dx; shifts[2 * p + 1] = res[p].dy; }
        if (n_matches) n_matches[p] = res[p].n_matches;
        if (best_index) best_index[p] = res[p].best;
        if (best_xyxy) memcpy(best_xyxy + 4 * p, res[p].xyxy, sizeof(float) * 4);
    }
    return 0;
}

int b200sift_get_pair_matches(b200sift_ctx *c, int p, int32_t *ia, int32_t *ib, float *xyxy)
{
    B200_ARG(c != nullptr);
    if (p < 0 || p >= c->pair_n || (int)c->pair_counts.size() != c->pair_n) {
        set_error("get_pair_matches: pair %d is not part of the last match_pairs call", p);
        return B200SIFT_ESTATE;
    }
    const int n = c->pair_counts[p];
    if (n == 0) return 0;
    B200_CUDA(cudaSetDevice(c->device));
    const size_t mo = (size_t)p * c->pair_rows_max;
    if (ia) B200_CUDA(cudaMemcpyAsync(ia, c->d_pair_ia + mo, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    if (ib) B200_CUDA(cudaMemcpyAsync(ib, c->d_pair_ib + mo, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    if (xyxy)
        B200_CUDA(cudaMemcpyAsync(xyxy, c->d_pair_xy + 4 * mo, sizeof(float) * 4 * n, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(b200::ctx_sync(c));
    return 0;
}

static int grow_results(b200sift_ctx *c, int used, int need);

int b200sift_append_results(b200sift_ctx *c, const uint8_t *desc, const float *xy, int n, int on_device,
                            int32_t *image_index)
{
    B200_ARG(c && image_index && n >= 0 && (n == 0 || (desc && xy)));
    if (!c->have_results) {
        set_error("append_results before a successful detect_describe");
        return B200SIFT_ESTATE;
    }
    B200_CUDA(cudaSetDevice(c->device));
    const int used = c->img_off.back();
    B200_CHECK(grow_results(c, used, used + n));
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (n > 0) {
        B200_CUDA(cudaMemcpyAsync(c->d_desc + (size_t)used * 128, desc, (size_t)n * 128, kind, c->stream));
        B200_CUDA(cudaMemsetAsync(c->d_kps + used, 0, sizeof(b200sift_keypoint) * (size_t)n, c->stream));
        // (x, y) are the first two floats of the 24-byte keypoint record
        B200_CUDA(cudaMemcpy2DAsync(c->d_kps + used, sizeof(b200sift_keypoint), xy, 2 * sizeof(float),
                                    2 * sizeof(float), n, kind, c->stream));
        if (!on_device) B200_CUDA(b200::ctx_sync(c));
    }
    *image_index = c->n_img_last;
    c->img_off.push_back(used + n);
    c->n_img_last += 1;
    return 0;
}

// ------------------------------------------------------------------ multi-GPU neighbour exchange
namespace xchg {
constexpr int kXRow = 136;   // bytes per wire row: 128 descriptor + 8 xy
struct XHeader { int32_t v[34]; };

__global__ void pack_exchange_kernel(const uint8_t *__restrict__ desc, const b200sift_keypoint *__restrict__ kps,
                                     int n_rows, uint8_t *__restrict__ dst, const XHeader hdr)
{
    // one 8-byte word per thread: 17 words per row; row 0 = header
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = t / 17, wd = t - row * 17;
    if (row > n_rows) return;
    uint2 val;
    if (row == 0) {
        val = make_uint2((unsigned)hdr.v[2 * wd], (unsigned)hdr.v[2 * wd + 1]);
    } else if (wd < 16) {
        val = reinterpret_cast<const uint2 *>(desc + (size_t)(row - 1) * 128)[wd];
    } else {
        const b200sift_keypoint &k = kps[row - 1];
        val = make_uint2(__float_as_uint(k.x), __float_as_uint(k.y));
    }
    reinterpret_cast<uint2 *>(dst + (size_t)row * kXRow)[wd] = val;
}

__global__ void unpack_exchange_kernel(const uint8_t *__restrict__ src, int n_rows, uint8_t *__restrict__ desc,
                                       b200sift_keypoint *__restrict__ kps)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = t / 17, wd = t - row * 17;
    if (row >= n_rows) return;
    const uint2 val = reinterpret_cast<const uint2 *>(src + (size_t)(row + 1) * kXRow)[wd];
    if (wd < 16) {
        reinterpret_cast<uint2 *>(desc + (size_t)row * 128)[wd] = val;
    } else {
        b200sift_keypoint k;
        k.x = __uint_as_float(val.x); k.y = __uint_as_float(val.y);
        k.size = 0.f; k.angle = 0.f; k.response = 0.f; k.octave = 0;
        kps[row] = k;
    }
}

// same as unpack_exchange_kernel with the row count read from the buffer's header on the device
__global__ void unpack_exchange_dev_kernel(const uint8_t *__restrict__ src, int cap, uint8_t *__restrict__ desc,
                                           b200sift_keypoint *__restrict__ kps)
{
    const int n_rows = max(0, min(*reinterpret_cast<const int32_t *>(src), cap));
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = t / 17, wd = t - row * 17;
    if (row >= n_rows) return;
    const uint2 val = reinterpret_cast<const uint2 *>(src + (size_t)(row + 1) * kXRow)[wd];
    if (wd < 16) {
        reinterpret_cast<uint2 *>(desc + (size_t)row * 128)[wd] = val;
    } else {
        b200sift_keypoint k;
        k.x = __uint_as_float(val.x); k.y = __uint_as_float(val.y);
        k.size = 0.f; k.angle = 0.f; k.response = 0.f; k.octave = 0;
        kps[row] = k;
    }
}

__global__ void pair_shifts_kernel(const PairResult *__restrict__ res, int n, uint8_t *__restrict__ dst, size_t stride)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    double *o = reinterpret_cast<double *>(dst + (size_t)p * stride);
    const bool any = res[p].n_matches > 0;
    o[0] = any ? res[p].dx : 0.0;
    o[1] = any ? res[p].dy : 0.0;
}

}  // namespace xchg
using namespace xchg;

// grow the compact result arrays to hold `need` records, keeping their contents
static int grow_results(b200sift_ctx *c, int used, int need)
{
    if (need <= c->out_cap) return 0;
    const int ncap = need + need / 4 + 1024;
    b200sift_keypoint *nk = nullptr;
    uint8_t *nd = nullptr;
    B200_CUDA(cudaMalloc((void **)&nk, sizeof(b200sift_keypoint) * (size_t)ncap));
    B200_CUDA(cudaMalloc((void **)&nd, (size_t)ncap * 128));
    B200_CUDA(cudaMemcpyAsync(nk, c->d_kps, sizeof(b200sift_keypoint) * (size_t)used, cudaMemcpyDeviceToDevice,
                              c->stream));
    B200_CUDA(cudaMemcpyAsync(nd, c->d_desc, (size_t)used * 128, cudaMemcpyDeviceToDevice, c->stream));
    B200_CUDA(b200::ctx_sync(c));
    cudaFree(c->d_kps);
    cudaFree(c->d_desc);
    c->d_kps = nk;
    c->d_desc = nd;
    c->out_cap = ncap;
    return 0;
}

int b200sift_pack_exchange(b200sift_ctx *c, int image, const int32_t *tail, int n_tail, void *dst, int cap)
{
    B200_ARG(c && dst && cap >= 0 && n_tail >= 0 && n_tail <= 33 && (n_tail == 0 || tail));
    if (!c->have_results) {
        set_error("pack_exchange before a successful detect_describe");
        return B200SIFT_ESTATE;
    }
    B200_ARG(image >= 0 && image < c->n_img_last);
    B200_CUDA(cudaSetDevice(c->device));
    const int off = c->img_off[image], n = c->img_off[image + 1] - off;
    XHeader h;
    memset(&h, 0, sizeof(h));
    h.v[0] = n;
    for (int i = 0; i < n_tail; ++i) h.v[1 + i] = tail[i];
    const int rows = n < cap ? n : cap;
    const int threads = (rows + 1) * 17;
    pack_exchange_kernel<<<(threads + 255) / 256, 256, 0, c->stream>>>(c->d_desc + (size_t)off * 128, c->d_kps + off,
                                                                        rows, static_cast<uint8_t *>(dst), h);
    c->launches++;
    B200_CUDA(cudaGetLastError());
    return 0;
}

int b200sift_unpack_exchange(b200sift_ctx *c, const void *gathered, int world, int cap, int src, int32_t *headers,
                             int32_t *image_index)
{
    B200_ARG(c && gathered && headers && image_index && world >= 1 && cap >= 0 && src < world);
    if (!c->have_results) {
        set_error("unpack_exchange before a successful detect_describe");
        return B200SIFT_ESTATE;
    }
    B200_CUDA(cudaSetDevice(c->device));
    *image_index = -1;
    const size_t block = (size_t)(cap + 1) * kXRow;
    const size_t need_pin = (size_t)world * kXRow;
    if (c->h_pin_cap < need_pin) {
        if (c->h_pin) cudaFreeHost(c->h_pin);
        c->h_pin = nullptr;
        c->h_pin_cap = 0;
        B200_CUDA(cudaMallocHost(&c->h_pin, need_pin * 2));
        c->h_pin_cap = need_pin * 2;
    }
    B200_CUDA(cudaMemcpy2DAsync(c->h_pin, kXRow, gathered, block, kXRow, world, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(b200::ctx_sync(c));
    memcpy(headers, c->h_pin, need_pin);
    if (src < 0) return 0;
    const int n = headers[(size_t)src * 34];
    B200_ARG(n >= 0);
    if (n > cap) return 0;  // truncated on the wire: the caller grows cap and repeats the exchange
    const int used = c->img_off.back();
    B200_CHECK(grow_results(c, used, used + n));
    if (n > 0) {
        const int threads = n * 17;
        unpack_exchange_kernel<<<(threads + 255) / 256, 256, 0, c->stream>>>(
            static_cast<const uint8_t *>(gathered) + (size_t)src * block, n, c->d_desc + (size_t)used * 128,
            c->d_kps + used);
        c->launches++;
        B200_CUDA(cudaGetLastError());
    }
    *image_index = c->n_img_last;
    c->img_off.push_back(used + n);
    c->n_img_last += 1;
    return 0;
}

int b200sift_append_exchange(b200sift_ctx *c, const void *wire, int cap, int32_t *image_index)
{
    B200_ARG(c && wire && image_index && cap >= 0);
    if (!c->have_results) {
        set_error("append_exchange before a successful detect_describe");
        return B200SIFT_ESTATE;
    }
    B200_CUDA(cudaSetDevice(c->device));
    const int used = c->img_off.back();
    B200_CHECK(grow_results(c, used, used + cap));
    if (cap > 0) {
        const int threads = cap * 17;
        unpack_exchange_dev_kernel<<<(threads + 255) / 256, 256, 0, c->stream>>>(
            static_cast<const uint8_t *>(wire), cap, c->d_desc + (size_t)used * 128, c->d_kps + used);
        c->launches++;
        B200_CUDA(cudaGetLastError());
    }
    *image_index = c->n_img_last;
    RemoteImage r;
    r.image = c->n_img_last;
    r.d_count = static_cast<const int32_t *>(wire);
    r.cap = cap;
    c->remote.push_back(r);
    c->img_off.push_back(used + cap);   // laid out with the capacity; the matcher's tables get the real count
    c->n_img_last += 1;
    return 0;
}

int b200sift_match_pairs_device(b200sift_ctx *c, int n_pairs, const int32_t *pairs, int desc_thresh, double vote_thr,
                                void *dst, size_t dst_stride)
{
    B200_ARG(c && n_pairs >= 0 && (n_pairs == 0 || (pairs && dst && dst_stride >= 2 * sizeof(double))));
    if (!c->have_results) {
        set_error("match_pairs_device before a successful detect_describe");
        return B200SIFT_ESTATE;
    }
    c->pair_n = 0;
    c->pair_counts.clear();
    if (n_pairs == 0) return 0;
    B200_CUDA(cudaSetDevice(c->device));
    for (int p = 0; p < n_pairs; ++p) {
        const int a = pairs[2 * p], b = pairs[2 * p + 1];
        B200_ARG(a >= 0 && a < c->n_img_last && b >= 0 && b < c->n_img_last);
    }
    B200_CHECK(run_match_pairs(c, n_pairs, pairs, desc_thresh, vote_thr));
    pair_shifts_kernel<<<(n_pairs + 127) / 128, 128, 0, c->stream>>>(c->d_pair_res, n_pairs, static_cast<uint8_t *>(dst),
                                                                     dst_stride);
    c->launches++;
    B200_CUDA(cudaGetLastError());
    c->pair_n = 0;  // match lists are not retrievable through get_pair_matches after this variant
    return 0;
}

int b200sift_ransac(b200sift_ctx *c, const double *matches, int n, double thr, double *move, int32_t *best)
{
    B200_ARG(c && move && best && n >= 0);
    move[0] = move[1] = 0;
    *best = -1;
    if (n == 0) return 0;
    B200_ARG(matches != nullptr);
    B200_CUDA(cudaSetDevice(c->device));
    size_t cap = c->mA_cap;
    B200_CHECK(ensure(&c->d_mA, &cap, (size_t)n * 32));
    c->mA_cap = cap;
    B200_CUDA(cudaMemcpyAsync(c->d_mA, matches, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
    return launch_ransac(c, (const double *)c->d_mA, n, thr, move, best);
}

// ------------------------------------------------------------------ stage API

int b200sift_gaussian_blur(b200sift_ctx *c, const float *src, int h, int w, double sigma, float *dst, int on_device)
{
    B200_ARG(c && src && dst && h >= 1 && w >= 1 && sigma > 0);
    B200_CUDA(cudaSetDevice(c->device));
    if (on_device) {
        // dense device rows; rows are float4-aligned only when w % 4 == 0 (the
        // strip kernel checks and otherwise the tile kernel runs)
        return launch_blur(c, src, dst, 1, h, w, w, (size_t)h * w, sigma, nullptr, 0, 0, 0, 0);
    }
    const int pitch = (w + 7) & ~7;
    size_t cap = c->up_cap;
    B200_CHECK(ensure(&c->d_up, &cap, (size_t)2 * h * pitch));
    c->up_cap = cap;
    float *d_src = c->d_up, *d_dst = c->d_up + (size_t)h * pitch;
    B200_CUDA(cudaMemcpy2DAsync(d_src, (size_t)pitch * 4, src, (size_t)w * 4, (size_t)w * 4, h,
                                cudaMemcpyHostToDevice, c->stream));
    Timer tm(c);
    B200_CHECK(launch_blur(c, d_src, d_dst, 1, h, w, pitch, (size_t)h * pitch, sigma, nullptr, 0, 0, 0, 0));
    tm.stop();
    B200_CUDA(cudaMemcpy2DAsync(dst, (size_t)w * 4, d_dst, (size_t)pitch * 4, (size_t)w * 4, h,
                                cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(b200::ctx_sync(c));
    return 0;
}

int b200sift_base_image(b200sift_ctx *c, const float *image, int h, int w, double sigma, double assumed_blur,
                        float *out)
{
    B200_ARG(c && image && out && h >= 1 && w >= 1);
    B200_CUDA(cudaSetDevice(c->device));
    if (!(sigma > 0)) sigma = 1.6;
    const int H = 2 * h, W = 2 * w, pitch = (W + 7) & ~7;
    size_t cap = c->in_cap;
    B200_CHECK(ensure(&c->d_in, &cap, (size_t)h * w * 4));
    c->in_cap = cap;
    cap = c->up_cap;
    B200_CHECK(ensure(&c->d_up, &cap, (size_t)2 * H * pitch));
    c->up_cap = cap;
    B200_CUDA(cudaMemcpyAsync(c->d_in, image, (size_t)h * w * 4, cudaMemcpyHostToDevice, c->stream));
    float *d_up = c->d_up, *d_dst = c->d_up + (size_t)H * pitch;
    B200_CHECK(launch_gray_upsample(c, c->d_in, 0, nullptr, (size_t)w * 4, 1, h, w, 1, B200SIFT_F32, d_up, pitch));
    const double d2 = sigma * sigma - (2 * assumed_blur) * (2 * assumed_blur);
    B200_CHECK(launch_blur(c, d_up, d_dst, 1, H, W, pitch, (size_t)H * pitch, sqrt(d2 > 0.01 ? d2 : 0.01), nullptr,
                           0, 0, 0, 0));
    B200_CUDA(cudaMemcpy2DAsync(out, (size_t)W * 4, d_dst, (size_t)pitch * 4, (size_t)W * 4, H,
                                cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(b200::ctx_sync(c));
    return 0;
}

int b200sift_gaussian_pyramid(b200sift_ctx *c, const float *base, int h, int w, int n_oct, const double *sigmas,
                              int n_layers, float *const *out_layers)
{
    B200_ARG(c && base && sigmas && out_layers && n_oct >= 1 && n_layers >= 4);
    B200_CUDA(cudaSetDevice(c->device));
    c->have_results = false;
    B200_CHECK(pyramid_layout(c, 1, h, w, n_oct, n_layers));
    const Pyramid &p = c->pyr;
    B200_CUDA(cudaMemcpy2DAsync(p.layer(0, 0), (size_t)p.pitch[0] * 4, base, (size_t)w * 4, (size_t)w * 4, h,
                                cudaMemcpyHostToDevice, c->stream));
    B200_CHECK(build_octaves(c, sigmas));
    for (int o = 0; o < n_oct; ++o)
        for (int l = 0; l < n_layers; ++l) {
            float *dst = out_layers[o * n_layers + l];
            B200_ARG(dst != nullptr);
            B200_CUDA(cudaMemcpy2DAsync(dst, (size_t)p.w[o] * 4, p.layer(o, l), (size_t)p.pitch[o] * 4,
                                        (size_t)p.w[o] * 4, p.h[o], cudaMemcpyDeviceToHost, c->stream));
        }
    B200_CUDA(b200::ctx_sync(c));
    return 0;
}

int b200sift_dog_pyramid(b200sift_ctx *c, const float *const *layers, int h, int w, int n_oct, int n_layers,
                         float *const *out_dogs)
{
    B200_ARG(c && out_dogs);
    B200_CUDA(cudaSetDevice(c->device));
    c->have_results = false;
    B200_CHECK(upload_pyramid(c, layers, h, w, n_oct, n_layers));
    const Pyramid &p = c->pyr;
    size_t need = 0;
    for (int o = 0; o < n_oct; ++o) need = need > p.img_stride(o) ? need : p.img_stride(o);
    size_t cap = c->dog_cap;
    B200_CHECK(ensure(&c->d_dog, &cap, need));
    c->dog_cap = cap;
    for (int o = 0; o < n_oct; ++o)
        for (int l = 0; l + 1 < n_layers; ++l) {
            float *dst = out_dogs[o * (n_layers - 1) + l];
            B200_ARG(dst != nullptr);
            B200_CHECK(launch_dog(c, p.layer(o, l), p.layer(o, l + 1), c->d_dog, p.img_stride(o)));
            B200_CUDA(cudaMemcpy2DAsync(dst, (size_t)p.w[o] * 4, c->d_dog, (size_t)p.pitch[o] * 4,
                                        (size_t)p.w[o] * 4, p.h[o], cudaMemcpyDeviceToHost, c->stream));
        }
    B200_CUDA(b200::ctx_sync(c));
    return 0;
}

int b200sift_find_extrema(b200sift_ctx *c, const b200sift_params *params, const float *const *layers,
                          const float *const *dog_layers, int h, int w, int n_oct, int n_layers,
                          b200sift_keypoint *kps, int capacity, int32_t *n)
{
    B200_ARG(c && n && capacity >= 0);
    b200sift_params P;
    if (params) P = *params; else b200sift_default_params(&P);
    B200_ARG(n_layers == P.num_intervals + 3);
    B200_CUDA(cudaSetDevice(c->device));
    c->have_results = false;
    B200_CHECK(upload_pyramid(c, layers, h, w, n_oct, n_layers));
    if (dog_layers) B200_CHECK(upload_pyramid_into(c, c->dog_pyr, dog_layers, h, w, n_oct, n_layers - 1));
    B200_CHECK(run_detect(c, P, dog_layers != nullptr));
    fill_stats(c);
    const int n_raw = c->h_counters[CNT_RAW];
    B200_CHECK(run_sort_gather(c, n_raw, 1, /*scan_order=*/1, /*dedupe=*/0, /*convert=*/0, /*with_desc=*/0));
    *n = n_raw;
    if (n_raw > capacity) {
        set_error("find_extrema: %d keypoints exceed the caller's capacity %d", n_raw, capacity);
        return B200SIFT_ECAPACITY;
    }
    if (n_raw > 0) {
        B200_ARG(kps != nullptr);
        B200_CUDA(cudaMemcpyAsync(kps, c->d_kps, sizeof(b200sift_keypoint) * n_raw, cudaMemcpyDeviceToHost,
                                  c->stream));
        B200_CUDA(b200::ctx_sync(c));
    }
    return 0;
}

int b200sift_extrema_candidates(b200sift_ctx *c, const b200sift_params *params, const float *const *layers, int h,
                                int w, int n_oct, int n_layers, int32_t *cand, int capacity, int32_t *n)
{
    B200_ARG(c && n && capacity >= 0);
    b200sift_params P;
    if (params) P = *params; else b200sift_default_params(&P);
    B200_ARG(n_layers == P.num_intervals + 3);
    B200_CUDA(cudaSetDevice(c->device));
    c->have_results = false;
    B200_CHECK(upload_pyramid(c, layers, h, w, n_oct, n_layers));
    B200_CHECK(run_detect(c, P, 0));
    const int nc = c->h_counters[CNT_CAND];
    *n = nc;
    if (nc > capacity) {
        set_error("extrema_candidates: %d candidates exceed the caller's capacity %d", nc, capacity);
        return B200SIFT_ECAPACITY;
    }
    if (nc == 0) return 0;
    std::vector<Candidate> hc(nc);
    B200_CUDA(cudaMemcpyAsync(hc.data(), c->d_cand, sizeof(Candidate) * nc, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(b200::ctx_sync(c));
    // device order is arbitrary (atomic compaction); hand back the reference's scan order
    std::vector<uint64_t> key(nc);
    for (int i = 0; i < nc; ++i)
        key[i] = ((uint64_t)(hc[i].img_o_l & 0xffff) << 32) | hc[i].yx;
    std::vector<int> ord(nc);
    for (int i = 0; i < nc; ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](int a, int b) { return key[a] < key[b]; });
    for (int i = 0; i < nc; ++i) {
        const Candidate &k = hc[ord[i]];
        cand[4 * i] = (k.img_o_l >> 8) & 255;
        cand[4 * i + 1] = k.img_o_l & 255;
        cand[4 * i + 2] = k.yx >> 16;
        cand[4 * i + 3] = k.yx & 0xffff;
    }
    return 0;
}

int b200sift_remove_duplicates(b200sift_ctx *c, b200sift_keypoint *kps, int n, int32_t *n_out)
{
    B200_ARG(c && n_out && n >= 0);
    *n_out = n;
    if (n < 2) return 0;  // sift_impl.py:318-319
    B200_ARG(kps != nullptr);
    B200_CUDA(cudaSetDevice(c->device));
    c->have_results = false;
    // stage as raw keypoints of one image; `order` = position in the caller's list (stable sort)
    std::vector<RawKeypoint> raw(n);
    for (int i = 0; i < n; ++i) {
        raw[i].x = kps[i].x; raw[i].y = kps[i].y; raw[i].size = kps[i].size; raw[i].angle = kps[i].angle;
        raw[i].response = kps[i].response; raw[i].octave_packed = kps[i].octave;
        raw[i].img = 0; raw[i].pad = 0; raw[i].order = (uint64_t)i;
    }
    B200_CHECK(ensure_sparse_for(c, 1, n));
    B200_CUDA(cudaMemcpyAsync(c->d_raw, raw.data(), sizeof(RawKeypoint) * n, cudaMemcpyHostToDevice, c->stream));
    B200_CHECK(run_sort_gather(c, n, 1, 0, 1, 0, 0));
    const int m = c->img_off[1];
    B200_CUDA(cudaMemcpyAsync(kps, c->d_kps, sizeof(b200sift_keypoint) * m, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(b200::ctx_sync(c));
    *n_out = m;
    return 0;
}

int b200sift_descriptors(b200sift_ctx *c, const b200sift_params *params, const b200sift_keypoint *kps, int n,
                         const float *const *layers, int h, int w, int n_oct, int n_layers, float *desc_f32)
{
    B200_ARG(c && n >= 0);
    if (n == 0) return 0;
    B200_ARG(kps && desc_f32);
    b200sift_params P;
    if (params) P = *params; else b200sift_default_params(&P);
    B200_CUDA(cudaSetDevice(c->device));
    c->have_results = false;
    B200_CHECK(upload_pyramid(c, layers, h, w, n_oct, n_layers));
    std::vector<RawKeypoint> raw(n);
    for (int i = 0; i < n; ++i) {
        raw[i].x = kps[i].x; raw[i].y = kps[i].y; raw[i].size = kps[i].size; raw[i].angle = kps[i].angle;
        raw[i].response = kps[i].response; raw[i].octave_packed = kps[i].octave;
        raw[i].img = 0; raw[i].pad = 0; raw[i].order = (uint64_t)i;
    }
    B200_ARG(P.window_width >= 1 && P.desc_bins >= 1 && P.window_width * P.window_width * P.desc_bins <= 1024);
    const size_t dlen = (size_t)P.window_width * P.window_width * P.desc_bins;
    B200_CHECK(ensure_sparse_for(c, 1, n));
    uint8_t *d_out = c->d_raw_desc;               // [raw_cap][128]
    if (dlen > 128) {
        size_t cap = c->misc_cap;
        B200_CHECK(ensure((uint8_t **)&c->d_misc, &cap, (size_t)n * dlen));
        c->misc_cap = cap;
        d_out = (uint8_t *)c->d_misc;
    }
    B200_CUDA(cudaMemcpyAsync(c->d_raw, raw.data(), sizeof(RawKeypoint) * n, cudaMemcpyHostToDevice, c->stream));
    B200_CHECK(run_describe(c, P, c->d_raw, n, /*converted=*/1, d_out, 0));
    std::vector<uint8_t> u8((size_t)n * dlen);
    B200_CUDA(cudaMemcpyAsync(u8.data(), d_out, u8.size(), cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(b200::ctx_sync(c));
    for (size_t i = 0; i < u8.size(); ++i) desc_f32[i] = (float)u8[i];
    return 0;
}

int b200sift_localize(b200sift_ctx *c, const b200sift_params *params, const float *const *layers, int is_dog, int h,
                      int w, int n_oct, int n_layers, int octave_base, const int32_t *cand, int n,
                      b200sift_keypoint *kps, int32_t *final_layer)
{
    B200_ARG(c && n >= 0);
    if (n == 0) return 0;
    B200_ARG(layers && cand && kps && final_layer && n_oct >= 1 && (octave_base < 0 || n_oct == 1));
    b200sift_params P;
    if (params) P = *params; else b200sift_default_params(&P);
    B200_ARG(n_layers == P.num_intervals + (is_dog ? 2 : 3));
    B200_CUDA(cudaSetDevice(c->device));
    c->have_results = false;
    if (is_dog) B200_CHECK(upload_pyramid_into(c, c->dog_pyr, layers, h, w, n_oct, n_layers));
    else B200_CHECK(upload_pyramid(c, layers, h, w, n_oct, n_layers));
    if (octave_base >= 0)
        for (int i = 0; i < n; ++i) B200_ARG(cand[4 * i] == octave_base);
    return run_localize_direct(c, P, is_dog, octave_base >= 0, cand, n, kps, final_layer);
}

int b200sift_orientations(b200sift_ctx *c, const b200sift_params *params, const b200sift_keypoint *kps, int n,
                          int octave, const float *gauss_img, int h, int w, b200sift_keypoint *out, int32_t *counts)
{
    B200_ARG(c && n >= 0);
    if (n == 0) return 0;
    B200_ARG(kps && gauss_img && out && counts && h >= 3 && w >= 3);
    b200sift_params P;
    if (params) P = *params; else b200sift_default_params(&P);
    B200_CUDA(cudaSetDevice(c->device));
    c->have_results = false;
    const float *one[1] = {gauss_img};
    B200_CHECK(upload_pyramid(c, one, h, w, 1, 1));
    return run_orient_direct(c, P, kps, n, octave, out, counts);
}

int b200sift_ratio_match(b200sift_ctx *c, const uint8_t *A, int nA, const uint8_t *B, int nB, int on_device,
                         int ratio_num, int ratio_den, int32_t *ia, int32_t *ib, int32_t *best_d2,
                         int32_t *second_d2, int32_t *n_good)
{
    B200_ARG(c && n_good && nA >= 0 && nB >= 0 && ratio_num > 0 && ratio_den > 0 && ratio_num < 32768 &&
             ratio_den < 32768);
    *n_good = 0;
    if (nA == 0) return 0;
    B200_ARG(A != nullptr && (nB == 0 || B != nullptr));
    B200_CUDA(cudaSetDevice(c->device));
    const uint8_t *dA = A, *dB = B;
    if (!on_device) {
        size_t cap = c->mA_cap;
        B200_CHECK(ensure(&c->d_mA, &cap, (size_t)nA * 128));
        c->mA_cap = cap;
        cap = c->mB_cap;
        B200_CHECK(ensure(&c->d_mB, &cap, (size_t)(nB > 0 ? nB : 1) * 128));
        c->mB_cap = cap;
        B200_CUDA(cudaMemcpyAsync(c->d_mA, A, (size_t)nA * 128, cudaMemcpyHostToDevice, c->stream));
        if (nB) B200_CUDA(cudaMemcpyAsync(c->d_mB, B, (size_t)nB * 128, cudaMemcpyHostToDevice, c->stream));
        dA = c->d_mA;
        dB = c->d_mB;
    }
    size_t cap = c->misc_cap;
    B200_CHECK(ensure((uint8_t **)&c->d_misc, &cap, (size_t)(nA * 5 + 4) * sizeof(int32_t)));
    c->misc_cap = cap;
    int32_t *d_idx = (int32_t *)c->d_misc, *d_b1 = d_idx + nA, *d_b2 = d_b1 + nA, *d_ia = d_b2 + nA, *d_ib = d_ia + nA,
            *d_cnt = d_ib + nA;
    Timer tm(c);
    B200_CHECK(run_match(c, dA, nA, dB, nB, d_idx, d_b1, d_b2));
    B200_CHECK(run_ratio_accept(c, d_idx, d_b1, d_b2, nA, ratio_num, ratio_den, d_ia, d_ib, d_cnt));
    tm.stop();
    int32_t n = 0;
    B200_CUDA(cudaMemcpyAsync(&n, d_cnt, sizeof(n), cudaMemcpyDeviceToHost, c->stream));
    if (best_d2) B200_CUDA(cudaMemcpyAsync(best_d2, d_b1, sizeof(int32_t) * nA, cudaMemcpyDeviceToHost, c->stream));
    if (second_d2) B200_CUDA(cudaMemcpyAsync(second_d2, d_b2, sizeof(int32_t) * nA, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(b200::ctx_sync(c));
    if (n > 0) {
        if (ia) B200_CUDA(cudaMemcpyAsync(ia, d_ia, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
        if (ib) B200_CUDA(cudaMemcpyAsync(ib, d_ib, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
        B200_CUDA(b200::ctx_sync(c));
    }
    *n_good = n;
    return 0;
}

int b200sift_match_grid(b200sift_ctx *c, int32_t *tiles_per_chunk, int32_t *n_chunks)
{
    B200_ARG(c != nullptr);
    if (tiles_per_chunk) *tiles_per_chunk = c->last_tiles_per_chunk;
    if (n_chunks) *n_chunks = c->last_n_chunks;
    return 0;
}

int b200sift_cylindrical_projection(b200sift_ctx *c, const uint8_t *src, int h, int w, int ch, double focal,
                                    uint8_t *dst)
{
    B200_ARG(c && src && dst && h >= 1 && w >= 1 && ch >= 1 && ch <= 4 && focal > 0);
    B200_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)h * w * ch;
    size_t cap = c->mA_cap;
    B200_CHECK(ensure(&c->d_mA, &cap, bytes));
    c->mA_cap = cap;
    cap = c->mB_cap;
    B200_CHECK(ensure(&c->d_mB, &cap, bytes));
    c->mB_cap = cap;
    B200_CUDA(cudaMemcpyAsync(c->d_mA, src, bytes, cudaMemcpyHostToDevice, c->stream));
    B200_CHECK(launch_cyl(c, c->d_mA, h, w, ch, focal, c->d_mB));
    B200_CUDA(cudaMemcpyAsync(dst, c->d_mB, bytes, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(b200::ctx_sync(c));
    return 0;
}

int b200sift_bench_blur(b200sift_ctx *c, int n_img, int h, int w, double sigma, int iters, int flush_l2,
                        float *ms_per_launch)
{
    B200_ARG(c && ms_per_launch && n_img >= 1 && h >= 1 && w >= 1 && iters >= 1 && sigma > 0);
    B200_CUDA(cudaSetDevice(c->device));
    c->have_results = false;
    const int pitch = (w + 7) & ~7;
    const size_t img = (size_t)h * pitch, total = img * n_img;
    size_t cap = c->pyr.capacity_floats;
    float *buf = c->pyr.base;
    B200_CHECK(ensure(&buf, &cap, 2 * total));
    c->pyr.base = buf;
    c->pyr.capacity_floats = cap;
    float *src = buf, *dst = buf + total;
    B200_CUDA(cudaMemsetAsync(src, 0x3c, total * sizeof(float), c->stream));  // finite floats
    const size_t flush_bytes = (size_t)256 << 20;
    if (flush_l2) {
        size_t fc = c->misc_cap;
        B200_CHECK(ensure((uint8_t **)&c->d_misc, &fc, flush_bytes));
        c->misc_cap = fc;
    }
    // warm-up (also uploads the taps)
    for (int i = 0; i < 3; ++i)
        B200_CHECK(launch_blur(c, src, dst, n_img, h, w, pitch, img, sigma, nullptr, 0, 0, 0, 0));
    double acc = 0;
    for (int i = 0; i < iters; ++i) {
        if (flush_l2) B200_CUDA(cudaMemsetAsync(c->d_misc, i & 0xff, flush_bytes, c->stream));
        B200_CUDA(cudaEventRecord(c->ev0, c->stream));
        B200_CHECK(launch_blur(c, src, dst, n_img, h, w, pitch, img, sigma, nullptr, 0, 0, 0, 0));
        B200_CUDA(cudaEventRecord(c->ev1, c->stream));
        B200_CUDA(cudaEventSynchronize(c->ev1));
        float ms = 0;
        B200_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        acc += ms;
    }
    *ms_per_launch = (float)(acc / iters);
    return 0;
}

int b200sift_bench_match(b200sift_ctx *c, const uint8_t *A, int nA, const uint8_t *B, int nB, int top2, int iters,
                         float *ms_per_launch)
{
    B200_ARG(c && ms_per_launch && nA >= 1 && nB >= 1 && iters >= 1 && ((A == nullptr) == (B == nullptr)));
    B200_CUDA(cudaSetDevice(c->device));
    return bench_match_tc(c, A, nA, B, nB, top2, iters, ms_per_launch);
}

}  // extern "C"
