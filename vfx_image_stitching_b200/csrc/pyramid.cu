// Dense front-end of the SIFT path: grey conversion + 2x upsample, separable
// Gaussian blur (HBM-bound packed ring kernel + generic tile kernel), octave
// decimation, DoG materialisation for the stage API.
//
// Replaces (all in /root/reference): sift_impl.py:27-29 (cv2.cvtColor + float
// cast), :45-56 generate_base_image (cv2.resize INTER_LINEAR + GaussianBlur),
// :82-97 generate_gaussian_images (incremental cv2.GaussianBlur chain +
// INTER_NEAREST decimation), :100-111 generate_DoG_images.
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"

namespace b200 {

// Up to 16 tap sets (centre..R) per CONTEXT, in global memory (b200sift_ctx::d_taps, host mirror and
// sigma cache in b200sift_ctx::taps): contexts that run concurrently on one GPU with different
// sigmas cannot overwrite each other's sets, and an update is an ordinary stream-ordered copy.
// A launch names its set; set 15 is reserved for the stand-alone blur entry point.  The packed ring
// kernel takes its taps as a kernel parameter (uniform-register operands), the tile / tail kernels
// read them from the table.

// cv2.getGaussianKernel(ksize, sigma, CV_32F) with ksize = cvRound(8*sigma+1)|1
// (the CV_32F branch of cv::createGaussianKernels): taps = float(exp(-x^2/2s^2)/sum).
static int gaussian_taps(double sigma, float *taps /*centre..R*/)
{
    int ks = ((int)rint(sigma * 8.0 + 1.0)) | 1;
    int r = ks / 2;
    if (r > kMaxBlurRadius) return -1;
    double tmp[2 * kMaxBlurRadius + 1], sum = 0, s2 = -0.5 / (sigma * sigma);
    for (int i = 0; i < ks; ++i) {
        double x = i - (ks - 1) * 0.5;
        tmp[i] = exp(s2 * x * x);
        sum += tmp[i];
    }
    sum = 1.0 / sum;
    for (int k = 0; k <= r; ++k) taps[k] = (float)(tmp[r + k] * sum);
    return r;
}

static int upload_taps(b200sift_ctx *c, int set, double sigma, int *radius)
{
    TapCache &tc = c->taps;
    if (tc.valid[set] && tc.sigma[set] == sigma) {
        *radius = tc.radius[set];
        return 0;
    }
    float taps[kMaxBlurRadius + 1] = {0};
    int r = gaussian_taps(sigma, taps);
    if (r < 0) {
        set_error("sigma %.3f needs a blur radius > %d", sigma, kMaxBlurRadius);
        return B200SIFT_EARG;
    }
    if (!c->d_taps) B200_CUDA(cudaMalloc((void **)&c->d_taps, sizeof(tc.taps)));
    // The host mirror of a set is rewritten below while an earlier asynchronous copy of the same set
    // could still be reading it: wait for the stream first (rare: only when a sigma changes).
    if (tc.valid[set]) B200_CUDA(b200::ctx_sync(c));
    memcpy(tc.taps[set], taps, sizeof(taps));
    // stream-ordered: later launches of this context see the new set, earlier ones the old
    B200_CUDA(cudaMemcpyAsync(c->d_taps + (size_t)set * (kMaxBlurRadius + 1), tc.taps[set], sizeof(taps),
                              cudaMemcpyHostToDevice, c->stream));
    tc.valid[set] = true;
    tc.sigma[set] = sigma;
    tc.radius[set] = r;
    *radius = r;
    return 0;
}

// cv::borderInterpolate(BORDER_REFLECT_101)
__device__ __forceinline__ int reflect101(int p, int len)
{
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        p = p < 0 ? -p : 2 * (len - 1) - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

// ---------------------------------------------------------------------------
// grey + 2x bilinear upsample: (u8 BGR | u8 grey | f32 grey) -> f32 (2h x 2w)
// cv2.cvtColor fixed point (B*3735 + G*19235 + R*9798 + 16384) >> 15, then
// cv2.resize INTER_LINEAR fx=fy=2: src = (dst+0.5)/2 - 0.5, weights .25/.75,
// replicate clamp, horizontal pass first.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float load_grey(const uint8_t *img, size_t row_stride, int y, int x, int channels,
                                           int dtype)
{
    const uint8_t *row = img + (size_t)y * row_stride;
    if (dtype == B200SIFT_F32) return reinterpret_cast<const float *>(row)[x];
    if (channels == 1) return (float)row[x];
    const uint8_t *p = row + 3 * x;
    return (float)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15);
}

// One CTA converts a 16 x 64 source tile (+1 replicate-clamped halo) to grey in shared memory
// once, then every thread produces 2 x 4 output pixels per step from 3 x 4 grey values: the
// .25/.75 pattern of the exact 2x resize is fixed per output parity, only the image border
// columns / rows differ (weight 1 on the edge pixel, as cv2's clamped coordinates give).
constexpr int kUpH = 16, kUpW = 64;

__global__ void __launch_bounds__(256) gray_upsample_kernel(const uint8_t *__restrict__ in, size_t img_stride_bytes,
                                                            const uint8_t *const *__restrict__ ptrs,
                                                            size_t row_stride, int h, int w, int channels, int dtype,
                                                            float *__restrict__ out, int out_pitch, int vec_ok)
{
    __shared__ __align__(16) float g[kUpH + 2][kUpW + 4];
    const int tid = threadIdx.x;
    const int sy0 = blockIdx.y * kUpH, sx0 = blockIdx.x * kUpW;
    // images either sit at a fixed stride behind `in` or are named one by one in `ptrs`
    const uint8_t *img = ptrs ? ptrs[blockIdx.z] : in + (size_t)blockIdx.z * img_stride_bytes;
    for (int i = tid; i < (kUpH + 2) * (kUpW + 3); i += 256) {
        const int r = i / (kUpW + 3), cc = i - r * (kUpW + 3);
        const int y = min(max(sy0 - 1 + r, 0), h - 1), x = min(max(sx0 - 1 + cc, 0), w - 1);
        g[r][cc] = load_grey(img, row_stride, y, x, channels, dtype);
    }
    __syncthreads();
    float *oimg = out + (size_t)blockIdx.z * (size_t)(2 * h) * out_pitch;
    for (int u = tid; u < kUpH * (kUpW / 2); u += 256) {
        const int i = u / (kUpW / 2), k = u - i * (kUpW / 2);
        const int si = sy0 + i, sj = sx0 + 2 * k;  // source row / first of the two source columns
        if (si >= h || sj >= w) continue;
        // horizontal pass on source rows si-1, si, si+1 (tile rows i, i+1, i+2)
        float H[3][4];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float2 q0 = *reinterpret_cast<const float2 *>(&g[i + r][2 * k]);      // columns sj-1, sj
            const float2 q1 = *reinterpret_cast<const float2 *>(&g[i + r][2 * k + 2]);  // columns sj+1, sj+2
            const float a = q0.x, b = q0.y, c = q1.x, d = q1.y;
            H[r][0] = sj == 0 ? b : __fadd_rn(__fmul_rn(a, 0.25f), __fmul_rn(b, 0.75f));
            H[r][1] = sj >= w - 1 ? b : __fadd_rn(__fmul_rn(b, 0.75f), __fmul_rn(c, 0.25f));
            H[r][2] = __fadd_rn(__fmul_rn(b, 0.25f), __fmul_rn(c, 0.75f));
            H[r][3] = sj + 1 >= w - 1 ? c : __fadd_rn(__fmul_rn(c, 0.75f), __fmul_rn(d, 0.25f));
        }
        float o0[4], o1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            o0[j] = si == 0 ? H[1][j] : __fadd_rn(__fmul_rn(H[0][j], 0.25f), __fmul_rn(H[1][j], 0.75f));
            o1[j] = si >= h - 1 ? H[1][j] : __fadd_rn(__fmul_rn(H[1][j], 0.75f), __fmul_rn(H[2][j], 0.25f));
        }
        float *p0 = oimg + (size_t)(2 * si) * out_pitch + 2 * sj, *p1 = p0 + out_pitch;
        if (vec_ok && sj + 1 < w) {
            *reinterpret_cast<float4 *>(p0) = make_float4(o0[0], o0[1], o0[2], o0[3]);
            *reinterpret_cast<float4 *>(p1) = make_float4(o1[0], o1[1], o1[2], o1[3]);
        } else {
            const int nv = sj + 1 < w ? 4 : 2;
            for (int j = 0; j < nv; ++j) { p0[j] = o0[j]; p1[j] = o1[j]; }
        }
    }
}

int launch_gray_upsample(b200sift_ctx *c, const void *d_in, size_t img_stride_bytes, const void *const *d_ptrs,
                         size_t row_stride, int n_img, int h, int w, int channels, int dtype, float *d_out,
                         int out_pitch)
{
    dim3 grid((w + kUpW - 1) / kUpW, (h + kUpH - 1) / kUpH, n_img);
    const int vec_ok = (out_pitch % 4 == 0) && (((size_t)(2 * h) * out_pitch) % 4 == 0) &&
                       (reinterpret_cast<uintptr_t>(d_out) % 16 == 0);
    gray_upsample_kernel<<<grid, 256, 0, c->stream>>>((const uint8_t *)d_in, img_stride_bytes,
                                                      (const uint8_t *const *)d_ptrs, row_stride, h, w, channels,
                                                      dtype, d_out, out_pitch, vec_ok);
    c->launches++;
    B200_CUDA(cudaGetLastError());
    return 0;
}

#include "blur_ring.cuh"

// Generic tile blur: any radius <= kMaxBlurRadius, any (tiny) image.  32x32 output tile, halo tile in
// shared memory, k0*c + sum k[k]*(a[+k]+a[-k]) with fmaf.  All 256 threads of the CTA call it together.
__device__ __forceinline__ void blur_tile(const float *__restrict__ src, float *__restrict__ dst,
                                          float *__restrict__ dst2, int h, int w, int pitch, int h2, int w2,
                                          int pitch2, int R, const float *__restrict__ taps, int x0, int y0, float *smem)
{
    constexpr int T = 32;
    const int IW = T + 2 * R;
    float *in_s = smem;            // [IW][IW]
    float *hs = smem + IW * IW;    // [IW][T]
    for (int i = threadIdx.x; i < IW * IW; i += 256) {
        const int yy = i / IW, xx = i - yy * IW;
        // L2 load: in the tail kernel the source layer was written by other CTAs of this launch
        in_s[i] = __ldcg(&src[(size_t)reflect101(y0 - R + yy, h) * pitch + reflect101(x0 - R + xx, w)]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < IW * T; i += 256) {
        const int yy = i / T, cx = i - yy * T;
        const float *p = in_s + yy * IW + cx + R;
        float acc = taps[0] * p[0];
        for (int k = 1; k <= R; ++k) acc = fmaf(taps[k], p[k] + p[-k], acc);
        hs[i] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < T * T; i += 256) {
        const int r = i / T, cx = i - r * T;
        const int y = y0 + r, x = x0 + cx;
        if (y >= h || x >= w) continue;
        const float *p = hs + (r + R) * T + cx;
        float acc = taps[0] * p[0];
        for (int k = 1; k <= R; ++k) acc = fmaf(taps[k], p[k * T] + p[-k * T], acc);
        dst[(size_t)y * pitch + x] = acc;
        if (dst2 && !((y | x) & 1) && (y >> 1) < h2 && (x >> 1) < w2)
            dst2[(size_t)(y >> 1) * pitch2 + (x >> 1)] = acc;
    }
    __syncthreads();  // the shared tiles may be refilled by the caller's next tile
}

__global__ void __launch_bounds__(256)
blur_tile_kernel(const float *__restrict__ src, float *__restrict__ dst, float *__restrict__ dst2, int h, int w,
                 int pitch, size_t img_stride, int h2, int w2, int pitch2, size_t img_stride2, int R,
                 const float *__restrict__ taps)
{
    extern __shared__ __align__(16) float smem[];
    src += (size_t)blockIdx.z * img_stride;
    dst += (size_t)blockIdx.z * img_stride;
    if (dst2) dst2 += (size_t)blockIdx.z * img_stride2;
    blur_tile(src, dst, dst2, h, w, pitch, h2, w2, pitch2, R, taps, blockIdx.x * 32, blockIdx.y * 32, smem);
}

// ---------------------------------------------------------------------------
// Pyramid tail: ALL layers of ALL small octaves (<= 128 px wide) in ONE launch, one CTA per image.
// With one kernel per layer the small octaves are pure launch latency (30 dependent launches of a
// few microseconds each for a 1024x768 base).  Here the whole octave image lives in shared memory:
//   A[h][w+2Rm]   current layer with a reflected x-halo      (row pass reads contiguous spans)
//   B[h+2Rm][w]   row-filtered layer with a reflected y-halo (column pass reads contiguous spans)
// Each thread produces 4 adjacent outputs per pass from 4+2R shared-memory values; the column pass
// writes the new layer to HBM, back into A for the next blur of the chain (sift_impl.py:90-92) and,
// for layer n_layers-3, its [::2, ::2] decimation as layer 0 of the next octave (:95-96).
// Arithmetic: k0*c + sum k[k]*(a[+k]+a[-k]), fmaf, BORDER_REFLECT_101.
// ---------------------------------------------------------------------------
struct TailArgs {
    float *base;                 // pyramid allocation
    const float *taps;           // the context's tap table (set l = layer l)
    size_t oct_off[kMaxOctaves]; // float offset of octave o
    int h[kMaxOctaves], w[kMaxOctaves], pitch[kMaxOctaves];
    int radius[kMaxLayers];      // per layer (tap set index = layer)
    int n_img, n_oct, n_layers, o_tail, r_max;
};
constexpr int kTailThreads = 512;

// One blur of the chain on the shared-memory resident octave: A (x-halo) -> B (y-halo) -> A, HBM.
// R is a template parameter so that the tap loops unroll into straight LDS / FADD / FFMA runs.
// Lanes past the right / bottom edge compute on whatever lies there (inside the allocation, see
// tail_smem_bytes) and only the stores are predicated.
// Fixed shared-memory pitches (tail octaves are <= 128 px wide, radii <= 16): every tap of both
// passes is an LDS with an immediate offset, and the flattened (row, unit) loops use shifts
// (lw = log2 of the width rounded up to a power of two) instead of divisions.
constexpr int kTailRm = 16;                 // halo reserved on each side of A and B
constexpr int kTailPA = 128 + 2 * kTailRm;  // A[h][kTailPA], interior starts at column kTailRm
constexpr int kTailPB = 128;                // B[h + 2*kTailRm + 3][kTailPB], interior starts at row kTailRm

template <int R>
__device__ __forceinline__ void tail_layer(float *A, float *B, const float *__restrict__ taps, int h, int w, int lw,
                                           float *__restrict__ dst, int pitch, float *__restrict__ dst2, int h2,
                                           int w2, int pitch2)
{
    constexpr int Rm = kTailRm, PA = kTailPA, PB = kTailPB;
    const int tid = threadIdx.x;
    float t[R + 1];
#pragma unroll
    for (int k = 0; k <= R; ++k) t[k] = taps[k];
    // reflected x-halo of A
    for (int i = tid; i < h * 2 * R; i += kTailThreads) {
        const int y = i / (2 * R), k = i - y * (2 * R);
        const int x = k < R ? k - R : w + (k - R);  // -R..-1, w..w+R-1
        A[y * PA + Rm + x] = A[y * PA + Rm + reflect101(x, w)];
    }
    __syncthreads();
    // row pass: 4 adjacent x per thread, A -> B interior rows
    const int lxb = lw >= 2 ? lw - 2 : 0;  // log2 of the (padded) number of 4-column units per row
    for (int i = tid; i < (h << lxb); i += kTailThreads) {
        const int y = i >> lxb, x0 = (i & ((1 << lxb) - 1)) * 4;
        if (x0 >= w) continue;
        const float *p = A + y * PA + Rm + x0 - R;
        float v[4 + 2 * R];
#pragma unroll
        for (int q = 0; q < 4 + 2 * R; ++q) v[q] = p[q];
        float acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc[j] = t[0] * v[j + R];
#pragma unroll
            for (int k = 1; k <= R; ++k) acc[j] = fmaf(t[k], v[j + R + k] + v[j + R - k], acc[j]);
        }
        // columns past w land in B's padding (PB >= w rounded up to 4) and are never read back
        *reinterpret_cast<float4 *>(B + (y + Rm) * PB + x0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
    __syncthreads();
    // reflected y-halo of B
    for (int i = tid; i < ((2 * R) << lw); i += kTailThreads) {
        const int k = i >> lw, x = i & ((1 << lw) - 1);
        if (x >= w) continue;
        const int y = k < R ? k - R : h + (k - R);
        B[(y + Rm) * PB + x] = B[(reflect101(y, h) + Rm) * PB + x];
    }
    __syncthreads();
    // column pass: 4 adjacent y per thread, B -> HBM layer, A interior, decimated seed
    const int yb = (h + 3) >> 2;
    for (int i = tid; i < (yb << lw); i += kTailThreads) {
        const int yblk = i >> lw, x = i & ((1 << lw) - 1), y0 = yblk * 4;
        if (x >= w) continue;
        const float *p = B + (y0 + Rm - R) * PB + x;
        float v[4 + 2 * R];
#pragma unroll
        for (int q = 0; q < 4 + 2 * R; ++q) v[q] = p[q * PB];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float acc = t[0] * v[j + R];
#pragma unroll
            for (int k = 1; k <= R; ++k) acc = fmaf(t[k], v[j + R + k] + v[j + R - k], acc);
            const int y = y0 + j;
            if (y < h) {
                A[y * PA + Rm + x] = acc;
                dst[(size_t)y * pitch + x] = acc;
                if (dst2 && !((y | x) & 1) && (y >> 1) < h2 && (x >> 1) < w2)
                    dst2[(size_t)(y >> 1) * pitch2 + (x >> 1)] = acc;
            }
        }
    }
    __syncthreads();
}

// bytes of shared memory for a first tail octave of h0 rows: A, B (with 3 rows of slack for the
// unpredicated reads of the column pass)
static size_t tail_smem_bytes(int h0)
{
    return ((size_t)h0 * kTailPA + (size_t)(h0 + 2 * kTailRm + 3) * kTailPB) * sizeof(float);
}

__global__ void __launch_bounds__(kTailThreads) pyramid_tail_kernel(const __grid_constant__ TailArgs a)
{
    extern __shared__ __align__(16) float smem[];
    const int img = blockIdx.x, tid = threadIdx.x;
    const int h0 = a.h[a.o_tail];
    float *A = smem;                // [h0][kTailPA]
    float *B = smem + h0 * kTailPA; // [h0 + 2*kTailRm + 3][kTailPB]
    for (int o = a.o_tail; o < a.n_oct; ++o) {
        const int h = a.h[o], w = a.w[o], pitch = a.pitch[o];
        const size_t istride = (size_t)h * pitch;
        int lw = 0;
        while ((1 << lw) < w) ++lw;
        // layer 0 of this octave -> A interior (written by the previous kernel, or by this CTA below)
        {
            const float *src = a.base + a.oct_off[o] + (size_t)img * istride;
            for (int i = tid; i < (h << lw); i += kTailThreads) {
                const int y = i >> lw, x = i & ((1 << lw) - 1);
                if (x < w) A[y * kTailPA + kTailRm + x] = __ldcg(&src[(size_t)y * pitch + x]);
            }
        }
        __syncthreads();
        for (int l = 1; l < a.n_layers; ++l) {
            const int R = a.radius[l];
            const float *taps = a.taps + (size_t)l * (kMaxBlurRadius + 1);
            float *dst = a.base + a.oct_off[o] + ((size_t)l * a.n_img + img) * istride;
            float *dst2 = nullptr;
            int h2 = 0, w2 = 0, pitch2 = 0;
            if (l == a.n_layers - 3 && o + 1 < a.n_oct) {
                h2 = a.h[o + 1]; w2 = a.w[o + 1]; pitch2 = a.pitch[o + 1];
                dst2 = a.base + a.oct_off[o + 1] + (size_t)img * ((size_t)h2 * pitch2);
            }
            switch (R) {
#define B200_TAIL_CASE(RR) case RR: tail_layer<RR>(A, B, taps, h, w, lw, dst, pitch, dst2, h2, w2, pitch2); break;
                B200_TAIL_CASE(1) B200_TAIL_CASE(2) B200_TAIL_CASE(3) B200_TAIL_CASE(4) B200_TAIL_CASE(5)
                B200_TAIL_CASE(6) B200_TAIL_CASE(7) B200_TAIL_CASE(8) B200_TAIL_CASE(9) B200_TAIL_CASE(10)
                B200_TAIL_CASE(11) B200_TAIL_CASE(12) B200_TAIL_CASE(13) B200_TAIL_CASE(14) B200_TAIL_CASE(15)
                B200_TAIL_CASE(16)
#undef B200_TAIL_CASE
            default: break;  // host never routes other radii here
            }
        }
        __threadfence_block();
    }
}

// Resident CTAs per SM of every ring instantiation (a property of the kernel and of sm_100, the
// same on every device): filled once by pyramid_init_device under the library's init lock.
template <int R, int STAGES>
struct RingOcc { static int occ; };
template <int R, int STAGES>
int RingOcc<R, STAGES>::occ = 1;

template <int R, int STAGES>
static int ring_setup()
{
    const size_t smem = RingCfg<R, STAGES>::smem;
    B200_CUDA(cudaFuncSetAttribute(blur_ring_kernel<R, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_CUDA(cudaFuncSetAttribute(blur_ring_kernel<R, STAGES>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int occ = 0;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, blur_ring_kernel<R, STAGES>, kRingThreads, smem));
    RingOcc<R, STAGES>::occ = occ < 1 ? 1 : occ;
    return 0;
}

// input stages (cp.async batches in flight + 1) per radius; build-time overridable for sweeps
#ifndef B200SIFT_RING_S_SMALL
#define B200SIFT_RING_S_SMALL 3
#endif
#ifndef B200SIFT_RING_S_10
#define B200SIFT_RING_S_10 3
#endif
#ifndef B200SIFT_RING_S_13
#define B200SIFT_RING_S_13 2
#endif
template <int R>
struct RingDepth { static constexpr int S = (R >= 12) ? B200SIFT_RING_S_13 : (R >= 10) ? B200SIFT_RING_S_10 : B200SIFT_RING_S_SMALL; };

constexpr size_t kTileSmemMax = (size_t)((32 + 2 * kMaxBlurRadius) * (32 + 2 * kMaxBlurRadius) +
                                         (32 + 2 * kMaxBlurRadius) * 32) * sizeof(float);
constexpr size_t kTailSmemMax = 200 * 1024;

// Function attributes are per DEVICE: called once for every device a context is created on
// (b200sift_create, under the init lock), never from a launch path.
int pyramid_init_device()
{
    B200_CHECK((ring_setup<5, RingDepth<5>::S>()));
    B200_CHECK((ring_setup<6, RingDepth<6>::S>()));
    B200_CHECK((ring_setup<8, RingDepth<8>::S>()));
    B200_CHECK((ring_setup<10, RingDepth<10>::S>()));
    B200_CHECK((ring_setup<13, RingDepth<13>::S>()));
    B200_CUDA(cudaFuncSetAttribute(blur_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmemMax));
    B200_CUDA(cudaFuncSetAttribute(pyramid_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailSmemMax));
    return 0;
}

template <int R, int STAGES>
static int launch_ring_s(b200sift_ctx *c, const float *src, float *dst, float *dst2, int n_img, int h, int w,
                         int pitch, size_t img_stride, int h2, int w2, int pitch2, size_t img_stride2, int tapset,
                         int seg)
{
    dim3 grid((w + kRingW - 1) / kRingW, (h + seg - 1) / seg, n_img);
    BlurTaps<R> taps;
    memcpy(taps.t, c->taps.taps[tapset], sizeof(taps.t));
    blur_ring_kernel<R, STAGES><<<grid, kRingThreads, RingCfg<R, STAGES>::smem,
                                  c->blur_stream ? c->blur_stream : c->stream>>>(
        src, dst, dst2, h, w, pitch, img_stride, h2, w2, pitch2, img_stride2, seg, taps);
    c->launches++;
    B200_CUDA(cudaGetLastError());
    return 0;
}

// Grid shape.  A CTA marches down `seg` rows of one 256-column strip and re-filters 8*ceil(2R/8) halo
// rows per segment, so segments should be long -- but a small problem must still fill the machine:
//   * 256-row segments (halo work 2R/256) whenever they give at least one full wave of CTAs;
//   * otherwise ONE wave: floor(slots / strip-columns) segments per column, at least 32 rows each (the
//     18-image pyramid octaves: 432 CTAs on 444 slots, 128-row segments).
// Three input stages (two batches in flight per CTA) except for R = 13, whose 48-row ring leaves room
// for two.  Two stages for R <= 8 would fit a fourth CTA per SM but measured 4-8 % slower
// (8 x 6144 x 8192, 256-row segments: 564 vs 523 us at R = 5).
template <int R>
static int launch_ring(b200sift_ctx *c, const float *src, float *dst, float *dst2, int n_img, int h, int w,
                       int pitch, size_t img_stride, int h2, int w2, int pitch2, size_t img_stride2, int tapset)
{
    constexpr int SDEEP = RingDepth<R>::S;
    const int cols = ((w + kRingW - 1) / kRingW) * n_img;
    const int slots = c->sm_count * RingOcc<R, SDEEP>::occ;
    const int n1 = slots / cols;
    int seg = 256;
    if (n1 >= 1 && (long long)cols * ((h + 255) / 256) < slots) {  // 256-row segments would not even fill one wave
        seg = (h + n1 - 1) / n1;
        seg = ((seg + kRingBR - 1) / kRingBR) * kRingBR;
        if (seg < 32) seg = 32;
    }
#ifdef B200SIFT_RING_SEG
    seg = B200SIFT_RING_SEG;   // build-time sweep of the segment height
#endif
    return launch_ring_s<R, SDEEP>(c, src, dst, dst2, n_img, h, w, pitch, img_stride, h2, w2, pitch2, img_stride2,
                                   tapset, seg);
}

static int launch_blur_set(b200sift_ctx *c, const float *src, float *dst, int n_img, int h, int w, int pitch,
                           size_t img_stride, int R, int tapset, float *dst2, int h2, int w2, int pitch2,
                           size_t img_stride2)
{
    // ring kernel: 16 B loads from src, 8 B stores to dst, both row-aligned; the radii of the
    // reference's sigma chain (sift_impl.py:66-79 with sigma = 1.6, num_intervals = 3, and the base blur)
    const bool ring_ok = (w >= 96) && (h >= 32) && (pitch % 4 == 0) &&
                         ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && (img_stride % 4 == 0);
    if (ring_ok) {
        switch (R) {
#define B200_RING(RR)                                                                                         \
    case RR:                                                                                                  \
        return launch_ring<RR>(c, src, dst, dst2, n_img, h, w, pitch, img_stride, h2, w2, pitch2, img_stride2, \
                               tapset);
            B200_RING(5)
            B200_RING(6)
            B200_RING(8)
            B200_RING(10)
            B200_RING(13)
#undef B200_RING
            default: break;
        }
    }
    // any other radius / tiny or unaligned image: generic tile kernel
    const int IW = 32 + 2 * R;
    const size_t smem = (size_t)(IW * IW + IW * 32) * sizeof(float);
    dim3 grid((w + 31) / 32, (h + 31) / 32, n_img);
    blur_tile_kernel<<<grid, 256, smem, c->blur_stream ? c->blur_stream : c->stream>>>(
        src, dst, dst2, h, w, pitch, img_stride, h2, w2, pitch2, img_stride2, R,
        c->d_taps + (size_t)tapset * (kMaxBlurRadius + 1));
    c->launches++;
    B200_CUDA(cudaGetLastError());
    return 0;
}

int launch_blur(b200sift_ctx *c, const float *src, float *dst, int n_img, int h, int w, int pitch, size_t img_stride,
                double sigma, float *dst2, int h2, int w2, int pitch2, size_t img_stride2)
{
    int R;
    B200_CHECK(upload_taps(c, 15, sigma, &R));
    return launch_blur_set(c, src, dst, n_img, h, w, pitch, img_stride, R, 15, dst2, h2, w2, pitch2, img_stride2);
}

// ---------------------------------------------------------------------------
// pyramid layout + octave builder
// ---------------------------------------------------------------------------
int pyramid_layout(b200sift_ctx *c, int n_img, int h0, int w0, int n_oct, int n_layers)
{
    c->oct_events_valid = false;
    return pyramid_layout_into(c->pyr, n_img, h0, w0, n_oct, n_layers);
}

int pyramid_layout_into(Pyramid &p, int n_img, int h0, int w0, int n_oct, int n_layers)
{
    B200_ARG(n_img >= 1 && h0 >= 1 && w0 >= 1);
    B200_ARG(n_oct >= 1 && n_oct <= kMaxOctaves);
    B200_ARG(n_layers >= 1 && n_layers <= kMaxLayers);
    p.n_img = n_img;
    p.n_oct = n_oct;
    p.n_layers = n_layers;
    size_t off = 0;
    int h = h0, w = w0;
    for (int o = 0; o < n_oct; ++o) {
        if (h < 1 || w < 1) {
            set_error("octave %d of a %dx%d base image is empty", o, h0, w0);
            return B200SIFT_EARG;
        }
        p.h[o] = h;
        p.w[o] = w;
        p.pitch[o] = (w + 7) & ~7;
        p.oct_off[o] = off;
        off += (size_t)n_layers * n_img * h * p.pitch[o];
        off = (off + 63) & ~(size_t)63;
        h /= 2;
        w /= 2;
    }
    p.floats = off;
    if (off > p.capacity_floats || !p.base) {
        float *nb = p.base;
        size_t cap = p.capacity_floats;
        B200_CHECK(ensure(&nb, &cap, off));
        p.base = nb;
        p.capacity_floats = cap;
    }
    return 0;
}

// generate_gaussian_images (sift_impl.py:82-97): layer 0 of octave 0 must be
// present; builds every other layer.  sigmas[l] is the incremental sigma of
// layer l (index 0 unused).
int build_octaves(b200sift_ctx *c, const double *sigmas)
{
    Pyramid &p = c->pyr;
    int R[kMaxLayers];
    for (int l = 1; l < p.n_layers; ++l) B200_CHECK(upload_taps(c, l, sigmas[l], &R[l]));
    // octaves of <= 128 x 128 px go to the tail kernel (one launch, one CTA per image) when the
    // image + halos fit in shared memory
    int o_tail = p.n_oct;
    for (int o = 1; o < p.n_oct; ++o)
        if (p.w[o] <= 128 && p.h[o] <= 160) { o_tail = o; break; }
    int max_r = 0;
    size_t tail_smem = 0;
    if (o_tail < p.n_oct) {
        for (int l = 1; l < p.n_layers; ++l) max_r = R[l] > max_r ? R[l] : max_r;
        const int h0 = p.h[o_tail], w0 = p.w[o_tail];
        (void)w0;
        tail_smem = tail_smem_bytes(h0);
        if (tail_smem > kTailSmemMax || max_r > kTailRm) o_tail = p.n_oct;
    }
    // Critical path of the pyramid: layers 1..n-3 of octave o, whose last one seeds octave o+1
    // (sift_impl.py:95-96).  The remaining layers of octave o feed nothing downstream but the
    // extrema scan, so they go to a second stream and overlap the (small, latency-bound) blurs of
    // the following octaves.
    const int l_seed = p.n_layers - 3;
    const bool split = l_seed >= 1 && l_seed < p.n_layers - 1 && p.n_oct > 1;
    bool side_used = false;
    for (int o = 0; o < o_tail; ++o) {
        for (int l = 1; l < p.n_layers; ++l) {
            float *dst2 = nullptr;
            int h2 = 0, w2 = 0, pitch2 = 0;
            size_t is2 = 0;
            if (l == l_seed && o + 1 < p.n_oct) {
                dst2 = p.layer(o + 1, 0);
                h2 = p.h[o + 1];
                w2 = p.w[o + 1];
                pitch2 = p.pitch[o + 1];
                is2 = p.img_stride(o + 1);
            }
            const bool on_side = split && l > l_seed;
            if (on_side && l == l_seed + 1) {
                B200_CUDA(cudaEventRecord(c->ev_seed, c->stream));
                B200_CUDA(cudaStreamWaitEvent(c->blur_side_stream, c->ev_seed, 0));
                side_used = true;
            }
            c->blur_stream = on_side ? c->blur_side_stream : nullptr;
            const int rc = launch_blur_set(c, p.layer(o, l - 1), p.layer(o, l), p.n_img, p.h[o], p.w[o], p.pitch[o],
                                           p.img_stride(o), R[l], l, dst2, h2, w2, pitch2, is2);
            tl_mark(on_side ? c->blur_side_stream : c->stream, "%s oct %d layer %d", on_side ? "blur2" : "main ", o, l);
            c->blur_stream = nullptr;
            B200_CHECK(rc);
        }
        // all layers of octave o are complete
        B200_CUDA(cudaEventRecord(c->ev_oct[o], split ? c->blur_side_stream : c->stream));
    }
    if (o_tail < p.n_oct) {
        TailArgs a;
        a.base = p.base;
        a.taps = c->d_taps;
        for (int o = 0; o < p.n_oct; ++o) {
            a.oct_off[o] = p.oct_off[o];
            a.h[o] = p.h[o]; a.w[o] = p.w[o]; a.pitch[o] = p.pitch[o];
        }
        for (int l = 0; l < kMaxLayers; ++l) a.radius[l] = (l >= 1 && l < p.n_layers) ? R[l] : 0;
        a.n_img = p.n_img; a.n_oct = p.n_oct; a.n_layers = p.n_layers; a.o_tail = o_tail; a.r_max = max_r;
        pyramid_tail_kernel<<<p.n_img, kTailThreads, tail_smem, c->stream>>>(a);
        B200_CUDA(cudaGetLastError());
        c->launches++;
        tl_mark(c->stream, "main  tail octaves %d..%d", o_tail, p.n_oct - 1);
        for (int o = o_tail; o < p.n_oct; ++o) B200_CUDA(cudaEventRecord(c->ev_oct[o], c->stream));
    }
    if (side_used) {  // later work on the main stream sees the whole pyramid
        B200_CUDA(cudaEventRecord(c->ev_blur_side, c->blur_side_stream));
        B200_CUDA(cudaStreamWaitEvent(c->stream, c->ev_blur_side, 0));
    }
    c->pyr_o_tail = o_tail < p.n_oct ? o_tail : 0;
    c->oct_events_valid = true;
    return 0;
}

int base_blur(b200sift_ctx *c, const float *d_up, double sigma_diff)
{
    Pyramid &p = c->pyr;
    int R;
    B200_CHECK(upload_taps(c, 0, sigma_diff, &R));
    return launch_blur_set(c, d_up, p.layer(0, 0), p.n_img, p.h[0], p.w[0], p.pitch[0], p.img_stride(0), R, 0,
                           nullptr, 0, 0, 0, 0);
}

// second - first (sift_impl.py:109), float4 where possible
__global__ void __launch_bounds__(256) dog_kernel(const float *__restrict__ a, const float *__restrict__ b,
                                                  float *__restrict__ out, size_t n)
{
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * 256;
    for (; i < n; i += stride) out[i] = __fsub_rn(b[i], a[i]);
}

int launch_dog(b200sift_ctx *c, const float *a, const float *b, float *out, size_t n)
{
    int blocks = (int)((n + 255) / 256);
    if (blocks > c->sm_count * 16) blocks = c->sm_count * 16;
    if (blocks < 1) blocks = 1;
    dog_kernel<<<blocks, 256, 0, c->stream>>>(a, b, out, n);
    c->launches++;
    B200_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200
