// The step AFTER the SIFT hot path: linear blend of two aligned images and the
// bounding box of the non-black region.
//
// Replaces /root/reference/image_stitching_sift.py:139-153 (pad_image), :156-202 (blend_two_images)
// and the reduction of :208-247 (rectangle_crop: cvtColor + mask + min/max of the coordinates).
// Compiled with --fmad=false: numpy rounds the two products and the sum of
// (1 - alpha) * colA + alpha * colB separately.
#include <math.h>
#include <string.h>
#include "common.cuh"

namespace b200 {

struct BlendGeom {
    int hA, wA, hB, wB;          // images after the dx < 0 swap
    int oyA, oxA, oyB, oxB;      // offset of each image on the canvas (the zero padding of pad_image)
    int HH, WW;                  // canvas
    double overlap_range;
};

// pad_image (:139-153): int(round(move)) with Python's half-to-even rounding; a non-negative move
// pads in front (the image moves right / down), a negative one pads behind.
static inline long py_round(double v) { return (long)rint(v); }

static void blend_geometry(int hA, int wA, int hB, int wB, double dy, const double rm[4], BlendGeom *g)
{
    // :167-169, evaluated left to right in float64
    const double padA_x = ((double)(wB - wA) + rm[0]) - rm[2];
    const double padB_x = rm[0] - rm[2];
    g->overlap_range = (rm[2] - rm[0]) + (double)wA;
    const long mxA = py_round(-padA_x), myA = py_round(-dy);
    const long mxB = py_round(padB_x), myB = py_round(dy);
    g->hA = hA; g->wA = wA; g->hB = hB; g->wB = wB;
    g->oxA = mxA >= 0 ? (int)mxA : 0; g->oyA = myA >= 0 ? (int)myA : 0;
    g->oxB = mxB >= 0 ? (int)mxB : 0; g->oyB = myB >= 0 ? (int)myB : 0;
    const int hA2 = hA + (int)labs(myA), wA2 = wA + (int)labs(mxA);
    const int hB2 = hB + (int)labs(myB), wB2 = wB + (int)labs(mxB);
    g->HH = hA2 > hB2 ? hA2 : hB2;
    g->WW = wA2 > wB2 ? wA2 : wB2;
}

// flags[x] bit 0: column x of canvas A holds a non-zero byte, bit 1: same for canvas B (:186-187)
__global__ void __launch_bounds__(128)
blend_column_flags_kernel(const uint8_t *__restrict__ A, const uint8_t *__restrict__ B, BlendGeom g,
                          uint8_t *__restrict__ flags)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= g.WW) return;
    unsigned f = 0;
    const int xa = x - g.oxA, xb = x - g.oxB;
    if (xa >= 0 && xa < g.wA) {
        unsigned any = 0;
        for (int y = 0; y < g.hA; ++y) {
            const uint8_t *p = A + ((size_t)y * g.wA + xa) * 3;
            any |= p[0] | p[1] | p[2];
        }
        if (any) f |= 1u;
    }
    if (xb >= 0 && xb < g.wB) {
        unsigned any = 0;
        for (int y = 0; y < g.hB; ++y) {
            const uint8_t *p = B + ((size_t)y * g.wB + xb) * 3;
            any |= p[0] | p[1] | p[2];
        }
        if (any) f |= 2u;
    }
    flags[x] = (uint8_t)f;
}

// overlap_counter of :189-193: the number of overlapping columns left of x (one block, any width)
__global__ void __launch_bounds__(1024)
blend_overlap_scan_kernel(const uint8_t *__restrict__ flags, int WW, int32_t *__restrict__ idx)
{
    __shared__ int warp_sum[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < WW; base += 1024) {
        const int x = base + threadIdx.x;
        const int v = (x < WW && flags[x] == 3) ? 1 : 0;
        int s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        if (lane == 31) warp_sum[wid] = s;
        __syncthreads();
        if (wid == 0) {
            int ws = warp_sum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, ws, d);
                if (lane >= d) ws += t;
            }
            warp_sum[lane] = ws;
        }
        __syncthreads();
        const int before = carry + (wid ? warp_sum[wid - 1] : 0) + s - v;  // exclusive
        if (x < WW) idx[x] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
}

// :195-200 + .astype(np.uint8).  STRONG = alpha is a numpy float64 scalar (float64 arithmetic, one
// rounding to float32 on assignment); otherwise alpha is a Python float and numpy computes in float32.
template <bool STRONG>
__global__ void __launch_bounds__(256)
blend_kernel(const uint8_t *__restrict__ A, const uint8_t *__restrict__ B, BlendGeom g,
             const uint8_t *__restrict__ flags, const int32_t *__restrict__ idx, uint8_t *__restrict__ out)
{
    const size_t n = (size_t)g.HH * g.WW;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / g.WW), x = (int)(i - (size_t)y * g.WW);
        const int ya = y - g.oyA, xa = x - g.oxA, yb = y - g.oyB, xb = x - g.oxB;
        const bool inA = ya >= 0 && ya < g.hA && xa >= 0 && xa < g.wA;
        const bool inB = yb >= 0 && yb < g.hB && xb >= 0 && xb < g.wB;
        const uint8_t *pa = A + ((size_t)(inA ? ya : 0) * g.wA + (inA ? xa : 0)) * 3;
        const uint8_t *pb = B + ((size_t)(inB ? yb : 0) * g.wB + (inB ? xb : 0)) * 3;
        const unsigned f = flags[x];
        double alpha = 0.0;
        if (f == 3 && g.overlap_range != 0.0) alpha = (double)idx[x] / g.overlap_range;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float a = inA ? (float)pa[ch] : 0.f, b = inB ? (float)pb[ch] : 0.f;
            float r;
            if (f == 3) {
                if (STRONG) r = (float)((1.0 - alpha) * (double)a + alpha * (double)b);
                else r = (float)(1.0 - alpha) * a + (float)alpha * b;
            } else if (f & 1) {
                r = a;
            } else if (f & 2) {
                r = b;
            } else {
                r = 0.f;
            }
            out[i * 3 + ch] = (uint8_t)(int)r;  // C cast of astype: truncate, wrap modulo 256
        }
    }
}

// rectangle_crop (:224-236): gray = cv2.cvtColor(BGR2GRAY) (15-bit fixed point, sift pyramid.cu), mask =
// gray > threshold, bounding box of the mask.  box = {y_min, y_max, x_min, x_max}
__global__ void __launch_bounds__(256)
crop_bbox_kernel(const uint8_t *__restrict__ img, int h, int w, int thr, int32_t *__restrict__ box)
{
    int y0 = INT_MAX, y1 = -1, x0 = INT_MAX, x1 = -1;
    const size_t n = (size_t)h * w;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint8_t *p = img + i * 3;
        const int gray = (p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15;
        if (gray > thr) {
            const int y = (int)(i / w), x = (int)(i - (size_t)y * w);
            y0 = min(y0, y); y1 = max(y1, y); x0 = min(x0, x); x1 = max(x1, x);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, d));
        y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, d));
        x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, d));
        x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, d));
    }
    if ((threadIdx.x & 31) == 0 && y1 >= 0) {
        atomicMin(&box[0], y0); atomicMax(&box[1], y1); atomicMin(&box[2], x0); atomicMax(&box[3], x1);
    }
}

}  // namespace b200

using namespace b200;

int b200sift_blend_two_images(b200sift_ctx *c, const uint8_t *imgA, int hA, int wA, const uint8_t *imgB, int hB,
                              int wB, double dx, double dy, const double *ref_match, int alpha_float64,
                              uint8_t *out, size_t out_capacity, int32_t *out_h, int32_t *out_w)
{
    B200_ARG(c && imgA && imgB && ref_match && out_h && out_w && hA >= 1 && wA >= 1 && hB >= 1 && wB >= 1);
    double rm[4] = {ref_match[0], ref_match[1], ref_match[2], ref_match[3]};
    if (dx < 0) {  // :160-164: work on (B, A) with the shift negated
        dy = -dy;
        const double t0 = rm[0], t1 = rm[1];
        rm[0] = rm[2]; rm[1] = rm[3]; rm[2] = t0; rm[3] = t1;
        const uint8_t *ti = imgA; imgA = imgB; imgB = ti;
        int t = hA; hA = hB; hB = t;
        t = wA; wA = wB; wB = t;
    }
    BlendGeom g;
    blend_geometry(hA, wA, hB, wB, dy, rm, &g);
    *out_h = g.HH;
    *out_w = g.WW;
    if (!out) return 0;  // size query
    const size_t nA = (size_t)hA * wA * 3, nB = (size_t)hB * wB * 3, nO = (size_t)g.HH * g.WW * 3;
    if (out_capacity < nO) {
        set_error("blend_two_images: output needs %zu bytes, capacity %zu", nO, out_capacity);
        return B200SIFT_EARG;
    }
    B200_CUDA(cudaSetDevice(c->device));
    size_t cap = c->mA_cap;
    B200_CHECK(ensure(&c->d_mA, &cap, nA + nB));
    c->mA_cap = cap;
    cap = c->mB_cap;
    B200_CHECK(ensure(&c->d_mB, &cap, nO));
    c->mB_cap = cap;
    cap = c->mout_cap;
    B200_CHECK(ensure(&c->d_mout, &cap, (size_t)g.WW + ((size_t)g.WW + 3) / 4 + 8));
    c->mout_cap = cap;
    uint8_t *dA = c->d_mA, *dB = c->d_mA + nA, *dO = c->d_mB;
    int32_t *d_idx = c->d_mout;
    uint8_t *d_flags = reinterpret_cast<uint8_t *>(c->d_mout + g.WW);
    B200_CUDA(cudaMemcpyAsync(dA, imgA, nA, cudaMemcpyHostToDevice, c->stream));
    B200_CUDA(cudaMemcpyAsync(dB, imgB, nB, cudaMemcpyHostToDevice, c->stream));
    blend_column_flags_kernel<<<(g.WW + 127) / 128, 128, 0, c->stream>>>(dA, dB, g, d_flags);
    blend_overlap_scan_kernel<<<1, 1024, 0, c->stream>>>(d_flags, g.WW, d_idx);
    const size_t npx = (size_t)g.HH * g.WW;
    const int blocks = (int)((npx + 255) / 256 < (size_t)c->sm_count * 8 ? (npx + 255) / 256 : (size_t)c->sm_count * 8);
    if (alpha_float64) blend_kernel<true><<<blocks, 256, 0, c->stream>>>(dA, dB, g, d_flags, d_idx, dO);
    else blend_kernel<false><<<blocks, 256, 0, c->stream>>>(dA, dB, g, d_flags, d_idx, dO);
    c->launches += 3;
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaMemcpyAsync(out, dO, nO, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(b200::ctx_sync(c));
    return 0;
}

int b200sift_crop_bbox(b200sift_ctx *c, const uint8_t *img, int h, int w, int black_threshold, int32_t *box)
{
    B200_ARG(c && img && box && h >= 1 && w >= 1);
    B200_CUDA(cudaSetDevice(c->device));
    const size_t n = (size_t)h * w * 3;
    size_t cap = c->mA_cap;
    B200_CHECK(ensure(&c->d_mA, &cap, n));
    c->mA_cap = cap;
    cap = c->mout_cap;
    B200_CHECK(ensure(&c->d_mout, &cap, (size_t)4));
    c->mout_cap = cap;
    const int32_t init[4] = {INT32_MAX, -1, INT32_MAX, -1};
    B200_CUDA(cudaMemcpyAsync(c->d_mout, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    B200_CUDA(cudaMemcpyAsync(c->d_mA, img, n, cudaMemcpyHostToDevice, c->stream));
    const size_t npx = (size_t)h * w;
    const int blocks = (int)((npx + 255) / 256 < (size_t)c->sm_count * 8 ? (npx + 255) / 256 : (size_t)c->sm_count * 8);
    crop_bbox_kernel<<<blocks, 256, 0, c->stream>>>(c->d_mA, h, w, black_threshold, c->d_mout);
    c->launches++;
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaMemcpyAsync(box, c->d_mout, sizeof(init), cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(b200::ctx_sync(c));
    return 0;
}
