// Separable Gaussian blur, strip kernel (included by pyramid.cu).
//
// Replaces cv2.GaussianBlur(img,(0,0),sigma) on float32 (/root/reference/sift_impl.py:56,91).
//
// One CTA owns a 256-column strip of `seg_rows` output rows of one image and marches down it in
// batches of 8 rows:
//   HBM --cp.async 16 B, kStages batches in flight--> in_s[stage][8][256+2RP]   raw rows + x halo
//   row pass   : warp <-> batch row, 4 adjacent outputs per thread from float4 LDS -> hb[2][8][256]
//   column pass: thread <-> column; the last 2R row-filtered values of the column stay in
//                registers (sliding window), 8 new ones come from hb, 8 outputs go to HBM
// The two passes are software-pipelined (row pass of batch b next to the column pass of batch
// b-1), so a batch costs ONE block barrier.
// Every input float is read once from HBM (+ 2RP/256 x-halo and 2R/seg_rows y-halo, both L2 hits)
// and every output written once: 8 B per pixel, the roofline figure of DESIGN.md.  With ~700 ns
// of HBM latency the kernel needs ~31 KB in flight per SM to reach the measured 6.5 TB/s; the
// cp.async ring keeps two 9 KB batches per resident CTA in flight.
// The taps are a __grid_constant__ kernel parameter: they sit in the constant bank at fixed
// offsets, so every FFMA reads its tap as an immediate-address constant operand (no tap
// registers, no loads) -- the kernel is FP32-issue bound for the wide kernels (27 taps: 54
// FMA-class instructions per pixel against 8 bytes).
// dst2 (optional) receives the [::2, ::2] decimation that seeds the next octave
// (sift_impl.py:95-96), which saves a separate pass over layer 3.
// Arithmetic: k0*c + sum_k k[k]*(a[+k] + a[-k]) in float32 (fmaf), rows then columns,
// BORDER_REFLECT_101: the structure of OpenCV's symmetric separable float filter.
#pragma once

constexpr int kStripW = 256;
constexpr int kStripBR = 8;
constexpr int kStripStages = 3;

template <int R>
struct BlurTaps {
    float t[R + 1];  // centre .. R
};

__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gmem_src)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template <int R>
constexpr size_t strip_smem_bytes()
{
    constexpr int RP = (R + 3) & ~3;
    return (size_t)(kStripStages * kStripBR * (kStripW + 2 * RP) + 2 * kStripBR * kStripW) * sizeof(float);
}

template <int R>
__global__ void __launch_bounds__(256, 3)
blur_strip_kernel(const float *__restrict__ src, float *__restrict__ dst, float *__restrict__ dst2, int h, int w,
                  int pitch, size_t img_stride, int h2, int w2, int pitch2, size_t img_stride2, int seg_rows,
                  const __grid_constant__ BlurTaps<R> taps)
{
    constexpr int TW = kStripW, BR = kStripBR, S = kStripStages;
    constexpr int RP = (R + 3) & ~3;
    constexpr int INW = TW + 2 * RP;
    static_assert(BR == 8 && TW == 256, "fill mapping assumes 8 rows x 64 float4 = 2 per thread");
    extern __shared__ __align__(16) float smem[];
    float *in_s = smem;                 // [S][BR][INW]
    float *hb = smem + S * BR * INW;    // [2][BR][TW] row-filtered values, double buffered

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * TW;
    const int ys = blockIdx.y * seg_rows;
    const int ye = min(ys + seg_rows, h);
    src += (size_t)blockIdx.z * img_stride;
    dst += (size_t)blockIdx.z * img_stride;
    if (dst2) dst2 += (size_t)blockIdx.z * img_stride2;

    // Asynchronous fill of one stage.  Body columns [x0, x0+TW) are two float4 per thread and never
    // need a reflection in x (except past the right image edge); the 2*RP halo columns are handled
    // by the first 4*RP threads only, so that the scalar BORDER_REFLECT_101 path of the first / last
    // strip does not drag every warp through divergent code.
    constexpr int HQ = RP / 4;               // float4 per halo side and row
    constexpr int NHALO = BR * 2 * HQ;       // halo float4 slots per batch (<= 64)
    const int b_row0 = tid >> 6, b_c4 = tid & 63;          // body slot i: row = b_row0 + 4*i
    const int b_x = x0 + 4 * b_c4;
    const bool b_in = b_x + 3 < w;
    const int h_row = tid / (2 * HQ), h_k = tid - h_row * (2 * HQ);
    const int h_x = h_k < HQ ? x0 - RP + 4 * h_k : x0 + TW + 4 * (h_k - HQ);
    const int h_off = h_k < HQ ? 4 * h_k : RP + TW + 4 * (h_k - HQ);
    const bool h_in = (h_x >= 0) && (h_x + 3 < w);
    auto issue = [&](int yb, int stage) {
        float *st = in_s + stage * (BR * INW);
        const bool rows_in = (yb >= 0) && (yb + BR <= h);  // CTA-uniform
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int row = b_row0 + 4 * i;
            const int y = rows_in ? yb + row : reflect101(yb + row, h);
            const float *p = src + (size_t)y * pitch;
            float *d = st + row * INW + RP + 4 * b_c4;
            if (b_in) {
                cp_async16(d, p + b_x);
            } else {
                float4 v;
                v.x = p[reflect101(b_x, w)];
                v.y = p[reflect101(b_x + 1, w)];
                v.z = p[reflect101(b_x + 2, w)];
                v.w = p[reflect101(b_x + 3, w)];
                *reinterpret_cast<float4 *>(d) = v;
            }
        }
        if (tid < NHALO) {
            const int y = rows_in ? yb + h_row : reflect101(yb + h_row, h);
            const float *p = src + (size_t)y * pitch;
            float *d = st + h_row * INW + h_off;
            if (h_in) {
                cp_async16(d, p + h_x);
            } else {
                float4 v;
                v.x = p[reflect101(h_x, w)];
                v.y = p[reflect101(h_x + 1, w)];
                v.z = p[reflect101(h_x + 2, w)];
                v.w = p[reflect101(h_x + 3, w)];
                *reinterpret_cast<float4 *>(d) = v;
            }
        }
    };

    const int n_batches = (ye - ys + 2 * R + BR - 1) / BR;
#pragma unroll
    for (int s = 0; s < S - 1; ++s) {
        if (s < n_batches) issue(ys - R + s * BR, s);
        cp_async_commit();
    }
    float win[2 * R];  // row-filtered values of this thread's column, rows yo0-R .. yo0+R-1
#pragma unroll
    for (int i = 0; i < 2 * R; ++i) win[i] = 0.f;
    const int x = x0 + tid;
    const bool col_ok = x < w;
    const bool dec_col = (dst2 != nullptr) && !(x & 1) && ((x >> 1) < w2);
    int stage = 0;
    // Software pipeline with ONE barrier per batch: iteration b runs the row pass of batch b
    // (in_s[stage] -> hb[b&1]) and the column pass of batch b-1 (hb[(b-1)&1] -> HBM); the extra
    // iteration b == n_batches drains the last column pass.
    for (int b = 0; b <= n_batches; ++b) {
        cp_async_wait<S - 2>();  // batch b has landed (this thread's part); one group may stay in flight
        __syncthreads();         // everyone's part of batch b; hb[(b-1)&1] complete; hb[b&1] and stage (b-1)%S free
        {
            int ps = stage + S - 1;
            if (ps >= S) ps -= S;
            if (b + S - 1 < n_batches) issue(ys - R + (b + S - 1) * BR, ps);
            cp_async_commit();
        }
        // ---- row pass of batch b: warp <-> batch row, 2 groups of 4 adjacent columns per lane
        if (b < n_batches) {
            const float *rowp = in_s + stage * (BR * INW) + warp * INW;
            float *outp = hb + (b & 1) * (BR * TW) + warp * TW;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const int cidx = g * 128 + 4 * lane;
                float v[4 + 2 * RP];
#pragma unroll
                for (int q = 0; q < (4 + 2 * RP) / 4; ++q) {
                    const float4 t = *reinterpret_cast<const float4 *>(rowp + cidx + 4 * q);
                    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                }
                float4 o;
                float *op = reinterpret_cast<float *>(&o);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float acc = taps.t[0] * v[j + RP];
#pragma unroll
                    for (int k = 1; k <= R; ++k) acc = fmaf(taps.t[k], v[j + RP + k] + v[j + RP - k], acc);
                    op[j] = acc;
                }
                *reinterpret_cast<float4 *>(outp + cidx) = o;
            }
        }
        // ---- column pass of batch b-1: window = win[2R] (registers) ++ nw[BR] (from hb)
        if (b >= 1) {
            const int bb = b - 1;
            const float *hp = hb + (bb & 1) * (BR * TW) + tid;
            float nw[BR];
#pragma unroll
            for (int t = 0; t < BR; ++t) nw[t] = hp[t * TW];
            const int yo0 = ys + bb * BR - 2 * R;  // output row of t = 0
            if (yo0 + BR - 1 >= ys) {
                float out[BR];
#pragma unroll
                for (int t = 0; t < BR; ++t) {
                    // value at window index i: i < 2R ? win[i] : nw[i - 2R]; centre of output t is t+R
                    auto at = [&](int i) -> float { return i < 2 * R ? win[i] : nw[i - 2 * R]; };
                    float acc = taps.t[0] * at(t + R);
#pragma unroll
                    for (int k = 1; k <= R; ++k) acc = fmaf(taps.t[k], at(t + R + k) + at(t + R - k), acc);
                    out[t] = acc;
                }
                if (yo0 >= ys && yo0 + BR <= ye) {
                    if (col_ok) {
                        float *o = dst + (size_t)yo0 * pitch + x;
#pragma unroll
                        for (int t = 0; t < BR; ++t) o[(size_t)t * pitch] = out[t];
                    }
                    if (dec_col) {
#pragma unroll
                        for (int t = 0; t < BR; ++t) {
                            const int yo = yo0 + t;
                            if (!(yo & 1) && (yo >> 1) < h2) dst2[(size_t)(yo >> 1) * pitch2 + (x >> 1)] = out[t];
                        }
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < BR; ++t) {
                        const int yo = yo0 + t;
                        if (yo >= ys && yo < ye) {
                            if (col_ok) dst[(size_t)yo * pitch + x] = out[t];
                            if (dec_col && !(yo & 1) && (yo >> 1) < h2)
                                dst2[(size_t)(yo >> 1) * pitch2 + (x >> 1)] = out[t];
                        }
                    }
                }
            }
            // slide the window down by BR rows
#pragma unroll
            for (int i = 0; i < 2 * R; ++i) win[i] = (i + BR < 2 * R) ? win[i + BR] : nw[i + BR - 2 * R];
        }
        if (++stage == S) stage = 0;
    }
    cp_async_wait<0>();
}
