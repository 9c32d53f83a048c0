// 4x4x8 SIFT descriptors, one warp per oriented keypoint (included by detect.cu, which is
// compiled with --fmad=false; fused multiply-adds below are explicit).
//
// Replaces /root/reference/sift_impl.py:349-358 (unpack_octave) and :361-526
// (generate_descriptors: window gather, trilinear scatter with np.add.at, threshold /
// normalise / quantise).
//
// The round-1 kernel (tools/experiments/describe_v1.cuh) spent ~300 thread instructions per
// accumulated window pixel (ncu: 471 M warp instructions for 34.4 k keypoints): float64 grid
// geometry, libdevice atan2f / expf / sqrtf with their slow-path branches, a four-way search
// that maps a flat survivor index back to its window row, and separate multiply / add in the
// histogram update.  This version does the same job in ~half the instructions:
//   * enumeration.  For a fixed window row the pixels inside the rotated 4x4 grid (:429-430) form
//     an INTERVAL of x; a lane derives the interval of its row analytically (32 rows = one band).
//     Intervals are cut into chunks of 8 pixels and a table of 16-bit entries (first x, row, pixel
//     count of the chunk; 480 B of shared memory) is filled by the row owners; an iteration then
//     takes 4U consecutive chunks: lane = (chunk slot, pixel of the chunk), ONE 16-bit load gives it
//     its row and x -- no search, no per-pixel test of the other half of the window, and lanes of a
//     quarter warp read 8 adjacent pixels (one 32 B sector per gathered row);
//   * arithmetic in float32 throughout the per-pixel part.  The reference computes the grid
//     coordinates and the Gaussian weight in float64 (:421-449) and the orientation in float32; the
//     trilinear split is continuous in all of them, so float32 rounding (1e-7 relative) moves a
//     histogram bin by parts in 1e7 -- two orders below what flips round(512 v) -- and the parity
//     tests hold this kernel to the same bars as before (|diff| <= 1 step, >= 97 % of rows identical
//     to the oracle on the same pyramid, RMS relative L2 < 1e-3 end to end);
//   * atan2 in degrees as one branch-free minimax polynomial (8 terms, <= 7e-6 deg, below the 3e-5
//     deg float32 spacing of an angle near 360), MUFU approximations for sqrt / exp2 / reciprocal;
//   * every histogram update is one FFMA between a shared-memory load and store;
//   * per keypoint: the work queue runs two keypoints ahead (the atomic, the class-table entry and the keypoint
//     record are fetched behind the previous windows' arithmetic), the row intervals of a band come from two
//     reciprocals per keypoint and start one pixel wide of the analytic bound.
// Each lane adds its shares to a lane-private float32 4x4x8 histogram in shared memory
// ([bin][lane]: conflict free, no atomics); only the inner 4x4 cells of the reference's 6x6 tensor
// are ever read (:509), so shares of the border ring are dropped.  The 32 private histograms are
// summed in a fixed order (deterministic), followed by the 0.2 clip, renormalisation and
// round(512 v) of :512-524 with warp shuffles.
#pragma once

constexpr int kDescWarps = 1;   // one warp per CTA: 13 CTAs (16.4 KB histogram + table + 1 KB reserve each) fit an SM
constexpr int kDescHistFloats = 128 * 32;
#ifndef B200SIFT_DESC_U
#define B200SIFT_DESC_U 3
#endif
constexpr int kDescU = B200SIFT_DESC_U;   // chunk slots per lane and iteration (independent dependency chains)
constexpr int kDescTab = 480;             // chunk table of one band of rows (bytes, 16-bit entries): 13 warps per SM fit
constexpr int kDescMaxRowLen = 128;       // rows of up to 64 px: bands of 30 rows (<= 240 chunks); up to 128 px: bands of
                                          // 15 rows (<= 240 chunks); longer rows (huge keypoints of the quirk): plain path
constexpr size_t kDescSmemPerWarp = kDescHistFloats * sizeof(float) + kDescTab;

// atan2(y, x) mod 2 pi in units of ORIENTATION BINS (2 pi = 8 bins; sift_impl.py:416-417 followed by
// the bins_per_degree scale of :454-455) without branches: atan(t) on [0, 1] as t * P(t^2) (minimax,
// max error 2.1e-6 deg, 9e-6 deg evaluated in float32 -- below the 3e-5 deg float32 spacing of an
// angle near 360), then the octant folds.
__device__ __forceinline__ float atan2_bins_fast(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float t = mn * fast_rcp(fmaxf(mx, 1e-30f));   // 0 when both are 0, like atan2(0, 0)
    const float s = t * t;
    float p = -5.162433314e-03f;
    p = __fmaf_rn(p, s, 2.783677512e-02f);
    p = __fmaf_rn(p, s, -7.118977452e-02f);
    p = __fmaf_rn(p, s, 1.227682611e-01f);
    p = __fmaf_rn(p, s, -1.770901683e-01f);
    p = __fmaf_rn(p, s, 2.539675610e-01f);
    p = __fmaf_rn(p, s, -4.243689677e-01f);
    p = __fmaf_rn(p, s, 1.273238699e+00f);
    float r = p * t;
    r = ay > ax ? 2.f - r : r;
    r = x < 0.f ? 4.f - r : r;
    r = y < 0.f ? 8.f - r : r;
    return r;
}

struct DescGeom {
    float cos_q, sin_q, angle_bins;   // cos, sin of the rotation divided by hist_width; angle in bins
};

// floor(x) for |x| < 2^22 without the conversion pipe: round-to-nearest of x - 0.5 through the
// 1.5 * 2^23 trick.  At an exact integer it may return x - 1 instead of x; the callers below use
// floor and the fraction x - floor together in a trilinear split, which is the same for both
// (weight 0 on the extra cell).  `xm` = x - 0.5 is supplied by the caller (folded into an earlier add).
constexpr float kMagic = 12582912.f;   // 1.5 * 2^23
__device__ __forceinline__ int floor_magic(float xm, float &fl)
{
    const float t = xm + kMagic;
    fl = t - kMagic;
    return __float_as_int(t) - 0x4B400000;
}

// floor(x + 1.5) given x: the same with the + 1 folded into the magic constant (2^23 * 1.5 + 1 is exact)
__device__ __forceinline__ int floor_magic1(float x, float &fl)
{
    const float t = x + (kMagic + 1.f);
    fl = t - kMagic;
    return __float_as_int(t) - 0x4B400000;
}

// U window pixels of this lane (offsets fx, fy from the keypoint, the four gradient neighbours g):
// grid coordinates, weight, orientation bin, trilinear split (:421-500) and the 8 histogram updates
// of each.  Straight-line code: the U dependency chains interleave; the updates of one pixel are
// applied before those of the next, which may hit the same bins.
template <int U>
__device__ __forceinline__ void desc_eval(float *__restrict__ hl, const float (&fx)[U], const float (&fy)[U],
                                          const bool (&live)[U], const float (&g)[U][4], const DescGeom &G)
{
    constexpr float kExp = -0.125f * 1.4426950408889634f;   // weight_mul (:447) * log2(e)
    int a0[U], a1[U];
    float v[U][4], w0[U], w1[U];
    bool p[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        // (x sin + y cos) / hist_width and (x cos - y sin) / hist_width (:421-426), the division folded
        // into the coefficients
        const float qr = __fmaf_rn(fx[u], G.sin_q, fy[u] * G.cos_q);
        const float qc = __fmaf_rn(fx[u], G.cos_q, -(fy[u] * G.sin_q));
        // r_bin = qr + 1.5 (:425); -1 < r_bin < 4 (:429-430)  <=>  |qr| < 2.5
        const bool in = live[u] && fabsf(qr) < 2.5f && fabsf(qc) < 2.5f;
        float r0f, c0f, o0f;
        const int r0 = floor_magic1(qr, r0f), c0 = floor_magic1(qc, c0f);   // floor(q + 1.5)
        const float rf = (qr + 1.5f) - r0f, cf = (qc + 1.5f) - c0f;
        const float gx = g[u][0] - g[u][1];
        const float gy = g[u][2] - g[u][3];
        const float mag = fast_sqrt(__fmaf_rn(gx, gx, gy * gy));
        const float wm = fast_ex2(kExp * __fmaf_rn(qr, qr, qc * qc)) * mag;  // :448-451
        float ob = atan2_bins_fast(gy, gx) - G.angle_bins;                   // np.mod(ob, 8) in float32 (:455-456):
        ob = ob >= 8.f ? ob - 8.f : ob;                                      //   |ob| <= 8 here
        ob = ob < 0.f ? ob + 8.f : ob;                                       //   may round up to 8.0, as numpy's does
        const int o0 = floor_magic(ob - 0.5f, o0f) & 7;                      // :461-462
        const float of = ob - (float)o0;                                     // :465 (8.0 - 0 in the rounded-up case)
        const float c1 = wm * rf, c0w = wm - c1, omcf = 1.f - cf;            // :469-476
        v[u][0] = c0w * omcf;   // (r0,   c0)
        v[u][1] = c0w * cf;     // (r0,   c0+1)
        v[u][2] = c1 * omcf;    // (r0+1, c0)
        v[u][3] = c1 * cf;      // (r0+1, c0+1)
        w1[u] = of;
        w0[u] = 1.f - of;
        const int cell8 = (r0 * 4 + c0) * 8;
        a0[u] = (cell8 + o0) * 32;
        a1[u] = (cell8 + ((o0 + 1) & 7)) * 32;
        // inner cells only: tensor index r0+dr+1 in [1,4]  <=>  r0+dr in [0,3]; `in` bounds r0, c0 to [-1, 3]
        const bool pr0 = in && r0 >= 0, pr1 = in && r0 <= 2, pc0 = c0 >= 0, pc1 = c0 <= 2;
        p[u][0] = pr0 && pc0; p[u][1] = pr0 && pc1; p[u][2] = pr1 && pc0; p[u][3] = pr1 && pc1;
    }
    constexpr int koff[4] = {0, 8 * 32, 4 * 8 * 32, 5 * 8 * 32};
#pragma unroll
    for (int u = 0; u < U; ++u) {
        float h0[4], h1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // the eight bins of one pixel are distinct: all loads before the first store
            h0[k] = p[u][k] ? hl[a0[u] + koff[k]] : 0.f;
            h1[k] = p[u][k] ? hl[a1[u] + koff[k]] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (p[u][k]) {
                hl[a0[u] + koff[k]] = __fmaf_rn(v[u][k], w0[u], h0[k]);
                hl[a1[u] + koff[k]] = __fmaf_rn(v[u][k], w1[u], h1[k]);
            }
        }
    }
}

// One keypoint's window on its pyramid layer (unpack_octave :349-358 and :373-388).
struct DescWin {
    const float *img;   // layer (octave + 1, layer) of the keypoint's image; nullptr: nothing to accumulate
    int pitch, rows, cols, ptx, pty, half_w;
    float hist_width, angle;
};

__device__ __forceinline__ DescWin desc_window(const PyrView &v, const DetectParams &dp, const RawKeypoint &K,
                                               int converted)
{
    // convert_keypoints_to_input_image_size (:333-343) unless already done
    const float kx = converted ? K.x : K.x * 0.5f, ky = converted ? K.y : K.y * 0.5f;
    const float ksize = converted ? K.size : K.size * 0.5f;
    const int koct = converted ? K.octave_packed : ((K.octave_packed & ~255) | ((K.octave_packed - 1) & 255));
    int octv = koct & 255;
    const int lyr = (koct >> 8) & 255;
    if (octv >= 128) octv |= -128;
    const float scl = octv >= 0 ? 1.f / (float)(1 << octv) : (float)(1 << -octv);
    const int po = octv + 1;
    const bool ok = (po >= 0 && po < v.n_oct && lyr < v.n_layers);
    DescWin W;
    W.rows = ok ? v.h[po] : 1;
    W.cols = ok ? v.w[po] : 1;
    W.pitch = ok ? v.pitch[po] : 1;
    W.img = ok ? v.layer(po, lyr, K.img) : nullptr;
    // per-keypoint scalars keep the reference's dtypes (:374-388)
    W.ptx = (int)rint((double)scl * (double)kx);
    W.pty = (int)rint((double)scl * (double)ky);
    W.hist_width = (float)dp.scale_multiplier_half * scl * ksize;
    int half_w = (int)rint((double)W.hist_width * 1.4142135623730951 * 5 * 0.5);
    const int diag = (int)sqrt((double)((long long)W.rows * W.rows + (long long)W.cols * W.cols));
    W.half_w = min(half_w, diag);
    W.angle = K.angle;
    return W;
}

// The gathers of a keypoint's window mostly miss L2 (the three layers the keypoints live on are
// larger than L2 and were last touched several kernels ago): while one keypoint is evaluated the
// warp asks for the 128 B lines of its NEXT keypoint's window, so that those gathers find L2.
__device__ __forceinline__ void desc_prefetch(const DescWin &W, int lane)
{
    if (!W.img) return;
    const int rlo = max(W.pty - W.half_w, 1) - 1, rhi = min(W.pty + W.half_w, W.rows - 2) + 1;
    const int clo = max(W.ptx - W.half_w, 1) - 1, chi = min(W.ptx + W.half_w, W.cols - 2) + 1;
    if (rhi < rlo || chi < clo) return;
    const uintptr_t row0 = reinterpret_cast<uintptr_t>(W.img + (size_t)rlo * W.pitch);
    const uintptr_t first = (row0 + (uintptr_t)clo * 4) & ~(uintptr_t)127;
    const int lpr = (int)(((row0 + (uintptr_t)chi * 4) >> 7) - (first >> 7)) + 2;   // lines per row (rows shift by pitch)
    const int n_rows = min(rhi - rlo + 1, 128);
    for (int r = lane; r < n_rows; r += 32) {
        uintptr_t a = (first + (uintptr_t)r * W.pitch * 4) & ~(uintptr_t)127;
        for (int l = 0; l < lpr && l < 6; ++l, a += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    }
}

__global__ void __launch_bounds__(kDescWarps * 32, 13)
describe_kernel(PyrView v, DetectParams dp, const RawKeypoint *__restrict__ raw, int n, int converted,
                uint8_t *__restrict__ desc_out, int32_t *__restrict__ counters, const int32_t *__restrict__ class_idx,
                int class_stride)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float *hist = reinterpret_cast<float *>(dsm + (size_t)wib * kDescSmemPerWarp);
    unsigned short *tab = reinterpret_cast<unsigned short *>(hist + kDescHistFloats);
    float *hl = hist + lane;
    constexpr int U = kDescU;
    const int slot = lane >> 3, pix = lane & 7;
    // keypoint windows differ 6x in size: warps take the next keypoint from a global counter
    // instead of a static stride, which removes the tail where a few warps still work
    // Queue position t -> keypoint: with work classes (orient_kernel filled class_idx, largest windows
    // in class 0) the classes are walked in order, so the small keypoints come last; n is then the sum
    // of the class sizes.  Without them the raw list is taken as it is.
    int cls_end[kDescClasses];
    {
        int acc = 0;
#pragma unroll
        for (int k = 0; k < kDescClasses; ++k) {
            acc += class_idx ? counters[CNT_CLASS + k] : 0;
            cls_end[k] = acc;
        }
    }
    // The queue runs TWO keypoints ahead so that the warp never waits for the chain atomic -> class
    // table -> keypoint record before it can start on the current window: `take` only issues the
    // atomic; its answer is mapped to a keypoint (`resolve`, lane 0) after the window loop, the record
    // is loaded before the normalisation epilogue and turned into a window after it.
    auto take = [&]() -> int {
        int t = 0;
        if (lane == 0) t = atomicAdd(&counters[CNT_WORK_DESC], 1);
        return t;   // valid in lane 0
    };
    auto resolve = [&](int t) -> int {   // queue position -> keypoint index (lane 0), n = none left
        if (lane == 0 && class_idx && t < n) {
            int k = 0, start = 0;
#pragma unroll
            for (int q = 0; q < kDescClasses - 1; ++q)
                if (t >= cls_end[q]) { k = q + 1; start = cls_end[q]; }
            // (a class can only outgrow its stride when the keypoint list overflowed; the stage is redone then)
            t = (t < cls_end[kDescClasses - 1] && t - start < class_stride) ? class_idx[(size_t)k * class_stride + (t - start)] : n;
        }
        return t;
    };
    int ki = __shfl_sync(0xffffffffu, resolve(take()), 0);
    int ki_next = __shfl_sync(0xffffffffu, resolve(take()), 0);
    DescWin W, Wn;
    if (ki < n) W = desc_window(v, dp, raw[ki], converted);
    if (ki_next < n) Wn = desc_window(v, dp, raw[ki_next], converted);
    while (ki < n) {
        const int t_after = take();   // the keypoint after the next one
        if (ki_next < n) desc_prefetch(Wn, lane);
        const bool ok = W.img != nullptr;
        const int rows = W.rows, cols = W.cols, pitch = W.pitch, ptx = W.ptx, pty = W.pty, half_w = W.half_w;
        const float *img = W.img;
        const float hist_width = W.hist_width;
        const double angle = 360. - (double)W.angle;
        const double rad = angle * (3.14159265358979323846 / 180.0);
        DescGeom G;
        const float cos_f = (float)cos(rad), sin_f = (float)sin(rad);
        G.cos_q = cos_f / hist_width;
        G.sin_q = sin_f / hist_width;
        G.angle_bins = (float)angle * (float)(8 / 360.);
        const float lim = 2.5f * hist_width * 1.0001f + 1e-3f;  // float32 pre-filter of the enumeration, exact test per pixel

        {
            float4 *h4 = reinterpret_cast<float4 *>(hist);
#pragma unroll 8
            for (int b = 0; b < kDescHistFloats / 128; ++b) h4[b * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        }

        // the window clipped to the pixels that pass the first mask (:400)
        const int rlo = max(pty - half_w, 1), rhi = min(pty + half_w, rows - 2);
        const int clo = max(ptx - half_w, 1), chi = min(ptx + half_w, cols - 2);
        const int nx = chi - clo + 1, ny = rhi - rlo + 1;
        const int total_px = (ok && nx > 0 && ny > 0) ? nx * ny : 0;
        // pixel (xs, ys) of the window = center[ys * pitch + xs]; one 32-bit offset per pixel, the row
        // above / below through a 64-bit pitch held in registers
        const char *center = reinterpret_cast<const char *>(img + (ptrdiff_t)pty * pitch + ptx);
        const ptrdiff_t pitch_b = (ptrdiff_t)pitch * 4;
        auto gather = [&](int xs, int ys, float (&g)[4]) {
            const char *p = center + (ys * pitch + xs) * 4;
            g[0] = __ldg(reinterpret_cast<const float *>(p + 4));
            g[1] = __ldg(reinterpret_cast<const float *>(p - 4));
            g[2] = __ldg(reinterpret_cast<const float *>(p - pitch_b));
            g[3] = __ldg(reinterpret_cast<const float *>(p + pitch_b));
        };
        // which pixels of a row pass  |x*sin + y*cos| < lim && |x*cos - y*sin| < lim ?  Each term is monotone
        // in x (also after float32 rounding), so they form an interval: derived analytically, widened by
        // two pixels and shrunk with the predicate itself.
        auto keep_px = [&](int xs, float fy) -> bool {
            const float fx = (float)xs;
            return (fabsf(fx * sin_f + fy * cos_f) < lim) && (fabsf(fx * cos_f - fy * sin_f) < lim);
        };
        if (total_px > 0 && nx <= kDescMaxRowLen) {
            const int xmin = clo - ptx, xmax = chi - ptx;
            const bool use_sin = fabsf(sin_f) > 1e-3f, use_cos = fabsf(cos_f) > 1e-3f;
            const float inv_sin = use_sin ? 1.f / sin_f : 0.f, inv_cos = use_cos ? 1.f / cos_f : 0.f;
            __syncwarp();
            const int band_rows = nx <= 64 ? 30 : 15;   // <= 8 resp. 16 chunks per row: at most 240 chunks per band
            for (int band0 = 0; band0 < ny; band0 += band_rows) {
                // ---- this lane's row of the band: interval [a, a + cnt) of window offsets
                const int r = band0 + lane;
                int a = 0, cnt = 0;
                if (r < ny && lane < band_rows) {
                    const float fy = (float)(rlo + r - pty);
                    float xl = (float)xmin, xh = (float)xmax;
                    bool none = false;
                    const float b1 = fy * cos_f, b2 = -(fy * sin_f);
                    // a term whose x coefficient is below 1e-3 moves by < 0.13 over the row: it is left to
                    // the exact predicate (and decides "no pixel at all" when it is off by more than 1);
                    // above 1e-3 the analytic bounds are good to 0.01 px (cancellation 1e-5 / 1e-3)
                    if (use_sin) {
                        const float t0 = (-lim - b1) * inv_sin, t1 = (lim - b1) * inv_sin;
                        xl = fmaxf(xl, fminf(t0, t1));
                        xh = fminf(xh, fmaxf(t0, t1));
                    } else if (!(fabsf(b1) < lim + 1.f)) {
                        none = true;
                    }
                    if (use_cos) {
                        const float t0 = (-lim - b2) * inv_cos, t1 = (lim - b2) * inv_cos;
                        xl = fmaxf(xl, fminf(t0, t1));
                        xh = fminf(xh, fmaxf(t0, t1));
                    } else if (!(fabsf(b2) < lim + 1.f)) {
                        none = true;
                    }
                    int b = -1;
                    if (!none && xl <= xh + 2.f) {
                        a = max(xmin, (int)ceilf(xl) - 1);    // one pixel of slack on each side, then shrink
                        b = min(xmax, (int)floorf(xh) + 1);
                        while (a <= b && !keep_px(a, fy)) ++a;
                        while (b >= a && !keep_px(b, fy)) --b;
                    } else {
                        a = 0;
                    }
                    cnt = b >= a ? b - a + 1 : 0;
                }
                const int nch = (cnt + 7) >> 3;
                int cend = nch;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, cend, d);
                    if (lane >= d) cend += t;
                }
                const int cstart = cend - nch;
                const int total = __shfl_sync(0xffffffffu, cend, 31);   // <= 240 chunks
                // chunk entry: bits 0-7 first x offset + 128 (|x| <= half_w + 1 <= 65 on this path),
                // bits 8-12 row of the band, bits 13-15 pixels in the chunk - 1
                for (int k = 0; k < nch; ++k)
                    tab[cstart + k] = (unsigned short)((a + 8 * k + 128) | (lane << 8) | ((min(cnt - 8 * k, 8) - 1) << 13));
                __syncwarp();
                const int ys_band = rlo + band0 - pty;
                // chunk slot u of an iteration at chunk c0: chunk c0 + 4u + slot, pixel `pix` of it
                auto lookup = [&](int c0, int (&xs)[U], int (&ys)[U], bool (&live)[U]) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int c = c0 + 4 * u + slot;
                        const bool have = c < total;
                        const int e = tab[have ? c : 0];            // total > 0 inside the loop
                        live[u] = have && pix <= (e >> 13);
                        xs[u] = (e & 255) - 128 + (live[u] ? pix : 0);   // dead lanes gather a valid pixel and drop it
                        ys[u] = ys_band + ((e >> 8) & 31);
                    }
                };
                int nxs[U], nys[U];
                bool nlive[U];
                float ng[U][4];
                if (total > 0) {
                    lookup(0, nxs, nys, nlive);
#pragma unroll
                    for (int u = 0; u < U; ++u) gather(nxs[u], nys[u], ng[u]);
                }
                for (int c0 = 0; c0 < total; c0 += 4 * U) {
                    float fx[U], fy[U], g[U][4];
                    bool live[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        fx[u] = (float)nxs[u]; fy[u] = (float)nys[u]; live[u] = nlive[u];
#pragma unroll
                        for (int k = 0; k < 4; ++k) g[u][k] = ng[u][k];
                    }
                    // the gather of the next iteration goes out before this one is evaluated: its L2
                    // latency is covered by arithmetic instead of stalling the warp
                    if (c0 + 4 * U < total) {
                        lookup(c0 + 4 * U, nxs, nys, nlive);
#pragma unroll
                        for (int u = 0; u < U; ++u) gather(nxs[u], nys[u], ng[u]);
                    }
                    desc_eval<U>(hl, fx, fy, live, g, G);
                }
                __syncwarp();   // the table is rewritten by the next band
            }
        } else if (total_px > 0) {
            // plain path (windows wider than the chunk table allows: keypoints of the non-convergence
            // quirk with a huge size): every pixel of the clipped window, one per lane
            for (int idx0 = 0; idx0 < total_px; idx0 += 32) {
                const int idx = idx0 + lane;
                const bool have = idx < total_px;
                const int yy = have ? idx / nx : 0, xx = have ? idx - yy * nx : 0;
                const int xs1 = clo + xx - ptx, ys1 = rlo + yy - pty;
                float fx[1] = {(float)xs1}, fy[1] = {(float)ys1}, g[1][4];
                bool live[1] = {have && keep_px(xs1, (float)ys1)};
                gather(xs1, ys1, g[0]);
                desc_eval<1>(hl, fx, fy, live, g, G);
            }
        }
        __syncwarp();
        const int r_after = resolve(t_after);   // lane 0: class-table load goes out, awaited after the sum below

        // fixed-order sum of the 32 private histograms: lane <-> elements lane + 32 q, 16 B loads in a
        // rotated order (the eight lanes of a quarter warp read eight different bank groups)
        float vq[4];
        double ss = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 *row4 = reinterpret_cast<const float4 *>(hist + (lane + 32 * q) * 32);
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // 4 chains
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 t = row4[(j + lane) & 7];
                s0 += t.x; s1 += t.y; s2 += t.z; s3 += t.w;
            }
            const float s = (s0 + s1) + (s2 + s3);
            vq[q] = s;
            ss += (double)(s * s);
        }
        const int ki_after = __shfl_sync(0xffffffffu, r_after, 0);
        RawKeypoint Ka;
        if (ki_after < n) Ka = raw[ki_after];   // first used after the stores of this descriptor
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, sft);
        const float thr = sqrtf((float)ss) * dp.descriptor_max_value_f;
        double ss2 = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (vq[q] > thr) vq[q] = thr;
            ss2 += (double)(vq[q] * vq[q]);
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) ss2 += __shfl_xor_sync(0xffffffffu, ss2, sft);
        float norm_v = sqrtf((float)ss2);
        if (norm_v < 1e-7f) norm_v = 1e-7f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float t = rintf(512.f * (vq[q] / norm_v));
            t = fminf(fmaxf(t, 0.f), 255.f);
            desc_out[(size_t)ki * 128 + lane + 32 * q] = (uint8_t)t;
        }
        __syncwarp();
        ki = ki_next;
        W = Wn;
        ki_next = ki_after;
        if (ki_after < n) Wn = desc_window(v, dp, Ka, converted);
    }
}

// ---------------------------------------------------------------------------
// generate_descriptors for any (window_width, num_bins) (sift_impl.py:361-362 keyword arguments;
// the kernel above is the 4 x 4 x 8 default).  One warp per keypoint, every lane walks its share of
// the clipped window and adds its trilinear shares to a lane-private histogram of the INNER
// window_width^2 x num_bins cells ([bin][lane] in dynamic shared memory, <= 1024 bins); the 32
// histograms are summed in a fixed order.  Same dtypes as the kernel above.
// ---------------------------------------------------------------------------
constexpr int kDescGenericMaxBins = 1024;

__global__ void __launch_bounds__(32)
describe_generic_kernel(PyrView v, DetectParams dp, int d, int nb, const RawKeypoint *__restrict__ raw, int n,
                        int converted, uint8_t *__restrict__ desc_out)
{
    extern __shared__ __align__(16) float ghist[];   // [d*d*nb][32]
    const int lane = threadIdx.x;
    const int dlen = d * d * nb;
    for (int ki = blockIdx.x; ki < n; ki += gridDim.x) {
        const RawKeypoint K = raw[ki];
        const float kx = converted ? K.x : K.x * 0.5f, ky = converted ? K.y : K.y * 0.5f;
        const float ksize = converted ? K.size : K.size * 0.5f;
        const int koct = converted ? K.octave_packed : ((K.octave_packed & ~255) | ((K.octave_packed - 1) & 255));
        int octv = koct & 255;
        const int lyr = (koct >> 8) & 255;
        if (octv >= 128) octv |= -128;
        const float scl = octv >= 0 ? 1.f / (float)(1 << octv) : (float)(1 << -octv);
        const int po = octv + 1;
        const bool ok = (po >= 0 && po < v.n_oct && lyr < v.n_layers);
        const int rows = ok ? v.h[po] : 1, cols = ok ? v.w[po] : 1, pitch = ok ? v.pitch[po] : 1;
        const float *img = ok ? v.layer(po, lyr, K.img) : nullptr;
        const int ptx = (int)rint((double)scl * (double)kx);
        const int pty = (int)rint((double)scl * (double)ky);
        const double angle = 360. - (double)K.angle;
        const double rad = angle * (3.14159265358979323846 / 180.0);
        const double cos_a = cos(rad), sin_a = sin(rad);
        const float hist_width = (float)dp.scale_multiplier_half * scl * ksize;
        int half_w = (int)rint((double)hist_width * 1.4142135623730951 * (d + 1) * 0.5);
        const int diag = (int)sqrt((double)((long long)rows * rows + (long long)cols * cols));
        half_w = min(half_w, diag);
        const double inv_hw = 1.0 / (double)hist_width;
        const float anglef = (float)angle;
        const float bins_per_deg = (float)(nb / 360.);
        const float wmul = (float)(-0.5 / ((0.5 * d) * (0.5 * d)));
        const double shift = 0.5 * d - 0.5;
        for (int b = 0; b < dlen; ++b) ghist[b * 32 + lane] = 0.f;
        const int rlo = max(pty - half_w, 1), rhi = min(pty + half_w, rows - 2);
        const int clo = max(ptx - half_w, 1), chi = min(ptx + half_w, cols - 2);
        const int nx = chi - clo + 1, ny = rhi - rlo + 1;
        const int total = (ok && nx > 0 && ny > 0) ? nx * ny : 0;
        for (int idx = lane; idx < total; idx += 32) {
            const int yy = idx / nx, xx = idx - yy * nx;
            const int ys = rlo + yy - pty, xs = clo + xx - ptx;
            const double r_rot = xs * sin_a + ys * cos_a;
            const double c_rot = xs * cos_a - ys * sin_a;
            const double qr = r_rot * inv_hw, qc = c_rot * inv_hw;
            const double r_bin = qr + shift, c_bin = qc + shift;
            if (!(r_bin > -1.0 && r_bin < (double)d && c_bin > -1.0 && c_bin < (double)d)) continue;
            const float *p = img + (size_t)(pty + ys) * pitch + (ptx + xs);
            const float gx = __ldg(p + 1) - __ldg(p - 1);
            const float gy = __ldg(p - pitch) - __ldg(p + pitch);
            const float mag = sqrtf(gx * gx + gy * gy);
            const float orient = mod360f(atan2f(gy, gx) * B200_RAD2DEGF);
            const float fqr = (float)qr, fqc = (float)qc;
            const float wm = expf(wmul * (fqr * fqr + fqc * fqc)) * mag;
            float ob = (orient - anglef) * bins_per_deg;          // np.mod(ob, nb) in float32
            ob = fmodf(ob, (float)nb);
            if (ob != 0.f) { if (ob < 0.f) ob += (float)nb; } else ob = 0.f;
            const int r0 = __double2int_rd(r_bin), c0 = __double2int_rd(c_bin);
            int o0 = (int)floorf(ob);                             // in [0, nb]: ob + nb can round up to nb
            if (o0 >= nb) o0 -= nb;                               // o0 % num_bins (:461); `of` is then nb (:465)
            const int o1 = o0 + 1 < nb ? o0 + 1 : 0;
            const float rf = (float)(r_bin - (double)r0), cf = (float)(c_bin - (double)c0);
            const float of = ob - (float)o0;
            const float c1 = wm * rf, c0w = wm - c1;
            const float mv[4] = {c0w * (1.f - cf), c0w * cf, c1 * (1.f - cf), c1 * cf};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int rb = r0 + (k >> 1), cb = c0 + (k & 1);
                if ((unsigned)rb < (unsigned)d && (unsigned)cb < (unsigned)d) {   // inner cells only (:509)
                    float *cell = ghist + ((rb * d + cb) * nb) * 32 + lane;
                    cell[o0 * 32] += mv[k] * (1.f - of);
                    cell[o1 * 32] += mv[k] * of;
                }
            }
        }
        __syncwarp();
        // fixed-order sum of the 32 private histograms -> element e in ghist[e*32] (lane e % 32 owns it)
        double ss = 0.0;
        for (int e = lane; e < dlen; e += 32) {
            float s = 0.f;
            for (int l = 0; l < 32; ++l) s += ghist[e * 32 + ((l + lane) & 31)];
            ghist[e * 32] = s;   // row e is read and rewritten by this lane only (no warp-level sync here:
                                 // dlen need not be a multiple of 32, so the trip count differs per lane)
            ss += (double)(s * s);
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, sft);
        const float thr = sqrtf((float)ss) * dp.descriptor_max_value_f;
        double ss2 = 0.0;
        for (int e = lane; e < dlen; e += 32) {
            float s = ghist[e * 32];
            if (s > thr) s = thr;
            ghist[e * 32] = s;
            ss2 += (double)(s * s);
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) ss2 += __shfl_xor_sync(0xffffffffu, ss2, sft);
        float norm_v = sqrtf((float)ss2);
        if (norm_v < 1e-7f) norm_v = 1e-7f;
        for (int e = lane; e < dlen; e += 32) {
            float t = rintf(512.f * (ghist[e * 32] / norm_v));
            t = fminf(fmaxf(t, 0.f), 255.f);
            desc_out[(size_t)ki * dlen + e] = (uint8_t)t;
        }
        __syncwarp();
    }
}
