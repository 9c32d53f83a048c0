// 4x4x8 SIFT descriptors, one warp per oriented keypoint (included by detect.cu, which is
// compiled with --fmad=false).
//
// Replaces /root/reference/sift_impl.py:349-358 (unpack_octave) and :361-526
// (generate_descriptors: window gather, trilinear scatter with np.add.at, threshold /
// normalise / quantise).
//
// Two phases per 32 window pixels:
//   filter  : every lane tests one pixel of the clipped (2*half_w+1)^2 window against the rotated
//             4x4 grid (|r_rot|, |c_rot| < 2.5*hist_width, float32 with a safety margin -- half
//             of the window fails, :429-430) and the survivors are compacted into a per-warp
//             queue with a ballot;
//   scatter : whenever 32 survivors are queued all lanes evaluate one each, with the reference's
//             dtypes (float64 geometry :421-426, float32 gradient / orientation :414-417,:455-456),
//             and add their 8 trilinear shares to a lane-private float32 4x4x8 histogram in
//             shared memory ([bin][lane]: conflict free, no atomics).  Only the inner 4x4 cells
//             of the reference's 6x6 tensor are ever read (:509), so shares of the border ring
//             are dropped.
// The 32 private histograms are then summed in a fixed order (deterministic), followed by the
// 0.2 clip, renormalisation and round(512 v) of :512-524 with warp shuffles.
#pragma once

constexpr int kDescWarps = 1;   // one warp per CTA: 13 CTAs (16.4 KB histogram + queue + 1 KB reserve each) fit an SM
constexpr int kDescHistFloats = 128 * 32;
constexpr int kDescU = 2;                       // surviving pixels evaluated per lane and batch
constexpr int kDescQueue = 32 * kDescU + 32 + 4;   // also the row table of the interval path (<= 96 rows + sentinel)
constexpr int kDescMaxRows = 95;   // + 5 sentinel entries
constexpr size_t kDescSmemPerWarp = kDescHistFloats * sizeof(float) + kDescQueue * sizeof(int);

__global__ void __launch_bounds__(kDescWarps * 32, 13)
describe_kernel(PyrView v, DetectParams dp, const RawKeypoint *__restrict__ raw, int n, int converted,
                uint8_t *__restrict__ desc_out, int32_t *__restrict__ work_counter)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float *hist = reinterpret_cast<float *>(dsm + (size_t)wib * kDescSmemPerWarp);
    // queue of surviving window offsets, packed (ys << 16) | (xs & 0xffff): |xs|, |ys| < 32768 because the
    // window is clipped to the image and pyramid dimensions are below 32768
    int *q = reinterpret_cast<int *>(hist + kDescHistFloats);
    auto qx_of = [](int v) { return (v << 16) >> 16; };
    auto qy_of = [](int v) { return v >> 16; };
    const int warps_total = gridDim.x * kDescWarps;
    constexpr int U = kDescU;
    const unsigned lt_mask = (1u << lane) - 1u;
    (void)warps_total;
    // keypoint windows differ 6x in size: warps take the next keypoint from a global counter
    // instead of a static stride, which removes the tail where a few warps still work
    for (;;) {
        int ki = 0;
        if (lane == 0) ki = atomicAdd(work_counter, 1);
        ki = __shfl_sync(0xffffffffu, ki, 0);
        if (ki >= n) break;
        const RawKeypoint K = raw[ki];
        // convert_keypoints_to_input_image_size (:333-343) unless already done
        const float kx = converted ? K.x : K.x * 0.5f, ky = converted ? K.y : K.y * 0.5f;
        const float ksize = converted ? K.size : K.size * 0.5f;
        const int koct = converted ? K.octave_packed : ((K.octave_packed & ~255) | ((K.octave_packed - 1) & 255));
        // unpack_octave (:349-358)
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
        int half_w = (int)rint((double)hist_width * 1.4142135623730951 * 5 * 0.5);
        const int diag = (int)sqrt((double)((long long)rows * rows + (long long)cols * cols));
        half_w = min(half_w, diag);
        const double hw = (double)hist_width;
        const float anglef = (float)angle;
        const float bins_per_deg = (float)(8 / 360.);
        const float cos_f = (float)cos_a, sin_f = (float)sin_a;
        const float lim = 2.5f * hist_width * 1.0001f + 1e-3f;  // float32 pre-filter, exact test below

#pragma unroll 8
        for (int b = 0; b < 128; ++b) hist[b * 32 + lane] = 0.f;

        // Evaluate TWO surviving pixels per lane (window offsets xs, ys) and scatter their shares.
        // Straight-line code on purpose: the two independent dependency chains (gather, sqrt,
        // atan2, exp, shared-memory read-modify-write) interleave and hide each other's latency;
        // with 16 KB of histogram per warp only 12 warps fit on an SM.
        const double inv_hw = 1.0 / hw;
        // gather4: the four neighbours of a queued pixel (every queued pixel lies inside the clipped
        // window, so the address is always valid).  Issued one batch AHEAD of scatter2: the L2
        // latency of the gather (the shared-memory histograms leave almost no L1) is covered by the
        // arithmetic of the previous batch instead of stalling the warp.
        auto gather4 = [&](const int (&xs)[U], const int (&ys)[U], float (&g)[U][4]) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float *p = img + (size_t)(pty + ys[u]) * pitch + (ptx + xs[u]);
                g[u][0] = __ldg(p + 1);
                g[u][1] = __ldg(p - 1);
                g[u][2] = __ldg(p - pitch);
                g[u][3] = __ldg(p + pitch);
            }
        };
        auto scatter2 = [&](const int (&xs)[U], const int (&ys)[U], const bool (&live)[U], const float (&g)[U][4]) {
            bool okc[U][4];
            float *cell[U][4];
            float mv[U][4], w0[U], w1[U];
            int o0[U], o1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double r_rot = xs[u] * sin_a + ys[u] * cos_a;
                const double c_rot = xs[u] * cos_a - ys[u] * sin_a;
                const double qr = r_rot * inv_hw, qc = c_rot * inv_hw;
                const double r_bin = qr + 1.5, c_bin = qc + 1.5;
                const bool in = live[u] && (r_bin > -1.0 && r_bin < 4.0 && c_bin > -1.0 && c_bin < 4.0);
                const float gx = in ? g[u][0] - g[u][1] : 0.f;
                const float gy = in ? g[u][2] - g[u][3] : 0.f;
                const float mag = sqrtf(gx * gx + gy * gy);
                const float orient = mod360f(atan2f(gy, gx) * B200_RAD2DEGF);
                const float fqr = (float)qr, fqc = (float)qc;
                const float wm = expf(-0.125f * (fqr * fqr + fqc * fqc)) * mag;
                float ob = (orient - anglef) * bins_per_deg;  // np.mod(ob, 8) in float32:
                ob = ob - 8.f * truncf(ob * 0.125f);          // exact fmod for |ob| < 16
                if (ob != 0.f) { if (ob < 0.f) ob += 8.f; } else ob = 0.f;
                const int r0 = __double2int_rd(r_bin), c0 = __double2int_rd(c_bin);
                o0[u] = ((int)floorf(ob)) & 7;
                o1[u] = (o0[u] + 1) & 7;
                const float rf = (float)(r_bin - (double)r0), cf = (float)(c_bin - (double)c0);
                const float of = ob - (float)o0[u];
                const float c1 = wm * rf, c0w = wm - c1;
                mv[u][0] = c0w * (1.f - cf);  // (r0,   c0)
                mv[u][1] = c0w * cf;          // (r0,   c0+1)
                mv[u][2] = c1 * (1.f - cf);   // (r0+1, c0)
                mv[u][3] = c1 * cf;           // (r0+1, c0+1)
                w1[u] = of;
                w0[u] = 1.f - of;
                // inner cells only: tensor index r0+dr+1 in [1,4]  <=>  r0+dr in [0,3]
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int rb = r0 + (k >> 1), cb = c0 + (k & 1);
                    okc[u][k] = in && ((unsigned)rb < 4u) && ((unsigned)cb < 4u);
                    cell[u][k] = hist + ((rb * 4 + cb) * 8) * 32 + lane;
                }
            }
            // The eight bins of one pixel are distinct: all loads before the first store.  The
            // second pixel may hit the same bins, so it is applied after the first one's stores.
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float h0[4], h1[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    h0[k] = okc[u][k] ? cell[u][k][o0[u] * 32] : 0.f;
                    h1[k] = okc[u][k] ? cell[u][k][o1[u] * 32] : 0.f;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (okc[u][k]) {
                        cell[u][k][o0[u] * 32] = h0[k] + mv[u][k] * w0[u];
                        cell[u][k][o1[u] * 32] = h1[k] + mv[u][k] * w1[u];
                    }
                }
            }
        };

        // the window clipped to the pixels that pass the first mask (:400)
        const int rlo = max(pty - half_w, 1), rhi = min(pty + half_w, rows - 2);
        const int clo = max(ptx - half_w, 1), chi = min(ptx + half_w, cols - 2);
        const int nx = chi - clo + 1, ny = rhi - rlo + 1;
        const int total = (ok && nx > 0 && ny > 0) ? nx * ny : 0;
        // ---- which window pixels pass the float32 pre-filter (the rotated 4x4 grid, :429-430)?
        // For a fixed row the filter  |x*sin + y*cos| < lim && |x*cos - y*sin| < lim  holds on an
        // INTERVAL of x (each term is monotone in x, also after float32 rounding), so the survivors
        // are enumerated row by row instead of testing every pixel and compacting with ballots: a
        // lane derives the interval of its rows analytically, widens it by two pixels and shrinks it
        // with the very predicate the per-pixel test used -- the survivor set and its row-major order
        // are exactly those of the per-pixel scan (kept below for windows taller than the row table).
        auto keep_px = [&](int xs, float fy) -> bool {
            const float fx = (float)xs;
            return (fabsf(fx * sin_f + fy * cos_f) < lim) && (fabsf(fx * cos_f - fy * sin_f) < lim);
        };
        if (total > 0 && ny <= kDescMaxRows && total < 32768) {
            const int xmin = clo - ptx, xmax = chi - ptx;
            int T = 0;  // survivors so far (warp-uniform)
            for (int r0 = 0; r0 < ny; r0 += 32) {
                const int r = r0 + lane;
                int a = 0, cnt = 0;
                if (r < ny) {
                    const float fy = (float)(rlo + r - pty);
                    float xl = (float)xmin, xh = (float)xmax;
                    bool none = false;
                    const float b1 = fy * cos_f, b2 = -(fy * sin_f);
                    if (fabsf(sin_f) > 1e-6f) {
                        const float t0 = (-lim - b1) / sin_f, t1 = (lim - b1) / sin_f;
                        xl = fmaxf(xl, fminf(t0, t1));
                        xh = fminf(xh, fmaxf(t0, t1));
                    } else if (!(fabsf(b1) < lim + 1.f)) {
                        none = true;
                    }
                    if (fabsf(cos_f) > 1e-6f) {
                        const float t0 = (-lim - b2) / cos_f, t1 = (lim - b2) / cos_f;
                        xl = fmaxf(xl, fminf(t0, t1));
                        xh = fminf(xh, fmaxf(t0, t1));
                    } else if (!(fabsf(b2) < lim + 1.f)) {
                        none = true;
                    }
                    int b = -1;
                    if (!none && xl <= xh + 4.f) {
                        a = max(xmin, (int)floorf(xl) - 2);
                        b = min(xmax, (int)ceilf(xh) + 2);
                        while (a <= b && !keep_px(a, fy)) ++a;
                        while (b >= a && !keep_px(b, fy)) --b;
                    } else {
                        a = 0;
                    }
                    cnt = b >= a ? b - a + 1 : 0;
                }
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                if (r < ny) q[r] = ((T + incl - cnt) << 16) | (a & 0xffff);
                T += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane < 5) q[ny + lane] = T << 16;  // sentinels: every item index is below them
            __syncwarp();
            // items in batches of 32*U, lane <-> item base + lane + 32u (the assignment of the queue path);
            // the gather of a batch is issued before the previous batch is evaluated
            int rw[U];
#pragma unroll
            for (int u = 0; u < U; ++u) rw[u] = 0;
            int px[U], py[U];
            float pg[U][4];
            bool plive[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { px[u] = 0; py[u] = 0; plive[u] = false; }
            bool pending = false;  // warp-uniform
            for (int base = 0; base < T; base += 32 * U) {
                int sx[U], sy[U];
                bool live[U];
                float ng[U][4];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = base + lane + 32 * u;
                    live[u] = i < T;
                    const int it = live[u] ? i : 0;   // dead lanes gather item 0 (valid address) and drop it
                    int r = live[u] ? rw[u] : 0;
                    // advance to the row that holds item `it`: starts are non-decreasing (rows without
                    // survivors have equal starts), so the number of the next four starts that are <= it
                    // is the number of rows to skip; four independent loads instead of a dependent chain
                    for (;;) {
                        const int s1 = q[r + 1] >> 16, s2 = q[r + 2] >> 16, s3 = q[r + 3] >> 16, s4 = q[r + 4] >> 16;
                        const int adv = (s1 <= it) + (s2 <= it) + (s3 <= it) + (s4 <= it);
                        r += adv;
                        if (adv < 4) break;
                    }
                    if (live[u]) rw[u] = r;
                    const int e = q[r];
                    sx[u] = ((e << 16) >> 16) + (it - (e >> 16));
                    sy[u] = rlo + r - pty;
                }
                gather4(sx, sy, ng);
                if (pending) scatter2(px, py, plive, pg);
                pending = true;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    px[u] = sx[u]; py[u] = sy[u]; plive[u] = live[u];
#pragma unroll
                    for (int k = 0; k < 4; ++k) pg[u][k] = ng[u][k];
                }
            }
            if (pending) scatter2(px, py, plive, pg);
        } else {
        int yy = lane / max(nx, 1), xx = lane - yy * max(nx, 1);
        int qn = 0;  // warp-uniform queue length
        int px[U], py[U];  // batch whose gather is in flight
        float pg[U][4];
        bool all_live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { px[u] = 0; py[u] = 0; all_live[u] = true; }
        bool pending = false;  // warp-uniform
        for (int idx0 = 0; idx0 < total; idx0 += 32) {
            const int ys = rlo + yy - pty, xs = clo + xx - ptx;
            xx += 32;
            while (xx >= nx) { xx -= nx; ++yy; }
            const float fx = (float)xs, fy = (float)ys;
            const bool keep = (idx0 + lane < total) && (fabsf(fx * sin_f + fy * cos_f) < lim) &&
                              (fabsf(fx * cos_f - fy * sin_f) < lim);
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int pos = qn + __popc(m & lt_mask);
                q[pos] = (ys << 16) | (xs & 0xffff);
            }
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32 * U) {
                int sx[U], sy[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int e = q[lane + 32 * u];
                    sx[u] = qx_of(e); sy[u] = qy_of(e);
                }
                const int te = q[lane + 32 * U];
                __syncwarp();
                qn -= 32 * U;
                if (lane < qn) q[lane] = te;
                float ng[U][4];
                gather4(sx, sy, ng);
                if (pending) scatter2(px, py, all_live, pg);
                pending = true;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    px[u] = sx[u]; py[u] = sy[u];
#pragma unroll
                    for (int k = 0; k < 4; ++k) pg[u][k] = ng[u][k];
                }
                __syncwarp();
            }
        }
        {
            // drain: gather of the last (partial) batch goes out before the pending one is evaluated
            int sx[U], sy[U];
            bool live[U];
            float ng[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                live[u] = lane + 32 * u < qn;
                // dead lanes gather queue entry 0 (a valid address when qn > 0) and drop the result
                const int e = qn > 0 ? q[live[u] ? lane + 32 * u : 0] : 0;
                sx[u] = qx_of(e); sy[u] = qy_of(e);
#pragma unroll
                for (int k = 0; k < 4; ++k) ng[u][k] = 0.f;
            }
            if (qn > 0) gather4(sx, sy, ng);
            if (pending) scatter2(px, py, all_live, pg);
            if (qn > 0) scatter2(sx, sy, live, ng);
        }
        }
        __syncwarp();

        float vq[4];
        double ss = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = lane + 32 * q;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // fixed order, 4 chains
#pragma unroll
            for (int l = 0; l < 32; l += 4) {
                s0 += hist[e * 32 + ((l + lane) & 31)];
                s1 += hist[e * 32 + ((l + 1 + lane) & 31)];
                s2 += hist[e * 32 + ((l + 2 + lane) & 31)];
                s3 += hist[e * 32 + ((l + 3 + lane) & 31)];
            }
            const float s = (s0 + s1) + (s2 + s3);
            vq[q] = s;
            ss += (double)(s * s);
        }
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
    }
}
