// Separable Gaussian blur, packed ring kernel (included by pyramid.cu).
//
// Replaces cv2.GaussianBlur(img,(0,0),sigma) on float32 (/root/reference/sift_impl.py:56,91).
//
// The scalar strip kernel (blur_strip.cuh) spends ~55 (R=5) .. ~95 (R=13) issue slots per pixel,
// 60 % of them outside the FP32 pipe, and waits on its block barrier; this kernel does the same
// 8 B/pixel job in ~18 .. ~40 slots:
//   * every FADD / FMUL / FFMA works on an aligned PAIR of adjacent columns (add/mul/fma.f32x2 ->
//     FADD2 / FMUL2 / FFMA2 on sm_100); the taps are uniform-register operands broadcast to both
//     halves (`FFMA2 R, R, UR.F32, R`), so they cost no vector registers;
//   * row pass: a lane produces 8 adjacent outputs of one row from 16 B shared-memory loads.  Even
//     taps combine aligned input pairs into aligned output pairs; odd taps are accumulated for the
//     output pairs shifted by one column (again aligned input pairs) and the two partial sums are
//     added at the end -- no register shuffling, no second copy of the row;
//   * column pass: a thread owns a column pair; the row-filtered rows live in a shared-memory RING
//     (2R + 16 rows), each ring row is read once per 8 output rows and scattered into 8
//     accumulator pairs (16 registers instead of a 2R-deep register window, no window moves);
//   * a warp fills (cp.async, 16 B) exactly the two input rows it row-filters, so the only
//     block-wide dependency is the ring: ONE barrier per batch of 8 rows, 128 threads per CTA;
//   * 16 B chunks of every shared row are XOR-swizzled (chunk ^ ((chunk >> 3) & 1)) so that lanes
//     32 B apart hit distinct bank groups.
// One CTA marches down a 256-column strip of `seg_rows` rows.  Every input float is read once from
// HBM (+ x halo and 2R/seg_rows y halo, L2 hits) and every output written once: 8 B per pixel.
//
// Arithmetic: row pass  (k0*c + sum_{k even} k[k]*(a[+k] + a[-k])) + sum_{k odd} k[k]*(a[+k] + a[-k]),
//             column pass sum_{d=-R..R} k[|d|]*a[d] accumulated top to bottom,
// float32 with fused multiply-add, BORDER_REFLECT_101.  The association differs from OpenCV's
// (k0*c + sum_k k[k]*(a[+k]+a[-k]), rows then columns) in the last bits only: max |gpu - cv2| and
// max |gpu - oracle| stay below 1e-4 on the 0..255 range (tests/test_gpu_parity.py asserts 2e-4;
// the reference's own IPP-on and IPP-off blurs differ by 7.6e-5).
// dst2 (optional) receives the [::2, ::2] decimation that seeds the next octave (sift_impl.py:95-96).
#pragma once

constexpr int kRingW = 256;      // strip width (columns)
constexpr int kRingBR = 8;       // rows per batch
constexpr int kRingThreads = 128;

template <int R>
struct RingCfg {
    static constexpr int E4 = ((R + (R & 1)) + 3) & ~3;       // x halo per side, floats (multiple of 4, >= R + (R odd))
    static constexpr int INW = kRingW + 2 * E4;                // floats per staged input row
    static constexpr int NCH = INW / 4;                        // 16 B chunks per staged input row
    static constexpr int NV = (2 * E4 + 8) / 4;                // chunks a lane loads per row in the row pass
    static constexpr int Q = (2 * R + kRingBR - 1) / kRingBR;  // ring slots reaching back from the newest one
    static constexpr int NS = Q + 2;                           // ring slots (8 rows each)
    static constexpr int S = (R >= 12) ? 2 : 3;                // input stages (cp.async ring)
    static constexpr size_t smem = (size_t)(S * kRingBR * INW + NS * kRingBR * kRingW) * sizeof(float);
};

__device__ __forceinline__ int ring_swz(int chunk) { return chunk ^ ((chunk >> 3) & 1); }

template <int R>
__global__ void __launch_bounds__(kRingThreads)
blur_ring_kernel(const float *__restrict__ src, float *__restrict__ dst, float *__restrict__ dst2, int h, int w,
                 int pitch, size_t img_stride, int h2, int w2, int pitch2, size_t img_stride2, int seg_rows,
                 const __grid_constant__ BlurTaps<R> taps)
{
    using C = RingCfg<R>;
    constexpr int TW = kRingW, BR = kRingBR, S = C::S, E4 = C::E4, INW = C::INW, NCH = C::NCH, NV = C::NV;
    constexpr int NS = C::NS, Q = C::Q;
    constexpr int SLOT = BR * TW;  // floats per ring slot
    extern __shared__ __align__(16) float smem[];
    float *in_s = smem;                  // [S][BR][INW], chunks swizzled
    float *ring = smem + S * BR * INW;   // [NS][BR][TW], chunks swizzled

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * TW;
    const int ys = blockIdx.y * seg_rows;
    const int ye = min(ys + seg_rows, h);
    src += (size_t)blockIdx.z * img_stride;
    dst += (size_t)blockIdx.z * img_stride;
    if (dst2) dst2 += (size_t)blockIdx.z * img_stride2;

    // ---- fill: warp `warp` stages rows 2*warp, 2*warp+1 of a batch; lane takes chunks lane, lane+32, lane+64.
    // A chunk is fetched when it lies inside the row's allocation [0, pitch); columns left of 0 and
    // right of w-1 are patched from their BORDER_REFLECT_101 sources after the batch has landed
    // (shared -> shared, by the warp that owns the row), so the fill itself never branches per lane
    // on the image border and never waits on a global load.
    constexpr int NQ = (NCH + 31) / 32;
    int f_gx[NQ], f_so[NQ];
    unsigned f_in = 0;
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
        const int q = lane + 32 * k;
        f_gx[k] = x0 - E4 + 4 * q;
        f_so[k] = 4 * ring_swz(q);
        if (q < NCH && f_gx[k] >= 0 && f_gx[k] + 4 <= pitch) f_in |= 1u << k;
    }
    auto issue = [&](int yb, int stage) {
        float *st = in_s + stage * (BR * INW) + (2 * warp) * INW;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int y = reflect101(yb + 2 * warp + r, h);  // warp-uniform
            const float *p = src + (size_t)y * pitch;
            float *d = st + r * INW;
#pragma unroll
            for (int k = 0; k < NQ; ++k)
                if (f_in >> k & 1) cp_async16(d + f_so[k], p + f_gx[k]);
        }
    };
    // border patch: staged index e <-> column x0 - E4 + e.  Lane i < E4 rewrites column -1-i from
    // column 1+i (left edge) and column w+i from column w-2-i (right edge).  Only columns within R
    // of a valid output column matter, and their sources are inside the staged window because
    // R <= E4 <= w/2 (the host routes narrower images to the tile kernel).
    const bool edge_l = (x0 == 0), edge_r = (x0 + TW + E4 > w);
    auto sidx = [&](int e) -> int { return 4 * ring_swz(e >> 2) + (e & 3); };
    const int pl_dst = sidx(E4 - 1 - (lane % E4)), pl_src = sidx(E4 + 1 + (lane % E4));
    const int e_r = E4 + (w - x0) + (lane % E4);            // staged index of column w + i
    const bool pr_ok = edge_r && lane < E4 && e_r < INW;
    const int pr_dst = pr_ok ? sidx(e_r) : 0, pr_src = pr_ok ? sidx(E4 + (w - x0) - 2 - (lane % E4)) : 0;
    auto patch = [&](int stage) {
        float *st = in_s + stage * (BR * INW) + (2 * warp) * INW;
        if (edge_l && lane < E4) {
            st[pl_dst] = st[pl_src];
            st[INW + pl_dst] = st[INW + pl_src];
        }
        if (pr_ok) {
            st[pr_dst] = st[pr_src];
            st[INW + pr_dst] = st[INW + pr_src];
        }
        __syncwarp();
    };

    const int n_batches = (ye - ys + 2 * R + BR - 1) / BR;
#pragma unroll
    for (int s = 0; s < S - 1; ++s) {
        if (s < n_batches) issue(ys - R + s * BR, s);
        cp_async_commit();
    }

    const int x = x0 + 2 * tid;                       // this thread's column pair in the column pass
    const bool col2 = x + 1 < w, col1 = x < w;
    const bool dec_col = (dst2 != nullptr) && ((x >> 1) < w2) && col1;  // x is even
    const int c_off = 4 * ring_swz(tid >> 1) + 2 * (tid & 1);             // float offset of the pair in a ring row
    int stage = 0;      // b % S
    int wslot = 0;      // b % NS: ring slot the row pass of batch b writes
    for (int b = 0; b <= n_batches; ++b) {
        cp_async_wait<S - 2>();  // this thread's part of batch b has landed
        __syncthreads();         // batch b visible; ring slot (b-1) complete; slot b%NS and stage (b-1)%S free
        {
            int ps = stage + S - 1;
            if (ps >= S) ps -= S;
            if (b + S - 1 < n_batches) issue(ys - R + (b + S - 1) * BR, ps);
            cp_async_commit();
        }
        // ---- row pass of batch b: warp <-> rows 2*warp, 2*warp+1; lane <-> columns 8*lane .. 8*lane+7
        if (b < n_batches) {
            if (edge_l | edge_r) patch(stage);  // CTA-uniform
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float *rowp = in_s + stage * (BR * INW) + (2 * warp + r) * INW;
                float v[4 * NV];  // v[i] = staged column 8*lane + i  (output column j sits at v[E4 + j])
#pragma unroll
                for (int m = 0; m < NV; ++m) {
                    const float4 t = *reinterpret_cast<const float4 *>(rowp + 4 * ring_swz(2 * lane + m));
                    v[4 * m] = t.x; v[4 * m + 1] = t.y; v[4 * m + 2] = t.z; v[4 * m + 3] = t.w;
                }
                auto P = [&](int i) -> float2 { return make_float2(v[i], v[i + 1]); };  // i even: aligned pair
                float2 accE[4], accT[5];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    accE[j] = __fmul2_rn(make_float2(taps.t[0], taps.t[0]), P(E4 + 2 * j));
#pragma unroll
                    for (int k = 2; k <= R; k += 2)
                        accE[j] = __ffma2_rn(make_float2(taps.t[k], taps.t[k]),
                                             __fadd2_rn(P(E4 + 2 * j + k), P(E4 + 2 * j - k)), accE[j]);
                }
#pragma unroll
                for (int j = 0; j < 5; ++j) {  // outputs (2j-1, 2j): odd taps reach aligned input pairs again
                    accT[j] = __fmul2_rn(make_float2(taps.t[1], taps.t[1]),
                                         __fadd2_rn(P(E4 + 2 * j), P(E4 + 2 * j - 2)));
#pragma unroll
                    for (int k = 3; k <= R; k += 2)
                        accT[j] = __ffma2_rn(make_float2(taps.t[k], taps.t[k]),
                                             __fadd2_rn(P(E4 + 2 * j + k - 1), P(E4 + 2 * j - k - 1)), accT[j]);
                }
                float4 o0, o1;
                o0.x = accE[0].x + accT[0].y; o0.y = accE[0].y + accT[1].x;
                o0.z = accE[1].x + accT[1].y; o0.w = accE[1].y + accT[2].x;
                o1.x = accE[2].x + accT[2].y; o1.y = accE[2].y + accT[3].x;
                o1.z = accE[3].x + accT[3].y; o1.w = accE[3].y + accT[4].x;
                float *outp = ring + wslot * SLOT + (2 * warp + r) * TW;
                *reinterpret_cast<float4 *>(outp + 4 * ring_swz(2 * lane)) = o0;
                *reinterpret_cast<float4 *>(outp + 4 * ring_swz(2 * lane + 1)) = o1;
            }
        }
        // ---- column pass of batch b-1: ring rows j = 8(b-1) - 2R + i, i = 0 .. 2R+7, scattered into
        //      the 8 outputs t = 0..7 (output row ys + 8(b-1) - 2R + t, tap |i - t - R|)
        if (b >= 1) {
            const int bb = b - 1;
            const int yo0 = ys + bb * BR - 2 * R;
            if (yo0 + BR - 1 >= ys) {
                constexpr int SH = Q * BR - 2 * R;  // first ring row of the window inside slot (bb - Q)
                // slot of window row i: (bb - Q + (SH + i) / 8) mod NS;  (bb - Q) mod NS == (wslot + 1) mod NS
                int sl = wslot + 1;
                if (sl >= NS) sl -= NS;
                const float *sp[Q + 1];
#pragma unroll
                for (int m = 0; m <= Q; ++m) {
                    sp[m] = ring + sl * SLOT + c_off;
                    if (++sl == NS) sl = 0;
                }
                float2 acc[BR];
#pragma unroll
                for (int i = 0; i < 2 * R + BR; ++i) {
                    const float2 val = *reinterpret_cast<const float2 *>(sp[(SH + i) / BR] + ((SH + i) % BR) * TW);
#pragma unroll
                    for (int t = 0; t < BR; ++t) {
                        const int d = i - t - R;  // row offset of this ring row from the centre of output t
                        if (d == -R)
                            acc[t] = __fmul2_rn(make_float2(taps.t[R], taps.t[R]), val);
                        else if (d > -R && d <= R)
                            acc[t] = __ffma2_rn(make_float2(taps.t[d < 0 ? -d : d], taps.t[d < 0 ? -d : d]), val,
                                                acc[t]);
                    }
                }
                if (yo0 >= ys && yo0 + BR <= ye) {
                    float *o = dst + (size_t)yo0 * pitch + x;
                    if (col2) {
#pragma unroll
                        for (int t = 0; t < BR; ++t) *reinterpret_cast<float2 *>(o + (size_t)t * pitch) = acc[t];
                    } else if (col1) {
#pragma unroll
                        for (int t = 0; t < BR; ++t) o[(size_t)t * pitch] = acc[t].x;
                    }
                    if (dec_col) {
#pragma unroll
                        for (int t = 0; t < BR; ++t) {
                            const int yo = yo0 + t;
                            if (!(yo & 1) && (yo >> 1) < h2) dst2[(size_t)(yo >> 1) * pitch2 + (x >> 1)] = acc[t].x;
                        }
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < BR; ++t) {
                        const int yo = yo0 + t;
                        if (yo >= ys && yo < ye) {
                            float *o = dst + (size_t)yo * pitch + x;
                            if (col2) *reinterpret_cast<float2 *>(o) = acc[t];
                            else if (col1) *o = acc[t].x;
                            if (dec_col && !(yo & 1) && (yo >> 1) < h2)
                                dst2[(size_t)(yo >> 1) * pitch2 + (x >> 1)] = acc[t].x;
                        }
                    }
                }
            }
        }
        if (++stage == S) stage = 0;
        if (++wslot == NS) wslot = 0;
    }
    cp_async_wait<0>();
}
