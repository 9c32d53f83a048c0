// Separable Gaussian blur, packed-pair strip kernel (included by pyramid.cu).
//
// Replaces cv2.GaussianBlur(img,(0,0),sigma) on float32 (/root/reference/sift_impl.py:56,91).
//
// Same decomposition as blur_strip.cuh (256-column strip, 8-row batches, row pass -> hb -> column
// pass with a register window, one barrier per batch) but every FADD / FFMA works on TWO pixels
// (add.f32x2 / fma.rn.f32x2, FADD2 / FFMA2 on sm_100): the strip kernel spends 54 of its ~110
// issue slots per pixel on scalar FADD / FFMA and is issue bound; packed, the same arithmetic
// takes 27 slots.  Each component of a packed operation is the IEEE round-to-nearest scalar
// operation, so the results are bit-identical to the scalar kernel and to the oracle.
//
// Packed operands must sit in aligned register pairs for EVERY tap offset, which fixes the layout:
//   row pass    pairs are (row 2p, row 2p+1) of the SAME column.  The input tile is stored
//               row-pair interleaved in shared memory (float2 per column), so a tap offset moves
//               by whole float2 elements.  A 4-column block takes 48 B (32 B data + 16 B pad):
//               with lanes 48 B apart every LDS.128 of a quarter warp hits 8 distinct 16 B bank
//               groups.  Global rows are read with 16 B loads into registers and interleaved on
//               the way into shared memory (cp.async cannot interleave).
//   column pass pairs are (column 2t, column 2t+1) of the SAME row, read as float2 from the
//               row-filtered buffer hb; thread t keeps the last 2R row-filtered pairs of its two
//               columns in registers.
// 128 threads per CTA: warp w row-filters row pair w of the batch, thread t column-filters
// column pair t.  Software pipeline per iteration b (one barrier):
//   row pass of batch b (stage b%3 -> hb[b&1]), global loads of batch b+2 into registers,
//   column pass of batch b-1 (hb[(b-1)&1] -> HBM), registers -> stage (b+2)%3.
// Arithmetic per pixel: k0*c + sum_k k[k]*(a[+k] + a[-k]) in float32, rows then columns,
// BORDER_REFLECT_101 -- the order of the oracle.
#pragma once

constexpr int kPairW = 256;
constexpr int kPairBR = 8;
constexpr int kPairNP = kPairBR / 2;  // row pairs per batch
constexpr int kPairStages = 3;
constexpr int kPairThreads = 128;
constexpr int kPairBlk = 6;           // float2 per 4-column block in shared memory (4 data + 2 pad)

template <int R>
struct BlurTaps2 {
    float2 t[R + 1];  // (k, k), centre .. R
};

template <int R>
constexpr size_t pair_smem_bytes()
{
    constexpr int RP = (R + 3) & ~3;
    constexpr int NBLK = (kPairW + 2 * RP) / 4;
    return (size_t)kPairStages * kPairNP * NBLK * kPairBlk * sizeof(float2) +
           (size_t)2 * kPairBR * kPairW * sizeof(float);
}

template <int R>
__global__ void __launch_bounds__(kPairThreads, 3)
blur_pair_kernel(const float *__restrict__ src, float *__restrict__ dst, float *__restrict__ dst2, int h, int w,
                 int pitch, size_t img_stride, int h2, int w2, int pitch2, size_t img_stride2, int seg_rows,
                 const __grid_constant__ BlurTaps2<R> taps)
{
    constexpr int TW = kPairW, BR = kPairBR, NP = kPairNP, S = kPairStages, NT = kPairThreads;
    constexpr int RP = (R + 3) & ~3;
    constexpr int NBLK = (TW + 2 * RP) / 4;   // 4-column blocks per tile row
    constexpr int PW = NBLK * kPairBlk;       // float2 per row pair in shared memory
    constexpr int NIT = NP * NBLK;            // (row pair, block) fill items per batch
    constexpr int ROUNDS = (NIT + NT - 1) / NT;
    static_assert(NP == NT / 32, "one warp per row pair");
    static_assert(TW == 2 * NT, "one thread per column pair");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *in_s = reinterpret_cast<float2 *>(smem_raw);          // [S][NP][PW]
    float *hb = reinterpret_cast<float *>(in_s + S * NP * PW);    // [2][BR][TW]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * TW;
    const int ys = blockIdx.y * seg_rows;
    const int ye = min(ys + seg_rows, h);
    src += (size_t)blockIdx.z * img_stride;
    dst += (size_t)blockIdx.z * img_stride;
    if (dst2) dst2 += (size_t)blockIdx.z * img_stride2;

    // ---- fill: item i = (row pair, 4-column block); rows 2p / 2p+1 as two float4, interleaved on store
    int it_p[ROUNDS], it_gx[ROUNDS], it_so[ROUNDS];
    bool it_ok[ROUNDS], it_in[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const int i = tid + NT * r;
        it_ok[r] = i < NIT;
        const int p = it_ok[r] ? i / NBLK : 0, k = it_ok[r] ? i - p * NBLK : 0;
        it_p[r] = p;
        it_gx[r] = x0 - RP + 4 * k;
        it_in[r] = it_gx[r] >= 0 && it_gx[r] + 3 < w;
        it_so[r] = p * PW + k * kPairBlk;
    }
    float4 sa[ROUNDS], sb[ROUNDS];
    auto load = [&](int yb) {
        const bool rows_in = (yb >= 0) && (yb + BR <= h);  // CTA-uniform
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            if (!it_ok[r]) continue;
            const int y0 = yb + 2 * it_p[r];
            const int ya = rows_in ? y0 : reflect101(y0, h), yc = rows_in ? y0 + 1 : reflect101(y0 + 1, h);
            const float *pa = src + (size_t)ya * pitch, *pb = src + (size_t)yc * pitch;
            const int gx = it_gx[r];
            if (it_in[r]) {
                sa[r] = __ldg(reinterpret_cast<const float4 *>(pa + gx));
                sb[r] = __ldg(reinterpret_cast<const float4 *>(pb + gx));
            } else {
                const int g0 = reflect101(gx, w), g1 = reflect101(gx + 1, w), g2 = reflect101(gx + 2, w),
                          g3 = reflect101(gx + 3, w);
                sa[r] = make_float4(pa[g0], pa[g1], pa[g2], pa[g3]);
                sb[r] = make_float4(pb[g0], pb[g1], pb[g2], pb[g3]);
            }
        }
    };
    auto store = [&](int stage) {
        float2 *st = in_s + stage * (NP * PW);
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            if (!it_ok[r]) continue;
            float4 *d = reinterpret_cast<float4 *>(st + it_so[r]);
            d[0] = make_float4(sa[r].x, sb[r].x, sa[r].y, sb[r].y);
            d[1] = make_float4(sa[r].z, sb[r].z, sa[r].w, sb[r].w);
        }
    };

    const int n_batches = (ye - ys + 2 * R + BR - 1) / BR;
    load(ys - R);
    store(0);
    if (n_batches > 1) {
        load(ys - R + BR);
        store(1);
    }

    float2 win[2 * R];  // row-filtered pairs of this thread's two columns, rows yo0-R .. yo0+R-1
#pragma unroll
    for (int i = 0; i < 2 * R; ++i) win[i] = make_float2(0.f, 0.f);
    const int x = x0 + 2 * tid;
    const bool col2 = x + 1 < w, col1 = x < w;
    const bool dec_col = (dst2 != nullptr) && ((x >> 1) < w2);  // x is even
    int stage = 0;
    for (int b = 0; b <= n_batches; ++b) {
        __syncthreads();  // stage b%S holds batch b; hb[(b-1)&1] complete; hb[b&1] and stage (b+2)%S free
        // ---- row pass of batch b: warp <-> row pair, 2 groups of 4 adjacent columns per lane
        if (b < n_batches) {
            const float2 *rowp = in_s + stage * (NP * PW) + warp * PW;
            float *outa = hb + (b & 1) * (BR * TW) + (2 * warp) * TW, *outb = outa + TW;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const int cb = 32 * g + lane;  // block of the 4 output columns; tile column 4*cb is input c-RP
                const float2 *p = rowp + cb * kPairBlk;
                float2 v[4 + 2 * RP];
#pragma unroll
                for (int m = 0; m <= RP / 2; ++m) {
                    const float4 q0 = *reinterpret_cast<const float4 *>(p + m * kPairBlk);
                    const float4 q1 = *reinterpret_cast<const float4 *>(p + m * kPairBlk + 2);
                    v[4 * m] = make_float2(q0.x, q0.y);
                    v[4 * m + 1] = make_float2(q0.z, q0.w);
                    v[4 * m + 2] = make_float2(q1.x, q1.y);
                    v[4 * m + 3] = make_float2(q1.z, q1.w);
                }
                float2 acc[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[j] = __fmul2_rn(taps.t[0], v[j + RP]);
#pragma unroll
                    for (int k = 1; k <= R; ++k)
                        acc[j] = __ffma2_rn(taps.t[k], __fadd2_rn(v[j + RP + k], v[j + RP - k]), acc[j]);
                }
                *reinterpret_cast<float4 *>(outa + 4 * cb) = make_float4(acc[0].x, acc[1].x, acc[2].x, acc[3].x);
                *reinterpret_cast<float4 *>(outb + 4 * cb) = make_float4(acc[0].y, acc[1].y, acc[2].y, acc[3].y);
            }
        }
        // ---- global loads of batch b+2 (in flight during the column pass)
        const bool more = b + 2 < n_batches;
        if (more) load(ys - R + (b + 2) * BR);
        // ---- column pass of batch b-1: window = win[2R] (registers) ++ nw[BR] (from hb)
        if (b >= 1) {
            const int bb = b - 1;
            const float *hp = hb + (bb & 1) * (BR * TW) + 2 * tid;
            float2 nw[BR];
#pragma unroll
            for (int t = 0; t < BR; ++t) nw[t] = *reinterpret_cast<const float2 *>(hp + t * TW);
            const int yo0 = ys + bb * BR - 2 * R;  // output row of t = 0
            if (yo0 + BR - 1 >= ys) {
                float2 out[BR];
#pragma unroll
                for (int t = 0; t < BR; ++t) {
                    auto at = [&](int i) -> float2 { return i < 2 * R ? win[i] : nw[i - 2 * R]; };
                    float2 acc = __fmul2_rn(taps.t[0], at(t + R));
#pragma unroll
                    for (int k = 1; k <= R; ++k)
                        acc = __ffma2_rn(taps.t[k], __fadd2_rn(at(t + R + k), at(t + R - k)), acc);
                    out[t] = acc;
                }
                if (yo0 >= ys && yo0 + BR <= ye) {
                    // whole batch inside the segment (yo0 is even: ys and BR are multiples of 8, 2R is even)
                    float *o = dst + (size_t)yo0 * pitch + x;
                    if (col2) {
#pragma unroll
                        for (int t = 0; t < BR; ++t) *reinterpret_cast<float2 *>(o + (size_t)t * pitch) = out[t];
                    } else if (col1) {
#pragma unroll
                        for (int t = 0; t < BR; ++t) o[(size_t)t * pitch] = out[t].x;
                    }
                    if (dec_col) {
                        float *o2 = dst2 + (size_t)(yo0 >> 1) * pitch2 + (x >> 1);
#pragma unroll
                        for (int t = 0; t < BR; t += 2)
                            if ((yo0 >> 1) + (t >> 1) < h2) o2[(size_t)(t >> 1) * pitch2] = out[t].x;
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < BR; ++t) {
                        const int yo = yo0 + t;
                        if (yo >= ys && yo < ye) {
                            float *o = dst + (size_t)yo * pitch + x;
                            if (col2) *reinterpret_cast<float2 *>(o) = out[t];
                            else if (col1) *o = out[t].x;
                            if (dec_col && !(yo & 1) && (yo >> 1) < h2)
                                dst2[(size_t)(yo >> 1) * pitch2 + (x >> 1)] = out[t].x;
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 2 * R; ++i) win[i] = (i + BR < 2 * R) ? win[i + BR] : nw[i + BR - 2 * R];
        }
        // ---- registers -> stage (b+2)%S (last read by the row pass of batch b-1)
        if (more) {
            int ps = stage + 2;
            if (ps >= S) ps -= S;
            store(ps);
        }
        if (++stage == S) stage = 0;
    }
}
