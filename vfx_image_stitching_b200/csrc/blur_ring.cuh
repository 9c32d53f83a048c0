// Separable Gaussian blur, packed ring kernel (included by pyramid.cu).
//
// Replaces cv2.GaussianBlur(img,(0,0),sigma) on float32 (/root/reference/sift_impl.py:56,91).
//
// A scalar strip kernel (round 1, now tools/experiments/blur_strip.cuh) spent 54 issue slots per
// pixel at R = 5 (more for the wider kernels), 60 % of them outside the FP32 pipe, and waited on its
// block barrier; this kernel does the same 8 B/pixel job in 32 (R = 5) .. 55 (R = 13) slots (ncu,
// warp instructions per 18 x 1024 x 768 layer at R = 5: 14.0 M against 24.1 M; R = 13: 24.5 M),
// without a block barrier in its loop:
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
//   * 256 threads in two roles.  Warps 0-3 (producers) fill -- 16 B cp.async, each warp exactly the
//     two input rows it row-filters, so the input stages need only __syncwarp -- and run the row
//     pass; warps 4-7 (consumers) run the column pass and the stores.  The ring is handed over
//     slot by slot with one mbarrier pair (full / empty) per slot; there is no __syncthreads in the
//     loop, and a producer may run one batch ahead of the consumers;
//   * image borders: chunks are fetched wherever they lie inside the row allocation, and the columns
//     left of 0 / right of w-1 are patched shared -> shared from their REFLECT_101 sources once the
//     batch has landed, so no lane ever waits on a scalar global load;
//   * shared rows are linear with an ODD number of 16 B chunks per row; even lanes work on the
//     warp's first row and odd lanes on its second, so the eight lanes of a quarter warp (32 B apart
//     within a row) hit eight distinct bank groups without any address swizzle;
//   * output rows of a batch are 8-aligned (the column pass lags the row pass by ceil(2R/8) batches),
//     so every batch is stored by the same straight-line code.
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

template <int R>
struct BlurTaps {
    float t[R + 1];  // centre .. R
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

constexpr int kRingW = 256;      // strip width (columns)
constexpr int kRingBR = 8;       // rows per batch
constexpr int kRingThreads = 256;   // warps 0-3: fill + row pass, warps 4-7: column pass

template <int R, int STAGES>
struct RingCfg {
    static constexpr int E4 = ((R + (R & 1)) + 3) & ~3;       // x halo per side, floats (multiple of 4, >= R + (R odd))
    static constexpr int INW = kRingW + 2 * E4;                // staged floats per input row
    static constexpr int INP = ((INW / 4) | 1) * 4;            // row stride: odd number of 16 B chunks
    static constexpr int TWP = kRingW + 4;                     // ring row stride: 65 chunks
    static constexpr int NCH = INW / 4;                        // 16 B chunks per staged input row
    static constexpr int NV = (2 * E4 + 8) / 4;                // chunks a lane loads per tile in the row pass
    static constexpr int Q = (2 * R + kRingBR - 1) / kRingBR;  // batches the column pass lags behind the row pass
    static constexpr int NS = Q + 2;                           // ring slots (8 rows each)
    static constexpr int S = STAGES;                           // input stages (cp.async ring): 2 or 3
    static constexpr size_t smem = (size_t)(S * kRingBR * INP + NS * kRingBR * TWP) * sizeof(float);
};

__device__ __forceinline__ void cp_async16s(unsigned smem_dst, const float *gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ void ring_mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void ring_mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void ring_mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "RING_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra RING_DONE_%=;\n\t"
        "bra RING_WAIT_%=;\n\t"
        "RING_DONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}

template <int R, int STAGES>
__global__ void __launch_bounds__(kRingThreads)
blur_ring_kernel(const float *__restrict__ src, float *__restrict__ dst, float *__restrict__ dst2, int h, int w,
                 int pitch, size_t img_stride, int h2, int w2, int pitch2, size_t img_stride2, int seg_rows,
                 const __grid_constant__ BlurTaps<R> taps)
{
    using C = RingCfg<R, STAGES>;
    constexpr int TW = kRingW, BR = kRingBR, S = C::S, E4 = C::E4, INW = C::INW, INP = C::INP, TWP = C::TWP;
    constexpr int NCH = C::NCH, NV = C::NV, NS = C::NS, Q = C::Q;
    constexpr int STG = BR * INP;   // floats per input stage
    constexpr int SLOT = BR * TWP;  // floats per ring slot
    extern __shared__ __align__(16) float smem[];
    float *in_s = smem;             // [S][BR][INP]
    float *ring = smem + S * STG;   // [NS][BR][TWP]

    const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 3;
    const bool row_role = threadIdx.x < 128;  // warp-uniform
    const int tid = threadIdx.x & 127;        // index inside the role
    const int x0 = blockIdx.x * TW;
    const int ys = blockIdx.y * seg_rows;   // multiple of 8
    const int ye = min(ys + seg_rows, h);
    src += (size_t)blockIdx.z * img_stride;
    dst += (size_t)blockIdx.z * img_stride;
    if (dst2) dst2 += (size_t)blockIdx.z * img_stride2;

    // ---- fill: warp `warp` stages rows 2*warp, 2*warp+1 of a batch; lane takes chunks lane, lane+32, lane+64.
    // A chunk is fetched when it lies inside the row's allocation [0, pitch); columns left of 0 and
    // right of w-1 are patched from their BORDER_REFLECT_101 sources after the batch has landed
    // (shared -> shared, by the warp that owns the row), so the fill never branches per lane on the
    // image border and never waits on a global load.
    constexpr int NQ = (NCH + 31) / 32;
    const int f_gx = x0 - E4 + 4 * lane;
    bool f_in[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k)
        f_in[k] = (lane + 32 * k < NCH) && (f_gx + 128 * k >= 0) && (f_gx + 128 * k + 4 <= pitch);
    const float *f_src = src + f_gx;  // only dereferenced at in-range chunks
    const unsigned f_dst = (unsigned)__cvta_generic_to_shared(in_s + (2 * warp) * INP + 4 * lane);
    auto issue = [&](int yb, int stage) {
        const bool interior = (yb >= 0) && (yb + BR <= h);  // CTA-uniform
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            int y = yb + 2 * warp + r;
            if (!interior) y = reflect101(y, h);
            const float *p = f_src + (size_t)y * pitch;
            const unsigned d = f_dst + (unsigned)((stage * STG + r * INP) * sizeof(float));
#pragma unroll
            for (int k = 0; k < NQ; ++k)
                if (f_in[k]) cp_async16s(d + 512 * k, p + 128 * k);
        }
    };
    // border patch: staged index e <-> column x0 - E4 + e.  Lane i < E4 rewrites column -1-i from
    // column 1+i (left edge) and column w+i from column w-2-i (right edge).  Only columns within R
    // of a valid output column matter, and their sources are inside the staged window because
    // R <= E4 <= w/2 (the host routes narrower images to the tile kernel).
    const bool edge_l = (x0 == 0), edge_r = (x0 + TW + E4 > w);
    const int e_r = E4 + (w - x0) + lane;            // staged index of column w + lane
    const bool pl_ok = edge_l && lane < E4;
    const bool pr_ok = edge_r && lane < E4 && e_r < INW;
    auto patch = [&](int stage) {
        float *st = in_s + stage * STG + (2 * warp) * INP;
        if (pl_ok) {
            st[E4 - 1 - lane] = st[E4 + 1 + lane];
            st[INP + E4 - 1 - lane] = st[INP + E4 + 1 + lane];
        }
        if (pr_ok) {
            st[e_r] = st[e_r - 2 - 2 * lane];
            st[INP + e_r] = st[INP + e_r - 2 - 2 * lane];
        }
        __syncwarp();
    };

    // Row batch b holds rows ys - 8Q + 8b .. +7 (row-filtered into ring slot b % NS); column batch
    // bb = b - 1 >= Q writes output rows ys + 8(bb - Q) .. +7 from ring rows 8(bb - Q) .. 8bb + 7 of
    // which it needs the first 2R + 8.
    const int n_out = (ye - ys + BR - 1) / BR;  // output batches of this segment
    const int n_batches = n_out + Q;
    const int y_first = ys - R;  // input row behind ring row 0 of row batch 0
    // ring hand-off: full[slot] <- 128 row-role arrivals, empty[slot] <- 128 column-role arrivals
    __shared__ __align__(8) unsigned long long bar_full[NS], bar_empty[NS];
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            ring_mbar_init(&bar_full[i], 128);
            ring_mbar_init(&bar_empty[i], 128);
        }
    }
    __syncthreads();

    if (row_role) {
        // ================= producer: fill + row pass, one warp per row pair, no block barrier =================
#pragma unroll
        for (int s = 0; s < S - 1; ++s) {
            if (s < n_batches) issue(y_first + s * BR, s);
            cp_async_commit();
        }
        const int rsel = lane & 1, tl = lane >> 1;
        int stage = 0;            // b % S
        int wslot = 0, wpar = 1;  // ring slot of batch b and the parity its "empty" barrier is waited with
        for (int b = 0; b < n_batches; ++b) {
            cp_async_wait<S - 2>();  // this lane's part of batch b has landed
            __syncwarp();            // ... and every other lane's; all lanes are done with stage (b-1)%S
            {
                int ps = stage + S - 1;
                if (ps >= S) ps -= S;
                if (b + S - 1 < n_batches) issue(y_first + (b + S - 1) * BR, ps);
                cp_async_commit();
            }
            if (edge_l | edge_r) patch(stage);  // CTA-uniform
            ring_mbar_wait(&bar_empty[wslot], wpar);  // column pass has released the slot (free on its first use)
            // even lanes row 2*warp, odd lanes row 2*warp+1; two passes of 16 tiles of 8 columns
            const float *rowp = in_s + stage * STG + (2 * warp + rsel) * INP + 8 * tl;
            float *outp = ring + wslot * SLOT + (2 * warp + rsel) * TWP + 8 * tl;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                float v[4 * NV];  // v[i] = staged column 8*tile + i  (output column j of the tile sits at v[E4 + j])
#pragma unroll
                for (int m = 0; m < NV; ++m) {
                    const float4 t = *reinterpret_cast<const float4 *>(rowp + 128 * pass + 4 * m);
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
                *reinterpret_cast<float4 *>(outp + 128 * pass) = o0;
                *reinterpret_cast<float4 *>(outp + 128 * pass + 4) = o1;
            }
            ring_mbar_arrive(&bar_full[wslot]);
            if (++stage == S) stage = 0;
            if (++wslot == NS) { wslot = 0; wpar ^= 1; }
        }
        cp_async_wait<0>();
    } else {
        // ================= consumer: column pass, one thread per column pair =================
        const int x = x0 + 2 * tid;
        const bool col2 = x + 1 < w, col1 = x < w;
        const bool dec_col = (dst2 != nullptr) && ((x >> 1) < w2) && col1;  // x is even
        int rslot = Q % NS, rpar = 0;   // newest slot of column batch bb (= bb % NS) and its "full" parity
        int oslot = 0;                  // oldest slot of the window, (bb - Q) % NS
        for (int bb = Q; bb < n_batches; ++bb) {
            ring_mbar_wait(&bar_full[rslot], rpar);  // rows up to batch bb are in the ring
            const int yo0 = ys + (bb - Q) * BR;      // multiple of 8
            const float *sp[Q + 1];
            {
                int sl = oslot;
#pragma unroll
                for (int m = 0; m <= Q; ++m) {
                    sp[m] = ring + sl * SLOT + 2 * tid;
                    if (++sl == NS) sl = 0;
                }
            }
            float2 acc[BR];
#pragma unroll
            for (int i = 0; i < 2 * R + BR; ++i) {
                const float2 val = *reinterpret_cast<const float2 *>(sp[i / BR] + (i % BR) * TWP);
#pragma unroll
                for (int t = 0; t < BR; ++t) {
                    const int d = i - t - R;  // row offset of this ring row from the centre of output t
                    if (d == -R)
                        acc[t] = __fmul2_rn(make_float2(taps.t[R], taps.t[R]), val);
                    else if (d > -R && d <= R)
                        acc[t] = __ffma2_rn(make_float2(taps.t[d < 0 ? -d : d], taps.t[d < 0 ? -d : d]), val, acc[t]);
                }
            }
            ring_mbar_arrive(&bar_empty[oslot]);  // slot (bb - Q) is not read again
            float *o = dst + (size_t)yo0 * pitch + x;
            if (yo0 + BR <= ye) {
                if (col2) {
#pragma unroll
                    for (int t = 0; t < BR; ++t) *reinterpret_cast<float2 *>(o + (size_t)t * pitch) = acc[t];
                } else if (col1) {
#pragma unroll
                    for (int t = 0; t < BR; ++t) o[(size_t)t * pitch] = acc[t].x;
                }
            } else {  // last batch of an image whose height is not a multiple of 8
#pragma unroll
                for (int t = 0; t < BR; ++t) {
                    if (yo0 + t < ye) {
                        if (col2) *reinterpret_cast<float2 *>(o + (size_t)t * pitch) = acc[t];
                        else if (col1) o[(size_t)t * pitch] = acc[t].x;
                    }
                }
            }
            if (dec_col) {
                float *o2 = dst2 + (size_t)(yo0 >> 1) * pitch2 + (x >> 1);
#pragma unroll
                for (int t = 0; t < BR; t += 2)
                    if (yo0 + t < ye && ((yo0 + t) >> 1) < h2) o2[(size_t)(t >> 1) * pitch2] = acc[t].x;
            }
            if (++rslot == NS) { rslot = 0; rpar ^= 1; }
            if (++oslot == NS) oslot = 0;
        }
    }
}
