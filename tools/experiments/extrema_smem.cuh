// EXPERIMENT (round 2, lost): the 3x3x3 extrema scan fed through shared memory.
//
// Idea: the register-sliding scan (extrema_rows_kernel) keeps 12 loads per warp in flight and looked
// latency-bound, so stage the six Gaussian layers of a 240-column tile by cp.async through a 3-deep
// ring of 4-row batches and run the same per-row arithmetic from shared memory.
// Measured (ncu --set full, 18 x 1024 x 768 octave): 187.8 us and 135.8 M warp instructions against
// 122 us and ~75 M for extrema_rows_kernel -- the scan is bound by instruction issue (72 % busy), not by
// load latency, and the staging loop, the block barrier per batch and the idle warps of the last column
// tile add instructions.  Candidates were bit-identical (all parity tests passed).  Kept for the record;
// not compiled into the library.

// ---------------------------------------------------------------------------
// The same scan for the large octaves, fed through shared memory.  The register-sliding kernel
// above keeps only 12 loads per warp in flight (6 layers, one row ahead) and is bound by their
// latency (round 1: 2.8 TB/s on the 1024 x 768 octave).  Here a CTA (8 warps = 240 output columns)
// marches down a segment of rows in batches of 4; the batches of all 6 Gaussian layers are
// staged by 16 B cp.async through a 3-deep ring (two batches = 48 KB per CTA in flight, three CTAs
// per SM), and the warps run the SAME per-row arithmetic as above from shared memory (a 30-cycle
// load instead of an L2 / HBM one).  Every Gaussian value is still read once from HBM (24 B per
// pixel) + the 2-row / 2-column tile halos from L2.
// ---------------------------------------------------------------------------
constexpr int kExsCols = 240;                 // output columns per CTA (8 warps x 30)
constexpr int kExsW = 248;                    // staged floats per row: 242 needed + alignment slack, 62 chunks of 16 B
constexpr int kExsBR = 4;                     // rows per batch
constexpr int kExsStages = 3;
constexpr int kExsLayers = 6;                 // NI + 3 Gaussian layers (NI = 3)
constexpr size_t kExsSmem = (size_t)kExsStages * kExsLayers * kExsBR * kExsW * sizeof(float);   // 71 424 B

__global__ void __launch_bounds__(256)
extrema_smem_kernel(PyrView v, int o, int n_ct, int seg_rows, int border, float thresh, Candidate *__restrict__ cand,
                    int cand_cap, int32_t *__restrict__ counters)
{
    constexpr int NI = 3, ND = NI + 2, NL = NI + 3;
    constexpr int BR = kExsBR, S = kExsStages, W = kExsW;
    constexpr int CH = W / 4;                          // chunks per staged row
    constexpr int STG = NL * BR * W;                   // floats per stage
    extern __shared__ __align__(16) float exs[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int img = blockIdx.y;
    const int ct = blockIdx.x % n_ct, rs = blockIdx.x / n_ct;
    const int h = v.h[o], w = v.w[o], pitch = v.pitch[o];
    const size_t lstride = (size_t)v.n_img * h * pitch;
    const float *g0 = v.layer(o, 0, img);
    const int x_first = border + kExsCols * ct;        // first output column of the CTA
    const int xs0 = (x_first - 1) & ~3;                // staged column 0 (16 B aligned; border >= 1)
    const int ybeg = border + rs * seg_rows, yend = min(ybeg + seg_rows, h - border);
    if (ybeg >= yend) return;
    const int n_rows = yend - ybeg + 2;                // rows ybeg - 1 .. yend
    const int n_batches = (n_rows + BR - 1) / BR;

    auto issue = [&](int b) {
        float *dst = exs + (b % S) * STG;
        const int y0 = ybeg - 1 + b * BR;
        for (int i = threadIdx.x; i < NL * BR * CH; i += 256) {
            const int c = i % CH, lr = i / CH, r = lr % BR, l = lr / BR;
            const int x = xs0 + 4 * c;
            if (x + 4 <= pitch) {                      // chunk inside the row allocation (pitch % 8 == 0)
                const int y = min(y0 + r, h - 1);
                const float *src = g0 + (size_t)l * lstride + (size_t)y * pitch + x;
                const unsigned d = (unsigned)__cvta_generic_to_shared(dst + (l * BR + r) * W + 4 * c);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
            }
        }
    };
#pragma unroll
    for (int b = 0; b < S - 1; ++b) {
        if (b < n_batches) issue(b);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }

    const int x = x_first + 30 * warp - 1 + lane;       // this lane's column (lanes 0 and 31: halo)
    const int col = x - xs0;                            // < 3 + 242 <= W
    const bool out_lane = (lane >= 1) && (lane <= 30) && (x < w - border);
    float hmx[ND][3], hmn[ND][3], dprev[NI], dcur[NI];
#pragma unroll
    for (int l = 0; l < ND; ++l)
#pragma unroll
        for (int r = 0; r < 3; ++r) { hmx[l][r] = 0.f; hmn[l][r] = 0.f; }
#pragma unroll
    for (int l = 0; l < NI; ++l) { dprev[l] = 0.f; dcur[l] = 0.f; }

    for (int b = 0; b < n_batches; ++b) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(S - 2) : "memory");
        __syncthreads();                                // batch b is visible to all; everyone is done with batch b - 1
        if (b + S - 1 < n_batches) issue(b + S - 1);    // into the stage batch b - 1 was read from
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        const float *st = exs + (b % S) * STG + col;
        const int nr = min(BR, n_rows - b * BR);
        for (int r = 0; r < nr; ++r) {
            const int y = ybeg - 1 + b * BR + r;
            float g[NL];
#pragma unroll
            for (int l = 0; l < NL; ++l) g[l] = st[(l * BR + r) * W];
#pragma unroll
            for (int l = 0; l < ND; ++l) {
                const float d = __fsub_rn(g[l + 1], g[l]);
                const float lf = __shfl_up_sync(0xffffffffu, d, 1), rt = __shfl_down_sync(0xffffffffu, d, 1);
                hmx[l][0] = hmx[l][1]; hmx[l][1] = hmx[l][2]; hmx[l][2] = fmaxf(fmaxf(lf, d), rt);
                hmn[l][0] = hmn[l][1]; hmn[l][1] = hmn[l][2]; hmn[l][2] = fminf(fminf(lf, d), rt);
                if (l >= 1 && l <= NI) { dprev[l - 1] = dcur[l - 1]; dcur[l - 1] = d; }
            }
            if (y < ybeg + 1) continue;                 // the 3-row window is not full yet (warp-uniform)
            float M[ND], m[ND];
#pragma unroll
            for (int l = 0; l < ND; ++l) {
                M[l] = fmaxf(fmaxf(hmx[l][0], hmx[l][1]), hmx[l][2]);
                m[l] = fminf(fminf(hmn[l][0], hmn[l][1]), hmn[l][2]);
            }
#pragma unroll
            for (int li = 0; li < NI; ++li) {
                const int l = li + 1;
                const float val = dprev[li];
                const bool ext = out_lane && ((val > thresh && val >= M[l - 1] && val >= M[l] && val >= M[l + 1]) ||
                                              (val < -thresh && val <= m[l - 1] && val <= m[l] && val <= m[l + 1]));
                const unsigned mk = __ballot_sync(0xffffffffu, ext);
                if (mk) {
                    int base = 0;
                    const int leader = __ffs(mk) - 1;
                    if (lane == leader) {
                        base = atomicAdd(&counters[CNT_CAND], __popc(mk));
                        atomicAdd(&counters[CNT_HDR + img * CNT_PER_IMG + 0], __popc(mk));
                    }
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (ext) {
                        const int idx = base + __popc(mk & ((1u << lane) - 1u));
                        if (idx < cand_cap) {
                            Candidate cd;
                            cd.img_o_l = ((uint32_t)img << 16) | ((uint32_t)o << 8) | (uint32_t)l;
                            cd.yx = ((uint32_t)(y - 1) << 16) | (uint32_t)x;
                            cand[idx] = cd;
                        }
                    }
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

