// Experiment (not built): extrema scan with two columns per lane.  See README.md.
// Drop into csrc/detect.cu next to extrema_rows_kernel; host side: n_cg = (sw + (border & 1) + 59) / 60.
// Same scan with TWO adjacent columns per lane (float2 loads) and the three-row history kept in
// statically indexed registers (the row loop is unrolled by three), which removes the per-row
// register moves and halves the shuffles and loads per pixel: ~60 instead of ~170 instructions
// per pixel.  A warp owns 60 output columns (lanes 1..30; lanes 0 and 31 only supply neighbours).
template <int NI>
__global__ void __launch_bounds__(256)
extrema_rows2_kernel(PyrView v, const __grid_constant__ ExGroup grp, int border, float thresh,
                     Candidate *__restrict__ cand, int cand_cap, int32_t *__restrict__ counters)
{
    constexpr int ND = NI + 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int img = blockIdx.y;
    int gi = 0;
    while (gi + 1 < grp.n_oct && (int)blockIdx.x >= grp.blk_off[gi + 1]) ++gi;
    const int o = grp.o_first + gi;
    const int n_cg = grp.n_cg[gi], n_rs = grp.n_rs[gi], seg_rows = grp.seg_rows[gi];
    const int item = ((int)blockIdx.x - grp.blk_off[gi]) * 8 + warp;
    if (item >= n_cg * n_rs) return;  // whole warp
    const int cg = item % n_cg, rs = item / n_cg;
    const int h = v.h[o], w = v.w[o], pitch = v.pitch[o];
    const size_t lstride = (size_t)v.n_img * h * pitch;
    const float *g0 = v.layer(o, 0, img);
    const int x = (border & ~1) + 60 * cg - 2 + 2 * lane;     // first of this lane's two columns (even)
    const int xl = max(0, min(x, pitch - 2));                 // clamped load column (row allocation is `pitch` wide)
    const bool mid = (lane >= 1) && (lane <= 30);
    const bool out0 = mid && (x >= border) && (x < w - border);
    const bool out1 = mid && (x + 1 >= border) && (x + 1 < w - border);
    const int ybeg = border + rs * seg_rows, yend = min(ybeg + seg_rows, h - border);
    // history of the last three rows, slot = (row - (ybeg-1)) % 3
    float hmx[ND][3][2], hmn[ND][3][2], dv[NI][3][2];
#pragma unroll
    for (int l = 0; l < ND; ++l)
#pragma unroll
        for (int r = 0; r < 3; ++r) { hmx[l][r][0] = hmx[l][r][1] = 0.f; hmn[l][r][0] = hmn[l][r][1] = 0.f; }
#pragma unroll
    for (int l = 0; l < NI; ++l)
#pragma unroll
        for (int r = 0; r < 3; ++r) { dv[l][r][0] = dv[l][r][1] = 0.f; }
    const float *pl[ND + 1];
    float2 gn[ND + 1];
#pragma unroll
    for (int l = 0; l <= ND; ++l) {
        pl[l] = g0 + (size_t)l * lstride + (size_t)(ybeg - 1) * pitch + xl;
        gn[l] = __ldg(reinterpret_cast<const float2 *>(pl[l]));
    }
    for (int yb = ybeg - 1; yb <= yend; yb += 3) {
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int y = yb + s;
            if (y > yend) break;  // warp-uniform
            float2 g[ND + 1];
#pragma unroll
            for (int l = 0; l <= ND; ++l) g[l] = gn[l];
            if (y < yend) {
#pragma unroll
                for (int l = 0; l <= ND; ++l) {
                    pl[l] += pitch;
                    gn[l] = __ldg(reinterpret_cast<const float2 *>(pl[l]));
                }
            }
#pragma unroll
            for (int l = 0; l < ND; ++l) {
                const float d0 = __fsub_rn(g[l + 1].x, g[l].x), d1 = __fsub_rn(g[l + 1].y, g[l].y);
                const float lf = __shfl_up_sync(0xffffffffu, d1, 1), rt = __shfl_down_sync(0xffffffffu, d0, 1);
                const float mx01 = fmaxf(d0, d1), mn01 = fminf(d0, d1);
                hmx[l][s][0] = fmaxf(lf, mx01); hmx[l][s][1] = fmaxf(mx01, rt);
                hmn[l][s][0] = fminf(lf, mn01); hmn[l][s][1] = fminf(mn01, rt);
                if (l >= 1 && l <= NI) { dv[l - 1][s][0] = d0; dv[l - 1][s][1] = d1; }
            }
            if (y < ybeg + 1) continue;
            constexpr int SC[3] = {2, 0, 1};  // slot of the centre row y-1
            const int sc = SC[s];
            float M[ND][2], m[ND][2];
#pragma unroll
            for (int l = 0; l < ND; ++l)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    M[l][j] = fmaxf(fmaxf(hmx[l][0][j], hmx[l][1][j]), hmx[l][2][j]);
                    m[l][j] = fminf(fminf(hmn[l][0][j], hmn[l][1][j]), hmn[l][2][j]);
                }
#pragma unroll
            for (int li = 0; li < NI; ++li) {
                const int l = li + 1;
                bool ext[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float val = dv[li][sc][j];
                    const float M3 = fmaxf(fmaxf(M[l - 1][j], M[l][j]), M[l + 1][j]);
                    const float m3 = fminf(fminf(m[l - 1][j], m[l][j]), m[l + 1][j]);
                    ext[j] = (j ? out1 : out0) && ((val > thresh && val >= M3) || (val < -thresh && val <= m3));
                }
                const unsigned mk0 = __ballot_sync(0xffffffffu, ext[0]), mk1 = __ballot_sync(0xffffffffu, ext[1]);
                if (mk0 | mk1) {
                    const int n0 = __popc(mk0), n1 = __popc(mk1);
                    int base = 0;
                    if (lane == 0) {
                        base = atomicAdd(&counters[CNT_CAND], n0 + n1);
                        atomicAdd(&counters[CNT_HDR + img * CNT_PER_IMG + 0], n0 + n1);
                    }
                    base = __shfl_sync(0xffffffffu, base, 0);
                    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (ext[j]) {
                            const int idx = base + (j ? n0 + __popc(mk1 & lt) : __popc(mk0 & lt));
                            if (idx < cand_cap) {
                                Candidate cd;
                                cd.img_o_l = ((uint32_t)img << 16) | ((uint32_t)o << 8) | (uint32_t)l;
                                cd.yx = ((uint32_t)(y - 1) << 16) | (uint32_t)(x + j);
                                cand[idx] = cd;
                            }
                        }
                    }
                }
            }
        }
    }
}

