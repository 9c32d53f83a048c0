// Sparse stage of the SIFT path: fused DoG + 3x3x3 extrema scan with
// warp-ballot compaction, quadratic-fit refinement, warp-per-keypoint
// orientation histograms and 4x4x8 descriptors, ordering + de-duplication.
//
// Replaces (all in /root/reference/sift_impl.py): :100-111 (DoG, fused, never
// materialised), :117-163 find_scale_space_extrema / is_pixel_an_extremum,
// :169-240 localize_extremum_via_quadratic_fit + gradient / Hessian,
// :246-293 compute_keypoints_with_orientations, :299-327 compare_keypoints /
// remove_duplicate_keypoints, :333-343 convert_keypoints_to_input_image_size,
// :349-526 unpack_octave / generate_descriptors.
//
// This file is compiled with --fmad=false: numpy rounds every float32
// operation once, so must we (the float64 solve then agrees bit for bit with
// a no-FMA CPU evaluation of the same expressions).
#include <cub/device/device_merge_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <math.h>
#include "common.cuh"

namespace b200 {

#include "fastmath.cuh"


PyrView make_view(const Pyramid &p)
{
    PyrView v;
    v.n_img = p.n_img;
    v.n_oct = p.n_oct;
    v.n_layers = p.n_layers;
    for (int o = 0; o < p.n_oct; ++o) {
        v.h[o] = p.h[o];
        v.w[o] = p.w[o];
        v.pitch[o] = p.pitch[o];
        v.oct[o] = p.base + p.oct_off[o];
    }
    return v;
}

PyrView make_view(const b200sift_ctx *c) { return make_view(c->pyr); }

DetectParams make_detect_params(const b200sift_params &p)
{
    DetectParams d;
    d.num_intervals = p.num_intervals;
    d.border = p.image_border_width;
    d.max_iter = p.max_iter;
    d.ori_bins = p.ori_bins;
    d.dog_thresh = (float)floor(0.5 * p.contrast_threshold / p.num_intervals * 255);  // sift_impl.py:122
    d.contrast_thr_f = (float)p.contrast_threshold;
    d.eigen_ratio_f = (float)p.eigen_ratio;
    d.sigma_f = (float)p.sigma;
    d.radius_factor_f = (float)p.radius_factor;
    d.scale_factor = p.scale_factor;
    d.peak_ratio = p.peak_ratio;
    d.scale_multiplier_half = p.scale_multiplier * 0.5;
    d.descriptor_max_value_f = (float)p.descriptor_max_value;
    return d;
}

// ---------------------------------------------------------------------------
// fused DoG + extrema scan (sift_impl.py:109, :124-132, :143-163)
// One CTA = 32x16 scan pixels of one image at one octave.  The DoG stack of
// the tile (+1 halo) is formed in shared memory from the Gaussian layers
// (float32 subtraction, never written to HBM: 24 B read per pixel for six
// layers); each thread tests its pixels in the num_intervals middle layers;
// hits are compacted with a warp ballot and one atomic per warp.
// ---------------------------------------------------------------------------
// kDog: `v` already holds DoG layers (find_scale_space_extrema with a caller-supplied dog_images,
// sift_impl.py:117-118) and they are loaded as they are.
constexpr int kExTW = 32, kExTH = 16;

template <bool kDog>
__global__ void __launch_bounds__(256)
extrema_kernel(PyrView v, int o, int border, int num_intervals, float thresh, Candidate *__restrict__ cand,
               int cand_cap, int32_t *__restrict__ counters)
{
    extern __shared__ float dog_s[];  // [n_dog][(TH+2)*(TW+2)]
    constexpr int SW = kExTW + 2, SH = kExTH + 2, SN = SW * SH;
    const int img = blockIdx.z;
    const int h = v.h[o], w = v.w[o], pitch = v.pitch[o];
    const int n_dog = kDog ? v.n_layers : v.n_layers - 1;
    const int x0 = border + blockIdx.x * kExTW, y0 = border + blockIdx.y * kExTH;
    const size_t lstride = (size_t)v.n_img * h * pitch;  // layer stride
    const float *g0 = v.layer(o, 0, img);
    for (int i = threadIdx.x; i < SN; i += 256) {
        const int yy = i / SW, xx = i - yy * SW;
        const int y = min(max(y0 - 1 + yy, 0), h - 1), x = min(max(x0 - 1 + xx, 0), w - 1);
        const float *p = g0 + (size_t)y * pitch + x;
        float prev = *p;
        for (int l = 0; l < n_dog; ++l) {
            if (kDog) {
                dog_s[l * SN + i] = p[(size_t)l * lstride];
            } else {
                p += lstride;
                const float cur = *p;
                dog_s[l * SN + i] = __fsub_rn(cur, prev);
                prev = cur;
            }
        }
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const unsigned lane = tx;
    for (int k = 0; k < kExTH / 8; ++k) {
        const int yl = ty + 8 * k;
        const int y = y0 + yl, x = x0 + tx;
        const bool inside = (y < h - border) && (x < w - border);
        const int ctr = (yl + 1) * SW + tx + 1;
        for (int l = 1; l <= num_intervals; ++l) {
            bool ext = false;
            if (inside) {
                const float *c = dog_s + l * SN + ctr;
                const float val = *c;
                if (fabsf(val) > thresh) {
                    ext = true;
                    if (val > 0.f) {
#pragma unroll
                        for (int dl = -1; dl <= 1; ++dl)
#pragma unroll
                            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                                for (int dx = -1; dx <= 1; ++dx) ext = ext && (val >= c[dl * SN + dy * SW + dx]);
                    } else {
#pragma unroll
                        for (int dl = -1; dl <= 1; ++dl)
#pragma unroll
                            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                                for (int dx = -1; dx <= 1; ++dx) ext = ext && (val <= c[dl * SN + dy * SW + dx]);
                    }
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, ext);
            if (m) {
                int base = 0;
                if (lane == (unsigned)(__ffs(m) - 1)) {
                    base = atomicAdd(&counters[CNT_CAND], __popc(m));
                    atomicAdd(&counters[CNT_HDR + img * CNT_PER_IMG + 0], __popc(m));
                }
                base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                if (ext) {
                    const int idx = base + __popc(m & ((1u << lane) - 1u));
                    if (idx < cand_cap) {
                        Candidate cd;
                        cd.img_o_l = ((uint32_t)img << 16) | ((uint32_t)o << 8) | (uint32_t)l;
                        cd.yx = ((uint32_t)y << 16) | (uint32_t)x;
                        cand[idx] = cd;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Register-sliding form of the same scan (used for the reference's
// num_intervals = 3): no shared memory.  A warp owns 30 output columns (+1
// halo lane each side) and marches down a segment of rows; per row every lane
// loads its pixel of the NI+3 Gaussian layers (coalesced 128 B per layer),
// forms the NI+2 DoG values in registers, gets the left / right neighbours by
// warp shuffle and keeps the horizontal 3-max / 3-min of the last three rows.
// "val >= all 26 neighbours" (ties pass, :151-156) is then
//   val >= max3x3(layer-1) && val >= max3x3(layer) && val >= max3x3(layer+1)
// (the centre is part of max3x3(layer), so that term is an equality test), and
// likewise with min for negative values.  Each Gaussian value is read once
// from HBM (24 B per pixel for 6 layers, + halo rows / lanes from L2).
// ---------------------------------------------------------------------------
// One launch scans a GROUP of consecutive octaves (the small octaves of the pyramid tail are
// ready at the same time; one launch per octave there is pure latency on the critical path).
struct ExGroup {
    int o_first, n_oct;
    int blk_off[kMaxOctaves + 1];  // first block of each octave of the group
    int n_cg[kMaxOctaves], n_rs[kMaxOctaves], seg_rows[kMaxOctaves];
};

template <int NI>
__global__ void __launch_bounds__(256)
extrema_rows_kernel(PyrView v, const __grid_constant__ ExGroup grp, int border, float thresh,
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
    const int x = border + 30 * cg - 1 + lane;
    const int xc = min(x, w - 1);
    const bool out_lane = (lane >= 1) && (lane <= 30) && (x < w - border);
    const int ybeg = border + rs * seg_rows, yend = min(ybeg + seg_rows, h - border);
    float hmx[ND][3], hmn[ND][3], dprev[NI], dcur[NI];
#pragma unroll
    for (int l = 0; l < ND; ++l)
#pragma unroll
        for (int r = 0; r < 3; ++r) { hmx[l][r] = 0.f; hmn[l][r] = 0.f; }
#pragma unroll
    for (int l = 0; l < NI; ++l) { dprev[l] = 0.f; dcur[l] = 0.f; }
    // one pointer per Gaussian layer, stepped by `pitch` per row (no 64-bit address arithmetic in
    // the loop), and the loads of row y+1 are issued before row y is evaluated
    const float *pl[ND + 1];
    float gn[ND + 1];
#pragma unroll
    for (int l = 0; l <= ND; ++l) {
        pl[l] = g0 + (size_t)l * lstride + (size_t)(ybeg - 1) * pitch + xc;
        gn[l] = __ldg(pl[l]);
    }
    for (int y = ybeg - 1; y <= yend; ++y) {
        float g[ND + 1];
#pragma unroll
        for (int l = 0; l <= ND; ++l) g[l] = gn[l];
        if (y < yend) {
#pragma unroll
            for (int l = 0; l <= ND; ++l) {
                pl[l] += pitch;
                gn[l] = __ldg(pl[l]);
            }
        }
#pragma unroll
        for (int l = 0; l < ND; ++l) {
            const float d = __fsub_rn(g[l + 1], g[l]);
            const float lf = __shfl_up_sync(0xffffffffu, d, 1), rt = __shfl_down_sync(0xffffffffu, d, 1);
            hmx[l][0] = hmx[l][1]; hmx[l][1] = hmx[l][2]; hmx[l][2] = fmaxf(fmaxf(lf, d), rt);
            hmn[l][0] = hmn[l][1]; hmn[l][1] = hmn[l][2]; hmn[l][2] = fminf(fminf(lf, d), rt);
            if (l >= 1 && l <= NI) { dprev[l - 1] = dcur[l - 1]; dcur[l - 1] = d; }
        }
        if (y < ybeg + 1) continue;
        // most rows of a warp have no pixel above the contrast pre-threshold (:122,148) in any of the
        // NI layers: skip the 3x3x3 comparisons there (the running row maxima above are kept up)
        bool any_big = false;
#pragma unroll
        for (int li = 0; li < NI; ++li) any_big |= fabsf(dprev[li]) > thresh;
        if (!__any_sync(0xffffffffu, any_big && out_lane)) continue;
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

// ---------------------------------------------------------------------------
// quadratic-fit refinement (sift_impl.py:169-240), one thread per candidate
// ---------------------------------------------------------------------------

// x = pinv(H) g for symmetric 3x3 H in float64 (cyclic Jacobi): the
// minimum-norm least-squares solution np.linalg.lstsq(hess, grad, rcond=None)
// returns (LAPACK gelsd, cut-off eps*3*|lambda|max).
__device__ void sym3_pinv_solve(const double H[3][3], const double g[3], double x[3])
{
    double a[3][3], v[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            a[i][j] = H[i][j];
            v[i][j] = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 60; ++sweep) {
        // converged once the off-diagonal part is far below one ulp of the diagonal (Jacobi
        // converges quadratically, so this costs at most one sweep more than needed; waiting for
        // an exact zero would spin through all 60 sweeps and stall the whole warp)
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off <= 1e-25 * (fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]))) break;
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    const double lmax = fmax(fabs(a[0][0]), fmax(fabs(a[1][1]), fabs(a[2][2])));
    const double cut = 2.220446049250313e-16 * 3.0 * lmax;
    x[0] = x[1] = x[2] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double lam = a[i][i];
        if (!(fabs(lam) > cut)) continue;
        const double proj = (v[0][i] * g[0] + v[1][i] * g[1] + v[2][i] * g[2]) / lam;
        x[0] += proj * v[0][i];
        x[1] += proj * v[1][i];
        x[2] += proj * v[2][i];
    }
}

// Fast path of the same solve: H is practically always well conditioned (smallest
// singular-value ratio seen on the reference's sets: 6e-5), where the minimum-norm solution is
// the ordinary one.  Adjugate / determinant in float64 (relative error ~cond*1e-16, invisible
// after the cast to float32); anything close to singular takes the Jacobi pseudo-inverse above.
__device__ __forceinline__ void sym3_solve(const double H[3][3], const double g[3], double x[3])
{
    const double a = H[0][0], b = H[0][1], c = H[0][2], d = H[1][1], e = H[1][2], f = H[2][2];
    const double A = d * f - e * e, B = c * e - b * f, C = b * e - c * d;
    const double det = a * A + b * B + c * C;
    const double m = fmax(fmax(fabs(a), fabs(d)), fmax(fabs(f), fmax(fabs(b), fmax(fabs(c), fabs(e)))));
    if (!(fabs(det) > 1e-9 * m * m * m)) {
        sym3_pinv_solve(H, g, x);
        return;
    }
    const double D = a * f - c * c, E = b * c - a * e, F = a * d - b * b;
    const double inv = 1.0 / det;
    x[0] = (A * g[0] + B * g[1] + C * g[2]) * inv;
    x[1] = (B * g[0] + D * g[1] + E * g[2]) * inv;
    x[2] = (C * g[0] + E * g[1] + F * g[2]) * inv;
}

// kDog: `v` holds DoG layers (a caller-supplied dog_images / dog_octave) instead of Gaussian layers.
// direct_n >= 0 is the per-call form of the stage API (localize_extremum_via_quadratic_fit on a
// caller's candidates): exactly direct_n candidates, result i written to loc[i] (img_o_l =
// 0xFFFFFFFF when the reference returns None), no counters; direct_single_octave: `v` holds only
// the candidate's octave (as octave 0) while the candidate's own octave number scales the result.
template <bool kDog>
__global__ void __launch_bounds__(128)
refine_kernel(PyrView v, DetectParams dp, const Candidate *__restrict__ cand, int cand_cap,
              Localized *__restrict__ loc, int loc_cap, int32_t *__restrict__ counters, int direct_n,
              int direct_single_octave)
{
    const bool direct = direct_n >= 0;
    const int n = direct ? direct_n : min(counters[CNT_CAND], cand_cap);
    for (int ci = blockIdx.x * blockDim.x + threadIdx.x; ci < n; ci += gridDim.x * blockDim.x) {
        const Candidate cd = cand[ci];
        const int img = cd.img_o_l >> 16, o = (cd.img_o_l >> 8) & 255;
        int layer = cd.img_o_l & 255;
        int y = cd.yx >> 16, x = cd.yx & 0xffff;
        const uint64_t order = ((((uint64_t)o << 4) | (uint64_t)layer) << 30) | ((uint64_t)y << 15) | (uint64_t)x;
        const int ov = direct_single_octave ? 0 : o;
        const int h = v.h[ov], w = v.w[ov], pitch = v.pitch[ov];
        const size_t lstride = (size_t)v.n_img * h * pitch;
        const float *g0 = v.layer(ov, 0, img);
        if (direct) {
            Localized none;
            none.x = none.y = none.size = none.response = 0.f;
            none.octave_packed = 0; none.img_o_l = 0xFFFFFFFFu; none.order = order;
            loc[ci] = none;
        }
        float cube[3][3][3], grad[3], hess[3][3], upd[3];
        bool alive = true;
        for (int it = 0; it < dp.max_iter; ++it) {
            // cube[s][j][i] = DoG(layer-1+s)(y-1+j, x-1+i) / 255  (float32)
            const float *p = g0 + (size_t)(layer - 1) * lstride + (size_t)(y - 1) * pitch + (x - 1);
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float *q = p + j * pitch + i;
                    if (kDog) {
                        cube[0][j][i] = __fdiv_rn(q[0], 255.f);
                        cube[1][j][i] = __fdiv_rn(q[lstride], 255.f);
                        cube[2][j][i] = __fdiv_rn(q[2 * lstride], 255.f);
                    } else {
                        const float a0 = q[0], a1 = q[lstride], a2 = q[2 * lstride], a3 = q[3 * lstride];
                        cube[0][j][i] = __fdiv_rn(__fsub_rn(a1, a0), 255.f);
                        cube[1][j][i] = __fdiv_rn(__fsub_rn(a2, a1), 255.f);
                        cube[2][j][i] = __fdiv_rn(__fsub_rn(a3, a2), 255.f);
                    }
                }
            grad[0] = 0.5f * (cube[1][1][2] - cube[1][1][0]);
            grad[1] = 0.5f * (cube[1][2][1] - cube[1][0][1]);
            grad[2] = 0.5f * (cube[2][1][1] - cube[0][1][1]);
            const float c = cube[1][1][1];
            const float dxx = cube[1][1][2] - 2 * c + cube[1][1][0];
            const float dyy = cube[1][2][1] - 2 * c + cube[1][0][1];
            const float dss = cube[2][1][1] - 2 * c + cube[0][1][1];
            const float dxy = 0.25f * (cube[1][2][2] - cube[1][2][0] - cube[1][0][2] + cube[1][0][0]);
            const float dxs = 0.25f * (cube[2][1][2] - cube[2][1][0] - cube[0][1][2] + cube[0][1][0]);
            const float dys = 0.25f * (cube[2][2][1] - cube[2][0][1] - cube[0][2][1] + cube[0][0][1]);
            hess[0][0] = dxx; hess[0][1] = dxy; hess[0][2] = dxs;
            hess[1][0] = dxy; hess[1][1] = dyy; hess[1][2] = dys;
            hess[2][0] = dxs; hess[2][1] = dys; hess[2][2] = dss;
            double Hd[3][3], gd[3], xd[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                gd[i] = grad[i];
#pragma unroll
                for (int j = 0; j < 3; ++j) Hd[i][j] = hess[i][j];
            }
            sym3_solve(Hd, gd, xd);
#pragma unroll
            for (int i = 0; i < 3; ++i) upd[i] = -(float)xd[i];
            if (fabsf(upd[0]) < 0.5f && fabsf(upd[1]) < 0.5f && fabsf(upd[2]) < 0.5f) break;
            x += (int)rintf(upd[0]);
            y += (int)rintf(upd[1]);
            layer += (int)rintf(upd[2]);
            if (y < dp.border || y >= h - dp.border || x < dp.border || x >= w - dp.border || layer < 1 ||
                layer > dp.num_intervals) {
                alive = false;
                break;
            }
            // no "failed to converge" rejection in the reference: after max_iter
            // moves the stale cube / grad / hess / upd are used with the moved x, y, layer
        }
        if (!alive) continue;
        const float p0 = grad[0] * upd[0], p1 = grad[1] * upd[1], p2 = grad[2] * upd[2];
        const float dot = (float)((double)p0 + (double)p1 + (double)p2);  // numpy float32 dot of 3
        const float val = cube[1][1][1] + 0.5f * dot;
        if (fabsf(val) * (float)dp.num_intervals < dp.contrast_thr_f) continue;
        const float tr = hess[0][0] + hess[1][1];
        // np.linalg.det (LU with partial pivoting in float64, then float32)
        const double a = hess[0][0], b = hess[0][1], cc = hess[1][0], d = hess[1][1];
        double det;
        if (fabs(a) >= fabs(cc)) {
            if (a == 0.0) det = 0.0;
            else { const double l = cc / a; det = a * (d - l * b); }
        } else {
            const double l = a / cc;
            det = -(cc * (b - l * d));
        }
        const float detf = (float)det;
        const float er = dp.eigen_ratio_f;
        if (detf <= 0.f || er * (tr * tr) >= ((er + 1.f) * (er + 1.f)) * detf) continue;
        Localized L;
        const float sc = (float)(1 << o);
        L.x = ((float)x + upd[0]) * sc;
        L.y = ((float)y + upd[1]) * sc;
        L.octave_packed = o + layer * 256 + (int)rintf((upd[2] + 0.5f) * 255.f) * 65536;
        const float e = ((float)layer + upd[2]) / (float)dp.num_intervals;
        L.size = dp.sigma_f * (float)exp2((double)e) * (float)(1 << (o + 1));
        L.response = fabsf(val);
        L.img_o_l = ((uint32_t)img << 16) | ((uint32_t)o << 8) | (uint32_t)layer;
        L.order = order;
        if (direct) {
            loc[ci] = L;
            continue;
        }
        const int slot = atomicAdd(&counters[CNT_LOC], 1);
        atomicAdd(&counters[CNT_HDR + img * CNT_PER_IMG + 1], 1);
        if (slot < loc_cap) loc[slot] = L;
    }
}

#define B200_RAD2DEGF (180.0f / 3.141592653589793238462643383279502884f)

__device__ __forceinline__ float mod360f(float a)  // np.float32 % 360 for |a| < 360
{
    if (a < 0.f) a += 360.f;
    else if (a == 0.f) a = 0.f;
    return a;
}

// ---------------------------------------------------------------------------
// orientation assignment (sift_impl.py:246-293), one warp per localized
// extremum.  Each lane accumulates its share of the (2r+1)^2 window into a
// private float64 36-bin histogram in shared memory ([bin][lane], bank
// conflict free); the lanes' histograms are then summed in a fixed order, so
// the result is deterministic (no atomics).
// ---------------------------------------------------------------------------
// 2^-k as a float (0 <= k <= 126): dividing by 2^k and multiplying by this give the same float
__device__ __forceinline__ float pow2_neg(int k) { return __int_as_float((127 - k) << 23); }

constexpr int kOriWarps = 4;
constexpr int kOriMaxBins = 36;
// five CTAs per SM (36.9 KB of private bins each, 94 registers); a sixth (80 registers) measured slower
__global__ void __launch_bounds__(kOriWarps * 32, 5)
orient_kernel(PyrView v, DetectParams dp, const Localized *__restrict__ loc, int loc_cap,
              RawKeypoint *__restrict__ raw, int raw_cap, int32_t *__restrict__ counters, int direct_n,
              int32_t *__restrict__ direct_counts, int32_t *__restrict__ class_idx)
{
    // direct_n >= 0: the per-call form of the stage API (compute_keypoints_with_orientations on a
    // caller's keypoints and ONE Gaussian image, held by `v` as octave 0 / layer 0): the peaks of
    // keypoint i go to raw[i * ori_bins ...] in ascending bin order, their number to direct_counts[i].
    const bool direct = direct_n >= 0;
    __shared__ __align__(16) double hist_s[kOriWarps][kOriMaxBins][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nb = dp.ori_bins;  // <= 36
    const int n = direct ? direct_n : min(counters[CNT_LOC], loc_cap);
    double(*hist)[32] = hist_s[wib];
    // the summed and the smoothed histogram reuse the first rows of the private bins once those are summed
    double *raw_h = &hist[0][0], *smooth_h = &hist[2][0];
    static_assert(kOriMaxBins <= 36, "bins past 31 are summed in one round of four");
    // dynamic work queue: windows are (2r+1)^2 with r = 7..17.  The 128 B lines of the NEXT item's window
    // are requested into L2 while this one is evaluated (the layer was last touched by the blur kernels
    // and mostly left L2 since).
    // The queue itself runs TWO items ahead: `take` only issues the atomic (its value is broadcast after the
    // pixel loop of the current keypoint, when it has long arrived), the Localized record of that item
    // is loaded before the histogram epilogue and first used at the top of the next keypoint -- the
    // warp never waits for an atomic followed by a dependent load before it can start its work.
    auto take = [&]() -> int {
        int t = 0;
        if (lane == 0) t = atomicAdd(&counters[CNT_WORK_ORI], 1);
        return t;   // valid in lane 0
    };
    auto prefetch_window = [&](const Localized &P) {
        const int po = (P.img_o_l >> 8) & 255, pov = direct ? 0 : po;
        const int ph = v.h[pov], pw = v.w[pov], pp = v.pitch[pov];
        const float *pimg = v.layer(pov, direct ? 0 : (int)(P.img_o_l & 255), P.img_o_l >> 16);
        const float pscale = (float)(dp.scale_factor * (double)P.size) * pow2_neg(po + 1);
        const int prad = (int)fminf(rintf(dp.radius_factor_f * pscale), 64.f);
        const int pcy = (int)rintf(P.y * pow2_neg(po)), pcx = (int)rintf(P.x * pow2_neg(po));
        const int y0 = max(pcy - prad - 1, 0), y1 = min(pcy + prad + 1, ph - 1);
        const int x0 = max(pcx - prad - 1, 0), x1 = min(pcx + prad + 1, pw - 1);
        if (y1 < y0 || x1 < x0) return;
        for (int yy = y0 + lane; yy <= y1; yy += 32) {
            const uintptr_t a0 = reinterpret_cast<uintptr_t>(pimg + (size_t)yy * pp + x0) & ~(uintptr_t)127;
            const uintptr_t a1 = reinterpret_cast<uintptr_t>(pimg + (size_t)yy * pp + x1) & ~(uintptr_t)127;
            for (uintptr_t a = a0; a <= a1; a += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
    };
    int li = __shfl_sync(0xffffffffu, take(), 0);
    int li_next = __shfl_sync(0xffffffffu, take(), 0);
    Localized L, Ln;
    if (li < n) L = loc[li];
    if (li_next < n) Ln = loc[li_next];
    while (li < n) {
        const int t_after = take();   // item after the next one
        if (li_next < n) prefetch_window(Ln);
        const int img = L.img_o_l >> 16, o = (L.img_o_l >> 8) & 255, layer = L.img_o_l & 255;
        const int ov = direct ? 0 : o;
        const int h = v.h[ov], w = v.w[ov], pitch = v.pitch[ov];
        const float *gimg = v.layer(ov, direct ? 0 : layer, img);
        // x / 2^o (:250-251, :259-260) as the exact product x * 2^-o
        const float scale = (float)(dp.scale_factor * (double)L.size) * pow2_neg(o + 1);
        const int radius = (int)fminf(rintf(dp.radius_factor_f * scale), 1048576.f);
        const float weight_fac = -0.5f / (scale * scale);
        const float weight_fac2 = weight_fac * 1.4426950408889634f;   // exp(w d) = 2^(w log2(e) d)
        const int cy = (int)rintf(L.y * pow2_neg(o));
        const int cx = (int)rintf(L.x * pow2_neg(o));
        {   // zero the nb x 32 private bins (256 B per bin) with 16 B stores
            float4 *h4 = reinterpret_cast<float4 *>(&hist[0][0]);
            for (int i = lane; i < nb * 16; i += 32) h4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
        // the window clipped to the pixels the reference does not skip (:262-264)
        const int ylo = max(cy - radius, 1), yhi = min(cy + radius, h - 2);
        const int xlo = max(cx - radius, 1), xhi = min(cx + radius, w - 2);
        const int nx = xhi - xlo + 1, ny = yhi - ylo + 1;
        const int total = (nx > 0 && ny > 0) ? nx * ny : 0;
        int yy = lane / max(nx, 1), xx = lane - yy * max(nx, 1);
        const int q32 = 32 / max(nx, 1), r32 = 32 - q32 * max(nx, 1);   // 32 pixels further = q32 rows and r32 columns
        const float bin_scale = (float)nb * (1.f / 360.f);
        // Two pixels per lane and iteration (straight-line code: the two sqrt / atan2 / exp chains
        // overlap); the histogram updates stay in pixel order.
        struct Stage {           // one iteration in flight: pixel coordinates and the four neighbours
            int y[2], x[2];
            bool live[2];
            float g[2][4];
        };
        auto issue = [&](Stage &S, int idx) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                S.live[u] = idx + 32 * u < total;
                S.y[u] = S.live[u] ? ylo + yy : ylo;
                S.x[u] = S.live[u] ? xlo + xx : xlo;
                xx += r32;
                yy += q32;
                if (xx >= nx) { xx -= nx; ++yy; }
                const float *p = gimg + (size_t)S.y[u] * pitch + S.x[u];
                S.g[u][0] = __ldg(p + 1);
                S.g[u][1] = __ldg(p - 1);
                S.g[u][2] = __ldg(p - pitch);
                S.g[u][3] = __ldg(p + pitch);
            }
        };
        // evaluate the iteration held by S; its registers are refilled with iteration idx_next first
        auto consume = [&](Stage &S, int idx_next) {
            int cyv[2], cxv[2];
            bool live[2];
            float g[2][4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                cyv[u] = S.y[u]; cxv[u] = S.x[u]; live[u] = S.live[u];
#pragma unroll
                for (int k = 0; k < 4; ++k) g[u][k] = S.g[u][k];
            }
            if (idx_next < total) issue(S, idx_next);
            int bin[2];
            float val[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int dy = cyv[u] - cy, dx = cxv[u] - cx;
                const float gx = g[u][0] - g[u][1];
                const float gy = g[u][2] - g[u][3];
                const float mag = fast_sqrt(__fmaf_rn(gx, gx, gy * gy));
                const float ang = atan2_deg_fast(gy, gx);
                const float wgt = fast_ex2(weight_fac2 * (float)(dx * dx + dy * dy));
                int bi = (int)rintf(ang * bin_scale);          // ang in [0, 360]: bi in [0, nb]
                if (bi >= nb) bi -= nb;                         // == bi % nb (:279)
                bin[u] = live[u] ? bi : -1;
                val[u] = wgt * mag;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (bin[u] >= 0) hist[bin[u]][lane] += (double)val[u];
        };
        // TWO iterations in flight (stages A, B; the issue order is the pixel order): one iteration of
        // arithmetic does not cover an L2 round trip
        Stage A, B;
        if (total > 0) issue(A, lane);
        if (lane + 64 < total) issue(B, lane + 64);
        for (int idx = lane; idx < total; idx += 128) {
            consume(A, idx + 128);
            if (idx + 64 < total) consume(B, idx + 192);
        }
        const int li_after = __shfl_sync(0xffffffffu, t_after, 0);
        Localized La;
        if (li_after < n) La = loc[li_after];
        __syncwarp();
        double r_main = 0.0;
        if (lane < nb) {   // bins 0..31: one per lane
            const int b = lane;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;  // fixed order, 4 chains
#pragma unroll
            for (int l = 0; l < 32; l += 4) {
                s0 += hist[b][(l + lane) & 31];
                s1 += hist[b][(l + 1 + lane) & 31];
                s2 += hist[b][(l + 2 + lane) & 31];
                s3 += hist[b][(l + 3 + lane) & 31];
            }
            r_main = (s0 + s1) + (s2 + s3);
        }
        double r_past = 0.0;
        const int b_past = 32 + (lane >> 3);
        if (nb > 32) {   // bins past 31 (four of the 36): eight lanes per bin, fixed tree
            const int sub = lane & 7;
            if (b_past < nb)
                r_past = (hist[b_past][4 * sub] + hist[b_past][4 * sub + 1]) +
                         (hist[b_past][4 * sub + 2] + hist[b_past][4 * sub + 3]);
            r_past += __shfl_xor_sync(0xffffffffu, r_past, 1);
            r_past += __shfl_xor_sync(0xffffffffu, r_past, 2);
            r_past += __shfl_xor_sync(0xffffffffu, r_past, 4);
        }
        __syncwarp();   // every private bin has been read: its first rows now hold the sums
        if (lane < nb) raw_h[lane] = r_main;
        if (b_past < nb && (lane & 7) == 0) raw_h[b_past] = r_past;
        __syncwarp();
        double mx = -1.0;
        for (int b = lane; b < nb; b += 32) {
            const double *r = raw_h;
            // circular neighbours (:282-283); nb >= 4, so one conditional wrap each
            const int m1 = b >= 1 ? b - 1 : b - 1 + nb, m2 = b >= 2 ? b - 2 : b - 2 + nb;
            const int p1 = b + 1 < nb ? b + 1 : b + 1 - nb, p2 = b + 2 < nb ? b + 2 : b + 2 - nb;
            const double s = (6 * r[b] + 4 * (r[m1] + r[p1]) + r[m2] + r[p2]) / 16.;
            smooth_h[b] = s;
            mx = fmax(mx, s);
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, sft));
        __syncwarp();
        // peaks in ascending bin order; rounds of 32 bins
        int emitted = 0;
        for (int b0 = 0; b0 < nb; b0 += 32) {
            const int b = b0 + lane;
            bool peak = false;
            double sl = 0, sr = 0, sc = 0;
            if (b < nb) {
                const double *s = smooth_h;
                sl = s[b >= 1 ? b - 1 : nb - 1];
                sr = s[b + 1 < nb ? b + 1 : 0];
                sc = s[b];
                peak = (sc > sl) && (sc > sr) && (sc >= dp.peak_ratio * mx);
            }
            const unsigned m = __ballot_sync(0xffffffffu, peak);
            if (m) {
                // work class of the descriptor kernel: by window half-width (sift_impl.py:386-388)
                const float hw = (float)dp.scale_multiplier_half * L.size * pow2_neg(o);   // hist_width (:386) in octave pixels
                const float half_w = hw * 3.5355339f;
                const int cls = half_w >= 40.f ? 0 : half_w >= 32.f ? 1 : half_w >= 26.f ? 2 : half_w >= 21.f ? 3 : 4;
                int base = 0, cb = 0;
                if (direct) {
                    base = li * nb + emitted;
                } else {
                    if (lane == __ffs(m) - 1) {   // both slot atomics go out before either answer is awaited
                        base = atomicAdd(&counters[CNT_RAW], __popc(m));
                        if (class_idx) cb = atomicAdd(&counters[CNT_CLASS + cls], __popc(m));
                        atomicAdd(&counters[CNT_HDR + img * CNT_PER_IMG + 2], __popc(m));
                    }
                    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                    cb = __shfl_sync(0xffffffffu, cb, __ffs(m) - 1);
                }
                if (class_idx) {
                    if (peak) {
                        const int slot = base + __popc(m & ((1u << lane) - 1u));
                        if (slot < raw_cap) class_idx[(size_t)cls * raw_cap + cb + __popc(m & ((1u << lane) - 1u))] = slot;
                    }
                }
                if (peak) {
                    const int slot = base + __popc(m & ((1u << lane) - 1u));
                    double t = (double)b + 0.5 * (sl - sr) / (sl - 2 * sc + sr);
                    // t % nb (:288): a strict peak has |t - b| <= 0.5, so -0.5 <= t < nb and the
                    // remainder is t itself or, for t < 0, t + nb
                    double interp = t < 0.0 ? t + (double)nb : t;
                    double angle = 360. - interp * 360. / (double)nb;
                    if (fabs(angle - 360.) < 1e-7) angle = 0;
                    if (slot < raw_cap) {
                        RawKeypoint k;
                        k.x = L.x; k.y = L.y; k.size = L.size; k.angle = (float)angle; k.response = L.response;
                        k.octave_packed = L.octave_packed;
                        k.img = img;
                        k.pad = 0;
                        k.order = (L.order << 6) | (uint64_t)b;
                        raw[slot] = k;
                    }
                }
                emitted += __popc(m);
            }
        }
        if (direct && lane == 0) direct_counts[li] = emitted;
        __syncwarp();
        li = li_next;
        L = Ln;
        li_next = li_after;
        Ln = La;
    }
}

#include "describe.cuh"

// ---------------------------------------------------------------------------
// ordering + de-duplication (sift_impl.py:299-327) + conversion (:333-343)
// ---------------------------------------------------------------------------
struct KpLess {
    const RawKeypoint *raw;
    int scan_order;
    __device__ __forceinline__ bool operator()(uint32_t ia, uint32_t ib) const
    {
        const RawKeypoint &a = raw[ia], &b = raw[ib];
        if (a.img != b.img) return a.img < b.img;
        if (!scan_order) {
            if (a.x != b.x) return a.x < b.x;
            if (a.y != b.y) return a.y < b.y;
            if (a.size != b.size) return a.size > b.size;
            if (a.angle != b.angle) return a.angle < b.angle;
            if (a.response != b.response) return a.response > b.response;
        }
        return a.order < b.order;  // reference: stable sort of the scan-order list
    }
};

__global__ void __launch_bounds__(256) iota_kernel(uint32_t *idx, int n)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) idx[i] = i;
}

__global__ void __launch_bounds__(256) zero_out_counts_kernel(int32_t *counters, int n_img)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i == 0) counters[CNT_OUT] = 0;
    if (i < n_img) counters[CNT_HDR + i * CNT_PER_IMG + 3] = 0;
}

__global__ void __launch_bounds__(256)
flag_kernel(const RawKeypoint *__restrict__ raw, const uint32_t *__restrict__ idx, int n, int dedupe,
            uint32_t *__restrict__ keep)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    uint32_t k = 1;
    if (dedupe && i > 0) {
        const RawKeypoint &a = raw[idx[i - 1]], &b = raw[idx[i]];
        if (a.img == b.img && a.x == b.x && a.y == b.y && a.size == b.size && a.angle == b.angle) k = 0;
    }
    keep[i] = k;
}

__global__ void __launch_bounds__(256)
gather_kernel(const RawKeypoint *__restrict__ raw, const uint8_t *__restrict__ raw_desc,
              const uint32_t *__restrict__ idx, const uint32_t *__restrict__ keep, const uint32_t *__restrict__ pos,
              int n, int convert, b200sift_keypoint *__restrict__ kps, uint8_t *__restrict__ desc,
              int32_t *__restrict__ counters, const int *__restrict__ img_base)
{
    // 8 threads per keypoint: each moves 16 B of the descriptor.  pos[] is either the global output
    // slot (scan over the whole list) or, with img_base, the slot inside the keypoint's image
    // (written by sort_image_kernel; the per-image counters are then already final).
    const int t = blockIdx.x * 256 + threadIdx.x;
    const int i = t >> 3, part = t & 7;
    if (i >= n || !keep[i]) return;
    const uint32_t src = idx[i];
    const uint32_t dst = img_base ? (uint32_t)img_base[raw[src].img] + pos[i] : pos[i];
    if (raw_desc)
        reinterpret_cast<uint4 *>(desc + (size_t)dst * 128)[part] =
            reinterpret_cast<const uint4 *>(raw_desc + (size_t)src * 128)[part];
    if (part == 0) {
        const RawKeypoint &r = raw[src];
        b200sift_keypoint k;
        if (convert) {
            k.x = r.x * 0.5f; k.y = r.y * 0.5f; k.size = r.size * 0.5f;
            k.octave = (r.octave_packed & ~255) | ((r.octave_packed - 1) & 255);
        } else {
            k.x = r.x; k.y = r.y; k.size = r.size; k.octave = r.octave_packed;
        }
        k.angle = r.angle;
        k.response = r.response;
        kps[dst] = k;
        if (!img_base) {
            atomicAdd(&counters[CNT_HDR + r.img * CNT_PER_IMG + 3], 1);
            atomicAdd(&counters[CNT_OUT], 1);
        }
    }
}

// ---------------------------------------------------------------------------
// Per-image ordering in shared memory (the common case: <= 4096 oriented keypoints per image).
// bucket_kernel groups the raw keypoint indices by image; sort_image_kernel (one CTA per image)
// loads the comparator keys of its segment into shared memory and runs a bitonic network on a
// permutation of slots.  The comparator is the total order of KpLess (compare_keypoints,
// sift_impl.py:299-311, then the scan-order key), so the result does not depend on the arbitrary
// order in which the atomics filled the segment.  Larger images take the CUB merge sort below.
// ---------------------------------------------------------------------------
constexpr int kSortMaxPerImage = 4096;
constexpr int kSortBytesPerSlot = 8 + 5 * 4 + 4 + 4;   // order key, 5 comparator floats, raw index, permutation
constexpr uint32_t kSortSentinel = 0xFFFFFFFFu;

__global__ void __launch_bounds__(256)
bucket_kernel(const RawKeypoint *__restrict__ raw, int n, const int *__restrict__ seg_off, int *__restrict__ cursor,
              uint32_t *__restrict__ out_idx)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int img = raw[i].img;
    const int slot = atomicAdd(&cursor[img], 1);
    out_idx[seg_off[img] + slot] = i;
}

struct SortKeys {
    float *x, *y, *size, *angle, *resp;
    unsigned long long *order;
    __device__ __forceinline__ bool less(uint32_t a, uint32_t b, int scan_order) const
    {
        if (a == kSortSentinel) return false;
        if (b == kSortSentinel) return true;
        if (!scan_order) {
            if (x[a] != x[b]) return x[a] < x[b];
            if (y[a] != y[b]) return y[a] < y[b];
            if (size[a] != size[b]) return size[a] > size[b];
            if (angle[a] != angle[b]) return angle[a] < angle[b];
            if (resp[a] != resp[b]) return resp[a] > resp[b];
        }
        return order[a] < order[b];
    }
};

__global__ void __launch_bounds__(1024)
sort_image_kernel(const RawKeypoint *__restrict__ raw, const int *__restrict__ seg_off,
                  const uint32_t *__restrict__ idx_in, uint32_t *__restrict__ idx_out, int scan_order, int P_max,
                  int dedupe, uint32_t *__restrict__ keep, uint32_t *__restrict__ pos_out,
                  int *__restrict__ img_kept)
{
    __shared__ int wsum[32];
    extern __shared__ __align__(16) unsigned char ssm[];
    SortKeys K;
    K.order = reinterpret_cast<unsigned long long *>(ssm);
    K.x = reinterpret_cast<float *>(K.order + P_max);
    K.y = K.x + P_max; K.size = K.y + P_max; K.angle = K.size + P_max; K.resp = K.angle + P_max;
    uint32_t *rid = reinterpret_cast<uint32_t *>(K.resp + P_max);
    uint32_t *perm = rid + P_max;
    const int base = seg_off[blockIdx.x], n = seg_off[blockIdx.x + 1] - base;
    if (n <= 0) {
        if (threadIdx.x == 0) img_kept[blockIdx.x] = 0;
        return;
    }
    int P = 2;
    while (P < n) P <<= 1;
    for (int s = threadIdx.x; s < P; s += 1024) {
        if (s < n) {
            const uint32_t r = idx_in[base + s];
            const RawKeypoint k = raw[r];
            K.x[s] = k.x; K.y[s] = k.y; K.size[s] = k.size; K.angle[s] = k.angle; K.resp[s] = k.response;
            K.order[s] = k.order;
            rid[s] = r;
            perm[s] = s;
        } else {
            perm[s] = kSortSentinel;
        }
    }
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += 1024) {
                const int i1 = 2 * t - (t & (j - 1));
                const int i2 = i1 + j;
                const uint32_t a = perm[i1], b = perm[i2];
                const bool up = (i1 & k) == 0;
                // ascending block: want perm[i1] <= perm[i2]
                const bool swap = up ? K.less(b, a, scan_order) : K.less(a, b, scan_order);
                if (swap) { perm[i1] = b; perm[i2] = a; }
            }
            __syncthreads();
        }
    }
    for (int s = threadIdx.x; s < n; s += 1024) idx_out[base + s] = rid[perm[s]];
    // remove_duplicate_keypoints (:313-327) on the sorted list: keep flag and the slot inside this
    // image's output (block-wide exclusive scan, 1024 entries per round).  idx_in aliases pos_out:
    // this CTA's segment of it was consumed above.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int running = 0;
    for (int s0 = 0; s0 < n; s0 += 1024) {
        const int s = s0 + threadIdx.x;
        int k = 0;
        if (s < n) {
            k = 1;
            if (dedupe && s > 0) {
                const uint32_t a = perm[s - 1], b = perm[s];
                if (K.x[a] == K.x[b] && K.y[a] == K.y[b] && K.size[a] == K.size[b] && K.angle[a] == K.angle[b]) k = 0;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int wv = 0; wv < 32; ++wv) {
            const int c = wsum[wv];
            if (wv < warp) woff += c;
            tot += c;
        }
        if (s < n) {
            keep[base + s] = (uint32_t)k;
            pos_out[base + s] = (uint32_t)(running + woff + __popc(m & ((1u << lane) - 1u)));
        }
        running += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) img_kept[blockIdx.x] = running;
}

// Exclusive scan of the per-image kept counts -> first output slot of every image; also the final
// per-image and total output counters (one CTA; n_img is small next to the keypoint lists).
__global__ void __launch_bounds__(1024)
image_offsets_kernel(const int *__restrict__ img_kept, int n_img, int *__restrict__ img_base,
                     int32_t *__restrict__ counters)
{
    __shared__ int wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int running = 0;
    for (int i0 = 0; i0 < n_img; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const int c = i < n_img ? img_kept[i] : 0;
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int woff = 0, tot = 0;
        for (int wv = 0; wv < 32; ++wv) {
            const int w = wsum[wv];
            if (wv < warp) woff += w;
            tot += w;
        }
        if (i < n_img) {
            img_base[i] = running + woff + incl - c;
            counters[CNT_HDR + i * CNT_PER_IMG + 3] = c;
        }
        running += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) counters[CNT_OUT] = running;
}

static int g_desc_occ = 1;  // resident describe CTAs per SM (kernel + sm_100 property; set under the init lock)
static int g_ori_occ = 5;   // resident orient CTAs per SM, same

// Function attributes are per DEVICE: called once for every device a context is created on
// (b200sift_create, under the init lock), never from a launch path.
int detect_init_device()
{
    const size_t smem = (size_t)kDescWarps * kDescSmemPerWarp;
    B200_CUDA(cudaFuncSetAttribute(describe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_CUDA(cudaFuncSetAttribute(describe_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int occ = 0;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, describe_kernel, kDescWarps * 32, smem));
    g_desc_occ = occ < 1 ? 1 : occ;
    B200_CUDA(cudaFuncSetAttribute(orient_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, orient_kernel, kOriWarps * 32, 0));
    g_ori_occ = occ < 1 ? 1 : occ;
    B200_CUDA(cudaFuncSetAttribute(sort_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kSortMaxPerImage * kSortBytesPerSlot));
    B200_CUDA(cudaFuncSetAttribute(describe_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kDescGenericMaxBins * 32 * (int)sizeof(float)));
    return 0;
}

static int ensure_sparse(b200sift_ctx *c, int cand_cap, int loc_cap, int raw_cap)
{
    size_t cap;
    cap = c->cand_cap; B200_CHECK(ensure(&c->d_cand, &cap, (size_t)cand_cap)); c->cand_cap = (int)cap;
    cap = c->loc_cap;  B200_CHECK(ensure(&c->d_loc, &cap, (size_t)loc_cap));   c->loc_cap = (int)cap;
    if (raw_cap > c->raw_cap || !c->d_raw) {
        cap = c->raw_cap;
        B200_CHECK(ensure(&c->d_raw, &cap, (size_t)raw_cap));
        const int rc = (int)cap;
        size_t cap2 = 0;
        if (c->d_raw_desc) { cudaFree(c->d_raw_desc); c->d_raw_desc = nullptr; }
        B200_CHECK(ensure(&c->d_raw_desc, &cap2, (size_t)rc * 128));
        cap2 = 0; if (c->d_class_idx) { cudaFree(c->d_class_idx); c->d_class_idx = nullptr; }
        B200_CHECK(ensure(&c->d_class_idx, &cap2, (size_t)rc * kDescClasses));
        cap2 = 0; if (c->d_sort_idx) { cudaFree(c->d_sort_idx); c->d_sort_idx = nullptr; }
        B200_CHECK(ensure(&c->d_sort_idx, &cap2, (size_t)rc));
        cap2 = 0; if (c->d_keep) { cudaFree(c->d_keep); c->d_keep = nullptr; }
        B200_CHECK(ensure(&c->d_keep, &cap2, (size_t)rc));
        cap2 = 0; if (c->d_pos) { cudaFree(c->d_pos); c->d_pos = nullptr; }
        B200_CHECK(ensure(&c->d_pos, &cap2, (size_t)rc));
        cap2 = 0; if (c->d_kps) { cudaFree(c->d_kps); c->d_kps = nullptr; }
        B200_CHECK(ensure(&c->d_kps, &cap2, (size_t)rc));
        cap2 = 0; if (c->d_desc) { cudaFree(c->d_desc); c->d_desc = nullptr; }
        B200_CHECK(ensure(&c->d_desc, &cap2, (size_t)rc * 128));
        c->raw_cap = rc;
        c->out_cap = rc;
    }
    return 0;
}

static int ensure_counters(b200sift_ctx *c, int n_img)
{
    const int need = CNT_HDR + n_img * CNT_PER_IMG;
    if (need > c->counters_len) {
        if (c->d_counters) cudaFree(c->d_counters);
        if (c->h_counters) cudaFreeHost(c->h_counters);
        B200_CUDA(cudaMalloc((void **)&c->d_counters, sizeof(int32_t) * need));
        B200_CUDA(cudaMallocHost((void **)&c->h_counters, sizeof(int32_t) * need));
        c->counters_len = need;
    }
    return 0;
}

// extrema -> refine -> orientation on the context's pyramid.  Leaves the raw
// (unsorted) oriented keypoints in c->d_raw and the counters on the host.
int run_detect(b200sift_ctx *c, const b200sift_params &p, int use_dog)
{
    // use_dog: extrema + quadratic fit read c->dog_pyr (a caller-supplied dog_images) instead of
    // forming the differences of the Gaussian layers; orientations always read c->pyr.
    const Pyramid &py = c->pyr;
    if (use_dog) {
        const Pyramid &dg = c->dog_pyr;
        B200_ARG(dg.n_oct == py.n_oct && dg.n_layers == py.n_layers - 1 && dg.n_img == py.n_img);
        for (int o = 0; o < py.n_oct; ++o) B200_ARG(dg.h[o] == py.h[o] && dg.w[o] == py.w[o]);
    }
    const PyrView vd = use_dog ? make_view(c->dog_pyr) : make_view(c);
    B200_ARG(p.num_intervals >= 1 && p.num_intervals + 3 == py.n_layers);
    B200_ARG(p.image_border_width >= 1);
    B200_ARG(p.ori_bins >= 4 && p.ori_bins <= kOriMaxBins);
    B200_ARG(p.max_iter >= 1);
    B200_ARG(py.h[0] < 32768 && py.w[0] < 32768 && py.n_img < 65536);
    const PyrView v = make_view(c);
    const DetectParams dp = make_detect_params(p);
    B200_CHECK(ensure_counters(c, py.n_img));
    const size_t px = (size_t)py.n_img * py.h[0] * py.w[0];
    int cand_cap = (int)(px / 48 + 16384), loc_cap = (int)(px / 96 + 8192), raw_cap = (int)(px / 64 + 8192);
    if (c->cand_cap > cand_cap) cand_cap = c->cand_cap;
    if (c->loc_cap > loc_cap) loc_cap = c->loc_cap;
    if (c->raw_cap > raw_cap) raw_cap = c->raw_cap;
    const int n_cnt = CNT_HDR + py.n_img * CNT_PER_IMG;
    for (int attempt = 0; attempt < 4; ++attempt) {
        B200_CHECK(ensure_sparse(c, cand_cap, loc_cap, raw_cap));
        // First attempt after build_octaves: the scan of octave o runs on the side stream as soon as
        // that octave's layers are complete (event), i.e. next to the blurs of the following octaves.
        const bool overlap = (attempt == 0) && c->oct_events_valid;
        c->oct_events_valid = false;
        cudaStream_t es = overlap ? c->side_stream : c->stream;
        B200_CUDA(cudaMemsetAsync(c->d_counters, 0, sizeof(int32_t) * n_cnt, es));
        const size_t ex_smem = (size_t)(py.n_layers - 1) * (kExTW + 2) * (kExTH + 2) * sizeof(float);
        // octaves at or past `o_merge` (the ones the pyramid tail kernel produced together) share one launch
        const int o_merge = (overlap && p.num_intervals == 3 && c->pyr_o_tail > 0) ? c->pyr_o_tail : py.n_oct;
        for (int o = 0; o < py.n_oct; ++o) {
            const int sh = py.h[o] - 2 * p.image_border_width, sw = py.w[o] - 2 * p.image_border_width;
            if (sh <= 0 || sw <= 0) continue;
            if (overlap) B200_CUDA(cudaStreamWaitEvent(es, c->ev_oct[o], 0));
            if (!use_dog && p.num_intervals == 3 && (sw >= 24 || o >= o_merge)) {
                ExGroup g;
                g.o_first = o;
                g.n_oct = 0;
                g.blk_off[0] = 0;
                const int o_end = o >= o_merge ? py.n_oct : o + 1;
                int oo = o;
                for (; oo < o_end; ++oo) {
                    const int sh2 = py.h[oo] - 2 * p.image_border_width, sw2 = py.w[oo] - 2 * p.image_border_width;
                    if (sh2 <= 0 || sw2 <= 0) break;  // every later octave is smaller still
                    const int j = g.n_oct++;
                    g.seg_rows[j] = py.h[oo] > 256 ? 32 : 8;
                    g.n_cg[j] = (sw2 + 29) / 30;
                    g.n_rs[j] = (sh2 + g.seg_rows[j] - 1) / g.seg_rows[j];
                    g.blk_off[j + 1] = g.blk_off[j] + (g.n_cg[j] * g.n_rs[j] + 7) / 8;
                }
                dim3 grid(g.blk_off[g.n_oct], py.n_img);
                extrema_rows_kernel<3><<<grid, 256, 0, es>>>(v, g, p.image_border_width, dp.dog_thresh, c->d_cand,
                                                            c->cand_cap, c->d_counters);
                c->launches++;
                tl_mark(es, "side  extrema oct %d..%d", o, o + g.n_oct - 1);
                if (o >= o_merge) break;  // the group covered every remaining octave
            } else {
                dim3 grid((sw + kExTW - 1) / kExTW, (sh + kExTH - 1) / kExTH, py.n_img);
                if (use_dog)
                    extrema_kernel<true><<<grid, 256, ex_smem, es>>>(vd, o, p.image_border_width, p.num_intervals,
                                                                     dp.dog_thresh, c->d_cand, c->cand_cap,
                                                                     c->d_counters);
                else
                    extrema_kernel<false><<<grid, 256, ex_smem, es>>>(v, o, p.image_border_width, p.num_intervals,
                                                                      dp.dog_thresh, c->d_cand, c->cand_cap,
                                                                      c->d_counters);
                c->launches++;
            }
        }
        if (overlap) {
            B200_CUDA(cudaEventRecord(c->ev_side, es));
            B200_CUDA(cudaStreamWaitEvent(c->stream, c->ev_side, 0));
        }
        B200_CUDA(cudaGetLastError());
        {
            int blocks = (c->cand_cap + 127) / 128;
            if (blocks > c->sm_count * 8) blocks = c->sm_count * 8;
            if (use_dog)
                refine_kernel<true><<<blocks, 128, 0, c->stream>>>(vd, dp, c->d_cand, c->cand_cap, c->d_loc,
                                                                   c->loc_cap, c->d_counters, -1, 0);
            else
                refine_kernel<false><<<blocks, 128, 0, c->stream>>>(v, dp, c->d_cand, c->cand_cap, c->d_loc,
                                                                    c->loc_cap, c->d_counters, -1, 0);
            c->launches++;
            tl_mark(c->stream, "main  refine");
            orient_kernel<<<c->sm_count * g_ori_occ, kOriWarps * 32, 0, c->stream>>>(v, dp, c->d_loc, c->loc_cap, c->d_raw,
                                                                             c->raw_cap, c->d_counters, -1, nullptr,
                                                                             c->d_class_idx);
            c->launches++;
            tl_mark(c->stream, "main  orient");
        }
        B200_CUDA(cudaGetLastError());
        B200_CUDA(cudaMemcpyAsync(c->h_counters, c->d_counters, sizeof(int32_t) * n_cnt, cudaMemcpyDeviceToHost,
                                  c->stream));
        B200_CUDA(b200::ctx_sync(c));
        tl_mark(c->stream, "main  counters on host");
        const int nc = c->h_counters[CNT_CAND], nl = c->h_counters[CNT_LOC], nr = c->h_counters[CNT_RAW];
        if (nc <= c->cand_cap && nl <= c->loc_cap && nr <= c->raw_cap) return 0;
        // a fixed-capacity list overflowed: grow to twice what was needed and redo the stage
        if (nc > c->cand_cap) cand_cap = 2 * nc;
        if (nl > c->loc_cap) loc_cap = 2 * nl;
        if (nr > c->raw_cap) raw_cap = 2 * nr + 1024;
        if (nc > c->cand_cap && loc_cap < nc) loc_cap = nc;
        if (raw_cap < 2 * loc_cap) raw_cap = 2 * loc_cap;
    }
    set_error("keypoint buffers still overflow after 4 attempts");
    return B200SIFT_ECAPACITY;
}

int run_describe(b200sift_ctx *c, const b200sift_params &p, const RawKeypoint *d_raw, int n, int converted,
                 uint8_t *d_out, int use_classes)
{
    if (n <= 0) return 0;
    B200_ARG(p.window_width >= 1 && p.desc_bins >= 1 &&
             p.window_width * p.window_width * p.desc_bins <= kDescGenericMaxBins);
    const PyrView v = make_view(c);
    const DetectParams dp = make_detect_params(p);
    if (p.window_width != 4 || p.desc_bins != 8) {   // d_out holds n * window_width^2 * desc_bins bytes
        const int dlen = p.window_width * p.window_width * p.desc_bins;
        int blocks = n < c->sm_count * 4 ? n : c->sm_count * 4;
        describe_generic_kernel<<<blocks, 32, (size_t)dlen * 32 * sizeof(float), c->stream>>>(
            v, dp, p.window_width, p.desc_bins, d_raw, n, converted, d_out);
        c->launches++;
        B200_CUDA(cudaGetLastError());
        return 0;
    }
    const size_t smem = (size_t)kDescWarps * kDescSmemPerWarp;
    int blocks = (n + kDescWarps - 1) / kDescWarps;
    if (blocks > c->sm_count * g_desc_occ) blocks = c->sm_count * g_desc_occ;
    B200_CHECK(ensure_counters(c, c->pyr.n_img > 0 ? c->pyr.n_img : 1));
    B200_CUDA(cudaMemsetAsync(c->d_counters + CNT_WORK_DESC, 0, sizeof(int32_t), c->stream));
    describe_kernel<<<blocks, kDescWarps * 32, smem, c->stream>>>(v, dp, d_raw, n, converted, d_out, c->d_counters,
                                                                  use_classes ? c->d_class_idx : nullptr, c->raw_cap);
    c->launches++;
    B200_CUDA(cudaGetLastError());
    return 0;
}

// Sort the n raw keypoints (by image, then compare_keypoints or scan order),
// optionally drop duplicates, convert and compact into c->d_kps / c->d_desc.
// Ordering of the n raw keypoints (by image, then compare_keypoints or scan order) into
// c->d_sort_idx, on the side stream; run_gather() joins it.
int run_sort_async(b200sift_ctx *c, int n_raw, int n_img, int scan_order, int dedupe)
{
    c->sort_fast = false;
    c->sort_dedupe = dedupe;
    B200_CHECK(ensure_counters(c, n_img));
    c->img_off.assign(n_img + 1, 0);
    if (n_raw <= 0) return 0;
    cudaStream_t ss = c->side_stream;
    B200_CUDA(cudaEventRecord(c->ev_main, c->stream));       // everything queued so far (uploads, d_raw)
    B200_CUDA(cudaStreamWaitEvent(ss, c->ev_main, 0));
    const int blocks = (n_raw + 255) / 256;
    // per-image raw counts (known on the host since the read-back that ended run_detect)
    std::vector<int> &seg = c->h_seg;
    seg.assign(n_img + 1, 0);
    int max_per = 0;
    bool have_counts = false;
    if (n_img == 1) {
        seg[1] = n_raw;
        max_per = n_raw;
        have_counts = true;
    } else if ((int)c->stat_raw.size() == n_img) {
        for (int i = 0; i < n_img; ++i) {
            seg[i + 1] = seg[i] + c->stat_raw[i];
            if (c->stat_raw[i] > max_per) max_per = c->stat_raw[i];
        }
        have_counts = (seg[n_img] == n_raw);
    }
    // temp storage for the scan of run_gather (and the merge-sort fall-back)
    size_t tmp = 0, tmp2 = 0;
    KpLess less{c->d_raw, scan_order};
    const bool fast = have_counts && max_per <= kSortMaxPerImage;
    if (!fast) B200_CUDA(cub::DeviceMergeSort::StableSortKeys(nullptr, tmp, c->d_sort_idx, n_raw, less, ss));
    B200_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, c->d_keep, c->d_pos, n_raw, ss));
    if (tmp2 > tmp) tmp = tmp2;
    if (tmp > c->cub_tmp_cap) {
        B200_CUDA(b200::ctx_sync(c));
        if (c->d_cub_tmp) cudaFree(c->d_cub_tmp);
        c->d_cub_tmp = nullptr;
        B200_CUDA(cudaMalloc(&c->d_cub_tmp, tmp + 1024));
        c->cub_tmp_cap = tmp + 1024;
    }
    if (fast) {
        size_t cap = c->seg_cap;
        B200_CHECK(ensure(&c->d_seg, &cap, (size_t)4 * (n_img + 1)));
        c->seg_cap = cap;
        int *d_off = c->d_seg, *d_cur = c->d_seg + (n_img + 1);
        int *d_kept = c->d_seg + 2 * (n_img + 1), *d_base = c->d_seg + 3 * (n_img + 1);
        B200_CUDA(cudaMemcpyAsync(d_off, seg.data(), sizeof(int) * (n_img + 1), cudaMemcpyHostToDevice, ss));
        B200_CUDA(cudaMemsetAsync(d_cur, 0, sizeof(int) * (n_img + 1), ss));
        int P = 2;
        while (P < max_per) P <<= 1;
        const size_t smem = (size_t)P * kSortBytesPerSlot;
        // d_pos doubles as the bucketed (unsorted) index list; run_gather's scan rewrites it afterwards
        bucket_kernel<<<blocks, 256, 0, ss>>>(c->d_raw, n_raw, d_off, d_cur, c->d_pos);
        sort_image_kernel<<<n_img, 1024, smem, ss>>>(c->d_raw, d_off, c->d_pos, c->d_sort_idx, scan_order, P, dedupe,
                                                     c->d_keep, c->d_pos, d_kept);
        image_offsets_kernel<<<1, 1024, 0, ss>>>(d_kept, n_img, d_base, c->d_counters);
        c->launches += 3;
        c->sort_fast = true;
        tl_mark(ss, "side  bucket + sort + offsets");
    } else {
        size_t t1 = c->cub_tmp_cap;
        iota_kernel<<<blocks, 256, 0, ss>>>(c->d_sort_idx, n_raw);
        B200_CUDA(cub::DeviceMergeSort::StableSortKeys(c->d_cub_tmp, t1, c->d_sort_idx, n_raw, less, ss));
        c->launches += 3;
    }
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaEventRecord(c->ev_side, ss));
    return 0;
}

// Join the sort, optionally drop duplicates, convert and compact into c->d_kps / c->d_desc.
int run_gather(b200sift_ctx *c, int n_raw, int n_img, int dedupe, int convert, int with_desc)
{
    if (n_raw <= 0) return 0;
    B200_CUDA(cudaStreamWaitEvent(c->stream, c->ev_side, 0));
    const int blocks = (n_raw + 255) / 256;
    if (c->sort_fast) {
        // flags, in-image slots and per-image counts came out of the per-image sort (side stream)
        if (dedupe != c->sort_dedupe) {
            set_error("run_gather: dedupe flag differs from the one the sort was run with");
            return B200SIFT_ESTATE;
        }
        gather_kernel<<<(n_raw * 8 + 255) / 256, 256, 0, c->stream>>>(
            c->d_raw, with_desc ? c->d_raw_desc : nullptr, c->d_sort_idx, c->d_keep, c->d_pos, n_raw, convert, c->d_kps,
            c->d_desc, c->d_counters, c->d_seg + 3 * (n_img + 1));
        c->launches += 1;
    } else {
        zero_out_counts_kernel<<<(n_img + 255) / 256, 256, 0, c->stream>>>(c->d_counters, n_img);
        flag_kernel<<<blocks, 256, 0, c->stream>>>(c->d_raw, c->d_sort_idx, n_raw, dedupe, c->d_keep);
        size_t t1 = c->cub_tmp_cap;
        B200_CUDA(cub::DeviceScan::ExclusiveSum(c->d_cub_tmp, t1, c->d_keep, c->d_pos, n_raw, c->stream));
        gather_kernel<<<(n_raw * 8 + 255) / 256, 256, 0, c->stream>>>(
            c->d_raw, with_desc ? c->d_raw_desc : nullptr, c->d_sort_idx, c->d_keep, c->d_pos, n_raw, convert, c->d_kps,
            c->d_desc, c->d_counters, nullptr);
        c->launches += 4;
    }
    B200_CUDA(cudaGetLastError());
    const int n_cnt = CNT_HDR + n_img * CNT_PER_IMG;
    B200_CUDA(cudaMemcpyAsync(c->h_counters, c->d_counters, sizeof(int32_t) * n_cnt, cudaMemcpyDeviceToHost,
                              c->stream));
    B200_CUDA(b200::ctx_sync(c));
    for (int i = 0; i < n_img; ++i) c->img_off[i + 1] = c->img_off[i] + c->h_counters[CNT_HDR + i * CNT_PER_IMG + 3];
    return 0;
}

int run_sort_gather(b200sift_ctx *c, int n_raw, int n_img, int scan_order, int dedupe, int convert, int with_desc)
{
    B200_CHECK(run_sort_async(c, n_raw, n_img, scan_order, dedupe));
    return run_gather(c, n_raw, n_img, dedupe, convert, with_desc);
}

int ensure_sparse_for(b200sift_ctx *c, int n_img, int n_raw)
{
    B200_CHECK(ensure_counters(c, n_img));
    return ensure_sparse(c, c->cand_cap > 0 ? c->cand_cap : 1024, c->loc_cap > 0 ? c->loc_cap : 1024,
                         n_raw > c->raw_cap ? n_raw : c->raw_cap);
}

// ---------------------------------------------------------------------------
// "next" rows f1 / f2
// ---------------------------------------------------------------------------

// ransac() vote (image_stitching_sift.py:86-111): one thread per candidate shift counts the
// matches within dist_sq_thresh (float64, as the python floats of the reference); the first
// maximum wins.  The candidate moves are streamed through shared memory in tiles, so the number of
// matches is not limited by the shared-memory size.
constexpr int kVoteTile = 1024;

__global__ void __launch_bounds__(256)
ransac_vote_kernel(const double *__restrict__ m, int n, double thr, int32_t *__restrict__ votes)
{
    __shared__ double sh[2 * kVoteTile];  // candidate moves of the current tile
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int ii = min(i, n - 1);
    const double dxr = m[4 * (size_t)ii] - m[4 * (size_t)ii + 2], dyr = m[4 * (size_t)ii + 1] - m[4 * (size_t)ii + 3];
    int cnt = 0;
    for (int j0 = 0; j0 < n; j0 += kVoteTile) {
        const int nt = min(kVoteTile, n - j0);
        __syncthreads();
        for (int j = threadIdx.x; j < nt; j += 256) {
            sh[2 * j] = m[4 * (size_t)(j0 + j)] - m[4 * (size_t)(j0 + j) + 2];       // :94-96
            sh[2 * j + 1] = m[4 * (size_t)(j0 + j) + 1] - m[4 * (size_t)(j0 + j) + 3];
        }
        __syncthreads();
        for (int j = 0; j < nt; ++j) {
            const double dx = sh[2 * j] - dxr, dy = sh[2 * j + 1] - dyr;
            cnt += (dx * dx + dy * dy < thr) ? 1 : 0;
        }
    }
    if (i < n) votes[i] = cnt;
}

// first maximum: maximise (votes << 32) | (~index); also hands back the winning move
__global__ void __launch_bounds__(256)
ransac_pick_kernel(const int32_t *__restrict__ votes, const double *__restrict__ m, int n, double *__restrict__ out)
{
    __shared__ unsigned long long red[256];
    unsigned long long k = 0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const unsigned long long key = ((unsigned long long)(uint32_t)votes[i] << 32) | (uint32_t)(~(uint32_t)i);
        if (key > k) k = key;
    }
    red[threadIdx.x] = k;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s && red[threadIdx.x + s] > red[threadIdx.x]) red[threadIdx.x] = red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint32_t best = ~(uint32_t)(red[0] & 0xffffffffu);
        out[0] = (double)best;
        out[1] = m[4 * (size_t)best] - m[4 * (size_t)best + 2];
        out[2] = m[4 * (size_t)best + 1] - m[4 * (size_t)best + 3];
    }
}

int launch_ransac(b200sift_ctx *c, const double *d_matches, int n, double thr, double *move, int32_t *best)
{
    move[0] = move[1] = 0;
    *best = -1;
    if (n <= 0) return 0;
    size_t cap = c->misc_cap;
    B200_CHECK(ensure((uint8_t **)&c->d_misc, &cap, (size_t)n * sizeof(int32_t) + 64));
    c->misc_cap = cap;
    double *d_out = (double *)c->d_misc;          // 3 doubles, then the votes
    int32_t *votes = (int32_t *)(d_out + 4);
    ransac_vote_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_matches, n, thr, votes);
    ransac_pick_kernel<<<1, 256, 0, c->stream>>>(votes, d_matches, n, d_out);
    c->launches += 2;
    B200_CUDA(cudaGetLastError());
    double res[3];
    B200_CUDA(cudaMemcpyAsync(res, d_out, sizeof(res), cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(b200::ctx_sync(c));
    *best = (int32_t)res[0];
    move[0] = res[1];
    move[1] = res[2];
    return 0;
}

// ---------------------------------------------------------------------------
// Per-call forms of the stage API (sift_impl.py:169-211 and :246-293 as the reference exposes
// them): the same kernels on a caller's candidates / keypoints, results in input order.
// ---------------------------------------------------------------------------

// localize_extremum_via_quadratic_fit for n candidates (octave, layer, y, x) on c->pyr (Gaussian
// layers) or c->dog_pyr (use_dog).  single_octave: the pyramid holds only the candidates' octave.
// final_layer[i] = -1 where the reference returns None.
int run_localize_direct(b200sift_ctx *c, const b200sift_params &p, int use_dog, int single_octave,
                        const int32_t *h_cand, int n, b200sift_keypoint *h_kps, int32_t *h_final_layer)
{
    if (n <= 0) return 0;
    const Pyramid &py = use_dog ? c->dog_pyr : c->pyr;
    const int n_dog = use_dog ? py.n_layers : py.n_layers - 1;
    B200_ARG(p.num_intervals >= 1 && n_dog == p.num_intervals + 2 && p.max_iter >= 1);
    std::vector<Candidate> hc(n);
    for (int i = 0; i < n; ++i) {
        const int o = h_cand[4 * i], l = h_cand[4 * i + 1], y = h_cand[4 * i + 2], x = h_cand[4 * i + 3];
        const int ov = single_octave ? 0 : o;
        B200_ARG(o >= 0 && o < 31 && ov < py.n_oct && l >= 1 && l <= p.num_intervals);
        B200_ARG(y >= 1 && y < py.h[ov] - 1 && x >= 1 && x < py.w[ov] - 1);   // the 3x3x3 cube must exist
        hc[i].img_o_l = ((uint32_t)o << 8) | (uint32_t)l;
        hc[i].yx = ((uint32_t)y << 16) | (uint32_t)x;
    }
    B200_CHECK(ensure_counters(c, 1));
    B200_CHECK(ensure_sparse(c, n > c->cand_cap ? n : c->cand_cap, n > c->loc_cap ? n : c->loc_cap,
                             c->raw_cap > 0 ? c->raw_cap : 1024));
    B200_CUDA(cudaMemcpyAsync(c->d_cand, hc.data(), sizeof(Candidate) * n, cudaMemcpyHostToDevice, c->stream));
    const PyrView v = make_view(py);
    const DetectParams dp = make_detect_params(p);
    const int blocks = (n + 127) / 128;
    if (use_dog)
        refine_kernel<true><<<blocks, 128, 0, c->stream>>>(v, dp, c->d_cand, n, c->d_loc, n, c->d_counters, n,
                                                           single_octave);
    else
        refine_kernel<false><<<blocks, 128, 0, c->stream>>>(v, dp, c->d_cand, n, c->d_loc, n, c->d_counters, n,
                                                            single_octave);
    c->launches++;
    B200_CUDA(cudaGetLastError());
    std::vector<Localized> hl(n);
    B200_CUDA(cudaMemcpyAsync(hl.data(), c->d_loc, sizeof(Localized) * n, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(b200::ctx_sync(c));   // also covers the copy out of `hc`
    for (int i = 0; i < n; ++i) {
        const bool ok = hl[i].img_o_l != 0xFFFFFFFFu;
        b200sift_keypoint k;
        k.x = ok ? hl[i].x : 0.f; k.y = ok ? hl[i].y : 0.f; k.size = ok ? hl[i].size : 0.f;
        k.angle = -1.f;                                  // cv2.KeyPoint() default (:206)
        k.response = ok ? hl[i].response : 0.f;
        k.octave = ok ? hl[i].octave_packed : 0;
        h_kps[i] = k;
        h_final_layer[i] = ok ? (int32_t)(hl[i].img_o_l & 255) : -1;
    }
    return 0;
}

// compute_keypoints_with_orientations for n keypoints on ONE Gaussian image (c->pyr holds it as
// octave 0 / layer 0); `octave` is the reference's octave argument.  Keypoint i yields counts[i]
// keypoints at out[i * ori_bins ...], ascending in histogram bin like the reference's list.
int run_orient_direct(b200sift_ctx *c, const b200sift_params &p, const b200sift_keypoint *h_kps, int n, int octave,
                      b200sift_keypoint *h_out, int32_t *h_counts)
{
    if (n <= 0) return 0;
    B200_ARG(p.ori_bins >= 4 && p.ori_bins <= kOriMaxBins && octave >= 0 && octave < 30);
    B200_ARG(c->pyr.n_oct >= 1 && c->pyr.n_layers >= 1 && c->pyr.n_img == 1);
    const int nb = p.ori_bins;
    B200_CHECK(ensure_counters(c, 1));
    B200_CHECK(ensure_sparse(c, c->cand_cap > 0 ? c->cand_cap : 1024, n > c->loc_cap ? n : c->loc_cap,
                             n * nb > c->raw_cap ? n * nb : c->raw_cap));
    std::vector<Localized> hl(n);
    for (int i = 0; i < n; ++i) {
        hl[i].x = h_kps[i].x; hl[i].y = h_kps[i].y; hl[i].size = h_kps[i].size; hl[i].response = h_kps[i].response;
        hl[i].octave_packed = h_kps[i].octave;
        hl[i].img_o_l = (uint32_t)octave << 8;
        hl[i].order = (uint64_t)i;
    }
    B200_CUDA(cudaMemcpyAsync(c->d_loc, hl.data(), sizeof(Localized) * n, cudaMemcpyHostToDevice, c->stream));
    B200_CUDA(cudaMemsetAsync(c->d_counters, 0, sizeof(int32_t) * CNT_HDR, c->stream));
    int32_t *d_counts = reinterpret_cast<int32_t *>(c->d_cand);   // scratch: n <= cand_cap * 2
    B200_ARG((size_t)n * sizeof(int32_t) <= (size_t)c->cand_cap * sizeof(Candidate));
    const PyrView v = make_view(c);
    const DetectParams dp = make_detect_params(p);
    int blocks = (n + kOriWarps - 1) / kOriWarps;
    if (blocks > c->sm_count * 5) blocks = c->sm_count * 5;
    orient_kernel<<<blocks, kOriWarps * 32, 0, c->stream>>>(v, dp, c->d_loc, n, c->d_raw, n * nb, c->d_counters, n,
                                                            d_counts, nullptr);
    c->launches++;
    B200_CUDA(cudaGetLastError());
    std::vector<RawKeypoint> hr((size_t)n * nb);
    B200_CUDA(cudaMemcpyAsync(h_counts, d_counts, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(cudaMemcpyAsync(hr.data(), c->d_raw, sizeof(RawKeypoint) * hr.size(), cudaMemcpyDeviceToHost,
                              c->stream));
    B200_CUDA(b200::ctx_sync(c));
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < h_counts[i]; ++k) {
            const RawKeypoint &r = hr[(size_t)i * nb + k];
            b200sift_keypoint q;
            q.x = r.x; q.y = r.y; q.size = r.size; q.angle = r.angle; q.response = r.response;
            q.octave = r.octave_packed;
            h_out[(size_t)i * nb + k] = q;
        }
    return 0;
}

// cylindrical_projection (image_stitching_sift.py:117-136).  The reference is
// a forward map where later source pixels (row-major) overwrite earlier ones;
// here every source pixel publishes its row-major index with atomicMax into a
// per-destination winner map, then each destination copies from its winner.
__global__ void __launch_bounds__(256) cyl_winner_kernel(int h, int w, double f, int32_t *__restrict__ winner)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= h * w) return;
    const int yy = i / w, xx = i - yy * w;
    const int xd = xx - w / 2, yd = yy - h / 2;
    const long long xm = (long long)rint(f * atan((double)xd / f)) + w / 2;
    const double denom = sqrt((double)(xd * xd) + f * f);
    const long long ym = (long long)rint(f * ((double)yd / denom)) + h / 2;
    if (xm >= 0 && xm < w && ym >= 0 && ym < h) atomicMax(&winner[ym * w + xm], i);
}

__global__ void __launch_bounds__(256)
cyl_copy_kernel(const uint8_t *__restrict__ src, int n, int ch, const int32_t *__restrict__ winner,
                uint8_t *__restrict__ dst)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int s = winner[i];
    for (int k = 0; k < ch; ++k) dst[(size_t)i * ch + k] = s >= 0 ? src[(size_t)s * ch + k] : 0;
}

int launch_cyl(b200sift_ctx *c, const uint8_t *d_src, int h, int w, int ch, double f, uint8_t *d_dst)
{
    const int n = h * w;
    size_t cap = c->misc_cap;
    B200_CHECK(ensure((uint8_t **)&c->d_misc, &cap, (size_t)n * sizeof(int32_t)));
    c->misc_cap = cap;
    int32_t *winner = (int32_t *)c->d_misc;
    B200_CUDA(cudaMemsetAsync(winner, 0xff, (size_t)n * sizeof(int32_t), c->stream));
    cyl_winner_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(h, w, f, winner);
    cyl_copy_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_src, n, ch, winner, d_dst);
    c->launches += 2;
    B200_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200
