/*
 * oracle/sift_oracle.c -- CPU restatement of the reference SIFT hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package
 * (vfx_image_stitching_b200/) may include, link, import or execute this file;
 * it exists so that tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs have something to check the CUDA path
 * against and to time on the host cores.
 *
 * What it restates (all citations are into /root/reference):
 *   sift_impl.py:15-526            the SIFT detect+describe pipeline
 *   image_stitching_sift.py:63-79  the brute-force nearest-neighbour matcher
 *   image_stitching_sift.py:86-111 ransac() translation vote          (row f1)
 *   image_stitching_sift.py:117-136 cylindrical_projection()          (row f2)
 * plus the third-party arithmetic those lines delegate to and that is NOT
 * under /root/reference (opencv-python 4.13.0, numpy 2.3.5 -- the reference
 * pins neither, README.md:10-12): cv2.cvtColor(BGR2GRAY) on uint8,
 * cv2.resize(INTER_LINEAR, fx=fy=2) on float32, cv2.GaussianBlur on float32
 * (BORDER_REFLECT_101), cv2.resize(INTER_NEAREST) by 1/2, numpy float32 /
 * float64 promotion (NumPy 2 / NEP 50), np.linalg.lstsq (LAPACK gelsd in
 * float64), np.linalg.det, np.round (half to even), np.add.at.
 *
 * Parity status: PINNED against outputs of the unmodified reference run in
 * the build container (tests/golden/make_golden.py -> tests/golden/*.npz,
 * checked by tests/test_oracle_golden.py).  Integer paths (grey conversion,
 * upsample, extrema scan, matcher, ransac, projection) are bit-exact.  The
 * float paths are exact restatements of the operation order and dtype of
 * every expression; they cannot be bit-identical everywhere because
 * cv2.GaussianBlur (IPP) and numpy's SIMD expf/atan2f/powf are closed
 * implementations -- the measured agreement is written in DESIGN.md and
 * asserted in the tests.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction,
 * every float32 operation rounds once, as numpy's scalar/ufunc loops do).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_OCT 32
#define ORC_MAX_LAYERS 16
/* numpy npy_rad2degf: x * (180.0f / NPY_PIf), constant folded in float */
#define RAD2DEGF (180.0f / 3.141592653589793238462643383279502884f)

typedef struct {
    float x, y, size, angle, response;
    int32_t octave;
} orc_kp;

/* ------------------------------------------------------------------ */
/* third-party semantics                                               */
/* ------------------------------------------------------------------ */

/* cv2.cvtColor(COLOR_BGR2GRAY) on uint8 (sift_impl.py:27-28): fixed point,
 * (B*3735 + G*19235 + R*9798 + 16384) >> 15 (OpenCV 4.13 uses 15-bit
 * coefficients; checked against cv2 on 2M random pixels: 0 mismatches; the
 * 14-bit set 1868/9617/4899 mismatches 0.26 %). */
void orc_bgr2gray(const uint8_t *bgr, int h, int w, int stride, uint8_t *gray)
{
    for (int y = 0; y < h; ++y) {
        const uint8_t *p = bgr + (size_t)y * stride;
        for (int x = 0; x < w; ++x)
            gray[(size_t)y * w + x] =
                (uint8_t)((p[3 * x] * 3735 + p[3 * x + 1] * 19235 + p[3 * x + 2] * 9798 + 16384) >> 15);
    }
}

/* cv2.resize(img,(0,0),fx=2,fy=2,INTER_LINEAR) on float32 (sift_impl.py:53):
 * src = (dst+0.5)/2-0.5, weights {0.25,0.75}, replicate clamp.  OpenCV's
 * float path interpolates horizontally first, then vertically. */
void orc_resize2x_linear(const float *src, int h, int w, float *dst)
{
    int W = 2 * w, H = 2 * h;
    float *rows = (float *)malloc(sizeof(float) * (size_t)W * 2);
    int have[2] = {-1, -1};
    for (int Y = 0; Y < H; ++Y) {
        float fy = (float)((Y + 0.5) * 0.5 - 0.5);
        int sy = (int)floorf(fy);
        fy -= (float)sy;
        int sy0 = sy, sy1 = sy + 1;
        if (sy0 < 0) { sy0 = 0; }
        if (sy1 < 0) { sy1 = 0; }
        if (sy0 > h - 1) sy0 = h - 1;
        if (sy1 > h - 1) sy1 = h - 1;
        if (sy < 0) fy = 0.f, sy0 = sy1 = 0;
        if (sy >= h - 1) { fy = 0.f; sy0 = sy1 = h - 1; }
        int want[2] = {sy0, sy1};
        for (int k = 0; k < 2; ++k) {
            (void)have;
            const float *s = src + (size_t)want[k] * w;
            float *r = rows + (size_t)k * W;
            for (int X = 0; X < W; ++X) {
                float fx = (float)((X + 0.5) * 0.5 - 0.5);
                int sx = (int)floorf(fx);
                fx -= (float)sx;
                if (sx < 0) { fx = 0.f; sx = 0; }
                float a, b;
                if (sx >= w - 1) { fx = 0.f; sx = w - 1; a = s[sx]; b = s[sx]; }
                else { a = s[sx]; b = s[sx + 1]; }
                r[X] = a * (1.f - fx) + b * fx;
            }
        }
        float *d = dst + (size_t)Y * W;
        const float *r0 = rows, *r1 = rows + W;
        for (int X = 0; X < W; ++X)
            d[X] = r0[X] * (1.f - fy) + r1[X] * fy;
    }
    free(rows);
}

/* cvRound(sigma*8+1)|1  (createGaussianKernels, CV_32F depth uses 4 sigma). */
int orc_gaussian_ksize(double sigma)
{
    int k = (int)rint(sigma * 4 * 2 + 1);
    return k | 1;
}

/* cv2.getGaussianKernel(ksize, sigma, CV_32F) for sigma > 0: double exp,
 * double normalisation, cast to float. */
void orc_gaussian_kernel(int ksize, double sigma, float *taps)
{
    double tmp[1024];
    double scale2x = -0.5 / (sigma * sigma), sum = 0;
    for (int i = 0; i < ksize; ++i) {
        double x = i - (ksize - 1) * 0.5;
        tmp[i] = exp(scale2x * x * x);
        sum += tmp[i];
    }
    sum = 1. / sum;
    for (int i = 0; i < ksize; ++i)
        taps[i] = (float)(tmp[i] * sum);
}

/* cv::borderInterpolate(BORDER_REFLECT_101) */
static inline int reflect101(int p, int len)
{
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

/* cv2.GaussianBlur(img,(0,0),sigma) on float32 (sift_impl.py:56,91):
 * separable, rows first then columns, float32 accumulation, symmetric taps
 * folded as k0*c + sum_k k[k]*(a[+k]+a[-k]) like OpenCV's SymmRowVec /
 * SymmColumnVec float kernels.  Not bit-identical with the IPP build (see
 * header); max abs difference ~1e-4 on a 0..255 range. */
void orc_gaussian_blur(const float *src, int h, int w, double sigma, float *dst)
{
    int ks = orc_gaussian_ksize(sigma);
    int r = ks / 2;
    float taps[1024];
    orc_gaussian_kernel(ks, sigma, taps);
    float *tmp = (float *)malloc(sizeof(float) * (size_t)h * w);
    int *ix = (int *)malloc(sizeof(int) * (size_t)(w + 2 * r));
    for (int x = -r; x < w + r; ++x) ix[x + r] = reflect101(x, w);
    for (int y = 0; y < h; ++y) {
        const float *s = src + (size_t)y * w;
        float *t = tmp + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            float acc = taps[r] * s[x];
            for (int k = 1; k <= r; ++k)
                acc += taps[r + k] * (s[ix[x + k + r]] + s[ix[x - k + r]]);
            t[x] = acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        float *d = dst + (size_t)y * w;
        const float *c = tmp + (size_t)y * w;
        for (int x = 0; x < w; ++x) d[x] = taps[r] * c[x];
        for (int k = 1; k <= r; ++k) {
            const float *a = tmp + (size_t)reflect101(y + k, h) * w;
            const float *b = tmp + (size_t)reflect101(y - k, h) * w;
            float tk = taps[r + k];
            for (int x = 0; x < w; ++x) d[x] += tk * (a[x] + b[x]);
        }
    }
    free(ix);
    free(tmp);
}

/* cv2.resize(base,(w//2,h//2),INTER_NEAREST) (sift_impl.py:96) == [::2,::2]. */
void orc_decimate(const float *src, int h, int w, float *dst)
{
    int H = h / 2, W = w / 2;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) dst[(size_t)y * W + x] = src[(size_t)(2 * y) * w + 2 * x];
}

/* sift_impl.py:59-63 */
int orc_num_octaves(int h, int w)
{
    int m = h < w ? h : w;
    return (int)rint(log((double)m) / log(2.0) - 1.0);
}

/* sift_impl.py:66-79 */
void orc_gaussian_sigmas(double sigma, int num_intervals, double *out)
{
    int n = num_intervals + 3;
    double k = pow(2.0, 1. / num_intervals);
    out[0] = sigma;
    for (int i = 1; i < n; ++i) {
        double prev = pow(k, (double)(i - 1)) * sigma;
        double tot = k * prev;
        out[i] = sqrt(tot * tot - prev * prev);
    }
}

/* ------------------------------------------------------------------ */
/* pyramid view                                                        */
/* ------------------------------------------------------------------ */
typedef struct {
    int n_oct, n_layers; /* n_layers = num_intervals + 3 Gaussian layers */
    const float *const *layers; /* [n_oct * n_layers] */
    const int *h, *w;
} pyr_t;

static inline float G(const pyr_t *p, int o, int l, int y, int x)
{
    return p->layers[o * p->n_layers + l][(size_t)y * p->w[o] + x];
}
/* sift_impl.py:109 second - first, float32 */
static inline float DOG(const pyr_t *p, int o, int l, int y, int x)
{
    return G(p, o, l + 1, y, x) - G(p, o, l, y, x);
}

/* sift_impl.py:143-163 -- ties pass, strict |v| > threshold */
static int is_extremum(const pyr_t *p, int o, int l, int y, int x, double thresh)
{
    float v = DOG(p, o, l, y, x);
    if (fabs((double)v) <= thresh) return 0;
    if (v > 0) {
        for (int dl = -1; dl <= 1; ++dl)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx)
                    if (!(v >= DOG(p, o, l + dl, y + dy, x + dx))) return 0;
    } else {
        for (int dl = -1; dl <= 1; ++dl)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx)
                    if (!(v <= DOG(p, o, l + dl, y + dy, x + dx))) return 0;
    }
    return 1;
}

/* Minimum-norm least squares x = pinv(H) g for symmetric 3x3 H in float64:
 * what LAPACK gelsd (np.linalg.lstsq, rcond=None -> eps*max(M,N)) returns.
 * For symmetric H the SVD is the eigen-decomposition; cyclic Jacobi. */
static void sym3_pinv_solve(const double H[3][3], const double g[3], double x[3])
{
    double a[3][3], v[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { a[i][j] = H[i][j]; v[i][j] = (i == j); }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    double lmax = fmax(fabs(a[0][0]), fmax(fabs(a[1][1]), fabs(a[2][2])));
    double cut = 2.220446049250313e-16 * 3.0 * lmax;
    x[0] = x[1] = x[2] = 0.0;
    for (int i = 0; i < 3; ++i) {
        double lam = a[i][i];
        if (!(fabs(lam) > cut)) continue;
        double proj = (v[0][i] * g[0] + v[1][i] * g[1] + v[2][i] * g[2]) / lam;
        x[0] += proj * v[0][i];
        x[1] += proj * v[1][i];
        x[2] += proj * v[2][i];
    }
}

/* sift_impl.py:169-211.  Returns 1 and fills *kp, *layer_out when a keypoint
 * is produced, else 0.  Reproduces the reference's missing "did not
 * converge" rejection: after max_iter non-converged steps the last cube /
 * gradient / Hessian / update are used with the moved x, y, layer. */
int orc_localize(const float *const *layers, const int *hs, const int *ws, int n_oct, int n_layers,
                 int x, int y, int layer, int octave, int num_intervals, double sigma,
                 double contrast_threshold, int border, double eigen_ratio, int max_iter,
                 orc_kp *kp, int *layer_out)
{
    pyr_t P = {n_oct, n_layers, layers, hs, ws};
    const pyr_t *p = &P;
    int h = hs[octave], w = ws[octave];
    float cube[3][3][3], grad[3] = {0, 0, 0}, hess[3][3] = {{0}}, upd[3] = {0, 0, 0};
    for (int it = 0; it < max_iter; ++it) {
        for (int s = 0; s < 3; ++s)
            for (int j = 0; j < 3; ++j)
                for (int i = 0; i < 3; ++i)
                    cube[s][j][i] = DOG(p, octave, layer - 1 + s, y - 1 + j, x - 1 + i) / 255.f;
        /* :217-224 */
        grad[0] = 0.5f * (cube[1][1][2] - cube[1][1][0]);
        grad[1] = 0.5f * (cube[1][2][1] - cube[1][0][1]);
        grad[2] = 0.5f * (cube[2][1][1] - cube[0][1][1]);
        /* :227-240 */
        float v = cube[1][1][1];
        float dxx = cube[1][1][2] - 2 * v + cube[1][1][0];
        float dyy = cube[1][2][1] - 2 * v + cube[1][0][1];
        float dss = cube[2][1][1] - 2 * v + cube[0][1][1];
        float dxy = 0.25f * (cube[1][2][2] - cube[1][2][0] - cube[1][0][2] + cube[1][0][0]);
        float dxs = 0.25f * (cube[2][1][2] - cube[2][1][0] - cube[0][1][2] + cube[0][1][0]);
        float dys = 0.25f * (cube[2][2][1] - cube[2][0][1] - cube[0][2][1] + cube[0][0][1]);
        hess[0][0] = dxx; hess[0][1] = dxy; hess[0][2] = dxs;
        hess[1][0] = dxy; hess[1][1] = dyy; hess[1][2] = dys;
        hess[2][0] = dxs; hess[2][1] = dys; hess[2][2] = dss;
        double Hd[3][3], gd[3], xd[3];
        for (int i = 0; i < 3; ++i) {
            gd[i] = grad[i];
            for (int j = 0; j < 3; ++j) Hd[i][j] = hess[i][j];
        }
        sym3_pinv_solve(Hd, gd, xd);
        for (int i = 0; i < 3; ++i) upd[i] = -(float)xd[i];
        if (fabsf(upd[0]) < 0.5f && fabsf(upd[1]) < 0.5f && fabsf(upd[2]) < 0.5f) break;
        x += (int)rintf(upd[0]);
        y += (int)rintf(upd[1]);
        layer += (int)rintf(upd[2]);
        if (y < border || y >= h - border || x < border || x >= w - border || layer < 1 ||
            layer > num_intervals)
            return 0;
    }
    /* val = cube[1,1,1] + 0.5*np.dot(grad, update): numpy's float32 dot of a
     * 3-vector rounds each product to float32 and sums them in float64. */
    float p0 = grad[0] * upd[0], p1 = grad[1] * upd[1], p2 = grad[2] * upd[2];
    float dot = (float)((double)p0 + (double)p1 + (double)p2);
    float val = cube[1][1][1] + 0.5f * dot;
    if (fabsf(val) * (float)num_intervals < (float)contrast_threshold) return 0;
    float tr = hess[0][0] + hess[1][1];
    /* np.linalg.det on float32 2x2: LU with partial pivoting in float64, cast. */
    double a = hess[0][0], b = hess[0][1], c = hess[1][0], d = hess[1][1], det;
    if (fabs(a) >= fabs(c)) {
        if (a == 0.0) det = 0.0;
        else { double l = c / a; det = a * (d - l * b); }
    } else {
        double l = a / c;
        det = -(c * (b - l * d));
    }
    float detf = (float)det;
    float er = (float)eigen_ratio;
    if (detf <= 0 || er * (tr * tr) >= ((er + 1) * (er + 1)) * detf) return 0;
    float sc = (float)(1 << octave); /* python int 2**octave, weak -> float32 */
    kp->x = ((float)x + upd[0]) * sc;
    kp->y = ((float)y + upd[1]) * sc;
    kp->octave = octave + layer * 256 + (int)rintf((upd[2] + 0.5f) * 255.f) * 65536;
    float e = ((float)layer + upd[2]) / (float)num_intervals;
    kp->size = (float)sigma * powf(2.f, e) * (float)(1 << (octave + 1));
    kp->response = fabsf(val);
    kp->angle = -1.f;
    *layer_out = layer;
    return 1;
}

/* np.float32 % 360 (python modulo, float32 remainder loop) */
static inline float pymodf(float a, float b)
{
    float m = fmodf(a, b);
    if (m != 0.f) {
        if ((b < 0) != (m < 0)) m += b;
    } else {
        m = copysignf(0.f, b);
    }
    return m;
}
static inline double pymod(double a, double b)
{
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0) != (m < 0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

/* sift_impl.py:246-293.  img = gauss[octave][layer] (h x w).  Returns the
 * number of oriented keypoints written to out (at most num_bins). */
int orc_orientations(const orc_kp *kp, int octave, const float *img, int h, int w,
                     double radius_factor, int num_bins, double peak_ratio, double scale_factor,
                     orc_kp *out)
{
    /* scale_factor*size is python-float (double) arithmetic, then / np.float32 -> float32 */
    float scale = (float)(scale_factor * (double)kp->size) / (float)(1 << (octave + 1));
    int radius = (int)rintf((float)radius_factor * scale);
    float weight_fac = -0.5f / (scale * scale);
    double raw[64], smooth[64];
    for (int i = 0; i < num_bins; ++i) raw[i] = 0.0;
    int cy = (int)rintf(kp->y / (float)(1 << octave));
    int cx = (int)rintf(kp->x / (float)(1 << octave));
    for (int dy = -radius; dy <= radius; ++dy) {
        int y = cy + dy;
        if (y <= 0 || y >= h - 1) continue;
        for (int dx = -radius; dx <= radius; ++dx) {
            int x = cx + dx;
            if (x <= 0 || x >= w - 1) continue;
            float gx = img[(size_t)y * w + x + 1] - img[(size_t)y * w + x - 1];
            float gy = img[(size_t)(y - 1) * w + x] - img[(size_t)(y + 1) * w + x];
            float mag = sqrtf(gx * gx + gy * gy);
            float ang = pymodf(atan2f(gy, gx) * RAD2DEGF, 360.f);
            float wgt = expf(weight_fac * (float)(dx * dx + dy * dy));
            int idx = (int)rintf(ang * (float)num_bins / 360.f) % num_bins;
            raw[idx] += (double)(wgt * mag);
        }
    }
    double maxv = -1e300;
    for (int i = 0; i < num_bins; ++i) {
        int im1 = (i - 1 + num_bins) % num_bins, im2 = (i - 2 + num_bins) % num_bins;
        smooth[i] = (6 * raw[i] + 4 * (raw[im1] + raw[(i + 1) % num_bins]) + raw[im2] +
                     raw[(i + 2) % num_bins]) / 16.;
        if (smooth[i] > maxv) maxv = smooth[i];
    }
    int n = 0;
    for (int pk = 0; pk < num_bins; ++pk) {
        double l = smooth[(pk - 1 + num_bins) % num_bins], r = smooth[(pk + 1) % num_bins];
        if (!(smooth[pk] > l && smooth[pk] > r)) continue;
        if (!(smooth[pk] >= peak_ratio * maxv)) continue;
        double interp = pymod(pk + 0.5 * (l - r) / (l - 2 * smooth[pk] + r), (double)num_bins);
        double angle = 360. - interp * 360. / num_bins;
        if (fabs(angle - 360.) < 1e-7) angle = 0;
        out[n] = *kp;
        out[n].angle = (float)angle;
        ++n;
    }
    return n;
}

/* sift_impl.py:117-140.  Scan order o, i, y, x; appends oriented keypoints.
 * cand (optional, 4 ints per candidate: octave, layer, y, x) receives every
 * pixel that passed is_pixel_an_extremum; stats[0]=#candidates,
 * stats[1]=#localized.  Returns number of keypoints, or -1 on overflow. */
int orc_find_scale_space_extrema(const float *const *layers, const int *hs, const int *ws,
                                 int n_oct, int num_intervals, double sigma, int border,
                                 double contrast_threshold, orc_kp *out, int cap, int *cand,
                                 int cand_cap, int *stats)
{
    int n_layers = num_intervals + 3;
    pyr_t P = {n_oct, n_layers, layers, hs, ws};
    double thresh = floor(0.5 * contrast_threshold / num_intervals * 255);
    int n = 0, nc = 0, nl = 0;
    for (int o = 0; o < n_oct; ++o) {
        int h = hs[o], w = ws[o];
        for (int i = 0; i < n_layers - 3; ++i) {
            for (int y = border; y < h - border; ++y)
                for (int x = border; x < w - border; ++x) {
                    if (!is_extremum(&P, o, i + 1, y, x, thresh)) continue;
                    if (cand && nc < cand_cap) {
                        cand[4 * nc] = o; cand[4 * nc + 1] = i + 1; cand[4 * nc + 2] = y; cand[4 * nc + 3] = x;
                    }
                    ++nc;
                    orc_kp kp;
                    int lyr;
                    if (!orc_localize(layers, hs, ws, n_oct, n_layers, x, y, i + 1, o, num_intervals,
                                      sigma, contrast_threshold, border, 10.0, 5, &kp, &lyr))
                        continue;
                    ++nl;
                    orc_kp tmp[64];
                    int k = orc_orientations(&kp, o, layers[o * n_layers + lyr], h, w, 3.0, 36, 0.8, 1.5, tmp);
                    if (n + k > cap) return -1;
                    memcpy(out + n, tmp, sizeof(orc_kp) * k);
                    n += k;
                }
        }
    }
    if (stats) { stats[0] = nc; stats[1] = nl; }
    return n;
}

/* sift_impl.py:299-311 (class_id is always -1) */
static int kp_cmp(const orc_kp *a, const orc_kp *b)
{
    if (a->x != b->x) return a->x < b->x ? -1 : 1;
    if (a->y != b->y) return a->y < b->y ? -1 : 1;
    if (a->size != b->size) return b->size < a->size ? -1 : 1;
    if (a->angle != b->angle) return a->angle < b->angle ? -1 : 1;
    if (a->response != b->response) return b->response < a->response ? -1 : 1;
    return 0;
}

static void merge_sort(orc_kp *a, orc_kp *tmp, int n)
{
    if (n < 2) return;
    int m = n / 2;
    merge_sort(a, tmp, m);
    merge_sort(a + m, tmp, n - m);
    int i = 0, j = m, k = 0;
    while (i < m && j < n) tmp[k++] = (kp_cmp(&a[j], &a[i]) < 0) ? a[j++] : a[i++];
    while (i < m) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(orc_kp) * n);
}

/* sift_impl.py:314-327: stable sort, drop a keypoint whose (pt,size,angle)
 * equal those of the previously kept one.  In place; returns new count. */
int orc_remove_duplicates(orc_kp *kps, int n)
{
    if (n < 2) return n;
    orc_kp *tmp = (orc_kp *)malloc(sizeof(orc_kp) * n);
    merge_sort(kps, tmp, n);
    free(tmp);
    int m = 1;
    for (int i = 1; i < n; ++i) {
        const orc_kp *last = &kps[m - 1], *k = &kps[i];
        if (last->x != k->x || last->y != k->y || last->size != k->size || last->angle != k->angle)
            kps[m++] = *k;
    }
    return m;
}

/* sift_impl.py:333-343 */
void orc_convert_to_input_size(orc_kp *kps, int n)
{
    for (int i = 0; i < n; ++i) {
        kps[i].x *= 0.5f;
        kps[i].y *= 0.5f;
        kps[i].size *= 0.5f;
        kps[i].octave = (kps[i].octave & ~255) | ((kps[i].octave - 1) & 255);
    }
}

/* sift_impl.py:361-526 */
void orc_descriptors(const orc_kp *kps, int n, const float *const *layers, const int *hs,
                     const int *ws, int n_oct, int n_layers, int window_width, int num_bins,
                     double scale_multiplier, double descriptor_max_value, float *out)
{
    (void)n_oct;
    int tw = window_width + 2;
    float *tensor = (float *)malloc(sizeof(float) * tw * tw * num_bins);
    int *cell0 = NULL, *obin = NULL;
    double *c8 = NULL;
    size_t scratch_cap = 0;
    int dlen = window_width * window_width * num_bins;
    for (int ki = 0; ki < n; ++ki) {
        const orc_kp *kp = &kps[ki];
        float *vec = out + (size_t)ki * dlen;
        /* :349-358 unpack_octave */
        int octv = kp->octave & 255, lyr = (kp->octave >> 8) & 255;
        if (octv >= 128) octv |= -128;
        float scl = octv >= 0 ? 1.f / (float)(1 << octv) : (float)(1 << -octv);
        const float *img = layers[(octv + 1) * n_layers + lyr];
        int rows = hs[octv + 1], cols = ws[octv + 1];
        int ptx = (int)rint((double)scl * (double)kp->x);
        int pty = (int)rint((double)scl * (double)kp->y);
        double angle = 360. - (double)kp->angle;
        double rad = angle * (M_PI / 180.0); /* np.deg2rad */
        double cos_a = cos(rad), sin_a = sin(rad);
        memset(tensor, 0, sizeof(float) * tw * tw * num_bins);
        float hist_width = (float)(scale_multiplier * 0.5) * scl * kp->size;
        int half_w = (int)rint((double)hist_width * sqrt(2.0) * (window_width + 1) * 0.5);
        int diag = (int)sqrt((double)((long long)rows * rows + (long long)cols * cols));
        if (diag < half_w) half_w = diag;
        double hw = (double)hist_width;
        float anglef = (float)angle;
        float bins_per_deg = (float)(num_bins / 360.);
        double weight_mul = -0.5 / ((0.5 * window_width) * (0.5 * window_width));
        /* per-pixel quantities (:389-466), valid pixels kept in row-major order */
        int side = 2 * half_w + 1;
        size_t need = (size_t)side * side;
        if (need > scratch_cap) {
            scratch_cap = need;
            free(cell0); free(obin); free(c8);
            cell0 = (int *)malloc(sizeof(int) * need);
            obin = (int *)malloc(sizeof(int) * need);
            c8 = (double *)malloc(sizeof(double) * need * 8);
        }
        size_t np_ = 0;
        for (int ys = -half_w; ys <= half_w; ++ys) {
            int rr = pty + ys;
            if (!(rr > 0 && rr < rows - 1)) continue;
            for (int xs = -half_w; xs <= half_w; ++xs) {
                int cc = ptx + xs;
                if (!(cc > 0 && cc < cols - 1)) continue;
                double r_rot = xs * sin_a + ys * cos_a;
                double c_rot = xs * cos_a - ys * sin_a;
                double r_bin = (r_rot / hw) + 0.5 * window_width - 0.5;
                double c_bin = (c_rot / hw) + 0.5 * window_width - 0.5;
                if (!(r_bin > -1.0 && r_bin < window_width && c_bin > -1.0 && c_bin < window_width))
                    continue;
                float gx = img[(size_t)rr * cols + cc + 1] - img[(size_t)rr * cols + cc - 1];
                float gy = img[(size_t)(rr - 1) * cols + cc] - img[(size_t)(rr + 1) * cols + cc];
                float mag = sqrtf(gx * gx + gy * gy);
                float orient = pymodf(atan2f(gy, gx) * RAD2DEGF, 360.f);
                double qr = r_rot / hw, qc = c_rot / hw;
                double wgt = exp(weight_mul * (qr * qr + qc * qc));
                double wmag = wgt * (double)mag;
                float ob = pymodf((orient - anglef) * bins_per_deg, (float)num_bins);
                long r0 = (long)floor(r_bin), c0 = (long)floor(c_bin);
                long o0 = (long)floorf(ob);
                o0 = ((o0 % num_bins) + num_bins) % num_bins;
                double rf = r_bin - (double)r0, cf = c_bin - (double)c0, of = (double)ob - (double)o0;
                double c1 = wmag * rf, c0w = wmag - c1;
                double c10 = c1 * (1 - cf), c11 = c1 * cf, c00 = c0w * (1 - cf), c01 = c0w * cf;
                double *q = c8 + np_ * 8;
                q[0] = c00 * (1 - of); q[1] = c00 * of;
                q[2] = c01 * (1 - of); q[3] = c01 * of;
                q[4] = c10 * (1 - of); q[5] = c10 * of;
                q[6] = c11 * (1 - of); q[7] = c11 * of;
                cell0[np_] = (int)((r0 + 1) * tw + (c0 + 1));
                obin[np_] = (int)o0;
                ++np_;
            }
        }
        /* the 8 np.add.at passes (:503-506, each scatter_orient = base then
         * plus) run one after the other over all pixels in row-major order. */
        for (int pass = 0; pass < 8; ++pass) {
            int dr = (pass >> 2) & 1, dc = (pass >> 1) & 1, dob = pass & 1;
            for (size_t i = 0; i < np_; ++i) {
                int ob_i = dob == 0 ? obin[i] % num_bins : (obin[i] + 1) % num_bins;
                float *cell = &tensor[(cell0[i] + dr * tw + dc) * num_bins + ob_i];
                /* np.add.at(float32 array, idx, float64): add in float64, store float32 */
                *cell = (float)((double)*cell + c8[i * 8 + pass]);
            }
        }
        for (int r = 0; r < window_width; ++r)
            for (int c = 0; c < window_width; ++c)
                for (int o = 0; o < num_bins; ++o)
                    vec[(r * window_width + c) * num_bins + o] = tensor[((r + 1) * tw + (c + 1)) * num_bins + o];
        /* :512-524; np.linalg.norm = sqrt(sdot(x,x)) (BLAS, order unspecified) */
        double ss = 0;
        for (int i = 0; i < dlen; ++i) ss += (double)(vec[i] * vec[i]);
        float thr = sqrtf((float)ss) * (float)descriptor_max_value;
        for (int i = 0; i < dlen; ++i)
            if (vec[i] > thr) vec[i] = thr;
        ss = 0;
        for (int i = 0; i < dlen; ++i) ss += (double)(vec[i] * vec[i]);
        float norm_v = sqrtf((float)ss);
        if (norm_v < 1e-7f) norm_v = 1e-7f;
        for (int i = 0; i < dlen; ++i) {
            float q = rintf(512.f * (vec[i] / norm_v));
            if (q < 0) q = 0;
            if (q > 255) q = 255;
            vec[i] = q;
        }
    }
    free(tensor);
    free(cell0); free(obin); free(c8);
}

/* ------------------------------------------------------------------ */
/* matcher / vote / projection                                         */
/* ------------------------------------------------------------------ */

/* image_stitching_sift.py:63-79: float32 difference, float32 dot, strict <.
 * (For integer-valued descriptors every partial sum is an exact integer
 * < 2^24, so the float32 result is order-independent.)  best_idx[i] = -1 and
 * best_d2[i] = inf when nb == 0. */
void orc_match_f32(const float *A, int na, const float *B, int nb, int dim, int *best_idx,
                   float *best_d2)
{
    for (int i = 0; i < na; ++i) {
        float best = INFINITY;
        int bi = -1;
        for (int j = 0; j < nb; ++j) {
            float s = 0;
            for (int k = 0; k < dim; ++k) {
                float d = A[(size_t)i * dim + k] - B[(size_t)j * dim + k];
                s += d * d;
            }
            if (s < best) { best = s; bi = j; }
        }
        best_idx[i] = bi;
        best_d2[i] = best;
    }
}

/* Exact integer nearest + second nearest on uint8 descriptors (the quantity
 * the fused top-2 epilogue produces; second = INT32_MAX when nb < 2). */
void orc_match_u8(const uint8_t *A, int na, const uint8_t *B, int nb, int dim, int32_t *best_idx,
                  int32_t *best_d2, int32_t *second_d2)
{
    for (int i = 0; i < na; ++i) {
        int32_t b1 = INT32_MAX, b2 = INT32_MAX, bi = -1;
        for (int j = 0; j < nb; ++j) {
            int32_t s = 0;
            for (int k = 0; k < dim; ++k) {
                int32_t d = (int32_t)A[(size_t)i * dim + k] - (int32_t)B[(size_t)j * dim + k];
                s += d * d;
            }
            if (s < b1) { b2 = b1; b1 = s; bi = j; }
            else if (s < b2) b2 = s;
        }
        best_idx[i] = bi;
        best_d2[i] = b1;
        second_d2[i] = b2;
    }
}

/* image_stitching_sift.py:86-111.  matches: n x 4 doubles (xA,yA,xB,yB).
 * Returns index of the winning match (first maximum), -1 if n == 0. */
int orc_ransac(const double *m, int n, double dist_sq_thresh, double *move)
{
    move[0] = move[1] = 0;
    int best = -1, best_score = -1;
    for (int i = 0; i < n; ++i) {
        double dxr = m[4 * i] - m[4 * i + 2], dyr = m[4 * i + 1] - m[4 * i + 3];
        int votes = 0;
        for (int j = 0; j < n; ++j) {
            double dx = (m[4 * j] - m[4 * j + 2]) - dxr, dy = (m[4 * j + 1] - m[4 * j + 3]) - dyr;
            if (dx * dx + dy * dy < dist_sq_thresh) ++votes;
        }
        if (votes > best_score) { best_score = votes; best = i; move[0] = dxr; move[1] = dyr; }
    }
    return best;
}

/* image_stitching_sift.py:117-136: forward map, python round() (half to
 * even), later source pixels overwrite earlier ones. */
void orc_cylindrical_projection(const uint8_t *src, int h, int w, int ch, double f, uint8_t *dst)
{
    memset(dst, 0, (size_t)h * w * ch);
    int cy = h / 2, cx = w / 2;
    for (int yy = 0; yy < h; ++yy)
        for (int xx = 0; xx < w; ++xx) {
            int xd = xx - cx, yd = yy - cy;
            long xm = (long)rint(f * atan((double)xd / f)) + cx;
            double denom = sqrt((double)(xd * xd) + f * f);
            long ym = (long)rint(f * ((double)yd / denom)) + cy;
            if (xm >= 0 && xm < w && ym >= 0 && ym < h)
                memcpy(dst + ((size_t)ym * w + xm) * ch, src + ((size_t)yy * w + xx) * ch, ch);
        }
}
