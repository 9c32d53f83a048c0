// Brute-force descriptor matcher, host orchestration + the small kernels around the contraction.
//
// Replaces /root/reference/image_stitching_sift.py:63-79 (nearest-neighbour loop + acceptance) and
// :86-111 (ransac() vote).  The contraction itself runs on the tensor cores (match_tc.cu,
// tcgen05.mma kind::i8); this file holds
//   * the merge of per-chunk top-2 results (B is split into chunks so that small problems still
//     fill the GPU; chunks ascend in j, so a strict "<" keeps the lowest j on ties),
//   * pair_finalize_kernel: acceptance best < desc_thresh (:74), ordered compaction of the match
//     list and the translation vote, one CTA per image pair,
//   * ratio_accept_kernel: the nearest / second-nearest ratio test of sift_visualizeUI.py:247-257
//     in exact integers.
// Descriptors are the 0..255 integers generate_descriptors emits (sift_impl.py:519-524), so
// |a-b|^2 = |a|^2 + |b|^2 - 2 a.b <= 128*255^2 < 2^24 is exact in int32 (and equals the
// reference's float32 value bit for bit).
#include <cub/device/device_scan.cuh>
#include <limits.h>
#include <string.h>
#include "common.cuh"

namespace b200 {

// match_tc.cu
int run_match_tc(b200sift_ctx *c, const uint8_t *d_src, int n_imgs, const int *h_src_off, const int *h_n,
                 int n_pairs, const int *h_pairs, int rows_max, int n_chunks_out, int tiles_per_chunk, bool top2,
                 int32_t *d_part);
void tc_chunking(const b200sift_ctx *c, int rows_max, int nb_max, int n_pairs, int *tiles_per_chunk, int *n_chunks);

// per-chunk (best j, best d, second d) -> global top-2; chunks ascend in j
__global__ void __launch_bounds__(256)
match_merge_kernel(const int32_t *__restrict__ part, int nA, int n_chunks, int32_t *__restrict__ best_idx,
                   int32_t *__restrict__ best_d2, int32_t *__restrict__ second_d2)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= nA) return;
    int b1 = INT_MAX, b2 = INT_MAX, bi = -1;
    for (int c = 0; c < n_chunks; ++c) {
        const int32_t *o = part + ((size_t)i * n_chunks + c) * 3;
        const int ci = o[0], c1 = o[1], c2 = o[2];
        if (ci < 0) continue;
        if (c1 < b1) { b2 = min(b1, c2); b1 = c1; bi = ci; }
        else { b2 = min(b2, c1); }
    }
    best_idx[i] = bi;
    best_d2[i] = b1;
    if (second_d2) second_d2[i] = b2;
}

// Generic A (nA,128) x B (nB,128), device pointers, row-major uint8.
int run_match(b200sift_ctx *c, const uint8_t *dA, int nA, const uint8_t *dB, int nB, int32_t *d_best_idx,
              int32_t *d_best_d2, int32_t *d_second_d2)
{
    if (nA <= 0) return 0;
    size_t cap;
    // tensor-core path: A and B become "image" 0 and 1 of one packed buffer
    int tpc, n_chunks;
    tc_chunking(c, nA, nB, 1, &tpc, &n_chunks);
    cap = c->mout_cap;
    B200_CHECK(ensure(&c->d_mout, &cap, (size_t)nA * n_chunks * 3));
    c->mout_cap = cap;
    // both sets must be addressable from one base pointer: stage them side by side
    cap = c->tcsrc_cap;
    B200_CHECK(ensure(&c->d_tcsrc, &cap, (size_t)(nA + (nB > 0 ? nB : 0)) * 128));
    c->tcsrc_cap = cap;
    B200_CUDA(cudaMemcpyAsync(c->d_tcsrc, dA, (size_t)nA * 128, cudaMemcpyDeviceToDevice, c->stream));
    if (nB > 0)
        B200_CUDA(cudaMemcpyAsync(c->d_tcsrc + (size_t)nA * 128, dB, (size_t)nB * 128, cudaMemcpyDeviceToDevice,
                                  c->stream));
    const int src_off[2] = {0, nA}, ns[2] = {nA, nB}, pr[2] = {0, 1};
    B200_CHECK(run_match_tc(c, c->d_tcsrc, 2, src_off, ns, 1, pr, nA, n_chunks, tpc, d_second_d2 != nullptr,
                            c->d_mout));
    match_merge_kernel<<<(nA + 255) / 256, 256, 0, c->stream>>>(c->d_mout, nA, n_chunks, d_best_idx, d_best_d2,
                                                                d_second_d2);
    c->launches++;
    B200_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// Batched form: every requested pair of the last detect in one pass
// (image_stitching_sift.py:312-327 runs compute_shift_sift per adjacent pair).
// pair_finalize_kernel: one CTA per pair merges the chunks, accepts best < thresh (:74),
// compacts in A order (block scan), then runs the ransac() vote (:86-111, float64, first
// maximum wins).
// ---------------------------------------------------------------------------
constexpr int kFinThreads = 1024;

// device-counted images (RemoteImage): the real keypoint count replaces the host's upper bound
__global__ void __launch_bounds__(128)
patch_remote_pairs_kernel(PairDesc *__restrict__ pd, int n_pairs, int image, const int32_t *__restrict__ d_count, int cap)
{
    const int n = max(0, min(*d_count, cap));
    for (int p = threadIdx.x; p < n_pairs; p += blockDim.x) {
        if (pd[p].imgA == image) pd[p].nA = n;
        if (pd[p].imgB == image) pd[p].nB = n;
    }
}

__global__ void __launch_bounds__(kFinThreads)
pair_finalize_kernel(const PairDesc *__restrict__ pd, const int32_t *__restrict__ part, int n_chunks_max,
                     int rows_max, const b200sift_keypoint *__restrict__ kps, int thresh, double vote_thr,
                     int32_t *__restrict__ m_ia, int32_t *__restrict__ m_ib, float *__restrict__ m_xy,
                     double *__restrict__ m_mv, PairResult *__restrict__ res)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    __shared__ unsigned long long s_red[32];
    const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const PairDesc P = pd[pair];
    const b200sift_keypoint *kA = kps + P.offA, *kB = kps + P.offB;
    const size_t mo = (size_t)pair * rows_max;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int row0 = 0; row0 < P.nA; row0 += kFinThreads) {
        const int i = row0 + tid;
        int b1 = INT_MAX, bi = -1;
        if (i < P.nA) {
            const int32_t *o = part + ((size_t)pair * rows_max + i) * n_chunks_max * 3;
            for (int c = 0; c < n_chunks_max; ++c) {
                const int ci = o[3 * c], c1 = o[3 * c + 1];
                if (ci >= 0 && c1 < b1) { b1 = c1; bi = ci; }   // strict <: the earlier chunk (lower j) wins ties
            }
        }
        const bool keep = (bi != -1) && (b1 < thresh);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = s_base, total = 0;
        for (int w2 = 0; w2 < kFinThreads / 32; ++w2) {
            const int c = s_warp[w2];
            if (w2 < warp) before += c;
            total += c;
        }
        if (keep) {
            const int d = before + __popc(m & ((1u << lane) - 1u));
            m_ia[mo + d] = i;
            m_ib[mo + d] = bi;
            const float xa = kA[i].x, ya = kA[i].y, xb = kB[bi].x, yb = kB[bi].y;
            m_xy[4 * (mo + d)] = xa; m_xy[4 * (mo + d) + 1] = ya; m_xy[4 * (mo + d) + 2] = xb; m_xy[4 * (mo + d) + 3] = yb;
            m_mv[2 * (mo + d)] = (double)xa - (double)xb;       // :94-96 python floats
            m_mv[2 * (mo + d) + 1] = (double)ya - (double)yb;
        }
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
    }
    const int n = s_base;
    unsigned long long key = 0;
    for (int i = tid; i < n; i += kFinThreads) {
        const double dxr = m_mv[2 * (mo + i)], dyr = m_mv[2 * (mo + i) + 1];
        unsigned votes = 0;
        for (int j = 0; j < n; ++j) {
            const double dx = m_mv[2 * (mo + j)] - dxr, dy = m_mv[2 * (mo + j) + 1] - dyr;
            votes += (dx * dx + dy * dy < vote_thr) ? 1u : 0u;
        }
        const unsigned long long k = ((unsigned long long)votes << 32) | (unsigned)(~(unsigned)i);
        if (k > key) key = k;
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, sft);
        if (o > key) key = o;
    }
    if (lane == 0) s_red[warp] = key;
    __syncthreads();
    if (tid == 0) {
        for (int w2 = 1; w2 < kFinThreads / 32; ++w2)
            if (s_red[w2] > key) key = s_red[w2];
        PairResult r;
        r.n_matches = n;
        if (n > 0) {
            const int best = (int)(~(unsigned)(key & 0xffffffffu));
            r.best = best;
            r.dx = m_mv[2 * (mo + best)];
            r.dy = m_mv[2 * (mo + best) + 1];
            for (int k2 = 0; k2 < 4; ++k2) r.xyxy[k2] = m_xy[4 * (mo + best) + k2];
        } else {
            r.best = -1;
            r.dx = r.dy = 0;
            for (int k2 = 0; k2 < 4; ++k2) r.xyxy[k2] = 0.f;
        }
        res[pair] = r;
    }
}

// Matcher + finalize for n_pairs (imgA, imgB) index pairs of the last detect; results land in
// c->d_pair_res / c->d_pair_ia / ...
int run_match_pairs(b200sift_ctx *c, int n_pairs, const int *h_pairs, int thresh, double vote_thr)
{
    const int n_img = c->n_img_last;
    std::vector<PairDesc> &h_pd = c->h_pair_desc;  // read by an asynchronous copy below
    h_pd.assign(n_pairs, PairDesc{0, 0, 0, 0, 0, 0});
    int rows_max = 1, nb_max = 0;
    for (int p = 0; p < n_pairs; ++p) {
        const int a = h_pairs[2 * p], b = h_pairs[2 * p + 1];
        h_pd[p].offA = c->img_off[a];
        h_pd[p].nA = c->img_off[a + 1] - c->img_off[a];
        h_pd[p].offB = c->img_off[b];
        h_pd[p].nB = c->img_off[b + 1] - c->img_off[b];
        h_pd[p].imgA = a;
        h_pd[p].imgB = b;
        rows_max = h_pd[p].nA > rows_max ? h_pd[p].nA : rows_max;
        nb_max = h_pd[p].nB > nb_max ? h_pd[p].nB : nb_max;
    }
    int tpc = 1, n_chunks;
    tc_chunking(c, rows_max, nb_max, n_pairs, &tpc, &n_chunks);
    size_t cap = c->mout_cap;
    B200_CHECK(ensure(&c->d_mout, &cap, (size_t)n_pairs * rows_max * n_chunks * 3));
    c->mout_cap = cap;
    // per-pair scratch: mv | res | xy | pd | ia | ib
    const size_t rows = (size_t)n_pairs * rows_max;
    const size_t bytes = (size_t)n_pairs * (sizeof(PairDesc) + sizeof(PairResult)) + rows * (4 + 4 + 16 + 16) + 256;
    cap = c->pair_cap;
    B200_CHECK(ensure(&c->d_pair, &cap, bytes));
    c->pair_cap = cap;
    uint8_t *base = c->d_pair;
    double *m_mv = (double *)base;                     base += rows * 16;
    PairResult *res = (PairResult *)base;              base += (size_t)n_pairs * sizeof(PairResult);
    float *m_xy = (float *)base;                       base += rows * 16;
    PairDesc *pd = (PairDesc *)base;                   base += (size_t)n_pairs * sizeof(PairDesc);
    int32_t *m_ia = (int32_t *)base;                   base += rows * 4;
    int32_t *m_ib = (int32_t *)base;
    c->pair_rows_max = rows_max;
    c->pair_n = n_pairs;
    c->d_pair_res = res; c->d_pair_ia = m_ia; c->d_pair_ib = m_ib; c->d_pair_xy = m_xy;
    B200_CUDA(cudaMemcpyAsync(pd, h_pd.data(), sizeof(PairDesc) * n_pairs, cudaMemcpyHostToDevice, c->stream));
    for (const RemoteImage &r : c->remote) {
        patch_remote_pairs_kernel<<<1, 128, 0, c->stream>>>(pd, n_pairs, r.image, r.d_count, r.cap);
        c->launches++;
    }
    {
        std::vector<int> off(n_img), cnt(n_img);
        for (int i = 0; i < n_img; ++i) { off[i] = c->img_off[i]; cnt[i] = c->img_off[i + 1] - c->img_off[i]; }
        B200_CHECK(run_match_tc(c, c->d_desc, n_img, off.data(), cnt.data(), n_pairs, h_pairs, rows_max, n_chunks, tpc,
                                false, c->d_mout));
    }
    pair_finalize_kernel<<<n_pairs, kFinThreads, 0, c->stream>>>(pd, c->d_mout, n_chunks, rows_max, c->d_kps, thresh,
                                                                vote_thr, m_ia, m_ib, m_xy, m_mv, res);
    c->launches++;
    B200_CUDA(cudaGetLastError());
    return 0;
}

// accepted matches of image_stitching_sift.py:74-79, compacted in A order (single pair, used by
// b200sift_match_images)
__global__ void __launch_bounds__(256)
accept_flag_kernel(const int32_t *__restrict__ best_idx, const int32_t *__restrict__ best_d2, int nA, int thresh,
                   uint32_t *__restrict__ keep)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < nA) keep[i] = (best_idx[i] != -1 && best_d2[i] < thresh) ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
accept_scatter_kernel(const int32_t *__restrict__ best_idx, const uint32_t *__restrict__ keep,
                      const uint32_t *__restrict__ pos, int nA, const b200sift_keypoint *__restrict__ kA,
                      const b200sift_keypoint *__restrict__ kB, int32_t *__restrict__ ia, int32_t *__restrict__ ib,
                      float *__restrict__ xyxy, int32_t *__restrict__ count)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= nA) return;
    if (i == nA - 1) *count = (int32_t)(pos[i] + keep[i]);
    if (!keep[i]) return;
    const uint32_t d = pos[i];
    const int j = best_idx[i];
    ia[d] = i;
    ib[d] = j;
    xyxy[4 * d] = kA[i].x; xyxy[4 * d + 1] = kA[i].y; xyxy[4 * d + 2] = kB[j].x; xyxy[4 * d + 3] = kB[j].y;
}

// exclusive scan of c->d_keep[0..n) into c->d_pos on the context stream
static int scan_keep(b200sift_ctx *c, int n)
{
    size_t tmp = 0;
    B200_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->d_keep, c->d_pos, n, c->stream));
    if (tmp > c->cub_tmp_cap) {
        B200_CUDA(b200::ctx_sync(c));
        if (c->d_cub_tmp) cudaFree(c->d_cub_tmp);
        c->d_cub_tmp = nullptr;
        B200_CUDA(cudaMalloc(&c->d_cub_tmp, tmp + 1024));
        c->cub_tmp_cap = tmp + 1024;
    }
    tmp = c->cub_tmp_cap;
    B200_CUDA(cub::DeviceScan::ExclusiveSum(c->d_cub_tmp, tmp, c->d_keep, c->d_pos, n, c->stream));
    return 0;
}

// best_idx / best_d2 (device) -> ordered accepted list (device); *d_count receives n
int run_accept(b200sift_ctx *c, const int32_t *d_idx, const int32_t *d_d2, int nA, int thresh,
               const b200sift_keypoint *kA, const b200sift_keypoint *kB, int32_t *d_ia, int32_t *d_ib,
               float *d_xyxy, int32_t *d_count)
{
    if (nA <= 0) return 0;
    B200_ARG(nA <= c->raw_cap);
    const int blocks = (nA + 255) / 256;
    accept_flag_kernel<<<blocks, 256, 0, c->stream>>>(d_idx, d_d2, nA, thresh, c->d_keep);
    B200_CHECK(scan_keep(c, nA));
    accept_scatter_kernel<<<blocks, 256, 0, c->stream>>>(d_idx, c->d_keep, c->d_pos, nA, kA, kB, d_ia, d_ib, d_xyxy,
                                                         d_count);
    c->launches += 3;
    B200_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// Nearest / second-nearest ratio test (sift_visualizeUI.py:247-257: knnMatch(k=2), keep m when
// m.distance < 0.7 * n.distance).  The reference gets its two neighbours from approximate FLANN
// KD-trees; here they are the exact ones, and the test is evaluated on the squared integer
// distances: sqrt(d1) < (num/den) sqrt(d2)  <=>  den^2 d1 < num^2 d2  (64-bit, exact).  A row
// without a second neighbour (nB < 2) is rejected.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ratio_flag_kernel(const int32_t *__restrict__ best_idx, const int32_t *__restrict__ d1, const int32_t *__restrict__ d2,
                  int nA, long long num2, long long den2, uint32_t *__restrict__ keep)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= nA) return;
    const bool ok = best_idx[i] != -1 && d2[i] != INT_MAX && den2 * (long long)d1[i] < num2 * (long long)d2[i];
    keep[i] = ok ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
ratio_scatter_kernel(const int32_t *__restrict__ best_idx, const uint32_t *__restrict__ keep,
                     const uint32_t *__restrict__ pos, int nA, int32_t *__restrict__ ia, int32_t *__restrict__ ib,
                     int32_t *__restrict__ count)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= nA) return;
    if (i == nA - 1) *count = (int32_t)(pos[i] + keep[i]);
    if (!keep[i]) return;
    ia[pos[i]] = i;
    ib[pos[i]] = best_idx[i];
}

int run_ratio_accept(b200sift_ctx *c, const int32_t *d_idx, const int32_t *d_d1, const int32_t *d_d2, int nA, int num,
                     int den, int32_t *d_ia, int32_t *d_ib, int32_t *d_count)
{
    if (nA <= 0) return 0;
    B200_CHECK(ensure_sparse_for(c, 1, nA));   // d_keep / d_pos hold nA flags
    const int blocks = (nA + 255) / 256;
    ratio_flag_kernel<<<blocks, 256, 0, c->stream>>>(d_idx, d_d1, d_d2, nA, (long long)num * num, (long long)den * den,
                                                     c->d_keep);
    B200_CHECK(scan_keep(c, nA));
    ratio_scatter_kernel<<<blocks, 256, 0, c->stream>>>(d_idx, c->d_keep, c->d_pos, nA, d_ia, d_ib, d_count);
    c->launches += 3;
    B200_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200
