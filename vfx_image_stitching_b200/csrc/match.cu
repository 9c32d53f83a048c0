// Brute-force descriptor matcher: for every row of A the nearest and second
// nearest row of B under squared L2 distance, exact in integers.
//
// Replaces /root/reference/image_stitching_sift.py:63-73 (the O(NA*NB) python
// double loop with np.dot(d,d) and a strict "<" arg-min).  Descriptors are the
// 0..255 integers generate_descriptors emits (sift_impl.py:519-524), so
// |a-b|^2 = |a|^2 + |b|^2 - 2 a.b <= 128*255^2 < 2^24 is exact in int32 (and
// equals the reference's float32 value bit for bit).
//
// This file holds the CUDA-core (dp4a) formulation: one thread owns one A row
// in registers and streams B tiles through shared memory.  B is split into
// chunks across blockIdx.y so small A sets still fill the GPU; the per-chunk
// top-2 are merged in chunk order (lowest j wins ties) by a second kernel.
#include <cub/device/device_scan.cuh>
#include <limits.h>
#include "common.cuh"

namespace b200 {

constexpr int kMatchRows = 128;   // A rows per CTA (one per thread)
constexpr int kMatchTile = 64;    // B rows staged per iteration
constexpr int kMatchChunk = 512;  // B rows per blockIdx.y

__global__ void __launch_bounds__(256) norms_kernel(const uint8_t *__restrict__ d, int n, int32_t *__restrict__ out)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint4 *p = reinterpret_cast<const uint4 *>(d + (size_t)i * 128);
    unsigned acc = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const uint4 v = p[q];
        acc = __dp4a(v.x, v.x, acc);
        acc = __dp4a(v.y, v.y, acc);
        acc = __dp4a(v.z, v.z, acc);
        acc = __dp4a(v.w, v.w, acc);
    }
    out[i] = (int32_t)acc;
}

__global__ void __launch_bounds__(kMatchRows)
match_dp4a_kernel(const uint8_t *__restrict__ A, int nA, const uint8_t *__restrict__ B, int nB,
                  const int32_t *__restrict__ nrmB, int n_chunks, int32_t *__restrict__ part /*[nA][n_chunks][3]*/)
{
    __shared__ uint4 b_s[kMatchTile * 8];
    __shared__ int32_t nb_s[kMatchTile];
    const int i = blockIdx.x * kMatchRows + threadIdx.x;
    const int chunk = blockIdx.y;
    const int j_begin = chunk * kMatchChunk, j_end = min(nB, j_begin + kMatchChunk);
    uint32_t a[32];
    unsigned na = 0;
    {
        const int ii = min(i, nA - 1);
        const uint4 *p = reinterpret_cast<const uint4 *>(A + (size_t)ii * 128);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint4 v = p[q];
            a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) na = __dp4a(a[q], a[q], na);
    }
    int b1 = INT_MAX, b2 = INT_MAX, bi = -1;
    for (int j0 = j_begin; j0 < j_end; j0 += kMatchTile) {
        const int nt = min(kMatchTile, j_end - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < nt * 8; t += kMatchRows)
            b_s[t] = reinterpret_cast<const uint4 *>(B + (size_t)j0 * 128)[t];
        for (int t = threadIdx.x; t < nt; t += kMatchRows) nb_s[t] = nrmB[j0 + t];
        __syncthreads();
        for (int t = 0; t < nt; ++t) {
            unsigned dot = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint4 v = b_s[t * 8 + q];
                dot = __dp4a(a[4 * q], v.x, dot);
                dot = __dp4a(a[4 * q + 1], v.y, dot);
                dot = __dp4a(a[4 * q + 2], v.z, dot);
                dot = __dp4a(a[4 * q + 3], v.w, dot);
            }
            const int d = (int)na + nb_s[t] - 2 * (int)dot;
            if (d < b1) { b2 = b1; b1 = d; bi = j0 + t; }
            else if (d < b2) b2 = d;
        }
    }
    if (i < nA) {
        int32_t *o = part + ((size_t)i * n_chunks + chunk) * 3;
        o[0] = bi; o[1] = b1; o[2] = b2;
    }
}

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

int run_match(b200sift_ctx *c, const uint8_t *dA, int nA, const uint8_t *dB, int nB, int32_t *d_best_idx,
              int32_t *d_best_d2, int32_t *d_second_d2)
{
    if (nA <= 0) return 0;
    const int n_chunks = nB > 0 ? (nB + kMatchChunk - 1) / kMatchChunk : 1;
    size_t cap = c->nrmB_cap;
    B200_CHECK(ensure(&c->d_nrmB, &cap, (size_t)(nB > 0 ? nB : 1)));
    c->nrmB_cap = cap;
    cap = c->mout_cap;
    B200_CHECK(ensure(&c->d_mout, &cap, (size_t)nA * n_chunks * 3));
    c->mout_cap = cap;
    if (nB > 0) {
        norms_kernel<<<(nB + 255) / 256, 256, 0, c->stream>>>(dB, nB, c->d_nrmB);
        c->launches++;
    }
    dim3 grid((nA + kMatchRows - 1) / kMatchRows, n_chunks);
    match_dp4a_kernel<<<grid, kMatchRows, 0, c->stream>>>(dA, nA, dB, nB, c->d_nrmB, n_chunks, c->d_mout);
    match_merge_kernel<<<(nA + 255) / 256, 256, 0, c->stream>>>(c->d_mout, nA, n_chunks, d_best_idx, d_best_d2,
                                                                d_second_d2);
    c->launches += 2;
    B200_CUDA(cudaGetLastError());
    return 0;
}

// accepted matches of image_stitching_sift.py:74-79, compacted in A order
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

// best_idx / best_d2 (device) -> ordered accepted list (device); *d_count receives n
int run_accept(b200sift_ctx *c, const int32_t *d_idx, const int32_t *d_d2, int nA, int thresh,
               const b200sift_keypoint *kA, const b200sift_keypoint *kB, int32_t *d_ia, int32_t *d_ib,
               float *d_xyxy, int32_t *d_count)
{
    if (nA <= 0) return 0;
    B200_ARG(nA <= c->raw_cap);
    const int blocks = (nA + 255) / 256;
    accept_flag_kernel<<<blocks, 256, 0, c->stream>>>(d_idx, d_d2, nA, thresh, c->d_keep);
    size_t tmp = 0;
    B200_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->d_keep, c->d_pos, nA, c->stream));
    if (tmp > c->cub_tmp_cap) {
        if (c->d_cub_tmp) cudaFree(c->d_cub_tmp);
        c->d_cub_tmp = nullptr;
        B200_CUDA(cudaMalloc(&c->d_cub_tmp, tmp + 1024));
        c->cub_tmp_cap = tmp + 1024;
    }
    tmp = c->cub_tmp_cap;
    B200_CUDA(cub::DeviceScan::ExclusiveSum(c->d_cub_tmp, tmp, c->d_keep, c->d_pos, nA, c->stream));
    accept_scatter_kernel<<<blocks, 256, 0, c->stream>>>(d_idx, c->d_keep, c->d_pos, nA, kA, kB, d_ia, d_ib, d_xyxy,
                                                         d_count);
    c->launches += 3;
    B200_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200
