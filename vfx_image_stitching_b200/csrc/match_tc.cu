// All-pairs descriptor matcher on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Replaces /root/reference/image_stitching_sift.py:63-73 (the python double loop
// `d = descA[i]-descB[j]; dist = np.dot(d,d); if dist < best_dist`).  With the 0..255 integer
// descriptors of sift_impl.py:519-524,
//     |a - b|^2 = |a|^2 + |b|^2 - 2 a.b
// and the contraction a.b over K = 128 is a dense u8 x u8 -> s32 GEMM (tcgen05.mma kind::i8),
// exact in integers, so the result equals the reference's float32 value bit for bit.
//
// Data layout.  A pre-pass (pack_kernel) rewrites every image's (n,128) uint8 descriptors into the
// UMMA canonical K-major "interleave" (no-swizzle) layout, 8-row groups of 1 KiB:
//     [group = row/8][k-chunk = 0..7][row%8][16 bytes]
// so that a 128-row A tile (16 KiB) and a 256-row B tile (32 KiB) are CONTIGUOUS in global memory
// and one cp.async.bulk (TMA bulk copy, SASS UBLKCP) brings a tile into shared memory already in
// the form the matrix descriptors expect (LBO = 128 B between the two 16-byte K chunks of one MMA,
// SBO = 1 KiB between 8-row groups).  Rows are padded with zeros to a multiple of 256 per image;
// the pre-pass also writes |row|^2 (padding rows get a sentinel that can never win).
//
// Kernel (one CTA per 128 A rows x chunk of B tiles, 10 warps, warp-specialised):
//   warp 0   TMA producer : A tile once, then B tiles (+ their 1 KiB of per-column key constants)
//                           through a 4-stage shared-memory ring
//   warp 1   MMA issuer   : one elected thread issues 4 x tcgen05.mma (M128 N256 K32, kind::i8)
//                           per B tile into one of two 256-column TMEM accumulators and
//                           tcgen05.commit's the stage / the accumulator to mbarriers
//   warps 2-9 epilogue    : two warps per TMEM lane quarter, each owning 128 of the 256 columns;
//                           thread <-> A row (TMEM lane).  tcgen05.ld of the next 32 columns is in
//                           flight while the current 32 are folded into the running minimum of
//                               key = (|b_j|^2 - 2 a.b_j) * 256 + (j mod 256)
//                           = one IMAD per element + one three-input minimum (VIMNMX3) per two
//                           elements (nearest) / min + max per element (top-2); the min of the
//                           packed key is the arg-min with the lowest j winning ties, exactly the
//                           strict "<" of the reference loop.  |a|^2 is added once per row at the end.
// The epilogue of tile t overlaps the MMAs of tile t+1 (double-buffered TMEM) and the TMA of
// tiles t+2.. (ring).  K = 128 is only four MMA K-steps (512 tensor-pipe cycles per tile at the
// kind::i8 rate), so the epilogue -- 32 768 accumulators per tile -- has to run at about one
// element per lane-cycle on every scheduler to keep up: see DESIGN.md for the roofline arithmetic.
#include <limits.h>
#include <string.h>
#include "common.cuh"

namespace b200 {

constexpr int kTcM = 128, kTcN = 256, kTcStages = 4;
constexpr int kTcKeySlots = kTcStages + 2;   // key-constant ring: a slot is reused kTcStages + 2 tiles later (see producer)
constexpr int kTcABytes = kTcM * 128, kTcBBytes = kTcN * 128, kTcKeyBytes = kTcN * 4;
constexpr int kTcEpiWarps = 8;
constexpr int kTcThreads = (2 + kTcEpiWarps) * 32;
constexpr int kPadSentinel = 0x7FFFFF;  // |row|^2 of a padding row: key = 0x7FFFFF00 + j > every real key

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
// TMA bulk copy global -> shared, completion on an mbarrier (transaction bytes)
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, u8 x u8 -> s32, M128 x N256 x K32
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld32(uint32_t taddr, int32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no swizzle: start address, LBO (K-chunk stride) = 128 B, SBO (8-row group stride) = 1 KiB,
// descriptor version 1 (sm_100), layout type 0.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           (1ull << 46);
}
// kind::i8: D = S32 (2 << 4), A = B = unsigned 8 bit (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kTcIdesc = (2u << 4) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);

// ------------------------------------------------------------------ pre-pass
// Row-major (n,128) uint8 -> packed canonical layout + squared norms.  grid.y = image.
struct PackImg { int src_off, n, dst_off; };  // rows: source offset, count, packed offset (multiple of 256)

__global__ void __launch_bounds__(256)
pack_kernel(const uint8_t *__restrict__ src, const PackImg *__restrict__ imgs, uint8_t *__restrict__ packed,
            int32_t *__restrict__ nrm, int32_t *__restrict__ ckey)
{
    const PackImg im = imgs[blockIdx.y];
    const int n_pad = (im.n + kTcN - 1) / kTcN * kTcN;
    const int idx = blockIdx.x * 256 + threadIdx.x;  // one 16-byte chunk
    const int r = idx >> 3, kc = idx & 7;
    if (r >= n_pad) return;  // whole 8-lane groups leave together (r is uniform in a group)
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < im.n) v = reinterpret_cast<const uint4 *>(src + (size_t)(im.src_off + r) * 128)[kc];
    const int pr = im.dst_off + r;
    reinterpret_cast<uint4 *>(packed + (size_t)(pr >> 3) * 1024 + kc * 128 + (pr & 7) * 16)[0] = v;
    unsigned s = __dp4a(v.x, v.x, 0u);
    s = __dp4a(v.y, v.y, s);
    s = __dp4a(v.z, v.z, s);
    s = __dp4a(v.w, v.w, s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (kc == 0) {
        const int32_t nr = r < im.n ? (int32_t)s : kPadSentinel;
        nrm[pr] = nr;
        ckey[pr] = nr * 256 + (pr & 255);   // per-column constant of the packed key (images start at multiples of 256)
    }
}

// ------------------------------------------------------------------ the matcher
struct TcPair { int offA, nA, offB, nB, imgA, imgB; };  // packed row offsets (multiples of 256), true counts, image indices

// Device-counted images (RemoteImage): write min(*d_count, cap) over the host's upper bound in the
// pack table and in every pair that names the image.  Runs after the table upload, before pack_kernel.
__global__ void __launch_bounds__(128)
patch_remote_tc_kernel(PackImg *__restrict__ imgs, TcPair *__restrict__ tp, int n_pairs, int image,
                       const int32_t *__restrict__ d_count, int cap)
{
    const int n = max(0, min(*d_count, cap));
    if (threadIdx.x == 0) imgs[image].n = n;
    for (int p = threadIdx.x; p < n_pairs; p += blockDim.x) {
        if (tp[p].imgA == image) tp[p].nA = n;
        if (tp[p].imgB == image) tp[p].nB = n;
    }
}

template <bool kTop2>
__global__ void __launch_bounds__(kTcThreads, 1)
match_tc_kernel(const uint8_t *__restrict__ packed, const int32_t *__restrict__ nrm, const int32_t *__restrict__ ckey,
                const TcPair *__restrict__ pairs, int tiles_per_chunk, int n_chunks, int rows_max,
                int32_t *__restrict__ part)
{
    extern __shared__ __align__(1024) uint8_t tsm[];
    uint8_t *a_s = tsm;                                   // 16 KiB
    uint8_t *b_s = tsm + kTcABytes;                       // kTcStages x 32 KiB
    int32_t *ck_s = reinterpret_cast<int32_t *>(b_s + kTcStages * kTcBBytes);   // [kTcKeySlots][256]
    int32_t *red_s = ck_s + kTcKeySlots * kTcN;                                  // [128][4]: upper column half's result
    uint64_t *bars = reinterpret_cast<uint64_t *>(red_s + kTcM * 4);
    uint64_t *full = bars, *empty = full + kTcStages, *kfull = empty + kTcStages, *tfull = kfull + kTcKeySlots,
             *tempty = tfull + 2, *afull = tempty + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(afull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TcPair P = pairs[blockIdx.z];
    const int m0 = blockIdx.x * kTcM;
    if (m0 >= P.nA) return;
    const int chunk = blockIdx.y;
    const int n_tiles_b = (P.nB + kTcN - 1) / kTcN;
    const int t_begin = chunk * tiles_per_chunk, t_end = min(n_tiles_b, t_begin + tiles_per_chunk);
    const int n_tiles = t_end - t_begin;
    if (n_tiles <= 0) {  // nothing of B in this chunk: sentinel (uniform over the CTA)
        if (warp >= 2 && warp < 6) {
            const int r2 = m0 + (warp & 3) * 32 + lane;
            if (r2 < P.nA) {
                int32_t *o = part + (((size_t)blockIdx.z * rows_max + r2) * n_chunks + chunk) * 3;
                o[0] = -1; o[1] = INT_MAX; o[2] = INT_MAX;
            }
        }
        return;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < kTcKeySlots; ++s) mbar_init(&kfull[s], 1);
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], kTcEpiWarps); }
        mbar_init(afull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: both accumulators = all 512 columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        // B tile t goes to stage t % kTcStages once the MMAs of tile t - kTcStages have retired.  Its key
        // constants go to slot t % kTcKeySlots: the previous user of that slot is tile t - kTcStages - 2,
        // whose epilogue is complete because the MMAs of tile t - kTcStages (just seen retired) could not
        // start before it handed its TMEM buffer back.
        if (lane == 0) {
            mbar_arrive_expect_tx(afull, kTcABytes);
            tma_bulk_g2s(a_s, packed + (size_t)(P.offA + m0) * 128, kTcABytes, afull);
            int s = 0, ks = 0;
            for (int t = 0; t < n_tiles; ++t) {
                mbar_wait(&empty[s], ((t / kTcStages) & 1) ^ 1);
                const size_t row0 = (size_t)P.offB + (size_t)(t_begin + t) * kTcN;
                mbar_arrive_expect_tx(&full[s], kTcBBytes);
                tma_bulk_g2s(b_s + (size_t)s * kTcBBytes, packed + row0 * 128, kTcBBytes, &full[s]);
                mbar_arrive_expect_tx(&kfull[ks], kTcKeyBytes);
                tma_bulk_g2s(ck_s + ks * kTcN, ckey + row0, kTcKeyBytes, &kfull[ks]);
                if (++s == kTcStages) s = 0;
                if (++ks == kTcKeySlots) ks = 0;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            mbar_wait(afull, 0);
            const uint32_t a_addr = smem_u32(a_s);
            int s = 0;
            for (int t = 0; t < n_tiles; ++t) {
                const int buf = t & 1;
                mbar_wait(&tempty[buf], ((t >> 1) & 1) ^ 1);  // epilogue drained this accumulator
                mbar_wait(&full[s], (t / kTcStages) & 1);     // B tile landed
                tc_fence_after();
                const uint32_t b_addr = smem_u32(b_s + (size_t)s * kTcBBytes);
                const uint32_t d = tmem_base + (uint32_t)(buf * kTcN);
#pragma unroll
                for (int k = 0; k < 4; ++k)  // K = 128 bytes = 4 x (two 16-byte chunks)
                    tc_mma_i8(d, make_smem_desc(a_addr + k * 256), make_smem_desc(b_addr + k * 256), kTcIdesc,
                              k > 0 ? 1u : 0u);
                tc_commit(&empty[s]);    // smem stage reusable once these MMAs retire
                tc_commit(&tfull[buf]);  // accumulator ready
                if (++s == kTcStages) s = 0;
            }
        }
    } else {
        // ===== epilogue: thread <-> A row, warp <-> 128 of the tile's 256 columns =====
        const int q = warp & 3;                 // TMEM lane quarter this warp may access (warp index mod 4)
        const int half = (warp - 2) >> 2;       // column half: warps 2-5 take columns 0-127, warps 6-9 columns 128-255
        const int row_l = q * 32 + lane;        // row inside the A tile
        const int r = m0 + row_l;
        int best_d = INT_MAX, best_j = -1, second_d = INT_MAX;
        int ks = 0;
        for (int t = 0; t < n_tiles; ++t) {
            const int buf = t & 1;
            mbar_wait(&kfull[ks], (t / kTcKeySlots) & 1);   // this tile's key constants c_j = |b_j|^2 * 256 + (j mod 256)
            mbar_wait(&tfull[buf], (t >> 1) & 1);
            tc_fence_after();
            int m1 = INT_MAX, m2 = INT_MAX;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kTcN + half * 128);
            const int4 *cj4 = reinterpret_cast<const int4 *>(ck_s + ks * kTcN + half * 128);
            int32_t acc[2][32];
            tc_ld32(taddr, acc[0]);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i < 3) tc_ld32(taddr + 32 * (i + 1), acc[(i + 1) & 1]);   // in flight during the arithmetic below
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const int4 c = cj4[i * 8 + g];
                    const int32_t *a = acc[i & 1] + 4 * g;
                    const int k0 = a[0] * -512 + c.x, k1 = a[1] * -512 + c.y;
                    const int k2 = a[2] * -512 + c.z, k3 = a[3] * -512 + c.w;
                    if (kTop2) {
                        m2 = min(m2, max(m1, k0)); m1 = min(m1, k0);
                        m2 = min(m2, max(m1, k1)); m1 = min(m1, k1);
                        m2 = min(m2, max(m1, k2)); m1 = min(m1, k2);
                        m2 = min(m2, max(m1, k3)); m1 = min(m1, k3);
                    } else {
                        m1 = min(m1, min(k0, k1));
                        m1 = min(m1, min(k2, k3));
                    }
                }
                if (i < 3) tc_wait_ld();
            }
            // accumulator drained: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
            // merge the tile-local result (tiles ascend in j: strict < keeps the lowest j)
            const int td = m1 >> 8, tj = (t_begin + t) * kTcN + (m1 & 255);
            if (td < best_d) {
                if (kTop2) second_d = min(best_d, m2 >> 8);
                best_d = td;
                best_j = tj;
            } else if (kTop2) {
                second_d = min(second_d, td);
            }
            if (++ks == kTcKeySlots) ks = 0;
        }
        // the two column halves of a row meet in shared memory; equal distances: the lower j wins
        if (half == 1) {
            red_s[row_l * 4] = best_d; red_s[row_l * 4 + 1] = best_j; red_s[row_l * 4 + 2] = second_d;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiWarps * 32) : "memory");
        if (half == 0 && r < P.nA) {
            const int od = red_s[row_l * 4], oj = red_s[row_l * 4 + 1], os = red_s[row_l * 4 + 2];
            if (od < best_d || (od == best_d && (unsigned)oj < (unsigned)best_j)) {
                if (kTop2) second_d = min(best_d, os);
                best_d = od;
                best_j = oj;
            } else if (kTop2) {
                second_d = min(second_d, od);
            }
            const int na = nrm[P.offA + r];
            int32_t *o = part + (((size_t)blockIdx.z * rows_max + r) * n_chunks + chunk) * 3;
            const bool real = best_j >= 0 && best_j < P.nB;  // padding rows of B can only win when nB == 0
            o[0] = real ? best_j : -1;
            o[1] = real ? best_d + na : INT_MAX;
            o[2] = (kTop2 && real && second_d < (kPadSentinel - 1)) ? second_d + na : INT_MAX;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

constexpr size_t kTcSmemBytes = kTcABytes + (size_t)kTcStages * kTcBBytes + (size_t)kTcKeySlots * kTcKeyBytes +
                                kTcM * 4 * 4 + (2 * kTcStages + kTcKeySlots + 5) * 8 + 64;

// Function attributes are per DEVICE: called once for every device a context is created on
// (b200sift_create, under the init lock), never from a launch path.
int match_init_device()
{
    B200_CUDA(cudaFuncSetAttribute(match_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)kTcSmemBytes));
    B200_CUDA(cudaFuncSetAttribute(match_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)kTcSmemBytes));
    return 0;
}

// Pack `n_imgs` descriptor sets (rows of c->d_desc or of an explicit source) and run every pair.
// part layout = the one pair_finalize_kernel / match_merge_kernel read.
int run_match_tc(b200sift_ctx *c, const uint8_t *d_src, int n_imgs, const int *h_src_off, const int *h_n,
                 int n_pairs, const int *h_pairs /*2 per pair: image indices*/, int rows_max, int n_chunks_out,
                 int tiles_per_chunk, bool top2, int32_t *d_part)
{
    // packed offsets
    std::vector<PackImg> imgs(n_imgs);
    int total = 0, max_pad = 0;
    for (int i = 0; i < n_imgs; ++i) {
        imgs[i].src_off = h_src_off[i];
        imgs[i].n = h_n[i];
        imgs[i].dst_off = total;
        const int pad = (h_n[i] + kTcN - 1) / kTcN * kTcN;
        total += pad;
        if (pad > max_pad) max_pad = pad;
    }
    if (total == 0 || n_pairs == 0) return 0;
    std::vector<TcPair> tp(n_pairs);
    for (int p = 0; p < n_pairs; ++p) {
        const int a = h_pairs[2 * p], b = h_pairs[2 * p + 1];
        tp[p].offA = imgs[a].dst_off; tp[p].nA = imgs[a].n; tp[p].imgA = a;
        tp[p].offB = imgs[b].dst_off; tp[p].nB = imgs[b].n; tp[p].imgB = b;
    }
    size_t cap = c->tc_cap;
    const size_t need = (size_t)total * 128 + (size_t)total * 8 + imgs.size() * sizeof(PackImg) +
                        tp.size() * sizeof(TcPair) + 1024;
    B200_CHECK(ensure(&c->d_tc, &cap, need));
    c->tc_cap = cap;
    uint8_t *packed = c->d_tc;
    int32_t *nrm = reinterpret_cast<int32_t *>(packed + (size_t)total * 128);
    int32_t *ckey = nrm + total;   // total is a multiple of 256: 1 KiB aligned like the TMA slices need
    PackImg *d_imgs = reinterpret_cast<PackImg *>(ckey + total);
    TcPair *d_tp = reinterpret_cast<TcPair *>(d_imgs + imgs.size());
    // the small host tables must outlive the asynchronous copies: keep them in the context
    c->h_tc_tables.resize(imgs.size() * sizeof(PackImg) + tp.size() * sizeof(TcPair));
    memcpy(c->h_tc_tables.data(), imgs.data(), imgs.size() * sizeof(PackImg));
    memcpy(c->h_tc_tables.data() + imgs.size() * sizeof(PackImg), tp.data(), tp.size() * sizeof(TcPair));
    B200_CUDA(cudaMemcpyAsync(d_imgs, c->h_tc_tables.data(), c->h_tc_tables.size(), cudaMemcpyHostToDevice,
                              c->stream));  // d_imgs and d_tp are adjacent
    for (const RemoteImage &r : c->remote)
        if (d_src == c->d_desc && r.image < n_imgs) {
            patch_remote_tc_kernel<<<1, 128, 0, c->stream>>>(d_imgs, d_tp, n_pairs, r.image, r.d_count, r.cap);
            c->launches++;
        }
    if (max_pad > 0) {
        dim3 pg((max_pad * 8 + 255) / 256, n_imgs);
        pack_kernel<<<pg, 256, 0, c->stream>>>(d_src, d_imgs, packed, nrm, ckey);
        c->launches++;
    }
    c->last_tiles_per_chunk = tiles_per_chunk;
    c->last_n_chunks = n_chunks_out;
    dim3 grid((rows_max + kTcM - 1) / kTcM, n_chunks_out, n_pairs);
    if (top2)
        match_tc_kernel<true><<<grid, kTcThreads, kTcSmemBytes, c->stream>>>(packed, nrm, ckey, d_tp, tiles_per_chunk,
                                                                           n_chunks_out, rows_max, d_part);
    else
        match_tc_kernel<false><<<grid, kTcThreads, kTcSmemBytes, c->stream>>>(packed, nrm, ckey, d_tp, tiles_per_chunk,
                                                                            n_chunks_out, rows_max, d_part);
    c->launches++;
    B200_CUDA(cudaGetLastError());
    return 0;
}

void tc_chunking(const b200sift_ctx *c, int rows_max, int nb_max, int n_pairs, int *tiles_per_chunk, int *n_chunks);

// ------------------------------------------------------------------ measurement hook
__global__ void __launch_bounds__(256) fill_desc_kernel(uint8_t *d, size_t n_bytes, uint32_t seed)
{
    // SIFT-like synthetic descriptors: mostly small values, clipped at 255 (xorshift hash per 4 bytes)
    size_t i = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4;
    const size_t stride = (size_t)gridDim.x * 256 * 4;
    for (; i + 3 < n_bytes; i += stride) {
        uint32_t x = (uint32_t)(i >> 2) * 2654435761u + seed;
        x ^= x << 13; x ^= x >> 17; x ^= x << 5;
        uint32_t out = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t r = (x >> (8 * b)) & 255u;
            const uint32_t v = (r * r * r) >> 18;  // 0..63 skewed, like quantised SIFT bins
            out |= (v > 255u ? 255u : v) << (8 * b);
        }
        *reinterpret_cast<uint32_t *>(d + i) = out;
    }
}

// Times the tensor-core kernel alone (CUDA events on the context stream) on synthetic nA x nB
// descriptors resident in HBM; `ms_kernel` = mean per launch.  Algorithmic work of one launch:
// 2*128*nA*nB integer operations.
int bench_match_tc(b200sift_ctx *c, const uint8_t *hA, int nA, const uint8_t *hB, int nB, int top2, int iters,
                   float *ms_kernel)
{
    size_t cap = c->tcsrc_cap;
    B200_CHECK(ensure(&c->d_tcsrc, &cap, (size_t)(nA + nB) * 128));
    c->tcsrc_cap = cap;
    if (hA && hB) {   // caller-supplied descriptors (host): the distributions of BASELINE.json configs[4]
        B200_CUDA(cudaMemcpyAsync(c->d_tcsrc, hA, (size_t)nA * 128, cudaMemcpyHostToDevice, c->stream));
        B200_CUDA(cudaMemcpyAsync(c->d_tcsrc + (size_t)nA * 128, hB, (size_t)nB * 128, cudaMemcpyHostToDevice, c->stream));
    } else {
        fill_desc_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->d_tcsrc, (size_t)(nA + nB) * 128, 12345u);
    }
    int tpc, n_chunks;
    tc_chunking(c, nA, nB, 1, &tpc, &n_chunks);
    cap = c->mout_cap;
    B200_CHECK(ensure(&c->d_mout, &cap, (size_t)nA * n_chunks * 3));
    c->mout_cap = cap;
    const int src_off[2] = {0, nA}, ns[2] = {nA, nB}, pr[2] = {0, 1};
    // first call packs + warms up
    B200_CHECK(run_match_tc(c, c->d_tcsrc, 2, src_off, ns, 1, pr, nA, n_chunks, tpc, top2 != 0, c->d_mout));
    const int padA = (nA + kTcN - 1) / kTcN * kTcN, padB = (nB + kTcN - 1) / kTcN * kTcN;
    uint8_t *packed = c->d_tc;
    int32_t *nrm = reinterpret_cast<int32_t *>(packed + (size_t)(padA + padB) * 128);
    int32_t *ckey = nrm + padA + padB;
    TcPair *d_tp = reinterpret_cast<TcPair *>(reinterpret_cast<PackImg *>(ckey + padA + padB) + 2);
    dim3 grid((nA + kTcM - 1) / kTcM, n_chunks, 1);
    double acc = 0;
    for (int it = 0; it < iters + 2; ++it) {
        B200_CUDA(cudaEventRecord(c->ev0, c->stream));
        if (top2)
            match_tc_kernel<true><<<grid, kTcThreads, kTcSmemBytes, c->stream>>>(packed, nrm, ckey, d_tp, tpc, n_chunks, nA,
                                                                               c->d_mout);
        else
            match_tc_kernel<false><<<grid, kTcThreads, kTcSmemBytes, c->stream>>>(packed, nrm, ckey, d_tp, tpc, n_chunks, nA,
                                                                                c->d_mout);
        B200_CUDA(cudaEventRecord(c->ev1, c->stream));
        B200_CUDA(cudaEventSynchronize(c->ev1));
        float ms = 0;
        B200_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (it >= 2) acc += ms;
        c->launches++;
    }
    B200_CUDA(cudaGetLastError());
    *ms_kernel = (float)(acc / iters);
    return 0;
}

// Chunking shared by both callers.  One CTA is resident per SM (it owns all 512 TMEM columns), so
// the launch runs in waves of sm_count CTAs; a CTA costs its B tiles plus a fixed start-up (TMEM
// allocation, A tile, first B tile, pipeline fill: about two tile times).  The number of B chunks is
// the one that minimises waves x (tiles per chunk + start-up): small problems are cut until they fill
// the machine, large ones until the last wave is nearly full (64k x 64k: 512 A tiles x 2 chunks =
// 6.9 waves instead of 3.5).
void tc_chunking(const b200sift_ctx *c, int rows_max, int nb_max, int n_pairs, int *tiles_per_chunk, int *n_chunks)
{
    const int a_tiles = (rows_max + kTcM - 1) / kTcM;
    const int b_tiles = nb_max > 0 ? (nb_max + kTcN - 1) / kTcN : 1;
    const long long ctas_one = (long long)a_tiles * n_pairs;  // with a single chunk
    long long best_cost = -1;
    int best_tpc = b_tiles;
    const int max_chunks = b_tiles < 64 ? b_tiles : 64;
    for (int ch = 1; ch <= max_chunks; ++ch) {
        const int tpc = (b_tiles + ch - 1) / ch;
        const int real_ch = (b_tiles + tpc - 1) / tpc;
        const long long waves = (ctas_one * real_ch + c->sm_count - 1) / c->sm_count;
        const long long cost = waves * (tpc + 2);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_tpc = tpc; }
    }
    *tiles_per_chunk = best_tpc;
    *n_chunks = (b_tiles + best_tpc - 1) / best_tpc;
}

}  // namespace b200
