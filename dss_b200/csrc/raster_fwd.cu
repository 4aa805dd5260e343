// raster_fwd.cu -- per-tile forward rasterization (+ optional fused blend).
//
// Replaces RasterizePointsFineCudaKernel (DSS/csrc/rasterize_points.cu:506-597), whose every pixel
// loops over all M = max(1e4, P) slots of its bin and keeps a 150-entry queue in local memory, and --
// when blending -- the weights/compositor/concat sequence of DSS/core/renderer.py:53-78.
//
// One 256-thread CTA per 16x16 pixel tile.  The tile's splat list (CSR ids) is streamed through shared
// memory in chunks of 256 records.  Each warp owns an 8x4 pixel patch and runs a two-level test:
//   level 1 (lane-parallel, 32 splats at a time): bbox-vs-patch overlap and a depth cull against the
//            warp's current K-th nearest depth -> ballot;
//   level 2 (per surviving splat, broadcast from shared memory): the reference's exact per-pixel test
//            (CheckPixelInsidePoint, rasterize_points.cu:87-97) and a sorted insert into K registers.
// Selection is by (z, id) lexicographic order, so the result does not depend on list order (the
// reference's own CPU path pops a max-heap of (z, idx, q) tuples: rasterize_points_cpu.cpp:87-121).
#include <math_constants.h>

#include "common.cuh"
#include "kernels.cuh"

namespace dss {

constexpr int RASTER_THREADS = RASTER_TILE * RASTER_TILE;  // 256
constexpr int RASTER_CHUNK = RASTER_THREADS;

__device__ __forceinline__ bool frag_less(float za, int ia, float zb, int ib) {
    return (za < zb) || (za == zb && ia < ib);
}

template <int KMAX, bool PER_POINT_CUTOFF, bool BLEND>
__global__ void __launch_bounds__(RASTER_THREADS)
raster_fwd_kernel(const __grid_constant__ RasterArgs a) {
    __shared__ float4 sA[RASTER_CHUNK];
    __shared__ float4 sB[RASTER_CHUNK];
    __shared__ int sId[RASTER_CHUNK];
    __shared__ float sCut[PER_POINT_CUTOFF ? RASTER_CHUNK : 1];

    const int S = a.S, B = a.B, K = a.K;
    const int n = blockIdx.y;
    const int tile = blockIdx.x;
    const int ty = tile / B, tx = tile - ty * B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr unsigned FULL = 0xffffffffu;

    // NDC-index-space pixel of this thread
    const int ix0 = tx * RASTER_TILE + (warp & 1) * 8;
    const int iy0 = ty * RASTER_TILE + (warp >> 1) * 4;
    const int xi = ix0 + (lane & 7);
    const int yi = iy0 + (lane >> 3);
    const bool valid = (xi < S) && (yi < S);
    const float xf = pix_to_ndc(xi, S);
    const float yf = pix_to_ndc(yi, S);
    // warp patch extents, widened by half a pixel (same slack as the reference's bins)
    const float half_pix = 1.0f / S;
    const bool patch_live = (ix0 < S) && (iy0 < S);
    const float pat_x0 = pix_to_ndc(ix0, S) - half_pix;
    const float pat_x1 = pix_to_ndc(min(ix0 + 7, S - 1), S) + half_pix;
    const float pat_y0 = pix_to_ndc(iy0, S) - half_pix;
    const float pat_y1 = pix_to_ndc(min(iy0 + 3, S - 1), S) + half_pix;

    float fz[KMAX], fq[KMAX];
    int fid[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        fz[k] = CUDART_INF_F;
        fq[k] = -1.0f;
        fid[k] = INT32_MAX;
    }
    float thresh = CUDART_INF_F;  // warp-wide max of the KMAX-th nearest depth

    const int64_t tbase = ((int64_t)n * B * B + tile) * a.NS;
    const int beg = min(a.tile_offsets[tbase], a.ids_capacity), end = min(a.tile_offsets[tbase + a.NS], a.ids_capacity);

    for (int base = beg; base < end; base += RASTER_CHUNK) {
        const int cnt = min(RASTER_CHUNK, end - base);
        if (tid < cnt) {
            const int id = (int)a.tile_keys[base + tid].y;
            sA[tid] = __ldg(&a.rec[2 * (int64_t)id]);
            sB[tid] = __ldg(&a.rec[2 * (int64_t)id + 1]);
            sId[tid] = id;
            if (PER_POINT_CUTOFF) sCut[tid] = __ldg(&a.cutoff[id]);
        }
        __syncthreads();
        if (patch_live) {
            for (int j0 = 0; j0 < cnt; j0 += 32) {
                const int j = j0 + lane;
                bool pass = false;
                if (j < cnt) {
                    const float4 A = sA[j];
                    const float ry = sB[j].x;
                    pass = (A.z >= 0.0f) && (A.x - A.w <= pat_x1) && (pat_x0 <= A.x + A.w) &&
                           (A.y - ry <= pat_y1) && (pat_y0 <= A.y + ry) && (A.z <= thresh);
                }
                unsigned mask = __ballot_sync(FULL, pass);
                bool inserted = false;
                while (mask) {
                    const int jj = j0 + __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float4 A = sA[jj];
                    const float4 Bv = sB[jj];
                    const float dx = xf - A.x;
                    const float dy = yf - A.y;
                    // rasterize_points.cu:92-97 -- same expression tree for q as the reference
                    const float qv = Bv.y * dx * dx + Bv.z * dx * dy + Bv.w * dy * dy;
                    const float cut = PER_POINT_CUTOFF ? sCut[jj] : a.cutoff_uniform;
                    const bool ok = valid && !(fabsf(dx) > A.w || fabsf(dy) > Bv.x) && !(qv > cut);
                    if (ok) {
                        const int id = sId[jj];
                        const float z = A.z + 0.0f;  // canonicalise -0
                        if (frag_less(z, id, fz[KMAX - 1], fid[KMAX - 1])) {
                            inserted = true;
#pragma unroll
                            for (int i = KMAX - 1; i >= 0; --i) {
                                if (frag_less(z, id, fz[i], fid[i])) {
                                    const bool shift = (i > 0) && frag_less(z, id, fz[i > 0 ? i - 1 : 0],
                                                                             fid[i > 0 ? i - 1 : 0]);
                                    fz[i] = shift ? fz[i > 0 ? i - 1 : 0] : z;
                                    fq[i] = shift ? fq[i > 0 ? i - 1 : 0] : qv;
                                    fid[i] = shift ? fid[i > 0 ? i - 1 : 0] : id;
                                }
                            }
                        }
                    }
                }
                if (__any_sync(FULL, inserted)) {
                    const float kz = valid ? fz[KMAX - 1] : -CUDART_INF_F;
                    thresh = __int_as_float(__reduce_max_sync(FULL, __float_as_int(kz)));
                }
            }
        }
        __syncthreads();
    }

    if (!valid) return;
    // ---- epilogue: depth merge, outputs (image row S-1-yi, col S-1-xi: rasterize_points.cu:577-580) ----
    const int64_t pix = ((int64_t)n * S + (S - 1 - yi)) * S + (S - 1 - xi);
    const bool any = fid[0] != INT32_MAX;
    const float z0 = fz[0];
    bool emit[KMAX];
    bool open = any;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        // rasterize_points.cu:586-595: stop at the first fragment farther than the merge threshold
        open = open && (k < K) && (fid[k] != INT32_MAX) && !(fz[k] - z0 > a.depth_merge);
        emit[k] = open;
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
            a.idx[pix * K + k] = emit[k] ? fid[k] : -1;
            if (a.zbuf) a.zbuf[pix * K + k] = emit[k] ? fz[k] : -1.0f;
            if (a.qvalue) a.qvalue[pix * K + k] = emit[k] ? fq[k] : -1.0f;
        }
    }
    const float occ = any ? 1.0f : 0.0f;
    if (a.occ) a.occ[pix] = occ;
    if (BLEND) {
        // renderer.py:53 w = exp(-0.5 q) * scaler ; norm_weighted_sum [ext]: sum w f / max(sum w, 1e-4)
        float w[KMAX];
        float wsum = 0.f, r = 0.f, g = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            w[k] = 0.f;
            if (emit[k]) {
                const int id = fid[k];
                w[k] = expf(-0.5f * fq[k]) * __ldg(&a.scaler[id]);
                const float *c = a.colours + ((int64_t)id - (int64_t)n * a.colour_P0) * 3;
                r += w[k] * __ldg(c + 0);
                g += w[k] * __ldg(c + 1);
                b += w[k] * __ldg(c + 2);
                wsum += w[k];
                if (a.visible) a.visible[id] = 1;
            }
        }
        const float inv = 1.0f / fmaxf(wsum, 1e-4f);
        reinterpret_cast<float4 *>(a.image)[pix] = make_float4(r * inv, g * inv, b * inv, occ);
        if (a.weights) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k)
                if (k < K) a.weights[pix * K + k] = w[k] * inv;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Splat-parallel ("scatter") rasterization: shared-memory K-lists (used by raster_sliced_kernel below).
//
// The pixel-parallel kernel above tests every splat that overlaps a warp's 8x4 patch in all 32 lanes,
// although a ~5 px splat covers only ~1/5 of them.  Here the roles are swapped: the per-pixel K-nearest
// lists of a 16x16 tile live in SHARED memory as sorted 64-bit keys (z bits << 32 | id), every THREAD
// takes its own splat of the tile list and walks only the pixels of that splat's bounding box inside
// the tile:  key >= pixel's current K-th key -> skip (one LDS.64 + compare);  otherwise the reference's
// exact test (rasterize_points.cu:87-97) and a lock-free sorted insert: a chain of atomicMin on the K
// slots, each step keeping the smaller key and carrying the larger one to the next slot.  The slots end
// up holding the K smallest keys in ascending (z, id) order whatever the interleaving, so the result is
// deterministic and identical to the pixel-parallel kernel.  q is recomputed for the K winners in the
// epilogue from the same record with the same expression, hence bit-identical to the accept test.
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long KEY_EMPTY = ~0ull;

__device__ __forceinline__ unsigned long long make_key(float z, int id) {
    return ((unsigned long long)__float_as_uint(z) << 32) | (unsigned int)id;
}

template <int KMAX, bool BLEND>
__device__ __forceinline__ void raster_epilogue(const RasterArgs &a, const unsigned long long *s_keys, bool tile_has_entries,
                                                int n, int tx0, int ty0, bool pow2, float inv_S) {
    const int S = a.S, K = a.K;
    const int tid = threadIdx.x;
    const int beg = 0, end = tile_has_entries ? 1 : 0;
    // one thread per pixel
    const int xi = tx0 + (tid & (RASTER_TILE - 1));
    const int yi = ty0 + (tid >> 4);
    if (xi >= S || yi >= S) return;
    const int64_t pix = ((int64_t)n * S + (S - 1 - yi)) * S + (S - 1 - xi);
    const float xf = pix_to_ndc_fast(xi, S, inv_S, pow2);
    const float yf = pix_to_ndc_fast(yi, S, inv_S, pow2);
    float fz[KMAX], fq[KMAX];
    int fid[KMAX];
    bool emit[KMAX];
    bool open = beg < end;
    float z0 = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        fz[k] = -1.0f;
        fq[k] = -1.0f;
        fid[k] = -1;
        emit[k] = false;
        if (open && k < K) {
            const unsigned long long key = s_keys[(tid * KMAX) + k];
            if (key == KEY_EMPTY) {
                open = false;
            } else {
                const float z = __uint_as_float((unsigned int)(key >> 32));
                if (k == 0) z0 = z;
                if (z - z0 > a.depth_merge) {   // rasterize_points.cu:586-595
                    open = false;
                } else {
                    const int id = (int)(unsigned int)(key & 0xffffffffull);
                    const float4 A = __ldg(&a.rec[2 * (int64_t)id]);
                    const float4 Bv = __ldg(&a.rec[2 * (int64_t)id + 1]);
                    const float dx = xf - A.x, dy = yf - A.y;
                    fz[k] = z;
                    fid[k] = id;
                    fq[k] = Bv.y * dx * dx + Bv.z * dx * dy + Bv.w * dy * dy;
                    emit[k] = true;
                }
            }
        }
    }
    const bool any = (beg < end) && (s_keys[tid * KMAX] != KEY_EMPTY);
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
            a.idx[pix * K + k] = fid[k];
            if (a.zbuf) a.zbuf[pix * K + k] = fz[k];
            if (a.qvalue) a.qvalue[pix * K + k] = fq[k];
        }
    }
    const float occ = any ? 1.0f : 0.0f;
    if (a.occ) a.occ[pix] = occ;
    if (BLEND) {
        float w[KMAX];
        float wsum = 0.f, r = 0.f, g = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            w[k] = 0.f;
            if (emit[k]) {
                const int id = fid[k];
                w[k] = expf(-0.5f * fq[k]) * __ldg(&a.scaler[id]);
                const float *c = a.colours + ((int64_t)id - (int64_t)n * a.colour_P0) * 3;
                r += w[k] * __ldg(c + 0);
                g += w[k] * __ldg(c + 1);
                b += w[k] * __ldg(c + 2);
                wsum += w[k];
                if (a.visible) a.visible[id] = 1;
            }
        }
        const float inv = 1.0f / fmaxf(wsum, 1e-4f);
        reinterpret_cast<float4 *>(a.image)[pix] = make_float4(r * inv, g * inv, b * inv, occ);
        if (a.weights) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k)
                if (k < K) a.weights[pix * K + k] = w[k] * inv;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Depth-sliced splat-parallel rasterizer (production path for K <= 8).
//
// Shared-memory K-lists as described above, plus the three things that remove most of the
// (splat, pixel) candidate tests on dense clouds:
//   1. the tile's list arrives in NS front-to-back depth slices (binning key = tile x slice); before a slice
//      is touched the kernel compares its lower depth bound with the largest K-th depth of the tile and
//      stops if nothing behind that bound can enter any pixel's list;
//   2. every entry is first tested against the K-th depths of the 4x4-pixel blocks its bbox touches (16
//      values per tile, refreshed after every processing phase) -- one cheap, fully lane-parallel test;
//   3. survivors are compacted into a shared-memory queue and rasterized in full batches, so the expensive
//      per-pixel loops run with (nearly) all lanes busy instead of ~18 % (ncu, profiles/r01_ncu_v2_*).
// ---------------------------------------------------------------------------------------------

// Exact range of pixel indices i in [lo_clip, hi_clip] with |pix_to_ndc(i) - c| <= r, i.e. the pixels that
// pass the reference's bbox test (rasterize_points.cu:92) for one axis: arithmetic estimate, then the
// estimate is corrected by evaluating the very same fp32 test on the neighbouring pixels.
__device__ __forceinline__ void pixel_range(float c, float r, int lo_clip, int hi_clip, int S, float inv_S, bool pow2,
                                            float half_S, int &lo, int &hi) {
    if (!(fabsf(c) < 4.0f) || !(r < 2.0f)) {   // far-away centre / huge splat: estimates lose accuracy, test every pixel
        lo = lo_clip;
        hi = hi_clip;
        return;
    }
    lo = (int)fminf(fmaxf(ceilf((c - r + 1.0f) * half_S - 0.5f), -1.0f), (float)S);
    hi = (int)fmaxf(fminf(floorf((c + r + 1.0f) * half_S - 0.5f), (float)S), -2.0f);
    if (lo > 0 && !(fabsf(pix_to_ndc_fast(lo - 1, S, inv_S, pow2) - c) > r)) --lo;
    else if (lo <= hi && lo < S && (fabsf(pix_to_ndc_fast(max(lo, 0), S, inv_S, pow2) - c) > r)) ++lo;
    if (hi < S - 1 && !(fabsf(pix_to_ndc_fast(hi + 1, S, inv_S, pow2) - c) > r)) ++hi;
    else if (hi >= lo && hi >= 0 && (fabsf(pix_to_ndc_fast(min(hi, S - 1), S, inv_S, pow2) - c) > r)) --hi;
    lo = max(lo, lo_clip);
    hi = min(hi, hi_clip);
}

// explicit shared-space accesses with 32-bit addresses: inside the hot pixel loop the compiler otherwise rebuilds the
// generic->shared window base (S2R CgaCtaId + LEA ...) for every access, a third of the loop's instructions
// (the mov through an opaque asm keeps the value in a register instead of being rematerialised in the loop)
__device__ __forceinline__ uint32_t sh_addr(const void *p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ unsigned long long lds_u64_volatile(uint32_t addr) {
    unsigned long long v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int lds_u32(uint32_t addr) {
    unsigned int v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u64(uint32_t addr, unsigned long long v) {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t addr, unsigned short v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}

constexpr int RASTER_QCAP = 384;     // survivors of the entry-level cull waiting to be rasterized (40 bytes each)
constexpr int RASTER_WPEND = 128;    // accepted fragments a warp buffers before it inserts them
constexpr int RASTER_NBKT = 1024;    // depth buckets per slice of a tile's list (in-kernel counting sort)
constexpr int RASTER_SORT_CAP = 4096;   // entries ordered at once (shared memory)
constexpr int RASTER_SORT_MIN = 384;    // shorter (sub)lists are walked as they come
constexpr int RASTER_FLUSH_MIN = 96;    // survivors that trigger a rasterization phase while walking an ordered list

// lock-free sorted insert of one key into a pixel's K slots (chain of atomicMin, see above)
template <int KMAX>
__device__ __forceinline__ void klist_insert(unsigned long long *slot, unsigned long long carry) {
    // most buffered fragments are stale by the time they are inserted (the pixel's K-th key has dropped below them
    // since the test): one load settles those
    if (carry >= slot[KMAX - 1]) return;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        if (carry < slot[k]) {
            const unsigned long long old = atomicMin(&slot[k], carry);
            carry = old > carry ? old : carry;
        }
    }
}

// A warp inserts the fragments it has buffered: 32 at a time, all lanes busy.  Other warps may be testing
// pixels concurrently; they read the K-th key, which only ever decreases, so a stale value merely lets a
// superfluous fragment into a buffer -- the insert itself re-checks.
template <int KMAX>
__device__ __forceinline__ void pend_flush(unsigned long long *s_keys, const unsigned long long *wkey,
                                           const unsigned short *wpix, int count, int lane) {
    __syncwarp();
    for (int i = lane; i < count; i += 32) klist_insert<KMAX>(s_keys + (int)wpix[i] * KMAX, wkey[i]);
    __syncwarp();
}

// Launch order of the tiles: longest lists first (a counting sort of the tiles by the power of two of their list
// length, one block).  CTAs are handed out in block-index order, so without this the last wave of a launch is whatever
// tiles happen to come last -- often dense ones, with most SMs already idle.
__global__ void __launch_bounds__(1024)
raster_tile_order_kernel(const int32_t *__restrict__ tile_offsets, int NS, int ntiles, int32_t *__restrict__ order) {
    __shared__ int s_cnt[32], s_base[32];
    if (threadIdx.x < 32) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < ntiles; t += blockDim.x) {
        const int len = tile_offsets[(int64_t)(t + 1) * NS] - tile_offsets[(int64_t)t * NS];
        atomicAdd(&s_cnt[len > 0 ? 31 - __clz(len) + 1 : 0], 1);      // bucket 0: empty, k: 2^(k-1) <= len < 2^k
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int b = 31; b >= 0; --b) {   // descending work
            s_base[b] = run;
            run += s_cnt[b];
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < ntiles; t += blockDim.x) {
        const int len = tile_offsets[(int64_t)(t + 1) * NS] - tile_offsets[(int64_t)t * NS];
        order[atomicAdd(&s_base[len > 0 ? 31 - __clz(len) + 1 : 0], 1)] = t;
    }
}

// ---------------------------------------------------------------------------------------------
// Depth-ordered splat-parallel rasterizer (production path for K <= 8).
//
// One CTA per 16x16 tile, shared-memory K-lists as described above.  What removes most of the (splat, pixel)
// candidate tests on dense clouds (~100 candidates per covered pixel at 1M points / 512^2, of which the K = 5 nearest
// are wanted) is the ORDER in which a tile's candidates are visited:
//   1. the binning delivers a tile's list as 8-byte {z bits, id} keys in NS coarse front-to-back depth slices;
//   2. inside a slice the CTA orders the keys by a 1024-bucket counting sort on z in shared memory (histogram pass,
//      scan, scatter pass; only the ids are kept) -- no splat record is touched for this;
//   3. the ordered ids are walked 256 at a time; before each group its depth lower bound is compared with the tile's
//      largest K-th depth and the whole rest of the list is dropped as soon as nothing behind the bound can enter any
//      pixel's list -- on a surface seen from the front that happens after the first few hundred of several thousand
//      entries;
//   4. every visited entry is first tested against the K-th depths of the 4x4-pixel blocks its bounding box touches
//      (16 values per tile, refreshed after every rasterization phase); survivors are queued and rasterized one per
//      thread over their exact pixel rectangle, accepted fragments go to per-warp buffers and are inserted 32 at a
//      time.
// The result does not depend on the visiting order (the K-lists keep the K smallest (z, id) keys whatever the order),
// so the ordering needs no exactness: the bucket bounds only have to be conservative.
//
// A tile whose list did not fit the key buffer (the buffer is sized from the PREVIOUS call, see bin_and_raster) takes
// its candidates from the view's records directly -- slow, correct, and it keeps the host out of the step.
// ---------------------------------------------------------------------------------------------
template <int KMAX>
struct RasterShared {
    unsigned long long keys[RASTER_THREADS * KMAX];                  // [pixel][k]
    unsigned long long pend_key[RASTER_THREADS / 32][RASTER_WPEND];  // per warp: accepted fragments waiting for insertion
    float4 qA[RASTER_QCAP];                                          // survivor queue: record halves ...
    float4 qB[RASTER_QCAP];
    uint2 qD[RASTER_QCAP];                                           // ... {id, exact pixel rectangle (tile-local, 4 x 4 bits)}
    unsigned int sorted[RASTER_SORT_CAP];                            // ids of the current round, ordered by bucket
    unsigned int off[RASTER_NBKT + 4];                               // bucket offsets (exclusive scan of the histogram)
    unsigned int grp_lb[RASTER_SORT_CAP / RASTER_THREADS + 1];       // depth lower bound (float bits) of every group
    unsigned short pend_pix[RASTER_THREADS / 32][RASTER_WPEND];
    float xf[RASTER_TILE], yf[RASTER_TILE];                          // exact pixel centres of the tile
    unsigned int closed[RASTER_TILE];   // per pixel row: bit x set = no entry still to come can enter that pixel's list
    unsigned int warp_tot[RASTER_THREADS / 32];
    int qcount;
    int bhi;
};
static_assert(sizeof(float4) * RASTER_QCAP >= sizeof(int) * RASTER_NBKT, "the scatter cursors alias the survivor queue");

template <int KMAX, bool PER_POINT_CUTOFF, bool STATS, bool IMM>
struct RasterTile {
    const RasterArgs &a;
    RasterShared<KMAX> &sh;
    int tx0, ty0, tx1, ty1, tid, lane, warp;
    bool pow2;
    float inv_S, half_S;
    uint32_t sa_kth, sa_xf, sa_yf, sa_wkey, sa_wpix, sa_closed;
    int nq;                  // survivors queued (uniform over the CTA)
    unsigned int tilemax;    // largest K-th depth of the tile's pixels, float bits (uniform; 0xffffffff: some list not full)
    unsigned int st_scanned, st_surv, st_tests, st_acc, st_groups, st_sorted;

    // Entry-level cull of one candidate: it survives iff its EXACT pixel rectangle (the pixels that pass the
    // reference's bounding-box test, rasterize_points.cu:92) contains an open pixel of the tile -- a pixel is closed
    // once its K-th depth lies below the depth lower bound of everything still to come.  Survivors go to the queue with
    // their record and rectangle.  Called by all threads; ends with a barrier and keeps `nq` up to date.
    __device__ __forceinline__ void cull_push(int id, const float4 A, const float4 Bv, bool have) {
        constexpr unsigned FULL = 0xffffffffu;
        const int S = a.S;
        bool survive = false;
        unsigned int rect = 0;
        if (have) {
            if (STATS) st_scanned++;
            if (A.z >= 0.0f) {
                int x0, x1, y0, y1;
                pixel_range(A.x, A.w, tx0, tx1, S, inv_S, pow2, half_S, x0, x1);
                pixel_range(A.y, Bv.x, ty0, ty1, S, inv_S, pow2, half_S, y0, y1);
                if (x0 <= x1 && y0 <= y1) {
                    const unsigned int rmask = ((2u << (x1 - x0)) - 1u) << (x0 - tx0);
                    unsigned int open = 0;
                    for (int yl = y0 - ty0; yl <= y1 - ty0; ++yl) open |= rmask & ~sh.closed[yl];
                    survive = open != 0;
                    rect = (unsigned int)(x0 - tx0) | ((unsigned int)(y0 - ty0) << 4) | ((unsigned int)(x1 - tx0) << 8) |
                           ((unsigned int)(y1 - ty0) << 12);
                }
            }
        }
        const unsigned m = __ballot_sync(FULL, survive);
        int wbase = 0;
        if (lane == 0 && m) wbase = atomicAdd(&sh.qcount, __popc(m));
        wbase = __shfl_sync(FULL, wbase, 0);
        if (survive) {
            const int slot = wbase + __popc(m & ((1u << lane) - 1u));
            sh.qA[slot] = A;
            sh.qB[slot] = Bv;
            sh.qD[slot] = make_uint2((unsigned int)id, rect);
            if (STATS) st_surv++;
        }
        nq += __syncthreads_count(survive);
    }

    // Rasterize the queued survivors.  Every warp owns a contiguous share of the queue; every LANE works on one
    // survivor at a time, stepping through the open pixels of its rectangle (row bit masks), and takes the next
    // survivor of the warp's share the moment it runs out of pixels -- so the lanes stay busy whatever the sizes of
    // the rectangles (one survivor per lane with a common trip count ran at 29 % lane utilisation,
    // profiles/r02_ncu_raster_v1.txt).  Accepted fragments go to the warp's own buffer (no atomics, no election: the
    // fill level is a warp-uniform register) and are inserted 32 at a time.  Then the closed-pixel masks and the
    // tile's largest K-th depth are refreshed; `lb_bits` = float bits of a depth lower bound of every entry that has
    // not been pushed yet.  Called by all threads (nq > 0); ends with a barrier.
    __device__ __forceinline__ void flush(unsigned int lb_bits) {
        constexpr unsigned FULL = 0xffffffffu;
        const int S = a.S;
        unsigned long long *wkey = sh.pend_key[warp];
        unsigned short *wpix = sh.pend_pix[warp];
        const int per = (nq + RASTER_THREADS / 32 - 1) / (RASTER_THREADS / 32);
        int whead = min(nq, warp * per);
        const int wend = min(nq, whead + per);
        const unsigned int lt_mask = (1u << lane) - 1u;
        float4 A = make_float4(0.f, 0.f, 0.f, 0.f), Bv = make_float4(0.f, 0.f, 0.f, 0.f);
        float cut = 0.f;
        unsigned long long key = 0;
        unsigned int rmask = 0, bits = 0;
        int yl = 0, yl1 = -1;
        int wcount = 0;
        for (;;) {
            // ---- lanes without an open pixel take the next survivors of the warp's share ----
            const unsigned need = __ballot_sync(FULL, bits == 0);
            const int avail = wend - whead;
            if (need && avail > 0) {
                const int r = __popc(need & lt_mask);
                if (bits == 0 && r < avail) {
                    const int i = whead + r;
                    A = sh.qA[i];
                    Bv = sh.qB[i];
                    const uint2 d = sh.qD[i];
                    key = make_key(A.z + 0.0f, (int)d.x);
                    cut = (PER_POINT_CUTOFF && a.cutoff) ? __ldg(&a.cutoff[d.x]) : a.cutoff_uniform;
                    const int x0l = d.y & 15, x1l = (d.y >> 8) & 15;
                    rmask = ((2u << (x1l - x0l)) - 1u) << x0l;
                    yl = (d.y >> 4) & 15;
                    yl1 = (d.y >> 12) & 15;
                    while (yl <= yl1) {   // first row with an open pixel
                        bits = rmask & ~lds_u32(sa_closed + yl * 4);
                        if (bits) break;
                        ++yl;
                    }
                }
                whead += min(__popc(need), avail);
            }
            if (!__any_sync(FULL, bits != 0)) {
                if (whead >= wend) break;
                continue;
            }
            // ---- one open pixel per lane ----
            const bool active = bits != 0;
            const int xl = active ? __ffs(bits) - 1 : 0;
            const int yc = active ? yl : 0;
            const int pixl = yc * RASTER_TILE + xl;
            const unsigned long long kth = lds_u64_volatile(sa_kth + pixl * (KMAX * 8));
            const float dx = lds_f32(sa_xf + xl * 4) - A.x;
            const float dy = lds_f32(sa_yf + yc * 4) - A.y;
            // rasterize_points.cu:92-97 -- same expression tree for q as the reference
            const float qv = Bv.y * dx * dx + Bv.z * dx * dy + Bv.w * dy * dy;
            const bool ok = active && (key < kth) && !(fabsf(dx) > A.w) && !(fabsf(dy) > Bv.x) && !(qv > cut);
            if (STATS) st_tests += active ? 1u : 0u;
            if (IMM) {
                // insert at once: the pixel's K-th key stays fresh, so that -- the candidates arriving front to back --
                // hardly anything beyond the K winners is ever accepted
                if (ok) {
                    if (STATS) st_acc++;
                    klist_insert<KMAX>(sh.keys + pixl * KMAX, key);
                }
            } else {
                const unsigned pm = __ballot_sync(FULL, ok);
                if (pm) {
                    if (wcount + 32 > RASTER_WPEND) {
                        pend_flush<KMAX>(sh.keys, wkey, wpix, wcount, lane);
                        wcount = 0;
                    }
                    if (ok) {
                        if (STATS) st_acc++;
                        const int pi = wcount + __popc(pm & lt_mask);
                        sts_u64(sa_wkey + pi * 8, key);
                        sts_u16(sa_wpix + pi * 2, (unsigned short)pixl);
                    }
                    wcount += __popc(pm);
                }
            }
            // next open pixel of this lane's rectangle
            bits &= bits - 1;
            if (active && !bits) {
                while (++yl <= yl1) {
                    bits = rmask & ~lds_u32(sa_closed + yl * 4);
                    if (bits) break;
                }
            }
        }
        if (!IMM) pend_flush<KMAX>(sh.keys, wkey, wpix, wcount, lane);
        __syncthreads();
        // refresh: a pixel is closed when its K-th depth lies below the bound of everything still to come (pixels
        // outside the image are always closed); the tile's largest K-th depth decides when the whole list can be dropped
        {
            const int pxl = tid & (RASTER_TILE - 1), pyl = tid >> 4;
            const bool in_img = (tx0 + pxl < S) && (ty0 + pyl < S);
            const unsigned int kz = in_img ? (unsigned int)(sh.keys[tid * KMAX + KMAX - 1] >> 32) : 0u;
            const unsigned int cm = __ballot_sync(FULL, !in_img || kz < lb_bits);
            const unsigned int wm = __reduce_max_sync(FULL, kz);
            if (lane == 0) {
                sh.closed[2 * warp] = cm & 0xffffu;
                sh.closed[2 * warp + 1] = cm >> 16;
                sh.warp_tot[warp] = wm;
            }
            if (tid == 0) sh.qcount = 0;
        }
        __syncthreads();
        {
            unsigned int mx = (lane < RASTER_THREADS / 32) ? sh.warp_tot[lane] : 0u;
            tilemax = __reduce_max_sync(FULL, mx);
        }
        nq = 0;
    }

    // Walk `count` candidates in chunks of 256 in the order they come.  fetch(i) -> packed id (or -1 to skip).
    template <typename Fetch>
    __device__ __forceinline__ void walk_unordered(int64_t count, unsigned int lb_bits, Fetch fetch) {
        int id0 = (tid < count) ? fetch((int64_t)tid) : -1;
        float4 pA = make_float4(0.f, 0.f, -1.f, 0.f), pB = make_float4(0.f, 0.f, 0.f, 0.f);
        if (id0 >= 0) {
            pA = __ldg(&a.rec[2 * (int64_t)id0]);
            pB = __ldg(&a.rec[2 * (int64_t)id0 + 1]);
        }
        for (int64_t base = 0; base < count; base += RASTER_THREADS) {
            const int id = id0;
            const float4 A = pA, Bv = pB;
            const int64_t jn = base + RASTER_THREADS + tid;
            id0 = (jn < count) ? fetch(jn) : -1;
            if (id0 >= 0) {
                pA = __ldg(&a.rec[2 * (int64_t)id0]);
                pB = __ldg(&a.rec[2 * (int64_t)id0 + 1]);
            }
            cull_push(id, A, Bv, id >= 0);
            if (nq > RASTER_QCAP - RASTER_THREADS) flush(lb_bits);
        }
    }
};

template <int KMAX, bool PER_POINT_CUTOFF, bool BLEND, bool STATS, int MINB, bool IMM>
__global__ void __launch_bounds__(RASTER_THREADS, MINB)
raster_sorted_kernel(const __grid_constant__ RasterArgs a) {
    extern __shared__ __align__(16) unsigned char raster_smem[];
    RasterShared<KMAX> &sh = *reinterpret_cast<RasterShared<KMAX> *>(raster_smem);

    const int S = a.S, B = a.B, NS = a.NS;
    const int gt = a.tile_order ? a.tile_order[blockIdx.x] : (int)blockIdx.x;   // global tile: view * B*B + tile
    const int n = gt / (B * B);
    const int tile = gt - n * B * B;
    const int ty = tile / B, tx = tile - ty * B;
    const int tid = threadIdx.x;
    const int64_t tb = ((int64_t)n * B * B + tile) * NS;
    const int lbeg = a.tile_offsets[tb], lend = a.tile_offsets[tb + NS];
    const bool overflow = lend > a.ids_capacity;   // (part of) this tile's list was not written: see below
    const bool work = overflow || lbeg < lend;
    constexpr unsigned FULL = 0xffffffffu;
    long long t_start = 0;
    if (STATS) t_start = clock64();

    RasterTile<KMAX, PER_POINT_CUTOFF, STATS, IMM> T{a, sh};
    T.tx0 = tx * RASTER_TILE;
    T.ty0 = ty * RASTER_TILE;
    T.tx1 = min(T.tx0 + RASTER_TILE, S) - 1;
    T.ty1 = min(T.ty0 + RASTER_TILE, S) - 1;
    T.tid = tid;
    T.lane = tid & 31;
    T.warp = tid >> 5;
    T.pow2 = (S & (S - 1)) == 0;
    T.inv_S = 1.0f / (float)S;
    T.half_S = 0.5f * (float)S;
    T.nq = 0;
    T.tilemax = 0xffffffffu;
    T.st_scanned = T.st_surv = T.st_tests = T.st_acc = T.st_groups = T.st_sorted = 0;

    if (work) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) sh.keys[k * RASTER_THREADS + tid] = KEY_EMPTY;
        if (tid < RASTER_TILE) {
            // nothing is closed yet except the pixels of a border tile that lie outside the image
            const unsigned int cols_out = (T.tx1 - T.tx0 + 1 >= RASTER_TILE) ? 0u : (0xffffu << (T.tx1 - T.tx0 + 1)) & 0xffffu;
            sh.closed[tid] = (T.ty0 + tid < S) ? cols_out : 0xffffu;
            sh.xf[tid] = pix_to_ndc(T.tx0 + tid, S);
            sh.yf[tid] = pix_to_ndc(T.ty0 + tid, S);
        }
        if (tid == 0) sh.qcount = 0;
        T.sa_kth = sh_addr(sh.keys) + (KMAX - 1) * 8;
        T.sa_xf = sh_addr(sh.xf);
        T.sa_yf = sh_addr(sh.yf);
        T.sa_wkey = sh_addr(sh.pend_key[T.warp]);
        T.sa_wpix = sh_addr(sh.pend_pix[T.warp]);
        T.sa_closed = sh_addr(sh.closed);
        __syncthreads();

        if (overflow) {
            // every splat of the view is a candidate; the cull's rectangle test drops those that miss the tile
            const ViewRange vr = view_range(a.first_idx, a.num_points, n, a.P0);
            const int64_t first = vr.first;
            T.walk_unordered(vr.count, 0u, [&](int64_t i) { return (int)(first + i); });
        } else {
            const SliceMap sm = make_slice_map(a.zrange, n, NS);
            const float bw = sm.dz * (1.0f / (float)RASTER_NBKT);            // bucket width
            const float inv_bw = (bw > 0.f) ? (float)RASTER_NBKT * sm.inv_dz : 0.f;
            int *cursors = reinterpret_cast<int *>(sh.qA);                    // free whenever the queue is empty
            bool done = false;
            for (int s = 0; s < NS && !done; ++s) {
                const int o0 = a.tile_offsets[tb + s], o1 = a.tile_offsets[tb + s + 1];
                const int ns = o1 - o0;
                if (ns == 0) continue;
                const float zlo = slice_bound(sm, s);
                const unsigned int lb_slice = __float_as_uint(fmaxf(zlo, 0.f));
                // everything in this and the later slices has z >= slice_bound(s)
                if (s > 0 && lb_slice > T.tilemax) break;
                // bound of everything behind this slice (the bound of the slice itself when it is the last one)
                const unsigned int lb_after = (s + 1 < NS) ? __float_as_uint(fmaxf(slice_bound(sm, s + 1), 0.f)) : lb_slice;
                const uint2 *keys = a.tile_keys + o0;
                if (ns < RASTER_SORT_MIN || !(inv_bw > 0.f)) {
                    T.walk_unordered(ns, lb_slice, [&](int64_t i) { return (int)keys[i].y; });
                } else {
                    // ---- histogram of the slice's keys over RASTER_NBKT depth buckets, exclusive scan ----
                    for (int b = tid; b < RASTER_NBKT + 4; b += RASTER_THREADS) sh.off[b] = 0;
                    if (T.nq > 0) T.flush(lb_slice);      // the scatter cursors alias the queue (ends with a barrier)
                    else __syncthreads();
                    for (int j = tid; j < ns; j += RASTER_THREADS) {
                        const float z = __uint_as_float(__ldg(&keys[j].x));
                        const int b = min(RASTER_NBKT - 1, max(0, (int)((z - zlo) * inv_bw)));
                        atomicAdd(&sh.off[b], 1u);
                    }
                    __syncthreads();
                    {
                        constexpr int PER = RASTER_NBKT / RASTER_THREADS;   // 4 buckets per thread
                        unsigned int c[PER], sum = 0;
#pragma unroll
                        for (int j = 0; j < PER; ++j) {
                            c[j] = sh.off[tid * PER + j];
                            sum += c[j];
                        }
                        unsigned int incl = sum;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const unsigned int t = __shfl_up_sync(FULL, incl, d);
                            if (T.lane >= d) incl += t;
                        }
                        if (T.lane == 31) sh.warp_tot[T.warp] = incl;
                        __syncthreads();
                        unsigned int wbase = 0;
#pragma unroll
                        for (int w = 0; w < RASTER_THREADS / 32; ++w) wbase += (w < T.warp) ? sh.warp_tot[w] : 0u;
                        unsigned int run = wbase + incl - sum;
#pragma unroll
                        for (int j = 0; j < PER; ++j) {
                            sh.off[tid * PER + j] = run;
                            run += c[j];
                        }
                        if (tid == RASTER_THREADS - 1) sh.off[RASTER_NBKT] = run;   // = ns
                        if (tid == 0) sh.bhi = 0;
                        __syncthreads();
                    }
                    // ---- rounds: as many buckets as fit the ordered array at once ----
                    int b_lo = 0;
                    while (b_lo < RASTER_NBKT && !done) {
                        const unsigned int base_off = sh.off[b_lo];
                        if (base_off == (unsigned int)ns) break;             // nothing left
                        {
                            // largest b_hi with off[b_hi] - off[b_lo] <= CAP  (off is non-decreasing); sh.bhi only grows
                            int best = b_lo + 1;
#pragma unroll
                            for (int j = 0; j < RASTER_NBKT / RASTER_THREADS; ++j) {
                                const int b = tid * (RASTER_NBKT / RASTER_THREADS) + j + 1;   // candidates 1..NBKT
                                if (b > b_lo && sh.off[b] - base_off <= (unsigned int)RASTER_SORT_CAP) best = max(best, b);
                            }
                            best = __reduce_max_sync(FULL, best);
                            if (T.lane == 0) atomicMax(&sh.bhi, best);
                        }
                        __syncthreads();
                        const int b_hi = sh.bhi;
                        const int cnt = (int)(sh.off[b_hi] - base_off);
                        const float lb_round = fmaxf(fmaf((float)(b_lo - 1), bw, zlo), 0.f);
                        if ((s > 0 || b_lo > 0) && __float_as_uint(lb_round) > T.tilemax) {
                            done = true;
                            break;
                        }
                        if (cnt > RASTER_SORT_CAP) {
                            // a single bucket that does not fit: walk its entries as they come
                            const int bsel = b_lo;
                            T.walk_unordered(ns, __float_as_uint(lb_round), [&](int64_t i) {
                                const uint2 k = keys[i];
                                const int b = min(RASTER_NBKT - 1, max(0, (int)((__uint_as_float(k.x) - zlo) * inv_bw)));
                                return b == bsel ? (int)k.y : -1;
                            });
                            if (T.nq > 0) T.flush(__float_as_uint(lb_round));   // the next round's cursors alias the queue
                            b_lo = b_hi;
                            continue;
                        }
                        // ---- scatter the round's ids into bucket order; depth lower bound of every group of 256 ----
                        for (int b = b_lo + tid; b < b_hi; b += RASTER_THREADS) {
                            const int r0 = (int)(sh.off[b] - base_off), r1 = (int)(sh.off[b + 1] - base_off);
                            cursors[b - b_lo] = r0;
                            const unsigned int lb = __float_as_uint(fmaxf(fmaf((float)(b - 1), bw, zlo), 0.f));
                            for (int g = (r0 + RASTER_THREADS - 1) / RASTER_THREADS; g * RASTER_THREADS < r1; ++g) sh.grp_lb[g] = lb;
                        }
                        // bound of everything behind this round
                        if (tid == 0)
                            sh.grp_lb[(cnt + RASTER_THREADS - 1) / RASTER_THREADS] =
                                __float_as_uint(fmaxf(fmaf((float)(b_hi - 1), bw, zlo), 0.f));
                        __syncthreads();
                        for (int j = tid; j < ns; j += RASTER_THREADS) {
                            const uint2 k = __ldg(&keys[j]);
                            const int b = min(RASTER_NBKT - 1, max(0, (int)((__uint_as_float(k.x) - zlo) * inv_bw)));
                            if (b >= b_lo && b < b_hi) sh.sorted[atomicAdd(&cursors[b - b_lo], 1)] = k.y;
                        }
                        __syncthreads();
                        if (STATS) T.st_sorted += (tid < cnt % RASTER_THREADS ? 1u : 0u) + cnt / RASTER_THREADS;
                        // ---- walk the ordered ids front to back ----
                        int id0 = (tid < cnt) ? (int)sh.sorted[tid] : -1;
                        float4 pA = make_float4(0.f, 0.f, -1.f, 0.f), pB = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (id0 >= 0) {
                            pA = __ldg(&a.rec[2 * (int64_t)id0]);
                            pB = __ldg(&a.rec[2 * (int64_t)id0 + 1]);
                        }
                        const int ngroups = (cnt + RASTER_THREADS - 1) / RASTER_THREADS;
                        for (int g = 0; g < ngroups; ++g) {
                            if ((g > 0 || b_lo > 0 || s > 0) && sh.grp_lb[g] > T.tilemax) {
                                done = true;
                                break;
                            }
                            if (STATS && tid == 0) T.st_groups++;
                            const int id = id0;
                            const float4 A = pA, Bv = pB;
                            const int jn = (g + 1) * RASTER_THREADS + tid;
                            id0 = (jn < cnt) ? (int)sh.sorted[jn] : -1;
                            if (id0 >= 0) {
                                pA = __ldg(&a.rec[2 * (int64_t)id0]);
                                pB = __ldg(&a.rec[2 * (int64_t)id0 + 1]);
                            }
                            T.cull_push(id, A, Bv, id >= 0);
                            if (T.nq >= a.flush_min || T.nq > RASTER_QCAP - RASTER_THREADS) T.flush(sh.grp_lb[g + 1]);
                        }
                        // the next round's cursors alias the queue: empty it
                        if (T.nq > 0) T.flush(sh.grp_lb[ngroups]);
                        b_lo = b_hi;
                    }
                }
                // thresholds up to date before the next slice's bound is tested
                if (T.nq > 0) T.flush(lb_after);
            }
        }
        if (T.nq > 0) T.flush(0u);
    }
    if (STATS && a.stats) {
        const long long t_end = clock64();
        // one atomic per warp and counter
        const unsigned int v[5] = {T.st_scanned, T.st_surv, T.st_tests, T.st_acc, T.st_sorted};
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const unsigned int w = __reduce_add_sync(FULL, v[i]);
            if (T.lane == 0 && w) atomicAdd(&a.stats[i], (unsigned long long)w);
        }
        if (tid == 0) {
            if (T.st_groups) atomicAdd(&a.stats[5], (unsigned long long)T.st_groups);
            if (overflow) atomicAdd(&a.stats[6], 1ull);
            // per-tile record: {cycles before the epilogue, list length, groups visited, launch position}
            if (a.tile_dbg) a.tile_dbg[gt] = make_uint4((unsigned int)(t_end - t_start), (unsigned int)(lend - lbeg), T.st_groups, blockIdx.x);
        }
    }
    raster_epilogue<KMAX, BLEND>(a, sh.keys, work, n, T.tx0, T.ty0, T.pow2, T.inv_S);
}


template <int KMAX>
static int launch_scatter(dss_ctx *ctx, const RasterArgs &a_in, cudaStream_t st) {
    RasterArgs a = a_in;
    const int ntiles = a.B * a.B * a.N;
    dim3 grid((unsigned)ntiles);
    StageScope prof(ctx, ST_RASTER_FWD, st);
    a.tile_order = nullptr;
    a.flush_min = ctx->raster_flush_min > 0 ? ctx->raster_flush_min : RASTER_FLUSH_MIN;
    if (!ctx->no_tile_order && ntiles > 4 * ctx->sm_count) {
        int32_t *order = nullptr;
        int rc = ctx_get(ctx, BUF_TILE_ORDER, (size_t)ntiles, &order);
        if (rc) return rc;
        raster_tile_order_kernel<<<1, 1024, 0, st>>>(a.tile_offsets, a.NS, ntiles, order);
        DSS_LAUNCH_CHECK(ctx);
        a.tile_order = order;
    }
    const bool blend = a.image != nullptr;
    constexpr size_t smem = sizeof(RasterShared<KMAX>);
    auto launch = [&](auto kern) -> int {
        DSS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, RASTER_THREADS, smem, st>>>(a);
        return DSS_OK;
    };
    int rc;
    const bool imm = ctx->raster_immediate != 0;
    if (a.stats) {   // debug counters on: one generic instantiation per insertion mode is enough
        if (imm) rc = blend ? launch(raster_sorted_kernel<KMAX, true, true, true, 3, true>) : launch(raster_sorted_kernel<KMAX, true, false, true, 3, true>);
        else rc = blend ? launch(raster_sorted_kernel<KMAX, true, true, true, 3, false>) : launch(raster_sorted_kernel<KMAX, true, false, true, 3, false>);
    } else if (blend) {
        if (a.cutoff) rc = launch(raster_sorted_kernel<KMAX, true, true, false, 4, false>);
        else rc = imm ? launch(raster_sorted_kernel<KMAX, false, true, false, 4, true>) : launch(raster_sorted_kernel<KMAX, false, true, false, 4, false>);
    } else {
        if (a.cutoff) rc = imm ? launch(raster_sorted_kernel<KMAX, true, false, false, 4, true>) : launch(raster_sorted_kernel<KMAX, true, false, false, 4, false>);
        else rc = launch(raster_sorted_kernel<KMAX, false, false, false, 4, false>);
    }
    if (rc) return rc;
    DSS_LAUNCH_CHECK(ctx);
    return DSS_OK;
}

template <int KMAX>
static int launch_raster(dss_ctx *ctx, const RasterArgs &a, cudaStream_t st) {
    dim3 grid((unsigned)(a.B * a.B), (unsigned)a.N);
    StageScope prof(ctx, ST_RASTER_FWD, st);
    const bool blend = a.image != nullptr;
    if (blend) {
        if (a.cutoff)
            raster_fwd_kernel<KMAX, true, true><<<grid, RASTER_THREADS, 0, st>>>(a);
        else
            raster_fwd_kernel<KMAX, false, true><<<grid, RASTER_THREADS, 0, st>>>(a);
    } else {
        if (a.cutoff)
            raster_fwd_kernel<KMAX, true, false><<<grid, RASTER_THREADS, 0, st>>>(a);
        else
            raster_fwd_kernel<KMAX, false, false><<<grid, RASTER_THREADS, 0, st>>>(a);
    }
    DSS_LAUNCH_CHECK(ctx);
    return DSS_OK;
}

int raster_forward(dss_ctx *ctx, const RasterArgs &a, cudaStream_t st) {
    if (a.N <= 0 || a.S <= 0) return DSS_OK;
    if (!a.force_pixel_parallel) {
        if (a.K <= 5) return launch_scatter<5>(ctx, a, st);
        if (a.K <= 8) return launch_scatter<8>(ctx, a, st);
    }
    if (a.K <= 5) return launch_raster<5>(ctx, a, st);
    if (a.K <= 8) return launch_raster<8>(ctx, a, st);
    if (a.K <= 16) return launch_raster<16>(ctx, a, st);
    return launch_raster<DSS_MAX_POINTS_PER_PIXEL>(ctx, a, st);
}

int bin_and_raster(dss_ctx *ctx, RasterArgs a, const int64_t *first_idx, const int64_t *num_points,
                   int64_t P0, cudaStream_t st) {
    const int S = a.S;
    const bool scatter_path = a.K <= 8 && !a.force_pixel_parallel;
    a.B = (S + RASTER_TILE - 1) / RASTER_TILE;
    a.NS = scatter_path ? choose_depth_slices(a.B) : 1;
    if (a.NS > 1 && ctx->ns_override > 0 && ctx->ns_override < a.NS) a.NS = ctx->ns_override;
    a.first_idx = first_idx;
    a.num_points = num_points;
    a.P0 = P0;
    const int64_t nb = (int64_t)a.N * a.B * a.B * a.NS;
    if (nb + 1 >= (int64_t)INT32_MAX) {
        set_error("too many tiles (%lld)", (long long)nb);
        return DSS_E_INVALID;
    }
    int32_t *counts = nullptr, *offsets = nullptr;
    uint2 *keys = nullptr;
    float *zrange = const_cast<float *>(a.zrange);   // non-null: the caller's preprocess already produced it
    int rc;
    if ((rc = ctx_get(ctx, BUF_TILE_COUNTS, (size_t)(nb + 1), &counts))) return rc;
    if ((rc = ctx_get(ctx, BUF_TILE_OFFSETS, (size_t)(nb + 1), &offsets))) return rc;
    if (scatter_path && zrange == nullptr) {
        if ((rc = ctx_get(ctx, BUF_ZRANGE, (size_t)(2 * a.N), &zrange))) return rc;
        if ((rc = compute_zrange(ctx, a.rec, first_idx, num_points, a.N, P0, zrange, st))) return rc;
    }
    a.zrange = zrange;
    if ((rc = bin_count_and_scan(ctx, a.rec, first_idx, num_points, a.N, P0, S, RASTER_TILE, a.NS, zrange, counts,
                                 offsets, st)))
        return rc;
    a.tile_offsets = offsets;
    a.stats = nullptr;
    a.tile_dbg = nullptr;
    if (ctx->raster_stats) {
        unsigned long long *sp = reinterpret_cast<unsigned long long *>(ctx->buf[BUF_STATS]);   // sized by dss_debug_raster_stats
        a.stats = sp;
        ctx->raster_dbg_tiles = 0;
        if (sp && ctx->cap[BUF_STATS] >= (8 + 2 * (size_t)(a.N * a.B * a.B)) * sizeof(unsigned long long)) {
            a.tile_dbg = reinterpret_cast<uint4 *>(sp + 8);
            ctx->raster_dbg_tiles = a.N * a.B * a.B;
        }
    }
    // Size of the CSR key list.  It is only known on the device (last entry of the scan), and the host never waits for
    // it in steady state: the scan's total is published into mapped pinned memory (no copy engine, no sync) and read
    // by the NEXT call, which grows the buffer if the lists have outgrown it.  Meanwhile scatter and rasterizer run
    // with the buffer they have: the scatter drops entries beyond the capacity, the rasterizer recognises the tiles
    // whose lists are incomplete (offset past the capacity) and takes their candidates from the view's records
    // directly -- slower for those tiles, same result.  Only a context that has never sized the buffer (first call)
    // and the pixel-parallel fallback for K > 8 (which has no such path) wait for the total.
    int64_t cap = (int64_t)(ctx->cap[BUF_TILE_IDS] / sizeof(uint2));
    volatile int32_t *h_total = reinterpret_cast<volatile int32_t *>(ctx->h_pinned);
    if ((rc = publish_words(ctx, offsets + nb, ctx->h_pinned, 1, st))) return rc;
    const bool must_wait = cap == 0 || !scatter_path || ctx->sync_forward;
    int64_t want = 0;
    if (must_wait) {
        DSS_CUDA_TRY(cudaStreamSynchronize(st));
        want = (int64_t)h_total[0];
    } else {
        want = ctx->tile_total_hint;    // what the previous call published (read below, after its kernels were queued)
    }
    if (want < 0) {
        set_error("tile list size overflowed int32");
        return DSS_E_INVALID;
    }
    if (want > cap) cap = want + want / 4 + 1024;   // headroom so that slowly growing lists never outgrow the buffer
    if (cap > (int64_t)INT32_MAX) cap = INT32_MAX;
    if (cap < 1) cap = 1;
    if ((rc = ctx_get(ctx, BUF_TILE_IDS, (size_t)cap, &keys))) return rc;      // grows when needed, otherwise a no-op
    {
        const size_t have = ctx->cap[BUF_TILE_IDS] / sizeof(uint2);
        a.ids_capacity = (int)(have > (size_t)INT32_MAX ? (size_t)INT32_MAX : have);
        // testing (dss_debug_limit_tile_capacity): pretend the buffer is smaller, which forces the overflow path
        if (ctx->tile_cap_limit > 0 && a.ids_capacity > ctx->tile_cap_limit) a.ids_capacity = (int)ctx->tile_cap_limit;
    }
    if (!scatter_path && (int64_t)h_total[0] > a.ids_capacity) {
        set_error("internal: tile list buffer too small for the pixel-parallel path");
        return DSS_E_CAPACITY;
    }
    if ((rc = bin_scatter(ctx, a.rec, first_idx, num_points, a.N, P0, S, RASTER_TILE, a.NS, zrange, offsets, counts, keys,
                          a.ids_capacity, true, st)))
        return rc;
    a.tile_keys = keys;
    if ((rc = raster_forward(ctx, a, st))) return rc;
    // hint for the next call: whatever total the device has published by now (this call's, if the GPU is ahead of
    // the host, otherwise an earlier one) -- it only sizes the buffer, correctness never depends on it
    ctx->tile_total_hint = (int64_t)h_total[0];
    return DSS_OK;
}

}  // namespace dss

extern "C" {

int dss_debug_raster_stats(dss_ctx *ctx, int enable, uint64_t out[8]) {
    using namespace dss;
    DSS_REQUIRE(ctx != nullptr, "ctx is null");
    unsigned long long *sp = nullptr;
    int rc = ctx_get(ctx, BUF_STATS, 8 + 2 * ((size_t)1 << 18), &sp);   // 8 counters + per-tile records of up to 256k tiles
    if (rc) return rc;
    DSS_CUDA_TRY(cudaDeviceSynchronize());
    if (out) DSS_CUDA_TRY(cudaMemcpy(out, sp, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (enable) DSS_CUDA_TRY(cudaMemset(sp, 0, 8 * sizeof(unsigned long long)));
    ctx->raster_stats = enable ? 1 : 0;
    return DSS_OK;
}

int dss_debug_tile_profile(dss_ctx *ctx, uint32_t *out, int64_t max_tiles) {
    using namespace dss;
    DSS_REQUIRE(ctx != nullptr && out != nullptr, "null pointer");
    int64_t n = ctx->raster_dbg_tiles < max_tiles ? ctx->raster_dbg_tiles : max_tiles;
    if (n <= 0 || !ctx->buf[BUF_STATS]) return 0;
    DSS_CUDA_TRY(cudaDeviceSynchronize());
    DSS_CUDA_TRY(cudaMemcpy(out, reinterpret_cast<unsigned long long *>(ctx->buf[BUF_STATS]) + 8, (size_t)n * 16, cudaMemcpyDeviceToHost));
    return (int)n;
}

int dss_debug_limit_tile_capacity(dss_ctx *ctx, int64_t max_entries) {
    DSS_REQUIRE(ctx != nullptr, "ctx is null");
    DSS_REQUIRE(max_entries >= 0 && max_entries <= (int64_t)INT32_MAX, "bad limit");
    ctx->tile_cap_limit = max_entries;
    return DSS_OK;
}

int dss_splat_points(dss_ctx *ctx, const float *points, const float *ellipse_params, const float *cutoff_thres,
                     const float *radii, const int64_t *first_idx, const int64_t *num_points, int N, int64_t P,
                     float depth_merging_thres, int image_size, int points_per_pixel, int bin_size, int32_t *idx,
                     float *zbuf, float *qvalue, float *occupancy, void *stream) {
    using namespace dss;
    (void)bin_size;
    cudaStream_t st = (cudaStream_t)stream;
    DSS_REQUIRE(ctx != nullptr, "ctx is null");
    DSS_REQUIRE(N >= 0 && P >= 0, "negative size");
    DSS_REQUIRE(image_size > 0, "image_size must be positive");
    DSS_REQUIRE(points_per_pixel > 0 && points_per_pixel <= DSS_MAX_POINTS_PER_PIXEL,
                "points_per_pixel must be in [1, 64]");
    DSS_REQUIRE(P < (int64_t)INT32_MAX, "more than 2^31-1 packed points");
    if (N == 0) return DSS_OK;
    DSS_REQUIRE(idx && occupancy && first_idx && num_points, "null pointer");
    DSS_REQUIRE(P == 0 || (points && ellipse_params && cutoff_thres && radii), "null input array");
    float4 *rec = nullptr;
    int rc;
    if ((rc = ctx_get(ctx, BUF_RECORDS, (size_t)(2 * (P > 0 ? P : 1)), &rec))) return rc;
    if ((rc = pack_records(ctx, points, radii, ellipse_params, P, rec, st))) return rc;
    RasterArgs a;
    memset(&a, 0, sizeof(a));
    a.rec = rec;
    a.cutoff = cutoff_thres;
    a.cutoff_uniform = 0.f;
    a.N = N;
    a.S = image_size;
    a.K = points_per_pixel;
    a.depth_merge = depth_merging_thres;
    a.idx = idx;
    a.zbuf = zbuf;
    a.qvalue = qvalue;
    a.occ = occupancy;
    return bin_and_raster(ctx, a, first_idx, num_points, P, st);
}

}  // extern "C"
