// backward.cu -- visibility, per-view search radius (exact radix select), occupancy gather,
// z-buffer and colour scatters.
//
// Replaces the fast branch of EllipticalRasterizer.backward (DSS/core/rasterizer.py:845-972) and
// RasterizePointsBackwardCudaFastKernel (DSS/csrc/rasterize_points_backward.cu:30-212).  The reference
// scatters from pixels to points with two float atomics per (pixel, point) pair after building an FRNN
// grid with per-view host loops; here every visible point GATHERS its pixel disc -- no grid, no
// atomics, deterministic -- which is the same sum because the grid is only an accelerator for the
// d^2 <= r^2 test.
#include <math_constants.h>

#include "common.cuh"
#include "kernels.cuh"

namespace dss {

static inline unsigned int nblocks(int64_t items, int threads, int sm_count, int per_sm) {
    int64_t b = (items + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned int)b;
}

// ---------------------------------------------------------------------------------------------
// visibility (rasterizer.py:854-860): any slot of a pixel whose first slot is occupied.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
visibility_kernel(const int32_t *__restrict__ idx, int64_t num_pixels, int K, int64_t P,
                  uint8_t *__restrict__ visible) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_pixels;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t *row = idx + i * K;
        if (row[0] < 0) continue;
        for (int k = 0; k < K; ++k) {
            const int p = row[k];
            if (p >= 0 && p < P) visible[p] = 1;
        }
    }
}

int visibility_from_idx(dss_ctx *ctx, const int32_t *idx, int64_t num_pixels, int K, int64_t P,
                        uint8_t *visible, cudaStream_t st) {
    DSS_CUDA_TRY(cudaMemsetAsync(visible, 0, (size_t)(P > 0 ? P : 0), st));
    if (num_pixels == 0 || P == 0) return DSS_OK;
    StageScope prof(ctx, ST_VISIBILITY, st);
    visibility_kernel<<<nblocks(num_pixels, 256, ctx->sm_count, 16), 256, 0, st>>>(idx, num_pixels, K, P, visible);
    DSS_LAUNCH_CHECK(ctx);
    return DSS_OK;
}

// ---------------------------------------------------------------------------------------------
// Search radius: radii_s * lower median of the flattened (rx, ry) of the view's visible points
// (rasterizer.py:888, torch.median = element (m-1)/2 of the ascending sort of m = 2 n_vis values).
// Exact 4-pass (8 bits each, MSB first) radix select on the order-preserving uint image of the floats.
// hist layout: (N, 4, 256) uint32.  Each block first re-derives the prefix chosen by the previous
// passes from their (complete) histograms -- 256-bin scans, negligible -- so no host involvement.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int float_key(float f) {
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned int k) {
    const unsigned int b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

// Walk histograms of passes [0, upto) and return (prefix, remaining rank).  Executed by one warp.
__device__ void select_resolve(const unsigned int *hist_n, int upto, unsigned int &prefix,
                               unsigned long long &rank, unsigned long long &total) {
    const int lane = threadIdx.x & 31;
    prefix = 0;
    rank = 0;
    total = 0;
    for (int pass = 0; pass < upto; ++pass) {
        const unsigned int *h = hist_n + pass * 256;
        // each lane owns 8 consecutive bins
        unsigned int c[8];
        unsigned long long s = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            c[j] = h[lane * 8 + j];
            s += c[j];
        }
        unsigned long long incl = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const unsigned long long tot = __shfl_sync(0xffffffffu, incl, 31);
        if (pass == 0) {
            total = tot;
            rank = (tot > 0) ? (tot - 1) / 2 : 0;  // lower median
        }
        unsigned long long excl = incl - s;
        // find the bin containing `rank`
        int found = -1;
        unsigned long long found_excl = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (found < 0 && rank >= excl && rank < excl + c[j]) {
                found = lane * 8 + j;
                found_excl = excl;
            }
            excl += c[j];
        }
        const unsigned int who = __ballot_sync(0xffffffffu, found >= 0);
        int digit = 0;
        unsigned long long dexcl = 0;
        if (who) {
            const int src = __ffs(who) - 1;
            digit = __shfl_sync(0xffffffffu, found, src);
            dexcl = __shfl_sync(0xffffffffu, found_excl, src);
        }
        prefix = (prefix << 8) | (unsigned int)digit;
        rank -= dexcl;
    }
}

__global__ void __launch_bounds__(256)
select_hist_kernel(const float4 *__restrict__ rec, const float *__restrict__ radii,
                   const uint8_t *__restrict__ visible, const int64_t *__restrict__ first_idx,
                   const int64_t *__restrict__ num_points, int64_t P0_shared, int pass,
                   unsigned int *__restrict__ hist) {
    __shared__ unsigned int s_hist[256];
    __shared__ unsigned int s_prefix;
    const int n = blockIdx.y;
    unsigned int *hist_n = hist + (int64_t)n * 4 * 256;
    s_hist[threadIdx.x] = 0;
    if (threadIdx.x < 32) {
        unsigned int prefix;
        unsigned long long rank, total;
        select_resolve(hist_n, pass, prefix, rank, total);
        if (threadIdx.x == 0) s_prefix = prefix;
    }
    __syncthreads();
    const unsigned int prefix = s_prefix;
    const int shift = 24 - 8 * pass;
    const ViewRange vr = view_range(first_idx, num_points, n, P0_shared);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < vr.count;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = vr.first + i;
        if (!visible[p]) continue;
        float rx, ry;
        if (rec) {
            rx = __ldg(&rec[2 * p]).w;
            ry = __ldg(&rec[2 * p + 1]).x;
        } else {
            rx = radii[p * 2];
            ry = radii[p * 2 + 1];
        }
        const unsigned int kx = float_key(rx), ky = float_key(ry);
        if (pass == 0 || (kx >> (shift + 8)) == prefix) atomicAdd(&s_hist[(kx >> shift) & 255u], 1u);
        if (pass == 0 || (ky >> (shift + 8)) == prefix) atomicAdd(&s_hist[(ky >> shift) & 255u], 1u);
    }
    __syncthreads();
    const unsigned int v = s_hist[threadIdx.x];
    if (v) atomicAdd(&hist_n[pass * 256 + threadIdx.x], v);
}

__global__ void select_final_kernel(const unsigned int *__restrict__ hist, float radii_s, float *__restrict__ rs) {
    const int n = blockIdx.x;
    unsigned int prefix;
    unsigned long long rank, total;
    select_resolve(hist + (int64_t)n * 4 * 256, 4, prefix, rank, total);
    if (threadIdx.x == 0) rs[n] = (total > 0) ? key_float(prefix) * radii_s : 0.0f;
}

int search_radius(dss_ctx *ctx, const float4 *rec, const float *radii, const uint8_t *visible,
                  const int64_t *first_idx, const int64_t *num_points, int N, int64_t P0, float radii_s,
                  float *rs, cudaStream_t st) {
    if (N <= 0) return DSS_OK;
    unsigned int *hist = nullptr;
    int rc = ctx_get(ctx, BUF_SELECT, (size_t)N * 4 * 256, &hist);
    if (rc) return rc;
    StageScope prof(ctx, ST_SEARCH_RADIUS, st);
    DSS_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)N * 4 * 256 * sizeof(unsigned int), st));
    if (P0 > 0) {
        dim3 grid(nblocks(P0, 256, ctx->sm_count, 4), N);
        for (int pass = 0; pass < 4; ++pass) {
            select_hist_kernel<<<grid, 256, 0, st>>>(rec, radii, visible, first_idx, num_points, P0, pass, hist);
            DSS_LAUNCH_CHECK(ctx);
        }
    }
    select_final_kernel<<<N, 32, 0, st>>>(hist, radii_s, rs);
    DSS_LAUNCH_CHECK(ctx);
    return DSS_OK;
}

// ---------------------------------------------------------------------------------------------
// Occupancy gather.  One warp per visible splat; lanes stride over the (2R+1)^2 pixel window that
// conservatively contains the disc of radius r_n, evaluate the reference's per-pair rule
// (rasterize_points_backward.cu:141-178) and warp-reduce.  32 consecutive splats per warp, the visible
// ones are processed in turn; each lane finally stores the result of "its" splat (coalesced float2).
// ---------------------------------------------------------------------------------------------
constexpr int OCC_WARPS = 8;
constexpr int OCC_TILE = 32;          // pixels per side of a backward tile
constexpr int OCC_TILE_THREADS = 256;

// Does the (tile + 2R)^2 window of alpha gradients fit the tile kernel's shared memory?  Evaluated
// identically on the device by both kernels so that every view is processed by exactly one of them.
__host__ __device__ __forceinline__ int occ_halo(float r, int S) { return (int)ceilf(r * 0.5f * (float)S) + 2; }
__host__ __device__ __forceinline__ bool occ_tile_fits(float r, int S, int smem_bytes) {
    if (!(r >= 0.0f) || !(r < 4.0f) || smem_bytes <= 0) return false;
    const int side = OCC_TILE + 2 * occ_halo(r, S);
    return (size_t)(2 * side * side + 2 * side) * sizeof(float) <= (size_t)smem_bytes;
}

__device__ __forceinline__ int centre_tile(float px, float py, int S, int OB) {
    const float half_S = 0.5f * (float)S;
    const int xi = min(max((int)floorf((px + 1.0f) * half_S), 0), S - 1);
    const int yi = min(max((int)floorf((py + 1.0f) * half_S), 0), S - 1);
    return (yi / OCC_TILE) * OB + (xi / OCC_TILE);
}

__device__ __forceinline__ bool occ_eligible(const float4 A) {
    // rasterize_points_backward.cu:145 -- outside the renderable area
    return !(A.z < 0.0f || fabsf(A.y) > 1.0f || fabsf(A.x) > 1.0f);
}

// Bin the visible, eligible splats by the 32x32 tile that contains their centre (one tile per splat,
// so the id list is bounded by P and no host round-trip is needed).  PASS 0 counts, PASS 1 scatters.
template <int PASS>
__global__ void __launch_bounds__(256)
occ_bin_kernel(const float4 *__restrict__ rec, const uint8_t *__restrict__ visible,
               const int64_t *__restrict__ first_idx, const int64_t *__restrict__ num_points,
               int64_t P0_shared, int S, int OB, int32_t *__restrict__ counters, int32_t *__restrict__ ids) {
    extern __shared__ int32_t s_hist[];
    const int n = blockIdx.y;
    const int nt = OB * OB;
    const ViewRange vr = view_range(first_idx, num_points, n, P0_shared);
    constexpr int ITEMS = 8;
    const int64_t chunk0 = (int64_t)blockIdx.x * (256 * ITEMS);
    if (chunk0 >= vr.count) return;
    for (int t = threadIdx.x; t < nt; t += 256) s_hist[t] = 0;
    __syncthreads();
    int tile[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        tile[j] = -1;
        const int64_t i = chunk0 + j * 256 + threadIdx.x;
        if (i < vr.count && visible[vr.first + i]) {
            const float4 A = __ldg(&rec[2 * (vr.first + i)]);
            if (occ_eligible(A)) {
                tile[j] = centre_tile(A.x, A.y, S, OB);
                atomicAdd(&s_hist[tile[j]], 1);
            }
        }
    }
    __syncthreads();
    int32_t *cnt = counters + (int64_t)n * nt;
    for (int t = threadIdx.x; t < nt; t += 256) {
        const int v = s_hist[t];
        if (v) {
            const int base = atomicAdd(&cnt[t], v);
            if (PASS == 1) s_hist[t] = base;
        }
    }
    if (PASS == 0) return;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
        if (tile[j] >= 0) ids[atomicAdd(&s_hist[tile[j]], 1)] = (int32_t)(vr.first + chunk0 + j * 256 + threadIdx.x);
}

// One CTA per (32x32 tile, view): the tile's window of alpha gradients (tile + halo of the search radius)
// is staged in shared memory once; each warp then takes splats of the tile's list and gathers over the
// splat's own (2R+1)^2 sub-window with conflict-free shared-memory reads.
__global__ void __launch_bounds__(OCC_TILE_THREADS)
occ_tile_kernel(const float4 *__restrict__ rec, const float *__restrict__ rs, const float *__restrict__ grad,
                int pix_stride, int pix_offset, const int32_t *__restrict__ tile_offsets,
                const int32_t *__restrict__ tile_ids, int S, int OB, int smem_bytes, float2 *__restrict__ grad_xy) {
    extern __shared__ float s_g[];
    const int n = blockIdx.y, tile = blockIdx.x;
    const float r = rs[n];
    if (!occ_tile_fits(r, S, smem_bytes)) return;
    const int64_t tb = (int64_t)n * OB * OB + tile;
    int beg = tile_offsets[tb], end = tile_offsets[tb + 1];
    if (beg == end) return;
    {
        // a tile's list is split over up to gridDim.z CTAs (>= 64 splats each) so that dense tiles do not
        // serialise on one CTA; every CTA stages the (small) window itself
        const int count = end - beg;
        const int parts = min((int)gridDim.z, (count + 63) / 64);
        if ((int)blockIdx.z >= parts) return;
        const int per = (count + parts - 1) / parts;
        beg += (int)blockIdx.z * per;
        end = min(end, beg + per);
        if (beg >= end) return;
    }
    const int R = occ_halo(r, S);
    const int side = OCC_TILE + 2 * R;
    const int ty = tile / OB, tx = tile - ty * OB;
    const int wx0 = tx * OCC_TILE - R, wy0 = ty * OCC_TILE - R;
    const float *gview = grad + ((int64_t)n * S * S) * pix_stride + pix_offset;
    float *s_gp = s_g + side * side, *s_xf = s_gp + side * side, *s_yf = s_xf + side;
    for (int wy = threadIdx.x >> 5; wy < side; wy += OCC_TILE_THREADS / 32) {
        const int yi = wy0 + wy;
        const bool row_in = yi >= 0 && yi < S;
        const float *grow = gview + ((int64_t)(S - 1 - yi) * S + (S - 1)) * pix_stride;
        for (int wx = threadIdx.x & 31; wx < side; wx += 32) {
            const int xi = wx0 + wx;
            float g = 0.0f;   // zero outside the image: such pixels then drop out of the sums
            if (row_in && xi >= 0 && xi < S) g = __ldg(grow - (int64_t)xi * pix_stride);
            // Negative gradients act on the whole search disc, positive ones only inside the splat's own
            // bounding box (rasterize_points_backward.cu:161-168): two planes, two loops, no per-pair sign tests.
            s_g[wy * side + wx] = fminf(g, 0.0f);
            s_gp[wy * side + wx] = fmaxf(g, 0.0f);
        }
    }
    // exact pixel-centre NDC coordinates of the window's columns / rows (the reference's PixToNdc, division
    // included, evaluated once per CTA instead of once per pair)
    for (int i = threadIdx.x; i < side; i += OCC_TILE_THREADS) {
        s_xf[i] = pix_to_ndc(wx0 + i, S);
        s_yf[i] = pix_to_ndc(wy0 + i, S);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float r2 = r * r;
    const float half_S = 0.5f * (float)S;
    constexpr unsigned FULL = 0xffffffffu;
    // Lane <-> window column, loop over window rows.  dx (hence dx^2 and the x half of the bbox test) is
    // a per-lane constant of the splat, dy is shared by the row, and since dx is constant along a column
    //     sum_rows dx * w = dx * sum_rows w ,   w = g / max(d2, 1e-10)
    // each pair costs one reciprocal, one add and one fma, branch-free.  Narrow windows pack several
    // splats per warp (8 or 16 lanes each).
    const int Rw = R - 1;                       // per-splat window half-width in pixels (covers the disc)
    const int Wwin = 2 * Rw + 1;
    const int lpp = Wwin <= 8 ? 8 : (Wwin <= 16 ? 16 : 32);   // lanes per splat
    const int groups = 32 / lpp;
    const int grp = lane / lpp, gl = lane - grp * lpp;
    const int nwarps = OCC_TILE_THREADS / 32;
    // software pipeline: the record of the next splat is fetched while the current one is being summed
    int n_p = 0;
    float4 n_A = make_float4(0.f, 0.f, 0.f, 0.f);
    float n_ry = 0.f;
    {
        const int k = beg + warp * groups + grp;
        if (k < end) {
            n_p = tile_ids[k];
            n_A = __ldg(&rec[2 * (int64_t)n_p]);
            n_ry = __ldg(&rec[2 * (int64_t)n_p + 1]).x;
        }
    }
    for (int k0 = beg + warp * groups; k0 < end; k0 += nwarps * groups) {
        const int k = k0 + grp;
        const bool have = k < end;
        const int p = n_p;
        const float px = n_A.x, py = n_A.y, rx = n_A.w, ry = n_ry;
        {
            const int kn = k + nwarps * groups;
            if (kn < end) {
                n_p = tile_ids[kn];
                n_A = __ldg(&rec[2 * (int64_t)n_p]);
                n_ry = __ldg(&rec[2 * (int64_t)n_p + 1]).x;
            }
        }
        // window origin in staged-window coordinates (always inside: the halo R = Rw + 1 covers it)
        const int cx = min(max((int)floorf((px + 1.0f) * half_S), tx * OCC_TILE), tx * OCC_TILE + OCC_TILE - 1);
        const int cy = min(max((int)floorf((py + 1.0f) * half_S), ty * OCC_TILE), ty * OCC_TILE + OCC_TILE - 1);
        const int ox = cx - Rw - wx0, oy = cy - Rw - wy0;
        float gx = 0.f, gy = 0.f;
        // windows wider than 32 columns: the first 32 columns go through the column-per-lane loop below, the
        // remaining (Wwin - 32) x Wwin strip is walked in flattened order with all lanes busy (a second
        // column block would run Wwin iterations with only Wwin - 32 lanes active)
        const int Wmain = (lpp == 32 && Wwin > 32 && Wwin <= 64) ? 32 : Wwin;
        if (Wmain < Wwin) {
            const int Wt = Wwin - 32, total = Wt * Wwin;
            int wy = lane / Wt, wx = lane - wy * Wt;
            const int step_y = 32 / Wt, step_x = 32 - step_y * Wt;
            for (int t = lane; t < total; t += 32) {
                const int sx = ox + 32 + wx, sy = oy + wy;
                const float g = s_g[sy * side + sx];
                const float dx = s_xf[sx] - px, dy = s_yf[sy] - py;
                const float d2 = fmaf(dy, dy, dx * dx);
                float inv;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(fmaxf(d2, 1e-10f)));
                const float w = (have && !(d2 > r2)) ? g * inv : 0.0f;
                gx = fmaf(dx, w, gx);
                gy = fmaf(dy, w, gy);
                wx += step_x;
                wy += step_y;
                if (wx >= Wt) {
                    wx -= Wt;
                    wy += 1;
                }
            }
        }
        for (int cb = 0; cb < Wmain; cb += lpp) {           // column blocks (one unless the window is > 64 wide)
            const bool col_ok = have && (cb + gl < Wmain);
            const int sx = col_ok ? ox + cb + gl : ox;      // inactive lanes read a valid column, result dropped
            const float dx = s_xf[sx] - px;
            const float dx2 = dx * dx;
            const float *col = s_g + oy * side + sx;
            const float *yfp = s_yf + oy;
            float sw = 0.f, swy = 0.f;
#pragma unroll 4
            for (int j = 0; j < Wwin; ++j) {
                const float g = col[j * side];                 // <= 0
                const float dy = yfp[j] - py;
                const float d2 = fmaf(dy, dy, dx2);
                float inv;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(fmaxf(d2, 1e-10f)));
                const float w = (d2 > r2) ? 0.0f : g * inv;    // rasterize_points_backward.cu:156, 170-172
                sw += w;
                swy = fmaf(w, dy, swy);
            }
            if (col_ok) {
                gx = fmaf(dx, sw, gx);
                gy += swy;
            }
        }
        // positive gradients: only pixels inside the splat's bounding box (and the disc) count
        {
            const int bw = min(Rw, (int)fminf(ceilf(rx * half_S) + 1.0f, 4096.0f));
            const int bh = min(Rw, (int)fminf(ceilf(ry * half_S) + 1.0f, 4096.0f));
            const int Wb = 2 * bw + 1, total = have ? Wb * (2 * bh + 1) : 0;
            const int bx0 = ox + Rw - bw, by0 = oy + Rw - bh;
            const float inv_Wb = 1.0f / (float)Wb;
            for (int t = gl; t < total; t += lpp) {
                const int wy = __float2int_rz(((float)t + 0.5f) * inv_Wb), wx = t - wy * Wb;   // exact for these sizes
                const int sx = bx0 + wx, sy = by0 + wy;
                const float g = s_gp[sy * side + sx];         // >= 0
                const float dx = s_xf[sx] - px, dy = s_yf[sy] - py;
                const float d2 = fmaf(dy, dy, dx * dx);
                const bool use = !(fabsf(dx) > rx) && !(fabsf(dy) > ry) && !(d2 > r2);
                float inv;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(fmaxf(d2, 1e-10f)));
                const float w = use ? g * inv : 0.0f;
                gx = fmaf(dx, w, gx);
                gy = fmaf(dy, w, gy);
            }
        }
        for (int d = lpp >> 1; d > 0; d >>= 1) {
            gx += __shfl_xor_sync(FULL, gx, d);
            gy += __shfl_xor_sync(FULL, gy, d);
        }
        if (have && gl == 0) grad_xy[p] = make_float2(gx, gy);
    }
}

// ---------------------------------------------------------------------------------------------
// Occupancy gather.  One warp per visible splat; lanes stride over the (2R+1)^2 pixel window that
// conservatively contains the disc of radius r_n, evaluate the reference's per-pair rule
// (rasterize_points_backward.cu:141-178) and warp-reduce.  32 consecutive splats per warp, the visible
// ones are processed in turn; each lane finally stores the result of "its" splat (coalesced float2).
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(OCC_WARPS * 32)
occ_backward_kernel(const float4 *__restrict__ rec, const uint8_t *__restrict__ visible,
                    const float *__restrict__ rs, const float *__restrict__ grad, int pix_stride,
                    int pix_offset, const int64_t *__restrict__ first_idx,
                    const int64_t *__restrict__ num_points, int64_t P0_shared, int S, int fast_smem_bytes,
                    float2 *__restrict__ grad_xy) {
    const int n = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const ViewRange vr = view_range(first_idx, num_points, n, P0_shared);
    const float r = rs[n];
    if (occ_tile_fits(r, S, fast_smem_bytes)) return;   // this view is handled by occ_tile_kernel
    const float r2 = r * r;
    const bool pow2 = (S & (S - 1)) == 0;
    const float inv_S = 1.0f / (float)S;
    const float half_S = 0.5f * (float)S;
    const float *gview = grad + ((int64_t)n * S * S) * pix_stride + pix_offset;
    constexpr unsigned FULL = 0xffffffffu;

    for (int64_t g0 = ((int64_t)blockIdx.x * OCC_WARPS + warp) * 32; g0 < vr.count;
         g0 += (int64_t)gridDim.x * OCC_WARPS * 32) {
        const int64_t i = g0 + lane;
        const bool in_range = i < vr.count;
        const int64_t p = vr.first + i;
        float4 A = make_float4(0.f, 0.f, -1.f, 0.f);
        float ry = 0.f;
        bool vis = false;
        if (in_range) {
            vis = visible[p] != 0;
            if (vis) {
                A = __ldg(&rec[2 * p]);
                ry = __ldg(&rec[2 * p + 1]).x;
                // rasterize_points_backward.cu:145 -- outside the renderable area
                if (A.z < 0.0f || fabsf(A.y) > 1.0f || fabsf(A.x) > 1.0f) vis = false;
            }
        }
        float out_x = 0.f, out_y = 0.f;
        unsigned todo = __ballot_sync(FULL, vis);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const float px = __shfl_sync(FULL, A.x, src);
            const float py = __shfl_sync(FULL, A.y, src);
            const float rx = __shfl_sync(FULL, A.w, src);
            const float ryb = __shfl_sync(FULL, ry, src);
            // conservative window in NDC-index space: pixel i has centre -1 + (2i+1)/S
            // (clamped in float first: saturating conversions of huge radii must not wrap)
            const float top = (float)(S - 1);
            const int xi_lo = (int)fminf(fmaxf(floorf((px - r + 1.0f) * half_S - 0.5f) - 1.0f, 0.0f), top + 1.0f);
            const int xi_hi = (int)fmaxf(fminf(ceilf((px + r + 1.0f) * half_S - 0.5f) + 1.0f, top), -1.0f);
            const int yi_lo = (int)fminf(fmaxf(floorf((py - r + 1.0f) * half_S - 0.5f) - 1.0f, 0.0f), top + 1.0f);
            const int yi_hi = (int)fmaxf(fminf(ceilf((py + r + 1.0f) * half_S - 0.5f) + 1.0f, top), -1.0f);
            const int W = xi_hi - xi_lo + 1, H = yi_hi - yi_lo + 1;
            float gx = 0.f, gy = 0.f;
            if (W > 0 && H > 0) {
                const int total = W * H;
                int wy = lane / W, wx = lane - wy * W;
                const int step_y = 32 / W, step_x = 32 - step_y * W;
                for (int w = lane; w < total; w += 32) {
                    const int xi = xi_lo + wx, yi = yi_lo + wy;
                    const float g = __ldg(gview + ((int64_t)(S - 1 - yi) * S + (S - 1 - xi)) * pix_stride);
                    if (g != 0.0f) {
                        const float xf = pix_to_ndc_fast(xi, S, inv_S, pow2);
                        const float yf = pix_to_ndc_fast(yi, S, inv_S, pow2);
                        const float dx = xf - px, dy = yf - py;
                        const float d2 = dx * dx + dy * dy;
                        const bool outside = (fabsf(dx) > rx) || (fabsf(dy) > ryb);
                        if (!(d2 > r2) && !(g > 0.0f && outside)) {
                            const float den = eps_denom(d2, 1e-10f);
                            gx += dx / den * g;
                            gy += dy / den * g;
                        }
                    }
                    wx += step_x;
                    wy += step_y;
                    if (wx >= W) {
                        wx -= W;
                        wy += 1;
                    }
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                gx += __shfl_xor_sync(FULL, gx, d);
                gy += __shfl_xor_sync(FULL, gy, d);
            }
            if (lane == src) {
                out_x = gx;
                out_y = gy;
            }
        }
        if (in_range) grad_xy[p] = make_float2(out_x, out_y);
    }
}

int occ_backward(dss_ctx *ctx, const float4 *rec, const uint8_t *visible, const float *rs,
                 const float *grad_occ, int pix_stride, int pix_offset, const int64_t *first_idx,
                 const int64_t *num_points, int N, int64_t P0, int S, float *grad_xy, cudaStream_t st) {
    if (N <= 0 || P0 <= 0) return DSS_OK;
    const int OB = (S + OCC_TILE - 1) / OCC_TILE;
    const int64_t nt = (int64_t)N * OB * OB;
    // shared-memory budget of the tile kernel from the radii seen by the previous call (a hint only:
    // both kernels re-derive "fits" from the current radius on the device)
    float hint = 0.0f;
    float *h_rs = reinterpret_cast<float *>(ctx->h_pinned + 8);
    for (int i = 0; i < (N < 96 ? N : 96); ++i) hint = fmaxf(hint, h_rs[i]);
    int smem = 64 * 1024;
    if (hint > 0.0f && hint < 4.0f) {
        const int side = OCC_TILE + 2 * (occ_halo(hint * 1.25f, S) + 1);
        smem = (2 * side * side + 2 * side) * (int)sizeof(float);
        if (smem < 16 * 1024) smem = 16 * 1024;
        if (smem > 200 * 1024) smem = 200 * 1024;
    }
    const bool tiles_ok = (size_t)OB * OB * sizeof(int32_t) <= 200 * 1024 && nt + 1 < (int64_t)INT32_MAX;
    if (!tiles_ok) smem = 0;
    int rc;
    StageScope prof(ctx, ST_OCC_BWD, st);
    if (smem > 0) {
        int32_t *counts = nullptr, *offsets = nullptr, *ids = nullptr;
        if ((rc = ctx_get(ctx, BUF_TILE_COUNTS, (size_t)(nt + 1), &counts))) return rc;
        if ((rc = ctx_get(ctx, BUF_TILE_OFFSETS, (size_t)(nt + 1), &offsets))) return rc;
        const int64_t Ptot = (first_idx == nullptr) ? (int64_t)N * P0 : P0;   // packed mode passes P0 = P
        if ((rc = ctx_get(ctx, BUF_TILE_IDS, (size_t)(Ptot > 0 ? Ptot : 1), &ids))) return rc;
        DSS_CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)(nt + 1) * sizeof(int32_t), st));
        dim3 bgrid((unsigned)((P0 + 2047) / 2048), N);
        const size_t hist = (size_t)OB * OB * sizeof(int32_t);
        if (hist > 48 * 1024) {
            DSS_CUDA_TRY(cudaFuncSetAttribute(occ_bin_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist));
            DSS_CUDA_TRY(cudaFuncSetAttribute(occ_bin_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist));
        }
        occ_bin_kernel<0><<<bgrid, 256, hist, st>>>(rec, visible, first_idx, num_points, P0, S, OB, counts, nullptr);
        DSS_LAUNCH_CHECK(ctx);
        if ((rc = exclusive_scan_i32(ctx, counts, offsets, nt + 1, st))) return rc;
        DSS_CUDA_TRY(cudaMemcpyAsync(counts, offsets, (size_t)nt * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        occ_bin_kernel<1><<<bgrid, 256, hist, st>>>(rec, visible, first_idx, num_points, P0, S, OB, counts, ids);
        DSS_LAUNCH_CHECK(ctx);
        DSS_CUDA_TRY(cudaMemsetAsync(grad_xy, 0, (size_t)Ptot * 2 * sizeof(float), st));
        if (smem > 48 * 1024)
            DSS_CUDA_TRY(cudaFuncSetAttribute(occ_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        dim3 tgrid((unsigned)(OB * OB), N, 32);
        occ_tile_kernel<<<tgrid, OCC_TILE_THREADS, smem, st>>>(rec, rs, grad_occ, pix_stride, pix_offset, offsets, ids,
                                                              S, OB, smem, reinterpret_cast<float2 *>(grad_xy));
        DSS_LAUNCH_CHECK(ctx);
    }
    // views whose window does not fit (very large search radius) take the direct global-memory gather
    dim3 grid(nblocks(P0, OCC_WARPS * 32, ctx->sm_count, 64), N);
    occ_backward_kernel<<<grid, OCC_WARPS * 32, 0, st>>>(rec, visible, rs, grad_occ, pix_stride, pix_offset,
                                                         first_idx, num_points, P0, S, smem,
                                                         reinterpret_cast<float2 *>(grad_xy));
    DSS_LAUNCH_CHECK(ctx);
    // refresh the hint for the next call (asynchronous; may be read stale, it is only a hint)
    DSS_CUDA_TRY(cudaMemcpyAsync(h_rs, rs, (size_t)(N < 96 ? N : 96) * sizeof(float), cudaMemcpyDeviceToHost, st));
    return DSS_OK;
}

// ---------------------------------------------------------------------------------------------
// z-buffer backward (rasterize_points.cu:823-846).  z_grad element stride in floats (1 for (P,),
// 3 to write the z column of a (P,3) gradient).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
zbuf_backward_kernel(const int32_t *__restrict__ idx, const float *__restrict__ grad_zbuf, int64_t num_pixels,
                     int K, float *z_grad, int z_stride) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_pixels;
         i += (int64_t)gridDim.x * blockDim.x) {
        for (int k = 0; k < K; ++k) {
            const float g = grad_zbuf[i * K + k];
            if (g == 0.0f) continue;
            const int p = idx[i * K + k];
            if (p < 0) break;
            atomicAdd(z_grad + (int64_t)p * z_stride, g);
        }
    }
}

int zbuf_backward(dss_ctx *ctx, const int32_t *idx, const float *grad_zbuf, int64_t num_pixels, int K,
                  float *z_grad, int z_stride, cudaStream_t st) {
    if (num_pixels == 0) return DSS_OK;
    StageScope prof(ctx, ST_ZBUF_BWD, st);
    zbuf_backward_kernel<<<nblocks(num_pixels, 256, ctx->sm_count, 16), 256, 0, st>>>(idx, grad_zbuf, num_pixels,
                                                                                      K, z_grad, z_stride);
    DSS_LAUNCH_CHECK(ctx);
    return DSS_OK;
}

// ---------------------------------------------------------------------------------------------
// Colour backward: dL/dcolour[idx_k] += g_rgb * w_k / max(sum w, 1e-4)  (norm_weighted_sum backward [ext]);
// `weights` already holds the normalised weights written by the forward pass.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colour_backward_kernel(const int32_t *__restrict__ idx, const float *__restrict__ weights,
                       const float4 *__restrict__ grad_image, int64_t num_pixels, int K,
                       float *grad_colours) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_pixels;
         i += (int64_t)gridDim.x * blockDim.x) {
        if (idx[i * K] < 0) continue;
        const float4 g = __ldg(&grad_image[i]);
        if (g.x == 0.0f && g.y == 0.0f && g.z == 0.0f) continue;
        for (int k = 0; k < K; ++k) {
            const int p = idx[i * K + k];
            if (p < 0) break;
            const float w = weights[i * K + k];
            float *dst = grad_colours + (int64_t)p * 3;
            atomicAdd(dst + 0, g.x * w);
            atomicAdd(dst + 1, g.y * w);
            atomicAdd(dst + 2, g.z * w);
        }
    }
}

int colour_backward(dss_ctx *ctx, const int32_t *idx, const float *weights, const float *grad_image,
                    int64_t num_pixels, int K, float *grad_colours, cudaStream_t st) {
    if (num_pixels == 0) return DSS_OK;
    StageScope prof(ctx, ST_COLOUR_BWD, st);
    colour_backward_kernel<<<nblocks(num_pixels, 256, ctx->sm_count, 16), 256, 0, st>>>(
        idx, weights, reinterpret_cast<const float4 *>(grad_image), num_pixels, K, grad_colours);
    DSS_LAUNCH_CHECK(ctx);
    return DSS_OK;
}

}  // namespace dss

extern "C" {

int dss_visibility_from_idx(dss_ctx *ctx, const int32_t *idx, int64_t num_pixels, int K, int64_t P,
                            uint8_t *visible, void *stream) {
    DSS_REQUIRE(ctx != nullptr, "ctx is null");
    DSS_REQUIRE(num_pixels >= 0 && K > 0 && P >= 0, "bad size");
    DSS_REQUIRE((num_pixels == 0 || idx) && (P == 0 || visible), "null pointer");
    return dss::visibility_from_idx(ctx, idx, num_pixels, K, P, visible, (cudaStream_t)stream);
}

int dss_search_radius(dss_ctx *ctx, const float *radii, const uint8_t *visible, const int64_t *first_idx,
                      const int64_t *num_points, int N, int64_t P, float radii_s, float *rs, void *stream) {
    DSS_REQUIRE(ctx != nullptr, "ctx is null");
    DSS_REQUIRE(N >= 0 && P >= 0, "negative size");
    if (N == 0) return DSS_OK;
    DSS_REQUIRE(rs && first_idx && num_points && (P == 0 || (radii && visible)), "null pointer");
    return dss::search_radius(ctx, nullptr, radii, visible, first_idx, num_points, N, P, radii_s, rs,
                              (cudaStream_t)stream);
}

int dss_occ_backward(dss_ctx *ctx, const float *points, const float *radii, const uint8_t *visible,
                     const float *rs, const float *grad_occ, int pix_stride, int pix_offset,
                     const int64_t *first_idx, const int64_t *num_points, int N, int64_t P, int image_size,
                     float *grad_xy, void *stream) {
    using namespace dss;
    cudaStream_t st = (cudaStream_t)stream;
    DSS_REQUIRE(ctx != nullptr, "ctx is null");
    DSS_REQUIRE(N >= 0 && P >= 0 && image_size > 0, "bad size");
    DSS_REQUIRE(pix_stride >= 1 && pix_offset >= 0 && pix_offset < pix_stride, "bad pixel stride/offset");
    if (N == 0 || P == 0) return DSS_OK;
    DSS_REQUIRE(points && radii && visible && rs && grad_occ && first_idx && num_points && grad_xy, "null pointer");
    float4 *rec = nullptr;
    int rc;
    if ((rc = ctx_get(ctx, BUF_RECORDS, (size_t)(2 * P), &rec))) return rc;
    if ((rc = pack_records(ctx, points, radii, nullptr, P, rec, st))) return rc;
    DSS_CUDA_TRY(cudaMemsetAsync(grad_xy, 0, (size_t)P * 2 * sizeof(float), st));
    return occ_backward(ctx, rec, visible, rs, grad_occ, pix_stride, pix_offset, first_idx, num_points, N, P,
                        image_size, grad_xy, st);
}

int dss_zbuf_backward(dss_ctx *ctx, const int32_t *idx, const float *grad_zbuf, int64_t num_pixels, int K,
                      float *z_grad, void *stream) {
    DSS_REQUIRE(ctx != nullptr, "ctx is null");
    DSS_REQUIRE(num_pixels >= 0 && K > 0, "bad size");
    if (num_pixels == 0) return DSS_OK;
    DSS_REQUIRE(idx && grad_zbuf && z_grad, "null pointer");
    return dss::zbuf_backward(ctx, idx, grad_zbuf, num_pixels, K, z_grad, 1, (cudaStream_t)stream);
}

}  // extern "C"
