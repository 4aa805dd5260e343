// kernels.cuh -- host-side entry points shared between the translation units of libdss_b200.
#pragma once
#include "common.cuh"

namespace dss {

constexpr int RASTER_TILE = 16;  // pixels per side of a raster tile (one 256-thread CTA)

// Small device -> host read-back WITHOUT the copy engine: a one-block kernel stores the words straight into pinned,
// device-mapped host memory.  A cudaMemcpyAsync would queue behind whatever large D2H copy another stream has in
// flight (copies of one direction are served in order) and, being in stream order, stall every kernel enqueued
// after it -- measured: ~1 ms per step in bench.py's end-to-end loop while the previous step's image is read back.
int publish_words(dss_ctx *ctx, const void *src_device, void *dst_pinned_host, int n_words, cudaStream_t st);

int exclusive_scan_i32(dss_ctx *ctx, const int32_t *in, int32_t *out, int64_t n, cudaStream_t st);

int pack_records(dss_ctx *ctx, const float *points, const float *radii, const float *ellipse, int64_t P,
                 float4 *rec, cudaStream_t st);

int choose_depth_slices(int B);
int init_zrange(dss_ctx *ctx, float *zrange, int N, cudaStream_t st);
int compute_zrange(dss_ctx *ctx, const float4 *rec, const int64_t *first_idx, const int64_t *num_points, int N,
                   int64_t P0, float *zrange, cudaStream_t st);

// rects (optional, one word per packed splat): the count pass stores every splat's tile rectangle + depth slice there and
// the scatter pass, given the same array, reuses them instead of deriving them again from the records
int bin_count_and_scan(dss_ctx *ctx, const float4 *rec, const int64_t *first_idx, const int64_t *num_points,
                       int N, int64_t P0, int S, int bin, int NS, const float *zrange, int32_t *counts,
                       int32_t *offsets, unsigned int *rects, cudaStream_t st);

int bin_scatter(dss_ctx *ctx, const float4 *rec, const int64_t *first_idx, const int64_t *num_points, int N,
                int64_t P0, int S, int bin, int NS, const float *zrange, const int32_t *offsets, int32_t *cursors,
                int32_t *ids, int64_t ids_capacity, const unsigned int *rects, cudaStream_t st);

// Depth slicing shared by the binning and raster kernels (identical arithmetic on both sides).
struct SliceMap {
    float zmin, dz, inv_dz;
    int NS;
};
__device__ __forceinline__ float slice_bound(const SliceMap &m, int s) { return fmaf((float)s, m.dz, m.zmin); }
__device__ __forceinline__ SliceMap make_slice_map(const float *zrange, int n, int NS) {
    SliceMap m;
    m.NS = NS;
    m.zmin = 0.f;
    m.dz = 0.f;
    m.inv_dz = 0.f;
    if (NS > 1 && zrange) {
        const float z0 = zrange[2 * n], z1 = zrange[2 * n + 1];
        if (z1 > z0) {
            m.zmin = z0;
            m.dz = (z1 - z0) / (float)NS;
            m.inv_dz = 1.0f / m.dz;
        }
    }
    return m;
}
// slice index with slice_bound(slice) <= z guaranteed (estimate + exact fix-up)
__device__ __forceinline__ int depth_slice(const SliceMap &m, float z) {
    if (m.NS <= 1 || !(m.dz > 0.f)) return 0;
    int s = (int)((z - m.zmin) * m.inv_dz);
    s = max(0, min(m.NS - 1, s));
    while (s > 0 && z < slice_bound(m, s)) --s;
    while (s < m.NS - 1 && z >= slice_bound(m, s + 1)) ++s;
    return s;
}

struct RasterArgs {
    const float4 *rec;        // 2 per splat
    const float *cutoff;      // per point (P,) or nullptr -> cutoff_uniform
    float cutoff_uniform;
    const int32_t *tile_offsets;  // (N*B*B + 1)
    const int32_t *tile_ids;      // CSR payload: packed splat ids
    const int32_t *tile_order;    // launch order of the tiles (longest lists first) or nullptr: block b = tile b
    int ids_capacity;             // entries tile_ids can hold; a tile whose list does not fit is rasterized from the
                                  // view's records directly (see bin_and_raster)
    const int64_t *first_idx;     // packed layout of the views (nullptr: shared cloud, view n = [n*P0, (n+1)*P0))
    const int64_t *num_points;
    int64_t P0;
    int N, S, K, B;
    int NS;                       // depth slices per tile list
    const float *zrange;          // (N,2) depth range per view (NS > 1)
    float depth_merge;
    // outputs
    int32_t *idx;
    float *zbuf;     // may be null
    float *qvalue;   // may be null
    float *occ;      // may be null (N,S,S)
    // blend (all null when not blending)
    const float *scaler;   // (P,)
    const float *colours;  // (P,3), or (P0,3) shared by every view when colour_P0 > 0
    int64_t colour_P0;     // > 0: colour row of packed splat id in view n is id - n * colour_P0
    float *image;          // (N,S,S,4)
    float *weights;        // (N,S,S,K)
    int32_t *cell_counts;  // optional (N*OB*OB*1024,), zeroed by the caller, OB = ceil(S/32): the blend epilogue counts every
                           // visible splat once in the 32x32-tile / pixel cell of its centre (the count pass of the occupancy
                           // backward's binning, fused); needs `visible` 4-byte aligned with capacity rounded up to 4
    uint8_t *visible;      // (P,) must be zeroed by the caller
    int64_t visible_count; // P (to re-zero `visible` when a pass has to be repeated)
    int force_pixel_parallel;  // testing: use the pixel-parallel kernel even for K <= 8
    unsigned long long *stats; // optional debug counters (dss_debug_raster_stats) or nullptr
};

int raster_forward(dss_ctx *ctx, const RasterArgs &a, cudaStream_t st);

// bin + rasterize, shared by dss_splat_points and dss_render_forward.  Synchronises once.
int bin_and_raster(dss_ctx *ctx, RasterArgs a, const int64_t *first_idx, const int64_t *num_points,
                   int64_t P0, cudaStream_t st);

int visibility_from_idx(dss_ctx *ctx, const int32_t *idx, int64_t num_pixels, int K, int64_t P,
                        uint8_t *visible, cudaStream_t st);
int search_radius(dss_ctx *ctx, const float4 *rec, const float *radii, const uint8_t *visible,
                  const int64_t *first_idx, const int64_t *num_points, int N, int64_t P0, float radii_s,
                  float *rs, cudaStream_t st);
// rs is an input when radii_s < 0, otherwise rs[n] = radii_s * lower median of the visible radii is computed here
// cell_counts (optional): per-cell counts of the visible splats produced by the forward (RasterArgs::cell_counts); when
// given and the cell-level binning is in use, the counting pass is replaced by a copy of these counts
int occ_backward(dss_ctx *ctx, const float4 *rec, const uint8_t *visible, float *rs, float radii_s,
                 const float *grad_occ, int pix_stride, int pix_offset, const int64_t *first_idx,
                 const int64_t *num_points, int N, int64_t P0, int S, float *grad_xy, const int32_t *cell_counts,
                 cudaStream_t st);
int zbuf_backward(dss_ctx *ctx, const int32_t *idx, const float *grad_zbuf, int64_t num_pixels, int K,
                  float *z_grad, int z_stride, cudaStream_t st);
// colour_P0 > 0: grad_colours is (P0,3) shared by the views (pixels_per_view = S*S locates a pixel's view)
int colour_backward(dss_ctx *ctx, const int32_t *idx, const float *weights, const float *grad_image,
                    int64_t num_pixels, int K, float *grad_colours, int64_t colour_P0, int64_t pixels_per_view,
                    cudaStream_t st);

}  // namespace dss
