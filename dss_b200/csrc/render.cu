// render.cu -- fused renderer path: per-(point,view) preprocess, forward orchestration, backward
// orchestration (gradient clip + chain to world space).
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

struct ShadeArgs {
    const float *albedo, *lights, *ambient, *cam;
    int n_lights, light_type;
    float shininess;
};

struct PreArgs {
    const float *pts, *nrm, *proj, *view, *h;
    const int64_t *first_idx, *num_points;
    int64_t P0;
    int shared_cloud, h_per_splat, S, backface;
    float cutoffC, sigma, znear, zfar;
    float4 *rec;
    float *ndc, *ellipse, *radii, *scaler;
    int32_t *zrange;   // (N,2) float bits: min / max view depth of the renderable splats, or null
    ShadeArgs sh;      // sh.albedo != nullptr: write the shaded colour of every (view, point) to `shaded`
    float *shaded;
};

// Fast reciprocal / division / square root (MUFU based, <= 2 ulp): the per-splat quantities are compared with the
// float64 oracle at 2e-5 .. 2e-3 relative, IEEE-exact division and sqrt were a third of this kernel's instructions.
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Python-side eps helpers of the reference (DSS/utils/mathHelper.py:10-22), zero counts as positive.
__device__ __forceinline__ float py_eps_denom(float d) { return eps_denom(d, 1e-17f); }
__device__ __forceinline__ float py_eps_sqrt(float s) { return fmaxf(fabsf(s), 1e-17f); }

// ---------------------------------------------------------------------------------------------
// Per-point shading: DSS/core/texture.py:74-127 (LightingTexture.forward:
// shaded = rgb * (ambient + diffuse) + specular) over DSS/core/lighting.py:10-172 (diffuse: Lambert with the
// renormalised normal and light direction, :62-69; specular: reflected ray against the view direction, masked where
// the light is behind the surface, :139-172).  Lights sit in shared memory as rows of 9 floats.
// ---------------------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 add3(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 sub3(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 mul3(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
// F.normalize(x, p=2, dim=-1, eps=1e-6): x / max(|x|, eps); returns the (clamped) norm through `len`
__device__ __forceinline__ V3 normalize3(V3 a, float &len) {
    len = fmaxf(sqrtf(dot3(a, a)), 1e-6f);
    return mul3(a, 1.0f / len);
}
// gradient of y = normalize(x) given dL/dy (for |x| > eps; below it the map is linear: g / eps)
__device__ __forceinline__ V3 normalize3_bwd(V3 y, float len, V3 gy) {
    if (len <= 1e-6f) return mul3(gy, 1.0f / len);
    return mul3(sub3(gy, mul3(y, dot3(y, gy))), 1.0f / len);
}

// forward (and, with gc != nullptr, backward) of the shading of one (view, point).  sL: lights in shared memory.
//   colour = albedo * (ambient + diffuse) + specular
// backward: accumulates d albedo, d normal (un-normalised input normal), d position into g_alb / g_nrm / g_pos.
template <bool BACKWARD>
__device__ __forceinline__ V3 shade_point(const float *sL, int n_lights, int light_type, float shininess, V3 amb, V3 p,
                                          V3 nrm, V3 alb, V3 cam, V3 gc, V3 &g_alb, V3 &g_nrm, V3 &g_pos) {
    float nlen, vlen;
    const V3 nh = normalize3(nrm, nlen);
    const V3 v = normalize3(sub3(cam, p), vlen);
    V3 diff = v3(0.f, 0.f, 0.f), spec = v3(0.f, 0.f, 0.f);
    V3 g_nh = v3(0.f, 0.f, 0.f), g_v = v3(0.f, 0.f, 0.f);
    for (int l = 0; l < n_lights; ++l) {
        const float *L = sL + l * 9;
        float dlen;
        const V3 draw = light_type ? sub3(v3(L[0], L[1], L[2]), p) : v3(L[0], L[1], L[2]);
        const V3 d = normalize3(draw, dlen);
        const V3 cd = v3(L[3], L[4], L[5]), cs = v3(L[6], L[7], L[8]);
        const float cosang = dot3(nh, d);
        const float lam = fmaxf(cosang, 0.0f);                           // lighting.py:65
        diff = add3(diff, mul3(cd, lam));
        const float mask = cosang > 0.0f ? 1.0f : 0.0f;                   // :158
        const V3 r = add3(mul3(d, -1.0f), mul3(nh, 2.0f * cosang));       // :163
        const float vr = dot3(v, r);
        const float alpha = fmaxf(vr, 0.0f) * mask;                       // :166-167
        const float pw = alpha > 0.0f ? powf(alpha, shininess) : 0.0f;    // :169
        spec = add3(spec, mul3(cs, pw));
        if (BACKWARD) {
            // diffuse: d/d cos = sum_c (gc*alb)_c cd_c  where cos > 0
            float g_cos = (cosang > 0.0f) ? (gc.x * alb.x * cd.x + gc.y * alb.y * cd.y + gc.z * alb.z * cd.z) : 0.0f;
            V3 g_d = v3(0.f, 0.f, 0.f);
            if (alpha > 0.0f) {
                const float g_alpha = dot3(gc, cs) * shininess * powf(alpha, shininess - 1.0f);
                const float g_vr = g_alpha;                                // alpha = vr where vr > 0 and mask = 1
                const V3 g_r = mul3(v, g_vr);
                g_v = add3(g_v, mul3(r, g_vr));
                g_cos += 2.0f * dot3(g_r, nh);
                g_nh = add3(g_nh, mul3(g_r, 2.0f * cosang));
                g_d = sub3(g_d, g_r);
            }
            g_nh = add3(g_nh, mul3(d, g_cos));
            g_d = add3(g_d, mul3(nh, g_cos));
            if (light_type) g_pos = sub3(g_pos, normalize3_bwd(d, dlen, g_d));   // d = normalize(loc - p)
        }
    }
    if (BACKWARD) {
        g_alb = add3(g_alb, v3(gc.x * (amb.x + diff.x), gc.y * (amb.y + diff.y), gc.z * (amb.z + diff.z)));
        g_nrm = add3(g_nrm, normalize3_bwd(nh, nlen, g_nh));
        g_pos = sub3(g_pos, normalize3_bwd(v, vlen, g_v));                // v = normalize(cam - p)
    }
    return v3(alb.x * (amb.x + diff.x) + spec.x, alb.y * (amb.y + diff.y) + spec.y, alb.z * (amb.z + diff.z) + spec.z);
}

// ---------------------------------------------------------------------------------------------
// One thread per (view, point).  Fuses DSS/core/rasterizer.py:183-217 (depth filter), :148-181
// (backface filter), :443-496 (_compute_WJk), :293-342 (global/isotropic Vrk with Sk^T Sk = I - n n^T),
// :404-441 (variance + detMk), :525-565 (conic, radii, scaler) and the pytorch3d transform of :614
// (x/t, y/t, view-space z) -- about 40 ATen launches incl. batched det/inverse in the reference.
// Filtered points keep their slot but get z = -1, which every later stage treats as "not renderable"
// exactly like the reference treats points behind the camera (rasterize_points.cu:87-88).
// ---------------------------------------------------------------------------------------------
template <bool SHADE>
__global__ void __launch_bounds__(256) preprocess_kernel(const __grid_constant__ PreArgs a) {
    __shared__ float sM[16], sV[16];
    __shared__ float sL[DSS_MAX_LIGHTS * 9 + 6];   // lights, ambient (3), camera centre (3)
    const int n = blockIdx.y;
    if (threadIdx.x < 16) {
        sM[threadIdx.x] = a.proj[n * 16 + threadIdx.x];
        sV[threadIdx.x] = a.view[n * 16 + threadIdx.x];
    }
    constexpr bool shade = SHADE;
    if (shade) {
        if (threadIdx.x < a.sh.n_lights * 9) sL[threadIdx.x] = a.sh.lights[threadIdx.x];
        if (threadIdx.x < 3) {
            sL[DSS_MAX_LIGHTS * 9 + threadIdx.x] = a.sh.ambient[threadIdx.x];
            sL[DSS_MAX_LIGHTS * 9 + 3 + threadIdx.x] = a.sh.cam[n * 3 + threadIdx.x];
        }
    }
    __syncthreads();
    const ViewRange vr = view_range(a.shared_cloud ? nullptr : a.first_idx, a.num_points, n, a.P0);
    const float pix = 2.0f / (float)a.S;
    const float aa = a.sigma * pix * pix;
    int zlo = 0x7f7fffff, zhi = 0;   // depth range of this thread's renderable splats (float bits)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < vr.count;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = vr.first + i;                     // packed slot
        const int64_t src = a.shared_cloud ? i : s;         // where the world-space point lives
        const float p0 = a.pts[src * 3 + 0], p1 = a.pts[src * 3 + 1], p2 = a.pts[src * 3 + 2];
        const float n0 = a.nrm[src * 3 + 0], n1 = a.nrm[src * 3 + 1], n2 = a.nrm[src * 3 + 2];
        // [p 1] . M  (row-vector convention, rasterizer.py:465-476)
        const float x = fmaf(p0, sM[0], fmaf(p1, sM[4], fmaf(p2, sM[8], sM[12])));
        const float y = fmaf(p0, sM[1], fmaf(p1, sM[5], fmaf(p2, sM[9], sM[13])));
        const float t = fmaf(p0, sM[3], fmaf(p1, sM[7], fmaf(p2, sM[11], sM[15])));
        float zv = fmaf(p0, sV[2], fmaf(p1, sV[6], fmaf(p2, sV[10], sV[14])));
        const float te = py_eps_denom(t);
        const float t2 = py_eps_denom(t * t);
        const float it = fast_rcp(te), it2 = fast_rcp(t2);
        // J = d(ndc xy)/d(world xyz), 3x2 (Mk = W @ Jk, rasterizer.py:483-494)
        float J0[3], J1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            J0[k] = sM[k * 4 + 0] * it - sM[k * 4 + 3] * x * it2;
            J1[k] = sM[k * 4 + 1] * it - sM[k * 4 + 3] * y * it2;
        }
        // T = J^T (I - n n^T) J
        const float nj0 = n0 * J0[0] + n1 * J0[1] + n2 * J0[2];
        const float nj1 = n0 * J1[0] + n1 * J1[1] + n2 * J1[2];
        const float T00 = (J0[0] * J0[0] + J0[1] * J0[1] + J0[2] * J0[2]) - nj0 * nj0;
        const float T01 = (J0[0] * J1[0] + J0[1] * J1[1] + J0[2] * J1[2]) - nj0 * nj1;
        const float T11 = (J1[0] * J1[0] + J1[1] * J1[1] + J1[2] * J1[2]) - nj1 * nj1;
        const float hh = a.h_per_splat ? a.h[s] : a.h[n];
        const float G00 = fmaf(hh, T00, aa), G01 = hh * T01, G11 = fmaf(hh, T11, aa);
        const float detG = G00 * G11 - G01 * G01;
        const float idet = fast_rcp(detG);
        const float ea = G11 * idet, eb = -2.0f * G01 * idet, ec = G00 * idet;   // rasterizer.py:543-551
        const float den = py_eps_denom(4.0f * ea * ec - eb * eb);                 // :509-519
        const float iden = fast_rcp(den);
        const float ry = fast_sqrt(py_eps_sqrt(4.0f * ea * a.cutoffC * iden));
        const float rx = fast_sqrt(py_eps_sqrt(4.0f * ec * a.cutoffC * iden));
        const float detT = fmaxf(T00 * T11 - T01 * T01, 0.0f);
        const float sc = fast_sqrt(detT) *
                         fast_rcp(py_eps_denom(fast_sqrt(py_eps_sqrt(detG * (4.0f * CUDART_PI_F * CUDART_PI_F)))));  // :558-559
        // filters (rasterizer.py:187-192, :152): view-space depth range, optional backface
        bool keep = (zv >= a.znear) && (zv <= a.zfar);
        if (a.backface) {
            const float nz = n0 * sV[2] + n1 * sV[6] + n2 * sV[10];
            keep = keep && (nz < 0.0f);
        }
        if (!keep) zv = -1.0f;
        if (zv >= 0.0f && zv < 3.0e38f) {
            const int zb = __float_as_int(zv + 0.0f);
            zlo = min(zlo, zb);
            zhi = max(zhi, zb);
        }
        const float irt = fast_rcp(t);
        const float xn = x * irt, yn = y * irt;
        a.rec[2 * s] = make_float4(xn, yn, zv, rx);
        a.rec[2 * s + 1] = make_float4(ry, ea, eb, ec);
        if (a.scaler) a.scaler[s] = sc;
        if (a.ndc) {
            a.ndc[s * 3 + 0] = xn;
            a.ndc[s * 3 + 1] = yn;
            a.ndc[s * 3 + 2] = zv;
        }
        if (a.ellipse) {
            a.ellipse[s * 3 + 0] = ea;
            a.ellipse[s * 3 + 1] = eb;
            a.ellipse[s * 3 + 2] = ec;
        }
        if (a.radii) {
            a.radii[s * 2 + 0] = rx;
            a.radii[s * 2 + 1] = ry;
        }
        if (shade) {
            V3 ga, gn, gp;
            const float *amb = sL + DSS_MAX_LIGHTS * 9;
            const V3 c = shade_point<false>(sL, a.sh.n_lights, a.sh.light_type, a.sh.shininess, v3(amb[0], amb[1], amb[2]),
                                            v3(p0, p1, p2), v3(n0, n1, n2),
                                            v3(a.sh.albedo[src * 3], a.sh.albedo[src * 3 + 1], a.sh.albedo[src * 3 + 2]),
                                            v3(amb[3], amb[4], amb[5]), v3(0.f, 0.f, 0.f), ga, gn, gp);
            a.shaded[s * 3 + 0] = c.x;
            a.shaded[s * 3 + 1] = c.y;
            a.shaded[s * 3 + 2] = c.z;
        }
    }
    if (a.zrange) {   // one pair of atomics per block
        __shared__ int s_lo[8], s_hi[8];
        zlo = __reduce_min_sync(0xffffffffu, zlo);
        zhi = __reduce_max_sync(0xffffffffu, zhi);
        if ((threadIdx.x & 31) == 0) {
            s_lo[threadIdx.x >> 5] = zlo;
            s_hi[threadIdx.x >> 5] = zhi;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) {
                zlo = min(zlo, s_lo[w]);
                zhi = max(zhi, s_hi[w]);
            }
            atomicMin(&a.zrange[2 * n], zlo);
            atomicMax(&a.zrange[2 * n + 1], zhi);
        }
    }
}

// d L / d shaded (N*P0,3) -> d albedo, d normal, d position (through the shading) per world point, summed over the views:
// one thread per point loops over the views (deterministic, no atomics) and re-evaluates the lighting of the
// (view, point) pairs that received a colour gradient (the visible ones, ~18 %).
struct ShadeBwdArgs {
    const float *pts, *nrm, *gsh;
    ShadeArgs sh;
    int64_t P0;
    int N;
    float *g_alb, *g_nrm, *g_pos;
};

__global__ void __launch_bounds__(256) shade_backward_kernel(const __grid_constant__ ShadeBwdArgs a) {
    extern __shared__ float sS[];   // lights (L*9), ambient (3), camera centres (N*3)
    const int nl9 = a.sh.n_lights * 9;
    for (int i = threadIdx.x; i < nl9; i += blockDim.x) sS[i] = a.sh.lights[i];
    for (int i = threadIdx.x; i < 3; i += blockDim.x) sS[nl9 + i] = a.sh.ambient[i];
    for (int i = threadIdx.x; i < a.N * 3; i += blockDim.x) sS[nl9 + 3 + i] = a.sh.cam[i];
    __syncthreads();
    const V3 amb = v3(sS[nl9], sS[nl9 + 1], sS[nl9 + 2]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.P0; i += (int64_t)gridDim.x * blockDim.x) {
        const V3 p = v3(a.pts[i * 3], a.pts[i * 3 + 1], a.pts[i * 3 + 2]);
        const V3 nrm = v3(a.nrm[i * 3], a.nrm[i * 3 + 1], a.nrm[i * 3 + 2]);
        const V3 alb = v3(a.sh.albedo[i * 3], a.sh.albedo[i * 3 + 1], a.sh.albedo[i * 3 + 2]);
        V3 ga = v3(0.f, 0.f, 0.f), gn = v3(0.f, 0.f, 0.f), gp = v3(0.f, 0.f, 0.f);
        for (int n = 0; n < a.N; ++n) {
            const float *g = a.gsh + ((int64_t)n * a.P0 + i) * 3;
            const V3 gc = v3(g[0], g[1], g[2]);
            if (gc.x == 0.0f && gc.y == 0.0f && gc.z == 0.0f) continue;
            const float *c = sS + nl9 + 3 + n * 3;
            shade_point<true>(sS, a.sh.n_lights, a.sh.light_type, a.sh.shininess, amb, p, nrm, alb, v3(c[0], c[1], c[2]), gc,
                              ga, gn, gp);
        }
        a.g_alb[i * 3] = ga.x, a.g_alb[i * 3 + 1] = ga.y, a.g_alb[i * 3 + 2] = ga.z;
        a.g_nrm[i * 3] = gn.x, a.g_nrm[i * 3 + 1] = gn.y, a.g_nrm[i * 3 + 2] = gn.z;
        a.g_pos[i * 3] = gp.x, a.g_pos[i * 3 + 1] = gp.y, a.g_pos[i * 3 + 2] = gp.z;
    }
}

// Fork/join of the context's side stream: once work has been forked, EVERY return path makes the caller's stream wait
// for it (an error return that skipped the join would let the caller free tensors the side stream still writes).
struct SideJoin {
    dss_ctx *ctx;
    cudaStream_t st;
    bool forked = false, joined = false;
    SideJoin(dss_ctx *c, cudaStream_t s) : ctx(c), st(s) {}
    int fork() {
        DSS_CUDA_TRY(cudaEventRecord(ctx->ev_fork, st));
        DSS_CUDA_TRY(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
        forked = true;
        return DSS_OK;
    }
    int join() {
        if (!forked || joined) return DSS_OK;
        joined = true;
        DSS_CUDA_TRY(cudaEventRecord(ctx->ev_join, ctx->side));
        DSS_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_join, 0));
        return DSS_OK;
    }
    ~SideJoin() {
        if (forked && !joined) {
            cudaEventRecord(ctx->ev_join, ctx->side);
            cudaStreamWaitEvent(st, ctx->ev_join, 0);
        }
    }
};

static int run_colour_backward(dss_ctx *ctx, const dss_render_args *g, cudaStream_t on) {
    const int N = g->n_views, S = g->image_size, K = g->points_per_pixel;
    const int64_t npix = (int64_t)N * S * S;
    const int64_t cP0 = (g->shared_cloud && g->shared_colours && !g->shade) ? g->P0 : 0;
    DSS_CUDA_TRY(cudaMemsetAsync(g->grad_colours, 0, (size_t)(cP0 > 0 ? cP0 : g->P) * 3 * sizeof(float), on));
    int rc = colour_backward(ctx, g->idx, g->weights, g->grad_image, npix, K, g->grad_colours, cP0, (int64_t)S * S, on);
    if (rc || !g->shade) return rc;
    // shading backward: d shaded (N*P0,3) -> d albedo, d normal, d position, per world point
    DSS_REQUIRE(g->grad_albedo && g->grad_normals_world && g->grad_points_shading && g->points_world,
                "shading backward needs grad_albedo, grad_normals_world, grad_points_shading");
    ShadeBwdArgs b;
    b.pts = g->points_world;
    b.nrm = g->normals_world;
    b.gsh = g->grad_colours;
    b.sh.albedo = g->albedo;
    b.sh.lights = g->lights;
    b.sh.ambient = g->ambient;
    b.sh.cam = g->cam_centres;
    b.sh.n_lights = g->n_lights;
    b.sh.light_type = g->light_type;
    b.sh.shininess = g->shininess;
    b.P0 = g->P0;
    b.N = N;
    b.g_alb = g->grad_albedo;
    b.g_nrm = g->grad_normals_world;
    b.g_pos = g->grad_points_shading;
    StageScope prof(ctx, ST_COLOUR_BWD, on);
    const size_t smem = (size_t)(g->n_lights * 9 + 3 + N * 3) * sizeof(float);
    shade_backward_kernel<<<nblocks(g->P0, 256, ctx->sm_count, 8), 256, smem, on>>>(b);
    DSS_LAUNCH_CHECK(ctx);
    return DSS_OK;
}

static int check_common(const dss_render_args *g) {
    DSS_REQUIRE(g != nullptr, "args is null");
    DSS_REQUIRE(g->n_views > 0, "n_views must be positive");
    DSS_REQUIRE(g->P0 >= 0 && g->P >= 0, "negative size");
    DSS_REQUIRE(g->P < (int64_t)INT32_MAX, "more than 2^31-1 packed points");
    DSS_REQUIRE(g->image_size > 0, "image_size must be positive");
    DSS_REQUIRE(g->points_per_pixel > 0 && g->points_per_pixel <= DSS_MAX_POINTS_PER_PIXEL,
                "points_per_pixel must be in [1, 64]");
    DSS_REQUIRE(g->shared_cloud || (g->first_idx && g->num_points), "first_idx/num_points required");
    DSS_REQUIRE(!g->shared_cloud || g->P == (int64_t)g->n_views * g->P0, "P != n_views * P0");
    DSS_REQUIRE(!g->shared_cloud || g->n_views <= DSS_MAX_SHARED_VIEWS, "a shared cloud takes at most 256 views per call");
    if (g->shade) {
        DSS_REQUIRE(g->shared_cloud, "fused shading needs a shared cloud");
        DSS_REQUIRE(g->n_lights >= 1 && g->n_lights <= DSS_MAX_LIGHTS, "n_lights must be in [1, 8]");
        DSS_REQUIRE(g->albedo && g->lights && g->ambient && g->cam_centres && g->normals_world, "shading needs albedo, lights, ambient, cam_centres, normals");
        DSS_REQUIRE(g->light_type == 0 || g->light_type == 1, "light_type must be 0 (directional) or 1 (point)");
    }
    return DSS_OK;
}

static int run_preprocess(dss_ctx *ctx, const dss_render_args *g, float4 *rec, cudaStream_t st, float *zrange = nullptr) {
    if (zrange) {
        int rc = init_zrange(ctx, zrange, g->n_views, st);
        if (rc) return rc;
    }
    if (g->P == 0 || g->P0 == 0) return DSS_OK;
    DSS_REQUIRE(g->points_world && g->normals_world && g->proj && g->view && g->h, "null input array");
    PreArgs a;
    a.pts = g->points_world;
    a.nrm = g->normals_world;
    a.proj = g->proj;
    a.view = g->view;
    a.h = g->h;
    a.first_idx = g->first_idx;
    a.num_points = g->num_points;
    a.P0 = g->P0;
    a.shared_cloud = g->shared_cloud;
    a.h_per_splat = g->h_per_splat;
    a.S = g->image_size;
    a.backface = g->backface_culling;
    a.cutoffC = g->cutoff_threshold;
    a.sigma = g->antialiasing_sigma;
    a.znear = g->znear;
    a.zfar = g->zfar;
    a.rec = rec;
    a.ndc = g->ndc;
    a.ellipse = g->ellipse;
    a.radii = g->radii;
    a.scaler = g->scaler;
    a.zrange = reinterpret_cast<int32_t *>(zrange);
    memset(&a.sh, 0, sizeof(a.sh));
    a.shaded = nullptr;
    if (g->shade && g->shaded) {
        a.sh.albedo = g->albedo;
        a.sh.lights = g->lights;
        a.sh.ambient = g->ambient;
        a.sh.cam = g->cam_centres;
        a.sh.n_lights = g->n_lights;
        a.sh.light_type = g->light_type;
        a.sh.shininess = g->shininess;
        a.shaded = g->shaded;
    }
    dim3 grid(nblocks(g->P0, 256, ctx->sm_count, 8), g->n_views);
    StageScope prof(ctx, ST_PREPROCESS, st);
    if (a.shaded) preprocess_kernel<true><<<grid, 256, 0, st>>>(a);
    else preprocess_kernel<false><<<grid, 256, 0, st>>>(a);
    DSS_LAUNCH_CHECK(ctx);
    return DSS_OK;
}

// ---------------------------------------------------------------------------------------------
// Backward tail: per-point gradient clip (rasterizer.py:667-673) and chain through
// ndc = (X/T, Y/T, z_view) to world space: d ndc_xy / d p = J (the same Jacobian as the forward
// pass), d z_view / d p = V[:3, 2].  shared_cloud: one thread per world point loops over the views
// and writes the sum (deterministic, no atomics); otherwise one thread per packed splat.
// ---------------------------------------------------------------------------------------------
struct ChainArgs {
    const float *gshade;   // (P0,3) position gradient through the shading, added to the result (shared cloud), or null
    const float *pts, *proj, *view;
    const float2 *gxy;     // (P,2) occupancy gradient
    const float *gz;       // (P,) z gradient or null
    const int64_t *first_idx, *num_points;
    int64_t P0;
    int N, shared_cloud;
    float clip;
    float *grad_ndc;       // (P,3) or null
    float *grad_world;     // (P0,3) or (P,3)
};

__device__ __forceinline__ void chain_one(const float *M, const float *V, float p0, float p1, float p2,
                                          float gx, float gy, float gz, float &w0, float &w1, float &w2) {
    const float x = fmaf(p0, M[0], fmaf(p1, M[4], fmaf(p2, M[8], M[12])));
    const float y = fmaf(p0, M[1], fmaf(p1, M[5], fmaf(p2, M[9], M[13])));
    const float t = fmaf(p0, M[3], fmaf(p1, M[7], fmaf(p2, M[11], M[15])));
    const float it = 1.0f / t, it2 = it * it;
    float w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float j0 = M[k * 4 + 0] * it - M[k * 4 + 3] * x * it2;
        const float j1 = M[k * 4 + 1] * it - M[k * 4 + 3] * y * it2;
        w[k] = j0 * gx + j1 * gy + V[k * 4 + 2] * gz;
    }
    w0 = w[0];
    w1 = w[1];
    w2 = w[2];
}

__device__ __forceinline__ void clip_grad(float clip, float &gx, float &gy, float &gz) {
    if (clip > 0.0f) {
        // grad.norm().clamp(0, clip) * normalize(grad)  (F.normalize eps = 1e-12)
        const float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
        const float s = fminf(nrm, clip) / fmaxf(nrm, 1e-12f);
        gx *= s;
        gy *= s;
        gz *= s;
    }
}

__global__ void __launch_bounds__(256) chain_kernel(const __grid_constant__ ChainArgs a) {
    extern __shared__ float sMat[];  // N * 32 floats: proj then view per view (shared_cloud) or 32 (packed)
    if (a.shared_cloud) {
        for (int i = threadIdx.x; i < a.N * 16; i += blockDim.x) {
            sMat[(i / 16) * 32 + (i % 16)] = a.proj[i];
            sMat[(i / 16) * 32 + 16 + (i % 16)] = a.view[i];
        }
        __syncthreads();
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.P0;
             i += (int64_t)gridDim.x * blockDim.x) {
            const float p0 = a.pts[i * 3], p1 = a.pts[i * 3 + 1], p2 = a.pts[i * 3 + 2];
            float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
            // views in batches of four: the (independent) gradient loads of a batch are issued together
            for (int n0 = 0; n0 < a.N; n0 += 4) {
                float2 g4[4];
                float gz4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t s = (int64_t)(n0 + j) * a.P0 + i;
                    g4[j] = (n0 + j < a.N) ? a.gxy[s] : make_float2(0.f, 0.f);
                    gz4[j] = (a.gz && n0 + j < a.N) ? a.gz[s] : 0.0f;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = n0 + j;
                    if (n >= a.N) break;
                    const int64_t s = (int64_t)n * a.P0 + i;
                    float gx = g4[j].x, gy = g4[j].y, gz = gz4[j];
                    clip_grad(a.clip, gx, gy, gz);
                    if (a.grad_ndc) {
                        a.grad_ndc[s * 3 + 0] = gx;
                        a.grad_ndc[s * 3 + 1] = gy;
                        a.grad_ndc[s * 3 + 2] = gz;
                    }
                    if (gx != 0.0f || gy != 0.0f || gz != 0.0f) {
                        float w0, w1, w2;
                        chain_one(sMat + n * 32, sMat + n * 32 + 16, p0, p1, p2, gx, gy, gz, w0, w1, w2);
                        acc0 += w0;
                        acc1 += w1;
                        acc2 += w2;
                    }
                }
            }
            if (a.gshade) {
                acc0 += a.gshade[i * 3 + 0];
                acc1 += a.gshade[i * 3 + 1];
                acc2 += a.gshade[i * 3 + 2];
            }
            a.grad_world[i * 3 + 0] = acc0;
            a.grad_world[i * 3 + 1] = acc1;
            a.grad_world[i * 3 + 2] = acc2;
        }
    } else {
        const int n = blockIdx.y;
        if (threadIdx.x < 16) {
            sMat[threadIdx.x] = a.proj[n * 16 + threadIdx.x];
            sMat[16 + threadIdx.x] = a.view[n * 16 + threadIdx.x];
        }
        __syncthreads();
        const ViewRange vr = view_range(a.first_idx, a.num_points, n, a.P0);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < vr.count;
             i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t s = vr.first + i;
            const float2 g = a.gxy[s];
            float gx = g.x, gy = g.y, gz = a.gz ? a.gz[s] : 0.0f;
            clip_grad(a.clip, gx, gy, gz);
            if (a.grad_ndc) {
                a.grad_ndc[s * 3 + 0] = gx;
                a.grad_ndc[s * 3 + 1] = gy;
                a.grad_ndc[s * 3 + 2] = gz;
            }
            float w0 = 0.f, w1 = 0.f, w2 = 0.f;
            if (gx != 0.0f || gy != 0.0f || gz != 0.0f)
                chain_one(sMat, sMat + 16, a.pts[s * 3], a.pts[s * 3 + 1], a.pts[s * 3 + 2], gx, gy, gz, w0, w1, w2);
            a.grad_world[s * 3 + 0] = w0;
            a.grad_world[s * 3 + 1] = w1;
            a.grad_world[s * 3 + 2] = w2;
        }
    }
}

}  // namespace dss

extern "C" {

int dss_preprocess(dss_ctx *ctx, const dss_render_args *g, void *stream) {
    using namespace dss;
    DSS_REQUIRE(ctx != nullptr, "ctx is null");
    int rc = check_common(g);
    if (rc) return rc;
    float4 *rec = reinterpret_cast<float4 *>(g->records);
    if (!rec && (rc = ctx_get(ctx, BUF_RECORDS, (size_t)(2 * (g->P > 0 ? g->P : 1)), &rec))) return rc;
    return run_preprocess(ctx, g, rec, (cudaStream_t)stream);
}

int dss_render_forward(dss_ctx *ctx, const dss_render_args *g, void *stream) {
    using namespace dss;
    cudaStream_t st = (cudaStream_t)stream;
    DSS_REQUIRE(ctx != nullptr, "ctx is null");
    int rc = check_common(g);
    if (rc) return rc;
    DSS_REQUIRE(g->image && g->idx && g->scaler && (g->shade ? (const void *)g->shaded : (const void *)g->colours),
                "forward needs image, idx, scaler and colours (or, with shade, the shaded output)");
    float4 *rec = reinterpret_cast<float4 *>(g->records);
    if (!rec && (rc = ctx_get(ctx, BUF_RECORDS, (size_t)(2 * (g->P > 0 ? g->P : 1)), &rec))) return rc;
    DSS_REQUIRE((reinterpret_cast<uintptr_t>(rec) & 15) == 0, "records must be 16-byte aligned");
    float *zrange = nullptr;
    if ((rc = ctx_get(ctx, BUF_ZRANGE, (size_t)(2 * g->n_views), &zrange))) return rc;
    if ((rc = run_preprocess(ctx, g, rec, st, zrange))) return rc;
    if (g->visible) DSS_CUDA_TRY(cudaMemsetAsync(g->visible, 0, (size_t)g->P, st));
    // fused count pass of the backward's binning (see RasterArgs::cell_counts): cells = pixels of 32x32 tiles
    int32_t *cell_counts = nullptr;
    if (g->cell_counts && g->visible && g->points_per_pixel <= 8 && (reinterpret_cast<uintptr_t>(g->visible) & 3) == 0) {
        const int64_t OB = (g->image_size + 31) / 32;
        cell_counts = g->cell_counts;
        DSS_CUDA_TRY(cudaMemsetAsync(cell_counts, 0, (size_t)g->n_views * OB * OB * 1024 * sizeof(int32_t), st));
    }
    RasterArgs a;
    memset(&a, 0, sizeof(a));
    a.rec = rec;
    a.cutoff = nullptr;
    a.cutoff_uniform = g->cutoff_threshold;
    a.N = g->n_views;
    a.S = g->image_size;
    a.K = g->points_per_pixel;
    a.depth_merge = g->depth_merging_threshold;
    a.idx = g->idx;
    a.zbuf = g->zbuf;
    a.qvalue = g->qvalue;
    a.occ = nullptr;
    a.scaler = g->scaler;
    a.colours = g->shade ? g->shaded : g->colours;
    a.colour_P0 = (g->shared_cloud && g->shared_colours && !g->shade) ? g->P0 : 0;
    a.image = g->image;
    a.weights = g->weights;
    a.visible = g->visible;
    a.cell_counts = cell_counts;
    a.visible_count = g->P;
    a.zrange = zrange;   // already filled by the preprocess kernel
    return bin_and_raster(ctx, a, g->shared_cloud ? nullptr : g->first_idx, g->num_points, g->P0, st);
}

int dss_colour_backward(dss_ctx *ctx, const dss_render_args *g, void *stream) {
    using namespace dss;
    DSS_REQUIRE(ctx != nullptr, "ctx is null");
    int rc = check_common(g);
    if (rc) return rc;
    DSS_REQUIRE(g->grad_image && g->idx && g->weights && g->grad_colours, "colour backward needs grad_image, idx, weights, grad_colours");
    if (g->P == 0) return DSS_OK;
    return run_colour_backward(ctx, g, (cudaStream_t)stream);
}

int dss_render_backward(dss_ctx *ctx, const dss_render_args *g, void *stream) {
    using namespace dss;
    cudaStream_t st = (cudaStream_t)stream;
    DSS_REQUIRE(ctx != nullptr, "ctx is null");
    int rc = check_common(g);
    if (rc) return rc;
    DSS_REQUIRE(g->grad_image && g->idx && g->weights && g->visible, "backward needs grad_image, idx, weights, visible");
    DSS_REQUIRE(g->records || (g->ndc && g->radii), "backward needs the records (or ndc and radii) saved by the forward pass");
    DSS_REQUIRE(g->grad_points_world && g->points_world && g->proj && g->view, "null pointer");
    const int N = g->n_views, S = g->image_size, K = g->points_per_pixel;
    const int64_t P = g->P, npix = (int64_t)N * S * S;
    const int64_t *fi = g->shared_cloud ? nullptr : g->first_idx;
    if (P == 0) return DSS_OK;
    float4 *rec = reinterpret_cast<float4 *>(g->records);
    float *rs = nullptr, *gxy = nullptr;
    if (!rec) {
        // rebuild from the tensors the caller saved (another forward may have reused the scratch)
        if ((rc = ctx_get(ctx, BUF_RECORDS, (size_t)(2 * P), &rec))) return rc;
        if ((rc = pack_records(ctx, g->ndc, g->radii, nullptr, P, rec, st))) return rc;
    }
    if ((rc = ctx_get(ctx, BUF_RS, (size_t)N, &rs))) return rc;
    if ((rc = ctx_get(ctx, BUF_GRADXY, (size_t)(2 * P), &gxy))) return rc;
    // The colour scatter does not depend on the occupancy path: fork it onto the context's side stream so that its
    // atomics overlap the (latency-bound) binning / median kernels, join before the chain kernel (and on every return).
    SideJoin side(ctx, st);
    if (g->grad_colours != nullptr) {
        if ((rc = side.fork())) return rc;
        if ((rc = run_colour_backward(ctx, g, ctx->side))) return rc;
    }
    if ((rc = occ_backward(ctx, rec, g->visible, rs, g->radii_backward_scaler, g->grad_image, 4, 3, fi, g->num_points, N,
                           g->P0, S, gxy,
                           (g->cell_counts && K <= 8 && (reinterpret_cast<uintptr_t>(g->visible) & 3) == 0) ? g->cell_counts : nullptr,
                           st)))
        return rc;
    if (g->search_radius)
        DSS_CUDA_TRY(cudaMemcpyAsync(g->search_radius, rs, (size_t)N * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if ((rc = side.join())) return rc;
    float *gz = nullptr;
    if (g->grad_zbuf) {
        if ((rc = ctx_get(ctx, BUF_MISC, (size_t)P, &gz))) return rc;
        DSS_CUDA_TRY(cudaMemsetAsync(gz, 0, (size_t)P * sizeof(float), st));
        if ((rc = zbuf_backward(ctx, g->idx, g->grad_zbuf, npix, K, gz, 1, st))) return rc;
    }
    ChainArgs c;
    c.gshade = (g->shade && g->grad_colours != nullptr) ? g->grad_points_shading : nullptr;
    c.pts = g->points_world;
    c.proj = g->proj;
    c.view = g->view;
    c.gxy = reinterpret_cast<const float2 *>(gxy);
    c.gz = gz;
    c.first_idx = fi;
    c.num_points = g->num_points;
    c.P0 = g->P0;
    c.N = N;
    c.shared_cloud = g->shared_cloud;
    c.clip = g->clip_pts_grad;
    c.grad_ndc = g->grad_ndc;
    c.grad_world = g->grad_points_world;
    StageScope prof(ctx, ST_CHAIN, st);
    if (g->shared_cloud) {
        chain_kernel<<<nblocks(g->P0, 256, ctx->sm_count, 8), 256, (size_t)N * 32 * sizeof(float), st>>>(c);
    } else {
        dim3 grid(nblocks(g->P0, 256, ctx->sm_count, 8), N);
        chain_kernel<<<grid, 256, 32 * sizeof(float), st>>>(c);
    }
    DSS_LAUNCH_CHECK(ctx);
    return DSS_OK;
}

}  // extern "C"
