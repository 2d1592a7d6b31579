/* m2s_oracle.c — CPU restatement of the mesh2splat conversion pass.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mesh2splat_b200/ may include, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the timed CPU baseline.
 *
 * What is restated, and from where (paths relative to the reference tree):
 *   per-triangle stage   src/shaders/conversion/converterGS.glsl:326-443 (main),
 *                        :131-183 (quat_cast), :206-235 (inverse2x2, mat mult), :269-300 (Jacobian)
 *   per-fragment stage   src/shaders/conversion/converterFS.glsl:44-104
 *   host state           src/renderer/renderPasses/ConversionPass.cpp:17-59,77-116
 *   sampler state        src/utils/glUtils.cpp:292-313 (RGBA8, REPEAT, trilinear, max level 4)
 *   bbox rule            src/utils/SceneManager.cpp:476-477,514-520,527
 *   export               src/utils/SceneManager.cpp:668, src/parsers/parsers.cpp:232-316,339-428,
 *                        431-514, src/utils/utils.cpp:45-49, src/utils/utils.hpp:270
 * The fixed-function parts the reference leaves to the GL driver are restated from the
 * OpenGL 4.6 core spec: 14.6.1 (polygon rasterisation: pixel-centre sampling, barycentric
 * interpolation, one fragment per shared-edge sample -> top-left rule, 8 sub-pixel bits) and
 * 8.14 (texture minification: rho, lambda = log2 rho, level clamp, trilinear weights).
 *
 * Pinning: the per-triangle and per-fragment arithmetic is checked against the reference's own
 * GLSL compiled as C++ (oracle/_ref, see build_ref.py) and against the analytic KATs of
 * SURVEY.md 8c; the rasteriser/sampler have no reference executable to be checked against
 * (driver behaviour), so that part of parity is unpinned.
 *
 * Arithmetic: fp32, one rounding per operation (build with -ffp-contract=off), operation order
 * of GLM's length/normalize/cross/dot so that oracle/_ref (GLSL through GLM) agrees bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/m2s.h"

#define ORC_API __attribute__((visibility("default")))
#define SH_C0 0.28209479177387814f /* params.hpp:17 */

typedef struct { float x, y, z; } v3;

static inline v3 v3sub(v3 a, v3 b) { v3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static inline v3 v3scale(v3 a, float s) { v3 r = {a.x * s, a.y * s, a.z * s}; return r; }
/* glm::dot<vec3>: (x*x + y*y) + z*z */
static inline float v3dot(v3 a, v3 b) { float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return tx + ty + tz; }
static inline float v3len(v3 a) { return sqrtf(v3dot(a, a)); }
/* glm::normalize: v * (1 / sqrt(dot(v,v))) */
static inline v3 v3norm(v3 a) { float inv = 1.0f / sqrtf(v3dot(a, a)); return v3scale(a, inv); }
static inline v3 v3cross(v3 x, v3 y) {
    v3 r = {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y};
    return r;
}

/* ------------------------------------------------------------------------------------------ */
/* per-triangle stage                                                                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct orc_setup {
    float ouv[3][2];    /* orthogonal uv, converterGS.glsl:354-397 */
    float quat[4];      /* (w,x,y,z), :407 */
    float scale[3];     /* raw (|Ju|,|Jv|,1e-7), :420-430 */
    int32_t X[3], Y[3]; /* window coords, 8 sub-pixel bits */
    int64_t area2;      /* signed doubled area in fixed point; 0 => no fragments */
    int32_t axis;       /* 0 X-dominant, 1 Y, 2 Z */
    int32_t valid;      /* 0 if a coordinate is NaN/inf (never rasterised) */
    /* sign-normalised edge functions at pixel (i,j): E_k = A_k i + B_k j + C_k >= 0 inside */
    int64_t A[3], B[3], C[3];
    int32_t incl[3];    /* edge owns its E==0 samples (top-left rule) */
    int32_t x0, y0, x1, y1; /* inclusive candidate pixel box, empty if x1<x0 or y1<y0 */
} orc_setup;

/* converterGS.glsl:131-183, m[col][row] */
static void quat_cast_wxyz(const v3 c0, const v3 c1, const v3 c2, float out_wxyz[4]) {
    const float m00 = c0.x, m01 = c0.y, m02 = c0.z;
    const float m10 = c1.x, m11 = c1.y, m12 = c1.z;
    const float m20 = c2.x, m21 = c2.y, m22 = c2.z;
    float fourX = m00 - m11 - m22;
    float fourY = m11 - m00 - m22;
    float fourZ = m22 - m00 - m11;
    float fourW = m00 + m11 + m22;
    int big = 0;
    float fourBig = fourW;
    if (fourX > fourBig) { fourBig = fourX; big = 1; }
    if (fourY > fourBig) { fourBig = fourY; big = 2; }
    if (fourZ > fourBig) { fourBig = fourZ; big = 3; }
    float bigVal = sqrtf(fourBig + 1.0f) * 0.5f;
    float mult = 0.25f / bigVal;
    float w, x, y, z;
    if (big == 0) {
        w = bigVal; x = (m12 - m21) * mult; y = (m20 - m02) * mult; z = (m01 - m10) * mult;
    } else if (big == 1) {
        w = (m12 - m21) * mult; x = bigVal; y = (m01 + m10) * mult; z = (m20 + m02) * mult;
    } else if (big == 2) {
        w = (m20 - m02) * mult; x = (m01 + m10) * mult; y = bigVal; z = (m12 + m21) * mult;
    } else {
        w = (m01 - m10) * mult; x = (m20 + m02) * mult; y = (m12 + m21) * mult; z = bigVal;
    }
    out_wxyz[0] = w; out_wxyz[1] = x; out_wxyz[2] = y; out_wxyz[3] = z;
}

static inline int64_t floor_div256(int64_t v) { return v >> 8; } /* arithmetic shift = floor */

/* tri: 3 x {pos3 nrm3 tan4 uv2}. Returns 1 if the triangle can produce fragments. */
ORC_API int orc_triangle_setup(const float* tri, const float bmin[3], const float bmax[3],
                               uint32_t R, orc_setup* s) {
    memset(s, 0, sizeof(*s));
    v3 P[3];
    for (int k = 0; k < 3; ++k) { P[k].x = tri[12 * k]; P[k].y = tri[12 * k + 1]; P[k].z = tri[12 * k + 2]; }

    /* converterGS.glsl:327-342 — longest edge first (strict >) */
    v3 e1 = v3sub(P[1], P[0]), e2 = v3sub(P[2], P[0]), e3 = v3sub(P[2], P[1]);
    float l1 = v3len(e1), l2 = v3len(e2), l3 = v3len(e3);
    if (l2 > l1 && l2 > l3) { v3 t = e1; e1 = e2; e2 = t; }
    else if (l3 > l1 && l3 > l2) { v3 t = e1; e1 = e3; e3 = t; }
    (void)e3;
    e1 = v3norm(e1);                       /* :345 */
    v3 n = v3norm(v3cross(e1, e2));        /* :347 */
    float ax = fabsf(n.x), ay = fabsf(n.y), az = fabsf(n.z);

    /* :354-397 — dominant-axis orthographic projection, normalised by the larger box side */
    int axis = (ax > ay && ax > az) ? 0 : ((ay > az) ? 1 : 2);
    s->axis = axis;
    for (int k = 0; k < 3; ++k) {
        float u, v;
        if (axis == 0) {
            float rY = bmax[1] - bmin[1], rZ = bmax[2] - bmin[2];
            float range = rY > rZ ? rY : rZ; /* glm::max(a,b) = (a<b)?b:a */
            if (rY < rZ) range = rZ; else range = rY;
            u = (P[k].y - bmin[1]) / range; v = (P[k].z - bmin[2]) / range;
        } else if (axis == 1) {
            float rX = bmax[0] - bmin[0], rZ = bmax[2] - bmin[2];
            float range; if (rX < rZ) range = rZ; else range = rX;
            u = (P[k].x - bmin[0]) / range; v = (P[k].z - bmin[2]) / range;
        } else {
            float rX = bmax[0] - bmin[0], rY = bmax[1] - bmin[1];
            float range; if (rX < rY) range = rY; else range = rX;
            u = (P[k].x - bmin[0]) / range; v = (P[k].y - bmin[1]) / range;
        }
        s->ouv[k][0] = u; s->ouv[k][1] = v;
    }

    /* :399-407 — rotation (longest edge, n x e, n) -> quaternion (w,x,y,z) */
    v3 xA = e1, yA = v3norm(v3cross(n, xA)), zA = n;
    quat_cast_wxyz(xA, yA, zA, s->quat);

    /* :269-300,206-235,414-430 — J = V * inverse(UV); scale = (|J col0|, |J col1|, 1e-7) */
    {
        float a = s->ouv[1][0] - s->ouv[0][0]; /* UV[0][0] */
        float b = s->ouv[2][0] - s->ouv[0][0]; /* UV[1][0] */
        float c = s->ouv[1][1] - s->ouv[0][1]; /* UV[0][1] */
        float d = s->ouv[2][1] - s->ouv[0][1]; /* UV[1][1] */
        float det = a * d - c * b;
        float i00, i10, i01, i11;
        if (det == 0.0f) { i00 = i10 = i01 = i11 = 0.0f; }
        else {
            float invDet = 1.0f / det;
            i00 = d * invDet; i10 = -b * invDet; i01 = -c * invDet; i11 = a * invDet;
        }
        v3 V0 = v3sub(P[1], P[0]), V1 = v3sub(P[2], P[0]);
        v3 Ju = {V0.x * i00 + V1.x * i01, V0.y * i00 + V1.y * i01, V0.z * i00 + V1.z * i01};
        v3 Jv = {V0.x * i10 + V1.x * i11, V0.y * i10 + V1.y * i11, V0.z * i10 + V1.z * i11};
        s->scale[0] = v3len(Ju); s->scale[1] = v3len(Jv); s->scale[2] = 1e-7f;
    }

    /* Rasteriser set-up (GL 4.6 14.6.1 / 13.8.1): gl_Position = ouv*2-1 (:439), viewport
     * (0,0,R,R) (ConversionPass.cpp:45): xw = ndc*R/2 + R/2, snapped to 1/256 pixel. */
    const float half = (float)R * 0.5f;
    s->valid = 1;
    for (int k = 0; k < 3; ++k) {
        float ndx = s->ouv[k][0] * 2.0f - 1.0f, ndy = s->ouv[k][1] * 2.0f - 1.0f;
        float xw = ndx * half + half, yw = ndy * half + half;
        if (!isfinite(xw) || !isfinite(yw) || fabsf(xw) > 8192.0f || fabsf(yw) > 8192.0f) { s->valid = 0; xw = yw = 0.0f; }
        s->X[k] = (int32_t)lrintf(xw * 256.0f);
        s->Y[k] = (int32_t)lrintf(yw * 256.0f);
    }
    if (!s->valid) { s->area2 = 0; s->x1 = -1; s->y1 = -1; return 0; }
    int64_t area2 = (int64_t)(s->X[1] - s->X[0]) * (s->Y[2] - s->Y[0]) - (int64_t)(s->X[2] - s->X[0]) * (s->Y[1] - s->Y[0]);
    s->area2 = area2;
    if (area2 == 0) { s->x1 = -1; s->y1 = -1; return 0; }
    const int64_t sg = area2 < 0 ? -1 : 1;
    for (int k = 0; k < 3; ++k) {
        int a = (k + 1) % 3, b = (k + 2) % 3;
        int64_t dx = s->X[b] - s->X[a], dy = s->Y[b] - s->Y[a];
        s->A[k] = sg * (-dy * 256);
        s->B[k] = sg * (dx * 256);
        s->C[k] = sg * (dx * (128 - (int64_t)s->Y[a]) - dy * (128 - (int64_t)s->X[a]));
        s->incl[k] = (s->A[k] > 0) || (s->A[k] == 0 && s->B[k] > 0);
    }
    int32_t xmin = s->X[0], xmax = s->X[0], ymin = s->Y[0], ymax = s->Y[0];
    for (int k = 1; k < 3; ++k) {
        if (s->X[k] < xmin) xmin = s->X[k];
        if (s->X[k] > xmax) xmax = s->X[k];
        if (s->Y[k] < ymin) ymin = s->Y[k];
        if (s->Y[k] > ymax) ymax = s->Y[k];
    }
    int64_t x0 = floor_div256((int64_t)xmin + 127), x1 = floor_div256((int64_t)xmax - 128);
    int64_t y0 = floor_div256((int64_t)ymin + 127), y1 = floor_div256((int64_t)ymax - 128);
    if (x0 < 0) x0 = 0;
    if (y0 < 0) y0 = 0;
    if (x1 > (int64_t)R - 1) x1 = (int64_t)R - 1;
    if (y1 > (int64_t)R - 1) y1 = (int64_t)R - 1;
    s->x0 = (int32_t)x0; s->x1 = (int32_t)x1; s->y0 = (int32_t)y0; s->y1 = (int32_t)y1;
    return (x1 >= x0 && y1 >= y0) ? 1 : 0;
}

static inline int orc_inside(const orc_setup* s, int i, int j, int64_t E[3]) {
    for (int k = 0; k < 3; ++k) {
        E[k] = s->A[k] * i + s->B[k] * j + s->C[k];
        if (E[k] < 0 || (E[k] == 0 && !s->incl[k])) return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* textures: mip chain + trilinear REPEAT sampler                                               */
/* ------------------------------------------------------------------------------------------ */
typedef struct orc_tex {
    uint8_t* level[M2S_MAX_MIP_LEVEL + 1];
    uint32_t w[M2S_MAX_MIP_LEVEL + 1], h[M2S_MAX_MIP_LEVEL + 1];
    uint32_t nlevels; /* q + 1 */
    int owns0;
} orc_tex;

/* number of levels = min(4, floor(log2(max(w,h)))) + 1 */
ORC_API uint32_t orc_mip_count(uint32_t w, uint32_t h) {
    uint32_t m = w > h ? w : h, q = 0;
    while ((m >> (q + 1)) != 0) ++q;
    if (q > M2S_MAX_MIP_LEVEL) q = M2S_MAX_MIP_LEVEL;
    return q + 1;
}

/* one 2x2 box step, round half up; odd sizes clamp the second tap (driver-defined in GL) */
ORC_API void orc_mip_down(const uint8_t* src, uint32_t sw, uint32_t sh, uint8_t* dst, uint32_t* dw_, uint32_t* dh_) {
    uint32_t dw = sw / 2 ? sw / 2 : 1, dh = sh / 2 ? sh / 2 : 1;
    for (uint32_t y = 0; y < dh; ++y) {
        uint32_t y0 = 2 * y < sh ? 2 * y : sh - 1, y1 = 2 * y + 1 < sh ? 2 * y + 1 : sh - 1;
        for (uint32_t x = 0; x < dw; ++x) {
            uint32_t x0 = 2 * x < sw ? 2 * x : sw - 1, x1 = 2 * x + 1 < sw ? 2 * x + 1 : sw - 1;
            for (int c = 0; c < 4; ++c) {
                uint32_t sum = src[(y0 * (size_t)sw + x0) * 4 + c] + src[(y0 * (size_t)sw + x1) * 4 + c] +
                               src[(y1 * (size_t)sw + x0) * 4 + c] + src[(y1 * (size_t)sw + x1) * 4 + c];
                dst[(y * (size_t)dw + x) * 4 + c] = (uint8_t)((sum + 2) >> 2);
            }
        }
    }
    *dw_ = dw; *dh_ = dh;
}

static void orc_tex_build(orc_tex* t, const m2s_texture* src) {
    memset(t, 0, sizeof(*t));
    t->level[0] = (uint8_t*)src->rgba; t->w[0] = src->width; t->h[0] = src->height;
    t->nlevels = orc_mip_count(src->width, src->height);
    for (uint32_t l = 1; l < t->nlevels; ++l) {
        uint32_t dw = t->w[l - 1] / 2 ? t->w[l - 1] / 2 : 1, dh = t->h[l - 1] / 2 ? t->h[l - 1] / 2 : 1;
        t->level[l] = (uint8_t*)malloc((size_t)dw * dh * 4);
        orc_mip_down(t->level[l - 1], t->w[l - 1], t->h[l - 1], t->level[l], &t->w[l], &t->h[l]);
    }
}
static void orc_tex_free(orc_tex* t) { for (uint32_t l = 1; l < t->nlevels; ++l) free(t->level[l]); }

/* level `l` of the chain of one image; dst must hold w*h*4 of that level. For parity tests. */
ORC_API int orc_mip_level(const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t level, uint8_t* dst, uint32_t* ow, uint32_t* oh) {
    m2s_texture src = {rgba, w, h};
    orc_tex t; orc_tex_build(&t, &src);
    if (level >= t.nlevels) { orc_tex_free(&t); return -1; }
    memcpy(dst, t.level[level], (size_t)t.w[level] * t.h[level] * 4);
    *ow = t.w[level]; *oh = t.h[level];
    orc_tex_free(&t);
    return 0;
}

static inline uint32_t wrap_repeat(int64_t i, uint32_t n) { int64_t m = i % (int64_t)n; if (m < 0) m += n; return (uint32_t)m; }

static void bilinear(const orc_tex* t, uint32_t l, float u, float v, float out[4]) {
    const uint32_t W = t->w[l], H = t->h[l];
    float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f;
    float fx0 = floorf(x), fy0 = floorf(y);
    float ax = x - fx0, ay = y - fy0;
    int64_t ix = (int64_t)fx0, iy = (int64_t)fy0;
    uint32_t x0 = wrap_repeat(ix, W), x1 = wrap_repeat(ix + 1, W), y0 = wrap_repeat(iy, H), y1 = wrap_repeat(iy + 1, H);
    const uint8_t* p00 = t->level[l] + ((size_t)y0 * W + x0) * 4;
    const uint8_t* p10 = t->level[l] + ((size_t)y0 * W + x1) * 4;
    const uint8_t* p01 = t->level[l] + ((size_t)y1 * W + x0) * 4;
    const uint8_t* p11 = t->level[l] + ((size_t)y1 * W + x1) * 4;
    const float k = 1.0f / 255.0f;
    for (int c = 0; c < 4; ++c) {
        float t0 = (float)p00[c] * k, t1 = (float)p10[c] * k, t2 = (float)p01[c] * k, t3 = (float)p11[c] * k;
        float top = t0 + ax * (t1 - t0), bot = t2 + ax * (t3 - t2);
        out[c] = top + ay * (bot - top);
    }
}

/* GL 4.6 8.14: lambda <= 0 -> magnification (LINEAR, level 0); else LINEAR_MIPMAP_LINEAR between
 * floor(lambda) and floor(lambda)+1, clamped to level q = nlevels-1. */
static void sample_trilinear(const orc_tex* t, float u, float v, float lambda, float out[4]) {
    const float q = (float)(t->nlevels - 1);
    if (!(lambda > 0.0f)) { bilinear(t, 0, u, v, out); return; }
    if (lambda >= q) { bilinear(t, t->nlevels - 1, u, v, out); return; }
    float d = floorf(lambda), f = lambda - d;
    float a[4], b[4];
    bilinear(t, (uint32_t)d, u, v, a);
    if (f == 0.0f) { memcpy(out, a, sizeof(a)); return; }
    bilinear(t, (uint32_t)d + 1, u, v, b);
    for (int c = 0; c < 4; ++c) out[c] = a[c] + f * (b[c] - a[c]);
}

/* rho from the screen-space derivatives of (u*W, v*H); UV is affine per triangle so lambda is a
 * per-(triangle,map) constant.  dudx etc are per-pixel steps of the mesh UV. */
static float lod_lambda(float dudx, float dvdx, float dudy, float dvdy, uint32_t W, uint32_t H) {
    float ax = dudx * (float)W, bx = dvdx * (float)H, ay = dudy * (float)W, by = dvdy * (float)H;
    float rx = sqrtf(ax * ax + bx * bx), ry = sqrtf(ay * ay + by * by);
    float rho = rx > ry ? rx : ry;
    return log2f(rho);
}

/* ------------------------------------------------------------------------------------------ */
/* per-fragment stage (converterFS.glsl:44-104) on already-fetched texels                       */
/* ------------------------------------------------------------------------------------------ */
/* in: P3 N3 T4 (interpolated), texels (albedo rgba, normal rgba, mr rgba), flags bit0 albedo,
 * bit1 normal, bit2 mr; factor; scale, quat from the per-triangle stage. out: 24 floats REF96. */
ORC_API void orc_fragment(const float P[3], const float N[3], const float T[4], const float scale[3],
                          const float quat[4], const float albedo[4], const float nrm[4], const float mr[4],
                          uint32_t flags, const float factor[4], float rec[24]) {
    float col[4];
    if (flags & 1u) { for (int c = 0; c < 4; ++c) col[c] = albedo[c]; }
    else { col[0] = col[1] = col[2] = col[3] = 1.0f; }
    float on[3];
    if (flags & 2u) {
        v3 nm = {nrm[0] * 2.0f - 1.0f, nrm[1] * 2.0f - 1.0f, nrm[2] * 2.0f - 1.0f};
        v3 rn = v3norm(nm);                                  /* :71 */
        v3 Nv = {N[0], N[1], N[2]}, Tv = {T[0], T[1], T[2]};
        v3 bit = v3scale(v3norm(v3cross(Nv, Tv)), T[3]);     /* :73 */
        v3 Nn = v3norm(Nv);                                  /* :74 */
        v3 o = {Tv.x * rn.x + bit.x * rn.y + Nn.x * rn.z,
                Tv.y * rn.x + bit.y * rn.y + Nn.y * rn.z,
                Tv.z * rn.x + bit.z * rn.y + Nn.z * rn.z};
        o = v3norm(o);                                       /* :76 */
        on[0] = o.x; on[1] = o.y; on[2] = o.z;
    } else { on[0] = N[0]; on[1] = N[1]; on[2] = N[2]; }       /* :79-81, un-normalised */
    float metal, rough;
    if (flags & 4u) { metal = mr[2]; rough = mr[1]; }          /* .bg, :88-92 */
    else { metal = 0.1f; rough = 0.5f; }                       /* :93-95 */
    rec[0] = P[0]; rec[1] = P[1]; rec[2] = P[2]; rec[3] = 1.0f;
    for (int c = 0; c < 4; ++c) rec[4 + c] = col[c] * factor[c];
    rec[8] = scale[0]; rec[9] = scale[1]; rec[10] = scale[2]; rec[11] = 0.0f;
    rec[12] = on[0]; rec[13] = on[1]; rec[14] = on[2]; rec[15] = 0.0f;
    rec[16] = quat[0]; rec[17] = quat[1]; rec[18] = quat[2]; rec[19] = quat[3];
    rec[20] = metal; rec[21] = rough; rec[22] = 0.0f; rec[23] = 1.0f;
}

/* ------------------------------------------------------------------------------------------ */
/* export arithmetic (parsers.cpp) — REF96 record -> other layouts                              */
/* ------------------------------------------------------------------------------------------ */
static inline float inv_sigmoid(float a) { /* utils.hpp:270 */
    if (a < 0.0f) a = 0.0f;
    if (a > 1.0f) a = 1.0f;
    return -logf((1.0f / (a + 1e-8f)) - 1.0f);
}
static inline uint8_t to_byte(float v) { /* parsers.cpp:370-375 */
    if (v < 0.0f) v = 0.0f;
    if (v > 1.0f) v = 1.0f;
    return (uint8_t)roundf(v * 255.0f);
}
static void encode_octa(const float n[3], float out[2]) { /* parsers.cpp:318-337 */
    float s = fabsf(n[0]) + fabsf(n[1]) + fabsf(n[2]) + 1e-8f;
    float x = n[0] / s, y = n[1] / s, z = n[2] / s;
    float rx, ry;
    if (z >= 0.0f) { rx = x; ry = y; }
    else {
        float m = (x >= 0.0f && y >= 0.0f) ? 1.0f : -1.0f;
        rx = (1.0f - fabsf(y)) * m; ry = (1.0f - fabsf(x)) * m;
    }
    out[0] = rx * 0.5f + 0.5f; out[1] = ry * 0.5f + 0.5f;
}
static inline uint8_t clamp_round_u8(float v) { /* parsers.cpp:411-412 */
    float r = roundf(v); if (r < 0.0f) r = 0.0f; if (r > 255.0f) r = 255.0f; return (uint8_t)r;
}

ORC_API uint32_t orc_record_stride(uint32_t layout) {
    switch (layout) {
        case M2S_LAYOUT_REF96: return 96; case M2S_LAYOUT_PACKED56: return 56;
        case M2S_LAYOUT_PLY_STANDARD: return 248; case M2S_LAYOUT_PLY_PBR: return 76;
        case M2S_LAYOUT_PLY_COMPRESSED: return 48; default: return 0;
    }
}

/* rec: REF96 (24 floats); mult = gaussianStd / R (SceneManager.cpp:668) */
ORC_API void orc_encode(uint32_t layout, const float* rec, float mult, uint8_t* dst) {
    float f[62];
    const float sh0[3] = {(rec[4] - 0.5f) / SH_C0, (rec[5] - 0.5f) / SH_C0, (rec[6] - 0.5f) / SH_C0};
    const float op = inv_sigmoid(rec[7]);
    const float ls[3] = {logf(rec[8] * mult), logf(rec[9] * mult), logf(rec[10] * mult)};
    switch (layout) {
    case M2S_LAYOUT_REF96: memcpy(dst, rec, 96); break;
    case M2S_LAYOUT_PACKED56:
        f[0] = rec[0]; f[1] = rec[1]; f[2] = rec[2];
        f[3] = rec[16]; f[4] = rec[17]; f[5] = rec[18]; f[6] = rec[19];
        f[7] = ls[0]; f[8] = ls[1]; f[9] = ls[2];
        f[10] = sh0[0]; f[11] = sh0[1]; f[12] = sh0[2]; f[13] = op;
        memcpy(dst, f, 56); break;
    case M2S_LAYOUT_PLY_STANDARD: /* parsers.cpp:469-511 */
        memset(f, 0, sizeof(f));
        f[0] = rec[0]; f[1] = rec[1]; f[2] = rec[2];
        f[3] = rec[12]; f[4] = rec[13]; f[5] = rec[14];
        f[6] = sh0[0]; f[7] = sh0[1]; f[8] = sh0[2];
        f[54] = op; f[55] = ls[0]; f[56] = ls[1]; f[57] = ls[2];
        f[58] = rec[16]; f[59] = rec[17]; f[60] = rec[18]; f[61] = rec[19];
        memcpy(dst, f, 248); break;
    case M2S_LAYOUT_PLY_PBR: /* parsers.cpp:268-313 */
        f[0] = rec[0]; f[1] = rec[1]; f[2] = rec[2];
        f[3] = rec[12]; f[4] = rec[13]; f[5] = rec[14];
        f[6] = sh0[0]; f[7] = sh0[1]; f[8] = sh0[2];
        f[9] = rec[20]; f[10] = rec[21]; f[11] = op;
        f[12] = ls[0]; f[13] = ls[1]; f[14] = ls[2];
        f[15] = rec[16]; f[16] = rec[17]; f[17] = rec[18]; f[18] = rec[19];
        memcpy(dst, f, 76); break;
    case M2S_LAYOUT_PLY_COMPRESSED: { /* parsers.cpp:378-424 */
        uint8_t* p = dst;
        memcpy(p, rec, 12); p += 12;
        *p++ = to_byte(rec[4]); *p++ = to_byte(rec[5]); *p++ = to_byte(rec[6]); *p++ = to_byte(rec[7]);
        memcpy(p, rec + 16, 16); p += 16;
        float mn = rec[8] < rec[9] ? rec[8] : rec[9]; /* std::min(x,y) = (y<x)?y:x */
        if (rec[9] < rec[8]) mn = rec[9]; else mn = rec[8];
        float cs[3] = {ls[0], ls[1], logf(mn * mult)};
        memcpy(p, cs, 12); p += 12;
        float oc[2]; encode_octa(rec + 12, oc);
        *p++ = clamp_round_u8(oc[0] * 255.0f); *p++ = clamp_round_u8(oc[1] * 255.0f);
        *p++ = to_byte(rec[21]); *p++ = to_byte(rec[20]);
        break; }
    default: break;
    }
}

ORC_API size_t orc_ply_header(uint32_t format, uint64_t count, char* dst, size_t cap) {
    char buf[4096]; size_t n = 0;
#define APP(...) n += (size_t)snprintf(buf + n, sizeof(buf) - n, __VA_ARGS__)
    APP("ply\nformat binary_little_endian 1.0\nelement vertex %llu\n", (unsigned long long)count);
    if (format == 1) {
        const char* p[] = {"x","y","z","nx","ny","nz","f_dc_0","f_dc_1","f_dc_2","metallicFactor","roughnessFactor",
                           "opacity","scale_0","scale_1","scale_2","rot_0","rot_1","rot_2","rot_3"};
        for (int i = 0; i < 19; ++i) APP("property float %s\n", p[i]);
    } else if (format == 2) {
        APP("property float x\nproperty float y\nproperty float z\n");
        APP("property uint8 red\nproperty uint8 green\nproperty uint8 blue\nproperty uint8 opacity\n");
        APP("property float rot_0\nproperty float rot_1\nproperty float rot_2\nproperty float rot_3\n");
        APP("property float scale_0\nproperty float scale_1\nproperty float scale_2\n");
        APP("property uint8 octa_nx\nproperty uint8 octa_ny\nproperty uint8 roughness\nproperty uint8 metallic\n");
    } else {
        const char* p[] = {"x","y","z","nx","ny","nz","f_dc_0","f_dc_1","f_dc_2"};
        for (int i = 0; i < 9; ++i) APP("property float %s\n", p[i]);
        for (int i = 0; i <= 44; ++i) APP("property float f_rest_%d\n", i);
        const char* t[] = {"opacity","scale_0","scale_1","scale_2","rot_0","rot_1","rot_2","rot_3"};
        for (int i = 0; i < 8; ++i) APP("property float %s\n", t[i]);
    }
    APP("end_header\n");
#undef APP
    if (dst && cap > n) memcpy(dst, buf, n + 1);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* bbox rule                                                                                    */
/* ------------------------------------------------------------------------------------------ */
ORC_API void orc_compute_bboxes(const float* tris, m2s_primitive* prims, uint32_t nprim, int cumulative) {
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    for (uint32_t p = 0; p < nprim; ++p) {
        if (!cumulative) for (int c = 0; c < 3; ++c) { mn[c] = 3.402823466e+38f; mx[c] = -3.402823466e+38f; }
        for (uint64_t t = prims[p].first_triangle; t < prims[p].first_triangle + prims[p].triangle_count; ++t)
            for (int k = 0; k < 3; ++k)
                for (int c = 0; c < 3; ++c) {
                    float v = tris[t * 36 + 12 * k + c];
                    if (v < mn[c]) mn[c] = v; /* std::min(a,b) = (b<a)?b:a */
                    if (mx[c] < v) mx[c] = v;
                }
        for (int c = 0; c < 3; ++c) { prims[p].bbox_min[c] = mn[c]; prims[p].bbox_max[c] = mx[c]; }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* the whole pass                                                                               */
/* ------------------------------------------------------------------------------------------ */
typedef struct tri_ctx {
    orc_setup s;
    float lam[3]; /* lod per map */
    uint32_t prim;
} tri_ctx;

static uint64_t count_fragments(const orc_setup* s) {
    if (s->area2 == 0 || s->x1 < s->x0 || s->y1 < s->y0) return 0;
    uint64_t n = 0; int64_t E[3];
    for (int j = s->y0; j <= s->y1; ++j)
        for (int i = s->x0; i <= s->x1; ++i) n += (uint64_t)orc_inside(s, i, j, E);
    return n;
}

/* Converts params->{first_triangle, triangle_count} of the scene. Output order is deterministic:
 * triangle-major, then row-major (y, x).  Returns written records; *total = all fragments. */
/* mip chains built once per scene (the reference builds them at load time, glUtils.cpp:305) */
typedef struct orc_prepared { orc_tex* tex; uint32_t ntex; } orc_prepared;

ORC_API orc_prepared* orc_prepare(const m2s_scene* sc) {
    orc_prepared* p = (orc_prepared*)calloc(1, sizeof(orc_prepared));
    p->ntex = sc->texture_count;
    p->tex = (orc_tex*)calloc(sc->texture_count ? sc->texture_count : 1, sizeof(orc_tex));
    for (uint32_t i = 0; i < sc->texture_count; ++i) orc_tex_build(&p->tex[i], &sc->textures[i]);
    return p;
}
ORC_API void orc_release(orc_prepared* p) {
    if (!p) return;
    for (uint32_t i = 0; i < p->ntex; ++i) orc_tex_free(&p->tex[i]);
    free(p->tex); free(p);
}

ORC_API uint64_t orc_convert_prepared(const m2s_scene* sc, const orc_prepared* prep, const m2s_params* pr, void* out,
                                      uint64_t out_capacity, uint64_t* keys, uint64_t* total_out, int threads) {
    const uint32_t R = pr->resolution;
    const uint32_t stride = orc_record_stride(pr->layout);
    uint64_t first = pr->first_triangle, cnt = pr->triangle_count;
    if (first > sc->triangle_count) first = sc->triangle_count;
    if (cnt == 0 || first + cnt > sc->triangle_count) cnt = sc->triangle_count - first;
    uint64_t cap = pr->max_gaussians;
    if (cap == 0) {
        if (pr->flags & M2S_FLAG_UNCAPPED) cap = out_capacity;
        else { /* ConversionPass.cpp:21-24 */
            uint64_t mc = sc->primitive_count ? sc->primitive_count : 1;
            cap = (uint64_t)R * R * 6ull * mc;
            if (cap > M2S_REFERENCE_MAX_GAUSSIANS) cap = M2S_REFERENCE_MAX_GAUSSIANS;
        }
    }
    if (cap > out_capacity) cap = out_capacity;
    const float mult = pr->gaussian_std / (float)R;
    const uint32_t row_begin = pr->row_begin < R ? pr->row_begin : R;
    const uint32_t row_end = (pr->row_end == 0 || pr->row_end > R) ? R : pr->row_end;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    const orc_tex* tex = prep->tex;

    /* triangle -> primitive */
    uint32_t* prim_of = (uint32_t*)malloc(sizeof(uint32_t) * (cnt ? cnt : 1));
    for (uint64_t t = 0; t < cnt; ++t) prim_of[t] = 0xffffffffu;
    for (uint32_t p = 0; p < sc->primitive_count; ++p) {
        uint64_t a = sc->primitives[p].first_triangle, b = a + sc->primitives[p].triangle_count;
        for (uint64_t t = (a > first ? a : first); t < b && t < first + cnt; ++t) prim_of[t - first] = p;
    }
    uint64_t* offs = (uint64_t*)malloc(sizeof(uint64_t) * (cnt + 1));
    tri_ctx* tc = (tri_ctx*)malloc(sizeof(tri_ctx) * (cnt ? cnt : 1));

#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t t = 0; t < (int64_t)cnt; ++t) {
        tri_ctx* c = &tc[t];
        c->prim = prim_of[t];
        offs[t + 1] = 0;
        if (c->prim == 0xffffffffu) { memset(&c->s, 0, sizeof(c->s)); c->s.x1 = -1; c->s.y1 = -1; continue; }
        const m2s_primitive* P = &sc->primitives[c->prim];
        const float* tri = sc->triangles + (first + (uint64_t)t) * 36;
        if (!orc_triangle_setup(tri, P->bbox_min, P->bbox_max, R, &c->s)) continue;
        /* pixel-row band of this call (m2s_params.row_begin/row_end; the whole grid by default) */
        if (c->s.y0 < (int32_t)row_begin) c->s.y0 = (int32_t)row_begin;
        if (c->s.y1 > (int32_t)row_end - 1) c->s.y1 = (int32_t)row_end - 1;
        if (c->s.y1 < c->s.y0) continue;
        offs[t + 1] = count_fragments(&c->s);
        /* per-pixel steps of the mesh uv: d(uv)/dx = sum_k uv_k * A_k / area2, d/dy with B_k */
        float ia = 1.0f / (float)(c->s.area2 < 0 ? -c->s.area2 : c->s.area2);
        float dudx = 0, dvdx = 0, dudy = 0, dvdy = 0;
        for (int k = 0; k < 3; ++k) {
            float a = (float)c->s.A[k] * ia, b = (float)c->s.B[k] * ia;
            dudx += tri[12 * k + 10] * a; dvdx += tri[12 * k + 11] * a;
            dudy += tri[12 * k + 10] * b; dvdy += tri[12 * k + 11] * b;
        }
        const int32_t ti[3] = {P->albedo_texture, P->normal_texture, P->metallic_roughness_texture};
        for (int m = 0; m < 3; ++m)
            c->lam[m] = (ti[m] >= 0 && (uint32_t)ti[m] < sc->texture_count)
                            ? lod_lambda(dudx, dvdx, dudy, dvdy, tex[ti[m]].w[0], tex[ti[m]].h[0]) : 0.0f;
    }
    offs[0] = 0;
    for (uint64_t t = 0; t < cnt; ++t) offs[t + 1] += offs[t];
    const uint64_t total = offs[cnt];

#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t t = 0; t < (int64_t)cnt; ++t) {
        if (offs[t + 1] == offs[t]) continue;
        const tri_ctx* c = &tc[t];
        const orc_setup* s = &c->s;
        const m2s_primitive* P = &sc->primitives[c->prim];
        const float* tri = sc->triangles + (first + (uint64_t)t) * 36;
        const int32_t ti[3] = {P->albedo_texture, P->normal_texture, P->metallic_roughness_texture};
        uint32_t flags = 0;
        for (int m = 0; m < 3; ++m) if (ti[m] >= 0 && (uint32_t)ti[m] < sc->texture_count) flags |= 1u << m;
        /* only the maps the layout's record carries are sampled (the product uploads and samples the same set):
         * PACKED56 has neither normal nor metallic/roughness, the standard .ply row no metallic/roughness; the
         * values that ARE stored do not depend on the skipped maps */
        if (pr->layout == M2S_LAYOUT_PACKED56) flags &= 1u;
        else if (pr->layout == M2S_LAYOUT_PLY_STANDARD) flags &= 3u;
        const float ia = 1.0f / (float)(s->area2 < 0 ? -s->area2 : s->area2);
        uint64_t idx = offs[t];
        int64_t E[3];
        for (int j = s->y0; j <= s->y1; ++j)
            for (int i = s->x0; i <= s->x1; ++i) {
                if (!orc_inside(s, i, j, E)) continue;
                uint64_t my = idx++;
                if (my >= cap) continue; /* converterFS.glsl:48-51 */
                float l[3] = {(float)E[0] * ia, (float)E[1] * ia, (float)E[2] * ia};
                float a[12];
                for (int q = 0; q < 12; ++q) a[q] = l[0] * tri[q] + l[1] * tri[12 + q] + l[2] * tri[24 + q];
                float tx[3][4] = {{0}};
                for (int m = 0; m < 3; ++m)
                    if (flags & (1u << m)) sample_trilinear(&tex[ti[m]], a[10], a[11], c->lam[m], tx[m]);
                float rec[24];
                orc_fragment(a, a + 3, a + 6, s->scale, s->quat, tx[0], tx[1], tx[2], flags, P->base_color_factor, rec);
                orc_encode(pr->layout, rec, mult, (uint8_t*)out + my * stride);
                if (keys) keys[my] = ((first + (uint64_t)t) << 24) | ((uint64_t)j << 12) | (uint64_t)i;
            }
    }
    free(prim_of); free(offs); free(tc);
    if (total_out) *total_out = total;
    return total < cap ? total : cap;
}

ORC_API uint64_t orc_convert(const m2s_scene* sc, const m2s_params* pr, void* out, uint64_t out_capacity,
                             uint64_t* keys, uint64_t* total_out, int threads) {
    orc_prepared* p = orc_prepare(sc);
    const uint64_t n = orc_convert_prepared(sc, p, pr, out, out_capacity, keys, total_out, threads);
    orc_release(p);
    return n;
}

/* single texture fetch, for sampler unit tests */
ORC_API void orc_sample(const uint8_t* rgba, uint32_t w, uint32_t h, float u, float v, float lambda, float out[4]) {
    m2s_texture src = {rgba, w, h};
    orc_tex t; orc_tex_build(&t, &src);
    sample_trilinear(&t, u, v, lambda, out);
    orc_tex_free(&t);
}

/* ---- viewer prepass: gaussianSplattingPrepassCS.glsl:58-204 + common.glsl ---------------------------
 * Plain-C restatement, one invocation after the other in gid order (survivors come out in input order).  Pinned to the
 * reference's own shader compiled here (oracle/_ref/libm2s_refprepass.so, tests/golden/ref_prepass_vectors.npz): same
 * survivors, values within 1e-5 relative (GLM's operation order inside inverse()/matrix products is not replicated bit
 * for bit).  Matrices are column-major: m[c * 4 + r].  format 0: GaussianVertex input as the converter writes it;
 * format 1 (ply_has_pbr 0): a loaded standard .ply.  render_mode 3 and the mesh depth test are not restated. */
static void m3mul(const float a[9], const float b[9], float r[9]) {   /* column-major 3x3: r = a * b */
    for (int c = 0; c < 3; ++c)
        for (int row = 0; row < 3; ++row)
            r[c * 3 + row] = a[0 * 3 + row] * b[c * 3 + 0] + a[1 * 3 + row] * b[c * 3 + 1] + a[2 * 3 + row] * b[c * 3 + 2];
}
static void m3transpose(const float a[9], float r[9]) { for (int c = 0; c < 3; ++c) for (int k = 0; k < 3; ++k) r[c * 3 + k] = a[k * 3 + c]; }
static void m3inverse(const float m[9], float r[9]) {
    const float a = m[0], b = m[3], c = m[6], d = m[1], e = m[4], f = m[7], g = m[2], h = m[5], i = m[8];   /* rows */
    const float det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g), inv = 1.0f / det;
    r[0] = (e * i - f * h) * inv; r[3] = (c * h - b * i) * inv; r[6] = (b * f - c * e) * inv;
    r[1] = (f * g - d * i) * inv; r[4] = (a * i - c * g) * inv; r[7] = (c * d - a * f) * inv;
    r[2] = (d * h - e * g) * inv; r[5] = (b * g - a * h) * inv; r[8] = (a * e - b * d) * inv;
}
static int m4inverse(const float m[16], float inv[16]) {   /* general 4x4 (cofactors), column-major */
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    if (det == 0.0f) return 0;
    const float id = 1.0f / det;
    for (int k = 0; k < 16; ++k) inv[k] *= id;
    return 1;
}
static void m4mulv(const float m[16], const float v[4], float r[4]) {
    for (int row = 0; row < 4; ++row) r[row] = m[0 + row] * v[0] + m[4 + row] * v[1] + m[8 + row] * v[2] + m[12 + row] * v[3];
}
static inline float clamp01(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }

ORC_API uint32_t orc_prepass(const float* gaussians, uint32_t n, const float* world_to_view, const float* view_to_clip, const float* model_to_world,
                             const float resolution[2], const float near_far[2], float std_dev, int render_mode, uint32_t format, uint32_t ply_has_pbr,
                             float* quads, float* depths) {
    const float* M = model_to_world;
    float Minv[16], nmat[16];   /* normal matrix: transpose(inverse(M)) (:119) */
    const int has_inv = m4inverse(M, Minv);
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) nmat[c * 4 + r] = has_inv ? Minv[r * 4 + c] : 0.0f;
    const float mrot[9] = {M[0], M[1], M[2], M[4], M[5], M[6], M[8], M[9], M[10]};   /* :103-107 */
    float mrot_inv[9];
    m3inverse(mrot, mrot_inv);
    const float len0 = sqrtf(M[0] * M[0] + M[1] * M[1] + M[2] * M[2] + M[3] * M[3]), len1 = sqrtf(M[4] * M[4] + M[5] * M[5] + M[6] * M[6] + M[7] * M[7]);
    const float mscale[3] = {len0, len0, len1};   /* :96 (sic: column 0 twice, then column 1) */
    const float W[9] = {world_to_view[0], world_to_view[1], world_to_view[2], world_to_view[4], world_to_view[5], world_to_view[6],
                        world_to_view[8], world_to_view[9], world_to_view[10]};   /* mat3(u_worldToView) :167 */
    uint32_t valid = 0;
    for (uint32_t gid = 0; gid < n; ++gid) {
        const float* g = gaussians + (size_t)gid * 24;   /* position color scale normal rotation pbr */
        const float p4[4] = {g[0], g[1], g[2], 1.0f};
        float ws[4], vs[4], clipv[4];
        m4mulv(M, p4, ws);                                   /* :66 */
        const float ws1[4] = {ws[0], ws[1], ws[2], 1.0f};
        m4mulv(world_to_view, ws1, vs);                      /* :68 */
        m4mulv(view_to_clip, vs, clipv);                     /* :70 */
        const float clip = 1.05f * clipv[3];                 /* :72-76 */
        if (clipv[2] < -clip || clipv[0] < -clip || clipv[0] > clip || clipv[1] < -clip || clipv[1] > clip) continue;
        const float multiplier = (format == 0 || format == 3) ? std_dev : 1.0f;   /* :95-97 */
        const float sc[3] = {g[8] * multiplier * (mscale[0] * mscale[0]), g[9] * multiplier * (mscale[1] * mscale[1]), g[10] * multiplier * (mscale[2] * mscale[2])};
        /* castQuatToMat3 (common.glsl:22-48): the three "rows" are the COLUMNS of the GLM matrix; quat = (w, x, y, z) in .x .y .z .w */
        const float qx = g[16], qy = g[17], qz = g[18], qw = g[19];
        const float rot0[9] = {1.f - 2.f * (qz * qz + qw * qw), 2.f * (qy * qz - qx * qw), 2.f * (qy * qw + qx * qz),
                               2.f * (qy * qz + qx * qw), 1.f - 2.f * (qy * qy + qw * qw), 2.f * (qz * qw - qx * qy),
                               2.f * (qy * qw - qx * qz), 2.f * (qz * qw + qx * qy), 1.f - 2.f * (qy * qy + qz * qz)};
        float rot[9];
        m3mul(rot0, mrot_inv, rot);                          /* :109 */
        /* computeCov3D (common.glsl:50-61): mMatrix = diag(scales) * rotMat; sigma = transpose(mMatrix) * mMatrix */
        float mm[9], mmT[9], cov3d[9];
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) mm[c * 3 + r] = sc[r] * rot[c * 3 + r];
        m3transpose(mm, mmT);
        m3mul(mmT, mm, cov3d);
        float out_color[4] = {0, 0, 0, 0}, nrm[4] = {1, 0, 0, 0};
        const float depth_exp = clamp01(expf(-20.0f * clamp01((-vs[2] - near_far[0]) / (near_far[1] - near_far[0]))));   /* :114, common.glsl:80-84 */
        if (format == 0 || (format == 1 && ply_has_pbr != 0) || format == 3) {   /* :117-121 */
            const float n4[4] = {g[12], g[13], g[14], 1.0f};
            float nw[4];
            m4mulv(nmat, n4, nw);
            nrm[0] = nw[0] * 0.5f + 0.5f; nrm[1] = nw[1] * 0.5f + 0.5f; nrm[2] = nw[2] * 0.5f + 0.5f; nrm[3] = g[7];
        } else if (format == 1) {                                                  /* :123-130 shortest axis */
            const unsigned idx = (unsigned)((g[9] < g[10]) && (g[9] < g[8])) + (unsigned)((g[10] < g[9]) && (g[10] < g[8])) * 2u;
            nrm[0] = rot[idx * 3 + 0] * 0.5f + 0.5f; nrm[1] = rot[idx * 3 + 1] * 0.5f + 0.5f; nrm[2] = rot[idx * 3 + 2] * 0.5f + 0.5f; nrm[3] = g[7];
        }
        if (render_mode == 0 || render_mode == 6) { out_color[0] = g[4]; out_color[1] = g[5]; out_color[2] = g[6]; out_color[3] = g[7]; }   /* :132-148 */
        else if (render_mode == 1) { out_color[0] = out_color[1] = out_color[2] = depth_exp; out_color[3] = g[7]; }
        else if (render_mode == 2) { out_color[0] = nrm[0]; out_color[1] = nrm[1]; out_color[2] = nrm[2]; out_color[3] = nrm[3]; }
        const float ndc[4] = {clipv[0] / clipv[3], clipv[1] / clipv[3], clipv[2] / clipv[3], clipv[3]};   /* :150 */
        /* :153-166 Jacobian of the projection; J = mat3(col0 = (jsx,0,0), col1 = (0,jsy,0), col2 = (jtx,jty,jtz)) */
        const float tzSq = vs[2] * vs[2];
        const float jsx = -(view_to_clip[0] * resolution[0]) / (2 * vs[2]), jsy = -(view_to_clip[5] * resolution[1]) / (2 * vs[2]);
        const float jtx = (view_to_clip[0] * vs[0] * resolution[0]) / (2 * tzSq), jty = (view_to_clip[5] * vs[1] * resolution[1]) / (2 * tzSq);
        const float jtz = ((near_far[1] - near_far[0]) * view_to_clip[3 * 4 + 2]) / (2 * tzSq);
        const float J[9] = {jsx, 0.f, 0.f, 0.f, jsy, 0.f, jtx, jty, jtz};
        float JW[9], JWT[9], t9[9], Vp[9];
        m3mul(J, W, JW);
        m3transpose(JW, JWT);
        m3mul(JW, cov3d, t9);
        m3mul(t9, JWT, Vp);                                  /* :170 */
        float c00 = Vp[0] + 0.3f, c01 = Vp[1], c10 = Vp[3], c11 = Vp[4] + 0.3f;   /* mat2(V'), :172-176; cXY = cov2d[X][Y] (column X, row Y) */
        const float mid = c00 + c11;
        const float dx = c00 - c11, dy = 2 * c01;
        const float delta = sqrtf(dx * dx + dy * dy);
        const float lambda1 = 0.5f * (mid + delta), lambda2 = 0.5f * (mid - delta);
        if (lambda2 < 0.0f) continue;                        /* :185 */
        float dvx = 1.0f, dvy = (-c00 + c01 + lambda1) / (c01 - c11 + lambda1);
        const float dinv = 1.0f / sqrtf(dvx * dvx + dvy * dvy);   /* normalize (GLM: v * inversesqrt(dot)) */
        dvx *= dinv; dvy *= dinv;
        const float r1 = fminf(3 * sqrtf(lambda1), 1024.0f), r2 = fminf(3 * sqrtf(lambda2), 1024.0f);
        const float majx = r1 * dvx, majy = r1 * dvy, minx = r2 * dvy, miny = r2 * -dvx;
        const float hx = resolution[0] * 0.5f, hy = resolution[1] * 0.5f;
        float* q = quads + (size_t)valid * 24;
        q[0] = ndc[0]; q[1] = ndc[1]; q[2] = ndc[2]; q[3] = ndc[3];
        q[4] = majx / hx; q[5] = majy / hy; q[6] = minx / hx; q[7] = miny / hy;
        q[8] = out_color[0]; q[9] = out_color[1]; q[10] = out_color[2]; q[11] = out_color[3];
        const float det = c00 * c11 - c01 * c10;             /* inverseMat2 (common.glsl:63-78) */
        float i00 = 0, i01 = 0, i11 = 0;
        if (det != 0) { i00 = c11 / det; i01 = -c01 / det; i11 = c00 / det; }
        q[12] = i00; q[13] = i01; q[14] = i11; q[15] = -vs[2];
        q[16] = nrm[0]; q[17] = nrm[1]; q[18] = nrm[2]; q[19] = g[20];
        q[20] = ws[0]; q[21] = ws[1]; q[22] = ws[2]; q[23] = g[21];
        depths[valid] = vs[2];
        ++valid;
    }
    return valid;
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
