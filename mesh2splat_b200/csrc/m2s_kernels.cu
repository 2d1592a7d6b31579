// m2s_kernels.cu — the conversion pass as hand-written sm_100a CUDA.
//
// ONE persistent kernel replaces the reference's geometry shader, fixed-function rasteriser,
// fragment shader and SSBO atomic append (converter{GS,FS}.glsl, ConversionPass.cpp:114-116):
//
//   per CTA, until the batch counter runs dry
//     TMA (cp.async.bulk, mbarrier complete_tx) stages the next 128-triangle batch (18 KB) into
//       the idle half of a double buffer while the current one is processed
//     per-triangle stage (one thread per triangle; converterGS.glsl:326-443): longest edge,
//       face normal, dominant axis, orthographic uv, quaternion, UV->3D Jacobian scale, then
//       rasteriser set-up: 24.8 fixed-point window coords, int64 edge functions, top-left
//       ownership bits, candidate pixel box, per-map texture LOD
//     block scan of candidate counts -> flat candidate index space over the whole batch
//     rounds of <= 2048 candidates:
//       phase A  every lane tests one candidate pixel centre (exact integer edge functions) and
//                survivors are compacted warp-ballot-wise into a shared-memory fragment queue
//       one global atomicAdd per ROUND reserves the output range (not one per fragment, as the
//                reference's atomicCounterIncrement does)
//       phase B  (converterFS.glsl:44-104) full warps take 32 queued fragments: barycentric
//                interpolation from the staged vertex data, trilinear RGBA8 fetches, TBN normal,
//                record encode; records are transposed through shared memory so each warp
//                writes one contiguous 32*stride-byte span with vector stores
//   triangles whose pixel box exceeds 8192 candidates are pushed to a global chunk queue and
//   rasterised by ALL CTAs after the batches are done (keeps a 2-triangle quad at R=2048 from
//   serialising on one SM).
//
// Bit-exactness: every float operation that feeds a DECISION (edge ordering, dominant axis,
// fixed-point snapping => coverage) is written with __f*_rn intrinsics in the operation order of
// the oracle (and of GLM, which the reference's GLSL-as-C++ build uses), so coverage is bit-exact.
// Per-fragment values may use FMA contraction and are compared with a tolerance.
#include "m2s_device.cuh"

namespace m2s {

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA engine; SASS UBLKCP)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ------------------------------------------------------------------------------------------
// exact fp32 (one rounding per operation, GLM operation order)
// ------------------------------------------------------------------------------------------
struct f3 {
    float x, y, z;
};
__device__ __forceinline__ f3 sub3(f3 a, f3 b) { return {__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)}; }
__device__ __forceinline__ f3 scale3(f3 a, float s) { return {__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)}; }
__device__ __forceinline__ float dot3(f3 a, f3 b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ float len3(f3 a) { return __fsqrt_rn(dot3(a, a)); }
__device__ __forceinline__ f3 norm3(f3 a) { return scale3(a, __fdiv_rn(1.0f, __fsqrt_rn(dot3(a, a)))); }
__device__ __forceinline__ f3 cross3(f3 x, f3 y) {
    return {__fsub_rn(__fmul_rn(x.y, y.z), __fmul_rn(y.y, x.z)), __fsub_rn(__fmul_rn(x.z, y.x), __fmul_rn(y.z, x.x)),
            __fsub_rn(__fmul_rn(x.x, y.y), __fmul_rn(y.x, x.y))};
}

// ------------------------------------------------------------------------------------------
// shared-memory records
// ------------------------------------------------------------------------------------------
struct __align__(16) TriRaster {  // 64 B
    long long C[3];
    int A[3];
    int B[3];
    unsigned short x0, y0, w, h;
    float inv_area;
    unsigned incl;
};
struct __align__(16) TriFrag {  // 48 B
    float quat[4];   // (w,x,y,z)
    float scale[3];  // raw (REF96) or log(scale * sigma/R)
    float lod[3];
    int prim;
    unsigned tri;    // global triangle index
};

template <int LAYOUT>
struct LayoutTraits;
template <>
struct LayoutTraits<0> {  // REF96
    static constexpr int kStride = 96;
    static constexpr int kStagePitch = 112;  // padded: conflict-free float4 staging writes
    static constexpr bool kNeedNormal = true, kNeedMR = true, kLogScale = false;
};
template <>
struct LayoutTraits<1> {  // PACKED56
    static constexpr int kStride = 56;
    static constexpr int kStagePitch = 56;
    static constexpr bool kNeedNormal = false, kNeedMR = false, kLogScale = true;
};

struct SmemLayout {
    static constexpr int kTriIn = 0;                                     // 2 * kBatch * 144
    static constexpr int kRast = kTriIn + 2 * kBatch * kTriBytes;        // kBatch * 64
    static constexpr int kFrag = kRast + kBatch * 64;                    // kBatch * 48
    static constexpr int kPrefix = kFrag + kBatch * 48;                  // (kBatch + 4) * 4
    static constexpr int kQueueOff = kPrefix + (kBatch + 4) * 4;         // kQueue * 4
    static constexpr int kStage = kQueueOff + kQueue * 4;                // kWarps * 32 * pitch
    __host__ __device__ static constexpr int stage_bytes(int pitch) { return kWarps * 32 * pitch; }
    __host__ __device__ static constexpr int misc(int pitch) { return kStage + stage_bytes(pitch); }  // 64 B of scalars
    __host__ __device__ static constexpr int total(int pitch) { return misc(pitch) + 64; }
};

struct Misc {
    uint64_t bar[2];
    int batch[2];
    unsigned qcount;
    unsigned item;
    unsigned long long base;
    unsigned wsum[4];
};

// ------------------------------------------------------------------------------------------
// per-triangle stage + rasteriser set-up.  t: 36 floats in shared memory.
// Returns the number of candidate pixels (0 => nothing to rasterise).
// ------------------------------------------------------------------------------------------
template <int LAYOUT>
__device__ uint32_t setup_triangle(const float4* __restrict__ t4, uint32_t tri_global, const ConvertArgs& a,
                                   TriRaster& tr, TriFrag& tf) {
    using LT = LayoutTraits<LAYOUT>;
    tr.w = 0; tr.h = 0; tr.x0 = 0; tr.y0 = 0; tr.incl = 0; tr.inv_area = 0.f;
    tf.tri = tri_global;
    tf.prim = -1;
    // triangle -> primitive (sorted disjoint ranges)
    int lo = 0, hi = (int)a.nranges - 1, found = -1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        DRange r = a.ranges[mid];
        if (tri_global < r.first) hi = mid - 1;
        else if (tri_global >= r.end) lo = mid + 1;
        else { found = (int)r.prim; break; }
    }
    if (found < 0) return 0;
    tf.prim = found;
    const DPrim pr = a.prims[found];

    // vertex data: 3 x {pos3 nrm3 tan4 uv2} = 9 float4
    const float4 q0 = t4[0], q3 = t4[3], q6 = t4[6];
    const float4 q2 = t4[2], q5 = t4[5], q8 = t4[8];
    const f3 P0 = {q0.x, q0.y, q0.z}, P1 = {q3.x, q3.y, q3.z}, P2 = {q6.x, q6.y, q6.z};
    const float uvx[3] = {q2.z, q5.z, q8.z}, uvy[3] = {q2.w, q5.w, q8.w};

    // converterGS.glsl:327-347
    f3 e1 = sub3(P1, P0), e2 = sub3(P2, P0), e3 = sub3(P2, P1);
    const float l1 = len3(e1), l2 = len3(e2), l3 = len3(e3);
    if (l2 > l1 && l2 > l3) { f3 tmp = e1; e1 = e2; e2 = tmp; }
    else if (l3 > l1 && l3 > l2) { e1 = e3; }
    e1 = norm3(e1);
    const f3 n = norm3(cross3(e1, e2));
    const float ax = fabsf(n.x), ay = fabsf(n.y), az = fabsf(n.z);
    const int axis = (ax > ay && ax > az) ? 0 : ((ay > az) ? 1 : 2);

    // :354-397 orthogonal uv
    float ou[3], ov[3];
    {
        const int ia = axis == 0 ? 1 : 0, ib = axis == 2 ? 1 : 2;
        const float ra = __fsub_rn(pr.bmax[ia], pr.bmin[ia]), rb = __fsub_rn(pr.bmax[ib], pr.bmin[ib]);
        const float range = (ra < rb) ? rb : ra;
        const f3 Ps[3] = {P0, P1, P2};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float pa = ia == 0 ? Ps[k].x : Ps[k].y;
            const float pb = ib == 1 ? Ps[k].y : Ps[k].z;
            ou[k] = __fdiv_rn(__fsub_rn(pa, pr.bmin[ia]), range);
            ov[k] = __fdiv_rn(__fsub_rn(pb, pr.bmin[ib]), range);
        }
    }

    // :399-407 rotation -> quaternion (w,x,y,z), quat_cast :131-183
    {
        const f3 xA = e1, yA = norm3(cross3(n, xA)), zA = n;
        const float m00 = xA.x, m01 = xA.y, m02 = xA.z, m10 = yA.x, m11 = yA.y, m12 = yA.z, m20 = zA.x, m21 = zA.y,
                    m22 = zA.z;
        const float fX = __fsub_rn(__fsub_rn(m00, m11), m22), fY = __fsub_rn(__fsub_rn(m11, m00), m22),
                    fZ = __fsub_rn(__fsub_rn(m22, m00), m11), fW = __fadd_rn(__fadd_rn(m00, m11), m22);
        int big = 0;
        float fB = fW;
        if (fX > fB) { fB = fX; big = 1; }
        if (fY > fB) { fB = fY; big = 2; }
        if (fZ > fB) { fB = fZ; big = 3; }
        const float bv = __fmul_rn(__fsqrt_rn(__fadd_rn(fB, 1.0f)), 0.5f);
        const float mult = __fdiv_rn(0.25f, bv);
        float w, x, y, z;
        if (big == 0) { w = bv; x = __fmul_rn(__fsub_rn(m12, m21), mult); y = __fmul_rn(__fsub_rn(m20, m02), mult); z = __fmul_rn(__fsub_rn(m01, m10), mult); }
        else if (big == 1) { w = __fmul_rn(__fsub_rn(m12, m21), mult); x = bv; y = __fmul_rn(__fadd_rn(m01, m10), mult); z = __fmul_rn(__fadd_rn(m20, m02), mult); }
        else if (big == 2) { w = __fmul_rn(__fsub_rn(m20, m02), mult); x = __fmul_rn(__fadd_rn(m01, m10), mult); y = bv; z = __fmul_rn(__fadd_rn(m12, m21), mult); }
        else { w = __fmul_rn(__fsub_rn(m01, m10), mult); x = __fmul_rn(__fadd_rn(m20, m02), mult); y = __fmul_rn(__fadd_rn(m12, m21), mult); z = bv; }
        tf.quat[0] = w; tf.quat[1] = x; tf.quat[2] = y; tf.quat[3] = z;
    }

    // :269-300,206-235,414-430 Jacobian scale
    {
        const float ja = __fsub_rn(ou[1], ou[0]), jb = __fsub_rn(ou[2], ou[0]);
        const float jc = __fsub_rn(ov[1], ov[0]), jd = __fsub_rn(ov[2], ov[0]);
        const float det = __fsub_rn(__fmul_rn(ja, jd), __fmul_rn(jc, jb));
        float i00 = 0.f, i10 = 0.f, i01 = 0.f, i11 = 0.f;
        if (det != 0.0f) {
            const float invDet = __fdiv_rn(1.0f, det);
            i00 = __fmul_rn(jd, invDet); i10 = __fmul_rn(-jb, invDet); i01 = __fmul_rn(-jc, invDet); i11 = __fmul_rn(ja, invDet);
        }
        const f3 V0 = sub3(P1, P0), V1 = sub3(P2, P0);
        const f3 Ju = {__fadd_rn(__fmul_rn(V0.x, i00), __fmul_rn(V1.x, i01)), __fadd_rn(__fmul_rn(V0.y, i00), __fmul_rn(V1.y, i01)),
                       __fadd_rn(__fmul_rn(V0.z, i00), __fmul_rn(V1.z, i01))};
        const f3 Jv = {__fadd_rn(__fmul_rn(V0.x, i10), __fmul_rn(V1.x, i11)), __fadd_rn(__fmul_rn(V0.y, i10), __fmul_rn(V1.y, i11)),
                       __fadd_rn(__fmul_rn(V0.z, i10), __fmul_rn(V1.z, i11))};
        const float sx = len3(Ju), sy = len3(Jv), sz = 1e-7f;
        if (LT::kLogScale) {  // parsers.cpp:497-499 log(scale * sigma/R)
            tf.scale[0] = logf(__fmul_rn(sx, a.mult)); tf.scale[1] = logf(__fmul_rn(sy, a.mult)); tf.scale[2] = logf(__fmul_rn(sz, a.mult));
        } else { tf.scale[0] = sx; tf.scale[1] = sy; tf.scale[2] = sz; }
    }

    // rasteriser set-up: gl_Position = ouv*2-1 (:439), viewport R x R, 8 sub-pixel bits
    int X[3], Y[3];
    bool valid = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float ndx = __fsub_rn(__fmul_rn(ou[k], 2.0f), 1.0f), ndy = __fsub_rn(__fmul_rn(ov[k], 2.0f), 1.0f);
        const float xw = __fadd_rn(__fmul_rn(ndx, a.half_R), a.half_R), yw = __fadd_rn(__fmul_rn(ndy, a.half_R), a.half_R);
        if (!(fabsf(xw) <= kGuard) || !(fabsf(yw) <= kGuard)) valid = false;  // also rejects NaN
        X[k] = __float2int_rn(__fmul_rn(xw, 256.0f));
        Y[k] = __float2int_rn(__fmul_rn(yw, 256.0f));
    }
    if (!valid) return 0;
    long long area2 = (long long)(X[1] - X[0]) * (Y[2] - Y[0]) - (long long)(X[2] - X[0]) * (Y[1] - Y[0]);
    if (area2 == 0) return 0;
    const long long sg = area2 < 0 ? -1 : 1;
    unsigned incl = 0;
    int Ak[3], Bk[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int va = (k + 1) % 3, vb = (k + 2) % 3;
        const long long dx = X[vb] - X[va], dy = Y[vb] - Y[va];
        const long long A = sg * (-dy * 256), B = sg * (dx * 256);
        tr.A[k] = Ak[k] = (int)A;
        tr.B[k] = Bk[k] = (int)B;
        tr.C[k] = sg * (dx * (128 - (long long)Y[va]) - dy * (128 - (long long)X[va]));
        if (A > 0 || (A == 0 && B > 0)) incl |= 1u << k;
    }
    tr.incl = incl;
    const float ia = 1.0f / __ll2float_rn(area2 < 0 ? -area2 : area2);
    tr.inv_area = ia;
    const int xmin = min(X[0], min(X[1], X[2])), xmax = max(X[0], max(X[1], X[2]));
    const int ymin = min(Y[0], min(Y[1], Y[2])), ymax = max(Y[0], max(Y[1], Y[2]));
    const int R1 = (int)a.R - 1;
    const int x0 = max(0, (xmin + 127) >> 8), x1 = min(R1, (xmax - 128) >> 8);
    const int y0 = max(0, (ymin + 127) >> 8), y1 = min(R1, (ymax - 128) >> 8);
    if (x1 < x0 || y1 < y0) return 0;
    tr.x0 = (unsigned short)x0; tr.y0 = (unsigned short)y0;
    tr.w = (unsigned short)(x1 - x0 + 1); tr.h = (unsigned short)(y1 - y0 + 1);

    // texture LOD (GL 4.6 8.14): per-pixel steps of the mesh uv are constant per triangle
    float dudx = 0.f, dvdx = 0.f, dudy = 0.f, dvdy = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float ca = (float)Ak[k] * ia, cb = (float)Bk[k] * ia;
        dudx += uvx[k] * ca; dvdx += uvy[k] * ca;
        dudy += uvx[k] * cb; dvdy += uvy[k] * cb;
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        float lam = 0.f;
        const bool need = (m == 0) || (m == 1 && LT::kNeedNormal) || (m == 2 && LT::kNeedMR);
        const int ti = pr.tex[m];
        if (need && ti >= 0) {
            const float W = (float)a.texs[ti].w[0], H = (float)a.texs[ti].h[0];
            const float axx = dudx * W, bxx = dvdx * H, ayy = dudy * W, byy = dvdy * H;
            const float rx = sqrtf(axx * axx + bxx * bxx), ry = sqrtf(ayy * ayy + byy * byy);
            lam = log2f(fmaxf(rx, ry));
        }
        tf.lod[m] = lam;
    }
    return (uint32_t)tr.w * (uint32_t)tr.h;
}

// ------------------------------------------------------------------------------------------
// sampler: RGBA8 unorm, REPEAT, bilinear within a level, linear between levels
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 unpack_rgba8(uint32_t t) {
    const float k = 1.0f / 255.0f;
    return make_float4((float)(t & 0xffu) * k, (float)((t >> 8) & 0xffu) * k, (float)((t >> 16) & 0xffu) * k,
                       (float)(t >> 24) * k);
}
__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float t) {
    return make_float4(a.x + t * (b.x - a.x), a.y + t * (b.y - a.y), a.z + t * (b.z - a.z), a.w + t * (b.w - a.w));
}
__device__ __forceinline__ float4 bilinear(const uint32_t* __restrict__ lv, uint32_t W, uint32_t H, float u, float v) {
    // REPEAT: wrap in the normalised domain (exact for u in [0,1)), then fix the one-texel overhang
    u -= floorf(u);
    v -= floorf(v);
    const float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float ax = x - fx, ay = y - fy;
    int ix = (int)fx, iy = (int)fy;
    int x0 = ix < 0 ? (int)W - 1 : (ix >= (int)W ? ix - (int)W : ix);
    int y0 = iy < 0 ? (int)H - 1 : (iy >= (int)H ? iy - (int)H : iy);
    int x1 = x0 + 1 >= (int)W ? 0 : x0 + 1;
    int y1 = y0 + 1 >= (int)H ? 0 : y0 + 1;
    const uint32_t* r0 = lv + (size_t)y0 * W;
    const uint32_t* r1 = lv + (size_t)y1 * W;
    const uint32_t t00 = __ldg(r0 + x0), t10 = __ldg(r0 + x1), t01 = __ldg(r1 + x0), t11 = __ldg(r1 + x1);
    const float4 top = lerp4(unpack_rgba8(t00), unpack_rgba8(t10), ax);
    const float4 bot = lerp4(unpack_rgba8(t01), unpack_rgba8(t11), ax);
    return lerp4(top, bot, ay);
}
__device__ __forceinline__ float4 sample_trilinear(const DTexture& t, float u, float v, float lambda) {
    const int q = (int)t.nlevels - 1;
    if (!(lambda > 0.0f)) return bilinear(t.level[0], t.w[0], t.h[0], u, v);
    if (lambda >= (float)q) return bilinear(t.level[q], t.w[q], t.h[q], u, v);
    const float d = floorf(lambda), f = lambda - d;
    const int l = (int)d;
    const float4 a = bilinear(t.level[l], t.w[l], t.h[l], u, v);
    if (f == 0.0f) return a;
    const float4 b = bilinear(t.level[l + 1], t.w[l + 1], t.h[l + 1], u, v);
    return lerp4(a, b, f);
}

__device__ __forceinline__ float inv_sigmoid(float a) {  // utils.hpp:270
    a = fminf(fmaxf(a, 0.0f), 1.0f);
    return -logf(__fdiv_rn(1.0f, a + 1e-8f) - 1.0f);
}

// ------------------------------------------------------------------------------------------
// rounds over a flat candidate range of the triangles currently set up in shared memory
// ------------------------------------------------------------------------------------------
template <int LAYOUT>
__device__ void raster_rounds(const ConvertArgs& a, unsigned char* smem, const float4* tri_in, uint32_t ntri,
                              uint32_t cand_begin, uint32_t cand_end) {
    using LT = LayoutTraits<LAYOUT>;
    constexpr int kPitch = LT::kStagePitch;
    const TriRaster* rast = reinterpret_cast<const TriRaster*>(smem + SmemLayout::kRast);
    const TriFrag* frag = reinterpret_cast<const TriFrag*>(smem + SmemLayout::kFrag);
    const uint32_t* prefix = reinterpret_cast<const uint32_t*>(smem + SmemLayout::kPrefix);
    uint32_t* queue = reinterpret_cast<uint32_t*>(smem + SmemLayout::kQueueOff);
    Misc* misc = reinterpret_cast<Misc*>(smem + SmemLayout::misc(kPitch));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char* stage = smem + SmemLayout::kStage + warp * 32 * kPitch;

    for (uint32_t r0 = cand_begin; r0 < cand_end; r0 += kQueue) {
        const uint32_t r1 = min(r0 + (uint32_t)kQueue, cand_end);
        if (tid == 0) misc->qcount = 0;
        __syncthreads();
        // ---- phase A: coverage test + compaction -------------------------------------------
        for (uint32_t cb = r0 + warp * 32; cb < r1; cb += kThreads) {
            const uint32_t c = cb + lane;
            bool pass = false;
            uint32_t id = 0;
            if (c < r1) {
                uint32_t lo = 0, hi = ntri;  // largest s with prefix[s] <= c
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (prefix[mid] <= c) lo = mid; else hi = mid;
                }
                const TriRaster& tr = rast[lo];
                const uint32_t k = c - prefix[lo];
                const uint32_t w = tr.w;
                const uint32_t row = k / w, col = k - row * w;
                const int px = tr.x0 + col, py = tr.y0 + row;
                pass = true;
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    const long long E = tr.C[e] + (long long)tr.A[e] * px + (long long)tr.B[e] * py;
                    pass = pass && (E > 0 || (E == 0 && ((tr.incl >> e) & 1u)));
                }
                id = (lo << 24) | ((uint32_t)py << 12) | (uint32_t)px;
            }
            const unsigned m = __ballot_sync(0xffffffffu, pass);
            if (m) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&misc->qcount, (unsigned)__popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (pass) queue[base + __popc(m & ((1u << lane) - 1u))] = id;
            }
        }
        __syncthreads();
        const uint32_t qn = misc->qcount;
        if (tid == 0 && qn) misc->base = atomicAdd(a.counter, (unsigned long long)qn);
        __syncthreads();
        if (qn == 0) continue;
        const unsigned long long base = misc->base;

        // ---- phase B: fragment stage ---------------------------------------------------------
        for (uint32_t fb = warp * 32; fb < qn; fb += kThreads) {
            const uint32_t nfr = min(32u, qn - fb);
            unsigned long long key = 0;
            if ((uint32_t)lane < nfr) {
                const uint32_t id = queue[fb + lane];
                const uint32_t slot = id >> 24;
                const int py = (id >> 12) & 0xfff, px = id & 0xfff;
                const TriRaster& tr = rast[slot];
                const TriFrag& tf = frag[slot];
                float l[3];
#pragma unroll
                for (int e = 0; e < 3; ++e)
                    l[e] = __ll2float_rn(tr.C[e] + (long long)tr.A[e] * px + (long long)tr.B[e] * py) * tr.inv_area;
                const float4* v = tri_in + slot * 9;
                float at[12];
                {
                    const float4 a0 = v[0], a1 = v[1], a2 = v[2], b0 = v[3], b1 = v[4], b2 = v[5], c0 = v[6], c1 = v[7], c2 = v[8];
                    at[0] = l[0] * a0.x + l[1] * b0.x + l[2] * c0.x; at[1] = l[0] * a0.y + l[1] * b0.y + l[2] * c0.y;
                    at[2] = l[0] * a0.z + l[1] * b0.z + l[2] * c0.z; at[3] = l[0] * a0.w + l[1] * b0.w + l[2] * c0.w;
                    at[4] = l[0] * a1.x + l[1] * b1.x + l[2] * c1.x; at[5] = l[0] * a1.y + l[1] * b1.y + l[2] * c1.y;
                    at[6] = l[0] * a1.z + l[1] * b1.z + l[2] * c1.z; at[7] = l[0] * a1.w + l[1] * b1.w + l[2] * c1.w;
                    at[8] = l[0] * a2.x + l[1] * b2.x + l[2] * c2.x; at[9] = l[0] * a2.y + l[1] * b2.y + l[2] * c2.y;
                    at[10] = l[0] * a2.z + l[1] * b2.z + l[2] * c2.z; at[11] = l[0] * a2.w + l[1] * b2.w + l[2] * c2.w;
                }
                const DPrim& pr = a.prims[tf.prim];
                const float u = at[10], vv = at[11];
                // converterFS.glsl:55-62,99
                float4 col = make_float4(1.f, 1.f, 1.f, 1.f);
                const int ta = pr.tex[0];
                if (ta >= 0) col = sample_trilinear(a.texs[ta], u, vv, tf.lod[0]);
                col.x *= pr.factor[0]; col.y *= pr.factor[1]; col.z *= pr.factor[2]; col.w *= pr.factor[3];
                float* srec = reinterpret_cast<float*>(stage + lane * kPitch);
                if (LAYOUT == 0) {
                    // :64-81 normal
                    float nx = at[3], ny = at[4], nz = at[5];
                    const int tn = pr.tex[1];
                    if (tn >= 0) {
                        const float4 nm = sample_trilinear(a.texs[tn], u, vv, tf.lod[1]);
                        float rx = nm.x * 2.0f - 1.0f, ry = nm.y * 2.0f - 1.0f, rz = nm.z * 2.0f - 1.0f;
                        float inv = 1.0f / sqrtf(rx * rx + ry * ry + rz * rz);
                        rx *= inv; ry *= inv; rz *= inv;
                        const float tx = at[6], ty = at[7], tz = at[8], tw = at[9];
                        float bx = ny * tz - ty * nz, by = nz * tx - tz * nx, bz = nx * ty - tx * ny;  // cross(N,T)
                        inv = tw / sqrtf(bx * bx + by * by + bz * bz);
                        bx *= inv; by *= inv; bz *= inv;
                        inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
                        const float nnx = nx * inv, nny = ny * inv, nnz = nz * inv;
                        float ox = tx * rx + bx * ry + nnx * rz, oy = ty * rx + by * ry + nny * rz, oz = tz * rx + bz * ry + nnz * rz;
                        inv = 1.0f / sqrtf(ox * ox + oy * oy + oz * oz);
                        nx = ox * inv; ny = oy * inv; nz = oz * inv;
                    }
                    // :83-95 metallic-roughness (.bg)
                    float metal = 0.1f, rough = 0.5f;
                    const int tm = pr.tex[2];
                    if (tm >= 0) {
                        const float4 mr = sample_trilinear(a.texs[tm], u, vv, tf.lod[2]);
                        metal = mr.z; rough = mr.y;
                    }
                    float4* s4 = reinterpret_cast<float4*>(srec);
                    s4[0] = make_float4(at[0], at[1], at[2], 1.0f);
                    s4[1] = col;
                    s4[2] = make_float4(tf.scale[0], tf.scale[1], tf.scale[2], 0.0f);
                    s4[3] = make_float4(nx, ny, nz, 0.0f);
                    s4[4] = make_float4(tf.quat[0], tf.quat[1], tf.quat[2], tf.quat[3]);
                    s4[5] = make_float4(metal, rough, 0.0f, 1.0f);
                } else {
                    // parsers.cpp:484-499: SH0, opacity logit, log scale (per triangle)
                    const float kC0 = 0.28209479177387814f;  // SH_COEFF0, params.hpp:17
                    float2* s2 = reinterpret_cast<float2*>(srec);
                    s2[0] = make_float2(at[0], at[1]);
                    s2[1] = make_float2(at[2], tf.quat[0]);
                    s2[2] = make_float2(tf.quat[1], tf.quat[2]);
                    s2[3] = make_float2(tf.quat[3], tf.scale[0]);
                    s2[4] = make_float2(tf.scale[1], tf.scale[2]);
                    s2[5] = make_float2(__fdiv_rn(col.x - 0.5f, kC0), __fdiv_rn(col.y - 0.5f, kC0));
                    s2[6] = make_float2(__fdiv_rn(col.z - 0.5f, kC0), inv_sigmoid(col.w));
                }
                key = ((unsigned long long)tf.tri << 24) | ((unsigned long long)py << 12) | (unsigned long long)px;
            }
            __syncwarp();
            // ---- coalesced copy-out of this warp's contiguous span --------------------------
            const unsigned long long wbase = base + fb;  // first output index of this warp's span
            uint32_t nvalid = 0;                          // converterFS.glsl:48-51: idx >= cap dropped
            if (wbase < a.cap) nvalid = (uint32_t)min((unsigned long long)nfr, a.cap - wbase);
            if (LAYOUT == 0) {
                float4* dst = reinterpret_cast<float4*>(a.out + wbase * 96ull);
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const uint32_t c = lane + 32 * j, rec = c / 6, part = c - rec * 6;
                    if (rec < nvalid) dst[c] = *reinterpret_cast<const float4*>(stage + rec * kPitch + part * 16);
                }
            } else {
                float2* dst = reinterpret_cast<float2*>(a.out + wbase * 56ull);
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const uint32_t c = lane + 32 * j, rec = c / 7, part = c - rec * 7;
                    if (rec < nvalid) dst[c] = *reinterpret_cast<const float2*>(stage + rec * kPitch + part * 8);
                }
            }
            if (a.keys && (uint32_t)lane < nvalid) a.keys[wbase + lane] = key;
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <int LAYOUT>
__global__ void __launch_bounds__(kThreads, 2) convert_kernel(const __grid_constant__ ConvertArgs a) {
    using LT = LayoutTraits<LAYOUT>;
    constexpr int kPitch = LT::kStagePitch;
    extern __shared__ __align__(128) unsigned char smem[];
    float4* tri_buf[2] = {reinterpret_cast<float4*>(smem + SmemLayout::kTriIn),
                          reinterpret_cast<float4*>(smem + SmemLayout::kTriIn + kBatch * kTriBytes)};
    TriRaster* rast = reinterpret_cast<TriRaster*>(smem + SmemLayout::kRast);
    TriFrag* frag = reinterpret_cast<TriFrag*>(smem + SmemLayout::kFrag);
    uint32_t* prefix = reinterpret_cast<uint32_t*>(smem + SmemLayout::kPrefix);
    Misc* misc = reinterpret_cast<Misc*>(smem + SmemLayout::misc(kPitch));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned char* tri_bytes = reinterpret_cast<const unsigned char*>(a.tris);

    auto issue = [&](int stage, uint32_t b) {  // thread 0 only
        const uint32_t ntri = min((uint32_t)kBatch, a.tri_count - b * kBatch);
        const uint32_t bytes = ntri * kTriBytes;
        fence_proxy_async();
        mbar_arrive_expect_tx(&misc->bar[stage], bytes);
        tma_load_1d(tri_buf[stage], tri_bytes + ((size_t)a.tri_first + (size_t)b * kBatch) * kTriBytes, bytes,
                    &misc->bar[stage]);
    };

    if (tid == 0) {
        mbar_init(&misc->bar[0], 1);
        mbar_init(&misc->bar[1], 1);
        fence_barrier_init();
        const uint32_t b = atomicAdd(&a.sched[0], 1u);
        misc->batch[0] = (int)min(b, a.n_batches);
        if (b < a.n_batches) issue(0, b);
    }
    __syncthreads();

    int stage = 0;
    uint32_t phase[2] = {0, 0};
    while (true) {
        const uint32_t b = (uint32_t)misc->batch[stage];
        if (b >= a.n_batches) break;
        if (tid == 0) {  // claim + prefetch the next batch into the idle buffer
            const uint32_t nb = atomicAdd(&a.sched[0], 1u);
            misc->batch[stage ^ 1] = (int)min(nb, a.n_batches);
            if (nb < a.n_batches) issue(stage ^ 1, nb);
        }
        mbar_wait(&misc->bar[stage], phase[stage]);
        phase[stage] ^= 1;

        const uint32_t ntri = min((uint32_t)kBatch, a.tri_count - b * kBatch);
        const float4* tin = tri_buf[stage];
        // ---- per-triangle stage ----------------------------------------------------------------
        uint32_t cnt = 0;
        if (tid < kBatch) {
            if ((uint32_t)tid < ntri) {
                const uint32_t tg = a.tri_first + b * kBatch + tid;
                cnt = setup_triangle<LAYOUT>(tin + tid * 9, tg, a, rast[tid], frag[tid]);
                if (cnt > kBigCand) {  // defer: push chunks to the global queue
                    const uint32_t nch = (cnt + kChunkCand - 1) / kChunkCand;
                    uint32_t old = a.sched[2];
                    bool ok = false;
                    while (old + nch <= a.queue_cap) {
                        const uint32_t prev = atomicCAS(&a.sched[2], old, old + nch);
                        if (prev == old) { ok = true; break; }
                        old = prev;
                    }
                    if (ok) {
                        for (uint32_t i = 0; i < nch; ++i) a.queue[old + i] = make_uint2(tg, i);
                        __threadfence();
                        cnt = 0;
                    }
                }
            }
            // inclusive warp scan of cnt over the first 4 warps
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += n;
            }
            if (lane == 31) misc->wsum[warp] = incl;
            cnt = incl;  // keep inclusive value
        }
        __syncthreads();
        if (tid < kBatch) {
            uint32_t off = 0;
            for (int w = 0; w < warp; ++w) off += misc->wsum[w];
            prefix[tid + 1] = off + cnt;
            if (tid == 0) prefix[0] = 0;
        }
        __syncthreads();
        const uint32_t total = prefix[kBatch];
        raster_rounds<LAYOUT>(a, smem, tin, kBatch, 0, total);
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(&a.sched[1], 1u);
        }
        stage ^= 1;
    }

    // ---- drain: deferred big triangles, chunk by chunk, all CTAs ----------------------------
    if (tid == 0) {
        while (ld_acquire_u32(&a.sched[1]) < a.n_batches) __nanosleep(64);
        __threadfence();
    }
    __syncthreads();
    const uint32_t tail = ld_acquire_u32(&a.sched[2]);
    while (tail) {
        if (tid == 0) misc->item = atomicAdd(&a.sched[3], 1u);
        __syncthreads();
        const uint32_t it = misc->item;
        if (it >= tail) break;
        const uint2 item = a.queue[it];
        if (tid < 9) tri_buf[0][tid] = a.tris[(size_t)item.x * 9 + tid];
        __syncthreads();
        if (tid == 0) {
            const uint32_t c = setup_triangle<LAYOUT>(tri_buf[0], item.x, a, rast[0], frag[0]);
            prefix[0] = 0;
            prefix[1] = c;
        }
        __syncthreads();
        const uint32_t c0 = item.y * kChunkCand, c1 = min(prefix[1], c0 + kChunkCand);
        raster_rounds<LAYOUT>(a, smem, tri_buf[0], 1, c0, c1);
        __syncthreads();
    }

    // ---- last CTA out publishes the count and re-arms the scheduler for the next launch ------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const uint32_t done = atomicAdd(&a.sched[4], 1u);
        if (done == gridDim.x - 1) {
            __threadfence();
            *a.total_out = *reinterpret_cast<volatile unsigned long long*>(a.counter);
            *a.counter = 0ull;
            a.sched[0] = 0; a.sched[1] = 0; a.sched[2] = 0; a.sched[3] = 0; a.sched[4] = 0;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------------------------------
// mip chain: 2x2 box, round half up (matches oracle orc_mip_down)
// ------------------------------------------------------------------------------------------
__global__ void mip_down_kernel(const uint32_t* __restrict__ src, uint32_t sw, uint32_t sh, uint32_t* __restrict__ dst,
                                uint32_t dw, uint32_t dh) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const uint32_t x0 = min(2 * x, sw - 1), x1 = min(2 * x + 1, sw - 1), y0 = min(2 * y, sh - 1), y1 = min(2 * y + 1, sh - 1);
    const uint32_t a = src[(size_t)y0 * sw + x0], b = src[(size_t)y0 * sw + x1], c = src[(size_t)y1 * sw + x0],
                   d = src[(size_t)y1 * sw + x1];
    uint32_t o = 0;
#pragma unroll
    for (int s = 0; s < 32; s += 8) {
        const uint32_t sum = ((a >> s) & 0xff) + ((b >> s) & 0xff) + ((c >> s) & 0xff) + ((d >> s) & 0xff);
        o |= ((sum + 2) >> 2) << s;
    }
    dst[(size_t)y * dw + x] = o;
}

// ------------------------------------------------------------------------------------------
// .ply body rows from REF96 records (parsers.cpp:232-316,339-428,431-514)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned char to_byte(float v) {
    v = fminf(fmaxf(v, 0.0f), 1.0f);
    return (unsigned char)roundf(v * 255.0f);
}
__global__ void ply_rows_kernel(const float4* __restrict__ rec, unsigned long long count,
                                const unsigned long long* __restrict__ d_count, uint32_t format, float mult,
                                unsigned char* __restrict__ rows) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (d_count) count = min(count, *d_count);  // device-side count (enqueue-only path)
    if (i >= count) return;
    const float4 pos = rec[i * 6 + 0], col = rec[i * 6 + 1], sc = rec[i * 6 + 2], nrm = rec[i * 6 + 3], rot = rec[i * 6 + 4],
                 pbr = rec[i * 6 + 5];
    const float kC0 = 0.28209479177387814f;
    const float sh0 = __fdiv_rn(col.x - 0.5f, kC0), sh1 = __fdiv_rn(col.y - 0.5f, kC0), sh2 = __fdiv_rn(col.z - 0.5f, kC0);
    const float op = inv_sigmoid(col.w);
    const float lx = logf(__fmul_rn(sc.x, mult)), ly = logf(__fmul_rn(sc.y, mult)), lz = logf(__fmul_rn(sc.z, mult));
    if (format == 1) {
        float* f = reinterpret_cast<float*>(rows + i * 76ull);
        f[0] = pos.x; f[1] = pos.y; f[2] = pos.z; f[3] = nrm.x; f[4] = nrm.y; f[5] = nrm.z;
        f[6] = sh0; f[7] = sh1; f[8] = sh2; f[9] = pbr.x; f[10] = pbr.y; f[11] = op;
        f[12] = lx; f[13] = ly; f[14] = lz; f[15] = rot.x; f[16] = rot.y; f[17] = rot.z; f[18] = rot.w;
    } else if (format == 2) {
        unsigned char* p = rows + i * 48ull;
        float* f = reinterpret_cast<float*>(p);
        f[0] = pos.x; f[1] = pos.y; f[2] = pos.z;
        p[12] = to_byte(col.x); p[13] = to_byte(col.y); p[14] = to_byte(col.z); p[15] = to_byte(col.w);
        f[4] = rot.x; f[5] = rot.y; f[6] = rot.z; f[7] = rot.w;
        const float mn = (sc.y < sc.x) ? sc.y : sc.x;
        f[8] = lx; f[9] = ly; f[10] = logf(__fmul_rn(mn, mult));
        // octahedral normal (parsers.cpp:318-337)
        const float s = __fadd_rn(__fadd_rn(__fadd_rn(fabsf(nrm.x), fabsf(nrm.y)), fabsf(nrm.z)), 1e-8f);
        const float nx = __fdiv_rn(nrm.x, s), ny = __fdiv_rn(nrm.y, s), nz = __fdiv_rn(nrm.z, s);
        float rx, ry;
        if (nz >= 0.0f) { rx = nx; ry = ny; }
        else {
            const float m = (nx >= 0.0f && ny >= 0.0f) ? 1.0f : -1.0f;
            rx = __fmul_rn(__fsub_rn(1.0f, fabsf(ny)), m); ry = __fmul_rn(__fsub_rn(1.0f, fabsf(nx)), m);
        }
        const float ox = __fadd_rn(__fmul_rn(rx, 0.5f), 0.5f), oy = __fadd_rn(__fmul_rn(ry, 0.5f), 0.5f);
        p[44] = (unsigned char)fminf(fmaxf(roundf(__fmul_rn(ox, 255.0f)), 0.0f), 255.0f);
        p[45] = (unsigned char)fminf(fmaxf(roundf(__fmul_rn(oy, 255.0f)), 0.0f), 255.0f);
        p[46] = to_byte(pbr.y); p[47] = to_byte(pbr.x);
    } else {
        float* f = reinterpret_cast<float*>(rows + i * 248ull);
        f[0] = pos.x; f[1] = pos.y; f[2] = pos.z; f[3] = nrm.x; f[4] = nrm.y; f[5] = nrm.z;
        f[6] = sh0; f[7] = sh1; f[8] = sh2;
        for (int k = 9; k < 54; ++k) f[k] = 0.0f;
        f[54] = op; f[55] = lx; f[56] = ly; f[57] = lz; f[58] = rot.x; f[59] = rot.y; f[60] = rot.z; f[61] = rot.w;
    }
}

// ------------------------------------------------------------------------------------------
// launch wrappers used by m2s_api.cu
// ------------------------------------------------------------------------------------------
size_t convert_smem_bytes(int layout) {
    return layout == 0 ? SmemLayout::total(LayoutTraits<0>::kStagePitch) : SmemLayout::total(LayoutTraits<1>::kStagePitch);
}

cudaError_t convert_configure(int layout, int* blocks_per_sm) {
    cudaError_t e;
    const size_t smem = convert_smem_bytes(layout);
    if (layout == 0) {
        e = cudaFuncSetAttribute(convert_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, convert_kernel<0>, kThreads, smem);
    }
    e = cudaFuncSetAttribute(convert_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, convert_kernel<1>, kThreads, smem);
}

cudaError_t convert_launch(int layout, const ConvertArgs& args, int grid, cudaStream_t stream) {
    const size_t smem = convert_smem_bytes(layout);
    if (layout == 0) convert_kernel<0><<<grid, kThreads, smem, stream>>>(args);
    else convert_kernel<1><<<grid, kThreads, smem, stream>>>(args);
    return cudaGetLastError();
}

cudaError_t mip_down_launch(const uint32_t* src, uint32_t sw, uint32_t sh, uint32_t* dst, uint32_t dw, uint32_t dh,
                            cudaStream_t stream) {
    dim3 blk(32, 8), grd((dw + 31) / 32, (dh + 7) / 8);
    mip_down_kernel<<<grd, blk, 0, stream>>>(src, sw, sh, dst, dw, dh);
    return cudaGetLastError();
}

cudaError_t ply_rows_launch(const void* ref96, unsigned long long count, const unsigned long long* d_count,
                            uint32_t format, float mult, void* rows, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((count + 255) / 256);
    ply_rows_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(ref96), count, d_count, format, mult,
                                                reinterpret_cast<unsigned char*>(rows));
    return cudaGetLastError();
}

}  // namespace m2s
