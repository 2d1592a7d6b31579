// m2s_kernels.cu — the conversion pass as hand-written sm_100a CUDA.
//
// Two kernels on one stream replace the reference's geometry shader, fixed-function rasteriser,
// fragment shader and SSBO atomic append (converter{GS,FS}.glsl, ConversionPass.cpp:114-116); the data
// between them (8-byte fragment ids, 144-176 B per-triangle records) stays in the 126 MB L2.  The second
// kernel is launched with programmatic dependent launch.
//
// raster_kernel (persistent, one CTA per SM, every WARP an autonomous pipeline; one __syncthreads after the
// descriptor tables are copied to shared memory, none in the steady state):
//   work unit = <= 32 consecutive triangles; a warp's first unit is static (global warp id), further ones
//     are claimed from a global counter one unit ahead and their bytes prefetched into L2
//   TMA (cp.async.bulk + mbarrier complete_tx) stages the unit's 144 B/triangle into shared memory
//   per-triangle stage, one LANE per triangle (converterGS.glsl:326-443): longest edge, face normal,
//     dominant axis, orthographic uv, quaternion, UV->3D Jacobian scale; rasteriser set-up: 24.8
//     fixed-point window coords, int64 edge functions, top-left ownership bits, candidate pixel box
//     (clamped to the call's pixel-row band); exact barycentric state (edge functions at the box origin,
//     per-pixel steps, 1/area); the triangle's resolved sampler state (mip level pair, blend fraction,
//     level offsets and sizes)
//   the unit's per-triangle records leave shared memory as ONE TMA bulk store (cp.async.bulk.global.shared::cta)
//   coverage, three regimes by candidate-pixel count:
//     small  (<= 64, fits int32)  lane-per-triangle, lock-step incremental edge functions into a 64-bit
//                                 hit mask; warp scan -> contiguous range per triangle; second pass over
//                                 the set bits writes the ids (fragments leave TRIANGLE-MAJOR)
//     medium (<= 1024)            warp-per-triangle, 32 candidates per step, int64 edge functions
//     big                         pushed as 512-candidate chunks to a global queue (one atomicAdd per warp)
//                                 and rasterised by ALL warps of the grid after the units are set up
//     one global atomicAdd per unit (small) / per <= 512 fragments (medium, big) reserves the output range
//     (the reference: one atomicCounterIncrement per fragment).  Fragment i of the id list IS output record i.
// fragment_kernel (converterFS.glsl:44-104; lean registers, high occupancy, grid-stride):
//   a warp takes 32 consecutive fragments: exact barycentrics from the int64 edge functions, attributes
//   from the original vertices, all texel loads of all bound maps issued back to back, trilinear filter on
//   the FMA pipe (u8->f32 by PRMT+FADD), TBN normal, encode; the 32 records are transposed through shared
//   memory and written as one contiguous span (16-byte stores when aligned) — locally, or into every
//   rank's final buffer over NVLink (fused multi-GPU gather), or after the earlier chunks' records
//   (appended launches of the pipelined host path).
//
// Bit-exactness: every float operation of the per-triangle stage is written with __f*_rn intrinsics in the
// operation order of the oracle (and of GLM, which the reference's GLSL-as-C++ build uses): coverage is
// bit-exact and Scale/Quaternion match converterGS.glsl bit for bit.  Per-fragment values may use FMA
// contraction and are compared with a tolerance.
#include <cstdio>
#include "m2s_device.cuh"

// resident warps per SM / register cap per layout (warps are a multiple of 4: register allocation granularity)
#ifndef M2S_WARPS_REF96
#define M2S_WARPS_REF96 16
#define M2S_REGS_REF96 128
#endif
#ifndef M2S_WARPS_PACKED56
#define M2S_WARPS_PACKED56 16
#define M2S_REGS_PACKED56 128
#endif
#ifndef M2S_FRAG_THREADS
#define M2S_FRAG_THREADS 256
#endif

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
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
#ifdef M2S_TRACE
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define STAMP(a, slot) do { if ((a).trace && lane == 0) (a).trace[((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * 16 + (slot)] = gtime(); } while (0)
#define STAMPV(a, slot, v) do { if ((a).trace && lane == 0) (a).trace[((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * 16 + (slot)] = (v); } while (0)
#define TNOW() gtime()
#else
#define STAMPV(a, slot, v) do { } while (0)
#define TNOW() 0ull
#define STAMP(a, slot) do { } while (0)
#endif
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
constexpr int kSchedStride = 32;  // scheduler words live on separate 128-byte lines
#define SCHED(a, i) ((a).sched + (i) * kSchedStride)
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
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
// per-warp shared-memory records
// ------------------------------------------------------------------------------------------
struct __align__(16) TriRaster {  // 64 B — sign-normalised edge functions E_k(i,j) = A_k i + B_k j + C_k
    long long C[3];
    int A[3];
    int B[3];
    unsigned short x0, y0, w, h;  // candidate pixel box
    float inv_area;
    unsigned incl;                // bit k: edge k owns its E == 0 samples (top-left rule)
};
struct __align__(8) TexRef {  // 16 B — one map, resolved for one triangle
    uint32_t off0, off1;          // texel offsets of the two mip levels in the arena; off0 == ~0u: no map
    unsigned short w0, h0, w1, h1;
};
template <int NMAPS>
struct __align__(16) TriFragT {  // 128 + 16*NMAPS bytes
    float quat[4];    // (w,x,y,z)
    float scale[3];   // raw (REF96) or log(scale * sigma/R)
    unsigned tri;     // global triangle index
    float factor[4];  // u_materialFactor
    float frac[3];    // trilinear blend per map (0 => single level)
    unsigned meta;    // bits 0-2: map m has the same level sizes as map 0 (=> same footprint and weights);
                      // bits 4-15: x0, bits 16-27: y0 of the candidate pixel box
    // exact barycentrics: lambda_k(px,py) = (E0_k + A_k (px-x0) + B_k (py-y0)) * inv_area  (GL 4.6 eq. 14.9, w = 1)
    long long E0[3];
    int A[3];
    int B[3];
    float inv_area;
    unsigned pad;
    TexRef tex[NMAPS];
};

template <int LAYOUT>
struct Cfg;
template <>
struct Cfg<0> {  // REF96
    static constexpr int kStride = 96;
    static constexpr int kPitch = 96;   // = stride: the warp's 32 records are one contiguous 3 KB span
    static constexpr int kMaps = 3;
    static constexpr bool kLogScale = false;
    // warps are allocated registers in groups of 4, so the warp count is a multiple of 4
    static constexpr int kWarps = M2S_WARPS_REF96;
    static constexpr int kMaxRegs = M2S_REGS_REF96;   // warps * 32 * regs <= 64 K registers
};
template <>
struct Cfg<1> {  // PACKED56
    static constexpr int kStride = 56;
    static constexpr int kPitch = 56;
    static constexpr int kMaps = 1;
    static constexpr bool kLogScale = true;
    static constexpr int kWarps = M2S_WARPS_PACKED56;
    static constexpr int kMaxRegs = M2S_REGS_PACKED56;
};

template <int LAYOUT>
struct __align__(128) WarpBlock {
    float4 tri[kUnitTris * 9];                   // 4608 B, TMA destination
    TriFragT<Cfg<LAYOUT>::kMaps> frag[kUnitTris];
    uint32_t queue[kQueue];                      // 2048 B: slot << 24 | y << 12 | x
    uint64_t bar;
};

// The scene's descriptor tables as the set-up sees them: in shared memory when they fit (each CTA copies
// them once — 2368 warps chasing range -> primitive -> texture through the same few L2 lines cost ~1 us per
// dependent step), else in global memory.
struct Tables {
    const DRange* ranges;
    const DPrim* prims;
    const DTexture* texs;
    uint32_t nranges;
};
constexpr uint32_t kTableSmemBytes = 24 * 1024;

// ------------------------------------------------------------------------------------------
// per-triangle stage + rasteriser set-up.  t4: 9 float4 in shared memory.
// Returns the number of candidate pixels (0 => nothing to rasterise).
// ------------------------------------------------------------------------------------------
// M2S_INLINE_SETUP (tuning build, off): inline the per-triangle stage into the unit loop — the raster state then lives
// in registers instead of going through local memory (the default build shows 36 STL / 28 LDL around the call) — and
// keep one out-of-line copy for the drain path.  Inlined at both sites it was 90 KB of code, hence the split.
#ifdef M2S_INLINE_SETUP
#define M2S_SETUP_QUAL __forceinline__
#else
#define M2S_SETUP_QUAL __noinline__
#endif
template <int LAYOUT>
__device__ M2S_SETUP_QUAL uint32_t setup_triangle(const float4* __restrict__ t4, uint32_t tri_global, const ConvertArgs& a,
                                   const Tables& tb, TriRaster& tr, TriFragT<Cfg<LAYOUT>::kMaps>& tf) {
    using C = Cfg<LAYOUT>;
    tr.w = 0; tr.h = 0; tr.x0 = 0; tr.y0 = 0; tr.incl = 0; tr.inv_area = 0.f;
    tf.tri = tri_global;
    // triangle -> primitive (sorted disjoint ranges)
    int lo = 0, hi = (int)tb.nranges - 1, found = -1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const DRange r = tb.ranges[mid];
        if (tri_global < r.first) hi = mid - 1;
        else if (tri_global >= r.end) lo = mid + 1;
        else { found = (int)r.prim; break; }
    }
    if (found < 0) return 0;
    const DPrim pr = tb.prims[found];

    // vertex data: 3 x {pos3 nrm3 tan4 uv2} = 9 float4
    const float4 q0 = t4[0], q3 = t4[3], q6 = t4[6];
    const float4 q2 = t4[2], q5 = t4[5], q8 = t4[8];
    const f3 P0 = {q0.x, q0.y, q0.z}, P1 = {q3.x, q3.y, q3.z}, P2 = {q6.x, q6.y, q6.z};

    // converterGS.glsl:327-347
    f3 e1 = sub3(P1, P0), e2 = sub3(P2, P0), e3 = sub3(P2, P1);
    const float l1 = len3(e1), l2 = len3(e2), l3 = len3(e3);
    if (l2 > l1 && l2 > l3) { f3 tmp = e1; e1 = e2; e2 = tmp; }
    else if (l3 > l1 && l3 > l2) { e1 = e3; }
    e1 = norm3(e1);
    const f3 n = norm3(cross3(e1, e2));
    const float ax = fabsf(n.x), ay = fabsf(n.y), az = fabsf(n.z);
    const int axis = (ax > ay && ax > az) ? 0 : ((ay > az) ? 1 : 2);

    // :354-397 orthogonal uv: X -> (y,z), Y -> (x,z), Z -> (x,y), over max(range_a, range_b)
    float ou[3], ov[3];
    {
        const float mina = axis == 0 ? pr.bmin[1] : pr.bmin[0], maxa = axis == 0 ? pr.bmax[1] : pr.bmax[0];
        const float minb = axis == 2 ? pr.bmin[1] : pr.bmin[2], maxb = axis == 2 ? pr.bmax[1] : pr.bmax[2];
        const float ra = __fsub_rn(maxa, mina), rb = __fsub_rn(maxb, minb);
        const float range = (ra < rb) ? rb : ra;
        const float pa0 = axis == 0 ? P0.y : P0.x, pa1 = axis == 0 ? P1.y : P1.x, pa2 = axis == 0 ? P2.y : P2.x;
        const float pb0 = axis == 2 ? P0.y : P0.z, pb1 = axis == 2 ? P1.y : P1.z, pb2 = axis == 2 ? P2.y : P2.z;
        ou[0] = __fdiv_rn(__fsub_rn(pa0, mina), range); ov[0] = __fdiv_rn(__fsub_rn(pb0, minb), range);
        ou[1] = __fdiv_rn(__fsub_rn(pa1, mina), range); ov[1] = __fdiv_rn(__fsub_rn(pb1, minb), range);
        ou[2] = __fdiv_rn(__fsub_rn(pa2, mina), range); ov[2] = __fdiv_rn(__fsub_rn(pb2, minb), range);
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
        // Scale and Quaternion are bit-identical to converterGS.glsl's (tests: golden GS vectors); only the
        // log of the packed layout is the device's logf
        const float sx = len3(Ju), sy = len3(Jv), sz = 1e-7f;
        if (C::kLogScale) {  // parsers.cpp:497-499 log(scale * sigma/R)
            tf.scale[0] = logf(__fmul_rn(sx, a.mult)); tf.scale[1] = logf(__fmul_rn(sy, a.mult)); tf.scale[2] = a.log_sz;
        } else { tf.scale[0] = sx; tf.scale[1] = sy; tf.scale[2] = sz; }
    }
    tf.factor[0] = pr.factor[0]; tf.factor[1] = pr.factor[1]; tf.factor[2] = pr.factor[2]; tf.factor[3] = pr.factor[3];

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
    const long long area2 = (long long)(X[1] - X[0]) * (Y[2] - Y[0]) - (long long)(X[2] - X[0]) * (Y[1] - Y[0]);
    if (area2 == 0) return 0;
    const int sg = area2 < 0 ? -1 : 1;
    unsigned incl = 0;
    int Ak[3], Bk[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int va = (k + 1) % 3, vb = (k + 2) % 3;
        const int dx = X[vb] - X[va], dy = Y[vb] - Y[va];
        const int A = sg * (-dy * 256), B = sg * (dx * 256);
        tr.A[k] = Ak[k] = A;
        tr.B[k] = Bk[k] = B;
        tr.C[k] = (long long)sg * ((long long)dx * (128 - Y[va]) - (long long)dy * (128 - X[va]));
        if (A > 0 || (A == 0 && B > 0)) incl |= 1u << k;
    }
    tr.incl = incl;
    const float ia = 1.0f / __ll2float_rn(area2 < 0 ? -area2 : area2);
    tr.inv_area = ia;
    const int xmin = min(X[0], min(X[1], X[2])), xmax = max(X[0], max(X[1], X[2]));
    const int ymin = min(Y[0], min(Y[1], Y[2])), ymax = max(Y[0], max(Y[1], Y[2]));
    const int R1 = (int)a.R - 1;
    const int x0 = max(0, (xmin + 127) >> 8), x1 = min(R1, (xmax - 128) >> 8);
    const int y0 = max((int)a.row_begin, (ymin + 127) >> 8), y1 = min((int)a.row_end - 1, (ymax - 128) >> 8);  // row band
    if (x1 < x0 || y1 < y0) return 0;
    tr.x0 = (unsigned short)x0; tr.y0 = (unsigned short)y0;
    tr.w = (unsigned short)(x1 - x0 + 1); tr.h = (unsigned short)(y1 - y0 + 1);

    // barycentric state for the fragment kernel (exact integers, relative to the box origin) and the
    // per-pixel steps of the mesh uv (constant per triangle: uv is affine in window space)
    float dudx = 0.f, dvdx = 0.f, dudy = 0.f, dvdy = 0.f;
    {
        const float uvx[3] = {q2.z, q5.z, q8.z}, uvy[3] = {q2.w, q5.w, q8.w};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            tf.E0[k] = tr.C[k] + (long long)Ak[k] * x0 + (long long)Bk[k] * y0;
            tf.A[k] = Ak[k]; tf.B[k] = Bk[k];
            const float ca = (float)Ak[k] * ia, cb = (float)Bk[k] * ia;
            dudx += uvx[k] * ca; dvdx += uvy[k] * ca;
            dudy += uvx[k] * cb; dvdy += uvy[k] * cb;
        }
        tf.inv_area = ia;
        tf.pad = 0;
    }

    // sampler state (GL 4.6 8.14): the steps of the mesh uv are constant per triangle, so lambda, the
    // level pair and the blend fraction are too
    unsigned share = 0;
    TexRef ref0;
#pragma unroll
    for (int m = 0; m < C::kMaps; ++m) {
        TexRef ref;
        ref.off0 = 0xffffffffu; ref.off1 = 0; ref.w0 = ref.h0 = ref.w1 = ref.h1 = 1;
        float frac = 0.f;
        const int ti = pr.tex[m];
        if (ti >= 0) {
            const DTexture t = tb.texs[ti];
            const float W = (float)t.w[0], H = (float)t.h[0];
            const float axx = dudx * W, bxx = dvdx * H, ayy = dudy * W, byy = dvdy * H;
            const float lam = 0.5f * __log2f(fmaxf(axx * axx + bxx * bxx, ayy * ayy + byy * byy));  // log2 of the longer step
            const int q = (int)t.nlevels - 1;
            int l0 = 0;
            if (!(lam > 0.0f)) { l0 = 0; }                       // magnification: LINEAR on level 0
            else if (lam >= (float)q) { l0 = q; }                 // clamped to the last level
            else { const float d = floorf(lam); l0 = (int)d; frac = lam - d; }
            const int l1 = min(l0 + 1, q);
            ref.off0 = t.off[l0]; ref.off1 = t.off[l1];
            ref.w0 = t.w[l0]; ref.h0 = t.h[l0]; ref.w1 = t.w[l1]; ref.h1 = t.h[l1];
        }
        if (m == 0) ref0 = ref;
        else if (ti >= 0 && ref0.off0 != 0xffffffffu && ref.w0 == ref0.w0 && ref.h0 == ref0.h0 && ref.w1 == ref0.w1 &&
                 ref.h1 == ref0.h1 && frac == tf.frac[0])
            share |= 1u << m;
        tf.tex[m] = ref;
        tf.frac[m] = frac;
    }
    tf.meta = share | ((unsigned)x0 << 4) | ((unsigned)y0 << 16);
    return (uint32_t)tr.w * (uint32_t)tr.h;
}
#ifdef M2S_INLINE_SETUP
template <int LAYOUT>
__device__ __noinline__ uint32_t setup_triangle_outofline(const float4* __restrict__ t4, uint32_t tri_global, const ConvertArgs& a,
                                                          const Tables& tb, TriRaster& tr, TriFragT<Cfg<LAYOUT>::kMaps>& tf) {
    return setup_triangle<LAYOUT>(t4, tri_global, a, tb, tr, tf);
}
#define M2S_SETUP_DRAIN setup_triangle_outofline
#else
#define M2S_SETUP_DRAIN setup_triangle
#endif

// ------------------------------------------------------------------------------------------
// sampler: RGBA8 unorm, REPEAT, bilinear within a level, linear between levels
// ------------------------------------------------------------------------------------------
struct Bilin {          // one bilinear footprint: texel indices relative to the level start + weights
    uint32_t i00, i10, i01, i11;
    float w00, w10, w01, w11;  // already scaled by 1/255
};
// small non-negative int -> float on the FMA pipe (no I2F): 2^23 | v is the float 2^23 + v
__device__ __forceinline__ float u2f(uint32_t v) { return __uint_as_float(0x4B000000u | v) - 8388608.0f; }
// floor for |x| < 2^22 on the FMA pipe: round-to-nearest of x - 0.5 via the 1.5*2^23 trick.  At exact
// integers it may return x - 1 with fraction 1, which selects the same texels with the same weights.
__device__ __forceinline__ float fast_floor(float x, int& i) {
    const float t = (x - 0.5f) + 12582912.0f;
    i = __float_as_int(t) - 0x4B400000;
    return t - 12582912.0f;
}
__device__ __forceinline__ Bilin bilin_setup(uint32_t W, uint32_t H, float u, float v) {
    // REPEAT: wrap in the normalised domain (exact for u in [0,1)), then fix the one-texel overhang
    int ix, iy;
    u -= fast_floor(u, ix);
    v -= fast_floor(v, iy);
    const float x = u * u2f(W) - 0.5f, y = v * u2f(H) - 0.5f;
    const float fx = fast_floor(x, ix), fy = fast_floor(y, iy);
    const float ax = x - fx, ay = y - fy;
    // ix in [-1, W]: wrap, then clamp so that even NaN/huge uv can never index outside the level
    int x0 = ix < 0 ? ix + (int)W : ix;
    int y0 = iy < 0 ? iy + (int)H : iy;
    x0 = min(max(x0, 0), (int)W - 1);
    y0 = min(max(y0, 0), (int)H - 1);
    const int x1 = x0 + 1 >= (int)W ? 0 : x0 + 1;
    const int y1 = y0 + 1 >= (int)H ? 0 : y0 + 1;
    Bilin b;
    const uint32_t r0 = (uint32_t)y0 * W, r1 = (uint32_t)y1 * W;
    b.i00 = r0 + x0; b.i10 = r0 + x1; b.i01 = r1 + x0; b.i11 = r1 + x1;
    const float k = 1.0f / 255.0f;
    const float bx = 1.0f - ax, by = (1.0f - ay) * k, cy = ay * k;
    b.w00 = bx * by; b.w10 = ax * by; b.w01 = bx * cy; b.w11 = ax * cy;
    return b;
}
// byte c of a texel as float, without the conversion pipe: 0x4B000000 | byte is 2^23 + byte
template <int CH>
__device__ __forceinline__ float tex_ch(uint32_t t) {
    return __uint_as_float(__byte_perm(t, 0x4B000000u, 0x7440u | CH)) - 8388608.0f;
}
template <int CH>
__device__ __forceinline__ float filt(const Bilin& b, uint32_t t00, uint32_t t10, uint32_t t01, uint32_t t11) {
    return b.w00 * tex_ch<CH>(t00) + b.w10 * tex_ch<CH>(t10) + b.w01 * tex_ch<CH>(t01) + b.w11 * tex_ch<CH>(t11);
}

__device__ __forceinline__ float inv_sigmoid(float a) {  // utils.hpp:270
    a = fminf(fmaxf(a, 0.0f), 1.0f);
    return -logf(__fdiv_rn(1.0f, a + 1e-8f) - 1.0f);
}

// ------------------------------------------------------------------------------------------
// flush: reserve the output range and write the warp's fragment ids (coalesced)
// ------------------------------------------------------------------------------------------
template <int LAYOUT>
__device__ __forceinline__ void flush_ids(const ConvertArgs& a, WarpBlock<LAYOUT>& wb, uint32_t qn, int lane) {
    if (qn == 0) return;
    __syncwarp();
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(a.counter, (unsigned long long)qn);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (uint32_t i = lane; i < qn; i += 32) {
        const uint32_t id = wb.queue[i];
        const unsigned long long idx = base + i;
        if (idx < a.cap)  // converterFS.glsl:48-51: the counter keeps counting, records beyond the cap are dropped
            a.frag_ids[idx] = make_uint2(wb.frag[id >> 24].tri, id & 0xffffffu);
    }
    __syncwarp();
}

// enqueue the lanes whose `inside` is set; flush when the queue could overflow on the next step
template <int LAYOUT>
__device__ __forceinline__ void enqueue(const ConvertArgs& a, WarpBlock<LAYOUT>& wb, uint32_t& qn, bool inside, uint32_t id,
                                        int lane) {
    const unsigned m = __ballot_sync(0xffffffffu, inside);
    if (m) {
        if (inside) wb.queue[qn + __popc(m & ((1u << lane) - 1u))] = id;
        qn += __popc(m);
        if (qn > kQueue - 32) {
            flush_ids<LAYOUT>(a, wb, qn, lane);
            qn = 0;
        }
    }
}

// warp-per-triangle coverage of candidates [c0, c1) of the triangle in `slot` (int64 edge functions)
#ifdef M2S_INLINE_RASTER  // tuning build (off): with the set-up inlined too, the raster state never needs an address
#define M2S_RASTER_QUAL __forceinline__
#else
#define M2S_RASTER_QUAL
#endif
template <int LAYOUT>
__device__ M2S_RASTER_QUAL void raster_one(const ConvertArgs& a, WarpBlock<LAYOUT>& wb, uint32_t& qn, uint32_t slot, uint32_t c0, uint32_t c1,
                           int lane, const TriRaster& mine) {
    // the raster state lives in the registers of lane `slot`: broadcast it
    const unsigned full = 0xffffffffu;
    const int src = (int)slot;
    const uint32_t w = __shfl_sync(full, (uint32_t)mine.w, src);
    const float rcp = 1.0f / (float)w;
    const long long C0 = __shfl_sync(full, mine.C[0], src), C1 = __shfl_sync(full, mine.C[1], src), C2 = __shfl_sync(full, mine.C[2], src);
    const int A0 = __shfl_sync(full, mine.A[0], src), A1 = __shfl_sync(full, mine.A[1], src), A2 = __shfl_sync(full, mine.A[2], src);
    const int B0 = __shfl_sync(full, mine.B[0], src), B1 = __shfl_sync(full, mine.B[1], src), B2 = __shfl_sync(full, mine.B[2], src);
    const unsigned incl = __shfl_sync(full, mine.incl, src);
    const int bx = __shfl_sync(full, (int)mine.x0, src), by = __shfl_sync(full, (int)mine.y0, src);
    for (uint32_t cb = c0; cb < c1; cb += 32) {
        const uint32_t c = cb + lane;
        bool inside = false;
        uint32_t id = 0;
        if (c < c1) {
            uint32_t row = (uint32_t)((float)c * rcp);  // c < 2^24: exact in fp32, quotient off by at most 1
            int col = (int)(c - row * w);
            if (col < 0) { --row; col += (int)w; }
            else if (col >= (int)w) { ++row; col -= (int)w; }
            const int px = bx + col, py = by + (int)row;
            const long long E0 = C0 + (long long)A0 * px + (long long)B0 * py;
            const long long E1 = C1 + (long long)A1 * px + (long long)B1 * py;
            const long long E2 = C2 + (long long)A2 * px + (long long)B2 * py;
            inside = (E0 > 0 || (E0 == 0 && (incl & 1u))) && (E1 > 0 || (E1 == 0 && (incl & 2u))) &&
                     (E2 > 0 || (E2 == 0 && (incl & 4u)));
            id = (slot << 24) | ((uint32_t)py << 12) | (uint32_t)px;
        }
        enqueue<LAYOUT>(a, wb, qn, inside, id, lane);
    }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <int LAYOUT>
__global__ void __launch_bounds__(Cfg<LAYOUT>::kWarps * 32) __maxnreg__(Cfg<LAYOUT>::kMaxRegs) raster_kernel(const __grid_constant__ ConvertArgs a) {
    using C = Cfg<LAYOUT>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpBlock<LAYOUT>& wb = *reinterpret_cast<WarpBlock<LAYOUT>*>(smem + (size_t)warp * sizeof(WarpBlock<LAYOUT>));
    const unsigned char* tri_bytes = reinterpret_cast<const unsigned char*>(a.tris);

#ifdef M2S_EARLY_TRIGGER
    // PDL early trigger (one fragment-kernel CTA per SM becomes resident beside this CTA and parks in
    // griddepcontrol.wait).  Measured SLOWER (PACKED56 40.5 vs 39.8 us, REF96 62.4 vs 59.5 us): off.
    asm volatile("griddepcontrol.launch_dependents;");
#endif
    if (lane == 0) {
        mbar_init(&wb.bar, 1);
        fence_barrier_init();
    }
    // descriptor tables -> shared memory (once per CTA) when they fit
    Tables tabs{a.ranges, a.prims, a.texs, a.nranges};
    {
        const uint32_t br = a.nranges * (uint32_t)sizeof(DRange), bp = a.nprims * (uint32_t)sizeof(DPrim), bt = a.ntex * (uint32_t)sizeof(DTexture);
#ifdef M2S_NO_TABLES
        if (false) {
#else
        if (br + bp + bt <= kTableSmemBytes) {  // uniform across the grid
#endif
            unsigned char* base = smem + (size_t)C::kWarps * sizeof(WarpBlock<LAYOUT>);
            uint32_t* dst = reinterpret_cast<uint32_t*>(base);  // word-wise: the structs are 16, 56 and 48 bytes
            const uint32_t* s0 = reinterpret_cast<const uint32_t*>(a.ranges);
            const uint32_t* s1 = reinterpret_cast<const uint32_t*>(a.prims);
            const uint32_t* s2 = reinterpret_cast<const uint32_t*>(a.texs);
            const uint32_t n0 = br / 4, n1 = bp / 4, n2 = bt / 4;
            for (uint32_t i = threadIdx.x; i < n0 + n1 + n2; i += blockDim.x)
                dst[i] = i < n0 ? s0[i] : (i < n0 + n1 ? s1[i - n0] : s2[i - n0 - n1]);
            tabs.ranges = reinterpret_cast<const DRange*>(base);
            tabs.prims = reinterpret_cast<const DPrim*>(base + br);
            tabs.texs = reinterpret_cast<const DTexture*>(base + br + bp);
            __syncthreads();  // the only CTA-wide barrier before the end of the kernel
        }
    }
    __syncwarp();
    uint32_t phase = 0, qn = 0;
    STAMP(a, 0);

    // ---- work units ---------------------------------------------------------------------------
    // the first unit of every warp is static (unit = global warp id): no atomic, and nobody can grab two
    // units while a neighbour gets none; further units are claimed dynamically one unit ahead
    const uint32_t nwarps_total = gridDim.x * (blockDim.x >> 5);
    uint32_t unit = blockIdx.x + gridDim.x * warp;  // warp w of every CTA before warp w+1 of any: SMs fill evenly
    while (unit < a.n_units) {
        const uint32_t t0 = unit * a.unit_tris;
        const uint32_t ntri = min(a.unit_tris, a.tri_count - t0);
        uint32_t next = 0xffffffffu;
        if (lane == 0) {
            const uint32_t bytes = ntri * kTriBytes;
            tma_store_wait_read();  // the previous unit's record stores have finished reading this slice
            fence_proxy_async();
            mbar_arrive_expect_tx(&wb.bar, bytes);
            tma_load_1d(wb.tri, tri_bytes + ((size_t)a.tri_first + t0) * kTriBytes, bytes, &wb.bar);
            if (a.n_units > nwarps_total) {  // more units than warps: claim the next one now, pull its bytes into L2
                next = nwarps_total + atomicAdd(SCHED(a, 0), 1u);
                if (next < a.n_units) {
                    const uint32_t nt0 = next * a.unit_tris;
                    prefetch_l2(tri_bytes + ((size_t)a.tri_first + nt0) * kTriBytes, min(a.unit_tris, a.tri_count - nt0) * kTriBytes);
                }
            }
        }
        mbar_wait(&wb.bar, phase);
        phase ^= 1;
        STAMP(a, 1);

        // per-triangle stage: one lane per triangle
        uint32_t cnt = 0;
        TriRaster tr;  // raster state stays with the lane that owns the triangle
        tr.w = 0; tr.h = 0; tr.x0 = 0; tr.y0 = 0; tr.incl = 0;
        tr.A[0] = tr.A[1] = tr.A[2] = tr.B[0] = tr.B[1] = tr.B[2] = 0; tr.C[0] = tr.C[1] = tr.C[2] = 0;
        if ((uint32_t)lane < ntri) cnt = setup_triangle<LAYOUT>(wb.tri + lane * 9, a.tri_first + t0 + lane, a, tabs, tr, wb.frag[lane]);

        // classify: small (lane-per-triangle, int32), medium (warp-per-triangle), big (deferred)
        int e0 = 0, e1 = 0, e2 = 0, a0 = 0, a1 = 0, a2 = 0, r0 = 0, r1 = 0, r2 = 0, w = 1, bx = 0, by = 0;
        bool small = false, deferred = false;
        if (cnt) {
            w = tr.w; bx = tr.x0; by = tr.y0;
            const int h = tr.h;
            a0 = tr.A[0]; a1 = tr.A[1]; a2 = tr.A[2];
            const int b0 = tr.B[0], b1 = tr.B[1], b2 = tr.B[2];
            // E at the box origin, with the ownership bias folded in: inside <=> all E' >= 0
            const long long E0 = tr.C[0] + (long long)a0 * bx + (long long)b0 * by - ((tr.incl & 1u) ? 0 : 1);
            const long long E1 = tr.C[1] + (long long)a1 * bx + (long long)b1 * by - ((tr.incl & 2u) ? 0 : 1);
            const long long E2 = tr.C[2] + (long long)a2 * bx + (long long)b2 * by - ((tr.incl & 4u) ? 0 : 1);
            const long long lim = 0x7fffffffll;
            const long long s0 = llabs(E0) + (long long)(w - 1) * abs(a0) + (long long)(h - 1) * abs(b0);
            const long long s1 = llabs(E1) + (long long)(w - 1) * abs(a1) + (long long)(h - 1) * abs(b1);
            const long long s2 = llabs(E2) + (long long)(w - 1) * abs(a2) + (long long)(h - 1) * abs(b2);
            small = cnt <= kSmallCand && s0 < lim && s1 < lim && s2 < lim;
            if (small) {
                e0 = (int)E0; e1 = (int)E1; e2 = (int)E2;
                r0 = b0 - (w - 1) * a0; r1 = b1 - (w - 1) * a1; r2 = b2 - (w - 1) * a2;  // step to the next row's first pixel
            }
        }
        // big triangles: push their chunks to the global queue.  ONE atomicAdd per warp reserves the slots
        // (a per-lane CAS loop collapses under contention: 15 k simultaneous pushers cost 19 ms)
        {
            const bool big = cnt > kBigCand && !small;
            const uint32_t nch = big ? (cnt + kChunkCand - 1) / kChunkCand : 0u;
            uint32_t incl = nch;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            const uint32_t wtotal = __shfl_sync(0xffffffffu, incl, 31);
            if (wtotal) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(SCHED(a, 2), wtotal);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (big) {
                    const uint32_t first = base + (incl - nch);
                    const uint32_t tg = a.tri_first + t0 + lane;
                    if (first + nch <= a.queue_cap) {
                        for (uint32_t i = 0; i < nch; ++i) a.queue[first + i] = make_uint2(tg, i);
                        cnt = 0;
                        deferred = true;
                    } else {  // queue full: the in-range slots become no-ops, the triangle is rasterised here
                        for (uint32_t i = first; i < min(first + nch, a.queue_cap); ++i) a.queue[i] = make_uint2(0xffffffffu, 0u);
                    }
                    __threadfence();
                }
            }
        }
        STAMP(a, 2);
        __syncwarp();
        // every chunk this unit defers is in the global queue now (pushers fenced): count the unit as
        // "past set-up" so idle warps only wait for set-ups in flight, not for whole units
        if (lane == 0) atomicAdd(SCHED(a, 1), 1u);
        // the unit's per-triangle records (barycentric + shading state) go to global memory for the fragment
        // kernel: one TMA bulk store straight out of this warp's shared-memory slice
        if (__any_sync(0xffffffffu, cnt != 0 || deferred)) {
            if (lane == 0) {
                fence_proxy_async();
                tma_store_1d(a.tri_frag + (size_t)t0 * sizeof(TriFragT<C::kMaps>), wb.frag, ntri * (uint32_t)sizeof(TriFragT<C::kMaps>));
                tma_store_commit();
            }
        }

        // small triangles: every lane walks its own pixel box in lock-step, twice.  Walk 1 records the
        // covered candidates in a 64-bit mask (no ballots, no stores); a warp scan of the hit counts gives
        // every triangle a contiguous output range; walk 2 writes the ids there.  Fragments therefore
        // leave TRIANGLE-MAJOR, which is what keeps the fragment kernel's record and texel loads coherent.
        {
            const uint32_t mine = small ? cnt : 0u;
            uint32_t maxc = mine;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, d));
            unsigned long long hits = 0;
            {
                int col = 0, f0 = e0, f1 = e1, f2 = e2;
                for (uint32_t it = 0; it < maxc; ++it) {
                    const bool inside = it < mine && (f0 | f1 | f2) >= 0;
                    hits |= (unsigned long long)inside << it;
                    if (++col == w) { col = 0; f0 += r0; f1 += r1; f2 += r2; }
                    else { f0 += a0; f1 += a1; f2 += a2; }
                }
            }
            STAMP(a, 3);
            const uint32_t nh = (uint32_t)__popcll(hits);
            uint32_t incl = nh;  // inclusive warp scan
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (total) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.counter, (unsigned long long)total);
                base = __shfl_sync(0xffffffffu, base, 0);
                STAMP(a, 4);
                unsigned long long idx = base + (incl - nh);
                const uint32_t tg = a.tri_first + t0 + lane;
                // walk 2 visits only the hits: bit b of the mask is candidate b = row*w + col; row = b/w by a
                // 16.16 reciprocal (exact for b < 64, w <= 64)
                const uint32_t inv = (65536u + (uint32_t)w - 1u) / (uint32_t)w;
                uint32_t maxh = nh;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) maxh = max(maxh, __shfl_xor_sync(0xffffffffu, maxh, d));
                unsigned long long m = hits;
                for (uint32_t it = 0; it < maxh; ++it) {
                    if (m) {
                        const uint32_t b = (uint32_t)__ffsll((long long)m) - 1u;
                        m &= m - 1ull;
                        const uint32_t row = (b * inv) >> 16, col = b - row * (uint32_t)w;
                        if (idx < a.cap)  // converterFS.glsl:48-51 beyond the cap
                            a.frag_ids[idx] = make_uint2(tg, ((uint32_t)(by + (int)row) << 12) | (uint32_t)(bx + (int)col));
                        ++idx;
                    }
                }
            }
        }
        STAMP(a, 5);
        // medium triangles: the whole warp covers one triangle at a time
        {
            unsigned mm = __ballot_sync(0xffffffffu, cnt != 0 && !small);
            while (mm) {
                const int s = __ffs(mm) - 1;
                mm &= mm - 1;
                const uint32_t cs = __shfl_sync(0xffffffffu, cnt, s);
                raster_one<LAYOUT>(a, wb, qn, (uint32_t)s, 0u, cs, lane, tr);
            }
        }
        flush_ids<LAYOUT>(a, wb, qn, lane);
        qn = 0;
        __syncwarp();
        unit = __shfl_sync(0xffffffffu, next, 0);
        STAMP(a, 6);
    }
    STAMP(a, 7);

    // ---- drain: deferred big triangles, chunk by chunk, all warps ------------------------------
    if (lane == 0) {
        unsigned ns = 100;
        while (ld_acquire_u32(SCHED(a, 1)) < a.n_units) { __nanosleep(ns); ns = min(ns * 2u, 1000u); }
    }
    __syncwarp();
    STAMP(a, 8);
    uint32_t tail = 0;
    if (lane == 0) tail = ld_acquire_u32(SCHED(a, 2));
    tail = min(__shfl_sync(0xffffffffu, tail, 0), a.queue_cap);
    unsigned long long n_items = 0, t_setup = 0, t_rast = 0, t_load = 0;
    while (tail) {
        const unsigned long long ta = TNOW();
        uint32_t it = 0;
        if (lane == 0) it = atomicAdd(SCHED(a, 3), 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= tail) break;
        const uint2 item = a.queue[it];
        if (item.x == 0xffffffffu) continue;  // slot of a push that did not fit
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        if (lane < 9) wb.tri[lane] = a.tris[(size_t)item.x * 9 + lane];
        __syncwarp();
        const unsigned long long tb = TNOW();
        uint32_t c = 0;
        TriRaster tr;
        tr.w = 1; tr.h = 0; tr.x0 = 0; tr.y0 = 0; tr.incl = 0;
        tr.A[0] = tr.A[1] = tr.A[2] = tr.B[0] = tr.B[1] = tr.B[2] = 0; tr.C[0] = tr.C[1] = tr.C[2] = 0;
        if (lane == 0) c = M2S_SETUP_DRAIN<LAYOUT>(wb.tri, item.x, a, tabs, tr, wb.frag[0]);
        c = __shfl_sync(0xffffffffu, c, 0);
        __syncwarp();
        const unsigned long long tc = TNOW();
        const uint32_t c0 = item.y * kChunkCand, c1 = min(c, c0 + kChunkCand);
        raster_one<LAYOUT>(a, wb, qn, 0u, c0, c1, lane, tr);
        flush_ids<LAYOUT>(a, wb, qn, lane);
        qn = 0;
        __syncwarp();
        const unsigned long long td = TNOW();
        ++n_items; t_load += tb - ta; t_setup += tc - tb; t_rast += td - tc;
    }
    STAMPV(a, 12, n_items); STAMPV(a, 13, t_load); STAMPV(a, 14, t_setup); STAMPV(a, 15, t_rast);

    STAMP(a, 9);
    if (lane == 0) tma_store_wait_all();  // record stores are complete (not just read) before the kernel ends
    // ---- last CTA out publishes the count and re-arms the scheduler for the next launch ---------
    STAMP(a, 10);
    __syncthreads();
    STAMP(a, 11);
    if (warp == 0) {
        uint32_t last = 0;
        unsigned long long tot = 0;
        if (lane == 0) {
            __threadfence();
            const uint32_t done = atomicAdd(SCHED(a, 4), 1u);
            if (done == gridDim.x - 1) {
                __threadfence();
                last = 1;
                tot = *reinterpret_cast<volatile unsigned long long*>(a.counter);
                *a.total_out = tot;
                *a.counter = 0ull;
                if (a.host_total) {  // zero-copy count for the host (PCIe posted write, ~1 us)
                    *reinterpret_cast<volatile unsigned long long*>(a.host_total) = tot;
                    __threadfence_system();
                    *reinterpret_cast<volatile unsigned long long*>(a.host_total + 1) = a.host_tag;
                }
                *SCHED(a, 0) = 0; *SCHED(a, 1) = 0; *SCHED(a, 2) = 0; *SCHED(a, 3) = 0; *SCHED(a, 4) = 0;
                __threadfence();
            }
        }
        if (a.world > 1) {  // fused gather: tell every peer how many records this rank will write, one lane per peer
            last = __shfl_sync(0xffffffffu, last, 0);
            tot = __shfl_sync(0xffffffffu, tot, 0);
            if (last && (uint32_t)lane < a.world) {
                const unsigned long long mine = tot < a.cap ? tot : a.cap;
                a.peer_xch[lane][a.rank * 4 + 0] = mine;
                __threadfence_system();
                st_release_sys(a.peer_xch[lane] + a.rank * 4 + 1, a.epoch);
            }
        }
    }
}

// A warp's staged records (shared memory, 16-byte aligned) -> one contiguous span of global memory at byte
// offset `boff` of `dstbase`.  16-byte stores when the span starts 16-byte aligned (REF96 always; PACKED56
// at even record offsets), else 8-byte stores (appended chunks / gathered ranks may start at an odd record).
template <int STRIDE>
__device__ __forceinline__ void copy_span(uint8_t* dstbase, unsigned long long boff, const unsigned char* stage, uint32_t nbytes, int lane) {
    if ((boff & 15ull) == 0) {
        float4* dst = reinterpret_cast<float4*>(dstbase + boff);
        const float4* src = reinterpret_cast<const float4*>(stage);
        const uint32_t n16 = nbytes / 16;
#pragma unroll
        for (int j = 0; j < (32 * STRIDE / 16 + 31) / 32; ++j) {
            const uint32_t c = lane + 32 * j;
            if (c < n16) dst[c] = src[c];
        }
        if ((nbytes & 8u) && lane == 0)  // odd number of 56-byte records: one trailing 8-byte piece
            reinterpret_cast<float2*>(dst)[n16 * 2] = reinterpret_cast<const float2*>(src)[n16 * 2];
    } else {
        float2* dst = reinterpret_cast<float2*>(dstbase + boff);
        const float2* src = reinterpret_cast<const float2*>(stage);
        const uint32_t n8 = nbytes / 8;
#pragma unroll
        for (int j = 0; j < 32 * STRIDE / 8 / 32; ++j) {
            const uint32_t c = lane + 32 * j;
            if (c < n8) dst[c] = src[c];
        }
    }
}

// ------------------------------------------------------------------------------------------
// fragment stage: one warp = 32 consecutive fragments = 32 consecutive output records
// ------------------------------------------------------------------------------------------
template <int LAYOUT>
__global__ void __launch_bounds__(M2S_FRAG_THREADS) fragment_kernel(const __grid_constant__ ConvertArgs a) {
    using C = Cfg<LAYOUT>;
    constexpr int kStride = C::kStride;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* stage = smem + (size_t)warp * 32 * kStride;  // this warp's 32 records
    // launched with programmatic stream serialisation: the CTAs of this grid are placed while the raster
    // kernel drains; everything it wrote is visible after this wait
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const unsigned long long total = *reinterpret_cast<const volatile unsigned long long*>(a.total_out);
    // appended launches (m2s_convert_host pipelines a scene in triangle chunks): this launch's records follow
    // those of the earlier chunks; the cap applies to the running index, as the reference's counter does
    unsigned long long base = 0;
    for (uint32_t j = 0; j < a.nprev; ++j) base += *reinterpret_cast<const volatile unsigned long long*>(a.prev_totals + j);
    const unsigned long long room = a.cap > base ? a.cap - base : 0ull;
    const unsigned long long n = total < room ? total : room;
    // fused gather: wait for every rank's count of this epoch, my records start after the lower ranks'
    __shared__ unsigned long long s_goff;
    unsigned long long goff = 0;
    if (a.world > 1) {
        if (threadIdx.x == 0) {
            const unsigned long long* x = a.peer_xch[a.rank];
            unsigned long long off = 0;
            for (uint32_t r = 0; r < a.world; ++r) {
                while (ld_acquire_sys(x + r * 4 + 1) != a.epoch) __nanosleep(100);
                if (r < a.rank) off += *reinterpret_cast<const volatile unsigned long long*>(x + r * 4);
            }
            s_goff = off;
        }
        __syncthreads();
        goff = s_goff;
    }
    const uint32_t* __restrict__ texb = a.tex_base;
    const bool want_keys = a.keys != nullptr;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    for (unsigned long long g = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + warp; g * 32 < n; g += nwarps) {
        const unsigned long long wbase = g * 32;
        const uint32_t nfr = (uint32_t)min(32ull, n - wbase);
        unsigned long long key = 0;
        if ((uint32_t)lane < nfr) {
            const uint2 fid = __ldg(a.frag_ids + wbase + lane);
            const uint32_t tl = fid.x - a.tri_first;
            const int py = (fid.y >> 12) & 0xfff, px = fid.y & 0xfff;
#ifdef M2S_FRAG_PREFETCH
            // Tuning build (off): the stall samples of this kernel sit on four serial load waits — fragment id ->
            // triangle record -> vertices -> texels (profiles/r01_fragment_kernel_*.md: 14.8 + 13.2 + 16.6 + 10.3 % of
            // the samples).  Pull the position vertices (read last, after the texel loads were issued) into L1 as soon
            // as the triangle is known, and the next iteration's fragment ids while this one is shaded.
            {
                const float4* pv = a.tris + (size_t)fid.x * 9;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pv));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pv + 3));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pv + 6));
                if ((g + nwarps) * 32 < n) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.frag_ids + (g + nwarps) * 32 + lane));
            }
#endif
            const TriFragT<C::kMaps> tf = *reinterpret_cast<const TriFragT<C::kMaps>*>(a.tri_frag + (size_t)tl * sizeof(TriFragT<C::kMaps>));
            const unsigned meta = tf.meta;
            const int dxi = px - (int)((meta >> 4) & 0xfffu), dyi = py - (int)((meta >> 16) & 0xfffu);
            const float l0 = __ll2float_rn(tf.E0[0] + (long long)tf.A[0] * dxi + (long long)tf.B[0] * dyi) * tf.inv_area;
            const float l1 = __ll2float_rn(tf.E0[1] + (long long)tf.A[1] * dxi + (long long)tf.B[1] * dyi) * tf.inv_area;
            const float l2 = __ll2float_rn(tf.E0[2] + (long long)tf.A[2] * dxi + (long long)tf.B[2] * dyi) * tf.inv_area;
            const float4* __restrict__ v = a.tris + (size_t)fid.x * 9;  // 3 x {pos3 nrm3 tan4 uv2}, L2-resident
            // uv first: the texel addresses depend on nothing else
            const float4 a2 = __ldg(v + 2), b2 = __ldg(v + 5), c2 = __ldg(v + 8);
            const float u = l0 * a2.z + l1 * b2.z + l2 * c2.z, vv = l0 * a2.w + l1 * b2.w + l2 * c2.w;

            // ---- issue every texel load of every bound map back to back ----
            uint32_t tx[C::kMaps][8];
            Bilin bl[C::kMaps][2];
            bool has[C::kMaps], two[C::kMaps];
            uint32_t offs0[C::kMaps], offs1[C::kMaps];
#pragma unroll
            for (int m = 0; m < C::kMaps; ++m) {
                const TexRef ref = tf.tex[m];
                has[m] = ref.off0 != 0xffffffffu;
                two[m] = has[m] && tf.frac[m] > 0.0f;
                offs0[m] = has[m] ? ref.off0 : 0u;
                offs1[m] = ref.off1;
                if (m == 0 || !((meta >> m) & 1u)) {
                    bl[m][0] = bilin_setup(ref.w0, ref.h0, u, vv);
                    bl[m][1] = bilin_setup(ref.w1, ref.h1, u, vv);
                } else { bl[m][0] = bl[0][0]; bl[m][1] = bl[0][1]; }
            }
#pragma unroll
            for (int m = 0; m < C::kMaps; ++m) {
                const uint32_t o0 = offs0[m], o1 = offs1[m];  // uniform base + 32-bit texel index
                tx[m][0] = has[m] ? __ldg(texb + (o0 + bl[m][0].i00)) : 0u; tx[m][1] = has[m] ? __ldg(texb + (o0 + bl[m][0].i10)) : 0u;
                tx[m][2] = has[m] ? __ldg(texb + (o0 + bl[m][0].i01)) : 0u; tx[m][3] = has[m] ? __ldg(texb + (o0 + bl[m][0].i11)) : 0u;
                tx[m][4] = two[m] ? __ldg(texb + (o1 + bl[m][1].i00)) : 0u; tx[m][5] = two[m] ? __ldg(texb + (o1 + bl[m][1].i10)) : 0u;
                tx[m][6] = two[m] ? __ldg(texb + (o1 + bl[m][1].i01)) : 0u; tx[m][7] = two[m] ? __ldg(texb + (o1 + bl[m][1].i11)) : 0u;
            }
            // ---- interpolate the remaining varyings while the loads are in flight ----
            const float4 a0 = __ldg(v + 0), b0 = __ldg(v + 3), c0 = __ldg(v + 6);
            const float Px = l0 * a0.x + l1 * b0.x + l2 * c0.x, Py = l0 * a0.y + l1 * b0.y + l2 * c0.y,
                        Pz = l0 * a0.z + l1 * b0.z + l2 * c0.z;
            float* srec = reinterpret_cast<float*>(stage + lane * kStride);

            // colour (converterFS.glsl:55-62,99)
            float cr = 1.f, cg = 1.f, cb = 1.f, ca = 1.f;
            if (has[0]) {
                const float f = tf.frac[0];
                cr = filt<0>(bl[0][0], tx[0][0], tx[0][1], tx[0][2], tx[0][3]);
                cg = filt<1>(bl[0][0], tx[0][0], tx[0][1], tx[0][2], tx[0][3]);
                cb = filt<2>(bl[0][0], tx[0][0], tx[0][1], tx[0][2], tx[0][3]);
                ca = filt<3>(bl[0][0], tx[0][0], tx[0][1], tx[0][2], tx[0][3]);
                if (two[0]) {
                    cr += f * (filt<0>(bl[0][1], tx[0][4], tx[0][5], tx[0][6], tx[0][7]) - cr);
                    cg += f * (filt<1>(bl[0][1], tx[0][4], tx[0][5], tx[0][6], tx[0][7]) - cg);
                    cb += f * (filt<2>(bl[0][1], tx[0][4], tx[0][5], tx[0][6], tx[0][7]) - cb);
                    ca += f * (filt<3>(bl[0][1], tx[0][4], tx[0][5], tx[0][6], tx[0][7]) - ca);
                }
            }
            cr *= tf.factor[0]; cg *= tf.factor[1]; cb *= tf.factor[2]; ca *= tf.factor[3];

            if (LAYOUT == 0) {
                const float Nx = l0 * a0.w + l1 * b0.w + l2 * c0.w;
                const float4 a1 = __ldg(v + 1), b1 = __ldg(v + 4), c1 = __ldg(v + 7);
                const float Ny = l0 * a1.x + l1 * b1.x + l2 * c1.x, Nz = l0 * a1.y + l1 * b1.y + l2 * c1.y;
                float nx = Nx, ny = Ny, nz = Nz;
                constexpr int MN = C::kMaps > 1 ? 1 : 0, MM = C::kMaps > 2 ? 2 : 0;
                if (has[MN]) {  // :64-77 TBN
                    const float f = tf.frac[MN];
                    float mx = filt<0>(bl[MN][0], tx[MN][0], tx[MN][1], tx[MN][2], tx[MN][3]);
                    float my = filt<1>(bl[MN][0], tx[MN][0], tx[MN][1], tx[MN][2], tx[MN][3]);
                    float mz = filt<2>(bl[MN][0], tx[MN][0], tx[MN][1], tx[MN][2], tx[MN][3]);
                    if (two[MN]) {
                        mx += f * (filt<0>(bl[MN][1], tx[MN][4], tx[MN][5], tx[MN][6], tx[MN][7]) - mx);
                        my += f * (filt<1>(bl[MN][1], tx[MN][4], tx[MN][5], tx[MN][6], tx[MN][7]) - my);
                        mz += f * (filt<2>(bl[MN][1], tx[MN][4], tx[MN][5], tx[MN][6], tx[MN][7]) - mz);
                    }
                    const float Tx = l0 * a1.z + l1 * b1.z + l2 * c1.z, Ty = l0 * a1.w + l1 * b1.w + l2 * c1.w;
                    const float Tz = l0 * a2.x + l1 * b2.x + l2 * c2.x, Tw = l0 * a2.y + l1 * b2.y + l2 * c2.y;
                    float rx = mx * 2.0f - 1.0f, ry = my * 2.0f - 1.0f, rz = mz * 2.0f - 1.0f;
                    float inv = rsqrtf(rx * rx + ry * ry + rz * rz);
                    rx *= inv; ry *= inv; rz *= inv;
                    float bx = Ny * Tz - Ty * Nz, by = Nz * Tx - Tz * Nx, bz = Nx * Ty - Tx * Ny;  // cross(N,T)
                    inv = Tw * rsqrtf(bx * bx + by * by + bz * bz);
                    bx *= inv; by *= inv; bz *= inv;
                    inv = rsqrtf(Nx * Nx + Ny * Ny + Nz * Nz);
                    const float ox = Tx * rx + bx * ry + Nx * inv * rz, oy = Ty * rx + by * ry + Ny * inv * rz,
                                oz = Tz * rx + bz * ry + Nz * inv * rz;
                    inv = rsqrtf(ox * ox + oy * oy + oz * oz);
                    nx = ox * inv; ny = oy * inv; nz = oz * inv;
                }
                float metal = 0.1f, rough = 0.5f;  // :83-95 (.bg)
                if (has[MM]) {
                    const float f = tf.frac[MM];
                    rough = filt<1>(bl[MM][0], tx[MM][0], tx[MM][1], tx[MM][2], tx[MM][3]);
                    metal = filt<2>(bl[MM][0], tx[MM][0], tx[MM][1], tx[MM][2], tx[MM][3]);
                    if (two[MM]) {
                        rough += f * (filt<1>(bl[MM][1], tx[MM][4], tx[MM][5], tx[MM][6], tx[MM][7]) - rough);
                        metal += f * (filt<2>(bl[MM][1], tx[MM][4], tx[MM][5], tx[MM][6], tx[MM][7]) - metal);
                    }
                }
                float4* s4 = reinterpret_cast<float4*>(srec);
                s4[0] = make_float4(Px, Py, Pz, 1.0f);
                s4[1] = make_float4(cr, cg, cb, ca);
                s4[2] = make_float4(tf.scale[0], tf.scale[1], tf.scale[2], 0.0f);
                s4[3] = make_float4(nx, ny, nz, 0.0f);
                s4[4] = make_float4(tf.quat[0], tf.quat[1], tf.quat[2], tf.quat[3]);
                s4[5] = make_float4(metal, rough, 0.0f, 1.0f);
            } else {
                // parsers.cpp:484-499: SH0, opacity logit, log scale (per triangle)
                float2* s2 = reinterpret_cast<float2*>(srec);
                s2[0] = make_float2(Px, Py);
                s2[1] = make_float2(Pz, tf.quat[0]);
                s2[2] = make_float2(tf.quat[1], tf.quat[2]);
                s2[3] = make_float2(tf.quat[3], tf.scale[0]);
                s2[4] = make_float2(tf.scale[1], tf.scale[2]);
                const float kC0 = 0.28209479177387814f;  // SH_COEFF0, params.hpp:17
                s2[5] = make_float2(__fdiv_rn(cr - 0.5f, kC0), __fdiv_rn(cg - 0.5f, kC0));
                s2[6] = make_float2(__fdiv_rn(cb - 0.5f, kC0), inv_sigmoid(ca));
            }
            if (want_keys) key = ((unsigned long long)fid.x << 24) | (unsigned long long)fid.y;
        }
        __syncwarp();
        // ---- the warp's records are one contiguous span: straight vector copy -----------------------
        if (a.world <= 1) {
            copy_span<kStride>(a.out, (base + wbase) * (unsigned long long)kStride, stage, nfr * kStride, lane);
        } else {
            // fused gather: the same span goes to the final buffer of EVERY rank (peer stores over NVLink)
            const unsigned long long gbase = goff + wbase;
            uint32_t nval = 0;
            if (gbase < a.gcap) nval = (uint32_t)min((unsigned long long)nfr, a.gcap - gbase);
            // destinations are visited in a rotated order (by rank and by span) so that at any moment the
            // grid's stores are spread over all peers' ingress ports instead of converging on peer 0
            uint32_t p = (a.rank + 1u + (uint32_t)g) % a.world;
            for (uint32_t i = 0; i < a.world; ++i) {
                copy_span<kStride>(a.peer_out[p], gbase * (unsigned long long)kStride, stage, nval * kStride, lane);
                p = p + 1 == a.world ? 0 : p + 1;
            }
        }
        if (want_keys && (uint32_t)lane < nfr) a.keys[base + wbase + lane] = key;
        __syncwarp();
    }
    if (a.world > 1) {  // last CTA out tells every peer that this rank's records have landed
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t done = atomicAdd(SCHED(a, 6), 1u);
            if (done == gridDim.x - 1) {
                *SCHED(a, 6) = 0;
                __threadfence_system();
                for (uint32_t i = 0; i < a.world; ++i) {
                    const uint32_t p = (a.rank + 1u + i) % a.world;
                    st_release_sys(a.peer_xch[p] + a.rank * 4 + 2, a.epoch);
                }
            }
        }
    }
}

// waits until every rank's records of this epoch have landed in THIS rank's final buffer, publishes the
// total; one thread — the data path never returns to the host
__global__ void gather_wait_kernel(const unsigned long long* xch, uint32_t world, unsigned long long epoch,
                                   unsigned long long gcap, unsigned long long* total_global) {
    unsigned long long tot = 0;
    for (uint32_t r = 0; r < world; ++r) {
        while (ld_acquire_sys(xch + r * 4 + 2) != epoch) __nanosleep(200);
        tot += *reinterpret_cast<const volatile unsigned long long*>(xch + r * 4);
    }
    if (total_global) *total_global = tot;
    (void)gcap;
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
size_t raster_smem_bytes(int layout) {
    return (layout == 0 ? sizeof(WarpBlock<0>) * Cfg<0>::kWarps : sizeof(WarpBlock<1>) * Cfg<1>::kWarps) + kTableSmemBytes;
}
size_t fragment_smem_bytes(int layout) { return (size_t)(M2S_FRAG_THREADS / 32) * 32 * (layout == 0 ? Cfg<0>::kStride : Cfg<1>::kStride); }
int convert_warps_per_cta(int layout) { return layout == 0 ? Cfg<0>::kWarps : Cfg<1>::kWarps; }
size_t tri_frag_bytes(int layout) { return layout == 0 ? sizeof(TriFragT<Cfg<0>::kMaps>) : sizeof(TriFragT<Cfg<1>::kMaps>); }

cudaError_t convert_configure(int layout, int* raster_blocks_per_sm, int* fragment_blocks_per_sm) {
    cudaError_t e;
    const size_t smem = raster_smem_bytes(layout), fsmem = fragment_smem_bytes(layout);
    if (layout == 0) {
        e = cudaFuncSetAttribute(raster_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(raster_blocks_per_sm, raster_kernel<0>, Cfg<0>::kWarps * 32, smem);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(fragment_blocks_per_sm, fragment_kernel<0>, M2S_FRAG_THREADS, fsmem);
    }
    e = cudaFuncSetAttribute(raster_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(raster_blocks_per_sm, raster_kernel<1>, Cfg<1>::kWarps * 32, smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(fragment_blocks_per_sm, fragment_kernel<1>, M2S_FRAG_THREADS, fsmem);
}

cudaError_t convert_launch(int layout, const ConvertArgs& args, int raster_grid, int fragment_grid, cudaStream_t stream) {
    const size_t smem = raster_smem_bytes(layout), fsmem = fragment_smem_bytes(layout);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)fragment_grid);
    cfg.blockDim = dim3(M2S_FRAG_THREADS);
    cfg.dynamicSmemBytes = fsmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // PDL: overlap this launch with the raster kernel's tail
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
#ifndef M2S_NO_PDL
    cfg.numAttrs = 1;
#else
    cfg.numAttrs = 0;
#endif
    if (layout == 0) {
        raster_kernel<0><<<raster_grid, Cfg<0>::kWarps * 32, smem, stream>>>(args);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        return cudaLaunchKernelEx(&cfg, fragment_kernel<0>, args);
    }
    raster_kernel<1><<<raster_grid, Cfg<1>::kWarps * 32, smem, stream>>>(args);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, fragment_kernel<1>, args);
}

cudaError_t gather_wait_launch(const unsigned long long* xch, uint32_t world, unsigned long long epoch, unsigned long long gcap,
                               unsigned long long* total_global, cudaStream_t stream) {
    gather_wait_kernel<<<1, 1, 0, stream>>>(xch, world, epoch, gcap, total_global);
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
