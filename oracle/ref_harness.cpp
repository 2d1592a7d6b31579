// ref_harness.cpp — hosts the REFERENCE's conversion shaders on the CPU (TEST INFRASTRUCTURE).
//
// oracle/build.py rewrites /root/reference/src/shaders/conversion/converter{GS,FS}.glsl into
// oracle/_ref/converter{GS,FS}.inc (qualifier/literal token rewrites only) and this file gives
// them the GLSL environment they expect: GLM types (the reference's vendored copy), the
// geometry-shader emit interface, the fragment shader's SSBO / atomic counter / sampler.
// Nothing here restates the algorithm — the arithmetic executed is the reference's own source.
//
// Exports (C ABI, used by tests/ and tests/golden/make_golden.py only):
//   ref_gs   : one geometry-shader invocation  (converterGS.glsl main)
//   ref_fs   : one fragment-shader invocation  (converterFS.glsl main), texels supplied by caller
#include <cmath>
#include <cstdint>
#include <cstring>
#include <glm/glm.hpp>

#define REF_API extern "C" __attribute__((visibility("default")))

namespace refgs {
using namespace glm;
static vec4 gl_Position;
struct Emitted { vec3 Position, Scale, Normal; vec2 UV; vec4 Tangent, Quaternion, glpos; };
static Emitted g_emitted[3];
static int g_nemit = 0;
static void EmitVertex();
static void EndPrimitive() {}
#include "converterGS.inc"
static void EmitVertex() {
    if (g_nemit < 3) {
        Emitted& e = g_emitted[g_nemit++];
        e.Position = Position; e.Scale = Scale; e.Normal = Normal; e.UV = UV;
        e.Tangent = Tangent; e.Quaternion = Quaternion; e.glpos = gl_Position;
    }
}
}  // namespace refgs

namespace reffs {
using namespace glm;
struct sampler2D { int unit; };
typedef unsigned int atomic_uint;
typedef unsigned int uint;
static vec4 g_texel[5];
static vec4 texture(const sampler2D& s, const vec2&) { return g_texel[s.unit]; }
static uint atomicCounterIncrement(atomic_uint& c) { return c++; }
static vec2 swz_bg(const vec4& v) { return vec2(v.b, v.g); }
#include "converterFS.inc"
}  // namespace reffs

// tri: 3 x {pos3 nrm3 tan4 uv2}.  Outputs per emitted vertex k: glpos[k][4]; flat outputs
// scale[3], quat[4] (as written by the shader: (w,x,y,z)); pass-through varyings are not
// returned (they are copies of the inputs).  Returns number of emitted vertices.
REF_API int ref_gs(const float* tri, const float* bmin, const float* bmax, float* glpos /*3x4*/,
                   float* scale /*3*/, float* quat /*4*/) {
    using namespace refgs;
    for (int k = 0; k < 3; ++k) {
        const float* v = tri + 12 * k;
        gs_in[k].position = glm::vec3(v[0], v[1], v[2]);
        gs_in[k].normal = glm::vec3(v[3], v[4], v[5]);
        gs_in[k].tangent = glm::vec4(v[6], v[7], v[8], v[9]);
        gs_in[k].uv = glm::vec2(v[10], v[11]);
        gs_in[k].normalizedUv = glm::vec2(0.0f);
        gs_in[k].scale = glm::vec3(0.0f);
    }
    u_bboxMin = glm::vec3(bmin[0], bmin[1], bmin[2]);
    u_bboxMax = glm::vec3(bmax[0], bmax[1], bmax[2]);
    g_nemit = 0;
    gs_main();
    for (int k = 0; k < g_nemit; ++k)
        for (int c = 0; c < 4; ++c) glpos[4 * k + c] = g_emitted[k].glpos[c];
    // Scale is assigned once before the emit loop (converterGS.glsl:430); Quaternion per vertex
    for (int c = 0; c < 3; ++c) scale[c] = g_emitted[0].Scale[c];
    for (int c = 0; c < 4; ++c) quat[c] = g_emitted[0].Quaternion[c];
    return g_nemit;
}

// One fragment.  varyings: P3 N3 T4 UV2 Scale3 Quat4 (19 floats); texels: albedo/normal/mr rgba
// (what texture() returns for the bound maps); flags bit0/1/2 = has albedo/normal/mr;
// counter_start/max_gaussians drive the overflow discard.  rec receives the 24-float
// GaussianVertex if written.  Returns 1 if the record was written, 0 if discarded; *counter_out
// is the counter after the invocation.
REF_API int ref_fs(const float* varyings, const float* albedo, const float* nrm, const float* mr,
                   uint32_t flags, const float* factor, uint32_t counter_start, int max_gaussians,
                   float* rec, uint32_t* counter_out) {
    using namespace reffs;
    Position = glm::vec3(varyings[0], varyings[1], varyings[2]);
    Normal = glm::vec3(varyings[3], varyings[4], varyings[5]);
    Tangent = glm::vec4(varyings[6], varyings[7], varyings[8], varyings[9]);
    UV = glm::vec2(varyings[10], varyings[11]);
    Scale = glm::vec3(varyings[12], varyings[13], varyings[14]);
    Quaternion = glm::vec4(varyings[15], varyings[16], varyings[17], varyings[18]);
    albedoTexture.unit = 0; normalTexture.unit = 1; metallicRoughnessTexture.unit = 2;
    occlusionTexture.unit = 3; emissiveTexture.unit = 4;
    g_texel[0] = glm::vec4(albedo[0], albedo[1], albedo[2], albedo[3]);
    g_texel[1] = glm::vec4(nrm[0], nrm[1], nrm[2], nrm[3]);
    g_texel[2] = glm::vec4(mr[0], mr[1], mr[2], mr[3]);
    hasAlbedoMap = (flags & 1u) ? 1 : 0;
    hasNormalMap = (flags & 2u) ? 1 : 0;
    hasMetallicRoughnessMap = (flags & 4u) ? 1 : 0;
    u_materialFactor = glm::vec4(factor[0], factor[1], factor[2], factor[3]);
    u_maxGaussians = max_gaussians;
    g_validCounter = counter_start;
    static GaussianVertex slot;
    const float sentinel = -12345.0f;
    slot.position = glm::vec4(sentinel);
    // the shader indexes vertices[index]; point the base so that [counter_start] is `slot`
    gaussianBuffer.vertices = &slot - counter_start;
    fs_main();
    *counter_out = g_validCounter;
    if (slot.position.x == sentinel && slot.position.w == sentinel) return 0;
    std::memcpy(rec, &slot, sizeof(float) * 24);
    return 1;
}
