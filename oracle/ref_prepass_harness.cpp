// ref_prepass_harness.cpp — hosts the REFERENCE's viewer prepass compute shader on the CPU (TEST INFRASTRUCTURE).
//
// oracle/build.py rewrites /root/reference/src/shaders/rendering/gaussianSplattingPrepassCS.glsl (+ common.glsl) into
// oracle/_ref/prepassCS.inc (qualifier / literal / swizzle token rewrites only) and this file gives it the GLSL
// environment it expects: GLM types (the reference's vendored copy), the SSBOs, the atomic counter, the invocation id.
// Nothing here restates the algorithm — the arithmetic executed is the reference's own source.  Invocations run one
// after the other in gid order, so the survivors come out in input order.
#define GLM_FORCE_SWIZZLE   // .xyz() / .xy() swizzle functions
#include <cmath>
#include <cstdint>
#include <cstring>
#include <glm/glm.hpp>

#define REF_API extern "C" __attribute__((visibility("default")))

namespace refpp {
using namespace glm;
typedef unsigned int atomic_uint;
typedef unsigned int uint;
struct sampler2D { int unit; };
static vec4 texture(const sampler2D&, const vec2&) { return vec4(1.0f); }   // the mesh depth texture: only read when u_depthTestMesh == 1
static uint atomicCounterIncrement(atomic_uint& c) { return c++; }
static uvec3 gl_NumWorkGroups, gl_WorkGroupSize(16, 16, 1), gl_GlobalInvocationID;
#include "prepassCS.inc"
}  // namespace refpp

// gaussians: n x 24 floats (GaussianVertex); mats column-major (glm::mat4); quads: capacity n x 24 floats
// (QuadNdcTransformation); depths: capacity n.  Dispatch geometry as GaussiansPrepass.cpp:43-48.  Returns the counter.
REF_API uint32_t ref_prepass(const float* gaussians, uint32_t n, const float* world_to_view, const float* view_to_clip, const float* model_to_world,
                             const float* resolution, const float* near_far, float std_dev, int render_mode, uint32_t format, uint32_t ply_has_pbr,
                             float* quads, float* depths) {
    using namespace refpp;
    std::memcpy(&u_worldToView, world_to_view, 64);
    std::memcpy(&u_viewToClip, view_to_clip, 64);
    std::memcpy(&u_modelToWorld, model_to_world, 64);
    u_resolution = glm::vec2(resolution[0], resolution[1]);
    u_nearFar = glm::vec2(near_far[0], near_far[1]);
    u_stdDev = std_dev; u_renderMode = render_mode; u_format = format; u_plyHasPbr = ply_has_pbr; u_depthTestMesh = 0;
    u_gaussianCount = (int)n;
    gaussianBuffer.gaussians = reinterpret_cast<GaussianVertex*>(const_cast<float*>(gaussians));
    perQuadTransformations.ndcTransformations = reinterpret_cast<QuadNdcTransformation*>(quads);
    gaussianDepthPostFiltering.depths_vs = depths;
    g_validCounter = 0;
    const unsigned groups_needed = (n + 255u) / 256u;
    const unsigned gx = (unsigned)std::ceil(std::sqrt((float)groups_needed));
    const unsigned gy = gx ? (unsigned)((groups_needed + gx - 1) / std::max(float(gx), 1.0f)) : 0u;
    gl_NumWorkGroups = glm::uvec3(gx, gy, 1);
    const unsigned width = gx * 16u;
    for (unsigned y = 0; y < gy * 16u; ++y)
        for (unsigned x = 0; x < width; ++x) {
            gl_GlobalInvocationID = glm::uvec3(x, y, 0);
            prepass_main();
        }
    return g_validCounter;
}
