// TEST INFRASTRUCTURE.  C entry points around the REFERENCE's own .glb parser: SceneManager::parseGltfFile
// (src/utils/SceneManager.cpp:195-459; tinygltf + stb_image from the reference's thirdParty/).  oracle/build.py
// compiles SceneManager.cpp, utils.cpp and parsers.cpp where they lie under /root/reference and links them with
// this file into oracle/_ref/libm2s_refloader.so.  Nothing of the GL side runs: the GLEW entry points the
// translation unit references are null data symbols (oracle/_ref/glew_null.c, generated), and the two functions
// of other reference files that only the GL paths call are defined here as traps.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <vector>

// the one translation unit that instantiates tinygltf and the stb libraries (from the reference's thirdParty/);
// the macros are dropped again before the reference's own headers re-include stb_image.h
#define TINYGLTF_IMPLEMENTATION
#define STB_IMAGE_IMPLEMENTATION
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "tiny_gltf.h"
#undef TINYGLTF_IMPLEMENTATION
#undef STB_IMAGE_IMPLEMENTATION
#undef STB_IMAGE_WRITE_IMPLEMENTATION
#define STB_IMAGE_RESIZE_IMPLEMENTATION
#include "stb_image_resize.h"
#undef STB_IMAGE_RESIZE_IMPLEMENTATION

#include "utils/utils.hpp"
#include "utils/glUtils.hpp"
#include "utils/normalizedUvUnwrapping.hpp"
#define private public  // parseGltfFile is a private member; the reference's sources are not modified
#include "utils/SceneManager.hpp"
#undef private

// only reached through loadModel's GL half, which this harness never calls
void uvUnwrapping::generateNormalizedUvCoordinatesPerMesh(int&, int&, std::vector<utils::Mesh>&) { std::abort(); }
void glUtils::generateTextures(std::map<std::string, std::map<std::string, utils::TextureDataGl>>&) { std::abort(); }

struct RefScene { std::vector<utils::Mesh> meshes; bool ok = false; };

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API void* ref_glb_parse(const char* path) {
    auto* s = new RefScene();
    // parseGltfFile never touches the render context (SceneManager.cpp:195-459): the manager is bound to raw zeroed
    // storage, so no RenderContext (GL objects, ShaderRegistry) is ever constructed or destroyed; the manager
    // itself is leaked because its destructor runs the GL clean-up
    alignas(RenderContext) static unsigned char rc_storage[sizeof(RenderContext)];
    SceneManager* sm = new SceneManager(*reinterpret_cast<RenderContext*>(rc_storage));
    s->ok = sm->parseGltfFile(std::string(path), std::string(""), s->meshes);
    return s;
}
REF_API int ref_glb_ok(void* h) { return static_cast<RefScene*>(h)->ok ? 1 : 0; }
REF_API int ref_glb_mesh_count(void* h) { return (int)static_cast<RefScene*>(h)->meshes.size(); }
REF_API const char* ref_glb_mesh_name(void* h, int i) { return static_cast<RefScene*>(h)->meshes[i].name.c_str(); }
REF_API int ref_glb_face_count(void* h, int i) { return (int)static_cast<RefScene*>(h)->meshes[i].faces.size(); }
// faces as 3 x {pos3 nrm3 tan4 uv2} floats (the layout of m2s_scene.triangles)
REF_API void ref_glb_faces(void* h, int i, float* out) {
    for (const utils::Face& f : static_cast<RefScene*>(h)->meshes[i].faces)
        for (int k = 0; k < 3; ++k) {
            *out++ = f.pos[k].x; *out++ = f.pos[k].y; *out++ = f.pos[k].z;
            *out++ = f.normal[k].x; *out++ = f.normal[k].y; *out++ = f.normal[k].z;
            *out++ = f.tangent[k].x; *out++ = f.tangent[k].y; *out++ = f.tangent[k].z; *out++ = f.tangent[k].w;
            *out++ = f.uv[k].x; *out++ = f.uv[k].y;
        }
}
REF_API void ref_glb_base_color(void* h, int i, float* out4) {
    const glm::vec4 c = static_cast<RefScene*>(h)->meshes[i].material.baseColorFactor;
    out4[0] = c.x; out4[1] = c.y; out4[2] = c.z; out4[3] = c.w;
}
// which: 0 base colour, 1 normal, 2 metallic-roughness.  Returns the byte count of the decoded image (0: none)
static const utils::TextureInfo& tex_of(void* h, int i, int which) {
    const utils::MaterialGltf& m = static_cast<RefScene*>(h)->meshes[i].material;
    return which == 0 ? m.baseColorTexture : (which == 1 ? m.normalTexture : m.metallicRoughnessTexture);
}
REF_API uint64_t ref_glb_texture(void* h, int i, int which, int* w, int* hgt, int* channels, const unsigned char** data) {
    const utils::TextureInfo& t = tex_of(h, i, which);
    *w = t.width; *hgt = t.height; *channels = (int)t.channels; *data = t.texture.data();
    return (uint64_t)t.texture.size();
}
REF_API void ref_glb_free(void* h) { delete static_cast<RefScene*>(h); }
