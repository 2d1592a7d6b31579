// debug driver: m2s_glb_load under AddressSanitizer (stubs for the CUDA-side symbols m2s_glb.cpp links against)
#include <cstdio>
#include <cstring>
#include "../include/m2s.h"  // build from the repository root: see scripts/fuzz_loader_asan.py
struct m2s_ctx;
namespace m2s { m2s_status convert_scene_to_ply(m2s_ctx*, const m2s_scene*, const m2s_params*, const char*, m2s_result*) { return M2S_E_NOGPU; } }
extern "C" {
void m2s_params_default(m2s_params* p) { std::memset(p, 0, sizeof(*p)); }
uint64_t m2s_reference_capacity(uint32_t, uint32_t) { return 0; }
m2s_status m2s_compute_bboxes(const float*, m2s_primitive*, uint32_t, int) { return M2S_OK; }
}
int main(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
        m2s_hscene* hs = nullptr;
        m2s_status st = m2s_glb_load(argv[i], 1, &hs);
        std::printf("%s: status %d\n", argv[i], (int)st);
        if (hs) m2s_hscene_free(hs);
    }
    return 0;
}
