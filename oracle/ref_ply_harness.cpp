// TEST INFRASTRUCTURE.  C entry point around the REFERENCE's own .ply writer: parsers::savePlyVector
// (src/parsers/parsers.cpp:631-651 -> writeBinaryPlyStandardFormat / writePbrPLY / writeCompressedPbrPLY) with
// utils::getShFromColor / invSigmoid (src/utils/utils.cpp:45-55, utils.hpp:269-270).  oracle/build.py compiles
// the reference's parsers.cpp and utils.cpp where they lie under /root/reference (stub <windows.h> / <crtdbg.h>
// generated into oracle/_ref/stubs) and links them with this file into oracle/_ref/libm2s_refply.so.
// The stb implementations the reference's translation units expect elsewhere are instantiated here from the
// reference's own vendored headers.
#define STB_IMAGE_IMPLEMENTATION
#define STB_IMAGE_RESIZE_IMPLEMENTATION
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "parsers/parsers.hpp"

extern "C" __attribute__((visibility("default"))) int ref_save_ply(const char* path, const float* rec96, uint64_t count,
                                                                    unsigned format, float scale_multiplier) {
    static_assert(sizeof(utils::GaussianDataSSBO) == 96, "GaussianDataSSBO is the 96-byte SSBO record");
    std::vector<utils::GaussianDataSSBO> v((size_t)count);
    if (count) std::memcpy(v.data(), rec96, (size_t)count * 96);
    parsers::savePlyVector(std::string(path), v, format, scale_multiplier);
    return 0;
}
