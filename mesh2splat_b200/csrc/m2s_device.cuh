// m2s_device.cuh — device-side data layout shared by the kernels and the C-ABI host code.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace m2s {

constexpr int kMaxPeers = 8;             // GPUs of one NVSwitch domain
constexpr int kMaxLevels = 5;            // levels 0..4 (GL_TEXTURE_MAX_LEVEL 4, glUtils.cpp:313)
constexpr int kUnitTris = 32;            // max triangles per work unit (one lane per triangle)
constexpr int kQueue = 512;              // fragment ids compacted per warp before a flush
constexpr int kTriBytes = 144;           // 36 floats
constexpr uint32_t kSmallCand = 64;      // <= this many candidate pixels: lane-per-triangle raster
constexpr uint32_t kBigCand = 1024;      // > this many: deferred, split into chunks over all warps
constexpr uint32_t kChunkCand = 512;     // candidates per deferred work item
constexpr float kGuard = 8192.0f;        // window-coordinate guard band (|xw| beyond -> triangle dropped)

// RGBA8 mip chain of one texture inside the texture arena (one allocation for all textures).
// Levels are pitch-linear, tightly packed, row 0 first; off[] are TEXEL offsets from the arena base.
struct DTexture {
    uint32_t off[kMaxLevels];
    uint16_t w[kMaxLevels];
    uint16_t h[kMaxLevels];
    uint32_t nlevels;  // q + 1, q = min(4, floor(log2(max(w,h))))
    uint32_t pad;
};
static_assert(sizeof(DTexture) == 48, "DTexture layout");

// One glTF primitive: the uniforms ConversionPass::conversion uploads per draw call
// (ConversionPass.cpp:77-112).
struct DPrim {
    float bmin[3];
    float bmax[3];
    float factor[4];
    int tex[3];  // albedo, normal, metallic-roughness; -1 = has*Map == 0
    int pad;
};

// sorted, disjoint triangle ranges -> primitive
static_assert(sizeof(DPrim) == 56, "DPrim layout");

// sorted, disjoint triangle ranges -> primitive
struct DRange {
    uint32_t first, end, prim, pad;
};

struct ConvertArgs {
    const float4* tris;   // 9 float4 per triangle (3 x {pos3 nrm3 tan4 uv2})
    uint32_t tri_first;   // shard
    uint32_t tri_count;
    const DRange* ranges;
    uint32_t nranges;
    const DPrim* prims;
    uint32_t nprims;
    const DTexture* texs;
    const uint32_t* tex_base;  // texture arena
    uint32_t ntex;
    uint32_t R;
    uint32_t row_begin, row_end;  // pixel-row band [row_begin, row_end) of the R x R grid (whole grid: 0, R)
    float half_R;
    float mult;  // sigma / R (SceneManager.cpp:668)
    float log_sz;  // ln(1e-7 * mult): the constant third log-scale of the packed layout
    // intermediates between the two kernels (context-owned scratch, L2-resident at the sizes of interest);
    // the fragment kernel reads the vertices themselves from `tris`
    uint2* frag_ids;                   // {global triangle, y << 12 | x} per fragment; index = output record index
    unsigned char* tri_frag;           // one TriFragT per triangle of the shard
    uint8_t* out;
    unsigned long long cap;
    unsigned long long* keys;          // optional
    unsigned long long* counter;       // fragments generated (the reference's atomic counter); context-owned,
                                       // zero at launch, re-zeroed by the last CTA
    unsigned long long* total_out;     // receives the final count (last CTA out)
    const unsigned long long* prev_totals;  // appended launches: counts of the earlier chunks (records start after them)
    uint32_t nprev;
    unsigned long long* host_total;    // optional, mapped pinned host memory: {count, tag} written by the raster kernel's
    unsigned long long host_tag;       // last CTA so the host can size the download while the fragment kernel still runs
    // scheduling state (zero at launch, re-armed by the last CTA)
    uint32_t* sched;                   // 5 words, 128 B apart: unit counter, units past set-up, queue tail, queue head,
                                       // CTAs finished
    uint32_t unit_tris;                // triangles per work unit (<= 32), chosen by the host for balance
    uint32_t n_units;
    uint2* queue;                      // deferred big-triangle chunks: (triangle, chunk)
    uint32_t queue_cap;
    unsigned long long* trace;         // M2S_TRACE builds only: 16 globaltimer stamps per raster warp
    // multi-GPU fused gather (world <= 1: off).  peer_out[p] / peer_xch[p] are rank p's final buffer and
    // exchange block mapped into this process (NVLink peer memory); every rank's fragment kernel stores its
    // records into ALL final buffers at its global offset.  xch block: [kMaxPeers][4] u64 =
    // {count, count_epoch, done_epoch, pad} per source rank.
    uint32_t world, rank;
    uint8_t* peer_out[kMaxPeers];
    unsigned long long* peer_xch[kMaxPeers];
    unsigned long long epoch;
    unsigned long long gcap;           // capacity of the final buffers (records)
};

}  // namespace m2s
