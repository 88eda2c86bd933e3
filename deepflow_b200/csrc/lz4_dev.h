/*
 * lz4_dev.h — the compressed-frame stage (lz4.cu): constants and the region table shared with api.cu.
 */
#pragma once
#include <cstdint>
#include <vector>

namespace dfgpu {

constexpr uint32_t LZ_CHUNK = 4096;            // uncompressed bytes per frame
constexpr uint32_t LZ_FRAME_HDR = 25;          // 16 B checksum slot + method + compressed size + uncompressed size
constexpr uint32_t LZ_SCRATCH_STRIDE = 4160;   // worst-case LZ4 block of a chunk (4096 + 4096/255 + 16), 16-byte multiple

struct LzRegion {
  uint64_t off;          // byte offset of the column buffer in the arena (256-byte aligned)
  uint64_t len;          // bytes
  uint32_t first_chunk;  // index of its first chunk / frame (filled by lz4_compress_regions)
  uint32_t stride;       // element size of a fixed-width buffer (match distance tried first), 0: none
};

}  // namespace dfgpu

struct dfgpu_ctx;
namespace dfgpu {
int lz4_compress_regions(::dfgpu_ctx* ctx, const uint8_t* arena, std::vector<LzRegion>& regs, const uint8_t** d_frames,
                         std::vector<uint64_t>& frame_pos);
void lz4_free_state(::dfgpu_ctx* ctx);
}  // namespace dfgpu
