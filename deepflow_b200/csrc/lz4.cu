/*
 * lz4.cu — ClickHouse compressed frames on the device ("per-column LZ4 framing").
 *
 * What crosses PCIe device->host is the e2e bound of the path; with this stage a column buffer
 * leaves the GPU as the byte stream ClickHouse's native protocol carries for it when compression is on: a
 * sequence of frames
 *     [16 B checksum slot][1 B method][4 B compressed size incl. this 9 B header][4 B uncompressed size][payload]
 * (ClickHouse CompressedWriteBuffer / ch-go compress.Writer, module not in the reference tree: format restated
 * from the protocol, pinned by decoding every frame with an independent LZ4 block decoder in the tests).
 * method 0x82 = LZ4 block, 0x02 = NONE (stored) when LZ4 would not shrink the chunk.  The checksum is
 * CityHash128 of header + payload; it is left ZERO here — the host shim fills it while it writes the frame
 * (go-faster/city, the hash ch-go itself uses).
 *
 * One warp per 4 KB chunk of a column buffer, everything in shared memory:
 *   1. match finding, position-parallel: the 4096 positions are visited 32 at a time in order; every lane hashes
 *      the 4 bytes at its position, reads the candidate the table holds (an EARLIER tile's position), publishes
 *      its own, and also tries the fixed distance `stride` (the element size of the column: equal or
 *      nearly equal neighbours are what fixed-width columns are made of).  A candidate that agrees in its first
 *      4 bytes is kept (offset per position + a bit mask of the positions that have one).
 *   2. parse, segment-parallel: lane j walks positions [128 j, 128 j + 128) greedily (first match wins; only
 *      the matches it takes are extended, so every byte is compared at most once; jump behind the match),
 *      matches cut at the segment end.  Literal runs may span segments: the end of the last match
 *      of the lower lanes comes from a warp max-scan.
 *   3. emit: per-lane output sizes -> exclusive scan -> every lane writes its own sequences; the final literal
 *      run is copied by the whole warp.
 * The result is a valid LZ4 block (last 5 bytes literals, last match starts >= 12 bytes before the end).
 */
#include "ctx.h"
#include "lz4_dev.h"

namespace dfgpu {

namespace {

constexpr uint32_t HT_BITS = 11, HT_SIZE = 1u << HT_BITS;
constexpr uint32_t SEG = LZ_CHUNK / 32;  // bytes per lane in the parse
constexpr uint32_t CW = 2;               // warps per CTA (16.5 KB of shared memory each: 6 CTAs per SM)

struct alignas(16) WarpSmem {
  uint32_t in[LZ_CHUNK / 4 + 4];  // the chunk (+ slack so that 4-byte reads at the end stay inside)
  uint16_t tab[HT_SIZE];
  uint16_t moff[LZ_CHUNK + 1];  // candidate offset per position; [p + 1] of a TAKEN match at p holds its length
  uint16_t pad_;
  uint32_t mask[LZ_CHUNK / 32];  // bit set: a match of >= 4 bytes starts here
};

__device__ __forceinline__ uint32_t ld4(const uint32_t* in, uint32_t pos) {
  const uint32_t w = pos >> 2, s = (pos & 3u) * 8u;
  return __funnelshift_r(in[w], in[w + 1], s);
}
__device__ __forceinline__ uint32_t ld1(const uint32_t* in, uint32_t pos) { return (in[pos >> 2] >> ((pos & 3u) * 8u)) & 0xffu; }

// bytes of agreement between the data at a and at b (a > b), at most `cap`
__device__ __forceinline__ uint32_t match_len(const uint32_t* in, uint32_t a, uint32_t b, uint32_t cap) {
  uint32_t n = 0;
  while (n + 4 <= cap) {
    const uint32_t x = ld4(in, a + n) ^ ld4(in, b + n);
    if (x) return n + ((uint32_t)__ffs((int)x) - 1u) / 8u;
    n += 4;
  }
  while (n < cap && ld1(in, a + n) == ld1(in, b + n)) n++;
  return n;
}

__device__ __forceinline__ uint32_t seq_size(uint32_t lit, uint32_t ml) {
  uint32_t n = 1 + lit + 2;
  if (lit >= 15) n += 1 + (lit - 15) / 255;
  if (ml - 4 >= 15) n += 1 + (ml - 4 - 15) / 255;
  return n;
}

// writes one sequence; returns the advanced output pointer
__device__ __forceinline__ uint8_t* put_seq(uint8_t* o, const uint32_t* in, uint32_t lit_start, uint32_t lit, uint32_t off, uint32_t ml) {
  const uint32_t mcode = ml - 4;
  *o++ = (uint8_t)((lit >= 15 ? 15u : lit) << 4 | (mcode >= 15 ? 15u : mcode));
  if (lit >= 15) {
    uint32_t r = lit - 15;
    while (r >= 255) {
      *o++ = 255;
      r -= 255;
    }
    *o++ = (uint8_t)r;
  }
  for (uint32_t i = 0; i < lit; i++) *o++ = (uint8_t)ld1(in, lit_start + i);
  *o++ = (uint8_t)off;
  *o++ = (uint8_t)(off >> 8);
  if (mcode >= 15) {
    uint32_t r = mcode - 15;
    while (r >= 255) {
      *o++ = 255;
      r -= 255;
    }
    *o++ = (uint8_t)r;
  }
  return o;
}

}  // namespace

// chunk c of the batch: region by binary search over first_chunk
__device__ __forceinline__ uint32_t lz_region_of(const LzRegion* regs, uint32_t nregs, uint32_t c) {
  uint32_t lo = 0, hi = nregs;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (regs[mid].first_chunk <= c) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(CW * 32) k_lz4_compress(const uint8_t* __restrict__ arena, const LzRegion* __restrict__ regs,
                                                           uint32_t nregs, uint32_t nchunks, uint8_t* __restrict__ scratch,
                                                           uint32_t* __restrict__ csize) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  WarpSmem& S = reinterpret_cast<WarpSmem*>(smem_raw)[threadIdx.x >> 5];
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gwarp = blockIdx.x * CW + (threadIdx.x >> 5), total_warps = gridDim.x * CW;
  for (uint32_t c = gwarp; c < nchunks; c += total_warps) {
    const uint32_t r = lz_region_of(regs, nregs, c);
    const LzRegion rg = regs[r];
    const uint64_t coff = (uint64_t)(c - rg.first_chunk) * LZ_CHUNK;
    const uint32_t n = (uint32_t)(rg.len - coff < LZ_CHUNK ? rg.len - coff : LZ_CHUNK);
    const uint8_t* src = arena + rg.off + coff;  // 16-byte aligned (regions are 256-byte aligned)
    const uint32_t stride = rg.stride;
    // ---- load, reset ----
    {
      const uint4* s4 = reinterpret_cast<const uint4*>(src);
      uint4* d4 = reinterpret_cast<uint4*>(S.in);
      const uint32_t n16 = (n + 15) / 16;  // reads at most 15 bytes past n: inside the arena's 256-byte padding
      for (uint32_t i = lane; i < LZ_CHUNK / 16 + 1; i += 32) d4[i] = i < n16 ? __ldg(s4 + i) : make_uint4(0, 0, 0, 0);
      uint4* t4 = reinterpret_cast<uint4*>(S.tab);
      if (stride < 8)
        for (uint32_t i = lane; i < HT_SIZE * 2 / 16; i += 32) t4[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
    }
    __syncwarp();
    // matches start at <= n - 12 and end at <= n - 5 (LZ4 block format end conditions)
    const uint32_t mstart_lim = n >= 13 ? n - 12 : 0, mend = n >= 5 ? n - 5 : 0;
    // ---- 1. candidates ----
    // A buffer of 8- or 16-byte elements only tries the distance of one element — equal high bytes of neighbouring
    // values are its redundancy, and the test needs no table; byte streams (strings) and narrow elements (whose
    // repeats sit at any distance: enums, ports, flags) use the hash table.  Either way a candidate only has to
    // agree in its first 4 bytes here; the parse extends what it takes.
    const bool by_stride = stride >= 8;
    if (by_stride) {
      for (uint32_t t = 0; t < LZ_CHUNK / 32; t++) {
        const uint32_t pos = t * 32 + lane;
        const bool hit = n >= 13 && pos <= mstart_lim && pos >= stride && ld4(S.in, pos - stride) == ld4(S.in, pos);
        S.moff[pos] = (uint16_t)(hit ? stride : 0u);
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) S.mask[t] = m;
      }
    } else {
      for (uint32_t t = 0; t < LZ_CHUNK / 32; t++) {
        const uint32_t pos = t * 32 + lane;
        const bool valid = n >= 13 && pos <= mstart_lim;
        uint32_t v4 = 0, h = 0, cand = 0xffffu;
        if (valid) {
          v4 = ld4(S.in, pos);
          h = (v4 * 2654435761u) >> (32 - HT_BITS);
          cand = S.tab[h];
        }
        __syncwarp();
        if (valid) S.tab[h] = (uint16_t)pos;
        const bool hit = valid && cand != 0xffffu && ld4(S.in, cand) == v4;
        S.moff[pos] = (uint16_t)(hit ? pos - cand : 0u);
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) S.mask[t] = m;
        __syncwarp();
      }
    }
    __syncwarp();
    // ---- 2. parse: sizes ----
    const uint32_t seg0 = lane * SEG, seg1 = seg0 + SEG;
    uint32_t my_first = 0xffffffffu, my_last_end = 0, my_bytes = 0;  // my_bytes: sequences after my first match
    uint32_t first_ml = 0;
    {
      uint32_t p = seg0, prev_end = 0;
      bool have = false;
      while (p < seg1) {
        const uint32_t w = S.mask[p >> 5] >> (p & 31u);
        if (!w) {
          p = (p | 31u) + 1;
          continue;
        }
        p += (uint32_t)__ffs((int)w) - 1u;
        if (p >= seg1) break;
        const uint32_t lim = seg1 < mend ? seg1 : mend;  // matches are cut at the segment end
        if (p + 4 > lim) {
          p++;
          continue;
        }
        const uint32_t ml = 4 + match_len(S.in, p + 4, p + 4 - S.moff[p], lim - p - 4);
        S.moff[p + 1] = (uint16_t)ml;  // p + 1 is inside the match: the walks never look at it as a start
        if (!have) {
          have = true;
          my_first = p;
          first_ml = ml;
        } else {
          my_bytes += seq_size(p - prev_end, ml);
        }
        prev_end = p + ml;
        p = prev_end;
      }
      if (have) my_last_end = prev_end;
    }
    // end of the last match below my segment: inclusive max-scan of my_last_end, shifted by one lane
    uint32_t incl = my_last_end;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if ((int)lane >= d && o > incl) incl = o;
    }
    uint32_t below = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) below = 0;
    const uint32_t last_end_all = __shfl_sync(0xffffffffu, incl, 31);
    if (my_first != 0xffffffffu) my_bytes += seq_size(my_first - below, first_ml);
    // exclusive scan of the bytes every lane emits
    uint32_t ex = my_bytes;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, ex, d);
      if ((int)lane >= d) ex += o;
    }
    const uint32_t seq_total = __shfl_sync(0xffffffffu, ex, 31);
    ex -= my_bytes;
    const uint32_t tail_lit = n - last_end_all;
    const uint32_t tail_hdr = 1 + (tail_lit >= 15 ? 1 + (tail_lit - 15) / 255 : 0);
    const uint32_t total = seq_total + tail_hdr + tail_lit;
    uint8_t* out = scratch + (size_t)c * LZ_SCRATCH_STRIDE;
    if (total >= n) {  // would not shrink: the frame stores the chunk (method NONE), the gather copies the source
      if (lane == 0) csize[c] = n | 0x80000000u;
      __syncwarp();
      continue;
    }
    // ---- 3. emit ----
    // round r = the r-th taken match of every lane: the lanes write their sequences side by side instead of each
    // running through its own list (ncu: the per-lane loop ran at 3 of 32 lanes, a quarter of the kernel)
    {
      uint8_t* o = out + ex;
      uint32_t p = seg0, prev_end = below;
      const uint32_t lim = seg1 < mend ? seg1 : mend;
      for (;;) {
        bool have = false;
        while (p < seg1) {
          const uint32_t w = S.mask[p >> 5] >> (p & 31u);
          if (!w) {
            p = (p | 31u) + 1;
            continue;
          }
          p += (uint32_t)__ffs((int)w) - 1u;
          if (p + 4 > lim) {
            p++;
            continue;
          }
          have = true;
          break;
        }
        if (!__any_sync(0xffffffffu, have)) break;
        if (have) {
          const uint32_t ml = S.moff[p + 1];
          o = put_seq(o, S.in, prev_end, p - prev_end, S.moff[p], ml);
          prev_end = p + ml;
          p = prev_end;
        }
        __syncwarp();
      }
    }
    {
      uint8_t* o = out + seq_total;
      if (lane == 0) {
        *o = (uint8_t)((tail_lit >= 15 ? 15u : tail_lit) << 4);
        if (tail_lit >= 15) {
          uint32_t r = tail_lit - 15, k = 1;
          while (r >= 255) {
            o[k++] = 255;
            r -= 255;
          }
          o[k] = (uint8_t)r;
        }
        csize[c] = total;
      }
      o += tail_hdr;
      for (uint32_t i = lane; i < tail_lit; i += 32) o[i] = (uint8_t)ld1(S.in, last_end_all + i);
    }
    __syncwarp();
  }
}

// frame positions: exclusive scan of (25 + payload bytes) over the chunks; per region the offset of its first frame
__global__ void __launch_bounds__(1024) k_lz4_scan(const uint32_t* __restrict__ csize, uint32_t nchunks, uint64_t* __restrict__ cpos,
                                                   const LzRegion* __restrict__ regs, uint32_t nregs, uint64_t* __restrict__ reg_pos) {
  __shared__ uint64_t sh[1024];
  const uint32_t per = (nchunks + 1023) / 1024;
  const uint32_t a = threadIdx.x * per, b = min(nchunks, a + per);
  uint64_t sum = 0;
  for (uint32_t i = a; i < b; i++) sum += LZ_FRAME_HDR + (csize[i] & 0x7fffffffu);
  sh[threadIdx.x] = sum;
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {
    const uint64_t t = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  uint64_t run = sh[threadIdx.x] - sum;
  for (uint32_t i = a; i < b; i++) {
    cpos[i] = run;
    run += LZ_FRAME_HDR + (csize[i] & 0x7fffffffu);
  }
  if (threadIdx.x == 1023) cpos[nchunks] = sh[1023];
  __syncthreads();
  for (uint32_t r = threadIdx.x; r <= nregs; r += 1024) {
    const uint32_t fc = r < nregs ? regs[r].first_chunk : nchunks;
    reg_pos[r] = cpos[fc];
  }
}

// frames, contiguous, in chunk order
__global__ void __launch_bounds__(256) k_lz4_gather(const uint8_t* __restrict__ arena, const LzRegion* __restrict__ regs, uint32_t nregs,
                                                    uint32_t nchunks, const uint8_t* __restrict__ scratch, const uint32_t* __restrict__ csize,
                                                    const uint64_t* __restrict__ cpos, uint8_t* __restrict__ out) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), total_warps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t c = gwarp; c < nchunks; c += total_warps) {
    const uint32_t r = lz_region_of(regs, nregs, c);
    const LzRegion rg = regs[r];
    const uint64_t coff = (uint64_t)(c - rg.first_chunk) * LZ_CHUNK;
    const uint32_t n = (uint32_t)(rg.len - coff < LZ_CHUNK ? rg.len - coff : LZ_CHUNK);
    const uint32_t cs = csize[c];
    const bool stored = (cs & 0x80000000u) != 0;
    const uint32_t payload = cs & 0x7fffffffu;
    const uint8_t* src = stored ? arena + rg.off + coff : scratch + (size_t)c * LZ_SCRATCH_STRIDE;
    uint8_t* o = out + cpos[c];
    if (lane < LZ_FRAME_HDR) {
      uint32_t v = 0;  // bytes 0..15: checksum slot, filled by the host
      const uint32_t csz = payload + 9;
      if (lane == 16) v = stored ? 0x02u : 0x82u;
      else if (lane >= 17 && lane < 21) v = (csz >> ((lane - 17) * 8)) & 0xffu;
      else if (lane >= 21) v = (n >> ((lane - 21) * 8)) & 0xffu;
      o[lane] = (uint8_t)v;
    }
    o += LZ_FRAME_HDR;
    // source is 16-byte aligned (scratch stride, region alignment); the destination is not
    for (uint32_t i = lane * 4; i < payload; i += 128) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(src + i);
      const uint32_t k = payload - i < 4 ? payload - i : 4;
      for (uint32_t j = 0; j < k; j++) o[i + j] = (uint8_t)(w >> (j * 8));
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
struct LzHost {
  LzRegion* d_regs = nullptr;
  size_t regs_cap = 0;
  uint8_t* d_scratch = nullptr;
  uint32_t* d_csize = nullptr;
  uint64_t* d_cpos = nullptr;
  size_t chunks_cap = 0;
  uint8_t* d_out = nullptr;
  size_t out_cap = 0;
  uint64_t* d_reg_pos = nullptr;
  uint64_t* h_reg_pos = nullptr;  // pinned
  LzRegion* h_regs = nullptr;     // pinned
  unsigned grid = 0;
};

void lz4_free_state(dfgpu_ctx* ctx) {
  LzHost* h = reinterpret_cast<LzHost*>(ctx->lz);
  if (!h) return;
  cudaFree(h->d_regs);
  cudaFree(h->d_scratch);
  cudaFree(h->d_csize);
  cudaFree(h->d_cpos);
  cudaFree(h->d_out);
  cudaFree(h->d_reg_pos);
  cudaFreeHost(h->h_reg_pos);
  cudaFreeHost(h->h_regs);
  delete h;
  ctx->lz = nullptr;
}

// Compresses the regions of `arena` (device) into frames.  On return frame_pos[r] .. frame_pos[r + 1] is the byte
// range of region r's frames inside *d_frames (device), frame_pos[nregs] the total.  Synchronises the stream
// once (the total is needed to size the host buffer).
int lz4_compress_regions(dfgpu_ctx* ctx, const uint8_t* arena, std::vector<LzRegion>& regs, const uint8_t** d_frames,
                         std::vector<uint64_t>& frame_pos) {
  cudaStream_t s = ctx->stream;
  if (!ctx->lz) {
    LzHost* h = new LzHost();
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const int per_sm = (int)((227u * 1024u) / (sizeof(WarpSmem) * CW + 1024));
    h->grid = (unsigned)(sms * (per_sm > 0 ? per_sm : 1));
    cudaFuncSetAttribute(k_lz4_compress, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(WarpSmem) * CW));
    ctx->lz = h;
  }
  LzHost* h = reinterpret_cast<LzHost*>(ctx->lz);
  uint32_t nchunks = 0;
  for (auto& r : regs) {
    r.first_chunk = nchunks;
    nchunks += (uint32_t)((r.len + LZ_CHUNK - 1) / LZ_CHUNK);
  }
  const uint32_t nregs = (uint32_t)regs.size();
  frame_pos.assign(nregs + 1, 0);
  *d_frames = nullptr;
  if (!nchunks) return 0;
  if (nregs + 1 > h->regs_cap) {
    cudaFree(h->d_regs);
    cudaFree(h->d_reg_pos);
    cudaFreeHost(h->h_reg_pos);
    cudaFreeHost(h->h_regs);
    h->regs_cap = (size_t)nregs * 2 + 64;
    CU(cudaMalloc(&h->d_regs, h->regs_cap * sizeof(LzRegion)));
    CU(cudaMalloc(&h->d_reg_pos, h->regs_cap * 8));
    CU(cudaHostAlloc(&h->h_reg_pos, h->regs_cap * 8, cudaHostAllocDefault));
    CU(cudaHostAlloc(&h->h_regs, h->regs_cap * sizeof(LzRegion), cudaHostAllocDefault));
  }
  if (nchunks > h->chunks_cap) {
    cudaFree(h->d_scratch);
    cudaFree(h->d_csize);
    cudaFree(h->d_cpos);
    cudaFree(h->d_out);
    h->chunks_cap = (size_t)nchunks + nchunks / 4 + 256;
    CU(cudaMalloc(&h->d_scratch, h->chunks_cap * LZ_SCRATCH_STRIDE));
    CU(cudaMalloc(&h->d_csize, h->chunks_cap * 4));
    CU(cudaMalloc(&h->d_cpos, (h->chunks_cap + 1) * 8));
    h->out_cap = h->chunks_cap * (size_t)(LZ_CHUNK + LZ_FRAME_HDR);
    CU(cudaMalloc(&h->d_out, h->out_cap));
  }
  memcpy(h->h_regs, regs.data(), nregs * sizeof(LzRegion));
  CU(cudaMemcpyAsync(h->d_regs, h->h_regs, nregs * sizeof(LzRegion), cudaMemcpyHostToDevice, s));
  k_lz4_compress<<<h->grid, CW * 32, sizeof(WarpSmem) * CW, s>>>(arena, h->d_regs, nregs, nchunks, h->d_scratch, h->d_csize);
  k_lz4_scan<<<1, 1024, 0, s>>>(h->d_csize, nchunks, h->d_cpos, h->d_regs, nregs, h->d_reg_pos);
  k_lz4_gather<<<h->grid * 2, 256, 0, s>>>(arena, h->d_regs, nregs, nchunks, h->d_scratch, h->d_csize, h->d_cpos, h->d_out);
  ctx->launches += 3;
  CU(cudaMemcpyAsync(h->h_reg_pos, h->d_reg_pos, (size_t)(nregs + 1) * 8, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  CU(cudaGetLastError());
  for (uint32_t r = 0; r <= nregs; r++) frame_pos[r] = h->h_reg_pos[r];
  *d_frames = h->d_out;
  return 0;
}

}  // namespace dfgpu
