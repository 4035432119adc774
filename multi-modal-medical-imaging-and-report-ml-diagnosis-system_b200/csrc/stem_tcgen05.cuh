// ResNet stem on the tensor cores, fused with its max-pool:
//   conv 7x7 stride 2 pad 3 (3 -> 64) + folded BatchNorm + ReLU  [+ MaxPool 3x3 stride 2 pad 1]
//   (torchvision resnet50 conv1/bn1/relu/maxpool = backbone[0..3], training_pipeline.py:178-183)
//
// Input: the zero-bordered 4-channel bf16 image written by K_pre, [B, hp, wp, 4] with the picture at (3,3):
// one pixel = 8 bytes, so the 8 pixels (7 taps + one zero-weight) under a filter row are 64 contiguous bytes
// and the window of the NEXT output pixel (stride 2) starts 16 bytes later.
//
// Zero-copy im2col: a strip of raw image rows is bulk-copied (cp.async.bulk, one contiguous range) into shared
// memory ONCE, and the tensor core reads the A operand straight out of it through a no-swizzle K-major descriptor
// whose 8-row core matrices are 128 contiguous bytes: row m (= output column) at +16 m, K chunk j at +16 j
// (LBO = 16 B, SBO = 128 B) - overlapping windows, nothing is ever unfolded.  One output row of 128 columns is
// 7 filter rows x 2 tcgen05.mma (M=128, N=64, K=16) into one of eight 64-column TMEM accumulators.  The
// weights (7 filter rows x [64 x 32], 28 KB, canonical no-swizzle layout) stay resident in shared memory.
// Every input pixel is fetched from L2 about 1.1 times (strip halo) instead of ~12 times.
//
// Epilogue (thread = output column = TMEM lane): +bias, ReLU, bf16; with pooling, a running vertical max over
// conv rows 2py-1..2py+1 in registers, then the horizontal 3-max through a swizzled shared row buffer, and
// 16-byte coalesced stores of the pooled row.  Out-of-range rows/columns contribute zeros, which equals the
// -inf padding of MaxPool2d because every ReLU output is >= 0.
// Roles (320 threads): warp 0 = bulk-copy producer, warp 1 = MMA issuer + TMEM allocator, warps 2-5 / 6-9 = two
// epilogue groups, each owning 32 of the 64 output channels of every conv row (one warp per SMSP was the limiter).
#pragma once
#include "ptx.cuh"

namespace mmdx {

constexpr int STEM_THREADS = 320;
constexpr int STEM_W_BYTES = 7 * 64 * 32 * 2;          // 28 KB resident weights
constexpr int STEM_ROWBUF_BYTES = 128 * 128;           // one conv row of vertical maxima: 128 columns x 64 ch bf16
constexpr int STEM_SLACK = 2048;                       // window reads of columns past the row end stay inside smem
constexpr int STEM_ACC = 8;                            // TMEM accumulator slots (64 columns each)

struct StemParams {
  const __nv_bfloat16* in_pad;   // [B, hp, wp, 4]
  const __nv_bfloat16* w;        // [7][4 k-chunks][8 n-groups][8 rows][8] bf16  (canonical no-swizzle, per filter row)
  const float* bias;             // [64]
  __nv_bfloat16* out;            // pool ? [B, PH, PW, 64] : [B, OH, OW, 64]
  int B, hp, wp, OH, OW, PH, PW;
  int pool;                      // 1 = fused max-pool
  int rows_per_strip;            // pooled rows (pool) or conv rows (no pool) per work unit
  int strips, col_blocks;        // per image
  int cols_per_block;            // pooled columns (pool, <= 63) or conv columns (<= 128) per block
  int num_units;                 // B * strips * col_blocks
  int in_buf_bytes;              // bytes reserved per input strip buffer (incl. slack)
};

struct StemUnit { int b, c_first, c_count, j0, ox0, cb; };   // conv rows [c_first, c_first + c_count), j0 = index of
                                                              // c_first in the pooling window sequence

__device__ __forceinline__ StemUnit stem_unit(const StemParams& p, int id) {
  StemUnit u;
  const int per_img = p.strips * p.col_blocks;
  u.b = id / per_img;
  const int rem = id - u.b * per_img;
  const int strip = rem / p.col_blocks;
  u.cb = rem - strip * p.col_blocks;
  if (p.pool) {
    const int py0 = strip * p.rows_per_strip;
    const int py1 = min(py0 + p.rows_per_strip, p.PH);         // exclusive
    const int c0 = 2 * py0 - 1, c1 = 2 * (py1 - 1) + 1;        // inclusive conv-row window of the strip
    u.c_first = max(c0, 0);
    u.c_count = min(c1, p.OH - 1) - u.c_first + 1;
    u.j0 = u.c_first - c0;
    u.ox0 = 2 * (u.cb * p.cols_per_block) - 1;                 // conv column of TMEM lane 0 (may be -1)
  } else {
    u.c_first = strip * p.rows_per_strip;
    u.c_count = min(p.rows_per_strip, p.OH - u.c_first);
    u.j0 = 0;
    u.ox0 = u.cb * p.cols_per_block;
  }
  return u;
}

// no-swizzle K-major descriptor words: LBO = K-chunk stride, SBO = 8-row-group stride (bytes)
__device__ __forceinline__ uint32_t stem_desc_lo(uint32_t saddr, uint32_t lbo) {
  return ((saddr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16);
}
__device__ __forceinline__ constexpr uint32_t stem_desc_hi(uint32_t sbo) { return (sbo >> 4) | (1u << 14); }

// Row-buffer swizzle: a bijection of (m & 7) with bits 1 and 2 swapped.  Writers (8 consecutive rows, same chunk) hit
// 8 distinct 16-byte positions; readers (rows mm and mm+2, the 4 chunks of one channel half) land in different
// halves of the 128-byte row - both conflict-free.
__device__ __forceinline__ int stem_swz(int m) { return (m & 1) | ((m & 2) << 1) | ((m & 4) >> 1); }

__global__ void __launch_bounds__(STEM_THREADS, 1) stem_pool_tcgen05_kernel(const __grid_constant__ StemParams p) {
  constexpr uint32_t IDESC = make_idesc_bf16(128, 64);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wbuf = smem;                                            // 28 KB
  uint8_t* rowbuf = smem + STEM_W_BYTES;                           // 16 KB
  uint8_t* inbuf0 = rowbuf + STEM_ROWBUF_BYTES;                    // 2 x in_buf_bytes (each starts with 16 B of slack)
  uint64_t* bars = reinterpret_cast<uint64_t*>(inbuf0 + 2 * p.in_buf_bytes);
  uint64_t* in_full = bars;            // [2]
  uint64_t* in_empty = bars + 2;       // [2]
  uint64_t* acc_full = bars + 4;       // [STEM_ACC]
  uint64_t* acc_empty = bars + 4 + STEM_ACC;
  uint64_t* w_bar = bars + 4 + 2 * STEM_ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pitch = p.wp * 8;                                      // bytes per padded image row

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 1); }
    for (int i = 0; i < STEM_ACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      // ================= producer: weights once, then one contiguous strip of image rows per unit =================
      mbar_arrive_expect_tx(w_bar, STEM_W_BYTES);
      bulk_load(wbuf, p.w, STEM_W_BYTES, w_bar);
      int n = 0;
      for (int id = blockIdx.x; id < p.num_units; id += gridDim.x, ++n) {
        const StemUnit u = stem_unit(p, id);
        const int buf = n & 1;
        mbar_wait(&in_empty[buf], ((n >> 1) & 1) ^ 1);
        const int rows = 2 * (u.c_count - 1) + 7;                  // input rows 2*c_first .. 2*c_last + 6
        const uint32_t bytes = static_cast<uint32_t>(rows) * pitch;
        const __nv_bfloat16* src = p.in_pad + (static_cast<size_t>(u.b) * p.hp + 2 * u.c_first) * p.wp * 4;
        mbar_arrive_expect_tx(&in_full[buf], bytes);
        bulk_load(inbuf0 + buf * p.in_buf_bytes + 16, src, bytes, &in_full[buf]);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (warp-uniform loop, one elected lane issues) =================
    constexpr uint32_t A_HI = stem_desc_hi(128), B_HI = stem_desc_hi(128);
    mbar_wait(w_bar, 0);
    const uint32_t w_lo = stem_desc_lo(smem_u32(wbuf), 1024);
    const uint32_t p16 = static_cast<uint32_t>(pitch) >> 4;        // row pitch in descriptor units (pitch % 16 == 0)
    int n = 0;
    uint32_t row_ctr = 0;                                          // conv rows issued so far (TMEM slot ring)
    for (int id = blockIdx.x; id < p.num_units; id += gridDim.x, ++n) {
      const StemUnit u = stem_unit(p, id);
      const int buf = n & 1;
      mbar_wait(&in_full[buf], (n >> 1) & 1);
      // lane 0 of the tile = conv column ox0: its window starts 16*ox0 bytes into the row (ox0 = -1: the slack)
      uint32_t a_lo = stem_desc_lo(smem_u32(inbuf0 + buf * p.in_buf_bytes + 16) + 16 * u.ox0, 16);
      for (int c = 0; c < u.c_count; ++c, ++row_ctr, a_lo += 2 * p16) {
        const int slot = row_ctr % STEM_ACC;
        mbar_wait(&acc_empty[slot], ((row_ctr / STEM_ACC) & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + slot * 64;
          uint32_t a = a_lo;
#pragma unroll
          for (int r = 0; r < 7; ++r, a += p16) {
            umma_bf16_words<false>(d, a, A_HI, w_lo + ((r * 4096) >> 4), B_HI, IDESC, r != 0 ? 1u : 0u);
            umma_bf16_words<false>(d, a + 2, A_HI, w_lo + ((r * 4096 + 2048) >> 4), B_HI, IDESC, 1u);
          }
          umma_commit(&acc_full[slot]);
          if (c == u.c_count - 1) umma_commit(&in_empty[buf]);     // every MMA reading this strip has finished
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue warps 2..9: thread = (conv column, channel half) =================
    const int hf = (warp - 2) >> 2;                                // channels [32*hf, 32*hf + 32)
    const int q = warp & 3;                                        // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;                                   // lane of the tile = conv column ox0 + m
    const int et = ((warp - 2) & 3) * 32 + lane;                   // 0..127 index inside the group
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    float bias[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) bias[i] = __ldg(p.bias + hf * 32 + i);
    uint32_t row_ctr = 0;
    for (int id = blockIdx.x; id < p.num_units; id += gridDim.x) {
      const StemUnit u = stem_unit(p, id);
      const int ox = u.ox0 + m;
      const bool col_ok = ox >= 0 && ox < p.OW && (p.pool ? m <= 2 * p.cols_per_block : m < p.cols_per_block);
      uint32_t cur[16];                                            // running vertical max, bf16 pairs
#pragma unroll
      for (int i = 0; i < 16; ++i) cur[i] = 0u;
      for (int c = 0; c < u.c_count; ++c, ++row_ctr) {
        const int slot = row_ctr % STEM_ACC;
        mbar_wait(&acc_full[slot], (row_ctr / STEM_ACC) & 1);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + lane_base + slot * 64 + hf * 32, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[slot]);
        uint32_t row[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a0 = col_ok ? fmaxf(__uint_as_float(v[2 * i]) + bias[2 * i], 0.f) : 0.f;
          const float a1 = col_ok ? fmaxf(__uint_as_float(v[2 * i + 1]) + bias[2 * i + 1], 0.f) : 0.f;
          row[i] = pack_bf16(a0, a1);
        }
        const int oy = u.c_first + c;
        if (!p.pool) {
          if (col_ok) {
            uint4* dst = reinterpret_cast<uint4*>(p.out + ((static_cast<size_t>(u.b) * p.OH + oy) * p.OW + ox) * 64 + hf * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_uint4(row[4 * i], row[4 * i + 1], row[4 * i + 2], row[4 * i + 3]);
          }
          continue;
        }
        // ---- fused max-pool: rows j = 0,1,2 | 2,3,4 | ... of the strip's window sequence feed pooled rows 0,1,..
        const int j = u.j0 + c;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&cur[i]);
          const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&row[i]);
          const __nv_bfloat162 mx = __hmax2(a, b);
          cur[i] = *reinterpret_cast<const uint32_t*>(&mx);
        }
        const bool last_row = (c == u.c_count - 1);
        if ((j >= 2 && (j & 1) == 0) || (last_row && (j & 1) == 1)) {
          // conv row 2py+1 (or the bottom border) completes pooled row py
          const int py = (u.c_first - u.j0 + 1) / 2 + (j - 1) / 2;      // strip's first pooled row + index
          named_bar_sync(1 + hf, 128);                                    // the group has read the previous pooled row
          uint8_t* my = rowbuf + m * 128;                                 // 128-byte row: 8 chunks, XOR-swizzled
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(my + (((hf * 4 + i) ^ stem_swz(m)) << 4)) =
                make_uint4(cur[4 * i], cur[4 * i + 1], cur[4 * i + 2], cur[4 * i + 3]);
          named_bar_sync(1 + hf, 128);
          const int px0 = u.cb * p.cols_per_block;
          const int nq = min(p.cols_per_block, p.PW - px0);
          for (int t = et; t < nq * 4; t += 128) {
            const int qq = t >> 2, ch = hf * 4 + (t & 3);
            __nv_bfloat162 acc[4];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const int mm = 2 * qq + d;
              const uint4 val = *reinterpret_cast<const uint4*>(rowbuf + mm * 128 + ((ch ^ stem_swz(mm)) << 4));
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&val);
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i] = d == 0 ? h2[i] : __hmax2(acc[i], h2[i]);
            }
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&acc[0]); o.y = *reinterpret_cast<uint32_t*>(&acc[1]);
            o.z = *reinterpret_cast<uint32_t*>(&acc[2]); o.w = *reinterpret_cast<uint32_t*>(&acc[3]);
            reinterpret_cast<uint4*>(p.out + ((static_cast<size_t>(u.b) * p.PH + py) * p.PW + px0 + qq) * 64)[ch] = o;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) cur[i] = row[i];                   // row 2py+1 is also row 2(py+1)-1
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace mmdx
