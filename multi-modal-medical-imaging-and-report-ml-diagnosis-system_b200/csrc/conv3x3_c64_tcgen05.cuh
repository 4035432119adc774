// 3x3 stride-1 convolution 64 -> 64 channels (+ folded BN bias, optional ReLU) over NHWC bf16: the conv2 of the three
// layer-1 bottlenecks (torchvision Bottleneck.conv2/bn2/relu, 56x56 at 224 input), the only convolutions whose
// implicit-GEMM form is L2-bound: with N = 64 a 128-pixel tile does so little math per operand byte that re-fetching
// the A tile for each of the nine taps (9 x 16 KB) plus the weights (72 KB) per tile saturates L2 -> SM bandwidth
// (measured 113 us against a 43 us tensor floor at B = 256).
//
// Here a tile is 16 rows x 8 columns of ONE image.  Its halo (18 x 16 pixels x 64 ch, one 4-D TMA box, SWIZZLE_128B,
// out-of-bounds pixels zero-filled = the padding) is loaded ONCE, and every tap reads the same shared-memory tile
// through a descriptor whose start address is shifted by (r*16 + s) pixel rows of 128 bytes: the 8-pixel groups of a
// tile row are spaced one halo row (2048 B) apart.  The 128-byte swizzle is a function of the ABSOLUTE shared-memory
// address bits - the same bits TMA used when it wrote the tile - so a start address shifted by whole pixel rows needs
// nothing else (measured on B200: declaring the shift in the descriptor's base-offset field gives wrong results,
// leaving it zero is exact).  All nine [64 x 64] weight tiles stay resident (72 KB).
// Per tile: 36 tcgen05.mma (M=128, N=64, K=16) into one slot of an 8-deep TMEM ring; thread = pixel in the epilogue.
// Roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issue, warps 2-9 = epilogue (two groups,
// one per 32-channel half).
#pragma once
#include "ptx.cuh"

namespace mmdx {

constexpr int C64_THREADS = 320;
constexpr int C64_HALO_W = 16, C64_HALO_H = 18;
constexpr int C64_HALO_BYTES = C64_HALO_W * C64_HALO_H * 128;      // 36 KB
constexpr int C64_W_BYTES = 9 * 64 * 128;                          // 72 KB
constexpr int C64_BUFS = 3;
constexpr int C64_ACC = 8;
constexpr int C64_SMEM = 1024 + C64_W_BYTES + C64_BUFS * C64_HALO_BYTES + 512;

struct C64Params {
  CUtensorMap tmA;     // input  [NB,H,W,64]: dims (64, W, H, NB), box (64, 16, 18, 1), SWIZZLE_128B
  CUtensorMap tmW;     // weights [64, 9*64] K-major: box (64, 64), SWIZZLE_128B
  const float* bias;   // [64]
  __nv_bfloat16* out;  // [NB,H,W,64]
  int NB, H, W, tiles_w, tiles_h, num_tiles, relu;
};

__global__ void __launch_bounds__(C64_THREADS, 1) conv3x3_c64_tcgen05_kernel(const __grid_constant__ C64Params p) {
  constexpr uint32_t IDESC = make_idesc_bf16(128, 64);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wres = smem;                                   // 9 x 8 KB weight tiles
  uint8_t* halo = smem + C64_W_BYTES;                     // C64_BUFS x 36 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(halo + C64_BUFS * C64_HALO_BYTES);
  uint64_t* in_full = bars;                // [C64_BUFS]
  uint64_t* in_empty = bars + C64_BUFS;    // [C64_BUFS]
  uint64_t* acc_full = bars + 2 * C64_BUFS;
  uint64_t* acc_empty = acc_full + C64_ACC;
  uint64_t* w_bar = acc_empty + C64_ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int per_img = p.tiles_w * p.tiles_h;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.tmA);
    prefetch_tensormap(&p.tmW);
    for (int i = 0; i < C64_BUFS; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 1); }
    for (int i = 0; i < C64_ACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
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
      // ================= producer: resident weights, then one halo box per tile =================
      mbar_arrive_expect_tx(w_bar, C64_W_BYTES);
      for (int t = 0; t < 9; ++t) tma_load_2d(wres + t * 8192, &p.tmW, w_bar, t * 64, 0);
      int n = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++n) {
        const int img = tile / per_img, rem = tile - img * per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        const int buf = n % C64_BUFS;
        mbar_wait(&in_empty[buf], ((n / C64_BUFS) & 1) ^ 1);
        mbar_arrive_expect_tx(&in_full[buf], C64_HALO_BYTES);
        tma_load_4d(halo + buf * C64_HALO_BYTES, &p.tmA, &in_full[buf], 0, tw * 8 - 1, th * 16 - 1, img);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issue (warp-uniform loop, one elected lane) =================
    // A: K-major SWIZZLE_128B, 8-row groups every 2048 B (one halo row of 16 pixels), base-offset field left 0
    constexpr uint32_t HI_B = sdesc_hi<128>();
    constexpr uint32_t HI_A = static_cast<uint32_t>(2048 >> 4) | (1u << 14) | (2u << 29);
    mbar_wait(w_bar, 0);
    const uint32_t w_lo = sdesc_lo<128>(smem_u32(wres));
    int n = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++n) {
      const int buf = n % C64_BUFS;
      const int slot = n % C64_ACC;
      mbar_wait(&in_full[buf], (n / C64_BUFS) & 1);
      mbar_wait(&acc_empty[slot], ((n / C64_ACC) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem_base + slot * 64;
        const uint32_t a0 = sdesc_lo<128>(smem_u32(halo + buf * C64_HALO_BYTES));
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const uint32_t a_lo = a0 + ((r * C64_HALO_W + s) * 128 >> 4);
            const uint32_t b_lo = w_lo + ((r * 3 + s) * 8192 >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_words<false>(d, a_lo + 2 * k, HI_A, b_lo + 2 * k, HI_B, IDESC, (r | s | k) != 0 ? 1u : 0u);
          }
        umma_commit(&acc_full[slot]);
        umma_commit(&in_empty[buf]);
      }
      __syncwarp();
    }
  } else {
    // ================= epilogue warps 2..9: thread = (pixel, 32-channel half) =================
    const int hf = (warp - 2) >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;                       // TMEM lane = pixel of the tile: row m >> 3, column m & 7
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    float bias[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) bias[i] = __ldg(p.bias + hf * 32 + i);
    int n = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++n) {
      const int img = tile / per_img, rem = tile - img * per_img;
      const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
      const int oh = th * 16 + (m >> 3), ow = tw * 8 + (m & 7);
      const int slot = n % C64_ACC;
      mbar_wait(&acc_full[slot], (n / C64_ACC) & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + lane_base + slot * 64 + hf * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[slot]);
      if (oh < p.H && ow < p.W) {
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a0 = __uint_as_float(v[2 * i]) + bias[2 * i], a1 = __uint_as_float(v[2 * i + 1]) + bias[2 * i + 1];
          if (p.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
          o[i] = pack_bf16(a0, a1);
        }
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((static_cast<size_t>(img) * p.H + oh) * p.W + ow) * 64 + hf * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
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
