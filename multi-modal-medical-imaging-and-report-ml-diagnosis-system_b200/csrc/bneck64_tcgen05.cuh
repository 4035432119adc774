// One kernel per layer-1 bottleneck of torchvision's ResNet-50 (training_pipeline.py:178-183; Bottleneck.forward),
// shifted by one convolution so that nothing needs a halo recompute:
//     t2  = relu(conv3x3(t1) + b2)                 64 -> 64, this block's conv2/bn2/relu
//     y   = relu(conv1x1(t2) + b3 + residual)      64 -> 256, conv3/bn3 + shortcut + relu   -> global (next shortcut)
//     t1' = relu(conv1x1(y) + b1')                 256 -> 64, the NEXT block's conv1/bn1/relu (C1 = 0: not computed)
// t2 and the A operand of the third GEMM never leave the SM: the accumulator of one tcgen05 GEMM is read by the epilogue
// warps (tcgen05.ld), biased / activated / rounded to bf16 and written to shared memory in the 128-byte-swizzled
// K-major layout that the next tcgen05.mma reads as its A operand (and that the TMA store of y reads as its source).
// At B = 256 a 56x56x256 tensor is 411 MB, so the three convolutions as separate launches move 1645 MB per block
// (they run at 5.5-6.2 TB/s, i.e. HBM-bound); this kernel moves t1 (103 MB, halos from L2), the residual (411), y (411)
// and t1' (103): 1028 MB.
//
// Tile = 16 rows x 8 columns of one image, as in conv3x3_c64_tcgen05.cuh: the halo (18 x 16 pixels x 64 ch, one 4-D
// TMA box, zero-filled outside the image = the padding) is loaded once and the nine taps read it through descriptors
// shifted by whole pixel rows.  CTA PAIRS (cta_group::2): the pair computes two tiles with M = 256 MMAs issued by the
// leader; each CTA holds HALF of every weight matrix (its N half), which is what lets all three weight sets stay
// resident: W2 9 x [32 x 64], W3 [128 x 64], W1' 4 x [32 x 64] = 68 KB per CTA instead of 136 KB.
// The shortcut tile arrives by TMA, one 64-channel chunk per slot of the four-slot y ring, a tile ahead of its use; the
// epilogue adds it IN PLACE (reads its row of the slot, writes the activated bf16 result back), after which the same
// slot is the source of the TMA store of y and the A operand of GEMM1'.  (A first version loaded the shortcut with
// per-thread global loads one chunk ahead: 35 % of all stall samples sat on those loads and the fused kernel was
// slower than the three launches it replaces - profiles/r01_bneck64_v1_full.md.)
//
// Roles (608 threads): warp 0 = TMA producer (weights, halos), warp 1 = TMEM allocator + MMA issue (leader CTA only),
// warps 2-17 = epilogue in four groups (16-column quarters of every 64-column chunk) x four TMEM lane quarters, thread =
// pixel - four epilogue warps per SM sub-partition, because a tile's epilogue work is a chain of latencies (tcgen05.ld,
// shared-memory round trips, proxy fences) rather than issue-bound; warp 18 = y ring: stores
// each finished chunk and refills its slot with the next tile's shortcut, so the epilogue warps never wait for each
// other or for a TMA instruction (they only arrive on mbarriers).
// MMA issue order per tile i:   GEMM3(i), GEMM2(i+1), GEMM1'(i) chunk by chunk as the epilogue delivers y.
// Epilogue order per tile i:    ep3(i) [y, four 64-channel chunks], ep2(i+1) [t2], ep1'(i) [t1'].
// Every reuse hazard except the two rings below is ordered by that program order:
//   halo ring (2): producer waits halo_empty (commit of GEMM2);  y ring (4 = the chunks of a tile): warp 10 stores chunk j
//   when the eight epilogue warps have arrived on y_ready[j], and refills slot j for the next tile once y_free[j] (commit
//   of the GEMM1' chunk that read it) has flipped and its own store of the slot has been read (bulk-group wait).
// TMEM (512 columns): D2 2 x 64 | D3 256 | D1' C1.
#pragma once
#include "ptx.cuh"

namespace mmdx {

constexpr int B64_EW = 16;                                         // epilogue warps: 4 column groups x 4 TMEM lane quarters
constexpr int B64_CW = 64 / (B64_EW / 4);                          // columns of a 64-column chunk per thread
static_assert(B64_CW == 16, "the epilogue is written for 16 columns per thread (tcgen05.ld 32x32b.x16)");
constexpr int B64_THREADS = 64 + 32 * B64_EW + 32;
constexpr int B64_HALO_W = 16, B64_HALO_H = 18;
constexpr int B64_HALO_BYTES = B64_HALO_W * B64_HALO_H * 128;     // 36 KB
constexpr int B64_W2_TAP_BYTES = 32 * 128;                         // this CTA's 32 output rows of one tap
constexpr int B64_W2_BYTES = 9 * B64_W2_TAP_BYTES;                 // 36 KB
constexpr int B64_W3_BYTES = 128 * 128;                            // this CTA's 128 of the 256 output rows
constexpr int B64_T2_BYTES = 128 * 128;
constexpr int B64_Y_BYTES = 128 * 128;                             // one 64-channel chunk of the y tile
constexpr int B64_COL_D3 = 128, B64_COL_D1 = 384;

// C1 = width of the next block's conv1 (0: not computed); DS = the shortcut is the block's 1x1 downsample conv of the
// 64-channel block input (layer1.0), computed here as a K extension of GEMM3 instead of being read back from memory;
// NH = halo buffers (1 where shared memory is short: 2-7 % slower, the next halo then loads under one epilogue phase only).
template <int C1, bool DS, int NH>
struct B64Smem {
  static constexpr int W1_CHUNK_BYTES = (C1 / 2) * 128;            // this CTA's C1/2 output rows of one 64-channel K chunk
  static constexpr int W1_BYTES = 4 * W1_CHUNK_BYTES;
  static constexpr int WD_BYTES = DS ? B64_W3_BYTES : 0;           // this CTA's 128 rows of the downsample weights
  static constexpr int XS_BYTES = DS ? B64_T2_BYTES : 0;           // the block-input tile (A operand next to t2)
  static constexpr int OFF_W3 = B64_W2_BYTES;
  static constexpr int OFF_WD = OFF_W3 + B64_W3_BYTES;
  static constexpr int OFF_W1 = OFF_WD + WD_BYTES;
  static constexpr int OFF_HALO = OFF_W1 + W1_BYTES;
  static constexpr int OFF_T2 = OFF_HALO + NH * B64_HALO_BYTES;
  static constexpr int OFF_XS = OFF_T2 + B64_T2_BYTES;
  static constexpr int OFF_Y = OFF_XS + XS_BYTES;
  static constexpr int OFF_BIAS = OFF_Y + 4 * B64_Y_BYTES;         // b2[64] | b3[256] | b1[C1] fp32
  static constexpr int OFF_BAR = OFF_BIAS + (64 + 256 + C1) * 4;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;               // + manual 1024-byte alignment slack
  static constexpr int W_BYTES = B64_W2_BYTES + B64_W3_BYTES + WD_BYTES + W1_BYTES;
  static_assert(TOTAL <= 232448, "shared-memory budget");
  static_assert(OFF_HALO % 1024 == 0 && OFF_T2 % 1024 == 0 && OFF_Y % 1024 == 0 && OFF_W1 % 1024 == 0 && OFF_XS % 1024 == 0,
                "swizzle atoms");
};

struct Bneck64Params {
  CUtensorMap tmA;     // t1  [NB,H,W,64]:  dims (64, W, H, NB), box (64, 16, 18, 1), SWIZZLE_128B
  CUtensorMap tmW2;    // W2  [64, 9*64] K-major, box (64, 32)
  CUtensorMap tmW3;    // W3  [256, 64],  box (64, 128)
  CUtensorMap tmW1;    // W1' [C1, 256],  box (64, C1/2)
  CUtensorMap tmY;     // y   [NB,H,W,256]: dims (256, W, H, NB), box (64, 8, 16, 1), SWIZZLE_128B
  CUtensorMap tmR;     // shortcut, same geometry as y                         (!DS)
  CUtensorMap tmX;     // block input [NB,H,W,64], box (64, 8, 16, 1)          (DS)
  CUtensorMap tmWd;    // downsample weights [256, 64], box (64, 128)          (DS)
  const float* b2;     // [64]
  const float* b3;     // [256]
  const float* b1;     // [C1]
  const float* bd;     // [256] downsample bias (DS), added to b3
  __nv_bfloat16* t1n;         // [NB,H,W,C1]
  int NB, H, W, tiles_w, tiles_h, num_tiles, num_items;   // item = two consecutive tiles (one per CTA of the pair)
  int reverse;         // 1: walk the items from the last to the first (see mmdx_engine::zigzag)
};

template <int C1, bool DS, int NH>
__global__ void __launch_bounds__(B64_THREADS, 1) bneck64_tcgen05_kernel(const __grid_constant__ Bneck64Params p) {
  using L = B64Smem<C1, DS, NH>;
  static_assert(C1 == 0 || C1 == 64 || C1 == 128, "next conv1 width (0 = none)");
  constexpr uint32_t IDESC2 = make_idesc_bf16(256, 64);
  constexpr uint32_t IDESC3 = make_idesc_bf16(256, 256);
  constexpr uint32_t IDESC1 = make_idesc_bf16(256, C1 > 0 ? C1 : 64);
  static_assert(B64_COL_D1 + C1 <= 512, "TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w2s = smem;
  uint8_t* w3s = smem + L::OFF_W3;
  uint8_t* w1s = smem + L::OFF_W1;
  uint8_t* halo = smem + L::OFF_HALO;
  uint8_t* t2s = smem + L::OFF_T2;
  uint8_t* xss = smem + L::OFF_XS;
  uint8_t* wds = smem + L::OFF_WD;
  uint8_t* ys = smem + L::OFF_Y;
  float* b2s = reinterpret_cast<float*>(smem + L::OFF_BIAS);
  float* b3s = b2s + 64;
  float* b1s = b3s + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* w_bar = bars;                 // leader: weights of both CTAs have landed
  uint64_t* halo_full = bars + 1;         // [2] leader: both CTAs' halo bytes
  uint64_t* halo_empty = bars + 3;        // [2] own: GEMM2 has finished reading the buffer (multicast commit)
  uint64_t* d2_full = bars + 5;           // [2] own
  uint64_t* t2_full = bars + 7;           // leader: the epilogue warps of both CTAs have written their t2 rows
  uint64_t* d3_full = bars + 8;           // own
  uint64_t* d1_full = bars + 9;           // own
  uint64_t* res_full = bars + 10;         // [4] own: the shortcut chunk has landed in slot j
  uint64_t* y_full = bars + 14;           // [4] leader: the epilogue warps of both CTAs have written y chunk j
  uint64_t* y_free = bars + 18;           // [4] own: GEMM1' has finished reading slot j
  uint64_t* y_ready = bars + 22;          // [4] own: the epilogue warps of this CTA have written y chunk j
  uint64_t* xs_full = bars + 26;          // leader: both CTAs' block-input tiles have landed (DS)
  uint64_t* xs_empty = bars + 27;         // own: GEMM3 has finished reading the tile (DS)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int per_img = p.tiles_w * p.tiles_h;
  const int n_items = p.num_items > pair ? (p.num_items - pair + num_pairs - 1) / num_pairs : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.tmA); prefetch_tensormap(&p.tmW2); prefetch_tensormap(&p.tmW3);
    if (C1 > 0) prefetch_tensormap(&p.tmW1);
    prefetch_tensormap(&p.tmY);
    if (DS) { prefetch_tensormap(&p.tmX); prefetch_tensormap(&p.tmWd); } else { prefetch_tensormap(&p.tmR); }
    mbar_init(w_bar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&halo_full[i], 1); mbar_init(&halo_empty[i], 1); mbar_init(&d2_full[i], 1); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&res_full[i], 1); mbar_init(&y_full[i], 2 * B64_EW); mbar_init(&y_free[i], 1); mbar_init(&y_ready[i], B64_EW);
    }
    mbar_init(t2_full, 2 * B64_EW); mbar_init(d3_full, 1); mbar_init(d1_full, 1);
    mbar_init(xs_full, 1); mbar_init(xs_empty, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
  for (int i = threadIdx.x; i < 64 + 256 + C1; i += B64_THREADS)
    b2s[i] = i < 64 ? __ldg(p.b2 + i)
                    : (i < 320 ? __ldg(p.b3 + i - 64) + (DS ? __ldg(p.bd + i - 64) : 0.f) : __ldg(p.b1 + i - 320));
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      // ================= producer: this CTA's half of the weights once, then one halo box per tile =================
      const uint32_t wb = mapa_rank(smem_u32(w_bar), 0);
      if (rank == 0) mbar_arrive_expect_tx(w_bar, 2u * L::W_BYTES);
      for (int t = 0; t < 9; ++t) tma_load_2d_pair(w2s + t * B64_W2_TAP_BYTES, &p.tmW2, wb, t * 64, rank * 32);
      tma_load_2d_pair(w3s, &p.tmW3, wb, 0, rank * 128);
      if constexpr (DS) tma_load_2d_pair(wds, &p.tmWd, wb, 0, rank * 128);
      if constexpr (C1 > 0)
        for (int j = 0; j < 4; ++j) tma_load_2d_pair(w1s + j * L::W1_CHUNK_BYTES, &p.tmW1, wb, j * 64, rank * (C1 / 2));
      const uint32_t hf0 = mapa_rank(smem_u32(&halo_full[0]), 0);
      const uint32_t xf = mapa_rank(smem_u32(xs_full), 0);
      int item = pair;
      for (int n = 0; n < n_items; ++n, item += num_pairs) {
        const int tile = 2 * (p.reverse ? p.num_items - 1 - item : item) + rank;               // a past-the-end tile decodes to image NB: zero-filled, clipped
        const int img = tile / per_img, rem = tile - img * per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        const int buf = n % NH;
        mbar_wait(&halo_empty[buf], ((n / NH) & 1) ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&halo_full[buf], 2u * B64_HALO_BYTES);
        tma_load_4d_pair(halo + buf * B64_HALO_BYTES, &p.tmA, hf0 + buf * 8, 0, tw * 8 - 1, th * 16 - 1, img);
        if constexpr (DS) {                             // the block-input tile, read by GEMM3 next to t2
          mbar_wait(xs_empty, (n & 1) ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(xs_full, 2u * B64_T2_BYTES);
          tma_load_4d_pair(xss, &p.tmX, xf, 0, tw * 8, th * 16, img);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ================= MMA issue (leader; warp-uniform loop, one elected lane) =================
      constexpr uint32_t HI = sdesc_hi<128>();                      // 8-row atoms 1024 B apart
      constexpr uint32_t HI_HALO = static_cast<uint32_t>(2048 >> 4) | (1u << 14) | (2u << 29);   // atoms one halo row apart
      const uint32_t w2_lo = sdesc_lo<128>(smem_u32(w2s));
      const uint32_t w3_lo = sdesc_lo<128>(smem_u32(w3s));
      const uint32_t w1_lo = sdesc_lo<128>(smem_u32(w1s));
      const uint32_t t2_lo = sdesc_lo<128>(smem_u32(t2s));
      const uint32_t xs_lo = sdesc_lo<128>(smem_u32(xss));
      const uint32_t wd_lo = sdesc_lo<128>(smem_u32(wds));
      const uint32_t y_lo = sdesc_lo<128>(smem_u32(ys));
      const uint32_t halo_lo = sdesc_lo<128>(smem_u32(halo));
      auto gemm2 = [&](int n) {           // t2 accumulator of local tile n: nine taps x four K steps, N = 64
        const int buf = n % NH, slot = n & 1;
        mbar_wait(&halo_full[buf], (n / NH) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + slot * 64;
          const uint32_t a0 = halo_lo + ((buf * B64_HALO_BYTES) >> 4);
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              const uint32_t a_lo = a0 + (((r * B64_HALO_W + s) * 128) >> 4);
              const uint32_t b_lo = w2_lo + (((r * 3 + s) * B64_W2_TAP_BYTES) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_words<true>(d, a_lo + 2 * k, HI_HALO, b_lo + 2 * k, HI, IDESC2, (r | s | k) != 0 ? 1u : 0u);
            }
          umma_commit_pair(&d2_full[slot]);
          umma_commit_pair(&halo_empty[buf]);
        }
        __syncwarp();
      };
      mbar_wait(w_bar, 0);
      if (n_items > 0) gemm2(0);
      for (int n = 0; n < n_items; ++n) {
        // ---- GEMM3(n): D3 = t2 * W3^T (K = 64, N = 256)  [+ x * Wds^T: the downsample shortcut as 64 more K]
        mbar_wait(t2_full, n & 1);
        if constexpr (DS) mbar_wait(xs_full, n & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_words<true>(tmem_base + B64_COL_D3, t2_lo + 2 * k, HI, w3_lo + 2 * k, HI, IDESC3, k != 0 ? 1u : 0u);
          if constexpr (DS) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_words<true>(tmem_base + B64_COL_D3, xs_lo + 2 * k, HI, wd_lo + 2 * k, HI, IDESC3, 1u);
            umma_commit_pair(xs_empty);
          }
          umma_commit_pair(d3_full);
        }
        __syncwarp();
        if (n + 1 < n_items) gemm2(n + 1);
        // ---- GEMM1'(n): D1 += y[:, 64j:64j+64] * W1'[:, 64j:64j+64]^T as the epilogue delivers the chunks
        if constexpr (C1 > 0) {
          for (int j = 0; j < 4; ++j) {
            mbar_wait(&y_full[j], n & 1);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_lo = y_lo + ((j * B64_Y_BYTES) >> 4);
              const uint32_t b_lo = w1_lo + ((j * L::W1_CHUNK_BYTES) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_words<true>(tmem_base + B64_COL_D1, a_lo + 2 * k, HI, b_lo + 2 * k, HI, IDESC1, (j | k) != 0 ? 1u : 0u);
              umma_commit_pair(&y_free[j]);
              if (j == 3) umma_commit_pair(d1_full);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 2 + B64_EW) {
    if (lane == 0) {
      // ================= y ring: store finished chunks, refill the slots with the next tile's shortcut =================
      auto load_res = [&](int n, int j) {              // shortcut chunk j of local tile n -> slot j
        const int tile = 2 * (p.reverse ? p.num_items - 1 - (pair + n * num_pairs) : pair + n * num_pairs) + rank;
        const int img = tile / per_img, rem = tile - img * per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        if constexpr (DS) {
          mbar_arrive(&res_full[j]);                   // nothing to load: the barrier only says "slot j is free"
        } else {
          mbar_arrive_expect_tx(&res_full[j], B64_Y_BYTES);
          tma_load_4d(ys + j * B64_Y_BYTES, &p.tmR, &res_full[j], j * 64, tw * 8, th * 16, img);
        }
      };
      if (n_items > 0)
        for (int j = 0; j < 4; ++j) load_res(0, j);
      for (int n = 0; n < n_items; ++n) {
        const int tile = 2 * (p.reverse ? p.num_items - 1 - (pair + n * num_pairs) : pair + n * num_pairs) + rank;
        const int img = tile / per_img, rem = tile - img * per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        for (int j = 0; j < 4; ++j) {
          mbar_wait(&y_ready[j], n & 1);
          tma_store_4d(&p.tmY, ys + j * B64_Y_BYTES, j * 64, tw * 8, th * 16, img);
          tma_store_commit();
          // refill the slot of the PREVIOUS chunk: its GEMM1' was triggered a chunk ago and its store is the second
          // most recent bulk group of this thread
          if (j > 0 && n + 1 < n_items) {
            if constexpr (C1 > 0) mbar_wait(&y_free[j - 1], n & 1);
            tma_store_wait_read<1>();
            load_res(n + 1, j - 1);
          }
        }
        if (n + 1 < n_items) {
          if constexpr (C1 > 0) mbar_wait(&y_free[3], n & 1);
          tma_store_wait_read<0>();
          load_res(n + 1, 3);
        }
      }
      tma_store_wait_all<0>();
    }
  } else {
    // ================= epilogue warps: thread = (pixel, 16-column group) =================
    constexpr int CW = B64_CW;                         // 16 columns = two 16-byte chunks of a 128-byte row
    const int e = warp - 2;
    const int g = e >> 2;                              // column group
    const int q = warp & 3;                            // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;                       // TMEM lane = pixel of the tile: row m >> 3, column m & 7
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t t2_full_r = mapa_rank(smem_u32(t2_full), 0);
    const uint32_t y_full_r = mapa_rank(smem_u32(&y_full[0]), 0);
    const int sw = m & 7;                              // 128-byte swizzle: 16-byte chunk index ^= row & 7

    // 16 accumulator columns + bias -> ReLU -> bf16, 16 bytes (8 channels) at a time into `row` (swizzled chunks g*2, g*2+1)
    auto bias_relu_store = [&](const uint32_t (&v)[16], const float* bias, uint8_t* row) {
      const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float4 ba = b4[2 * c], bb = b4[2 * c + 1];
        float x[8];
        unpack_f32x2(add_f32x2(pack_f32x2(__uint_as_float(v[8 * c]), __uint_as_float(v[8 * c + 1])), pack_f32x2(ba.x, ba.y)), x[0], x[1]);
        unpack_f32x2(add_f32x2(pack_f32x2(__uint_as_float(v[8 * c + 2]), __uint_as_float(v[8 * c + 3])), pack_f32x2(ba.z, ba.w)), x[2], x[3]);
        unpack_f32x2(add_f32x2(pack_f32x2(__uint_as_float(v[8 * c + 4]), __uint_as_float(v[8 * c + 5])), pack_f32x2(bb.x, bb.y)), x[4], x[5]);
        unpack_f32x2(add_f32x2(pack_f32x2(__uint_as_float(v[8 * c + 6]), __uint_as_float(v[8 * c + 7])), pack_f32x2(bb.z, bb.w)), x[6], x[7]);
        *reinterpret_cast<uint4*>(row + (((g * 2 + c) ^ sw) << 4)) =
            make_uint4(pack_bf16_relu(x[0], x[1]), pack_bf16_relu(x[2], x[3]), pack_bf16_relu(x[4], x[5]), pack_bf16_relu(x[6], x[7]));
      }
    };
    auto ep2 = [&](int n) {                            // t2 rows of local tile n -> shared memory (A operand of GEMM3)
      const int buf = n & 1;
      mbar_wait(&d2_full[buf], (n >> 1) & 1);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld_32x16(lane_base + buf * 64 + g * CW, v);
      tmem_ld_wait();
      bias_relu_store(v, b2s + g * CW, t2s + m * 128);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(t2_full_r);
    };

    if (n_items > 0) ep2(0);
    for (int n = 0; n < n_items; ++n) {
      // ---- ep3(n): y = relu(D3 + b3 + shortcut) in place in the four slots; the next chunk's accumulator load is in
      //      flight while this one is processed
      mbar_wait(d3_full, n & 1);
      tc_fence_after();
      uint32_t v[2][16];
      tmem_ld_32x16(lane_base + B64_COL_D3 + g * CW, v[0]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mbar_wait(&res_full[j], n & 1);
        uint8_t* row = ys + j * B64_Y_BYTES + m * 128;
        uint4 rv[2];
        if constexpr (!DS) {
#pragma unroll
          for (int c = 0; c < 2; ++c) rv[c] = *reinterpret_cast<const uint4*>(row + (((g * 2 + c) ^ sw) << 4));
        }
        tmem_ld_wait();
        if (j < 3) tmem_ld_32x16(lane_base + B64_COL_D3 + (j + 1) * 64 + g * CW, v[(j + 1) & 1]);
        const uint32_t (&a)[16] = v[j & 1];
        if constexpr (DS) {
          bias_relu_store(a, b3s + j * 64 + g * CW, row);
        } else {
          const float4* b4 = reinterpret_cast<const float4*>(b3s + j * 64 + g * CW);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const float4 ba = b4[2 * c], bb = b4[2 * c + 1];
            const uint32_t* ru = &rv[c].x;
            float x[8];
            unpack_f32x2(add_f32x2(add_f32x2(pack_f32x2(__uint_as_float(a[8 * c]), __uint_as_float(a[8 * c + 1])), pack_f32x2(ba.x, ba.y)), bf16x2_to_f32x2(ru[0])), x[0], x[1]);
            unpack_f32x2(add_f32x2(add_f32x2(pack_f32x2(__uint_as_float(a[8 * c + 2]), __uint_as_float(a[8 * c + 3])), pack_f32x2(ba.z, ba.w)), bf16x2_to_f32x2(ru[1])), x[2], x[3]);
            unpack_f32x2(add_f32x2(add_f32x2(pack_f32x2(__uint_as_float(a[8 * c + 4]), __uint_as_float(a[8 * c + 5])), pack_f32x2(bb.x, bb.y)), bf16x2_to_f32x2(ru[2])), x[4], x[5]);
            unpack_f32x2(add_f32x2(add_f32x2(pack_f32x2(__uint_as_float(a[8 * c + 6]), __uint_as_float(a[8 * c + 7])), pack_f32x2(bb.z, bb.w)), bf16x2_to_f32x2(ru[3])), x[6], x[7]);
            *reinterpret_cast<uint4*>(row + (((g * 2 + c) ^ sw) << 4)) =
                make_uint4(pack_bf16_relu(x[0], x[1]), pack_bf16_relu(x[2], x[3]), pack_bf16_relu(x[4], x[5]), pack_bf16_relu(x[6], x[7]));
          }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&y_ready[j]);
          if constexpr (C1 > 0) mbar_arrive_cluster(y_full_r + j * 8);
        }
      }
      // ---- ep2(n+1): the next tile's t2 (its GEMM2 ran under ep3)
      if (n + 1 < n_items) ep2(n + 1);
      // ---- ep1'(n): t1' = relu(D1 + b1') -> global
      if constexpr (C1 > 0) {
        const int tile = 2 * (p.reverse ? p.num_items - 1 - (pair + n * num_pairs) : pair + n * num_pairs) + rank;
        const int img = tile / per_img, rem = tile - img * per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        const int oh = th * 16 + (m >> 3), ow = tw * 8 + (m & 7);
        const bool pix_ok = img < p.NB && oh < p.H && ow < p.W;
        const size_t pix = (static_cast<size_t>(img) * p.H + oh) * p.W + ow;
        mbar_wait(d1_full, n & 1);
        tc_fence_after();
#pragma unroll
        for (int h2 = 0; h2 < C1 / 64; ++h2) {
          const int col0 = g * (C1 / 4) + h2 * CW;
          uint32_t v1[16];
          tmem_ld_32x16(lane_base + B64_COL_D1 + col0, v1);
          tmem_ld_wait();
          if (pix_ok) {
            const float4* b4 = reinterpret_cast<const float4*>(b1s + col0);
            uint4* dst = reinterpret_cast<uint4*>(p.t1n + pix * C1 + col0);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const float4 ba = b4[2 * c], bb = b4[2 * c + 1];
              dst[c] = make_uint4(pack_bf16_relu(__uint_as_float(v1[8 * c]) + ba.x, __uint_as_float(v1[8 * c + 1]) + ba.y),
                                  pack_bf16_relu(__uint_as_float(v1[8 * c + 2]) + ba.z, __uint_as_float(v1[8 * c + 3]) + ba.w),
                                  pack_bf16_relu(__uint_as_float(v1[8 * c + 4]) + bb.x, __uint_as_float(v1[8 * c + 5]) + bb.y),
                                  pack_bf16_relu(__uint_as_float(v1[8 * c + 6]) + bb.z, __uint_as_float(v1[8 * c + 7]) + bb.w));
            }
          }
        }
        tc_fence_before();
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace mmdx
