// One persistent, warp-specialised tcgen05 kernel for every dense contraction on the path:
//   * Linear layers of the text encoder / heads:  C[M,N] = A[M,K] * W[N,K]^T  (+bias, +GELU, +residual)
//   * ResNet-50 convolutions as implicit GEMM over NHWC activations: the K loop walks
//     (filter tap, 64-channel chunk); each step is one TMA box load of a spatial tile shifted
//     by the tap offset, with TMA out-of-bounds zero fill providing the conv padding.
//     Stride-2 convs read one of four "parity" views of the input (even/odd rows x even/odd cols).
//   * The 7x7/2 stem over a zero-padded 4-channel image: K chunk = one filter row
//     (8 pixels x 4 channels = 32 contiguous bf16), fetched through a tensor map whose
//     W dimension advances by 2 pixels (16 B) - overlapping windows, no im2col buffer.
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> bias/act/residual -> global).
// Pipelines: smem ring full/empty (TMA <-> MMA), double-buffered TMEM accumulator full/empty
// (MMA <-> epilogue), static persistent tile schedule (tile = blockIdx.x + i*gridDim.x).
#pragma once
#include "ptx.cuh"

namespace mmdx {

constexpr int kMaxTaps = 12;
enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

struct alignas(64) GemmParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  int num_k_blocks;   // taps * kb_per_tap
  int kb_per_tap;     // channel chunks per filter tap (K/BK for a plain GEMM)
  int a_box_bytes;    // bytes one A box load lands (Wb*Hb*Nb*BK*2)
  int num_tiles;      // m_tiles * n_tiles
  int n_tiles;
  int tiles_w, tiles_h;        // spatial tile grid (tiles over batch follow)
  int Wb, Hb, Nb;              // tile = Wb x Hb pixels x Nb images (<=128 rows)
  int OW, OH, NB;              // valid output extents (plain GEMM: OW=M, OH=NB=1)
  int act;                     // ACT_*
  int out_f32;                 // 0: bf16 output, 1: fp32 output
  long long ldc, ldr;          // output / residual row pitch in elements
  const float* bias;           // [N] or null
  const __nv_bfloat16* residual;   // [M, ldr] or null
  void* out;                   // [M, ldc]
  signed char tap_map[kMaxTaps], tap_dw[kMaxTaps], tap_dh[kMaxTaps];
};

template <int BN, int BK, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = 128 * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;   // +1024: manual alignment slack
};

template <int BN, int BK, int STAGES>
__global__ void __launch_bounds__(192, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  using L = GemmSmem<BN, BK, STAGES>;
  constexpr int SWZ = BK * 2;                 // swizzle span = one K-chunk row (128 B or 64 B)
  constexpr uint32_t IDESC = make_idesc_bf16(128, BN);
  constexpr uint32_t TMEM_COLS = 2 * BN;      // double-buffered fp32 accumulator
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  static_assert(BK == 64 || BK == 32, "BK");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * L::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) prefetch_tensormap(&p.tmA[i]);
    prefetch_tensormap(&p.tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_hw = p.tiles_w * p.tiles_h;

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer =================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n_t = tile % p.n_tiles;
        const int m_t = tile / p.n_tiles;
        const int tn = m_t / tiles_hw;
        const int rem = m_t - tn * tiles_hw;
        const int th = rem / p.tiles_w;
        const int tw = rem - th * p.tiles_w;
        const int w0 = tw * p.Wb, h0 = th * p.Hb, n0 = tn * p.Nb;
        int tap = 0, cc = 0;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.a_box_bytes + L::B_BYTES));
          tma_load_4d(sa, &p.tmA[p.tap_map[tap]], &full_bar[stage], cc * BK, w0 + p.tap_dw[tap], h0 + p.tap_dh[tap],
                      n0);
          tma_load_2d(sb, &p.tmB, &full_bar[stage], kb * BK, n_t * BN);
          if (++cc == p.kb_per_tap) { cc = 0; ++tap; }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer (one thread) =================
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);     // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);         // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_bf16(tmem_d, make_sdesc<SWZ>(sa + k * 32), make_sdesc<SWZ>(sb + k * 32), IDESC,
                      (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);             // frees the smem slot when these MMAs finish
          if (kb == p.num_k_blocks - 1) umma_commit(&tfull_bar[as]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ================= epilogue warps 2..5 =================
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;            // row inside the 128-row tile
    const int rows_valid = p.Wb * p.Hb * p.Nb;
    const int wi = r % p.Wb;
    const int hi = (r / p.Wb) % p.Hb;
    const int ni = r / (p.Wb * p.Hb);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int n_t = tile % p.n_tiles;
      const int m_t = tile / p.n_tiles;
      const int tn = m_t / tiles_hw;
      const int rem = m_t - tn * tiles_hw;
      const int th = rem / p.tiles_w;
      const int tw = rem - th * p.tiles_w;
      const int w = tw * p.Wb + wi, h = th * p.Hb + hi, n = tn * p.Nb + ni;
      const bool valid = (r < rows_valid) && (w < p.OW) && (h < p.OH) && (n < p.NB);
      const long long orow = (static_cast<long long>(n) * p.OH + h) * p.OW + w;

      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + c * 32, v);
        tmem_ld_wait();
        if (c == BN / 32 - 1) {             // accumulator fully read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        if (valid) {
          const int col0 = n_t * BN + c * 32;
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
          if (p.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(b4 + j);
              x[4 * j + 0] += b.x; x[4 * j + 1] += b.y; x[4 * j + 2] += b.z; x[4 * j + 3] += b.w;
            }
          }
          if (p.residual != nullptr) {
            const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + orow * p.ldr + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 u = __ldg(r4 + j);
              float2 f;
              f = unpack_bf16(u.x); x[8 * j + 0] += f.x; x[8 * j + 1] += f.y;
              f = unpack_bf16(u.y); x[8 * j + 2] += f.x; x[8 * j + 3] += f.y;
              f = unpack_bf16(u.z); x[8 * j + 4] += f.x; x[8 * j + 5] += f.y;
              f = unpack_bf16(u.w); x[8 * j + 6] += f.x; x[8 * j + 7] += f.y;
            }
          }
          if (p.act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.0f);
          } else if (p.act == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = gelu_erf(x[j]);
          }
          if (p.out_f32) {
            float4* o4 = reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow * p.ldc + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) o4[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
          } else {
            uint4* o4 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + orow * p.ldc + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o4[j] = make_uint4(pack_bf16(x[8 * j], x[8 * j + 1]), pack_bf16(x[8 * j + 2], x[8 * j + 3]),
                                 pack_bf16(x[8 * j + 4], x[8 * j + 5]), pack_bf16(x[8 * j + 6], x[8 * j + 7]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace mmdx
