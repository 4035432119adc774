// One persistent, warp-specialised tcgen05 kernel for every dense contraction on the path:
//   * Linear layers of the text encoder / heads:  C[M,N] = A[M,K] * W[N,K]^T  (+bias, +GELU, +residual)
//   * ResNet-50 convolutions as implicit GEMM over NHWC activations: the K loop walks
//     (filter tap, 64-channel chunk); each step is one TMA box load of a spatial tile shifted
//     by the tap offset, with TMA out-of-bounds zero fill providing the conv padding.
//     Stride-2 convs read one of four "parity" views of the input (even/odd rows x even/odd cols).
//   (The 7x7/2 stem has its own kernel, stem_tcgen05.cuh.)
//
// Roles (640 threads): warp 0 = TMA producer (A/B ring), warp 1 = TMEM allocator + single-thread MMA
// issuer, warps 4..19 = epilogue in two groups of eight (group g owns the 32-column chunks with
// index = g mod 2; inside a group four warps - one per TMEM lane quarter - own each 16-column half).
// Residual add (bottleneck shortcut, BERT skip connections) rides the MMA pipeline: after the main K loop
// the producer streams the residual tile through the same smem ring as extra A blocks (128 rows x 64
// columns) next to a 64x64 identity B block, and the issuer accumulates D[:, 64j:64j+64] += R_j * I
// with N=64 MMAs.  bf16 x 1.0 is exact in the fp32 accumulator, the loads are prefetched as deep as the
// ring, and the epilogue never touches the residual.
// Epilogue data path (EPI_TMA): TMEM -> registers (+bias, activation) -> bf16 -> swizzled staging buffer
// -> TMA store.  The output box is the same (Wb x Hb x Nb) spatial tile as the A operand, so image
// borders and ragged M are clipped by the TMA unit.
// CTA pairs (CG = 2, launched as clusters of two): the pair computes a 256 x BN tile with
// tcgen05.mma.cta_group::2 issued by the leader (cluster rank 0) only.  Each CTA TMA-loads its own 128 A rows
// and HALF of the B tile (BN/2 weight rows) into its own smem - 32 KB instead of 48 KB of L2 traffic per
// k-block at BN = 256, which is what bounds the single-CTA tiles - and owns the accumulator rows of its
// A half in its own TMEM, so the epilogue is unchanged.  Completion bytes of both CTAs' loads are posted to
// the leader's full barrier; the leader's commits are multicast to both CTAs' empty / accumulator-full
// barriers; both CTAs' epilogue warps arrive on the leader's accumulator-empty barrier.  Only the leader arms
// the full barriers (expect_tx of both CTAs' bytes): no cluster-scope release (a MEMBAR.ALL.GPU) in the loop.
// Pipelines: smem ring full/empty (TMA <-> MMA), double-buffered TMEM accumulator full/empty
// (MMA <-> epilogue), one staging buffer per epilogue group guarded by bulk-group waits + named
// barriers, static persistent tile schedule (tile = blockIdx.x + i*gridDim.x).
#pragma once
#include "ptx.cuh"

namespace mmdx {

constexpr int kMaxTaps = 12;
enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };
enum : int { EPI_DIRECT = 0, EPI_TMA = 1 };
constexpr int kEpiCW = 32;                    // epilogue chunk width (columns): 64-byte bf16 rows, SWIZZLE_64B
constexpr int kEpiGroups = 2;                 // epilogue groups (four warps each)
constexpr int kIdentBytes = 64 * 64 * 2;      // resident 64x64 identity (B operand of the residual MMAs)
constexpr int kEpiBufBytes = 128 * kEpiCW * 2;
constexpr int kEpiWarps = 16;                 // epilogue warps (two groups of eight)
constexpr int kGemmThreads = 128 + 32 * kEpiWarps;

struct alignas(64) GemmParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  CUtensorMap tmC;    // output  [.., N] bf16, box (32, Wb, Hb, Nb), SWIZZLE_64B   (EPI_TMA)
  CUtensorMap tmR;    // residual as an A operand: box (64, Wb, Hb, Nb), SWIZZLE_128B (res_blocks > 0)
  CUtensorMap tmI;    // 64x64 bf16 identity, box (64, 64 / CG), SWIZZLE_128B
  int num_k_blocks;   // taps * kb_per_tap
  int kb_per_tap;     // channel chunks per filter tap (K/BK for a plain GEMM)
  int a_box_bytes;    // bytes one A box load lands (Wb*Hb*Nb*BK*2)
  int res_blocks;     // residual k-blocks per tile (BN/64) when the residual is added by the tensor core, else 0
  int num_tiles;      // work items: ceil(m_tiles / CG) * n_tiles   (CG = 2: one item = two consecutive m-tiles)
  int n_tiles;
  int cg;             // CTAs per tile group (1, or 2 = cta_group::2 pair)
  int tiles_w, tiles_h;        // spatial tile grid (tiles over batch follow)
  int Wb, Hb, Nb;              // tile = Wb x Hb pixels x Nb images (<=128 rows)
  int OW, OH, NB;              // valid output extents (plain GEMM: OW=M, OH=NB=1)
  int act;                     // ACT_*
  int out_f32;                 // 0: bf16 output, 1: fp32 output (EPI_DIRECT only)
  int epi_mode;                // EPI_*
  long long ldc, ldr;          // output / residual row pitch in elements
  const float* bias;           // [N] or null
  const __nv_bfloat16* residual;   // [M, ldr] or null
  void* out;                   // [M, ldc]
  signed char tap_map[kMaxTaps], tap_dw[kMaxTaps], tap_dh[kMaxTaps];
  unsigned char tap_kb[kMaxTaps];   // k-blocks of each tap (all kb_per_tap, except conv3 + downsample: two "taps" over two tensors)
  // Fused row LayerNorm (template LN > 0; plain GEMMs whose N tiles cover the whole row): the schedule turns
  // item-major - one CTA group computes all n_tiles tiles of its 256 (128) rows back to back - and after the last of
  // them the epilogue warps normalise the rows this CTA has just stored (L2-hot) into ln_out.
  int reverse;                 // 1: walk the tiles from the last to the first (see mmdx_engine::zigzag)
  int item_major;              // 1: tiles of one m-item are consecutive work items of the same CTA group
  float ln_eps;
  const float* ln_gamma;       // [N]
  const float* ln_beta;        // [N]
  __nv_bfloat16* ln_out;       // [M, ld_ln]; may alias `out` (in place)
  long long ld_ln;
  // Folded LayerNorm (text encoder; DESIGN.md "LayerNorm folded into the GEMMs"): the activation tensor in memory is the
  // PRE-LayerNorm tensor x with its per-row sums (sum x, sum x^2 as 2^24 fixed-point int64, so that atomic accumulation
  // is order-independent and the result deterministic).  With mu, rstd from those sums:
  //   consumer GEMM  LN(x) W^T + b      = rstd (x W"^T) + c2,   W" = (gamma o W)(I - 11^T / n): the mean is projected out
  //                  by the weights themselves (x W"^T = (x - mu) (gamma o W)^T since mu = x 1 / n),  c2 = W beta + b
  //   residual GEMM  A W^T + b + LN(x)  = rstd ((A / rstd) W^T + x diag(gamma) - mu gamma) + (b + beta)
  //                  (the producer of A has already scaled its rows by 1 / rstd; x diag(gamma) is the residual MMA with
  //                   diag(gamma) blocks in place of the identity)
  // i.e. consumer epilogue v = rstd_row * D + bias[n] (one FMA per element, like a plain bias add); residual epilogue
  // v = rstd_row * (D - mu_row * c1[n]) + bias[n], c1 = gamma.
  int lnf;                     // 1: epilogue as above (bias = c2)
  int lnf_post;                // 1: after the activation multiply by 1 / rstd_row (FFN1: its output feeds a residual GEMM)
  int lnf_diag;                // 1: residual MMAs take diag(gamma) blocks from tmI (rows = output columns) through the ring
  float lnf_inv_n;             // 1 / width of the LayerNorm
  const float* lnf_c1;         // [N]
  const long long* lnf_stats;  // [M][2]
  unsigned long long* stat_out;   // [M][2] or null: row sums of THIS GEMM's output are accumulated here (zeroed by the host)
};

constexpr float kStatScale = 16777216.0f;     // 2^24 fixed point of the row sums

__device__ __forceinline__ void ln_row_coeffs(const long long* st, float inv_n, float eps, float& mu, float& rstd) {
  const float s1 = static_cast<float>(static_cast<double>(st[0]) * (1.0 / 16777216.0));
  const float s2 = static_cast<float>(static_cast<double>(st[1]) * (1.0 / 16777216.0));
  mu = s1 * inv_n;
  const float var = fmaxf(s2 * inv_n - mu * mu, 0.0f);
  rstd = rsqrtf(var + eps);
}

// Persistent schedule.  Default: tile = group + i * num_groups.  Item-major: group g owns items g, g + num_groups, ...
// and walks each item's n_tiles tiles consecutively (tile = item * n_tiles + n).
__device__ __forceinline__ int first_tile(const GemmParams& p, int group) {
  return p.item_major ? group * p.n_tiles : group;
}
__device__ __forceinline__ int next_tile(const GemmParams& p, int tile, int num_groups) {
  if (!p.item_major) return tile + num_groups;
  return ((tile + 1) % p.n_tiles != 0) ? tile + 1 : tile + 1 + (num_groups - 1) * p.n_tiles;
}

// LayerNorm of 2 rows of LNC * 256 columns by one warp, fp32 two-pass statistics in registers (same arithmetic as
// layernorm_kernel).  Plain coherent loads: the rows were written by this CTA's TMA stores a moment ago.
__device__ __forceinline__ uint4 ld_global_v4(const void* ptr) {
  uint4 v;
  asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr) : "memory");
  return v;
}
template <int LNC>
__device__ __forceinline__ void ln_two_rows(const __nv_bfloat16* x, long long ldx, __nv_bfloat16* y, long long ldy, int row0,
                                            int rows, const float* gamma, const float* beta, float eps, int lane) {
  constexpr int N = LNC * 256;
  float v[2][LNC][8];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = min(row0 + r, rows - 1);
    const uint4* src = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * ldx);
    uint4 raw[LNC];
#pragma unroll
    for (int c = 0; c < LNC; ++c) raw[c] = ld_global_v4(src + c * 32 + lane);
#pragma unroll
    for (int c = 0; c < LNC; ++c) {
      const uint32_t* u = &raw[c].x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16(u[j]);
        v[r][c][2 * j] = f.x; v[r][c][2 * j + 1] = f.y;
      }
    }
  }
  float mean[2], rstd[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < LNC; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[r][c][j];
    mean[r] = s;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < 2; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mean[r] *= (1.0f / N);
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < LNC; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[r][c][j] - mean[r]; q += d * d; }
    rstd[r] = q;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < 2; ++r) rstd[r] += __shfl_xor_sync(0xffffffffu, rstd[r], o);
#pragma unroll
  for (int r = 0; r < 2; ++r) rstd[r] = rsqrtf(rstd[r] * (1.0f / N) + eps);
#pragma unroll
  for (int c = 0; c < LNC; ++c) {
    const int col = (c * 32 + lane) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (row0 + r >= rows) continue;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[r][c][j] - mean[r]) * rstd[r] * gg[j] + bb[j];
      reinterpret_cast<uint4*>(y + static_cast<size_t>(row0 + r) * ldy)[c * 32 + lane] =
          make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
    }
  }
}

// EB = staging buffers per epilogue group: with two, the TMA store of chunk c reads its buffer while chunk c+2 (same
// group) is already being converted and written into the other one, and one named barrier per chunk is enough.
// RES = the GEMM can add a residual on the tensor core (needs the resident identity block).  The ring takes whatever
// shared memory is left (at most 8 stages).
template <int BN, int BK, int CG, int EB, int RES, int LNF = 0>
struct GemmSmem {
  static constexpr int A_BYTES = 128 * BK * 2;
  static constexpr int B_BYTES = (BN / CG) * BK * 2;     // this CTA's share of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_BYTES = kEpiGroups * EB * kEpiBufBytes + (RES ? kIdentBytes : 0);
  static constexpr int BIAS_BYTES = (LNF ? 4 : 2) * BN * 4;   // bias (and c1) slices of two tiles in flight
  static constexpr int ROW_BYTES = LNF ? 2 * 128 * 8 : 0;     // folded LayerNorm: (rstd, -rstd * mu) of the 128 rows, two tiles
  static constexpr int BAR_BYTES = 512 + BIAS_BYTES + ROW_BYTES;
  static constexpr int FIXED = STG_BYTES + BAR_BYTES + 1024;                // +1024: manual alignment slack
  static constexpr int STAGES = (232448 - FIXED) / STAGE_BYTES > 8 ? 8 : (232448 - FIXED) / STAGE_BYTES;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = RING_BYTES + FIXED;
  static_assert(STAGES >= 3 && TOTAL <= 232448, "shared-memory budget");
};

struct TileCoord { int n_t, w0, h0, n0; };

__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile, int rank) {
  TileCoord t;
  if (p.reverse) tile = p.num_tiles - 1 - tile;
  t.n_t = tile % p.n_tiles;
  const int m_t = (tile / p.n_tiles) * p.cg + rank;   // past-the-end m-tiles of an odd pair land out of bounds:
                                                      // TMA zero-fills the loads and clips the stores
  const int tiles_hw = p.tiles_w * p.tiles_h;
  const int tn = m_t / tiles_hw;
  const int rem = m_t - tn * tiles_hw;
  const int th = rem / p.tiles_w;
  const int tw = rem - th * p.tiles_w;
  t.w0 = tw * p.Wb; t.h0 = th * p.Hb; t.n0 = tn * p.Nb;
  return t;
}

// LNF (folded LayerNorm, see GemmParams::lnf): 0 = off (every conv and head GEMM: the code below is compiled out),
// 1 = consumer GEMM (QKV), 2 = consumer + rows scaled by 1 / rstd after the activation (FFN1), 3 = residual GEMM with
// diag(gamma) residual MMAs + row sums of the output (attention output, FFN2).
template <int BN, int BK, int CG, int EB, int RES, int LN = 0, int LNF = 0>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  using L = GemmSmem<BN, BK, CG, EB, RES, LNF>;
  constexpr int STAGES = L::STAGES;
  constexpr int SWZ = BK * 2;                 // swizzle span = one K-chunk row (128 B or 64 B)
  constexpr uint32_t IDESC = make_idesc_bf16(128 * CG, BN);
  static_assert(CG == 1 || (CG == 2 && BK == 64 && BN >= 128), "CTA pairs: BK 64, BN >= 128");
  constexpr uint32_t TMEM_COLS = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);   // double-buffered fp32 accumulator, power of two
  constexpr int NC = BN / kEpiCW;             // epilogue chunks per tile (even)
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "BN");
  static_assert(BK == 64 || BK == 32, "BK");

  static_assert(LNF == 0 || (BK == 64 && LN == 0 && (LNF == 3) == (RES == 1)), "folded LayerNorm variants");
  constexpr bool kDiag = LNF == 3;      // residual B operand = diag(gamma) block through the ring (else the resident identity)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + L::RING_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg + L::STG_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* ident_bar = tempty_bar + 2;
  uint64_t* cfull_bar = ident_bar + 1;        // tile coefficients (bias / c1 slices, row coefficients) staged  (coef warps -> epilogue)
  uint64_t* cempty_bar = cfull_bar + 2;       // ... and no longer read                                             (epilogue -> coef warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cempty_bar + 2);
  static_assert((2 * 8 + 9) * 8 + 4 <= 512, "barrier area");
  float* sbias = reinterpret_cast<float*>(stg + L::STG_BYTES + 512);       // [2][BN] bias, then [2][BN] c1 (folded LN)
  float2* srow = reinterpret_cast<float2*>(stg + L::STG_BYTES + 512 + L::BIAS_BYTES);   // [2][128] (LNF only)
  uint8_t* ident = stg + kEpiGroups * EB * kEpiBufBytes;      // 1024-byte aligned (RES only)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool res_direct = p.residual != nullptr && p.res_blocks == 0;   // EPI_DIRECT fallback only
  const int rank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;  // 0 = leader (issues the MMAs)
  const int group = blockIdx.x / CG;            // persistent schedule: item = group + i * num_groups
  const int num_groups = gridDim.x / CG;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) prefetch_tensormap(&p.tmA[i]);
    prefetch_tensormap(&p.tmB);
    if (p.epi_mode == EPI_TMA) prefetch_tensormap(&p.tmC);
    if (p.res_blocks > 0) { prefetch_tensormap(&p.tmR); prefetch_tensormap(&p.tmI); }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);               // armed by the leader's producer with the bytes of the whole group
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps * CG);        // epilogue warps of every CTA of the group
    }
    mbar_init(ident_bar, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&cfull_bar[s], 64);                      // every thread of the two coefficient warps
      mbar_init(&cempty_bar[s], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) { tmem_alloc_pair(tmem_slot, TMEM_COLS); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // the kernel that produced A / the residual has completed (everything above overlapped it)
  pdl_trigger();
  // (setmaxnreg note: registers can only be redistributed INSIDE the CTA's launch allocation, 640 x 96 = 61440; the
  // epilogue warp groups could gain at most 8 registers (128 x 64 + 512 x 104) and an .inc beyond the pool never
  // returns - measured as a hang - so the kernel keeps the uniform 96.)

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer: A/B ring (every CTA loads its own A rows and its share of B) =================
      int stage = 0;
      uint32_t phase = 0;
      // CG = 2: completion is tracked by the LEADER's full barriers (shared::cluster addresses of rank 0)
      const uint32_t full0 = CG == 2 ? mapa_rank(smem_u32(&full_bar[0]), 0) : 0;
      if (RES && BK == 64 && p.res_blocks > 0 && !kDiag) {   // identity block (this CTA's rows of it): loaded once, stays resident
        if constexpr (CG == 2) {
          if (rank == 0) mbar_arrive_expect_tx(ident_bar, kIdentBytes);      // both halves
          tma_load_2d_pair(ident, &p.tmI, mapa_rank(smem_u32(ident_bar), 0), 0, rank * 32);
        } else {
          mbar_arrive_expect_tx(ident_bar, kIdentBytes);
          tma_load_2d(ident, &p.tmI, ident_bar, 0, 0);
        }
      }
      for (int tile = first_tile(p, group); tile < p.num_tiles; tile = next_tile(p, tile, num_groups)) {
        const TileCoord t = decode_tile(p, tile, rank);
        int tap = 0, cc = 0;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if constexpr (CG == 2) {
            // The leader arms its barrier with the bytes of BOTH CTAs; the peer only issues its loads.  Peer bytes
            // that land before the leader's expect_tx just drive the tx-count negative for a moment: the phase
            // cannot complete before the (single) pending arrival.
            const uint32_t fb = full0 + stage * 8;
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * static_cast<uint32_t>(p.a_box_bytes + L::B_BYTES));
            tma_load_4d_pair(sa, &p.tmA[p.tap_map[tap]], fb, cc * BK, t.w0 + p.tap_dw[tap], t.h0 + p.tap_dh[tap], t.n0);
            tma_load_2d_pair(sb, &p.tmB, fb, kb * BK, t.n_t * BN + rank * (BN / 2));
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.a_box_bytes + L::B_BYTES));
            tma_load_4d(sa, &p.tmA[p.tap_map[tap]], &full_bar[stage], cc * BK, t.w0 + p.tap_dw[tap],
                        t.h0 + p.tap_dh[tap], t.n0);
            tma_load_2d(sb, &p.tmB, &full_bar[stage], kb * BK, t.n_t * BN);
          }
          if (++cc == p.tap_kb[tap]) { cc = 0; ++tap; }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (BK == 64 && RES == 1) {
          for (int j = 0; j < p.res_blocks; ++j) {      // residual tile as extra A blocks + identity B block
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * L::STAGE_BYTES;
            uint8_t* sb = sa + L::A_BYTES;
            // folded LN: the B operand of this residual block is diag(gamma[64 b .. 64 b + 63]), rows 64 b.. of tmI
            constexpr uint32_t dg = kDiag ? static_cast<uint32_t>(kIdentBytes / CG) : 0u;
            const int drow = t.n_t * BN + j * 64;
            if constexpr (CG == 2) {
              const uint32_t fb = full0 + stage * 8;
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * (static_cast<uint32_t>(p.a_box_bytes) + dg));
              tma_load_4d_pair(sa, &p.tmR, fb, t.n_t * BN + j * 64, t.w0, t.h0, t.n0);
              if (dg) tma_load_2d_pair(sb, &p.tmI, fb, 0, drow + rank * 32);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.a_box_bytes) + dg);
              tma_load_4d(sa, &p.tmR, &full_bar[stage], t.n_t * BN + j * 64, t.w0, t.h0, t.n0);
              if (dg) tma_load_2d(sb, &p.tmI, &full_bar[stage], 0, drow);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ================= MMA issuer (leader CTA) =================
      // The whole warp walks the pipeline (warp-uniform control flow keeps addresses and descriptors on the
      // uniform datapath); one elected lane issues the tcgen05 instructions.  Descriptor words are computed once:
      // stage s / K step k only add (s * STAGE_BYTES + k * 32) >> 4 to the low word.  At ~140 instructions per
      // k-block the single issuing thread was as slow as the four MMAs it feeds (ncu: 80 % of its samples in
      // issue code, tensor pipe 72 % busy) - this loop is what keeps the tensor pipe fed.
      auto commit = [](uint64_t* bar) {
        if constexpr (CG == 2) umma_commit_pair(bar); else umma_commit(bar);
      };
      constexpr uint32_t HI = sdesc_hi<SWZ>();
      constexpr uint32_t STAGE_STEP = L::STAGE_BYTES >> 4;
      const uint32_t a_lo0 = sdesc_lo<SWZ>(smem_u32(smem));
      const uint32_t b_lo0 = sdesc_lo<SWZ>(smem_u32(smem) + L::A_BYTES);
      const uint32_t id_lo = sdesc_lo<128>(smem_u32(ident));
      int stage = 0;
      uint32_t phase = 0, soff = 0;              // soff = stage * STAGE_STEP
      int it = 0;
      if (RES && BK == 64 && p.res_blocks > 0 && !kDiag) mbar_wait(ident_bar, 0);
      const int nkb = p.num_k_blocks, nres = (BK == 64 && RES) ? p.res_blocks : 0;
      for (int tile = first_tile(p, group); tile < p.num_tiles; tile = next_tile(p, tile, num_groups), ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);     // the epilogue (of both CTAs) has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        uint32_t acc = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);         // TMA bytes (of both CTAs) have landed
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              umma_bf16_words<CG == 2>(tmem_d, a_lo0 + soff + 2 * k, HI, b_lo0 + soff + 2 * k, HI, IDESC, acc);
              acc = 1;
            }
            commit(&empty_bar[stage]);                // frees the smem slot (in both CTAs) when these MMAs finish
            if (kb == nkb - 1 && nres == 0) commit(&tfull_bar[as]);
          }
          __syncwarp();
          acc = 1;
          soff += STAGE_STEP;
          if (++stage == STAGES) { stage = 0; phase ^= 1; soff = 0; }
        }
        if constexpr (BK == 64 && RES == 1) {
          constexpr uint32_t IDESC64 = make_idesc_bf16(128 * CG, 64);
          constexpr uint32_t HI128 = sdesc_hi<128>();
          for (int j = 0; j < nres; ++j) {              // D[:, 64j:64j+64] += R_j * I64
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_words<CG == 2>(tmem_d + j * 64, a_lo0 + soff + 2 * k, HI128,
                                         (kDiag ? b_lo0 + soff : id_lo) + 2 * k, HI128, IDESC64, 1u);
              commit(&empty_bar[stage]);
              if (j == nres - 1) commit(&tfull_bar[as]);
            }
            __syncwarp();
            soff += STAGE_STEP;
            if (++stage == STAGES) { stage = 0; phase ^= 1; soff = 0; }
          }
        }
      }
    }
  } else if (warp < 4) {
    // ================= coefficient warps 2, 3 =================
    // Everything the epilogue of a tile needs from global memory - the bias (and, folded LayerNorm, c1) slice of its
    // columns and the (rstd, -rstd * mu) pair of each of its 128 rows - is fetched here, up to two tiles ahead, and
    // handed over through shared memory: the epilogue warps never wait for a global load.
    if (p.epi_mode == EPI_TMA) {
      const int ct = threadIdx.x - 64;        // 0..63
      int it = 0;
      for (int tile = first_tile(p, group); tile < p.num_tiles; tile = next_tile(p, tile, num_groups), ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        const TileCoord t = decode_tile(p, tile, rank);
        float vb[BN / 64], vc[BN / 64];
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) {
          vb[j] = p.bias != nullptr ? __ldg(p.bias + t.n_t * BN + ct + 64 * j) : 0.0f;
          if constexpr (LNF == 3) vc[j] = __ldg(p.lnf_c1 + t.n_t * BN + ct + 64 * j);
        }
        float2 rc[2];
        if constexpr (LNF > 0) {
          // plain GEMM: global rows of this CTA's tile (not t.w0: the past-the-end m-tile of an odd pair decodes to w0 = 0
          // of the next "image" - its loads are zero-filled and its stores clipped, but its rows must stay past the end)
          const int row0 = (((p.reverse ? p.num_tiles - 1 - tile : tile) / p.n_tiles) * CG + rank) * 128;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            float mu, rstd;
            ln_row_coeffs(p.lnf_stats + 2 * static_cast<size_t>(min(row0 + ct + 64 * j, p.OW - 1)), p.lnf_inv_n, p.ln_eps, mu, rstd);
            rc[j] = make_float2(rstd, -rstd * mu);
          }
        }
        mbar_wait(&cempty_bar[as], aphase ^ 1);          // the epilogue of two tiles ago has finished with this slot
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) {
          sbias[as * BN + ct + 64 * j] = vb[j];
          if constexpr (LNF == 3) sbias[(2 + as) * BN + ct + 64 * j] = vc[j];
        }
        if constexpr (LNF > 0) { srow[as * 128 + ct] = rc[0]; srow[as * 128 + ct + 64] = rc[1]; }
        mbar_arrive(&cfull_bar[as]);
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue warps 4..19 =================
    // Two groups of eight warps; group g owns the 32-column chunks with (c & 1) == g and, inside the group, each set
    // of four warps (one per TMEM lane quarter) owns one 16-column half of the chunk.  Four epilogue warps per SM
    // sub-partition: a chunk is a chain of latencies (tcgen05.ld, MUFU, shared-memory round trip, proxy fence), and with
    // two warps per sub-partition the erf-GELU epilogue of FFN1 did not fit under the MMAs of the next tile.
    constexpr int TW = kEpiCW / 2;          // columns per thread
    const int e = warp - 4;
    const int g = e >> 3;                   // group: chunks c with (c & 1) == g
    const int hh = (e >> 2) & 1;            // column half of the chunk
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;            // row inside the 128-row tile
    const bool issuer = (e & 7) == 0 && lane == 0;
    // accumulator-empty barriers live in the leader CTA (its MMA thread waits on them)
    const uint32_t tempty0 = CG == 2 ? mapa_rank(smem_u32(&tempty_bar[0]), 0) : 0;
    auto release_acc = [&](int as) {
      if constexpr (CG == 2) mbar_arrive_cluster(tempty0 + as * 8); else mbar_arrive(&tempty_bar[as]);
    };
    int it = 0;
    uint32_t nstore = 0;                    // TMA stores issued by this group so far (selects the staging buffer)
    for (int tile = first_tile(p, group); tile < p.num_tiles; tile = next_tile(p, tile, num_groups), ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const TileCoord t = decode_tile(p, tile, rank);
      // plain GEMM: global row of this thread (see the coefficient warps for why this is not t.w0 + r)
      const int grow = (((p.reverse ? p.num_tiles - 1 - tile : tile) / p.n_tiles) * CG + rank) * 128 + r;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + hh * TW;

      if (p.epi_mode == EPI_TMA) {
        const int sw = (r >> 1) & 3;          // SWIZZLE_64B: 16-byte chunk index ^= address bits [7:8]
        // bias (and c1) slices of this tile and - folded LayerNorm - the row coefficients were staged by the coefficient warps
        const float* sb = sbias + as * BN;
        const float* sc = sbias + (2 + as) * BN;
        mbar_wait(&cfull_bar[as], aphase);
        // folded LayerNorm: this thread's row coefficients  v = ra * D + (rb * c1[n] + c2[n])
        float ra = 1.0f, rb = 0.0f, rsig = 1.0f;
        uint64_t st_sum = 0, st_sq = 0;       // packed (even, odd column) sums of v and v^2 over this thread's share of the row (LNF 3)
        if constexpr (LNF > 0) {
          const float2 rc = srow[as * 128 + r];
          ra = rc.x; rb = rc.y;
          if constexpr (LNF == 2) rsig = 1.0f / ra;
        }
        uint32_t v[TW];
        tmem_ld_32x16(tbase + g * kEpiCW, v);
#pragma unroll 1
        for (int c = g; c < NC; c += 2, ++nstore) {
          uint8_t* buf = stg + (g * EB + (EB == 2 ? (nstore & 1) : 0)) * kEpiBufBytes;
          tmem_ld_wait();
          const int col0 = t.n_t * BN + c * kEpiCW;
          float x[TW];
#pragma unroll
          for (int j = 0; j < TW; ++j) x[j] = __uint_as_float(v[j]);
          if (c + 2 < NC) {                   // the next chunk's accumulator load is in flight during this chunk's math
            tmem_ld_32x16(tbase + (c + 2) * kEpiCW, v);
          } else {                            // this warp's last read of the accumulator: hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(as);
          }
          if constexpr (LNF == 1 || LNF == 2) {
            const float4* b4 = reinterpret_cast<const float4*>(sb + c * kEpiCW + hh * TW);
            const uint64_t ra2 = pack_f32x2(ra, ra);
#pragma unroll
            for (int j = 0; j < TW / 4; ++j) {              // v = ra * D + c2, two columns per instruction
              const float4 b = b4[j];
              unpack_f32x2(fma_f32x2(ra2, pack_f32x2(x[4 * j], x[4 * j + 1]), pack_f32x2(b.x, b.y)), x[4 * j], x[4 * j + 1]);
              unpack_f32x2(fma_f32x2(ra2, pack_f32x2(x[4 * j + 2], x[4 * j + 3]), pack_f32x2(b.z, b.w)), x[4 * j + 2], x[4 * j + 3]);
            }
          } else if constexpr (LNF == 3) {
            const float4* b4 = reinterpret_cast<const float4*>(sb + c * kEpiCW + hh * TW);
            const float4* c4 = reinterpret_cast<const float4*>(sc + c * kEpiCW + hh * TW);
            const uint64_t ra2 = pack_f32x2(ra, ra), rb2 = pack_f32x2(rb, rb);
#pragma unroll
            for (int j = 0; j < TW / 4; ++j) {              // v = ra * D + (rb * c1 + c2), two columns per instruction
              const float4 b = b4[j], cc = c4[j];
              unpack_f32x2(fma_f32x2(ra2, pack_f32x2(x[4 * j], x[4 * j + 1]),
                                     fma_f32x2(rb2, pack_f32x2(cc.x, cc.y), pack_f32x2(b.x, b.y))), x[4 * j], x[4 * j + 1]);
              unpack_f32x2(fma_f32x2(ra2, pack_f32x2(x[4 * j + 2], x[4 * j + 3]),
                                     fma_f32x2(rb2, pack_f32x2(cc.z, cc.w), pack_f32x2(b.z, b.w))), x[4 * j + 2], x[4 * j + 3]);
            }
          } else {
            const float4* b4 = reinterpret_cast<const float4*>(sb + c * kEpiCW + hh * TW);
#pragma unroll
            for (int j = 0; j < TW / 4; ++j) {              // packed adds: two columns per instruction
              const float4 b = b4[j];
              unpack_f32x2(add_f32x2(pack_f32x2(x[4 * j], x[4 * j + 1]), pack_f32x2(b.x, b.y)), x[4 * j], x[4 * j + 1]);
              unpack_f32x2(add_f32x2(pack_f32x2(x[4 * j + 2], x[4 * j + 3]), pack_f32x2(b.z, b.w)), x[4 * j + 2],
                           x[4 * j + 3]);
            }
          }
          if (p.act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < TW; ++j) x[j] = fmaxf(x[j], 0.0f);
          } else if (p.act == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < TW; j += 2) gelu_erf_x2(x[j], x[j + 1]);
          }
          if constexpr (LNF == 2) {
            const uint64_t sg2 = pack_f32x2(rsig, rsig);
#pragma unroll
            for (int j = 0; j < TW; j += 2) unpack_f32x2(mul_f32x2(pack_f32x2(x[j], x[j + 1]), sg2), x[j], x[j + 1]);
          }
          if constexpr (LNF == 3) {
#pragma unroll
            for (int j = 0; j < TW; j += 2) {
              const uint64_t xx = pack_f32x2(x[j], x[j + 1]);
              st_sum = add_f32x2(st_sum, xx);
              st_sq = fma_f32x2(xx, xx, st_sq);
            }
          }
          if constexpr (EB == 1) {
            if (issuer) tma_store_wait_read<0>();   // the previous store has finished reading `buf`
            named_bar_sync(1 + g, 16 * kEpiWarps);
          }
          uint8_t* my_row = buf + r * 64;
#pragma unroll
          for (int j = 0; j < TW / 8; ++j)
            *reinterpret_cast<uint4*>(my_row + (((hh * (TW / 8) + j) ^ sw) << 4)) =
                make_uint4(pack_bf16(x[8 * j], x[8 * j + 1]), pack_bf16(x[8 * j + 2], x[8 * j + 3]),
                           pack_bf16(x[8 * j + 4], x[8 * j + 5]), pack_bf16(x[8 * j + 6], x[8 * j + 7]));
          fence_proxy_async();                // make the generic-proxy smem writes visible to the TMA unit
          if constexpr (EB == 2) {
            // One barrier per chunk: before it the issuer makes sure the store that last used the OTHER buffer (issued a
            // whole chunk ago) has finished reading it, so after the barrier everybody may write the next chunk there.
            if (issuer) tma_store_wait_read<0>();
          }
          named_bar_sync(1 + g, 16 * kEpiWarps);
          if (issuer) {
            tma_store_4d(&p.tmC, buf, col0, t.w0, t.h0, t.n0);
            tma_store_commit();
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&cempty_bar[as]);       // this warp no longer reads the tile's staged coefficients
        if constexpr (LNF == 3) {
          if (grow < p.OW) {
            // this thread's share of the row sums of the tile it has just written; int64 fixed point: integer atomics
            // commute, so the totals do not depend on the order in which tiles and threads arrive
            float a0, a1, q0, q1;
            unpack_f32x2(st_sum, a0, a1);
            unpack_f32x2(st_sq, q0, q1);
            atomicAdd(p.stat_out + 2 * static_cast<size_t>(grow), static_cast<unsigned long long>(__float2ll_rn((a0 + a1) * kStatScale)));
            atomicAdd(p.stat_out + 2 * static_cast<size_t>(grow) + 1, static_cast<unsigned long long>(__float2ll_rn((q0 + q1) * kStatScale)));
          }
        }
        if constexpr (LN > 0) {
          if (t.n_t == p.n_tiles - 1) {
            // Row LayerNorm of the 128 rows this CTA has just completed.  The accumulators were released above, so the
            // MMA warp is already two tiles into the next item while this runs.
            if (issuer) tma_store_wait_all<0>();      // this group's stores of the item have landed (not just been read)
            named_bar_sync(3, 32 * kEpiWarps);
            const int rows = p.OW;                    // plain GEMM: OW = M
            const int m_row0 = (((p.reverse ? p.num_tiles - 1 - tile : tile) / p.n_tiles) * CG + rank) * 128;   // (not t.w0: a past-the-end m-tile wraps to 0)
            const __nv_bfloat16* xin = static_cast<const __nv_bfloat16*>(p.out);
            constexpr int RW = 128 / kEpiWarps;       // rows per warp
#pragma unroll 1
            for (int rr = e * RW; rr < e * RW + RW; rr += 2) {
              const int row0 = m_row0 + rr;
              if (row0 < rows) ln_two_rows<LN>(xin, p.ldc, p.ln_out, p.ld_ln, row0, rows, p.ln_gamma, p.ln_beta, p.ln_eps, lane);
            }
          }
        }
      } else {
        // ---------- EPI_DIRECT: per-thread row-contiguous global I/O (fp32 outputs, unaligned pitches) ----------
        const int rows_valid = p.Wb * p.Hb * p.Nb;
        const int wi = r % p.Wb;
        const int hi = (r / p.Wb) % p.Hb;
        const int ni = r / (p.Wb * p.Hb);
        const int w = t.w0 + wi, h = t.h0 + hi, n = t.n0 + ni;
        const bool valid = (r < rows_valid) && (w < p.OW) && (h < p.OH) && (n < p.NB);
        const long long orow = (static_cast<long long>(n) * p.OH + h) * p.OW + w;
#pragma unroll 1
        for (int c = g; c < NC; c += 2) {
          uint32_t v[TW];
          tmem_ld_32x16(tbase + c * kEpiCW, v);
          tmem_ld_wait();
          if (c == NC - 2 + g) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(as);
          }
          if (valid) {
            const int col0 = t.n_t * BN + c * kEpiCW + hh * TW;
            float x[TW];
#pragma unroll
            for (int j = 0; j < TW; ++j) x[j] = __uint_as_float(v[j]);
            if (p.bias != nullptr) {
              const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
              for (int j = 0; j < TW / 4; ++j) {
                const float4 b = __ldg(b4 + j);
                x[4 * j + 0] += b.x; x[4 * j + 1] += b.y; x[4 * j + 2] += b.z; x[4 * j + 3] += b.w;
              }
            }
            if (res_direct) {
              const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + orow * p.ldr + col0);
#pragma unroll
              for (int j = 0; j < TW / 8; ++j) {
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
              for (int j = 0; j < TW; ++j) x[j] = fmaxf(x[j], 0.0f);
            } else if (p.act == ACT_GELU) {
#pragma unroll
              for (int j = 0; j < TW; ++j) x[j] = gelu_erf(x[j]);
            }
            if (p.out_f32) {
              float4* o4 = reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow * p.ldc + col0);
#pragma unroll
              for (int j = 0; j < TW / 4; ++j) o4[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
            } else {
              uint4* o4 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + orow * p.ldc + col0);
#pragma unroll
              for (int j = 0; j < TW / 8; ++j)
                o4[j] = make_uint4(pack_bf16(x[8 * j], x[8 * j + 1]), pack_bf16(x[8 * j + 2], x[8 * j + 3]),
                                   pack_bf16(x[8 * j + 4], x[8 * j + 5]), pack_bf16(x[8 * j + 6], x[8 * j + 7]));
            }
          }
        }
      }
    }
    if (issuer && p.epi_mode == EPI_TMA) tma_store_wait_all<0>();   // all output bytes are in global memory
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // pair: the peer may still signal this CTA's barriers
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace mmdx
