// conv3 of a ResNet bottleneck and conv1 of the NEXT block as ONE persistent launch (layers 2-3).
//
//   y   [M, N1] = relu(t2 [M, K1] * W3^T  + b3 + shortcut [M, N1])          (torchvision Bottleneck conv3 + bn3 + add + relu)
//   t1n [M, N2] = relu(y  [M, N1] * W1n^T + b1n)                            (next block's conv1 + bn1 + relu)
//
// Both are 1x1 convolutions over NHWC, i.e. plain GEMMs over the M = B*H*W pixel rows.  Run as two launches, the second
// one re-reads y from DRAM a whole kernel later (layer 2: 205 MB at 5 TB/s = the entire 44-50 us of that launch); here a
// CTA pair owns a 256-row ITEM, computes the N1 / BN tiles of y for it, and then the N2 / BN tiles of t1n from the rows
// it has just stored - read back through TMA while they are still in L2.  y is written once (the next block's shortcut
// needs it) and never read from DRAM by this kernel.
//
// Schedule per CTA pair (static, persistent): G1(item 0), G1(item 1), G2(item 0), G1(item 2), G2(item 1), ... so that the
// stores of an item have long landed when its G2 loads are issued.  The dependency is explicit all the same: after the
// last G1 tile of an item each epilogue group's issuing thread waits for ITS bulk stores to complete (not just to have
// been read) and arrives on `ydone`; the TMA producer waits on it before the first G2 load of that item.
//
// Everything else is gemm_tcgen05.cuh's CTA-pair path, specialised: 640 threads (warp 0 TMA producer, warp 1 TMEM
// allocator + MMA issue, warps 2-3 bias staging, warps 4-19 epilogue in two groups with two staging buffers each),
// cta_group::2 MMAs issued by the leader, double-buffered TMEM accumulators alternating per job, residual added by the
// tensor core (D += R * I64), SWIZZLE_128B operands, SWIZZLE_64B TMA stores.
#pragma once
#include "gemm_tcgen05.cuh"

namespace mmdx {

struct alignas(64) Gemm2Params {
  CUtensorMap tmA1;   // t2   [M, K1]  box (64, 128, 1, 1) SWIZZLE_128B
  CUtensorMap tmB1;   // W3   [N1, K1] box (64, BN / 2)
  CUtensorMap tmR;    // shortcut [M, N1] box (64, 128, 1, 1) SWIZZLE_128B
  CUtensorMap tmI;    // 64x64 identity, box (64, 32)
  CUtensorMap tmC1;   // y    [M, N1]  box (32, 128, 1, 1) SWIZZLE_64B
  CUtensorMap tmA2;   // y    [M, N1]  box (64, 128, 1, 1) SWIZZLE_128B
  CUtensorMap tmB2;   // W1n  [N2, N1] box (64, BN / 2)
  CUtensorMap tmC2;   // t1n  [M, N2]  box (32, 128, 1, 1) SWIZZLE_64B
  const float* bias1; // [N1]
  const float* bias2; // [N2]
  int kb1, nt1;       // K1 / 64, N1 / BN
  int kb2, nt2;       // N1 / 64, N2 / BN
  int num_items;      // ceil(ceil(M / 128) / 2)
  int reverse;        // zigzag traversal (see mmdx_engine::zigzag)
};

struct Gemm2Job { int type, seq, n_t, item; };

// Job sequence of one CTA pair over its n items: block b = [G1 tiles of item b (b < n)] then [G2 tiles of item b - 1 (b >= 1)]
struct Gemm2Sched {
  int n, n1, n2, b = 0, r = 0, group, num_groups, num_items, reverse;
  __host__ __device__ Gemm2Sched(int num_items_, int nt1, int nt2, int reverse_, int group_, int num_groups_)
      : n(group_ < num_items_ ? (num_items_ - group_ + num_groups_ - 1) / num_groups_ : 0), n1(nt1), n2(nt2), group(group_),
        num_groups(num_groups_), num_items(num_items_), reverse(reverse_) {}
  __device__ Gemm2Sched(const Gemm2Params& p, int group_, int num_groups_)
      : Gemm2Sched(p.num_items, p.nt1, p.nt2, p.reverse, group_, num_groups_) {}
  __host__ __device__ __forceinline__ int item_of(int seq) const {
    const int it = group + seq * num_groups;
    return reverse ? num_items - 1 - it : it;
  }
  __host__ __device__ __forceinline__ bool next(Gemm2Job& j) {
    while (b <= n) {
      if (r < n1) {
        if (b < n) { j.type = 0; j.seq = b; j.n_t = r; j.item = item_of(b); ++r; return true; }
        r = n1;
      } else if (r < n1 + n2) {
        if (b >= 1) { j.type = 1; j.seq = b - 1; j.n_t = r - n1; j.item = item_of(b - 1); ++r; return true; }
        r = n1 + n2;
      } else {
        r = 0; ++b;
      }
    }
    return false;
  }
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm2_tcgen05_kernel(const __grid_constant__ Gemm2Params p) {
  constexpr int CG = 2, BK = 64, EB = 2;
  using L = GemmSmem<BN, BK, CG, EB, 1, 0>;
  constexpr int STAGES = L::STAGES;
  constexpr uint32_t IDESC = make_idesc_bf16(256, BN);
  constexpr uint32_t IDESC64 = make_idesc_bf16(256, 64);
  constexpr uint32_t TMEM_COLS = 2 * BN <= 256 ? 256 : 512;
  constexpr int NC = BN / kEpiCW;
  static_assert(BN == 128 || BN == 256, "BN");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + L::RING_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg + L::STG_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* ident_bar = tempty_bar + 2;
  uint64_t* cfull_bar = ident_bar + 1;
  uint64_t* cempty_bar = cfull_bar + 2;
  uint64_t* ydone_bar = cempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ydone_bar + 2);
  static_assert((2 * 8 + 11) * 8 + 4 <= 512, "barrier area");
  float* sbias = reinterpret_cast<float*>(stg + L::STG_BYTES + 512);       // [2][BN]
  uint8_t* ident = stg + kEpiGroups * EB * kEpiBufBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int group = blockIdx.x / CG;
  const int num_groups = gridDim.x / CG;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.tmA1); prefetch_tensormap(&p.tmB1); prefetch_tensormap(&p.tmR); prefetch_tensormap(&p.tmI);
    prefetch_tensormap(&p.tmC1); prefetch_tensormap(&p.tmA2); prefetch_tensormap(&p.tmB2); prefetch_tensormap(&p.tmC2);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps * CG);
      mbar_init(&cfull_bar[s], 64);
      mbar_init(&cempty_bar[s], kEpiWarps);
      mbar_init(&ydone_bar[s], kEpiGroups);
    }
    mbar_init(ident_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc_pair(tmem_slot, TMEM_COLS); tmem_relinquish_pair(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer =================
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full0 = mapa_rank(smem_u32(&full_bar[0]), 0);
      if (rank == 0) mbar_arrive_expect_tx(ident_bar, kIdentBytes);
      tma_load_2d_pair(ident, &p.tmI, mapa_rank(smem_u32(ident_bar), 0), 0, rank * 32);
      Gemm2Sched sch(p, group, num_groups);
      Gemm2Job j;
      while (sch.next(j)) {
        const int row0 = (j.item * CG + rank) * 128;
        const int col0 = j.n_t * BN;
        if (j.type == 1) {
          // the y rows of this item have been stored by this CTA's epilogue (all N1 / BN tiles) and have landed
          mbar_wait(&ydone_bar[j.seq & 1], (j.seq >> 1) & 1);
          fence_proxy_async();
        }
        const int nkb = j.type == 0 ? p.kb1 : p.kb2;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          const uint32_t fb = full0 + stage * 8;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * static_cast<uint32_t>(L::A_BYTES + L::B_BYTES));
          tma_load_4d_pair(sa, j.type == 0 ? &p.tmA1 : &p.tmA2, fb, kb * BK, row0, 0, 0);
          tma_load_2d_pair(sb, j.type == 0 ? &p.tmB1 : &p.tmB2, fb, kb * BK, col0 + rank * (BN / 2));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (j.type == 0) {
          for (int q = 0; q < BN / 64; ++q) {           // shortcut tile as extra A blocks (B = resident identity)
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * L::STAGE_BYTES;
            const uint32_t fb = full0 + stage * 8;
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * static_cast<uint32_t>(L::A_BYTES));
            tma_load_4d_pair(sa, &p.tmR, fb, col0 + q * 64, row0, 0, 0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ================= MMA issuer (leader CTA) =================
      constexpr uint32_t HI = sdesc_hi<128>();
      constexpr uint32_t STAGE_STEP = L::STAGE_BYTES >> 4;
      const uint32_t a_lo0 = sdesc_lo<128>(smem_u32(smem));
      const uint32_t b_lo0 = sdesc_lo<128>(smem_u32(smem) + L::A_BYTES);
      const uint32_t id_lo = sdesc_lo<128>(smem_u32(ident));
      int stage = 0;
      uint32_t phase = 0, soff = 0;
      int it = 0;
      mbar_wait(ident_bar, 0);
      Gemm2Sched sch(p, group, num_groups);
      Gemm2Job j;
      while (sch.next(j)) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        ++it;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        const int nkb = j.type == 0 ? p.kb1 : p.kb2;
        const int nres = j.type == 0 ? BN / 64 : 0;
        uint32_t acc = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16_words<true>(tmem_d, a_lo0 + soff + 2 * k, HI, b_lo0 + soff + 2 * k, HI, IDESC, acc);
              acc = 1;
            }
            umma_commit_pair(&empty_bar[stage]);
            if (kb == nkb - 1 && nres == 0) umma_commit_pair(&tfull_bar[as]);
          }
          __syncwarp();
          acc = 1;
          soff += STAGE_STEP;
          if (++stage == STAGES) { stage = 0; phase ^= 1; soff = 0; }
        }
        for (int q = 0; q < nres; ++q) {                // D[:, 64q : 64q + 64] += R_q * I64
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_words<true>(tmem_d + q * 64, a_lo0 + soff + 2 * k, HI, id_lo + 2 * k, HI, IDESC64, 1u);
            umma_commit_pair(&empty_bar[stage]);
            if (q == nres - 1) umma_commit_pair(&tfull_bar[as]);
          }
          __syncwarp();
          soff += STAGE_STEP;
          if (++stage == STAGES) { stage = 0; phase ^= 1; soff = 0; }
        }
      }
    }
  } else if (warp < 4) {
    // ================= bias staging warps 2, 3 =================
    const int ct = threadIdx.x - 64;
    int it = 0;
    Gemm2Sched sch(p, group, num_groups);
    Gemm2Job j;
    while (sch.next(j)) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      ++it;
      const float* bias = j.type == 0 ? p.bias1 : p.bias2;
      float vb[BN / 64];
#pragma unroll
      for (int q = 0; q < BN / 64; ++q) vb[q] = __ldg(bias + j.n_t * BN + ct + 64 * q);
      mbar_wait(&cempty_bar[as], aphase ^ 1);
#pragma unroll
      for (int q = 0; q < BN / 64; ++q) sbias[as * BN + ct + 64 * q] = vb[q];
      mbar_arrive(&cfull_bar[as]);
    }
  } else {
    // ================= epilogue warps 4..19 =================
    constexpr int TW = kEpiCW / 2;
    const int e = warp - 4;
    const int g = e >> 3;
    const int hh = (e >> 2) & 1;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const bool issuer = (e & 7) == 0 && lane == 0;
    const uint32_t tempty0 = mapa_rank(smem_u32(&tempty_bar[0]), 0);
    const int sw = (r >> 1) & 3;
    int it = 0;
    uint32_t nstore = 0;
    Gemm2Sched sch(p, group, num_groups);
    Gemm2Job j;
    while (sch.next(j)) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      ++it;
      const int row0 = (j.item * CG + rank) * 128;
      const void* tmC = j.type == 0 ? static_cast<const void*>(&p.tmC1) : static_cast<const void*>(&p.tmC2);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + hh * TW;
      const float* sb = sbias + as * BN;
      mbar_wait(&cfull_bar[as], aphase);
      uint32_t v[TW];
      tmem_ld_32x16(tbase + g * kEpiCW, v);
#pragma unroll 1
      for (int c = g; c < NC; c += 2, ++nstore) {
        uint8_t* buf = stg + (g * EB + (nstore & 1)) * kEpiBufBytes;
        tmem_ld_wait();
        float x[TW];
#pragma unroll
        for (int k = 0; k < TW; ++k) x[k] = __uint_as_float(v[k]);
        if (c + 2 < NC) {
          tmem_ld_32x16(tbase + (c + 2) * kEpiCW, v);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty0 + as * 8);
        }
        const float4* b4 = reinterpret_cast<const float4*>(sb + c * kEpiCW + hh * TW);
        uint32_t pk[TW / 2];
#pragma unroll
        for (int k = 0; k < TW / 4; ++k) {
          const float4 b = b4[k];
          float y0, y1, y2, y3;
          unpack_f32x2(add_f32x2(pack_f32x2(x[4 * k], x[4 * k + 1]), pack_f32x2(b.x, b.y)), y0, y1);
          unpack_f32x2(add_f32x2(pack_f32x2(x[4 * k + 2], x[4 * k + 3]), pack_f32x2(b.z, b.w)), y2, y3);
          pk[2 * k] = pack_bf16_relu(y0, y1);            // both convolutions end in ReLU
          pk[2 * k + 1] = pack_bf16_relu(y2, y3);
        }
        uint8_t* my_row = buf + r * 64;
#pragma unroll
        for (int k = 0; k < TW / 8; ++k)
          *reinterpret_cast<uint4*>(my_row + (((hh * (TW / 8) + k) ^ sw) << 4)) =
              make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
        fence_proxy_async();
        if (issuer) tma_store_wait_read<0>();            // the store that last used the OTHER buffer has read it
        named_bar_sync(1 + g, 16 * kEpiWarps);
        if (issuer) {
          tma_store_4d(tmC, buf, j.n_t * BN + c * kEpiCW, row0, 0, 0);
          tma_store_commit();
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&cempty_bar[as]);
      if (j.type == 0 && j.n_t == p.nt1 - 1 && issuer) {
        tma_store_wait_all<0>();                         // this group's y stores of the item are in global memory
        mbar_arrive(&ydone_bar[j.seq & 1]);
      }
    }
    if (issuer) tma_store_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

}  // namespace mmdx
