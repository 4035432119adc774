// Self-attention over packed (unpadded) tokens on the 5th-gen tensor cores:
//   ctx = softmax(Q K^T / sqrt(64)) V   per (sequence, head)       (HF modeling_bert.py:115-140, BertSelfAttention)
// qkv: bf16 [T, 3*HID], columns [Q | K | V], head h at columns h*64.  Padded keys never exist here (tokens are
// packed), which equals the reference's additive -inf key-padding mask; padded query rows are never computed.
//
// Persistent kernel, one CTA per SM, 384 threads.  A work unit is (sequence, head, block of 128 query rows);
// keys/values stream in blocks of 128 (one block for L <= 128, flash-style online softmax beyond).
//   warp 0      TMA producer: per key block one stage = {Q, K, V} boxes of [128 rows x 64 cols] bf16 (SWIZZLE_128B)
//   warp 1, 2   MMA issuers, one per softmax group (single thread each):
//                 S[128 x 128] = Q K^T        4 x tcgen05.mma M=128 N=128 K=16   (A, B K-major from smem)
//                 O[128 x 64]  = P V          8 x tcgen05.mma M=128 N=64  K=16   (A = P K-major from smem,
//                                                                                  B = V MN-major: V is [key][d])
//   warp 3      TMEM allocator (2 x S + 2 x O accumulators)
//   warps 4-7 / 8-11   softmax groups 0 / 1 (units alternate between them so one group's exp/pack overlaps the
//               other's MMAs): thread = one query row; tcgen05.ld S -> row max -> p = exp2(s*scale - m) ->
//               bf16 P into swizzled smem (the A operand of the PV MMA) -> O from TMEM, rescale/accumulate in
//               registers across key blocks, normalise, 128-byte row store.
#pragma once
#include "ptx.cuh"

namespace mmdx {

constexpr int ATC_THREADS = 384;
constexpr int ATC_STAGES = 3;
constexpr int ATC_BOX = 128 * 64 * 2;              // one [128 x 64] bf16 box = 16 KB
constexpr int ATC_STAGE_BYTES = 3 * ATC_BOX;       // Q | K | V
constexpr int ATC_P_BYTES = 2 * ATC_BOX;           // P [128 x 128] bf16 as two K-major halves of 64 keys
constexpr int ATC_BAR_BYTES = 256;
constexpr int ATC_SMEM = ATC_STAGES * ATC_STAGE_BYTES + 2 * ATC_P_BYTES + ATC_BAR_BYTES + 1024;
constexpr uint32_t ATC_TMEM_COLS = 512;            // S0 | S1 (128 each) | O0 | O1 (64 each) -> 384, power of two
static_assert(ATC_SMEM <= 232448, "exceeds 227 KB of shared memory");

struct AttnParams {
  CUtensorMap tm;          // qkv as [T, 3*HID] bf16, box (64, 128), SWIZZLE_128B
  const int* cu_seqlens;   // [n_seq + 1]
  __nv_bfloat16* ctx;      // [T, HID]
  int n_seq, heads, hidden, nqb, num_units;   // nqb = query blocks per sequence at max_len
  float scale_log2;        // 1/sqrt(64) * log2(e)
};

// B operand in MN-major form (rows = K index, 128-byte rows of 64 consecutive N elements, SWIZZLE_128B):
// 8-row groups every 1024 B.  N = 64 is a single MN atom, so the MN stride is unused; both stride fields
// carry the K-group stride.
__device__ __forceinline__ uint64_t make_sdesc_mn128(uint32_t saddr) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(1024 >> 4) << 16) |
         (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

struct AttnUnit { int seq, head, qb, tok0, len; };

__device__ __forceinline__ bool attn_unit(const AttnParams& p, int id, AttnUnit& u) {
  const int per_seq = p.heads * p.nqb;
  u.seq = id / per_seq;
  const int rem = id - u.seq * per_seq;
  u.qb = rem / p.heads;
  u.head = rem - u.qb * p.heads;
  u.tok0 = __ldg(p.cu_seqlens + u.seq);
  u.len = __ldg(p.cu_seqlens + u.seq + 1) - u.tok0;
  return u.qb * 128 < u.len;
}

__global__ void __launch_bounds__(ATC_THREADS, 1) attention_tcgen05_kernel(const __grid_constant__ AttnParams p) {
  constexpr uint32_t IDESC_S = make_idesc_bf16(128, 128);
  constexpr uint32_t IDESC_PV = make_idesc_bf16(128, 64) | (1u << 16);    // B is MN-major
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* psm = smem + ATC_STAGES * ATC_STAGE_BYTES;
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(psm + 2 * ATC_P_BYTES);
  uint64_t* kv_empty = kv_full + ATC_STAGES;
  uint64_t* s_full = kv_empty + ATC_STAGES;     // [2]
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 2;
  uint64_t* o_free = o_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.tm);
    for (int s = 0; s < ATC_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 2); }   // both MMA threads release a stage
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1); mbar_init(&p_full[g], 4); mbar_init(&o_full[g], 1); mbar_init(&o_free[g], 4);
    }
    fence_barrier_init();
  }
  if (warp == 3) {
    tmem_alloc(tmem_slot, ATC_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer =================
      uint32_t blk = 0;
      for (int id = blockIdx.x; id < p.num_units; id += gridDim.x) {
        AttnUnit u;
        if (!attn_unit(p, id, u)) continue;
        const int nkb = (u.len + 127) >> 7;
        for (int kb = 0; kb < nkb; ++kb, ++blk) {
          const int stage = blk % ATC_STAGES;
          const uint32_t ph = (blk / ATC_STAGES) & 1;
          mbar_wait(&kv_empty[stage], ph ^ 1);
          uint8_t* st = smem + stage * ATC_STAGE_BYTES;
          mbar_arrive_expect_tx(&kv_full[stage], ATC_STAGE_BYTES);
          tma_load_2d(st, &p.tm, &kv_full[stage], u.head * 64, u.tok0 + u.qb * 128);
          tma_load_2d(st + ATC_BOX, &p.tm, &kv_full[stage], p.hidden + u.head * 64, u.tok0 + kb * 128);
          tma_load_2d(st + 2 * ATC_BOX, &p.tm, &kv_full[stage], 2 * p.hidden + u.head * 64, u.tok0 + kb * 128);
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    if (lane == 0) {
      // ================= MMA issuer of softmax group g =================
      const int g = warp - 1;
      const uint32_t tS = tmem_base + g * 128;
      const uint32_t tO = tmem_base + 256 + g * 64;
      const uint32_t sP = smem_u32(psm + g * ATC_P_BYTES);
      uint32_t blk = 0, n = 0;
      int nvalid = 0;
      for (int id = blockIdx.x; id < p.num_units; id += gridDim.x) {
        AttnUnit u;
        if (!attn_unit(p, id, u)) continue;
        const int nkb = (u.len + 127) >> 7;
        const bool mine = (nvalid++ & 1) == g;
        if (!mine) {
          // The ring is shared by both groups.  A parity wait is only unambiguous while the waiter is never a whole
          // phase away from the barrier, so this thread observes the other group's blocks too and co-signs their
          // release: the producer cannot recycle a stage before BOTH MMA threads have seen it filled.
          for (int kb = 0; kb < nkb; ++kb, ++blk) {
            const int stage = blk % ATC_STAGES;
            mbar_wait(&kv_full[stage], (blk / ATC_STAGES) & 1);
            mbar_arrive(&kv_empty[stage]);
          }
          continue;
        }
        for (int kb = 0; kb < nkb; ++kb, ++blk, ++n) {
          const int stage = blk % ATC_STAGES;
          const uint32_t ph = (blk / ATC_STAGES) & 1;
          mbar_wait(&kv_full[stage], ph);
          tc_fence_after();
          const uint32_t sQ = smem_u32(smem + stage * ATC_STAGE_BYTES);
          const uint32_t sK = sQ + ATC_BOX, sV = sQ + 2 * ATC_BOX;
          // S is free: this thread waited for the previous block's P (written after S was read) before its PV
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tS, make_sdesc<128>(sQ + k * 32), make_sdesc<128>(sK + k * 32), IDESC_S, k != 0 ? 1u : 0u);
          umma_commit(&s_full[g]);
          mbar_wait(&p_full[g], n & 1);
          if (n > 0) mbar_wait(&o_free[g], (n - 1) & 1);     // the group has drained the previous block's O
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tO, make_sdesc<128>(sP + (k >> 2) * ATC_BOX + (k & 3) * 32), make_sdesc_mn128(sV + k * 2048),
                      IDESC_PV, k != 0 ? 1u : 0u);
          umma_commit(&o_full[g]);
          umma_commit(&kv_empty[stage]);
        }
      }
    }
  } else if (warp >= 4) {
    // ================= softmax groups =================
    const int g = (warp - 4) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;                          // query row inside the block = TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + g * 128;
    const uint32_t tO = tmem_base + lane_base + 256 + g * 64;
    uint8_t* prow = psm + g * ATC_P_BYTES + r * 128;
    const int sw = r & 7;
    uint32_t n = 0;
    int nvalid = 0;
    for (int id = blockIdx.x; id < p.num_units; id += gridDim.x) {
      AttnUnit u;
      if (!attn_unit(p, id, u)) continue;
      if ((nvalid++ & 1) != g) continue;
      const int nkb = (u.len + 127) >> 7;
      float m = -INFINITY, l = 0.f;
      float oacc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) oacc[j] = 0.f;
      for (int kb = 0; kb < nkb; ++kb, ++n) {
        const int nv = min(128, u.len - kb * 128);        // valid keys in this block (>= 1)
        mbar_wait(&s_full[g], n & 1);
        tc_fence_after();
        // ---- pass 1: row maximum
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          if (c * 32 >= nv) break;
          uint32_t v[32];
          tmem_ld_32x32(tS + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, (c * 32 + j < nv) ? __uint_as_float(v[j]) : -INFINITY);
        }
        const float m_new = fmaxf(m, mx * p.scale_log2);
        const float corr = fast_exp2(m - m_new);               // first block: exp2(-inf) = 0
        m = m_new;
        if (kb > 0) {
          // previous block's O = P V is complete: fold it in (it was computed relative to the old maximum)
          mbar_wait(&o_full[g], (n - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tO + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) oacc[c * 32 + j] += __uint_as_float(v[j]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_free[g]);
#pragma unroll
          for (int j = 0; j < 64; ++j) oacc[j] *= corr;
        }
        l *= corr;
        // ---- pass 2: p = exp2(s*scale - m), bf16 P into the swizzled A-operand buffer
        float rs = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
          if (c * 32 < nv) {
            uint32_t v[32];
            tmem_ld_32x32(tS + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float x0 = (c * 32 + j < nv) ? __uint_as_float(v[j]) : -INFINITY;
              const float x1 = (c * 32 + j + 1 < nv) ? __uint_as_float(v[j + 1]) : -INFINITY;
              const float p0 = fast_exp2(fmaf(x0, p.scale_log2, -m_new));
              const float p1 = fast_exp2(fmaf(x1, p.scale_log2, -m_new));
              rs += p0 + p1;
              pk[j >> 1] = pack_bf16(p0, p1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = 0u;
          }
          uint8_t* half = prow + (c >> 1) * ATC_BOX;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = (c & 1) * 4 + j;                // 16-byte chunk (8 keys) inside the 64-key half
            *reinterpret_cast<uint4*>(half + ((ch ^ sw) << 4)) =
                make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          }
        }
        l += rs;
        fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
        tc_fence_before();            // orders this thread's TMEM reads of S before the next S MMA
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
      }
      // ---- last block's O, normalise, store this row
      mbar_wait(&o_full[g], (n - 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tO + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) oacc[c * 32 + j] += __uint_as_float(v[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[g]);
      const int row = u.qb * 128 + r;
      if (row < u.len) {
        const float inv = 1.0f / l;
        uint4* dst = reinterpret_cast<uint4*>(p.ctx + static_cast<size_t>(u.tok0 + row) * p.hidden + u.head * 64);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_uint4(pack_bf16(oacc[8 * j] * inv, oacc[8 * j + 1] * inv),
                              pack_bf16(oacc[8 * j + 2] * inv, oacc[8 * j + 3] * inv),
                              pack_bf16(oacc[8 * j + 4] * inv, oacc[8 * j + 5] * inv),
                              pack_bf16(oacc[8 * j + 6] * inv, oacc[8 * j + 7] * inv));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATC_TMEM_COLS);
  }
}

}  // namespace mmdx
