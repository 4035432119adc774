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
  int reverse;             // 1: walk the units from the last to the first (see mmdx_engine::zigzag)
  // Folded LayerNorm (gemm_tcgen05.cuh): when set, row t of ctx is written as ctx[t] / rstd[t], rstd from the row sums of
  // the pre-LayerNorm tensor the NEXT GEMM re-creates its residual from (it multiplies the whole accumulator by rstd).
  const long long* row_stats;   // [T][2] 2^24 fixed-point (sum x, sum x^2), or null
  float inv_n, eps;
};

__device__ __forceinline__ float attn_row_sigma(const AttnParams& p, int tok) {
  if (p.row_stats == nullptr) return 1.0f;
  const float s1 = static_cast<float>(static_cast<double>(p.row_stats[2 * static_cast<size_t>(tok)]) * (1.0 / 16777216.0));
  const float s2 = static_cast<float>(static_cast<double>(p.row_stats[2 * static_cast<size_t>(tok) + 1]) * (1.0 / 16777216.0));
  const float mu = s1 * p.inv_n;
  const float var = fmaxf(s2 * p.inv_n - mu * mu, 0.0f);
  return 1.0f / rsqrtf(var + p.eps);        // exactly the reciprocal of the rstd the consumer's epilogue computes
}

// B operand in MN-major form (rows = K index, 128-byte rows of 64 consecutive N elements, SWIZZLE_128B):
// 8-row groups every 1024 B.  N = 64 is a single MN atom, so the MN stride is unused; both stride fields
// carry the K-group stride.
__device__ __forceinline__ uint64_t make_sdesc_mn128(uint32_t saddr) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(1024 >> 4) << 16) |
         (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

struct AttnUnit { int seq, head, qb, tok0, len; };

__device__ __forceinline__ bool attn_unit(const AttnParams& p, int id, AttnUnit& u) {
  if (p.reverse) id = p.num_units - 1 - id;
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
  pdl_wait();
  pdl_trigger();

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
        const float inv = attn_row_sigma(p, u.tok0 + row) / l;
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


// =================================================================================================================
// Short-sequence variant (every sequence <= 128 tokens: the reference's own max_len = 96, and BASELINE C2's 128).
// One unit = (sequence, head): a single 128-key block, so no online softmax.  Four units are in flight per SM:
// unit i uses smem stage i%4 and TMEM region i%4 (128 columns), in which S, P and O alias each other:
//     S = Q K^T   fp32, columns [0,128)                      tcgen05.mma, A/B K-major from smem
//     P = softmax numerators, bf16 pairs, columns [0,64)     written in place by tcgen05.st as S is consumed
//     O = P V     fp32, columns [64,128)                     tcgen05.mma with A = P read from TENSOR MEMORY
// so the probabilities never touch shared memory.  Roles (640 threads): warp 0 TMA producer, warp 1 issues the S MMAs,
// warp 3 the P V MMAs (independently, each in unit order), warp 2 TMEM allocator, warps 4-19 four softmax groups
// (thread = query row).  Warps whose 32 rows are all padding skip the math and only keep the barriers moving;
// key chunks beyond the sequence are neither exponentiated nor multiplied.
// =================================================================================================================
constexpr int ATS_THREADS = 640;
constexpr int ATS_GROUPS = 4;
constexpr int ATS_SMEM = ATS_GROUPS * ATC_STAGE_BYTES + ATC_BAR_BYTES + 1024;

struct AttnShortUnit { int tok0, len, head; };

__device__ __forceinline__ AttnShortUnit attn_short_unit(const AttnParams& p, int id) {
  AttnShortUnit u;
  if (p.reverse) id = p.n_seq * p.heads - 1 - id;
  const int seq = id / p.heads;
  u.head = id - seq * p.heads;
  u.tok0 = __ldg(p.cu_seqlens + seq);
  u.len = __ldg(p.cu_seqlens + seq + 1) - u.tok0;
  return u;
}

__global__ void __launch_bounds__(ATS_THREADS, 1) attention_short_tcgen05_kernel(const __grid_constant__ AttnParams p) {
  constexpr uint32_t IDESC_S = make_idesc_bf16(128, 128);
  constexpr uint32_t IDESC_PV = make_idesc_bf16(128, 64) | (1u << 16);    // B (= V) is MN-major
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(smem + ATS_GROUPS * ATC_STAGE_BYTES);
  uint64_t* kv_empty = kv_full + ATS_GROUPS;
  uint64_t* s_full = kv_empty + ATS_GROUPS;
  uint64_t* p_full = s_full + ATS_GROUPS;
  uint64_t* o_full = p_full + ATS_GROUPS;
  uint64_t* o_free = o_full + ATS_GROUPS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + ATS_GROUPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_units = p.n_seq * p.heads;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.tm);
    for (int g = 0; g < ATS_GROUPS; ++g) {
      mbar_init(&kv_full[g], 1); mbar_init(&kv_empty[g], 1); mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 4); mbar_init(&o_full[g], 1); mbar_init(&o_free[g], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
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
      // ================= TMA producer =================
      int i = 0;
      for (int id = blockIdx.x; id < num_units; id += gridDim.x, ++i) {
        const AttnShortUnit u = attn_short_unit(p, id);
        const int g = i & 3;
        mbar_wait(&kv_empty[g], ((i >> 2) & 1) ^ 1);
        uint8_t* st = smem + g * ATC_STAGE_BYTES;
        mbar_arrive_expect_tx(&kv_full[g], ATC_STAGE_BYTES);
        tma_load_2d(st, &p.tm, &kv_full[g], u.head * 64, u.tok0);
        tma_load_2d(st + ATC_BOX, &p.tm, &kv_full[g], p.hidden + u.head * 64, u.tok0);
        tma_load_2d(st + 2 * ATC_BOX, &p.tm, &kv_full[g], 2 * p.hidden + u.head * 64, u.tok0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer 1: S(i) = Q K^T into region i%4 =================
      // S and P V are issued by two different warps: a single in-order issuer would sit on the next unit's K/V load
      // while a finished P waits for its P V (ncu: 39 % of all samples were softmax warps waiting for o_full).
      const int n_mine = blockIdx.x < num_units ? (num_units - 1 - blockIdx.x) / static_cast<int>(gridDim.x) + 1 : 0;
      for (int i = 0; i < n_mine; ++i) {
        const int g = i & 3;
        const uint32_t ph = (i >> 2) & 1;
        mbar_wait(&kv_full[g], ph);
        if (i >= ATS_GROUPS) mbar_wait(&o_free[g], ph ^ 1);      // unit i-4's O has been read out of this region
        tc_fence_after();
        const uint32_t sQ = smem_u32(smem + g * ATC_STAGE_BYTES);
        const uint32_t sK = sQ + ATC_BOX;
        const uint32_t tS = tmem_base + g * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tS, make_sdesc<128>(sQ + k * 32), make_sdesc<128>(sK + k * 32), IDESC_S, k != 0 ? 1u : 0u);
        umma_commit(&s_full[g]);
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      // ================= MMA issuer 2: O(j) = P V, P from tensor memory =================
      const int n_mine = blockIdx.x < num_units ? (num_units - 1 - blockIdx.x) / static_cast<int>(gridDim.x) + 1 : 0;
      for (int j = 0; j < n_mine; ++j) {
        const int g = j & 3;
        const int id = blockIdx.x + j * static_cast<int>(gridDim.x);
        const AttnShortUnit u = attn_short_unit(p, id);
        const int nks = (min(u.len, 128) + 15) >> 4;             // 16-key steps that hold any valid key
        mbar_wait(&p_full[g], (j >> 2) & 1);
        tc_fence_after();
        const uint32_t sV = smem_u32(smem + g * ATC_STAGE_BYTES) + 2 * ATC_BOX;
        const uint32_t tP = tmem_base + g * 128;
        const uint32_t tO = tP + 64;
        for (int k = 0; k < nks; ++k)
          umma_bf16_ts(tO, tP + k * 8, make_sdesc_mn128(sV + k * 2048), IDESC_PV, k != 0 ? 1u : 0u);
        umma_commit(&o_full[g]);
        umma_commit(&kv_empty[g]);                               // S of this unit completed long ago: stage is free
      }
    }
  } else if (warp >= 4) {
    // ================= softmax groups =================
    const int g = (warp - 4) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t tS = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 128;
    const uint32_t tO = tS + 64;
    const int stride = ATS_GROUPS * static_cast<int>(gridDim.x);
    int id = blockIdx.x + g * static_cast<int>(gridDim.x);
    AttnShortUnit u{0, 0, 0};
    if (id < num_units) u = attn_short_unit(p, id);
    for (uint32_t n = 0; id < num_units; id += stride, ++n) {
      AttnShortUnit nxt{0, 0, 0};
      if (id + stride < num_units) nxt = attn_short_unit(p, id + stride);     // prefetch: off the critical path
      const int len = min(u.len, 128);
      const bool active = q * 32 < len;               // warp-uniform: any valid query row in this quarter
      // folded LayerNorm: this row's 1 / rstd is requested before waiting for S (its global loads hide under the wait)
      const float sigma = (active && r < len) ? attn_row_sigma(p, u.tok0 + r) : 1.0f;
      mbar_wait(&s_full[g], n & 1);
      tc_fence_after();
      float l = 0.f;
      if (active) {
        const int nch = (len + 31) >> 5;              // 32-key chunks holding valid keys
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tS + c * 32, v);
          tmem_ld_wait();
          if (c * 32 + 32 <= len) {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, (c * 32 + j < len) ? __uint_as_float(v[j]) : -INFINITY);
          }
        }
        const float msc = mx * p.scale_log2;
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
          uint32_t v[32];
          uint32_t pk[16];
          tmem_ld_32x32(tS + c * 32, v);
          tmem_ld_wait();
          if (c * 32 + 32 <= len) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float p0 = fast_exp2(fmaf(__uint_as_float(v[j]), p.scale_log2, -msc));
              const float p1 = fast_exp2(fmaf(__uint_as_float(v[j + 1]), p.scale_log2, -msc));
              l += p0 + p1;
              pk[j >> 1] = pack_bf16(p0, p1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float p0 = (c * 32 + j < len) ? fast_exp2(fmaf(__uint_as_float(v[j]), p.scale_log2, -msc)) : 0.f;
              const float p1 = (c * 32 + j + 1 < len) ? fast_exp2(fmaf(__uint_as_float(v[j + 1]), p.scale_log2, -msc)) : 0.f;
              l += p0 + p1;
              pk[j >> 1] = pack_bf16(p0, p1);
            }
          }
          // keys [32c, 32c+32) -> packed columns [16c, 16c+16): inside S columns this thread has already consumed
          tmem_st_32x16(tS + c * 16, pk);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      // ---- O = P V
      mbar_wait(&o_full[g], n & 1);
      tc_fence_after();
      if (active) {
        const float inv = sigma / l;
        uint4* dst = reinterpret_cast<uint4*>(p.ctx + static_cast<size_t>(u.tok0 + r) * p.hidden + u.head * 64);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tO + c * 32, v);
          tmem_ld_wait();
          if (r < len) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dst[c * 4 + j] = make_uint4(
                  pack_bf16(__uint_as_float(v[8 * j]) * inv, __uint_as_float(v[8 * j + 1]) * inv),
                  pack_bf16(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv),
                  pack_bf16(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv),
                  pack_bf16(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[g]);
      u = nxt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace mmdx
