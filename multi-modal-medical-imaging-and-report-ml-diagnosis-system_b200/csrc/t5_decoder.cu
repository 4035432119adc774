// KV-cached T5 decoder step: the model call inside the reference's report generation
// (FusionTransformerModel.generate -> HF T5ForConditionalGeneration.generate, training_pipeline.py:613-618;
// inference_pipeline.py:190-196) as hand-written CUDA kernels behind a C ABI (include/mmdx.h, mmdx_t5_*).
//
// One step = one new token for each of R = batch x beams rows.  With R <= 64 every contraction is a skinny GEMV-like
// product whose cost is reading the weights once (t5-small decoder + tied LM head: 41 M parameters, 165 MB fp32 per
// step), so this is HBM/latency-bound work on the CUDA cores: fp32 weights, fp32 arithmetic (the beam search compares sums
// of log-probabilities; fp32 keeps the token sequence identical to HF's eager fp32 path):
//   x = E[token]
//   per block:  qkv = RMSNorm(x) [Wq;Wk;Wv]^T ; self-attention over the cache (+ relative-position bias, no 1/sqrt(d)) ;
//               x += attn Wo^T ; q = RMSNorm(x) Wq^T ; cross-attention over the projected conditioning tokens ;
//               x += attn Wo^T ; x += relu(RMSNorm(x) Wi^T) Wo^T
//   logits = (RMSNorm(x) * d_model^-0.5) E^T            (tied embeddings; lm_head.weight when untied)
// The whole step is ONE cooperative launch (t5_step_mega_kernel: one 512-thread CTA per SM, the 50 dependent phases
// separated by grid-wide barriers instead of kernel boundaries): at 8 rows a phase is a few microseconds of work, so 50
// launches cost ~1 ms per token where the phases themselves need ~0.15 ms.  The self-attention cache is never moved by
// the beam search: a hypothesis keeps, per position, the row slot that wrote that position (its ancestry), and
// mmdx_t5_reorder permutes those small index rows instead of copying keys and values.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mmdx.h"

namespace {

thread_local std::string g_t5_err;
int t5_fail(const std::string& m) { g_t5_err = m; return 1; }
#define T5_CK(call)                                                                                             \
  do {                                                                                                          \
    cudaError_t _e = (call);                                                                                    \
    if (_e != cudaSuccess) return t5_fail(std::string(#call) + " failed: " + cudaGetErrorString(_e));           \
  } while (0)
#define T5_REQUIRE(cond, msg)                                                \
  do {                                                                       \
    if (!(cond)) return t5_fail(std::string("mmdx_t5: ") + msg);             \
  } while (0)

constexpr int kRowChunk = 8;          // rows of x a linear launch keeps in shared memory
constexpr int kColsPerWarp = 4;

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[r, n] (+= res) = act( (RMSNorm(x[r]) * in_scale) . W[n] ),  W [N, K] row-major, no bias (T5 Linear layers have none).
// ln_w == null: x is used as is.  Rows are staged (normalised) in shared memory once per block; every warp then streams
// kColsPerWarp weight rows with 16-byte loads and keeps kRowChunk accumulators per row of W.
__global__ void __launch_bounds__(256) t5_linear_kernel(const float* __restrict__ x, long long ldx,
                                                        const float* __restrict__ ln_w, float eps, float in_scale,
                                                        const float* __restrict__ W, float* __restrict__ out, long long ldo,
                                                        const float* __restrict__ res, int R, int K, int N, int relu) {
  extern __shared__ float xs[];       // [kRowChunk][K]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.y * kRowChunk;
  const int nr = min(kRowChunk, R - r0);
  for (int r = warp; r < kRowChunk; r += 8) {
    float* dst = xs + static_cast<size_t>(r) * K;
    if (r >= nr) { for (int k = lane; k < K; k += 32) dst[k] = 0.f; continue; }
    const float* src = x + static_cast<size_t>(r0 + r) * ldx;
    float sc = in_scale;
    if (ln_w != nullptr) {            // T5LayerNorm: x * rsqrt(mean(x^2) + eps) * w  (no mean subtraction, no bias)
      float q = 0.f;
      for (int k = lane; k < K; k += 32) q = fmaf(src[k], src[k], q);
      sc *= rsqrtf(warp_sum_f(q) / K + eps);
    }
    for (int k = lane; k < K; k += 32) dst[k] = src[k] * sc * (ln_w != nullptr ? ln_w[k] : 1.0f);
  }
  __syncthreads();
  const int c0 = (blockIdx.x * 8 + warp) * kColsPerWarp;
#pragma unroll 1
  for (int c = c0; c < min(c0 + kColsPerWarp, N); ++c) {
    const float4* w4 = reinterpret_cast<const float4*>(W + static_cast<size_t>(c) * K);
    float acc[kRowChunk];
#pragma unroll
    for (int r = 0; r < kRowChunk; ++r) acc[r] = 0.f;
    for (int k4 = lane; k4 < K / 4; k4 += 32) {
      const float4 w = __ldg(w4 + k4);
#pragma unroll
      for (int r = 0; r < kRowChunk; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(xs + static_cast<size_t>(r) * K + 4 * k4);
        acc[r] = fmaf(w.x, v.x, fmaf(w.y, v.y, fmaf(w.z, v.z, fmaf(w.w, v.w, acc[r]))));
      }
    }
#pragma unroll
    for (int r = 0; r < kRowChunk; ++r) acc[r] = warp_sum_f(acc[r]);
    if (lane == 0) {
      for (int r = 0; r < nr; ++r) {
        float v = acc[r];
        if (relu) v = fmaxf(v, 0.f);
        const size_t o = static_cast<size_t>(r0 + r) * ldo + c;
        if (res != nullptr) v += res[o];
        out[o] = v;
      }
    }
  }
}

// ---- the decoder step as one cooperative kernel ---------------------------------------------------------------------
namespace cg = cooperative_groups;
constexpr int kMegaThreads = 512, kMegaWarps = kMegaThreads / 32;

struct T5Layer {                      // device copy of one block's pointers
  const float *ln0, *qkv, *so, *ln1, *cq, *co, *ln2, *wi, *wo;
  float *sk, *sv;                     // self-attention cache [R, H, Tmax, dk]: slot (r, j) is written once, at step j, by row r
  const float *ck, *cv;               // projected conditioning tokens [R, H, n_enc, dk]
  const float *kq, *w2;               // cross-attention folded per generation (t5_cross_fold_kernel), or null
};

constexpr int kMaxT5Layers = 24;
// Kernel parameters live in the constant bank: nothing here is reached through a pointer in local or global memory, because
// every grid barrier flushes L1 and each such access would then cost an L2 round trip on the critical path of a phase.
struct T5StepArgs {
  const int32_t* tok; float* logits;
  const float *E, *lm, *final_ln, *bias;
  T5Layer layers[kMaxT5Layers];
  int32_t* anc;                       // [R, Tmax]: anc[r][j] = the row slot that holds position j of row r's hypothesis
  const int32_t* anc_src;             // with bidx: the beam reordering happens here (anc[r] = anc_src[bidx[r]]), else unused
  const int32_t* bidx;
  const int32_t* copy_src; int32_t* copy_dst; int copy_n;     // optional: words the last CTA copies (mapped host -> device) for the scoring kernels
  float *x, *qkv, *att, *q, *hid;
  float eps, lm_scale;
  int R, d, H, dk, ff, L, vocab, Tmax, t, n_enc, kmax;
  int xfold, xk;                      // cross-attention folded: xfold = R * H * n_enc score columns, xk = xfold padded to 128
  unsigned long long* prof;           // optional: globaltimer stamp per phase boundary, written by CTA 0 (mmdx_t5_step_profile)
};

__device__ __forceinline__ float dot4(const float4 w, const float4 v, float acc) {
  return fmaf(w.x, v.x, fmaf(w.y, v.y, fmaf(w.z, v.z, fmaf(w.w, v.w, acc))));
}

// One linear layer spread over the whole grid: out[r, c] (+= res) = act((RMSNorm(x[r]) * in_scale) . W[c]).  The NR rows of
// x are staged (normalised) in every CTA's shared memory; a warp owns one column (two when there are more columns than
// 2 x warps: the x reads from shared memory are then shared by both, which keeps the 65 MB LM head on the HBM side of
// the shared-memory roofline), or a quarter of a column's K range when there are fewer columns than a quarter of the
// warps (the four partial sums meet in shared memory, in a fixed order).  Columns are dealt round-robin over the CTAs so
// that a narrow layer still streams through every SM.
// The grid barrier that separates this phase from its producer sits INSIDE, behind the first batch of weight loads: the
// weights do not depend on the previous phase, so their HBM / L2 latency runs under the barrier and the staging of x.
template <int NR>
__device__ __noinline__ void mega_linear(const int R, const float eps, const float* x, long long ldx,
                                         const float* __restrict__ ln_w, float in_scale, const float* __restrict__ W, float* out,
                                         long long ldo, const float* res, int K, int N, int relu, bool stream_w, float* xs,
                                         float* part, int bd_m = 0, int bd_nenc = 0) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nb = gridDim.x, bid = blockIdx.x;
  const int GW = nb * kMegaWarps;
  const int KS = (N * 4 <= GW && K % 512 == 0) ? 4 : 1;
  const int NC = N >= 2 * GW ? 2 : 1;
  const int slots = kMegaWarps / KS, slot = warp / KS, kp = warp % KS;
  const int kq = K / 4 / KS;                                     // float4 per warp task (a multiple of 32)
  const int groups = (N + slots * NC - 1) / (slots * NC);
  float4 wa[4], wb[4];
  auto load_batch = [&](int c0, int i0) {                        // four independent 16-byte loads per column in flight
    const float4* w0 = reinterpret_cast<const float4*>(W + static_cast<size_t>(c0) * K) + kp * kq;
    const bool two = NC == 2 && c0 + 1 < N;
    const float4* w1 = w0 + (two ? K / 4 : 0);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + 32 * u;
      if (i < kq) {
        wa[u] = stream_w ? __ldcs(w0 + i) : __ldg(w0 + i);
        if (two) wb[u] = stream_w ? __ldcs(w1 + i) : __ldg(w1 + i);
      }
    }
  };
  bool preloaded = false;
  if (bid < groups && (bid * slots + slot) * NC < N) { load_batch((bid * slots + slot) * NC, lane); preloaded = true; }
  constexpr int XR = 4;                                          // float4 per lane of a row kept in registers (K <= 512)
  float4 gw[XR];                                                 // the norm weights are constants too
  const bool in_regs = ln_w != nullptr && K <= XR * 128;
  if (in_regs) {
#pragma unroll
    for (int u = 0; u < XR; ++u) if (lane + 32 * u < K / 4) gw[u] = __ldg(reinterpret_cast<const float4*>(ln_w) + lane + 32 * u);
  }
  cg::this_grid().sync();
  for (int r0 = 0; r0 < R; r0 += NR) {
    __syncthreads();                                             // the previous readers of xs are done
    if (bd_m > 0) {
      // folded cross-attention, second half: x holds the scores S [R, R * bd_m] of every row against every row's keys
      // (only the diagonal blocks mean anything); the staged operand is the block-diagonal matrix of softmax(S) per
      // head - row r carries its probabilities in columns [r * bd_m, (r + 1) * bd_m) and zeros elsewhere - so the
      // product with W2 [N, K] (values x output projection, per row) is that row's cross-attention output.
      for (int i = tid; i < NR * K; i += kMegaThreads) {
        const int r = i / K, col = i - r * K, row = r0 + r;
        float v = 0.f;
        if (row < R && col >= row * bd_m && col < (row + 1) * bd_m) {
          const int m = col - row * bd_m, h = m / bd_nenc;
          const float* sc = x + static_cast<size_t>(row) * ldx + row * bd_m + h * bd_nenc;
          float mx = -INFINITY, sum = 0.f;
          for (int j = 0; j < bd_nenc; ++j) mx = fmaxf(mx, sc[j]);
          for (int j = 0; j < bd_nenc; ++j) sum += expf(sc[j] - mx);
          v = expf(sc[m - h * bd_nenc] - mx) / sum;
        }
        xs[i] = v;
      }
    } else if (ln_w == nullptr) {                                // plain copy (x in_scale): every thread, independent 16-byte loads
      const int k4 = K / 4, tot = NR * k4;
#pragma unroll 4
      for (int i = tid; i < tot; i += kMegaThreads) {
        const int r = i / k4, k = i - r * k4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + r < R) v = reinterpret_cast<const float4*>(x + static_cast<size_t>(r0 + r) * ldx)[k];
        v.x *= in_scale; v.y *= in_scale; v.z *= in_scale; v.w *= in_scale;
        reinterpret_cast<float4*>(xs)[i] = v;
      }
    }
    for (int r = warp; r < NR && ln_w != nullptr; r += kMegaWarps) {
      float4* dst = reinterpret_cast<float4*>(xs + static_cast<size_t>(r) * K);
      if (r0 + r >= R) { for (int k = lane; k < K / 4; k += 32) dst[k] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
      const float4* src = reinterpret_cast<const float4*>(x + static_cast<size_t>(r0 + r) * ldx);
      if (in_regs) {                  // T5LayerNorm: x * rsqrt(mean(x^2) + eps) * w  (no mean subtraction, no bias); one read of x
        float4 xv4[XR];
        float q = 0.f;
#pragma unroll
        for (int u = 0; u < XR; ++u) if (lane + 32 * u < K / 4) xv4[u] = src[lane + 32 * u];
#pragma unroll
        for (int u = 0; u < XR; ++u) if (lane + 32 * u < K / 4) q = dot4(xv4[u], xv4[u], q);
        const float sc = in_scale * rsqrtf(warp_sum_f(q) / K + eps);
#pragma unroll
        for (int u = 0; u < XR; ++u) {
          if (lane + 32 * u < K / 4) {
            float4 v = xv4[u];
            v.x = v.x * sc * gw[u].x; v.y = v.y * sc * gw[u].y; v.z = v.z * sc * gw[u].z; v.w = v.w * sc * gw[u].w;
            dst[lane + 32 * u] = v;
          }
        }
        continue;
      }
      float sc = in_scale;
      if (ln_w != nullptr) {
        float q = 0.f;
        for (int k = lane; k < K / 4; k += 32) { const float4 v = src[k]; q = dot4(v, v, q); }
        sc *= rsqrtf(warp_sum_f(q) / K + eps);
      }
      for (int k = lane; k < K / 4; k += 32) {
        float4 v = src[k];
        const float4 g = ln_w != nullptr ? reinterpret_cast<const float4*>(ln_w)[k] : make_float4(1.f, 1.f, 1.f, 1.f);
        v.x = v.x * sc * g.x; v.y = v.y * sc * g.y; v.z = v.z * sc * g.z; v.w = v.w * sc * g.w;
        dst[k] = v;
      }
    }
    __syncthreads();
    for (int g = bid; g < groups; g += nb) {                     // trip count is uniform over the CTA
      const int c0 = (g * slots + slot) * NC;
      const bool on = c0 < N, two = NC == 2 && c0 + 1 < N;
      const int oc = c0 + lane / NR, orow = r0 + lane % NR;       // the output element this lane writes (kp == 0 warps)
      const bool writes = on && kp == 0 && lane < NR * NC && oc < N && orow < R;
      float resv = 0.f;
      if (writes && res != nullptr) resv = res[static_cast<size_t>(orow) * ldo + oc];     // in flight under the products
      float acc0[NR], acc1[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) { acc0[r] = 0.f; acc1[r] = 0.f; }
      if (stream_w && g + nb < groups) {                        // LM head: pull the next task's rows from HBM into L2 now
        const int cn = ((g + nb) * slots + slot) * NC;
        if (cn < N) {
          const float4* n0 = reinterpret_cast<const float4*>(W + static_cast<size_t>(cn) * K) + kp * kq;
          for (int i = lane * 8; i < kq * (NC == 2 && cn + 1 < N ? 2 : 1); i += 256)      // one prefetch per 128-byte line
            asm volatile("prefetch.global.L2 [%0];" ::"l"(n0 + i));
        }
      }
      if (on) {
        const float4* xv = reinterpret_cast<const float4*>(xs) + kp * kq;
        for (int i0 = lane; i0 < kq; i0 += 128) {
          if (!(preloaded && i0 == lane)) load_batch(c0, i0);
          preloaded = false;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + 32 * u;
            if (i < kq) {
#pragma unroll
              for (int r = 0; r < NR; ++r) {
                const float4 v = xv[static_cast<size_t>(r) * (K / 4) + i];
                acc0[r] = dot4(wa[u], v, acc0[r]);
                if (two) acc1[r] = dot4(wb[u], v, acc1[r]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < NR; ++r) { acc0[r] = warp_sum_f(acc0[r]); if (NC == 2) acc1[r] = warp_sum_f(acc1[r]); }
      float v = 0.f;                                             // lane l holds (column l / NR, row l % NR)
#pragma unroll
      for (int r = 0; r < NR; ++r) { if (lane == r) v = acc0[r]; if (lane == NR + r) v = acc1[r]; }
      if (KS > 1) {
        if (lane < NR) part[warp * NR + lane] = v;
        __syncthreads();
        if (kp == 0 && lane < NR) {
          v = part[warp * NR + lane];
          for (int q = 1; q < KS; ++q) v += part[(warp + q) * NR + lane];
        }
        __syncthreads();
      }
      if (writes) {
        if (relu) v = fmaxf(v, 0.f);
        out[static_cast<size_t>(orow) * ldo + oc] = v + resv;
      }
    }
  }
}

// Self-attention of the new token against the cache, one (row, head) per CTA.  qkv [R, 3*inner] (q | k | v); position t is
// written into the row's own slot first; key j < t is read from slot anc[r][j].  scores[j] = q . k_j + bias[(t - j) * H + h]
// (no scaling in T5), softmax in fp32, out [R, inner].
__device__ __forceinline__ void mega_self_attn(const T5StepArgs& a, const T5Layer& L, float* sm) {
  float* q = sm;                      // [64]
  float* red = sm + 64;               // [32]
  float* part = sm + 96;              // [32 key groups][64]
  float* p = sm + 96 + 2048;          // [Tmax]
  int* slot = reinterpret_cast<int*>(p + a.Tmax);      // [Tmax]: row slot of every position of this hypothesis
  const int tid = threadIdx.x, H = a.H, dk = a.dk, Tmax = a.Tmax, t = a.t, inner = H * dk;
  for (int item = blockIdx.x; item < a.R * H; item += gridDim.x) {
    const int r = item / H, h = item - r * H;
    __syncthreads();
    const float* row = a.qkv + static_cast<size_t>(r) * 3 * inner;
    const int32_t* an = a.anc + static_cast<size_t>(r) * Tmax;
    for (int j = tid; j <= t; j += kMegaThreads) slot[j] = j == t ? r : an[j];
    if (tid < dk) {
      const size_t own = ((static_cast<size_t>(r) * H + h) * Tmax + t) * dk + tid;
      q[tid] = row[h * dk + tid];
      L.sk[own] = row[inner + h * dk + tid];
      L.sv[own] = row[2 * inner + h * dk + tid];
    }
    if (h == 0 && tid == 0) a.anc[static_cast<size_t>(r) * Tmax + t] = r;
    __syncthreads();
    float mx = -INFINITY;
    for (int j = tid; j <= t; j += kMegaThreads) {
      const float4* kj = reinterpret_cast<const float4*>(L.sk + ((static_cast<size_t>(slot[j]) * H + h) * Tmax + j) * dk);
      const float4* q4 = reinterpret_cast<const float4*>(q);
      float s = 0.f;
#pragma unroll 4
      for (int d = 0; d < dk / 4; ++d) s = dot4(q4[d], kj[d], s);
      s += a.bias[static_cast<size_t>(t - j) * H + h];
      p[j] = s;
      mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < kMegaWarps; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j <= t; j += kMegaThreads) { const float e = expf(p[j] - mx); p[j] = e; sum += e; }
    sum = warp_sum_f(sum);
    if ((tid & 31) == 0) red[tid >> 5] = sum;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < kMegaWarps; ++w) tot += red[w];
    const float inv = 1.0f / tot;
    // P V: 32 groups of 16 threads (four dims each, 16-byte loads) take every 32nd key; partial sums meet in shared memory
    const int d4 = tid & 15, g = tid >> 4;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d4 * 4 < dk) {
#pragma unroll 4
      for (int j = g; j <= t; j += 32) {
        const float4 v = *reinterpret_cast<const float4*>(L.sv + ((static_cast<size_t>(slot[j]) * H + h) * Tmax + j) * dk + d4 * 4);
        const float w = p[j] * inv;
        o.x = fmaf(w, v.x, o.x); o.y = fmaf(w, v.y, o.y); o.z = fmaf(w, v.z, o.z); o.w = fmaf(w, v.w, o.w);
      }
    }
    reinterpret_cast<float4*>(part)[g * 16 + d4] = o;
    __syncthreads();
    if (tid < dk) {
      float v = part[tid];
      for (int q8 = 1; q8 < 32; ++q8) v += part[q8 * 64 + tid];
      a.att[static_cast<size_t>(r) * inner + h * dk + tid] = v;
    }
  }
}

// Cross-attention of the new token against the projected conditioning tokens: q [R, inner], K / V [R, H, n_enc, dk].
__device__ __forceinline__ void mega_cross_attn(const T5StepArgs& a, const T5Layer& L, float* sm) {
  float* p = sm;                      // [64]
  const int tid = threadIdx.x, H = a.H, dk = a.dk, n_enc = a.n_enc, inner = H * dk;
  for (int item = blockIdx.x; item < a.R * H; item += gridDim.x) {
    const int r = item / H, h = item - r * H;
    __syncthreads();
    const float4* q4 = reinterpret_cast<const float4*>(a.q + static_cast<size_t>(r) * inner + h * dk);
    const float* kc = L.ck + (static_cast<size_t>(r) * H + h) * n_enc * dk;
    const float* vc = L.cv + (static_cast<size_t>(r) * H + h) * n_enc * dk;
    if (tid < n_enc) {
      const float4* kj = reinterpret_cast<const float4*>(kc + static_cast<size_t>(tid) * dk);
      float s = 0.f;
      for (int d = 0; d < dk / 4; ++d) s = dot4(q4[d], kj[d], s);
      p[tid] = s;
    }
    __syncthreads();
    if (tid < dk) {
      float mx = -INFINITY, sum = 0.f;
      for (int j = 0; j < n_enc; ++j) mx = fmaxf(mx, p[j]);
      for (int j = 0; j < n_enc; ++j) sum += expf(p[j] - mx);
      float o = 0.f;
      for (int j = 0; j < n_enc; ++j) o = fmaf(expf(p[j] - mx) / sum, vc[static_cast<size_t>(j) * dk + tid], o);
      a.att[static_cast<size_t>(r) * inner + h * dk + tid] = o;
    }
  }
}

template <int NR>
__global__ void __launch_bounds__(kMegaThreads, 1) t5_step_mega_kernel(const T5StepArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ float sm[];       // linear phases: xs [NR][kmax] | part [warps][NR]; attention phases: see there
  float* xs = sm;
  float* part = sm + static_cast<size_t>(NR) * a.kmax;
  const int d = a.d, inner = a.H * a.dk;
  int np = 0;
  auto stamp = [&]() {
    if (a.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long tns;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
      a.prof[np] = tns;
    }
    ++np;
  };
  stamp();
  // tok / bidx / copy_src may live in mapped host memory (mmdx_t5_generate): one PCIe read per warp or row, not per thread
  if (a.bidx != nullptr) {            // beam reordering of the last search step; read two barriers later by layer 0's attention
    __shared__ int s_src;
    for (int r = blockIdx.x; r < a.R; r += gridDim.x) {
      if (threadIdx.x == 0) s_src = min(max(a.bidx[r], 0), a.R - 1);
      __syncthreads();
      for (int j = threadIdx.x; j < a.t; j += kMegaThreads)
        a.anc[static_cast<size_t>(r) * a.Tmax + j] = a.anc_src[static_cast<size_t>(s_src) * a.Tmax + j];
      __syncthreads();
    }
  }
  if (a.copy_n > 0 && blockIdx.x == gridDim.x - 1)
    for (int i = threadIdx.x; i < a.copy_n; i += kMegaThreads) a.copy_dst[i] = a.copy_src[i];
  for (int i = blockIdx.x * kMegaThreads + threadIdx.x; i < a.R * d; i += gridDim.x * kMegaThreads) {
    const int r = i / d, k = i - r * d;                          // d % 128 == 0: a warp stays inside one row
    int id = 0;
    if ((threadIdx.x & 31) == 0) id = min(max(a.tok[r], 0), a.vocab - 1);
    id = __shfl_sync(0xffffffffu, id, 0);
    a.x[i] = a.E[static_cast<size_t>(id) * d + k];
  }
  for (int l = 0; l < a.L; ++l) {                                // every mega_linear starts with the barrier behind its producer
    const T5Layer& L = a.layers[l];
    mega_linear<NR>(a.R, a.eps, a.x, d, L.ln0, 1.0f, L.qkv, a.qkv, 3 * inner, nullptr, d, 3 * inner, 0, false, xs, part);
    stamp();
    grid.sync();
    mega_self_attn(a, L, sm);
    stamp();
    mega_linear<NR>(a.R, a.eps, a.att, inner, nullptr, 1.0f, L.so, a.x, d, a.x, inner, d, 0, false, xs, part);
    stamp();
    if (a.xfold > 0) {
      // cross-attention with the query projection folded into the keys and the output projection into the values (both are
      // constants of the generation): scores = RMSNorm(x) . KQ, x += softmax(scores) . W2 - two phases instead of three
      mega_linear<NR>(a.R, a.eps, a.x, d, L.ln1, 1.0f, L.kq, a.q, a.xfold, nullptr, d, a.xfold, 0, false, xs, part);
      stamp();
      stamp();
      mega_linear<NR>(a.R, a.eps, a.q, a.xfold, nullptr, 1.0f, L.w2, a.x, d, a.x, a.xk, d, 0, false, xs, part, a.xfold / a.R, a.n_enc);
      stamp();
    } else {
      mega_linear<NR>(a.R, a.eps, a.x, d, L.ln1, 1.0f, L.cq, a.q, inner, nullptr, d, inner, 0, false, xs, part);
      stamp();
      grid.sync();
      mega_cross_attn(a, L, sm);
      stamp();
      mega_linear<NR>(a.R, a.eps, a.att, inner, nullptr, 1.0f, L.co, a.x, d, a.x, inner, d, 0, false, xs, part);
      stamp();
    }
    mega_linear<NR>(a.R, a.eps, a.x, d, L.ln2, 1.0f, L.wi, a.hid, a.ff, nullptr, d, a.ff, 1, false, xs, part);
    stamp();
    mega_linear<NR>(a.R, a.eps, a.hid, a.ff, nullptr, 1.0f, L.wo, a.x, d, a.x, a.ff, d, 0, false, xs, part);
    stamp();
  }
  // the LM head streams 65 MB once per step: evict-first loads keep the 88 MB of layer weights in the 126 MB L2
  mega_linear<NR>(a.R, a.eps, a.x, d, a.final_ln, a.lm_scale, a.lm, a.logits, a.vocab, nullptr, d, a.vocab, 0, true, xs, part);
  stamp();
}

// enc-side K / V: proj [R * n_enc, inner] -> [R, H, n_enc, dk]
__global__ void t5_split_heads_kernel(const float* __restrict__ in, int R, int n_enc, int H, int dk, float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(R) * n_enc * H * dk;
  if (i >= total) return;
  const int d = static_cast<int>(i % dk);
  long long x = i / dk;
  const int j = static_cast<int>(x % n_enc); x /= n_enc;
  const int h = static_cast<int>(x % H);
  const int r = static_cast<int>(x / H);
  out[i] = in[(static_cast<size_t>(r) * n_enc + j) * H * dk + h * dk + d];
}

// Folds the two cross-attention projections into the per-generation constants (one CTA per (row, head, token)):
//   kq [R * M, d]  : kq[r * M + m][k] = sum_dd Wq[h * dk + dd][k] * K[r, h, j, dd]          (m = h * n_enc + j, M = H * n_enc)
//   w2 [d, xk]     : w2[c][r * M + m] = sum_dd Wo[c][h * dk + dd] * V[r, h, j, dd]          (columns >= R * M are zero)
// so that scores = RMSNorm(x) . kq^T and output = softmax(scores) . w2^T need no q and no attention phase per token.
__global__ void __launch_bounds__(256) t5_cross_fold_kernel(const float* __restrict__ Wq, const float* __restrict__ Wo,
                                                            const float* __restrict__ Kc, const float* __restrict__ Vc, int H,
                                                            int dk, int n_enc, int d, int xk, float* __restrict__ kq,
                                                            float* __restrict__ w2) {
  __shared__ float kv[64], vv[64];
  const int M = H * n_enc, rm = blockIdx.x, r = rm / M, m = rm - r * M, h = m / n_enc, j = m - h * n_enc, inner = H * dk;
  const size_t src = ((static_cast<size_t>(r) * H + h) * n_enc + j) * dk;
  if (threadIdx.x < dk) { kv[threadIdx.x] = Kc[src + threadIdx.x]; vv[threadIdx.x] = Vc[src + threadIdx.x]; }
  __syncthreads();
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int dd = 0; dd < dk; ++dd) {
      a = fmaf(Wq[static_cast<size_t>(h * dk + dd) * d + k], kv[dd], a);
      b = fmaf(Wo[static_cast<size_t>(k) * inner + h * dk + dd], vv[dd], b);
    }
    kq[static_cast<size_t>(rm) * d + k] = a;
    w2[static_cast<size_t>(k) * xk + rm] = b;
  }
}

// beam reordering: row r continues the hypothesis of row idx[r] - only the ancestry rows move, positions [0, t)
__global__ void t5_reorder_anc_kernel(const int32_t* __restrict__ src, int32_t* __restrict__ dst, const int32_t* __restrict__ idx,
                                      int Tmax, int t, int R) {
  const int r = blockIdx.x;
  const int s = min(max(idx[r], 0), R - 1);
  for (int j = threadIdx.x; j < t; j += blockDim.x) dst[static_cast<size_t>(r) * Tmax + j] = src[static_cast<size_t>(s) * Tmax + j];
}

// ---- beam-search scoring (native search, mmdx_b200/t5_fast.py NativeBeamSearch / mmdx_t5_generate)
// score(row, token) = log_softmax(logits)[row, token] + beam_score[row] in HF's formulation ((x - max) - log(sum)) + score,
// minus infinity for banned tokens (no-repeat-n-gram lists, -1 padded) and for EOS while below the minimum length; per study
// the k best (row, token) pairs over its `beams` rows, descending, ties to the lower flat index row_in_study * V + token.
// Three small launches over kScoreChunks slices of the vocabulary (one CTA scanning 4 x 32128 logits took 115 us):
//   t5_lse_ban_kernel  (chunk, row)   : partial (max, sum exp) of the slice, THEN the row's bans inside the slice are written
//                                       into the logits (HF applies its processors to log_softmax(logits), not to the logits)
//   t5_topk_chunk_kernel (chunk, study): row statistics from the partials, thread-local top-k of the slice, k block-wide pops
//   t5_topk_merge_kernel (study)      : k pops over the chunks' candidates - the same total order, so the same result as one pass
constexpr int kTopK = 8;
constexpr int kScoreChunks = 32;
constexpr int kScoreThreads = 256;

__device__ __forceinline__ int score_chunk_width(int V) { return (V + kScoreChunks - 1) / kScoreChunks; }

__global__ void __launch_bounds__(kScoreThreads) t5_lse_ban_kernel(float* __restrict__ logits, int V, float* __restrict__ part,
                                                                   const int32_t* __restrict__ banned, int max_ban, int ban_eos,
                                                                   int eos) {
  __shared__ float red[kScoreThreads / 32];
  const int c = blockIdx.x, r = blockIdx.y, tid = threadIdx.x;
  const int cw = score_chunk_width(V), lo = c * cw, hi = min(V, lo + cw);
  float* x = logits + static_cast<size_t>(r) * V;
  float mx = -INFINITY;
  for (int i = lo + tid; i < hi; i += kScoreThreads) mx = fmaxf(mx, x[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < kScoreThreads / 32; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sm = 0.f;
  for (int i = lo + tid; i < hi; i += kScoreThreads) sm += expf(x[i] - mx);
  sm = warp_sum_f(sm);
  if ((tid & 31) == 0) red[tid >> 5] = sm;
  __syncthreads();                                   // also: every read of the slice is done before the bans are written
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < kScoreThreads / 32; ++i) t += red[i];
    part[(static_cast<size_t>(r) * kScoreChunks + c) * 2] = mx;
    part[(static_cast<size_t>(r) * kScoreChunks + c) * 2 + 1] = t;
  }
  if (ban_eos && tid == 0 && eos >= lo && eos < hi) x[eos] = -INFINITY;
  for (int q = tid; q < max_ban; q += kScoreThreads) {
    const int bt = banned[static_cast<size_t>(r) * max_ban + q];
    if (bt >= lo && bt < hi) x[bt] = -INFINITY;
  }
}

// The k best of the threads' sorted candidate lists, in descending order (ties: the lower index first).  Every warp pops
// its own k winners with shuffles only (no block barrier per round), the 8 x k survivors meet in shared memory and warp 0
// pops the final k from them.  The order is total, so the result equals k rounds of a block-wide arg-max.  kScoreThreads threads.
__device__ __forceinline__ void warp_argmax(float& v, int& ix, int& th) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
    const int ot = __shfl_xor_sync(0xffffffffu, th, o);
    if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; th = ot; }
  }
}

__device__ void block_pop_topk(const float (&best)[kTopK], const int (&bidx)[kTopK], int k, float* __restrict__ out_scores,
                               int32_t* __restrict__ out_idx) {
  constexpr int W = kScoreThreads / 32;
  __shared__ float sv[W * kTopK];
  __shared__ int si[W * kTopK];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int head = 0;                               // this thread's next candidate
  for (int round = 0; round < k; ++round) {
    float v = -INFINITY;
    int ix = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < kTopK; ++j) if (j == head) { v = best[j]; ix = bidx[j]; }
    int th = lane;
    warp_argmax(v, ix, th);
    if (lane == th) ++head;
    if (lane == 0) { sv[warp * kTopK + round] = v; si[warp * kTopK + round] = ix; }
  }
  __syncthreads();
  if (warp == 0) {                            // W * k <= 64 survivors: two per lane, each warp's list already sorted
    float c0 = -INFINITY, c1 = -INFINITY;
    int i0 = 0x7fffffff, i1 = 0x7fffffff;
    const int a = lane, b = lane + 32;
    if (a < W * kTopK && (a % kTopK) < k) { c0 = sv[a]; i0 = si[a]; }
    if (b < W * kTopK && (b % kTopK) < k) { c1 = sv[b]; i1 = si[b]; }
    if (c1 > c0 || (c1 == c0 && i1 < i0)) { const float tv = c0; c0 = c1; c1 = tv; const int ti = i0; i0 = i1; i1 = ti; }
    int used = 0;                             // 0: c0 is next, 1: c1, 2: none left
    for (int round = 0; round < k; ++round) {
      float v = used == 0 ? c0 : (used == 1 ? c1 : -INFINITY);
      int ix = used == 0 ? i0 : (used == 1 ? i1 : 0x7fffffff);
      int th = lane;
      warp_argmax(v, ix, th);
      if (lane == th) ++used;
      if (lane == 0) { out_scores[round] = v; out_idx[round] = ix; }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kScoreThreads) t5_topk_chunk_kernel(const float* __restrict__ logits, const float* __restrict__ part,
                                                                      const float* __restrict__ beam_scores, int beams, int V,
                                                                      int k, float* __restrict__ cand_s, int32_t* __restrict__ cand_i) {
  extern __shared__ float stat[];             // [beams][3]: row max, log(sum exp), beam score
  const int c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int row = warp; row < beams; row += kScoreThreads / 32) {
    const int R = b * beams + row;
    static_assert(kScoreChunks == 32, "one partial per lane");
    const float m = part[(static_cast<size_t>(R) * kScoreChunks + lane) * 2];
    const float sc = part[(static_cast<size_t>(R) * kScoreChunks + lane) * 2 + 1];
    float M = m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
    const float S = warp_sum_f(sc > 0.f ? sc * expf(m - M) : 0.f);
    if (lane == 0) { stat[3 * row] = M; stat[3 * row + 1] = logf(S); stat[3 * row + 2] = beam_scores[R]; }
  }
  __syncthreads();
  float best[kTopK];
  int bidx[kTopK];
#pragma unroll
  for (int j = 0; j < kTopK; ++j) { best[j] = -INFINITY; bidx[j] = 0x7fffffff; }
  const int cw = score_chunk_width(V), lo = c * cw, n = max(0, min(V, lo + cw) - lo);
  // rows outermost (no index divisions), tokens strided over the threads: ascending flat index per thread; the loads of a
  // row are issued together, then the (data-dependent) insertions run as compare-exchanges on registers
  for (int row = 0; row < beams; ++row) {
    const float* x = logits + (static_cast<size_t>(b) * beams + row) * V + lo;
    const float mx = stat[3 * row], ls = stat[3 * row + 1], bs = stat[3 * row + 2];
    for (int t0 = tid; t0 < n; t0 += 4 * kScoreThreads) {
      float x4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x4[u] = t0 + u * kScoreThreads < n ? x[t0 + u * kScoreThreads] : -INFINITY;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v = (x4[u] - mx) - ls;
        v += bs;
        if (v > best[kTopK - 1]) {           // strictly greater: an equal later (higher) index never displaces an earlier one
          best[kTopK - 1] = v; bidx[kTopK - 1] = row * V + lo + t0 + u * kScoreThreads;
#pragma unroll
          for (int j = kTopK - 1; j > 0; --j) {
            if (best[j] > best[j - 1]) {     // bubble up past strictly smaller entries only
              const float tv = best[j]; best[j] = best[j - 1]; best[j - 1] = tv;
              const int ti = bidx[j]; bidx[j] = bidx[j - 1]; bidx[j - 1] = ti;
            }
          }
        }
      }
    }
  }
  const size_t o = (static_cast<size_t>(b) * kScoreChunks + c) * k;
  block_pop_topk(best, bidx, k, cand_s + o, cand_i + o);
}

__global__ void __launch_bounds__(kScoreThreads) t5_topk_merge_kernel(const float* __restrict__ cand_s, const int32_t* __restrict__ cand_i,
                                                                      int k, float* __restrict__ out_scores, int32_t* __restrict__ out_idx,
                                                                      int* __restrict__ done_ctr, volatile int* host_flag, int flag_value) {
  const int b = blockIdx.x, tid = threadIdx.x;
  float best[kTopK];
  int bidx[kTopK];
#pragma unroll
  for (int j = 0; j < kTopK; ++j) { best[j] = -INFINITY; bidx[j] = 0x7fffffff; }
  static_assert(kScoreChunks * kTopK <= kScoreThreads, "one candidate per thread");
  if (tid < kScoreChunks * k) {
    best[0] = cand_s[static_cast<size_t>(b) * kScoreChunks * k + tid];
    bidx[0] = cand_i[static_cast<size_t>(b) * kScoreChunks * k + tid];
  }
  block_pop_topk(best, bidx, k, out_scores + static_cast<size_t>(b) * k, out_idx + static_cast<size_t>(b) * k);
  // mmdx_t5_generate: the outputs live in mapped host memory and the host spins on the studies' flags instead of
  // synchronising the stream
  (void)done_ctr;
  if (host_flag != nullptr && tid == 0) {     // one flag per study: results, one system-scope fence, flag
    __threadfence_system();
    host_flag[b] = flag_value;
  }
}

struct Block { float *ln0, *qkv, *so, *ln1, *cq, *ck, *cv, *co, *ln2, *wi, *wo; };

}  // namespace

struct mmdx_t5 {
  int device = 0, d = 512, H = 8, dk = 64, ff = 2048, L = 6, vocab = 32128, tied = 1;
  float eps = 1e-6f;
  std::recursive_mutex mu;         // recursive: mmdx_t5_generate drives begin / step / score / reorder under one lock
  std::map<std::string, std::vector<float>> host;
  bool finalized = false;
  float* arena = nullptr; size_t arena_floats = 0, used = 0;
  float *E = nullptr, *lm = nullptr, *final_ln = nullptr;
  std::vector<Block> blocks;
  // per generation
  int R = 0, n_enc = 0, Tmax = 0, t = 0;
  float* ws = nullptr; size_t ws_floats = 0;
  float *x = nullptr, *qkv = nullptr, *att = nullptr, *q = nullptr, *hid = nullptr, *bias = nullptr, *score_ws = nullptr;
  std::vector<float*> sk, sv, ck, cv;
  int32_t* anc[2] = {nullptr, nullptr};   // ancestry rows [R, Tmax], double-buffered for the permutation
  int cur = 0;                        // which ancestry copy is live
  std::vector<T5Layer> h_layers;      // per-block pointers (refreshed by mmdx_t5_begin), passed to the step kernel by value
  unsigned long long* prof = nullptr; // device stamps of the last step (MMDX_T5_PROF=1)
  int xfold = 0, xk = 0;              // folded cross-attention of the current generation (0 = off)
  int mega_blocks = 0;                // co-resident CTAs of the step kernel (one per SM)
  float* gen_dev = nullptr; size_t gen_dev_words = 0;       // mmdx_t5_generate scratch (device / pinned host), grow-only
  float* gen_host = nullptr; size_t gen_host_words = 0;
  int64_t launches = 0;
};

static int t5_linear(mmdx_t5* e, const float* x, long long ldx, const float* ln_w, float in_scale, const float* W, float* out,
                     long long ldo, const float* res, int R, int K, int N, int relu, cudaStream_t s) {
  T5_REQUIRE(K % 128 == 0 && K <= 4096, "linear: K must be a multiple of 128 (<= 4096)");
  const size_t smem = static_cast<size_t>(kRowChunk) * K * 4;
  static int attr_done[64] = {};
  if (smem > 48 * 1024 && !attr_done[e->device & 63]) {
    T5_CK(cudaFuncSetAttribute(t5_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowChunk * 4096 * 4));
    attr_done[e->device & 63] = 1;
  }
  dim3 grid((N + 8 * kColsPerWarp - 1) / (8 * kColsPerWarp), (R + kRowChunk - 1) / kRowChunk);
  t5_linear_kernel<<<grid, 256, smem, s>>>(x, ldx, ln_w, e->eps, in_scale, W, out, ldo, res, R, K, N, relu);
  e->launches++;
  T5_CK(cudaGetLastError());
  return 0;
}

extern "C" const char* mmdx_t5_last_error(void) { return g_t5_err.c_str(); }

extern "C" int mmdx_t5_create(int device, int d_model, int n_heads, int d_kv, int d_ff, int n_layers, int vocab, float eps,
                              int tied_embeddings, mmdx_t5** out) {
  T5_REQUIRE(out != nullptr, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return t5_fail("mmdx_t5: no CUDA device (there is no CPU fallback)");
  T5_REQUIRE(device >= 0 && device < ndev, "bad device ordinal");
  T5_REQUIRE(d_model % 128 == 0 && d_ff % 128 == 0 && (n_heads * d_kv) % 128 == 0 && d_kv <= 64 && d_kv % 4 == 0 && n_layers > 0 && n_layers <= kMaxT5Layers &&
                 vocab > 0, "unsupported T5 dimensions");
  mmdx_t5* e = new mmdx_t5();
  e->device = device; e->d = d_model; e->H = n_heads; e->dk = d_kv; e->ff = d_ff; e->L = n_layers; e->vocab = vocab; e->eps = eps;
  e->tied = tied_embeddings;
  *out = e;
  return 0;
}

extern "C" void mmdx_t5_destroy(mmdx_t5* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  if (e->arena) cudaFree(e->arena);
  if (e->ws) cudaFree(e->ws);
  if (e->prof) cudaFree(e->prof);
  if (e->gen_dev) cudaFree(e->gen_dev);
  if (e->gen_host) cudaFreeHost(e->gen_host);
  delete e;
}

// name = key of T5ForConditionalGeneration.state_dict() ("shared.weight", "decoder.block.0.layer.0.SelfAttention.q.weight", ...)
extern "C" int mmdx_t5_load_tensor(mmdx_t5* e, const char* name, const float* h_data, int64_t n_elems) {
  T5_REQUIRE(e && name && h_data && n_elems > 0, "bad argument");
  T5_REQUIRE(!e->finalized, "weights already finalized");
  e->host[name].assign(h_data, h_data + n_elems);
  return 0;
}

extern "C" int mmdx_t5_finalize(mmdx_t5* e) {
  T5_REQUIRE(e && !e->finalized, "bad state");
  T5_CK(cudaSetDevice(e->device));
  size_t total = 0;
  for (auto& kv : e->host) total += kv.second.size() + 64;
  T5_CK(cudaMalloc(&e->arena, total * 4));
  e->arena_floats = total; e->used = 0;
  auto up = [&](const std::vector<float>& v, float** out) -> int {
    T5_REQUIRE(e->used + v.size() <= e->arena_floats, "arena overflow");
    *out = e->arena + e->used;
    T5_CK(cudaMemcpy(*out, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
    e->used += (v.size() + 63) & ~size_t(63);
    return 0;
  };
  auto get = [&](const std::string& k, size_t n, float** out) -> int {
    auto it = e->host.find(k);
    if (it == e->host.end()) return t5_fail("mmdx_t5: missing weight tensor " + k);
    if (it->second.size() != n) return t5_fail("mmdx_t5: wrong size for " + k);
    return up(it->second, out);
  };
  const int d = e->d, inner = e->H * e->dk;
  if (get("shared.weight", (size_t)e->vocab * d, &e->E)) return 1;
  e->lm = e->E;
  if (e->tied == 0 && get("lm_head.weight", (size_t)e->vocab * d, &e->lm)) return 1;
  if (get("decoder.final_layer_norm.weight", d, &e->final_ln)) return 1;
  e->blocks.clear();
  for (int i = 0; i < e->L; ++i) {
    const std::string p = "decoder.block." + std::to_string(i) + ".layer.";
    Block b{};
    if (get(p + "0.layer_norm.weight", d, &b.ln0)) return 1;
    {   // q | k | v stacked: one launch
      std::vector<float> w3;
      for (const char* n : {"q", "k", "v"}) {
        auto it = e->host.find(p + "0.SelfAttention." + n + ".weight");
        if (it == e->host.end() || it->second.size() != (size_t)inner * d) return t5_fail("mmdx_t5: missing / bad self-attention weight");
        w3.insert(w3.end(), it->second.begin(), it->second.end());
      }
      if (up(w3, &b.qkv)) return 1;
    }
    if (get(p + "0.SelfAttention.o.weight", (size_t)d * inner, &b.so)) return 1;
    if (get(p + "1.layer_norm.weight", d, &b.ln1)) return 1;
    if (get(p + "1.EncDecAttention.q.weight", (size_t)inner * d, &b.cq)) return 1;
    if (get(p + "1.EncDecAttention.k.weight", (size_t)inner * d, &b.ck)) return 1;
    if (get(p + "1.EncDecAttention.v.weight", (size_t)inner * d, &b.cv)) return 1;
    if (get(p + "1.EncDecAttention.o.weight", (size_t)d * inner, &b.co)) return 1;
    if (get(p + "2.layer_norm.weight", d, &b.ln2)) return 1;
    if (get(p + "2.DenseReluDense.wi.weight", (size_t)e->ff * d, &b.wi)) return 1;
    if (get(p + "2.DenseReluDense.wo.weight", (size_t)d * e->ff, &b.wo)) return 1;
    e->blocks.push_back(b);
  }
  e->host.clear();
  e->finalized = true;
  T5_CK(cudaDeviceSynchronize());
  return 0;
}

// Start a generation: R rows (batch x beams, beams of one study adjacent), d_enc [R, n_enc, d_model] conditioning tokens
// per row, at most max_steps tokens.  h_bias [max_steps][heads]: relative-position bias by distance (host; computed by the
// caller with the reference's own bucket arithmetic).  Projects the cross-attention keys / values once.
extern "C" int mmdx_t5_begin(mmdx_t5* e, const float* d_enc, int R, int n_enc, int max_steps, const float* h_bias, void* stream) {
  T5_REQUIRE(e && d_enc && h_bias && R > 0 && n_enc > 0 && n_enc <= 64 && max_steps > 0 && max_steps <= 4096, "bad argument");
  std::lock_guard<std::recursive_mutex> lk(e->mu);
  T5_REQUIRE(e->finalized, "weights not finalized");
  T5_CK(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int d = e->d, inner = e->H * e->dk;
  const size_t cache = (size_t)R * e->H * max_steps * e->dk, cross = (size_t)R * e->H * n_enc * e->dk;
  // cross-attention folding (see t5_cross_fold_kernel): worth it while the R x R score matrix stays narrow
  const int xfold = R * e->H * n_enc <= inner && e->d == inner && std::getenv("MMDX_T5_NOFOLD") == nullptr ? R * e->H * n_enc : 0;
  const int xk = (xfold + 127) / 128 * 128;
  const size_t fold_words = xfold > 0 ? (size_t)xfold * d + (size_t)d * xk : 0;
  const size_t anc_words = ((size_t)R * max_steps + 3) / 4 * 4;          // int32 rows, kept 16-byte aligned
  const size_t layer_words = (sizeof(T5Layer) * e->L + 15) / 16 * 4;
  const size_t need = (size_t)R * (d + 3 * inner + inner + inner + e->ff) + (size_t)max_steps * e->H + 4 + (size_t)e->L * (2 * cache + 2 * cross) +
                      (size_t)R * n_enc * inner + (size_t)e->L * fold_words + 2 * anc_words + layer_words + ((size_t)R * kScoreChunks * (2 + 2 * kTopK) + 4) + 1024;
  if (need > e->ws_floats) {
    if (e->ws) { T5_CK(cudaDeviceSynchronize()); cudaFree(e->ws); e->ws = nullptr; e->ws_floats = 0; }
    // headroom: a later generation with more rows or more steps should not have to free and reallocate (a device-wide
    // synchronisation plus a cudaFree / cudaMalloc pair in the middle of serving)
    const size_t grant = std::max(need + need / 2, (size_t)16 << 20);
    T5_CK(cudaMalloc(&e->ws, grant * 4));
    e->ws_floats = grant;
  }
  float* p = e->ws;
  e->x = p; p += (size_t)R * d;
  e->qkv = p; p += (size_t)R * 3 * inner;
  e->att = p; p += (size_t)R * inner;
  e->q = p; p += (size_t)R * inner;
  e->hid = p; p += (size_t)R * e->ff;
  e->bias = p; p += ((size_t)max_steps * e->H + 3) / 4 * 4;
  float* tmp = p; p += (size_t)R * n_enc * inner;
  e->sk.clear(); e->sv.clear(); e->ck.clear(); e->cv.clear();
  std::vector<T5Layer> hl(e->L);
  for (int l = 0; l < e->L; ++l) {
    e->sk.push_back(p); p += cache;
    e->sv.push_back(p); p += cache;
    e->ck.push_back(p); p += cross;
    e->cv.push_back(p); p += cross;
    const Block& b = e->blocks[l];
    float *kq = nullptr, *w2 = nullptr;
    if (xfold > 0) { kq = p; p += (size_t)xfold * d; w2 = p; p += (size_t)d * xk; }
    hl[l] = T5Layer{b.ln0, b.qkv, b.so, b.ln1, b.cq, b.co, b.ln2, b.wi, b.wo, e->sk[l], e->sv[l], e->ck[l], e->cv[l], kq, w2};
  }
  e->xfold = xfold; e->xk = xk;
  e->anc[0] = reinterpret_cast<int32_t*>(p); p += anc_words;
  e->anc[1] = reinterpret_cast<int32_t*>(p); p += anc_words;
  (void)layer_words;
  e->score_ws = p; p += ((size_t)R * kScoreChunks * (2 + 2 * kTopK) + 4);
  T5_CK(cudaMemsetAsync(e->score_ws + (size_t)R * kScoreChunks * (2 + 2 * kTopK), 0, 16, s));      // the merge kernel's arrival counter
  e->h_layers = hl;
  e->R = R; e->n_enc = n_enc; e->Tmax = max_steps; e->t = 0; e->cur = 0;
  T5_CK(cudaMemcpyAsync(e->bias, h_bias, (size_t)max_steps * e->H * 4, cudaMemcpyHostToDevice, s));
  const long long tot = (long long)R * n_enc * inner;
  for (int l = 0; l < e->L; ++l) {
    if (t5_linear(e, d_enc, d, nullptr, 1.0f, e->blocks[l].ck, tmp, inner, nullptr, R * n_enc, d, inner, 0, s)) return 1;
    t5_split_heads_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(tmp, R, n_enc, e->H, e->dk, e->ck[l]);
    if (t5_linear(e, d_enc, d, nullptr, 1.0f, e->blocks[l].cv, tmp, inner, nullptr, R * n_enc, d, inner, 0, s)) return 1;
    t5_split_heads_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(tmp, R, n_enc, e->H, e->dk, e->cv[l]);
    e->launches += 2;
    if (xfold > 0) {
      float* w2 = const_cast<float*>(e->h_layers[l].w2);
      if (xk > xfold) T5_CK(cudaMemsetAsync(w2, 0, (size_t)d * xk * 4, s));
      t5_cross_fold_kernel<<<xfold, 256, 0, s>>>(e->blocks[l].cq, e->blocks[l].co, e->ck[l], e->cv[l], e->H, e->dk, n_enc, d, xk,
                                                 const_cast<float*>(e->h_layers[l].kq), w2);
      e->launches++;
    }
  }
  T5_CK(cudaGetLastError());
  return 0;
}

// Beam reordering between steps: row r continues the hypothesis that lived in row d_beam_idx[r].
extern "C" int mmdx_t5_reorder(mmdx_t5* e, const int32_t* d_beam_idx, void* stream) {
  T5_REQUIRE(e && d_beam_idx, "null argument");
  std::lock_guard<std::recursive_mutex> lk(e->mu);
  T5_REQUIRE(e->R > 0, "mmdx_t5_reorder before mmdx_t5_begin");
  T5_CK(cudaSetDevice(e->device));
  if (e->t == 0) return 0;
  const int nxt = e->cur ^ 1;
  t5_reorder_anc_kernel<<<e->R, 128, 0, (cudaStream_t)stream>>>(e->anc[e->cur], e->anc[nxt], d_beam_idx, e->Tmax, e->t, e->R);
  e->launches++;
  e->cur = nxt;
  T5_CK(cudaGetLastError());
  return 0;
}

// One decoder step: d_tokens [R] -> d_logits [R, vocab] fp32; the new position joins the cache.  d_beam_idx (optional): the
// beam reordering that precedes this step, done inside the same launch.
static int t5_step_impl(mmdx_t5* e, const int32_t* d_tokens, float* d_logits, const int32_t* d_beam_idx, void* stream,
                        const int32_t* copy_src = nullptr, int32_t* copy_dst = nullptr, int copy_n = 0) {
  T5_REQUIRE(e && d_tokens && d_logits, "null argument");
  std::lock_guard<std::recursive_mutex> lk(e->mu);
  T5_REQUIRE(e->R > 0, "mmdx_t5_step before mmdx_t5_begin");
  T5_REQUIRE(e->t < e->Tmax, "more steps than mmdx_t5_begin reserved");
  T5_CK(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int R = e->R, d = e->d, inner = e->H * e->dk;
  T5StepArgs a{};
  a.tok = d_tokens; a.logits = d_logits; a.E = e->E; a.lm = e->lm; a.final_ln = e->final_ln; a.bias = e->bias;
  for (int l = 0; l < e->L; ++l) a.layers[l] = e->h_layers[l];
  a.anc = e->anc[e->cur];
  a.copy_src = copy_src; a.copy_dst = copy_dst; a.copy_n = copy_n;
  if (d_beam_idx != nullptr && e->t > 0) {
    a.anc_src = e->anc[e->cur]; a.anc = e->anc[e->cur ^ 1]; a.bidx = d_beam_idx;
    e->cur ^= 1;
  }
  a.x = e->x; a.qkv = e->qkv; a.att = e->att; a.q = e->q; a.hid = e->hid;
  a.eps = e->eps;
  a.lm_scale = e->tied == 1 ? 1.0f / std::sqrt((float)d) : 1.0f;     // HF scales the decoder output only in the default tied setup
  a.R = R; a.d = d; a.H = e->H; a.dk = e->dk; a.ff = e->ff; a.L = e->L; a.vocab = e->vocab; a.Tmax = e->Tmax; a.t = e->t;
  a.n_enc = e->n_enc; a.kmax = std::max(std::max(d, inner), e->ff);
  a.prof = e->prof;
  a.xfold = e->xfold; a.xk = e->xk;
  const int NR = R <= 4 ? 4 : (R <= 8 ? 8 : 16);       // rows per pass over the weights
  const size_t smem = std::max((size_t)NR * a.kmax + (size_t)kMegaWarps * NR, (size_t)96 + 2048 + 2 * (size_t)e->Tmax) * 4;
  T5_REQUIRE(smem <= 200 * 1024, "step: d_ff / max_steps too large for the shared-memory staging");
  void* fn = NR == 4 ? (void*)t5_step_mega_kernel<4> : (NR == 8 ? (void*)t5_step_mega_kernel<8> : (void*)t5_step_mega_kernel<16>);
  if (e->mega_blocks == 0) {
    cudaDeviceProp prop;
    T5_CK(cudaGetDeviceProperties(&prop, e->device));
    T5_REQUIRE(prop.cooperativeLaunch, "device lacks cooperative launch");
    e->mega_blocks = prop.multiProcessorCount;
    T5_CK(cudaFuncSetAttribute(t5_step_mega_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    T5_CK(cudaFuncSetAttribute(t5_step_mega_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    T5_CK(cudaFuncSetAttribute(t5_step_mega_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const char* pf = std::getenv("MMDX_T5_PROF");
    if (pf && pf[0] == '1') { T5_CK(cudaMalloc(&e->prof, 1024 * 8)); T5_CK(cudaMemset(e->prof, 0, 1024 * 8)); }
  }
  void* args[] = {&a};
  T5_CK(cudaLaunchCooperativeKernel(fn, dim3(e->mega_blocks), dim3(kMegaThreads), args, smem, s));
  e->launches++;
  T5_CK(cudaGetLastError());
  e->t++;
  return 0;
}

extern "C" int mmdx_t5_step(mmdx_t5* e, const int32_t* d_tokens, float* d_logits, void* stream) {
  return t5_step_impl(e, d_tokens, d_logits, nullptr, stream);
}

// Beam-search scoring of the logits mmdx_t5_step has just produced: per study the k best (row, token) continuations of
// log_softmax(logits) + beam_score, with EOS and the per-row banned tokens (int32 [R, max_ban], -1 padded; may be null when
// max_ban = 0) masked out.  d_out_scores / d_out_idx [R / num_beams, k], descending, idx = row_in_study * vocab + token.
static int t5_score_topk_impl(mmdx_t5* e, float* d_logits, const float* d_beam_scores, const int32_t* d_banned,
                              int max_ban, int ban_eos, int eos_id, int num_beams, int k, float* d_out_scores,
                              int32_t* d_out_idx, int* d_host_flag, int flag_value, void* stream) {
  T5_REQUIRE(e && d_logits && d_beam_scores && d_out_scores && d_out_idx, "null argument");
  std::lock_guard<std::recursive_mutex> lk(e->mu);
  T5_REQUIRE(e->R > 0 && num_beams > 0 && e->R % num_beams == 0, "mmdx_t5_score_topk: rows must be studies x beams");
  T5_REQUIRE(k >= 1 && k <= kTopK && (max_ban == 0 || d_banned != nullptr), "mmdx_t5_score_topk: 1 <= k <= 8");
  T5_CK(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  T5_REQUIRE(!(ban_eos || max_ban > 0) || (eos_id >= 0 && eos_id < e->vocab), "eos id outside the vocabulary");
  const int B = e->R / num_beams;
  float* part = e->score_ws;                                          // [R][chunks][2]
  float* cand_s = part + (size_t)e->R * kScoreChunks * 2;             // [B][chunks][k]
  int32_t* cand_i = reinterpret_cast<int32_t*>(cand_s + (size_t)e->R * kScoreChunks * kTopK);
  t5_lse_ban_kernel<<<dim3(kScoreChunks, e->R), kScoreThreads, 0, s>>>(d_logits, e->vocab, part, d_banned, max_ban, ban_eos, eos_id);
  t5_topk_chunk_kernel<<<dim3(kScoreChunks, B), kScoreThreads, (size_t)num_beams * 3 * 4, s>>>(d_logits, part, d_beam_scores, num_beams,
                                                                                               e->vocab, k, cand_s, cand_i);
  int* done_ctr = reinterpret_cast<int*>(cand_i + (size_t)e->R * kScoreChunks * kTopK);
  t5_topk_merge_kernel<<<B, kScoreThreads, 0, s>>>(cand_s, cand_i, k, d_out_scores, d_out_idx, done_ctr, d_host_flag, flag_value);
  e->launches += 3;
  T5_CK(cudaGetLastError());
  return 0;
}

extern "C" int mmdx_t5_score_topk(mmdx_t5* e, float* d_logits, const float* d_beam_scores, const int32_t* d_banned,
                                  int max_ban, int ban_eos, int eos_id, int num_beams, int k, float* d_out_scores,
                                  int32_t* d_out_idx, void* stream) {
  return t5_score_topk_impl(e, d_logits, d_beam_scores, d_banned, max_ban, ban_eos, eos_id, num_beams, k, d_out_scores, d_out_idx,
                            nullptr, 0, stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// Whole generation in one call: the beam search of t5_fast.NativeBeamSearch (= HF GenerationMixin._beam_search for
// do_sample = False, one EOS id, num_return_sequences = 1, with the min-new-tokens and no-repeat-n-gram processors) with
// the per-token bookkeeping here on the host instead of in Python: per token one decoder step, the scoring / top-k
// launches, one 2K-candidate read-back per study and a few dozen scalar operations.
// early_stopping: 0 = False, 1 = True, 2 = "never".  h_out [B, max_new_tokens + 1] (start token first) is filled with HF's
// fill value past each hypothesis; *h_out_len = the common output length HF would return (longest best hypothesis).
__global__ void t5_repeat_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int K, long long row_elems) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * K * row_elems;
  if (i >= total) return;
  const long long r = i / row_elems, c = i - r * row_elems;
  out[i] = in[(r / K) * row_elems + c];
}

extern "C" int mmdx_t5_generate(mmdx_t5* e, const float* d_cond, int B, int n_enc, int num_beams, int max_new_tokens,
                                int min_new_tokens, int no_repeat_ngram, float length_penalty, int early_stopping, int eos_id,
                                int pad_id, int start_id, const float* h_bias, int32_t* h_out, int32_t* h_out_len, void* stream) {
  T5_REQUIRE(e && d_cond && h_bias && h_out && h_out_len, "null argument");
  T5_REQUIRE(B > 0 && n_enc > 0 && num_beams >= 1 && 2 * num_beams <= kTopK && max_new_tokens >= 1 && max_new_tokens < 4096,
             "mmdx_t5_generate: 1 <= num_beams <= 4, 1 <= max_new_tokens < 4096");
  std::lock_guard<std::recursive_mutex> lk(e->mu);
  T5_REQUIRE(e->finalized, "weights not finalized");
  T5_CK(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int K = num_beams, K2 = 2 * K, R = B * K, V = e->vocab, prompt = 1, max_length = 1 + max_new_tokens;
  const int fill = pad_id != 0 ? pad_id : eos_id;                 // HF: `pad_token_id or eos_token_id[0]`
  const float NEG = -1.0e9f;
  const double lp = (double)length_penalty;
  // ---- device scratch of the search (freed on every exit path by the guard)
  // Per token the host writes one block (beam indices of the last step | tokens | running scores | banned lists) and reads
  // one (2K candidate scores | indices per study | a sequence flag).  Both live in pinned host memory that the kernels
  // read and write directly (zero-copy over PCIe: a few hundred bytes), and the host spins on the flag the merge kernel
  // raises instead of synchronising the stream: no copy-engine launches and no stream synchronisation per token.  The
  // buffers belong to the engine and only ever grow.
  const long long enc_row = (long long)n_enc * e->d;
  const int max_ban = max_length;                                 // a row can ban at most one token per earlier position
  const size_t stage_words = (size_t)R * (3 + max_ban), cand_words = (size_t)2 * B * K2;
  const size_t dev_words = (size_t)R * enc_row + (size_t)R * V + (size_t)R * (1 + max_ban) + 64;
  if (dev_words > e->gen_dev_words) {
    if (e->gen_dev) { T5_CK(cudaDeviceSynchronize()); cudaFree(e->gen_dev); e->gen_dev = nullptr; e->gen_dev_words = 0; }
    const size_t grant = std::max(dev_words + dev_words / 2, (size_t)4 << 20);
    T5_CK(cudaMalloc(&e->gen_dev, grant * 4));
    e->gen_dev_words = grant;
  }
  if (stage_words + cand_words + B + 16 > e->gen_host_words) {
    if (e->gen_host) { T5_CK(cudaDeviceSynchronize()); cudaFreeHost(e->gen_host); e->gen_host = nullptr; e->gen_host_words = 0; }
    const size_t grant = std::max(2 * (stage_words + cand_words + B + 16), (size_t)1 << 18);
    T5_CK(cudaHostAlloc(&e->gen_host, grant * 4, cudaHostAllocMapped));
    e->gen_host_words = grant;
  }
  float* host_dev = nullptr;                                      // the device's view of the pinned block
  T5_CK(cudaHostGetDevicePointer(&host_dev, e->gen_host, 0));
  struct { float *enc, *logits; int32_t *h_stage, *h_cand, *score_in; } w;
  w.enc = e->gen_dev;
  w.logits = w.enc + (size_t)R * enc_row;
  w.score_in = reinterpret_cast<int32_t*>(w.logits + (size_t)R * V);      // device copy of (running scores | banned lists)
  w.h_stage = reinterpret_cast<int32_t*>(e->gen_host);
  w.h_cand = w.h_stage + stage_words;
  volatile int* h_flag = reinterpret_cast<volatile int*>(w.h_cand + cand_words);
  for (int b = 0; b < B; ++b) h_flag[b] = 0;
  int32_t* const dv_stage = reinterpret_cast<int32_t*>(host_dev);
  int32_t* const dv_cand = dv_stage + stage_words;
  int* const dv_flag = dv_cand + cand_words;
  int32_t* const h_bidx = w.h_stage;
  int32_t* const h_tok = w.h_stage + R;
  float* const h_bscore = reinterpret_cast<float*>(w.h_stage + 2 * R);
  int32_t* const h_ban = w.h_stage + 3 * R;
  const int32_t* d_bidx = dv_stage;
  const int32_t* d_tok = dv_stage + R;
  const float* d_bscore = reinterpret_cast<const float*>(w.score_in);
  const int32_t* d_ban = w.score_in + R;
  float* d_oscore = reinterpret_cast<float*>(dv_cand);
  int32_t* d_oidx = dv_cand + (size_t)B * K2;
  const float* h_oscore = reinterpret_cast<const float*>(w.h_cand);
  const int32_t* h_oidx = w.h_cand + (size_t)B * K2;
  {
    const long long tot = (long long)R * enc_row;
    t5_repeat_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d_cond, w.enc, B, K, enc_row);
    e->launches++;
  }
  if (mmdx_t5_begin(e, w.enc, R, n_enc, max_length, h_bias, s)) return 1;
  // ---- host state (HF names): running = hypotheses still being extended, finished pool = `sequences` / `beam_scores`
  std::vector<int32_t> run_seq((size_t)R * max_length, fill), seqs((size_t)R * max_length, fill);
  std::vector<int32_t> tk_seq((size_t)B * K2 * max_length), m_seq((size_t)(K + K2) * max_length), new_run((size_t)R * max_length);
  std::vector<float> run_scores(R, NEG), beam_scores(R, NEG);
  std::vector<int> fin_len(R, 0);
  std::vector<char> finished(R, 0), unsat(B, 1);
  for (int b = 0; b < B; ++b) {
    run_scores[b * K] = 0.f;
    for (int k = 0; k < K; ++k) { run_seq[(size_t)(b * K + k) * max_length] = start_id; seqs[(size_t)(b * K + k) * max_length] = start_id; }
  }
  std::vector<int32_t> h_banned((size_t)R * max_ban);
  for (int r = 0; r < R; ++r) h_bidx[r] = r;
  int cur_len = 1;
  while (true) {
    for (int r = 0; r < R; ++r) { h_tok[r] = run_seq[(size_t)r * max_length + cur_len - 1]; h_bscore[r] = run_scores[r]; }
    // no-repeat-n-gram bans of every running row (HF NoRepeatNGramLogitsProcessor over the tokens generated so far)
    int ban_w = 0;
    const int n = no_repeat_ngram;
    if (n > 0 && cur_len + 1 >= n) {
      std::fill(h_banned.begin(), h_banned.end(), -1);
      for (int r = 0; r < R; ++r) {
        const int32_t* q = &run_seq[(size_t)r * max_length];
        int cnt = 0;
        for (int i = 0; i + n - 1 < cur_len; ++i) {              // window q[i .. i+n-2] against the current suffix
          bool same = true;
          for (int j = 0; j < n - 1 && same; ++j) same = q[i + j] == q[cur_len - n + 1 + j];
          if (same) h_banned[(size_t)r * max_ban + cnt++] = q[i + n - 1];
        }
        ban_w = std::max(ban_w, cnt);
      }
      for (int r = 0; r < R; ++r)                                 // repack to [R, ban_w]
        for (int c = 0; c < ban_w; ++c) h_ban[(size_t)r * ban_w + c] = h_banned[(size_t)r * max_ban + c];
    }
    std::atomic_thread_fence(std::memory_order_seq_cst);          // the block is complete before the kernels are enqueued
    if (t5_step_impl(e, d_tok, w.logits, cur_len > 1 ? d_bidx : nullptr, s, dv_stage + 2 * R, w.score_in, R + R * ban_w)) return 1;
    const int ban_eos = (cur_len - prompt) < min_new_tokens ? 1 : 0;
    if (t5_score_topk_impl(e, w.logits, d_bscore, ban_w > 0 ? d_ban : nullptr, ban_w, ban_eos, eos_id, K, K2, d_oscore, d_oidx,
                           dv_flag, cur_len, s))
      return 1;
    auto all_flags = [&]() { for (int b = 0; b < B; ++b) if (h_flag[b] != cur_len) return false; return true; };
    for (unsigned spins = 0; !all_flags(); ++spins) {             // the merge kernel raises a study's flag behind its results
      if ((spins & 0xfff) == 0xfff) {
        const cudaError_t q = cudaStreamQuery(s);
        if (q == cudaSuccess && !all_flags()) return t5_fail("mmdx_t5_generate: the stream drained without the completion flags");
        if (q != cudaSuccess && q != cudaErrorNotReady) return t5_fail(std::string("mmdx_t5_generate: ") + cudaGetErrorString(q));
      }
    }
    std::atomic_thread_fence(std::memory_order_seq_cst);
    bool any_unsat = false, all_finished = true, all_hits = true;
    for (int b = 0; b < B; ++b) {
      const float* tks = h_oscore + (size_t)b * K2;
      const int32_t* tki = h_oidx + (size_t)b * K2;
      int src[kTopK], tokn[kTopK]; bool hit[kTopK]; float pool[kTopK];
      for (int c = 0; c < K2; ++c) {
        src[c] = tki[c] / V; tokn[c] = tki[c] % V;
        int32_t* dst = &tk_seq[((size_t)b * K2 + c) * max_length];
        std::copy_n(&run_seq[(size_t)(b * K + src[c]) * max_length], max_length, dst);
        dst[cur_len] = tokn[c];
        hit[c] = tokn[c] == eos_id || cur_len + 1 >= max_length;
        pool[c] = tks[c] + (hit[c] ? 1.0f : 0.0f) * NEG;
        all_hits = all_hits && hit[c];
      }
      // running beams of the next step: the best K continuations that did not stop (descending, ties: candidate order)
      int order[kTopK];
      for (int c = 0; c < K2; ++c) order[c] = c;
      std::stable_sort(order, order + K2, [&](int a, int c2) { return pool[a] > pool[c2]; });
      for (int k = 0; k < K; ++k) {
        const int c = order[k];
        std::copy_n(&tk_seq[((size_t)b * K2 + c) * max_length], max_length, &new_run[(size_t)(b * K + k) * max_length]);
        run_scores[b * K + k] = pool[c];
        h_bidx[b * K + k] = b * K + src[c];
      }
      // finished pool: only the top K of the 2K continuations may finish; merge with the stored ones, keep the best K
      const float denom = (float)std::pow((double)(cur_len + 1 - prompt), lp);
      bool full = true;
      for (int k = 0; k < K; ++k) full = full && finished[b * K + k];
      full = full && early_stopping == 1;
      float m_scores[kTopK + 4]; char m_fin[kTopK + 4]; int m_len[kTopK + 4];
      for (int k = 0; k < K; ++k) {
        m_scores[k] = beam_scores[b * K + k]; m_fin[k] = finished[b * K + k]; m_len[k] = fin_len[b * K + k];
        std::copy_n(&seqs[(size_t)(b * K + k) * max_length], max_length, &m_seq[(size_t)k * max_length]);
      }
      for (int c = 0; c < K2; ++c) {
        const bool just = hit[c] && c < K;
        float fs = tks[c] / denom;
        fs = fs + (full ? 1.0f : 0.0f) * NEG;
        fs = fs + (unsat[b] ? 0.0f : 1.0f) * NEG;
        fs = fs + (just ? 0.0f : 1.0f) * NEG;
        m_scores[K + c] = fs; m_fin[K + c] = just; m_len[K + c] = cur_len + 1 - prompt;
        std::copy_n(&tk_seq[((size_t)b * K2 + c) * max_length], max_length, &m_seq[(size_t)(K + c) * max_length]);
      }
      int mo[kTopK + 4];
      for (int c = 0; c < K + K2; ++c) mo[c] = c;
      std::stable_sort(mo, mo + K + K2, [&](int a, int c2) { return m_scores[a] > m_scores[c2]; });
      for (int k = 0; k < K; ++k) {
        const int c = mo[k];
        beam_scores[b * K + k] = m_scores[c]; finished[b * K + k] = m_fin[c]; fin_len[b * K + k] = m_len[c];
        std::copy_n(&m_seq[(size_t)c * max_length], max_length, &seqs[(size_t)(b * K + k) * max_length]);
      }
    }
    run_seq.swap(new_run);                                        // h_bidx goes up with the next token's block
    ++cur_len;
    // can a running beam still beat the worst finished hypothesis?
    const int best_len = (early_stopping == 2 && lp > 0.0) ? (max_length - prompt) : (cur_len - prompt);
    const float bden = (float)std::pow((double)best_len, lp);
    for (int b = 0; b < B; ++b) {
      const float best_possible = run_scores[b * K] / bden;
      float mn = beam_scores[b * K];
      for (int k = 1; k < K; ++k) mn = std::min(mn, beam_scores[b * K + k]);
      bool any = false;
      for (int k = 0; k < K; ++k) any = any || best_possible > (finished[b * K + k] ? mn : NEG);
      unsat[b] = unsat[b] && any;
      any_unsat = any_unsat || unsat[b];
      for (int k = 0; k < K; ++k) all_finished = all_finished && finished[b * K + k];
    }
    const bool go_on = any_unsat && !(all_finished && early_stopping == 1) && !all_hits;
    if (!go_on) break;
  }
  T5_CK(cudaStreamSynchronize(s));
  int out_len = 0;
  for (int b = 0; b < B; ++b) out_len = std::max(out_len, fin_len[b * K]);
  out_len += prompt;
  for (int b = 0; b < B; ++b) std::copy_n(&seqs[(size_t)(b * K) * max_length], max_length, h_out + (size_t)b * max_length);
  *h_out_len = out_len;
  return 0;
}

extern "C" int64_t mmdx_t5_launch_count(mmdx_t5* e) { return e ? e->launches : 0; }

// Phase boundaries of the LAST decoder step as globaltimer nanoseconds (CTA 0's view): entry, then one stamp behind every
// phase (per block: qkv, self-attention, o, cross q, cross-attention, o, wi, wo; finally the LM head).  Needs MMDX_T5_PROF=1
// in the environment when the first step runs; synchronises the device.
extern "C" int mmdx_t5_step_profile(mmdx_t5* e, uint64_t* h_out, int cap, int* n_out) {
  T5_REQUIRE(e && h_out && n_out, "null argument");
  std::lock_guard<std::recursive_mutex> lk(e->mu);
  T5_REQUIRE(e->prof != nullptr, "mmdx_t5_step_profile: set MMDX_T5_PROF=1 before the first step");
  const int n = std::min(cap, 2 + 8 * e->L);
  T5_CK(cudaSetDevice(e->device));
  T5_CK(cudaDeviceSynchronize());
  T5_CK(cudaMemcpy(h_out, e->prof, (size_t)n * 8, cudaMemcpyDeviceToHost));
  *n_out = n;
  return 0;
}
