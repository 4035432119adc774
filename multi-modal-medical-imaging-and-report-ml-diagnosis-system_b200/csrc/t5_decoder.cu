// KV-cached T5 decoder step: the model call inside the reference's report generation
// (FusionTransformerModel.generate -> HF T5ForConditionalGeneration.generate, training_pipeline.py:613-618;
// inference_pipeline.py:190-196) as hand-written CUDA kernels behind a C ABI (include/mmdx.h, mmdx_t5_*).
//
// One step = one new token for each of R = batch x beams rows.  With R <= 64 every contraction is a skinny GEMV-like
// product whose cost is reading the weights once (t5-small decoder + tied LM head: 41 M parameters, 165 MB fp32 per
// step), so this is HBM-bound work on the CUDA cores: fp32 weights, fp32 arithmetic (the beam search compares sums of
// log-probabilities; fp32 keeps the token sequence identical to HF's eager fp32 path), 50 launches per step:
//   x = E[token]
//   per block:  qkv = RMSNorm(x) [Wq;Wk;Wv]^T ; self-attention over the cache (+ relative-position bias, no 1/sqrt(d)) ;
//               x += attn Wo^T ; q = RMSNorm(x) Wq^T ; cross-attention over the projected conditioning tokens ;
//               x += attn Wo^T ; x += relu(RMSNorm(x) Wi^T) Wo^T
//   logits = (RMSNorm(x) * d_model^-0.5) E^T            (tied embeddings; lm_head.weight when untied)
// The beam search itself stays HF's code (mmdx_b200/t5_fast.py swaps only the model call), which reorders the cache
// through mmdx_t5_reorder.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
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

__global__ void t5_embed_kernel(const int32_t* __restrict__ tok, const float* __restrict__ E, int d, int vocab,
                                float* __restrict__ x) {
  const int r = blockIdx.x;
  const int id = min(max(tok[r], 0), vocab - 1);
  for (int i = threadIdx.x; i < d; i += blockDim.x) x[static_cast<size_t>(r) * d + i] = E[static_cast<size_t>(id) * d + i];
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

// Self-attention of the new token against the cache.  grid (heads, rows), 128 threads.
// qkv [R, 3*inner] (q | k | v), cache_k / cache_v [R, H, Tmax, dk]; position t is written first.
// scores[j] = q . k_j + bias[(t - j) * H + h]  (no scaling in T5), softmax in fp32, out [R, inner].
__global__ void __launch_bounds__(128) t5_self_attn_kernel(const float* __restrict__ qkv, float* __restrict__ ck,
                                                           float* __restrict__ cv, const float* __restrict__ bias, int H,
                                                           int dk, int Tmax, int t, float* __restrict__ out) {
  extern __shared__ float sm[];       // q[dk] | p[Tmax]
  __shared__ float red[4];
  float* q = sm;
  float* p = sm + dk;
  const int h = blockIdx.x, r = blockIdx.y, tid = threadIdx.x, inner = H * dk;
  const float* row = qkv + static_cast<size_t>(r) * 3 * inner;
  float* kc = ck + (static_cast<size_t>(r) * H + h) * Tmax * dk;
  float* vc = cv + (static_cast<size_t>(r) * H + h) * Tmax * dk;
  if (tid < dk) {
    q[tid] = row[h * dk + tid];
    kc[static_cast<size_t>(t) * dk + tid] = row[inner + h * dk + tid];
    vc[static_cast<size_t>(t) * dk + tid] = row[2 * inner + h * dk + tid];
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = tid; j <= t; j += 128) {
    const float4* kj = reinterpret_cast<const float4*>(kc + static_cast<size_t>(j) * dk);      // 16-byte loads of this key's row
    const float4* q4 = reinterpret_cast<const float4*>(q);
    float s = 0.f;
    for (int d = 0; d < dk / 4; ++d) {
      const float4 kv = kj[d], qv = q4[d];
      s = fmaf(qv.x, kv.x, fmaf(qv.y, kv.y, fmaf(qv.z, kv.z, fmaf(qv.w, kv.w, s))));
    }
    s += bias[static_cast<size_t>(t - j) * H + h];
    p[j] = s;
    mx = fmaxf(mx, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j <= t; j += 128) { const float e = expf(p[j] - mx); p[j] = e; sum += e; }
  sum = warp_sum_f(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  const float inv = 1.0f / (red[0] + red[1] + red[2] + red[3]);
  // P V: the two halves of the block take alternate keys (independent loads, four in flight), then meet in shared memory
  __shared__ float part[64];
  const int d = tid & 63, half = tid >> 6;
  float o = 0.f;
  if (d < dk) {
    int j = half;
    for (; j + 6 <= t; j += 8) {
      const float v0 = vc[static_cast<size_t>(j) * dk + d], v1 = vc[static_cast<size_t>(j + 2) * dk + d];
      const float v2 = vc[static_cast<size_t>(j + 4) * dk + d], v3 = vc[static_cast<size_t>(j + 6) * dk + d];
      o = fmaf(p[j] * inv, v0, o); o = fmaf(p[j + 2] * inv, v1, o);
      o = fmaf(p[j + 4] * inv, v2, o); o = fmaf(p[j + 6] * inv, v3, o);
    }
    for (; j <= t; j += 2) o = fmaf(p[j] * inv, vc[static_cast<size_t>(j) * dk + d], o);
  }
  if (half == 1 && d < dk) part[d] = o;
  __syncthreads();
  if (half == 0 && d < dk) out[static_cast<size_t>(r) * inner + h * dk + d] = o + part[d];
}

// Cross-attention of the new token against the projected conditioning tokens: q [R, inner], K / V [R, H, n_enc, dk].
__global__ void __launch_bounds__(64) t5_cross_attn_kernel(const float* __restrict__ qm, const float* __restrict__ K,
                                                           const float* __restrict__ V, int H, int dk, int n_enc,
                                                           float* __restrict__ out) {
  __shared__ float p[64];
  const int h = blockIdx.x, r = blockIdx.y, tid = threadIdx.x, inner = H * dk;
  const float* q = qm + static_cast<size_t>(r) * inner + h * dk;
  const float* kc = K + (static_cast<size_t>(r) * H + h) * n_enc * dk;
  const float* vc = V + (static_cast<size_t>(r) * H + h) * n_enc * dk;
  if (tid < n_enc) {
    float s = 0.f;
    for (int d = 0; d < dk; ++d) s = fmaf(q[d], kc[static_cast<size_t>(tid) * dk + d], s);
    p[tid] = s;
  }
  __syncthreads();
  float mx = -INFINITY, sum = 0.f;
  for (int j = 0; j < n_enc; ++j) mx = fmaxf(mx, p[j]);
  for (int j = 0; j < n_enc; ++j) sum += expf(p[j] - mx);
  if (tid < dk) {
    float o = 0.f;
    for (int j = 0; j < n_enc; ++j) o = fmaf(expf(p[j] - mx) / sum, vc[static_cast<size_t>(j) * dk + tid], o);
    out[static_cast<size_t>(r) * inner + h * dk + tid] = o;
  }
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

// beam reordering of the self-attention cache: dst[r] = src[idx[r]] for positions [0, t)
__global__ void t5_reorder_kernel(const float* __restrict__ src, float* __restrict__ dst, const int32_t* __restrict__ idx, int H,
                                  int Tmax, int dk, int t, int R) {
  const int r = blockIdx.y, h = blockIdx.x;
  const int s = min(max(idx[r], 0), R - 1);
  const float4* a = reinterpret_cast<const float4*>(src + (static_cast<size_t>(s) * H + h) * Tmax * dk);
  float4* b = reinterpret_cast<float4*>(dst + (static_cast<size_t>(r) * H + h) * Tmax * dk);
  for (int i = threadIdx.x; i < t * dk / 4; i += blockDim.x) b[i] = a[i];
}

// ---- beam-search scoring (native search, mmdx_b200/t5_fast.py NativeBeamSearch)
// Row statistics of log_softmax: (max, log(sum exp(x - max))) per row, in HF's formulation lp = (x - max) - log(sum).
__global__ void __launch_bounds__(1024) t5_row_lse_kernel(const float* __restrict__ logits, int V, float* __restrict__ stat) {
  __shared__ float red[32];
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* x = logits + static_cast<size_t>(r) * V;
  float mx = -INFINITY;
  for (int i = tid; i < V; i += 1024) mx = fmaxf(mx, x[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < 32; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sm = 0.f;
  for (int i = tid; i < V; i += 1024) sm += expf(x[i] - mx);
  sm = warp_sum_f(sm);
  if ((tid & 31) == 0) red[tid >> 5] = sm;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < 32; ++i) t += red[i];
    stat[2 * r] = mx; stat[2 * r + 1] = logf(t);
  }
}

// Top-k continuations of one study over its `beams` rows: score(row, token) = log_softmax(logits)[row, token] + beam_score
// [row], minus infinity for banned tokens (no-repeat-n-gram lists, -1 padded) and for EOS while below the minimum length.
// One block per study; every thread keeps its own top-k (k <= 8) of a strided share, then k rounds of a block-wide arg-max
// pop the winners in descending order (ties: the lower flat index first).  out_idx = row_in_study * V + token.
// (The masks are written into the logits by t5_ban_kernel AFTER the row statistics were taken: HF applies its logits
// processors to log_softmax(logits), not to the logits.)
__global__ void t5_ban_kernel(float* __restrict__ logits, const int32_t* __restrict__ banned, int max_ban, int ban_eos, int eos,
                              int V) {
  const int r = blockIdx.x;
  float* x = logits + static_cast<size_t>(r) * V;
  if (ban_eos && threadIdx.x == 0) x[eos] = -INFINITY;
  for (int q = threadIdx.x; q < max_ban; q += blockDim.x) {
    const int bt = banned[static_cast<size_t>(r) * max_ban + q];
    if (bt >= 0 && bt < V) x[bt] = -INFINITY;
  }
}

constexpr int kTopK = 8;
__global__ void __launch_bounds__(1024) t5_topk_kernel(const float* __restrict__ logits, const float* __restrict__ stat,
                                                       const float* __restrict__ beam_scores, int beams, int V, int k,
                                                       float* __restrict__ out_scores, int32_t* __restrict__ out_idx) {
  __shared__ float sv[32];
  __shared__ int si[32];
  __shared__ int s_win_thread;
  const int b = blockIdx.x, tid = threadIdx.x;
  float best[kTopK];
  int bidx[kTopK];
#pragma unroll
  for (int j = 0; j < kTopK; ++j) { best[j] = -INFINITY; bidx[j] = 0x7fffffff; }
  const int total = beams * V;
  for (int i = tid; i < total; i += 1024) {
    const int row = i / V, tok = i - row * V;
    const int R = b * beams + row;
    float v = (logits[static_cast<size_t>(R) * V + tok] - stat[2 * R]) - stat[2 * R + 1];
    v += beam_scores[R];
    if (v > best[kTopK - 1]) {               // strictly greater: an equal later (higher) index never displaces an earlier one
      int j = kTopK - 1;
      while (j > 0 && v > best[j - 1]) { best[j] = best[j - 1]; bidx[j] = bidx[j - 1]; --j; }
      best[j] = v; bidx[j] = i;
    }
  }
  int head = 0;                               // this thread's next candidate
  for (int round = 0; round < k; ++round) {
    float v = head < kTopK ? best[head] : -INFINITY;
    int ix = head < kTopK ? bidx[head] : 0x7fffffff;
    int th = tid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
      const int ot = __shfl_xor_sync(0xffffffffu, th, o);
      if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; th = ot; }
    }
    if ((tid & 31) == 0) { sv[tid >> 5] = v; si[tid >> 5] = ix; }
    __shared__ int st[32];
    if ((tid & 31) == 0) st[tid >> 5] = th;
    __syncthreads();
    if (tid == 0) {
      float bv = sv[0]; int bi = si[0], bt = st[0];
      for (int w = 1; w < 32; ++w)
        if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; bt = st[w]; }
      out_scores[b * k + round] = bv;
      out_idx[b * k + round] = bi;
      s_win_thread = bt;
    }
    __syncthreads();
    if (tid == s_win_thread) ++head;
    __syncthreads();
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
  float *x = nullptr, *qkv = nullptr, *att = nullptr, *q = nullptr, *hid = nullptr, *bias = nullptr;
  std::vector<float*> sk[2], sv[2], ck, cv;
  int cur = 0;                        // which of the two self-attention cache copies is live
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
  T5_REQUIRE(d_model % 128 == 0 && d_ff % 128 == 0 && (n_heads * d_kv) % 128 == 0 && d_kv <= 64 && d_kv % 4 == 0 && n_layers > 0 &&
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
  const size_t need = (size_t)R * (d + 3 * inner + inner + inner + e->ff) + (size_t)max_steps * e->H + (size_t)e->L * (4 * cache + 2 * cross) +
                      (size_t)R * n_enc * inner + 1024;
  if (need > e->ws_floats) {
    if (e->ws) { T5_CK(cudaDeviceSynchronize()); cudaFree(e->ws); e->ws = nullptr; }
    T5_CK(cudaMalloc(&e->ws, need * 4));
    e->ws_floats = need;
  }
  float* p = e->ws;
  e->x = p; p += (size_t)R * d;
  e->qkv = p; p += (size_t)R * 3 * inner;
  e->att = p; p += (size_t)R * inner;
  e->q = p; p += (size_t)R * inner;
  e->hid = p; p += (size_t)R * e->ff;
  e->bias = p; p += (size_t)max_steps * e->H;
  float* tmp = p; p += (size_t)R * n_enc * inner;
  for (int c = 0; c < 2; ++c) { e->sk[c].clear(); e->sv[c].clear(); }
  e->ck.clear(); e->cv.clear();
  for (int l = 0; l < e->L; ++l) {
    for (int c = 0; c < 2; ++c) { e->sk[c].push_back(p); p += cache; e->sv[c].push_back(p); p += cache; }
    e->ck.push_back(p); p += cross;
    e->cv.push_back(p); p += cross;
  }
  e->R = R; e->n_enc = n_enc; e->Tmax = max_steps; e->t = 0; e->cur = 0;
  T5_CK(cudaMemcpyAsync(e->bias, h_bias, (size_t)max_steps * e->H * 4, cudaMemcpyHostToDevice, s));
  const long long tot = (long long)R * n_enc * inner;
  for (int l = 0; l < e->L; ++l) {
    if (t5_linear(e, d_enc, d, nullptr, 1.0f, e->blocks[l].ck, tmp, inner, nullptr, R * n_enc, d, inner, 0, s)) return 1;
    t5_split_heads_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(tmp, R, n_enc, e->H, e->dk, e->ck[l]);
    if (t5_linear(e, d_enc, d, nullptr, 1.0f, e->blocks[l].cv, tmp, inner, nullptr, R * n_enc, d, inner, 0, s)) return 1;
    t5_split_heads_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(tmp, R, n_enc, e->H, e->dk, e->cv[l]);
    e->launches += 2;
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
  for (int l = 0; l < e->L; ++l) {
    t5_reorder_kernel<<<dim3(e->H, e->R), 128, 0, (cudaStream_t)stream>>>(e->sk[e->cur][l], e->sk[nxt][l], d_beam_idx, e->H, e->Tmax, e->dk, e->t, e->R);
    t5_reorder_kernel<<<dim3(e->H, e->R), 128, 0, (cudaStream_t)stream>>>(e->sv[e->cur][l], e->sv[nxt][l], d_beam_idx, e->H, e->Tmax, e->dk, e->t, e->R);
    e->launches += 2;
  }
  e->cur = nxt;
  T5_CK(cudaGetLastError());
  return 0;
}

// One decoder step: d_tokens [R] -> d_logits [R, vocab] fp32; the new position joins the cache.
extern "C" int mmdx_t5_step(mmdx_t5* e, const int32_t* d_tokens, float* d_logits, void* stream) {
  T5_REQUIRE(e && d_tokens && d_logits, "null argument");
  std::lock_guard<std::recursive_mutex> lk(e->mu);
  T5_REQUIRE(e->R > 0, "mmdx_t5_step before mmdx_t5_begin");
  T5_REQUIRE(e->t < e->Tmax, "more steps than mmdx_t5_begin reserved");
  T5_CK(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int R = e->R, d = e->d, inner = e->H * e->dk;
  t5_embed_kernel<<<R, 128, 0, s>>>(d_tokens, e->E, d, e->vocab, e->x);
  e->launches++;
  for (int l = 0; l < e->L; ++l) {
    const Block& b = e->blocks[l];
    if (t5_linear(e, e->x, d, b.ln0, 1.0f, b.qkv, e->qkv, 3 * inner, nullptr, R, d, 3 * inner, 0, s)) return 1;
    t5_self_attn_kernel<<<dim3(e->H, R), 128, (size_t)(e->dk + e->Tmax) * 4, s>>>(e->qkv, e->sk[e->cur][l], e->sv[e->cur][l], e->bias, e->H,
                                                                                  e->dk, e->Tmax, e->t, e->att);
    if (t5_linear(e, e->att, inner, nullptr, 1.0f, b.so, e->x, d, e->x, R, inner, d, 0, s)) return 1;
    if (t5_linear(e, e->x, d, b.ln1, 1.0f, b.cq, e->q, inner, nullptr, R, d, inner, 0, s)) return 1;
    t5_cross_attn_kernel<<<dim3(e->H, R), 64, 0, s>>>(e->q, e->ck[l], e->cv[l], e->H, e->dk, e->n_enc, e->att);
    if (t5_linear(e, e->att, inner, nullptr, 1.0f, b.co, e->x, d, e->x, R, inner, d, 0, s)) return 1;
    if (t5_linear(e, e->x, d, b.ln2, 1.0f, b.wi, e->hid, e->ff, nullptr, R, d, e->ff, 1, s)) return 1;
    if (t5_linear(e, e->hid, e->ff, nullptr, 1.0f, b.wo, e->x, d, e->x, R, e->ff, d, 0, s)) return 1;
    e->launches += 2;
  }
  const float sc = e->tied == 1 ? 1.0f / std::sqrt((float)d) : 1.0f;     // HF scales the decoder output only in the default tied setup
  if (t5_linear(e, e->x, d, e->final_ln, sc, e->lm, d_logits, e->vocab, nullptr, R, d, e->vocab, 0, s)) return 1;
  T5_CK(cudaGetLastError());
  e->t++;
  return 0;
}

// Beam-search scoring of the logits mmdx_t5_step has just produced: per study the k best (row, token) continuations of
// log_softmax(logits) + beam_score, with EOS and the per-row banned tokens (int32 [R, max_ban], -1 padded; may be null when
// max_ban = 0) masked out.  d_out_scores / d_out_idx [R / num_beams, k], descending, idx = row_in_study * vocab + token.
extern "C" int mmdx_t5_score_topk(mmdx_t5* e, float* d_logits, const float* d_beam_scores, const int32_t* d_banned,
                                  int max_ban, int ban_eos, int eos_id, int num_beams, int k, float* d_out_scores,
                                  int32_t* d_out_idx, void* stream) {
  T5_REQUIRE(e && d_logits && d_beam_scores && d_out_scores && d_out_idx, "null argument");
  std::lock_guard<std::recursive_mutex> lk(e->mu);
  T5_REQUIRE(e->R > 0 && num_beams > 0 && e->R % num_beams == 0, "mmdx_t5_score_topk: rows must be studies x beams");
  T5_REQUIRE(k >= 1 && k <= kTopK && (max_ban == 0 || d_banned != nullptr), "mmdx_t5_score_topk: 1 <= k <= 8");
  T5_CK(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  float* stat = e->hid;                        // 2 floats per row: scratch that the next step overwrites anyway
  t5_row_lse_kernel<<<e->R, 1024, 0, s>>>(d_logits, e->vocab, stat);
  if (ban_eos || max_ban > 0) {
    T5_REQUIRE(eos_id >= 0 && eos_id < e->vocab, "eos id outside the vocabulary");
    t5_ban_kernel<<<e->R, 128, 0, s>>>(d_logits, d_banned, max_ban, ban_eos, eos_id, e->vocab);
    e->launches++;
  }
  t5_topk_kernel<<<e->R / num_beams, 1024, 0, s>>>(d_logits, stat, d_beam_scores, num_beams, e->vocab, k, d_out_scores, d_out_idx);
  e->launches += 2;
  T5_CK(cudaGetLastError());
  return 0;
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
  struct Scratch {
    float *enc = nullptr, *logits = nullptr, *bscore = nullptr, *oscore = nullptr;
    int32_t *tok = nullptr, *bidx = nullptr, *banned = nullptr, *oidx = nullptr;
    float* h_oscore = nullptr; int32_t* h_oidx = nullptr;
    ~Scratch() {
      for (void* p : {(void*)enc, (void*)logits, (void*)bscore, (void*)oscore, (void*)tok, (void*)bidx, (void*)banned, (void*)oidx})
        if (p) cudaFree(p);
      if (h_oscore) cudaFreeHost(h_oscore);
      if (h_oidx) cudaFreeHost(h_oidx);
    }
  } w;
  const long long enc_row = (long long)n_enc * e->d;
  const int max_ban = max_length;                                 // a row can ban at most one token per earlier position
  T5_CK(cudaMalloc(&w.enc, (size_t)R * enc_row * 4));
  T5_CK(cudaMalloc(&w.logits, (size_t)R * V * 4));
  T5_CK(cudaMalloc(&w.bscore, (size_t)R * 4));
  T5_CK(cudaMalloc(&w.oscore, (size_t)B * K2 * 4));
  T5_CK(cudaMalloc(&w.tok, (size_t)R * 4));
  T5_CK(cudaMalloc(&w.bidx, (size_t)R * 4));
  T5_CK(cudaMalloc(&w.banned, (size_t)R * max_ban * 4));
  T5_CK(cudaMalloc(&w.oidx, (size_t)B * K2 * 4));
  T5_CK(cudaMallocHost(&w.h_oscore, (size_t)B * K2 * 4));
  T5_CK(cudaMallocHost(&w.h_oidx, (size_t)B * K2 * 4));
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
  std::vector<int32_t> h_tok(R), h_bidx(R), h_banned((size_t)R * max_ban);
  std::vector<float> h_bscore(R);
  int cur_len = 1;
  while (true) {
    for (int r = 0; r < R; ++r) h_tok[r] = run_seq[(size_t)r * max_length + cur_len - 1];
    T5_CK(cudaMemcpyAsync(w.tok, h_tok.data(), (size_t)R * 4, cudaMemcpyHostToDevice, s));
    if (mmdx_t5_step(e, w.tok, w.logits, s)) return 1;
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
      if (ban_w > 0) {
        // repack to [R, ban_w]
        std::vector<int32_t> packed((size_t)R * ban_w, -1);
        for (int r = 0; r < R; ++r)
          for (int c = 0; c < ban_w; ++c) packed[(size_t)r * ban_w + c] = h_banned[(size_t)r * max_ban + c];
        T5_CK(cudaMemcpyAsync(w.banned, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice, s));
        T5_CK(cudaStreamSynchronize(s));                          // `packed` goes out of scope
      }
    }
    const int ban_eos = (cur_len - prompt) < min_new_tokens ? 1 : 0;
    for (int r = 0; r < R; ++r) h_bscore[r] = run_scores[r];
    T5_CK(cudaMemcpyAsync(w.bscore, h_bscore.data(), (size_t)R * 4, cudaMemcpyHostToDevice, s));
    if (mmdx_t5_score_topk(e, w.logits, w.bscore, ban_w > 0 ? w.banned : nullptr, ban_w, ban_eos, eos_id, K, K2, w.oscore, w.oidx, s))
      return 1;
    T5_CK(cudaMemcpyAsync(w.h_oscore, w.oscore, (size_t)B * K2 * 4, cudaMemcpyDeviceToHost, s));
    T5_CK(cudaMemcpyAsync(w.h_oidx, w.oidx, (size_t)B * K2 * 4, cudaMemcpyDeviceToHost, s));
    T5_CK(cudaStreamSynchronize(s));
    bool any_unsat = false, all_finished = true, all_hits = true;
    for (int b = 0; b < B; ++b) {
      const float* tks = w.h_oscore + (size_t)b * K2;
      const int32_t* tki = w.h_oidx + (size_t)b * K2;
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
    run_seq.swap(new_run);
    T5_CK(cudaMemcpyAsync(w.bidx, h_bidx.data(), (size_t)R * 4, cudaMemcpyHostToDevice, s));
    if (mmdx_t5_reorder(e, w.bidx, s)) return 1;
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
