// fp32 mode of the path (BASELINE north_star: "probabilities must agree within 1e-2 in bf16, or 1e-5 if run in fp32";
// SURVEY.md 8c golden set G4).  The reference itself runs fp32 (inference_pipeline.py:156-157): this mode keeps every
// weight and activation in fp32 on the CUDA cores (the long K sums of the contractions run in fp64 accumulators and are
// rounded to fp32 once), so the thresholded label vector is the reference's with no safety margin.  It is a PARITY mode - plain tiled SIMT kernels, one per reference op, an order of
// magnitude slower than the tcgen05 path - and shares nothing with it except preprocessing's integer resample
// (bit-exact either way) and the head tail.  No tensor cores: bf16/tf32 operands could not hold 1e-5.
#pragma once
#include "kernels.cuh"

namespace mmdx {

enum : int { F32_ACT_NONE = 0, F32_ACT_RELU = 1, F32_ACT_GELU = 2 };

struct F32ConvParams {
  const float* in;      // [NB, H, W, Cin]  (a Linear layer is NB = M, H = W = 1, k = 1)
  const float* w;       // [Cout][k*k][Cin], BatchNorm scale folded in
  const float* bias;    // [Cout] (BN shift / Linear bias) or null
  const float* res;     // [NB, OH, OW, Cout] added before the activation, or null
  float* out;           // row pitch ld_out floats
  int NB, H, W, Cin, Cout, k, stride, pad, OH, OW, act;
  long long ld_out;
};

// Implicit-GEMM convolution on CUDA cores: C[M = NB*OH*OW, N = Cout] = A[M, K = k*k*Cin] * W[N, K]^T.
// 64 x 64 output tile per block, K in steps of 16 through shared memory, 4 x 4 outputs per thread.  Operands and results
// are fp32; the K sum runs in a double accumulator (fp32 products are exact in fp64), so the only rounding per output is
// the final one - closer to the exact dot product than any fp32 summation order, the reference's blocked one included,
// which is what lets 53 stacked layers stay within 1e-5 of it.
__global__ void __launch_bounds__(256) f32_conv_kernel(const F32ConvParams p) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x;
  const int M = p.NB * p.OH * p.OW, K = p.k * p.k * p.Cin;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  // loader role: row lr of the tile, 4 consecutive k starting at lk
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int am = m0 + lr;
  int a_n = 0, a_oh = 0, a_ow = 0;
  const bool a_ok = am < M;
  if (a_ok) { a_n = am / (p.OH * p.OW); const int r = am - a_n * p.OH * p.OW; a_oh = r / p.OW; a_ow = r - a_oh * p.OW; }
  const int bn = n0 + lr;
  const bool b_ok = bn < p.Cout;
  const int ty = tid >> 4, tx = tid & 15;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int kk = k0 + lk + u;
      float av = 0.f, bv = 0.f;
      if (kk < K) {
        if (a_ok) {
          const int tap = kk / p.Cin, c = kk - tap * p.Cin;
          const int r = tap / p.k, s = tap - r * p.k;
          const int ih = a_oh * p.stride + r - p.pad, iw = a_ow * p.stride + s - p.pad;
          if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
            av = __ldg(p.in + ((static_cast<size_t>(a_n) * p.H + ih) * p.W + iw) * p.Cin + c);
        }
        if (b_ok) bv = __ldg(p.w + static_cast<size_t>(bn) * K + kk);
      }
      As[lk + u][lr] = av;
      Bs[lk + u][lr] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = static_cast<double>(As[kk][ty * 4 + i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = static_cast<double>(Bs[kk][tx * 4 + j]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.Cout) continue;
      float v = static_cast<float>(acc[i][j]);
      if (p.bias) v += p.bias[n];
      if (p.res) v += p.res[static_cast<size_t>(m) * p.Cout + n];
      if (p.act == F32_ACT_RELU) v = fmaxf(v, 0.f);
      else if (p.act == F32_ACT_GELU) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
      p.out[static_cast<size_t>(m) * p.ld_out + n] = v;
    }
  }
}

// ToTensor + gray->3ch + Normalize in the reference's own op order (training_pipeline.py:115-117): x = u8 / 255, then
// (x - mean) / std - IEEE fp32 division and subtraction, so the result is bit-identical to torchvision's.
__global__ void f32_normalize_kernel(const uint8_t* __restrict__ in, long long npix, int C, float3 mean, float3 std,
                                     float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  int px[3];
  if (C == 1) { px[0] = px[1] = px[2] = in[i]; }
  else { px[0] = in[i * 3]; px[1] = in[i * 3 + 1]; px[2] = in[i * 3 + 2]; }
  out[i * 3 + 0] = (static_cast<float>(px[0]) / 255.0f - mean.x) / std.x;
  out[i * 3 + 1] = (static_cast<float>(px[1]) / 255.0f - mean.y) / std.y;
  out[i * 3 + 2] = (static_cast<float>(px[2]) / 255.0f - mean.z) / std.z;
}

// MaxPool2d(3, stride 2, padding 1), NHWC
__global__ void f32_maxpool_kernel(const float* __restrict__ in, int NB, int H, int W, int C, int OH, int OW,
                                   float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(NB) * OH * OW * C;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  long long r = i / C;
  const int ow = static_cast<int>(r % OW); r /= OW;
  const int oh = static_cast<int>(r % OH);
  const int n = static_cast<int>(r / OH);
  float m = -INFINITY;
  for (int dy = 0; dy < 3; ++dy) {
    const int ih = oh * 2 + dy - 1;
    if (ih < 0 || ih >= H) continue;
    for (int dx = 0; dx < 3; ++dx) {
      const int iw = ow * 2 + dx - 1;
      if (iw < 0 || iw >= W) continue;
      m = fmaxf(m, in[((static_cast<size_t>(n) * H + ih) * W + iw) * C + c]);
    }
  }
  out[i] = m;
}

// AdaptiveAvgPool2d(1): [NB, HW, C] -> [NB, C]
__global__ void f32_avgpool_kernel(const float* __restrict__ in, int NB, int HW, int C, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NB * C) return;
  const int n = i / C, c = i - n * C;
  double s = 0.0;
  for (int k = 0; k < HW; ++k) s += static_cast<double>(in[(static_cast<size_t>(n) * HW + k) * C + c]);
  out[i] = static_cast<float>(s / static_cast<double>(HW));
}

// LayerNorm of fp32 rows (one warp per row, two-pass statistics).  EMBED: row = word[id] + type[tt] + position[pos]
// (HF BertEmbeddings order).  Any width.
template <bool EMBED>
__global__ void __launch_bounds__(256) f32_layernorm_kernel(const float* __restrict__ x, int rows, int N,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, float* __restrict__ y, const int* __restrict__ ids,
                                                            const int* __restrict__ pos, const int* __restrict__ tts,
                                                            const float* __restrict__ word, const float* __restrict__ ptab,
                                                            const float* __restrict__ ttab, int n_word, int n_pos,
                                                            int n_type) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float *a, *b = nullptr, *c = nullptr;
  if (EMBED) {
    a = word + static_cast<size_t>(min(max(ids[row], 0), n_word - 1)) * N;
    b = ttab + static_cast<size_t>(min(max(tts[row], 0), n_type - 1)) * N;
    c = ptab + static_cast<size_t>(min(max(pos[row], 0), n_pos - 1)) * N;
  } else {
    a = x + static_cast<size_t>(row) * N;
  }
  float s = 0.f;
  for (int i = lane; i < N; i += 32) s += EMBED ? (a[i] + b[i]) + c[i] : a[i];
  const float mean = warp_sum(s) / N;
  float q = 0.f;
  for (int i = lane; i < N; i += 32) { const float d = (EMBED ? (a[i] + b[i]) + c[i] : a[i]) - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / N + eps);
  for (int i = lane; i < N; i += 32) {
    const float v = EMBED ? (a[i] + b[i]) + c[i] : a[i];
    y[static_cast<size_t>(row) * N + i] = (v - mean) * rstd * gamma[i] + beta[i];
  }
}

// softmax(Q K^T / sqrt(64)) V over packed tokens, one warp per (sequence, head, query); qkv fp32 [T, 3*hidden].
// Padded keys do not exist in the packed layout (the reference masks them to -inf: exp() = 0 exactly).
constexpr int F32_ATT_MAXL = 512;
__global__ void __launch_bounds__(128) f32_attention_kernel(const float* __restrict__ qkv, const int* __restrict__ cu,
                                                            int heads, int hidden, float scale, float* __restrict__ ctx) {
  __shared__ float sq[4][64];
  __shared__ float sp[4][F32_ATT_MAXL];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.z, head = blockIdx.y;
  const int t0 = cu[seq], len = cu[seq + 1] - t0;
  const int qi = blockIdx.x * 4 + w;
  if (qi >= len) return;
  const size_t ld = 3 * static_cast<size_t>(hidden);
  const float* q = qkv + (t0 + qi) * ld + head * 64;
  sq[w][lane] = q[lane]; sq[w][lane + 32] = q[lane + 32];
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < len; j += 32) {
    const float* k = qkv + (t0 + j) * ld + hidden + head * 64;
    float s = 0.f;
#pragma unroll 16
    for (int d = 0; d < 64; ++d) s = fmaf(sq[w][d], k[d], s);
    s *= scale;
    sp[w][j] = s;
    mx = fmaxf(mx, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < len; j += 32) { const float e = expf(sp[w][j] - mx); sp[w][j] = e; sum += e; }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.0f / sum;
  float o0 = 0.f, o1 = 0.f;
  for (int j = 0; j < len; ++j) {
    const float* v = qkv + (t0 + j) * ld + 2 * hidden + head * 64;
    const float pj = sp[w][j] * inv;
    o0 = fmaf(pj, v[lane], o0);
    o1 = fmaf(pj, v[lane + 32], o1);
  }
  float* out = ctx + static_cast<size_t>(t0 + qi) * hidden + head * 64;
  out[lane] = o0; out[lane + 32] = o1;
}

// mean_pool (training_pipeline.py:452-459) over each sequence's packed tokens
__global__ void f32_seq_mean_pool_kernel(const float* __restrict__ h, const int* __restrict__ cu, int hidden,
                                         float* __restrict__ out) {
  const int seq = blockIdx.x;
  const int t0 = cu[seq], t1 = cu[seq + 1];
  for (int c = threadIdx.x; c < hidden; c += blockDim.x) {
    double s = 0.0;
    for (int t = t0; t < t1; ++t) s += static_cast<double>(h[static_cast<size_t>(t) * hidden + c]);
    out[static_cast<size_t>(seq) * hidden + c] = static_cast<float>(s) / fmaxf(static_cast<float>(t1 - t0), 1e-6f);
  }
}

}  // namespace mmdx
