// Engine + C ABI (include/mmdx.h): weight packing (BN fold, QKV fuse, bf16, K-major), per-shape
// launch plans (tensor maps + workspace), and the forward: K_pre -> stem/maxpool/16 bottlenecks/avgpool
// -> BERT-base over packed tokens -> fused head.  All compute is in hand-written sm_100a kernels.
#include <cmath>
#include <dlfcn.h>
#include <nvjpeg.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mmdx.h"
#include "attention_tcgen05.cuh"
#include "bneck64_tcgen05.cuh"
#include "conv3x3_c64_tcgen05.cuh"
#include "gemm_tcgen05.cuh"
#include "gemm2_tcgen05.cuh"
#include "stem_tcgen05.cuh"
#include "kernels.cuh"
#include "fp32_kernels.cuh"

using namespace mmdx;
typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t _e = (call);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return fail(std::string(#call) + " failed: " + cudaGetErrorString(_e) + " at " + __FILE__ + ":" +  \
                  std::to_string(__LINE__));                                                             \
  } while (0)
#define REQUIRE(cond, msg)                                                            \
  do {                                                                                \
    if (!(cond)) return fail(std::string("mmdx: ") + msg + " [" #cond "]");           \
  } while (0)
#define TRY(expr)                  \
  do {                             \
    int _r = (expr);               \
    if (_r != 0) return _r;        \
  } while (0)

extern "C" const char* mmdx_last_error(void) { return g_err.c_str(); }
extern "C" const char* mmdx_version(void) { return "mmdx-b200 0.1 (sm_100a, tcgen05/TMA)"; }

// ------------------------------------------------------------------------------------------ Pillow tables
extern "C" int mmdx_resample_coeffs(int in_size, int out_size, int out_first, int n, int32_t* first, int32_t* count,
                                    int32_t* weights, int weights_capacity) {
  // Pillow ImagingResample precompute_coeffs + normalize_coeffs_8bpc, triangle filter (support 1.0)
  if (in_size <= 0 || out_size <= 0 || n < 0) return -1;
  const double scale = static_cast<double>(in_size) / out_size;
  const double fs = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * fs;
  const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  if (static_cast<long long>(n) * ksize > weights_capacity) return -1;
  const double ss = 1.0 / fs;
  std::vector<double> k(ksize);
  for (int i = 0; i < n; ++i) {
    const int xx = out_first + i;
    const double center = (xx + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      const double w = a < 1.0 ? 1.0 - a : 0.0;
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    first[i] = xmin;
    count[i] = xmax;
    for (int x = 0; x < ksize; ++x) {
      int32_t q = 0;
      if (x < xmax) q = k[x] < 0 ? static_cast<int32_t>(-0.5 + k[x] * (1 << 22)) : static_cast<int32_t>(0.5 + k[x] * (1 << 22));
      weights[static_cast<size_t>(i) * ksize + x] = q;
    }
  }
  return ksize;
}

extern "C" int mmdx_resize_geometry(int h, int w, int resize_short, int crop, int* out_h, int* out_w, int* top,
                                    int* left) {
  int oh = h, ow = w;
  if (resize_short > 0) {
    const int s = w <= h ? w : h, l = w <= h ? h : w;
    const int ns = resize_short, nl = static_cast<int>(static_cast<double>(resize_short) * l / s);
    if (w <= h) { ow = ns; oh = nl; } else { ow = nl; oh = ns; }
  }
  *out_h = oh; *out_w = ow;
  if (crop > 0) {
    if (oh < crop || ow < crop) return 1;
    *top = static_cast<int>(std::nearbyint((oh - crop) / 2.0));    // Python round(): half-to-even
    *left = static_cast<int>(std::nearbyint((ow - crop) / 2.0));
  } else { *top = 0; *left = 0; }
  return 0;
}

extern "C" int mmdx_padded_dims(int H, int W, int* hp, int* wp) {
  *hp = H + 8 + (H & 1);
  *wp = W + 8 + (W & 1);
  return 0;
}

// ------------------------------------------------------------------------------------------ engine types
struct HostTensor { std::vector<float> data; std::vector<int64_t> shape; };

struct DevBuf {
  void* p = nullptr; size_t bytes = 0;
  int64_t* gen = nullptr;     // bumped when the buffer moves: captured CUDA graphs that point into it are stale
  ~DevBuf() { if (p) cudaFree(p); }
  int ensure(size_t n) {
    if (n <= bytes) return 0;
    if (gen) ++*gen;
    if (p) { cudaDeviceSynchronize(); cudaFree(p); p = nullptr; bytes = 0; }
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) return fail(std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    bytes = n;
    return 0;
  }
};

struct GemmLaunch {
  GemmParams p; int bn = 128; int bk = 64; int cg = 1; int eb = 1;
  int ln = 0;                         // > 0: row LayerNorm fused behind the last N tile of every m-item (row width ln * 256)
  bool c64 = false; C64Params c;      // layer-1 style 3x3 64->64 conv: dedicated halo-tile kernel instead of the GEMM kernel
  int b64 = 0; Bneck64Params b;       // > 0: fused layer-1 bottleneck kernel, variant code (see build_b64)
  int g2 = 0; Gemm2Params d;          // > 0 (= BN): conv3 + next conv1 as one two-GEMM launch (gemm2_tcgen05.cuh)
};

struct ConvW {      // one folded conv: weights [Cout][taps][Cin] bf16, bias fp32
  bf16* w = nullptr; float* bias = nullptr; int cin = 0, cout = 0, k = 1, stride = 1;
};
struct LinW { bf16* w = nullptr; float* bias = nullptr; int nin = 0, nout = 0; };
// fp32 mode (fp32_kernels.cuh): the same tensors kept in fp32 in a second arena (BN folded in fp32, QKV concatenated)
struct Conv32 { float* w = nullptr; float* bias = nullptr; int cin = 0, cout = 0, k = 1, stride = 1; };
struct Lin32 { float* w = nullptr; float* bias = nullptr; int nin = 0, nout = 0; };
struct Block32 { Conv32 c1, c2, c3, ds; bool has_ds = false; };
struct LnW { float* g = nullptr; float* b = nullptr; };
struct Bert32 { Lin32 qkv, ao, ff1, ff2; LnW ln1, ln2; };
// Folded LayerNorm (gemm_tcgen05.cuh): what a GEMM needs to consume LN(x) - or add it as a residual - straight from the
// pre-LayerNorm tensor x and its row sums.
struct LnFold {
  bf16* wf = nullptr;      // consumer GEMM: gamma folded in and the row mean projected out, bf16(gamma[k] W[n][k] - mean_k); null for residual GEMMs
  float* c1 = nullptr;     // consumer: row sums of gamma o W (unused by the kernel: folded into wf); residual: float(bf16(gamma))
  float* c2 = nullptr;     // consumer: W beta + b;                residual: b + beta
  bf16* gdiag = nullptr;   // residual GEMM: [N][64] bf16, row n holds gamma[n] at column n % 64 (diag(gamma) in 64-blocks)
};
struct BertLayerW { LinW qkv, ao, ff1, ff2; LnW ln1, ln2; LnFold f_qkv, f_ao, f_ff1, f_ff2; };
struct Bottleneck {
  ConvW c1, c2, c3, ds; bool has_ds = false;
  ConvW c3ds;      // has_ds: conv3 and the downsample conv as ONE GEMM: weights [Cout][c3.cin + ds.cin], bias b3 + bd
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct ImagePlan {
  int B = 0, H = 0, W = 0, C = 0;
  int crop_h = 0, crop_w = 0, hp = 0, wp = 0;
  int has_x = 0, has_y = 0, off_x = 0, off_y = 0;
  ResampleTable tx{}, ty{};
  PreStrip strip{};
  bf16* in_pad = nullptr; bf16* pool_out = nullptr;
  int sh = 0, sw = 0, ph = 0, pw = 0;   // stem / pool output sizes
  struct Step { int kind; GemmLaunch g; };   // kind 0 = gemm
  StemParams stem_p{};                      // fused conv1 + bn + relu + maxpool
  std::vector<GemmLaunch> convs;            // the 52 bottleneck convolutions in launch order
  bf16* last = nullptr; int last_hw = 0;
  size_t in_pad_bytes = 0;
  DevBuf tables;        // Pillow coefficient tables of this geometry
};
struct TextPlan {
  int T = 0, B = 0;
  bool folded = false;             // LayerNorm folded into the GEMMs: no LayerNorm launches, 5 launches per layer
  bool ln_fused = false;           // ao / ff2 normalise their own rows (no separate LayerNorm launches)
  std::vector<GemmLaunch> gemms;   // per layer: qkv, ao, ff1, ff2
};
struct HeadPlan { int B = 0; GemmLaunch proj_img, proj_txt, fuse; };

struct mmdx_engine {
  mmdx_config cfg{};
  int num_sms = 148;
  EncodeTiledFn encode = nullptr;
  std::mutex mu;
  int64_t launches = 0;
  // optional per-kernel-class timing (bench.py roofline): CUDA events around every launch
  bool profiling = false;
  struct ProfRec { int cls; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;
  int cur_cls = 8;
  cudaStream_t cur_stream = nullptr;
  std::map<std::string, HostTensor> host;
  bool finalized = false;
  // dims
  int d_img = 0, d_txt = 0, d_fuse = 0, n_cls = 0, hidden = 0, n_layers = 0, ffn = 0, feat_dim = 2048;
  int vocab = 0, max_pos = 0, type_vocab = 2;     // rows of the word / position / token-type embedding tables
  // weights (one arena)
  DevBuf warena; size_t wused = 0;
  ConvW stem; bf16* stem_w2 = nullptr;   // stem weights in the stem kernel's resident layout
  std::vector<Bottleneck> blocks;
  bf16 *word = nullptr, *ptab = nullptr, *ttab = nullptr; LnW emb_ln; std::vector<BertLayerW> layers;
  LinW proj_img, proj_txt, fuse; LnW fuse_ln; float* head_w = nullptr; float* head_b = nullptr;
  // fp32 mode: filled by mmdx_finalize_weights when cfg.keep_fp32 != 0
  bool has_f32 = false;
  DevBuf w32; size_t w32_used = 0;
  Conv32 stem32; std::vector<Block32> blocks32; Lin32 proj_img32, proj_txt32, fuse32;
  float *word32 = nullptr, *ptab32 = nullptr, *ttab32 = nullptr; std::vector<Bert32> layers32;
  DevBuf f32_ws;
  LinW cond;             // optional: fusion.cond_proj.0 (z_fuse -> the T5 decoder's conditioning tokens)
  cudaStream_t copy_stream = nullptr;   // H2D of the image batch overlaps the text branch (mmdx_forward_host)
  cudaEvent_t copy_done = nullptr, copy_ready = nullptr;
  // The text branch runs on its own stream next to the image branch (they only meet at the fusion head): both are
  // chains of persistent whole-chip kernels, so they do not share SMs, but each launch's prologue and its last,
  // partly empty wave are filled by the other branch's CTAs.  MMDX_STREAMS=1 (or profiling mode) serialises them.
  cudaStream_t text_stream = nullptr;
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
  bool two_streams = true;
  // Zigzag traversal: consecutive kernels of a branch walk their tiles / rows / units in opposite directions, so a
  // consumer starts with what its producer wrote LAST - the part of a streamed tensor that is still in the 126 MB L2
  // (read in the producer's order, an LRU cache smaller than the tensor gives no hits at all).  MMDX_ZIGZAG=0 turns it off.
  bool zigzag = true;
  int zz_txt = 1, zz_img = 1;   // direction of the next launch on the text / image branch
  DevBuf pre_lut;        // 3 x 256 bf16: ((v / 255) - mean[c]) / std[c], the ToTensor + Normalize arithmetic per byte value
  DevBuf ident;          // 64x64 bf16 identity: B operand of the residual-add MMAs
  CUtensorMap tm_ident{};
  CUtensorMap tm_ident_half{};   // same identity, box of 32 rows: each CTA of a pair holds half of the B operand
  bool attn_force_general = false;   // MMDX_ATTN=general: use the flash-style kernel for short sequences too (experiments)
  int epi_bufs = 0;      // MMDX_EB=1|2 pins the staging buffers per epilogue group of the CTA-pair kernels; 0 = heuristic
  int force_bn = 0;      // MMDX_BN=64|128|192|256 pins the tile width wherever it divides N (experiments)
  int force_cg = 0;      // MMDX_CG=1|2 pins the CTA-group size of every eligible GEMM (experiments); 0 = heuristic
  // workspaces
  DevBuf img_ws, txt_ws, head_ws, tab_ws, io_ws;
  // mmdx_forward_host_submit / _wait: two request slots, each with its own device I/O buffers and events
  DevBuf io_slot[2];
  cudaEvent_t slot_copy_done[2] = {nullptr, nullptr}, slot_tok_done[2] = {nullptr, nullptr}, slot_done[2] = {nullptr, nullptr};
  bool slot_busy[2] = {false, false};
  // Small host requests (B <= graph_max_b) replay a captured CUDA graph of the whole call: at B = 1 the 135 launches
  // of a step cost more host time than GPU time.  One graph per request shape; inputs / outputs go through engine-owned
  // pinned staging buffers so that every pointer inside the graph is fixed.  MMDX_GRAPH_MAX_B=0 turns it off.
  struct HostGraph {
    cudaGraphExec_t exec = nullptr; int64_t launches = 0; int seen = 0;
    int64_t gen = -1;          // ws_gen the plans / workspaces of this shape were last laid out under
    void* h_in = nullptr; void* h_out = nullptr; size_t in_bytes = 0, out_bytes = 0;
  };
  std::map<std::string, HostGraph> host_graphs;
  // Workspace generation: every reallocation of an arena a captured graph points into (img_ws, txt_ws, head_ws, the
  // request slots) bumps it; a graph captured under another generation holds dangling pointers and is dropped.
  int max_pass = 512;        // images of 224 x 224 per pass of the conv stack (MMDX_MAX_PASS, read at mmdx_create; 0 = no limit)
  int64_t ws_gen = 0;
  // Cross-call ordering: the activation / head workspaces are shared by every call, so a call on another stream than
  // the previous one waits for the previous call's last kernel (calls on one stream are ordered anyway).
  cudaEvent_t last_done = nullptr; cudaStream_t last_stream = nullptr; bool have_last = false;
  // optional GPU JPEG decode: nvJPEG, loaded with dlopen on first use so that libmmdx.so has no hard
  // dependency on it
  struct Jpeg {
    void* lib = nullptr; nvjpegHandle_t h = nullptr; nvjpegJpegState_t st = nullptr; int batch = 0; int backend = -1;
    decltype(&nvjpegCreateEx) create_ex = nullptr; decltype(&nvjpegCreateSimple) create_simple = nullptr;
    decltype(&nvjpegDestroy) destroy = nullptr; decltype(&nvjpegJpegStateCreate) state_create = nullptr;
    decltype(&nvjpegJpegStateDestroy) state_destroy = nullptr; decltype(&nvjpegGetImageInfo) image_info = nullptr;
    decltype(&nvjpegDecodeBatchedInitialize) batched_init = nullptr; decltype(&nvjpegDecodeBatched) batched = nullptr;
  } jpeg;
  cudaStream_t graph_stream = nullptr;
  cudaEvent_t graph_fork = nullptr;
  int graph_max_b = 8;
  bool capturing = false;
  std::map<std::string, std::unique_ptr<ImagePlan>> img_plans;
  ImagePlan* img_last = nullptr;   // plan whose zero borders currently sit in img_ws
  std::map<std::string, std::unique_ptr<TextPlan>> txt_plans;
  std::map<int, std::unique_ptr<HeadPlan>> head_plans;
  // head-side persistent buffers (capacity B)
  int head_cap = 0;
  bf16* feats_bf = nullptr; bf16* pooled_bf = nullptr; bf16* zcat = nullptr; float* fuse_h = nullptr;
  bool defer_proj = false;    // inside mmdx_forward* at B <= kHeadFusedMaxB with head_mode 2: the encoders leave their projections to head_fused_kernel
  bool fused_tail = false;    // inside mmdx_forward* at B <= kHeadFusedMaxB: fusion MLP + head tail as one launch
  int head_mode = 2;          // MMDX_HEAD_FUSED at the first forward: 0 = four launches, 1 = F1 + O1 fused, 2 = I2 + T8 + F1 + O1 fused (default)
  int head_cluster = 0;       // cluster size of head_fused_kernel on this device (16, else 8; -1 = unusable)
  bf16* zfuse_bf = nullptr;   // bf16 copy of z_fuse (A operand of cond_proj)
  float* thr_default = nullptr;
  // attention: tensor map over the packed qkv buffer, rebuilt when (pointer, T, hidden) changes
  const void* attn_qkv = nullptr; int attn_T = 0, attn_hidden = 0; CUtensorMap attn_tm{};
};

// ------------------------------------------------------------------------------------------ profiling
enum { CLS_PRE = 0, CLS_STEM, CLS_POOL, CLS_CONV, CLS_GEMM_TEXT, CLS_ATTN, CLS_LN, CLS_HEAD, CLS_MISC, CLS_COUNT };
struct ProfScope {       // counts the launch; in profiling mode brackets it with CUDA events on the launch stream
  mmdx_engine* e; cudaEvent_t a = nullptr, b = nullptr;
  explicit ProfScope(mmdx_engine* e_);
  ~ProfScope();
};

// ------------------------------------------------------------------------------------------ tensor maps
static int make_tmap(mmdx_engine* e, CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = e->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gd, gs, bx, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu box %u %u %u %u",
             (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
             (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1],
             rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return fail(buf);
  }
  return 0;
}

// cudaFuncSetAttribute and cluster occupancy are per DEVICE: one process may drive several GPUs (inference(device="cuda:1")
// after "cuda:0"), so the "done once" flags of the launch helpers are indexed by the current device.
static int cur_dev() {
  int d = 0;
  cudaGetDevice(&d);
  return d & 63;
}

// Every kernel goes through here: programmatic dependent launch (see ptx.cuh pdl_wait / pdl_trigger) so the next
// kernel's CTAs are scheduled and run their prologue while this one drains; clusters of `cluster` CTAs along x.
// MMDX_PDL=0 turns the attribute off (plain stream order) for A/B timing.
static bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* v = getenv("MMDX_PDL"); on = (v && atoi(v) == 0) ? 0 : 1; }
  return on == 1;
}
template <typename... KArgs>
static cudaError_t launch_typed(const cudaLaunchConfig_t& cfg, void (*kfn)(KArgs...), KArgs... args) {
  return cudaLaunchKernelEx(&cfg, kfn, args...);
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kfn)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (cluster > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)cluster; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at; cfg.numAttrs = n;
  return launch_typed<KArgs...>(cfg, kfn, std::forward<Args>(args)...);
}

ProfScope::ProfScope(mmdx_engine* e_) : e(e_) {
  e->launches++;
  if (!e->profiling) return;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a, e->cur_stream);
}
ProfScope::~ProfScope() {
  if (!a) return;
  cudaEventRecord(b, e->cur_stream);
  e->prof.push_back({e->cur_cls, a, b});
}

// Tile shape (BN) and CTA-group size (cg) of one GEMM: minimise  waves x BN x penalty, where a wave is one tile per
// CTA group (148 / cg groups) - i.e. the per-SM time of the persistent schedule including its last, partly empty
// wave.  Penalties (from tools/opbench.py on B200): single-CTA tiles and BN < 192 are shared-memory-bandwidth bound
// (A + B operand reads plus the TMA fill exceed 128 B/clk), CTA pairs with BN >= 192 are not.
// e.g. N = 768 (attention output / FFN2, M = 32768): BN 256 -> 6 waves x 256, BN 192 -> 7 waves x 192 (10 % less).
static void pick_tile_shape(mmdx_engine* e, long long m_tiles, int N, int bn_req, int* bn_out, int* cg_out) {
  struct Cand { int bn, cg; double pen; };
  const Cand cands[] = {{256, 2, 1.00}, {192, 2, 1.03}, {128, 2, 1.15}, {256, 1, 1.35}, {128, 1, 1.35}, {64, 1, 1.60}};
  double best = -1.0; *bn_out = 0; *cg_out = 1;
  if (bn_req == 0 && e->force_bn != 0 && N % e->force_bn == 0) bn_req = e->force_bn;      // MMDX_BN: tile-shape sweeps
  for (const Cand& c : cands) {
    if (N % c.bn != 0) continue;
    if (bn_req != 0 && c.bn != bn_req) continue;
    if (c.cg == 2 && m_tiles < 2) continue;
    if (e->force_cg == 1 && c.cg != 1) continue;
    if (e->force_cg == 2 && c.cg != 2 && m_tiles >= 2 && N % 128 == 0) continue;
    const long long groups = e->num_sms / c.cg;
    const long long tiles = ((m_tiles + c.cg - 1) / c.cg) * (N / c.bn);
    const long long waves = (tiles + groups - 1) / groups;
    const double cost = (double)waves * c.bn * c.pen;
    if (best < 0 || cost < best) { best = cost; *bn_out = c.bn; *cg_out = c.cg; }
  }
}

// Epilogue wiring.  Must be called after build_gemm/build_conv/build_stem (needs the tile geometry).
// bf16 outputs with 16-byte-aligned pitches go through the TMA-staged epilogue; fp32 outputs use direct stores.
static int fill_epilogue(mmdx_engine* e, GemmLaunch& g, const float* bias, const bf16* residual, long long ldr, void* out,
                         long long ldc, int act, int out_f32) {
  GemmParams& p = g.p;
  p.bias = bias; p.residual = residual; p.ldr = ldr; p.out = out; p.ldc = ldc; p.act = act; p.out_f32 = out_f32;
  const bool out_aligned = (ldc % 8 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  const bool res_aligned = residual && (ldr % 8 == 0) && (reinterpret_cast<uintptr_t>(residual) % 16 == 0);
  p.epi_mode = (!out_f32 && out_aligned) ? EPI_TMA : EPI_DIRECT;
  const uint64_t N = (uint64_t)p.n_tiles * g.bn;
  const uint64_t dims[4] = {N, (uint64_t)p.OW, (uint64_t)p.OH, (uint64_t)p.NB};
  p.tmC = p.tmB; p.tmR = p.tmB; p.tmI = g.cg == 2 ? e->tm_ident_half : e->tm_ident;     // always valid descriptors
  p.res_blocks = 0;
  if (p.epi_mode == EPI_TMA) {
    const uint32_t box[4] = {(uint32_t)kEpiCW, (uint32_t)p.Wb, (uint32_t)p.Hb, (uint32_t)p.Nb};
    const uint64_t cs[3] = {(uint64_t)ldc * 2, (uint64_t)p.OW * ldc * 2, (uint64_t)p.OH * p.OW * ldc * 2};
    TRY(make_tmap(e, &p.tmC, out, 4, dims, cs, box, 64));
  }
  if (res_aligned && g.bk == 64) {      // residual added by the tensor core: D += R * I (see gemm_tcgen05.cuh)
    const uint32_t rbox[4] = {64, (uint32_t)p.Wb, (uint32_t)p.Hb, (uint32_t)p.Nb};
    const uint64_t rs[3] = {(uint64_t)ldr * 2, (uint64_t)p.OW * ldr * 2, (uint64_t)p.OH * p.OW * ldr * 2};
    TRY(make_tmap(e, &p.tmR, residual, 4, dims, rs, rbox, 128));
    p.res_blocks = g.bn / 64;
  } else if (residual) {
    REQUIRE(p.epi_mode == EPI_DIRECT, "unaligned residual needs the direct epilogue (fp32 or unaligned output)");
  }
  // CTA-pair kernels trade one ring stage for a second staging buffer per epilogue group.  Measured on one box
  // (tools/opbench.py, MMDX_EB=1|2): that pays when the main loop is short and the tile is epilogue/HBM-bound
  // (layer1/2 conv3: -10 / -4 us) or the epilogue is heavy (erf-GELU of FFN1: -8 us) and costs 2-4 us on the other long-K
  // text GEMMs.
  g.eb = (e->epi_bufs == 1 || e->epi_bufs == 2) ? e->epi_bufs : ((p.num_k_blocks + p.res_blocks <= 6 || act == ACT_GELU) ? 2 : 1);
  return 0;
}

// Row LayerNorm fused behind a plain GEMM whose N tiles cover whole rows (see GemmParams::item_major).  Legal for
// N = 768 with 256-wide CTA-pair tiles, tensor-core residual and the TMA epilogue.  OFF by default: measured on B200 at
// M = 32768 (tools/opbench.py gemm_ln) the fused launch is slower than GEMM + layernorm_kernel under programmatic
// dependent launch (K = 768: 71.7 vs 66.6 us, K = 3072: 153.6 vs 130.0 us) - 128 items on 74 CTA pairs quantise to
// 6 tile times instead of 5.25 (7 waves of 192-wide tiles), and the eight epilogue warps re-reading their rows through
// an L2 that is busy feeding the operand ring take ~25 us per 128 rows, which stalls the MMA warp two tiles later.
// MMDX_LNFUSE=1 turns it on wherever the shape allows (parity-tested in tests/test_ops_gpu.py).
static bool ln_fusable(const mmdx_engine* e, int M, int N) {
  (void)e;
  if (N != 768 || M < 256) return false;
  static int mode = -2;
  if (mode == -2) { const char* v = getenv("MMDX_LNFUSE"); mode = v ? atoi(v) : 0; }
  return mode == 1;
}
static int fuse_ln(GemmLaunch& g, const float* gamma, const float* beta, float eps, bf16* ln_out, long long ld_ln) {
  GemmParams& p = g.p;
  REQUIRE(!g.c64 && g.bn == 256 && g.cg == 2 && g.bk == 64 && p.n_tiles == 3 && p.epi_mode == EPI_TMA && p.res_blocks > 0 &&
              p.Wb == 128 && p.Hb == 1 && p.Nb == 1,
          "fused LayerNorm needs a plain 768-wide CTA-pair GEMM with residual");
  REQUIRE(ld_ln % 8 == 0 && p.ldc % 8 == 0, "fused LayerNorm needs 16-byte aligned rows");
  g.ln = 3; g.eb = 1;
  p.item_major = 1; p.ln_eps = eps; p.ln_gamma = gamma; p.ln_beta = beta; p.ln_out = ln_out; p.ld_ln = ld_ln;
  return 0;
}

// C[M,N] = A[M,K] * W[N,K]^T
static int build_gemm(mmdx_engine* e, GemmLaunch& g, const bf16* A, long long lda, const bf16* Wt, int M, int N, int K,
                      int bn_req) {
  REQUIRE(K % 64 == 0 && N % 64 == 0 && M > 0, "gemm needs K%64==0, N%64==0");
  REQUIRE(lda % 8 == 0, "gemm lda must be a multiple of 8 elements");
  GemmParams& p = g.p;
  memset(&p, 0, sizeof p);
  g.ln = 0; g.c64 = false; g.b64 = 0;
  const long long m_tiles = (M + 127) / 128;
  pick_tile_shape(e, m_tiles, N, bn_req, &g.bn, &g.cg);
  REQUIRE(g.bn != 0, "no BN tile divides N");
  g.bk = 64;
  const uint64_t dims[4] = {(uint64_t)K, (uint64_t)M, 1, 1};
  const uint64_t str[3] = {(uint64_t)lda * 2, (uint64_t)lda * 2 * (uint64_t)M, (uint64_t)lda * 2 * (uint64_t)M};
  const uint32_t box[4] = {64, 128, 1, 1};
  TRY(make_tmap(e, &p.tmA[0], A, 4, dims, str, box, 128));
  for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
  const uint64_t bd[2] = {(uint64_t)K, (uint64_t)N};
  const uint64_t bs[1] = {(uint64_t)K * 2};
  const uint32_t bb[2] = {64, (uint32_t)(g.bn / g.cg)};
  TRY(make_tmap(e, &p.tmB, Wt, 2, bd, bs, bb, 128));
  p.num_k_blocks = K / 64; p.kb_per_tap = K / 64; p.a_box_bytes = 128 * 64 * 2;
  REQUIRE(K / 64 <= 255, "K too large");
  p.tap_kb[0] = (unsigned char)(K / 64);
  p.n_tiles = N / g.bn; p.cg = g.cg; p.num_tiles = (int)(((m_tiles + g.cg - 1) / g.cg) * p.n_tiles);
  p.tiles_w = (int)m_tiles; p.tiles_h = 1;
  p.Wb = 128; p.Hb = 1; p.Nb = 1; p.OW = M; p.OH = 1; p.NB = 1;
  return 0;
}

// choose a spatial tile (Wb x Hb pixels x Nb images, <=128 rows) that minimises the number of tiles
static void pick_tile(int OW, int OH, int NB, int& Wb, int& Hb, int& Nb) {
  long long best = -1; Wb = 1; Hb = 1; Nb = 1;
  for (int w = 1; w <= OW && w <= 128; ++w) {
    for (int h = 1; h <= OH && w * h <= 128; ++h) {
      int n = 128 / (w * h);
      if (n > NB) n = NB;
      if (n < 1) continue;
      const long long tiles = (long long)((OW + w - 1) / w) * ((OH + h - 1) / h) * ((NB + n - 1) / n);
      // tie-break: wider rows (longer contiguous runs for TMA)
      if (best < 0 || tiles < best || (tiles == best && w > Wb)) { best = tiles; Wb = w; Hb = h; Nb = n; }
    }
  }
}

// NHWC conv as implicit GEMM.  in [NB,H,W,Cin], w [Cout][k*k][Cin], out [NB,OH,OW,Cout]
static int build_conv(mmdx_engine* e, GemmLaunch& g, const bf16* in, int NB, int H, int W, int Cin, const bf16* w,
                      int Cout, int k, int stride) {
  REQUIRE((k == 1 || k == 3) && (stride == 1 || stride == 2), "conv supports k in {1,3}, stride in {1,2}");
  REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "conv channels must be multiples of 64");
  const int pad = k / 2;
  const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
  if (k == 1 && stride == 1) return build_gemm(e, g, in, Cin, w, NB * H * W, Cout, Cin, 0);
  GemmParams& p = g.p;
  memset(&p, 0, sizeof p);
  int Wb, Hb, Nb;
  pick_tile(OW, OH, NB, Wb, Hb, Nb);
  const long long m_tiles = (long long)((OW + Wb - 1) / Wb) * ((OH + Hb - 1) / Hb) * ((NB + Nb - 1) / Nb);
  pick_tile_shape(e, m_tiles, Cout, 0, &g.bn, &g.cg);
  REQUIRE(g.bn != 0, "no BN tile divides Cout");
  g.bk = 64;
  const uint32_t box[4] = {64, (uint32_t)Wb, (uint32_t)Hb, (uint32_t)Nb};
  if (stride == 1) {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
    const uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    TRY(make_tmap(e, &p.tmA[0], in, 4, dims, str, box, 128));
    for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
  } else {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const int Hh = (H - ph + 1) / 2, Wh = (W - pw + 1) / 2;
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)(Wh > 0 ? Wh : 1), (uint64_t)(Hh > 0 ? Hh : 1), (uint64_t)NB};
        const uint64_t str[3] = {(uint64_t)2 * Cin * 2, (uint64_t)2 * W * Cin * 2, (uint64_t)H * W * Cin * 2};
        TRY(make_tmap(e, &p.tmA[ph * 2 + pw], in + ((size_t)ph * W + pw) * Cin, 4, dims, str, box, 128));
      }
  }
  int t = 0;
  for (int r = 0; r < k; ++r)
    for (int s = 0; s < k; ++s, ++t) {
      if (stride == 1) { p.tap_map[t] = 0; p.tap_dh[t] = (signed char)(r - pad); p.tap_dw[t] = (signed char)(s - pad); }
      else {
        // input index = 2*o + r - pad ; parity view index = floor((2*o + r - pad)/2)
        const int rr = r - pad, sr = s - pad;
        const int phh = ((rr % 2) + 2) % 2, pww = ((sr % 2) + 2) % 2;
        p.tap_map[t] = (signed char)(phh * 2 + pww);
        p.tap_dh[t] = (signed char)((rr - phh) / 2);
        p.tap_dw[t] = (signed char)((sr - pww) / 2);
      }
    }
  const uint64_t bd[2] = {(uint64_t)k * k * Cin, (uint64_t)Cout};
  const uint64_t bs[1] = {(uint64_t)k * k * Cin * 2};
  const uint32_t bb[2] = {64, (uint32_t)(g.bn / g.cg)};
  TRY(make_tmap(e, &p.tmB, w, 2, bd, bs, bb, 128));
  p.kb_per_tap = Cin / 64; p.num_k_blocks = k * k * p.kb_per_tap; p.a_box_bytes = Wb * Hb * Nb * 64 * 2;
  for (int i = 0; i < k * k; ++i) p.tap_kb[i] = (unsigned char)(Cin / 64);
  p.n_tiles = Cout / g.bn; p.cg = g.cg; p.num_tiles = (int)(((m_tiles + g.cg - 1) / g.cg) * p.n_tiles);
  p.tiles_w = (OW + Wb - 1) / Wb; p.tiles_h = (OH + Hb - 1) / Hb;
  p.Wb = Wb; p.Hb = Hb; p.Nb = Nb; p.OW = OW; p.OH = OH; p.NB = NB;
  return 0;
}

// conv3 (1x1, on t2 [NB,OH,OW,Cmid]) + the block's downsample conv (1x1 stride s, on x [NB,H,W,Cin]) as ONE implicit GEMM
// whose K loop walks two "taps" over two tensors: out = t2 * W3^T + x[::s, ::s] * Wd^T (+ b3 + bd, ReLU in the epilogue).
// The downsample output (as large as the block output) is never written or read back, and the shortcut add is a
// longer K loop instead of residual MMAs.  w = [Cout][Cmid + Cin] (Bottleneck::c3ds).
static int build_c3ds(mmdx_engine* e, GemmLaunch& g, const bf16* t2, int NB, int OH, int OW, int Cmid, const bf16* x, int H,
                      int W, int Cin, int stride, const bf16* w, int Cout) {
  REQUIRE(Cmid % 64 == 0 && Cin % 64 == 0 && Cout % 64 == 0 && (stride == 1 || stride == 2), "conv3+downsample shape");
  REQUIRE(OH == (H - 1) / stride + 1 && OW == (W - 1) / stride + 1, "conv3+downsample geometry");
  GemmParams& p = g.p;
  memset(&p, 0, sizeof p);
  g.ln = 0; g.c64 = false; g.b64 = 0;
  int Wb, Hb, Nb;
  pick_tile(OW, OH, NB, Wb, Hb, Nb);
  const long long m_tiles = (long long)((OW + Wb - 1) / Wb) * ((OH + Hb - 1) / Hb) * ((NB + Nb - 1) / Nb);
  pick_tile_shape(e, m_tiles, Cout, 0, &g.bn, &g.cg);
  REQUIRE(g.bn != 0, "no BN tile divides Cout");
  g.bk = 64;
  const uint32_t box[4] = {64, (uint32_t)Wb, (uint32_t)Hb, (uint32_t)Nb};
  {
    const uint64_t dims[4] = {(uint64_t)Cmid, (uint64_t)OW, (uint64_t)OH, (uint64_t)NB};
    const uint64_t str[3] = {(uint64_t)Cmid * 2, (uint64_t)OW * Cmid * 2, (uint64_t)OH * OW * Cmid * 2};
    TRY(make_tmap(e, &p.tmA[0], t2, 4, dims, str, box, 128));
  }
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)OW, (uint64_t)OH, (uint64_t)NB};
    const uint64_t str[3] = {(uint64_t)stride * Cin * 2, (uint64_t)stride * W * Cin * 2, (uint64_t)H * W * Cin * 2};
    TRY(make_tmap(e, &p.tmA[1], x, 4, dims, str, box, 128));
  }
  p.tmA[2] = p.tmA[0]; p.tmA[3] = p.tmA[0];
  p.tap_map[0] = 0; p.tap_map[1] = 1;
  p.tap_kb[0] = (unsigned char)(Cmid / 64); p.tap_kb[1] = (unsigned char)(Cin / 64);
  const int K = Cmid + Cin;
  const uint64_t bd[2] = {(uint64_t)K, (uint64_t)Cout};
  const uint64_t bs[1] = {(uint64_t)K * 2};
  const uint32_t bb[2] = {64, (uint32_t)(g.bn / g.cg)};
  TRY(make_tmap(e, &p.tmB, w, 2, bd, bs, bb, 128));
  p.kb_per_tap = Cmid / 64; p.num_k_blocks = K / 64; p.a_box_bytes = Wb * Hb * Nb * 64 * 2;
  p.n_tiles = Cout / g.bn; p.cg = g.cg; p.num_tiles = (int)(((m_tiles + g.cg - 1) / g.cg) * p.n_tiles);
  p.tiles_w = (OW + Wb - 1) / Wb; p.tiles_h = (OH + Hb - 1) / Hb;
  p.Wb = Wb; p.Hb = Hb; p.Nb = Nb; p.OW = OW; p.OH = OH; p.NB = NB;
  return 0;
}

// 3x3 stride-1 64->64 conv through the halo-tile kernel (conv3x3_c64_tcgen05.cuh)
static bool c64_applicable(int Cin, int Cout, int k, int stride, const void* residual) {
  static int off = -1;
  if (off < 0) { const char* v = getenv("MMDX_C64"); off = (v && atoi(v) == 0) ? 1 : 0; }
  return !off && Cin == 64 && Cout == 64 && k == 3 && stride == 1 && residual == nullptr;
}
static int build_c64(mmdx_engine* e, GemmLaunch& g, const bf16* in, int NB, int H, int W, const bf16* w, const float* bias,
                     bf16* out, int act) {
  REQUIRE(act == ACT_NONE || act == ACT_RELU, "c64 conv supports no activation or ReLU");
  g.c64 = true;
  C64Params& p = g.c;
  memset(&p, 0, sizeof p);
  const uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
  const uint64_t str[3] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128};
  const uint32_t box[4] = {64, C64_HALO_W, C64_HALO_H, 1};
  TRY(make_tmap(e, &p.tmA, in, 4, dims, str, box, 128));
  const uint64_t wd[2] = {576, 64};
  const uint64_t ws[1] = {576 * 2};
  const uint32_t wb[2] = {64, 64};
  TRY(make_tmap(e, &p.tmW, w, 2, wd, ws, wb, 128));
  p.bias = bias; p.out = out; p.NB = NB; p.H = H; p.W = W;
  p.tiles_w = (W + 7) / 8; p.tiles_h = (H + 15) / 16; p.num_tiles = NB * p.tiles_w * p.tiles_h;
  p.relu = act == ACT_RELU;
  return 0;
}
static int launch_c64(mmdx_engine* e, const C64Params& p, cudaStream_t s) {
  static bool attr_set_[64] = {};
  bool& attr_set = attr_set_[cur_dev()];
  if (!attr_set) {
    CK(cudaFuncSetAttribute(conv3x3_c64_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C64_SMEM));
    attr_set = true;
  }
  const int grid = p.num_tiles < e->num_sms ? p.num_tiles : e->num_sms;
  ProfScope _ps(e);
  CK(launch_k(conv3x3_c64_tcgen05_kernel, dim3(grid), dim3(C64_THREADS), C64_SMEM, s, 1, p));
  CK(cudaGetLastError());
  return 0;
}

// Fused layer-1 bottleneck (bneck64_tcgen05.cuh): t1 [NB,H,W,64] -> y [NB,H,W,256] and the next block's conv1 output.
static bool b64_enabled() {
  static int on = -1;
  if (on < 0) { const char* v = getenv("MMDX_B64"); on = (v && atoi(v) == 0) ? 0 : 1; }
  return on == 1;
}
// g.b64 variant codes: 1 = <0,false,2>  2 = <64,false,2>  3 = <128,false,1>  4 = <64,true,1> (shortcut = downsample of x)
static bool c3ds_enabled() {
  static int on = -1;
  if (on < 0) { const char* v = getenv("MMDX_C3DS"); on = (v && atoi(v) == 0) ? 0 : 1; }
  return on == 1;
}
static int build_b64(mmdx_engine* e, GemmLaunch& g, const bf16* t1, const bf16* res, bf16* y, bf16* t1n, int NB, int H, int W,
                     const ConvW& c2, const ConvW& c3, const ConvW* c1n, const ConvW* ds, const bf16* x) {
  REQUIRE(c2.cin == 64 && c2.cout == 64 && c2.k == 3 && c2.stride == 1 && c3.cin == 64 && c3.cout == 256 && c3.k == 1,
          "fused bottleneck: 64 -3x3-> 64 -1x1-> 256");
  REQUIRE(!c1n || (c1n->cin == 256 && c1n->k == 1 && c1n->stride == 1 && (c1n->cout == 64 || c1n->cout == 128) && t1n),
          "fused bottleneck: the next conv1 must be 256 -1x1-> 64|128");
  REQUIRE(!ds || (ds->cin == 64 && ds->cout == 256 && ds->k == 1 && ds->stride == 1 && x && c1n && c1n->cout == 64),
          "fused bottleneck: in-kernel downsample needs 64 -1x1-> 256 and a 64-wide next conv1");
  REQUIRE(ds || res, "fused bottleneck: shortcut tensor missing");
  g.c64 = false; g.ln = 0;
  g.b64 = ds ? 4 : (!c1n ? 1 : (c1n->cout == 64 ? 2 : 3));
  Bneck64Params& p = g.b;
  memset(&p, 0, sizeof p);
  const uint64_t st64[3] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128};
  const uint64_t d64[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
  {
    const uint32_t box[4] = {64, B64_HALO_W, B64_HALO_H, 1};
    TRY(make_tmap(e, &p.tmA, t1, 4, d64, st64, box, 128));
  }
  {
    const uint64_t d[2] = {576, 64}; const uint64_t st[1] = {576 * 2}; const uint32_t bx[2] = {64, 32};
    TRY(make_tmap(e, &p.tmW2, c2.w, 2, d, st, bx, 128));
  }
  {
    const uint64_t d[2] = {64, 256}; const uint64_t st[1] = {64 * 2}; const uint32_t bx[2] = {64, 128};
    TRY(make_tmap(e, &p.tmW3, c3.w, 2, d, st, bx, 128));
    p.tmWd = p.tmW3;
    if (ds) TRY(make_tmap(e, &p.tmWd, ds->w, 2, d, st, bx, 128));
  }
  p.tmW1 = p.tmW3;
  if (c1n) {
    const uint64_t d[2] = {256, (uint64_t)c1n->cout}; const uint64_t st[1] = {256 * 2};
    const uint32_t bx[2] = {64, (uint32_t)c1n->cout / 2};
    TRY(make_tmap(e, &p.tmW1, c1n->w, 2, d, st, bx, 128));
  }
  {
    const uint64_t dims[4] = {256, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
    const uint64_t str[3] = {512, (uint64_t)W * 512, (uint64_t)H * W * 512};
    const uint32_t box[4] = {64, 8, 16, 1};
    TRY(make_tmap(e, &p.tmY, y, 4, dims, str, box, 128));
    p.tmR = p.tmY; p.tmX = p.tmY;
    if (ds) TRY(make_tmap(e, &p.tmX, x, 4, d64, st64, box, 128));
    else TRY(make_tmap(e, &p.tmR, res, 4, dims, str, box, 128));
  }
  p.b2 = c2.bias; p.b3 = c3.bias; p.b1 = c1n ? c1n->bias : c3.bias; p.bd = ds ? ds->bias : c3.bias; p.t1n = t1n;
  p.NB = NB; p.H = H; p.W = W;
  p.tiles_w = (W + 7) / 8; p.tiles_h = (H + 15) / 16; p.num_tiles = NB * p.tiles_w * p.tiles_h;
  p.num_items = (p.num_tiles + 1) / 2;
  return 0;
}
template <int C1, bool DS, int NH>
static int launch_b64_inst(mmdx_engine* e, const Bneck64Params& p, cudaStream_t s, int reverse) {
  static bool attr_set_[64] = {};
  static int max_clusters_[64] = {};
  bool& attr_set = attr_set_[cur_dev()];
  int& max_clusters = max_clusters_[cur_dev()];
  auto* kfn = bneck64_tcgen05_kernel<C1, DS, NH>;
  constexpr int SMEM = B64Smem<C1, DS, NH>::TOTAL;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * 64, 1, 1); cfg.blockDim = dim3(B64_THREADS, 1, 1); cfg.dynamicSmemBytes = SMEM; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaOccupancyMaxActiveClusters(&max_clusters, kfn, &cfg));
    REQUIRE(max_clusters > 0, "no CTA pair fits on this device");
    attr_set = true;
  }
  int pairs = p.num_items < e->num_sms / 2 ? p.num_items : e->num_sms / 2;
  if (pairs > max_clusters) pairs = max_clusters;
  ProfScope _ps(e);
  Bneck64Params prm = p;
  prm.reverse = reverse;
  CK(launch_k(kfn, dim3((unsigned)pairs * 2), dim3(B64_THREADS), SMEM, s, 2, prm));
  CK(cudaGetLastError());
  return 0;
}

// conv3 (+ shortcut, ReLU) of a bottleneck and conv1 (+ ReLU) of the next block as one launch (gemm2_tcgen05.cuh).
// MMDX_DUAL=0 keeps them as two launches.
static bool dual_enabled() {
  static int on = -1;
  if (on < 0) { const char* v = getenv("MMDX_DUAL"); on = (v && atoi(v) == 0) ? 0 : 1; }
  return on == 1;
}
static bool dual_applicable(const mmdx_engine* e, long long M, int K1, int N1, int N2) {
  if (!dual_enabled() || K1 % 64 != 0 || N1 % 128 != 0 || N2 % 128 != 0) return false;
  const int bn = (N2 % 256 == 0 && N1 % 256 == 0) ? 256 : 128;
  if (N1 % bn != 0 || N2 % bn != 0) return false;
  // item-major schedule: one item = 256 rows on one CTA pair; below ~2 items per pair the coarse items quantise worse
  // than the two separate launches (layer 4 at B = 256: 49 items on 74 pairs)
  return (M + 255) / 256 >= 2 * (e->num_sms / 2);
}
static int build_dual(mmdx_engine* e, GemmLaunch& g, const bf16* t2, int K1, const bf16* w3, const float* b3, const bf16* res,
                      bf16* y, int N1, const bf16* w1n, const float* b1n, bf16* t1n, int N2, long long M) {
  REQUIRE(K1 % 64 == 0 && N1 % 128 == 0 && N2 % 128 == 0 && M > 0 && M < (1ll << 31), "conv3 + conv1 shapes");
  g.ln = 0; g.c64 = false; g.b64 = 0;
  const int bn = (N2 % 256 == 0 && N1 % 256 == 0) ? 256 : 128;
  g.g2 = bn;
  Gemm2Params& p = g.d;
  memset(&p, 0, sizeof p);
  auto rows_map = [&](CUtensorMap* m, const void* base, int cols, int box_cols, int swz) -> int {
    const uint64_t dims[4] = {(uint64_t)cols, (uint64_t)M, 1, 1};
    const uint64_t str[3] = {(uint64_t)cols * 2, (uint64_t)cols * 2 * (uint64_t)M, (uint64_t)cols * 2 * (uint64_t)M};
    const uint32_t box[4] = {(uint32_t)box_cols, 128, 1, 1};
    return make_tmap(e, m, base, 4, dims, str, box, swz);
  };
  auto w_map = [&](CUtensorMap* m, const void* base, int n, int k) -> int {
    const uint64_t d[2] = {(uint64_t)k, (uint64_t)n};
    const uint64_t st[1] = {(uint64_t)k * 2};
    const uint32_t bx[2] = {64, (uint32_t)(bn / 2)};
    return make_tmap(e, m, base, 2, d, st, bx, 128);
  };
  TRY(rows_map(&p.tmA1, t2, K1, 64, 128));
  TRY(w_map(&p.tmB1, w3, N1, K1));
  TRY(rows_map(&p.tmR, res, N1, 64, 128));
  p.tmI = e->tm_ident_half;
  TRY(rows_map(&p.tmC1, y, N1, kEpiCW, 64));
  TRY(rows_map(&p.tmA2, y, N1, 64, 128));
  TRY(w_map(&p.tmB2, w1n, N2, N1));
  TRY(rows_map(&p.tmC2, t1n, N2, kEpiCW, 64));
  p.bias1 = b3; p.bias2 = b1n;
  p.kb1 = K1 / 64; p.nt1 = N1 / bn; p.kb2 = N1 / 64; p.nt2 = N2 / bn;
  const long long m_tiles = (M + 127) / 128;
  p.num_items = (int)((m_tiles + 1) / 2);
  return 0;
}
template <int BN>
static int launch_dual_inst(mmdx_engine* e, const Gemm2Params& p, cudaStream_t s, int reverse) {
  static bool attr_set_[64] = {};
  static int max_clusters_[64] = {};
  bool& attr_set = attr_set_[cur_dev()];
  int& max_clusters = max_clusters_[cur_dev()];
  auto* kfn = gemm2_tcgen05_kernel<BN>;
  constexpr int SMEM = GemmSmem<BN, 64, 2, 2, 1, 0>::TOTAL;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * 64, 1, 1); cfg.blockDim = dim3(kGemmThreads, 1, 1); cfg.dynamicSmemBytes = SMEM; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaOccupancyMaxActiveClusters(&max_clusters, kfn, &cfg));
    REQUIRE(max_clusters > 0, "no CTA pair fits on this device");
    attr_set = true;
  }
  int pairs = p.num_items < e->num_sms / 2 ? p.num_items : e->num_sms / 2;
  if (pairs > max_clusters) pairs = max_clusters;
  ProfScope _ps(e);
  Gemm2Params prm = p;
  prm.reverse = reverse;
  CK(launch_k(kfn, dim3((unsigned)pairs * 2), dim3(kGemmThreads), SMEM, s, 2, prm));
  CK(cudaGetLastError());
  return 0;
}

// direction of this launch on its branch (and flip it for the next one)
static int next_direction(mmdx_engine* e, cudaStream_t s) {
  if (!e->zigzag) return 0;
  int& z = (s == e->text_stream || e->cur_cls == CLS_GEMM_TEXT || e->cur_cls == CLS_ATTN || e->cur_cls == CLS_LN) ? e->zz_txt : e->zz_img;
  const int d = z;
  z ^= 1;
  return d;
}

template <int BN, int BK, int CG, int EB, int RES, int LN = 0, int LNF = 0>
static int launch_inst(const GemmLaunch& g, int groups, cudaStream_t s, int reverse) {
  static bool attr_set_[64] = {};
  bool& attr_set = attr_set_[cur_dev()];
  auto* kfn = gemm_tcgen05_kernel<BN, BK, CG, EB, RES, LN, LNF>;
  constexpr int SMEM = GemmSmem<BN, BK, CG, EB, RES, LNF>::TOTAL;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  if (CG == 2) {     // CTA pairs: clusters of two along x
    // The schedule is static and persistent: every cluster must be resident at once, or the stragglers run as a
    // second wave.  Not every SM can be paired (GPCs with an odd number of enabled SMs), so ask the driver.
    static int max_clusters_[64] = {};
    int& max_clusters = max_clusters_[cur_dev()];
    if (max_clusters == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)groups * CG, 1, 1);
      cfg.blockDim = dim3(kGemmThreads, 1, 1);
      cfg.dynamicSmemBytes = SMEM;
      cfg.stream = s;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      CK(cudaOccupancyMaxActiveClusters(&n, kfn, &cfg));
      REQUIRE(n > 0, "no CTA pair fits on this device");
      max_clusters = n;
      if (getenv("MMDX_DEBUG"))
        fprintf(stderr, "mmdx: gemm<BN %d, CG %d, EB %d, RES %d> %d stages, max active clusters %d\n", BN, CG, EB, RES,
                GemmSmem<BN, BK, CG, EB, RES, LNF>::STAGES, n);
    }
    if (groups > max_clusters) groups = max_clusters;
  }
  GemmParams prm = g.p;
  prm.reverse = reverse;
  CK(launch_k(kfn, dim3((unsigned)groups * CG), dim3(kGemmThreads), SMEM, s, CG, prm));
  CK(cudaGetLastError());
  return 0;
}

// Instantiations: the residual variants carry the identity block; every variant takes as many ring stages as fit.
template <int BN, int CG>
static int launch_bn(const GemmLaunch& g, int groups, cudaStream_t s, int rv) {
  const bool res = g.p.res_blocks > 0;
  if (g.p.lnf) {       // folded LayerNorm (text encoder): one variant per role, with the staging buffers that role gets
    if (g.p.lnf_diag) return launch_inst<BN, 64, CG, 1, 1, 0, 3>(g, groups, s, rv);          // attention output, FFN2
    if (g.p.lnf_post) return launch_inst<BN, 64, CG, 2, 0, 0, 2>(g, groups, s, rv);          // FFN1 (erf-GELU)
    return launch_inst<BN, 64, CG, 1, 0, 0, 1>(g, groups, s, rv);                            // QKV
  }
  if (g.eb == 2) return res ? launch_inst<BN, 64, CG, 2, 1>(g, groups, s, rv) : launch_inst<BN, 64, CG, 2, 0>(g, groups, s, rv);
  return res ? launch_inst<BN, 64, CG, 1, 1>(g, groups, s, rv) : launch_inst<BN, 64, CG, 1, 0>(g, groups, s, rv);
}

static int launch_gemm(mmdx_engine* e, const GemmLaunch& g, cudaStream_t s) {
  const int rv = next_direction(e, s);
  switch (g.b64) {
    case 1: return launch_b64_inst<0, false, 2>(e, g.b, s, rv);
    case 2: return launch_b64_inst<64, false, 2>(e, g.b, s, rv);
    case 3: return launch_b64_inst<128, false, 1>(e, g.b, s, rv);
    case 4: return launch_b64_inst<64, true, 1>(e, g.b, s, rv);
    default: break;
  }
  if (g.c64) return launch_c64(e, g.c, s);
  if (g.g2 == 256) return launch_dual_inst<256>(e, g.d, s, rv);
  if (g.g2 == 128) return launch_dual_inst<128>(e, g.d, s, rv);
  const int max_groups = e->num_sms / g.cg;
  const int groups = g.p.num_tiles < max_groups ? g.p.num_tiles : max_groups;
  ProfScope _ps(e);
  if (g.cg == 2) {
    if (g.ln == 3) return launch_inst<256, 64, 2, 1, 1, 3>(g, groups, s, rv);
    switch (g.bn) {
      case 256: return launch_bn<256, 2>(g, groups, s, rv);
      case 192: return launch_bn<192, 2>(g, groups, s, rv);
      case 128: return launch_bn<128, 2>(g, groups, s, rv);
    }
    return fail("mmdx: bad BN for a CTA pair");
  }
  switch (g.bn) {
    case 256: return launch_bn<256, 1>(g, groups, s, rv);
    case 128: return launch_bn<128, 1>(g, groups, s, rv);
    case 64: return launch_bn<64, 1>(g, groups, s, rv);
  }
  return fail("mmdx: bad BN");
}

// ------------------------------------------------------------------------------------------ stem (conv1 + maxpool)
// conv1 weight [64,3,7,7] (x optional per-output-channel scale = folded BN) -> bf16 in the layout the stem kernel
// keeps resident: per filter row r, a [64 x 32] K-major operand (k = 4*s + c; tap s = 7 and channel c = 3 are zero)
// stored as no-swizzle core matrices [k-chunk j][n-group g][row i][8 elements].
extern "C" int mmdx_pack_stem_weights(const float* w_oihw, const float* scale, uint16_t* out_bf16) {
  if (!w_oihw || !out_bf16) return 1;
  for (int r = 0; r < 7; ++r)
    for (int j = 0; j < 4; ++j)
      for (int g = 0; g < 8; ++g)
        for (int i = 0; i < 8; ++i)
          for (int el = 0; el < 8; ++el) {
            const int o = 8 * g + i, k = 8 * j + el, s = k >> 2, c = k & 3;
            float v = 0.f;
            if (s < 7 && c < 3) v = w_oihw[((o * 3 + c) * 7 + r) * 7 + s] * (scale ? scale[o] : 1.0f);
            const bf16 b = __float2bfloat16(v);
            out_bf16[(((r * 4 + j) * 8 + g) * 8 + i) * 8 + el] = *reinterpret_cast<const uint16_t*>(&b);
          }
  return 0;
}

static int plan_stem(mmdx_engine* e, StemParams& p, const bf16* in_pad, int B, int H, int W, const bf16* w,
                     const float* bias, bf16* out, int pool) {
  memset(&p, 0, sizeof p);
  int hp, wp;
  mmdx_padded_dims(H, W, &hp, &wp);
  p.in_pad = in_pad; p.w = w; p.bias = bias; p.out = out; p.B = B; p.hp = hp; p.wp = wp; p.pool = pool;
  p.OH = (H - 1) / 2 + 1; p.OW = (W - 1) / 2 + 1;
  p.PH = (p.OH - 1) / 2 + 1; p.PW = (p.OW - 1) / 2 + 1;
  const int pitch = wp * 8;
  const int fixed = STEM_W_BYTES + STEM_ROWBUF_BYTES + 512 + 1024;
  const int budget = (232448 - fixed) / 2 - 16 - STEM_SLACK;             // bytes of image rows per strip buffer
  const int max_rows_in = budget / pitch;
  REQUIRE(max_rows_in >= (pool ? 11 : 7), "image too wide for the stem kernel's shared-memory strip");
  const int out_rows = pool ? p.PH : p.OH, out_cols = pool ? p.PW : p.OW;
  p.cols_per_block = pool ? (out_cols < 63 ? out_cols : 63) : (out_cols < 128 ? out_cols : 128);
  p.col_blocks = (out_cols + p.cols_per_block - 1) / p.cols_per_block;
  // rows per strip: fewest waves x conv rows per strip (halo rows are recomputed)
  int max_r = pool ? (max_rows_in - 7) / 4 : (max_rows_in - 5) / 2;
  if (max_r > out_rows) max_r = out_rows;
  long long best = -1; int best_r = 1;
  for (int r = 1; r <= max_r; ++r) {
    const long long units = (long long)B * ((out_rows + r - 1) / r) * p.col_blocks;
    const long long waves = (units + e->num_sms - 1) / e->num_sms;
    const long long cost = waves * (pool ? 2 * r + 1 : r);
    if (best < 0 || cost < best || (cost == best && r > best_r)) { best = cost; best_r = r; }
  }
  p.rows_per_strip = best_r;
  p.strips = (out_rows + best_r - 1) / best_r;
  p.num_units = B * p.strips * p.col_blocks;
  const int rows_in = pool ? 4 * best_r + 7 : 2 * best_r + 5;
  p.in_buf_bytes = ((16 + rows_in * pitch + STEM_SLACK + 1023) / 1024) * 1024;
  return 0;
}

static int launch_stem(mmdx_engine* e, const StemParams& p, cudaStream_t s) {
  const int smem = 1024 + STEM_W_BYTES + STEM_ROWBUF_BYTES + 2 * p.in_buf_bytes + 512;
  REQUIRE(smem <= 232448, "stem strip does not fit in shared memory");
  static int attr_smem_[64] = {};
  int& attr_smem = attr_smem_[cur_dev()];
  if (smem > attr_smem) {
    CK(cudaFuncSetAttribute(stem_pool_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_smem = 232448;
  }
  const int grid = p.num_units < e->num_sms ? p.num_units : e->num_sms;
  ProfScope _ps(e);
  CK(launch_k(stem_pool_tcgen05_kernel, dim3(grid), dim3(STEM_THREADS), smem, s, 1, p));
  CK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ create / weights
extern "C" int mmdx_create(const mmdx_config* cfg, mmdx_engine** out) {
  REQUIRE(cfg && out, "null argument");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(std::string("mmdx: no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(ce));
  REQUIRE(cfg->device >= 0 && cfg->device < ndev, "bad device ordinal");
  CK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail("mmdx: this library is built for sm_100a (B200) only; found sm_" +
                                    std::to_string(prop.major) + std::to_string(prop.minor));
  std::unique_ptr<mmdx_engine> e(new mmdx_engine());
  e->cfg = *cfg;
  if (e->cfg.n_heads <= 0) e->cfg.n_heads = 12;
  if (const char* v = getenv("MMDX_KEEP_FP32")) e->cfg.keep_fp32 = atoi(v);
  if (const char* v = getenv("MMDX_MAX_PASS")) e->max_pass = atoi(v);
  e->num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
  e->encode = reinterpret_cast<EncodeTiledFn>(fn);
  {
    std::vector<bf16> id(64 * 64, __float2bfloat16(0.f));
    for (int i = 0; i < 64; ++i) id[i * 64 + i] = __float2bfloat16(1.f);
    TRY(e->ident.ensure(id.size() * 2));
    CK(cudaMemcpy(e->ident.p, id.data(), id.size() * 2, cudaMemcpyHostToDevice));
    const uint64_t d[2] = {64, 64};
    const uint64_t st[1] = {128};
    const uint32_t bx[2] = {64, 64};
    TRY(make_tmap(e.get(), &e->tm_ident, e->ident.p, 2, d, st, bx, 128));
    const uint32_t bh[2] = {64, 32};
    TRY(make_tmap(e.get(), &e->tm_ident_half, e->ident.p, 2, d, st, bh, 128));
  }
  CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&e->copy_done, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&e->copy_ready, cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) {
    CK(cudaEventCreateWithFlags(&e->slot_copy_done[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->slot_tok_done[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->slot_done[i], cudaEventDisableTiming));
  }
  CK(cudaStreamCreateWithFlags(&e->text_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&e->fork_ev, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&e->join_ev, cudaEventDisableTiming));
  if (const char* v = getenv("MMDX_STREAMS")) e->two_streams = atoi(v) != 1;
  {
    std::vector<bf16> lut(3 * 256);
    for (int c = 0; c < 3; ++c)
      for (int v = 0; v < 256; ++v)      // ToTensor (/255) then Normalize ((x - mean) / std), fp32 like the reference
        lut[c * 256 + v] = __float2bfloat16((static_cast<float>(v) / 255.0f - e->cfg.mean[c]) / e->cfg.std[c]);
    TRY(e->pre_lut.ensure(lut.size() * 2));
    CK(cudaMemcpy(e->pre_lut.p, lut.data(), lut.size() * 2, cudaMemcpyHostToDevice));
  }
  if (const char* v = getenv("MMDX_ZIGZAG")) e->zigzag = atoi(v) != 0;
  if (const char* v = getenv("MMDX_GRAPH_MAX_B")) e->graph_max_b = atoi(v);
  CK(cudaStreamCreateWithFlags(&e->graph_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&e->graph_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&e->last_done, cudaEventDisableTiming));
  e->img_ws.gen = e->txt_ws.gen = e->head_ws.gen = e->io_slot[0].gen = e->io_slot[1].gen = &e->ws_gen;
  if (const char* v = getenv("MMDX_CG")) e->force_cg = atoi(v);
  if (const char* v = getenv("MMDX_BN")) e->force_bn = atoi(v);
  if (const char* v = getenv("MMDX_EB")) e->epi_bufs = atoi(v);
  if (const char* v = getenv("MMDX_ATTN")) e->attn_force_general = (strcmp(v, "general") == 0);
  *out = e.release();
  return 0;
}

extern "C" void mmdx_destroy(mmdx_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaDeviceSynchronize();
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  if (e->copy_done) cudaEventDestroy(e->copy_done);
  if (e->copy_ready) cudaEventDestroy(e->copy_ready);
  for (int i = 0; i < 2; ++i) {
    if (e->slot_copy_done[i]) cudaEventDestroy(e->slot_copy_done[i]);
    if (e->slot_tok_done[i]) cudaEventDestroy(e->slot_tok_done[i]);
    if (e->slot_done[i]) cudaEventDestroy(e->slot_done[i]);
  }
  if (e->jpeg.st) e->jpeg.state_destroy(e->jpeg.st);
  if (e->jpeg.h) e->jpeg.destroy(e->jpeg.h);
  if (e->jpeg.lib) dlclose(e->jpeg.lib);
  for (auto& kv : e->host_graphs) {
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (kv.second.h_in) cudaFreeHost(kv.second.h_in);
    if (kv.second.h_out) cudaFreeHost(kv.second.h_out);
  }
  if (e->graph_stream) cudaStreamDestroy(e->graph_stream);
  if (e->graph_fork) cudaEventDestroy(e->graph_fork);
  if (e->last_done) cudaEventDestroy(e->last_done);
  if (e->text_stream) cudaStreamDestroy(e->text_stream);
  if (e->fork_ev) cudaEventDestroy(e->fork_ev);
  if (e->join_ev) cudaEventDestroy(e->join_ev);
  delete e;
}

extern "C" int mmdx_num_sms(mmdx_engine* e) { return e ? e->num_sms : 0; }
extern "C" int64_t mmdx_launch_count(mmdx_engine* e) { return e ? e->launches : 0; }

extern "C" int mmdx_load_tensor(mmdx_engine* e, const char* name, const float* h_data, int ndim, const int64_t* shape) {
  REQUIRE(e && name && h_data && ndim >= 0 && ndim <= 4, "bad argument");
  REQUIRE(!e->finalized, "weights already finalized");
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= (size_t)shape[i]; }
  t.data.assign(h_data, h_data + n);
  e->host[name] = std::move(t);
  return 0;
}

static const HostTensor* get(mmdx_engine* e, const std::string& k) {
  auto it = e->host.find(k);
  return it == e->host.end() ? nullptr : &it->second;
}
#define GET(var, key)                                                       \
  const HostTensor* var = get(e, key);                                      \
  if (!var) return fail(std::string("mmdx: missing weight tensor ") + (key))

template <typename T>
static int upload(mmdx_engine* e, const std::vector<T>& v, T** out) {
  const size_t bytes = (v.size() * sizeof(T) + 255) & ~size_t(255);
  REQUIRE(e->wused + bytes <= e->warena.bytes, "weight arena overflow");
  T* p = reinterpret_cast<T*>(static_cast<char*>(e->warena.p) + e->wused);
  CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  e->wused += bytes;
  *out = p;
  return 0;
}

// conv weight [Cout,Cin,k,k] + BN(eval) -> bf16 [Cout][k*k][Cin] with scale folded, fp32 bias
static int pack_conv(mmdx_engine* e, const std::string& wkey, const std::string& bnkey, int stride, ConvW* out) {
  GET(w, wkey + ".weight"); GET(g, bnkey + ".weight"); GET(b, bnkey + ".bias");
  GET(m, bnkey + ".running_mean"); GET(v, bnkey + ".running_var");
  REQUIRE(w->shape.size() == 4, "conv weight rank");
  const int cout = (int)w->shape[0], cin = (int)w->shape[1], k = (int)w->shape[2];
  std::vector<bf16> pw((size_t)cout * k * k * cin);
  std::vector<float> bias(cout);
  for (int o = 0; o < cout; ++o) {
    const float sc = g->data[o] / std::sqrt(v->data[o] + 1e-5f);     // BatchNorm2d eps 1e-5, running stats
    bias[o] = b->data[o] - m->data[o] * sc;
    for (int c = 0; c < cin; ++c)
      for (int r = 0; r < k; ++r)
        for (int s = 0; s < k; ++s)
          pw[((size_t)o * k * k + r * k + s) * cin + c] =
              __float2bfloat16(w->data[(((size_t)o * cin + c) * k + r) * k + s] * sc);
  }
  out->cin = cin; out->cout = cout; out->k = k; out->stride = stride;
  TRY(upload(e, pw, &out->w));
  TRY(upload(e, bias, &out->bias));
  return 0;
}

// conv3 | downsample concatenated along K (device to host, once): [Cout][c3.cin + ds.cin] bf16, bias b3 + bd
static int pack_c3ds(mmdx_engine* e, Bottleneck* bk) {
  const ConvW& a = bk->c3; const ConvW& d = bk->ds;
  REQUIRE(a.k == 1 && d.k == 1 && a.cout == d.cout, "conv3 / downsample shapes");
  std::vector<bf16> wa((size_t)a.cout * a.cin), wd((size_t)d.cout * d.cin), wc((size_t)a.cout * (a.cin + d.cin));
  std::vector<float> ba(a.cout), bd(d.cout);
  CK(cudaMemcpy(wa.data(), a.w, wa.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(wd.data(), d.w, wd.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ba.data(), a.bias, ba.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(bd.data(), d.bias, bd.size() * 4, cudaMemcpyDeviceToHost));
  const int K = a.cin + d.cin;
  for (int o = 0; o < a.cout; ++o) {
    memcpy(&wc[(size_t)o * K], &wa[(size_t)o * a.cin], (size_t)a.cin * 2);
    memcpy(&wc[(size_t)o * K + a.cin], &wd[(size_t)o * d.cin], (size_t)d.cin * 2);
    ba[o] += bd[o];
  }
  bk->c3ds.cin = K; bk->c3ds.cout = a.cout; bk->c3ds.k = 1; bk->c3ds.stride = d.stride;
  TRY(upload(e, wc, &bk->c3ds.w));
  TRY(upload(e, ba, &bk->c3ds.bias));
  return 0;
}

static int pack_linear(mmdx_engine* e, const std::string& key, LinW* out) {
  GET(w, key + ".weight"); GET(b, key + ".bias");
  REQUIRE(w->shape.size() == 2, "linear weight rank");
  out->nout = (int)w->shape[0]; out->nin = (int)w->shape[1];
  std::vector<bf16> pw(w->data.size());
  for (size_t i = 0; i < pw.size(); ++i) pw[i] = __float2bfloat16(w->data[i]);
  TRY(upload(e, pw, &out->w));
  TRY(upload(e, b->data, &out->bias));
  return 0;
}
static int pack_ln(mmdx_engine* e, const std::string& key, LnW* out) {
  GET(g, key + ".weight"); GET(b, key + ".bias");
  TRY(upload(e, g->data, &out->g));
  TRY(upload(e, b->data, &out->b));
  return 0;
}
static int pack_table(mmdx_engine* e, const std::string& key, bf16** out) {
  GET(w, key);
  std::vector<bf16> pw(w->data.size());
  for (size_t i = 0; i < pw.size(); ++i) pw[i] = __float2bfloat16(w->data[i]);
  return upload(e, pw, out);
}

// ---- folded LayerNorm weights.  `lin` = the GEMM's Linear (host fp32), (g, b) = the LayerNorm whose output it consumes
static int fold_consumer(mmdx_engine* e, const HostTensor* w, const HostTensor* bias, const HostTensor* g, const HostTensor* b,
                         LnFold* out) {
  const int n = (int)w->shape[0], k = (int)w->shape[1];
  REQUIRE((int)g->data.size() == k && (int)b->data.size() == k, "LayerNorm width != Linear input width");
  std::vector<bf16> wf((size_t)n * k);
  std::vector<float> c1(n), c2(n);
  for (int o = 0; o < n; ++o) {
    // W"[o][i] = gamma[i] W[o][i] - mean_i(gamma[i] W[o][i]): x W"^T = (x - mean(x)) (gamma o W)^T, the mean subtraction of
    // the LayerNorm is done by the tensor core; the epilogue only scales by rstd and adds c2 = W beta + b
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < k; ++i) {
      s1 += (double)w->data[(size_t)o * k + i] * (double)g->data[i];
      s2 += (double)w->data[(size_t)o * k + i] * (double)b->data[i];
    }
    const double m = s1 / k;
    for (int i = 0; i < k; ++i)
      wf[(size_t)o * k + i] = __float2bfloat16((float)((double)w->data[(size_t)o * k + i] * (double)g->data[i] - m));
    c1[o] = (float)s1; c2[o] = (float)(s2 + (double)bias->data[o]);
  }
  TRY(upload(e, wf, &out->wf));
  TRY(upload(e, c1, &out->c1));
  return upload(e, c2, &out->c2);
}
static int fold_residual(mmdx_engine* e, const HostTensor* bias, const HostTensor* g, const HostTensor* b, LnFold* out) {
  const int n = (int)g->data.size();
  REQUIRE((int)bias->data.size() == n && n % 64 == 0, "LayerNorm width != Linear output width");
  std::vector<bf16> gd((size_t)n * 64, __float2bfloat16(0.f));
  std::vector<float> c1(n), c2(n);
  for (int o = 0; o < n; ++o) {
    const bf16 q = __float2bfloat16(g->data[o]);
    gd[(size_t)o * 64 + (o & 63)] = q;
    c1[o] = __bfloat162float(q);
    c2[o] = bias->data[o] + b->data[o];
  }
  TRY(upload(e, gd, &out->gdiag));
  TRY(upload(e, c1, &out->c1));
  return upload(e, c2, &out->c2);
}

// ---- fp32 arena (fp32 mode)
static int upload32(mmdx_engine* e, const std::vector<float>& v, float** out) {
  const size_t bytes = (v.size() * 4 + 255) & ~size_t(255);
  REQUIRE(e->w32_used + bytes <= e->w32.bytes, "fp32 weight arena overflow");
  float* p = reinterpret_cast<float*>(static_cast<char*>(e->w32.p) + e->w32_used);
  CK(cudaMemcpy(p, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  e->w32_used += bytes;
  *out = p;
  return 0;
}
static int pack_conv32(mmdx_engine* e, const std::string& wkey, const std::string& bnkey, int stride, Conv32* out) {
  GET(w, wkey + ".weight"); GET(g, bnkey + ".weight"); GET(b, bnkey + ".bias");
  GET(m, bnkey + ".running_mean"); GET(v, bnkey + ".running_var");
  const int cout = (int)w->shape[0], cin = (int)w->shape[1], k = (int)w->shape[2];
  std::vector<float> pw((size_t)cout * k * k * cin), bias(cout);
  for (int o = 0; o < cout; ++o) {
    // y = (conv - mean) / sqrt(var + eps) * gamma + beta, folded in double and rounded once
    const double sc = (double)g->data[o] / std::sqrt((double)v->data[o] + 1e-5);
    bias[o] = (float)((double)b->data[o] - (double)m->data[o] * sc);
    for (int c = 0; c < cin; ++c)
      for (int r = 0; r < k; ++r)
        for (int s = 0; s < k; ++s)
          pw[((size_t)o * k * k + r * k + s) * cin + c] = (float)((double)w->data[(((size_t)o * cin + c) * k + r) * k + s] * sc);
  }
  out->cin = cin; out->cout = cout; out->k = k; out->stride = stride;
  TRY(upload32(e, pw, &out->w));
  return upload32(e, bias, &out->bias);
}
static int pack_linear32(mmdx_engine* e, const std::string& key, Lin32* out) {
  GET(w, key + ".weight"); GET(b, key + ".bias");
  out->nout = (int)w->shape[0]; out->nin = (int)w->shape[1];
  TRY(upload32(e, w->data, &out->w));
  return upload32(e, b->data, &out->bias);
}
static int finalize_f32(mmdx_engine* e) {
  size_t total = 0;
  for (auto& kv : e->host) total += kv.second.data.size() * 4 + 512;
  TRY(e->w32.ensure(total + (4 << 20)));
  e->w32_used = 0;
  TRY(pack_conv32(e, "image.backbone.0", "image.backbone.1", 2, &e->stem32));
  const int nblocks[4] = {3, 4, 6, 3};
  e->blocks32.clear();
  for (int li = 0; li < 4; ++li)
    for (int bi = 0; bi < nblocks[li]; ++bi) {
      const std::string pfx = "image.backbone." + std::to_string(4 + li) + "." + std::to_string(bi);
      Block32 bk;
      const int s = (bi == 0 && li > 0) ? 2 : 1;
      TRY(pack_conv32(e, pfx + ".conv1", pfx + ".bn1", 1, &bk.c1));
      TRY(pack_conv32(e, pfx + ".conv2", pfx + ".bn2", s, &bk.c2));
      TRY(pack_conv32(e, pfx + ".conv3", pfx + ".bn3", 1, &bk.c3));
      if (get(e, pfx + ".downsample.0.weight")) {
        bk.has_ds = true;
        TRY(pack_conv32(e, pfx + ".downsample.0", pfx + ".downsample.1", s, &bk.ds));
      }
      e->blocks32.push_back(bk);
    }
  TRY(pack_linear32(e, "image.proj", &e->proj_img32));
  const std::string eb = "text.encoder.embeddings.";
  { GET(t, eb + "word_embeddings.weight"); TRY(upload32(e, t->data, &e->word32)); }
  { GET(t, eb + "position_embeddings.weight"); TRY(upload32(e, t->data, &e->ptab32)); }
  { GET(t, eb + "token_type_embeddings.weight"); TRY(upload32(e, t->data, &e->ttab32)); }
  e->layers32.clear();
  for (int l = 0;; ++l) {
    const std::string p = "text.encoder.encoder.layer." + std::to_string(l) + ".";
    if (!get(e, p + "attention.self.query.weight")) break;
    Bert32 L;
    GET(q, p + "attention.self.query.weight"); GET(k, p + "attention.self.key.weight");
    GET(v, p + "attention.self.value.weight"); GET(qb, p + "attention.self.query.bias");
    GET(kb, p + "attention.self.key.bias"); GET(vb, p + "attention.self.value.bias");
    std::vector<float> w3, b3;
    for (const HostTensor* t : {q, k, v}) w3.insert(w3.end(), t->data.begin(), t->data.end());
    for (const HostTensor* t : {qb, kb, vb}) b3.insert(b3.end(), t->data.begin(), t->data.end());
    L.qkv.nin = (int)q->shape[1]; L.qkv.nout = 3 * (int)q->shape[0];
    TRY(upload32(e, w3, &L.qkv.w));
    TRY(upload32(e, b3, &L.qkv.bias));
    TRY(pack_linear32(e, p + "attention.output.dense", &L.ao));
    TRY(pack_linear32(e, p + "intermediate.dense", &L.ff1));
    TRY(pack_linear32(e, p + "output.dense", &L.ff2));
    e->layers32.push_back(L);
  }
  TRY(pack_linear32(e, "text.proj", &e->proj_txt32));
  TRY(pack_linear32(e, "fusion.fusion_mlp.0", &e->fuse32));
  e->has_f32 = true;
  return 0;
}

extern "C" int mmdx_finalize_weights(mmdx_engine* e) {
  REQUIRE(e && !e->finalized, "bad engine state");
  CK(cudaSetDevice(e->cfg.device));
  size_t total = 0;
  for (auto& kv : e->host) total += kv.second.data.size() * 4 + 512;
  // + the conv3|downsample copies (7 MB bf16), the gamma-folded QKV / FFN1 weights and diag(gamma) blocks of the text
  // encoder (~100 MB bf16: already counted above at 4 bytes per fp32 source element, which is twice their bf16 size) and
  // alignment slack
  TRY(e->warena.ensure(total + (32 << 20)));
  e->wused = 0;
  // ---- image encoder: stem
  {
    GET(w, "image.backbone.0.weight"); GET(g, "image.backbone.1.weight"); GET(b, "image.backbone.1.bias");
    GET(m, "image.backbone.1.running_mean"); GET(v, "image.backbone.1.running_var");
    REQUIRE(w->shape.size() == 4 && w->shape[0] == 64 && w->shape[1] == 3 && w->shape[2] == 7, "stem shape");
    std::vector<float> bias(64), sc(64);
    for (int o = 0; o < 64; ++o) {
      sc[o] = g->data[o] / std::sqrt(v->data[o] + 1e-5f);          // BatchNorm2d eps 1e-5, running stats
      bias[o] = b->data[o] - m->data[o] * sc[o];
    }
    e->stem.cin = 3; e->stem.cout = 64; e->stem.k = 7; e->stem.stride = 2;
    TRY(upload(e, bias, &e->stem.bias));
    std::vector<uint16_t> w2(STEM_W_BYTES / 2);
    REQUIRE(mmdx_pack_stem_weights(w->data.data(), sc.data(), w2.data()) == 0, "stem weight packing");
    uint16_t* dw2 = nullptr;
    TRY(upload(e, w2, &dw2));
    e->stem_w2 = reinterpret_cast<bf16*>(dw2);
  }
  const int nblocks[4] = {3, 4, 6, 3};
  e->blocks.clear();
  for (int li = 0; li < 4; ++li)
    for (int bi = 0; bi < nblocks[li]; ++bi) {
      const std::string pfx = "image.backbone." + std::to_string(4 + li) + "." + std::to_string(bi);
      Bottleneck bk;
      const int s = (bi == 0 && li > 0) ? 2 : 1;
      TRY(pack_conv(e, pfx + ".conv1", pfx + ".bn1", 1, &bk.c1));
      TRY(pack_conv(e, pfx + ".conv2", pfx + ".bn2", s, &bk.c2));      // v1.5: stride on the 3x3
      TRY(pack_conv(e, pfx + ".conv3", pfx + ".bn3", 1, &bk.c3));
      if (get(e, pfx + ".downsample.0.weight")) {
        bk.has_ds = true;
        TRY(pack_conv(e, pfx + ".downsample.0", pfx + ".downsample.1", s, &bk.ds));
        TRY(pack_c3ds(e, &bk));
      }
      e->blocks.push_back(bk);
    }
  e->feat_dim = e->blocks.back().c3.cout;
  TRY(pack_linear(e, "image.proj", &e->proj_img));
  e->d_img = e->proj_img.nout;
  // ---- text encoder
  {
    const std::string eb = "text.encoder.embeddings.";
    TRY(pack_table(e, eb + "word_embeddings.weight", &e->word));
    TRY(pack_table(e, eb + "position_embeddings.weight", &e->ptab));
    TRY(pack_table(e, eb + "token_type_embeddings.weight", &e->ttab));
    TRY(pack_ln(e, eb + "LayerNorm", &e->emb_ln));
    e->hidden = (int)get(e, eb + "word_embeddings.weight")->shape[1];
    e->vocab = (int)get(e, eb + "word_embeddings.weight")->shape[0];
    e->max_pos = (int)get(e, eb + "position_embeddings.weight")->shape[0];
    e->type_vocab = (int)get(e, eb + "token_type_embeddings.weight")->shape[0];
    e->layers.clear();
    for (int l = 0;; ++l) {
      const std::string p = "text.encoder.encoder.layer." + std::to_string(l) + ".";
      if (!get(e, p + "attention.self.query.weight")) break;
      BertLayerW L;
      {   // fuse Q,K,V -> [3H, H]
        GET(q, p + "attention.self.query.weight"); GET(k, p + "attention.self.key.weight");
        GET(v, p + "attention.self.value.weight"); GET(qb, p + "attention.self.query.bias");
        GET(kb, p + "attention.self.key.bias"); GET(vb, p + "attention.self.value.bias");
        const size_t n = q->data.size();
        std::vector<bf16> pw(3 * n);
        for (size_t i = 0; i < n; ++i) {
          pw[i] = __float2bfloat16(q->data[i]); pw[n + i] = __float2bfloat16(k->data[i]);
          pw[2 * n + i] = __float2bfloat16(v->data[i]);
        }
        std::vector<float> bias;
        bias.insert(bias.end(), qb->data.begin(), qb->data.end());
        bias.insert(bias.end(), kb->data.begin(), kb->data.end());
        bias.insert(bias.end(), vb->data.begin(), vb->data.end());
        L.qkv.nin = e->hidden; L.qkv.nout = 3 * e->hidden;
        TRY(upload(e, pw, &L.qkv.w));
        TRY(upload(e, bias, &L.qkv.bias));
      }
      TRY(pack_linear(e, p + "attention.output.dense", &L.ao));
      TRY(pack_ln(e, p + "attention.output.LayerNorm", &L.ln1));
      TRY(pack_linear(e, p + "intermediate.dense", &L.ff1));
      TRY(pack_linear(e, p + "output.dense", &L.ff2));
      TRY(pack_ln(e, p + "output.LayerNorm", &L.ln2));
      {   // folded LayerNorm: the LN in front of this layer (embeddings, or the previous layer's output LN) and ln1
        const std::string prev = l == 0 ? eb + "LayerNorm" : "text.encoder.encoder.layer." + std::to_string(l - 1) + ".output.LayerNorm";
        GET(pg, prev + ".weight"); GET(pb, prev + ".bias");
        GET(g1, p + "attention.output.LayerNorm.weight"); GET(b1, p + "attention.output.LayerNorm.bias");
        GET(q, p + "attention.self.query.weight"); GET(k, p + "attention.self.key.weight");
        GET(v, p + "attention.self.value.weight"); GET(qb, p + "attention.self.query.bias");
        GET(kb, p + "attention.self.key.bias"); GET(vb, p + "attention.self.value.bias");
        HostTensor w3, b3;
        for (const HostTensor* t : {q, k, v}) w3.data.insert(w3.data.end(), t->data.begin(), t->data.end());
        for (const HostTensor* t : {qb, kb, vb}) b3.data.insert(b3.data.end(), t->data.begin(), t->data.end());
        w3.shape = {3 * q->shape[0], q->shape[1]};
        TRY(fold_consumer(e, &w3, &b3, pg, pb, &L.f_qkv));
        GET(aob, p + "attention.output.dense.bias");
        TRY(fold_residual(e, aob, pg, pb, &L.f_ao));
        GET(w1, p + "intermediate.dense.weight"); GET(bb1, p + "intermediate.dense.bias");
        TRY(fold_consumer(e, w1, bb1, g1, b1, &L.f_ff1));
        GET(b2, p + "output.dense.bias");
        TRY(fold_residual(e, b2, g1, b1, &L.f_ff2));
      }
      e->layers.push_back(L);
    }
    e->n_layers = (int)e->layers.size();
    REQUIRE(e->n_layers > 0, "no BERT layers found");
    e->ffn = e->layers[0].ff1.nout;
    REQUIRE(e->hidden == 768 || e->hidden == 1024 || e->hidden == 512 || e->hidden == 256, "unsupported hidden size");
    REQUIRE(e->hidden == e->cfg.n_heads * 64, "head dim must be 64");
    TRY(pack_linear(e, "text.proj", &e->proj_txt));
    e->d_txt = e->proj_txt.nout;
  }
  // ---- fusion head
  TRY(pack_linear(e, "fusion.fusion_mlp.0", &e->fuse));
  TRY(pack_ln(e, "fusion.fusion_mlp.3", &e->fuse_ln));
  e->d_fuse = e->fuse.nout;
  REQUIRE(e->fuse.nin == e->d_img + e->d_txt, "fusion_mlp.0 input width != d_img + d_txt");
  if (get(e, "fusion.cond_proj.0.weight")) {     // optional: report-generation conditioning (training_pipeline.py:553-558)
    TRY(pack_linear(e, "fusion.cond_proj.0", &e->cond));
    REQUIRE(e->cond.nin == e->d_fuse && e->cond.nout % 64 == 0, "cond_proj.0 shape");
  }
  {
    GET(w, "fusion.disease_head.weight"); GET(b, "fusion.disease_head.bias");
    e->n_cls = (int)w->shape[0];
    TRY(upload(e, w->data, &e->head_w));
    TRY(upload(e, b->data, &e->head_b));
    std::vector<float> thr(e->n_cls, 0.5f);
    TRY(upload(e, thr, &e->thr_default));
  }
  if (e->cfg.keep_fp32) {
    TRY(finalize_f32(e));
    for (size_t l = 0; l < e->layers32.size(); ++l) { e->layers32[l].ln1 = e->layers[l].ln1; e->layers32[l].ln2 = e->layers[l].ln2; }
  }
  e->host.clear();
  e->finalized = true;
  CK(cudaDeviceSynchronize());
  return 0;
}

// ------------------------------------------------------------------------------------------ packed weight file
// The finalized weight arena (BN folded, QKV fused, bf16, kernel-ready layouts) as one file: a header, the engine's
// weight table - every dimension and every device pointer as an offset into the arena, in the fixed order of
// walk_weights - and the arena bytes.  Loading is a file read and ONE host-to-device copy: no torch modules, no
// state_dict, no re-packing (replaces the bundle rebuild of backend/api/views.py:188-258 and
// training_pipeline.py:773-796 for serving processes).
struct PackHeader {
  char magic[8];            // "MMDXPACK"
  uint32_t version;         // layout version of walk_weights
  uint32_t n_words;         // entries of the weight table (int64 each)
  uint64_t arena_bytes;
  uint64_t checksum;        // FNV-1a of table + arena
};
static const uint32_t kPackVersion = 6;

struct PackWalker {
  bool loading; char* base; std::vector<int64_t> words; size_t pos = 0; bool ok = true;
  void i(int& v) {
    if (loading) { if (pos < words.size()) v = (int)words[pos++]; else ok = false; }
    else words.push_back(v);
  }
  void b(bool& v) { int t = v ? 1 : 0; i(t); v = t != 0; }
  template <typename T> void p(T*& ptr) {
    if (loading) {
      if (pos >= words.size()) { ok = false; return; }
      const int64_t off = words[pos++];
      ptr = off < 0 ? nullptr : reinterpret_cast<T*>(base + off);
    } else {
      words.push_back(ptr ? reinterpret_cast<char*>(ptr) - base : -1);
    }
  }
  void conv(ConvW& c) { p(c.w); p(c.bias); i(c.cin); i(c.cout); i(c.k); i(c.stride); }
  void lin(LinW& l) { p(l.w); p(l.bias); i(l.nin); i(l.nout); }
  void ln(LnW& l) { p(l.g); p(l.b); }
  void fold(LnFold& f) { p(f.wf); p(f.c1); p(f.c2); p(f.gdiag); }
};
// one traversal for both directions: every field mmdx_finalize_weights sets
static void walk_weights(mmdx_engine* e, PackWalker& w) {
  w.i(e->d_img); w.i(e->d_txt); w.i(e->d_fuse); w.i(e->n_cls); w.i(e->hidden); w.i(e->n_layers); w.i(e->ffn); w.i(e->feat_dim);
  w.i(e->cfg.n_heads); w.i(e->vocab); w.i(e->max_pos); w.i(e->type_vocab);
  w.conv(e->stem); w.p(e->stem_w2);
  int nb = (int)e->blocks.size();
  w.i(nb);
  if (w.loading) e->blocks.assign(nb < 0 || nb > 64 ? 0 : nb, Bottleneck());
  for (Bottleneck& bk : e->blocks) {
    w.conv(bk.c1); w.conv(bk.c2); w.conv(bk.c3); w.b(bk.has_ds); w.conv(bk.ds); w.conv(bk.c3ds);
  }
  w.lin(e->proj_img);
  w.p(e->word); w.p(e->ptab); w.p(e->ttab); w.ln(e->emb_ln);
  int nl = (int)e->layers.size();
  w.i(nl);
  if (w.loading) e->layers.assign(nl < 0 || nl > 64 ? 0 : nl, BertLayerW());
  for (BertLayerW& L : e->layers) {
    w.lin(L.qkv); w.lin(L.ao); w.lin(L.ff1); w.lin(L.ff2); w.ln(L.ln1); w.ln(L.ln2);
    w.fold(L.f_qkv); w.fold(L.f_ao); w.fold(L.f_ff1); w.fold(L.f_ff2);
  }
  w.lin(e->proj_txt); w.lin(e->fuse); w.ln(e->fuse_ln); w.lin(e->cond);
  w.p(e->head_w); w.p(e->head_b); w.p(e->thr_default);
}
static uint64_t fnv1a(uint64_t h, const void* data, size_t n) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

extern "C" int mmdx_save_packed(mmdx_engine* e, const char* path) {
  REQUIRE(e && path, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  REQUIRE(e->finalized, "weights not finalized");
  CK(cudaSetDevice(e->cfg.device));
  PackWalker w{false, static_cast<char*>(e->warena.p)};
  walk_weights(e, w);
  std::vector<char> arena(e->wused);
  CK(cudaMemcpy(arena.data(), e->warena.p, e->wused, cudaMemcpyDeviceToHost));
  PackHeader h;
  memcpy(h.magic, "MMDXPACK", 8);
  h.version = kPackVersion; h.n_words = (uint32_t)w.words.size(); h.arena_bytes = e->wused;
  h.checksum = fnv1a(fnv1a(14695981039346656037ull, w.words.data(), w.words.size() * 8), arena.data(), arena.size());
  FILE* f = fopen(path, "wb");
  REQUIRE(f != nullptr, "cannot open the packed weight file for writing");
  const bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(w.words.data(), 8, w.words.size(), f) == w.words.size() &&
                  fwrite(arena.data(), 1, arena.size(), f) == arena.size();
  const bool closed = fclose(f) == 0;
  REQUIRE(ok && closed, "short write to the packed weight file");
  return 0;
}

extern "C" int mmdx_load_packed(mmdx_engine* e, const char* path) {
  REQUIRE(e && path, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  REQUIRE(!e->finalized, "weights already finalized");
  CK(cudaSetDevice(e->cfg.device));
  FILE* f = fopen(path, "rb");
  REQUIRE(f != nullptr, "cannot open the packed weight file");
  PackHeader h;
  std::vector<int64_t> words;
  std::vector<char> arena;
  bool ok = fread(&h, sizeof h, 1, f) == 1 && memcmp(h.magic, "MMDXPACK", 8) == 0 && h.version == kPackVersion &&
            h.n_words < (1u << 20) && h.arena_bytes < (1ull << 36);
  if (ok) {
    words.resize(h.n_words); arena.resize(h.arena_bytes);
    ok = fread(words.data(), 8, words.size(), f) == words.size() && fread(arena.data(), 1, arena.size(), f) == arena.size();
  }
  fclose(f);
  REQUIRE(ok, "not a packed weight file of this layout version (or truncated)");
  REQUIRE(h.checksum == fnv1a(fnv1a(14695981039346656037ull, words.data(), words.size() * 8), arena.data(), arena.size()),
          "packed weight file checksum mismatch");
  for (int64_t v : words) REQUIRE(v < (int64_t)h.arena_bytes, "packed weight table entry out of range");
  TRY(e->warena.ensure(h.arena_bytes + (1 << 20)));
  CK(cudaMemcpy(e->warena.p, arena.data(), arena.size(), cudaMemcpyHostToDevice));
  e->wused = h.arena_bytes;
  const int heads_cfg = e->cfg.n_heads;
  PackWalker w{true, static_cast<char*>(e->warena.p), std::move(words)};
  walk_weights(e, w);
  REQUIRE(w.ok && w.pos == w.words.size(), "packed weight table does not match this build's layout");
  REQUIRE(e->n_layers == (int)e->layers.size() && e->n_layers > 0 && !e->blocks.empty(), "packed weight table is inconsistent");
  REQUIRE(heads_cfg == 0 || heads_cfg == e->cfg.n_heads, "engine was created for a different head count than the file");
  REQUIRE(e->hidden == e->cfg.n_heads * 64, "head dim must be 64");
  e->host.clear();
  e->finalized = true;
  CK(cudaDeviceSynchronize());
  return 0;
}

extern "C" int mmdx_dims(mmdx_engine* e, int32_t out[8]) {
  REQUIRE(e && e->finalized, "weights not finalized");
  out[0] = e->d_img; out[1] = e->d_txt; out[2] = e->d_fuse; out[3] = e->n_cls; out[4] = e->hidden; out[5] = e->n_layers;
  out[6] = e->cond.w ? e->cond.nout : 0; out[7] = e->max_pos;
  return 0;
}

extern "C" int mmdx_table_sizes(mmdx_engine* e, int32_t out[3]) {
  REQUIRE(e && e->finalized && out, "weights not finalized");
  out[0] = e->vocab; out[1] = e->max_pos; out[2] = e->type_vocab;
  return 0;
}

// ------------------------------------------------------------------------------------------ preprocessing
struct PreGeom { int oh, ow, top, left, crop_h, crop_w, has_x, has_y; };

static int pre_geometry(mmdx_engine* e, int H, int W, PreGeom* g) {
  int rc = mmdx_resize_geometry(H, W, e->cfg.resize_short, e->cfg.crop, &g->oh, &g->ow, &g->top, &g->left);
  REQUIRE(rc == 0, "image too small for the crop");
  g->crop_h = e->cfg.crop > 0 ? e->cfg.crop : g->oh;
  g->crop_w = e->cfg.crop > 0 ? e->cfg.crop : g->ow;
  g->has_x = g->ow != W; g->has_y = g->oh != H;
  return 0;
}

// builds the two device coefficient tables inside `buf`
static int build_tables(mmdx_engine* e, DevBuf& buf, int H, int W, int C, const PreGeom& g, ResampleTable* tx,
                        ResampleTable* ty, PreStrip* strip) {
  std::vector<int32_t> host;
  size_t off_x[3] = {0, 0, 0}, off_y[3] = {0, 0, 0};
  int kx = 0, ky = 0;
  int maxc_x = 1, maxc_y = 1;
  std::vector<int32_t> fx, cx, fy, cy;
  auto add = [&](int in, int out, int first, int n, size_t* off, int* ks, int* maxc, std::vector<int32_t>& fo,
                 std::vector<int32_t>& co) -> int {
    const double sc = (double)in / out;
    const int ksize = (int)std::ceil(sc < 1.0 ? 1.0 : sc) * 2 + 1;
    std::vector<int32_t> f(n), c(n), w((size_t)n * ksize);
    if (mmdx_resample_coeffs(in, out, first, n, f.data(), c.data(), w.data(), (int)w.size()) != ksize) return 1;
    for (int v : c) if (v > *maxc) *maxc = v;
    off[0] = host.size(); host.insert(host.end(), f.begin(), f.end());
    off[1] = host.size(); host.insert(host.end(), c.begin(), c.end());
    off[2] = host.size(); host.insert(host.end(), w.begin(), w.end());
    *ks = ksize;
    fo = f; co = c;
    return 0;
  };
  if (g.has_x) REQUIRE(add(W, g.ow, g.left, g.crop_w, off_x, &kx, &maxc_x, fx, cx) == 0, "coefficient table");
  if (g.has_y) REQUIRE(add(H, g.oh, g.top, g.crop_h, off_y, &ky, &maxc_y, fy, cy) == 0, "coefficient table");
  if (host.empty()) host.push_back(0);
  TRY(buf.ensure(host.size() * 4));
  CK(cudaMemcpy(buf.p, host.data(), host.size() * 4, cudaMemcpyHostToDevice));
  const int* base = static_cast<const int*>(buf.p);
  *tx = ResampleTable{base + off_x[0], base + off_x[1], base + off_x[2], kx, maxc_x};
  *ty = ResampleTable{base + off_y[0], base + off_y[1], base + off_y[2], ky, maxc_y};
  // strip geometry of the tiled kernel: x-span touched by the crop window, rows staged per strip
  PreStrip st{};
  const int xs = g.has_x ? fx.front() : g.left;
  const int xe = g.has_x ? fx.back() + cx.back() : g.left + g.crop_w;
  st.xs = xs; st.span_bytes = (xe - xs) * C;
  st.in_pitch = ((st.span_bytes + 3 + 3) / 4) * 4 + 4;
  st.h_pitch = ((g.crop_w * C + 3) / 4) * 4;
  st.rows_per_block = 0;
  for (int ty_rows = 16; ty_rows >= 1; ty_rows /= 2) {
    int rows_in = 0;
    for (int oy0 = 0; oy0 < g.crop_h; oy0 += ty_rows) {
      const int oy1 = oy0 + ty_rows < g.crop_h ? oy0 + ty_rows : g.crop_h;
      const int n = g.has_y ? fy[oy1 - 1] + cy[oy1 - 1] - fy[oy0] : oy1 - oy0;
      if (n > rows_in) rows_in = n;
    }
    const size_t smem = (size_t)rows_in * (st.in_pitch + st.h_pitch) + 256 * 3 * 2 + 64;
    if (rows_in <= 256 && smem <= 160 * 1024) { st.rows_per_block = ty_rows; st.max_rows_in = rows_in; break; }
  }
  *strip = st;
  return 0;
}

static int launch_preprocess(mmdx_engine* e, const uint8_t* d_images, int B, int H, int W, int C, const PreGeom& g,
                             const ResampleTable& tx, const ResampleTable& ty, const PreStrip& st, bf16* out, int hp, int wp,
                             cudaStream_t s) {
  REQUIRE(C == 1 || C == 3, "images must have 1 or 3 channels");
  ProfScope _ps(e);
  if (st.rows_per_block > 0) {                  // strip-tiled two-pass kernel
    const size_t smem = (size_t)st.max_rows_in * (st.in_pitch + st.h_pitch) + 256 * 3 * 2 + 64;
    static bool attr_set_[64] = {};
    bool& attr_set = attr_set_[cur_dev()];
    if (!attr_set) {
      CK(cudaFuncSetAttribute(preprocess_tiled_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024 + 2048));
      CK(cudaFuncSetAttribute(preprocess_tiled_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024 + 2048));
      attr_set = true;
    }
    dim3 grid((g.crop_h + st.rows_per_block - 1) / st.rows_per_block, B), block(256);
    const size_t total = (size_t)B * H * W * C;
    const bf16* lut = static_cast<const bf16*>(e->pre_lut.p);
    if (C == 3)
      CK(launch_k(preprocess_tiled_kernel<3>, dim3(grid), dim3(block), smem, s, 1, d_images, total, H, W, tx, ty, g.crop_h, g.crop_w, g.has_x, g.has_y,
                                                          g.left, g.top, st, lut, out, hp, wp, 3, 3));
    else
      CK(launch_k(preprocess_tiled_kernel<1>, dim3(grid), dim3(block), smem, s, 1, d_images, total, H, W, tx, ty, g.crop_h, g.crop_w, g.has_x, g.has_y,
                                                          g.left, g.top, st, lut, out, hp, wp, 3, 3));
  } else {                                      // a strip does not fit in shared memory (extreme down-scaling)
    const float3 sc = make_float3(e->cfg.std[0], e->cfg.std[1], e->cfg.std[2]);
    const float3 sh = make_float3(e->cfg.mean[0], e->cfg.mean[1], e->cfg.mean[2]);
    dim3 grid((g.crop_w + 255) / 256, g.crop_h, B), block(256);
    if (C == 3)
      CK(launch_k(preprocess_kernel<3>, dim3(grid), dim3(block), 0, s, 1, d_images, B, H, W, tx, ty, g.crop_h, g.crop_w, g.has_x, g.has_y, g.left,
                                                 g.top, out, hp, wp, 3, 3, sc, sh));
    else
      CK(launch_k(preprocess_kernel<1>, dim3(grid), dim3(block), 0, s, 1, d_images, B, H, W, tx, ty, g.crop_h, g.crop_w, g.has_x, g.has_y, g.left,
                                                 g.top, out, hp, wp, 3, 3, sc, sh));
  }
  CK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ plans
static size_t al(size_t n) { return (n + 1023) & ~size_t(1023); }

static int get_image_plan(mmdx_engine* e, int B, int H, int W, int C, ImagePlan** out) {
  char key[96];
  snprintf(key, sizeof key, "%d_%d_%d_%d", B, H, W, C);
  auto it = e->img_plans.find(key);
  if (it != e->img_plans.end()) { *out = it->second.get(); return 0; }
  std::unique_ptr<ImagePlan> pl(new ImagePlan());
  pl->B = B; pl->H = H; pl->W = W; pl->C = C;
  PreGeom g;
  TRY(pre_geometry(e, H, W, &g));
  pl->crop_h = g.crop_h; pl->crop_w = g.crop_w; pl->has_x = g.has_x; pl->has_y = g.has_y; pl->off_x = g.left; pl->off_y = g.top;
  mmdx_padded_dims(g.crop_h, g.crop_w, &pl->hp, &pl->wp);
  const int IH = g.crop_h, IW = g.crop_w;
  pl->sh = (IH - 1) / 2 + 1; pl->sw = (IW - 1) / 2 + 1;
  pl->ph = (pl->sh - 1) / 2 + 1; pl->pw = (pl->sw - 1) / 2 + 1;
  // workspace layout: in_pad | pool_out/X | Y | O1 | O2 | DS   (the conv1 output never reaches memory)
  const size_t in_pad_b = al((size_t)B * pl->hp * pl->wp * 4 * 2);
  const size_t stem_b = 0;
  const size_t act_b = al((size_t)B * pl->ph * pl->pw * 256 * 2);      // largest bottleneck tensor (layer1 output)
  const size_t total = in_pad_b + stem_b + 5 * act_b;
  char* old = static_cast<char*>(e->img_ws.p);
  TRY(e->img_ws.ensure(total));
  if (old != e->img_ws.p) { e->img_plans.clear(); e->img_last = nullptr; }   // buffers moved: cached tensor maps are stale
  char* base = static_cast<char*>(e->img_ws.p);
  pl->in_pad = reinterpret_cast<bf16*>(base);
  bf16* bufs[5];
  for (int i = 0; i < 5; ++i) bufs[i] = reinterpret_cast<bf16*>(base + in_pad_b + stem_b + i * act_b);
  pl->in_pad_bytes = in_pad_b;
  TRY(build_tables(e, pl->tables, H, W, C, g, &pl->tx, &pl->ty, &pl->strip));
  pl->convs.clear();
  GemmLaunch gl;
  pl->pool_out = bufs[0];
  TRY(plan_stem(e, pl->stem_p, pl->in_pad, B, IH, IW, e->stem_w2, e->stem.bias, pl->pool_out, 1));
  bf16* x = bufs[0];
  bf16* y = bufs[1];
  bf16* o1 = bufs[2];
  bf16* o2 = bufs[3];
  bf16* ds = bufs[4];
  int h = pl->ph, w = pl->pw;
  bool have_t1 = false;                      // o1 already holds this block's conv1 output (written by the previous fused block)
  for (size_t bi = 0; bi < e->blocks.size(); ++bi) {
    const Bottleneck& bk = e->blocks[bi];
    const int s = bk.c2.stride;
    const int oh = (h - 1) / s + 1, ow = (w - 1) / s + 1;
    if (!have_t1) {
      TRY(build_conv(e, gl, x, B, h, w, bk.c1.cin, bk.c1.w, bk.c1.cout, 1, 1));
      TRY(fill_epilogue(e, gl, bk.c1.bias, nullptr, 0, o1, bk.c1.cout, ACT_RELU, 0));
      pl->convs.push_back(gl);
    }
    have_t1 = false;
    // layer-1 shape (64 -3x3-> 64 -1x1-> 256): conv2, conv3 + shortcut and the next block's conv1 (256 -> 64, or -> 128
    // for layer2.0) run as ONE kernel (bneck64_tcgen05.cuh); in layer1.0 the downsample conv of the 64-channel block
    // input joins it as 64 more K of the conv3 GEMM, so its 256-channel output is never written or read back.
    const bool b64 = b64_enabled() && c64_applicable(bk.c2.cin, bk.c2.cout, 3, s, nullptr) && bk.c3.cin == 64 &&
                     bk.c3.cout == 256 && bk.c3.k == 1;
    const Bottleneck* nx = bi + 1 < e->blocks.size() ? &e->blocks[bi + 1] : nullptr;
    const bool fuse_next = b64 && nx && nx->c1.cin == 256 && nx->c1.k == 1 && (nx->c1.cout == 64 || nx->c1.cout == 128);
    const bool fuse_ds = b64 && bk.has_ds && fuse_next && nx->c1.cout == 64 && bk.ds.cin == 64 && bk.ds.cout == 256 &&
                         bk.ds.k == 1 && bk.ds.stride == 1;
    const bf16* idt = x;
    // other blocks with a downsample conv: conv3 and the downsample run as one GEMM over [t2 | x] (build_c3ds), unless
    // the block is a fused layer-1 shape that needs the shortcut as a tensor
    const bool cat_ds = bk.has_ds && !fuse_ds && !b64 && c3ds_enabled();
    if (bk.has_ds && !fuse_ds && !cat_ds) {
      TRY(build_conv(e, gl, x, B, h, w, bk.ds.cin, bk.ds.w, bk.ds.cout, 1, s));
      TRY(fill_epilogue(e, gl, bk.ds.bias, nullptr, 0, ds, bk.ds.cout, ACT_NONE, 0));
      pl->convs.push_back(gl);
      idt = ds;
    }
    if (b64) {
      TRY(build_b64(e, gl, o1, fuse_ds ? nullptr : idt, y, fuse_next ? o2 : nullptr, B, h, w, bk.c2, bk.c3,
                    fuse_next ? &nx->c1 : nullptr, fuse_ds ? &bk.ds : nullptr, x));
      pl->convs.push_back(gl);
      gl.b64 = 0;
      if (fuse_next) { bf16* t = o1; o1 = o2; o2 = t; have_t1 = true; }
    } else {
      gl.c64 = false;
      if (c64_applicable(bk.c2.cin, bk.c2.cout, 3, s, nullptr)) {
        TRY(build_c64(e, gl, o1, B, h, w, bk.c2.w, bk.c2.bias, o2, ACT_RELU));
      } else {
        TRY(build_conv(e, gl, o1, B, h, w, bk.c2.cin, bk.c2.w, bk.c2.cout, 3, s));
        TRY(fill_epilogue(e, gl, bk.c2.bias, nullptr, 0, o2, bk.c2.cout, ACT_RELU, 0));
      }
      pl->convs.push_back(gl);
      gl.c64 = false;
      // conv3 (+ shortcut + ReLU) and the NEXT block's conv1 (+ ReLU) as one two-GEMM launch: y is read back by the
      // second GEMM while it is still in L2 instead of from DRAM a launch later (gemm2_tcgen05.cuh)
      const bool dual = !cat_ds && !bk.has_ds && nx && nx->c1.k == 1 && nx->c1.stride == 1 && nx->c1.cin == bk.c3.cout &&
                        bk.c3.k == 1 && dual_applicable(e, (long long)B * oh * ow, bk.c3.cin, bk.c3.cout, nx->c1.cout);
      if (dual) {
        TRY(build_dual(e, gl, o2, bk.c3.cin, bk.c3.w, bk.c3.bias, idt, y, bk.c3.cout, nx->c1.w, nx->c1.bias, o1, nx->c1.cout,
                       (long long)B * oh * ow));
        have_t1 = true;              // o1 now holds the next block's conv1 output
      } else if (cat_ds) {
        TRY(build_c3ds(e, gl, o2, B, oh, ow, bk.c3.cin, x, h, w, bk.ds.cin, s, bk.c3ds.w, bk.c3.cout));
        TRY(fill_epilogue(e, gl, bk.c3ds.bias, nullptr, 0, y, bk.c3.cout, ACT_RELU, 0));
      } else {
        TRY(build_conv(e, gl, o2, B, oh, ow, bk.c3.cin, bk.c3.w, bk.c3.cout, 1, 1));
        TRY(fill_epilogue(e, gl, bk.c3.bias, idt, bk.c3.cout, y, bk.c3.cout, ACT_RELU, 0));
      }
      pl->convs.push_back(gl);
      gl.g2 = 0;
    }
    bf16* t = x; x = y; y = t;
    h = oh; w = ow;
  }
  pl->last = x; pl->last_hw = h * w;
  *out = pl.get();
  e->img_plans[key] = std::move(pl);
  return 0;
}

static int ensure_head_buffers(mmdx_engine* e, int B) {
  if (B <= e->head_cap) return 0;
  const size_t f = al((size_t)B * e->feat_dim * 2), p = al((size_t)B * e->hidden * 2),
               z = al((size_t)B * (e->d_img + e->d_txt) * 2), h = al((size_t)B * e->d_fuse * 4),
               zb = al((size_t)B * e->d_fuse * 2);
  TRY(e->head_ws.ensure(f + p + z + h + zb));
  char* b = static_cast<char*>(e->head_ws.p);
  e->feats_bf = reinterpret_cast<bf16*>(b);
  e->pooled_bf = reinterpret_cast<bf16*>(b + f);
  e->zcat = reinterpret_cast<bf16*>(b + f + p);
  e->fuse_h = reinterpret_cast<float*>(b + f + p + z);
  e->zfuse_bf = reinterpret_cast<bf16*>(b + f + p + z + h);
  e->head_cap = B;
  e->head_plans.clear();
  return 0;
}

static int get_head_plan(mmdx_engine* e, int B, HeadPlan** out) {
  TRY(ensure_head_buffers(e, B));
  auto it = e->head_plans.find(B);
  if (it != e->head_plans.end()) { *out = it->second.get(); return 0; }
  std::unique_ptr<HeadPlan> pl(new HeadPlan());
  pl->B = B;
  const int dz = e->d_img + e->d_txt;
  TRY(build_gemm(e, pl->proj_img, e->feats_bf, e->feat_dim, e->proj_img.w, B, e->d_img, e->feat_dim, 0));
  TRY(fill_epilogue(e, pl->proj_img, e->proj_img.bias, nullptr, 0, e->zcat, dz, ACT_NONE, 0));
  TRY(build_gemm(e, pl->proj_txt, e->pooled_bf, e->hidden, e->proj_txt.w, B, e->d_txt, e->hidden, 0));
  TRY(fill_epilogue(e, pl->proj_txt, e->proj_txt.bias, nullptr, 0, e->zcat + e->d_img, dz, ACT_NONE, 0));
  TRY(build_gemm(e, pl->fuse, e->zcat, dz, e->fuse.w, B, e->d_fuse, dz, 0));
  TRY(fill_epilogue(e, pl->fuse, e->fuse.bias, nullptr, 0, e->fuse_h, e->d_fuse, ACT_GELU, 1));
  *out = pl.get();
  e->head_plans[B] = std::move(pl);
  return 0;
}

struct TextBufs { bf16 *hid, *qkv, *ctx, *pre, *ffn, *hid2; long long* stats; size_t stats_bytes; };

// LayerNorm folded into the text GEMMs (gemm_tcgen05.cuh, GemmParams::lnf).  MMDX_LNFOLD=0 restores the separate
// LayerNorm launches (A/B timing; also the path the fused-LN experiment uses).
static bool lnfold_enabled() {
  static int on = -1;
  if (on < 0) { const char* v = getenv("MMDX_LNFOLD"); on = (v && atoi(v) == 0) ? 0 : 1; }
  return on == 1;
}
static int apply_fold(mmdx_engine* e, GemmLaunch& g, const LnFold& f, const long long* stats, long long* stat_out, bool post,
                      bool diag, int ln_width) {
  GemmParams& p = g.p;
  REQUIRE(p.epi_mode == EPI_TMA && p.Wb == 128 && p.Hb == 1 && p.Nb == 1, "folded LayerNorm needs a plain GEMM with the TMA epilogue");
  p.lnf = 1; p.lnf_post = post ? 1 : 0; p.lnf_diag = diag ? 1 : 0;
  g.eb = post ? 2 : 1;                                               // the variants launch_bn instantiates
  p.lnf_inv_n = 1.0f / (float)ln_width; p.ln_eps = 1e-12f;          // BertLayerNorm eps
  p.lnf_c1 = f.c1; p.lnf_stats = stats; p.stat_out = reinterpret_cast<unsigned long long*>(stat_out);
  if (diag) {
    REQUIRE(p.res_blocks > 0 && f.gdiag, "folded residual needs the tensor-core residual path");
    const uint64_t d[2] = {64, (uint64_t)p.n_tiles * g.bn};
    const uint64_t st[1] = {128};
    const uint32_t bx[2] = {64, (uint32_t)(64 / g.cg)};
    TRY(make_tmap(e, &p.tmI, f.gdiag, 2, d, st, bx, 128));
  }
  return 0;
}

static int get_text_plan(mmdx_engine* e, int T, int B, TextPlan** out, TextBufs* tb) {
  const int H = e->hidden;
  const size_t hb = al((size_t)T * H * 2), qb = al((size_t)T * 3 * H * 2), fb = al((size_t)T * e->ffn * 2);
  const size_t sb = al((size_t)(1 + 2 * e->n_layers) * T * 16);      // row sums per LayerNorm input (folded LN)
  const size_t total = 4 * hb + qb + fb + sb;
  char* old = static_cast<char*>(e->txt_ws.p);
  TRY(e->txt_ws.ensure(total));
  if (old != e->txt_ws.p) e->txt_plans.clear();
  // layout depends on T; plans are keyed by T and laid out from the arena base
  char* base = static_cast<char*>(e->txt_ws.p);
  tb->hid = reinterpret_cast<bf16*>(base);
  tb->hid2 = reinterpret_cast<bf16*>(base + hb);
  tb->ctx = reinterpret_cast<bf16*>(base + 2 * hb);
  tb->pre = reinterpret_cast<bf16*>(base + 3 * hb);
  tb->qkv = reinterpret_cast<bf16*>(base + 4 * hb);
  tb->ffn = reinterpret_cast<bf16*>(base + 4 * hb + qb);
  tb->stats = reinterpret_cast<long long*>(base + 4 * hb + qb + fb);
  tb->stats_bytes = sb;
  char key[48];
  snprintf(key, sizeof key, "%d", T);
  auto it = e->txt_plans.find(key);
  if (it != e->txt_plans.end()) { *out = it->second.get(); return 0; }
  if (e->txt_plans.size() > 64) e->txt_plans.clear();
  std::unique_ptr<TextPlan> pl(new TextPlan());
  pl->T = T; pl->B = B;
  pl->folded = lnfold_enabled() && e->layers[0].f_qkv.wf != nullptr;
  pl->ln_fused = !pl->folded && ln_fusable(e, T, H);
  const int bn_ln = pl->ln_fused ? 256 : 0;
  for (int l = 0; l < e->n_layers && pl->folded; ++l) {
    // x_a = tb->hid (pre-LN input of the layer), x_b = tb->hid2 (pre-LN attention output); stats[0] = embeddings,
    // stats[1 + 2l] = x_b of layer l, stats[2 + 2l] = x_a leaving layer l
    const BertLayerW& L = e->layers[l];
    long long* s_in = tb->stats + (size_t)(l == 0 ? 0 : 2 * l) * T * 2;
    long long* s_mid = tb->stats + (size_t)(1 + 2 * l) * T * 2;
    long long* s_out = tb->stats + (size_t)(2 + 2 * l) * T * 2;
    GemmLaunch g;
    TRY(build_gemm(e, g, tb->hid, H, L.f_qkv.wf, T, 3 * H, H, 0));
    TRY(fill_epilogue(e, g, L.f_qkv.c2, nullptr, 0, tb->qkv, 3 * H, ACT_NONE, 0));
    TRY(apply_fold(e, g, L.f_qkv, s_in, nullptr, false, false, H));
    pl->gemms.push_back(g);
    TRY(build_gemm(e, g, tb->ctx, H, L.ao.w, T, H, H, 0));
    TRY(fill_epilogue(e, g, L.f_ao.c2, tb->hid, H, tb->hid2, H, ACT_NONE, 0));
    TRY(apply_fold(e, g, L.f_ao, s_in, s_mid, false, true, H));
    pl->gemms.push_back(g);
    TRY(build_gemm(e, g, tb->hid2, H, L.f_ff1.wf, T, e->ffn, H, 0));
    TRY(fill_epilogue(e, g, L.f_ff1.c2, nullptr, 0, tb->ffn, e->ffn, ACT_GELU, 0));
    TRY(apply_fold(e, g, L.f_ff1, s_mid, nullptr, true, false, H));
    pl->gemms.push_back(g);
    TRY(build_gemm(e, g, tb->ffn, e->ffn, L.ff2.w, T, H, e->ffn, 0));
    TRY(fill_epilogue(e, g, L.f_ff2.c2, tb->hid2, H, tb->hid, H, ACT_NONE, 0));
    TRY(apply_fold(e, g, L.f_ff2, s_mid, s_out, false, true, H));
    pl->gemms.push_back(g);
  }
  for (int l = 0; l < e->n_layers && !pl->folded; ++l) {
    const BertLayerW& L = e->layers[l];
    GemmLaunch g;
    TRY(build_gemm(e, g, tb->hid, H, L.qkv.w, T, 3 * H, H, 0));
    TRY(fill_epilogue(e, g, L.qkv.bias, nullptr, 0, tb->qkv, 3 * H, ACT_NONE, 0));
    pl->gemms.push_back(g);
    TRY(build_gemm(e, g, tb->ctx, H, L.ao.w, T, H, H, bn_ln));
    TRY(fill_epilogue(e, g, L.ao.bias, tb->hid, H, tb->pre, H, ACT_NONE, 0));      // + residual (pre-LN)
    if (pl->ln_fused) TRY(fuse_ln(g, L.ln1.g, L.ln1.b, 1e-12f, tb->hid2, H));
    pl->gemms.push_back(g);
    TRY(build_gemm(e, g, tb->hid2, H, L.ff1.w, T, e->ffn, H, 0));
    TRY(fill_epilogue(e, g, L.ff1.bias, nullptr, 0, tb->ffn, e->ffn, ACT_GELU, 0));
    pl->gemms.push_back(g);
    TRY(build_gemm(e, g, tb->ffn, e->ffn, L.ff2.w, T, H, e->ffn, bn_ln));
    TRY(fill_epilogue(e, g, L.ff2.bias, tb->hid2, H, tb->pre, H, ACT_NONE, 0));    // + residual (pre-LN)
    if (pl->ln_fused) TRY(fuse_ln(g, L.ln2.g, L.ln2.b, 1e-12f, tb->hid, H));
    pl->gemms.push_back(g);
  }
  *out = pl.get();
  e->txt_plans[key] = std::move(pl);
  return 0;
}

// ------------------------------------------------------------------------------------------ launch helpers
static int launch_ln(mmdx_engine* e, const bf16* x, int rows, int N, const float* g, const float* b, float eps, bf16* y,
                     cudaStream_t s) {
  constexpr int R = 4;                          // rows per warp
  const int grid = (rows + 8 * R - 1) / (8 * R);
  const int rv = next_direction(e, s);
  ProfScope _ps(e);
  switch (N) {
    case 256: CK(launch_k(layernorm_kernel<256, false, R>, dim3(grid), dim3(256), 0, s, 1, x, rows, g, b, eps, y, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, rv, 0, 0, 0, (long long*)nullptr)); break;
    case 512: CK(launch_k(layernorm_kernel<512, false, R>, dim3(grid), dim3(256), 0, s, 1, x, rows, g, b, eps, y, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, rv, 0, 0, 0, (long long*)nullptr)); break;
    case 768: CK(launch_k(layernorm_kernel<768, false, 2, 3>, dim3((rows + 15) / 16), dim3(256), 0, s, 1, x, rows, g, b, eps, y, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, rv, 0, 0, 0, (long long*)nullptr)); break;
    case 1024: CK(launch_k(layernorm_kernel<1024, false, 2>, dim3((rows + 15) / 16), dim3(256), 0, s, 1, x, rows, g, b, eps, y, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, rv, 0, 0, 0, (long long*)nullptr)); break;
    default: return fail("mmdx: layernorm width must be 256/512/768/1024");
  }
  CK(cudaGetLastError());
  return 0;
}
static int launch_embed(mmdx_engine* e, const int* ids, const int* pos, const int* tt, int rows, int N, const bf16* word,
                        const bf16* ptab, const bf16* ttab, const float* g, const float* b, float eps, bf16* y,
                        cudaStream_t s, int n_word = 0x7fffffff, int n_pos = 0x7fffffff, int n_type = 0x7fffffff,
                        long long* stats_out = nullptr) {
  constexpr int R = 2;
  const int grid = (rows + 8 * R - 1) / (8 * R);
  ProfScope _ps(e);
  switch (N) {
    case 256: CK(launch_k(layernorm_kernel<256, true, R>, dim3(grid), dim3(256), 0, s, 1, nullptr, rows, g, b, eps, y, ids, pos, tt, word, ptab, ttab, 0, n_word, n_pos, n_type, stats_out)); break;
    case 512: CK(launch_k(layernorm_kernel<512, true, R>, dim3(grid), dim3(256), 0, s, 1, nullptr, rows, g, b, eps, y, ids, pos, tt, word, ptab, ttab, 0, n_word, n_pos, n_type, stats_out)); break;
    case 768: CK(launch_k(layernorm_kernel<768, true, R>, dim3(grid), dim3(256), 0, s, 1, nullptr, rows, g, b, eps, y, ids, pos, tt, word, ptab, ttab, 0, n_word, n_pos, n_type, stats_out)); break;
    case 1024: CK(launch_k(layernorm_kernel<1024, true, R>, dim3(grid), dim3(256), 0, s, 1, nullptr, rows, g, b, eps, y, ids, pos, tt, word, ptab, ttab, 0, n_word, n_pos, n_type, stats_out)); break;
    default: return fail("mmdx: hidden width must be 256/512/768/1024");
  }
  CK(cudaGetLastError());
  return 0;
}
static int launch_attention(mmdx_engine* e, const bf16* qkv, const int* cu, int n_seq, int T, int max_len, int heads,
                            int hidden, bf16* ctx, cudaStream_t s, const long long* row_stats = nullptr) {
  REQUIRE(hidden == heads * 64, "attention head dim must be 64");
  REQUIRE(n_seq > 0 && T > 0 && max_len > 0, "bad attention batch");
  static bool attr_set_[64] = {};
  bool& attr_set = attr_set_[cur_dev()];
  if (!attr_set) {
    CK(cudaFuncSetAttribute(attention_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM));
    CK(cudaFuncSetAttribute(attention_short_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATS_SMEM));
    attr_set = true;
  }
  if (e->attn_qkv != qkv || e->attn_T != T || e->attn_hidden != hidden) {
    const uint64_t dims[2] = {(uint64_t)3 * hidden, (uint64_t)T};
    const uint64_t str[1] = {(uint64_t)3 * hidden * 2};
    const uint32_t box[2] = {64, 128};
    TRY(make_tmap(e, &e->attn_tm, qkv, 2, dims, str, box, 128));
    e->attn_qkv = qkv; e->attn_T = T; e->attn_hidden = hidden;
  }
  AttnParams p;
  p.tm = e->attn_tm; p.cu_seqlens = cu; p.ctx = ctx; p.n_seq = n_seq; p.heads = heads; p.hidden = hidden;
  p.nqb = (max_len + 127) / 128; p.num_units = n_seq * heads * p.nqb;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  p.row_stats = row_stats; p.inv_n = 1.0f / (float)hidden; p.eps = 1e-12f;
  p.reverse = next_direction(e, s);
  ProfScope _ps(e);
  if (max_len <= 128 && !e->attn_force_general) {       // one key block per sequence: the 4-deep TMEM-resident variant
    const int units = n_seq * heads;
    const int grid = units < e->num_sms ? units : e->num_sms;
    CK(launch_k(attention_short_tcgen05_kernel, dim3(grid), dim3(ATS_THREADS), ATS_SMEM, s, 1, p));
  } else {
    const int grid = p.num_units < e->num_sms ? p.num_units : e->num_sms;
    CK(launch_k(attention_tcgen05_kernel, dim3(grid), dim3(ATC_THREADS), ATC_SMEM, s, 1, p));
  }
  CK(cudaGetLastError());
  return 0;
}

__global__ void bf16_to_f32_kernel(const bf16* __restrict__ in, long long ld, int rows, int cols, float* __restrict__ out) {
  mmdx::pdl_wait();
  mmdx::pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(rows) * cols) return;
  const int r = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
  out[i] = __bfloat162float(in[r * ld + c]);
}
static int launch_cvt(mmdx_engine* e, const bf16* in, long long ld, int rows, int cols, float* out, cudaStream_t s) {
  const long long n = (long long)rows * cols;
  ProfScope _ps(e);
  CK(launch_k(bf16_to_f32_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, 1, in, ld, rows, cols, out));
  CK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ hot path
// Calls share the activation workspaces.  Work of one call is ordered on its stream(s); a call that arrives on a
// DIFFERENT stream than the previous one (torch per-thread streams) first waits for the previous call's last kernel.
static int order_begin(mmdx_engine* e, cudaStream_t s) {
  if (e->capturing) return 0;
  if (e->have_last && e->last_stream != s) CK(cudaStreamWaitEvent(s, e->last_done, 0));
  return 0;
}
static int order_end(mmdx_engine* e, cudaStream_t s) {
  if (e->capturing) return 0;
  CK(cudaEventRecord(e->last_done, s));
  e->last_stream = s; e->have_last = true;
  return 0;
}

// Images per pass of the conv stack.  Beyond ~512 images of 224 x 224 the activation tensors of a pass outgrow what the
// 126 MB L2 can hold between producer and consumer kernels and throughput falls (BASELINE config C5: 104.9 k img/s at
// B = 512, 100.5 k at B = 1024 in one pass): larger batches run as several passes over the same workspace, so
// throughput is monotone in B and the workspace stops growing.  MMDX_MAX_PASS overrides (0 = one pass).
static int max_pass_images(const mmdx_engine* e, int H, int W) {
  const int cap = e->max_pass;
  if (cap <= 0) return 1 << 30;
  const long long px = (long long)H * W;
  long long n = (long long)cap * 224 * 224 / (px > 0 ? px : 1);      // same activation volume for other image sizes
  return (int)(n < 16 ? 16 : n);
}

static int image_backbone_locked(mmdx_engine* e, const uint8_t* d_images, int B0, int B, int H, int W, int C, float* d_feats,
                                 cudaStream_t s);

static int image_encode_locked(mmdx_engine* e, const uint8_t* d_images, int B, int H, int W, int C, float* d_feats,
                               float* d_z_img, cudaStream_t s) {
  REQUIRE(e->finalized, "weights not finalized");
  REQUIRE(B > 0 && H > 0 && W > 0, "bad image batch");
  REQUIRE(C == 1 || C == 3, "images must have 1 or 3 channels");
  TRY(ensure_head_buffers(e, B));
  PreGeom pg;
  TRY(pre_geometry(e, H, W, &pg));
  const int cap = max_pass_images(e, pg.crop_h, pg.crop_w);
  for (int b0 = 0; b0 < B; b0 += cap) {
    const int bc = B - b0 < cap ? B - b0 : cap;
    TRY(image_backbone_locked(e, d_images + (size_t)b0 * H * W * C, b0, bc, H, W, C, d_feats, s));
  }
  if (e->defer_proj) return 0;
  e->cur_cls = CLS_HEAD;
  HeadPlan* hp;
  TRY(get_head_plan(e, B, &hp));
  TRY(launch_gemm(e, hp->proj_img, s));
  if (d_z_img) TRY(launch_cvt(e, e->zcat, e->d_img + e->d_txt, B, e->d_img, d_z_img, s));
  return 0;
}

// preprocess + stem + 16 bottlenecks + avgpool of images [B0, B0 + B) of the batch: rows B0.. of the feature buffers
static int image_backbone_locked(mmdx_engine* e, const uint8_t* d_images, int B0, int B, int H, int W, int C, float* d_feats,
                                 cudaStream_t s) {
  ImagePlan* pl;
  TRY(get_image_plan(e, B, H, W, C, &pl));
  PreGeom g{0, 0, pl->off_y, pl->off_x, pl->crop_h, pl->crop_w, pl->has_x, pl->has_y};
  e->cur_stream = s; e->cur_cls = CLS_PRE;
  e->zz_img = 1;                                   // preprocess / stem write forward: the first conv walks backwards
  // Another geometry may have used the arena since this plan last ran: restore the zero border + zero 4th channel.
  // A captured graph cannot know what ran before each of its replays, so it always carries the memset (and a replay
  // resets img_last, see forward_host_graphed).
  if (e->capturing || e->img_last != pl) {
    CK(cudaMemsetAsync(pl->in_pad, 0, pl->in_pad_bytes, s));
    if (!e->capturing) e->img_last = pl;
  }
  TRY(launch_preprocess(e, d_images, B, H, W, C, g, pl->tx, pl->ty, pl->strip, pl->in_pad, pl->hp, pl->wp, s));
  e->cur_cls = CLS_STEM;
  TRY(launch_stem(e, pl->stem_p, s));                // conv1 + bn1 + relu + maxpool in one kernel
  e->cur_cls = CLS_CONV;
  for (const GemmLaunch& g : pl->convs) TRY(launch_gemm(e, g, s));
  e->cur_cls = CLS_POOL;
  {
    const int n = B * (e->feat_dim / 8);
    ProfScope _ps(e);
    CK(launch_k(avgpool_kernel, dim3((n + 255) / 256), dim3(256), 0, s, 1, pl->last, B, pl->last_hw, e->feat_dim,
                e->feats_bf + (size_t)B0 * e->feat_dim, d_feats ? d_feats + (size_t)B0 * e->feat_dim : (float*)nullptr));
    CK(cudaGetLastError());
  }
  return 0;
}

static int text_encode_locked(mmdx_engine* e, const int32_t* d_ids, const int32_t* d_pos, const int32_t* d_tt,
                              const int32_t* d_cu, int B, int T, int max_len, float* d_pooled, float* d_z_txt,
                              cudaStream_t s) {
  REQUIRE(e->finalized, "weights not finalized");
  REQUIRE(B > 0 && T > 0 && max_len > 0, "bad token batch");
  REQUIRE(max_len <= e->max_pos, "sequence longer than the position embedding table (BERT: 512)");
  REQUIRE((long long)T <= (long long)B * max_len, "more packed tokens than B * max_len");
  TRY(ensure_head_buffers(e, B));
  TextPlan* pl;
  TextBufs tb;
  TRY(get_text_plan(e, T, B, &pl, &tb));
  const int H = e->hidden;
  e->cur_stream = s; e->cur_cls = CLS_LN;
  e->zz_txt = 1;                                   // the embedding kernel writes forward: the first consumer walks backwards
  if (pl->folded) {
    // LayerNorm folded into the GEMMs: the embedding kernel stores the raw sum + its row sums, every GEMM applies /
    // re-creates the LayerNorm in its epilogue, AO and FFN2 accumulate the row sums of what they write: 5 launches per layer
    CK(cudaMemsetAsync(tb.stats + (size_t)T * 2, 0, (size_t)2 * e->n_layers * T * 16, s));
    TRY(launch_embed(e, d_ids, d_pos, d_tt, T, H, e->word, e->ptab, e->ttab, e->emb_ln.g, e->emb_ln.b, 1e-12f, tb.hid, s,
                     e->vocab, e->max_pos, e->type_vocab, tb.stats));
    for (int l = 0; l < e->n_layers; ++l) {
      const long long* s_in = tb.stats + (size_t)(l == 0 ? 0 : 2 * l) * T * 2;
      e->cur_cls = CLS_GEMM_TEXT;
      TRY(launch_gemm(e, pl->gemms[4 * l + 0], s));                                       // QKV of LN(x_a)
      e->cur_cls = CLS_ATTN;
      TRY(launch_attention(e, tb.qkv, d_cu, B, T, max_len, e->cfg.n_heads, H, tb.ctx, s, s_in));   // ctx / rstd
      e->cur_cls = CLS_GEMM_TEXT;
      TRY(launch_gemm(e, pl->gemms[4 * l + 1], s));                                       // x_b = out-proj + LN(x_a)
      TRY(launch_gemm(e, pl->gemms[4 * l + 2], s));                                       // GELU(FFN1 of LN(x_b)) / rstd
      TRY(launch_gemm(e, pl->gemms[4 * l + 3], s));                                       // x_a = FFN2 + LN(x_b)
    }
    e->cur_cls = CLS_POOL;
    {
      const LnW& ln = e->layers.back().ln2;
      const long long* s_last = tb.stats + (size_t)(2 * e->n_layers) * T * 2;
      ProfScope _ps(e);
      CK(launch_k(seq_mean_pool_kernel, dim3(dim3(B, (H + 63) / 64)), dim3(256), 0, s, 1, tb.hid, d_cu, H, e->pooled_bf, H, d_pooled,
                  s_last, ln.g, ln.b, 1e-12f));
      CK(cudaGetLastError());
    }
    if (e->defer_proj) return 0;
    e->cur_cls = CLS_HEAD;
    HeadPlan* hp;
    TRY(get_head_plan(e, B, &hp));
    TRY(launch_gemm(e, hp->proj_txt, s));
    if (d_z_txt) TRY(launch_cvt(e, e->zcat + e->d_img, e->d_img + e->d_txt, B, e->d_txt, d_z_txt, s));
    return 0;
  }
  TRY(launch_embed(e, d_ids, d_pos, d_tt, T, H, e->word, e->ptab, e->ttab, e->emb_ln.g, e->emb_ln.b, 1e-12f, tb.hid, s,
                   e->vocab, e->max_pos, e->type_vocab));
  for (int l = 0; l < e->n_layers; ++l) {
    const BertLayerW& L = e->layers[l];
    e->cur_cls = CLS_GEMM_TEXT;
    TRY(launch_gemm(e, pl->gemms[4 * l + 0], s));                                         // QKV
    e->cur_cls = CLS_ATTN;
    TRY(launch_attention(e, tb.qkv, d_cu, B, T, max_len, e->cfg.n_heads, H, tb.ctx, s));
    e->cur_cls = CLS_GEMM_TEXT;
    TRY(launch_gemm(e, pl->gemms[4 * l + 1], s));                                         // out-proj + residual
    e->cur_cls = CLS_LN;
    if (!pl->ln_fused) TRY(launch_ln(e, tb.pre, T, H, L.ln1.g, L.ln1.b, 1e-12f, tb.hid2, s));
    e->cur_cls = CLS_GEMM_TEXT;
    TRY(launch_gemm(e, pl->gemms[4 * l + 2], s));                                         // FFN1 + GELU
    TRY(launch_gemm(e, pl->gemms[4 * l + 3], s));                                         // FFN2 + residual
    e->cur_cls = CLS_LN;
    if (!pl->ln_fused) TRY(launch_ln(e, tb.pre, T, H, L.ln2.g, L.ln2.b, 1e-12f, tb.hid, s));
  }
  e->cur_cls = CLS_POOL;
  {
    ProfScope _ps(e);
    CK(launch_k(seq_mean_pool_kernel, dim3(dim3(B, (H + 63) / 64)), dim3(256), 0, s, 1, tb.hid, d_cu, H, e->pooled_bf, H, d_pooled,
                (const long long*)nullptr, (const float*)nullptr, (const float*)nullptr, 0.f));
    CK(cudaGetLastError());
  }
  if (e->defer_proj) return 0;
  e->cur_cls = CLS_HEAD;
  HeadPlan* hp;
  TRY(get_head_plan(e, B, &hp));
  TRY(launch_gemm(e, hp->proj_txt, s));
  if (d_z_txt) TRY(launch_cvt(e, e->zcat + e->d_img, e->d_img + e->d_txt, B, e->d_txt, d_z_txt, s));
  return 0;
}

// K_head for the reference's request shape (B <= 2), MMDX_HEAD_FUSED read once per engine:
//   2 (default)  both projections, the fusion MLP, LayerNorm, head, sigmoid and thresholds are ONE launch behind the join
//                (I2 + T8 + F1 + O1 in one kernel);
//   1            the projections stay GEMM + bias at the ends of their branches and only F1 + O1 are one launch;
//   0            the four-launch tensor path everywhere.
// B = 1 request latency, same box, two runs each (tools/latency_b1.py): 0.610 / 0.614 ms (0), 0.611 / 0.619 (1),
// 0.623 / 0.617 (2) - the end of the chain is not what a request waits for; the launch count goes 111 -> 110 -> 108.
static void head_fused_setup(mmdx_engine* e) {
  if (e->head_cluster != 0) return;
  e->head_cluster = -1;
  { const char* v = getenv("MMDX_HEAD_FUSED"); e->head_mode = v ? atoi(v) : 2; }
  if (e->head_mode <= 0 || e->feat_dim % 8 || e->hidden % 8 || (e->d_img + e->d_txt) % 8) return;
  const size_t words = std::max<size_t>((size_t)kHeadFusedMaxB * std::max(e->feat_dim + e->hidden, e->d_img + e->d_txt), (size_t)e->d_fuse);
  if (words * 4 > 48 * 1024) return;
  for (int cs : {16, 8}) {
    if (cs > 8) {
      cudaFuncSetAttribute(head_fused_kernel<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaFuncSetAttribute(head_fused_kernel<2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs); cfg.blockDim = dim3(kHeadFusedThreads); cfg.dynamicSmemBytes = words * 4;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, head_fused_kernel<2>, &cfg) == cudaSuccess && n >= 1) { e->head_cluster = cs; break; }
  }
  cudaGetLastError();                // a refused cluster size is not an error of the engine
}
struct DeferProj {                   // scope of one mmdx_forward* call (under the engine mutex)
  mmdx_engine* e;
  DeferProj(mmdx_engine* e_, int B) : e(e_) {
    head_fused_setup(e);
    e->fused_tail = e->head_cluster > 0 && B <= kHeadFusedMaxB;
    e->defer_proj = e->fused_tail && e->head_mode >= 2;
  }
  ~DeferProj() { e->defer_proj = false; e->fused_tail = false; }
};

static int head_locked(mmdx_engine* e, int B, const float* d_thr, float* d_z_fuse, float* d_logits, float* d_probs,
                       uint8_t* d_vector, cudaStream_t s) {
  REQUIRE(e->finalized && B > 0 && B <= e->head_cap, "head called before the encoders");
  REQUIRE(d_logits && d_probs && d_vector, "null output");
  e->cur_stream = s; e->cur_cls = CLS_HEAD;
  if (e->fused_tail) {               // kernels.cuh head_fused_kernel: (projections +) fusion MLP + tail as one launch
    HeadFusedParams p{};
    p.feats = e->feats_bf; p.feat_dim = e->feat_dim; p.pooled = e->pooled_bf; p.hidden = e->hidden;
    p.w_img = e->proj_img.w; p.b_img = e->proj_img.bias; p.d_img = e->d_img;
    p.w_txt = e->proj_txt.w; p.b_txt = e->proj_txt.bias; p.d_txt = e->d_txt;
    p.w_fuse = e->fuse.w; p.b_fuse = e->fuse.bias; p.d_fuse = e->d_fuse;
    p.ln_g = e->fuse_ln.g; p.ln_b = e->fuse_ln.b; p.eps = 1e-5f;
    p.w_head = e->head_w; p.b_head = e->head_b; p.n_cls = e->n_cls;
    p.thr = d_thr ? d_thr : e->thr_default;
    p.zcat = e->zcat; p.fuse_h = e->fuse_h;
    p.z_fuse = d_z_fuse; p.logits = d_logits; p.probs = d_probs; p.vec = d_vector; p.z_fuse_bf = e->cond.w ? e->zfuse_bf : nullptr;
    p.B = B; p.do_proj = e->defer_proj ? 1 : 0;
    const int nr = B <= 1 ? 1 : 2;
    const size_t words = std::max<size_t>((size_t)nr * std::max(e->feat_dim + e->hidden, e->d_img + e->d_txt), (size_t)e->d_fuse);
    ProfScope _ps(e);
    if (nr == 1) CK(launch_k(head_fused_kernel<1>, dim3(e->head_cluster), dim3(kHeadFusedThreads), words * 4, s, e->head_cluster, p));
    else CK(launch_k(head_fused_kernel<2>, dim3(e->head_cluster), dim3(kHeadFusedThreads), words * 4, s, e->head_cluster, p));
    CK(cudaGetLastError());
    return 0;
  }
  HeadPlan* hp;
  TRY(get_head_plan(e, B, &hp));
  TRY(launch_gemm(e, hp->fuse, s));
  ProfScope _ps(e);
  CK(launch_k(head_tail_kernel, dim3(B), dim3(256), e->d_fuse * sizeof(float), s, 1, e->fuse_h, e->d_fuse, e->fuse_ln.g, e->fuse_ln.b, 1e-5f,
                                                             e->head_w, e->head_b, e->n_cls,
                                                             d_thr ? d_thr : e->thr_default, d_z_fuse, d_logits, d_probs,
                                                             d_vector, e->cond.w ? e->zfuse_bf : nullptr));
  CK(cudaGetLastError());
  return 0;
}

extern "C" int mmdx_image_encode(mmdx_engine* e, const uint8_t* d_images, int B, int H, int W, int C, float* d_feats,
                                 float* d_z_img, void* stream) {
  REQUIRE(e, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  TRY(order_begin(e, (cudaStream_t)stream));
  TRY(image_encode_locked(e, d_images, B, H, W, C, d_feats, d_z_img, (cudaStream_t)stream));
  return order_end(e, (cudaStream_t)stream);
}
extern "C" int mmdx_text_encode(mmdx_engine* e, const int32_t* d_ids, const int32_t* d_pos, const int32_t* d_tt,
                                const int32_t* d_cu, int B, int T, int max_len, float* d_pooled, float* d_z_txt,
                                void* stream) {
  REQUIRE(e, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  TRY(order_begin(e, (cudaStream_t)stream));
  TRY(text_encode_locked(e, d_ids, d_pos, d_tt, d_cu, B, T, max_len, d_pooled, d_z_txt, (cudaStream_t)stream));
  return order_end(e, (cudaStream_t)stream);
}
extern "C" int mmdx_head(mmdx_engine* e, int B, const float* d_thr, float* d_z_fuse, float* d_logits, float* d_probs,
                         uint8_t* d_vector, void* stream) {
  REQUIRE(e, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  TRY(order_begin(e, (cudaStream_t)stream));
  TRY(head_locked(e, B, d_thr, d_z_fuse, d_logits, d_probs, d_vector, (cudaStream_t)stream));
  return order_end(e, (cudaStream_t)stream);
}
// FusionTransformerModel._make_encoder_outputs (training_pipeline.py:574-578) for the batch mmdx_head has just processed:
// cond = GELU(z_fuse * Wc^T + bc), fp32 [B, n_cond * h_dec] - the "encoder output" the T5 decoder is conditioned on.
extern "C" int mmdx_cond_tokens(mmdx_engine* e, int B, float* d_cond, void* stream) {
  REQUIRE(e && d_cond, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  REQUIRE(e->finalized && e->cond.w, "the bundle has no cond_proj weights");
  REQUIRE(B > 0 && B <= e->head_cap, "cond_tokens called before mmdx_head");
  e->cur_stream = (cudaStream_t)stream; e->cur_cls = CLS_HEAD;
  GemmLaunch g;
  TRY(build_gemm(e, g, e->zfuse_bf, e->d_fuse, e->cond.w, B, e->cond.nout, e->d_fuse, 0));
  TRY(fill_epilogue(e, g, e->cond.bias, nullptr, 0, d_cond, e->cond.nout, ACT_GELU, 1));
  TRY(order_begin(e, (cudaStream_t)stream));
  TRY(launch_gemm(e, g, (cudaStream_t)stream));
  return order_end(e, (cudaStream_t)stream);
}
extern "C" int mmdx_forward(mmdx_engine* e, const uint8_t* d_images, int B, int H, int W, int C, const int32_t* d_ids,
                            const int32_t* d_pos, const int32_t* d_tt, const int32_t* d_cu, int T, int max_len,
                            const float* d_thr, float* d_logits, float* d_probs, uint8_t* d_vector, void* stream) {
  REQUIRE(e, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const bool fork = e->two_streams && !e->profiling;
  cudaStream_t ts = fork ? e->text_stream : s;
  TRY(order_begin(e, s));
  if (fork) {
    CK(cudaEventRecord(e->fork_ev, s));                 // inputs (and earlier work on `s`) are ready
    CK(cudaStreamWaitEvent(ts, e->fork_ev, 0));
  }
  DeferProj _dp(e, B);
  TRY(text_encode_locked(e, d_ids, d_pos, d_tt, d_cu, B, T, max_len, nullptr, nullptr, ts));
  if (fork) CK(cudaEventRecord(e->join_ev, ts));
  TRY(image_encode_locked(e, d_images, B, H, W, C, nullptr, nullptr, s));
  if (fork) CK(cudaStreamWaitEvent(s, e->join_ev, 0));
  TRY(head_locked(e, B, d_thr, nullptr, d_logits, d_probs, d_vector, s));
  return order_end(e, s);
}

// Request slot `slot`: H2D of its inputs, the forward, D2H of its results - enqueued, not waited for.  The image batch
// (the bulk of the bytes) crosses PCIe on the copy stream into the slot's own buffer, so with two slots in flight the
// copy of request k+1 runs under the kernels of request k; the compute itself stays ordered on `s` / the text stream
// (the activation workspaces are shared), and the D2H of request k runs under the first kernels of request k+1.
static int forward_host_submit_locked(mmdx_engine* e, int slot, const uint8_t* h_images, int B, int H, int W, int C,
                                      const int32_t* h_ids, const int32_t* h_pos, const int32_t* h_tt, const int32_t* h_cu,
                                      int T, int max_len, const float* h_thr, float* h_logits, float* h_probs,
                                      uint8_t* h_vector, cudaStream_t s) {
  REQUIRE(e->finalized, "weights not finalized");
  if (e->slot_busy[slot] && !e->capturing) { CK(cudaEventSynchronize(e->slot_done[slot])); e->slot_busy[slot] = false; }
  const size_t img_b = al((size_t)B * H * W * C), tok_b = al((size_t)T * 4), cu_b = al((size_t)(B + 1) * 4);
  const size_t out_f = al((size_t)B * e->n_cls * 4), out_u = al((size_t)B * e->n_cls), thr_b = al((size_t)e->n_cls * 4);
  TRY(e->io_slot[slot].ensure(img_b + 3 * tok_b + cu_b + 2 * out_f + out_u + thr_b));
  char* b = static_cast<char*>(e->io_slot[slot].p);
  uint8_t* d_img = reinterpret_cast<uint8_t*>(b); b += img_b;
  int32_t* d_ids = reinterpret_cast<int32_t*>(b); b += tok_b;
  int32_t* d_pos = reinterpret_cast<int32_t*>(b); b += tok_b;
  int32_t* d_tt = reinterpret_cast<int32_t*>(b); b += tok_b;
  int32_t* d_cu = reinterpret_cast<int32_t*>(b); b += cu_b;
  float* d_logits = reinterpret_cast<float*>(b); b += out_f;
  float* d_probs = reinterpret_cast<float*>(b); b += out_f;
  uint8_t* d_vec = reinterpret_cast<uint8_t*>(b); b += out_u;
  float* d_thr = reinterpret_cast<float*>(b);
  // the slot's previous request has been waited for (above), so nothing on the device still reads these buffers
  // All inputs go over the copy stream at submit time - the token arrays (a few hundred KB) FIRST, then the image batch:
  // one DMA engine serves every host-to-device copy in the order they become ready, and token copies queued behind the
  // 38 MB image copy would hold the text branch (and with it the whole GPU) back for the 0.75 ms the images take.
  if (e->capturing) {                                          // the copy stream has to join the capture of `s`
    CK(cudaEventRecord(e->graph_fork, s));
    CK(cudaStreamWaitEvent(e->copy_stream, e->graph_fork, 0));
  }
  CK(cudaMemcpyAsync(d_ids, h_ids, (size_t)T * 4, cudaMemcpyHostToDevice, e->copy_stream));
  CK(cudaMemcpyAsync(d_pos, h_pos, (size_t)T * 4, cudaMemcpyHostToDevice, e->copy_stream));
  CK(cudaMemcpyAsync(d_tt, h_tt, (size_t)T * 4, cudaMemcpyHostToDevice, e->copy_stream));
  CK(cudaMemcpyAsync(d_cu, h_cu, (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, e->copy_stream));
  if (h_thr) CK(cudaMemcpyAsync(d_thr, h_thr, (size_t)e->n_cls * 4, cudaMemcpyHostToDevice, e->copy_stream));
  CK(cudaEventRecord(e->slot_tok_done[slot], e->copy_stream));
  CK(cudaMemcpyAsync(d_img, h_images, (size_t)B * H * W * C, cudaMemcpyHostToDevice, e->copy_stream));
  CK(cudaEventRecord(e->slot_copy_done[slot], e->copy_stream));
  const bool fork = e->two_streams && !e->profiling;
  cudaStream_t ts = fork ? e->text_stream : s;
  TRY(order_begin(e, s));
  if (fork) {                                                  // the previous request on `s` still uses the shared workspaces
    CK(cudaEventRecord(e->copy_ready, s));
    CK(cudaStreamWaitEvent(ts, e->copy_ready, 0));
  }
  CK(cudaStreamWaitEvent(ts, e->slot_tok_done[slot], 0));
  if (h_thr && fork) CK(cudaStreamWaitEvent(s, e->slot_tok_done[slot], 0));
  DeferProj _dp(e, B);
  TRY(text_encode_locked(e, d_ids, d_pos, d_tt, d_cu, B, T, max_len, nullptr, nullptr, ts));
  if (fork) CK(cudaEventRecord(e->join_ev, ts));
  CK(cudaStreamWaitEvent(s, e->slot_copy_done[slot], 0));
  TRY(image_encode_locked(e, d_img, B, H, W, C, nullptr, nullptr, s));
  if (fork) CK(cudaStreamWaitEvent(s, e->join_ev, 0));
  TRY(head_locked(e, B, h_thr ? d_thr : nullptr, nullptr, d_logits, d_probs, d_vec, s));
  TRY(order_end(e, s));
  CK(cudaMemcpyAsync(h_logits, d_logits, (size_t)B * e->n_cls * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(h_probs, d_probs, (size_t)B * e->n_cls * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(h_vector, d_vec, (size_t)B * e->n_cls, cudaMemcpyDeviceToHost, s));
  if (!e->capturing) {
    CK(cudaEventRecord(e->slot_done[slot], s));
    e->slot_busy[slot] = true;
  }
  return 0;
}

// mmdx_forward_host for a small request: host copy into the staging buffer, one cudaGraphLaunch, host copy out.
// The first call of a shape runs the ordinary path (it creates plans and workspaces), the second captures the graph.
static int forward_host_graphed(mmdx_engine* e, const uint8_t* h_images, int B, int H, int W, int C, const int32_t* h_ids,
                                const int32_t* h_pos, const int32_t* h_tt, const int32_t* h_cu, int T, int max_len,
                                const float* h_thr, float* h_logits, float* h_probs, uint8_t* h_vector, cudaStream_t s,
                                bool* done, mmdx_engine::HostGraph** warm) {
  *done = false; *warm = nullptr;
  char key[96];
  snprintf(key, sizeof key, "%d_%d_%d_%d_%d_%d_%d", B, H, W, C, T, max_len, h_thr ? 1 : 0);
  if (e->host_graphs.size() > 32 && e->host_graphs.find(key) == e->host_graphs.end()) return 0;   // bounded
  mmdx_engine::HostGraph& hg = e->host_graphs[key];
  if (hg.seen < 0) return 0;                                   // capture failed once for this shape: ordinary path
  if (hg.exec && hg.gen != e->ws_gen) {
    // A workspace moved since the capture (cudaFree + cudaMalloc in DevBuf::ensure: a longer report, a larger image or a
    // bigger batch came by): every pointer and tensor map inside the graph is stale.  Drop it and start over.
    cudaGraphExecDestroy(hg.exec);
    hg.exec = nullptr; hg.seen = 0;
  }
  if (hg.seen == 0 || hg.gen != e->ws_gen) {
    // first call of this shape (or first after a move): the ordinary path creates the plans and sizes the workspaces;
    // the caller stamps hg.gen when it has run
    hg.seen = 1; *warm = &hg;
    return 0;
  }
  ++hg.seen;
  const size_t img_b = al((size_t)B * H * W * C), tok_b = al((size_t)T * 4), cu_b = al((size_t)(B + 1) * 4);
  const size_t thr_b = al((size_t)e->n_cls * 4), of = al((size_t)B * e->n_cls * 4), ou = al((size_t)B * e->n_cls);
  if (!hg.h_in) {
    hg.in_bytes = img_b + 3 * tok_b + cu_b + thr_b; hg.out_bytes = 2 * of + ou;
    CK(cudaMallocHost(&hg.h_in, hg.in_bytes));
    CK(cudaMallocHost(&hg.h_out, hg.out_bytes));
  }
  char* hi = static_cast<char*>(hg.h_in);
  uint8_t* s_img = reinterpret_cast<uint8_t*>(hi);
  int32_t* s_ids = reinterpret_cast<int32_t*>(hi + img_b);
  int32_t* s_pos = reinterpret_cast<int32_t*>(hi + img_b + tok_b);
  int32_t* s_tt = reinterpret_cast<int32_t*>(hi + img_b + 2 * tok_b);
  int32_t* s_cu = reinterpret_cast<int32_t*>(hi + img_b + 3 * tok_b);
  float* s_thr = reinterpret_cast<float*>(hi + img_b + 3 * tok_b + cu_b);
  char* ho = static_cast<char*>(hg.h_out);
  float* s_logits = reinterpret_cast<float*>(ho);
  float* s_probs = reinterpret_cast<float*>(ho + of);
  uint8_t* s_vec = reinterpret_cast<uint8_t*>(ho + 2 * of);
  if (!hg.exec) {
    // slot 0 must be idle and its buffers allocated (the warm-up call did both); capture on the engine's own stream
    // (the caller's may be the legacy default stream, which cannot be captured)
    if (e->slot_busy[0]) { CK(cudaEventSynchronize(e->slot_done[0])); e->slot_busy[0] = false; }
    const int64_t n0 = e->launches, gen0 = e->ws_gen;
    const int zt = e->zz_txt, zi = e->zz_img;
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(e->graph_stream, cudaStreamCaptureModeRelaxed));
    e->capturing = true;
    const int rc = forward_host_submit_locked(e, 0, s_img, B, H, W, C, s_ids, s_pos, s_tt, s_cu, T, max_len,
                                              h_thr ? s_thr : nullptr, s_logits, s_probs, s_vec, e->graph_stream);
    e->capturing = false;
    const cudaError_t ce = cudaStreamEndCapture(e->graph_stream, &graph);
    e->zz_txt = zt; e->zz_img = zi;
    if (rc != 0 || ce != cudaSuccess || graph == nullptr) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      hg.seen = -1000000;                                      // do not try again for this shape
      return rc != 0 ? rc : 0;
    }
    hg.launches = e->launches - n0;
    e->launches = n0;
    if (e->ws_gen != gen0) {                                   // a buffer moved while capturing: the graph is unusable
      cudaGraphDestroy(graph);
      hg.seen = 0;
      return 0;
    }
    const cudaError_t ie = cudaGraphInstantiate(&hg.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { cudaGetLastError(); hg.exec = nullptr; hg.seen = -1000000; return 0; }
    hg.gen = e->ws_gen;
  }
  if (hg.seen < 0) return 0;
  memcpy(s_img, h_images, (size_t)B * H * W * C);
  memcpy(s_ids, h_ids, (size_t)T * 4); memcpy(s_pos, h_pos, (size_t)T * 4); memcpy(s_tt, h_tt, (size_t)T * 4);
  memcpy(s_cu, h_cu, (size_t)(B + 1) * 4);
  if (h_thr) memcpy(s_thr, h_thr, (size_t)e->n_cls * 4);
  CK(cudaEventRecord(e->graph_fork, s));                       // order behind whatever the caller queued on `s`
  CK(cudaStreamWaitEvent(e->graph_stream, e->graph_fork, 0));
  TRY(order_begin(e, e->graph_stream));                        // ... and behind the previous call, whatever its stream
  CK(cudaGraphLaunch(hg.exec, e->graph_stream));
  CK(cudaStreamSynchronize(e->graph_stream));
  e->have_last = false;                                        // everything has completed
  e->img_last = nullptr;           // the replay laid its own plan's tensors over the image arena (and zeroed only its own in_pad)
  e->launches += hg.launches;
  memcpy(h_logits, s_logits, (size_t)B * e->n_cls * 4);
  memcpy(h_probs, s_probs, (size_t)B * e->n_cls * 4);
  memcpy(h_vector, s_vec, (size_t)B * e->n_cls);
  *done = true;
  return 0;
}

// nn.Embedding raises IndexError on an out-of-range index (a tokenizer / vocabulary mismatch, caller-made tokens); the
// host entry points see the ids and reject them before anything is launched.
static int validate_tokens(mmdx_engine* e, const int32_t* ids, const int32_t* pos, const int32_t* tt, const int32_t* cu,
                           int B, int T, int max_len) {
  REQUIRE(B > 0 && T > 0 && cu[0] == 0 && cu[B] == T, "cu_seqlens must run from 0 to T");
  for (int b = 0; b < B; ++b) REQUIRE(cu[b + 1] > cu[b] && cu[b + 1] - cu[b] <= max_len, "sequence lengths must be in 1..max_len");
  unsigned bad = 0;
  for (int i = 0; i < T; ++i)
    bad |= (unsigned)((unsigned)ids[i] >= (unsigned)e->vocab) | (unsigned)((unsigned)pos[i] >= (unsigned)e->max_pos) |
           (unsigned)((unsigned)tt[i] >= (unsigned)e->type_vocab);
  REQUIRE(bad == 0, "token id / position / token type outside the embedding tables (index out of range)");
  return 0;
}

extern "C" int mmdx_forward_host_submit(mmdx_engine* e, int slot, const uint8_t* h_images, int B, int H, int W, int C,
                                        const int32_t* h_ids, const int32_t* h_pos, const int32_t* h_tt,
                                        const int32_t* h_cu, int T, int max_len, const float* h_thr, float* h_logits,
                                        float* h_probs, uint8_t* h_vector, void* stream) {
  REQUIRE(e && h_images && h_ids && h_pos && h_tt && h_cu && h_logits && h_probs && h_vector, "null argument");
  REQUIRE(slot == 0 || slot == 1, "request slot must be 0 or 1");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  REQUIRE(e->finalized, "weights not finalized");
  TRY(validate_tokens(e, h_ids, h_pos, h_tt, h_cu, B, T, max_len));
  return forward_host_submit_locked(e, slot, h_images, B, H, W, C, h_ids, h_pos, h_tt, h_cu, T, max_len, h_thr, h_logits,
                                    h_probs, h_vector, (cudaStream_t)stream);
}

extern "C" int mmdx_forward_host_wait(mmdx_engine* e, int slot) {
  REQUIRE(e, "null engine");
  REQUIRE(slot == 0 || slot == 1, "request slot must be 0 or 1");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  if (!e->slot_busy[slot]) return 0;
  CK(cudaEventSynchronize(e->slot_done[slot]));
  e->slot_busy[slot] = false;
  return 0;
}

extern "C" int mmdx_forward_host(mmdx_engine* e, const uint8_t* h_images, int B, int H, int W, int C,
                                 const int32_t* h_ids, const int32_t* h_pos, const int32_t* h_tt, const int32_t* h_cu,
                                 int T, int max_len, const float* h_thr, float* h_logits, float* h_probs,
                                 uint8_t* h_vector, void* stream) {
  REQUIRE(e && h_images && h_ids && h_pos && h_tt && h_cu && h_logits && h_probs && h_vector, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  REQUIRE(e->finalized, "weights not finalized");
  TRY(validate_tokens(e, h_ids, h_pos, h_tt, h_cu, B, T, max_len));
  mmdx_engine::HostGraph* warm = nullptr;
  if (B <= e->graph_max_b && !e->profiling) {
    bool done = false;
    TRY(forward_host_graphed(e, h_images, B, H, W, C, h_ids, h_pos, h_tt, h_cu, T, max_len, h_thr, h_logits, h_probs, h_vector,
                             (cudaStream_t)stream, &done, &warm));
    if (done) return 0;
  }
  TRY(forward_host_submit_locked(e, 0, h_images, B, H, W, C, h_ids, h_pos, h_tt, h_cu, T, max_len, h_thr, h_logits, h_probs,
                                 h_vector, (cudaStream_t)stream));
  CK(cudaEventSynchronize(e->slot_done[0]));
  e->slot_busy[0] = false;
  if (warm) warm->gen = e->ws_gen;      // plans and workspaces of this shape exist under this generation: capture next time
  return 0;
}

// ------------------------------------------------------------------------------------------ fp32 mode
// The whole path in fp32 on the CUDA cores (fp32_kernels.cuh), one kernel per reference op, for the north_star's
// "1e-5 if run in fp32" bar: labels equal the reference's with no margin.  Plain launches (no programmatic dependent
// launch: these kernels do not carry the griddepcontrol handshake), one stream.
static int f32_conv(mmdx_engine* e, const float* in, int NB, int H, int W, const Conv32& c, const float* res, float* out,
                    int act, cudaStream_t s, long long ld_out = 0) {
  F32ConvParams p;
  p.in = in; p.w = c.w; p.bias = c.bias; p.res = res; p.out = out;
  p.NB = NB; p.H = H; p.W = W; p.Cin = c.cin; p.Cout = c.cout; p.k = c.k; p.stride = c.stride; p.pad = c.k / 2;
  p.OH = (H + 2 * p.pad - c.k) / c.stride + 1; p.OW = (W + 2 * p.pad - c.k) / c.stride + 1; p.act = act;
  p.ld_out = ld_out ? ld_out : c.cout;
  const long long M = (long long)NB * p.OH * p.OW;
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)((c.cout + 63) / 64));
  e->launches++;
  f32_conv_kernel<<<grid, 256, 0, s>>>(p);
  CK(cudaGetLastError());
  return 0;
}
static int f32_linear(mmdx_engine* e, const float* in, int M, const Lin32& l, const float* res, float* out, int act,
                      cudaStream_t s, long long ld_out = 0) {
  Conv32 c; c.w = l.w; c.bias = l.bias; c.cin = l.nin; c.cout = l.nout; c.k = 1; c.stride = 1;
  return f32_conv(e, in, M, 1, 1, c, res, out, act, s, ld_out);
}

extern "C" int mmdx_forward_f32(mmdx_engine* e, const uint8_t* d_images, int B, int H, int W, int C, const int32_t* d_ids,
                                const int32_t* d_pos, const int32_t* d_tt, const int32_t* d_cu, int T, int max_len,
                                const float* d_thr, float* d_feats, float* d_z_img, float* d_pooled, float* d_z_txt,
                                float* d_z_fuse, float* d_logits, float* d_probs, uint8_t* d_vector, void* stream) {
  REQUIRE(e && d_images && d_ids && d_pos && d_tt && d_cu && d_logits && d_probs && d_vector, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  REQUIRE(e->finalized && e->has_f32, "fp32 weights were not kept: create the engine with mmdx_config.keep_fp32 = 1 "
                                      "(not available for engines loaded from a packed bf16 file)");
  REQUIRE(B > 0 && T > 0 && max_len > 0 && max_len <= e->max_pos && max_len <= F32_ATT_MAXL, "bad batch");
  REQUIRE(C == 1 || C == 3, "images must have 1 or 3 channels");
  cudaStream_t s = (cudaStream_t)stream;
  TRY(order_begin(e, s));
  PreGeom g;
  TRY(pre_geometry(e, H, W, &g));
  const int IH = g.crop_h, IW = g.crop_w;
  const int sh = (IH - 1) / 2 + 1, sw = (IW - 1) / 2 + 1, ph = (sh - 1) / 2 + 1, pw = (sw - 1) / 2 + 1;
  const int Hd = e->hidden, dz = e->d_img + e->d_txt;
  // workspace: u8 crop | x | stem | 4 activation buffers | text buffers | head buffers
  const size_t u8_b = al((size_t)B * IH * IW * C), x_b = al((size_t)B * IH * IW * 3 * 4), stem_b = al((size_t)B * sh * sw * 64 * 4);
  const size_t act_b = al((size_t)B * ph * pw * 256 * 4);
  const size_t hid_b = al((size_t)T * Hd * 4), qkv_b = al((size_t)T * 3 * Hd * 4), ffn_b = al((size_t)T * e->ffn * 4);
  const size_t head_b = al((size_t)B * (e->feat_dim + Hd + dz + e->d_fuse) * 4);
  TRY(e->f32_ws.ensure(u8_b + x_b + stem_b + 4 * act_b + 4 * hid_b + qkv_b + ffn_b + head_b));
  char* b = static_cast<char*>(e->f32_ws.p);
  uint8_t* u8 = reinterpret_cast<uint8_t*>(b); b += u8_b;
  float* x = reinterpret_cast<float*>(b); b += x_b;
  float* stem = reinterpret_cast<float*>(b); b += stem_b;
  float* act[4];
  for (int i = 0; i < 4; ++i) { act[i] = reinterpret_cast<float*>(b); b += act_b; }
  float* hid = reinterpret_cast<float*>(b); b += hid_b;
  float* hid2 = reinterpret_cast<float*>(b); b += hid_b;
  float* ctx = reinterpret_cast<float*>(b); b += hid_b;
  float* pre = reinterpret_cast<float*>(b); b += hid_b;
  float* qkv = reinterpret_cast<float*>(b); b += qkv_b;
  float* ffn = reinterpret_cast<float*>(b); b += ffn_b;
  float* feats = reinterpret_cast<float*>(b);
  float* pooled = feats + (size_t)B * e->feat_dim;
  float* zcat = pooled + (size_t)B * Hd;
  float* fuse_h = zcat + (size_t)B * dz;
  // ---- image branch: integer resample (bit-exact), ToTensor + Normalize, ResNet-50
  const uint8_t* crop = d_images;
  if (g.has_x || g.has_y || g.crop_h != H || g.crop_w != W) {
    ResampleTable tx, ty;
    PreStrip strip;
    TRY(build_tables(e, e->tab_ws, H, W, C, g, &tx, &ty, &strip));
    dim3 grid((g.crop_w + 255) / 256, g.crop_h, B), block(256);
    e->launches++;
    if (C == 3) resample_u8_kernel<3><<<grid, block, 0, s>>>(d_images, B, H, W, tx, ty, g.crop_h, g.crop_w, g.has_x, g.has_y, g.left, g.top, u8);
    else resample_u8_kernel<1><<<grid, block, 0, s>>>(d_images, B, H, W, tx, ty, g.crop_h, g.crop_w, g.has_x, g.has_y, g.left, g.top, u8);
    CK(cudaGetLastError());
    crop = u8;
  }
  {
    const long long npix = (long long)B * IH * IW;
    e->launches++;
    f32_normalize_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(
        crop, npix, C, make_float3(e->cfg.mean[0], e->cfg.mean[1], e->cfg.mean[2]),
        make_float3(e->cfg.std[0], e->cfg.std[1], e->cfg.std[2]), x);
    CK(cudaGetLastError());
  }
  TRY(f32_conv(e, x, B, IH, IW, e->stem32, nullptr, stem, F32_ACT_RELU, s));
  {
    const long long n = (long long)B * ph * pw * 64;
    e->launches++;
    f32_maxpool_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(stem, B, sh, sw, 64, ph, pw, act[0]);
    CK(cudaGetLastError());
  }
  float *cur = act[0], *nxt = act[1], *t1 = act[2], *t2 = act[3];
  int h = ph, w = pw;
  for (const Block32& bk : e->blocks32) {
    const int st = bk.c2.stride, oh = (h - 1) / st + 1, ow = (w - 1) / st + 1;
    TRY(f32_conv(e, cur, B, h, w, bk.c1, nullptr, t1, F32_ACT_RELU, s));
    TRY(f32_conv(e, t1, B, h, w, bk.c2, nullptr, t2, F32_ACT_RELU, s));
    const float* idt = cur;
    if (bk.has_ds) { TRY(f32_conv(e, cur, B, h, w, bk.ds, nullptr, t1, F32_ACT_NONE, s)); idt = t1; }
    TRY(f32_conv(e, t2, B, oh, ow, bk.c3, idt, nxt, F32_ACT_RELU, s));
    float* t = cur; cur = nxt; nxt = t;
    h = oh; w = ow;
  }
  e->launches++;
  f32_avgpool_kernel<<<(B * e->feat_dim + 255) / 256, 256, 0, s>>>(cur, B, h * w, e->feat_dim, feats);
  CK(cudaGetLastError());
  TRY(f32_linear(e, feats, B, e->proj_img32, nullptr, zcat, F32_ACT_NONE, s, dz));
  // ---- text branch: BERT over packed tokens
  e->launches++;
  f32_layernorm_kernel<true><<<(T + 7) / 8, 256, 0, s>>>(nullptr, T, Hd, e->emb_ln.g, e->emb_ln.b, 1e-12f, hid, d_ids, d_pos, d_tt,
                                                          e->word32, e->ptab32, e->ttab32, e->vocab, e->max_pos, e->type_vocab);
  CK(cudaGetLastError());
  for (const Bert32& L : e->layers32) {
    TRY(f32_linear(e, hid, T, L.qkv, nullptr, qkv, F32_ACT_NONE, s));
    e->launches++;
    f32_attention_kernel<<<dim3((max_len + 3) / 4, e->cfg.n_heads, B), 128, 0, s>>>(qkv, d_cu, e->cfg.n_heads, Hd, 0.125f, ctx);
    CK(cudaGetLastError());
    TRY(f32_linear(e, ctx, T, L.ao, hid, pre, F32_ACT_NONE, s));
    e->launches++;
    f32_layernorm_kernel<false><<<(T + 7) / 8, 256, 0, s>>>(pre, T, Hd, L.ln1.g, L.ln1.b, 1e-12f, hid2, nullptr, nullptr, nullptr,
                                                             nullptr, nullptr, nullptr, 0, 0, 0);
    TRY(f32_linear(e, hid2, T, L.ff1, nullptr, ffn, F32_ACT_GELU, s));
    TRY(f32_linear(e, ffn, T, L.ff2, hid2, pre, F32_ACT_NONE, s));
    e->launches++;
    f32_layernorm_kernel<false><<<(T + 7) / 8, 256, 0, s>>>(pre, T, Hd, L.ln2.g, L.ln2.b, 1e-12f, hid, nullptr, nullptr, nullptr,
                                                             nullptr, nullptr, nullptr, 0, 0, 0);
    CK(cudaGetLastError());
  }
  e->launches++;
  f32_seq_mean_pool_kernel<<<B, 256, 0, s>>>(hid, d_cu, Hd, pooled);
  CK(cudaGetLastError());
  TRY(f32_linear(e, pooled, B, e->proj_txt32, nullptr, zcat + e->d_img, F32_ACT_NONE, s, dz));
  // ---- fusion head: Linear + GELU (fp32), then LayerNorm + disease head + sigmoid + threshold (head_tail_kernel is fp32)
  TRY(f32_linear(e, zcat, B, e->fuse32, nullptr, fuse_h, F32_ACT_GELU, s));
  e->launches++;
  head_tail_kernel<<<B, 256, e->d_fuse * sizeof(float), s>>>(fuse_h, e->d_fuse, e->fuse_ln.g, e->fuse_ln.b, 1e-5f, e->head_w, e->head_b,
                                                            e->n_cls, d_thr ? d_thr : e->thr_default, d_z_fuse, d_logits, d_probs,
                                                            d_vector, nullptr);
  CK(cudaGetLastError());
  if (d_feats) CK(cudaMemcpyAsync(d_feats, feats, (size_t)B * e->feat_dim * 4, cudaMemcpyDeviceToDevice, s));
  if (d_pooled) CK(cudaMemcpyAsync(d_pooled, pooled, (size_t)B * Hd * 4, cudaMemcpyDeviceToDevice, s));
  if (d_z_img) CK(cudaMemcpy2DAsync(d_z_img, (size_t)e->d_img * 4, zcat, (size_t)dz * 4, (size_t)e->d_img * 4, B, cudaMemcpyDeviceToDevice, s));
  if (d_z_txt) CK(cudaMemcpy2DAsync(d_z_txt, (size_t)e->d_txt * 4, zcat + e->d_img, (size_t)dz * 4, (size_t)e->d_txt * 4, B, cudaMemcpyDeviceToDevice, s));
  return order_end(e, s);
}

// ------------------------------------------------------------------------------------------ GPU JPEG decode (N3)
// A batch of baseline JPEGs of one size -> uint8 HWC RGB on the device (the input layout of mmdx_forward), decoded by
// nvJPEG (hardware JPEG engine where the backend is available, GPU-hybrid otherwise).  NOT bit-identical to Pillow /
// libjpeg-turbo (different IDCT and chroma up-sampling rounding: ~2 % of the bytes differ by one), which is why the
// drop-in inference() keeps Pillow and this is an opt-in entry point with its own tolerance test.
static int jpeg_init(mmdx_engine* e) {
  mmdx_engine::Jpeg& j = e->jpeg;
  if (j.h) return 0;
  if (!j.lib) {
    for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
      j.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (j.lib) break;
    }
    REQUIRE(j.lib != nullptr, "libnvjpeg is not available on this machine (GPU JPEG decode is optional; use Pillow)");
    j.create_ex = reinterpret_cast<decltype(j.create_ex)>(dlsym(j.lib, "nvjpegCreateEx"));
    j.create_simple = reinterpret_cast<decltype(j.create_simple)>(dlsym(j.lib, "nvjpegCreateSimple"));
    j.destroy = reinterpret_cast<decltype(j.destroy)>(dlsym(j.lib, "nvjpegDestroy"));
    j.state_create = reinterpret_cast<decltype(j.state_create)>(dlsym(j.lib, "nvjpegJpegStateCreate"));
    j.state_destroy = reinterpret_cast<decltype(j.state_destroy)>(dlsym(j.lib, "nvjpegJpegStateDestroy"));
    j.image_info = reinterpret_cast<decltype(j.image_info)>(dlsym(j.lib, "nvjpegGetImageInfo"));
    j.batched_init = reinterpret_cast<decltype(j.batched_init)>(dlsym(j.lib, "nvjpegDecodeBatchedInitialize"));
    j.batched = reinterpret_cast<decltype(j.batched)>(dlsym(j.lib, "nvjpegDecodeBatched"));
    REQUIRE(j.create_ex && j.create_simple && j.destroy && j.state_create && j.state_destroy && j.image_info &&
                j.batched_init && j.batched, "libnvjpeg lacks the batched decode API");
  }
  const char* want = getenv("MMDX_NVJPEG_BACKEND");          // "hardware" | "hybrid" | unset: hardware, then default
  nvjpegStatus_t st = NVJPEG_STATUS_NOT_INITIALIZED;
  if (!want || strcmp(want, "hardware") == 0) {
    st = j.create_ex(NVJPEG_BACKEND_HARDWARE, nullptr, nullptr, 0, &j.h);
    if (st == NVJPEG_STATUS_SUCCESS) j.backend = (int)NVJPEG_BACKEND_HARDWARE;
  }
  if (st != NVJPEG_STATUS_SUCCESS && (!want || strcmp(want, "hybrid") == 0)) {
    st = j.create_ex(NVJPEG_BACKEND_GPU_HYBRID, nullptr, nullptr, 0, &j.h);
    if (st == NVJPEG_STATUS_SUCCESS) j.backend = (int)NVJPEG_BACKEND_GPU_HYBRID;
  }
  if (st != NVJPEG_STATUS_SUCCESS) {
    st = j.create_simple(&j.h);
    if (st == NVJPEG_STATUS_SUCCESS) j.backend = (int)NVJPEG_BACKEND_DEFAULT;
  }
  REQUIRE(st == NVJPEG_STATUS_SUCCESS, "nvjpegCreate failed");
  REQUIRE(j.state_create(j.h, &j.st) == NVJPEG_STATUS_SUCCESS, "nvjpegJpegStateCreate failed");
  j.batch = 0;
  return 0;
}

extern "C" int mmdx_decode_jpeg_batch(mmdx_engine* e, const uint8_t* const* h_blobs, const size_t* h_sizes, int n, int H, int W,
                                      uint8_t* d_out_rgb, void* stream) {
  REQUIRE(e && h_blobs && h_sizes && d_out_rgb && n > 0 && H > 0 && W > 0, "bad argument");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  TRY(jpeg_init(e));
  mmdx_engine::Jpeg& j = e->jpeg;
  for (int i = 0; i < n; ++i) {                                // every image must have the batch's size
    int nc = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t ss;
    REQUIRE(j.image_info(j.h, h_blobs[i], h_sizes[i], &nc, &ss, ws, hs) == NVJPEG_STATUS_SUCCESS, "not a decodable JPEG");
    REQUIRE(ws[0] == W && hs[0] == H, "every JPEG of a batch must be H x W");
  }
  if (j.batch != n) {
    REQUIRE(j.batched_init(j.h, j.st, n, 1, NVJPEG_OUTPUT_RGBI) == NVJPEG_STATUS_SUCCESS, "nvjpegDecodeBatchedInitialize failed");
    j.batch = n;
  }
  std::vector<nvjpegImage_t> out(n);
  for (int i = 0; i < n; ++i) {
    memset(&out[i], 0, sizeof(nvjpegImage_t));
    out[i].channel[0] = d_out_rgb + (size_t)i * H * W * 3;
    out[i].pitch[0] = (size_t)W * 3;
  }
  const nvjpegStatus_t st = j.batched(j.h, j.st, h_blobs, h_sizes, out.data(), (cudaStream_t)stream);
  if (st != NVJPEG_STATUS_SUCCESS) {
    char buf[96];
    snprintf(buf, sizeof buf, "nvjpegDecodeBatched failed (status %d, backend %d)", (int)st, j.backend);
    return fail(buf);
  }
  return 0;
}
extern "C" int mmdx_jpeg_backend(mmdx_engine* e) { return e ? e->jpeg.backend : -1; }

// ------------------------------------------------------------------------------------------ single-op entry points
extern "C" int mmdx_op_gemm(mmdx_engine* e, const void* d_a, int64_t lda, const void* d_w, const float* d_bias,
                            const void* d_residual, int64_t ldr, void* d_out, int64_t ldc, int M, int N, int K, int act,
                            int out_f32, int bn, void* stream) {
  if (e) { e->cur_stream = (cudaStream_t)stream; e->cur_cls = CLS_MISC; }
  REQUIRE(e, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  GemmLaunch g;
  TRY(build_gemm(e, g, static_cast<const bf16*>(d_a), lda, static_cast<const bf16*>(d_w), M, N, K, bn));
  TRY(fill_epilogue(e, g, d_bias, static_cast<const bf16*>(d_residual), ldr, d_out, ldc, act, out_f32));
  return launch_gemm(e, g, (cudaStream_t)stream);
}
extern "C" int mmdx_op_gemm_ln(mmdx_engine* e, const void* d_a, int64_t lda, const void* d_w, const float* d_bias,
                               const void* d_residual, int64_t ldr, void* d_pre, int64_t ldc, const float* d_gamma,
                               const float* d_beta, float eps, void* d_out, int64_t ldo, int M, int N, int K, void* stream) {
  if (e) { e->cur_stream = (cudaStream_t)stream; e->cur_cls = CLS_MISC; }
  REQUIRE(e, "null engine");
  REQUIRE(N == 768 && M >= 256 && d_residual && d_gamma && d_beta && d_out, "gemm_ln: N = 768, M >= 256, residual required");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  GemmLaunch g;
  TRY(build_gemm(e, g, static_cast<const bf16*>(d_a), lda, static_cast<const bf16*>(d_w), M, N, K, 256));
  TRY(fill_epilogue(e, g, d_bias, static_cast<const bf16*>(d_residual), ldr, d_pre, ldc, ACT_NONE, 0));
  TRY(fuse_ln(g, d_gamma, d_beta, eps, static_cast<bf16*>(d_out), ldo));
  return launch_gemm(e, g, (cudaStream_t)stream);
}
extern "C" int mmdx_op_conv(mmdx_engine* e, const void* d_in, int NB, int H, int W, int Cin, const void* d_w,
                            const float* d_bias, const void* d_residual, void* d_out, int Cout, int k, int stride,
                            int act, void* stream) {
  if (e) { e->cur_stream = (cudaStream_t)stream; e->cur_cls = CLS_MISC; }
  REQUIRE(e, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  GemmLaunch g;
  if (c64_applicable(Cin, Cout, k, stride, d_residual) && d_bias != nullptr && (act == ACT_NONE || act == ACT_RELU)) {
    TRY(build_c64(e, g, static_cast<const bf16*>(d_in), NB, H, W, static_cast<const bf16*>(d_w), d_bias,
                  static_cast<bf16*>(d_out), act));
    return launch_gemm(e, g, (cudaStream_t)stream);
  }
  TRY(build_conv(e, g, static_cast<const bf16*>(d_in), NB, H, W, Cin, static_cast<const bf16*>(d_w), Cout, k, stride));
  TRY(fill_epilogue(e, g, d_bias, static_cast<const bf16*>(d_residual), Cout, d_out, Cout, act, 0));
  return launch_gemm(e, g, (cudaStream_t)stream);
}
extern "C" int mmdx_op_bneck64(mmdx_engine* e, const void* d_t1, const void* d_res, const void* d_w2, const float* d_b2,
                               const void* d_w3, const float* d_b3, const void* d_w1n, const float* d_b1n, int c1n,
                               const void* d_x, const void* d_wd, const float* d_bd, void* d_y, void* d_t1n, int NB, int H,
                               int W, void* stream) {
  if (e) { e->cur_stream = (cudaStream_t)stream; e->cur_cls = CLS_MISC; }
  REQUIRE(e && d_t1 && d_w2 && d_b2 && d_w3 && d_b3 && d_y, "null argument");
  REQUIRE(c1n == 0 || ((c1n == 64 || c1n == 128) && d_w1n && d_b1n && d_t1n), "bneck64: next conv1 width 0 (none), 64 or 128");
  REQUIRE((d_x && d_wd && d_bd) || d_res, "bneck64: shortcut tensor, or block input + downsample weights");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  ConvW c2, c3, c1, cd;
  c2.w = (bf16*)d_w2; c2.bias = (float*)d_b2; c2.cin = 64; c2.cout = 64; c2.k = 3; c2.stride = 1;
  c3.w = (bf16*)d_w3; c3.bias = (float*)d_b3; c3.cin = 64; c3.cout = 256; c3.k = 1; c3.stride = 1;
  c1.w = (bf16*)d_w1n; c1.bias = (float*)d_b1n; c1.cin = 256; c1.cout = c1n; c1.k = 1; c1.stride = 1;
  cd.w = (bf16*)d_wd; cd.bias = (float*)d_bd; cd.cin = 64; cd.cout = 256; cd.k = 1; cd.stride = 1;
  const bool ds = d_x && d_wd && d_bd;
  GemmLaunch g;
  TRY(build_b64(e, g, static_cast<const bf16*>(d_t1), ds ? nullptr : static_cast<const bf16*>(d_res), static_cast<bf16*>(d_y),
                static_cast<bf16*>(d_t1n), NB, H, W, c2, c3, c1n ? &c1 : nullptr, ds ? &cd : nullptr,
                static_cast<const bf16*>(d_x)));
  return launch_gemm(e, g, (cudaStream_t)stream);
}
extern "C" int mmdx_op_conv3_ds(mmdx_engine* e, const void* d_t2, const void* d_x, const void* d_wcat, const float* d_bias,
                                void* d_out, int NB, int H, int W, int Cin, int Cmid, int Cout, int stride, void* stream) {
  if (e) { e->cur_stream = (cudaStream_t)stream; e->cur_cls = CLS_MISC; }
  REQUIRE(e && d_t2 && d_x && d_wcat && d_bias && d_out, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1;
  GemmLaunch g;
  TRY(build_c3ds(e, g, static_cast<const bf16*>(d_t2), NB, OH, OW, Cmid, static_cast<const bf16*>(d_x), H, W, Cin, stride,
                 static_cast<const bf16*>(d_wcat), Cout));
  TRY(fill_epilogue(e, g, d_bias, nullptr, 0, d_out, Cout, ACT_RELU, 0));
  return launch_gemm(e, g, (cudaStream_t)stream);
}
// Host-side enumeration of the two-GEMM kernel's static schedule (the SAME iterator the device roles run): the jobs of CTA
// pair `group` of `num_groups`, as (type, item, n_tile) triples.  No GPU needed; tests/test_capi_cpu.py checks that every job
// runs exactly once and that an item's second GEMM never precedes its first.  Returns the number of jobs (or -1).
extern "C" int mmdx_gemm2_schedule(int num_items, int nt1, int nt2, int reverse, int group, int num_groups, int32_t* out, int cap) {
  if (num_items < 0 || nt1 <= 0 || nt2 <= 0 || group < 0 || num_groups <= 0 || !out) return -1;
  Gemm2Sched sch(num_items, nt1, nt2, reverse, group, num_groups);
  Gemm2Job j;
  int n = 0;
  while (sch.next(j)) {
    if (n < cap) { out[3 * n] = j.type; out[3 * n + 1] = j.item; out[3 * n + 2] = j.n_t; }
    ++n;
  }
  return n;
}
extern "C" int mmdx_op_conv3_conv1(mmdx_engine* e, const void* d_t2, const void* d_w3, const float* d_b3, const void* d_res,
                                   void* d_y, const void* d_w1n, const float* d_b1n, void* d_t1n, int64_t M, int K1, int N1,
                                   int N2, void* stream) {
  if (e) { e->cur_stream = (cudaStream_t)stream; e->cur_cls = CLS_MISC; }
  REQUIRE(e && d_t2 && d_w3 && d_b3 && d_res && d_y && d_w1n && d_b1n && d_t1n, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  GemmLaunch g;
  TRY(build_dual(e, g, static_cast<const bf16*>(d_t2), K1, static_cast<const bf16*>(d_w3), d_b3, static_cast<const bf16*>(d_res),
                 static_cast<bf16*>(d_y), N1, static_cast<const bf16*>(d_w1n), d_b1n, static_cast<bf16*>(d_t1n), N2, M));
  return launch_gemm(e, g, (cudaStream_t)stream);
}
extern "C" int mmdx_op_stem_pool(mmdx_engine* e, const void* d_in_padded, int NB, int H, int W, const void* d_w_packed,
                                 const float* d_bias, void* d_out, int pool, void* stream) {
  if (e) { e->cur_stream = (cudaStream_t)stream; e->cur_cls = CLS_MISC; }
  REQUIRE(e, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  StemParams p;
  TRY(plan_stem(e, p, static_cast<const bf16*>(d_in_padded), NB, H, W, static_cast<const bf16*>(d_w_packed), d_bias,
                static_cast<bf16*>(d_out), pool ? 1 : 0));
  return launch_stem(e, p, (cudaStream_t)stream);
}
extern "C" int mmdx_op_preprocess(mmdx_engine* e, const uint8_t* d_images, int B, int H, int W, int C,
                                  void* d_out_padded, int* out_h, int* out_w, void* stream) {
  if (e) { e->cur_stream = (cudaStream_t)stream; e->cur_cls = CLS_MISC; }
  REQUIRE(e, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  PreGeom g;
  TRY(pre_geometry(e, H, W, &g));
  ResampleTable tx, ty;
  PreStrip strip;
  TRY(build_tables(e, e->tab_ws, H, W, C, g, &tx, &ty, &strip));
  int hp, wp;
  mmdx_padded_dims(g.crop_h, g.crop_w, &hp, &wp);
  if (out_h) *out_h = g.crop_h;
  if (out_w) *out_w = g.crop_w;
  return launch_preprocess(e, d_images, B, H, W, C, g, tx, ty, strip, static_cast<bf16*>(d_out_padded), hp, wp,
                           (cudaStream_t)stream);
}
extern "C" int mmdx_op_resample_u8(mmdx_engine* e, const uint8_t* d_images, int B, int H, int W, int C, uint8_t* d_out,
                                   void* stream) {
  if (e) { e->cur_stream = (cudaStream_t)stream; e->cur_cls = CLS_MISC; }
  REQUIRE(e, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  PreGeom g;
  TRY(pre_geometry(e, H, W, &g));
  ResampleTable tx, ty;
  PreStrip strip;
  TRY(build_tables(e, e->tab_ws, H, W, C, g, &tx, &ty, &strip));
  dim3 grid((g.crop_w + 255) / 256, g.crop_h, B), block(256);
  ProfScope _ps(e);
  if (C == 3)
    CK(launch_k(resample_u8_kernel<3>, dim3(grid), dim3(block), 0, (cudaStream_t)stream, 1, d_images, B, H, W, tx, ty, g.crop_h, g.crop_w, g.has_x,
                                                                   g.has_y, g.left, g.top, d_out));
  else if (C == 1)
    CK(launch_k(resample_u8_kernel<1>, dim3(grid), dim3(block), 0, (cudaStream_t)stream, 1, d_images, B, H, W, tx, ty, g.crop_h, g.crop_w, g.has_x,
                                                                   g.has_y, g.left, g.top, d_out));
  else
    return fail("mmdx: images must have 1 or 3 channels");
  CK(cudaGetLastError());
  return 0;
}
extern "C" int mmdx_op_avgpool(mmdx_engine* e, const void* d_in, int B, int HW, int C, void* d_out_bf16,
                               float* d_out_f32, void* stream) {
  if (e) { e->cur_stream = (cudaStream_t)stream; e->cur_cls = CLS_MISC; }
  REQUIRE(e && C % 8 == 0, "avgpool needs C%8==0");
  const int n = B * (C / 8);
  ProfScope _ps(e);
  CK(launch_k(avgpool_kernel, dim3((n + 255) / 256), dim3(256), 0, (cudaStream_t)stream, 1, static_cast<const bf16*>(d_in), B, HW, C,
                                                                    static_cast<bf16*>(d_out_bf16), d_out_f32));
  CK(cudaGetLastError());
  return 0;
}
extern "C" int mmdx_op_layernorm(mmdx_engine* e, const void* d_x, int rows, int N, const float* d_gamma,
                                 const float* d_beta, float eps, void* d_y, void* stream) {
  if (e) { e->cur_stream = (cudaStream_t)stream; e->cur_cls = CLS_MISC; }
  REQUIRE(e, "null engine");
  return launch_ln(e, static_cast<const bf16*>(d_x), rows, N, d_gamma, d_beta, eps, static_cast<bf16*>(d_y),
                   (cudaStream_t)stream);
}
extern "C" int mmdx_op_embed_ln(mmdx_engine* e, const int32_t* d_ids, const int32_t* d_pos, const int32_t* d_tt, int rows,
                                const void* d_word, const void* d_ptab, const void* d_ttab, const float* d_gamma,
                                const float* d_beta, float eps, void* d_y, void* stream) {
  if (e) { e->cur_stream = (cudaStream_t)stream; e->cur_cls = CLS_MISC; }
  REQUIRE(e, "null engine");
  return launch_embed(e, d_ids, d_pos, d_tt, rows, 768, static_cast<const bf16*>(d_word), static_cast<const bf16*>(d_ptab),
                      static_cast<const bf16*>(d_ttab), d_gamma, d_beta, eps, static_cast<bf16*>(d_y), (cudaStream_t)stream);
}
extern "C" int mmdx_op_attention(mmdx_engine* e, const void* d_qkv, const int32_t* d_cu, int n_seq, int T, int max_len,
                                 int n_heads, int hidden, void* d_ctx, void* stream) {
  if (e) { e->cur_stream = (cudaStream_t)stream; e->cur_cls = CLS_MISC; }
  REQUIRE(e, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  return launch_attention(e, static_cast<const bf16*>(d_qkv), d_cu, n_seq, T, max_len, n_heads, hidden,
                          static_cast<bf16*>(d_ctx), (cudaStream_t)stream);
}
extern "C" int mmdx_op_seq_mean_pool(mmdx_engine* e, const void* d_h, const int32_t* d_cu, int n_seq, int hidden,
                                     void* d_out_bf16, float* d_out_f32, void* stream) {
  if (e) { e->cur_stream = (cudaStream_t)stream; e->cur_cls = CLS_MISC; }
  REQUIRE(e && hidden % 8 == 0, "hidden % 8");
  ProfScope _ps(e);
  CK(launch_k(seq_mean_pool_kernel, dim3(dim3(n_seq, (hidden + 63) / 64)), dim3(256), 0, (cudaStream_t)stream, 1, static_cast<const bf16*>(d_h), d_cu, hidden,
                                                               static_cast<bf16*>(d_out_bf16), hidden, d_out_f32,
                                                               (const long long*)nullptr, (const float*)nullptr, (const float*)nullptr, 0.f));
  CK(cudaGetLastError());
  return 0;
}
extern "C" int mmdx_op_head_tail(mmdx_engine* e, const float* d_hidden, int B, int D, const float* d_ln_g,
                                 const float* d_ln_b, float eps, const float* d_w, const float* d_b, int n_cls,
                                 const float* d_thr, float* d_z_fuse, float* d_logits, float* d_probs, uint8_t* d_vector,
                                 void* stream) {
  if (e) { e->cur_stream = (cudaStream_t)stream; e->cur_cls = CLS_MISC; }
  REQUIRE(e && D <= 8192, "head width");
  ProfScope _ps(e);
  CK(launch_k(head_tail_kernel, dim3(B), dim3(256), D * sizeof(float), (cudaStream_t)stream, 1, d_hidden, D, d_ln_g, d_ln_b, eps, d_w, d_b, n_cls,
                                                                        d_thr, d_z_fuse, d_logits, d_probs, d_vector, nullptr));
  CK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ profiling API
extern "C" int mmdx_profile_begin(mmdx_engine* e) {
  REQUIRE(e, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  for (auto& r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  e->prof.clear();
  e->profiling = true;
  return 0;
}
// Per-launch variant of mmdx_profile_end: device time and class of every launch since mmdx_profile_begin, in launch
// order (tools/launch_times.py).  Returns the number of launches written (<= cap), or -1.
extern "C" int mmdx_profile_end_list(mmdx_engine* e, float* ms, int32_t* cls, int cap) {
  if (!e || !ms || !cls) return -1;
  std::lock_guard<std::mutex> lk(e->mu);
  if (cudaSetDevice(e->cfg.device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return -1;
  int n = 0;
  for (auto& r : e->prof) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    if (n < cap) { ms[n] = t; cls[n] = r.cls; ++n; }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  e->prof.clear();
  e->profiling = false;
  return n;
}
extern "C" int mmdx_profile_end(mmdx_engine* e, float* ms_by_class, int64_t* launches_by_class, int n_classes) {
  REQUIRE(e && ms_by_class && launches_by_class && n_classes >= CLS_COUNT, "bad argument");
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaDeviceSynchronize());
  for (int i = 0; i < n_classes; ++i) { ms_by_class[i] = 0.f; launches_by_class[i] = 0; }
  for (auto& r : e->prof) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, r.a, r.b));
    ms_by_class[r.cls] += ms;
    launches_by_class[r.cls] += 1;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  e->prof.clear();
  e->profiling = false;
  return CLS_COUNT;
}
