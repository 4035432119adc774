// Native BERT WordPiece tokenizer (host side of SURVEY.md 8f N4; replaces the per-call HF tokenizer of
// tokenize_patient_details, backend/ml/pipelines/training_pipeline.py:323,335-342, for ASCII input).
//
// Restates the HF `BertTokenizer` pipeline (tokenizers: BertNormalizer -> BertPreTokenizer -> WordPiece ->
// TemplateProcessing "[CLS] $A [SEP]", truncation to max_length, padding to max_length) for 7-bit ASCII text:
//   normalise : drop NUL and control characters (Unicode category C: 0x00-0x08, 0x0B, 0x0C, 0x0E-0x1F, 0x7F), map
//               \t \n \r to ' ', lower-case A-Z (NFD accent stripping and CJK spacing are no-ops for ASCII);
//   pre-token : split on spaces, every ASCII punctuation character (33-47, 58-64, 91-96, 123-126) is its own word;
//   WordPiece : greedy longest-match-first per word, continuation pieces carry the "##" prefix, a word of more than
//               100 characters or with an unmatched remainder is ONE [UNK];
//   post      : [CLS] + pieces[: max_len - 2] + [SEP], padded with [PAD] to max_len.
// Strings with a byte >= 0x80 (the Unicode tables live in HF) or that contain the literal text of a special token
// ("[SEP]" in the input is matched as the special token by HF) are NOT tokenised here: they are flagged and the
// caller sends exactly those strings through the bundle's HF tokenizer.  tests/test_tokenizer_cpu.py holds this code to
// bit-identical ids against HF on the patient-details grammar, the reference's sample strings and adversarial ASCII.
//
// Host-only, no CUDA: the ids are an input of the GPU path (they cross PCIe with the images).  Work is split over
// threads by string; one string costs well under a microsecond.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mmdx.h"

namespace {

thread_local std::string g_tok_err;

struct Table {                       // open addressing, 64-bit FNV-1a of the token bytes, linear probing
  std::vector<uint64_t> hash;        // 0 = empty slot
  std::vector<int32_t> id;
  std::vector<uint32_t> off, len;    // token bytes inside `pool` (collisions are resolved by comparing them)
  uint64_t mask = 0;
  void init(size_t n) {
    size_t cap = 64;
    while (cap < n * 3) cap <<= 1;
    hash.assign(cap, 0); id.assign(cap, -1); off.assign(cap, 0); len.assign(cap, 0);
    mask = cap - 1;
  }
};

inline uint64_t fnv_step(uint64_t h, unsigned char c) { return (h ^ c) * 1099511628211ull; }
constexpr uint64_t kFnvInit = 14695981039346656037ull;

}  // namespace

struct mmdx_tokenizer {
  std::string pool;                  // all token bytes back to back
  Table first, cont;                 // whole-word / word-initial pieces, and "##" continuation pieces (prefix stripped)
  int32_t pad = 0, unk = 100, cls = 101, sep = 102;
  bool lower = true;
  int max_piece = 1;                 // longest piece in bytes: longer candidates cannot match
  std::vector<std::string> specials; // literal texts HF matches as special tokens before normalisation
  void insert(Table& t, const char* s, uint32_t n, int32_t tid) {
    uint64_t h = kFnvInit;
    for (uint32_t i = 0; i < n; ++i) h = fnv_step(h, (unsigned char)s[i]);
    if (h == 0) h = 1;
    size_t slot = h & t.mask;
    while (t.hash[slot] != 0) {
      if (t.hash[slot] == h && t.len[slot] == n && memcmp(pool.data() + t.off[slot], s, n) == 0) return;   // first id wins
      slot = (slot + 1) & t.mask;
    }
    t.hash[slot] = h; t.id[slot] = tid; t.off[slot] = (uint32_t)pool.size(); t.len[slot] = n;
    pool.append(s, n);
  }
  inline int32_t find(const Table& t, uint64_t h, const char* s, uint32_t n) const {
    if (h == 0) h = 1;
    size_t slot = h & t.mask;
    while (t.hash[slot] != 0) {
      if (t.hash[slot] == h && t.len[slot] == n && memcmp(pool.data() + t.off[slot], s, n) == 0) return t.id[slot];
      slot = (slot + 1) & t.mask;
    }
    return -1;
  }
};

static inline bool ascii_punct(unsigned char c) {
  return (c >= 33 && c <= 47) || (c >= 58 && c <= 64) || (c >= 91 && c <= 96) || (c >= 123 && c <= 126);
}

// One string -> ids[max_len] (padded), returns the number of valid tokens, or -1 if the string needs the HF fallback.
static int tokenize_one(const mmdx_tokenizer* t, const char* s, int n, int max_len, int32_t* ids) {
  for (int i = 0; i < n; ++i)
    if ((unsigned char)s[i] >= 0x80) return -1;
  for (const std::string& sp : t->specials) {
    if ((int)sp.size() <= n && std::search(s, s + n, sp.begin(), sp.end()) != s + n) return -1;
  }
  const int budget = max_len - 2;                 // room between [CLS] and [SEP]
  int cnt = 0;
  ids[cnt++] = t->cls;
  char word[128];
  uint64_t pre[128];
  int i = 0, emitted = 0;
  auto emit = [&](int32_t v) { if (emitted < budget) ids[cnt++] = v; ++emitted; };
  auto wordpiece = [&](const char* w, int wn) {   // wn >= 1 normalised characters
    if (wn > 100) { emit(t->unk); return; }
    int32_t pieces[100];
    int np = 0, start = 0;
    while (start < wn) {
      const Table& tab = start == 0 ? t->first : t->cont;
      const int lim = std::min(wn - start, t->max_piece);
      uint64_t h = kFnvInit;
      for (int k = 0; k < lim; ++k) { h = fnv_step(h, (unsigned char)w[start + k]); pre[k] = h; }
      int32_t hit = -1; int end = lim;
      for (; end >= 1; --end) {
        hit = t->find(tab, pre[end - 1], w + start, (uint32_t)end);
        if (hit >= 0) break;
      }
      if (hit < 0) { emit(t->unk); return; }      // the whole word becomes one [UNK]
      pieces[np++] = hit;
      start += end;
    }
    for (int k = 0; k < np; ++k) emit(pieces[k]);
  };
  while (i < n && emitted < budget) {
    unsigned char c = (unsigned char)s[i];
    if (c == ' ' || c == '\t' || c == '\n' || c == '\r') { ++i; continue; }
    if (c < 0x20 || c == 0x7F) { ++i; continue; }                      // removed by clean_text (between words here)
    if (ascii_punct(c)) { char p = (char)c; wordpiece(&p, 1); ++i; continue; }
    int wn = 0; bool too_long = false;
    while (i < n) {
      c = (unsigned char)s[i];
      if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || ascii_punct(c)) break;
      ++i;
      if (c < 0x20 || c == 0x7F) continue;                               // control characters vanish INSIDE a word
      if (t->lower && c >= 'A' && c <= 'Z') c = (unsigned char)(c + 32);
      if (wn < 127) word[wn++] = (char)c; else too_long = true;
    }
    if (wn == 0) continue;                                               // a run of control characters only
    if (too_long) emit(t->unk); else wordpiece(word, wn);
  }
  ids[cnt++] = t->sep;
  const int valid = cnt;
  while (cnt < max_len) ids[cnt++] = t->pad;
  return valid;
}

extern "C" const char* mmdx_tokenizer_last_error(void) { return g_tok_err.c_str(); }

extern "C" int mmdx_tokenizer_create(const char* vocab_txt, size_t vocab_bytes, int lower_case, mmdx_tokenizer** out) {
  if (!vocab_txt || !out) { g_tok_err = "mmdx_tokenizer_create: null argument"; return 1; }
  mmdx_tokenizer* t = new mmdx_tokenizer();
  t->lower = lower_case != 0;
  size_t lines = 1;
  for (size_t i = 0; i < vocab_bytes; ++i) lines += vocab_txt[i] == '\n';
  t->first.init(lines); t->cont.init(lines);
  t->pool.reserve(vocab_bytes);
  int32_t id = 0;
  int have = 0;
  size_t p = 0;
  while (p < vocab_bytes) {
    size_t e = p;
    while (e < vocab_bytes && vocab_txt[e] != '\n') ++e;
    size_t n = e - p;
    if (n > 0 && vocab_txt[p + n - 1] == '\r') --n;
    const char* s = vocab_txt + p;
    if (n == 5 && memcmp(s, "[PAD]", 5) == 0) { t->pad = id; have |= 1; }
    else if (n == 5 && memcmp(s, "[UNK]", 5) == 0) { t->unk = id; have |= 2; }
    else if (n == 5 && memcmp(s, "[CLS]", 5) == 0) { t->cls = id; have |= 4; }
    else if (n == 5 && memcmp(s, "[SEP]", 5) == 0) { t->sep = id; have |= 8; }
    if (n > 0) {
      if (n > 2 && s[0] == '#' && s[1] == '#') {
        t->insert(t->cont, s + 2, (uint32_t)(n - 2), id);
        t->max_piece = std::max(t->max_piece, (int)n - 2);
      } else {
        t->insert(t->first, s, (uint32_t)n, id);
        t->max_piece = std::max(t->max_piece, (int)n);
      }
    }
    ++id;
    p = e + 1;
  }
  if (have != 15) { delete t; g_tok_err = "vocabulary lacks [PAD] / [UNK] / [CLS] / [SEP]"; return 1; }
  if (t->max_piece > 100) t->max_piece = 100;
  t->specials = {"[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"};
  *out = t;
  return 0;
}

extern "C" void mmdx_tokenizer_destroy(mmdx_tokenizer* t) { delete t; }

extern "C" int mmdx_tokenize_batch(mmdx_tokenizer* t, const char* text, const int64_t* offsets, int n, int max_len,
                                   int n_threads, int32_t* ids, int32_t* lens, uint8_t* fallback) {
  if (!t || !text || !offsets || !ids || !lens || !fallback || n < 0) { g_tok_err = "mmdx_tokenize_batch: bad argument"; return 1; }
  if (max_len < 2) { g_tok_err = "mmdx_tokenize_batch: max_len must be >= 2 ([CLS] and [SEP])"; return 1; }
  auto work = [&](int lo, int hi) {
    for (int i = lo; i < hi; ++i) {
      const int64_t a = offsets[i], b = offsets[i + 1];
      const int v = tokenize_one(t, text + a, (int)(b - a), max_len, ids + (size_t)i * max_len);
      fallback[i] = v < 0;
      lens[i] = v < 0 ? 0 : v;
    }
  };
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, (n + 63) / 64));            // a thread is worth starting for >= 64 strings
  if (nt == 1) { work(0, n); return 0; }
  std::vector<std::thread> th;
  const int per = (n + nt - 1) / nt;
  for (int k = 0; k < nt; ++k) {
    const int lo = k * per, hi = std::min(n, lo + per);
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto& x : th) x.join();
  return 0;
}
