// amt_model: eval-mode forward of CNNRNNModel / CNNRNNModelLarge
// (reference models/cnn_rnn_model.py:57-74 and :262-349) as a fixed sequence of
// kernel launches on the caller's stream.  All intermediate tensors live in the
// caller-provided workspace; weights are the packed tensors registered by name.
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace amt {

struct Tensor {
  const void* ptr = nullptr;
  size_t nbytes = 0;
};

}  // namespace amt

namespace amt {

// CUDA-event stage timer: one (start, stop) pair per stage per forward call; pending pairs are
// folded into the totals at the next forward call or at read time.
struct Profiler {
  struct Stage { std::string name; double ms = 0; int launches = 0; };
  struct Pending { int stage; cudaEvent_t a, b; };
  bool enabled = false;
  std::vector<Stage> stages;
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;
  int stage_id(const std::string& n) {
    for (size_t i = 0; i < stages.size(); ++i) if (stages[i].name == n) return (int)i;
    stages.push_back(Stage{n});
    return (int)stages.size() - 1;
  }
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void begin(const std::string& n, cudaStream_t s) {
    if (!enabled) return;
    Pending p{stage_id(n), get(), get()};
    cudaEventRecord(p.a, s);
    pending.push_back(p);
  }
  void end(cudaStream_t s) {
    if (!enabled || pending.empty()) return;
    cudaEventRecord(pending.back().b, s);
  }
  void fold() {
    for (Pending& p : pending) {
      cudaEventSynchronize(p.b);
      float ms = 0;
      if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { stages[p.stage].ms += ms; stages[p.stage].launches += 1; }
      pool.push_back(p.a); pool.push_back(p.b);
    }
    pending.clear();
  }
  void reset() { fold(); stages.clear(); }
  ~Profiler() { for (auto& p : pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); } for (auto e : pool) cudaEventDestroy(e); }
};

}  // namespace amt

struct amt_model {
  amt::Profiler prof;
  amt_model_config cfg;
  std::map<std::string, amt::Tensor> tensors;
  bool finalized = false;
  void* arena = nullptr;      // packed weights produced by amt_model_load (owned; amt_model_set_tensor pointers are borrowed)
  int arena_device = -1;
  // derived sizes
  int F1, F2, F3;       // frequency bins after 1, 2, 3 pools
  int Kfeat;            // LSTM input width
  int H, Hl, D;         // hidden, local hidden, rnn output width
  int Ng0;              // gate columns of layer 0 (all sequences)
  int n_out, n_out_pad; // head columns (88 * n_heads) and padded GEMM N
  // precise mode (cfg.precision == 1): split-bf16 operands, K axis of every contraction tripled
  int KM;               // K multiplier: 1 (fast) or 3 (precise: [hi | lo | hi] x [Wh | Wh | Wl])
  int C1;               // channels of one stem-output pixel: 32, or the 128-channel group [hi | lo | hi | 0]
};

namespace amt {

struct Expect { std::string name; size_t nbytes; };

static std::vector<Expect> expected_tensors(const amt_model& m) {
  const amt_model_config& c = m.cfg;
  std::vector<Expect> e;
  auto bf = [](size_t n) { return n * 2; };
  auto f32 = [](size_t n) { return n * 4; };
  const size_t km = m.KM, c1 = m.C1;       // operand channels: c1 for the stem output, km * C elsewhere
  e.push_back({"conv1.w", f32(32 * 9)});
  e.push_back({"conv1.b", f32(32)});
  if (c.kind == AMT_MODEL_CNN_RNN) {
    e.push_back({"c2.w", bf(64ull * 9 * c1)});
    e.push_back({"c2.b", f32(64)});
  } else {
    e.push_back({"res1.c1.w", bf(64ull * 9 * c1)});
    e.push_back({"res1.c1.b", f32(64)});
    e.push_back({"res1.c2.w", bf(64ull * (9 * 64 * km + c1))});
    e.push_back({"res1.c2.b", f32(64)});
    e.push_back({"res2.c1.w", bf(128ull * 9 * 64 * km)});
    e.push_back({"res2.c1.b", f32(128)});
    e.push_back({"res2.c2.w", bf(128ull * (9 * 128 + 64) * km)});
    e.push_back({"res2.c2.b", f32(128)});
    e.push_back({"freq.w", bf(256ull * 21 * 128 * km)});
    e.push_back({"freq.b", f32(256)});
  }
  for (int l = 0; l < c.layers; ++l) {
    const size_t N = l == 0 ? m.Ng0 : 8ull * m.H;
    const size_t K = km * (l == 0 ? m.Kfeat : 2ull * m.H);
    e.push_back({"rnn" + std::to_string(l) + ".wih", bf(N * K)});
    e.push_back({"rnn" + std::to_string(l) + ".b", f32(N)});
    for (int d = 0; d < 2; ++d) e.push_back({"rnn" + std::to_string(l) + ".whh" + std::to_string(d), bf(4ull * m.H * m.H)});
  }
  if (c.kind == AMT_MODEL_CNN_RNN_LARGE) {
    for (int d = 0; d < 2; ++d) e.push_back({"loc.whh" + std::to_string(d), bf(4ull * m.Hl * m.Hl)});
    if (c.use_attention) {
      e.push_back({"attn.qkv.w", bf(3ull * m.D * m.D * km)});
      e.push_back({"attn.qkv.b", f32(3ull * m.D)});
      e.push_back({"attn.proj.w", bf(1ull * m.D * m.D * km)});
      e.push_back({"attn.proj.b", f32(m.D)});
      e.push_back({"ln.w", f32(m.D)});
      e.push_back({"ln.b", f32(m.D)});
    }
    if (c.use_onset_offset) {
      e.push_back({"fc1.w", bf(1ull * m.H * m.D * km)});
      e.push_back({"fc1.b", f32(m.H)});
      e.push_back({"heads.w", bf(1ull * m.n_out_pad * m.H * km)});
    } else {
      e.push_back({"heads.w", bf(1ull * m.n_out_pad * m.D * km)});
    }
  } else {
    e.push_back({"heads.w", bf(1ull * m.n_out_pad * m.D * km)});
  }
  e.push_back({"heads.b", f32(m.n_out_pad)});
  return e;
}

// workspace carving --------------------------------------------------------
struct Workspace {
  uint8_t* base;
  size_t off = 0;
  explicit Workspace(void* p) : base(static_cast<uint8_t*>(p)) {}
  void* take(size_t bytes) {
    void* r = base ? base + off : nullptr;
    off += align_up(bytes, 1024);
    return r;
  }
};

struct Buffers {
  void *act1, *h1, *act2, *h2, *act3, *feat;    // CNN activations (bf16)
  float* gx;                                     // gate pre-activations (f32), max over layers
  void* seq_a; void* seq_b;                      // inter-layer LSTM outputs (bf16) ping-pong
  void* rnn_bf16; float* rnn_f32;                // final LSTM features
  void *qkv, *att; float* proj; void* normed; void* shared; float* logits;
  void* lstm_scratch; size_t lstm_scratch_bytes;
  // precise mode only: fp32 outputs of the stages whose consumers need split operands
  float* seq_f32; float* shared_f32; void* att_split;
};

static size_t carve(const amt_model& m, int B, int T, void* ws, Buffers* b) {
  Workspace w(ws);
  const size_t BT = static_cast<size_t>(B) * T;
  const bool large = m.cfg.kind == AMT_MODEL_CNN_RNN_LARGE;
  const size_t km = m.KM;                   // bf16 operand tensors are km x wider in precise mode
  const bool precise = km > 1;
  b->act1 = w.take(BT * m.F1 * m.C1 * 2);
  if (large) {
    b->h1 = w.take(BT * m.F1 * 64 * km * 2);
    b->act2 = w.take(BT * m.F2 * 64 * km * 2);
    b->h2 = w.take(BT * m.F2 * 128 * km * 2);
    b->act3 = w.take(BT * m.F2 * 128 * km * 2);
    b->feat = w.take(BT * m.F3 * 256 * km * 2);
  } else {
    b->h1 = b->act2 = b->h2 = b->act3 = nullptr;
    b->feat = w.take(BT * m.F2 * 64 * km * 2);
  }
  const size_t gcols = std::max<size_t>(m.Ng0, 8ull * m.H);
  b->gx = static_cast<float*>(w.take(BT * gcols * 4));
  b->seq_a = w.take(BT * 2 * m.H * km * 2);
  b->seq_b = w.take(BT * 2 * m.H * km * 2);
  b->rnn_bf16 = w.take(BT * m.D * km * 2);
  b->rnn_f32 = static_cast<float*>(w.take(BT * m.D * 4));
  if (large && m.cfg.use_attention) {
    b->qkv = w.take(BT * 3 * m.D * 2);
    b->att = w.take(BT * m.D * 2);
    b->proj = static_cast<float*>(w.take(BT * m.D * 4));
    b->normed = w.take(BT * m.D * km * 2);
  } else {
    b->qkv = b->att = b->normed = nullptr;
    b->proj = nullptr;
  }
  b->shared = (large && m.cfg.use_onset_offset) ? w.take(BT * m.H * km * 2) : nullptr;
  b->logits = static_cast<float*>(w.take(BT * m.n_out_pad * 4));
  b->seq_f32 = precise ? static_cast<float*>(w.take(BT * 2 * m.H * 4)) : nullptr;
  b->shared_f32 = (precise && large && m.cfg.use_onset_offset) ? static_cast<float*>(w.take(BT * m.H * 4)) : nullptr;
  b->att_split = (precise && large && m.cfg.use_attention) ? w.take(BT * m.D * km * 2) : nullptr;
  // recurrence scratch: worst case is layer 0 of the large model (4 sequences)
  amt_lstm_seq seqs[4];
  int n = 0;
  for (int d = 0; d < 2; ++d) { seqs[n] = amt_lstm_seq{}; seqs[n++].H = m.H; }
  if (large) for (int d = 0; d < 2; ++d) { seqs[n] = amt_lstm_seq{}; seqs[n++].H = m.Hl; }
  b->lstm_scratch_bytes = std::max(lstm_scratch_bytes(seqs, n, B), lstm_scratch_bytes(seqs, 2, B));
  b->lstm_scratch = w.take(b->lstm_scratch_bytes);
  return w.off;
}

static const void* T_(const amt_model& m, const std::string& n) { return m.tensors.at(n).ptr; }
static const float* F_(const amt_model& m, const std::string& n) { return static_cast<const float*>(m.tensors.at(n).ptr); }

static int conv(const amt_model& m, const void* X, int C, const void* X2, int C2, int B, int T, int F, const void* W,
                const float* bias, int N, int kf, int kt, void* out, int pool, cudaStream_t s) {
  return run_conv_halo(X, C, X2, C2, B, T, F, W, bias, N, kf, kt, out, 1, pool, m.KM > 1, s);
}

static int forward(amt_model& m, const float* logmel, const float* chunk_max, float top_db, int B, int T, float* frame,
                   float* onset, float* offset, void* ws, cudaStream_t s) {
  const amt_model_config& c = m.cfg;
  const bool large = c.kind == AMT_MODEL_CNN_RNN_LARGE;
  const bool precise = m.KM > 1;
  const int km = m.KM;
  Buffers b;
  carve(m, B, T, ws, &b);
  const int BT = B * T;
  if (m.prof.enabled) m.prof.fold();      // previous call's events (host sync; profiling mode only)
#define STAGE(name, call)               \
  do {                                  \
    m.prof.begin(name, s);              \
    const int _st = (call);             \
    m.prof.end(s);                      \
    if (_st != 0) return _st;           \
  } while (0)

  // ---- CNN ----  (precise: every activation tensor carries km = 3 x the channels, [hi | lo | hi] per pixel)
  STAGE("conv1", run_conv1(logmel, chunk_max, top_db, F_(m, "conv1.w"), F_(m, "conv1.b"), b.act1, B, c.n_mels, T, precise, s));
  if (large) {
    STAGE("res1.c1", conv(m, b.act1, m.C1, nullptr, 0, B, T, m.F1, T_(m, "res1.c1.w"), F_(m, "res1.c1.b"), 64, 3, 3, b.h1, 0, s));
    STAGE("res1.c2", conv(m, b.h1, 64 * km, b.act1, m.C1, B, T, m.F1, T_(m, "res1.c2.w"), F_(m, "res1.c2.b"), 64, 3, 3, b.act2, 1, s));
    STAGE("res2.c1", conv(m, b.act2, 64 * km, nullptr, 0, B, T, m.F2, T_(m, "res2.c1.w"), F_(m, "res2.c1.b"), 128, 3, 3, b.h2, 0, s));
    STAGE("res2.c2", conv(m, b.h2, 128 * km, b.act2, 64 * km, B, T, m.F2, T_(m, "res2.c2.w"), F_(m, "res2.c2.b"), 128, 3, 3, b.act3, 0, s));
    STAGE("freq", conv(m, b.act3, 128 * km, nullptr, 0, B, T, m.F2, T_(m, "freq.w"), F_(m, "freq.b"), 256, 7, 3, b.feat, 1, s));
  } else {
    STAGE("c2", conv(m, b.act1, m.C1, nullptr, 0, B, T, m.F1, T_(m, "c2.w"), F_(m, "c2.b"), 64, 3, 3, b.feat, 1, s));
  }

  // ---- BiLSTM stack (+ the parallel local BiLSTM of the large model on layer 0) ----
  // precise: the recurrences write fp32 outputs, which split3 turns into the next GEMM's [hi | lo | hi] operand
  const void* x = b.feat;
  int K = m.Kfeat * km;
  for (int l = 0; l < c.layers; ++l) {
    const std::string pre = "rnn" + std::to_string(l);
    const int N = l == 0 ? m.Ng0 : 8 * m.H;
    STAGE(pre + ".gemm", run_gemm(x, T_(m, pre + ".wih"), F_(m, pre + ".b"), b.gx, BT, N, K, N, 0, 1, s));
    const bool last = l == c.layers - 1;
    void* out_bf = last ? b.rnn_bf16 : ((l & 1) ? b.seq_b : b.seq_a);
    const int ld_out = last ? m.D : 2 * m.H;
    amt_lstm_seq seqs[4];
    int n = 0;
    for (int d = 0; d < 2; ++d) {
      amt_lstm_seq& q = seqs[n++];
      q.whh = T_(m, pre + ".whh" + std::to_string(d));
      q.gx = b.gx + d * 4 * m.H;
      q.out_bf16 = precise ? nullptr : static_cast<__nv_bfloat16*>(out_bf) + d * m.H;
      q.out_f32 = last ? b.rnn_f32 + d * m.H : (precise ? b.seq_f32 + d * m.H : nullptr);
      q.H = m.H; q.reverse = d; q.ld_gx = N; q.ld_out = ld_out; q.ld_out32 = last ? m.D : 2 * m.H;
    }
    if (large && l == 0) {
      for (int d = 0; d < 2; ++d) {
        amt_lstm_seq& q = seqs[n++];
        q.whh = T_(m, "loc.whh" + std::to_string(d));
        q.gx = b.gx + 8 * m.H + d * 4 * m.Hl;
        q.out_bf16 = precise ? nullptr : static_cast<__nv_bfloat16*>(b.rnn_bf16) + 2 * m.H + d * m.Hl;
        q.out_f32 = b.rnn_f32 + 2 * m.H + d * m.Hl;
        q.H = m.Hl; q.reverse = d; q.ld_gx = N; q.ld_out = m.D; q.ld_out32 = m.D;
      }
    }
    STAGE(pre + ".rec", run_lstm(seqs, n, B, T, b.lstm_scratch, b.lstm_scratch_bytes, s));
    if (precise && !last) STAGE(pre + ".split", run_split3(b.seq_f32, 1, out_bf, BT, 2 * m.H, s));
    x = out_bf;
    K = 2 * m.H * km;
  }
  if (precise) STAGE("rnn.split", run_split3(b.rnn_f32, 1, b.rnn_bf16, BT, m.D, s));

  // ---- attention + residual LayerNorm ----
  const void* head_in = b.rnn_bf16;
  if (large && c.use_attention) {
    STAGE("attn.qkv", run_gemm(b.rnn_bf16, T_(m, "attn.qkv.w"), F_(m, "attn.qkv.b"), b.qkv, BT, 3 * m.D, m.D * km, 3 * m.D, 0, 0, s));
    STAGE("attn.core", run_attention(b.qkv, b.att, B, T, c.heads, m.D / c.heads, 10.0f, s));
    const void* att_in = b.att;
    if (precise) {      // the attention core keeps bf16 q, k, v, P (< 1e-4 on the probabilities); its output has lo = 0
      STAGE("attn.split", run_split3(b.att, 0, b.att_split, BT, m.D, s));
      att_in = b.att_split;
    }
    STAGE("attn.proj", run_gemm(att_in, T_(m, "attn.proj.w"), F_(m, "attn.proj.b"), b.proj, BT, m.D, m.D * km, m.D, 0, 1, s));
    STAGE("add_ln", run_add_layernorm(b.rnn_f32, b.proj, F_(m, "ln.w"), F_(m, "ln.b"), b.normed, BT, m.D, 1e-6f, precise, s));
    head_in = b.normed;
  }

  // ---- output heads ----
  // The stacked head GEMM has frame | onset | offset in columns 0..87 | 88..175 | 176..263 (padded to 384).  When the
  // caller asks for the frame head only (main.py's path), only the first 128-column n-tile is computed and transposed:
  // the same accumulations, a third of the GEMM and of the transpose.
  int n_heads = 1;
  const bool three = large && c.use_onset_offset && (onset != nullptr || offset != nullptr);
  const int n_head_cols = three ? m.n_out_pad : 128;
  if (large && c.use_onset_offset) {
    if (precise) {
      STAGE("fc1", run_gemm(head_in, T_(m, "fc1.w"), F_(m, "fc1.b"), b.shared_f32, BT, m.H, m.D * km, m.H, 1, 1, s));
      STAGE("fc1.split", run_split3(b.shared_f32, 1, b.shared, BT, m.H, s));
    } else {
      STAGE("fc1", run_gemm(head_in, T_(m, "fc1.w"), F_(m, "fc1.b"), b.shared, BT, m.H, m.D, m.H, 1, 0, s));
    }
    STAGE("heads", run_gemm(b.shared, T_(m, "heads.w"), F_(m, "heads.b"), b.logits, BT, n_head_cols, m.H * km, m.n_out_pad, 0, 1, s));
    n_heads = three ? 3 : 1;
  } else {
    STAGE("heads", run_gemm(head_in, T_(m, "heads.w"), F_(m, "heads.b"), b.logits, BT, m.n_out_pad, m.D * km, m.n_out_pad, 0, 1, s));
  }
  STAGE("heads.transpose", run_heads_transpose(b.logits, m.n_out_pad, B, T, n_heads, frame, n_heads == 3 ? onset : nullptr,
                                               n_heads == 3 ? offset : nullptr, s));
#undef STAGE
  return 0;
}

}  // namespace amt

extern "C" {

int amt_model_create(const amt_model_config* cfg, amt_model** out) {
  using namespace amt;
  AMT_REQUIRE(cfg && out, "model_create: NULL argument");
  AMT_REQUIRE(cfg->kind == AMT_MODEL_CNN_RNN || cfg->kind == AMT_MODEL_CNN_RNN_LARGE, "model_create: unknown kind %d", cfg->kind);
  AMT_REQUIRE(cfg->hidden % 128 == 0 && cfg->hidden >= 128 && cfg->hidden <= 640,
              "model_create: hidden_size %d unsupported (multiple of 128 in 128..640)", cfg->hidden);
  AMT_REQUIRE(cfg->layers >= 1 && cfg->layers <= 8, "model_create: num_layers %d unsupported", cfg->layers);
  const bool large = cfg->kind == AMT_MODEL_CNN_RNN_LARGE;
  AMT_REQUIRE(cfg->n_mels >= (large ? 8 : 4) && cfg->n_mels <= 4096, "model_create: n_mels %d unsupported", cfg->n_mels);
  if (large && cfg->use_attention) {
    AMT_REQUIRE(cfg->heads == 8, "model_create: the reference fixes 8 attention heads");
    const int hd = 3 * cfg->hidden / 8;
    AMT_REQUIRE(hd == 48 || hd == 96 || hd == 144 || hd == 192, "model_create: head_dim %d unsupported (hidden <= 512 with attention)", hd);
  }
  AMT_REQUIRE(cfg->precision == AMT_PRECISION_FAST || cfg->precision == AMT_PRECISION_PRECISE,
              "model_create: precision %d unknown (0 = fast bf16, 1 = precise split-bf16)", cfg->precision);
  auto* m = new amt_model();
  m->cfg = *cfg;
  m->KM = cfg->precision == AMT_PRECISION_PRECISE ? 3 : 1;
  m->C1 = cfg->precision == AMT_PRECISION_PRECISE ? 128 : 32;
  m->F1 = cfg->n_mels / 2;
  m->F2 = m->F1 / 2;
  m->F3 = m->F2 / 2;
  m->H = cfg->hidden;
  m->Hl = large ? cfg->hidden / 2 : 0;
  m->D = large ? 2 * m->H + 2 * m->Hl : 2 * m->H;
  m->Kfeat = large ? 256 * m->F3 : 64 * m->F2;
  m->Ng0 = 8 * m->H + 8 * m->Hl;
  m->n_out = (large && cfg->use_onset_offset) ? 3 * 88 : 88;
  m->n_out_pad = (m->n_out + 127) / 128 * 128;
  *out = m;
  return 0;
}

static void free_arena(amt_model* m) {
  if (!m->arena) return;
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != m->arena_device) cudaSetDevice(m->arena_device);
  cudaFree(m->arena);
  if (cur != m->arena_device) cudaSetDevice(cur);
  m->arena = nullptr;
}

int amt_model_destroy(amt_model* m) {
  if (m) free_arena(m);
  delete m;
  return 0;
}

namespace amt {
struct LoadView : ModelLoadView {
  amt_model* m;
  explicit LoadView(amt_model* mm) : m(mm) {}
  const amt_model_config& cfg() const override { return m->cfg; }
  size_t expected_bytes() const override {
    size_t tot = 0;
    for (const Expect& e : expected_tensors(*m)) tot += align_up(e.nbytes, 256);
    return tot;
  }
  int alloc_arena(size_t bytes, uint8_t** out) override {
    free_arena(m);
    m->tensors.clear();
    m->finalized = false;
    AMT_CUDA(cudaGetDevice(&m->arena_device));
    AMT_CUDA(cudaMalloc(&m->arena, bytes));
    *out = static_cast<uint8_t*>(m->arena);
    return 0;
  }
  void set(const std::string& name, void* p, size_t nbytes) override { m->tensors[name] = Tensor{p, nbytes}; }
};
}  // namespace amt

int amt_model_load(amt_model* m, const char* const* names, const void* const* ptrs, const int64_t* numels, int n,
                   amt_stream_t stream) {
  using namespace amt;
  AMT_REQUIRE(m && names && ptrs && numels && n > 0, "model_load: bad arguments");
  AMT_TRY(ensure_device());
  LoadView view(m);
  const int st = model_load_impl(&view, names, ptrs, numels, n, static_cast<cudaStream_t>(stream));
  if (st != 0) {
    free_arena(m);
    m->tensors.clear();
    return st;
  }
  return amt_model_finalize(m);
}

int amt_model_get_tensor(const amt_model* m, const char* name, const void** dev_ptr, size_t* nbytes) {
  using namespace amt;
  AMT_REQUIRE(m && name && dev_ptr && nbytes, "model_get_tensor: NULL argument");
  auto it = m->tensors.find(name);
  if (it == m->tensors.end()) return set_error(AMT_ERR_STATE, "model_get_tensor: no packed tensor '%s'", name);
  *dev_ptr = it->second.ptr;
  *nbytes = it->second.nbytes;
  return 0;
}

int amt_model_set_tensor(amt_model* m, const char* name, const void* dev_ptr, size_t nbytes) {
  using namespace amt;
  AMT_REQUIRE(m && name && dev_ptr, "model_set_tensor: NULL argument");
  AMT_REQUIRE(reinterpret_cast<uintptr_t>(dev_ptr) % 16 == 0, "model_set_tensor: %s must be 16-byte aligned", name);
  m->tensors[name] = Tensor{dev_ptr, nbytes};
  m->finalized = false;
  return 0;
}

int amt_model_finalize(amt_model* m) {
  using namespace amt;
  AMT_REQUIRE(m, "model_finalize: NULL model");
  for (const Expect& e : expected_tensors(*m)) {
    auto it = m->tensors.find(e.name);
    if (it == m->tensors.end()) return set_error(AMT_ERR_STATE, "model_finalize: missing tensor '%s'", e.name.c_str());
    if (it->second.nbytes != e.nbytes)
      return set_error(AMT_ERR_STATE, "model_finalize: tensor '%s' has %zu bytes, expected %zu", e.name.c_str(),
                       it->second.nbytes, e.nbytes);
  }
  m->finalized = true;
  return 0;
}

int amt_model_profile_enable(amt_model* m, int enable) {
  using namespace amt;
  AMT_REQUIRE(m, "model_profile_enable: NULL model");
  m->prof.reset();
  m->prof.enabled = enable != 0;
  return 0;
}

int amt_model_profile_read(amt_model* m, char* names, float* total_ms, int* launches, int cap) {
  using namespace amt;
  AMT_REQUIRE(m && names && total_ms && launches && cap >= 0, "model_profile_read: bad arguments");
  m->prof.fold();
  const int n = static_cast<int>(m->prof.stages.size());
  for (int i = 0; i < n && i < cap; ++i) {
    snprintf(names + 32 * i, 32, "%s", m->prof.stages[i].name.c_str());
    total_ms[i] = static_cast<float>(m->prof.stages[i].ms);
    launches[i] = m->prof.stages[i].launches;
  }
  return n;
}

int amt_model_profile_in_flight(amt_model* m, char* name, int cap) {
  using namespace amt;
  AMT_REQUIRE(m && name && cap > 0, "model_profile_in_flight: bad arguments");
  int idx = 0;
  for (const Profiler::Pending& p : m->prof.pending) {
    if (cudaEventQuery(p.b) != cudaSuccess) {
      cudaGetLastError();
      snprintf(name, cap, "%s", m->prof.stages[p.stage].name.c_str());
      return idx;
    }
    ++idx;
  }
  name[0] = 0;
  return -1;
}

int amt_model_workspace_layout(const amt_model* m, int B, int T, const char* name, size_t* offset, size_t* nbytes) {
  using namespace amt;
  AMT_REQUIRE(m && name && offset && nbytes && B >= 1 && T >= 1, "model_workspace_layout: bad arguments");
  Buffers b;
  uint8_t* const base = reinterpret_cast<uint8_t*>(uintptr_t(1) << 20);       // fake base: only differences are used
  const size_t total = carve(*m, B, T, base, &b);
  const struct { const char* n; const void* p; } tab[] = {
      {"act1", b.act1}, {"h1", b.h1}, {"act2", b.act2}, {"h2", b.h2}, {"act3", b.act3}, {"feat", b.feat}, {"gx", b.gx},
      {"seq_a", b.seq_a}, {"seq_b", b.seq_b}, {"rnn_bf16", b.rnn_bf16}, {"rnn_f32", b.rnn_f32}, {"qkv", b.qkv},
      {"att", b.att}, {"proj", b.proj}, {"normed", b.normed}, {"shared", b.shared}, {"logits", b.logits},
      {"lstm_scratch", b.lstm_scratch}};
  std::vector<size_t> offs;
  for (const auto& e : tab) if (e.p) offs.push_back(static_cast<const uint8_t*>(e.p) - base);
  offs.push_back(total);
  for (const auto& e : tab) {
    if (strcmp(e.n, name) != 0) continue;
    if (!e.p) return set_error(AMT_ERR_ARG, "model_workspace_layout: this configuration has no buffer '%s'", name);
    const size_t off = static_cast<const uint8_t*>(e.p) - base;
    size_t next = total;
    for (size_t o : offs) if (o > off && o < next) next = o;
    *offset = off;
    *nbytes = next - off;                 // up to the next buffer (includes the 1 KB alignment padding)
    return 0;
  }
  return set_error(AMT_ERR_ARG, "model_workspace_layout: unknown buffer '%s'", name);
}

size_t amt_model_workspace_bytes(const amt_model* m, int B, int T) {
  if (!m || B <= 0 || T <= 0) return 0;
  amt::Buffers b;
  return amt::carve(*m, B, T, nullptr, &b);
}

int amt_model_forward(amt_model* m, const float* logmel, int B, int T, float* frame, float* onset, float* offset,
                      void* workspace, size_t workspace_bytes, amt_stream_t stream) {
  return amt_model_forward_db(m, logmel, nullptr, 0.0f, B, T, frame, onset, offset, workspace, workspace_bytes, stream);
}

int amt_model_forward_db(amt_model* m, const float* logmel, const float* chunk_max, float top_db, int B, int T, float* frame,
                         float* onset, float* offset, void* workspace, size_t workspace_bytes, amt_stream_t stream) {
  using namespace amt;
  AMT_REQUIRE(m && logmel && frame && workspace, "model_forward: NULL argument");
  AMT_REQUIRE(chunk_max == nullptr || top_db >= 0.0f, "model_forward: top_db must be >= 0 when chunk_max is given");
  AMT_REQUIRE(B >= 1 && T >= 1, "model_forward: B and T must be >= 1 (the reference's conv rejects T == 0 too)");
  if (!m->finalized) return set_error(AMT_ERR_STATE, "model_forward: call amt_model_finalize first");
  AMT_TRY(ensure_device());
  AMT_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 1024 == 0, "model_forward: workspace must be 1024-byte aligned");
  const size_t need = amt_model_workspace_bytes(m, B, T);
  if (workspace_bytes < need) return set_error(AMT_ERR_WORKSPACE, "model_forward: workspace %zu < %zu bytes", workspace_bytes, need);
  AMT_REQUIRE(static_cast<long long>(B) * T < (1ll << 30), "model_forward: B*T too large");
  return forward(*m, logmel, chunk_max, top_db, B, T, frame, onset, offset, workspace, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
