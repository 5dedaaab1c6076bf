// amt_model_load: a reference checkpoint (the fp32 tensors of TranscriptionModel.state_dict(), reference
// models/cnn_rnn_model.py:28-55 and :179-260) -> the packed tensors the kernels consume,
// entirely on the device and inside the library, so that a host in ANY language goes from a .pth's tensors to a
// runnable handle through the C ABI alone:
//   * BatchNorm (eval, eps 1e-5) folded into the conv weights / biases in double precision,
//   * conv weights re-laid out [Cout][(kf, kt, cin)] with the residual block's 1x1 skip conv appended along K,
//   * LSTM gate rows permuted into slice order, b_ih + b_hh summed, layer-0 columns permuted from the reference's
//     c*F + f feature index to the kernels' f*C + c, all sequences of a layer stacked along N,
//   * head weights stacked (frame | onset | offset) and zero-padded to a multiple of 128 rows,
//   * bf16 conversion -- or, in precise mode, the split-bf16 K layout [Wh | Wh | Wl (| 0)] per channel group.
// music_transcription_b200/packing.py states the same layouts in torch and is what the tests compare against, bit for bit.
#include <map>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace amt {

// One contraction weight element, written in the layout of its mode.  `j` indexes the packed K axis:
// fast: j = k.  precise: K is cut into groups of G columns, each stored as P = 3 (or 4 when G == 32) blocks of G:
// [hi | hi | lo (| 0)].
struct KLayout {
  int G, P;        // group size, blocks per group (1 = fast)
  __device__ __forceinline__ void decode(long long j, long long* k, int* part) const {
    if (P == 1) { *k = j; *part = 0; return; }
    const long long gi = j / (static_cast<long long>(G) * P);
    const int r = static_cast<int>(j - gi * G * P);
    *part = r / G;
    *k = gi * G + (r - *part * G);
  }
};

__device__ __forceinline__ __nv_bfloat16 emit_f32(float v, int part, int P) {
  if (P == 1) return __float2bfloat16_rn(v);
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  if (part < 2) return hi;
  if (part == 2) return __float2bfloat16_rn(v - __bfloat162float(hi));
  return __float2bfloat16_rn(0.0f);
}

// ---- conv weights: w [Co][Ci][kf][kt] f32, BN folded, -> out[co][col0 + j], j over taps x (Cg channels, laid out) ----
struct ConvSrc {
  const float *w, *gamma, *var;
  int Ci, Cg, taps;          // real / padded-to-block input channels; kf*kt
};

__global__ void __launch_bounds__(256) pack_conv_kernel(ConvSrc s, KLayout lay, __nv_bfloat16* __restrict__ out, float* __restrict__ out_f32,
                                                        int Co, long long ld, long long col0) {
  const long long Kp = static_cast<long long>(s.taps) * s.Cg * lay.P;       // packed columns of this conv
  const long long total = Co * Kp;
  for (long long e = blockIdx.x * 256ll + threadIdx.x; e < total; e += gridDim.x * 256ll) {
    const int co = static_cast<int>(e / Kp);
    const long long j = e - co * Kp;
    long long k;
    int part;
    lay.decode(j, &k, &part);
    const int tap = static_cast<int>(k / s.Cg), c = static_cast<int>(k - static_cast<long long>(tap) * s.Cg);
    double v = 0.0;
    if (c < s.Ci) {
      const double scale = static_cast<double>(s.gamma[co]) / sqrt(static_cast<double>(s.var[co]) + 1e-5);
      v = static_cast<double>(s.w[(static_cast<long long>(co) * s.Ci + c) * s.taps + tap]) * scale;
    }
    if (out_f32) out_f32[co * ld + col0 + j] = static_cast<float>(v);                 // stem conv: fp32 [32][9]
    else out[co * ld + col0 + j] = emit_f32(static_cast<float>(v), part, lay.P);      // double -> float -> bf16, as torch's .to(bfloat16) rounds
  }
}

struct BnBias { const float *b, *gamma, *beta, *mean, *var; };

__global__ void fold_bias_kernel(BnBias a, BnBias s, int has_skip, float* __restrict__ out, int Co) {
  const int co = blockIdx.x * blockDim.x + threadIdx.x;
  if (co >= Co) return;
  auto fold = [&](const BnBias& x) {
    const double scale = static_cast<double>(x.gamma[co]) / sqrt(static_cast<double>(x.var[co]) + 1e-5);
    return (static_cast<double>(x.b[co]) - static_cast<double>(x.mean[co])) * scale + static_cast<double>(x.beta[co]);
  };
  double v = fold(a);
  if (has_skip) v += fold(s);
  out[co] = static_cast<float>(v);
}

// ---- linear weights: src [Ns][Ks] f32 -> out rows [row0, row0 + Ns), optional slice-order row permutation (LSTM gates)
// and c*F + f -> f*C + c column permutation (layer-0 input features) ----
struct LinSrc {
  const float* w;
  int Ns, Ks;
  int H;            // > 0: rows are LSTM gates of hidden size H, packed row n <- reference row g*H + 32*s + ul
  int C, F;         // > 0: packed column f*C + c <- reference column c*F + f
};

__device__ __forceinline__ int slice_src_row(int n, int H) {
  const int s = n >> 7, r = n & 127;
  return (r & 3) * H + 32 * s + (r >> 2);
}

__global__ void __launch_bounds__(256) pack_linear_kernel(LinSrc s, KLayout lay, __nv_bfloat16* __restrict__ out, long long ld, int row0) {
  const long long Kp = static_cast<long long>(s.Ks) * lay.P;
  const long long total = s.Ns * Kp;
  for (long long e = blockIdx.x * 256ll + threadIdx.x; e < total; e += gridDim.x * 256ll) {
    const int n = static_cast<int>(e / Kp);
    const long long j = e - n * Kp;
    long long k;
    int part;
    lay.decode(j, &k, &part);
    const int sr = s.H > 0 ? slice_src_row(n, s.H) : n;
    long long sc = k;
    if (s.C > 0) {
      const int f = static_cast<int>(k / s.C), c = static_cast<int>(k - static_cast<long long>(f) * s.C);
      sc = static_cast<long long>(c) * s.F + f;
    }
    out[(row0 + n) * ld + j] = emit_f32(s.w[static_cast<long long>(sr) * s.Ks + sc], part, lay.P);
  }
}

__global__ void pack_lstm_bias_kernel(const float* __restrict__ bih, const float* __restrict__ bhh, float* __restrict__ out, int n4h, int H,
                                      int row0) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n4h) return;
  const int sr = slice_src_row(n, H);
  out[row0 + n] = static_cast<float>(static_cast<double>(bih[sr]) + static_cast<double>(bhh[sr]));
}

// ---------------------------------------------------------------------------------------------------------------
struct Src { const void* p; long long numel; };

struct Loader {
  ModelLoadView* v;       // (model internals the loader may touch; see model.cu)
  std::map<std::string, Src> src;
  cudaStream_t stream;
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0, off = 0;

  int get(const std::string& key, long long numel, const float** out) const {
    auto it = src.find(key);
    if (it == src.end()) return set_error(AMT_ERR_STATE, "model_load: checkpoint has no tensor '%s'", key.c_str());
    if (it->second.numel != numel)
      return set_error(AMT_ERR_STATE, "model_load: tensor '%s' has %lld elements, expected %lld", key.c_str(), it->second.numel, numel);
    *out = static_cast<const float*>(it->second.p);
    return 0;
  }
  void* take(const std::string& name, size_t nbytes) {
    void* p = arena + off;
    off += align_up(nbytes, 256);
    v->set(name, p, nbytes);
    return p;
  }
};

static unsigned grid_for(long long total) {
  const long long b = (total + 255) / 256;
  return static_cast<unsigned>(b < 1 ? 1 : (b > 65535 ? 65535 : b));
}

struct ConvSpec { std::string conv, bn; int Ci, Co, kf, kt; };

static int conv_src(const Loader& L, const ConvSpec& c, ConvSrc* s, BnBias* b) {
  const float *w, *bias, *gamma, *beta, *mean, *var;
  AMT_TRY(L.get(c.conv + ".weight", 1ll * c.Co * c.Ci * c.kf * c.kt, &w));
  AMT_TRY(L.get(c.conv + ".bias", c.Co, &bias));
  AMT_TRY(L.get(c.bn + ".weight", c.Co, &gamma));
  AMT_TRY(L.get(c.bn + ".bias", c.Co, &beta));
  AMT_TRY(L.get(c.bn + ".running_mean", c.Co, &mean));
  AMT_TRY(L.get(c.bn + ".running_var", c.Co, &var));
  *s = ConvSrc{w, gamma, var, c.Ci, c.Ci == 32 || c.Ci == 1 ? c.Ci : (c.Ci + 63) / 64 * 64, c.kf * c.kt};
  *b = BnBias{bias, gamma, beta, mean, var};
  return 0;
}

static KLayout layout_for(int group, bool precise) { return KLayout{group, precise ? (group == 32 ? 4 : 3) : 1}; }

// <name>.w / <name>.b from a conv (+ BN) and an optional 1x1 skip conv (+ BN) appended along K
static int pack_conv(Loader& L, const std::string& name, const ConvSpec& main, const ConvSpec* skip, bool precise) {
  ConvSrc ms, ss{};
  BnBias mb, sb{};
  AMT_TRY(conv_src(L, main, &ms, &mb));
  if (skip) AMT_TRY(conv_src(L, *skip, &ss, &sb));
  const KLayout ml = layout_for(ms.Cg, precise), sl = layout_for(skip ? ss.Cg : 64, precise);
  const long long Km = 1ll * ms.taps * ms.Cg * ml.P, Ks = skip ? 1ll * ss.taps * ss.Cg * sl.P : 0;
  auto* w = static_cast<__nv_bfloat16*>(L.take(name + ".w", static_cast<size_t>(main.Co) * (Km + Ks) * 2));
  auto* b = static_cast<float*>(L.take(name + ".b", static_cast<size_t>(main.Co) * 4));
  pack_conv_kernel<<<grid_for(main.Co * Km), 256, 0, L.stream>>>(ms, ml, w, nullptr, main.Co, Km + Ks, 0);
  AMT_CHECK_LAUNCH();
  if (skip) {
    pack_conv_kernel<<<grid_for(main.Co * Ks), 256, 0, L.stream>>>(ss, sl, w, nullptr, main.Co, Km + Ks, Km);
    AMT_CHECK_LAUNCH();
  }
  fold_bias_kernel<<<ceil_div(main.Co, 128), 128, 0, L.stream>>>(mb, sb, skip ? 1 : 0, b, main.Co);
  AMT_CHECK_LAUNCH();
  return 0;
}

struct LinPart { std::string weight; int Ns; int H; };     // one source stacked along N; H > 0: LSTM gate rows

// <name> = the listed sources stacked along N (rows zero-padded to n_pad), K = Ks, optional feature permutation
static int pack_linear(Loader& L, const std::string& name, const std::vector<LinPart>& parts, int Ks, int n_pad, int C, int F, int group,
                       bool precise) {
  const KLayout lay = layout_for(group, precise);
  const long long ld = 1ll * Ks * lay.P;
  int n_tot = 0;
  for (const auto& p : parts) n_tot += p.Ns;
  if (n_pad < n_tot) n_pad = n_tot;
  auto* out = static_cast<__nv_bfloat16*>(L.take(name, static_cast<size_t>(n_pad) * ld * 2));
  if (n_pad > n_tot) AMT_CUDA(cudaMemsetAsync(out + n_tot * ld, 0, static_cast<size_t>(n_pad - n_tot) * ld * 2, L.stream));
  int row0 = 0;
  for (const auto& p : parts) {
    const float* w;
    AMT_TRY(L.get(p.weight, 1ll * p.Ns * Ks, &w));
    pack_linear_kernel<<<grid_for(p.Ns * ld), 256, 0, L.stream>>>(LinSrc{w, p.Ns, Ks, p.H, C, F}, lay, out, ld, row0);
    AMT_CHECK_LAUNCH();
    row0 += p.Ns;
  }
  return 0;
}

static int copy_f32(Loader& L, const std::string& name, const std::vector<std::string>& keys, int n_each, int n_pad) {
  const int n_tot = n_each * static_cast<int>(keys.size());
  if (n_pad < n_tot) n_pad = n_tot;
  auto* out = static_cast<float*>(L.take(name, static_cast<size_t>(n_pad) * 4));
  if (n_pad > n_tot) AMT_CUDA(cudaMemsetAsync(out + n_tot, 0, static_cast<size_t>(n_pad - n_tot) * 4, L.stream));
  for (size_t i = 0; i < keys.size(); ++i) {
    const float* s;
    AMT_TRY(L.get(keys[i], n_each, &s));
    AMT_CUDA(cudaMemcpyAsync(out + i * n_each, s, static_cast<size_t>(n_each) * 4, cudaMemcpyDeviceToDevice, L.stream));
  }
  return 0;
}

int model_load_impl(ModelLoadView* view, const char* const* names, const void* const* ptrs, const int64_t* numels, int n,
                    cudaStream_t stream) {
  const amt_model_config& c = view->cfg();
  const bool large = c.kind == AMT_MODEL_CNN_RNN_LARGE, precise = c.precision == AMT_PRECISION_PRECISE;
  const int H = c.hidden, Hl = large ? H / 2 : 0, D = large ? 2 * H + 2 * Hl : 2 * H;
  Loader L;
  L.v = view;
  L.stream = stream;
  for (int i = 0; i < n; ++i) {
    AMT_REQUIRE(names[i] && ptrs[i] && numels[i] >= 0, "model_load: entry %d is NULL", i);
    L.src[names[i]] = Src{ptrs[i], numels[i]};
  }
  L.arena_bytes = view->expected_bytes() + 4096;
  AMT_TRY(view->alloc_arena(L.arena_bytes, &L.arena));

  // ---- stem conv (fp32, [32][9]) ----
  const std::string stem = large ? "model.conv1.0" : "model.cnn.0", stem_bn = large ? "model.conv1.1" : "model.cnn.1";
  {
    ConvSrc s;
    BnBias b;
    AMT_TRY(conv_src(L, ConvSpec{stem, stem_bn, 1, 32, 3, 3}, &s, &b));
    auto* w = static_cast<float*>(L.take("conv1.w", 32 * 9 * 4));
    auto* bo = static_cast<float*>(L.take("conv1.b", 32 * 4));
    pack_conv_kernel<<<2, 256, 0, stream>>>(s, KLayout{1, 1}, nullptr, w, 32, 9, 0);
    AMT_CHECK_LAUNCH();
    fold_bias_kernel<<<1, 128, 0, stream>>>(b, BnBias{}, 0, bo, 32);
    AMT_CHECK_LAUNCH();
  }
  int C, F;
  std::string rnn;
  if (large) {
    const ConvSpec r1c1{"model.res_block1.conv1", "model.res_block1.bn1", 32, 64, 3, 3};
    const ConvSpec r1c2{"model.res_block1.conv2", "model.res_block1.bn2", 64, 64, 3, 3};
    const ConvSpec r1s{"model.res_block1.skip.0", "model.res_block1.skip.1", 32, 64, 1, 1};
    const ConvSpec r2c1{"model.res_block2.conv1", "model.res_block2.bn1", 64, 128, 3, 3};
    const ConvSpec r2c2{"model.res_block2.conv2", "model.res_block2.bn2", 128, 128, 3, 3};
    const ConvSpec r2s{"model.res_block2.skip.0", "model.res_block2.skip.1", 64, 128, 1, 1};
    const ConvSpec fq{"model.freq_aware_conv.0", "model.freq_aware_conv.1", 128, 256, 7, 3};
    AMT_TRY(pack_conv(L, "res1.c1", r1c1, nullptr, precise));
    AMT_TRY(pack_conv(L, "res1.c2", r1c2, &r1s, precise));
    AMT_TRY(pack_conv(L, "res2.c1", r2c1, nullptr, precise));
    AMT_TRY(pack_conv(L, "res2.c2", r2c2, &r2s, precise));
    AMT_TRY(pack_conv(L, "freq", fq, nullptr, precise));
    C = 256;
    F = c.n_mels / 8;
    rnn = "model.rnn_main";
  } else {
    AMT_TRY(pack_conv(L, "c2", ConvSpec{"model.cnn.4", "model.cnn.5", 32, 64, 3, 3}, nullptr, precise));
    C = 64;
    F = c.n_mels / 4;
    rnn = "model.rnn";
  }

  // ---- LSTM stack ----
  const char* suf[2] = {"", "_reverse"};
  for (int l = 0; l < c.layers; ++l) {
    const std::string ls = std::to_string(l);
    const int Ks = l == 0 ? C * F : 2 * H;
    std::vector<LinPart> parts;
    std::vector<std::pair<std::string, int>> biases;      // (prefix of bias_ih/bias_hh key, hidden)
    for (int d = 0; d < 2; ++d) {
      parts.push_back({rnn + ".weight_ih_l" + ls + suf[d], 4 * H, H});
      biases.push_back({rnn + ".bias_", H});
      AMT_TRY(pack_linear(L, "rnn" + ls + ".whh" + std::to_string(d), {{rnn + ".weight_hh_l" + ls + suf[d], 4 * H, H}}, H, 0, 0, 0, H, false));
    }
    if (large && l == 0) {
      for (int d = 0; d < 2; ++d) {
        parts.push_back({std::string("model.rnn_local.weight_ih_l0") + suf[d], 4 * Hl, Hl});
        biases.push_back({"model.rnn_local.bias_", Hl});
        AMT_TRY(pack_linear(L, "loc.whh" + std::to_string(d), {{std::string("model.rnn_local.weight_hh_l0") + suf[d], 4 * Hl, Hl}}, Hl, 0, 0, 0, Hl, false));
      }
    }
    AMT_TRY(pack_linear(L, "rnn" + ls + ".wih", parts, Ks, 0, l == 0 ? C : 0, l == 0 ? F : 0, l == 0 ? C : 2 * H, precise));
    int n_tot = 0;
    for (const auto& p : parts) n_tot += p.Ns;
    auto* bout = static_cast<float*>(L.take("rnn" + ls + ".b", static_cast<size_t>(n_tot) * 4));
    int row0 = 0;
    for (size_t i = 0; i < parts.size(); ++i) {
      const int Hh = biases[i].second;
      const std::string lsuf = (i < 2 ? "l" + ls : std::string("l0")) + suf[i & 1];
      const float *bih, *bhh;
      AMT_TRY(L.get(biases[i].first + "ih_" + lsuf, 4 * Hh, &bih));
      AMT_TRY(L.get(biases[i].first + "hh_" + lsuf, 4 * Hh, &bhh));
      pack_lstm_bias_kernel<<<ceil_div(4 * Hh, 128), 128, 0, stream>>>(bih, bhh, bout, 4 * Hh, Hh, row0);
      AMT_CHECK_LAUNCH();
      row0 += 4 * Hh;
    }
  }

  // ---- attention, heads ----
  const int n_out = (large && c.use_onset_offset) ? 3 * 88 : 88, n_out_pad = (n_out + 127) / 128 * 128;
  if (large && c.use_attention) {
    AMT_TRY(pack_linear(L, "attn.qkv.w", {{"model.attention.qkv.weight", 3 * D, 0}}, D, 0, 0, 0, D, precise));
    AMT_TRY(copy_f32(L, "attn.qkv.b", {"model.attention.qkv.bias"}, 3 * D, 0));
    AMT_TRY(pack_linear(L, "attn.proj.w", {{"model.attention.proj.weight", D, 0}}, D, 0, 0, 0, D, precise));
    AMT_TRY(copy_f32(L, "attn.proj.b", {"model.attention.proj.bias"}, D, 0));
    AMT_TRY(copy_f32(L, "ln.w", {"model.attention_norm.weight"}, D, 0));
    AMT_TRY(copy_f32(L, "ln.b", {"model.attention_norm.bias"}, D, 0));
  }
  if (large && c.use_onset_offset) {
    AMT_TRY(pack_linear(L, "fc1.w", {{"model.shared_fc.weight", H, 0}}, D, 0, 0, 0, D, precise));
    AMT_TRY(copy_f32(L, "fc1.b", {"model.shared_fc.bias"}, H, 0));
    AMT_TRY(pack_linear(L, "heads.w", {{"model.frame_head.weight", 88, 0}, {"model.onset_head.weight", 88, 0}, {"model.offset_head.weight", 88, 0}},
                        H, n_out_pad, 0, 0, H, precise));
    AMT_TRY(copy_f32(L, "heads.b", {"model.frame_head.bias", "model.onset_head.bias", "model.offset_head.bias"}, 88, n_out_pad));
  } else {
    AMT_TRY(pack_linear(L, "heads.w", {{"model.fc.weight", 88, 0}}, D, n_out_pad, 0, 0, D, precise));
    AMT_TRY(copy_f32(L, "heads.b", {"model.fc.bias"}, 88, n_out_pad));
  }
  if (L.off > L.arena_bytes) return set_error(AMT_ERR_STATE, "model_load: internal arena overflow (%zu > %zu)", L.off, L.arena_bytes);
  // the source tensors are borrowed for the duration of THIS call only
  AMT_CUDA(cudaStreamSynchronize(stream));
  return 0;
}

}  // namespace amt
