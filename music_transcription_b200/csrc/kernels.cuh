// Internal (non-ABI) launch entry points shared between translation units.
#pragma once
#include <string>

#include "common.cuh"

namespace amt {

// halo-tile implicit-GEMM convolution (conv_halo.cu): X [B][T][F][C], optional 1x1 skip source X2 [B][T][F][C2],
// W [N][kf*kt*C + C2] (K index = (kf, kt, c)), out [B][T][F or F/2][N] bf16
// split != 0: out [B][T][F'][3N] = [hi(N) | lo(N) | hi(N)] per pixel
int run_conv_halo(const void* X, int C, const void* X2, int C2, int B, int T, int F, const void* W, const float* bias,
                  int N, int kf, int kt, void* out, int relu, int pool, int split, cudaStream_t stream);
int run_gemm(const void* A, const void* W, const float* bias, void* C, int M, int N, int K, long long ldc, int relu,
             int out_f32, cudaStream_t stream);
int run_lstm(const amt_lstm_seq* seqs, int n_seq, int B, int T, void* scratch, size_t scratch_bytes, cudaStream_t stream);
size_t lstm_scratch_bytes(const amt_lstm_seq* seqs, int n_seq, int B);
int run_attention(const void* qkv, void* out, int B, int T, int heads, int head_dim, float clip, cudaStream_t stream);
int run_attention_tc(const void* qkv, void* out, int B, int T, int heads, int head_dim, float clip, cudaStream_t stream);
// split != 0 (precise mode): outputs in the split-bf16 layout [hi | lo | hi (| 0)] per channel group
// chunk_max != nullptr: x is the unfloored log-mel; max(x, chunk_max[b] - top_db) is applied on load
int run_conv1(const float* x, const float* chunk_max, float top_db, const float* w, const float* bias, void* out, int B, int Fin,
              int T, int split, cudaStream_t stream);
int run_add_layernorm(const float* a, const float* b, const float* gamma, const float* beta, void* out, long long rows,
                      int D, float eps, int split, cudaStream_t stream);
int run_split3(const void* x, int in_f32, void* out, long long rows, int K, cudaStream_t stream);
int run_heads_transpose(const float* in, int ld, int B, int T, int n_heads, float* o0, float* o1, float* o2,
                        cudaStream_t stream);

// What amt_model_load (model_load.cu) may touch of an amt_model (model.cu)
struct ModelLoadView {
  virtual const amt_model_config& cfg() const = 0;
  virtual size_t expected_bytes() const = 0;                      // sum of the packed tensors' sizes (+ alignment)
  virtual int alloc_arena(size_t bytes, uint8_t** out) = 0;        // device memory owned by the handle, freed on destroy / reload
  virtual void set(const std::string& name, void* p, size_t nbytes) = 0;
  virtual ~ModelLoadView() {}
};
int model_load_impl(ModelLoadView* view, const char* const* names, const void* const* ptrs, const int64_t* numels, int n,
                    cudaStream_t stream);

}  // namespace amt
