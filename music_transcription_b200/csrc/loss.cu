// Validation-loss VALUE of the reference's TranscriptionModel.compute_loss for the CNN-RNN models
// (reference models/transcription_model.py:110-217): binary cross-entropy with logits against the piano roll,
// optionally masked to each sample's valid frames, and for the three-head Large model the weighted sum
// 0.5 frame + 0.25 onset + 0.25 offset with onset / offset targets derived from the roll on the fly.
// Forward only: training (autograd) stays outside the hot path.
//
// One pass over the roll: every thread handles (b, p, t) cells grid-stride, reads the target and its two
// time neighbours and up to three logits, accumulates the three loss sums in fp64, block-reduces and issues
// one atomicAdd(double) per head and block; a one-thread kernel turns the sums into the fp32 scalar.
// HBM bound: (heads + 1) x 4 bytes per cell.
#include "kernels.cuh"

namespace amt {

struct LossParams {
  const float* logits[3];      // frame, onset, offset: [B][P][Tl]; onset/offset NULL for a single head
  const float* targets;        // [B][P][Tt]
  const int* lengths;          // [B] or NULL
  int B, P, Tl, Tt, heads;
  float scale;                 // Tl / Tt (F.interpolate linear, align_corners=False, when Tl != Tt)
  double* acc;                 // [0..2] loss sums, [3] valid (b, t) pairs
};

// F.interpolate(mode="linear", align_corners=False) of one row at output index t (fp32 index arithmetic as ATen's
// area_pixel_compute_source_index): src = max((t + 0.5) * scale - 0.5, 0)
__device__ __forceinline__ float logit_at(const float* row, int t, int Tl, int Tt, float scale) {
  if (Tl == Tt) return __ldg(row + t);
  float src = (static_cast<float>(t) + 0.5f) * scale - 0.5f;
  src = src < 0.0f ? 0.0f : src;
  int i0 = static_cast<int>(src);
  i0 = i0 > Tl - 1 ? Tl - 1 : i0;
  const int i1 = i0 + (i0 < Tl - 1 ? 1 : 0);
  const float lam1 = src - static_cast<float>(i0), lam0 = 1.0f - lam1;
  return lam0 * __ldg(row + i0) + lam1 * __ldg(row + i1);
}

// max(x, 0) - x*y + log(1 + exp(-|x|))
__device__ __forceinline__ float bce_with_logits(float x, float y) {
  return fmaxf(x, 0.0f) - x * y + log1pf(expf(-fabsf(x)));
}

__global__ void __launch_bounds__(256) bce_loss_kernel(const LossParams p) {
  const long long n = static_cast<long long>(p.B) * p.P * p.Tt;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) {
    const int t = static_cast<int>(i % p.Tt);
    const long long row = i / p.Tt;                       // b * P + pitch
    const int b = static_cast<int>(row / p.P);
    if (p.lengths && t >= __ldg(p.lengths + b)) continue;
    const float* trow = p.targets + row * p.Tt;
    const float y = __ldg(trow + t);
    s0 += bce_with_logits(logit_at(p.logits[0] + row * p.Tl, t, p.Tl, p.Tt, p.scale), y);
    if (p.heads == 3) {
      // onset: the roll rises into t (0 at t = 0); offset: it falls after t (0 at the last frame)
      const float y_on = t > 0 ? fmaxf(y - __ldg(trow + t - 1), 0.0f) : 0.0f;
      const float y_off = t < p.Tt - 1 ? fmaxf(y - __ldg(trow + t + 1), 0.0f) : 0.0f;
      s1 += bce_with_logits(logit_at(p.logits[1] + row * p.Tl, t, p.Tl, p.Tt, p.scale), y_on);
      s2 += bce_with_logits(logit_at(p.logits[2] + row * p.Tl, t, p.Tl, p.Tt, p.scale), y_off);
    }
  }
  __shared__ double red[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; red[2][warp] = s2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    if (threadIdx.x < p.heads) atomicAdd(p.acc + threadIdx.x, s);
  }
  // valid (b, t) pairs: sum_b clamp(lengths[b], 0, Tt)  (mask.sum() of the reference)
  if (blockIdx.x == 0 && threadIdx.x == 32) {
    double c = 0.0;
    for (int b = 0; b < p.B; ++b) {
      int l = p.lengths ? p.lengths[b] : p.Tt;
      l = l < 0 ? 0 : (l > p.Tt ? p.Tt : l);
      c += l;
    }
    p.acc[3] = c;
  }
}

// loss_h = sum_h / max(pairs * P, 1) when masked, sum_h / (B * P * Tt) otherwise (nn.BCEWithLogitsLoss mean);
// three heads: 0.5 frame + 0.25 onset + 0.25 offset (reference :187-189), combined in fp32 like the reference
__global__ void bce_finalize_kernel(const double* acc, int heads, int P, int masked, float* out) {
  double denom = acc[3] * P;
  if (masked && denom < 1.0) denom = 1.0;
  const float f = static_cast<float>(acc[0] / denom);
  if (heads == 3) {
    const float on = static_cast<float>(acc[1] / denom), off = static_cast<float>(acc[2] / denom);
    out[0] = 0.5f * f + 0.25f * on + 0.25f * off;
    out[1] = f; out[2] = on; out[3] = off;
  } else {
    out[0] = f;
    out[1] = f; out[2] = 0.0f; out[3] = 0.0f;
  }
}

int run_bce_loss(const float* frame, const float* onset, const float* offset, const float* targets, const int* lengths,
                 int B, int P, int Tl, int Tt, double* acc, float* out, cudaStream_t stream) {
  AMT_TRY(ensure_device());
  AMT_REQUIRE(frame && targets && acc && out, "bce_loss: NULL pointer");
  AMT_REQUIRE((onset == nullptr) == (offset == nullptr), "bce_loss: onset and offset logits come together");
  AMT_REQUIRE(B >= 1 && P >= 1 && Tl >= 1 && Tt >= 1, "bce_loss: empty input");
  LossParams p{};
  p.logits[0] = frame; p.logits[1] = onset; p.logits[2] = offset;
  p.targets = targets;
  p.lengths = lengths;
  p.B = B; p.P = P; p.Tl = Tl; p.Tt = Tt;
  p.heads = onset ? 3 : 1;
  p.scale = static_cast<float>(Tl) / static_cast<float>(Tt);
  p.acc = acc;
  AMT_CUDA(cudaMemsetAsync(acc, 0, 4 * sizeof(double), stream));
  const long long n = static_cast<long long>(B) * P * Tt;
  const long long want = (n + 256 * 8 - 1) / (256 * 8);
  const int grid = static_cast<int>(want < 1 ? 1 : (want > 4ll * num_sms() ? 4ll * num_sms() : want));
  bce_loss_kernel<<<grid, 256, 0, stream>>>(p);
  AMT_CHECK_LAUNCH();
  bce_finalize_kernel<<<1, 1, 0, stream>>>(acc, p.heads, P, lengths != nullptr, out);
  AMT_CHECK_LAUNCH();
  return 0;
}

}  // namespace amt

extern "C" int amt_bce_loss(const float* frame, const float* onset, const float* offset, const float* targets,
                            const int32_t* lengths, int B, int P, int T_logits, int T_targets, double* acc, float* out,
                            amt_stream_t stream) {
  return amt::run_bce_loss(frame, onset, offset, targets, lengths, B, P, T_logits, T_targets, acc, out,
                           static_cast<cudaStream_t>(stream));
}
