// Polyphase FIR sample-rate conversion on the GPU (the step right before
// the hot path -- librosa.load(path, sr=16000) at reference main.py:76 resamples every recording to
// 16 kHz before it is split into 30-s chunks, main.py:82-97).
//
//   y[n] = sum_k h[k] * x_up[n*down - k + centre],   x_up = x zero-stuffed by `up`
// evaluated in polyphase form: only every up-th tap meets a non-zero sample, so output n reads the
// taps of phase (n*down + centre) % up against ~len(h)/up consecutive input samples.  HBM-bound
// (one read of x, one write of y): a CTA produces a tile of outputs from an input segment staged in
// shared memory with coalesced loads; taps are read through the read-only cache.  The taps are an
// argument, so the host chooses the filter (the Python side uses scipy-compatible Kaiser-windowed sinc).
#include "common.cuh"

namespace amt {

constexpr int kRsTile = 1024;        // outputs per CTA
constexpr int kRsThreads = 256;

__global__ void __launch_bounds__(kRsThreads)
resample_poly_kernel(const float* __restrict__ x, long long n_in, float* __restrict__ y, long long n_out,
                     const float* __restrict__ h, int n_taps, int up, int down, int centre, int seg_len) {
  extern __shared__ float s_x[];
  const long long o0 = static_cast<long long>(blockIdx.x) * kRsTile;
  const long long o1 = min(o0 + kRsTile, n_out);
  const int taps_per_phase = (n_taps + up - 1) / up;
  // input index of the newest sample output n needs: floor((n*down + centre) / up); it reaches back
  // taps_per_phase - 1 samples
  const long long hi = ((o1 - 1) * down + centre) / up;
  const long long lo = (o0 * down + centre) / up - (taps_per_phase - 1);
  const int seg = static_cast<int>(hi - lo + 1);            // <= seg_len
  for (int i = threadIdx.x; i < seg; i += kRsThreads) {
    const long long g = lo + i;
    s_x[i] = (g >= 0 && g < n_in) ? __ldg(x + g) : 0.0f;
  }
  __syncthreads();
  for (long long n = o0 + threadIdx.x; n < o1; n += kRsThreads) {
    const long long pos = n * down + centre;
    const int phase = static_cast<int>(pos % up);
    const int newest = static_cast<int>(pos / up - lo);
    float acc = 0.0f;
    int k = phase;
    for (int j = 0; k < n_taps; ++j, k += up) acc = fmaf(__ldg(h + k), s_x[newest - j], acc);
    y[n] = acc;
  }
  (void)seg_len;
}

// 16-bit PCM frames [n_frames][channels] (interleaved, as they lie in a WAVE file) -> mono float32 in [-1, 1):
// sample / 32768 (exact in fp32), channels averaged as numpy's float32 mean does (exact sum of 16-bit values, one
// division).  Uploading int16 and converting here moves half the bytes over PCIe and skips the host conversion.
__global__ void __launch_bounds__(256) pcm16_to_mono_kernel(const int16_t* __restrict__ pcm, long long n_frames, int ch,
                                                            float* __restrict__ out) {
  const float n_ch = static_cast<float>(ch);
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n_frames; i += 256ll * gridDim.x) {
    float sum = 0.0f;
    for (int c = 0; c < ch; ++c) sum += static_cast<float>(__ldg(pcm + i * ch + c)) * (1.0f / 32768.0f);
    out[i] = ch == 1 ? sum : sum / n_ch;
  }
}

}  // namespace amt

extern "C" int amt_resample_poly_f32(const float* x, int64_t n_in, float* y, int64_t n_out, const float* taps,
                                     int n_taps, int up, int down, amt_stream_t stream_) {
  using namespace amt;
  AMT_REQUIRE(x && y && taps, "resample: NULL argument");
  AMT_REQUIRE(n_in > 0 && n_out > 0 && up >= 1 && down >= 1 && n_taps >= 1 && (n_taps & 1), "resample: bad sizes (odd tap count)");
  AMT_TRY(ensure_device());
  const int centre = (n_taps - 1) / 2;                       // zero-phase: output 0 is aligned with input 0
  const int taps_per_phase = (n_taps + up - 1) / up;
  const long long span = (static_cast<long long>(kRsTile - 1) * down + up - 1) / up + taps_per_phase + 2;
  AMT_REQUIRE(span * 4 <= 200 * 1024, "resample: ratio %d/%d with %d taps needs too large an input tile", up, down, n_taps);
  const size_t smem = static_cast<size_t>(span) * 4;
  if (smem > 48 * 1024) AMT_FUNC_ATTR(resample_poly_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const long long grid = (n_out + kRsTile - 1) / kRsTile;
  AMT_REQUIRE(grid < (1ll << 31), "resample: output too long");
  resample_poly_kernel<<<static_cast<unsigned>(grid), kRsThreads, smem, static_cast<cudaStream_t>(stream_)>>>(
      x, n_in, y, n_out, taps, n_taps, up, down, centre, static_cast<int>(span));
  AMT_CHECK_LAUNCH();
  return 0;
}

extern "C" int amt_pcm16_to_mono_f32(const int16_t* pcm, int64_t n_frames, int channels, float* out, amt_stream_t stream_) {
  using namespace amt;
  AMT_REQUIRE(pcm && out && n_frames >= 0 && channels >= 1 && channels <= 64, "pcm16_to_mono: bad arguments");
  AMT_TRY(ensure_device());
  if (n_frames == 0) return 0;
  const long long want = (n_frames + 256 * 4 - 1) / (256 * 4);
  const int grid = static_cast<int>(want > 8ll * num_sms() ? 8ll * num_sms() : want);
  pcm16_to_mono_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(pcm, n_frames, channels, out);
  AMT_CHECK_LAUNCH();
  return 0;
}
