// Integer post-processing on probability rolls:
//   * threshold + contiguous-frame note grouping (reference main.py:204-223) as a
//     ballot/popc warp-scan kernel, on the virtual concatenation of per-chunk rolls
//     (main.py:164-186) so seam-crossing notes merge exactly as in the reference;
//   * framewise TP/FP/FN for a whole threshold grid in ONE pass over the data
//     (reference scripts/evaluate.py:524-553 re-runs the model per threshold).
#include "common.cuh"

namespace amt {

struct RollView {
  const float* vals;
  int n_seg, T;
  long long seg_stride, pitch_stride;
  float thr;
  __device__ __forceinline__ bool active(int p, long long tt, long long total) const {
    if (tt < 0 || tt >= total) return false;
    const int seg = static_cast<int>(tt / T);
    const int t = static_cast<int>(tt - static_cast<long long>(seg) * T);
    return __ldg(vals + seg * seg_stride + p * pitch_stride + t) > thr;
  }
};

__global__ void __launch_bounds__(256) notes_count_kernel(RollView rv, int32_t* __restrict__ counts) {
  const int p = blockIdx.x;
  const long long total = static_cast<long long>(rv.n_seg) * rv.T;
  int local = 0;
  for (long long tt = threadIdx.x; tt < total; tt += 256)
    local += (rv.active(p, tt, total) && !rv.active(p, tt - 1, total)) ? 1 : 0;
  __shared__ int wsum[8];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int i = 0; i < 8; ++i) s += wsum[i];
    counts[p] = s;
  }
}

__global__ void __launch_bounds__(256) notes_emit_kernel(RollView rv, const int32_t* __restrict__ counts_in,
                                                         int32_t* __restrict__ counts_total, int n_pitch,
                                                         int32_t* __restrict__ notes, int cap) {
  const int p = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long total = static_cast<long long>(rv.n_seg) * rv.T;
  __shared__ int s_base;
  __shared__ int w_on[8], w_off[8];
  if (warp == 0) {
    int s = 0;
    for (int i = lane; i < p; i += 32) s += counts_in[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) {
      s_base = s;
      if (p == n_pitch - 1) counts_total[0] = s + counts_in[p];
    }
  }
  __syncthreads();
  const int base = s_base;
  int on_run = 0, off_run = 0;                 // block-uniform running ranks
  for (long long c0 = 0; c0 < total; c0 += 256) {
    const long long tt = c0 + tid;
    const bool a = rv.active(p, tt, total);
    const bool is_on = a && !rv.active(p, tt - 1, total);
    const bool is_off = a && !rv.active(p, tt + 1, total);     // note ends after frame tt -> offset index tt + 1
    const unsigned m_on = __ballot_sync(0xffffffffu, is_on), m_off = __ballot_sync(0xffffffffu, is_off);
    if (lane == 0) { w_on[warp] = __popc(m_on); w_off[warp] = __popc(m_off); }
    __syncthreads();
    int pre_on = 0, pre_off = 0, tot_on = 0, tot_off = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      pre_on += i < warp ? w_on[i] : 0;
      pre_off += i < warp ? w_off[i] : 0;
      tot_on += w_on[i];
      tot_off += w_off[i];
    }
    const unsigned lt = (1u << lane) - 1u;
    if (is_on) {
      const int idx = base + on_run + pre_on + __popc(m_on & lt);
      if (idx < cap) { notes[3 * idx + 0] = p; notes[3 * idx + 1] = static_cast<int32_t>(tt); }
    }
    if (is_off) {
      const int idx = base + off_run + pre_off + __popc(m_off & lt);
      if (idx < cap) notes[3 * idx + 2] = static_cast<int32_t>(tt + 1);
    }
    on_run += tot_on;
    off_run += tot_off;
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------
// TP/FP/FN over a sorted threshold grid, one read of probs + targets.
// For every cell: k = #thresholds strictly below p  ->  predicted positive for
// thresholds j < k.  Histogram k per class, then suffix sums.
// ----------------------------------------------------------------------------
constexpr int kF1MaxThr = 512;

__global__ void __launch_bounds__(256) f1_counts_kernel(const float* __restrict__ probs, const float* __restrict__ target,
                                                        const int32_t* __restrict__ lengths, int n_pitch, int T_stride,
                                                        const float* __restrict__ thresholds, int n_thr,
                                                        unsigned long long* __restrict__ out) {
  extern __shared__ uint32_t s_hist[];                 // [8 warps][2][n_thr+1] then thresholds
  const int nb = n_thr + 1;
  float* s_thr = reinterpret_cast<float*>(s_hist + 8 * 2 * nb);
  const int piece = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 8 * 2 * nb; i += 256) s_hist[i] = 0u;
  for (int i = tid; i < n_thr; i += 256) s_thr[i] = thresholds[i];
  __syncthreads();
  const int L = min(max(lengths[piece], 0), T_stride);
  const long long cells = static_cast<long long>(n_pitch) * L;
  const float* pp = probs + static_cast<size_t>(piece) * n_pitch * T_stride;
  const float* yy = target + static_cast<size_t>(piece) * n_pitch * T_stride;
  uint32_t* my = s_hist + warp * 2 * nb;
  for (long long e = blockIdx.x * 256ll + tid; e < cells; e += gridDim.x * 256ll) {
    const int p = static_cast<int>(e / L), t = static_cast<int>(e - static_cast<long long>(p) * L);
    const float pr = __ldg(pp + static_cast<size_t>(p) * T_stride + t);
    const int y = __ldg(yy + static_cast<size_t>(p) * T_stride + t) > 0.5f ? 1 : 0;
    int lo = 0, hi = n_thr;                              // first index with thr >= pr  == #thr < pr
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_thr[mid] < pr) lo = mid + 1; else hi = mid;
    }
    atomicAdd(&my[y * nb + lo], 1u);
  }
  __syncthreads();
  // reduce warps into warp 0's histogram
  for (int i = tid; i < 2 * nb; i += 256) {
    uint32_t s = 0;
    for (int w = 0; w < 8; ++w) s += s_hist[w * 2 * nb + i];
    s_hist[i] = s;
  }
  __syncthreads();
  for (int j = tid; j < n_thr; j += 256) {
    unsigned long long tp = 0, fp = 0, pos = 0;
    for (int k = 0; k < nb; ++k) {
      pos += s_hist[nb + k];
      if (k > j) { tp += s_hist[nb + k]; fp += s_hist[k]; }
    }
    unsigned long long* o = out + (static_cast<size_t>(piece) * n_thr + j) * 3;
    if (tp) atomicAdd(o + 0, tp);
    if (fp) atomicAdd(o + 1, fp);
    if (pos - tp) atomicAdd(o + 2, pos - tp);
  }
}

}  // namespace amt

extern "C" {

int amt_threshold_notes(const float* vals, int n_seg, int n_pitch, int T, int64_t seg_stride, int64_t pitch_stride,
                        float thr, int32_t* notes, int cap, int32_t* counts, amt_stream_t stream_) {
  using namespace amt;
  AMT_REQUIRE(vals && notes && counts, "threshold_notes: NULL argument");
  AMT_REQUIRE(n_seg >= 1 && n_pitch >= 1 && T >= 1 && cap >= 0, "threshold_notes: bad sizes");
  AMT_TRY(ensure_device());
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RollView rv{vals, n_seg, T, seg_stride, pitch_stride, thr};
  notes_count_kernel<<<n_pitch, 256, 0, stream>>>(rv, counts);
  AMT_CHECK_LAUNCH();
  notes_emit_kernel<<<n_pitch, 256, 0, stream>>>(rv, counts, counts + n_pitch, n_pitch, notes, cap);
  AMT_CHECK_LAUNCH();
  return 0;
}

int amt_f1_counts(const float* probs, const float* target, const int32_t* lengths, int n_pieces, int n_pitch,
                  int T_stride, const float* thresholds, int n_thr, int64_t* out, amt_stream_t stream_) {
  using namespace amt;
  AMT_REQUIRE(probs && target && lengths && thresholds && out, "f1_counts: NULL argument");
  AMT_REQUIRE(n_pieces >= 1 && n_pitch >= 1 && T_stride >= 1, "f1_counts: bad sizes");
  AMT_REQUIRE(n_thr >= 1 && n_thr <= kF1MaxThr, "f1_counts: n_thr must be in 1..%d", kF1MaxThr);
  AMT_REQUIRE(n_pieces <= 65535, "f1_counts: at most 65535 pieces per call");
  AMT_TRY(ensure_device());
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  AMT_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t) * 3 * static_cast<size_t>(n_pieces) * n_thr, stream));
  const size_t smem = sizeof(uint32_t) * 8 * 2 * (n_thr + 1) + sizeof(float) * n_thr;
  if (smem > 48 * 1024) AMT_CUDA(cudaFuncSetAttribute(f1_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long cells = static_cast<long long>(n_pitch) * T_stride;
  int bx = static_cast<int>(std::min<long long>((cells + 256 * 8 - 1) / (256 * 8), 64));
  if (bx < 1) bx = 1;
  dim3 grid(bx, n_pieces);
  f1_counts_kernel<<<grid, 256, smem, stream>>>(probs, target, lengths, n_pitch, T_stride, thresholds, n_thr,
                                                reinterpret_cast<unsigned long long*>(out));
  AMT_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
