// Integer post-processing on probability rolls:
//   * threshold + contiguous-frame note grouping (reference main.py:204-223) as a
//     ballot/popc warp-scan kernel, on the virtual concatenation of per-chunk rolls
//     (main.py:164-186) so seam-crossing notes merge exactly as in the reference;
//   * framewise TP/FP/FN for a whole threshold grid in ONE pass over the data
//     (reference scripts/evaluate.py:524-553 re-runs the model per threshold).
#include <algorithm>

#include "common.cuh"

namespace amt {

// The roll is the VIRTUAL concatenation of n_seg per-chunk rolls (main.py:164-186): frame tt of pitch p
// lives in segment tt / T at offset tt % T.  One warp owns one (pitch, segment): 88 x n_seg warps scan
// the roll in parallel (a CTA per pitch walked 60 000 frames serially); a tiny scan kernel turns the
// per-warp onset / offset counts into output ranks, so the note list comes out pitch-major and
// onset-ascending exactly as the reference loop emits it (main.py:204-223).
struct RollView {
  const float* vals;          // float roll / probabilities, or nullptr when `bits` is set
  const uint32_t* bits;       // bit-packed roll [n_seg][n_pitch][words]: bit t%32 of word t/32 (amt_pack_roll_u32)
  int n_seg, T, words, n_pitch;
  long long seg_stride, pitch_stride;
  float thr;
  // frame t of segment s, where t may be -1 (last frame of the previous segment) or T (first of the next)
  __device__ __forceinline__ bool active(int p, int s, int t) const {
    if (t < 0) { --s; t = T - 1; }
    else if (t >= T) { ++s; t = 0; }
    if (s < 0 || s >= n_seg) return false;
    if (bits) return (__ldg(bits + (static_cast<long long>(s) * n_pitch + p) * words + (t >> 5)) >> (t & 31)) & 1u;
    return __ldg(vals + s * seg_stride + p * pitch_stride + t) > thr;
  }
};

// pass 1 (EMIT = false): onset / offset counts of every (pitch, segment);  pass 3 (EMIT = true): write the
// onset frames and offset frames at their ranks.  A note is (pitch, first active frame, last active frame + 1).
template <bool EMIT>
__global__ void __launch_bounds__(256) notes_scan_kernel(RollView rv, int n_pitch, int32_t* __restrict__ cnt_on,
                                                         int32_t* __restrict__ cnt_off, int32_t* __restrict__ notes, int cap) {
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= n_pitch * rv.n_seg) return;
  const int p = w / rv.n_seg, s = w - p * rv.n_seg;
  int on_run = EMIT ? cnt_on[w] : 0, off_run = EMIT ? cnt_off[w] : 0;      // EMIT: exclusive ranks from the scan
  const unsigned lt = (1u << lane) - 1u;
  for (int t0 = 0; t0 < rv.T; t0 += 32) {
    const int t = t0 + lane;
    const bool a = t < rv.T && rv.active(p, s, t);
    const bool is_on = a && !rv.active(p, s, t - 1);
    const bool is_off = a && !rv.active(p, s, t + 1);
    const unsigned m_on = __ballot_sync(0xffffffffu, is_on), m_off = __ballot_sync(0xffffffffu, is_off);
    if (EMIT) {
      const int tt = s * rv.T + t;
      if (is_on) {
        const int idx = on_run + __popc(m_on & lt);
        if (idx < cap) { notes[3 * idx + 0] = p; notes[3 * idx + 1] = tt; }
      }
      if (is_off) {
        const int idx = off_run + __popc(m_off & lt);
        if (idx < cap) notes[3 * idx + 2] = tt + 1;
      }
    }
    on_run += __popc(m_on);
    off_run += __popc(m_off);
  }
  if (!EMIT && lane == 0) { cnt_on[w] = on_run; cnt_off[w] = off_run; }
}

// pass 2: in-place exclusive scan of both count arrays (pitch-major), per-pitch note counts and the total
__global__ void __launch_bounds__(1024) notes_rank_kernel(int32_t* __restrict__ cnt_on, int32_t* __restrict__ cnt_off, int n,
                                                          int n_seg, int n_pitch, int32_t* __restrict__ counts) {
  __shared__ int wsum[2][32];
  __shared__ int carry[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 2) carry[tid] = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + tid;
    const int v0 = i < n ? cnt_on[i] : 0, v1 = i < n ? cnt_off[i] : 0;
    int x0 = v0, x1 = v1;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int y0 = __shfl_up_sync(0xffffffffu, x0, off), y1 = __shfl_up_sync(0xffffffffu, x1, off);
      if (lane >= off) { x0 += y0; x1 += y1; }
    }
    if (lane == 31) { wsum[0][warp] = x0; wsum[1][warp] = x1; }
    __syncthreads();
    if (warp == 0) {
      int a0 = wsum[0][lane], a1 = wsum[1][lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int y0 = __shfl_up_sync(0xffffffffu, a0, off), y1 = __shfl_up_sync(0xffffffffu, a1, off);
        if (lane >= off) { a0 += y0; a1 += y1; }
      }
      wsum[0][lane] = a0;
      wsum[1][lane] = a1;
    }
    __syncthreads();
    const int b0 = carry[0] + (warp > 0 ? wsum[0][warp - 1] : 0), b1 = carry[1] + (warp > 0 ? wsum[1][warp - 1] : 0);
    if (i < n) {
      const int e0 = b0 + x0 - v0;                          // exclusive rank of this (pitch, segment)
      cnt_on[i] = e0;
      cnt_off[i] = b1 + x1 - v1;
      if (i % n_seg == 0) counts[i / n_seg] = e0;           // start rank of the pitch (differenced below)
    }
    __syncthreads();
    if (tid == 0) { carry[0] += wsum[0][31]; carry[1] += wsum[1][31]; }
    __syncthreads();
  }
  if (tid == 0) counts[n_pitch] = carry[0];
  __syncthreads();
  // start ranks -> per-pitch counts  (read all, then write: one CTA, so a barrier separates the phases)
  int mine = 0;
  if (tid < n_pitch) mine = counts[tid + 1] - counts[tid];
  __syncthreads();
  if (tid < n_pitch) counts[tid] = mine;
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

size_t amt_threshold_notes_scratch_ints(int n_seg, int n_pitch) {
  if (n_seg < 1 || n_pitch < 1) return 0;
  return 2 * static_cast<size_t>(n_seg) * static_cast<size_t>(n_pitch);
}

static int group_notes(amt::RollView rv, int n_pitch, int32_t* notes, int cap, int32_t* counts, int32_t* scratch,
                       size_t scratch_ints, cudaStream_t stream) {
  using namespace amt;
  AMT_REQUIRE(n_pitch <= 1024, "threshold_notes: at most 1024 pitches");
  AMT_REQUIRE(static_cast<long long>(rv.n_seg) * rv.T < (1ll << 31), "threshold_notes: roll too long");
  const int n = n_pitch * rv.n_seg;
  // per-(pitch, segment) onset / offset counts -> ranks, in CALLER scratch (the library owns no device memory:
  // nothing here is tied to one device, one stream or one host thread, and the call is graph-capturable)
  if (scratch_ints < 2 * static_cast<size_t>(n))
    return set_error(AMT_ERR_WORKSPACE, "threshold_notes: scratch of %zu ints < %zu (amt_threshold_notes_scratch_ints)",
                     scratch_ints, 2 * static_cast<size_t>(n));
  const int grid = ceil_div(n, 8);
  notes_scan_kernel<false><<<grid, 256, 0, stream>>>(rv, n_pitch, scratch, scratch + n, notes, cap);
  AMT_CHECK_LAUNCH();
  notes_rank_kernel<<<1, 1024, 0, stream>>>(scratch, scratch + n, n, rv.n_seg, n_pitch, counts);
  AMT_CHECK_LAUNCH();
  notes_scan_kernel<true><<<grid, 256, 0, stream>>>(rv, n_pitch, scratch, scratch + n, notes, cap);
  AMT_CHECK_LAUNCH();
  return 0;
}

int amt_threshold_notes(const float* vals, int n_seg, int n_pitch, int T, int64_t seg_stride, int64_t pitch_stride,
                        float thr, int32_t* notes, int cap, int32_t* counts, int32_t* scratch, size_t scratch_ints,
                        amt_stream_t stream_) {
  using namespace amt;
  AMT_REQUIRE(vals && notes && counts && scratch, "threshold_notes: NULL argument");
  AMT_REQUIRE(n_seg >= 1 && n_pitch >= 1 && T >= 1 && cap >= 0, "threshold_notes: bad sizes");
  AMT_TRY(ensure_device());
  RollView rv{vals, nullptr, n_seg, T, 0, n_pitch, seg_stride, pitch_stride, thr};
  return group_notes(rv, n_pitch, notes, cap, counts, scratch, scratch_ints, static_cast<cudaStream_t>(stream_));
}

int amt_bits_notes(const uint32_t* bits, int n_seg, int n_pitch, int T, int32_t* notes, int cap, int32_t* counts,
                   int32_t* scratch, size_t scratch_ints, amt_stream_t stream_) {
  using namespace amt;
  AMT_REQUIRE(bits && notes && counts && scratch, "bits_notes: NULL argument");
  AMT_REQUIRE(n_seg >= 1 && n_pitch >= 1 && T >= 1 && cap >= 0, "bits_notes: bad sizes");
  AMT_TRY(ensure_device());
  RollView rv{nullptr, bits, n_seg, T, (T + 31) / 32, n_pitch, 0, 0, 0.0f};
  return group_notes(rv, n_pitch, notes, cap, counts, scratch, scratch_ints, static_cast<cudaStream_t>(stream_));
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
  if (smem > 48 * 1024) AMT_FUNC_ATTR(f1_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
