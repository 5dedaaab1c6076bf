// Onset / offset-aware note decoding on bit-packed rolls.
//
// CNNRNNModelLarge computes onset and offset heads next to the frame head (reference models/cnn_rnn_model.py:333-345)
// but the reference's inference path drops them (main.py:150-160 thresholds the frame head only).  This is the decoder
// those heads are trained for, in the form the Onsets-and-Frames literature uses; the reference has no counterpart, so
// the rule is DEFINED here and restated on the CPU in oracle/notes.py (group_notes_onset_aware), which the kernel must
// match bit for bit.  Per pitch, over the time axis of the concatenated segments (frame index = seg * T + t), with the
// thresholded heads F, ON, OFF:
//     sounding   Fa[t] = F[t] | ON[t]                       (an onset implies the frame sounds)
//     start      S[t]  = ON[t] & !ON[t-1]                   (rising edge of the onset head; ON[-1] = 0)
//     boundary   B[t]  = !Fa[t] | S[t] | OFF[t]             (OFF = 0 when the offset head is not given)
//   every start t opens a note (pitch, t, e) with e = the first boundary after t (e > t), or the end of the roll;
//   sounding frames that no onset opened are ignored; a new onset re-strikes (ends the running note, starts the next).
// Output like amt_bits_notes: rows (pitch, onset, offset) pitch-major, onset ascending; counts[n_pitch] = total.
//
// One warp per pitch walks the roll left to right, 32 words (1024 frames) at a time, lane = word: starts and boundaries
// are word-wide bit operations, "first boundary after" is a find-first-set inside the word or a suffix-min over the
// later lanes, and at most ONE note per pitch is open across a block / segment seam (a later start is itself a
// boundary), carried in registers.  Every start yields exactly one note, so the counting pass is a popcount.
#include "kernels.cuh"

namespace amt {

struct OnsetView {
  const uint32_t* f;
  const uint32_t* on;
  const uint32_t* off;      // may be nullptr
  int n_seg, n_pitch, T, words;
};

constexpr int kNoBoundary = 0x7fffffff;

template <bool EMIT>
__global__ void __launch_bounds__(256) onset_notes_kernel(const OnsetView v, const int32_t* __restrict__ offsets,
                                                          int32_t* __restrict__ counts, int32_t* __restrict__ notes, int cap) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (p >= v.n_pitch) return;
  const int total_frames = v.n_seg * v.T;
  int n_emitted = EMIT ? offsets[p] : 0;            // next output row of this pitch
  int n_count = 0;
  uint32_t prev_on = 0;                             // ON[t-1] of the frame before the current word block
  int pending = -1;                                 // start frame of the note that is still open, or -1
  for (int seg = 0; seg < v.n_seg; ++seg) {
    const size_t row = (static_cast<size_t>(seg) * v.n_pitch + p) * v.words;
    for (int w0 = 0; w0 < v.words; w0 += 32) {
      const int w = w0 + lane;
      const bool have = w < v.words;
      const int nbits = have ? min(32, v.T - 32 * w) : 0;
      const uint32_t valid = nbits >= 32 ? 0xffffffffu : ((1u << nbits) - 1u);
      const uint32_t fw = have ? __ldg(v.f + row + w) & valid : 0u;
      const uint32_t ow = have ? __ldg(v.on + row + w) & valid : 0u;
      const uint32_t xw = (have && v.off) ? __ldg(v.off + row + w) & valid : 0u;
      // ON[t-1] for bit 0 of this word: the last valid bit of the previous word (lane - 1) or of the previous block
      uint32_t left = __shfl_up_sync(0xffffffffu, ow >> 31, 1);
      if (lane == 0) left = prev_on;
      const uint32_t S = ow & ~((ow << 1) | left);
      const uint32_t B = (~(fw | ow) | S | xw) & valid;
      if (!EMIT) {
        n_count += __popc(S);
      } else {
        const int base = seg * v.T + 32 * w;
        // first boundary of each word, and of all LATER words of the block (exclusive suffix-min over lanes)
        const int fb = B ? base + __ffs(B) - 1 : kNoBoundary;
        int nxt = fb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int o = __shfl_down_sync(0xffffffffu, nxt, d);
          if (lane + d < 32) nxt = min(nxt, o);
        }
        const int block_first = __shfl_sync(0xffffffffu, nxt, 0);     // first boundary of the whole block
        nxt = __shfl_down_sync(0xffffffffu, nxt, 1);
        if (lane == 31) nxt = kNoBoundary;
        // the note carried in from the left ends at the block's first boundary, if it has one
        if (pending >= 0 && block_first != kNoBoundary) {
          if (lane == 0 && n_emitted < cap) {
            notes[3 * n_emitted + 0] = p;
            notes[3 * n_emitted + 1] = pending;
            notes[3 * n_emitted + 2] = block_first;
          }
          ++n_emitted;
          pending = -1;
        }
        // notes opened in this word: all but (possibly) the last one end inside the word; the last one ends at the next
        // boundary of a later word -- or stays open (only the block's very last start can: later starts are boundaries)
        const int ns = __popc(S);
        const bool last_open = ns > 0 && (B >> (31 - __clz(S)) >> 1) == 0u && nxt == kNoBoundary;
        const int mine = ns - (last_open ? 1 : 0);
        int pos = mine;                                               // exclusive prefix over lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int o = __shfl_up_sync(0xffffffffu, pos, d);
          if (lane >= d) pos += o;
        }
        const int block_total = __shfl_sync(0xffffffffu, pos, 31);
        pos = n_emitted + pos - mine;
        uint32_t s = S;
        for (int k = 0; k < mine; ++k) {
          const int b = __ffs(s) - 1;
          s &= s - 1;
          const uint32_t after = b == 31 ? 0u : (B >> (b + 1));
          const int e = after ? base + b + 1 + __ffs(after) - 1 : nxt;
          if (pos < cap) {
            notes[3 * pos + 0] = p;
            notes[3 * pos + 1] = base + b;
            notes[3 * pos + 2] = e;
          }
          ++pos;
        }
        n_emitted += block_total;
        // at most one lane holds an open start: it becomes the carried note
        const int open_start = last_open ? base + 31 - __clz(S) : -1;
        int op = open_start;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) op = max(op, __shfl_xor_sync(0xffffffffu, op, d));
        if (op >= 0) pending = op;
      }
      // ON at the last valid frame of this block -> `left` of the next one (the last word's valid bits may be < 32)
      const int last_lane = min(31, v.words - 1 - w0);
      const uint32_t top = nbits > 0 ? (ow >> (nbits - 1)) & 1u : 0u;
      prev_on = __shfl_sync(0xffffffffu, top, last_lane);
    }
  }
  if (!EMIT) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n_count += __shfl_xor_sync(0xffffffffu, n_count, d);
    if (lane == 0) counts[p] = n_count;
  } else if (pending >= 0 && lane == 0 && n_emitted < cap) {          // still sounding at the end of the roll
    notes[3 * n_emitted + 0] = p;
    notes[3 * n_emitted + 1] = pending;
    notes[3 * n_emitted + 2] = total_frames;
  }
}

// exclusive prefix of the per-pitch counts (n_pitch <= 1024) -> offsets; counts[n_pitch] = total
__global__ void __launch_bounds__(1024) onset_offsets_kernel(int32_t* __restrict__ counts, int32_t* __restrict__ offsets, int n_pitch) {
  __shared__ int s[1024];
  const int t = threadIdx.x;
  const int c = t < n_pitch ? counts[t] : 0;
  s[t] = c;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const int o = t >= d ? s[t - d] : 0;
    __syncthreads();
    s[t] += o;
    __syncthreads();
  }
  if (t < n_pitch) offsets[t] = s[t] - c;
  if (t == 1023) counts[n_pitch] = s[1023];
}

}  // namespace amt

extern "C" {

size_t amt_onset_notes_scratch_ints(int n_pitch) { return n_pitch < 1 ? 0 : static_cast<size_t>(n_pitch); }

int amt_onset_notes(const uint32_t* frame_bits, const uint32_t* onset_bits, const uint32_t* offset_bits, int n_seg, int n_pitch,
                    int T, int32_t* notes, int cap, int32_t* counts, int32_t* scratch, size_t scratch_ints, amt_stream_t stream_) {
  using namespace amt;
  AMT_REQUIRE(frame_bits && onset_bits && notes && counts && scratch, "onset_notes: NULL argument");
  AMT_REQUIRE(n_seg >= 1 && n_pitch >= 1 && n_pitch <= 1024 && T >= 1 && cap >= 0, "onset_notes: bad sizes (at most 1024 pitches)");
  AMT_REQUIRE(static_cast<long long>(n_seg) * T < (1ll << 31) - 1, "onset_notes: roll too long");
  if (scratch_ints < static_cast<size_t>(n_pitch))
    return set_error(AMT_ERR_WORKSPACE, "onset_notes: scratch of %zu ints < %d (amt_onset_notes_scratch_ints)", scratch_ints, n_pitch);
  AMT_TRY(ensure_device());
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const OnsetView v{frame_bits, onset_bits, offset_bits, n_seg, n_pitch, T, (T + 31) / 32};
  const int grid = ceil_div(n_pitch, 8);
  onset_notes_kernel<false><<<grid, 256, 0, stream>>>(v, nullptr, counts, notes, cap);
  AMT_CHECK_LAUNCH();
  onset_offsets_kernel<<<1, 1024, 0, stream>>>(counts, scratch, n_pitch);
  AMT_CHECK_LAUNCH();
  onset_notes_kernel<true><<<grid, 256, 0, stream>>>(v, scratch, counts, notes, cap);
  AMT_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
