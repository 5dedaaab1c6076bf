/*
 * amt.h -- C ABI of libamt_sm100.so: the B200 (sm_100a) implementation of the
 * audio -> piano-roll inference path of cs4247/music-transcription.
 *
 * The reference has no FFI layer (it is pure Python); each entry point below
 * replaces the arithmetic behind one reference call site, cited as file:line
 * into the reference tree.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative amt_status; it never
 *     throws or aborts across the boundary.  amt_last_error() returns a
 *     thread-local message for the last failure on the calling thread.
 *   - all data pointers are DEVICE pointers unless the name says "host";
 *     they are borrowed for the duration of the call.  The library allocates
 *     nothing the caller must free except opaque handles (*_destroy).
 *   - `stream` is a cudaStream_t passed as void*.  Calls are asynchronous with
 *     respect to the host unless stated otherwise.
 *   - there is no CPU fallback: a missing device or a non-sm_100 device is an
 *     error (AMT_ERR_DEVICE).
 */
#ifndef AMT_H_
#define AMT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* amt_stream_t;

enum amt_status {
  AMT_OK = 0,
  AMT_ERR_ARG = -1,       /* bad argument / unsupported configuration (ValueError in Python) */
  AMT_ERR_DEVICE = -2,    /* no CUDA device, or not compute capability 10.x */
  AMT_ERR_CUDA = -3,      /* a CUDA runtime / driver call failed */
  AMT_ERR_STATE = -4,     /* handle not fully initialised (missing weights, ...) */
  AMT_ERR_WORKSPACE = -5  /* caller-provided workspace too small */
};

/* ---- library ------------------------------------------------------------ */
const char* amt_version(void);
const char* amt_last_error(void);
/* 0 when the current device can run the kernels (compute capability 10.x). */
int amt_device_check(void);
/* Number of CUDA kernels this library has launched in this process (all threads). */
uint64_t amt_launch_count(void);

/* ---- log-mel frontend --------------------------------------------------- */
/* Replaces librosa.feature.melspectrogram + librosa.power_to_db as called at
 * reference main.py:117-125 (and data/dataset.py:155-156,195-196): centred,
 * zero-padded STFT (periodic Hann, n_fft), |X|^2, slaney mel filterbank
 * (fmin..fmax, slaney norm), 10*log10(max(1e-10, .)), floor at max-top_db per
 * chunk. */
typedef struct amt_frontend amt_frontend;
int amt_frontend_create(int sr, int n_fft, int hop, int n_mels, double fmin, double fmax,
                        amt_frontend** out);
int amt_frontend_destroy(amt_frontend* fe);
/* 1 + n_samples / hop  (centred STFT frame count; 938 for a 30-s chunk). */
int amt_frontend_num_frames(const amt_frontend* fe, int n_samples);
/* Host copy of the (n_mels x (1+n_fft/2)) float32 filterbank, for inspection. */
int amt_frontend_filterbank_host(const amt_frontend* fe, float* out_host);
/* wav [B][n_samples] f32 (row stride wav_stride elements) -> out_db [B][n_mels][T] f32.
 * chunk_max: [B] f32 scratch (device).  top_db < 0 disables the floor. */
int amt_logmel_f32(amt_frontend* fe, const float* wav, int B, int n_samples, int64_t wav_stride,
                   float* out_db, float top_db, float* chunk_max, amt_stream_t stream);

/* ---- model forward ------------------------------------------------------ */
/* Replaces TranscriptionModel.forward -> CNNRNNModel / CNNRNNModelLarge.forward
 * (reference models/transcription_model.py:91-108, models/cnn_rnn_model.py:57-74
 * and :262-349), eval mode. */
enum amt_model_kind { AMT_MODEL_CNN_RNN = 0, AMT_MODEL_CNN_RNN_LARGE = 1 };
/* Arithmetic of the tensor-core contractions (the reference is fp32 throughout, models/cnn_rnn_model.py:69-70,309-311):
 *   FAST    bf16 operands, fp32 accumulation: probabilities within 2e-3 of the fp32 reference at the canonical config;
 *   PRECISE split-bf16 operands x = hi + lo (16 mantissa bits), three products hi*hi + lo*hi + hi*lo per contraction
 *           (the K axis of every conv / linear tripled), fp32 accumulation: within 3e-4, at about a third of the speed.
 * Packed-weight layouts differ between the two. */
enum amt_precision { AMT_PRECISION_FAST = 0, AMT_PRECISION_PRECISE = 1 };

typedef struct amt_model_config {
  int kind;             /* amt_model_kind */
  int n_mels;
  int hidden;           /* LSTM hidden size: multiple of 128, <= 640 (<= 512 with attention) */
  int layers;           /* layers of the main BiLSTM */
  int heads;            /* attention heads (8 in the reference) */
  int use_attention;    /* large only */
  int use_onset_offset; /* large only: shared_fc + frame/onset/offset heads */
  int precision;        /* amt_precision */
} amt_model_config;

typedef struct amt_model amt_model;
int amt_model_create(const amt_model_config* cfg, amt_model** out);
int amt_model_destroy(amt_model* m);
/* Register one packed weight tensor (device pointer, borrowed until destroy or
 * replacement).  Names and layouts: DESIGN.md "Packed weights";
 * music_transcription_b200/packing.py produces them from a reference .pth. */
int amt_model_set_tensor(amt_model* m, const char* name, const void* dev_ptr, size_t nbytes);
/* Verifies every tensor the configuration needs is present with the right size. */
int amt_model_finalize(amt_model* m);
/* Load a reference checkpoint: the n tensors of TranscriptionModel.state_dict() (reference models/cnn_rnn_model.py:28-55,
 * :179-260; keys as torch.save writes them, e.g. "model.res_block1.conv1.weight") given as DEVICE pointers to contiguous
 * float32 data with their element counts.  The library packs them on the device into memory owned by the handle -- eval
 * BatchNorm folded into the convolutions, (kf, kt, cin) K order with the residual 1x1 skip conv appended, slice-ordered
 * LSTM gate rows, b_ih + b_hh, permuted layer-0 columns, stacked + padded heads, bf16 (or the split-bf16 layout of
 * AMT_PRECISION_PRECISE) -- and finalizes the handle.  Integer tensors (num_batches_tracked) and unknown keys are
 * ignored; a missing or mis-sized tensor is AMT_ERR_STATE.  The sources are borrowed only until the call returns (it
 * synchronises `stream`).  Replaces amt_model_set_tensor + amt_model_finalize for hosts without the Python packer. */
int amt_model_load(amt_model* m, const char* const* names, const void* const* ptrs, const int64_t* numels, int n,
                   amt_stream_t stream);
/* The packed tensor registered under `name` ("res1.c2.w", "rnn0.wih", ... -- DESIGN.md "Packed weights"). */
int amt_model_get_tensor(const amt_model* m, const char* name, const void** dev_ptr, size_t* nbytes);
size_t amt_model_workspace_bytes(const amt_model* m, int B, int T);
/* Where an intermediate tensor of amt_model_forward(m, ., B, T, ...) lives inside the caller's workspace: byte offset
 * and extent of buffer `name` ("act1", "feat", "gx", "seq_a", "seq_b", "rnn_f32", "qkv", "att", "proj", "normed",
 * "shared", "logits", ... -- DESIGN.md section 3).  The workspace is the caller's memory, so after a forward the
 * tensors can be read back from it: this is how the parity tests attribute an output difference to a stage
 * (reference internals: models/cnn_rnn_model.py:262-349).  AMT_ERR_ARG for a buffer the configuration lacks. */
int amt_model_workspace_layout(const amt_model* m, int B, int T, const char* name, size_t* offset, size_t* nbytes);
/* logmel [B][1][n_mels][T] f32 -> frame/onset/offset logits [B][88][T] f32.
 * onset/offset may be NULL.  workspace: device scratch of at least
 * amt_model_workspace_bytes(m,B,T) bytes, 1024-byte aligned. */
int amt_model_forward(amt_model* m, const float* logmel, int B, int T, float* frame, float* onset,
                      float* offset, void* workspace, size_t workspace_bytes, amt_stream_t stream);
/* The same forward fed by amt_logmel_f32(..., top_db < 0, chunk_max): `logmel` is the UNFLOORED dB spectrogram and the
 * per-chunk floor max(x, chunk_max[b] - top_db) of librosa.power_to_db (reference main.py:125) is applied by the stem
 * convolution as it loads its input -- bitwise the result of flooring first, without the extra read + write of the
 * spectrogram.  chunk_max: [B] f32 device (as written by amt_logmel_f32); NULL = `logmel` is already floored. */
int amt_model_forward_db(amt_model* m, const float* logmel, const float* chunk_max, float top_db, int B, int T,
                         float* frame, float* onset, float* offset, void* workspace, size_t workspace_bytes,
                         amt_stream_t stream);

/* Per-stage device timing of amt_model_forward (CUDA events on the caller's stream around every
 * kernel launch).  enable != 0 starts/resets accumulation; read synchronises the pending events
 * and returns the number of stages, filling names (cap x 32 chars), total ms and launch counts. */
int amt_model_profile_enable(amt_model* m, int enable);
int amt_model_profile_read(amt_model* m, char* names, float* total_ms, int* launches, int cap);
/* Without blocking: the first profiled stage (since the last read / forward) whose end event has not completed yet --
 * its position in launch order, name copied to `name`; -1 when nothing is in flight.  Watchdog / hang diagnosis. */
int amt_model_profile_in_flight(amt_model* m, char* name, int cap);

/* ---- validation loss value (forward only) --------- */
/* TranscriptionModel.compute_loss for the CNN-RNN models (reference models/transcription_model.py:110-217):
 * mean binary cross-entropy with logits of frame [B][P][T_logits] against targets [B][P][T_targets]
 * (logits linearly interpolated along time like F.interpolate(mode="linear", align_corners=False) when the
 * lengths differ, :140-142); with lengths[B] != NULL only frames t < lengths[b] count and the sum is divided
 * by max(valid_frames * P, 1) (:148-163).  With onset and offset logits (both or neither) the result is
 * 0.5 frame + 0.25 onset + 0.25 offset against onset / offset targets derived from the roll (:165-190).
 * acc: 4 doubles of device scratch; out: 4 floats on the device = {loss, frame, onset, offset}.  Asynchronous. */
int amt_bce_loss(const float* frame, const float* onset, const float* offset, const float* targets,
                 const int32_t* lengths, int B, int P, int T_logits, int T_targets, double* acc, float* out,
                 amt_stream_t stream);

/* ---- sigmoid / threshold / notes ---------------------------------------- */
/* probs = sigmoid(logits); roll = (probs > thr) as float {0,1}
 * (reference main.py:153-156, models/transcription_model.py:263-266).
 * probs and/or roll may be NULL. */
int amt_sigmoid_threshold(const float* logits, int64_t n, float thr, float* probs, float* roll,
                          amt_stream_t stream);

/* Bit-packed roll: row r of vals [n_rows][T] (one pitch of one chunk) -> bits [n_rows][ceil(T/32)] uint32,
 * bit (t % 32) of word (t / 32) = (v > thr) with v = sigmoid(vals) when apply_sigmoid != 0, else vals --
 * the same float32 strict compare (and the same sigmoid) as amt_sigmoid_threshold, 1/32 of the bytes:
 * what the streaming path downloads instead of the float roll of main.py:153-160. */
int amt_pack_roll_u32(const float* vals, int64_t n_rows, int T, float thr, int apply_sigmoid, uint32_t* bits,
                      amt_stream_t stream);

/* Note grouping of reference main.py:204-223 on the roll formed by concatenating
 * `n_seg` segments along time (main.py:164-186):
 *   x[p][seg*T + t] = vals[seg*seg_stride + p*pitch_stride + t],  active iff x > thr
 * (float32 strict compare; pass thr = 0 for an already-binarised roll).
 * Output rows (pitch_idx, onset_frame, offset_frame), pitch-major then onset
 * ascending.  notes: int32 [cap][3]; counts: int32 [n_pitch + 1] -- per-pitch
 * note counts and, last, the total (which may exceed cap: then only the first
 * `cap` rows were written).  scratch: device int32 [amt_threshold_notes_scratch_ints(n_seg, n_pitch)]
 * (= 2 * n_seg * n_pitch: onset / offset counts per (pitch, segment)), owned by the caller like every
 * other buffer -- the library keeps no device memory of its own, so calls on different devices,
 * streams or host threads never share state.  Asynchronous. */
size_t amt_threshold_notes_scratch_ints(int n_seg, int n_pitch);
int amt_threshold_notes(const float* vals, int n_seg, int n_pitch, int T, int64_t seg_stride,
                        int64_t pitch_stride, float thr, int32_t* notes, int cap, int32_t* counts,
                        int32_t* scratch, size_t scratch_ints, amt_stream_t stream);

/* The same grouping on BIT-PACKED rolls (amt_pack_roll_u32): bits [n_seg][n_pitch][ceil(T/32)], segments concatenated
 * along time.  This is the form in which rolls cross NVLink: every rank all-gathers its chunks' packed rolls (10.6 KB per
 * chunk) and ONE grouping pass over the gathered roll yields the recording's note list, seams included. */
int amt_bits_notes(const uint32_t* bits, int n_seg, int n_pitch, int T, int32_t* notes, int cap, int32_t* counts,
                   int32_t* scratch, size_t scratch_ints, amt_stream_t stream);

/* Onset / offset-aware note decoding on bit-packed rolls of the three heads of CNNRNNModelLarge
 * (reference models/cnn_rnn_model.py:333-345 computes them; its inference path, main.py:150-160, thresholds the frame
 * head only and has no such decoder -- the rule is defined here and restated in oracle/notes.py).  frame_bits /
 * onset_bits / offset_bits (offset_bits may be NULL): [n_seg][n_pitch][ceil(T/32)] from amt_pack_roll_u32, segments
 * concatenated along time.  Per pitch: a rising edge of the onset head starts a note; it ends at the first later frame
 * where neither frame nor onset head is active, the offset head is active, or another onset starts -- or at the end of
 * the roll; sounding frames no onset opened are ignored.  Output and counts as amt_bits_notes.  scratch: device int32
 * [amt_onset_notes_scratch_ints(n_pitch)].  n_pitch <= 1024.  Asynchronous. */
size_t amt_onset_notes_scratch_ints(int n_pitch);
int amt_onset_notes(const uint32_t* frame_bits, const uint32_t* onset_bits, const uint32_t* offset_bits, int n_seg,
                    int n_pitch, int T, int32_t* notes, int cap, int32_t* counts, int32_t* scratch, size_t scratch_ints,
                    amt_stream_t stream);

/* Framewise TP/FP/FN of reference scripts/evaluate.py:524-553 for every piece
 * and every threshold in one pass.  probs/target: [n_pieces][n_pitch][T_stride]
 * f32, only the first lengths[i] frames of piece i count; thresholds: sorted
 * ascending, float32; out: int64 [n_pieces][n_thr][3] = (TP, FP, FN), overwritten. */
int amt_f1_counts(const float* probs, const float* target, const int32_t* lengths, int n_pieces,
                  int n_pitch, int T_stride, const float* thresholds, int n_thr, int64_t* out,
                  amt_stream_t stream);

/* ---- audio decode helper ----------------------------- */
/* Interleaved 16-bit PCM frames pcm [n_frames][channels] (device) -> mono float32 out [n_frames]: sample / 32768,
 * channels averaged in float32 -- bit-identical to decoding on the host and taking numpy's float32 mean
 * (librosa.load(..., mono=True) at reference main.py:76), at half the upload bytes. */
int amt_pcm16_to_mono_f32(const int16_t* pcm, int64_t n_frames, int channels, float* out, amt_stream_t stream);

/* ---- sample-rate conversion -------------------------- */
/* Polyphase FIR resampling y = decimate_down(filter_h(zero_stuff_up(x))), zero phase (output 0 is
 * aligned with input 0), i.e. scipy.signal.resample_poly(x, up, down, window=taps) without its gain /
 * trimming conventions -- the caller passes taps already scaled by `up`.  Stands in for the
 * soxr resampling inside librosa.load(path, sr=16000) at reference main.py:76.  x [n_in] f32,
 * y [n_out] f32 with n_out <= ceil(n_in * up / down), taps [n_taps] f32 (odd count), all device. */
int amt_resample_poly_f32(const float* x, int64_t n_in, float* y, int64_t n_out, const float* taps,
                          int n_taps, int up, int down, amt_stream_t stream);

/* ---- building blocks exported for tests and profiling ------------------- */
/* C[M][ldc] (+bias, optional ReLU) = A[M][K] (bf16, row-major) * W[N][K]^T (bf16).
 * tcgen05/TMA kernel; K % 64 == 0, N % 64 == 0.  out_f32 selects f32 or bf16 output. */
int amt_gemm_bf16(const void* A, const void* W, const float* bias, void* C, int M, int N, int K,
                  int ldc, int relu, int out_f32, amt_stream_t stream);
/* Halo-tile implicit-GEMM convolution (reference models/cnn_rnn_model.py:35-38, :83-99, :196-201 with
 * BatchNorm folded) over activations X [B][T][F][Cin] bf16 (Cin = 32 or a multiple of 64):
 * Y[B][T][F or F/2][Cout] = act(conv_{kf x kt}(X) (+ 1x1 conv of X2 [B][T][F][Cin2]) + bias), optional
 * 2:1 max-pool over F.  W [Cout][kf*kt*Cin (+Cin2)] bf16 with K index (kf, kt, cin).
 * Built filters: 3x3 and 7x3 (kf x kt); Cout in {64, 128, 256}; Cin2 in {0, 32, 64k} and <= the
 * main channel block.  pool: bit 0 = 2:1 max-pool over F; bit 1 = split-bf16 output, Y [..][3*Cout] =
 * [hi | lo | hi] per pixel (precise mode). */
int amt_conv_bf16(const void* X, const void* X2, const void* W, const float* bias, void* Y, int B, int T,
                  int F, int Cin, int Cin2, int Cout, int kf, int kt, int relu, int pool,
                  amt_stream_t stream);
/* Precise-mode operand split: x [rows][K] (f32 when in_f32, else bf16) -> out bf16 [rows][3K] =
 * [hi(K) | lo(K) | hi(K)], hi = bf16(x), lo = bf16(x - hi) (0 for bf16 input).  K % 8 == 0. */
int amt_split3_bf16(const void* x, int in_f32, void* out, int64_t rows, int K, amt_stream_t stream);
/* LSTM recurrence (nn.LSTM eval forward, gate order i,f,g,o; reference models/cnn_rnn_model.py:45-52,
 * :212-228) over precomputed input projections gx = x W_ih^T + b_ih + b_hh, for n_seq independent
 * sequences (directions / stacked LSTMs) at once; see DESIGN.md section 4. */
typedef struct amt_lstm_seq {
  const void* whh;      /* bf16 [4H][H], rows in slice order */
  const float* gx;      /* f32 [B*T][ld_gx], this sequence's first column */
  void* out_bf16;       /* bf16 [B*T][ld_out], this sequence's first column (or NULL) */
  float* out_f32;       /* f32  [B*T][ld_out32] (or NULL) */
  int H;
  int reverse;
  int ld_gx, ld_out, ld_out32;
} amt_lstm_seq;
size_t amt_lstm_scratch_bytes(const amt_lstm_seq* seqs, int n_seq, int B);
int amt_lstm_recurrence(const amt_lstm_seq* seqs_host, int n_seq, int B, int T, void* scratch,
                        size_t scratch_bytes, amt_stream_t stream);
/* Clamped softmax attention (reference models/cnn_rnn_model.py:118-139, middle part):
 * qkv bf16 [B*T][3*D] (q|k|v, each [heads][hd]) -> out bf16 [B*T][D].  head_dim 64/128/192 run the
 * tcgen05 kernel, 48/96/144 the mma.sync one. */
int amt_attention_bf16(const void* qkv, void* out, int B, int T, int heads, int head_dim, float clip,
                       amt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AMT_H_ */
