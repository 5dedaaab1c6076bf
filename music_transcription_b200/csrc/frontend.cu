// Fused log-mel frontend (replaces librosa.feature.melspectrogram + power_to_db as
// called at reference main.py:117-125): centred zero-padded framing, periodic Hann,
// real FFT, |X|^2, slaney mel projection, 10*log10, per-chunk (max - top_db) floor.
//
// One CTA = 8 warps handles 8 consecutive frames of one chunk (one frame per warp).  The samples those
// frames share (hop 512 / n_fft 2048 = 4x overlap) are staged once in shared memory
// with coalesced loads; each warp then runs a 1024-point complex FFT of one real
// 2048-sample frame entirely in registers + one shared-memory transpose
// (32 lanes x radix-32 in registers, twiddle, transpose, radix-32), recovers the 1025
// real-FFT bins, and applies the banded (0.6 % dense) mel filterbank as exact fp32
// dot products.  The complex STFT (7.7 MB/chunk) never exists in HBM.  Results are
// staged as a [n_mels][8] tile so the (B, n_mels, T) output is written in 32-byte runs.
#include <cmath>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace amt {

constexpr int kNfft = 2048;
constexpr int kHalf = 1024;
constexpr int kFramesPerCta = 8;
constexpr int kWarpsPerCta = 8;

constexpr int kMaxRounds = 32;   // n_mels <= 1024: rounds of 32 filters (one per lane)
struct FrontendDev {
  const float2* tw1024;    // [1024] exp(-2*pi*i*k/1024)
  const float2* tw2048;    // [1025] exp(-2*pi*i*k/2048)
  const int* fb_start;     // [rounds * 32] first FFT bin of each filter (0 for the padding filters of the last round)
  // banded filterbank in ELL form: round r = filters 32r .. 32r+31, taps[r] = its widest band; the weight of tap i
  // of filter 32r + l sits at fb_w[(off[r] + i) * 32 + l] (zero past the filter's own band), so a warp reads one
  // contiguous 128-byte row per tap
  const float* fb_w;
  int n_mels, hop, rounds, fb_rows;          // fb_rows = sum of taps[]
  unsigned short taps[kMaxRounds], off[kMaxRounds];
};
// padded power spectrum of a frame: bin b at word b + b/32 (1057 words), + 32 zeroed words a zero-weight tap may touch;
// it overwrites the frame's complex spectrum (32 x 33 float2 = 2112 words per warp)
constexpr int kPowWords = 1025 + 1025 / 32 + 32;
static_assert(kPowWords <= 2 * 32 * 33, "the power spectrum must fit the per-warp FFT scratch");

// ---- 32-point in-register FFT (radix-2 DIF, output in bit-reversed order) ----
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __constant__ float2 c_tw32[16];   // exp(-2*pi*i*k/32), k = 0..15

template <int N>
__device__ __forceinline__ void fft_dif_stage(float2 (&v)[32]) {
  // butterflies of span N/2 inside blocks of N
#pragma unroll
  for (int blk = 0; blk < 32; blk += N) {
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
      const float2 a = v[blk + k], b = v[blk + k + N / 2];
      v[blk + k] = make_float2(a.x + b.x, a.y + b.y);
      const float2 d = make_float2(a.x - b.x, a.y - b.y);
      if (k == 0) v[blk + k + N / 2] = d;
      else if (k * (32 / N) == 8) v[blk + k + N / 2] = make_float2(d.y, -d.x);   // * (-i)
      else v[blk + k + N / 2] = cmul(d, c_tw32[k * (32 / N)]);
    }
  }
}
__device__ __forceinline__ void fft32(float2 (&v)[32]) {
  fft_dif_stage<32>(v);
  fft_dif_stage<16>(v);
  fft_dif_stage<8>(v);
  fft_dif_stage<4>(v);
  fft_dif_stage<2>(v);
}
__host__ __device__ constexpr int bitrev5(int x) {
  return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4);
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  // monotone int encoding: works for mixed signs
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void __launch_bounds__(kWarpsPerCta * 32, 2)
logmel_kernel(const float* __restrict__ wav, long long wav_stride, int n_samples, int T, const __grid_constant__ FrontendDev fe,
              float* __restrict__ out, float* __restrict__ chunk_max, int vec_ok) {
  extern __shared__ __align__(16) uint8_t smem_fe[];
  constexpr int kOutPitch = kFramesPerCta + 1;                      // odd pitch: lanes (= filters) hit different banks
  const int span = (kFramesPerCta - 1) * fe.hop + kNfft;            // samples shared by the CTA's frames
  float* s_x = reinterpret_cast<float*>(smem_fe);                   // [span]
  float2* s_z = reinterpret_cast<float2*>(s_x + ((span + 3) & ~3));  // [warps][32*33] transpose / spectrum / power
  float* s_out = reinterpret_cast<float*>(s_z + kWarpsPerCta * 32 * 33);   // [rounds * 32][9]
  int* s_fb_start = reinterpret_cast<int*>(s_out + fe.rounds * 32 * kOutPitch);   // [rounds * 32]
  float* s_fb_w = reinterpret_cast<float*>(s_fb_start + fe.rounds * 32);          // [fb_rows][32]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kFramesPerCta;
  const float* x = wav + static_cast<long long>(b) * wav_stride;

  // stage samples [t0*hop - 1024, ... + span) with zero padding outside [0, n_samples)
  // (all loads of a thread are in flight at once: 16-byte cp.async with zero fill outside the chunk;
  //  a load -> store loop serialises ~44 DRAM round trips per thread)
  const int start = t0 * fe.hop - kHalf;
  if (vec_ok) {
    for (int i = tid * 4; i < span; i += blockDim.x * 4) {
      const int g = start + i;                                 // multiple of 4; n_samples % 4 == 0
      const bool ok = g >= 0 && g < n_samples;
      const int sz = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ptx::smem_u32(s_x + i)), "l"(x + (ok ? g : 0)), "r"(sz)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // (filterbank tables: L2-resident, 16-byte copies issued while the samples are in flight)
  {
    const int n16 = (fe.rounds * 32 + fe.fb_rows * 32) >> 2;   // s_fb_start and s_fb_w are contiguous in both places
    const float4* src = reinterpret_cast<const float4*>(fe.fb_start);
    for (int i = tid; i < n16; i += blockDim.x)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(ptx::smem_u32(reinterpret_cast<float4*>(s_fb_start) + i)), "l"(src + i)
                   : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (!vec_ok) {
    for (int i0 = tid; i0 < span; i0 += blockDim.x * 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int g = start + i0 + u * blockDim.x;
        v[u] = (i0 + u * blockDim.x < span && g >= 0 && g < n_samples) ? __ldg(x + g) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u * blockDim.x < span) s_x[i0 + u * blockDim.x] = v[u];
    }
  }
  __syncthreads();

  float2* zw = s_z + warp * 32 * 33;
  float* pw = reinterpret_cast<float*>(zw);          // the same words, later: the padded power spectrum
  float wmax = -INFINITY;
  // Periodic Hann window without a table: sample n = 64 n1 + 2 lane + j has
  //   w = 0.5 - 0.5 cos(2 pi n1 / 32 + beta_j),  beta_j = 2 pi (2 lane + j) / 2048
  //     = 0.5 - cos(a) * (0.5 cos beta_j) + sin(a) * (0.5 sin beta_j)
  // cos(a), sin(a) are the radix-32 twiddles (constant bank, n1 is a compile-time index); the two beta terms are four
  // registers per lane for the whole kernel.  Two FMAs per sample instead of 32 eight-byte table gathers per frame.
  const float2 tb0 = __ldg(fe.tw2048 + 2 * lane), tb1 = __ldg(fe.tw2048 + 2 * lane + 1);   // (cos beta, -sin beta)
  const float hc0 = 0.5f * tb0.x, hs0 = -0.5f * tb0.y, hc1 = 0.5f * tb1.x, hs1 = -0.5f * tb1.y;

  for (int fi = warp; fi < kFramesPerCta; fi += kWarpsPerCta) {
    const int t = t0 + fi;
    if (t < T) {
      // z[n] = xw[2n] + i xw[2n+1];  n = 32*n1 + n2,  lane = n2, register index = n1
      float2 v[32];
      const float* fx = s_x + fi * fe.hop;
#pragma unroll
      for (int n1 = 0; n1 < 32; ++n1) {
        const int n = 64 * n1 + 2 * lane;
        const float2 xs = *reinterpret_cast<const float2*>(fx + n);
        // (cos a, sin a) for a = 2 pi n1 / 32: c_tw32[k] = (cos, -sin)(2 pi k / 32), and a + pi flips both signs
        const float ca = n1 < 16 ? c_tw32[n1 & 15].x : -c_tw32[n1 & 15].x;
        const float sa = n1 < 16 ? -c_tw32[n1 & 15].y : c_tw32[n1 & 15].y;
        const float w0 = fmaf(-ca, hc0, fmaf(sa, hs0, 0.5f));
        const float w1 = fmaf(-ca, hc1, fmaf(sa, hs1, 0.5f));
        v[n1] = make_float2(xs.x * w0, xs.y * w1);
      }
      fft32(v);                                           // over n1 -> k1 (bit-reversed slots)
      // twiddle W_1024^(n2*k1), then transpose through smem: zw[k1][n2]
      // (powers of w = W_1024^lane by recurrence, re-seeded from the table every 8 steps: 4 table gathers per
      //  lane instead of 32 -- a 32-address gather costs the LSU ~32 cycles -- at < 1e-6 relative error)
      {
        float2 tw = make_float2(1.0f, 0.0f);
        const float2 w1 = __ldg(fe.tw1024 + lane);
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) {
          if ((k1 & 7) == 0 && k1 > 0) tw = __ldg(fe.tw1024 + ((lane * k1) & 1023));
          zw[k1 * 33 + lane] = cmul(v[bitrev5(k1)], tw);
          tw = cmul(tw, w1);
        }
      }
      __syncwarp();
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) v[n2] = zw[lane * 33 + n2];      // lane = k1
      __syncwarp();
      fft32(v);                                           // over n2 -> k2 ; Z[k1 + 32*k2]
#pragma unroll
      for (int sidx = 0; sidx < 32; ++sidx) zw[bitrev5(sidx) * 32 + lane] = v[sidx];   // natural order, k = k1 + 32*k2
      __syncwarp();
      // real-FFT recovery.  With a = Z[k], c = Z[1024-k] (Z[1024] == Z[0]):
      //   E = (a + conj c)/2,  O = (a - conj c)/(2i),  X[k] = E + W_2048^k O,  X[1024-k] = conj(E - W_2048^k O)
      // so one pass over k = 0..512 yields both |X[k]|^2 and |X[1024-k]|^2.  The 17 pairs of a lane stay in the
      // registers the FFT no longer needs until every lane has read its Z values; then the power spectrum
      // overwrites Z as a plain float array, bin b at word b + b/32 (the padding makes the filterbank's strided
      // gathers below conflict-free for every filter spacing).
      float pk[17], qk[17];
#pragma unroll
      for (int i = 0; i < 17; ++i) {
        const int k = 32 * i + lane;
        if (k <= kHalf / 2) {
          const float2 a = zw[k];
          const float2 c = zw[(kHalf - k) & 1023];
          const float2 e = make_float2(0.5f * (a.x + c.x), 0.5f * (a.y - c.y));
          const float2 o = make_float2(0.5f * (a.y + c.y), -0.5f * (a.x - c.x));   // (a - conj(c)) / (2i)
          const float2 w = __ldg(fe.tw2048 + k);
          const float2 wo = make_float2(o.x * w.x - o.y * w.y, o.x * w.y + o.y * w.x);
          const float px = e.x + wo.x, py = e.y + wo.y, qx = e.x - wo.x, qy = e.y - wo.y;
          pk[i] = px * px + py * py;                       // |X[k]|^2
          qk[i] = qx * qx + qy * qy;                       // |X[1024-k]|^2
        }
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 17; ++i) {
        const int k = 32 * i + lane;
        if (k <= kHalf / 2) {
          const int k2 = kHalf - k;
          pw[k + (k >> 5)] = pk[i];
          pw[k2 + (k2 >> 5)] = qk[i];                      // (k = 512 writes the same value twice)
        }
      }
      pw[kPowWords - 32 + lane] = 0.0f;                    // words a zero-weight tap past bin 1024 may read: no NaN * 0
      __syncwarp();
      // banded mel projection + dB: lane l of round r owns filter 32r + l
      const float* wrow = s_fb_w + lane;
      for (int r = 0; r < fe.rounds; ++r) {
        const int m = 32 * r + lane;
        const int st = s_fb_start[m];
        const float* wr = wrow + fe.off[r] * 32;
        const int nt = fe.taps[r];
        float acc = 0.0f;
        for (int i = 0; i < nt; ++i) {
          const int bin = st + i;
          acc = fmaf(wr[i * 32], pw[bin + (bin >> 5)], acc);
        }
        const float db = 10.0f * log10f(fmaxf(acc, 1e-10f));
        s_out[m * kOutPitch + fi] = db;
        if (m < fe.n_mels) wmax = fmaxf(wmax, db);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // coalesced-ish store of the [n_mels][8] tile
  for (int e = tid; e < fe.n_mels * kFramesPerCta; e += blockDim.x) {
    const int m = e / kFramesPerCta, fi = e - m * kFramesPerCta;
    if (t0 + fi < T) out[(static_cast<long long>(b) * fe.n_mels + m) * T + t0 + fi] = s_out[m * kOutPitch + fi];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
  if (lane == 0 && wmax > -INFINITY) atomic_max_float(chunk_max + b, wmax);
}

__global__ void fill_kernel(float* p, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void topdb_floor_kernel(float* __restrict__ x, const float* __restrict__ chunk_max, long long per_chunk,
                                   float top_db, int vec) {
  const int b = blockIdx.y;
  const float floor_v = chunk_max[b] - top_db;
  float* xb = x + b * per_chunk;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (vec) {   // per_chunk % 4 == 0 and 16-byte aligned base
    float4* p = reinterpret_cast<float4*>(xb);
    for (long long i = i0; i < (per_chunk >> 2); i += stride) {
      float4 v = p[i];
      v.x = fmaxf(v.x, floor_v); v.y = fmaxf(v.y, floor_v); v.z = fmaxf(v.z, floor_v); v.w = fmaxf(v.w, floor_v);
      p[i] = v;
    }
  } else {
    for (long long i = i0; i < per_chunk; i += stride) xb[i] = fmaxf(xb[i], floor_v);
  }
}

}  // namespace amt

// ----------------------------------------------------------------------------
// host: filterbank / tables (float64, following SURVEY.md Appendix A)
// ----------------------------------------------------------------------------
struct amt_frontend {
  int sr, n_fft, hop, n_mels;
  std::vector<float> fb_dense;   // n_mels x (1 + n_fft/2)
  amt::FrontendDev dev;
  void* dev_blob;
  int max_band;
};

namespace {

double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

}  // namespace

extern "C" {

int amt_frontend_create(int sr, int n_fft, int hop, int n_mels, double fmin, double fmax, amt_frontend** out) {
  using namespace amt;
  AMT_REQUIRE(out != nullptr, "frontend: out is NULL");
  AMT_REQUIRE(n_fft == kNfft, "frontend: n_fft must be 2048 (librosa default used by the reference), got %d", n_fft);
  AMT_REQUIRE(hop > 0 && hop <= 1024 && hop % 2 == 0, "frontend: hop must be even and in (0, 1024], got %d", hop);
  AMT_REQUIRE(n_mels > 0 && n_mels <= 1024 && sr > 0, "frontend: bad n_mels / sr");
  if (fmax <= 0) fmax = sr / 2.0;
  AMT_TRY(ensure_device());
  const int bins = 1 + n_fft / 2;
  auto* fe = new amt_frontend();
  fe->sr = sr; fe->n_fft = n_fft; fe->hop = hop; fe->n_mels = n_mels;
  fe->fb_dense.assign(static_cast<size_t>(n_mels) * bins, 0.0f);
  // mel points (np.linspace semantics) and triangular slaney-normalised filters
  std::vector<double> mel_f(n_mels + 2);
  const double m0 = hz_to_mel(fmin), m1 = hz_to_mel(fmax);
  for (int i = 0; i < n_mels + 2; ++i) {
    const double step = (m1 - m0) / (n_mels + 1);
    mel_f[i] = mel_to_hz(i == n_mels + 1 ? m1 : m0 + step * i);
  }
  // banded filters -> ELL rounds of 32 filters (see FrontendDev)
  const int rounds = (n_mels + 31) / 32;
  AMT_REQUIRE(rounds <= kMaxRounds, "frontend: n_mels %d exceeds %d", n_mels, 32 * kMaxRounds);
  std::vector<int> start(rounds * 32, 0), count(rounds * 32, 0);
  int max_band = 0;
  for (int i = 0; i < n_mels; ++i) {
    const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
    const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
    int first = -1, last = -1;
    for (int k = 0; k < bins; ++k) {
      const double fk = static_cast<double>(k) * (static_cast<double>(sr) / n_fft);
      const double lower = -(mel_f[i] - fk) / fd0, upper = (mel_f[i + 2] - fk) / fd1;
      const double v = std::max(0.0, std::min(lower, upper)) * enorm;
      const float vf = static_cast<float>(v);
      fe->fb_dense[static_cast<size_t>(i) * bins + k] = vf;
      if (vf > 0.0f) { if (first < 0) first = k; last = k; }
    }
    start[i] = first < 0 ? 0 : first;
    count[i] = first < 0 ? 0 : last - first + 1;
    max_band = std::max(max_band, count[i]);
  }
  int fb_rows = 0;
  for (int r = 0; r < rounds; ++r) {
    int taps = 0;
    for (int l = 0; l < 32; ++l) taps = std::max(taps, count[32 * r + l]);
    fe->dev.taps[r] = static_cast<unsigned short>(taps);
    fe->dev.off[r] = static_cast<unsigned short>(fb_rows);
    fb_rows += taps;
    // a zero-weight tap past a filter's own band still reads the power spectrum: it must stay inside the padded array
    for (int l = 0; l < 32; ++l)
      if (start[32 * r + l] + taps > bins + 31) {
        delete fe;
        return set_error(AMT_ERR_ARG, "frontend: filter %d (round width %d) reaches past the spectrum", 32 * r + l, taps);
      }
  }
  if (fb_rows > 65535) { delete fe; return set_error(AMT_ERR_ARG, "frontend: filterbank too wide"); }
  std::vector<float> w(static_cast<size_t>(fb_rows) * 32, 0.0f);
  for (int r = 0; r < rounds; ++r)
    for (int l = 0; l < 32; ++l) {
      const int m = 32 * r + l;
      for (int k = 0; k < count[m]; ++k)
        w[(static_cast<size_t>(fe->dev.off[r]) + k) * 32 + l] = fe->fb_dense[static_cast<size_t>(m) * bins + start[m] + k];
    }
  fe->max_band = max_band;
  // tables
  std::vector<float2> tw1024(1024), tw2048(1025);
  const double PI = 3.14159265358979323846;
  for (int k = 0; k < 1024; ++k) tw1024[k] = make_float2((float)std::cos(2.0 * PI * k / 1024.0), (float)-std::sin(2.0 * PI * k / 1024.0));
  for (int k = 0; k <= 1024; ++k) tw2048[k] = make_float2((float)std::cos(2.0 * PI * k / 2048.0), (float)-std::sin(2.0 * PI * k / 2048.0));
  float2 tw32[16];
  for (int k = 0; k < 16; ++k) tw32[k] = make_float2((float)std::cos(2.0 * PI * k / 32.0), (float)-std::sin(2.0 * PI * k / 32.0));
  cudaError_t ce = cudaMemcpyToSymbol(c_tw32, tw32, sizeof(tw32));
  if (ce != cudaSuccess) { delete fe; return set_error(AMT_ERR_CUDA, "frontend: constant upload failed: %s", cudaGetErrorString(ce)); }

  // one device blob: [tw1024 | tw2048 (+ pad) | fb_start | fb_w]  (fb_start and fb_w contiguous: one copy loop in the kernel)
  const size_t o_t1 = 0, o_t2 = o_t1 + sizeof(float2) * 1024, o_st = align_up(o_t2 + sizeof(float2) * 1025, 16),
               o_w = o_st + sizeof(int) * rounds * 32, total = o_w + sizeof(float) * w.size() + 16;
  std::vector<uint8_t> blob(total, 0);
  memcpy(blob.data() + o_t1, tw1024.data(), sizeof(float2) * 1024);
  memcpy(blob.data() + o_t2, tw2048.data(), sizeof(float2) * 1025);
  memcpy(blob.data() + o_st, start.data(), sizeof(int) * rounds * 32);
  if (!w.empty()) memcpy(blob.data() + o_w, w.data(), sizeof(float) * w.size());
  void* d = nullptr;
  ce = cudaMalloc(&d, total);
  if (ce == cudaSuccess) ce = cudaMemcpy(d, blob.data(), total, cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) { if (d) cudaFree(d); delete fe; return set_error(AMT_ERR_CUDA, "frontend: table upload failed: %s", cudaGetErrorString(ce)); }
  uint8_t* db = static_cast<uint8_t*>(d);
  fe->dev_blob = d;
  fe->dev.tw1024 = reinterpret_cast<const float2*>(db + o_t1);
  fe->dev.tw2048 = reinterpret_cast<const float2*>(db + o_t2);
  fe->dev.fb_start = reinterpret_cast<const int*>(db + o_st);
  fe->dev.fb_w = reinterpret_cast<const float*>(db + o_w);
  fe->dev.n_mels = n_mels;
  fe->dev.rounds = rounds;
  fe->dev.fb_rows = fb_rows;
  fe->dev.hop = hop;
  *out = fe;
  return 0;
}

int amt_frontend_destroy(amt_frontend* fe) {
  if (!fe) return 0;
  if (fe->dev_blob) cudaFree(fe->dev_blob);
  delete fe;
  return 0;
}

int amt_frontend_num_frames(const amt_frontend* fe, int n_samples) {
  if (!fe || n_samples < 0) return amt::set_error(AMT_ERR_ARG, "frontend: bad arguments");
  return 1 + n_samples / fe->hop;
}

int amt_frontend_filterbank_host(const amt_frontend* fe, float* out_host) {
  if (!fe || !out_host) return amt::set_error(AMT_ERR_ARG, "frontend: bad arguments");
  memcpy(out_host, fe->fb_dense.data(), fe->fb_dense.size() * sizeof(float));
  return 0;
}

int amt_logmel_f32(amt_frontend* fe, const float* wav, int B, int n_samples, int64_t wav_stride, float* out_db,
                   float top_db, float* chunk_max, amt_stream_t stream_) {
  using namespace amt;
  AMT_REQUIRE(fe && wav && out_db && chunk_max, "logmel: NULL argument");
  AMT_REQUIRE(B > 0 && n_samples > 0 && wav_stride >= n_samples, "logmel: bad sizes");
  AMT_TRY(ensure_device());
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int T = 1 + n_samples / fe->hop;
  const int span = (kFramesPerCta - 1) * fe->hop + kNfft;
  const size_t smem = sizeof(float) * ((span + 3) & ~3) + sizeof(float2) * kWarpsPerCta * 32 * 33 +
                      sizeof(float) * fe->dev.rounds * 32 * (kFramesPerCta + 1) + sizeof(int) * fe->dev.rounds * 32 +
                      sizeof(float) * fe->dev.fb_rows * 32;
  AMT_REQUIRE(smem <= 227 * 1024, "logmel: filterbank (n_mels %d) does not fit shared memory", fe->n_mels);
  AMT_FUNC_ATTR(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  fill_kernel<<<ceil_div(B, 256), 256, 0, stream>>>(chunk_max, B, -INFINITY);
  AMT_CHECK_LAUNCH();
  dim3 grid(ceil_div(T, kFramesPerCta), B);
  AMT_REQUIRE(B <= 65535, "logmel: B must be <= 65535");
  // 16-byte staging needs aligned rows and a span that is a whole number of float4
  const int vec_ok = (reinterpret_cast<uintptr_t>(wav) % 16 == 0) && (wav_stride % 4 == 0) && (n_samples % 4 == 0) &&
                     (fe->hop % 4 == 0);
  logmel_kernel<<<grid, kWarpsPerCta * 32, smem, stream>>>(wav, wav_stride, n_samples, T, fe->dev, out_db, chunk_max, vec_ok);
  AMT_CHECK_LAUNCH();
  if (top_db >= 0.0f) {
    const long long per_chunk = static_cast<long long>(fe->n_mels) * T;
    const int vec = (per_chunk % 4 == 0) && (reinterpret_cast<uintptr_t>(out_db) % 16 == 0);
    const long long work = vec ? per_chunk / 4 : per_chunk;
    dim3 g2(static_cast<unsigned>(std::min<long long>((work + 255) / 256, 64)), B);
    topdb_floor_kernel<<<g2, 256, 0, stream>>>(out_db, chunk_max, per_chunk, top_db, vec);
    AMT_CHECK_LAUNCH();
  }
  return 0;
}

}  // extern "C"
