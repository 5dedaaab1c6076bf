// HBM-bound kernels around the GEMMs: the Cin=1 stem convolution, residual+LayerNorm,
// head finalisation (transpose to (B,88,T)) and sigmoid/threshold.
#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"

namespace amt {

// ----------------------------------------------------------------------------
// conv1: Conv2d(1,32,3x3,pad 1) + BatchNorm(eval, folded) + ReLU + MaxPool(2,1)
// (reference models/cnn_rnn_model.py:30-33 and :179-182).  K = 9 is not a tensor-core
// shape: this is an fp32 stencil whose floor is the 64 B/position bf16 write.  Reads logmel
// [B][F][T] f32, writes activations [B][T][F/2][32] bf16 (channels-last, 64-byte rows: the next
// layer's SWIZZLE_64B K block).
//
// Thread = (4 output channels, 1 pooled bin), marching along time with a 4x3 register window:
// per frame 4 shared-memory loads feed 72 FMAs (weights live in registers), and a warp
// (8 channel groups x 4 bins) stores 256 contiguous bytes.
// ----------------------------------------------------------------------------
constexpr int kC1T = 64, kC1F = 32;   // outputs per CTA: 64 frames x 32 pooled bins
constexpr int kC1Rows = 2 * kC1F + 2, kC1Pitch = kC1T + 3;   // 66 input rows, odd pitch (conflict-free)

// SPLIT (precise mode): the pixel's 32 channels are written as a 128-channel group [hi | lo | hi | 0] with
// hi = bf16(x), lo = bf16(x - hi) -- the split-bf16 operand layout of DESIGN.md section 5.
__device__ __forceinline__ uint32_t pack_bf16_lo(float x0, float x1, uint32_t hi) {
  return ptx::pack_bf16(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xffff0000u));
}

// Stage the CTA's input tile (rows 2 fo0 - 1 .., frames t0 - 1 ..; zero outside the spectrogram = the conv padding) with
// ALL of a thread's loads in flight before the first use: a load -> floor -> store loop pays one DRAM round trip per
// element (ncu: 40 % of the kernel's stall samples sat on that dependency), 18 times per thread.
constexpr int kC1Fill = (kC1Rows * (kC1T + 2) + 255) / 256;
template <typename Store>
__device__ __forceinline__ void conv1_fill_tile(const float* __restrict__ xb, int Fin, int T, int fo0, int t0, float floor_v, int tid,
                                                Store store) {
  float v[kC1Fill];
  uint32_t ok = 0;
#pragma unroll
  for (int u = 0; u < kC1Fill; ++u) {
    const int i = tid + u * 256;
    const int rr = i / (kC1T + 2), cc = i - rr * (kC1T + 2);
    const int f = 2 * fo0 - 1 + rr, t = t0 - 1 + cc;
    const bool in = rr < kC1Rows && f >= 0 && f < Fin && t >= 0 && t < T;
    v[u] = in ? __ldg(xb + static_cast<size_t>(f) * T + t) : 0.0f;
    ok |= static_cast<uint32_t>(in) << u;
  }
#pragma unroll
  for (int u = 0; u < kC1Fill; ++u) {
    const int i = tid + u * 256;
    const int rr = i / (kC1T + 2), cc = i - rr * (kC1T + 2);
    if (rr < kC1Rows) store(rr, cc, ((ok >> u) & 1u) ? fmaxf(v[u], floor_v) : 0.0f);
  }
}

template <bool SPLIT>
__global__ void __launch_bounds__(256) conv1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                    const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                    int Fin, int T, int Fout, const float* __restrict__ chunk_max,
                                                    float top_db) {
  constexpr int CS = SPLIT ? 128 : 32;             // channel stride of one output pixel
  __shared__ float tile[kC1Rows][kC1Pitch];
  const int tid = threadIdx.x;
  const int t0 = blockIdx.x * kC1T, fo0 = blockIdx.y * kC1F, b = blockIdx.z;
  const float* xb = x + static_cast<size_t>(b) * Fin * T;
  // power_to_db's per-chunk floor (max - top_db, reference main.py:125) applied on load when the frontend left it to us
  const float floor_v = chunk_max ? __ldg(chunk_max + b) - top_db : -INFINITY;
  conv1_fill_tile(xb, Fin, T, fo0, t0, floor_v, tid, [&](int rr, int cc, float v) { tile[rr][cc] = v; });
  const int lane = tid & 31, warp = tid >> 5;
  const int cg = lane & 7;                         // channels 4*cg .. 4*cg+3
  const int fl = warp * 4 + (lane >> 3);           // pooled bin inside the CTA tile
  float wr[4][9], br[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    br[c] = __ldg(bias + 4 * cg + c);
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[c][k] = __ldg(w + (4 * cg + c) * 9 + k);   // k = kf*3 + kt
  }
  __syncthreads();
  const int fo = fo0 + fl;
  float win[4][3];                                 // input rows 2*fo-1 .. 2*fo+2, frames t-1 .. t+1
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    win[a][1] = tile[2 * fl + a][0];
    win[a][2] = tile[2 * fl + a][1];
  }
  __nv_bfloat16* orow = out + ((static_cast<size_t>(b) * T + t0) * Fout + fo) * CS + 4 * cg;
  const int tmax = min(kC1T, T - t0);
  for (int tl = 0; tl < tmax; ++tl) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      win[a][0] = win[a][1];
      win[a][1] = win[a][2];
      win[a][2] = tile[2 * fl + a][tl + 2];
    }
    float r[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float a0 = br[c], a1 = br[c];
#pragma unroll
      for (int kf = 0; kf < 3; ++kf)
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          a0 = fmaf(wr[c][kf * 3 + kt], win[kf][kt], a0);
          a1 = fmaf(wr[c][kf * 3 + kt], win[kf + 1][kt], a1);
        }
      r[c] = fmaxf(fmaxf(a0, a1), 0.0f);
    }
    if (fo < Fout) {
      const uint32_t h0 = ptx::pack_bf16(r[0], r[1]), h1 = ptx::pack_bf16(r[2], r[3]);
      uint2* o = reinterpret_cast<uint2*>(orow + static_cast<size_t>(tl) * Fout * CS);
      o[0] = make_uint2(h0, h1);
      if constexpr (SPLIT) {
        o[8] = make_uint2(pack_bf16_lo(r[0], r[1], h0), pack_bf16_lo(r[2], r[3], h1));     // + 32 channels
        o[16] = make_uint2(h0, h1);                                                         // + 64
        o[24] = make_uint2(0u, 0u);                                                         // + 96: padding
      }
    }
  }
}

// ----------------------------------------------------------------------------
// conv1 on the tensor pipe (fast mode).  The FFMA stencil above is bound by the fp32 pipe (72 FMAs per thread and
// frame: 0.36-0.42 ms for 60 chunks, 2.3x its 64-byte-per-pixel write floor).  K = 9 is not a tensor-core shape, but
// 27 is: with x = hi + lo and w = wh + wl (bf16 pairs, |residual| < 2^-17 |x|) the stencil becomes a K = 27 (padded
// to 32) contraction  sum_tap (hi wh + lo wh + hi wl)  with fp32 accumulation -- three bf16 products per tap, exact
// to ~2^-16 relative like the split-bf16 operands of the precise mode.  mma.sync m16n8k16 (the warp-level path: a
// 16-frame x 32-channel tile per warp is far below a tcgen05 tile): M = 16 consecutive frames of one frequency row,
// N = 32 channels = 4 n-tiles, K = 2 k-steps.
//   * The input tile is staged ONCE as packed words (hi | lo << 16).  K slot pairs are laid out so that every A
//     register is one such word: thread (g = lane / 4, c = lane % 4) owns taps c and c + 4 --
//       k-step 0: slots (2c, 2c+1)   = tap c   (hi, lo) x (wh, wh);   slots (2c+8, 2c+9)   = tap c+4 (hi, lo) x (wh, wh)
//       k-step 1: slots (16+2c, 17+2c) = (hi[c], hi[c+4]) x (wl[c], wl[c+4])  -- one PRMT of the two words it holds;
//                 slots (24+2c, 25+2c) = tap 8 (hi, lo) x (wh, wh) for c = 0, x (wl, 0) for c = 1, x 0 otherwise
//     so a fragment costs 3 shared-memory loads per frame row (taps c, c+4, 8) for 8 MMAs.
//   * The weight fragments (16 registers) are built once per thread from the fp32 folded weights.
//   * Both pre-pool rows of a pooled bin are accumulated by the same thread (bias in the accumulator init), so
//     ReLU(max(.,.)) is register math; each quad then writes a pixel's 64 channel bytes.
// ----------------------------------------------------------------------------
constexpr int kC1P = 76;     // tile pitch in words: rows 0/1/2 of a fragment's taps fall into disjoint banks
constexpr int kC1StagePitch = 20;   // words per staged output row (16 used): conflict-free fragment writes, 16-byte aligned rows

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t split_word(float x) {      // bf16(x) | bf16(x - bf16(x)) << 16
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  return static_cast<uint32_t>(__bfloat16_as_ushort(h)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l)) << 16);
}

__global__ void __launch_bounds__(256) conv1_mma_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int Fin,
                                                        int T, int Fout, const float* __restrict__ chunk_max, float top_db) {
  __shared__ uint32_t tile[kC1Rows * kC1P];
  __shared__ __align__(16) uint32_t stage[8 * 16 * kC1StagePitch];
  const int tid = threadIdx.x;
  const int t0 = blockIdx.x * kC1T, fo0 = blockIdx.y * kC1F, b = blockIdx.z;
  const float* xb = x + static_cast<size_t>(b) * Fin * T;
  const float floor_v = chunk_max ? __ldg(chunk_max + b) - top_db : -INFINITY;
  conv1_fill_tile(xb, Fin, T, fo0, t0, floor_v, tid, [&](int rr, int cc, float v) { tile[rr * kC1P + cc] = split_word(v); });
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, c = lane & 3;
  // weight fragments: n-tile j covers channels 8j .. 8j+7, this thread's column is channel 8j + g
  uint32_t wb[4][4];                               // [j][k-step 0: b0, b1 | k-step 1: b0, b1]
  float bs[4][2];                                  // bias of this thread's accumulator columns: channels 8j + 2c, + 1
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float* wr = w + (8 * j + g) * 9;
    const uint32_t s0 = split_word(__ldg(wr + c)), s1 = split_word(__ldg(wr + c + 4)), s8 = split_word(__ldg(wr + 8));
    wb[j][0] = __byte_perm(s0, s0, 0x1010);        // (wh[c], wh[c])
    wb[j][1] = __byte_perm(s1, s1, 0x1010);        // (wh[c+4], wh[c+4])
    wb[j][2] = __byte_perm(s0, s1, 0x7632);        // (wl[c], wl[c+4])
    wb[j][3] = c == 0 ? __byte_perm(s8, s8, 0x1010) : (c == 1 ? (s8 >> 16) : 0u);   // (wh[8], wh[8]) | (wl[8], 0) | 0
    bs[j][0] = __ldg(bias + 8 * j + 2 * c);
    bs[j][1] = __ldg(bias + 8 * j + 2 * c + 1);
  }
  __syncthreads();
  const int offA = (c / 3) * kC1P + (c % 3);                       // tap c      = (kf, kt) = (c / 3, c % 3)
  const int offB = ((c + 4) / 3) * kC1P + ((c + 4) % 3);           // tap c + 4
  const int off8 = 2 * kC1P + 2;                                   // tap 8
#pragma unroll 1
  for (int bi = 0; bi < 4; ++bi) {
    const int fl = warp * 4 + bi;                                  // pooled bin inside the CTA tile
    const int fo = fo0 + fl;
    if (fo >= Fout) break;
#pragma unroll 1
    for (int mt = 0; mt < kC1T / 16; ++mt) {
      if (t0 + 16 * mt >= T) break;
      float acc[2][4][4];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[r][j][0] = acc[r][j][2] = bs[j][0];
          acc[r][j][1] = acc[r][j][3] = bs[j][1];
        }
#pragma unroll
      for (int r = 0; r < 2; ++r) {                                // the two pre-pool rows 2 fo, 2 fo + 1
        const uint32_t* p0 = tile + (2 * fl + r) * kC1P + 16 * mt + g;
        const uint32_t a_lo = p0[offA], a_hi = p0[offA + 8];       // tap c:   frames g, g + 8
        const uint32_t b_lo = p0[offB], b_hi = p0[offB + 8];       // tap c+4
        const uint32_t e_lo = p0[off8], e_hi = p0[off8 + 8];       // tap 8
        const uint32_t m_lo = __byte_perm(a_lo, b_lo, 0x5410), m_hi = __byte_perm(a_hi, b_hi, 0x5410);   // (hi[c], hi[c+4])
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mma_bf16_16816(acc[r][j], a_lo, a_hi, b_lo, b_hi, wb[j][0], wb[j][1]);
          mma_bf16_16816(acc[r][j], m_lo, m_hi, e_lo, e_hi, wb[j][2], wb[j][3]);
        }
      }
      // ReLU(max over the pooled pair) -> bf16 pairs, transposed through a warp-private staging tile so that a lane
      // stores 16 contiguous bytes and a quad a pixel's whole 64-byte channel row (the fragment layout gives a lane
      // 4 bytes of every 16: eight half-filled sectors per store instruction, ncu: l1tex 73 % from those)
      uint32_t* stg = stage + warp * (16 * kC1StagePitch);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        stg[g * kC1StagePitch + 4 * j + c] =
            ptx::pack_bf16(fmaxf(fmaxf(acc[0][j][0], acc[1][j][0]), 0.0f), fmaxf(fmaxf(acc[0][j][1], acc[1][j][1]), 0.0f));
        stg[(g + 8) * kC1StagePitch + 4 * j + c] =
            ptx::pack_bf16(fmaxf(fmaxf(acc[0][j][2], acc[1][j][2]), 0.0f), fmaxf(fmaxf(acc[0][j][3], acc[1][j][3]), 0.0f));
      }
      __syncwarp();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = g + 8 * h, tt = t0 + 16 * mt + row;       // lane (g, c): frame row, channels 8c .. 8c + 7
        const uint4 v = *reinterpret_cast<const uint4*>(stg + row * kC1StagePitch + 4 * c);
        if (tt < T) *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(b) * T + tt) * Fout + fo) * 32 + 8 * c) = v;
      }
      __syncwarp();
    }
  }
}

int run_conv1(const float* x, const float* chunk_max, float top_db, const float* w, const float* bias, void* out, int B, int Fin,
              int T, int split, cudaStream_t stream) {
  const int Fout = Fin / 2;
  AMT_REQUIRE(Fout >= 1 && T >= 1 && B >= 1 && B <= 65535, "conv1: bad sizes");
  dim3 grid(ceil_div(T, kC1T), ceil_div(Fout, kC1F), B);
  static const bool force_ffma = getenv("AMT_CONV1_FFMA") != nullptr;       // bring-up switch: the fp32 stencil in fast mode too
  if (split) conv1_kernel<true><<<grid, 256, 0, stream>>>(x, w, bias, static_cast<__nv_bfloat16*>(out), Fin, T, Fout, chunk_max, top_db);
  else if (force_ffma) conv1_kernel<false><<<grid, 256, 0, stream>>>(x, w, bias, static_cast<__nv_bfloat16*>(out), Fin, T, Fout, chunk_max, top_db);
  else conv1_mma_kernel<<<grid, 256, 0, stream>>>(x, w, bias, static_cast<__nv_bfloat16*>(out), Fin, T, Fout, chunk_max, top_db);
  AMT_CHECK_LAUNCH();
  return 0;
}

// ----------------------------------------------------------------------------
// split3: x [rows][K] (f32, or bf16 whose lo part is then zero) -> bf16 [rows][3K] = [hi(K) | lo(K) | hi(K)]:
// the A operand of a precise-mode GEMM (weights packed as [Wh | Wh | Wl]).  HBM-bound, 8 elements per thread.
// ----------------------------------------------------------------------------
template <typename TIn>
__global__ void __launch_bounds__(256) split3_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ out, long long rows,
                                                     int K) {
  const int kv = K >> 3;
  const long long total = rows * kv;
  for (long long e = blockIdx.x * 256ll + threadIdx.x; e < total; e += gridDim.x * 256ll) {
    const long long row = e / kv;
    const int c = static_cast<int>(e - row * kv) << 3;
    uint4 hi, lo;
    if constexpr (sizeof(TIn) == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(x + row * K + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(x + row * K + c) + 1);
      hi = make_uint4(ptx::pack_bf16(a.x, a.y), ptx::pack_bf16(a.z, a.w), ptx::pack_bf16(b.x, b.y), ptx::pack_bf16(b.z, b.w));
      lo = make_uint4(pack_bf16_lo(a.x, a.y, hi.x), pack_bf16_lo(a.z, a.w, hi.y), pack_bf16_lo(b.x, b.y, hi.z),
                      pack_bf16_lo(b.z, b.w, hi.w));
    } else {
      hi = __ldg(reinterpret_cast<const uint4*>(x + row * K + c));
      lo = make_uint4(0u, 0u, 0u, 0u);
    }
    __nv_bfloat16* o = out + row * 3 * K + c;
    *reinterpret_cast<uint4*>(o) = hi;
    *reinterpret_cast<uint4*>(o + K) = lo;
    *reinterpret_cast<uint4*>(o + 2 * K) = hi;
  }
}

int run_split3(const void* x, int in_f32, void* out, long long rows, int K, cudaStream_t stream) {
  AMT_REQUIRE(rows >= 1 && K >= 8 && K % 8 == 0, "split3: K (%d) must be a positive multiple of 8", K);
  const long long total = rows * (K >> 3);
  const unsigned grid = static_cast<unsigned>(std::min<long long>((total + 255) / 256, 16ll * num_sms()));
  if (in_f32) split3_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), static_cast<__nv_bfloat16*>(out), rows, K);
  else split3_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), rows, K);
  AMT_CHECK_LAUNCH();
  return 0;
}

// ----------------------------------------------------------------------------
// y = LayerNorm(a + b) * gamma + beta  (reference models/cnn_rnn_model.py:243,322; eps 1e-6)
// one warp per row, row held in registers, two-pass variance.
// ----------------------------------------------------------------------------
constexpr int kLnMaxVec = 18;   // D <= 2304

// NV = float4 per lane the row needs (D / 128) when it is one of the model's widths (1536 -> 12, 768 -> 6, 384 -> 3:
// hidden 512 / 256 / 128), else the generic bound: sizing the register row to the real width (48 instead of 72 registers
// at D = 1536) lets a third CTA share the SM.
template <bool SPLIT, int NV>
__global__ void __launch_bounds__(256) add_layernorm_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            __nv_bfloat16* __restrict__ out, long long rows, int D, float eps) {
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nvec = D >> 7;                       // float4 per lane
  const float4* pa = reinterpret_cast<const float4*>(a + row * D);
  const float4* pb = reinterpret_cast<const float4*>(b + row * D);
  float4 v[NV];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < nvec) {
      const float4 x = __ldg(pa + i * 32 + lane), y = __ldg(pb + i * 32 + lane);
      v[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  const float mean = sum / D;
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < nvec) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
  const float rstd = rsqrtf(sq / D + eps);
  uint2* po = reinterpret_cast<uint2*>(out + row * (SPLIT ? 3 * D : D));     // SPLIT: row = [hi(D) | lo(D) | hi(D)]
  const float4* pg = reinterpret_cast<const float4*>(gamma);
  const float4* pbt = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < nvec) {
      const float4 g = __ldg(pg + i * 32 + lane), bt = __ldg(pbt + i * 32 + lane);
      const float y0 = (v[i].x - mean) * rstd * g.x + bt.x, y1 = (v[i].y - mean) * rstd * g.y + bt.y;
      const float y2 = (v[i].z - mean) * rstd * g.z + bt.z, y3 = (v[i].w - mean) * rstd * g.w + bt.w;
      const uint32_t h0 = ptx::pack_bf16(y0, y1), h1 = ptx::pack_bf16(y2, y3);
      po[i * 32 + lane] = make_uint2(h0, h1);
      if constexpr (SPLIT) {
        po[(D >> 2) + i * 32 + lane] = make_uint2(pack_bf16_lo(y0, y1, h0), pack_bf16_lo(y2, y3, h1));
        po[(D >> 1) + i * 32 + lane] = make_uint2(h0, h1);
      }
    }
  }
}

int run_add_layernorm(const float* a, const float* b, const float* gamma, const float* beta, void* out, long long rows,
                      int D, float eps, int split, cudaStream_t stream) {
  AMT_REQUIRE(D % 128 == 0 && D <= 128 * kLnMaxVec, "layernorm: D (%d) must be a multiple of 128 and <= %d", D, 128 * kLnMaxVec);
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
#define AMT_LN_LAUNCH(NV_)                                                                                      \
  do {                                                                                                          \
    if (split) add_layernorm_kernel<true, NV_><<<grid, 256, 0, stream>>>(a, b, gamma, beta, o, rows, D, eps);   \
    else add_layernorm_kernel<false, NV_><<<grid, 256, 0, stream>>>(a, b, gamma, beta, o, rows, D, eps);        \
  } while (0)
  switch (D >> 7) {
    case 3: AMT_LN_LAUNCH(3); break;
    case 6: AMT_LN_LAUNCH(6); break;
    case 12: AMT_LN_LAUNCH(12); break;
    default: AMT_LN_LAUNCH(kLnMaxVec); break;
  }
#undef AMT_LN_LAUNCH
  AMT_CHECK_LAUNCH();
  return 0;
}

// ----------------------------------------------------------------------------
// heads: [B*T][ld] f32 (head h occupies columns h*88 .. h*88+87) -> out_h [B][88][T]
// (the .transpose(1, 2) of reference models/cnn_rnn_model.py:74,337-345)
// ----------------------------------------------------------------------------
struct HeadPtrs { float* out[3]; };

__global__ void __launch_bounds__(256) heads_transpose_kernel(const float* __restrict__ in, int ld, int T, int n_out,
                                                              HeadPtrs outs) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;       // c = head*88 + pitch
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    tile[i][tx] = (t < T && c < n_out) ? __ldg(in + (static_cast<size_t>(b) * T + t) * ld + c) : 0.0f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    if (c < n_out && t < T) {
      const int h = c / 88, p = c - h * 88;
      float* o = outs.out[h];
      if (o) o[(static_cast<size_t>(b) * 88 + p) * T + t] = tile[tx][i];
    }
  }
}

int run_heads_transpose(const float* in, int ld, int B, int T, int n_heads, float* o0, float* o1, float* o2,
                        cudaStream_t stream) {
  HeadPtrs hp{{o0, o1, o2}};
  const int n_out = n_heads * 88;
  dim3 grid(ceil_div(T, 32), ceil_div(n_out, 32), B);
  heads_transpose_kernel<<<grid, 256, 0, stream>>>(in, ld, T, n_out, hp);
  AMT_CHECK_LAUNCH();
  return 0;
}

// ----------------------------------------------------------------------------
// probs = sigmoid(logits); roll = probs > thr  (reference main.py:153-156)
// ----------------------------------------------------------------------------
// ONE definition of the probability for every kernel that thresholds, so the float roll, the bit-packed roll
// and the note grouper can never disagree on a cell
__device__ __forceinline__ float sigmoid_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void sigmoid_threshold_kernel(const float* __restrict__ logits, long long n, float thr, float* __restrict__ probs,
                                         float* __restrict__ roll) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float p = sigmoid_ref(logits[i]);
    if (probs) probs[i] = p;
    if (roll) roll[i] = p > thr ? 1.0f : 0.0f;
  }
}

// Bit-packed piano roll: row r (one pitch of one chunk, T frames) -> ceil(T/32) little-endian words, bit t%32 of
// word t/32 = (value > thr).  A warp ballots 32 frames into one word: 88 x 938 floats (330 KB) become 10.6 KB per
// chunk, which is what leaves the device in the streaming path ("bit-packed rolls").
__global__ void __launch_bounds__(256) pack_roll_kernel(const float* __restrict__ vals, long long n_rows, int T, int words,
                                                        float thr, int apply_sigmoid, uint32_t* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  const long long n_words = n_rows * words;
  for (long long w = blockIdx.x * 8ll + (threadIdx.x >> 5); w < n_words; w += gridDim.x * 8ll) {
    const long long row = w / words;
    const int t = static_cast<int>(w - row * words) * 32 + lane;
    bool on = false;
    if (t < T) {
      const float x = __ldg(vals + row * T + t);
      on = (apply_sigmoid ? sigmoid_ref(x) : x) > thr;
    }
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (lane == 0) bits[w] = m;
  }
}

}  // namespace amt

extern "C" int amt_sigmoid_threshold(const float* logits, int64_t n, float thr, float* probs, float* roll,
                                     amt_stream_t stream) {
  using namespace amt;
  AMT_REQUIRE(logits && n >= 0, "sigmoid_threshold: bad arguments");
  AMT_TRY(ensure_device());
  if (n == 0) return 0;
  const long long blocks = std::min<long long>((n + 255) / 256, 8ll * num_sms());
  sigmoid_threshold_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, n, thr, probs, roll);
  AMT_CHECK_LAUNCH();
  return 0;
}

extern "C" int amt_pack_roll_u32(const float* vals, int64_t n_rows, int T, float thr, int apply_sigmoid, uint32_t* bits,
                                 amt_stream_t stream) {
  using namespace amt;
  AMT_REQUIRE(vals && bits && n_rows >= 0 && T >= 1, "pack_roll: bad arguments");
  AMT_TRY(ensure_device());
  if (n_rows == 0) return 0;
  const int words = (T + 31) / 32;
  const long long n_words = n_rows * words;
  const long long blocks = std::min<long long>((n_words + 7) / 8, 16ll * num_sms());
  pack_roll_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(vals, n_rows, T, words, thr,
                                                                                               apply_sigmoid, bits);
  AMT_CHECK_LAUNCH();
  return 0;
}
