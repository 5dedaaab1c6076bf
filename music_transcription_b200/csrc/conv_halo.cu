// Halo-tile implicit-GEMM convolution for sm_100a (tcgen05 / TMA / TMEM, persistent).
//
// Covers the 3x3 and 7x3 convolutions of the CNN (reference models/cnn_rnn_model.py:35-38,
// :83-99, :196-201) over channels-last activations [B][T][F][C] (bf16), with BatchNorm folded
// into the weights, the residual 1x1 skip conv (:88-92) accumulated into the same TMEM tile,
// and bias + ReLU + the 2:1 frequency max-pool fused into the epilogue.
//
// Why a halo tile: a tap-by-tap implicit GEMM (one shifted TMA box per filter tap) re-reads every
// activation kf*kt times from L2 -- 9x (3x3) or 21x (7x3) -- and the 64/128-channel layers of this
// model were L2->SMEM bound (10.7 TB/s of TMA traffic at 230-990 TFLOP/s).  Here the M tile is
// 16 frames x 8 bins; per 64- (or 32-) channel block ONE TMA box {KC, 16 bins, 18 frames} brings the
// tile plus its halo (out-of-range coordinates are zero-filled = the conv padding), and every filter
// tap is just a different START ROW of the same shared-memory tile: the 8-row core-matrix groups of
// the UMMA descriptor are the 16 frame rows (stride = 16 bins x row bytes), and tap (kf, kt) starts
// at row kt*16 + kf.  The 128-byte (or 64-byte) swizzle is a function of the absolute shared-memory
// address, so a start address that is not a multiple of the swizzle atom reads exactly what TMA wrote.
// L2->SMEM activation traffic drops from taps x 16 KB to 36 KB per tile and channel block.
//
// CTA pairs: the layers with 128 / 256 output channels run as clusters of 2 (tcgen05.mma.cta_group::2,
// M = 256 = two adjacent tiles, one per CTA): each CTA loads its own halo tile but only HALF of every weight
// block, which takes a third off the shared-memory traffic these layers are bound by.
// Roles (352 threads): warp 0 = activation (A) producer, warp 1 = weight (B) producer, warp 2 = MMA
// issuer (+ TMEM alloc), warps 3..10 = epilogue.  Two independent smem rings (A: halo tiles, B: one
// [BN x KC] weight block per tap) and a double-buffered TMEM accumulator.
//
// Epilogue: with short K (9 taps x 64 channels) the tile's epilogue, not its MMAs, is the long pole
// (a single warp per scheduler retires ~150 dependent instructions per 32 columns), so it is spread
// over 8 warps (two per TMEM lane quarter), reads its bias from shared memory as float4, and never
// touches global memory itself: each 64-channel chunk of the output tile is staged in a swizzled
// shared-memory buffer (double buffered) and written by ONE TMA tensor store, which also clips the
// rows that fall outside T / F.
#include "kernels.cuh"

namespace amt {

constexpr int kHaloF = 16;      // bins per halo row (tile 8 + up to 8 halo; pitch must be a multiple of 8 rows)
constexpr int kHaloT = 18;      // frames per halo tile (tile 16 + 2)
constexpr int kTileF = 8;
constexpr int kTileT = 16;

struct ConvHaloParams {
  int cblks, cblks2, kc2;       // main / skip channel blocks; channels per skip block (32 or 64)
  int kf, kt, padF, padT;
  int F, T, tilesF, tilesT, num_tiles;
  int kmain;                    // weight columns of the main conv = kf*kt*C
  const float* bias;
  int pool, relu;
  int resident;                 // all weight blocks of a tile fit the B ring: loaded once, never released
};

template <int KC, int BN>
struct ConvHaloCfg {
  static constexpr bool kPair = true;                                  // CTA pairs (cta_group::2)
  static constexpr int kRowBytes = KC * 2;
  static constexpr int kABytes = kHaloF * kHaloT * kRowBytes;          // 36 KB (KC 64) / 18 KB (KC 32)
  static constexpr int kBRows = kPair ? BN / 2 : BN;                   // weight rows this CTA loads per block
  static constexpr int kBBytes = kBRows * kRowBytes;
  // activation tiles in flight: 3 when the layer is tensor-bound (BN >= 128); the 64-output layers are HBM-bound and
  // 3 x 18 KB per SM (8 MB chip-wide) is less than HBM latency x bandwidth, so they get the smem their small weight
  // blocks leave free
  static constexpr int kAStages = BN == 64 ? (KC == 32 ? 6 : 4) : 3;
  static constexpr int kOutBytes = 2 * 128 * 128;                      // two staged [128 rows x 64 ch] output chunks
  static constexpr int kBudget = 225 * 1024 - kAStages * kABytes - kOutBytes - BN * 4 - 1024 - 512;
  static constexpr int kBStagesRaw = kBudget / kBBytes;
  static constexpr int kBStages = kBStagesRaw > 16 ? 16 : kBStagesRaw;   // 16 covers every resident case (<= 10 blocks)
  static constexpr int kSmemBytes = kAStages * kABytes + kBStages * kBBytes + kOutBytes + BN * 4 + 1024 + 512;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  static_assert(kBStages >= 3, "weight ring too small");
};

// smem operand descriptor, K-major, rows of `row_bytes` (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B),
// 8-row groups `sbo_bytes` apart
__device__ __forceinline__ uint64_t conv_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t row_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(row_bytes == 128 ? 2 : 4) << 61;
  return d;
}

constexpr int kConvThreads = 352;
constexpr int kEpiWarp0 = 3;        // first epilogue warp
constexpr int kEpiThreads = 256;

// SPLIT (precise mode): the epilogue writes every output value as a split-bf16 pair, hi = bf16(x) and
// lo = bf16(x - hi), into a [hi(BN) | lo(BN) | hi(BN)] channel group of 3*BN channels per pixel -- the operand layout
// whose weights are packed [Wh | Wh | Wl], so the next layer's MMAs compute hi*Wh + lo*Wh + hi*Wl.
template <int KC, int BN, int KF, bool SPLIT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ CUtensorMap tmOut, const ConvHaloParams p) {
  using Cfg = ConvHaloCfg<KC, BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + Cfg::kAStages * Cfg::kABytes;
  uint8_t* o_smem = b_smem + Cfg::kBStages * Cfg::kBBytes;          // 2 x 16 KB output staging (1024-B aligned)
  float* sbias = reinterpret_cast<float*>(o_smem + Cfg::kOutBytes);
  uint64_t* afull = reinterpret_cast<uint64_t*>(sbias + BN);
  uint64_t* aempty = afull + Cfg::kAStages;
  uint64_t* bfull = aempty + Cfg::kAStages;
  uint64_t* bempty = bfull + Cfg::kBStages;
  uint64_t* tfull = bempty + Cfg::kBStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr bool PAIR = Cfg::kPair;
  // work units: a tile (single CTA) or a pair of consecutive tiles (CTA pair; rank r takes tile 2u + r --
  // a tile index past the end decodes to chunk == B: its loads are zero filled, its stores clipped)
  const int rank = PAIR ? static_cast<int>(ptx::cluster_ctarank()) : 0;
  const int unit0 = PAIR ? blockIdx.x >> 1 : blockIdx.x;
  const int unit_step = PAIR ? gridDim.x >> 1 : gridDim.x;
  const int num_units = PAIR ? (p.num_tiles + 1) >> 1 : p.num_tiles;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA0);
    ptx::prefetch_tmap(&tmA1);
  }
  if (warp == 1 && lane == 0) {
    ptx::prefetch_tmap(&tmB0);
    ptx::prefetch_tmap(&tmB1);
    ptx::prefetch_tmap(&tmOut);
  }
  for (int i = threadIdx.x; i < BN; i += kConvThreads) sbias[i] = p.bias[i];
  if (warp == 2) {
    if (lane == 0) {
      for (int i = 0; i < Cfg::kAStages; ++i) {
        ptx::mbar_init(&afull[i], 1);
        ptx::mbar_init(&aempty[i], 1);
      }
      for (int i = 0; i < Cfg::kBStages; ++i) {
        ptx::mbar_init(&bfull[i], 1);
        ptx::mbar_init(&bempty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&tfull[i], 1);
        ptx::mbar_init(&tempty[i], PAIR ? 16 : 8);
      }
      ptx::mbar_fence_init();
    }
    __syncwarp();
    if constexpr (PAIR) {
      ptx::tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if constexpr (PAIR) ptx::cluster_sync_all();             // the peer's barriers exist before anything targets them
  const uint32_t tmem_base = *tmem_slot;

  constexpr int taps = KF * 3;

  if (warp == 0) {
    // -------------------- activation producer: one halo box per channel block --------------------
    const bool leader = ptx::elect_one_sync();
    const uint32_t afull0 = PAIR ? ptx::mapa(ptx::smem_u32(afull), 0) : ptx::smem_u32(afull);   // (the pair leader's)
    constexpr uint32_t kShare = PAIR ? 2 : 1;              // CTAs whose bytes complete on one barrier
    uint32_t s = 0, ph = 0;
    for (int u = unit0; u < num_units; u += unit_step) {
      int m = PAIR ? 2 * u + rank : u;
      const int f0 = (m % p.tilesF) * kTileF;
      m /= p.tilesF;
      const int t0 = (m % p.tilesT) * kTileT;
      const int b = m / p.tilesT;
      for (int e = 0; e < p.cblks + p.cblks2; ++e) {
        ptx::mbar_wait(&aempty[s], ph ^ 1);
        if (leader) {
          uint8_t* dst = a_smem + s * Cfg::kABytes;
          if (e < p.cblks) {
            if (rank == 0) ptx::mbar_expect_tx(&afull[s], kShare * Cfg::kABytes);
            ptx::tma_load_4d_to<PAIR>(dst, &tmA0, afull0 + 8 * s, e * KC, f0 - p.padF, t0 - p.padT, b);
          } else {
            if (rank == 0) ptx::mbar_expect_tx(&afull[s], kShare * static_cast<uint32_t>(kTileF * kTileT * p.kc2 * 2));
            ptx::tma_load_4d_to<PAIR>(dst, &tmA1, afull0 + 8 * s, (e - p.cblks) * p.kc2, f0, t0, b);
          }
        }
        __syncwarp();
        if (++s == Cfg::kAStages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // -------------------- weight producer: one [BN x KC] block per (channel block, tap) ----------
    const bool leader = ptx::elect_one_sync();
    const uint32_t bfull0 = PAIR ? ptx::mapa(ptx::smem_u32(bfull), 0) : ptx::smem_u32(bfull);
    constexpr uint32_t kShare = PAIR ? 2 : 1;
    const int row0 = rank * Cfg::kBRows;                   // this CTA's half of the output channels
    uint32_t s = 0, ph = 0;
    for (int u = unit0; u < num_units; u += unit_step) {
      if (p.resident && u != unit0) break;                 // weights stay in smem after the first tile
      for (int e = 0; e < p.cblks; ++e) {
        int col = e * KC;                       // weight column of (tap 0, block e); taps are cblks*KC apart
        for (int tap = 0; tap < taps; ++tap) {
          ptx::mbar_wait(&bempty[s], ph ^ 1);
          if (leader) {
            if (rank == 0) ptx::mbar_expect_tx(&bfull[s], kShare * Cfg::kBBytes);
            ptx::tma_load_2d_to<PAIR>(b_smem + s * Cfg::kBBytes, &tmB0, bfull0 + 8 * s, col, row0);
          }
          __syncwarp();
          col += p.cblks * KC;
          if (++s == Cfg::kBStages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
      for (int e = 0; e < p.cblks2; ++e) {
        ptx::mbar_wait(&bempty[s], ph ^ 1);
        if (leader) {
          if (rank == 0) ptx::mbar_expect_tx(&bfull[s], kShare * static_cast<uint32_t>(Cfg::kBRows * p.kc2 * 2));
          ptx::tma_load_2d_to<PAIR>(b_smem + s * Cfg::kBBytes, &tmB1, bfull0 + 8 * s, p.kmain + e * p.kc2, row0);
        }
        __syncwarp();
        if (++s == Cfg::kBStages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 2) {
    // -------------------- MMA issuer --------------------
    // The loop bodies are issue-bound for the small-N layers (a tap is only KC/16 MMAs of 32-64 tensor
    // cycles), so all per-tap state is carried incrementally and the filter loop is unrolled (kt == 3,
    // KF compile-time); with resident weights a tile is one straight-line burst of MMAs.
    const bool leader = ptx::elect_one_sync() && rank == 0;       // in a pair only the leader CTA issues
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(PAIR ? 256 : 128, BN);
    constexpr uint64_t kRow16 = Cfg::kRowBytes >> 4;            // one halo row, in descriptor address units
    constexpr uint64_t kBStage16 = Cfg::kBBytes >> 4;
    const uint32_t a_addr0 = ptx::smem_u32(a_smem), b_addr0 = ptx::smem_u32(b_smem);
    // halo tile: frame rows are kHaloF rows apart; plain (skip) tile and weights: 8-row groups contiguous
    const uint64_t a_halo0 = conv_desc(a_addr0, kHaloF * Cfg::kRowBytes, Cfg::kRowBytes);
    const uint64_t b_main0 = conv_desc(b_addr0, 8 * Cfg::kRowBytes, Cfg::kRowBytes);
    const uint32_t row2 = static_cast<uint32_t>(p.kc2 * 2);
    const uint64_t a_skip0 = conv_desc(a_addr0, 8 * row2, row2);
    const uint64_t b_skip0 = conv_desc(b_addr0, 8 * row2, row2);
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0, tl = 0;
    uint64_t a_stage = 0, b_stage = 0;                          // descriptor offsets of A stage sa / B stage sb
    for (int u = unit0; u < num_units && rank == 0; u += unit_step, ++tl) {
      const uint32_t acc = tl & 1;
      ptx::mbar_wait(&tempty[acc], ((tl >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      uint32_t first = 0;                       // 0 until the tile's first MMA was issued
      if (p.resident) {
        if (tl == 0)                            // weights arrive once; they are never released
          for (int i = 0; i < p.cblks * KF * 3 + p.cblks2; ++i) ptx::mbar_wait(&bfull[i], 0);
        uint64_t b_desc = b_main0;
        for (int e = 0; e < p.cblks; ++e) {
          ptx::mbar_wait(&afull[sa], pa);
          ptx::tc_fence_after();
          if (leader) {
            const uint64_t a_tile = a_halo0 + a_stage;
#pragma unroll
            for (int kfi = 0; kfi < KF; ++kfi)
#pragma unroll
              for (int kti = 0; kti < 3; ++kti)
#pragma unroll
                for (int k = 0; k < KC / 16; ++k)
                  ptx::umma_ss<PAIR>(d_tmem, a_tile + (kti * kHaloF + kfi) * kRow16 + 2 * k,
                                    b_desc + (kfi * 3 + kti) * kBStage16 + 2 * k, idesc, first | kfi | kti | k);
            ptx::umma_commit_to<PAIR>(&aempty[sa]);
          }
          __syncwarp();
          first = 1;
          b_desc += KF * 3 * kBStage16;
          a_stage += Cfg::kABytes >> 4;
          if (++sa == Cfg::kAStages) {
            sa = 0;
            pa ^= 1;
            a_stage = 0;
          }
        }
        const uint64_t b_skip = b_skip0 + (b_desc - b_main0);
        for (int e = 0; e < p.cblks2; ++e) {
          ptx::mbar_wait(&afull[sa], pa);
          ptx::tc_fence_after();
          if (leader) {
            for (int k = 0; k < p.kc2 / 16; ++k)
              ptx::umma_ss<PAIR>(d_tmem, a_skip0 + a_stage + 2 * k, b_skip + e * kBStage16 + 2 * k, idesc, first | k);
            ptx::umma_commit_to<PAIR>(&aempty[sa]);
          }
          __syncwarp();
          first = 1;
          a_stage += Cfg::kABytes >> 4;
          if (++sa == Cfg::kAStages) {
            sa = 0;
            pa ^= 1;
            a_stage = 0;
          }
        }
      } else {
        for (int e = 0; e < p.cblks; ++e) {
          ptx::mbar_wait(&afull[sa], pa);
          uint64_t a_row = a_halo0 + a_stage;   // tap (kf = 0, kt = 0)
#pragma unroll 1
          for (int kfi = 0; kfi < KF; ++kfi) {
#pragma unroll
            for (int kti = 0; kti < 3; ++kti) {   // weight K order is (kf, kt, c): kt fastest
              ptx::mbar_wait(&bfull[sb], pb);
              ptx::tc_fence_after();
              if (leader) {
                const uint64_t a_desc = a_row + kti * kHaloF * kRow16;
                const uint64_t b_desc = b_main0 + b_stage;
#pragma unroll
                for (int k = 0; k < KC / 16; ++k)
                  ptx::umma_ss<PAIR>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, first | kti | k);
                ptx::umma_commit_to<PAIR>(&bempty[sb]);
              }
              __syncwarp();
              b_stage += kBStage16;
              if (++sb == Cfg::kBStages) {
                sb = 0;
                pb ^= 1;
                b_stage = 0;
              }
            }
            first = 1;
            a_row += kRow16;
          }
          if (leader) ptx::umma_commit_to<PAIR>(&aempty[sa]);
          __syncwarp();
          a_stage += Cfg::kABytes >> 4;
          if (++sa == Cfg::kAStages) {
            sa = 0;
            pa ^= 1;
            a_stage = 0;
          }
        }
        for (int e = 0; e < p.cblks2; ++e) {    // residual 1x1 skip conv: plain tile, centre tap only
          ptx::mbar_wait(&afull[sa], pa);
          ptx::mbar_wait(&bfull[sb], pb);
          ptx::tc_fence_after();
          if (leader) {
            for (int k = 0; k < p.kc2 / 16; ++k)
              ptx::umma_ss<PAIR>(d_tmem, a_skip0 + a_stage + 2 * k, b_skip0 + b_stage + 2 * k, idesc, first | k);
            ptx::umma_commit_to<PAIR>(&bempty[sb]);
            ptx::umma_commit_to<PAIR>(&aempty[sa]);
          }
          __syncwarp();
          first = 1;
          b_stage += kBStage16;
          if (++sb == Cfg::kBStages) {
            sb = 0;
            pb ^= 1;
            b_stage = 0;
          }
          a_stage += Cfg::kABytes >> 4;
          if (++sa == Cfg::kAStages) {
            sa = 0;
            pa ^= 1;
            a_stage = 0;
          }
        }
      }
      if (leader) ptx::umma_commit_to<PAIR>(&tfull[acc]);
      __syncwarp();
    }
  } else {
    // -------------------- epilogue: TMEM -> bias / ReLU / freq max-pool -> smem -> TMA store -----
    const int q = warp & 3;                        // TMEM lane quarter this warp may read
    const int half = (warp - kEpiWarp0) >> 2;      // which 32 columns of each 64-column chunk
    const bool issuer = threadIdx.x == kEpiWarp0 * 32;
    const int r = q * 32 + lane;                   // tile row: frame r/8, bin r%8
    const int fl = r & (kTileF - 1);
    // staged row: pooled tiles keep the even bins only (row = frame*4 + bin/2)
    const int ro = p.pool ? ((r >> 3) * 4 + (fl >> 1)) : r;
    const bool writer = !p.pool || (fl & 1) == 0;
    const uint32_t o_row = static_cast<uint32_t>(ro) * 128u;
    const uint32_t tempty0 = PAIR ? ptx::mapa(ptx::smem_u32(tempty), 0) : ptx::smem_u32(tempty);
    uint32_t tl = 0, chunk_no = 0;
    for (int u = unit0; u < num_units; u += unit_step, ++tl) {
      int m = PAIR ? 2 * u + rank : u;
      const int f0 = (m % p.tilesF) * kTileF;
      m /= p.tilesF;
      const int t0 = (m % p.tilesT) * kTileT;
      const int b = m / p.tilesT;
      const uint32_t acc = tl & 1;
      ptx::mbar_wait(&tfull[acc], (tl >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + half * 32;
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c, ++chunk_no) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(taddr + c * 64, v);
        ptx::tmem_ld_wait();
        if (c == BN / 64 - 1) {                    // accumulator fully read: hand it back to the MMA warp
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) ptx::mbar_arrive_remote_relaxed(tempty0 + 8 * acc);     // the leader's barrier
            else ptx::mbar_arrive(&tempty[acc]);
          }
        }
        const float4* b4 = reinterpret_cast<const float4*>(sbias + c * 64 + half * 32);
        uint32_t pk[16], pl[SPLIT ? 16 : 1];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = b4[j];
          float x0 = __uint_as_float(v[4 * j]) + bb.x, x1 = __uint_as_float(v[4 * j + 1]) + bb.y;
          float x2 = __uint_as_float(v[4 * j + 2]) + bb.z, x3 = __uint_as_float(v[4 * j + 3]) + bb.w;
          if (p.relu) {
            x0 = fmaxf(x0, 0.0f); x1 = fmaxf(x1, 0.0f); x2 = fmaxf(x2, 0.0f); x3 = fmaxf(x3, 0.0f);
          }
          if (p.pool) {
            x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 1));
            x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 1));
            x2 = fmaxf(x2, __shfl_xor_sync(0xffffffffu, x2, 1));
            x3 = fmaxf(x3, __shfl_xor_sync(0xffffffffu, x3, 1));
          }
          pk[2 * j] = ptx::pack_bf16(x0, x1);
          pk[2 * j + 1] = ptx::pack_bf16(x2, x3);
          if constexpr (SPLIT) {
            pl[2 * j] = ptx::pack_bf16(x0 - __uint_as_float(pk[2 * j] << 16), x1 - __uint_as_float(pk[2 * j] & 0xffff0000u));
            pl[2 * j + 1] = ptx::pack_bf16(x2 - __uint_as_float(pk[2 * j + 1] << 16), x3 - __uint_as_float(pk[2 * j + 1] & 0xffff0000u));
          }
        }
        if constexpr (SPLIT) {
          // both staging buffers are used per chunk (hi, lo): wait until the previous chunk's stores have read them
          if (issuer) ptx::bulk_wait_group_read0();
          ptx::named_bar_sync(1, kEpiThreads);
          if (writer) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t off = o_row + ((static_cast<uint32_t>(half * 4 + j) ^ (ro & 7)) << 4);
              *reinterpret_cast<uint4*>(o_smem + off) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
              *reinterpret_cast<uint4*>(o_smem + 128 * 128 + off) = make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
            }
          }
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(1, kEpiThreads);
          if (issuer) {
            const int fo = p.pool ? (f0 >> 1) : f0;
            ptx::tma_store_4d(&tmOut, o_smem, c * 64, fo, t0, b);
            ptx::tma_store_4d(&tmOut, o_smem + 128 * 128, BN + c * 64, fo, t0, b);
            ptx::tma_store_4d(&tmOut, o_smem, 2 * BN + c * 64, fo, t0, b);
            ptx::bulk_commit_group();
          }
        } else {
          uint8_t* obuf = o_smem + (chunk_no & 1) * (128 * 128);
          if (writer) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t c16 = static_cast<uint32_t>(half * 4 + j);
              *reinterpret_cast<uint4*>(obuf + o_row + ((c16 ^ (ro & 7)) << 4)) =
                  make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            }
          }
          ptx::fence_proxy_async_smem();
          // the buffer written NEXT (other parity) was last read by the store issued one chunk ago
          if (issuer) ptx::bulk_wait_group_read0();
          ptx::named_bar_sync(1, kEpiThreads);
          if (issuer) {
            ptx::tma_store_4d(&tmOut, obuf, c * 64, p.pool ? (f0 >> 1) : f0, t0, b);
            ptx::bulk_commit_group();
          }
        }
      }
    }
    if (issuer) ptx::bulk_wait_group0();           // all output stores complete before the CTA exits
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) ptx::cluster_sync_all();             // the leader's MMAs also wrote the peer's tensor memory
  if (warp == 2) {
    __syncwarp();
    if constexpr (PAIR) ptx::tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int KC, int BN, int KF, bool SPLIT>
static int launch_conv_halo(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0, const CUtensorMap& b1,
                            const CUtensorMap& o, const ConvHaloParams& p, cudaStream_t stream) {
  using Cfg = ConvHaloCfg<KC, BN>;
  AMT_FUNC_ATTR((conv_halo_kernel<KC, BN, KF, SPLIT>), cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  ConvHaloParams q = p;
  q.resident = (p.cblks * KF * 3 + p.cblks2 <= Cfg::kBStages) ? 1 : 0;
  if constexpr (Cfg::kPair) {
    const int units = (p.num_tiles + 1) / 2;
    const int pairs = units < num_sms() / 2 ? units : num_sms() / 2;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kConvThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    AMT_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<KC, BN, KF, SPLIT>, a0, a1, b0, b1, o, q));
    count_launch();
  } else {
    const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    conv_halo_kernel<KC, BN, KF, SPLIT><<<grid, kConvThreads, Cfg::kSmemBytes, stream>>>(a0, a1, b0, b1, o, q);
    AMT_CHECK_LAUNCH();
  }
  return 0;
}

static CUtensorMapSwizzle swizzle_for(int kc) { return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B; }

int run_conv_halo(const void* X, int C, const void* X2, int C2, int B, int T, int F, const void* W, const float* bias,
                  int N, int kf, int kt, void* out, int relu, int pool, int split, cudaStream_t stream) {
  AMT_TRY(ensure_device());
  AMT_REQUIRE(B > 0 && T > 0 && F > 0, "conv: empty problem");
  AMT_REQUIRE(C == 32 || C % 64 == 0, "conv: Cin (%d) must be 32 or a multiple of 64", C);
  AMT_REQUIRE(X2 == nullptr || C2 == 32 || C2 % 64 == 0, "conv: skip Cin (%d) must be 32 or a multiple of 64", C2);
  AMT_REQUIRE(N == 64 || N == 128 || N == 256, "conv: Cout (%d) must be 64, 128 or 256", N);
  AMT_REQUIRE((kf == 3 || kf == 7) && kt == 3, "conv: filter %dx%d unsupported (3x3 and 7x3 are built)", kf, kt);
  AMT_REQUIRE(kf == 3 || C % 64 == 0, "conv: the 7x3 filter needs Cin %% 64 == 0");
  const int KC = C == 32 ? 32 : 64;
  const int c2 = X2 ? C2 : 0;
  const int kc2 = c2 == 0 ? KC : (c2 == 32 ? 32 : 64);
  AMT_REQUIRE(kc2 <= KC, "conv: skip channel block wider than the main one");
  const long long Ktot = static_cast<long long>(kf) * kt * C + c2;

  CUtensorMap a0, a1, b0, b1, om;
  {
    const int Fout = pool ? F / 2 : F;
    AMT_REQUIRE(Fout >= 1, "conv: pooled output is empty");
    const uint64_t No = split ? 3ull * N : N;          // split: [hi(N) | lo(N) | hi(N)] per pixel
    uint64_t dims[4] = {No, (uint64_t)Fout, (uint64_t)T, (uint64_t)B};
    uint64_t str[3] = {No * 2, (uint64_t)Fout * No * 2, (uint64_t)T * Fout * No * 2};
    uint32_t box[4] = {64, (uint32_t)(pool ? kTileF / 2 : kTileF), kTileT, 1};
    AMT_TRY(encode_tmap_bf16(&om, out, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)F, (uint64_t)T, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)F * C * 2, (uint64_t)T * F * C * 2};
    uint32_t box[4] = {(uint32_t)KC, kHaloF, kHaloT, 1};
    AMT_TRY(encode_tmap_bf16(&a0, X, 4, dims, str, box, swizzle_for(KC)));
  }
  if (X2) {
    uint64_t dims[4] = {(uint64_t)c2, (uint64_t)F, (uint64_t)T, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)c2 * 2, (uint64_t)F * c2 * 2, (uint64_t)T * F * c2 * 2};
    uint32_t box[4] = {(uint32_t)kc2, kTileF, kTileT, 1};
    AMT_TRY(encode_tmap_bf16(&a1, X2, 4, dims, str, box, swizzle_for(kc2)));
  } else {
    a1 = a0;
  }
  {
    uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)Ktot * 2};
    const uint32_t brows = N / 2;                      // each CTA of a pair loads half of a weight block
    uint32_t box[2] = {(uint32_t)KC, brows};
    AMT_TRY(encode_tmap_bf16(&b0, W, 2, dims, str, box, swizzle_for(KC)));
    uint32_t box2[2] = {(uint32_t)kc2, brows};
    AMT_TRY(encode_tmap_bf16(&b1, W, 2, dims, str, box2, swizzle_for(kc2)));
  }

  ConvHaloParams p;
  p.cblks = C / KC;
  p.cblks2 = c2 / kc2;
  p.kc2 = kc2;
  p.kf = kf;
  p.kt = kt;
  p.padF = kf / 2;
  p.padT = kt / 2;
  p.F = F;
  p.T = T;
  p.tilesF = ceil_div(F, kTileF);
  p.tilesT = ceil_div(T, kTileT);
  const long long nt = static_cast<long long>(B) * p.tilesF * p.tilesT;
  AMT_REQUIRE(nt < (1ll << 31), "conv: too many tiles");
  p.num_tiles = static_cast<int>(nt);
  p.kmain = kf * kt * C;
  p.bias = bias;
  p.pool = pool;
  p.relu = relu;
  p.resident = 0;

#define AMT_CONV_DISPATCH(KC_, N_, KF_)                                                                   \
  return split ? launch_conv_halo<KC_, N_, KF_, true>(a0, a1, b0, b1, om, p, stream)                    \
               : launch_conv_halo<KC_, N_, KF_, false>(a0, a1, b0, b1, om, p, stream)
  if (kf == 7) {
    if (N == 64) AMT_CONV_DISPATCH(64, 64, 7);
    if (N == 128) AMT_CONV_DISPATCH(64, 128, 7);
    AMT_CONV_DISPATCH(64, 256, 7);
  }
  if (KC == 32) {
    if (N == 64) AMT_CONV_DISPATCH(32, 64, 3);
    if (N == 128) AMT_CONV_DISPATCH(32, 128, 3);
    AMT_CONV_DISPATCH(32, 256, 3);
  }
  if (N == 64) AMT_CONV_DISPATCH(64, 64, 3);
  if (N == 128) AMT_CONV_DISPATCH(64, 128, 3);
  AMT_CONV_DISPATCH(64, 256, 3);
#undef AMT_CONV_DISPATCH
}

}  // namespace amt
