// tcgen05 / TMA / TMEM persistent implicit-GEMM for sm_100a.
//
// One kernel serves every dense contraction on the path:
//   * the 3x3 / 7x3 convolutions of the CNN (reference models/cnn_rnn_model.py:35-38,
//     :83-99, :196-201) as implicit GEMM over NHWC-like activations [B][T][F][C]:
//     for every filter tap the A tile is ONE shifted 4-D TMA box (zero fill outside
//     the tensor = the conv's zero padding), so no im2col buffer ever exists;
//   * the residual 1x1 skip conv (:88-92) as extra K blocks from a second tensor map,
//     accumulated into the same TMEM tile (BatchNorm is folded into the weights);
//   * all nn.Linear / LSTM input projections (:45-55, :212-260) as the 1-tap case.
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
// warps 2..5 = epilogue (tcgen05.ld -> bias/ReLU/freq-max-pool -> global).  The
// accumulator is double buffered in TMEM so the epilogue of tile i overlaps the
// MMAs of tile i+1.  Tile = 128 (M) x BN (N), BLOCK_K = 64 bf16 = one 128-B swizzle atom.
#include <cstdlib>

#include "kernels.cuh"

namespace amt {

struct GemmParams {
  int kblocks0, cblk0, ntapT, padF, padT;
  int kblocks1;
  int boxF_log2, boxT;
  int F, T, Bn;
  int tilesF, tilesT, n_tiles, num_tiles;
  int m_tiles, group_n;   // rasterisation: n-tiles are swept in groups of group_n (weight slab stays in L2)
  const float* bias;
  void* out;
  long long ld_out;
  int Fout;
  int pool, relu;
};

// tile id -> (m tile, n tile).  n-tiles are visited in groups of group_n: inside a group the order is
// m-major with n fastest, so the ~148 concurrently running tiles share group_n weight tiles (the
// slab stays L2-resident while the activations stream past once per group).
__device__ __forceinline__ int decode_tile(const GemmParams& p, int tile, int& m) {
  const int per_group = p.m_tiles * p.group_n;
  const int ng = tile / per_group;
  const int rem = tile - ng * per_group;
  const int gsize = min(p.group_n, p.n_tiles - ng * p.group_n);
  m = rem / gsize;
  return ng * p.group_n + (rem - m * gsize);
}

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB

template <int BN>
struct GemmCfg {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, bool OUT_F32>
__global__ void __launch_bounds__(192, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* tfull = empty + Cfg::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA0);
    ptx::prefetch_tmap(&tmA1);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < Cfg::kStages; ++i) {
        ptx::mbar_init(&full[i], 1);
        ptx::mbar_init(&empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&tfull[i], 1);
        ptx::mbar_init(&tempty[i], 4);
      }
      ptx::mbar_fence_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kblocks = p.kblocks0 + p.kblocks1;
  const int boxF = 1 << p.boxF_log2;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    // The whole warp runs the (warp-uniform) loop; only the issuing instructions are predicated on
    // one elected lane, so coordinates / addresses stay in uniform registers.
    const bool leader = ptx::elect_one_sync();
    // all per-k-block state is carried incrementally (no div/mod in the loop: the producer must stay
    // well under the 128-cycle MMA time of a BN=64 k-block)
    uint32_t s = 0, ph = 0;
    uint8_t* a_dst = smem;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int m;
      const int n0 = decode_tile(p, tile, m) * BN;
      const int f0 = (m % p.tilesF) * boxF;
      m /= p.tilesF;
      const int t0 = (m % p.tilesT) * p.boxT;
      const int b = m / p.tilesT;
      int cb = 0, kt = 0;
      int c1 = f0 - p.padF;                   // f coordinate of the current tap row
      for (int kb = 0; kb < kblocks; ++kb) {
        ptx::mbar_wait(&empty[s], ph ^ 1);
        const bool first = kb < p.kblocks0;
        const CUtensorMap* am = first ? &tmA0 : &tmA1;
        const int c0 = first ? cb * kBlockK : (kb - p.kblocks0) * kBlockK;
        const int cf = first ? c1 : f0;
        const int ct = first ? t0 + kt - p.padT : t0;
        if (leader) {
          ptx::mbar_expect_tx(&full[s], Cfg::kStageBytes);
          ptx::tma_load_4d(a_dst, am, &full[s], c0, cf, ct, b);
          ptx::tma_load_2d(a_dst + kABytes, &tmB, &full[s], kb * kBlockK, n0);
        }
        __syncwarp();
        if (++cb == p.cblk0) {                // next tap: (kf, kt) row-major
          cb = 0;
          if (++kt == p.ntapT) {
            kt = 0;
            ++c1;
          }
        }
        a_dst += Cfg::kStageBytes;
        if (++s == Cfg::kStages) {
          s = 0;
          ph ^= 1;
          a_dst = smem;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    const bool leader = ptx::elect_one_sync();
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(kBlockM, BN);
    const uint64_t desc0 = ptx::umma_desc_sw128(ptx::smem_u32(smem));     // stage 0, A tile, k = 0
    uint32_t s = 0, ph = 0, tl = 0;
    uint64_t a_desc = desc0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
      const uint32_t acc = tl & 1;
      const uint32_t aph = (tl >> 1) & 1;
      ptx::mbar_wait(&tempty[acc], aph ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < kblocks; ++kb) {
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after();
        // descriptors differ only in the 14-bit start-address field (units of 16 bytes)
        const uint64_t b_desc = a_desc + static_cast<uint64_t>(kABytes >> 4);
        if (leader) {
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            ptx::umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit(&empty[s]);   // frees the smem slot once these MMAs retire
        }
        __syncwarp();
        a_desc += static_cast<uint64_t>(Cfg::kStageBytes >> 4);
        if (++s == Cfg::kStages) {
          s = 0;
          ph ^= 1;
          a_desc = desc0;
        }
      }
      if (leader) ptx::umma_commit(&tfull[acc]);   // accumulator ready for the epilogue
      __syncwarp();
    }
  } else {
    // ------------------------------ epilogue ----------------------------------
    const int q = warp & 3;              // TMEM lane quarter this warp may read
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
      int m;
      const int n0 = decode_tile(p, tile, m) * BN;
      const int f0 = (m % p.tilesF) * boxF;
      m /= p.tilesF;
      const int t0 = (m % p.tilesT) * p.boxT;
      const int b = m / p.tilesT;
      const uint32_t acc = tl & 1;
      const uint32_t aph = (tl >> 1) & 1;

      const int r = q * 32 + lane;
      const int fl = r & (boxF - 1);
      const int f = f0 + fl;
      const int t = t0 + (r >> p.boxF_log2);
      bool ok = (f < p.F) && (t < p.T);
      long long orow;
      if (p.pool) {
        ok = ok && ((fl & 1) == 0) && (f + 1 < p.F);
        orow = (static_cast<long long>(b) * p.T + t) * p.Fout + (f >> 1);
      } else {
        orow = (static_cast<long long>(b) * p.T + t) * p.Fout + f;
      }

      ptx::mbar_wait(&tfull[acc], aph);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(taddr + c * 32, v);
        ptx::tmem_ld_wait();
        float x[32];
        const float* bias = p.bias + n0 + c * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          x[j] = __uint_as_float(v[j]) + __ldg(bias + j);
          if (p.relu) x[j] = fmaxf(x[j], 0.0f);
        }
        if (p.pool) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], __shfl_xor_sync(0xffffffffu, x[j], 1));
        }
        if (ok) {
          if (OUT_F32) {
            float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow * p.ld_out + n0 + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
          } else {
            uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + orow * p.ld_out + n0 + c * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dst[j] = make_uint4(ptx::pack_bf16(x[8 * j], x[8 * j + 1]), ptx::pack_bf16(x[8 * j + 2], x[8 * j + 3]),
                                  ptx::pack_bf16(x[8 * j + 4], x[8 * j + 5]), ptx::pack_bf16(x[8 * j + 6], x[8 * j + 7]));
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ----------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------
template <int BN, bool OUT_F32>
static int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const GemmParams& p,
                  cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    AMT_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, OUT_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::kSmemBytes));
    attr_set = true;
  }
  int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  tc_gemm_kernel<BN, OUT_F32><<<grid, 192, Cfg::kSmemBytes, stream>>>(a0, a1, b, p);
  AMT_CHECK_LAUNCH();
  return 0;
}

int run_conv_gemm(const ConvGemmDesc& d, cudaStream_t stream) {
  AMT_TRY(ensure_device());
  AMT_REQUIRE(d.C % 64 == 0 && (d.X2 == nullptr || d.C2 % 64 == 0), "conv/gemm: channel counts must be multiples of 64");
  AMT_REQUIRE(d.N % 64 == 0, "conv/gemm: N (%d) must be a multiple of 64", d.N);
  AMT_REQUIRE(d.boxF * d.boxT == kBlockM && (d.boxF & (d.boxF - 1)) == 0, "conv/gemm: bad M-tile box");
  AMT_REQUIRE(d.B > 0 && d.T > 0 && d.F > 0, "conv/gemm: empty problem");
  const int BN = d.N % 256 == 0 ? 256 : (d.N % 128 == 0 ? 128 : 64);
  const int c2 = d.X2 ? d.C2 : 0;
  const long long Ktot = static_cast<long long>(d.kf) * d.kt * d.C + c2;

  CUtensorMap a0, a1, bm;
  {
    uint64_t dims[4] = {(uint64_t)d.C, (uint64_t)d.F, (uint64_t)d.T, (uint64_t)d.B};
    uint64_t str[3] = {(uint64_t)d.C * 2, (uint64_t)d.F * d.C * 2, (uint64_t)d.T * d.F * d.C * 2};
    uint32_t box[4] = {64, (uint32_t)d.boxF, (uint32_t)d.boxT, 1};
    AMT_TRY(encode_tmap_bf16(&a0, d.X, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  if (d.X2) {
    uint64_t dims[4] = {(uint64_t)d.C2, (uint64_t)d.F, (uint64_t)d.T, (uint64_t)d.B};
    uint64_t str[3] = {(uint64_t)d.C2 * 2, (uint64_t)d.F * d.C2 * 2, (uint64_t)d.T * d.F * d.C2 * 2};
    uint32_t box[4] = {64, (uint32_t)d.boxF, (uint32_t)d.boxT, 1};
    AMT_TRY(encode_tmap_bf16(&a1, d.X2, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  } else {
    a1 = a0;
  }
  {
    uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)d.N};
    uint64_t str[1] = {(uint64_t)Ktot * 2};
    uint32_t box[2] = {64, (uint32_t)BN};
    AMT_TRY(encode_tmap_bf16(&bm, d.W, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }

  GemmParams p;
  p.cblk0 = d.C / 64;
  p.kblocks0 = d.kf * d.kt * p.cblk0;
  p.ntapT = d.kt;
  p.padF = d.kf / 2;
  p.padT = d.kt / 2;
  p.kblocks1 = c2 / 64;
  p.boxF_log2 = 0;
  while ((1 << p.boxF_log2) < d.boxF) ++p.boxF_log2;
  p.boxT = d.boxT;
  p.F = d.F;
  p.T = d.T;
  p.Bn = d.B;
  p.tilesF = ceil_div(d.F, d.boxF);
  p.tilesT = ceil_div(d.T, d.boxT);
  p.n_tiles = d.N / BN;
  p.m_tiles = d.B * p.tilesF * p.tilesT;
  p.group_n = p.n_tiles > 8 ? 8 : p.n_tiles;
  const long long nt = static_cast<long long>(d.B) * p.tilesF * p.tilesT * p.n_tiles;
  AMT_REQUIRE(nt < (1ll << 31), "conv/gemm: too many tiles");
  p.num_tiles = static_cast<int>(nt);
  p.bias = d.bias;
  p.out = d.out;
  p.ld_out = d.ld_out;
  p.pool = d.pool;
  p.relu = d.relu;
  p.Fout = d.pool ? d.F / 2 : d.F;

  if (d.out_f32) {
    if (BN == 256) return launch<256, true>(a0, a1, bm, p, stream);
    if (BN == 128) return launch<128, true>(a0, a1, bm, p, stream);
    return launch<64, true>(a0, a1, bm, p, stream);
  }
  if (BN == 256) return launch<256, false>(a0, a1, bm, p, stream);
  if (BN == 128) return launch<128, false>(a0, a1, bm, p, stream);
  return launch<64, false>(a0, a1, bm, p, stream);
}

int run_gemm(const void* A, const void* W, const float* bias, void* C, int M, int N, int K, long long ldc, int relu,
             int out_f32, cudaStream_t stream) {
  AMT_REQUIRE(K % 64 == 0, "gemm: K (%d) must be a multiple of 64", K);
  ConvGemmDesc d{};
  d.X = A; d.C = K; d.X2 = nullptr; d.C2 = 0;
  d.B = 1; d.T = 1; d.F = M;
  d.W = W; d.bias = bias; d.N = N;
  d.kf = 1; d.kt = 1;
  d.out = C; d.ld_out = ldc;
  d.relu = relu; d.pool = 0; d.out_f32 = out_f32;
  d.boxF = 128; d.boxT = 1;
  return run_conv_gemm(d, stream);
}

}  // namespace amt

extern "C" {

int amt_gemm_bf16(const void* A, const void* W, const float* bias, void* C, int M, int N, int K, int ldc, int relu,
                  int out_f32, amt_stream_t stream) {
  return amt::run_gemm(A, W, bias, C, M, N, K, ldc, relu, out_f32, static_cast<cudaStream_t>(stream));
}

int amt_conv_bf16(const void* X, const void* X2, const void* W, const float* bias, void* Y, int B, int T, int F,
                  int Cin, int Cin2, int Cout, int kf, int kt, int relu, int pool, amt_stream_t stream) {
  static const bool taps = getenv("AMT_CONV_TAPS") != nullptr;     // bring-up switch: tap-by-tap TMA boxes
  if (!taps)
    return amt::run_conv_halo(X, Cin, X2, Cin2, B, T, F, W, bias, Cout, kf, kt, Y, relu, pool,
                              static_cast<cudaStream_t>(stream));
  amt::ConvGemmDesc d{};
  d.X = X; d.C = Cin; d.X2 = X2; d.C2 = Cin2;
  d.B = B; d.T = T; d.F = F;
  d.W = W; d.bias = bias; d.N = Cout;
  d.kf = kf; d.kt = kt;
  d.out = Y; d.ld_out = Cout;
  d.relu = relu; d.pool = pool; d.out_f32 = 0;
  d.boxF = 16; d.boxT = 8;
  return amt::run_conv_gemm(d, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
