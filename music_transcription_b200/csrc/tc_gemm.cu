// tcgen05 / TMA / TMEM persistent GEMM for sm_100a:  C[M][N] = A[M][K] * W[N][K]^T + bias  (bf16 in,
// fp32 accumulate, bf16 or f32 out, optional ReLU).
//
// Serves every nn.Linear of the path and the LSTM input projections (reference
// models/cnn_rnn_model.py:45-55, :212-260): x_t W_ih^T for all t at once, qkv / proj of the
// attention block, shared_fc and the stacked frame|onset|offset heads.  (The convolutions have their
// own halo-tile kernel, conv_halo.cu.)
//
// CTA PAIRS (cluster of 2, tcgen05.mma.cta_group::2): one tile = 256 (M) x BN (N); each CTA of the pair
// owns 128 rows of A and of the accumulator and loads only HALF of the weight tile (BN/2 rows) -- the
// tensor cores of both SMs read each other's half, so L2->SMEM traffic and smem reads per MMA drop by a
// third against two independent 128 x BN tiles (the single-CTA kernel was L2/smem-bandwidth bound at
// ~1470 TFLOP/s).  BLOCK_K = 64 bf16 = one 128-B swizzle atom, 6 smem stages.
// Roles (320 threads per CTA): warp 0 = TMA producer (both CTAs; all loads complete on the LEADER's
// full barrier), warp 1 = MMA issuer (leader CTA only; commits multicast to both CTAs' barriers)
// + TMEM alloc, warps 2..9 = epilogue (both CTAs, own 128 rows).  Double-buffered TMEM accumulator
// (the epilogue of tile i overlaps the MMAs of tile i+1).
// Epilogue: two warps per TMEM lane quarter; each 128-byte-wide column chunk of the tile is staged in
// a swizzled shared-memory buffer (double buffered) and written by ONE TMA tensor store, which also
// clips the rows beyond M -- the warps never issue global stores themselves.
#include "kernels.cuh"

namespace amt {

struct GemmParams {
  int kblocks;
  int n_tiles, m_tiles, num_tiles, group_n;   // rasterisation: n-tiles are swept in groups of group_n
  const float* bias;
  int relu;
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

constexpr int kBlockM = 128;                     // rows per CTA; a pair's tile is 2 * kBlockM
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kGemmThreads = 320;
constexpr int kGemmEpiWarp0 = 2;
constexpr int kGemmEpiThreads = 256;
constexpr int kOutChunkBytes = 128 * 128;        // one staged chunk: 128 rows x 128 bytes

template <int BN>
struct GemmCfg {
  static constexpr int kStages = 6;
  static constexpr int kBBytes = (BN / 2) * kBlockK * 2;      // this CTA's half of the weight tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 2 * kOutChunkBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, bool OUT_F32>
__global__ void __launch_bounds__(kGemmThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kChunkCols = OUT_F32 ? 32 : 64;          // output columns per 128-byte staged row
  constexpr int kWarpCols = kChunkCols / 2;              // columns per epilogue warp and chunk
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* o_smem = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(o_smem + 2 * kOutChunkBytes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* tfull = empty + Cfg::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();          // 0 = leader of the pair
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmC);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < Cfg::kStages; ++i) {
        ptx::mbar_init(&full[i], 1);
        ptx::mbar_init(&empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&tfull[i], 1);
        ptx::mbar_init(&tempty[i], 16);                  // 8 epilogue warps of each CTA of the pair
      }
      ptx::mbar_fence_init();
    }
    __syncwarp();
    ptx::tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  ptx::cluster_sync_all();                               // the peer's barriers exist before anything targets them
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    // The whole warp runs the (warp-uniform) loop; only the issuing instructions are predicated on
    // one elected lane, so coordinates / addresses stay in uniform registers.
    const bool leader = ptx::elect_one_sync();
    const uint32_t full0_leader = ptx::mapa(ptx::smem_u32(full), 0);      // stage s: + 8*s
    uint32_t s = 0, ph = 0;
    uint8_t* a_dst = smem;
    // L2 priorities: operands evict_last, the streamed-out C tiles evict_first (epilogue) -- the 1.5 GB of fp32 gate
    // pre-activations a layer-0 projection writes must not push the operand slabs the running tiles share out of
    // L2.  ncu, M 60032 x N 6144 x K 10240: DRAM reads 10.3 -> 9.2 GB, 5.33 -> 5.15 ms (algorithmic 1.36 GB: the
    // rest is re-reads by tiles that share a panel but drift apart in time).
    const uint64_t pol_ld = ptx::l2_policy_evict_last();
    for (int tile = pair; tile < p.num_tiles; tile += num_pairs) {
      int m;
      const int n0 = decode_tile(p, tile, m) * BN + static_cast<int>(rank) * (BN / 2);
      const int m0 = m * (2 * kBlockM) + static_cast<int>(rank) * kBlockM;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        ptx::mbar_wait(&empty[s], ph ^ 1);
        if (leader) {
          if (rank == 0) ptx::mbar_expect_tx(&full[s], 2 * Cfg::kStageBytes);   // both CTAs' bytes land on this barrier
          ptx::tma_load_2d_pair_hint(a_dst, &tmA, full0_leader + 8 * s, kb * kBlockK, m0, pol_ld);
          ptx::tma_load_2d_pair_hint(a_dst + kABytes, &tmB, full0_leader + 8 * s, kb * kBlockK, n0, pol_ld);
        }
        __syncwarp();
        a_dst += Cfg::kStageBytes;
        if (++s == Cfg::kStages) {
          s = 0;
          ph ^= 1;
          a_dst = smem;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA) --------------------
    if (rank == 0) {
      const bool leader = ptx::elect_one_sync();
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(2 * kBlockM, BN);
      const uint64_t desc0 = ptx::umma_desc_sw128(ptx::smem_u32(smem));     // stage 0, A tile, k = 0
      uint32_t s = 0, ph = 0, tl = 0;
      uint64_t a_desc = desc0;
      for (int tile = pair; tile < p.num_tiles; tile += num_pairs, ++tl) {
        const uint32_t acc = tl & 1;
        const uint32_t aph = (tl >> 1) & 1;
        ptx::mbar_wait(&tempty[acc], aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          // descriptors differ only in the 14-bit start-address field (units of 16 bytes)
          const uint64_t b_desc = a_desc + static_cast<uint64_t>(kABytes >> 4);
          if (leader) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              ptx::umma_bf16_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            ptx::umma_commit_pair(&empty[s], 3);   // frees the smem slot of BOTH CTAs once these MMAs retire
          }
          __syncwarp();
          a_desc += static_cast<uint64_t>(Cfg::kStageBytes >> 4);
          if (++s == Cfg::kStages) {
            s = 0;
            ph ^= 1;
            a_desc = desc0;
          }
        }
        if (leader) ptx::umma_commit_pair(&tfull[acc], 3);   // accumulator ready for both epilogues
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ epilogue ----------------------------------
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    const int half = (warp - kGemmEpiWarp0) >> 2;    // which half of every 128-byte column chunk
    const bool issuer = threadIdx.x == kGemmEpiWarp0 * 32;
    const uint64_t pol_st = ptx::l2_policy_evict_first();
    const int r = q * 32 + lane;                     // tile row
    const uint32_t o_row = static_cast<uint32_t>(r) * 128u;
    const uint32_t tempty0_leader = ptx::mapa(ptx::smem_u32(tempty), 0);
    uint32_t tl = 0, chunk_no = 0;
    for (int tile = pair; tile < p.num_tiles; tile += num_pairs, ++tl) {
      int m;
      const int n0 = decode_tile(p, tile, m) * BN;
      const int m0 = m * (2 * kBlockM) + static_cast<int>(rank) * kBlockM;
      const uint32_t acc = tl & 1;
      ptx::mbar_wait(&tfull[acc], (tl >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + half * kWarpCols;
#pragma unroll 1
      for (int c = 0; c < BN / kChunkCols; ++c, ++chunk_no) {
        uint32_t v[kWarpCols];
        ptx::tmem_ld_cols<kWarpCols>(taddr + c * kChunkCols, v);
        ptx::tmem_ld_wait();
        if (c == BN / kChunkCols - 1) {              // accumulator fully read: hand it back to the leader's MMA warp
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_remote_relaxed(tempty0_leader + 8 * acc);
        }
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + c * kChunkCols + half * kWarpCols);
        uint8_t* obuf = o_smem + (chunk_no & 1) * kOutChunkBytes;
        uint4 pk[4];
        if constexpr (OUT_F32) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 bb = __ldg(b4 + j);
            float x0 = __uint_as_float(v[4 * j]) + bb.x, x1 = __uint_as_float(v[4 * j + 1]) + bb.y;
            float x2 = __uint_as_float(v[4 * j + 2]) + bb.z, x3 = __uint_as_float(v[4 * j + 3]) + bb.w;
            if (p.relu) {
              x0 = fmaxf(x0, 0.0f); x1 = fmaxf(x1, 0.0f); x2 = fmaxf(x2, 0.0f); x3 = fmaxf(x3, 0.0f);
            }
            pk[j] = make_uint4(__float_as_uint(x0), __float_as_uint(x1), __float_as_uint(x2), __float_as_uint(x3));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 ba = __ldg(b4 + 2 * j), bb = __ldg(b4 + 2 * j + 1);
            float x0 = __uint_as_float(v[8 * j]) + ba.x, x1 = __uint_as_float(v[8 * j + 1]) + ba.y;
            float x2 = __uint_as_float(v[8 * j + 2]) + ba.z, x3 = __uint_as_float(v[8 * j + 3]) + ba.w;
            float x4 = __uint_as_float(v[8 * j + 4]) + bb.x, x5 = __uint_as_float(v[8 * j + 5]) + bb.y;
            float x6 = __uint_as_float(v[8 * j + 6]) + bb.z, x7 = __uint_as_float(v[8 * j + 7]) + bb.w;
            if (p.relu) {
              x0 = fmaxf(x0, 0.0f); x1 = fmaxf(x1, 0.0f); x2 = fmaxf(x2, 0.0f); x3 = fmaxf(x3, 0.0f);
              x4 = fmaxf(x4, 0.0f); x5 = fmaxf(x5, 0.0f); x6 = fmaxf(x6, 0.0f); x7 = fmaxf(x7, 0.0f);
            }
            pk[j] = make_uint4(ptx::pack_bf16(x0, x1), ptx::pack_bf16(x2, x3), ptx::pack_bf16(x4, x5), ptx::pack_bf16(x6, x7));
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t c16 = static_cast<uint32_t>(half * 4 + j);
          *reinterpret_cast<uint4*>(obuf + o_row + ((c16 ^ (r & 7)) << 4)) = pk[j];
        }
        ptx::fence_proxy_async_smem();
        // the buffer written NEXT (other parity) was last read by the store issued one chunk ago
        if (issuer) ptx::bulk_wait_group_read0();
        ptx::named_bar_sync(1, kGemmEpiThreads);
        if (issuer) {
          ptx::tma_store_2d_hint(&tmC, obuf, n0 + c * kChunkCols, m0, pol_st);
          ptx::bulk_commit_group();
        }
      }
    }
    if (issuer) ptx::bulk_wait_group0();             // all output stores complete before the CTA exits
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();                               // the leader's MMAs also wrote the peer's tensor memory
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}


// ----------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------
template <int BN, bool OUT_F32>
static int launch(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const GemmParams& p,
                  cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  AMT_FUNC_ATTR((tc_gemm_kernel<BN, OUT_F32>), cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  const int pairs = p.num_tiles < num_sms() / 2 ? p.num_tiles : num_sms() / 2;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  AMT_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<BN, OUT_F32>, a, b, c, p));
  count_launch();
  return 0;
}

int run_gemm(const void* A, const void* W, const float* bias, void* C, int M, int N, int K, long long ldc, int relu,
             int out_f32, cudaStream_t stream) {
  AMT_TRY(ensure_device());
  AMT_REQUIRE(M > 0, "gemm: empty problem");
  AMT_REQUIRE(K % 64 == 0 && K > 0, "gemm: K (%d) must be a positive multiple of 64", K);
  AMT_REQUIRE(N % 64 == 0 && N > 0, "gemm: N (%d) must be a positive multiple of 64", N);
  const int esize = out_f32 ? 4 : 2;
  AMT_REQUIRE(ldc >= N && (ldc * esize) % 16 == 0, "gemm: ldc (%lld) must be >= N with 16-byte aligned rows", ldc);
  AMT_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
              "gemm: C and bias must be 16-byte aligned");
  const int BN = N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : 64);

  CUtensorMap am, bm, cm;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {64, 128};
    AMT_TRY(encode_tmap_bf16(&am, A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {64, (uint32_t)(BN / 2)};      // each CTA of a pair loads half of the weight tile
    AMT_TRY(encode_tmap_bf16(&bm, W, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)ldc * esize};
    uint32_t box[2] = {(uint32_t)(out_f32 ? 32 : 64), 128};
    if (out_f32) AMT_TRY(encode_tmap_f32(&cm, C, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
    else AMT_TRY(encode_tmap_bf16(&cm, C, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }

  GemmParams p;
  p.kblocks = K / 64;
  p.n_tiles = N / BN;
  p.m_tiles = ceil_div(M, 2 * kBlockM);
  // n-tiles per rasterisation group: as many as keep the group's weight slab (group_n x BN x K bf16) around
  // 42 MB, i.e. L2-resident next to the streaming activations -- 8 tiles for K = 10240 (measured best),
  // every tile for the K <= 1536 projections, whose activations are then read from DRAM once.
  const long long tile_bytes = static_cast<long long>(BN) * K * 2;
  long long g = (42ll << 20) / tile_bytes;
  g = g < 1 ? 1 : g;
  p.group_n = p.n_tiles > g ? static_cast<int>(g) : p.n_tiles;
  const long long nt = static_cast<long long>(p.m_tiles) * p.n_tiles;
  AMT_REQUIRE(nt < (1ll << 31), "gemm: too many tiles");
  p.num_tiles = static_cast<int>(nt);
  p.bias = bias;
  p.relu = relu;

  if (out_f32) {
    if (BN == 256) return launch<256, true>(am, bm, cm, p, stream);
    if (BN == 128) return launch<128, true>(am, bm, cm, p, stream);
    return launch<64, true>(am, bm, cm, p, stream);
  }
  if (BN == 256) return launch<256, false>(am, bm, cm, p, stream);
  if (BN == 128) return launch<128, false>(am, bm, cm, p, stream);
  return launch<64, false>(am, bm, cm, p, stream);
}

}  // namespace amt

extern "C" {

int amt_gemm_bf16(const void* A, const void* W, const float* bias, void* C, int M, int N, int K, int ldc, int relu,
                  int out_f32, amt_stream_t stream) {
  return amt::run_gemm(A, W, bias, C, M, N, K, ldc, relu, out_f32, static_cast<cudaStream_t>(stream));
}

int amt_conv_bf16(const void* X, const void* X2, const void* W, const float* bias, void* Y, int B, int T, int F,
                  int Cin, int Cin2, int Cout, int kf, int kt, int relu, int pool, amt_stream_t stream) {
  return amt::run_conv_halo(X, Cin, X2, Cin2, B, T, F, W, bias, Cout, kf, kt, Y, relu, pool & 1, (pool >> 1) & 1,
                            static_cast<cudaStream_t>(stream));
}

int amt_split3_bf16(const void* x, int in_f32, void* out, int64_t rows, int K, amt_stream_t stream) {
  AMT_REQUIRE(x && out, "split3: NULL argument");
  AMT_TRY(amt::ensure_device());
  return amt::run_split3(x, in_f32, out, rows, K, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
