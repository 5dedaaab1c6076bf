// tcgen05 / TMA / TMEM fused clamped-softmax attention for sm_100a (head_dim a multiple of 64; the
// canonical model has 8 heads x 192).  Reference models/cnn_rnn_model.py:118-139, the part between the
// qkv and proj Linear layers:
//     S = clamp(Q K^T * hd^-0.5, -clip, +clip);  P = softmax(S);  O = P V
//
// Because the logits are clamped to +-clip (= 10) BEFORE the softmax, exp() is bounded by e^+-10 and a
// plain running sum is exact in fp32: no running maximum, so the un-normalised O accumulates in tensor
// memory over all key blocks without ever being rescaled, and the T x T scores never leave the SM.
//
// One tile = 128 queries of one (chunk, head); keys/values stream in blocks of 64.
//   warp 0   TMA loader: Q once per tile, K and V blocks through two 2-deep rings
//   warp 1   MMA issuer:  S_j = Q K_j^T        (M 128, N 64,  K = hd)   -> TMEM S[j&1]
//                         O  += P_j V_j        (M 128, N hd,  K = 64)   -> TMEM O
//            Both A operands (Q, copied once per tile by the softmax warps, and P) live in tensor memory
//            (TS-form MMAs): shared memory only feeds K and V.
//            V is consumed as it lies in memory ([key][hd], hd contiguous) as an MN-major B operand,
//            so no transposed copy of V exists; S_{j+1} is issued before O += P_j V_j so the tensor
//            pipe works while the softmax warps turn S_j into P_j.
//   warps 4-19  softmax: tcgen05.ld S (two warps per TMEM lane quarter, 32 keys each), scale, clamp,
//            exp, mask keys >= T, row sums in fp32, P as bf16 pairs back into TENSOR MEMORY (tcgen05.st),
//            where it is the A operand of the TS-form MMA O += P V;
//            at the end of the tile O / rowsum -> bf16 -> smem -> TMA store (rows >= T clipped).
// Persistent over (chunk, head, query tile); every ring is indexed by a global key-block counter.
#include "kernels.cuh"

namespace amt {

constexpr int kAttQ = 128;          // queries per tile
constexpr int kAttK = 64;           // keys per block
constexpr int kAttStages = 3;      // K and V rings (a TMA round trip outlasts two blocks of MMAs)
constexpr int kAttParts = 4;        // softmax warps per TMEM lane quarter
constexpr int kAttSoftWarps = 4 * kAttParts;
constexpr int kAttSoftWarp0 = 4;    // (a multiple of 4: warp % 4 is the lane quarter a warp may touch)
constexpr int kAttSoftThreads = 32 * kAttSoftWarps;
constexpr int kAttThreads = 32 * kAttSoftWarp0 + kAttSoftThreads;

struct AttParams {
  int T, heads, B, q_tiles, num_tiles, nb;     // nb = key blocks per tile
  int D;                                       // heads * hd
  float scale, clip;
  long long* trace;                            // AMT_ATT_TRACE builds: per-phase cycle counters of CTA 0
};

#ifdef AMT_ATT_TRACE
#define ATT_TR_BEGIN() long long tr_[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long tr_t_ = clock64(); const long long tr_t0_ = tr_t_
#define ATT_TR(i) do { const long long n_ = clock64(); tr_[i] += n_ - tr_t_; tr_t_ = n_; } while (0)
#define ATT_TR_END(base, who) do { if (blockIdx.x == 0 && (who)) { tr_[9] = clock64() - tr_t0_; for (int i_ = 0; i_ < 10; ++i_) p.trace[(base) + i_] = tr_[i_]; } } while (0)
#else
#define ATT_TR_BEGIN() do { } while (0)
#define ATT_TR(i) do { } while (0)
#define ATT_TR_END(base, who) do { } while (0)
#endif

template <int HD>
struct AttCfg {
  static constexpr int kNB = HD / 64;                       // 64-wide head-dim blocks
  static constexpr int kQBytes = kNB * kAttQ * 128;         // Q tile
  static constexpr int kKVBytes = kNB * kAttK * 128;        // one K (or V) block
  static constexpr int kPBytes = kAttQ * 128;               // one P block (128 q x 64 keys)
  static constexpr int kSmemBytes = kQBytes + 2 * kAttStages * kKVBytes + 2 * kPBytes + 1024 + 256;
  static constexpr int kColO = 128;                         // TMEM: S0 [0,64), S1 [64,128), O [128, 128+HD)
  static constexpr int kColP = 320;                         //       P0 [320,352), P1 [352,384): bf16 pairs, A operand of O += P V
  static constexpr int kColQ = 384;                         //       Q  [384, 384 + HD/2): bf16 pairs, A operand of S = Q K^T
  static constexpr int kTmemCols = 512;
};

// kind::f16 instruction descriptor with an MN-major B operand (bit 16)
__host__ __device__ constexpr uint32_t att_idesc_b_mn(int M, int N) { return ptx::umma_idesc_bf16(M, N) | (1u << 16); }

// MN-major, 128-byte-swizzled operand: 64-element MN atoms `lbo` bytes apart, 8-row K groups 1024 B apart
__device__ __forceinline__ uint64_t att_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int HD>
__global__ void __launch_bounds__(kAttThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const __grid_constant__ CUtensorMap tmO, const AttParams p) {
  using Cfg = AttCfg<HD>;
  constexpr int NB = Cfg::kNB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;
  uint8_t* k_smem = q_smem + Cfg::kQBytes;                  // kAttStages stages
  uint8_t* v_smem = k_smem + kAttStages * Cfg::kKVBytes;    // kAttStages stages
  uint8_t* p_smem = v_smem + kAttStages * Cfg::kKVBytes;    // 2 output staging buffers
  // [kAttParts][128 rows] partial row sums; aliases staging buffer 1, which is first written after the barrier that
  // follows the last read of the sums and is drained before the next tile's sums are written
  float* lsum_x = reinterpret_cast<float*>(p_smem + Cfg::kPBytes);
  static_assert(kAttParts * kAttQ * 4 <= Cfg::kPBytes, "row sums must fit the staging buffer");
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_smem + 2 * Cfg::kPBytes);
  uint64_t* q_full = bars;          // [1]
  uint64_t* q_empty = bars + 1;     // [1]
  uint64_t* k_full = bars + 2;      // [kAttStages]
  uint64_t* k_empty = bars + 5;
  uint64_t* v_full = bars + 8;
  uint64_t* v_empty = bars + 11;
  uint64_t* s_full = bars + 14;     // [2]
  uint64_t* s_empty = bars + 16;
  uint64_t* p_full = bars + 18;
  uint64_t* p_empty = bars + 20;
  uint64_t* o_full = bars + 22;     // [1]
  uint64_t* o_empty = bars + 23;    // [1]
  uint64_t* qt_full = bars + 24;    // [1] Q tile copied into tensor memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);
  static_assert(kAttStages == 3, "barrier layout");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmKV);
    ptx::prefetch_tmap(&tmO);
  }
  if (warp == 1) {
    if (lane == 0) {
      ptx::mbar_init(q_full, 1);
      ptx::mbar_init(q_empty, kAttSoftWarps);
      ptx::mbar_init(qt_full, kAttSoftWarps);
      for (int i = 0; i < kAttStages; ++i) {
        ptx::mbar_init(&k_full[i], 1);
        ptx::mbar_init(&k_empty[i], 1);
        ptx::mbar_init(&v_full[i], 1);
        ptx::mbar_init(&v_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&s_full[i], 1);
        ptx::mbar_init(&s_empty[i], kAttSoftWarps / 2);
        ptx::mbar_init(&p_full[i], kAttSoftWarps / 2);
        ptx::mbar_init(&p_empty[i], 1);
      }
      ptx::mbar_init(o_full, 1);
      ptx::mbar_init(o_empty, kAttSoftWarps);
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

  if (warp == 0) {
    // ------------------------------ TMA loader ------------------------------
    const bool leader = ptx::elect_one_sync();
    uint32_t s = 0, ph = 0, tl = 0;                  // K / V ring slot and phase
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
      int m = tile;
      const int t0 = (m % p.q_tiles) * kAttQ;
      m /= p.q_tiles;
      const int head = m % p.heads;
      const int b = m / p.heads;
      ptx::mbar_wait(q_empty, (tl & 1) ^ 1);
      if (leader) {
        ptx::mbar_expect_tx(q_full, Cfg::kQBytes);
        for (int i = 0; i < NB; ++i) ptx::tma_load_4d(q_smem + i * (kAttQ * 128), &tmQ, q_full, i * 64, head, t0, b);
      }
      __syncwarp();
      for (int j = 0; j < p.nb; ++j) {
        ptx::mbar_wait(&k_empty[s], ph ^ 1);
        if (leader) {
          ptx::mbar_expect_tx(&k_full[s], Cfg::kKVBytes);
          for (int i = 0; i < NB; ++i)
            ptx::tma_load_4d(k_smem + s * Cfg::kKVBytes + i * (kAttK * 128), &tmKV, &k_full[s], i * 64, p.heads + head,
                             j * kAttK, b);
        }
        __syncwarp();
        ptx::mbar_wait(&v_empty[s], ph ^ 1);
        if (leader) {
          ptx::mbar_expect_tx(&v_full[s], Cfg::kKVBytes);
          for (int i = 0; i < NB; ++i)
            ptx::tma_load_4d(v_smem + s * Cfg::kKVBytes + i * (kAttK * 128), &tmKV, &v_full[s], i * 64, 2 * p.heads + head,
                             j * kAttK, b);
        }
        __syncwarp();
        if (++s == kAttStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    const bool leader = ptx::elect_one_sync();
    constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(kAttQ, kAttK);
    constexpr uint32_t idesc_o = att_idesc_b_mn(kAttQ, HD);
    const uint64_t k_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(k_smem));
    const uint64_t v_desc0 = att_desc_mn(ptx::smem_u32(v_smem), kAttK * 128);
    const uint32_t o_tmem = tmem_base + Cfg::kColO;
    uint32_t g = 0, tl = 0;
    ATT_TR_BEGIN();
    // O += P_g V_g for global block gp
    uint32_t ks = 0, kph = 0, vs = 0, vph = 0;      // K / V ring slots and phases (V runs one block behind K)
    auto issue_pv = [&](uint32_t gp, bool first_of_tile) {
      const uint32_t s = gp & 1, ph = (gp >> 1) & 1;
      ptx::mbar_wait(&p_full[s], ph);
      ATT_TR(4);
      ptx::mbar_wait(&v_full[vs], vph);
      ATT_TR(5);
      if (first_of_tile) ptx::mbar_wait(o_empty, (tl & 1) ^ 1);    // previous tile's O has been read out
      ATT_TR(6);
      ptx::tc_fence_after();
      if (leader) {
        const uint32_t pt = tmem_base + Cfg::kColP + s * (kAttK / 2);
        const uint64_t vd = v_desc0 + static_cast<uint64_t>((vs * Cfg::kKVBytes) >> 4);
#pragma unroll
        for (int k = 0; k < kAttK / 16; ++k)       // 16 keys per MMA: 8 TMEM columns of P, 2 KB down the V block
          ptx::umma_bf16_ts(o_tmem, pt + 8 * k, vd + static_cast<uint64_t>(k * (2048 >> 4)), idesc_o,
                            (first_of_tile && k == 0) ? 0u : 1u);
        ptx::umma_commit(&p_empty[s]);
        ptx::umma_commit(&v_empty[vs]);
      }
      __syncwarp();
      if (++vs == kAttStages) { vs = 0; vph ^= 1; }
      ATT_TR(7);
    };
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
      ptx::mbar_wait(qt_full, tl & 1);
      ATT_TR(0);
      for (int j = 0; j < p.nb; ++j, ++g) {
        const uint32_t s = g & 1, ph = (g >> 1) & 1;
        ptx::mbar_wait(&k_full[ks], kph);
        ATT_TR(1);
        ptx::mbar_wait(&s_empty[s], ph ^ 1);
        ATT_TR(2);
        ptx::tc_fence_after();
        if (leader) {
          const uint64_t kd = k_desc0 + static_cast<uint64_t>((ks * Cfg::kKVBytes) >> 4);
#pragma unroll
          for (int i = 0; i < NB; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16_ts(tmem_base + s * kAttK, tmem_base + Cfg::kColQ + i * 32 + k * 8,
                                kd + static_cast<uint64_t>(i * ((kAttK * 128) >> 4) + 2 * k), idesc_s, (i | k) != 0 ? 1u : 0u);
          ptx::umma_commit(&k_empty[ks]);
          ptx::umma_commit(&s_full[s]);
        }
        __syncwarp();
        if (++ks == kAttStages) { ks = 0; kph ^= 1; }
        ATT_TR(3);
        if (j > 0) issue_pv(g - 1, j == 1);
      }
      issue_pv(g - 1, p.nb == 1);
      if (leader) ptx::umma_commit(o_full);
      __syncwarp();
    }
    ATT_TR_END(0, leader);
  } else if (warp >= kAttSoftWarp0) {
    // ------------------------------ softmax + output ------------------------------
    // Four warps share each TMEM lane quarter (warp % 4).  For the softmax they form two groups of 8 warps that take
    // ALTERNATE key blocks (group = S / P buffer index): one group's barrier / tcgen05.ld / tcgen05.st latencies
    // overlap the other's exponentials.  Within a group, a warp handles row q*32 + lane and 32 of the block's 64 keys.
    // For the Q copy and the output epilogue all 16 warps split the columns four ways (`part`).
    constexpr int SW = kAttK / 2;                      // keys per warp and block (softmax)
    constexpr int KW = kAttK / kAttParts;              // output columns per warp and 64-column chunk (epilogue)
    const int q = warp & 3;                            // TMEM lane quarter
    const int part = (warp - kAttSoftWarp0) >> 2;      // 0..3
    const uint32_t grp = part >> 1;                    // softmax group: key blocks with (g & 1) == grp
    const int half = part & 1;                         // which 32 keys of the group's block
    const bool issuer = threadIdx.x == kAttSoftWarp0 * 32;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f, cl2 = p.clip * 1.4426950408889634f;   // work in log2 units
    uint32_t g = 0, tl = 0;
    ATT_TR_BEGIN();
    // Q tile: shared memory (TMA) -> tensor memory, where it is the A operand of the TS-form S = Q K^T for all key
    // blocks of the tile (an smem-resident Q would be re-read by every one of the 15 x 12 MMAs).  This warp's share:
    // row `row`, HD/8/kAttParts sixteen-byte chunks = 4 columns each.
    auto copy_q = [&](uint32_t tq) {
      ptx::mbar_wait(q_full, tq & 1);
      ATT_TR(0);
      constexpr int kChunks = HD / 8 / kAttParts;          // 2 (hd 64), 4 (128), 6 (192)
#pragma unroll
      for (int c2 = 0; c2 < kChunks; c2 += 2) {
        uint32_t w[8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int gc = part * kChunks + c2 + u;          // 16-byte chunk of the row: block gc/8, chunk gc%8
          const uint4 x = *reinterpret_cast<const uint4*>(q_smem + (gc >> 3) * (kAttQ * 128) + ptx::sw128_offset(row, gc & 7));
          w[4 * u] = x.x; w[4 * u + 1] = x.y; w[4 * u + 2] = x.z; w[4 * u + 3] = x.w;
        }
        ptx::tmem_st_32x32b_x8(tmem_base + lane_addr + Cfg::kColQ + (part * kChunks + c2) * 4, w);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(qt_full);
        ptx::mbar_arrive(q_empty);                         // the smem tile may be refilled with the next tile's Q
      }
      ATT_TR(1);
    };
    copy_q(0);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
      int m = tile;
      const int t0 = (m % p.q_tiles) * kAttQ;
      m /= p.q_tiles;
      const int head = m % p.heads;
      const int b = m / p.heads;
      float lsum = 0.0f;
      // This group takes the key blocks j with (j & 1) == jpar in THIS tile (the S / P buffer index runs over a global
      // block counter, so with an odd number of blocks per tile the groups swap roles from tile to tile).  The partial
      // row sums are filed by block parity, not by group, so the final sum always adds the same four partials in the same
      // order: a tile's output does not depend on how many tiles its CTA processed before it (batch invariance).
      const uint32_t jpar = grp ^ (g & 1u);
      ATT_TR(8);
      for (int j = 0; j < p.nb; ++j, ++g) {
        const uint32_t s = g & 1, ph = (g >> 1) & 1;
        if (s != grp) continue;
        ptx::mbar_wait(&s_full[s], ph);
        ATT_TR(2);
        ptx::tc_fence_after();
        uint32_t v[SW];
        ptx::tmem_ld_cols<SW>(tmem_base + lane_addr + s * kAttK + half * SW, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&s_empty[s]);
        const int key0 = j * kAttK + half * SW;
        uint32_t pk[SW / 2];
        float ls0 = 0.0f, ls1 = 0.0f;                // two independent sum chains
        if (key0 + SW <= p.T) {                      // every key of this warp's slice is real: no mask
#pragma unroll
          for (int i = 0; i < SW / 2; ++i) {
            const float p0 = ptx::ex2_approx(fminf(fmaxf(__uint_as_float(v[2 * i]) * sl2, -cl2), cl2));
            const float p1 = ptx::ex2_approx(fminf(fmaxf(__uint_as_float(v[2 * i + 1]) * sl2, -cl2), cl2));
            ls0 += p0;
            ls1 += p1;
            pk[i] = ptx::pack_bf16(p0, p1);
          }
        } else {                                     // the block that straddles T: keys >= T contribute 0
#pragma unroll
          for (int i = 0; i < SW / 2; ++i) {
            const float x0 = fminf(fmaxf(__uint_as_float(v[2 * i]) * sl2, -cl2), cl2);
            const float x1 = fminf(fmaxf(__uint_as_float(v[2 * i + 1]) * sl2, -cl2), cl2);
            const float p0 = key0 + 2 * i < p.T ? ptx::ex2_approx(x0) : 0.0f;
            const float p1 = key0 + 2 * i + 1 < p.T ? ptx::ex2_approx(x1) : 0.0f;
            ls0 += p0;
            ls1 += p1;
            pk[i] = ptx::pack_bf16(p0, p1);
          }
        }
        lsum += ls0 + ls1;
        ATT_TR(3);
        ptx::mbar_wait(&p_empty[s], ph ^ 1);       // O += P_{g-2} V_{g-2} has consumed this buffer
        ATT_TR(4);
        // P stays in TENSOR MEMORY (lane = query row, 32-bit column = two keys): it is the A operand of the
        // TS-form MMA, so it never crosses shared memory
        ptx::tmem_st_32x32b_x16(tmem_base + lane_addr + Cfg::kColP + s * (kAttK / 2) + half * (SW / 2), pk);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[s]);
        ATT_TR(5);
      }
      // The next tile's Q goes to tensor memory as soon as this tile's last S = Q K^T has completed (the tensor pipe
      // retires in order, so that one barrier covers them all; the group that did not consume the block observes its
      // phase without side effects): S of the next tile's first blocks then runs under this tile's epilogue.
      if (tile + static_cast<int>(gridDim.x) < p.num_tiles) {
        ptx::mbar_wait(&s_full[(g - 1) & 1], ((g - 1) >> 1) & 1);
        ptx::tc_fence_after();
        copy_q(tl + 1);
      }
      // ---- tile epilogue: O / rowsum -> bf16 -> smem -> TMA store ----
      lsum_x[(jpar * 2 + half) * kAttQ + row] = lsum;
      ptx::mbar_wait(o_full, tl & 1);              // every MMA of the tile retired
      ATT_TR(6);
      ptx::tc_fence_after();
      ptx::named_bar_sync(1, kAttSoftThreads);
      float tot = 0.0f;
#pragma unroll
      for (int i = 0; i < kAttParts; ++i) tot += lsum_x[i * kAttQ + row];
      const float inv = 1.0f / tot;
#pragma unroll 1
      for (int cc = 0; cc < NB; ++cc) {
        uint32_t v[KW];
        ptx::tmem_ld_cols<KW>(tmem_base + lane_addr + Cfg::kColO + cc * 64 + part * KW, v);
        ptx::tmem_ld_wait();
        if (cc == NB - 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(o_empty);
        }
        uint8_t* obuf = p_smem + (cc & 1) * Cfg::kPBytes;
#pragma unroll
        for (int c = 0; c < KW / 8; ++c) {
          const uint4 val = make_uint4(
              ptx::pack_bf16(__uint_as_float(v[8 * c]) * inv, __uint_as_float(v[8 * c + 1]) * inv),
              ptx::pack_bf16(__uint_as_float(v[8 * c + 2]) * inv, __uint_as_float(v[8 * c + 3]) * inv),
              ptx::pack_bf16(__uint_as_float(v[8 * c + 4]) * inv, __uint_as_float(v[8 * c + 5]) * inv),
              ptx::pack_bf16(__uint_as_float(v[8 * c + 6]) * inv, __uint_as_float(v[8 * c + 7]) * inv));
          *reinterpret_cast<uint4*>(obuf + ptx::sw128_offset(row, part * (KW / 8) + c)) = val;
        }
        ptx::fence_proxy_async_smem();
        if (issuer) ptx::bulk_wait_group_read0();  // (the buffer written next was read by the store one chunk ago)
        ptx::named_bar_sync(1, kAttSoftThreads);
        if (issuer) {
          ptx::tma_store_4d(&tmO, obuf, cc * 64, head, t0, b);
          ptx::bulk_commit_group();
        }
      }
      // the staging buffers are rewritten by the next tile's epilogue: the stores must have read them out
      if (issuer) ptx::bulk_wait_group_read0();
      ptx::named_bar_sync(1, kAttSoftThreads);
      ATT_TR(7);
    }
    ATT_TR_END(10 + 10 * part, lane == 0 && q == 0);
    if (issuer) ptx::bulk_wait_group0();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// HD = head dim padded to a multiple of 64 (the kernel's tile width), hd = the real one (48 / 96 / 144 / 192 for
// hidden 128 .. 512).  The tensor maps describe qkv as [B][T][3 * heads][hd] and the output as [B][T][heads][hd] with
// the REAL hd as the innermost extent: a 64-wide box that reaches past hd is zero-filled on load (the padded q / k
// columns add 0 to every score, the padded v columns produce zeros) and clipped on store.
template <int HD>
static int launch_attention_tc(const void* qkv, void* out, int B, int T, int heads, int hd, float clip, cudaStream_t stream) {
  using Cfg = AttCfg<HD>;
  AMT_REQUIRE(hd <= HD && hd > HD - 64 && hd % 8 == 0, "attention (tcgen05): head_dim %d does not pad to %d", hd, HD);
  AMT_FUNC_ATTR(attention_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  const int D = heads * hd;
  CUtensorMap tq, tkv, to;
  {
    uint64_t dims[4] = {(uint64_t)hd, (uint64_t)3 * heads, (uint64_t)T, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)hd * 2, (uint64_t)3 * D * 2, (uint64_t)T * 3 * D * 2};
    uint32_t boxq[4] = {64, 1, kAttQ, 1}, boxk[4] = {64, 1, kAttK, 1};
    AMT_TRY(encode_tmap_bf16(&tq, qkv, 4, dims, str, boxq, CU_TENSOR_MAP_SWIZZLE_128B));
    AMT_TRY(encode_tmap_bf16(&tkv, qkv, 4, dims, str, boxk, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    uint64_t dims[4] = {(uint64_t)hd, (uint64_t)heads, (uint64_t)T, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)hd * 2, (uint64_t)D * 2, (uint64_t)T * D * 2};
    uint32_t box[4] = {64, 1, kAttQ, 1};
    AMT_TRY(encode_tmap_bf16(&to, out, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  AttParams p;
  p.T = T;
  p.heads = heads;
  p.B = B;
  p.q_tiles = ceil_div(T, kAttQ);
  const long long nt = static_cast<long long>(B) * heads * p.q_tiles;
  AMT_REQUIRE(nt < (1ll << 31), "attention: too many tiles");
  p.num_tiles = static_cast<int>(nt);
  p.nb = ceil_div(T, kAttK);
  p.D = D;
  p.scale = 1.0f / sqrtf(static_cast<float>(hd));
  p.clip = clip;
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  p.trace = nullptr;
#ifdef AMT_ATT_TRACE
  static long long* d_trace = nullptr;
  if (!d_trace) AMT_CUDA(cudaMalloc(&d_trace, 50 * sizeof(long long)));
  p.trace = d_trace;
#endif
  attention_tc_kernel<HD><<<grid, kAttThreads, Cfg::kSmemBytes, stream>>>(tq, tkv, to, p);
  AMT_CHECK_LAUNCH();
#ifdef AMT_ATT_TRACE
  {
    long long h[50];
    AMT_CUDA(cudaStreamSynchronize(stream));
    AMT_CUDA(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
    const int tiles0 = (p.num_tiles + grid - 1) / grid;
    fprintf(stderr, "[att trace] B=%d T=%d hd=%d tiles/CTA=%d blocks/tile=%d\n  mma : qt_full %lld k_full %lld s_empty %lld S-issue %lld p_full %lld v_full %lld o_empty %lld PV-issue %lld total %lld\n",
            B, T, hd, tiles0, p.nb, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[9]);
    for (int w = 0; w < kAttParts; ++w) {
      const long long* t = h + 10 + 10 * w;
      fprintf(stderr, "  soft%d: q_full %lld q-copy %lld s_full %lld ld+exp %lld p_empty %lld st+arrive %lld o_full %lld epilogue %lld misc %lld total %lld\n",
              w, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9]);
    }
  }
#endif
  return 0;
}

int run_attention_tc(const void* qkv, void* out, int B, int T, int heads, int head_dim, float clip, cudaStream_t stream) {
  AMT_REQUIRE(head_dim >= 8 && head_dim <= 192 && head_dim % 8 == 0,
              "attention (tcgen05): head_dim %d unsupported (a multiple of 8 up to 192)", head_dim);
  if (head_dim <= 64) return launch_attention_tc<64>(qkv, out, B, T, heads, head_dim, clip, stream);
  if (head_dim <= 128) return launch_attention_tc<128>(qkv, out, B, T, heads, head_dim, clip, stream);
  return launch_attention_tc<192>(qkv, out, B, T, heads, head_dim, clip, stream);
}

}  // namespace amt
