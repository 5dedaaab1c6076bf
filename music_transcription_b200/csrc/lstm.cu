// Persistent LSTM recurrence for sm_100a (nn.LSTM eval forward, gate order i,f,g,o;
// reference models/cnn_rnn_model.py:45-52,69-70 and :212-228,309-312).
//
// The input projections x_t * W_ih^T + b_ih + b_hh are precomputed for all t by the
// tcgen05 GEMM (tc_gemm.cu); this kernel runs only the T dependent steps
//     gates = gx[t] + h_{t-1} * W_hh^T ;  c' = s(f) c + s(i) tanh(g) ;  h' = s(o) tanh(c')
// for several independent sequences at once (forward / reverse directions, the main
// and the local LSTM, and groups of BC chunks).
//
// Two kernels share this decomposition -- one CTA (512 threads) per (batch group of BC chunks, sequence, slice of
// 32 hidden units), 16 warps finishing the cell update (warp w reads TMEM lane quarter w%4, column group w/4; the
// 4 gates of a unit sit in 4 adjacent lanes of ONE warp, so the gate exchange is a warp-private smem transpose;
// every thread owns BC/16 cells whose fp32 cell state never leaves its registers):
//
//   lstm_cluster_kernel (the product path, H <= 512): the slices of a sequence form a thread-block cluster of CTA
//     pairs; W_hh lives in TENSOR MEMORY, h_t travels between the CTAs through distributed shared memory with bulk
//     copies -- see the comment above that kernel.
//   lstm_recurrence_kernel (fallback for hidden sizes whose slices do not fit one cluster): W_hh slice resident in
//     shared memory as the K-major swizzled A operand, h_{t-1} gathered from L2 (published by the sibling slices),
//     one release/acquire counter per step in global memory, cooperative launch so all CTAs are co-resident.
//
// The step is latency-bound, so everything is arranged to shorten the dependent chain: gx prefetched a step ahead,
// no per-thread fences, 4 warps per scheduler for the transcendental-heavy epilogue, outputs stored a step late.
#include <cooperative_groups.h>

#include <cstdlib>

#include <algorithm>

#include "kernels.cuh"

namespace amt {

constexpr int kMaxSeq = 8;
constexpr int kLstmThreads = 512;

struct LstmSeqDev {
  const __nv_bfloat16* whh;
  const float* gx;
  __nv_bfloat16* out_bf16;
  float* out_f32;
  int H, reverse, ld_gx, ld_out, ld_out32, n_slices, cta_begin;
};

struct LstmParams {
  LstmSeqDev seq[kMaxSeq];
  int n_seq, ctas_per_group, B, T, n_groups, Hmax;
  __nv_bfloat16* hbuf;   // [n_groups][n_seq][2][BC][Hmax]
  uint32_t* flags;       // [n_groups][n_seq]
  long long* trace;      // debug: per-phase cycle totals of CTA 0 (AMT_LSTM_TRACE=1), else nullptr
};

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

template <int BC>
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_recurrence_kernel(const LstmParams p) {
  constexpr int NC = BC / 4;                       // accumulator columns per warp
  constexpr int XP = NC + 1;                       // padded pitch of the exchange tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int group = blockIdx.x / p.ctas_per_group;
  const int within = blockIdx.x - group * p.ctas_per_group;
  int q = 0;
  for (int i = 1; i < p.n_seq; ++i)
    if (within >= p.seq[i].cta_begin) q = i;
  const LstmSeqDev sq = p.seq[q];
  const int slice = within - sq.cta_begin;
  const int H = sq.H;
  const int kblocks = H >> 6;

  uint8_t* w_smem = smem;                                       // kblocks x 16 KB
  uint8_t* h_smem = w_smem + kblocks * 16384;                   // kblocks x BC*128 B
  float* xch = reinterpret_cast<float*>(h_smem + kblocks * BC * 128);   // [16 warps][32][XP]
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(xch + 16 * 32 * XP);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

  // ---- one-time setup: barrier, TMEM, resident W_hh slice ----
  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_init(mma_bar, 1);
      ptx::mbar_fence_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, BC < 32 ? 32 : BC);
    ptx::tmem_relinquish();
  }
  {
    const int chunks_per_row = H >> 3;
    const uint4* src = reinterpret_cast<const uint4*>(sq.whh + static_cast<size_t>(slice) * 128 * H);
    for (int e = tid; e < 128 * chunks_per_row; e += kLstmThreads) {
      const int row = e / chunks_per_row;
      const int cc = e - row * chunks_per_row;
      *reinterpret_cast<uint4*>(w_smem + (cc >> 3) * 16384 + ptx::sw128_offset(row, cc & 7)) = __ldg(src + e);
    }
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int b0 = group * BC;                       // first chunk of this batch group
  const int nvalid = min(BC, p.B - b0);            // chunks >= nvalid are padding
  const int quarter = warp & 3;                    // TMEM lane quarter this warp may read
  const int cg = warp >> 2;                        // column group: columns cg*NC .. cg*NC+NC-1
  const int r = quarter * 32 + lane;               // gate row inside the slice: 4*unit_local + gate
  const int gate = lane & 3;
  const int unit = slice * 32 + (r >> 2);          // hidden unit this thread updates
  const int jlane = lane & 3;                      // batch sub-column this thread updates
  // activation as a*sigmoid(k*x)+c: tanh(x) = 2*sigmoid(2x)-1 for the g gate
  const float act_k = gate == 2 ? 2.0f : 1.0f;
  const float act_a = gate == 2 ? 2.0f : 1.0f;
  const float act_c = gate == 2 ? -1.0f : 0.0f;

  __nv_bfloat16* hbuf = p.hbuf + static_cast<size_t>(group * p.n_seq + q) * 2 * BC * p.Hmax;
  uint32_t* flag = p.flags + group * p.n_seq + q;
  const float* gx_row = sq.gx + slice * 128 + r;
  float* xw = xch + warp * 32 * XP;                // warp-private exchange tile
  constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, BC);
  const int chunks_per_row = H >> 3;
  const int gather_total = BC * chunks_per_row;    // 16-byte chunks of h_{t-1}

  const bool mma_leader = ptx::elect_one_sync();   // one lane per warp; only warp 0's is used
  const uint64_t w_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(w_smem));
  const uint64_t h_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(h_smem));

  float cstate[NC / 4];
#pragma unroll
  for (int i = 0; i < NC / 4; ++i) cstate[i] = 0.0f;
  uint32_t parity = 0;

  long long tr[6] = {0, 0, 0, 0, 0, 0};
  const bool tracing = p.trace != nullptr && blockIdx.x == 0 && tid == 0;
#define TRACE_MARK(i) do { if (tracing) { const long long _c = clock64(); tr[i] += _c - tlast; tlast = _c; } } while (0)
  long long tlast = tracing ? clock64() : 0;

  for (int step = 0; step < p.T; ++step) {
    const int t = sq.reverse ? p.T - 1 - step : step;

    // prefetch this step's input projections (independent of h_{t-1})
    float gxv[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int b = cg * NC + j;
      gxv[j] = b < nvalid ? __ldg(gx_row + (static_cast<size_t>(b0 + b) * p.T + t) * sq.ld_gx) : 0.0f;
    }

    if (step > 0) {
      if (tid == 0) {
        const uint32_t target = static_cast<uint32_t>(step) * sq.n_slices;
        while (ptx::ld_acquire_gpu(flag) < target) {
        }
      }
      __syncthreads();
      TRACE_MARK(0);   // gx issue + flag wait
      // gather h_{t-1} (BC x H bf16) from L2 into the swizzled B-operand tile; up to 4 independent
      // 16-byte loads in flight per thread before the first dependent smem store
      const uint4* hsrc = reinterpret_cast<const uint4*>(hbuf + static_cast<size_t>((step - 1) & 1) * BC * p.Hmax);
      for (int e0 = tid; e0 < gather_total; e0 += kLstmThreads * 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * kLstmThreads;
          if (e < gather_total) {
            const int row = e / chunks_per_row;
            v[u] = ptx::ld_cg_v4(hsrc + static_cast<size_t>(row) * (p.Hmax >> 3) + (e - row * chunks_per_row));
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * kLstmThreads;
          if (e < gather_total) {
            const int row = e / chunks_per_row;
            const int cc = e - row * chunks_per_row;
            *reinterpret_cast<uint4*>(h_smem + (cc >> 3) * (BC * 128) + ptx::sw128_offset(row, cc & 7)) = v[u];
          }
        }
      }
      ptx::fence_proxy_async_smem();
      __syncthreads();
      TRACE_MARK(1);   // h gather
      if (warp == 0) {
        // whole warp runs the uniform loop, one elected lane issues (keeps descriptors in uniform
        // registers: ~3 SASS instructions per MMA instead of an ELECT/BRA.U.ANY loop each)
        ptx::tc_fence_after();
        if (mma_leader) {
          for (int kb = 0; kb < kblocks; ++kb) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16_ss(tmem_base, w_desc0 + static_cast<uint64_t>(kb * 1024 + 2 * k),
                                h_desc0 + static_cast<uint64_t>(kb * (BC * 8) + 2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(mma_bar);
        }
        __syncwarp();
      }
      ptx::mbar_wait(mma_bar, parity);
      parity ^= 1;
      ptx::tc_fence_after();
      TRACE_MARK(2);   // MMA issue + completion
    }

    // ---- gates -> activations -> warp-private exchange -> cell update ----
    uint32_t v[NC];
    if (step > 0) {
      ptx::tmem_ld_cols<NC>(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + cg * NC, v);
      ptx::tmem_ld_wait();
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j) v[j] = 0u;
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const float pre = __uint_as_float(v[j]) + gxv[j];
      xw[lane * XP + j] = act_a * sigmoid_fast(act_k * pre) + act_c;
    }
    __syncwarp();
    __nv_bfloat16* hdst = hbuf + static_cast<size_t>(step & 1) * BC * p.Hmax;
    const float* g4 = xw + (lane & ~3) * XP;       // rows of this thread's unit: i, f, g, o
#pragma unroll
    for (int m = 0; m < NC / 4; ++m) {
      const int j = jlane + 4 * m;
      const float gi = g4[j], gf = g4[XP + j], gg = g4[2 * XP + j], go = g4[3 * XP + j];
      const float c = gf * cstate[m] + gi * gg;
      cstate[m] = c;
      const float h = go * (2.0f * sigmoid_fast(2.0f * c) - 1.0f);
      const int b = cg * NC + j;
      const __nv_bfloat16 hb = __float2bfloat16_rn(h);
      hdst[static_cast<size_t>(b) * p.Hmax + unit] = hb;
      if (b < nvalid) {
        const size_t row = static_cast<size_t>(b0 + b) * p.T + t;
        if (sq.out_bf16) sq.out_bf16[row * sq.ld_out + unit] = hb;
        if (sq.out_f32) sq.out_f32[row * sq.ld_out32 + unit] = h;
      }
    }

    // publish h_t to the sibling slices.  bar.sync orders every thread's h stores before thread
    // 0's gpu-scope release (cumulativity), so no per-thread __threadfence() is needed.
    ptx::tc_fence_before();
    TRACE_MARK(3);     // epilogue of this thread
    __syncthreads();
    TRACE_MARK(4);     // wait for the slowest warp
    if (tid == 0) ptx::red_release_gpu_add(flag, 1u);
    TRACE_MARK(5);     // release
  }
  if (tracing)
    for (int i = 0; i < 6; ++i) p.trace[i] = tr[i];
#undef TRACE_MARK

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, BC < 32 ? 32 : BC);
  }
}

static size_t lstm_smem_bytes(int Hmax, int BC) {
  return static_cast<size_t>(Hmax / 64) * 16384 + static_cast<size_t>(Hmax / 64) * BC * 128 +
         16 * 32 * (BC / 4 + 1) * 4 + 64 + 1024;
}

struct LstmPlan {
  int BC, n_groups, ctas_per_group, Hmax;
  size_t flags_bytes, hbuf_bytes;
};

static int lstm_plan(const amt_lstm_seq* seqs, int n_seq, int B, LstmPlan* plan) {
  AMT_REQUIRE(n_seq >= 1 && n_seq <= kMaxSeq, "lstm: n_seq must be in 1..%d", kMaxSeq);
  AMT_REQUIRE(B >= 1, "lstm: empty batch");
  int ctas = 0, Hmax = 0;
  for (int i = 0; i < n_seq; ++i) {
    AMT_REQUIRE(seqs[i].H % 64 == 0 && seqs[i].H >= 64 && seqs[i].H <= 704, "lstm: hidden size %d unsupported (multiple of 64, <= 704: the W_hh slice must fit shared memory)",
                seqs[i].H);
    ctas += seqs[i].H / 32;
    Hmax = seqs[i].H > Hmax ? seqs[i].H : Hmax;
  }
  const int sms = num_sms();
  AMT_REQUIRE(ctas <= sms, "lstm: %d CTAs per batch group exceed the %d SMs", ctas, sms);
  const int max_groups = sms / ctas;
  int BC = 0;
  for (int cand : {16, 32, 64}) {
    if (lstm_smem_bytes(Hmax, cand) > 227 * 1024) break;
    BC = cand;
    if (ceil_div(B, cand) <= max_groups) break;
  }
  AMT_REQUIRE(BC > 0, "lstm: hidden size %d does not fit in shared memory", Hmax);
  plan->BC = BC;
  plan->n_groups = ceil_div(B, BC) < max_groups ? ceil_div(B, BC) : max_groups;   // per launch
  plan->ctas_per_group = ctas;
  plan->Hmax = Hmax;
  plan->flags_bytes = align_up(static_cast<size_t>(plan->n_groups) * n_seq * 4, 256);
  plan->hbuf_bytes = static_cast<size_t>(plan->n_groups) * n_seq * 2 * BC * Hmax * 2;
  return 0;
}

// ============================================================================
// Cluster variant: the slices of one sequence form (part of) a thread-block cluster and
// exchange h_t through distributed shared memory instead of L2.  Per step every CTA stages
// its 32 x BC new h values (pre-swizzled) in its own smem and 16 warps each push one half of that
// block with ONE bulk DSMEM copy (cp.async.bulk.shared::cluster) straight into the double-buffered
// B-operand tile of a peer; the copy completes (complete_tx) on the peer's mbarrier, which the
// consumer waits on.  No L2 round trips, no membar, no spinning on global memory, no cooperative
// launch: clusters are co-scheduled by hardware and independent of each other.
// ============================================================================
constexpr int kMaxClusterCtas = 64;     // CTAs of one batch group (all its clusters)
constexpr int kWCol0 = 64;              // TMEM columns [0,64): accumulator D; [64, 64 + H/2): W_hh slice
constexpr int kTmemColsCluster = 512;

struct LstmClusterParams {
  LstmSeqDev seq[kMaxSeq];
  int n_seq, ctas_per_group, B, T, cluster_size;
  unsigned char cta_seq[kMaxClusterCtas];     // sequence hosted by CTA w of a group (255 = idle padding)
  unsigned char cta_slice[kMaxClusterCtas];   // its slice of 32 hidden units
  unsigned char cta_peer0[kMaxClusterCtas];   // cluster rank of slice 0 of that sequence
  int Hmax;
  long long* trace;
};

template <int BC>
__global__ void __launch_bounds__(kLstmThreads, 1)
lstm_cluster_kernel(const LstmClusterParams p) {
  constexpr int NC = BC / 4;
  constexpr int XP = NC + 1;
  constexpr int kSliceBytes = BC * 64;             // one slice's h_t: BC rows x 32 units bf16, SWIZZLE_64B rows
  constexpr int kHalfBytes = kSliceBytes / 2;      // the BC/2 chunk rows one CTA of a pair keeps (cta_group::2 splits B along N)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int group = blockIdx.x / p.ctas_per_group;
  const int within = blockIdx.x - group * p.ctas_per_group;
  const int q = p.cta_seq[within];
  const bool idle = q == 255;
  const LstmSeqDev sq = p.seq[idle ? 0 : q];
  const int slice = p.cta_slice[within];
  const int peer0 = p.cta_peer0[within];
  const int H = sq.H;
  const int n_peers = sq.n_slices;
  const int hbuf_bytes = n_peers * kHalfBytes;
  const bool pair_leader = (slice & 1) == 0;       // even slice = leader CTA of the pair (cluster ranks 2i, 2i+1)

  // CTA PAIRS: slices 2i and 2i+1 run ONE tcgen05.mma.cta_group::2 per K step (M = 256: 128 gate rows
  // from each CTA's tensor memory); the B operand h_{t-1} is split along N between the two CTAs, so each
  // CTA keeps -- and RECEIVES -- only BC/2 of the BC chunk rows.  The per-step all-to-all of h is bound by
  // the ~17 B/clk/SM DSMEM bandwidth, so halving the bytes per CTA halves the longest phase of the step.
  // W_hh slice lives in TENSOR MEMORY (A operand of tcgen05.mma, TS form): lane = gate row, 32-bit
  // column j of the W region = (W[row][2j], W[row][2j+1]).  The MMA then reads only the small h tile
  // from shared memory (the SS form re-reads the 128 KB slice every step: ~1000 smem-bound cycles).
  // B operand: h_{t-1} as n_peers blocks of [BC/2 rows][32 units] (64-byte rows, SWIZZLE_64B), block s
  // written by slice s -- each block is ONE contiguous region, so a peer delivers it with a single
  // bulk DSMEM copy (rows 0..BC/2-1 of its staging block to the even CTAs, the rest to the odd ones).
  uint8_t* h_smem = smem;                                         // 2 x hbuf_bytes, double buffered
  uint8_t* stage = h_smem + 2 * hbuf_bytes;                       // 2 x kSliceBytes: this slice's new h (pre-swizzled)
  float* xch = reinterpret_cast<float*>(stage + 2 * kSliceBytes); // [16 warps][32][XP]
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(xch + 16 * 32 * XP);
  uint64_t* hbar = mma_bar + 1;                                   // [2 buffers]: this CTA's half of h has landed
  uint64_t* pair_bar = hbar + 2;                                  // [2 buffers] (leader): the odd CTA's half has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pair_bar + 2);

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_init(mma_bar, 1);
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&hbar[i], 1);
        ptx::mbar_init(&pair_bar[i], 1);
      }
      ptx::mbar_fence_init();
    }
    __syncwarp();
    ptx::tmem_alloc_pair(tmem_slot, kTmemColsCluster);     // both CTAs of the pair, same warp, same smem slot
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (!idle) {
    // warp (quarter, cg) fills rows 32*quarter.. of columns [cg*H/8, (cg+1)*H/8) of the W region
    const int q4 = warp & 3, cgw = warp >> 2;
    const int cols_per_warp = H >> 3;                               // 32-bit columns
    const __nv_bfloat16* wrow = sq.whh + (static_cast<size_t>(slice) * 128 + q4 * 32 + lane) * H + cgw * (H >> 2);
    const uint32_t tdst = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + kWCol0 + cgw * cols_per_warp;
    for (int c = 0; c < cols_per_warp; c += 8) {
      const uint4 lo = __ldg(reinterpret_cast<const uint4*>(wrow + 2 * c));
      const uint4 hi = __ldg(reinterpret_cast<const uint4*>(wrow + 2 * c + 8));
      const uint32_t regs[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      ptx::tmem_st_32x32b_x8(tdst + c, regs);
    }
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  ptx::cluster_sync_all();            // every peer's mbarriers are initialised before any remote arrive

  if (!idle) {
    const int b0 = group * BC;
    const int nvalid = min(BC, p.B - b0);
    const int quarter = warp & 3;
    const int cg = warp >> 2;
    const int r = quarter * 32 + lane;
    const int gate = lane & 3;
    const int ul = r >> 2;                           // unit inside the slice
    const int unit = slice * 32 + ul;
    const int jlane = lane & 3;
    // activation as A*tanh(K*x)+C with ONE MUFU op: sigmoid(x) = 0.5*tanh(0.5x)+0.5 for i,f,o; tanh for g
    const float act_k = gate == 2 ? 1.0f : 0.5f;
    const float act_a = gate == 2 ? 1.0f : 0.5f;
    const float act_c = gate == 2 ? 0.0f : 0.5f;
    const float* gx_row = sq.gx + slice * 128 + r;
    float* xw = xch + warp * 32 * XP;
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(256, BC);     // the pair's 256 gate rows x BC chunks
    const bool mma_leader = ptx::elect_one_sync();
    const uint32_t w_tmem = tmem_base + kWCol0;    // A operand: 8 columns (16 bf16) per MMA
    const uint64_t h_desc0 = ptx::umma_desc_sw64(ptx::smem_u32(h_smem));
    // Lane 0 of warp w < n_peers delivers one half of this slice's block to slice (slice + w) % n_peers
    // (rows 0..BC/2-1 to an even slice, the rest to an odd one): 16 different warps issue the 16 bulk
    // copies in parallel.  shared::cluster addresses of the peer's block `slice` (buffer 0) and its hbar[0].
    const int peer_slice = warp < n_peers ? (slice + warp) % n_peers : 0;
    const uint32_t peer_rank = static_cast<uint32_t>(peer0 + peer_slice);
    const uint32_t peer_dst = ptx::mapa(ptx::smem_u32(h_smem + slice * kHalfBytes), peer_rank);
    const uint32_t peer_bar = ptx::mapa(ptx::smem_u32(hbar), peer_rank);
    const uint32_t peer_src_off = static_cast<uint32_t>((peer_slice & 1) * kHalfBytes);
    // odd CTA -> leader: "my half of h has landed";  leader -> both: MMAs of the step retired
    const uint32_t leader_pair_bar = ptx::mapa(ptx::smem_u32(pair_bar), static_cast<uint32_t>(peer0 + (slice & ~1)));
    const uint16_t pair_mask = static_cast<uint16_t>(3u << (peer0 + (slice & ~1)));
    // this thread's cell (unit ul, chunk b) in the pre-swizzled staging block: row b, 16-byte chunk ul/8
    auto stage_off = [&](int b) { return b * 64 + ((((ul >> 3) ^ (b >> 1)) & 3) << 4) + (ul & 7) * 2; };

    float cstate[NC / 4];
#pragma unroll
    for (int i = 0; i < NC / 4; ++i) cstate[i] = 0.0f;
    uint32_t parity = 0;

    long long tr[6] = {0, 0, 0, 0, 0, 0};
    const bool tracing = p.trace != nullptr && blockIdx.x == 0 && tid == 0;
#define TRACE_MARK(i) do { if (tracing) { const long long _c = clock64(); tr[i] += _c - tlast; tlast = _c; } } while (0)
    long long tlast = tracing ? clock64() : 0;

    // input projections are prefetched one full step ahead (DRAM latency never on the step's chain)
    float gxn[NC];
    {
      const int t0 = sq.reverse ? p.T - 1 : 0;
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int b = cg * NC + j;
        gxn[j] = b < nvalid ? __ldg(gx_row + (static_cast<size_t>(b0 + b) * p.T + t0) * sq.ld_gx) : 0.0f;
      }
    }

    // Layer outputs of step s are written during step s+1 (between MMA issue and MMA completion): their
    // stores are then long retired when the next publish runs its gpu-scope membar, which otherwise has
    // to drain them on the dependent chain.
    const bool is_pub = tid < BC * 4;              // publishing thread: chunk pb, 8 units from pc*8 of the slice
    const int pb = tid >> 2, pc = tid & 3;
    uint4 pub_val = make_uint4(0u, 0u, 0u, 0u);
    float hval[NC / 4];
    auto store_outputs = [&](int tt) {
      if (is_pub && sq.out_bf16 && pb < nvalid)
        *reinterpret_cast<uint4*>(sq.out_bf16 + (static_cast<size_t>(b0 + pb) * p.T + tt) * sq.ld_out + slice * 32 + pc * 8) = pub_val;
      if (sq.out_f32) {
#pragma unroll
        for (int m = 0; m < NC / 4; ++m) {
          const int b = cg * NC + jlane + 4 * m;
          if (b < nvalid) sq.out_f32[(static_cast<size_t>(b0 + b) * p.T + tt) * sq.ld_out32 + unit] = hval[m];
        }
      }
    };

    for (int step = 0; step < p.T; ++step) {
      const int t = sq.reverse ? p.T - 1 - step : step;
      float gxv[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) gxv[j] = gxn[j];
      if (step + 1 < p.T) {
        const int tn = sq.reverse ? t - 1 : t + 1;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          const int b = cg * NC + j;
          gxn[j] = b < nvalid ? __ldg(gx_row + (static_cast<size_t>(b0 + b) * p.T + tn) * sq.ld_gx) : 0.0f;
        }
      }

      if (step > 0) {
        const int buf = (step - 1) & 1;
        store_outputs(sq.reverse ? t + 1 : t - 1);
        if (warp == 0) {
          const uint64_t hd = h_desc0 + static_cast<uint64_t>((buf * hbuf_bytes) >> 4);
          const uint32_t hpar = ((step - 1) >> 1) & 1;
          ptx::mbar_wait(&hbar[buf], hpar);                  // this CTA's half of h_{t-1} has landed
          if (!pair_leader) {
            if (lane == 0) ptx::mbar_arrive_remote_relaxed(leader_pair_bar + buf * 8);
          } else {
            ptx::mbar_wait(&pair_bar[buf], hpar);            // ... and so has the odd CTA's half
            TRACE_MARK(0);
            ptx::tc_fence_after();
            if (mma_leader) {
              for (int sb = 0; sb < n_peers; ++sb) {         // one 32-unit block per slice, two K=16 MMAs each
#pragma unroll
                for (int k = 0; k < 2; ++k)
                  ptx::umma_bf16_ts_pair(tmem_base, w_tmem + static_cast<uint32_t>(sb * 16 + k * 8),
                                         hd + static_cast<uint64_t>(sb * (kHalfBytes >> 4) + 2 * k), idesc, (sb | k) != 0 ? 1u : 0u);
              }
              ptx::umma_commit_pair(mma_bar, pair_mask);     // arrives on mma_bar of BOTH CTAs
            }
          }
          __syncwarp();
        }
        ptx::mbar_wait(mma_bar, parity);
        parity ^= 1;
        ptx::tc_fence_after();
        TRACE_MARK(1);
      }

      uint8_t* stage_t = stage + (step & 1) * kSliceBytes;
      uint32_t v[NC];
      if (step > 0) {
        ptx::tmem_ld_cols<NC>(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + cg * NC, v);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < NC; ++j) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const float pre = __uint_as_float(v[j]) + gxv[j];
        xw[lane * XP + j] = fmaf(act_a, ptx::tanh_approx(act_k * pre), act_c);
      }
      __syncwarp();
      const float* g4 = xw + (lane & ~3) * XP;
#pragma unroll
      for (int m = 0; m < NC / 4; ++m) {
        const int j = jlane + 4 * m;
        const float gi = g4[j], gf = g4[XP + j], gg = g4[2 * XP + j], go = g4[3 * XP + j];
        const float c = fmaf(gf, cstate[m], gi * gg);
        cstate[m] = c;
        hval[m] = go * ptx::tanh_approx(c);
        *reinterpret_cast<__nv_bfloat16*>(stage_t + stage_off(cg * NC + j)) = __float2bfloat16_rn(hval[m]);
      }
      ptx::fence_proxy_async_smem();   // generic-proxy staging writes -> async-proxy (bulk copy) reads
      ptx::tc_fence_before();
      __syncthreads();                 // staging block complete; all TMEM reads of this step retired
      TRACE_MARK(2);

      // publish h_t (skipped after the last step: nobody consumes it): 16 lanes each push this slice's
      // 2 KB block with ONE bulk DSMEM copy into block `slice` of buffer step&1 of a peer's B-operand tile;
      // the copy completes (complete_tx) on that peer's hbar.  No L2 round trip, no membar, no barrier:
      // a slice only ever waits for data.  Buffer step&1 of a peer was last read by its MMAs of step-1, and
      // the peer published h_{step-1} -- which this CTA's step needed -- only after those MMAs retired.
      if (is_pub) {
        const int chunk = pc ^ ((pb >> 1) & 3);
        pub_val = *reinterpret_cast<const uint4*>(stage_t + pb * 64 + chunk * 16);
      }
      if (step + 1 < p.T && lane == 0) {
        const int buf = step & 1;
        if (warp == 15) ptx::mbar_expect_tx(&hbar[buf], static_cast<uint32_t>(hbuf_bytes));
        if (warp < n_peers)
          ptx::bulk_copy_to_peer(peer_dst + buf * hbuf_bytes, ptx::smem_u32(stage_t) + peer_src_off, kHalfBytes,
                                 peer_bar + buf * 8);
      }
      TRACE_MARK(3);
    }
    store_outputs(sq.reverse ? 0 : p.T - 1);
    if (tracing)
      for (int i = 0; i < 6; ++i) p.trace[i] = tr[i];
#undef TRACE_MARK
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();             // no CTA exits while a peer may still write its smem / use its TMEM
  if (warp == 0) {
    __syncwarp();
    ptx::tmem_dealloc_pair(tmem_base, kTmemColsCluster);
  }
}

static size_t lstm_cluster_smem_bytes(int Hmax, int BC) {
  return 2 * static_cast<size_t>(Hmax / 32) * (BC / 2) * 64 + 2 * static_cast<size_t>(BC) * 64 +
         16 * 32 * (BC / 4 + 1) * 4 + 128 + 1024;
}

struct ClusterPlan {
  bool ok;
  int BC, CS, ctas_per_group, Hmax;
  unsigned char cta_seq[kMaxClusterCtas], cta_slice[kMaxClusterCtas], cta_peer0[kMaxClusterCtas];
};

template <int BC>
static int lstm_cluster_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int grid, int CS, size_t smem,
                               cudaStream_t stream) {
  AMT_FUNC_ATTR(lstm_cluster_kernel<BC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (CS > 8) AMT_FUNC_ATTR(lstm_cluster_kernel<BC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  *cfg = cudaLaunchConfig_t{};
  cfg->gridDim = dim3(grid);
  cfg->blockDim = dim3(kLstmThreads);
  cfg->dynamicSmemBytes = smem;
  cfg->stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
  return 0;
}

template <int BC>
static int lstm_cluster_max_active(int CS, size_t smem, int* out) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  AMT_TRY(lstm_cluster_config<BC>(&cfg, attr, CS, CS, smem, nullptr));
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, lstm_cluster_kernel<BC>, &cfg);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *out = n;
  return 0;
}

template <int BC>
static int lstm_cluster_launch(const LstmClusterParams& p, int grid, int CS, size_t smem,
                               cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  AMT_TRY(lstm_cluster_config<BC>(&cfg, attr, grid, CS, smem, stream));
  AMT_CUDA(cudaLaunchKernelEx(&cfg, lstm_cluster_kernel<BC>, p));
  count_launch();
  return 0;
}

// Pack sequences into clusters of CS = max(n_slices): a sequence with CS slices fills a cluster,
// smaller sequences of equal H share one (e.g. the two directions of the local LSTM).
static int lstm_cluster_plan(const amt_lstm_seq* seqs, int n_seq, int B, ClusterPlan* plan) {
  plan->ok = false;
  int CS = 0, Hmax = 0;
  for (int i = 0; i < n_seq; ++i) {
    if (seqs[i].H % 64 != 0 || seqs[i].H < 64 || kWCol0 + seqs[i].H / 2 > kTmemColsCluster) return 0;
    CS = std::max(CS, seqs[i].H / 32);
    Hmax = std::max(Hmax, seqs[i].H);
  }
  if (CS > 16) return 0;
  int n_ctas = 0;
  bool placed[kMaxSeq] = {false};
  for (int i = 0; i < n_seq; ++i) {
    if (placed[i]) continue;
    int rank = 0;
    for (int j = i; j < n_seq; ++j) {           // fill this cluster with sequences of the same H
      if (placed[j] || seqs[j].H != seqs[i].H) continue;
      const int ns = seqs[j].H / 32;
      if (rank + ns > CS) break;
      if (n_ctas + rank + ns > kMaxClusterCtas) return 0;
      for (int s = 0; s < ns; ++s) {
        plan->cta_seq[n_ctas + rank + s] = static_cast<unsigned char>(j);
        plan->cta_slice[n_ctas + rank + s] = static_cast<unsigned char>(s);
        plan->cta_peer0[n_ctas + rank + s] = static_cast<unsigned char>(rank);
      }
      rank += ns;
      placed[j] = true;
    }
    for (; rank < CS; ++rank) {                 // idle padding
      if (n_ctas + rank >= kMaxClusterCtas) return 0;
      plan->cta_seq[n_ctas + rank] = 255;
      plan->cta_slice[n_ctas + rank] = 0;
      plan->cta_peer0[n_ctas + rank] = 0;
    }
    n_ctas += CS;
    if (n_ctas > kMaxClusterCtas) return 0;
  }
  const int clusters_per_group = n_ctas / CS;
  int BC = 0;
  for (int cand : {16, 32, 64}) {
    const size_t smem = lstm_cluster_smem_bytes(Hmax, cand);
    if (smem > 227 * 1024) break;
    int max_active = 0;
    if (cand == 16) AMT_TRY(lstm_cluster_max_active<16>(CS, smem, &max_active));
    else if (cand == 32) AMT_TRY(lstm_cluster_max_active<32>(CS, smem, &max_active));
    else AMT_TRY(lstm_cluster_max_active<64>(CS, smem, &max_active));
    if (max_active < 1) break;
    if (getenv("AMT_LSTM_TRACE")) fprintf(stderr, "[lstm plan] BC=%d CS=%d clusters/group=%d max_active_clusters=%d\n", cand, CS, clusters_per_group, max_active);
    BC = cand;
    if (ceil_div(B, cand) * clusters_per_group <= max_active) break;   // everything co-resident
  }
  if (BC == 0) return 0;
  plan->ok = true;
  plan->BC = BC;
  plan->CS = CS;
  plan->ctas_per_group = n_ctas;
  plan->Hmax = Hmax;
  return 0;
}

template <int BC>
static int lstm_launch(const LstmParams& p, int grid, size_t smem, cudaStream_t stream) {
  AMT_FUNC_ATTR(lstm_recurrence_kernel<BC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  void* args[] = {const_cast<LstmParams*>(&p)};
  AMT_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lstm_recurrence_kernel<BC>), dim3(grid),
                                       dim3(kLstmThreads), args, smem, stream));
  count_launch();
  return 0;
}

int run_lstm(const amt_lstm_seq* seqs, int n_seq, int B, int T, void* scratch, size_t scratch_bytes,
             cudaStream_t stream) {
  AMT_TRY(ensure_device());
  AMT_REQUIRE(n_seq >= 1 && n_seq <= kMaxSeq && B >= 1 && T >= 1, "lstm: bad sizes");
  static const bool trace_on = getenv("AMT_LSTM_TRACE") != nullptr;
  static const bool force_l2 = getenv("AMT_LSTM_L2") != nullptr;
  ClusterPlan cp;
  AMT_TRY(lstm_cluster_plan(seqs, n_seq, B, &cp));
  if (cp.ok && !force_l2) {
    LstmClusterParams p{};
    for (int i = 0; i < n_seq; ++i) {
      LstmSeqDev& s = p.seq[i];
      s.whh = static_cast<const __nv_bfloat16*>(seqs[i].whh);
      s.gx = seqs[i].gx;
      s.out_bf16 = static_cast<__nv_bfloat16*>(seqs[i].out_bf16);
      s.out_f32 = seqs[i].out_f32;
      s.H = seqs[i].H;
      s.reverse = seqs[i].reverse;
      s.ld_gx = seqs[i].ld_gx;
      s.ld_out = seqs[i].ld_out;
      s.ld_out32 = seqs[i].ld_out32;
      s.n_slices = seqs[i].H / 32;
      s.cta_begin = 0;
    }
    p.n_seq = n_seq;
    p.ctas_per_group = cp.ctas_per_group;
    p.B = B;
    p.T = T;
    p.cluster_size = cp.CS;
    memcpy(p.cta_seq, cp.cta_seq, sizeof(p.cta_seq));
    memcpy(p.cta_slice, cp.cta_slice, sizeof(p.cta_slice));
    memcpy(p.cta_peer0, cp.cta_peer0, sizeof(p.cta_peer0));
    long long* trace_dev = nullptr;
    if (trace_on) AMT_CUDA(cudaMalloc(&trace_dev, 6 * sizeof(long long)));
    p.trace = trace_dev;
    const int n_groups = ceil_div(B, cp.BC);
    const int grid = n_groups * cp.ctas_per_group;
    const size_t smem = lstm_cluster_smem_bytes(cp.Hmax, cp.BC);
    p.Hmax = cp.Hmax;
    (void)scratch;
    (void)scratch_bytes;          // the cluster path exchanges h through distributed shared memory only
    if (cp.BC == 16) AMT_TRY(lstm_cluster_launch<16>(p, grid, cp.CS, smem, stream));
    else if (cp.BC == 32) AMT_TRY(lstm_cluster_launch<32>(p, grid, cp.CS, smem, stream));
    else AMT_TRY(lstm_cluster_launch<64>(p, grid, cp.CS, smem, stream));
    if (trace_on) {   // debug only: host sync + print
      long long h[6];
      AMT_CUDA(cudaStreamSynchronize(stream));
      AMT_CUDA(cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost));
      cudaFree(trace_dev);
      fprintf(stderr, "[lstm cluster trace] n_seq=%d BC=%d CS=%d grid=%d T=%d cycles/step: h-wait %.0f mma %.0f epilogue %.0f publish %.0f\n",
              n_seq, cp.BC, cp.CS, grid, T, (double)h[0] / T, (double)h[1] / T, (double)h[2] / T, (double)h[3] / T);
    }
    return 0;
  }
  LstmPlan plan;
  AMT_TRY(lstm_plan(seqs, n_seq, B, &plan));
  AMT_REQUIRE(T >= 1, "lstm: T must be >= 1");
  if (scratch_bytes < plan.flags_bytes + plan.hbuf_bytes)
    return set_error(AMT_ERR_WORKSPACE, "lstm: scratch %zu < %zu bytes", scratch_bytes, plan.flags_bytes + plan.hbuf_bytes);
  const int per_launch = plan.n_groups * plan.BC;
  for (int bstart = 0; bstart < B; bstart += per_launch) {
    const int Bl = B - bstart < per_launch ? B - bstart : per_launch;
    LstmParams p{};
    int begin = 0;
    for (int i = 0; i < n_seq; ++i) {
      LstmSeqDev& s = p.seq[i];
      const size_t row0 = static_cast<size_t>(bstart) * T;
      s.whh = static_cast<const __nv_bfloat16*>(seqs[i].whh);
      s.gx = seqs[i].gx + row0 * seqs[i].ld_gx;
      s.out_bf16 = seqs[i].out_bf16 ? static_cast<__nv_bfloat16*>(seqs[i].out_bf16) + row0 * seqs[i].ld_out : nullptr;
      s.out_f32 = seqs[i].out_f32 ? seqs[i].out_f32 + row0 * seqs[i].ld_out32 : nullptr;
      s.H = seqs[i].H;
      s.reverse = seqs[i].reverse;
      s.ld_gx = seqs[i].ld_gx;
      s.ld_out = seqs[i].ld_out;
      s.ld_out32 = seqs[i].ld_out32;
      s.n_slices = seqs[i].H / 32;
      s.cta_begin = begin;
      begin += s.n_slices;
    }
    p.n_seq = n_seq;
    p.ctas_per_group = plan.ctas_per_group;
    p.B = Bl;
    p.T = T;
    p.n_groups = ceil_div(Bl, plan.BC);
    p.Hmax = plan.Hmax;
    p.flags = static_cast<uint32_t*>(scratch);
    p.hbuf = reinterpret_cast<__nv_bfloat16*>(static_cast<uint8_t*>(scratch) + plan.flags_bytes);
    AMT_CUDA(cudaMemsetAsync(p.flags, 0, plan.flags_bytes, stream));
    long long* trace_dev = nullptr;
    if (trace_on) AMT_CUDA(cudaMalloc(&trace_dev, 6 * sizeof(long long)));
    p.trace = trace_dev;
    const int grid = p.n_groups * plan.ctas_per_group;
    const size_t smem = lstm_smem_bytes(plan.Hmax, plan.BC);
    if (plan.BC == 16) AMT_TRY(lstm_launch<16>(p, grid, smem, stream));
    else if (plan.BC == 32) AMT_TRY(lstm_launch<32>(p, grid, smem, stream));
    else AMT_TRY(lstm_launch<64>(p, grid, smem, stream));
    if (trace_on) {   // debug only: host sync + print
      long long h[6];
      AMT_CUDA(cudaStreamSynchronize(stream));
      AMT_CUDA(cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost));
      cudaFree(trace_dev);
      fprintf(stderr, "[lstm trace] n_seq=%d BC=%d grid=%d T=%d cycles/step: wait %.0f gather %.0f mma %.0f epi %.0f barrier %.0f release %.0f\n",
              n_seq, plan.BC, grid, T, (double)h[0] / T, (double)h[1] / T, (double)h[2] / T, (double)h[3] / T,
              (double)h[4] / T, (double)h[5] / T);
    }
  }
  return 0;
}

size_t lstm_scratch_bytes(const amt_lstm_seq* seqs, int n_seq, int B) {
  LstmPlan plan;
  if (lstm_plan(seqs, n_seq, B, &plan) != 0) return 0;
  // cluster path: [groups*n_seq*2][BC][Hmax] bf16 with groups*BC <= B + 63
  const size_t cluster_need = static_cast<size_t>(B + 64) * n_seq * 2 * plan.Hmax * 2;
  return std::max(plan.flags_bytes + plan.hbuf_bytes, cluster_need) + 1024;
}

}  // namespace amt

extern "C" {

size_t amt_lstm_scratch_bytes(const amt_lstm_seq* seqs, int n_seq, int B) {
  return amt::lstm_scratch_bytes(seqs, n_seq, B);
}

int amt_lstm_recurrence(const amt_lstm_seq* seqs_host, int n_seq, int B, int T, void* scratch, size_t scratch_bytes,
                        amt_stream_t stream) {
  return amt::run_lstm(seqs_host, n_seq, B, T, scratch, scratch_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
