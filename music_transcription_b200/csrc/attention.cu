// Fused clamped-softmax attention (reference models/cnn_rnn_model.py:118-139, the part
// between the qkv and proj Linear layers):
//     S = clamp(Q K^T * hd^-0.5, -clip, +clip);  P = softmax(S);  O = P V
// The T x T score matrix never touches HBM.  Because the logits are clamped to
// +-clip (=10) BEFORE the softmax, exp() is bounded by e^+-10, so a single pass with a
// plain running sum (no running-max rescaling) is exact in fp32; keys beyond T are
// masked explicitly (a padded zero logit would contribute e^0, not 0).
//
// This file is the generic fallback for head dims that are not a multiple of 64 (hidden sizes other
// than 512): warp-level mma.sync (m16n8k16 bf16, fp32 accumulate), 64 queries per CTA (4 warps x 16
// rows), key/value blocks of 64 staged in shared memory with cp.async.  Head dims 64/128/192 (the
// canonical model: 8 x 192) run the tcgen05 kernel in attention_tc.cu.
#include <cstdlib>

#include "kernels.cuh"

namespace amt {

__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3,
                                                  uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;    // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

constexpr int kQTile = 64;
constexpr int kKTile = 64;

template <int HD>
__global__ void __launch_bounds__(128) attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                        __nv_bfloat16* __restrict__ out, int T, int heads, float scale,
                                                        float clip) {
  constexpr int PITCH = HD + 8;                 // elements; (PITCH*2/16) odd -> conflict-free ldmatrix
  extern __shared__ __align__(16) uint8_t smem_att[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_att);
  __nv_bfloat16* sK = sQ + kQTile * PITCH;
  __nv_bfloat16* sV = sK + kKTile * PITCH;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kQTile;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int D = heads * HD;
  const size_t ld = 3 * static_cast<size_t>(D);
  const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * T * ld + head * HD;
  constexpr int CPR = HD / 8;                   // 16-byte chunks per row

  for (int e = tid; e < kQTile * CPR; e += 128) {
    const int r = e / CPR, c = e - r * CPR;
    const bool ok = q0 + r < T;
    cp_async16(ptx::smem_u32(sQ + r * PITCH + c * 8), base + static_cast<size_t>(ok ? q0 + r : 0) * ld + c * 8, ok);
  }

  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f;
  float lsum0 = 0.0f, lsum1 = 0.0f;             // rows g and g+8 of this warp's 16 queries

  const int g = lane >> 2, tq = lane & 3;
  const uint32_t sQ_u = ptx::smem_u32(sQ), sK_u = ptx::smem_u32(sK), sV_u = ptx::smem_u32(sV);

  for (int k0 = 0; k0 < T; k0 += kKTile) {
    __syncthreads();                            // previous block fully consumed
    for (int e = tid; e < kKTile * CPR; e += 128) {
      const int r = e / CPR, c = e - r * CPR;
      const bool ok = k0 + r < T;
      const __nv_bfloat16* src = base + static_cast<size_t>(ok ? k0 + r : 0) * ld + c * 8;
      cp_async16(sK_u + (r * PITCH + c * 8) * 2, src + D, ok);
      cp_async16(sV_u + (r * PITCH + c * 8) * 2, src + 2 * D, ok);
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- S = Q K^T for this warp's 16 queries x 64 keys ----
    float s[kKTile / 8][4];
#pragma unroll
    for (int i = 0; i < kKTile / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < HD / 16; kk += 2) {    // two k-steps (32 d) per iteration
      uint32_t qa[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (kk + u < HD / 16) {
          const int row = warp * 16 + (lane & 15);
          const int col = (kk + u) * 16 + (lane >> 4) * 8;
          ldmatrix_x4(qa[u][0], qa[u][1], qa[u][2], qa[u][3], sQ_u + (row * PITCH + col) * 2);
        }
      }
#pragma unroll
      for (int nt = 0; nt < kKTile / 8; ++nt) {
        // x4: (keys nt*8.., d kk*16 .. +31) -> b0,b1 of k-step kk and b0,b1 of k-step kk+1
        uint32_t kb0, kb1, kb2, kb3;
        const int row = nt * 8 + (lane & 7);
        int col = kk * 16 + (lane >> 3) * 8;
        if (col >= HD) col = HD - 8;             // odd number of k-steps: second half unused
        ldmatrix_x4(kb0, kb1, kb2, kb3, sK_u + (row * PITCH + col) * 2);
        mma_bf16_16816(s[nt], qa[0][0], qa[0][1], qa[0][2], qa[0][3], kb0, kb1);
        if (kk + 1 < HD / 16) mma_bf16_16816(s[nt], qa[1][0], qa[1][1], qa[1][2], qa[1][3], kb2, kb3);
      }
    }

    // ---- clamp, exp, mask, row sums; P as bf16 A fragments ----
    uint32_t pa[kKTile / 16][4];
#pragma unroll
    for (int nt = 0; nt < kKTile / 8; ++nt) {
      const int key = k0 + nt * 8 + tq * 2;
      float p[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float x = fminf(fmaxf(s[nt][i] * scale, -clip), clip);
        const bool ok = key + (i & 1) < T;
        p[i] = ok ? __expf(x) : 0.0f;
      }
      lsum0 += p[0] + p[1];
      lsum1 += p[2] + p[3];
      pa[nt >> 1][(nt & 1) * 2 + 0] = ptx::pack_bf16(p[0], p[1]);
      pa[nt >> 1][(nt & 1) * 2 + 1] = ptx::pack_bf16(p[2], p[3]);
    }

    // ---- O += P V ----
#pragma unroll
    for (int ks = 0; ks < kKTile / 16; ++ks) {
#pragma unroll
      for (int nt = 0; nt < HD / 8; nt += 2) {
        // x4.trans: V[keys ks*16 + 0..15][d nt*8 .. +15] -> (b0,b1) for n-tile nt and nt+1
        uint32_t v0, v1, v2, v3;
        const int row = ks * 16 + (lane & 15);
        const int col = nt * 8 + (lane >> 4) * 8;
        ldmatrix_x4_trans(v0, v1, v2, v3, sV_u + (row * PITCH + col) * 2);
        mma_bf16_16816(o[nt], pa[ks][0], pa[ks][1], pa[ks][2], pa[ks][3], v0, v1);
        mma_bf16_16816(o[nt + 1], pa[ks][0], pa[ks][1], pa[ks][2], pa[ks][3], v2, v3);
      }
    }
  }

  // ---- normalise and store ----
  lsum0 += __shfl_xor_sync(0xffffffffu, lsum0, 1);
  lsum0 += __shfl_xor_sync(0xffffffffu, lsum0, 2);
  lsum1 += __shfl_xor_sync(0xffffffffu, lsum1, 1);
  lsum1 += __shfl_xor_sync(0xffffffffu, lsum1, 2);
  const float inv0 = 1.0f / lsum0, inv1 = 1.0f / lsum1;
  const int qr0 = q0 + warp * 16 + g, qr1 = qr0 + 8;
  __nv_bfloat16* obase = out + static_cast<size_t>(b) * T * D + head * HD;
#pragma unroll
  for (int nt = 0; nt < HD / 8; ++nt) {
    const int col = nt * 8 + tq * 2;
    if (qr0 < T) *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(qr0) * D + col) = ptx::pack_bf16(o[nt][0] * inv0, o[nt][1] * inv0);
    if (qr1 < T) *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(qr1) * D + col) = ptx::pack_bf16(o[nt][2] * inv1, o[nt][3] * inv1);
  }
}

template <int HD>
static int launch_attention(const void* qkv, void* out, int B, int T, int heads, float clip, cudaStream_t stream) {
  constexpr int PITCH = HD + 8;
  const int smem = (kQTile + 2 * kKTile) * PITCH * 2;
  AMT_FUNC_ATTR(attention_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  dim3 grid(ceil_div(T, kQTile), heads, B);
  attention_kernel<HD><<<grid, 128, smem, stream>>>(static_cast<const __nv_bfloat16*>(qkv),
                                                     static_cast<__nv_bfloat16*>(out), T, heads,
                                                     1.0f / sqrtf(static_cast<float>(HD)), clip);
  AMT_CHECK_LAUNCH();
  return 0;
}

int run_attention(const void* qkv, void* out, int B, int T, int heads, int head_dim, float clip, cudaStream_t stream) {
  AMT_TRY(ensure_device());
  AMT_REQUIRE(B > 0 && T > 0 && heads > 0, "attention: empty problem");
  static const bool force_sync = getenv("AMT_ATT_MMA_SYNC") != nullptr;    // bring-up switch
  // every head dim of the path (48 / 96 / 144 / 192 = hidden 128 .. 512) runs the tcgen05 kernel, padded to the
  // next multiple of 64 by the tensor maps' zero fill; the mma.sync kernel below stays as a bring-up cross-check
  if (!force_sync && head_dim % 8 == 0 && head_dim <= 192)
    return run_attention_tc(qkv, out, B, T, heads, head_dim, clip, stream);
  switch (head_dim) {
    case 48: return launch_attention<48>(qkv, out, B, T, heads, clip, stream);
    case 96: return launch_attention<96>(qkv, out, B, T, heads, clip, stream);
    case 144: return launch_attention<144>(qkv, out, B, T, heads, clip, stream);
    case 192: return launch_attention<192>(qkv, out, B, T, heads, clip, stream);
    default:
      return set_error(AMT_ERR_ARG, "attention: head_dim %d unsupported (48, 96, 144, 192 = hidden 128..512)", head_dim);
  }
}

}  // namespace amt

extern "C" int amt_attention_bf16(const void* qkv, void* out, int B, int T, int heads, int head_dim, float clip,
                                  amt_stream_t stream) {
  return amt::run_attention(qkv, out, B, T, heads, head_dim, clip, static_cast<cudaStream_t>(stream));
}
