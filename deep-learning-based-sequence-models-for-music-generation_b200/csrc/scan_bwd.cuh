// scan_bwd.cuh — declarations shared by the two selective-scan backward kernels (scan_bwd.cu: lane<->channel kernel
// for any d_state; scan_bwd_fused.cu: fused recompute/reverse kernel with tensor-pipe channel sums for d_state 32/64).
#pragma once
#include "common.cuh"

namespace mb {

constexpr int kBD = 32;        // channels per CTA
constexpr int kBHelperWarps = 4;
constexpr int kBHelperThreads = kBHelperWarps * 32;
constexpr int kBMaxWarps = 8;  // scan warps per CTA
constexpr int kBRing = 2;      // raw-tile ring depth (a chunk is >= 2 us of work at d_state 64)

struct ScanBwdParams {
  int B, L, D, N, NW, NPT, nck, ck, flags, ntiles, helper_teams;
  const void *u, *delta, *Bm, *Cm, *z, *dout, *ypre;
  int64_t u_bs, u_ls, delta_bs, delta_ls, B_bs, B_ls, C_bs, C_ls, z_bs, z_ls, dout_bs, dout_ls, ypre_bs, ypre_ls;
  void *du, *ddelta, *dz, *dB, *dC;
  int64_t du_bs, du_ls, ddelta_bs, ddelta_ls, dz_bs, dz_ls, dB_bs, dB_ls, dC_bs, dC_ls;
  const float *A, *Dv, *dbias, *ckpt;
  float *dA, *dD, *ddbias;
  // workspace carve-up (fp32)
  float *ws_dB, *ws_dC;  // [B][ntiles][L][N]
  float *ws_dA;          // [B][N][D]
  float *ws_dD, *ws_db;  // [B][D]
  int vec_u, vec_delta, vec_z, vec_dout, vec_ypre, vec_B, vec_C, vec_ck, vec_du, vec_ddelta, vec_dz;
  // fixed_acc: the channel tiles add their dB / dC partials into ONE [B][L][N] accumulator (the memory of ws_dB /
  // ws_dC reused as int64) with 64-bit integer atomics on fixed-point values — integer addition is associative, so
  // the result does not depend on the order in which the tiles arrive (deterministic), and the [B][ntiles][L][N]
  // partial tensors (134 MB written + 269 MB re-read per layer at the training shape) disappear.
  int fixed_acc;
};

// fixed point of the dB / dC accumulators: 2^-36 resolution (1.5e-11), +-1.3e8 range per accumulator
constexpr float kAccScale = 68719476736.f;        // 2^36
constexpr float kAccInvScale = 1.f / 68719476736.f;
__device__ __forceinline__ void red_add_fixed(float* slot_as_i64, float v) {
  const long long q = __float2ll_rn(v * kAccScale);
  asm volatile("red.global.add.u64 [%0], %1;" ::"l"(slot_as_i64), "l"(q) : "memory");
}

template <int NPER>
__device__ __forceinline__ void lds_vec(float (&dst)[NPER], const float* src) {
  if constexpr (NPER == 4) {
    const float4 v = *reinterpret_cast<const float4*>(src);
    dst[0] = v.x, dst[1] = v.y, dst[2] = v.z, dst[3] = v.w;
  } else if constexpr (NPER == 2) {
    const float2 v = *reinterpret_cast<const float2*>(src);
    dst[0] = v.x, dst[1] = v.y;
  } else {
#pragma unroll
    for (int j = 0; j < NPER; ++j) dst[j] = src[j];
  }
}
template <int NPER>
__device__ __forceinline__ void sts_vec(float* dst, const float (&v)[NPER]) {
  if constexpr (NPER == 4) {
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
  } else if constexpr (NPER == 2) {
    *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
  } else {
#pragma unroll
    for (int j = 0; j < NPER; ++j) dst[j] = v[j];
  }
}
__device__ __forceinline__ float2 shfl_xor2(float2 v, int m) {
  return make_float2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
// One reduce-scatter step over the lane pair (lane, lane ^ MASK): on entry every lane holds CNT partial values
// v[0..CNT); on exit v[0..CNT/2) holds the pair-sums of the lower half on lanes with the MASK bit clear and of the
// upper half on lanes with it set.  CNT/2 shuffles instead of CNT (the selects run on the idle ALU pipe); the
// shuffle/shared-memory pipe is the scarce one in this kernel.  With CNT == 1 it is a plain butterfly step.
template <int CNT, int MASK, int LEN>
__device__ __forceinline__ void reduce_scatter_step(float (&v)[LEN], bool hi) {
  if constexpr (CNT >= 2) {
    constexpr int H = CNT / 2;
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const float send = hi ? v[i] : v[i + H];
      const float keep = hi ? v[i + H] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, MASK);
    }
  } else {
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], MASK);
  }
}


// D[16x8] += A[16x8] * B[8x8] on the tensor pipe (legacy mma.sync, tf32 operands, fp32 accumulate).  Fragment
// layout (g = lane >> 2, t = lane & 3): a0 = A[g][t], a1 = A[g+8][t], a2 = A[g][t+4], a3 = A[g+8][t+4];
// b0 = B[t][g], b1 = B[t+4][g]; c0 = D[g][2t], c1 = D[g][2t+1], c2 = D[g+8][2t], c3 = D[g+8][2t+1].
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// scan_bwd_fused.cu: launches the fused kernel only (the caller runs the finalize kernel afterwards).
template <typename T>
int launch_scan_bwd_fused(const ScanBwdParams& p, cudaStream_t stream);

}  // namespace mb
