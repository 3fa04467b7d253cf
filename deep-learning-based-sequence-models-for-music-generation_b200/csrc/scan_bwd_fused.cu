// scan_bwd_fused.cu — selective-scan backward for d_state 32 / 64 (the repo's training shape), sm_100a.
//
// Same mathematics as scan_bwd.cu (adjoint of MambaBlock.selective_scan, simple_mamba.cpython-311.pyc @L310-333:
// in-chunk recompute from the forward's checkpoints, dh_t = C_t*dy_t + a_{t+1}*dh_{t+1} in registers), organised
// around the two things that bound that kernel at d_state 64 — issue slots spent on cross-lane reductions, and the
// MUFU-heavy recompute phase never overlapping the FMA-heavy reverse phase:
//
//   * ONE LOOP does both phases: while the reverse sweep walks chunk c from its last step to its first, the same
//     instruction stream recomputes chunk c-1 from its first step to its last (two steps of each per iteration), so
//     every warp always has independent MUFU (recompute) and FMA (reverse) work in flight.  The h_t of the even
//     steps go through thread-private shared memory; the slot the reverse sweep has just read is the slot the
//     recompute writes next (slot order alternates per chunk), so one chunk's worth of slots is enough.
//   * sums over CHANNELS (dB, dC) run on the tensor pipe.  warp = (channel octet, state half); lane = (g = lane >> 2:
//     quad of 4 states, t = lane & 3: channel pair); a thread owns 2 channels x 4 states.  Its values ARE an
//     mma.m16n8k8 A-fragment (rows = states, k = the octet's 8 channels); the B-fragment is one-hot in the column of
//     the current timestep, so eight steps accumulate in 16 registers and are flushed with four 16-byte stores.
//     fp32 I/O: A = the fp32 products split into tf32 hi + lo (two MMAs, error ~2^-21); bf16 I/O: A = dh / h
//     directly, B = one-hot * (delta*u | dy) (operands truncated to tf32, 2^-10, far inside the bf16 tolerance).
//   * sums over STATES: 4 states inside the thread, then 4 values over the 8 g-lanes (4 shuffles).
//   * helper warps (4) as in scan_bwd.cu, running two chunks ahead of the reverse sweep: cp.async ring (2 slots),
//     pre-pass into a 3-slot operand ring (softplus(delta+bias), delta*u, dy = dout*silu(z), B/C as fp32), post-pass
//     from a 2-slot result ring (re-reads its few per-(t,d) inputs through L2 instead of pinning the raw tiles).
//   * registers: the scan warpgroups raise their budget with setmaxnreg (helpers give theirs up).
#include "scan_bwd.cuh"

namespace mb {

constexpr int kFIn = 3;   // operand ring depth (chunk it is read by the reverse sweep of iteration it and by the
                          // recompute of iteration it-1, and once more by the post-pass)
constexpr int kFRaw = 2;  // cp.async ring depth
constexpr int kFOut = 2;  // result ring depth

constexpr int align128(int x) { return (x + 127) & ~127; }

// Shared-memory map (all offsets compile-time constants so that every shared access is base register + immediate).
template <typename T, int NW, int kCK>
struct FusedLayout {
  static constexpr int S = (int)sizeof(T);
  static constexpr int NPT = NW * 8;                 // d_state
  static constexpr int RS = 2 * NPT + 4;             // row stride (floats) of the dB|dC octet partials: == 4 mod 16
  static constexpr int NSH = NW >= 4 ? NW / 4 : 1;   // state halves
  // raw slot (cp.async targets, element type T)
  static constexpr int raw_u = 0;
  static constexpr int raw_dl = raw_u + kCK * kBD * S;
  static constexpr int raw_z = raw_dl + kCK * kBD * S;
  static constexpr int raw_do = raw_z + kCK * kBD * S;
  static constexpr int raw_B = raw_do + kCK * kBD * S;
  static constexpr int raw_C = raw_B + kCK * NPT * S;
  static constexpr int raw_bytes = align128(raw_C + kCK * NPT * S);
  // operand slot (fp32)
  static constexpr int w_dl = 0;
  static constexpr int w_du = w_dl + kCK * kBD * 4;
  static constexpr int w_dy = w_du + kCK * kBD * 4;
  static constexpr int w_Bf = w_dy + kCK * kBD * 4;
  static constexpr int w_Cf = w_Bf + kCK * NPT * 4;
  static constexpr int in_bytes = align128(w_Cf + kCK * NPT * 4);
  // result slot (fp32)
  static constexpr int o_pg = 0;
  static constexpr int o_pS = o_pg + NSH * kCK * kBD * 4;
  static constexpr int o_bc = o_pS + NSH * kCK * kBD * 4;  // [channel octet][step][dB | dC | pad]
  static constexpr int out_bytes = align128(o_bc + 4 * kCK * RS * 4);
  static constexpr int off_raw = 0;
  static constexpr int off_in = off_raw + kFRaw * raw_bytes;
  static constexpr int off_out = off_in + kFIn * in_bytes;
  static constexpr int off_hs = off_out + kFOut * out_bytes;
  static constexpr int hs_slot = 2 * NW * 32 * 16;          // one even step: float4 [2 channels][scan threads]
  static constexpr int total = off_hs + (kCK / 2) * hs_slot;
};

// shared-memory accesses by 32-bit shared address (no generic->shared conversion inside the loops)
__device__ __forceinline__ float2 lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

// 4 consecutive channels of one (b, t) row from global memory (post-pass operands; L2 hits)
template <typename T>
__device__ __forceinline__ void ld4_global(const T* p, bool vec, int nvalid, float (&v)[4]) {
  if (vec && nvalid >= 4) {
    V4<T>::ld(p, v);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = e < nvalid ? IO<T>::cvt(p[e]) : 0.f;
  }
}

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

template <typename T, int NW, int kCK>
__global__ void __launch_bounds__((NW + kBHelperWarps) * 32, 1) scan_bwd_fused_kernel(const ScanBwdParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  using Lay = FusedLayout<T, NW, kCK>;
  constexpr Lay lay{};
  constexpr int NPT = Lay::NPT, RS = Lay::RS, NSH = Lay::NSH;
  constexpr int K = kCK / 2;        // (odd, even) step pairs per chunk
  constexpr int nscan_threads = NW * 32;
  constexpr int bar_count = nscan_threads + kBHelperThreads;
  constexpr bool kF32 = sizeof(T) == 4;
  const int b = blockIdx.y, tile = blockIdx.x;
  const int d0 = tile * kBD;
  const int dvalid = min(kBD, p.D - d0);
  unsigned char* const raw_base = smem + Lay::off_raw;
  unsigned char* const in_base = smem + Lay::off_in;
  unsigned char* const out_base = smem + Lay::off_out;
  const int nck = p.nck;
  const bool has_z = p.flags & MAMBA_FLAG_HAS_Z;
  // named barriers: 1..3 = READY[operand slot], 4..5 = DONE[result slot], 6 = helpers only.
  // Iteration `it` handles chunk nck-1-it.

  if (warp < NW) {
    // ============================================ SCAN WARPS ===============================================
    if constexpr (NW == 8) setmaxnreg_inc<208>();
    const int t4 = lane & 3, g = lane >> 2;
    const int cg = warp & 3, sh = warp >> 2;
    const int cl = 8 * cg + 2 * t4;  // first (tile-local) channel of this thread
    const int n0 = 32 * sh + 4 * g;  // first state of this thread
    const int d = d0 + cl;
    float2 A2[2][2], dAacc[2][2], dhc[2][2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = d + ch < p.D ? load_A(p.A, (int64_t)(d + ch) * p.N + n0 + j, p.flags) * kLog2e : 0.f;
        if (j & 1) A2[ch][j / 2].y = a;
        else A2[ch][j / 2].x = a;
        dAacc[ch][j / 2] = dhc[ch][j / 2] = make_float2(0.f, 0.f);
      }
    // checkpoint layout [B][nck][N/4][D][4] (scan_fwd.cu); chunk 0 starts from h = 0
    auto load_ckpt = [&](int c, float2 (&h)[2][2]) {
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c > 0 && d + ch < p.D)
          v = __ldg(reinterpret_cast<const float4*>(p.ckpt) + (((int64_t)b * nck + c) * (NPT / 4) + (n0 >> 2)) * p.D + d + ch);
        h[ch][0] = make_float2(v.x, v.y), h[ch][1] = make_float2(v.z, v.w);
      }
    };
    float accB[2][4], accC[2][4];  // [state pair][mma accumulator]: 16 states x 8 timesteps per warp
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) accB[i][k] = accC[i][k] = 0.f;
    const bool st_writer = lane < 16;

    // 32-bit shared addresses of this thread's operands (row strides: 128 B per step for the per-channel arrays,
    // NPT*4 B per step for B / C); every access below is one of these registers + an immediate
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t in0 = sbase + Lay::off_in + cl * 4;           // + slot*in_bytes + {w_dl,w_du,w_dy} + t*128
    const uint32_t bc0 = sbase + Lay::off_in + Lay::w_Bf + n0 * 4;  // + slot*in_bytes + {0, w_Cf-w_Bf} + t*NPT*4
    const uint32_t hs0 = sbase + Lay::off_hs + tid * 16;         // + slot*hs_slot + ch*nscan_threads*16
    const uint32_t st0 = sbase + Lay::off_out + ((lane & 8) ? Lay::o_pS : Lay::o_pg) +
                         (sh * kCK * kBD + cl + ((lane & 4) ? 1 : 0)) * 4;          // + oslot*out_bytes + t*128
    const uint32_t pb0 = sbase + Lay::off_out + Lay::o_bc + ((cg * kCK + 2 * t4) * RS + n0) * 4;  // + oslot*.. + t*RS*4
    constexpr int kRowC = kBD * 4, kRowS = NPT * 4, kCoff = Lay::w_Cf - Lay::w_Bf;
    constexpr int kHch = nscan_threads * 16;

    // one recompute step: h <- exp2(delta*A) * h + delta*u*B
    auto rec_step = [&](const float2 dlv, const float2 duv, const float4 B4, float2 (&h)[2][2]) {
      const float2 Bp[2] = {make_float2(B4.x, B4.y), make_float2(B4.z, B4.w)};
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const float dl = ch ? dlv.y : dlv.x, du = ch ? duv.y : duv.x;
        const float2 dl2 = make_float2(dl, dl), du2 = make_float2(du, du);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float2 ga = __fmul2_rn(dl2, A2[ch][r]);
          const float2 a = make_float2(ex2_approx(ga.x), ex2_approx(ga.y));
          h[ch][r] = __ffma2_rn(a, h[ch][r], __fmul2_rn(du2, Bp[r]));
        }
      }
    };
    auto hs_store = [&](uint32_t a, const float2 (&h)[2][2]) {
      sts128(a, make_float4(h[0][0].x, h[0][0].y, h[0][1].x, h[0][1].y));
      sts128(a + kHch, make_float4(h[1][0].x, h[1][0].y, h[1][1].x, h[1][1].y));
    };

    // one reverse step; returns this lane's state-sum value BEFORE the last butterfly stage (lane ^ 16), which the
    // caller finishes one loop iteration later so that the shuffle latency is off the critical path
    auto rev_compute = [&](const float2 dlv, const float2 duv, const float2 dyv, const float4 B4, const float4 C4,
                           const float2 (&hprev)[2][2], const bool odd, const bool onehot) -> float {
      const float2 Bp[2] = {make_float2(B4.x, B4.y), make_float2(B4.z, B4.w)};
      const float2 Cp[2] = {make_float2(C4.x, C4.y), make_float2(C4.z, C4.w)};
      float2 xb[2][2], xc[2][2];  // A-fragments of the dB / dC column sums
      float v[4];                 // gs(ch0), S(ch0), gs(ch1), S(ch1): sums over this thread's 4 states
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const float dl = ch ? dlv.y : dlv.x, du = ch ? duv.y : duv.x, dy = ch ? dyv.y : dyv.x;
        const float2 dl2 = make_float2(dl, dl), du2 = make_float2(du, du), dy2 = make_float2(dy, dy);
        const float2 ndu2 = make_float2(-du, -du);
        float2 gs2 = make_float2(0.f, 0.f), S2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float2 ga = __fmul2_rn(dl2, A2[ch][r]);
          const float2 a = make_float2(ex2_approx(ga.x), ex2_approx(ga.y));
          float2 hm, hc;  // hm = a_t * h_{t-1}, hc = h_t
          if (odd) {
            hm = __fmul2_rn(a, hprev[ch][r]);
            hc = __ffma2_rn(du2, Bp[r], hm);
          } else {
            hc = hprev[ch][r];
            hm = __ffma2_rn(ndu2, Bp[r], hc);
          }
          const float2 dh = __ffma2_rn(Cp[r], dy2, dhc[ch][r]);
          if constexpr (kF32) {
            xc[ch][r] = __fmul2_rn(dy2, hc);
            xb[ch][r] = __fmul2_rn(dh, du2);
          } else {
            xc[ch][r] = hc, xb[ch][r] = dh;
          }
          const float2 gg = __fmul2_rn(dh, hm);  // dL/d(delta*A) for these two states
          gs2 = __ffma2_rn(gg, A2[ch][r], gs2);
          dAacc[ch][r] = __ffma2_rn(gg, dl2, dAacc[ch][r]);
          S2 = __ffma2_rn(dh, Bp[r], S2);
          dhc[ch][r] = __fmul2_rn(a, dh);
        }
        v[2 * ch] = gs2.x + gs2.y, v[2 * ch + 1] = S2.x + S2.y;
      }
      // sums over the octet's channels on the tensor pipe, into the accumulator column of this timestep
      if constexpr (kF32) {
        const uint32_t oh = onehot ? 0x3f800000u : 0u;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            const float2 x0 = w ? xc[0][r] : xb[0][r], x1 = w ? xc[1][r] : xb[1][r];
            const uint32_t h0 = __float_as_uint(x0.x) & 0xffffe000u, h1 = __float_as_uint(x0.y) & 0xffffe000u;
            const uint32_t h2 = __float_as_uint(x1.x) & 0xffffe000u, h3 = __float_as_uint(x1.y) & 0xffffe000u;
            float(&acc)[4] = w ? accC[r] : accB[r];
            mma_tf32(acc, h0, h1, h2, h3, oh, oh);
            mma_tf32(acc, __float_as_uint(x0.x - __uint_as_float(h0)), __float_as_uint(x0.y - __uint_as_float(h1)),
                     __float_as_uint(x1.x - __uint_as_float(h2)), __float_as_uint(x1.y - __uint_as_float(h3)), oh, oh);
          }
        }
      } else {
        const uint32_t bu0 = onehot ? __float_as_uint(duv.x) : 0u, bu1 = onehot ? __float_as_uint(duv.y) : 0u;
        const uint32_t by0 = onehot ? __float_as_uint(dyv.x) : 0u, by1 = onehot ? __float_as_uint(dyv.y) : 0u;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          mma_tf32(accB[r], __float_as_uint(xb[0][r].x), __float_as_uint(xb[0][r].y), __float_as_uint(xb[1][r].x),
                   __float_as_uint(xb[1][r].y), bu0, bu1);
          mma_tf32(accC[r], __float_as_uint(xc[0][r].x), __float_as_uint(xc[0][r].y), __float_as_uint(xc[1][r].x),
                   __float_as_uint(xc[1][r].y), by0, by1);
        }
      }
      // sums over states: 4 values over the 8 g-lanes (lane bits 2..4); the last stage is the caller's
      reduce_scatter_step<4, 4>(v, lane & 4);
      reduce_scatter_step<2, 8>(v, lane & 8);
      return v[0];  // lane bit 2: channel, lane bit 3: gs / S
    };

    // ---- prologue: recompute the last chunk (iteration 0) on its own -------------------------------------------
    float2 hR[2][2], hnext[2][2];
    load_ckpt(nck - 1, hR);
    load_ckpt(nck - 2, hnext);  // start state of iteration 1's chunk (zero when nck == 1 or for chunk 0)
    bar_sync(1, bar_count);
#pragma unroll 2
    for (int i = 0; i < K; ++i) {
      const uint32_t a = in0 + 2 * i * kRowC, bq = bc0 + 2 * i * kRowS;
      const float2 dle = lds64(a + Lay::w_dl), due = lds64(a + Lay::w_du);
      const float2 dlo = lds64(a + Lay::w_dl + kRowC), duo = lds64(a + Lay::w_du + kRowC);
      const float4 Be = lds128(bq), Bo = lds128(bq + kRowS);
      rec_step(dle, due, Be, hR);
      hs_store(hs0 + i * Lay::hs_slot, hR);  // iteration 0: even step 2i -> slot i
      rec_step(dlo, duo, Bo, hR);
    }

    int vslot = 0;                 // operand slot of iteration `it`
    float pend1 = 0.f, pend0 = 0.f;  // state sums of the previous step pair, one butterfly stage short
    uint32_t pend_addr = 0;          // where they go (0: nothing pending)
    auto finish_pending = [&]() {
      const float f1 = pend1 + __shfl_xor_sync(0xffffffffu, pend1, 16);
      const float f0 = pend0 + __shfl_xor_sync(0xffffffffu, pend0, 16);
      if (st_writer && pend_addr) {
        sts32(pend_addr + kRowC, f1);
        sts32(pend_addr, f0);
      }
    };
    for (int it = 0; it < nck; ++it) {
      const int rslot = vslot + 1 == kFIn ? 0 : vslot + 1;
      const bool has_next = it + 1 < nck;
      const bool par = it & 1;  // even step j of this chunk lives in hs slot (par ? K-1-j : j)
      const uint32_t oslot = (it & 1) * Lay::out_bytes;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) hR[ch][0] = hnext[ch][0], hR[ch][1] = hnext[ch][1];
      // operands of the next chunk are prepared — and (also on the last iteration, where the helpers arrive without
      // a pre-pass) the post-pass that read this iteration's result slot two chunks ago is over
      bar_sync(1 + rslot, bar_count);
      if (has_next) load_ckpt(nck - 3 - it, hnext);  // its latency hides behind this whole iteration
      // running addresses: reverse sweep walks down from step pair K-1, recompute walks up from step pair 0
      uint32_t va = in0 + vslot * Lay::in_bytes + (kCK - 2) * kRowC;   // step 2tp of the swept chunk
      uint32_t vb = bc0 + vslot * Lay::in_bytes + (kCK - 2) * kRowS;
      uint32_t ra = in0 + rslot * Lay::in_bytes;                       // step 2i of the recomputed chunk
      uint32_t rb = bc0 + rslot * Lay::in_bytes;
      uint32_t ha = hs0 + (par ? 0 : (K - 1) * Lay::hs_slot);          // hs slot read (and then rewritten)
      const int hstep = par ? Lay::hs_slot : -Lay::hs_slot;
      uint32_t sa = st0 + oslot + (kCK - 2) * kRowC;                   // state sums of step 2tp
      uint32_t pa = pb0 + oslot + kCK * RS * 4;                        // dB|dC flush base: 8 steps down per flush
      // delta of the first iteration's four steps (the MUFU inputs), prefetched one iteration ahead from here on
      float2 dl1 = lds64(va + Lay::w_dl + kRowC), dl0 = lds64(va + Lay::w_dl);
      float2 dle = lds64(ra + Lay::w_dl), dlo = lds64(ra + Lay::w_dl + kRowC);

      auto body = [&](const int i, auto with_rec) {
        constexpr bool kRec = decltype(with_rec)::value;
        const int tp = K - 1 - i;
        // loads of this iteration
        const float4 he0 = lds128(ha), he1 = lds128(ha + kHch);
        const float2 du1 = lds64(va + Lay::w_du + kRowC), dy1 = lds64(va + Lay::w_dy + kRowC);
        const float4 B1 = lds128(vb + kRowS), C1 = lds128(vb + kCoff + kRowS);
        const float2 du0 = lds64(va + Lay::w_du), dy0 = lds64(va + Lay::w_dy);
        const float4 B0 = lds128(vb), C0 = lds128(vb + kCoff);
        float2 due, duo;
        float4 Be, Bo;
        if constexpr (kRec) {
          due = lds64(ra + Lay::w_du), duo = lds64(ra + Lay::w_du + kRowC);
          Be = lds128(rb), Bo = lds128(rb + kRowS);
        }
        finish_pending();  // last butterfly stage + store of the previous pair's state sums
        const float2 he[2][2] = {{make_float2(he0.x, he0.y), make_float2(he0.z, he0.w)},
                                 {make_float2(he1.x, he1.y), make_float2(he1.z, he1.w)}};
        const int col = (2 * tp) & 7;
        if constexpr (kRec) rec_step(dle, due, Be, hR);
        pend1 = rev_compute(dl1, du1, dy1, B1, C1, he, true, g == col + 1);
        if constexpr (kRec) {
          hs_store(ha, hR);  // the slot just read: next chunk's even step 2i
          rec_step(dlo, duo, Bo, hR);
        }
        pend0 = rev_compute(dl0, du0, dy0, B0, C0, he, false, g == col);
        pend_addr = sa;
        // advance; prefetch the next iteration's delta (past the chunk on the last iteration: in-bounds, unused)
        va -= 2 * kRowC, vb -= 2 * kRowS, sa -= 2 * kRowC, ha += hstep;
        dl1 = lds64(va + Lay::w_dl + kRowC), dl0 = lds64(va + Lay::w_dl);
        if constexpr (kRec) {
          ra += 2 * kRowC, rb += 2 * kRowS;
          dle = lds64(ra + Lay::w_dl), dlo = lds64(ra + Lay::w_dl + kRowC);
        }
      };
      // four step pairs (8 timesteps) fill the accumulator columns; then flush dB | dC: this thread holds steps
      // 8*grp' + 2*t4 (+1), states n0..n0+3
      auto flush = [&]() {
        pa -= 8 * RS * 4;
        sts128(pa, make_float4(accB[0][0], accB[0][2], accB[1][0], accB[1][2]));
        sts128(pa + RS * 4, make_float4(accB[0][1], accB[0][3], accB[1][1], accB[1][3]));
        sts128(pa + NPT * 4, make_float4(accC[0][0], accC[0][2], accC[1][0], accC[1][2]));
        sts128(pa + (RS + NPT) * 4, make_float4(accC[0][1], accC[0][3], accC[1][1], accC[1][3]));
#pragma unroll
        for (int w = 0; w < 2; ++w)
#pragma unroll
          for (int k = 0; k < 4; ++k) accB[w][k] = accC[w][k] = 0.f;
      };
#pragma unroll 1
      for (int grp = 0; grp < K / 4; ++grp) {
        if (has_next) {
#pragma unroll 1
          for (int j = 0; j < 4; ++j) body(4 * grp + j, std::true_type{});
        } else {
#pragma unroll 1
          for (int j = 0; j < 4; ++j) body(4 * grp + j, std::false_type{});
        }
        flush();
      }
      finish_pending();  // the results of this iteration must be complete before DONE
      pend_addr = 0;
      bar_arrive(4 + (it & 1), bar_count);  // chunk swept
      vslot = rslot;
    }
    // ---- epilogue: dA partial of this batch element -----------------------------------------------------------
#pragma unroll
    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (d + ch < p.D)
          p.ws_dA[((int64_t)b * p.N + n0 + j) * p.D + d + ch] = (j & 1) ? dAacc[ch][j / 2].y : dAacc[ch][j / 2].x;
    return;
  }

  // =========================================== HELPER WARPS ===============================================
  if constexpr (NW == 8) setmaxnreg_dec<88>();
  const int ht = tid - nscan_threads;  // 0..127
  const bool do_softplus = p.flags & MAMBA_FLAG_DELTA_SOFTPLUS;
  const int my_t = ht >> 3, my_c = 4 * (ht & 7);  // this thread's (timestep, 4 channels) of every chunk
  const bool my_row = my_t < kCK;                  // with 8-step chunks half of the helper threads only load
  const int my_nv = dvalid - my_c;                 // valid channels of this thread's quad (may be <= 0)
  float bias4[4], D4[4], dD_acc[4], db_acc[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int dd = d0 + my_c + e;
    bias4[e] = ((p.flags & MAMBA_FLAG_HAS_DELTA_BIAS) && dd < p.D) ? p.dbias[dd] : 0.f;
    D4[e] = ((p.flags & MAMBA_FLAG_HAS_D) && dd < p.D) ? p.Dv[dd] : 0.f;
    dD_acc[e] = 0.f, db_acc[e] = 0.f;
  }
  const T* gu = static_cast<const T*>(p.u) + (int64_t)b * p.u_bs + d0;
  const T* gdl = static_cast<const T*>(p.delta) + (int64_t)b * p.delta_bs + d0;
  const T* gz = has_z ? static_cast<const T*>(p.z) + (int64_t)b * p.z_bs + d0 : nullptr;
  const T* gyp = has_z ? static_cast<const T*>(p.ypre) + (int64_t)b * p.ypre_bs + d0 : nullptr;
  const T* gdo = static_cast<const T*>(p.dout) + (int64_t)b * p.dout_bs + d0;
  const T* gB = static_cast<const T*>(p.Bm) + (int64_t)b * p.B_bs;
  const T* gC = static_cast<const T*>(p.Cm) + (int64_t)b * p.C_bs;

  auto issue_loads = [&](int it) {
    if (it < nck) {
      const int c = nck - 1 - it;
      unsigned char* base = raw_base + (size_t)(it & 1) * lay.raw_bytes;
      const int t0 = c * kCK;
      const int rv = min(kCK, p.L - t0);
      load_tile_async<T>(reinterpret_cast<T*>(base + lay.raw_u), kBD, gu + (int64_t)t0 * p.u_ls, p.u_ls, kCK, rv, dvalid,
                         p.vec_u, ht, kBHelperThreads);
      load_tile_async<T>(reinterpret_cast<T*>(base + lay.raw_dl), kBD, gdl + (int64_t)t0 * p.delta_ls, p.delta_ls, kCK, rv,
                         dvalid, p.vec_delta, ht, kBHelperThreads);
      load_tile_async<T>(reinterpret_cast<T*>(base + lay.raw_do), kBD, gdo + (int64_t)t0 * p.dout_ls, p.dout_ls, kCK, rv,
                         dvalid, p.vec_dout, ht, kBHelperThreads);
      if (has_z)
        load_tile_async<T>(reinterpret_cast<T*>(base + lay.raw_z), kBD, gz + (int64_t)t0 * p.z_ls, p.z_ls, kCK, rv, dvalid,
                           p.vec_z, ht, kBHelperThreads);
      load_tile_async<T>(reinterpret_cast<T*>(base + lay.raw_B), NPT, gB + (int64_t)t0 * p.B_ls, p.B_ls, kCK, rv, p.N,
                         p.vec_B, ht, kBHelperThreads);
      load_tile_async<T>(reinterpret_cast<T*>(base + lay.raw_C), NPT, gC + (int64_t)t0 * p.C_ls, p.C_ls, kCK, rv, p.N,
                         p.vec_C, ht, kBHelperThreads);
    }
    cp_async_commit();
  };

  auto pre_pass = [&](int it, int islot) {
    const int c = nck - 1 - it;
    unsigned char* rbase = raw_base + (size_t)(it & 1) * lay.raw_bytes;
    unsigned char* wbase = in_base + (size_t)islot * lay.in_bytes;
    const int rv = min(kCK, p.L - c * kCK);
    const int o = my_t * kBD + my_c;
    if (my_row) {
      float dl[4], uu[4], dy[4], du[4];
      V4<T>::ld(reinterpret_cast<const T*>(rbase + lay.raw_dl) + o, dl);
      V4<T>::ld(reinterpret_cast<const T*>(rbase + lay.raw_u) + o, uu);
      V4<T>::ld(reinterpret_cast<const T*>(rbase + lay.raw_do) + o, dy);
      if (has_z) {
        float zz[4];
        V4<T>::ld(reinterpret_cast<const T*>(rbase + lay.raw_z) + o, zz);
#pragma unroll
        for (int e = 0; e < 4; ++e) dy[e] *= silu_fast(zz[e]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = dl[e] + bias4[e];
        if (do_softplus) v = softplus_fast(v);
        if (my_t >= rv) v = 0.f, dy[e] = 0.f;  // padded timestep: contributes nothing
        dl[e] = v, du[e] = v * uu[e];
      }
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(wbase + lay.w_dl) + o) = make_float4(dl[0], dl[1], dl[2], dl[3]);
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(wbase + lay.w_du) + o) = make_float4(du[0], du[1], du[2], du[3]);
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(wbase + lay.w_dy) + o) = make_float4(dy[0], dy[1], dy[2], dy[3]);
    }
    // B / C of the chunk as fp32 (converted or copied: the raw slot is recycled as soon as this pass is over)
    float* Bf = reinterpret_cast<float*>(wbase + lay.w_Bf);
    float* Cf = reinterpret_cast<float*>(wbase + lay.w_Cf);
    if constexpr (kF32) {
      const float4* sB = reinterpret_cast<const float4*>(rbase + lay.raw_B);
      const float4* sC = reinterpret_cast<const float4*>(rbase + lay.raw_C);
      for (int i = ht; i < kCK * NPT / 4; i += kBHelperThreads) {
        reinterpret_cast<float4*>(Bf)[i] = sB[i];
        reinterpret_cast<float4*>(Cf)[i] = sC[i];
      }
    } else {
      const T* sB = reinterpret_cast<const T*>(rbase + lay.raw_B);
      const T* sC = reinterpret_cast<const T*>(rbase + lay.raw_C);
      for (int i = ht; i < kCK * NPT / 8; i += kBHelperThreads) {
        cvt8_bf16_f32(sB + 8 * i, Bf + 8 * i);
        cvt8_bf16_f32(sC + 8 * i, Cf + 8 * i);
      }
    }
  };

  auto store4 = [&](void* base, int64_t bs, int64_t ls, int64_t tg, bool vec, const float (&v)[4]) {
    T* o = static_cast<T*>(base) + (int64_t)b * bs + tg * ls + d0 + my_c;
    if (vec && my_nv >= 4) {
      V4<T>::st_global(o, v);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (e < my_nv) IO<T>::st(o + e, v[e]);
    }
  };

  auto post_pass = [&](int it, int islot) {
    const int c = nck - 1 - it;
    unsigned char* wbase = in_base + (size_t)islot * lay.in_bytes;
    unsigned char* obase = out_base + (size_t)(it & 1) * lay.out_bytes;
    const int t0 = c * kCK;
    const int rv = min(kCK, p.L - t0);
    if (my_row && my_t < rv) {
      const int o = my_t * kBD + my_c;
      const int64_t tg = t0 + my_t;
      // per-(t, d) operands again, through L2 (they were streamed two chunks ago)
      float uu[4], raw[4], go[4];
      ld4_global<T>(gu + tg * p.u_ls + my_c, p.vec_u, my_nv, uu);
      ld4_global<T>(gdl + tg * p.delta_ls + my_c, p.vec_delta, my_nv, raw);
      ld4_global<T>(gdo + tg * p.dout_ls + my_c, p.vec_dout, my_nv, go);
      float zz[4], yp[4];
      if (has_z) {
        ld4_global<T>(gz + tg * p.z_ls + my_c, p.vec_z, my_nv, zz);
        ld4_global<T>(gyp + tg * p.ypre_ls + my_c, p.vec_ypre, my_nv, yp);
      }
      const float* pg = reinterpret_cast<const float*>(obase + lay.o_pg) + o;
      const float* pS = reinterpret_cast<const float*>(obase + lay.o_pS) + o;
      float4 g4 = *reinterpret_cast<const float4*>(pg), S4 = *reinterpret_cast<const float4*>(pS);
#pragma unroll
      for (int w = 1; w < NSH; ++w) {
        const float4 a = *reinterpret_cast<const float4*>(pg + w * kCK * kBD);
        const float4 s4 = *reinterpret_cast<const float4*>(pS + w * kCK * kBD);
        g4.x += a.x, g4.y += a.y, g4.z += a.z, g4.w += a.w;
        S4.x += s4.x, S4.y += s4.y, S4.z += s4.z, S4.w += s4.w;
      }
      const float gs[4] = {g4.x, g4.y, g4.z, g4.w}, S[4] = {S4.x, S4.y, S4.z, S4.w};
      const float4 dl4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(wbase + lay.w_dl) + o);
      const float4 dy4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(wbase + lay.w_dy) + o);
      const float dl[4] = {dl4.x, dl4.y, dl4.z, dl4.w}, dy[4] = {dy4.x, dy4.y, dy4.z, dy4.w};
      float ddl[4], du[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = fmaf(uu[e], S[e], gs[e] * kLn2);
        if (do_softplus) {
          const float x = raw[e] + bias4[e];
          v *= x > 20.f ? 1.f : rcp_approx(1.f + ex2_approx(-x * kLog2e));  // softplus' = sigmoid
        }
        ddl[e] = v;
        du[e] = fmaf(dl[e], S[e], dy[e] * D4[e]);
        dD_acc[e] = fmaf(dy[e], uu[e], dD_acc[e]);
        db_acc[e] += v;
      }
      store4(p.du, p.du_bs, p.du_ls, tg, p.vec_du, du);
      store4(p.ddelta, p.ddelta_bs, p.ddelta_ls, tg, p.vec_ddelta, ddl);
      if (has_z) {
        float dz[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float sg = rcp_approx(1.f + ex2_approx(-zz[e] * kLog2e));
          dz[e] = go[e] * yp[e] * sg * fmaf(zz[e], 1.f - sg, 1.f);
        }
        store4(p.dz, p.dz_bs, p.dz_ls, tg, p.vec_dz, dz);
      }
    }
    // per-CTA partial of dB / dC for this chunk -> workspace [B][ntiles][L][N]: sum of the four octet partials
    {
      float* wsB = p.ws_dB + (((int64_t)b * p.ntiles + tile) * p.L + t0) * p.N;
      float* wsC = p.ws_dC + (((int64_t)b * p.ntiles + tile) * p.L + t0) * p.N;
      const float* bc = reinterpret_cast<const float*>(obase + lay.o_bc);
      constexpr int Q = NPT / 4;
      for (int i = ht; i < rv * Q; i += kBHelperThreads) {
        const int t = i / Q, q = i - t * Q;
        const float* src = bc + t * RS + 4 * q;
        float4 sB = *reinterpret_cast<const float4*>(src), sC = *reinterpret_cast<const float4*>(src + NPT);
#pragma unroll
        for (int oc = 1; oc < 4; ++oc) {
          const float4 xb = *reinterpret_cast<const float4*>(src + oc * kCK * RS);
          const float4 xc = *reinterpret_cast<const float4*>(src + oc * kCK * RS + NPT);
          sB.x += xb.x, sB.y += xb.y, sB.z += xb.z, sB.w += xb.w;
          sC.x += xc.x, sC.y += xc.y, sC.z += xc.z, sC.w += xc.w;
        }
        if (p.fixed_acc) {
          // one accumulator per (b, t, n): 8-byte slots in the same workspace, [B][L][N]
          long long* aB = reinterpret_cast<long long*>(p.ws_dB) + ((int64_t)b * p.L + t0) * p.N + 4 * (int64_t)i;
          long long* aC = reinterpret_cast<long long*>(p.ws_dC) + ((int64_t)b * p.L + t0) * p.N + 4 * (int64_t)i;
          red_add_fixed(reinterpret_cast<float*>(aB + 0), sB.x), red_add_fixed(reinterpret_cast<float*>(aB + 1), sB.y);
          red_add_fixed(reinterpret_cast<float*>(aB + 2), sB.z), red_add_fixed(reinterpret_cast<float*>(aB + 3), sB.w);
          red_add_fixed(reinterpret_cast<float*>(aC + 0), sC.x), red_add_fixed(reinterpret_cast<float*>(aC + 1), sC.y);
          red_add_fixed(reinterpret_cast<float*>(aC + 2), sC.z), red_add_fixed(reinterpret_cast<float*>(aC + 3), sC.w);
        } else {
          reinterpret_cast<float4*>(wsB)[i] = sB;
          reinterpret_cast<float4*>(wsC)[i] = sC;
        }
      }
    }
  };

  // Helper iteration j: pre-pass of chunk j (operand slot j % 3), then — once the scan warps have finished
  // iteration j-2 — its post-pass.  The raw slot is recycled right after the pre-pass (loads run 2 chunks ahead).
  issue_loads(0);
  issue_loads(1);
  int islot = 0, pslot = 1;  // operand slots of chunk j and of chunk j-2
  for (int j = 0; j < nck + 2; ++j) {
    if (j < nck) {
      cp_async_wait<1>();
      bar_sync(6, kBHelperThreads);
      pre_pass(j, islot);
      bar_arrive(1 + islot, bar_count);  // READY
      bar_sync(6, kBHelperThreads);      // every helper is done with raw slot j & 1
      issue_loads(j + 2);
    } else if (j == nck) {
      bar_arrive(1 + islot, bar_count);  // releases the scan warps' last iteration (its result slot is free)
    }
    if (j >= 2) {
      bar_sync(4 + (j & 1), bar_count);  // DONE of iteration j-2
      post_pass(j - 2, pslot);
    }
    islot = islot + 1 == kFIn ? 0 : islot + 1;
    pslot = pslot + 1 == kFIn ? 0 : pslot + 1;
  }

  // ---- dD / d_bias: sum this CTA's timestep-rows in shared memory, one partial per (b, d) ----------------------
  bar_sync(6, kBHelperThreads);
  float* fin = reinterpret_cast<float*>(in_base);  // [2][16][32]; the operand slots are free now
  constexpr int kRows = kBHelperThreads / 8;
  *reinterpret_cast<float4*>(fin + my_t * kBD + my_c) = make_float4(dD_acc[0], dD_acc[1], dD_acc[2], dD_acc[3]);
  *reinterpret_cast<float4*>(fin + (kRows + my_t) * kBD + my_c) = make_float4(db_acc[0], db_acc[1], db_acc[2], db_acc[3]);
  bar_sync(6, kBHelperThreads);
  if (ht < kBD && d0 + ht < p.D) {
    float sD = 0.f, sb = 0.f;
    for (int r = 0; r < kRows; ++r) {
      sD += fin[r * kBD + ht];
      sb += fin[(kRows + r) * kBD + ht];
    }
    p.ws_dD[(int64_t)b * p.D + d0 + ht] = sD;
    p.ws_db[(int64_t)b * p.D + d0 + ht] = sb;
  }
}

template <typename T, int NW, int kCK>
static int launch_fused(const ScanBwdParams& p, cudaStream_t stream) {
  const size_t smem = (size_t)FusedLayout<T, NW, kCK>::total;
  if (smem > 227 * 1024) return set_error(MAMBA_ESIZE, "scan_bwd (fused): needs %zu B of shared memory", smem);
  auto kern = scan_bwd_fused_kernel<T, NW, kCK>;
  static thread_local SmemConfig cfg;
  if (int rc = ensure_dynamic_smem(kern, smem, cfg, "scan_bwd (fused)")) return rc;
  dim3 grid(p.ntiles, p.B);
  kern<<<grid, (NW + kBHelperWarps) * 32, smem, stream>>>(p);
  count_launch();
  return check_launch("scan_bwd_fused");
}

template <typename T>
int launch_scan_bwd_fused(const ScanBwdParams& p, cudaStream_t stream) {
  if (p.N == 64) return p.ck == 8 ? launch_fused<T, 8, 8>(p, stream) : launch_fused<T, 8, 16>(p, stream);
  if (p.N == 32) return p.ck == 8 ? launch_fused<T, 4, 8>(p, stream) : launch_fused<T, 4, 16>(p, stream);
  return set_error(MAMBA_EINVAL, "scan_bwd (fused): d_state must be 32 or 64 (got %d)", p.N);
}
template int launch_scan_bwd_fused<float>(const ScanBwdParams&, cudaStream_t);
template int launch_scan_bwd_fused<__nv_bfloat16>(const ScanBwdParams&, cudaStream_t);

}  // namespace mb
