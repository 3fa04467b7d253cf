// scan_fwd.cu — fused selective-scan forward for sm_100a (warp-specialised, MUFU-paced).
//
// Replaces MambaBlock.selective_scan (reference models/mamba/__pycache__/simple_mamba.cpython-311.pyc
// @L310-333) fused with softplus(dt) (@L276), the D skip (@L331) and the z gate (@L241).
//
// The op is one ex2 per (b, t, d, n) on the MUFU pipe (16 lanes/clk/SM measured, tools/microbench.cu) plus four
// fp32 multiply-adds; at d_state 64 that is ~6x more time than its HBM traffic, at d_state 16 about 1.4x.  The
// kernel is organised so that the MUFU pipe is the only thing that is ever busy:
//   * one CTA owns (batch b, 32 consecutive channels) for the whole sequence; lane <-> channel, so every global
//     access of u / delta / z / out is a coalesced row and no shuffle is ever needed.
//   * SCAN warps (one per slice of NPER states) do nothing but the recurrence.  Per timestep: one 8-byte LDS of
//     (delta, delta*u), broadcast 16-byte LDS of B and C, and per PAIR of states {FMUL2, 2 x MUFU.EX2, FMUL2,
//     FFMA2, FFMA2} using Blackwell's packed fp32x2 pipe ops (3 issue slots per state instead of 5), then one
//     STS of the partial <h, C>.  h[NPER] stays in registers for all L steps; operands of step t+1 are fetched
//     before step t is computed.
//   * HELPER warps (4) run everything else through shared-memory rings: cp.async (LDGSTS, 16 B) loads RR stages
//     ahead (enough bytes in flight to cover HBM latency with one CTA per SM), the pre-pass (softplus(delta +
//     bias), delta*u), and the post-pass (sum of the per-warp partials, + D*u, * silu(z),
//     128-bit coalesced store).  Each helper thread owns (timestep, 4 consecutive channels) of a stage.
//   * producer/consumer hand-off uses named barriers (bar.arrive / bar.sync), two per work slot.
//   * every 16 timesteps the scan warps checkpoint h to HBM ([B, nck, N, D], coalesced) for the backward.
//     deltaA / deltaB_u ([B, L, D, N] in the reference) never exist in memory.
#include "scan_fwd.cuh"

namespace mb {

constexpr int kDT = 32;   // channels per CTA (= warp width)
constexpr int kTS = 16;   // timesteps per stage (= checkpoint interval)
constexpr int kHelperWarps = 4;
constexpr int kHelperThreads = kHelperWarps * 32;
constexpr int kMaxScanWarps = 16;
constexpr int kMaxBCVec = 4;  // precomputed B/C cp.async slots per helper thread (fast path)


// 4 consecutive shared-memory elements as a float4 (bf16 widened in registers)
__device__ __forceinline__ float4 ld4_as_f32(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4_as_f32(const __nv_bfloat16* p) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                     __uint_as_float(v.y & 0xffff0000u));
}

// Shared-memory layout.  RAW ring slot (cp.async targets, element type T):
//     u[TS][32], delta[TS][32], z[TS][32], B[TS][NPT], C[TS][NPT]
// WORK slot (two of them):  dlu[TS][32] float2 = (softplus(delta+bias), that * u),  ypart[NS][TS][32] fp32.
template <typename T>
struct FwdLayout {
  int raw_u, raw_dl, raw_z, raw_B, raw_C, raw_bytes;
  int w_dlu, w_y, w_Bf, w_Cf, work_bytes;
  __host__ __device__ FwdLayout(int NS, int NPT) {
    int o = 0;
    raw_u = o, o += kTS * kDT * (int)sizeof(T);
    raw_dl = o, o += kTS * kDT * (int)sizeof(T);
    raw_z = o, o += kTS * kDT * (int)sizeof(T);
    raw_B = o, o += kTS * NPT * (int)sizeof(T);
    raw_C = o, o += kTS * NPT * (int)sizeof(T);
    raw_bytes = (o + 127) & ~127;
    o = 0;
    w_dlu = o, o += kTS * kDT * 8;
    w_y = o, o += NS * kTS * kDT * 4;
    w_Bf = o, o += (sizeof(T) == 4 ? 0 : kTS * NPT * 4);  // bf16 I/O: B / C widened by the helpers' pre-pass
    w_Cf = o, o += (sizeof(T) == 4 ? 0 : kTS * NPT * 4);
    work_bytes = (o + 127) & ~127;
  }
};

template <typename T, int NPER, int RR, int CKI>
__global__ void __launch_bounds__((kMaxScanWarps + kHelperWarps) * 32) scan_fwd_kernel(const ScanFwdParams p) {
  static_assert(kTS % CKI == 0, "a stage holds whole checkpoint chunks");
  static_assert(NPER % 4 == 0, "states per thread come in float4 groups");
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NS = p.NS, NPT = p.NPT;
  const int b = blockIdx.y, d0 = blockIdx.x * kDT;
  const FwdLayout<T> lay(NS, NPT);
  unsigned char* const raw_base = smem;
  unsigned char* const work_base = smem + (size_t)RR * lay.raw_bytes;
  const int nst = (p.L + kTS - 1) / kTS;
  const int nscan_threads = NS * 32;
  const int bar_count = nscan_threads + kHelperThreads;
  // barrier ids: 1,2 = READY[work slot]; 3,4 = DONE[work slot]; 5 = helpers only

  if (warp < NS) {
    // ======================================= SCAN WARPS ===============================================
    const int s = warp, d = d0 + lane;
    float2 A2[NPER / 2], h[NPER / 2];
#pragma unroll
    for (int j = 0; j < NPER; ++j) {
      const int n = s * NPER + j;
      const bool ok = (d < p.D) && (n < p.N);
      const float a = ok ? load_A(p.A, (int64_t)d * p.N + n, p.flags) * kLog2e : 0.f;
      const float h0 = (ok && p.h_init) ? p.h_init[((int64_t)b * p.D + d) * p.N + n] : 0.f;
      if (j & 1) A2[j / 2].y = a, h[j / 2].y = h0;
      else A2[j / 2].x = a, h[j / 2].x = h0;
    }
    int rslot = 0;
    for (int c = 0; c < nst; ++c) {
      const int ws = c & 1;
      unsigned char* wbase = work_base + (size_t)ws * lay.work_bytes;
      unsigned char* rbase = raw_base + (size_t)rslot * lay.raw_bytes;
      const float2* dlu = reinterpret_cast<const float2*>(wbase + lay.w_dlu) + lane;
      // B / C as fp32: the raw cp.async tile (fp32 I/O) or the copy widened by the helpers' pre-pass (bf16 I/O;
      // widening in the scan warps costs 16 issue slots per thread and step, a third of this loop)
      const float* Bf = (sizeof(T) == 4 ? reinterpret_cast<const float*>(rbase + lay.raw_B)
                                        : reinterpret_cast<const float*>(wbase + lay.w_Bf)) + s * NPER;
      const float* Cf = (sizeof(T) == 4 ? reinterpret_cast<const float*>(rbase + lay.raw_C)
                                        : reinterpret_cast<const float*>(wbase + lay.w_Cf)) + s * NPER;
      float* yp = reinterpret_cast<float*>(wbase + lay.w_y) + (s * kTS) * kDT + lane;
      bar_sync(1 + ws, bar_count);  // stage c prepared
      // Operands of timestep t+1 are fetched BEFORE timestep t is computed and stored: ptxas will not move a
      // shared load above an earlier shared store it cannot disambiguate.
      // Software pipeline: (delta, delta*u) is fetched two steps ahead, B / C one step ahead, and the decay
      // a_{t+1} = exp2(delta_{t+1} * A) is formed while step t's multiply-adds run, so that the MUFU results are
      // never waited for.
      float2 dd_cur, dd_nxt, dd_n2;
      float4 Bc[NPER / 4], Cc[NPER / 4], Bn[NPER / 4], Cn[NPER / 4];
      float2 a_cur[NPER / 2], a_nxt[NPER / 2];
      auto fetch_bc = [&](int t, float4 (&Bv)[NPER / 4], float4 (&Cv)[NPER / 4]) {
#pragma unroll
        for (int q = 0; q < NPER / 4; ++q) {
          Bv[q] = ld4_as_f32(Bf + t * NPT + 4 * q);
          Cv[q] = ld4_as_f32(Cf + t * NPT + 4 * q);
        }
      };
      auto decay = [&](const float2 dd, float2 (&a)[NPER / 2]) {
        const float2 dl2 = make_float2(dd.x, dd.x);
#pragma unroll
        for (int k = 0; k < NPER / 2; ++k) {
          const float2 gk = __fmul2_rn(dl2, A2[k]);
          a[k] = make_float2(ex2_approx(gk.x), ex2_approx(gk.y));
        }
      };
      dd_cur = dlu[0];
      dd_nxt = dlu[(kTS > 1 ? 1 : 0) * kDT];
      fetch_bc(0, Bc, Cc);
      decay(dd_cur, a_cur);
#pragma unroll
      for (int t = 0; t < kTS; ++t) {
        if (t + 2 < kTS) dd_n2 = dlu[(t + 2) * kDT];
        if (t + 1 < kTS) {
          fetch_bc(t + 1, Bn, Cn);
          decay(dd_nxt, a_nxt);
        }
        const float2 du2 = make_float2(dd_cur.y, dd_cur.y);
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < NPER / 4; ++q) {
          const float2 B01 = make_float2(Bc[q].x, Bc[q].y), B23 = make_float2(Bc[q].z, Bc[q].w);
          const float2 C01 = make_float2(Cc[q].x, Cc[q].y), C23 = make_float2(Cc[q].z, Cc[q].w);
          h[2 * q] = __ffma2_rn(a_cur[2 * q], h[2 * q], __fmul2_rn(du2, B01));
          h[2 * q + 1] = __ffma2_rn(a_cur[2 * q + 1], h[2 * q + 1], __fmul2_rn(du2, B23));
          acc = __ffma2_rn(h[2 * q], C01, acc);
          acc = __ffma2_rn(h[2 * q + 1], C23, acc);
        }
        sts_f32(yp + t * kDT, acc.x + acc.y);
        dd_cur = dd_nxt, dd_nxt = dd_n2;
#pragma unroll
        for (int q = 0; q < NPER / 4; ++q) Bc[q] = Bn[q], Cc[q] = Cn[q];
#pragma unroll
        for (int k = 0; k < NPER / 2; ++k) a_cur[k] = a_nxt[k];
        if ((t + 1) % CKI == 0) {
          // h is now the state at the start of checkpoint chunk (c*kTS + t + 1) / CKI
          const int tg_next = c * kTS + t + 1;
          if (p.ckpt != nullptr && tg_next < p.L && d < p.D) {
            // layout [B][nck][ceil(N/4)][D][4]: one 16-byte store per 4 states, 512 contiguous bytes per warp
            float4* ck = reinterpret_cast<float4*>(p.ckpt) +
                         (((int64_t)b * p.nck + tg_next / CKI) * p.N4 + s * (NPER / 4)) * p.D + d;
#pragma unroll
            for (int q = 0; q < NPER / 4; ++q)
              if (s * NPER + 4 * q < p.N) ck[(int64_t)q * p.D] = make_float4(h[2 * q].x, h[2 * q].y, h[2 * q + 1].x, h[2 * q + 1].y);
          }
        }
      }
      bar_arrive(3 + ws, bar_count);  // stage c scanned
      rslot = (rslot + 1 == RR) ? 0 : rslot + 1;
    }
    if (p.h_last != nullptr && d < p.D) {
#pragma unroll
      for (int j = 0; j < NPER; ++j) {
        const int n = s * NPER + j;
        if (n < p.N) p.h_last[((int64_t)b * p.D + d) * p.N + n] = (j & 1) ? h[j / 2].y : h[j / 2].x;
      }
    }
    return;
  }

  // ========================================= HELPER WARPS ===============================================
  const int ht = tid - nscan_threads;  // 0..127
  const int dvalid = min(kDT, p.D - d0);
  const bool has_z = p.flags & MAMBA_FLAG_HAS_Z;
  const bool do_softplus = p.flags & MAMBA_FLAG_DELTA_SOFTPLUS;
  // this thread's element of every stage: timestep ht >> 3, channels 4*(ht & 7) .. +3
  const int my_t = ht >> 3, my_c = 4 * (ht & 7);
  float bias4[4], D4[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int d = d0 + my_c + e;
    bias4[e] = ((p.flags & MAMBA_FLAG_HAS_DELTA_BIAS) && d < p.D) ? p.dbias[d] : 0.f;
    D4[e] = ((p.flags & MAMBA_FLAG_HAS_D) && d < p.D) ? p.Dv[d] : 0.f;
  }

  const T* gu = static_cast<const T*>(p.u) + (int64_t)b * p.u_bs + d0;
  const T* gdl = static_cast<const T*>(p.delta) + (int64_t)b * p.delta_bs + d0;
  const T* gz = has_z ? static_cast<const T*>(p.z) + (int64_t)b * p.z_bs + d0 : nullptr;
  const T* gB = static_cast<const T*>(p.Bm) + (int64_t)b * p.B_bs;
  const T* gC = static_cast<const T*>(p.Cm) + (int64_t)b * p.C_bs;
  T* gout = static_cast<T*>(p.out) + (int64_t)b * p.out_bs + d0;

  // ---- fast load path: full 32-channel tile, 16-byte aligned everything: one precomputed cp.async per tile ----
  constexpr int VEC = 16 / (int)sizeof(T);         // elements per 16-byte vector
  constexpr int VPR = kDT / VEC;                    // vectors per activation row (8 fp32 / 4 bf16)
  const bool fast = p.vec_u && p.vec_delta && (!has_z || p.vec_z) && p.vec_B && p.vec_C && dvalid == kDT;
  const bool act_mine = ht < kTS * VPR;             // this thread owns one vector of each activation tile
  const int a_row = ht / VPR, a_col = (ht % VPR) * VEC;
  const int bc_vpr = NPT / VEC;                     // vectors per B/C row (NPT is a multiple of 8)
  const int bc_nvec = kTS * bc_vpr;
  int bc_row[kMaxBCVec], bc_col[kMaxBCVec];
#pragma unroll
  for (int k = 0; k < kMaxBCVec; ++k) {
    const int i = ht + k * kHelperThreads;
    bc_row[k] = i / bc_vpr;
    bc_col[k] = (i - bc_row[k] * bc_vpr) * VEC;
  }
  const bool bc_fast = fast && bc_nvec <= kMaxBCVec * kHelperThreads;

  auto zero16 = [](void* dst) { *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u); };

  auto issue_loads = [&](int c, int rslot) {
    if (c < nst) {
      unsigned char* base = raw_base + (size_t)rslot * lay.raw_bytes;
      const int t0 = c * kTS;
      const int rv = min(kTS, p.L - t0);
      T* su = reinterpret_cast<T*>(base + lay.raw_u);
      T* sdl = reinterpret_cast<T*>(base + lay.raw_dl);
      T* sz = reinterpret_cast<T*>(base + lay.raw_z);
      T* sB = reinterpret_cast<T*>(base + lay.raw_B);
      T* sC = reinterpret_cast<T*>(base + lay.raw_C);
      if (fast) {
        if (act_mine) {
          const int so = a_row * kDT + a_col;
          if (a_row < rv) {
            const int64_t t = t0 + a_row;
            cp_async16(su + so, gu + t * p.u_ls + a_col);
            cp_async16(sdl + so, gdl + t * p.delta_ls + a_col);
            if (has_z) cp_async16(sz + so, gz + t * p.z_ls + a_col);
          } else {
            zero16(su + so), zero16(sdl + so);
            if (has_z) zero16(sz + so);
          }
        }
      } else {
        load_tile_async<T>(su, kDT, gu + (int64_t)t0 * p.u_ls, p.u_ls, kTS, rv, dvalid, p.vec_u, ht, kHelperThreads);
        load_tile_async<T>(sdl, kDT, gdl + (int64_t)t0 * p.delta_ls, p.delta_ls, kTS, rv, dvalid, p.vec_delta, ht,
                           kHelperThreads);
        if (has_z)
          load_tile_async<T>(sz, kDT, gz + (int64_t)t0 * p.z_ls, p.z_ls, kTS, rv, dvalid, p.vec_z, ht, kHelperThreads);
      }
      if (bc_fast) {
#pragma unroll
        for (int k = 0; k < kMaxBCVec; ++k) {
          if (ht + k * kHelperThreads < bc_nvec) {
            const int so = bc_row[k] * NPT + bc_col[k];
            if (bc_row[k] < rv && bc_col[k] + VEC <= p.N) {
              const int64_t t = t0 + bc_row[k];
              cp_async16(sB + so, gB + t * p.B_ls + bc_col[k]);
              cp_async16(sC + so, gC + t * p.C_ls + bc_col[k]);
            } else if (bc_row[k] < rv && bc_col[k] < p.N) {  // vector straddles N: guarded scalars
              const int64_t t = t0 + bc_row[k];
              for (int e = 0; e < VEC; ++e) {
                const bool ok = bc_col[k] + e < p.N;
                sB[so + e] = ok ? gB[t * p.B_ls + bc_col[k] + e] : T(0.f);
                sC[so + e] = ok ? gC[t * p.C_ls + bc_col[k] + e] : T(0.f);
              }
            } else {
              zero16(sB + so), zero16(sC + so);
            }
          }
        }
      } else {
        load_tile_async<T>(sB, NPT, gB + (int64_t)t0 * p.B_ls, p.B_ls, kTS, rv, p.N, p.vec_B, ht, kHelperThreads);
        load_tile_async<T>(sC, NPT, gC + (int64_t)t0 * p.C_ls, p.C_ls, kTS, rv, p.N, p.vec_C, ht, kHelperThreads);
      }
    }
    cp_async_commit();  // always commit so that the group arithmetic below is uniform
  };

  auto pre_pass = [&](int c, int rslot) {
    unsigned char* rbase = raw_base + (size_t)rslot * lay.raw_bytes;
    unsigned char* wbase = work_base + (size_t)(c & 1) * lay.work_bytes;
    const int rv = min(kTS, p.L - c * kTS);
    float dl[4], uu[4];
    V4<T>::ld(reinterpret_cast<const T*>(rbase + lay.raw_dl) + my_t * kDT + my_c, dl);
    V4<T>::ld(reinterpret_cast<const T*>(rbase + lay.raw_u) + my_t * kDT + my_c, uu);
    float r[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = dl[e] + bias4[e];
      if (do_softplus) v = softplus_fast(v);
      if (my_t >= rv) v = 0.f;  // padded timestep: a = 1, input 0 -> state unchanged
      r[2 * e] = v, r[2 * e + 1] = v * uu[e];
    }
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float2*>(wbase + lay.w_dlu) + my_t * kDT + my_c);
    dst[0] = make_float4(r[0], r[1], r[2], r[3]);
    dst[1] = make_float4(r[4], r[5], r[6], r[7]);
    if (sizeof(T) != 4) {
      const T* sB = reinterpret_cast<const T*>(rbase + lay.raw_B);
      const T* sC = reinterpret_cast<const T*>(rbase + lay.raw_C);
      float* Bf = reinterpret_cast<float*>(wbase + lay.w_Bf);
      float* Cf = reinterpret_cast<float*>(wbase + lay.w_Cf);
      for (int i = ht; i < kTS * NPT / 8; i += kHelperThreads) {  // NPT is a multiple of 8
        cvt8_bf16_f32(sB + 8 * i, Bf + 8 * i);
        cvt8_bf16_f32(sC + 8 * i, Cf + 8 * i);
      }
    }
  };

  auto post_pass = [&](int c, int rslot) {
    unsigned char* rbase = raw_base + (size_t)rslot * lay.raw_bytes;
    unsigned char* wbase = work_base + (size_t)(c & 1) * lay.work_bytes;
    const int t0 = c * kTS;
    if (t0 + my_t >= p.L) return;
    const float* yp = reinterpret_cast<const float*>(wbase + lay.w_y) + my_t * kDT + my_c;
    float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
    for (int w = 0; w < NS; ++w) {
      const float4 v = *reinterpret_cast<const float4*>(yp + w * kTS * kDT);
      s01 = __fadd2_rn(s01, make_float2(v.x, v.y));
      s23 = __fadd2_rn(s23, make_float2(v.z, v.w));
    }
    float y[4] = {s01.x, s01.y, s23.x, s23.y};
    float uu[4];
    V4<T>::ld(reinterpret_cast<const T*>(rbase + lay.raw_u) + my_t * kDT + my_c, uu);
#pragma unroll
    for (int e = 0; e < 4; ++e) y[e] = fmaf(D4[e], uu[e], y[e]);
    if (p.ypre != nullptr) {  // pre-gate output, kept for the backward's dz
      T* yo = static_cast<T*>(p.ypre) + (int64_t)b * p.ypre_bs + (int64_t)(t0 + my_t) * p.ypre_ls + d0 + my_c;
      if (p.vec_ypre && my_c + 4 <= dvalid) {
        V4<T>::st_global(yo, y);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (my_c + e < dvalid) IO<T>::st(yo + e, y[e]);
      }
    }
    if (has_z) {
      float zz[4];
      V4<T>::ld(reinterpret_cast<const T*>(rbase + lay.raw_z) + my_t * kDT + my_c, zz);
#pragma unroll
      for (int e = 0; e < 4; ++e) y[e] *= silu_fast(zz[e]);
    }
    T* o = gout + (int64_t)(t0 + my_t) * p.out_ls + my_c;
    if (p.vec_out && my_c + 4 <= dvalid) {
      V4<T>::st_global(o, y);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (my_c + e < dvalid) IO<T>::st(o + e, y[e]);
    }
  };

  // prologue: RR - 1 stages in flight
  for (int c = 0; c < RR - 1; ++c) issue_loads(c, c);
  int rslot = 0;                 // raw slot of stage c
  int pslot = RR - 1;            // raw slot of stage c - 1 (the one that is refilled at the end of iteration c)
  for (int c = 0; c < nst; ++c) {
    cp_async_wait<RR - 2>();       // this thread's loads of stage c have landed
    bar_sync(5, kHelperThreads);   // ... and every helper's (also: every helper is past post_pass(c - 2))
    pre_pass(c, rslot);
    bar_arrive(1 + (c & 1), bar_count);  // READY[work slot]
    if (c >= 1) {
      bar_sync(3 + ((c - 1) & 1), bar_count);  // DONE[work slot of c - 1]
      post_pass(c - 1, pslot);
    }
    bar_sync(5, kHelperThreads);   // every helper has finished reading stage c-1's raw tiles
    issue_loads(c + RR - 1, pslot);
    pslot = rslot;
    rslot = (rslot + 1 == RR) ? 0 : rslot + 1;
  }
  bar_sync(3 + ((nst - 1) & 1), bar_count);
  post_pass(nst - 1, pslot);
}

template <typename T>
static size_t fwd_smem_bytes(int NS, int NPT, int RR) {
  const FwdLayout<T> lay(NS, NPT);
  return (size_t)RR * lay.raw_bytes + (size_t)2 * lay.work_bytes;
}

template <typename T, int NPER, int RR, int CKI>
static int launch_fwd(const ScanFwdParams& p, cudaStream_t stream) {
  const size_t smem = fwd_smem_bytes<T>(p.NS, p.NPT, RR);
  if (smem > 227 * 1024) return set_error(MAMBA_ESIZE, "scan_fwd: d_state %d needs %zu B of shared memory", p.N, smem);
  auto kern = scan_fwd_kernel<T, NPER, RR, CKI>;
  static thread_local SmemConfig cfg;  // per instantiation and device: raise the dynamic-smem limit once per size
  if (int rc = ensure_dynamic_smem(kern, smem, cfg, "scan_fwd")) return rc;
  dim3 grid(ceil_div(p.D, kDT), p.B);
  kern<<<grid, (p.NS + kHelperWarps) * 32, smem, stream>>>(p);
  count_launch();
  return check_launch("scan_fwd");
}

template <typename T, int NPER, int CKI>
static int dispatch_ring(const ScanFwdParams& p, cudaStream_t stream) {
  // deep ring when a stage is short (few states): bytes in flight must cover HBM latency with one CTA per SM
  if (p.N <= 32 && fwd_smem_bytes<T>(p.NS, p.NPT, 8) <= 100 * 1024) return launch_fwd<T, NPER, 8, CKI>(p, stream);
  return launch_fwd<T, NPER, 4, CKI>(p, stream);
}

template <typename T, int NPER>
static int dispatch_cki(const ScanFwdParams& p, cudaStream_t stream) {
  return p.cki == 8 ? dispatch_ring<T, NPER, 8>(p, stream) : dispatch_ring<T, NPER, 16>(p, stream);
}

template <typename T>
static int dispatch_nper(ScanFwdParams& p, int nper, cudaStream_t stream) {
  p.NS = ceil_div(p.N, nper);
  const int np = p.NS * nper;
  p.NPT = (np + 7) & ~7;
  if (p.NS > kMaxScanWarps) return set_error(MAMBA_ESIZE, "scan_fwd: d_state %d too large for %d states/thread", p.N, nper);
  switch (nper) {
    case 4: return dispatch_cki<T, 4>(p, stream);
    case 8: return dispatch_cki<T, 8>(p, stream);
    case 16: return dispatch_cki<T, 16>(p, stream);
  }
  return set_error(MAMBA_EINVAL, "scan_fwd: variant must be 0, 4, 8 or 16 (got %d)", nper);
}

static bool vec_ok(const void* ptr, int64_t bs, int64_t ls, size_t elt) {
  return aligned16(ptr) && (bs * elt) % 16 == 0 && (ls * elt) % 16 == 0;
}

}  // namespace mb

extern "C" size_t mamba_scan_ckpt_elems(int batch, int seqlen, int dim, int dstate, int chunk) {
  if (batch <= 0 || seqlen <= 0 || dim <= 0 || dstate <= 0 || chunk <= 0) return 0;
  return (size_t)batch * mb::ceil_div(seqlen, chunk) * (4 * mb::ceil_div(dstate, 4)) * dim;
}

extern "C" int mamba_scan_fwd(const MambaScanFwdArgs* a, void* stream) {
  using namespace mb;
  if (!a || a->struct_size != (int32_t)sizeof(MambaScanFwdArgs))
    return set_error(MAMBA_EINVAL, "scan_fwd: bad args pointer or struct_size");
  if (a->batch <= 0 || a->seqlen <= 0 || a->dim <= 0 || a->dstate <= 0)
    return set_error(MAMBA_EINVAL, "scan_fwd: batch/seqlen/dim/dstate must be positive (got %d/%d/%d/%d)", a->batch,
                     a->seqlen, a->dim, a->dstate);
  if (a->batch > 65535) return set_error(MAMBA_ESIZE, "scan_fwd: batch %d above 65535", a->batch);
  if (!a->u || !a->delta || !a->A || !a->B || !a->C || !a->out)
    return set_error(MAMBA_EINVAL, "scan_fwd: null u/delta/A/B/C/out");
  if ((a->flags & MAMBA_FLAG_HAS_Z) && !a->z) return set_error(MAMBA_EINVAL, "scan_fwd: HAS_Z but z == NULL");
  if ((a->flags & MAMBA_FLAG_HAS_D) && !a->D) return set_error(MAMBA_EINVAL, "scan_fwd: HAS_D but D == NULL");
  if ((a->flags & MAMBA_FLAG_HAS_DELTA_BIAS) && !a->delta_bias)
    return set_error(MAMBA_EINVAL, "scan_fwd: HAS_DELTA_BIAS but delta_bias == NULL");
  if (a->dtype != MAMBA_F32 && a->dtype != MAMBA_BF16) return set_error(MAMBA_EDTYPE, "scan_fwd: dtype %d", a->dtype);
  if (a->ckpt && !aligned16(a->ckpt)) return set_error(MAMBA_EALIGN, "scan_fwd: ckpt must be 16-byte aligned");
  if (a->ckpt && a->chunk != 8 && a->chunk != 16)
    return set_error(MAMBA_EINVAL, "scan_fwd: chunk must be 8 or 16 (got %d)", a->chunk);

  ScanFwdParams p{};
  p.B = a->batch, p.L = a->seqlen, p.D = a->dim, p.N = a->dstate, p.flags = a->flags;
  p.N4 = (p.N + 3) / 4;
  p.cki = a->ckpt ? a->chunk : 16;
  p.nck = ceil_div(p.L, p.cki);
  p.u = a->u, p.delta = a->delta, p.Bm = a->B, p.Cm = a->C, p.z = a->z, p.out = a->out;
  p.u_bs = a->u_bs, p.u_ls = a->u_ls, p.delta_bs = a->delta_bs, p.delta_ls = a->delta_ls;
  p.B_bs = a->B_bs, p.B_ls = a->B_ls, p.C_bs = a->C_bs, p.C_ls = a->C_ls;
  p.z_bs = a->z_bs, p.z_ls = a->z_ls, p.out_bs = a->out_bs, p.out_ls = a->out_ls;
  p.A = a->A, p.Dv = a->D, p.dbias = a->delta_bias, p.h_init = a->h_init;
  p.ckpt = a->ckpt, p.h_last = a->h_last;
  const size_t elt = a->dtype == MAMBA_F32 ? 4 : 2;
  // 16-byte cp.async needs the 32-channel tile start (d0 * elt, a multiple of 64 B) on an aligned base.
  p.vec_u = vec_ok(a->u, a->u_bs, a->u_ls, elt);
  p.vec_delta = vec_ok(a->delta, a->delta_bs, a->delta_ls, elt);
  p.vec_z = a->z ? vec_ok(a->z, a->z_bs, a->z_ls, elt) : 0;
  p.vec_B = vec_ok(a->B, a->B_bs, a->B_ls, elt);
  p.vec_C = vec_ok(a->C, a->C_bs, a->C_ls, elt);
  p.vec_out = vec_ok(a->out, a->out_bs, a->out_ls, elt);
  p.ypre = a->y_pre, p.ypre_bs = a->y_pre_bs, p.ypre_ls = a->y_pre_ls;
  p.vec_ypre = a->y_pre ? vec_ok(a->y_pre, a->y_pre_bs, a->y_pre_ls, elt) : 0;

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // variant 100 + tune: the TMA-staged kernel (scan_fwd_tma.cu).  Automatic choice: d_state <= 32 takes it when the
  // tensors can be described by tensor maps (config 5: 384 us against 529 us for this file's kernel at B=2, L=8192,
  // N=16 fp32); at d_state 64 the two measure the same and this file's kernel stays the default.
  if (a->variant >= 100 || (a->variant == 0 && p.N <= 32)) {
    const int rc = launch_scan_fwd_tma(p, a->dtype, a->variant >= 100 ? a->variant - 100 : 0, st);
    if (rc != kTmaNotEligible) return rc;
    if (a->variant >= 100)
      return set_error(MAMBA_EALIGN, "scan_fwd: variant %d (TMA) needs 16-byte aligned bases/strides and d_state <= 128", a->variant);
  }
  int nper = a->variant;
  if (nper == 0) nper = p.N >= 64 ? 8 : 4;
  while (nper < 16 && ceil_div(p.N, nper) > kMaxScanWarps) nper *= 2;
  return a->dtype == MAMBA_F32 ? dispatch_nper<float>(p, nper, st) : dispatch_nper<__nv_bfloat16>(p, nper, st);
}
