// scan_bwd.cu — fused selective-scan backward for sm_100a (warp-specialised).
//
// The reference has no explicit backward: torch autograd differentiates the python loop of
// MambaBlock.selective_scan (models/mamba/__pycache__/simple_mamba.cpython-311.pyc @L310-333), saving the
// [B, L, D, N] tensors deltaA / deltaB_u and every per-step state.  Here the forward is recomputed chunk by chunk
// (16 timesteps) from the checkpoints written by scan_fwd and the adjoint recurrence
//     dh_t = C_t * dy_t + a_{t+1} * dh_{t+1}
// runs in registers; nothing of size B*L*D*N touches HBM.
//
// Per (b, t, d, n) the work is 2 MUFU.EX2 (a_t in the recompute and again in the reverse sweep — keeping a_t in
// shared memory instead would saturate the shared-memory pipe) and 13 fp32 multiply-adds, i.e. the kernel is
// paced by the MUFU pipe with the FMA pipe and the issue slots close behind; its HBM traffic is ~10x smaller
// than that.  Organisation (same producer/consumer scheme as scan_fwd.cu):
//   * one CTA owns (batch b, 32 channels); chunks are walked from the end of the sequence to the start.
//   * SCAN warps: lane <-> channel, each thread owns NPER states of its channel (warps tile the state axis), so
//     the per-channel operands are 4-byte shared loads (one wavefront per warp) and B / C are warp broadcasts.
//     Recompute: h_t of the EVEN steps of the chunk -> shared memory (thread-private float4; odd steps are
//     rebuilt in the reverse sweep).  Reverse sweep: all arithmetic in packed fp32x2 (FFMA2/FMUL2) on state
//     pairs; a_t*h_{t-1} is h_t - delta*u*B on even steps and a_t*h_even on odd ones, so h_{t-1} is never
//     re-read.  Sums over states (d_delta, d_u) finish inside the thread (+ one per-warp partial in shared
//     memory); sums over the 32 channels (dB, dC) are a reduce-scatter across the lanes (2*NPER shuffles for
//     2*NPER values instead of 5 per value).
//   * HELPER warps (4): cp.async loads one chunk ahead, pre-pass (softplus(delta+bias), delta*u,
//     dy = dout*silu(z)), post-pass (finish d_delta (softplus'), d_u (+D*dy), dz from the saved pre-gate output,
//     accumulate dD / d_bias, 128-bit stores, per-CTA dB/dC partial -> workspace).
//   * a finalize kernel reduces the per-CTA / per-batch partials in a fixed order (deterministic, no atomics).
#include <cstdlib>

#include "scan_bwd.cuh"

namespace mb {

// Shared-memory layout.
//   RAW ring slot (cp.async targets, element type T): u, delta, z, dout, ypre [CK][32];  B, C [CK][NPT]
//   WORK slot (two): wdl, wdu, wdy fp32 [CK][32];  (bf16 I/O only) Bf, Cf fp32 [CK][NPT];
//                    pg, pS [NW][CK][32] (per-warp partial sums over states);  redB, redC [CK][NPT]
//   hs: float4 [CK/2][NPER/4][scan threads]  (h_t of the even steps of the chunk, 4 states per float4, thread-private)
template <typename T, int kCK>
struct BwdLayout {
  int raw_u, raw_dl, raw_z, raw_do, raw_yp, raw_B, raw_C, raw_bytes;
  int w_dl, w_du, w_dy, w_Bf, w_Cf, w_pg, w_pS, w_rB, w_rC, work_bytes;
  int hs_bytes;
  __host__ __device__ BwdLayout(int NW, int NPT, int NPER) {
    int o = 0;
    raw_u = o, o += kCK * kBD * (int)sizeof(T);
    raw_dl = o, o += kCK * kBD * (int)sizeof(T);
    raw_z = o, o += kCK * kBD * (int)sizeof(T);
    raw_do = o, o += kCK * kBD * (int)sizeof(T);
    raw_yp = o, o += kCK * kBD * (int)sizeof(T);
    raw_B = o, o += kCK * NPT * (int)sizeof(T);
    raw_C = o, o += kCK * NPT * (int)sizeof(T);
    raw_bytes = (o + 127) & ~127;
    o = 0;
    w_dl = o, o += kCK * kBD * 4;
    w_du = o, o += kCK * kBD * 4;
    w_dy = o, o += kCK * kBD * 4;
    w_Bf = o, o += (sizeof(T) == 4 ? 0 : kCK * NPT * 4);
    w_Cf = o, o += (sizeof(T) == 4 ? 0 : kCK * NPT * 4);
    w_pg = o, o += NW * kCK * kBD * 4;
    w_pS = o, o += NW * kCK * kBD * 4;
    w_rB = o, o += kCK * NPT * 4;
    w_rC = o, o += kCK * NPT * 4;
    work_bytes = (o + 127) & ~127;
    hs_bytes = (kCK / 2) * NW * 32 * (NPER >= 4 ? (NPER / 4) * 16 : 8);  // even steps only: one float4 per 4 states and thread (NPER 2: one float2)
  }
};

// HT = helper teams (of kBHelperWarps warps each).  With two teams the chunks alternate between them and every team
// has TWO raw and TWO work slots of its own, so that a team's pre-pass of its next chunk is finished before the scan
// warps are done with its current one: the per-(t, d) work then has two scan-chunks of time per chunk instead of one.
// That is for small d_state (the scan warps' share per chunk shrinks with d_state, the helpers' does not) on grids
// that cannot fill the machine (otherwise two resident CTAs per SM do the same job, see the launch bounds).
template <typename T, int NPER, int NW, int kCK, int HT>
__global__ void __launch_bounds__((NW + HT * kBHelperWarps) * 32, (NPER == 4 && NW <= 4 && HT == 1) ? 2 : 1)
    scan_bwd_kernel(const ScanBwdParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NPT = ((NW * NPER + 7) / 8) * 8;  // padded d_state of the shared tiles
  const int b = blockIdx.y, tile = blockIdx.x;
  const int d0 = tile * kBD;
  const int dvalid = min(kBD, p.D - d0);
  const BwdLayout<T, kCK> lay(NW, NPT, NPER);
  unsigned char* const raw_base = smem;
  constexpr int NSLOT = 2 * HT;   // raw slots = work slots; iteration i uses slot i % NSLOT (team i % HT)
  static_assert(kBRing == 2, "slot = iteration % NSLOT assumes the two-deep raw ring per team");
  unsigned char* const work_base = smem + (size_t)NSLOT * lay.raw_bytes;
  float4* const hs = reinterpret_cast<float4*>(work_base + (size_t)NSLOT * lay.work_bytes);
  const int nck = p.nck;
  constexpr int nscan_threads = NW * 32;
  constexpr int bar_count = nscan_threads + kBHelperThreads;   // the scan warps + ONE helper team
  const bool has_z = p.flags & MAMBA_FLAG_HAS_Z;
  // barrier ids: 1 + slot = READY[slot]; 1 + NSLOT + slot = DONE[slot]; 1 + 2*NSLOT + team = that team's helpers only;
  // 15 = all helpers (epilogue).  Iteration i handles chunk nck-1-i.
  constexpr int kBarReady = 1, kBarDone = 1 + NSLOT, kBarTeam = 1 + 2 * NSLOT, kBarHelpers = 15;

  if (warp < NW) {
    // ========================================= SCAN WARPS ===============================================
    // lane <-> channel (32 channels), this warp's slice of NPER states in registers, packed in pairs along n.
    // Per-channel operands (delta, delta*u, dy) are then ONE 4-byte shared load per lane (a single wavefront per
    // warp instead of four for a 16-byte load), B / C are full-warp broadcasts, sums over states stay inside the
    // thread, and only the sums over channels (dB, dC) cross lanes.
    constexpr int NQ = NPER / 4;   // float4 groups of states per thread (0 when the thread owns one state pair)
    constexpr int NQ1 = NQ > 0 ? NQ : 1;
    constexpr bool kPair = NPER == 2;
    constexpr int NP = NPER / 2;   // state pairs per thread
    const int n0 = warp * NPER;    // first state of this thread
    const int d = d0 + lane;
    float2 A2[NP], dAacc[NP], dhc[NP];
#pragma unroll
    for (int j = 0; j < NPER; ++j) {
      const int n = n0 + j;
      const float a = (d < p.D && n < p.N) ? load_A(p.A, (int64_t)d * p.N + n, p.flags) * kLog2e : 0.f;
      if (j & 1) A2[j / 2].y = a;
      else A2[j / 2].x = a;
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) dAacc[q] = dhc[q] = make_float2(0.f, 0.f);
    // checkpoint layout [B][nck][ceil(N/4)][D][4] (scan_fwd.cu): 16 bytes per lane, 512 contiguous bytes per warp
    const int N4 = (p.N + 3) >> 2;
    auto load_ckpt = [&](int c, float2 (&h)[NP]) {
      if constexpr (kPair) {   // half of a 16-byte checkpoint group
        float2 v = make_float2(0.f, 0.f);
        if (c > 0 && d < p.D && n0 < p.N)
          v = __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float4*>(p.ckpt) + (((int64_t)b * nck + c) * N4 + (n0 >> 2)) * p.D + d) + ((n0 >> 1) & 1));
        h[0] = v;
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c > 0 && d < p.D && n0 + 4 * q < p.N)
          v = __ldg(reinterpret_cast<const float4*>(p.ckpt) + (((int64_t)b * nck + c) * N4 + (n0 >> 2) + q) * p.D + d);
        h[2 * q] = make_float2(v.x, v.y), h[2 * q + 1] = make_float2(v.z, v.w);
      }
    };
    float2 hnext[NP];
    load_ckpt(nck - 1, hnext);
    float4* const hst = hs + tid;  // + (tp * NQ + q) * nscan_threads
    float2* const hst2 = reinterpret_cast<float2*>(hs) + tid;  // NPER 2: + tp * nscan_threads

    for (int it = 0; it < nck; ++it) {
      const int c = nck - 1 - it;
      const int ws = it & (NSLOT - 1);
      unsigned char* wbase = work_base + (size_t)ws * lay.work_bytes;
      unsigned char* rbase = raw_base + (size_t)ws * lay.raw_bytes;
      const float* wdl = reinterpret_cast<const float*>(wbase + lay.w_dl) + lane;
      const float* wdu = reinterpret_cast<const float*>(wbase + lay.w_du) + lane;
      const float* wdy = reinterpret_cast<const float*>(wbase + lay.w_dy) + lane;
      const float* Bf = (sizeof(T) == 4 ? reinterpret_cast<const float*>(rbase + lay.raw_B)
                                        : reinterpret_cast<const float*>(wbase + lay.w_Bf)) + n0;
      const float* Cf = (sizeof(T) == 4 ? reinterpret_cast<const float*>(rbase + lay.raw_C)
                                        : reinterpret_cast<const float*>(wbase + lay.w_Cf)) + n0;
      float* pg = reinterpret_cast<float*>(wbase + lay.w_pg) + (warp * kCK) * kBD + lane;
      float* pS = reinterpret_cast<float*>(wbase + lay.w_pS) + (warp * kCK) * kBD + lane;
      float* redB = reinterpret_cast<float*>(wbase + lay.w_rB) + n0;
      float* redC = reinterpret_cast<float*>(wbase + lay.w_rC) + n0;

      float2 h[NP];
#pragma unroll
      for (int q = 0; q < NP; ++q) h[q] = hnext[q];
      bar_sync(kBarReady + ws, bar_count);  // chunk c prepared

      // ---- forward recompute: h_t of the EVEN steps of the chunk -> shared memory --------------------------
#pragma unroll 4
      for (int t = 0; t < kCK; ++t) {
        const float dl = wdl[t * kBD], du = wdu[t * kBD];
        const float2 dl2 = make_float2(dl, dl), du2 = make_float2(du, du);
        if constexpr (kPair) {
          const float2 b2 = *reinterpret_cast<const float2*>(Bf + t * NPT);
          const float2 g0 = __fmul2_rn(dl2, A2[0]);
          const float2 a0 = make_float2(ex2_approx(g0.x), ex2_approx(g0.y));
          h[0] = __ffma2_rn(a0, h[0], __fmul2_rn(du2, b2));
          if ((t & 1) == 0) hst2[(t >> 1) * nscan_threads] = h[0];
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const float4 b4 = *reinterpret_cast<const float4*>(Bf + t * NPT + 4 * q);
          const float2 g0 = __fmul2_rn(dl2, A2[2 * q]), g1 = __fmul2_rn(dl2, A2[2 * q + 1]);
          const float2 a0 = make_float2(ex2_approx(g0.x), ex2_approx(g0.y));
          const float2 a1 = make_float2(ex2_approx(g1.x), ex2_approx(g1.y));
          h[2 * q] = __ffma2_rn(a0, h[2 * q], __fmul2_rn(du2, make_float2(b4.x, b4.y)));
          h[2 * q + 1] = __ffma2_rn(a1, h[2 * q + 1], __fmul2_rn(du2, make_float2(b4.z, b4.w)));
          if ((t & 1) == 0)  // odd steps are rebuilt in the reverse sweep from the even one before them
            hst[((t >> 1) * NQ + q) * nscan_threads] = make_float4(h[2 * q].x, h[2 * q].y, h[2 * q + 1].x, h[2 * q + 1].y);
        }
      }
      // prefetch the next chunk's start state: its latency hides behind the reverse sweep
      if (c > 0) load_ckpt(c - 1, hnext);

      // ---- reverse sweep: adjoint recurrence ------------------------------------------------------------------
      // Steps are taken in (odd, even) pairs: only h of the even step is in shared memory; the odd step's state
      // is h_odd = a_odd * h_even + delta*u*B, whose first term is exactly the a_t * h_{t-1} the gradient needs.
      // Each step is split into load / compute / store so that the shared loads of the second step of a pair sit
      // BEFORE the shared stores of the first in program order (ptxas does not move a load above a store it cannot
      // disambiguate) and can overlap the first step's arithmetic.
      struct RevOps {
        float dl, du, dy;
        float4 B[NQ1], C[NQ1];   // (NPER 2: the pair sits in .x, .y)
      };
      struct RevOut {
        float g, S, red;
      };
      auto rev_load = [&](const int t, RevOps& o) {
        o.dl = wdl[t * kBD], o.du = wdu[t * kBD], o.dy = wdy[t * kBD];
        if constexpr (kPair) {
          const float2 b2 = *reinterpret_cast<const float2*>(Bf + t * NPT), c2 = *reinterpret_cast<const float2*>(Cf + t * NPT);
          o.B[0] = make_float4(b2.x, b2.y, 0.f, 0.f), o.C[0] = make_float4(c2.x, c2.y, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          o.B[q] = *reinterpret_cast<const float4*>(Bf + t * NPT + 4 * q);
          o.C[q] = *reinterpret_cast<const float4*>(Cf + t * NPT + 4 * q);
        }
      };
      auto rev_compute = [&](const RevOps& o, const float2 (&hprev)[NP], const bool odd, RevOut& out) {
        const float2 dl2 = make_float2(o.dl, o.dl), du2 = make_float2(o.du, o.du), dy2 = make_float2(o.dy, o.dy);
        const float2 ndu2 = make_float2(-o.du, -o.du);
        float2 gs2 = make_float2(0.f, 0.f), S2 = make_float2(0.f, 0.f);
        float red[2 * NPER];  // dB_0..dB_{NPER-1}, dC_0..dC_{NPER-1} of this lane's channel
#pragma unroll
        for (int q = 0; q < NQ1; ++q) {
          const float2 Bp[2] = {make_float2(o.B[q].x, o.B[q].y), make_float2(o.B[q].z, o.B[q].w)};
          const float2 Cp[2] = {make_float2(o.C[q].x, o.C[q].y), make_float2(o.C[q].z, o.C[q].w)};
#pragma unroll
          for (int r = 0; r < (kPair ? 1 : 2); ++r) {
            const int k = 2 * q + r;  // state pair index
            const float2 ga = __fmul2_rn(dl2, A2[k]);
            const float2 a = make_float2(ex2_approx(ga.x), ex2_approx(ga.y));
            float2 hm, hc;  // hm = a_t * h_{t-1}, hc = h_t
            if (odd) {
              hm = __fmul2_rn(a, hprev[k]);
              hc = __ffma2_rn(du2, Bp[r], hm);
            } else {
              hc = hprev[k];
              hm = __ffma2_rn(ndu2, Bp[r], hc);
            }
            const float2 dh = __ffma2_rn(Cp[r], dy2, dhc[k]);
            const float2 dc = __fmul2_rn(dy2, hc);
            const float2 db = __fmul2_rn(dh, du2);
            const float2 g = __fmul2_rn(dh, hm);  // dL/d(delta*A) for these two states
            gs2 = __ffma2_rn(g, A2[k], gs2);
            dAacc[k] = __ffma2_rn(g, dl2, dAacc[k]);
            S2 = __ffma2_rn(dh, Bp[r], S2);
            dhc[k] = __fmul2_rn(a, dh);
            red[2 * k] = db.x, red[2 * k + 1] = db.y;
            red[NPER + 2 * k] = dc.x, red[NPER + 2 * k + 1] = dc.y;
          }
        }
        // sums over states are complete inside the thread
        out.g = gs2.x + gs2.y, out.S = S2.x + S2.y;
        // sums over the 32 channels of the warp: reduce-scatter of the 2*NPER values, then plain butterflies
        constexpr int V = 2 * NPER;
        reduce_scatter_step<V, 16>(red, lane & 16);
        reduce_scatter_step<(V >= 2 ? V / 2 : 1), 8>(red, lane & 8);
        reduce_scatter_step<(V >= 4 ? V / 4 : 1), 4>(red, lane & 4);
        reduce_scatter_step<(V >= 8 ? V / 8 : 1), 2>(red, lane & 2);
        reduce_scatter_step<(V >= 16 ? V / 16 : 1), 1>(red, lane & 1);
        out.red = red[0];
      };
      // V = 32: every lane holds one value (index = lane); V = 16: index = lane >> 1; V = 8: lane >> 2; V = 4: lane >> 3
      constexpr int SH = 2 * NPER >= 32 ? 0 : (2 * NPER == 16 ? 1 : (2 * NPER == 8 ? 2 : 3));
      const int ridx = lane >> SH;
      const bool rwriter = (lane & ((1 << SH) - 1)) == 0;
      float* const rdst = (ridx >= NPER ? redC : redB) + (ridx % NPER);
      auto rev_store = [&](const int t, const RevOut& out) {
        pg[t * kBD] = out.g;
        pS[t * kBD] = out.S;
        if (rwriter) rdst[t * NPT] = out.red;
      };
#pragma unroll 1
      for (int tp = kCK / 2 - 1; tp >= 0; --tp) {
        float2 he[NP];
        if constexpr (kPair) he[0] = hst2[tp * nscan_threads];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const float4 h4 = hst[(tp * NQ + q) * nscan_threads];
          he[2 * q] = make_float2(h4.x, h4.y), he[2 * q + 1] = make_float2(h4.z, h4.w);
        }
        RevOps o1, o0;
        RevOut r1, r0;
        rev_load(2 * tp + 1, o1);
        rev_load(2 * tp, o0);
        rev_compute(o1, he, true, r1);
        rev_compute(o0, he, false, r0);
        rev_store(2 * tp + 1, r1);
        rev_store(2 * tp, r0);
      }
      bar_arrive(kBarDone + ws, bar_count);  // chunk c swept
    }
    // ---- epilogue: dA partial of this batch element -----------------------------------------------------------
    if (d < p.D) {
#pragma unroll
      for (int j = 0; j < NPER; ++j) {
        const int n = n0 + j;
        if (n < p.N) p.ws_dA[((int64_t)b * p.N + n) * p.D + d] = (j & 1) ? dAacc[j / 2].y : dAacc[j / 2].x;
      }
    }
    return;
  }

  // =========================================== HELPER WARPS ===============================================
  const int team = (tid - nscan_threads) / kBHelperThreads;   // 0 .. HT-1
  const int ht = (tid - nscan_threads) % kBHelperThreads;     // 0..127 inside the team
  const bool do_softplus = p.flags & MAMBA_FLAG_DELTA_SOFTPLUS;
  const int my_t = ht >> 3, my_c = 4 * (ht & 7);  // this thread's (timestep, 4 channels) of every chunk
  const bool my_row = my_t < kCK;                  // with 8-step chunks half of the helper threads only load
  float bias4[4], D4[4], dD_acc[4], db_acc[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int d = d0 + my_c + e;
    bias4[e] = ((p.flags & MAMBA_FLAG_HAS_DELTA_BIAS) && d < p.D) ? p.dbias[d] : 0.f;
    D4[e] = ((p.flags & MAMBA_FLAG_HAS_D) && d < p.D) ? p.Dv[d] : 0.f;
    dD_acc[e] = 0.f, db_acc[e] = 0.f;
  }
  const T* gu = static_cast<const T*>(p.u) + (int64_t)b * p.u_bs + d0;
  const T* gdl = static_cast<const T*>(p.delta) + (int64_t)b * p.delta_bs + d0;
  const T* gz = has_z ? static_cast<const T*>(p.z) + (int64_t)b * p.z_bs + d0 : nullptr;
  const T* gyp = has_z ? static_cast<const T*>(p.ypre) + (int64_t)b * p.ypre_bs + d0 : nullptr;
  const T* gdo = static_cast<const T*>(p.dout) + (int64_t)b * p.dout_bs + d0;
  const T* gB = static_cast<const T*>(p.Bm) + (int64_t)b * p.B_bs;
  const T* gC = static_cast<const T*>(p.Cm) + (int64_t)b * p.C_bs;

  auto issue_loads = [&](int it, int rslot) {
    if (it < nck) {
      const int c = nck - 1 - it;
      unsigned char* base = raw_base + (size_t)rslot * lay.raw_bytes;
      const int t0 = c * kCK;
      const int rv = min(kCK, p.L - t0);
      load_tile_async<T>(reinterpret_cast<T*>(base + lay.raw_u), kBD, gu + (int64_t)t0 * p.u_ls, p.u_ls, kCK, rv, dvalid,
                         p.vec_u, ht, kBHelperThreads);
      load_tile_async<T>(reinterpret_cast<T*>(base + lay.raw_dl), kBD, gdl + (int64_t)t0 * p.delta_ls, p.delta_ls, kCK, rv,
                         dvalid, p.vec_delta, ht, kBHelperThreads);
      load_tile_async<T>(reinterpret_cast<T*>(base + lay.raw_do), kBD, gdo + (int64_t)t0 * p.dout_ls, p.dout_ls, kCK, rv,
                         dvalid, p.vec_dout, ht, kBHelperThreads);
      if (has_z) {
        load_tile_async<T>(reinterpret_cast<T*>(base + lay.raw_z), kBD, gz + (int64_t)t0 * p.z_ls, p.z_ls, kCK, rv, dvalid,
                           p.vec_z, ht, kBHelperThreads);
        load_tile_async<T>(reinterpret_cast<T*>(base + lay.raw_yp), kBD, gyp + (int64_t)t0 * p.ypre_ls, p.ypre_ls, kCK, rv,
                           dvalid, p.vec_ypre, ht, kBHelperThreads);
      }
      load_tile_async<T>(reinterpret_cast<T*>(base + lay.raw_B), NPT, gB + (int64_t)t0 * p.B_ls, p.B_ls, kCK, rv, p.N,
                         p.vec_B, ht, kBHelperThreads);
      load_tile_async<T>(reinterpret_cast<T*>(base + lay.raw_C), NPT, gC + (int64_t)t0 * p.C_ls, p.C_ls, kCK, rv, p.N,
                         p.vec_C, ht, kBHelperThreads);
    }
    cp_async_commit();
  };

  auto pre_pass = [&](int it, int rslot) {
    const int c = nck - 1 - it;
    unsigned char* rbase = raw_base + (size_t)rslot * lay.raw_bytes;
    unsigned char* wbase = work_base + (size_t)(it & (NSLOT - 1)) * lay.work_bytes;
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
    if (sizeof(T) != 4) {
      const T* sB = reinterpret_cast<const T*>(rbase + lay.raw_B);
      const T* sC = reinterpret_cast<const T*>(rbase + lay.raw_C);
      float* Bf = reinterpret_cast<float*>(wbase + lay.w_Bf);
      float* Cf = reinterpret_cast<float*>(wbase + lay.w_Cf);
      for (int i = ht; i < kCK * NPT / 8; i += kBHelperThreads) {  // NPT is a multiple of 8
        cvt8_bf16_f32(sB + 8 * i, Bf + 8 * i);
        cvt8_bf16_f32(sC + 8 * i, Cf + 8 * i);
      }
    }
  };

  auto store4 = [&](void* base, int64_t bs, int64_t ls, int64_t tg, bool vec, const float (&v)[4]) {
    T* o = static_cast<T*>(base) + (int64_t)b * bs + tg * ls + d0 + my_c;
    if (vec && my_c + 4 <= dvalid) {
      V4<T>::st_global(o, v);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (my_c + e < dvalid) IO<T>::st(o + e, v[e]);
    }
  };

  auto post_pass = [&](int it, int rslot) {
    const int c = nck - 1 - it;
    unsigned char* rbase = raw_base + (size_t)rslot * lay.raw_bytes;
    unsigned char* wbase = work_base + (size_t)(it & (NSLOT - 1)) * lay.work_bytes;
    const int t0 = c * kCK;
    const int rv = min(kCK, p.L - t0);
    if (my_row && my_t < rv) {
      const int o = my_t * kBD + my_c;
      const float* pg = reinterpret_cast<const float*>(wbase + lay.w_pg) + o;
      const float* pS = reinterpret_cast<const float*>(wbase + lay.w_pS) + o;
      float2 g01 = make_float2(0.f, 0.f), g23 = g01, S01 = g01, S23 = g01;
      for (int w = 0; w < NW; ++w) {
        const float4 a = *reinterpret_cast<const float4*>(pg + w * kCK * kBD);
        const float4 s4 = *reinterpret_cast<const float4*>(pS + w * kCK * kBD);
        g01 = __fadd2_rn(g01, make_float2(a.x, a.y)), g23 = __fadd2_rn(g23, make_float2(a.z, a.w));
        S01 = __fadd2_rn(S01, make_float2(s4.x, s4.y)), S23 = __fadd2_rn(S23, make_float2(s4.z, s4.w));
      }
      const float g[4] = {g01.x, g01.y, g23.x, g23.y}, S[4] = {S01.x, S01.y, S23.x, S23.y};
      float uu[4], raw[4], go[4];
      V4<T>::ld(reinterpret_cast<const T*>(rbase + lay.raw_u) + o, uu);
      V4<T>::ld(reinterpret_cast<const T*>(rbase + lay.raw_dl) + o, raw);
      V4<T>::ld(reinterpret_cast<const T*>(rbase + lay.raw_do) + o, go);
      const float4 dl4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(wbase + lay.w_dl) + o);
      const float4 dy4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(wbase + lay.w_dy) + o);
      const float dl[4] = {dl4.x, dl4.y, dl4.z, dl4.w}, dy[4] = {dy4.x, dy4.y, dy4.z, dy4.w};
      float ddl[4], du[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = fmaf(uu[e], S[e], g[e] * kLn2);
        if (do_softplus) {
          const float x = raw[e] + bias4[e];
          v *= x > 20.f ? 1.f : rcp_approx(1.f + ex2_approx(-x * kLog2e));  // softplus' = sigmoid
        }
        ddl[e] = v;
        du[e] = fmaf(dl[e], S[e], dy[e] * D4[e]);
        dD_acc[e] = fmaf(dy[e], uu[e], dD_acc[e]);
        db_acc[e] += v;
      }
      store4(p.du, p.du_bs, p.du_ls, t0 + my_t, p.vec_du, du);
      store4(p.ddelta, p.ddelta_bs, p.ddelta_ls, t0 + my_t, p.vec_ddelta, ddl);
      if (has_z) {
        float zz[4], yp[4], dz[4];
        V4<T>::ld(reinterpret_cast<const T*>(rbase + lay.raw_z) + o, zz);
        V4<T>::ld(reinterpret_cast<const T*>(rbase + lay.raw_yp) + o, yp);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float sg = rcp_approx(1.f + ex2_approx(-zz[e] * kLog2e));
          dz[e] = go[e] * yp[e] * sg * fmaf(zz[e], 1.f - sg, 1.f);
        }
        store4(p.dz, p.dz_bs, p.dz_ls, t0 + my_t, p.vec_dz, dz);
      }
    }
    // per-CTA partial of dB / dC for this chunk -> workspace [B][ntiles][L][N]
    {
      float* wsB = p.ws_dB + (((int64_t)b * p.ntiles + tile) * p.L + t0) * p.N;
      float* wsC = p.ws_dC + (((int64_t)b * p.ntiles + tile) * p.L + t0) * p.N;
      const float* redB = reinterpret_cast<const float*>(wbase + lay.w_rB);
      const float* redC = reinterpret_cast<const float*>(wbase + lay.w_rC);
      if (NPT == p.N && (p.N & 3) == 0) {
        const int nv = rv * p.N / 4;
        for (int i = ht; i < nv; i += kBHelperThreads) {
          reinterpret_cast<float4*>(wsB)[i] = reinterpret_cast<const float4*>(redB)[i];
          reinterpret_cast<float4*>(wsC)[i] = reinterpret_cast<const float4*>(redC)[i];
        }
      } else {
        for (int i = ht; i < rv * p.N; i += kBHelperThreads) {
          const int t = i / p.N, n = i - t * p.N;
          wsB[i] = redB[t * NPT + n];
          wsC[i] = redC[t * NPT + n];
        }
      }
    }
  };

  if constexpr (HT == 1) {
    // prologue
    for (int it = 0; it < kBRing - 1; ++it) issue_loads(it, it);
    int rslot = 0, pslot = kBRing - 1;
    for (int it = 0; it < nck; ++it) {
      cp_async_wait<kBRing - 2>();
      bar_sync(kBarTeam, kBHelperThreads);
      pre_pass(it, rslot);
      bar_arrive(kBarReady + (it & 1), bar_count);  // READY
      if (it >= 1) {
        bar_sync(kBarDone + ((it - 1) & 1), bar_count);  // DONE of the previous chunk
        post_pass(it - 1, pslot);
      }
      bar_sync(kBarTeam, kBHelperThreads);
      issue_loads(it + kBRing - 1, pslot);
      pslot = rslot;
      rslot = (rslot + 1 == kBRing) ? 0 : rslot + 1;
    }
    bar_sync(kBarDone + ((nck - 1) & 1), bar_count);
    post_pass(nck - 1, pslot);
  } else {
    // team t owns iterations t, t + HT, ... and slots t, t + HT (raw and work alike: slot = iteration % NSLOT)
    const int bt = kBarTeam + team;
    issue_loads(team, team);                       // (an empty group when there is no such chunk)
    issue_loads(team + HT, team + HT);
    if (team < nck) {
      cp_async_wait<1>();
      bar_sync(bt, kBHelperThreads);
      pre_pass(team, team);
      bar_arrive(kBarReady + team, bar_count);
    }
    for (int it = team; it < nck; it += HT) {
      const int slot = it & (NSLOT - 1), nxt = it + HT;
      if (nxt < nck) {   // the team's next chunk is prepared BEFORE it waits for the scan warps to finish this one
        cp_async_wait<0>();
        bar_sync(bt, kBHelperThreads);
        pre_pass(nxt, nxt & (NSLOT - 1));
        bar_arrive(kBarReady + (nxt & (NSLOT - 1)), bar_count);
      }
      bar_sync(kBarDone + slot, bar_count);
      post_pass(it, slot);
      bar_sync(bt, kBHelperThreads);           // every thread of the team is done with the raw slot
      issue_loads(it + 2 * HT, slot);
    }
  }

  // ---- dD / d_bias: sum this CTA's 16 timestep-rows in shared memory, one partial per (b, d) ------------------
  bar_sync(kBarHelpers, HT * kBHelperThreads);
  float* fin = reinterpret_cast<float*>(work_base);  // [2][HT * 16][32], the work slots are free now
  constexpr int kRows = HT * kBHelperThreads / 8;
  const int my_r = team * (kBHelperThreads / 8) + my_t;
  *reinterpret_cast<float4*>(fin + my_r * kBD + my_c) = make_float4(dD_acc[0], dD_acc[1], dD_acc[2], dD_acc[3]);
  *reinterpret_cast<float4*>(fin + (kRows + my_r) * kBD + my_c) = make_float4(db_acc[0], db_acc[1], db_acc[2], db_acc[3]);
  bar_sync(kBarHelpers, HT * kBHelperThreads);
  if (team == 0 && ht < kBD && d0 + ht < p.D) {
    float sD = 0.f, sb = 0.f;
    for (int r = 0; r < kRows; ++r) {
      sD += fin[r * kBD + ht];
      sb += fin[(kRows + r) * kBD + ht];
    }
    p.ws_dD[(int64_t)b * p.D + d0 + ht] = sD;
    p.ws_db[(int64_t)b * p.D + d0 + ht] = sb;
  }
}

// Finalize: fixed-order reductions of the per-CTA / per-batch partials.
//   blocks [0, nrow_blocks): dB/dC[b, t, :] = sum over channel tiles; remaining blocks: dA, dD, d_bias.
template <typename T>
__global__ void scan_bwd_finalize_kernel(const ScanBwdParams p, int rows_per_block, int nrow_blocks) {
  const int tid = threadIdx.x;
  if ((int)blockIdx.x < nrow_blocks) {
    const int64_t row0 = (int64_t)blockIdx.x * rows_per_block;  // row = b * L + t
    const int64_t nrows = (int64_t)p.B * p.L;
    const int64_t ts = (int64_t)p.L * p.N;
    if (p.fixed_acc) {
      // the tiles already added their partials into one fixed-point accumulator per (b, t, n): convert and store
      const long long* aB = reinterpret_cast<const long long*>(p.ws_dB);
      const long long* aC = reinterpret_cast<const long long*>(p.ws_dC);
      const int nel = rows_per_block * p.N;
      for (int i = tid; i < nel; i += blockDim.x) {
        const int64_t row = row0 + i / p.N;
        const int n = i % p.N;
        if (row >= nrows) break;
        const int b = (int)(row / p.L);
        const int64_t t = row - (int64_t)b * p.L;
        IO<T>::st(static_cast<T*>(p.dB) + (int64_t)b * p.dB_bs + t * p.dB_ls + n, (float)aB[row * p.N + n] * kAccInvScale);
        IO<T>::st(static_cast<T*>(p.dC) + (int64_t)b * p.dC_bs + t * p.dC_ls + n, (float)aC[row * p.N + n] * kAccInvScale);
      }
      return;
    }
    if ((p.N & 3) == 0) {
      // 4 states per thread: 16-byte loads of the per-tile partials, 8 of them in flight (4 interleaved partial
      // sums per array, combined in a fixed order)
      const int nq = p.N >> 2;
      for (int i = tid; i < rows_per_block * nq; i += blockDim.x) {
        const int64_t row = row0 + i / nq;
        const int n = (i % nq) * 4;
        if (row >= nrows) break;
        const int b = (int)(row / p.L);
        const int64_t t = row - (int64_t)b * p.L;
        const float* sB = p.ws_dB + (((int64_t)b * p.ntiles) * p.L + t) * p.N + n;
        const float* sC = p.ws_dC + (((int64_t)b * p.ntiles) * p.L + t) * p.N + n;
        float4 aB[4], aC[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) aB[q] = aC[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        auto add4 = [](float4& a, const float4 v) { a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w; };
        int k = 0;
        for (; k + 4 <= p.ntiles; k += 4) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            add4(aB[q], __ldcs(reinterpret_cast<const float4*>(sB + (k + q) * ts)));
            add4(aC[q], __ldcs(reinterpret_cast<const float4*>(sC + (k + q) * ts)));
          }
        }
        for (; k < p.ntiles; ++k) {
          add4(aB[0], __ldcs(reinterpret_cast<const float4*>(sB + k * ts)));
          add4(aC[0], __ldcs(reinterpret_cast<const float4*>(sC + k * ts)));
        }
        const float rB[4] = {(aB[0].x + aB[1].x) + (aB[2].x + aB[3].x), (aB[0].y + aB[1].y) + (aB[2].y + aB[3].y),
                             (aB[0].z + aB[1].z) + (aB[2].z + aB[3].z), (aB[0].w + aB[1].w) + (aB[2].w + aB[3].w)};
        const float rC[4] = {(aC[0].x + aC[1].x) + (aC[2].x + aC[3].x), (aC[0].y + aC[1].y) + (aC[2].y + aC[3].y),
                             (aC[0].z + aC[1].z) + (aC[2].z + aC[3].z), (aC[0].w + aC[1].w) + (aC[2].w + aC[3].w)};
        T* oB = static_cast<T*>(p.dB) + (int64_t)b * p.dB_bs + t * p.dB_ls + n;
        T* oC = static_cast<T*>(p.dC) + (int64_t)b * p.dC_bs + t * p.dC_ls + n;
#pragma unroll
        for (int e = 0; e < 4; ++e) IO<T>::st(oB + e, rB[e]), IO<T>::st(oC + e, rC[e]);
      }
      return;
    }
    const int nel = rows_per_block * p.N;
    for (int i = tid; i < nel; i += blockDim.x) {
      const int64_t row = row0 + i / p.N;
      const int n = i % p.N;
      if (row >= nrows) break;
      const int b = (int)(row / p.L);
      const int64_t t = row - (int64_t)b * p.L;
      const float* sB = p.ws_dB + (((int64_t)b * p.ntiles) * p.L + t) * p.N + n;
      const float* sC = p.ws_dC + (((int64_t)b * p.ntiles) * p.L + t) * p.N + n;
      // fixed-order sum with 8 loads in flight (4 interleaved partial sums, combined in a fixed order)
      float aB[4] = {0.f, 0.f, 0.f, 0.f}, aC[4] = {0.f, 0.f, 0.f, 0.f};
      int k = 0;
      for (; k + 4 <= p.ntiles; k += 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          aB[q] += sB[(k + q) * ts];
          aC[q] += sC[(k + q) * ts];
        }
      }
      for (; k < p.ntiles; ++k) {
        aB[0] += sB[k * ts];
        aC[0] += sC[k * ts];
      }
      const float accB = (aB[0] + aB[1]) + (aB[2] + aB[3]), accC = (aC[0] + aC[1]) + (aC[2] + aC[3]);
      IO<T>::st(static_cast<T*>(p.dB) + (int64_t)b * p.dB_bs + t * p.dB_ls + n, accB);
      IO<T>::st(static_cast<T*>(p.dC) + (int64_t)b * p.dC_bs + t * p.dC_ls + n, accC);
    }
  } else {
    const int64_t idx0 = ((int64_t)blockIdx.x - nrow_blocks) * blockDim.x + tid;
    const int64_t stride = (int64_t)(gridDim.x - nrow_blocks) * blockDim.x;
    const int64_t DN = (int64_t)p.D * p.N;
    for (int64_t i = idx0; i < DN; i += stride) {  // i = n * D + d  (coalesced reads)
      const int n = (int)(i / p.D), d = (int)(i % p.D);
      float acc = 0.f;
      for (int b = 0; b < p.B; ++b) acc += p.ws_dA[(int64_t)b * DN + i];
      // with A given as A_log: d/dA_log = dA * A, A = -exp(A_log)
      p.dA[(int64_t)d * p.N + n] = (p.flags & MAMBA_FLAG_A_IS_LOG) ? acc * load_A(p.A, (int64_t)d * p.N + n, p.flags) : acc;
    }
    for (int64_t d = idx0; d < p.D; d += stride) {
      float sD = 0.f, sb = 0.f;
      for (int b = 0; b < p.B; ++b) {
        sD += p.ws_dD[(int64_t)b * p.D + d];
        sb += p.ws_db[(int64_t)b * p.D + d];
      }
      if (p.dD) p.dD[d] = sD;
      if (p.ddbias) p.ddbias[d] = sb;
    }
  }
}

template <typename T, int kCK>
static size_t bwd_smem_bytes(int NW, int NPT, int NPER, int HT) {
  const BwdLayout<T, kCK> lay(NW, NPT, NPER);
  return (size_t)2 * HT * lay.raw_bytes + (size_t)2 * HT * lay.work_bytes + lay.hs_bytes;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t bwd_workspace_layout(int B, int L, int D, int N, size_t* o_dB, size_t* o_dC, size_t* o_dA, size_t* o_dD,
                                   size_t* o_db) {
  const size_t ntiles = (size_t)ceil_div(D, kBD);
  size_t off = 0;
  *o_dB = off, off = align256(off + (size_t)4 * B * ntiles * L * N);
  *o_dC = off, off = align256(off + (size_t)4 * B * ntiles * L * N);
  *o_dA = off, off = align256(off + (size_t)4 * B * N * D);
  *o_dD = off, off = align256(off + (size_t)4 * B * D);
  *o_db = off, off = align256(off + (size_t)4 * B * D);
  return off;
}

template <typename T>
static int launch_finalize(const ScanBwdParams& p, cudaStream_t stream);

template <typename T, int NPER, int NW, int kCK, int HT = 1>
static int launch_bwd(const ScanBwdParams& p, cudaStream_t stream) {
  const size_t smem = bwd_smem_bytes<T, kCK>(NW, p.NPT, NPER, HT);
  if (smem > 227 * 1024)
    return set_error(MAMBA_ESIZE, "scan_bwd: d_state %d needs %zu B of shared memory", p.N, smem);
  auto kern = scan_bwd_kernel<T, NPER, NW, kCK, HT>;
  static thread_local SmemConfig cfg;
  if (int rc = ensure_dynamic_smem(kern, smem, cfg, "scan_bwd")) return rc;
  dim3 grid(p.ntiles, p.B);
  kern<<<grid, (NW + HT * kBHelperWarps) * 32, smem, stream>>>(p);
  count_launch();
  int rc = check_launch("scan_bwd");
  if (rc) return rc;
  return launch_finalize<T>(p, stream);
}

template <typename T>
static int launch_finalize(const ScanBwdParams& p, cudaStream_t stream) {
  const int rows_per_block = (p.N & 3) == 0 ? 1024 / p.N > 0 ? 1024 / p.N : 1 : 8;  // 256 threads x 4 states each
  const int nrow_blocks = (int)(((int64_t)p.B * p.L + rows_per_block - 1) / rows_per_block);
  const int nsmall = ceil_div(p.D * p.N, 256 * 4);
  scan_bwd_finalize_kernel<T><<<nrow_blocks + nsmall, 256, 0, stream>>>(p, rows_per_block, nrow_blocks);
  count_launch();
  return check_launch("scan_bwd_finalize");
}

template <typename T, int NPER, int kCK>
static int bwd_dispatch_nw(const ScanBwdParams& p, cudaStream_t stream) {
  switch (p.NW) {
    case 1: return launch_bwd<T, NPER, 1, kCK>(p, stream);
    case 2: return launch_bwd<T, NPER, 2, kCK>(p, stream);
    case 4: return launch_bwd<T, NPER, 4, kCK>(p, stream);
    case 8: return launch_bwd<T, NPER, 8, kCK>(p, stream);
  }
  return set_error(MAMBA_ESIZE, "scan_bwd: d_state %d too large (max %d)", p.N, kBMaxWarps * 16);
}

template <typename T, int NPER>
static int bwd_dispatch_ck(const ScanBwdParams& p, cudaStream_t stream) {
  if (NPER == 4 && p.helper_teams == 2 && p.NW == 4)   // d_state 16, four scan warps + two helper teams
    return p.ck == 8 ? launch_bwd<T, 4, 4, 8, 2>(p, stream) : launch_bwd<T, 4, 4, 16, 2>(p, stream);
  return p.ck == 8 ? bwd_dispatch_nw<T, NPER, 8>(p, stream) : bwd_dispatch_nw<T, NPER, 16>(p, stream);
}

template <typename T>
static int bwd_dispatch(ScanBwdParams& p, int nper, cudaStream_t stream) {
  if (nper == 1) {  // fused recompute/reverse kernel with tensor-pipe channel sums (scan_bwd_fused.cu)
    if (p.N != 64 && p.N != 32)
      return set_error(MAMBA_EINVAL, "scan_bwd: variant 1 (fused, tensor-pipe reductions) needs d_state 32 or 64 (got %d)", p.N);
    p.NW = p.N / 8, p.NPT = p.N;
    // Fixed-point accumulation of dB / dC across the channel tiles (scan_bwd.cuh), opt-in with
    // MAMBA_B200_BWD_FIXED_ACC=1: it removes the [B][ntiles][L][N] partial tensors (DRAM traffic per launch 366 MB ->
    // ~235 MB, finalize 40 us -> 5 us) but the 33.6 M 64-bit atomics per launch cost more than they save — measured
    // 765 us against 737 us per backward at the training shape (profiles/r02_summary.md) — so the per-tile partials
    // stay the default.  Needs the 8-byte slots to fit in the partial workspace (two tiles or more).
    static const bool fixed = [] { const char* e = getenv("MAMBA_B200_BWD_FIXED_ACC"); return e && e[0] == '1'; }();
    p.fixed_acc = (fixed && p.ntiles >= 2) ? 1 : 0;
    if (p.fixed_acc) {
      const size_t bytes = (size_t)p.B * p.L * p.N * 8;
      cudaMemsetAsync(p.ws_dB, 0, bytes, stream);
      cudaMemsetAsync(p.ws_dC, 0, bytes, stream);
    }
    const int rc = launch_scan_bwd_fused<T>(p, stream);
    return rc ? rc : launch_finalize<T>(p, stream);
  }
  p.NW = ceil_div(p.N, nper);
  if (p.NW > kBMaxWarps) return set_error(MAMBA_ESIZE, "scan_bwd: d_state %d too large (max %d)", p.N, kBMaxWarps * 16);
  p.NW = p.NW <= 2 ? p.NW : (p.NW <= 4 ? 4 : 8);  // instantiated warp counts
  p.NPT = (p.NW * nper + 7) & ~7;
  switch (nper) {
    case 2: return bwd_dispatch_ck<T, 2>(p, stream);
    case 4: return bwd_dispatch_ck<T, 4>(p, stream);
    case 8: return bwd_dispatch_ck<T, 8>(p, stream);
    case 16: return bwd_dispatch_ck<T, 16>(p, stream);
  }
  return set_error(MAMBA_EINVAL, "scan_bwd: variant must be 0, 1, 2, 4, 8, 16 or 104 (got %d)", nper);
}

static bool bvec_ok(const void* ptr, int64_t bs, int64_t ls, size_t elt) {
  return aligned16(ptr) && (bs * elt) % 16 == 0 && (ls * elt) % 16 == 0;
}

}  // namespace mb

extern "C" size_t mamba_scan_bwd_workspace_bytes(int batch, int seqlen, int dim, int dstate) {
  if (batch <= 0 || seqlen <= 0 || dim <= 0 || dstate <= 0) return 0;
  size_t a, b, c, d, e;
  return mb::bwd_workspace_layout(batch, seqlen, dim, dstate, &a, &b, &c, &d, &e);
}

extern "C" int mamba_scan_bwd(const MambaScanBwdArgs* a, void* stream) {
  using namespace mb;
  if (!a || a->struct_size != (int32_t)sizeof(MambaScanBwdArgs))
    return set_error(MAMBA_EINVAL, "scan_bwd: bad args pointer or struct_size");
  if (a->batch <= 0 || a->seqlen <= 0 || a->dim <= 0 || a->dstate <= 0)
    return set_error(MAMBA_EINVAL, "scan_bwd: batch/seqlen/dim/dstate must be positive (got %d/%d/%d/%d)", a->batch,
                     a->seqlen, a->dim, a->dstate);
  if (a->batch > 65535) return set_error(MAMBA_ESIZE, "scan_bwd: batch %d above 65535", a->batch);
  if (!a->u || !a->delta || !a->A || !a->B || !a->C || !a->dout || !a->du || !a->ddelta || !a->dB || !a->dC || !a->dA)
    return set_error(MAMBA_EINVAL, "scan_bwd: null input or output pointer");
  if (a->chunk != 8 && a->chunk != 16) return set_error(MAMBA_EINVAL, "scan_bwd: chunk must be 8 or 16 (got %d)", a->chunk);
  if (a->seqlen > a->chunk && !a->ckpt) return set_error(MAMBA_EINVAL, "scan_bwd: ckpt == NULL");
  if (a->ckpt && !aligned16(a->ckpt)) return set_error(MAMBA_EALIGN, "scan_bwd: ckpt must be 16-byte aligned");
  if ((a->flags & MAMBA_FLAG_HAS_Z) && (!a->z || !a->dz || !a->y_pre))
    return set_error(MAMBA_EINVAL, "scan_bwd: HAS_Z but z/dz/y_pre NULL");
  if ((a->flags & MAMBA_FLAG_HAS_D) && !a->D) return set_error(MAMBA_EINVAL, "scan_bwd: HAS_D but D == NULL");
  if ((a->flags & MAMBA_FLAG_HAS_DELTA_BIAS) && !a->delta_bias)
    return set_error(MAMBA_EINVAL, "scan_bwd: HAS_DELTA_BIAS but delta_bias == NULL");
  if (a->dtype != MAMBA_F32 && a->dtype != MAMBA_BF16) return set_error(MAMBA_EDTYPE, "scan_bwd: dtype %d", a->dtype);

  ScanBwdParams p{};
  p.B = a->batch, p.L = a->seqlen, p.D = a->dim, p.N = a->dstate, p.flags = a->flags;
  p.ck = a->chunk;
  p.nck = ceil_div(p.L, p.ck);
  p.ntiles = ceil_div(p.D, kBD);
  p.u = a->u, p.delta = a->delta, p.Bm = a->B, p.Cm = a->C, p.z = a->z, p.dout = a->dout, p.ypre = a->y_pre;
  p.u_bs = a->u_bs, p.u_ls = a->u_ls, p.delta_bs = a->delta_bs, p.delta_ls = a->delta_ls;
  p.B_bs = a->B_bs, p.B_ls = a->B_ls, p.C_bs = a->C_bs, p.C_ls = a->C_ls;
  p.z_bs = a->z_bs, p.z_ls = a->z_ls, p.dout_bs = a->dout_bs, p.dout_ls = a->dout_ls;
  p.ypre_bs = a->y_pre_bs, p.ypre_ls = a->y_pre_ls;
  p.du = a->du, p.ddelta = a->ddelta, p.dz = a->dz, p.dB = a->dB, p.dC = a->dC;
  p.du_bs = a->du_bs, p.du_ls = a->du_ls, p.ddelta_bs = a->ddelta_bs, p.ddelta_ls = a->ddelta_ls;
  p.dz_bs = a->dz_bs, p.dz_ls = a->dz_ls, p.dB_bs = a->dB_bs, p.dB_ls = a->dB_ls, p.dC_bs = a->dC_bs, p.dC_ls = a->dC_ls;
  p.A = a->A, p.Dv = a->D, p.dbias = a->delta_bias, p.ckpt = a->ckpt;
  p.dA = a->dA, p.dD = a->dD, p.ddbias = a->ddelta_bias;

  size_t o_dB, o_dC, o_dA, o_dD, o_db;
  const size_t need = bwd_workspace_layout(p.B, p.L, p.D, p.N, &o_dB, &o_dC, &o_dA, &o_dD, &o_db);
  if (!a->workspace || a->workspace_bytes < need)
    return set_error(MAMBA_ESIZE, "scan_bwd: workspace %zu B < required %zu B", a->workspace_bytes, need);
  if (!aligned16(a->workspace)) return set_error(MAMBA_EALIGN, "scan_bwd: workspace must be 16-byte aligned");
  char* ws = static_cast<char*>(a->workspace);
  p.ws_dB = reinterpret_cast<float*>(ws + o_dB), p.ws_dC = reinterpret_cast<float*>(ws + o_dC);
  p.ws_dA = reinterpret_cast<float*>(ws + o_dA), p.ws_dD = reinterpret_cast<float*>(ws + o_dD);
  p.ws_db = reinterpret_cast<float*>(ws + o_db);

  const size_t elt = a->dtype == MAMBA_F32 ? 4 : 2;
  p.vec_u = bvec_ok(a->u, a->u_bs, a->u_ls, elt);
  p.vec_delta = bvec_ok(a->delta, a->delta_bs, a->delta_ls, elt);
  p.vec_z = a->z ? bvec_ok(a->z, a->z_bs, a->z_ls, elt) : 0;
  p.vec_ypre = a->y_pre ? bvec_ok(a->y_pre, a->y_pre_bs, a->y_pre_ls, elt) : 0;
  p.vec_dout = bvec_ok(a->dout, a->dout_bs, a->dout_ls, elt);
  p.vec_B = bvec_ok(a->B, a->B_bs, a->B_ls, elt);
  p.vec_C = bvec_ok(a->C, a->C_bs, a->C_ls, elt);
  p.vec_du = bvec_ok(a->du, a->du_bs, a->du_ls, elt);
  p.vec_ddelta = bvec_ok(a->ddelta, a->ddelta_bs, a->ddelta_ls, elt);
  p.vec_dz = a->dz ? bvec_ok(a->dz, a->dz_bs, a->dz_ls, elt) : 0;
  p.vec_ck = a->ckpt && aligned16(a->ckpt) && (p.D % 4 == 0);

  // variant: 0 = auto, 1 = fused kernel with tensor-pipe channel reductions (d_state 32 / 64), 2 / 4 / 8 / 16 =
  // lane<->channel kernel with that many states per thread (2: d_state <= 16, eight scan warps; measured 1451 against
  // 1491 us at config 5, B=2 — the four helper warps pace the CTA, not the scan warps — so auto keeps 4)
  int nper = a->variant;
  // 104: the 4-states-per-thread kernel with TWO helper teams (d_state 9..16).  Auto picks it when the grid is a single
  // wave of one CTA per SM (the second team then does what a second resident CTA's helpers would do)
  p.helper_teams = 1;
  if (nper == 104) nper = 4, p.helper_teams = 2;
  // auto: the fused kernel wins with bf16 I/O (plain tf32 column sums); with fp32 I/O its split-tf32 MMAs cost
  // more than the shuffles they replace
  if (nper == 0) nper = ((p.N == 64 || p.N == 32) && a->dtype == MAMBA_BF16) ? 1 : (p.N <= 32 ? 4 : 8);
  while (nper > 1 && nper < 16 && ceil_div(p.N, nper) > kBMaxWarps) nper *= 2;
  if (a->variant == 0 && nper == 4 && p.N > 8 && p.N <= 16 && (int64_t)p.ntiles * p.B <= kNumSMs) p.helper_teams = 2;
  if (p.helper_teams == 2 && !(nper == 4 && p.N > 8 && p.N <= 16))
    return set_error(MAMBA_EINVAL, "scan_bwd: variant 104 (two helper teams) needs d_state 9..16 (got %d)", p.N);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return a->dtype == MAMBA_F32 ? bwd_dispatch<float>(p, nper, st) : bwd_dispatch<__nv_bfloat16>(p, nper, st);
}
