// scan_fwd_tma.cu — selective-scan forward, second generation: TMA-staged, homogeneous warps, MUFU/FMA co-issue.
//
// Same mathematics and outputs as scan_fwd.cu (MambaBlock.selective_scan, reference
// models/mamba/__pycache__/simple_mamba.cpython-311.pyc @L310-333, fused with softplus @L276, the D skip @L331 and the
// z gate @L241); what changed is how the SM is kept busy:
//
//   * operands arrive by TMA.  One elected thread issues five `cp.async.bulk.tensor.3d` copies per 16-step stage
//     (u, delta, z: box [32 channels x 16 steps]; B, C: box [d_state x 16 steps]) into a ring of RR slots; each slot
//     has one mbarrier armed with the stage's byte count.  Rows past the end of the sequence and channels past D are
//     zero-filled by the copy engine, so there is no ragged-edge code on the load side and no helper warp spends issue
//     slots on LDGSTS address arithmetic.
//   * the producer is one thread of an extra warp; it waits on per-slot "empty" mbarriers and never takes part in the
//     scan warps' barrier (when a scan warp issued the copies, the other warps waited for it at the next stage barrier).
//   * there are no helper warps.  lane <-> channel, warp s owns states [s*NPER, (s+1)*NPER) for the whole sequence
//     (h in registers).  The per-(t, d) work — softplus(delta + bias), delta*u, the sum of the per-warp partials
//     <h, C>, + D*u, * silu(z), the stores — is dealt out row by row to the same warps and sits INSIDE the unrolled
//     16-step scan loop of the neighbouring stages (post-pass of stage c-1 and pre-pass of stage c+1 while stage c is
//     scanned), so ptxas interleaves it with the recurrence and its latencies hide behind the MUFU stream.  One
//     CTA-wide barrier per stage.
//   * the scan is bound by the MUFU pipe (16 ex2/clk/SM).  PK of every thread's NPER/2 state pairs take their decay
//     exp2(delta*A) from the FMA pipe instead: round-to-nearest split with the 1.5*2^23 trick, a degree-5 (fp32 I/O)
//     or degree-4 (bf16 I/O) polynomial on [-1/2, 1/2] in packed fp32x2, exponent insertion on the ALU pipe.  Written
//     as 1 + f*q(f) so that the error near delta*A = 0 is relative to 1 - a (what the recurrence is sensitive to).
//
// Eligibility (else the caller uses scan_fwd.cu): 16-byte aligned bases and strides (tensor maps), d_state <= 128.
#include <cuda.h>

#include <type_traits>

#include "scan_fwd.cuh"

namespace mb {

namespace {

constexpr int kDT = 32;  // channels per CTA (= warp width)
constexpr int kTS = 16;  // timesteps per stage

// ---- mbarrier / TMA primitives ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---- exp2 on the FMA pipe ------------------------------------------------------------------------------------------
// 2^g for a pair of arguments: g = i + f, i = rne(g), f in [-1/2, 1/2]; 2^f = 1 + f*q(f); result = 2^f with i added to the
// exponent field.  Arguments below -125 are clamped (2^-125 instead of a denormal / zero: far below anything the
// recurrence can resolve).  DEG 5: |rel err| < 2.1e-7 (fp32-grade); DEG 4: 7.2e-6 (bf16 I/O).
template <int DEG>
__device__ __forceinline__ float2 exp2_fma2(float2 g) {
  g.x = fmaxf(g.x, -125.f), g.y = fmaxf(g.y, -125.f);
  const float2 M = make_float2(12582912.f, 12582912.f), nM = make_float2(-12582912.f, -12582912.f);
  const float2 t = __fadd2_rn(g, M);      // integer part in the low mantissa bits
  const float2 r = __fadd2_rn(t, nM);     // rne(g) as a float
  const float2 f = __fadd2_rn(g, make_float2(-r.x, -r.y));
  float2 q;
  if constexpr (DEG >= 5) {
    q = __ffma2_rn(f, make_float2(1.3390867e-3f, 1.3390867e-3f), make_float2(9.6663740e-3f, 9.6663740e-3f));
    q = __ffma2_rn(q, f, make_float2(5.5503570e-2f, 5.5503570e-2f));
  } else {
    q = __ffma2_rn(f, make_float2(9.6663740e-3f, 9.6663740e-3f), make_float2(5.5838343e-2f, 5.5838343e-2f));
  }
  q = __ffma2_rn(q, f, make_float2(2.4022348e-1f, 2.4022348e-1f));
  q = __ffma2_rn(q, f, DEG >= 5 ? make_float2(6.9314718e-1f, 6.9314718e-1f) : make_float2(6.9313675e-1f, 6.9313675e-1f));
  const float2 pm = __ffma2_rn(q, f, make_float2(1.f, 1.f));
  float2 o;
  o.x = __int_as_float(__float_as_int(pm.x) + (__float_as_int(t.x) << 23));
  o.y = __int_as_float(__float_as_int(pm.y) + (__float_as_int(t.y) << 23));
  return o;
}

// ---- shared-memory layout ---------------------------------------------------------------------------------------------
template <typename T>
struct TmaLayout {
  int raw_u, raw_dl, raw_z, raw_B, raw_C, raw_bytes;
  int w_dlu, w_y, w_Bf, w_Cf, work_bytes;
  int off_work, off_bar, total;
  __host__ __device__ TmaLayout(int NS, int NPT, int RR) {
    constexpr int S = (int)sizeof(T);
    auto up = [](int x) { return (x + 127) & ~127; };
    int o = 0;
    raw_u = o, o = up(o + kTS * kDT * S);
    raw_dl = o, o = up(o + kTS * kDT * S);
    raw_z = o, o = up(o + kTS * kDT * S);
    raw_B = o, o = up(o + kTS * NPT * S);
    raw_C = o, o = up(o + kTS * NPT * S);
    raw_bytes = o;
    o = 0;
    w_dlu = o, o = up(o + kTS * kDT * 8);
    w_y = o, o = up(o + NS * kTS * kDT * 4);
    w_Bf = o, o = up(o + (S == 4 ? 0 : kTS * NPT * 4));  // bf16 I/O: B / C widened by the pre-pass
    w_Cf = o, o = up(o + (S == 4 ? 0 : kTS * NPT * 4));
    work_bytes = o;
    off_work = RR * raw_bytes;
    off_bar = off_work + 2 * work_bytes;  // full[RR] then empty[RR]
    total = off_bar + 16 * RR;
  }
};

__device__ __forceinline__ float4 ld4f(const float* p) { return *reinterpret_cast<const float4*>(p); }

// Predicated global stores and always-executed MUFU ops.  The stage body must stay ONE basic block (ptxas interleaves
// the pre-/post-pass rows with the recurrence only inside a block): nvcc turns `if (ok) store` and selects with an
// expensive arm into branches, so these are spelled as predicated / volatile PTX.
__device__ __forceinline__ void stg_if(float* p, float v, bool ok) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.s32 q, %2, 0;\n@q st.global.f32 [%0], %1;\n}" ::"l"(p), "f"(v), "r"((int)ok));
}
__device__ __forceinline__ void stg_if(__nv_bfloat16* p, float v, bool ok) {
  const __nv_bfloat16 hv = __float2bfloat16_rn(v);
  asm volatile("{\n.reg .pred q;\nsetp.ne.s32 q, %2, 0;\n@q st.global.b16 [%0], %1;\n}" ::"l"(p),
               "h"(*reinterpret_cast<const unsigned short*>(&hv)), "r"((int)ok));
}
__device__ __forceinline__ void stg4_if(float4* p, float4 v, bool ok) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.s32 q, %5, 0;\n@q st.global.v4.f32 [%0], {%1,%2,%3,%4};\n}" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w), "r"((int)ok));
}
__device__ __forceinline__ void stg2_if(float2* p, float2 v, bool ok) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.s32 q, %3, 0;\n@q st.global.v2.f32 [%0], {%1,%2};\n}" ::"l"(p), "f"(v.x), "f"(v.y),
               "r"((int)ok));
}
__device__ __forceinline__ float ex2_always(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// softplus_fast / silu_fast (common.cuh) with the first MUFU op pinned, so that a select on a uniform flag cannot
// become a branch around them
__device__ __forceinline__ float softplus_nobranch(float x) {
  const float e = ex2_always(x * kLog2e);
  const float series = e * fmaf(e, fmaf(e, fmaf(e, -0.25f, 0.33333334f), -0.5f), 1.f);
  const float full = lg2_approx(1.f + e) * kLn2;
  const float r = x < -4.f ? series : full;
  return x > 20.f ? x : r;
}
__device__ __forceinline__ float silu_nobranch(float x) { return x * rcp_approx(1.f + ex2_always(-x * kLog2e)); }

// NPER states per thread, NS scan warps, PK state pairs per thread on the FMA-pipe exp2 (every step, or only every
// second step when ALT).
template <typename T, int NPER, int NS, int PK, bool ALT, bool SP>
__global__ void __launch_bounds__((NS + 1) * 32, 1)
    scan_fwd_tma_kernel(const ScanFwdParams p, const int RR, const __grid_constant__ CUtensorMap tm_u,
                        const __grid_constant__ CUtensorMap tm_dl, const __grid_constant__ CUtensorMap tm_z,
                        const __grid_constant__ CUtensorMap tm_B, const __grid_constant__ CUtensorMap tm_C) {
  static_assert(NPER % 2 == 0 && kTS % NS == 0 && PK <= NPER / 2, "tile shape");
  constexpr int NPT = NS * NPER;      // padded d_state
  constexpr int RPW = kTS / NS;       // rows of every stage that one warp pre-/post-processes
  constexpr int NT = NS * 32;
  constexpr bool kF32 = sizeof(T) == 4;
  constexpr int DEG = kF32 ? 5 : 4;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, s = tid >> 5;
  const int b = blockIdx.y, d0 = blockIdx.x * kDT, d = d0 + lane;
  const bool dok = d < p.D;
  const TmaLayout<T> lay(NS, NPT, RR);
  unsigned char* const work_base = smem + lay.off_work;
  const uint32_t bar0 = smem_u32(smem + lay.off_bar);
  const uint32_t raw0 = smem_u32(smem);
  const int nst = (p.L + kTS - 1) / kTS;
  const bool has_z = p.flags & MAMBA_FLAG_HAS_Z;
  const uint32_t stage_bytes = (uint32_t)((has_z ? 3 : 2) * kTS * kDT * sizeof(T) + 2 * kTS * NPT * sizeof(T));

  auto issue_stage = [&](int k) {  // one thread: arm the slot's mbarrier and start the five copies of stage k
    const int slot = k % RR;
    const uint32_t bar = bar0 + 8 * slot, base = raw0 + slot * lay.raw_bytes;
    mbar_expect_tx(bar, stage_bytes);
    tma_load_3d(base + lay.raw_u, &tm_u, bar, d0, k * kTS, b);
    tma_load_3d(base + lay.raw_dl, &tm_dl, bar, d0, k * kTS, b);
    if (has_z) tma_load_3d(base + lay.raw_z, &tm_z, bar, d0, k * kTS, b);
    tma_load_3d(base + lay.raw_B, &tm_B, bar, 0, k * kTS, b);
    tma_load_3d(base + lay.raw_C, &tm_C, bar, 0, k * kTS, b);
  };

  // ---- producer warp: one thread feeds the ring; it never joins the scan warps' barriers ------------------------------
  if (tid == NT) {
    for (int i = 0; i < 2 * RR; ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();  // mbarriers initialised
  if (s == NS) {
    if (lane == 0) {
      for (int k = 0; k < RR && k < nst; ++k) issue_stage(k);
      int slot = 0;
      uint32_t par = 0;
      for (int k = RR; k < nst; ++k) {
        mbar_wait(bar0 + 8 * (RR + slot), par);  // the scan warps have released the slot (post-pass of stage k - RR done)
        issue_stage(k);
        if (++slot == RR) slot = 0, par ^= 1u;
      }
    }
    return;
  }

  // ---- per-thread constants -------------------------------------------------------------------------------------------
  float2 A2[NPER / 2], h[NPER / 2];
#pragma unroll
  for (int j = 0; j < NPER; ++j) {
    const int n = s * NPER + j;
    const bool ok = dok && (n < p.N);
    const float a = ok ? load_A(p.A, (int64_t)d * p.N + n, p.flags) * kLog2e : 0.f;
    const float h0 = (ok && p.h_init) ? p.h_init[((int64_t)b * p.D + d) * p.N + n] : 0.f;
    if (j & 1) A2[j / 2].y = a, h[j / 2].y = h0;
    else A2[j / 2].x = a, h[j / 2].x = h0;
  }
  const float bias = ((p.flags & MAMBA_FLAG_HAS_DELTA_BIAS) && dok) ? p.dbias[d] : 0.f;
  const float Dd = ((p.flags & MAMBA_FLAG_HAS_D) && dok) ? p.Dv[d] : 0.f;
  T* const gout = static_cast<T*>(p.out) + (int64_t)b * p.out_bs + d;
  const bool has_ypre = p.ypre != nullptr;
  T* const gyp = static_cast<T*>(has_ypre ? p.ypre : p.out) + (int64_t)b * p.ypre_bs + d;
  // checkpoints [B][nck][ceil(N/4)][D][4]: running pointer to this thread's first float4 of the next chunk start
  const bool has_ck = p.ckpt != nullptr && dok;
  const int64_t ck_stride = (int64_t)p.N4 * p.D;  // float4s per checkpoint
  float4* ckp = reinterpret_cast<float4*>(has_ck ? p.ckpt : reinterpret_cast<float*>(p.out)) +
                ((int64_t)b * p.nck + 1) * ck_stride + (int64_t)((s * NPER) / 4) * p.D + d;
  const bool ck8 = p.cki == 8;

  // pre-pass of one row of stage k: (softplus(delta + bias), that * u) for this lane's channel
  auto pre_row = [&](int k, int slot, int row) {
    const unsigned char* rb = smem + (size_t)slot * lay.raw_bytes;
    unsigned char* wb = work_base + (size_t)(k & 1) * lay.work_bytes;
    const float dl = IO<T>::ld(reinterpret_cast<const T*>(rb + lay.raw_dl) + row * kDT + lane);
    const float uu = IO<T>::ld(reinterpret_cast<const T*>(rb + lay.raw_u) + row * kDT + lane);
    float v = dl + bias;
    if constexpr (SP) v = softplus_nobranch(v);  // (a uniform runtime flag here becomes a branch that splits the block)
    v = (k * kTS + row >= p.L) ? 0.f : v;     // padded timestep: a = 1, input 0 -> state unchanged
    reinterpret_cast<float2*>(wb + lay.w_dlu)[row * kDT + lane] = make_float2(v, v * uu);
  };
  // bf16 I/O: B / C of stage k as fp32 for the scan's broadcast loads
  auto widen_bc = [&](int k, int slot) {
    if constexpr (!kF32) {
      const unsigned char* rb = smem + (size_t)slot * lay.raw_bytes;
      unsigned char* wb = work_base + (size_t)(k & 1) * lay.work_bytes;
      constexpr int items = kTS * NPT / 8;
#pragma unroll
      for (int i0 = 0; i0 < 2 * items; i0 += NT) {
        const int i = i0 + tid;
        if (i < 2 * items) {
          const bool isC = i >= items;
          const int j = isC ? i - items : i;
          cvt8_bf16_f32(reinterpret_cast<const T*>(rb + (isC ? lay.raw_C : lay.raw_B)) + 8 * j,
                        reinterpret_cast<float*>(wb + (isC ? lay.w_Cf : lay.w_Bf)) + 8 * j);
        }
      }
    }
  };
  // post-pass of one row of stage k: sum of the per-warp partials, + D*u, gate, stores
  auto post_row = [&](int k, int slot, int row) {
    const unsigned char* rb = smem + (size_t)slot * lay.raw_bytes;
    const unsigned char* wb = work_base + (size_t)(k & 1) * lay.work_bytes;
    const float* yp = reinterpret_cast<const float*>(wb + lay.w_y) + row * kDT + lane;
    float y0 = 0.f, y1 = 0.f;
#pragma unroll
    for (int w = 0; w < NS; w += 2) y0 += yp[w * kTS * kDT], y1 += yp[(w + 1) * kTS * kDT];
    const float uu = IO<T>::ld(reinterpret_cast<const T*>(rb + lay.raw_u) + row * kDT + lane);
    float y = fmaf(Dd, uu, y0 + y1);
    const int64_t t = (int64_t)k * kTS + row;
    const bool ok = dok && t < p.L;
    const float zz = IO<T>::ld(reinterpret_cast<const T*>(rb + lay.raw_z) + row * kDT + lane);  // unused garbage if !has_z
    stg_if(gyp + t * p.ypre_ls, y, has_ypre && ok);
    y *= has_z ? silu_nobranch(zz) : 1.f;
    stg_if(gout + t * p.out_ls, y, ok);
  };

  // ---- prologue: pre-pass of stage 0 ------------------------------------------------------------------------------------
  mbar_wait(bar0, 0);
#pragma unroll
  for (int r = 0; r < RPW; ++r) pre_row(0, 0, s + r * NS);
  widen_bc(0, 0);
  bar_sync(1, NT);

  // One stage.  kPrev / kNext (is there a stage before / after this one) are compile-time so that the steady-state
  // body is a single basic block: ptxas interleaves the post- and pre-pass rows with the recurrence.
  int slot_prev = 0, slot_cur = 0, slot_next = RR > 1 ? 1 : 0;  // raw-ring slots of stages c-1, c, c+1
  uint32_t par_next = 0;                                        // phase parity of stage c+1's mbarrier
  auto stage = [&](const int c, auto kPrevT, auto kNextT) {
    constexpr bool has_prev = decltype(kPrevT)::value, has_next = decltype(kNextT)::value;
    const int ws = c & 1;
    const unsigned char* wbase = work_base + (size_t)ws * lay.work_bytes;
    const unsigned char* rbase = smem + (size_t)slot_cur * lay.raw_bytes;
    const float2* dlu = reinterpret_cast<const float2*>(wbase + lay.w_dlu) + lane;
    const float* Bf = (kF32 ? reinterpret_cast<const float*>(rbase + lay.raw_B) : reinterpret_cast<const float*>(wbase + lay.w_Bf)) + s * NPER;
    const float* Cf = (kF32 ? reinterpret_cast<const float*>(rbase + lay.raw_C) : reinterpret_cast<const float*>(wbase + lay.w_Cf)) + s * NPER;
    float* yp = reinterpret_cast<float*>(const_cast<unsigned char*>(wbase) + lay.w_y) + (s * kTS) * kDT + lane;
    if (has_next) mbar_wait(bar0 + 8 * slot_next, par_next);  // issued RR-1 stages ago: normally no wait

    float2 dd_cur, dd_nxt, dd_n2;
    float2 Bc[NPER / 2], Cc[NPER / 2], Bn[NPER / 2], Cn[NPER / 2];  // state pairs
    float2 a_cur[NPER / 2], a_nxt[NPER / 2];
    auto fetch_bc = [&](int t, float2(&Bv)[NPER / 2], float2(&Cv)[NPER / 2]) {
      if constexpr (NPER % 4 == 0) {
#pragma unroll
        for (int q = 0; q < NPER / 4; ++q) {
          const float4 bq = ld4f(Bf + t * NPT + 4 * q), cq = ld4f(Cf + t * NPT + 4 * q);
          Bv[2 * q] = make_float2(bq.x, bq.y), Bv[2 * q + 1] = make_float2(bq.z, bq.w);
          Cv[2 * q] = make_float2(cq.x, cq.y), Cv[2 * q + 1] = make_float2(cq.z, cq.w);
        }
      } else {
#pragma unroll
        for (int k = 0; k < NPER / 2; ++k) {
          Bv[k] = *reinterpret_cast<const float2*>(Bf + t * NPT + 2 * k);
          Cv[k] = *reinterpret_cast<const float2*>(Cf + t * NPT + 2 * k);
        }
      }
    };
    auto decay = [&](const float2 dd, float2(&a)[NPER / 2], const bool fma_step) {
      const float2 dl2 = make_float2(dd.x, dd.x);
#pragma unroll
      for (int k = 0; k < NPER / 2; ++k) {
        const float2 gk = __fmul2_rn(dl2, A2[k]);
        if (k < PK && fma_step) a[k] = exp2_fma2<DEG>(gk);
        else a[k] = make_float2(ex2_approx(gk.x), ex2_approx(gk.y));
      }
    };
    dd_cur = dlu[0];
    dd_nxt = dlu[kDT];
    fetch_bc(0, Bc, Cc);
    decay(dd_cur, a_cur, !ALT);
#pragma unroll
    for (int t = 0; t < kTS; ++t) {
      // the neighbouring stages' per-(t, d) rows, dealt out over the unrolled steps
      if (t % NS == 0 && has_prev) post_row(c - 1, slot_prev, s + (t / NS) * NS);
      if (t % NS == NS / 2 && has_next) pre_row(c + 1, slot_next, s + (t / NS) * NS);
      if (t == 1 && has_next) widen_bc(c + 1, slot_next);
      if (t + 2 < kTS) dd_n2 = dlu[(t + 2) * kDT];
      if (t + 1 < kTS) {
        fetch_bc(t + 1, Bn, Cn);
        decay(dd_nxt, a_nxt, !ALT || ((t + 1) & 1));
      }
      const float2 du2 = make_float2(dd_cur.y, dd_cur.y);
      float2 acc = make_float2(0.f, 0.f);
      // (packed FFMA2 with three distinct register pairs runs at half the lane rate of scalar FFMA — 60 vs 117
      // lane-ops/clk/SM, tools/microbench3.cu — but the loop is short of issue slots, not of FMA lanes: scalar
      // multiply-adds measured 5 % slower here)
#pragma unroll
      for (int k = 0; k < NPER / 2; ++k) {
        h[k] = __ffma2_rn(a_cur[k], h[k], __fmul2_rn(du2, Bc[k]));
        acc = __ffma2_rn(h[k], Cc[k], acc);
      }
      sts_f32(yp + t * kDT, acc.x + acc.y);
      dd_cur = dd_nxt, dd_nxt = dd_n2;
#pragma unroll
      for (int k = 0; k < NPER / 2; ++k) Bc[k] = Bn[k], Cc[k] = Cn[k], a_cur[k] = a_nxt[k];
      if (t == 7 || t == 15) {
        // h is now the state at the start of checkpoint chunk (c*kTS + t + 1) / cki
        const int tg_next = c * kTS + t + 1;
        const bool doit = has_ck && (t == 15 || ck8) && tg_next < p.L;
        if constexpr (NPER % 4 == 0) {
#pragma unroll
          for (int q = 0; q < NPER / 4; ++q)
            stg4_if(ckp + (int64_t)q * p.D, make_float4(h[2 * q].x, h[2 * q].y, h[2 * q + 1].x, h[2 * q + 1].y),
                    doit && s * NPER + 4 * q < p.N);
        } else {  // two states per thread: this thread's half of the float4
          stg2_if(reinterpret_cast<float2*>(ckp) + ((s * NPER) & 2) / 2, h[0], doit && s * NPER < p.N);
        }
        ckp += (t == 15 || ck8) ? ck_stride : 0;
      }
    }
    bar_sync(1, NT);  // stage c scanned by every warp; post(c-1) and pre(c+1) complete
    if (tid == 0 && has_prev) mbar_arrive(bar0 + 8 * (RR + slot_prev));  // slot of stage c-1 is free: tell the producer
    slot_prev = slot_cur, slot_cur = slot_next;
    slot_next = slot_next + 1 == RR ? 0 : slot_next + 1;
    if (slot_next == 0) par_next ^= 1u;
  };
  if (nst == 1) {
    stage(0, std::false_type{}, std::false_type{});
  } else {
    stage(0, std::false_type{}, std::true_type{});
#pragma unroll 1
    for (int c = 1; c < nst - 1; ++c) stage(c, std::true_type{}, std::true_type{});
    stage(nst - 1, std::true_type{}, std::false_type{});
  }
  // epilogue: post-pass of the last stage (slot_prev after the final rotation)
#pragma unroll
  for (int r = 0; r < RPW; ++r) post_row(nst - 1, slot_prev, s + r * NS);
  if (p.h_last != nullptr && dok) {
#pragma unroll
    for (int j = 0; j < NPER; ++j) {
      const int n = s * NPER + j;
      if (n < p.N) p.h_last[((int64_t)b * p.D + d) * p.N + n] = (j & 1) ? h[j / 2].y : h[j / 2].x;
    }
  }
}

// ---- host side ------------------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}

// Tensor map of a [batch, seqlen, X] activation with element strides (bs, ls, 1); box = [bx, kTS, 1].
// Returns false when the tensor cannot be described (alignment) — the caller falls back to the LDGSTS kernel.
bool make_map(CUtensorMap* tm, int dtype, const void* base, int64_t X, int64_t L, int64_t B, int64_t bs, int64_t ls, int bx) {
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr) return false;
  const int64_t elt = dtype == MAMBA_F32 ? 4 : 2;
  if (B == 1) bs = L * ls;
  if (!aligned16(base) || ls <= 0 || bs <= 0 || (ls * elt) % 16 || (bs * elt) % 16 || (bx * elt) % 16) return false;
  if (L > 1 && ls < X) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)X, (cuuint64_t)L, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)(ls * elt), (cuuint64_t)(bs * elt)};
  const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)kTS, 1u};
  const cuuint32_t es[3] = {1u, 1u, 1u};
  const CUresult r = enc(tm, dtype == MAMBA_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                         const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <typename T, int NPER, int NS, int PK, bool ALT, bool SP>
int launch_sp(const ScanFwdParams& p, int dtype, cudaStream_t stream) {
  constexpr int NPT = NS * NPER;
  // ring depth: enough bytes in flight per SM to cover HBM latency (a stage is short when d_state is small)
  const int stage = (int)TmaLayout<T>(NS, NPT, 1).raw_bytes;
  int RR = NPT <= 16 ? 8 : (NPT <= 32 ? 6 : 5);
  while (RR > 3 && TmaLayout<T>(NS, NPT, RR).total > 200 * 1024) --RR;
  (void)stage;
  const TmaLayout<T> lay(NS, NPT, RR);
  const size_t smem = (size_t)lay.total;
  if (smem > 227 * 1024) return kTmaNotEligible;
  CUtensorMap tu, tdl, tz, tB, tC;
  if (!make_map(&tu, dtype, p.u, p.D, p.L, p.B, p.u_bs, p.u_ls, kDT)) return kTmaNotEligible;
  if (!make_map(&tdl, dtype, p.delta, p.D, p.L, p.B, p.delta_bs, p.delta_ls, kDT)) return kTmaNotEligible;
  if (p.flags & MAMBA_FLAG_HAS_Z) {
    if (!make_map(&tz, dtype, p.z, p.D, p.L, p.B, p.z_bs, p.z_ls, kDT)) return kTmaNotEligible;
  } else {
    tz = tu;
  }
  if (!make_map(&tB, dtype, p.Bm, p.N, p.L, p.B, p.B_bs, p.B_ls, NPT)) return kTmaNotEligible;
  if (!make_map(&tC, dtype, p.Cm, p.N, p.L, p.B, p.C_bs, p.C_ls, NPT)) return kTmaNotEligible;
  auto kern = scan_fwd_tma_kernel<T, NPER, NS, PK, ALT, SP>;
  static thread_local SmemConfig cfg;
  if (int rc = ensure_dynamic_smem(kern, smem, cfg, "scan_fwd_tma")) return rc;
  dim3 grid(ceil_div(p.D, kDT), p.B);
  kern<<<grid, (NS + 1) * 32, smem, stream>>>(p, RR, tu, tdl, tz, tB, tC);
  count_launch();
  return check_launch("scan_fwd_tma");
}

template <typename T, int NPER, int NS, int PK, bool ALT>
int launch(const ScanFwdParams& p, int dtype, cudaStream_t stream) {
  return (p.flags & MAMBA_FLAG_DELTA_SOFTPLUS) ? launch_sp<T, NPER, NS, PK, ALT, true>(p, dtype, stream)
                                               : launch_sp<T, NPER, NS, PK, ALT, false>(p, dtype, stream);
}

// tune = 10 * shape + split.  shape 0: automatic, 1 / 2: first / second tiling for this d_state (states per thread x
// scan warps: d_state <= 16: 2x8 | 4x4; <= 32: 4x8 | 2x16; <= 64: 8x8 | 4x16; <= 128: 8x16).  split: share of the exp2
// work on the FMA pipe — 0 none, 1 one state pair per thread on every second step.  Measured (profiles/r02_*): the
// split never pays — with three-operand FFMA2 at half rate the FMA pipe is as loaded as the MUFU pipe — so 0 is the
// default and 1 is kept as the record of the experiment.
template <typename T, int NPER, int NS>
int dispatch_split(const ScanFwdParams& p, int dtype, int split, cudaStream_t stream) {
  if (!(p.flags & MAMBA_FLAG_DELTA_SOFTPLUS)) return launch_sp<T, NPER, NS, 0, false, false>(p, dtype, stream);
  switch (split) {
    case 0: return launch_sp<T, NPER, NS, 0, false, true>(p, dtype, stream);
    case 1: return launch_sp<T, NPER, NS, 1, true, true>(p, dtype, stream);
  }
  return set_error(MAMBA_EINVAL, "scan_fwd_tma: split must be 0 or 1 (got %d)", split);
}

template <typename T>
int dispatch_shape(const ScanFwdParams& p, int dtype, int tune, cudaStream_t stream) {
  int shape = tune / 10;
  const int split = tune % 10;
  // automatic: few CTAs per SM -> more, lighter warps per CTA; several co-resident CTAs -> fewer, heavier warps
  const bool crowded = (int64_t)ceil_div(p.D, kDT) * p.B >= 2 * kNumSMs;
  if (shape == 0) shape = (p.N <= 16 && crowded) ? 2 : 1;
  if (p.N <= 16) return shape == 1 ? dispatch_split<T, 2, 8>(p, dtype, split, stream) : dispatch_split<T, 4, 4>(p, dtype, split, stream);
  if (p.N <= 32) return shape == 1 ? dispatch_split<T, 4, 8>(p, dtype, split, stream) : dispatch_split<T, 2, 16>(p, dtype, split, stream);
  if (p.N <= 64) return shape == 1 ? dispatch_split<T, 8, 8>(p, dtype, split, stream) : dispatch_split<T, 4, 16>(p, dtype, split, stream);
  if (p.N <= 128) return dispatch_split<T, 8, 16>(p, dtype, split, stream);
  return kTmaNotEligible;
}

}  // namespace

int launch_scan_fwd_tma(const ScanFwdParams& p, int dtype, int tune, cudaStream_t stream) {
  return dtype == MAMBA_F32 ? dispatch_shape<float>(p, dtype, tune, stream) : dispatch_shape<__nv_bfloat16>(p, dtype, tune, stream);
}

}  // namespace mb
