// scan_fwd.cu — fused selective-scan forward for sm_100a.
//
// Replaces MambaBlock.selective_scan (reference models/mamba/__pycache__/simple_mamba.cpython-311.pyc
// @L310-333) fused with softplus(dt) (@L276), the D skip (@L331) and the z gate (@L241).
//
// Mapping (B200-first, not the reference's python loop over L):
//   * one CTA owns (batch b, 32 consecutive channels); lane <-> channel d, warp <-> slice of NPER
//     states n.  Every global access of u/delta/z/out is therefore a fully coalesced 128-byte row.
//   * the state h[NPER] lives in registers for the whole sequence; the CTA walks L sequentially in
//     stages of LCS timesteps that are double-buffered into shared memory with cp.async (16-byte
//     LDGSTS), so HBM latency is hidden behind the recurrence of the previous stage.
//   * per stage: a cooperative pre-pass turns raw delta into softplus(delta+bias) and delta*u once
//     per (t, d) (not once per state), the main loop does exp2/FMA per (t, d, n), and a cooperative
//     post-pass reduces the per-warp partial <h, C> sums, adds D*u, applies silu(z) and stores.
//   * every CK timesteps the state is checkpointed to HBM ([B, nck, N, D], coalesced) so the
//     backward can recompute the forward chunk by chunk.  deltaA / deltaB_u ([B, L, D, N] in the
//     reference) never exist in memory.
#include "common.cuh"

namespace mb {

constexpr int kDT = 32;   // channels per CTA (= warp width)
constexpr int kLCS = 32;  // timesteps per shared-memory stage
constexpr int kMaxWarps = 16;  // state slices per CTA (<= 512 threads, <= 128 registers each)

struct ScanFwdParams {
  int B, L, D, N, NS, NPT, nck, flags;
  const void *u, *delta, *Bm, *Cm, *z;
  void* out;
  int64_t u_bs, u_ls, delta_bs, delta_ls, B_bs, B_ls, C_bs, C_ls, z_bs, z_ls, out_bs, out_ls;
  const float *A, *Dv, *dbias, *h_init;
  float *ckpt, *h_last;
  int vec_u, vec_delta, vec_z, vec_B, vec_C, vec_out;
};

template <typename T>
struct FwdSmem {
  // raw (cp.async targets), per stage
  T *u, *dl, *z, *Bm, *Cm;
};

template <typename T, int NPER, int CK>
__global__ void __launch_bounds__(kMaxWarps * 32) scan_fwd_kernel(const ScanFwdParams p) {
  static_assert(kLCS % CK == 0, "stage must hold whole checkpoint chunks");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, s = tid >> 5;
  const int nthreads = blockDim.x;
  const int b = blockIdx.y;
  const int d0 = blockIdx.x * kDT;
  const int d = d0 + lane;
  const int dvalid = min(kDT, p.D - d0);
  const int NPT = p.NPT, NS = p.NS;
  const bool has_z = p.flags & MAMBA_FLAG_HAS_Z;

  // ---- carve shared memory ---------------------------------------------------------------
  const int raw_stage_elems = 3 * kLCS * kDT + 2 * kLCS * NPT;
  T* raw = reinterpret_cast<T*>(smem_raw);
  size_t raw_bytes = (size_t)2 * raw_stage_elems * sizeof(T);
  raw_bytes = (raw_bytes + 15) & ~(size_t)15;
  float* wdl = reinterpret_cast<float*>(smem_raw + raw_bytes);
  float* wdu = wdl + kLCS * kDT;
  float* wB = wdu + kLCS * kDT;
  float* wC = wB + kLCS * NPT;
  float* ypart = wC + kLCS * NPT;  // [NS][kLCS][kDT]

  auto stage = [&](int st) {
    FwdSmem<T> r;
    T* base = raw + (size_t)st * raw_stage_elems;
    r.u = base;
    r.dl = r.u + kLCS * kDT;
    r.z = r.dl + kLCS * kDT;
    r.Bm = r.z + kLCS * kDT;
    r.Cm = r.Bm + kLCS * NPT;
    return r;
  };

  const T* gu = static_cast<const T*>(p.u) + (int64_t)b * p.u_bs + d0;
  const T* gdl = static_cast<const T*>(p.delta) + (int64_t)b * p.delta_bs + d0;
  const T* gz = has_z ? static_cast<const T*>(p.z) + (int64_t)b * p.z_bs + d0 : nullptr;
  const T* gB = static_cast<const T*>(p.Bm) + (int64_t)b * p.B_bs;
  const T* gC = static_cast<const T*>(p.Cm) + (int64_t)b * p.C_bs;
  T* gout = static_cast<T*>(p.out) + (int64_t)b * p.out_bs + d0;

  auto issue_loads = [&](int st, int c) {
    FwdSmem<T> r = stage(st);
    const int t0 = c * kLCS;
    const int rows_valid = min(kLCS, p.L - t0);
    load_tile_async<T>(r.u, kDT, gu + (int64_t)t0 * p.u_ls, p.u_ls, kLCS, rows_valid, dvalid, p.vec_u, tid, nthreads);
    load_tile_async<T>(r.dl, kDT, gdl + (int64_t)t0 * p.delta_ls, p.delta_ls, kLCS, rows_valid, dvalid, p.vec_delta,
                       tid, nthreads);
    if (has_z)
      load_tile_async<T>(r.z, kDT, gz + (int64_t)t0 * p.z_ls, p.z_ls, kLCS, rows_valid, dvalid, p.vec_z, tid,
                         nthreads);
    load_tile_async<T>(r.Bm, NPT, gB + (int64_t)t0 * p.B_ls, p.B_ls, kLCS, rows_valid, p.N, p.vec_B, tid, nthreads);
    load_tile_async<T>(r.Cm, NPT, gC + (int64_t)t0 * p.C_ls, p.C_ls, kLCS, rows_valid, p.N, p.vec_C, tid, nthreads);
  };

  // ---- per-thread constants and state ----------------------------------------------------
  float A2[NPER], h[NPER];
#pragma unroll
  for (int j = 0; j < NPER; ++j) {
    const int n = s * NPER + j;
    const bool ok = (d < p.D) && (n < p.N);
    A2[j] = ok ? p.A[(int64_t)d * p.N + n] * kLog2e : 0.f;
    h[j] = (ok && p.h_init) ? p.h_init[((int64_t)b * p.D + d) * p.N + n] : 0.f;
  }
  const float bias_d = ((p.flags & MAMBA_FLAG_HAS_DELTA_BIAS) && d < p.D) ? p.dbias[d] : 0.f;
  const float D_d = ((p.flags & MAMBA_FLAG_HAS_D) && d < p.D) ? p.Dv[d] : 0.f;
  const bool do_softplus = p.flags & MAMBA_FLAG_DELTA_SOFTPLUS;

  const int nst = (p.L + kLCS - 1) / kLCS;
  issue_loads(0, 0);
  cp_async_commit();

  for (int c = 0; c < nst; ++c) {
    if (c + 1 < nst) issue_loads((c + 1) & 1, c + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    FwdSmem<T> r = stage(c & 1);
    const int t0 = c * kLCS;
    const int rows_valid = min(kLCS, p.L - t0);

    // ---- pre-pass: discretisation inputs, once per (t, d) ---------------------------------
    for (int i = tid; i < kLCS * kDT; i += nthreads) {  // i % 32 == lane
      const int t = i >> 5;
      float dl = 0.f, du = 0.f;
      if (t < rows_valid) {
        dl = IO<T>::cvt(r.dl[i]) + bias_d;
        if (do_softplus) dl = softplus_f(dl);
        du = dl * IO<T>::cvt(r.u[i]);
      }
      wdl[i] = dl;
      wdu[i] = du;
    }
    for (int i = tid; i < kLCS * NPT; i += nthreads) {
      wB[i] = IO<T>::cvt(r.Bm[i]);
      wC[i] = IO<T>::cvt(r.Cm[i]);
    }
    __syncthreads();

    // ---- main loop: the recurrence, h in registers ----------------------------------------
#pragma unroll 1
    for (int tc = 0; tc < kLCS; tc += CK) {
      if (tc >= rows_valid) break;
#pragma unroll 4
      for (int tt = 0; tt < CK; ++tt) {
        const int t = tc + tt;
        const float dl = wdl[t * kDT + lane];
        const float du = wdu[t * kDT + lane];
        float Bv[NPER], Cv[NPER];
        const float4* b4 = reinterpret_cast<const float4*>(wB + t * NPT + s * NPER);
        const float4* c4 = reinterpret_cast<const float4*>(wC + t * NPT + s * NPER);
#pragma unroll
        for (int q = 0; q < NPER / 4; ++q) {
          float4 bb = b4[q], cc = c4[q];
          Bv[4 * q] = bb.x, Bv[4 * q + 1] = bb.y, Bv[4 * q + 2] = bb.z, Bv[4 * q + 3] = bb.w;
          Cv[4 * q] = cc.x, Cv[4 * q + 1] = cc.y, Cv[4 * q + 2] = cc.z, Cv[4 * q + 3] = cc.w;
        }
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < NPER; ++j) {
          const float a = ex2_approx(dl * A2[j]);
          h[j] = fmaf(a, h[j], du * Bv[j]);
          acc = fmaf(h[j], Cv[j], acc);
        }
        ypart[(s * kLCS + t) * kDT + lane] = acc;
      }
      const int tg_next = t0 + tc + CK;  // h now = state at the start of checkpoint chunk tg_next / CK
      if (p.ckpt != nullptr && tg_next < p.L && d < p.D) {
        float* ck = p.ckpt + (((int64_t)b * p.nck + tg_next / CK) * p.N) * p.D + d;
#pragma unroll
        for (int j = 0; j < NPER; ++j) {
          const int n = s * NPER + j;
          if (n < p.N) ck[(int64_t)n * p.D] = h[j];
        }
      }
    }
    __syncthreads();

    // ---- post-pass: reduce over state slices, D skip, z gate, coalesced store --------------
    for (int i = tid; i < kLCS * kDT; i += nthreads) {  // i % 32 == lane
      const int t = i >> 5;
      if (t < rows_valid && d < p.D) {
        float y = 0.f;
        for (int w = 0; w < NS; ++w) y += ypart[w * kLCS * kDT + i];
        y = fmaf(D_d, IO<T>::cvt(r.u[i]), y);
        if (has_z) y *= silu_f(IO<T>::cvt(r.z[i]));
        IO<T>::st(gout + (int64_t)(t0 + t) * p.out_ls + lane, y);
      }
    }
    __syncthreads();
  }

  if (p.h_last != nullptr && d < p.D) {
#pragma unroll
    for (int j = 0; j < NPER; ++j) {
      const int n = s * NPER + j;
      if (n < p.N) p.h_last[((int64_t)b * p.D + d) * p.N + n] = h[j];
    }
  }
}

static size_t fwd_smem_bytes(size_t elt, int NS, int NPT) {
  size_t raw = (size_t)2 * (3 * kLCS * kDT + 2 * kLCS * NPT) * elt;
  raw = (raw + 15) & ~(size_t)15;
  size_t work = (size_t)4 * (2 * kLCS * kDT + 2 * kLCS * NPT + (size_t)NS * kLCS * kDT);
  return raw + work;
}

template <typename T, int NPER, int CK>
static int launch_fwd(const ScanFwdParams& p, cudaStream_t stream) {
  const size_t smem = fwd_smem_bytes(sizeof(T), p.NS, p.NPT);
  if (smem > 227 * 1024) return set_error(MAMBA_ESIZE, "scan_fwd: d_state %d needs %zu B of shared memory", p.N, smem);
  auto kern = scan_fwd_kernel<T, NPER, CK>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error(MAMBA_ELAUNCH, "scan_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  dim3 grid(ceil_div(p.D, kDT), p.B);
  kern<<<grid, p.NS * 32, smem, stream>>>(p);
  count_launch();
  return check_launch("scan_fwd");
}

template <typename T, int NPER>
static int dispatch_ck(const ScanFwdParams& p, int chunk, cudaStream_t stream) {
  switch (chunk) {
    case 8: return launch_fwd<T, NPER, 8>(p, stream);
    case 16: return launch_fwd<T, NPER, 16>(p, stream);
    case 32: return launch_fwd<T, NPER, 32>(p, stream);
  }
  return set_error(MAMBA_EINVAL, "scan_fwd: chunk must be 8, 16 or 32 (got %d)", chunk);
}

template <typename T>
static int dispatch_nper(ScanFwdParams& p, int nper, int chunk, cudaStream_t stream) {
  p.NS = ceil_div(p.N, nper);
  const int np = p.NS * nper;
  p.NPT = (np + 7) & ~7;
  if (p.NS > kMaxWarps) return set_error(MAMBA_ESIZE, "scan_fwd: d_state %d too large for %d states/thread", p.N, nper);
  switch (nper) {
    case 4: return dispatch_ck<T, 4>(p, chunk, stream);
    case 8: return dispatch_ck<T, 8>(p, chunk, stream);
    case 16: return dispatch_ck<T, 16>(p, chunk, stream);
  }
  return set_error(MAMBA_EINVAL, "scan_fwd: variant must be 0, 4, 8 or 16 (got %d)", nper);
}

static bool vec_ok(const void* ptr, int64_t bs, int64_t ls, size_t elt) {
  return aligned16(ptr) && (bs * elt) % 16 == 0 && (ls * elt) % 16 == 0;
}

}  // namespace mb

extern "C" size_t mamba_scan_ckpt_elems(int batch, int seqlen, int dim, int dstate, int chunk) {
  if (batch <= 0 || seqlen <= 0 || dim <= 0 || dstate <= 0 || chunk <= 0) return 0;
  return (size_t)batch * mb::ceil_div(seqlen, chunk) * dstate * dim;
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

  ScanFwdParams p{};
  p.B = a->batch, p.L = a->seqlen, p.D = a->dim, p.N = a->dstate, p.flags = a->flags;
  const int chunk = a->ckpt ? a->chunk : 16;
  p.nck = ceil_div(p.L, chunk);
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

  int nper = a->variant;
  if (nper == 0) nper = p.N >= 64 ? 16 : (p.N >= 32 ? 8 : 4);
  while (nper < 16 && ceil_div(p.N, nper) > kMaxWarps) nper *= 2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return a->dtype == MAMBA_F32 ? dispatch_nper<float>(p, nper, chunk, st)
                               : dispatch_nper<__nv_bfloat16>(p, nper, chunk, st);
}
