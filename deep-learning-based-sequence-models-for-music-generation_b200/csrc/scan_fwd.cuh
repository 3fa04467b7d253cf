// scan_fwd.cuh — parameters shared by the selective-scan forward kernels (scan_fwd.cu, scan_fwd_tma.cu).
#pragma once
#include "common.cuh"

namespace mb {

struct ScanFwdParams {
  int B, L, D, N, N4, NS, NPT, nck, cki, flags;
  const void *u, *delta, *Bm, *Cm, *z;
  void* out;
  int64_t u_bs, u_ls, delta_bs, delta_ls, B_bs, B_ls, C_bs, C_ls, z_bs, z_ls, out_bs, out_ls;
  const float *A, *Dv, *dbias, *h_init;
  float *ckpt, *h_last;
  void* ypre;
  int64_t ypre_bs, ypre_ls;
  int vec_u, vec_delta, vec_z, vec_B, vec_C, vec_out, vec_ypre;
};

// TMA-staged forward (scan_fwd_tma.cu).  Returns MAMBA_OK after a launch, kTmaNotEligible when the problem cannot be
// described by tensor maps (unaligned base or strides, unsupported d_state) — the caller then uses scan_fwd.cu —
// or a negative MAMBA_E* code.
constexpr int kTmaNotEligible = 1;
int launch_scan_fwd_tma(const ScanFwdParams& p, int dtype, int tune, cudaStream_t stream);

}  // namespace mb
