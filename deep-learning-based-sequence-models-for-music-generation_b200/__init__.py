"""B200-native (sm_100a) Mamba hot path for composer-conditioned MIDI-token modelling.

Product code only: hand-written CUDA kernels behind a C-ABI (`libmamba_b200.so`, declared in
`include/mamba_b200.h`), `torch.autograd.Function` wrappers (`ops`), and the host-side mirror of
the reference's module interface (`models.mamba`, `configs`, `train`, `generate`, plus `synthetic` batches).  There is no CPU
fallback: every op raises if the CUDA library is missing or a tensor is not on a CUDA device.
"""
__version__ = "0.1.0"
