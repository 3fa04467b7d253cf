"""Mirror of `configs.mamba` (reference configs/mamba/config.yaml:1-6 plus the three keys that only the
older yaml carries, configs/mamba/.ipynb_checkpoints/config-checkpoint.yaml:7-9, and that the
pure-PyTorch model reads)."""
from types import SimpleNamespace

config = SimpleNamespace(model_values=SimpleNamespace(
    d_model=1024, n_layer=10, d_state=64, expand=2, d_conv=4,
    pad_vocab_size_multiple=8, conv_bias=True, bias=False))
