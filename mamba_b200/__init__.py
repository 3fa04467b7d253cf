"""Importable alias of the product package.

The product lives in the directory `deep-learning-based-sequence-models-for-music-generation_b200/`
whose name (fixed by the project layout) is not a valid Python identifier; this two-line alias puts
that directory on `mamba_b200.__path__` so that `import mamba_b200.ops` etc. work.
"""
from pathlib import Path as _Path

_impl = _Path(__file__).resolve().parent.parent / "deep-learning-based-sequence-models-for-music-generation_b200"
__path__.insert(0, str(_impl))  # noqa: F821  (module attribute)
exec(compile((_impl / "__init__.py").read_text(), str(_impl / "__init__.py"), "exec"))
