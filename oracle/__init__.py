"""ORACLE — test infrastructure only.

CPU restatements of the reference's algorithm for the Mamba hot path.  Nothing in the product package
(`deep-learning-based-sequence-models-for-music-generation_b200/`, alias `mamba_b200`) imports this; only
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs do.
"""
