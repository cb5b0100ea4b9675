"""jegal_b200 — B200-native (sm_100a) cross-modal scoring path for JEGAL embeddings.

Only the scoring hot path of Sindhu-Hegde/jegal lives here:
CUDA kernels + C ABI in ``csrc/`` (built to ``libjegal_b200.so``), device-tensor
wrappers in ``ops``, and the reference-compatible host functions in ``scoring``.
"""
__version__ = "0.1.0"
