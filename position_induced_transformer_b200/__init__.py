"""B200-native position-attention for the Position-induced Transformer (PiT).

Public surface:
    position_induced_transformer_b200.pit      drop-in for the reference module ``pit.py``
    position_induced_transformer_b200.utils    drop-in for the reference module ``utils.py``
    position_induced_transformer_b200.posatt   functional fused position-attention (autograd aware)

The CUDA kernels live in ``csrc/`` and are reached through the C ABI of ``libpit_posatt.so``
(``include/pit_posatt.h``).  There is no CPU fallback.
"""
__version__ = "0.1.0"
