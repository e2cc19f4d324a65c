"""vocoder_b200 - B200 (sm_100a) native implementation of the fish-vocoder generator forward path.

Host-side mirror of the reference module surface (``Generator.forward(mel[, template]) -> wav``, same
constructor kwargs and state_dict layout) over hand-written CUDA kernels in ``libfv_b200.so``.
"""
__version__ = "0.1.0"
