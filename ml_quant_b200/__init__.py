"""ml_quant_b200 -- B200-native (sm_100a) implementation of apple/ml-quant's binary-quantized
inference path: the least-squares k-bit quantizer and the QuantConv2d forward that consumes its
codes, behind the reference's own module surface (``quant.binary.*``; see the ``quant`` shim package
at the repository root).  The arithmetic lives in hand-written CUDA behind a C ABI
(include/lsq_b200.h, csrc/); this package is the host-side mirror of the reference interface.
"""
__version__ = '0.1.0'
