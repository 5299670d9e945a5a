"""Drop-in for the reference's quant/binary/quantization.py: re-exports ml_quant_b200.binary.quantization."""
from ml_quant_b200.binary.quantization import *  # noqa: F401,F403
from ml_quant_b200.binary import quantization as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
