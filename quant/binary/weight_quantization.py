"""Drop-in for the reference's quant/binary/weight_quantization.py: re-exports ml_quant_b200.binary.weight_quantization."""
from ml_quant_b200.binary.weight_quantization import *  # noqa: F401,F403
from ml_quant_b200.binary import weight_quantization as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
