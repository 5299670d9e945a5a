"""Drop-in for the reference's quant/binary/binary_conv.py: re-exports ml_quant_b200.binary.binary_conv."""
from ml_quant_b200.binary.binary_conv import *  # noqa: F401,F403
from ml_quant_b200.binary import binary_conv as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
