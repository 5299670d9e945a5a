"""Drop-in for the reference's quant/binary/ste.py: re-exports ml_quant_b200.binary.ste."""
from ml_quant_b200.binary.ste import *  # noqa: F401,F403
from ml_quant_b200.binary import ste as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
