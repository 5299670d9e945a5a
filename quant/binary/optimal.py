"""Drop-in for the reference's quant/binary/optimal.py: re-exports ml_quant_b200.binary.optimal."""
from ml_quant_b200.binary.optimal import *  # noqa: F401,F403
from ml_quant_b200.binary import optimal as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
