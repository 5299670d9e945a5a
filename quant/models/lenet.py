"""Drop-in for quant/models/lenet.py."""
from ml_quant_b200.nets import QLeNet5  # noqa: F401
