"""Drop-in for quant/models/resnet.py."""
from ml_quant_b200.nets import QResNet, RegularBasicBlock, XnorBasicBlock, non_linearity_map  # noqa: F401
