"""Drop-in ``quant`` package: ``quant.binary`` and ``quant.models`` resolve to the B200 implementation
(ml_quant_b200), everything else (``quant.common``, ``quant.data``, ``quant.utils`` ...) to the
reference checkout named by $ML_QUANT_REFERENCE (default /root/reference) when it is present, so the
reference's ``examples/*.py`` and tests run unmodified with this repository first on PYTHONPATH.

The package-level names of the reference (quant/__init__.py:10-29: ``__version__``, ``MetricDict``, ``Hook``) are
part of that surface -- ``quant.common.tasks`` / ``training`` import them -- and are provided here.
"""
import os as _os
from typing import Any, Dict, Optional

from typing_extensions import Protocol

_ref = _os.path.join(_os.environ.get('ML_QUANT_REFERENCE', '/root/reference'), 'quant')
if _os.path.isdir(_ref) and _ref not in __path__:
    __path__.append(_ref)

__version__ = '0.2.0'

try:                                            # the metrics classes live in the reference's orchestration layer
    from quant.common.metrics import Metric
except ImportError:                             # no reference checkout on this box: only the hot path is importable
    Metric = Any                                # type: ignore[misc,assignment]

MetricDict = Dict[str, Metric]                  # type: ignore[valid-type]


class Hook(Protocol):
    """Signature of a logging hook (called by the training / evaluation loops of quant.common.training)."""

    def __call__(self, epoch: int, global_step: int, log_interval: int = 10,
                 values_dict: Optional[dict] = None) -> None:
        ...
