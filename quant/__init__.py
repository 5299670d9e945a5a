"""Drop-in ``quant`` package: ``quant.binary`` and ``quant.models`` resolve to the B200 implementation
(ml_quant_b200), everything else (``quant.common``, ``quant.data``, ``quant.utils`` ...) to the
reference checkout named by $ML_QUANT_REFERENCE (default /root/reference) when it is present, so the
reference's ``examples/*.py`` run unmodified with this repository first on PYTHONPATH.
"""
import os as _os

_ref = _os.path.join(_os.environ.get('ML_QUANT_REFERENCE', '/root/reference'), 'quant')
if _os.path.isdir(_ref) and _ref not in __path__:
    __path__.append(_ref)
