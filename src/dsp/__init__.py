"""Top-level import names of the reference (``dsp.*``): aliases of sdrterm_b200.dsp, so that code
and tests written against the reference's ``src/`` layout run against this build unchanged
(``PYTHONPATH=src python -m sdrterm ...``, ``import dsp.demodulation``)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from sdrterm_b200.dsp import data_processor, demodulation, dsp_processor, vfo_processor  # noqa: E402

for _m in (data_processor, demodulation, dsp_processor, vfo_processor):
    sys.modules[__name__ + '.' + _m.__name__.rsplit('.', 1)[1]] = _m
