"""Top-level import names of the reference (``misc.*``): aliases of sdrterm_b200.misc."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from sdrterm_b200.misc import (file_util, general_util, io_args, keyboard_interruptable_thread,  # noqa: E402
                               mappable_enum, read_file)

for _m in (file_util, general_util, io_args, keyboard_interruptable_thread, mappable_enum, read_file):
    sys.modules[__name__ + '.' + _m.__name__.rsplit('.', 1)[1]] = _m
