"""``PYTHONPATH=src python -m sdrterm``: the reference's entry point name, this build's CLI."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from sdrterm_b200.sdrterm import buildParser, main, makeProcessor  # noqa: E402,F401

if __name__ == '__main__':
    sys.exit(main(lifecycle=True))
