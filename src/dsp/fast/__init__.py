"""``dsp.fast``: alias of sdrterm_b200.dsp.fast (the reference installs its Cython plug-in here)."""
import sys

from sdrterm_b200.dsp.fast import iq_correction  # noqa: E402

sys.modules[__name__ + '.iq_correction'] = iq_correction
