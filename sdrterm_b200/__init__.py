"""sdrterm_b200: B200-native (sm_100a) implementation of peads/sdrterm's streamed IQ demodulation
chain behind the reference's own processor API.

Layout:
  plan.py      host-side tables (filter design by SciPy, modal block form in extended precision)
  engine.py    one C-ABI handle per processor (ctypes)
  _native.py   loader/builder of libsdrterm_b200.so (csrc/*.cu, include/sdrterm_b200.h)
  dsp/, misc/, sdrterm.py   drop-in mirrors of the reference's src/dsp, src/misc, src/sdrterm.py
"""
__version__ = '0.1.0'
