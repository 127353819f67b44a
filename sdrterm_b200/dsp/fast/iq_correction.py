"""``dsp.fast.iq_correction``: the reference's compiled-plugin seam.  ``src/misc/read_file.py:58-63``
tries ``from dsp.fast.iq_correction import IQCorrection`` and, when the import works, routes every
chunk through ``IQCorrection(fs).correctIq(data, off)`` instead of its numba loop.  The reference's
plug-in is a Cython class (extra/src/iq_correction.pyx:37-70); this one has the same constructor,
properties and method and runs the recurrence on the B200 (``sdrb_correct_iq``: the serial
``z[i] -= off; off += z[i] * L`` evaluated as a scan of affine maps).  No CPU implementation."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ... import _native as nat


class IQCorrection:
    """``IQCorrection(fs, impedance=50)``; ``inductance = impedance / fs`` (pyx:41-44)."""

    def __init__(self, fs: int, impedance: int = 50, device: int = 0):
        self._fs = int(fs)
        self._impedance = int(impedance)
        self._device = int(device)
        self._update()

    def _update(self) -> None:
        self._inductance = self._impedance / self._fs

    @property
    def fs(self) -> int:
        return self._fs

    @fs.setter
    def fs(self, fs: int) -> None:
        self._fs = int(fs)
        self._update()

    @property
    def impedance(self) -> int:
        return self._impedance

    @impedance.setter
    def impedance(self, impedance: int) -> None:
        self._impedance = int(impedance)
        self._update()

    @property
    def inductance(self) -> float:
        return self._inductance

    def correctIq(self, data: np.ndarray, off: np.ndarray) -> None:
        """In place on ``data`` (1-D complex128, C-contiguous) and on ``off`` (one complex128)."""
        if not (isinstance(data, np.ndarray) and data.dtype == np.complex128 and data.flags.c_contiguous):
            raise TypeError('data must be a C-contiguous complex128 array')
        if not (isinstance(off, np.ndarray) and off.dtype == np.complex128 and off.size >= 1):
            raise TypeError('off must be a complex128 array holding the carried offset')
        st = (C.c_double * 2)(float(off.flat[0].real), float(off.flat[0].imag))
        nat.check(nat.lib().sdrb_correct_iq(self._device, data.ctypes.data, data.size, st, self._inductance))
        off.flat[0] = complex(st[0], st[1])
