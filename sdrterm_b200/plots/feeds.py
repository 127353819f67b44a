"""What the reference's plot consumers compute per received chunk, as device calls.

* ``powerSpectrum`` / ``SpectrumFeed.update`` -- ``SpectrumAnalyzerPlot.update``
  (src/plots/spectrum_analyzer_plot.py:75-92): ``shiftFreq(y, shift, y)``,
  ``amp = abs(fftshift(fftn(y, norm='forward')))``, ``amp = log10(amp*amp)``,
  ``freq = fftshift(fftfreq(n, dt))``.
* ``stftDb`` / ``WaterfallFeed.update`` -- ``WaterfallPlot.update`` (src/plots/waterfall_plot.py:
  44-51, 90-99): ``10*log10(abs(ShortTimeFFT.stft(y)))`` with the Kaiser(5) window of 256 samples,
  hop 128, ``mfft=1024``, ``fft_mode='centered'``, ``scale_to='magnitude'``, ``phase_shift=None``.

The NCO vector is formed on the host exactly as ``AbstractPlot.receiveData`` forms it
(src/plots/abstract_plot.py:75,151-152: ``exp(-2j*pi*(offset/fs) * arange(n))``).  One ctypes call
per chunk (``sdrb_power_spectrum`` / ``sdrb_stft_db``); no CPU fallback."""
from __future__ import annotations

import numpy as np

from .. import _native as nat


def _c128(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.complex128)


def powerSpectrum(y, shift=None, device: int = 0) -> np.ndarray:
    """``log10(abs(fftshift(fftn(y*shift, norm='forward')))**2)`` of every row of ``y`` (one row or
    a (rows, n) array; n a power of two).  ``shift``: n complex values applied to every row."""
    z = _c128(y)
    rows = z.reshape(-1, z.shape[-1])
    n = rows.shape[1]
    sh = None
    if shift is not None:
        sh = _c128(shift).reshape(-1)
        if sh.size != n:
            raise ValueError(f'shift has {sh.size} values, the rows have {n}')
    out = np.empty(rows.shape, dtype=np.float64)
    nat.check(nat.lib().sdrb_power_spectrum(device, rows.ctypes.data, sh.ctypes.data if sh is not None else None,
                                            n, rows.shape[0], out.ctypes.data))
    return out.reshape(z.shape)


def stftDb(y, win, hop: int, mfft: int, p_num: int, shift=None, device: int = 0) -> np.ndarray:
    """``10*log10(abs(S))`` for the centred short-time transform S[q, p] of one row ``y``: slice p
    covers samples ``p*hop - len(win)//2 + (0..len(win)-1)`` (zero outside), times ``win``, zero-
    padded to ``mfft``.  Returns (mfft, p_num)."""
    z = _c128(y).reshape(-1)
    w = np.ascontiguousarray(win, dtype=np.float64)
    sh = None
    if shift is not None:
        sh = _c128(shift).reshape(-1)
        if sh.size != z.size:
            raise ValueError(f'shift has {sh.size} values, the row has {z.size}')
    out = np.empty((mfft, p_num), dtype=np.float64)
    nat.check(nat.lib().sdrb_stft_db(device, z.ctypes.data, sh.ctypes.data if sh is not None else None, z.size,
                                     w.ctypes.data, w.size, hop, mfft, p_num, out.ctypes.data))
    return out


class _Feed:
    def __init__(self, fs: int, center: int = 0, tuned: int = 0, device: int = 0):
        if fs is None:
            raise ValueError('a feed cannot be used without a sampling rate: fs')
        self.fs, self.offset, self.tuned, self.device = fs, center, tuned, device
        self.dt = 1 / fs
        self.nyquistFs = fs >> 1
        self._omega = -2j * np.pi * (self.offset / self.fs)           # abstract_plot.py:75
        self._shift = None

    def _shiftFor(self, n: int) -> np.ndarray:
        if self._shift is None or self._shift.size != n:
            self._shift = np.exp(self._omega * np.arange(n))          # abstract_plot.py:151-152
        return self._shift


class SpectrumFeed(_Feed):
    """``update(y) -> (freq, amp)``: the two arrays ``SpectrumAnalyzerPlot`` hands to ``setData``."""

    def update(self, y):
        z = _c128(y).reshape(-1)
        amp = powerSpectrum(z, self._shiftFor(z.size), self.device)
        from scipy.fft import fftfreq, fftshift
        return fftshift(fftfreq(amp.size, self.dt)), amp


class WaterfallFeed(_Feed):
    """``update(y) -> image``: the (NFFT, n // NOOVERLAP + 1) array ``WaterfallPlot`` shows."""
    _NPERSEG = 256
    _NOOVERLAP = _NPERSEG >> 1
    _NFFT = 1024

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        from scipy.signal import ShortTimeFFT
        self._SFT = ShortTimeFFT.from_window(('kaiser', 5), self.fs, self._NPERSEG, self._NOOVERLAP, mfft=self._NFFT,
                                             fft_mode='centered', scale_to='magnitude', phase_shift=None)

    def update(self, y):
        z = _c128(y).reshape(-1)
        s = self._SFT
        if s.p_min != 0 or s.m_num_mid != self._NPERSEG // 2:
            raise RuntimeError('unexpected ShortTimeFFT slice geometry')
        return stftDb(z, s.win, s.hop, self._NFFT, s.p_max(z.size), self._shiftFor(z.size), self.device)
