"""CPU oracle for sdrterm's streamed IQ demodulation chain -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm may
import this module.  The product path (``sdrterm_b200``) never does; it fails loudly when the CUDA
library is missing.

This is a restatement (numpy + the plain-C helpers in ``sdr_oracle.c``) of what the reference
computes on its hot path, function by function:

=====================  ==========================================================================
oracle function        reference (relative to /root/reference)
=====================  ==========================================================================
``struct_dtype``       src/misc/file_util.py:46-111 (DataType, parseRawType, ':' => big-endian),
                       src/misc/read_file.py:48-51 (-X swap, [('re',T),('im',T)])
``decode``             src/misc/read_file.py:100-101,124  (z = y['re'] + 1j*y['im'])
``generate_domain``    src/misc/read_file.py:177-196
``normalize``          src/misc/read_file.py:82-96
``correct_iq``         src/misc/read_file.py:65-77; extra/src/iq_correction.pyx:46-52
``nco_table``          src/dsp/dsp_processor.py:185-187; src/dsp/vfo_processor.py:42-48,71-74
``decimate``           src/dsp/dsp_processor.py:147 -> scipy.signal.decimate (SciPy 1.18.1:
                       _signaltools.py:5206-5369; sosfiltfilt :5091-5203; sosfilt_zi :4464-4545)
``fm_demod``           src/dsp/demodulation.py:25-38 (+ scipy.signal.resample :3586-3881)
``am_demod`` etc.      src/dsp/demodulation.py:41-68
``output_filter``      src/dsp/dsp_processor.py:32-45,116-128,149
``frame``              src/dsp/dsp_processor.py:162; src/dsp/vfo_processor.py:84
``Chain``              src/misc/read_file.py:119-125 + src/dsp/dsp_processor.py:140-183
``plot_shift``         src/plots/abstract_plot.py:75,151-152
``power_spectrum``     src/plots/spectrum_analyzer_plot.py:75-82
``stft_db``            src/plots/waterfall_plot.py:44-51,97-99 (+ scipy.signal.ShortTimeFFT.stft,
                       restated slice by slice; pinned against SciPy's own in test_oracle_pins)
=====================  ==========================================================================

Third-party arithmetic: the decimator and both IIR filters live in SciPy (pyproject.toml:12-18,
unpinned; 1.18.1 in this image).  Filter *design* (cheby1 / ellip / sosfilt_zi) is setup and is
taken from SciPy directly, exactly as the reference does; the per-sample recurrences are restated
(here and in C) and pinned bit-exactly against scipy.signal.decimate / sosfilt by
tests/test_oracle_pins.py, and the whole chain is pinned against the reference itself run
in-process (tests/golden/make_golden.py -> tests/golden/*.npz).

Documented quirk switches: see ``Chain``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np
from scipy import signal as _sig

CHUNK_BYTES = 131072  # src/misc/read_file.py:38 (readSize)
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile sdr_oracle.c -> oracle/_build/liboracle.so (gcc, no FMA contraction)."""
    so = os.path.join(_HERE, '_build', 'liboracle.so')
    src = os.path.join(_HERE, 'sdr_oracle.c')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(['make', '-C', _HERE, '-B' if force else '-s'], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, '_build', 'liboracle.so')
        if not os.path.exists(so):
            so = build()
        L = ctypes.CDLL(so)
        dp = ctypes.POINTER(ctypes.c_double)
        L.orc_decode.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_char, ctypes.c_int, dp]
        L.orc_correct_iq.argtypes = [dp, ctypes.c_long, ctypes.c_double, dp]
        L.orc_normalize.argtypes = [dp, ctypes.c_long, ctypes.c_double, ctypes.c_double]
        L.orc_sosfilt_r.argtypes = [dp, ctypes.c_int, dp, ctypes.c_long, dp]
        L.orc_sosfilt_c.argtypes = [dp, ctypes.c_int, dp, ctypes.c_long, dp]
        L.orc_decimate.argtypes = [dp, ctypes.c_int, dp, ctypes.c_int, dp, ctypes.c_long,
                                   ctypes.c_int, dp, dp]
        L.orc_decimate.restype = ctypes.c_long
        L.orc_batch.argtypes = [dp, ctypes.c_int, dp, ctypes.c_int, dp, dp, ctypes.c_int,
                                ctypes.c_long, ctypes.c_int, ctypes.c_long, dp, ctypes.c_int]
        L.orc_batch.restype = ctypes.c_long
        L.orc_fm_pairs.argtypes = [dp, ctypes.c_long, dp]
        L.orc_am.argtypes = [dp, ctypes.c_long, dp]
        L.orc_max_threads.restype = ctypes.c_int
        L.orc_decode_batch.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_char, ctypes.c_int,
                                       dp, ctypes.c_int]
        L.orc_fm_pairs_rows.argtypes = [dp, ctypes.c_long, ctypes.c_long, dp, ctypes.c_int]
        L.orc_sosfilt_rows.argtypes = [dp, ctypes.c_int, dp, ctypes.c_long, ctypes.c_long,
                                       ctypes.c_int]
        _LIB = L
    return _LIB


def _dp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def max_threads() -> int:
    return int(_lib().orc_max_threads())


# ----------------------------------------------------------------------------- input layout
_BASE = {'b': '|i1', 'B': '|u1', 'h': '=i2', 'H': '=u2', 'i': '=i4', 'I': '=u4', 'f': '=f4',
         'd': '=f8'}


def sample_dtype(enc: str, big_endian: bool | None = None, swap: bool = False) -> np.dtype:
    """file_util.py:46-111 + read_file.py:48-49.  ``big_endian=None``: native order (raw file);
    True/False: forced by a RIFX/RIFF header or a host:port input.  ``swap`` is -X."""
    dt = np.dtype(_BASE[enc])
    if big_endian is not None:
        dt = dt.newbyteorder('>' if big_endian else '<')
    if swap:
        # read_file.py:49: '<' if byteorder == '>' else '>'  (numpy reports native as '=', and
        # single-byte types as '|': both become '>' there, which is a no-op for 1-byte types)
        dt = dt.newbyteorder('<' if dt.byteorder == '>' else '>')
    return dt


def struct_dtype(dt: np.dtype) -> np.dtype:
    return np.dtype([('re', dt), ('im', dt)])


def needs_swap(dt: np.dtype) -> bool:
    """True when the stored byte order differs from this host's."""
    if dt.itemsize == 1:
        return False
    native_little = np.little_endian
    bo = dt.byteorder
    if bo == '=':
        return False
    return (bo == '>') == native_little


def decode(raw: np.ndarray | bytes, dt: np.dtype, upcast_f32: bool = True) -> np.ndarray:
    """read_file.py:100-101: view as [('re',T),('im',T)], z = re + 1j*im (complex128).
    For float32 input the reference's expression yields complex64 and the whole chain runs in
    single precision; ``upcast_f32`` (default) widens exactly to complex128
    instead -- the documented divergence."""
    y = np.frombuffer(raw, dtype=struct_dtype(dt))
    z = y['re'] + 1j * y['im']
    if z.dtype != np.complex128 and upcast_f32:
        z = z.astype(np.complex128)
    return z


def generate_domain(char: str):
    """read_file.py:177-196 (same table, same quirk: 'H' uses xmax = 65536)."""
    table = {'B': (0, 255), 'h': (-32768, 32767), 'b': (-128, 127),
             'i': (-2147483648, 2147483647), 'H': (0, 65536), 'I': (0, 4294967295),
             'L': (0, 18446744073709551615), 'l': (-9223372036854775808, 9223372036854775807)}
    if char not in table:
        return None
    xmin, xmax = table[char]
    return xmin, 1 / (-xmin + xmax)


def normalize(z: np.ndarray, char: str) -> np.ndarray:
    """read_file.py:88-96 in place: 1.6*(z - xmin)*k - 0.8 with real constants on complex z."""
    dom = generate_domain(char)
    if dom is None:
        return z
    xmin, k = dom
    _lib().orc_normalize(_dp(z.view(np.float64)), z.size, float(xmin), float(k))
    return z


def correct_iq(z: np.ndarray, off: np.ndarray, fs: int, impedance: int = 50) -> None:
    """read_file.py:65-77 in place; ``off`` is a 1-element complex128 array carried across
    chunks (read_file.py:53)."""
    L = impedance / fs
    _lib().orc_correct_iq(_dp(z.view(np.float64)), z.size, L, _dp(off.view(np.float64)))


def correct_iq_py(z: np.ndarray, off: np.ndarray, fs: int, impedance: int = 50) -> None:
    """Same recurrence as a plain Python loop (used to pin the C helper on small inputs)."""
    L = impedance / fs
    o = complex(off[0])
    for i in range(z.shape[0]):
        z[i] = z[i] - o
        o += z[i] * L
    off[0] = o


# ----------------------------------------------------------------------------- NCO
def nco_rows(center: int, vfos: str | None) -> list[int]:
    """vfo_processor.py:42-46: rows = [int(v)+center for v in vfos] + [center]; standard mode
    has the single row [center]."""
    if vfos is None:
        return [center]
    return [int(x) + center for x in vfos.split(',') if x is not None] + [center]


def nco_table(freqs, fs: int, n: int, simo: bool) -> np.ndarray | None:
    """dsp_processor.py:185-187 (standard: exp(-2j*pi*(fc/fs)*arange(n)), None when fc == 0)
    and vfo_processor.py:48,71-74 (SIMO: w = -2j*pi*(vfos/fs); exp(w*arange(n)) per row).
    Expression order kept so the rounded phase products are identical."""
    if not simo:
        fc = freqs[0]
        if not fc:
            return None
        return np.array([np.exp(-2j * np.pi * (fc / fs) * np.arange(n))])
    omega = -2j * np.pi * (np.array(freqs) / fs)
    out = np.ones((len(freqs), n), dtype=np.complex128)
    for i, w in enumerate(omega):
        out[i][:] = np.exp(w * np.arange(n))
    return out


# ----------------------------------------------------------------------------- filters
def decimation_filter(q: int):
    """scipy.signal.decimate's IIR default: cheby1(8, 0.05, 0.8/q) SOS, sosfilt_zi, edge = 27."""
    sos = np.ascontiguousarray(_sig.cheby1(8, 0.05, 0.8 / q, output='sos'), dtype=np.float64)
    zi = np.ascontiguousarray(_sig.sosfilt_zi(sos), dtype=np.float64)
    nsec = sos.shape[0]
    ntaps = 2 * nsec + 1 - min(int((sos[:, 2] == 0).sum()), int((sos[:, 5] == 0).sum()))
    return sos, zi, 3 * ntaps


def output_filter_sos(decimated_fs: int, omega_out: int) -> np.ndarray:
    """dsp_processor.py:39-45,116-128: ellip(3, 1, 30, Wn=omegaOut, lowpass, fs=decimatedFs)."""
    return np.ascontiguousarray(
        _sig.ellip(3, 1, 30, omega_out, btype='lowpass', analog=False, output='sos',
                   fs=decimated_fs), dtype=np.float64)


def decimate(x: np.ndarray, q: int, filt=None) -> np.ndarray:
    """Restated scipy.signal.decimate(x, q) (zero-phase IIR) on the last axis of complex128 x."""
    sos, zi, edge = filt if filt is not None else decimation_filter(q)
    x2 = np.ascontiguousarray(x, dtype=np.complex128).reshape(-1, x.shape[-1])
    n = x2.shape[1]
    if n <= edge:
        raise ValueError(f'The length of the input vector x must be greater than padlen, '
                         f'which is {edge}.')
    M = (n + q - 1) // q
    out = np.empty((x2.shape[0], M), dtype=np.complex128)
    work = np.empty(2 * (n + 2 * edge), dtype=np.float64)
    L = _lib()
    for r in range(x2.shape[0]):
        L.orc_decimate(_dp(sos), sos.shape[0], _dp(zi), edge, _dp(x2[r].view(np.float64)), n, q,
                       _dp(work), _dp(out[r].view(np.float64)))
    return out.reshape(x.shape[:-1] + (M,))


def decimate_py(x: np.ndarray, q: int) -> np.ndarray:
    """Pure-numpy/Python restatement of the same thing (slow; pins the C helper on small n)."""
    sos, zi, edge = decimation_filter(q)
    x = np.asarray(x, dtype=np.complex128)
    ext = np.concatenate((2 * x[0] - x[edge:0:-1], x, 2 * x[-1] - x[-2:-(edge + 2):-1]))

    def sosfilt(v, z):
        z = z.copy()
        y = np.empty_like(v)
        for i in range(v.shape[0]):
            xc = v[i]
            for s in range(sos.shape[0]):
                b0, b1, b2, _, a1, a2 = sos[s]
                xn = b0 * xc + z[s, 0]
                z[s, 0] = (b1 * xc - a1 * xn) + z[s, 1]
                z[s, 1] = b2 * xc - a2 * xn
                xc = xn
            y[i] = xc
        return y

    y = sosfilt(ext, zi.astype(np.complex128) * ext[0])
    y = sosfilt(y[::-1], zi.astype(np.complex128) * y[-1])[::-1]
    return y[edge:-edge][::q]


def sosfilt_real(sos: np.ndarray, x: np.ndarray) -> np.ndarray:
    """scipy.signal.sosfilt(sos, x) along the last axis, zero initial state, real x."""
    x2 = np.array(x, dtype=np.float64, copy=True).reshape(-1, x.shape[-1])
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    for r in range(x2.shape[0]):
        zi = np.zeros(2 * sos.shape[0])
        _lib().orc_sosfilt_r(_dp(sos), sos.shape[0], _dp(x2[r]), x2.shape[1], _dp(zi))
    return x2.reshape(x.shape)


# ----------------------------------------------------------------------------- demodulation
def resample_2x(r: np.ndarray, num: int) -> np.ndarray:
    """scipy.signal.resample(r, num) for real r, restated (rfft -> unpaired-bin fix -> irfft)."""
    n_x = r.shape[-1]
    s_fac = n_x / num
    m = min(num, n_x)
    m2 = m // 2 + 1
    X = np.fft.rfft(r)[..., :m2].copy()
    if m % 2 == 0 and num != n_x:
        X[..., m // 2] *= 2 if num < n_x else 0.5
    return np.fft.irfft(X / s_fac, n=num)


def fm_demod(y: np.ndarray) -> np.ndarray:
    """demodulation.py:25-38: angle(y[2k]*conj(y[2k+1])) on non-overlapping pairs, then
    resample(res[:M>>1], M) per row.  Odd M (only reachable with the fixed-allocation oracle,
    SURVEY 8-Q1) uses the M>>1 complete pairs."""
    y2 = np.ascontiguousarray(y, dtype=np.complex128).reshape(-1, y.shape[-1])
    M = y2.shape[1]
    out = np.empty((y2.shape[0], M), dtype=np.float64)
    for r in range(y2.shape[0]):
        res = np.zeros(M >> 1, dtype=np.float64)
        _lib().orc_fm_pairs(_dp(y2[r].view(np.float64)), (M >> 1) * 2, _dp(res))
        out[r] = resample_2x(res, M)
    return out.reshape(y.shape)


def am_demod(y: np.ndarray) -> np.ndarray:
    """demodulation.py:41-48: abs(square(z))."""
    return np.abs(np.square(y))


def real_output(y: np.ndarray) -> np.ndarray:
    return np.real(y).copy()


def imag_output(y: np.ndarray) -> np.ndarray:
    return np.imag(y).copy()


def frame(z: np.ndarray, simo: bool) -> list[bytes]:
    """dsp_processor.py:162 (one native-endian block) / vfo_processor.py:84 (one big-endian
    block per row)."""
    if not simo:
        return [np.ascontiguousarray(z, dtype='=f8').tobytes()]
    return [np.ascontiguousarray(row, dtype='>f8').tobytes() for row in z]


# ----------------------------------------------------------------------------- plot feeds
def plot_shift(center: int, fs: int, n: int) -> np.ndarray:
    """abstract_plot.py:75 ``_omega = -2j*pi*(offset/fs)`` and :151-152 ``exp(_omega * arange(n))``."""
    return np.exp(-2j * np.pi * (center / fs) * np.arange(n))


def power_spectrum(y: np.ndarray, shift: np.ndarray | None) -> np.ndarray:
    """spectrum_analyzer_plot.py:75-82: ``shiftFreq(y, shift, y)``;
    ``amp = abs(fftshift(fftn(y, norm='forward')))``; ``amp = log10(amp * amp)`` on the (1, n) array."""
    from scipy.fft import fftn, fftshift
    z = np.array([np.asarray(y, dtype=np.complex128).reshape(-1)])
    if shift is not None:
        z = z * np.asarray(shift).reshape(1, -1)
    amp = abs(fftshift(fftn(z, norm='forward')))
    with np.errstate(divide='ignore'):
        return np.log10(amp * amp)[0]


def stft_geometry(n: int, nperseg: int = 256, hop: int = 128):
    """(first sample of slice 0, number of slices) of ``ShortTimeFFT.stft`` on an n-sample row for a
    window whose half length is a multiple of the hop (the reference's 256 / 128): slice p starts
    at p*hop - nperseg//2; p runs from 0 to the last slice that still begins inside the row
    (SciPy's p_min .. p_max(n)-1; checked against SciPy in tests/test_oracle_pins.py)."""
    mid = nperseg // 2
    p_max = -(-(n + mid) // hop)
    while (p_max - 1) * hop - mid >= n:
        p_max -= 1
    return -mid, p_max


def stft_db(y: np.ndarray, shift: np.ndarray | None, win: np.ndarray, hop: int, mfft: int, p_num: int) -> np.ndarray:
    """waterfall_plot.py:97-99 ``10. * log10(abs(self._SFT.stft(self._y)))`` after ``shiftFreq``; the
    transform restated per slice: window * segment (zeros outside the row), zero-padded to mfft,
    FFT, fftshift ('centered'), no phase shift, window already scaled ('magnitude')."""
    z = np.asarray(y, dtype=np.complex128).reshape(-1)
    if shift is not None:
        z = z * np.asarray(shift).reshape(-1)
    n, m = z.size, len(win)
    out = np.empty((mfft, p_num), dtype=np.complex128)
    for p in range(p_num):
        seg = np.zeros(mfft, dtype=np.complex128)
        k0 = p * hop - m // 2
        lo, hi = max(k0, 0), min(k0 + m, n)
        if hi > lo:
            seg[lo - k0:hi - k0] = z[lo:hi] * np.conj(win[lo - k0:hi - k0])
        out[:, p] = np.fft.fftshift(np.fft.fft(seg))
    with np.errstate(divide='ignore'):
        return 10. * np.log10(abs(out))


# ----------------------------------------------------------------------------- the chain
@dataclass
class Chain:
    """Reference-equivalent processing of a raw byte stream, chunk by chunk.

    Quirk switches, defaults = what the tests call "reference semantics":

    * ``ceil_alloc`` (Q1): size the decimated buffers ceil(N/q) so that non-dividing ``dec``
      (e.g. -d 50) produces output; the unmodified reference produces nothing there.
    * re/im (Q2): intended semantics (no output filter).
    * float32 (Q3): chain runs in complex128 (``decode(upcast_f32=True)``).
    * ``stale_tail`` (Q5): a trailing partial read is processed as a full chunk whose tail
      holds the previous chunk's bytes, as the reference's reused buffer does.
    * NCO phase restarts every chunk (Q8); SIMO processes K+1 rows (Q9).
    """
    fs: int
    enc: str = 'h'
    big_endian: bool | None = None
    swap: bool = False
    center: int = 0
    vfos: str | None = None
    simo: bool = False
    dec: int = 2
    demod: str = 'fm'
    omega_out: int = 12500
    correct_iq: bool = False
    normalize: bool = False
    impedance: int = 50
    chunk_bytes: int = CHUNK_BYTES
    stale_tail: bool = True
    nthreads: int = 1
    _off: np.ndarray = field(default_factory=lambda: np.array([0j], dtype=np.complex128))

    def __post_init__(self):
        self.dt = sample_dtype(self.enc, self.big_endian, self.swap)
        self.n = self.chunk_bytes // (2 * self.dt.itemsize)
        self.rows = nco_rows(self.center, self.vfos if self.simo else None)
        self.R = len(self.rows)
        self.M = (self.n + self.dec - 1) // self.dec
        self.filt = decimation_filter(self.dec)
        self.decimated_fs = self.fs // self.dec
        self.out_sos = (output_filter_sos(self.decimated_fs, self.omega_out)
                        if self.demod in ('fm', 'am') else None)
        self.shift = nco_table(self.rows, self.fs, self.n, self.simo)
        self._buf = np.zeros(self.chunk_bytes, dtype=np.uint8)

    # -- producer half (read_file.py:100-103)
    def ingest(self, raw: bytes | np.ndarray) -> np.ndarray:
        """One read of <= chunk_bytes bytes -> corrected complex chunk (always n samples)."""
        raw = np.frombuffer(raw, dtype=np.uint8)
        if raw.size == self.chunk_bytes or not self.stale_tail:
            self._buf[:raw.size] = raw
            if raw.size < self.chunk_bytes and not self.stale_tail:
                self._buf[raw.size:] = 0
        else:
            self._buf[:raw.size] = raw  # tail keeps the previous chunk's bytes (Q5)
        z = decode(self._buf, self.dt)
        if self.normalize:
            normalize(z, self.dt.char)
        if self.correct_iq:
            correct_iq(z, self._off, self.fs, self.impedance)
        return z

    # -- consumer half (dsp_processor.py:140-149)
    def process_chunk(self, z: np.ndarray) -> np.ndarray:
        """corrected complex chunk (n,) -> demodulated (R, M) float64."""
        y = self.decimated(z[None, :])[0]
        return self.demodulate(y)

    def decimated(self, zs: np.ndarray) -> np.ndarray:
        """(nchunks, n) corrected chunks -> (nchunks, R, M) complex decimator output."""
        zs = np.ascontiguousarray(zs, dtype=np.complex128)
        nchunks = zs.shape[0]
        sos, zi, edge = self.filt
        y = np.empty((nchunks, self.R, self.M), dtype=np.complex128)
        sh = self.shift
        rc = _lib().orc_batch(_dp(sos), sos.shape[0], _dp(zi), edge, _dp(zs.view(np.float64)),
                              _dp(sh.view(np.float64)) if sh is not None else None,
                              self.R, self.n, self.dec, nchunks, _dp(y.view(np.float64)),
                              self.nthreads)
        if rc < 0:
            raise MemoryError('oracle batch allocation failed')
        return y

    def demodulate(self, y: np.ndarray) -> np.ndarray:
        if self.demod == 'fm':
            z = fm_demod(y)
        elif self.demod == 'am':
            z = am_demod(y)
        elif self.demod == 're':
            z = real_output(y)
        elif self.demod == 'im':
            z = imag_output(y)
        else:
            raise ValueError(f'Invalid demod type {self.demod}')
        if self.out_sos is not None:
            z = sosfilt_real(self.out_sos, z)
        return z

    def run(self, stream: bytes | np.ndarray) -> np.ndarray:
        """Whole byte stream -> (R, nchunks*M) float64, reading chunk_bytes at a time like
        read_file.py:119-125."""
        stream = np.frombuffer(stream, dtype=np.uint8)
        chunks = []
        for o in range(0, stream.size, self.chunk_bytes):
            chunks.append(self.ingest(stream[o:o + self.chunk_bytes]).copy())
        if not chunks:
            return np.empty((self.R, 0), dtype=np.float64)
        y = self.decimated(np.stack(chunks))
        out = np.empty((self.R, len(chunks) * self.M), dtype=np.float64)
        for c in range(len(chunks)):
            out[:, c * self.M:(c + 1) * self.M] = self.demodulate(y[c])
        return out

    def run_fast(self, stream: bytes | np.ndarray) -> np.ndarray:
        """Same results as ``run`` for whole-chunk streams, organised for speed (bench.py's CPU
        baseline): decode, decimate, pair phase and the output filter are chunk/row-parallel in
        C with ``nthreads`` threads; the IQ recurrence is one serial pass; the FM resample is one
        batched numpy FFT."""
        stream = np.frombuffer(stream, dtype=np.uint8)
        nch, tail = divmod(stream.size, self.chunk_bytes)
        if tail or self.normalize or nch == 0:
            return self.run(stream)
        L = _lib()
        z = np.empty((nch, self.n), dtype=np.complex128)
        L.orc_decode_batch(stream.ctypes.data, nch * self.n, self.enc.encode(),
                           int(needs_swap(self.dt)), _dp(z.view(np.float64)), self.nthreads)
        if self.correct_iq:
            L.orc_correct_iq(_dp(z.view(np.float64)), z.size, self.impedance / self.fs,
                             _dp(self._off.view(np.float64)))
        y = self.decimated(z)                      # (nch, R, M)
        rows = nch * self.R
        if self.demod == 'fm':
            h = self.M >> 1
            res = np.empty((rows, h), dtype=np.float64)
            L.orc_fm_pairs_rows(_dp(y.view(np.float64)), rows, self.M, _dp(res), self.nthreads)
            zz = np.ascontiguousarray(resample_2x(res, self.M))
        elif self.demod == 'am':
            zz = np.ascontiguousarray(am_demod(y).reshape(rows, self.M))
        elif self.demod == 're':
            zz = np.ascontiguousarray(y.real.reshape(rows, self.M))
        else:
            zz = np.ascontiguousarray(y.imag.reshape(rows, self.M))
        if self.out_sos is not None:
            L.orc_sosfilt_rows(_dp(self.out_sos), self.out_sos.shape[0], _dp(zz), rows, self.M,
                               self.nthreads)
        return np.ascontiguousarray(zz.reshape(nch, self.R, self.M).transpose(1, 0, 2)).reshape(
            self.R, nch * self.M)

    def run_framed(self, stream) -> list[bytes]:
        """Byte streams as written: one native-endian file stream (standard) or one big-endian
        stream per row (SIMO)."""
        out = self.run(stream)
        if not self.simo:
            # per chunk the reference packs z.flat == rows concatenated; R == 1 here
            return [np.ascontiguousarray(out[0], dtype='=f8').tobytes()]
        return [np.ascontiguousarray(out[r], dtype='>f8').tobytes() for r in range(self.R)]
