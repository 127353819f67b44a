"""Deterministic synthetic IQ byte streams for the five BASELINE.json configurations
(SURVEY.md section 8d table).  Raw bytes are the artefact; everything derives from
numpy.random.default_rng(seed)."""
from __future__ import annotations

import struct

import numpy as np

CHUNK_BYTES = 131072


def _fm_carrier(n, fs, fc, fmod, dev, amp, phase0=0.0, t0=0):
    t = (np.arange(n) + t0) / fs
    ph = 2 * np.pi * fc * t + (dev / fmod) * np.sin(2 * np.pi * fmod * t) + phase0
    return amp * np.exp(1j * ph)


def _interleave(z, dtype, lo, hi, rnd=True):
    iq = np.empty(2 * z.size, dtype=np.float64)
    iq[0::2] = z.real
    iq[1::2] = z.imag
    if rnd:
        iq = np.clip(np.rint(iq), lo, hi)
    return iq.astype(dtype)


def wav_header(fs: int, bits: int, nbytes: int, fmt: int = 1, rifx: bool = False) -> bytes:
    """Canonical 44-byte RIFF/RIFX header, 2 channels (I, Q)."""
    e = '>' if rifx else '<'
    block = 2 * bits // 8
    return (b'RIFX' if rifx else b'RIFF') + struct.pack(e + 'I', 36 + nbytes) + b'WAVEfmt ' + \
        struct.pack(e + 'IHHIIHH', 16, fmt, 2, fs, fs * block, block, bits) + b'data' + \
        struct.pack(e + 'I', nbytes)


def c1_bytes(nsamples: int, seed: int = 0, header: bool = True) -> bytes:
    """C1: int16 LE, fs 1.024 MS/s; FM carrier at +15 kHz, 1 kHz modulation, +-2.5 kHz
    deviation, amplitude 8000, AWGN sigma 200, DC (+37, -21)."""
    fs = 1_024_000
    rng = np.random.default_rng(seed)
    z = _fm_carrier(nsamples, fs, 15_000, 1_000, 2_500, 8000.0)
    z = z + rng.normal(0, 200, nsamples) + 1j * rng.normal(0, 200, nsamples) + (37 - 21j)
    body = _interleave(z, '<i2', -32768, 32767).tobytes()
    return (wav_header(fs, 16, len(body)) if header else b'') + body


def c2_bytes(nsamples: int, seed: int = 2) -> bytes:
    """C2: uint8, fs 2.4 MS/s; AM 40*(1+0.5 sin 2pi 800 t) + 127.5 offset, sigma 3, clipped."""
    fs = 2_400_000
    rng = np.random.default_rng(seed)
    t = np.arange(nsamples) / fs
    env = 40.0 * (1 + 0.5 * np.sin(2 * np.pi * 800 * t))
    z = env * np.exp(1j * 0.3) + (127.5 + 127.5j)
    z = z + rng.normal(0, 3, nsamples) + 1j * rng.normal(0, 3, nsamples)
    return _interleave(z, 'u1', 0, 255).tobytes()


def vfo_grid(k: int, step: int) -> list[int]:
    """k VFO offsets on a `step` Hz grid, symmetric about 0 and excluding 0."""
    half = k // 2
    offs = [i * step for i in range(-half, 0)] + [i * step for i in range(1, k - half + 1)]
    return offs


def c3_bytes(nsamples: int, seed: int = 3, k: int = 16, step: int = 100_000):
    """C3: int16 BIG-endian, fs 2.4 MS/s; one FM carrier per VFO offset (+ centre), amp 1500
    each, sigma 100.  Returns (bytes, vfos-csv)."""
    fs = 2_400_000
    rng = np.random.default_rng(seed)
    offs = vfo_grid(k, step)
    z = np.zeros(nsamples, dtype=np.complex128)
    for i, f in enumerate(offs + [0]):
        z += _fm_carrier(nsamples, fs, f, 700 + 40 * i, 2_000, 1500.0, phase0=0.37 * i)
    z = z + rng.normal(0, 100, nsamples) + 1j * rng.normal(0, 100, nsamples)
    return _interleave(z, '>i2', -32768, 32767).tobytes(), ','.join(str(o) for o in offs)


def c4_bytes(nsamples: int, seed: int = 4, k: int = 256, step: int = 200_000):
    """C4: float32 LE, fs 61.44 MS/s; k+1 FM carriers, unit total amplitude, sigma 0.01."""
    fs = 61_440_000
    rng = np.random.default_rng(seed)
    offs = vfo_grid(k, step)
    amp = 1.0 / np.sqrt(k + 1)
    z = np.zeros(nsamples, dtype=np.complex128)
    for i, f in enumerate(offs + [0]):
        z += _fm_carrier(nsamples, fs, f, 900 + 13 * i, 40_000, amp, phase0=0.11 * i)
    z = z + rng.normal(0, 0.01, nsamples) + 1j * rng.normal(0, 0.01, nsamples)
    return _interleave(z, '<f4', 0, 0, rnd=False).tobytes(), ','.join(str(o) for o in offs)


def generic_bytes(enc: str, nsamples: int, seed: int, fs: int, fc: float,
                  big_endian: bool = False) -> bytes:
    """A single FM carrier at fc scaled to ~1/4 of the encoding's range, for the encodings the
    named configs do not cover (b, H, i, I, d)."""
    rng = np.random.default_rng(seed)
    spec = {'b': ('i1', -128, 127, 30.0, 0.0), 'B': ('u1', 0, 255, 30.0, 127.5),
            'h': ('i2', -32768, 32767, 8000.0, 0.0), 'H': ('u2', 0, 65535, 8000.0, 32768.0),
            'i': ('i4', -2 ** 31, 2 ** 31 - 1, 5e8, 0.0), 'I': ('u4', 0, 2 ** 32 - 1, 5e8, 2.0 ** 31),
            'f': ('f4', 0, 0, 0.25, 0.0), 'd': ('f8', 0, 0, 0.25, 0.0)}[enc]
    kind, lo, hi, amp, dc = spec
    z = _fm_carrier(nsamples, fs, fc, 1_000, 2_500, amp) + dc * (1 + 1j)
    z = z + amp * 0.02 * (rng.normal(0, 1, nsamples) + 1j * rng.normal(0, 1, nsamples))
    dt = np.dtype(kind).newbyteorder('>' if big_endian else '<')
    return _interleave(z, dt, lo, hi, rnd=kind[0] in 'iu').tobytes()
