#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (peads/sdrterm, imported from
/root/reference/src) in-process on the seeded synthetic inputs of tests/signals.py.

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Producer half: the reference's own ``misc.read_file.readFile`` is driven with a capture queue,
so chunking, dataOffset, stale-tail reads (8-Q5), -X, normalisation and the numba
IQ corrector are the reference's code.  Consumer half: ``DspProcessor`` / ``VfoProcessor``
objects built as ``IOArgs`` builds them, ``_generateShift`` / ``_processChunk`` called as
``_processData`` calls them (dsp_processor.py:164-183).  Deviations, each recorded in the
fixture's ``notes``:
  * SIMO shift table is built with the reference's expression from the object's own omega
    instead of ``_generateShift`` (which blocks on TCP clients, vfo_processor.py:71-78);
  * ``ceil`` cases size y/z with ceil(N/q) (8-Q1), ``re``/``im`` skip the (empty) output
    filter (8-Q2), float32 chunks are widened to complex128 before the chain (8-Q3).
Outputs: tests/golden/<case>.npz with the demodulated float64 rows, the decimator output of the
first chunk, the IQ-corrected first chunk head, and a sha256 of the input bytes.
"""
import hashlib
import os
import sys
import tempfile

os.environ.setdefault('NUMBA_CACHE_DIR', os.path.join(tempfile.gettempdir(), 'numba_ref_cache'))
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, '/root/reference/src')

import numpy as np  # noqa: E402

import signals  # noqa: E402  (tests/signals.py)
from dsp.dsp_processor import DspProcessor  # noqa: E402  (reference)
from dsp.vfo_processor import VfoProcessor  # noqa: E402  (reference)
from misc.file_util import checkWavHeader  # noqa: E402  (reference)
from misc.read_file import readFile  # noqa: E402  (reference)


class _Val:
    value = 0


class _Proc:
    exitcode = None
    name = 'capture'


class _Capture:
    def __init__(self):
        self.chunks = []

    def put_nowait(self, z):
        if isinstance(z, bytes):
            return
        self.chunks.append(np.array(z, copy=True))

    def close(self):
        pass


def produce(path, fs, enc, swap, correct_iq, normalize):
    info = checkWavHeader(path, fs, enc)
    cap = _Capture()
    readFile(buffers=[cap], processes=[_Proc()], isDead=_Val(), inFile=path,
             swapEndianness=swap, correctIq=correct_iq, normalize=normalize,
             **{k: v for k, v in info.items() if k in ('bitsPerSample', 'dataOffset', 'isSocket')},
             fs=info['sampRate'])
    return info, cap.chunks


def consume(chunks, fs, center, omega, dec, demod, simo, vfos, ceil_alloc):
    kw = dict(center=center, omegaOut=omega, tuned=0, dec=dec, smooth=0,
              fileInfo={'bitsPerSample': 'x'})
    if simo:
        p = VfoProcessor(fs, vfoHost='localhost:0', vfos=vfos, **kw)
    else:
        p = DspProcessor(fs, **kw)
    {'fm': p.selectOutputFm, 'am': p.selectOutputAm, 're': p.selectOutputReal,
     'im': p.selectOutputImag}[demod]()
    n = chunks[0].size
    if simo:
        om = p._VfoProcessor__omega
        p._shift = np.ones((p._nFreq, n), dtype=np.complex128)
        for i, w in enumerate(om):
            p._shift[i][:] = np.exp(w * np.arange(n))  # vfo_processor.py:72-74
    else:
        p._generateShift(n)
    M = -(-n // dec) if ceil_alloc else n // dec
    x = np.empty((p._nFreq, n), dtype=np.complex128)
    y = np.empty((p._nFreq, M), dtype=np.complex128)
    z = np.empty((p._nFreq, M), dtype=np.float64)
    outs, y0 = [], None
    for c in chunks:
        x[0, :] = c.astype(np.complex128)  # 8-Q3 widening is a no-op for complex128 chunks
        if demod in ('re', 'im'):
            # 8-Q2: _processChunk would raise in sosfilt([]); run its first three steps
            from dsp.demodulation import shiftFreq
            from scipy.signal import decimate
            if p._shift is not None:
                shiftFreq(x[0], p._shift, x)
            y[:] = decimate(x, dec)
            p.demod(y, z)
        else:
            p._processChunk(x, y, z)
        if y0 is None:
            y0 = y.copy()
        outs.append(z.copy())
    return np.concatenate(outs, axis=1), y0


from cases import CASES  # noqa: E402  (tests/cases.py)


def main(only=None):
    for name, c in CASES.items():
        if only and name not in only:
            continue
        raw = c['make']()
        with tempfile.NamedTemporaryFile(suffix=c['suffix'], delete=False) as f:
            f.write(raw)
            path = f.name
        try:
            info, chunks = produce(path, c['fs'], c['enc'], c['swap'], c['correct_iq'],
                                   c['normalize'])
        finally:
            os.unlink(path)
        fs = info['sampRate']
        widened = chunks[0].dtype != np.complex128
        out, y0 = consume(chunks, fs, c['center'], c['omega'], c['dec'], c['demod'], c['simo'],
                          c['vfos'], c['ceil'])
        notes = []
        if c['simo']:
            notes.append('shift table built from VfoProcessor.__omega (no TCP clients)')
        if c['ceil']:
            notes.append('8-Q1: y/z sized ceil(N/q)')
        if c['demod'] in ('re', 'im'):
            notes.append('8-Q2: output filter skipped')
        if widened:
            notes.append('8-Q3: complex64 chunks widened to complex128 before the chain')
        np.savez_compressed(
            os.path.join(HERE, name + '.npz'),
            out=out, y0=y0, z0_head=chunks[0][:64].astype(np.complex128),
            nchunks=len(chunks), fs=fs, data_offset=info['dataOffset'],
            dtype=str(info['bitsPerSample'].str), sha256=hashlib.sha256(raw).hexdigest(),
            notes='; '.join(notes),
            params=repr({k: v for k, v in c.items() if k != 'make'}))
        print(f'{name}: chunks={len(chunks)} out={out.shape} max|out|={np.abs(out).max():.4g} '
              f'dtype={info["bitsPerSample"].str} off={info["dataOffset"]} [{"; ".join(notes)}]')


if __name__ == '__main__':
    main(sys.argv[1:])
