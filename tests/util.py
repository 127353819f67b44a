"""Helpers shared by the parity tests."""
import hashlib
import os
import struct

import numpy as np

from cases import CASES

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def rel_err(a, b):
    """SURVEY 8d error metric: max|a-b| / max|b| over the whole output."""
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def parse_wav(raw: bytes):
    """Minimal RIFF/RIFX parse mirroring file_util.py:125-195 for the canonical 44-byte header,
    including its dataOffset quirk (4 bytes short).  Returns
    (fs, enc, big_endian, data_offset)."""
    e = '>' if raw[:4] == b'RIFX' else '<'
    fmt, ch, fs, _, _, bits = struct.unpack(e + 'HHIIHH', raw[20:36])
    off = raw.find(b'data', 36) + 4
    enc = {(1, 8): 'B', (1, 16): 'h', (1, 32): 'i', (3, 32): 'f', (3, 64): 'd'}[(fmt, bits)]
    return fs, enc, e == '>', off


def case_stream(name):
    """(raw data bytes after dataOffset, oracle Chain kwargs) for a parity case."""
    c = CASES[name]
    raw = c['make']()
    kw = dict(center=c['center'], vfos=c['vfos'], simo=c['simo'], dec=c['dec'], demod=c['demod'],
              omega_out=c['omega'], correct_iq=c['correct_iq'], normalize=c['normalize'],
              swap=c['swap'])
    if c['suffix'] == '.wav':
        fs, enc, be, off = parse_wav(raw)
        kw.update(fs=fs, enc=enc, big_endian=be)
        body = raw[off:]
    else:
        kw.update(fs=c['fs'], enc=c['enc'], big_endian=None)
        body = raw
    return raw, body, kw


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    return g


def sha(raw):
    return hashlib.sha256(raw).hexdigest()
