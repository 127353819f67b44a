"""Input description: encoding letter -> numpy dtype, RIFF/RIFX header -> dtype, rate and data
offset (reference: src/misc/file_util.py:46-195).  Returns the same ``fileInfo`` dictionary keys
the reference's processors and reader consume (``bitsPerSample``, ``sampRate``, ``dataOffset``,
``isSocket``)."""
from __future__ import annotations

import struct

import numpy as np

# file_util.py:46-60: single bytes have no order, the rest are native for files
ENCODINGS = {'b': '|i1', 'B': '|u1', 'h': '=i2', 'H': '=u2', 'i': '=i4', 'I': '=u4', 'f': '=f4', 'd': '=f8'}
_WAVE_PCM, _WAVE_FLOAT, _WAVE_EXT = 0x0001, 0x0003, 0xFFFE
_EX_SIGNED = {0x01: True, 0x02: True, 0x05: False, 0x06: False}        # PCM_S_LE/BE, PCM_U_LE/BE
_EX_BIG = {0x01: False, 0x02: True, 0x05: False, 0x06: True}
_GUID_TAIL = b'\x00\x00\x00\x00\x10\x00\x80\x00\x00\xAA\x00\x38\x9B\x71'


def dtypeOf(enc: str) -> np.dtype:
    try:
        return np.dtype(ENCODINGS[enc])
    except KeyError:
        raise ValueError(f'unknown encoding {enc!r}; one of {"|".join(ENCODINGS)}') from None


def parseRawType(file: str | None, fs: int | None, enc: str | None, isSocket: bool = False) -> dict:
    """Headerless input (:99-111): rate and encoding are mandatory; anything that looks like
    ``host:port`` is taken as network byte order."""
    if fs is None or fs < 1 or enc is None:
        raise ValueError('Valid sampling rate, encoding type and bit-size are all required for raw pcm input')
    dt = dtypeOf(enc)
    if file is not None and ':' in file:
        dt = dt.newbyteorder('>')
    return {'subchunk1Size': 0, 'audioFormat': 0, 'numChannels': 0, 'sampRate': int(fs),
            'byteRate': int(fs), 'blockAlign': 0, 'bitsPerSample': dt, 'dataOffset': 0,
            'isSocket': isSocket}


def _wavDtype(bits: int, fmt: int, sub: int | None, rifx: bool) -> np.dtype:
    big = rifx
    if fmt == _WAVE_FLOAT:
        ch = {32: 'f4', 64: 'f8'}.get(bits)
    elif fmt in (_WAVE_PCM, _WAVE_EXT):
        signed = None
        if sub is not None and sub in _EX_SIGNED:
            signed, big = _EX_SIGNED[sub], rifx or _EX_BIG[sub]
        if signed is None:
            signed = bits != 8                                         # 8-bit PCM is unsigned (:66-68)
        ch = {8: 'i1', 16: 'i2', 32: 'i4'}.get(bits)
        if ch is not None and not signed:
            ch = 'u' + ch[1:]
    else:
        ch = None
    if ch is None:
        raise ValueError(f'Unsupported format: {fmt:#x} @ {bits} bits')
    return np.dtype(('>' if big else '<') + ch)


def checkWavHeader(f, fs: int | None, enc: str | None) -> dict:
    """(:125-195).  Replicates the reference's ``dataOffset``: the position right after the
    ``data`` tag, i.e. the 4-byte length field is read as samples."""
    if f is None:
        return parseRawType(f, fs, enc)
    if isinstance(f, str) and ':' in f:
        import socket
        host, port = f.split(':')
        try:
            if socket.getaddrinfo(host, port):
                return parseRawType(f, fs, enc, True)
        except socket.gaierror:
            pass
    with open(f, 'rb') as fh:
        head = fh.read(4)
        if head[:3] != b'RIF':
            if '.wav' in f:
                raise ValueError('Invalid: Expected raw pcm file, but got malformed RIFF header')
            return parseRawType(f, fs, enc)
        if head[3:4] not in (b'F', b'X'):
            raise ValueError('Invalid: Malformed RIFF/X header')
        e = '>' if head[3:4] == b'X' else '<'
        fh.read(4)
        if fh.read(4) != b'WAVE':
            raise ValueError('Invalid: Expected a wave file')
        if fh.read(4) != b'fmt ':
            raise ValueError('Invalid: Format section not found')
        size, fmt, nch, rate, brate, align, bits = struct.unpack(e + 'IHHIIHH', fh.read(20))
        sub = None
        if fmt == _WAVE_EXT:
            extra, = struct.unpack(e + 'H', fh.read(2))
            fh.read(extra - 16)
            sub, = struct.unpack(e + 'H', fh.read(2))
            if fh.read(14) != _GUID_TAIL:
                raise ValueError('Invalid: SubFormat GUID malformed')
        info = {'subchunk1Size': size, 'audioFormat': fmt, 'numChannels': nch, 'sampRate': rate,
                'byteRate': brate, 'blockAlign': align,
                'bitsPerSample': _wavDtype(bits, fmt, sub, e == '>'), 'isSocket': False,
                'bitRate': (bits * brate * align) >> 3}
        pos = fh.tell()
        rest = fh.read(4096)
        k = rest.find(b'data')
        if k < 0:
            raise ValueError('Invalid: data section not found')
        info['dataOffset'] = pos + k + 4
    return info


def parseIntString(value) -> int:
    """``10k`` -> 10000, ``2.4M`` -> 2400000 (sdrterm.py:39-51)."""
    if value is None:
        raise ValueError('Value cannot be None')
    if isinstance(value, int):
        return value
    for suffix, mul in (('k', 1e3), ('M', 1e6)):
        if suffix in value:
            return int(float(value.replace(suffix, '')) * mul)
    return int(float(value))
