"""Input description: encoding letter -> numpy dtype, RIFF/RIFX header -> dtype, rate and data
offset (reference: src/misc/file_util.py:46-195).  Returns the same ``fileInfo`` dictionary keys
the reference's processors and reader consume (``bitsPerSample``, ``sampRate``, ``dataOffset``,
``isSocket``), and exposes the same public types (``DataType``, ``WaveFormat``, ``ExWaveFormat``)."""
from __future__ import annotations

import struct
from enum import Enum

import numpy as np

from .mappable_enum import MappableEnum


class WaveFormat(Enum):
    """``wFormatTag`` values the header parser knows (file_util.py:29-34)."""
    WAVE_FORMAT_PCM = 0x0001
    WAVE_FORMAT_IEEE_FLOAT = 0x0003
    WAVE_FORMAT_ALAW = 0x0006
    WAVE_FORMAT_MULAW = 0x0007
    WAVE_FORMAT_EXTENSIBLE = 0xFFFE


class ExWaveFormat(Enum):
    """SubFormat codes of WAVE_FORMAT_EXTENSIBLE (file_util.py:37-43): signedness and byte order."""
    PCM_S_LE = 0x01
    PCM_S_BE = 0x02
    PCM_U_LE = 0x05
    PCM_U_BE = 0x06


class DataType(MappableEnum):
    """``-e`` letters -> numpy dtype (file_util.py:46-60): single bytes have no order, the rest are
    native-endian for files; ``str(member)`` is the letter."""
    b = np.dtype('|i1')
    B = np.dtype('|u1')
    h = np.dtype('=i2')
    H = np.dtype('=u2')
    i = np.dtype('=i4')
    I = np.dtype('=u4')  # noqa: E741
    f = np.dtype('=f4')
    d = np.dtype('=f8')

    def __str__(self):
        return self.name

    @classmethod
    def fromWav(cls, bits: int, aFormat: Enum, bFormat: Enum | None, isRifx: bool) -> np.dtype:
        """Header fields -> dtype with explicit byte order (file_util.py:62-97).  PCM without a
        SubFormat: 8 bits unsigned, wider signed; the SubFormat of an extensible header decides
        signedness and may force big-endian; RIFX forces big-endian."""
        member, big = None, bool(isRifx)
        if aFormat == WaveFormat.WAVE_FORMAT_IEEE_FLOAT:
            member = {32: cls.f, 64: cls.d}.get(bits)
        elif aFormat in (WaveFormat.WAVE_FORMAT_PCM, WaveFormat.WAVE_FORMAT_EXTENSIBLE):
            kind = None
            if isinstance(bFormat, ExWaveFormat):
                _, kind, order = bFormat.name.split('_')
                big = big or order == 'BE'
            table = {8: {'S': cls.b, 'U': cls.B, None: cls.B},
                     16: {'S': cls.h, 'U': cls.H, None: cls.h},
                     32: {'S': cls.i, 'U': cls.I, None: cls.i}}.get(bits)
            member = table[kind] if table is not None else None
        if member is None:
            raise ValueError(f'Unsupported format: {aFormat} @ {bits} bits')
        return member.value.newbyteorder('>' if big else '<')


ENCODINGS = {m.name: m.value.str for m in DataType}
_GUID_TAIL = b'\x00\x00\x00\x00\x10\x00\x80\x00\x00\xAA\x00\x38\x9B\x71'


def dtypeOf(enc: str) -> np.dtype:
    """``DataType[enc].value``; an unknown letter is a ``KeyError`` as in the reference."""
    return DataType[enc].value


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
        if fmt == WaveFormat.WAVE_FORMAT_EXTENSIBLE.value:
            extra, = struct.unpack(e + 'H', fh.read(2))
            fh.read(extra - 16)
            code, = struct.unpack(e + 'H', fh.read(2))
            if fh.read(14) != _GUID_TAIL:
                raise ValueError('Invalid: SubFormat GUID malformed')
            sub = ExWaveFormat(code)                       # unknown codes: ValueError
        info = {'subchunk1Size': size, 'audioFormat': fmt, 'numChannels': nch, 'sampRate': rate,
                'byteRate': brate, 'blockAlign': align,
                'bitsPerSample': DataType.fromWav(bits, WaveFormat(fmt), sub, e == '>'),
                'isSocket': False, 'bitRate': (bits * brate * align) >> 3}
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
