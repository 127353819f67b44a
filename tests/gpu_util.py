"""Shared helpers of the GPU parity tests: build an Engine for a parity case."""
import numpy as np

from oracle import oracle as orc
from sdrterm_b200.engine import Engine
from sdrterm_b200.plan import build_plan
from util import case_stream

CB = 131072


def plan_for(kw):
    ch = orc.Chain(**kw)
    return build_plan(kw['fs'], kw['enc'], kw['dec'], ch.rows, simo=kw['simo'],
                      swap=orc.needs_swap(ch.dt), correct_iq=kw['correct_iq'],
                      normalize=kw['normalize'], demod=kw['demod'], omega_out=kw['omega_out'])


def chunked(body: bytes, stale_tail: bool = True) -> np.ndarray:
    """Whole-chunk byte matrix as the reference's reused read buffer presents it (8-Q5)."""
    raw = np.frombuffer(body, dtype=np.uint8)
    n = -(-raw.size // CB)
    out = np.zeros((n, CB), dtype=np.uint8)
    buf = np.zeros(CB, dtype=np.uint8)
    for c in range(n):
        part = raw[c * CB:(c + 1) * CB]
        buf[:part.size] = part
        out[c] = buf
    return out


def run_case(name, max_chunks=8, use_tc=True):
    """use_tc=True takes the tensor-core block front end (k_tc) where the shape supports it;
    False forces the FP64 block kernel (k_main)."""
    raw, body, kw = case_stream(name)
    pl = plan_for(kw)
    chunks = chunked(body)
    with Engine(pl, max_chunks=max_chunks, use_tc=use_tc, keep_decimated=True) as eng:
        out = eng.process(chunks)
        y = eng.decimated(min(chunks.shape[0], max_chunks))
        off = eng.iq_state
    return kw, pl, chunks, out, y, off
