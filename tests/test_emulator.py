"""CPU: the numpy twins of the CUDA kernels (tests/emulator.py), driven by the SAME plan tables the
kernels consume, against the oracle.  This validates the host-side table construction -- the modal
block form, the decoupled IQ-corrector terms and the int8 digit matrices of the
tensor-core front end (exact integer GEMM emulated in int64) -- without a GPU."""
import numpy as np
import pytest

import emulator as emu
from oracle import oracle as orc
from sdrterm_b200.plan import build_plan, build_tc
from util import case_stream, rel_err

TOL = 1e-9


def _plan(kw):
    ch = orc.Chain(**kw)
    return build_plan(kw['fs'], kw['enc'], kw['dec'], ch.rows, simo=kw['simo'], swap=orc.needs_swap(ch.dt),
                      correct_iq=kw['correct_iq'], normalize=kw['normalize'], demod=kw['demod'],
                      omega_out=kw['omega_out'])


@pytest.mark.parametrize('name', ['c1_fm_wav_int16', 'c2_am_u8_d50_ceil', 'enc_I_norm_am', 'enc_h_d2_fm'])
def test_block_kernel_twin_matches_oracle(name):
    raw, body, kw = case_stream(name)
    pl = _plan(kw)
    out, ys, off = emu.emu_stream(pl, body)
    ch = orc.Chain(**kw)
    ref = ch.run(body)
    assert out.shape == ref.shape and rel_err(out, ref) < 1e-11
    assert abs(off - ch._off[0]) <= 1e-9 * max(1.0, abs(ch._off[0]))


@pytest.mark.parametrize('name', ['c1_fm_wav_int16', 'c2_am_u8_d64', 'enc_H_swap_im', 'c3_simo16_int16be'])
def test_tensor_core_twin_matches_oracle(name):
    """Super-block GEMM rows, 5 digit columns, re-rounded low-byte coefficients, separate scale of
    the local outputs: the whole chain stays ~1e-12 from the oracle (tolerance 1e-9)."""
    raw, body, kw = case_stream(name)
    pl = _plan(kw)
    tc = build_tc(pl)
    assert tc is not None and tc.NCOL == 5 and tc.Npad == 208 and tc.K in (256, 512)
    assert tc.col_l1 * (128 if tc.a_signed else 255) * 257 < 2 ** 31   # digit-pair sums fit int32 on the device
    assert tc.a_signed == (kw['enc'] in 'bhH') and bool(tc.xor_mask.any()) == (kw['enc'] in 'hH')
    body = body[:len(body) // 131072 * 131072] or body
    out, ys, off = emu.emu_stream_tc(pl, tc, body)
    ch = orc.Chain(**kw)
    ref = ch.run(body)
    assert out.shape == ref.shape and rel_err(out, ref) < 2e-11


def test_iq_decoupling_constants():
    """w_i = beta_i W~_i - alpha_i s: beta -> 1, gamma -> 0 without --correct-iq; with it, gamma is
    the two-sided response of the decimation filter to the offset mode lam e^{jw}."""
    pl0 = build_plan(1_024_000, 'h', 64, [15000], correct_iq=False, demod='fm', omega_out=5000)
    assert np.all(pl0.beta == 1) and np.all(pl0.alpha == 0) and np.all(pl0.gamma == 0)
    pl = build_plan(1_024_000, 'h', 64, [0], correct_iq=True, demod='fm', omega_out=5000)
    # centre 0: the offset mode is (almost) DC, where the low-pass has unit gain (0.05 dB ripple)
    assert abs(abs(pl.gamma[0]) - 1) < 0.02
    pl = build_plan(1_024_000, 'h', 64, [200_000], correct_iq=True, demod='fm', omega_out=5000)
    assert abs(pl.gamma[0]) < 1e-6                      # far in the stop band


@pytest.mark.parametrize('enc,q,demod,iq,norm,center', [
    ('h', 5, 'am', True, False, 30000), ('B', 25, 'am', True, True, 0), ('f', 63, 'fm', False, False, -20000),
    ('b', 3, 're', True, False, 10000)])
def test_odd_decimation_factors(enc, q, demod, iq, norm, center):
    """An odd block has a middle sample without a mirror and, for the lane past the last pair, an
    empty descending run (plan.run_len must not go negative): the reference takes any `-d`."""
    import signals
    isz = {'b': 1, 'B': 1, 'h': 2, 'f': 4}[enc]
    n = 2 * (131072 // (2 * isz))
    body = signals.generic_bytes(enc, n, 17 + q, 1_000_000, center or 40_000, big_endian=False)
    kw = dict(fs=1_000_000, enc=enc, center=center, dec=q, demod=demod, omega_out=4000, correct_iq=iq,
              vfos=None, simo=False, normalize=norm, swap=False, big_endian=None)
    pl = _plan(kw)
    assert pl.run_len.sum() == q and (pl.run_len >= 0).all()
    out, ys, off = emu.emu_stream(pl, body)
    ch = orc.Chain(**kw)
    ref = ch.run(body)
    assert out.shape == ref.shape and rel_err(out, ref) < 1e-10
    if iq:
        assert abs(off - ch._off[0]) <= 1e-9 * max(1.0, abs(ch._off[0]))
