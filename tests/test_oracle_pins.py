"""Pins the CPU oracle (oracle/) -- against SciPy bit-exactly, against the reference's own
known-answer vectors, and against golden outputs of the reference itself
(tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import math

import numpy as np
import pytest
from scipy import signal

from cases import CASES
from oracle import oracle as orc
from util import case_stream, load_golden, rel_err, sha


@pytest.mark.parametrize('q,n', [(64, 32768), (50, 65536), (2, 8192), (7, 1000), (128, 16384)])
def test_decimate_bit_exact_vs_scipy(q, n):
    rng = np.random.default_rng(q)
    x = (rng.normal(size=(2, n)) + 1j * rng.normal(size=(2, n))) * 1000
    ref = signal.decimate(x, q)
    got = orc.decimate(x, q)
    assert got.shape == ref.shape
    if q == 2:
        # SciPy's build differs from the plain recurrence in 14 of 4096 outputs by 1 ulp here
        # (seen only for q=2); everything else is bit-identical.
        assert np.max(np.abs(got - ref)) <= 1e-15 * np.max(np.abs(ref))
    else:
        assert np.array_equal(got, np.ascontiguousarray(ref))


def test_decimate_py_matches_c():
    rng = np.random.default_rng(1)
    x = rng.normal(size=900) + 1j * rng.normal(size=900)
    a = orc.decimate_py(x, 8)
    b = orc.decimate(x, 8)
    assert np.array_equal(np.ascontiguousarray(a), b)


def test_sosfilt_real_bit_exact_vs_scipy():
    rng = np.random.default_rng(2)
    sos = orc.output_filter_sos(16000, 5000)
    x = rng.normal(size=(3, 777))
    assert np.array_equal(orc.sosfilt_real(sos, x), signal.sosfilt(sos, x))


def test_resample_matches_scipy():
    rng = np.random.default_rng(3)
    for m in (4, 256, 655, 512):
        r = rng.normal(size=m)
        num = 2 * m if m != 655 else 1311
        assert np.max(np.abs(orc.resample_2x(r, num) - signal.resample(r, num))) < 1e-14


def test_reference_known_answer_demod_vectors():
    """test/dsp/demodulation_test.py:14-50 of the reference: 8-sample vector, FM/AM within
    1e-12, re/im exact."""
    inp = np.array([[0j, 1 + 2j, 2 + 3j, 3 + 4j, 4 + 5j, 5 + 6j, 6 + 7j, 7 + 8j]])
    exp_fm = [0.0] * 4
    for i in range(0, 8, 2):
        t = inp[0][i] * inp[0][i + 1].conjugate()
        exp_fm[i >> 1] = math.atan2(t.imag, t.real)
    exp_fm = signal.resample(exp_fm, 8)
    assert np.max(np.abs(orc.fm_demod(inp)[0] - exp_fm)) < 1e-12
    exp_am = [z.real ** 2 + z.imag ** 2 for z in inp[0]]
    assert np.max(np.abs(orc.am_demod(inp)[0] - exp_am)) < 1e-12
    assert np.array_equal(orc.real_output(inp)[0], inp[0].real)
    assert np.array_equal(orc.imag_output(inp)[0], inp[0].imag)


def test_reference_known_answer_nco():
    """test/dsp/dsp_processor_test.py:50-58: None for centre 0, exact e^{-2j pi (fc/fs) k}."""
    assert orc.nco_table([0], 48000, 8, False) is None
    t = orc.nco_table([-1000], 48000, 8, False)
    for k in range(8):
        assert t[0][k] == np.pow(math.e, -2j * math.pi * (-1000 / 48000) * k)


def test_reference_generate_domain_table():
    """test/misc/read_file_test.py:12-19 of the reference."""
    assert orc.generate_domain('B') == (0, 1 / 255)
    assert orc.generate_domain('b') == (-128, 1 / 255)
    assert orc.generate_domain('h') == (-32768, 1 / 65535)
    assert orc.generate_domain('H') == (0, 1 / 65536)
    assert orc.generate_domain('i') == (-2147483648, 1 / 4294967295)
    assert orc.generate_domain('f') is None and orc.generate_domain('d') is None


def test_iq_correction_c_matches_python_loop():
    """iq_correction_test.py:24-49 compares against a python loop within 0.1; here exact."""
    rng = np.random.default_rng(5)
    z = rng.normal(size=500) + 1j * rng.normal(size=500) + (3 - 2j)
    a, b = z.copy(), z.copy()
    oa, ob = np.array([0.25 - 0.5j]), np.array([0.25 - 0.5j])
    orc.correct_iq(a, oa, 48000)
    orc.correct_iq_py(b, ob, 48000)
    assert np.array_equal(a, b) and oa[0] == ob[0]


@pytest.mark.parametrize('name', sorted(CASES))
def test_chain_matches_reference_golden(name):
    """Whole chain vs the reference run in-process.  Residual is numba fastmath (shiftFreq,
    angle, IQ loop) and pocketfft-vs-numpy FFT rounding: observed <= 1e-12."""
    raw, body, kw = case_stream(name)
    g = load_golden(name)
    assert sha(raw) == str(g['sha256']), 'synthetic input drifted from the fixture'
    ch = orc.Chain(**kw)
    assert ch.M == g['out'].shape[1] // int(g['nchunks'])
    z0 = ch.ingest(body[:orc.CHUNK_BYTES]).copy()
    assert rel_err(z0[:64], g['z0_head']) < 1e-14
    if kw['enc'] not in 'fd' and not kw['normalize'] and not kw['correct_iq']:
        assert np.array_equal(z0[:64], g['z0_head'])  # integer decode is exact
    ch = orc.Chain(**kw)
    out = ch.run(body)
    assert out.shape == g['out'].shape
    y0 = orc.Chain(**kw).decimated(z0[None, :])[0]
    assert rel_err(y0, g['y0']) < 1e-12
    assert rel_err(out, g['out']) < 2e-11


def test_stale_tail_and_chunk_count():
    """a trailing partial read is processed as a full chunk over the reused
    buffer; 3.37 chunks in -> 4 chunks out."""
    raw, body, kw = case_stream('c1_fm_wav_int16')
    ch = orc.Chain(**kw)
    out = ch.run(body)
    assert out.shape == (1, 4 * 512)


def test_plot_feed_restatements_match_scipy():
    """the oracle's slice-by-slice STFT equals scipy.signal.ShortTimeFFT.stft with the
    reference's parameters (waterfall_plot.py:44-51) exactly, its slice count equals SciPy's, and
    the power spectrum is the reference's expression (spectrum_analyzer_plot.py:77-82)."""
    from scipy.fft import fftn, fftshift
    from scipy.signal import ShortTimeFFT
    rng = np.random.default_rng(5)
    fs = 1_024_000
    S = ShortTimeFFT.from_window(('kaiser', 5), fs, 256, 128, mfft=1024, fft_mode='centered', scale_to='magnitude',
                                 phase_shift=None)
    for n in (128, 129, 1000, 8192, 32768):
        assert orc.stft_geometry(n) == (-S.m_num_mid, S.p_max(n) - S.p_min) and S.p_min == 0
    y = rng.standard_normal(4096) + 1j * rng.standard_normal(4096)
    sh = orc.plot_shift(15000, fs, y.size)
    want = 10. * np.log10(abs(S.stft(y * sh)))
    got = orc.stft_db(y, sh, S.win, S.hop, 1024, S.p_max(y.size))
    assert got.shape == want.shape == (1024, 33)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-9)
    assert np.array_equal(sh, np.exp(-2j * np.pi * (15000 / fs) * np.arange(y.size)))
    z = np.array([y * sh])
    amp = abs(fftshift(fftn(z, norm='forward')))
    assert np.array_equal(orc.power_spectrum(y, sh), np.log10(amp * amp)[0])
