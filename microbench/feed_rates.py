"""Rate of the plot consumers' FFT feed through the public calls (host arrays in and
out, so H2D / D2H and the allocations are inside the timed region) beside the reference's own
arithmetic (scipy.fft / scipy.signal.ShortTimeFFT) on this host: frames per second and Msamples/s.
    python microbench/feed_rates.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
from scipy.fft import fftn, fftshift
from scipy.signal import ShortTimeFFT
from sdrterm_b200.plots import SpectrumFeed, WaterfallFeed, powerSpectrum

fs, n = 1_024_000, 32768
rng = np.random.default_rng(0)
y = rng.standard_normal(n) + 1j * rng.standard_normal(n)


def rate(fn, reps):
    fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    return reps / (time.perf_counter() - t)


sf, wf = SpectrumFeed(fs, center=15000), WaterfallFeed(fs, center=15000)
sh = np.exp(-2j * np.pi * (15000 / fs) * np.arange(n))
S = ShortTimeFFT.from_window(('kaiser', 5), fs, 256, 128, mfft=1024, fft_mode='centered', scale_to='magnitude', phase_shift=None)
rows = np.stack([y] * 64)


def ref_spec():
    z = np.array([y * sh])
    a = abs(fftshift(fftn(z, norm='forward')))
    return np.log10(a * a)


def ref_spec64():
    z = rows * sh
    a = abs(fftshift(np.fft.fft(z, axis=1, norm='forward'), axes=1))
    return np.log10(a * a)


res = {'n': n,
       'spectrum_fps_device': rate(lambda: sf.update(y), 200), 'spectrum_fps_scipy': rate(ref_spec, 200),
       'spectrum_batch64_fps_device': 64 * rate(lambda: powerSpectrum(rows, sh), 20),
       'spectrum_batch64_fps_numpy': 64 * rate(ref_spec64, 5),
       'waterfall_fps_device': rate(lambda: wf.update(y), 100),
       'waterfall_fps_scipy': rate(lambda: 10. * np.log10(abs(S.stft(y * sh))), 20)}
for k in list(res):
    if k.endswith('_device'):
        res[k.replace('fps_device', 'msps_device')] = res[k] * n / 1e6
print(json.dumps(res))
