"""Device-resident throughput of the BASELINE.json configurations that are NOT the headline bench
line (they are parity-test cases; this records where they stand): C2 (uint8 AM, -d 50 and -d 64),
C3 (17 rows int16 SIMO), C4 (257 rows float32 SIMO).  One line per config: path taken, Msamples/s,
VFO*Msamples/s, fraction of the HBM roof for that config's algorithmic bytes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import torch
import signals
from sdrterm_b200.engine import Engine
from sdrterm_b200.plan import build_plan

PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0


def run(name, pl, nch, itemsize):
    eng = Engine(pl, max_chunks=nch)
    raw = torch.randint(0, 256, (nch * 131072,), dtype=torch.uint8, device='cuda')
    if pl.enc == 'f':
        raw = (torch.randn(nch * 131072 // 4, device='cuda') * 0.1).view(torch.uint8)
    out = torch.empty((pl.R, nch * pl.M), dtype=torch.float64, device='cuda')
    for _ in range(3):
        eng.process_device(raw.data_ptr(), nch, out.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        eng.process_device(raw.data_ptr(), nch, out.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    eng.set_profiling(True)
    eng.process_device(raw.data_ptr(), nch, out.data_ptr())
    torch.cuda.synchronize()
    kt = [round(v, 4) for v in eng.kernel_times()]     # front end | IQ kernels | k_finish or k_fixup | k_demod
    eng.set_profiling(False)
    nsamp = nch * pl.N
    msps = nsamp / (ms * 1e-3) / 1e6
    bps = 2 * itemsize + 8 * pl.R / pl.q
    print(json.dumps({'config': name, 'front_end': 'k_tc' if eng.tc is not None else 'k_main', 'rows': pl.R,
                      'chunks': nch, 'ms': ms, 'input_msps': msps, 'vfo_msps': msps * pl.R,
                      'bytes_per_sample': bps, 'frac_hbm': msps * 1e6 * bps / 1e9 / PEAK, 'kernel_ms': kt}))
    eng.close()


offs16 = signals.vfo_grid(16, 100_000)
offs256 = signals.vfo_grid(256, 200_000)
run('C2 uint8 AM -d 50 --correct-iq', build_plan(2_400_000, 'B', 50, [0], correct_iq=True, demod='am', omega_out=5000), 2048, 1)
run('C2 uint8 AM -d 64 --correct-iq', build_plan(2_400_000, 'B', 64, [0], correct_iq=True, demod='am', omega_out=5000), 2048, 1)
run('C3 int16be SIMO 16+1 FM -d 64', build_plan(2_400_000, 'h', 64, offs16 + [0], simo=True, swap=True, demod='fm', omega_out=5000), 512, 2)
run('C4 float32 SIMO 256+1 FM -d 64', build_plan(61_440_000, 'f', 64, offs256 + [0], simo=True, demod='fm', omega_out=12500), 64, 4)
