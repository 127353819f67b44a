"""Rate of ONE rank's share of a bank against the number of rows it holds (no broadcast): what a
row-group split costs on the kernels alone.  Per-kernel CUDA-event times from the library.
    python microbench/bank_rows.py f 256 257 129 65 33      (encoding, chunks, rows...)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import torch
import signals
from sdrterm_b200.engine import Engine
from sdrterm_b200.plan import build_plan

enc = sys.argv[1] if len(sys.argv) > 1 else 'f'
nch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
widths = [int(v) for v in sys.argv[3:]] or [257, 129, 65, 33]
fs = 61_440_000 if enc == 'f' else 2_400_000
offs = signals.vfo_grid(256, 200_000) if enc == 'f' else signals.vfo_grid(16, 100_000)
rows_all = list(offs) + [0]
for R in widths:
    pl = build_plan(fs, enc, 64, rows_all[:R], simo=True, swap=(enc != 'f'), demod='fm', omega_out=12500 if enc == 'f' else 5000)
    eng = Engine(pl, max_chunks=nch)
    if enc == 'f':
        raw = (torch.randn(nch * 131072 // 4, device='cuda') * 0.1).view(torch.uint8)
    else:
        raw = torch.randint(0, 256, (nch * 131072,), dtype=torch.uint8, device='cuda')
    out = torch.empty((R, nch * pl.M), dtype=torch.float64, device='cuda')
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
    kt = eng.kernel_times()
    eng.set_profiling(False)
    print(json.dumps({'enc': enc, 'rows': R, 'chunks': nch, 'ms': round(ms, 4), 'ms_per_row': round(ms / R, 5),
                      'vfo_gsps': round(nch * pl.N * R / ms / 1e6, 1),
                      'front_end': 'k_tc' if eng.tc is not None else 'k_main', 'kernel_ms': [round(v, 4) for v in kt]}))
    eng.close()
