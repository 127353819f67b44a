"""Config 3 (16 VFOs + centre = 17 rows, int16 big-endian, fs 2.4 MS/s, FM, -d 64) on one GPU, a few
batches: the command profiled for profiles/*simo* (tensor-pipe utilisation of the channelizer GEMM).
    python microbench/simo_run.py [chunks] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import torch

import signals
from bench import synth_c3_device
from sdrterm_b200 import multigpu

sch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device('cuda', 0)
offs = signals.vfo_grid(16, 100_000)
rows = list(offs) + [0]
bank = multigpu.RowShardedBank(2_400_000, 'h', 64, rows, sch, 0, None, torch, demod='fm', swap=True, omega_out=5000)
raw = synth_c3_device(torch, sch * 32768, 3, dev, offs)
out = torch.empty((len(rows), sch * bank.M), dtype=torch.float64, device=dev)
bank.run([raw] * 2, sch, [out] * 2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
bank.run([raw] * steps, sch, [out] * steps)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f'{len(rows)} rows, {sch} chunks: {ms:.4f} ms per batch = {sch * 32768 / ms / 1e3:.1f} M input samples/s = '
      f'{sch * 32768 * len(rows) / ms / 1e6:.1f} G VFO*samples/s, front end {"k_tc" if bank.engine.tc is not None else "k_main"}')
