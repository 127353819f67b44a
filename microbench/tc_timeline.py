"""k_tc pipeline timeline of CTA 0 (diagnostic): SDRB_TC_DEBUG=1 python microbench/tc_timeline.py"""
import os, sys
os.environ['SDRB_TC_DEBUG'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import numpy as np, torch
import signals
from sdrterm_b200 import _native as nat
from sdrterm_b200.engine import Engine
from sdrterm_b200.plan import build_plan
nch = 8192
pl = build_plan(1_024_000, 'h', 64, [15000], correct_iq=True, demod='fm', omega_out=5000)
eng = Engine(pl, max_chunks=nch)
base = np.frombuffer(signals.c1_bytes(16 * 32768, seed=0, header=False), dtype=np.uint8)
raw = torch.from_numpy(np.tile(base, nch // 16)).cuda()
out = torch.empty((1, nch * pl.M), dtype=torch.float64, device='cuda')
for _ in range(3):
    eng.process_device(raw.data_ptr(), nch, out.data_ptr(), 0)
torch.cuda.synchronize()
buf = np.zeros(1024, dtype=np.uint64)
nat.check(nat.lib().sdrb_read_debug(eng._h, buf.ctypes.data), eng._h)
t = buf.reshape(64, 16).astype(np.int64)
t0 = t[0, 0]
names = ['tma_issue', 'xor_start', 'xor_done', 'mma_start', 'mma_issued', 'epi_start', 'tmem_free', 'epi_end',
         'iq', 'scanF', 'dotF', 'scanB']
print('tile ' + ' '.join(f'{n:>10s}' for n in names))
for it in range(3, 32, 3):      # epilogue events are recorded by warp 4 (epilogue group 0: every third tile)
    print(f'{it:4d} ' + ' '.join(f'{(t[it, e] - t0):10d}' for e in range(12)))
ev = np.arange(9, 31, 3)
d = np.diff(t[8:31, 0])
print('cycles per MMA tile (tma_issue to tma_issue):', d.mean())
for a, b in ((0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 8), (8, 9), (9, 10), (10, 11), (11, 7), (5, 7), (0, 7)):
    print(f'{names[a]:>10s} -> {names[b]:<10s}: mean {np.mean(t[ev, b] - t[ev, a]):8.0f}')

f = t[32:48, :11]          # k_finish events of warp 0 / CTA 0 live in the second half of the buffer
fn = ['start', '1a loads', '1b iq', '1c nco', '1d dots', '1e tiles', '1f zeta', '2 outputs+phase', '3b fft', '4 sos', '5 store']
print('k_finish phases of warp 0 / CTA 0 (cycles, mean over its items 1..3):')
for e in range(10):
    print(f'  {fn[e + 1]:>16s}: {np.mean(f[1:4, e + 1] - f[1:4, e]):8.0f}')
print(f'  {"item":>16s}: {np.mean(f[1:4, 10] - f[1:4, 0]):8.0f}')
