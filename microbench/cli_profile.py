"""Where the wall time of `python -m sdrterm -i <wav> -o <file> ...` goes (in-process, timers around the
reader's readinto / pool waits and the consumer's queue waits, submits, waits and writes).
    python microbench/cli_profile.py [chunks]"""
import os, sys, time, wave
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import numpy as np

nch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
path, out = '/dev/shm/cli_prof.wav', '/dev/shm/cli_prof.out'
rng = np.random.default_rng(0)
blk = rng.integers(-2000, 2000, size=16 * 32768 * 2, dtype=np.int16)
with wave.open(path, 'wb') as w:
    w.setnchannels(2); w.setsampwidth(2); w.setframerate(1_024_000)
    for _ in range(nch // 16):
        w.writeframesraw(blk.tobytes())

T = {}


def timed(name, fn):
    def wrap(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            T[name] = T.get(name, 0.0) + time.perf_counter() - t
    return wrap


t_import = time.perf_counter()
from sdrterm_b200 import sdrterm as cli
from sdrterm_b200.dsp import dsp_processor as dp
from sdrterm_b200.misc import read_file as rf
from sdrterm_b200 import engine as eng_mod
T['import'] = time.perf_counter() - t_import

eng_mod.Engine.submit = timed('engine.submit', eng_mod.Engine.submit)
eng_mod.Engine.wait = timed('engine.wait', eng_mod.Engine.wait)
eng_mod.Engine.__init__ = timed('Engine()', eng_mod.Engine.__init__)
dp.DspProcessor._makeEngine = timed('_makeEngine (plan + Engine)', dp.DspProcessor._makeEngine)
dp.DspProcessor._staging = timed('_staging (pinned alloc)', dp.DspProcessor._staging)
dp.DspProcessor._emit = timed('_emit (write)', dp.DspProcessor._emit)
dp.DspProcessor._next = timed('consumer: queue wait', dp.DspProcessor._next)
eng_mod.Engine.close = timed('Engine.close', eng_mod.Engine.close)
dp.DspProcessor._drain = timed('_drain (wait + write)', dp.DspProcessor._drain)
dp.DspProcessor._processData = timed('_processData total', dp.DspProcessor._processData)
dp.DspProcessor.processData = timed('processData total (incl. close)', dp.DspProcessor.processData)
cli.checkWavHeader = timed('checkWavHeader', cli.checkWavHeader)
cli.makeProcessor = timed('makeProcessor', cli.makeProcessor)
rf.ChunkPool.get = timed('reader: pool wait', rf.ChunkPool.get)
rf.ChunkPool.__init__ = timed('ChunkPool() (pinned alloc)', rf.ChunkPool.__init__)
_chunks = rf.chunks


def chunks(reader, *a, **k):
    class R:
        def readinto(self, mv):
            t = time.perf_counter()
            n = reader.readinto(mv)
            T['reader: readinto'] = T.get('reader: readinto', 0.0) + time.perf_counter() - t
            return n
    return _chunks(R(), *a, **k)


rf.chunks = chunks
t0 = time.perf_counter()
cli.main(['-i', path, '-o', out, '-c', '15k', '-w', '5k', '-d', '64', '--correct-iq'])
wall = time.perf_counter() - t0
print(f'{nch} chunks ({nch * 131072 / 2**30:.2f} GiB): wall {wall:.3f} s in main() = {nch * 32768 / wall / 1e6:.1f} Msamples/s')
for k, v in sorted(T.items(), key=lambda kv: -kv[1]):
    print(f'  {k:32s} {v:7.3f} s')
os.remove(path); os.remove(out)
