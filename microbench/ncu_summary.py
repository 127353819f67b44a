#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` report for profiles/ (runs where ncu is installed,
no GPU needed):  python microbench/ncu_summary.py gpurun_out/prof.ncu-rep "<command that was profiled>"

Per kernel: duration, DRAM bytes, pipe utilisations, occupancy, instruction count, warp-stall
reasons, and the source lines with the most stall samples (needs -lineinfo)."""
import collections
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed.avg.per_cycle_elapsed']


def ncu(rep, *args):
    return subprocess.run(['ncu', '-i', rep, *args, '--csv'], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    cmd = sys.argv[2] if len(sys.argv) > 2 else ''
    print(f'# ncu --set full --clock-control none --import-source on: {cmd}')
    print('# per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes\n')
    rows = list(csv.reader(io.StringIO(ncu(rep, '--page', 'raw'))))
    hdr, units = rows[0], rows[1]
    names = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d['Kernel Name']
        names.append(name)
        print(f'Kernel Name: {name}')
        for k in WANT:
            if k in d:
                print(f'{k}: {d[k]} {units[hdr.index(k)]}')
        st = {k: float(v.replace(',', '')) for k, v in d.items()
              if 'smsp__pcsamp_warps_issue_stalled' in k and not k.endswith('_not_issued') and v}
        tot = sum(st.values()) or 1.0
        top = sorted(st.items(), key=lambda kv: -kv[1])[:9]
        print('warp-stall samples by reason (%): ' + ', '.join(f"{k.split('stalled_')[1]} {100 * v / tot:.1f}" for k, v in top))
        try:
            rd, wr = float(d['dram__bytes_read.sum'].replace(',', '')), float(d['dram__bytes_write.sum'].replace(',', ''))
            un = units[hdr.index('dram__bytes_read.sum')]
            print(f'dram bytes per launch: {rd + wr:.3f} {un}')
        except Exception:
            pass
        print()
    # source lines
    rows = list(csv.reader(io.StringIO(ncu(rep, '--page', 'source', '--print-source', 'cuda,sass'))))
    agg = collections.defaultdict(lambda: collections.defaultdict(int))
    src = {}
    cur_file = cur_fn = None
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
        elif r[0] == 'Function Name':
            cur_fn = r[1]
        elif r[0].isdigit():
            try:
                s = int(r[4])
            except (ValueError, IndexError):
                continue
            agg[cur_fn][(cur_file, int(r[0]))] += s
            src[(cur_file, int(r[0]))] = r[1][:110]
    for fn, lines in agg.items():
        tot = sum(lines.values()) or 1
        print(f'top source lines by warp-stall samples: {fn[:70]} ({tot} samples)')
        for (f, ln), s in sorted(lines.items(), key=lambda kv: -kv[1])[:14]:
            print(f'  {100 * s / tot:5.1f}%  {f}:{ln}  {src[(f, ln)]}')
        print()


if __name__ == '__main__':
    main()
