#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + source page) into the handful of numbers DESIGN.md quotes."""
import csv, subprocess, sys, io

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']

def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))

def main(rep, top=25):
    r = page(rep, 'raw')
    hdr, units, vals = r[0], r[1], r[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print(f"== {rep}\nkernel: {d.get('Kernel Name', ('?',))[0][:100]}")
    for k in KEYS:
        if k in d:
            print(f"{k:75s} {d[k][0]:>18s} {d[k][1]}")
    rows = page(rep, 'source')
    h = rows[1]; data = rows[2:]
    ix = {n: i for i, n in enumerate(h)}
    f = lambda x: float(x) if x not in ('', None) else 0.0
    tot = sum(f(r_[ix['# Samples']]) for r_ in data) or 1.0
    stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
    agg = {n: sum(f(r_[ix[n]]) for r_ in data) for n in stalls}
    print("stall samples by reason (all warps):")
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        print(f"   {n:28s} {100 * v / tot:6.2f}%")
    print(f"top {top} instructions by samples:")
    for r_ in sorted(data, key=lambda r_: -f(r_[ix['# Samples']]))[:top]:
        s = f(r_[ix['# Samples']])
        st = sorted(((f(r_[ix[n]]), n) for n in stalls), reverse=True)[0]
        print(f"   {100 * s / tot:5.2f}% {r_[ix['Source']][:64]:64s} x{r_[ix['Instructions Executed']]:>11s} {st[1]}")

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
