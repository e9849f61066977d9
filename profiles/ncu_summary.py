"""Summarises an `ncu --set full` report: per captured launch the duration, DRAM / L2->SM traffic, tensor-pipe and
issue utilisation, plus the top stall reasons.  usage: python profiles/ncu_summary.py report.ncu-rep [--json out.json]"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'regs/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dyn smem/block'),
    ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
    ('l1tex__m_xbar2l1tex_read_bytes.sum', 'L2->SM read bytes'),
    ('l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', '  of which TMA (weights)'),
    ('sm__sass_l1tex_m_xbar2l1tex_read_bytes_mem_global_op_ldgsts_cache_bypass.sum', '  of which cp.async gathers'),
    ('l1tex__m_l1tex2xbar_write_bytes.sum', 'SM->L2 write bytes'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput %'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg', 'tensor imma cycles active (avg/SM)'),
    ('sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg', 'tensor hmma cycles active (avg/SM)'),
    ('sm__cycles_active.avg', 'SM cycles active (avg)'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print('kernel:', d['Kernel Name'][:110])
        rec = {'kernel': d['Kernel Name']}
        for k, label in KEYS:
            hit = [h for h in hdr if h == k or h.endswith('.' + k)]
            if hit and d[hit[0]] not in ('', 'n/a'):
                print(f'  {label:40s} {d[hit[0]]} {u[hit[0]]}')
                rec[k] = (d[hit[0]], u[hit[0]])
        out.append(rec)
        print()
    if '--json' in sys.argv:
        with open(sys.argv[sys.argv.index('--json') + 1], 'w') as f:
            json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
