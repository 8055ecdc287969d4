"""Summarise one kernel of an `ncu --set full` report (read on the build container: `ncu -i … --page raw --csv`)
into the small JSON committed under profiles/.  Usage: python tools/ncu_summary.py report.ncu-rep out.json [launch#]"""
import csv
import json
import subprocess
import sys

KEEP = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
    'l1tex__m_xbar2l1tex_read_bytes.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor', 'sm__cycles_elapsed.avg',
    'sm__cycles_active.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'smsp__sass_inst_executed_op_tmem_ldt.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(line for line in txt.splitlines() if line.startswith('"')))
    head, units, data = rows[0], rows[1], rows[2 + which]
    rec = {}
    for name, unit, val in zip(head, units, data):
        if name in KEEP:
            rec[name] = dict(value=val.replace(',', ''), unit=unit)
    rec['_kernel'] = data[head.index('Kernel Name')]
    rec['_source'] = f'ncu --set full --clock-control none --import-source on ({rep.split("/")[-1]}), launch {which}'
    json.dump(rec, open(out, 'w'), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == '__main__':
    main()
