#!/usr/bin/env python
"""Key readings of one `ncu --set full` raw page (ncu -i x.ncu-rep --page raw --csv > x.raw.csv) as JSON:
    python tools/ncu_extract.py profiles/x.raw.csv ["source note"]"""
import csv
import json
import sys

KEYS = {'Kernel Name': 'kernel', 'gpu__time_duration.sum': 'gpu_time_us', 'dram__bytes_read.sum': 'dram_read',
        'dram__bytes_write.sum': 'dram_write', 'sm__cycles_elapsed.avg.per_second': 'sm_clock_ghz',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pipe_active_pct',
        'sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed': 'utchmma_fp16_pct_of_peak',
        'launch__registers_per_thread': 'registers_per_thread', 'launch__grid_size': 'grid', 'launch__block_size': 'block',
        'lts__t_bytes.sum': 'l2_bytes', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'smem_bank_conflicts',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_throughput_pct',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram_throughput_pct',
        'smsp__inst_executed.sum': 'warp_instructions'}
UNIT = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    head, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for k, name in KEYS.items():
        if k in head:
            i = head.index(k)
            v = vals[i]
            try:
                v = float(v.replace(',', ''))
                if units[i] in UNIT:
                    v *= UNIT[units[i]]
                    name += '_bytes'
            except ValueError:
                pass
            out[name] = v
    if 'dram_read_bytes' in out:
        out['dram_bytes_read'], out['dram_bytes_write'] = out.pop('dram_read_bytes'), out.pop('dram_write_bytes')
    if len(sys.argv) > 2:
        out['source'] = sys.argv[2]
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
