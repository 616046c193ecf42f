"""Print the headline metrics of an `ncu --page raw --csv` export (first kernel)."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1]))); h=rows[0]; r=rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active_realtime.avg.pct','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
 'sm__pipe_fp64_cycles_active_realtime','smsp__issue_active.avg.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct','l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum','l1tex__data_pipe_tc_wavefronts_mem_shared.sum',
 'lts__throughput.avg.pct','lts__t_bytes.sum ','launch__registers_per_thread','sm__cycles_elapsed.max','smsp__inst_executed.sum','sm__throughput.avg.pct','gpu__dram_throughput.avg.pct','lts__t_sectors_srcunit_tex_op_read.sum','sm__inst_executed_pipe_tensor']
for i,c in enumerate(h):
    if any(c.startswith(w) or w in c for w in want) and '.max.' not in c and '.min.' not in c:
        print(f'{c:95s} {rows[1][i]:12s} {r[i]}')

# --json OUT: the figures bench.py quotes from a committed capture (dram bytes per launch, pipe utilisation of the launch)
if '--json' in sys.argv:
    import json
    out_path = sys.argv[sys.argv.index('--json') + 1]
    def val(name):
        for i, c in enumerate(h):
            if c == name:
                try: return float(r[i].replace(',', ''))
                except Exception: return None
        return None
    unit = {c: rows[1][i] for i, c in enumerate(h)}
    def to_bytes(name):
        v = val(name)
        if v is None: return None
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit.get(name, 'byte'), 1)
    rd, wr = to_bytes('dram__bytes_read.sum'), to_bytes('dram__bytes_write.sum')
    kname = r[h.index('Kernel Name')] if 'Kernel Name' in h else None
    json.dump({'dram_bytes_per_launch': (rd or 0) + (wr or 0), 'dram_bytes_read': rd, 'dram_bytes_write': wr,
               'source': 'ncu --set full --clock-control none, ' + sys.argv[1].split('/')[-1] + ' (profiles/), kernel ' + str(kname),
               'binding': {'issue_slots_pct': val('smsp__issue_active.avg.pct_of_peak_sustained_active'),
                           'xu_pipe_pct': val('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
                           'tensor_pipe_active_pct': val('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'),
                           'duration_ms': val('gpu__time_duration.sum')}},
              open(out_path, 'w'), indent=1)
