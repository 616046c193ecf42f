"""Print the headline metrics of an `ncu --page raw --csv` export (first kernel)."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1]))); h=rows[0]; r=rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active_realtime.avg.pct','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
 'sm__pipe_fp64_cycles_active_realtime','smsp__issue_active.avg.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct','l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum','l1tex__data_pipe_tc_wavefronts_mem_shared.sum',
 'lts__throughput.avg.pct','lts__t_bytes.sum ','launch__registers_per_thread','sm__cycles_elapsed.max','smsp__inst_executed.sum','sm__throughput.avg.pct','gpu__dram_throughput.avg.pct','lts__t_sectors_srcunit_tex_op_read.sum','sm__inst_executed_pipe_tensor']
for i,c in enumerate(h):
    if any(c.startswith(w) or w in c for w in want) and '.max.' not in c and '.min.' not in c:
        print(f'{c:95s} {rows[1][i]:12s} {r[i]}')
