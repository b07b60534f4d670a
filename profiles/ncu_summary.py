import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print(d['Kernel Name'][:60], d['Grid Size'], d['Block Size'])
    keys=['gpu__time_duration.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','launch__occupancy_limit_registers','launch__waves_per_multiprocessor','smsp__thread_inst_executed_per_inst_executed.ratio','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_write.sum','sm__cycles_elapsed.max','l1tex__data_pipe_lsu_wavefronts.sum','smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_shared_st.sum','smsp__inst_executed_op_global_ld.sum','smsp__inst_executed_op_global_st.sum','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum']
    for k in keys:
        if k in d: print('  ',k, d[k], units[hdr.index(k)])
    st=[(float(d[k]),k) for k in hdr if 'issue_stalled' in k and k.endswith('_per_warp_active.pct') and d[k]]
    for v,k in sorted(st,reverse=True)[:8]: print('   stall',k.replace('smsp__average_warp_latency_issue_stalled_','').replace('smsp__average_warps_issue_stalled_','').replace('_per_warp_active.pct',''),v)
