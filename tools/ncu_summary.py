import csv, sys, subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','details','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines())); h=rows[0]
keep=('Duration','DRAM Throughput','Memory Throughput','L2 Hit Rate','L1/TEX Hit Rate','Registers Per Thread','Achieved Occupancy','Theoretical Occupancy','Executed Ipc Active','Issue Slots Busy','Block Size','Grid Size','Mem Busy','Max Bandwidth','Compute (SM) Throughput','Warp Cycles Per Issued Instruction','Avg. Active Threads Per Warp','L2 Cache Throughput')
for row in rows[1:]:
    d=dict(zip(h,row))
    if d.get('Metric Name') in keep: print(d['Section Name'],'|',d['Metric Name'],'|',d['Metric Unit'],'|',d['Metric Value'])
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines())); hdr,units,vals=rows[0],rows[1],rows[2]
for k in ('dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','gpu__time_duration.sum','lts__t_sector_hit_rate.pct','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio'):
    for hh,u,v in zip(hdr,units,vals):
        if hh==k: print(f"raw | {hh} | {u} | {v}")
