import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
title, cmd = sys.argv[2], sys.argv[3]
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__grid_size','launch__block_size','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
print("# %s\n" % title)
print("Command: `%s`\n" % cmd)
for r in rows[2:]:
    print("## %s\n" % r[hdr.index('Kernel Name')])
    print("| metric | value |\n|---|---|")
    for w in want:
        if w in hdr:
            i = hdr.index(w); print("| %s | %s %s |" % (w, r[i], units[i]))
    tot = 0; items=[]
    for i,h in enumerate(hdr):
        if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued'):
            v = float(r[i].replace(',','')); items.append((v,h)); tot += v
    if tot: print("\nWarp stall sampling (share of samples): " + ", ".join("%s %.1f%%" % (h.replace('smsp__pcsamp_warps_issue_stalled_',''), 100*v/tot) for v,h in sorted(items, reverse=True)[:8]) + "\n")
