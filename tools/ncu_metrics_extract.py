import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]
keep=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_tensor.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__grid_size","launch__block_size","launch__waves_per_multiprocessor","l1tex__t_sector_hit_rate.pct","lts__t_sector_hit_rate.pct","l1tex__throughput.avg.pct_of_peak_sustained_elapsed","lts__throughput.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct_of_peak_sustained_elapsed","smsp__inst_executed.sum","l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum","sm__cycles_elapsed.avg","smsp__cycles_active.avg"]
stall=[h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
with open(sys.argv[2],"w") as f:
    for r in rows[2:]:
        d=dict(zip(hdr,r))
        for k in keep:
            if k in d: f.write("%-70s %s\n"%(k,d[k]))
        st=[]
        for k in stall:
            try: st.append((float(d[k].replace(",","")),k.replace("smsp__pcsamp_warps_issue_stalled_","")))
            except: pass
        f.write("stall samples (top): "+", ".join("%s %d"%(n,v) for v,n in sorted(st,reverse=True)[:8])+"\n\n")
