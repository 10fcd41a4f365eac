import csv,json
rows=list(csv.reader(open("gpurun_out/prof_attn_w128_raw.csv"))); hdr=rows[0]; un=rows[1]; idx={h:i for i,h in enumerate(hdr)}
out=[]
for r in rows[2:]:
    d={}
    for k in ["Kernel Name","launch__grid_size","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","smsp__inst_executed.sum","sm__cycles_elapsed.max","lts__t_bytes.sum","sm__throughput.avg.pct_of_peak_sustained_elapsed"]:
        d[k]=(r[idx[k]][:70], un[idx[k]])
        print(k, r[idx[k]][:70], un[idx[k]])
    out.append(d)
    print("--")
