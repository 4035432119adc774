import csv, json, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
n_step=int(sys.argv[2])
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit"); ii=h.index("ID")
gi=h.index("Grid Size"); bi=h.index("Block Size")
per={}
for r in rows[1:]:
    d=per.setdefault(int(r[ii]),{"k":r[ki],"grid":r[gi],"block":r[bi]})
    v=float(r[vi].replace(",",""))
    if r[ui]=="ns": v/=1e3
    d[r[mi]]=v
ids=sorted(per)[-n_step:]
agg=collections.OrderedDict()
for i in ids:
    d=per[i]; n=d["k"].split("(")[0].replace("void ","")
    cls="gemm_tcgen05_kernel" if "gemm_tcgen05" in n else n.split("<")[0]
    a=agg.setdefault(cls,[0,0.0,0.0,0.0]); a[0]+=1; a[1]+=d["gpu__time_duration.sum"]; a[2]+=d["dram__bytes_read.sum"]; a[3]+=d["dram__bytes_write.sum"]
tot=sum(a[1] for a in agg.values())
out={}
print(f"# ncu launch list, last step ({n_step} launches): gpu__time_duration + DRAM bytes (cold-cache, serialised, --clock-control none)\n")
print(f"total {tot:.1f} us\n")
print("| kernel | launches | total us | share | DRAM read MB | DRAM write MB | MB per launch |\n|---|---|---|---|---|---|---|")
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"| {k} | {a[0]} | {a[1]:.1f} | {100*a[1]/tot:.1f}% | {a[2]/1e6:.1f} | {a[3]/1e6:.1f} | {(a[2]+a[3])/a[0]/1e6:.1f} |")
    out[k]={"launches_per_step":a[0],"gpu_time_us_total":round(a[1],1),"dram_bytes_read":a[2],"dram_bytes_write":a[3],"dram_bytes_per_launch":(a[2]+a[3])/a[0]}
print("\n| # | kernel | grid | block | us | DRAM MB |\n|---|---|---|---|---|---|")
for j,i in enumerate(ids):
    d=per[i]
    print(f"| {j} | {d['k'].split('(')[0].replace('void ','')} | {d['grid']} | {d['block']} | {d['gpu__time_duration.sum']:.1f} | {(d['dram__bytes_read.sum']+d['dram__bytes_write.sum'])/1e6:.1f} |")
md = sys.argv[4] if len(sys.argv) > 4 else "profiles/r01_launches_v6.md"
out["gemm_tcgen05_kernel"]["source"]="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over one bench.py step (B=256, L=128): " + md + "; writes still dirty in L2 at kernel end are charged to later kernels"
json.dump(out,open(sys.argv[3],"w"),indent=1)
