import csv,collections,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); im=hdr.index("Metric Name"); iv=hdr.index("Metric Value"); iid=hdr.index("ID")
d=collections.defaultdict(dict)
for r in rows[1:]:
    d[(r[iid],r[ik][:60])][r[im]]=float(r[iv].replace(',',''))
agg=collections.defaultdict(list)
for (i,k),m in d.items(): agg[k].append(m)
for k,ms in sorted(agg.items(), key=lambda kv:-sum(m['gpu__time_duration.sum'] for m in kv[1])):
    n=len(ms); t=sum(m['gpu__time_duration.sum'] for m in ms)/n
    if t<2e4: continue
    rd=sum(m['dram__bytes_read.sum'] for m in ms)/n; wr=sum(m['dram__bytes_write.sum'] for m in ms)/n
    print(f"{n:3d} {t/1e3:9.1f} us  rd {rd/1e6:7.1f} MB wr {wr/1e6:7.1f} MB  {(rd+wr)/t:7.1f} GB/s  {k}")
