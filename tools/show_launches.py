import csv, collections, sys
path = sys.argv[1] if len(sys.argv) > 1 else '/root/repo/gpurun_out/step_launches.csv'
lines=[l for l in open(path) if not l.startswith('==')]
rows=list(csv.DictReader(lines))
by=collections.OrderedDict()
for r in rows:
    key=(r['ID'], r['Kernel Name'][:44])
    by.setdefault(key,{})[r['Metric Name']]=(r['Metric Value'], r['Metric Unit'])
tot=0
for (i,k),m in by.items():
    def g(n):
        v,u=m.get(n,('0',''))
        return v+u
    t=float(m['gpu__time_duration.sum'][0].replace(',','')); u=m['gpu__time_duration.sum'][1]
    t = t/1e3 if u=='ns' else t*1e3 if u=='ms' else t
    tot+=t
    print("%4s %-44s t=%9.1fus rd=%-16s wr=%-16s sm=%-6s dram=%-6s"%(i,k,t,g('dram__bytes_read.sum'),g('dram__bytes_write.sum'),g('sm__throughput.avg.pct_of_peak_sustained_elapsed')[:5],g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')[:5]))
print("total %.1f us"%tot)
