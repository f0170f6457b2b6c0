import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); mi=hdr.index("Metric Name")
for r in rows[1:]: print(r[ki][:40], r[mi], r[vi])
