import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
for r in rows[1:][-int(sys.argv[2]):]: print(r[ki][:70], r[vi])
