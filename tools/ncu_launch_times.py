import csv
for f in ("gpurun_out/corr_launches.csv","gpurun_out/corrsweep_launches.csv"):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value"); ig=hdr.index("Grid Size")
    print(f, [(r[ig], r[iv]) for r in rows[1:] if "corr_tc" in r[ik]][-5:])
