"""Parse ONE kernel launch of an `ncu --set full` report into the small JSON bench.py reads for `roofline.traffic`:
    python tools/ncu_traffic.py gpurun_out/<name>.ncu-rep profiles/gemm_traffic.json <algorithmic bytes> "<what was captured>"
dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum of that launch."""
import csv
import json
import subprocess
import sys

rep, out, alg, what = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, first = rows[0], rows[1], rows[2]
d = dict(zip(hdr, first))
u = dict(zip(hdr, units))


def to_bytes(key):
    v = float(d[key].replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[key]]


rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
res = {"kernel": d["Kernel Name"], "dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "algorithmic_bytes": alg,
       "duration_us_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) * {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u["gpu__time_duration.sum"], 1.0),
       "source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full --clock-control none: %s (%s)" % (what, rep.split("/")[-1])}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
