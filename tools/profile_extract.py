"""Turns the raw artefacts of tools/profile_run.sh (gpurun_out/) into the committed evidence under profiles/."""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput", "sm__throughput",
        "sm__warps_active", "launch__", "smsp__inst_executed.sum", "sm__inst_executed_pipe", "smsp__cycles_active.avg",
        "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate", "sm__cycles_active.avg", "smsp__warp_issue_stalled",
        "smsp__average_warp", "sm__pipe_fma", "sm__inst_executed.avg.per_cycle", "gpu__dram_throughput", "dram__cycles_active")

for name in ("bench.json", "bench_reference.json", "c3_1024.json", "c3_4096.json", "cqpsk_bench.jsonl", "bench_cqpsk.json", "bench_fec.json"):
    src = os.path.join(OUT, name)
    if os.path.exists(src) and os.path.getsize(src) > 2:
        shutil.copy(src, os.path.join(PROF, f"{TAG}_{name}"))
src = os.path.join(OUT, "launches.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(PROF, f"{TAG}_launches.csv"))
traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures "
                       f"(profiles/{TAG}_ncu_*.csv), C2 step size"}
for rep in sorted(os.listdir(OUT)):
    if not (rep.startswith("ncu_") and rep.endswith(".ncu-rep")):
        continue
    raw = subprocess.run(["ncu", "-i", os.path.join(OUT, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for launch in rows[2:]:
        kname = launch[hdr.index("Kernel Name")].split("(")[0].split("::")[-1].split("<")[0]
        dst = os.path.join(PROF, f"{TAG}_ncu_{kname}.csv")
        rd = wr = None
        with open(dst, "w") as f:
            f.write("metric,unit,value\n")
            f.write(f"kernel,,{launch[hdr.index('Kernel Name')]}\n")
            f.write(f"grid,,{launch[hdr.index('Grid Size')]}\nblock,,{launch[hdr.index('Block Size')]}\n")
            for h, u, v in zip(hdr, units, launch):
                if any(k in h for k in KEEP):
                    f.write(f"{h},{u},{v}\n")
                if h == "dram__bytes_read.sum":
                    rd = (float(v), u)
                if h == "dram__bytes_write.sum":
                    wr = (float(v), u)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        if rd and wr:
            traffic[kname] = int(rd[0] * scale.get(rd[1], 1.0) + wr[0] * scale.get(wr[1], 1.0))
        print("wrote", dst)
with open(os.path.join(PROF, "traffic.json"), "w") as f:
    json.dump(traffic, f, indent=2)
print(json.dumps(traffic, indent=1))
