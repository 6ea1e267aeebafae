"""profiles/ncu_counters.json: the per-launch counters bench.py quotes as "offline ncu capture" (DRAM traffic and warp
instructions of the fused step kernel), extracted from an `ncu --set full` report instead of being typed in.

    ncu -i gpurun_out/r02_step.ncu-rep --page raw --csv | python tools/ncu_counters.py hrp_step_kernel 4096 profiles/r02_step_kernel_ncu_full.csv
"""
import csv
import json
import os
import sys

kernel, envs, source = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(sys.stdin))
h, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(h)}
picked = [r for r in data if kernel in r[col["Kernel Name"]]]
assert picked, f"no launch of {kernel} in the report"


def val(r, name):
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "": 1.0}.get(u, 1.0) if "byte" in u else v


def dur_us(r):
    v, u = float(r[col["gpu__time_duration.sum"]].replace(",", "")), units[col["gpu__time_duration.sum"]]
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)


r = picked[-1]
out = {"envs": envs, "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
       "warp_instructions": val(r, "smsp__inst_executed.sum"),
       "thread_instructions_per_warp_instruction": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
       "registers": int(val(r, "launch__registers_per_thread")), "duration_us_under_ncu": dur_us(r), "source": source}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_counters.json")
allc = json.load(open(path)) if os.path.exists(path) else {}
allc[kernel] = out
json.dump(allc, open(path, "w"), indent=1)
print(json.dumps(out))
