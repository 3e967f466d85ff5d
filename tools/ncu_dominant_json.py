"""profiles/r02_ncu_dominant.json from an `ncu --set full` report of tools/prof_block.py (full-resolution 32->32 block of
cfg2): DRAM bytes per launch of the dominant tcgen05 conv instantiation, which bench.py reports as roofline.traffic.
Usage: python tools/ncu_dominant_json.py gpurun_out/r02_block.ncu-rep [workload=cfg2]"""
import csv, json, os, subprocess, sys

rep = sys.argv[1]
workload = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]


def val(r, name):
    v = float(r[h.index(name)].replace(",", ""))
    u = units[h.index(name)].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3,
             "nsecond": 1e-3, "%": 1}.get(u, 1)
    return v * scale


launches = []
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    if "conv_tc_kernel" not in name:
        continue
    launches.append({"kernel": name[:80],
                     "dram_read_MB": val(r, "dram__bytes_read.sum") / 1e6,
                     "dram_write_MB": val(r, "dram__bytes_write.sum") / 1e6,
                     "duration_us": val(r, "gpu__time_duration.sum"),
                     "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")})
assert launches, "no conv_tc_kernel launch in the report"
traffic = sum(l["dram_read_MB"] + l["dram_write_MB"] for l in launches) / len(launches) * 1e6
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_ncu_dominant.json")
try:
    doc = json.load(open(path))
except (OSError, ValueError):
    doc = {}
doc[workload] = {"kernel": "conv_tc_kernel<32,32,4,S1K3,WRES> (full-res 32->32 3x3x3: fprop with fused GN statistics / dgrad "
                           "with fused GN-backward reduction)",
                 "launches": launches, "traffic_bytes_per_launch": traffic,
                 "source": f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum averaged over the {len(launches)} "
                           f"conv_tc launches of tools/prof_block.py ({os.path.basename(rep)}); profiles/r02_ncu_block.md"}
json.dump(doc, open(path, "w"), indent=1)
print(json.dumps(doc[workload], indent=1))
