"""profiles/r2_scan_traffic.json from an `ncu --set full` capture of the scan (tools/final_profiles.sh):
    python tools/make_scan_traffic.py gpurun_out/r2_scan_fwd_b32.ncu-rep 32"""
import csv, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, nseq = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
def get(name):
    i = hdr.index(name)
    v, u = float(vals[i].replace(",", "")), units[i]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1)
    return v * mult
src = os.path.join(ROOT, "audio-mamba-aum_b200", "csrc", "scan_fwd_tma.cu")
d = {"dram__bytes_read.sum": int(get("dram__bytes_read.sum")), "dram__bytes_write.sum": int(get("dram__bytes_write.sum")),
     "sequences_per_launch": nseq, "gpu__time_duration_us_under_ncu": get("gpu__time_duration.sum"),
     "kernel": vals[hdr.index("Kernel Name")],
     "kernel_source_sha256": hashlib.sha256(open(src, "rb").read()).hexdigest(),
     "how": "ncu --set full --clock-control none -k regex:scan_fwd_tma -s 2 -c 1, tools/scan_once.py PREGATED=1 DELTA16=1 B=%d (tools/final_profiles.sh)" % nseq}
json.dump(d, open(os.path.join(ROOT, "profiles", "r2_scan_traffic.json"), "w"), indent=1)
print(d)
