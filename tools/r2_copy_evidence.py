"""Copy the outputs of tools/r2_evidence.sh (gpurun_out/ev_*) into profiles/ under their tracked names and regenerate
profiles/r2_traffic.json from the raw ncu exports.  Run here, after the evidence run has been merged back."""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
SRC = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")

COPIES = {
    "ev_bench_cfg1.json": "r2_bench_cfg1.json", "ev_bench_cfg2.json": "r2_bench_cfg2.json", "ev_bench_cfg3.json": "r2_bench_cfg3.json",
    "ev_bench_cfg5.json": "r2_bench_cfg5.json", "ev_bench_cfg4_1gpu_tf32.json": "r2_bench_cfg4_1gpu_tf32.json",
    "ev_bench_cfg4_1gpu_fp32.json": "r2_bench_cfg4_1gpu_fp32.json", "ev_bench_features.json": "r2_bench_features.json",
    "ev_ggs_cfg3_raw.csv": "r2_ggs_cfg3_ncu_raw.csv", "ev_ggs_cfg5_raw.csv": "r2_ggs_cfg5_ncu_raw.csv",
    "ev_den_n20_raw.csv": "r2_denoiser_n20_ncu_raw.csv", "ev_tc_b8_raw.csv": "r2_tc_linear_b8_ncu_raw.csv",
    "ev_launches_cfg3.csv": "r2_launches_cfg3.csv", "ev_probe_ggs_cfg3.txt": "r2_probe_ggs_cfg3.txt",
    "ev_probe_ggs_cfg5.txt": "r2_probe_ggs_cfg5.txt", "ev_probe_den_n20.txt": "r2_probe_den_n20.txt",
    "ev_tc_gemm_probe.txt": "r2_tc_gemm_probe.txt", "ev_stage_probe.txt": "r2_negative_den_flagged_stage_probe.txt",
    "ev_gputests.log": "r2_gputests.log",
}
for a, b in COPIES.items():
    p = os.path.join(SRC, a)
    if os.path.exists(p) and os.path.getsize(p) > 0:
        shutil.copyfile(p, os.path.join(DST, b))
        print("copied", a, "->", b)
    else:
        print("MISSING", a)


def metric(path, name):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    i = hdr.index(name)
    return float(rows[2][i].replace(",", "")), rows[1][i]


traffic = {}
for tag, fname, cmd in (("cfg3", "r2_ggs_cfg3_ncu_raw.csv", "python tools/ggs_stage_probe.py 20 2048"),
                        ("cfg5", "r2_ggs_cfg5_ncu_raw.csv", "python tools/ggs_stage_probe.py 80 4096")):
    path = os.path.join(DST, fname)
    if not os.path.exists(path):
        continue
    rd, ru = metric(path, "dram__bytes_read.sum")
    wr, wu = metric(path, "dram__bytes_write.sum")
    dur, du = metric(path, "gpu__time_duration.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    traffic[f"ggs_entry<false>@{tag}"] = {
        "dram_bytes_read": int(rd * scale[ru]), "dram_bytes_write": int(wr * scale[wu]), "launch_ms_under_ncu": dur * tscale[du],
        "source": f"ncu --set full --clock-control none --import-source on -k regex:ggs_entry -s 2 -c 1 {cmd} (profiles/{fname}; paired layout, "
                  "one-hop exchange, evict-first stream policy; the probe instantiation of the kernel)"}
json.dump(traffic, open(os.path.join(DST, "r2_traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
