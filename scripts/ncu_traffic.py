"""Writes profiles/r02_traffic.json: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernel families that
bench.py reports, read from the `ncu --set full` reports in gpurun_out/ (scripts/gpu_ncu_full.sh).  bench.py copies the number into
`roofline.traffic` with this file as `traffic_src`, so the figure always names the capture (and commit) it came from."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# bench.py kernel family -> report (cfg2: kernels launched alone at B=64, T=512, hidden 300; cfg3: hidden 768, B=8)
MAP = {"cfg2": {"xattn_fwd": "full_attn_fwd", "xattn_bwd": "full_attn_bwd", "gemm_fc1": "full_gemm_fc1", "gemm_fc2": "full_gemm_fc2",
                "gemm_qproj": "full_gemm_q", "wgrad_fc1": "full_wgrad_fc1", "wgrad_qproj": "full_wgrad_q", "layernorm_fwd": "full_ln_fwd",
                "layernorm_bwd": "full_ln_bwd"},
       "cfg3": {"xattn_fwd": "full_attn128_fwd_b8", "xattn_bwd": "full_attn128_bwd_b8"}}


def read(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        return None
    hdr, units, vals = rows[0], rows[1], rows[2]

    def get(key):
        for h, u, v in zip(hdr, units, vals):
            if h.endswith(key):
                mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
                return float(v) * mul
        return 0.0
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    return dict(dram_bytes=get("dram__bytes_read.sum") + get("dram__bytes_write.sum"), dram_read=get("dram__bytes_read.sum"),
                dram_write=get("dram__bytes_write.sum"), kernel=name[:80])


def main():
    rev = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out = {}
    for cfg, fam in MAP.items():
        for k, rep in fam.items():
            path = os.path.join(ROOT, "gpurun_out", rep + ".ncu-rep")
            if not os.path.exists(path):
                continue
            d = read(path)
            if d:
                d["src"] = "ncu --set full, %s.ncu-rep (scripts/gpu_ncu_full.sh), summarised at commit %s" % (rep, rev)
                out.setdefault(cfg, {})[k] = d
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
    for cfg, fam in out.items():
        for k, d in fam.items():
            print("%s %-14s %8.1f MB  (%s)" % (cfg, k, d["dram_bytes"] / 1e6, d["kernel"][:50]))


if __name__ == "__main__":
    main()
