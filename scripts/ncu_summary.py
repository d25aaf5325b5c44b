"""Text summary of `ncu --set full` reports (gpurun_out/full_*.ncu-rep) for profiles/: duration, SM clock, DRAM traffic, pipe
utilisation, issue-slot utilisation and the top warp-stall reasons of each captured kernel."""
import csv
import glob
import io
import os
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "duration"), ("sm__cycles_elapsed.avg.per_second", "SM clock"), ("dram__bytes_read.sum", "DRAM read"),
        ("dram__bytes_write.sum", "DRAM write"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        # tcgen05 work: the hmma-path counters do not see UTCHMMA (they read 0 on a GEMM at 50 % of peak); the cycles in which tensor memory
        # is active track the achieved tensor throughput (wgrad_kv: 46.7 % here vs 52 % of the bf16 peak by FLOPs / time)
        ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor memory (tcgen05) active % of elapsed"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "MMA operand reads from shared memory % of peak"),
        ("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "legacy hmma-path counter (0 for tcgen05)"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"), ("smsp__inst_executed.sum", "warp instructions")]

out = []
for rep in sorted(glob.glob(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/full_*.ncu-rep")):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
    name = d.get("Kernel Name", ("?", ""))[0]
    out.append("== %s   (%s)" % (os.path.basename(rep), name[:90]))
    for k, label in KEYS:
        hit = [h for h in hdr if h.endswith(k)]
        if hit:
            v, u = d[hit[0]]
            out.append("   %-44s %s %s" % (label, v, u))
    st = []
    for h in hdr:
        if "pcsamp_warps_issue_stalled_" in h and "not_issued" not in h:
            try:
                st.append((float(d[h][0]), h.split("issue_stalled_")[1]))
            except ValueError:
                pass
    tot = sum(x for x, _ in st) or 1.0
    out.append("   warp stall samples: " + ", ".join("%s %.0f%%" % (n, 100 * x / tot) for x, n in sorted(st, reverse=True)[:7]))
print("\n".join(out))
