"""Condenses an `ncu --set full` report into the few counters the roofline discussion uses (run here, no GPU needed):

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/rN_x_ncu.txt

One block per profiled launch: duration, issue-slot and pipe utilisation, DRAM bytes, occupancy, registers, and the warp
stall reasons per issued instruction (largest first).  Also prints a JSON line per launch that tools/ncu_traffic.py-style
consumers (bench.py's roofline.traffic) can keep."""
import csv, io, json, subprocess, sys

KEEP = [
    ("gpu__time_duration.sum", "duration"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles %"),
    ("sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active", "FMA pipe, fp16 instructions %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe cycles %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe cycles %"),
    ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tcgen05 (tc) pipe cycles %"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform datapath %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts (LSU) %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts (tensor core) %"),
    ("lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "L2 sectors %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of max"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def main():
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, launches = rows[0], rows[1], rows[2:]
        ik = hdr.index("Kernel Name")
        for r in launches:
            print(f"== {path}: {r[ik][:140]}")
            dram = 0.0
            for key, label in KEEP:
                if key in hdr:
                    i = hdr.index(key)
                    print(f"   {label:42s} {r[i]} {units[i]}")
                    if key.startswith("dram__bytes_"):
                        dram += float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
            stalls = []
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                    stalls.append((float(r[i].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            print("   warp stalls per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:7]))
            print("   " + json.dumps({"kernel": r[ik][:100], "dram_bytes": dram, "source": path.split("/")[-1]}))


if __name__ == "__main__":
    main()
