"""DRAM traffic of ONE op call from an `ncu --set full --profile-from-start off` report of tools/profile_call.py (run here, no
GPU needed) -> an entry of profiles/ncu_traffic.json, which bench.py's roofline.traffic reads (never a literal in bench.py):

    python tools/ncu_traffic.py gpurun_out/r2_attn_call.ncu-rep "attention[16, 65536, 64, 8]" "<command that produced the report>"

The entry is the SUM of dram__bytes_read.sum + dram__bytes_write.sum over every launch in the report (the op may consist of
several kernels), with per-kernel detail and its provenance."""
import csv, io, json, os, subprocess, sys

UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "Tbyte": 1e12}
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, key, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    ik, ir, iw, it = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    tot, detail = 0.0, []
    for r in launches:
        rd = float(r[ir].replace(",", "")) * UNIT.get(units[ir], 1.0)
        wr = float(r[iw].replace(",", "")) * UNIT.get(units[iw], 1.0)
        tot += rd + wr
        detail.append({"kernel": r[ik][:80], "dram_read": rd, "dram_written": wr, "duration": f"{r[it]} {units[it]}"})
    out = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    d = json.load(open(out)) if os.path.isfile(out) else {}
    d[key] = {"dram_bytes": tot, "launches": len(launches), "source": f"{os.path.basename(path)} (ncu --set full, cold-cache replay; {cmd})",
              "kernels": detail}
    json.dump(d, open(out, "w"), indent=1)
    print(f"{key}: {tot / 1e6:.1f} MB over {len(launches)} launches -> {out}")


if __name__ == "__main__":
    main()
