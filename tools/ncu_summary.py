"""Summarise an `ncu --set full` report (.ncu-rep) into the small JSON / text files committed under profiles/.

    python tools/ncu_summary.py gpurun_out/r02_main.ncu-rep profiles/r02_ncu_main_kernels

writes <prefix>.json (what bench.py reads for `roofline.traffic` and the ncu tensor-pipe figures) and <prefix>.txt."""
import csv
import json
import subprocess
import sys

METRICS = [
    ("time_ms", "gpu__time_duration.sum"), ("sm_clock_ghz", "sm__cycles_elapsed.avg.per_second"),
    ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("xu_pipe_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("dram_read_bytes", "dram__bytes_read.sum"), ("dram_write_bytes", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("registers", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
    ("cluster_x", "launch__cluster_dim_x"),
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3, "Ghz": 1.0, "Mhz": 1e-3}


def main(rep, prefix):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        rec = {"kernel": name.split("(")[0].replace("void ", "").strip(), "id": int(r[hdr.index("ID")])}
        for key, metric in METRICS:
            if metric in hdr:
                i = hdr.index(metric)
                try:
                    rec[key] = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
                except ValueError:
                    rec[key] = None
        rec["dram_bytes"] = (rec.get("dram_read_bytes") or 0) + (rec.get("dram_write_bytes") or 0)
        out.append(rec)
    json.dump({"source": rep, "how": "ncu --set full --clock-control none (one replayed launch each; cold caches)", "kernels": out},
              open(prefix + ".json", "w"), indent=1)
    with open(prefix + ".txt", "w") as fh:
        fh.write(f"# {rep}: ncu --set full --clock-control none; per launch\n")
        fh.write(f"{'kernel':56s} {'ms':>8s} {'GHz':>5s} {'tensor%':>8s} {'xu%':>6s} {'dramGB':>7s} {'dram%':>6s} {'L2%':>6s} {'regs':>5s} {'grid':>6s} {'blk':>4s} {'cl':>3s}\n")
        for k in out:
            fh.write(f"{k['kernel'][-56:]:56s} {k['time_ms']:8.3f} {k['sm_clock_ghz']:5.2f} {k['tensor_pipe_pct']:8.1f} {k.get('xu_pipe_pct') or 0:6.1f} "
                     f"{k['dram_bytes'] / 1e9:7.3f} {k['dram_pct']:6.1f} {k['l2_pct']:6.1f} {int(k['registers']):5d} {int(k['grid']):6d} {int(k['block']):4d} {int(k.get('cluster_x') or 0):3d}\n")
    print(open(prefix + ".txt").read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
