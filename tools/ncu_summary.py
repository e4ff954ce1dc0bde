"""Summarise an .ncu-rep (read here, no GPU needed) into a small table for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]"""
import csv
import subprocess
import sys

KEYS = [
    ("dur_us", "gpu__time_duration.sum"),
    ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor_rt_pct", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("dram_read_MB", "dram__bytes_read.sum"),
    ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("lsu_wavefront_pct", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
    ("xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("fma_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("issue_pct", "sm__inst_executed.sum.pct_of_peak_sustained_elapsed"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("smem_KB", "launch__shared_mem_per_block_dynamic"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["| # | kernel | grid | " + " | ".join(k for k, _ in KEYS) + " |", "|" + "---|" * (3 + len(KEYS))]
    for n, r in enumerate(rows[2:]):
        vals = []
        for k, m in KEYS:
            if m in idx:
                v = r[idx[m]]
                u = units[idx[m]]
                try:
                    f = float(v.replace(",", ""))
                    if k == "dur_us":
                        f = f * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
                    if k.endswith("_MB"):
                        f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
                    if k == "smem_KB":
                        f = f * {"byte": 1e-3, "Kbyte": 1, "Mbyte": 1e3}.get(u, 1)
                    v = f"{f:.1f}"
                except ValueError:
                    pass
                vals.append(v)
            else:
                vals.append("-")
        name = r[idx["Kernel Name"]].split("(")[0][-40:]
        lines.append(f"| {n} | {name} | {r[idx['Grid Size']]} | " + " | ".join(vals) + " |")
    out = "\n".join(lines)
    print(out)
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            f.write(f"# ncu --set full summary of {rep}\n\n" + out + "\n")


if __name__ == "__main__":
    main()
