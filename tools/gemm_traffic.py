"""profiles/gemm_traffic.json from an `ncu --set full` capture of bench.py: mean dram bytes (read + write) per launch of
the dominant kernel (gemm2_kernel), stamped with the library version so that bench.py only quotes it for the build it
was measured on.   usage: python tools/gemm_traffic.py gpurun_out/prof.ncu-rep [model]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rep = sys.argv[1]
    model = sys.argv[2] if len(sys.argv) > 2 else "vit_b16"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = []
    for r in rows[2:]:
        if "gemm2_kernel" not in r[idx["Kernel Name"]]:
            continue
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[idx[m]].replace(",", "")) * scale[units[idx[m]]]
        per.append((r[idx["Kernel Name"]].split("(")[0][-28:], tot))
    from vit_pytorch_b200 import _lib
    out = {"model": model, "lib_version": int(_lib.lib().b200vit_version()), "source": os.path.basename(rep),
           "dram_bytes_per_launch": sum(t for _, t in per) / len(per), "launches": [{"kernel": k, "dram_bytes": t} for k, t in per]}
    json.dump(out, open(os.path.join(ROOT, "profiles", "gemm_traffic.json"), "w"), indent=1)
    print(json.dumps(out)[:400])


if __name__ == "__main__":
    main()
