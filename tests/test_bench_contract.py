"""bench.py output contract (CPU side): the reference arm prints ONE JSON line with the keys the driver reads, and the
flop / metric bookkeeping of the GPU arm matches SURVEY.md 8d."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-batch", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "images/sec"
    assert d["metric"].startswith("images/sec ViT-B/16") and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "oracle" in cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_algorithmic_flops_match_the_survey():
    sys.path.insert(0, ROOT)
    import bench
    from oracle import vit_oracle as O
    assert abs(O.flops_per_image(**bench.VIT_B16) / 1e9 - 35.128) < 0.01      # ViT-B/16 (cls)
    assert abs(O.flops_per_image(**bench.MODELS["vit_l16"]) / 1e9 - 123.109) < 0.05
    assert abs(O.flops_per_image(**bench.MODELS["vit_h14"]) / 1e9 - 310.867) < 0.1
