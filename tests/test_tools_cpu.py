"""The measurement tools that turn an ncu launch list into the committed summaries must keep working on the
committed lists (no GPU): profiles/*_launches_page2800x2000.csv -> per-layer table and DRAM traffic per kernel group,
which bench.py reads for `roofline.traffic`."""
import glob
import json
import os
import subprocess
import sys

from conftest import ROOT


def newest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    assert files, pattern
    return files[-1]


def test_launch_list_tools_parse_the_committed_list(tmp_path):
    csv = newest("*_launches_page2800x2000.csv")
    tag = os.path.basename(csv).split("_")[0]
    log = tmp_path / "prof_page.log"
    # layer names in launch order, as tools/prof_page.py prints them
    names = [l.split()[0] for l in open(csv.replace(".csv", ".txt")).read().splitlines()[1:] if not l.startswith("sum of")]
    log.write_text("LAYERS " + ",".join(names) + "\n")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_launches.py"), csv, str(log)],
                         capture_output=True, text=True, check=True).stdout
    assert out.splitlines()[1].startswith("stem_pad") and "sum of kernel durations" in out
    assert len(names) == 58
    traffic = tmp_path / "traffic.json"
    committed_file = json.load(open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json")))
    env = dict(os.environ, **committed_file.get("plan_env", {}))     # the kernel plan the list was measured with
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py"), csv, str(log), str(traffic)],
                   capture_output=True, text=True, check=True, env=env)
    got = json.load(open(traffic))["groups"]
    committed = committed_file["groups"]
    assert got.keys() == committed.keys()
    for k in got:
        assert got[k]["launches"] == committed[k]["launches"]
        assert abs(got[k]["dram_bytes"] - committed[k]["dram_bytes"]) <= 1e-6 * committed[k]["dram_bytes"] + 1
    assert sum(g["launches"] for g in got.values()) == 58


def test_bench_takes_traffic_only_from_a_list_measured_with_the_current_kernels(tmp_path, monkeypatch):
    """roofline.traffic comes from the newest profiles/*_traffic.json -- but only while the kernel sources are the
    ones that list was measured with (csrc digest stamped by tools/ncu_traffic.py); otherwise null + the reason."""
    sys.path.insert(0, ROOT)
    import bench
    prof = tmp_path / "profiles"
    prof.mkdir()
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    monkeypatch.setattr(bench, "csrc_digest", lambda: "abc")
    old = {"groups": {"conv_gemm_tc<BN=128>": {"launches": 47, "us": 1.0, "dram_bytes": 47e8}}}
    (prof / "r01_traffic.json").write_text(json.dumps(old))                       # round-1 file: no stamp at all
    per_launch, src = bench.ncu_traffic("conv_gemm_tc<BN=128>")
    assert per_launch is None and "stale" in src["refused"]
    (prof / "r02_traffic.json").write_text(json.dumps(dict(old, csrc_digest="abc", git_head="1234567")))
    per_launch, src = bench.ncu_traffic("conv_gemm_tc<BN=128>")
    assert per_launch == 1e8 and src["launches_per_page"] == 47 and src["git_head"] == "1234567"
    assert src["file"] == os.path.join("profiles", "r02_traffic.json")
    monkeypatch.setattr(bench, "csrc_digest", lambda: "changed")
    assert bench.ncu_traffic("conv_gemm_tc<BN=128>")[0] is None
    monkeypatch.undo()
    assert len(bench.csrc_digest()) == 16


def test_bench_arms_share_one_config_block(monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    for k in ("SBB_PAIR", "SBB_PAIR_HEAD", "SBB_PAIR64", "SBB_DEC4_MERGED", "SBB_DEC5_MERGED"):
        monkeypatch.delenv(k, raising=False)
    for name, cfg in bench.CONFIGS.items():
        assert bench.config_block(cfg, 8) == bench.config_block(cfg, 8) and set(bench.config_block(cfg, 1)) == {"workload", "parallelism", "l2"}
    # kernel groups follow the plan: multi-tap N = 128 launches on the CTA-pair kernel unless switched off
    assert bench.kernel_group("dec5") == "conv_gemm_pair<BN=128,head>" and bench.kernel_group("dec4") == "conv_gemm_pair<BN=128>"
    assert bench.kernel_group("res4b_branch2b") == bench.kernel_group("res4b_branch2c") == "conv_gemm_pair<BN=128>"
    assert bench.kernel_group("res2b_branch2c") == "conv_gemm_tc<BN=128>"
    assert bench.kernel_group("conv1") == bench.kernel_group("res2a_branch2b") == "conv_gemm_pair<BN=64>"
    assert bench.kernel_group("res2a_branch2a") == "conv_gemm_tc<BN=64>"
