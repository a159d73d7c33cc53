"""CPU: the parts of bench.py's contract that do not need a GPU — the workload description shared
by both arms, the algorithmic bytes per row behind `roofline.achieved`, which ops are new surface
(`pinned: false`), the traffic table behind `roofline.traffic`, and the reference arm's JSON line
(run here on a small column: it is the one place besides tests/ where bench.py executes oracle/)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402

pytestmark = pytest.mark.timeout(300)


def test_config2_is_55_ops_over_the_four_sub_word_types():
    ops = bench.cfg2_ops()
    assert len(ops) == 55 and len({s[0] for s in ops}) == 55
    by_type = {}
    for s in ops:
        by_type.setdefault(s[0].split(".")[0], []).append(s[0])
    assert {k: len(v) for k, v in by_type.items()} == {"i8": 11, "u8": 11, "i16": 11, "u16": 11, "cast": 11}
    # BASELINE.json configs[1] / SURVEY §8d: bytes per row = sizeof(inputs) + sizeof(output)
    bpr = {s[0]: bench.bytes_per_row(s) for s in ops}
    assert bpr["i8.add"] == 3 and bpr["u16.add"] == 6 and bpr["i8.shl"] == 6 and bpr["u16.shr"] == 8
    assert bpr["cast.i8->f32"] == 5 and bpr["cast.u16->u32"] == 6 and bpr["cast.f32->u8"] == 5 and bpr["u8.not"] == 2


def test_new_surface_ops_are_the_ones_the_reference_lacks():
    """SURVEY §8a "new surface": sub-word array arithmetic and scalar arithmetic other than u16 + scalar"""
    unpinned = sorted(s[0] for s in bench.cfg2_ops() if not bench.pinned_by_reference(s))
    assert len(unpinned) == 19
    assert "u16.add_scalar" not in unpinned                      # arithmetic/compute_shaders/u16/scalar.wgsl:15-23
    assert all(name.split(".")[1] in ("add", "sub", "mul", "add_scalar", "mul_scalar") for name in unpinned)
    assert not any(name.startswith("cast.") or name.endswith((".and", ".or", ".xor", ".not", ".shl", ".shr")) for name in unpinned)


def test_both_arms_describe_the_same_workload():
    a, b = bench.config_dict(1 << 28), bench.config_dict(1 << 28)
    assert a == b and a["rows_per_gpu"] == 1 << 28 and a["ops_per_step"] == 55
    assert "configs[1]" in a["workload"] and "model" not in a


def test_every_op_has_a_measured_dram_traffic_entry():
    table = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    rows = 1 << 28
    for s in bench.cfg2_ops():
        got = bench.known_traffic(s[0])
        assert got == table[s[0]]
        # ncu's dram bytes for one launch: within 25 % of the algorithmic bytes (part of the output is
        # still dirty in L2 when the kernel ends; nothing is read twice)
        assert 0.75 <= got / (bench.bytes_per_row(s) * rows) <= 1.05, s[0]


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "65536", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=280)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["unit"] == "rows/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["config"] == bench.config_dict(65536)
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(line["value"] - 55 * 65536 / (line["ms_per_step"] * 1e-3)) / line["value"] < 0.01


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """the driver launches the reference arm like ours: under torchrun at N > 1 only rank 0 works and prints"""
    env = dict(os.environ, OMP_NUM_THREADS="2")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29593", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--rows", "65536", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, cwd=ROOT,
                         timeout=280)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["config"] == bench.config_dict(65536)
