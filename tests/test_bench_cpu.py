"""CPU suite: the reference arm of bench.py (the compiled reference, or the C restatement, timed on the host cores) prints ONE JSON
line on stdout and nothing else, with the keys the bench contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout[:2000]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_committed_ncu_numbers_belong_to_the_committed_kernels():
    """profiles/traffic.json records the hashes of the sources its ncu numbers were measured on; bench.py reports roofline.traffic /
    issue_roofline from it only while they match.  The committed tree must be in that state for the three hot kernels."""
    sys.path.insert(0, ROOT)
    import bench
    for kernel in ("k_env_rollout", "k_nn_conv_tc3", "k_mcts_sim"):
        d, why = bench.ncu_capture(kernel)
        assert d is not None, why
        assert os.path.exists(os.path.join(ROOT, d["capture"])), d["capture"]
