"""CPU tier: the reference arm of bench.py (the reference's CPU path on the host cores) runs
without a GPU and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                                   "--steps", "1", "--warmup", "0", "--ref-sets-per-core", "1"], text=True,
                                  env=env, timeout=600)
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "sims/s" and line["higher_is_better"] is True
    assert line["metric"] == "TRPL forward sims/sec (nx=128, FP64)"
    assert line["value"] > 0 and line["steps"] == 1
    # oracle/_ref (the unmodified reference modules, oracle/make_ref.py) when present, else the port
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "trial_move_evaluation.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and "hmax" in line["config"]
    assert line["e2e"] == {"value": line["value"], "unit": "sims/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("configs[1]")


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                                   "--steps", "1", "--warmup", "0"], text=True, env=env, timeout=120)
    assert out.strip() == ""
