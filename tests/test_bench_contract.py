"""bench.py's CPU arm (`--impl reference`: the C restatement of the reference path on the host cores) runs without a GPU
and prints the contract's JSON line; under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

from helpers import ROOT


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip()


def test_reference_arm_prints_one_contract_line():
    out = _run()
    line = json.loads(out.splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "poseidon_perms_per_sec" and line["unit"] == "perms/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "perms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["steps"] == 1 and line["warmup"] == 0
    # the CPU arm's tree is the first 1/16 of the 2^24-leaf job: its root is node 15 of the committed oracle tree
    assert line["cpu_baseline"]["root_matches_oracle_golden"] is True
    assert line["cpu_baseline"]["host"]["threads"] == line["cpu_baseline"]["cores"]


def test_reference_arm_other_ranks_stay_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == ""
