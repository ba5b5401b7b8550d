"""Repeats tests/test_gpu_parity.py::test_two_label_visits_trajectory N times in one process and prints every failure
(flake hunting: float atomics make the summation order - and so the Adam-amplified round-off - vary run to run).

    python tools/stress_trajectory.py [N]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_gpu_parity as T  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
bad = 0
for i in range(n):
    try:
        T.test_two_label_visits_trajectory()
    except AssertionError as ex:
        bad += 1
        print(f"run {i}: FAIL\n{str(ex)[:1500]}", flush=True)
print(f"{bad} failures in {n} runs (CVG_STREAMS={os.environ.get('CVG_STREAMS', '1')})")
