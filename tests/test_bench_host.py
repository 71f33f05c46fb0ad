"""Host-side pieces of bench.py that run without a GPU: workload arithmetic, CPU thread selection, the N > 1 watchdog."""
import json
import os
import subprocess
import sys
import textwrap

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import bench  # noqa: E402


def test_code_bytes_match_survey_8d():
    # one matvec moves out * (in / g) * K * ceil(nbits / 8) bytes of codes; Llama-3-8B 1x16 = 1.625 GiB per token
    assert bench.code_bytes(4096, 14336, 1, 16) == 14336 * 512 * 2
    assert bench.code_bytes(4096, 11008, 2, 8) == 11008 * 512 * 2
    assert bench.code_bytes(4096, 11008, 8, 8) == 11008 * 512 * 8
    assert bench.model_code_bytes("llama3-8b", 1, 16, 32) == 1744830464
    per_layer = sum(bench.code_bytes(fin, fout, 1, 16) for _, fin, fout in bench.layer_linears("llama3-70b"))
    assert bench.model_code_bytes("llama3-70b", 1, 16, 80) == 80 * per_layer


def test_host_threads_ignores_torchrun_omp_default(monkeypatch):
    monkeypatch.setenv("OMP_NUM_THREADS", "1")  # what torchrun exports to its workers
    n = bench.host_threads()
    assert 1 <= n <= (os.cpu_count() or 1)
    monkeypatch.setenv("AQLM_BENCH_CPU_THREADS", "3")
    assert bench.host_threads() == 3


def test_watchdog_ends_a_stuck_phase_and_respects_cancel():
    code = textwrap.dedent("""
        import sys, time
        sys.path.insert(0, %r)
        import bench
        w = bench.Watchdog(0, 0.4); w.phase("a"); w.phase("b"); w.cancel(); time.sleep(0.7); print("survived", flush=True)
        off = bench.Watchdog(0, 0.1, enabled=False); off.phase("single GPU"); time.sleep(0.3); print("disabled ok", flush=True)
        w = bench.Watchdog(0, 0.3); w.phase("stuck exchange"); time.sleep(5); print("NOT REACHED", flush=True)
    """ % REPO)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
    assert r.returncode == 5, (r.returncode, r.stderr[-400:])
    lines = r.stdout.strip().splitlines()
    assert lines[:2] == ["survived", "disabled ok"] and "NOT REACHED" not in r.stdout
    err = json.loads(lines[-1])
    assert "stuck exchange" in err["error"] and err["rank"] == 0
