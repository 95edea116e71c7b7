"""Every Python entry point and helper script of the repo must at least parse: they run on the GPU
box where a syntax error costs a whole gpurun call (tools/kernel_times.py once shipped broken)."""
import ast
import glob
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_all_python_sources_parse():
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    for sub in ("tools", "cqs_b200", "oracle", "tests", os.path.join("tests", "golden")):
        files += glob.glob(os.path.join(ROOT, sub, "*.py"))
    assert len(files) > 20
    for f in files:
        with open(f) as fh:
            ast.parse(fh.read(), filename=f)


def test_bench_cli_contract_flags():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert f'"{flag}"' in src
    for key in ('"roofline"', '"cpu_baseline"', '"e2e"', '"gpu_launches"', '"clocks"', '"vs_baseline"'):
        assert key in src
