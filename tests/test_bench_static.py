"""Static checks of bench.py that need no GPU: the multi-rank control flow must issue the same
number of collectives on every rank (a rank-dependent loop around an all-reduce deadlocked an
8-GPU run once), and the JSON contract keys are all present in the printed line."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLLECTIVE = {"all_reduce", "barrier", "step_device", "step_host", "max_over_ranks", "broadcast", "all_gather"}


def _calls(node):
    for n in ast.walk(node):
        if isinstance(n, ast.Call):
            f = n.func
            yield f.attr if isinstance(f, ast.Attribute) else getattr(f, "id", "")


def test_no_collective_inside_data_dependent_loops():
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    bad = []
    for loop in ast.walk(tree):
        if not isinstance(loop, (ast.For, ast.While)):
            continue
        has_break = any(isinstance(n, ast.Break) for n in ast.walk(loop))
        if has_break and COLLECTIVE & set(_calls(loop)):
            bad.append(loop.lineno)
    assert not bad, "loops with an early exit must not contain collectives: bench.py lines %s" % bad


def test_rank_conditionals_contain_no_collectives():
    """`if rank == 0:` (or any test on rank / local) must not guard a collective."""
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    bad = []
    for node in ast.walk(tree):
        if isinstance(node, ast.If):
            names = {n.id for n in ast.walk(node.test) if isinstance(n, ast.Name)}
            if names & {"rank", "local"} and COLLECTIVE & set(_calls(node)):
                bad.append(node.lineno)
    assert not bad, "rank-dependent branches with collectives: bench.py lines %s" % bad


def test_json_contract_keys_present():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks",
                "roofline", "cpu_baseline", "h2d_bytes_per_step", "d2h_bytes_per_step", "impl"):
        assert '"%s"' % key in src, key
