"""bench.py hygiene that the round-end driver checks from the outside: both arms print the same `config.workload` for
the same configuration, the reference arm imports nothing of the product (its process must map oracle/_ref only), the
product arm executes oracle/ only in its cpu_baseline leg, and every BASELINE.json configuration has a workload."""
import ast
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _functions():
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    return {n.name: n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)}


def _imports(fn):
    mods = set()
    for n in ast.walk(fn):
        if isinstance(n, ast.Import):
            mods.update(a.name.split(".")[0] for a in n.names)
        elif isinstance(n, ast.ImportFrom) and n.module:
            mods.add(n.module.split(".")[0])
    return mods


def test_workloads_cover_the_baseline_configs_and_strings_are_shared():
    import bench
    configs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert sorted(bench.WORKLOADS) == list(range(1, len(configs) + 1))
    for cid, w in bench.WORKLOADS.items():
        name = w.get("model", "spatial_conv")
        assert name.lower().replace("_", "") in configs[cid - 1].lower().replace("_", "").replace(" ", ""), (cid, name)
        # one function, no arm argument: the two arms cannot print different strings for the same configuration
        s = bench.workload_string(w, w["B"])
        assert str(w["P"]) in s and s == bench.workload_string(dict(w), w["B"])
    assert bench.net_args(bench.WORKLOADS[2], 32, "ours")[:4] == bench.net_args(bench.WORKLOADS[2], 32, "reference")


def test_reference_arm_never_imports_the_product():
    fns = _functions()
    for name in ("run_reference", "measure_reference"):
        assert "mccnn_b200" not in _imports(fns[name]), name
    # run_config1 serves both arms: the product import sits on the arm == "ours" branches only
    src = ast.get_source_segment(open(os.path.join(ROOT, "bench.py")).read(), fns["run_config1"])
    for line in src.splitlines():
        if "mccnn_b200" in line and "import" in line:
            assert line.startswith("        "), line  # nested under `if arm == "ours":`
    top = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    top_imports = {a.name.split(".")[0] for n in top.body if isinstance(n, ast.Import) for a in n.names} | \
                  {n.module.split(".")[0] for n in top.body if isinstance(n, ast.ImportFrom) and n.module}
    assert "mccnn_b200" not in top_imports and "oracle" not in top_imports and "common" not in top_imports


def test_product_arm_executes_the_oracle_only_as_cpu_baseline():
    fns = _functions()
    for name, fn in fns.items():
        if "oracle" in _imports(fn):
            assert name in ("cpu_baseline_line", "run_config1", "run_reference", "measure_reference"), name
    # run_config1 (both arms): oracle.ref is the reference arm's op module, oracle.numpy_ops the product arm's
    # cpu_baseline leg -- never the C oracle as input preparation
    src = ast.get_source_segment(open(os.path.join(ROOT, "bench.py")).read(), fns["run_config1"])
    assert "oracle import cpu" not in src
