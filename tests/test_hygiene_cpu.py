"""CPU: structural rules of the build — the product package never imports the oracle (test infrastructure only), nothing
that runs on the GPU box reads /root/reference at run time, and the package fails loudly without its CUDA library."""
import re
import subprocess
import sys

from conftest import ROOT


def _py(dirname):
    return sorted((ROOT / dirname).glob("*.py"))


def test_product_package_never_touches_the_oracle():
    for f in _py("simxns_b200"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f.name} imports the oracle"
        assert "oracle." not in re.sub(r'""".*?"""', "", src, flags=re.S), f"{f.name} references the oracle"


def test_nothing_on_the_gpu_box_reads_the_reference_tree():
    """citations in docstrings / comments are fine; opening or importing from /root/reference is not"""
    files = _py("simxns_b200") + _py("tests") + [ROOT / "bench.py", ROOT / "__graft_entry__.py"]
    for f in files:
        code = re.sub(r'""".*?"""', "", f.read_text(), flags=re.S)
        code = "\n".join(line.split("#", 1)[0] for line in code.splitlines())
        if f.name == "test_hygiene_cpu.py":
            continue  # this file spells the path in order to look for it
        assert "/root/" + "reference" not in code, f"{f.name} touches the reference tree at run time"


def test_missing_library_fails_loudly():
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from simxns_b200 import _lib\n"
            "from pathlib import Path\n"
            "_lib.LIB_PATH = Path('/nonexistent/libsimx_b200.so')\n"
            "try:\n"
            "    _lib.lib()\n"
            "except RuntimeError as e:\n"
            "    assert 'no fallback' in str(e).lower(), e\n"
            "    print('raised')\n") % str(ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.stdout.strip() == "raised", out.stderr[-500:]


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to the GPU arm): ONE JSON line with the metric / unit of
    BASELINE.json, `impl`, a `cpu_baseline` describing the run and an `e2e` that moves no bytes"""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--layers", "1", "--ref-queries", "1"], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("query+passage pairs/sec") and base["metric"].startswith("query+passage pairs/sec")
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert abs(d["cpu_baseline"]["value"] - d["value"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]
