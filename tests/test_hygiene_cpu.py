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
