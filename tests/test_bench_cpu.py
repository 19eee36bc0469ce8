"""CPU: bench.py's reference arm (`--impl reference`) -- the reference's own renderer on the host cores -- prints the contract's
JSON line, renders the GPU arm's configuration (16 strips), and never maps the product library."""
import json
import subprocess
import sys

import pytest

from conftest import ROOT


def test_reference_arm_line_and_no_product_library(ref):
    code = r'''
import sys, runpy
sys.argv = ["bench.py", "--impl", "reference", "--workload", "c1_dragon_720p", "--steps", "2", "--warmup", "1"]
try:
    runpy.run_path("bench.py", run_name="__main__")
finally:
    maps = open("/proc/self/maps").read()
    libs = sorted({l.split()[-1].rsplit("/", 1)[-1] for l in maps.splitlines() if ".so" in l and "%s" in l})
    print("REPO_LIBS " + " ".join(libs), file=sys.stderr)
''' % str(ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(ROOT), timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mrays/s" and line["value"] > 0
    assert line["config"]["workload"] == "c1_dragon_720p" and line["config"]["strips"] == 16
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0
    libs = [ln for ln in out.stderr.splitlines() if ln.startswith("REPO_LIBS")][-1].split()[1:]
    assert libs == ["libsvo_ref.so"], libs          # the oracle's reference object code and nothing of the product


def test_own_arm_refuses_without_a_gpu(pysvo):
    if pysvo.device_count() > 0:
        pytest.skip("a GPU is visible")
    out = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         cwd=str(ROOT), timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_committed_ncu_capture_matches_the_built_kernel(pysvo):
    """profiles/traffic.json is quoted by bench.py (roofline.traffic and the ncu counters) only while the fine-pass kernel's
    SASS in the library being run hashes to the value stamped into the entry when the capture was filed. This keeps the
    committed entry and the built library in step: a change of the kernel's machine code fails here until the capture is
    retaken (tools/gpu_session.sh ncu) or the entry removed."""
    import shutil
    from tools.ncu_summary import kernel_sass_hash
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    entry = json.loads((ROOT / "profiles" / "traffic.json").read_text())["c3_ico8192_4k"]
    built = kernel_sass_hash(pysvo.LIB_PATH)
    assert built is not None and len(built) == 16
    assert entry["kernel_sass_hash"] == built
    assert entry["dram_bytes_per_launch"] > 1e8 and entry["warp_instructions"] > 1e8     # bytes and counts, not scaled units
    assert kernel_sass_hash(pysvo.LIB_PATH, pattern="noSuchKernel") is None
