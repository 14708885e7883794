"""The staging-copy pools of the pageable host path (csrc/copy_pool.h) are plain C++: a multi-threaded stress test on the CPU
(random multi-segment jobs, two pools at once, two callers on one pool, workers falling asleep in between)."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
@pytest.mark.parametrize("threads", ["3", "16"])
def test_copy_pool_stress(tmp_path, threads):
    exe = tmp_path / "copy_pool_stress"
    subprocess.run(["g++", "-std=c++17", "-O2", "-pthread", "-Wall", "-Werror", f"-I{ROOT / 'dxt_lossless_transform_b200' / 'csrc'}",
                    str(ROOT / "tests" / "native" / "copy_pool_stress.cpp"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), "200"], capture_output=True, text=True, timeout=300, env={"DLTCUDA_COPY_THREADS": threads})
    assert out.returncode == 0 and "copy pool ok" in out.stdout, out.stdout + out.stderr
