"""The C ABI driven from plain C (tests/c/game_loop.c): the reference's renderer call order without any Python in
the loop.  On a CPU-only box the program must stop with UVT_ERR_NO_DEVICE (exit 3), never fall back."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, camera_k0


def _build(tmp_path):
    exe = str(tmp_path / "game_loop")
    pkg = os.path.join(ROOT, "unnamed-voxel-tracer_b200")
    subprocess.run(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-O1", "-Wall", "-Werror", "-std=c11",
                    "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "game_loop.c"),
                    "-o", exe, "-L", pkg, "-luvt", f"-Wl,-rpath,{pkg}"], check=True)
    models = np.load(os.path.join(GOLDEN, "atlas_models.npy"))
    mpath = str(tmp_path / "models.bin")
    models.astype("<u4").tofile(mpath)
    return exe, mpath


def test_c_program_links_and_refuses_to_run_without_a_gpu(uvt, tmp_path):
    import torch
    uvt._native.load()  # make sure libuvt.so is built
    exe, mpath = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe, mpath, "512", "64", "36", str(tmp_path / "f.bin")], capture_output=True, text=True)
    assert r.returncode == 3, r.stderr
    assert "no CPU path" in r.stderr


@pytest.mark.gpu
def test_c_program_frame_matches_oracle(uvt, oracle, world512, tmp_path):
    uvt._native.load()
    exe, mpath = _build(tmp_path)
    out = str(tmp_path / "frame.bin")
    r = subprocess.run([exe, mpath, "512", "320", "180", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    frame = np.fromfile(out, dtype="<u4").reshape(180, 320)
    ref = oracle.render(world512.oracle_world, camera_k0(oracle), 320, 180)
    d = np.abs(frame.view(np.uint8).astype(np.int16) - ref["frame"].view(np.uint8).astype(np.int16)).max()
    assert d <= 1
    assert "kernel launches" in r.stdout
