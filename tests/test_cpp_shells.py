"""The C++14 header shells (superviseddescent_b200/include/) compiled against libsd_b200.so.

CPU: the translation unit that uses them like the reference's tests/apps must compile, and the binary must
fail loudly without a GPU.  GPU: it must reproduce the reference's gtest literals, train on the device route,
and detect() on a reference example frame must match the committed reference-HOG golden.
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shells_binary(tmp_path_factory):
    from superviseddescent_b200 import build
    lib = build.build()
    out = str(tmp_path_factory.mktemp("cpp") / "test_shells")
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-Werror=return-type", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "superviseddescent_b200", "include"), os.path.join(ROOT, "tests", "cpp", "test_shells.cpp"),
           "-L", os.path.dirname(lib), "-lsd_b200", f"-Wl,-rpath,{os.path.dirname(lib)}", "-lpthread", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return out


def test_shells_compile_as_cxx14(shells_binary):
    assert os.path.exists(shells_binary)


def test_shells_fail_loudly_without_gpu(shells_binary):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([shells_binary], capture_output=True, text=True)
    assert r.returncode != 0 and "no usable CUDA device" in r.stdout


@pytest.mark.gpu
def test_shells_reference_literals_train_and_detect(shells_binary, golden, tmp_path):
    gray = golden.examples["gray1"]
    box = golden.examples["boxes"][1]
    raw = tmp_path / "frame.raw"
    raw.write_bytes(np.ascontiguousarray(gray).tobytes())
    out_model = tmp_path / "saved.bin"
    r = subprocess.run([shells_binary, golden.model_path, str(raw), str(gray.shape[1]), str(gray.shape[0]),
                        str(box[0]), str(box[1]), str(box[2]), str(box[3]), str(out_model)], capture_output=True, text=True, timeout=300)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "ALL OK" in r.stdout
    line = [l for l in r.stdout.splitlines() if l.startswith("LANDMARKS")][0].split()[1:]
    names = line[0::3]
    xs = np.array(line[1::3], dtype=np.float32)
    ys = np.array(line[2::3], dtype=np.float32)
    ref = golden.detect["landmarks1"]
    assert names[:3] == ["9", "31", "32"]
    assert np.max(np.abs(np.concatenate([xs, ys]) - ref)) <= 1e-4 * np.max(np.abs(ref))
    assert out_model.read_bytes() == open(golden.model_path, "rb").read()
    assert "The given model file could not be opened" in r.stdout        # model.hpp:199 message
    assert "At * A (ms)" in r.stdout and "Decomposition (ms)" in r.stdout  # verbose_solver.hpp:66-103 lines
    assert "(The rank is 3, full rank would be 4). Increase lambda." in r.stdout   # regressors.hpp:290-293 through ColPivHouseholderQRSolver
