"""The C++ host-side mirror (include/lz4_flex.hpp) compiles against the C ABI; without a GPU it fails loudly,
with a GPU its block/frame calls round-trip and FrameEncoder reproduces the one-shot frame byte for byte."""
import os
import subprocess

import pytest

from lz4_flex_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_mirror")


def build():
    so = _native.build()
    pkg = os.path.dirname(so)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_mirror.cpp"), "-o", EXE, "-L", pkg,
                           "-l:liblz4b200.so", f"-Wl,-rpath,{pkg}"])
    return EXE


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_cpp_mirror_compiles_and_fails_loudly_without_gpu():
    exe = build()
    r = subprocess.run([exe, "nogpu"], capture_output=True, text=True)
    assert r.returncode == 0 and "nogpu ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_mirror_on_gpu():
    exe = build()
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0 and "gpu ok" in r.stdout, r.stdout + r.stderr
