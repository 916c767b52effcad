"""The C++17 host mirror (rupphash_b200/host/rupphash.hpp) compiles against the C ABI everywhere
and, on a GPU box, runs the reference-style checks of harness.cpp."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "rupphash_b200", "host")
EXE = os.path.join(HOST, "build", "harness")


def _build():
    import __graft_entry__ as g
    from rupphash_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        g.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    so_dir = os.path.join(ROOT, "rupphash_b200")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", os.path.join(HOST, "harness.cpp"), "-o", EXE,
                    "-L" + so_dir, "-lrupphash_b200", "-Wl,-rpath," + so_dir, "-pthread"], check=True, capture_output=True)


def test_host_mirror_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_host_mirror_harness_runs():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, f"harness exit {r.returncode}: {r.stdout}{r.stderr}"
    assert "harness ok" in r.stdout
