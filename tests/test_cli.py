"""CPU tests of the CLI surface that is decided before any CUDA call: usage text, return codes and
error messages of lorastencil_{1d,2d,3d} (src/1d/main.cu:43-75, src/2d/main.cu:97-135,
src/3d/main.cu:71-106)."""
import os
import subprocess

import pytest

import lorastencil_b200 as ls

BIN = os.path.join(os.path.dirname(ls.lib_path()), "..", "bin")


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not ls.library_built() or not os.path.exists(os.path.join(BIN, "lorastencil_2d")):
        ls.build()


def run(exe, *args):
    return subprocess.run([os.path.join(BIN, exe), *args], capture_output=True, text=True)


@pytest.mark.parametrize("exe,good,nargs", [("lorastencil_1d", "1d2r", 2), ("lorastencil_2d", "box2d1r", 3),
                                            ("lorastencil_3d", "star3d1r", 4)])
def test_usage_and_errors(exe, good, nargs):
    r = run(exe)  # too few arguments -> help, return 1
    assert r.returncode == 1 and f"Program name: {exe}" in r.stdout and "Usage:" in r.stdout
    r = run(exe, good, *(["64"] * (nargs - 1)))  # one argument short
    assert r.returncode == 1 and "Usage:" in r.stdout
    r = run(exe, "nonsense", *(["64"] * nargs))  # unknown shape -> help, return 1
    assert r.returncode == 1 and "Shape:" in r.stdout
    r = run(exe, good, *(["abc"] * nargs))
    assert r.returncode == 1 and "Invalid argument: cannot convert the parameter(s) to integer." in r.stderr
    r = run(exe, good, *(["99999999999999"] * nargs))
    assert r.returncode == 1 and "Argument out of range: the parameter(s) is(are) too large." in r.stderr


def test_shape_names_of_each_driver():
    assert "1d1r or 1d2r" in run("lorastencil_1d").stdout
    assert "box2d1r or star2d1r or box2d3r or star2d3r" in run("lorastencil_2d").stdout
    assert "box3d1r or star3d1r" in run("lorastencil_3d").stdout
