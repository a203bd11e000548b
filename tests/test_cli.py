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


@pytest.mark.parametrize("exe,args,count", [("lorastencil_1d", ["1d2r", "1024", "2"], 9),
                                            ("lorastencil_2d", ["box2d3r", "64", "64", "2"], 49),
                                            ("lorastencil_3d", ["box3d1r", "8", "8", "64", "2"], 27),
                                            ("lorastencil_3d", ["star3d2r", "8", "8", "64", "2"], 125)])
def test_weight_file_errors(tmp_path, exe, args, count):
    """--weights FILE is validated before any CUDA call: a missing file, a short table and a non-number each give
    'Invalid argument: ...' on stderr and return 1, like the reference's own argument errors (src/2d/main.cu:128-131)."""
    r = run(exe, *args, "--weights", str(tmp_path / "absent.txt"))
    assert r.returncode == 1 and "Invalid argument: cannot open the weight file" in r.stderr
    short = tmp_path / "short.txt"
    short.write_text(" ".join(["1.5"] * (count - 1)))
    r = run(exe, *args, "--weights", str(short))
    assert r.returncode == 1 and f"must hold exactly {count} numbers (found {count - 1})" in r.stderr
    long = tmp_path / "long.txt"
    long.write_text("\n".join(["2"] * (count + 1)))
    r = run(exe, *args, "--weights", str(long))
    assert r.returncode == 1 and f"(found {count + 1})" in r.stderr
    junk = tmp_path / "junk.txt"
    junk.write_text(" ".join(["1"] * 3 + ["x"] + ["1"] * (count - 4)))
    r = run(exe, *args, "--weights", str(junk))
    assert r.returncode == 1 and "not a number" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("exe,args,count", [("lorastencil_1d", ["1d1r", "5000", "3"], 9),
                                            ("lorastencil_2d", ["star2d1r", "96", "130", "3"], 49),
                                            ("lorastencil_2d", ["box2d3r", "64", "64", "2"], 49),
                                            ("lorastencil_3d", ["star3d1r", "9", "16", "64", "3"], 27),
                                            ("lorastencil_3d", ["box3d2r", "9", "16", "64", "3"], 125)])
def test_weight_file_is_honoured(tmp_path, exe, args, count):
    """A caller's table (dense, asymmetric, no zero anywhere) from a file: the --check protocol of the reference
    (one direct-tap CPU step with EVERY weight against one launch, src/2d/main.cu:282-328) finds no mismatch."""
    import numpy as np
    w = np.random.default_rng(count).uniform(-1.0, 1.0, count)
    f = tmp_path / "w.txt"
    f.write_text("\n".join(" ".join(f"{v:.17g}" for v in w[i:i + 7]) for i in range(0, count, 7)))
    r = run(exe, *args, "--weights", str(f), "--check")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert f"INFO: weights = {f} ({count} values, every one honoured)" in r.stdout
    assert "Correct!" in r.stdout and "naive = " not in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("shape,info", [("box3d2r", "box_3d2r"), ("star3d2r", "star_3d2r")])
def test_radius2_shapes_pass_the_check_protocol(shape, info):
    """The radius-2 extensions through the same driver: INFO line, banner, one direct-tap CPU step (5x5x5 window)
    against one launch."""
    r = run("lorastencil_3d", shape, "12", "10", "130", "4", "--check")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert f"INFO: shape = {info}, h = 12, m = 10, n = 130, times = 4" in r.stdout
    assert f"LoRAStencil(3D {info}): " in r.stdout and "GStencil/s = " in r.stdout
    assert "Correct!" in r.stdout and "naive = " not in r.stdout
