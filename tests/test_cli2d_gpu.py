"""The 2D `nbco` command line (C++ host code over the C ABI; surface of the reference's main.cu, which does not
compile at HEAD): options, file naming, fp64 snapshot format, -test table, and one step against the oracle."""
import os
import subprocess

import numpy as np
import pytest

import coulomb_oscillators_b200 as nb
from refs2d import Oracle2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "coulomb_oscillators_b200", "nbco")


def run(args, cwd):
    return subprocess.run([CLI] + args, cwd=cwd, capture_output=True, text=True, timeout=300)


def test_cli2d_snapshots_format_and_first_step(tmp_path):
    out = tmp_path / "out"
    out.mkdir()
    n = 6000
    r = run(["-n", str(n), "-iters", "2", "-steps", "2", "-integ", "pefrl", "-o", str(out)], tmp_path)
    assert r.returncode == 0, r.stderr
    assert sorted(os.listdir(out)) == ["args.txt", "out0_0.000500.bin", "out2_0.000500.bin"]
    assert os.path.getsize(out / "out0_0.000500.bin") == 32 * n      # all positions then all velocities, fp64 x 2
    assert "perveance" in r.stdout and "dep. phase adv." in r.stdout
    # snapshot 0 = the KV beam of main.cu after ONE PEFRL step of coulombOscillatorFMM at p = 5 (defaults)
    got = np.fromfile(out / "out0_0.000500.bin", np.float64).reshape(2, n, 2)
    st = nb.init_kv2(n)
    par = nb.default_param2(n)
    buf = np.concatenate([st[0], st[1], np.zeros((n, 2))]).copy()
    orc = Oracle2(order=5)
    orc.eval(3, buf, n, par)
    orc.integrate(nb.PEFRL, 3, buf, n, par, 5e-4, 1)
    assert np.abs(got[0] - buf[:n]).max() <= 1e-12 * np.abs(buf[:n]).max()
    assert np.abs(got[1] - buf[n:2 * n]).max() <= 1e-12 * np.abs(buf[n:2 * n]).max()
    # resume from a snapshot
    out2 = tmp_path / "out2"
    out2.mkdir()
    r = run(["-iters", "0", "-steps", "1", "-o", str(out2), str(out / "out2_0.000500.bin")], tmp_path)
    assert r.returncode == 0 and os.path.getsize(out2 / "out0_0.000500.bin") == 32 * n


def test_cli2d_options_and_errors(tmp_path):
    out = tmp_path / "o"
    out.mkdir()
    ok = ["-n", "3000", "-iters", "0", "-steps", "1", "-o", str(out)]
    assert run(ok + ["-ga", "-p", "3", "-r", "2", "-eps", "1e-6", "-i", "2", "-gpu", "256", "-gridsize", "64"], tmp_path).returncode == 0
    assert run(ok + ["-A", "2e-3", "1e-3", "-omega", "30", "28", "-xi", "1e-3", "-omega0", "39", "38", "-ncoll"], tmp_path).returncode == 0
    assert run(ok + ["-x", "1e-3", "5e-4", "-u", "0.03", "0.02", "-integ", "fr"], tmp_path).returncode == 0
    assert "unrecognised option" in run(["-bogus"], tmp_path).stderr
    assert run(["-p", "11"], tmp_path).returncode != 0
    assert run(["-r", "0.5"], tmp_path).returncode != 0
    assert run(["-x", "1"], tmp_path).returncode != 0
    assert "no CPU path" in run(["-cpu"], tmp_path).stderr
    r = run(["-n", "1000", "-iters", "0", "-o", str(tmp_path / "missing")], tmp_path)
    assert r.returncode != 0 and "Create it if not" in r.stderr
    assert run(["-h"], tmp_path).returncode == 0


def test_cli2d_test_mode_errors_fall_with_the_order(tmp_path):
    r = run(["-test", "-n", "20000"], tmp_path)
    assert r.returncode == 0, r.stderr
    errs = [float(l.split(":")[-1]) for l in r.stdout.splitlines() if "Relative error" in l]
    assert len(errs) == 10 and "Time elapsed" in r.stdout
    assert errs[0] > errs[4] > errs[9] and errs[9] < 1e-5
    # the same numbers from the oracle (reference algorithm): FMM against the direct sum, p = 1, 5, 10
    st = nb.init_kv2(20000)
    par = nb.default_param2(20000)
    for p in (1, 5, 10):
        o = Oracle2(order=p).fmm(st[0], None, par)
        d = Oracle2().direct(o["pos"], par)
        e = np.sqrt(((o["acc"] - d) ** 2).sum(1) / ((d ** 2).sum(1) + 1e-18)).mean()
        assert abs(errs[p - 1] - e) <= 1e-6 * e + 1e-15, (p, errs[p - 1], e)
