"""The drop-in itself: oracle/_ref/vicNl is the reference's own executable (its main, its readers, its
runModel), oracle/_ref/vicNl_gpu the same objects with runModel() replaced by vic_b200/host/vicNl_gpu.cpp + libvicgpu.so.  Both are
run on the same global parameter file (the reference's own input formats, written by vic_b200/synth.py); the output stream
(OutputData::aggdata of every cell and variable at every output step, as raw doubles) and the state file written by the reference's
write_model_state() (write_model_state.c:107-371) -- for the drop-in: from structs refilled through vicgpu_get_state +
vicgpu_unpack_hrurec -- must be identical byte for byte."""
import dataclasses
import os
import subprocess

import numpy as np
import pytest

from vic_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _run(exe, cfgname, tmp_path, tag, ndays, stateday, seed, extra=()):
    p = os.path.join(REF, exe)
    if not os.path.exists(p):
        pytest.skip(f"{p} not built (oracle/Makefile)")
    res = tmp_path / f"res_{tag}"
    res.mkdir()
    cfg = dataclasses.replace(synth.CONFIGS[cfgname], ndays=ndays, out_step=24,
                              extra_global=["STATENAME state", "STATEYEAR 2001", "STATEMONTH 1", f"STATEDAY {stateday}", "STATE_FORMAT ASCII", *extra])
    r = synth.generate(str(tmp_path / f"in_{tag}"), cfg, 2, 3, seed, result_dir=str(res))
    # get_global_param.c:1142 builds the state file name with an overlapping sprintf: the file lands in the working directory as
    # "_<yyyy>-<mm>-<dd>" whatever STATENAME says
    o = subprocess.run([p, "-g", r["global_file"]], cwd=str(res), capture_output=True, text=True)
    assert o.returncode == 0, o.stderr[-2000:]
    out = np.fromfile(res / "results.nc.f64", dtype=np.float64)
    state = (res / f"_2001-01-{stateday:02d}").read_bytes()
    return out, state, o.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("cfgname,ndays,stateday,seed", [("fe_hourly", 5, 3, 501), ("glacier", 4, 4, 502), ("frozen_bands", 2, 2, 503), ("wb_daily", 40, 31, 504)])
def test_vicNl_gpu_matches_stock_vicNl(cfgname, ndays, stateday, seed, tmp_path):
    out_cpu, state_cpu, _ = _run("vicNl", cfgname, tmp_path, "cpu", ndays, stateday, seed)
    out_gpu, state_gpu, err = _run("vicNl_gpu", cfgname, tmp_path, "gpu", ndays, stateday, seed)
    assert "Model execution time (GPU)" in err
    assert out_cpu.size == out_gpu.size and out_cpu.size > 0
    nsteps = ndays if cfgname != "wb_daily" else ndays  # one output step per day
    per_step = out_cpu.size // nsteps
    assert per_step * nsteps == out_cpu.size
    # the reference never initialises aggdata before its first output step (output_list_utils.c:20-24): compare from the second on
    a, b = out_cpu.reshape(nsteps, per_step)[1:], out_gpu.reshape(nsteps, per_step)[1:]
    assert np.array_equal(a, b, equal_nan=True), np.argwhere(~((a == b) | ((a != a) & (b != b))))[:5]
    assert len(state_cpu) > 1000 and state_cpu == state_gpu


def test_stock_vicNl_runs_and_writes_state(tmp_path):
    """CPU-only part: the reference's own executable builds from its sources, runs the synthetic case and writes output + state"""
    out, state, err = _run("vicNl", "fe_hourly", tmp_path, "cpu", 2, 2, 505)
    assert "VIC model run done" in err
    assert out.size > 0 and len(state) > 1000
