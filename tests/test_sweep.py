"""ParameterSweep as a replica scheduler: combination order and result structure of the
reference (optimization/sweep.py:85-195), one worker process per device, dynamic assignment, errors per combination."""
import json

import numpy as np
import pytest

from prismo_b200.sweep import ParameterSweep, SweepParameter
from tests import sweep_funcs


def _params():
    return [SweepParameter("w", [1.0, 2.0, 3.0], "um"), SweepParameter("h", [10, 20])]


def test_combinations_follow_meshgrid_ij_order(tmp_path):
    sw = ParameterSweep(_params(), sweep_funcs.toy, output_dir=tmp_path)
    assert [(c["w"], c["h"]) for c in sw.parameter_combinations] == [(1.0, 10), (1.0, 20), (2.0, 10), (2.0, 20), (3.0, 10), (3.0, 20)]
    assert repr(sw) == "ParameterSweep(2 parameters, 6 combinations)"


@pytest.mark.reference
def test_combinations_equal_the_reference(tmp_path, ref, monkeypatch):
    monkeypatch.setenv("PRISMO_B200_DEVICE", "0")
    from prismo.optimization.sweep import ParameterSweep as RP, SweepParameter as RS

    r = RP([RS("w", [1.0, 2.0, 3.0]), RS("h", [10, 20])], sweep_funcs.toy, output_dir=tmp_path / "r")
    m = ParameterSweep(_params(), sweep_funcs.toy, output_dir=tmp_path / "m")
    assert r.parameter_combinations == m.parameter_combinations
    r.run(show_progress=False)
    m.run(show_progress=False)
    strip = lambda rs: [{k: v for k, v in x.items() if k != "pid"} for x in rs]          # noqa: E731
    monkey = [dict(x, env_device=0) if "env_device" in x else x for x in strip(m.results)]
    assert strip(r.results) == monkey
    assert np.array_equal(r.get_result_array("metric"), m.get_result_array("metric"), equal_nan=True)
    assert r.find_optimal("metric")[0] == m.find_optimal("metric")[0]


def test_sequential_results_and_helpers(tmp_path, monkeypatch):
    monkeypatch.setenv("PRISMO_B200_DEVICE", "0")
    sw = ParameterSweep(_params(), sweep_funcs.toy, output_dir=tmp_path)
    res = sw.run(show_progress=False)
    assert [r["status"] for r in res] == ["success"] * 5 + ["error"]
    assert res[-1]["error"] == "bad corner" and res[-1]["parameters"] == {"w": 3.0, "h": 20}
    arr = sw.get_result_array("metric")
    assert arr.shape == (3, 2) and arr[1, 1] == 40.0 and np.isnan(arr[2, 1])
    best, best_res = sw.find_optimal("metric")
    assert best == {"w": 2.0, "h": 20} and best_res["metric"] == 40.0         # ties: first in combination order
    assert sw.find_optimal("metric", maximize=False)[0] == {"w": 1.0, "h": 10}
    saved = json.load(open(sw.save_results()))
    assert len(saved) == 6 and saved[0]["parameters"] == {"w": 1.0, "h": 10}


def test_replicas_one_process_per_device(tmp_path):
    sw = ParameterSweep(_params(), sweep_funcs.toy, output_dir=tmp_path, parallel=True, devices=[0, 1, 5])
    res = sw.run(show_progress=False)
    assert [r["parameters"] for r in res] == sw.parameter_combinations          # combination order, not completion order
    ok = [r for r in res if r["status"] == "success"]
    assert len(ok) == 5 and res[-1]["status"] == "error"
    by_dev = {}
    for r in ok:
        assert r["env_device"] == r["device"] and r["device"] in (0, 1, 5)
        by_dev.setdefault(r["device"], set()).add(r["pid"])
    assert all(len(p) == 1 for p in by_dev.values())                              # one persistent worker per device
    assert len(by_dev) >= 2                                                       # the queue is shared: work spreads
    assert sw.get_result_array("metric").shape == (3, 2)


def test_replica_binds_the_engine_device(tmp_path):
    sw = ParameterSweep([SweepParameter("x", [0, 1])], sweep_funcs.configured_device, output_dir=tmp_path, parallel=True,
                        devices=[3], num_workers=1)
    res = sw.run(show_progress=False)
    assert [r["dev"] for r in res] == [3, 3]
