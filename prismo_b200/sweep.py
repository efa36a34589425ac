"""Parameter sweeps as a multi-GPU REPLICA scheduler (interface of
/root/reference/src/prismo/optimization/sweep.py:19-279).

The reference fans a sweep out over CPU processes (``ProcessPoolExecutor``), every worker stepping on NumPy.  Here a
sweep is N independent simulations on N GPUs: one persistent worker process per device, bound to it once
(``PRISMO_B200_DEVICE`` / ``configure(device=…)``), pulling parameter combinations from one shared queue — a GPU that
finishes early takes the next combination, and no simulation ever shares a device with another.  There is no data-path
collective: "replicas only".  Results come back in combination order (the reference's progress-bar mode returns them in
completion order).  Plotting helpers are not part of the step path and are not provided.
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Callable, Optional

import numpy as np


@dataclass
class SweepParameter:
    name: str
    values: list
    unit: str = ""

    def __len__(self) -> int:
        return len(self.values)


def _run_one(func, params):
    """One combination: the user's function, never raising (sweep.py:169-195)."""
    try:
        result = func(params)
        result["parameters"] = params
        result["status"] = "success"
        return result
    except Exception as e:                                   # reported per combination, like the reference
        return {"parameters": params, "status": "error", "error": str(e)}


def _worker(device, func, jobs, done):
    """Replica worker: binds this process to one GPU, then serves combinations until the sentinel arrives."""
    os.environ["PRISMO_B200_DEVICE"] = str(device)
    os.environ.setdefault("CUDA_DEVICE_ORDER", "PCI_BUS_ID")
    try:
        from . import session

        session.configure(device=device)
    except Exception:
        pass
    while True:
        job = jobs.get()
        if job is None:
            return
        idx, params = job
        res = _run_one(func, params)
        res.setdefault("device", device)
        done.put((idx, res))


def visible_devices() -> list:
    try:
        import torch

        n = torch.cuda.device_count()
    except Exception:
        n = 0
    return list(range(n)) or [0]


class ParameterSweep:
    def __init__(self, parameters: list, simulation_func: Callable[[dict], dict], output_dir: Optional[Path] = None,
                 parallel: bool = False, num_workers: Optional[int] = None, devices: Optional[list] = None):
        """``parallel=True`` runs one replica worker per entry of ``devices`` (default: every visible GPU, capped by
        ``num_workers``).  ``simulation_func`` must be importable from the workers (module level), as with the
        reference's process pool."""
        self.parameters = parameters
        self.simulation_func = simulation_func
        self.output_dir = Path(output_dir) if output_dir else Path("./sweep_results")
        self.output_dir.mkdir(parents=True, exist_ok=True)
        self.parallel, self.num_workers, self.devices = parallel, num_workers, devices
        self.results: list = []
        self.parameter_combinations: list = []
        self._generate_combinations()

    def _generate_combinations(self) -> None:
        if not self.parameters:
            return
        grids = np.meshgrid(*[p.values for p in self.parameters], indexing="ij")
        for i in range(grids[0].size):
            self.parameter_combinations.append({p.name: grids[k].flat[i] for k, p in enumerate(self.parameters)})

    # ---- execution --------------------------------------------------------------------------------------------
    def run(self, show_progress: bool = True) -> list:
        self.results = self._run_replicas(show_progress) if self.parallel else self._run_sequential(show_progress)
        return self.results

    def _progress(self, show):
        if not show:
            return lambda: None
        try:
            from tqdm import tqdm

            bar = tqdm(total=len(self.parameter_combinations), desc="Parameter Sweep")
            return lambda: bar.update(1)
        except Exception:
            return lambda: None

    def _run_sequential(self, show_progress: bool) -> list:
        tick = self._progress(show_progress)
        out = []
        for params in self.parameter_combinations:
            out.append(_run_one(self.simulation_func, params))
            tick()
        return out

    def _run_replicas(self, show_progress: bool) -> list:
        devices = list(self.devices) if self.devices is not None else visible_devices()
        if self.num_workers:
            devices = devices[: self.num_workers]
        devices = devices[: max(1, len(self.parameter_combinations))]
        ctx = mp.get_context("spawn")                        # CUDA contexts do not survive fork
        jobs, done = ctx.Queue(), ctx.Queue()
        for item in enumerate(self.parameter_combinations):
            jobs.put(item)
        for _ in devices:
            jobs.put(None)
        procs = [ctx.Process(target=_worker, args=(d, self.simulation_func, jobs, done), daemon=True) for d in devices]
        for p in procs:
            p.start()
        tick = self._progress(show_progress)
        out: list = [None] * len(self.parameter_combinations)
        got = 0
        while got < len(out):
            try:
                idx, res = done.get(timeout=1.0)
            except Exception:                                # queue.Empty: check that the replicas are still alive
                if not any(p.is_alive() for p in procs) and done.empty():
                    break
                continue
            out[idx] = res
            got += 1
            tick()
        for p in procs:
            p.join(timeout=10)
        for i, r in enumerate(out):                          # a replica died (e.g. the device was lost)
            if r is None:
                out[i] = {"parameters": self.parameter_combinations[i], "status": "error", "error": "replica worker died"}
        return out

    # ---- result helpers (sweep.py:197-279) ------------------------------------------------------------------
    def save_results(self, filename: str = "sweep_results.json") -> Path:
        path = self.output_dir / filename

        def plain(v):
            if isinstance(v, np.ndarray):
                return v.tolist()
            if isinstance(v, np.generic):
                return v.item()
            if isinstance(v, dict):
                return {k: plain(x) for k, x in v.items()}
            return v

        with open(path, "w") as f:
            json.dump([{k: plain(v) for k, v in r.items()} for r in self.results], f, indent=2)
        return path

    def get_result_array(self, result_key: str) -> np.ndarray:
        values = [r.get(result_key, np.nan) for r in self.results]
        return np.array(values).reshape(tuple(len(p.values) for p in self.parameters))

    def find_optimal(self, metric: str, maximize: bool = True):
        values = [r.get(metric, np.nan) for r in self.results]
        best = int(np.nanargmax(values) if maximize else np.nanargmin(values))
        return self.results[best]["parameters"], self.results[best]

    def __repr__(self) -> str:
        return f"ParameterSweep({len(self.parameters)} parameters, {len(self.parameter_combinations)} combinations)"
