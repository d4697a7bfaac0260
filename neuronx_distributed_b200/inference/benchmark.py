"""Latency benchmark harness with the reference's report format (``examples/inference/modules/benchmark.py:9-73``):
1 warm-up + N timed runs, p50/p90/p95/p99/p100/avg in ms, throughput = runs·max_length·batch / total_time; per-submodule
collectors through forward hooks.  Timing is on-device (CUDA events) when a GPU is present."""
from __future__ import annotations

import time
from typing import Any, Callable, Dict, List, Optional

import numpy as np
import torch


class LatencyCollector:
    def __init__(self):
        self.latency_list: List[float] = []
        self._t0: Optional[Any] = None

    def pre_hook(self, *args):
        if torch.cuda.is_available():
            self._t0 = torch.cuda.Event(enable_timing=True)
            self._t0.record()
        else:
            self._t0 = time.perf_counter()

    def hook(self, *args):
        if torch.cuda.is_available():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.latency_list.append((self._t0, e))
        else:
            self.latency_list.append(time.perf_counter() - self._t0)

    def latencies_s(self) -> List[float]:
        out = []
        for x in self.latency_list:
            if isinstance(x, tuple):
                x[1].synchronize()
                out.append(x[0].elapsed_time(x[1]) / 1e3)
            else:
                out.append(x)
        return out

    def percentile(self, p: float) -> float:
        l = self.latencies_s()
        return float(np.percentile(l, p)) if l else 0.0


def generate_report(latencies_s: List[float], max_length: int, max_batch_size: int, n_runs: Optional[int] = None) -> Dict[str, float]:
    l = np.array(latencies_s)
    total = float(l.sum())
    n = n_runs if n_runs is not None else len(l)
    return {
        "latency_ms_p50": float(np.percentile(l, 50)) * 1e3, "latency_ms_p90": float(np.percentile(l, 90)) * 1e3,
        "latency_ms_p95": float(np.percentile(l, 95)) * 1e3, "latency_ms_p99": float(np.percentile(l, 99)) * 1e3,
        "latency_ms_p100": float(np.percentile(l, 100)) * 1e3, "latency_ms_avg": float(l.mean()) * 1e3,
        "throughput": n * max_length * max_batch_size / total if total > 0 else 0.0,
    }


class Benchmark:
    def __init__(self, benchmark_func: Callable, input_param: Any = None, config: Any = None, num_runs: int = 20,
                 preprocess_func: Optional[Callable] = None):
        self.f, self.inp, self.num_runs, self.pre = benchmark_func, input_param, num_runs, preprocess_func
        self.latency_list: List[float] = []

    def _call(self):
        if self.pre is not None:
            self.pre()
        if isinstance(self.inp, (tuple, list)):
            return self.f(*self.inp)
        if isinstance(self.inp, dict):
            return self.f(**self.inp)
        return self.f() if self.inp is None else self.f(self.inp)

    def run(self) -> List[float]:
        self._call()    # warm-up
        cuda = torch.cuda.is_available()
        for _ in range(self.num_runs):
            if cuda:
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                self._call()
                e.record()
                torch.cuda.synchronize()
                self.latency_list.append(s.elapsed_time(e) / 1e3)
            else:
                t0 = time.perf_counter()
                self._call()
                self.latency_list.append(time.perf_counter() - t0)
        return self.latency_list
