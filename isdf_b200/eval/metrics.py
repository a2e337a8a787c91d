"""Step timing -- mirror of reference isdf/eval/metrics.py:13-38 (milliseconds, CUDA events on GPU)."""
import time

import torch


def start_timing():
    if torch.cuda.is_available():
        torch.cuda.synchronize()
        start = torch.cuda.Event(enable_timing=True)
        end = torch.cuda.Event(enable_timing=True)
        start.record()
        return start, end
    return time.perf_counter(), None


def end_timing(start, end):
    if torch.cuda.is_available():
        end.record()
        end.synchronize()
        return start.elapsed_time(end)
    return (time.perf_counter() - start) * 1000.0
