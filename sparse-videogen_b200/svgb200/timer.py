"""TIME_BENCH stage timers with the reference's interface and label table (svg/timer.py:1-86), so the reference's
analysis scripts (svg/utils/extract_time.py) keep working on logs produced on top of this engine.

TIME_BENCH=0 (default): no events, no synchronisation — the hot path is untouched.
TIME_BENCH=1: accumulate per-label milliseconds in `operator_log_data` (CUDA events + a synchronise per stage,
              like the reference — this serialises the streams, so never bench with it on).
TIME_BENCH=2: as 1, and `print_operator_log_data` prints milliseconds and clears after every print.
"""
from __future__ import annotations

import os
from contextlib import ContextDecorator

import torch

ENABLE_LOGGING = int(os.getenv("TIME_BENCH", "0")) >= 1
CLEAR_LOG_DATA = int(os.getenv("TIME_BENCH", "0")) == 2

operator_log_data: dict = {}


def clear_operator_log_data():
    operator_log_data.clear()


class TimeLoggingContext(ContextDecorator):
    def __init__(self, operation_type):
        self.operation_type = operation_type
        self.start_event = None
        self.end_event = None

    def __enter__(self):
        if ENABLE_LOGGING:
            self.start_event = torch.cuda.Event(enable_timing=True)
            self.end_event = torch.cuda.Event(enable_timing=True)
            self.start_event.record()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        if ENABLE_LOGGING:
            self.end_event.record()
            torch.cuda.synchronize()
            ms = self.start_event.elapsed_time(self.end_event)
            operator_log_data[self.operation_type] = operator_log_data.get(self.operation_type, 0) + ms
        return False


time_logging_decorator = TimeLoggingContext


def format_aligned_decimal(value, max_integer_digits=8, decimal_places=2):
    return f"{value:>{max_integer_digits + 1 + decimal_places}.{decimal_places}f}"


def format_operator_log_data() -> str:
    """The table `print_operator_log_data` prints: `<label padded> : <value> ms|s`, sorted by label."""
    if not operator_log_data:
        return ""
    width = max(len(str(k)) for k in operator_log_data)
    lines = []
    for key, value in sorted(operator_log_data.items()):
        if CLEAR_LOG_DATA:
            lines.append(f"{key:<{width}} : {format_aligned_decimal(value):>4} ms")
        else:
            lines.append(f"{key:<{width}} : {format_aligned_decimal(value / 1000):>4} s")
    return "\n".join(lines)


def print_operator_log_data(module=None, input=None, output=None):
    """Forward-hook signature, like the reference (registered on the transformer, svg/timer.py:49)."""
    if not ENABLE_LOGGING:
        return
    text = format_operator_log_data()
    if CLEAR_LOG_DATA:
        clear_operator_log_data()
    print("\n\n")
    print(text)
