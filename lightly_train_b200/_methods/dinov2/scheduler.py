"""Schedules used by the DINOv2 method.

linear_warmup_schedule mirrors LT/_methods/dinov2/scheduler.py:13-34; cosine_schedule and the
CosineWarmupScheduler factor restate lightly.utils.scheduler (lightly 1.5.26), which the reference calls at
LT/_methods/dinov2/dinov2.py:576-583,602-607,648-653.
"""
from __future__ import annotations

import math


def linear_warmup_schedule(step: int, warmup_steps: int, start_value: float, end_value: float) -> float:
    """Linear ramp start_value -> end_value over `warmup_steps`, then constant.  ValueError on the same argument
    violations the reference rejects (negative step / warm-up length, negative start, non-positive end, start > end)."""
    problems = (
        (warmup_steps < 0, f"warmup_steps must be >= 0, got {warmup_steps}"),
        (step < 0, f"step must be >= 0, got {step}"),
        (start_value < 0, f"start_value must be >= 0, got {start_value}"),
        (end_value <= 0, f"end_value must be > 0, got {end_value}"),
        (start_value > end_value, f"start_value {start_value} exceeds end_value {end_value}"),
    )
    for bad, msg in problems:
        if bad:
            raise ValueError(msg)
    frac = min(step / warmup_steps, 1.0) if warmup_steps > 0 else 1.0
    return end_value if frac >= 1.0 else start_value + frac * (end_value - start_value)


def cosine_schedule(step: int, max_steps: int, start_value: float, end_value: float) -> float:
    if step < 0:
        raise ValueError(f"Current step number {step} can't be negative.")
    if max_steps < 1:
        raise ValueError(f"Total step number {max_steps} must be >= 1.")
    if max_steps == 1 or step >= max_steps:
        return end_value
    return end_value - (end_value - start_value) * (math.cos(math.pi * step / (max_steps - 1)) + 1) / 2


def cosine_warmup_factor(step: int, warmup_steps: int, max_steps: int, end_value: float, start_value: float = 1.0) -> float:
    """lr multiplier of CosineWarmupScheduler at scheduler epoch `step` (0-based)."""
    if step < warmup_steps:
        return start_value * (step + 1) / warmup_steps
    return cosine_schedule(step - warmup_steps, max_steps - warmup_steps, start_value, end_value)
