"""Stand-in for the reference's ``elegantrl/logger.py``, which is missing from its checkout (swallowed by the ``log*`` rule
of its .gitignore) although six files import it.  The call sites fix the interface -- a Stable-Baselines3
style module-level logger:

    elegantrl/utils.py:26-47      logger.configure(save_path, ["stdout", "tensorboard", "csv"]) / configure(format_strings=[""])
    elegantrl/agent.py:335,659-662 logger.record(key, value[, exclude=...])
    elegantrl/run.py:183,222      logger.dump(step=total_step)
    utils/test.py:11-12           from elegantrl.logger import Figure

Outputs: ``progress.csv`` (header + one row per dump, rewritten when a new key appears), a tensorboard event file when
``torch.utils.tensorboard`` is importable (the reference has the same condition, elegantrl/utils.py:5-8), a key / value
table on stdout.  Nothing here touches the GPU; values may be python numbers, numpy scalars or 0-d tensors.
"""
from __future__ import annotations

import csv
import os
import sys
import time
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple, Union

try:  # same optional dependency as the reference (elegantrl/utils.py:5-8)
    from torch.utils.tensorboard import SummaryWriter
except Exception:  # noqa: BLE001
    SummaryWriter = None

DEBUG, INFO, WARN, ERROR, DISABLED = 10, 20, 30, 40, 50


class Figure:
    """A matplotlib figure handed to record(); only the tensorboard writer can store it (SB3 semantics)."""

    def __init__(self, figure, close: bool = True):
        self.figure, self.close = figure, close


def _as_float(v) -> Optional[float]:
    try:
        return float(v.item()) if hasattr(v, "item") else float(v)
    except (TypeError, ValueError):
        return None


class _Writer:
    name = ""

    def write(self, row: Dict[str, Any], excluded: Dict[str, Tuple[str, ...]], step: int) -> None:
        raise NotImplementedError

    def close(self) -> None:
        pass

    def _visible(self, row, excluded):
        return {k: v for k, v in row.items() if self.name not in excluded.get(k, ())}


class _Stdout(_Writer):
    name = "stdout"

    def __init__(self, stream=None):
        self.stream = stream or sys.stdout

    def write(self, row, excluded, step):
        items = [(k, v) for k, v in sorted(self._visible(row, excluded).items()) if not isinstance(v, Figure)]
        if not items:
            return
        cells = [(k, f"{v:.4g}" if isinstance(v, float) else str(v)) for k, v in items]
        wk, wv = max(len(k) for k, _ in cells), max(len(v) for _, v in cells)
        bar = "-" * (wk + wv + 7)
        self.stream.write("\n".join([bar] + [f"| {k:<{wk}} | {v:<{wv}} |" for k, v in cells] + [bar]) + "\n")
        self.stream.flush()


class _Csv(_Writer):
    """progress.csv: one header line, one line per dump; when a dump brings a new key the file is rewritten with the wider
    header (earlier rows get empty cells), so that the file always parses as one table."""
    name = "csv"

    def __init__(self, path: str):
        self.path, self.keys, self.rows = path, [], []

    def write(self, row, excluded, step):
        vis = {k: v for k, v in self._visible(row, excluded).items() if not isinstance(v, Figure)}
        vis["step"] = step
        new = [k for k in sorted(vis) if k not in self.keys]
        self.rows.append(vis)
        if new or not os.path.exists(self.path):
            self.keys += new
            with open(self.path, "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(self.keys)
                for r in self.rows:
                    w.writerow([r.get(k, "") for k in self.keys])
        else:
            with open(self.path, "a", newline="") as f:
                csv.writer(f).writerow([vis.get(k, "") for k in self.keys])


class _Tensorboard(_Writer):
    name = "tensorboard"

    def __init__(self, folder: str):
        self.writer = SummaryWriter(log_dir=folder)

    def write(self, row, excluded, step):
        for k, v in sorted(self._visible(row, excluded).items()):
            if isinstance(v, Figure):
                self.writer.add_figure(k, v.figure, step, close=v.close)
            elif isinstance(v, str):
                self.writer.add_text(k, v, step)
            elif _as_float(v) is not None:
                self.writer.add_scalar(k, _as_float(v), step)
        self.writer.flush()

    def close(self):
        self.writer.close()


def _make_writer(fmt: str, folder: Optional[str]) -> Optional[_Writer]:
    if fmt == "stdout":
        return _Stdout()
    if fmt in ("", "none"):
        return None
    if folder is None:
        return None
    if fmt == "csv":
        return _Csv(os.path.join(folder, "progress.csv"))
    if fmt == "tensorboard":
        return _Tensorboard(folder) if SummaryWriter is not None else None
    if fmt == "log":
        return _Stdout(open(os.path.join(folder, "log.txt"), "a"))
    raise ValueError(f"unknown logger format {fmt!r}")


# ---------------------------------------------------------------------------------------------------- module state
# `values` and `history` are kept as the SAME objects for the life of the module (callers hold references to them).
values: Dict[str, Any] = {}                 # key -> value recorded since the last dump
history: List[Dict[str, Any]] = []          # one dict per dump (+ "step"), kept in memory for tests and notebooks
_counts: Dict[str, int] = {}
_excluded: Dict[str, Tuple[str, ...]] = {}
_writers: List[_Writer] = []
_dir: Optional[str] = None
_level = INFO


def configure(folder: Optional[str] = None, format_strings: Optional[Sequence[str]] = None, verbose: Optional[int] = None) -> None:
    """configure(save_path, ["stdout", "tensorboard", "csv"]) / configure(format_strings=[""]) (elegantrl/utils.py:40-47).
    ``verbose`` is accepted for the callers of the round-1 helper (1: also print)."""
    global _dir
    for w in _writers:
        w.close()
    _writers.clear()
    if format_strings is None:
        format_strings = ["csv", "tensorboard"] if folder else []
        if verbose:
            format_strings = ["stdout"] + list(format_strings)
    if folder:
        os.makedirs(folder, exist_ok=True)
    _dir = folder
    for fmt in format_strings:
        w = _make_writer(fmt, folder)
        if w is not None:
            _writers.append(w)
    values.clear()
    _counts.clear()
    _excluded.clear()


def get_dir() -> Optional[str]:
    return _dir


def _excl(exclude) -> Tuple[str, ...]:
    if exclude is None:
        return ()
    return (exclude,) if isinstance(exclude, str) else tuple(exclude)


def record(key: str, value: Any, exclude: Optional[Union[str, Iterable[str]]] = None) -> None:
    """Keep ``value`` under ``key`` until the next dump; ``exclude`` names the output formats that must skip the key."""
    if isinstance(value, Figure) or isinstance(value, str):
        values[key] = value
    else:
        f = _as_float(value)
        values[key] = value if f is None else f
    _excluded[key] = _excl(exclude)


def record_mean(key: str, value: Any, exclude: Optional[Union[str, Iterable[str]]] = None) -> None:
    f = _as_float(value)
    if f is None:
        return
    n = _counts.get(key, 0)
    values[key] = f if n == 0 else (values[key] * n + f) / (n + 1)
    _counts[key] = n + 1
    _excluded[key] = _excl(exclude)


def dump(step: int = 0) -> None:
    """Write everything recorded since the last dump to every configured output and clear it."""
    row = dict(values)
    history.append(dict(row, step=int(step)))
    for w in _writers:
        w.write(row, _excluded, int(step))
    values.clear()
    _counts.clear()
    _excluded.clear()


def set_level(level: int) -> None:
    global _level
    _level = level


def log(*args, level: int = INFO) -> None:
    if level >= _level:
        print(time.strftime("[%H:%M:%S]"), *args)


def debug(*args) -> None:
    log(*args, level=DEBUG)


def info(*args) -> None:
    log(*args, level=INFO)


def warn(*args) -> None:
    log(*args, level=WARN)


def error(*args) -> None:
    log(*args, level=ERROR)


def close() -> None:
    for w in _writers:
        w.close()
    _writers.clear()
