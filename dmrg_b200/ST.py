"""Suzuki-Trotter task lines — same interface as the reference's `DMRG/ST.py`.

A *task line* is a list of `(task_type, duration)` steps executed in order; a *pipeline* is a list
of worker factories, one per task type: `funs[t](duration)` builds a zero-argument worker that
performs task `t` for that duration (`DMRG/ST.py:6-34`).  `do_tasks` first merges neighbouring steps
of the same type, builds ONE worker per distinct (type, duration) pair — for TEBD that is one
`expm` and one gate upload per distinct time step — and then calls them in order.

First- and second-order lines follow the reference (`ST1_tasks`, `ST2_tasks`, `DMRG/ST.py:62-93`).
`ST4_tasks` / `ST4` (fourth-order Forest-Ruth/Suzuki composition) are new: BASELINE config 3 names a
4th-order scheme that the reference does not have.
"""
from typing import Callable, List, Tuple

Taskline = List[Tuple[int, float]]
Pipeline = List[Callable[[float], Callable]]


def streamline(tasks: Taskline) -> Taskline:
    """Merge runs of equal task type by adding their durations (`DMRG/ST.py:37-51`).

    >>> streamline([(1, 0.5), (1, 0.8), (2, 0.4)])
    [(1, 1.3), (2, 0.4)]
    """
    merged = []
    for kind, span in tasks:
        if merged and merged[-1][0] == kind:
            merged[-1][1] += span
        else:
            merged.append([kind, span])
    return [tuple(step) for step in merged]


def do_tasks(tasks: Taskline, funs: Pipeline) -> None:
    """Streamline, build one worker per distinct step, run them in order (`DMRG/ST.py:54-59`)."""
    line = streamline(tasks)
    workers = {}
    for step in line:
        if step not in workers:
            workers[step] = funs[step[0]](step[1])
    for step in line:
        workers[step]()


def ST1_tasks(L: int, n: int) -> Taskline:
    """First order: every operator for 1/n, repeated n times (`DMRG/ST.py:62-75`)."""
    return [[k, 1. / n] for _ in range(n) for k in range(L)]


def ST2_tasks(L: int, n: int) -> Taskline:
    """Second order (symmetric): operators forward then backward for 1/(2n) each, n times (`DMRG/ST.py:78-93`)."""
    forward = [[k, 0.5 / n] for k in range(L)]
    return [list(step) for _ in range(n) for step in forward + forward[::-1]]


def ST4_tasks(L: int, n: int) -> Taskline:
    """Fourth order: S4(tau) = S2(p tau)^2 S2((1-4p) tau) S2(p tau)^2 with p = 1/(4 - 4^(1/3)),
    written out as a task line so that `streamline` merges the half steps like it does for ST2."""
    p = 1.0 / (4.0 - 4.0 ** (1.0 / 3.0))
    line = []
    for _ in range(n):
        for frac in (p, p, 1.0 - 4.0 * p, p, p):
            half = [[k, 0.5 * frac / n] for k in range(L)]
            line += [list(step) for step in half + half[::-1]]
    return line


def ST1(funs: Pipeline, n: int) -> None:
    """Run the pipeline with the first-order line (`DMRG/ST.py:96-105`)."""
    do_tasks(ST1_tasks(len(funs), n), funs)


def ST2(funs: Pipeline, n: int) -> None:
    """Run the pipeline with the second-order line (`DMRG/ST.py:108-117`)."""
    do_tasks(ST2_tasks(len(funs), n), funs)


def ST4(funs: Pipeline, n: int) -> None:
    """Run the pipeline with the fourth-order line (new)."""
    do_tasks(ST4_tasks(len(funs), n), funs)
