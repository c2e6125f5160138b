"""Multi-GPU sharding of the path: the unit is the reference view.

Inside a stage, ProcessProblem(i) touches only view i's outputs (reference main.cpp:431-446), so views are dealt
out to the ranks with no data-path collective.  The only cross-view input is the neighbours' DEPTH maps of the two
geometric-consistency stages (ACMMP.cpp:653-678 -> ComputeGeomConsistencyCost): between stages every rank
all-gathers the depth maps of the views it owns.  This module is the host logic of that plan; it runs on any
torch.distributed backend (NCCL over NVLink on the B200 box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def owner_of(view: int, world: int) -> int:
    """Round-robin deal: view v lives on rank v mod G (neighbouring views land on different GPUs, so the
    per-stage load is even whatever the view order)."""
    return view % world


def views_of(rank: int, n_views: int, world: int) -> list:
    return [v for v in range(n_views) if owner_of(v, world) == rank]


def rounds(n_views: int, world: int) -> int:
    """Views are processed in lock-step rounds (one view per rank per round) so that the all-gather between
    stages has the same shape on every rank; ranks that run out of views contribute a dummy map."""
    return -(-n_views // world)


def gather_slot(view: int, world: int) -> tuple:
    """Where view's depth map sits after the all-gather of round r: (round, rank)."""
    return view // world, owner_of(view, world)


def neighbour_sources(src_ids, round_index: int, world: int, computed_rounds=None) -> list:
    """For every source view of a problem: ("gathered", rank) when that view's map is part of THIS round's
    all-gather, ("stored", view) when an earlier round of the same stage produced it (computed_rounds = rounds
    already finished), else ("input", view): the map the stage started from (previous stage's output)."""
    out = []
    for v in src_ids:
        r, k = gather_slot(v, world)
        if r == round_index:
            out.append(("gathered", k))
        elif computed_rounds is not None and r in computed_rounds:
            out.append(("stored", v))
        else:
            out.append(("input", v))
    return out


class DepthExchange:
    """All-gather of one depth map per rank; returns the gathered tensor [world, H, W] on the ranks' device."""

    def __init__(self, dist, world: int):
        self.dist, self.world = dist, world
        self.bytes = 0

    def all_gather(self, mine):
        import torch
        gathered = torch.empty((self.world,) + tuple(mine.shape), dtype=mine.dtype, device=mine.device)
        if self.world == 1:
            gathered[0].copy_(mine)
            return gathered
        if self.dist.get_backend() == "nccl":
            self.dist.all_gather_into_tensor(gathered.view(-1), mine.contiguous().view(-1))
        else:       # gloo (CPU tests): list form
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            self.dist.all_gather(parts, mine.contiguous())
            for k, part in enumerate(parts):
                gathered[k].copy_(part)
        self.bytes += mine.numel() * mine.element_size() * (self.world - 1)
        return gathered

    def pick(self, gathered, src_ids, round_index: int, fallbacks):
        """One map per source view: the freshly gathered one where a rank computed it this round, else fallbacks[k]."""
        plan = neighbour_sources(src_ids, round_index, self.world)
        return [gathered[w] if kind == "gathered" and gathered[w].shape == fallbacks[k].shape else fallbacks[k]
                for k, (kind, w) in enumerate(plan)]
