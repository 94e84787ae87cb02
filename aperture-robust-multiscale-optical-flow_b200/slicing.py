"""Host-side plan for time-sliced multi-GPU runs; pure Python, no device code.

The stream is cut into equal time slices, one per rank.  Rank g
  * owns the events with stream time in [g*D, (g+1)*D)            -> it reports outputs for these only,
  * also processes the causal halo [g*D - 499, g*D)               -> pooling admits |dt| < 500 us
                                                                     (reference src/vFlow.cpp:1002),
  * contributes "last event per pixel" over [g*D - 499, (g+1)*D - 499) to the exchange.  Those ranges tile
    the time axis, so folding the surfaces of ranks < g in rank order (later wins) gives the exact surface of
    active events at rank g's halo start -- needed because the SAE never forgets (src/vFlow.cpp:267).
"""
from dataclasses import dataclass

import numpy as np

HALO_US = 499


@dataclass
class SlicePlan:
    rank: int
    world: int
    t_lo: int       # first stream time this rank processes (halo start)
    t_begin: int    # first stream time this rank owns
    t_end: int      # end (exclusive) of the owned range
    surf_end: int   # events before this time are this rank's share of the surface exchange


def slice_plan(rank, world, slice_us, halo_us=HALO_US):
    t_begin = rank * slice_us
    return SlicePlan(rank, world, max(0, t_begin - halo_us), t_begin, (rank + 1) * slice_us,
                     (rank + 1) * slice_us - halo_us)


def split_counts(t_stream, plan):
    """t_stream: sorted stream times (us) of the events this rank holds (from plan.t_lo).
    Returns (n_halo, n_surf): outputs of the first n_halo events are discarded; the first n_surf events
    enter the surface exchange."""
    n_halo = int(np.searchsorted(t_stream, plan.t_begin, side="left"))
    n_surf = int(np.searchsorted(t_stream, plan.surf_end, side="left"))
    return n_halo, n_surf


def last_event_surface(x, y, t_rel, width, height):
    """Reference (numpy) 'last event per pixel' of a slice, flat index x*height + y like EventMatrix
    (include/EventMatrix.h:32-34).  The CUDA path does the same with farms_slice_surface."""
    last_t = np.zeros(width * height, np.uint32)
    hit = np.zeros(width * height, np.uint8)
    q = x.astype(np.int64) * height + y.astype(np.int64)
    last_t[q] = t_rel  # numpy assigns in order: the last occurrence wins
    hit[q] = 1
    return last_t, hit


def fold(acc_t, acc_hit, new_t, new_hit):
    """Later slice wins where it has an event."""
    m = new_hit.astype(bool)
    acc_t[m] = new_t[m]
    acc_hit[m] = 1
    return acc_t, acc_hit
