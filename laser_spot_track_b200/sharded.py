"""Frame sharding of one stack over several GPUs (north_star item 4).

One process per GPU.  The stack is cut into ``world`` contiguous chunks; rank r tracks chunk r.
Nothing is reduced between GPUs -- there is no data-path collective.  The only coupling between
chunks is the reference's frame-to-frame recurrence: the search windows of the first frame of
chunk r are placed from the *integer-truncated* displacements left by the last accepted frame of
chunk r-1 (LST4:1327-1328, 832-846).  That is 10 numbers per boundary, handled by
speculate-and-verify:

1. every rank tracks its chunk from a *speculated* start state (optionally refined by tracking a
   short lead-in of the previous chunk's last frames first);
2. the ranks exchange their end states and speculations (21 doubles each, host side, one all_gather);
3. walking the boundaries in order, a rank whose speculated start state places the windows
   differently from the true state re-tracks its chunk (rare: the state only matters through
   ``(int)dis``), and passes its corrected end state on.

``comm`` is any object with ``all_gather(list_of_10_floats) -> list over ranks`` and
``broadcast(list_or_None, src) -> list`` (see TorchComm); ``tracker`` is anything with
``set_state(dis)``, ``get_state()`` and ``track(frames, first_frame_index) -> records``
(LaserSpotTrack4, or the CPU stand-in of the tests).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence


def _jint(v: float) -> int:
    if v != v:
        return 0
    return int(math.trunc(v))


def same_windows(a: Sequence[float], b: Sequence[float], doubling: bool = False) -> bool:
    """Two displacement states give the same records from here on: they place all five search windows identically
    (LST4:1327-1328) and -- with search-area doubling, where mark1's un-truncated displacement shifts the other four
    windows when mark1 is lost and found again (LST4:1757) -- agree exactly in the two mark1 components."""
    if not all(_jint(x) == _jint(y) for x, y in zip(a, b)):
        return False
    return not doubling or (a[2] == b[2] and a[3] == b[3])


def chunk_bounds(n_frames: int, world: int):
    """Contiguous chunks, sizes differing by at most one frame."""
    base, extra = divmod(n_frames, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def _tail(seq, k: int):
    """seq[k:] -- for ctypes arrays (slice pointers, records) a view of the same memory, not a copy."""
    import ctypes as C
    if k <= 0:
        return seq
    if isinstance(seq, C.Array):
        n = len(seq) - k
        return (seq._type_ * n).from_buffer(seq, k * C.sizeof(seq._type_))
    return seq[k:]


class TorchComm:
    """Host-side exchange over a torch.distributed process group (gloo on CPU tensors; with an NCCL
    group the 10 doubles travel as a tiny device tensor).  Not on the data path."""

    def __init__(self, group=None, device="cpu"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group, self.device = torch, dist, group, device
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_gather(self, vals: Sequence[float]) -> List[List[float]]:
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.device)
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t, group=self.group)
        return [o.cpu().tolist() for o in out]

    def broadcast(self, vals: Optional[Sequence[float]], src: int, n: int = 10) -> List[float]:
        t = self.torch.zeros(n, dtype=self.torch.float64, device=self.device)
        if self.rank == src:
            t[:] = self.torch.tensor(list(vals), dtype=self.torch.float64)
        self.dist.broadcast(t, src=src, group=self.group)
        return t.cpu().tolist()


class LocalComm:
    """world == 1."""
    rank, world = 0, 1

    def all_gather(self, vals):
        return [list(vals)]

    def broadcast(self, vals, src, n=10):
        return list(vals)


class ShardedTracker:
    def __init__(self, tracker, comm=None):
        self.tracker = tracker
        self.comm = comm or LocalComm()
        self.rank, self.world = self.comm.rank, self.comm.world
        self.retracked = 0
        params = getattr(tracker, "params", None)      # LaserSpotTrack4 keeps its lst_params; the CPU stand-ins of the tests may not
        self.doubling = bool(getattr(params, "doubling", 0)) or bool(getattr(tracker, "doubling", False))

    def _track(self, frames, first_frame_index, start_state, lead_in, lead_in_count=0):
        tr = self.tracker
        guess = list(start_state)
        if lead_in_count > 0:
            # `frames` starts with the last lead_in_count slices of the previous chunk: ONE tracking call covers them and the
            # chunk (no extra call with its own passes and synchronisations); the state after the lead-in -- the refined
            # speculation -- is what the last lead-in record carries (a skipped slice carries the state it kept)
            L = lead_in_count
            chunk = _tail(frames, L)
            if self.rank > 0:
                tr.set_state(guess)
                recs_all = tr.track(frames, first_frame_index - L)
                guess = [float(v) for v in recs_all[L - 1].dis]
                recs = _tail(recs_all, L)
            else:
                tr.set_state(guess)
                recs = tr.track(chunk, first_frame_index)
            end = list(tr.get_state())
            accepted_any = any(r.status == 0 for r in recs)
            return recs, end, guess, end + guess + [1.0 if accepted_any else 0.0], chunk
        if self.rank > 0 and lead_in is not None and len(lead_in) > 0:
            tr.set_state(guess)
            tr.track(lead_in, first_frame_index - len(lead_in))
            guess = list(tr.get_state())
        tr.set_state(guess)
        recs = tr.track(frames, first_frame_index)
        end = list(tr.get_state())
        accepted_any = any(r.status == 0 for r in recs)
        return recs, end, guess, end + guess + [1.0 if accepted_any else 0.0], frames

    def _verify(self, gathered, frames, first_frame_index, recs, end, guess):
        """Boundary verification.  One all_gather of (end state, speculated start, any-accepted flag)
        lets every rank walk the chain of boundaries identically; in the common case (every
        speculation placed the windows right) that is all.  From the first wrong speculation on,
        the ranks fall back to handing the verified state down one boundary at a time."""
        tr = self.tracker
        verified = list(gathered[0][:10])          # rank 0 started from the true state
        first_bad = None
        for r in range(1, self.world):
            g_end, g_guess, g_any = gathered[r][:10], gathered[r][10:20], gathered[r][20] != 0.0
            if not same_windows(verified, g_guess, self.doubling):
                first_bad = r
                break
            if r == self.rank:
                recs, end = self._patch_carried_state(recs, end, guess, verified)
            verified = g_end if g_any else verified
        if first_bad is not None:
            for r in range(first_bad, self.world):
                # `verified` is the true state before chunk r on every rank that needs it (r's owner
                # got it from the walk above or from the previous iteration's broadcast)
                if self.rank == r:
                    if not same_windows(verified, guess, self.doubling):
                        tr.set_state(verified)
                        recs = tr.track(frames, first_frame_index)
                        end = list(tr.get_state())
                        self.retracked += 1
                    else:
                        recs, end = self._patch_carried_state(recs, end, guess, verified)
                    accepted_any = any(x.status == 0 for x in recs)
                    if not accepted_any:
                        end = list(verified)
                verified = self.comm.broadcast(end if self.rank == r else None, src=r, n=10)
        return recs, end

    def track_chunk(self, frames, first_frame_index: int, start_state: Sequence[float],
                    lead_in: Optional[Sequence] = None, lead_in_count: int = 0):
        """Track this rank's chunk.  ``start_state``: speculated displacement state before the chunk
        (rank 0: the true one).  ``lead_in``: optionally the last few frames of the previous chunk,
        tracked first from ``start_state`` to refine the speculation -- or ``lead_in_count`` > 0: ``frames``
        itself starts with that many slices of the previous chunk and one tracking call covers both
        (``first_frame_index`` is still the chunk's).  Returns (records of the chunk, end_state);
        both are final after the boundary verification."""
        import os, time
        dbg = os.environ.get("LST_SHARD_DEBUG")
        t0 = time.perf_counter()
        recs, end, guess, vals, frames = self._track(frames, first_frame_index, start_state, lead_in, lead_in_count)
        t1 = time.perf_counter()
        gathered = self.comm.all_gather(vals)
        t2 = time.perf_counter()
        before = self.retracked
        recs, end = self._verify(gathered, frames, first_frame_index, recs, end, guess)
        t3 = time.perf_counter()
        self.tracker.set_state(end)
        if dbg:
            import sys
            print(f"[shard rank {self.rank}] track {1e3 * (t1 - t0):.1f} ms, gather {1e3 * (t2 - t1):.1f} ms, verify {1e3 * (t3 - t2):.1f} ms"
                  f"{' (re-tracked)' if self.retracked != before else ''}", file=sys.stderr, flush=True)
        return recs, end

    # ---- the same with the boundary exchange off the critical path --------------------------------
    def track_chunk_pipelined(self, frames, first_frame_index: int, start_state: Sequence[float],
                              lead_in: Optional[Sequence] = None, lead_in_count: int = 0):
        """Like track_chunk, but the exchange of THIS chunk's boundary states travels (on a helper thread) while the
        caller goes on -- typically into the next independent stack -- and the PREVIOUS call's exchange is completed
        after this chunk has been tracked.  Returns (records, end_state, previous) where ``previous`` is the verified
        (records, end_state) of the previous call, or None.  ``finish()`` completes the last one.  A boundary that
        fails verification re-tracks that chunk as track_chunk does; its records must not have been overwritten, so
        callers alternate their record buffers.  ``retracked_late`` counts such late re-tracks: whatever was tracked
        in between started from that chunk's unverified end state and has to be looked at by the caller."""
        recs, end, guess, vals, frames = self._track(frames, first_frame_index, start_state, lead_in, lead_in_count)
        prev, self._pending = getattr(self, "_pending", None), None
        if not hasattr(self, "_pool"):
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=1)
            self.retracked_late = 0
        done = self._complete(prev) if prev else None
        self._pending = (self._pool.submit(self.comm.all_gather, vals), frames, first_frame_index, recs, end, guess)
        return recs, end, done

    def _complete(self, pending):
        fut, frames, first, recs, end, guess = pending
        before = self.retracked
        keep = list(self.tracker.get_state())
        out = self._verify(fut.result(), frames, first, recs, end, guess)
        if self.retracked != before:
            self.retracked_late += 1
        self.tracker.set_state(keep)       # the tracker has moved on; a late re-track must not pull it back
        return out

    def finish(self):
        """Completes the exchange left in flight by the last track_chunk_pipelined call."""
        pending, self._pending = getattr(self, "_pending", None), None
        return self._complete(pending) if pending else None

    @staticmethod
    def _patch_carried_state(recs, end, guess, true_start):
        all_skipped = True
        for r in recs:
            if r.status == 0:
                all_skipped = False
                break
            for k in range(10):
                r.dis[k] = true_start[k]
        if all_skipped:
            end = list(true_start)
        return recs, end
