"""N > 1 on CPU: two gloo ranks shard one stack (contiguous chunks), speculate the chunk-boundary
state, verify it by exchanging 10 doubles, and must reproduce the sequential loop exactly.  The
device is replaced by the oracle behind the library's resolver; the sharding logic under test
(laser_spot_track_b200.sharded) is the one bench.py uses on GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _stack():
    from laser_spot_track_b200.synth import StackConfig, SyntheticStack
    cfg = StackConfig(360, 280, "u8", 24, 12, 40, spot_travel=1.2)
    st = SyntheticStack(cfg)
    return cfg, st, [st.frame(i) for i in range(0, 25)]


def _worker(rank, world, port, mode, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import util
        from laser_spot_track_b200 import native as N
        from laser_spot_track_b200.sharded import ShardedTracker, TorchComm, chunk_bounds
        cfg, st, frames = _stack()
        op = util.oracle_params(cfg, match_intensity=False)
        trk = util.CpuResolverTracker(op, frames[0], st.ref_xy, N.PIX_U8, max_batch=5)
        sh = ShardedTracker(trk, TorchComm())
        lo, hi = chunk_bounds(len(frames) - 1, world)[rank]
        chunk = frames[1 + lo:1 + hi]
        lead = frames[1 + lo - 2:1 + lo] if (mode == "lead_in" and rank > 0) else None
        if mode == "lead_in_folded":      # the chunk handed over WITH the two slices before it (rank 0: the last two of nothing -- its own first two twice, ignored)
            chunk = (frames[1 + lo - 2:1 + lo] if rank > 0 else frames[1:3]) + chunk
        if mode == "pipelined":
            # two independent passes over the same stack: the exchange of the first travels while the second is tracked
            _, _, none = sh.track_chunk_pipelined(chunk, 1 + lo, [0.0] * 10)
            assert none is None
            _, _, first = sh.track_chunk_pipelined(chunk, 1 + lo, [0.0] * 10)
            second = sh.finish()
            assert first is not None and second is not None
            key = lambda rr: [(int(r.frame), r.status, [(t.ix, t.iy, t.score) for t in r.tm]) for r in rr]
            assert key(first[0]) == key(second[0]) and first[1] == second[1]
            recs, end = second
        else:
            recs, end = sh.track_chunk(chunk, 1 + lo, [0.0] * 10, lead_in=lead, lead_in_count=2 if mode == "lead_in_folded" else 0)
        rows = [(int(r.frame), r.status, [(t.ix, t.iy, t.wx, t.wy, t.score, t.x, t.y) for t in r.tm],
                 list(r.dis), (r.dX_pix, r.dY_pix, r.x_abs, r.y_abs, r.dL)) for r in recs]
        q.put((rank, rows, end, sh.retracked))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["zero_guess", "lead_in", "lead_in_folded", "pipelined"])
def test_two_ranks_reproduce_sequential_loop(built_lib, oracle_clib, mode):
    import torch.multiprocessing as mp
    import util
    from oracle import lst_oracle as O
    cfg, st, frames = _stack()
    op = util.oracle_params(cfg, match_intensity=False)
    seq = O.OracleTracker(op)
    seq.set_reference(frames[0], st.ref_xy)
    want = seq.run(frames[1:])
    # the state at the chunk boundary must differ from the reference state, else nothing is tested
    assert any(O.jint(v) != 0 for d in want[11].dis for v in d)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = got[0][1] + got[1][1]
    assert len(rows) == len(want)
    for i, (row, w) in enumerate(zip(rows, want)):
        frame, status, tms, dis, out = row
        assert frame == i + 1 and status == w.status
        for k in range(5):
            t = w.tm[k]
            assert tms[k] == (t.ix, t.iy, t.wx, t.wy, t.score, t.x, t.y), (i, k)
        assert dis == [v for d in w.dis for v in d]
        assert out == (w.dX_pix, w.dY_pix, w.x_abs, w.y_abs, w.dL)
    assert got[1][2] == [v for d in seq.dis for v in d]
    if mode == "zero_guess":
        assert got[1][3] == 1, "rank 1 speculated the reference state and must have re-tracked"
    elif mode == "pipelined":
        assert got[1][3] == 2, "both passes of rank 1 started from the reference state and must have been re-tracked"
    else:
        assert got[1][3] == 0, "the lead-in should have found the right windows"


def test_chunk_bounds_and_window_equivalence():
    from laser_spot_track_b200.sharded import chunk_bounds, same_windows
    assert chunk_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert chunk_bounds(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert same_windows([0.3, -0.9] * 5, [0.9, -0.1] * 5)          # (int) truncates toward zero
    assert not same_windows([1.0] + [0.0] * 9, [0.999] + [0.0] * 9)
