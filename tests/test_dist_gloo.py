"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU paths.

PDLP: every rank holds a row block of A (easylp_b200.partition.row_block) and a replica of x.  The identities the
CUDA path relies on are checked with real collectives: local A_g x is the rank's slice of A x; the allreduce(sum) of
the partial A_g' y_g is A' y; scalar partial sums ride in the tail of the same buffer (one collective per check).
One full distributed PDHG iteration built from those pieces equals the single-process iteration of the oracle.
Batched simplex: contiguous LP ranges, no collective — concatenating the ranks' results equals the whole batch."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as td
    from easylp_b200.partition import lp_ranges, row_block, row_cuts
    from oracle import cbind, gen
    from oracle.pdlp_ref import dual_prox, row_bounds

    td.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        p = gen.sparse_planted(600, seed=5)
        m, n = p["m"], p["n"]
        q, r0, r1 = row_block(p, rank, world)
        cuts = row_cuts(p["row_ptr"], world)
        assert cuts[0] == 0 and cuts[-1] == m and (r0, r1) == (cuts[rank], cuts[rank + 1])
        nnz_local = int(q["row_ptr"][-1])
        assert abs(nnz_local - int(p["row_ptr"][m]) / world) <= 20          # balanced by non-zeros

        rng = np.random.default_rng(1)                                        # same stream on every rank = replicas
        x = rng.normal(size=n)
        y = rng.normal(size=m)
        rp_full = p["row_ptr"].astype(np.int64)
        ax_full = gen._csr_matvec(rp_full, p["col_idx"], p["vals"], x, m)
        aty_full = gen._csr_rmatvec(rp_full, p["col_idx"], p["vals"], y, n)

        rp = q["row_ptr"].astype(np.int64)
        ax_loc = gen._csr_matvec(rp, q["col_idx"], q["vals"], x, q["m"])
        assert np.array_equal(ax_loc, ax_full[r0:r1])                          # local rows need no exchange

        # partial A_g' y_g with the scalar partials packed behind it: ONE allreduce (pdlp.cu check_iteration)
        part = gen._csr_rmatvec(rp, q["col_idx"], q["vals"], y[r0:r1], n)
        tail = np.array([np.sum(ax_loc ** 2), np.dot(y[r0:r1], ax_loc)])
        buf = torch.from_numpy(np.concatenate([part, tail]))
        td.all_reduce(buf, op=td.ReduceOp.SUM)
        buf = buf.numpy()
        assert np.allclose(buf[:n], aty_full, rtol=1e-13, atol=1e-13)
        assert np.isclose(buf[n], np.sum(ax_full ** 2), rtol=1e-12) and np.isclose(buf[n + 1], np.dot(y, ax_full), rtol=1e-12)

        # one PDHG step T(z) assembled the distributed way == the single-process step
        tau, sigma = 0.3, 0.2
        lc, uc = row_bounds(p["sense"], p["rhs"])
        xp = np.clip(x - tau * (p["c"] - buf[:n]), p["lb"], p["ub"])          # replicated primal update
        xbar = 2 * xp - x
        axb_loc = gen._csr_matvec(rp, q["col_idx"], q["vals"], xbar, q["m"])
        yp_loc = dual_prox(y[r0:r1] - sigma * axb_loc, sigma, lc[r0:r1], uc[r0:r1])
        xp1 = np.clip(x - tau * (p["c"] - aty_full), p["lb"], p["ub"])
        yp1 = dual_prox(y - sigma * gen._csr_matvec(rp_full, p["col_idx"], p["vals"], 2 * xp1 - x, m), sigma, lc, uc)
        assert np.allclose(xp, xp1, rtol=1e-12, atol=1e-12)
        assert np.allclose(yp_loc, yp1[r0:r1], rtol=1e-12, atol=1e-12)

        # batched path: contiguous ranges, no collective
        d = gen.dense_batch(B=101, seed=3)
        b0, b1 = lp_ranges(d["B"], world)[rank]
        st, obj, xs, _ = cbind.simplex_batch(d["A"][b0:b1], d["b"][b0:b1], d["c"][b0:b1], d["lb"][b0:b1], d["ub"][b0:b1],
                                             d["sense"][b0:b1])
        gathered = [None] * world
        td.all_gather_object(gathered, (b0, b1, st.tolist(), obj.tolist()))      # test-side gather only
        if rank == 0:
            s_all, o_all, _, _ = cbind.simplex_batch(d["A"], d["b"], d["c"], d["lb"], d["ub"], d["sense"])
            cat_s = sum((g[2] for g in sorted(gathered)), [])
            cat_o = sum((g[3] for g in sorted(gathered)), [])
            assert [g[:2] for g in sorted(gathered)] == lp_ranges(d["B"], world)
            assert cat_s == s_all.tolist() and cat_o == o_all.tolist()
        out.put((rank, "ok"))
    except BaseException as e:       # noqa: BLE001 - report to the parent
        out.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        td.destroy_process_group()


@pytest.mark.timeout(180)
def test_row_partition_and_single_collective_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(out.get(timeout=150) for _ in procs)
    for pr in procs:
        pr.join(30)
    assert res == [(0, "ok"), (1, "ok")], res


def test_partition_edge_cases():
    sys.path.insert(0, ROOT)
    from easylp_b200.partition import lp_ranges, row_cuts
    assert row_cuts(np.array([0]), 4) == [0, 0, 0, 0, 0]                      # no rows
    assert row_cuts(np.array([0, 5]), 3)[-1] == 1                              # fewer rows than ranks
    c = row_cuts(np.array([0, 0, 0, 10, 10, 20]), 2)
    assert c[0] == 0 and c[-1] == 5 and c[1] in (2, 3)
    assert lp_ranges(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert lp_ranges(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
