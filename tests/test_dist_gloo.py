"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU paths.

PDLP: every rank holds a ROW block of A (easylp_b200.partition.row_block; K2 = A x-bar + dual update on its rows,
gathering the whole x-bar) and a COLUMN block of A over all rows (K1 = A'y + primal update on its columns, gathering
the whole y in the padded layout y_full[N][mb]).  Nothing is replicated; the blocks of x-bar and y are all-gathered.
The identities the CUDA path relies on are checked with real collectives: the column block assembled from the ranks'
row-block transposes equals the column slice of A; K1/K2 on the blocks + all-gathers reproduce the single-process
PDHG step; the scalar partials of a check travel in one allreduce.
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

        import scipy.sparse as sp
        A = sp.csr_matrix((p["vals"], p["col_idx"], p["row_ptr"]), shape=(m, n))
        rp = q["row_ptr"].astype(np.int64)
        Ar = sp.csr_matrix((q["vals"], q["col_idx"], q["row_ptr"]), shape=(q["m"], n))      # my row block
        # block sizes exactly as pdlp.cu: equal column blocks (multiple of 4), rows padded to the largest block
        nb = (-(-n // world) + 3) & ~3
        n0 = min(n, rank * nb)
        nl = max(0, min(n, (rank + 1) * nb) - n0)
        sizes = [None] * world
        td.all_gather_object(sizes, q["m"])
        mb = (max(max(sizes), 1) + 3) & ~3
        # the column block is assembled from every rank's transposed row block (pdlp.cu::build_column_block):
        # each rank sends the slice of its local CSC that falls into my columns, row ids shifted into y_full
        pieces = [None] * world
        loc = Ar.tocsc()
        mine = [(g, loc[:, min(n, g * nb):min(n, (g + 1) * nb)].tocoo()) for g in range(world)]
        td.all_gather_object(pieces, [(g, c.row + rank * mb, c.col, c.data) for g, c in mine])
        rows_, cols_, vals_ = [], [], []
        for src in range(world):                                  # source order = ascending global row
            g, r_, c_, v_ = pieces[src][rank]
            assert g == rank
            rows_.append(r_); cols_.append(c_); vals_.append(v_)
        Ac = sp.coo_matrix((np.concatenate(vals_), (np.concatenate(rows_), np.concatenate(cols_))),
                           shape=(world * mb, nl)).tocsc()          # [padded rows] x [my columns]
        ref = A[:, n0:n0 + nl].tocsc()
        pad_of = np.concatenate([np.arange(cuts[g], cuts[g + 1]) - cuts[g] + g * mb for g in range(world)])
        assert (Ac[pad_of, :] != ref).nnz == 0                    # it IS the column slice of A

        rng = np.random.default_rng(1)                             # same stream on every rank: a common (x, y)
        x = rng.normal(size=n)
        y = rng.normal(size=m)
        tau, sigma = 0.3, 0.2
        lc, uc = row_bounds(p["sense"], p["rhs"])
        # single-process reference step T(z)
        xp1 = np.clip(x - tau * (p["c"] - A.T @ y), p["lb"], p["ub"])
        yp1 = dual_prox(y - sigma * (A @ (2 * xp1 - x)), sigma, lc, uc)

        def gather(block, width):                                 # in-place all-gather of equal-sized blocks
            buf = torch.zeros(world * width, dtype=torch.float64)
            buf[rank * width:rank * width + block.size] = torch.from_numpy(block)
            outs = [torch.zeros(width, dtype=torch.float64) for _ in range(world)]
            td.all_gather(outs, buf[rank * width:(rank + 1) * width].clone())
            return torch.cat(outs).numpy()

        y_full = gather(y[r0:r1], mb)                              # padded row layout
        # K1 on my column block: g = A' y over ALL rows, primal update of my x block, x-bar block published
        gcol = Ac.T @ y_full
        xp_blk = np.clip(x[n0:n0 + nl] - tau * (p["c"][n0:n0 + nl] - gcol), p["lb"][n0:n0 + nl], p["ub"][n0:n0 + nl])
        assert np.allclose(xp_blk, xp1[n0:n0 + nl], rtol=1e-12, atol=1e-12)
        xbar_full = gather(2 * xp_blk - x[n0:n0 + nl], nb)[:n]      # flat column layout
        # K2 on my row block
        yp_blk = dual_prox(y[r0:r1] - sigma * (Ar @ xbar_full), sigma, lc[r0:r1], uc[r0:r1])
        assert np.allclose(yp_blk, yp1[r0:r1], rtol=1e-12, atol=1e-12)
        # scalar partials of a check: row-side from the row block, column-side from the column block, ONE allreduce
        part = torch.tensor([float(np.sum((Ar @ xbar_full) ** 2)), float(np.sum(gcol ** 2))], dtype=torch.float64)
        td.all_reduce(part, op=td.ReduceOp.SUM)
        assert np.isclose(part[0].item(), np.sum((A @ (2 * xp1 - x)) ** 2), rtol=1e-12)
        assert np.isclose(part[1].item(), np.sum((A.T @ y) ** 2), rtol=1e-12)

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
