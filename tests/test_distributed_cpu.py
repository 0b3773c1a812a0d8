"""The multi-rank merge logic on CPU with the gloo backend (world_size 2): range selection from the first global
non-zero call, gathering of per-batch integer histograms, and the global replay order -- checked against a sequential
oracle run over the interleaved batch list. (The device kernels the same logic drives are covered by
tests/test_gpu_parity.py::test_ordered_replay_reproduces_sequential_pdf.)"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

WORLD = 2
N_LOCAL = 3
CALLS = 2
Q = 3
LOG_WORDS = 514


def make_batch(b, q, call):
    rng = np.random.default_rng(1000 * b + 10 * q + call)
    x = (rng.standard_normal(2000 + 100 * q) * (1 + q) + call).astype(np.float32)
    if q == 1 and b == 0:
        x[:] = 0            # quantizer 1: global batch 0 is all zeros -> the range must come from batch 1 (rank 1)
    if q == 2:
        x = np.maximum(x, 0)
    return x


def calls_of(q):
    return 2 if q == 0 else 1          # quantizer 0 is a module used twice per forward


def worker(rank, port, result_queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from aimet_b200.distributed import (_all_gather, choose_first_ranges, first_call_positions,
                                            global_replay_offsets)
        from oracle.bindings import Oracle, OracleTfe
        o = Oracle()
        # ---- exchange 1: (min, max) table of this rank's first local batch (global batch == rank) ----
        table = torch.full((Q, CALLS, 2), float("inf"))
        table[..., 1] = float("-inf")
        for q in range(Q):
            for c in range(calls_of(q)):
                mn, mx = o.get_min_max(make_batch(rank, q, c))
                table[q, c, 0], table[q, c, 1] = mn, mx
        gathered = _all_gather(table, None)
        chosen = choose_first_ranges(gathered)
        first_pos = first_call_positions(gathered)
        # ---- local statistics with the agreed ranges: integer histogram log ----
        states = []
        for q in range(Q):
            s = OracleTfe(o)
            s.init_pdf(float(chosen[q, 0]), float(chosen[q, 1]))
            states.append(s)
        log = torch.zeros((N_LOCAL * CALLS, Q, LOG_WORDS), dtype=torch.int32)
        for i in range(N_LOCAL):
            b = i * WORLD + rank
            for q in range(Q):
                bucket, offset = states[q].bucket_params()
                for c in range(calls_of(q)):
                    x = make_batch(b, q, c)
                    h = o.histogram(x, bucket, offset)
                    log[i * CALLS + c, q, :512] = torch.from_numpy(h.astype(np.int32))
                    log[i * CALLS + c, q, 512] = x.size
        pos = rank * CALLS + torch.arange(CALLS)
        void = pos[:, None] < first_pos[None, :]
        log[:CALLS][void] = 0
        all_logs = _all_gather(log, None)
        offsets = global_replay_offsets(WORLD, N_LOCAL, CALLS, Q)
        # ---- replay on the host in the planned order ----
        flat = all_logs.reshape(-1).numpy().view(np.uint32)
        merged = []
        for q in range(Q):
            s = OracleTfe(o)
            s.init_pdf(float(chosen[q, 0]), float(chosen[q, 1]))
            for off in offsets.tolist():
                e = flat[off + q * LOG_WORDS: off + (q + 1) * LOG_WORDS]
                if e[512] == 0:
                    continue
                s.fold_histogram(e[:512].copy(), int(e[512]))
            merged.append((s.histogram()[1], s.compute(8), s.s.iterations))
        if rank == 0:
            result_queue.put([(m[0].tolist(), m[1], m[2]) for m in merged])
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_merge_equals_sequential_run(oracle):
    from oracle.bindings import OracleTfe
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=worker, args=(r, port, queue)) for r in range(WORLD)]
    for p in procs:
        p.start()
    merged = queue.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # sequential reference: one process sees global batches 0 .. WORLD*N_LOCAL-1 in order
    for q in range(Q):
        s = OracleTfe(oracle)
        for b in range(WORLD * N_LOCAL):
            for c in range(calls_of(q)):
                s.update(make_batch(b, q, c))
        pdf, enc, iters = merged[q]
        assert iters == s.s.iterations, q
        assert np.array_equal(np.array(pdf), s.histogram()[1]), q
        assert tuple(enc) == s.compute(8), q


def test_choose_first_ranges_rules():
    from aimet_b200.distributed import choose_first_ranges, first_call_positions
    inf = float("inf")
    g = torch.tensor([  # [W=2, Q=3, C=2, 2]
        [[[0.0, 0.0], [inf, -inf]], [[-1.0, 2.0], [0.5, 0.7]], [[0.0, 0.0], [0.0, 0.0]]],
        [[[-3.0, 4.0], [inf, -inf]], [[-9.0, 9.0], [inf, -inf]], [[0.0, 0.0], [inf, -inf]]],
    ])
    chosen = choose_first_ranges(g)
    assert chosen.tolist() == [[-3.0, 4.0], [-1.0, 2.0], [0.0, 0.0]]
    assert first_call_positions(g).tolist() == [2, 0, 4]


def test_global_replay_offsets_order():
    from aimet_b200.distributed import global_replay_offsets
    off = global_replay_offsets(world=2, local_batches=2, calls=2, num_quantizers=1) // 514
    # slots are [rank][local batch][call]; global order is batch 0 (rank 0), batch 1 (rank 1), batch 2 (rank 0), ...
    assert off.tolist() == [0, 1, 4, 5, 2, 3, 6, 7]


def test_replay_plan_orders_calls_as_a_single_process_would():
    """Per-call log rows tagged (local batch, record): rank w's local batch i is global batch i * W + w; inside a batch the
    calls of a record keep their row order; padding rows (record -1) are ignored."""
    from aimet_b200.distributed import replay_plan
    # W = 2, R = 5 rows per rank. rank 0: batch 0 -> records 0, 1, 0 ; batch 1 -> record 0 ; one padding row
    #                              rank 1: batch 0 -> records 1, 0    ; batch 1 -> records 0, 1, 1
    meta = torch.tensor([[[0, 0], [0, 1], [0, 0], [1, 0], [-1, -1]],
                         [[0, 1], [0, 0], [1, 0], [1, 1], [1, 1]]], dtype=torch.int32)
    rows, begin = replay_plan(meta, num_records=3)
    assert begin.tolist() == [0, 5, 9, 9]                       # record 2 never called
    # record 0: global batch 0 (rank 0 rows 0, 2), batch 1 (rank 1 row 1 -> 5 + 1), batch 2 (rank 0 row 3), batch 3 (rank 1 row 2)
    assert rows[:5].tolist() == [0, 2, 6, 3, 7]
    # record 1: batch 0 (rank 0 row 1), batch 1 (rank 1 row 0), batch 3 (rank 1 rows 3, 4)
    assert rows[5:9].tolist() == [1, 5, 8, 9]


def plan_worker(rank, port, result_queue):
    """Two ranks log per-call rows (different numbers of calls per batch for quantizer 0), exchange them with ONE gather of
    rows + tags as ShardedCalibrator._merge does, and replay on the host through replay_plan."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from aimet_b200.distributed import _all_gather, choose_first_ranges, first_call_positions, replay_plan
        from oracle.bindings import Oracle, OracleTfe
        o = Oracle()
        table = torch.full((Q, CALLS, 2), float("inf"))
        table[..., 1] = float("-inf")
        for q in range(Q):
            for c in range(calls_of(q)):
                mn, mx = o.get_min_max(make_batch(rank, q, c))
                table[q, c, 0], table[q, c, 1] = mn, mx
        gathered = _all_gather(table, None)
        chosen = choose_first_ranges(gathered)
        first_pos = first_call_positions(gathered).tolist()
        states = []
        for q in range(Q):
            s = OracleTfe(o)
            s.init_pdf(float(chosen[q, 0]), float(chosen[q, 1]))
            states.append(s)
        rows, tags = [], []
        for i in range(N_LOCAL):
            b = i * WORLD + rank
            for q in range(Q):
                bucket, offset = states[q].bucket_params()
                for c in range(calls_of(q)):
                    if i == 0 and rank * CALLS + c < first_pos[q]:
                        continue                       # all-zero call before the range existed: never logged
                    x = make_batch(b, q, c)
                    row = torch.zeros(LOG_WORDS, dtype=torch.int32)
                    row[:512] = torch.from_numpy(o.histogram(x, bucket, offset).astype(np.int32))
                    row[512] = x.size
                    rows.append(row)
                    tags.append((i, q))
        used = torch.tensor([len(rows)])
        dist.all_reduce(used, op=dist.ReduceOp.MAX)
        r_max = int(used)
        payload = torch.zeros((r_max + 1, LOG_WORDS), dtype=torch.int32)
        payload[:len(rows)] = torch.stack(rows)
        t = torch.full((r_max, 2), -1, dtype=torch.int32)
        t[:len(tags)] = torch.tensor(tags, dtype=torch.int32)
        payload[r_max, :2 * r_max] = t.view(-1)
        everything = _all_gather(payload, None)
        meta = everything[:, r_max, :2 * r_max].reshape(WORLD, r_max, 2)
        entry_rows, begin = replay_plan(meta, Q)
        flat = everything.reshape(-1, LOG_WORDS).numpy().view(np.uint32)
        merged = []
        for q in range(Q):
            s = OracleTfe(o)
            s.init_pdf(float(chosen[q, 0]), float(chosen[q, 1]))
            for e in entry_rows[begin[q]:begin[q + 1]].tolist():
                w, r = divmod(e, r_max)
                entry = flat[w * (r_max + 1) + r]
                s.fold_histogram(entry[:512].copy(), int(entry[512]))
            merged.append((s.histogram()[1].tolist(), s.compute(8), s.s.iterations))
        if rank == 0:
            result_queue.put(merged)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_per_call_log_merge_equals_sequential_run(oracle):
    from oracle.bindings import OracleTfe
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=plan_worker, args=(r, port, queue)) for r in range(WORLD)]
    for p in procs:
        p.start()
    merged = queue.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for q in range(Q):
        s = OracleTfe(oracle)
        for b in range(WORLD * N_LOCAL):
            for c in range(calls_of(q)):
                s.update(make_batch(b, q, c))
        pdf, enc, iters = merged[q]
        assert iters == s.s.iterations, q
        assert np.array_equal(np.array(pdf), s.histogram()[1]), q
        assert tuple(enc) == s.compute(8), q
