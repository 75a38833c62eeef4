"""Model check (CPU) of the wait graph of csrc/attn.cu (softmax warpgroup, correction warpgroup, MMA issuer, in-order tensor
pipe) in its product form and with the opt-in -DSAB_DEFER_PST / -DSAB_PREMAX changes, with the same barrier model as
tests/test_alt_protocol_model.py (parity semantics of mbarrier.try_wait, random interleavings): no deadlock for n_kv = 0..9,
PV(j) only after P(j) was stored and the correction warps are done with tile j, the correction of tile j only after PV(j-1)
completed, S buffers rewritten only after their P was consumed and their owner finished reading."""
import random

import pytest

from test_alt_protocol_model import MBar


def simulate(n_kv, rescale, rng, defer=False, premax=False):
    s_full = [MBar(1), MBar(1)]
    p_full = [MBar(2), MBar(2)]       # softmax (128 arrivals -> 1) + correction (128 -> 1)
    a_full = [MBar(1), MBar(1)]
    log, pipe = [], []

    def softmax():
        have_pre = False
        for j in range(n_kv):
            if not have_pre:
                yield lambda j=j: s_full[j & 1].passed((j >> 1) & 1)
            log.append(("S_read", j))
            if defer and j > 0:
                p_full[(j - 1) & 1].arrive(); log.append(("p_arrive", j - 1))
            a_full[j & 1].arrive()
            yield None
            have_pre = False
            if premax and j + 1 < n_kv and s_full[(j + 1) & 1].passed(((j + 1) >> 1) & 1):   # mbarrier.test_wait, non-blocking
                have_pre = True
                log.append(("premax", j + 1))
            log.append(("P_store", j))
            yield None
            if not defer or j == n_kv - 1:
                p_full[j & 1].arrive(); log.append(("p_arrive", j))
        if n_kv > 0:
            yield lambda: s_full[(n_kv + 1) & 1].passed(((n_kv + 1) >> 1) & 1)
            log.append(("epilogue", 0))

    def correction():
        for j in range(n_kv):
            yield lambda j=j: a_full[j & 1].passed((j >> 1) & 1)
            if j > 0 and rescale(j):
                yield lambda j=j: s_full[(j + 1) & 1].passed(((j + 1) >> 1) & 1)
                log.append(("rescale", j))
            p_full[j & 1].arrive()
            log.append(("c_arrive", j))

    def mma():
        if n_kv > 0:
            pipe.append(("QK", 0)); pipe.append(("commit", s_full[0]))
            if n_kv > 1:
                pipe.append(("QK", 1)); pipe.append(("commit", s_full[1]))
            yield None
            for j in range(n_kv):
                yield lambda j=j: p_full[j & 1].passed((j >> 1) & 1)
                pipe.append(("PV", j))
                if j + 2 < n_kv:
                    pipe.append(("QK", j + 2))
                pipe.append(("commit", s_full[j & 1]))
                yield None

    def tensor_pipe():
        while True:
            if pipe:
                kind, x = pipe.pop(0)
                if kind == "commit":
                    x.arrive()
                else:
                    log.append((kind + "_done", x))
            yield None

    agents = {"softmax": softmax(), "correction": correction(), "mma": mma(), "pipe": tensor_pipe()}
    waiting, live, steps = {}, set(agents), 0
    while live - {"pipe"} or pipe:
        steps += 1
        assert steps < 20000, "livelock"
        runnable = [a for a in live if a not in waiting or waiting[a]()]
        if not [a for a in runnable if a != "pipe"] and not pipe:
            raise AssertionError(f"deadlock at n_kv={n_kv}: waiting {sorted(waiting)}; log tail {log[-8:]}")
        a = rng.choice(runnable)
        waiting.pop(a, None)
        try:
            r = next(agents[a])
        except StopIteration:
            live.discard(a)
            continue
        if callable(r):
            waiting[a] = r
    return log


@pytest.mark.parametrize("defer,premax", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("n_kv", list(range(0, 10)))
def test_attn_wait_graph(n_kv, defer, premax):
    rng = random.Random(99 + n_kv)
    for trial in range(200):
        mode = trial % 3
        rescale = (lambda j: True) if mode == 0 else (lambda j: False) if mode == 1 else (lambda j, r=rng: r.random() < 0.5)
        log = simulate(n_kv, rescale, rng, defer, premax)
        pos = {e: i for i, e in enumerate(log)}
        for j in range(n_kv):
            assert pos[("QK_done", j)] < pos[("S_read", j)]
            assert pos[("P_store", j)] < pos[("p_arrive", j)] < pos[("PV_done", j)]
            assert pos[("c_arrive", j)] < pos[("PV_done", j)]
            if ("rescale", j) in pos:
                assert pos[("PV_done", j - 1)] < pos[("rescale", j)] < pos[("PV_done", j)]
            if ("premax", j) in pos:
                assert pos[("QK_done", j)] < pos[("premax", j)]
            if j + 2 < n_kv:
                assert pos[("PV_done", j)] < pos[("QK_done", j + 2)] and pos[("S_read", j)] < pos[("QK_done", j + 2)]
        if n_kv > 0:
            assert pos[("PV_done", n_kv - 1)] < pos[("epilogue", 0)]
