"""Model check (CPU) of the synchronisation protocol of csrc/attn_q4.cu — one CTA per SM, four softmax warpgroups on key tiles
j = w, w+4, ..., four S buffers and four separate P buffers in TMEM, one QK^T issuer and one PV issuer.

Same method as tests/test_alt_protocol_model.py: mbarriers with the phase-parity semantics of `mbarrier.try_wait.parity`
(a waiter that is two completions off passes or blocks wrongly — the aliasing hazard to rule out), one agent per role, random
interleavings.  The statements mirror the kernel: per-buffer barriers, completion c of buffer b belongs to tile 4c + b.
Checked for n_kv in 0..13:
  * no deadlock;
  * S(t) is read only after QK(t) retired, and QK(t+4) overwrites the buffer only after the owner's LAST read of S(t);
  * P(t) is written only after PV(t-4) retired (its buffer), PV(t) is issued only after P(t) was handed off;
  * an in-line O rescale for tile j lies between PV(j-1) and PV(j);
  * m(j-1) is read from its slot before the slot is overwritten with m(j+3);
  * the epilogue of every warpgroup comes after PV(n_kv-1)."""
import random

import pytest


class MBar:
    def __init__(self):
        self.phase = 0                                                  # number of completed phases

    def arrive(self):                                                   # 128 arrivals of one warpgroup / one commit, modelled as 1
        self.phase += 1

    def passed(self, parity):                                           # try_wait.parity semantics
        return (self.phase & 1) != parity


NW = 4


def simulate(n_kv, rescale, rng):
    s_full, s_free, p_full, pv_done, m_full = ([MBar() for _ in range(NW)] for _ in range(5))
    log, pipe = [], []                                                  # in-order tensor pipe shared by the two issuers
    m_slot = [None] * NW
    epi = {"arrived": 0}

    def wg(w):
        for j in range(w, n_kv, NW):
            b, c = j & 3, j >> 2
            yield lambda b=b, c=c: s_full[b].passed(c & 1)
            log.append(("S_read", j))
            if j > 0:
                yield lambda j=j: m_full[(j - 1) & 3].passed(((j - 1) >> 2) & 1)
                assert m_slot[(j - 1) & 3] == j - 1, f"tile {j} read an m slot holding {m_slot[(j - 1) & 3]}"
            m_slot[b] = j
            m_full[b].arrive()
            yield None
            log.append(("S_last_read", j))
            s_free[b].arrive()
            yield None
            if j >= NW:
                yield lambda b=b, c=c: pv_done[b].passed((c - 1) & 1)
            log.append(("P_store", j))
            if j > 0 and rescale(j):
                yield lambda j=j: pv_done[(j - 1) & 3].passed(((j - 1) >> 2) & 1)
                log.append(("rescale", j))
            yield None
            log.append(("p_arrive", j))
            p_full[b].arrive()
        if n_kv > 0 and w == ((n_kv - 1) & 3):
            yield lambda: pv_done[(n_kv - 1) & 3].passed(((n_kv - 1) >> 2) & 1)
        epi["arrived"] += 1
        yield lambda: epi["arrived"] == NW                               # bar.sync 1, 512
        if n_kv > 0:
            log.append(("epilogue", w))

    def qk_issuer():
        for t in range(n_kv):
            if t >= NW:
                yield lambda t=t: s_free[t & 3].passed(((t >> 2) - 1) & 1)
            pipe.append(("QK", t)); pipe.append(("commit", s_full[t & 3]))
            yield None

    def pv_issuer():
        for t in range(n_kv):
            yield lambda t=t: p_full[t & 3].passed((t >> 2) & 1)
            pipe.append(("PV", t)); pipe.append(("commit", pv_done[t & 3]))
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

    agents = {f"wg{w}": wg(w) for w in range(NW)}
    agents.update(qk=qk_issuer(), pv=pv_issuer(), pipe=tensor_pipe())
    waiting, live, steps = {}, set(agents), 0
    while live - {"pipe"} or pipe:
        steps += 1
        assert steps < 100000, "livelock"
        runnable = [a for a in live if (a != "pipe" or pipe) and (a not in waiting or waiting[a]())]
        if not runnable:
            raise AssertionError(f"deadlock at n_kv={n_kv}: waiting {sorted(waiting)}; log tail {log[-8:]}")
        a = rng.choice(sorted(runnable))
        waiting.pop(a, None)
        try:
            r = next(agents[a])
        except StopIteration:
            live.discard(a)
            continue
        if callable(r):
            waiting[a] = r
    return log


def _check(n_kv, log):
    pos = {e: i for i, e in enumerate(log)}
    assert len(pos) == len(log), "an event happened twice"
    for j in range(n_kv):
        assert pos[("QK_done", j)] < pos[("S_read", j)]
        assert pos[("p_arrive", j)] < pos[("PV_done", j)]
        if j + NW < n_kv:
            assert pos[("S_last_read", j)] < pos[("QK_done", j + NW)]
        if j >= NW:
            assert pos[("PV_done", j - NW)] < pos[("P_store", j)]
        if j > 0:
            assert pos[("PV_done", j - 1)] < pos[("PV_done", j)]
        if ("rescale", j) in pos:
            assert pos[("PV_done", j - 1)] < pos[("rescale", j)] < pos[("PV_done", j)]
    if n_kv > 0:
        for w in range(NW):
            assert pos[("PV_done", n_kv - 1)] < pos[("epilogue", w)]


@pytest.mark.parametrize("n_kv", list(range(0, 14)))
def test_no_deadlock_and_ordering(n_kv):
    rng = random.Random(4321 + n_kv)
    for trial in range(200):
        mode = trial % 3
        rescale = (lambda j: True) if mode == 0 else (lambda j: False) if mode == 1 else (lambda j, r=rng: r.random() < 0.5)
        _check(n_kv, simulate(n_kv, rescale, rng))


def test_starved_agents_do_not_alias():
    """Adversarial schedules: one agent at a time is starved (runs only when nothing else can), which is what exposes parity
    aliasing — a waiter overtaken by two completions of the barrier it is about to test."""
    for starved in ["wg0", "wg1", "wg2", "wg3", "qk", "pv", "pipe"]:
        class Biased(random.Random):
            def choice(self, seq):
                rest = [a for a in seq if a != starved]
                return super().choice(rest if rest else seq)
        for n_kv in (5, 8, 9, 13):
            for seed in range(20):
                _check(n_kv, simulate(n_kv, lambda j: True, Biased(seed)))
