"""Model check (CPU) of the synchronisation protocol of csrc/attn_alt.cu — two softmax warpgroups on alternate key tiles.

The kernel cannot be run here (no GPU), so its wait graph is restated as a small discrete-event model: mbarriers with phase
parity exactly as `mbarrier.try_wait.parity` sees them (a wait on parity P passes iff the barrier's current phase has the
other parity — so a waiter that is overtaken by two completions blocks, which is the aliasing hazard to rule out), one agent
per role (warpgroup 0, warpgroup 1, the MMA issuer, the in-order tensor pipe that completes committed work asynchronously),
random interleavings.  Checked for every n_kv in 0..9 over many schedules:
  * no deadlock (every agent terminates),
  * PV(j) is issued only after P(j) was stored and O was rescaled for tile j; the rescale for tile j happens after PV(j-1)
    completed; QK(j+2) overwrites an S buffer only after PV(j) consumed P(j) (in-order pipe) and the owner finished S(j),
  * m(j) is read by the owner of tile j+1 before the slot is overwritten with m(j+2).
The statements mirror the code: wait_S_ready / wait_step_retired parities, p_full / m_full parities (j >> 1) & 1."""
import random

import pytest


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0      # phase = number of completed phases

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def passed(self, parity):                                         # try_wait.parity semantics
        return (self.phase & 1) != parity


def s_parity(t):
    return (t >> 1) & 1


def simulate(n_kv, rescale, rng, owner_epilogue=True):
    """rescale(j) -> bool: does the warp of tile j need the in-line O rescale.  Returns the event log.
    owner_epilogue=True is the kernel (only the owner of the last tile waits for step(n_kv-1), the other warpgroup learns it
    through the named barrier); False is the first draft (both warpgroups wait on that phase by parity)."""
    s_full = [MBar(1), MBar(1)]
    p_full = [MBar(1), MBar(1)]       # 128 arrivals of one warpgroup, modelled as 1
    m_full = [MBar(1), MBar(1)]
    log = []
    pipe = []                          # in-order tensor pipe: list of ("op", payload) / ("commit", bar)
    m_slot = [None, None]

    epi = {"arrived": 0}

    def wg(w):                         # generator: yields a wait predicate or None (a step)
        def wait_completion(bar, completion):      # bare parity wait, as mbarrier.try_wait.parity
            return lambda: s_full[bar].passed(completion & 1)

        for j in range(w, n_kv, 2):
            yield wait_completion(j & 1, j >> 1)                         # wait_S_ready(j)
            log.append(("S_read", j))
            if j > 0:
                yield lambda j=j: m_full[(j - 1) & 1].passed(((j - 1) >> 1) & 1)
                assert m_slot[(j - 1) & 1] == j - 1, f"tile {j} read m slot holding {m_slot[(j - 1) & 1]}"
                log.append(("m_read", j - 1))
            m_slot[j & 1] = j
            m_full[j & 1].arrive()
            yield None
            log.append(("P_store", j))
            if j > 0 and rescale(j):
                yield wait_completion((j - 1) & 1, ((j - 1) >> 1) + 1)   # wait_step_retired(j - 1)
                log.append(("rescale", j))
            yield None
            p_full[j & 1].arrive()
            log.append(("p_arrive", j))
        # epilogue: the owner of the last tile waits for step(n_kv - 1); both meet at the named barrier
        if n_kv > 0 and (owner_epilogue is False or w == ((n_kv - 1) & 1)):
            yield wait_completion((n_kv - 1) & 1, ((n_kv - 1) >> 1) + 1)
        epi["arrived"] += 1
        if owner_epilogue:
            yield lambda: epi["arrived"] == 2                            # bar.sync 1, 256
        if n_kv > 0:
            log.append(("epilogue", w))

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

    def tensor_pipe():                 # completes one queued item per step, in order; runs until everything else is done
        while True:
            if pipe:
                kind, x = pipe.pop(0)
                if kind == "commit":
                    x.arrive()
                else:
                    log.append((kind + "_done", x))
            yield "idle" if not pipe else None

    agents = {"wg0": wg(0), "wg1": wg(1), "mma": mma(), "pipe": tensor_pipe()}
    waiting = {}
    live = set(agents)
    steps = 0
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


def _check(n_kv, log):
    pos = {e: i for i, e in enumerate(log)}
    for j in range(n_kv):
        assert pos[("P_store", j)] < pos[("PV_done", j)]
        assert pos[("p_arrive", j)] < pos[("PV_done", j)]                 # PV(j) issued only after the hand-off
        if ("rescale", j) in pos:
            assert pos[("PV_done", j - 1)] < pos[("rescale", j)] < pos[("PV_done", j)]
        if j + 2 < n_kv:                                                   # S buffer j&1 is rewritten by QK(j+2)
            assert pos[("PV_done", j)] < pos[("QK_done", j + 2)]
            assert pos[("S_read", j)] < pos[("QK_done", j + 2)]
        assert pos[("QK_done", j)] < pos[("S_read", j)]
    if n_kv > 0:
        assert ("epilogue", 0) in pos and ("epilogue", 1) in pos
        assert pos[("PV_done", n_kv - 1)] < min(pos[("epilogue", 0)], pos[("epilogue", 1)])


@pytest.mark.parametrize("n_kv", list(range(0, 10)))
def test_no_deadlock_and_ordering(n_kv):
    rng = random.Random(1234 + n_kv)
    for trial in range(300):
        mode = trial % 3
        rescale = (lambda j: True) if mode == 0 else (lambda j: False) if mode == 1 else (lambda j, r=rng: r.random() < 0.5)
        _check(n_kv, simulate(n_kv, rescale, rng))


def test_bare_parity_waits_are_caught():
    """The first draft let both warpgroups wait for step(n_kv-1) by parity; the warpgroup that does not own the last tile has
    not followed that barrier and is a full parity cycle off (passes before the step retired, or blocks for ever when
    n_kv == 1) — the model has to see that."""
    rng = random.Random(7)
    bad = 0
    for n_kv in (1, 3, 4, 6):
        for trial in range(200):
            try:
                _check(n_kv, simulate(n_kv, lambda j: False, rng, owner_epilogue=False))
            except (AssertionError, KeyError):
                bad += 1
    assert bad > 0

