"""Randomised discrete-event model of the barrier protocol of the EXPERIMENTAL one-CTA attention variant
(freefine_b200/csrc/attn_onecta.cuh; not in the product build, never run on a GPU in round 1).  The model restates the
roles exactly as the kernel codes them -- TMA producer (8-stage K/V ring), MMA issuer (QK four tiles ahead, PV(t) then
QK(t+4), commits to kv_empty / o_done / s_full), two softmax warpgroups (tile parity, two S/P buffers each, o_done
observed every tile before the lazy rescale and at the end of a pass) -- with mbarrier PARITY semantics (a wait on
parity P passes iff the phase in progress has the other parity, so a skipped phase aliases) and an in-order asynchronous
tensor pipe, runs them under random interleavings and checks: no deadlock, every wait that passes has really seen its
event, S/P buffers and O accumulators are never written while a reader is outstanding."""
import random

import pytest

NSTAGE = 8


def sbuf(it):
    return 2 * (it & 1) + ((it >> 1) & 1)


class Bar:
    def __init__(self, count=1):
        self.count, self.pending, self.completed = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.pending = self.count
            self.completed += 1

    def passes(self, parity):
        return (self.completed & 1) != parity


class Sim:
    def __init__(self, passes, seed):
        self.rng = random.Random(seed)
        self.passes = passes                      # tiles per pass
        self.n_total = sum(passes)
        self.bar_s = [Bar() for _ in range(4)]
        self.bar_p = [Bar() for _ in range(4)]   # (one arrival per warpgroup in the model)
        self.bar_o = [Bar() for _ in range(2)]
        self.kv_full = [Bar() for _ in range(NSTAGE)]
        self.kv_empty = [Bar() for _ in range(NSTAGE)]
        self.pipe = []                            # in-order tensor pipe: ("qk", t) / ("pv", t) / ("commit", bar)
        self.buf = [("empty", -1)] * 4            # S/P buffer state
        self.stage_tile = [-1] * NSTAGE           # which tile's K/V a ring stage holds
        self.pv_done = [0, 0]                     # PVs executed per warpgroup
        self.pv_issued = [0, 0]
        self.o_readers = 0
        self.log = []

    # ---- roles as generators: yield ("wait", bar, parity, check) blocks until the wait passes
    def producer(self):
        for it in range(self.n_total):
            stage, use = it % NSTAGE, it // NSTAGE
            if use > 0:
                yield ("wait", self.kv_empty[stage], (use - 1) & 1, lambda it=it: self._stage_free(it))
            self.stage_tile[stage] = it
            self.kv_full[stage].arrive()          # (TMA completion, modelled as immediate)

    def _stage_free(self, it):
        # the tile whose K/V the stage held: its PV must have executed
        assert self._pv_executed(it - NSTAGE), ("K/V stage reused early", it)

    def _pv_executed(self, t):
        return self.pv_done[t & 1] >= (t >> 1) + 1

    def issuer(self):
        qk_t, q_stage, q_phase, p_stage = 0, 0, 0, 0
        bounds = []
        acc = 0
        for n in self.passes:
            bounds.append(acc)
            acc += n

        def issue_qk():
            nonlocal qk_t, q_stage, q_phase
            assert self.stage_tile[q_stage] == qk_t, ("QK reads a stage that holds another tile", qk_t)
            self.pipe.append(("qk", qk_t))
            self.pipe.append(("commit", self.bar_s[sbuf(qk_t)]))
            qk_t += 1
            q_stage += 1
            if q_stage == NSTAGE:
                q_stage, q_phase = 0, q_phase ^ 1

        for _ in range(4):
            if qk_t < self.n_total:
                yield ("wait", self.kv_full[q_stage], q_phase, None)
                issue_qk()
        for t in range(self.n_total):
            yield ("wait", self.bar_p[sbuf(t)], (t >> 2) & 1, lambda t=t: self._p_ready(t))
            pstart = max(b for b in bounds if b <= t)
            self.pipe.append(("pv", t, t - pstart < 2))
            self.pv_issued[t & 1] += 1
            self.pipe.append(("commit", self.kv_empty[p_stage]))
            self.pipe.append(("commit", self.bar_o[t & 1]))
            if qk_t < self.n_total:
                yield ("wait", self.kv_full[q_stage], q_phase, None)
                issue_qk()
            p_stage = (p_stage + 1) % NSTAGE

    def _p_ready(self, t):
        assert self.buf[sbuf(t)] == ("p", t), ("p_full passed before P was written", t, self.buf[sbuf(t)])

    def softmax(self, wg):
        it, m_glob, o_seen = 0, 0, 0
        for n in self.passes:
            it0, n_mine = it, 0
            for itj in range(it, it + n):
                if (itj & 1) != wg:
                    continue
                sb = sbuf(itj)
                yield ("wait", self.bar_s[sb], (itj >> 2) & 1, lambda itj=itj, sb=sb: self._s_ready(itj, sb))
                self.buf[sb] = ("read", itj)      # row in registers
                if o_seen < m_glob:
                    yield ("wait", self.bar_o[wg], (m_glob - 1) & 1, lambda m=m_glob, wg=wg: self._o_done(wg, m))
                    o_seen = m_glob
                # lazy rescale: read-modify-write of O_wg -- no PV of this warpgroup may be outstanding
                assert self.pv_done[wg] == self.pv_issued[wg] == m_glob, ("O rescaled under a running PV", wg, itj)
                self.buf[sb] = ("p", itj)
                self.bar_p[sb].arrive()
                n_mine += 1
                m_glob += 1
            it += n
            if n == 0:
                continue
            if o_seen < m_glob:
                yield ("wait", self.bar_o[wg], (m_glob - 1) & 1, lambda m=m_glob, wg=wg: self._o_done(wg, m))
                o_seen = m_glob
            yield ("named", 1)
            # merge: reads BOTH accumulators -- every PV of this pass must have executed
            for g in (0, 1):
                mine = sum(1 for k in range(it0, it) if (k & 1) == g) + sum(
                    1 for k in range(0, it0) if (k & 1) == g)
                assert self.pv_done[g] == mine, ("merge read an accumulator under a running PV", g, it)
            self.o_readers += 1
            yield ("nop",)                        # (the reads take time)
            self.o_readers -= 1
            yield ("named", 2)

    def _s_ready(self, itj, sb):
        assert self.buf[sb] == ("s", itj), ("s_full passed before S was written", itj, self.buf[sb])

    def _o_done(self, wg, m):
        assert self.pv_done[wg] >= m, ("o_done passed before PV completed (parity aliasing?)", wg, m, self.pv_done[wg])

    # ---- tensor pipe: executes queued ops in order, at arbitrary times
    def pipe_step(self):
        op = self.pipe.pop(0)
        if op[0] == "qk":
            t = op[1]
            sb = sbuf(t)
            assert self.buf[sb][0] == "empty", ("QK overwrote a live S/P buffer", t, self.buf[sb])
            self.buf[sb] = ("s", t)
        elif op[0] == "pv":
            t, first = op[1], op[2]
            sb = sbuf(t)
            assert self.buf[sb] == ("p", t), ("PV read a buffer that does not hold its P", t, self.buf[sb])
            if first:
                assert self.o_readers == 0, ("PV restarted an accumulator while the merge was reading it", t)
            self.buf[sb] = ("empty", -1)
            self.pv_done[t & 1] += 1
        else:
            op[1].arrive()

    def run(self):
        roles = {"prod": self.producer(), "mma": self.issuer(), "sm0": self.softmax(0), "sm1": self.softmax(1)}
        blocked = {k: None for k in roles}
        steps = 0
        while roles or self.pipe:
            steps += 1
            assert steps < 200000, "livelock"
            ready = []
            for k in roles:
                b = blocked[k]
                if b is None or b[0] == "nop" or (b[0] == "wait" and b[1].passes(b[2])):
                    ready.append(k)
            choices = ready + (["pipe"] if self.pipe else [])
            if not choices:
                raise AssertionError(f"deadlock: blocked={ {k: (v[0], v[2] if v[0] == 'wait' else v[1]) for k, v in blocked.items() if v} }")
            c = self.rng.choice(choices)
            if c == "pipe":
                self.pipe_step()
                continue
            b = blocked[c]
            if b is not None and b[0] == "wait" and b[3] is not None:
                b[3]()                              # the wait has passed: its event must really have happened
            try:
                nxt = next(roles[c])
            except StopIteration:
                del roles[c]
                blocked.pop(c)
                continue
            blocked[c] = nxt
            if nxt[0] == "named":                   # both softmax warpgroups must arrive (they run the same pass list)
                waiting = [k for k in ("sm0", "sm1") if blocked.get(k) and blocked[k][0] == "named" and blocked[k][1] == nxt[1]]
                if len(waiting) == 2:
                    for k in waiting:
                        blocked[k] = None
        assert self.pv_done[0] + self.pv_done[1] == self.n_total
        assert all(b[0] == "empty" for b in self.buf)


@pytest.mark.parametrize("passes", [[1], [2], [3], [4], [5], [9], [64], [5, 3], [1, 1, 1, 1], [2, 2, 7], [8, 1, 16, 3],
                                    [0, 5], [5, 0, 4], [17, 17], [33, 2, 1]])
def test_onecta_protocol_random_interleavings(passes):
    for seed in range(40):
        Sim(passes, seed).run()
