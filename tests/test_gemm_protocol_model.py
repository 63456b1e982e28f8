"""Executable model of the tcgen05 GEMM kernel's synchronisation protocol (csrc/gemm_tc.cu), single-CTA
and CTA-pair (cta_group::2) forms, run under thousands of random schedules.

What is modelled, role by role and loop by loop as in the kernel: the TMA producer (one per CTA), the MMA
issuer (the leader CTA's in the pair form), the epilogue warps (kEpiWarps per CTA), the shared-memory ring
(S stages) with its full / empty mbarriers, the two TMEM accumulator stages with their full / empty
mbarriers, chunked accumulation, asynchronous completion of TMA loads and of tcgen05.commit arrivals.
mbarriers follow the PTX semantics: a phase completes when the pending-arrival count AND the transaction
count reach zero; `try_wait.parity(P)` succeeds once the phase of parity P has completed.

What is checked under every schedule: no deadlock; no barrier ever receives more arrivals than it was
initialised for; an MMA only reads shared-memory stages that hold the operands of exactly its (tile, K step)
in EVERY CTA it reads from; the producer never overwrites a stage an MMA in flight is still reading; an
accumulator is only overwritten after every epilogue warp of every CTA has drained it; the epilogue always
drains the (tile, chunk) it expects.

The pair form has not run on hardware yet.  This model checks its barrier arithmetic — init
counts, who arrives where, phase bookkeeping — not the hardware semantics of the instructions themselves."""
import random

import pytest

K_EPI_WARPS = 8
ACC_STAGES = 2


class MBar:
    def __init__(self, count):
        self.init, self.pending, self.tx, self.phase = count, count, 0, 0

    def _settle(self):
        assert self.pending >= 0, "more arrivals than the barrier was initialised for"
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1
            self.pending = self.init

    def arrive(self):
        self.pending -= 1
        self._settle()

    def arrive_expect_tx(self, nbytes):
        self.tx += nbytes
        self.pending -= 1
        self._settle()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._settle()

    def passed(self, parity):
        return self.phase != parity


class Cta:
    def __init__(self, stages, pair, leader):
        self.full = [MBar(2 if pair else 1) for _ in range(stages)]
        self.empty = [MBar(1) for _ in range(stages)]
        self.tfull = [MBar(1) for _ in range(ACC_STAGES)]
        self.tempty = [MBar((2 if pair else 1) * K_EPI_WARPS) for _ in range(ACC_STAGES)]
        self.smem = [None] * stages          # (tile, it) the stage currently holds
        self.reading = [0] * stages          # MMAs in flight that read the stage
        self.tmem = [None] * ACC_STAGES      # (tile, chunk) the accumulator stage holds
        self.drained = [K_EPI_WARPS] * ACC_STAGES  # epilogue warps that have drained the current content


def simulate(pair, stages, tiles_per_cta, k_iters, chunk_iters, seed):
    rng = random.Random(seed)
    ncta = 2 if pair else 1
    ctas = [Cta(stages, pair, r == 0) for r in range(ncta)]
    leader = ctas[0]
    n_chunks = (k_iters + chunk_iters - 1) // chunk_iters
    stage_bytes = 100
    async_events = []    # independent completions (TMA): any order
    mma_fifo = []        # tcgen05 work of the issuing thread completes in order

    def producer(rank):
        me = ctas[rank]
        s, ph = 0, 0
        for tile in range(tiles_per_cta):
            for it in range(k_iters):
                yield lambda: me.empty[s].passed(ph ^ 1)
                assert me.reading[s] == 0, "producer overwrites a stage an MMA is still reading"
                me.smem[s] = None
                if rank == 0:
                    leader.full[s].arrive_expect_tx(ncta * stage_bytes)

                def landed(me=me, s=s, tile=tile, it=it):
                    me.smem[s] = (tile, it)
                    leader.full[s].complete_tx(stage_bytes)   # pair: bytes are counted on the leader's barrier
                async_events.append(landed)
                if pair and rank != 0:
                    leader.full[s].arrive()                   # mbar_arrive_remote(&full_bar[s], 0)
                s += 1
                if s == stages:
                    s, ph = 0, ph ^ 1

    def mma_issuer():
        s, ph, acc, acc_ph = 0, 0, 0, 0
        for tile in range(tiles_per_cta):
            it = 0
            for c in range(n_chunks):
                yield lambda: leader.tempty[acc].passed(acc_ph ^ 1)
                for cta in ctas:
                    assert cta.drained[acc] == K_EPI_WARPS, "accumulator overwritten before it was drained"
                it_end = min(it + chunk_iters, k_iters)
                while it < it_end:
                    yield lambda: leader.full[s].passed(ph)
                    for cta in ctas:
                        assert cta.smem[s] == (tile, it), f"MMA reads stage {s} holding {cta.smem[s]}, wants {(tile, it)}"
                        cta.reading[s] += 1

                    def mma_done(s=s):
                        for cta in ctas:
                            cta.reading[s] -= 1
                            cta.empty[s].arrive()             # commit (pair: multicast to both CTAs)
                    mma_fifo.append(mma_done)
                    it += 1
                    s += 1
                    if s == stages:
                        s, ph = 0, ph ^ 1

                def chunk_done(acc=acc, tile=tile, c=c):
                    for cta in ctas:
                        cta.tmem[acc] = (tile, c)
                        cta.drained[acc] = 0
                        cta.tfull[acc].arrive()               # commit (pair: multicast)
                mma_fifo.append(chunk_done)
                acc += 1
                if acc == ACC_STAGES:
                    acc, acc_ph = 0, acc_ph ^ 1

    def epilogue_warp(rank):
        me = ctas[rank]
        acc, acc_ph = 0, 0
        for tile in range(tiles_per_cta):
            for c in range(n_chunks):
                yield lambda: me.tfull[acc].passed(acc_ph)
                assert me.tmem[acc] == (tile, c), f"epilogue drains {me.tmem[acc]}, expects {(tile, c)}"
                me.drained[acc] += 1
                leader.tempty[acc].arrive()                   # pair: remote arrive on the leader's barrier
                acc += 1
                if acc == ACC_STAGES:
                    acc, acc_ph = 0, acc_ph ^ 1

    threads = [producer(r) for r in range(ncta)] + [mma_issuer()]
    threads += [epilogue_warp(r) for r in range(ncta) for _ in range(K_EPI_WARPS)]
    waiting = {}
    for i, t in enumerate(threads):
        try:
            waiting[i] = next(t)
        except StopIteration:
            pass
    steps = 0
    while waiting or async_events or mma_fifo:
        steps += 1
        assert steps < 10_000_000
        ready = [i for i, cond in waiting.items() if cond()]
        choices = [("t", i) for i in ready] + [("a", j) for j in range(len(async_events))]
        if mma_fifo:
            choices.append(("m", 0))
        assert choices, "deadlock: every role is blocked on a barrier and nothing is in flight"
        kind, idx = rng.choice(choices)
        if kind == "t":
            try:
                waiting[idx] = next(threads[idx])
            except StopIteration:
                del waiting[idx]
        elif kind == "a":
            async_events.pop(idx)()
        else:
            mma_fifo.pop(0)()
    for cta in ctas:
        assert all(r == 0 for r in cta.reading)
        assert all(d == K_EPI_WARPS for d in cta.drained)
    return steps


@pytest.mark.parametrize("pair", [False, True], ids=["single_cta", "cta_pair"])
@pytest.mark.parametrize("stages,tiles,k_iters,chunk", [(4, 3, 13, 16),    # one chunk per tile (int8 products)
                                                         (4, 4, 26, 4),     # fp16-split: several chunks per tile
                                                         (4, 2, 3, 1),      # fewer K steps than stages
                                                         (4, 5, 1, 16),     # a single K step per tile
                                                         (6, 3, 8, 3),      # ring depth not a divisor of the K loop
                                                         (4, 1, 64, 16)])
def test_gemm_pipeline_protocol_has_no_deadlock_or_hazard(pair, stages, tiles, k_iters, chunk):
    for seed in range(60):
        simulate(pair, stages, tiles, k_iters, chunk, seed)


def test_model_detects_a_wrong_barrier_count():
    """The model is not vacuous: the single-CTA barrier counts applied to the pair form deadlock or over-arrive."""
    global K_EPI_WARPS
    real = Cta.__init__

    def wrong(self, stages, pair, leader):
        real(self, stages, False, leader)     # full: 1 instead of 2, tempty: 8 instead of 16
    Cta.__init__ = wrong
    try:
        with pytest.raises(AssertionError):
            for seed in range(20):
                simulate(True, 4, 3, 8, 4, seed)
    finally:
        Cta.__init__ = real
