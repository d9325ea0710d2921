"""Discrete-event model of the mbarrier / bulk-group protocols of the GEMM kernels (CPU).

The warp-specialised kernels in csrc/gemm.cu synchronise a TMA producer lane, an MMA issuer lane and the epilogue
threads through mbarriers with hand-computed phase parities, tcgen05.commit arrivals, named barriers and
cp.async.bulk wait_group.read counts. A wrong parity or arrival count shows up on hardware as a hang or as a silent
overwrite of a tile that is still in use. This file transcribes the loops of

    v1  gemm_async_epi_kernel            (validated on B200)
    v2  gemm_async_epi2_kernel           (BF16: double-buffered staging; GELU_BWD: in-place aux ring)
    2g  gemm_async_gelu2g_kernel         (two epilogue groups on alternate accumulator buffers)
    sk  gemm_async_smallk_kernel         (weight tile resident per column block: b_full / b_empty)

into Python coroutines with the same variable names and runs them under a randomised scheduler with asynchronous
completion of TMA loads, MMAs and bulk stores. Checked: every role terminates (no deadlock) and no shared-memory /
TMEM buffer is written while a previous user still reads it. The model covers the protocol, not the arithmetic.
"""
import random

import pytest


class Deadlock(Exception):
    pass


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self):
        assert self.pending > 0, "more arrivals than the barrier expects in one phase"
        self.pending -= 1
        self._check()

    def expect_tx(self, nbytes):  # mbarrier.arrive.expect_tx
        self.tx += nbytes
        self.arrive()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._check()

    def test(self, parity):  # mbarrier.try_wait.parity: has the phase with this parity completed?
        return (self.phase & 1) != parity


class Buf:
    """shadow state of a buffer: who may touch it"""

    def __init__(self, name):
        self.name, self.state, self.readers = name, "free", 0

    def begin_write(self, who):
        assert self.state == "free" and self.readers == 0, f"{who} writes {self.name} while it is {self.state}/{self.readers}"
        self.state = "writing"

    def end_write(self):
        self.state = "full"

    def begin_read(self, who):
        assert self.state == "full", f"{who} reads {self.name} while it is {self.state}"
        self.readers += 1

    def end_read(self):
        self.readers -= 1

    def release(self):  # all readers done, contents dead
        assert self.readers == 0, f"{self.name} released with readers"
        self.state = "free"

    def rewrite_in_place(self, who):  # generic-proxy rewrite of a full buffer by its readers (v2 GELU_BWD)
        assert self.state == "full", f"{who} rewrites {self.name} while it is {self.state}"


class Sim:
    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.roles, self.events, self.tick = [], [], 0
        self.mma_queue = []      # in-order tensor pipe: callables run when the MMA / commit "completes"
        self.named = {}          # named barrier id -> [arrived set, generation]

    def spawn(self, name, gen):
        self.roles.append([name, gen, None])  # pending wait condition

    def later(self, fn, lo=1, hi=40):
        self.events.append([self.tick + self.rng.randint(lo, hi), fn])

    def run(self, max_ticks=2_000_000):
        while self.roles or self.events or self.mma_queue:
            self.tick += 1
            if self.tick > max_ticks:
                raise Deadlock("tick limit")
            progressed = False
            due = [e for e in self.events if e[0] <= self.tick]
            for e in due:
                self.events.remove(e)
                e[1]()
                progressed = True
            if self.mma_queue and self.rng.random() < 0.5:
                self.mma_queue.pop(0)()
                progressed = True
            self.rng.shuffle(self.roles)
            for role in list(self.roles):
                name, gen, cond = role
                if cond is not None and not cond():
                    continue
                role[2] = None
                try:
                    role[2] = next(gen)  # a role yields the condition it waits for (or None)
                except StopIteration:
                    self.roles.remove(role)
                progressed = True
                if self.rng.random() < 0.5:
                    break
            if not progressed and not self.events and not self.mma_queue:
                blocked = [r[0] for r in self.roles]
                raise Deadlock(f"no runnable role, blocked: {blocked}")

    # named barrier (bar.sync id, count): returns a wait condition
    def bar_sync(self, bid, count, who):
        st = self.named.setdefault(bid, [set(), 0])
        gen = st[1]
        st[0].add(who)
        if len(st[0]) == count:
            st[0] = set()
            st[1] += 1
        return lambda: self.named[bid][1] > gen


class BulkGroups:
    """cp.async.bulk commit groups of one issuing thread; a store 'reads' its shared-memory source asynchronously, in order"""

    def __init__(self, sim):
        self.sim, self.unread = sim, 0

    def store(self, bufs, on_read=None):
        self.unread += 1
        for b in bufs:
            b.begin_read("TMA store")

        def done():
            for b in bufs:
                b.end_read()
                b.release()
            self.unread -= 1
            if on_read:
                on_read()
        # in-order completion: chain behind the previous store
        prev = getattr(self, "_last_due", 0)
        due = max(prev, self.sim.tick) + self.sim.rng.randint(5, 120)
        self._last_due = due
        self.sim.events.append([due, done])

    def wait_read(self, n):
        return lambda: self.unread <= n


NUM_EPI = 4  # epilogue "threads" in the model (stands for the 256 of the kernel; arrival counts scale accordingly)


def build_kernel(sim, variant, tiles_m, n_tiles, kblocks, num_stages, t_begin=0):
    """variant in {'v1', 'v2_bf16', 'v2_gelu', 'v2_gelu_bwd', 'v1_gelu_bwd', '2g', 'sk_bf16'}"""
    t_end = t_begin + n_tiles
    has_aux = variant in ("v1_gelu_bwd", "v2_gelu_bwd")
    two_group = variant == "2g"
    smallk = variant.startswith("sk")
    full_bar = [MBar(1) for _ in range(num_stages)]
    empty_bar = [MBar(1) for _ in range(num_stages)]
    tmem_full = [MBar(1) for _ in range(2)]
    tmem_empty = [MBar(NUM_EPI // 2 if two_group else NUM_EPI) for _ in range(2)]
    aux_full = [MBar(1) for _ in range(2)]
    aux_empty = [MBar(1 if variant == "v2_gelu_bwd" else NUM_EPI) for _ in range(2)]
    b_full, b_empty = MBar(1), MBar(1)
    stage = [Buf(f"stage{s}") for s in range(num_stages)]
    tmem = [Buf(f"tmem{b}") for b in range(2)]
    aux = [Buf(f"aux{b}") for b in range(2)]
    bres = Buf("B")
    n_out = 2 if variant in ("v2_bf16", "2g", "sk_bf16") else 1
    out = [Buf(f"out{i}") for i in range(n_out)]

    def wait(bar, parity):
        return lambda: bar.test(parity)

    def read_and_arrive(buf_obj, bar, who):
        """a consumer thread reads the buffer and arrives on its 'empty' barrier; the arrival that completes the phase
        is the point from which the other side may overwrite the buffer"""
        buf_obj.begin_read(who)
        buf_obj.end_read()
        before = bar.phase
        bar.arrive()
        if bar.phase != before:
            buf_obj.release()

    def producer():
        it, lt, cur_cb, b_use = 0, 0, -1, 0
        for t in range(t_begin, t_end):
            cb = t // tiles_m
            if smallk and cb != cur_cb:
                yield wait(b_empty, (b_use & 1) ^ 1)
                bres.begin_write("producer")
                b_full.expect_tx(1)
                sim.later(lambda: (bres.end_write(), b_full.complete_tx(1)))
                cur_cb, b_use = cb, b_use + 1
            if has_aux:
                a_s = lt & 1
                yield wait(aux_empty[a_s], ((lt >> 1) & 1) ^ 1)
                aux[a_s].begin_write("producer")
                aux_full[a_s].expect_tx(1)
                sim.later(lambda a_s=a_s: (aux[a_s].end_write(), aux_full[a_s].complete_tx(1)))
            for _kb in range(kblocks):
                s, ph = it % num_stages, (it // num_stages) & 1
                yield wait(empty_bar[s], ph ^ 1)
                stage[s].begin_write("producer")
                full_bar[s].expect_tx(1)
                sim.later(lambda s=s: (stage[s].end_write(), full_bar[s].complete_tx(1)))
                it += 1
            lt += 1

    def mma():
        it, lt, cur_cb, b_use = 0, 0, -1, 0
        for t in range(t_begin, t_end):
            cb = t // tiles_m
            if smallk and cb != cur_cb:
                yield wait(b_full, b_use & 1)
                cur_cb, b_use = cb, b_use + 1
            buf = lt & 1
            yield wait(tmem_empty[buf], ((lt >> 1) & 1) ^ 1)
            tmem[buf].begin_write("MMA")
            for _i in range(kblocks):
                s, ph = it % num_stages, (it // num_stages) & 1
                yield wait(full_bar[s], ph)
                stage[s].begin_read("MMA")
                if smallk:
                    bres.begin_read("MMA")

                def mma_done(s=s):
                    stage[s].end_read()
                    stage[s].release()
                    if smallk:
                        bres.end_read()
                sim.mma_queue.append(mma_done)
                sim.mma_queue.append(lambda s=s: empty_bar[s].arrive())       # umma_commit(&empty_bar[s])
                it += 1
            sim.mma_queue.append(lambda buf=buf: (tmem[buf].end_write(), tmem_full[buf].arrive()))  # commit(tmem_full)
            if smallk and ((t + 1 >= t_end) or ((t + 1) // tiles_m != cb)):
                sim.mma_queue.append(lambda: (bres.release(), b_empty.arrive()))                      # commit(b_empty)
            lt += 1

    def epilogue(et, grp=0, group_size=NUM_EPI, bulk=None):
        issuer = (et == 0)
        bar1, bar2 = (1 + 2 * grp, 2 + 2 * grp)
        start = grp if two_group else 0
        step = 2 if two_group else 1
        lt = start
        while t_begin + lt < t_end:
            buf = lt & 1
            yield wait(tmem_full[buf], (lt >> 1) & 1)
            if variant in ("v1", "v1_gelu_bwd"):
                # v1: tcgen05.ld issued, then the staging wait, then the arithmetic
                if issuer:
                    yield bulk.wait_read(0)
                yield sim.bar_sync(bar1, group_size, (grp, et))
                read_and_arrive(tmem[buf], tmem_empty[buf], f"epi{et}")
                if has_aux:
                    a_s = lt & 1
                    yield wait(aux_full[a_s], (lt >> 1) & 1)
                    read_and_arrive(aux[a_s], aux_empty[a_s], f"epi{et}")
                ob = out[0]
            elif variant in ("v2_bf16", "v2_gelu", "v2_gelu_bwd", "sk_bf16", "bug_wait1_single_buffer"):
                read_and_arrive(tmem[buf], tmem_empty[buf], f"epi{et}")
                if variant == "v2_gelu_bwd":
                    if issuer and lt > 0:
                        yield bulk.wait_read(0)          # the in-place result of the previous tile has been stored
                        aux_empty[buf ^ 1].arrive()
                    yield wait(aux_full[buf], (lt >> 1) & 1)
                    aux[buf].rewrite_in_place(f"epi{et}")
                    ob = aux[buf]
                else:
                    ob = out[buf] if variant in ("v2_bf16", "sk_bf16") and n_out == 2 else out[0]
                    if issuer:
                        yield bulk.wait_read(1 if (n_out == 2 or variant.startswith("bug")) else 0)
                    yield sim.bar_sync(bar1, group_size, (grp, et))
            else:  # two-group GELU kernel: one staging tile per group, used for gelu' then gelu
                ob = out[grp]
                if issuer:
                    yield bulk.wait_read(0)
                yield sim.bar_sync(bar1, group_size, (grp, et))
                read_and_arrive(tmem[buf], tmem_empty[buf], f"epi{grp}.{et}")
            # staging writes of this thread (generic proxy) ...
            if variant != "v2_gelu_bwd":
                if et == (0):
                    ob.begin_write(f"epi{grp}.{et}")
            yield sim.bar_sync(bar2, group_size, (grp, et))
            if issuer:
                if variant != "v2_gelu_bwd":
                    ob.end_write()
                    bulk.store([ob])
                else:
                    bulk.store([ob])  # store straight from the aux stage; the stage is released when it has been read
                if two_group:
                    yield bulk.wait_read(0)
            if two_group:
                yield sim.bar_sync(bar1, group_size, (grp, et))
                if issuer:
                    ob.begin_write(f"epi{grp}.{et} (gelu)")
                yield sim.bar_sync(bar2, group_size, (grp, et))
                if issuer:
                    ob.end_write()
                    bulk.store([ob])
            lt += step

    sim.spawn("producer", producer())
    sim.spawn("mma", mma())
    if two_group:
        for grp in range(2):
            bulk = BulkGroups(sim)
            for et in range(NUM_EPI // 2):
                sim.spawn(f"epi{grp}.{et}", epilogue(et, grp, NUM_EPI // 2, bulk))
    else:
        bulk = BulkGroups(sim)
        for et in range(NUM_EPI):
            sim.spawn(f"epi{et}", epilogue(et, 0, NUM_EPI, bulk))


@pytest.mark.parametrize("variant,kblocks,stages", [
    ("v1", 2, 3), ("v1", 6, 3), ("v1_gelu_bwd", 2, 2), ("v1_gelu_bwd", 12, 2),
    ("v2_bf16", 2, 3), ("v2_bf16", 12, 3), ("v2_gelu", 2, 3), ("v2_gelu_bwd", 2, 3), ("v2_gelu_bwd", 6, 3),
    ("2g", 2, 3), ("2g", 6, 3), ("sk_bf16", 2, 3), ("sk_bf16", 1, 2),
])
@pytest.mark.parametrize("n_tiles,tiles_m,t_begin", [(1, 4, 0), (2, 4, 3), (7, 3, 1), (11, 4, 2)])
def test_protocol_terminates_without_hazards(variant, kblocks, stages, n_tiles, tiles_m, t_begin):
    for seed in range(12):
        sim = Sim(seed * 7919 + n_tiles)
        build_kernel(sim, variant, tiles_m, n_tiles, kblocks, stages, t_begin)
        sim.run()


def test_model_detects_a_missing_staging_wait():
    """wait_group.read 1 is only correct with two staging tiles: with one, the next tile is staged while the store reads"""
    caught = 0
    for seed in range(20):
        sim = Sim(seed)
        build_kernel(sim, "bug_wait1_single_buffer", 4, 9, 2, 3)
        try:
            sim.run()
        except AssertionError:
            caught += 1
    assert caught >= 15


def test_model_detects_a_wrong_parity():
    """sanity of the model itself: a producer that waits on the wrong phase of the empty barriers overwrites a live stage"""
    caught = 0
    for seed in range(20):
        sim = Sim(seed)
        full, empty = [MBar(1) for _ in range(2)], [MBar(1) for _ in range(2)]
        st = [Buf("s0"), Buf("s1")]

        def producer():
            for it in range(8):
                s, ph = it % 2, (it // 2) & 1
                yield (lambda s=s, ph=ph: empty[s].test(ph))  # BUG on purpose: should be ph ^ 1
                st[s].begin_write("producer")
                full[s].expect_tx(1)
                sim.later(lambda s=s: (st[s].end_write(), full[s].complete_tx(1)))

        def consumer():
            for it in range(8):
                s, ph = it % 2, (it // 2) & 1
                yield (lambda s=s, ph=ph: full[s].test(ph))
                st[s].begin_read("consumer")
                sim.later(lambda s=s: (st[s].end_read(), st[s].release(), empty[s].arrive()))

        sim.spawn("p", producer())
        sim.spawn("c", consumer())
        try:
            sim.run(20000)
        except (Deadlock, AssertionError):
            caught += 1
    assert caught == 20
