"""Host-side model of the mbarrier choreography of csrc/attention_tc.cu (the experimental single-kernel attention):
the TMA thread, the MMA thread and the four softmax warps are replayed as coroutines over barriers with the
hardware's phase/parity semantics, under random interleavings, with every shared resource (K buffers, V buffer, the S
tile, the P ring, Q, O) carrying an owner tag that is asserted on each use.  It checks that the protocol written in
the kernel cannot deadlock and never reads a resource before it is produced or overwrites it before it is consumed —
the part of the kernel that cannot be checked by compiling it.  (tcgen05.commit is modelled as an arrival that happens
at a random later scheduling point, in issue order, like the in-order tensor pipe.)

    python tools/attention_protocol_sim.py [trials]
"""
import random
import sys


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, 0, 0

    def arrive(self):
        self.pending += 1
        assert self.pending <= self.count, f"{self.name}: more arrivals than the barrier expects in one phase"
        if self.pending == self.count:
            self.pending = 0
            self.phase += 1

    def passed(self, parity):  # mbarrier.try_wait.parity: true once the phase with this parity has completed
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, n_tiles, nb, rng):
        self.n_tiles, self.nb, self.rng = n_tiles, nb, rng
        B = lambda n, c: Bar(n, c)  # noqa: E731
        self.q_full, self.s_full, self.s_empty = B("q_full", 4), B("s_full", 1), B("s_empty", 4)
        self.k_full, self.k_empty = [B("k_full0", 1), B("k_full1", 1)], [B("k_empty0", 1), B("k_empty1", 1)]
        self.p_full, self.p_empty = [B("p_full0", 4), B("p_full1", 4)], [B("p_empty0", 1), B("p_empty1", 1)]
        self.v_full, self.v_empty, self.o_full, self.o_empty = B("v_full", 1), B("v_empty", 1), B("o_full", 1), B("o_empty", 4)
        # resources: tag of what they hold, and who still has to consume it
        self.kbuf = [None, None]
        self.vbuf = None
        self.S, self.S_readers = None, 0
        self.P, self.P_writers = [None, None], [0, 0]
        self.Q, self.Q_writers = None, 0
        self.O_tile, self.O_readers = None, 0
        self.mma_queue = []  # issued, not yet "completed" tensor-core work: (kind, payload) in order

    # ---- tensor pipe: completes queued items in order at random times
    def pump(self):
        while self.mma_queue and self.rng.random() < 0.6:
            kind, payload = self.mma_queue.pop(0)
            if kind == "commit":
                payload.arrive()
            elif kind == "S":  # the S tile now holds (tile, pass, j)
                tile, ps, j, buf = payload
                assert self.Q == tile, f"S MMA read Q of tile {self.Q}, wanted {tile}"
                assert self.kbuf[buf] == (tile, ps, j), f"S MMA read K buffer {self.kbuf[buf]}, wanted {(tile, ps, j)}"
                assert self.S_readers == 0, "S overwritten while softmax warps still read it"
                self.S, self.S_readers = (tile, ps, j), 4
            elif kind == "Odone":  # the tile's accumulator is final: four warps have to read it before it may be overwritten
                assert self.O_tile == payload and self.O_readers == 0
                self.O_readers = 4
            elif kind == "PV":
                tile, j, cc, slot = payload
                assert self.P[slot] == (tile, j, cc) and self.P_writers[slot] == 0, f"PV read P slot {self.P[slot]}, wanted {(tile, j, cc)}"
                assert self.vbuf == (tile, j), f"PV read V buffer {self.vbuf}, wanted {(tile, j)}"
                if j == 0 and cc == 0:
                    assert self.O_readers == 0, "O overwritten before the previous tile was read out"
                self.O_tile = tile
                self.P[slot] = None  # consumed

    # ---- roles
    def tma(self):
        kc = vc = 0
        for tile in range(self.n_tiles):
            for ps in range(2):
                for j in range(self.nb):
                    buf = kc & 1
                    yield ("wait", self.k_empty[buf], ((kc >> 1) & 1) ^ 1)
                    self.kbuf[buf] = (tile, ps, j)  # (the copy lands before the full barrier completes)
                    self.k_full[buf].arrive()
                    kc += 1
                    if ps == 1:
                        yield ("wait", self.v_empty, (vc & 1) ^ 1)
                        self.vbuf = (tile, j)
                        self.v_full.arrive()
                        vc += 1

    def mma(self):
        kc = vc = sc = pc = 0
        for tcount, tile in enumerate(range(self.n_tiles)):
            yield ("wait", self.q_full, tcount & 1)
            for ps in range(2):
                for j in range(self.nb):
                    buf = kc & 1
                    yield ("wait", self.k_full[buf], (kc >> 1) & 1)
                    yield ("wait", self.s_empty, (sc & 1) ^ 1)
                    self.mma_queue += [("S", (tile, ps, j, buf)), ("commit", self.k_empty[buf]), ("commit", self.s_full)]
                    kc += 1
                    sc += 1
                    if ps == 1:
                        yield ("wait", self.v_full, vc & 1)
                        if j == 0:
                            yield ("wait", self.o_empty, (tcount & 1) ^ 1)
                        for cc in range(4):
                            slot = pc & 1
                            yield ("wait", self.p_full[slot], (pc >> 1) & 1)
                            self.mma_queue += [("PV", (tile, j, cc, slot)), ("commit", self.p_empty[slot])]
                            pc += 1
                        self.mma_queue.append(("commit", self.v_empty))
                        vc += 1
            self.mma_queue += [("Odone", tile), ("commit", self.o_full)]

    def softmax(self, w):
        sc = pc = 0
        for tcount, tile in enumerate(range(self.n_tiles)):
            # Q operand of this tile
            assert not any(k == "S" for k, _ in self.mma_queue), "Q rewritten while S products of the previous tile are in flight"
            if self.Q_writers == 0:
                self.Q_writers = 4
            self.Q_writers -= 1
            if self.Q_writers == 0:
                self.Q = tile
            self.q_full.arrive()
            yield ("step",)
            for ps in range(2):
                for j in range(self.nb):
                    yield ("wait", self.s_full, sc & 1)
                    assert self.S == (tile, ps, j), f"warp {w} read S {self.S}, wanted {(tile, ps, j)}"
                    if ps == 1:
                        for cc in range(4):
                            slot = pc & 1
                            yield ("wait", self.p_empty[slot], ((pc >> 1) & 1) ^ 1)
                            if self.P_writers[slot] == 0:
                                assert self.P[slot] is None, f"P slot {slot} overwritten before it was consumed"
                                self.P_writers[slot] = 4
                                self.P[slot] = (tile, j, cc)
                            assert self.P[slot] == (tile, j, cc)
                            self.P_writers[slot] -= 1
                            self.p_full[slot].arrive()
                            pc += 1
                            yield ("step",)
                    self.S_readers -= 1
                    self.s_empty.arrive()
                    sc += 1
            yield ("wait", self.o_full, tcount & 1)
            assert self.O_tile == tile and self.O_readers > 0, f"warp {w} read O of tile {self.O_tile}, wanted {tile}"
            self.O_readers -= 1
            self.o_empty.arrive()

    def run(self):
        roles = [self.tma(), self.mma()] + [self.softmax(w) for w in range(4)]
        blocked = [None] * len(roles)
        done = [False] * len(roles)
        steps = 0
        while not all(done):
            self.pump()
            order = [i for i in range(len(roles)) if not done[i]]
            self.rng.shuffle(order)
            progressed = False
            for i in order:
                if blocked[i] is not None:
                    _, bar, parity = blocked[i]
                    if not bar.passed(parity):
                        continue
                    blocked[i] = None
                try:
                    ev = next(roles[i])
                except StopIteration:
                    done[i] = True
                    progressed = True
                    continue
                progressed = True
                if ev[0] == "wait" and not ev[1].passed(ev[2]):
                    blocked[i] = ev
                if self.rng.random() < 0.5:
                    break
            if not progressed and not self.mma_queue:
                raise RuntimeError("deadlock: " + ", ".join(f"role {i} waits {b[1].name} parity {b[2]}" for i, b in enumerate(blocked) if b))
            steps += 1
            assert steps < 10_000_000
        while self.mma_queue:
            self.pump()


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    for t in range(trials):
        rng = random.Random(t)
        Sim(n_tiles=rng.choice([1, 2, 3, 5]), nb=rng.choice([1, 2, 3, 8]), rng=rng).run()
    print(f"ok: {trials} random interleavings, no deadlock, no resource hazard")


if __name__ == "__main__":
    main()
