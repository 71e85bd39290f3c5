"""Host bookkeeping behind the chain overlap (flmip.cpp: stream_run), driven through the test hook flmip_overlap_bookkeeping -- no GPU.
The invariant it has to keep: the first kernel of a chain may only start without waiting for its predecessor if no kernel that can
still be running touches its image, i.e. if the image is not in the stream's open run."""
import numpy as np

OPT, CHAIN, LITERAL_CHAIN, OTHER, FORGET = 0, 1, 2, 3, 4


class Model:
    """what the rule says, restated: the open run of a stream and whether a chain head may start late"""

    def __init__(self):
        self.enabled, self.run = False, []

    def chain(self, img, kernels, pdl_head=True):
        late = self.enabled and pdl_head and bool(self.run) and len(self.run) < 64 and img not in self.run
        if self.enabled and kernels:
            if not late:
                self.run = []      # only a first kernel that waited at its start opens a new run (later kernels release their dependents first)
            self.run.append(img)
        return late

    def other(self):
        self.run = []


def test_overlap_is_off_by_default_and_per_stream(built_lib):
    L = built_lib
    s1, s2 = 0x1000, 0x2000
    assert L.flmip_overlap_bookkeeping(s1, 1, CHAIN, 1) == 0 and L.flmip_overlap_bookkeeping(s1, 2, CHAIN, 1) == 0  # never opted in
    L.flmip_overlap_bookkeeping(s1, 0, OPT, 1)
    assert L.flmip_overlap_bookkeeping(s1, 1, CHAIN, 1) == 0   # nothing of ours in front: waits (and opens the run)
    assert L.flmip_overlap_bookkeeping(s1, 2, CHAIN, 1) == 1   # another image: starts late
    assert L.flmip_overlap_bookkeeping(s2, 3, CHAIN, 1) == 0   # the other stream has not opted in
    assert L.flmip_overlap_bookkeeping(s1, 1, CHAIN, 1) == 0   # image 1 is still in the open run: waits, new run = {1}
    assert L.flmip_overlap_bookkeeping(s1, 2, CHAIN, 1) == 1   # image 2 is not in the new run
    assert L.flmip_overlap_bookkeeping(s1, 2, CHAIN, 1) == 0   # the same image twice in a row
    L.flmip_overlap_bookkeeping(s1, 0, OTHER, 0)               # a copy / fill / event / fence
    assert L.flmip_overlap_bookkeeping(s1, 5, CHAIN, 1) == 0   # ... closes the run
    assert L.flmip_overlap_bookkeeping(s1, 6, CHAIN, 3) == 1   # a three-kernel chain may start late ...
    assert L.flmip_overlap_bookkeeping(s1, 5, CHAIN, 1) == 0   # ... and its later kernels release their dependents before they wait: image 5 is still in the run
    assert L.flmip_overlap_bookkeeping(s1, 6, CHAIN, 1) == 1   # that chain waited at its start and opened a new run {5}
    assert L.flmip_overlap_bookkeeping(s1, 6, CHAIN, 1) == 0
    assert L.flmip_overlap_bookkeeping(s1, 7, LITERAL_CHAIN, 2) == 0   # a chain the literal kernel starts never skips the wait
    assert L.flmip_overlap_bookkeeping(s1, 8, CHAIN, 1) == 1
    L.flmip_overlap_bookkeeping(s1, 0, OPT, 0)
    assert L.flmip_overlap_bookkeeping(s1, 9, CHAIN, 1) == 0   # opted out again
    for s in (s1, s2):
        L.flmip_overlap_bookkeeping(s, 0, FORGET, 0)


def test_bookkeeping_matches_the_rule_on_random_sequences(built_lib):
    L = built_lib
    rng = np.random.default_rng(7)
    s = 0x3000
    m = Model()
    for step in range(20000):
        r = rng.random()
        if r < 0.02:
            en = int(rng.integers(0, 2))
            L.flmip_overlap_bookkeeping(s, 0, OPT, en)
            m.enabled, m.run = bool(en), []
        elif r < 0.10:
            L.flmip_overlap_bookkeeping(s, 0, OTHER, 0)
            m.other()
        else:
            img = int(rng.integers(1, 100 if step % 3000 < 1500 else 6))   # long runs of distinct images (the 64-image bound) and short ones
            kernels = int(rng.choice([1, 1, 1, 2, 3, 0]))
            pdl = rng.random() > 0.1
            got = L.flmip_overlap_bookkeeping(s, img, CHAIN if pdl else LITERAL_CHAIN, kernels)
            assert got == int(m.chain(img, kernels, pdl)), step
    L.flmip_overlap_bookkeeping(s, 0, FORGET, 0)


def test_a_late_head_never_shares_an_image_with_the_open_run(built_lib):
    """the safety property itself, tracked independently of the model above: every image of the kernels enqueued since the last kernel
    that waited at its start"""
    L = built_lib
    rng = np.random.default_rng(11)
    s = 0x4000
    L.flmip_overlap_bookkeeping(s, 0, OPT, 1)
    maybe_running = set()
    for step in range(20000):
        if rng.random() < 0.05:
            L.flmip_overlap_bookkeeping(s, 0, OTHER, 0)
            maybe_running = set()      # a stream-ordered op: everything in front of it completes before anything behind it starts
            continue
        img = int(rng.integers(1, 12))
        kernels = int(rng.choice([1, 1, 2, 4]))
        late = L.flmip_overlap_bookkeeping(s, img, CHAIN, kernels)
        if late:
            assert img not in maybe_running, step
        if not late:
            maybe_running = set()      # the first kernel of this chain waited for everything in front of it before anything behind it could start
        maybe_running.add(img)
    L.flmip_overlap_bookkeeping(s, 0, FORGET, 0)
