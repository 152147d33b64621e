"""Properties of FramesSignal's serial f32 cursor `offset += ds` (frames.rs:195) that DESIGN.md argues from (numpy
only; no device, no oracle).

* the closed form offset0 + k * ds is NOT the chain: truncated indices flip - which is why
  the kernels walk the chain literally;
* inside one binade the chain advances by a constant number of ulps after its first step there, ties to even
  included (the closed form per binade is exact, it just does not pay per frame);
* a checkpoint every 4th value plus <= 3 literal steps reproduces every cursor bit for bit (what the staged kernels
  do), and adding +0.0 is exact (how lanes with fewer steps share the instruction)."""
import numpy as np

F32 = np.float32


def chain(o0, ds, n=256):
    out = np.empty(n + 1, dtype=F32)
    o = F32(o0)
    for k in range(n + 1):
        out[k] = o
        o = F32(o + ds)
    return out


def cases(n, seed):
    rng = np.random.default_rng(seed)
    for t in range(n):
        ds = F32(rng.uniform(0.85, 1.17)) if t % 2 else F32(rng.uniform(0.5, 2.0))
        if t % 7 == 0:
            ds = F32(round(float(ds) * 64) / 64)  # few mantissa bits: every step in the upper binades is a tie
        yield F32(rng.uniform(0.0, 1.0)), ds


def test_closed_form_flips_indices():
    flips = 0
    for o0, ds in cases(300, 1):
        lit = chain(o0, ds)
        k = np.arange(257, dtype=F32)
        closed = (o0 + k * ds).astype(F32)
        flips += int(np.count_nonzero(np.trunc(lit) != np.trunc(closed)))
    assert flips > 0


def test_constant_ulp_increment_inside_a_binade():
    segments = 0
    for o0, ds in cases(3000, 2):
        bits = chain(o0, ds).view(np.uint32).astype(np.int64)
        expo = bits >> 23
        i = 0
        while i < bits.size:
            j = i
            while j + 1 < bits.size and expo[j + 1] == expo[i]:
                j += 1
            if j - i >= 3:
                d = np.diff(bits[i:j + 1])
                assert (d[1:] == d[1]).all(), (float(o0), float(ds), i)
                segments += 1
            i = j + 1
    assert segments > 10000


def test_checkpoint_plus_literal_steps_is_the_chain():
    for o0, ds in cases(200, 3):
        lit = chain(o0, ds, 255)
        ck = lit[0::4]
        for r in range(4):
            o = ck.copy()
            for step in range(3):  # three adds for every lane: ds for the first r of them, +0.0 after that
                o = (o + (ds if step < r else F32(0.0))).astype(F32)
            np.testing.assert_array_equal(o.view(np.uint32), lit[r::4].view(np.uint32))
