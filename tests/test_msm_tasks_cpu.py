"""The task / slot algebra of the segmented level-1 accumulation (csrc/msm.cu: msm_accum_l1_seg_kernel, scan_input mode 2)
restated in Python and checked exhaustively on small cases: thread t walks the aligned window [L t, L t + L) of
the sorted entry list, emits one partial per bucket the window meets into slot pbase[b] + (t - off[b] // L), and

  * every bucket's slots are exactly [pbase[b], pbase[b + 1]) and together hold the bucket's entries once each, in order;
  * the number of partials is at most ceil(total / L) + (buckets - 1), the workspace bound of plan_levels;
  * a bucket of cnt entries gets at most ceil(cnt / L) + 1 partials, which is why plan_levels adds one to the worst case.

This pins the ALGORITHM the kernel implements; the kernel itself is compared with the default path byte for byte on the device
(tests/gpu_msm_variants.py)."""
import random


def run_tasks(counts, L):
    nb = len(counts)
    off = [0]
    for c in counts:
        off.append(off[-1] + c)
    total = off[-1]
    windows = [0 if counts[b] == 0 else (off[b + 1] - 1) // L - off[b] // L + 1 for b in range(nb)]      # scan_input, mode 2
    pbase = [0]
    for w in windows:
        pbase.append(pbase[-1] + w)
    slots = {}

    def emit(b, t, acc):
        s = pbase[b] + (t - off[b] // L)
        assert s not in slots
        slots[s] = (b, acc)

    for t in range((total + L - 1) // L):
        e0, e1 = t * L, min(t * L + L, total)
        b = max(i for i in range(nb) if off[i] <= e0)               # find_segment: the largest b with off[b] <= e0
        assert off[b + 1] > e0
        nxt, acc = off[b + 1], []
        for e in range(e0, e1):
            if e == nxt:
                emit(b, t, acc)
                acc = []
                b += 1
                while off[b + 1] <= e:
                    b += 1
                nxt = off[b + 1]
            acc.append(e)
        emit(b, t, acc)
    return off, pbase, windows, slots


def test_segmented_level1_tasks():
    rnd = random.Random(1)
    for trial in range(400):
        nb, L = rnd.randrange(1, 40), rnd.choice([4, 16, 64])
        kind = trial % 4
        if kind == 0:
            counts = [rnd.randrange(0, 3 * L) for _ in range(nb)]
        elif kind == 1:
            counts = [rnd.choice([0, 0, 0, 1, L, 2 * L, L - 1, L + 1]) for _ in range(nb)]
        elif kind == 2:
            counts = [0] * nb
            counts[rnd.randrange(nb)] = rnd.randrange(1, 40 * L)        # one bucket holds everything (TinyRAM's 0/1 columns)
        else:
            counts = [rnd.randrange(0, L // 2 + 1) for _ in range(nb)]      # many buckets per window (wide windows)
        if not sum(counts):
            continue
        off, pbase, windows, slots = run_tasks(counts, L)
        for b in range(nb):
            got = []
            for s in range(pbase[b], pbase[b + 1]):
                bb, acc = slots[s]
                assert bb == b and acc
                got += acc
            assert got == list(range(off[b], off[b + 1]))
            assert windows[b] <= -(-counts[b] // L) + 1
        assert len(slots) == pbase[-1] <= -(-off[-1] // L) + sum(1 for c in counts if c) - 1 + 1
