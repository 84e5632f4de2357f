"""Model check of the CNVM kernel's optimistic commit (native_cnvm.cu, `process`): the warp protocol executed in
Python on random tiny instances with an arbitrary serialization order of the atomics, against the sequential
semantics.  Development aid; the GPU tests compare the real kernel with the oracle.
    python tools/commit_model.py [trials]"""
import random
import sys


def sequential(x, lanes, table):
    x = list(x)
    acc_n = 0
    for (noise, a, nbr, newop, r3) in lanes:
        m = x[a]
        nw = newop if noise else x[nbr]
        if r3 < table[noise][m][nw]:
            x[a] = nw
            acc_n += 1
    return x, acc_n


def warp(x, lanes, table, rng):
    x = list(x)
    L = len(lanes)
    pending = [True] * L
    acc_n = 0
    rounds = 0
    while True:
        rounds += 1
        ev = []
        for i, (noise, a, nbr, newop, r3) in enumerate(lanes):
            m = x[a]
            nw = newop if noise else x[nbr]
            acc = pending[i] and r3 < table[noise][m][nw]
            ev.append((m, nw, acc))
        order = [i for i in range(L) if ev[i][2]]
        rng.shuffle(order)  # the hardware serializes the atomics of one instruction in some order
        won = [False] * L
        lost = [False] * L
        for i in order:  # compare-and-swap on the field: the first writer of a field wins, the others write nothing
            a = lanes[i][1]
            if ev[i][0] == ev[i][1]:
                continue
            if x[a] == ev[i][0]:
                x[a] = ev[i][1]
                won[i] = True
            else:
                lost[i] = True
        bad = [False] * L
        for i, (noise, a, nbr, newop, r3) in enumerate(lanes):
            m, nw, acc = ev[i]
            chk = x[a] ^ (nw if (acc and not lost[i]) else m)
            if not noise:
                chk |= x[nbr] ^ nw
            bad[i] = (pending[i] and chk != 0) or lost[i]
        if not any(bad):
            acc_n += sum(1 for e in ev if e[2])
            return x, acc_n, rounds
        for i in range(L):  # roll back
            if won[i]:
                x[lanes[i][1]] ^= ev[i][0] ^ ev[i][1]
        writers = [i for i in range(L) if ev[i][2] and ev[i][0] != ev[i][1]]
        w0 = writers[0] if writers else -1
        f = bad.index(True)
        g = max(f, w0 + 1)
        for i in range(g):
            if ev[i][2]:
                x[lanes[i][1]] ^= ev[i][0] ^ ev[i][1]
                acc_n += 1
            pending[i] = False


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    rng = random.Random(1)
    worst = 0
    for t in range(trials):
        n = rng.choice([2, 3, 4, 6, 10, 40])
        M = rng.choice([2, 3, 4])
        L = rng.choice([3, 5, 8, 32])
        x = [rng.randrange(M) for _ in range(n)]
        table = [[[0 if m == w and rng.random() < 0.9 else rng.choice([0, 30, 60, 100]) for w in range(M)] for m in range(M)] for _ in range(2)]
        lanes = []
        for _ in range(L):
            a = rng.randrange(n)
            nbr = rng.randrange(n)
            lanes.append((rng.random() < 0.3, a, nbr, rng.randrange(M), rng.randrange(100)))
        xs, an = sequential(x, lanes, table)
        xw, aw, rounds = warp(x, lanes, table, rng)
        worst = max(worst, rounds)
        if xs != xw or an != aw:
            print("MISMATCH", t, n, M, L, x, lanes, table, xs, xw, an, aw)
            return 1
    print(f"{trials} instances agree (max rounds {worst})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
