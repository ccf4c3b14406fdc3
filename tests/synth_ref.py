"""Independent numpy restatement of the synthetic road-network rule that the
device generator k_synth implements; used to cross-check the generator and to feed the oracle."""
import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = np.asarray(x, np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def synth_hash(seed, tag, a, b):
    with np.errstate(over="ignore"):
        k = splitmix64(np.uint64(seed) ^ (np.uint64(tag) * np.uint64(0xD6E8FEB86659FD93)))
    k = splitmix64(k ^ np.asarray(a, np.uint64))
    return splitmix64(k ^ np.asarray(b, np.uint64))


def ring_heading(cw, ly, lx):
    """outgoing travel direction of ring cell (ly, lx) in a 13x13 ring"""
    if cw:
        if ly == 0 and lx < 12: return 0, 1
        if lx == 12 and ly < 12: return 1, 0
        if ly == 12 and lx > 0: return 0, -1
        return -1, 0
    if ly == 0 and lx > 0: return 0, -1
    if lx == 0 and ly < 12: return 1, 0
    if ly == 12 and lx < 12: return 0, 1
    return -1, 0


def generate(gh, gw, seed=1, density16=6554, ring_full16=2):
    Y, X = np.meshgrid(np.arange(gh, dtype=np.int64), np.arange(gw, dtype=np.int64), indexing="ij")
    hc = synth_hash(seed, 2, Y, X)
    car_draw = (hc & np.uint64(0xFFFF)) < np.uint64(density16)
    uy = np.zeros((gh, gw), np.int8); ux = np.zeros((gh, gw), np.int8)
    dy = np.zeros((gh, gw), np.int8); dx = np.zeros((gh, gw), np.int8)
    arow = (Y % 16) == 0
    acol = (X % 16) == 8
    ux[arow] = 1
    uy[acol] = 1
    alongx = arow & (~acol | (((hc >> np.uint64(16)) & np.uint64(1)) == 1))
    av = (arow | acol) & car_draw
    dx[av & alongx] = 1
    dy[av & ~alongx] = 1
    rings = []
    nby = (gh - 15) // 16 + 1 if gh >= 15 else 0
    nbx = (gw - 23) // 16 + 1 if gw >= 23 else 0
    for by in range(nby):
        for bx in range(nbx):
            hb = int(synth_hash(seed, 1, by, bx))
            if not (hb & 1):
                continue
            cw = bool((hb >> 1) & 1)
            full = ((hb >> 8) & 15) < ring_full16
            s = 1 if cw else -1
            cells = []
            for ly in range(13):
                for lx in range(13):
                    if not (ly in (0, 12) or lx in (0, 12)):
                        continue
                    y, x = 16 * by + 2 + ly, 16 * bx + 10 + lx
                    if ly == 0: ux[y, x] = s
                    if ly == 12: ux[y, x] = -s
                    if lx == 12: uy[y, x] = s
                    if lx == 0: uy[y, x] = -s
                    hy, hx = ring_heading(cw, ly, lx)
                    if full or car_draw[y, x]:
                        dy[y, x], dx[y, x] = hy, hx
                    cells.append((y, x, hy, hx))
            rings.append(cells)
    occ = (dy != 0) | (dx != 0)
    color = np.full((gh, gw), 0xFFFFFFFF, np.uint32)
    hcol = (synth_hash(seed, 3, Y, X) & np.uint64(0xFFFFFF)).astype(np.uint32) | np.uint32(0xFF000000)
    color[occ] = hcol[occ]
    rng_state = synth_hash(seed, 4, Y, X)
    pref = np.zeros((gh, gw), np.uint8)
    inc = ((0xda3e39cb94b95bdb << 1) | 1) & 0xFFFFFFFFFFFFFFFF
    arrs = dict(uy=uy, ux=ux, dy=dy, dx=dx, color=color, pref=pref, rng_state=rng_state,
                rng_inc=np.full((gh, gw), inc, np.uint64))
    # sparse rings with the grant rule of steppers.fut:206-211 evaluated on the ring itself
    off, idx, cy, cx, grant = [0], [], [], [], []
    for cells in rings:
        d = {(y, x): (hy, hx) for y, x, hy, hx in cells}
        for y, x, hy, hx in cells:
            idx.append(y * gw + x); cy.append(hy); cx.append(hx)
            if uy[y, x] != 0 and ux[y, x] != 0:
                q = d.get((y - int(uy[y, x]), x))
                grant.append(1 if (q is not None and q == (int(uy[y, x]), 0)) else 2)
            else:
                grant.append(1 if uy[y, x] != 0 else 2)
        off.append(len(idx))
    ring_arrays = (np.asarray(off, np.int64), np.asarray(idx, np.int64), np.asarray(cy, np.int8),
                   np.asarray(cx, np.int8), np.asarray(grant, np.uint8))
    return arrs, ring_arrays
