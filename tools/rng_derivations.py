#!/usr/bin/env python
"""Offline research tool: enumerate short DERIVATION PROGRAMS for the pcg32 state of cell (14, 14) of the
README grid -- sequences of up to DEPTH steps, each either an engine advance or `state op f(k)` with op in {xor, mul, add,
sub}, f in {identity, hash32 unsigned, hash32 signed (sign-extended)} and k in a small set of constants that could enter a
derivation (seed 123, flat index 434, row / column 14, grid sizes 30 / 900, their +-1) -- from a few start states, and test
the resulting stream against the GIF's 398-bit fork stream with the top-bit and low-bit draw rules, plain and complemented,
with 0..3 leading draws skipped.  Covers e.g. two-level splits (rows, then columns), seeds mixed with indices, re-seeding per
cell.  numpy, a few minutes.  python tools/rng_derivations.py [depth]"""
import itertools
import sys

import numpy as np

TAPE = ("0111111001011111111101100111110101100100011010011001100101000011101010011110001011000000001100011100"
        "1000101011010101101101110111010000101101101001001000010011010001100101110011100110110110110100000101"
        "0110101100110001000010001100011011010011100000111011100011110001001011011000110010000111101010010011"
        "00111100111110010100111010010101111101001010111000001101101001111110111101011100100000111011000001")
MULT = np.uint64(6364136223846793005)
U = np.uint64


def hash_u(x):
    x = np.uint32(x & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        x = ((x >> np.uint32(16)) ^ x) * np.uint32(0x45d9f3b)
        x = ((x >> np.uint32(16)) ^ x) * np.uint32(0x45d9f3b)
        return int((x >> np.uint32(16)) ^ x)


def hash_s(x):   # i32 arithmetic with arithmetic right shifts, as Futhark's >> on i32
    def wrap(v):
        v &= 0xFFFFFFFF
        return v - (1 << 32) if v >= (1 << 31) else v
    x = wrap(x)
    x = wrap(((x >> 16) ^ x) * 0x45d9f3b)
    x = wrap(((x >> 16) ^ x) * 0x45d9f3b)
    return wrap((x >> 16) ^ x)


def operands(small=False):
    ks = sorted({434, 14, 13, 15, 30, 435} if small else {123, 434, 433, 435, 14, 13, 15, 30, 29, 900, 899, 0, 1, 2})
    out = {}
    for k in ks:
        out[f"{k}"] = k & (2**64 - 1)
        out[f"hu({k})"] = hash_u(k)
        hs = hash_s(k)
        out[f"hs({k})"] = hs & (2**64 - 1)
    vals = {}
    for name, v in out.items():
        vals.setdefault(v, name)
    return [(n, v) for v, n in vals.items()]


def main():
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    from_master = len(sys.argv) > 2 and sys.argv[2] == "master"     # start from rng_from_seed [123] (and variants) instead of 0
    tape = np.array([int(c) for c in TAPE], np.uint8)
    ops = operands(small=from_master)
    incs = [((0xda3e39cb94b95bdb << 1) | 1) & (2**64 - 1), 0xda3e39cb94b95bdb | 1]
    steps = [("adv", None, None)] + [(o, n, v) for o in ("xor", "mul", "add", "sub") for n, v in ops]
    print(f"{len(steps)} step kinds, depth <= {depth}: {sum(len(steps) ** d for d in range(1, depth + 1)) * 4:.3g} programs x 16 readings")
    hits = 0
    for inc in incs:
        incv = U(inc | 1)
        def adv(x):
            return (x * int(MULT) + (inc | 1)) & (2**64 - 1)
        starts = (("0", 0), ("initstate", 0x853c49e6748fea9b))
        if from_master:
            c = adv(0)
            starts = (("rng_from_seed[123]", adv((c + 123) & (2**64 - 1))), ("adv(adv(0)^123)", adv(c ^ 123)),
                      ("adv(0)+123", (c + 123) & (2**64 - 1)), ("adv(123)", adv(123)), ("adv(adv(123))", adv(adv(123))))
        for start_name, start in starts:
            # frontier of states after d steps, as numpy arrays, with program ids
            states = np.array([start], np.uint64)
            progs = [()]
            for d in range(depth):
                new_states, new_progs = [], []
                with np.errstate(over="ignore"):
                    for si, (o, n, v) in enumerate(steps):
                        if o == "adv":
                            ns = states * MULT + incv
                        elif o == "xor":
                            ns = states ^ U(v)
                        elif o == "mul":
                            ns = states * U(v)
                        elif o == "add":
                            ns = states + U(v)
                        else:
                            ns = states - U(v)
                        new_states.append(ns)
                        new_progs.append(si)
                states = np.concatenate(new_states)
                nprev = len(progs)
                progs = [p + (si,) for si in new_progs for p in progs]
                assert len(progs) == len(states) and nprev
                # test every state: 44 outputs, skip 0..3, 4 readings
                s = states.copy()
                top = np.zeros((44, len(s)), np.uint8)
                low = np.zeros((44, len(s)), np.uint8)
                with np.errstate(over="ignore"):
                    for t in range(44):
                        old = s
                        xs = (((old >> U(18)) ^ old) >> U(27)).astype(np.uint32)
                        rot = (old >> U(59)).astype(np.uint32)
                        x = (xs >> rot) | (xs << ((np.uint32(0) - rot) & np.uint32(31)))
                        top[t] = x >= np.uint32(0x7FFFFFFF)
                        low[t] = x & np.uint32(1)
                        s = old * MULT + incv
                for skip in range(4):
                    want = tape[:40][:, None]
                    for name, bits in (("top", top), ("low", low)):
                        got = bits[skip:skip + 40]
                        for comp in (0, 1):
                            ok = np.all((got ^ comp) == want, axis=0)
                            for i in np.flatnonzero(ok):
                                hits += 1
                                prog = " ; ".join(steps[k][0] + ("" if steps[k][1] is None else " " + steps[k][1]) for k in progs[i])
                                print(f"HIT inc={inc:#x} start={start_name} program [{prog}] skip {skip} rule {name} complement {comp}")
                if len(states) > 3e7:
                    break
    print("hits:", hits)


if __name__ == "__main__":
    main()
