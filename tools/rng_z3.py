#!/usr/bin/env python
"""Ask an SMT solver whether ANY pcg32 start state reproduces the README-GIF fork stream.

Model: standard PCG-XSH-RR 64/32 (fa_bits.h pcg32_next), the draw rule of uniform_int_distribution (0,1), i.e.
choice = x / 0x7FFFFFFF (x >= 0xFFFFFFFE is redrawn: probability 2^-31 per draw, ignored here and checked on the model),
unknown 64-bit start state; `inc` fixed to cpprandom's constant (default) or free (--free-inc).  The first N tape bits
are constrained (default 128); a model, if found, is replayed against all 398 bits in plain Python.

  SAT   => engine + distribution are right; only the rng_from_seed / split_rng derivation is unknown.
  UNSAT => the engine or the distribution restatement is wrong (or the GIF predates the current plumbing).

This is an offline research tool (not product, not test): `python tools/rng_z3.py --bits 128 --timeout 3600`.
"""
import argparse
import sys
import time

import z3

TAPE = ("0111111001011111111101100111110101100100011010011001100101000011101010011110001011000000001100011100"
        "1000101011010101101101110111010000101101101001001000010011010001100101110011100110110110110100000101"
        "0110101100110001000010001100011011010011100000111011100011110001001011011000110010000111101010010011"
        "00111100111110010100111010010101111101001010111000001101101001111110111101011100100000111011000001")
MULT = 6364136223846793005
INC = ((0xda3e39cb94b95bdb << 1) | 1) & (2**64 - 1)
M64 = 2**64 - 1


def pcg_out(old):
    xs = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
    rot = old >> 59
    return ((xs >> rot) | (xs << ((-rot) & 31))) & 0xFFFFFFFF


def replay(state, inc, n, complement=False, out_new=False):
    s = state
    bits = []
    for _ in range(n):
        while True:
            old = s
            s = (old * MULT + (inc | 1)) & M64
            x = pcg_out(s if out_new else old)
            if x < 0xFFFFFFFE:
                break
        b = x // 0x7FFFFFFF
        bits.append(str(b ^ 1 if complement else b))
    return "".join(bits)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bits", type=int, default=128)
    ap.add_argument("--offset", type=int, default=0, help="tape alignment: constrain tape[offset:]")
    ap.add_argument("--timeout", type=int, default=3600)
    ap.add_argument("--free-inc", action="store_true")
    ap.add_argument("--complement", action="store_true", help="choice 0 <-> 1")
    ap.add_argument("--out-new", action="store_true", help="output computed from the NEW state")
    a = ap.parse_args()
    tape = TAPE[a.offset:]
    n = min(a.bits, len(tape))
    s0 = z3.BitVec("s0", 64)
    inc = z3.BitVec("inc", 64) if a.free_inc else z3.BitVecVal(INC, 64)
    mult = z3.BitVecVal(MULT, 64)
    sol = z3.SolverFor("QF_BV")
    sol.set("timeout", a.timeout * 1000)
    s = s0
    for t in range(n):
        old = s
        s = old * mult + (inc | 1)
        src = s if a.out_new else old
        xs = z3.Extract(31, 0, z3.LShR(z3.LShR(src, 18) ^ src, 27))
        rot = z3.ZeroExt(27, z3.Extract(63, 59, src))
        x = z3.RotateRight(xs, rot)
        want = int(tape[t]) ^ (1 if a.complement else 0)
        # x / 0x7FFFFFFF == want, x < 0xFFFFFFFE
        if want:
            sol.add(z3.UGE(x, z3.BitVecVal(0x7FFFFFFF, 32)), z3.ULT(x, z3.BitVecVal(0xFFFFFFFE, 32)))
        else:
            sol.add(z3.ULT(x, z3.BitVecVal(0x7FFFFFFF, 32)))
    t0 = time.time()
    r = sol.check()
    dt = time.time() - t0
    print(f"bits={n} offset={a.offset} free_inc={a.free_inc} complement={a.complement} out_new={a.out_new}: {r} in {dt:.0f} s",
          flush=True)
    if r == z3.sat:
        m = sol.model()
        st = m[s0].as_long()
        ic = m[inc].as_long() if a.free_inc else INC
        got = replay(st, ic, len(tape), a.complement, a.out_new)
        agree = sum(x == y for x, y in zip(got, tape))
        print(f"state=0x{st:016x} inc=0x{ic:016x}: {agree}/{len(tape)} tape bits reproduced", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
