#!/usr/bin/env python3
"""Generates csrc/mont_chains.cuh: the carry-chain primitives of the Montgomery multiplier.

Every primitive is one PTX carry chain (add.cc / madc.lo.cc / madc.hi.cc ...).  On the
device it is emitted as ONE `asm volatile` block, so the condition-code register never
lives across statements the compiler could reorder; ptxas fuses each
(mad.lo.cc, madc.hi.cc) pair on the same operands into a single IMAD.WIDE.U32(.X).
For host builds (unit tests of the algorithm without a GPU, and the few serial field
operations the host side of the library does) the same op list is emitted as portable
C with an explicit carry variable, so both paths run the identical sequence.

Run:  python gen_chains.py > mont_chains.cuh
"""
import sys

NS = (8, 12)


class Chain:
    def __init__(self, name, params, tparams=None):
        self.name = name
        self.params = params  # C++ parameter list (string)
        self.tparams = tparams
        self.ops = []  # (op, dst, a, b, c)   operand = C++ lvalue/rvalue expression or int literal
        self.out = []  # expressions written
        self.inp = []  # expressions read only
        self.imm = []  # compile-time constants

    def op(self, o, d, a, b, c=None):
        self.ops.append((o, d, a, b, c))

    # ---- emit -------------------------------------------------------------------------
    def _operands(self):
        written, read = [], []
        for o, d, a, b, c in self.ops:
            if d not in written:
                written.append(d)
        for o, d, a, b, c in self.ops:
            for s in (a, b, c):
                if s is None or isinstance(s, int):
                    continue
                if s not in written and s not in read:
                    read.append(s)
        return written, read

    def emit(self):
        written, read = self._operands()
        imms = [s for s in read if s.startswith("F::")]
        regs = [s for s in read if not s.startswith("F::")]
        idx = {}
        for e in written + regs + imms:
            idx[e] = len(idx)

        def ref(s):
            if isinstance(s, int):
                return str(s)
            return "%" + str(idx[s])

        lines = []
        for o, d, a, b, c in self.ops:
            srcs = ", ".join(ref(s) for s in (a, b, c) if s is not None)
            lines.append('"%s.u32 %s, %s;\\n\\t"' % (o, ref(d), srcs))
        def constraint(e):
            # read at or before its first write -> read-write; otherwise a pure, early-clobber output
            for o, d, a, b, c in self.ops:
                if e in (a, b, c):
                    return '"+r"'
                if d == e:
                    return '"=&r"'
            return '"+r"'

        outs = ", ".join('%s(%s)' % (constraint(e), e) for e in written)
        ins = ", ".join(['"r"(%s)' % e for e in regs] + ['"n"(%s)' % e for e in imms])
        dev = "    asm volatile(\n        " + "\n        ".join(lines) + "\n        : %s\n        : %s);\n" % (outs, ins)

        host = ["    uint32_t cc = 0; uint64_t t_;"]
        for o, d, a, b, c in self.ops:
            A = str(a)
            B = str(b)
            base = o.split(".")
            name = base[0]
            use_cc = name.endswith("c") and name not in ("sub",) and name in ("addc", "subc", "madc")
            set_cc = base[-1] == "cc"
            cin = "cc" if use_cc else "0"
            if name in ("add", "addc"):
                host.append("    t_ = (uint64_t)(uint32_t)(%s) + (uint32_t)(%s) + %s; %s = (uint32_t)t_;%s" %
                            (A, B, cin, d, " cc = (uint32_t)(t_ >> 32);" if set_cc else ""))
            elif name in ("sub", "subc"):
                host.append("    t_ = (uint64_t)(uint32_t)(%s) - (uint32_t)(%s) - %s; %s = (uint32_t)t_;%s" %
                            (A, B, cin, d, " cc = (uint32_t)(t_ >> 63);" if set_cc else ""))
            elif name in ("mad", "madc"):
                part = base[1]
                prod = "((uint64_t)(uint32_t)(%s) * (uint32_t)(%s))" % (A, B)
                sel = "(uint32_t)%s" % prod if part == "lo" else "(uint32_t)(%s >> 32)" % prod
                host.append("    t_ = (uint64_t)%s + (uint32_t)(%s) + %s; %s = (uint32_t)t_;%s" %
                            (sel, c, cin, d, " cc = (uint32_t)(t_ >> 32);" if set_cc else ""))
            else:
                raise ValueError(o)
        host.append("    (void)cc; (void)t_;")
        if not self.ops:
            return "__host__ __device__ __forceinline__ void %s(%s) {}\n" % (self.name, self.params)
        t = "template <class F> " if self.tparams else ""
        s = "%s__host__ __device__ __forceinline__ void %s(%s) {\n" % (t, self.name, self.params)
        s += "#ifdef __CUDA_ARCH__\n" + dev + "#else\n" + "\n".join(host) + "\n#endif\n}\n"
        return s


def gen(N):
    out = []
    arr = "uint32_t (&%s)[" + str(N) + "]"
    carr = "const uint32_t (&%s)[" + str(N) + "]"

    # acc += sum_{j even} a[j]*b * 2^(32 j);   ci += carry-out
    for imm in (False, True):
        A = (lambda j: "F::P%d" % j) if imm else (lambda j: "a[%d]" % j)
        c = Chain("chain_mad_even" + ("_p" if imm else ""),
                  (arr % "acc") + ("" if imm else ", " + carr % "a") + ", uint32_t b, uint32_t &ci", imm)
        for j in range(0, N, 2):
            c.op("mad.lo.cc" if j == 0 else "madc.lo.cc", "acc[%d]" % j, A(j), "b", "acc[%d]" % j)
            c.op("madc.hi.cc", "acc[%d]" % (j + 1), A(j), "b", "acc[%d]" % (j + 1))
        c.op("addc", "ci", "ci", 0)
        out.append(c.emit())

        # acc += sum_{j odd} a[j]*b * 2^(32 (j-1));  carry-out is provably zero
        c = Chain("chain_mad_odd" + ("_p" if imm else ""),
                  (arr % "acc") + ("" if imm else ", " + carr % "a") + ", uint32_t b", imm)
        for j in range(1, N, 2):
            c.op("mad.lo.cc" if j == 1 else "madc.lo.cc", "acc[%d]" % (j - 1), A(j), "b", "acc[%d]" % (j - 1))
            c.op("madc.hi.cc" if j != N - 1 else "madc.hi", "acc[%d]" % j, A(j), "b", "acc[%d]" % j)
        out.append(c.emit())

    # e0 += x[1] (carry into the chain);  x = (x >> 64) + sum_{j odd} a[j]*b * 2^(32 (j-1))
    c = Chain("chain_shift_mad_odd", (arr % "x") + ", uint32_t &e0, " + (carr % "a") + ", uint32_t b")
    c.op("add.cc", "e0", "e0", "x[1]")
    for j in range(1, N, 2):
        k = j - 1
        lo_add = "x[%d]" % (k + 2) if k + 2 < N else 0
        hi_add = "x[%d]" % (k + 3) if k + 3 < N else 0
        c.op("madc.lo.cc", "x[%d]" % k, "a[%d]" % j, "b", lo_add)
        c.op("madc.hi.cc" if j != N - 1 else "madc.hi", "x[%d]" % (k + 1), "a[%d]" % j, "b", hi_add)
    out.append(c.emit())

    # Squaring rows (Fp::sqr): row I of a square only takes the products with j >= I (the others were added, doubled,
    # by the rows before it).  Same chains as above with the skipped products replaced by the carry propagation alone
    # (odd chain: the 64-bit shift still has to happen) or dropped (even chain: nothing to add below the first product).
    for I in range(1, N):
        c = Chain("chain_shift_mad_odd_from%d" % I, (arr % "x") + ", uint32_t &e0, " + (carr % "a") + ", uint32_t b")
        c.op("add.cc", "e0", "e0", "x[1]")
        for j in range(1, N, 2):
            k = j - 1
            lo_add = "x[%d]" % (k + 2) if k + 2 < N else 0
            hi_add = "x[%d]" % (k + 3) if k + 3 < N else 0
            if j >= I:
                c.op("madc.lo.cc", "x[%d]" % k, "a[%d]" % j, "b", lo_add)
                c.op("madc.hi.cc" if j != N - 1 else "madc.hi", "x[%d]" % (k + 1), "a[%d]" % j, "b", hi_add)
            else:
                c.op("addc.cc", "x[%d]" % k, lo_add, 0)
                c.op("addc.cc", "x[%d]" % (k + 1), hi_add, 0)
        out.append(c.emit())
        c = Chain("chain_mad_even_from%d" % I, (arr % "acc") + ", " + (carr % "a") + ", uint32_t b, uint32_t &ci")
        first = True
        for j in range(0, N, 2):
            if j < I:
                continue
            c.op("mad.lo.cc" if first else "madc.lo.cc", "acc[%d]" % j, "a[%d]" % j, "b", "acc[%d]" % j)
            c.op("madc.hi.cc", "acc[%d]" % (j + 1), "a[%d]" % j, "b", "acc[%d]" % (j + 1))
            first = False
        if not first:
            c.op("addc", "ci", "ci", 0)
        out.append(c.emit())
    disp = []
    for nm, params, args in (("chain_shift_mad_odd_from", (arr % "x") + ", uint32_t &e0, " + (carr % "a") + ", uint32_t b", "x, e0, a, b"),
                             ("chain_mad_even_from", (arr % "acc") + ", " + (carr % "a") + ", uint32_t b, uint32_t &ci", "acc, a, b, ci")):
        body = " else ".join("if constexpr (I == %d) %s%d(%s);" % (I, nm, I, args) for I in range(1, N))
        disp.append("template <int I> __host__ __device__ __forceinline__ void %s(%s) {\n    %s\n}\n" % (nm, params, body))
    out.append("\n".join(disp))

    # r = (e >> 32) + o
    c = Chain("chain_merge", (arr % "r") + ", " + (carr % "e") + ", " + (carr % "o"))
    for k in range(N):
        opn = "add.cc" if k == 0 else ("addc.cc" if k < N - 1 else "addc")
        c.op(opn, "r[%d]" % k, "e[%d]" % (k + 1) if k + 1 < N else 0, "o[%d]" % k)
    out.append(c.emit())

    # r = a + b, carry -> cout (0/1)
    c = Chain("chain_add", (arr % "r") + ", " + (carr % "a") + ", " + (carr % "b") + ", uint32_t &cout")
    for k in range(N):
        c.op("add.cc" if k == 0 else "addc.cc", "r[%d]" % k, "a[%d]" % k, "b[%d]" % k)
    c.op("addc", "cout", 0, 0)
    out.append(c.emit())

    # r = a - b, borrow -> bout (0 or 0xffffffff)
    c = Chain("chain_sub", (arr % "r") + ", " + (carr % "a") + ", " + (carr % "b") + ", uint32_t &bout")
    for k in range(N):
        c.op("sub.cc" if k == 0 else "subc.cc", "r[%d]" % k, "a[%d]" % k, "b[%d]" % k)
    c.op("subc", "bout", 0, 0)
    out.append(c.emit())

    # r = a - p, borrow -> bout
    c = Chain("chain_sub_p", (arr % "r") + ", " + (carr % "a") + ", uint32_t &bout", True)
    for k in range(N):
        c.op("sub.cc" if k == 0 else "subc.cc", "r[%d]" % k, "a[%d]" % k, "F::P%d" % k)
    c.op("subc", "bout", 0, 0)
    out.append(c.emit())

    # r = a + (p & mask)
    c = Chain("chain_add_masked", (arr % "r") + ", " + (carr % "a") + ", " + (carr % "pm"))
    for k in range(N):
        opn = "add.cc" if k == 0 else ("addc.cc" if k < N - 1 else "addc")
        c.op(opn, "r[%d]" % k, "a[%d]" % k, "pm[%d]" % k)
    out.append(c.emit())
    return "\n".join(out)


def main():
    print("// GENERATED by gen_chains.py -- do not edit.  Carry-chain primitives for 8- and 12-limb")
    print("// (256- / 384-bit) Montgomery arithmetic; see gen_chains.py for the rationale.")
    print("#pragma once")
    print("#include <stdint.h>")
    print()
    for N in NS:
        print("// ---------------------------------------------------------------- N = %d" % N)
        print(gen(N))


if __name__ == "__main__":
    main()
