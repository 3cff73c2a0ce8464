#!/usr/bin/env python
"""Decode the scheduling control fields of a `cuobjdump -sass` listing (Volta+ 128-bit encoding).

  cuobjdump -sass -fun <mangled> lib.so | python tools/sass_ctrl.py [--from ADDR --to ADDR]

Per instruction prints: address, stall count, yield, write barrier, read barrier, wait mask, opcode text; at the end the
sum of the static stall counts per opcode for the selected address range (a lower bound on the issue time of one warp
running that range alone).  Bits of the second 64-bit word: [41:44] stall, [45] yield, [46:48] write barrier,
[49:51] read barrier, [52:57] wait mask, [58:61] reuse.
"""
import argparse
import collections
import re
import sys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--from", dest="lo", default=None)
    ap.add_argument("--to", dest="hi", default=None)
    ap.add_argument("--quiet", action="store_true")
    args = ap.parse_args()
    lo = int(args.lo, 16) if args.lo else 0
    hi = int(args.hi, 16) if args.hi else 1 << 60
    ins = []
    cur = None
    for line in sys.stdin:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", line)
        if m:
            cur = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]
            ins.append(cur)
            continue
        m = re.match(r"\s*/\* (0x[0-9a-f]{16}) \*/", line)
        if m and cur is not None and cur[3] is None:
            cur[3] = int(m.group(1), 16)
    stalls = collections.Counter()
    counts = collections.Counter()
    total = 0
    for addr, text, w0, w1 in ins:
        if w1 is None or addr < lo or addr > hi:
            continue
        ctrl = w1 >> 41
        stall = ctrl & 0xF
        yld = (ctrl >> 4) & 1
        wb = (ctrl >> 5) & 7
        rb = (ctrl >> 8) & 7
        wm = (ctrl >> 11) & 0x3F
        op = text.split()[0] if not text.startswith("@") else text.split()[1]
        op = op.split(".")[0]
        stalls[op] += stall
        counts[op] += 1
        total += stall
        if not args.quiet:
            print(f"{addr:06x} st={stall:2d} y={yld} wb={wb if wb != 7 else '-'} rb={rb if rb != 7 else '-'} "
                  f"wait={wm:06b} {text}")
    print(f"# instructions {sum(counts.values())}, static stall cycles {total}")
    for op, c in counts.most_common():
        print(f"#   {op:10s} n={c:5d} stall_sum={stalls[op]:6d} avg={stalls[op] / c:.2f}")


if __name__ == "__main__":
    main()
