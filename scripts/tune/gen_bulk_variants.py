#!/usr/bin/env python3
"""Writes variants.inc for the bulk-copy staged kernel: per N and mode the candidates the dispatch table
(pick_bulk, lub_launch.cuh) is chosen from -- search kind (N <= 16), lean step, block shape."""
def cdiv(a, b): return (a + b - 1) // b
def cfg(n, es=4):
    epv = 16 // es
    ch = epv if n % epv == 0 else (2 if (epv == 4 and n % 2 == 0) else 1)
    cpr = n // ch
    budget = 64 if es == 4 else 36
    g = 1
    while g <= 32:
        if g != 2:
            best = None
            gr = 1
            while gr <= g:
                gc = g // gr
                lr, lc = cdiv(n, gr), cdiv(cpr, gc) * ch
                if lr * lc <= budget:
                    cost = 64 * ((lr if gc > 1 else 0) + (lc if gr > 1 else 0)) + lr
                    if best is None or cost < best[0]: best = (cost, gr, gc)
                gr *= 2
            if best: return best[1], best[2]
        g *= 2
    return 4, 8
def minb(n, piv):
    if n <= 6: return 4
    if n == 7 or 9 <= n <= 11: return 3
    if n == 12 and piv: return 3
    return 2
out = []
for n in (6, 7, 9, 10, 11, 13, 14, 15):
    gr, gc = cfg(n)
    for mode in (1, 2):
        for opt in (0, 16):
            out.append("VARB(float, %d, %d, %d, %d, %d, %d, 256)," % (n, gr, gc, mode, minb(n, True), opt))
for n in (17, 18, 19, 21, 22, 23):
    gr, gc = cfg(n)
    for mode in (0, 1, 2):
        for lean in (0, 1):
            out.append("VARB(float, %d, %d, %d, %d, 2, %d, 256)," % (n, gr, gc, mode, lean))
            out.append("VARB(float, %d, %d, %d, %d, 1, %d, 384)," % (n, gr, gc, mode, lean))
open(__file__.rsplit("/", 1)[0] + "/variants.inc", "w").write("\n".join(out) + "\n")
print(len(out), "variants")
