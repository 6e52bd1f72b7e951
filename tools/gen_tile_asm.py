#!/usr/bin/env python
"""Generates beamforming-lk_b200/csrc/das_tile_asm.inc: the per-channel accumulate step of das_tile as
hand-scheduled inline PTX, one block per channel covering the four directions of a tile.

Why PTX: the step is "for each direction, jump to the body whose register operands match the direction's
offset inside the shared window".  Written in C++ the compiler re-derives each branch predicate right before
its branch (a 20+ cycle ISETP -> BRA dependency per tree level) and wraps every tree in divergence
bookkeeping (BSSY / BSYNC).  Here the bit predicates of a direction are formed together up front, the
branches are `bra.uni` (the offsets are warp-uniform by construction) and the FFMA2 / FADD2 pairs are
software-pipelined four deep.

usage: python tools/gen_tile_asm.py > beamforming-lk_b200/csrc/das_tile_asm.inc
"""
K = 8


import os
VOTE = os.environ.get('BFLK_GEN_VOTE', '0') == '1'   # measured slower on B200 (extra VOTE + BRA.DIV per direction)
BRX = os.environ.get('BFLK_GEN_BRX', '0') == '1'     # indexed jump instead of the bit tree
DEPTH = 4   # FFMA2 -> FADD2 software-pipeline depth (temporaries); set per variant in gen()


def body(r, D, lines, ind="    "):
    """acc[r][k] = acc[r][k] + fma(ff, d[D+k], w[D+k+1]), k = 0..7 (delay.cpp:24), pipelined DEPTH deep."""
    def fma(t, k):
        lines.append(f"{ind}fma.rn.f32x2 t{t}, ff, %{DOP(D + k)}, %{WOP(D + k + 1)};")

    def add(k, t):
        lines.append(f"{ind}add.rn.f32x2 %{AOP(r, k)}, %{AOP(r, k)}, t{t};")
    for k in range(DEPTH):
        fma(k, k)
    for k in range(K):
        add(k, k % DEPTH)
        if k + DEPTH < K:
            fma(k % DEPTH, k + DEPTH)


def gen(nch):
    global AOP, WOP, DOP, DEPTH
    DEPTH = 8 if nch <= 8 else 4
    nw = 2 * nch           # w[0 .. 2nch-1]; w[0] is never an FMA operand
    nd = 2 * nch - 1       # d[0 .. 2nch-2]
    kmax = 2 * nch - 9
    nbits = max(1, kmax.bit_length())
    # operand numbering: acc (32, "+l"), w[1..nw-1], d[0..nd-1], f0..f3, e1
    AOP = lambda r, k: r * K + k
    WOP = lambda j: 32 + (j - 1)
    DOP = lambda j: 32 + (nw - 1) + j
    FOP = lambda r: 32 + (nw - 1) + nd + r
    EOP = 32 + (nw - 1) + nd + 4
    L = []
    L.append("{")
    dbl = 2 * nbits <= 6               # two predicate sets fit the 7 predicate registers
    sets = ["pa", "pb"] if dbl else ["pa", "pa"]
    L.append(f"    .reg .pred pa<{nbits}>, pb<{nbits}>;")
    L.append("    .reg .b32 x;")
    L.append(f"    .reg .b64 ff, t<{DEPTH}>;")

    def preds(r, ps):
        L.append(f"    // predicates of direction {r}: bits of delta = (e1 >> {6 * r}) & 63")
        for b in range(nbits):
            L.append(f"    and.b32 x, %{EOP}, {1 << (6 * r + b)};")
            L.append(f"    setp.ne.b32 {ps}{b}, x, 0;")
            if VOTE:
                # the offsets are warp-uniform by construction; a vote makes that visible to ptxas, which then
                # uses uniform predicates + BRA.U and drops the BSSY / BSYNC reconvergence bookkeeping
                L.append(f"    vote.sync.any.pred {ps}{b}, {ps}{b}, 0xffffffff;")

    preds(0, sets[0])
    for r in range(4):
        ps = sets[r % 2]
        # the next direction's predicates are formed before this direction's body runs, so its branches never
        # wait on a compare (when both sets fit the predicate file)
        if dbl and r + 1 < 4:
            preds(r + 1, sets[(r + 1) % 2])
        elif not dbl and r > 0:
            preds(r, ps)
        L.append(f"    // ---- direction {r}")
        L.append(f"    mov.b64 ff, {{%{FOP(r)}, %{FOP(r)}}};")
        tree_start = len(L)

        def tree(d0, bit):
            # dispatches among deltas d0 .. d0 + 2^(bit+1) - 1 (clipped to kmax)
            if bit < 0:
                body(r, d0, L)
                L.append(f"    bra.uni J{r};")
                return
            hi = d0 + (1 << bit)
            if hi > kmax:
                tree(d0, bit - 1)
                return
            L.append(f"    @{ps}{bit} bra.uni T{r}_{hi}_{bit};")
            tree(d0, bit - 1)
            L.append(f"T{r}_{hi}_{bit}:")
            tree(hi, bit - 1)
        if BRX:
            # undo the tree: one indexed jump (LDC + BRX) to the case body
            del L[tree_start:]
            labels = ", ".join(f"C{r}_{dd}" for dd in range(kmax + 1))
            L.append(f"    bfe.u32 x, %{EOP}, {6 * r}, 6;")
            L.append(f"    min.u32 x, x, {kmax};")
            L.append(f"TS{r}: .branchtargets {labels};")
            L.append(f"    brx.idx.uni x, TS{r};")
            for dd in range(kmax + 1):
                L.append(f"C{r}_{dd}:")
                body(r, dd, L)
                L.append(f"    bra.uni J{r};")
        else:
            tree(0, nbits - 1)
        L.append(f"J{r}:")
    L.append("}")
    asm = "\n".join(f'        "{ln}\\n"' for ln in L)
    outs = ", ".join(f'"+l"(acc[{r}][{k}])' for r in range(4) for k in range(K))
    ins = ", ".join([f'"l"(w[{j}])' for j in range(1, nw)] + [f'"l"(d[{j}])' for j in range(nd)] +
                    [f'"f"(f{r})' for r in range(4)] + ['"r"(e1)'])
    return f"""// NCH = {nch}: window of {nw} sample pairs, deltas 0..{kmax}
__device__ __forceinline__ void tile_channel_asm(u64 (&acc)[4][{K}], const u64 (&w)[{nw}], const u64 (&d)[{nd}],
                                                 float f0, float f1, float f2, float f3, uint32_t e1) {{
    asm volatile(
{asm}
        : {outs}
        : {ins});
}}
"""


print("// GENERATED by tools/gen_tile_asm.py -- do not edit.  See that script for the why.")
for nch in (6, 8, 10):
    print(gen(nch))
