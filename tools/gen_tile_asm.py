#!/usr/bin/env python
"""Generates beamforming-lk_b200/csrc/das_tile_asm.inc: one channel step of das_tile as hand-written PTX.

One asm block per (warp, channel) does: window loads (LDS.128 with the pad correction folded into two base
registers), the differences s[i] - s[i+1] (once per window), then for each of the tile's four directions a
jump to the body whose register operands match the direction's offset inside the window, and finally the
prefetch of the next channel's table entry into the same registers.

Why PTX instead of C++ (measured on B200, profiles/README.md):
 * C++ re-derives every branch predicate right before its branch (a ~20-cycle ISETP -> BRA dependency per
   tree level).  Here the bit predicates of a direction are formed one direction ahead (two predicate sets).
 * `bra.uni` + no join points ("chain"): each case body ends with the NEXT direction's dispatch, so a
   direction costs only its tree branches; bodies are shared, only the small trees are replicated.
 * the table-entry prefetch lands in place (no register moves), the loop around the block is two adds.
The kernel is issue-bound (an FFMA2 / FADD2 holds the issue port two cycles), so every instruction removed
from this block is throughput.

Rejected variants, kept behind flags: BFLK_GEN_VOTE=1 (vote-uniform predicates, -7 %), BFLK_GEN_BRX=1
(brx.idx jump table, -9 .. -22 %), BFLK_GEN_CHAIN=0 (join after every direction).

usage: python tools/gen_tile_asm.py > beamforming-lk_b200/csrc/das_tile_asm.inc
       python tools/gen_tile_asm.py --fast > beamforming-lk_b200/csrc/das_tile_fast_asm.inc
"""
import os

K = 8
VOTE = os.environ.get("BFLK_GEN_VOTE", "0") == "1"
BRX = os.environ.get("BFLK_GEN_BRX", "0") == "1"
CHAIN = os.environ.get("BFLK_GEN_CHAIN", "1") == "1"
# exact-triple bodies: FADD2 / FSUB2 written as FFMA2 with a unit multiplier (same single rounding, one pipe)
ADD_AS_FMA = os.environ.get("BFLK_GEN_ADD_AS_FMA", "0") == "1"
# two-FMA variants: next channel's window loads issued INSIDE the last direction's body, each chunk right after the last
# FFMA2 that reads the registers it lands in (no extra registers, the load latency hides behind the rest of the body)
# Measured: single-window flavour cfg5 0.588 -> 0.595, cfg1 +-0; two-window flavour cfg3 0.560 -> 0.522, cfg2 0.595 -> 0.559
# (shared memory is already 81 % busy there; earlier loads only lengthen its queue) -> "single"
PIPE_MODE = os.environ.get("BFLK_GEN_PIPE", "single")
# two-FMA variants: one loop tail (fraction prefetch, next window addresses, loop branch) shared by all last-direction
# bodies ("1"), or a copy per body ("0"); "dual" = shared for the two-window flavour only.  Round 1 measured the shared
# tail +0.3 .. +0.5 % for the two-window flavour (fewer register moves across the back edge); with the round-2 stage loop
# (last-arriver refill) ptxas splits the live ranges of four accumulators of the second direction around the shared tail
# instead (8 MOVs per channel): a tail per body removes them, cfg3 0.559 -> 0.577, cfg2 0.600 -> 0.620 -> "0"
SHARED_TAIL = os.environ.get("BFLK_GEN_SHARED_TAIL", "0")
# wide flavour (gen_wide): next channel's window loads from inside the last body (same idea as PIPE_MODE)
PIPE_MODE_WIDE = os.environ.get("BFLK_GEN_PIPE_WIDE", "0")
# two-window two-FMA flavour: no join point after the window-B block (measured: see profiles/README.md)
NOJOIN = os.environ.get("BFLK_GEN_NOJOIN", "0") == "1"
# two-FMA variants: fetch the last window chunk (4 of 24 shared-memory wavefronts) only where the table says a direction
# reads it (bits 29 / 30 of the delta word; cfg3: 33.5 instead of 36 window wavefronts per (warp, channel), cfg5: 20 instead
# of 24).  ADAPT_MODE pred = predicated LDS (ptxas keeps the chunk's old registers alive and moves accumulators around the
# window-B block: 135 instead of 82 MOVs), sel = always load, from a warp-uniform address (one broadcast wavefront) when not
# needed (+1 LOP3 +1 SEL per window).  Measured on one B200, same box, sel mode vs off: cfg3 0.5794 vs 0.5768, cfg2 0.6210 vs
# 0.6225, cfg1 0.6485 vs 0.6501, cfg5 0.6046 vs 0.6051 -- the wavefronts saved are worth no more than the two instructions
# added (the loop is bound by issue slots, not by shared-memory wavefronts) -> off
ADAPT = os.environ.get("BFLK_GEN_ADAPT", "0") == "1"
ADAPT_MODE = os.environ.get("BFLK_GEN_ADAPT_MODE", "sel")
ADAPT_PWB_EARLY = os.environ.get("BFLK_GEN_ADAPT_PWB_EARLY", "0") == "1"

# operand numbers of the asm block: acc[4][8] "+l" 0..31, e0 32, e1 33 ("+r"), f0..f3 34..37 ("+f"),
# row 38 ("r": shared address of this lane's row start), nxt 39 ("r": shared address of the next entry)
A = lambda r, k: f"%{r * K + k}"
E0, E1, ROW, NXT = "%32", "%33", "%38", "%39"
F = lambda r: f"%{34 + r}"


def gen(nch):
    nw, nd = 2 * nch, 2 * nch - 1
    kmax = 2 * nch - 9                      # largest delta whose 9-pair window fits the loaded chunks
    nbits = max(1, kmax.bit_length())
    depth = 8 if nch <= 8 else 4            # FFMA2 -> FADD2 software-pipeline depth (temporaries)
    dbl = 2 * nbits <= 6                    # two predicate sets fit the 7 predicate registers
    chain = CHAIN and dbl and not BRX
    sets = ["pa", "pb"] if dbl else ["pa", "pa"]
    L = []
    emit = L.append

    def body(r, D):
        """acc[r][k] = acc[r][k] + fma(f, d[D+k], w[D+k+1]), k = 0..7  (delay.cpp:24), pipelined."""
        fma = lambda t, k: emit(f"    fma.rn.f32x2 t{t}, ff, d{D + k}, w{D + k + 1};")
        if ADD_AS_FMA:   # acc = fma(t, 1, acc): the same single rounding as the add, FFMA2 only on the FP pipe
            add = lambda k, t: emit(f"    fma.rn.f32x2 {A(r, k)}, t{t}, one, {A(r, k)};")
        else:
            add = lambda k, t: emit(f"    add.rn.f32x2 {A(r, k)}, {A(r, k)}, t{t};")
        for k in range(depth):
            fma(k, k)
        for k in range(K):
            add(k, k % depth)
            if k + depth < K:
                fma(k % depth, k + depth)

    def preds(r, ps):
        for b in range(nbits):
            emit(f"    and.b32 x, {E1}, {1 << (6 * r + b)};")
            emit(f"    setp.ne.b32 {ps}{b}, x, 0;")
            if VOTE:
                emit(f"    vote.sync.any.pred {ps}{b}, {ps}{b}, 0xffffffff;")

    def setf(r):
        emit(f"    mov.b64 ff, {{{F(r)}, {F(r)}}};")

    def prefetch():
        # e0, e1, f0..f3 are dead from here on: fetch the next channel's entry into them
        emit(f"    ld.shared.v2.u32 {{{E0}, {E1}}}, [{NXT}];")
        emit(f"    ld.shared.v4.f32 {{{F(0)}, {F(1)}, {F(2)}, {F(3)}}}, [{NXT}+16];")

    emit("{")
    emit(f"    .reg .pred pa<{nbits}>, pb<{nbits}>, q<4>;")
    emit("    .reg .b32 x, rr, a<4>;")
    emit(f"    .reg .b64 ff, t<{depth}>, w<{nw}>, d<{nd}>;")
    if ADD_AS_FMA:
        emit("    .reg .b64 one, mone;")
        emit("    mov.b64 one, 0x3F8000003F800000;")
        emit("    mov.b64 mone, 0xBF800000BF800000;")
    # ---- window: chunk m sits at padded chunk m + ((r + m) >> 2), r = bits 24-25 of e1 -----------------
    emit(f"    add.u32 a0, {ROW}, {E0};")
    emit(f"    shr.u32 rr, {E1}, 24;")
    emit("    add.u32 x, a0, 16;")
    for cls, thr in ((1, 3), (2, 2), (3, 1)):   # chunk class m & 3 needs the +16 when r >= 4 - (m & 3)
        emit(f"    setp.ge.u32 q{cls}, rr, {thr};")
        emit(f"    selp.u32 a{cls}, x, a0, q{cls};")
    for m in range(nch):
        emit(f"    ld.shared.v2.b64 {{w{2 * m}, w{2 * m + 1}}}, [a{m & 3}+{16 * (m + (m >> 2))}];")
    if chain:
        preds(0, sets[0])
        preds(1, sets[1])
    else:
        preds(0, sets[0])
    for j in range(nd):
        if ADD_AS_FMA:
            emit(f"    fma.rn.f32x2 d{j}, w{j + 1}, mone, w{j};")   # s[i] - s[i+1] = fma(s[i+1], -1, s[i])
        else:
            emit(f"    sub.rn.f32x2 d{j}, w{j}, w{j + 1};")       # s[i] - s[i+1], once per window

    if chain:
        uid = [0]

        def jump_tree(r, ps, d0, bit):
            if bit < 0:
                emit(f"    bra.uni B{r}_{d0};")
                return
            hi = d0 + (1 << bit)
            if hi > kmax:
                jump_tree(r, ps, d0, bit - 1)
                return
            uid[0] += 1
            lab = f"N{uid[0]}"
            emit(f"    @{ps}{bit} bra.uni {lab};")
            jump_tree(r, ps, d0, bit - 1)
            emit(f"{lab}:")
            jump_tree(r, ps, hi, bit - 1)

        setf(0)
        jump_tree(0, sets[0], 0, nbits - 1)
        for r in range(4):
            for dd in range(kmax + 1):
                emit(f"B{r}_{dd}:")
                if r == 3:
                    prefetch()
                body(r, dd)
                if r == 3:
                    emit("    bra.uni DONE;")
                    continue
                if r + 2 < 4:
                    preds(r + 2, sets[r % 2])          # the set this direction just consumed
                setf(r + 1)
                jump_tree(r + 1, sets[(r + 1) % 2], 0, nbits - 1)
        emit("DONE:")
    else:
        for r in range(4):
            ps = sets[r % 2]
            if dbl and r + 1 < 4:
                preds(r + 1, sets[(r + 1) % 2])
            elif not dbl and r > 0:
                preds(r, ps)
            setf(r)
            if r == 3:
                prefetch()
            if BRX:
                labels = ", ".join(f"C{r}_{dd}" for dd in range(kmax + 1))
                emit(f"    bfe.u32 x, {E1}, {6 * r}, 6;")
                emit(f"    min.u32 x, x, {kmax};")
                emit(f"TS{r}: .branchtargets {labels};")
                emit(f"    brx.idx.uni x, TS{r};")
                for dd in range(kmax + 1):
                    emit(f"C{r}_{dd}:")
                    body(r, dd)
                    emit(f"    bra.uni J{r};")
            else:
                def tree(d0, bit):
                    if bit < 0:
                        body(r, d0)
                        emit(f"    bra.uni J{r};")
                        return
                    hi = d0 + (1 << bit)
                    if hi > kmax:
                        tree(d0, bit - 1)
                        return
                    emit(f"    @{ps}{bit} bra.uni T{r}_{hi}_{bit};")
                    tree(d0, bit - 1)
                    emit(f"T{r}_{hi}_{bit}:")
                    tree(hi, bit - 1)
                tree(0, nbits - 1)
            emit(f"J{r}:")
    emit("}")
    if BRX or (not chain and not dbl):
        # with a single predicate set direction 3's predicates read e1 before the prefetch overwrites it: fine,
        # the prefetch is emitted after preds(3)
        pass
    asm = "\n".join(f'        "{ln}\\n"' for ln in L)
    outs = ", ".join([f'"+l"(acc[{r}][{k}])' for r in range(4) for k in range(K)] + ['"+r"(e0)', '"+r"(e1)'] +
                     [f'"+f"(f{r})' for r in range(4)])
    return f"""// NCH = {nch}: window of {nw} sample pairs, deltas 0..{kmax}, {'chained' if chain else 'joined'} dispatch
template <>
__device__ __forceinline__ void tile_channel_step<{nch}>(u64 (&acc)[4][{K}], uint32_t &e0, uint32_t &e1, float &f0, float &f1,
                                                     float &f2, float &f3, uint32_t row, uint32_t nxt) {{
    asm volatile(
{asm}
        : {outs}
        : "r"(row), "r"(nxt)
        : "memory");
}}
"""


def gen_dual(nch):
    """Two windows per channel, each shared by a pair of directions (slots 0,1 -> window A, slots 2,3 -> window B),
    processed one after the other through the same registers.  For coarse grids whose 2x2 tiles spread over more
    samples than one 6-chunk window holds: the pairs along the array's short axis still fit, the kernel keeps the
    small 4-case dispatch, 128 registers and 16 warps."""
    nw, nd = 2 * nch, 2 * nch - 1
    kmax = 2 * nch - 9
    nbits = max(1, kmax.bit_length())
    depth = 8
    assert 2 * nbits <= 6
    sets = ["pa", "pb"]
    L = []
    emit = L.append

    def body(r, D):
        fma = lambda t, k: emit(f"    fma.rn.f32x2 t{t}, ff, d{D + k}, w{D + k + 1};")
        if ADD_AS_FMA:   # acc = fma(t, 1, acc): the same single rounding as the add, FFMA2 only on the FP pipe
            add = lambda k, t: emit(f"    fma.rn.f32x2 {A(r, k)}, t{t}, one, {A(r, k)};")
        else:
            add = lambda k, t: emit(f"    add.rn.f32x2 {A(r, k)}, {A(r, k)}, t{t};")
        for k in range(depth):
            fma(k, k)
        for k in range(K):
            add(k, k % depth)
            if k + depth < K:
                fma(k % depth, k + depth)

    def preds(r, ps):
        for b in range(nbits):
            emit(f"    and.b32 x, {E1}, {1 << (6 * r + b)};")
            emit(f"    setp.ne.b32 {ps}{b}, x, 0;")

    def setf(r):
        emit(f"    mov.b64 ff, {{{F(r)}, {F(r)}}};")

    def window(w):
        # window w: byte offset in the low / high half of e0, pad phase in bits 24-25 / 26-27 of e1
        if w == 0:
            emit(f"    and.b32 x, {E0}, 0xffff;")
        else:
            emit(f"    shr.u32 x, {E0}, 16;")
        emit(f"    add.u32 a0, {ROW}, x;")
        emit(f"    bfe.u32 rr, {E1}, {24 + 2 * w}, 2;")
        emit("    add.u32 x, a0, 16;")
        for cls, thr in ((1, 3), (2, 2), (3, 1)):
            emit(f"    setp.ge.u32 q{cls}, rr, {thr};")
            emit(f"    selp.u32 a{cls}, x, a0, q{cls};")
        for m in range(nch):
            emit(f"    ld.shared.v2.b64 {{w{2 * m}, w{2 * m + 1}}}, [a{m & 3}+{16 * (m + (m >> 2))}];")

    def subs():
        for j in range(nd):
            if ADD_AS_FMA:
                emit(f"    fma.rn.f32x2 d{j}, w{j + 1}, mone, w{j};")
            else:
                emit(f"    sub.rn.f32x2 d{j}, w{j}, w{j + 1};")

    uid = [0]

    def jump_tree(r, ps, d0, bit):
        if bit < 0:
            emit(f"    bra.uni B{r}_{d0};")
            return
        hi = d0 + (1 << bit)
        if hi > kmax:
            jump_tree(r, ps, d0, bit - 1)
            return
        uid[0] += 1
        lab = f"N{uid[0]}"
        emit(f"    @{ps}{bit} bra.uni {lab};")
        jump_tree(r, ps, d0, bit - 1)
        emit(f"{lab}:")
        jump_tree(r, ps, hi, bit - 1)

    emit("{")
    emit(f"    .reg .pred pa<{nbits}>, pb<{nbits}>, q<4>;")
    emit("    .reg .b32 x, rr, a<4>;")
    emit(f"    .reg .b64 ff, t<{depth}>, w<{nw}>, d<{nd}>;")
    if ADD_AS_FMA:
        emit("    .reg .b64 one, mone;")
        emit("    mov.b64 one, 0x3F8000003F800000;")
        emit("    mov.b64 mone, 0xBF800000BF800000;")
    window(0)
    preds(0, sets[0])
    preds(1, sets[1])
    subs()
    setf(0)
    jump_tree(0, sets[0], 0, nbits - 1)
    for dd in range(kmax + 1):            # slot 0 bodies chain into slot 1's dispatch
        emit(f"B0_{dd}:")
        body(0, dd)
        preds(2, sets[0])
        setf(1)
        jump_tree(1, sets[1], 0, nbits - 1)
    for dd in range(kmax + 1):            # slot 1 bodies join before window B is loaded
        emit(f"B1_{dd}:")
        body(1, dd)
        emit("    bra.uni HALF;")
    emit("HALF:")
    # bit 28 of e1: the whole 2x2 tile fits window A for this channel -> slots 2,3 reuse its registers
    emit(f"    and.b32 x, {E1}, {1 << 28};")
    emit("    setp.ne.b32 q0, x, 0;")
    preds(3, sets[1])
    emit("    @q0 bra.uni SAMEWIN;")
    window(1)
    subs()
    emit("SAMEWIN:")
    setf(2)
    jump_tree(2, sets[0], 0, nbits - 1)
    for dd in range(kmax + 1):
        emit(f"B2_{dd}:")
        body(2, dd)
        setf(3)
        jump_tree(3, sets[1], 0, nbits - 1)
    for dd in range(kmax + 1):
        emit(f"B3_{dd}:")
        # e0, e1, f0..f3 are dead from here on: fetch the next channel's entry into them
        emit(f"    ld.shared.v2.u32 {{{E0}, {E1}}}, [{NXT}];")
        emit(f"    ld.shared.v4.f32 {{{F(0)}, {F(1)}, {F(2)}, {F(3)}}}, [{NXT}+16];")
        body(3, dd)
        emit("    bra.uni DONE;")
    emit("DONE:")
    emit("}")
    asm = "\n".join(f'        "{ln}\\n"' for ln in L)
    outs = ", ".join([f'"+l"(acc[{r}][{k}])' for r in range(4) for k in range(K)] + ['"+r"(e0)', '"+r"(e1)'] +
                     [f'"+f"(f{r})' for r in range(4)])
    return f"""// two {nch}-chunk windows per channel (one per direction pair), deltas 0..{kmax}
template <>
__device__ __forceinline__ void tile_channel_step_dual<{nch}>(u64 (&acc)[4][{K}], uint32_t &e0, uint32_t &e1, float &f0, float &f1,
                                                          float &f2, float &f3, uint32_t row, uint32_t nxt) {{
    asm volatile(
{asm}
        : {outs}
        : "r"(row), "r"(nxt)
        : "memory");
}}
"""



def gen_fast(nch, dual=False):
    """Two-FMA form (tolerance mode): acc += g * s[i+1]; acc += f * s[i], g = fl(1 - f) from the table.  No differences,
    no temporaries: 16 FFMA2 per (direction, channel) and nothing else on the FP pipe.

    One asm block runs ALL channels of a pipeline stage; there is no join point per channel:
      * one chain per delta value: B0_d, B1_d, B2_d, B3_d.  The dispatch at the end of a body tests the next direction's
        delta bits with the polarity of its OWN delta, so "same delta as the previous direction" needs no taken branch;
        any other delta leaves through one taken branch into a shared subtree;
      * the last direction's body advances the entry pointer, fetches the next channel's table entry, forms the window
        addresses (the entry carries the four pad-class offsets, row offset inside the stage included, ready to add to
        the lane's base) and the first direction's predicates, then loops: the next iteration starts with its window
        loads.  No register that an in-flight LDS still reads as its address is written (ncu: that WAR wait on the
        short scoreboard was 15 % of all stall samples in the first version).
    dual: two windows per channel (slots 0,1 -> window A, slots 2,3 -> window B, through the same registers); bit 28 of
    the delta word says the whole tile fits window A for this channel (window B is not loaded).
    operands: acc[4][8] "+l" 0..31, ent 32 "+r" (shared address of the current entry), row 33 "r" (lane's base inside
    the stage's rows), end 34 "r" (shared address one past the stage's last entry of this warp).
    entry (single, 64 B): o[4] | dl, -, -, - | f[4] | g[4];   (dual, 80 B): oA[4] | oB[4] | f[4] | g[4] | dl, -, -, -"""
    nw = 2 * nch
    kmax = 2 * nch - 9
    nbits = max(1, kmax.bit_length())
    ENT, ROWR, END = "%32", "%33", "%34"
    esz = 80 if dual else 64
    o_dl, o_f, o_g = (64, 32, 48) if dual else (16, 32, 48)
    shared_tail = SHARED_TAIL == "1" or (SHARED_TAIL == "dual" and dual)
    PIPE = PIPE_MODE == "1" or (PIPE_MODE == "single" and not dual)
    # the last window chunk is loaded only for the (window, channel)s whose largest delta reaches it (kmax - 1 or more):
    # a predicated-off LDS still issues but moves no shared-memory wavefronts.  nch = 5 always needs its last chunk.
    adapt = ADAPT and kmax >= 2
    L = []
    emit = L.append

    def body(r, D):
        for k in range(K):
            emit(f"    fma.rn.f32x2 {A(r, k)}, gg{r}, w{D + k + 1}, {A(r, k)};")
        for k in range(K):
            emit(f"    fma.rn.f32x2 {A(r, k)}, ff{r}, w{D + k}, {A(r, k)};")

    def preds(r):
        for b in range(nbits):
            emit(f"    and.b32 x, dl, {1 << (6 * r + b)};")
            emit(f"    setp.ne.b32 p{b}, x, 0;")

    first_window = [True]

    def window(o):
        if os.environ.get("BFLK_GEN_EXP") == "nolds" and not first_window[0]:
            return      # timing experiment only (wrong results): how much do the window loads cost?
        for m in range(nch):
            # the last chunk (pairs 2 nch - 2, 2 nch - 1) is read only by bodies with delta >= kmax - 1: the table says per
            # (window, channel) whether any direction needs it (bit 29 window A, bit 30 window B of the delta word)
            if adapt and m == nch - 1 and ADAPT_MODE == "sel":
                # (a full definition of the chunk's registers: a predicated load would keep their old values alive and
                # ptxas then moves accumulators around the window-B block -- 135 instead of 82 MOVs in the dual kernel)
                emit(f"    selp.u32 xl, {o}{m & 3}, {ENT}, pw{o[1]};")
                emit(f"    ld.shared.v2.b64 {{w{2 * m}, w{2 * m + 1}}}, [xl+{16 * (m + (m >> 2))}];")
                continue
            pred = f"@pw{o[1]} " if adapt and m == nch - 1 else ""
            emit(f"    {pred}ld.shared.v2.b64 {{w{2 * m}, w{2 * m + 1}}}, [{o}{m & 3}+{16 * (m + (m >> 2))}];")

    def need_last(w, extra=""):
        """pw{w} = the window's last chunk is needed (and `extra`, a predicate)"""
        emit(f"    and.b32 x, dl, {1 << (29 if w == 'a' else 30)};")
        emit(f"    setp.ne{'.and' if extra else ''}.b32 pw{w}, x, 0{', ' + extra if extra else ''};")

    def load_entry_head():
        emit(f"    ld.shared.v4.u32 {{oa0, oa1, oa2, oa3}}, [{ENT}];")
        if dual:
            emit(f"    ld.shared.v4.u32 {{ob0, ob1, ob2, ob3}}, [{ENT}+16];")
        emit(f"    ld.shared.u32 dl, [{ENT}+{o_dl}];")

    def load_entry_fracs():
        emit(f"    ld.shared.v4.f32 {{f0, f1, f2, f3}}, [{ENT}+{o_f}];")
        emit(f"    ld.shared.v4.f32 {{g0, g1, g2, g3}}, [{ENT}+{o_g}];")

    prologue = [True]

    def entry_tail():
        """addresses of the next window A (pipelined variant: formed at the start of B3 instead), first direction's predicates"""
        if prologue[0] or not PIPE:
            for c in range(4):
                emit(f"    add.u32 oa{c}, oa{c}, {ROWR};")
            if adapt:
                need_last("a")
        prologue[0] = False
        preds(0)

    def subtree(r, lo, bit, tag):
        """standard tree over targets [lo, lo + 2^(bit+1)) & <= kmax using bits bit..0, leaves jump to the bodies"""
        if bit < 0:
            emit(f"    bra.uni B{r}_{lo};")
            return
        hi = lo + (1 << bit)
        if hi > kmax:
            subtree(r, lo, bit - 1, tag)
            return
        lab = f"T{tag}_{hi}_{bit}"
        emit(f"    @p{bit} bra.uni {lab};")
        subtree(r, lo, bit - 1, tag)
        emit(f"{lab}:")
        subtree(r, hi, bit - 1, tag)

    need = set()   # shared subtrees S{r}_{base}_{b}: targets [base, base + 2^b)

    def tree_from(r, dd):
        """dispatch of direction r, falling through into B{r}_{dd}"""
        for b in reversed(range(nbits)):
            want = (dd >> b) & 1
            base = ((dd >> (b + 1)) << (b + 1)) | ((1 - want) << b)
            if base > kmax:
                continue
            need.add((r, base, b))
            emit(f"    @{'!' if want else ''}p{b} bra.uni S{r}_{base}_{b};")

    emit("{")
    emit(f"    .reg .pred p<{nbits}>, ploop, q, pwa, pwb;")
    emit("    .reg .b32 x, xl, dl, oa<4>, ob<4>;")
    emit("    .reg .f32 f<4>, g<4>;")
    emit(f"    .reg .b64 ff<4>, gg<4>, w<{nw}>;")
    load_entry_head()
    load_entry_fracs()
    entry_tail()
    if os.environ.get("BFLK_GEN_EXP") == "nolds":
        window("oa")
        first_window[0] = False
    if PIPE:
        window("oa")          # first channel of the stage; later ones are loaded from inside the previous B3 body
    emit("TOP:")
    if not PIPE:
        window("oa")
    for r in range(4):
        emit(f"    mov.b64 ff{r}, {{f{r}, f{r}}};")
        emit(f"    mov.b64 gg{r}, {{g{r}, g{r}}};")
    tree_from(0, 0)

    def body_recycling(r, D):
        """body(r, D) in an order that frees window registers early (g0 g1 g2 g3 | f0 g4 | f1 g5 | f2 g6 | f3 g7 | f4..f7:
        a dependent pair is never closer than four FFMA2), with the next channel's window chunk m loaded (if the stage
        goes on) as soon as both registers of the chunk have been read for the last time."""
        g = lambda k: f"    fma.rn.f32x2 {A(r, k)}, gg{r}, w{D + k + 1}, {A(r, k)};"
        f = lambda k: f"    fma.rn.f32x2 {A(r, k)}, ff{r}, w{D + k}, {A(r, k)};"
        order = [("g", 0), ("g", 1), ("g", 2), ("g", 3)]
        for k in range(4):
            order += [("f", k), ("g", k + 4)]
        order += [("f", k) for k in range(4, 8)]
        last_use = {}
        for pos, (kind, k) in enumerate(order):
            last_use[D + k + (1 if kind == "g" else 0)] = pos
        loaded = set()

        def issue_free(pos):
            for m in range(nch):
                if m in loaded:
                    continue
                if all(last_use.get(j, -1) <= pos for j in (2 * m, 2 * m + 1)):
                    loaded.add(m)
                    if adapt and m == nch - 1 and ADAPT_MODE == "sel":
                        emit(f"    @ploop ld.shared.v2.b64 {{w{2 * m}, w{2 * m + 1}}}, [xl+{16 * (m + (m >> 2))}];")
                        continue
                    pred = "pwa" if adapt and m == nch - 1 else "ploop"
                    emit(f"    @{pred} ld.shared.v2.b64 {{w{2 * m}, w{2 * m + 1}}}, [oa{m & 3}+{16 * (m + (m >> 2))}];")

        issue_free(-1)
        for pos, (kind, k) in enumerate(order):
            emit(g(k) if kind == "g" else f(k))
            issue_free(pos)
        assert len(loaded) == nch

    for dd in range(kmax + 1):
        for r in range(4):
            emit(f"B{r}_{dd}:")
            if r == 0:
                preds(1)
                body(r, dd)
                tree_from(1, dd)
            elif r == 1:
                preds(2)
                if dual:
                    emit(f"    and.b32 x, dl, {1 << 28};")
                    emit("    setp.ne.b32 q, x, 0;")
                    if adapt and ADAPT_PWB_EARLY:
                        need_last("b")
                body(r, dd)
                if dual and NOJOIN:
                    # no join point after the window-B block: the same-window path gets its own copy of the next
                    # dispatch (out of line, SWS_dd), so ptxas sees no if-then to wrap in BSSY / BSYNC
                    emit(f"    @q bra.uni SWS_{dd};")
                    for c in range(4):
                        emit(f"    add.u32 ob{c}, ob{c}, {ROWR};")
                    if adapt:
                        need_last("b")
                    window("ob")
                elif dual:
                    # address adds inside the skipped block: ten instructions are too many for ptxas to if-convert, so a
                    # tile that fits window A really branches around them instead of issuing six predicated-off loads
                    emit(f"    @q bra.uni SW_{dd};")
                    for c in range(4):
                        emit(f"    add.u32 ob{c}, ob{c}, {ROWR};")
                    if adapt and not ADAPT_PWB_EARLY:
                        need_last("b")
                    window("ob")
                    emit(f"SW_{dd}:")
                tree_from(2, dd)
            elif r == 2:
                preds(3)
                if PIPE:
                    # the next channel's entry head already here: its window addresses are needed inside B3
                    emit(f"    add.u32 {ENT}, {ENT}, {esz};")
                    load_entry_head()
                    emit(f"    setp.ne.u32 ploop, {ENT}, {END};")
                body(r, dd)
                tree_from(3, dd)
            else:
                if PIPE:
                    for c in range(4):
                        emit(f"    add.u32 oa{c}, oa{c}, {ROWR};")
                    if adapt and ADAPT_MODE == "sel":
                        need_last("a")
                        emit(f"    selp.u32 xl, oa{(nch - 1) & 3}, {ENT}, pwa;")
                    elif adapt:
                        need_last("a", "ploop")
                    body_recycling(r, dd)
                else:
                    emit(f"    add.u32 {ENT}, {ENT}, {esz};")
                    load_entry_head()
                    emit(f"    setp.ne.u32 ploop, {ENT}, {END};")
                    body(r, dd)
                if shared_tail:
                    emit("    bra.uni TAIL;")
                else:
                    load_entry_fracs()
                    entry_tail()
                    emit("    @ploop bra.uni TOP;")
                    emit("    bra.uni DONE;")
    if shared_tail:
        emit("TAIL:")
        load_entry_fracs()
        entry_tail()
        emit("    @ploop bra.uni TOP;")
        emit("    bra.uni DONE;")
    if dual and NOJOIN:
        for dd in range(kmax + 1):
            emit(f"SWS_{dd}:")
            tree_from(2, dd)
            emit(f"    bra.uni B2_{dd};")
    for (r, base, b) in sorted(need):
        emit(f"S{r}_{base}_{b}:")
        subtree(r, base, b - 1, f"{r}_{base}_{b}")
    emit("DONE:")
    emit("}")
    asm = "\n".join(f'        "{ln}\\n"' for ln in L)
    outs = ", ".join([f'"+l"(acc[{r}][{k}])' for r in range(4) for k in range(K)] + ['"+r"(ent)'])
    name = "tile_stage_fast_dual" if dual else "tile_stage_fast"
    return f"""// NCH = {nch}: {'two windows' if dual else 'window'} of {nw} sample pairs, deltas 0..{kmax}
template <>
__device__ __forceinline__ void {name}<{nch}>(u64 (&acc)[4][{K}], uint32_t ent, uint32_t row, uint32_t end) {{
    asm volatile(
{asm}
        : {outs}
        : "r"(row), "r"(end)
        : "memory");
}}
"""



def gen_wide(nch, exact=False):
    """Wide flavour for coarse grids: ONE direction pair per warp, lane l owns 16 consecutive sample pairs (a 512-sample block
    per slot of the block pair), one window of 2*nch pairs shared by the two directions.  Per (warp, channel): 64 FFMA2 behind
    TWO dispatches (32 FFMA2 each) and nch window loads -- half the dispatch and table traffic per FLOP of the 2x2 tile, which
    on coarse grids (cfg3: a 2x2 tile spreads over up to 9 samples) needs two windows and four dispatches for the same work.
    Packed rows carry one 16-byte pad per EIGHT chunks (lane stride 9 chunks = 144 B: bank-conflict free for LDS.128), so a
    window chunk m sits at padded chunk m + ((r + m) >> 3), r = (first chunk) & 7: eight pad classes, offsets from the table.
    entry (48 B): o[8] | dl, f0, f1, -   (g = 1 - f formed here with the table builder's rounding: sub.rn.f32)
    exact=True: the reference's operation triple (differences once per window, fma + add per direction).
    operands: acc[2][16] "+l" 0..31, ent 32 "+r", row 33 "r" (lane base inside the stage's rows), end 34 "r"."""
    KW = 16
    nw = 2 * nch
    kmax = nw - (KW + 1)
    assert kmax >= 0
    nbits = max(1, kmax.bit_length())
    ENT, ROWR, END = "%32", "%33", "%34"
    esz = 48
    PIPE = PIPE_MODE_WIDE == "1" and not exact
    AW = lambda r, k: f"%{r * KW + k}"
    L = []
    emit = L.append

    def body(r, D):
        if exact:
            # acc = acc + fma(f, s[i] - s[i+1], s[i+1]) (delay.cpp:24), FFMA2 -> FADD2 software-pipelined 8 deep
            depth = 8
            fma = lambda t, k: emit(f"    fma.rn.f32x2 t{t}, ff{r}, d{D + k}, w{D + k + 1};")
            if ADD_AS_FMA:   # acc = fma(t, 1, acc): the same single rounding as the add, on the FFMA2 path only
                add = lambda k, t: emit(f"    fma.rn.f32x2 {AW(r, k)}, t{t}, one, {AW(r, k)};")
            else:
                add = lambda k, t: emit(f"    add.rn.f32x2 {AW(r, k)}, {AW(r, k)}, t{t};")
            for k in range(depth):
                fma(k, k)
            for k in range(KW):
                add(k, k % depth)
                if k + depth < KW:
                    fma(k % depth, k + depth)
            return
        for k in range(KW):
            emit(f"    fma.rn.f32x2 {AW(r, k)}, gg{r}, w{D + k + 1}, {AW(r, k)};")
        for k in range(KW):
            emit(f"    fma.rn.f32x2 {AW(r, k)}, ff{r}, w{D + k}, {AW(r, k)};")

    def preds(r):
        for b in range(nbits):
            emit(f"    and.b32 x, dl, {1 << (6 * r + b)};")
            emit(f"    setp.ne.b32 p{b}, x, 0;")

    def window(pred=""):
        for m in range(nch):
            emit(f"    {pred}ld.shared.v2.b64 {{w{2 * m}, w{2 * m + 1}}}, [oa{m & 7}+{16 * (m + (m >> 3))}];")

    def subs():
        for j in range(nw - 1):
            if ADD_AS_FMA:   # s[i] - s[i+1] = fma(s[i+1], -1, s[i]): one rounding, same result
                emit(f"    fma.rn.f32x2 d{j}, w{j + 1}, mone, w{j};")
            else:
                emit(f"    sub.rn.f32x2 d{j}, w{j}, w{j + 1};")

    def load_entry_head():
        emit(f"    ld.shared.v4.u32 {{oa0, oa1, oa2, oa3}}, [{ENT}];")
        emit(f"    ld.shared.v4.u32 {{oa4, oa5, oa6, oa7}}, [{ENT}+16];")

    def load_entry_fracs():
        emit(f"    ld.shared.v4.b32 {{dl, fb0, fb1, x}}, [{ENT}+32];")
        for r in range(2):
            emit(f"    mov.b32 f{r}, fb{r};")
            if not exact:
                emit(f"    sub.rn.f32 g{r}, 0f3F800000, f{r};")

    def add_row():
        for c in range(8):
            emit(f"    add.u32 oa{c}, oa{c}, {ROWR};")

    def subtree(r, lo, bit, tag):
        if bit < 0:
            emit(f"    bra.uni B{r}_{lo};")
            return
        hi = lo + (1 << bit)
        if hi > kmax:
            subtree(r, lo, bit - 1, tag)
            return
        lab = f"T{tag}_{hi}_{bit}"
        emit(f"    @p{bit} bra.uni {lab};")
        subtree(r, lo, bit - 1, tag)
        emit(f"{lab}:")
        subtree(r, hi, bit - 1, tag)

    need = set()

    def tree_from(r, dd):
        for b in reversed(range(nbits)):
            want = (dd >> b) & 1
            base = ((dd >> (b + 1)) << (b + 1)) | ((1 - want) << b)
            if base > kmax:
                continue
            need.add((r, base, b))
            emit(f"    @{'!' if want else ''}p{b} bra.uni S{r}_{base}_{b};")

    def body_recycling(r, D):
        """two-FMA body in an order that frees window registers early; the next channel's window chunk m is loaded (if the
        stage goes on) as soon as both its registers have been read for the last time"""
        g = lambda k: f"    fma.rn.f32x2 {AW(r, k)}, gg{r}, w{D + k + 1}, {AW(r, k)};"
        f = lambda k: f"    fma.rn.f32x2 {AW(r, k)}, ff{r}, w{D + k}, {AW(r, k)};"
        order = [("g", 0), ("g", 1), ("g", 2), ("g", 3)]
        for k in range(KW - 4):
            order += [("f", k), ("g", k + 4)]
        order += [("f", k) for k in range(KW - 4, KW)]
        last_use = {}
        for pos, (kind, k) in enumerate(order):
            last_use[D + k + (1 if kind == "g" else 0)] = pos
        loaded = set()

        def issue_free(pos):
            for m in range(nch):
                if m in loaded:
                    continue
                if all(last_use.get(j, -1) <= pos for j in (2 * m, 2 * m + 1)):
                    loaded.add(m)
                    emit(f"    @ploop ld.shared.v2.b64 {{w{2 * m}, w{2 * m + 1}}}, [oa{m & 7}+{16 * (m + (m >> 3))}];")

        issue_free(-1)
        for pos, (kind, k) in enumerate(order):
            emit(g(k) if kind == "g" else f(k))
            issue_free(pos)
        assert len(loaded) == nch

    emit("{")
    emit(f"    .reg .pred p<{nbits}>, ploop;")
    emit("    .reg .b32 x, dl, fb<2>, oa<8>;")
    emit("    .reg .f32 f<2>, g<2>;")
    emit(f"    .reg .b64 ff<2>, gg<2>, w<{nw}>" + (f", d<{nw - 1}>, t<8>, one, mone;" if exact else ";"))
    if exact and ADD_AS_FMA:
        emit("    mov.b64 one, 0x3F8000003F800000;")
        emit("    mov.b64 mone, 0xBF800000BF800000;")
    load_entry_head()
    load_entry_fracs()
    add_row()
    preds(0)
    if PIPE:
        window()
    emit("TOP:")
    if not PIPE:
        window()
    for r in range(2):
        emit(f"    mov.b64 ff{r}, {{f{r}, f{r}}};")
        if not exact:
            emit(f"    mov.b64 gg{r}, {{g{r}, g{r}}};")
    if exact:
        subs()
    tree_from(0, 0)
    for dd in range(kmax + 1):
        emit(f"B0_{dd}:")
        preds(1)
        # the next channel's entry head already here: the window addresses must be ready before the last body ends
        emit(f"    add.u32 {ENT}, {ENT}, {esz};")
        if PIPE:
            load_entry_head()
        emit(f"    setp.ne.u32 ploop, {ENT}, {END};")
        body(0, dd)
        tree_from(1, dd)
        emit(f"B1_{dd}:")
        if PIPE:
            add_row()
            body_recycling(1, dd)
        else:
            load_entry_head()
            body(1, dd)
        emit("    bra.uni TAIL;")
    emit("TAIL:")
    load_entry_fracs()
    if not PIPE:
        add_row()
    preds(0)
    emit("    @ploop bra.uni TOP;")
    emit("    bra.uni DONE;")
    for (r, base, b) in sorted(need):
        emit(f"S{r}_{base}_{b}:")
        subtree(r, base, b - 1, f"{r}_{base}_{b}")
    emit("DONE:")
    emit("}")
    asm = "\n".join(f'        "{ln}\\n"' for ln in L)
    outs = ", ".join([f'"+l"(acc[{r}][{k}])' for r in range(2) for k in range(KW)] + ['"+r"(ent)'])
    name = "tile_stage_wide_exact" if exact else "tile_stage_wide"
    return f"""// NCH = {nch}: direction pair, 16 sample pairs per lane, window of {nw} sample pairs, deltas 0..{kmax}{', exact triple' if exact else ''}
template <>
__device__ __forceinline__ void {name}<{nch}>(u64 (&acc)[2][{KW}], uint32_t ent, uint32_t row, uint32_t end) {{
    asm volatile(
{asm}
        : {outs}
        : "r"(row), "r"(end)
        : "memory");
}}
"""



def gen_k16(nch, dual=False):
    """Two-FMA form with SIXTEEN sample pairs per lane (512-sample blocks) and the usual 2x2 direction tile: 128 accumulator
    registers, so 8 warps per CTA at 255 registers -- but per (warp, channel) 128 FFMA2 behind the same four dispatches, the
    same entry loads and a window only 25 % longer than the 8-pair flavour's: half the non-FP instructions and ~0.8x the
    shared-memory wavefronts per FLOP.  Packed rows carry one 16-byte pad per EIGHT chunks (lane stride 9 chunks = 144 B),
    eight pad classes.  Structure as gen_fast (one chain per delta value, loop tail per body, no pipelined window loads).
    operands: acc[4][16] "+l" 0..63, ent 64 "+r", row 65 "r", end 66 "r".
    entry (single, 80 B): o[8] | f[4] | g[4] | dl, -, -, -;   (dual, 112 B): oA[8] | oB[8] | f[4] | g[4] | dl, -, -, -"""
    KW = 16
    nw = 2 * nch
    kmax = nw - (KW + 1)
    assert kmax >= 0
    nbits = max(1, kmax.bit_length())
    ENT, ROWR, END = "%64", "%65", "%66"
    esz = 112 if dual else 80
    o_f, o_g, o_dl = (64, 80, 96) if dual else (32, 48, 64)
    AW = lambda r, k: f"%{r * KW + k}"
    L = []
    emit = L.append

    def body(r, D):
        for k in range(KW):
            emit(f"    fma.rn.f32x2 {AW(r, k)}, gg{r}, w{D + k + 1}, {AW(r, k)};")
        for k in range(KW):
            emit(f"    fma.rn.f32x2 {AW(r, k)}, ff{r}, w{D + k}, {AW(r, k)};")

    def preds(r):
        for b in range(nbits):
            emit(f"    and.b32 x, dl, {1 << (6 * r + b)};")
            emit(f"    setp.ne.b32 p{b}, x, 0;")

    def window(o):
        for m in range(nch):
            emit(f"    ld.shared.v2.b64 {{w{2 * m}, w{2 * m + 1}}}, [{o}{m & 7}+{16 * (m + (m >> 3))}];")

    def load_entry_head():
        emit(f"    ld.shared.v4.u32 {{oa0, oa1, oa2, oa3}}, [{ENT}];")
        emit(f"    ld.shared.v4.u32 {{oa4, oa5, oa6, oa7}}, [{ENT}+16];")
        if dual:
            emit(f"    ld.shared.v4.u32 {{ob0, ob1, ob2, ob3}}, [{ENT}+32];")
            emit(f"    ld.shared.v4.u32 {{ob4, ob5, ob6, ob7}}, [{ENT}+48];")
        emit(f"    ld.shared.u32 dl, [{ENT}+{o_dl}];")

    def load_entry_fracs():
        emit(f"    ld.shared.v4.f32 {{f0, f1, f2, f3}}, [{ENT}+{o_f}];")
        emit(f"    ld.shared.v4.f32 {{g0, g1, g2, g3}}, [{ENT}+{o_g}];")

    def entry_tail():
        for c in range(8):
            emit(f"    add.u32 oa{c}, oa{c}, {ROWR};")
        preds(0)

    def subtree(r, lo, bit, tag):
        if bit < 0:
            emit(f"    bra.uni B{r}_{lo};")
            return
        hi = lo + (1 << bit)
        if hi > kmax:
            subtree(r, lo, bit - 1, tag)
            return
        lab = f"T{tag}_{hi}_{bit}"
        emit(f"    @p{bit} bra.uni {lab};")
        subtree(r, lo, bit - 1, tag)
        emit(f"{lab}:")
        subtree(r, hi, bit - 1, tag)

    need = set()

    def tree_from(r, dd):
        for b in reversed(range(nbits)):
            want = (dd >> b) & 1
            base = ((dd >> (b + 1)) << (b + 1)) | ((1 - want) << b)
            if base > kmax:
                continue
            need.add((r, base, b))
            emit(f"    @{'!' if want else ''}p{b} bra.uni S{r}_{base}_{b};")

    emit("{")
    emit(f"    .reg .pred p<{nbits}>, ploop, q;")
    emit("    .reg .b32 x, dl, oa<8>, ob<8>;")
    emit("    .reg .f32 f<4>, g<4>;")
    emit(f"    .reg .b64 ff<4>, gg<4>, w<{nw}>;")
    load_entry_head()
    load_entry_fracs()
    entry_tail()
    emit("TOP:")
    window("oa")
    for r in range(4):
        emit(f"    mov.b64 ff{r}, {{f{r}, f{r}}};")
        emit(f"    mov.b64 gg{r}, {{g{r}, g{r}}};")
    tree_from(0, 0)
    for dd in range(kmax + 1):
        for r in range(4):
            emit(f"B{r}_{dd}:")
            if r == 0:
                preds(1)
                body(r, dd)
                tree_from(1, dd)
            elif r == 1:
                preds(2)
                if dual:
                    emit(f"    and.b32 x, dl, {1 << 28};")
                    emit("    setp.ne.b32 q, x, 0;")
                body(r, dd)
                if dual:
                    emit(f"    @q bra.uni SW_{dd};")
                    for c in range(8):
                        emit(f"    add.u32 ob{c}, ob{c}, {ROWR};")
                    window("ob")
                    emit(f"SW_{dd}:")
                tree_from(2, dd)
            elif r == 2:
                preds(3)
                body(r, dd)
                tree_from(3, dd)
            else:
                emit(f"    add.u32 {ENT}, {ENT}, {esz};")
                load_entry_head()
                emit(f"    setp.ne.u32 ploop, {ENT}, {END};")
                body(r, dd)
                load_entry_fracs()
                entry_tail()
                emit("    @ploop bra.uni TOP;")
                emit("    bra.uni DONE;")
    for (r, base, b) in sorted(need):
        emit(f"S{r}_{base}_{b}:")
        subtree(r, base, b - 1, f"{r}_{base}_{b}")
    emit("DONE:")
    emit("}")
    asm = "\n".join(f'        "{ln}\\n"' for ln in L)
    outs = ", ".join([f'"+l"(acc[{r}][{k}])' for r in range(4) for k in range(KW)] + ['"+r"(ent)'])
    name = "tile_stage_k16_dual" if dual else "tile_stage_k16"
    return f"""// NCH = {nch}: 16 sample pairs per lane, {'two windows' if dual else 'window'} of {nw} sample pairs, deltas 0..{kmax}
template <>
__device__ __forceinline__ void {name}<{nch}>(u64 (&acc)[4][{KW}], uint32_t ent, uint32_t row, uint32_t end) {{
    asm volatile(
{asm}
        : {outs}
        : "r"(row), "r"(end)
        : "memory");
}}
"""


# Measured and rejected: a double-buffered flavour (window B of a channel in flight to a second register set while slots
# 0,1 run on window A, then window A of the next channel while slots 2,3 run): ~165 registers -> 12 warps per CTA (the
# register file is split per scheduler, so 14 warps still cap a thread at 128), always two window loads per channel ->
# shared memory 79 % busy and 0.45 of the FP32 peak on every configuration (profiles/README.md).



def gen_k6(dual=True, KK=6):
    """Experiment (tile_bench only): the two-FMA stage loop with KK = 6 sample pairs per lane (192-sample blocks).  The lane
    stride is three 16-byte chunks -- odd, so LDS.128 is conflict free WITHOUT pad chunks: one address per window instead of
    four class offsets, 48-byte entries (oa | ob | dl | - ; f[4] ; g[4]), a 5-chunk window (6 + 1 + span 3 pairs), 48
    accumulator registers -> ~90 registers, 20 warps per CTA (five per scheduler).  Fewer FFMA2 per dispatch (12 instead
    of 16) against fewer address / load instructions and a fifth warp to fill the issue slots."""
    nch = (KK + 1 + 3 + 1) // 2
    nw = 2 * nch
    kmax = nw - (KK + 1)
    nbits = max(1, kmax.bit_length())
    n_acc = 4 * KK
    ENT, ROWR, END = f"%{n_acc}", f"%{n_acc + 1}", f"%{n_acc + 2}"
    esz = 48
    AK = lambda r, k: f"%{r * KK + k}"
    L = []
    emit = L.append

    def body(r, D):
        for k in range(KK):
            emit(f"    fma.rn.f32x2 {AK(r, k)}, gg{r}, w{D + k + 1}, {AK(r, k)};")
        for k in range(KK):
            emit(f"    fma.rn.f32x2 {AK(r, k)}, ff{r}, w{D + k}, {AK(r, k)};")

    def preds(r):
        for b in range(nbits):
            emit(f"    and.b32 x, dl, {1 << (6 * r + b)};")
            emit(f"    setp.ne.b32 p{b}, x, 0;")

    def window(o):
        for m in range(nch):
            emit(f"    ld.shared.v2.b64 {{w{2 * m}, w{2 * m + 1}}}, [{o}+{16 * m}];")

    def load_entry_head():
        emit(f"    ld.shared.v4.u32 {{oa, ob, dl, x}}, [{ENT}];")

    def load_entry_fracs():
        emit(f"    ld.shared.v4.f32 {{f0, f1, f2, f3}}, [{ENT}+16];")
        emit(f"    ld.shared.v4.f32 {{g0, g1, g2, g3}}, [{ENT}+32];")

    def entry_tail():
        emit(f"    add.u32 oa, oa, {ROWR};")
        preds(0)

    def subtree(r, lo, bit, tag):
        if bit < 0:
            emit(f"    bra.uni B{r}_{lo};")
            return
        hi = lo + (1 << bit)
        if hi > kmax:
            subtree(r, lo, bit - 1, tag)
            return
        lab = f"T{tag}_{hi}_{bit}"
        emit(f"    @p{bit} bra.uni {lab};")
        subtree(r, lo, bit - 1, tag)
        emit(f"{lab}:")
        subtree(r, hi, bit - 1, tag)

    need = set()

    def tree_from(r, dd):
        for b in reversed(range(nbits)):
            want = (dd >> b) & 1
            base = ((dd >> (b + 1)) << (b + 1)) | ((1 - want) << b)
            if base > kmax:
                continue
            need.add((r, base, b))
            emit(f"    @{'!' if want else ''}p{b} bra.uni S{r}_{base}_{b};")

    emit("{")
    emit(f"    .reg .pred p<{nbits}>, ploop, q;")
    emit("    .reg .b32 x, dl, oa, ob;")
    emit("    .reg .f32 f<4>, g<4>;")
    emit(f"    .reg .b64 ff<4>, gg<4>, w<{nw}>;")
    load_entry_head()
    load_entry_fracs()
    entry_tail()
    emit("TOP:")
    window("oa")
    for r in range(4):
        emit(f"    mov.b64 ff{r}, {{f{r}, f{r}}};")
        emit(f"    mov.b64 gg{r}, {{g{r}, g{r}}};")
    tree_from(0, 0)
    for dd in range(kmax + 1):
        for r in range(4):
            emit(f"B{r}_{dd}:")
            if r == 0:
                preds(1)
                body(r, dd)
                tree_from(1, dd)
            elif r == 1:
                preds(2)
                if dual:
                    emit(f"    and.b32 x, dl, {1 << 28};")
                    emit("    setp.ne.b32 q, x, 0;")
                body(r, dd)
                if dual:
                    # no join point: a block this short would be if-converted into five predicated loads that still issue
                    # when the tile fits window A; the same-window path gets its own copy of the dispatch instead (SWS_dd)
                    emit(f"    @q bra.uni SWS_{dd};")
                    emit(f"    add.u32 ob, ob, {ROWR};")
                    window("ob")
                tree_from(2, dd)
            elif r == 2:
                preds(3)
                body(r, dd)
                tree_from(3, dd)
            else:
                emit(f"    add.u32 {ENT}, {ENT}, {esz};")
                load_entry_head()
                emit(f"    setp.ne.u32 ploop, {ENT}, {END};")
                body(r, dd)
                load_entry_fracs()
                entry_tail()
                emit("    @ploop bra.uni TOP;")
                emit("    bra.uni DONE;")
    if dual:
        for dd in range(kmax + 1):
            emit(f"SWS_{dd}:")
            tree_from(2, dd)
            emit(f"    bra.uni B2_{dd};")
    for (r, base, b) in sorted(need):
        emit(f"S{r}_{base}_{b}:")
        subtree(r, base, b - 1, f"{r}_{base}_{b}")
    emit("DONE:")
    emit("}")
    asm = "\n".join(f'        "{ln}\\n"' for ln in L)
    outs = ", ".join([f'"+l"(acc[{r}][{k}])' for r in range(4) for k in range(KK)] + ['"+r"(ent)'])
    name = "tile_stage_k6_dual" if dual else "tile_stage_k6"
    return f"""// KK = {KK}: {'two windows' if dual else 'window'} of {nw} sample pairs, deltas 0..{kmax}, unpadded rows (lane stride {KK // 2} chunks)
__device__ __forceinline__ void {name}(u64 (&acc)[4][{KK}], uint32_t ent, uint32_t row, uint32_t end) {{
    asm volatile(
{asm}
        : {outs}
        : "r"(row), "r"(end)
        : "memory");
}}
"""


import sys
if "--fast" in sys.argv:
    print("// GENERATED by tools/gen_tile_asm.py --fast -- do not edit.  See that script for the why.")
    print("""// All channels of one pipeline stage for one tile, two-FMA form (tolerance mode).
template <int NCH>
__device__ __forceinline__ void tile_stage_fast(u64 (&acc)[4][8], uint32_t ent, uint32_t row, uint32_t end);
template <int NCH>
__device__ __forceinline__ void tile_stage_fast_dual(u64 (&acc)[4][8], uint32_t ent, uint32_t row, uint32_t end);
""")
    for nch in (5, 6, 7, 8, 9, 10):
        print(gen_fast(nch))
    for nch in (6, 7):
        print(gen_fast(nch, dual=True))
    sys.exit(0)
if "--k6" in sys.argv:
    print("// GENERATED by tools/gen_tile_asm.py --k6 -- do not edit.  Experiment for tools/ubench/tile_bench.cu.")
    print(gen_k6(dual=True))
    print(gen_k6(dual=False))
    sys.exit(0)
if "--k16" in sys.argv:
    print("// GENERATED by tools/gen_tile_asm.py --k16 -- do not edit.  See that script for the why.")
    print("""// All channels of one pipeline stage for one 2x2 tile, 16 sample pairs per lane (512-sample blocks), two-FMA form.
template <int NCH>
__device__ __forceinline__ void tile_stage_k16(u64 (&acc)[4][16], uint32_t ent, uint32_t row, uint32_t end);
template <int NCH>
__device__ __forceinline__ void tile_stage_k16_dual(u64 (&acc)[4][16], uint32_t ent, uint32_t row, uint32_t end);
""")
    for nch in (9, 10, 11):
        print(gen_k16(nch))
    for nch in (10, 11):
        print(gen_k16(nch, dual=True))
    sys.exit(0)
if "--wide" in sys.argv:
    print("// GENERATED by tools/gen_tile_asm.py --wide -- do not edit.  See that script for the why.")
    print("""// All channels of one pipeline stage for one direction PAIR, 16 sample pairs per lane (512-sample blocks).
template <int NCH>
__device__ __forceinline__ void tile_stage_wide(u64 (&acc)[2][16], uint32_t ent, uint32_t row, uint32_t end);
template <int NCH>
__device__ __forceinline__ void tile_stage_wide_exact(u64 (&acc)[2][16], uint32_t ent, uint32_t row, uint32_t end);
""")
    for nch in (9, 10, 11):
        print(gen_wide(nch))
    for nch in (9, 10):
        print(gen_wide(nch, exact=True))
    sys.exit(0)

print("// GENERATED by tools/gen_tile_asm.py -- do not edit.  See that script for the why.")
print("""// One channel of one tile: window loads, differences, the four directions' accumulate bodies, and the prefetch of
// the next channel's table entry (e0 = window byte offset, e1 = 4 x 6-bit deltas | pad phase << 24, f = fractions).
template <int NCH>
__device__ __forceinline__ void tile_channel_step(u64 (&acc)[4][8], uint32_t &e0, uint32_t &e1, float &f0, float &f1,
                                                  float &f2, float &f3, uint32_t row, uint32_t nxt);
""")
for nch in (5, 6, 7, 8, 10):
    print(gen(nch))
print("""template <int NCH>
__device__ __forceinline__ void tile_channel_step_dual(u64 (&acc)[4][8], uint32_t &e0, uint32_t &e1, float &f0, float &f1,
                                                       float &f2, float &f3, uint32_t row, uint32_t nxt);
""")
for nch in (6, 7):
    print(gen_dual(nch))
