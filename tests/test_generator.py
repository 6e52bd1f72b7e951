"""The PTX channel loops are generated (tools/gen_tile_asm.py): the committed .inc files must be what the generator
emits, and the generated two-FMA bodies must be what DESIGN.md says they are -- checked statically on the PTX text:
every direction body accumulates g * w[D+k+1] and f * w[D+k] into acc[r][k] for k = 0..7 exactly once, and a window
chunk that the last body reloads early (pipelined single-window flavour) is never read again afterwards."""
import os
import re
import subprocess
import sys

import pytest

from conftest import PKG, ROOT

GEN = os.path.join(ROOT, "tools", "gen_tile_asm.py")
CSRC = os.path.join(PKG, "csrc")


def _gen(*args):
    env = {k: v for k, v in os.environ.items() if not k.startswith("BFLK_GEN_")}
    return subprocess.run([sys.executable, GEN, *args], capture_output=True, text=True, check=True, env=env).stdout


def test_committed_asm_is_the_generator_output():
    assert _gen() == open(os.path.join(CSRC, "das_tile_asm.inc")).read()
    assert _gen("--fast") == open(os.path.join(CSRC, "das_tile_fast_asm.inc")).read()


def _functions(text):
    """{(name, nch): [ptx lines]} of the generated specialisations."""
    out = {}
    for m in re.finditer(r"void (tile_stage_fast(?:_dual)?)<(\d+)>\(.*?asm volatile\(\n(.*?)\n        :", text, re.S):
        lines = [ln.strip()[1:-3].strip() for ln in m.group(3).splitlines()]
        out[(m.group(1), int(m.group(2)))] = [ln for ln in lines if ln]
    return out


@pytest.mark.parametrize("name,nch", [("tile_stage_fast", n) for n in (5, 6, 7, 8, 9, 10)] + [("tile_stage_fast_dual", n) for n in (6, 7)])
def test_two_fma_bodies(name, nch):
    lines = _functions(open(os.path.join(CSRC, "das_tile_fast_asm.inc")).read())[(name, nch)]
    kmax = 2 * nch - 9
    labels = {ln[:-1]: i for i, ln in enumerate(lines) if ln.endswith(":")}
    fma = re.compile(r"fma\.rn\.f32x2 %(\d+), (gg|ff)(\d), w(\d+), %(\d+);")
    # (the last chunk of a window of 6 or more comes through xl: its class address, or a broadcast dummy when no direction reads it)
    ld = re.compile(r"(@ploop )?ld\.shared\.v2\.b64 \{w(\d+), w(\d+)\}, \[(?:o([ab])(\d)|xl)\+(\d+)\];")
    for r in range(4):
        for D in range(kmax + 1):
            start = labels[f"B{r}_{D}"]
            seen_g, seen_f, reloaded = set(), set(), set()
            for ln in lines[start + 1:]:
                if ln.endswith(":") and re.match(r"B\d_\d+:|TAIL:|SW_\d+:|S\d_", ln) and not ln.startswith("SW_"):
                    break
                if ln.startswith("bra.uni TAIL") or ln.startswith("@ploop bra.uni TOP"):
                    break
                m = fma.match(ln)
                if m:
                    dst, kind, rr, w, src = int(m[1]), m[2], int(m[3]), int(m[4]), int(m[5])
                    if rr != r:
                        break                                        # fell through into the next direction's body
                    k = dst - 8 * r
                    assert dst == src and 0 <= k < 8                 # accumulates in place into acc[r][k]
                    assert w == D + k + (1 if kind == "gg" else 0)   # g * s[i+1], f * s[i] at this delta
                    assert w not in reloaded, (name, nch, r, D, ln)  # never reads a register the next window already owns
                    (seen_g if kind == "gg" else seen_f).add(k)
                    if kind == "ff":
                        assert k in seen_g                           # acc = fma(f, s[i], fma(g, s[i+1], acc))
                    continue
                m = ld.match(ln)
                if m and r == 3 and m[1]:                            # pipelined reload of a window chunk inside B3
                    a, b, imm = int(m[2]), int(m[3]), int(m[6])
                    chunk = a // 2
                    assert b == a + 1 and a % 2 == 0 and imm == 16 * (chunk + (chunk >> 2))
                    assert (int(m[5]) == chunk & 3) if m[5] is not None else (chunk == nch - 1 and nch >= 6)
                    reloaded.update((a, b))
            assert seen_g == set(range(8)) and seen_f == set(range(8)), (name, nch, r, D)
            if r == 3 and reloaded:
                assert reloaded == set(range(2 * nch))               # the whole next window, each chunk once
