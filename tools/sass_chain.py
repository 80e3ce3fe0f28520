"""Static single-warp timing model of a SASS region (no GPU needed).

The LunarLander step kernel is ONE warp's dependency chain per SM sub-partition (3.8 % of the warp slots, profiles/r2/
lunar_step_ncu_summary.md): no other warp hides its stalls, so the time of a straight-line region is what ptxas wrote into the
instructions' control fields — the stall count after each issue plus the waits on the six scoreboard barriers of variable-latency
operations.  This tool decodes those fields from `cuobjdump -sass` (sm_100: bits 105..125 of the 128-bit word: stall[4] yield[1]
write-barrier[3] read-barrier[3] wait-mask[6] reuse[4]) and replays a region in order.

  python tools/sass_chain.py OBJ --func lunar_step --loops            # backward branches = loops, with body size / stall sum
  python tools/sass_chain.py OBJ --func lunar_step --range 0x1a00 0x2400 [--taken 0x1b40,...] [--dump]

`--range A B` replays from address A until the instruction at B has issued, following unconditional branches, taking the
conditional branches listed in --taken and falling through the others; it prints issue cycles = the model's time for one pass.
Variable-latency results are modelled with fixed figures (LAT below; shared/local loads as L1 hits) — the model ranks variants of
the same loop, it does not replace a measurement.
"""
import argparse
import re
import subprocess
import sys

LAT = {"MUFU": 18, "LDL": 33, "LDS": 30, "LDC": 30, "LDCU": 30, "LDG": 350, "I2F": 14, "F2I": 14, "F2F": 14, "I2FP": 6, "F2FP": 6,
       "DADD": 10, "DMUL": 10, "DFMA": 10, "DSETP": 14, "S2R": 25, "S2UR": 25, "CS2R": 25, "SHFL": 25, "ATOM": 400, "ATOMG": 400,
       "RED": 30, "STL": 10, "STG": 10, "STS": 10, "POPC": 14, "FLO": 14, "BREV": 14, "R2UR": 14, "VOTE": 14, "VOTEU": 14,
       "IMAD.WIDE": 6, "DEFAULT": 12}


def disasm(obj):
    return subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout


def parse(text, func):
    """-> list of dict(addr, op, txt, stall, yld, wr, rd, wait) for the first function whose name contains `func`."""
    out, on = [], False
    lines = text.split("\n")
    i = 0
    ins_re = re.compile(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*;\s*/\* 0x([0-9a-f]{16}) \*/")
    hi_re = re.compile(r"^\s+/\* 0x([0-9a-f]{16}) \*/")
    while i < len(lines):
        l = lines[i]
        if "Function :" in l:
            if on:
                break
            on = func in l
        elif on:
            m = ins_re.match(l)
            if m and i + 1 < len(lines):
                h = hi_re.match(lines[i + 1])
                if h:
                    hi = int(h.group(1), 16)
                    txt = m.group(2)
                    body = re.sub(r"^@!?U?P\d+\s+", "", txt)
                    out.append(dict(addr=int(m.group(1), 16), txt=txt, op=body.split()[0] if body else "", pred=txt.startswith("@"),
                                    stall=(hi >> 41) & 0xF, yld=(hi >> 45) & 1, wr=(hi >> 46) & 7, rd=(hi >> 49) & 7,
                                    wait=(hi >> 52) & 0x3F))
                    i += 1
        i += 1
    return out


def latency(op):
    if op in LAT:
        return LAT[op]
    base = op.split(".")[0]
    return LAT.get(base, LAT["DEFAULT"])


def branch_target(ins):
    m = re.search(r"\b(?:BRA|BRA\.U|BRA\.DIV)\S*\s+(?:[!U]*P\d+,\s*)?(?:`\([^)]*\)|0x([0-9a-f]+))", ins["txt"])
    m2 = re.search(r"0x([0-9a-f]+)\s*$", ins["txt"])
    if ins["op"].startswith("BRA") and m2:
        return int(m2.group(1), 16)
    return None


def loops(code):
    res = []
    for k, ins in enumerate(code):
        t = branch_target(ins)
        if t is not None and t <= ins["addr"]:
            body = [c for c in code if t <= c["addr"] <= ins["addr"]]
            res.append((t, ins["addr"], len(body), sum(max(c["stall"], 1) for c in body), ins["txt"]))
    return res


def replay(code, a, b, taken=(), dump=False, max_steps=200000):
    by_addr = {c["addr"]: k for k, c in enumerate(code)}
    k = by_addr[a]
    t = 0
    ready = [0] * 6
    n = 0
    waited = 0
    while n < max_steps:
        ins = code[k]
        t0 = t
        for bit in range(6):
            if ins["wait"] >> bit & 1:
                t = max(t, ready[bit])
        waited += t - t0
        if dump:
            print(f"{t:7d} {'+%d' % (t - t0) if t > t0 else '':>5s} {ins['addr']:06x} s{ins['stall']:<2d} w{ins['wr']} r{ins['rd']} m{ins['wait']:02x}  {ins['txt']}")
        if ins["wr"] != 7:
            ready[ins["wr"]] = t + latency(ins["op"])
        if ins["rd"] != 7:
            ready[ins["rd"]] = max(ready[ins["rd"]], t + 6)
        t += max(ins["stall"], 1)
        n += 1
        if ins["addr"] == b:
            break
        tgt = branch_target(ins)
        if tgt is not None and (not ins["pred"] or ins["addr"] in taken):
            k = by_addr[tgt]
        elif ins["op"] in ("EXIT", "RET"):
            break
        else:
            k += 1
    return t, n, waited


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("--func", required=True)
    ap.add_argument("--loops", action="store_true")
    ap.add_argument("--range", nargs=2)
    ap.add_argument("--taken", default="")
    ap.add_argument("--dump", action="store_true")
    a = ap.parse_args()
    code = parse(disasm(a.obj), a.func)
    print(f"{len(code)} instructions in *{a.func}*", file=sys.stderr)
    if a.loops:
        for t, e, n, s, txt in loops(code):
            print(f"loop {t:06x}..{e:06x}: {n:5d} instructions, stall sum {s:6d}   {txt}")
    if a.range:
        taken = {int(x, 16) for x in a.taken.split(",") if x}
        cyc, n, waited = replay(code, int(a.range[0], 16), int(a.range[1], 16), taken, a.dump)
        print(f"replay {a.range[0]}..{a.range[1]}: {n} instructions issued, {cyc} cycles ({waited} waiting on scoreboards)")


if __name__ == "__main__":
    main()
