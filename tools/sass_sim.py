#!/usr/bin/env python3
"""Development tool: crude in-order issue model of ONE warp over a SASS loop body (register / predicate dependencies, fixed
latencies), to compare how well the band loop of a kernel variant is interleaved before spending GPU time.
  python tools/sass_sim.py variants/libwgk_f1.so _ZN3wgk15k_cells_pre_tpcE9WgkParamsiii"""
import re
import subprocess
import sys

LAT = {"D": 9, "I2F": 18, "F2F": 10, "MUFU": 18, "LDS": 29, "LDG": 600, "LDC": 30}
ISSUE = {"D": 2, "I2F": 4, "F2F": 2, "MUFU": 4, "LDS": 2, "LDGSTS": 2, "STG": 2}


def kernel_sass(lib, fn):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    take, rows = False, []
    for l in out.split("\n"):
        if "Function :" in l:
            take = l.strip().endswith(fn)
            continue
        if take:
            m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
            if m:
                rows.append((int(m.group(1), 16), m.group(2).strip()))
    return rows


def regs_of(tok, wide):
    out = []
    for m in re.finditer(r"\b(U?R)(\d+)\b", tok):
        n = int(m.group(2))
        out.append(f"{m.group(1)}{n}")
        if wide:
            out.append(f"{m.group(1)}{n + 1}")
    for m in re.finditer(r"\b(U?P)(\d)\b", tok):
        out.append(f"{m.group(1)}{m.group(2)}")
    return out


def parse(ins):
    pred = None
    m = re.match(r"@(!?)(U?P\d)\s+(.*)", ins)
    if m:
        pred, ins = m.group(2), m.group(3)
    op, _, rest = ins.partition(" ")
    ops = [o.strip() for o in rest.split(",")] if rest else []
    base = op.split(".")[0]
    wide = base in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX") or ".64" in op
    dst, src = [], []
    if base in ("STG", "STS", "ST", "BRA", "DEPBAR", "LDGSTS", "BAR", "EXIT", "WARPSYNC", "BSYNC", "BSSY", "NOP", "RED"):
        for o in ops:
            src += regs_of(o, ".64" in op and base != "LDGSTS")
    else:
        ndst = 1
        if base in ("DSETP", "ISETP", "FSETP", "PLOP3"):
            ndst = 2
        if base in ("IADD3", "LEA", "IMAD") and len(ops) > 1 and re.match(r"^P\d$", ops[1]):
            ndst = 2
        for k, o in enumerate(ops):
            if k < ndst:
                w = wide and base not in ("DSETP",)
                if base == "I2F" and ".F64" in op:
                    w = True
                dst += regs_of(o, w)
            else:
                w = wide
                if base == "I2F":
                    w = False
                src += regs_of(o, w)
    if pred:
        src.append(pred)
    return base, op, dst, src


def simulate(body, iters=3):
    ready, t, per_iter = {}, 0, []
    for it in range(iters):
        t0 = t
        for _, ins in body:
            base, op, dst, src = parse(ins)
            key = "D" if base in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX") else base
            start = max([t] + [ready.get(r, 0) for r in src if r not in ("RZ", "URZ", "PT", "UPT")])
            lat = LAT.get(key, 5)
            for r in dst:
                ready[r] = start + lat
            t = start + ISSUE.get(key, 1)
        per_iter.append(t - t0)
    return per_iter


if __name__ == "__main__":
    rows = kernel_sass(sys.argv[1], sys.argv[2])
    addr = {a: i for i, (a, _) in enumerate(rows)}
    # loops = backward branches; report every loop body that contains LDS (the staged band loop) or is long
    for i, (a, ins) in enumerate(rows):
        m = re.search(r"BRA.*?0x([0-9a-f]+)", ins)
        if m and m.group(1):
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                body = rows[addr[tgt]:i + 1]
                nd = sum(1 for _, x in body if re.match(r"(@!?U?P\d\s+)?D(ADD|MUL|FMA|SETP)", x))
                nl = sum(1 for _, x in body if "LDS" in x)
                if len(body) > 60:
                    pi = simulate(body)
                    print(f"loop {tgt:#x}..{a:#x}: {len(body)} instr, {nd} FP64, {nl} LDS; modelled cycles per iteration {pi}")
