#!/usr/bin/env python
"""Print the instruction mix of the innermost (hottest) loop of each kernel in a cubin / .so / executable.

Usage: sass_loop.py <binary> [kernel-substring]
A loop is a backward branch; for every kernel the loop with the most instructions is reported.
Used to check instruction budgets per cell before spending GPU time ("Instruction budget").
"""
import re, subprocess, sys, collections

def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name: yield name, body
            name, body = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            body.append((int(m.group(1), 16), m.group(2).strip()))
    if name: yield name, body

def loops(body):
    res = []
    for idx, (addr, ins) in enumerate(body):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", ins)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= addr:
                res.append([i for a, i in body if tgt <= a <= addr])
    return res

def main():
    path = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ""
    for name, body in kernels(path):
        if sub not in name: continue
        ls = loops(body)
        if not ls:
            print(f"{name}: no loop ({len(body)} instrs)"); continue
        for l in sorted(ls, key=len, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 1]:
            c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0].split(".")[0] for i in l)
            print(f"{name}: loop {len(l)} instrs: " + " ".join(f"{k}={v}" for k, v in c.most_common()))

if __name__ == "__main__":
    main()
