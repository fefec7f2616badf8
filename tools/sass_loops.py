"""Static view of a kernel's SASS: for every backward branch (a loop), the opcode histogram of its body.

usage: python tools/sass_loops.py <object-or-so> <mangled-or-substring-of-demangled-name> [--min N]
Prints the loops sorted by body length.  Used to check instruction counts per solver step before spending GPU time."""
import collections
import re
import subprocess
import sys


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
        elif cur:
            body[cur].append(line)
    return body


def parse(lines):
    ins = []
    for l in lines:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            addr = int(m.group(1), 16)
            text = m.group(2).strip()
            pred = ""
            pm = re.match(r"(@!?U?P\d+)\s+(.*)", text)
            if pm:
                pred, text = pm.group(1), pm.group(2)
            op = text.split()[0]
            ins.append((addr, op, text, pred))
    return ins


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    minlen = int(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 20
    ks = kernels(obj)
    names = list(ks)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for n, d in zip(names, dem):
        if pat not in n and pat not in d:
            continue
        ins = parse(ks[n])
        print("==", d[:200])
        print("   total instructions:", len(ins))
        idx = {a: i for i, (a, _, _, _) in enumerate(ins)}
        loops = []
        for i, (a, op, text, pred) in enumerate(ins):
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", text)
                if m:
                    t = int(m.group(1), 16)
                    if t <= a and t in idx:
                        loops.append((idx[t], i))
        for s, e in sorted(loops, key=lambda x: x[0] - x[1]):
            if e - s + 1 < minlen:
                continue
            h = collections.Counter(re.sub(r"\..*", "", op) for _, op, _, _ in ins[s:e + 1])
            full = collections.Counter(op for _, op, _, _ in ins[s:e + 1])
            print(f"   loop {ins[s][0]:#x}..{ins[e][0]:#x}: {e - s + 1} instr")
            print("     ", ", ".join(f"{k} {v}" for k, v in h.most_common()))
            if "--full" in sys.argv:
                print("     ", ", ".join(f"{k} {v}" for k, v in full.most_common()))


if __name__ == "__main__":
    main()
