#!/usr/bin/env python3
"""tools_sass_diff.py OLD.so NEW.so — which kernels of two builds of libvfd_dfsph.so differ, instruction for instruction
(cuobjdump -sass; addresses and encodings stripped, so code that merely moved compares equal apart from its branch targets).
Used to show what of the library at the end of a round is NOT the code that last ran on hardware (profiles/r02_sass_vs_validated.md):
    git worktree add /tmp/wt <commit>; (cd /tmp/wt && python -m vfd_b200.build --force); python tools_sass_diff.py /tmp/wt/vfd_b200/lib/libvfd_dfsph.so vfd_b200/lib/libvfd_dfsph.so"""
import difflib
import re
import subprocess
import sys


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    res, name, body = {}, None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                res[name] = body
            name, body = m.group(1), []
        elif name is not None:
            mm = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)
            if mm:
                body.append(re.sub(r"\s+", " ", mm.group(1)).strip())
    if name:
        res[name] = body
    return res


def demangled(name):
    try:
        return subprocess.run(["c++filt", "-p", name], stdout=subprocess.PIPE, text=True).stdout.strip() or name
    except OSError:
        return name


def main(old, new):
    a, b = kernels(old), kernels(new)
    same = [k for k in a if k in b and a[k] == b[k]]
    print("%d kernels in OLD, %d in NEW, %d identical instruction for instruction" % (len(a), len(b), len(same)))
    for k in sorted(set(a) | set(b)):
        if k not in a:
            print("only in NEW: %s (%d instructions)" % (demangled(k), len(b[k])))
        elif k not in b:
            print("only in OLD: %s" % demangled(k))
        elif a[k] != b[k]:
            # a moved call / branch target is not a change of the code: compare again with the targets masked
            mask = lambda body: [re.sub(r"0x[0-9a-f]+", "0x?", i) if re.match(r"(@!?U?P\d+ )?(CALL|BRA|BSSY|WARPSYNC|MOV R\d+, 0x)", i) else i for i in body]
            d = [l for l in difflib.unified_diff(mask(a[k]), mask(b[k]), lineterm="", n=0) if l[:1] in "+-" and l[:3] not in ("+++", "---")]
            print("differs: %s: %d -> %d instructions, %d changed lines once call/branch targets are masked" % (demangled(k), len(a[k]), len(b[k]), len(d)))
            for l in d[:24]:
                print("    " + l)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
