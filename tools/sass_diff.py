"""Per-kernel SASS comparison of two builds of the CUDA library (run in the build container; no GPU needed).

  python tools/sass_diff.py OLD_CSRC_DIR [NEW_CSRC_DIR]      # directories holding the *.o files of a `make`

For every kernel of grid / neighbors / solver / level / adapt / dist / capi it prints whether the instruction text
(`cuobjdump -sass`, addresses and encodings stripped) is identical in both builds, how many instructions differ otherwise,
and which kernels exist on one side only; a kernel whose only
differences are constant-bank offsets of its parameters (a by-value parameter struct grew) is reported as such.  Template instantiations whose parameter list grew by trailing defaulted
arguments (k_sweep<P, HMWIN, PEER, W2020> -> k_sweep<..., R4 = false>) are paired with their old selves.
Used to show that kernels measured on hardware are still the ones in the library after code was added around them
(profiles/r1_sass_audit.md).
"""
import difflib
import os
import re
import subprocess
import sys


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    res, name, body = {}, None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                res[name] = body
            name, body = m.group(1), []
        elif name:
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", line)
            if m:
                body.append(m.group(1))
    if name:
        res[name] = body
    return res


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", "").replace("void ", "")) for d in out]


def main():
    old = sys.argv[1]
    new = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "adaptive-sph_b200", "csrc")
    for unit in ("grid", "neighbors", "solver", "level", "adapt", "dist", "capi"):
        fa, fb = functions(os.path.join(old, unit + ".o")), functions(os.path.join(new, unit + ".o"))
        a = dict(zip(demangle(list(fa)), fa.values()))
        b = dict(zip(demangle(list(fb)), fb.values()))
        paired = set()
        print(f"{unit}.o: {len(a)} -> {len(b)} kernels")
        for name, body in sorted(a.items()):
            partner = name if name in b else next((k for k in b if k.startswith(name[:-1] + ", ") and set(t.strip() for t in k[len(name) - 1:-1].split(",")) <= {"", "false"}), None)
            if partner is None:
                print(f"  removed    {name}")
                continue
            paired.add(partner)
            if body == b[partner]:
                print(f"  identical  {name}  ({len(body)} instructions)")
            elif [re.sub(r"c\[0x0\]\[0x[0-9a-f]+\]", "c[0x0][.]", l) for l in body] == [re.sub(r"c\[0x0\]\[0x[0-9a-f]+\]", "c[0x0][.]", l) for l in b[partner]]:
                print(f"  identical  {name}  ({len(body)} instructions; kernel-parameter offsets moved: a by-value struct grew)")
            else:
                d = [l for l in difflib.unified_diff(body, b[partner], lineterm="", n=0) if l[:1] in "+-" and l[:3] not in ("---", "+++")]
                print(f"  CHANGED    {name}  ({len(body)} -> {len(b[partner])} instructions, {len(d)} differing lines)")
        for name in sorted(set(b) - paired):
            print(f"  new        {name}")


if __name__ == "__main__":
    main()
