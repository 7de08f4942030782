#!/usr/bin/env python
"""Compare the SASS of every kernel in build/obj/*.o against the same objects built from another commit.

    python tools/sass_diff.py <base-commit>

Used when a change must provably leave the measured kernels untouched (e.g. work done without GPU access): a kernel counts
as identical when its instruction stream (addresses and operands included, encodings stripped) is the same.  New template
instantiations show up under "new".  Needs only nvcc / cuobjdump (no GPU)."""
import hashlib
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJECTS = ["api_core", "api_sampler", "api_tc", "api_vit", "api_post"]


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    res, name, buf = {}, None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                res[name] = hashlib.md5("\n".join(buf).encode()).hexdigest()
            name, buf = m.group(1), []
        elif re.match(r"^\s+/\*[0-9a-f]{4,6}\*/", line):
            buf.append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip())
    if name:
        res[name] = hashlib.md5("\n".join(buf).encode()).hexdigest()
    return res


def norm(name):  # anonymous-namespace hashes differ between source trees
    return re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+", "ANON", name)


def main():
    base = sys.argv[1]
    with tempfile.TemporaryDirectory() as tmp:
        tree = os.path.join(tmp, "base")
        subprocess.run(["git", "-C", ROOT, "worktree", "add", "-q", tree, base], check=True)
        try:
            subprocess.run([sys.executable, "-m", "posediffusion_b200.build"], cwd=tree, check=True, capture_output=True)
            subprocess.run([sys.executable, "-m", "posediffusion_b200.build"], cwd=ROOT, check=True, capture_output=True)
            changed = 0
            for obj in OBJECTS:
                a = {norm(k): v for k, v in kernels(os.path.join(tree, "build", "obj", obj + ".o")).items()}
                b = {norm(k): v for k, v in kernels(os.path.join(ROOT, "build", "obj", obj + ".o")).items()}
                diff = sorted(k for k in a if k in b and a[k] != b[k])
                gone = sorted(k for k in a if k not in b)
                new = sorted(k for k in b if k not in a)
                # a kernel whose name changed (e.g. a new template parameter) but whose instruction stream did not
                renamed = {k: next((n for n in new if b[n] == a[k]), None) for k in gone}
                moved = sorted(k for k, n in renamed.items() if n)
                gone = [k for k in gone if not renamed[k]]
                new = [n for n in new if n not in renamed.values()]
                print(f"{obj}: {len(a)} kernels in {base}: {len(a) - len(diff) - len(gone) - len(moved)} identical, {len(moved)} renamed but "
                      f"identical, {len(diff)} changed, {len(gone)} removed; {len(new)} new")
                for k in moved:
                    print("   renamed, identical:", k[:110], "->", renamed[k][:110])
                for k in diff:
                    print("   changed:", k[:140])
                for k in gone:
                    print("   removed:", k[:140])
                for k in new:
                    print("   new:", k[:140])
                changed += len(diff) + len(gone)
            return 1 if changed else 0
        finally:
            subprocess.run(["git", "-C", ROOT, "worktree", "remove", "--force", tree])


if __name__ == "__main__":
    sys.exit(main())
