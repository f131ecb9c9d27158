"""Pack the reference's Python files (unmodified) into one archive for the GPU-side caller tests.  TEST INFRASTRUCTURE ONLY:
usage  python pack_reference_py.py <reference root> <out.zip>   (run by oracle/Makefile where the reference is mounted)."""
import os
import sys
import zipfile

root, out = sys.argv[1], sys.argv[2]
n = 0
with zipfile.ZipFile(out, "w", zipfile.ZIP_DEFLATED) as z:
    for d, _, files in os.walk(root):
        if os.path.relpath(d, root).split(os.sep)[0] == "docs":
            continue
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(d, f)
                z.write(p, os.path.relpath(p, root))
                n += 1
print("oracle: packed %d reference python files into %s" % (n, out))
