"""Are two cubins the same device code, kernel by kernel?  (Refactors of the .cu/.cuh files are checked with this when no
GPU is at hand: `nvcc -cubin` both trees, then `python scripts/sass_same.py a.cubin b.cubin`.)"""
import re
import subprocess
import sys


def load(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    d, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); d[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
            d[cur].append(re.sub(r"/\*[0-9a-f]{4,5}\*/|/\* 0x[0-9a-f]+ \*/", "", line).strip())
    return d


a, b = load(sys.argv[1]), load(sys.argv[2])
only = sorted(set(a) ^ set(b))
diff = [k for k in a if k in b and a[k] != b[k]]
print(f"{len(a)} / {len(b)} kernels; only in one: {only}; different: {diff}")
sys.exit(1 if only or diff else 0)
