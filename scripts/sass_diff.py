"""Which kernels differ between two builds of libskyb200.so?  (No GPU needed.)

    python scripts/sass_diff.py old/libskyb200.so skyrl_b200/libskyb200.so

Compares the SASS of every kernel by name (instruction text without the address column).  Used to prove that a
change leaves the GPU-validated kernels byte-identical, and to list exactly what has to be re-validated.
"""
import collections
import hashlib
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    body, name = collections.defaultdict(list), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
        elif name and re.match(r"\s+(/\*[0-9a-f]{4}\*/|/\* 0x[0-9a-f]+ \*/|\.L_)", line):  # instruction and label lines only
            body[name].append(re.sub(r"/\*[0-9a-f]{4}\*/", "", line))
    return {k: (hashlib.md5("\n".join(v).encode()).hexdigest(), len(v)) for k, v in body.items()}


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    by_hash = {h: k for k, (h, _) in a.items()}
    same = 0
    for k in sorted(set(a) | set(b)):
        if k not in a:
            twin = by_hash.get(b[k][0])
            print(f"NEW      {k} ({b[k][1]} lines)" + (f"  == old {twin}" if twin else ""))
        elif k not in b:
            print(f"GONE     {k} ({a[k][1]} lines)")
        elif a[k][0] != b[k][0]:
            print(f"CHANGED  {k} ({a[k][1]} -> {b[k][1]} lines)")
        else:
            same += 1
    print(f"{same} kernels identical, {len(a)} in {sys.argv[1]}, {len(b)} in {sys.argv[2]}")


if __name__ == "__main__":
    main()
