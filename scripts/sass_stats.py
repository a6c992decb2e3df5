"""Static SASS statistics of librtgrff_b200.so: instructions per kernel and opcode mix of one kernel.

    python scripts/sass_stats.py [substring-of-mangled-name]
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "raytracinggrff_b200" / "librtgrff_b200.so"


def main():
    pat = sys.argv[1] if len(sys.argv) > 1 else None
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    name = None
    counts = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            counts[name] += 1
            ops[name][m.group(2).split(".")[0]] += 1
    for k, v in sorted(counts.items(), key=lambda kv: -kv[1]):
        if pat is None or pat in k:
            print(f"{v:7d}  {k}")
            if pat is not None:
                print("   ", ", ".join(f"{o}:{n}" for o, n in ops[k].most_common(25)))


if __name__ == "__main__":
    main()
