#!/usr/bin/env python
"""Static SASS instruction counts per source line of one kernel (no GPU needed).

    python tools/sass_lines.py ps_raster raster_fwd6_kernelILi3ELb0 [--min 3]

Extracts the sm_100a cubin of pose_splatter_b200/csrc/<object>.o, disassembles it with `nvdisasm -g` (the library is built
with -lineinfo) and counts the instructions attributed to every source line of the first kernel whose mangled name
contains the pattern.  Inlined device functions are attributed to their own lines (ps_contract.cuh etc.).  The counts are
static: multiply by trip counts (bench.py's pair / chunk counters) for a dynamic estimate."""
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    min_count = int(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 1
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "pose_splatter_b200" / "csrc" / f"{obj}.o")], cwd=tmp, check=True,
                       stdout=subprocess.DEVNULL)
        cubin = next(Path(tmp).glob("*.cubin"))
        text = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], check=True, capture_output=True, text=True).stdout
    counts, ops = collections.Counter(), collections.defaultdict(collections.Counter)
    inside, cur, total, name = False, None, 0, None
    for line in text.splitlines():
        if line.startswith("//---") and ".text." in line:
            if inside:
                break
            inside = pat in line
            if inside:
                name = line.split(".text.")[1].split()[0]
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            counts[cur] += 1
            ops[cur][m.group(1).split(".")[0]] += 1
            total += 1
    print(f"# {name}: {total} SASS instructions (static), by source line")
    src_cache = {}
    for (fname, ln), c in sorted(counts.items()):
        if c < min_count:
            continue
        if fname not in src_cache:
            hits = list(ROOT.rglob(fname))
            src_cache[fname] = hits[0].read_text().splitlines() if hits else []
        src = src_cache[fname][ln - 1].strip()[:110] if 0 < ln <= len(src_cache[fname]) else ""
        top = " ".join(f"{k}{v}" for k, v in ops[(fname, ln)].most_common(4))
        print(f"{fname}:{ln}\t{c}\t[{top}]\t{src}")


if __name__ == "__main__":
    main()
