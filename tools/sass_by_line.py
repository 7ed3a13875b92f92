"""Static SASS instruction counts per source line of one kernel (no GPU needed):

    python tools/sass_by_line.py semi_detr_b200/lib/obj/msda_backward.o msda_bwd_d32_kernelILi128ELi4ELi8ELi4ELb1E [top]

Uses `cuobjdump -xelf` + `nvdisasm --print-line-info` (objects are built with -lineinfo).  The unrolled hot loop of the
MSDA kernels dominates their SASS, so static counts track the dynamic instruction mix closely; per-item setup code
(tile cursor, 64-bit divisions) is over-represented."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main(obj, kernel_substr, top=40):
    obj = os.path.abspath(obj)
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        text = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], capture_output=True, text=True,
                              check=True).stdout
    inside, cur = False, None
    cnt, ops = collections.Counter(), collections.defaultdict(collections.Counter)
    for line in text.splitlines():
        if line.startswith(".text."):
            inside = kernel_substr in line
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            cnt[cur] += 1
            ops[cur][m.group(2).split(".")[0]] += 1
    total = sum(cnt.values())
    print(f"# {kernel_substr}: {total} SASS instructions")
    cache = {}
    for (f, ln), c in cnt.most_common(top):
        if f not in cache:
            try:
                cache[f] = open(f).read().splitlines()
            except OSError:
                cache[f] = []
        src = cache[f][ln - 1].strip()[:80] if ln - 1 < len(cache[f]) else ""
        mix = ",".join(f"{k}{v}" for k, v in ops[(f, ln)].most_common(4))
        print(f"{c:5d} {100 * c / total:4.1f}%  {os.path.basename(f)}:{ln:<4d} {src}   [{mix}]")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
