"""Attribute SASS instructions (and local-memory spill traffic) of one kernel to source lines.
    python tools/sass_lines.py <object.o> <kernel-name-substring> [top]
Needs -lineinfo objects; uses cuobjdump -xelf + nvdisasm -g (CPU only, no GPU needed)."""
import collections, os, re, subprocess, sys, tempfile
obj, pat = os.path.abspath(sys.argv[1]), sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(d, cub)], capture_output=True, text=True).stdout
sec, cur = None, None
tot, spill = collections.Counter(), collections.Counter()
for line in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        sec = m.group(1)
        continue
    if sec is None or pat not in sec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        tot[cur] += 1
        if re.search(r"\b(STL|LDL)\b", line):
            spill[cur] += 1
print(f"kernel ~{pat}: {sum(tot.values())} SASS instructions, {sum(spill.values())} local-memory (STL/LDL)")
print("-- most instructions by source line")
for k, v in tot.most_common(top):
    print(f"   {k[0]}:{k[1]:<5} {v:6d}   spill {spill.get(k, 0)}")
print("-- local-memory instructions by source line")
for k, v in spill.most_common(top):
    print(f"   {k[0]}:{k[1]:<5} {v:6d}")
