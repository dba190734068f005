"""Device code size per function of libarmour_b200.so (the reach-set kernel is instruction-fetch bound: track it)."""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "armour_b200/libarmour_b200.so"
out = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
rows = []
for line in out.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+)\s+\S+\s+\S+\s+(\S+)", line)
    if m and m.group(3) in ("0x2", "0x12", "0x22"):
        rows.append((int(m.group(2), 16), m.group(4)))
names = subprocess.run(["c++filt"], input="\n".join(n.split("$")[-1] for _, n in rows), capture_output=True, text=True).stdout.splitlines()
for (sz, raw), nm in sorted(zip(rows, names), key=lambda x: -x[0][0]):
    print(f"{sz:8d} B  {sz // 16:6d} instr  {nm[:110]}")
