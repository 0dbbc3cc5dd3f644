"""List the loops (backward branches) of one kernel in the built library with their instruction mix (development aid).
usage: python scripts/sass_loops.py <mangled-name-substring> [min_len]"""
import re, subprocess, sys
from collections import Counter
pat = sys.argv[1]; minlen = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["cuobjdump", "-sass", "bsalign_b200/libbsalign_b200.so"], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
ALU = {"VIADDMNMX", "VIMNMX", "VIMNMX3", "PRMT", "SHF", "LOP3", "IADD3", "ISETP", "VIADD", "SEL", "MOV", "IABS", "LEA", "FLO", "POPC", "IMNMX"}
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if pat not in name: continue
    ins = [(int(m.group(1), 16), m.group(2)) for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", f)]
    print(name, len(ins), "instructions")
    for addr, txt in ins:
        m = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", txt)
        if m and int(m.group(1), 16) < addr:
            tgt = int(m.group(1), 16)
            body = [re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for a, t in ins if tgt <= a <= addr]
            if len(body) < minlen or len(body) > 400: continue
            c = Counter(body)
            print("  loop %#x-%#x n=%d alu=%d" % (tgt, addr, len(body), sum(v for k, v in c.items() if k in ALU)), dict(c.most_common(16)))
    break
