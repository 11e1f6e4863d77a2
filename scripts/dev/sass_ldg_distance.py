"""For every 128-bit texel gather (LDG.E.128.CONSTANT) of a kernel: how many instructions later is its result first read?
Follows the fall-through order; at a backward branch whose target lies before the load the search continues at the target
(the hot loops are single back-edge loops).  usage: sass_ldg_distance.py file.sass kernel_substring [stop_substring]"""
import re, sys
path, sub = sys.argv[1], sys.argv[2]
stop = sys.argv[3] if len(sys.argv) > 3 else None
ins = []; labels = {}; inside = False
for ln in open(path):
    if ln.startswith("//---") and ".text." in ln:
        inside = sub in ln and not (stop and stop in ln)
        continue
    if not inside: continue
    m = re.match(r"^(\.L_x_\d+):", ln)
    if m: labels[m.group(1)] = len(ins); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m: ins.append(m.group(2).strip())
def regs(tok):
    return [int(x) for x in re.findall(r"\bR(\d+)\b", tok)]
for k, s in enumerate(ins):
    if "LDG.E.128.CONSTANT" not in s: continue
    d0 = regs(s.split(",")[0])[0]; dest = set(range(d0, d0 + 4))
    pos = k + 1; dist = 0; seen_back = 0; first = None
    while pos < len(ins) and dist < 600:
        t = ins[pos]
        body = re.sub(r"^@!?U?P\d+\s+", "", t)
        op = body.split()[0]
        operands = body[len(op):]
        parts = operands.split(",")
        srcs = regs(",".join(parts[1:])) if not op.startswith(("ST", "BRA", "RED", "ATOM")) else regs(operands)
        if dest & set(srcs): first = (dist, t); break
        dst = regs(parts[0]) if parts and not op.startswith(("ST", "BRA")) else []
        # (a redefinition of a dest register before any read would end the search; not expected)
        m = re.search(r"BRA\s+`\((\.L_x_\d+)\)", t)
        if m and m.group(1) in labels and labels[m.group(1)] <= k and seen_back < 1 and not t.startswith("@!") :
            pass
        if m and m.group(1) in labels and labels[m.group(1)] <= k and seen_back < 1:
            # backward branch of the enclosing loop: assume taken
            pos = labels[m.group(1)]; seen_back += 1; continue
        pos += 1; dist += 1
    spill = sum(1 for t in ins[max(0,k-150):k+150] if t.split()[0] in ("STL","LDL") or " STL" in t or " LDL" in t)
    print(f"instr {k:5d}: {s[:60]:60s} first use after {first[0] if first else '>600':>4} instrs   ({first[1][:50] if first else ''})  local-mem ops within +-150: {spill}")
