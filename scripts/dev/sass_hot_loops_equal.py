"""Do two builds of a sweep unit have the same hot loops?  Compares, instruction by instruction, the windows around the texel
gathers of the Hessian pass (LDG.E.128.CONSTANT) and of the cost pass (LDG.E.CONSTANT) of the first kernel in both cubins
(ptxas was seen to schedule the sample loops differently after changes to cold code: profiles/r2_history.md).
usage: sass_hot_loops_equal.py a.cubin b.cubin"""
import re, subprocess, sys
def load(cubin):
    out = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout
    ins = []
    for l in out.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
        if m:
            s = re.sub(r"\.L_x_\d+", "L", m.group(1).strip())
            ins.append(re.sub(r"MOV (R\d+), 0x[0-9a-f]+", r"MOV \1, X", s))
    return ins
a, b = load(sys.argv[1]), load(sys.argv[2])
ok = True
for pat, lo, hi, name in ((r"^LDG\.E\.128\.CONSTANT", 200, 300, "Hessian-pass sample loops"), (r"^LDG\.E\.CONSTANT ", 300, 150, "cost-pass sample loop")):
    ia = [k for k, s in enumerate(a) if re.match(pat, s)]
    ib = [k for k, s in enumerate(b) if re.match(pat, s)]
    n = 6 if "128" in pat else 7
    wa, wb = a[ia[0] - lo:ia[n - 1] + hi], b[ib[0] - lo:ib[n - 1] + hi]
    same = wa == wb
    ok &= same
    print(f"{name}: {'identical' if same else 'DIFFERENT'} ({len(wa)} / {len(wb)} instructions; gathers at {ia[:n]} / {ib[:n]})")
sys.exit(0 if ok else 1)
