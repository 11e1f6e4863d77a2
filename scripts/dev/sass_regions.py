"""Static code-size profile of a kernel: SASS instructions per source region (nvdisasm -g -c output).
usage: sass_regions.py file.sass kernel_substring"""
import re, sys, collections
path, sub = sys.argv[1], sys.argv[2]
regions = [  # (file suffix, first line, last line, name) in track_kernel.cuh / pose_device.cuh
    ("track_kernel.cuh", 152, 196, "setup_pixel (phase A)"),
    ("track_kernel.cuh", 225, 357, "sample_step"),
    ("track_kernel.cuh", 358, 376, "SegmentLoop"),
    ("track_kernel.cuh", 377, 390, "huber"),
    ("track_kernel.cuh", 400, 532, "ldlt regs/rows + rcp_newton"),
    ("track_kernel.cuh", 533, 596, "build/factor block"),
    ("track_kernel.cuh", 597, 797, "gn_solve_step"),
    ("track_kernel.cuh", 832, 888, "wait_pass/ready"),
    ("track_kernel.cuh", 889, 899, "pose_records_block"),
    ("track_kernel.cuh", 917, 1156, "pass_finish"),
    ("track_kernel.cuh", 1158, 1276, "track_pass prologue"),
    ("track_kernel.cuh", 1277, 1396, "track_pass A/B/rows"),
    ("track_kernel.cuh", 1397, 1445, "track_pass C + patch cost"),
    ("track_kernel.cuh", 1446, 1507, "track_pass epilogue"),
    ("track_kernel.cuh", 1509, 1700, "sweep_kernel"),
    ("pose_device.cuh", 0, 10000, "pose_device"),
    ("mbavo_device.h", 0, 10000, "mbavo_device.h"),
]
cnt = collections.Counter(); inside = False; cur = ("?", 0); total = 0
for ln in open(path):
    if ln.startswith("//---") and ".text." in ln:
        inside = sub in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        total += 1
        name = "other:" + cur[0].split("/")[-1]
        for suf, a, b, n in regions:
            if cur[0].endswith(suf) and a <= cur[1] <= b:
                name = n; break
        cnt[name] += 1
print("total instructions", total, "=", total * 16 / 1024, "KB")
for k, v in cnt.most_common():
    print(f"{v:7d}  {v*16/1024:7.1f} KB  {k}")
