"""Per-source-line totals of an ncu report's SASS page (development aid).
    python scripts/ncu_lines.py report.ncu-rep 'regex:k_nms' extract.sm_100a.cubin [launch_skip] [top]
Joins `ncu --page source --csv` (per-instruction executed counts and stall samples) with the line table of the
cubin (`nvdisasm -g`), and prints the source lines of csrc/*.cu that executed the most warp instructions."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def main():
    rep, kern, cubin = sys.argv[1:4]
    skip = sys.argv[4] if len(sys.argv) > 4 else "0"
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--launch-skip", skip,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    name = rows[0][1]
    head = rows[1]
    ix = {h: i for i, h in enumerate(head)}
    mangled = None
    # function name -> find in nvdisasm by matching the demangled prefix
    short = re.sub(r"\(.*", "", name).replace("void ", "").strip()
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    # split per function
    line_of = {}
    cur_fn, cur_line, cur_file = None, None, None
    want = None
    for ln in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1)
            dem = subprocess.run(["cu++filt", cur_fn], capture_output=True, text=True).stdout.strip()
            want = dem.replace("void ", "").startswith(short.split("<")[0]) and (("<" not in short) or short.replace(" ", "").replace("(int)", "") in dem.replace(" ", "").replace("(int)", ""))
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_file, cur_line = m.group(1), int(m.group(2))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and want:
            line_of[int(m.group(1), 16)] = (cur_file, cur_line)
    inst, samp, thr = defaultdict(float), defaultdict(float), defaultdict(float)
    tot_i = tot_s = 0.0
    base = None
    for r in rows[2:]:
        if len(r) < len(head) - 2 or not r[0].startswith("0x"):
            continue
        addr = int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]])
        if base is None:
            base = addr
        key = line_of.get(addr - base, ("?", 0))
        i = float(r[ix["Instructions Executed"]] or 0)
        s = float(r[ix["# Samples"]] or 0)
        t = float(r[ix["Thread Instructions Executed"]] or 0)
        inst[key] += i; samp[key] += s; thr[key] += t
        tot_i += i; tot_s += s
    print(f"{name}: {tot_i:.0f} warp instructions, {tot_s:.0f} samples")
    src_cache = {}
    for key in sorted(inst, key=lambda k: -samp[k])[:top]:
        f, l = key
        if f not in src_cache:
            try:
                src_cache[f] = open(f).read().splitlines()
            except Exception:
                src_cache[f] = []
        text = src_cache[f][l - 1].strip()[:110] if 0 < l <= len(src_cache[f]) else ""
        print(f"{100 * inst[key] / tot_i:5.1f}% inst {100 * samp[key] / max(tot_s, 1):5.1f}% samp  thr/inst {thr[key] / max(inst[key], 1):4.1f}  {f.split('/')[-1]}:{l}  {text}")


if __name__ == "__main__":
    main()
