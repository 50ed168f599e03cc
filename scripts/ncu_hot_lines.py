"""Where a kernel's time goes, by source function: aggregates the source page of an ncu report (captured with
`--import-source on`, kernels built with -lineinfo) into stall samples and executed warp instructions per enclosing function
of one .cu/.cuh file, plus the hottest source lines.

  python scripts/ncu_hot_lines.py gpurun_out/r2x/ncu_lander.ncu-rep gymnasium_b200/csrc/lunarlander.cu [top]
"""
import csv
import io
import re
import subprocess
import sys


def functions_of(path):
    """[(first line, name)] of the function definitions of a C++ source (a line that starts at column 0 and opens a body)."""
    out = []
    pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:extern \"C\" )?(?:static\s+)?(?:DI|__device__|__global__|__host__|inline|static|void|int|float|double|bool|V2|Rot|Xf)[^;]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;]*$")
    with open(path) as f:
        lines = f.read().split("\n")
    for i, ln in enumerate(lines, 1):
        if ln[:1] in (" ", "\t", "/", "#", "}", "") or ln.startswith("struct") or ln.startswith("constexpr") or ln.startswith("namespace"):
            continue
        m = pat.match(ln)
        if m and "=" not in ln.split("(")[0]:
            out.append((i, m.group(1)))
    return out


def main():
    rep, src = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True,
                         text=True).stdout
    base = src.split("/")[-1]
    per_line, sass_rows = {}, []
    cur_file, header, cur_line, cur_text = None, None, None, ""
    total_s = total_i = 0
    for row in csv.reader(io.StringIO(raw)):
        if not row:
            continue
        if row[0] == "File Path":
            cur_file = row[1]
            continue
        if row[0] == "Function Name":
            continue
        if row[0] == "Line No":
            header = row
            i_samp, i_inst = header.index("# Samples"), header.index("Instructions Executed")
            continue
        if header is None or len(row) <= max(i_samp, i_inst):
            continue
        if row[0] != "":  # a CUDA source line; its SASS rows follow with an empty first column
            cur_line, cur_text = int(row[0]), row[1]
            continue
        try:
            s, n = int(row[i_samp] or 0), int(row[i_inst] or 0)
        except ValueError:
            continue
        total_s += s
        total_i += n
        try:
            sass_rows.append((int(row[2], 16), cur_file.split("/")[-1] if cur_file else "?", cur_line, s, n))
        except ValueError:
            pass
        key = (cur_file.split("/")[-1] if cur_file else "?", cur_line)
        e = per_line.setdefault(key, [0, 0, cur_text])
        e[0] += s
        e[1] += n
    fns = functions_of(src)

    def fn_of(fname, line):
        if fname != base or not fns or line < fns[0][0]:
            return None  # an inlined helper (vector operators, intrinsics): attributed to the code around it, by address
        name = None
        for first, fn in fns:
            if first <= line:
                name = fn
            else:
                break
        return name

    per_fn = {}
    last = "?"
    for _addr, fname, line, s, n in sorted(sass_rows):
        name = fn_of(fname, line)
        if name is None:
            name = last
        last = name
        e = per_fn.setdefault(name, [0, 0])
        e[0] += s
        e[1] += n
    print(f"report {rep}: {total_s} stall samples, {total_i} executed warp instructions (inlined code is attributed to the line it came from)")
    print(f"\n| function ({base}) | samples | % | warp instructions | % |\n|---|---|---|---|---|")
    for name, (s, n) in sorted(per_fn.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"| {name} | {s} | {100.0 * s / max(total_s, 1):.1f} | {n} | {100.0 * n / max(total_i, 1):.1f} |")
    print("\n| line | samples | % | warp instructions | source |\n|---|---|---|---|---|")
    for (fname, line), (s, n, t) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"| {fname}:{line} | {s} | {100.0 * s / max(total_s, 1):.1f} | {n} | `{t.strip()[:90]}` |")


if __name__ == "__main__":
    main()
