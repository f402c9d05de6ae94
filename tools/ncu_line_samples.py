"""Per-source-line warp-state sample counts out of an .ncu-rep captured with --import-source on (ncu --page source):
usage: python tools/ncu_line_samples.py report.ncu-rep [top_n]   ->  "samples  file:line  source" rows, most sampled first,
then the mbarrier try_wait sites in SASS with the barrier offset they poll (which hand-shake a warp is waiting in)."""
import csv
import io
import subprocess
import sys


def page(rep, extra):
    return list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + extra,
                                                      capture_output=True, text=True).stdout)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = page(rep, ["--print-source", "cuda,sass"])
    cur, isamp, out, total = None, None, [], 0
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            isamp = r.index("# Samples")
        elif r and r[0] not in ("", "Function Name") and isamp is not None and len(r) > isamp:
            try:
                n = int(r[isamp])
            except ValueError:
                continue
            out.append((n, "%s:%s" % (cur, r[0]), r[1].strip()[:110]))
            total += n
    # inlined helpers are listed under their own file AND roll up into the caller's line: the total counts both
    print("# %s" % rep)
    print("# samples per CUDA source line (inlined helper lines also roll up into their call sites)")
    for n, where, src in sorted(out, reverse=True)[:top]:
        print("%7d  %-28s %s" % (n, where, src))
    sass = page(rep, [])
    hdr = sass[1]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = sass[2:]
    print("# mbarrier try_wait sites (SASS): executed, samples on the wait loop, instruction")
    for i, r in enumerate(data):
        if "TRYWAIT" in r[isrc]:
            n = sum(int(data[k][isamp] or 0) for k in range(i, min(i + 3, len(data))))
            if n >= 100:
                print("%9s %7d  %s" % (r[iex], n, r[isrc].strip()[:90]))
    print("# fences / remote arrives (SASS): executed, samples, instruction")
    for r in data:
        s = r[isrc]
        if ("MEMBAR" in s or "ERRBAR" in s or "SYNCS.ARRIVE" in s or "STAS" in s) and int(r[isamp] or 0) >= 40:
            print("%9s %7s  %s" % (r[iex], r[isamp], s.strip()[:90]))


if __name__ == "__main__":
    main()
