"""Per-source-line instruction counts, stall samples and shared-memory wavefronts from an .ncu-rep captured with
--import-source on:  python tools/ncu_lines.py report.ncu-rep [top=40] [kernel-name substring]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    want = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    out, hdr, fname, func = [], None, "", ""
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif len(r) >= 2 and r[0] == "Function Name":
            func = r[1]
        elif len(r) > 8 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-" and want in func:     # a source line row (its SASS rows carry an address)
            ci, ct, cn = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
            cw, cx = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive")
            try:
                out.append((int(r[ci]), int(r[ct]), int(r[cn]), "%s:%s" % (fname, r[0]), r[1].strip()[:90], int(r[cw] or 0), int(r[cx] or 0)))
            except ValueError:
                pass
    tot = sum(o[0] for o in out)
    samp = sum(o[2] for o in out)
    wav = sum(o[5] for o in out)
    print("warp instructions %d, samples %d, shared-memory wavefronts %d (excessive %d)" % (tot, samp, wav, sum(o[6] for o in out)))
    for o in sorted(out, reverse=True)[:top]:
        print("%5.1f%% inst  %5.1f%% samp  thr/inst %4.1f  smem wav %5.1f%% (x%.2f)  %-18s %s" % (
            100.0 * o[0] / max(tot, 1), 100.0 * o[2] / max(samp, 1), o[1] / max(o[0], 1), 100.0 * o[5] / max(wav, 1),
            o[5] / max(o[5] - o[6], 1), o[3], o[4]))


if __name__ == "__main__":
    main()
