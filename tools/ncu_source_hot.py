"""Hot source lines of an ncu report: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv`,
then `python tools/ncu_source_hot.py src.csv [kernel-instance-index] [top]`.  Aggregates warp-stall samples per CUDA line."""
import csv
import sys


def main(path, instance=0, top=25):
    rows = list(csv.reader(open(path)))
    # split into blocks per (kernel instance, file): a block starts with "File Path"
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "File Path":
            cur = {"file": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    # kernel instances repeat the same file sequence; group by first file name recurrence
    first = blocks[0]["file"]
    inst, groups = -1, []
    for b in blocks:
        if b["file"] == first:
            inst += 1
            groups.append([])
        groups[inst].append(b)
    lines = []
    for b in groups[instance]:
        hdr = None
        for r in b["rows"]:
            if r and r[0] == "Line No":
                hdr = r
                continue
            if hdr is None or len(r) < 8 or not r[0].isdigit():
                continue
            if r[2] != "-":      # sass rows carry an address; cuda rows have "-"
                continue
            try:
                samples = int(r[4])
                inst_exec = int(r[7])
            except ValueError:
                continue
            lines.append((samples, inst_exec, b["file"].split("/")[-1], int(r[0]), r[1].strip()))
    tot = sum(l[0] for l in lines) or 1
    lines.sort(reverse=True)
    print("instance %d: %d stall samples" % (instance, tot))
    for s, ie, f, ln, src in lines[:top]:
        print("%5.1f%% %7d smp %9d inst  %s:%d  %s" % (100.0 * s / tot, s, ie, f, ln, src[:90]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 25)
