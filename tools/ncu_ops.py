"""opcode histogram (warp instructions) of every kernel in an .ncu-rep source page (development helper)"""
import collections, csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"] + sys.argv[2:], capture_output=True, text=True).stdout
name = None
by = collections.Counter()
def flush():
    if name and by:
        tot = sum(by.values())
        print(name[:90], "total warp inst", tot)
        print("   " + "  ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in by.most_common(22)))
for r in csv.reader(io.StringIO(raw)):
    if r and r[0] == "Kernel Name":
        flush()
        name, by = r[1], collections.Counter()
        hdr = None
        continue
    if r and r[0] == "Address":
        hdr = r
        ie, si = hdr.index("Instructions Executed"), hdr.index("Source")
        continue
    if name and hdr and len(r) > ie:
        try:
            n = int(r[ie])
        except ValueError:
            continue
        toks = r[si].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        by[op.split(".")[0]] += n
flush()
