"""Dynamic SASS opcode histogram of an `ncu --page source --csv --print-source sass` export:
    python tools/ncu_sass_hist.py file.csv [units]    (units = frames or words the launch processed)"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ia = hdr.index("Source"); ie = hdr.index("Instructions Executed"); iw=hdr.index("L1 Wavefronts Shared"); istall=hdr.index("# Samples")
frames = float(sys.argv[2]) if len(sys.argv)>2 else 203420
h = collections.Counter(); wf=collections.Counter(); samp=collections.Counter()
tot=0
for r in rows[2:]:
    s = r[ia].strip(); n = int(r[ie] or 0)
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', s)
    op = m.group(2) if m else s
    h[op]+=n; tot+=n; wf[op]+=int(r[iw] or 0); samp[op]+=int(r[istall] or 0)
print("total inst/frame %.1f"%(tot/frames))
ts=sum(samp.values())
for op,n in h.most_common(40):
    print("%-12s %8.1f /frame  wavefronts/frame %7.1f  samples %5.1f%%"%(op, n/frames, wf[op]/frames, 100*samp[op]/ts))
