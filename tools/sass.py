"""Dump / summarise the SASS of one kernel of libvslam_b200.so:  python tools/sass.py k_blur [--full]"""
import collections, re, subprocess, sys
lib = "orb-slam3-rust_b200/vslam_b200/libvslam_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name = sys.argv[1]
blocks = re.split(r"\n\s*Function : ", out)
for b in blocks[1:]:
    fn = b.split("\n", 1)[0]
    if name not in fn: continue
    lines = [l for l in b.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    print(fn, len(lines), "instructions")
    if "--full" in sys.argv:
        for l in lines: print(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).rstrip())
    else:
        h = collections.Counter()
        for l in lines:
            t = l.split()[1:]
            op = t[0] if not t[0].startswith("@") else t[1]
            h[".".join(op.split(".")[:2]) if op.startswith("VIMNMX") else op.split(".")[0].rstrip(";")] += 1
        print(h.most_common(25))
