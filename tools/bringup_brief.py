import re, sys
for l in sys.stdin:
    l = l.rstrip()
    if l.startswith(("PASS", "FAIL")):
        name = l.split()[1]
        m = re.search(r'"ms": ([0-9.]+), "tflops": ([0-9.]+), "cublas_tflops": ([0-9.]+)', l)
        g = re.search(r'"ms": ([0-9.]+), "gbps": ([0-9.]+)', l)
        extra = f"ms={float(m.group(1)):.3f} tflops={float(m.group(2)):.0f} cublas={float(m.group(3)):.0f}" if m else \
            (f"ms={float(g.group(1)):.3f} gbps={float(g.group(2)):.0f}" if g else "")
        print(l.split()[0], name, extra, "" if l.startswith("PASS") else l[:600])
    elif l.startswith("SUMMARY"):
        print(l)
