"""Static SASS of one kernel of the built library: instruction count, opcode histogram and the
step loop(s) (development aid; the executed-instruction numbers in profiles/ come from ncu).
usage: sass_static.py [lib.so] [mangled-name substring] [out.sass]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "hestonexotics_b200/lib/libhexo_gpu.so"
pat = sys.argv[2] if len(sys.argv) > 2 else "heston_qe_paths_kernelILi0ELi0ELi2ENS_7ShishuaELb0ELb0"
out = sys.argv[3] if len(sys.argv) > 3 else None
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
hit = [b for b in blocks if b.split("\n", 1)[0].find(pat) >= 0]
if not hit:
    sys.exit(f"no kernel matching {pat}")
body = hit[0]
if out:
    open(out, "w").write(body)
ins = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", body)
hist = collections.Counter(i.split(".")[0] for i in ins)
print(body.split("\n", 1)[0])
print("instructions:", len(ins))
fp64 = sum(v for k, v in hist.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print("FP64:", fp64, " MUFU:", hist["MUFU"], " FFMA2:", hist.get("FFMA2", 0), " BSSY:", hist.get("BSSY", 0))
print(", ".join(f"{k} {v}" for k, v in hist.most_common(40)))
