"""SASS of one kernel of a built library annotated with source lines: python scripts/sass_dump.py lib.so pattern lo hi"""
import os, re, subprocess, sys, tempfile
lib, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
with tempfile.TemporaryDirectory() as d:
  subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
  cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
  txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
m = re.search(r"\.section\s+\.text\.(\S*%s\S*)," % re.escape(pat), txt)
i = m.start(); j = txt.find("\n//---------------------", i + 10)
line = None
for l in txt[i:j].splitlines():
  mm = re.search(r'//## File ".*", line (\d+)', l)
  if mm: line = int(mm.group(1)); continue
  if line is not None and lo <= line <= hi and (re.match(r"\s+/\*[0-9a-f]{4,}\*/", l) or l.startswith('.L_')):
    print(f"{line:5d} {l.strip()[:110]}")
