"""Per-source-line static SASS summary of one kernel of a built library (no GPU needed):
   python scripts/sass_lines.py build/variants/v2.so replay_reg_kernelILi12ELb1 [--spills]
prints the spill instructions (STL/LDL) with their source lines, and the instruction count per source line."""
import collections, os, re, subprocess, sys, tempfile

lib, pat = sys.argv[1], sys.argv[2]
with tempfile.TemporaryDirectory() as d:
  subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
  cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
  txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
m = re.search(r"\.section\s+\.text\.(\S*%s\S*)," % re.escape(pat), txt)
i = m.start(); j = txt.find("\n//---------------------", i + 10)
line, per, spills, n = None, collections.Counter(), [], 0
for l in txt[i:j].splitlines():
  mm = re.search(r'//## File ".*", line (\d+)', l)
  if mm: line = int(mm.group(1)); continue
  if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
    n += 1; per[line] += 1
    if re.search(r"\b(STL|LDL)", l): spills.append((line, l.strip()[:90]))
print(m.group(1), "instructions:", n, "spill instructions:", len(spills))
for ln, l in spills: print("  spill @ line", ln, l)
if "--lines" in sys.argv:
  for ln, c in sorted(per.items(), key=lambda x: -x[1])[:40]: print(f"  line {ln}: {c}")
