"""Summarise `nvcc -Xptxas -v` output: one line per kernel with registers, spills, stack (developer tool)."""
import re, subprocess, sys
text = sys.stdin.read()
name = None
for line in text.splitlines():
    m = re.search(r"Compiling entry function '([^']+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*$", "", name).replace("b200::", "").replace("void ", "")
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and name:
        stack, sst, sld = m.groups(); continue
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        print(f"{m.group(1):>4} regs  stack {stack:>4}  spill {sst}/{sld}  {name}")
        name = None
    if re.search(r"error|warning", line): print(line)
