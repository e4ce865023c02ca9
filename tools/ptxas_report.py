#!/usr/bin/env python3
"""Summarise `nvcc -Xptxas -v` output: one line per function (registers, stack, spills)."""
import re, sys
lines = open(sys.argv[1]).read().splitlines()
for i, l in enumerate(lines):
    m = re.search(r"Function properties for (\S+)", l)
    if not m:
        continue
    name = re.sub(r"_INTERNAL_[0-9a-f_]+b2k_\w+_cu_[0-9a-f]+", "", m.group(1))
    props = lines[i + 1].strip() if i + 1 < len(lines) else ""
    used = lines[i + 2].strip() if i + 2 < len(lines) and "Used" in lines[i + 2] else ""
    st = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", props)
    rg = re.search(r"Used (\d+) registers", used)
    print(f"{name[:100]:100s} stack={st.group(1) if st else '?':>6} spill_st={st.group(2) if st else '?':>5} spill_ld={st.group(3) if st else '?':>5} regs={rg.group(1) if rg else '-'}")
