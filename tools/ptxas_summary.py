#!/usr/bin/env python
"""Summarise ptxas -v logs (registers / spills / smem per kernel)."""
import glob, re, subprocess, sys
for f in sorted(glob.glob('transflow_b200/csrc/*.ptxas.log')):
    txt = open(f).read()
    blocks = txt.split("Compiling entry function ")[1:]
    for b in blocks:
        name = re.match(r"'(\S+)'", b).group(1)
        dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r'\(.*', '', dem)[:60]
        st = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
        u = re.search(r"Used (\d+) registers", b)
        sm = re.search(r"(\d+) bytes smem", b)
        print(f"{dem:62s} regs={u.group(1):>3s} stack={st.group(1):>4s} spill={st.group(2)}/{st.group(3)} smem={sm.group(1) if sm else 0}")
