"""summarise a ptxas -v log: per entry kernel, registers / stack / spills (python scratch/ptxas_summary.py <log>)"""
import re, subprocess, sys
txt = open(sys.argv[1]).read()
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?Function properties for \1\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt, re.S):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"qmc::\(anonymous namespace\)::|void |\(.*", "", name)
    print("%-36s regs %3s stack %4s spill st/ld %4s/%4s" % (name, m.group(5), m.group(2), m.group(3), m.group(4)))
