"""Prints registers / spills of the hot kernels from the ptxas log of the last libmag build."""
import re, subprocess, sys
t = open(sys.argv[1] if len(sys.argv) > 1 else 'core_b200/lib/mag_kernels.ptxas.log').read()
for b in re.split(r"ptxas info\s+: Compiling entry function '", t)[1:]:
    name = b.split("'")[0]
    dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    m = re.search(r"Used (\d+) registers", b); sp = re.search(r"(\d+) bytes spill stores", b)
    if any(k in dem for k in ('k_edges', 'k_tets', 'k_vertex', 'k_edge_rows', 'k_tet_rows')):
        print(dem.replace('(anonymous namespace)::', '')[:48], m.group(1), 'regs, spill', sp.group(1) if sp else '?')
