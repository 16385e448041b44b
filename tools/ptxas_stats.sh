#!/bin/bash
# registers / spills / smem per kernel of engine.cu for a set of -D flags (no GPU needed): tools/ptxas_stats.sh [-DFLAG ...]
cd "$(dirname "$0")/.."
nvcc -ccbin /usr/bin/g++ -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -Xptxas -v "$@" -o /tmp/ptxas_stats.so ionization_b200/csrc/engine.cu 2>&1 | python3 -c "
import sys,re,subprocess
name=None; cur=None; rows={}
for line in sys.stdin:
    m=re.search(r\"Compiling entry function '(\S+)'\", line)
    if m: name=m.group(1); rows[name]=['','']; continue
    m=re.search(r\"Function properties for (\S+)\", line)
    if m: cur=m.group(1); continue
    if name and 'spill' in line and cur==name: rows[name][0]=line.strip().replace('bytes ','B ')
    if name and 'Used' in line: rows[name][1]=line.strip().replace('ptxas info    : ','')
names=list(rows)
dem=subprocess.run(['c++filt']+names,capture_output=True,text=True).stdout.splitlines()
for n,d in zip(names,dem):
    d=d.replace('ion::','').replace('(int)','').replace('(bool)','')
    d=re.sub(r'\(.*\)$','',d).replace('void ','')
    print(f'{d:44s} {rows[n][1][:60]:60s} {rows[n][0]}')
"
