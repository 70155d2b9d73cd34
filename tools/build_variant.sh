#!/bin/bash
# build a tuning variant of the library: build_variant.sh <name> [-DFLAG=... ...]  ->  exp/<name>.so (+ exp/<name>.ptxas)
NAME=$1; shift
mkdir -p exp
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off -shared \
  -Xptxas -v "$@" -o exp/$NAME.so fargocpt_b200/csrc/fargo_api.cu -ldl 2> exp/$NAME.ptxas
c++filt < exp/$NAME.ptxas | grep -A1 "Compiling entry function '.*\(k_fused_sources<true>\|k_fused_artvisc<true>\|k_fused_viscosity<true>\|k_transport_azimuthal<0, true>\|k_transport_radial<0, true>\|k_cfl\)" | grep -v "^--" | paste - - | sed -e "s/.*function '\(void \)\?\([a-z_]*<*[0-9a-z, ]*>*\)(.*Function properties for.*/\2/" | head -0
python3 - exp/$NAME.ptxas <<'PY'
import re,sys,subprocess
txt=subprocess.run(['c++filt'],input=open(sys.argv[1]).read(),capture_output=True,text=True).stdout
cur=None
for line in txt.splitlines():
    m=re.search(r"Compiling entry function '(?:void )?(\w+(?:<[^>]*>)?)",line)
    if m: cur=m.group(1)
    m=re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads",line)
    if m and cur: spill=m.groups()
    m=re.search(r"Used (\d+) registers",line)
    if m and cur and re.match(r"k_fused_\w+<true>|k_transport_\w+<0, true>|k_cfl",cur or ''):
        print(f"  {cur:40s} regs {m.group(1):>4s} stack {spill[0]} spill st/ld {spill[1]}/{spill[2]}")
PY
