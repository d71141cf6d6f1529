#!/bin/bash
# usage: tools/prof_split.sh <tag>   -> /tmp/prof/<kernel>.csv per kernel + nvdisasm -gi of the current libbmpc.so
TAG=$1
mkdir -p /tmp/prof && cd /tmp/prof && rm -f *.cubin *.csv
cuobjdump -xelf all /root/repo/bipedal_control_b200/libbmpc.so >/dev/null
nvdisasm -gi -c bmpc_api.sm_100a.cubin > disgi.txt 2>/dev/null
python - <<PY
import re
src=open('/root/repo/gpurun_out/${TAG}_full_source.csv').read().split('"Kernel Name",')
seen=set()
for s in src[1:]:
    name=s.split('\n',1)[0]
    k=re.search(r'bmpc::(\w+)<',name).group(1)
    if k in seen: continue
    seen.add(k)
    open(f'/tmp/prof/{k}.csv','w').write('"Kernel Name",'+s)
print(sorted(seen))
PY
