#!/bin/bash
for a in 0 768 896 640; do
  echo "== ABLATE=$a"
  ISDFB_ABLATE=$a python tools/kernel_timeline.py bf16x3 701 2>&1 | grep "timeline" | python -c "
import sys,re
for line in sys.stdin:
    pe=re.search(r'PE took (\d+)',line).group(1)
    steps=re.findall(r's(\d+) epi(\d+) wait_end=(-?\d+) epi_end=(-?\d+)',line)
    prev=0; out=[]
    for s,e,w,x in steps:
        w=int(w);x=int(x); out.append('s%s/e%s wait %d epi %d'%(s,e,w-prev,x-w)); prev=x
    print('PE',pe,'total',prev); print(' | '.join(out))
"
done
