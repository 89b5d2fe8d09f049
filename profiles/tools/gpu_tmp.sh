#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ab.log gpurun_out/dump_*.txt
TF_PROFILE_DUMP=gpurun_out/dump_half.txt python bench.py --frames 96 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-detection 2>&1 | python profiles/tools/ab_parse.py | head -2
TF_PYR_NO_HALF=1 TF_PROFILE_DUMP=gpurun_out/dump_nohalf.txt python bench.py --frames 96 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-detection 2>&1 | python profiles/tools/ab_parse.py | head -2
python - <<'PY'
for name in ("half","nohalf"):
    recs=[tuple(map(float,l.split())) for l in open(f"gpurun_out/dump_{name}.txt")]
    print(name, len(recs))
    prev_end=None
    gaps=[]
    for i,(k,t0,ms) in enumerate(recs):
        if prev_end is not None and t0-prev_end>0.3: gaps.append((i,int(k),round(t0,2),round(t0-prev_end,2), int(recs[i-1][0])))
        prev_end=t0+ms
    print(" gaps>0.3ms:", gaps[:30])
PY
