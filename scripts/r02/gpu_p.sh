#!/bin/bash
# round 2, call P: issuer clock64 stamps of collect_hi_kernel at several places (debug build)
set +e
O=gpurun_out; mkdir -p $O
L=$PWD/audio-compression_b200/fwav_b200/libfwav_b200_dbg.so
for d in 0 256; do FWAV_LIB=$L FWAV_UMMA_DEBUG=$d timeout 200 python scripts/time_topk.py 1.0 umma 2 2>/dev/null | cut -c1-160; done
# (first stage / 64, CTA / 32): own neighbourhood, far away, a late CTA
for w in "0 0" "1 0" "60 0" "4 60" "60 60" "100 120"; do
  set -- $w
  d=$(( 64 + ($1 << 16) + ($2 << 24) ))
  FWAV_LIB=$L FWAV_UMMA_DEBUG=$d timeout 200 python scripts/time_topk.py 1.0 umma 1 > $O/r02p_t.out 2> $O/r02p_trace_$1_$2.err
  python - "$O/r02p_trace_$1_$2.err" "$1" "$2" <<'PY'
import sys
rows=[l.split() for l in open(sys.argv[1]) if l.strip() and l.split()[0].rstrip(':').isdigit() and ':' in l.split()[0]]
rows=[[int(r[0].rstrip(':'))]+[int(x) for x in r[1:]] for r in rows][-64:]
ev=[r for r in rows if r[0]%2==0 and r[4]!=-1]; od=[r for r in rows if r[0]%2==1 and r[4]!=-1]; rows=ev+od
def per(rs): 
    return (rs[-1][4]-rs[0][4])/(len(rs)-1) if len(rs)>1 else -1
print(f"stage0={int(sys.argv[2])*64} cta={int(sys.argv[3])*32}: cycles between issues, accumulator 0: {per(ev):.0f}, accumulator 1: {per(od):.0f}; free->issued {sum(r[4]-r[3] for r in rows)/len(rows):.0f}; issued->next free (same accumulator) {sum(b[3]-a[4] for a,b in zip(ev,ev[1:]))/max(1,len(ev)-1):.0f}")
PY
done
