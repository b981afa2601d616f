#!/bin/bash
# A/B of CF_TC_FLAGS values on the pyramid build at 64 x 60x80 with the SM clock / power draw sampled during each run
# usage: power_ab.sh FLAGS [FLAGS ...]
for f in "$@"; do
  nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 50 > /tmp/smi_$f.txt &
  SMI=$!
  CF_TC_FLAGS=$f timeout 300 python scripts/scale_bench.py --only build --cases cfg5_480x640_b64 --out gpurun_out/sb_p.json 2>&1 | grep -E "^cfg|Error"
  kill $SMI
  python - "$f" <<'P'
import sys
rows = [l.strip().split(",") for l in open(f"/tmp/smi_{sys.argv[1]}.txt") if l.count(",") >= 2]
clk = sorted(float(r[0]) for r in rows); pw = [float(r[1]) for r in rows]; cap = sum("Active" in r[2] for r in rows)
load = clk[len(clk) // 2:]
print(f"   CF_TC_FLAGS={sys.argv[1]}: {len(rows)} samples, SM clock under load min {min(load):.0f} median {load[len(load)//2]:.0f} MHz, power max {max(pw):.0f} W, sw_power_cap active in {cap} samples")
P
done
