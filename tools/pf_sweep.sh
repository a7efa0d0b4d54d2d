# start-up L2 prefetch per decoder lane (RCB_DEC_PF bytes; unset = sized by the host): decode time and DRAM bytes
for pf in auto 0 256 512 2048; do
  if [ $pf = auto ]; then unset RCB_DEC_PF; else export RCB_DEC_PF=$pf; fi
  for mode in static adaptive; do
    python bench.py --mode $mode --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('pf=$pf mode=$mode value=%.2f decode_ms=%.4f kernels=%s' % (d['value'], d['phase_ms']['decode'], d['roofline']['kernels_ms']))"
  done
done
for pf in auto 2048; do
  if [ $pf = auto ]; then unset RCB_DEC_PF; else export RCB_DEC_PF=$pf; fi
  for mode in static adaptive; do
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"decode" -s 4 -c 1 --csv python bench.py --mode $mode --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity 2>/dev/null | grep -E "dram__|gpu__time" | awk -F'","' -v t="pf=$pf mode=$mode" '{print t, $5, $(NF-2), $(NF-1), $NF}'
  done
done
